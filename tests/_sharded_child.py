"""Child process (rank 1) of tests/test_multi_gpu.py::test_sharded_chain_two_processes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bio_b200 import _cabi as cabi, synth

device, n_total, chunk, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
handles = [bytes.fromhex(h) for h in sys.argv[5:10]]  # root's val, pos, off, status, state
dev = torch.device("cuda", device)
torch.cuda.set_device(dev)
ctx = cabi.Context(device)
n_tiles = (n_total + 31) // 32
my_handle, my_state = ctx.gather_create(n_tiles + 1)
print("HANDLE", my_handle.hex(), flush=True)
val, pos, off, status, state0 = (ctx.gather_open(h) for h in handles)
b, o = synth.uniform_reads(n_total, 150, 91)
b = b.copy()
b[150 * 40:150 * 40 + 30] = ord("N")  # the same patch as the parent's
# rank 1 of 2: the odd chunks, back to back
mine = [c for c in range((n_total + chunk - 1) // chunk) if c % 2 == 1]
segs = [b[c * chunk * 150:min(n_total, (c + 1) * chunk) * 150] for c in mine]
lb = np.concatenate(segs + [np.zeros(64, dtype=np.uint8)])
n = (len(lb) - 64) // 150
bases = torch.from_numpy(lb).to(dev)
loff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 150
flags = torch.zeros(1, dtype=torch.int32, device=dev)
if mode == "syncmer":
    p = cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=150, pos_width=1)
else:
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150, pos_width=1)
assert sys.stdin.readline().strip() == "GO"
for epoch in (1, 2):
    spec = cabi.ShardSpec()
    spec.rank, spec.n_ranks, spec.chunk_reads, spec.epoch, spec.n_reads_global = 1, 2, chunk, epoch, n_total
    spec.state[0], spec.state[1] = state0, my_state
    ctx.enqueue_device_sharded(p, spec, bases, loff, n * 150, val, pos, off, status, int(sys.argv[10]), flags)
    torch.cuda.synchronize()
    assert int(flags.item()) == 0
    print("DONE", epoch, flush=True)
    assert sys.stdin.readline().strip() == "NEXT"
for a in (val, pos, off, status, state0):
    ctx.gather_close(a, False)
ctx.gather_close(my_state, True)
ctx.close()
