"""Child process of tests/test_multi_gpu.py::test_ipc_gather_two_processes: maps the root's gather buffer over CUDA IPC
and lets the sketching kernel store its shard's minimizers straight into its segment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bio_b200 import _cabi as cabi, synth

handle = bytes.fromhex(sys.argv[1])
seg_base, cap, device, n_total, r0, r1 = (int(x) for x in sys.argv[2:8])
dev = torch.device("cuda", device)
torch.cuda.set_device(dev)
ctx = cabi.Context(device)
b, o = synth.uniform_reads(n_total, 150, 77)
bases = torch.from_numpy(np.concatenate([b[r0 * 150:r1 * 150], np.zeros(64, dtype=np.uint8)])).to(dev)
off = torch.from_numpy((o[r0:r1 + 1] - o[r0]).astype(np.int64)).to(dev)
n = r1 - r0
p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150)
gaddr = ctx.gather_open(handle)
pos = torch.empty(cap, dtype=torch.int32, device=dev)
ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
st = torch.empty(n, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
ctx.enqueue_device_raw(p, bases, off, n * 150, gaddr + seg_base * 8, cap, pos, ooff, st, flags)
torch.cuda.synchronize()
assert int(flags.item()) == 0
print("COUNT", int(ooff[n].item()), flush=True)
ctx.gather_close(gaddr, False)
ctx.close()
