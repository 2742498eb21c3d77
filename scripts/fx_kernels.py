"""Per-kernel view of the record feeder (run under ncu --metrics gpu__time_duration.sum): one parse of a FASTQ text
(4M x 150 bp) and one of a 60-column FASTA text (100k x 9960 bp), after a warm-up parse each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bio_b200 import _cabi as cabi
from bench_fastx import fastq_text, fasta_text
ctx = cabi.Context(0)
for text in (fastq_text(4_000_000, 150, 46), fasta_text(100_000, 9960, 60, 47)):
    nb = text.size
    pad = np.zeros((nb + 15) // 16 * 16 + 16, dtype=np.uint8); pad[:nb] = text
    d = torch.from_numpy(pad).cuda()
    for _ in range(2):
        info = ctx.fastx_parse_device(d, nb, 0, True)
    torch.cuda.synchronize()
    print(nb, info.n_records, info.n_lines)
