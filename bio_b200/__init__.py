"""bio_b200: B200-native sketching path (k-mer / ntHash / minimizer / syncmer / protein k-mer)
behind the sketches.Iterator / sketches.Sketch interface of shenwei356/bio.

The compute lives in bio_b200/lib/libb200sketch.so (hand-written sm_100a CUDA behind the C ABI of
include/b200sketch.h).  There is no CPU fallback.
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]
