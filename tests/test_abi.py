"""The C-ABI library: loads, exports every symbol include/b200sketch.h declares, and its
device-free entry points (constructor checks, error strings, bounds) behave like the reference's
constructors.  No compute calls here -- this file runs without a GPU."""
import ctypes as C
import os
import re

import pytest

from bio_b200 import _cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    L = cabi.lib()
    assert L.b200sk_version() >= 100


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "b200sketch.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200sk_[a-z_]+)\s*\(", hdr))
    assert declared == set(cabi.EXPORTS)
    L = cabi.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_product_does_not_link_oracle():
    # the product library must not reference the oracle (no CPU fallback)
    import subprocess
    out = subprocess.run(["nm", "-D", cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "ora_" not in out
    for f in os.listdir(os.path.join(ROOT, "bio_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "bio_b200", f)).read()
            assert "import oracle" not in src and "from oracle" not in src


def test_constructor_checks():
    # iterator.go:616,669; sketch.go:86-91,143-148; iterator-protein.go:47; codon_tables.go:209
    L = cabi.lib()
    ck = lambda **kw: L.b200sk_check_params(C.byref(cabi.make_params(**kw)))
    assert ck(mode=cabi.MODE_NTHASH, k=21) == 0
    assert ck(mode=cabi.MODE_NTHASH, k=0) == cabi.ERR_INVALID_K
    assert ck(mode=cabi.MODE_KMER, k=33) == cabi.ERR_K_OVERFLOW
    assert ck(mode=cabi.MODE_KMER, k=32) == 0
    assert ck(mode=cabi.MODE_MINIMIZER, k=21, w=0) == cabi.ERR_INVALID_W
    assert ck(mode=cabi.MODE_MINIMIZER, k=0, w=0) == cabi.ERR_INVALID_K
    assert ck(mode=cabi.MODE_MINIMIZER, k=21, w=1) == 0
    assert ck(mode=cabi.MODE_SYNCMER, k=21, s=22) == cabi.ERR_INVALID_S
    assert ck(mode=cabi.MODE_SYNCMER, k=21, s=0) == cabi.ERR_INVALID_S
    assert ck(mode=cabi.MODE_SYNCMER, k=21, s=21) == 0
    assert ck(mode=cabi.MODE_PROTEIN, k=11, frame=0) == cabi.ERR_INVALID_FRAME
    assert ck(mode=cabi.MODE_PROTEIN, k=11, frame=-3) == 0
    assert ck(mode=9, k=11) == cabi.ERR_BAD_ARG


def test_error_strings_are_the_references():
    # iterator.go:34-53, sketch.go:32-42
    assert cabi.strerror(cabi.ERR_INVALID_K) == "sketches: invalid k-mer size"
    assert cabi.strerror(cabi.ERR_SHORT_SEQ) == "sketches: sequence too short"
    assert cabi.strerror(cabi.ERR_INVALID_W) == "kmers: invalid minimimzer window"
    assert cabi.strerror(cabi.ERR_INVALID_S) == "kmers: invalid s-mer size"
    assert cabi.strerror(cabi.ERR_ILLEGAL_BASE) == "sketches: illegal base"


def test_output_bound():
    L = cabi.lib()
    p = cabi.make_params(cabi.MODE_NTHASH, 21)
    assert L.b200sk_output_bound(C.byref(p), 1500, 10, 1) >= 1300
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11)
    assert L.b200sk_output_bound(C.byref(p), 1500, 10, 1) >= 1200
    assert L.b200sk_output_bound(C.byref(p), 150000, 1000, 0) >= 22200


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cabi.SketchError) as e:
        cabi.Context(0)
    assert e.value.code == cabi.ERR_NO_DEVICE
    # the pipelined reader refuses the same way (and bad parameters come back before any device is touched)
    with pytest.raises(cabi.SketchError) as e:
        cabi.FastxStream(cabi.make_params(cabi.MODE_MINIMIZER, k=21, w=11), b"@r\nACGT\n+\nIIII\n")
    assert e.value.code == cabi.ERR_NO_DEVICE
    with pytest.raises(cabi.SketchError) as e:
        cabi.FastxStream(cabi.make_params(cabi.MODE_MINIMIZER, k=21, w=0), b"@r\nACGT\n+\nIIII\n")
    assert e.value.code == -3  # ErrInvalidW, sketches/sketch.go:89
