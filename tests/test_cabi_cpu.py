"""CPU-only checks of the product boundary: the C-ABI library builds/loads, exports every symbol
declared in include/fxg.h, fails loudly without a GPU, and the host-side helpers behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pyfastx_b200 import _cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    declared = _cabi.declared_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(_cabi.SIGNATURES), set(declared) ^ set(_cabi.SIGNATURES)
    assert lib.fxg_abi_version() == 2


def test_row_struct_layouts_match_header():
    text = open(os.path.join(ROOT, "include", "fxg.h")).read()
    assert "} fxg_fasta_row;" in text and "} fxg_fastq_row;" in text
    assert _cabi.FASTA_ROW.itemsize == 48 and _cabi.FASTQ_ROW.itemsize == 32
    assert _cabi.FASTA_ROW.fields["dlen"][1] == 32 and _cabi.FASTA_ROW.fields["elen"][1] == 40
    assert _cabi.FASTQ_ROW.fields["dlen"][1] == 24
    assert C.sizeof(_cabi.ScanStats) == 64


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "fxg.h")).read()
    for needle in ("src/index.c:226-361", "src/fastq.c:84-171", "src/sequence.c:498-510", "src/util.c:166-194",
                   "src/read.c:37-45"):
        assert needle in text


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_fails_loudly():
    lib = _cabi.lib()
    assert lib.fxg_device_count() == 0
    h = C.c_void_p()
    rc = lib.fxg_ctx_create(0, C.byref(h))
    assert rc == _cabi.FXG_ENODEV
    assert b"no CPU fallback" in lib.fxg_last_error()
    from pyfastx_b200 import engine
    with pytest.raises(_cabi.NoDeviceError):
        engine.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyfastx_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "fxo_" not in text and "libfxo" not in text, fn


def test_synth_is_deterministic_and_well_formed():
    a = synth.synth_fasta(30, seed=5)
    assert a == synth.synth_fasta(30, seed=5) and a != synth.synth_fasta(30, seed=6)
    lens = synth.fasta_lengths(30, 5)
    assert int(synth.fasta_record_sizes(lens).sum()) == len(a)
    assert a.startswith(b">seq1 synthetic len=%d\n" % lens[0])
    q = synth.synth_fastq(10, seed=3)
    lines = q.split(b"\n")
    assert lines[0] == b"@read1 1:N:0:ACGT" and len(lines[1]) == 150 and lines[2] == b"+" and len(lines[3]) == 150
    assert min(lines[3]) >= 35 and max(lines[3]) <= 70
    rid, s, e, minus = synth.random_queries(lens, 1000, seed=123)
    assert ((e - s) == 1000).all() and (s >= 0).all() and (e <= lens[rid]).all()


def test_compiled_object_layer_is_a_cpython_extension():
    """the object layer is a compiled extension module exporting PyInit_pyfastx with the reference's type names
    (src/module.c:61-138); it must import without a GPU (compute calls then fail with NoDeviceError)"""
    import ctypes
    import importlib
    import pyfastx_b200
    mod = importlib.import_module("pyfastx_b200.pyfastx")
    assert mod.__file__.endswith(".so") and pyfastx_b200.COMPILED
    lib = ctypes.CDLL(mod.__file__)
    assert hasattr(lib, "PyInit_pyfastx")
    for name in ("Fasta", "Fastq", "Fastx", "Sequence", "Read", "FastaKeys", "FastqKeys", "version", "gzip_check",
                 "reverse_complement"):
        assert hasattr(mod, name) and getattr(pyfastx_b200, name) is getattr(mod, name)
    fast = importlib.import_module("pyfastx_b200._fast")
    assert fast.__file__.endswith(".so") and all(hasattr(fast, n) for n in ("extract_one", "read_one", "name_find"))
