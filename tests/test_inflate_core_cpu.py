"""The per-member DEFLATE decoder that the inflate kernel runs one thread per BGZF member
(pyfastx_b200/csrc/fxg_inflate_core.cuh), compiled for the host and checked against zlib: dynamic, fixed and
stored blocks, long codes (slow path), overlapping matches, corrupt and truncated members."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

import gen
from pyfastx_b200 import synth
from test_bgzf import bgzf_compress, members

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("inflate_core") / "inflate_core_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "native", "inflate_core_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.fxi_host_inflate.restype = C.c_int
    lib.fxi_host_inflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    return lib


def inflate(core, z):
    rc, co, uo, tot = members(z)
    assert rc == 0
    a = np.frombuffer(z, dtype=np.uint8).copy()
    out = np.zeros(tot + 64, dtype=np.uint8)
    st = np.zeros(len(co) - 1, dtype=np.int32)
    bad = core.fxi_host_inflate(a.ctypes.data, a.size, co.ctypes.data, uo.ctypes.data, len(co) - 1, out.ctypes.data, tot, st.ctypes.data)
    return bad, out[:tot].tobytes(), st


@pytest.mark.parametrize("kind", ["dna", "text", "stored", "tiny", "fastq", "binary", "runs", "huff_only", "long_codes"])
def test_matches_zlib(core, kind):
    rng = np.random.default_rng(7)
    level, block, strategy = 6, 0xff00, zlib.Z_DEFAULT_STRATEGY
    if kind == "dna":
        data = synth.synth_fasta(120, seed=11)
    elif kind == "text":
        data, level = (b"the quick brown fox jumps over the lazy dog. " * 20000)[:700_000], 9
    elif kind == "stored":
        data, level = rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes(), 0
    elif kind == "tiny":
        data, block = synth.synth_fastq(60, seed=2), 97                   # many tiny members (fixed Huffman blocks)
    elif kind == "fastq":
        data, level = synth.synth_fastq(6000, seed=20240602), 1
    elif kind == "binary":
        data, block = rng.integers(0, 256, size=300_000, dtype=np.uint8).tobytes(), 30000
    elif kind == "runs":
        data = b"A" * 100_000 + b"ACGT" * 30_000 + bytes(range(256)) * 300    # distance-1 and short-period overlaps
    elif kind == "huff_only":
        data, strategy = gen.random_fasta(3, n_records=200), zlib.Z_HUFFMAN_ONLY
    else:
        # a skewed alphabet gives codes longer than the 9-bit primary table
        p = np.array([2.0 ** -i for i in range(1, 41)]); p /= p.sum()
        data = rng.choice(np.arange(40, 80, dtype=np.uint8), size=400_000, p=p).tobytes()
    if strategy == zlib.Z_DEFAULT_STRATEGY:
        z = bgzf_compress(data, level, block)
    else:
        import struct
        parts = []
        for a in list(range(0, len(data), block)) + [None]:
            chunk = b"" if a is None else data[a:a + block]
            co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, strategy)
            comp = co.compress(chunk) + co.flush()
            parts.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25)
                         + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        z = b"".join(parts)
    bad, out, st = inflate(core, z)
    assert bad == 0 and not st.any()
    assert out == data


def test_corrupt_members_are_reported(core):
    data = synth.synth_fasta(50, seed=5)
    z = bytearray(bgzf_compress(data))
    z[200] ^= 0xff                                   # inside the first member's deflate data
    bad, out, st = inflate(core, bytes(z))
    assert bad >= 1 and st[0] != 0 and not st[1:].any()
    assert out[0xff00:] == data[0xff00:]             # the other members are unaffected


# ---- segments: decoding from zran checkpoints (generic gzip) --------------------------------------------
@pytest.fixture(scope="module")
def core_points(core):
    core.fxi_host_inflate_points.restype = C.c_int
    core.fxi_host_inflate_points.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    return core


@pytest.mark.parametrize("kind,level,spacing", [("dna", 6, 1 << 16), ("dna", 1, 1 << 15), ("fastq", 9, 1 << 17),
                                                ("text", 6, 1 << 16), ("stored", 0, 1 << 16), ("runs", 6, 1 << 15)])
def test_segments_from_checkpoints_match_zlib(core_points, kind, level, spacing):
    """every checkpoint that the one sequential pass (csrc/fxg_gzip.cpp) collects starts an independent segment: the decoder
    that the GPU runs one thread per segment, started at (compressed offset, bit offset) with the checkpoint's 32 KiB
    window, reproduces the bytes up to the next checkpoint -- all segments together give the file"""
    import gzip as _gzip
    from test_gzip_cpu import inflate_host
    from pyfastx_b200 import _cabi
    rng = np.random.default_rng(3)
    if kind == "dna":
        data = synth.synth_fasta(400, seed=21)
    elif kind == "fastq":
        data = synth.synth_fastq(20000, seed=22)
    elif kind == "text":
        data = (b"the quick brown fox jumps over the lazy dog. " * 40000)[:1_500_000]
    elif kind == "stored":
        data = rng.integers(0, 256, size=600_000, dtype=np.uint8).tobytes()
    else:
        data = b"A" * 300_000 + b"ACGT" * 100_000 + bytes(range(256)) * 1000
    z = _gzip.compress(data, compresslevel=level)
    got, gz, pts, h = inflate_host(z, spacing)
    try:
        assert got == data
        n = len(pts["cmp"])
        assert n >= 1 and (n >= 3 or kind not in ("dna", "fastq"))
        a = np.frombuffer(z, dtype=np.uint8).copy()
        ucmp = np.concatenate([pts["ucmp"], [len(data)]]).astype(np.int64)
        out = np.zeros(len(data) + 64, dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        win = np.frombuffer(pts["win"], dtype=np.uint8).copy() if pts["win"] else np.zeros(1, np.uint8)
        bad = core_points.fxi_host_inflate_points(a.ctypes.data, a.size, n, pts["cmp"].ctypes.data, pts["bits"].ctypes.data,
                                                  ucmp.ctypes.data, pts["has"].ctypes.data, win.ctypes.data, 32768,
                                                  out.ctypes.data, len(data), st.ctypes.data)
        assert bad == 0 and not st.any(), st[:8]
        assert out[:len(data)].tobytes() == data
    finally:
        _cabi.lib().fxg_gzip_free(h)
