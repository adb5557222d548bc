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
