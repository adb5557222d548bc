"""Deterministic synthetic FASTA/FASTQ of the BASELINE.json shapes (bench + tests tooling).

Record lengths come from ``numpy.random.default_rng(seed)``; every base/quality byte is a
counter-based hash of (seed, record, position), so the CUDA generator
(``csrc/fxg_synth.cu``, used by bench.py to fill HBM directly) and this numpy generator
produce byte-identical files.

Shapes:
  C2  FASTA  ``>seq{i} synthetic len={L}``, L ~ U[9000, 11000], 80 bases per line, LF
  C4  FASTQ  ``@read{i} 1:N:0:ACGT``, 150 bp, ``+``, qualities U[35, 70]
"""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
MULK = np.uint64(0xBF58476D1CE4E5B9)
QSALT = np.uint64(0xD6E8FEB86659FD93)
_M1 = np.uint64(0xFF51AFD7ED558CCD)
_M2 = np.uint64(0xC4CEB9FE1A85EC53)
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mix(x):
    """murmur3 fmix64 on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(33))
        x = x * _M1
        x = x ^ (x >> np.uint64(33))
        x = x * _M2
        x = x ^ (x >> np.uint64(33))
    return x


def _key(seed, rec, k):
    with np.errstate(over="ignore"):
        return np.uint64(seed) + np.uint64(rec) * GOLD + k.astype(np.uint64) * MULK


def bases(seed, rec, length):
    k = np.arange(length, dtype=np.uint64)
    return _BASES[(_mix(_key(seed, rec, k)) >> np.uint64(62)).astype(np.int64)]


def quals(seed, rec, length):
    k = np.arange(length, dtype=np.uint64)
    z = _mix(_key(seed, rec, k) ^ QSALT)
    return (np.uint64(35) + (z >> np.uint64(32)) % np.uint64(36)).astype(np.uint8)


def fasta_lengths(n_records, seed, min_len=9000, max_len=11000):
    rng = np.random.default_rng(seed)
    return rng.integers(min_len, max_len + 1, size=n_records, dtype=np.int64)


def fasta_header(i, length):
    return b">seq%d synthetic len=%d" % (i + 1, length)


def fasta_record_sizes(lengths, width=80, crlf=False):
    """Bytes per record (header + wrapped sequence) -- used for the HBM layout."""
    eol = 2 if crlf else 1
    idx = np.arange(1, lengths.size + 1, dtype=np.int64)
    ndig = lambda v: np.floor(np.log10(np.maximum(v, 1))).astype(np.int64) + 1
    hdr = len(b">seq synthetic len=") + ndig(idx) + ndig(lengths) + eol
    nlines = (lengths + width - 1) // width
    return hdr + lengths + nlines * eol


def synth_fasta(n_records, seed=20240601, min_len=9000, max_len=11000, width=80, crlf=False,
                trailing_newline=True):
    """Return the file as a bytes object (CPU path; fine up to a few hundred MB)."""
    lengths = fasta_lengths(n_records, seed, min_len, max_len)
    eol = b"\r\n" if crlf else b"\n"
    parts = []
    for i, L in enumerate(lengths):
        L = int(L)
        parts.append(fasta_header(i, L) + eol)
        b = bases(seed, i, L)
        nfull = L // width
        if nfull:
            body = b[:nfull * width].reshape(nfull, width)
            wrapped = np.concatenate([body, np.broadcast_to(np.frombuffer(eol, np.uint8), (nfull, len(eol)))], axis=1)
            parts.append(wrapped.tobytes())
        if L % width:
            parts.append(b[nfull * width:].tobytes() + eol)
    data = b"".join(parts)
    if not trailing_newline and data.endswith(eol):
        data = data[:-len(eol)]
    return data


def fastq_header(i):
    return b"@read%d 1:N:0:ACGT" % (i + 1)


def synth_fastq(n_reads, seed=20240602, read_len=150, crlf=False):
    eol = b"\r\n" if crlf else b"\n"
    parts = []
    for i in range(n_reads):
        parts.append(fastq_header(i) + eol)
        parts.append(bases(seed, i, read_len).tobytes() + eol)
        parts.append(b"+" + eol)
        parts.append(quals(seed, i, read_len).tobytes() + eol)
    return b"".join(parts)


def random_queries(slens, n_queries, seed=123, window=1000, minus_prob=0.5, mixed=False):
    """C3 query set: uniform record, fixed ``window`` (or L ~ U[50, 5000] when ``mixed``),
    0-based half-open [s, e), strand '-' with probability ``minus_prob``.
    Returns (row_id, s, e, minus) int64/bool arrays."""
    rng = np.random.default_rng(seed)
    slens = np.asarray(slens, dtype=np.int64)
    rid = rng.integers(0, slens.size, size=n_queries, dtype=np.int64)
    if mixed:
        L = rng.integers(50, 5001, size=n_queries, dtype=np.int64)
    else:
        L = np.full(n_queries, window, dtype=np.int64)
    L = np.minimum(L, slens[rid])
    u = rng.random(n_queries)
    s = np.floor(u * (slens[rid] - L + 1)).astype(np.int64)
    e = s + L
    minus = rng.random(n_queries) < minus_prob
    return rid, s, e, minus
