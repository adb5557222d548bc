"""pyfastx_b200 -- B200-native implementation of the pyfastx hot path.

    import pyfastx_b200 as pyfastx
    fa = pyfastx.Fasta("genome.fa")            # GPU index-build scan, reference-compatible .fxi
    fa["chr1"][1000:2000].antisense            # GPU extraction (strip + reverse complement)
    fa.fetch_many(names, starts, ends, strands)  # batched: millions of queries per call

The native code lives in libfxg.so (C-ABI: include/fxg.h); there is no CPU fallback for the hot path.
The object layer is served by the compiled CPython extension `pyfastx_b200.pyfastx` (api.py compiled by Cython,
exports PyInit_pyfastx like the reference's src/module.c) when it has been built (`__graft_entry__.build()` /
csrc/build_ext.sh); the identical Python source `api.py` is the fallback for a tree without a C compiler.
"""
try:
    from . import pyfastx as _impl          # compiled object layer
    COMPILED = True
except ImportError:                          # pragma: no cover - source tree without the built extension
    from . import api as _impl
    COMPILED = False

Fasta, Fastq, Fastx = _impl.Fasta, _impl.Fastq, _impl.Fastx
Sequence, Read = _impl.Sequence, _impl.Read
FastaKeys, FastqKeys = _impl.FastaKeys, _impl.FastqKeys
gzip_check, reverse_complement, version = _impl.gzip_check, _impl.reverse_complement, _impl.version

__version__ = version()
