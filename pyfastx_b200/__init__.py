"""pyfastx_b200 -- B200-native implementation of the pyfastx hot path.

    import pyfastx_b200 as pyfastx
    fa = pyfastx.Fasta("genome.fa")            # GPU index-build scan, reference-compatible .fxi
    fa["chr1"][1000:2000].antisense            # GPU extraction (strip + reverse complement)
    fa.fetch_many(names, starts, ends, strands)  # batched: millions of queries per call

The native code lives in libfxg.so (C-ABI: include/fxg.h); there is no CPU fallback.
"""
from .api import Fasta, Fastq, Read, Sequence, gzip_check, reverse_complement, version  # noqa: F401

__version__ = version()
