# cython: language_level=3, boundscheck=False, wraparound=False
"""_fast -- compiled bridge between the object layer and the C-ABI for the per-object getters.

The reference's getters (Sequence.seq / .antisense ..., Read.seq / .qual; src/sequence.c:337-398, src/read.c:152-249)
are C functions inside a CPython extension; here the equivalent calls go from compiled code straight into libfxg.so
(no ctypes marshalling), with the GIL released while the GPU works, and write into the bytes object that is returned.
"""
from cpython.bytes cimport PyBytes_FromStringAndSize, PyBytes_AS_STRING
from libc.stdint cimport int64_t, int32_t, uint8_t, uintptr_t

cdef extern from "fxg.h":
    ctypedef struct fxg_ctx
    ctypedef struct fxg_file
    ctypedef struct fxg_fasta_row
    ctypedef struct fxg_fastq_row
    ctypedef struct fxg_nametab
    const char *fxg_last_error() nogil
    int fxg_extract_one_host(fxg_ctx *ctx, const fxg_file *f, const fxg_fasta_row *d_rows, int64_t n_rows,
                             int64_t row_id, int64_t s, int64_t e, int32_t flags, uint8_t *out_host, int64_t out_cap) nogil
    int fxg_read_one_host(fxg_ctx *ctx, const fxg_file *f, const fxg_fastq_row *d_rows, int64_t n_rows, int64_t read_id,
                          int which, int32_t flags, int64_t rlen, uint8_t *out_host, int64_t out_cap) nogil
    int64_t fxg_nametab_find(const fxg_nametab *t, const uint8_t *name, int64_t len) nogil


class FxgError(RuntimeError):
    pass


cdef inline object _fail(int rc):
    raise FxgError("libfxg error %d: %s" % (rc, fxg_last_error().decode("utf-8", "replace")))


def extract_one(uintptr_t ctx, uintptr_t dfile, uintptr_t drows, int64_t n_rows, int64_t row_id, int64_t s, int64_t e,
                int32_t flags):
    """bytes of ONE query [s, e) of record row_id: one kernel launch, one synchronisation"""
    cdef int64_t n = e - s
    if n <= 0:
        return b""
    cdef object out = PyBytes_FromStringAndSize(NULL, n)
    cdef uint8_t *p = <uint8_t *>PyBytes_AS_STRING(out)
    cdef int rc
    with nogil:
        rc = fxg_extract_one_host(<fxg_ctx *>ctx, <const fxg_file *>dfile, <const fxg_fasta_row *>drows, n_rows,
                                  row_id, s, e, flags, p, n)
    if rc != 0:
        _fail(rc)
    return out


def read_one(uintptr_t ctx, uintptr_t dfile, uintptr_t drows, int64_t n_rows, int64_t read_id, int which, int32_t flags,
             int64_t rlen):
    """sequence (which = 0) or quality (which = 1) bytes of ONE read"""
    if rlen <= 0:
        return b""
    cdef object out = PyBytes_FromStringAndSize(NULL, rlen)
    cdef uint8_t *p = <uint8_t *>PyBytes_AS_STRING(out)
    cdef int rc
    with nogil:
        rc = fxg_read_one_host(<fxg_ctx *>ctx, <const fxg_file *>dfile, <const fxg_fastq_row *>drows, n_rows, read_id,
                               which, flags, rlen, p, rlen)
    if rc != 0:
        _fail(rc)
    return out


def name_find(uintptr_t table, bytes name):
    """row of `name` in a fxg_nametab, -1 if absent"""
    cdef const uint8_t *p = <const uint8_t *>PyBytes_AS_STRING(name)
    cdef int64_t n = len(name)
    cdef int64_t r
    with nogil:
        r = fxg_nametab_find(<const fxg_nametab *>table, p, n)
    return r
