/* Stand-in for indexed_gzip 1.10.3 zran.h (pinned by the reference's setup.py, but NOT
 * vendored under /root/reference and not fetchable here).  TEST INFRASTRUCTURE ONLY.
 * It lets the unmodified reference sources compile into oracle/_ref/.  Plain-file paths
 * never enter zran; for .gz inputs the reference's index scan runs on zlib's gzread, so
 * every seq/read/stat row and every extracted byte is still the genuine reference
 * result.  Only the gzindex checkpoint blobs come from this stand-in ("parity unpinned"
 * for blob contents). */
#ifndef FXO_ZRAN_SHIM_H
#define FXO_ZRAN_SHIM_H
#include <stdint.h>
#include <stdio.h>
#include <Python.h>

typedef struct zran_point {
    uint64_t cmp_offset;
    uint64_t uncmp_offset;
    uint8_t  bits;
    uint8_t *data;
} zran_point_t;

typedef struct zran_index {
    FILE     *fd;
    PyObject *f;
    uint64_t  compressed_size;
    uint64_t  uncompressed_size;
    uint32_t  spacing;
    uint32_t  window_size;
    uint32_t  log_window_size;
    uint32_t  readbuf_size;
    uint32_t  npoints;
    uint32_t  size;
    zran_point_t *list;
    uint64_t  uncmp_seek_offset;
    uint16_t  flags;
    /* stand-in private state */
    void     *gz;          /* gzFile over a dup of fd */
    uint64_t  gz_pos;
} zran_index_t;

enum { ZRAN_AUTO_BUILD = 1, ZRAN_SKIP_CRC_CHECK = 2 };
enum { ZRAN_BUILD_INDEX_OK = 0, ZRAN_BUILD_INDEX_FAIL = -1 };
enum { ZRAN_SEEK_FAIL = -1, ZRAN_SEEK_OK = 0, ZRAN_SEEK_NOT_COVERED = 1, ZRAN_SEEK_EOF = 2 };
enum { ZRAN_READ_NOT_COVERED = -1, ZRAN_READ_EOF = -2, ZRAN_READ_FAIL = -3 };
enum { ZRAN_EXPORT_OK = 0, ZRAN_EXPORT_WRITE_ERROR = -1 };
enum { ZRAN_IMPORT_OK = 0, ZRAN_IMPORT_FAIL = -1, ZRAN_IMPORT_EOF = -2, ZRAN_IMPORT_READ_ERROR = -3,
       ZRAN_IMPORT_INCONSISTENT = -4, ZRAN_IMPORT_MEMORY_ERROR = -5, ZRAN_IMPORT_UNKNOWN_FORMAT = -6,
       ZRAN_IMPORT_UNSUPPORTED_VERSION = -7 };

extern const char    ZRAN_INDEX_FILE_ID[5];
extern const uint8_t ZRAN_INDEX_FILE_VERSION;

int     zran_init(zran_index_t *index, FILE *fd, PyObject *f, uint32_t spacing, uint32_t window_size,
                  uint32_t readbuf_size, uint16_t flags);
void    zran_free(zran_index_t *index);
int     zran_build_index(zran_index_t *index, uint64_t from, uint64_t until);
int     zran_seek(zran_index_t *index, int64_t offset, uint8_t whence, zran_point_t **point);
int64_t zran_read(zran_index_t *index, void *buf, uint64_t len);
#endif
