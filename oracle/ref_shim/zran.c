/* Naive stand-in for indexed_gzip's zran.c (see zran.h in this directory).
 * Random access = gzseek + gzread on a private gzFile: correct, O(file) per backward seek.
 * TEST INFRASTRUCTURE ONLY; never linked into the product. */
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/stat.h>
#include <zlib.h>
#include "zran.h"

const char    ZRAN_INDEX_FILE_ID[5]   = {'G', 'Z', 'I', 'D', 'X'};
const uint8_t ZRAN_INDEX_FILE_VERSION = 1;

int zran_init(zran_index_t *index, FILE *fd, PyObject *f, uint32_t spacing, uint32_t window_size,
              uint32_t readbuf_size, uint16_t flags) {
    struct stat st;
    memset(index, 0, sizeof(*index));
    index->fd = fd;
    index->f = f;
    index->spacing = spacing ? spacing : 1048576;
    index->window_size = window_size ? window_size : 32768;
    index->readbuf_size = readbuf_size ? readbuf_size : 16384;
    index->flags = flags;
    if (fd && fstat(fileno(fd), &st) == 0) index->compressed_size = (uint64_t)st.st_size;
    index->size = 8;
    index->list = (zran_point_t *)calloc(index->size, sizeof(zran_point_t));
    if (fd) {
        int d = dup(fileno(fd));
        lseek(d, 0, SEEK_SET);
        index->gz = (void *)gzdopen(d, "rb");
        if (index->gz) gzbuffer((gzFile)index->gz, 1 << 20);
    }
    return index->list ? 0 : -1;
}

void zran_free(zran_index_t *index) {
    uint32_t i;
    if (!index) return;
    for (i = 0; i < index->npoints; ++i) free(index->list[i].data);
    free(index->list);
    if (index->gz) gzclose((gzFile)index->gz);
    index->list = NULL;
    index->gz = NULL;
}

/* One pass over the stream to learn the uncompressed size; records a single point at
 * offset 0 (no window) so the exported blob passes the reference importer's checks. */
int zran_build_index(zran_index_t *index, uint64_t from, uint64_t until) {
    static char buf[1 << 20];
    gzFile g = (gzFile)index->gz;
    int n;
    uint64_t total = 0;
    (void)from; (void)until;
    if (!g) return ZRAN_BUILD_INDEX_FAIL;
    gzrewind(g);
    while ((n = gzread(g, buf, sizeof(buf))) > 0) total += (uint64_t)n;
    gzrewind(g);
    index->uncompressed_size = total;
    index->npoints = 1;
    index->list[0].cmp_offset = 0;
    index->list[0].uncmp_offset = 0;
    index->list[0].bits = 0;
    index->list[0].data = NULL;
    return ZRAN_BUILD_INDEX_OK;
}

int zran_seek(zran_index_t *index, int64_t offset, uint8_t whence, zran_point_t **point) {
    gzFile g = (gzFile)index->gz;
    if (point) *point = NULL;
    if (!g) return ZRAN_SEEK_FAIL;
    if (whence == SEEK_CUR) offset += (int64_t)index->uncmp_seek_offset;
    if (offset < 0) return ZRAN_SEEK_FAIL;
    if (gzseek(g, (z_off_t)offset, SEEK_SET) < 0) return ZRAN_SEEK_FAIL;
    index->uncmp_seek_offset = (uint64_t)offset;
    return ZRAN_SEEK_OK;
}

int64_t zran_read(zran_index_t *index, void *buf, uint64_t len) {
    gzFile g = (gzFile)index->gz;
    uint64_t done = 0;
    if (!g) return ZRAN_READ_FAIL;
    while (done < len) {
        unsigned want = (len - done) > (1u << 30) ? (1u << 30) : (unsigned)(len - done);
        int n = gzread(g, (char *)buf + done, want);
        if (n < 0) return ZRAN_READ_FAIL;
        if (n == 0) break;
        done += (uint64_t)n;
    }
    index->uncmp_seek_offset += done;
    if (done == 0 && len > 0) return ZRAN_READ_EOF;
    return (int64_t)done;
}
