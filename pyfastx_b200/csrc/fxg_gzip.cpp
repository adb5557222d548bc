// fxg_gzip.cpp -- generic (non-BGZF) gzip inputs, host side of libfxg.so.
//
// A plain .gz file is ONE serial deflate stream (or a few concatenated members): there are no independent entry
// points, so -- exactly like the reference, whose scan reads it through zlib's gzread (src/kseq.c:70) and whose
// zran_build_index (src/index.c:381-387, indexed_gzip) then inflates it a SECOND time to collect checkpoints -- the
// bytes have to go through a sequential inflate once.  This does that one pass with zlib (Z_BLOCK stepping, the
// published zran.c method of zlib/examples) and collects, in the same pass, the checkpoints the `.fxi` must carry
// (src/util.c:442-540): one access point per >= `spacing` bytes of output at a deflate block boundary --
// {compressed offset, bit offset, uncompressed offset, the 32 KiB of output in front of it}.  The inflated bytes are
// then staged into HBM like any plain file (all later work runs on the GPU); a reader holding the checkpoints can
// resume inflation at any of them: inflateInit2(raw) + inflatePrime(bits) + inflateSetDictionary(window), which is
// what tests/test_gzip_cpu.py does with every point.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <zlib.h>
#include "../../include/fxg.h"

void fxg_set_error(const char *fmt, ...);

struct fxg_gzip_result {
    uint8_t *data = nullptr;                 // inflated bytes (malloc)
    int64_t  size = 0;
    std::vector<int64_t> cmp_offset, uncmp_offset;
    std::vector<uint8_t> bits, has_data;
    std::vector<uint8_t> windows;            // window_size bytes per point with has_data
    int64_t  compressed_size = 0;
    uint32_t spacing = 0, window_size = 0;
};

extern "C" int fxg_gzip_inflate_host(const void *comp, int64_t nbytes, uint32_t spacing, fxg_gzip_result **out) {
    if (!comp || nbytes < 18 || !out) { fxg_set_error("invalid argument: fxg_gzip_inflate_host"); return FXG_EINVAL; }
    *out = nullptr;
    const uint32_t WINDOW = 32768;
    if (spacing < WINDOW) spacing = 1u << 20;                 // reference: zran_init(spacing 1048576, window 32768), index.c:70
    fxg_gzip_result *r = new fxg_gzip_result();
    r->compressed_size = nbytes; r->spacing = spacing; r->window_size = WINDOW;
    size_t cap = (size_t)nbytes * 4 + (1u << 20);
    r->data = (uint8_t *)malloc(cap);
    if (!r->data) { delete r; fxg_set_error("out of memory (gzip output)"); return FXG_ENOMEM; }
    z_stream strm;
    memset(&strm, 0, sizeof(strm));
    if (inflateInit2(&strm, 47) != Z_OK) { free(r->data); delete r; fxg_set_error("inflateInit2 failed"); return FXG_EFORMAT; }
    const uint8_t *in = (const uint8_t *)comp;
    int64_t totin = 0, totout = 0, last = 0;
    bool first_point = true;
    int ret = Z_OK;
    while (totin < nbytes) {
        strm.next_in = const_cast<Bytef *>(in + totin);
        const int64_t chunk_in = nbytes - totin < (1 << 30) ? nbytes - totin : (1 << 30);
        strm.avail_in = (uInt)chunk_in;
        for (;;) {
            if ((size_t)totout + (1u << 16) > cap) {
                cap = cap * 2;
                uint8_t *nd = (uint8_t *)realloc(r->data, cap);
                if (!nd) { inflateEnd(&strm); free(r->data); delete r; fxg_set_error("out of memory (gzip output)"); return FXG_ENOMEM; }
                r->data = nd;
            }
            const size_t room = cap - (size_t)totout;
            strm.next_out = r->data + totout;
            strm.avail_out = (uInt)(room < (1u << 30) ? room : (1u << 30));
            const uInt in0 = strm.avail_in, out0 = strm.avail_out;
            ret = inflate(&strm, Z_BLOCK);
            totin += in0 - strm.avail_in;
            totout += out0 - strm.avail_out;
            if (ret == Z_NEED_DICT || ret == Z_DATA_ERROR || ret == Z_MEM_ERROR || ret == Z_STREAM_ERROR) {
                inflateEnd(&strm); free(r->data); delete r;
                fxg_set_error("corrupt gzip stream (zlib error %d at compressed offset %lld)", ret, (long long)totin);
                return FXG_EFORMAT;
            }
            if (ret == Z_STREAM_END) break;
            // at a deflate block boundary, not the last block: a possible access point (zran.c)
            if ((strm.data_type & 128) && !(strm.data_type & 64) && (first_point || totout - last >= (int64_t)spacing)) {
                const bool with_window = totout > 0;
                r->cmp_offset.push_back(totin);
                r->uncmp_offset.push_back(totout);
                r->bits.push_back((uint8_t)(strm.data_type & 7));
                r->has_data.push_back(with_window ? 1 : 0);
                if (with_window) {
                    const size_t w0 = r->windows.size();
                    r->windows.resize(w0 + WINDOW, 0);
                    const int64_t have = totout < (int64_t)WINDOW ? totout : (int64_t)WINDOW;
                    memcpy(r->windows.data() + w0 + (WINDOW - have), r->data + totout - have, (size_t)have);
                }
                last = totout;
                first_point = false;
            }
            if (strm.avail_in == 0) break;
        }
        if (ret == Z_STREAM_END) {
            // a further gzip member may follow (concatenated streams); trailing zero padding ends the file
            while (totin < nbytes && in[totin] == 0) ++totin;
            if (totin >= nbytes) break;
            if (inflateReset(&strm) != Z_OK) break;
            first_point = true;                                // the next member's first block is an access point again
        }
    }
    inflateEnd(&strm);
    if (ret != Z_STREAM_END) { free(r->data); delete r; fxg_set_error("truncated gzip stream"); return FXG_EFORMAT; }
    r->size = totout;
    *out = r;
    return FXG_OK;
}

extern "C" const uint8_t *fxg_gzip_data(const fxg_gzip_result *r, int64_t *size) {
    if (size) *size = r ? r->size : 0;
    return r ? r->data : nullptr;
}

// fills `gz` with pointers INTO the result (valid until fxg_gzip_free)
extern "C" int fxg_gzip_index(const fxg_gzip_result *r, fxg_gzindex *gz) {
    if (!r || !gz) { fxg_set_error("invalid argument: fxg_gzip_index"); return FXG_EINVAL; }
    memset(gz, 0, sizeof(*gz));
    gz->compressed_size = r->compressed_size; gz->uncompressed_size = r->size;
    gz->spacing = r->spacing; gz->window_size = r->window_size;
    gz->npoints = (int64_t)r->cmp_offset.size();
    gz->cmp_offset = r->cmp_offset.data(); gz->uncmp_offset = r->uncmp_offset.data();
    gz->bits = r->bits.data(); gz->has_data = r->has_data.data(); gz->windows = r->windows.data();
    return FXG_OK;
}

extern "C" void fxg_gzip_free(fxg_gzip_result *r) {
    if (!r) return;
    free(r->data);
    delete r;
}

// ---- BGZF writer (bench / test tooling: the image has no bgzip) ------------------------------------------
// `data` is cut into 0xff00-byte blocks, each deflated on its own (raw deflate, `level`) into a gzip member with the
// 'BC' extra field, by all host threads; an empty EOF member ends the file.  *out is malloc'ed (fxg_free_host).
#include <thread>
#include <atomic>
extern "C" int fxg_bgzf_compress_host(const void *data, int64_t nbytes, int level, uint8_t **out, int64_t *out_len) {
    if ((!data && nbytes) || nbytes < 0 || !out || !out_len) { fxg_set_error("invalid argument: fxg_bgzf_compress_host"); return FXG_EINVAL; }
    const int64_t BLOCK = 0xff00;
    const int64_t nblk = (nbytes + BLOCK - 1) / BLOCK;
    const size_t bound = 18 + compressBound((uLong)BLOCK) + 8 + 64;
    std::vector<uint32_t> sizes((size_t)nblk, 0);
    uint8_t *scratch = (uint8_t *)malloc((size_t)nblk * bound + 64);
    if (!scratch) { fxg_set_error("out of memory (bgzf scratch)"); return FXG_ENOMEM; }
    unsigned T = std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 128) T = 128;
    if ((int64_t)T > nblk) T = (unsigned)(nblk > 0 ? nblk : 1);
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&] {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            for (;;) {
                const int64_t b = next.fetch_add(1);
                if (b >= nblk) break;
                const uint8_t *src = (const uint8_t *)data + b * BLOCK;
                const int64_t len = nbytes - b * BLOCK < BLOCK ? nbytes - b * BLOCK : BLOCK;
                uint8_t *dst = scratch + (size_t)b * bound;
                static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
                memcpy(dst, hdr, 16);
                deflateReset(&zs);
                zs.next_in = const_cast<Bytef *>(src); zs.avail_in = (uInt)len;
                zs.next_out = dst + 18; zs.avail_out = (uInt)(bound - 18 - 8);
                if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad = 1; break; }
                const uint32_t clen = (uint32_t)zs.total_out;
                const uint32_t total = 18 + clen + 8;
                if (total > 65536) { bad = 1; break; }
                dst[16] = (uint8_t)((total - 1) & 0xff); dst[17] = (uint8_t)((total - 1) >> 8);
                const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)len);
                uint8_t *tl = dst + 18 + clen;
                for (int i = 0; i < 4; ++i) { tl[i] = (uint8_t)(crc >> (8 * i)); tl[4 + i] = (uint8_t)((uint32_t)len >> (8 * i)); }
                sizes[(size_t)b] = total;
            }
            deflateEnd(&zs);
        });
    for (auto &x : th) x.join();
    if (bad) { free(scratch); fxg_set_error("deflate failed"); return FXG_EFORMAT; }
    int64_t total = 28;
    for (int64_t b = 0; b < nblk; ++b) total += sizes[(size_t)b];
    uint8_t *o = (uint8_t *)malloc((size_t)total);
    if (!o) { free(scratch); fxg_set_error("out of memory (bgzf output)"); return FXG_ENOMEM; }
    int64_t p = 0;
    for (int64_t b = 0; b < nblk; ++b) { memcpy(o + p, scratch + (size_t)b * bound, sizes[(size_t)b]); p += sizes[(size_t)b]; }
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    memcpy(o + p, eof, 28);
    free(scratch);
    *out = o; *out_len = total;
    return FXG_OK;
}
