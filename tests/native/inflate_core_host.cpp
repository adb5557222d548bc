// Host build of the per-member DEFLATE decoder (pyfastx_b200/csrc/fxg_inflate_core.cuh) for the CPU test
// suite: the same functions the CUDA kernel runs one thread per member, checked here against zlib.
// Test infrastructure only -- the product never loads this.
#include "../../pyfastx_b200/csrc/fxg_inflate_core.cuh"

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

extern "C" int fxi_host_inflate(const uint8_t *in, int64_t in_size, const int64_t *cmp_off, const int64_t *ucmp_off,
                                int64_t n_members, uint8_t *out, int64_t out_cap, int32_t *status) {
    int bad = 0;
    for (int64_t m = 0; m < n_members; ++m) {
        fxi::MemberTables T;
        const fxi::DeflateConsts K = {LEN_BASE, LEN_EXTRA, DIST_BASE, DIST_EXTRA, CL_ORDER};
        status[m] = fxi::inflate_member(in, in_size, cmp_off[m], cmp_off[m + 1], out, out_cap, ucmp_off[m], ucmp_off[m + 1], T, K);
        bad += status[m] != 0;
    }
    return bad;
}

// one segment per zran checkpoint: (cmp_off, bits, ucmp_off[i] .. ucmp_off[i + 1]) with the 32 KiB window of the points
// that have one (in point order); ucmp_off has n_points + 1 entries (the last = uncompressed size)
extern "C" int fxi_host_inflate_points(const uint8_t *in, int64_t in_size, int64_t n_points, const int64_t *cmp_off,
                                       const uint8_t *bits, const int64_t *ucmp_off, const uint8_t *has_data,
                                       const uint8_t *windows, int window_size, uint8_t *out, int64_t out_cap, int32_t *status) {
    int bad = 0;
    int64_t k = 0;
    for (int64_t i = 0; i < n_points; ++i) {
        fxi::MemberTables T;
        const fxi::DeflateConsts K = {LEN_BASE, LEN_EXTRA, DIST_BASE, DIST_EXTRA, CL_ORDER};
        const uint8_t *w = nullptr;
        int wl = 0;
        if (has_data && has_data[i]) { w = windows + (int64_t)k * window_size; wl = window_size; ++k; }
        status[i] = fxi::inflate_segment(in, in_size, cmp_off[i], bits ? bits[i] : 0, out, out_cap, ucmp_off[i], ucmp_off[i + 1], w, wl, T, K);
        bad += status[i] != 0;
    }
    return bad;
}
