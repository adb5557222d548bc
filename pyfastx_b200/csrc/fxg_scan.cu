// fxg_scan.cu -- K1 (FASTA) and K2 (FASTQ) index-build scans for sm_100a.
//
// Replaces the per-line loops of pyfastx_create_index (reference src/index.c:226-361) and
// pyfastx_fastq_create_index (src/fastq.c:84-171), both driven by ks_getuntil2
// (src/kseq.c:59-109), with ONE pass over the file bytes resident in HBM.
//
// Design:
//   * persistent CTAs claim 16 KiB tiles in file order from an atomic counter; each tile
//     (+ a 256 B left halo) is brought into shared memory by a 1-D TMA bulk copy
//     (cp.async.bulk -> SASS UBLKCP) through a 3-stage mbarrier ring;
//   * phase A: every thread tests 4 x 16 B for '\n' with 3 ALU ops per 32-bit word and the
//     warp turns the per-chunk masks into ordered newline indices with ballots/popc;
//   * the tile publishes {#newlines, #header starts, last two newline positions} and obtains
//     the exclusive prefix over all earlier tiles with a decoupled look-back (single pass,
//     no second read of the file);
//   * phase C: one thread per LINE.  Every quantity the reference carries from line to line
//     is re-expressed as a local rule on (this line, previous line, global line index,
//     global header ordinal):
//       - a header line writes boff / dlen / elen / name length / its line index;
//       - a sequence line that follows a header writes llen;
//       - a sequence line whose length differs from the previous sequence line raises an
//         "event" (count, min/max line index, sum of length deltas) on its record;
//     a tiny finalize kernel then derives blen, slen, norm per record from neighbouring
//     headers and the event summary (proof of equivalence with index.c:325-342 in DESIGN.md);
//   * FASTQ needs no per-record state at all: line k of the file writes field k%4 of row k/4.
#include "fxg_common.cuh"
#include <stdlib.h>

namespace fxg {

#ifdef FXG_SCAN_PROFILE
#define FXG_DBG(x) x
#else
#define FXG_DBG(x)
#endif

constexpr int TILE    = 16384;          // bytes per tile
constexpr int HALO    = 256;            // left halo kept in smem (previous line starts)
constexpr int LAG     = 2;          // phase C of tile i-LAG runs after phase A of tile i
constexpr int STAGES  = LAG + 2;    // tiles i-LAG..i-1 (awaiting phase C) | tile i (phase A) | tile i+1 (loading)
constexpr int NSLOT   = LAG + 1;    // line-list / mailbox slots
constexpr int THREADS = 512;          // worker threads (16 warps) + one prefix warp
constexpr int NWARPS  = THREADS / 32;
constexpr int REGION  = TILE / NWARPS;  // contiguous bytes per warp (2048)
constexpr int BPT     = TILE / THREADS; // contiguous bytes per worker thread (32)
constexpr int NCH     = BPT / 16;       // 16-byte chunks per thread (2)
constexpr int LB      = 512;            // line-list capacity of a regular tile (lines >= 32 B on average)
constexpr int SEGCAP  = 32;             // per-warp entry segment (2 KiB region: lines >= 21 B on average)
constexpr int STAGE_BYTES = HALO + TILE;
static_assert(SEGCAP * (THREADS / 32) <= LB, "a tile without segment overflow must fit the line list");

constexpr int64_t NOPOS = INT64_MIN / 4;

// ---- decoupled look-back state ---------------------------------------------------------------
// Two 16-byte entries per tile, every 64-bit word self-validating (0 = not written yet), so no
// fence / flag ordering is needed and a reader costs ONE L2 round trip:
//   cnt[t] = { st<<62 | newlines , st<<62 | header starts }   st 1 = tile aggregate, 2 = inclusive prefix
//   pos[t] = { p_last + 2 , p_prev + 2 }  positions of the tile's last two newlines; 1 = none
constexpr uint64_t ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;

struct Agg {          // exclusive prefix handed to phase C
    uint64_t nl, hdr;
    int64_t  p_last, p_prev;
    int      k;
};

__device__ __forceinline__ ulonglong2 ld_desc(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(ulonglong2 *p, uint64_t x, uint64_t y) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) { return (uint64_t)shfl_i64((int64_t)v, src); }

// Exclusive (newline, header) counts of tile t: warp-wide look-back over windows of 128 tiles
// (4 independent 16-byte loads per lane in flight).  Sums are commutative, so a window reduces
// with two REDUX instructions plus one shuffle of the (large) inclusive value it ends on.
__device__ void lookback_counts(const ulonglong2 *cnt, int64_t t, int lane, uint64_t &ex_nl, uint64_t &ex_hdr, int &nwin) {
    uint64_t snl = 0, shdr = 0;
    nwin = 0;
    int64_t j0 = t - 1;
    while (true) {
        ulonglong2 v[4];
        ++nwin;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int64_t j = j0 - lane - 32 * m;
            v[m] = (j >= 0) ? ld_desc(cnt + j) : make_ulonglong2(ST_INC, ST_INC);   // before tile 0: prefix 0
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int64_t j = j0 - lane - 32 * m;
            while ((v[m].x & ST_MASK) == 0 || ((v[m].x ^ v[m].y) & ST_MASK) != 0) v[m] = ld_desc(cnt + j);
            const bool inc = (v[m].x & ST_MASK) == ST_INC;
            const uint32_t ball = __ballot_sync(0xffffffffu, inc);
            const int first = ball ? (__ffs(ball) - 1) : 32;
            const uint32_t a = lane < first ? (uint32_t)(v[m].x & ~ST_MASK) : 0u;   // tile aggregates are < 2^15
            const uint32_t b = lane < first ? (uint32_t)(v[m].y & ~ST_MASK) : 0u;
            snl += __reduce_add_sync(0xffffffffu, a);
            shdr += __reduce_add_sync(0xffffffffu, b);
            if (ball) {
                snl += shfl_u64(v[m].x & ~ST_MASK, first);
                shdr += shfl_u64(v[m].y & ~ST_MASK, first);
                ex_nl = snl; ex_hdr = shdr;
                return;
            }
        }
        j0 -= 128;
    }
}

// The two newlines preceding tile t (buffer-relative); -1 is the virtual newline before byte 0.
__device__ void lookback_positions(const ulonglong2 *pos, int64_t t, int64_t &p_last, int64_t &p_prev, int &k) {
    int64_t got[2] = {NOPOS, NOPOS};
    int n = 0;
    for (int64_t j = t - 1; n < 2; --j) {
        if (j < 0) { got[n++] = -1; break; }
        ulonglong2 pv;
        do { pv = ld_desc(pos + j); } while (pv.x == 0 || pv.y == 0);
        if (pv.x != 1) {
            got[n++] = (int64_t)pv.x - 2;
            if (n < 2 && pv.y != 1) got[n++] = (int64_t)pv.y - 2;
        }
    }
    p_last = got[0]; p_prev = got[1]; k = n;
}

struct __align__(16) FastaTmp {   // per header slot (slot 0 = lines before the first header)
    int64_t  boff;       // header thread
    int64_t  lineidx;    // header thread: buffer-local line index of the header line
    int64_t  llen;       // first sequence line (len + 1)
    uint64_t S;          // sum of (L - prevL) over events (wrapping)
    uint64_t evmax;      // max line index of an event
    uint64_t evminc;     // max of ~lineidx  (== ~min)
    uint32_t D;          // number of events
    int32_t  dlen;
    int32_t  nlen;
    uint32_t elen;
};
static_assert(sizeof(FastaTmp) == 64, "FastaTmp layout");

struct ScanTotals {     // written by the CTA that owns the last tile (+ finalize)
    uint64_t nl;
    uint64_t hdr;
    int64_t  n_eff;     // n + 1 if the last line has no '\n'
    uint64_t sum_len;   // FASTA: sum(slen) (finalize); FASTQ: sum(rlen)
    int64_t  lead_lines, lead_bytes, lead_llen;
    uint64_t pad;
};

struct ScanParams {
    const uint8_t *file;
    int64_t   n;            // bytes
    int64_t   capacity;     // readable bytes at file (multiple of 16)
    int64_t   ntiles;
    int64_t   base_offset;  // added to every file offset written to rows
    int64_t   first_line;   // FASTQ: global index of the first line of this buffer
    int       flags;
    ulonglong2 *cnt;        // look-back: counts (aggregate -> inclusive, in place)
    ulonglong2 *pos;        // look-back: last two newline positions per tile
    uint32_t *tile_counter;
    ScanTotals *totals;
    FastaTmp *tmp;          // FASTA
    int64_t   tmp_cap;      // slots
    fxg_fastq_row *qrows;   // FASTQ
    int64_t   qrows_cap;
    unsigned long long *dbg; // optional cycle counters (FXG_SCAN_DEBUG=1)
};

// Per-CTA roles (warp specialisation).  All hand-offs are mbarriers on a ring of RING tile slots; no
// role ever executes another role's code, so the instruction cost of a tile is the sum of
//   16 byte warps   phase A only: newline masks of their 1 KiB region -> ordered entries in the slot
//    1 publisher    16 per-warp counts -> tile aggregate, published for the other CTAs
//    1 prefix warp  decoupled look-back (the only code that waits on other CTAs)
//    1 producer     claims tiles in file order and issues their TMA loads as ring slots free up
//    4 line warps   phase C: one thread per line, for the (few) lines of a tile
// and byte warps stream tile after tile without ever waiting for a look-back.
constexpr int NBYTE = NWARPS;                  // 16 byte warps
constexpr int NLINE = 4;                       // line warps
constexpr int W_PUB = NBYTE, W_PREF = NBYTE + 1, W_PROD = NBYTE + 2, W_LINE0 = NBYTE + 3;
constexpr int CTA_THREADS = (NBYTE + 3 + NLINE) * 32;
constexpr int LINE_THREADS = NLINE * 32;
constexpr int RING = 5;                        // tile slots (stage buffer + metadata) in flight per CTA
constexpr int SEG_PER_LINE_WARP = NBYTE / NLINE;

struct Pref {           // prefix warp -> line warps
    uint64_t ex_nl, ex_hdr;
    int64_t  cpos[2];   // the two newlines preceding the tile: [1] = nearest, [0] = the one before (NOPOS if none)
    uint32_t cflag[2];  // bit31: the line starting after that newline begins with '>' ; low bits: header count (only [1])
    uint32_t pad[2];
};
struct Slot {           // metadata of one tile in flight
    uint16_t seg_pos[NBYTE][SEGCAP];     // per byte-warp newline positions (tile relative), file order
    uint16_t seg_flag[NBYTE][SEGCAP];    // bit15: next line starts with '>', low bits: warp-local inclusive header count
    uint32_t wcnt[NBYTE];                // per warp: newlines | header starts << 16
    uint32_t wstart[NBYTE];              // exclusive prefix of wcnt (publisher)
    int64_t  tile;                       // tile id, -1 = no more tiles
    int      T_nl, dense;                // newlines in the tile; some warp overflowed its segment
    uint32_t tsh, pad0;                  // the tile's first byte starts a header line
    Pref     pref;
};

__device__ __forceinline__ void line_bar() { asm volatile("bar.sync 2, %0;" ::"n"(LINE_THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// First k in [0, limit) with h[k] == a or h[k] == b (h in shared memory), else limit.  Word-wise:
// four aligned 32-bit loads in flight per step instead of one dependent byte load per character.
// (May read up to 15 bytes past h + limit: the dynamic shared buffer carries 16 bytes of slack.)
__device__ __forceinline__ int64_t find_first_of2(const uint8_t *h, int64_t limit, uint32_t a4, uint32_t b4, int *which) {
    const int mis = (int)((uintptr_t)h & 3);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(h - mis);
    *which = 0;
    for (int64_t k = -mis; k < limit; k += 16, wp += 4) {
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        uint32_t ma[4] = {byte_eq_mask(w0, a4), byte_eq_mask(w1, a4), byte_eq_mask(w2, a4), byte_eq_mask(w3, a4)};
        uint32_t mb[4] = {byte_eq_mask(w0, b4), byte_eq_mask(w1, b4), byte_eq_mask(w2, b4), byte_eq_mask(w3, b4)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t m = ma[i] | mb[i];
            const int64_t kk = k + 4 * i;
            if (kk < 0) m &= 0xffffffffu << (8 * (int)(-kk));        // bytes before h (only the first word)
            if (m) {
                const int byte = (__ffs(m) - 1) >> 3;
                const int64_t pos = kk + byte;
                if (pos >= limit) return limit;
                *which = ((mb[i] >> (8 * byte)) & 0x80u) ? 2 : 1;
                return pos;
            }
        }
    }
    return limit;
}

template <int MODE>   // 0 = FASTA, 1 = FASTQ
__global__ void __launch_bounds__(CTA_THREADS, 2) scan_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[RING];   // TMA data landed            (tx)    -> byte warps
    __shared__ __align__(8) uint64_t fill_bar[RING];   // byte warps wrote the slot   (16)    -> publisher
    __shared__ __align__(8) uint64_t mail_bar[RING];   // aggregate published         (1)     -> prefix warp
    __shared__ __align__(8) uint64_t pref_bar[RING];   // exclusive prefix ready      (1)     -> line warps
    __shared__ __align__(8) uint64_t free_bar[RING];   // line warps done             (NLINE) -> TMA issuer
    __shared__ __align__(16) Slot slots[RING];
    __shared__ int64_t  l_pos[LB + 2];       // dense path only: line list of the tile ([0],[1] = carry)
    __shared__ uint32_t l_flag[LB + 2];
    __shared__ uint32_t s_scan[NLINE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t n = P.n;
    const bool virt = (n > 0) && (P.file[n - 1] != '\n');
    const int64_t n_eff = n + (virt ? 1 : 0);
    const bool full_name = (P.flags & FXG_SCAN_FULL_NAME) != 0;

    if (tid == 0) {
        for (int s = 0; s < RING; ++s) {
            mbar_init(&full_bar[s], 1); mbar_init(&fill_bar[s], NBYTE); mbar_init(&mail_bar[s], 1);
            mbar_init(&pref_bar[s], 1); mbar_init(&free_bar[s], NLINE);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto stage_ptr = [&](int s) -> uint8_t * { return dyn_smem + (size_t)s * STAGE_BYTES + HALO; };
    // byte at buffer-relative position x, given the tile (id t, base) whose stage is tb
    auto byte_at = [&](const uint8_t *tb, int64_t t, int64_t base, int64_t x) -> uint8_t {
        const int64_t r = x - base;
        if (r >= -(int64_t)HALO && (t > 0 || r >= 0)) return tb[r];
        return P.file[x];
    };

    // =========================================================================================
    // byte warps: phase A, tile after tile
    // =========================================================================================
    if (warp < NBYTE) {
        const uint32_t k0a = reg_const(0x0a0a0a0au), k7f = reg_const(0x7f7f7f7fu), k80 = reg_const(0x80808080u);
        const int lbase = warp * REGION + lane * BPT;              // this thread's BPT contiguous bytes
        const int rot = (lane / (8 / NCH)) % NCH;                  // chunk rotation: conflict-free LDS.128
        int s = 0;
        uint32_t par = 0;
        for (;;) {
            mbar_wait(&full_bar[s], par);
            Slot &sl = slots[s];
            const int64_t t = sl.tile;
            if (t < 0) {                                           // no more tiles: pass the baton and leave
                __syncwarp();
                if (lane == 0) mbar_arrive(&fill_bar[s]);
                break;
            }
            uint8_t *tb = stage_ptr(s);
            const int64_t base = t * TILE;
            // the tile that contains EOF: every warp neutralises the bytes past n in its own region and the
            // owner of position n plants the virtual newline
            if (base + TILE > n) {
                for (int x = warp * REGION + lane; x < (warp + 1) * REGION; x += 32)
                    if (base + x >= n) tb[x] = (virt && base + x == n) ? (uint8_t)'\n' : (uint8_t)0;
                __syncwarp();
            }
            // ---------------- phase A ------------------------------------------------------------------
            uint4 v[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) v[j] = *reinterpret_cast<const uint4 *>(tb + lbase + 16 * ((j + rot) % NCH));
            uint32_t m[NCH];
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < NCH; ++j) { m[j] = chunk_eq_mask_r(v[j], k0a, k7f, k80); cnt += __popc(m[j]); }
            uint32_t run;         // newlines | header starts << 16 of the whole warp
            if (!__any_sync(0xffffffffu, cnt > 1)) {
                int x = 0;
                uint32_t nh = 0;
                if (cnt) {
                    int j = 0;
                    uint32_t mm = 0;
#pragma unroll
                    for (int jj = NCH - 1; jj >= 0; --jj) if (m[jj]) { j = jj; mm = m[jj]; }
                    x = lbase + 16 * ((j + rot) % NCH) + chunk_bit_to_off(__ffs(mm) - 1);
                    if (MODE == 0) nh = (x + 1 < TILE && base + x + 1 < n && tb[x + 1] == '>') ? 1u : 0u;
                }
                const uint32_t bn = __ballot_sync(0xffffffffu, cnt != 0);
                const uint32_t bh = (MODE == 0) ? __ballot_sync(0xffffffffu, nh != 0) : 0u;
                if (cnt) {
                    const int wi = __popc(bn & lt_mask);
                    sl.seg_pos[warp][wi] = (uint16_t)x;
                    sl.seg_flag[warp][wi] = (uint16_t)((nh << 15) | (__popc(bh & lt_mask) + nh));
                }
                run = __popc(bn) | (__popc(bh) << 16);
            } else if (!__any_sync(0xffffffffu, cnt > 2)) {
                // at most two newlines per thread (e.g. a header line inside the 32 bytes): two ballots
                int xa = 0x7fffffff, xb = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    if (m[j]) {
                        const int qoff = lbase + 16 * ((j + rot) % NCH);
                        const int x1 = qoff + chunk_bit_to_off(__ffs(m[j]) - 1);
                        if (x1 < xa) { xb = xa; xa = x1; } else if (x1 < xb) xb = x1;
                        const uint32_t m2 = m[j] & (m[j] - 1);
                        if (m2) {
                            const int x2 = qoff + chunk_bit_to_off(__ffs(m2) - 1);
                            if (x2 < xa) { xb = xa; xa = x2; } else if (x2 < xb) xb = x2;
                        }
                    }
                }
                uint32_t nha = 0, nhb = 0;
                if (MODE == 0) {
                    if (cnt >= 1) nha = (xa + 1 < TILE && base + xa + 1 < n && tb[xa + 1] == '>') ? 1u : 0u;
                    if (cnt >= 2) nhb = (xb + 1 < TILE && base + xb + 1 < n && tb[xb + 1] == '>') ? 1u : 0u;
                }
                const uint32_t bn1 = __ballot_sync(0xffffffffu, cnt >= 1), bn2 = __ballot_sync(0xffffffffu, cnt >= 2);
                const uint32_t bh1 = (MODE == 0) ? __ballot_sync(0xffffffffu, nha != 0) : 0u;
                const uint32_t bh2 = (MODE == 0) ? __ballot_sync(0xffffffffu, nhb != 0) : 0u;
                const int wi = __popc(bn1 & lt_mask) + __popc(bn2 & lt_mask);
                const uint32_t hcb = __popc(bh1 & lt_mask) + __popc(bh2 & lt_mask);
                if (cnt >= 1 && wi < SEGCAP) { sl.seg_pos[warp][wi] = (uint16_t)xa; sl.seg_flag[warp][wi] = (uint16_t)((nha << 15) | (hcb + nha)); }
                if (cnt >= 2 && wi + 1 < SEGCAP) { sl.seg_pos[warp][wi + 1] = (uint16_t)xb; sl.seg_flag[warp][wi + 1] = (uint16_t)((nhb << 15) | (hcb + nha + nhb)); }
                run = (__popc(bn1) + __popc(bn2)) | ((__popc(bh1) + __popc(bh2)) << 16);
            } else {
                // short lines (three or more newlines in some thread's bytes): shuffle scan + ordered
                // iteration over the masks
                uint32_t h = 0;
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        uint32_t mm = m[j];
                        while (mm) {
                            const int x = lbase + 16 * ((j + rot) % NCH) + chunk_bit_to_off(__ffs(mm) - 1);
                            mm &= mm - 1;
                            if (x + 1 < TILE && base + x + 1 < n && tb[x + 1] == '>') ++h;
                        }
                    }
                }
                const uint32_t my_cnt = cnt | (h << 16);
                uint32_t incl = my_cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += o;
                }
                run = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t wi = (incl - my_cnt) & 0xffffu, hc = (incl - my_cnt) >> 16;
                if (cnt && (run & 0xffffu) <= (uint32_t)SEGCAP) {
#pragma unroll 1
                    for (int qc = 0; qc < NCH; ++qc) {
                        const int j = (qc - rot + NCH) % NCH;
                        uint32_t mq = 0;
#pragma unroll
                        for (int jj = 0; jj < NCH; ++jj) if (jj == j) mq = m[jj];
#pragma unroll 1
                        for (int w = 0; w < 4; ++w) {                 // byte order: word, then byte
                            uint32_t mw = mq & (0x80808080u >> w);
                            while (mw) {
                                const int x = lbase + 16 * qc + chunk_bit_to_off(__ffs(mw) - 1);
                                mw &= mw - 1;
                                uint32_t nh = 0;
                                if (MODE == 0) { nh = (x + 1 < TILE && base + x + 1 < n && tb[x + 1] == '>') ? 1u : 0u; hc += nh; }
                                sl.seg_pos[warp][wi] = (uint16_t)x;
                                sl.seg_flag[warp][wi] = (uint16_t)((nh << 15) | hc);
                                ++wi;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { sl.wcnt[warp] = run; mbar_arrive(&fill_bar[s]); }
            if (++s == RING) { s = 0; par ^= 1u; }
        }
        return;
    }

    // =========================================================================================
    // producer: claims tiles in file order and issues their TMA loads as ring slots are released.
    // A tile is claimed only when its load can start at once, so it is published a short, fixed time
    // later -- the byte warps never wait for anything but TMA data.
    // =========================================================================================
    if (warp == W_PROD) {
        if (lane == 0) {
            for (int64_t seq = 0;; ++seq) {
                const int s = (int)(seq % RING);
                if (seq >= RING) mbar_wait(&free_bar[s], (uint32_t)(((seq / RING) - 1) & 1));
                const int64_t t = (int64_t)atomicAdd(P.tile_counter, 1u);
                uint8_t *buf = dyn_smem + (size_t)s * STAGE_BYTES;
                bool loaded = false;
                if (t < P.ntiles) {
                    slots[s].tile = t;
                    const int64_t base = t * TILE;
                    int64_t src = base - HALO, dst = 0, want = STAGE_BYTES;
                    if (t == 0) { src = 0; dst = HALO; want = TILE; }
                    int64_t avail = P.capacity - src;
                    if (avail < want) want = avail > 0 ? (avail & ~(int64_t)15) : 0;
                    if (want > 0) {
                        fence_proxy_async();
                        mbar_expect_tx(&full_bar[s], (uint32_t)want);
                        tma_load_1d(buf + dst, P.file + src, (uint32_t)want, &full_bar[s]);
                        loaded = true;
                    }
                } else {
                    slots[s].tile = -1;
                }
                if (!loaded) mbar_arrive(&full_bar[s]);
                if (t >= P.ntiles) break;
            }
        }
        return;
    }

    // =========================================================================================
    // publisher warp: per-warp counts -> tile aggregate, published for the other CTAs right away
    // =========================================================================================
    if (warp == W_PUB) {
        int s = 0;
        uint32_t par = 0;
        for (;;) {
            mbar_wait(&fill_bar[s], par);
            Slot &sl = slots[s];
            const int64_t t = sl.tile;
            if (t < 0) { if (lane == 0) mbar_arrive(&mail_bar[s]); break; }
            const int64_t base = t * TILE;
            const uint8_t *tb = stage_ptr(s);
            const uint32_t c = lane < NBYTE ? sl.wcnt[lane] : 0u;
            const bool dense = __any_sync(0xffffffffu, (c & 0xffffu) > (uint32_t)SEGCAP);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < NBYTE; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane < NBYTE) sl.wstart[lane] = incl - c;
            const uint32_t ttot = __shfl_sync(0xffffffffu, incl, NBYTE - 1);
            const int T_nl = (int)(ttot & 0xffffu);
            if (lane == 0) {
                const bool line_start_at_base = (base == 0) || (tb[-1] == '\n');
                const uint32_t tsh = (MODE == 0 && line_start_at_base && base < n && tb[0] == '>') ? 1u : 0u;
                const uint32_t T_h = (ttot >> 16) + tsh;
                int64_t l0 = NOPOS, l1 = NOPOS;
                if (!dense) {
                    int found = 0;
                    for (int w = NBYTE - 1; w >= 0 && found < 2; --w) {
                        const int cw = (int)(sl.wcnt[w] & 0xffffu);
                        for (int k = cw - 1; k >= 0 && found < 2; --k) {
                            const int64_t pp = base + sl.seg_pos[w][k];
                            if (found == 0) l0 = pp; else l1 = pp;
                            ++found;
                        }
                    }
                } else {
                    int found = 0;
                    for (int x = TILE - 1; x >= 0 && found < 2; --x)
                        if (tb[x] == '\n') { if (found == 0) l0 = base + x; else l1 = base + x; ++found; }
                }
                st_desc(&P.cnt[t], ST_AGG | (uint64_t)T_nl, ST_AGG | (uint64_t)T_h);
                st_desc(&P.pos[t], T_nl >= 1 ? (uint64_t)(l0 + 2) : 1ull, T_nl >= 2 ? (uint64_t)(l1 + 2) : 1ull);
                sl.T_nl = T_nl; sl.dense = dense ? 1 : 0; sl.tsh = tsh;
                mbar_arrive(&mail_bar[s]);
            }
            __syncwarp();
            if (++s == RING) { s = 0; par ^= 1u; }
        }
        return;
    }

    // =========================================================================================
    // prefix warp: decoupled look-back for one tile after the other
    // =========================================================================================
    if (warp == W_PREF) {
        int s = 0;
        uint32_t par = 0;
        for (;;) {
            mbar_wait(&mail_bar[s], par);
            Slot &sl = slots[s];
            const int64_t t = sl.tile;
            if (t < 0) { if (lane == 0) mbar_arrive(&pref_bar[s]); break; }
            const int64_t base = t * TILE;
            const uint8_t *tb = stage_ptr(s);
            uint64_t ex_nl = 0, ex_hdr = 0;
            int nwin = 0;
            lookback_counts(P.cnt, t, lane, ex_nl, ex_hdr, nwin);
            if (lane == 0) {
                const ulonglong2 own = ld_desc(&P.cnt[t]);
                const uint64_t in_nl = ex_nl + (own.x & ~ST_MASK), in_hdr = ex_hdr + (own.y & ~ST_MASK);
                st_desc(&P.cnt[t], ST_INC | in_nl, ST_INC | in_hdr);
                if (t == P.ntiles - 1) { P.totals->nl = in_nl; P.totals->hdr = in_hdr; P.totals->n_eff = n_eff; }
                int64_t pl, pp; int k;
                lookback_positions(P.pos, t, pl, pp, k);
                uint32_t nh1 = 0, hc1 = 0, nh2 = 0;
                if (MODE == 0) {
                    if (pl + 1 == base) { nh1 = sl.tsh; hc1 = sl.tsh; }
                    else nh1 = (byte_at(tb, t, base, pl + 1) == '>') ? 1u : 0u;
                    if (k >= 2) nh2 = (byte_at(tb, t, base, pp + 1) == '>') ? 1u : 0u;
                }
                Pref &pr = sl.pref;
                pr.ex_nl = ex_nl; pr.ex_hdr = ex_hdr;
                pr.cpos[1] = pl;                    pr.cflag[1] = (nh1 << 31) | hc1;
                pr.cpos[0] = k >= 2 ? pp : NOPOS;   pr.cflag[0] = (nh2 << 31);
                mbar_arrive(&pref_bar[s]);
            }
            __syncwarp();
            if (++s == RING) { s = 0; par ^= 1u; }
        }
        return;
    }

    // =========================================================================================
    // line warps: phase C
    // =========================================================================================
    const int lw = warp - W_LINE0;                 // 0..NLINE-1
    const int ltid = lw * 32 + lane;               // 0..LINE_THREADS-1
    unsigned long long my_size = 0;                // FASTQ: sum of rlen seen by this thread
    struct TileState { int64_t t, base; };

    // ---- one line: newline at p, previous newlines pm1, pm2, their flags (bit31: the line that STARTS
    //      after that newline is a header; low bits: tile-level header count up to there), line index ----
    auto do_line_v = [&](const TileState &S, const uint8_t *tb, const Pref &pr, int64_t p, int64_t pm1, int64_t pm2,
                         uint32_t f1, uint32_t f2, int idx) {
        const int64_t s = pm1 + 1;
        const int64_t L = p - pm1;                            // len + 1
        const int64_t lineidx = (int64_t)pr.ex_nl + idx;      // buffer-local line index
        const int rp = (int)(p - S.base);                     // newline, tile relative (>= 0)
        const bool near = (s - S.base) >= -(int64_t)HALO && (S.t > 0 || s >= S.base);   // line start inside smem window
        const int rs = (int)(s - S.base);
        if (MODE == 0) {
            const bool is_hdr = (f1 >> 31) != 0;
            const int64_t slot = (int64_t)pr.ex_hdr + (int64_t)(f1 & 0x7fffffffu);   // rec + 1
            if (is_hdr) {
                const uint8_t before = (rp >= 1 || S.t > 0) ? tb[rp - 1] : P.file[p - 1];
                const int elen = (before == '\r') ? 2 : 1;
                const int64_t dlen = L - 1 - elen;
                int64_t nlen = dlen;
                if (!full_name) {
                    nlen = 0;
                    if (near) {
                        int which;
                        nlen = find_first_of2(tb + rs + 1, dlen, 0x20202020u, 0x09090909u, &which);
                    } else {
                        while (nlen < dlen) {
                            const uint8_t ch = byte_at(tb, S.t, S.base, s + 1 + nlen);
                            if (ch == ' ' || ch == '\t') break;
                            ++nlen;
                        }
                    }
                }
                if (slot < P.tmp_cap) {
                    FastaTmp *r = &P.tmp[slot];
                    r->boff = P.base_offset + p + 1;
                    r->lineidx = lineidx;
                    r->dlen = (int32_t)dlen;
                    r->nlen = (int32_t)nlen;
                    r->elen = (uint32_t)elen;
                }
            } else if (slot < P.tmp_cap) {
                const bool prev_exists = pm1 >= 0;
                const bool prev_is_hdr = (f2 >> 31) != 0;
                if (!prev_exists || prev_is_hdr) {
                    P.tmp[slot].llen = L;
                } else {
                    const int64_t prevL = pm1 - pm2;
                    if (L != prevL) {
                        FastaTmp *r = &P.tmp[slot];
                        atomicAdd(&r->D, 1u);
                        atomicMax((unsigned long long *)&r->evmax, (unsigned long long)lineidx);
                        atomicMax((unsigned long long *)&r->evminc, ~(unsigned long long)lineidx);
                        atomicAdd((unsigned long long *)&r->S, (unsigned long long)(L - prevL));
                    }
                }
            }
        } else {
            const int64_t gline = P.first_line + lineidx;
            const int ph = (int)(gline & 3);
            const int64_t row = (gline >> 2) - (P.first_line >> 2);
            const int64_t len = L - 1;
            if (ph == 1) {
                const uint8_t before = (rp >= 1 || S.t > 0) ? tb[rp - 1] : (p >= 1 ? P.file[p - 1] : (uint8_t)0);
                const int64_t rlen = (len > 0 && before == '\r') ? len - 1 : len;
                my_size += (unsigned long long)rlen;
                if (row < P.qrows_cap) { P.qrows[row].soff = P.base_offset + s; P.qrows[row].rlen = rlen; }
            } else if (row < P.qrows_cap) {
                if (ph == 0) {
                    const uint8_t before = (rp >= 1 || S.t > 0) ? tb[rp - 1] : (p >= 1 ? P.file[p - 1] : (uint8_t)0);
                    int64_t l = len - 1;
                    if (l > 0 && before == '\r') --l;
                    if (l < 0) l = 0;
                    int64_t k = 0;
                    if (near) {
                        int which;
                        k = find_first_of2(tb + rs + 1, l, 0x20202020u, 0x00000000u, &which);
                        if (which == 2) k = l;          // a NUL before any space: strchr() finds nothing (fastq.c:112)
                    } else {
                        for (; k < l; ++k) {
                            const uint8_t ch = byte_at(tb, S.t, S.base, s + 1 + k);
                            if (ch == 0) { k = l; break; }
                            if (ch == ' ') break;
                        }
                    }
                    *reinterpret_cast<int2 *>(&P.qrows[row].dlen) = make_int2((int)len, (int)k);
                } else if (ph == 3) {
                    P.qrows[row].qoff = P.base_offset + s;
                }
            }
        }
    };

    // ---- regular tile: line warp lw owns the segments of byte warps [lw*4, lw*4+4); their lines are
    //      flattened over the 32 lanes; predecessors come from the segments (or the tile carry) ---------
    auto phase_c = [&](const TileState &S, const uint8_t *tb, const Slot &sl) {
        const int w0 = lw * SEG_PER_LINE_WARP;
        const uint32_t tsh = sl.tsh;
        auto eflag = [&](int w2, int k) -> uint32_t {
            const uint32_t f = sl.seg_flag[w2][k];
            return ((f >> 15) << 31) | (tsh + (sl.wstart[w2] >> 16) + (f & 0x7fffu));
        };
        // entry `back` (1 or 2) positions before the first entry of byte-warp segment w2
        auto before_seg = [&](int w2, int back, int64_t &pp, uint32_t &ff) {
            for (int w3 = w2 - 1; w3 >= 0; --w3) {
                const int c = (int)(sl.wcnt[w3] & 0xffffu);
                if (c >= back) { pp = S.base + sl.seg_pos[w3][c - back]; ff = eflag(w3, c - back); return; }
                back -= c;
            }
            pp = sl.pref.cpos[2 - back]; ff = sl.pref.cflag[2 - back];      // back 1 -> cpos[1], back 2 -> cpos[0]
        };
        int cn[SEG_PER_LINE_WARP], tot = 0;
#pragma unroll
        for (int b = 0; b < SEG_PER_LINE_WARP; ++b) { cn[b] = (int)(sl.wcnt[w0 + b] & 0xffffu); tot += cn[b]; }
        for (int i = lane; i < tot; i += 32) {
            int b = 0, k = i;
#pragma unroll
            for (int bb = 0; bb < SEG_PER_LINE_WARP - 1; ++bb) if (b == bb && k >= cn[bb]) { k -= cn[bb]; ++b; }
            const int w = w0 + b;
            const int64_t p = S.base + sl.seg_pos[w][k];
            int64_t pm1, pm2;
            uint32_t f1, f2;
            if (k >= 1) { pm1 = S.base + sl.seg_pos[w][k - 1]; f1 = eflag(w, k - 1); } else before_seg(w, 1, pm1, f1);
            if (k >= 2) { pm2 = S.base + sl.seg_pos[w][k - 2]; f2 = eflag(w, k - 2); } else before_seg(w, 2 - k, pm2, f2);
            do_line_v(S, tb, sl.pref, p, pm1, pm2, f1, f2, (int)(sl.wstart[w] & 0xffffu) + k);
        }
    };

    // ---- dense tile (a byte warp found more than SEGCAP newlines in its region): the line warps rebuild
    //      the newline list byte-wise in batches of LB lines; slow but fully general --------------------------
    auto dense_tile = [&](const TileState &S, const uint8_t *tb, const Slot &sl) {
        constexpr int DB = TILE / LINE_THREADS;          // bytes per line thread
        const int dbase = ltid * DB;
        uint32_t my_cnt = 0;
        for (int b = 0; b < DB; ++b) {
            const int x = dbase + b;
            if (tb[x] == '\n') { my_cnt += 1; if (MODE == 0 && x + 1 < TILE && S.base + x + 1 < n && tb[x + 1] == '>') my_cnt += 1u << 16; }
        }
        uint32_t incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_scan[lw] = incl;
        line_bar();
        uint32_t wb = 0;
        for (int w = 0; w < lw; ++w) wb += s_scan[w];
        const uint32_t excl = wb + incl - my_cnt;
        const int T_nl = sl.T_nl;
        for (int b0 = 0; b0 < T_nl; b0 += LB) {
            line_bar();                                   // previous batch fully consumed
            int64_t c0 = 0, c1 = 0; uint32_t g0 = 0, g1 = 0;
            if (b0 > 0 && ltid == 0) { c0 = l_pos[LB]; c1 = l_pos[LB + 1]; g0 = l_flag[LB]; g1 = l_flag[LB + 1]; }
            line_bar();
            if (ltid == 0) {
                if (b0 == 0) { l_pos[0] = sl.pref.cpos[0]; l_pos[1] = sl.pref.cpos[1]; l_flag[0] = sl.pref.cflag[0]; l_flag[1] = sl.pref.cflag[1]; }
                else { l_pos[0] = c0; l_pos[1] = c1; l_flag[0] = g0; l_flag[1] = g1; }
            }
            int idx = (int)(excl & 0xffffu);
            uint32_t hc = sl.tsh + (excl >> 16);
            if ((my_cnt & 0xffffu) && idx < b0 + LB && idx + (int)(my_cnt & 0xffffu) > b0) {
                for (int b = 0; b < DB; ++b) {
                    const int x = dbase + b;
                    if (tb[x] != '\n') continue;
                    uint32_t nh = 0;
                    if (MODE == 0) { nh = (x + 1 < TILE && S.base + x + 1 < n && tb[x + 1] == '>') ? 1u : 0u; hc += nh; }
                    const int rel = idx - b0;
                    if (rel >= 0 && rel < LB) { l_pos[2 + rel] = S.base + x; l_flag[2 + rel] = (nh << 31) | hc; }
                    ++idx;
                }
            }
            line_bar();
            const int nb = T_nl - b0 < LB ? T_nl - b0 : LB;
            for (int i = ltid; i < nb; i += LINE_THREADS)
                do_line_v(S, tb, sl.pref, l_pos[2 + i], l_pos[1 + i], l_pos[i], l_flag[1 + i], l_flag[i], b0 + i);
        }
        line_bar();
    };

    {
        int s = 0;
        uint32_t par = 0;
        for (;;) {
            mbar_wait(&pref_bar[s], par);
            Slot &sl = slots[s];
            const int64_t t = sl.tile;
            if (t < 0) break;
            TileState S;
            S.t = t; S.base = t * TILE;
            if (sl.dense) dense_tile(S, stage_ptr(s), sl);
            else phase_c(S, stage_ptr(s), sl);
            __syncwarp();
            if (lane == 0) mbar_arrive(&free_bar[s]);
            if (++s == RING) { s = 0; par ^= 1u; }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_size += (unsigned long long)shfl_down_i64((int64_t)my_size, d);
        if (lane == 0 && my_size) atomicAdd((unsigned long long *)&P.totals->sum_len, my_size);
    }
}

// ---- FASTA finalize: per-record fields from neighbouring headers + event summary ------------
// blen  = next header start - boff (or end position)                       index.c:243,348
// slen  = blen - n_lines * elen  (sum over lines of len - elen + 1)        index.c:335-338
// norm  = [#lines differing from the first <= 1], from the events (DESIGN.md proof)  index.c:325-342
__global__ void fasta_finalize_kernel(const FastaTmp *tmp, int64_t nrows, int64_t base_offset,
                                      ScanTotals *tot, fxg_fasta_row *rows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long slen_acc = 0;
    if (r < nrows) {
        const FastaTmp t = tmp[r + 1];
        int64_t next_h, next_line;
        if (r + 1 < nrows) {
            const FastaTmp nx = tmp[r + 2];
            next_h = nx.boff - (1 + nx.dlen + (int64_t)nx.elen);
            next_line = nx.lineidx;
        } else {
            next_h = base_offset + tot->n_eff;
            next_line = (int64_t)tot->nl;
        }
        const int64_t blen = next_h - t.boff;
        const int64_t nlines = next_line - t.lineidx - 1;
        const int64_t slen = blen - nlines * (int64_t)t.elen;
        int norm;
        if (t.D == 0) norm = 1;
        else if (t.D == 1) norm = ((int64_t)t.evmax == t.lineidx + nlines) ? 1 : 0;
        else if (t.D == 2) norm = ((t.evmax - (~t.evminc) == 1) && t.S == 0) ? 1 : 0;
        else norm = 0;
        fxg_fasta_row o;
        o.boff = t.boff; o.blen = blen; o.slen = slen; o.llen = t.llen;
        o.dlen = t.dlen; o.nlen = t.nlen; o.elen = (uint8_t)t.elen; o.norm = (uint8_t)norm;
        for (int i = 0; i < 6; ++i) o.pad[i] = 0;
        // uniform lines: no length change at all, or a single SHORTER line at the very end
        o.pad[0] = (t.D == 0 || (t.D == 1 && (int64_t)t.evmax == t.lineidx + nlines && (int64_t)t.S < 0)) ? 1 : 0;
        rows[r] = o;
        slen_acc = (unsigned long long)slen;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) slen_acc += (unsigned long long)shfl_down_i64((int64_t)slen_acc, d);
    if ((threadIdx.x & 31) == 0 && slen_acc) atomicAdd((unsigned long long *)&tot->sum_len, slen_acc);
    if (r == 0) {
        // lead part = lines before the first header of this buffer (multi-GPU shard merge)
        const FastaTmp l = tmp[0];
        tot->lead_llen = l.llen;
        if (nrows > 0) {
            const FastaTmp f = tmp[1];
            tot->lead_lines = f.lineidx;
            tot->lead_bytes = f.boff - base_offset - (1 + f.dlen + (int64_t)f.elen);
        }
    }
}

// ---- density sample: headers / newlines in evenly spaced windows (capacity estimate) -------
__global__ void sample_density_kernel(const uint8_t *file, int64_t n, int64_t nwin, int64_t win,
                                      unsigned long long *out /* [0]=newlines, [1]=header starts */) {
    const int64_t w = blockIdx.x;
    const int64_t start = (nwin > 1) ? (int64_t)((__int128)(n - win) * w / (nwin - 1)) : 0;
    unsigned long long nl = 0, h = 0;
    for (int64_t i = start + threadIdx.x; i < start + win && i < n; i += blockDim.x) {
        if (file[i] == '\n') { ++nl; if (i + 1 < n && file[i + 1] == '>') ++h; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        nl += (unsigned long long)shfl_down_i64((int64_t)nl, d);
        h += (unsigned long long)shfl_down_i64((int64_t)h, d);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], nl); atomicAdd(&out[1], h); }
}

// ---- plain newline count (FASTQ multi-GPU phase pass) ----------------------------------------
__global__ void count_newlines_kernel(const uint8_t *file, int64_t n, unsigned long long *out) {
    const int64_t nvec = n / 16;
    unsigned long long cnt = 0;
    const uint4 *v = reinterpret_cast<const uint4 *>(file);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
        cnt += __popc(chunk_eq_mask(v[i], 0x0a0a0a0au));
    if (blockIdx.x == 0)
        for (int64_t i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) cnt += (file[i] == '\n');
    for (int d = 16; d > 0; d >>= 1) cnt += (unsigned long long)shfl_down_i64((int64_t)cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

}  // namespace fxg

// =============================================================================================
// host side
// =============================================================================================
using namespace fxg;

static int scan_launch_config(fxg_ctx *ctx, int mode, int *grid, size_t *smem) {
    *smem = (size_t)RING * STAGE_BYTES + 16;     // + slack for word-wise header reads
    int per_sm = 0;
    if (mode == 0) {
        FXG_CUDA(cudaFuncSetAttribute(scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
        FXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<0>, CTA_THREADS, *smem));
    } else {
        FXG_CUDA(cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
        FXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<1>, CTA_THREADS, *smem));
    }
    if (per_sm < 1) per_sm = 1;
    *grid = ctx->sm_count * per_sm;
    return FXG_OK;
}

// estimate (#newlines, #headers) of the whole buffer from 256 evenly spaced 16 KiB windows
static int sample_density(fxg_ctx *ctx, const fxg_file *f, double *nl_per_byte, double *hdr_per_byte) {
    const int64_t n = f->size;
    const int64_t win = 16384;
    int64_t nwin = n / win;
    if (nwin > 256) nwin = 256;
    if (nwin < 1) nwin = 1;
    FXG_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, 64, ctx->stream));
    ctx->launches += 1;
    sample_density_kernel<<<(unsigned)nwin, 256, 0, ctx->stream>>>(f->d, n, nwin, win < n ? win : n,
                                                                  (unsigned long long *)ctx->counters.ptr);
    FXG_CUDA(cudaGetLastError());
    unsigned long long h[2];
    FXG_CUDA(cudaMemcpyAsync(h, ctx->counters.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream));
    FXG_CUDA(cudaStreamSynchronize(ctx->stream));
    const double sampled = (double)nwin * (double)(win < n ? win : n);
    *nl_per_byte = sampled > 0 ? (double)h[0] / sampled : 0;
    *hdr_per_byte = sampled > 0 ? (double)h[1] / sampled : 0;
    return FXG_OK;
}

static int run_scan(fxg_ctx *ctx, const fxg_file *f, int mode, int64_t base_offset, int64_t first_line, int flags,
                    void **d_rows_out, fxg_scan_stats *stats) {
    FXG_CHECK_ARG(ctx && f && stats, "null ctx/file/stats");
    FXG_CUDA(cudaSetDevice(ctx->device));
    memset(stats, 0, sizeof(*stats));
    const int64_t n = f->size;
    if (d_rows_out) *d_rows_out = nullptr;
    if (n == 0) return FXG_OK;
    const int64_t ntiles = (n + 1 + TILE - 1) / TILE;   // room for a virtual newline at n
    int rc;
    if ((rc = ctx->counters.reserve(256))) return rc;
    if ((rc = ctx->tile_desc.reserve((size_t)ntiles * 2 * sizeof(ulonglong2)))) return rc;

    double nlpb = 0, hpb = 0;
    if ((rc = sample_density(ctx, f, &nlpb, &hpb))) return rc;
    int64_t cap;
    if (mode == 0) cap = (int64_t)(hpb * (double)n * 1.5) + 4096;
    else cap = (int64_t)(nlpb * (double)n * 1.25 / 4.0) + 4096;

    int grid = 0; size_t smem = 0;
    if ((rc = scan_launch_config(ctx, mode, &grid, &smem))) return rc;
    if ((int64_t)grid > ntiles) grid = (int)ntiles;

    for (int attempt = 0; attempt < 3; ++attempt) {
        if (mode == 0) {
            if ((rc = ctx->row_tmp.reserve((size_t)(cap + 2) * sizeof(FastaTmp)))) return rc;
            FXG_CUDA(cudaMemsetAsync(ctx->row_tmp.ptr, 0, (size_t)(cap + 2) * sizeof(FastaTmp), ctx->stream));
        } else {
            if ((rc = ctx->rows.reserve((size_t)(cap + 2) * sizeof(fxg_fastq_row)))) return rc;
            // rows are fully overwritten except partially-owned boundary rows
            FXG_CUDA(cudaMemsetAsync(ctx->rows.ptr, 0, sizeof(fxg_fastq_row), ctx->stream));
        }
        FXG_CUDA(cudaMemsetAsync(ctx->tile_desc.ptr, 0, (size_t)ntiles * 2 * sizeof(ulonglong2), ctx->stream));
        FXG_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, 256, ctx->stream));

        ScanParams P;
        memset(&P, 0, sizeof(P));
        P.file = f->d; P.n = n; P.capacity = f->capacity & ~(int64_t)15; P.ntiles = ntiles;
        P.base_offset = base_offset; P.first_line = first_line; P.flags = flags;
        P.cnt = (ulonglong2 *)ctx->tile_desc.ptr; P.pos = P.cnt + ntiles;
        P.tile_counter = (uint32_t *)ctx->counters.ptr;
        P.totals = (ScanTotals *)((uint8_t *)ctx->counters.ptr + 64);
        P.tmp = (FastaTmp *)ctx->row_tmp.ptr; P.tmp_cap = cap + 1;
        P.qrows = (fxg_fastq_row *)ctx->rows.ptr; P.qrows_cap = cap;
        const bool dbg = getenv("FXG_SCAN_DEBUG") != nullptr;
        P.dbg = dbg ? (unsigned long long *)((uint8_t *)ctx->counters.ptr + 128) : nullptr;

        {
            FxgProfScope prof(ctx, FXG_PROF_SCAN);
            if (mode == 0) scan_kernel<0><<<grid, CTA_THREADS, smem, ctx->stream>>>(P);
            else scan_kernel<1><<<grid, CTA_THREADS, smem, ctx->stream>>>(P);
        }
        FXG_CUDA(cudaGetLastError());

        ScanTotals tot;
        FXG_CUDA(cudaMemcpyAsync(&tot, P.totals, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
        FXG_CUDA(cudaStreamSynchronize(ctx->stream));

        if (dbg) {
            unsigned long long h[8];
            cudaMemcpy(h, P.dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[fxg scan dbg] grid=%d tiles=%lld | worker cycles/tile: wait_data=%.0f phaseA=%.0f wait_pref=%.0f phaseC=%.0f | prefix warp: lookback=%.0f cyc/tile, windows=%.2f/tile, wait_mail=%.0f\n",
                    grid, (long long)ntiles, (double)h[0] / ntiles, (double)h[1] / ntiles, (double)h[2] / ntiles, (double)h[3] / ntiles,
                    (double)h[4] / (h[6] ? h[6] : 1), (double)h[5] / (h[6] ? h[6] : 1), (double)h[7] / (h[6] ? h[6] : 1));
        }
        const int64_t nrows = mode == 0 ? (int64_t)tot.hdr
                                        : (int64_t)((first_line + (int64_t)tot.nl + 3) / 4 - first_line / 4);
        if (nrows > cap) { cap = nrows + 16; continue; }   // estimate too small: exact rerun

        stats->n_lines = (int64_t)tot.nl;
        stats->end_position = tot.n_eff;
        if (mode == 0) {
            stats->n_rows = nrows;
            if ((rc = ctx->rows.reserve((size_t)(nrows + 1) * sizeof(fxg_fasta_row)))) return rc;
            const int64_t work = nrows > 0 ? nrows : 1;
            {
                FxgProfScope prof(ctx, FXG_PROF_FINALIZE);
                fasta_finalize_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(
                    P.tmp, nrows, base_offset, P.totals, (fxg_fasta_row *)ctx->rows.ptr);
            }
            FXG_CUDA(cudaGetLastError());
            FXG_CUDA(cudaMemcpyAsync(&tot, P.totals, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
            FXG_CUDA(cudaStreamSynchronize(ctx->stream));
            stats->total_len = (int64_t)tot.sum_len;
            stats->lead_llen = tot.lead_llen;
            if (nrows > 0) { stats->lead_lines = tot.lead_lines; stats->lead_bytes = tot.lead_bytes; }
            else { stats->lead_lines = (int64_t)tot.nl; stats->lead_bytes = n; }
        } else {
            // complete reads only (fastq.c:132-146,159); rows owned partially by this buffer are
            // still present in the array for the shard merge
            stats->n_rows = (first_line + (int64_t)tot.nl) / 4 - first_line / 4;
            stats->total_len = (int64_t)tot.sum_len;
        }
        if (d_rows_out) *d_rows_out = ctx->rows.ptr;
        return FXG_OK;
    }
    fxg_set_error("scan: row capacity estimate failed to converge");
    return FXG_ECUDA;
}

extern "C" int fxg_fasta_scan(fxg_ctx *ctx, const fxg_file *f, int64_t base_offset, int flags,
                              fxg_fasta_row **d_rows_out, fxg_scan_stats *stats) {
    return run_scan(ctx, f, 0, base_offset, 0, flags, (void **)d_rows_out, stats);
}

extern "C" int fxg_fastq_scan(fxg_ctx *ctx, const fxg_file *f, int64_t base_offset, int64_t first_line,
                              fxg_fastq_row **d_rows_out, fxg_scan_stats *stats) {
    FXG_CHECK_ARG(first_line >= 0, "first_line < 0");
    return run_scan(ctx, f, 1, base_offset, first_line, 0, (void **)d_rows_out, stats);
}

extern "C" int fxg_count_lines(fxg_ctx *ctx, const fxg_file *f, int64_t *n_newlines, int *ends_with_newline) {
    FXG_CHECK_ARG(ctx && f && n_newlines, "null argument");
    FXG_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->counters.reserve(256))) return rc;
    FXG_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, 64, ctx->stream));
    unsigned long long h = 0;
    uint8_t last = '\n';
    if (f->size > 0) {
        ctx->launches += 1;
        count_newlines_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(f->d, f->size,
                                                                         (unsigned long long *)ctx->counters.ptr);
        FXG_CUDA(cudaGetLastError());
        FXG_CUDA(cudaMemcpyAsync(&last, f->d + f->size - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FXG_CUDA(cudaMemcpyAsync(&h, ctx->counters.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FXG_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_newlines = (int64_t)h;
    if (ends_with_newline) *ends_with_newline = (last == '\n');
    return FXG_OK;
}
