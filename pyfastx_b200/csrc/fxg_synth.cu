// fxg_synth.cu -- synthetic FASTA/FASTQ generated directly in HBM (bench / test tooling).
// Byte-identical to pyfastx_b200/synth.py: every base / quality byte is a counter-based
// hash of (seed, record, position), so CPU and GPU produce the same file.
#include "fxg_common.cuh"

namespace fxg {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // murmur3 fmix64
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ uint64_t key64(uint64_t seed, uint64_t rec, uint64_t k) {
    return seed + rec * 0x9E3779B97F4A7C15ull + k * 0xBF58476D1CE4E5B9ull;
}
__device__ __forceinline__ uint8_t base_at(uint64_t seed, uint64_t rec, uint64_t k) {
    return (uint8_t)("ACGT"[mix64(key64(seed, rec, k)) >> 62]);
}
__device__ __forceinline__ uint8_t qual_at(uint64_t seed, uint64_t rec, uint64_t k) {
    const uint64_t z = mix64(key64(seed, rec, k) ^ 0xD6E8FEB86659FD93ull);
    return (uint8_t)(35 + (z >> 32) % 36);
}
__device__ int utoa10(uint64_t v, char *out) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
    return n;
}
__device__ int append(char *dst, int pos, const char *s) {
    while (*s) dst[pos++] = *s++;
    return pos;
}
// total decimal digits of 1..i
__device__ uint64_t digits_upto(uint64_t i) {
    uint64_t total = 0, lo = 1;
    for (int d = 1; d <= 20 && lo <= i; ++d) {
        const uint64_t hi = (lo > UINT64_MAX / 10) ? UINT64_MAX : lo * 10 - 1;
        const uint64_t top = i < hi ? i : hi;
        total += (top - lo + 1) * (uint64_t)d;
        if (hi == UINT64_MAX) break;
        lo *= 10;
    }
    return total;
}

// one CTA per record: ">seq{i} synthetic len={L}\n" + bases wrapped at `width`
__global__ void synth_fasta_kernel(uint64_t seed, const int64_t *lengths, const int64_t *rec_off, int64_t n_records,
                                   int64_t first_record, int width, uint8_t *out) {
    __shared__ char hdr[64];
    __shared__ int hlen;
    for (int64_t r = blockIdx.x; r < n_records; r += gridDim.x) {
        const int64_t L = lengths[r];
        const uint64_t rec = (uint64_t)(first_record + r);
        __syncthreads();
        if (threadIdx.x == 0) {
            int p = append(hdr, 0, ">seq");
            p += utoa10(rec + 1, hdr + p);
            p = append(hdr, p, " synthetic len=");
            p += utoa10((uint64_t)L, hdr + p);
            hdr[p++] = '\n';
            hlen = p;
        }
        __syncthreads();
        uint8_t *o = out + rec_off[r];
        const int64_t body = L + (L + width - 1) / width;
        const int64_t total = hlen + body;
        for (int64_t p = threadIdx.x; p < total; p += blockDim.x) {
            uint8_t ch;
            if (p < hlen) ch = (uint8_t)hdr[p];
            else {
                const int64_t q = p - hlen;
                const int64_t line = q / (width + 1), col = q % (width + 1);
                const int64_t k = line * width + col;
                ch = (col == width || k >= L) ? (uint8_t)'\n' : base_at(seed, rec, (uint64_t)k);
            }
            o[p] = ch;
        }
    }
}

// one warp per read: "@read{i} 1:N:0:ACGT\n" bases "\n+\n" quals "\n"
__global__ void synth_fastq_kernel(uint64_t seed, int64_t n_reads, int64_t first_read, int read_len,
                                   const int64_t *rec_off, uint8_t *out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int fixed = 5 + 11 + 1 + read_len + 1 + 2 + read_len + 1;   // "@read" + " 1:N:0:ACGT\n" + seq\n + "+\n" + qual\n
    for (int64_t r = warp; r < n_reads; r += nwarps) {
        const uint64_t rec = (uint64_t)(first_read + r);
        char hdr[40];
        int p = append(hdr, 0, "@read");
        p += utoa10(rec + 1, hdr + p);
        p = append(hdr, p, " 1:N:0:ACGT");
        hdr[p++] = '\n';
        // offset = sum over reads first_read..rec-1 of (fixed + ndigits(j+1))
        const int64_t off = rec_off ? rec_off[r]
                                    : (int64_t)((uint64_t)r * (uint64_t)fixed + (digits_upto(rec) - digits_upto((uint64_t)first_read)));
        uint8_t *o = out + off;
        const int total = p + read_len + 3 + read_len + 1;
        for (int i = lane; i < total; i += 32) {
            uint8_t ch;
            if (i < p) ch = (uint8_t)hdr[i];
            else if (i < p + read_len) ch = base_at(seed, rec, (uint64_t)(i - p));
            else if (i == p + read_len) ch = '\n';
            else if (i == p + read_len + 1) ch = '+';
            else if (i == p + read_len + 2) ch = '\n';
            else if (i < p + read_len + 3 + read_len) ch = qual_at(seed, rec, (uint64_t)(i - p - read_len - 3));
            else ch = '\n';
            o[i] = ch;
        }
    }
}

}  // namespace fxg

using namespace fxg;

extern "C" int fxg_synth_fasta_dev(fxg_ctx *ctx, uint64_t seed, const int64_t *d_lengths, const int64_t *d_rec_off,
                                   int64_t n_records, int64_t first_record, int width, uint8_t *d_out) {
    FXG_CHECK_ARG(ctx && d_lengths && d_rec_off && d_out && n_records >= 0 && width > 0, "bad arguments");
    FXG_LOCK(ctx);
    if (n_records == 0) return FXG_OK;
    FXG_CUDA(cudaSetDevice(ctx->device));
    int64_t grid = n_records < (int64_t)ctx->sm_count * 32 ? n_records : (int64_t)ctx->sm_count * 32;
    ctx->launches += 1;
    synth_fasta_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(seed, d_lengths, d_rec_off, n_records, first_record, width, d_out);
    FXG_CUDA(cudaGetLastError());
    return FXG_OK;
}

extern "C" int fxg_synth_fastq_dev(fxg_ctx *ctx, uint64_t seed, int64_t n_reads, int64_t first_read, int read_len,
                                   const int64_t *d_rec_off, uint8_t *d_out) {
    FXG_CHECK_ARG(ctx && d_out && n_reads >= 0 && read_len > 0 && first_read >= 0, "bad arguments");
    FXG_LOCK(ctx);
    if (n_reads == 0) return FXG_OK;
    FXG_CUDA(cudaSetDevice(ctx->device));
    ctx->launches += 1;
    synth_fastq_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(seed, n_reads, first_read, read_len, d_rec_off, d_out);
    FXG_CUDA(cudaGetLastError());
    return FXG_OK;
}
