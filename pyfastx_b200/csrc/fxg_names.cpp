// fxg_names.cpp -- batched name -> row resolution, host side of libfxg.so.
//
// Replaces one prepared-statement probe of the sqlite name index per query -- pyfastx_index_get_seq_by_name
// (src/index.c:527-566: "SELECT * FROM seq WHERE chrom=?", src/fasta.c:122) and pyfastx_fastq_get_read_by_name
// (src/fastq.c:487-519) -- by an open-addressing hash table over the packed name buffer the index build already
// holds, built once (in parallel, lock-free) and probed for a whole batch of query names by several threads.
// Duplicate names resolve to the first record, like sqlite's rowid-ordered scan does when the UNIQUE index is absent.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/fxg.h"

void fxg_set_error(const char *fmt, ...);

struct fxg_nametab {
    const uint8_t *names = nullptr;
    const int64_t *off = nullptr;
    int64_t n = 0;
    uint64_t mask = 0;
    std::atomic<uint32_t> *slots = nullptr;     // row + 1, 0 = empty (row counts stay below 2^32 - 1)
};

namespace {

inline uint64_t hash_bytes(const uint8_t *p, int64_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 32;
        p += 8; n -= 8;
    }
    uint64_t w = 0;
    if (n) memcpy(&w, p, (size_t)n);
    h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

inline bool same(const fxg_nametab *t, uint32_t row, const uint8_t *q, int64_t ql) {
    const int64_t l = t->off[row + 1] - t->off[row];
    return l == ql && memcmp(t->names + t->off[row], q, (size_t)ql) == 0;
}

unsigned threads_for(int64_t n) {
    unsigned t = std::thread::hardware_concurrency();
    if (t > 32) t = 32;
    if (t < 1) t = 1;
    if ((int64_t)t > (n + 49999) / 50000) t = (unsigned)((n + 49999) / 50000);
    return t < 1 ? 1 : t;
}

}  // namespace

extern "C" int fxg_nametab_build(const uint8_t *names, const int64_t *name_off, int64_t n, fxg_nametab **out) {
    if (!out || n < 0 || (n && (!names || !name_off)) || n >= 0xfffffff0ll) { fxg_set_error("invalid argument: fxg_nametab_build"); return FXG_EINVAL; }
    fxg_nametab *t = new fxg_nametab();
    t->names = names; t->off = name_off; t->n = n;
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    t->mask = cap - 1;
    t->slots = (std::atomic<uint32_t> *)calloc(cap, sizeof(std::atomic<uint32_t>));
    if (!t->slots) { delete t; fxg_set_error("out of memory (name table)"); return FXG_ENOMEM; }
    const unsigned T = threads_for(n);
    std::vector<std::thread> th;
    for (unsigned k = 0; k < T; ++k)
        th.emplace_back([=] {
            for (int64_t i = n * k / T; i < n * (k + 1) / T; ++i) {
                const uint8_t *p = names + name_off[i];
                const int64_t l = name_off[i + 1] - name_off[i];
                uint64_t s = hash_bytes(p, l) & t->mask;
                const uint32_t me = (uint32_t)i + 1;
                for (;;) {
                    uint32_t cur = t->slots[s].load(std::memory_order_acquire);
                    if (cur == 0) {
                        if (t->slots[s].compare_exchange_weak(cur, me, std::memory_order_acq_rel)) break;
                        continue;                                   // lost the race: look at the slot again
                    }
                    if (same(t, cur - 1, p, l)) {                   // duplicate name: the smaller row stays
                        if (cur <= me) break;
                        if (t->slots[s].compare_exchange_weak(cur, me, std::memory_order_acq_rel)) break;
                        continue;
                    }
                    s = (s + 1) & t->mask;
                }
            }
        });
    for (auto &x : th) x.join();
    *out = t;
    return FXG_OK;
}

extern "C" int64_t fxg_nametab_find(const fxg_nametab *t, const uint8_t *name, int64_t len) {
    if (!t || len < 0 || (len && !name)) return -1;
    uint64_t s = hash_bytes(name, len) & t->mask;
    for (;;) {
        const uint32_t cur = t->slots[s].load(std::memory_order_relaxed);
        if (cur == 0) return -1;
        if (same(t, cur - 1, name, len)) return (int64_t)cur - 1;
        s = (s + 1) & t->mask;
    }
}

extern "C" int fxg_nametab_lookup(const fxg_nametab *t, const uint8_t *q, const int64_t *q_off, int64_t nq, int64_t *ids_out) {
    if (!t || nq < 0 || (nq && (!q_off || !ids_out))) { fxg_set_error("invalid argument: fxg_nametab_lookup"); return FXG_EINVAL; }
    const unsigned T = threads_for(nq);
    std::vector<std::thread> th;
    for (unsigned k = 0; k < T; ++k)
        th.emplace_back([=] {
            for (int64_t i = nq * k / T; i < nq * (k + 1) / T; ++i)
                ids_out[i] = fxg_nametab_find(t, q + q_off[i], q_off[i + 1] - q_off[i]);
        });
    for (auto &x : th) x.join();
    return FXG_OK;
}

extern "C" void fxg_nametab_free(fxg_nametab *t) {
    if (!t) return;
    free((void *)t->slots);
    delete t;
}
