// fxg_fxi.cpp -- bulk writer of `.fxi` index files, host side of libfxg.so.
//
// Replaces the per-row INSERT loops of the reference index build -- 9 binds + sqlite3_step per record
// (src/index.c:223-251), 6 binds per read (src/fastq.c:81-146) -- and the CREATE UNIQUE INDEX that follows them
// (src/index.c:366, src/fastq.c:155).  The rows come out of the GPU scan as one array and the names as one
// packed buffer, in final (rowid) order, so nothing has to be INSERTed: the file is written directly in the
// SQLite database file format (https://www.sqlite.org/fileformat2.html), bottom-up --
//   table b-trees   leaf pages filled left to right with (rowid, record) cells, interior pages from the
//                   per-leaf maximum rowids;  leaf runs are built by several threads at once (leaf pages
//                   hold no page numbers, so runs are independent) and written with pwrite;
//   index b-tree    (name, rowid) entries sorted by a parallel sample sort on 16-byte key prefixes, then
//                   loaded bottom-up the same way (an index interior cell holds a real entry);
//   schema          the reference's tables / columns / index names (src/index.c:178-207, src/fastq.c:29-60),
//                   so the reference's own SELECT statements read the file unchanged.
// Duplicate names make the reference's CREATE UNIQUE INDEX fail silently (no index is created); the same here.
// No sqlite library is involved in writing; tests/test_fxi_cpu.py checks the files with sqlite's own
// PRAGMA integrity_check and SELECT-compares them with the reference's.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
#include <errno.h>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "../../include/fxg.h"

void fxg_set_error(const char *fmt, ...);

#include <chrono>
namespace {

struct StageTimer {
    bool on; std::chrono::steady_clock::time_point t0;
    StageTimer() : on(getenv("FXG_FXI_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fxi] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

constexpr uint32_t PAGE = 4096;                 // usable size U (no reserved bytes)
constexpr uint32_t TBL_MAXLOCAL = PAGE - 35;    // table leaf: X = U - 35
constexpr uint32_t MINLOCAL = (PAGE - 12) * 32 / 255 - 23;
constexpr uint32_t IDX_MAXLOCAL = (PAGE - 12) * 64 / 255 - 23;

inline int put_varint(uint8_t *p, uint64_t v) {
    if (v <= 0x7f) { p[0] = (uint8_t)v; return 1; }
    if (v <= 0x3fff) { p[0] = (uint8_t)((v >> 7) | 0x80); p[1] = (uint8_t)(v & 0x7f); return 2; }
    uint8_t buf[10];
    int n = 0;
    if (v & 0xff00000000000000ull) {            // 9-byte form: 8 x 7 bits + 8 bits
        buf[8] = (uint8_t)v;
        v >>= 8;
        for (int i = 7; i >= 0; --i) { buf[i] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; }
        memcpy(p, buf, 9);
        return 9;
    }
    do { buf[n++] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; } while (v);
    buf[0] &= 0x7f;
    for (int i = 0; i < n; ++i) p[i] = buf[n - 1 - i];
    return n;
}
inline int varint_len(uint64_t v) {
    int n = 1;
    while (v > 0x7f && n < 9) { v >>= 7; ++n; }
    return n;
}
inline void put_be16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
inline void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

// serial type + big-endian body of an integer (schema format 4: 0 and 1 are type-only)
inline int int_serial(int64_t v, uint8_t *body, int *blen) {
    if (v == 0) { *blen = 0; return 8; }
    if (v == 1) { *blen = 0; return 9; }
    int n, t;
    if (v >= -128 && v <= 127) { n = 1; t = 1; }
    else if (v >= -32768 && v <= 32767) { n = 2; t = 2; }
    else if (v >= -8388608 && v <= 8388607) { n = 3; t = 3; }
    else if (v >= -2147483648ll && v <= 2147483647ll) { n = 4; t = 4; }
    else if (v >= -140737488355328ll && v <= 140737488355327ll) { n = 6; t = 5; }
    else { n = 8; t = 6; }
    for (int i = 0; i < n; ++i) body[i] = (uint8_t)((uint64_t)v >> (8 * (n - 1 - i)));
    *blen = n;
    return t;
}

// One value of a generic row (small tables, schema rows)
struct Val {
    enum Kind { NUL, INT, REAL, TEXT, BLOB } kind = NUL;
    int64_t i = 0;
    double d = 0;
    const void *p = nullptr;
    size_t n = 0;
    static Val null() { return Val(); }
    static Val integer(int64_t v) { Val x; x.kind = INT; x.i = v; return x; }
    static Val real(double v) { Val x; x.kind = REAL; x.d = v; return x; }
    static Val text(const void *s, size_t n) { Val x; x.kind = TEXT; x.p = s; x.n = n; return x; }
    static Val text(const char *s) { return text(s, strlen(s)); }
    static Val blob(const void *s, size_t n) { Val x; x.kind = BLOB; x.p = s; x.n = n; return x; }
};

void encode_record(const Val *v, int n, std::vector<uint8_t> &out) {
    uint8_t hdr[9 * 16 + 9], body_small[16 * 8];
    std::vector<uint8_t> big;
    int hl = 0;
    size_t bl = 0;
    // first pass: serial types (header without its own length varint)
    struct Piece { const void *p; size_t n; bool small; size_t off; };
    Piece pieces[16];
    for (int k = 0; k < n; ++k) {
        pieces[k].p = nullptr; pieces[k].n = 0; pieces[k].small = true; pieces[k].off = bl;
        switch (v[k].kind) {
        case Val::NUL: hl += put_varint(hdr + hl, 0); break;
        case Val::INT: {
            int b;
            const int t = int_serial(v[k].i, body_small + bl, &b);
            hl += put_varint(hdr + hl, (uint64_t)t);
            pieces[k].n = (size_t)b; bl += (size_t)b;
            break;
        }
        case Val::REAL: {
            uint64_t u;
            memcpy(&u, &v[k].d, 8);
            for (int i = 0; i < 8; ++i) body_small[bl + i] = (uint8_t)(u >> (8 * (7 - i)));
            hl += put_varint(hdr + hl, 7);
            pieces[k].n = 8; bl += 8;
            break;
        }
        case Val::TEXT: case Val::BLOB:
            hl += put_varint(hdr + hl, (uint64_t)v[k].n * 2 + (v[k].kind == Val::TEXT ? 13 : 12));
            pieces[k].p = v[k].p; pieces[k].n = v[k].n; pieces[k].small = false;
            break;
        }
    }
    int hsize = hl + 1;
    if (hsize > 127) hsize = hl + 2;
    uint8_t hv[9];
    const int hvn = put_varint(hv, (uint64_t)hsize);
    out.clear();
    out.insert(out.end(), hv, hv + hvn);
    out.insert(out.end(), hdr, hdr + hl);
    for (int k = 0; k < n; ++k) {
        if (pieces[k].small) out.insert(out.end(), body_small + pieces[k].off, body_small + pieces[k].off + pieces[k].n);
        else out.insert(out.end(), (const uint8_t *)pieces[k].p, (const uint8_t *)pieces[k].p + pieces[k].n);
    }
}

// ---- a run of b-tree pages built in memory by one thread ----------------------------------------------
// Pages are numbered first_page, first_page + 1, ... in creation order (overflow pages included).
struct PageRun {
    std::vector<uint8_t> bytes;                 // whole pages
    uint32_t first_page = 0;                    // assigned when the runs are laid out in the file
    uint32_t npages() const { return (uint32_t)(bytes.size() / PAGE); }
    uint8_t *new_page() {
        bytes.resize(bytes.size() + PAGE, 0);
        return bytes.data() + bytes.size() - PAGE;
    }
    uint8_t *page(uint32_t i) { return bytes.data() + (size_t)i * PAGE; }
};

// Local page index (within the run) is turned into a file page number by adding first_page; cells that
// contain page numbers (overflow pointers, interior children inside one run) are patched at layout time.
struct Patch { uint32_t page_idx; uint32_t offset; uint32_t target_idx; };   // be32 at page[offset] = first_page + target_idx

// Leaf-page writer: cells grow down from the page end, pointers up from the header.
struct LeafWriter {
    PageRun *run;
    std::vector<Patch> *patches;
    uint8_t type;                               // 0x0D table leaf, 0x0A index leaf
    uint32_t hdr_off = 0;                       // 100 on page 1, else 0
    int32_t cur = -1;                           // local index of the open page
    uint32_t ncell = 0, content = PAGE;
    uint32_t maxlocal;
    std::vector<uint32_t> leaf_pages;           // local indices of finished leaves, in key order

    LeafWriter(PageRun *r, std::vector<Patch> *p, uint8_t t) : run(r), patches(p), type(t) {
        maxlocal = (t == 0x0D) ? TBL_MAXLOCAL : IDX_MAXLOCAL;
    }
    void open() {
        run->new_page();
        cur = (int32_t)run->npages() - 1;
        ncell = 0; content = PAGE;
    }
    void close() {
        if (cur < 0) return;
        uint8_t *pg = run->page((uint32_t)cur) + hdr_off;
        pg[0] = type;
        put_be16(pg + 1, 0);
        put_be16(pg + 3, ncell);
        put_be16(pg + 5, content);              // PAGE == 4096 fits 16 bits
        pg[7] = 0;
        leaf_pages.push_back((uint32_t)cur);
        cur = -1;
    }
    // bytes a cell takes in the page for a payload of `plen` (+ prefix bytes `pre`), and its local part
    uint32_t local_size(uint64_t plen) const {
        if (plen <= maxlocal) return (uint32_t)plen;
        const uint32_t k = MINLOCAL + (uint32_t)((plen - MINLOCAL) % (PAGE - 4));
        return k <= maxlocal ? k : MINLOCAL;
    }
    bool fits(uint32_t cell_bytes) const { return 8 + hdr_off + 2 * (ncell + 1) + cell_bytes <= content; }
    // cell = prefix (varints) + payload (possibly spilling to overflow pages)
    // returns false if it does not fit the open page (caller closes / opens and retries)
    bool add(const uint8_t *prefix, uint32_t pre, const uint8_t *payload, uint64_t plen) {
        const uint32_t loc = local_size(plen);
        const uint32_t cell = pre + loc + (loc < plen ? 4 : 0);
        if (cur < 0) open();
        if (!fits(cell)) return false;
        content -= cell;
        const uint32_t my_page = (uint32_t)cur;
        uint8_t *pg = run->page(my_page);
        memcpy(pg + content, prefix, pre);
        memcpy(pg + content + pre, payload, loc);
        put_be16(pg + hdr_off + 8 + 2 * ncell, content);
        ++ncell;
        if (loc < plen) {                       // overflow chain
            uint64_t done = loc;
            uint32_t patch_page = my_page, patch_off = content + pre + loc;
            while (done < plen) {
                run->new_page();
                const uint32_t ov = run->npages() - 1;
                patches->push_back({patch_page, patch_off, ov});
                uint8_t *op = run->page(ov);
                const uint64_t take = std::min<uint64_t>(PAGE - 4, plen - done);
                memcpy(op + 4, payload + done, (size_t)take);
                done += take;
                patch_page = ov; patch_off = 0;  // next pointer at offset 0 (stays 0 on the last page)
            }
        }
        return true;
    }
};

// ---- database assembly ------------------------------------------------------------------------------
struct Child { uint32_t page; int64_t key; };                    // table tree: max rowid of the subtree

struct Db {
    int fd = -1;
    uint32_t next_page = 2;                                      // page 1 = schema
    std::string err;
    bool pwrite_all(const void *buf, size_t n, uint64_t off) {
        const uint8_t *p = (const uint8_t *)buf;
        while (n) {
            const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
            if (w <= 0) { err = std::string("write failed: ") + strerror(errno); return false; }
            p += w; n -= (size_t)w; off += (uint64_t)w;
        }
        return true;
    }
    // place a run in the file: assign page numbers, apply patches, write
    bool place(PageRun &run, std::vector<Patch> &patches) {
        run.first_page = next_page;
        for (const Patch &pt : patches) put_be32(run.page(pt.page_idx) + pt.offset, run.first_page + pt.target_idx);
        const bool ok = run.bytes.empty() || pwrite_all(run.bytes.data(), run.bytes.size(), (uint64_t)(run.first_page - 1) * PAGE);
        next_page += run.npages();
        return ok;
    }
};

// interior levels of a TABLE b-tree over `kids` (in key order); returns the root page number.
// Every interior page gets at least two children (one cell + right-most pointer).
uint32_t build_table_interior(Db &db, std::vector<Child> kids) {
    while (kids.size() > 1) {
        PageRun run;
        std::vector<Child> up;
        size_t i = 0;
        const size_t m = kids.size();
        while (i < m) {
            uint8_t *pg = run.new_page();
            uint32_t ncell = 0, content = PAGE;
            size_t j = i;
            // children i..b: cells for i..b-1, right-most = b
            while (j + 1 < m) {
                uint8_t cell[16];
                put_be32(cell, kids[j].page);
                const int cl = 4 + put_varint(cell + 4, (uint64_t)kids[j].key);
                if (12 + 2 * (ncell + 1) + (uint32_t)cl > content) break;
                // keep at least two children for the page after this one
                if (ncell >= 1 && m - (j + 1) == 1 && 12 + 2 * (ncell + 2) + (uint32_t)cl + 16 > content) break;
                content -= (uint32_t)cl;
                memcpy(pg + content, cell, (size_t)cl);
                put_be16(pg + 12 + 2 * ncell, content);
                ++ncell; ++j;
            }
            // j is the right-most child of this page.  If exactly one child would remain after it, it cannot
            // form a page of its own: give it this page's last cell (there are >= 2 when the page is full)
            if (m - (j + 1) == 1 && ncell >= 2) { --ncell; --j; }
            pg[0] = 0x05;
            put_be16(pg + 1, 0);
            put_be16(pg + 3, ncell);
            // recompute content start from the remaining cells
            uint32_t cs = PAGE;
            for (uint32_t c = 0; c < ncell; ++c) { const uint32_t o = ((uint32_t)pg[12 + 2 * c] << 8) | pg[12 + 2 * c + 1]; if (o < cs) cs = o; }
            put_be16(pg + 5, cs);
            pg[7] = 0;
            put_be32(pg + 8, kids[j].page);
            up.push_back({db.next_page + run.npages() - 1, kids[j].key});
            i = j + 1;
        }
        std::vector<Patch> none;
        if (!db.place(run, none)) return 0;
        kids.swap(up);
    }
    return kids[0].page;
}

struct IdxChild { uint32_t page; };
// interior levels of an INDEX b-tree: kids[k] separated from kids[k+1] by entry seps[k] (a full payload)
uint32_t build_index_interior(Db &db, std::vector<uint32_t> kids, std::vector<std::vector<uint8_t>> seps) {
    while (kids.size() > 1) {
        PageRun run;
        std::vector<Patch> patches;
        std::vector<uint32_t> up_kids;
        std::vector<std::vector<uint8_t>> up_seps;
        const size_t m = kids.size();
        size_t i = 0;
        while (i < m) {
            const uint32_t pidx = run.npages();
            run.new_page();
            uint32_t ncell = 0, content = PAGE;
            size_t j = i;
            std::vector<uint32_t> cell_sizes;
            while (j + 1 < m) {
                const std::vector<uint8_t> &pl = seps[j];
                uint8_t pre[16];
                put_be32(pre, kids[j]);
                const int pn = 4 + put_varint(pre + 4, pl.size());
                // local part (index interior uses the same X as index leaves)
                uint32_t loc = (uint32_t)pl.size();
                if (pl.size() > IDX_MAXLOCAL) {
                    const uint32_t k = MINLOCAL + (uint32_t)((pl.size() - MINLOCAL) % (PAGE - 4));
                    loc = k <= IDX_MAXLOCAL ? k : MINLOCAL;
                }
                const uint32_t cell = (uint32_t)pn + loc + (loc < pl.size() ? 4 : 0);
                if (12 + 2 * (ncell + 1) + cell > content) break;
                content -= cell;
                uint8_t *pg = run.page(pidx);
                memcpy(pg + content, pre, (size_t)pn);
                memcpy(pg + content + pn, pl.data(), loc);
                put_be16(pg + 12 + 2 * ncell, content);
                if (loc < pl.size()) {
                    uint64_t done = loc;
                    uint32_t ppage = pidx, poff = content + (uint32_t)pn + loc;
                    while (done < pl.size()) {
                        run.new_page();
                        const uint32_t ov = run.npages() - 1;
                        patches.push_back({ppage, poff, ov});
                        const uint64_t take = std::min<uint64_t>(PAGE - 4, pl.size() - done);
                        memcpy(run.page(ov) + 4, pl.data() + done, (size_t)take);
                        done += take;
                        ppage = ov; poff = 0;
                    }
                }
                cell_sizes.push_back(cell);
                ++ncell; ++j;
            }
            if (m - (j + 1) == 1 && ncell >= 2) { --ncell; --j; }   // never leave a single child for the last page
            uint8_t *pg = run.page(pidx);
            pg[0] = 0x02;
            put_be16(pg + 1, 0);
            put_be16(pg + 3, ncell);
            uint32_t cs = PAGE;
            for (uint32_t c = 0; c < ncell; ++c) { const uint32_t o = ((uint32_t)pg[12 + 2 * c] << 8) | pg[12 + 2 * c + 1]; if (o < cs) cs = o; }
            put_be16(pg + 5, cs);
            pg[7] = 0;
            put_be32(pg + 8, kids[j]);
            up_kids.push_back(db.next_page + pidx);
            if (j + 1 < m) up_seps.push_back(seps[j]);
            i = j + 1;
        }
        if (!db.place(run, patches)) return 0;
        kids.swap(up_kids);
        seps.swap(up_seps);
    }
    return kids[0];
}

unsigned worker_count(int64_t n) {
    unsigned t = std::thread::hardware_concurrency();
    if (t > 32) t = 32;
    if (t < 1) t = 1;
    const int64_t per = 50000;                  // not worth a thread below this many rows
    if ((int64_t)t > (n + per - 1) / per) t = (unsigned)((n + per - 1) / per);
    return t < 1 ? 1 : t;
}

// ---- the big table: rows [0, n) with rowid = i + 1, record written by `enc(i, buf)` -> payload length --------
template <class Enc>
uint32_t build_big_table(Db &db, int64_t n, Enc enc) {
    if (n == 0) {
        PageRun run; std::vector<Patch> none;
        uint8_t *pg = run.new_page();
        pg[0] = 0x0D; put_be16(pg + 3, 0); put_be16(pg + 5, PAGE);
        if (!db.place(run, none)) return 0;
        return run.first_page;
    }
    const unsigned T = worker_count(n);
    std::vector<PageRun> runs(T);
    std::vector<std::vector<Patch>> patches(T);
    std::vector<std::vector<uint32_t>> leaves(T);
    std::vector<std::vector<int64_t>> maxkey(T);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) {
        const int64_t a = n * t / T, b = n * (t + 1) / T;
        th.emplace_back([&, t, a, b] {
            runs[t].bytes.reserve((size_t)((b - a) * 72 + 65536));
            LeafWriter w(&runs[t], &patches[t], 0x0D);
            std::vector<uint8_t> big;
            uint8_t small[4096 + 64];
            for (int64_t i = a; i < b; ++i) {
                uint8_t *buf = small;
                uint64_t plen = enc(i, small, sizeof(small), big);
                if (plen > sizeof(small)) buf = big.data();
                uint8_t pre[20];
                int pn = put_varint(pre, plen);
                pn += put_varint(pre + pn, (uint64_t)(i + 1));
                if (!w.add(pre, (uint32_t)pn, buf, plen)) {
                    w.close();
                    maxkey[t].push_back(i);      // rowid of the last row of the closed leaf
                    w.add(pre, (uint32_t)pn, buf, plen);
                }
            }
            w.close();
            maxkey[t].push_back(b);
            leaves[t] = w.leaf_pages;
        });
    }
    for (auto &x : th) x.join();
    std::vector<Child> kids;
    for (unsigned t = 0; t < T; ++t) {
        if (!db.place(runs[t], patches[t])) return 0;
        for (size_t k = 0; k < leaves[t].size(); ++k) kids.push_back({runs[t].first_page + leaves[t][k], maxkey[t][k]});
        runs[t].bytes.clear(); runs[t].bytes.shrink_to_fit();
    }
    return build_table_interior(db, kids);
}

// ---- small generic table (schema, stat, gzindex, comp, base, meta) ----------------------------------------
uint32_t build_small_table(Db &db, const std::vector<std::vector<Val>> &rows) {
    PageRun run;
    std::vector<Patch> patches;
    LeafWriter w(&run, &patches, 0x0D);
    std::vector<Child> kids_local;
    std::vector<int64_t> maxkey;
    std::vector<uint8_t> rec;
    for (size_t i = 0; i < rows.size(); ++i) {
        encode_record(rows[i].data(), (int)rows[i].size(), rec);
        uint8_t pre[20];
        int pn = put_varint(pre, rec.size());
        pn += put_varint(pre + pn, (uint64_t)(i + 1));
        if (!w.add(pre, (uint32_t)pn, rec.data(), rec.size())) {
            w.close();
            maxkey.push_back((int64_t)i);
            w.add(pre, (uint32_t)pn, rec.data(), rec.size());
        }
    }
    if (w.cur < 0) w.open();
    w.close();
    maxkey.push_back((int64_t)rows.size());
    if (!db.place(run, patches)) return 0;
    std::vector<Child> kids;
    for (size_t k = 0; k < w.leaf_pages.size(); ++k) kids.push_back({run.first_page + w.leaf_pages[k], maxkey[k]});
    return build_table_interior(db, kids);
}

// ---- an index b-tree from entries that are already in key order -------------------------------------------
// payload(i, out) = the record of the i-th entry (indexed columns + rowid).  The range is cut into contiguous
// runs, one per thread; leaf pages carry no page numbers, so the runs are built independently.  Between two
// leaves one entry moves up into the parent (an index interior cell IS an entry).
template <class Payload>
uint32_t build_index_from_sorted(Db &db, int64_t n, Payload payload) {
    if (n == 0) {
        PageRun run; std::vector<Patch> none;
        uint8_t *pg = run.new_page();
        pg[0] = 0x0A; put_be16(pg + 3, 0); put_be16(pg + 5, PAGE);
        if (!db.place(run, none)) return 0;
        return run.first_page;
    }
    unsigned T = worker_count(n);
    while (T > 1 && n / T < 8) --T;                                // every run needs a few entries
    struct RunOut { PageRun run; std::vector<Patch> patches; std::vector<uint32_t> leaves; std::vector<std::vector<uint8_t>> seps;
                    std::vector<uint8_t> tail_sep; };
    std::vector<RunOut> outs(T);
    std::vector<std::thread> th;
    for (unsigned r = 0; r < T; ++r)
        th.emplace_back([&, r] {
            RunOut &o = outs[r];
            const int64_t a = n * r / T;
            int64_t b = n * (r + 1) / T;
            if (r + 1 < T) --b;                                    // the run's last entry separates it from the next run
            LeafWriter w(&o.run, &o.patches, 0x0A);
            std::vector<uint8_t> pl;
            for (int64_t i = a; i < b; ++i) {
                payload(i, pl);
                uint8_t pre[10];
                const int pn = put_varint(pre, pl.size());
                if (w.add(pre, (uint32_t)pn, pl.data(), pl.size())) continue;
                // leaf full: this entry becomes the separator -- unless it is the run's last one, which would leave no
                // leaf to its right; then the closed leaf gives up its last entry instead and this one opens a new leaf
                w.close();
                if (i + 1 < b) { o.seps.push_back(pl); continue; }
                uint8_t *pg = o.run.page(w.leaf_pages.back());
                uint32_t nc = ((uint32_t)pg[3] << 8) | pg[4];
                std::vector<uint8_t> prev;
                payload(i - 1, prev);
                o.seps.push_back(prev);
                --nc;                                              // drop the last cell (its bytes stay as dead space)
                put_be16(pg + 3, nc);
                uint32_t cs = PAGE;
                for (uint32_t c = 0; c < nc; ++c) { const uint32_t x = ((uint32_t)pg[8 + 2 * c] << 8) | pg[8 + 2 * c + 1]; if (x < cs) cs = x; }
                put_be16(pg + 5, cs);
                w.add(pre, (uint32_t)pn, pl.data(), pl.size());
            }
            w.close();
            o.leaves = w.leaf_pages;
            if (r + 1 < T) payload(b, o.tail_sep);
        });
    for (auto &x : th) x.join();
    std::vector<uint32_t> kids;
    std::vector<std::vector<uint8_t>> seps;
    for (unsigned r = 0; r < T; ++r) {
        RunOut &o = outs[r];
        if (!db.place(o.run, o.patches)) return 0;
        for (size_t k = 0; k < o.leaves.size(); ++k) {
            kids.push_back(o.run.first_page + o.leaves[k]);
            if (k < o.seps.size()) seps.push_back(o.seps[k]);
        }
        if (r + 1 < T) seps.push_back(o.tail_sep);
        o.run.bytes.clear(); o.run.bytes.shrink_to_fit();
    }
    return build_index_interior(db, kids, seps);
}

// ---- UNIQUE index on the name column --------------------------------------------------------------------------
struct SortKey { uint64_t k0, k1; uint32_t idx; };

inline uint64_t be_prefix(const uint8_t *p, int64_t len) {
    uint64_t v = 0;
    const int n = len >= 8 ? 8 : (int)(len > 0 ? len : 0);
    for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * (7 - i));
    return v;
}

struct NameCmp {
    const uint8_t *names; const int64_t *off;
    // memcmp order, shorter first on a common prefix (SQLite BINARY collation), then rowid
    bool operator()(const SortKey &a, const SortKey &b) const {
        if (a.k0 != b.k0) return a.k0 < b.k0;
        if (a.k1 != b.k1) return a.k1 < b.k1;
        const int64_t la = off[a.idx + 1] - off[a.idx], lb = off[b.idx + 1] - off[b.idx];
        if (la > 16 && lb > 16) {
            const int64_t m = std::min(la, lb) - 16;
            const int c = memcmp(names + off[a.idx] + 16, names + off[b.idx] + 16, (size_t)m);
            if (c) return c < 0;
        }
        if (la != lb) return la < lb;
        return a.idx < b.idx;
    }
    bool equal_names(const SortKey &a, const SortKey &b) const {
        const int64_t la = off[a.idx + 1] - off[a.idx], lb = off[b.idx + 1] - off[b.idx];
        return la == lb && a.k0 == b.k0 && a.k1 == b.k1 && (la <= 16 || memcmp(names + off[a.idx] + 16, names + off[b.idx] + 16, (size_t)(la - 16)) == 0);
    }
};

// returns root page (0 on I/O error); *created = false when duplicate names forbid the UNIQUE index
uint32_t build_name_index(Db &db, const uint8_t *names, const int64_t *off, int64_t n, bool *created) {
    *created = true;
    if (n == 0) {
        PageRun run; std::vector<Patch> none;
        uint8_t *pg = run.new_page();
        pg[0] = 0x0A; put_be16(pg + 3, 0); put_be16(pg + 5, PAGE);
        if (!db.place(run, none)) return 0;
        return run.first_page;
    }
    const unsigned T = worker_count(n);
    StageTimer tm;
    std::vector<SortKey> keys((size_t)n);
    NameCmp cmp{names, off};
    {   // keys, in parallel
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) {
                    const int64_t l = off[i + 1] - off[i];
                    keys[(size_t)i] = {be_prefix(names + off[i], l), be_prefix(names + off[i] + 8, l - 8), (uint32_t)i};
                }
            });
        for (auto &x : th) x.join();
    }
    // sample sort: T buckets by splitters, every bucket sorted by its own thread
    std::vector<int64_t> bstart(T + 1, 0);
    std::vector<SortKey> sorted((size_t)n);
    if (T == 1) {
        sorted = keys;
        std::sort(sorted.begin(), sorted.end(), cmp);
        bstart[1] = n;
    } else {
        std::vector<SortKey> sample;
        const int64_t ns = std::min<int64_t>(n, (int64_t)T * 256);
        for (int64_t s = 0; s < ns; ++s) sample.push_back(keys[(size_t)(s * n / ns)]);
        std::sort(sample.begin(), sample.end(), cmp);
        std::vector<SortKey> split;
        for (unsigned t = 1; t < T; ++t) split.push_back(sample[(size_t)(t * sample.size() / T)]);
        std::vector<std::vector<int64_t>> cnt(T, std::vector<int64_t>(T, 0));
        std::vector<uint8_t> bucket((size_t)n);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) {
                    const unsigned b = (unsigned)(std::upper_bound(split.begin(), split.end(), keys[(size_t)i], cmp) - split.begin());
                    bucket[(size_t)i] = (uint8_t)b;
                    ++cnt[t][b];
                }
            });
        for (auto &x : th) x.join();
        th.clear();
        std::vector<std::vector<int64_t>> pos(T, std::vector<int64_t>(T, 0));
        int64_t acc = 0;
        for (unsigned b = 0; b < T; ++b) {
            bstart[b] = acc;
            for (unsigned t = 0; t < T; ++t) { pos[t][b] = acc; acc += cnt[t][b]; }
        }
        bstart[T] = n;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::vector<int64_t> p = pos[t];
                for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) sorted[(size_t)p[bucket[(size_t)i]]++] = keys[(size_t)i];
            });
        for (auto &x : th) x.join();
        th.clear();
        for (unsigned b = 0; b < T; ++b)
            th.emplace_back([&, b] { std::sort(sorted.begin() + bstart[b], sorted.begin() + bstart[b + 1], cmp); });
        for (auto &x : th) x.join();
    }
    keys.clear(); keys.shrink_to_fit();
    tm.lap("  index: sort");
    // duplicates: UNIQUE index cannot be created (reference: sqlite3_exec fails, error ignored, src/index.c:366)
    {
        std::atomic<bool> dup(false);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (int64_t i = std::max<int64_t>(1, n * t / T); i < n * (t + 1) / T && !dup; ++i)
                    if (cmp.equal_names(sorted[(size_t)i - 1], sorted[(size_t)i])) dup = true;
            });
        for (auto &x : th) x.join();
        if (dup) { *created = false; return 0; }
    }
    auto entry_payload = [&](int64_t i, std::vector<uint8_t> &out) {
        const SortKey &k = sorted[(size_t)i];
        const int64_t l = off[k.idx + 1] - off[k.idx];
        uint8_t ib[8];
        int il;
        const int it = int_serial((int64_t)k.idx + 1, ib, &il);
        uint8_t h[24];
        int hl = put_varint(h, (uint64_t)l * 2 + 13);
        hl += put_varint(h + hl, (uint64_t)it);
        uint8_t hv[4];
        const int hvn = put_varint(hv, (uint64_t)(hl + 1 <= 127 ? hl + 1 : hl + 2));
        out.resize((size_t)(hvn + hl + l + il));
        memcpy(out.data(), hv, (size_t)hvn);
        memcpy(out.data() + hvn, h, (size_t)hl);
        memcpy(out.data() + hvn + hl, names + off[k.idx], (size_t)l);
        memcpy(out.data() + hvn + hl + l, ib, (size_t)il);
    };
    const uint32_t root = build_index_from_sorted(db, n, entry_payload);
    tm.lap("  index: leaves");
    return root;
}

// ---- schema page + header ----------------------------------------------------------------------------------
struct SchemaRow { const char *type, *name, *tbl; uint32_t root; std::string sql; };

bool write_page1(Db &db, const std::vector<SchemaRow> &schema) {
    PageRun run;
    std::vector<Patch> patches;
    LeafWriter w(&run, &patches, 0x0D);
    w.hdr_off = 100;
    w.open();
    std::vector<uint8_t> rec;
    for (size_t i = 0; i < schema.size(); ++i) {
        const SchemaRow &s = schema[i];
        Val v[5] = {Val::text(s.type), Val::text(s.name), Val::text(s.tbl), Val::integer(s.root), Val::text(s.sql.data(), s.sql.size())};
        encode_record(v, 5, rec);
        uint8_t pre[20];
        int pn = put_varint(pre, rec.size());
        pn += put_varint(pre + pn, (uint64_t)(i + 1));
        if (!w.add(pre, (uint32_t)pn, rec.data(), rec.size())) { db.err = "schema does not fit page 1"; return false; }
    }
    w.close();
    if (run.npages() != 1) { db.err = "schema overflowed page 1"; return false; }
    uint8_t *p = run.page(0);
    memcpy(p, "SQLite format 3\0", 16);
    put_be16(p + 16, PAGE);
    p[18] = 1; p[19] = 1; p[20] = 0; p[21] = 64; p[22] = 32; p[23] = 32;
    put_be32(p + 24, 1);                        // file change counter
    put_be32(p + 28, db.next_page - 1);         // database size in pages
    put_be32(p + 40, 1);                        // schema cookie
    put_be32(p + 44, 4);                        // schema format
    put_be32(p + 56, 1);                        // UTF-8
    put_be32(p + 92, 1);                        // version-valid-for
    put_be32(p + 96, 3045001);                  // SQLITE_VERSION_NUMBER of the writer (informational)
    return db.pwrite_all(p, PAGE, 0);
}

bool open_db(Db &db, const char *path) {
    db.fd = ::open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (db.fd < 0) { db.err = std::string("cannot create ") + path + ": " + strerror(errno); return false; }
    return true;
}

// the gzindex rows of a BGZF input: one row per field, exactly as pyfastx_gzip_index_export writes them
// (src/util.c:442-540); points sit at gzip member starts (deflate data right after the member header, bits = 0)
// and need no window data (members are independent), import checks of src/util.c:575-609 hold:
// window_size >= 32768, spacing >= window_size, compressed_size == the file's size.
void gzindex_rows(const fxg_gzindex *gz, std::vector<std::vector<Val>> &rows, std::vector<std::vector<uint8_t>> &store) {
    if (!gz || gz->npoints <= 0) return;
    auto add = [&](const void *p, size_t n) {
        store.emplace_back((const uint8_t *)p, (const uint8_t *)p + n);
        rows.push_back({Val::null(), Val::blob(nullptr, 0)});    // pointer fixed up below (store may reallocate)
    };
    const uint8_t version = 1, flags = 0;
    const uint64_t csz = (uint64_t)gz->compressed_size, usz = (uint64_t)gz->uncompressed_size;
    const uint32_t spacing = gz->spacing, wsz = gz->window_size, np = (uint32_t)gz->npoints;
    add("GZIDX", 5); add(&version, 1); add(&flags, 1); add(&csz, 8); add(&usz, 8); add(&spacing, 4); add(&wsz, 4); add(&np, 4);
    for (int64_t i = 0; i < gz->npoints; ++i) {
        const uint64_t c = (uint64_t)gz->cmp_offset[i], u = (uint64_t)gz->uncmp_offset[i];
        const uint8_t bits = gz->bits ? gz->bits[i] : 0, has_data = gz->has_data ? gz->has_data[i] : 0;
        add(&c, 8); add(&u, 8); add(&bits, 1); add(&has_data, 1);
    }
    // window data of the points that have one, in point order (src/util.c:514-527)
    if (gz->has_data && gz->windows) {
        int64_t k = 0;
        for (int64_t i = 0; i < gz->npoints; ++i)
            if (gz->has_data[i]) { add(gz->windows + (size_t)k * gz->window_size, gz->window_size); ++k; }
    }
    for (size_t i = 0; i < rows.size(); ++i) rows[i][1] = Val::blob(store[i].data(), store[i].size());
}

const char *SEQ_SQL = "CREATE TABLE seq ( \n\t\t\tID INTEGER PRIMARY KEY, --seq identifier\n \t\t\tchrom TEXT, --seq name\n \t\t\tboff INTEGER, --seq offset start\n \t\t\tblen INTEGER, --seq byte length\n \t\t\tslen INTEGER, --seq length\n \t\t\tllen INTEGER, --line length\n \t\t\telen INTEGER, --end length\n \t\t\tnorm INTEGER, --line with the same length or not\n \t\t\tdlen INTEGER --description header line length\n \t\t)";

}  // namespace

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" int fxg_fxi_write_fasta(const char *path, const fxg_fasta_row *rows, int64_t n_rows, const uint8_t *names,
                                   const int64_t *name_off, int64_t total_slen, const fxg_gzindex *gz,
                                   const fxg_comp_row *comp, int64_t n_comp) {
    if (!path || n_rows < 0 || (n_rows && (!rows || !names || !name_off)) || n_comp < 0 || (n_comp && !comp)) {
        fxg_set_error("invalid argument: fxg_fxi_write_fasta");
        return FXG_EINVAL;
    }
    Db db;
    if (!open_db(db, path)) { fxg_set_error("%s", db.err.c_str()); return FXG_EIO; }
    auto enc = [&](int64_t i, uint8_t *buf, size_t cap, std::vector<uint8_t> &big) -> uint64_t {
        const fxg_fasta_row &r = rows[i];
        const int64_t nl = name_off[i + 1] - name_off[i];
        uint8_t body[8 * 7];
        uint8_t hdr[9 + 9 + 7];
        int hl = 0, bl = 0, b;
        hdr[hl++] = 0;                                                       // ID: NULL (rowid alias)
        hl += put_varint(hdr + hl, (uint64_t)nl * 2 + 13);                   // chrom TEXT
        const int64_t iv[7] = {r.boff, r.blen, r.slen, r.llen, (int64_t)r.elen, (int64_t)r.norm, (int64_t)r.dlen};
        for (int k = 0; k < 7; ++k) { hdr[hl++] = (uint8_t)int_serial(iv[k], body + bl, &b); bl += b; }
        const int hs = hl + 1;
        const uint64_t plen = (uint64_t)hs + (uint64_t)nl + (uint64_t)bl;
        uint8_t *o = buf;
        if (plen > cap) { big.resize((size_t)plen); o = big.data(); }
        o[0] = (uint8_t)hs;
        memcpy(o + 1, hdr, (size_t)hl);
        memcpy(o + hs, names + name_off[i], (size_t)nl);
        memcpy(o + hs + nl, body, (size_t)bl);
        return plen;
    };
    StageTimer tm;
    const uint32_t seq_root = build_big_table(db, n_rows, enc);
    tm.lap("seq table");
    bool ok = seq_root != 0;
    uint32_t stat_root = 0, comp_root = 0, gz_root = 0, idx_root = 0, seqidx_root = 0;
    bool idx_created = false;
    if (ok) {
        std::vector<std::vector<Val>> stat = {{Val::integer(n_rows), Val::integer(total_slen), Val::null(), Val::null(), Val::null(), Val::null()}};
        stat_root = build_small_table(db, stat);
        ok = stat_root != 0;
    }
    if (ok) {
        auto cenc = [&](int64_t i, uint8_t *buf, size_t, std::vector<uint8_t> &) -> uint64_t {
            uint8_t body[24];
            uint8_t hdr[8];
            int hl = 0, bl = 0, b;
            hdr[hl++] = 0;
            const int64_t iv[3] = {comp[i].seqid, comp[i].abc, comp[i].num};
            for (int k = 0; k < 3; ++k) { hdr[hl++] = (uint8_t)int_serial(iv[k], body + bl, &b); bl += b; }
            buf[0] = (uint8_t)(hl + 1);
            memcpy(buf + 1, hdr, (size_t)hl);
            memcpy(buf + 1 + hl, body, (size_t)bl);
            return (uint64_t)(1 + hl + bl);
        };
        comp_root = build_big_table(db, n_comp, cenc);
        ok = comp_root != 0;
        if (ok && n_comp > 0) {
            // CREATE INDEX seqidx ON comp (seqid)  (src/fasta.c:953).  Entries ordered by (seqid, rowid): rows come in
            // seqid order except the whole-file rows (seqid 0), which the reference appends LAST -- they sort first.
            int64_t nz = 0;
            while (nz < n_comp && comp[n_comp - 1 - nz].seqid == 0) ++nz;
            bool sorted_ok = true;
            for (int64_t i = 1; i < n_comp - nz && sorted_ok; ++i) sorted_ok = comp[i - 1].seqid <= comp[i].seqid && comp[i].seqid != 0;
            if (sorted_ok) {
                auto pay = [&](int64_t i, std::vector<uint8_t> &out) {
                    const int64_t row = i < nz ? n_comp - nz + i : i - nz;          // 0-based row of the i-th entry
                    uint8_t b1[8], b2[8];
                    int l1, l2;
                    const int t1 = int_serial(comp[row].seqid, b1, &l1), t2 = int_serial(row + 1, b2, &l2);
                    out.resize((size_t)(3 + l1 + l2));
                    out[0] = 3; out[1] = (uint8_t)t1; out[2] = (uint8_t)t2;
                    memcpy(out.data() + 3, b1, (size_t)l1);
                    memcpy(out.data() + 3 + l1, b2, (size_t)l2);
                };
                seqidx_root = build_index_from_sorted(db, n_comp, pay);
                ok = seqidx_root != 0;
            }
        }
    }
    if (ok) {
        std::vector<std::vector<Val>> grows;
        std::vector<std::vector<uint8_t>> store;
        store.reserve(gz && gz->npoints > 0 ? (size_t)gz->npoints * 5 + 8 : 0);
        gzindex_rows(gz, grows, store);
        gz_root = build_small_table(db, grows);
        ok = gz_root != 0;
    }
    if (ok) {
        tm.lap("small tables");
        idx_root = build_name_index(db, names, name_off, n_rows, &idx_created);
        tm.lap("name index");
        ok = idx_root != 0 || !idx_created;
        if (!db.err.empty()) ok = false;
    }
    if (ok) {
        std::vector<SchemaRow> schema = {
            {"table", "seq", "seq", seq_root, SEQ_SQL},
            {"table", "stat", "stat", stat_root, "CREATE TABLE stat ( \n\t\t\tseqnum INTEGER, --total seq counts \n \t\t\tseqlen INTEGER, --total seq length \n \t\t\tavglen REAL, --average seq length \n \t\t\tmedlen REAL, --median seq length \n \t\t\tn50 INTEGER, --N50 seq length \n \t\t\tl50 INTEGER --L50 seq count \n \t\t)"},
            {"table", "comp", "comp", comp_root, "CREATE TABLE comp ( \n\t\t\tID INTEGER PRIMARY KEY, \n \t\t\tseqid INTEGER, --seq id \n \t\t\tabc INTEGER, --seq letter \n \t\t\tnum INTEGER -- letter count \n \t\t)"},
            {"table", "gzindex", "gzindex", gz_root, "CREATE TABLE gzindex ( \n\t\t\tID INTEGER PRIMARY KEY, \n \t\t\tcontent BLOB \n \t\t)"},
        };
        if (idx_created) schema.push_back({"index", "chromidx", "seq", idx_root, "CREATE UNIQUE INDEX chromidx ON seq (chrom)"});
        if (seqidx_root) schema.push_back({"index", "seqidx", "comp", seqidx_root, "CREATE INDEX seqidx ON comp (seqid)"});
        ok = write_page1(db, schema);
    }
    ::close(db.fd);
    if (!ok) { fxg_set_error("fxi write failed: %s", db.err.c_str()); ::unlink(path); return FXG_EIO; }
    return FXG_OK;
}

extern "C" int fxg_fxi_write_fastq(const char *path, const fxg_fastq_row *rows, int64_t n_rows, const uint8_t *names,
                                   const int64_t *name_off, int64_t n_lines, int64_t total_size, const fxg_gzindex *gz,
                                   const fxg_fastq_meta *meta) {
    if (!path || n_rows < 0 || (n_rows && (!rows || !names || !name_off))) {
        fxg_set_error("invalid argument: fxg_fxi_write_fastq");
        return FXG_EINVAL;
    }
    Db db;
    if (!open_db(db, path)) { fxg_set_error("%s", db.err.c_str()); return FXG_EIO; }
    auto enc = [&](int64_t i, uint8_t *buf, size_t cap, std::vector<uint8_t> &big) -> uint64_t {
        const fxg_fastq_row &r = rows[i];
        const int64_t nl = name_off[i + 1] - name_off[i];
        uint8_t body[8 * 4];
        uint8_t hdr[9 + 9 + 4];
        int hl = 0, bl = 0, b;
        hdr[hl++] = 0;                                                       // ID
        hl += put_varint(hdr + hl, (uint64_t)nl * 2 + 13);                   // name TEXT
        const int64_t iv[4] = {(int64_t)r.dlen, r.rlen, r.soff, r.qoff};
        for (int k = 0; k < 4; ++k) { hdr[hl++] = (uint8_t)int_serial(iv[k], body + bl, &b); bl += b; }
        const int hs = hl + 1;
        const uint64_t plen = (uint64_t)hs + (uint64_t)nl + (uint64_t)bl;
        uint8_t *o = buf;
        if (plen > cap) { big.resize((size_t)plen); o = big.data(); }
        o[0] = (uint8_t)hs;
        memcpy(o + 1, hdr, (size_t)hl);
        memcpy(o + hs, names + name_off[i], (size_t)nl);
        memcpy(o + hs + nl, body, (size_t)bl);
        return plen;
    };
    const uint32_t read_root = build_big_table(db, n_rows, enc);
    bool ok = read_root != 0;
    uint32_t gz_root = 0, stat_root = 0, base_root = 0, meta_root = 0, idx_root = 0;
    bool idx_created = false;
    if (ok) {
        std::vector<std::vector<Val>> grows;
        std::vector<std::vector<uint8_t>> store;
        store.reserve(gz && gz->npoints > 0 ? (size_t)gz->npoints * 5 + 8 : 0);
        gzindex_rows(gz, grows, store);
        gz_root = build_small_table(db, grows);
        ok = gz_root != 0;
    }
    if (ok) {
        const int64_t counts = n_lines / 4;                                  // fastq.c:159
        const double avg = counts ? (double)total_size * 1.0 / (double)counts : 0.0 / 0.0;   // fastq.c:161
        std::vector<std::vector<Val>> stat = {{Val::integer(counts), Val::integer(total_size), Val::real(avg)}};
        stat_root = build_small_table(db, stat);
        ok = stat_root != 0;
    }
    if (ok) {
        std::vector<std::vector<Val>> b, m;
        if (meta) {
            b.push_back({Val::integer(meta->a), Val::integer(meta->c), Val::integer(meta->g), Val::integer(meta->t), Val::integer(meta->n)});
            m.push_back({Val::integer(meta->maxlen), Val::integer(meta->minlen), Val::integer(meta->minqs), Val::integer(meta->maxqs), Val::integer(meta->phred)});
        }
        base_root = build_small_table(db, b);
        meta_root = ok && base_root ? build_small_table(db, m) : 0;
        ok = base_root != 0 && meta_root != 0;
    }
    if (ok) {
        idx_root = build_name_index(db, names, name_off, n_rows, &idx_created);
        ok = idx_root != 0 || !idx_created;
        if (!db.err.empty()) ok = false;
    }
    if (ok) {
        std::vector<SchemaRow> schema = {
            {"table", "read", "read", read_root, "CREATE TABLE read ( \n\t\t\tID INTEGER PRIMARY KEY, --read id \n \t\t\tname TEXT, --read name \n \t\t\tdlen INTEGER, --description length \n \t\t\trlen INTEGER, --read length \n \t\t\tsoff INTEGER, --read seq offset \n \t\t\tqoff INTEGER --read qual offset \n \t\t)"},
            {"table", "gzindex", "gzindex", gz_root, "CREATE TABLE gzindex ( \n\t\t\tID INTEGER PRIMARY KEY, \n \t\t\tcontent BLOB \n \t\t)"},
            {"table", "stat", "stat", stat_root, "CREATE TABLE stat ( \n\t\t\tcounts INTEGER, --read counts \n \t\t\tsize INTEGER, --all read length \n \t\t\tavglen REAL --average read length \n \t\t)"},
            {"table", "base", "base", base_root, "CREATE TABLE base ( \n\t\t\ta INTEGER,  \n \t\t\tc INTEGER,  \n \t\t\tg INTEGER,  \n \t\t\tt INTEGER,  \n \t\t\tn INTEGER  \n \t\t)"},
            {"table", "meta", "meta", meta_root, "CREATE TABLE meta ( \n\t\t\tmaxlen INTEGER, --maximum read length \n \t\t\tminlen INTEGER, --minimum read length \n \t\t\tminqs INTEGER, --max quality score \n \t\t\tmaxqs INTEGER, --min quality score \n \t\t\tphred INTEGER --phred value \n \t\t)"},
        };
        if (idx_created) schema.push_back({"index", "readidx", "read", idx_root, "CREATE UNIQUE INDEX readidx ON read (name)"});
        ok = write_page1(db, schema);
    }
    ::close(db.fd);
    if (!ok) { fxg_set_error("fxi write failed: %s", db.err.c_str()); ::unlink(path); return FXG_EIO; }
    return FXG_OK;
}
