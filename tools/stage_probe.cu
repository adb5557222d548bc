// Host-staging probe (run on the GPU box): how fast can file bytes reach HBM?
//   nvcc -O2 -o /tmp/stage_probe tools/stage_probe.cu && /tmp/stage_probe /dev/shm/x.bin
// Variants: (a) N pread threads -> pinned chunk (no DMA); (b) the same overlapped with cudaMemcpyAsync;
// (c) mmap + cudaHostRegister per chunk + DMA from the page cache; (d) cudaMemcpy from the pageable mapping.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#include <sched.h>
static int g_aff = -1;          // -1 none, 0 / 1: NUMA node whose CPUs the readers are pinned to (2 x 32 cores x 2 HT layout)
static void pin_self() {
    if (g_aff < 0) return;
    cpu_set_t cs;
    CPU_ZERO(&cs);
    for (int c = 0; c < 128; ++c) if (((c % 64) / 32) == g_aff) CPU_SET(c, &cs);
    sched_setaffinity(0, sizeof(cs), &cs);
}
static void pread_par(int fd, char *dst, int64_t off, int64_t len, int nt) {
    std::vector<std::thread> th;
    const int64_t per = ((len + nt - 1) / nt + 4095) & ~4095ll;
    for (int i = 0; i < nt; ++i) {
        const int64_t lo = (int64_t)i * per;
        if (lo >= len) break;
        const int64_t cnt = lo + per <= len ? per : len - lo;
        th.emplace_back([=] {
            pin_self();
            int64_t done = 0;
            while (done < cnt) {
                ssize_t r = pread(fd, dst + lo + done, (size_t)(cnt - done), (off_t)(off + lo + done));
                if (r <= 0) return;
                done += r;
            }
        });
    }
    for (auto &t : th) t.join();
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "/dev/shm/stage_probe.bin";
    int64_t want = argc > 2 ? atoll(argv[2]) : (int64_t)8 << 30;
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        fd = open(path, O_RDWR | O_CREAT, 0600);
        std::vector<char> buf(64 << 20, 'A');
        for (int64_t o = 0; o < want; o += (int64_t)buf.size()) if (write(fd, buf.data(), buf.size()) < 0) return 1;
        close(fd);
        fd = open(path, O_RDONLY);
    }
    struct stat st;
    fstat(fd, &st);
    const int64_t n = st.st_size;
    printf("file %s: %.2f GB, hw threads %u\n", path, n / 1e9, std::thread::hardware_concurrency());
    char *d;
    cudaMalloc(&d, n);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const int64_t CH = (int64_t)256 << 20;
    char *pin[4];
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; ++i) { cudaHostAlloc(&pin[i], CH, cudaHostAllocDefault); memset(pin[i], 1, CH); cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); }

    // (0) pure DMA from pinned
    {
        double t0 = now();
        for (int64_t o = 0; o < n; o += CH) cudaMemcpyAsync(d + o, pin[(o / CH) & 3], (size_t)(n - o < CH ? n - o : CH), cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        printf("(0) DMA only from pinned: %.1f GB/s\n", n / (now() - t0) / 1e9);
    }
    // (a) pread only
    for (int nt : {1, 16, 48}) {
        double t0 = now();
        int64_t lim = n < ((int64_t)4 << 30) ? n : ((int64_t)4 << 30);
        if (nt == 1) lim = (int64_t)1 << 30;
        for (int64_t o = 0; o < lim; o += CH) pread_par(fd, pin[(o / CH) & 3], o, lim - o < CH ? lim - o : CH, nt);
        printf("(a) pread only, %3d threads: %.1f GB/s\n", nt, lim / (now() - t0) / 1e9);
    }
    // (a2) pread with persistent threads pulling 4 MiB pieces from an atomic counter (no per-chunk spawn)
    for (int nt : {16}) {
        double t0 = now();
        std::atomic<int64_t> next(0);
        const int64_t P = (int64_t)4 << 20;
        int64_t lim = n < ((int64_t)4 << 30) ? n : ((int64_t)4 << 30);
        std::vector<std::thread> th;
        for (int i = 0; i < nt; ++i) th.emplace_back([&] {
            for (;;) {
                int64_t o = next.fetch_add(P);
                if (o >= lim) return;
                int64_t len = lim - o < P ? lim - o : P;
                char *dst = pin[(o / CH) & 3] + (o % CH);
                int64_t done = 0;
                while (done < len) { ssize_t r = pread(fd, dst + done, len - done, o + done); if (r <= 0) return; done += r; }
            }
        });
        for (auto &t : th) t.join();
        printf("(a2) pread persistent, %3d threads, 4 MiB pieces: %.1f GB/s\n", nt, lim / (now() - t0) / 1e9);
    }
    // (b) pread + DMA overlapped, 4 buffers
    {
        char bus[64] = {0};
        cudaDeviceGetPCIBusId(bus, 64, 0);
        for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
        char pth[256];
        snprintf(pth, sizeof pth, "/sys/bus/pci/devices/%s/numa_node", bus);
        FILE *nf = fopen(pth, "r");
        int node = -9;
        if (nf) { if (fscanf(nf, "%d", &node) != 1) node = -9; fclose(nf); }
        printf("GPU 0 at %s, numa_node %d\n", bus, node);
    }
    for (int aff : {-1, 0, 1}) for (int nt : {8, 12, 16, 20, 24, 32}) {
        g_aff = aff;
        if (aff >= 0) {
            double t0 = now();
            int64_t lim = n < ((int64_t)4 << 30) ? n : ((int64_t)4 << 30);
            for (int64_t o = 0; o < lim; o += CH) pread_par(fd, pin[(o / CH) & 3], o, lim - o < CH ? lim - o : CH, nt);
            printf("(a) pread only, aff %d, %3d threads: %.1f GB/s\n", aff, nt, lim / (now() - t0) / 1e9);
        }
        double t0 = now();
        int k = 0;
        for (int64_t o = 0; o < n; o += CH, k = (k + 1) & 3) {
            const int64_t len = n - o < CH ? n - o : CH;
            cudaEventSynchronize(ev[k]);
            pread_par(fd, pin[k], o, len, nt);
            cudaMemcpyAsync(d + o, pin[k], (size_t)len, cudaMemcpyHostToDevice, s);
            cudaEventRecord(ev[k], s);
        }
        cudaStreamSynchronize(s);
        printf("(b) aff %d pread(%d) + DMA overlapped: %.1f GB/s (%.3f s)\n", aff, nt, n / (now() - t0) / 1e9, now() - t0);
    }
    // (b2) fine-grained: persistent readers fill 16 MiB pieces; a dispatcher issues the DMA of each piece as it completes
    g_aff = -1;
    for (int nt : {16}) {
        const int64_t P = (int64_t)16 << 20;
        const int64_t RING = 4 * CH;                        // pinned ring = the four buffers (not contiguous: piece -> buffer)
        const int64_t np = (n + P - 1) / P;
        std::vector<std::atomic<int>> ready(np);
        for (auto &r : ready) r = 0;
        std::atomic<int64_t> next(0), dma_done(0);
        double t0 = now();
        std::vector<std::thread> th;
        for (int i = 0; i < nt; ++i) th.emplace_back([&] {
            for (;;) {
                int64_t p = next.fetch_add(1);
                if (p >= np) return;
                // ring slot free?  piece p reuses the slot of piece p - RING/P
                while (p - dma_done.load(std::memory_order_acquire) >= RING / P) std::this_thread::yield();
                const int64_t o = p * P, len = n - o < P ? n - o : P;
                char *dst = pin[((o % RING) / CH)] + (o % CH);
                int64_t done = 0;
                while (done < len) { ssize_t r = pread(fd, dst + done, len - done, o + done); if (r <= 0) break; done += r; }
                ready[p].store(1, std::memory_order_release);
            }
        });
        std::vector<cudaEvent_t> evs(np);
        int64_t issued = 0, retired = 0;
        while (retired < np) {
            while (issued < np && ready[issued].load(std::memory_order_acquire)) {
                const int64_t o = issued * P, len = n - o < P ? n - o : P;
                cudaMemcpyAsync(d + o, pin[((o % RING) / CH)] + (o % CH), (size_t)len, cudaMemcpyHostToDevice, s);
                cudaEventCreateWithFlags(&evs[issued], cudaEventDisableTiming);
                cudaEventRecord(evs[issued], s);
                ++issued;
            }
            while (retired < issued && cudaEventQuery(evs[retired]) == cudaSuccess) { cudaEventDestroy(evs[retired]); ++retired; dma_done.store(retired, std::memory_order_release); }
            if (issued < np && !ready[issued].load(std::memory_order_acquire)) std::this_thread::yield();
        }
        for (auto &t : th) t.join();
        cudaStreamSynchronize(s);
        printf("(b2) persistent readers(%d) + per-piece DMA: %.1f GB/s (%.3f s)\n", nt, n / (now() - t0) / 1e9, now() - t0);
    }
    // (c) mmap + register per chunk
    {
        char *m = (char *)mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
        if (m != MAP_FAILED) {
            double t0 = now(), treg = 0;
            int64_t lim = n < ((int64_t)2 << 30) ? n : ((int64_t)2 << 30);
            bool ok = true;
            for (int64_t o = 0; o < lim && ok; o += CH) {
                const int64_t len = lim - o < CH ? lim - o : CH;
                double a = now();
                cudaError_t e = cudaHostRegister(m + o, len, cudaHostRegisterReadOnly);
                if (e != cudaSuccess) { cudaGetLastError(); e = cudaHostRegister(m + o, len, cudaHostRegisterDefault); }
                treg += now() - a;
                if (e != cudaSuccess) { printf("(c) cudaHostRegister failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); ok = false; break; }
                cudaMemcpyAsync(d + o, m + o, (size_t)len, cudaMemcpyHostToDevice, s);
                cudaStreamSynchronize(s);
                cudaHostUnregister(m + o);
            }
            if (ok) printf("(c) mmap + register + DMA (serial): %.1f GB/s, register alone %.1f GB/s\n", lim / (now() - t0) / 1e9, lim / treg / 1e9);
            // (d) pageable
            t0 = now();
            int64_t lim2 = n < ((int64_t)2 << 30) ? n : ((int64_t)2 << 30);
            cudaMemcpy(d, m, lim2, cudaMemcpyHostToDevice);
            printf("(d) cudaMemcpy from the pageable mapping: %.1f GB/s\n", lim2 / (now() - t0) / 1e9);
            munmap(m, n);
        }
    }
    return 0;
}
