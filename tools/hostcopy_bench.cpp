// hostcopy_bench.cpp -- how fast can a window of a page-cache-resident file reach a (pinned-like) staging buffer?
// Compares, per 32 MB slot and with T threads: memcpy from a fresh file mapping (what upload() does for a
// video_source_yuv_file window), the same after madvise(MADV_POPULATE_READ), and pread() straight into the slot.
//   g++ -O2 -pthread -o /tmp/hostcopy_bench tools/hostcopy_bench.cpp && /tmp/hostcopy_bench
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F>
static void par(size_t n, int threads, F f) {
    const size_t part = ((n + threads - 1) / threads + 4095) / 4096 * 4096;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t o = (size_t)t * part;
        if (o >= n) break;
        pool.emplace_back([=]() { f(o, std::min(part, n - o)); });
    }
    f(0, std::min(part, n));
    for (auto &th : pool) th.join();
}
int main() {
    const size_t total = (size_t)1536 << 20, slot = (size_t)32 << 20;
    const char *path = "/dev/shm/hostcopy_bench.bin";
    int fd = open(path, O_RDWR | O_CREAT, 0600);
    if (ftruncate(fd, total) != 0) return 1;
    {
        char *w = (char *)mmap(0, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        memset(w, 1, total);
        munmap(w, total);
    }
    char *pin = (char *)aligned_alloc(4096, slot);
    memset(pin, 0, slot);
    for (int threads : {4, 8, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                const char *src = (const char *)mmap(0, total, PROT_READ, MAP_SHARED, fd, 0);
                const double t0 = now();
                for (size_t off = 0; off < total; off += slot) {
                    const size_t len = std::min(slot, total - off);
                    if (mode == 0) par(len, threads, [=](size_t o, size_t l) { memcpy(pin + o, src + off + o, l); });
                    if (mode == 1)
                        par(len, threads, [=](size_t o, size_t l) {
                            madvise((void *)(src + off + o), l, MADV_POPULATE_READ);
                            memcpy(pin + o, src + off + o, l);
                        });
                    if (mode == 2)
                        par(len, threads, [=](size_t o, size_t l) {
                            size_t done = 0;
                            while (done < l) {
                                ssize_t r = pread(fd, pin + o + done, l - done, (off_t)(off + o + done));
                                if (r <= 0) break;
                                done += (size_t)r;
                            }
                        });
                }
                best = std::min(best, now() - t0);
                munmap((void *)src, total);
            }
            printf("threads %2d  %-28s %7.1f ms  %5.1f GB/s\n", threads,
                   mode == 0 ? "memcpy from fresh mapping" : (mode == 1 ? "populate_read + memcpy" : "pread into the slot"), best * 1e3,
                   total / 1e9 / best);
        }
    }
    close(fd);
    unlink(path);
    return 0;
}
