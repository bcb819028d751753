// Host->device staging probe: pinned DMA bandwidth and multi-threaded staging copy bandwidth.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv)
{
    size_t n = 43u << 20;
    uint8_t *pin, *dev, *src = (uint8_t*)malloc(n);
    memset(src, 1, n);
    cudaMallocHost(&pin, n);
    cudaMalloc(&dev, n);
    memset(pin, 2, n);
    for (int rep = 0; rep < 3; rep++)
    {
        double t0 = now();
        cudaMemcpy(dev, pin, n, cudaMemcpyHostToDevice);
        double t1 = now();
        printf("pinned H2D %zu MB: %.3f ms = %.1f GB/s\n", n >> 20, t1 - t0, n / (t1 - t0) / 1e6);
    }
    for (int T : {1, 2, 4, 8, 16})
    {
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++)
        {
            double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] { size_t a = n * t / T, b = n * (t + 1) / T; memcpy(pin + a, src + a, b - a); });
            for (auto& x : th) x.join();
            best = std::min(best, now() - t0);
        }
        printf("staging memcpy %d threads: %.3f ms = %.1f GB/s\n", T, best, n / best / 1e6);
    }
    if (argc > 1)
    {
        int fd = open(argv[1], O_RDONLY);
        size_t fsz = lseek(fd, 0, SEEK_END);
        if (fsz > n) fsz = n;
        for (int T : {1, 4, 8, 16})
        {
            double best = 1e9;
            for (int rep = 0; rep < 3; rep++)
            {
                double t0 = now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back([&, t] { size_t a = fsz * t / T, b = fsz * (t + 1) / T; while (a < b) { ssize_t g = pread(fd, pin + a, b - a, a); if (g <= 0) break; a += g; } });
                for (auto& x : th) x.join();
                best = std::min(best, now() - t0);
            }
            printf("pread %s %d threads: %.3f ms = %.1f GB/s\n", argv[1], T, best, fsz / best / 1e6);
        }
        void* m = mmap(nullptr, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
        for (int T : {1, 4, 8})
        {
            double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] { size_t a = fsz * t / T, b = fsz * (t + 1) / T; memcpy(pin + a, (uint8_t*)m + a, b - a); });
            for (auto& x : th) x.join();
            double t1 = now();
            printf("mmap memcpy (warm mapping after first) %d threads: %.3f ms = %.1f GB/s\n", T, t1 - t0, fsz / (t1 - t0) / 1e6);
        }
    }
    // write-combined pinned buffer
    uint8_t* wc;
    cudaHostAlloc(&wc, n, cudaHostAllocWriteCombined);
    for (int T : {4, 8})
    {
        double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] { size_t a = n * t / T, b = n * (t + 1) / T; memcpy(wc + a, src + a, b - a); });
        for (auto& x : th) x.join();
        double t1 = now();
        cudaMemcpy(dev, wc, n, cudaMemcpyHostToDevice);
        double t2 = now();
        printf("WC staging %d threads: %.3f ms; H2D from WC %.3f ms = %.1f GB/s\n", T, t1 - t0, t2 - t1, n / (t2 - t1) / 1e6);
    }
    return 0;
}
