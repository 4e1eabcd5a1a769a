// Probe (not part of the product): which ways of page-locking a file mapping does this platform accept, and how fast is
// H2D from each of them next to a cudaHostAlloc buffer?   usage: pin_probe <file>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void h2d_rate(const char* what, void* dev, const void* host, size_t n) {
    cudaStream_t s; cudaStreamCreate(&s);
    cudaMemcpyAsync(dev, host, n, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);
    double t = now();
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(dev, host, n, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    t = (now() - t) / 3;
    printf("  H2D %-40s %.1f GB/s (%s)\n", what, n / t / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaStreamDestroy(s);
}

int main(int argc, char** argv) {
    const char* path = argv[1];
    int a = 0, b = 0;
    cudaDeviceGetAttribute(&a, cudaDevAttrHostRegisterSupported, 0);
    cudaDeviceGetAttribute(&b, cudaDevAttrHostRegisterReadOnlySupported, 0);
    int c = 0; cudaDeviceGetAttribute(&c, cudaDevAttrPageableMemoryAccess, 0);
    int d = 0; cudaDeviceGetAttribute(&d, cudaDevAttrPageableMemoryAccessUsesHostPageTables, 0);
    printf("HostRegisterSupported %d ReadOnlySupported %d PageableMemoryAccess %d UsesHostPageTables %d\n", a, b, c, d);
    int fd = open(path, O_RDONLY);
    struct stat st; fstat(fd, &st);
    size_t n = std::min<size_t>(st.st_size, size_t(1) << 30) & ~size_t(4095);
    void* dev; cudaMalloc(&dev, n);
    struct { const char* name; int prot, flags, oflag; unsigned reg; } tries[] = {
        {"PROT_READ MAP_SHARED  reg RO|portable", PROT_READ, MAP_SHARED, O_RDONLY, cudaHostRegisterReadOnly | cudaHostRegisterPortable},
        {"PROT_READ MAP_SHARED  reg RO", PROT_READ, MAP_SHARED, O_RDONLY, cudaHostRegisterReadOnly},
        {"PROT_READ MAP_SHARED  reg default", PROT_READ, MAP_SHARED, O_RDONLY, cudaHostRegisterDefault},
        {"PROT_READ MAP_PRIVATE reg RO", PROT_READ, MAP_PRIVATE, O_RDONLY, cudaHostRegisterReadOnly},
        {"PROT_RW   MAP_PRIVATE reg default", PROT_READ | PROT_WRITE, MAP_PRIVATE, O_RDONLY, cudaHostRegisterDefault},
        {"PROT_RW   MAP_PRIVATE reg RO", PROT_READ | PROT_WRITE, MAP_PRIVATE, O_RDONLY, cudaHostRegisterReadOnly},
        {"PROT_RW   MAP_SHARED(O_RDWR) reg default", PROT_READ | PROT_WRITE, MAP_SHARED, O_RDWR, cudaHostRegisterDefault},
    };
    for (auto& t : tries) {
        int f = open(path, t.oflag);
        if (f < 0) { printf("%s: open failed\n", t.name); continue; }
        void* p = mmap(nullptr, n, t.prot, t.flags | MAP_POPULATE, f, 0);
        if (p == MAP_FAILED) { printf("%s: mmap failed\n", t.name); close(f); continue; }
        double t0 = now();
        cudaError_t e = cudaHostRegister(p, n, t.reg);
        double dt = now() - t0;
        printf("%s: %s (%.1f ms for %.2f GB)\n", t.name, cudaGetErrorString(e), dt * 1e3, n / 1e9);
        cudaGetLastError();
        if (e == cudaSuccess) { h2d_rate("from the registered mapping", dev, p, n); cudaHostUnregister(p); }
        munmap(p, n); close(f);
    }
    // pageable mapping straight into cudaMemcpyAsync (the driver stages it)
    void* p = mmap(nullptr, n, PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
    h2d_rate("from the pageable mapping (driver staging)", dev, p, n);
    // pinned buffer + memcpy by T threads
    void* pin; cudaHostAlloc(&pin, n, cudaHostAllocDefault);
    for (int T : {1, 4, 8, 16, 32}) {
        double t0 = now();
        std::vector<std::thread> th;
        for (int k = 0; k < T; ++k) th.emplace_back([&, k] { size_t lo = n / T * k, hi = k == T - 1 ? n : n / T * (k + 1); memcpy((char*)pin + lo, (char*)p + lo, hi - lo); });
        for (auto& x : th) x.join();
        double dt = now() - t0;
        printf("  memcpy mapping -> pinned, %2d threads: %.1f GB/s\n", T, n / dt / 1e9);
    }
    h2d_rate("from cudaHostAlloc", dev, pin, n);
    // anonymous memory registered (what a copy of the file into private memory would allow)
    void* anon = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_POPULATE, -1, 0);
    double t0 = now();
    cudaError_t e = cudaHostRegister(anon, n, cudaHostRegisterDefault);
    printf("anonymous memory reg default: %s (%.1f ms)\n", cudaGetErrorString(e), (now() - t0) * 1e3);
    return 0;
}
