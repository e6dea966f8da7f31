// sinkhorn_cluster_phases.cu -- phase clocks of the cluster Sinkhorn kernel (measurement aid, not part of libotgan.so).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -o tools/bin/sinkhorn_cluster_phases tools/sinkhorn_cluster_phases.cu
//   tools/bin/sinkhorn_cluster_phases nblk h T
//
// Compiles the kernel source itself with OTGAN_SKC_CLOCKS, runs it on synthetic cost blocks (L0 = -500 * uniform[0.2, 1.4)) and
// prints the time of one launch and, for thread 0 of CTA 0, the cycles per iteration spent in each phase of the loop.
#define OTGAN_SKC_CLOCKS 1
#include "../otgan_b200/csrc/sinkhorn_cluster.cu"
#include <vector>
#include <cstdlib>

namespace otgan {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch(int) {}
}  // namespace otgan

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s nblk h T\n", argv[0]); return 2; }
    const int nblk = atoi(argv[1]), h = atoi(argv[2]), T = atoi(argv[3]);
    const size_t n = (size_t)nblk * h * h;
    std::vector<float> L0(n);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        L0[i] = -500.f * (0.2f + 1.2f * (float)((s >> 11) & 0xFFFFFF) / 16777216.f);
    }
    float *dL, *dP, *dE, *dC;
    cudaMalloc(&dL, n * 4); cudaMalloc(&dP, n * 4); cudaMalloc(&dE, nblk * 4); cudaMalloc(&dC, nblk * 4);
    cudaMemcpy(dL, L0.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    static const char* names[8] = {"row step", "column partials (4 rows)", "__syncthreads", "slab LSE + DSMEM push", "wait for the 8 slabs",
                                   "combine 8 slabs", "__syncthreads", "subtract"};
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        int rc = otgan::sinkhorn_cluster_launch(nblk, h, h, T, 500.f, dL, dP, dE, dC, 0);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (rc != 0 || err != cudaSuccess) { fprintf(stderr, "launch failed rc=%d %s\n", rc, cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long clk[8];
        cudaMemcpyFromSymbol(clk, otgan::g_skc_clk, sizeof(clk));
        printf("rep %d: %.1f us (event pair around one launch)\n", rep, ms * 1e3);
        if (rep == 2) {
            long long tot = 0;
            for (int i = 0; i < 8; ++i) tot += clk[i];
            for (int i = 0; i < 8; ++i) printf("  %-28s %7.0f cycles / iteration\n", names[i], (double)clk[i] / (T > 0 ? T : 1));
            printf("  %-28s %7.0f cycles / iteration\n", "total", (double)tot / (T > 0 ? T : 1));
        }
    }
    return 0;
}
