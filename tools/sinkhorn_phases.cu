// sinkhorn_phases.cu -- phase clocks of the persistent Sinkhorn kernel (measurement aid, not part of libotgan.so).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o gpurun_out/sinkhorn_phases tools/sinkhorn_phases.cu
//   gpurun_out/sinkhorn_phases L0.bin nblk h T        (L0.bin: nblk*h*h raw floats, e.g. dumped by tools/sinkhorn_timing.py --dump)
//
// Compiles the kernel source itself with OTGAN_SINKHORN_CLOCKS, launches it a few times and prints, per block, the cycles spent in
// staging / first (slow) half-step / main loop / epilogue and the total of the slow steps.
#define OTGAN_SINKHORN_CLOCKS 1
#include "../otgan_b200/csrc/sinkhorn_fast.cu"
#include <vector>
#include <cstdlib>

namespace otgan {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch(int) {}
}  // namespace otgan

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s L0.bin nblk h T\n", argv[0]); return 2; }
    const int nblk = atoi(argv[2]), h = atoi(argv[3]), T = atoi(argv[4]);
    const size_t n = (size_t)nblk * h * h;
    std::vector<float> L0(n);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(L0.data(), sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %zu floats from %s\n", n, argv[1]); return 2; }
    fclose(f);
    float *dL, *dP, *dE, *dC;
    int* dS;
    cudaMalloc(&dL, n * 4); cudaMalloc(&dP, n * 4); cudaMalloc(&dE, nblk * 4); cudaMalloc(&dC, nblk * 4); cudaMalloc(&dS, nblk * 4);
    cudaMemcpy(dL, L0.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        int rc = otgan::sinkhorn_fast_launch(nblk, h, h, T, 500.f, dL, dP, dE, dC, dS, 0);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (rc != 0 || err != cudaSuccess) { fprintf(stderr, "launch failed rc=%d %s\n", rc, cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long clk[OTGAN_MAX_BLOCKS][8];
        cudaMemcpyFromSymbol(clk, otgan::g_sinkhorn_clk, sizeof(clk));
        std::vector<int> slow(nblk);
        cudaMemcpy(slow.data(), dS, nblk * 4, cudaMemcpyDeviceToHost);
        printf("rep %d: %.1f us (event pair around one launch)\n", rep, ms * 1e3);
        for (int b = 0; b < nblk && b < OTGAN_MAX_BLOCKS; ++b)
            printf("  block %d: staging %lld | first half-step %lld | loop %lld (slow steps: %d, %lld cycles incl. the first) | epilogue %lld | total %lld cycles\n",
                   b, clk[b][1] - clk[b][0], clk[b][2] - clk[b][1], clk[b][3] - clk[b][2], slow[b], clk[b][5], clk[b][4] - clk[b][3],
                   clk[b][4] - clk[b][0]);
    }
    return 0;
}
