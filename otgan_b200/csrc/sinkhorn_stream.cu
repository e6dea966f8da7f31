// sinkhorn_stream.cu -- Sinkhorn for blocks that fit neither one SM nor one cluster (side > 512; sides in (128, 512] are
// sinkhorn_cluster.cu's, this rung remains their OTGAN_IMPL_SIMT comparison path).
//
// Literal log-domain iteration of utils/matching.py:50-57 with the block held in global memory (it stays L2-resident:
// 6 x 1024^2 fp32 = 24 MB << 126 MB L2): one kernel per half-step,
//   row kernel : one warp per row, the row cached in registers (cols <= 1024), max-subtracted LSE, in-place subtract
//   col kernel : one CTA per 32-column strip, the strip cached in shared memory (rows <= 1024 -> 128 KB)
//   final      : P = softmax(log_a), entropy, <P,C>, one CTA per block, fixed-order reductions
// The caller's P buffer is the working storage (no hidden allocation).  Launch-bound for small T*h, L2-bound for h = 1024.
// Blocks with a side above 1024 (the reference's default flags: --batch_size 625 --nr_gpu 8 -> h = 2500, N = 5000 with
// --single_batch) take the *_big kernels: the same half-steps with an online (running max / rescaled sum) log-sum-exp that
// streams the row / column strip twice from L2 / HBM instead of caching it on chip -- any size.
#include "common.cuh"
#include <math.h>

namespace otgan {

namespace {

constexpr int MAXS = 1024;                  // largest supported block side
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// L_out[row] = L_in[row] - LSE(L_in[row]); grid = ceil(nblk*rows / 8), 8 warps per CTA
__global__ void __launch_bounds__(256)
sk_row_kernel(const float* __restrict__ Lin, float* __restrict__ Lout, int nrows_total, int cols)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nrows_total) return;
    const float* src = Lin + (size_t)row * cols;
    float* dst = Lout + (size_t)row * cols;
    float x[MAXS / 32];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXS / 32; ++i) {
        const int c = lane + 32 * i;
        x[i] = (c < cols) ? src[c] : -INFINITY;
        m = fmaxf(m, x[i]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXS / 32; ++i) s += ex2_approx((x[i] - m) * LOG2E);
    s = warp_sum(s);
    const float lse = m + LN2 * lg2_approx(s);
#pragma unroll
    for (int i = 0; i < MAXS / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < cols) dst[c] = x[i] - lse;
    }
}

// in-place column step; grid = (ceil(cols/32), nblk), block = (32, 32); dynamic smem = rows * 33 floats
__global__ void __launch_bounds__(1024)
sk_col_kernel(float* __restrict__ L, int rows, int cols)
{
    extern __shared__ float strip[];                 // [rows][33]
    __shared__ float red[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    float* base = L + (size_t)blockIdx.y * rows * cols;
    float m = -INFINITY;
    for (int r = ty; r < rows; r += 32) {
        const float v = (c < cols) ? base[(size_t)r * cols + c] : -INFINITY;
        strip[r * 33 + tx] = v;
        m = fmaxf(m, v);
    }
    red[ty][tx] = m;
    __syncthreads();
    m = red[0][tx];
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, red[i][tx]);
    if (!(m > -INFINITY)) m = 0.f;
    __syncthreads();
    float s = 0.f;
    for (int r = ty; r < rows; r += 32) s += ex2_approx((strip[r * 33 + tx] - m) * LOG2E);
    red[ty][tx] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][tx];          // fixed order
    const float lse = m + LN2 * lg2_approx(s);
    if (c < cols)
        for (int r = ty; r < rows; r += 32) base[(size_t)r * cols + c] = strip[r * 33 + tx] - lse;
}

// one CTA (1024 threads = 32 warps) per block: row softmax, entropy, <P,C>; P overwrites L in place
__global__ void __launch_bounds__(1024)
sk_final_kernel(float* __restrict__ LP, const float* __restrict__ L0, float* __restrict__ entropy, float* __restrict__ pc,
                int rows, int cols, float lam)
{
    __shared__ float red[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* base = LP + (size_t)blockIdx.x * rows * cols;
    const float* l0 = L0 + (size_t)blockIdx.x * rows * cols;
    float ent = 0.f, pcs = 0.f;
    for (int r = warp; r < rows; r += 32) {
        float x[MAXS / 32];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < MAXS / 32; ++i) {
            const int c = lane + 32 * i;
            x[i] = (c < cols) ? base[(size_t)r * cols + c] : -INFINITY;
            m = fmaxf(m, x[i]);
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXS / 32; ++i) s += ex2_approx((x[i] - m) * LOG2E);
        s = warp_sum(s);
        const float ls = LN2 * lg2_approx(s);
        float e_row = 0.f, p_row = 0.f;
#pragma unroll
        for (int i = 0; i < MAXS / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < cols) {
                const float p = __fdiv_rn(ex2_approx((x[i] - m) * LOG2E), s);
                base[(size_t)r * cols + c] = p;
                e_row -= p * ((x[i] - m) - ls);
                p_row += p * l0[(size_t)r * cols + c];
            }
        }
        ent += warp_sum(e_row);
        pcs += warp_sum(p_row);
    }
    if (lane == 0) { red[0][warp] = ent; red[1][warp] = pcs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 32; ++w) { a += red[0][w]; b += red[1][w]; }
        if (entropy) entropy[blockIdx.x] = a / (float)rows;
        if (pc) pc[blockIdx.x] = -b / lam;
    }
}

// ---- any-size variants: online log-sum-exp, two passes over global memory -------------------------------------------
struct Lse { float m, s; };
__device__ __forceinline__ void lse_push(Lse& a, float v)
{
    if (v > a.m) { a.s = a.s * ex2_approx((a.m - v) * LOG2E) + 1.f; a.m = v; }
    else if (v > -INFINITY) a.s += ex2_approx((v - a.m) * LOG2E);
}
__device__ __forceinline__ float lse_scale(float m_part, float m_all) { return m_part > -INFINITY ? ex2_approx((m_part - m_all) * LOG2E) : 0.f; }

__global__ void __launch_bounds__(256)
sk_row_big_kernel(const float* __restrict__ Lin, float* __restrict__ Lout, int nrows_total, int cols)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nrows_total) return;
    const float* src = Lin + (size_t)row * cols;
    float* dst = Lout + (size_t)row * cols;
    Lse a = {-INFINITY, 0.f};
    for (int c = lane; c < cols; c += 32) lse_push(a, src[c]);
    const float m = warp_max(a.m);
    const float s = warp_sum(a.s * lse_scale(a.m, m));
    const float lse = m + LN2 * lg2_approx(s);
    for (int c = lane; c < cols; c += 32) dst[c] = src[c] - lse;
}

__global__ void __launch_bounds__(1024)
sk_col_big_kernel(float* __restrict__ L, int rows, int cols)
{
    __shared__ float red_m[32][33], red_s[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    float* base = L + (size_t)blockIdx.y * rows * cols;
    Lse a = {-INFINITY, 0.f};
    if (c < cols)
        for (int r = ty; r < rows; r += 32) lse_push(a, base[(size_t)r * cols + c]);
    red_m[ty][tx] = a.m; red_s[ty][tx] = a.s;
    __syncthreads();
    float m = red_m[0][tx];
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, red_m[i][tx]);
    if (!(m > -INFINITY)) m = 0.f;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red_s[i][tx] * lse_scale(red_m[i][tx], m);          // fixed order
    const float lse = m + LN2 * lg2_approx(s);
    if (c < cols)
        for (int r = ty; r < rows; r += 32) base[(size_t)r * cols + c] -= lse;
}

__global__ void __launch_bounds__(1024)
sk_final_big_kernel(float* __restrict__ LP, const float* __restrict__ L0, float* __restrict__ entropy, float* __restrict__ pc,
                    int rows, int cols, float lam)
{
    __shared__ float red[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* base = LP + (size_t)blockIdx.x * rows * cols;
    const float* l0 = L0 + (size_t)blockIdx.x * rows * cols;
    float ent = 0.f, pcs = 0.f;
    for (int r = warp; r < rows; r += 32) {
        float* row = base + (size_t)r * cols;
        Lse a = {-INFINITY, 0.f};
        for (int c = lane; c < cols; c += 32) lse_push(a, row[c]);
        const float m = warp_max(a.m);
        const float s = warp_sum(a.s * lse_scale(a.m, m));
        const float ls = LN2 * lg2_approx(s);
        float e_row = 0.f, p_row = 0.f;
        for (int c = lane; c < cols; c += 32) {
            const float x = row[c];
            const float p = __fdiv_rn(ex2_approx((x - m) * LOG2E), s);
            row[c] = p;
            e_row -= p * ((x - m) - ls);
            p_row += p * l0[(size_t)r * cols + c];
        }
        ent += warp_sum(e_row);
        pcs += warp_sum(p_row);
    }
    if (lane == 0) { red[0][warp] = ent; red[1][warp] = pcs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 32; ++w) { a += red[0][w]; b += red[1][w]; }
        if (entropy) entropy[blockIdx.x] = a / (float)rows;
        if (pc) pc[blockIdx.x] = -b / lam;
    }
}

}  // namespace

int sinkhorn_stream_max_side() { return 1 << 20; }     // the *_big kernels take any block that fits in memory

int sinkhorn_stream_launch(int nblk, int rows, int cols, int T, float lam, const float* L0, float* P, float* entropy,
                           float* pc, cudaStream_t stream)
{
    const int nrows_total = nblk * rows;
    const size_t col_smem = (size_t)rows * 33 * sizeof(float);
    if (T == 0) {
        OTGAN_CUDA(cudaMemcpyAsync(P, L0, (size_t)nrows_total * cols * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    }
    if (rows > MAXS || cols > MAXS) {
        for (int it = 0; it < T; ++it) {
            sk_row_big_kernel<<<ceil_div(nrows_total, 8), 256, 0, stream>>>(it == 0 ? L0 : P, P, nrows_total, cols);
            OTGAN_CHECK_LAUNCH("sk_row_big_kernel");
            sk_col_big_kernel<<<dim3(ceil_div(cols, 32), nblk), dim3(32, 32), 0, stream>>>(P, rows, cols);
            OTGAN_CHECK_LAUNCH("sk_col_big_kernel");
        }
        sk_final_big_kernel<<<nblk, 1024, 0, stream>>>(P, L0, entropy, pc, rows, cols, lam);
        OTGAN_CHECK_LAUNCH("sk_final_big_kernel");
        return OTGAN_OK;
    }
    OTGAN_SET_MAX_SMEM(sk_col_kernel, MAXS * 33 * sizeof(float));
    for (int it = 0; it < T; ++it) {
        sk_row_kernel<<<ceil_div(nrows_total, 8), 256, 0, stream>>>(it == 0 ? L0 : P, P, nrows_total, cols);   // :53
        OTGAN_CHECK_LAUNCH("sk_row_kernel");
        sk_col_kernel<<<dim3(ceil_div(cols, 32), nblk), dim3(32, 32), col_smem, stream>>>(P, rows, cols);      // :54
        OTGAN_CHECK_LAUNCH("sk_col_kernel");
    }
    sk_final_kernel<<<nblk, 1024, 0, stream>>>(P, L0, entropy, pc, rows, cols, lam);                           // :56-57
    OTGAN_CHECK_LAUNCH("sk_final_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
