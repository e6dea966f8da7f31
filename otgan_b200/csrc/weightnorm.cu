// weightnorm.cu -- weight-norm reparameterisation of utils/nn.py:176-180 fused with the layout change the conv needs.
//   forward : V [K, C] (HWIO / [in, out]: output channel contiguous), g [C]  ->  Wt [C, K] = g[c] * V[k][c] / ||V[:,c]||
//             (C-major == OHWI == channels-last OIHW, what cuDNN / the implicit-GEMM kernels read), inv[c] = 1/||V[:,c]||
//   backward: dWt [C, K] -> dV [K, C] = s_c (dWt[c][k] - V[k][c] inv_c^2 dot_c),  dg[c] = dot_c inv_c,
//             dot_c = sum_k dWt[c][k] V[k][c],  s_c = g_c inv_c
// Generic form (DenseNet): `perm` (nullable) permutes the K rows -- Wt[c][k] is built from V[perm[k]][c] (the crelu8 channel
// order of the concatenated feature buffers) -- and Wt / dWt may be addressed as [c][tap][ci] with strides (ldrow, ldtap) when
// `cin` > 0 (a row block of a wider filter-gradient tensor), k = tap * cin + ci.
// l2_normalize semantics: inv = rsqrt(max(sum V^2, 1e-12)).  HBM-bound: forward reads V twice (norm pass + scale pass) and
// writes Wt once; column reductions go through fixed-order partials (deterministic).  The reference runs ~6 TensorFlow
// elementwise/reduction ops per layer per forward for this.
#include "common.cuh"

namespace otgan {

namespace {

constexpr int TS = 32;

// partial[ks][c] = sum over the K-slice of (mode 0: V^2, mode 1: dWt * V)
struct WtAddr {                 // address of Wt / dWt element (c, k)
    int cin;                    // 0: dense [C][K]
    long long ldrow, ldtap;
    __device__ __forceinline__ size_t operator()(int c, int k, int K) const {
        if (cin == 0) return (size_t)c * K + k;
        const int t = k / cin;
        return (size_t)c * ldrow + (size_t)t * ldtap + (k - t * cin);
    }
};

template <int MODE>
__global__ void __launch_bounds__(256)
wn_col_partial_kernel(int K, int C, int rows_per_slice, const float* __restrict__ V, const float* __restrict__ dWt,
                      float* __restrict__ partial, const int* __restrict__ perm, WtAddr wa)
{
    __shared__ float tile[TS][TS + 1];
    __shared__ float red[8][TS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * TS, c = c0 + tx;
    const int kbeg = blockIdx.y * rows_per_slice;
    int kend = kbeg + rows_per_slice;
    kend = kend > K ? K : kend;
    float s = 0.f;
    if (MODE == 0) {
        if (c < C)
            for (int k = kbeg + ty; k < kend; k += 8) { const float v = V[(size_t)k * C + c]; s = fmaf(v, v, s); }
    } else {
        for (int k0 = kbeg; k0 < kend; k0 += TS) {
            // dWt tile [32 c][32 k] read coalesced along k, transposed through shared memory
#pragma unroll
            for (int r = 0; r < TS / 8; ++r) {
                const int cc = c0 + ty + 8 * r, kk = k0 + tx;
                tile[ty + 8 * r][tx] = (cc < C && kk < kend) ? dWt[wa(cc, kk, K)] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < TS / 8; ++r) {
                const int kk = k0 + ty + 8 * r;
                if (c < C && kk < kend) s = fmaf(tile[tx][ty + 8 * r], V[(size_t)(perm ? perm[kk] : kk) * C + c], s);
            }
            __syncthreads();
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][tx];
        partial[(size_t)blockIdx.y * C + c] = t;
    }
}

// grid (C/32, K/32): Wt[c][k] = V[k][c] * g[c] * inv[c]
__global__ void __launch_bounds__(256)
wn_fwd_apply_kernel(int K, int C, int KS, const float* __restrict__ V, const float* __restrict__ g,
                    const float* __restrict__ partial, float* __restrict__ Wt, float* __restrict__ inv,
                    const int* __restrict__ perm, WtAddr wa)
{
    __shared__ float tile[TS][TS + 1];
    __shared__ float scale[TS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * TS, k0 = blockIdx.y * TS;
    if (ty == 0) {
        const int c = c0 + tx;
        float t = 0.f;
        if (c < C) for (int s = 0; s < KS; ++s) t += partial[(size_t)s * C + c];
        const float r = rsqrtf(fmaxf(t, 1e-12f));
        scale[tx] = (c < C) ? g[c] * r : 0.f;
        if (blockIdx.y == 0 && c < C) inv[c] = r;
    }
#pragma unroll
    for (int r = 0; r < TS / 8; ++r) {
        const int k = k0 + ty + 8 * r, c = c0 + tx;
        tile[ty + 8 * r][tx] = (k < K && c < C) ? V[(size_t)(perm ? perm[k] : k) * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TS / 8; ++r) {
        const int c = c0 + ty + 8 * r, k = k0 + tx;
        if (c < C && k < K) Wt[wa(c, k, K)] = tile[tx][ty + 8 * r] * scale[ty + 8 * r];
    }
}

// grid (C/32, K/32): dV[k][c] = s_c (dWt[c][k] - V[k][c] inv_c^2 dot_c); dg[c] = dot_c inv_c
__global__ void __launch_bounds__(256)
wn_bwd_apply_kernel(int K, int C, int KS, const float* __restrict__ V, const float* __restrict__ g,
                    const float* __restrict__ inv, const float* __restrict__ partial, const float* __restrict__ dWt,
                    float* __restrict__ dV, float* __restrict__ dg, const int* __restrict__ perm, WtAddr wa)
{
    __shared__ float tile[TS][TS + 1];
    __shared__ float s_c[TS], q_c[TS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * TS, k0 = blockIdx.y * TS;
    if (ty == 0) {
        const int c = c0 + tx;
        float dot = 0.f;
        if (c < C) for (int s = 0; s < KS; ++s) dot += partial[(size_t)s * C + c];
        const float r = (c < C) ? inv[c] : 0.f, gg = (c < C) ? g[c] : 0.f;
        s_c[tx] = gg * r;
        q_c[tx] = r * r * dot;
        if (blockIdx.y == 0 && c < C) dg[c] = dot * r;
    }
#pragma unroll
    for (int r = 0; r < TS / 8; ++r) {
        const int c = c0 + ty + 8 * r, k = k0 + tx;
        tile[ty + 8 * r][tx] = (c < C && k < K) ? dWt[wa(c, k, K)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TS / 8; ++r) {
        const int k = k0 + ty + 8 * r, c = c0 + tx;
        if (k < K && c < C) {
            const size_t vi = (size_t)(perm ? perm[k] : k) * C + c;
            dV[vi] = s_c[tx] * (tile[tx][ty + 8 * r] - V[vi] * q_c[tx]);
        }
    }
}

int plan_slices(int K, int C, int* rows_per_slice)
{
    const int col_blocks = ceil_div(C, TS);
    int KS = (4 * kNumSMs) / (col_blocks > 0 ? col_blocks : 1);
    KS = KS < 1 ? 1 : (KS > 64 ? 64 : KS);
    int rps = ceil_div(K, KS);
    rps = ceil_div(rps, TS) * TS;                 // whole 32-row tiles per slice
    *rows_per_slice = rps;
    return ceil_div(K, rps);
}

}  // namespace

size_t weightnorm_workspace_bytes(int K, int C) { (void)K; return (size_t)64 * C * sizeof(float); }

int weightnorm_fwd_ex_launch(int K, int C, const float* V, const float* g, const int* perm, int cin, long long ldtap, long long ldrow,
                             float* Wt, float* inv, void* ws, cudaStream_t stream)
{
    int rps;
    const int KS = plan_slices(K, C, &rps);
    float* partial = reinterpret_cast<float*>(ws);
    const WtAddr wa{cin, ldrow, ldtap};
    wn_col_partial_kernel<0><<<dim3(ceil_div(C, TS), KS), 256, 0, stream>>>(K, C, rps, V, nullptr, partial, perm, wa);
    OTGAN_CHECK_LAUNCH("wn_col_partial_kernel<0>");
    wn_fwd_apply_kernel<<<dim3(ceil_div(C, TS), ceil_div(K, TS)), 256, 0, stream>>>(K, C, KS, V, g, partial, Wt, inv, perm, wa);
    OTGAN_CHECK_LAUNCH("wn_fwd_apply_kernel");
    return OTGAN_OK;
}

int weightnorm_fwd_launch(int K, int C, const float* V, const float* g, float* Wt, float* inv, void* ws, cudaStream_t stream)
{
    return weightnorm_fwd_ex_launch(K, C, V, g, nullptr, 0, 0, 0, Wt, inv, ws, stream);
}

int weightnorm_bwd_ex_launch(int K, int C, const float* V, const float* g, const float* inv, const int* perm, int cin, long long ldtap,
                             long long ldrow, const float* dWt, float* dV, float* dg, void* ws, cudaStream_t stream)
{
    int rps;
    const int KS = plan_slices(K, C, &rps);
    float* partial = reinterpret_cast<float*>(ws);
    const WtAddr wa{cin, ldrow, ldtap};
    wn_col_partial_kernel<1><<<dim3(ceil_div(C, TS), KS), 256, 0, stream>>>(K, C, rps, V, dWt, partial, perm, wa);
    OTGAN_CHECK_LAUNCH("wn_col_partial_kernel<1>");
    wn_bwd_apply_kernel<<<dim3(ceil_div(C, TS), ceil_div(K, TS)), 256, 0, stream>>>(K, C, KS, V, g, inv, partial, dWt, dV, dg, perm, wa);
    OTGAN_CHECK_LAUNCH("wn_bwd_apply_kernel");
    return OTGAN_OK;
}

int weightnorm_bwd_launch(int K, int C, const float* V, const float* g, const float* inv, const float* dWt, float* dV,
                          float* dg, void* ws, cudaStream_t stream)
{
    return weightnorm_bwd_ex_launch(K, C, V, g, inv, nullptr, 0, 0, 0, dWt, dV, dg, ws, stream);
}

}  // namespace otgan
