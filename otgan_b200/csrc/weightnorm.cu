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

// ---- streaming forms (round 2): float4 along the contiguous output-channel axis, no shared-memory transposes except the one
// the OHWI layout itself needs.
// partial[ks][c] = sum over the K-slice of A[k][c] * B[k][c]   (A == B: squared norms; A = dW (HWIO), B = V: the weight-norm dot)
__global__ void __launch_bounds__(256)
wn_coldot4_kernel(int K, int C4, int rows_per_slice, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ partial)
{
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c4 = blockIdx.x * 32 + tx;
    const int kbeg = blockIdx.y * rows_per_slice;
    int kend = kbeg + rows_per_slice;
    kend = kend > K ? K : kend;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < C4)
        for (int k = kbeg + ty; k < kend; k += 8) {
            const float4 a = A[(size_t)k * C4 + c4], b = B[(size_t)k * C4 + c4];
            s.x = fmaf(a.x, b.x, s.x); s.y = fmaf(a.y, b.y, s.y); s.z = fmaf(a.z, b.z, s.z); s.w = fmaf(a.w, b.w, s.w);
        }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c4 < C4) {
#pragma unroll
        for (int r = 1; r < 8; ++r) { const float4 v = red[r][tx]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        partial[(size_t)blockIdx.y * C4 + c4] = s;
    }
}

// Forward apply, 64 x 64 tiles: Wt[c][k] = V[k][c] * g[c] * inv[c] (OHWI, through a shared-memory transpose, float4 on both sides)
// and, when ihwo != null, ihwo[(ci * T + t)][c] = V[(t * cin + ci)][c] * g[c] * inv[c] (the dgrad operand: a scaled row
// permutation of V, written straight from the loaded tile -- no second transpose pass).  K % 4 == 0, C % 4 == 0.
__global__ void __launch_bounds__(256)
wn_fwd_apply64_kernel(int K, int C, int KS, int T, const float* __restrict__ V, const float* __restrict__ g,
                      const float* __restrict__ partial, float* __restrict__ Wt, float* __restrict__ ihwo, float* __restrict__ inv)
{
    __shared__ float tile[64][65];
    __shared__ float scale[64];
    const int c0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    if (threadIdx.x < 64) {
        const int c = c0 + threadIdx.x;
        float t = 0.f;
        if (c < C) for (int s = 0; s < KS; ++s) t += partial[(size_t)s * C + c];
        const float r = rsqrtf(fmaxf(t, 1e-12f));
        scale[threadIdx.x] = (c < C) ? g[c] * r : 0.f;
        if (blockIdx.y == 0 && c < C) inv[c] = r;
    }
    __syncthreads();
    const int cin = K / T;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;       // 16 float4 columns x 16 rows per pass
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int kk = ly + 16 * r, k = k0 + kk, c = c0 + 4 * lx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K && c < C) {
            v = *reinterpret_cast<const float4*>(V + (size_t)k * C + c);
            v.x *= scale[4 * lx]; v.y *= scale[4 * lx + 1]; v.z *= scale[4 * lx + 2]; v.w *= scale[4 * lx + 3];
            if (ihwo) {
                const int t = k / cin, ci = k - t * cin;
                *reinterpret_cast<float4*>(ihwo + ((size_t)ci * T + t) * C + c) = v;
            }
        }
        tile[kk][4 * lx] = v.x; tile[kk][4 * lx + 1] = v.y; tile[kk][4 * lx + 2] = v.z; tile[kk][4 * lx + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int cc = ly + 16 * r, c = c0 + cc, k = k0 + 4 * lx;
        if (c < C && k < K)
            *reinterpret_cast<float4*>(Wt + (size_t)c * K + k) =
                make_float4(tile[4 * lx][cc], tile[4 * lx + 1][cc], tile[4 * lx + 2][cc], tile[4 * lx + 3][cc]);
    }
}

// Backward apply on an HWIO gradient: dV[k][c] = s_c (dW[k][c] - V[k][c] q_c), dg[c] = dot_c inv_c; pure streaming (float4 along c).
__global__ void __launch_bounds__(256)
wn_bwd_apply_hwio_kernel(int K, int C4, int KS, int rows_per_block, const float4* __restrict__ V, const float* __restrict__ g,
                         const float* __restrict__ inv, const float* __restrict__ partial, const float4* __restrict__ dW,
                         float4* __restrict__ dV, float* __restrict__ dg)
{
    __shared__ float4 s_c[32], q_c[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c4 = blockIdx.x * 32 + tx, C = 4 * C4;
    if (ty == 0) {
        float dot[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f}, qc[4] = {0.f, 0.f, 0.f, 0.f};
        if (c4 < C4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * c4 + e;
                for (int s = 0; s < KS; ++s) dot[e] += partial[(size_t)s * C + c];
                const float r = inv[c];
                sc[e] = g[c] * r;
                qc[e] = r * r * dot[e];
                if (blockIdx.y == 0) dg[c] = dot[e] * r;
            }
        }
        s_c[tx] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        q_c[tx] = make_float4(qc[0], qc[1], qc[2], qc[3]);
    }
    __syncthreads();
    if (c4 >= C4) return;
    const float4 s = s_c[tx], q = q_c[tx];
    const int kbeg = blockIdx.y * rows_per_block;
    int kend = kbeg + rows_per_block;
    kend = kend > K ? K : kend;
    for (int k = kbeg + ty; k < kend; k += 8) {
        const float4 d = dW[(size_t)k * C4 + c4], v = V[(size_t)k * C4 + c4];
        dV[(size_t)k * C4 + c4] = make_float4(s.x * (d.x - v.x * q.x), s.y * (d.y - v.y * q.y), s.z * (d.z - v.z * q.z), s.w * (d.w - v.w * q.w));
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

// K-slices of the float4 streaming kernels: their column blocks are 128 floats wide (32 float4), so the tiled plan above (32-float
// blocks) gave the 1024-channel layers a grid of 8 x 18 = 144 CTAs -- one per SM, one or two loads in flight per thread, 1.3 TB/s.
// Eight CTAs per SM's worth of slices (at most the 64 partial rows the workspace and the apply kernels are sized for).
int plan_slices4(int K, int C, int* rows_per_slice)
{
    const int col_blocks = ceil_div(C / 4, 32);
    int KS = (8 * kNumSMs) / (col_blocks > 0 ? col_blocks : 1);
    KS = KS < 1 ? 1 : (KS > 64 ? 64 : KS);
    int rps = ceil_div(K, KS);
    rps = ceil_div(rps, 8) * 8;                   // whole 8-row passes of a CTA per slice
    *rows_per_slice = rps;
    return ceil_div(K, rps);
}

}  // namespace

size_t weightnorm_workspace_bytes(int K, int C) { (void)K; return (size_t)64 * C * sizeof(float); }

int weightnorm_fwd_ex_launch(int K, int C, const float* V, const float* g, const int* perm, int cin, long long ldtap, long long ldrow,
                             float* Wt, float* inv, void* ws, cudaStream_t stream)
{
    // The squared-norm pass takes the K-slices of the streaming form whenever that form could have been used (same slices, same
    // summation order => bit-identical 1/||V|| from the fused and the stand-alone nodes; the tensor cores read W as TF32 by
    // truncation, so a last-bit difference of a filter's scale moves a few of its weights by 2^-10 and, through ReLU sign flips,
    // single gradient entries by percents: tests/test_conv_gpu.py::test_wn_fusion_matches_the_unfused_nodes compares the two)
    int rps;
    const int KS = (perm == nullptr && (K & 3) == 0 && (C & 3) == 0) ? plan_slices4(K, C, &rps) : plan_slices(K, C, &rps);
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

// W = g V / ||V|| as OHWI (+ the IHWO dgrad operand in the same pass); K = T * cin rows, K % 4 == 0, C % 4 == 0
int weightnorm_fwd2_launch(int K, int C, int T, const float* V, const float* g, float* Wt, float* ihwo, float* inv, void* ws, cudaStream_t stream)
{
    int rps;
    const int KS = plan_slices4(K, C, &rps);
    float* partial = reinterpret_cast<float*>(ws);
    wn_coldot4_kernel<<<dim3(ceil_div(C / 4, 32), KS), 256, 0, stream>>>(K, C / 4, rps, reinterpret_cast<const float4*>(V),
                                                                       reinterpret_cast<const float4*>(V), reinterpret_cast<float4*>(partial));
    OTGAN_CHECK_LAUNCH("wn_coldot4_kernel");
    wn_fwd_apply64_kernel<<<dim3(ceil_div(C, 64), ceil_div(K, 64)), 256, 0, stream>>>(K, C, KS, T, V, g, partial, Wt, ihwo, inv);
    OTGAN_CHECK_LAUNCH("wn_fwd_apply64_kernel");
    return OTGAN_OK;
}

// dW given in HWIO layout ([K][C], the layout of V): no transposes at all
int weightnorm_bwd_hwio_launch(int K, int C, const float* V, const float* g, const float* inv, const float* dW, float* dV, float* dg,
                               void* ws, cudaStream_t stream)
{
    int rps;
    const int KS = plan_slices4(K, C, &rps);
    float* partial = reinterpret_cast<float*>(ws);
    wn_coldot4_kernel<<<dim3(ceil_div(C / 4, 32), KS), 256, 0, stream>>>(K, C / 4, rps, reinterpret_cast<const float4*>(dW),
                                                                       reinterpret_cast<const float4*>(V), reinterpret_cast<float4*>(partial));
    OTGAN_CHECK_LAUNCH("wn_coldot4_kernel");
    int rows_per_block = ceil_div(K, ceil_div(8 * kNumSMs, ceil_div(C / 4, 32)));
    rows_per_block = rows_per_block < 8 ? 8 : rows_per_block;
    wn_bwd_apply_hwio_kernel<<<dim3(ceil_div(C / 4, 32), ceil_div(K, rows_per_block)), 256, 0, stream>>>(
        K, C / 4, KS, rows_per_block, reinterpret_cast<const float4*>(V), g, inv, partial, reinterpret_cast<const float4*>(dW),
        reinterpret_cast<float4*>(dV), dg);
    OTGAN_CHECK_LAUNCH("wn_bwd_apply_hwio_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
