// sinkhorn.cu -- persistent log-domain Sinkhorn kernel (utils/matching.py:46-57).
//
// One CTA owns one h x h block for all T iterations: the block lives in REGISTERS (512 threads x 32 values for
// h <= 128), row log-sum-exps are warp-shuffle reductions, column log-sum-exps go through a two-stage shared-memory
// reduction, and HBM is touched exactly twice (read L0 = -lambda*C at the start; write P at the end, re-reading L0
// once for <P,C>).  The reference instead unrolls ~14 TensorFlow ops per iteration per block into the graph.
//
// exp/log are single MUFU ex2/lg2 ops with log2(e) applied to the (small) max-subtracted differences; the reference's
// update order is kept literally:   log_a -= LSE(log_a, axis=1);  log_a -= LSE(log_a, axis=0)   (max-subtracted LSE,
// like tf.reduce_logsumexp), then P = softmax(log_a, -1), entropy = mean_i(-sum_j P log_softmax(log_a)).
#include "common.cuh"
#include <math.h>

namespace otgan {

namespace {

constexpr int HMAX = 128;          // largest block side of the single-CTA kernel
constexpr int NWARPS = 16, NTHREADS = NWARPS * 32, RPW = HMAX / NWARPS;   // 8 rows per warp, 4 columns per lane
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct ColScratch {
    float part[NWARPS][HMAX];   // per-warp column partials
    float fin[HMAX];            // reduced column values
};

// Reduce per-thread column values v[4] (columns 4*lane..4*lane+3) over all 16 warps; every thread gets the result for
// its own 4 columns.  IS_MAX selects max vs sum.  Two __syncthreads; fixed reduction order (deterministic).
template <bool IS_MAX>
__device__ __forceinline__ void col_allreduce(float (&v)[4], ColScratch& sc, int warp, int lane)
{
    *reinterpret_cast<float4*>(&sc.part[warp][4 * lane]) = make_float4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    {
        const int p = lane >> 1, g = lane & 1;          // warp w reduces columns [8w, 8w+8): 16 partials x 2 float4
        float4 t = *reinterpret_cast<const float4*>(&sc.part[p][8 * warp + 4 * g]);
#pragma unroll
        for (int o = 2; o <= 16; o <<= 1) {
            float4 u;
            u.x = __shfl_xor_sync(0xffffffffu, t.x, o);
            u.y = __shfl_xor_sync(0xffffffffu, t.y, o);
            u.z = __shfl_xor_sync(0xffffffffu, t.z, o);
            u.w = __shfl_xor_sync(0xffffffffu, t.w, o);
            if (IS_MAX) { t.x = fmaxf(t.x, u.x); t.y = fmaxf(t.y, u.y); t.z = fmaxf(t.z, u.z); t.w = fmaxf(t.w, u.w); }
            else        { t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        }
        if (lane < 2) *reinterpret_cast<float4*>(&sc.fin[8 * warp + 4 * g]) = t;
    }
    __syncthreads();
    const float4 r = *reinterpret_cast<const float4*>(&sc.fin[4 * lane]);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}

__global__ void __launch_bounds__(NTHREADS, 1)
sinkhorn_reg_kernel(const float* __restrict__ L0, float* __restrict__ P, float* __restrict__ entropy,
                    float* __restrict__ pc, int rows, int cols, int T, float lam)
{
    __shared__ __align__(16) ColScratch sc;
    __shared__ float red[2][NWARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = warp * RPW, c0 = lane * 4;
    const size_t boff = (size_t)blockIdx.x * rows * cols;
    const float* __restrict__ L0b = L0 + boff;
    const bool vec = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(L0b) & 15u) == 0);

    float x[RPW][4];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = r0 + i;
        if (vec && r < rows && c0 < cols) {
            const float4 t = *reinterpret_cast<const float4*>(L0b + (size_t)r * cols + c0);
            x[i][0] = t.x; x[i][1] = t.y; x[i][2] = t.z; x[i][3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                x[i][j] = (r < rows && c0 + j < cols) ? L0b[(size_t)r * cols + c0 + j] : -INFINITY;
        }
    }

    for (int it = 0; it < T; ++it) {
        // ---- log_a -= reduce_logsumexp(log_a, axis=1)            utils/matching.py:53
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            float m = fmaxf(fmaxf(x[i][0], x[i][1]), fmaxf(x[i][2], x[i][3]));
            m = warp_max(m);
            if (m == -INFINITY) m = 0.f;
            float s = ex2_approx((x[i][0] - m) * LOG2E) + ex2_approx((x[i][1] - m) * LOG2E) +
                      ex2_approx((x[i][2] - m) * LOG2E) + ex2_approx((x[i][3] - m) * LOG2E);
            s = warp_sum(s);
            const float lse = (r0 + i < rows) ? m + LN2 * lg2_approx(s) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) x[i][j] -= lse;
        }
        // ---- log_a -= reduce_logsumexp(log_a, axis=0)            utils/matching.py:54
        float cm[4], cs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float m = x[0][j];
#pragma unroll
            for (int i = 1; i < RPW; ++i) m = fmaxf(m, x[i][j]);
            cm[j] = m;
        }
        col_allreduce<true>(cm, sc, warp, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (cm[j] == -INFINITY) cm[j] = 0.f;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < RPW; ++i) s += ex2_approx((x[i][j] - cm[j]) * LOG2E);
            cs[j] = s;
        }
        col_allreduce<false>(cs, sc, warp, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lse = (c0 + j < cols) ? cm[j] + LN2 * lg2_approx(cs[j]) : 0.f;
#pragma unroll
            for (int i = 0; i < RPW; ++i) x[i][j] -= lse;
        }
    }

    // ---- P = softmax(log_a); entropy = mean_i(-sum_j P log_softmax(log_a)); <P,C>      utils/matching.py:56-57
    float ent = 0.f, pcs = 0.f;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = r0 + i;
        float m = fmaxf(fmaxf(x[i][0], x[i][1]), fmaxf(x[i][2], x[i][3]));
        m = warp_max(m);
        if (m == -INFINITY) m = 0.f;
        float e[4];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { e[j] = ex2_approx((x[i][j] - m) * LOG2E); s += e[j]; }
        s = warp_sum(s);
        const float ls = LN2 * lg2_approx(s);
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = (r < rows) && (c0 + j < cols);
            p[j] = ok ? __fdiv_rn(e[j], s) : 0.f;
            if (ok) ent -= p[j] * ((x[i][j] - m) - ls);
        }
        if (r < rows && c0 < cols) {
            const size_t off = (size_t)r * cols + c0;
            if (vec) {
                if (pc) {
                    const float4 l = *reinterpret_cast<const float4*>(L0b + off);
                    pcs += p[0] * l.x + p[1] * l.y + p[2] * l.z + p[3] * l.w;
                }
                if (P) *reinterpret_cast<float4*>(P + boff + off) = make_float4(p[0], p[1], p[2], p[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < cols) {
                        if (pc) pcs += p[j] * L0b[off + j];
                        if (P) P[boff + off + j] = p[j];
                    }
            }
        }
    }
    ent = warp_sum(ent);
    pcs = warp_sum(pcs);
    if (lane == 0) { red[0][warp] = ent; red[1][warp] = pcs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) { a += red[0][w]; b += red[1][w]; }
        if (entropy) entropy[blockIdx.x] = a / (float)rows;
        if (pc) pc[blockIdx.x] = -b / lam;       // C = -L0 / lambda
    }
}

__global__ void distance_from_pc_kernel(const float* __restrict__ pc, const float* __restrict__ entropy, int n_total,
                                        float* __restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // (sum_{k=2..5} <P_k,C_k> - 2<P_0,C_0> - 2<P_1,C_1>) / (2N)   == utils/matching.py:139-153
        out[0] = (((pc[2] + pc[3]) + (pc[4] + pc[5])) - 2.f * pc[0] - 2.f * pc[1]) / (2.f * (float)n_total);
        out[1] = (entropy[0] + entropy[1] + entropy[2] + entropy[3] + entropy[4] + entropy[5]) / 6.f;  // :61
    }
}

}  // namespace

int sinkhorn_reg_launch(int nblk, int rows, int cols, int T, float lam, const float* L0, float* P, float* entropy,
                        float* pc, cudaStream_t stream)
{
    sinkhorn_reg_kernel<<<nblk, NTHREADS, 0, stream>>>(L0, P, entropy, pc, rows, cols, T, lam);
    OTGAN_CHECK_LAUNCH("sinkhorn_reg_kernel");
    return OTGAN_OK;
}

int sinkhorn_reg_max_side() { return HMAX; }

int distance_from_pc_launch(const float* pc, const float* entropy, int n_total, float* out, cudaStream_t stream)
{
    distance_from_pc_kernel<<<1, 32, 0, stream>>>(pc, entropy, n_total, out);
    OTGAN_CHECK_LAUNCH("distance_from_pc_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
