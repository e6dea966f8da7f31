// sinkhorn_cluster.cu -- persistent log-domain Sinkhorn for blocks of 128 < side <= 512 (utils/matching.py:46-57).
//
// A block of this size does not fit one SM's registers, and the streaming rung (sinkhorn_stream.cu) pays two launches per
// iteration.  Here a thread-block CLUSTER of 8 CTAs owns one block for all T iterations: CTA q keeps the row slab
// [q*R, (q+1)*R) of log_a in REGISTERS (warp = 4 rows, lane = CPL columns), so
//   * the row half-step (log_a -= LSE over axis 1, :53) is warp-local, exactly like sinkhorn.cu;
//   * the column half-step (:54) reduces in three levels: thread (4 rows) -> CTA (shared memory, (max, sum) pairs combined
//     online) -> cluster: every CTA pushes ONE float per column -- its slab's log-sum-exp -- into the other seven CTAs' shared
//     memory (DSMEM), one barrier.cluster later every CTA combines the eight slab values itself (LSE of LSEs is exact, no second
//     exchange).  One value per column keeps the push at 8 KB per CTA (DSMEM moves ~20 B/clk); the exchange buffer is
//     double-buffered by iteration parity so that one cluster barrier per iteration is enough;
//   * HBM is touched twice: L0 in, P out (+ L0 once more for <P,C>), like the single-CTA kernels.
// Arithmetic = the literal rung's (max-subtracted LSE with MUFU ex2 / lg2 on differences); the update order of the reference is
// kept.  Two shapes: side <= 256 (8 warps, 8 columns per lane) and side <= 512 (16 warps, 16 columns per lane).
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace otgan {

namespace {

constexpr int CL = 8;                        // CTAs per cluster (portable maximum)
constexpr int RPW = 4;                       // rows per warp
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

template <int NW, int CPL>
struct ClusterSmem {
    static constexpr int HC = 32 * CPL;      // columns held by one warp = largest block width
    float pm[NW][HC];                        // per-warp column partials: max ...
    float ps[NW][HC];                        // ... and sum of 2^((x - max) log2 e)
    float xch[2][CL][HC];                    // [iteration parity][source CTA][column] slab log-sum-exps (written through DSMEM)
    float lse[HC];                           // column log-sum-exp of the whole block
    float red[2][NW];                        // entropy / <P,L0> partials of this CTA
    float fin[2][CL];                        // ... of every CTA (rank 0's copy is the one that is read)
};

template <int NW, int CPL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NW * 32, 1)
sinkhorn_cluster_kernel(const float* __restrict__ L0, float* __restrict__ P, float* __restrict__ entropy,
                        float* __restrict__ pc, int rows, int cols, int T, float lam)
{
    using Smem = ClusterSmem<NW, CPL>;
    constexpr int HC = Smem::HC, NG = CPL / 4;            // NG float4 groups per lane
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int blk = blockIdx.x / CL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (rank * NW + warp) * RPW;
    const size_t boff = (size_t)blk * rows * cols;
    const float* __restrict__ L0b = L0 + boff;
    const bool vec = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(L0b) & 15u) == 0);

    float x[RPW][CPL];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = r0 + i;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int c0 = 128 * g + 4 * lane;
            if (vec && r < rows && c0 < cols) {
                const float4 t = *reinterpret_cast<const float4*>(L0b + (size_t)r * cols + c0);
                x[i][4 * g + 0] = t.x; x[i][4 * g + 1] = t.y; x[i][4 * g + 2] = t.z; x[i][4 * g + 3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    x[i][4 * g + j] = (r < rows && c0 + j < cols) ? L0b[(size_t)r * cols + c0 + j] : -INFINITY;
            }
        }
    }

    for (int it = 0; it < T; ++it) {
        // ---- log_a -= reduce_logsumexp(log_a, axis=1)            utils/matching.py:53   (warp-local)
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            float m = x[i][0];
#pragma unroll
            for (int j = 1; j < CPL; ++j) m = fmaxf(m, x[i][j]);
            m = warp_max(m);
            if (m == -INFINITY) m = 0.f;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) s += ex2_approx((x[i][j] - m) * LOG2E);
            s = warp_sum(s);
            const float lse = (r0 + i < rows) ? m + LN2 * lg2_approx(s) : 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) x[i][j] -= lse;
        }

        // ---- log_a -= reduce_logsumexp(log_a, axis=0)            utils/matching.py:54
        // level 1: this thread's 4 rows
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float mm[4], ss[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float m = x[0][4 * g + j];
#pragma unroll
                for (int i = 1; i < RPW; ++i) m = fmaxf(m, x[i][4 * g + j]);
                const float mz = (m == -INFINITY) ? 0.f : m;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < RPW; ++i) s += ex2_approx((x[i][4 * g + j] - mz) * LOG2E);
                mm[j] = m; ss[j] = s;
            }
            *reinterpret_cast<float4*>(&sm.pm[warp][128 * g + 4 * lane]) = make_float4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<float4*>(&sm.ps[warp][128 * g + 4 * lane]) = make_float4(ss[0], ss[1], ss[2], ss[3]);
        }
        __syncthreads();
        // level 2: the CTA's slab; one thread per column, fixed order; the slab's log-sum-exp goes to all eight CTAs
        const int par = it & 1;
        for (int c = threadIdx.x; c < HC; c += NW * 32) {
            float m = sm.pm[0][c];
#pragma unroll
            for (int w = 1; w < NW; ++w) m = fmaxf(m, sm.pm[w][c]);
            float l = -INFINITY;
            if (m > -INFINITY) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const float pmw = sm.pm[w][c];
                    s += (pmw > -INFINITY) ? sm.ps[w][c] * ex2_approx((pmw - m) * LOG2E) : 0.f;
                }
                l = m + LN2 * lg2_approx(s);
            }
#pragma unroll
            for (int q = 0; q < CL; ++q) *cluster.map_shared_rank(&sm.xch[par][rank][c], q) = l;
        }
        cluster.sync();
        // level 3: the eight slabs
        for (int c = threadIdx.x; c < HC; c += NW * 32) {
            float l[CL];
            float m = -INFINITY;
#pragma unroll
            for (int q = 0; q < CL; ++q) { l[q] = sm.xch[par][q][c]; m = fmaxf(m, l[q]); }
            float lse = 0.f;                                 // empty column (c >= cols): nothing to subtract
            if (m > -INFINITY) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < CL; ++q) s += ex2_approx((l[q] - m) * LOG2E);
                lse = m + LN2 * lg2_approx(s);
            }
            sm.lse[c] = lse;
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(&sm.lse[128 * g + 4 * lane]);
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                x[i][4 * g + 0] -= t.x; x[i][4 * g + 1] -= t.y; x[i][4 * g + 2] -= t.z; x[i][4 * g + 3] -= t.w;
            }
        }
    }

    // ---- P = softmax(log_a); entropy = mean_i(-sum_j P log_softmax(log_a)); <P,C>      utils/matching.py:56-57
    float ent = 0.f, pcs = 0.f;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = r0 + i;
        float m = x[i][0];
#pragma unroll
        for (int j = 1; j < CPL; ++j) m = fmaxf(m, x[i][j]);
        m = warp_max(m);
        if (m == -INFINITY) m = 0.f;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) s += ex2_approx((x[i][j] - m) * LOG2E);
        s = warp_sum(s);
        const float ls = LN2 * lg2_approx(s);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int c0 = 128 * g + 4 * lane;
            float p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = (r < rows) && (c0 + j < cols);
                const float d = x[i][4 * g + j] - m;
                p[j] = ok ? __fdiv_rn(ex2_approx(d * LOG2E), s) : 0.f;
                if (ok) ent -= p[j] * (d - ls);
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
    }
    ent = warp_sum(ent);
    pcs = warp_sum(pcs);
    if (lane == 0) { sm.red[0][warp] = ent; sm.red[1][warp] = pcs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { a += sm.red[0][w]; b += sm.red[1][w]; }
        *cluster.map_shared_rank(&sm.fin[0][rank], 0) = a;
        *cluster.map_shared_rank(&sm.fin[1][rank], 0) = b;
    }
    cluster.sync();                      // also keeps every CTA's shared memory alive until all remote stores have landed
    if (rank == 0 && threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int q = 0; q < CL; ++q) { a += sm.fin[0][q]; b += sm.fin[1][q]; }      // fixed order
        if (entropy) entropy[blk] = a / (float)rows;
        if (pc) pc[blk] = -b / lam;      // C = -L0 / lambda
    }
}

}  // namespace

int sinkhorn_cluster_max_side() { return 512; }

int sinkhorn_cluster_launch(int nblk, int rows, int cols, int T, float lam, const float* L0, float* P, float* entropy,
                            float* pc, cudaStream_t stream)
{
    if (rows <= 256 && cols <= 256) {
        OTGAN_SET_MAX_SMEM((sinkhorn_cluster_kernel<8, 8>), sizeof(ClusterSmem<8, 8>));
        sinkhorn_cluster_kernel<8, 8><<<nblk * CL, 8 * 32, sizeof(ClusterSmem<8, 8>), stream>>>(L0, P, entropy, pc, rows, cols, T, lam);
        OTGAN_CHECK_LAUNCH("sinkhorn_cluster_kernel<8,8>");
        return OTGAN_OK;
    }
    if (rows <= 512 && cols <= 512) {
        OTGAN_SET_MAX_SMEM((sinkhorn_cluster_kernel<16, 16>), sizeof(ClusterSmem<16, 16>));
        sinkhorn_cluster_kernel<16, 16><<<nblk * CL, 16 * 32, sizeof(ClusterSmem<16, 16>), stream>>>(L0, P, entropy, pc, rows, cols, T,
                                                                                                  lam);
        OTGAN_CHECK_LAUNCH("sinkhorn_cluster_kernel<16,16>");
        return OTGAN_OK;
    }
    set_error("sinkhorn_cluster: block %dx%d larger than %d", rows, cols, sinkhorn_cluster_max_side());
    return OTGAN_EUNSUPPORTED;
}

}  // namespace otgan
