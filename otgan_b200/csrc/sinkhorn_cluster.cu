// sinkhorn_cluster.cu -- persistent log-domain Sinkhorn for blocks of 128 < side <= 512 (utils/matching.py:46-57).
//
// A block of this size does not fit one SM's registers, and the streaming rung (sinkhorn_stream.cu) pays two launches per
// iteration.  Here a thread-block CLUSTER of 8 CTAs owns one block for all T iterations: CTA q keeps the row slab
// [q*R, (q+1)*R) of log_a in REGISTERS (warp = 4 rows, lane = CPL columns), so
//   * the row half-step (log_a -= LSE over axis 1, :53) is warp-local, exactly like sinkhorn.cu;
//   * the column half-step (:54) reduces in three levels: thread (4 rows) -> CTA (shared memory, (max, sum) pairs combined
//     online) -> cluster: every CTA pushes ONE float per column -- its slab's log-sum-exp -- into all eight CTAs' shared memory
//     with st.async (DSMEM store that reports its bytes to an mbarrier of the DESTINATION CTA), every CTA waits on its own
//     mbarrier for the 8 x HC floats and combines the eight slab values itself (LSE of LSEs is exact, no second exchange).  One
//     value per column keeps the push at 8 KB per CTA (DSMEM moves ~20 B/clk); buffer and mbarrier are doubled by iteration
//     parity, so the loop has no cluster-wide barrier at all (measured with barrier.cluster instead: 991 vs 534 cycles per
//     iteration for the exchange, 202 vs 182 us per call);
//   * HBM is touched twice: L0 in, P out (+ L0 once more for <P,C>), like the single-CTA kernels.
// Arithmetic = the literal rung's (max-subtracted LSE with MUFU ex2 / lg2 on differences); the update order of the reference is
// kept.  Two shapes: side <= 256 (8 warps, 8 columns per lane) and side <= 512 (16 warps, 16 columns per lane).
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace otgan {

namespace {

constexpr int CL = 8;                        // CTAs per cluster (portable maximum; 4 CTAs x 16 warps measured 270 vs 210 us)
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
    unsigned long long bar[2];               // mbarriers: "all eight slabs of parity p have landed in xch[p]"
};

// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// asynchronous 4-byte store into another CTA's shared memory that reports its bytes to an mbarrier of THAT CTA
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 :: "r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}

// Phase clocks (tools/sinkhorn_cluster_phases.cu compiles this file with OTGAN_SKC_CLOCKS): cycles of thread 0 of CTA 0, summed over the
// iterations, per phase of the loop.
#ifdef OTGAN_SKC_CLOCKS
__device__ long long g_skc_clk[8];
#define SKC_CLK_INIT long long clk_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long clk_last = clock64();
#define SKC_CLK(i) { const long long t__ = clock64(); clk_acc[i] += t__ - clk_last; clk_last = t__; }
#define SKC_CLK_STORE if (threadIdx.x == 0 && blockIdx.x == 0) { for (int i__ = 0; i__ < 8; ++i__) g_skc_clk[i__] = clk_acc[i__]; }
#else
#define SKC_CLK_INIT
#define SKC_CLK(i)
#define SKC_CLK_STORE
#endif

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

    // Exchange protocol: xch[p] / bar[p] serve the iterations of parity p.  bar[p] is armed with the byte count of one full exchange
    // (8 slabs x HC floats) BEFORE any store for that phase can be issued anywhere in the cluster: initially here, ahead of the
    // cluster barrier; afterwards right after the wait of iteration k, which precedes this CTA's own push of iteration k + 1, without
    // which no CTA can get past ITS wait of k + 1 and start pushing k + 2.  The same chain protects xch[p] against being overwritten
    // while it is still read.
    constexpr uint32_t XBYTES = CL * HC * sizeof(float);
    const uint32_t bar0 = tc::smem_u32(&sm.bar[0]);
    if (threadIdx.x == 0) {
        tc::mbar_init(bar0, 1);
        tc::mbar_init(bar0 + 8, 1);
        tc::fence_barrier_init();
        if (T >= 1) tc::mbar_arrive_expect_tx(bar0, XBYTES);
        if (T >= 2) tc::mbar_arrive_expect_tx(bar0 + 8, XBYTES);
    }
    cluster.sync();
    SKC_CLK_INIT
    for (int it = 0; it < T; ++it) {
        // ---- log_a -= reduce_logsumexp(log_a, axis=1)            utils/matching.py:53   (warp-local; the four rows' butterflies are
        // independent, which hides the shuffle latency better than a transposing reduction with fewer shuffles: measured)
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
        SKC_CLK(0)

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
        SKC_CLK(1)
        __syncthreads();
        SKC_CLK(2)
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
            const uint32_t dst = tc::smem_u32(&sm.xch[par][rank][c]);
#pragma unroll
            for (int q = 0; q < CL; ++q) st_async_f32(map_to_rank(dst, q), l, map_to_rank(bar0 + 8 * par, q));
        }
        SKC_CLK(3)
        tc::mbar_wait(bar0 + 8 * par, (it >> 1) & 1);
        if (threadIdx.x == 0 && it + 2 < T) tc::mbar_arrive_expect_tx(bar0 + 8 * par, XBYTES);
        SKC_CLK(4)
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
        SKC_CLK(5)
        __syncthreads();
        SKC_CLK(6)
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(&sm.lse[128 * g + 4 * lane]);
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                x[i][4 * g + 0] -= t.x; x[i][4 * g + 1] -= t.y; x[i][4 * g + 2] -= t.z; x[i][4 * g + 3] -= t.w;
            }
        }
        SKC_CLK(7)
    }

    SKC_CLK_STORE
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
        constexpr int NW = 256 / (CL * RPW);
        OTGAN_SET_MAX_SMEM((sinkhorn_cluster_kernel<NW, 8>), sizeof(ClusterSmem<NW, 8>));
        sinkhorn_cluster_kernel<NW, 8><<<nblk * CL, NW * 32, sizeof(ClusterSmem<NW, 8>), stream>>>(L0, P, entropy, pc, rows, cols, T, lam);
        OTGAN_CHECK_LAUNCH("sinkhorn_cluster_kernel<side 256>");
        return OTGAN_OK;
    }
    if (CL == 8 && rows <= 512 && cols <= 512) {
        OTGAN_SET_MAX_SMEM((sinkhorn_cluster_kernel<16, 16>), sizeof(ClusterSmem<16, 16>));
        sinkhorn_cluster_kernel<16, 16><<<nblk * CL, 16 * 32, sizeof(ClusterSmem<16, 16>), stream>>>(L0, P, entropy, pc, rows, cols, T,
                                                                                                  lam);
        OTGAN_CHECK_LAUNCH("sinkhorn_cluster_kernel<side 512>");
        return OTGAN_OK;
    }
    set_error("sinkhorn_cluster: block %dx%d larger than %d", rows, cols, sinkhorn_cluster_max_side());
    return OTGAN_EUNSUPPORTED;
}

}  // namespace otgan
