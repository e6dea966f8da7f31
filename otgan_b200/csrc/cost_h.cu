// cost_h.cu -- pairwise-dot cost blocks on the tensor cores for UNIT-RANGE embeddings (|x| < 4: the critic head's L2-normalised
// rows): coalesced LDG -> registers -> fp16 hi/lo split -> shared memory -> tcgen05.mma kind::f16 -> TMEM -> fp32 registers.
// Same work decomposition, accumulation discipline and finalize kernel as cost_tc.cu (utils/matching.py:29-43); what changes is
// how the operands reach the tensor core.
//
// Why.  cost_tc.cu (TMA -> smem -> hi/lo TF32 split in smem -> 3 x kind::tf32) is bound by SHARED-MEMORY BANDWIDTH, not by the
// tensor pipe or HBM: per K = 32 it moves 224 KB through the 128 B/clk port (TMA write 32, split read 32 + write 64, MMA operand
// reads 96) = 1750 cycles against 768 cycles of MMA -- 0.34 of its roofline time.  Here
//   * x = h1 + h2 with h1 = the leading 11 bits of 2^14 x (Veltkamp split, exact) and h2 = fp16(2^14 x - h1): the same 22-bit
//     operand precision as the TF32 hi/lo pair at HALF the bytes, and kind::f16 runs at twice the TF32 rate;
//   * the fp32 tile never touches shared memory: converter warps load it straight from global memory (each 8-lane group reads one
//     128-byte row segment), split in registers and store only the two fp16 planes, already in the K-major SWIZZLE_128B layout.
// Per K = 32: 32 KB of stores + 48 KB of MMA operand reads = 80 KB = 625 cycles (2.8x less), 384 cycles of MMA.
// Accuracy: products h1*h1, h1*h2, h2*h1 are exact in the fp32 accumulator, the dropped h2*h2 term is 2^-22 relative (same class
// as 3xTF32); the scale 2^14 * 2^14 is undone exactly in the epilogue.  fp16 range: |2^14 x| <= 65504 needs |x| < 4 -- the
// caller's contract (OTGAN_IMPL_TCGEN05_UNIT); h2 of entries below 2^-17 of that range goes subnormal (absolute error 2^-39).
// The TMEM accumulator is restarted every K = 64 and drained into fp32 registers with round-to-nearest adds (cost_tc.cu, note 2).
//
// Pipeline per CTA (320 threads), 3 smem stages of K = 64 (fp16 rows of 128 bytes == the swizzle span):
//   warp 0   : MMA issuer  -- 4 K-steps x 3 products of tcgen05.mma (one thread) into TMEM buffer c&1, commit -> empty[s], tfull[c&1]
//   warps 2-5: converters  -- two K = 32 units per stage, software-pipelined: the loads of unit u+1 are in flight while unit u is split
//                             and stored; then fence.proxy.async + arrive on ready[s]
//   warps 6-9: epilogue    -- wait tfull, tcgen05.ld 128 columns, acc += (RN), arrive tempty; at the end acc * 2^-28 -> partial
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace otgan {

int cost_finalize_launch(const float* partial, int S, int nblk, int rows, int cols, int D, int cost_kind,
                         const float* const* X, const float* const* Y, int ldx, int ldy, const float* diag, float lam,
                         float* L, float* sq, cudaStream_t stream);

namespace {

using namespace tc;

constexpr int UK = 32;                         // floats per load unit (one 128-byte row segment)
constexpr int SK = 64;                         // K per smem stage: fp16 rows of 128 bytes
constexpr int TILE_ROWS = 128;
constexpr int PLANE_BYTES = TILE_ROWS * SK * 2;   // 16 KB
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;   // [M h1 | N h1 | M h2 | N h2]
constexpr int NUM_CONV_THREADS = 128, NUM_EPI_THREADS = 128, NUM_THREADS = 64 + NUM_CONV_THREADS + NUM_EPI_THREADS;
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
constexpr uint32_t SWIZZLE_128B_CODE = 2;
constexpr uint32_t SBO_BYTES = 8 * 128;
constexpr float IN_SCALE = 16384.f, OUT_SCALE = 1.f / (16384.f * 16384.f);

struct Params {
    const float* x[OTGAN_MAX_BLOCKS];
    const float* y[OTGAN_MAX_BLOCKS];
    int nblk, rows, cols, D, ldx, ldy, nstages, stages_per_split;
    int tiles_r, tiles_c;
    float* partial;                            // [splits][nblk][rows][cols]
};

// kind::f16, fp16 operands (format 0), fp32 accumulate, both K-major: bit layout in tc_common.cuh
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// 2^14 x = h1 + h2 (+ 2^-22): h1 = leading 11 bits (Veltkamp, factor 2^13 + 1), both packed as fp16 pairs
__device__ __forceinline__ void split4(const float4 x, uint2& h1, uint2& h2) {
    const float xs[4] = {x.x * IN_SCALE, x.y * IN_SCALE, x.z * IN_SCALE, x.w * IN_SCALE};
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float c = __fmul_rn(xs[e], 8193.f);
        hi[e] = __fsub_rn(c, __fsub_rn(c, xs[e]));
        lo[e] = __fsub_rn(xs[e], hi[e]);
    }
    const __half2 a0 = __floats2half2_rn(hi[0], hi[1]), a1 = __floats2half2_rn(hi[2], hi[3]);
    const __half2 b0 = __floats2half2_rn(lo[0], lo[1]), b1 = __floats2half2_rn(lo[2], lo[3]);
    h1 = make_uint2(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
    h2 = make_uint2(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
cost_h_kernel(const __grid_constant__ Params p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    auto ready_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int blk = blockIdx.y / (p.tiles_r * p.tiles_c), tile = blockIdx.y % (p.tiles_r * p.tiles_c);
    const int ti = tile / p.tiles_c, tj = tile % p.tiles_c;
    const bool same_tile = p.x[blk] == p.y[blk] && p.ldx == p.ldy && ti == tj;   // X == Y, diagonal tile: one operand
    const int stage0 = split * p.stages_per_split;
    int nst = p.nstages - stage0;
    nst = nst > p.stages_per_split ? p.stages_per_split : nst;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(ready_bar(s), NUM_CONV_THREADS);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), NUM_EPI_THREADS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 128);
            for (int c = 0; c < nst; ++c) {
                const int s = c % STAGES, b = c & 1;
                mbar_wait(ready_bar(s), (uint32_t)(c / STAGES) & 1u);
                mbar_wait(tempty_bar(b), ((uint32_t)(c >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t m_h1 = smem_base + s * STAGE_BYTES, n_h1 = same_tile ? m_h1 : m_h1 + PLANE_BYTES;
                const uint32_t m_h2 = m_h1 + 2 * PLANE_BYTES, n_h2 = same_tile ? m_h2 : m_h2 + PLANE_BYTES;
                const uint32_t d = tmem_base + (uint32_t)(b * 128);
#pragma unroll
                for (int k = 0; k < SK / 16; ++k) {             // f16 MMA K = 16 elements = 32 bytes
                    const uint64_t a1 = umma_desc_kmajor(m_h1 + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t a2 = umma_desc_kmajor(m_h2 + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t b1 = umma_desc_kmajor(n_h1 + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t b2 = umma_desc_kmajor(n_h2 + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    umma_f16(d, a1, b2, idesc, k > 0 ? 1u : 0u);        // small terms first; the stage restarts at zero
                    umma_f16(d, a2, b1, idesc, 1u);
                    umma_f16(d, a1, b1, idesc, 1u);
                }
                umma_commit(empty_bar(s));
                umma_commit(tfull_bar(b));
            }
        }
    } else if (warp >= 2 && warp < 6) {
        // ===================================================== converter warps
        // Load i of a unit (16 per thread): tile = i >> 3 (0: M side, 1: N side), row = 16 (i & 7) + prow, float4 column c4 of the
        // 128-byte segment.  8 consecutive lanes read one full segment (coalesced); the two rows of a half-warp differ in bit 2, so
        // the 64-bit stores of a half-warp cover all 32 banks once (row & 7 selects the 16-byte chunk permutation of SWIZZLE_128B).
        const int t = threadIdx.x - 64;
        const int q = t >> 3, c4 = t & 7;
        const int prow = ((q & 1) << 2) | ((q >> 1) & 3) | ((q >> 3) << 3);
        const int ntiles = same_tile ? 1 : 2;
        const float* gsrc[2];
        int rvalid[2];
        gsrc[0] = p.x[blk] + (size_t)(ti * TILE_ROWS + prow) * p.ldx + 4 * c4;
        gsrc[1] = p.y[blk] + (size_t)(tj * TILE_ROWS + prow) * p.ldy + 4 * c4;
        rvalid[0] = p.rows - ti * TILE_ROWS - prow;                // row 16 j + prow is inside the tensor iff 16 j < rvalid
        rvalid[1] = p.cols - tj * TILE_ROWS - prow;
        const size_t rstep[2] = {(size_t)16 * p.ldx, (size_t)16 * p.ldy};
        // byte offset of this thread's 8-byte slot inside a plane, for half h of the stage: row * 128 + ((chunk ^ (row & 7)) << 4) + ...
        uint32_t soff[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int chunk = 4 * h + (c4 >> 1);
            soff[h] = (uint32_t)(prow * 128 + ((chunk ^ (prow & 7)) << 4) + ((c4 & 1) << 3));
        }
        // Work items: half-units = (stage c, sub = 2 * half + tile), 8 float4 per thread each; three rotating register buffers keep two
        // half-units (16 loads per thread, 32 KB per SM) in flight while a third is split and stored.  Unrolled by 12 = lcm(3, 4)
        // so that buffer and sub-index are compile-time constants.
        float4 buf[3][8];
        const int unit0 = stage0 * 2;
        const int total = 4 * nst;
        auto load_half = [&](float4 (&r)[8], int hu) {
            const int c = hu >> 2, sub = hu & 3, tl = sub & 1;
            const int unit = unit0 + 2 * c + (sub >> 1);
            const bool on = hu < total && tl < ntiles && unit * UK + 4 * c4 < p.D;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (on && 16 * j < rvalid[tl]) r[j] = ldg_stream(gsrc[tl] + j * rstep[tl] + (size_t)unit * UK);
            }
        };
        auto store_half = [&](const float4 (&r)[8], int hu) {
            const int c = hu >> 2, sub = hu & 3, tl = sub & 1, h = sub >> 1;
            const int s = c % STAGES;
            if (sub == 0) mbar_wait(empty_bar(s), ((uint32_t)(c / STAGES) & 1u) ^ 1u);
            if (tl < ntiles) {
                uint8_t* st = smem_gen + s * STAGE_BYTES + tl * PLANE_BYTES + soff[h];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint2 h1, h2;
                    split4(r[j], h1, h2);
                    *reinterpret_cast<uint2*>(st + j * (16 * 128)) = h1;
                    *reinterpret_cast<uint2*>(st + j * (16 * 128) + 2 * PLANE_BYTES) = h2;
                }
            }
            if (sub == 3) {
                fence_proxy_async_smem();
                mbar_arrive(ready_bar(s));
            }
        };
        load_half(buf[0], 0);
        load_half(buf[1], 1);
        for (int base = 0; base < total; base += 12) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int hu = base + j;
                load_half(buf[(j + 2) % 3], hu + 2);
                if (hu < total) store_half(buf[j % 3], hu);
            }
        }
    } else if (warp >= 6) {
        // ===================================================== epilogue warps: drain TMEM every stage, RN accumulate
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        for (int c = 0; c < nst; ++c) {
            const int b = c & 1;
            mbar_wait(tfull_bar(b), (uint32_t)(c >> 1) & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {                       // 16 columns at a time: 128 accumulators + 16 in flight fit the register cap
                uint32_t r[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * 128 + cc * 16), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[cc * 16 + j] += __uint_as_float(r[j]);
            }
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(b));
        }
        const int row = ti * TILE_ROWS + m, col0 = tj * TILE_ROWS;
        if (row < p.rows) {
            float* out = p.partial + (((size_t)split * p.nblk + blk) * p.rows + row) * p.cols + col0;
            if ((p.cols & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 128; j += 4)
                    if (col0 + j < p.cols)
                        *reinterpret_cast<float4*>(out + j) =
                            make_float4(acc[j] * OUT_SCALE, acc[j + 1] * OUT_SCALE, acc[j + 2] * OUT_SCALE, acc[j + 3] * OUT_SCALE);
            } else {
#pragma unroll
                for (int j = 0; j < 128; ++j)
                    if (col0 + j < p.cols) out[j] = acc[j] * OUT_SCALE;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int plan_splits(int ntiles, int D, int* stages_per_split)
{
    const int nstages = ceil_div(D, SK);
    int S = kNumSMs / ntiles;
    S = S < 1 ? 1 : (S > nstages ? nstages : S);
    const int sps = ceil_div(nstages, S);
    if (stages_per_split) *stages_per_split = sps;
    return ceil_div(nstages, sps);
}

}  // namespace

bool cost_h_supported(int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx, int ldy)
{
    if (rows < 1 || cols < 1 || (long long)nblk * ceil_div(rows, TILE_ROWS) * ceil_div(cols, TILE_ROWS) > 65535) return false;
    if ((ldx & 3) || (ldy & 3) || (D & 3) || D < UK) return false;
    for (int k = 0; k < nblk; ++k)
        if (!aligned16(X[k]) || !aligned16(Y[k])) return false;
    return true;
}

size_t cost_h_workspace_bytes(int nblk, int rows, int cols, int D)
{
    const int ntiles = nblk * ceil_div(rows, TILE_ROWS) * ceil_div(cols, TILE_ROWS);
    return (size_t)plan_splits(ntiles, D, nullptr) * nblk * rows * cols * sizeof(float) + (size_t)nblk * (rows + cols) * sizeof(float) + 256;
}

int cost_h_launch(int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx, int ldy,
                  int cost_kind, const float* diag, float lam, float* L, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(ws_bytes >= cost_h_workspace_bytes(nblk, rows, cols, D), "cost(tcgen05 f16): workspace too small");
    Params p;
    memset(&p, 0, sizeof(p));
    for (int k = 0; k < nblk; ++k) { p.x[k] = X[k]; p.y[k] = Y[k]; }
    p.nblk = nblk; p.rows = rows; p.cols = cols; p.D = D; p.ldx = ldx; p.ldy = ldy;
    p.tiles_r = ceil_div(rows, TILE_ROWS); p.tiles_c = ceil_div(cols, TILE_ROWS);
    p.nstages = ceil_div(D, SK);
    const int splits = plan_splits(nblk * p.tiles_r * p.tiles_c, D, &p.stages_per_split);
    float* partial = reinterpret_cast<float*>(ws);
    float* sq = partial + (size_t)splits * nblk * rows * cols;
    p.partial = partial;
    OTGAN_SET_MAX_SMEM((cost_h_kernel), SMEM_BYTES);
    dim3 grid(splits, nblk * p.tiles_r * p.tiles_c);
    cost_h_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
    OTGAN_CHECK_LAUNCH("cost_h_kernel");
    return cost_finalize_launch(partial, splits, nblk, rows, cols, D, cost_kind, X, Y, ldx, ldy, diag, lam, L, sq, stream);
}

}  // namespace otgan
