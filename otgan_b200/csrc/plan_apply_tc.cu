// plan_apply_tc.cu -- transport-plan application on the tensor cores (TMA + tcgen05 kind::tf32, 3xTF32).
//
//   out[o] ([h, D]) = sum_t coef[o][t] * op(P[blk[o][t]]) * F[src[o][t]]          (utils/matching.py:63-83, train.py:111,125-126)
//
// GEMM view per output group o and 128-row tile:  D[M = 128 rows i][N = 128 columns d] += A[i][k] * B[k][d],  K = h per term
// (blocks larger than 128 -- h = 256 of the 64x64-image configuration, N x N single-batch blocks -- are more row tiles and more
// K chunks of the same kernel; the prepared A planes are [HP x HP] with HP = h rounded up to 128).
//   A = coef * op(P)  : tiny, reused by every column tile.  A prep kernel writes it once per call, zero-padded to
//                       128 x 128, already split into hi/lo TF32 planes and K-major (so transposes, the 0.5 averaging and
//                       the f_aa - f_ab subtraction cost nothing in the main loop).
//   B = F (fp32, [k][d], d contiguous): streamed once by TMA as [32 k x 32 d] boxes in the only shared-memory layout the
//                       tensor core accepts for MN-major 32-bit operands -- SWIZZLE_128B with 32-byte atoms (4-row groups,
//                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / UMMA layout type 1); split into hi/lo by four warps
//                       (layout-agnostic: same byte offsets in a second buffer).
// Persistent CTAs (one per SM) walk (output group, column tile) items; per item 3 terms x 4 K-chunks of 32 accumulate
// into a 128-column TMEM buffer (double buffered: the epilogue of one tile overlaps the MMAs of the next); the epilogue
// reads TMEM (lane = output row) and writes 128-byte row segments.  The accumulation chain is 144 MMAs on O(0.01) values
// (truncation bias < 2e-6 relative), so no intermediate drain is needed here (cf. cost_tc.cu, where K = 32768).
#include "tc_common.cuh"
#include <string.h>

namespace otgan {

namespace {

using namespace tc;

constexpr int BK = 32, TM_ = 128, TN_ = 128;
constexpr int A_TILE = TM_ * BK * 4;             // 16 KB  [128 i][32 k] K-major
constexpr int B_TILE = BK * TN_ * 4;             // 16 KB  4 boxes of [32 k][32 d]
constexpr int B_BOX = BK * 32 * 4;               // 4 KB
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;          // [A hi | A lo | B hi (raw) | B lo]
constexpr int NUM_SPLIT_THREADS = 128, NUM_EPI_THREADS = 128, NUM_THREADS = 64 + NUM_SPLIT_THREADS + NUM_EPI_THREADS;
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
constexpr uint32_t SW128 = 2, SW128_BASE32B = 1;

struct Params {
    CUtensorMap map_a_hi, map_a_lo;              // planes [n_out * 3 * 128, 128]
    CUtensorMap map_f[OTGAN_MAX_OUTPUTS];        // sources [h, D]
    otgan_plan_t plan;
    float* out[OTGAN_MAX_OUTPUTS];
    int h, hp, row_tiles, rt_lo, D, ldo, n_col_tiles, n_items, kchunks, box_bytes;   // row_tiles = tiles walked, starting at tile rt_lo
};

// Aop[(o*3 + t)*128 + i][k] = coef * op(P)[i][k] (zero padded), split into hi / lo TF32 planes
__global__ void plan_prep_kernel(otgan_plan_t plan, int h, int hp, const float* __restrict__ P, float* __restrict__ a_hi,
                                 float* __restrict__ a_lo)
{
    const int o = blockIdx.y, t = blockIdx.z;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over hp*hp
    const int i = idx / hp, k = idx - i * hp;
    float v = 0.f;
    if (t < plan.nterms[o] && i < h && k < h) {
        const float* Pm = P + (size_t)plan.blk[o][t] * h * h;
        v = plan.coef[o][t] * (plan.trans[o][t] ? Pm[(size_t)k * h + i] : Pm[(size_t)i * h + k]);
    }
    const uint32_t hi = cvt_rna_tf32(v);
    const uint32_t lo = cvt_rna_tf32(v - __uint_as_float(hi));
    const size_t off = ((size_t)(o * OTGAN_MAX_TERMS + t) * hp + i) * hp + k;
    a_hi[off] = __uint_as_float(hi);
    a_lo[off] = __uint_as_float(lo);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
plan_apply_tc_kernel(const __grid_constant__ Params p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto ready_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (3 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (3 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a_hi);
        tma_prefetch_desc(&p.map_a_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), NUM_SPLIT_THREADS);
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
    // B regions start as zeros: with h < 32 the TMA box has only h rows, and the K rows it never writes must contribute 0
    for (int i = threadIdx.x; i < STAGES * (2 * B_TILE / 16); i += NUM_THREADS) {
        const int s = i / (2 * B_TILE / 16), r = i % (2 * B_TILE / 16);
        reinterpret_cast<uint4*>(smem_gen + s * STAGE_BYTES + 2 * A_TILE)[r] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // every role walks the same item / stage sequence
    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int c = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int rt = p.rt_lo + item % p.row_tiles, oc = item / p.row_tiles;
                const int o = oc / p.n_col_tiles, d0 = (oc % p.n_col_tiles) * TN_;
                for (int t = 0; t < p.plan.nterms[o]; ++t) {
                    const CUtensorMap* mf = &p.map_f[p.plan.src[o][t]];
                    for (int kc = 0; kc < p.kchunks; ++kc, ++c) {
                        const int s = c % STAGES;
                        mbar_wait(empty_bar(s), (((uint32_t)(c / STAGES)) & 1u) ^ 1u);
                        int nbox = (p.D - d0 + 31) / 32;                    // column boxes that start inside the tensor
                        nbox = nbox > TN_ / 32 ? TN_ / 32 : nbox;
                        mbar_arrive_expect_tx(full_bar(s), 2 * A_TILE + nbox * p.box_bytes);
                        const uint32_t dst = smem_base + s * STAGE_BYTES;
                        const int arow = (o * OTGAN_MAX_TERMS + t) * p.hp + rt * TM_;
                        tma_load_2d(dst, &p.map_a_hi, full_bar(s), kc * BK, arow);
                        tma_load_2d(dst + A_TILE, &p.map_a_lo, full_bar(s), kc * BK, arow);
                        for (int j = 0; j < nbox; ++j)
                            tma_load_2d(dst + 2 * A_TILE + j * B_BOX, mf, full_bar(s), d0 + 32 * j, kc * BK);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 128, /*a_mn_major=*/0, /*b_mn_major=*/1);
            int c = 0, n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                const int o = (item / p.row_tiles) / p.n_col_tiles, b = n & 1;
                mbar_wait(tempty_bar(b), (((uint32_t)(n >> 1)) & 1u) ^ 1u);
                const uint32_t d = tmem_base + (uint32_t)(b * 128);
                const int nst = p.plan.nterms[o] * p.kchunks;
                for (int st = 0; st < nst; ++st, ++c) {
                    const int s = c % STAGES;
                    mbar_wait(ready_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a_hi0 = smem_base + s * STAGE_BYTES, a_lo0 = a_hi0 + A_TILE;
                    const uint32_t b_hi0 = a_hi0 + 2 * A_TILE, b_lo0 = b_hi0 + B_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        // A: K-major SW128 (rows of 128 B, 8-row groups 1024 B apart); K-step = +32 B
                        const uint64_t a_hi = umma_desc_kmajor(a_hi0 + k * 32, 1024, SW128);
                        const uint64_t a_lo = umma_desc_kmajor(a_lo0 + k * 32, 1024, SW128);
                        // B: MN-major SW128 with 32-byte atoms: 32-wide column blocks LBO = 4096 B apart, 4-row K groups
                        // SBO = 512 B apart (one K = 8 MMA spans two groups); K-step = 8 rows = +1024 B
                        const uint64_t b_hi = umma_desc_mnmajor(b_hi0 + k * 1024, B_BOX, 512, SW128_BASE32B);
                        const uint64_t b_lo = umma_desc_mnmajor(b_lo0 + k * 1024, B_BOX, 512, SW128_BASE32B);
                        umma_tf32(d, a_hi, b_lo, idesc, (st > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(d, a_lo, b_hi, idesc, 1u);
                        umma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(b));
            }
        }
    } else if (warp < 6) {
        // ===================================================== split warps: B (embedding) tile -> hi / lo
        const int st_id = threadIdx.x - 64;
        int c = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int o = (item / p.row_tiles) / p.n_col_tiles;
            const int nst = p.plan.nterms[o] * p.kchunks;
            for (int st = 0; st < nst; ++st, ++c) {
                const int s = c % STAGES;
                mbar_wait(full_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                float4* hi = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES + 2 * A_TILE);
                float4* lo = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES + 2 * A_TILE + B_TILE);
                float4 x[B_TILE / 16 / NUM_SPLIT_THREADS];                  // all 8 loads first, then the splits and the stores
#pragma unroll
                for (int j = 0; j < B_TILE / 16 / NUM_SPLIT_THREADS; ++j) x[j] = hi[st_id + j * NUM_SPLIT_THREADS];
#pragma unroll
                for (int j = 0; j < B_TILE / 16 / NUM_SPLIT_THREADS; ++j) {
                    uint4 hh, ll;
                    split_tf32(x[j].x, hh.x, ll.x); split_tf32(x[j].y, hh.y, ll.y); split_tf32(x[j].z, hh.z, ll.z); split_tf32(x[j].w, hh.w, ll.w);
                    reinterpret_cast<uint4*>(hi)[st_id + j * NUM_SPLIT_THREADS] = hh;
                    reinterpret_cast<uint4*>(lo)[st_id + j * NUM_SPLIT_THREADS] = ll;
                }
                fence_proxy_async_smem();
                mbar_arrive(ready_bar(s));
            }
        }
    } else {
        // ===================================================== epilogue warps: TMEM -> global
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        int n = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
            const int rt = p.rt_lo + item % p.row_tiles, oc = item / p.row_tiles;
            const int o = oc / p.n_col_tiles, d0 = (oc % p.n_col_tiles) * TN_, b = n & 1;
            const int row = rt * TM_ + m;
            mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
            tcgen05_fence_after();
            float* out = p.out[o] + (size_t)row * p.ldo + d0;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * 128 + cc * 32), r);
                tmem_ld_wait();
                if (row < p.h) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (d0 + cc * 32 + j < p.D)          // D % 4 == 0
                            *reinterpret_cast<uint4*>(out + cc * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(b));
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

}  // namespace

bool plan_apply_tc_supported(const otgan_plan_t* plan, int h, int D, const float* P, const float* const* F, int ldf,
                             float* const* out, int ldo)
{
    if (h < 1 || h > 4096 || D < 32 || (D & 3) || (ldf & 3) || (ldo & 3) || !aligned16(P)) return false;
    for (int o = 0; o < plan->n_out; ++o) {
        if (!aligned16(out[o])) return false;
        for (int t = 0; t < plan->nterms[o]; ++t)
            if (!aligned16(F[plan->src[o][t]])) return false;
    }
    return true;
}

size_t plan_apply_tc_workspace_bytes(int n_out, int h)
{
    const size_t hp = (size_t)ceil_div(h < 1 ? 1 : h, TM_) * TM_;
    return (size_t)2 * n_out * OTGAN_MAX_TERMS * hp * hp * sizeof(float) + 256;
}

int plan_apply_tc_launch(const otgan_plan_t* plan, int h, int D, const float* P, const float* const* F, int ldf,
                         float* const* out, int ldo, void* ws, size_t ws_bytes, cudaStream_t stream, int row_lo, int row_hi)
{
    OTGAN_REQUIRE(ws && ws_bytes >= plan_apply_tc_workspace_bytes(plan->n_out, h), "plan_apply(tcgen05): workspace too small");
    const int hp = ceil_div(h, TM_) * TM_;
    const size_t plane = (size_t)plan->n_out * OTGAN_MAX_TERMS * hp * hp;
    float* a_hi = reinterpret_cast<float*>(ws);
    float* a_lo = a_hi + plane;
    plan_prep_kernel<<<dim3(hp * hp / 256, plan->n_out, OTGAN_MAX_TERMS), 256, 0, stream>>>(*plan, h, hp, P, a_hi, a_lo);
    OTGAN_CHECK_LAUNCH("plan_prep_kernel");

    Params p;
    memset(&p, 0, sizeof(p));
    p.plan = *plan;
    const int arows = plan->n_out * OTGAN_MAX_TERMS * hp;
    if (!make_tensor_map_2d(&p.map_a_hi, a_hi, arows, hp, hp, 128, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    if (!make_tensor_map_2d(&p.map_a_lo, a_lo, arows, hp, hp, 128, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    bool used[OTGAN_MAX_OUTPUTS] = {false};
    for (int o = 0; o < plan->n_out; ++o) {
        p.out[o] = out[o];
        for (int t = 0; t < plan->nterms[o]; ++t) used[plan->src[o][t]] = true;
    }
    // box = [min(h,32) k x 32 d]; a box that starts inside the tensor but crosses its edge is zero-filled by TMA
    const int box_rows = h < BK ? h : BK;
    p.box_bytes = box_rows * 32 * 4;
    for (int s = 0; s < OTGAN_MAX_OUTPUTS; ++s)
        if (used[s] && !make_tensor_map_2d(&p.map_f[s], F[s], h, D, ldf, box_rows, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return OTGAN_EUNSUPPORTED;
    // rows [row_lo, row_hi) of every output group only (row_hi <= 0: all rows): whole 128-row tiles that intersect the range
    if (row_hi <= 0) { row_lo = 0; row_hi = h; }
    p.rt_lo = row_lo / TM_;
    p.h = h; p.hp = hp; p.row_tiles = ceil_div(row_hi, TM_) - p.rt_lo; p.D = D; p.ldo = ldo;
    p.n_col_tiles = ceil_div(D, TN_);
    p.n_items = plan->n_out * p.n_col_tiles * p.row_tiles;
    p.kchunks = ceil_div(h, BK);
    OTGAN_SET_MAX_SMEM((plan_apply_tc_kernel), SMEM_BYTES);
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    plan_apply_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
    OTGAN_CHECK_LAUNCH("plan_apply_tc_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
