// cost_tc.cu -- pairwise-dot cost blocks on the 5th-gen tensor cores: TMA -> smem -> (hi/lo split) -> tcgen05.mma
// kind::tf32 (3xTF32, fp32-class accuracy) -> TMEM -> fp32 registers -> split-K partials.  Replaces the six
// tf.matmul(transpose_b=True) of utils/matching.py:29-43; the cost epilogue (1 - g, *(-lambda), +999 I, Euclidean form)
// is the shared finalize kernel.
//
// Work decomposition: grid = (splits, nblk * row tiles * column tiles).  A CTA owns one 128x128 tile of one block and a
// K-slice of D/splits; splits is chosen so that the grid fills the 148 SMs once (persistent over the K-slice).  Blocks larger
// than 128 (h = 256 of the 64x64-image configuration, N x N blocks of the single-batch variant) are simply more tiles: the TMA
// box of tile (ti, tj) starts at row 128 ti / 128 tj of the same tensor maps, ragged edges are zero-filled.  Tiles that share
// an embedding slab run concurrently on the same K range, so HBM sees every element once (the rest hits in L2).
//
// Accuracy.  (1) Operands: x = hi + lo with hi = rna_tf32(x), lo = x - hi (tc_common.cuh: split_tf32); D += Ahi*Blo + Alo*Bhi + Ahi*Bhi
// reproduces the fp32 product to 2^-22 (measured 2e-9 on the headline shape; a single TF32 pass is 1e-5 and would be
// amplified 500x by lambda).  (2) Accumulation: the tensor-core accumulator TRUNCATES (measured on B200: a K=32768 chain
// of positive products comes out biased by -3e-6 relative).  The TMEM accumulator is therefore restarted every K = 32 and
// drained into fp32 registers with round-to-nearest adds by dedicated epilogue warps (double-buffered TMEM, so the drain
// overlaps the next chunk's MMAs); the residual bias is < 1e-7 relative.
//
// Pipeline per CTA (320 threads), 3 smem stages of BK = 32 floats (128-byte rows, SWIZZLE_128B):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor of the raw fp32 tiles, completion on full[s]
//   warps 2-5: split          -- x -> hi (in place), lo (second buffer); layout-agnostic (same byte offsets), then
//                                 fence.proxy.async + arrive on ready[s]
//   warp 1   : MMA issuer     -- 4 K-steps x 3 products of tcgen05.mma (one thread) into TMEM buffer c&1,
//                                 tcgen05.commit -> empty[s] and -> tfull[c&1]
//   warps 6-9: epilogue       -- wait tfull, tcgen05.ld 128 columns, acc += (RN), arrive tempty; at the end acc -> partial
#include "tc_common.cuh"
#include <string.h>

namespace otgan {

// shared with cost_simt.cu
int cost_finalize_launch(const float* partial, int S, int nblk, int rows, int cols, int D, int cost_kind,
                         const float* const* X, const float* const* Y, int ldx, int ldy, const float* diag, float lam,
                         float* L, float* sq, cudaStream_t stream);

namespace {

using namespace tc;

constexpr int BK = 32;                         // floats per K-chunk: 128-byte rows == SWIZZLE_128B span
constexpr int TILE_ROWS = 128;
constexpr int TILE_BYTES = TILE_ROWS * BK * 4; // 16 KB
constexpr int MAX_MAPS = 2 * OTGAN_MAX_BLOCKS;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // [M hi | N hi | M lo | N lo]
constexpr int NUM_SPLIT_THREADS = 128, NUM_EPI_THREADS = 128, NUM_THREADS = 64 + NUM_SPLIT_THREADS + NUM_EPI_THREADS;
constexpr int TMEM_COLS = 256;                 // two 128-column accumulator buffers
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + 256;
constexpr uint32_t SWIZZLE_128B_CODE = 2;
constexpr uint32_t SBO_BYTES = 8 * BK * 4;     // 8 rows x 128 B

struct Params {
    CUtensorMap maps[MAX_MAPS];
    int map_m[OTGAN_MAX_BLOCKS], map_n[OTGAN_MAX_BLOCKS];   // tensor map of the M-side / N-side tile of each block
    int nblk, rows, cols, nchunks, chunks_per_split, tx_m, tx_n;
    int tiles_r, tiles_c;                                    // 128-row / 128-column tiles per block
    float* partial;                                          // [splits][nblk][rows][cols]
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
cost_tc_kernel(const __grid_constant__ Params p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms need 1024 B alignment
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
    const int split = blockIdx.x;
    const int blk = blockIdx.y / (p.tiles_r * p.tiles_c), tile = blockIdx.y % (p.tiles_r * p.tiles_c);
    const int ti = tile / p.tiles_c, tj = tile % p.tiles_c;
    const CUtensorMap* map_m = &p.maps[p.map_m[blk]];
    const CUtensorMap* map_n = &p.maps[p.map_n[blk]];
    const bool same_tile = p.map_m[blk] == p.map_n[blk] && ti == tj;   // X == Y (single-batch aa / bb blocks), diagonal tile: load once
    const int chunk0 = split * p.chunks_per_split;
    int nch = p.nchunks - chunk0;
    nch = nch > p.chunks_per_split ? p.chunks_per_split : nch;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(map_m);
        tma_prefetch_desc(map_n);
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
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const int ntiles = same_tile ? 1 : 2;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            for (int c = 0; c < nch; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), (uint32_t)(same_tile ? p.tx_m : p.tx_m + p.tx_n));
                const uint32_t dst = smem_base + s * STAGE_BYTES;
                tma_load_2d(dst, map_m, full_bar(s), (chunk0 + c) * BK, ti * TILE_ROWS);
                if (!same_tile) tma_load_2d(dst + TILE_BYTES, map_n, full_bar(s), (chunk0 + c) * BK, tj * TILE_ROWS);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 128);
            for (int c = 0; c < nch; ++c) {
                const int s = c % STAGES, b = c & 1;
                mbar_wait(ready_bar(s), (uint32_t)(c / STAGES) & 1u);
                mbar_wait(tempty_bar(b), ((uint32_t)(c >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t m_hi = smem_base + s * STAGE_BYTES, n_hi = same_tile ? m_hi : m_hi + TILE_BYTES;
                const uint32_t m_lo = m_hi + 2 * TILE_BYTES, n_lo = same_tile ? m_lo : m_lo + TILE_BYTES;
                const uint32_t d = tmem_base + (uint32_t)(b * 128);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {              // tf32 MMA K = 8 elements = 32 bytes
                    const uint64_t a_hi = umma_desc_kmajor(m_hi + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t a_lo = umma_desc_kmajor(m_lo + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t b_hi = umma_desc_kmajor(n_hi + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    const uint64_t b_lo = umma_desc_kmajor(n_lo + k * 32, SBO_BYTES, SWIZZLE_128B_CODE);
                    umma_tf32(d, a_hi, b_lo, idesc, k > 0 ? 1u : 0u);   // small terms first; chunk restarts at zero
                    umma_tf32(d, a_lo, b_hi, idesc, 1u);
                    umma_tf32(d, a_hi, b_hi, idesc, 1u);
                }
                umma_commit(empty_bar(s));                      // smem slot free once these MMAs have read it
                umma_commit(tfull_bar(b));                      // this chunk's partial product is complete in TMEM
            }
        }
    } else if (warp < 6) {
        // ===================================================== split warps: hi/lo transform
        const int st = threadIdx.x - 64;
        for (int c = 0; c < nch; ++c) {
            const int s = c % STAGES;
            mbar_wait(full_bar(s), (uint32_t)(c / STAGES) & 1u);
            float4* hi = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES + 2 * TILE_BYTES);
            // all loads of the chunk first (8 per tile and thread: one round trip to shared memory instead of a dependent
            // load -> split -> store chain per float4), then the splits and the stores
            const int nit = ntiles * (TILE_BYTES / 16 / NUM_SPLIT_THREADS);        // 8 or 16
            float4 x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nit) x[j] = hi[st + j * NUM_SPLIT_THREADS];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j < nit) {
                    uint4 h, l;
                    split_tf32(x[j].x, h.x, l.x); split_tf32(x[j].y, h.y, l.y); split_tf32(x[j].z, h.z, l.z); split_tf32(x[j].w, h.w, l.w);
                    reinterpret_cast<uint4*>(hi)[st + j * NUM_SPLIT_THREADS] = h;
                    reinterpret_cast<uint4*>(lo)[st + j * NUM_SPLIT_THREADS] = l;
                }
            }
            fence_proxy_async_smem();                           // generic-proxy writes -> visible to the tensor core
            mbar_arrive(ready_bar(s));
        }
    } else {
        // ===================================================== epilogue warps: drain TMEM every chunk, RN accumulate
        const int quad = warp & 3;                               // TMEM lanes [32*quad, 32*quad+32) belong to this warp
        const int m = quad * 32 + lane;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        for (int c = 0; c < nch; ++c) {
            const int b = c & 1;
            mbar_wait(tfull_bar(b), (uint32_t)(c >> 1) & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * 128 + cc * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(r[j]);
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
                    if (col0 + j < p.cols) *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 128; ++j)
                    if (col0 + j < p.cols) out[j] = acc[j];
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

// ---- host: tensor maps + split plan -----------------------------------------------------------------------------
struct Plan {
    Params prm;
    int splits;
};

int plan_splits(int ntiles, int D, int* chunks_per_split)
{
    const int nchunks = ceil_div(D, BK);
    int S = kNumSMs / ntiles;
    S = S < 1 ? 1 : (S > nchunks ? nchunks : S);
    const int cps = ceil_div(nchunks, S);
    if (chunks_per_split) *chunks_per_split = cps;
    return ceil_div(nchunks, cps);
}

bool build_plan(Plan& pl, int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx, int ldy)
{
    Params& p = pl.prm;
    memset(&p, 0, sizeof(p));
    struct Src { const float* ptr; int nrows, ld; };
    Src srcs[MAX_MAPS];
    int nsrc = 0;
    auto find_src = [&](const float* ptr, int nrows, int ld) -> int {
        for (int i = 0; i < nsrc; ++i)
            if (srcs[i].ptr == ptr && srcs[i].nrows == nrows && srcs[i].ld == ld) return i;
        srcs[nsrc] = {ptr, nrows, ld};
        return nsrc++;
    };
    for (int k = 0; k < nblk; ++k) {
        p.map_m[k] = find_src(X[k], rows, ldx);
        p.map_n[k] = find_src(Y[k], cols, ldy);
    }
    for (int i = 0; i < nsrc; ++i) {
        const int box_rows = srcs[i].nrows < TILE_ROWS ? srcs[i].nrows : TILE_ROWS;
        if (!make_tensor_map_2d(&p.maps[i], srcs[i].ptr, srcs[i].nrows, D, srcs[i].ld, box_rows, BK, CU_TENSOR_MAP_SWIZZLE_128B))
            return false;
    }
    p.tx_m = (rows < TILE_ROWS ? rows : TILE_ROWS) * BK * 4;
    p.tx_n = (cols < TILE_ROWS ? cols : TILE_ROWS) * BK * 4;
    p.nblk = nblk; p.rows = rows; p.cols = cols;
    p.tiles_r = ceil_div(rows, TILE_ROWS); p.tiles_c = ceil_div(cols, TILE_ROWS);
    p.nchunks = ceil_div(D, BK);
    pl.splits = plan_splits(nblk * p.tiles_r * p.tiles_c, D, &p.chunks_per_split);
    return true;
}

}  // namespace

bool cost_tc_supported(int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx, int ldy)
{
    if (rows < 1 || cols < 1 || (long long)nblk * ceil_div(rows, TILE_ROWS) * ceil_div(cols, TILE_ROWS) > 65535) return false;
    if ((ldx & 3) || (ldy & 3) || D < BK) return false;
    for (int k = 0; k < nblk; ++k)
        if (!aligned16(X[k]) || !aligned16(Y[k])) return false;
    return true;
}

size_t cost_tc_workspace_bytes(int nblk, int rows, int cols, int D)
{
    const int ntiles = nblk * ceil_div(rows, TILE_ROWS) * ceil_div(cols, TILE_ROWS);
    return (size_t)plan_splits(ntiles, D, nullptr) * nblk * rows * cols * sizeof(float) + (size_t)nblk * (rows + cols) * sizeof(float) + 256;
}

int cost_tc_launch(int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx, int ldy,
                   int cost_kind, const float* diag, float lam, float* L, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(ws_bytes >= cost_tc_workspace_bytes(nblk, rows, cols, D), "cost(tcgen05): workspace too small");
    Plan pl;
    if (!build_plan(pl, nblk, rows, cols, D, X, Y, ldx, ldy)) return OTGAN_EUNSUPPORTED;
    float* partial = reinterpret_cast<float*>(ws);
    float* sq = partial + (size_t)pl.splits * nblk * rows * cols;
    pl.prm.partial = partial;
    OTGAN_SET_MAX_SMEM((cost_tc_kernel), SMEM_BYTES);
    dim3 grid(pl.splits, nblk * pl.prm.tiles_r * pl.prm.tiles_c);
    cost_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(pl.prm);
    OTGAN_CHECK_LAUNCH("cost_tc_kernel");
    return cost_finalize_launch(partial, pl.splits, nblk, rows, cols, D, cost_kind, X, Y, ldx, ldy, diag, lam, L, sq, stream);
}

}  // namespace otgan
