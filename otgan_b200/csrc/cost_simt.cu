// cost_simt.cu -- pairwise-dot cost blocks on the FP32 FMA pipe (exact-fp32 rung of the cost stage).
//
// Computes L[k] = -lam * (cost(X_k, Y_k) + diag_k I) for up to 8 blocks, replacing the tf.matmul(transpose_b=True)
// graph of utils/matching.py:21-43 / :101-111 and toy_example/matching_cpu.py:10-49 plus the -lambda scaling of :50.
//
// Shape of the work: M = N = h (<= a few hundred) but K = D up to 131072, i.e. a tall-K Gram matrix with only a handful
// of output tiles.  The kernel therefore splits K across CTAs (grid sized to ~2 CTAs per SM), each CTA producing a
// 128x128 partial tile into a workspace, and a second kernel reduces the partials in a FIXED order (deterministic, no
// float atomics) and applies the cost epilogue.  HBM traffic = each embedding row read once per block it appears in.
#include "common.cuh"

namespace otgan {

namespace {

constexpr int TM = 128, TN = 128, BK = 32, LDK = BK + 4, STAGES = 3, NT = 256;
constexpr int STAGE_FLOATS = (TM + TN) * LDK;
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_FLOATS * sizeof(float);

struct CostArgs {
    const float* x[OTGAN_MAX_BLOCKS];
    const float* y[OTGAN_MAX_BLOCKS];
    float diag[OTGAN_MAX_BLOCKS];
};

// loads one [128 x BK] K-major tile (rows r0.., k range [k0, kend)) into smem laid out [row][LDK]
template <bool VEC>
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, int ld, int r0, int nrows, int k0,
                                          int kend, int tid)
{
    if (VEC) {
#pragma unroll
        for (int q = 0; q < (TM * BK / 4) / NT; ++q) {
            const int idx = tid + q * NT, row = idx >> 3, c4 = idx & 7;
            const int gr = r0 + row, k = k0 + c4 * 4;
            int bytes = (kend - k) * 4;
            bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
            if (gr >= nrows) bytes = 0;
            const float* src = g + (size_t)(gr < nrows ? gr : 0) * ld + (bytes > 0 ? k : 0);
            cp_async16(s + row * LDK + c4 * 4, src, bytes);
        }
    } else {
#pragma unroll 4
        for (int q = 0; q < (TM * BK) / NT; ++q) {
            const int idx = tid + q * NT, row = idx >> 5, c = idx & 31;
            const int gr = r0 + row, k = k0 + c;
            const bool ok = (gr < nrows) && (k < kend);
            const float* src = g + (ok ? (size_t)gr * ld + k : 0);
            cp_async4(s + row * LDK + c, src, ok ? 4 : 0);
        }
    }
}

// grid = (S, tiles_m * tiles_n, nblk); partial[(s * nblk + blk) * rows * cols + i * cols + j]
template <bool VEC>
__global__ void __launch_bounds__(NT, 2)
cost_gram_splitk_kernel(CostArgs args, int rows, int cols, int D, int ldx, int ldy, int ktiles_per_split, int tiles_n,
                        float* __restrict__ partial)
{
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int split = blockIdx.x, blk = blockIdx.z;
    const int m0 = (blockIdx.y / tiles_n) * TM, n0 = (blockIdx.y % tiles_n) * TN;
    const float* __restrict__ X = args.x[blk];
    const float* __restrict__ Y = args.y[blk];
    const int kbeg = split * ktiles_per_split * BK;
    int kend = kbeg + ktiles_per_split * BK;
    kend = kend > D ? D : kend;
    const int nkt = (kend - kbeg + BK - 1) / BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) {
            float* st = smem + s * STAGE_FLOATS;
            load_tile<VEC>(st, X, ldx, m0, rows, kbeg + s * BK, kend, tid);
            load_tile<VEC>(st + TM * LDK, Y, ldy, n0, cols, kbeg + s * BK, kend, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < nkt) {
                float* st = smem + (nk % STAGES) * STAGE_FLOATS;
                load_tile<VEC>(st, X, ldx, m0, rows, kbeg + nk * BK, kend, tid);
                load_tile<VEC>(st + TM * LDK, Y, ldy, n0, cols, kbeg + nk * BK, kend, tid);
            }
            cp_async_commit();
        }
        const float* xs = smem + (kt % STAGES) * STAGE_FLOATS;
        const float* ys = xs + TM * LDK;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(xs + (ty + 16 * i) * LDK + kk);
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(ys + (tx + 16 * j) * LDK + kk);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();
    float* out = partial + ((size_t)split * gridDim.z + blk) * rows * cols;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty + 16 * i;
        if (r < rows) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = n0 + tx + 16 * j;
                if (c < cols) out[(size_t)r * cols + c] = acc[i][j];
            }
        }
    }
}

// one warp per row: sq[r] = mean_d x[r][d]^2 for the X rows then the Y rows of every block (Euclidean cost only)
__global__ void row_sqmean_kernel(CostArgs args, int nblk, int rows, int cols, int D, int ldx, int ldy, float* __restrict__ sq)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int per_blk = rows + cols;
    const int blk = warp / per_blk, r = warp % per_blk;
    if (blk >= nblk) return;
    const float* p = (r < rows) ? args.x[blk] + (size_t)r * ldx : args.y[blk] + (size_t)(r - rows) * ldy;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(p[d], p[d], s);
    s = warp_sum(s);
    if (lane == 0) sq[blk * per_blk + r] = s / (float)D;
}

// L = -lam * (cost + diag*I); partial sums reduced in the fixed order s = 0..S-1
__global__ void cost_finalize_kernel(const float* __restrict__ partial, int S, int nblk, int rows, int cols, int D,
                                     int cost_kind, CostArgs args, const float* __restrict__ sq, float lam,
                                     float* __restrict__ L)
{
    const size_t per_blk = (size_t)rows * cols, total = per_blk * nblk;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        // fixed summation order s = 0..S-1; the loads of a batch of 8 partials are issued together (the serial load -> add chain
        // made this kernel 9 us for 9 MB: one L2 round trip per partial)
        float g = 0.f;
        int s = 0;
        for (; s + 8 <= S; s += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(partial + (size_t)(s + j) * total + e);
#pragma unroll
            for (int j = 0; j < 8; ++j) g += v[j];
        }
        for (; s < S; ++s) g += __ldg(partial + (size_t)s * total + e);
        const int blk = (int)(e / per_blk);
        const int rem = (int)(e % per_blk), i = rem / cols, j = rem % cols;
        float c;
        if (cost_kind == OTGAN_COST_COSINE) {
            c = __fsub_rn(1.0f, g);                                             // utils/matching.py:31 `1. - matmul`
        } else {
            const float xs = 0.5f * sq[blk * (rows + cols) + i], ys = 0.5f * sq[blk * (rows + cols) + rows + j];
            c = __fsub_rn(__fadd_rn(xs, ys), __fdiv_rn(g, (float)D));          // toy_example/matching_cpu.py:17-21
        }
        if (i == j && args.diag[blk] != 0.f) c = __fadd_rn(c, args.diag[blk]);  // utils/matching.py:109 `+ 999*eye`
        L[e] = __fmul_rn(-lam, c);                                              // utils/matching.py:50
    }
}

struct Split { int S, ktiles_per_split, tiles_m, tiles_n; };
Split plan_split(int nblk, int rows, int cols, int D)
{
    Split p;
    p.tiles_m = ceil_div(rows, TM);
    p.tiles_n = ceil_div(cols, TN);
    const int ktiles = ceil_div(D, BK);
    const int tiles = nblk * p.tiles_m * p.tiles_n;
    int S = (2 * kNumSMs) / (tiles > 0 ? tiles : 1);     // ~2 resident CTAs per SM
    S = S < 1 ? 1 : S;
    S = S > ktiles ? ktiles : S;
    p.ktiles_per_split = ceil_div(ktiles, S);
    p.S = ceil_div(ktiles, p.ktiles_per_split);
    return p;
}

}  // namespace

// shared by the SIMT and tcgen05 cost paths: fixed-order split-K reduction + cost epilogue
int cost_finalize_launch(const float* partial, int S, int nblk, int rows, int cols, int D, int cost_kind,
                         const float* const* X, const float* const* Y, int ldx, int ldy, const float* diag, float lam,
                         float* L, float* sq, cudaStream_t stream)
{
    CostArgs args;
    for (int k = 0; k < OTGAN_MAX_BLOCKS; ++k) {
        args.x[k] = k < nblk ? X[k] : nullptr;
        args.y[k] = k < nblk ? Y[k] : nullptr;
        args.diag[k] = (k < nblk && diag) ? diag[k] : 0.f;
    }
    if (cost_kind == OTGAN_COST_EUCLID_MEAN) {
        const int warps = nblk * (rows + cols);
        row_sqmean_kernel<<<ceil_div(warps * 32, 256), 256, 0, stream>>>(args, nblk, rows, cols, D, ldx, ldy, sq);
        OTGAN_CHECK_LAUNCH("row_sqmean_kernel");
    }
    const size_t total = (size_t)nblk * rows * cols;
    int fgrid = (int)((total + 255) / 256);
    fgrid = fgrid > 4 * kNumSMs ? 4 * kNumSMs : fgrid;
    cost_finalize_kernel<<<fgrid, 256, 0, stream>>>(partial, S, nblk, rows, cols, D, cost_kind, args, sq, lam, L);
    OTGAN_CHECK_LAUNCH("cost_finalize_kernel");
    return OTGAN_OK;
}

size_t cost_simt_workspace_bytes(int nblk, int rows, int cols, int D)
{
    const Split p = plan_split(nblk, rows, cols, D);
    return (size_t)p.S * nblk * rows * cols * sizeof(float) + (size_t)nblk * (rows + cols) * sizeof(float) + 256;
}

int cost_simt_launch(int nblk, int rows, int cols, int D, const float* const* X, const float* const* Y, int ldx,
                     int ldy, int cost_kind, const float* diag, float lam, float* L, void* ws, size_t ws_bytes,
                     cudaStream_t stream)
{
    OTGAN_REQUIRE(ws_bytes >= cost_simt_workspace_bytes(nblk, rows, cols, D), "cost: workspace too small (%zu < %zu)",
                  ws_bytes, cost_simt_workspace_bytes(nblk, rows, cols, D));
    CostArgs args;
    bool vec = (ldx % 4 == 0) && (ldy % 4 == 0);
    for (int k = 0; k < OTGAN_MAX_BLOCKS; ++k) {
        args.x[k] = k < nblk ? X[k] : nullptr;
        args.y[k] = k < nblk ? Y[k] : nullptr;
        args.diag[k] = (k < nblk && diag) ? diag[k] : 0.f;
        if (k < nblk) vec = vec && aligned16(X[k]) && aligned16(Y[k]);
    }
    const Split p = plan_split(nblk, rows, cols, D);
    float* partial = reinterpret_cast<float*>(ws);
    float* sq = partial + (size_t)p.S * nblk * rows * cols;
    OTGAN_SET_MAX_SMEM((cost_gram_splitk_kernel<true>), SMEM_BYTES);
    OTGAN_SET_MAX_SMEM((cost_gram_splitk_kernel<false>), SMEM_BYTES);
    dim3 grid(p.S, p.tiles_m * p.tiles_n, nblk);
    if (vec)
        cost_gram_splitk_kernel<true><<<grid, NT, SMEM_BYTES, stream>>>(args, rows, cols, D, ldx, ldy, p.ktiles_per_split, p.tiles_n, partial);
    else
        cost_gram_splitk_kernel<false><<<grid, NT, SMEM_BYTES, stream>>>(args, rows, cols, D, ldx, ldy, p.ktiles_per_split, p.tiles_n, partial);
    OTGAN_CHECK_LAUNCH("cost_gram_splitk_kernel");
    return cost_finalize_launch(partial, p.S, nblk, rows, cols, D, cost_kind, X, Y, ldx, ldy, diag, lam, L, sq, stream);
}

}  // namespace otgan
