// plan_apply_h.cu -- transport-plan application on the tensor cores for UNIT-RANGE embeddings (|x| < 4: L2-normalised rows):
// fp16 hi/lo operand planes + tcgen05.mma kind::f16 instead of plan_apply_tc.cu's 3xTF32 with an in-smem split.
//
//   out[o] ([h, D]) = sum_t coef[o][t] * op(P[blk[o][t]]) * F[src[o][t]]          (utils/matching.py:63-83, train.py:111,125-126)
//
// Same GEMM view as plan_apply_tc.cu (per output group o and 128-row tile: D[128 rows i][128 columns d] += A[i][k] B[k][d], K = h per
// term), same persistent walk over (output group, column tile, row tile) items; what changes is how the operands reach the tensor core:
//   A = coef * op(P): the prep kernel writes 2^14 A as two fp16 planes (h1 = leading 11 bits, h2 = remainder: 22 bits together, the
//       precision of a TF32 hi/lo pair at half the bytes), K-major; TMA brings [128 x 64] chunks (128-byte rows, SWIZZLE_128B).
//   B = F ([k][d], d contiguous): converter warps read it straight from global memory -- lane = one column d, eight consecutive k per
//       16-byte store, i.e. the TRANSPOSE to the K-major layout happens in registers -- split it in registers and store only the two
//       fp16 planes [128 d][64 k] (SWIZZLE_128B).  The fp32 tile never touches shared memory and both operands are plain K-major.
// plan_apply_tc.cu moves 1.9 MB per item through the 128 B/clk shared-memory port (TMA write, split read + write, three TF32 operand
// reads) for 9.2 K cycles of MMA; here it is 0.77 MB for 4.6 K cycles (kind::f16 runs at twice the TF32 rate).
// Contract (OTGAN_IMPL_TCGEN05_UNIT): |F| < 4; P is a plan (entries in [0, 1]) and |coef| <= 1.  The 2^28 scale is undone exactly in
// the epilogue; products h1*h1, h1*h2, h2*h1 are exact in the fp32 accumulator, the dropped h2*h2 term is 2^-22 relative.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace otgan {

namespace {

using namespace tc;

constexpr int SK = 64, TM_ = 128, TN_ = 128;
constexpr int PLANE = TM_ * SK * 2;              // 16 KB: [128 rows][64 k] fp16
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * PLANE;           // [A h1 | A h2 | B h1 | B h2]
constexpr int NUM_CONV_THREADS = 128, NUM_EPI_THREADS = 128, NUM_THREADS = 64 + NUM_CONV_THREADS + NUM_EPI_THREADS;
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
constexpr uint32_t SW128 = 2;
constexpr float IN_SCALE = 16384.f, OUT_SCALE = 1.f / (16384.f * 16384.f);

struct Params {
    CUtensorMap map_a1, map_a2;                  // fp16 planes [n_out * 3 * hp, hp] described as fp32 [.., hp / 2]
    const float* f[OTGAN_MAX_OUTPUTS];           // sources [h, D]
    otgan_plan_t plan;
    float* out[OTGAN_MAX_OUTPUTS];
    int h, hp, row_tiles, rt_lo, D, ldf, ldo, n_col_tiles, n_items, kchunks;
};

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
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// 2^14 x = hi + lo (+ 2^-22): hi = leading 11 bits (Veltkamp, factor 2^13 + 1): both exactly / correctly rounded fp16
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
    const float xs = x * IN_SCALE;
    const float c = __fmul_rn(xs, 8193.f);
    hi = __fsub_rn(c, __fsub_rn(c, xs));
    lo = __fsub_rn(xs, hi);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// Aop[(o*3 + t)*hp + i][k] = 2^14 coef * op(P)[i][k] (zero padded) as two fp16 planes
__global__ void plan_prep_h_kernel(otgan_plan_t plan, int h, int hp, const float* __restrict__ P, __half* __restrict__ a1,
                                   __half* __restrict__ a2)
{
    const int o = blockIdx.y, t = blockIdx.z;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over hp*hp
    const int i = idx / hp, k = idx - i * hp;
    float v = 0.f;
    if (t < plan.nterms[o] && i < h && k < h) {
        const float* Pm = P + (size_t)plan.blk[o][t] * h * h;
        v = plan.coef[o][t] * (plan.trans[o][t] ? Pm[(size_t)k * h + i] : Pm[(size_t)i * h + k]);
    }
    float hi, lo;
    split1(v, hi, lo);
    const size_t off = ((size_t)(o * OTGAN_MAX_TERMS + t) * hp + i) * hp + k;
    a1[off] = __float2half_rn(hi);
    a2[off] = __float2half_rn(lo);
}

struct Walk { int item, t, kc; };

__global__ void __launch_bounds__(NUM_THREADS, 1)
plan_apply_h_kernel(const __grid_constant__ Params p)
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
        tma_prefetch_desc(&p.map_a1);
        tma_prefetch_desc(&p.map_a2);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
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

    auto item_o = [&](int item) { return (item / p.row_tiles) / p.n_col_tiles; };

    // every role walks the same item / stage sequence
    if (warp == 0) {
        // ===================================================== TMA producer: the A chunks
        if (lane == 0) {
            int c = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int rt = p.rt_lo + item % p.row_tiles, o = item_o(item);
                for (int t = 0; t < p.plan.nterms[o]; ++t) {
                    for (int kc = 0; kc < p.kchunks; ++kc, ++c) {
                        const int s = c % STAGES;
                        mbar_wait(empty_bar(s), (((uint32_t)(c / STAGES)) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(full_bar(s), 2 * PLANE);
                        const uint32_t dst = smem_base + s * STAGE_BYTES;
                        const int arow = (o * OTGAN_MAX_TERMS + t) * p.hp + rt * TM_;
                        tma_load_2d(dst, &p.map_a1, full_bar(s), kc * (SK / 2), arow);          // coordinates in fp32 units
                        tma_load_2d(dst + PLANE, &p.map_a2, full_bar(s), kc * (SK / 2), arow);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 128);
            int c = 0, n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                const int o = item_o(item), b = n & 1;
                mbar_wait(tempty_bar(b), (((uint32_t)(n >> 1)) & 1u) ^ 1u);
                const uint32_t d = tmem_base + (uint32_t)(b * 128);
                const int nst = p.plan.nterms[o] * p.kchunks;
                for (int st = 0; st < nst; ++st, ++c) {
                    const int s = c % STAGES;
                    mbar_wait(full_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                    mbar_wait(ready_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a1_0 = smem_base + s * STAGE_BYTES, a2_0 = a1_0 + PLANE, b1_0 = a1_0 + 2 * PLANE, b2_0 = a1_0 + 3 * PLANE;
#pragma unroll
                    for (int k = 0; k < SK / 16; ++k) {          // f16 MMA K = 16 elements = 32 bytes
                        const uint64_t a1 = umma_desc_kmajor(a1_0 + k * 32, 1024, SW128), a2 = umma_desc_kmajor(a2_0 + k * 32, 1024, SW128);
                        const uint64_t b1 = umma_desc_kmajor(b1_0 + k * 32, 1024, SW128), b2 = umma_desc_kmajor(b2_0 + k * 32, 1024, SW128);
                        umma_f16(d, a1, b2, idesc, (st > 0 || k > 0) ? 1u : 0u);
                        umma_f16(d, a2, b1, idesc, 1u);
                        umma_f16(d, a1, b1, idesc, 1u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(b));
            }
        }
    } else if (warp < 6) {
        // ===================================================== converter warps: F tile -> fp16 planes, transposed to K-major
        // Thread = one column d of the 128-column tile; a half-stage = 4 octets of 8 consecutive k (32 coalesced scalar loads: a warp
        // reads 128 contiguous bytes of one source row per load).  Three rotating register buffers keep two half-stages in flight
        // while a third is split and stored; the load stream runs exactly one stage ahead of the store stream.
        const int tcol = threadIdx.x - 64;
        auto advance = [&](Walk w) {
            if (w.item >= p.n_items) return w;
            if (++w.kc == p.kchunks) {
                w.kc = 0;
                if (++w.t == p.plan.nterms[item_o(w.item)]) { w.t = 0; w.item += gridDim.x; }
            }
            return w;
        };
        float buf[3][32];
        auto load_half = [&](float (&r)[32], const Walk& w, int hf) {
            const bool valid = w.item < p.n_items;
            int o = 0, d = 0;
            const float* src = nullptr;
            if (valid) {
                o = item_o(w.item);
                d = ((w.item / p.row_tiles) % p.n_col_tiles) * TN_ + tcol;
                src = p.f[p.plan.src[o][w.t]] + d;
            }
            const int k0 = w.kc * SK + 32 * hf;
            const bool on = valid && d < p.D;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                r[j] = 0.f;
                if (on && k0 + j < p.h) r[j] = ldg_stream1(src + (size_t)(k0 + j) * p.ldf);
            }
        };
        auto store_half = [&](const float (&r)[32], int s, int hf) {
            uint8_t* row = smem_gen + s * STAGE_BYTES + 2 * PLANE + tcol * 128;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split1(r[8 * a + e], hi[e], lo[e]);
                const int chunk = (4 * hf + a) ^ (tcol & 7);
                *reinterpret_cast<uint4*>(row + 16 * chunk) =
                    make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
                *reinterpret_cast<uint4*>(row + PLANE + 16 * chunk) =
                    make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
            }
        };
        Walk cur = {(int)blockIdx.x, 0, 0};
        Walk nxt = advance(cur);
        load_half(buf[0], cur, 0);
        load_half(buf[1], cur, 1);
        for (int c = 0; cur.item < p.n_items; c += 3) {
            const uint32_t par = ((uint32_t)(c / STAGES)) & 1u;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {                   // stage c + jj lives in smem slot jj (STAGES == 3)
                if (cur.item < p.n_items) {
                    load_half(buf[(2 * jj + 2) % 3], nxt, 0);
                    mbar_wait(empty_bar(jj), par ^ 1u);
                    store_half(buf[(2 * jj) % 3], jj, 0);
                    load_half(buf[(2 * jj + 3) % 3], nxt, 1);
                    store_half(buf[(2 * jj + 1) % 3], jj, 1);
                    fence_proxy_async_smem();
                    mbar_arrive(ready_bar(jj));
                    cur = nxt;
                    nxt = advance(nxt);
                }
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
                            *reinterpret_cast<float4*>(out + cc * 32 + j) =
                                make_float4(__uint_as_float(r[j]) * OUT_SCALE, __uint_as_float(r[j + 1]) * OUT_SCALE,
                                            __uint_as_float(r[j + 2]) * OUT_SCALE, __uint_as_float(r[j + 3]) * OUT_SCALE);
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

bool plan_apply_h_supported(const otgan_plan_t* plan, int h, int D, const float* P, const float* const* F, int ldf,
                            float* const* out, int ldo)
{
    if (h < 1 || h > 4096 || D < 32 || (D & 3) || (ldo & 3) || !aligned16(P)) return false;
    (void)F; (void)ldf;
    for (int o = 0; o < plan->n_out; ++o)
        if (!aligned16(out[o])) return false;
    return true;
}

size_t plan_apply_h_workspace_bytes(int n_out, int h)
{
    const size_t hp = (size_t)ceil_div(h < 1 ? 1 : h, TM_) * TM_;
    return (size_t)2 * n_out * OTGAN_MAX_TERMS * hp * hp * sizeof(__half) + 256;
}

int plan_apply_h_launch(const otgan_plan_t* plan, int h, int D, const float* P, const float* const* F, int ldf,
                        float* const* out, int ldo, void* ws, size_t ws_bytes, cudaStream_t stream, int row_lo, int row_hi)
{
    OTGAN_REQUIRE(ws && ws_bytes >= plan_apply_h_workspace_bytes(plan->n_out, h), "plan_apply(tcgen05 f16): workspace too small");
    const int hp = ceil_div(h, TM_) * TM_;
    const size_t plane = (size_t)plan->n_out * OTGAN_MAX_TERMS * hp * hp;
    __half* a1 = reinterpret_cast<__half*>(ws);
    __half* a2 = a1 + plane;
    plan_prep_h_kernel<<<dim3(hp * hp / 256, plan->n_out, OTGAN_MAX_TERMS), 256, 0, stream>>>(*plan, h, hp, P, a1, a2);
    OTGAN_CHECK_LAUNCH("plan_prep_h_kernel");

    Params p;
    memset(&p, 0, sizeof(p));
    p.plan = *plan;
    const int arows = plan->n_out * OTGAN_MAX_TERMS * hp;
    // fp16 [arows, hp] described as fp32 [arows, hp / 2]: a [128 x 32] fp32 box is the [128 x 64] fp16 chunk, byte for byte
    if (!make_tensor_map_2d(&p.map_a1, reinterpret_cast<const float*>(a1), arows, hp / 2, hp / 2, 128, SK / 2, CU_TENSOR_MAP_SWIZZLE_128B))
        return OTGAN_EUNSUPPORTED;
    if (!make_tensor_map_2d(&p.map_a2, reinterpret_cast<const float*>(a2), arows, hp / 2, hp / 2, 128, SK / 2, CU_TENSOR_MAP_SWIZZLE_128B))
        return OTGAN_EUNSUPPORTED;
    for (int o = 0; o < plan->n_out; ++o) p.out[o] = out[o];
    for (int o = 0; o < plan->n_out; ++o)
        for (int t = 0; t < plan->nterms[o]; ++t) p.f[plan->src[o][t]] = F[plan->src[o][t]];      // only the sources the plan names
    if (row_hi <= 0) { row_lo = 0; row_hi = h; }
    p.rt_lo = row_lo / TM_;
    p.h = h; p.hp = hp; p.row_tiles = ceil_div(row_hi, TM_) - p.rt_lo; p.D = D; p.ldf = ldf; p.ldo = ldo;
    p.n_col_tiles = ceil_div(D, TN_);
    p.n_items = plan->n_out * p.n_col_tiles * p.row_tiles;
    p.kchunks = ceil_div(h, SK);
    OTGAN_SET_MAX_SMEM((plan_apply_h_kernel), SMEM_BYTES);
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    plan_apply_h_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
    OTGAN_CHECK_LAUNCH("plan_apply_h_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
