// sinkhorn_fast.cu -- persistent Sinkhorn kernel in stabilised scaling form (utils/matching.py:46-57).
//
// Same iteration as the reference (T x { row log-normalise ; column log-normalise }, then row softmax + entropy), but the
// iterate is carried as   log_a = L0 - f_i - g_j   with potentials split into an ABSORBED part (f, g: log domain, shared
// memory) and a SCALING part (u_i, v_j):   exp(log_a)_ij = K_ij * u_i * v_j,   K_ij = exp(L0_ij - f_i - g_j).
//   fast half-step:  u_i = 1 / sum_j K_ij v_j        (a 128x128 mat-vec: FMAs only, no exp/log)
//   slow half-step:  absorb (f -= log u, g -= log v), r_i = max-subtracted LSE of (L0 - f - g), f += r, K rebuilt
// A half-step takes the fast path unless some u_i (v_j) leaves [2^-51, 2^51] (checked every step, CTA-uniform via a
// shared-memory flag read after the step's barrier); then the same half-step is redone on the slow path from the last good state.  The first row step is
// always slow.  Exact algebra, fp32 rounding of the same class as the log-domain form: K keeps its entries down to the smallest
// DENORMAL (2^-149; FMAs take denormal operands at full rate), so what a rebuild flushes or quantises is < 2^-149 absolutely and
// can regain at most 2^102 before the next rebuild, i.e. stays < 2^-47 of a row sum.  (Round 2a flushed at 2^-126 and
// allowed [2^-40, 2^40]: the wider window saves about one 6 us rebuild in four.)
// Precision of the absorbed potentials.  With lambda = 500 they reach ~350, where one fp32 ulp is 3e-5, and exp(L0 - f - g) is
// a difference of such numbers: in plain fp32 that rounding went straight into P (measured 2.9e-5 on a peaked h = 32 problem, 8x
// the log-domain kernel).  Round 2 first took the difference in double -- accurate (P error 1.3e-7) but the fp64 pipe of this
// part made ONE rebuild of K cost 12 us (measured: T = 1 took 37 us, every slow step +12 us, 132 us at T = 100 of which only 65
// are fast half-steps).  Now the exponent is formed in fp32 with error-free transformations and no fp64 instruction at all:
// potentials are double-float pairs (hi, lo) in LOG2 units; L0 * log2(e) is an exact product (FMA residual + the low word of
// the constant); t = TL - f - g is accumulated with two TwoSums (hi words, exact) plus the low words; 2^t = 2^n * ex2(t - n)
// with n = rint(t_hi) taken BEFORE the low words are added, so the fraction carries the full fp32 precision whatever the
// magnitude of the terms; logarithms of the scalings are exponent + lg2(mantissa) (absolute error 2^-22).  ~35 fp32 operations
// per element instead of ~10 fp64 ones.
//
// Layout: one CTA (256 threads) per block.  K is held twice in registers -- a row-oriented copy and a column-oriented
// copy, each as 4x16 tiles per thread -- so that a half-step is: 4 LDS.128 of the scaling vector, 64 FMAs, a 4-value
// reduce-scatter over the 8 lanes that share the tile rows (4 shuffles, 3 stages), one reciprocal, one barrier.  Shared-
// memory traffic per half-step is 16 KB (a 1x32 tiling needs 64 KB and is crossbar-bound; 8x8 needs a 4th shuffle stage).  L0 stays in
// shared memory, plain and transposed, for the slow path and the epilogue.
//
// The slow path and the epilogue are ROLLED loops over rows (one warp per row, four rows in flight, results through the KX staging
// tile, then 32 LDS.128 into the two register copies), one call site each.  Written on the register tiles (fully unrolled, three
// call sites) they were 12 000 SASS instructions of straight-line code executed once or twice per launch; rolling them did not
// change the kernel time by itself (the fp64 arithmetic was the cost, see above) but is what made the error-free fp32 form -- ~35
// operations per element -- affordable in code size (4 800 instructions in total).  The epilogue needs no exponential at all:
// P_ij = u'_i K_ij v_j with u' from one more row half-step on the register copy of K.
// Phase clocks: tools/sinkhorn_phases.cu compiles this file with OTGAN_SINKHORN_CLOCKS.
#include "common.cuh"
#include <math.h>
#include <stddef.h>

namespace otgan {

namespace {

// Phase clocks for tools/sinkhorn_phases.cu (compiled in only there)
#ifdef OTGAN_SINKHORN_CLOCKS
__device__ long long g_sinkhorn_clk[OTGAN_MAX_BLOCKS][8];
#define SK_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x < OTGAN_MAX_BLOCKS) g_sinkhorn_clk[blockIdx.x][i] = clock64(); } while (0)
#define SK_CLK_ADD(i, t0) do { if (threadIdx.x == 0 && blockIdx.x < OTGAN_MAX_BLOCKS) g_sinkhorn_clk[blockIdx.x][i] += clock64() - (t0); } while (0)
#else
#define SK_CLK(i) do { } while (0)
#define SK_CLK_ADD(i, t0) do { } while (0)
#endif

constexpr int H = 128, NTHREADS = 256, LDS_ = 132;   // 132-word rows: float4 aligned, conflict-free for both copies
constexpr float LN2 = 0.6931471805599453f;
constexpr float CH = 1.44269502162933349609375f, CL = 1.925963033500011e-08f;   // log2(e) = CH + CL (fp32 pair)
constexpr float MAGIC = 12582912.f;                 // 1.5 * 2^23: (x + MAGIC) - MAGIC = rint(x) for |x| < 2^22
constexpr int MAGIC_BITS = 0x4B400000;
constexpr float S_LO = 4.440892098500626e-16f /* 2^-51 */, S_HI = 2251799813685248.f /* 2^51 */;

struct Smem {
    float L0[H * LDS_];        // L0[r][c] = -lambda*C (natural-log units, exactly the caller's fp32), -inf outside [rows, cols)
    float L0T[H * LDS_];       // L0T[c][r]
    float KX[H * LDS_];        // slow path: the freshly normalised copy of K, to transpose it into the other copy without new exps
    float fh[H], fl[H], gh[H], gl[H];   // absorbed potentials in LOG2 units as double-float pairs (hi, lo): see the header comment
    float u[2][H], v[2][H];    // scaling vectors, double buffered
    float red[2][NTHREADS / 32];
    int flag;                  // index of the last half-step whose scalings left [S_LO, S_HI] (see the main loop)
};

// s[4]: per-lane partials for the 4 tile lines; returns the total over the 8 lanes that share them for tile line
// 2*b2 + b1 of the lane index (both lanes of a b0 pair hold it).  Fixed order -> deterministic.
__device__ __forceinline__ float oct_reduce_scatter4(const float (&s)[4], int lane)
{
    const bool b2 = lane & 4, b1 = lane & 2;
    float k[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float snd = b2 ? s[i] : s[2 + i];
        k[i] = (b2 ? s[2 + i] : s[i]) + __shfl_xor_sync(0xffffffffu, snd, 4);
    }
    const float snd = b1 ? k[0] : k[1];
    float t = (b1 ? k[1] : k[0]) + __shfl_xor_sync(0xffffffffu, snd, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// ---- shared-memory accesses of the fast half-step through 32-bit shared-window addresses computed ONCE: with generic pointers
// the compiler re-derived the window base (S2UR SR_CgaCtaId -> ULEA -> LEA) at the top of every half-step, i.e. on the critical
// path right behind the barrier
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds32i(uint32_t a) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// 1/x for x in [2^-40, 2^40] (the range check of the fast half-step guarantees it): one MUFU, no denormal guard
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32i(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- error-free fp32 building blocks (intrinsics: never contracted or re-associated)
// (h, l) += x in double-float arithmetic
__device__ __forceinline__ void dd_add(float& h, float& l, float x)
{
    const float s = __fadd_rn(h, x), bb = __fsub_rn(s, h);
    const float e = __fadd_rn(__fsub_rn(h, __fsub_rn(s, bb)), __fsub_rn(x, bb));
    const float l2 = __fadd_rn(l, e), h2 = __fadd_rn(s, l2);
    l = __fsub_rn(l2, __fsub_rn(h2, s));
    h = h2;
}
// log2 of a positive normal float as exponent (exact) + lg2(mantissa in [1, 2)) (absolute error ~2^-22)
__device__ __forceinline__ void log2_split(float x, float& ipart, float& fpart)
{
    const int b = __float_as_int(x);
    ipart = (float)((b >> 23) - 127);
    fpart = lg2_approx(__int_as_float((b & 0x007FFFFF) | 0x3F800000));
}
// exponent t = l * log2(e) - (ph + pl) - (qh + ql) of one matrix entry as (t_hi, t_lo): t_hi is the EXACT fp32 sum chain's head,
// t_lo collects every rounding error and the low words (|t_lo| << 1).  l must be finite.
__device__ __forceinline__ void exponent_pair(float l, float ph, float pl, float qh, float ql, float& t_hi, float& t_lo, bool zero_potentials = false)
{
    const float th = __fmul_rn(l, CH);
    const float tl = __fmaf_rn(l, CL, __fmaf_rn(l, CH, -th));
    if (zero_potentials) {                                         // the very first rebuild (f = g = 0): nothing to subtract
        t_hi = th;
        t_lo = tl;
        return;
    }
    const float s1 = __fsub_rn(th, ph), b1 = __fsub_rn(s1, th);
    const float e1 = __fsub_rn(__fsub_rn(th, __fsub_rn(s1, b1)), __fadd_rn(ph, b1));
    const float s2 = __fsub_rn(s1, qh), b2 = __fsub_rn(s2, s1);
    const float e2 = __fsub_rn(__fsub_rn(s1, __fsub_rn(s2, b2)), __fadd_rn(qh, b2));
    t_hi = s2;
    t_lo = __fadd_rn(__fsub_rn(__fsub_rn(tl, pl), ql), __fadd_rn(e1, e2));
}
// 2^(t_hi + t_lo - shift) for an integer shift >= rint(t_hi): n = rint(t_hi), fraction = (t_hi - n) + t_lo in [-0.5, 0.5] exactly
// representable, ex2 of the fraction, the integer part through the exponent field.  Results below 2^-126 come out as DENORMALS
// (second factor 2^-64 through a multiply, which rounds correctly) and only values below 2^-149 flush to zero: the 23 extra binades
// are what lets the scalings range over [2^-51, 2^51] between two rebuilds (see the header).
// ex_out = n - shift (for the entropy term), fr_out = the fraction.
__device__ __forceinline__ float exp2_pair(float t_hi, float t_lo, int shift, int& ex_out, float& fr_out)
{
    const float A = __fadd_rn(t_hi, MAGIC);
    const float nf = __fsub_rn(A, MAGIC);
    const int ex = (__float_as_int(A) - MAGIC_BITS) - shift;
    const float fr = __fadd_rn(__fsub_rn(t_hi, nf), t_lo);
    const float k = ex2_approx(fr);
    ex_out = ex;
    fr_out = fr;
    if (ex < -150) return 0.f;
    const bool low = ex < -100;
    const float r = __int_as_float(__float_as_int(k) + ((low ? ex + 64 : ex) << 23));
    return low ? __fmul_rn(r, 5.421010862427522e-20f /* 2^-64 */) : r;
}

__global__ void __launch_bounds__(NTHREADS, 1)
sinkhorn_fast_kernel(const float* __restrict__ L0g, float* __restrict__ P, float* __restrict__ entropy,
                     float* __restrict__ pc, int* __restrict__ slow_steps, int rows, int cols, int T, float lam)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lg = lane >> 3, g8 = lane & 7;
    const int base4 = 16 * warp + 4 * lg;        // first of the 4 tile rows (row copy) / tile columns (column copy)
    // the 4 x 4 columns (row copy) / rows (column copy) of this lane: float4 groups g8 + 8m, m = 0..3
    const int my_idx = (lane >> 1) & 3;          // tile line whose total this lane holds after the reduce-scatter
    const int my_line = base4 + my_idx;
    const size_t boff = (size_t)blockIdx.x * rows * cols;
    const float* __restrict__ Lb = L0g + boff;
    const bool vec = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(Lb) & 15u) == 0);
    SK_CLK(0);
#ifdef OTGAN_SINKHORN_CLOCKS
    if (tid == 0 && blockIdx.x < OTGAN_MAX_BLOCKS) { g_sinkhorn_clk[blockIdx.x][5] = 0; g_sinkhorn_clk[blockIdx.x][6] = 0; }
#endif

    // ---- stage L0 into shared memory, plain and transposed; lanes = 32 consecutive rows -> conflict-free.
    // Vector path: all 16 loads of a thread are issued before the first store (one L2 round trip instead of sixteen); ragged
    // or unaligned blocks take a rolled scalar loop (code size, see the header).
    if (vec) {
        float4 x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = tid + j * NTHREADS;
            const int r = (t >> 10) * 32 + (t & 31), c4 = ((t >> 5) & 31) * 4;
            x[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (r < rows && c4 < cols) x[j] = __ldg(reinterpret_cast<const float4*>(Lb + (size_t)r * cols + c4));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = tid + j * NTHREADS;
            const int r = (t >> 10) * 32 + (t & 31), c4 = ((t >> 5) & 31) * 4;
            const float4 y = x[j];
            *reinterpret_cast<float4*>(&sm.L0[r * LDS_ + c4]) = y;
            sm.L0T[(c4 + 0) * LDS_ + r] = y.x; sm.L0T[(c4 + 1) * LDS_ + r] = y.y;
            sm.L0T[(c4 + 2) * LDS_ + r] = y.z; sm.L0T[(c4 + 3) * LDS_ + r] = y.w;
        }
    } else {
#pragma unroll 1
        for (int t = tid; t < H * H; t += NTHREADS) {
            const int r = (t >> 12) * 32 + (t & 31), c = (t >> 5) & 127;
            const float y = (r < rows && c < cols) ? __ldg(Lb + (size_t)r * cols + c) : -INFINITY;
            sm.L0[r * LDS_ + c] = y;
            sm.L0T[c * LDS_ + r] = y;
        }
    }
    if (tid < H) {
        sm.fh[tid] = 0.f; sm.fl[tid] = 0.f; sm.gh[tid] = 0.f; sm.gl[tid] = 0.f;
        sm.u[0][tid] = 1.f; sm.u[1][tid] = 1.f; sm.v[0][tid] = 1.f; sm.v[1][tid] = 1.f;
    }
    if (tid == 0) sm.flag = -1;
    __syncthreads();
    SK_CLK(1);

    float Kr[4][16];    // row copy:    Kr[i][4m + e] = K[base4 + i][4 (g8 + 8m) + e]
    float Kc[4][16];    // column copy: Kc[j][4m + e] = K[4 (g8 + 8m) + e][base4 + j]
    int ub = 0, vb = 0, n_slow = 0;

    // ---- slow path, rolled: one warp per LINE of M (a row of L0, or a row of L0T = a column), lane = 4 consecutive entries, four
    // lines in flight.  Entry exponent t = M[line][o] * log2(e) - p_line - q_o as an fp32 pair (exponent_pair); the line maximum of
    // the heads (rounded to an integer: the shift), K = 2^(t - shift), the line sum, KX[line][:] = K / sum, p_line += shift + log2(sum).
    auto rebuild = [&](const float* __restrict__ M, float* ph, float* pl, const float* qh, const float* ql, int nvalid, bool first) {
        float qhv[4], qlv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { qhv[e] = qh[4 * lane + e]; qlv[e] = ql[4 * lane + e]; }
        float keep_t = 1.f, keep_shift = 0.f;
#pragma unroll 1
        for (int rb = 0; rb < 4; ++rb) {
            float thi[4][4], tlo[4][4], mx[4], t[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = warp + 8 * (4 * rb + a);
                const float4 l4 = *reinterpret_cast<const float4*>(&M[r * LDS_ + 4 * lane]);
                const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
                const float p_h = ph[r], p_l = pl[r];
                mx[a] = -INFINITY;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool live = lv[e] > -3.0e38f;                   // -inf marks entries outside [rows, cols)
                    exponent_pair(live ? lv[e] : 0.f, p_h, p_l, qhv[e], qlv[e], thi[a][e], tlo[a][e], first);
                    if (!(live && thi[a][e] > -4.0e6f)) thi[a][e] = -INFINITY;   // also: beyond the range of the integer split (K = 0 anyway)
                    mx[a] = fmaxf(mx[a], thi[a][e]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < 4; ++a) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            float4 k[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!(mx[a] > -INFINITY)) mx[a] = 0.f;            // fully masked line
                mx[a] = rintf(mx[a]);
                const int shift = (int)mx[a];
                float kv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int ex; float fr;
                    const float kk = exp2_pair(thi[a][e], tlo[a][e], shift, ex, fr);
                    kv[e] = (thi[a][e] > -INFINITY) ? kk : 0.f;
                }
                k[a] = make_float4(kv[0], kv[1], kv[2], kv[3]);
                t[a] = (kv[0] + kv[1]) + (kv[2] + kv[3]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < 4; ++a) t[a] += __shfl_xor_sync(0xffffffffu, t[a], o);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = warp + 8 * (4 * rb + a);
                const float inv = (r < nvalid) ? __frcp_rn(t[a]) : 0.f;
                *reinterpret_cast<float4*>(&sm.KX[r * LDS_ + 4 * lane]) = make_float4(k[a].x * inv, k[a].y * inv, k[a].z * inv, k[a].w * inv);
                if (lane == 4 * rb + a) { keep_t = t[a]; keep_shift = mx[a]; }
            }
        }
        __syncwarp();
        if (lane < 16) {                                          // lane k publishes the potential of line warp + 8k
            const int r = warp + 8 * lane;
            if (r < nvalid) {
                float ip, fp, h = ph[r], l = pl[r];
                log2_split(keep_t, ip, fp);
                dd_add(h, l, keep_shift + ip);                    // integers: exact
                dd_add(h, l, fp);
                ph[r] = h; pl[r] = l;
            }
        }
        __syncthreads();                                          // KX complete, potentials published
    };
    // register copies from the staging tile: `same` = the orientation KX was written in, `other` = its transpose
    auto load_tiles = [&](float (&same)[4][16], float (&other)[4][16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float4 v = *reinterpret_cast<const float4*>(&sm.KX[(base4 + i) * LDS_ + 4 * (g8 + 8 * m)]);
                same[i][4 * m] = v.x; same[i][4 * m + 1] = v.y; same[i][4 * m + 2] = v.z; same[i][4 * m + 3] = v.w;
            }
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 v = *reinterpret_cast<const float4*>(&sm.KX[(4 * (g8 + 8 * m) + e) * LDS_ + base4]);
                other[0][4 * m + e] = v.x; other[1][4 * m + e] = v.y; other[2][4 * m + e] = v.z; other[3][4 * m + e] = v.w;
            }
    };
    // absorb: f -= log2 u, g -= log2 v; afterwards u = v = 1 in the current buffers.  Then the half-step itself in the log domain:
    // row:    f_i += LSE_j(L0 - f - g), K rebuilt row-normalised;   column: g_j += LSE_i(L0 - f - g), K rebuilt column-normalised
    auto slow_step = [&](bool row_step, bool first) {
        if (tid < H) {
            float ip, fp, h = sm.fh[tid], l = sm.fl[tid];
            log2_split(sm.u[ub][tid], ip, fp);
            dd_add(h, l, -ip); dd_add(h, l, -fp);
            sm.fh[tid] = h; sm.fl[tid] = l;
            h = sm.gh[tid]; l = sm.gl[tid];
            log2_split(sm.v[vb][tid], ip, fp);
            dd_add(h, l, -ip); dd_add(h, l, -fp);
            sm.gh[tid] = h; sm.gl[tid] = l;
            sm.u[ub][tid] = 1.f;
            sm.v[vb][tid] = 1.f;
        }
        __syncthreads();
        if (row_step) rebuild(sm.L0, sm.fh, sm.fl, sm.gh, sm.gl, rows, first); else rebuild(sm.L0T, sm.gh, sm.gl, sm.fh, sm.fl, cols, false);
        if (row_step) load_tiles(Kr, Kc); else load_tiles(Kc, Kr);
        ++n_slow;
    };
    // fast half-step: total_line = sum_o K[line][o] * x_o  (the caller takes the reciprocal); xa = shared address of this lane's
    // first float4 of the scaling vector
    auto matvec4 = [&](const float (&K)[4][16], uint32_t xa, float (&s)[4]) {
        float4 xv[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) xv[m] = lds128(xa + 128u * m);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t0 = K[i][0] * xv[0].x, t1 = K[i][4] * xv[1].x, t2 = K[i][8] * xv[2].x, t3 = K[i][12] * xv[3].x;
            t0 = fmaf(K[i][1], xv[0].y, t0); t1 = fmaf(K[i][5], xv[1].y, t1); t2 = fmaf(K[i][9], xv[2].y, t2); t3 = fmaf(K[i][13], xv[3].y, t3);
            t0 = fmaf(K[i][2], xv[0].z, t0); t1 = fmaf(K[i][6], xv[1].z, t1); t2 = fmaf(K[i][10], xv[2].z, t2); t3 = fmaf(K[i][14], xv[3].z, t3);
            t0 = fmaf(K[i][3], xv[0].w, t0); t1 = fmaf(K[i][7], xv[1].w, t1); t2 = fmaf(K[i][11], xv[2].w, t2); t3 = fmaf(K[i][15], xv[3].w, t3);
            s[i] = (t0 + t1) + (t2 + t3);
        }
    };
    const uint32_t sm_base = static_cast<uint32_t>(__cvta_generic_to_shared(&sm));
    const uint32_t u_sa = sm_base + (uint32_t)offsetof(Smem, u), v_sa = sm_base + (uint32_t)offsetof(Smem, v);
    const uint32_t flag_sa = sm_base + (uint32_t)offsetof(Smem, flag);

    // 2T half-steps, ONE loop body (a single inlined copy of the slow path): hs even = row step (utils/matching.py:53), odd = column
    // step (:54).  The first row step is always slow (K does not exist yet); T = 0 runs it alone so that the epilogue finds K = the
    // row softmax of L0.
    const int nhs = T > 0 ? 2 * T : 1;
#pragma unroll 1
    for (int hs = 0; hs < nhs; ++hs) {
        const bool row_step = (hs & 1) == 0;
        bool slow = hs == 0;
        if (!slow) {
            float s4[4];
            if (row_step) matvec4(Kr, v_sa + 512u * vb + 16u * g8, s4); else matvec4(Kc, u_sa + 512u * ub + 16u * g8, s4);
            const float s = oct_reduce_scatter4(s4, lane);
            const bool ok_line = my_line < (row_step ? rows : cols);
            const bool bad = ok_line && !(s >= S_LO && s <= S_HI);
            if ((lane & 1) == 0)
                sts32((row_step ? u_sa + 512u * (ub ^ 1) : v_sa + 512u * (vb ^ 1)) + 4u * my_line, ok_line ? rcp_approx(s) : 1.f);
            // CTA-uniform "some scaling left the range": the offending threads store the half-step index (same value from all of
            // them), everybody reads it back after the barrier that publishes the new scalings anyway -- a plain BAR.SYNC + one LDS
            // instead of BAR.RED + B2R.  (Tried and dropped: loading the next half-step's operands in the same breath as the flag,
            // before it is examined -- 6 us slower at T = 100, the 16 extra live registers cost more than the flag's latency.)
            if (bad) sts32i(flag_sa, hs);
            __syncthreads();
            slow = lds32i(flag_sa) == hs;
            if (!slow) { if (row_step) ub ^= 1; else vb ^= 1; }
        }
        if (slow) {
#ifdef OTGAN_SINKHORN_CLOCKS
            const long long t0 = clock64();
#endif
            slow_step(row_step, hs == 0);
            SK_CLK_ADD(5, t0);
        }
        if (hs == 0) SK_CLK(2);
    }
    if (T == 0) n_slow = 0;
    SK_CLK(3);

    // ================= P = softmax(log_a, -1), entropy, <P,C>                            utils/matching.py:56-57
    // The row softmax of log_a = L0 - f - (g - log v) IS one more row half-step in scaling form: P_ij = u'_i K_ij v_j with
    // u'_i = 1 / sum_j K_ij v_j (the row potential and the old u cancel).  So P comes from the register copy of K: no exponential.
    // log P_ij = log2 K_ij + log2 v_j + log2 u'_i, each as exponent (exact) + lg2(mantissa) (absolute error 2^-22); the integer and
    // the fractional parts of the entropy sum are accumulated separately so that neither loses the other's low bits.
    if (tid < H) {                                                // log2 v_j as (integer, fraction) for the threads that hold column j
        float ip, fp;
        log2_split(sm.v[vb][tid], ip, fp);
        sm.gh[tid] = ip; sm.gl[tid] = fp;
    }
    __syncthreads();
    float ent = 0.f, pcs = 0.f;
    {
        float s4[4], up[4], lui[4], luf[4];
        matvec4(Kr, v_sa + 512u * vb + 16u * g8, s4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s4[i] += __shfl_xor_sync(0xffffffffu, s4[i], o);
            const bool okr = base4 + i < rows && s4[i] > 0.f;
            up[i] = okr ? __frcp_rn(s4[i]) : 0.f;
            lui[i] = 0.f; luf[i] = 0.f;
            if (okr) log2_split(up[i], lui[i], luf[i]);
        }
        float entI = 0.f, entF = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int c = 4 * (g8 + 8 * m);
            const float4 xv = lds128(v_sa + 512u * vb + 4u * c);
            const float4 vi = *reinterpret_cast<const float4*>(&sm.gh[c]);
            const float4 vf = *reinterpret_cast<const float4*>(&sm.gl[c]);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, vis[4] = {vi.x, vi.y, vi.z, vi.w}, vfs[4] = {vf.x, vf.y, vf.z, vf.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = base4 + i;
                const float4 l4 = *reinterpret_cast<const float4*>(&sm.L0[r * LDS_ + c]);
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
                float p[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float k = Kr[i][4 * m + e];
                    p[e] = (k * xs[e]) * up[i];
                    // branch-free: an entry with p = 0 (K = 0, masked row / column) adds exactly 0 -- every factor below is finite
                    const bool tiny = k < 1.0e-30f;                                                      // denormal K: normalise before taking its exponent
                    const int b = __float_as_int(tiny ? k * 18446744073709551616.f /* 2^64 */ : k);
                    const float ek = __int_as_float(MAGIC_BITS + ((b >> 23) - (tiny ? 191 : 127))) - MAGIC;       // exponent of K as a float, no I2F
                    const float mk = lg2_approx(__int_as_float((b & 0x007FFFFF) | 0x3F800000));
                    entI = fmaf(p[e], ek + (vis[e] + lui[i]), entI);
                    entF = fmaf(p[e], mk + (vfs[e] + luf[i]), entF);
                    pcs = fmaf(p[e], p[e] > 0.f ? ls[e] : 0.f, pcs);                                      // L0 is -inf outside [rows, cols)
                }
                if (P && r < rows && c < cols) {
                    float* dst = P + boff + (size_t)r * cols + c;
                    if (vec) *reinterpret_cast<float4*>(dst) = make_float4(p[0], p[1], p[2], p[3]);
                    else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (c + e < cols) dst[e] = p[e];
                    }
                }
            }
        }
        ent = -LN2 * (entI + entF);
    }
    SK_CLK(4);
    ent = warp_sum(ent);
    pcs = warp_sum(pcs);
    if (lane == 0) { sm.red[0][warp] = ent; sm.red[1][warp] = pcs; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < NTHREADS / 32; ++w) { a += sm.red[0][w]; b += sm.red[1][w]; }
        if (entropy) entropy[blockIdx.x] = a / (float)rows;
        if (pc) pc[blockIdx.x] = -b / lam;                  // C = -L0/lambda
        if (slow_steps) slow_steps[blockIdx.x] = n_slow;
    }
}

}  // namespace

int sinkhorn_fast_launch(int nblk, int rows, int cols, int T, float lam, const float* L0, float* P, float* entropy,
                         float* pc, int* slow_steps, cudaStream_t stream)
{
    OTGAN_SET_MAX_SMEM((sinkhorn_fast_kernel), sizeof(Smem));
    sinkhorn_fast_kernel<<<nblk, NTHREADS, sizeof(Smem), stream>>>(L0, P, entropy, pc, slow_steps, rows, cols, T, lam);
    OTGAN_CHECK_LAUNCH("sinkhorn_fast_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
