// sinkhorn_fast.cu -- persistent Sinkhorn kernel in stabilised scaling form (utils/matching.py:46-57).
//
// Same iteration as the reference (T x { row log-normalise ; column log-normalise }, then row softmax + entropy), but the
// iterate is carried as   log_a = L0 - f_i - g_j   with potentials split into an ABSORBED part (f, g: log domain, shared
// memory) and a SCALING part (u_i, v_j):   exp(log_a)_ij = K_ij * u_i * v_j,   K_ij = exp(L0_ij - f_i - g_j).
//   fast half-step:  u_i = 1 / sum_j K_ij v_j        (a 128x128 mat-vec: FMAs only, no exp/log)
//   slow half-step:  absorb (f -= log u, g -= log v), r_i = max-subtracted LSE of (L0 - f - g), f += r, K rebuilt
// A half-step takes the fast path unless some u_i (v_j) leaves [2^-40, 2^40] (checked every step, CTA-uniform via a
// shared-memory flag read after the step's barrier); then the same half-step is redone on the slow path from the last good state.  The first row step is
// always slow.  Exact algebra, fp32 rounding of the same class as the log-domain form (entries of K that underflow at an
// absorption are < 2^-126 and can regain at most 2^80, i.e. stay < 2^-46 of a row sum).
// The absorbed potentials f, g are kept in DOUBLE precision: with lambda = 500 they reach ~350, where one fp32 ulp is 3e-5, and
// exp(L0 - f - g) is a difference of such numbers -- in fp32 that rounding went straight into P (measured 2.9e-5 on a peaked
// h = 32 problem, 8x the log-domain kernel).  They are touched only on the slow path and in the epilogue (a few thousand fp64
// operations per block in total), the fast half-steps are unchanged.
//
// Layout: one CTA (256 threads) per block.  K is held twice in registers -- a row-oriented copy and a column-oriented
// copy, each as 4x16 tiles per thread -- so that a half-step is: 4 LDS.128 of the scaling vector, 64 FMAs, a 4-value
// reduce-scatter over the 8 lanes that share the tile rows (4 shuffles, 3 stages), one reciprocal, one barrier.  Shared-
// memory traffic per half-step is 16 KB (a 1x32 tiling needs 64 KB and is crossbar-bound; 8x8 needs a 4th shuffle stage).  L0 stays in
// shared memory, plain and transposed, for the slow path and the epilogue.
//
// Code size matters as much as arithmetic here: the slow path and the epilogue run once or twice per launch, COLD.  Written
// on the register tiles (fully unrolled: 64 double-precision exponentials per thread, three call sites) they were 12 000
// instructions = 190 KB of straight-line code, and the ncu source page showed half of the kernel's 140 us at T = 100 in
// instruction-fetch stalls of code executed once.  They are therefore ROLLED loops over rows (one warp per row, four rows in
// flight, results through the KX staging tile, then 32 LDS.128 into the two register copies), one call site each: ~10x less code.
#include "common.cuh"
#include <math.h>

namespace otgan {

namespace {

constexpr int H = 128, NTHREADS = 256, LDS_ = 132;   // 132-word rows: float4 aligned, conflict-free for both copies
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float S_LO = 9.094947017729282e-13f /* 2^-40 */, S_HI = 1099511627776.f /* 2^40 */;

struct Smem {
    float L0[H * LDS_];        // L0[r][c] = -lambda*C (natural-log units, exactly the caller's fp32), -inf outside [rows, cols)
    float L0T[H * LDS_];       // L0T[c][r]
    float KX[H * LDS_];        // slow path: the freshly normalised copy of K, to transpose it into the other copy without new exps
    double f[H], g[H];         // absorbed potentials (natural-log units), double precision: see the header comment
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
        sm.f[tid] = 0.0; sm.g[tid] = 0.0;
        sm.u[0][tid] = 1.f; sm.u[1][tid] = 1.f; sm.v[0][tid] = 1.f; sm.v[1][tid] = 1.f;
    }
    if (tid == 0) sm.flag = -1;
    __syncthreads();

    float Kr[4][16];    // row copy:    Kr[i][4m + e] = K[base4 + i][4 (g8 + 8m) + e]
    float Kc[4][16];    // column copy: Kc[j][4m + e] = K[4 (g8 + 8m) + e][base4 + j]
    int ub = 0, vb = 0, n_slow = 0;

    // exp of a DOUBLE exponent to fp32 accuracy (~3e-7 relative) whatever its magnitude.  A plain fp32 exp(float(x)) carries half
    // an ulp of |x| (4e-6 at |x| ~ 64) -- and entries that are negligible when K is rebuilt become the significant ones after the
    // scalings have moved by e^+-55, so that rounding went straight into P (measured 1.2e-5 on a peaked h = 8 block).
    // t = exponent in LOG2 units, double.  2^t to fp32 accuracy: t = n + r, |r| <= 1/2, 2^r by ex2.approx, 2^n by the exponent field.
    auto exp2_d = [](double t) -> float {
        if (!(t > -125.5)) return 0.f;                             // below the fp32 normal range (flush to zero), also -inf
        if (t > 127.0) return INFINITY;
        const int n = __double2int_rn(t);
        const float r = ex2_approx((float)(t - (double)n));
        return r * __int_as_float((n + 127) << 23);
    };
    constexpr double LOG2E_D = 1.4426950408889634074;

    // ---- slow path, rolled: one warp per LINE of M (a row of L0, or a row of L0T = a column), lane = 4 consecutive entries, four
    // lines in flight.  line value = (M[line][o] - p_line) - q_o; the line maximum (fp32, only a shift that keeps the exponentials in
    // range), K = exp(. - max) with the difference of the large terms taken in double (log2 units: one cvt, one DFMA, one DADD per
    // element, then exp2_d), the line sum, KX[line][:] = K / sum, p_line += max + log(sum).
    auto rebuild = [&](const float* __restrict__ M, double* pl, const double* qo, int nvalid) {
        double qd[4];
        float qf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const double q = qo[4 * lane + e]; qd[e] = q * LOG2E_D; qf[e] = (float)q; }
        float keep_t = 1.f, keep_mx = 0.f;
#pragma unroll 1
        for (int rb = 0; rb < 4; ++rb) {
            float4 l[4];
            float mx[4], t[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = warp + 8 * (4 * rb + a);
                l[a] = *reinterpret_cast<const float4*>(&M[r * LDS_ + 4 * lane]);
                const float pf = (float)pl[r];
                mx[a] = fmaxf(fmaxf((l[a].x - pf) - qf[0], (l[a].y - pf) - qf[1]), fmaxf((l[a].z - pf) - qf[2], (l[a].w - pf) - qf[3]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < 4; ++a) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            float4 k[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = warp + 8 * (4 * rb + a);
                if (!(mx[a] > -INFINITY)) mx[a] = 0.f;            // fully masked line
                const double pe = -(pl[r] + (double)mx[a]) * LOG2E_D;
                k[a].x = exp2_d(fma((double)l[a].x, LOG2E_D, pe) - qd[0]); k[a].y = exp2_d(fma((double)l[a].y, LOG2E_D, pe) - qd[1]);
                k[a].z = exp2_d(fma((double)l[a].z, LOG2E_D, pe) - qd[2]); k[a].w = exp2_d(fma((double)l[a].w, LOG2E_D, pe) - qd[3]);
                t[a] = (k[a].x + k[a].y) + (k[a].z + k[a].w);
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
                if (lane == 4 * rb + a) { keep_t = t[a]; keep_mx = mx[a]; }
            }
        }
        __syncwarp();
        if (lane < 16) {                                          // lane k publishes the potential of line warp + 8k
            const int r = warp + 8 * lane;
            if (r < nvalid) pl[r] += (double)keep_mx + log((double)keep_t);
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
    // absorb: f -= log u, g -= log v; afterwards u = v = 1 in the current buffers.  Then the half-step itself in the log domain:
    // row:    f_i += LSE_j(L0 - f - g), K rebuilt row-normalised;   column: g_j += LSE_i(L0 - f - g), K rebuilt column-normalised
    auto slow_step = [&](bool row_step) {
        if (tid < H) {
            sm.f[tid] -= log((double)sm.u[ub][tid]);
            sm.g[tid] -= log((double)sm.v[vb][tid]);
            sm.u[ub][tid] = 1.f;
            sm.v[vb][tid] = 1.f;
        }
        __syncthreads();
        rebuild(row_step ? sm.L0 : sm.L0T, row_step ? sm.f : sm.g, row_step ? sm.g : sm.f, row_step ? rows : cols);   // one inlined copy
        if (row_step) load_tiles(Kr, Kc); else load_tiles(Kc, Kr);
        ++n_slow;
    };
    // fast half-step: total_line = sum_o K[line][o] * x_o  (the caller takes the reciprocal)
    auto matvec = [&](const float (&K)[4][16], const float* x) -> float {
        float4 xv[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) xv[m] = *reinterpret_cast<const float4*>(&x[4 * (g8 + 8 * m)]);
        float s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t0 = K[i][0] * xv[0].x, t1 = K[i][4] * xv[1].x, t2 = K[i][8] * xv[2].x, t3 = K[i][12] * xv[3].x;
            t0 = fmaf(K[i][1], xv[0].y, t0); t1 = fmaf(K[i][5], xv[1].y, t1); t2 = fmaf(K[i][9], xv[2].y, t2); t3 = fmaf(K[i][13], xv[3].y, t3);
            t0 = fmaf(K[i][2], xv[0].z, t0); t1 = fmaf(K[i][6], xv[1].z, t1); t2 = fmaf(K[i][10], xv[2].z, t2); t3 = fmaf(K[i][14], xv[3].z, t3);
            t0 = fmaf(K[i][3], xv[0].w, t0); t1 = fmaf(K[i][7], xv[1].w, t1); t2 = fmaf(K[i][11], xv[2].w, t2); t3 = fmaf(K[i][15], xv[3].w, t3);
            s[i] = (t0 + t1) + (t2 + t3);
        }
        return oct_reduce_scatter4(s, lane);
    };

    // 2T half-steps, ONE loop body (a single inlined copy of the slow path): hs even = row step (utils/matching.py:53), odd = column
    // step (:54).  The first row step is always slow (K does not exist yet).
#pragma unroll 1
    for (int hs = 0; hs < 2 * T; ++hs) {
        const bool row_step = (hs & 1) == 0;
        bool slow = hs == 0;
        if (!slow) {
            float s;
            if (row_step) s = matvec(Kr, sm.v[vb]); else s = matvec(Kc, sm.u[ub]);
            const bool ok_line = my_line < (row_step ? rows : cols);
            const bool bad = ok_line && !(s >= S_LO && s <= S_HI);
            if ((lane & 1) == 0) {
                float* dst = row_step ? sm.u[ub ^ 1] : sm.v[vb ^ 1];
                dst[my_line] = ok_line ? __fdividef(1.f, s) : 1.f;
            }
            // CTA-uniform "some scaling left the range": the offending threads store the half-step index (same value from all of
            // them), everybody reads it back after the barrier that publishes the new scalings anyway -- a plain BAR.SYNC + one LDS
            // that travels with the next half-step's operand loads, instead of BAR.RED + B2R on the critical path
            if (bad) sm.flag = hs;
            __syncthreads();
            slow = sm.flag == hs;
            if (!slow) { if (row_step) ub ^= 1; else vb ^= 1; }
        }
        if (slow) slow_step(row_step);
    }

    // ================= P = softmax(log_a, -1), entropy, <P,C>                            utils/matching.py:56-57
    // log_a_ij = L0_ij - f_i - (g_j - log v_j); the row potential (and u) cancels in the row softmax.  Rolled like the slow path:
    // one warp per row, lane = 4 consecutive columns (a row of P leaves as one coalesced 512-byte store).
    if (tid < H) sm.g[tid] -= log((double)sm.v[vb][tid]);
    __syncthreads();
    float ent = 0.f, pcs = 0.f;
    {
        double qd[4], qn[4];
        float qf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { qn[e] = sm.g[4 * lane + e]; qd[e] = qn[e] * LOG2E_D; qf[e] = (float)qn[e]; }
        const int c = 4 * lane;
#pragma unroll 1
        for (int rb = 0; rb < 8; ++rb) {
            float4 l[2];
            float mx[2], s[2], ev[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int r = warp + 8 * (2 * rb + a);
                l[a] = *reinterpret_cast<const float4*>(&sm.L0[r * LDS_ + c]);
                const float pf = (float)sm.f[r];
                mx[a] = fmaxf(fmaxf((l[a].x - pf) - qf[0], (l[a].y - pf) - qf[1]), fmaxf((l[a].z - pf) - qf[2], (l[a].w - pf) - qf[3]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < 2; ++a) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int r = warp + 8 * (2 * rb + a);
                if (!(mx[a] > -INFINITY)) mx[a] = 0.f;
                const double pe = -(sm.f[r] + (double)mx[a]) * LOG2E_D;
                ev[a][0] = exp2_d(fma((double)l[a].x, LOG2E_D, pe) - qd[0]); ev[a][1] = exp2_d(fma((double)l[a].y, LOG2E_D, pe) - qd[1]);
                ev[a][2] = exp2_d(fma((double)l[a].z, LOG2E_D, pe) - qd[2]); ev[a][3] = exp2_d(fma((double)l[a].w, LOG2E_D, pe) - qd[3]);
                s[a] = (ev[a][0] + ev[a][1]) + (ev[a][2] + ev[a][3]);     // e = exp(log_a - row max), fp32-accurate for every magnitude
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < 2; ++a) s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int r = warp + 8 * (2 * rb + a);
                const float ls = LN2 * lg2_approx(s[a]);
                const double pr = sm.f[r] + (double)mx[a];
                const float l0v[4] = {l[a].x, l[a].y, l[a].z, l[a].w};
                float p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = (r < rows) && (c + k < cols);
                    p[k] = ok ? __fdiv_rn(ev[a][k], s[a]) : 0.f;
                    if (ok && p[k] > 0.f) {
                        const float am = (float)(((double)l0v[k] - pr) - qn[k]);          // log_a - row max
                        ent -= p[k] * (am - ls);
                        pcs += p[k] * l0v[k];
                    }
                }
                if (P && r < rows && c < cols) {
                    float* dst = P + boff + (size_t)r * cols + c;
                    if (vec) *reinterpret_cast<float4*>(dst) = make_float4(p[0], p[1], p[2], p[3]);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (c + k < cols) dst[k] = p[k];
                    }
                }
            }
        }
    }
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
