// sinkhorn_fast.cu -- persistent Sinkhorn kernel in stabilised scaling form (utils/matching.py:46-57).
//
// Same iteration as the reference (T x { row log-normalise ; column log-normalise }, then row softmax + entropy), but the
// iterate is carried as   log_a = L0 - f_i - g_j   with potentials split into an ABSORBED part (f, g: log domain, shared
// memory) and a SCALING part (u_i, v_j):   exp(log_a)_ij = K_ij * u_i * v_j,   K_ij = exp(L0_ij - f_i - g_j).
//   fast half-step:  u_i = 1 / sum_j K_ij v_j        (a 128x128 mat-vec: FMAs only, no exp/log)
//   slow half-step:  absorb (f -= log u, g -= log v), r_i = max-subtracted LSE of (L0 - f - g), f += r, K rebuilt
// A half-step takes the fast path unless some u_i (v_j) leaves [2^-40, 2^40] (checked every step, CTA-uniform via
// __syncthreads_or); then the same half-step is redone on the slow path from the last good state.  The first row step is
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
    // All 16 loads of a thread are issued before the first store (one L2 round trip instead of sixteen).
    {
        float4 x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = tid + j * NTHREADS;
            const int r = (t >> 10) * 32 + (t & 31), c4 = ((t >> 5) & 31) * 4;
            x[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (r < rows) {
                if (vec && c4 + 4 <= cols) {
                    x[j] = __ldg(reinterpret_cast<const float4*>(Lb + (size_t)r * cols + c4));
                } else {
                    if (c4 + 0 < cols) x[j].x = __ldg(Lb + (size_t)r * cols + c4 + 0);
                    if (c4 + 1 < cols) x[j].y = __ldg(Lb + (size_t)r * cols + c4 + 1);
                    if (c4 + 2 < cols) x[j].z = __ldg(Lb + (size_t)r * cols + c4 + 2);
                    if (c4 + 3 < cols) x[j].w = __ldg(Lb + (size_t)r * cols + c4 + 3);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = tid + j * NTHREADS;
            const int r = (t >> 10) * 32 + (t & 31), c4 = ((t >> 5) & 31) * 4;
            float4 y = x[j];
            *reinterpret_cast<float4*>(&sm.L0[r * LDS_ + c4]) = y;
            sm.L0T[(c4 + 0) * LDS_ + r] = y.x; sm.L0T[(c4 + 1) * LDS_ + r] = y.y;
            sm.L0T[(c4 + 2) * LDS_ + r] = y.z; sm.L0T[(c4 + 3) * LDS_ + r] = y.w;
        }
    }
    if (tid < H) {
        sm.f[tid] = 0.0; sm.g[tid] = 0.0;
        sm.u[0][tid] = 1.f; sm.u[1][tid] = 1.f; sm.v[0][tid] = 1.f; sm.v[1][tid] = 1.f;
    }
    __syncthreads();

    float Kr[4][16];    // row copy:    Kr[i][4m + e] = K[base4 + i][4 (g8 + 8m) + e]
    float Kc[4][16];    // column copy: Kc[j][4m + e] = K[4 (g8 + 8m) + e][base4 + j]
    int ub = 0, vb = 0, n_slow = 0;

    auto absorb = [&]() {   // f -= log2 u, g -= log2 v; afterwards u = v = 1 in the current buffers
        if (tid < H) {
            sm.f[tid] -= log((double)sm.u[ub][tid]);
            sm.g[tid] -= log((double)sm.v[vb][tid]);
            sm.u[ub][tid] = 1.f;
            sm.v[vb][tid] = 1.f;
        }
        __syncthreads();
    };
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
    // approximate (fp32) line maxima of (M[line][o] - p_line) - q_o: only a shift that keeps the exponentials in range
    auto line_max = [&](const float* M, const double* pl, const double* qo, float (&mx)[4]) {
        float pf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { mx[i] = -INFINITY; pf[i] = (float)pl[base4 + i]; }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int o = 4 * (g8 + 8 * m);
            const float q0 = (float)qo[o], q1 = (float)qo[o + 1], q2 = (float)qo[o + 2], q3 = (float)qo[o + 3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 l = *reinterpret_cast<const float4*>(&M[(base4 + i) * LDS_ + o]);
                mx[i] = fmaxf(mx[i], fmaxf(fmaxf((l.x - pf[i]) - q0, (l.y - pf[i]) - q1), fmaxf((l.z - pf[i]) - q2, (l.w - pf[i]) - q3)));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
            if (!(mx[i] > -INFINITY)) mx[i] = 0.f;                // fully masked line
        }
    };
    // K[i][e] = exp((M[line][o] - p_line) - q_o - shift_i): the difference of the large terms is taken in double, in log2 units
    // (one cvt, one DFMA, one DADD per element), then reduced and exponentiated (exp2_d)
    auto exp_tile = [&](const float* M, const double* pl, const double* qo, const float (&shift)[4], float (&K)[4][16]) {
        double pe[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pe[i] = -(pl[base4 + i] + (double)shift[i]) * LOG2E_D;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int o = 4 * (g8 + 8 * m);
            const double q0 = qo[o] * LOG2E_D, q1 = qo[o + 1] * LOG2E_D, q2 = qo[o + 2] * LOG2E_D, q3 = qo[o + 3] * LOG2E_D;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 l = *reinterpret_cast<const float4*>(&M[(base4 + i) * LDS_ + o]);
                K[i][4 * m + 0] = exp2_d(fma((double)l.x, LOG2E_D, pe[i]) - q0); K[i][4 * m + 1] = exp2_d(fma((double)l.y, LOG2E_D, pe[i]) - q1);
                K[i][4 * m + 2] = exp2_d(fma((double)l.z, LOG2E_D, pe[i]) - q2); K[i][4 * m + 3] = exp2_d(fma((double)l.w, LOG2E_D, pe[i]) - q3);
            }
        }
    };
    // max-subtracted log-sum-exp over each of the 4 tile lines of (M - p - q); on return K = exp(. - lse) (line-normalised) and
    // the lane holding line my_idx returns its lse (double).  nvalid = number of valid lines (rows or cols).
    auto normalise_lines = [&](const float* M, const double* pl, const double* qo, float (&K)[4][16], int nvalid) -> double {
        float mx[4];
        line_max(M, pl, qo, mx);
        exp_tile(M, pl, qo, mx, K);
        double lse = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) t += K[i][e];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            const float inv = (base4 + i < nvalid) ? __frcp_rn(t) : 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) K[i][e] *= inv;
            if (i == my_idx) lse = (double)mx[i] + log((double)t);
        }
        return lse;
    };
    // The other copy of K is the transpose of the one just normalised: through shared memory (16 STS.128 + 16 LDS.128 per
    // thread) instead of 64 more double-precision exponentials per thread.  src tile: src[i][4m+e] = X[base4+i][4(g8+8m)+e];
    // dst[j][4m+e] = X[4(g8+8m)+e][base4+j].
    auto transpose_into = [&](const float (&src)[4][16], float (&dst)[4][16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m)
                *reinterpret_cast<float4*>(&sm.KX[(base4 + i) * LDS_ + 4 * (g8 + 8 * m)]) =
                    make_float4(src[i][4 * m], src[i][4 * m + 1], src[i][4 * m + 2], src[i][4 * m + 3]);
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 v = *reinterpret_cast<const float4*>(&sm.KX[(4 * (g8 + 8 * m) + e) * LDS_ + base4]);
                dst[0][4 * m + e] = v.x; dst[1][4 * m + e] = v.y; dst[2][4 * m + e] = v.z; dst[3][4 * m + e] = v.w;
            }
    };
    auto slow_row = [&]() {   // f_i += LSE_j(L0 - f - g); K rebuilt; u = v = 1
        absorb();
        const double lse = normalise_lines(sm.L0, sm.f, sm.g, Kr, rows);
        if ((lane & 1) == 0 && my_line < rows) sm.f[my_line] += lse;
        transpose_into(Kr, Kc);                // includes the barrier that publishes f
        ++n_slow;
    };
    auto slow_col = [&]() {
        absorb();
        const double lse = normalise_lines(sm.L0T, sm.g, sm.f, Kc, cols);
        if ((lane & 1) == 0 && my_line < cols) sm.g[my_line] += lse;
        transpose_into(Kc, Kr);
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

    for (int it = 0; it < T; ++it) {
        // ================= row half-step: log_a -= reduce_logsumexp(log_a, axis=1)      utils/matching.py:53
        if (it == 0) {
            slow_row();
        } else {
            const float s = matvec(Kr, sm.v[vb]);
            const bool ok_line = my_line < rows;
            const bool bad = ok_line && !(s >= S_LO && s <= S_HI);
            if ((lane & 1) == 0) sm.u[ub ^ 1][my_line] = ok_line ? __fdividef(1.f, s) : 1.f;
            if (__syncthreads_or(bad)) slow_row(); else ub ^= 1;
        }
        // ================= column half-step: log_a -= reduce_logsumexp(log_a, axis=0)   utils/matching.py:54
        {
            const float s = matvec(Kc, sm.u[ub]);
            const bool ok_line = my_line < cols;
            const bool bad = ok_line && !(s >= S_LO && s <= S_HI);
            if ((lane & 1) == 0) sm.v[vb ^ 1][my_line] = ok_line ? __fdividef(1.f, s) : 1.f;
            if (__syncthreads_or(bad)) slow_col(); else vb ^= 1;
        }
    }

    // ================= P = softmax(log_a, -1), entropy, <P,C>                            utils/matching.py:56-57
    // log_a_ij = L0_ij - f_i - (g_j - log v_j); the row potential (and u) cancels in the row softmax.
    if (tid < H) sm.g[tid] -= log((double)sm.v[vb][tid]);
    __syncthreads();
    float ent = 0.f, pcs = 0.f;
    {
        float mx[4];
        line_max(sm.L0, sm.f, sm.g, mx);
        float e[4][16];
        exp_tile(sm.L0, sm.f, sm.g, mx, e);                    // e = exp(log_a - row max), fp32-accurate for every magnitude
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = base4 + i;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s += e[i][k];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float ls = LN2 * lg2_approx(s);
            const double pr = sm.f[r] + (double)mx[i];
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                const int c = 4 * (g8 + 8 * mm);
                float p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = (r < rows) && (c + k < cols);
                    p[k] = ok ? __fdiv_rn(e[i][4 * mm + k], s) : 0.f;
                    if (ok && p[k] > 0.f) {
                        const float l0 = sm.L0[r * LDS_ + c + k];
                        const float am = (float)(((double)l0 - pr) - sm.g[c + k]);       // log_a - row max
                        ent -= p[k] * (am - ls);
                        pcs += p[k] * l0;
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
