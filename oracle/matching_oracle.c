/* CPU oracle (C + OpenMP) for the OT-GAN matching hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the reference algorithm, independent of oracle/matching_oracle.py, used
 *   (1) to cross-check the numpy oracle (tests/test_oracle.py), and
 *   (2) as the multi-core CPU baseline timed by bench.py (cpu_baseline / --impl reference).
 * Nothing under otgan_b200/ links or calls this file.
 *
 * PARITY UNPINNED: the reference has no tests / golden vectors and needs TensorFlow 1.x, which cannot
 * be installed here; see oracle/matching_oracle.py for how the oracle is pinned instead.
 *
 * Reference lines followed (relative to /root/reference):
 *   cost blocks, block order            utils/matching.py:16-43      (cosine)   toy_example/matching_cpu.py:17-49 (euclid/n)
 *   Sinkhorn loop, softmax, entropy     utils/matching.py:50-61
 *   twelve products + regrouping        utils/matching.py:64-83
 *   calc_distance                       utils/matching.py:139-153
 *   single-batch variant (+999 I)       utils/matching.py:88-136
 *
 * The file is compiled once; REAL-typed bodies are instantiated for float and double by re-including itself.
 */
#ifndef OTGAN_ORACLE_BODY
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

int otgan_oracle_max_threads(void) { return omp_get_max_threads(); }
void otgan_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

#define OTGAN_ORACLE_BODY
#define REAL float
#define SUF(x) x##_f32
#define EXP expf
#define LOG logf
#include "matching_oracle.c"
#undef REAL
#undef SUF
#undef EXP
#undef LOG
#define REAL double
#define SUF(x) x##_f64
#define EXP exp
#define LOG log
#include "matching_oracle.c"
#undef REAL
#undef SUF
#undef EXP
#undef LOG

#else /* ---------------------------------------------------------------- typed body */

/* G[i][j] = sum_d X[i][d] * Y[j][d]   (tf.matmul(x, y, transpose_b=True), utils/matching.py:31) */
static void SUF(gram)(const REAL* X, const REAL* Y, int m, int n, int D, REAL* G)
{
    const int TI = 4, TJ = 4;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int i0 = 0; i0 < m; i0 += TI)
        for (int j0 = 0; j0 < n; j0 += TJ) {
            REAL acc[4][4];
            for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = 0;
            const int im = (m - i0 < TI) ? m - i0 : TI, jm = (n - j0 < TJ) ? n - j0 : TJ;
            if (im == 4 && jm == 4) {
                const REAL *x0 = X + (size_t)(i0 + 0) * D, *x1 = X + (size_t)(i0 + 1) * D,
                           *x2 = X + (size_t)(i0 + 2) * D, *x3 = X + (size_t)(i0 + 3) * D;
                const REAL *y0 = Y + (size_t)(j0 + 0) * D, *y1 = Y + (size_t)(j0 + 1) * D,
                           *y2 = Y + (size_t)(j0 + 2) * D, *y3 = Y + (size_t)(j0 + 3) * D;
                REAL a00 = 0, a01 = 0, a02 = 0, a03 = 0, a10 = 0, a11 = 0, a12 = 0, a13 = 0,
                     a20 = 0, a21 = 0, a22 = 0, a23 = 0, a30 = 0, a31 = 0, a32 = 0, a33 = 0;
                #pragma omp simd reduction(+:a00,a01,a02,a03,a10,a11,a12,a13,a20,a21,a22,a23,a30,a31,a32,a33)
                for (int d = 0; d < D; ++d) {
                    a00 += x0[d] * y0[d]; a01 += x0[d] * y1[d]; a02 += x0[d] * y2[d]; a03 += x0[d] * y3[d];
                    a10 += x1[d] * y0[d]; a11 += x1[d] * y1[d]; a12 += x1[d] * y2[d]; a13 += x1[d] * y3[d];
                    a20 += x2[d] * y0[d]; a21 += x2[d] * y1[d]; a22 += x2[d] * y2[d]; a23 += x2[d] * y3[d];
                    a30 += x3[d] * y0[d]; a31 += x3[d] * y1[d]; a32 += x3[d] * y2[d]; a33 += x3[d] * y3[d];
                }
                acc[0][0] = a00; acc[0][1] = a01; acc[0][2] = a02; acc[0][3] = a03;
                acc[1][0] = a10; acc[1][1] = a11; acc[1][2] = a12; acc[1][3] = a13;
                acc[2][0] = a20; acc[2][1] = a21; acc[2][2] = a22; acc[2][3] = a23;
                acc[3][0] = a30; acc[3][1] = a31; acc[3][2] = a32; acc[3][3] = a33;
            } else {
                for (int a = 0; a < im; ++a)
                    for (int b = 0; b < jm; ++b) {
                        const REAL* x = X + (size_t)(i0 + a) * D; const REAL* y = Y + (size_t)(j0 + b) * D;
                        REAL s = 0;
                        #pragma omp simd reduction(+:s)
                        for (int d = 0; d < D; ++d) s += x[d] * y[d];
                        acc[a][b] = s;
                    }
            }
            for (int a = 0; a < im; ++a) for (int b = 0; b < jm; ++b) G[(size_t)(i0 + a) * n + j0 + b] = acc[a][b];
        }
}

static void SUF(row_sqmean)(const REAL* X, int m, int D, REAL* out)
{
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        const REAL* x = X + (size_t)i * D; REAL s = 0;
        #pragma omp simd reduction(+:s)
        for (int d = 0; d < D; ++d) s += x[d] * x[d];
        out[i] = s / (REAL)D;
    }
}

/* cost_kind 0: 1 - x.y (matching.py:31);  1: .5 mean(x^2) + .5 mean(y^2) - x.y/D (matching_cpu.py:17-21);
 * diag_add (999 for the single-batch aa/bb blocks, matching.py:109-110) is added on i==j. */
void SUF(otgan_oracle_cost)(const REAL* X, const REAL* Y, int m, int n, int D, int cost_kind, REAL diag_add, REAL* C)
{
    SUF(gram)(X, Y, m, n, D, C);
    REAL *xs = NULL, *ys = NULL;
    if (cost_kind == 1) {
        xs = (REAL*)malloc(sizeof(REAL) * m); ys = (REAL*)malloc(sizeof(REAL) * n);
        SUF(row_sqmean)(X, m, D, xs); SUF(row_sqmean)(Y, n, D, ys);
    }
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            REAL g = C[(size_t)i * n + j], c;
            if (cost_kind == 0) c = (REAL)1 - g;
            else c = (REAL)0.5 * xs[i] + (REAL)0.5 * ys[j] - g / (REAL)D;
            if (i == j && diag_add != 0) c += diag_add;
            C[(size_t)i * n + j] = c;
        }
    free(xs); free(ys);
}

/* nblk blocks [m,n] of cost C -> plans P, per-block entropy and <P,C>.  utils/matching.py:50-57.
 * All blocks advance together so every core has work: (block,row) tasks then (block,column-chunk) tasks. */
void SUF(otgan_oracle_sinkhorn)(const REAL* C, int nblk, int m, int n, REAL lam, int T,
                                REAL* P, REAL* entropy, REAL* pc)
{
    const size_t bsz = (size_t)m * n;
    REAL* L = (REAL*)malloc(sizeof(REAL) * bsz * nblk);
    REAL* ent_rows = (REAL*)malloc(sizeof(REAL) * (size_t)m * nblk);
    REAL* pc_rows = (REAL*)malloc(sizeof(REAL) * (size_t)m * nblk);
    const int CH = 16, nch = (n + CH - 1) / CH;
    #pragma omp parallel
    {
        #pragma omp for schedule(static)
        for (size_t e = 0; e < bsz * nblk; ++e) L[e] = -lam * C[e];                    /* :50 */
        for (int it = 0; it < T; ++it) {
            #pragma omp for schedule(static)
            for (int t = 0; t < nblk * m; ++t) {                                        /* :53 axis=1 */
                REAL* r = L + (size_t)t * n;
                REAL mx = r[0];
                for (int j = 1; j < n; ++j) mx = r[j] > mx ? r[j] : mx;
                REAL s = 0;
                for (int j = 0; j < n; ++j) s += EXP(r[j] - mx);
                const REAL lse = LOG(s) + mx;
                for (int j = 0; j < n; ++j) r[j] -= lse;
            }
            #pragma omp for schedule(static)
            for (int t = 0; t < nblk * nch; ++t) {                                      /* :54 axis=0 */
                const int b = t / nch, j0 = (t % nch) * CH, jn = (n - j0 < CH) ? n - j0 : CH;
                REAL* base = L + (size_t)b * bsz + j0;
                REAL mx[16], s[16];
                for (int j = 0; j < jn; ++j) { mx[j] = base[j]; s[j] = 0; }
                for (int i = 1; i < m; ++i) for (int j = 0; j < jn; ++j) { REAL v = base[(size_t)i * n + j]; mx[j] = v > mx[j] ? v : mx[j]; }
                for (int i = 0; i < m; ++i) for (int j = 0; j < jn; ++j) s[j] += EXP(base[(size_t)i * n + j] - mx[j]);
                for (int j = 0; j < jn; ++j) s[j] = LOG(s[j]) + mx[j];
                for (int i = 0; i < m; ++i) for (int j = 0; j < jn; ++j) base[(size_t)i * n + j] -= s[j];
            }
        }
        #pragma omp for schedule(static)
        for (int t = 0; t < nblk * m; ++t) {                                            /* :56-57 */
            const REAL* r = L + (size_t)t * n; const REAL* c = C + (size_t)t * n; REAL* p = P + (size_t)t * n;
            REAL mx = r[0];
            for (int j = 1; j < n; ++j) mx = r[j] > mx ? r[j] : mx;
            REAL s = 0;
            for (int j = 0; j < n; ++j) { p[j] = EXP(r[j] - mx); s += p[j]; }
            const REAL ls = LOG(s);
            REAL e = 0, q = 0;
            for (int j = 0; j < n; ++j) { p[j] /= s; e -= p[j] * ((r[j] - mx) - ls); q += p[j] * c[j]; }
            ent_rows[t] = e; pc_rows[t] = q;
        }
    }
    for (int b = 0; b < nblk; ++b) {
        REAL e = 0, q = 0;
        for (int i = 0; i < m; ++i) { e += ent_rows[(size_t)b * m + i]; q += pc_rows[(size_t)b * m + i]; }
        entropy[b] = e / (REAL)m; pc[b] = q;
    }
    free(L); free(ent_rows); free(pc_rows);
}

/* O[i][:] (+)= coef * sum_k W(i,k) * F[k][:],  W = Pm (trans=0, [m,kk]) or Pm^T (trans=1, Pm is [kk,m]). */
static void SUF(apply_plan)(const REAL* Pm, int trans, int m, int kk, REAL coef, const REAL* F, int D, REAL* O, int accumulate)
{
    /* Packed, register-blocked micro-kernel.  A [kk x 64] column panel of F is first packed contiguously (rows of F
     * are D*sizeof(REAL) apart -- a power-of-two stride that would alias every row onto one cache set), then an
     * 8-row x (2 SIMD vectors) output tile stays in registers across the whole k loop. */
    typedef REAL vreg __attribute__((vector_size(32)));
    typedef REAL vmem __attribute__((vector_size(32), aligned(sizeof(REAL))));
    enum { RT = 8, VL = (int)(32 / sizeof(REAL)), CT = 2 * VL, PW = 4 * CT };
    const int np_ = (D + PW - 1) / PW, nr = (m + RT - 1) / RT;
    #pragma omp parallel
    {
        REAL* Fp = (REAL*)aligned_alloc(64, sizeof(REAL) * (size_t)kk * PW);
        #pragma omp for schedule(static)
        for (int pnl = 0; pnl < np_; ++pnl) {
            const int p0 = pnl * PW, pn = (D - p0 < PW) ? D - p0 : PW;
            for (int k = 0; k < kk; ++k) {
                const REAL* f = F + (size_t)k * D + p0;
                for (int d = 0; d < pn; ++d) Fp[(size_t)k * PW + d] = f[d];
                for (int d = pn; d < PW; ++d) Fp[(size_t)k * PW + d] = 0;
            }
            for (int it = 0; it < nr; ++it) {
                const int i0 = it * RT, rn = (m - i0 < RT) ? m - i0 : RT;
                for (int cs = 0; cs < PW && cs < pn; cs += CT) {
                    const int dn = (pn - cs < CT) ? pn - cs : CT;
                    REAL tile[RT][CT];
                    vreg a00 = {0}, a01 = {0}, a10 = {0}, a11 = {0}, a20 = {0}, a21 = {0}, a30 = {0}, a31 = {0},
                         a40 = {0}, a41 = {0}, a50 = {0}, a51 = {0}, a60 = {0}, a61 = {0}, a70 = {0}, a71 = {0};
                    REAL wz[RT] = {0};
                    for (int k = 0; k < kk; ++k) {
                        const REAL* f = Fp + (size_t)k * PW + cs;
                        const vreg f0 = *(const vmem*)f, f1 = *(const vmem*)(f + VL);
                        for (int r = 0; r < RT; ++r)
                            wz[r] = (r < rn) ? (trans ? Pm[(size_t)k * m + i0 + r] : Pm[(size_t)(i0 + r) * kk + k]) : (REAL)0;
                        a00 += wz[0] * f0; a01 += wz[0] * f1; a10 += wz[1] * f0; a11 += wz[1] * f1;
                        a20 += wz[2] * f0; a21 += wz[2] * f1; a30 += wz[3] * f0; a31 += wz[3] * f1;
                        a40 += wz[4] * f0; a41 += wz[4] * f1; a50 += wz[5] * f0; a51 += wz[5] * f1;
                        a60 += wz[6] * f0; a61 += wz[6] * f1; a70 += wz[7] * f0; a71 += wz[7] * f1;
                    }
                    *(vmem*)&tile[0][0] = a00; *(vmem*)&tile[0][VL] = a01; *(vmem*)&tile[1][0] = a10; *(vmem*)&tile[1][VL] = a11;
                    *(vmem*)&tile[2][0] = a20; *(vmem*)&tile[2][VL] = a21; *(vmem*)&tile[3][0] = a30; *(vmem*)&tile[3][VL] = a31;
                    *(vmem*)&tile[4][0] = a40; *(vmem*)&tile[4][VL] = a41; *(vmem*)&tile[5][0] = a50; *(vmem*)&tile[5][VL] = a51;
                    *(vmem*)&tile[6][0] = a60; *(vmem*)&tile[6][VL] = a61; *(vmem*)&tile[7][0] = a70; *(vmem*)&tile[7][VL] = a71;
                    for (int r = 0; r < rn; ++r) {
                        REAL* o = O + (size_t)(i0 + r) * D + p0 + cs;
                        if (accumulate) { for (int d = 0; d < dn; ++d) o[d] += coef * tile[r][d]; }
                        else            { for (int d = 0; d < dn; ++d) o[d] = coef * tile[r][d]; }
                    }
                }
            }
        }
        free(Fp);
    }
}

/* Whole two-batch matching call (utils/matching.py:11-85 + calc_distance :139-153) on concatenated features.
 * A = [A1;A2], B = [B1;B2], each [N,D] row-major, h = N/2.  cost_kind as above.
 * Outputs: f_aa,f_bb,f_ab,f_ba [N,D]; P [6,h,h] (may be NULL); entropy (mean of 6); dist (calc_distance, cosine
 * normalisation 1/(2N)); phase_ms[3] = {cost, sinkhorn, matched+distance} wall milliseconds (may be NULL). */
void SUF(otgan_oracle_two_batch)(const REAL* A, const REAL* B, int N, int D, int cost_kind, REAL lam, int T,
                                 REAL* f_aa, REAL* f_bb, REAL* f_ab, REAL* f_ba, REAL* P_out,
                                 REAL* entropy, REAL* dist, double* phase_ms)
{
    const int h = N / 2; const size_t bsz = (size_t)h * h;
    const REAL *A1 = A, *A2 = A + (size_t)h * D, *B1 = B, *B2 = B + (size_t)h * D;
    REAL* C = (REAL*)malloc(sizeof(REAL) * bsz * 6);
    REAL* P = P_out ? P_out : (REAL*)malloc(sizeof(REAL) * bsz * 6);
    REAL ent[6], pc[6];
    double t0 = omp_get_wtime();
    const REAL* X[6] = {A1, B2, A1, A1, A2, A2};          /* matching.py:41-43 block order */
    const REAL* Y[6] = {A2, B1, B1, B2, B1, B2};
    for (int k = 0; k < 6; ++k) SUF(otgan_oracle_cost)(X[k], Y[k], h, h, D, cost_kind, 0, C + k * bsz);
    double t1 = omp_get_wtime();
    SUF(otgan_oracle_sinkhorn)(C, 6, h, h, lam, T, P, ent, pc);
    double t2 = omp_get_wtime();
    const REAL *P0 = P, *P1 = P + bsz, *P2 = P + 2 * bsz, *P3 = P + 3 * bsz, *P4 = P + 4 * bsz, *P5 = P + 5 * bsz;
    const size_t hD = (size_t)h * D;
    SUF(apply_plan)(P0, 0, h, h, 1, A2, D, f_aa, 0);            /* a1_a2  :64 */
    SUF(apply_plan)(P0, 1, h, h, 1, A1, D, f_aa + hD, 0);       /* a2_a1  :70 */
    SUF(apply_plan)(P1, 1, h, h, 1, B2, D, f_bb, 0);            /* b1_b2  :65 */
    SUF(apply_plan)(P1, 0, h, h, 1, B1, D, f_bb + hD, 0);       /* b2_b1  :71 */
    SUF(apply_plan)(P2, 0, h, h, (REAL)0.5, B1, D, f_ab, 0);    /* .5(a1_b1 + a1_b2) :80 */
    SUF(apply_plan)(P3, 0, h, h, (REAL)0.5, B2, D, f_ab, 1);
    SUF(apply_plan)(P4, 0, h, h, (REAL)0.5, B1, D, f_ab + hD, 0);
    SUF(apply_plan)(P5, 0, h, h, (REAL)0.5, B2, D, f_ab + hD, 1);
    SUF(apply_plan)(P2, 1, h, h, (REAL)0.5, A1, D, f_ba, 0);    /* .5(b1_a1 + b1_a2) :82 */
    SUF(apply_plan)(P4, 1, h, h, (REAL)0.5, A2, D, f_ba, 1);
    SUF(apply_plan)(P3, 1, h, h, (REAL)0.5, A1, D, f_ba + hD, 0);
    SUF(apply_plan)(P5, 1, h, h, (REAL)0.5, A2, D, f_ba + hD, 1);
    /* calc_distance :147-152 */
    REAL naa = 0, nbb = 0, nab = 0;
    #pragma omp parallel for reduction(+:naa,nbb,nab) schedule(static)
    for (int i = 0; i < N; ++i) {
        const REAL *a = A + (size_t)i * D, *b = B + (size_t)i * D;
        const REAL *faa = f_aa + (size_t)i * D, *fbb = f_bb + (size_t)i * D, *fab = f_ab + (size_t)i * D;
        REAL s0 = 0, s1 = 0, s2 = 0;
        #pragma omp simd reduction(+:s0,s1,s2)
        for (int d = 0; d < D; ++d) { s0 += a[d] * faa[d]; s1 += b[d] * fbb[d]; s2 += a[d] * fab[d]; }
        naa += s0; nbb += s1; nab += s2;
    }
    double t3 = omp_get_wtime();
    REAL e = 0; for (int k = 0; k < 6; ++k) e += ent[k];
    *entropy = e / (REAL)6;
    *dist = (nbb + naa - (REAL)2 * nab) / (REAL)(2 * N);
    if (phase_ms) { phase_ms[0] = (t1 - t0) * 1e3; phase_ms[1] = (t2 - t1) * 1e3; phase_ms[2] = (t3 - t2) * 1e3; }
    free(C); if (!P_out) free(P);
}

/* Single-batch variant (utils/matching.py:88-136): three N x N blocks, +999 on the aa/bb diagonals. */
void SUF(otgan_oracle_single_batch)(const REAL* A, const REAL* B, int N, int D, int cost_kind, REAL lam, int T,
                                    REAL* f_aa, REAL* f_bb, REAL* f_ab, REAL* f_ba, REAL* P_out, REAL* entropy)
{
    const size_t bsz = (size_t)N * N;
    REAL* C = (REAL*)malloc(sizeof(REAL) * bsz * 3);
    REAL* P = P_out ? P_out : (REAL*)malloc(sizeof(REAL) * bsz * 3);
    REAL ent[3], pc[3];
    SUF(otgan_oracle_cost)(A, A, N, N, D, cost_kind, (REAL)999, C);
    SUF(otgan_oracle_cost)(B, B, N, N, D, cost_kind, (REAL)999, C + bsz);
    SUF(otgan_oracle_cost)(A, B, N, N, D, cost_kind, 0, C + 2 * bsz);
    SUF(otgan_oracle_sinkhorn)(C, 3, N, N, lam, T, P, ent, pc);
    SUF(apply_plan)(P, 0, N, N, 1, A, D, f_aa, 0);
    SUF(apply_plan)(P + bsz, 0, N, N, 1, B, D, f_bb, 0);
    SUF(apply_plan)(P + 2 * bsz, 0, N, N, 1, B, D, f_ab, 0);
    SUF(apply_plan)(P + 2 * bsz, 1, N, N, 1, A, D, f_ba, 0);
    *entropy = (ent[0] + ent[1] + ent[2]) / (REAL)3;
    free(C); if (!P_out) free(P);
}

#endif /* OTGAN_ORACLE_BODY */
