// plan_apply_simt.cu -- transport-plan application on the FP32 FMA pipe (exact-fp32 rung of the matched-feature stage).
//
//   out[o] = sum_t coef[o][t] * op(P[blk[o][t]]) * F[src[o][t]]          op = identity | transpose
//
// One launch covers every output group of a matching call: the twelve tf.matmul + tf.split + 0.5*(f1+f2) of
// utils/matching.py:63-83 (API form, 8 output groups / 12 terms), its single-batch form (:131-134), or the fused
// feature-gradient form (4 output groups x 3 terms, which also absorbs train.py:111,125-126).  The contraction length
// is only h, the free dimension is D (tens of thousands): each CTA owns a 128-row x 128-column output tile, streams the
// source rows F[k][d0:d0+128] with cp.async (every F element is read once per output group that uses it) and keeps the
// small plan tiles in shared memory, scaled by coef on the way in.
#include "common.cuh"

namespace otgan {

namespace {

constexpr int TM = 128, TN = 128, BK = 16, NT = 256;

struct PlanArgs {
    otgan_plan_t plan;
    const float* src[OTGAN_MAX_OUTPUTS];
    float* out[OTGAN_MAX_OUTPUTS];
};

// grid = (ceil(D/128), ceil(h/128), n_out)
template <bool VEC_F, bool VEC_P>
__global__ void __launch_bounds__(NT, 2)
plan_apply_kernel(PlanArgs args, int h, int D, const float* __restrict__ P, int ldf, int ldo)
{
    __shared__ __align__(16) float As[2][BK][TM];   // As[k][i] = coef * op(P)[m0+i][k0+k]
    __shared__ __align__(16) float Bs[2][BK][TN];   // Bs[k][c] = F[k0+k][d0+c]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int d0 = blockIdx.x * TN, m0 = blockIdx.y * TM, o = blockIdx.z;
    const int nterms = args.plan.nterms[o];
    const int ksteps = (h + BK - 1) / BK, total = nterms * ksteps;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float areg[8];
    auto load_a = [&](int step) {       // global -> registers (scaled by coef)
        const int t = step / ksteps, k0 = (step % ksteps) * BK;
        const float* __restrict__ Pm = P + (size_t)args.plan.blk[o][t] * h * h;
        const float coef = args.plan.coef[o][t];
        if (!args.plan.trans[o][t]) {   // As[k][i] = P[m0+i][k0+k]; thread reads 8 consecutive k of one row
            const int i = tid & 127, kh = tid >> 7, r = m0 + i, kb = k0 + kh * 8;
            if (VEC_P && r < h && kb + 8 <= h) {
                const float4 u = *reinterpret_cast<const float4*>(Pm + (size_t)r * h + kb);
                const float4 v = *reinterpret_cast<const float4*>(Pm + (size_t)r * h + kb + 4);
                areg[0] = u.x; areg[1] = u.y; areg[2] = u.z; areg[3] = u.w;
                areg[4] = v.x; areg[5] = v.y; areg[6] = v.z; areg[7] = v.w;
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) areg[c] = (r < h && kb + c < h) ? Pm[(size_t)r * h + kb + c] : 0.f;
            }
        } else {                        // As[k][i] = P[k0+k][m0+i]; thread reads 2 x 4 consecutive i of one row k
            const int k = tid >> 4, i4 = (tid & 15) * 4, kr = k0 + k;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c = m0 + i4 + 64 * half;
                if (VEC_P && kr < h && c + 4 <= h) {
                    const float4 u = *reinterpret_cast<const float4*>(Pm + (size_t)kr * h + c);
                    areg[4 * half + 0] = u.x; areg[4 * half + 1] = u.y; areg[4 * half + 2] = u.z; areg[4 * half + 3] = u.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) areg[4 * half + q] = (kr < h && c + q < h) ? Pm[(size_t)kr * h + c + q] : 0.f;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) areg[c] *= coef;
    };
    auto store_a = [&](int step, int buf) {
        const int t = step / ksteps;
        if (!args.plan.trans[o][t]) {
            const int i = tid & 127, kh = tid >> 7;
#pragma unroll
            for (int c = 0; c < 8; ++c) As[buf][kh * 8 + c][i] = areg[c];
        } else {
            const int k = tid >> 4, i4 = (tid & 15) * 4;
            *reinterpret_cast<float4*>(&As[buf][k][i4]) = make_float4(areg[0], areg[1], areg[2], areg[3]);
            *reinterpret_cast<float4*>(&As[buf][k][i4 + 64]) = make_float4(areg[4], areg[5], areg[6], areg[7]);
        }
    };
    auto load_b = [&](int step, int buf) {   // cp.async global -> smem
        const int t = step / ksteps, k0 = (step % ksteps) * BK;
        const float* __restrict__ F = args.src[args.plan.src[o][t]];
        if (VEC_F) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT, k = idx >> 5, c4 = (idx & 31) * 4;
                const int kr = k0 + k, d = d0 + c4;
                int bytes = (D - d) * 4;
                bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
                if (kr >= h) bytes = 0;
                const float* srcp = F + (bytes > 0 ? (size_t)kr * ldf + d : 0);
                cp_async16(&Bs[buf][k][c4], srcp, bytes);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int idx = tid + q * NT, k = idx >> 7, c = idx & 127;
                const int kr = k0 + k, d = d0 + c;
                const bool ok = kr < h && d < D;
                cp_async4(&Bs[buf][k][c], F + (ok ? (size_t)kr * ldf + d : 0), ok ? 4 : 0);
            }
        }
    };

    load_a(0);
    load_b(0, 0);
    cp_async_commit();
    store_a(0, 0);
    cp_async_wait<0>();
    __syncthreads();

    for (int step = 0; step < total; ++step) {
        const int buf = step & 1;
        const bool more = step + 1 < total;
        if (more) {
            load_a(step + 1);
            load_b(step + 1, buf ^ 1);
            cp_async_commit();
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + 4 * tx]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store_a(step + 1, buf ^ 1);
            cp_async_wait<0>();
        }
        __syncthreads();
    }

    float* __restrict__ out = args.out[o];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
        if (r >= h) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int d = d0 + 4 * tx + 64 * half;
            float* dst = out + (size_t)r * ldo + d;
            if (VEC_F && d + 4 <= D) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][4 * half], acc[i][4 * half + 1], acc[i][4 * half + 2], acc[i][4 * half + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (d + q < D) dst[q] = acc[i][4 * half + q];
            }
        }
    }
}

}  // namespace

int plan_apply_simt_launch(const otgan_plan_t* plan, int h, int D, const float* P, const float* const* F, int ldf,
                           float* const* out, int ldo, cudaStream_t stream)
{
    PlanArgs args;
    args.plan = *plan;
    bool vec_f = (D % 4 == 0) && (ldf % 4 == 0) && (ldo % 4 == 0);
    int nsrc = 0;
    for (int o = 0; o < plan->n_out; ++o)
        for (int t = 0; t < plan->nterms[o]; ++t) nsrc = plan->src[o][t] + 1 > nsrc ? plan->src[o][t] + 1 : nsrc;
    for (int s = 0; s < OTGAN_MAX_OUTPUTS; ++s) {
        args.src[s] = s < nsrc ? F[s] : nullptr;
        args.out[s] = s < plan->n_out ? out[s] : nullptr;
        if (s < nsrc) vec_f = vec_f && aligned16(F[s]);
        if (s < plan->n_out) vec_f = vec_f && aligned16(out[s]);
    }
    const bool vec_p = (h % 4 == 0) && aligned16(P);
    dim3 grid(ceil_div(D, TN), ceil_div(h, TM), plan->n_out);
    if (vec_f && vec_p) plan_apply_kernel<true, true><<<grid, NT, 0, stream>>>(args, h, D, P, ldf, ldo);
    else if (vec_f)     plan_apply_kernel<true, false><<<grid, NT, 0, stream>>>(args, h, D, P, ldf, ldo);
    else if (vec_p)     plan_apply_kernel<false, true><<<grid, NT, 0, stream>>>(args, h, D, P, ldf, ldo);
    else                plan_apply_kernel<false, false><<<grid, NT, 0, stream>>>(args, h, D, P, ldf, ldo);
    OTGAN_CHECK_LAUNCH("plan_apply_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
