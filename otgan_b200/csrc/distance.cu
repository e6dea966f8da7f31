// distance.cu -- calc_distance (utils/matching.py:139-153; toy_example/matching_cpu.py:155-164) as one streaming pass.
//   out = ( sum(B * f_bb) + sum(A * f_aa) - 2 sum(A * f_ab) ) * scale
// HBM-bound: five [n, D] arrays read once (20 n D bytes).  One CTA per row, per-row partials, fixed-order final sum.
#include "common.cuh"

namespace otgan {

namespace {

constexpr int NT = 256;

template <bool VEC>
__global__ void __launch_bounds__(NT)
distance_rows_kernel(int D, const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Faa,
                     const float* __restrict__ Fbb, const float* __restrict__ Fab, int ld, float* __restrict__ part)
{
    __shared__ float red[3][NT / 32];
    const size_t off = (size_t)blockIdx.x * ld;
    const float *a = A + off, *b = B + off, *faa = Faa + off, *fbb = Fbb + off, *fab = Fab + off;
    float saa = 0.f, sbb = 0.f, sab = 0.f;
    if (VEC) {
        for (int d = threadIdx.x * 4; d < D; d += NT * 4) {
            const float4 va = *reinterpret_cast<const float4*>(a + d), vb = *reinterpret_cast<const float4*>(b + d);
            const float4 x = *reinterpret_cast<const float4*>(faa + d), y = *reinterpret_cast<const float4*>(fbb + d);
            const float4 z = *reinterpret_cast<const float4*>(fab + d);
            saa += va.x * x.x + va.y * x.y + va.z * x.z + va.w * x.w;
            sbb += vb.x * y.x + vb.y * y.y + vb.z * y.z + vb.w * y.w;
            sab += va.x * z.x + va.y * z.y + va.z * z.z + va.w * z.w;
        }
    } else {
        for (int d = threadIdx.x; d < D; d += NT) {
            saa = fmaf(a[d], faa[d], saa);
            sbb = fmaf(b[d], fbb[d], sbb);
            sab = fmaf(a[d], fab[d], sab);
        }
    }
    saa = warp_sum(saa); sbb = warp_sum(sbb); sab = warp_sum(sab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = saa; red[1][warp] = sbb; red[2][warp] = sab; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[threadIdx.x][w];
        part[threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void distance_final_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out)
{
    __shared__ float tot[3];
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += part[threadIdx.x * n + i];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = (tot[1] + tot[0] - 2.f * tot[2]) * scale;   // nd_b_b + nd_a_a - 2*nd_a_b  (:150)
}

}  // namespace

size_t distance_workspace_bytes(int n, int) { return (size_t)3 * n * sizeof(float); }

int distance_launch(int n, int D, const float* A, const float* B, const float* f_aa, const float* f_bb,
                    const float* f_ab, int ld, float scale, float* out, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(ws_bytes >= distance_workspace_bytes(n, D), "calc_distance: workspace too small");
    float* part = reinterpret_cast<float*>(ws);
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && aligned16(A) && aligned16(B) && aligned16(f_aa) &&
                     aligned16(f_bb) && aligned16(f_ab);
    if (vec) distance_rows_kernel<true><<<n, NT, 0, stream>>>(D, A, B, f_aa, f_bb, f_ab, ld, part);
    else     distance_rows_kernel<false><<<n, NT, 0, stream>>>(D, A, B, f_aa, f_bb, f_ab, ld, part);
    OTGAN_CHECK_LAUNCH("distance_rows_kernel");
    distance_final_kernel<<<1, 32, 0, stream>>>(part, n, scale, out);
    OTGAN_CHECK_LAUNCH("distance_final_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
