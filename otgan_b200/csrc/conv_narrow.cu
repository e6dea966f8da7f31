// conv_narrow.cu -- layout helpers for the two 3-channel convolutions of DCGAN (critic conv2d_0: 3 -> 128, generator
// conv2d_3: 128 -> 3; models/dcgan.py:11,52).  With only 3 channels on one side the per-tap implicit GEMM of conv_tc.cu
// would re-read the wide tensor once per filter tap with almost no reuse (measured: 407 us, L2-bound), so these layers are
// computed as ONE plain GEMM over all 75 (tap, channel) columns plus a cheap shift on the narrow side:
//   narrow OUTPUT (generator conv2d_3 fprop, critic conv2d_0 dgrad):  z[px][t*C + c] = sum_k wide[px][k] * w2[t*C + c][k]
//       (conv_gemm_tc_kernel as a 1x1 convolution, the wide tensor is read once), then  col2im_narrow:
//       y[px][c] = bias[c] + sum_t z[px + off_t][t*C + c]
//   narrow INPUT in a filter gradient (critic conv2d_0 wgrad, generator conv2d_3 wgrad):  im2col_narrow:
//       col[px][t*C + c] = narrow[px + off_t][c], then dW = wide^T * col  (conv_wgrad_tc_kernel as a 1x1 convolution).
// off_t = (kh - pad_top, kw - pad_left) for a forward-oriented shift, negated with flip = 1 (gradient orientation).
// Both kernels are HBM-bound streaming passes over the [pixels, 128] matrix; fixed summation order (deterministic).
#include "common.cuh"

namespace otgan {

namespace {

// col: [P, ldc]; one thread per float4 of a row.  The column -> (row shift, column shift, channel) decode is the same for
// every pixel, so it is tabulated once per block in shared memory (ldc <= 256) instead of being re-derived with integer
// divisions for each of the 4 elements a thread writes.
__global__ void __launch_bounds__(256)
im2col_narrow_kernel(size_t n4, int H, int W, int C, int kh, int kw, int pt, int pl, int flip, int ldc,
                     const float* __restrict__ x, float* __restrict__ col)
{
    __shared__ int tab[256];                                  // (dh + 64) | (dw + 64) << 8 | c << 16, or -1 past the last column
    const int g4 = ldc >> 2, ncol = kh * kw * C;
    for (int cidx = threadIdx.x; cidx < ldc; cidx += blockDim.x) {
        int v = -1;
        if (cidx < ncol) {
            const int t = cidx / C, c = cidx - t * C;
            const int a = t / kw, b = t - a * kw;
            const int dh = flip ? pt - a : a - pt, dw = flip ? pl - b : b - pl;
            v = (dh + 64) | ((dw + 64) << 8) | (c << 16);
        }
        tab[cidx] = v;
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % g4);
        const size_t px = i / g4;
        const int w = (int)(px % W), h = (int)((px / W) % H);
        const float* xpx = x + px * C;                        // x[n, h, w, 0]
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int code = tab[4 * j + e];
            float val = 0.f;
            if (code >= 0) {
                const int dh = (code & 0xff) - 64, dw = ((code >> 8) & 0xff) - 64, c = code >> 16;
                if (h + dh >= 0 && h + dh < H && w + dw >= 0 && w + dw < W) val = __ldg(xpx + ((long long)dh * W + dw) * C + c);
            }
            v[e] = val;
        }
        *reinterpret_cast<float4*>(col + px * ldc + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// y: [P, C]; one thread per output element
__global__ void __launch_bounds__(256)
col2im_narrow_kernel(size_t n, int H, int W, int C, int kh, int kw, int pt, int pl, int flip, int ldz,
                     const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ y)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t px = i / C;
        const int w = (int)(px % W), h = (int)((px / W) % H);
        float acc = bias ? __ldg(bias + c) : 0.f;
        for (int a = 0; a < kh; ++a) {
            const int dh = flip ? pt - a : a - pt;
            if (h + dh < 0 || h + dh >= H) continue;
            for (int b = 0; b < kw; ++b) {
                const int dw = flip ? pl - b : b - pl;
                if (w + dw < 0 || w + dw >= W) continue;
                acc += __ldg(z + (px + (long long)dh * W + dw) * ldz + (a * kw + b) * C + c);
            }
        }
        y[i] = acc;
    }
}

inline unsigned grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 16;
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int im2col_narrow_launch(int B, int H, int W, int C, int kh, int kw, int pt, int pl, int flip, const float* x, float* col,
                         int ldc, cudaStream_t stream)
{
    const size_t n4 = (size_t)B * H * W * (ldc / 4);
    im2col_narrow_kernel<<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, kh, kw, pt, pl, flip, ldc, x, col);
    OTGAN_CHECK_LAUNCH("im2col_narrow_kernel");
    return OTGAN_OK;
}

int col2im_narrow_launch(int B, int H, int W, int C, int kh, int kw, int pt, int pl, int flip, const float* z, int ldz,
                         const float* bias, float* y, cudaStream_t stream)
{
    const size_t n = (size_t)B * H * W * C;
    col2im_narrow_kernel<<<grid_for(n), 256, 0, stream>>>(n, H, W, C, kh, kw, pt, pl, flip, ldz, z, bias, y);
    OTGAN_CHECK_LAUNCH("col2im_narrow_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
