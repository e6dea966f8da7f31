// act.cu -- fused activation kernels of the conv stacks (utils/nn.py:190-206,234-241; models/dcgan.py:35-48):
//   crelu_pad : z = pad_SAME( relu(concat([x, -x], C)) )  -- CReLU pre-activation written straight into the zero-padded
//               NHWC buffer the following (stride-2, asymmetric TensorFlow 'SAME') convolution reads: one pass instead of
//               neg + concat + relu + pad (4 passes over tensors of up to 0.5 GB).
//   glu_up    : out = upsample_nn_x{1,2}( a * sigmoid(l) ), (a, l) = split(y, 2, C)  -- the generator's gated linear unit
//               fused with tf.image.resize_nearest_neighbor (src = dst // 2).
// Both are HBM-bound streaming kernels (float4, one read + one write of the big tensor), forward and backward.
#include "common.cuh"

namespace otgan {

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// x: [B,H,W,C] -> z: [B,Hp,Wp,2C] interior; C % 4 == 0
__global__ void __launch_bounds__(256)
crelu_pad_fwd_kernel(size_t n4, int H, int W, int C, int pt, int pl, int Hp, int Wp, const float* __restrict__ x,
                     float* __restrict__ z)
{
    const int C4 = C >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const size_t b = r / H;
        const float4 v = ld4(x + i * 4);
        float* dst = z + (((b * Hp + h + pt) * Wp + w + pl) * 2 * (size_t)C) + c4 * 4;
        st4(dst, make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)));
        st4(dst + C, make_float4(fmaxf(-v.x, 0.f), fmaxf(-v.y, 0.f), fmaxf(-v.z, 0.f), fmaxf(-v.w, 0.f)));
    }
}

// zero the padding frame of z: one thread per float4 of a border pixel
__global__ void __launch_bounds__(256)
pad_border_zero_kernel(int B, int H, int W, int C2, int pt, int pl, int Hp, int Wp, float* __restrict__ z)
{
    const int C4 = C2 >> 2;
    const int nborder = Hp * Wp - H * W;
    const size_t total = (size_t)B * nborder * C4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        int q = (int)(r % nborder);
        const size_t b = r / nborder;
        // enumerate border pixels: top rows, bottom rows, then left/right columns of the middle rows
        int hp, wp;
        const int top = pt * Wp, bottom = (Hp - pt - H) * Wp;
        if (q < top) { hp = q / Wp; wp = q % Wp; }
        else if (q < top + bottom) { q -= top; hp = pt + H + q / Wp; wp = q % Wp; }
        else {
            q -= top + bottom;
            const int side = Wp - W;                       // border pixels per middle row
            hp = pt + q / side;
            const int s = q % side;
            wp = s < pl ? s : W + s;
        }
        st4(z + ((b * Hp + hp) * Wp + wp) * (size_t)C2 + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

__global__ void __launch_bounds__(256)
crelu_pad_bwd_kernel(size_t n4, int H, int W, int C, int pt, int pl, int Hp, int Wp, const float* __restrict__ x,
                     const float* __restrict__ dz, float* __restrict__ dx)
{
    const int C4 = C >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const size_t b = r / H;
        const float4 v = ld4(x + i * 4);
        const float* src = dz + (((b * Hp + h + pt) * Wp + w + pl) * 2 * (size_t)C) + c4 * 4;
        const float4 gp = ld4(src), gn = ld4(src + C);
        float4 o;
        o.x = (v.x > 0.f ? gp.x : 0.f) - (v.x < 0.f ? gn.x : 0.f);
        o.y = (v.y > 0.f ? gp.y : 0.f) - (v.y < 0.f ? gn.y : 0.f);
        o.z = (v.z > 0.f ? gp.z : 0.f) - (v.z < 0.f ? gn.z : 0.f);
        o.w = (v.w > 0.f ? gp.w : 0.f) - (v.w < 0.f ? gn.w : 0.f);
        st4(dx + i * 4, o);
    }
}

// dy[px][c] = (z[px][c] > 0 ? dz[px][c] : 0) - (z[px][C + c] > 0 ? dz[px][C + c] : 0): the CReLU backward taken from the ACTIVATED tensor
// z = [relu(y) | relu(-y)] (z_pos > 0 <=> y > 0, z_neg > 0 <=> y < 0; relu'(0) = 0 on both sides like the pre-activation form)
__global__ void __launch_bounds__(256)
crelu_bwd_z_kernel(size_t n4, int C4, const float* __restrict__ z, const float* __restrict__ dz, float* __restrict__ dy)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t px = i / (unsigned)C4;
        const unsigned c4 = (unsigned)(i - px * (unsigned)C4);
        const size_t off = (px * 2 * (size_t)C4 + c4) * 4;
        const float4 zp = ld4(z + off), zn = ld4(z + off + 4 * (size_t)C4);
        const float4 gp = ld4(dz + off), gn = ld4(dz + off + 4 * (size_t)C4);
        float4 o;
        o.x = (zp.x > 0.f ? gp.x : 0.f) - (zn.x > 0.f ? gn.x : 0.f);
        o.y = (zp.y > 0.f ? gp.y : 0.f) - (zn.y > 0.f ? gn.y : 0.f);
        o.z = (zp.z > 0.f ? gp.z : 0.f) - (zn.z > 0.f ? gn.z : 0.f);
        o.w = (zp.w > 0.f ? gp.w : 0.f) - (zn.w > 0.f ? gn.w : 0.f);
        st4(dy + i * 4, o);
    }
}

__device__ __forceinline__ float sigmoidf_(float t) { return 1.f / (1.f + __expf(-t)); }

// y: [B,H,W,2C] -> out: [B,UP*H,UP*W,C]
template <int UP>
__global__ void __launch_bounds__(256)
glu_up_fwd_kernel(size_t n4, int H, int W, int C, const float* __restrict__ y, float* __restrict__ out)
{
    const int C4 = C >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const size_t b = r / H;
        const float* src = y + (((b * H + h) * W + w) * 2 * (size_t)C) + c4 * 4;
        const float4 a = ld4(src), l = ld4(src + C);
        const float4 o = make_float4(a.x * sigmoidf_(l.x), a.y * sigmoidf_(l.y), a.z * sigmoidf_(l.z), a.w * sigmoidf_(l.w));
#pragma unroll
        for (int di = 0; di < UP; ++di)
#pragma unroll
            for (int dj = 0; dj < UP; ++dj)
                st4(out + (((b * (UP * H) + UP * h + di) * (size_t)(UP * W) + UP * w + dj) * C) + c4 * 4, o);
    }
}

template <int UP>
__global__ void __launch_bounds__(256)
glu_up_bwd_kernel(size_t n4, int H, int W, int C, const float* __restrict__ y, const float* __restrict__ dout,
                  float* __restrict__ dy)
{
    const int C4 = C >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const size_t b = r / H;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int di = 0; di < UP; ++di)
#pragma unroll
            for (int dj = 0; dj < UP; ++dj) {
                const float4 t = ld4(dout + (((b * (UP * H) + UP * h + di) * (size_t)(UP * W) + UP * w + dj) * C) + c4 * 4);
                g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
            }
        const size_t off = (((b * H + h) * W + w) * 2 * (size_t)C) + c4 * 4;
        const float4 a = ld4(y + off), l = ld4(y + off + C);
        const float4 s = make_float4(sigmoidf_(l.x), sigmoidf_(l.y), sigmoidf_(l.z), sigmoidf_(l.w));
        st4(dy + off, make_float4(g.x * s.x, g.y * s.y, g.z * s.z, g.w * s.w));
        st4(dy + off + C, make_float4(g.x * a.x * s.x * (1.f - s.x), g.y * a.y * s.y * (1.f - s.y),
                                      g.z * a.z * s.z * (1.f - s.z), g.w * a.w * s.w * (1.f - s.w)));
    }
}

unsigned grid_for(size_t n)
{
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 16;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

int crelu_pad_fwd_launch(int B, int H, int W, int C, int pt, int pl, int pb, int pr, const float* x, float* z, cudaStream_t stream)
{
    const int Hp = H + pt + pb, Wp = W + pl + pr;
    const size_t n4 = (size_t)B * H * W * C / 4;
    crelu_pad_fwd_kernel<<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, pt, pl, Hp, Wp, x, z);
    OTGAN_CHECK_LAUNCH("crelu_pad_fwd_kernel");
    if (Hp * Wp > H * W) {
        const size_t nb = (size_t)B * (Hp * Wp - H * W) * (2 * C / 4);
        pad_border_zero_kernel<<<grid_for(nb), 256, 0, stream>>>(B, H, W, 2 * C, pt, pl, Hp, Wp, z);
        OTGAN_CHECK_LAUNCH("pad_border_zero_kernel");
    }
    return OTGAN_OK;
}

int crelu_pad_bwd_launch(int B, int H, int W, int C, int pt, int pl, int pb, int pr, const float* x, const float* dz, float* dx,
                         cudaStream_t stream)
{
    const int Hp = H + pt + pb, Wp = W + pl + pr;
    const size_t n4 = (size_t)B * H * W * C / 4;
    crelu_pad_bwd_kernel<<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, pt, pl, Hp, Wp, x, dz, dx);
    OTGAN_CHECK_LAUNCH("crelu_pad_bwd_kernel");
    return OTGAN_OK;
}

int crelu_bwd_z_launch(size_t P, int C, const float* z, const float* dz, float* dy, cudaStream_t stream)
{
    const size_t n4 = P * (size_t)C / 4;
    crelu_bwd_z_kernel<<<grid_for(n4), 256, 0, stream>>>(n4, C / 4, z, dz, dy);
    OTGAN_CHECK_LAUNCH("crelu_bwd_z_kernel");
    return OTGAN_OK;
}

int glu_up_fwd_launch(int B, int H, int W, int C, int up, const float* y, float* out, cudaStream_t stream)
{
    const size_t n4 = (size_t)B * H * W * C / 4;
    if (up == 2) glu_up_fwd_kernel<2><<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, y, out);
    else         glu_up_fwd_kernel<1><<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, y, out);
    OTGAN_CHECK_LAUNCH("glu_up_fwd_kernel");
    return OTGAN_OK;
}

int glu_up_bwd_launch(int B, int H, int W, int C, int up, const float* y, const float* dout, float* dy, cudaStream_t stream)
{
    const size_t n4 = (size_t)B * H * W * C / 4;
    if (up == 2) glu_up_bwd_kernel<2><<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, y, dout, dy);
    else         glu_up_bwd_kernel<1><<<grid_for(n4), 256, 0, stream>>>(n4, H, W, C, y, dout, dy);
    OTGAN_CHECK_LAUNCH("glu_up_bwd_kernel");
    return OTGAN_OK;
}

}  // namespace otgan
