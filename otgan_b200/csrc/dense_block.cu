// dense_block.cu -- DenseNet dense blocks (models/densenet.py:10-15 `block`, utils/nn.py:190-206 CReLU over LISTS, :234-241 conv)
// on the generic tcgen05 convolution kernels of conv_tc.cu.
//
// The reference keeps a Python list x = [e_0, e_1, ...] and every layer r computes
//     e_{r+1} = conv3x3( relu(concat([e_0, -e_0, e_1, -e_1, ...])), W_r ) + b_r                (growth = 16 new channels)
// i.e. it re-concatenates and re-activates all previous outputs for every layer (O(L^2) copies).  Here each block owns ONE
// NHWC buffer Z [B, H, W, Ctot], Ctot = 2 (c0 + 16 L), holding the ACTIVATED features in "crelu8" slot order: list element i with
// c_i channels at raw offset o_i occupies channels [2 o_i, 2 o_i + 2 c_i) as blocks of 8 positive parts followed by their 8
// negative parts.  Layer r reads the channel PREFIX [0, 2 c0 + 32 r) of Z through a TMA view (no copy) and its epilogue writes
// relu(y), relu(-y) straight into the next slot -- the concatenation and the CReLU cost nothing.  The input-channel order of
// every filter is permuted accordingly when W = g V / ||V|| is built (weightnorm.cu `perm`), so the variables keep the
// reference's HWIO layout and names.
//
// Forward, "contribution" form.  A 16-filter layer is a poor tensor-core shape: with N = 16 the MMA is bound by reading the
// 128-pixel activation operand from shared memory (4 KB per K = 8 step, 32 cycles) while the tensor pipe needs 8.  So the
// forward pass is re-associated: as soon as slot j of Z exists, ONE convolution adds its contribution to the pre-activations of
// ALL later layers -- K = the slot's channels, N = 16 (L - j) columns of an accumulator S [B, H, W, 16 L] (read-modify-write in
// the epilogue) -- and the first 16 of those columns, which are complete at that point, leave the same epilogue as crelu8 into
// slot j + 1.  Same products, same sums (per layer the contributions are added slot by slot instead of tap by tap), 3.9x fewer
// tensor-pipe cycles; launch 0 (the base slots, K = 2 c0) also carries the biases.  Weights for this form: WF_all [16 L][9][Ctot],
// WF_all[16 q + co][t][ci] = W_q[co][t][ci] (0 for ci >= cin_q) -- the layout of dW_all -- read as a K-slice / row-slice.
//
// Backward (what tf.gradients emits for the list graph), per block, with dZ = gradient w.r.t. Z from the consumer:
//     for r = L-1 .. 0:   g      = dZ[slot r+1] + sum_{q > r} conv_T( dy_q, W_q[:, slot r+1] )        "gather" form: ONE convolution with
//                         dy_r   = crelu'(g)                                                            K = 16 (L-1-r) per tap, N = 32, the
//                                                                                                       CReLU backward fused in its epilogue
//     base slots likewise (N = 2 c_i, K = 16 L);   dW_all [16 L][9][Ctot] = ONE wgrad GEMM of dY [.., 16 L] against Z;
//     db = column sums of dY.
// dY [B, H, W, 16 L] holds all dy_r; the weight operand of the gather convolutions is WB [Ctot][9][16 L],
// WB[ci][t][16 q + co] = W_q[co][t][ci] (0 where layer q does not see channel ci), built once per weight update.
#include "conv_ex.cuh"
#include <string.h>

namespace otgan {

namespace {

// z[p][(c/8)*16 + c%8] = relu(x[p][c]), z[p][(c/8)*16 + 8 + c%8] = relu(-x[p][c]); one thread per 8 channels of a pixel
__global__ void __launch_bounds__(256)
crelu8_fwd_kernel(long long P, int C8, const float* __restrict__ x, int ldx, float* __restrict__ z, int ldz)
{
    const long long total = P * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8;
        const int b = (int)(i - p * C8);
        const float4 a0 = *reinterpret_cast<const float4*>(x + p * ldx + 8 * b);
        const float4 a1 = *reinterpret_cast<const float4*>(x + p * ldx + 8 * b + 4);
        float* o = z + p * ldz + 16 * b;
        *reinterpret_cast<float4*>(o) = make_float4(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f), fmaxf(a0.z, 0.f), fmaxf(a0.w, 0.f));
        *reinterpret_cast<float4*>(o + 4) = make_float4(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f), fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f));
        *reinterpret_cast<float4*>(o + 8) = make_float4(fmaxf(-a0.x, 0.f), fmaxf(-a0.y, 0.f), fmaxf(-a0.z, 0.f), fmaxf(-a0.w, 0.f));
        *reinterpret_cast<float4*>(o + 12) = make_float4(fmaxf(-a1.x, 0.f), fmaxf(-a1.y, 0.f), fmaxf(-a1.z, 0.f), fmaxf(-a1.w, 0.f));
    }
}

// dx[p][c] = (z_pos > 0 ? dz_pos : 0) - (z_neg > 0 ? dz_neg : 0)   (z, dz in crelu8 slot order)
__global__ void __launch_bounds__(256)
crelu8_bwd_kernel(long long P, int C8, const float* __restrict__ z, int ldz, const float* __restrict__ dz, int lddz,
                  float* __restrict__ dx, int lddx)
{
    const long long total = P * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8;
        const int b = (int)(i - p * C8);
        const float* zz = z + p * ldz + 16 * b;
        const float* dd = dz + p * lddz + 16 * b;
        float o[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 zp = *reinterpret_cast<const float4*>(zz + 4 * h), zn = *reinterpret_cast<const float4*>(zz + 8 + 4 * h);
            const float4 gp = *reinterpret_cast<const float4*>(dd + 4 * h), gn = *reinterpret_cast<const float4*>(dd + 8 + 4 * h);
            o[4 * h + 0] = (zp.x > 0.f ? gp.x : 0.f) - (zn.x > 0.f ? gn.x : 0.f);
            o[4 * h + 1] = (zp.y > 0.f ? gp.y : 0.f) - (zn.y > 0.f ? gn.y : 0.f);
            o[4 * h + 2] = (zp.z > 0.f ? gp.z : 0.f) - (zn.z > 0.f ? gn.z : 0.f);
            o[4 * h + 3] = (zp.w > 0.f ? gp.w : 0.f) - (zn.w > 0.f ? gn.w : 0.f);
        }
        float* out = dx + p * lddx + 8 * b;
        *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(out + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// WB[ci][t][k] = WF_all[k][t][ci]  (k = 16 q + co; the zeros of WF_all -- layer q does not see channel ci -- carry over);
// 32 x 32 tiles through shared memory, one tap per blockIdx.z
__global__ void __launch_bounds__(256)
dense_wb_kernel(int KY, int taps, int Ctot, const float* __restrict__ WF, float* __restrict__ WB)
{
    __shared__ float tile[32][33];
    const int t = blockIdx.z, ci0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int k = k0 + i, ci = ci0 + tx;
        tile[i][tx] = (k < KY && ci < Ctot) ? WF[((size_t)k * taps + t) * Ctot + ci] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int ci = ci0 + i, k = k0 + tx;
        if (ci < Ctot && k < KY) WB[((size_t)ci * taps + t) * KY + k] = tile[tx][i];
    }
}

inline unsigned ew_blocks(long long n)
{
    const long long b = (n + 255) / 256, cap = (long long)kNumSMs * 16;
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

struct Geom {
    int B, H, W, n_base, base_ch[4], L, G;
    int c0() const { int s = 0; for (int i = 0; i < n_base; ++i) s += base_ch[i]; return s; }
    int ctot() const { return 2 * (c0() + L * G); }
    int cin(int r) const { return 2 * (c0() + r * G); }
};

bool geom_ok(const otgan_dense_geom_t* g, Geom& out)
{
    if (!g || g->B < 1 || g->H < 1 || g->W < 1 || g->n_base < 1 || g->n_base > 4 || g->L < 1 || g->L > 32 || g->growth != 16) return false;
    for (int i = 0; i < g->n_base; ++i)
        if (g->base_ch[i] < 8 || (g->base_ch[i] & 7)) return false;
    out.B = g->B; out.H = g->H; out.W = g->W; out.n_base = g->n_base; out.L = g->L; out.G = g->growth;
    for (int i = 0; i < 4; ++i) out.base_ch[i] = i < g->n_base ? g->base_ch[i] : 0;
    return out.G * out.L <= 1024;
}

}  // namespace

int crelu8_fwd_launch(long long P, int C, const float* x, int ldx, float* z, int ldz, cudaStream_t stream)
{
    OTGAN_REQUIRE(P >= 1 && C >= 8 && !(C & 7) && !(ldx & 3) && !(ldz & 3) && aligned16(x) && aligned16(z), "crelu8_fwd: C %% 8, 16-byte aligned rows");
    crelu8_fwd_kernel<<<ew_blocks(P * (C / 8)), 256, 0, stream>>>(P, C / 8, x, ldx, z, ldz);
    OTGAN_CHECK_LAUNCH("crelu8_fwd_kernel");
    return OTGAN_OK;
}

int crelu8_bwd_launch(long long P, int C, const float* z, int ldz, const float* dz, int lddz, float* dx, int lddx, cudaStream_t stream)
{
    OTGAN_REQUIRE(P >= 1 && C >= 8 && !(C & 7) && !(ldz & 3) && !(lddz & 3) && !(lddx & 3) && aligned16(z) && aligned16(dz) && aligned16(dx),
                  "crelu8_bwd: C %% 8, 16-byte aligned rows");
    crelu8_bwd_kernel<<<ew_blocks(P * (C / 8)), 256, 0, stream>>>(P, C / 8, z, ldz, dz, lddz, dx, lddx);
    OTGAN_CHECK_LAUNCH("crelu8_bwd_kernel");
    return OTGAN_OK;
}

int dense_channels(const otgan_dense_geom_t* g)
{
    Geom q;
    return geom_ok(g, q) ? q.ctot() : -1;
}

size_t dense_wb_floats(const otgan_dense_geom_t* g)
{
    Geom q;
    return geom_ok(g, q) ? (size_t)q.ctot() * 9 * q.G * q.L : 0;
}

int dense_build_wb_launch(const otgan_dense_geom_t* g, const float* wf_all, float* WB, cudaStream_t stream)
{
    Geom q;
    OTGAN_REQUIRE(geom_ok(g, q) && wf_all && WB, "dense_build_wb: bad geometry / null pointer");
    const int KY = q.G * q.L, Ctot = q.ctot();
    dense_wb_kernel<<<dim3(ceil_div(Ctot, 32), ceil_div(KY, 32), 9), 256, 0, stream>>>(KY, 9, Ctot, wf_all, WB);
    OTGAN_CHECK_LAUNCH("dense_wb_kernel");
    return OTGAN_OK;
}

// Z base slots must already hold crelu8(e_0 ...) (crelu8_fwd_launch per base element).  wf_all: [16L][9][Ctot] in Z channel order
// (zero where a layer does not see a channel), bias_all: [16L] or null, S: scratch [B,H,W,16L] (the pre-activation accumulator).
int dense_block_fprop_launch(const otgan_dense_geom_t* g, const float* wf_all, const float* bias_all, float* Z, float* S, cudaStream_t stream)
{
    Geom q;
    OTGAN_REQUIRE(geom_ok(g, q) && wf_all && Z && S, "dense_block_fprop: bad geometry / null pointer");
    const int Ctot = q.ctot(), KY = q.G * q.L;
    for (int j = 0; j < q.L; ++j) {
        ConvEx c;
        memset(&c, 0, sizeof(c));
        c.B = q.B; c.H = q.H; c.W = q.W; c.kh = 3; c.kw = 3; c.stride = 1; c.pt = 1; c.pl = 1;
        const int k0 = j == 0 ? 0 : q.cin(j - 1);                      // input slot j: channels [k0, cin(j))
        c.a = Z + k0; c.Ka = q.cin(j) - k0; c.lda = Ctot;
        c.out = S + q.G * j; c.N = q.G * (q.L - j); c.ldo = KY;
        c.w = wf_all; c.w_K = Ctot; c.w_taps = 9; c.w_rows = KY; c.w_ldtap = Ctot; c.w_ldrow = 9LL * Ctot; c.k0 = k0; c.row0 = q.G * j;
        c.bias = (j == 0 && bias_all) ? bias_all : nullptr;
        c.epi_mode = EPI_DENSE_FWD; c.accumulate = j > 0; c.e_out2 = Z + q.cin(j); c.e_ld = Ctot;
        const int rc = conv_fprop_ex_launch(c, stream);
        if (rc != OTGAN_OK) return rc;
    }
    return OTGAN_OK;
}

size_t dense_bgrad_workspace_bytes(const otgan_dense_geom_t* g)
{
    Geom q;
    if (!geom_ok(g, q)) return 0;
    const size_t a = conv_wgrad_ex_workspace_bytes(q.B, q.H, q.W, q.ctot(), q.G * q.L, 3, 3);
    const size_t b = colsum_workspace_bytes(q.B * q.H * q.W, q.G * q.L);
    return (a > b ? a : b) + 256;
}

// dZ: gradient w.r.t. Z from the block's consumer (read only).  Outputs: dY [B,H,W,16L] (scratch + db source), dbase[i] [B,H,W,c_i],
// dW_all [16L][9][Ctot] (rows of layer q valid for ci < cin_q), db_all [16L]; either may be null (no parameter gradients wanted).
int dense_block_bgrad_launch(const otgan_dense_geom_t* g, const float* Z, const float* dZ, const float* WB, float* dY,
                             float* const* dbase, float* dW_all, float* db_all, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    Geom q;
    OTGAN_REQUIRE(geom_ok(g, q) && Z && dZ && WB && dY && dbase, "dense_block_bgrad: bad geometry / null pointer");
    OTGAN_REQUIRE(ws && ws_bytes >= dense_bgrad_workspace_bytes(g), "dense_block_bgrad: workspace too small");
    const int Ctot = q.ctot(), KY = q.G * q.L;
    const long long P = (long long)q.B * q.H * q.W;
    // last layer: nothing downstream inside the block
    {
        const int slot = q.cin(q.L - 1);
        const int rc = crelu8_bwd_launch(P, q.G, Z + slot, Ctot, dZ + slot, Ctot, dY + (size_t)q.G * (q.L - 1), KY, stream);
        if (rc != OTGAN_OK) return rc;
    }
    auto gather = [&](int k_first, int row0, int nrows, float* out, int ldo) -> int {
        ConvEx c;
        memset(&c, 0, sizeof(c));
        c.B = q.B; c.H = q.H; c.W = q.W; c.kh = 3; c.kw = 3; c.stride = 1; c.pt = 1; c.pl = 1;
        c.a = dY + k_first; c.Ka = KY - k_first; c.lda = KY;
        c.out = out; c.N = nrows; c.ldo = ldo;
        c.w = WB; c.w_K = KY; c.w_taps = 9; c.w_rows = Ctot; c.w_ldtap = KY; c.w_ldrow = 9LL * KY; c.k0 = k_first; c.row0 = row0;
        c.epi_mode = EPI_CRELU8_BWD; c.e_add = dZ + row0; c.e_z = Z + row0; c.e_ld = Ctot;
        return conv_dgrad_ex_launch(c, stream);
    };
    for (int r = q.L - 2; r >= 0; --r) {
        const int rc = gather(q.G * (r + 1), q.cin(r), 2 * q.G, dY + (size_t)q.G * r, KY);
        if (rc != OTGAN_OK) return rc;
    }
    int off = 0;
    for (int i = 0; i < q.n_base; ++i) {
        OTGAN_REQUIRE(dbase[i], "dense_block_bgrad: null base gradient %d", i);
        const int rc = gather(0, 2 * off, 2 * q.base_ch[i], dbase[i], q.base_ch[i]);
        if (rc != OTGAN_OK) return rc;
        off += q.base_ch[i];
    }
    int rc = OTGAN_OK;
    if (dW_all) rc = conv_wgrad_ex_launch(q.B, q.H, q.W, Ctot, Ctot, KY, KY, 3, 3, 1, 1, 1, dY, Z, dW_all, ws, ws_bytes, stream);
    if (rc != OTGAN_OK) return rc;
    if (db_all) rc = colsum_launch((int)P, KY, dY, db_all, ws, ws_bytes, stream);
    return rc;
}

}  // namespace otgan
