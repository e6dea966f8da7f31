// conv_tc.cu -- im2col-free implicit-GEMM convolutions on the 5th-gen tensor cores (TMA -> smem -> tcgen05.mma kind::tf32
// -> TMEM -> registers -> global) for the conv stacks of models/dcgan.py (utils/nn.py:234-241 -> tf.nn.conv2d NHWC, 'SAME').
//
// All three passes are sums over the filter taps of small GEMMs whose operand tiles are plain TMA boxes of the NHWC
// tensors -- no im2col buffer, no padded copy:
//   fprop  y[n,oh,ow,co]  = sum_{kh,kw,ci} x[n, s*oh+kh-pt, s*ow+kw-pl, ci] * W[co,kh,kw,ci]
//          A = a [bw x bh x bn] = 128-pixel box of x shifted by the tap (K-major: the 32 channels of a K-chunk are the
//          contiguous 128-byte rows), B = 256 (or 128) rows of the OHWI weight matrix [Cout, kh*kw*Cin].
//   dgrad  dx[n,ih,iw,ci] = sum_{kh,kw,co} dy[n, (ih+pt-kh)/s, (iw+pl-kw)/s, co] * W[co,kh,kw,ci]
//          the same kernel on dy with the IHWO weight matrix [Cin, kh*kw*Cout]; for stride 2 the output pixels are
//          split into the s*s parity classes (each class sees only the taps with kh = ih+pt mod s: 2x2, 2x3, 3x2, 3x3
//          of the 5x5 filter), so no zero-insertion FLOPs are spent.
//   wgrad  dW[co,kh,kw,ci] = sum_{n,oh,ow} dy[n,oh,ow,co] * x[n, s*oh+kh-pt, s*ow+kw-pl, ci]
//          K = pixels: both operands are MN-major (channels contiguous), staged as [32 pixel x 32 channel] boxes in the
//          SWIZZLE_128B / 32-byte-atom layout (the layout plan_apply_tc.cu established for MN-major 32-bit operands).
// Padding: TMA zero-fills the part of a box that lies outside the tensor (coordinates are signed), which is exactly
// TensorFlow's 'SAME' zero padding, including the asymmetric stride-2 case (5x5/s2: 1 before, 2 after).  Stride 2 is
// expressed with four parity views of x (base offset (ph*W+pw)*C, dims [C, W/2, H/2, B]) so every tap is a dense box.
//
// Precision: operands are the fp32 tensors read as TF32 by the tensor core (10-bit mantissa, the same math class as the
// cuDNN TF32 kernels this replaces), fp32 accumulation in TMEM.
//
// Kernel shape: warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 2-5 epilogue (TMEM lane quadrant = warp & 3);
// the wgrad kernel adds two more producer warps (its 12 small boxes per stage need three issuing threads).  Persistent CTAs
// (one per SM) walk the tile list; 4-8 smem stages of BK = 32; accumulator 128 x TN fp32 in TMEM, double buffered so the
// epilogue of one tile overlaps the MMAs of the next.
//
// Also in this file: the fused 2x-upsample + convolution forms (same kernels, sub-pixel tap tables, make_submap), the
// 256 x 256-tile variant conv_gemm2_tc_kernel, split-K over filter taps for launches with fewer tiles than SMs, the
// weight-layout helpers (OHWI -> IHWO, sub-filter pre-sum / un-sum), the bias-gradient column sum, and the host-only plan
// capture behind otgan_conv_plan_describe that lets CPU tests replay the exact tap tables a launch would use.
#include "tc_common.cuh"
#include "conv_ex.cuh"
#include <string.h>

namespace otgan {

namespace {

using namespace tc;

constexpr int BK = 32;                        // K-chunk: 32 fp32 = 128-byte rows == swizzle span
constexpr int TM = 128;
constexpr int A_TILE = TM * BK * 4;           // 16 KB
constexpr int BOX32 = 32 * 32 * 4;            // 4 KB: one [32 x 32] fp32 box of the MN-major layout
constexpr int MAX_TAPS = 40;                  // 5x5 = 25; the fused upsample dgrad / wgrad use 4 classes x 9 = 36
constexpr int NUM_EPI_THREADS = 128, NUM_THREADS = 64 + NUM_EPI_THREADS;
constexpr uint32_t SW128 = 2, SW128_BASE32B = 1;
constexpr int SMEM_BUDGET = 200 * 1024;

template <int TN> struct Cfg {
    static constexpr int B_TILE = TN * BK * 4;
    static constexpr int STAGE_BYTES = A_TILE + B_TILE;
    static constexpr int STAGES_ = SMEM_BUDGET / STAGE_BYTES;         // TN = 256: 4 stages of 48 KB; TN = 128: 6 of 32 KB
    static constexpr int STAGES = STAGES_ > 8 ? 8 : STAGES_;          // TN = 16 (18 KB stages): 8 is plenty
    static constexpr int TMEM_COLS = 2 * TN < 32 ? 32 : 2 * TN;
    static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
};

// One filter tap of a launch: which activation view it reads (map) at which pixel shift (dw, dh), the column of the
// weight matrix it multiplies (wcol), a row offset into the weight matrix / filter-gradient output (brow: the fused
// upsample kernels keep one sub-filter per output parity class) and, for wgrad, which dy view it pairs with (dmap).
struct Tap { int map, dw, dh, wcol, brow, dmap; };

// ---------------------------------------------------------------------------------------------- fprop / dgrad
struct GemmParams {
    CUtensorMap amap[4];                      // activation views (stride 2 fprop: the four parity views)
    CUtensorMap bmap;                         // weight matrix [N rows, Ktot] K-major
    Tap taps[MAX_TAPS];
    int cls_tap_begin[5];                     // class c sums taps [begin[c], begin[c+1])
    long long cls_out_off[4];                 // output offset of the class (floats)
    int n_cls, m_tiles, n_tiles, n_items;
    int splits;                               // split-K over the taps of a class (small-M launches); partials split_stride apart
    long long split_stride;
    // tail split: the tiles of the last, partial wave (tiles % 148) are each cut into tail_splits shares of their taps so that
    // the wave fills all SMs; their raw partial tiles go to tail_ws ([tail tile][share][128][TN]) and conv_tail_fixup sums them
    int n_full, tail_splits;
    float* tail_ws;
    int bw, bh, bn, tiles_w, tiles_h;         // pixel box (bw*bh*bn = 128) and tile grid of one class
    int kchunks;                              // 32-channel chunks per tap
    int n_valid;                              // real output channels (TN = 16 variant: N <= 16, rest of the tile is ignored)
    int b_box_bytes;                          // bytes of one weight box (rows actually present in the weight matrix)
    long long osW, osH, osN;                  // output pixel strides (floats)
    float* out;
    const float* bias;                        // [N] or null
    // ---- generic mode (any channel counts / batch, channel slices of wider buffers, fused epilogues): DenseNet and every
    // shape outside the DCGAN family.  The weight operand is a 3-D tensor [K, taps, rows] so that TMA zero-fills both the K
    // tail of a tap and the rows past N; the accumulator N is a run-time multiple of 16 (n_inst <= TN); rows / columns outside
    // the tensor are masked in the epilogue.
    int crelu_half;                           // fused CReLU output of the plain fprop epilogue: 0 = off, else Cout -- row = [relu(y) (Cout) | relu(-y) (Cout)]
    int generic;
    int b_k0;                                 // K coordinate of the first weight column used (a K-slice of a wider weight tensor)
    int n_inst;                               // UMMA N of this launch (= the N tile stride: tile nt covers columns [nt * n_inst, ..))
    int m_w, m_h, m_b;                        // valid extents of the tile grid (row mask)
    int epi_mode;                             // EPI_PLAIN / EPI_CRELU8 / EPI_CRELU8_BWD
    const float* e_add;                       // EPI_CRELU8_BWD: g = acc + e_add (may be null), sign pattern from e_z; both are
    const float* e_z;                         //   [.., 2 * N_out] slices addressed with the strides below
    long long esW, esH, esN;
    int accumulate;                           // EPI_DENSE_FWD: v += previous contents of out (pre-activation accumulator S)
    float* e_out2;                            // EPI_DENSE_FWD: crelu8 of the first 16 columns goes here (strides es*), the rest to out
};

// Epilogues of the generic mode.  CReLU slot layout ("crelu8"): channel c of a tensor with C channels is stored as
// relu(x_c) at (c / 8) * 16 + c % 8 and relu(-x_c) at (c / 8) * 16 + 8 + c % 8 -- blocks of 8 positive then 8 negative parts,
// so that one thread's 32 accumulator columns map to 64 (forward) / 16 (backward) contiguous output floats.

template <int TN>
__device__ __forceinline__ void epilogue_generic(const GemmParams& p, uint32_t tmem_acc, int nt, float* orow, long long epix, bool row_ok)
{
    int lim = (nt + 1) * p.n_inst;                                // first column this tile does NOT own
    lim = lim > p.n_valid ? p.n_valid : lim;
    const int ncols = lim - nt * p.n_inst;
#pragma unroll 1
    for (int cc = 0; cc * 32 < ncols; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_acc + (uint32_t)(cc * 32), v);        // warp-collective: every lane executes it, masked rows included
        tmem_ld_wait();
        if (!row_ok) continue;
        const int col0 = nt * p.n_inst + cc * 32;
        if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < lim) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + col0 + j));
        }
        if (p.epi_mode == EPI_PLAIN) {
            float* o = orow + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (col0 + j + 3 < lim) {
                    *reinterpret_cast<uint4*>(o + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (col0 + j + e < lim) o[j + e] = __uint_as_float(v[j + e]);
                }
            }
        } else if (p.epi_mode == EPI_CRELU8) {
            float* o = orow + 2 * col0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (col0 + 8 * q < lim) {
                    float pos[8], neg[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float x = __uint_as_float(v[8 * q + i]);
                        pos[i] = fmaxf(x, 0.f);
                        neg[i] = fmaxf(-x, 0.f);
                    }
                    *reinterpret_cast<float4*>(o + 16 * q) = make_float4(pos[0], pos[1], pos[2], pos[3]);
                    *reinterpret_cast<float4*>(o + 16 * q + 4) = make_float4(pos[4], pos[5], pos[6], pos[7]);
                    *reinterpret_cast<float4*>(o + 16 * q + 8) = make_float4(neg[0], neg[1], neg[2], neg[3]);
                    *reinterpret_cast<float4*>(o + 16 * q + 12) = make_float4(neg[4], neg[5], neg[6], neg[7]);
                }
            }
        } else if (p.epi_mode == EPI_DENSE_FWD) {
            // DenseNet forward in "contribution" form: this launch adds the contribution of one input slot to the pre-activations
            // of ALL later layers (columns); the first 16 columns belong to the layer that is now complete: they leave as crelu8
            // into its slot of the feature buffer, the others go back to the accumulator S.
            float* o = orow + col0;
            if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (col0 + j < lim) {
                        const float4 a = *reinterpret_cast<const float4*>(o + j);
                        v[j] = __float_as_uint(__uint_as_float(v[j]) + a.x); v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + a.y);
                        v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + a.z); v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + a.w);
                    }
            }
            int jbeg = 0;
            if (col0 == 0) {
                float* z2 = p.e_out2 + epix;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float pos[8], neg[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float x = __uint_as_float(v[8 * q + i]);
                        pos[i] = fmaxf(x, 0.f);
                        neg[i] = fmaxf(-x, 0.f);
                    }
                    *reinterpret_cast<float4*>(z2 + 16 * q) = make_float4(pos[0], pos[1], pos[2], pos[3]);
                    *reinterpret_cast<float4*>(z2 + 16 * q + 4) = make_float4(pos[4], pos[5], pos[6], pos[7]);
                    *reinterpret_cast<float4*>(z2 + 16 * q + 8) = make_float4(neg[0], neg[1], neg[2], neg[3]);
                    *reinterpret_cast<float4*>(z2 + 16 * q + 12) = make_float4(neg[4], neg[5], neg[6], neg[7]);
                }
                jbeg = 16;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j >= jbeg && col0 + j < lim) *reinterpret_cast<uint4*>(o + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {    // EPI_CRELU8_BWD: 32 accumulator columns (gradient of a crelu8 slot) -> 16 gradients of the pre-activation
            float* o = orow + (col0 >> 1);
            const float* z = p.e_z + epix + col0;
            const float* a = p.e_add ? p.e_add + epix + col0 : nullptr;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (col0 + 16 * q < lim) {
                    float d[8];
#pragma unroll
                    for (int i4 = 0; i4 < 8; i4 += 4) {
                        const float4 zp = __ldg(reinterpret_cast<const float4*>(z + 16 * q + i4));
                        const float4 zn = __ldg(reinterpret_cast<const float4*>(z + 16 * q + 8 + i4));
                        float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), an = ap;
                        if (a) {
                            ap = __ldg(reinterpret_cast<const float4*>(a + 16 * q + i4));
                            an = __ldg(reinterpret_cast<const float4*>(a + 16 * q + 8 + i4));
                        }
                        const float gp[4] = {__uint_as_float(v[16 * q + i4]) + ap.x, __uint_as_float(v[16 * q + i4 + 1]) + ap.y,
                                             __uint_as_float(v[16 * q + i4 + 2]) + ap.z, __uint_as_float(v[16 * q + i4 + 3]) + ap.w};
                        const float gn[4] = {__uint_as_float(v[16 * q + 8 + i4]) + an.x, __uint_as_float(v[16 * q + 8 + i4 + 1]) + an.y,
                                             __uint_as_float(v[16 * q + 8 + i4 + 2]) + an.z, __uint_as_float(v[16 * q + 8 + i4 + 3]) + an.w};
                        const float zpv[4] = {zp.x, zp.y, zp.z, zp.w}, znv[4] = {zn.x, zn.y, zn.z, zn.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) d[i4 + e] = (zpv[e] > 0.f ? gp[e] : 0.f) - (znv[e] > 0.f ? gn[e] : 0.f);
                    }
                    *reinterpret_cast<float4*>(o + 8 * q) = make_float4(d[0], d[1], d[2], d[3]);
                    *reinterpret_cast<float4*>(o + 8 * q + 4) = make_float4(d[4], d[5], d[6], d[7]);
                }
            }
        }
    }
}

template <int TN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ GemmParams p)
{
    using C = Cfg<TN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * C::STAGE_BYTES + 8 * (2 * STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), NUM_EPI_THREADS); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // item -> (split, class, N tile, M tile); a split sums a contiguous share of the class's taps.  Items >= n_full are the
    // shares of the tail tiles (sp = share, tail = index of the tail tile, -1 for a whole tile).
    auto decode = [&](int item, int& mt, int& nt, int& cls, int& sp, int& t0, int& t1, int& tail) {
        int nsplit = p.splits;
        tail = -1;
        if (item >= p.n_full) {
            const int q = item - p.n_full;
            tail = q / p.tail_splits;
            sp = q - tail * p.tail_splits;
            item = p.n_full + tail;
            nsplit = p.tail_splits;
        }
        mt = item % p.m_tiles; item /= p.m_tiles;
        nt = item % p.n_tiles; item /= p.n_tiles;
        cls = item % p.n_cls;
        if (tail < 0) sp = item / p.n_cls;
        const int tb = p.cls_tap_begin[cls], T = p.cls_tap_begin[cls + 1] - tb;
        t0 = tb + (T * sp) / nsplit;
        t1 = tb + (T * (sp + 1)) / nsplit;
    };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int c = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int mt, nt, cls, sp, t0, t1, tail;
                decode(item, mt, nt, cls, sp, t0, t1, tail);
                const int w0 = (mt % p.tiles_w) * p.bw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
                const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
                for (int t = t0; t < t1; ++t) {
                    const Tap tap = p.taps[t];
                    const CUtensorMap* am = &p.amap[tap.map];
                    for (int kc = 0; kc < p.kchunks; ++kc, ++c) {
                        const int s = c % STAGES;
                        mbar_wait(empty_bar(s), (((uint32_t)(c / STAGES)) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(full_bar(s), (uint32_t)(A_TILE + p.b_box_bytes));
                        const uint32_t dst = smem_base + s * C::STAGE_BYTES;
                        tma_load_4d(dst, am, full_bar(s), kc * BK, w0 + tap.dw, h0 + tap.dh, n0);
                        if (p.generic) tma_load_3d(dst + A_TILE, &p.bmap, full_bar(s), p.b_k0 + kc * BK, tap.dmap, tap.brow + nt * p.n_inst);
                        else tma_load_2d(dst + A_TILE, &p.bmap, full_bar(s), tap.wcol + kc * BK, tap.brow + nt * TN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TM, p.generic ? p.n_inst : TN);
            int c = 0, n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                int mt, nt, cls, sp, t0, t1, tail;
                decode(item, mt, nt, cls, sp, t0, t1, tail);
                const int b = n & 1;
                mbar_wait(tempty_bar(b), (((uint32_t)(n >> 1)) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(b * TN);
                const int nst = (t1 - t0) * p.kchunks;
                for (int st = 0; st < nst; ++st, ++c) {
                    const int s = c % STAGES;
                    mbar_wait(full_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_base + s * C::STAGE_BYTES, b0 = a0 + A_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {          // tf32 MMA K = 8 elements = 32 bytes inside the swizzled row
                        const uint64_t ad = umma_desc_kmajor(a0 + k * 32, 1024, SW128);
                        const uint64_t bd = umma_desc_kmajor(b0 + k * 32, 1024, SW128);
                        umma_tf32(d, ad, bd, idesc, (st > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(b));
            }
        }
    } else {
        // ===================================================== epilogue warps: TMEM -> (+bias) -> global
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                          // tile row = pixel index inside the box (w fastest)
        const int rw = r % p.bw, rh = (r / p.bw) % p.bh, rn = r / (p.bw * p.bh);
        int n = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
            int mt, nt, cls, sp, t0, t1, tail;
            decode(item, mt, nt, cls, sp, t0, t1, tail);
            const int b = n & 1;
            const int w0 = (mt % p.tiles_w) * p.bw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
            const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
            float* out = p.out + sp * p.split_stride + p.cls_out_off[cls] + (long long)(n0 + rn) * p.osN +
                         (long long)(h0 + rh) * p.osH + (long long)(w0 + rw) * p.osW + nt * TN;
            const float* bias = (p.bias && sp == 0) ? p.bias + nt * TN : nullptr;     // the first partial carries the bias
            if (tail >= 0) {                             // share of a tail tile: raw partial, compact [128][TN], bias in the fix-up
                out = p.tail_ws + ((size_t)tail * p.tail_splits + sp) * (size_t)(TM * TN) + (size_t)r * TN;
                bias = nullptr;
            }
            if (p.generic) {
                if constexpr (TN != 16) {
                    const bool row_ok = (w0 + rw < p.m_w) && (h0 + rh < p.m_h) && (n0 + rn < p.m_b);
                    const long long epix = (long long)(n0 + rn) * p.esN + (long long)(h0 + rh) * p.esH + (long long)(w0 + rw) * p.esW;
                    float* orow = p.out + p.cls_out_off[cls] + (long long)(n0 + rn) * p.osN + (long long)(h0 + rh) * p.osH +
                                  (long long)(w0 + rw) * p.osW;
                    // The read-modify-write epilogues (dense-block forward accumulator, crelu8 backward inputs) are latency-bound on
                    // short tiles: pull this row's operands towards L2 / L1 while the tile's MMAs are still running.
                    if (row_ok) {
                        int lim = (nt + 1) * p.n_inst;
                        lim = lim > p.n_valid ? p.n_valid : lim;
                        if (p.epi_mode == EPI_DENSE_FWD && p.accumulate) {
                            for (int c = nt * p.n_inst; c < lim; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(orow + c));
                        } else if (p.epi_mode == EPI_CRELU8_BWD) {
                            for (int c = nt * p.n_inst; c < lim; c += 32) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e_z + epix + c));
                                if (p.e_add) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e_add + epix + c));
                            }
                        }
                    }
                    mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
                    tcgen05_fence_after();
                    epilogue_generic<TN>(p, tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN), nt, orow, epix, row_ok);
                } else {
                    mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
                    tcgen05_fence_after();
                }
            } else if constexpr (TN == 16) {
                mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
                tcgen05_fence_after();
                // narrow outputs (the generator's 3-channel image, the critic's image gradient): only n_valid columns exist;
                // the weight box has n_valid rows, the other accumulator columns hold garbage and are never stored
                uint32_t v[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < p.n_valid) out[j] = __uint_as_float(v[j]) + (bias ? __ldg(bias + j) : 0.f);
            } else {
            mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < TN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + cc * 32), v);
                tmem_ld_wait();
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + j));
                        v[j] = __float_as_uint(__uint_as_float(v[j]) + bb.x);
                        v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + bb.y);
                        v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + bb.z);
                        v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + bb.w);
                    }
                }
                if (p.crelu_half) {
                    // CReLU of the next layer (utils/nn.py:198-200 on one tensor: concat([y, -y], 3) then relu) leaves this epilogue
                    // directly: relu(y) into channels [0, Cout), relu(-y) into [Cout, 2 Cout) of a 2 Cout-wide row
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]), a2 = __uint_as_float(v[j + 2]), a3 = __uint_as_float(v[j + 3]);
                        *reinterpret_cast<float4*>(out + cc * 32 + j) = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
                        *reinterpret_cast<float4*>(out + p.crelu_half + cc * 32 + j) = make_float4(fmaxf(-a0, 0.f), fmaxf(-a1, 0.f), fmaxf(-a2, 0.f), fmaxf(-a3, 0.f));
                    }
                } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(out + cc * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
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
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------- fprop / dgrad, 256 x 256 tiles
// Measured on the kernel above (ncu, fused-upsample fprop of generator conv2d_0): every SM ingests 81 B/clk of operand tiles
// from L2 while it is active and the tensor pipe is busy exactly 81/96 = 84.5% of that time -- 96 B/clk (48 KB per 512-cycle
// K-chunk) is what a 128 x 256 tile needs at the full MMA rate, so the kernel is bound by operand delivery, not by the tensor
// pipe or by wave quantisation (cutting the last wave into shares changed nothing).  This variant gives each CTA TWO 128-pixel
// sub-tiles that share one 256-row weight tile: 64 KB per 1024 cycles of MMA = 64 B/clk.  Both accumulators (2 x 256 columns)
// fill TMEM, so the epilogue no longer overlaps the next tile's MMAs.  Measured: +12% on launches whose tiles have >= 500
// K-chunks, -10% on short tiles -- the convolutions run power-capped (~1.7 GHz), so less data movement only pays where the
// un-overlapped epilogue is negligible.  Used for tiles of >= 500 K-chunks in launches with >= 0.75 waves of the bigger tiles.
constexpr int G2_TN = 256;
constexpr int G2_B_TILE = G2_TN * BK * 4;                     // 32 KB
constexpr int G2_STAGE_BYTES = 2 * A_TILE + G2_B_TILE;        // 64 KB
constexpr int G2_STAGES = 3;
constexpr size_t G2_SMEM_BYTES = 1024 + (size_t)G2_STAGES * G2_STAGE_BYTES + 256;

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm2_tc_kernel(const __grid_constant__ GemmParams p)
{
    constexpr int STAGES = G2_STAGES, TN = G2_TN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES * G2_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bars + 8u * (2 * STAGES), tempty_bar = bars + 8u * (2 * STAGES + 1);
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 2);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * G2_STAGE_BYTES + 8 * (2 * STAGES + 2));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, NUM_EPI_THREADS);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // item -> (split, class, N tile, pair of M tiles): sub-tile j of pair m2 is the 128-pixel tile 2 m2 + j
    const int m_pairs = p.m_tiles >> 1;
    auto decode = [&](int item, int& m2, int& nt, int& cls, int& sp, int& t0, int& t1) {
        m2 = item % m_pairs; item /= m_pairs;
        nt = item % p.n_tiles; item /= p.n_tiles;
        cls = item % p.n_cls;
        sp = item / p.n_cls;
        const int tb = p.cls_tap_begin[cls], T = p.cls_tap_begin[cls + 1] - tb;
        t0 = tb + (T * sp) / p.splits;
        t1 = tb + (T * (sp + 1)) / p.splits;
    };
    auto tile_origin = [&](int mt, int& w0, int& h0, int& n0) {
        w0 = (mt % p.tiles_w) * p.bw;
        h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
    };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int m2, nt, cls, sp, t0, t1, wa, ha, na, wb, hb, nb;
                decode(item, m2, nt, cls, sp, t0, t1);
                tile_origin(2 * m2, wa, ha, na);
                tile_origin(2 * m2 + 1, wb, hb, nb);
                for (int t = t0; t < t1; ++t) {
                    const Tap tap = p.taps[t];
                    const CUtensorMap* am = &p.amap[tap.map];
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        mbar_arrive_expect_tx(full_bar(s), (uint32_t)G2_STAGE_BYTES);
                        const uint32_t dst = smem_base + s * G2_STAGE_BYTES;
                        tma_load_4d(dst, am, full_bar(s), kc * BK, wa + tap.dw, ha + tap.dh, na);
                        tma_load_4d(dst + A_TILE, am, full_bar(s), kc * BK, wb + tap.dw, hb + tap.dh, nb);
                        tma_load_2d(dst + 2 * A_TILE, &p.bmap, full_bar(s), tap.wcol + kc * BK, tap.brow + nt * TN);
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TM, TN);
            int s = 0, n = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                int m2, nt, cls, sp, t0, t1;
                decode(item, m2, nt, cls, sp, t0, t1);
                mbar_wait(tempty_bar, ((uint32_t)n & 1u) ^ 1u);          // both accumulators drained by the epilogue
                tcgen05_fence_after();
                const int nst = (t1 - t0) * p.kchunks;
                for (int st = 0; st < nst; ++st) {
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_base + s * G2_STAGE_BYTES, a1 = a0 + A_TILE, b0 = a0 + 2 * A_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t bd = umma_desc_kmajor(b0 + k * 32, 1024, SW128);
                        umma_tf32(tmem_base, umma_desc_kmajor(a0 + k * 32, 1024, SW128), bd, idesc, (st > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_base + TN, umma_desc_kmajor(a1 + k * 32, 1024, SW128), bd, idesc, (st > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull_bar);
            }
        }
    } else {
        // ===================================================== epilogue warps: both accumulators -> (+bias) -> global
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const int rw = r % p.bw, rh = (r / p.bw) % p.bh, rn = r / (p.bw * p.bh);
        int n = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
            int m2, nt, cls, sp, t0, t1;
            decode(item, m2, nt, cls, sp, t0, t1);
            const float* bias = (p.bias && sp == 0) ? p.bias + nt * TN : nullptr;
            mbar_wait(tfull_bar, (uint32_t)n & 1u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                int w0, h0, n0;
                tile_origin(2 * m2 + j, w0, h0, n0);
                float* out = p.out + sp * p.split_stride + p.cls_out_off[cls] + (long long)(n0 + rn) * p.osN +
                             (long long)(h0 + rh) * p.osH + (long long)(w0 + rw) * p.osW + nt * TN;
#pragma unroll 1
                for (int cc = 0; cc < TN / 32; ++cc) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * TN + cc * 32), v);
                    tmem_ld_wait();
                    if (bias) {
#pragma unroll
                        for (int q = 0; q < 32; q += 4) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + q));
                            v[q] = __float_as_uint(__uint_as_float(v[q]) + bb.x);
                            v[q + 1] = __float_as_uint(__uint_as_float(v[q + 1]) + bb.y);
                            v[q + 2] = __float_as_uint(__uint_as_float(v[q + 2]) + bb.z);
                            v[q + 3] = __float_as_uint(__uint_as_float(v[q + 3]) + bb.w);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 32; q += 4)
                        *reinterpret_cast<uint4*>(out + cc * 32 + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tempty_bar);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Sums the shares of every tail tile (fixed order), adds the bias and scatters the rows to their output pixels.
template <int TN>
__global__ void __launch_bounds__(256)
conv_tail_fixup_kernel(const __grid_constant__ GemmParams p)
{
    const int tail = blockIdx.x;
    int item = p.n_full + tail;
    const int mt = item % p.m_tiles; item /= p.m_tiles;
    const int nt = item % p.n_tiles; item /= p.n_tiles;
    const int cls = item % p.n_cls;
    const int w0 = (mt % p.tiles_w) * p.bw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
    const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
    const float4* src = reinterpret_cast<const float4*>(p.tail_ws + (size_t)tail * p.tail_splits * (size_t)(TM * TN));
    constexpr int C4 = TN / 4;
    for (int idx = threadIdx.x; idx < TM * C4; idx += blockDim.x) {
        const int r = idx / C4, c4 = idx - r * C4;
        float4 a = src[idx];
        for (int s = 1; s < p.tail_splits; ++s) {
            const float4 b = src[(size_t)s * (TM * C4) + idx];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (p.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + nt * TN) + c4);
            a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
        }
        const int rw = r % p.bw, rh = (r / p.bw) % p.bh, rn = r / (p.bw * p.bh);
        float* out = p.out + p.cls_out_off[cls] + (long long)(n0 + rn) * p.osN + (long long)(h0 + rh) * p.osH +
                     (long long)(w0 + rw) * p.osW + nt * TN;
        *reinterpret_cast<float4*>(out + 4 * c4) = a;
    }
}

// ---------------------------------------------------------------------------------------------- wgrad
struct WgradParams {
    CUtensorMap dymap[4];                     // dy views (parity views for the fused upsample), box [32 co] x [bw x bh x bn = 32 px]
    CUtensorMap xmap[4];                      // x views (parity views for stride 2), box [32 ci] x [bw x bh x bn = 32 px]
    Tap taps[MAX_TAPS];                       // map, dw, dh, wcol = tap * Cin
    int ntaps, co_tiles, ci_tiles, splits, n_items;
    int bw, bh, bn, tiles_w, tiles_h;         // 32-pixel chunk box and the chunk grid over the OUTPUT pixels
    int nchunks, chunks_per_split;
    int ldw;                                  // row stride of dW (= ntaps * Cin)
    long long split_stride;                   // floats between split partials
    float* out;
    int masked, m_valid, n_valid;             // generic shapes: rows (co) >= m_valid / columns (ci) >= n_valid are not stored
    int hwio, hwio_c, hwio_rows;              // hwio != 0: write dW as [class][tap * Cin + ci][co] (the layout of the HWIO variable V,
                                              // hwio_c = Cout, hwio_rows = Cout rows per class of `brow`): coalesced across the 32 lanes
};

// wgrad thread layout: warp 0 = producer of the dy boxes (+ expect_tx), warp 1 = MMA issuer, warps 2-5 = epilogue,
// warps 6-7 = producers of the x boxes.  Measured with one producer thread: 12 small TMA copies per stage cost ~1200 issue
// cycles (address arithmetic + the uniform-register hand-off around every UTMALDG) against 512 cycles of MMA, tensor pipe
// 48% busy; three producer threads with incremental coordinates (no division in the loop) keep the pipe fed.
constexpr int WGRAD_THREADS = 256;

template <int TN>
__global__ void __launch_bounds__(WGRAD_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgradParams p)
{
    using C = Cfg<TN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * C::STAGE_BYTES + 8 * (2 * STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) { tma_prefetch_desc(&p.dymap[i]); tma_prefetch_desc(&p.xmap[i]); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), NUM_EPI_THREADS); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // item -> (split, tap, ci tile, co tile); co fastest so neighbouring CTAs share the x boxes, all CTAs walk the
    // pixel range in step so dy / x are read from HBM once and then hit in L2
    auto decode = [&](int item, int& cot, int& cit, int& t, int& sp) {
        cot = item % p.co_tiles; item /= p.co_tiles;
        cit = item % p.ci_tiles; item /= p.ci_tiles;
        t = item % p.ntaps;
        sp = item / p.ntaps;
    };
    auto chunk_range = [&](int sp, int& c0, int& c1) {
        c0 = sp * p.chunks_per_split;
        c1 = c0 + p.chunks_per_split;
        c1 = c1 > p.nchunks ? p.nchunks : c1;
    };

    if (warp == 0 || warp >= 6) {
        // ===================================================== TMA producers: warp 0 -> the TM/32 dy boxes, warps 6, 7 -> half of
        // the TN/32 x boxes each.  All three walk the same (item, chunk) sequence and wait on the same empty barrier.
        if (lane == 0) {
            constexpr int XB = TN / 64;                       // x boxes per x-producer warp
            const int role = warp == 0 ? 0 : warp - 5;        // 0: dy, 1 / 2: x halves
            int s = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int cot, cit, t, sp, c0, c1;
                decode(item, cot, cit, t, sp);
                chunk_range(sp, c0, c1);
                const Tap tap = p.taps[t];
                const CUtensorMap* xm = &p.xmap[tap.map];
                const CUtensorMap* dm = &p.dymap[tap.dmap];
                int tw = c0 % p.tiles_w, th = (c0 / p.tiles_w) % p.tiles_h, tn = c0 / (p.tiles_w * p.tiles_h);
                const int co0 = cot * TM, ci0 = cit * TN + (role == 2 ? 32 * XB : 0);
                for (int ch = c0; ch < c1; ++ch) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t dst = smem_base + s * C::STAGE_BYTES;
                    if (role == 0) {
                        mbar_arrive_expect_tx(full_bar(s), (uint32_t)C::STAGE_BYTES);
#pragma unroll
                        for (int j = 0; j < TM / 32; ++j)
                            tma_load_4d(dst + j * BOX32, dm, full_bar(s), co0 + 32 * j, tw * p.bw, th * p.bh, tn * p.bn);
                    } else {
                        const uint32_t dstx = dst + A_TILE + (role == 2 ? XB * BOX32 : 0);
                        const int cw = tw * p.bw + tap.dw, chh = th * p.bh + tap.dh, cn = tn * p.bn;
#pragma unroll
                        for (int j = 0; j < XB; ++j)
                            tma_load_4d(dstx + j * BOX32, xm, full_bar(s), ci0 + 32 * j, cw, chh, cn);
                    }
                    if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TM, TN, /*a_mn_major=*/1, /*b_mn_major=*/1);
            int c = 0, n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                int cot, cit, t, sp, c0, c1;
                decode(item, cot, cit, t, sp);
                chunk_range(sp, c0, c1);
                const int b = n & 1;
                mbar_wait(tempty_bar(b), (((uint32_t)(n >> 1)) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(b * TN);
                for (int ch = c0; ch < c1; ++ch, ++c) {
                    const int s = c % STAGES;
                    mbar_wait(full_bar(s), ((uint32_t)(c / STAGES)) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_base + s * C::STAGE_BYTES, b0 = a0 + A_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        // MN-major, SWIZZLE_128B with 32-byte atoms: 32-channel blocks LBO = 4096 B apart, 4-pixel K groups
                        // SBO = 512 B apart; one K = 8 MMA spans two groups; K-step = 8 pixel rows = +1024 B
                        const uint64_t ad = umma_desc_mnmajor(a0 + k * 1024, BOX32, 512, SW128_BASE32B);
                        const uint64_t bd = umma_desc_mnmajor(b0 + k * 1024, BOX32, 512, SW128_BASE32B);
                        umma_tf32(d, ad, bd, idesc, (ch > c0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(b));
            }
        }
    } else {
        // ===================================================== epilogue warps: TMEM -> dW (or the split partial)
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                          // output channel inside the tile
        int n = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
            int cot, cit, t, sp;
            decode(item, cot, cit, t, sp);
            const int b = n & 1;
            float* out = p.out + (long long)sp * p.split_stride + (long long)(p.taps[t].brow + cot * TM + r) * p.ldw + p.taps[t].wcol + cit * TN;
            mbar_wait(tfull_bar(b), ((uint32_t)(n >> 1)) & 1u);
            tcgen05_fence_after();
            const bool row_ok = !p.masked || (cot * TM + r < p.m_valid);
            float* out_t = nullptr;
            if (p.hwio) {          // [class][tap * Cin + ci][co]: lane r -> consecutive co, one 128-byte segment per ci column
                const int cls = p.taps[t].brow / p.hwio_rows;
                out_t = p.out + (long long)sp * p.split_stride + ((long long)cls * p.ldw + p.taps[t].wcol + cit * TN) * p.hwio_c + cot * TM + r;
            }
#pragma unroll 1
            for (int cc = 0; cc < TN / 32; ++cc) {
                if (p.masked && cit * TN + cc * 32 >= p.n_valid) break;          // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + cc * 32), v);
                tmem_ld_wait();
                if (!row_ok) continue;
                if (p.hwio) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (!p.masked || cit * TN + cc * 32 + j < p.n_valid) out_t[(long long)(cc * 32 + j) * p.hwio_c] = __uint_as_float(v[j]);
                } else if (!p.masked || cit * TN + cc * 32 + 32 <= p.n_valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(out + cc * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (cit * TN + cc * 32 + j < p.n_valid) out[cc * 32 + j] = __uint_as_float(v[j]);
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
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// out[i] = sum_s partial[s][i]  (fixed order: deterministic split-K)
__global__ void __launch_bounds__(256)
split_reduce_kernel(size_t n4, int S, size_t stride4, const float4* __restrict__ partial, float4* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = partial[i];
        for (int s = 1; s < S; ++s) {
            const float4 b = partial[i + s * stride4];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}

// Wt[ci][t][co] = W[co][t][ci]   (OHWI -> IHWO, the weight matrix of dgrad); one (co, ci) 32x32 tile per block, per tap
__global__ void __launch_bounds__(256)
ohwi_to_ihwo_kernel(int Cout, int T, int Cin, const float* __restrict__ w, float* __restrict__ wt)
{
    __shared__ float tile[32][33];
    const int t = blockIdx.z, ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int co = co0 + i, ci = ci0 + tx;
        tile[i][tx] = (co < Cout && ci < Cin) ? w[((size_t)co * T + t) * Cin + ci] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int ci = ci0 + i, co = co0 + tx;
        if (ci < Cin && co < Cout) wt[((size_t)ci * T + t) * Cout + co] = tile[tx][i];
    }
}

// column sums of a [P, C] matrix (the bias gradient, tf.nn.bias_add backward): partial[slab][C] then a fixed-order reduce
__global__ void __launch_bounds__(256)
colsum_partial_kernel(int P, int C, int rows_per_slab, const float* __restrict__ x, float* __restrict__ partial)
{
    __shared__ float4 red[8][32];
    const int c4 = blockIdx.x * 32 + (threadIdx.x & 31);         // float4 column
    const int ry = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_slab;
    int r1 = r0 + rows_per_slab;
    r1 = r1 > P ? P : r1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 * 4 < C)
        for (int r = r0 + ry; r < r1; r += 8) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * C + c4 * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    red[ry][threadIdx.x & 31] = a;
    __syncthreads();
    if (ry == 0 && c4 * 4 < C) {
        for (int i = 1; i < 8; ++i) {
            const float4 v = red[i][threadIdx.x];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * C + c4 * 4) = a;
    }
}

// ---- fused 2x nearest-neighbour upsample + convolution (generator: resize_nearest_neighbor -> conv2d, models/dcgan.py:37-46)
// Output pixel (2i+a, 2j+b) of the convolution over the upsampled image reads x_low rows i + floor((a + kh - pad) / 2): the kh
// taps of a 5x5 filter collapse onto 3 low-resolution rows (3x3 filters: 2), so each of the 4 output parity classes (a, b) is a
// 3x3 convolution of the LOW-resolution input with a pre-summed sub-filter -- 9 taps instead of 25 (2.8x fewer FLOPs), and the
// upsampled tensor is never materialised.  SubMap holds the tap -> sub-row tables for both axes.
struct SubMap {
    int kh, kw, n1h, n1w;
    int dmin_h[2], dmin_w[2];          // first low-res offset of parity 0 / 1
    int idx_h[2][8], idx_w[2][8];      // idx[a][k] = floor((a + k - pad) / 2) - dmin[a]  in [0, n1)
    // idx is non-decreasing in k, so the taps that fold onto one sub-row / sub-column are CONSECUTIVE: [lo, hi)
    unsigned char h_lo[2][8], h_hi[2][8], w_lo[2][8], w_hi[2][8];
};

// w_sub[cls = 2a+b][co][(ri * n1w + ci) * Cin + c] = sum_{kh: idx_h[a][kh] == ri} sum_{kw: idx_w[b][kw] == ci} w[co][(kh * KW + kw) * Cin + c]
// One CTA per (filter co, parity class): it walks the class's slots and, per slot, only the 1-4 taps that fold onto it; threads =
// float4s of input channels.  The four classes of a filter are neighbouring CTAs, so the filter row (kh*kw*Cin floats) comes from
// HBM once and from L2 three times.  (Round 1 ran one flat index over all outputs: four 64-bit div/mods and a scan over all kh x kw
// taps per float4 made it instruction-bound at 2.5 TB/s.)
__global__ void __launch_bounds__(256)
up2_presum_kernel(SubMap m, int Cout, int Cin, const float* __restrict__ w, float* __restrict__ w_sub)
{
    const int slots = m.n1h * m.n1w, C4 = Cin >> 2;
    const int co = blockIdx.x >> 2, cls = blockIdx.x & 3, a = cls >> 1, b = cls & 1;
    const float4* wrow = reinterpret_cast<const float4*>(w + (size_t)co * m.kh * m.kw * Cin);
    float4* orow = reinterpret_cast<float4*>(w_sub) + ((size_t)cls * Cout + co) * slots * C4;
    for (int slot = 0; slot < slots; ++slot) {
        const int ri = slot / m.n1w, ci = slot - ri * m.n1w;
        const int h0 = m.h_lo[a][ri], h1 = m.h_hi[a][ri], w0 = m.w_lo[b][ci], w1 = m.w_hi[b][ci];
        for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int kh = h0; kh < h1; ++kh)
                for (int kw = w0; kw < w1; ++kw) {
                    const float4 v = __ldg(wrow + (size_t)(kh * m.kw + kw) * C4 + c4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            orow[(size_t)slot * C4 + c4] = acc;
        }
    }
}

// dw[co][(kh * KW + kw) * Cin + c] = sum_{a, b} dw_sub[2a+b][co][(idx_h[a][kh] * n1w + idx_w[b][kw]) * Cin + c]   (chain rule of the presum)
__global__ void __launch_bounds__(256)
up2_unsum_kernel(SubMap m, int Cout, int Cin, const float* __restrict__ dw_sub, float* __restrict__ dw)
{
    const int slots = m.n1h * m.n1w, taps = m.kh * m.kw, C4 = Cin >> 2;
    const size_t total = (size_t)Cout * taps * C4;
    const float4* src = reinterpret_cast<const float4*>(dw_sub);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int t = (int)(r % taps);
        const int co = (int)(r / taps);
        const int kh = t / m.kw, kw = t % m.kw;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int slot = m.idx_h[a][kh] * m.n1w + m.idx_w[b][kw];
                const float4 v = __ldg(src + (((size_t)(2 * a + b) * Cout + co) * slots + slot) * C4 + c4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        reinterpret_cast<float4*>(dw)[i] = acc;
    }
}

// The same two maps in the layouts whose contiguous axis is the OUTPUT channel (float4 along co):
//   w_sub_t[cls][ci][slot][co] = sum_{taps t in (cls, slot)} w_ihwo[ci][t][co]      (the dgrad operand of the fused-upsample layers,
//                                                                                  built from the IHWO filter in one pass)
// One CTA per (input channel ci, parity class), like up2_presum_kernel.
__global__ void __launch_bounds__(256)
up2_presum_ihwo_kernel(SubMap m, int Cout, int Cin, const float* __restrict__ w_ihwo, float* __restrict__ w_sub_t)
{
    const int slots = m.n1h * m.n1w, C4 = Cout >> 2, taps = m.kh * m.kw;
    const int ci = blockIdx.x >> 2, cls = blockIdx.x & 3, a = cls >> 1, b = cls & 1;
    const float4* src = reinterpret_cast<const float4*>(w_ihwo + (size_t)ci * taps * Cout);
    float4* orow = reinterpret_cast<float4*>(w_sub_t) + ((size_t)cls * Cin + ci) * slots * C4;
    for (int slot = 0; slot < slots; ++slot) {
        const int ri = slot / m.n1w, cj = slot - ri * m.n1w;
        const int h0 = m.h_lo[a][ri], h1 = m.h_hi[a][ri], w0 = m.w_lo[b][cj], w1 = m.w_hi[b][cj];
        for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int kh = h0; kh < h1; ++kh)
                for (int kw = w0; kw < w1; ++kw) {
                    const float4 v = __ldg(src + (size_t)(kh * m.kw + kw) * C4 + c4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            orow[(size_t)slot * C4 + c4] = acc;
        }
    }
}
//   dw_hwio[t][ci][co] = sum_{a, b} dw_sub_hwio[2a+b][slot(a, b, t)][ci][co]          (chain rule of the pre-sum on HWIO gradients)
// One CTA per (tap, input channel) row of the gradient; threads = float4s of output channels.
__global__ void __launch_bounds__(256)
up2_unsum_hwio_kernel(SubMap m, int Cout, int Cin, const float* __restrict__ dw_sub, float* __restrict__ dw)
{
    const int slots = m.n1h * m.n1w, C4 = Cout >> 2;
    const int t = blockIdx.x / Cin, ci = blockIdx.x - t * Cin;
    const int kh = t / m.kw, kw = t - kh * m.kw;
    const float4* src = reinterpret_cast<const float4*>(dw_sub);
    const float4* rows[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int slot = m.idx_h[a][kh] * m.n1w + m.idx_w[b][kw];
            rows[2 * a + b] = src + (((size_t)(2 * a + b) * slots + slot) * Cin + ci) * C4;
        }
    float4* orow = reinterpret_cast<float4*>(dw) + (size_t)blockIdx.x * C4;
    for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(rows[q] + c4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        orow[c4] = acc;
    }
}

// out[c] = sum_s partial[s][c]: 8 float4 columns x 32 slab lanes per block, fixed-order tree (deterministic)
__global__ void __launch_bounds__(256)
colsum_final_kernel(int C4, int S, const float4* __restrict__ partial, float4* __restrict__ out)
{
    __shared__ float4 red[32][8];
    const int cx = threadIdx.x & 7, sy = threadIdx.x >> 3, col = blockIdx.x * 8 + cx;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < C4)
        for (int s = sy; s < S; s += 32) {
            const float4 v = partial[(size_t)s * C4 + col];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    red[sy][cx] = acc;
    __syncthreads();
    if (sy == 0 && col < C4) {
        for (int i = 1; i < 32; ++i) {
            const float4 v = red[i][cx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[col] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- host helpers
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// pixel box with `npix` pixels over a [Wt, Ht, Bt] grid (w fastest): bw*bh*bn == npix, each dividing its extent
bool pixel_box(int npix, int Wt, int Ht, int Bt, int* bw, int* bh, int* bn)
{
    if (!is_pow2(Wt) || !is_pow2(Ht)) return false;
    *bw = Wt < npix ? Wt : npix;
    const int rem = npix / *bw;
    *bh = Ht < rem ? Ht : rem;
    *bn = rem / *bh;
    return (*bw) * (*bh) * (*bn) == npix && Bt % *bn == 0 && *bn <= 256;
}

// ---- plan capture (otgan_conv_plan_describe): the launch functions run unchanged up to the point of the kernel launch, with
// tensor-map encoding skipped (no driver needed), and hand their kernel parameters over instead of launching.  Lets the
// CPU tests check the very tap tables / classes / strides / splits the GPU kernels would receive.
struct PlanCapture {
    bool active = false;
    int kind = -1;            // 0: conv_gemm_tc_kernel, 1: conv_gemm2_tc_kernel, 2: conv_wgrad_tc_kernel
    int TN = 0, n_tail = 0;
    GemmParams gemm;
    WgradParams wgrad;
};
thread_local PlanCapture* t_capture = nullptr;
inline bool capturing() { return t_capture && t_capture->active; }

bool conv_map_2d(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, int box_cols, CUtensorMapSwizzle swz)
{
    if (capturing()) return true;
    return make_tensor_map_2d(map, base, rows, cols, ld, box_rows, box_cols, swz);
}

// 4-D view of an NHWC tensor [B, H, W, C] sub-sampled by `s` starting at pixel (ph, pw): dims [C, W/s, H/s, B]
// `ld` = pixel stride in floats (0: the tensor is dense, ld = C); with ld > C the view is a channel SLICE of a wider buffer
// (DenseNet's concatenated feature buffer): channels past C are outside the tensor and zero-filled like the spatial padding.
bool make_view_map(CUtensorMap* map, const float* base, int B, int H, int W, int C, int s, int ph, int pw,
                   const unsigned box[4], CUtensorMapSwizzle swz, int ld = 0)
{
    if (capturing()) return true;
    const unsigned long long L = (unsigned long long)(ld > 0 ? ld : C);
    const unsigned long long dims[4] = {(unsigned long long)C, (unsigned long long)(W / s), (unsigned long long)(H / s), (unsigned long long)B};
    const unsigned long long str[3] = {(unsigned long long)s * L * 4, (unsigned long long)s * W * L * 4, (unsigned long long)H * W * L * 4};
    return make_tensor_map_nd(map, base + ((size_t)ph * W + pw) * L, 4, dims, str, box, swz);
}

// 3-D weight tensor of the generic mode: [K (innermost), taps, rows], row stride ldrow floats, tap stride ldtap floats
bool make_weight_map_3d(CUtensorMap* map, const float* base, int K, int taps, int rows, long long ldtap, long long ldrow, int box_rows)
{
    if (capturing()) return true;
    const unsigned long long dims[3] = {(unsigned long long)K, (unsigned long long)taps, (unsigned long long)rows};
    const unsigned long long str[2] = {(unsigned long long)ldtap * 4, (unsigned long long)ldrow * 4};
    const unsigned box[3] = {(unsigned)BK, 1u, (unsigned)box_rows};
    return make_tensor_map_nd(map, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// generic pixel box: like pixel_box, but the batch need not be a multiple of bn (the overhang is zero-filled / masked)
bool pixel_box_generic(int npix, int Wt, int Ht, int* bw, int* bh, int* bn)
{
    if (!is_pow2(Wt) || !is_pow2(Ht)) return false;
    *bw = Wt < npix ? Wt : npix;
    const int rem = npix / *bw;
    *bh = Ht < rem ? Ht : rem;
    *bn = rem / *bh;
    return (*bw) * (*bh) * (*bn) == npix && *bn <= 256;
}

// Split-K of fprop / dgrad: when a launch has fewer tiles than SMs (small batch per GPU), the taps of each class are
// shared out over up to 148 / tiles CTAs that write partial outputs, summed in fixed order afterwards.
int gemm_splits(int tiles, int min_taps)
{
    if (tiles >= (kNumSMs * 3) / 4) return 1;
    int S = kNumSMs / tiles;
    S = S > min_taps ? min_taps : S;
    S = S > 8 ? 8 : S;
    return S < 1 ? 1 : S;
}

// Shares per tail tile: r = tiles % 148 tiles are left for the last wave; cutting each into S shares of its taps makes the
// wave last ceil(r S / 148) / S of a tile time instead of 1.  Pick the S (<= 8, <= taps) with the shortest tail, preferring
// small S (each share costs a partial tile of workspace traffic); 1 = leave the tail alone.
int tail_splits_for(int r, int min_taps)
{
    if (r <= 0) return 1;
    int best = 1;
    double best_t = 0.9;                                 // at least a 10% shorter tail wave, else leave it alone
    const int smax = min_taps < 8 ? min_taps : 8;
    for (int S = 2; S <= smax; ++S) {
        const double t = (double)((r * S + kNumSMs - 1) / kNumSMs) / S + 0.02 * S;
        if (t < best_t - 1e-9) { best_t = t; best = S; }
    }
    return best;
}

// A/B switches (otgan_conv_set_option).  Both experiments were measured on B200 with the kernels power-capped (sw_power_cap,
// SM clock ~1.7 GHz, 700-780 TFLOP/s TF32):
//  * 256 x 256 tiles: +12% on launches with very long tiles (critic conv2d_3 fprop, fused-upsample dgrad of generator
//    conv2d_1), -10% on short ones (un-overlapped epilogue) -> used only for tiles of >= 500 K-chunks;
//  * tail split: no measurable change (the chip, not the per-SM schedule, limits throughput) -> off by default.
bool g_use_gemm2 = true;
bool g_tail_split = false;
int g_gemm2_min_chunks = 500;

// finish a fprop / dgrad launch: pick the split, point the kernel at the workspace, launch, reduce
template <int TN>
int launch_gemm(const GemmParams& p, cudaStream_t stream);

int run_gemm(GemmParams& p, int TN, size_t out_numel, float* out, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    int min_taps = MAX_TAPS;
    for (int c = 0; c < p.n_cls; ++c) {
        const int T = p.cls_tap_begin[c + 1] - p.cls_tap_begin[c];
        min_taps = T < min_taps ? T : min_taps;
    }
    const int tiles = p.m_tiles * p.n_tiles * p.n_cls;
    p.splits = ((out_numel % 4) || p.generic || p.crelu_half) ? 1 : gemm_splits(tiles, min_taps);   // generic: slices / fused epilogues are not split
    if (p.splits > 1 && (!ws || ws_bytes < (size_t)p.splits * out_numel * sizeof(float))) p.splits = 1;   // no room: unsplit
    p.split_stride = (long long)out_numel;
    p.out = p.splits > 1 ? reinterpret_cast<float*>(ws) : out;
    p.n_full = tiles * p.splits;
    p.tail_splits = 1;
    p.tail_ws = nullptr;
    // 256 x 256 tiles (two sub-tiles share the weight tile) when the launch keeps >= 0.75 waves of them and a tile is long
    // enough to amortise the un-overlapped epilogue
    if (g_use_gemm2 && !p.generic && !p.crelu_half && p.splits == 1 && TN == 256 && (p.m_tiles & 1) == 0 && tiles / 2 >= (kNumSMs * 3) / 4 && min_taps * p.kchunks >= g_gemm2_min_chunks) {
        p.n_items = tiles / 2;
        if (capturing()) { t_capture->kind = 1; t_capture->TN = 256; t_capture->gemm = p; return OTGAN_OK; }
        OTGAN_SET_MAX_SMEM((conv_gemm2_tc_kernel), G2_SMEM_BYTES);
        const int grid2 = p.n_items < kNumSMs ? p.n_items : kNumSMs;
        conv_gemm2_tc_kernel<<<grid2, NUM_THREADS, G2_SMEM_BYTES, stream>>>(p);
        OTGAN_CHECK_LAUNCH("conv_gemm2_tc_kernel");
        return OTGAN_OK;
    }
    int n_tail = 0;
    if (g_tail_split && !p.generic && !p.crelu_half && p.splits == 1 && TN >= 128 && ws) {   // cut the last, partial wave of tiles along the taps
        n_tail = tiles % kNumSMs;
        const int S = tail_splits_for(n_tail, min_taps);
        if (S > 1 && ws_bytes >= (size_t)n_tail * S * TM * TN * sizeof(float)) {
            p.n_full = tiles - n_tail;
            p.tail_splits = S;
            p.tail_ws = reinterpret_cast<float*>(ws);
        } else {
            n_tail = 0;
        }
    }
    p.n_items = p.n_full + n_tail * p.tail_splits;
    const int rc = TN == 256 ? launch_gemm<256>(p, stream) : TN == 128 ? launch_gemm<128>(p, stream)
                 : TN == 32 ? launch_gemm<32>(p, stream) : launch_gemm<16>(p, stream);
    if (rc != OTGAN_OK) return rc;
    if (capturing()) { t_capture->n_tail = n_tail; return OTGAN_OK; }
    if (n_tail > 0) {
        if (TN == 256) conv_tail_fixup_kernel<256><<<n_tail, 256, 0, stream>>>(p);
        else conv_tail_fixup_kernel<128><<<n_tail, 256, 0, stream>>>(p);
        OTGAN_CHECK_LAUNCH("conv_tail_fixup_kernel");
        return OTGAN_OK;
    }
    if (p.splits == 1) return rc;
    const size_t n4 = out_numel / 4;
    const size_t blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < (size_t)(8 * kNumSMs) ? blocks : (size_t)(8 * kNumSMs));
    split_reduce_kernel<<<grid, 256, 0, stream>>>(n4, p.splits, n4, reinterpret_cast<const float4*>(p.out), reinterpret_cast<float4*>(out));
    OTGAN_CHECK_LAUNCH("split_reduce_kernel");
    return OTGAN_OK;
}

template <int TN>
int launch_gemm(const GemmParams& p, cudaStream_t stream)
{
    if (capturing()) { t_capture->kind = 0; t_capture->TN = TN; t_capture->gemm = p; return OTGAN_OK; }
    OTGAN_SET_MAX_SMEM((conv_gemm_tc_kernel<TN>), Cfg<TN>::SMEM_BYTES);
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    conv_gemm_tc_kernel<TN><<<grid, NUM_THREADS, Cfg<TN>::SMEM_BYTES, stream>>>(p);
    OTGAN_CHECK_LAUNCH("conv_gemm_tc_kernel");
    return OTGAN_OK;
}

template <int TN>
int launch_wgrad(const WgradParams& p, cudaStream_t stream)
{
    if (capturing()) { t_capture->kind = 2; t_capture->TN = TN; t_capture->wgrad = p; return OTGAN_OK; }
    OTGAN_SET_MAX_SMEM((conv_wgrad_tc_kernel<TN>), Cfg<TN>::SMEM_BYTES);
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    conv_wgrad_tc_kernel<TN><<<grid, WGRAD_THREADS, Cfg<TN>::SMEM_BYTES, stream>>>(p);
    OTGAN_CHECK_LAUNCH("conv_wgrad_tc_kernel");
    return OTGAN_OK;
}

bool conv_dims_ok(int B, int H, int W, int Cin, int Cout, int kh, int kw, int s, int pt, int pl, int Ho, int Wo)
{
    if (B < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || kh < 1 || kw < 1 || kh * kw > MAX_TAPS) return false;
    if (s != 1 && s != 2) return false;
    if (pt < 0 || pl < 0 || pt >= kh || pl >= kw) return false;
    if (H % s || W % s || Ho != H / s || Wo != W / s) return false;      // TensorFlow 'SAME' on even extents
    return true;
}

// Split-K factor of wgrad.  CTAs are persistent and take items round-robin, so the launch lasts ceil(n / 148) item times:
// pick S (<= 148, >= 16 chunks per split) minimising  compute / wave-efficiency + the partial write / reduce traffic.
int wgrad_splits(int items, int nchunks, double flops, double dw_bytes)
{
    int best = 1;
    double best_t = 1e30;
    for (int S = 1; S <= kNumSMs; ++S) {
        if (S > 1 && nchunks / S < 16) break;
        const int cps = (nchunks + S - 1) / S;
        const int Se = (nchunks + cps - 1) / cps;
        const double n = (double)items * Se;
        const double waves = (double)((long long)((n + kNumSMs - 1) / kNumSMs));
        const double eff = n / (waves * kNumSMs);
        const double t = flops / (7.0e14 * eff) + (Se > 1 ? (2.0 * Se + 1.0) * dw_bytes / 5.0e12 : 0.0);
        if (t < best_t * 0.98) { best_t = t; best = Se; }
    }
    return best;
}

}  // namespace

// y[B,Ho,Wo,Cout] = conv(x[B,H,W,Cin], w[Cout, kh*kw*Cin]) + bias
// workspace of a fprop / dgrad launch whose OUTPUT is [B, H, W, C]: room for the split-K partials (0 when the launch already
// has enough tiles to fill the SMs)
size_t conv_gemm_workspace_bytes(int B, int H, int W, int C)
{
    const long long pix = (long long)B * H * W;
    if (C % 128) return 256;                                          // narrow (TN = 16) outputs are never split
    const int TN = (C % 256 == 0) ? 256 : 128;
    const long long tiles = (pix / TM) * (C / TN);
    if (tiles < 1) return 256;
    const int S = gemm_splits((int)tiles, 8);
    if (S > 1) return (size_t)S * pix * C * sizeof(float) + 256;      // whole-launch split-K partials
    const int r = (int)(tiles % kNumSMs);                             // tail split: r tiles x up to 8 shares
    return r ? (size_t)r * 8 * TM * TN * sizeof(float) + 256 : 256;
}

// crelu != 0: y is [B, Ho, Wo, 2 Cout] = relu(concat([conv + bias, -(conv + bias)], 3)) (the plain epilogue's fused CReLU; Cout % 128 == 0,
// never split over the filter taps: the caller keeps it for launches with at least one tile per SM)
int conv_fprop_launch(int B, int H, int W, int Cin, int Cout, int kh, int kw, int s, int pt, int pl, int Ho, int Wo,
                      const float* x, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t stream, int crelu)
{
    OTGAN_REQUIRE(conv_dims_ok(B, H, W, Cin, Cout, kh, kw, s, pt, pl, Ho, Wo), "conv_fprop: unsupported geometry");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (Cin % BK || (Cout % 128 && Cout > 16) || !pixel_box(TM, Wo, Ho, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_fprop(tcgen05): needs Cin %% 32 == 0, Cout %% 128 == 0 or Cout <= 16, power-of-two output extent tiling into "
                  "128-pixel boxes (B=%d H=%d W=%d Cin=%d Cout=%d)", B, H, W, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 16;
    const int brows = Cout < TN ? Cout : TN;
    p.n_valid = Cout; p.b_box_bytes = brows * BK * 4;
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (!make_view_map(&p.amap[ph * s + pw], x, B, H, W, Cin, s, ph, pw, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    for (int i = s * s; i < 4; ++i) p.amap[i] = p.amap[0];
    if (!conv_map_2d(&p.bmap, w, Cout, kh * kw * Cin, kh * kw * Cin, brows, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    int nt = 0;
    for (int a = 0; a < kh; ++a)
        for (int b = 0; b < kw; ++b) {
            const int oh = a - pt, ow = b - pl;                  // input offset relative to s*oh, s*ow
            const int ph = oh & (s - 1), pw = ow & (s - 1);      // parity (two's complement: -1 & 1 == 1)
            p.taps[nt++] = Tap{ph * s + pw, (ow - pw) / s, (oh - ph) / s, (a * kw + b) * Cin, 0, 0};
        }
    p.n_cls = 1;
    p.cls_tap_begin[0] = 0; p.cls_tap_begin[1] = nt;
    p.cls_out_off[0] = 0;
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * (B / p.bn);
    p.n_tiles = Cout < TN ? 1 : Cout / TN;
    p.kchunks = Cin / BK;
    const long long ldo = crelu ? 2LL * Cout : Cout;
    if (crelu) {
        OTGAN_REQUIRE(TN >= 128, "conv_fprop: the fused CReLU output needs Cout %% 128 == 0");
        p.crelu_half = Cout;
    }
    p.osW = ldo; p.osH = (long long)Wo * ldo; p.osN = (long long)Ho * Wo * ldo;
    p.bias = bias;
    return run_gemm(p, TN, (size_t)B * Ho * Wo * Cout, y, ws, ws_bytes, stream);
}

// dx[B,H,W,Cin] = conv_transpose(dy[B,Ho,Wo,Cout], wt[Cin, kh*kw*Cout])
int conv_dgrad_launch(int B, int H, int W, int Cin, int Cout, int kh, int kw, int s, int pt, int pl, int Ho, int Wo,
                      const float* dy, const float* wt, float* dx, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(conv_dims_ok(B, H, W, Cin, Cout, kh, kw, s, pt, pl, Ho, Wo), "conv_dgrad: unsupported geometry");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (Cout % BK || (Cin % 128 && Cin > 16) || !pixel_box(TM, W / s, H / s, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_dgrad(tcgen05): needs Cout %% 32 == 0, Cin %% 128 == 0 or Cin <= 16, power-of-two extents tiling into "
                  "128-pixel boxes (B=%d H=%d W=%d Cin=%d Cout=%d)", B, H, W, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0) ? 128 : 16;
    const int brows = Cin < TN ? Cin : TN;
    p.n_valid = Cin; p.b_box_bytes = brows * BK * 4;
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    if (!make_view_map(&p.amap[0], dy, B, Ho, Wo, Cout, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.amap[i] = p.amap[0];
    if (!conv_map_2d(&p.bmap, wt, Cin, kh * kw * Cout, kh * kw * Cout, brows, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    // parity classes of the input pixel (ih, iw): ih = s*j + ph gets the taps with (ph + pt - a) % s == 0, from output row
    // oh = j + (ph + pt - a) / s
    int nt = 0;
    p.n_cls = s * s;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const int cls = ph * s + pw;
            p.cls_tap_begin[cls] = nt;
            p.cls_out_off[cls] = ((long long)ph * W + pw) * Cin;
            for (int a = 0; a < kh; ++a) {
                if ((ph + pt - a) % s) continue;
                for (int b = 0; b < kw; ++b) {
                    if ((pw + pl - b) % s) continue;
                    p.taps[nt++] = Tap{0, (pw + pl - b) / s, (ph + pt - a) / s, (a * kw + b) * Cout, 0, 0};
                }
            }
            if (nt == p.cls_tap_begin[cls]) { set_error("conv_dgrad: a parity class has no filter tap (k < stride)"); return OTGAN_EUNSUPPORTED; }
        }
    p.cls_tap_begin[p.n_cls] = nt;
    p.tiles_w = (W / s) / p.bw; p.tiles_h = (H / s) / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * (B / p.bn);
    p.n_tiles = Cin < TN ? 1 : Cin / TN;
    p.kchunks = Cout / BK;
    p.osW = (long long)s * Cin; p.osH = (long long)s * W * Cin; p.osN = (long long)H * W * Cin;
    p.bias = nullptr;
    return run_gemm(p, TN, (size_t)B * H * W * Cin, dx, ws, ws_bytes, stream);
}

// ---------------------------------------------------------------------------------------------- generic fprop / dgrad
// Any channel counts (multiples of 4), any batch, power-of-two spatial extents; operands may be channel SLICES of wider NHWC
// buffers (pixel strides lda / ldo); the weight operand is a K-slice / row-slice of a 3-D tensor [w_K, taps, w_rows].
static int tn_for(int N, int* n_inst)
{
    if (N <= 32) { *n_inst = (N + 15) / 16 * 16; return 32; }
    const int tiles = (N + 255) / 256;
    *n_inst = ((N + tiles - 1) / tiles + 15) / 16 * 16;
    return 256;
}

static bool ex_common(GemmParams& p, const ConvEx& c, int TN)
{
    p.generic = 1;
    p.b_k0 = c.k0;
    p.n_valid = c.N;
    p.b_box_bytes = p.n_inst * BK * 4;
    p.kchunks = (c.Ka + BK - 1) / BK;
    p.n_tiles = (c.N + p.n_inst - 1) / p.n_inst;
    p.bias = c.bias;
    p.epi_mode = c.epi_mode; p.e_add = c.e_add; p.e_z = c.e_z;
    p.accumulate = c.accumulate; p.e_out2 = c.e_out2;
    (void)TN;
    return make_weight_map_3d(&p.bmap, c.w, c.w_K, c.w_taps, c.w_rows, c.w_ldtap, c.w_ldrow, p.n_inst);
}

static bool ex_args_ok(const ConvEx& c)
{
    if (c.B < 1 || c.H < 1 || c.W < 1 || c.kh < 1 || c.kw < 1 || c.kh * c.kw > MAX_TAPS) return false;
    if ((c.stride != 1 && c.stride != 2) || c.pt < 0 || c.pl < 0 || c.pt >= c.kh || c.pl >= c.kw || c.H % c.stride || c.W % c.stride) return false;
    if (c.Ka < 1 || c.N < 1 || (c.lda & 3) || (c.ldo & 3) || (c.w_ldtap & 3) || (c.w_ldrow & 3) || (c.k0 & 3)) return false;
    if (!aligned16(c.a) || !aligned16(c.out) || !aligned16(c.w)) return false;
    if (c.epi_mode == EPI_CRELU8 && (c.N & 7)) return false;
    if (c.epi_mode == EPI_CRELU8_BWD && ((c.N & 15) || c.stride != 1 || !c.e_z || (c.e_ld & 3))) return false;
    if (c.epi_mode == EPI_DENSE_FWD && ((c.N & 15) || c.stride != 1 || !c.e_out2 || (c.e_ld & 3) || !aligned16(c.e_out2))) return false;
    return true;
}

// y[B, H/s, W/s, N] = conv(x[B, H, W, Ka], w) (+ bias) (+ epilogue)
int conv_fprop_ex_launch(const ConvEx& c, cudaStream_t stream)
{
    OTGAN_REQUIRE(ex_args_ok(c), "conv_fprop_ex: bad arguments / alignment");
    const int s = c.stride, Ho = c.H / s, Wo = c.W / s;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (!pixel_box_generic(TM, Wo, Ho, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_fprop_ex(tcgen05): output extent %dx%d must be powers of two", Ho, Wo);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = tn_for(c.N, &p.n_inst);
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (!make_view_map(&p.amap[ph * s + pw], c.a, c.B, c.H, c.W, c.Ka, s, ph, pw, box, CU_TENSOR_MAP_SWIZZLE_128B, c.lda)) return OTGAN_EUNSUPPORTED;
    for (int i = s * s; i < 4; ++i) p.amap[i] = p.amap[0];
    if (!ex_common(p, c, TN)) return OTGAN_EUNSUPPORTED;
    int nt = 0;
    for (int a = 0; a < c.kh; ++a)
        for (int b = 0; b < c.kw; ++b) {
            const int oh = a - c.pt, ow = b - c.pl;
            const int ph = oh & (s - 1), pw = ow & (s - 1);
            p.taps[nt++] = Tap{ph * s + pw, (ow - pw) / s, (oh - ph) / s, 0, c.row0, a * c.kw + b};
        }
    p.n_cls = 1;
    p.cls_tap_begin[0] = 0; p.cls_tap_begin[1] = nt;
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * ceil_div(c.B, p.bn);
    p.m_w = Wo; p.m_h = Ho; p.m_b = c.B;
    p.osW = c.ldo; p.osH = (long long)Wo * c.ldo; p.osN = (long long)Ho * Wo * c.ldo;
    p.esW = c.e_ld; p.esH = (long long)Wo * c.e_ld; p.esN = (long long)Ho * Wo * c.e_ld;
    return run_gemm(p, TN, 0, c.out, nullptr, 0, stream);
}

// dx[B, H, W, N] = conv_transpose(dy[B, H/s, W/s, Ka], w[N rows][taps][K]) (+ epilogue); the weight rows are the INPUT channels
int conv_dgrad_ex_launch(const ConvEx& c, cudaStream_t stream)
{
    OTGAN_REQUIRE(ex_args_ok(c), "conv_dgrad_ex: bad arguments / alignment");
    const int s = c.stride, Ho = c.H / s, Wo = c.W / s;
    OTGAN_REQUIRE(c.kh >= s && c.kw >= s, "conv_dgrad_ex: filter smaller than the stride");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (!pixel_box_generic(TM, Wo, Ho, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_dgrad_ex(tcgen05): extent %dx%d must be powers of two", Ho, Wo);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = tn_for(c.N, &p.n_inst);
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    if (!make_view_map(&p.amap[0], c.a, c.B, Ho, Wo, c.Ka, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B, c.lda)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.amap[i] = p.amap[0];
    if (!ex_common(p, c, TN)) return OTGAN_EUNSUPPORTED;
    int nt = 0;
    p.n_cls = s * s;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const int cls = ph * s + pw;
            p.cls_tap_begin[cls] = nt;
            p.cls_out_off[cls] = ((long long)ph * c.W + pw) * c.ldo;
            for (int a = 0; a < c.kh; ++a) {
                if ((ph + c.pt - a) % s) continue;
                for (int b = 0; b < c.kw; ++b) {
                    if ((pw + c.pl - b) % s) continue;
                    p.taps[nt++] = Tap{0, (pw + c.pl - b) / s, (ph + c.pt - a) / s, 0, c.row0, a * c.kw + b};
                }
            }
            if (nt == p.cls_tap_begin[cls]) { set_error("conv_dgrad_ex: a parity class has no filter tap"); return OTGAN_EUNSUPPORTED; }
        }
    p.cls_tap_begin[p.n_cls] = nt;
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * ceil_div(c.B, p.bn);
    p.m_w = Wo; p.m_h = Ho; p.m_b = c.B;
    p.osW = (long long)s * c.ldo; p.osH = (long long)s * c.W * c.ldo; p.osN = (long long)c.H * c.W * c.ldo;
    p.esW = c.e_ld; p.esH = (long long)c.W * c.e_ld; p.esN = (long long)c.H * c.W * c.e_ld;      // epilogue inputs: stride 1 only
    return run_gemm(p, TN, 0, c.out, nullptr, 0, stream);
}

size_t conv_wgrad_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int kh, int kw)
{
    const int TN = (Cin % 256 == 0) ? 256 : 128;
    const int items = (Cout / TM) * (Cin / TN) * kh * kw;
    const long long P = (long long)B * Ho * Wo;
    const double dw_bytes = 4.0 * Cout * kh * kw * Cin;
    const int S = wgrad_splits(items < 1 ? 1 : items, (int)(P / 32 < 1 ? 1 : P / 32), 0.5 * dw_bytes * (double)P, dw_bytes);
    return S > 1 ? (size_t)S * Cout * kh * kw * Cin * sizeof(float) + 256 : 256;
}

// dw[Cout, kh*kw*Cin] = sum over pixels dy (x) x
int conv_wgrad_launch(int B, int H, int W, int Cin, int Cout, int kh, int kw, int s, int pt, int pl, int Ho, int Wo,
                      const float* dy, const float* x, float* dw, void* ws, size_t ws_bytes, cudaStream_t stream, int hwio)
{
    OTGAN_REQUIRE(conv_dims_ok(B, H, W, Cin, Cout, kh, kw, s, pt, pl, Ho, Wo), "conv_wgrad: unsupported geometry");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    const long long P = (long long)B * Ho * Wo;
    if (Cin % 128 || Cout % TM || Wo > 32 || !pixel_box(32, Wo, Ho, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_wgrad(tcgen05): needs Cin %% 128 == 0, Cout %% 128 == 0, power-of-two output extent <= 32 wide "
                  "(B=%d H=%d W=%d Cin=%d Cout=%d)", B, H, W, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cin % 256 == 0) ? 256 : 128;
    const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    if (!make_view_map(&p.dymap[0], dy, B, Ho, Wo, Cout, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.dymap[i] = p.dymap[0];
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (!make_view_map(&p.xmap[ph * s + pw], x, B, H, W, Cin, s, ph, pw, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return OTGAN_EUNSUPPORTED;
    for (int i = s * s; i < 4; ++i) p.xmap[i] = p.xmap[0];
    int nt = 0;
    for (int a = 0; a < kh; ++a)
        for (int b = 0; b < kw; ++b) {
            const int oh = a - pt, ow = b - pl;
            const int ph = oh & (s - 1), pw = ow & (s - 1);
            p.taps[nt++] = Tap{ph * s + pw, (ow - pw) / s, (oh - ph) / s, (a * kw + b) * Cin, 0, 0};
        }
    p.ntaps = nt;
    p.co_tiles = Cout / TM; p.ci_tiles = Cin / TN;
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh;
    p.nchunks = (int)(P / 32);
    const int items = p.co_tiles * p.ci_tiles * nt;
    const double dw_bytes = 4.0 * Cout * nt * Cin;
    p.splits = wgrad_splits(items, p.nchunks, 0.5 * dw_bytes * (double)P, dw_bytes);
    p.chunks_per_split = ceil_div(p.nchunks, p.splits);
    p.splits = ceil_div(p.nchunks, p.chunks_per_split);
    p.n_items = items * p.splits;
    p.ldw = nt * Cin;
    p.split_stride = (long long)Cout * p.ldw;
    p.hwio = hwio; p.hwio_c = Cout; p.hwio_rows = Cout;
    if (p.splits > 1) {
        OTGAN_REQUIRE(ws && ws_bytes >= (size_t)p.splits * p.split_stride * sizeof(float), "conv_wgrad: workspace too small");
        p.out = reinterpret_cast<float*>(ws);
    } else {
        p.out = dw;
    }
    const int rc = TN == 256 ? launch_wgrad<256>(p, stream) : launch_wgrad<128>(p, stream);
    if (rc != OTGAN_OK || p.splits == 1 || capturing()) return rc;
    const size_t n4 = (size_t)p.split_stride / 4;
    const int grid = (int)((n4 + 255) / 256 < (size_t)(8 * kNumSMs) ? (n4 + 255) / 256 : (size_t)(8 * kNumSMs));
    split_reduce_kernel<<<grid, 256, 0, stream>>>(n4, p.splits, n4, reinterpret_cast<const float4*>(p.out), reinterpret_cast<float4*>(dw));
    OTGAN_CHECK_LAUNCH("split_reduce_kernel");
    return OTGAN_OK;
}

// ---------------------------------------------------------------------------------------------- generic wgrad
// dw[Cout][kh*kw][Cin] = sum over pixels dy[.., co] x[.. + tap, ci] for any channel counts (multiples of 4) / batch; dy and x may
// be channel slices of wider buffers (pixel strides ldy / ldx).  Tiles past Cout / Cin read zeros (TMA) and are not stored.
static int wgrad_ex_tn(int Cin)
{
    const int w256 = ceil_div(Cin, 256) * 256, w128 = ceil_div(Cin, 128) * 128;
    return w128 < w256 ? 128 : 256;
}

size_t conv_wgrad_ex_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int kh, int kw)
{
    const int TN = wgrad_ex_tn(Cin);
    const int items = ceil_div(Cout, TM) * ceil_div(Cin, TN) * kh * kw;
    const long long P = (long long)B * Ho * Wo;
    const double dw_bytes = 4.0 * Cout * kh * kw * Cin;
    const int S = wgrad_splits(items, (int)((P + 31) / 32), 0.5 * dw_bytes * (double)P, dw_bytes);
    return S > 1 ? (size_t)S * Cout * kh * kw * Cin * sizeof(float) + 256 : 256;
}

int conv_wgrad_ex_launch(int B, int H, int W, int Cin, int ldx, int Cout, int ldy, int kh, int kw, int s, int pt, int pl,
                         const float* dy, const float* x, float* dw, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(B >= 1 && H >= 1 && W >= 1 && kh >= 1 && kw >= 1 && kh * kw <= MAX_TAPS && (s == 1 || s == 2) && pt >= 0 && pl >= 0 &&
                  pt < kh && pl < kw && H % s == 0 && W % s == 0, "conv_wgrad_ex: unsupported geometry");
    OTGAN_REQUIRE(Cin >= 4 && Cout >= 4 && !(Cin & 3) && !(Cout & 3) && !(ldx & 3) && !(ldy & 3) && aligned16(dy) && aligned16(x) && aligned16(dw),
                  "conv_wgrad_ex: channel counts / strides must be multiples of 4, pointers 16-byte aligned");
    const int Ho = H / s, Wo = W / s;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    if (!pixel_box_generic(32, Wo, Ho, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_wgrad_ex(tcgen05): output extent %dx%d must be powers of two", Ho, Wo);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = wgrad_ex_tn(Cin);
    const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    if (!make_view_map(&p.dymap[0], dy, B, Ho, Wo, Cout, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, ldy)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.dymap[i] = p.dymap[0];
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (!make_view_map(&p.xmap[ph * s + pw], x, B, H, W, Cin, s, ph, pw, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, ldx)) return OTGAN_EUNSUPPORTED;
    for (int i = s * s; i < 4; ++i) p.xmap[i] = p.xmap[0];
    int nt = 0;
    for (int a = 0; a < kh; ++a)
        for (int b = 0; b < kw; ++b) {
            const int oh = a - pt, ow = b - pl;
            const int ph = oh & (s - 1), pw = ow & (s - 1);
            p.taps[nt++] = Tap{ph * s + pw, (ow - pw) / s, (oh - ph) / s, (a * kw + b) * Cin, 0, 0};
        }
    p.ntaps = nt;
    p.co_tiles = ceil_div(Cout, TM); p.ci_tiles = ceil_div(Cin, TN);
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh;
    p.nchunks = p.tiles_w * p.tiles_h * ceil_div(B, p.bn);
    p.masked = 1; p.m_valid = Cout; p.n_valid = Cin;
    const int items = p.co_tiles * p.ci_tiles * nt;
    const double dw_bytes = 4.0 * Cout * nt * Cin;
    p.splits = wgrad_splits(items, p.nchunks, 0.5 * dw_bytes * (double)B * Ho * Wo, dw_bytes);
    p.chunks_per_split = ceil_div(p.nchunks, p.splits);
    p.splits = ceil_div(p.nchunks, p.chunks_per_split);
    p.n_items = items * p.splits;
    p.ldw = nt * Cin;
    p.split_stride = (long long)Cout * p.ldw;
    if (p.splits > 1 && (!ws || ws_bytes < (size_t)p.splits * p.split_stride * sizeof(float))) {       // no room: unsplit
        p.splits = 1; p.chunks_per_split = p.nchunks; p.n_items = items;
    }
    p.out = p.splits > 1 ? reinterpret_cast<float*>(ws) : dw;
    const int rc = TN == 256 ? launch_wgrad<256>(p, stream) : launch_wgrad<128>(p, stream);
    if (rc != OTGAN_OK || p.splits == 1 || capturing()) return rc;
    const size_t n4 = (size_t)p.split_stride / 4;
    const int grid = (int)((n4 + 255) / 256 < (size_t)(8 * kNumSMs) ? (n4 + 255) / 256 : (size_t)(8 * kNumSMs));
    split_reduce_kernel<<<grid, 256, 0, stream>>>(n4, p.splits, n4, reinterpret_cast<const float4*>(p.out), reinterpret_cast<float4*>(dw));
    OTGAN_CHECK_LAUNCH("split_reduce_kernel");
    return OTGAN_OK;
}

int ohwi_to_ihwo_launch(int Cout, int T, int Cin, const float* w, float* wt, cudaStream_t stream)
{
    dim3 grid(ceil_div(Cin, 32), ceil_div(Cout, 32), T);
    ohwi_to_ihwo_kernel<<<grid, 256, 0, stream>>>(Cout, T, Cin, w, wt);
    OTGAN_CHECK_LAUNCH("ohwi_to_ihwo_kernel");
    return OTGAN_OK;
}

size_t colsum_workspace_bytes(int P, int C)
{
    (void)P;
    return (size_t)512 * C * sizeof(float) + 256;
}

int colsum_launch(int P, int C, const float* x, float* out, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(ws && ws_bytes >= colsum_workspace_bytes(P, C), "colsum: workspace too small");
    int slabs = ceil_div(P, 128);                      // enough CTAs to stream the [P, C] matrix at HBM speed
    slabs = slabs > 512 ? 512 : slabs;
    const int rows_per_slab = ceil_div(P, slabs);
    slabs = ceil_div(P, rows_per_slab);
    float* partial = reinterpret_cast<float*>(ws);
    colsum_partial_kernel<<<dim3(ceil_div(C, 128), slabs), 256, 0, stream>>>(P, C, rows_per_slab, x, partial);
    OTGAN_CHECK_LAUNCH("colsum_partial_kernel");
    const int C4 = C / 4;
    colsum_final_kernel<<<ceil_div(C4, 8), 256, 0, stream>>>(C4, slabs, reinterpret_cast<const float4*>(partial), reinterpret_cast<float4*>(out));
    OTGAN_CHECK_LAUNCH("colsum_final_kernel");
    return OTGAN_OK;
}

// ---------------------------------------------------------------------------------------------- fused upsample + conv
namespace {

inline int floor_div2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

bool make_submap(SubMap& m, int kh, int kw, int pt, int pl)
{
    if (kh < 1 || kw < 1 || kh > 8 || kw > 8) return false;
    m.kh = kh; m.kw = kw;
    int n1[2][2];
    for (int a = 0; a < 2; ++a) {
        m.dmin_h[a] = floor_div2(a - pt);
        m.dmin_w[a] = floor_div2(a - pl);
        for (int k = 0; k < kh; ++k) m.idx_h[a][k] = floor_div2(a + k - pt) - m.dmin_h[a];
        for (int k = 0; k < kw; ++k) m.idx_w[a][k] = floor_div2(a + k - pl) - m.dmin_w[a];
        n1[0][a] = m.idx_h[a][kh - 1] + 1;
        n1[1][a] = m.idx_w[a][kw - 1] + 1;
    }
    if (n1[0][0] != n1[0][1] || n1[1][0] != n1[1][1]) return false;     // both parities must see the same number of sub-taps
    m.n1h = n1[0][0]; m.n1w = n1[1][0];
    for (int a = 0; a < 2; ++a) {
        for (int r = 0; r < 8; ++r) { m.h_lo[a][r] = m.h_hi[a][r] = m.w_lo[a][r] = m.w_hi[a][r] = 0; }
        for (int k = kh - 1; k >= 0; --k) m.h_lo[a][m.idx_h[a][k]] = (unsigned char)k;
        for (int k = 0; k < kh; ++k) m.h_hi[a][m.idx_h[a][k]] = (unsigned char)(k + 1);
        for (int k = kw - 1; k >= 0; --k) m.w_lo[a][m.idx_w[a][k]] = (unsigned char)k;
        for (int k = 0; k < kw; ++k) m.w_hi[a][m.idx_w[a][k]] = (unsigned char)(k + 1);
    }
    return 4 * m.n1h * m.n1w <= MAX_TAPS;
}

bool up2_dims_ok(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pt, int pl)
{
    return B >= 1 && Hl >= 1 && Wl >= 1 && Cin >= 1 && Cout >= 1 && pt >= 0 && pl >= 0 && pt < kh && pl < kw;
}

unsigned ew_grid(size_t n)
{
    const size_t b = (n + 255) / 256, cap = (size_t)kNumSMs * 16;
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int up2_subtaps(int k, int pad)
{
    SubMap m;
    return make_submap(m, k, k, pad, pad) ? m.n1h : 0;
}

int up2_presum_launch(int Cout, int kh, int kw, int Cin, int pt, int pl, const float* w, float* w_sub, cudaStream_t stream)
{
    SubMap m;
    if (!make_submap(m, kh, kw, pt, pl)) { set_error("up2_presum: unsupported filter geometry %dx%d pad %d,%d", kh, kw, pt, pl); return OTGAN_EUNSUPPORTED; }
    OTGAN_REQUIRE(Cin % 4 == 0 && aligned16(w) && aligned16(w_sub), "up2_presum: Cin must be a multiple of 4, buffers 16-byte aligned");
    up2_presum_kernel<<<4 * Cout, (Cin / 4) < 256 ? ((Cin / 4 + 31) / 32) * 32 : 256, 0, stream>>>(m, Cout, Cin, w, w_sub);
    OTGAN_CHECK_LAUNCH("up2_presum_kernel");
    return OTGAN_OK;
}

int up2_unsum_launch(int Cout, int kh, int kw, int Cin, int pt, int pl, const float* dw_sub, float* dw, cudaStream_t stream)
{
    SubMap m;
    if (!make_submap(m, kh, kw, pt, pl)) { set_error("up2_unsum: unsupported filter geometry %dx%d pad %d,%d", kh, kw, pt, pl); return OTGAN_EUNSUPPORTED; }
    OTGAN_REQUIRE(Cin % 4 == 0 && aligned16(dw_sub) && aligned16(dw), "up2_unsum: Cin must be a multiple of 4, buffers 16-byte aligned");
    up2_unsum_kernel<<<ew_grid((size_t)Cout * kh * kw * (Cin / 4)), 256, 0, stream>>>(m, Cout, Cin, dw_sub, dw);
    OTGAN_CHECK_LAUNCH("up2_unsum_kernel");
    return OTGAN_OK;
}

int up2_presum_ihwo_launch(int Cout, int kh, int kw, int Cin, int pt, int pl, const float* w_ihwo, float* w_sub_t, cudaStream_t stream)
{
    SubMap m;
    if (!make_submap(m, kh, kw, pt, pl)) { set_error("up2_presum_ihwo: unsupported filter geometry %dx%d pad %d,%d", kh, kw, pt, pl); return OTGAN_EUNSUPPORTED; }
    OTGAN_REQUIRE(Cout % 4 == 0 && aligned16(w_ihwo) && aligned16(w_sub_t), "up2_presum_ihwo: Cout must be a multiple of 4, buffers 16-byte aligned");
    up2_presum_ihwo_kernel<<<4 * Cin, (Cout / 4) < 256 ? ((Cout / 4 + 31) / 32) * 32 : 256, 0, stream>>>(m, Cout, Cin, w_ihwo, w_sub_t);
    OTGAN_CHECK_LAUNCH("up2_presum_ihwo_kernel");
    return OTGAN_OK;
}

int up2_unsum_hwio_launch(int Cout, int kh, int kw, int Cin, int pt, int pl, const float* dw_sub, float* dw, cudaStream_t stream)
{
    SubMap m;
    if (!make_submap(m, kh, kw, pt, pl)) { set_error("up2_unsum_hwio: unsupported filter geometry %dx%d pad %d,%d", kh, kw, pt, pl); return OTGAN_EUNSUPPORTED; }
    OTGAN_REQUIRE(Cout % 4 == 0 && aligned16(dw_sub) && aligned16(dw), "up2_unsum_hwio: Cout must be a multiple of 4, buffers 16-byte aligned");
    up2_unsum_hwio_kernel<<<kh * kw * Cin, (Cout / 4) < 256 ? ((Cout / 4 + 31) / 32) * 32 : 256, 0, stream>>>(m, Cout, Cin, dw_sub, dw);
    OTGAN_CHECK_LAUNCH("up2_unsum_hwio_kernel");
    return OTGAN_OK;
}

// y[B, 2Hl, 2Wl, Cout] = conv(upsample2x(x_low[B, Hl, Wl, Cin]), W) + bias, W given as the 4 pre-summed sub-filters
// w_sub [4][Cout][n1h*n1w*Cin]
int conv_up2_fprop_launch(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pt, int pl, const float* x_low,
                          const float* w_sub, const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(up2_dims_ok(B, Hl, Wl, Cin, Cout, kh, kw, pt, pl), "conv_up2_fprop: bad geometry");
    SubMap m;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (!make_submap(m, kh, kw, pt, pl) || Cin % BK || Cout % 128 || !pixel_box(TM, Wl, Hl, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_up2_fprop(tcgen05): needs Cin %% 32 == 0, Cout %% 128 == 0, power-of-two low-res extents tiling into 128-pixel "
                  "boxes (B=%d Hl=%d Wl=%d Cin=%d Cout=%d)", B, Hl, Wl, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cout % 256 == 0) ? 256 : 128, slots = m.n1h * m.n1w;
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    if (!make_view_map(&p.amap[0], x_low, B, Hl, Wl, Cin, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.amap[i] = p.amap[0];
    if (!conv_map_2d(&p.bmap, w_sub, 4 * Cout, slots * Cin, slots * Cin, TN, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    p.n_valid = Cout; p.b_box_bytes = TN * BK * 4;
    int nt = 0;
    p.n_cls = 4;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int cls = 2 * a + b;
            p.cls_tap_begin[cls] = nt;
            p.cls_out_off[cls] = ((long long)a * (2 * Wl) + b) * Cout;
            for (int ri = 0; ri < m.n1h; ++ri)
                for (int ci = 0; ci < m.n1w; ++ci)
                    p.taps[nt++] = Tap{0, m.dmin_w[b] + ci, m.dmin_h[a] + ri, (ri * m.n1w + ci) * Cin, cls * Cout, 0};
        }
    p.cls_tap_begin[4] = nt;
    p.tiles_w = Wl / p.bw; p.tiles_h = Hl / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * (B / p.bn);
    p.n_tiles = Cout / TN;
    p.kchunks = Cin / BK;
    p.osW = 2LL * Cout; p.osH = 4LL * Wl * Cout; p.osN = 4LL * Hl * Wl * Cout;
    p.bias = bias;
    return run_gemm(p, TN, (size_t)B * 4 * Hl * Wl * Cout, y, ws, ws_bytes, stream);
}

// dx_low[B, Hl, Wl, Cin] from dy[B, 2Hl, 2Wl, Cout]; w_sub_t [4][Cin][n1h*n1w*Cout] (each sub-filter transposed to IHWO)
int conv_up2_dgrad_launch(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pt, int pl, const float* dy,
                          const float* w_sub_t, float* dx_low, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    OTGAN_REQUIRE(up2_dims_ok(B, Hl, Wl, Cin, Cout, kh, kw, pt, pl), "conv_up2_dgrad: bad geometry");
    SubMap m;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (!make_submap(m, kh, kw, pt, pl) || Cout % BK || Cin % 128 || !pixel_box(TM, Wl, Hl, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_up2_dgrad(tcgen05): needs Cout %% 32 == 0, Cin %% 128 == 0, power-of-two low-res extents tiling into 128-pixel "
                  "boxes (B=%d Hl=%d Wl=%d Cin=%d Cout=%d)", B, Hl, Wl, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cin % 256 == 0) ? 256 : 128, slots = m.n1h * m.n1w;
    const unsigned box[4] = {(unsigned)BK, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            if (!make_view_map(&p.amap[2 * a + b], dy, B, 2 * Hl, 2 * Wl, Cout, 2, a, b, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    if (!conv_map_2d(&p.bmap, w_sub_t, 4 * Cin, slots * Cout, slots * Cout, TN, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return OTGAN_EUNSUPPORTED;
    p.n_valid = Cin; p.b_box_bytes = TN * BK * 4;
    int nt = 0;
    p.n_cls = 1;
    p.cls_tap_begin[0] = 0;
    p.cls_out_off[0] = 0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int ri = 0; ri < m.n1h; ++ri)
                for (int ci = 0; ci < m.n1w; ++ci)      // y(2i'+a, 2j'+b) read x_low(i'+dh, j'+dw): dx_low(i, j) collects dy view (a,b) at (i-dh, j-dw)
                    p.taps[nt++] = Tap{2 * a + b, -(m.dmin_w[b] + ci), -(m.dmin_h[a] + ri), (ri * m.n1w + ci) * Cout, (2 * a + b) * Cin, 0};
    p.cls_tap_begin[1] = nt;
    p.tiles_w = Wl / p.bw; p.tiles_h = Hl / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * (B / p.bn);
    p.n_tiles = Cin / TN;
    p.kchunks = Cout / BK;
    p.osW = Cin; p.osH = (long long)Wl * Cin; p.osN = (long long)Hl * Wl * Cin;
    p.bias = nullptr;
    return run_gemm(p, TN, (size_t)B * Hl * Wl * Cin, dx_low, ws, ws_bytes, stream);
}

size_t conv_up2_wgrad_workspace_bytes(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pt, int pl)
{
    SubMap m;
    if (!make_submap(m, kh, kw, pt, pl)) return 256;
    const int TN = (Cin % 256 == 0) ? 256 : 128, slots = m.n1h * m.n1w;
    const int items = (Cout / TM) * (Cin / TN) * 4 * slots;
    const long long P = (long long)B * Hl * Wl;
    const double dw_bytes = 4.0 * 4 * Cout * slots * Cin;
    const int S = wgrad_splits(items < 1 ? 1 : items, (int)(P / 32 < 1 ? 1 : P / 32), 0.5 * dw_bytes * (double)P, dw_bytes);
    return S > 1 ? (size_t)S * 4 * Cout * slots * Cin * sizeof(float) + 256 : 256;
}

// dw_sub [4][Cout][n1h*n1w*Cin]: filter gradients of the 4 sub-filters (up2_unsum turns them into dW of the 5x5 filter)
int conv_up2_wgrad_launch(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pt, int pl, const float* dy,
                          const float* x_low, float* dw_sub, void* ws, size_t ws_bytes, cudaStream_t stream, int hwio)
{
    OTGAN_REQUIRE(up2_dims_ok(B, Hl, Wl, Cin, Cout, kh, kw, pt, pl), "conv_up2_wgrad: bad geometry");
    SubMap m;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    const long long P = (long long)B * Hl * Wl;
    if (!make_submap(m, kh, kw, pt, pl) || Cin % 128 || Cout % TM || Wl > 32 || !pixel_box(32, Wl, Hl, B, &p.bw, &p.bh, &p.bn)) {
        set_error("conv_up2_wgrad(tcgen05): needs Cin %% 128 == 0, Cout %% 128 == 0, power-of-two low-res extent <= 32 wide "
                  "(B=%d Hl=%d Wl=%d Cin=%d Cout=%d)", B, Hl, Wl, Cin, Cout);
        return OTGAN_EUNSUPPORTED;
    }
    const int TN = (Cin % 256 == 0) ? 256 : 128, slots = m.n1h * m.n1w;
    const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            if (!make_view_map(&p.dymap[2 * a + b], dy, B, 2 * Hl, 2 * Wl, Cout, 2, a, b, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return OTGAN_EUNSUPPORTED;
    if (!make_view_map(&p.xmap[0], x_low, B, Hl, Wl, Cin, 1, 0, 0, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return OTGAN_EUNSUPPORTED;
    for (int i = 1; i < 4; ++i) p.xmap[i] = p.xmap[0];
    int nt = 0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int ri = 0; ri < m.n1h; ++ri)
                for (int ci = 0; ci < m.n1w; ++ci)
                    p.taps[nt++] = Tap{0, m.dmin_w[b] + ci, m.dmin_h[a] + ri, (ri * m.n1w + ci) * Cin, (2 * a + b) * Cout, 2 * a + b};
    p.ntaps = nt;
    p.co_tiles = Cout / TM; p.ci_tiles = Cin / TN;
    p.tiles_w = Wl / p.bw; p.tiles_h = Hl / p.bh;
    p.nchunks = (int)(P / 32);
    const int items = p.co_tiles * p.ci_tiles * nt;
    const double dw_bytes = 4.0 * 4 * Cout * slots * Cin;
    p.splits = wgrad_splits(items, p.nchunks, 0.5 * dw_bytes * (double)P, dw_bytes);
    p.chunks_per_split = ceil_div(p.nchunks, p.splits);
    p.splits = ceil_div(p.nchunks, p.chunks_per_split);
    p.n_items = items * p.splits;
    p.ldw = slots * Cin;
    p.split_stride = 4LL * Cout * p.ldw;
    p.hwio = hwio; p.hwio_c = Cout; p.hwio_rows = Cout;
    if (p.splits > 1) {
        OTGAN_REQUIRE(ws && ws_bytes >= (size_t)p.splits * p.split_stride * sizeof(float), "conv_up2_wgrad: workspace too small");
        p.out = reinterpret_cast<float*>(ws);
    } else {
        p.out = dw_sub;
    }
    const int rc = TN == 256 ? launch_wgrad<256>(p, stream) : launch_wgrad<128>(p, stream);
    if (rc != OTGAN_OK || p.splits == 1 || capturing()) return rc;
    const size_t n4 = (size_t)p.split_stride / 4;
    const int grid = (int)((n4 + 255) / 256 < (size_t)(8 * kNumSMs) ? (n4 + 255) / 256 : (size_t)(8 * kNumSMs));
    split_reduce_kernel<<<grid, 256, 0, stream>>>(n4, p.splits, n4, reinterpret_cast<const float4*>(p.out), reinterpret_cast<float4*>(dw_sub));
    OTGAN_CHECK_LAUNCH("split_reduce_kernel");
    return OTGAN_OK;
}

// Serialises the kernel parameters one of the six convolution passes WOULD be launched with (no GPU, no driver needed).
// op: 0 fprop, 1 dgrad, 2 wgrad, 3 fused-upsample fprop, 4 fused-upsample dgrad, 5 fused-upsample wgrad (H, W = low-res then).
// Layout of out[] (all long long):
//   gemm kernels (kind 0 / 1): kind, TN, n_cls, n_items, splits, m_tiles, n_tiles, bw, bh, bn, tiles_w, tiles_h, kchunks,
//        n_valid, osW, osH, osN, n_full, tail_splits, n_tail, ntaps, cls_tap_begin[5], cls_out_off[4], then ntaps x
//        (map, dw, dh, wcol, brow, dmap)
//   wgrad kernel (kind 2): kind, TN, ntaps, co_tiles, ci_tiles, splits, n_items, bw, bh, bn, tiles_w, tiles_h, nchunks,
//        chunks_per_split, ldw, split_stride, then ntaps x (map, dw, dh, wcol, brow, dmap)
int conv_plan_describe(int op, int B, int H, int W, int Cin, int Cout, int kh, int kw, int s, int pt, int pl,
                       long long* out, int cap)
{
    static float dummy_storage[64];
    float* dummy = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dummy_storage) + 63) & ~(uintptr_t)63);
    void* big_ws = reinterpret_cast<void*>(dummy);            // never dereferenced: nothing is launched while capturing
    const size_t big = (size_t)1 << 40;
    PlanCapture cap_state;
    cap_state.active = true;
    t_capture = &cap_state;
    int rc = OTGAN_EINVAL;
    switch (op) {
    case 0: rc = conv_fprop_launch(B, H, W, Cin, Cout, kh, kw, s, pt, pl, H / s, W / s, dummy, dummy, nullptr, dummy, big_ws, big, nullptr, 0); break;
    case 1: rc = conv_dgrad_launch(B, H, W, Cin, Cout, kh, kw, s, pt, pl, H / s, W / s, dummy, dummy, dummy, big_ws, big, nullptr); break;
    case 2: rc = conv_wgrad_launch(B, H, W, Cin, Cout, kh, kw, s, pt, pl, H / s, W / s, dummy, dummy, dummy, big_ws, big, nullptr, 0); break;
    case 3: rc = conv_up2_fprop_launch(B, H, W, Cin, Cout, kh, kw, pt, pl, dummy, dummy, nullptr, dummy, big_ws, big, nullptr); break;
    case 4: rc = conv_up2_dgrad_launch(B, H, W, Cin, Cout, kh, kw, pt, pl, dummy, dummy, dummy, big_ws, big, nullptr); break;
    case 5: rc = conv_up2_wgrad_launch(B, H, W, Cin, Cout, kh, kw, pt, pl, dummy, dummy, dummy, big_ws, big, nullptr, 0); break;
    default: set_error("conv_plan_describe: unknown op %d", op);
    }
    t_capture = nullptr;
    if (rc != OTGAN_OK) return rc;
    int n = 0;
    auto put = [&](long long v) { if (n < cap) out[n] = v; ++n; };
    auto put_taps = [&](const Tap* taps, int nt) {
        for (int t = 0; t < nt; ++t) { put(taps[t].map); put(taps[t].dw); put(taps[t].dh); put(taps[t].wcol); put(taps[t].brow); put(taps[t].dmap); }
    };
    if (cap_state.kind == 0 || cap_state.kind == 1) {
        const GemmParams& g = cap_state.gemm;
        const int nt = g.cls_tap_begin[g.n_cls];
        put(cap_state.kind); put(cap_state.TN); put(g.n_cls); put(g.n_items); put(g.splits); put(g.m_tiles); put(g.n_tiles);
        put(g.bw); put(g.bh); put(g.bn); put(g.tiles_w); put(g.tiles_h); put(g.kchunks); put(g.n_valid);
        put(g.osW); put(g.osH); put(g.osN); put(g.n_full); put(g.tail_splits); put(cap_state.n_tail); put(nt);
        for (int c = 0; c < 5; ++c) put(g.cls_tap_begin[c]);
        for (int c = 0; c < 4; ++c) put(g.cls_out_off[c]);
        put_taps(g.taps, nt);
    } else if (cap_state.kind == 2) {
        const WgradParams& g = cap_state.wgrad;
        put(2); put(cap_state.TN); put(g.ntaps); put(g.co_tiles); put(g.ci_tiles); put(g.splits); put(g.n_items);
        put(g.bw); put(g.bh); put(g.bn); put(g.tiles_w); put(g.tiles_h); put(g.nchunks); put(g.chunks_per_split);
        put(g.ldw); put(g.split_stride);
        put_taps(g.taps, g.ntaps);
    } else {
        set_error("conv_plan_describe: nothing was captured");
        return OTGAN_EINVAL;
    }
    if (n > cap) { set_error("conv_plan_describe: needs %d values, buffer holds %d", n, cap); return OTGAN_ENOSPC; }
    return n;
}

// A/B switches for benchmarking / tests: 0 = use the 256 x 256 tile variant of the fprop / dgrad kernel (default 1),
// 1 = cut the last partial wave of tiles into shares of the taps (default 0), 2 = minimum K-chunks per tile for option 0 (500)
int conv_set_option(int option, int value)
{
    if (option == 0) { g_use_gemm2 = value != 0; return OTGAN_OK; }
    if (option == 1) { g_tail_split = value != 0; return OTGAN_OK; }
    if (option == 2) { g_gemm2_min_chunks = value < 1 ? 1 : value; return OTGAN_OK; }
    set_error("conv_set_option: unknown option %d", option);
    return OTGAN_EINVAL;
}

}  // namespace otgan
