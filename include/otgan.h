/* otgan.h -- C ABI of libotgan.so: the B200 (sm_100a) implementation of the OT-GAN matching hot path.
 *
 * The reference (openai/ot-gan) has no FFI layer: its boundary for this path is the Python signatures of
 * utils/matching.py (TensorFlow 1.x graph ops).  This header is the C ABI that sits one level below those
 * signatures; otgan_b200/matching.py is the host-side mirror that binds it with ctypes, and INTEGRATION.md shows the
 * stub a reference maintainer would add.  Each entry point cites the reference lines it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (row-major fp32) unless the name ends in _host;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), launches only this library's own
 *     kernels, allocates nothing (workspace is sized by otgan_workspace_bytes_* and passed in) and keeps no global
 *     mutable state except the thread-local last-error string / launch counter and the process-wide A/B switches of
 *     otgan_conv_set_option (measurement aids; they select between kernel variants that compute the same result);
 *   - return value 0 = OTGAN_OK, negative = error (nothing was launched for OTGAN_EINVAL); otgan_last_error() gives text;
 *   - results are deterministic (fixed reduction orders, no floating-point atomics): the same inputs give the same bits.
 */
#ifndef OTGAN_H
#define OTGAN_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTGAN_ABI_VERSION 1

#if defined(__GNUC__)
#define OTGAN_API __attribute__((visibility("default")))
#else
#define OTGAN_API
#endif

enum { OTGAN_OK = 0, OTGAN_EINVAL = -1, OTGAN_ECUDA = -2, OTGAN_ENOSPC = -3, OTGAN_EUNSUPPORTED = -4 };

/* cost kinds (epilogue of the pairwise-dot kernel) */
enum {
    OTGAN_COST_COSINE = 0,      /* 1 - x.y                                         utils/matching.py:31-39   */
    OTGAN_COST_EUCLID_MEAN = 1  /* .5 mean(x^2) + .5 mean(y^2) - x.y / D           toy_example/matching_cpu.py:17-45 */
};

/* implementation selectors (for A/B parity tests and benchmarking; AUTO picks the fastest valid one) */
enum { OTGAN_IMPL_AUTO = 0, OTGAN_IMPL_SIMT = 1, OTGAN_IMPL_TCGEN05 = 2 };

#define OTGAN_MAX_BLOCKS 8   /* two-batch matching uses 6 blocks, single-batch 3 */
#define OTGAN_MAX_TERMS 3
#define OTGAN_MAX_OUTPUTS 8

OTGAN_API int otgan_abi_version(void);
OTGAN_API const char* otgan_last_error(void);
/* number of this library's kernels launched by the calling thread since the last reset (bench.py's gpu_launches) */
OTGAN_API uint64_t otgan_launch_count(void);
OTGAN_API void otgan_reset_launch_count(void);

/* ---- cost blocks ------------------------------------------------------------------------------------------------
 * L[k] = -lam * ( cost(X_k, Y_k) + diag_add[k] * I ),  k < nblk;  X_k: [rows, D] (row stride ldx), Y_k: [cols, D].
 * Replaces the tf.matmul(..., transpose_b=True) / `1. - ` / concat graph of utils/matching.py:21-43 (two-batch, six
 * blocks), :101-111 (single batch, +999 on the diagonal) and toy_example/matching_cpu.py:10-49, fused with the
 * `log_a = -sinkhorn_lambda * distances[i]` scaling of utils/matching.py:50.
 * X_host / Y_host are HOST arrays of nblk device pointers.  ws: otgan_workspace_bytes_cost(...) bytes. */
OTGAN_API size_t otgan_workspace_bytes_cost(int nblk, int rows, int cols, int D, int impl);
OTGAN_API int otgan_cost_blocks_f32(int nblk, int rows, int cols, int D,
                          const float* const* X_host, const float* const* Y_host, int ldx, int ldy,
                          int cost_kind, const float* diag_add_host /* [nblk] or NULL */, float lam,
                          float* L /* [nblk, rows, cols] */, void* ws, size_t ws_bytes, int impl, void* stream);

/* ---- Sinkhorn ---------------------------------------------------------------------------------------------------
 * For each block k: log_a = L0[k]; T x { log_a -= logsumexp(log_a, axis=1); log_a -= logsumexp(log_a, axis=0) };
 * P[k] = softmax(log_a, axis=-1); entropy[k] = mean_i( -sum_j P log_softmax(log_a) ); pc[k] = sum_ij P * (-L0/lam).
 * Replaces utils/matching.py:46-57 (== :113-125, toy_example/matching_cpu.py:51-62).  One persistent kernel: the block
 * stays in registers/shared memory for all T iterations.  P, entropy, pc may each be NULL.  rows, cols <= 128: one CTA per
 * block; <= 512: one 8-CTA thread-block cluster per block (row slabs in registers, column reductions through distributed
 * shared memory); larger blocks (any size), or 128 < side with OTGAN_IMPL_SIMT, stream the L2-resident block with one
 * kernel per half-step and need P (it is the working storage).  L0 and P are float-aligned and, when cols % 4 == 0, share their
 * offset within 16 bytes (both 16-byte aligned in practice): rows move as float4 when L0's block addresses allow it. */
OTGAN_API int otgan_sinkhorn_f32(int nblk, int rows, int cols, int T, float lam,
                       const float* L0 /* [nblk, rows, cols] */, float* P /* [nblk, rows, cols] */,
                       float* entropy /* [nblk] */, float* pc /* [nblk] */, int impl, void* stream);

/* Same, additionally reporting per block how many half-steps took the slow (log-domain, max-subtracted) path of the
 * scaling-form kernel (slow_steps: [nblk] ints or NULL); impl = OTGAN_IMPL_SIMT selects the literal log-domain kernel. */
OTGAN_API int otgan_sinkhorn_ex_f32(int nblk, int rows, int cols, int T, float lam, const float* L0, float* P,
                                    float* entropy, float* pc, int* slow_steps, int impl, void* stream);

/* ---- plan application (matched features / feature gradients) ------------------------------------------------------
 * out[o] = sum_{t < nterms[o]} coef[o][t] * op(P[blk[o][t]]) * F[src[o][t]],   op = transpose if trans[o][t].
 * P blocks are [h, h]; F sources are [h, D] (row stride ldf); outputs are [h, D] (row stride ldo).
 * Replaces the twelve tf.matmul + regroup + 0.5*(f1+f2) of utils/matching.py:63-83 (and :131-134), and with the
 * fused plan of otgan_grad_features_f32 also train.py:111,125-126. */
typedef struct {
    int n_out;
    int nterms[OTGAN_MAX_OUTPUTS];
    int blk[OTGAN_MAX_OUTPUTS][OTGAN_MAX_TERMS];
    int trans[OTGAN_MAX_OUTPUTS][OTGAN_MAX_TERMS];
    int src[OTGAN_MAX_OUTPUTS][OTGAN_MAX_TERMS];
    float coef[OTGAN_MAX_OUTPUTS][OTGAN_MAX_TERMS];
} otgan_plan_t;

/* ws: otgan_workspace_bytes_plan_h(h) bytes (pre-split plan planes of the tensor-core path, h rounded up to 128;
 * otgan_workspace_bytes_plan() is the h <= 128 size); may be NULL, which selects the SIMT kernel. */
OTGAN_API size_t otgan_workspace_bytes_plan(void);
OTGAN_API size_t otgan_workspace_bytes_plan_h(int h);
OTGAN_API int otgan_plan_apply_f32(const otgan_plan_t* plan_host, int h, int D,
                         const float* P /* [nblk, h, h] */, const float* const* F_host /* sources */, int ldf,
                         float* const* out_host /* outputs */, int ldo, void* ws, size_t ws_bytes, int impl, void* stream);

/* Two-batch matched features in the reference's output form: A = [A1;A2], B = [B1;B2] ([2h, D], row stride ld),
 * P = the six plans in the order [a1a2, b2b1, a1b1, a1b2, a2b1, a2b2]; f_* are [2h, D] (row stride ldo).
 * utils/matching.py:63-83. */
OTGAN_API int otgan_matched_two_batch_f32(int h, int D, const float* P, const float* A, const float* B, int ld,
                                float* f_aa, float* f_bb, float* f_ab, float* f_ba, int ldo, void* ws, size_t ws_bytes,
                                int impl, void* stream);
/* Fused form: Ga = f_aa - f_ab, Gb = f_bb - f_ba written directly (train.py:111,125-126). */
OTGAN_API int otgan_grad_features_f32(int h, int D, const float* P, const float* A, const float* B, int ld,
                            float* Ga, float* Gb, int ldo, void* ws, size_t ws_bytes, int impl, void* stream);
/* Same for the rows [row_lo, row_hi) of Ga / Gb only (what one data-parallel rank back-propagates: its own towers, train.py:72-85):
 * computes the half-blocks that intersect the range and leaves the other rows untouched. */
OTGAN_API int otgan_grad_features_rows_f32(int h, int D, const float* P, const float* A, const float* B, int ld, float* Ga,
                                           float* Gb, int ldo, int row_lo, int row_hi, void* ws, size_t ws_bytes, int impl,
                                           void* stream);
/* Single-batch matched features: P = [P_aa, P_bb, P_ab] ([n, n] each), A, B: [n, D].  utils/matching.py:131-134. */
OTGAN_API int otgan_matched_single_batch_f32(int n, int D, const float* P, const float* A, const float* B, int ld,
                                   float* f_aa, float* f_bb, float* f_ab, float* f_ba, int ldo, void* ws,
                                   size_t ws_bytes, int impl, void* stream);

/* ---- distance ---------------------------------------------------------------------------------------------------
 * out[0] = ( sum(B*f_bb) + sum(A*f_aa) - 2 sum(A*f_ab) ) * scale      utils/matching.py:139-153 (scale = 1/(2 bs G)),
 * toy_example/matching_cpu.py:155-164 (scale = 1/(2 n D)).  A, B, f_*: [n, D] with row stride ld.
 * ws: otgan_workspace_bytes_distance(n, D) bytes. */
OTGAN_API size_t otgan_workspace_bytes_distance(int n, int D);
OTGAN_API int otgan_calc_distance_f32(int n, int D, const float* A, const float* B, const float* f_aa, const float* f_bb,
                            const float* f_ab, int ld, float scale, float* out /* [1] */, void* ws, size_t ws_bytes,
                            void* stream);
/* out[0] = (pc[2]+pc[3]+pc[4]+pc[5] - 2 pc[0] - 2 pc[1]) / (2 N); out[1] = mean(entropy[0..5]).
 * The <P,C> form of calc_distance (valid because every plan row sums to one). */
OTGAN_API int otgan_distance_from_pc_f32(const float* pc /* [6] */, const float* entropy /* [6] */, int n_total,
                               float* out /* [2] */, void* stream);

/* ---- optimiser / critic head -------------------------------------------------------------------------------------
 * One fused pass of nn.adam_updates (utils/nn.py:50-73; epsilon inside the root) over a flat parameter buffer, plus the
 * generator's ExponentialMovingAverage update (train.py:63-64, 223) when ema != NULL:
 *   v = mom1 v + (1-mom1) g; mg = mom2 mg + (1-mom2) g^2; p -= lr * (v/d1) / sqrt(mg/d2 + 1e-8); ema -= (1-decay)(ema - p)
 * d1 = 1 - mom1^t, d2 = 1 - mom2^t (t starts at 1).  v may be NULL (mom1 == 0).  n % 4 == 0, 16-byte aligned buffers. */
OTGAN_API int otgan_adam_ema_f32(size_t n, float* p, const float* g, float* v, float* mg, float* ema, float lr, float mom1,
                                 float mom2, float d1, float d2, float ema_decay, void* stream);
/* Same update with the step-dependent scalars read from DEVICE memory: hyper_dev = [lr, d1, d2].  The launch arguments are
 * then identical every step, so a training step captured in a CUDA graph can be replayed (the host refreshes hyper_dev). */
OTGAN_API int otgan_adam_ema_dev_f32(size_t n, float* p, const float* g, float* v, float* mg, float* ema,
                                     const float* hyper_dev, float mom1, float mom2, float ema_decay, void* stream);
/* Critic head (models/dcgan.py:16-19, models/densenet.py:37-42): y = z / ||z||, z = concat(relu(x), relu(-x)) over the
 * channel axis then flattened; x: [B, HW, C] (NHWC), y: [B, HW*2C], inv_norm: [B].  Backward: dx from dy. */
OTGAN_API int otgan_crelu_l2norm_fwd_f32(int B, int HW, int C, const float* x, float* y, float* inv_norm, void* stream);
OTGAN_API int otgan_crelu_l2norm_bwd_f32(int B, int HW, int C, const float* x, const float* y, const float* inv_norm,
                                         const float* dy, float* dx, void* stream);

/* Weight-norm reparameterisation (utils/nn.py:176-180: W = l2_normalize(V, all-but-last) * g), fused with the layout
 * change the convolution needs.  V: [K, C] (HWIO kernel flattened, or the [in, out] dense matrix; C = output channels
 * contiguous), g: [C].  Forward writes Wt: [C, K] (OHWI == channels-last OIHW) and inv_norm: [C].  Backward takes dWt [C, K]
 * and writes dV [K, C], dg [C].  ws: otgan_workspace_bytes_weightnorm(K, C) bytes. */
OTGAN_API size_t otgan_workspace_bytes_weightnorm(int K, int C);
OTGAN_API int otgan_weightnorm_fwd_f32(int K, int C, const float* V, const float* g, float* Wt, float* inv_norm, void* ws,
                                       size_t ws_bytes, void* stream);
OTGAN_API int otgan_weightnorm_bwd_f32(int K, int C, const float* V, const float* g, const float* inv_norm,
                                       const float* dWt, float* dV, float* dg, void* ws, size_t ws_bytes, void* stream);

/* ---- fused activations of the conv stacks (NHWC fp32, C % 4 == 0) ---------------------------------------------------
 * crelu_pad: z [B, H+pt+pb, W+pl+pr, 2C] = zero-pad( relu(concat([x, -x], channel)) ), x: [B,H,W,C] -- the CReLU
 *            pre-activation of nn.conv2d (utils/nn.py:198-200) written into the TensorFlow-'SAME'-padded input of the conv.
 * glu_up   : out [B, up*H, up*W, C] = nearest_upsample_x{1,2}( y[..., :C] * sigmoid(y[..., C:]) ), y: [B,H,W,2C] -- the
 *            generator's gated linear unit + tf.image.resize_nearest_neighbor (models/dcgan.py:39-48). */
OTGAN_API int otgan_crelu_pad_fwd_f32(int B, int H, int W, int C, int pad_top, int pad_left, int pad_bottom, int pad_right,
                                      const float* x, float* z, void* stream);
OTGAN_API int otgan_crelu_pad_bwd_f32(int B, int H, int W, int C, int pad_top, int pad_left, int pad_bottom, int pad_right,
                                      const float* x, const float* dz, float* dx, void* stream);
OTGAN_API int otgan_glu_up_fwd_f32(int B, int H, int W, int C, int up, const float* y, float* out, void* stream);
OTGAN_API int otgan_glu_up_bwd_f32(int B, int H, int W, int C, int up, const float* y, const float* dout, float* dy, void* stream);

/* ---- convolutions (utils/nn.py:234-241, 328-338: tf.nn.conv2d(x, W, [1,s,s,1], 'SAME') + tf.nn.bias_add, NHWC fp32) ------
 * im2col-free implicit GEMM on the tensor cores (TMA boxes of the NHWC tensors, tcgen05 kind::tf32, fp32 accumulation).
 * Geometry: x [B,H,W,Cin], y / dy [B,H/stride,W/stride,Cout], filter kh x kw, stride 1 or 2, pad_top / pad_left = the
 * TensorFlow 'SAME' leading padding (5x5/s1: 2, 5x5/s2: 1, 3x3/s1: 1, 3x3/s2: 0); the trailing padding is implied.
 * Weights: w_ohwi [Cout, kh*kw*Cin] (what otgan_weightnorm_fwd_f32 writes), w_ihwo [Cin, kh*kw*Cout] (otgan_ohwi_to_ihwo_f32).
 * Shapes the tensor-core kernels do not tile (channels not a multiple of 32/128, extents not powers of two) return
 * OTGAN_EUNSUPPORTED.  Pointers 16-byte aligned.
 *   fprop : y  = conv(x, w) + bias                       (bias may be NULL)
 *   dgrad : dx = d conv / d x  applied to dy             (what tf.gradients emits as Conv2DBackpropInput)
 *           fprop / dgrad take an optional workspace of otgan_workspace_bytes_conv_gemm(dims of the OUTPUT tensor) bytes:
 *           with it, launches that have fewer output tiles than SMs (small per-GPU batches) split the filter taps over
 *           more CTAs and sum the partials in a fixed order; ws = NULL runs unsplit.
 *   wgrad : dw_ohwi = d conv / d w applied to dy         (Conv2DBackpropFilter); ws: otgan_workspace_bytes_conv_wgrad
 *   colsum: out[C] = column sums of x [P, C]              (BiasAddGrad); ws: otgan_workspace_bytes_colsum */
OTGAN_API size_t otgan_workspace_bytes_conv_gemm(int B, int H, int W, int C);
OTGAN_API int otgan_conv2d_fprop_tf32(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                      int pad_left, const float* x, const float* w_ohwi, const float* bias, float* y,
                                      void* ws, size_t ws_bytes, void* stream);
/* Same convolution with the NEXT layer's CReLU pre-activation (utils/nn.py:198-200 on one tensor: relu(concat([y, -y], 3))) written by
 * the epilogue: z [B, H/s, W/s, 2 Cout] = [relu(conv + bias) | relu(-(conv + bias))].  Cout % 128 == 0; the launch is never split over
 * the filter taps (use it when the output has at least one 128-pixel x TN tile per SM).  models/dcgan.py:10-13. */
OTGAN_API int otgan_conv2d_fprop_crelu_tf32(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                            int pad_left, const float* x, const float* w_ohwi, const float* bias, float* z,
                                            void* stream);
/* Backward of that CReLU from the ACTIVATED tensor: dy [P, C] = (z_pos > 0 ? dz_pos : 0) - (z_neg > 0 ? dz_neg : 0), z / dz: [P, 2C]. */
OTGAN_API int otgan_crelu_bwd_from_activated_f32(long long P, int C, const float* z, const float* dz, float* dy, void* stream);
OTGAN_API int otgan_conv2d_dgrad_tf32(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                      int pad_left, const float* dy, const float* w_ihwo, float* dx, void* ws,
                                      size_t ws_bytes, void* stream);
OTGAN_API size_t otgan_workspace_bytes_conv_wgrad(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride);
OTGAN_API int otgan_conv2d_wgrad_tf32(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                      int pad_left, const float* dy, const float* x, float* dw_ohwi, void* ws,
                                      size_t ws_bytes, void* stream);
/* A/B switches for measurements and tests (process-wide).  option 0: let fprop / dgrad use the 256 x 256-tile kernel variant
 * where it applies (default 1); option 1: cut the last partial wave of tiles into shares of the filter taps (default 0: no
 * measured gain); option 2: minimum K-chunks per tile for option 0 (default 500).  Results are identical either way up to
 * the floating-point summation order. */
OTGAN_API int otgan_conv_set_option(int option, int value);
/* Host-only introspection (no GPU, no driver): the kernel parameters -- tile boxes, parity classes, filter-tap tables
 * (view, pixel shift, weight column / row), output strides, split factors, kernel variant -- that one convolution pass would
 * be launched with.  op: 0 fprop, 1 dgrad, 2 wgrad, 3 / 4 / 5 the fused-upsample fprop / dgrad / wgrad (H, W = the LOW-
 * resolution extent).  Writes long long values into out_host (layout documented at conv_plan_describe in conv_tc.cu) and
 * returns their number, or a negative error.  tests/test_conv_plan_host.py replays these tables in numpy against the
 * reference convolution, so the host logic of the kernels is checked on a machine without a GPU. */
OTGAN_API int otgan_conv_plan_describe(int op, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                       int pad_left, long long* out_host, int capacity);
OTGAN_API int otgan_ohwi_to_ihwo_f32(int Cout, int taps, int Cin, const float* w_ohwi, float* w_ihwo, void* stream);
/* ---- fused 2x nearest-neighbour upsample + convolution (models/dcgan.py:37-46: resize_nearest_neighbor -> nn.conv2d) -----------
 * y [B, 2Hl, 2Wl, Cout] = conv(upsample2x(x_low [B, Hl, Wl, Cin]), W, stride 1, 'SAME') + bias without materialising the
 * upsampled tensor: output pixel (2i+a, 2j+b) reads x_low rows i + floor((a + kh - pad)/2), so the kh taps of the filter
 * collapse onto n1 = otgan_up2_subtaps(k, pad) low-resolution rows (5x5: 3, 3x3: 2) and each of the 4 output parity classes
 * (a, b) is an n1 x n1 convolution of x_low with a pre-summed sub-filter: 9 taps instead of 25 for the generator's layers.
 *   w_sub      [4][Cout][n1*n1*Cin]   = otgan_up2_weight_presum_f32(w_ohwi)             (class index 2a + b)
 *   w_sub_ihwo [4][Cin][n1*n1*Cout]   = otgan_ohwi_to_ihwo_f32 of each class of w_sub   (dgrad)
 *   dw_ohwi = otgan_up2_weight_unsum_f32(dw_sub): the chain rule of the pre-sum          (after wgrad)
 * Exact algebra of the reference's two ops; only the floating-point summation order differs. */
OTGAN_API int otgan_up2_subtaps(int k, int pad);
OTGAN_API int otgan_up2_weight_presum_f32(int Cout, int kh, int kw, int Cin, int pad_top, int pad_left, const float* w_ohwi,
                                          float* w_sub, void* stream);
OTGAN_API int otgan_up2_weight_unsum_f32(int Cout, int kh, int kw, int Cin, int pad_top, int pad_left, const float* dw_sub,
                                         float* dw_ohwi, void* stream);
OTGAN_API int otgan_conv2d_up2_fprop_tf32(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pad_top, int pad_left,
                                          const float* x_low, const float* w_sub, const float* bias, float* y, void* ws,
                                          size_t ws_bytes, void* stream);
OTGAN_API int otgan_conv2d_up2_dgrad_tf32(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pad_top, int pad_left,
                                          const float* dy, const float* w_sub_ihwo, float* dx_low, void* ws, size_t ws_bytes,
                                          void* stream);
OTGAN_API size_t otgan_workspace_bytes_conv_up2_wgrad(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pad_top,
                                                      int pad_left);
OTGAN_API int otgan_conv2d_up2_wgrad_tf32(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pad_top, int pad_left,
                                          const float* dy, const float* x_low, float* dw_sub, void* ws, size_t ws_bytes,
                                          void* stream);

/* Narrow (<= 16 channel) side of the 3-channel layers, computed as one GEMM over all (tap, channel) columns + a shift
 * (conv_narrow.cu).  off_t = (kh - pad_top, kw - pad_left), negated when flip != 0 (gradient orientation).
 *   im2col: col[px][t*C + c] = x[px + off_t][c] (0 outside the image), columns kh*kw*C .. ldc-1 zero; x: [B,H,W,C]
 *   col2im: y[px][c] = bias[c] + sum_t z[px + off_t][t*C + c] (rows outside the image contribute 0); z: [B*H*W, ldz] */
OTGAN_API int otgan_im2col_narrow_f32(int B, int H, int W, int C, int kh, int kw, int pad_top, int pad_left, int flip,
                                      const float* x, float* col, int ldc, void* stream);
OTGAN_API int otgan_col2im_narrow_f32(int B, int H, int W, int C, int kh, int kw, int pad_top, int pad_left, int flip,
                                      const float* z, int ldz, const float* bias, float* y, void* stream);
OTGAN_API size_t otgan_workspace_bytes_colsum(int P, int C);
OTGAN_API int otgan_colsum_f32(int P, int C, const float* x, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- parameter-space pipeline in the variable's own layout ---------------------------------------------------------------------------
 * The variables are HWIO (utils/nn.py:123: [kh, kw, Cin, Cout] == [K, C] with the OUTPUT channel contiguous).  These forms keep the
 * whole gradient path in that layout, so that only ONE transpose (V -> the OHWI operand of fprop) is left per layer and step:
 *   weightnorm_fwd2     W (OHWI) and, in the same pass over V, W_ihwo [Cin][taps][Cout] = the dgrad operand (a scaled row permutation
 *                       of V -- no transpose); W_ihwo may be NULL
 *   wgrad_hwio          the filter gradient written as [taps * Cin][Cout] (dV's layout; the 32 lanes of a warp own 32 consecutive
 *                       output channels: 128-byte segments); the fused-upsample form writes [4][n1*n1 * Cin][Cout]
 *   up2_..._ihwo/_hwio  the sub-filter pre-sum for dgrad built from W_ihwo, and the chain rule of the pre-sum on HWIO gradients
 *   weightnorm_bwd_hwio dV, dg from an HWIO filter gradient: pure streaming, float4 along Cout */
OTGAN_API int otgan_weightnorm_fwd2_f32(int K, int C, int taps, const float* V, const float* g, float* Wt, float* W_ihwo,
                                        float* inv_norm, void* ws, size_t ws_bytes, void* stream);
OTGAN_API int otgan_weightnorm_bwd_hwio_f32(int K, int C, const float* V, const float* g, const float* inv_norm, const float* dW_hwio,
                                            float* dV, float* dg, void* ws, size_t ws_bytes, void* stream);
OTGAN_API int otgan_conv2d_wgrad_hwio_tf32(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad_top,
                                           int pad_left, const float* dy, const float* x, float* dw_hwio, void* ws, size_t ws_bytes,
                                           void* stream);
OTGAN_API int otgan_conv2d_up2_wgrad_hwio_tf32(int B, int Hl, int Wl, int Cin, int Cout, int kh, int kw, int pad_top, int pad_left,
                                               const float* dy, const float* x_low, float* dw_sub_hwio, void* ws, size_t ws_bytes,
                                               void* stream);
OTGAN_API int otgan_up2_weight_presum_ihwo_f32(int Cout, int kh, int kw, int Cin, int pad_top, int pad_left, const float* w_ihwo,
                                               float* w_sub_ihwo, void* stream);
OTGAN_API int otgan_up2_weight_unsum_hwio_f32(int Cout, int kh, int kw, int Cin, int pad_top, int pad_left, const float* dw_sub_hwio,
                                              float* dw_hwio, void* stream);

/* ---- generic convolutions (any channel counts / batch; utils/nn.py:234-241, 328-338 accept any shape) ---------------------------------
 * The same tcgen05 kernels in their generic mode: channel counts are multiples of 4, spatial extents powers of two, the batch is
 * arbitrary; x / y / dy / dx may be channel SLICES of wider NHWC buffers (pixel strides ldx / ldy in floats, pointers at the first
 * channel used) -- this is how DenseNet's growing list input (models/densenet.py:10-15) is read without a concatenation.
 * Weights: w [Cout][kh*kw][Cin] (fprop), w_ihwo [Cin][kh*kw][Cout] (dgrad), dense.  TMA zero-fills the K tail of every tap and the
 * rows past Cout, partial tiles are masked in the epilogue.
 * epilogue (fprop): OTGAN_EPI_NONE  y = conv + bias                                       [.., Cout] at pixel stride ldy
 *                   OTGAN_EPI_CRELU8 y = crelu8(conv + bias): channel c -> relu at (c/8)*16 + c%8, relu(-.) at (c/8)*16 + 8 + c%8
 *                                                                                          [.., 2 Cout] at pixel stride ldy (Cout % 8 == 0)
 * The crelu8 order is this library's layout of CReLU-activated tensors (utils/nn.py:198-200 concatenates [x, -x] per list
 * element); otgan_crelu8_perm_host gives the matching permutation of a filter's input channels. */
enum { OTGAN_EPI_NONE = 0, OTGAN_EPI_CRELU8 = 1 };
OTGAN_API int otgan_conv2d_fprop_ex_tf32(int B, int H, int W, int Cin, int ldx, int Cout, int ldy, int kh, int kw, int stride,
                                         int pad_top, int pad_left, const float* x, const float* w, const float* bias, float* y,
                                         int epilogue, void* stream);
OTGAN_API int otgan_conv2d_dgrad_ex_tf32(int B, int H, int W, int Cin, int ldx, int Cout, int ldy, int kh, int kw, int stride,
                                         int pad_top, int pad_left, const float* dy, const float* w_ihwo, float* dx, void* stream);
OTGAN_API size_t otgan_workspace_bytes_conv_wgrad_ex(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride);
OTGAN_API int otgan_conv2d_wgrad_ex_tf32(int B, int H, int W, int Cin, int ldx, int Cout, int ldy, int kh, int kw, int stride,
                                         int pad_top, int pad_left, const float* dy, const float* x, float* dw, void* ws,
                                         size_t ws_bytes, void* stream);
/* crelu8 activation of a [P, C] tensor (row stride ldx) into a 2C-channel slot (row stride ldz), and its backward:
 * dx[c] = (z_pos > 0 ? dz_pos : 0) - (z_neg > 0 ? dz_neg : 0).  C % 8 == 0. */
OTGAN_API int otgan_crelu8_fwd_f32(long long P, int C, const float* x, int ldx, float* z, int ldz, void* stream);
OTGAN_API int otgan_crelu8_bwd_f32(long long P, int C, const float* z, int ldz, const float* dz, int lddz, float* dx, int lddx,
                                   void* stream);
/* perm_host[t * C2 + cz] = t * C2 + cref: row of the reference's HWIO kernel (input channel order [x_0, -x_0, x_1, -x_1, ...] over
 * the list elements, utils/nn.py:198-200) that feeds crelu8 channel cz; C2 = 2 * sum(elem_ch).  Returns taps * C2 or < 0. */
OTGAN_API int otgan_crelu8_perm_host(int n_elem, const int* elem_ch, int taps, int* perm_host, int capacity);
/* Weight norm with an optional row permutation (device int array, NULL = identity) and an optionally strided Wt / dWt
 * ([C][taps][cin] with strides ldrow / ldtap floats when cin > 0): the generic form of otgan_weightnorm_{fwd,bwd}_f32. */
OTGAN_API int otgan_weightnorm_fwd_ex_f32(int K, int C, const float* V, const float* g, const int* perm_dev, int cin,
                                          long long ldtap, long long ldrow, float* Wt, float* inv_norm, void* ws, size_t ws_bytes,
                                          void* stream);
OTGAN_API int otgan_weightnorm_bwd_ex_f32(int K, int C, const float* V, const float* g, const float* inv_norm, const int* perm_dev,
                                          int cin, long long ldtap, long long ldrow, const float* dWt, float* dV, float* dg, void* ws,
                                          size_t ws_bytes, void* stream);

/* ---- DenseNet dense block (models/densenet.py:10-15, 56-61: `block`) ---------------------------------------------------------------
 * x = [e_0 .. e_{n_base-1}] (base elements, base_ch[i] channels each, multiples of 8); L times: x.append(conv3x3(crelu(x), 16)).
 * Z [B,H,W,Ctot], Ctot = otgan_dense_channels = 2 (sum base_ch + 16 L): the crelu8-activated list, element after element.  The
 * caller fills the base slots (otgan_crelu8_fwd_f32); fprop runs the L layers in "contribution" form: launch j convolves slot j of Z
 * with the filters of ALL later layers (N = 16 (L - j) columns of the pre-activation accumulator S [B,H,W,16L], scratch) and writes
 * the layer that is complete at that point into slot j + 1 from its epilogue.
 * wf_all [16L][9][Ctot]: wf_all[16 r + co][t][ci] = W_r[co][t][ci] with the input channels in Z order, ZERO for ci >= cin_r =
 * 2 (sum base_ch + 16 r) (otgan_weightnorm_fwd_ex_f32 with the otgan_crelu8_perm_host permutation and strides ldtap = Ctot,
 * ldrow = 9 Ctot into a zeroed buffer); bias_all [16L] or NULL.
 * bgrad: dZ = gradient w.r.t. Z from the block's consumer; writes dY [B,H,W,16L] (gradients of the L pre-activations),
 * dbase_host[i] [B,H,W,base_ch[i]], dW_all [16L][9][Ctot] (row block r valid for ci < cin_r) and db_all [16L] (either may be NULL).
 * WB [Ctot][9][16L] = otgan_dense_build_wb_f32(wf_all): the weight operand of the backward "gather" convolutions. */
typedef struct { int B, H, W, n_base, base_ch[4], L, growth; } otgan_dense_geom_t;
OTGAN_API int otgan_dense_channels(const otgan_dense_geom_t* geom);
OTGAN_API size_t otgan_dense_wb_floats(const otgan_dense_geom_t* geom);
OTGAN_API int otgan_dense_build_wb_f32(const otgan_dense_geom_t* geom, const float* wf_all, float* WB, void* stream);
OTGAN_API int otgan_dense_block_fprop_tf32(const otgan_dense_geom_t* geom, const float* wf_all, const float* bias_all, float* Z,
                                           float* S, void* stream);
OTGAN_API size_t otgan_workspace_bytes_dense_bgrad(const otgan_dense_geom_t* geom);
OTGAN_API int otgan_dense_block_bgrad_tf32(const otgan_dense_geom_t* geom, const float* Z, const float* dZ, const float* WB,
                                           float* dY, float* const* dbase_host, float* dW_all, float* db_all, void* ws,
                                           size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTGAN_H */
