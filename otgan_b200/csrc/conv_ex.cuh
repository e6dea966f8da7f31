// conv_ex.cuh -- interface of the generic convolution launchers of conv_tc.cu (used by dense_block.cu and abi.cu).
#pragma once
#include "common.cuh"

namespace otgan {

// Epilogues of the generic mode ("crelu8" slot order: see conv_tc.cu).
enum { EPI_PLAIN = 0, EPI_CRELU8 = 1, EPI_CRELU8_BWD = 2, EPI_DENSE_FWD = 3 };

// One generic fprop / dgrad launch.  Any channel counts (multiples of 4), any batch, power-of-two spatial extents; operands
// may be channel SLICES of wider NHWC buffers (pixel strides lda / ldo); the weight operand is a K-slice / row-slice of a 3-D
// tensor [w_K, taps, w_rows].
struct ConvEx {
    int B, H, W;                 // extent of the convolution's INPUT tensor x (dgrad: of dx)
    int kh, kw, stride, pt, pl;
    const float* a; int Ka, lda; // activation operand: Ka channels used (contraction length per tap), pixel stride lda
    float* out; int N, ldo;      // N output columns of the GEMM; out pixel stride ldo
    const float* w; int w_K, w_taps, w_rows; long long w_ldtap, w_ldrow; int k0, row0;
    const float* bias;
    int epi_mode; const float* e_add; const float* e_z; int e_ld;
    int accumulate;              // EPI_DENSE_FWD: add the previous contents of `out` (read-modify-write of the pre-activation buffer)
    float* e_out2;               // EPI_DENSE_FWD: the first 16 columns (the layer that just completed) go here as crelu8 (pixel stride e_ld)
};
int conv_fprop_ex_launch(const ConvEx& c, cudaStream_t stream);
int conv_dgrad_ex_launch(const ConvEx& c, cudaStream_t stream);
size_t conv_wgrad_ex_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int kh, int kw);
int conv_wgrad_ex_launch(int B, int H, int W, int Cin, int ldx, int Cout, int ldy, int kh, int kw, int s, int pt, int pl,
                         const float* dy, const float* x, float* dw, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t colsum_workspace_bytes(int P, int C);
int colsum_launch(int P, int C, const float* x, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace otgan
