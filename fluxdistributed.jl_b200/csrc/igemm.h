// Host entry points of the tcgen05 implicit-GEMM kernels (igemm_sm100.cu). All tensors bf16 NHWC / KRSC.
#pragma once
#include <cuda_runtime.h>

int fd_num_sms();
void fd_set_cluster_multicast(int on);
void fd_set_epilogue8(int on);  // eight-epilogue-warp conv kernels (default since round 2; env FLUXDIST_EPI8=0 for four)
void fd_set_tile_shape(int bn, int mt);  // force the conv CTA tile (bn in {64,128,256}, mt m-tiles per CTA; 0 = automatic)
int fd_gemm_nt(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s);
int fd_gemm_nn(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s);
int fd_gemm_tn(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s);
int fd_conv_fprop(const void* x, const void* w, void* y, float* sum, float* sumsq, int N, int H, int W, int C, int K, int R, int S, int stride,
                  int pad, cudaStream_t s);
// "dz epilogue" of the stride-1 data gradient (see KParams::bz_* in igemm_sm100.cu): BN layer whose output gradient this is
struct FdBnEpilogue {
  const void* y;      // raw conv output of that layer [N,H,W,C] bf16
  const void* out;    // its post-residual activation (mask source) or null (mask recomputed from y)
  const float *mean, *rstd, *gamma, *beta;
  float *s0, *s1;     // accumulators (zeroed by the consumer): sum dz*(y-mean), sum dz
};
int fd_conv_dgrad(const void* dy, const void* w, void* dx, const void* add, const FdBnEpilogue* bz, int N, int H, int W, int C,
                  int K, int R, int S, int pad, cudaStream_t s);
// stride-2 data gradient without a zero-upsampled tensor: 4 output-parity classes, each a stride-1 correlation with the
// taps that reach it (R,S <= 3; H,W even).  `add_compact` ([N,H/2,W/2,C], optional) is summed into the (even,even) class:
// it carries the data gradient of a parallel 1x1/stride-2 projection.  Parity classes no tap reaches are NOT written.
int fd_conv_dgrad_s2(const void* dy, const void* w, void* dx, const void* add_compact, int N, int H, int W, int C, int K,
                     int R, int S, int pad, cudaStream_t s);
int fd_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int C, int K, int R, int S,
                  int stride, int pad, int splits, int accumulate, cudaStream_t s);

// halo-tile 3x3 / stride-1 / pad-1 kernels (conv_halo_sm100.cu); variant 0 = auto
int fd_conv3x3_fprop(const void* x, const void* w, void* y, float* sum, float* sumsq, int N, int H, int W, int C, int K,
                     int variant, cudaStream_t s);
int fd_conv3x3_dgrad(const void* dy, const void* w, void* dx, const void* add, const FdBnEpilogue* bz, int N, int H, int W,
                     int C, int K, int variant, cudaStream_t s);
int fd_conv3x3_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int C, int K, int splits,
                     cudaStream_t s);
// stem: 7x7/2 conv as a 4x4/1 conv over the space-to-depth input (conv_halo_sm100.cu)
int fd_stem_fprop(const void* x2, const void* w2, void* y, float* sum, float* sumsq, int prezeroed, int N, int HS, int WS,
                  cudaStream_t s);
int fd_stem_wgrad(const void* dy, const void* x2, float* dw2, int N, int HS, int WS, cudaStream_t s);
