// Host entry points of the tcgen05 implicit-GEMM kernels (igemm_sm100.cu). All tensors bf16 NHWC / KRSC.
#pragma once
#include <cuda_runtime.h>

int fd_num_sms();
int fd_gemm_nt(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s);
int fd_conv_fprop(const void* x, const void* w, void* y, int N, int H, int W, int C, int K, int R, int S, int stride,
                  int pad, cudaStream_t s);
int fd_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int K, int R, int S, int pad,
                  cudaStream_t s);
int fd_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int C, int K, int R, int S,
                  int stride, int pad, int splits, int accumulate, cudaStream_t s);
