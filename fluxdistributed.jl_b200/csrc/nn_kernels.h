// Host entry points of the bandwidth-bound NHWC kernels (nn_kernels.cu).
#pragma once
#include <cuda_runtime.h>

int fd_bn_stats(const void* y, float* sum, float* sumsq, long M, int C, cudaStream_t s);
// `consumed` (optional): device counter; the last CTA that has read sum/sumsq zeroes them for the producer's next launch
int fd_bn_apply(const void* y, const void* res, void* out, float* sum, float* sumsq, const float* gamma,
                const float* beta, float* run_mean, float* run_var, float* save_mean, float* save_rstd, long M, int C,
                float eps, float momentum, int training, int relu, unsigned* consumed, cudaStream_t s);
// single-pass BN backward for a pre-masked gradient dz with sums s0 = sum dz*(y-mean), s1 = sum dz already reduced by the
// producer (igemm dz epilogue); clears s0/s1 through `consumed`
int fd_bn_bwd_apply(const void* dz, const void* y, const float* mean, const float* rstd, const float* gamma, float* s0,
                    float* s1, float* dgamma, float* dbeta, void* dy, long M, int C, unsigned* consumed, cudaStream_t s);
// BN (+ReLU) + 3x3/2/pad-1 max-pool of y[N,H,W,C] in one pass -> out[N,P,Q,C], idx (window-relative arg-max bytes)
int fd_bn_relu_pool(const void* y, void* out, void* idx, float* sum, float* sumsq, const float* gamma, const float* beta,
                    float* run_mean, float* run_var, float* save_mean, float* save_rstd, int N, int H, int W, int C,
                    float eps, float momentum, int training, unsigned* consumed, cudaStream_t s);
int fd_bn_bwd(const void* dout, const void* y, const void* out, const float* mean, const float* rstd,
              const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* dres, long M, int C,
              int relu, cudaStream_t s);
int fd_maxpool_fwd(const void* x, void* out, void* idx, int N, int H, int W, int C, int k, int stride, int pad,
                   cudaStream_t s);
int fd_maxpool_bwd_idx(const void* idx, const void* dout, void* dx, int N, int H, int W, int C, int k, int stride, int pad,
                       cudaStream_t s);
int fd_maxpool_bwd(const void* x, const void* out, const void* dout, void* dx, int N, int H, int W, int C, int k,
                   int stride, int pad, cudaStream_t s);
int fd_upsample_zero(const void* dy, void* up, int N, int P, int Q, int HU, int WU, int C, int stride, cudaStream_t s);
int fd_gap_fwd(const void* x, void* out, int N, int HW, int C, cudaStream_t s);
int fd_gap_bwd(const void* dout, void* dx, int N, int HW, int C, cudaStream_t s);
int fd_softmax_ce(const void* logits, const long long* labels, void* dlogits, float* row_loss, float* loss,
                  unsigned* counter, int B, int Cn, long ld, float scale, int is_bf16, cudaStream_t s);
int fd_normalize_u8(const void* in, void* out, long pixels, int Cpad, const float* mean, const float* stdv,
                    cudaStream_t s);
int fd_cast_f32_bf16(const float* in, void* out, long n, cudaStream_t s);
// direct (CUDA-core) NHWC convolution for channel counts the tensor-core tiles do not cover (direct_conv.cu)
int fd_direct_conv_fprop(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int C, int K, int R,
                         int S, int stride, int pad, cudaStream_t s);
int fd_direct_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int K, int R, int S, int stride,
                         int pad, cudaStream_t s);
int fd_direct_conv_wgrad(const void* dy, const void* x, float* dw, float* dbias, int N, int H, int W, int C, int K, int R,
                         int S, int stride, int pad, long sk, long sr, long ss, long sc, cudaStream_t s);
int fd_small_gemm(const void* A, const void* B, float* Cf, void* Cb, int M, int N, int K, long sam, long sak, long sbk,
                  long sbn, long ldc, cudaStream_t s);
int fd_bias_add(void* y, const float* bias, int M, int N, cudaStream_t s);
int fd_colsum(const void* d, float* out, int M, int N, cudaStream_t s);
int fd_s2d_pad(const void* x, void* x2, int N, int H, int W, int pad, cudaStream_t s);
int fd_stem_w_s2d(const void* w, void* w2, int K, cudaStream_t s);
int fd_stem_dw_from_s2d(const float* dw2, float* dw, int K, cudaStream_t s);
#define FD_BN_MAX_GRID 296
// workspace bytes for the single-launch BN kernels with at most `C` channels
// [barrier counter, generation, fwd-stats-consumed counter, bwd-sums-consumed counter] + 2 x [2][C] barrier-kernel
// accumulators + [2][C] forward epilogue statistics + [2][C] backward epilogue sums
void fd_bn_set_reserve(int ctas);
inline size_t fd_bn_ws_bytes(int C) { return 16 + (size_t)8 * C * sizeof(float); }
int fd_bn_fwd_fused(const void* y, const void* res, void* out, const float* gamma, const float* beta, float* run_mean,
                    float* run_var, float* save_mean, float* save_rstd, void* ws, long M, int C, float eps,
                    float momentum, int relu, cudaStream_t s);
int fd_bn_bwd_pool_fused(const void* idx, const void* g, const void* y, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* ws, int N,
                         int H, int W, int C, cudaStream_t s);
int fd_bn_bwd_fused(const void* dout, const void* y, const void* out, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* dres, void* ws,
                    long M, int C, int relu, cudaStream_t s);

// programmatic dependent launch switch (error.cu)
int fd_pdl_mask();
void fd_set_pdl(int mask);
