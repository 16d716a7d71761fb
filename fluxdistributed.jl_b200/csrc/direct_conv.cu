// Direct (CUDA-core) NHWC convolution for layers the tcgen05 tiles do not cover: channel counts that are not
// multiples of 64 (the reference's test network is Conv((7,7), 3=>3) -> flatten -> Dense(2028,10),
// test/single_device.jl:117-119).  Round 1 sent those to cuDNN; with these kernels every layer of every model in the
// zoo runs on this repository's own code.  They are written for correctness and small sizes, not for speed: one thread
// per output element (fprop, dgrad), one warp per filter element (wgrad).  bf16 in/out, fp32 accumulation.
#include "common.cuh"
#include "nn_kernels.h"

namespace {

__global__ void __launch_bounds__(256)
direct_fprop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ bias,
                    bf16* __restrict__ y, int N, int H, int W, int C, int K, int R, int S, int P, int Q, int stride, int pad) {
  const long total = (long)N * P * Q * K;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    long t = i / K;
    const int q = (int)(t % Q); t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float acc = bias ? bias[k] : 0.f;
    for (int r = 0; r < R; ++r) {
      const int h = p * stride + r - pad;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < S; ++s) {
        const int ww = q * stride + s - pad;
        if (ww < 0 || ww >= W) continue;
        const bf16* xp = x + (((long)n * H + h) * W + ww) * C;
        const bf16* wp = w + (((long)k * R + r) * S + s) * C;
        for (int c = 0; c < C; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
      }
    }
    y[i] = __float2bfloat16_rn(acc);
  }
}

__global__ void __launch_bounds__(256)
direct_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w, bf16* __restrict__ dx, int N, int H, int W,
                    int C, int K, int R, int S, int P, int Q, int stride, int pad) {
  const long total = (long)N * H * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int ww = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int hp = h + pad - r;
      if (hp < 0 || hp % stride) continue;
      const int p = hp / stride;
      if (p >= P) continue;
      for (int s = 0; s < S; ++s) {
        const int wq = ww + pad - s;
        if (wq < 0 || wq % stride) continue;
        const int q = wq / stride;
        if (q >= Q) continue;
        const bf16* dp = dy + (((long)n * P + p) * Q + q) * K;
        for (int k = 0; k < K; ++k) acc += __bfloat162float(dp[k]) * __bfloat162float(w[(((long)k * R + r) * S + s) * C + c]);
      }
    }
    dx[i] = __float2bfloat16_rn(acc);
  }
}

// one warp per filter element (k, r, s, c); dw addressed through explicit strides so the result can land in a
// [K,C,R,S]-contiguous or a channels_last ([K,R,S,C]) gradient buffer; accumulates (+=)
__global__ void __launch_bounds__(256)
direct_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, float* __restrict__ dw, float* __restrict__ dbias,
                    int N, int H, int W, int C, int K, int R, int S, int P, int Q, int stride, int pad, long sk, long sr,
                    long ss, long sc) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const long total = (long)K * R * S * C;
  const long npq = (long)N * P * Q;
  for (long e = warp; e < total + (dbias ? K : 0); e += nw) {
    float acc = 0.f;
    if (e < total) {
      const int c = (int)(e % C);
      long t = e / C;
      const int s = (int)(t % S); t /= S;
      const int r = (int)(t % R);
      const int k = (int)(t / R);
      for (long j = lane; j < npq; j += 32) {
        const int q = (int)(j % Q);
        long u = j / Q;
        const int p = (int)(u % P);
        const int n = (int)(u / P);
        const int h = p * stride + r - pad, ww = q * stride + s - pad;
        if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
        acc += __bfloat162float(dy[j * K + k]) * __bfloat162float(x[(((long)n * H + h) * W + ww) * C + c]);
      }
      acc = warp_sum(acc);
      if (lane == 0) dw[k * sk + r * sr + s * ss + c * sc] += acc;
    } else {  // bias gradient: column sums of dy
      const int k = (int)(e - total);
      for (long j = lane; j < npq; j += 32) acc += __bfloat162float(dy[j * K + k]);
      acc = warp_sum(acc);
      if (lane == 0) dbias[k] += acc;
    }
  }
}

inline int grid_of(long n, int per) {
  long g = (n + per - 1) / per;
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return (int)g;
}

}  // namespace

int fd_direct_conv_fprop(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int C, int K, int R,
                         int S, int stride, int pad, cudaStream_t s) {
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  direct_fprop_kernel<<<grid_of((long)N * P * Q * K, 256), 256, 0, s>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w), bias, reinterpret_cast<bf16*>(y), N, H, W, C, K, R,
      S, P, Q, stride, pad);
  FD_LAUNCH_OK();
  return 0;
}

int fd_direct_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int K, int R, int S, int stride,
                         int pad, cudaStream_t s) {
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  direct_dgrad_kernel<<<grid_of((long)N * H * W * C, 256), 256, 0, s>>>(
      reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(w), reinterpret_cast<bf16*>(dx), N, H, W, C, K, R, S,
      P, Q, stride, pad);
  FD_LAUNCH_OK();
  return 0;
}

int fd_direct_conv_wgrad(const void* dy, const void* x, float* dw, float* dbias, int N, int H, int W, int C, int K, int R,
                         int S, int stride, int pad, long sk, long sr, long ss, long sc, cudaStream_t s) {
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  const long warps = (long)K * R * S * C + (dbias ? K : 0);
  direct_wgrad_kernel<<<grid_of(warps * 32, 256), 256, 0, s>>>(
      reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), dw, dbias, N, H, W, C, K, R, S, P, Q, stride, pad,
      sk, sr, ss, sc);
  FD_LAUNCH_OK();
  return 0;
}
