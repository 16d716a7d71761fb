// Bandwidth-bound NHWC bf16 kernels of the ResNet/CIFAR path:
// BatchNorm statistics / apply(+residual,+ReLU) / backward, max-pool, global-average-pool,
// fused softmax-cross-entropy (loss + dlogits in one pass), uint8 -> bf16 input normalisation.
// All use 16-byte vector accesses along the channel (innermost) dimension and fp32 arithmetic.
#include <cstdlib>
#include "common.cuh"
#include "nn_kernels.h"

namespace {

struct alignas(16) V8 {
  uint32_t u[4];
};
FD_DEVINL void unpack8(const V8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = unpack_bf16x2(v.u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
FD_DEVINL V8 pack8(const float* f) {
  V8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.u[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return v;
}
// Flat-index split for kernels whose per-thread work is small (pooling): four 64-bit div/mod pairs by runtime values
// were ~250 of the ~390 instructions per thread of maxpool_bwd_idx (ncu: 118 M warp instructions for 9.6 M vectors).
// All tensors on the ResNet path have < 2^32 vectors, so the split runs in 32-bit unsigned arithmetic.
FD_DEVINL void split4(long i, int d0, int d1, int d2, int& r0, int& r1, int& r2, int& r3) {
  if ((unsigned long long)i < 0x100000000ull) {
    uint32_t t = (uint32_t)i;
    r0 = (int)(t % (uint32_t)d0); t /= (uint32_t)d0;
    r1 = (int)(t % (uint32_t)d1); t /= (uint32_t)d1;
    r2 = (int)(t % (uint32_t)d2);
    r3 = (int)(t / (uint32_t)d2);
  } else {
    long t = i;
    r0 = (int)(t % d0); t /= d0;
    r1 = (int)(t % d1); t /= d1;
    r2 = (int)(t % d2);
    r3 = (int)(t / d2);
  }
}
FD_DEVINL V8 ldg8(const bf16* p) { return *reinterpret_cast<const V8*>(p); }
FD_DEVINL void stg8(bf16* p, const V8& v) { *reinterpret_cast<V8*>(p) = v; }

// ---------------------------------------------------------------------------------------------
// BN statistics: per-channel sum and sum of squares over M rows of y[M, C].
// Block = 256 threads arranged as (256 / (C/8)) row-lanes x (C/8) channel-vectors.
// ---------------------------------------------------------------------------------------------
template <int NQ>  // number of quantities reduced per channel (2: sum,sumsq)
FD_DEVINL void block_channel_reduce(float (*acc)[8], int C8, int cv, int rl, int RL, float* smem, float* out0,
                                    float* out1) {
  // smem layout [RL][C8*8][NQ]
  const int Cc = C8 * 8;
  if (rl < RL) {
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) smem[(rl * Cc + cv * 8 + j) * NQ + q] = acc[q][j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cc * NQ; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < RL; ++r) s += smem[r * Cc * NQ + c];
    const int ch = c / NQ, q = c % NQ;
    atomicAdd((q == 0 ? out0 : out1) + ch, s);
  }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ y, float* __restrict__ sum,
                                                       float* __restrict__ sumsq, long M, int C) {
  extern __shared__ float smem[];
  const int C8 = C / 8;
  const int RL = 256 / C8;
  const int cv = threadIdx.x % C8, rl = threadIdx.x / C8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (rl < RL) {
    // each block owns a contiguous slab of rows; 4 independent 16-byte loads in flight per thread
    const long per = (M + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
    long r = r0 + rl;
    for (; r + 3L * RL < r1; r += 4L * RL) {
      V8 v0 = ldg8(y + r * C + cv * 8), v1 = ldg8(y + (r + RL) * C + cv * 8), v2 = ldg8(y + (r + 2L * RL) * C + cv * 8),
         v3 = ldg8(y + (r + 3L * RL) * C + cv * 8);
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(v0, f0); unpack8(v1, f1); unpack8(v2, f2); unpack8(v3, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        acc[1][j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
      }
    }
    for (; r < r1; r += RL) {
      float f[8];
      unpack8(ldg8(y + r * C + cv * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += f[j];
        acc[1][j] += f[j] * f[j];
      }
    }
  }
  block_channel_reduce<2>(acc, C8, cv, rl, RL, smem, sum, sumsq);
}

// BN apply: out = relu( (y-mean)*rstd*gamma + beta (+ residual) ).  mean/rstd derived from sum/sumsq
// (training) or running stats (eval). Block 0 also emits saved mean/rstd and updates running stats.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res, bf16* __restrict__ out,
                float* sum, float* sumsq, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ run_mean, float* __restrict__ run_var,
                float* __restrict__ save_mean, float* __restrict__ save_rstd, long M, int C, float eps,
                float momentum, int training, int relu, unsigned* consumed) {
  extern __shared__ float smem[];  // scale[C], shift[C]
  __shared__ int s_last;
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  float* sc = smem;
  float* sh = smem + C;
  const float invM = 1.f / (float)M;
  const int C8 = C / 8;
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int U = 4;  // independent 16-byte loads per tensor in flight per thread
  // The activation loads do not depend on the per-channel coefficients: the first batch is requested BEFORE the
  // coefficient prologue (dependent global loads, rsqrt, barrier, the `consumed` atomic), whose ~2 us of latency would
  // otherwise be fully exposed — the grid is a single wave, every CTA pays it exactly once (round 2).
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  V8 yv[U], rv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long ii = i0 + u * stride;
    if (ii < total) {
      yv[u] = ldg8(y + ii * 8);
      if (res) rv[u] = ldg8(res + ii * 8);
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, var;
    if (training) {
      mean = sum[c] * invM;
      var = fmaxf(sumsq[c] * invM - mean * mean, 0.f);
    } else {
      mean = run_mean[c];
      var = run_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float g = gamma[c];
    sc[c] = g * rstd;
    sh[c] = beta[c] - mean * g * rstd;
    if (blockIdx.x == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_rstd) save_rstd[c] = rstd;
      if (training && run_mean) {
        const float unbiased = var * ((float)M / fmaxf((float)M - 1.f, 1.f));
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
      }
    }
  }
  __syncthreads();
  if (consumed) {
    // the statistics were accumulated by the producing conv's epilogue into buffers that must be zero again before
    // its next launch: every CTA counts itself once it has read them; the last one clears them (no waiting involved)
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(consumed, 1u);
      s_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) sum[c] = sumsq[c] = 0.f;
      if (threadIdx.x == 0) *consumed = 0u;
    }
  }
  for (long i = i0; i < total; i += U * stride) {
    if (i != i0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long ii = i + u * stride;
        if (ii < total) {
          yv[u] = ldg8(y + ii * 8);
          if (res) rv[u] = ldg8(res + ii * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        const int ch = (int)(ii % C8) * 8;
        float f[8], r[8];
        unpack8(yv[u], f);
        if (res) unpack8(rv[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = f[j] * sc[ch + j] + sh[ch + j];
          if (res) v += r[j];
          if (relu) v = fmaxf(v, 0.f);
          f[j] = v;
        }
        stg8(out + ii * 8, pack8(f));
      }
    }
  }
}

// BN backward pass 1: dgamma[c] = sum dz*xhat, dbeta[c] = sum dz, with dz = dout * (out > 0) if relu
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const bf16* __restrict__ out,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int C,
                     int relu) {
  // relu mask: from the saved activation `out` when given (residual layers: the pre-activation includes the
  // skip input), otherwise recomputed from y as (gamma*xhat + beta > 0) — one tensor less to read.
  extern __shared__ float smem[];
  const int C8 = C / 8;
  const int RL = 256 / C8;
  const int cv = threadIdx.x % C8, rl = threadIdx.x / C8;
  float acc[2][8];
  float mu[8], rs[8], gsc[8], gsh[8];
  const bool from_y = relu && (out == nullptr);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[0][j] = acc[1][j] = 0.f;
    mu[j] = mean[cv * 8 + j];
    rs[j] = rstd[cv * 8 + j];
    gsc[j] = gamma[cv * 8 + j] * rs[j];
    gsh[j] = beta[cv * 8 + j] - mu[j] * gsc[j];
  }
  if (rl < RL) {
    const long per = (M + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
    long r = r0 + rl;
    for (; r + RL < r1; r += 2L * RL) {  // two rows (six 16-byte loads) in flight per thread
      const long o0 = r * C + cv * 8, o1 = (r + RL) * C + cv * 8;
      V8 a0 = ldg8(dout + o0), a1 = ldg8(dout + o1), b0 = ldg8(y + o0), b1 = ldg8(y + o1), c0, c1;
      if (relu && !from_y) {
        c0 = ldg8(out + o0);
        c1 = ldg8(out + o1);
      }
      float d0[8], d1[8], y0[8], y1[8], q0[8], q1[8];
      unpack8(a0, d0); unpack8(a1, d1); unpack8(b0, y0); unpack8(b1, y1);
      if (relu && !from_y) {
        unpack8(c0, q0);
        unpack8(c1, q1);
      }
      if (from_y) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          q0[j] = y0[j] * gsc[j] + gsh[j];
          q1[j] = y1[j] * gsc[j] + gsh[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z0 = (relu && !(q0[j] > 0.f)) ? 0.f : d0[j];
        const float z1 = (relu && !(q1[j] > 0.f)) ? 0.f : d1[j];
        acc[0][j] += (z0 * (y0[j] - mu[j]) + z1 * (y1[j] - mu[j])) * rs[j];
        acc[1][j] += z0 + z1;
      }
    }
    for (; r < r1; r += RL) {
      float d[8], yy[8], o[8];
      unpack8(ldg8(dout + r * C + cv * 8), d);
      unpack8(ldg8(y + r * C + cv * 8), yy);
      if (relu && !from_y) unpack8(ldg8(out + r * C + cv * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = d[j];
        const float pre = from_y ? (yy[j] * gsc[j] + gsh[j]) : o[j];
        if (relu && !(pre > 0.f)) dz = 0.f;
        acc[0][j] += dz * (yy[j] - mu[j]) * rs[j];
        acc[1][j] += dz;
      }
    }
  }
  block_channel_reduce<2>(acc, C8, cv, rl, RL, smem, dgamma, dbeta);
}

// BN backward pass 2: dy = gamma*rstd*(dz - dbeta/M - xhat*dgamma/M); optionally dres = dz
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const bf16* __restrict__ out,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                    bf16* __restrict__ dy, bf16* __restrict__ dres, long M, int C, int relu) {
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  extern __shared__ float smem[];  // a[C], b[C], c[C], sc[C], sh[C]:  dy = a*dz + b*y + c ; pre-act = sc*y + sh
  float* ka = smem;
  float* kb = smem + C;
  float* kc = smem + 2 * C;
  float* ksc = smem + 3 * C;
  float* ksh = smem + 4 * C;
  const bool from_y = relu && (out == nullptr);
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma[c], rs = rstd[c], mu = mean[c];
    const float dg = dgamma[c] * invM, db = dbeta[c] * invM;
    // dy = g*rs*(dz - db - (y-mu)*rs*dg) = g*rs*dz - g*rs*rs*dg*y + g*rs*(mu*rs*dg - db)
    ka[c] = g * rs;
    kb[c] = -g * rs * rs * dg;
    kc[c] = g * rs * (mu * rs * dg - db);
    ksc[c] = g * rs;
    ksh[c] = beta[c] - mu * g * rs;
  }
  __syncthreads();
  const int C8 = C / 8;
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const long i1 = i + stride;
    const bool has1 = i1 < total;
    V8 a0 = ldg8(dout + i * 8), b0 = ldg8(y + i * 8), a1, b1, c0, c1;
    if (has1) {
      a1 = ldg8(dout + i1 * 8);
      b1 = ldg8(y + i1 * 8);
    }
    if (relu && !from_y) {
      c0 = ldg8(out + i * 8);
      if (has1) c1 = ldg8(out + i1 * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has1) break;
      const long ii = u ? i1 : i;
      const int cv = (int)(ii % C8);
      float d[8], yy[8], o[8], dz[8];
      unpack8(u ? a1 : a0, d);
      unpack8(u ? b1 : b0, yy);
      if (relu && !from_y) unpack8(u ? c1 : c0, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = from_y ? (yy[j] * ksc[cv * 8 + j] + ksh[cv * 8 + j]) : o[j];
        dz[j] = (relu && !(pre > 0.f)) ? 0.f : d[j];
        d[j] = ka[cv * 8 + j] * dz[j] + kb[cv * 8 + j] * yy[j] + kc[cv * 8 + j];
      }
      stg8(dy + ii * 8, pack8(d));
      if (dres) stg8(dres + ii * 8, pack8(dz));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Single-launch BatchNorm forward / backward.  Most ResNet layers have small activations (5-40 MB) for which
// the 2-kernel + 2-memset formulation is launch/latency-bound (scripts/bn_bench.py: 10-35 us fixed cost per
// layer).  These kernels do both passes in one launch: phase 1 block-partial reductions added atomically into a
// per-channel accumulator -> grid barrier -> every CTA derives the per-channel coefficients -> phase 2 applies
// them (the second read of a small tensor is an L2 hit).  The accumulator is double-buffered by launch parity:
// a launch adds into buffer (gen & 1) and clears the other one for the next launch, so no memset is needed.
// All CTAs are co-resident (grid bounded by the occupancy query).
// ---------------------------------------------------------------------------------------------
struct GridSync {
  unsigned* counter;  // arrivals of this launch
  unsigned* gen;      // generation, bumped once per launch by the last arriver
};

FD_DEVINL void grid_barrier(const GridSync& g, unsigned gen0) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(g.counter, 1u);
    if (prev == gridDim.x - 1) {
      *g.counter = 0;
      __threadfence();
      atomicAdd(g.gen, 1u);
    } else {
      while (*reinterpret_cast<volatile unsigned*>(g.gen) == gen0) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

FD_DEVINL void block_accumulate2(float (*acc)[8], int C8, int cv, int rl, int RL, float* smem, float* accum) {
  // smem [RL][C][2] -> atomicAdd into accum[2][C]
  const int Cc = C8 * 8;
  if (rl < RL) {
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) smem[(rl * Cc + cv * 8 + j) * 2 + q] = acc[q][j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cc * 2; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < RL; ++r) s += smem[r * Cc * 2 + c];
    atomicAdd(accum + (long)(c & 1) * Cc + (c >> 1), s);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
bn_fwd_fused_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res, bf16* __restrict__ out,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ run_mean,
                    float* __restrict__ run_var, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                    float* __restrict__ accbuf, GridSync gs, long M, int C, float eps, float momentum, int relu) {
  extern __shared__ float smem[];
  __shared__ unsigned s_gen0;
  const int C8 = C / 8;
  const int RL = 256 / C8;
  const int cv = threadIdx.x % C8, rl = threadIdx.x / C8;
  if (threadIdx.x == 0) s_gen0 = *reinterpret_cast<volatile unsigned*>(gs.gen);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (rl < RL) {
    const long per = (M + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
    long r = r0 + rl;
    for (; r + 3L * RL < r1; r += 4L * RL) {
      V8 v0 = ldg8(y + r * C + cv * 8), v1 = ldg8(y + (r + RL) * C + cv * 8), v2 = ldg8(y + (r + 2L * RL) * C + cv * 8),
         v3 = ldg8(y + (r + 3L * RL) * C + cv * 8);
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(v0, f0); unpack8(v1, f1); unpack8(v2, f2); unpack8(v3, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        acc[1][j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
      }
    }
    for (; r < r1; r += RL) {
      float f[8];
      unpack8(ldg8(y + r * C + cv * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += f[j];
        acc[1][j] += f[j] * f[j];
      }
    }
  }
  __syncthreads();
  const unsigned gen0 = s_gen0;
  float* accum = accbuf + (size_t)(gen0 & 1) * 2 * C;
  float* other = accbuf + (size_t)((gen0 & 1) ^ 1) * 2 * C;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) other[c] = 0.f;
  block_accumulate2(acc, C8, cv, rl, RL, smem, accum);
  grid_barrier(gs, gen0);
  float* sc = smem;
  float* sh = smem + C;
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = __ldcg(accum + c) * invM;
    const float var = fmaxf(__ldcg(accum + C + c) * invM - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float g = gamma[c];
    sc[c] = g * rstd;
    sh[c] = beta[c] - mean * g * rstd;
    if (blockIdx.x == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_rstd) save_rstd[c] = rstd;
      if (run_mean) {
        const float unbiased = var * ((float)M / fmaxf((float)M - 1.f, 1.f));
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
      }
    }
  }
  __syncthreads();
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const long i1 = i + stride;
    const bool has1 = i1 < total;
    V8 y0 = ldg8(y + i * 8), y1, q0, q1;
    if (has1) y1 = ldg8(y + i1 * 8);
    if (res) {
      q0 = ldg8(res + i * 8);
      if (has1) q1 = ldg8(res + i1 * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has1) break;
      const long ii = u ? i1 : i;
      const int ch = (int)(ii % C8);
      float f[8], r[8];
      unpack8(u ? y1 : y0, f);
      if (res) unpack8(u ? q1 : q0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = f[j] * sc[ch * 8 + j] + sh[ch * 8 + j];
        if (res) v += r[j];
        if (relu) v = fmaxf(v, 0.f);
        f[j] = v;
      }
      stg8(out + ii * 8, pack8(f));
    }
  }
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_fused_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const bf16* __restrict__ out,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    bf16* __restrict__ dy, bf16* __restrict__ dres, float* __restrict__ accbuf, GridSync gs, long M, int C,
                    int relu) {
  // Both passes are latency-bound unless many 16-byte loads are in flight per thread: 4 rows x (dout, y[, out]) are
  // requested before the first use (8-12 outstanding loads per thread, 2 CTAs of 256 threads per SM).
  constexpr int U = 4;
  extern __shared__ float smem[];
  __shared__ unsigned s_gen0;
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  const int C8 = C / 8;
  const int RL = 256 / C8;
  const int cv = threadIdx.x % C8, rl = threadIdx.x / C8;
  const bool from_y = relu && (out == nullptr);
  const bool use_out = relu && !from_y;
  if (threadIdx.x == 0) s_gen0 = *reinterpret_cast<volatile unsigned*>(gs.gen);
  float acc[2][8];  // sum dz*(y-mean), sum dz   (the rstd factor is applied after the reduction)
  float mu[8], gsc[8], gsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[0][j] = acc[1][j] = 0.f;
    mu[j] = mean[cv * 8 + j];
    gsc[j] = gamma[cv * 8 + j] * rstd[cv * 8 + j];
    gsh[j] = beta[cv * 8 + j] - mu[j] * gsc[j];
  }
  if (rl < RL) {
    const long per = (M + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
    for (long r = r0 + rl; r < r1; r += (long)U * RL) {
      V8 a[U], b[U], c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long rr = r + (long)u * RL;
        if (rr < r1) {
          const long o = rr * C + cv * 8;
          a[u] = ldg8(dout + o);
          b[u] = ldg8(y + o);
          if (use_out) c[u] = ldg8(out + o);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (long)u * RL < r1) {
          float d[8], yy[8], o[8];
          unpack8(a[u], d);
          unpack8(b[u], yy);
          if (use_out) unpack8(c[u], o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pre = from_y ? (yy[j] * gsc[j] + gsh[j]) : (use_out ? o[j] : 1.f);
            const float dz = (relu && !(pre > 0.f)) ? 0.f : d[j];
            acc[0][j] += dz * (yy[j] - mu[j]);
            acc[1][j] += dz;
          }
        }
      }
    }
  }
  __syncthreads();
  const unsigned gen0 = s_gen0;
  float* accum = accbuf + (size_t)(gen0 & 1) * 2 * C;
  float* other = accbuf + (size_t)((gen0 & 1) ^ 1) * 2 * C;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) other[c] = 0.f;
  block_accumulate2(acc, C8, cv, rl, RL, smem, accum);
  grid_barrier(gs, gen0);
  float* ka = smem;
  float* kb = smem + C;
  float* kc = smem + 2 * C;
  float* ksh = smem + 3 * C;
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma[c], r_ = rstd[c], m_ = mean[c];
    const float dg = __ldcg(accum + c) * r_, db = __ldcg(accum + C + c);
    if (blockIdx.x == 0) {
      dgamma[c] = dg;
      dbeta[c] = db;
    }
    ka[c] = g * r_;
    kb[c] = -g * r_ * r_ * dg * invM;
    kc[c] = g * r_ * (m_ * r_ * dg * invM - db * invM);
    ksh[c] = beta[c] - m_ * g * r_;
  }
  __syncthreads();
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += U * stride) {
    V8 a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        a[u] = ldg8(dout + ii * 8);
        b[u] = ldg8(y + ii * 8);
        if (use_out) c[u] = ldg8(out + ii * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        const int ch = (int)(ii % C8) * 8;
        float d[8], yy[8], o[8], dz[8];
        unpack8(a[u], d);
        unpack8(b[u], yy);
        if (use_out) unpack8(c[u], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = from_y ? (yy[j] * ka[ch + j] + ksh[ch + j]) : (use_out ? o[j] : 1.f);
          dz[j] = (relu && !(pre > 0.f)) ? 0.f : d[j];
          d[j] = ka[ch + j] * dz[j] + kb[ch + j] * yy[j] + kc[ch + j];
        }
        stg8(dy + ii * 8, pack8(d));
        if (dres) stg8(dres + ii * 8, pack8(dz));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem tail backward in one launch: max-pool backward (arg-max gather) + ReLU mask + BN backward (round 2).
// The forward fuses BN + ReLU + 3x3/2 max-pool (bn_relu_pool3s2_kernel); its backward used to be maxpool_bwd_idx
// (writes the full-resolution gradient, 154 MB at batch 96) followed by bn_bwd_fused (reads it twice).  Here the
// full-resolution gradient is never materialised: both passes of the grid-barrier BN backward rebuild it on the fly
// from the pooled gradient and the arg-max bytes (<= 2x2 windows of 24 bytes each per output vector, L1/L2 hits).
// ---------------------------------------------------------------------------------------------
FD_DEVINL V8 pool3s2_gather(const uint8_t* __restrict__ idx, const bf16* __restrict__ g, int n, int h, int w, int cv, int C8,
                            int P, int Q) {
  // windows (p, q) of the 3x3 / stride-2 / pad-1 pool that contain pixel (h, w): p in [ceil((h-1)/2), floor((h+1)/2)]
  const int p_lo = max(0, h >> 1), p_hi = min(P - 1, (h + 1) >> 1);
  const int q_lo = max(0, w >> 1), q_hi = min(Q - 1, (w + 1) >> 1);
  uint2 pk[4];
  V8 dv[4];
  uint32_t pos[4];
  bool ok[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = p_lo + (u >> 1), q = q_lo + (u & 1);
    ok[u] = (p <= p_hi) && (q <= q_hi);
    if (ok[u]) {
      const long oi = (((long)n * P + p) * Q + q) * C8 + cv;
      pos[u] = (uint32_t)((h - (p * 2 - 1)) * 3 + (w - (q * 2 - 1)));
      pk[u] = *reinterpret_cast<const uint2*>(idx + oi * 8);
      dv[u] = ldg8(g + oi * 8);
    }
  }
  __nv_bfloat162 acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (!ok[u]) continue;
    const uint32_t rep = pos[u] * 0x01010101u;
    const uint32_t m_lo = __vcmpeq4(pk[u].x, rep), m_hi = __vcmpeq4(pk[u].y, rep);  // 0xFF per matching channel
    const uint32_t w0 = dv[u].u[0] & __byte_perm(m_lo, 0, 0x1100), w1 = dv[u].u[1] & __byte_perm(m_lo, 0, 0x3322);
    const uint32_t w2 = dv[u].u[2] & __byte_perm(m_hi, 0, 0x1100), w3 = dv[u].u[3] & __byte_perm(m_hi, 0, 0x3322);
    acc[0] = __hadd2(acc[0], *reinterpret_cast<const __nv_bfloat162*>(&w0));
    acc[1] = __hadd2(acc[1], *reinterpret_cast<const __nv_bfloat162*>(&w1));
    acc[2] = __hadd2(acc[2], *reinterpret_cast<const __nv_bfloat162*>(&w2));
    acc[3] = __hadd2(acc[3], *reinterpret_cast<const __nv_bfloat162*>(&w3));
  }
  V8 o;
#pragma unroll
  for (int e = 0; e < 4; ++e) o.u[e] = *reinterpret_cast<const uint32_t*>(&acc[e]);
  return o;
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_pool_fused_kernel(const uint8_t* __restrict__ idx, const bf16* __restrict__ g, const bf16* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         bf16* __restrict__ dy, float* __restrict__ accbuf, GridSync gs, int N, int H, int W, int C, int P,
                         int Q) {
  constexpr int U = 2;
  extern __shared__ float smem[];
  __shared__ unsigned s_gen0;
  const int C8 = C / 8;
  const int RL = 256 / C8;
  const int cv = threadIdx.x % C8, rl = threadIdx.x / C8;
  const long M = (long)N * H * W;
  if (threadIdx.x == 0) s_gen0 = *reinterpret_cast<volatile unsigned*>(gs.gen);
  float acc[2][8];  // sum dz*(y-mean), sum dz
  float mu[8], gsc[8], gsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[0][j] = acc[1][j] = 0.f;
    mu[j] = mean[cv * 8 + j];
    gsc[j] = gamma[cv * 8 + j] * rstd[cv * 8 + j];
    gsh[j] = beta[cv * 8 + j] - mu[j] * gsc[j];
  }
  if (rl < RL) {
    const long per = (M + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
    for (long r = r0 + rl; r < r1; r += (long)U * RL) {
      V8 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long rr = r + (long)u * RL;
        if (rr < r1) {
          const uint32_t t = (uint32_t)rr;
          const int w = (int)(t % (uint32_t)W), h = (int)((t / (uint32_t)W) % (uint32_t)H), n = (int)(t / (uint32_t)(W * H));
          b[u] = ldg8(y + rr * C + cv * 8);
          a[u] = pool3s2_gather(idx, g, n, h, w, cv, C8, P, Q);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (long)u * RL < r1) {
          float d[8], yy[8];
          unpack8(a[u], d);
          unpack8(b[u], yy);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pre = yy[j] * gsc[j] + gsh[j];
            const float dz = (pre > 0.f) ? d[j] : 0.f;
            acc[0][j] += dz * (yy[j] - mu[j]);
            acc[1][j] += dz;
          }
        }
      }
    }
  }
  __syncthreads();
  const unsigned gen0 = s_gen0;
  float* accum = accbuf + (size_t)(gen0 & 1) * 2 * C;
  float* other = accbuf + (size_t)((gen0 & 1) ^ 1) * 2 * C;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) other[c] = 0.f;
  block_accumulate2(acc, C8, cv, rl, RL, smem, accum);
  grid_barrier(gs, gen0);
  float* ka = smem;
  float* kb = smem + C;
  float* kc = smem + 2 * C;
  float* ksh = smem + 3 * C;
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g_ = gamma[c], r_ = rstd[c], m_ = mean[c];
    const float dg = __ldcg(accum + c) * r_, db = __ldcg(accum + C + c);
    if (blockIdx.x == 0) {
      dgamma[c] = dg;
      dbeta[c] = db;
    }
    ka[c] = g_ * r_;
    kb[c] = -g_ * r_ * r_ * dg * invM;
    kc[c] = g_ * r_ * (m_ * r_ * dg * invM - db * invM);
    ksh[c] = beta[c] - m_ * g_ * r_;
  }
  __syncthreads();
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += U * stride) {
    V8 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        int c_, w, h, n;
        split4(ii, C8, W, H, c_, w, h, n);
        b[u] = ldg8(y + ii * 8);
        a[u] = pool3s2_gather(idx, g, n, h, w, c_, C8, P, Q);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        const int ch = (int)(ii % C8) * 8;
        float d[8], yy[8];
        unpack8(a[u], d);
        unpack8(b[u], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = yy[j] * ka[ch + j] + ksh[ch + j];
          const float dz = (pre > 0.f) ? d[j] : 0.f;
          d[j] = ka[ch + j] * dz + kb[ch + j] * yy[j] + kc[ch + j];
        }
        stg8(dy + ii * 8, pack8(d));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Second (and only remaining) pass of the BN backward when the producer of the output gradient already applied the
// ReLU mask and reduced sum(dz), sum(dz*(y-mean)) in its epilogue (igemm "dz epilogue"):
//   dy = g*rstd * dz - g*rstd^2*dgamma/M * (y-mean)... folded into  ka*dz + kb*y + kc.
// No grid barrier, no second read of the inputs.  The last CTA to have read the accumulators clears them.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, float* s0, float* s1,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, bf16* __restrict__ dy, long M, int C,
                    unsigned* consumed) {
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  extern __shared__ float smem[];
  __shared__ int s_last;
  float* ka = smem;
  float* kb = smem + C;
  float* kc = smem + 2 * C;
  const float invM = 1.f / (float)M;
  constexpr int U = 4;
  const int C8 = C / 8;
  const long total = M * C8;
  const long stride = (long)gridDim.x * blockDim.x;
  // first batch of loads before the coefficient prologue (see bn_apply_kernel)
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  V8 a[U], b[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long ii = i0 + u * stride;
    if (ii < total) {
      a[u] = ldg8(dz + ii * 8);
      b[u] = ldg8(y + ii * 8);
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma[c], r_ = rstd[c], m_ = mean[c];
    const float dg = s0[c] * r_, db = s1[c];
    if (blockIdx.x == 0) {
      dgamma[c] = dg;
      dbeta[c] = db;
    }
    ka[c] = g * r_;
    kb[c] = -g * r_ * r_ * dg * invM;
    kc[c] = g * r_ * (m_ * r_ * dg * invM - db * invM);
  }
  __syncthreads();
  if (consumed) {
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(consumed, 1u);
      s_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) s0[c] = s1[c] = 0.f;
      if (threadIdx.x == 0) *consumed = 0u;
    }
  }
  for (long i = i0; i < total; i += U * stride) {
    if (i != i0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long ii = i + u * stride;
        if (ii < total) {
          a[u] = ldg8(dz + ii * 8);
          b[u] = ldg8(y + ii * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long ii = i + u * stride;
      if (ii < total) {
        const int ch = (int)(ii % C8) * 8;
        float d[8], yy[8];
        unpack8(a[u], d);
        unpack8(b[u], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = ka[ch + j] * d[j] + kb[ch + j] * yy[j] + kc[ch + j];
        stg8(dy + ii * 8, pack8(d));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem tail in one pass: BatchNorm (statistics from the conv epilogue) + ReLU + 3x3/2/pad-1 max-pool.
// relu(sc*y+sh) is monotone in y, so the pool runs on the RAW conv output (sign-flipped where sc < 0) with packed
// bf16x2 max; the arg-max is recovered by SIMD equality against the stored taps; only the pooled value goes through
// the affine + ReLU.  The 154 MB full-resolution activation of ResNet's stem is never written or re-read.
// Output: pooled [N,P,Q,C] bf16 and window-relative arg-max bytes (same format as maxpool_fwd_kernel).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_relu_pool3s2_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, uint8_t* __restrict__ idx, float* sum,
                       float* sumsq, const float* __restrict__ gamma, const float* __restrict__ beta,
                       float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ save_mean,
                       float* __restrict__ save_rstd, int N, int H, int W, int C, int P, int Q, float eps, float momentum,
                       int training, unsigned* consumed) {
  extern __shared__ float smem[];  // scale[C], shift[C]
  __shared__ int s_last;
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  float* sc = smem;
  float* sh = smem + C;
  const long M = (long)N * H * W;
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, var;
    if (training) {
      mean = sum[c] * invM;
      var = fmaxf(sumsq[c] * invM - mean * mean, 0.f);
    } else {
      mean = run_mean[c];
      var = run_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float g = gamma[c];
    sc[c] = g * rstd;
    sh[c] = beta[c] - mean * g * rstd;
    if (blockIdx.x == 0) {
      if (save_mean) save_mean[c] = mean;
      if (save_rstd) save_rstd[c] = rstd;
      if (training && run_mean) {
        const float unbiased = var * ((float)M / fmaxf((float)M - 1.f, 1.f));
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
      }
    }
  }
  __syncthreads();
  if (consumed) {  // see bn_apply_kernel: the last CTA to have read the statistics clears them
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(consumed, 1u);
      s_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) sum[c] = sumsq[c] = 0.f;
      if (threadIdx.x == 0) *consumed = 0u;
    }
  }
  const int C8 = C / 8;
  const long total = (long)N * P * Q * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, q, p, n;
    split4(i, C8, Q, P, cv, q, p, n);
    // sign-flip mask: channels with a negative scale are pooled by their minimum
    uint32_t sm[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      sm[e] = (sc[cv * 8 + 2 * e] < 0.f ? 0x8000u : 0u) | (sc[cv * 8 + 2 * e + 1] < 0.f ? 0x80000000u : 0u);
    uint32_t wv[9][4];
    bool ok[9];
    const int h0 = 2 * p - 1, w0 = 2 * q - 1;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        const int h = h0 + r, w = w0 + c3;
        const int tp = r * 3 + c3;
        ok[tp] = (h >= 0) && (h < H) && (w >= 0) && (w < W);
        if (ok[tp]) {
          const V8 v = ldg8(y + (((long)n * H + h) * W + w) * C + cv * 8);
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[tp][e] = v.u[e] ^ sm[e];
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[tp][e] = 0xff80ff80u;  // (-inf, -inf)
        }
      }
    __nv_bfloat162 mx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) mx[e] = *reinterpret_cast<const __nv_bfloat162*>(&wv[0][e]);
#pragma unroll
    for (int tp = 1; tp < 9; ++tp)
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e] = __hmax2(mx[e], *reinterpret_cast<const __nv_bfloat162*>(&wv[tp][e]));
    // first tap (scan order) whose value equals the maximum: walk the taps backwards, overwriting on equality
    uint32_t am[4] = {0x00ff00ffu, 0x00ff00ffu, 0x00ff00ffu, 0x00ff00ffu};
#pragma unroll
    for (int tp = 8; tp >= 0; --tp) {
      if (!ok[tp]) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t eq = __vcmpeq2(wv[tp][e], *reinterpret_cast<const uint32_t*>(&mx[e]));  // 0xFFFF per equal half
        am[e] = (am[e] & ~eq) | ((uint32_t)tp * 0x00010001u & eq);
      }
    }
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&mx[e]) ^ sm[e]);
      f[2 * e] = fmaxf(v2.x * sc[cv * 8 + 2 * e] + sh[cv * 8 + 2 * e], 0.f);
      f[2 * e + 1] = fmaxf(v2.y * sc[cv * 8 + 2 * e + 1] + sh[cv * 8 + 2 * e + 1], 0.f);
    }
    uint2 pk;
    pk.x = (am[0] & 0xffu) | ((am[0] >> 8) & 0xff00u) | ((am[1] & 0xffu) << 16) | ((am[1] & 0xff0000u) << 8);
    pk.y = (am[2] & 0xffu) | ((am[2] >> 8) & 0xff00u) | ((am[3] & 0xffu) << 16) | ((am[3] & 0xff0000u) << 8);
    *reinterpret_cast<uint2*>(idx + i * 8) = pk;
    stg8(out + i * 8, pack8(f));
  }
}

// ---------------------------------------------------------------------------------------------
// Max-pool (k x k, stride, pad) NHWC, 8 channels per thread; backward recomputes the arg-max
// (first maximum in scan order, matching PyTorch) instead of storing indices.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, uint8_t* __restrict__ idx, int N, int H, int W,
                   int C, int P, int Q, int k, int stride, int pad) {
  const int C8 = C / 8;
  const long total = (long)N * P * Q * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, q, p, n;
    split4(i, C8, Q, P, cv, q, p, n);
    float m[8];
    uint32_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m[j] = -INFINITY;
      am[j] = 255u;
    }
    for (int r = 0; r < k; ++r) {
      const int h = p * stride - pad + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = q * stride - pad + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ldg8(x + (((long)n * H + h) * W + w) * C + cv * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > m[j]) {  // strict: the first maximum in scan order wins, like PyTorch
            m[j] = f[j];
            am[j] = (uint32_t)(r * k + s);
          }
      }
    }
    if (idx) {  // one byte per element: window-relative arg-max, consumed by the atomic-free backward
      uint2 pk;
      pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = pk;
    }
    stg8(out + i * 8, pack8(m));
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout, bf16* __restrict__ dx, int N, int H,
                   int W, int C, int P, int Q, int k, int stride, int pad) {
  // one thread per (output window, 8 channels): recompute the arg-max (first maximum in scan order, like
  // PyTorch) and scatter the window's gradient there with native bf16x2 atomics; dx is pre-zeroed.
  const int C8 = C / 8;
  const long total = (long)N * P * Q * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, q, p, n;
    split4(i, C8, Q, P, cv, q, p, n);
    float m[8];
    int am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m[j] = -INFINITY;
      am[j] = -1;
    }
    for (int r = 0; r < k; ++r) {
      const int h = p * stride - pad + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = q * stride - pad + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ldg8(x + (((long)n * H + h) * W + w) * C + cv * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > m[j]) {
            m[j] = f[j];
            am[j] = h * W + w;
          }
      }
    }
    float d[8];
    unpack8(ldg8(dout + i * 8), d);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      bf16* base = dx + (long)n * H * W * C + cv * 8 + j;
      if (am[j] == am[j + 1]) {
        if (am[j] >= 0)
          atomicAdd(reinterpret_cast<bf162*>(base + (long)am[j] * C), __floats2bfloat162_rn(d[j], d[j + 1]));
      } else {
        if (am[j] >= 0) atomicAdd(reinterpret_cast<bf162*>(base + (long)am[j] * C), __floats2bfloat162_rn(d[j], 0.f));
        if (am[j + 1] >= 0)
          atomicAdd(reinterpret_cast<bf162*>(base + (long)am[j + 1] * C), __floats2bfloat162_rn(0.f, d[j + 1]));
      }
    }
  }
}

// atomic-free max-pool backward: one thread per (input pixel, 8 channels) gathers from the <= ceil(k/stride)^2 windows
// that contain it, using the arg-max bytes written by the forward kernel.
// KT/ST/PT > 0: window geometry known at compile time (3/2/1 for ResNet) so the window-range divisions fold to shifts
template <int KT, int ST, int PT>
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const uint8_t* __restrict__ idx, const bf16* __restrict__ dout, bf16* __restrict__ dx, int N, int H,
                       int W, int C, int P, int Q, int k_rt, int stride_rt, int pad_rt) {
  const int k = KT ? KT : k_rt, stride = KT ? ST : stride_rt, pad = KT ? PT : pad_rt;
  const int C8 = C / 8;
  const long total = (long)N * H * W * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, w, h, n;
    split4(i, C8, W, H, cv, w, h, n);
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    const int p_lo = max(0, (h + pad - k + stride) / stride), p_hi = min(P - 1, (h + pad) / stride);
    const int q_lo = max(0, (w + pad - k + stride) / stride), q_hi = min(Q - 1, (w + pad) / stride);
    if (p_hi - p_lo <= 1 && q_hi - q_lo <= 1) {
      // common case (k <= 2*stride): at most 2x2 windows.  All loads are issued first; the arg-max bytes are then
      // compared 4 at a time (SIMD-in-a-word), the byte mask is widened to bf16 lanes with PRMT and the selected
      // gradients are summed with packed bf16 adds: ~14 instructions per window instead of ~130 scalar ones (the
      // kernel was issue-bound, not memory-bound).
      uint2 pk[4];
      V8 dv[4];
      uint32_t pos[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p_lo + (u >> 1), q = q_lo + (u & 1);
        ok[u] = (p <= p_hi) && (q <= q_hi);
        if (ok[u]) {
          const long oi = (((long)n * P + p) * Q + q) * C8 + cv;
          pos[u] = (uint32_t)((h - (p * stride - pad)) * k + (w - (q * stride - pad)));
          pk[u] = *reinterpret_cast<const uint2*>(idx + oi * 8);
          dv[u] = ldg8(dout + oi * 8);
        }
      }
      __nv_bfloat162 acc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const uint32_t rep = pos[u] * 0x01010101u;
        const uint32_t m_lo = __vcmpeq4(pk[u].x, rep), m_hi = __vcmpeq4(pk[u].y, rep);  // 0xFF per matching channel
        const uint32_t w0 = dv[u].u[0] & __byte_perm(m_lo, 0, 0x1100), w1 = dv[u].u[1] & __byte_perm(m_lo, 0, 0x3322);
        const uint32_t w2 = dv[u].u[2] & __byte_perm(m_hi, 0, 0x1100), w3 = dv[u].u[3] & __byte_perm(m_hi, 0, 0x3322);
        acc[0] = __hadd2(acc[0], *reinterpret_cast<const __nv_bfloat162*>(&w0));
        acc[1] = __hadd2(acc[1], *reinterpret_cast<const __nv_bfloat162*>(&w1));
        acc[2] = __hadd2(acc[2], *reinterpret_cast<const __nv_bfloat162*>(&w2));
        acc[3] = __hadd2(acc[3], *reinterpret_cast<const __nv_bfloat162*>(&w3));
      }
      V8 o;
#pragma unroll
      for (int e = 0; e < 4; ++e) o.u[e] = *reinterpret_cast<const uint32_t*>(&acc[e]);
      stg8(dx + i * 8, o);
      continue;
    } else {
      for (int p = p_lo; p <= p_hi; ++p)
        for (int q = q_lo; q <= q_hi; ++q) {
          const long oi = (((long)n * P + p) * Q + q) * C8 + cv;
          const uint32_t pos = (uint32_t)((h - (p * stride - pad)) * k + (w - (q * stride - pad)));
          const uint2 pk = *reinterpret_cast<const uint2*>(idx + oi * 8);
          float d[8];
          unpack8(ldg8(dout + oi * 8), d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = ((j < 4 ? pk.x : pk.y) >> ((j & 3) * 8)) & 255u;
            if (a == pos) g[j] += d[j];
          }
        }
    }
    stg8(dx + i * 8, pack8(g));
  }
}

// zero-insertion upsample for strided dgrad: up[n, st*p, st*q, :] = dy[n,p,q,:], zeros elsewhere (one pass)
__global__ void __launch_bounds__(256)
upsample_zero_kernel(const bf16* __restrict__ dy, bf16* __restrict__ up, int N, int P, int Q, int HU, int WU, int C,
                     int stride) {
  const int C8 = C / 8;
  const long total = (long)N * HU * WU * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, w, h, n;
    split4(i, C8, WU, HU, cv, w, h, n);
    V8 v;
    v.u[0] = v.u[1] = v.u[2] = v.u[3] = 0u;
    if (h % stride == 0 && w % stride == 0 && h / stride < P && w / stride < Q)
      v = ldg8(dy + ((((long)n * P + h / stride) * Q + w / stride) * C8 + cv) * 8);
    stg8(up + i * 8, v);
  }
}

// Global average pool: x[N, HW, C] -> out[N, C] (bf16), and its backward (broadcast / HW)
__global__ void __launch_bounds__(256) gap_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int HW, int C) {
  const int C8 = C / 8;
  const long total = (long)N * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % C8);
    const int n = (int)(i / C8);
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      float f[8];
      unpack8(ldg8(x + ((long)n * HW + p) * C + cv * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= inv;
    stg8(out + i * 8, pack8(a));
  }
}
__global__ void __launch_bounds__(256) gap_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ dx, int N, int HW, int C) {
  const int C8 = C / 8;
  const long total = (long)N * HW * C8;
  const float inv = 1.f / (float)HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % C8);
    const int n = (int)(i / ((long)HW * C8));
    float f[8];
    unpack8(ldg8(dout + ((long)n * C8 + cv) * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    stg8(dx + i * 8, pack8(f));
  }
}

// ---------------------------------------------------------------------------------------------
// Fused softmax cross-entropy, integer labels: one warp per row; emits per-row loss, the mean loss
// (last-block reduction) and dlogits = (softmax - onehot) * scale in the logits dtype.
// ---------------------------------------------------------------------------------------------
template <typename T>
FD_DEVINL float to_f(T v);
template <>
FD_DEVINL float to_f<float>(float v) { return v; }
template <>
FD_DEVINL float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T>
FD_DEVINL T from_f(float v);
template <>
FD_DEVINL float from_f<float>(float v) { return v; }
template <>
FD_DEVINL bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const T* __restrict__ logits, const long long* __restrict__ labels, T* __restrict__ dlogits,
                  float* __restrict__ row_loss, float* __restrict__ loss, unsigned* __restrict__ counter, int B,
                  int Cn, long ld, float scale) {
  const int warps = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row = blockIdx.x * warps + warp;
  if (row < B) {
    const T* x = logits + (long)row * ld;
    float mx = -INFINITY;
    for (int c = lane; c < Cn; c += 32) mx = fmaxf(mx, to_f<T>(x[c]));
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < Cn; c += 32) se += __expf(to_f<T>(x[c]) - mx);
    se = warp_sum(se);
    const int y = (int)labels[row];
    const float lse = __logf(se) + mx;
    if (lane == 0) row_loss[row] = lse - to_f<T>(x[y]);
    const float inv = 1.f / se;
    T* d = dlogits + (long)row * ld;
    for (int c = lane; c < Cn; c += 32) {
      float p = __expf(to_f<T>(x[c]) - mx) * inv;
      if (c == y) p -= 1.f;
      d[c] = from_f<T>(p * scale);
    }
  }
  // mean over rows: last block to finish sums row_loss deterministically
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) s += row_loss[r];
    s = warp_sum(s);
    __shared__ float part[8];
    if (lane == 0) part[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < warps; ++w) tot += part[w];
      *loss = tot / (float)B;
      *counter = 0;
    }
  }
}

// uint8 NHWC [N,H,W,3] -> bf16 NHWC [N,H,W,Cpad] normalised with ImageNet mean/std; extra channels zero
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, bf16* __restrict__ out, long pixels, int Cpad, float m0, float m1,
                    float m2, float s0, float s1, float s2) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long)gridDim.x * blockDim.x) {
    const uint8_t* p = in + i * 3;
    float f[8] = {((float)p[0] * (1.f / 255.f) - m0) * s0, ((float)p[1] * (1.f / 255.f) - m1) * s1,
                  ((float)p[2] * (1.f / 255.f) - m2) * s2, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (Cpad == 8) {
      stg8(out + i * 8, pack8(f));
    } else {
      for (int c = 0; c < Cpad; ++c) out[i * Cpad + c] = __float2bfloat16_rn(c < 3 ? f[c] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem helpers: the 7x7/stride-2/pad-3 conv over 3 channels is run as a 4x4/stride-1 conv over the 2x2
// space-to-depth transform of the padded image (12 real + 4 zero channels = 32 B per pixel), which gives the
// tensor cores K = 16 taps x 16 channels instead of a 3-channel gather.
//   X2[n, i, j, (di*2+dj)*3 + c] = xpad[n, 2i+di, 2j+dj, c],  xpad = x zero-padded by 3
//   W2[k, a, b, (di*2+dj)*3 + c] = w[k, 2a+di, 2b+dj, c]  (zero where 2a+di or 2b+dj == 7)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
s2d_pad_kernel(const bf16* __restrict__ x, bf16* __restrict__ x2, int N, int H, int W, int HS, int WS, int pad) {
  const long total = (long)N * HS * WS;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int j = (int)(t % WS);
    long r = t / WS;
    const int i = (int)(r % HS);
    const int n = (int)(r / HS);
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = 0.f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        const int h = 2 * i + di - pad, w = 2 * j + dj - pad;
        if (h >= 0 && h < H && w >= 0 && w < W) {
          const bf16* px = x + (((long)n * H + h) * W + w) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[(di * 2 + dj) * 3 + c] = __bfloat162float(px[c]);
        }
      }
    stg8(x2 + t * 16, pack8(f));
    stg8(x2 + t * 16 + 8, pack8(f + 8));
  }
}

__global__ void stem_w_s2d_kernel(const bf16* __restrict__ w, bf16* __restrict__ w2, int K) {
  // w: [K,7,7,3] (KRSC) -> w2: [K,4,4,16]
  const int total = K * 256;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int e = t & 15, b = (t >> 4) & 3, a = (t >> 6) & 3, k = t >> 8;
    float v = 0.f;
    if (e < 12) {
      const int di = e / 6, dj = (e / 3) & 1, c = e % 3;
      const int r = 2 * a + di, s = 2 * b + dj;
      if (r < 7 && s < 7) v = __bfloat162float(w[((k * 7 + r) * 7 + s) * 3 + c]);
    }
    w2[t] = __float2bfloat16_rn(v);
  }
}

__global__ void stem_dw_from_s2d_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int K) {
  // dw[k,r,s,c] += dw2[k, r/2, s/2, ((r&1)*2 + (s&1))*3 + c]
  const int total = K * 147;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int c = t % 3, s = (t / 3) % 7, r = (t / 21) % 7, k = t / 147;
    dw[t] += dw2[((k * 4 + (r >> 1)) * 4 + (s >> 1)) * 16 + ((r & 1) * 2 + (s & 1)) * 3 + c];
  }
}

// fp32 -> bf16 cast of a flat buffer (weights arena -> compute copy)
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    st_v2_u32(out + i * 4, pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// Small CUDA-core GEMM for the tiny FC-backward products (<= 0.1 GFLOP): C[M,N] (+)= op(A) op(B), fp32 out.
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
__global__ void __launch_bounds__(256)
small_gemm_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ Cf, bf16* __restrict__ Cb,
                  int M, int N, int K, long sam, long sak, long sbk, long sbn, long ldc) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int kk = e / 64, mm = e % 64;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __bfloat162float(A[m * sam + k * sak]) : 0.f;
      const int n = n0 + mm;
      Bs[kk][mm] = (n < N && k < K) ? __bfloat162float(B[k * sbk + n * sbn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        if (Cf) Cf[m * ldc + n] = acc[i][j];
        if (Cb) Cb[m * ldc + n] = __float2bfloat16_rn(acc[i][j]);
      }
    }
}

// bias add (+ column sums for the bias gradient)
__global__ void bias_add_kernel(bf16* __restrict__ y, const float* __restrict__ bias, int M, int N) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)M * N; i += (long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(__bfloat162float(y[i]) + bias[i % N]);
}
__global__ void colsum_kernel(const bf16* __restrict__ d, float* __restrict__ out, int M, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += __bfloat162float(d[(long)m * N + n]);
  out[n] = s;
}

inline int grid_for(long work, int block, int cap) {
  long g = (work + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int fd_bn_stats(const void* y, float* sum, float* sumsq, long M, int C, cudaStream_t s) {
  if (C % 8 || C > 2048) { fd_set_error("bn_stats: C must be a multiple of 8 and <= 2048", __FILE__, __LINE__); return -1; }
  const int C8 = C / 8;
  if (256 % C8 && C8 < 256) { /* RL = floor; leftover threads idle */ }
  const int RL = 256 / C8 > 0 ? 256 / C8 : 0;
  if (RL == 0) { fd_set_error("bn_stats: C too large for this tiling", __FILE__, __LINE__); return -1; }
  FD_CUDA_CHECK(cudaMemsetAsync(sum, 0, C * sizeof(float), s));
  FD_CUDA_CHECK(cudaMemsetAsync(sumsq, 0, C * sizeof(float), s));
  int grid = grid_for((M + 4L * RL - 1) / (4L * RL), 1, 148 * 6);
  size_t sm = (size_t)RL * C * 2 * sizeof(float);
  bn_stats_kernel<<<grid, 256, sm, s>>>(reinterpret_cast<const bf16*>(y), sum, sumsq, M, C);
  FD_LAUNCH_OK();
  return 0;
}

int fd_bn_apply(const void* y, const void* res, void* out, float* sum, float* sumsq, const float* gamma,
                const float* beta, float* run_mean, float* run_var, float* save_mean, float* save_rstd, long M, int C,
                float eps, float momentum, int training, int relu, unsigned* consumed, cudaStream_t s) {
  int grid = grid_for(M * (C / 8), 256 * 4, 148 * 8);
  FD_CUDA_CHECK(fd_launch_pdl(PDL_BN_FWD, bn_apply_kernel, dim3(grid), dim3(256), 2 * C * sizeof(float), s,
                              reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(res),
                              reinterpret_cast<bf16*>(out), sum, sumsq, gamma, beta, run_mean, run_var, save_mean, save_rstd,
                              M, C, eps, momentum, training, relu, consumed));
  FD_LAUNCH_OK();
  return 0;
}

int fd_bn_bwd(const void* dout, const void* y, const void* out, const float* mean, const float* rstd,
              const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* dres, long M, int C,
              int relu, cudaStream_t s) {
  const int C8 = C / 8;
  const int RL = 256 / C8;
  if (RL == 0) { fd_set_error("bn_bwd: C too large", __FILE__, __LINE__); return -1; }
  FD_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, C * sizeof(float), s));
  FD_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, C * sizeof(float), s));
  int grid = grid_for((M + 2L * RL - 1) / (2L * RL), 1, 148 * 6);
  bn_bwd_reduce_kernel<<<grid, 256, (size_t)RL * C * 2 * sizeof(float), s>>>(
      reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(out), mean,
      rstd, gamma, beta, dgamma, dbeta, M, C, relu);
  FD_LAUNCH_OK();
  int grid2 = grid_for(M * C8, 256 * 2, 148 * 8);
  bn_bwd_apply_kernel<<<grid2, 256, 5 * C * sizeof(float), s>>>(
      reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(out), mean,
      rstd, gamma, beta, dgamma, dbeta, reinterpret_cast<bf16*>(dy), reinterpret_cast<bf16*>(dres), M, C, relu);
  FD_LAUNCH_OK();
  return 0;
}

int fd_bn_bwd_apply(const void* dz, const void* y, const float* mean, const float* rstd, const float* gamma, float* s0,
                    float* s1, float* dgamma, float* dbeta, void* dy, long M, int C, unsigned* consumed, cudaStream_t s) {
  if (C % 8) { fd_set_error("bn_bwd_apply: C must be a multiple of 8", __FILE__, __LINE__); return -1; }
  int grid = grid_for(M * (C / 8), 256 * 4, 148 * 8);
  void (*kern)(const bf16*, const bf16*, const float*, const float*, const float*, float*, float*, float*, float*, bf16*,
               long, int, unsigned*) = bn_bwd_apply_kernel;  // the dz overload
  FD_CUDA_CHECK(fd_launch_pdl(PDL_BN_BWD_APPLY, kern, dim3(grid), dim3(256), 3 * C * sizeof(float), s, reinterpret_cast<const bf16*>(dz),
                              reinterpret_cast<const bf16*>(y), mean, rstd, gamma, s0, s1, dgamma, dbeta,
                              reinterpret_cast<bf16*>(dy), M, C, consumed));
  FD_LAUNCH_OK();
  return 0;
}

int fd_bn_relu_pool(const void* y, void* out, void* idx, float* sum, float* sumsq, const float* gamma, const float* beta,
                    float* run_mean, float* run_var, float* save_mean, float* save_rstd, int N, int H, int W, int C,
                    float eps, float momentum, int training, unsigned* consumed, cudaStream_t s) {
  if (C % 8) { fd_set_error("bn_relu_pool: C must be a multiple of 8", __FILE__, __LINE__); return -1; }
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  int grid = grid_for((long)N * P * Q * (C / 8), 256, 148 * 8);
  FD_CUDA_CHECK(fd_launch_pdl(PDL_BN_FWD, bn_relu_pool3s2_kernel, dim3(grid), dim3(256), 2 * C * sizeof(float), s,
                              reinterpret_cast<const bf16*>(y), reinterpret_cast<bf16*>(out),
                              reinterpret_cast<uint8_t*>(idx), sum, sumsq, gamma, beta, run_mean, run_var, save_mean,
                              save_rstd, N, H, W, C, P, Q, eps, momentum, training, consumed));
  FD_LAUNCH_OK();
  return 0;
}

int fd_maxpool_fwd(const void* x, void* out, void* idx, int N, int H, int W, int C, int k, int stride, int pad,
                   cudaStream_t s) {
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  int grid = grid_for((long)N * P * Q * (C / 8), 256, 148 * 16);
  maxpool_fwd_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out),
                                          reinterpret_cast<uint8_t*>(idx), N, H, W, C, P, Q, k, stride, pad);
  FD_LAUNCH_OK();
  return 0;
}
int fd_maxpool_bwd(const void* x, const void* out, const void* dout, void* dx, int N, int H, int W, int C, int k,
                   int stride, int pad, cudaStream_t s) {
  (void)out;
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  FD_CUDA_CHECK(cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * 2, s));
  int grid = grid_for((long)N * P * Q * (C / 8), 256, 148 * 16);
  maxpool_bwd_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dout),
                                          reinterpret_cast<bf16*>(dx), N, H, W, C, P, Q, k, stride, pad);
  FD_LAUNCH_OK();
  return 0;
}
// 3x3 / stride-2 / pad-1 max-pool backward, one thread per 2x2 block of input pixels (x 8 channels).  The four pixels
// (2p+a, 2q+b), a,b in {0,1}, are covered by the windows {p, p+1} x {q, q+1} only — pixel (2p, 2q) by window (p, q)
// alone, (2p+1, 2q) by (p, q) and (p+1, q), ... — so FOUR window loads (arg-max bytes + gradient) serve four output
// vectors where the per-pixel kernel issues 1+2+2+4 = 9, and the index arithmetic is paid once per block (the per-pixel
// kernel was issue-bound: ~390 instructions per output vector, profiles/r2/ncu_hot_kernels_summary.txt [15]).
// Window-relative position of pixel (2p+a, 2q+b) inside window (p+dp, q+dq): row a+1-2dp, col b+1-2dq.
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_block_kernel(const uint8_t* __restrict__ idx, const bf16* __restrict__ dout, bf16* __restrict__ dx, int N,
                            int H, int W, int C, int P, int Q) {
  pdl_wait();  // PDL: scheduled early, global memory only once the previous kernel of the stream has completed
  pdl_trigger();
  const int C8 = C / 8;
  const int HB = (H + 1) / 2, WB = (W + 1) / 2;
  const long total = (long)N * HB * WB * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cv, qb, pb, n;
    split4(i, C8, WB, HB, cv, qb, pb, n);
    uint2 pk[4];
    V8 dv[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // window (pb + dp, qb + dq), u = 2*dp + dq
      const int p = pb + (u >> 1), q = qb + (u & 1);
      ok[u] = (p < P) && (q < Q);
      if (ok[u]) {
        const long oi = (((long)n * P + p) * Q + q) * C8 + cv;
        pk[u] = *reinterpret_cast<const uint2*>(idx + oi * 8);
        dv[u] = ldg8(dout + oi * 8);
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int h = 2 * pb + a, w = 2 * qb + b;
        if (h >= H || w >= W) continue;
        __nv_bfloat162 acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int dp = u >> 1, dq = u & 1;
          if ((dp == 1 && a == 0) || (dq == 1 && b == 0)) continue;  // even rows/cols belong to one window only
          if (!ok[u]) continue;
          const uint32_t pos = (uint32_t)((a + 1 - 2 * dp) * 3 + (b + 1 - 2 * dq));
          const uint32_t rep = pos * 0x01010101u;
          const uint32_t m_lo = __vcmpeq4(pk[u].x, rep), m_hi = __vcmpeq4(pk[u].y, rep);
          const uint32_t w0 = dv[u].u[0] & __byte_perm(m_lo, 0, 0x1100), w1 = dv[u].u[1] & __byte_perm(m_lo, 0, 0x3322);
          const uint32_t w2 = dv[u].u[2] & __byte_perm(m_hi, 0, 0x1100), w3 = dv[u].u[3] & __byte_perm(m_hi, 0, 0x3322);
          acc[0] = __hadd2(acc[0], *reinterpret_cast<const __nv_bfloat162*>(&w0));
          acc[1] = __hadd2(acc[1], *reinterpret_cast<const __nv_bfloat162*>(&w1));
          acc[2] = __hadd2(acc[2], *reinterpret_cast<const __nv_bfloat162*>(&w2));
          acc[3] = __hadd2(acc[3], *reinterpret_cast<const __nv_bfloat162*>(&w3));
        }
        V8 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) o.u[e] = *reinterpret_cast<const uint32_t*>(&acc[e]);
        stg8(dx + ((((long)n * H + h) * W + w) * C8 + cv) * 8, o);
      }
    }
  }
}

int fd_maxpool_bwd_idx(const void* idx, const void* dout, void* dx, int N, int H, int W, int C, int k, int stride, int pad,
                       cudaStream_t s) {
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  int grid = grid_for((long)N * H * W * (C / 8), 256, 148 * 16);
  static int use_block = -1;
  if (use_block < 0) {
    const char* e = getenv("FLUXDIST_POOL_BWD_BLOCK");
    use_block = (e && e[0] == '0') ? 0 : 1;
  }
  if (k == 3 && stride == 2 && pad == 1 && use_block) {
    const int gridb = grid_for((long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8), 256, 148 * 16);
    FD_CUDA_CHECK(fd_launch_pdl(PDL_POOL, maxpool3s2_bwd_block_kernel, dim3(gridb), dim3(256), 0, s,
                                reinterpret_cast<const uint8_t*>(idx), reinterpret_cast<const bf16*>(dout),
                                reinterpret_cast<bf16*>(dx), N, H, W, C, P, Q));
  } else if (k == 3 && stride == 2 && pad == 1)
    maxpool_bwd_idx_kernel<3, 2, 1><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(idx),
                                                         reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dx), N,
                                                         H, W, C, P, Q, k, stride, pad);
  else
    maxpool_bwd_idx_kernel<0, 0, 0><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(idx),
                                                         reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dx), N,
                                                         H, W, C, P, Q, k, stride, pad);
  FD_LAUNCH_OK();
  return 0;
}
int fd_upsample_zero(const void* dy, void* up, int N, int P, int Q, int HU, int WU, int C, int stride, cudaStream_t s) {
  int grid = grid_for((long)N * HU * WU * (C / 8), 256, 148 * 16);
  upsample_zero_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(up), N, P, Q, HU,
                                            WU, C, stride);
  FD_LAUNCH_OK();
  return 0;
}
int fd_gap_fwd(const void* x, void* out, int N, int HW, int C, cudaStream_t s) {
  gap_fwd_kernel<<<grid_for((long)N * (C / 8), 256, 1024), 256, 0, s>>>(reinterpret_cast<const bf16*>(x),
                                                                       reinterpret_cast<bf16*>(out), N, HW, C);
  FD_LAUNCH_OK();
  return 0;
}
int fd_gap_bwd(const void* dout, void* dx, int N, int HW, int C, cudaStream_t s) {
  gap_bwd_kernel<<<grid_for((long)N * HW * (C / 8), 256, 148 * 8), 256, 0, s>>>(reinterpret_cast<const bf16*>(dout),
                                                                               reinterpret_cast<bf16*>(dx), N, HW, C);
  FD_LAUNCH_OK();
  return 0;
}

int fd_softmax_ce(const void* logits, const long long* labels, void* dlogits, float* row_loss, float* loss,
                  unsigned* counter, int B, int Cn, long ld, float scale, int is_bf16, cudaStream_t s) {
  const int warps = 8;
  int grid = (B + warps - 1) / warps;
  if (is_bf16)
    softmax_ce_kernel<bf16><<<grid, warps * 32, 0, s>>>(reinterpret_cast<const bf16*>(logits), labels,
                                                       reinterpret_cast<bf16*>(dlogits), row_loss, loss, counter, B,
                                                       Cn, ld, scale);
  else
    softmax_ce_kernel<float><<<grid, warps * 32, 0, s>>>(reinterpret_cast<const float*>(logits), labels,
                                                        reinterpret_cast<float*>(dlogits), row_loss, loss, counter, B,
                                                        Cn, ld, scale);
  FD_LAUNCH_OK();
  return 0;
}

int fd_normalize_u8(const void* in, void* out, long pixels, int Cpad, const float* mean, const float* stdv,
                    cudaStream_t s) {
  normalize_u8_kernel<<<grid_for(pixels, 256, 148 * 16), 256, 0, s>>>(
      reinterpret_cast<const uint8_t*>(in), reinterpret_cast<bf16*>(out), pixels, Cpad, mean[0], mean[1], mean[2],
      1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]);
  FD_LAUNCH_OK();
  return 0;
}

int fd_cast_f32_bf16(const float* in, void* out, long n, cudaStream_t s) {
  cast_f32_bf16_kernel<<<grid_for(n / 4, 256, 148 * 8), 256, 0, s>>>(in, reinterpret_cast<bf16*>(out), n / 4);
  FD_LAUNCH_OK();
  return 0;
}

int fd_small_gemm(const void* A, const void* B, float* Cf, void* Cb, int M, int N, int K, long sam, long sak, long sbk,
                  long sbn, long ldc, cudaStream_t s) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  small_gemm_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(B), Cf,
                                         reinterpret_cast<bf16*>(Cb), M, N, K, sam, sak, sbk, sbn, ldc);
  FD_LAUNCH_OK();
  return 0;
}
int fd_bias_add(void* y, const float* bias, int M, int N, cudaStream_t s) {
  bias_add_kernel<<<grid_for((long)M * N, 256, 1024), 256, 0, s>>>(reinterpret_cast<bf16*>(y), bias, M, N);
  FD_LAUNCH_OK();
  return 0;
}
int fd_colsum(const void* d, float* out, int M, int N, cudaStream_t s) {
  colsum_kernel<<<(N + 255) / 256, 256, 0, s>>>(reinterpret_cast<const bf16*>(d), out, M, N);
  FD_LAUNCH_OK();
  return 0;
}

int fd_s2d_pad(const void* x, void* x2, int N, int H, int W, int pad, cudaStream_t s) {
  const int HS = (H + 2 * pad) / 2, WS = (W + 2 * pad) / 2;
  s2d_pad_kernel<<<grid_for((long)N * HS * WS, 256, 148 * 16), 256, 0, s>>>(reinterpret_cast<const bf16*>(x),
                                                                          reinterpret_cast<bf16*>(x2), N, H, W, HS, WS, pad);
  FD_LAUNCH_OK();
  return 0;
}
int fd_stem_w_s2d(const void* w, void* w2, int K, cudaStream_t s) {
  stem_w_s2d_kernel<<<(K * 256 + 255) / 256, 256, 0, s>>>(reinterpret_cast<const bf16*>(w), reinterpret_cast<bf16*>(w2), K);
  FD_LAUNCH_OK();
  return 0;
}
int fd_stem_dw_from_s2d(const float* dw2, float* dw, int K, cudaStream_t s) {
  stem_dw_from_s2d_kernel<<<(K * 147 + 255) / 256, 256, 0, s>>>(dw2, dw, K);
  FD_LAUNCH_OK();
  return 0;
}

// ---- single-launch BN (grid-barrier kernels).  `ws` is a per-layer workspace: [counter, gen, pad, pad] (uint)
// followed by two [2][C] float accumulators (double-buffered by launch parity).
// CTA slots the grid-barrier kernels leave free: a bucket of the fused all-reduce+update kernel may be running on the
// communication stream at the same time (parallel/native_comm.py), and every CTA of a grid-barrier kernel must be
// resident for the barrier to complete.
static int g_bn_reserve = 0;
void fd_bn_set_reserve(int ctas) { g_bn_reserve = ctas < 0 ? 0 : ctas; }

static int bn_fused_grid(const void* fn, size_t smem, long M, int C) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem);
  if (per_sm < 1) return 0;
  if (per_sm > 4) per_sm = 4;  // more CTAs = more 16-byte loads in flight (HBM needs ~4.5 MB outstanding)
  // one CTA per ~64 KB of input: small layers must not pay for hundreds of same-address atomics and a wide
  // grid barrier (measured: 592 CTAs cost a fixed ~22 us per launch regardless of tensor size)
  static long chunk = 0;
  if (!chunk) {
    const char* e = getenv("FLUXDIST_BN_CHUNK_KB");  // tuning knob (scripts/bn_bench.py)
    chunk = (e && atoi(e) > 0) ? 1024L * atoi(e) : 0;
  }
  // measured (scripts/bn_bench.py, B200): 64 KB per CTA is best from 9.6 MB up; the 4.8 MB layers (7x7x512) gain 25%
  // from 32 KB per CTA (twice the CTAs -> half the dependent load rounds)
  const long bytes = M * C * 2;
  const long ck = chunk ? chunk : (bytes < (8L << 20) ? 32768 : 65536);
  long want = (bytes + ck - 1) / ck;
  long cap = (long)sms * per_sm - g_bn_reserve;
  if (cap < sms / 2) cap = sms / 2;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

int fd_bn_fwd_fused(const void* y, const void* res, void* out, const float* gamma, const float* beta, float* run_mean,
                    float* run_var, float* save_mean, float* save_rstd, void* ws, long M, int C, float eps,
                    float momentum, int relu, cudaStream_t s) {
  const int C8 = C / 8;
  const int RL = C8 ? 256 / C8 : 0;
  if (C % 8 || RL == 0) { fd_set_error("bn_fwd_fused: unsupported C", __FILE__, __LINE__); return -1; }
  size_t smem = (size_t)RL * C * 2 * sizeof(float);
  if (smem < (size_t)2 * C * sizeof(float)) smem = 2 * C * sizeof(float);
  const int grid = bn_fused_grid((const void*)bn_fwd_fused_kernel, smem, M, C);
  if (grid < 1) { fd_set_error("bn_fwd_fused: occupancy query failed", __FILE__, __LINE__); return -1; }
  GridSync gs{reinterpret_cast<unsigned*>(ws), reinterpret_cast<unsigned*>(ws) + 1};
  float* accbuf = reinterpret_cast<float*>(ws) + 4;
  bn_fwd_fused_kernel<<<grid, 256, smem, s>>>(reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(res),
                                              reinterpret_cast<bf16*>(out), gamma, beta, run_mean, run_var, save_mean,
                                              save_rstd, accbuf, gs, M, C, eps, momentum, relu);
  FD_LAUNCH_OK();
  return 0;
}

int fd_bn_bwd_fused(const void* dout, const void* y, const void* out, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* dres, void* ws,
                    long M, int C, int relu, cudaStream_t s) {
  const int C8 = C / 8;
  const int RL = C8 ? 256 / C8 : 0;
  if (C % 8 || RL == 0) { fd_set_error("bn_bwd_fused: unsupported C", __FILE__, __LINE__); return -1; }
  size_t smem = (size_t)RL * C * 2 * sizeof(float);
  if (smem < (size_t)5 * C * sizeof(float)) smem = 5 * C * sizeof(float);
  const int grid = bn_fused_grid((const void*)bn_bwd_fused_kernel, smem, M, C);
  if (grid < 1) { fd_set_error("bn_bwd_fused: occupancy query failed", __FILE__, __LINE__); return -1; }
  GridSync gs{reinterpret_cast<unsigned*>(ws), reinterpret_cast<unsigned*>(ws) + 1};
  float* accbuf = reinterpret_cast<float*>(ws) + 4;
  FD_CUDA_CHECK(fd_launch_pdl(PDL_BN_BWD_FUSED, bn_bwd_fused_kernel, dim3(grid), dim3(256), smem, s, reinterpret_cast<const bf16*>(dout),
                              reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(out), mean, rstd, gamma, beta,
                              dgamma, dbeta, reinterpret_cast<bf16*>(dy), reinterpret_cast<bf16*>(dres), accbuf, gs, M, C,
                              relu));
  FD_LAUNCH_OK();
  return 0;
}

// stem tail: pooled gradient g [N,P,Q,C] + arg-max bytes idx -> dy [N,H,W,C] of the stem conv output, dgamma, dbeta
int fd_bn_bwd_pool_fused(const void* idx, const void* g, const void* y, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, float* dgamma, float* dbeta, void* dy, void* ws, int N,
                         int H, int W, int C, cudaStream_t s) {
  const int C8 = C / 8;
  const int RL = C8 ? 256 / C8 : 0;
  if (C % 8 || RL == 0) { fd_set_error("bn_bwd_pool_fused: unsupported C", __FILE__, __LINE__); return -1; }
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const long M = (long)N * H * W;
  if (M * C8 >= 0x100000000LL) { fd_set_error("bn_bwd_pool_fused: tensor too large", __FILE__, __LINE__); return -1; }
  size_t smem = (size_t)RL * C * 2 * sizeof(float);
  if (smem < (size_t)5 * C * sizeof(float)) smem = 5 * C * sizeof(float);
  const int grid = bn_fused_grid((const void*)bn_bwd_pool_fused_kernel, smem, M, C);
  if (grid < 1) { fd_set_error("bn_bwd_pool_fused: occupancy query failed", __FILE__, __LINE__); return -1; }
  GridSync gs{reinterpret_cast<unsigned*>(ws), reinterpret_cast<unsigned*>(ws) + 1};
  float* accbuf = reinterpret_cast<float*>(ws) + 4;
  bn_bwd_pool_fused_kernel<<<grid, 256, smem, s>>>(
      reinterpret_cast<const uint8_t*>(idx), reinterpret_cast<const bf16*>(g), reinterpret_cast<const bf16*>(y), mean, rstd,
      gamma, beta, dgamma, dbeta, reinterpret_cast<bf16*>(dy), accbuf, gs, N, H, W, C, P, Q);
  FD_LAUNCH_OK();
  return 0;
}
