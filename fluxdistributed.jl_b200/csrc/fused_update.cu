// Fused gradient all-reduce + optimiser update over peer-mapped memory (NVLink 5 / NVSwitch).
//
// Replaces, in ONE launch per replica, the reference's whole exchange path
// (src/ddp_tasks.jl:59-109,163-172 — markbuffer! gather, (N-1) allocating add passes,
// a divide pass, getbuffer! broadcast, per-tensor Optimisers kernels):
//
//   barrier-in  -> sum the slice from every replica's gradient arena (P2P ld, or in-switch
//   multimem.ld_reduce) -> x 1/N -> Descent/Momentum/Nesterov/Adam/AdamW in registers on the
//   fp32 master + moments -> store the new weights (fp32 and/or bf16) into every replica
//   (P2P st, or multimem.st) -> barrier-out
//
// Variants (picked per bucket by the host, parallel/native_comm.py):
//   two-shot : rank r owns elements [r*n/N, (r+1)*n/N) — reads N-1 peers' slices, writes the
//              result to N-1 peers; 2 barriers; moments/master sharded by owner.
//   one-shot : every rank reduces the whole bucket itself in fixed rank order (bit-identical
//              results everywhere), no weight broadcast; for small, latency-bound buckets.
//   NVLS     : two-shot with multimem.ld_reduce / multimem.st on multicast addresses.
//
// Barriers are per-CTA epoch flags on signal pads (st.release.sys / ld.acquire.sys): CTA b of
// rank r raises flag[b][r] on every peer and waits for flag[b][*] on its own pad, so no grid
// sync and no host involvement; epochs are monotonically increasing so pads are never reset
// and the kernel can be replayed from a CUDA graph.  A bounded spin turns a dead peer into an
// error flag instead of a GPU hang.
#include "common.cuh"
#include "fused_update.h"

namespace {

constexpr int kThreads = 512;
// float4 per thread per iteration: 2 (round 1) or 4.  More 16-byte requests in flight per thread cover more of the
// NVLink / NVSwitch round trip (the in-switch reduction of multimem.ld_reduce has a multi-microsecond latency); picked per
// plan (FusedArgs::unroll), measured by bench_allreduce.py.

struct Hyper {
  float eta, a, b, eps, decay;
};

template <int RULE>
FD_DEVINL void apply_rule(float g, float& w, float& s0, float& s1, const Hyper& h, float c1inv, float c2inv) {
  if constexpr (RULE == FD_RULE_DESCENT) {
    w -= h.eta * g;
  } else if constexpr (RULE == FD_RULE_MOMENTUM) {
    s0 = h.a * s0 - h.eta * g;  // v <- rho v - eta g ; x <- x + v   (Optimisers.jl 0.1)
    w += s0;
  } else if constexpr (RULE == FD_RULE_NESTEROV) {
    float d = -(h.a * h.a) * s0 + (1.f + h.a) * h.eta * g;
    s0 = h.a * s0 - h.eta * g;
    w -= d;
  } else {  // ADAM / ADAMW
    s0 = h.a * s0 + (1.f - h.a) * g;
    s1 = h.b * s1 + (1.f - h.b) * g * g;
    float d = (s0 * c1inv) / (sqrtf(s1 * c2inv) + h.eps) * h.eta;
    if constexpr (RULE == FD_RULE_ADAMW) d += h.eta * h.decay * w;
    w -= d;
  }
}

// per-CTA cross-GPU barrier
FD_DEVINL bool cta_barrier(const FusedArgs& A, uint32_t epoch) {
  __syncthreads();
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  if (threadIdx.x < (unsigned)A.world) {
    const int peer = threadIdx.x;
    uint32_t* remote = A.flags[peer] + (size_t)blockIdx.x * FD_MAX_WORLD + A.rank;
    st_release_sys(remote, epoch);  // cumulative with the bar.sync above: publishes this CTA's prior writes
    const uint32_t* mine = A.flags[A.rank] + (size_t)blockIdx.x * FD_MAX_WORLD + peer;
    long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > A.max_spins) {
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_fail) {
    if (threadIdx.x == 0) atomicExch(A.status + 1, 1);  // status[1] = barrier timeout
    return false;
  }
  return true;
}

template <int RULE, bool NVLS, int kUnroll = 2>
__global__ void __launch_bounds__(kThreads) fused_allreduce_update_kernel(const FusedArgs A) {
  const int W = A.world;
  const float inv_w = 1.f / (float)W;
  // step counter (Adam bias correction) lives on the device so a captured graph stays valid
  const int t = *A.step_dev + 1;
  Hyper h{A.h[0], A.h[1], A.h[2], A.h[3], A.h[4]};
  float grad_scale = A.grad_scale;
  if (A.hyper_dev != nullptr) {  // hyper-parameters live in device memory: graph replays see set_opt()
    h.eta = A.hyper_dev[0]; h.a = A.hyper_dev[1]; h.b = A.hyper_dev[2]; h.eps = A.hyper_dev[3]; h.decay = A.hyper_dev[4];
    grad_scale = A.hyper_dev[7];
  }
  float c1inv = 1.f, c2inv = 1.f;
  if constexpr (RULE == FD_RULE_ADAM || RULE == FD_RULE_ADAMW) {
    c1inv = 1.f / (1.f - powf(h.a, (float)t));
    c2inv = 1.f / (1.f - powf(h.b, (float)t));
  }
  uint32_t epoch = 0;
  if (W > 1) {
    epoch = A.epoch[blockIdx.x] + 1;  // read by every thread before thread 0 bumps it at the end
    if (!cta_barrier(A, epoch)) return;
  }

  const bool two_shot = (A.mode == FD_MODE_TWO_SHOT) && W > 1;
  const long n4 = A.n / 4;
  const long slice4 = two_shot ? n4 / W : n4;
  const long base4 = two_shot ? slice4 * A.rank : 0;
  bool bad = false;

  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < slice4; i += (long)gridDim.x * kThreads * kUnroll) {
    float4 g[kUnroll];
    long idx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      idx[u] = i + (long)u * gridDim.x * kThreads;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if constexpr (NVLS) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (idx[u] < slice4) g[u] = multimem_ld_reduce_add_v4(A.mc_g + (base4 + idx[u]) * 4);
    } else {
      // fixed rank order => identical rounding on every replica (one-shot stays bit-identical)
      for (int p = 0; p < W; ++p) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (idx[u] < slice4) {
            const float* src = A.g[p] + (base4 + idx[u]) * 4;
            float4 v = (W > 1) ? ld_relaxed_sys_v4(src) : ld_nc_v4(src);
            g[u].x += v.x; g[u].y += v.y; g[u].z += v.z; g[u].w += v.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (idx[u] >= slice4) continue;
      const long e = (base4 + idx[u]) * 4;
      const float gs = inv_w * grad_scale;
      float gg[4] = {g[u].x * gs, g[u].y * gs, g[u].z * gs, g[u].w * gs};
      float4 w4 = *reinterpret_cast<const float4*>(A.w32 + e);
      float ww[4] = {w4.x, w4.y, w4.z, w4.w};
      float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
      if constexpr (RULE != FD_RULE_DESCENT) {
        float4 t0 = *reinterpret_cast<const float4*>(A.s0 + e);
        s0[0] = t0.x; s0[1] = t0.y; s0[2] = t0.z; s0[3] = t0.w;
      }
      if constexpr (RULE == FD_RULE_ADAM || RULE == FD_RULE_ADAMW) {
        float4 t1 = *reinterpret_cast<const float4*>(A.s1 + e);
        s1[0] = t1.x; s1[1] = t1.y; s1[2] = t1.z; s1[3] = t1.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bad |= !isfinite(gg[k]);
        apply_rule<RULE>(gg[k], ww[k], s0[k], s1[k], h, c1inv, c2inv);
      }
      float4 wn = make_float4(ww[0], ww[1], ww[2], ww[3]);
      *reinterpret_cast<float4*>(A.w32 + e) = wn;
      if constexpr (RULE != FD_RULE_DESCENT) *reinterpret_cast<float4*>(A.s0 + e) = make_float4(s0[0], s0[1], s0[2], s0[3]);
      if constexpr (RULE == FD_RULE_ADAM || RULE == FD_RULE_ADAMW)
        *reinterpret_cast<float4*>(A.s1 + e) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      const uint32_t lo = pack_bf16x2(ww[0], ww[1]), hi = pack_bf16x2(ww[2], ww[3]);
      if (two_shot) {
        if constexpr (NVLS) {
          if (A.mc_w32 && e >= A.fp32_bcast_from) multimem_st_v4(A.mc_w32 + e, wn);
          if (A.mc_w16) multimem_st_bf16x4(A.mc_w16 + e, lo, hi);
        } else {
          const bool b32 = e >= A.fp32_bcast_from;
          for (int p = 0; p < W; ++p) {
            if (b32 && A.wout32[p] && A.wout32[p] != A.w32) st_v4(A.wout32[p] + e, wn);
            if (A.wout16[p]) st_v2_u32(A.wout16[p] + e, lo, hi);
          }
        }
      } else {
        float* o32 = A.wout32[A.rank];
        if (o32 && o32 != A.w32) st_v4(o32 + e, wn);
        if (A.wout16[A.rank]) st_v2_u32(A.wout16[A.rank] + e, lo, hi);
      }
    }
  }
  if (bad) atomicExch(A.status, 1);  // status[0] = non-finite reduced gradient (check_nans, K18)

  if (W > 1) {
    // barrier-out: (a) every peer's stores into my weights are visible to my next kernel,
    // (b) every peer is done reading my gradients before backward overwrites them.
    if (!cta_barrier(A, epoch + 1)) return;
    if (threadIdx.x == 0) A.epoch[blockIdx.x] = epoch + 1;
  }
  // last CTA out bumps the step counter (all CTAs have read it by now)
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned done = atomicAdd(A.done_ctr, 1u);
    if (done == gridDim.x - 1) {
      *A.done_ctr = 0;
      *A.step_dev = t;
    }
  }
}

template <int RULE>
int launch_rule(const FusedArgs& a, int grid, cudaStream_t s) {
  if (a.unroll >= 4) {
    if (a.use_nvls)
      fused_allreduce_update_kernel<RULE, true, 4><<<grid, kThreads, 0, s>>>(a);
    else
      fused_allreduce_update_kernel<RULE, false, 4><<<grid, kThreads, 0, s>>>(a);
  } else {
    if (a.use_nvls)
      fused_allreduce_update_kernel<RULE, true, 2><<<grid, kThreads, 0, s>>>(a);
    else
      fused_allreduce_update_kernel<RULE, false, 2><<<grid, kThreads, 0, s>>>(a);
  }
  FD_LAUNCH_OK();
  return 0;
}

}  // namespace

int fd_fused_update_grid(long n, int world, int mode, int max_grid) {
  constexpr int kUnroll = 2;
  long n4 = n / 4;
  long slice4 = (mode == FD_MODE_TWO_SHOT && world > 1) ? n4 / world : n4;
  long want = (slice4 + (long)kThreads * kUnroll - 1) / ((long)kThreads * kUnroll);
  if (want < 1) want = 1;
  if (want > max_grid) want = max_grid;
  return (int)want;
}

int fd_fused_update_launch(const FusedArgs& a, int grid, cudaStream_t stream) {
  switch (a.rule) {
    case FD_RULE_DESCENT: return launch_rule<FD_RULE_DESCENT>(a, grid, stream);
    case FD_RULE_MOMENTUM: return launch_rule<FD_RULE_MOMENTUM>(a, grid, stream);
    case FD_RULE_NESTEROV: return launch_rule<FD_RULE_NESTEROV>(a, grid, stream);
    case FD_RULE_ADAM: return launch_rule<FD_RULE_ADAM>(a, grid, stream);
    case FD_RULE_ADAMW: return launch_rule<FD_RULE_ADAMW>(a, grid, stream);
  }
  fd_set_error("unknown optimiser rule", __FILE__, __LINE__);
  return -1;
}

// ------------------------------------------------------------------ all-gather of owner shards
// Two-shot keeps master weights / moments authoritative only on the owner slice; this kernel
// re-assembles a full local copy by reading each owner's slice from its arena (state(), checkpoints).
__global__ void gather_shards_kernel(float* dst, const float* const* src, int world, int rank, long n) {
  const long n4 = n / 4, slice4 = n4 / world;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    int owner = (int)(i / slice4);
    if (owner >= world) owner = world - 1;
    if (owner == rank) continue;
    reinterpret_cast<float4*>(dst)[i] = ld_relaxed_sys_v4(src[owner] + i * 4);
  }
}

int fd_gather_shards(float* dst, const float* const* src_dev_array, int world, int rank, long n, cudaStream_t s) {
  gather_shards_kernel<<<296, 512, 0, s>>>(dst, src_dev_array, world, rank, n);
  FD_LAUNCH_OK();
  return 0;
}
