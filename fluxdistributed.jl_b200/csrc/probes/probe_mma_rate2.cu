// Hardware probe (scripts only): sustained tcgen05.mma rate, issued the way the product kernels issue it (convergent
// warp, predicated single-lane issue, descriptors = precomputed base + immediate).  M=128, K=16, N in {64,128,256};
// operands K-major or MN-major; all operands resident in shared memory (no TMA), so the number is the tensor pipe +
// shared-memory operand-read limit alone.  cta2 = 1: cta_group::2 pair (M=256 across two SMs, B split between them).
#include <cuda.h>

#include "../common.cuh"

namespace {

FD_DEVINL void umma_f16_ss_2cta_p(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                  uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}

template <int BN, bool AMN, bool BMN, bool CTA2>
__global__ void __launch_bounds__(128, 1) mma_rate2_kernel(long long* out, int reps, int a_stride16, int b_stride16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(slot, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync();
  tc_fence_after();
  const uint32_t tm = *slot;
  const int crank = CTA2 ? (int)cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && crank == 0) {
    const uint32_t leader = (lane == 0) ? 1u : 0u;
    constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, BN, AMN ? 1 : 0, BMN ? 1 : 0);
    const uint64_t dA = (AMN ? umma_smem_desc(0, 8192, 1024, UMMA_LAYOUT_SW128) : umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128)) +
                        (uint64_t)(smem_u32(smem) >> 4);
    const uint64_t dB = (BMN ? umma_smem_desc(0, 8192, 1024, UMMA_LAYOUT_SW128) : umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128)) +
                        (uint64_t)((smem_u32(smem) + 96 * 1024) >> 4);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // walk over a few tiles like a main loop does (stride in 16-byte units, wraps inside the 96 KB operand areas)
      const uint64_t da = dA + (uint64_t)((r & 3) * a_stride16);
      const uint64_t db = dB + (uint64_t)((r & 3) * b_stride16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (CTA2)
          umma_f16_ss_2cta_p(tm + (r & 1) * BN, da + (uint64_t)(k * (AMN ? 128 : 2)), db + (uint64_t)(k * (BMN ? 128 : 2)), idesc,
                             (r > 1 || k > 0) ? 1u : 0u, leader);
        else
          umma_f16_ss_p(tm + (r & 1) * BN, da + (uint64_t)(k * (AMN ? 128 : 2)), db + (uint64_t)(k * (BMN ? 128 : 2)), idesc,
                        (r > 1 || k > 0) ? 1u : 0u, leader);
      }
    }
    if (CTA2) {
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
          "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(done)),
          "r"(leader)
          : "memory");
    } else {
      umma_commit_p(done, leader);
    }
    mbar_wait(done, 0);
    t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync();
  if (warp == 0) {
    if (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    else
      tmem_dealloc(tm, 512);
  }
}

template <int BN, bool AMN, bool BMN, bool CTA2>
int launch_rate(long long* out, int reps, int grid, cudaStream_t s) {
  const int smem = 194 * 1024;
  auto fn = mma_rate2_kernel<BN, AMN, BMN, CTA2>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // tile strides: K-major tile of R rows = R*128 B; MN-major tile of 64 k x (R/64 blocks) = R/64 * 8 KB
  const int a16 = 16384 >> 4;
  const int b16 = (BN * 128) >> 4;
  if (CTA2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid & ~1);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // with cta_group::2 every CTA holds N/2 rows of B
    return cudaLaunchKernelEx(&cfg, fn, out, reps, a16, (int)(((BN / 2) * 128) >> 4)) == cudaSuccess ? 0 : -1;
  }
  fn<<<grid, 128, smem, s>>>(out, reps, a16, b16 > (16384 >> 4) ? (16384 >> 4) : b16);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// form: 0 = A,B K-major (fprop), 1 = B MN-major (dgrad), 2 = A,B MN-major (wgrad)
int fd_probe_mma_rate2(long long* out, int bn, int form, int cta2, int reps, int grid, cudaStream_t s) {
#define FD_RATE(BN_)                                                                               \
  if (bn == BN_) {                                                                                 \
    if (cta2) return form == 0 ? launch_rate<BN_, false, false, true>(out, reps, grid, s)          \
                   : form == 1 ? launch_rate<BN_, false, true, true>(out, reps, grid, s)           \
                               : launch_rate<BN_, true, true, true>(out, reps, grid, s);          \
    return form == 0 ? launch_rate<BN_, false, false, false>(out, reps, grid, s)                   \
         : form == 1 ? launch_rate<BN_, false, true, false>(out, reps, grid, s)                    \
                     : launch_rate<BN_, true, true, false>(out, reps, grid, s);                   \
  }
  FD_RATE(64)
  FD_RATE(128)
  FD_RATE(256)
#undef FD_RATE
  return -1;
}
