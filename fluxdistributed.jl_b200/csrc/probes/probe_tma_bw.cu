// Hardware probe (scripts only, built into _Cprobe, never into the product extension): shared-memory fill bandwidth
// of tiled TMA loads out of L2.  Every CTA streams 16 KB boxes (64 bf16 x 128 rows, 128B swizzle) through an 8-deep
// mbarrier ring and does nothing with them.  `share` CTAs read the SAME region (1 = every CTA its own, 148 = all the
// same): separates the L2 -> SM fill limit from what request merging / multicast recovers.  mc = 1: clusters of 2,
// each CTA loads half of every box and multicasts it to its peer (what the 2-CTA conv variants do for shared operands).
#include <cuda.h>

#include "../common.cuh"

namespace {

typedef CUresult (*PFN_encTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kStages = 8;
constexpr int kBox = 16384;

template <bool MC>
__global__ void __launch_bounds__(64, 1)
tma_bw_kernel(const __grid_constant__ CUtensorMap map, long long* cycles, int boxes_per_region, int share, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kBox);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MC ? 2 : 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (MC) cluster_sync();
  const int crank = MC ? (int)cluster_ctarank() : 0;
  const int unit = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int region = unit / share;
  const int total = boxes_per_region * reps;
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < total; ++i) {
        const int box = i % boxes_per_region;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], kBox);
        const int row = (region * boxes_per_region + box) * 128;
        if (MC)
          tma_load_2d_mc_p(smem + st * kBox + crank * (kBox / 2), &map, &full[st], 0, row + crank * 64, 3, 1u);
        else
          tma_load_2d(smem + st * kBox, &map, &full[st], 0, row);
        if (++st == kStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (lane == 0) {
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[st], ph);
      if (MC) {
        // release the slot in both CTAs (the peer multicasts into it)
        asm volatile(
            "{\n\t.reg .b32 ra;\n\t"
            "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
            "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(&empty[st])),
            "r"(crank ^ 1)
            : "memory");
      }
      mbar_arrive(&empty[st]);
      if (++st == kStages) {
        st = 0;
        ph ^= 1;
      }
    }
  }
  __syncthreads();
  if (MC) cluster_sync();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

}  // namespace

void fd_set_error(const char* msg, const char* file, int line);

// buf: bf16 [rows][64]; every region is boxes_per_region*128 rows
int fd_probe_tma_bw(const void* buf, long rows, long long* cycles, int grid, int boxes_per_region, int share, int reps,
                    int mc, cudaStream_t s) {
  static PFN_encTiled enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return -1;
    enc = (PFN_encTiled)f;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t st[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)(mc ? 64 : 128)};
  cuuint32_t es[2] = {1, 1};
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(buf), dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    fd_set_error("probe_tma_bw: encode", __FILE__, __LINE__);
    return -1;
  }
  const int smem = kStages * kBox + 1024 + 256;
  if (mc) {
    cudaFuncSetAttribute(tma_bw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid & ~1);
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, tma_bw_kernel<true>, m, cycles, boxes_per_region, share, reps) != cudaSuccess) {
      fd_set_error(cudaGetErrorString(cudaGetLastError()), __FILE__, __LINE__);
      return -1;
    }
  } else {
    cudaFuncSetAttribute(tma_bw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tma_bw_kernel<false><<<grid, 64, smem, s>>>(m, cycles, boxes_per_region, share, reps);
    if (cudaGetLastError() != cudaSuccess) {
      fd_set_error("probe_tma_bw launch", __FILE__, __LINE__);
      return -1;
    }
  }
  return 0;
}
