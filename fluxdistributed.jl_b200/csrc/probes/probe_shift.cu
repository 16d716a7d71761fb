// Hardware probe (test-only): do UMMA SW128 shared-memory descriptors accept a start address shifted by whole
// 128-byte rows and an 8-row-group stride (SBO) that is not a multiple of 1024 B?  If the swizzle XOR is a
// function of the absolute shared-memory address (as TMA's is), one halo tile in smem can serve every filter tap
// of a 3x3 convolution through descriptor arithmetic alone.  Used by tests/test_probe_gpu.py.
#include <cuda.h>

#include "../common.cuh"
#include "../igemm.h"

namespace {

typedef CUresult (*PFN_encodeTiledP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


// A: [rows x 64] bf16 row-major global, loaded whole into smem (SW128) by TMA in boxes of 64 rows.
// K-major probe : D[m][n] = sum_k A[row(m)][k] * B[n][k],  row(m) = shift + (m/8)*(sbo/128) + m%8
// MN-major probe: D[m][n] = sum_k At[krow(k)][m] * B[n][k] with At tile [k rows][64 m], krow(k) = shift + (k/8)*(sbo/128)+k%8
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* D, int rows,
             int shift, int sbo, int base_off, int mn_major, int nk16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // rows*128 B
  uint8_t* sB = smem + 48 * 1024;           // 64 x 64 bf16 = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 60 * 1024);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rows * 128 + 64 * 128);
    for (int r = 0; r < rows; r += 64) tma_load_2d(sA + r * 128, &mapA, bar, 0, r);
    tma_load_2d(sB, &mapB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, mn_major, 0);
    for (int k = 0; k < nk16; ++k) {
      uint64_t da, db;
      if (!mn_major) {
        da = umma_smem_desc(smem_u32(sA) + shift * 128 + k * 32, 16, sbo, UMMA_LAYOUT_SW128);
      } else {  // A tile is [k rows][64 m]; each UMMA_K=16 consumes two 8-row groups
        da = umma_smem_desc(smem_u32(sA) + shift * 128 + k * 2 * sbo, 0, sbo, UMMA_LAYOUT_SW128);  // LBO=0: rows 64..127 alias 0..63
      }
      da |= (uint64_t)(base_off & 7) << 49;
      db = umma_smem_desc(smem_u32(sB) + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
      umma_f16_ss(tm, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}


// 32-byte-row variant of the probe (SWIZZLE_32B): A is [rows x 16] bf16 (one 16-channel pixel per row).
__global__ void __launch_bounds__(128, 1)
probe32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* D, int rows,
               int shift, int sbo, int mn_major) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // rows*32 B
  uint8_t* sB = smem + 32 * 1024;  // 64 x 16 bf16 = 2 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 40 * 1024);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rows * 32 + 64 * 32);
    for (int r = 0; r < rows; r += 64) tma_load_2d(sA + r * 32, &mapA, bar, 0, r);
    tma_load_2d(sB, &mapB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, mn_major, 0);
    uint64_t da = mn_major ? umma_smem_desc(smem_u32(sA) + shift * 32, 0, sbo, UMMA_LAYOUT_SW32)
                           : umma_smem_desc(smem_u32(sA) + shift * 32, 16, sbo, UMMA_LAYOUT_SW32);
    uint64_t db = umma_smem_desc(smem_u32(sB), 16, 256, UMMA_LAYOUT_SW32);
    umma_f16_ss(tm, da, db, idesc, 0u);
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

// Timing probe: cycles for `reps` back-to-back tcgen05.mma (M=128, N=BN, K=16) with a given A start shift / SBO.
template <int BN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int reps, int shift, int sbo, int flavor) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 196 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    // flavor bit4: A MN-major, bit5: B MN-major (descriptor forms of the wgrad / dgrad main loops)
    const int amn = (flavor >> 4) & 1, bmn = (flavor >> 5) & 1;
    const uint32_t idesc = umma_idesc_bf16(128, BN, amn, bmn);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      // flavor bit0: alternate accumulator buffers every 9 groups; bit1: commit + wait every 9 groups (tile boundary)
      const uint32_t d = tm + (((flavor & 1) && ((r / 9) & 1)) ? BN : 0);
      const int tapshift = (flavor & 4) ? ((r % 9) / 3 * 10 + (r % 3)) : 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da = amn ? umma_smem_desc(a + k * 2048, 8192, 1024, UMMA_LAYOUT_SW128)
                          : umma_smem_desc(a + (shift + tapshift) * 128 + k * 32, 16, sbo, UMMA_LAYOUT_SW128);
        uint64_t db = bmn ? umma_smem_desc(b + k * 2048, 8192, 1024, UMMA_LAYOUT_SW128)
                          : umma_smem_desc(b + ((flavor & 4) ? (r % 9) * BN * 128 : 0) + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
        umma_f16_ss(d, da, db, idesc, (r % 9 == 0 && k == 0) ? 0u : 1u);
      }
      if ((flavor & 2) && (r % 9 == 8)) {
        umma_commit(done);
        if (flavor & 8) {
          mbar_wait(done, ph);
          ph ^= 1;
        }
      }
    }
    if (!(flavor & 8)) {
      umma_commit(done);
      // only the final phase is awaited when commits were issued without waiting: count completions
    }
    if (!(flavor & 2) || (flavor & 8)) {
      if (!(flavor & 8)) mbar_wait(done, 0);
    } else {
      // (reps/9 + 1) arrivals on a count-1 barrier: wait for each phase in turn
      const int n = reps / 9 + 1;
      for (int i = 0; i < n; ++i) mbar_wait(done, i & 1);
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace

int fd_probe_mma_rate(long long* out, int bn, int reps, int shift, int sbo, int grid, int flavor, cudaStream_t s) {
  const int smem = 200 * 1024;
  if (bn == 64) {
    FD_CUDA_CHECK(cudaFuncSetAttribute(mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<64><<<grid, 128, smem, s>>>(out, reps, shift, sbo, flavor);
  } else if (bn == 128) {
    FD_CUDA_CHECK(cudaFuncSetAttribute(mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<128><<<grid, 128, smem, s>>>(out, reps, shift, sbo, flavor);
  } else {
    FD_CUDA_CHECK(cudaFuncSetAttribute(mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<256><<<grid, 128, smem, s>>>(out, reps, shift, sbo, flavor);
  }
  FD_LAUNCH_OK();
  return 0;
}

int fd_probe_shift(const void* A, const void* B, float* D, int rows, int shift, int sbo, int base_off, int mn_major,
                   int nk16, cudaStream_t s);

int fd_probe_shift32(const void* A, const void* B, float* D, int rows, int shift, int sbo, int mn_major, cudaStream_t s) {
  static PFN_encodeTiledP enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    FD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    enc = (PFN_encodeTiledP)f;
  }
  CUtensorMap ma, mb;
  cuuint32_t es[2] = {1, 1};
  cuuint64_t dims[2] = {16, (cuuint64_t)rows};
  cuuint64_t st[1] = {32};
  cuuint32_t box[2] = {16, 64};
  cuuint64_t dimsb[2] = {16, 64};
  if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), dimsb, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    fd_set_error("probe32: encode", __FILE__, __LINE__);
    return -1;
  }
  FD_CUDA_CHECK(cudaFuncSetAttribute(probe32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  probe32_kernel<<<1, 128, 48 * 1024, s>>>(ma, mb, D, rows, shift, sbo, mn_major);
  FD_LAUNCH_OK();
  return 0;
}

int fd_probe_shift(const void* A, const void* B, float* D, int rows, int shift, int sbo, int base_off, int mn_major,
                   int nk16, cudaStream_t s) {
  static PFN_encodeTiledP enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    FD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    enc = (PFN_encodeTiledP)f;
  }
  CUtensorMap ma, mb;
  cuuint32_t es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t st[1] = {128};
    cuuint32_t box[2] = {64, 64};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dims, st, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      fd_set_error("probe: encode A", __FILE__, __LINE__);
      return -1;
    }
    cuuint64_t dimsb[2] = {64, 64};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), dimsb, st, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      fd_set_error("probe: encode B", __FILE__, __LINE__);
      return -1;
    }
  }
  FD_CUDA_CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  probe_kernel<<<1, 128, 64 * 1024, s>>>(ma, mb, D, rows, shift, sbo, base_off, mn_major, nk16);
  FD_LAUNCH_OK();
  return 0;
}
