// Shared device helpers for the sm_100a kernels: vector/bf16 utilities and thin
// inline-PTX wrappers (system-scope flags, mbarrier, TMA, tcgen05/TMEM).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FD_DEVINL __device__ __forceinline__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

#define FD_CUDA_CHECK(x)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fd_set_error(cudaGetErrorString(e_), __FILE__, __LINE__);                                 \
      return (int)e_;                                                                           \
    }                                                                                           \
  } while (0)

void fd_set_error(const char* msg, const char* file, int line);
void fd_count_launch();
long fd_launch_count();
#define FD_LAUNCH_OK()     \
  do {                   \
    fd_count_launch();   \
    FD_CUDA_CHECK(cudaGetLastError()); \
  } while (0)
const char* fd_last_error();

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A kernel launched through fd_launch_pdl() may be scheduled while the previous kernel of the stream is still
// draining: its CTAs run their on-chip prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in
// pdl_wait() until the previous grid has completed and its writes are visible.  Every kernel launched this way MUST
// execute pdl_wait() before its first global-memory access in every CTA that touches memory (that also keeps the
// chain transitive: "my predecessor is complete" implies "its predecessor is complete").  pdl_trigger() lets the NEXT
// kernel of the stream start being scheduled; both are no-ops for a kernel launched without the attribute.
// The attribute is applied per kernel CLASS (PDL_* below): FLUXDIST_PDL is a bit mask over the classes (0 = plain
// stream-ordered launches everywhere); fd_set_pdl() overrides it at run time.
FD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
FD_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
enum { PDL_IGEMM = 0, PDL_IGEMM_WGRAD = 1, PDL_HALO = 2, PDL_HALO_WGRAD = 3, PDL_BN_FWD = 4, PDL_BN_BWD_FUSED = 5,
       PDL_BN_BWD_APPLY = 6, PDL_POOL = 7 };
int fd_pdl_mask();
void fd_set_pdl(int mask);

template <typename... KArgs, typename... Args>
inline cudaError_t fd_launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ((fd_pdl_mask() >> cls) & 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ packing
FD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
FD_DEVINL float2 unpack_bf16x2(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}
FD_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------ warp / block reductions
FD_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
FD_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Column sums of a 32x32 fp32 block held one row per lane (v[c] = this lane's value of column c): a transpose-reduce
// butterfly on the shuffle network — after it, lane L holds the sum over all 32 rows of column L.  ~125 instructions
// and NO shared-memory traffic: the tcgen05 main loops are bound by shared-memory operand bandwidth (an M=128,K=16 MMA
// costs max(80, N/2) clk, scripts/probe_mma_rate.py), so every byte the epilogue warps move through shared memory
// slows the MMAs of the next tile — the round-2 measurement behind switching the statistics / dz epilogues from the
// shared-memory tile reduction (fewer instructions) back to shuffles.  Destroys v.
FD_DEVINL float warp_colsum_shfl(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float keep = up ? v[j + off] : v[j];
      const float send = up ? v[j] : v[j + off];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------ system-scope flags (cross-GPU barriers)
FD_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
FD_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
FD_DEVINL void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

FD_DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// streaming 128-bit accesses (peer or local)
FD_DEVINL float4 ld_nc_v4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
FD_DEVINL float4 ld_relaxed_sys_v4(const float* p) {
  // peer gradient reads: must not be served from a stale non-coherent path
  float4 r;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
FD_DEVINL void st_v4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
FD_DEVINL void st_v2_u32(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// NVSwitch multicast (NVLS): in-switch reduction load and broadcast store
FD_DEVINL float4 multimem_ld_reduce_add_v4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
FD_DEVINL void multimem_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
FD_DEVINL void multimem_st_bf16x4(void* mc, uint32_t a, uint32_t b) {
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1,%2};" ::"l"(mc), "r"(a), "r"(b) : "memory");
}

// ------------------------------------------------------------------ mbarrier
FD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

FD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FD_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FD_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
FD_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FD_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
FD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------ TMA
FD_DEVINL void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
FD_DEVINL void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
FD_DEVINL void tma_load_4d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// im2col-mode load of an NHWC tensor: coordinates {c, w, h, n} select the base pixel, {off_w, off_h} the filter tap
FD_DEVINL void tma_load_im2col_4d(void* smem_dst, const void* desc, uint64_t* bar, int c, int w, int h, int n,
                                  uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2], {%7, %8};" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
FD_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
FD_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
FD_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
FD_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FD_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate
FD_DEVINL void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// make an mbarrier track completion of all prior tcgen05.mma of this thread
FD_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns per warp (thread i owns lane i)
FD_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
FD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version=1 for sm_100)
FD_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
enum { UMMA_LAYOUT_NONE = 0, UMMA_LAYOUT_SW128 = 2, UMMA_LAYOUT_SW64 = 4, UMMA_LAYOUT_SW32 = 6 };

// instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor bit layout)
FD_DEVINL constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | ((uint32_t)a_mn_major << 15)  // a_major
         | ((uint32_t)b_mn_major << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}

// ---- predicated single-lane issue from CONVERGENT code -------------------------------------------------------
// tcgen05.mma / commit / TMA take their operands from uniform registers.  Issuing them under `if (lane == 0)` puts
// them in a divergent region where the compiler must assume per-thread values and wraps each one in an
// ELECT + R2UR.BROADCAST loop (~19 SASS instructions per MMA; see profiles/).  Keeping the warp convergent and
// predicating the instruction itself (`@q`) lets descriptors live in uniform registers: ~3 instructions per MMA.
FD_DEVINL void umma_f16_ss_p(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                             uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
FD_DEVINL void umma_commit_p(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(issue)
      : "memory");
}
FD_DEVINL void mbar_arrive_expect_tx_p(uint64_t* bar, uint32_t bytes, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(issue)
      : "memory");
}
FD_DEVINL void mbar_arrive_p(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(issue)
      : "memory");
}
FD_DEVINL void tma_load_2d_p(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      "\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(issue)
      : "memory");
}
FD_DEVINL void tma_load_4d_p(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3,
                             uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(issue)
      : "memory");
}
FD_DEVINL void tma_load_im2col_4d_p(void* smem_dst, const void* desc, uint64_t* bar, int c, int w, int h, int n,
                                    uint16_t off_w, uint16_t off_h, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %9, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2], {%7, %8};\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h), "r"(issue)
      : "memory");
}
// ---- thread-block clusters: rank, barrier, multicast TMA, multicast commit
FD_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
FD_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one tile lands at the same CTA-relative smem offset of every CTA in `mask`, and signals the mbarrier at the same
// CTA-relative offset in each of them
FD_DEVINL void tma_load_2d_mc_p(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, uint16_t mask,
                                uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask), "r"(issue)
      : "memory");
}
FD_DEVINL void umma_commit_mc_p(uint64_t* bar, uint16_t mask, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::
          "r"(smem_u32(bar)),
      "h"(mask), "r"(issue)
      : "memory");
}

// warp-uniform broadcast of a value every lane already holds (tells the compiler it is uniform)
FD_DEVINL uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

FD_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .b32 r;\n\t"
      "elect.sync r|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
