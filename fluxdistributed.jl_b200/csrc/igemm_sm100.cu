// Implicit-GEMM convolution and GEMM on 5th-gen tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel template serves
//   GEMM  : D[M,N] = A[M,K] * B[N,K]^T                       (FC head, 1x1 convs)
//   FPROP : Y[n,p,q,k]  = sum_{r,s,c} X[n,p*st+r-pad,q*st+s-pad,c] W[k,r,s,c]        
//   DGRAD : dX[n,h,w,c] = sum_{r,s,k} dY[n,h+pad-r,w+pad-s,k] W[k,r,s,c]   (stride 1)  
//   WGRAD : dW[k,r,s,c] = sum_{n,p,q} dY[n,p,q,k] X[n,p*st+r-pad,q*st+s-pad,c]        
// with bf16 operands and fp32 accumulation:
//   * operands are staged in shared memory by TMA — activations through *im2col-mode* tensor
//     maps (cp.async.bulk.tensor.4d...im2col) so padding/stride/halo handling is done by the
//     copy engine, weights/plain matrices through tiled maps — into 128B-swizzled tiles;
//   * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=64|128, K=16) reading the
//     tiles through UMMA shared-memory descriptors; accumulators live in TMEM, double-buffered
//     so the epilogue of tile i overlaps the main loop of tile i+1;
//   * four epilogue warps pull the accumulator out with tcgen05.ld (32 lanes x 32 columns),
//     convert and store (bf16 NHWC rows, or fp32 red.add for split-K wgrad).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <string>
#include <cstring>

#include "common.cuh"
#include "igemm.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block = one 128B swizzle row
constexpr int kThreads = 192;

enum { MODE_GEMM = 0, MODE_FPROP = 1, MODE_DGRAD = 2, MODE_WGRAD = 3, MODE_GEMM_NN = 4, MODE_GEMM_TN = 5, MODE_DGRAD_S2 = 6 };

struct KParams {
  int mode;
  int M, N;           // GEMM extents seen by the epilogue (rows, cols)
  int num_kb;         // k-blocks per tile (per split for wgrad)
  int m_tiles, n_tiles, splits;
  // conv geometry
  int PQ, Q;          // output (M-space) image extent for m -> (n,p,q)
  int stride;         // traversal stride of the im2col walk
  int lower_w, lower_h;
  int S, C_chunks;    // taps per filter row; 64-channel chunks per tap (of the reduction operand)
  int Ctot;           // channels per tap in the weight matrix row (R*S*C row => column = tap*Ctot + c)
  int flip, R;        // dgrad: use tap (R-1-r, S-1-s) offsets
  // wgrad
  int taps;           // R*S
  int kb_total;       // total pixel k-blocks
  // output
  void* out;
  long ldo;           // row stride in elements
  int out_fp32;       // 1: float output (store or atomic add)
  int out_atomic;
  const void* add;    // optional bf16 tensor added to the result before the store (fused residual-gradient sum)
  // stride-2 data gradient (MODE_DGRAD_S2): the output is split into its 4 (h%2, w%2) parity classes; each class is a
  // stride-1 correlation of dy with the subset of filter taps that hit that parity (1+2+2+4 = 9 taps for a 3x3 filter:
  // no zero-upsampled tensor, no multiplications by zero).  Classes ride on the `splits` tile dimension.
  int sub_ntaps[4];           // taps of class c
  unsigned char sub_ow[4][4]; // im2col offsets (relative to the lower corner) of tap t of class c
  unsigned char sub_oh[4][4];
  unsigned char sub_wi[4][4]; // filter tap index r*S+s
  unsigned char sub_ph[4], sub_pw[4];  // output parity of class c
  int sub_H, sub_W;           // full-resolution output extent
  int sub_add_cls;            // class that receives `add` (a COMPACT [N,H/2,W/2,C] tensor), -1: none
  // fprop: per-output-channel sum / sum of squares of the fp32 accumulators, added atomically (fused BN statistics)
  float* st_sum;
  float* st_sq;
  // dgrad (stride 1): "dz epilogue".  The tensor this dgrad produces is the gradient w.r.t. the OUTPUT of a BN(+ReLU)
  // layer.  With bz_y set the epilogue applies that layer's ReLU mask, stores dz = mask * dx instead of dx and reduces
  // sum(dz) and sum(dz * (y - mean)) per channel, i.e. the whole first pass of the BN backward; what remains of it is
  // one elementwise pass (bn_bwd_apply).  bz_out != null: mask from the layer's post-residual activation (out > 0),
  // else recomputed from y.
  const void* bz_y;
  const void* bz_out;
  const float* bz_mean;
  const float* bz_rstd;
  const float* bz_gamma;
  const float* bz_beta;
  float* bz_s0;  // += sum dz * (y - mean)
  float* bz_s1;  // += sum dz
};

// Column sums of a 32x32 fp32 block held one row per lane, through a padded shared-memory tile of 16 rows: the two
// half-warps take turns storing their rows with 8 STS.128 (row pitch 36 floats: conflict-free per quarter-warp), and
// after each turn lane l walks column l (bank 4r+l: conflict-free) accumulating the sum and, for free, the sum of
// squares.  ~120 instructions and no shuffles, against ~280 for two shuffle butterflies (scripts/stats_bench.py); the
// half-height tile keeps the scratch of eight epilogue warps at 18 KB so the wide-tile kernels keep four operand stages.
FD_DEVINL void warp_column_sums_smem(const float (&f)[32], float* buf, int lane, float& sum, float& sumsq) {
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent chains: the adds are latency-bound
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if ((lane >> 4) == half) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(buf + (lane & 15) * 36 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float x = buf[r * 36 + lane];
      a[r & 3] += x;
      b[r & 3] += x * x;
    }
    __syncwarp();
  }
  sum += (a[0] + a[1]) + (a[2] + a[3]);
  sumsq += (b[0] + b[1]) + (b[2] + b[3]);
}

// Shuffle-network variant (no shared-memory traffic, see common.cuh::warp_colsum_shfl): used by the eight-warp
// epilogues, where the extra instructions are spread over two warps per scheduler and the shared-memory bandwidth is
// what the MMAs of the next tile need.  want_sq: also accumulate the sum of squares.
template <bool WANT_SQ>
FD_DEVINL void warp_column_sums_shfl(const float (&f)[32], int lane, float& sum, float& sumsq) {
  float a[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = f[j];
  sum += warp_colsum_shfl(a, lane);
  if (WANT_SQ) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = f[j] * f[j];
    sumsq += warp_colsum_shfl(a, lane);
  }
}

// E8: eight epilogue warps (two per SM sub-partition, BN/2 columns each) instead of four.  One warp per scheduler issues
// its dependent epilogue instructions back to back (IPC ~0.2); a second warp on the same scheduler fills those slots.
// One operand stage is given up for the larger per-warp reduction scratch.
// MT: m-tiles per CTA tile.  MT = 2 makes the CTA tile 256 x BN: both m-tiles consume the SAME filter tile of a k-block,
// which halves the B bytes per MAC; BN = 256 does the same from the other side (and its N=256 MMA is the only shape that
// reaches the tensor-pipe peak: scripts/probe_mma_rate.py measures max(80, N/2) clk per M=128,K=16 MMA).  Both "wide"
// shapes use 256 TMEM columns per accumulator stage (512 double-buffered) and 48 KB of operands per k-block, i.e.
// ~737 clk of TMA fill at the ~66 B/clk an SM can take in (scripts/probe_tma_bw.py) for 512-640 clk of MMA work,
// where the 128x128 tile needs 492 clk of fill for 320 clk of MMA work.
template <int BN, int MODE, bool E8 = false, int MT = 1, bool BZ = false>
struct Cfg {
  static constexpr int kEpiWarps = E8 ? 8 : 4;
  static constexpr int kThreadsCfg = 64 + 32 * kEpiWarps;
  static constexpr int A_BYTES = MT * BM * BK * 2;  // 16 KB per m-tile
  static constexpr int B_BYTES = BN * BK * 2;       // 8/16/32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool kStats = (MODE == 1 /*FPROP*/) || BZ;
  static constexpr int kColredFloats = (kStats && !E8) ? kEpiWarps * 16 * 36 : 4;  // E8 reduces on the shuffle network
  static constexpr int kBnparFloats = BZ ? kEpiWarps * 3 * (E8 ? BN / 2 : BN) : 4;
  static constexpr int kStaticBytes = (kColredFloats + kBnparFloats) * 4 + 64;
  static constexpr int kBudget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - kStaticBytes;
  static constexpr int kStagesMax = kBudget / STAGE_BYTES;
  // The weight-gradient kernels run on a side stream next to the bandwidth-bound BN-backward kernels of the main
  // stream: they keep their operand ring at <= 160 KB so that those CTAs (16-35 KB of shared memory each) can still
  // become resident on the same SM (round 2: with a 224 KB ring the step got SLOWER although the kernel alone was
  // faster).  160 KB in flight is 2400 clk of TMA fill at 66 B/clk, well above the load latency.
  static constexpr int kStagesCap = (MODE == 3 /*WGRAD*/) ? (160 * 1024) / STAGE_BYTES : 8;
  static constexpr int kStages = kStagesMax > kStagesCap ? kStagesCap : kStagesMax;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * MT * BN;  // double-buffered accumulator(s)
  static_assert(MT * BN <= 256, "accumulators of one CTA tile must fit 256 TMEM columns (512 double-buffered)");
  static_assert(kStages >= 3, "too few operand stages");
};

// MC = true: 2-CTA thread-block cluster along M.  The two CTAs work on neighbouring m-tiles of the same n-tile, so
// they need the same B (filter) tile: each CTA fetches HALF of it and the copy engine multicasts it into both CTAs'
// shared memory (.multicast::cluster), halving the L2 traffic of the operand that every m-tile re-reads.  Shared-memory
// slots are released cluster-wide (tcgen05.commit ... .multicast::cluster on both CTAs' empty barriers).
template <int BN, int MODE, bool MC = false, bool BZ = false, bool E8 = false, int MT = 1>
__global__ void __launch_bounds__(E8 ? 320 : kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const KParams P) {
  using C = Cfg<BN, MODE, E8, MT, BZ>;
  static_assert(!(MC && MT > 1), "the multicast-cluster variant pairs m-tiles across CTAs; MT > 1 pairs them inside one");
  constexpr int kStages = C::kStages;
  constexpr int A_TILE = BM * BK * 2;  // one m-tile of a stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kStages * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = (int)uniform_u32(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const uint32_t leader = (lane == 0) ? 1u : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MC ? 2 : 1);  // cluster mode: both CTAs must have consumed a slot before either refills it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 32 * C::kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync();  // peers' barriers are initialised before any remote arrive / multicast copy
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  // everything above is on-chip set-up and may overlap the tail of the previous kernel of the stream (PDL); nothing
  // below touches global memory before the previous grid has completed
  pdl_wait();
  pdl_trigger();
  const int crank = MC ? (int)cluster_ctarank() : 0;

  // a work unit = MT consecutive m-tiles (MT = 2) or, in cluster mode, a pair of m-tiles split over the two CTAs; an
  // m-tile past the end is harmless: its loads are TMA zero-fill and the epilogue masks its rows
  constexpr int kGroup = MC ? 2 : MT;
  const int m_units = (P.m_tiles + kGroup - 1) / kGroup;
  const int total_tiles = m_units * P.n_tiles * P.splits;
  const int t_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ===================================================== TMA producer (convergent warp, lane 0 issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        int split = t / (m_units * P.n_tiles);
        int rem = t - split * (m_units * P.n_tiles);
        const int mt = kGroup * (rem / P.n_tiles) + crank, nt = rem % P.n_tiles;
        const int m0 = mt * BM, n0 = nt * BN;
        // base pixel of every m-tile of this unit (conv modes)
        int bn[MT], bh[MT], bw[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          bn[i] = bh[i] = bw[i] = 0;
          if (MODE == MODE_FPROP || MODE == MODE_DGRAD || MODE == MODE_DGRAD_S2) {
            const int mm = m0 + i * BM;
            bn[i] = mm / P.PQ;
            int r2 = mm - bn[i] * P.PQ;
            int p = r2 / P.Q, q = r2 - p * P.Q;
            bh[i] = P.lower_h + p * P.stride;
            bw[i] = P.lower_w + q * P.stride;
          }
        }
        int kb_begin = 0, kb_end = P.num_kb;
        int tap = 0;
        if (MODE == MODE_DGRAD_S2) kb_end = P.sub_ntaps[split] * P.C_chunks;
        if (MODE == MODE_WGRAD) {
          // n-tile index enumerates (tap, c-chunk); split enumerates pixel ranges
          tap = nt / P.C_chunks;
          kb_begin = (int)(((long)P.kb_total * split) / P.splits);
          kb_end = (int)(((long)P.kb_total * (split + 1)) / P.splits);
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a = smemA + stage * C::A_BYTES;
          uint8_t* b = smemB + stage * C::B_BYTES;
          mbar_arrive_expect_tx_p(&full[stage], C::STAGE_BYTES, leader);
          if (MODE == MODE_GEMM) {
#pragma unroll
            for (int i = 0; i < MT; ++i) tma_load_2d_p(a + i * A_TILE, &mapA, &full[stage], kb * BK, m0 + i * BM, leader);
            if (MC)
              tma_load_2d_mc_p(b + crank * (BN / 2) * 128, &mapB, &full[stage], kb * BK, n0 + crank * (BN / 2), 3, leader);
            else
              tma_load_2d_p(b, &mapB, &full[stage], kb * BK, n0, leader);
          } else if (MODE == MODE_GEMM_NN) {  // A[M,K] K-major, B[K,N] N-major
#pragma unroll
            for (int i = 0; i < MT; ++i) tma_load_2d_p(a + i * A_TILE, &mapA, &full[stage], kb * BK, m0 + i * BM, leader);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_p(b + j * (BK * 128), &mapB, &full[stage], n0 + j * 64, kb * BK, leader);
          } else if (MODE == MODE_GEMM_TN) {  // A[K,M] M-major, B[K,N] N-major
#pragma unroll
            for (int j = 0; j < MT * BM / 64; ++j)
              tma_load_2d_p(a + j * (BK * 128), &mapA, &full[stage], m0 + j * 64, kb * BK, leader);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_p(b + j * (BK * 128), &mapB, &full[stage], n0 + j * 64, kb * BK, leader);
          } else if (MODE == MODE_FPROP) {
            const int tp = kb / P.C_chunks, cc = kb - tp * P.C_chunks;
            const int r = tp / P.S, s = tp - r * P.S;
#pragma unroll
            for (int i = 0; i < MT; ++i)
              tma_load_im2col_4d_p(a + i * A_TILE, &mapA, &full[stage], cc * BK, bw[i], bh[i], bn[i], (uint16_t)s, (uint16_t)r,
                                   leader);
            if (MC)
              tma_load_2d_mc_p(b + crank * (BN / 2) * 128, &mapB, &full[stage], tp * P.Ctot + cc * BK,
                               n0 + crank * (BN / 2), 3, leader);
            else
              tma_load_2d_p(b, &mapB, &full[stage], tp * P.Ctot + cc * BK, n0, leader);
          } else if (MODE == MODE_DGRAD) {
            // reduction over (tap, k-chunk); A = im2col(dY) with flipped taps, B = W[k, (r,s,c)] (N-major)
            const int tp = kb / P.C_chunks, kc = kb - tp * P.C_chunks;
            const int r = tp / P.S, s = tp - r * P.S;
#pragma unroll
            for (int i = 0; i < MT; ++i)
              tma_load_im2col_4d_p(a + i * A_TILE, &mapA, &full[stage], kc * BK, bw[i], bh[i], bn[i], (uint16_t)(P.S - 1 - s),
                                   (uint16_t)(P.R - 1 - r), leader);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_p(b + j * (BK * 128), &mapB, &full[stage], tp * P.Ctot + n0 + j * 64, kc * BK, leader);
          } else if (MODE == MODE_DGRAD_S2) {
            const int tp = kb / P.C_chunks, kc = kb - tp * P.C_chunks;
            const int wi = P.sub_wi[split][tp];
#pragma unroll
            for (int i = 0; i < MT; ++i)
              tma_load_im2col_4d_p(a + i * A_TILE, &mapA, &full[stage], kc * BK, bw[i], bh[i], bn[i],
                                   (uint16_t)P.sub_ow[split][tp], (uint16_t)P.sub_oh[split][tp], leader);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_p(b + j * (BK * 128), &mapB, &full[stage], wi * P.Ctot + n0 + j * 64, kc * BK, leader);
          } else {  // WGRAD: A = dY[pixels, k] (M-major), B = im2col(X)[pixels, c] (N-major), both 64-pixel k-blocks
            const int cc = nt - tap * P.C_chunks;
            const int r = tap / P.S, s = tap - r * P.S;
            const int pix0 = kb * BK;
            int pn = pix0 / P.PQ;
            int r2 = pix0 - pn * P.PQ;
            int p = r2 / P.Q, q = r2 - p * P.Q;
#pragma unroll
            for (int j = 0; j < MT * BM / 64; ++j)
              tma_load_2d_p(a + j * (BK * 128), &mapA, &full[stage], m0 + j * 64, pix0, leader);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)  // one im2col box per 64 channels (N-major 64-wide blocks, LBO apart)
              tma_load_im2col_4d_p(b + j * (BK * 128), &mapB, &full[stage], cc * BN + j * 64, P.lower_w + q * P.stride,
                                   P.lower_h + p * P.stride, pn, (uint16_t)s, (uint16_t)r, leader);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr bool A_MN = (MODE == MODE_WGRAD || MODE == MODE_GEMM_TN);
    constexpr bool B_MN = (MODE == MODE_DGRAD || MODE == MODE_DGRAD_S2 || MODE == MODE_WGRAD || MODE == MODE_GEMM_NN || MODE == MODE_GEMM_TN);
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint64_t descA0 = A_MN ? umma_smem_desc(0, BK * 128, 1024, UMMA_LAYOUT_SW128) : umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t descB0 = B_MN ? umma_smem_desc(0, BK * 128, 1024, UMMA_LAYOUT_SW128) : umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
    const uint32_t smemA16 = smem_u32(smemA) >> 4, smemB16 = smem_u32(smemB) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = t_first; t < total_tiles; t += t_step) {
      int kb_count = P.num_kb;
      if (MODE == MODE_DGRAD_S2) kb_count = P.sub_ntaps[t / (m_units * P.n_tiles)] * P.C_chunks;
      if (MODE == MODE_WGRAD) {
        int split = t / (m_units * P.n_tiles);
        int b0 = (int)(((long)P.kb_total * split) / P.splits);
        int b1 = (int)(((long)P.kb_total * (split + 1)) / P.splits);
        kb_count = b1 - b0;
      }
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (MT * BN);
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        {
          // K-major SW128: 8-row groups 1024B apart, advance 32B per UMMA_K inside the swizzle row.
          // MN-major SW128: tile is [k rows][64 mn elems]; 8-k groups 1024B apart (SBO), 64-mn blocks
          // BK*128B apart (LBO); advance 16 rows = 2048B per UMMA_K.  Descriptors = precomputed base + address
          // field, so the single issuing thread spends ~3 instructions per MMA.
          const uint64_t da0 = descA0 + (uint64_t)(smemA16 + stage * (C::A_BYTES >> 4));
          const uint64_t db0 = descB0 + (uint64_t)(smemB16 + stage * (C::B_BYTES >> 4));
#pragma unroll
          for (int i = 0; i < MT; ++i) {
            // m-tile i: its own A tile and accumulator columns, the same B tile.  For M-major A (wgrad / TN) the second
            // m-tile is the next pair of 64-wide blocks of the stage: 2 * BK*128 bytes further.
            const uint64_t da_i = da0 + (uint64_t)(i * (A_TILE >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss_p(d_tmem + i * BN, da_i + (uint64_t)(k * (A_MN ? 128 : 2)), db0 + (uint64_t)(k * (B_MN ? 128 : 2)),
                            idesc, (kb > 0 || k > 0) ? 1u : 0u, leader);
          }
          if (MC) umma_commit_mc_p(&empty[stage], 3, leader);           // release the slot in both CTAs of the cluster
          else umma_commit_p(&empty[stage], leader);                    // smem slot free once these MMAs retire
          if (kb == kb_count - 1) umma_commit_p(&tfull[acc], leader);   // accumulator(s) ready for the epilogue
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (kb_count == 0) umma_commit_p(&tfull[acc], leader);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================================================== epilogue (warps 2..5, or 2..9 with E8)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr bool bz = BZ && (MODE == MODE_DGRAD);  // dz epilogue compiled into its own instantiation
    const bool do_stats = ((MODE == MODE_FPROP) && P.st_sum != nullptr) || bz;
    float* const red0 = bz ? P.bz_s1 : P.st_sum;  // st_s: sum x (fprop) / sum dz (dgrad)
    float* const red1 = bz ? P.bz_s0 : P.st_sq;   // st_q: sum x^2 (fprop) / sum dz*(y-mean) (dgrad)
    // columns this warp drains: all of them (4 epilogue warps) or one half (8: warps 2-5 the lower, 6-9 the upper half)
    const int c_begin = E8 ? ((warp - 2) >> 2) * (BN / 2) : 0;
    constexpr int c_count = E8 ? BN / 2 : BN;
    // per epilogue warp: mean | gamma*rstd | shift of the channels it drains (static shared: plain LDS, no generic loads)
    __shared__ float bnpar_all[C::kBnparFloats];
    __shared__ __align__(16) float colred_all[C::kColredFloats];
    float* const colred = colred_all + ((C::kStats && !E8) ? (warp - 2) * 16 * 36 : 0);
    float* const bnpar = bnpar_all + (bz ? (warp - 2) * 3 * c_count : 0);
    float st_s[c_count / 32], st_q[c_count / 32];  // this lane's channel (c0 + lane) of each 32-column chunk, over this CTA's tiles
#pragma unroll
    for (int i = 0; i < c_count / 32; ++i) st_s[i] = st_q[i] = 0.f;
    int st_nt = -1;
    auto flush_stats = [&]() {
      if (st_nt >= 0) {
#pragma unroll
        for (int i = 0; i < c_count / 32; ++i) {
          atomicAdd(red0 + st_nt * BN + c_begin + i * 32 + lane, st_s[i]);
          atomicAdd(red1 + st_nt * BN + c_begin + i * 32 + lane, st_q[i]);
          st_s[i] = st_q[i] = 0.f;
        }
      }
    };
    for (int t = t_first; t < total_tiles; t += t_step) {
      int split = t / (m_units * P.n_tiles);
      int rem = t - split * (m_units * P.n_tiles);
      const int mt0 = kGroup * (rem / P.n_tiles) + crank, nt = rem % P.n_tiles;
      if (do_stats && nt != st_nt) {
        flush_stats();
        st_nt = nt;
        if (bz) {  // this warp's copy of the BN coefficients of the channels it drains
          __syncwarp();
          for (int c = lane; c < c_count; c += 32) {
            const int ch = nt * BN + c_begin + c;
            const float mu = P.bz_mean[ch], g = P.bz_gamma[ch] * P.bz_rstd[ch];
            bnpar[c] = mu;
            bnpar[c_count + c] = g;
            bnpar[2 * c_count + c] = P.bz_beta[ch] - mu * g;
          }
          __syncwarp();
        }
      }
      bool has_k = true;
      if (MODE == MODE_WGRAD) {
        int b0 = (int)(((long)P.kb_total * split) / P.splits);
        int b1 = (int)(((long)P.kb_total * (split + 1)) / P.splits);
        has_k = b1 > b0;
      }
#pragma unroll 1
      for (int mi = 0; mi < MT; ++mi) {
        const int m0 = (mt0 + mi) * BM;
        const int row = m0 + quad * 32 + lane;
        // dz epilogue: the BN layer's y tile is requested BEFORE waiting for the accumulator, so its global-memory
        // latency overlaps the MMAs of this tile instead of serialising with the TMEM reads
        uint4 ypre[bz ? c_count / 8 : 1];
        if (bz) {
          const bf16* yrow = reinterpret_cast<const bf16*>(P.bz_y) + (long)row * P.ldo + nt * BN + c_begin;
#pragma unroll
          for (int j = 0; j < c_count / 8; ++j) ypre[j] = (row < P.M) ? *reinterpret_cast<const uint4*>(yrow + j * 8) : make_uint4(0, 0, 0, 0);
          // the residual addend and the mask source are read after the accumulator arrives (no registers to spare for
          // them here): pull their rows into L2 now so those reads do not pay a DRAM round trip per 32-column chunk
          if (row < P.M) {
            const long off = (long)row * P.ldo + nt * BN + c_begin;
            if (P.add) prefetch_l2(reinterpret_cast<const bf16*>(P.add) + off);
            if (P.bz_out) prefetch_l2(reinterpret_cast<const bf16*>(P.bz_out) + off);
          }
        }
        if (mi == 0) {
          mbar_wait(&tfull[acc], acc_phase);
          tc_fence_after();
        }
#pragma unroll
        for (int cc0 = 0; cc0 < c_count; cc0 += 32) {
          const int c0 = c_begin + cc0;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (MT * BN) + mi * BN + c0, v);
          tmem_ld_wait();
          bool added = false;
          if (bz) {
            float f[32], q[32];
            const bool live = row < P.M;
            const long off = (long)row * P.ldo + nt * BN + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 yv = ypre[(cc0 + j) / 8], av = make_uint4(0, 0, 0, 0), ov = av;
              if (live) {
                if (P.add) av = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(P.add) + off + j);
                if (P.bz_out) ov = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(P.bz_out) + off + j);
              }
              const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, aw[4] = {av.x, av.y, av.z, av.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 y2 = unpack_bf16x2(yw[e]), a2 = unpack_bf16x2(aw[e]), o2 = unpack_bf16x2(ow[e]);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const int jj = j + 2 * e + h2;
                  const float yy = h2 ? y2.y : y2.x;
                  const float x = __uint_as_float(v[jj]) + (h2 ? a2.y : a2.x);
                  const float pre = P.bz_out ? (h2 ? o2.y : o2.x) : (yy * bnpar[c_count + cc0 + jj] + bnpar[2 * c_count + cc0 + jj]);
                  const float dz = (live && pre > 0.f) ? x : 0.f;
                  f[jj] = dz;
                  q[jj] = dz * (yy - bnpar[cc0 + jj]);
                  v[jj] = __float_as_uint(dz);
                }
              }
            }
            float unused = 0.f;
            if (E8) {
              warp_column_sums_shfl<false>(f, lane, st_s[cc0 / 32], unused);
              warp_column_sums_shfl<false>(q, lane, st_q[cc0 / 32], unused);
            } else {
              warp_column_sums_smem(f, colred, lane, st_s[cc0 / 32], unused);
              warp_column_sums_smem(q, colred, lane, st_q[cc0 / 32], unused);
            }
            added = true;  // `add` is already inside v
          } else if (do_stats) {  // rows past M are TMA zero-fill: they add nothing
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (E8) warp_column_sums_shfl<true>(f, lane, st_s[cc0 / 32], st_q[cc0 / 32]);
            else warp_column_sums_smem(f, colred, lane, st_s[cc0 / 32], st_q[cc0 / 32]);
          }
          if (row < P.M && has_k) {
            if (MODE == MODE_WGRAD) {
              // row = output channel k; columns = (tap, c) block: out[k][tap*Ctot + cc*BN + c]
              const int tap = nt / P.C_chunks, cc = nt - tap * P.C_chunks;
              float* o = reinterpret_cast<float*>(P.out) + (long)row * P.ldo + (long)tap * P.Ctot + cc * BN + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {  // 16-byte RED / STG: a quarter of the instructions of scalar atomics
                const float4 f4 = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                              __uint_as_float(v[j + 3]));
                if (P.out_atomic)
                  atomicAdd(reinterpret_cast<float4*>(o + j), f4);
                else
                  *reinterpret_cast<float4*>(o + j) = f4;
              }
            } else {
              const int col0 = nt * BN + c0;
              if (P.out_fp32) {
                float* o = reinterpret_cast<float*>(P.out) + (long)row * P.ldo + col0;
                const bool vec = (P.ldo & 3) == 0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  if (vec && col0 + j + 4 <= P.N) {
                    *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      if (col0 + j + e < P.N) o[j + e] = __uint_as_float(v[j + e]);
                  }
                }
              } else {
                long orow = row;  // row of the output matrix this accumulator row lands in
                bool use_add = P.add != nullptr && !added;
                if (MODE == MODE_DGRAD_S2) {
                  const int n_ = row / P.PQ, r2 = row - n_ * P.PQ;
                  const int i_ = r2 / P.Q, j_ = r2 - i_ * P.Q;
                  orow = ((long)n_ * P.sub_H + 2 * i_ + P.sub_ph[split]) * P.sub_W + 2 * j_ + P.sub_pw[split];
                  use_add = use_add && split == P.sub_add_cls;
                }
                bf16* o = reinterpret_cast<bf16*>(P.out) + orow * P.ldo + col0;
                const bool vec = (P.ldo & 7) == 0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  if (!(vec && col0 + j + 8 <= P.N)) {  // ragged / unaligned rows (e.g. 100-class head)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (col0 + j + e < P.N) o[j + e] = __float2bfloat16_rn(__uint_as_float(v[j + e]));
                  } else {
                    if (use_add) {
                      const uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(P.add) + (long)row * P.ldo + col0 + j);
                      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float2 f2 = unpack_bf16x2(aw[e]);
                        v[j + 2 * e] = __float_as_uint(__uint_as_float(v[j + 2 * e]) + f2.x);
                        v[j + 2 * e + 1] = __float_as_uint(__uint_as_float(v[j + 2 * e + 1]) + f2.y);
                      }
                    }
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                    u.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    u.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                    u.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                    *reinterpret_cast<uint4*>(o + j) = u;
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (do_stats) flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync();  // no CTA may exit while its peer can still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled g_encodeTiled = nullptr;
PFN_encodeIm2col g_encodeIm2col = nullptr;
int g_num_sms = 0;
std::mutex g_mu;

int ensure_driver() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_encodeTiled && g_encodeIm2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  FD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  g_encodeTiled = (PFN_encodeTiled)f;
  FD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q));
  g_encodeIm2col = (PFN_encodeIm2col)f;
  int dev = 0;
  FD_CUDA_CHECK(cudaGetDevice(&dev));
  FD_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (!g_encodeTiled || !g_encodeIm2col) {
    fd_set_error("cuTensorMapEncode* entry points unavailable", __FILE__, __LINE__);
    return -1;
  }
  return 0;
}

// 2D tiled map over a row-major [rows, cols] bf16 matrix; box = {64 cols (128B), box_rows}
int make_map_2d(CUtensorMap* m, const void* base, long rows, long cols, long ld_elems, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fd_set_error("cuTensorMapEncodeTiled failed", __FILE__, (int)r);
    return -1;
  }
  return 0;
}

// im2col map over an NHWC bf16 tensor; box = {64 channels, pixels}
int make_map_im2col(CUtensorMap* m, const void* base, int N, int H, int W, int Cc, int lower_w, int lower_h,
                    int upper_w, int upper_h, int stride, int pixels) {
  cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encodeIm2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                              upper, 64, (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fd_set_error("cuTensorMapEncodeIm2col failed", __FILE__, (int)r);
    return -1;
  }
  // driver workaround carried by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors need bit 21 of the
  // second descriptor word cleared on drivers <= 13.1
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (long)N * H * W * Cc * 2 < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  return 0;
}

template <int BN, int MODE, bool BZ = false, bool E8 = false, int MT = 1>
int launch(const CUtensorMap& a, const CUtensorMap& b, const KParams& p, cudaStream_t s) {
  using C = Cfg<BN, MODE, E8, MT, BZ>;
  static_assert(C::SMEM_BYTES + C::kStaticBytes <= 227 * 1024, "shared-memory budget");
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 15]) {
    FD_CUDA_CHECK(cudaFuncSetAttribute(igemm_kernel<BN, MODE, false, BZ, E8, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[dev & 15] = true;
  }
  int tiles = ((p.m_tiles + MT - 1) / MT) * p.n_tiles * p.splits;
  int grid = tiles < g_num_sms ? tiles : g_num_sms;
  if (grid < 1) grid = 1;
  FD_CUDA_CHECK(fd_launch_pdl(MODE == MODE_WGRAD ? PDL_IGEMM_WGRAD : PDL_IGEMM, igemm_kernel<BN, MODE, false, BZ, E8, MT>, dim3(grid), dim3(C::kThreadsCfg), C::SMEM_BYTES, s, a, b, p));
  FD_LAUNCH_OK();
  return 0;
}

// Tile shape of the conv kernels.  Wide CTA tiles (256 accumulator columns: BN = 256, or two m-tiles of BN <= 128 sharing
// every filter tile) halve the operand bytes per MAC and are preferred whenever they still give (nearly) every SM a tile;
// FLUXDIST_WIDE=0 keeps the round-1 128 x BN tiles.
struct TileShape { int bn, mt; };
static int g_wide = -1;
static int g_force_bn = 0, g_force_mt = 0;  // tests / benches: force a tile shape (0 = automatic)
static TileShape pick_shape(long m_tiles, int n_out, bool heavy_epilogue = false) {
  if (g_wide < 0) {
    const char* e = getenv("FLUXDIST_WIDE");
    g_wide = (e && e[0] == '0') ? 0 : 1;
  }
  const int bn_narrow = (n_out % 128 == 0) ? 128 : 64;
  if (g_force_bn > 0 && n_out % g_force_bn == 0 && g_force_bn * g_force_mt <= 256) return {g_force_bn, g_force_mt};
  if (!g_wide) return {bn_narrow, 1};
  auto waves_eff = [&](int bn, int mt) {  // fraction of the SM-time of the last wave that does work
    const long tiles = ((m_tiles + mt - 1) / mt) * (n_out / bn);
    const long waves = (tiles + g_num_sms - 1) / g_num_sms;
    return (double)tiles / (double)(waves * g_num_sms);
  };
  TileShape best{bn_narrow, 1};
  double best_score = waves_eff(bn_narrow, 1) * (bn_narrow == 128 ? 0.65 : 0.55);  // narrow tiles are fill-bound
  if (n_out % 256 == 0) {
    // measured (scripts/conv_bench_r2.py, 14x14x256): plain dgrad 18.2 us with 128x256 vs 19.2 with 256x128; with the
    // BN-statistics epilogue 24.0 vs 22.4 (each epilogue warp drains 128 columns instead of 64)
    const double sc = waves_eff(256, 1) * (heavy_epilogue ? 0.8 : 0.92);
    if (sc > best_score) { best = {256, 1}; best_score = sc; }
  }
  {
    const double sc = waves_eff(bn_narrow, 2) * (bn_narrow == 128 ? 0.87 : 0.6);
    if (sc > best_score) { best = {bn_narrow, 2}; best_score = sc; }
  }
  return best;
}

template <int MODE, bool BZ>
int launch_conv(TileShape ts, bool e8, const CUtensorMap& a, const CUtensorMap& b, const KParams& p, cudaStream_t s) {
  if (ts.bn == 256) return launch<256, MODE, BZ, true, 1>(a, b, p, s);
  if (ts.mt == 2) return ts.bn == 128 ? launch<128, MODE, BZ, true, 2>(a, b, p, s) : launch<64, MODE, BZ, true, 2>(a, b, p, s);
  if (e8) return ts.bn == 128 ? launch<128, MODE, BZ, true, 1>(a, b, p, s) : launch<64, MODE, BZ, true, 1>(a, b, p, s);
  return ts.bn == 128 ? launch<128, MODE, BZ, false, 1>(a, b, p, s) : launch<64, MODE, BZ, false, 1>(a, b, p, s);
}

template <int BN, int MODE>
int launch_mc(const CUtensorMap& a, const CUtensorMap& b, const KParams& p, cudaStream_t s) {
  using C = Cfg<BN, MODE, false, 1, false>;
  FD_CUDA_CHECK(cudaFuncSetAttribute(igemm_kernel<BN, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles * p.splits;
  int clusters = g_num_sms / 2;
  if (clusters > pairs) clusters = pairs;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, igemm_kernel<BN, MODE, true>, a, b, p));
  fd_count_launch();
  return 0;
}

}  // namespace

int fd_num_sms() { return g_num_sms; }

int fd_gemm_nt(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s) {
  if (ensure_driver()) return -1;
  if (K % 64) {
    fd_set_error("gemm_nt: K must be a multiple of 64", __FILE__, __LINE__);
    return -1;
  }
  CUtensorMap ma, mb;
  const int bn = (N <= 64) ? 64 : 128;
  if (make_map_2d(&ma, A, M, K, K, BM) || make_map_2d(&mb, B, N, K, K, bn)) return -1;
  KParams p{};
  p.mode = MODE_GEMM;
  p.M = M; p.N = N;
  p.num_kb = K / 64;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + bn - 1) / bn;
  p.splits = 1;
  p.out = D; p.ldo = N; p.out_fp32 = out_fp32;
  return bn == 64 ? launch<64, MODE_GEMM>(ma, mb, p, s) : launch<128, MODE_GEMM>(ma, mb, p, s);
}

// D[M,N] = A[M,K] * B[K,N]   (A row-major K contiguous, B row-major N contiguous); K need not be a multiple of 64
int fd_gemm_nn(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s) {
  if (ensure_driver()) return -1;
  if (N % 64 || K % 8) {
    fd_set_error("gemm_nn: N must be a multiple of 64 and K of 8", __FILE__, __LINE__);
    return -1;
  }
  const int bn = (N % 128 == 0) ? 128 : 64;
  CUtensorMap ma, mb;
  if (make_map_2d(&ma, A, M, K, K, BM) || make_map_2d(&mb, B, K, N, N, 64)) return -1;
  KParams p{};
  p.mode = MODE_GEMM_NN;
  p.M = M; p.N = N;
  p.num_kb = (K + 63) / 64;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = N / bn;
  p.splits = 1;
  p.out = D; p.ldo = N; p.out_fp32 = out_fp32;
  return bn == 64 ? launch<64, MODE_GEMM_NN>(ma, mb, p, s) : launch<128, MODE_GEMM_NN>(ma, mb, p, s);
}

// D[M,N] = A^T * B with A stored [K,M] (M contiguous) and B stored [K,N] (N contiguous)
int fd_gemm_tn(const void* A, const void* B, void* D, int M, int N, int K, int out_fp32, cudaStream_t s) {
  if (ensure_driver()) return -1;
  if (N % 128 || M % 8) {
    fd_set_error("gemm_tn: N must be a multiple of 128 and M of 8", __FILE__, __LINE__);
    return -1;
  }
  CUtensorMap ma, mb;
  if (make_map_2d(&ma, A, K, M, M, 64) || make_map_2d(&mb, B, K, N, N, 64)) return -1;
  KParams p{};
  p.mode = MODE_GEMM_TN;
  p.M = M; p.N = N;
  p.num_kb = (K + 63) / 64;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = N / 128;
  p.splits = 1;
  p.out = D; p.ldo = N; p.out_fp32 = out_fp32;
  return launch<128, MODE_GEMM_TN>(ma, mb, p, s);
}

static int g_use_mc = 1;
void fd_set_cluster_multicast(int on) { g_use_mc = on; }
void fd_set_tile_shape(int bn, int mt) { g_force_bn = bn; g_force_mt = mt < 1 ? 1 : mt; }
// eight-epilogue-warp instantiations (Cfg::E8) of the conv kernels: the default since round 2 (validated on B200: all
// numerics tests pass, 18.5k -> 18.8k img/s alone, 19.2k with the dz epilogue on every stage); FLUXDIST_EPI8=0 or
// fd_set_epilogue8(0) selects the four-warp epilogues
static int g_epi8 = -1;
void fd_set_epilogue8(int on) { g_epi8 = on; }
static bool use_epi8() {
  if (g_epi8 < 0) {
    const char* e = getenv("FLUXDIST_EPI8");
    g_epi8 = (e && e[0] == '0') ? 0 : 1;
  }
  return g_epi8 == 1;
}

int fd_conv_fprop(const void* x, const void* w, void* y, float* sum, float* sumsq, int N, int H, int W, int C, int K, int R,
                  int S, int stride, int pad, cudaStream_t s) {
  if (ensure_driver()) return -1;
  if (C % 64 || K % 64) {
    fd_set_error("conv_fprop: C and K must be multiples of 64", __FILE__, __LINE__);
    return -1;
  }
  const int P_ = (H + 2 * pad - R) / stride + 1, Q_ = (W + 2 * pad - S) / stride + 1;
  const long m_tiles = ((long)N * P_ * Q_ + BM - 1) / BM;
  const TileShape ts = pick_shape(m_tiles, K, sum != nullptr);
  const int bn = ts.bn;
  CUtensorMap ma, mb;
  // bounding box of base pixels: [-pad, dim-1 + pad-(R-1)] (CUTLASS compute_{lower,upper}_corner_whd, fprop)
  if (make_map_im2col(&ma, x, N, H, W, C, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, BM)) return -1;
  // the 2-CTA cluster + multicast variant (round 1: 28x28x128 25.0 -> 24.2 us) is superseded by the wide tiles, which
  // cut the same filter traffic without a cluster; FLUXDIST_WIDE=0 brings it back
  const bool mc = g_use_mc && !g_wide && bn == 128 && m_tiles >= 2L * g_num_sms;
  if (make_map_2d(&mb, w, K, (long)R * S * C, (long)R * S * C, mc ? bn / 2 : bn)) return -1;
  KParams p{};
  p.mode = MODE_FPROP;
  p.M = N * P_ * Q_; p.N = K;
  p.num_kb = R * S * (C / 64);
  p.m_tiles = (int)m_tiles;
  p.n_tiles = K / bn;
  p.splits = 1;
  p.PQ = P_ * Q_; p.Q = Q_; p.stride = stride; p.lower_w = -pad; p.lower_h = -pad;
  p.S = S; p.R = R; p.C_chunks = C / 64; p.Ctot = C;
  p.out = y; p.ldo = K; p.out_fp32 = 0;
  p.st_sum = sum; p.st_sq = sumsq;  // accumulated into: the caller provides zeroed buffers
  if (mc) return launch_mc<128, MODE_FPROP>(ma, mb, p, s);
  return launch_conv<MODE_FPROP, false>(ts, use_epi8(), ma, mb, p, s);
}

int fd_conv_dgrad(const void* dy, const void* w, void* dx, const void* add, const FdBnEpilogue* bz, int N, int H, int W, int C,
                  int K, int R, int S, int pad, cudaStream_t s) {
  // stride-1 data gradient: dx[N,H,W,C] from dy[N,P,Q,K]; P = H + 2pad - R + 1
  if (ensure_driver()) return -1;
  if (C % 64 || K % 64) {
    fd_set_error("conv_dgrad: C and K must be multiples of 64", __FILE__, __LINE__);
    return -1;
  }
  const int P_ = H + 2 * pad - R + 1, Q_ = W + 2 * pad - S + 1;
  // the dz epilogue is long (mask + two column reductions + up to three extra tensor reads per element) and only hides
  // behind the MMAs of the NEXT tile: it keeps the 128 x BN tiles, which give a CTA 2-4 tiles (scripts/dz_bench.py,
  // 28x28x128: 36.0 us narrow vs 39.5 wide; with the chained mask + addend 49.2 vs 55.1; 14x14 and 7x7: equal or better)
  TileShape ts = pick_shape(((long)N * H * W + BM - 1) / BM, C, bz != nullptr);
  if (bz != nullptr && g_force_bn == 0) ts = {(C % 128 == 0) ? 128 : 64, 1};
  const int bn = ts.bn;
  CUtensorMap ma, mb;
  // CUTLASS dgrad corners: lower = pad-(S-1), upper = pad-(S-1) + (W - Q)
  if (make_map_im2col(&ma, dy, N, P_, Q_, K, pad - (S - 1), pad - (R - 1), pad - (S - 1) + (W - Q_),
                      pad - (R - 1) + (H - P_), 1, BM))
    return -1;
  if (make_map_2d(&mb, w, K, (long)R * S * C, (long)R * S * C, 64)) return -1;
  KParams p{};
  p.mode = MODE_DGRAD;
  p.M = N * H * W; p.N = C;
  p.num_kb = R * S * (K / 64);
  p.m_tiles = (p.M + BM - 1) / BM;
  p.n_tiles = C / bn;
  p.splits = 1;
  p.PQ = H * W; p.Q = W; p.stride = 1; p.lower_w = pad - (S - 1); p.lower_h = pad - (R - 1);
  p.S = S; p.R = R; p.C_chunks = K / 64; p.Ctot = C;
  p.out = dx; p.ldo = C; p.out_fp32 = 0; p.add = add;
  if (bz) {
    p.bz_y = bz->y; p.bz_out = bz->out; p.bz_mean = bz->mean; p.bz_rstd = bz->rstd; p.bz_gamma = bz->gamma;
    p.bz_beta = bz->beta; p.bz_s0 = bz->s0; p.bz_s1 = bz->s1;
    return launch_conv<MODE_DGRAD, true>(ts, use_epi8(), ma, mb, p, s);
  }
  return launch_conv<MODE_DGRAD, false>(ts, use_epi8(), ma, mb, p, s);
}

int fd_conv_dgrad_s2(const void* dy, const void* w, void* dx, const void* add_compact, int N, int H, int W, int C, int K,
                     int R, int S, int pad, cudaStream_t s) {
  // stride-2 data gradient by output-parity decomposition: dx[N,H,W,C] from dy[N,H/2,W/2,K], filter [K,R,S,C], R,S <= 3.
  // Output pixel h = 2i+ph receives tap r iff (ph + pad - r) is even, from dy row i + (ph + pad - r)/2.
  if (ensure_driver()) return -1;
  const int P_ = (H + 2 * pad - R) / 2 + 1, Q_ = (W + 2 * pad - S) / 2 + 1;
  if (C % 64 || K % 64 || R > 3 || S > 3 || (H & 1) || (W & 1) || P_ != H / 2 || Q_ != W / 2) {
    fd_set_error("conv_dgrad_s2: unsupported geometry", __FILE__, __LINE__);
    return -1;
  }
  const TileShape ts = pick_shape(((long)N * P_ * Q_ + BM - 1) / BM, C);
  const int bn = ts.bn;
  int lo_h = 0, lo_w = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int r = 0; r < R; ++r)
      if (((ph + pad - r) & 1) == 0) lo_h = std::min(lo_h, (ph + pad - r) / 2);
  for (int pw = 0; pw < 2; ++pw)
    for (int q = 0; q < S; ++q)
      if (((pw + pad - q) & 1) == 0) lo_w = std::min(lo_w, (pw + pad - q) / 2);
  KParams p{};
  struct Cls { int ph, pw, n; unsigned char ow[4], oh[4], wi[4]; } cls[4];
  int ncls = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      Cls c{ph, pw, 0, {0}, {0}, {0}};
      for (int r = 0; r < R; ++r)
        for (int q = 0; q < S; ++q)
          if (((ph + pad - r) & 1) == 0 && ((pw + pad - q) & 1) == 0) {
            c.oh[c.n] = (unsigned char)((ph + pad - r) / 2 - lo_h);
            c.ow[c.n] = (unsigned char)((pw + pad - q) / 2 - lo_w);
            c.wi[c.n] = (unsigned char)(r * S + q);
            ++c.n;
          }
      cls[ncls++] = c;
    }
  std::sort(cls, cls + 4, [](const Cls& a, const Cls& b) { return a.n > b.n; });  // longest tiles first
  p.sub_add_cls = -1;
  int used = 0;
  for (int i = 0; i < 4; ++i) {
    if (cls[i].n == 0) continue;  // parity classes no tap reaches (1x1 filters): the caller pre-zeroes dx
    p.sub_ntaps[used] = cls[i].n;
    p.sub_ph[used] = (unsigned char)cls[i].ph;
    p.sub_pw[used] = (unsigned char)cls[i].pw;
    for (int t = 0; t < 4; ++t) {
      p.sub_ow[used][t] = cls[i].ow[t];
      p.sub_oh[used][t] = cls[i].oh[t];
      p.sub_wi[used][t] = cls[i].wi[t];
    }
    if (cls[i].ph == 0 && cls[i].pw == 0 && add_compact) p.sub_add_cls = used;
    ++used;
  }
  CUtensorMap ma, mb;
  if (make_map_im2col(&ma, dy, N, P_, Q_, K, lo_w, lo_h, lo_w, lo_h, 1, BM)) return -1;
  if (make_map_2d(&mb, w, K, (long)R * S * C, (long)R * S * C, 64)) return -1;
  p.mode = MODE_DGRAD_S2;
  p.M = N * P_ * Q_; p.N = C;
  p.num_kb = 0;
  p.m_tiles = (p.M + BM - 1) / BM;
  p.n_tiles = C / bn;
  p.splits = used;
  p.PQ = P_ * Q_; p.Q = Q_; p.stride = 1; p.lower_w = lo_w; p.lower_h = lo_h;
  p.S = S; p.R = R; p.C_chunks = K / 64; p.Ctot = C;
  p.sub_H = H; p.sub_W = W;
  p.out = dx; p.ldo = C; p.out_fp32 = 0; p.add = (p.sub_add_cls >= 0) ? add_compact : nullptr;
  return launch_conv<MODE_DGRAD_S2, false>(ts, use_epi8(), ma, mb, p, s);
}

int fd_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int C, int K, int R, int S,
                  int stride, int pad, int splits, int accumulate, cudaStream_t s) {
  // dw[K, R, S, C] (fp32) += sum over pixels; rows (M) = K output channels, columns = (tap, c)
  if (ensure_driver()) return -1;
  if (C % 64 || K % 64) {
    fd_set_error("conv_wgrad: C and K must be multiples of 64", __FILE__, __LINE__);
    return -1;
  }
  const int P_ = (H + 2 * pad - R) / stride + 1, Q_ = (W + 2 * pad - S) / stride + 1;
  const long pixels = (long)N * P_ * Q_;
  const int bn = (C % 128 == 0) ? 128 : 64;
  CUtensorMap ma, mb;
  if (make_map_2d(&ma, dy, pixels, K, K, BK)) return -1;                       // box {64 k, 64 pixels}
  // the im2col box has 64 channels; for bn=128 the kernel would need two boxes -> keep bn=64 boxes per load
  if (make_map_im2col(&mb, x, N, H, W, C, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, BK)) return -1;
  KParams p{};
  p.mode = MODE_WGRAD;
  p.M = K; p.N = R * S * C;
  p.taps = R * S;
  p.kb_total = (int)((pixels + BK - 1) / BK);
  p.m_tiles = (K + BM - 1) / BM;  // K=64: the upper half of the M=128 tile is TMA zero-fill
  // Round 2: N tile of 128 (tap, c) columns wherever C allows.  The kernel is bound by the ~64 B/clk an SM can take in
  // through TMA (scripts/probe_tma_bw.py): a 128x64 tile needs 24 KB per 64-pixel k-block (369 clk of fill for 128
  // clk-equivalents of MMA work), a 128x128 tile 32 KB for twice the work.  Split-K is sized so that ONE wave of CTAs
  // covers the GPU (7x7x512: 4 x 36 tiles, no split; 14x14x256: 36 tiles x 4; 28x28x128: 9 x 16).
  p.C_chunks = C / bn;
  p.n_tiles = R * S * p.C_chunks;
  if (splits < 1) {
    int base = p.m_tiles * p.n_tiles;
    splits = g_num_sms / base;
    if (splits > p.kb_total) splits = p.kb_total;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.num_kb = 0;
  p.PQ = P_ * Q_; p.Q = Q_; p.stride = stride; p.lower_w = -pad; p.lower_h = -pad;
  p.S = S; p.R = R; p.Ctot = C;
  p.out = dw; p.ldo = (long)R * S * C; p.out_fp32 = 1; p.out_atomic = (splits > 1 || accumulate) ? 1 : 0;
  return bn == 128 ? launch<128, MODE_WGRAD>(ma, mb, p, s) : launch<64, MODE_WGRAD>(ma, mb, p, s);
}
