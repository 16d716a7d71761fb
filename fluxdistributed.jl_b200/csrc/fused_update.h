// Host/device argument block of the fused all-reduce + optimiser kernel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FD_MAX_WORLD 8
enum { FD_RULE_DESCENT = 0, FD_RULE_MOMENTUM = 1, FD_RULE_NESTEROV = 2, FD_RULE_ADAM = 3, FD_RULE_ADAMW = 4 };
enum { FD_MODE_TWO_SHOT = 0, FD_MODE_ONE_SHOT = 1 };

struct FusedArgs {
  int world, rank, rule, mode, use_nvls;
  int unroll;                           // 16-byte requests per thread per iteration (2 or 4)
  long n;                               // elements in this bucket (multiple of 4*world)
  long fp32_bcast_from;                 // elements >= this offset are also broadcast in fp32
  float grad_scale;                     // extra scale on the mean gradient (loss-scale inverse)
  float h[8];                           // eta, a(rho|beta1), b(beta2), eps, decay (used when hyper_dev is null)
  const float* hyper_dev;               // device copy of {h[0..4], -, -, grad_scale}: read at run time, so a captured
                                        // CUDA graph follows set_opt / LR schedules without re-capture
  const float* g[FD_MAX_WORLD];         // gradient arena of every rank (peer-mapped), bucket base
  float* w32;                           // local fp32 master
  float* s0;                            // local moment 1 (momentum / Adam m)
  float* s1;                            // local moment 2 (Adam v)
  float* wout32[FD_MAX_WORLD];          // fp32 compute weights of every rank (nullable)
  __nv_bfloat16* wout16[FD_MAX_WORLD];  // bf16 compute weights of every rank (nullable)
  uint32_t* flags[FD_MAX_WORLD];        // signal pads: [grid][FD_MAX_WORLD] u32 per rank
  uint32_t* epoch;                      // local per-CTA barrier epochs [grid]
  int* step_dev;                        // local optimiser step counter
  unsigned* done_ctr;                   // local CTA completion counter
  int* status;                          // [0]=non-finite gradient seen, [1]=barrier timeout
  long long max_spins;                  // bounded spin per barrier wait
  const float* mc_g;                    // multicast (NVLS) addresses, bucket base (nullable)
  float* mc_w32;
  __nv_bfloat16* mc_w16;
};

int fd_fused_update_grid(long n, int world, int mode, int max_grid);
int fd_fused_update_launch(const FusedArgs& a, int grid, cudaStream_t stream);
int fd_gather_shards(float* dst, const float* const* src_dev_array, int world, int rank, long n, cudaStream_t s);
