// Data-parallel gradient sum over NVLink / NVSwitch peer memory: one kernel per gradient bucket that
// reads every rank's bucket straight out of the peers' HBM (or lets the switch add them: NVLS multimem) instead of an
// NCCL ring / tree launch.  The reference has no distributed code (single cuda:0, train.py:181); the contract is the
// north star's: shard by video, SUM the 176 gradient tensors (one flat fp32 buffer here) across ranks.
//
// Every rank launches the same grid over the same bucket.  CTA j of every rank owns slice j of the bucket and talks only
// to CTA j of the peers, through per-(channel, CTA, phase, source rank) epoch words in each rank's flag area:
//   phase 0  "my bucket is final"   (stream order put this kernel behind the kernels that wrote it)
//   one-shot (peer reads)           acc = sum over ranks r = 0..W-1, in rank order, of peer_r[i]  -> identical bits on every rank
//   phase 1  "I have read yours"    then the sum is written over the local bucket (in place: no second buffer)
//   NVLS                            rank r adds sub-slice r of slice j inside the switch (multimem.ld_reduce) and
//                                   broadcasts it to all ranks (multimem.st); phase 1 = "my stores are out"
// The CTAs are 128 threads with a small register budget and no shared memory, so that one fits beside a resident
// chain-kernel CTA (352 threads, 224 KB smem): the sum of stage s+1's gradients hides under stage s's backward chain.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mstcn {
namespace dp {

constexpr int kThreads = 128;
constexpr int kMaxCtas = 160;        // flag rows per channel
constexpr int kMaxRanks = 16;
constexpr int kChannels = 8;         // concurrent buckets in flight (one per stage bucket + spare)
constexpr int kVecPerThread = 4;     // float4 accumulators per thread and round
// per-rank flag area (uint32 words): [channel][cta][phase 2][source rank] epoch words, then [channel][cta] local epochs
constexpr int64_t kFlagWords = (int64_t)kChannels * kMaxCtas * 2 * kMaxRanks + kChannels * kMaxCtas;
__host__ __device__ inline int64_t flag_slot(int ch, int cta, int phase, int src) {
  return (((int64_t)ch * kMaxCtas + cta) * 2 + phase) * kMaxRanks + src;
}
__host__ __device__ inline int64_t epoch_slot(int ch, int cta) {
  return (int64_t)kChannels * kMaxCtas * 2 * kMaxRanks + (int64_t)ch * kMaxCtas + cta;
}

struct DpArgs {
  float* const* bufs;          // device array [world]: every rank's flat gradient buffer (peer-mapped, same layout)
  uint32_t* const* flags;      // device array [world]: every rank's flag area (kFlagWords words, zeroed once)
  float* mc;                   // NVLS multicast mapping of the buffers (NULL: one-shot peer reads)
  long long offset, n;         // bucket = floats [offset, offset + n) of the flat buffer
  int rank, world, channel;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// all ranks' CTA j meet: thread t signals rank t and waits for rank t's word.  Bounded spin: a rank that never arrives
// (a crashed peer) ends in a trap after ~4 s instead of a hung GPU.
__device__ __forceinline__ void cta_rendezvous(const DpArgs& a, int phase, uint32_t epoch) {
  const int t = threadIdx.x;
  __syncthreads();                                 // every thread's earlier reads / writes are issued
  if (t < a.world && t != a.rank) {
    st_release_sys(a.flags[t] + flag_slot(a.channel, blockIdx.x, phase, a.rank), epoch);
    const uint32_t* mine = a.flags[a.rank] + flag_slot(a.channel, blockIdx.x, phase, t);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {        // epochs only grow: a peer may already be one ahead
      __nanosleep(64);
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 10) dp_allreduce_kernel(DpArgs a) {
  const int tid = threadIdx.x, W = a.world;
  uint32_t* const my_flags = a.flags[a.rank];
  uint32_t epoch = my_flags[epoch_slot(a.channel, blockIdx.x)];       // only this CTA index ever touches the word
  float* const mine = a.bufs[a.rank] + a.offset;
  // slice of this CTA, in float4 units when the bucket is 16-byte aligned (it is for every supported shape)
  const bool vec = ((a.offset | a.n) & 3) == 0;
  const long long units = vec ? a.n / 4 : a.n;
  const long long per_cta = (units + gridDim.x - 1) / gridDim.x;
  const long long u0 = per_cta * blockIdx.x, u1 = (u0 + per_cta < units) ? u0 + per_cta : units;
  const long long round_units = (long long)kThreads * kVecPerThread;
  const long long rounds = (per_cta + round_units - 1) / round_units;   // identical on every rank and CTA

  cta_rendezvous(a, 0, ++epoch);                                        // every rank's bucket is final
  if (a.mc != nullptr && vec) {
    // NVLS: this rank adds sub-slice `rank` of the CTA's slice inside the switch and broadcasts the sum to every rank;
    // nothing is held in registers, so one pass and one closing rendezvous
    const long long len = u1 > u0 ? u1 - u0 : 0;
    const long long lo = u0 + (len * a.rank) / W, hi = u0 + (len * (a.rank + 1)) / W;
    for (long long u = lo + tid; u < hi; u += kThreads) {
      float* p = a.mc + a.offset + 4 * u;
      multimem_st_v4(p, multimem_ld_reduce_v4(p));
    }
    __threadfence_system();
    cta_rendezvous(a, 1, ++epoch);                                      // every rank's multicast stores are out
  } else {
    for (long long r = 0; r < rounds; ++r) {
      const long long base = u0 + r * round_units;
      if (vec) {
        float4 acc[kVecPerThread];
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
          const long long u = base + (long long)k * kThreads + tid;
          acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (u < u1)
            for (int rk = 0; rk < W; ++rk) {
              const float4 v = ld_peer_v4(a.bufs[rk] + a.offset + 4 * u);
              acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
            }
        }
        cta_rendezvous(a, 1, ++epoch);                                  // every rank has read this round: overwrite
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
          const long long u = base + (long long)k * kThreads + tid;
          if (u < u1) reinterpret_cast<float4*>(mine)[u] = acc[k];
        }
      } else {
        float acc[kVecPerThread];
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
          const long long u = base + (long long)k * kThreads + tid;
          acc[k] = 0.f;
          if (u < u1)
            for (int rk = 0; rk < W; ++rk) acc[k] += ld_peer_f(a.bufs[rk] + a.offset + u);
        }
        cta_rendezvous(a, 1, ++epoch);
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
          const long long u = base + (long long)k * kThreads + tid;
          if (u < u1) mine[u] = acc[k];
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) my_flags[epoch_slot(a.channel, blockIdx.x)] = epoch;
}

}  // namespace dp
}  // namespace mstcn
