// reduce.cuh -- deterministic (fixed-order) sum reductions for one chain group.
//
// A chain is owned by one thread-block cluster of G CTAs (G = 1, 2, 4 or 8).
// Every evaluation of a sampler's log-density is a reduction over the chain's
// connected factors, and every thread of the group must see the same total
// because the slice sampler's control flow (stepping out, shrinkage, the
// doubling acceptance test) is executed redundantly by all of them.
//   lane level   : xor-butterfly with __shfl_xor_sync (all lanes get the sum)
//   CTA level    : one partial per warp in shared memory, one __syncthreads,
//                  then every warp re-reduces the partials with a butterfly
//   cluster level: thread 0 pushes its CTA total into every peer's shared
//                  memory (DSMEM), one cluster barrier, fixed-order sum.
// Buffers are double-buffered on `parity`, so one barrier per reduction is
// enough.  No floating-point atomics anywhere: results are bit-reproducible
// for a fixed (THREADS, G).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace bpt {
namespace cg = cooperative_groups;

struct ReduceSmem {
  double wsum[2][32];
  int wcnt[2][32];
  double csum[2][8];
  int ccnt[2][8];
};

struct GroupReducer {
  ReduceSmem* sm;
  int parity;
  int G, crank;   // cluster size and this CTA's rank in it
  int nwarps;

  BPT_D void init(ReduceSmem* s, int G_, int crank_) {
    sm = s; parity = 0; G = G_; crank = crank_; nwarps = (blockDim.x + 31) >> 5;
  }
  // one chain per warp (fused small-model kernel): shuffles only, no shared memory, no barrier
  BPT_D void init_warp() { sm = nullptr; parity = 0; G = 1; crank = 0; nwarps = 1; }

  // all threads of the group call this; on return s / c hold the group totals
  BPT_D void reduce(double& s, int& c) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(full, s, o);
      c += __shfl_xor_sync(full, c, o);
    }
    if (nwarps > 1) {
      if (lane == 0) { sm->wsum[parity][warp] = s; sm->wcnt[parity][warp] = c; }
      __syncthreads();
      s = lane < nwarps ? sm->wsum[parity][lane] : 0.0;
      c = lane < nwarps ? sm->wcnt[parity][lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(full, s, o);
        c += __shfl_xor_sync(full, c, o);
      }
    }
    if (G > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      if (threadIdx.x == 0) {
        for (int r = 0; r < G; r++) {
          ReduceSmem* peer = cluster.map_shared_rank(sm, r);
          peer->csum[parity][crank] = s;
          peer->ccnt[parity][crank] = c;
        }
      }
      cluster.sync();
      double t = 0.0; int tc = 0;
      for (int r = 0; r < G; r++) { t += sm->csum[parity][r]; tc += sm->ccnt[parity][r]; }
      s = t; c = tc;
    }
    parity ^= 1;
  }
};

}  // namespace bpt
