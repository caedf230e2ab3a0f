// ising.cuh -- C4: 2-D Ising model (F/Ising.bl, F/Functions.xtend:8-23) with checkerboard sweeps.
//
// Reference semantics per spin: a CategoricalSampler (R/mcmc/CategoricalSampler.xtend:24-28)
// = IntSliceSampler with the fixed window [0,2) over the spin's 5 connected factors
// (4 annealed edge LogPotentials + 1 fixed Bernoulli node factor).  The GPU keeps that move
// exactly (same draws: Exp(1) slice height, nextInt(2), nextInt(1)) but visits the spins in
// checkerboard order, colour 0 then colour 1, instead of the reference's shuffled order
// (SampledModel.java:258-282): same-colour spins share no factor, so a colour class can be
// updated in parallel and the invariant law is unchanged.  The Philox
// counter of a spin update is (draw, pass*nSpins + v, scan, position), so the oracle's
// ORC_ORDER_CHECKERBOARD sweep reproduces it draw for draw.
//
// One CTA per chain; the whole lattice (1 byte per spin) is staged in shared memory for the
// sweep (256x256 = 64 KB) and written back once.  Energy by an integer block reduction.
#pragma once
#include <algorithm>
#include "kernels.cuh"

namespace bpt {

struct IsingData {
  uint8_t* spins;   // [localSlots][N*N], row-major
  int N;
  double coupling, moment, p1;   // p1 = logistic(-2*moment) = P(spin = 1) under the node factor
};

constexpr int kIsingThreads = 1024;

__global__ void __launch_bounds__(kIsingThreads) ising_kernel(DevEngine E, IsingData D, int mode) {
  extern __shared__ uint8_t sp[];
  __shared__ long long red_i[32];
  __shared__ unsigned long long red_u[32];
  const int tid = threadIdx.x, l = blockIdx.x, N = D.N, S = N * N;
  const int g = E.slot_lo + l;
  const int rep = g / E.N, s_in_rep = g - rep * E.N;
  const int pos = E.pos_of_slot[rep * E.N + s_in_rep];
  const double beta = E.beta_override == E.beta_override ? E.beta_override : E.ladder[rep * E.N + pos];
  const uint32_t scan = E.scan_counter[rep];
  const uint32_t key1 = E.key1 ^ (E.replica_offset + (uint32_t)rep);
  uint8_t* gs = D.spins + (size_t)l * S;

  for (int v = tid; v < S; v += blockDim.x) sp[v] = gs[v];
  __syncthreads();

  unsigned long long evals = 0, execs = 0;
  const double lp1 = log(D.p1), lp0 = log(1.0 - D.p1);
  const bool forward = (mode == MODE_FORWARD_INIT) || (mode == MODE_EXPLORE && beta == 0.0 && E.use_prior);
  if (forward) {   // generate(rand) of each Bernoulli node in declaration order (Bernoulli.bl:17-19)
    for (int v = tid; v < S; v += blockDim.x) {
      Rng rng(E.key0, key1, mode == MODE_FORWARD_INIT ? P_INIT : P_FORWARD, (uint32_t)pos,
              mode == MODE_FORWARD_INIT ? 0u : scan, 0u);
      rng.draw = (uint32_t)v;
      sp[v] = rng.bernoulli(D.p1) ? 1 : 0;
    }
    __syncthreads();
  } else if (mode == MODE_EXPLORE) {
    // steps_override (SCM: "one sampler step per temperature") is served by one full checkerboard pass
    const int n_pass = E.steps_override > 0 ? 1 : (int)floor(E.n_passes);
    const double bj = beta * D.coupling;
    const int half = S / 2;
    for (int pass = 0; pass < n_pass; pass++)
      for (int colour = 0; colour < 2; colour++) {
        for (int q = tid; q < half; q += blockDim.x) {
          const int i = (2 * q) / N;
          const int j = ((2 * q) % N) + ((i + colour) & 1);
          const int v = i * N + j;
          const int up = (i == 0 ? N - 1 : i - 1) * N + j, dn = (i == N - 1 ? 0 : i + 1) * N + j;
          const int lf = i * N + (j == 0 ? N - 1 : j - 1), rt = i * N + (j == N - 1 ? 0 : j + 1);
          const int nb = 2 * ((int)sp[up] + (int)sp[dn] + (int)sp[lf] + (int)sp[rt]) - 4;
          Rng rng(E.key0, key1, P_MOVE, (uint32_t)pos, scan, (uint32_t)(pass * S + v));
          auto lp = [&](int x) -> double {
            evals++;
            const double fx = x == 1 ? lp1 : lp0;
            if (beta == 0.0) return fx;
            return fx + bj * (double)((2 * x - 1) * nb);
          };
          const int x0 = sp[v];
          int serr = 0;   // a spin window [0, 2) always holds the current state, whose density is finite
          const int xn = int_slice(lp, rng, x0, true, 0, 2, serr);
          execs++;
          sp[v] = (uint8_t)xn;
        }
        __syncthreads();
      }
  }

  // full evaluation: loglik = J * sum_edges s_u s_v (pre-annealing), fixed = sum_v log p(s_v)
  long long agree = 0, ones = 0;
  for (int v = tid; v < S; v += blockDim.x) {
    const int i = v / N, j = v - i * N;
    const int s = 2 * (int)sp[v] - 1;
    const int r = 2 * (int)sp[i * N + (j == N - 1 ? 0 : j + 1)] - 1;
    const int d = 2 * (int)sp[(i == N - 1 ? 0 : i + 1) * N + j] - 1;
    agree += s * (r + d);
    ones += sp[v];
    if (mode != MODE_EVAL) gs[v] = sp[v];
  }
  long long packed = agree * 4294967296LL + ones;   // |agree| <= 2S < 2^31, ones <= S
  const int lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) { packed += __shfl_xor_sync(0xffffffffu, packed, o); evals += __shfl_xor_sync(0xffffffffu, evals, o); execs += __shfl_xor_sync(0xffffffffu, execs, o); }
  __shared__ unsigned long long red_x[32];
  if (lane == 0) { red_i[warp] = packed; red_u[warp] = evals; red_x[warp] = execs; }
  __syncthreads();
  if (tid == 0) {
    long long tot = 0; unsigned long long ev = 0, ex = 0;
    for (int w = 0; w < nw; w++) { tot += red_i[w]; ev += red_u[w]; ex += red_x[w]; }
    // decode: ones is the low 32 bits (non-negative), agree the signed high part
    const long long n1 = tot & 0xffffffffLL;
    const long long ag = (tot - n1) / 4294967296LL;
    const double loglik = D.coupling * (double)ag;
    const double fixed = (double)n1 * lp1 + (double)(S - n1) * lp0;
    double* t = E.table + (size_t)g * 4;
    t[0] = loglik; t[1] = 0.0; t[2] = fixed; t[3] = (double)ev;
    double* sd = E.send + (size_t)l * 4;
    sd[0] = loglik; sd[1] = 0.0; sd[2] = fixed; sd[3] = (double)ev;
    E.evals[l] += ev;
    E.execs[l] += ex;
    peer_publish(E, g, pos, loglik, 0, fixed, ev, nullptr, 0);
  }
}

__global__ void ising_record_kernel(DevEngine E, IsingData D, uint8_t* out, int scan_in_batch) {
  const int rep = blockIdx.x, S = D.N * D.N;
  const int s = E.slot_of_pos[(size_t)rep * E.N];   // position 0 = beta 1
  const int l = rep * E.N + s - E.slot_lo;
  if (l < 0 || l >= E.M_local) return;
  uint8_t* dst = out + ((size_t)scan_in_batch * E.R + rep) * S;
  const uint8_t* src = D.spins + (size_t)l * S;
  for (int v = threadIdx.x; v < S; v += blockDim.x) dst[v] = src[v];
}

inline int ising_launch(const DevEngine& E, const IsingData& D, int mode, cudaStream_t stream) {
  const size_t smem = (size_t)D.N * D.N;
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  if (smem > 48 * 1024) {   // per device: set on every launch (cheap) rather than remembering which devices have it
    cudaError_t e = cudaFuncSetAttribute(ising_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  const int threads = D.N * D.N / 2 >= kIsingThreads ? kIsingThreads : std::max(32, ((D.N * D.N / 2 + 31) / 32) * 32);
  ising_kernel<<<E.M_local, threads, smem, stream>>>(E, D, mode);
  return (int)cudaGetLastError();
}

inline void ising_record_target(const DevEngine& E, const IsingData& D, uint8_t* out, int scan_in_batch, cudaStream_t stream) {
  ising_record_kernel<<<E.R, 256, 0, stream>>>(E, D, out, scan_in_batch);
}

}  // namespace bpt
