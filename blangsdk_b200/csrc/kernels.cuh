// kernels.cuh -- the three kernels of one NRPT scan.
//
//   explore_kernel<Fam>  local exploration of every chain slot: the reference's
//                        ParallelTempering.moveKernel (R/engines/ParallelTempering.java:77-92)
//                        = SampledModel.posteriorSamplingScan (R/runtime/SampledModel.java:249-282)
//                        or forwardSample at beta = 0 (:284-291).  One thread-block
//                        cluster per chain; ends with a from-scratch evaluation of
//                        (loglik, nOutOfSupport, fixed) into the per-slot table.
//   swap_kernel          energy accumulators + DEO swap pass + index process:
//                        ParallelTempering.swapKernel / swapAcceptPr / doSwap (:64-75,99-124,155-164),
//                        the accumulators of :88-90 and Paths.nRejuvenations
//                        (R/engines/internals/ptanalysis/Paths.xtend:40-50).  Swaps permute
//                        temperature POSITIONS between fixed slots; states never move.
//   reset_round_kernel   ParallelTempering.setAnnealingParameters :197-204.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "slice.cuh"
#include "families.cuh"

namespace bpt {

enum ExploreMode : int { MODE_EXPLORE = 0, MODE_EVAL = 1, MODE_FORWARD_INIT = 2 };

// ---- multi-GPU exchange through peer-mapped mailboxes (one process per GPU, no collective library) ----
// After its scan a chain's 32-byte record (loglik, nOutOfSupport, fixed, evaluations) is STORED into the mailbox of
// every rank (NVLink peer writes through CUDA-IPC mappings), as two 16-byte halves that each carry the exchange's
// sequence number, so a reader that sees the number it waits for in both halves holds the whole record.  The chain
// at position 0 also publishes its reals (one 16-byte {value, sequence, index} record each): they are the sample the
// swap pass records.  gather_kernel then fills the replicated table; two parities of buffers, because a fast rank may
// already publish exchange x+1 while a slow one still reads exchange x (it cannot get two ahead: its next gather
// needs the slow rank's records).
constexpr int kMaxPeers = 16;
struct __align__(16) MailHalf { double v; uint32_t seq; int i; };
struct PeerBox {
  int world;                 // 0 = exchange off (single process, or split-phase driven by the host)
  uint32_t xseq;             // sequence number of the exchange the NEXT exploration pass publishes (host-incremented)
  int n_reals;
  MailHalf* record[kMaxPeers];   // rank r's mailbox: [2 parities][M][2 halves]
  MailHalf* sample[kMaxPeers];   // rank r's sample mailbox: [2 parities][R][n_reals]
  MailHalf* my_record; MailHalf* my_sample;
  double* sample_buf;        // [R][n_reals] the gathered position-0 reals (read by the swap pass)
};
BPT_D void peer_st(MailHalf* p, double v, uint32_t seq, int i) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"((uint32_t)__double2loint(v)),
               "r"((uint32_t)__double2hiint(v)), "r"(seq), "r"((uint32_t)i) : "memory");
}
BPT_D MailHalf peer_ld(const MailHalf* p) {
  uint32_t a, b, c, d;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  MailHalf h; h.v = __hiloint2double((int)b, (int)a); h.seq = c; h.i = (int)d;
  return h;
}

struct DevEngine {
  int N, R;                 // chains per replica, replicas in this handle
  int slot_lo, M_local;     // first local global-slot, number of local slots
  uint32_t key0, key1;      // Philox key: (seed lo, seed hi); replica folded into key1
  uint32_t replica_offset;
  double n_passes;
  double beta_override;     // NaN = use the ladder; otherwise every chain runs at this exponent (SCM particles)
  int steps_override;       // 0 = floor(n_passes * S) sampler executions per scan; > 0 = exactly that many
  int use_prior, anneal_support;
  int nan_neginf;           // treatNaNAsNegativeInfinity: a NaN factor counts as -inf (ExponentiatedFactor.enclosedLogDensity)
  int s_max;                // stride of the order array
  const double* ladder;     // [R][N] descending
  int* pos_of_slot;         // [R][N] (slot-in-replica -> position)
  int* slot_of_pos;         // [R][N]
  double* table;            // [R*N][4]: loglik, noos, fixed, evals-this-scan
  double* send;             // [M_local][4] copy of the local segment (multi-GPU all-gather source)
  double* prev_loglik;      // [R*N] loglik at the start of the scan (energy "before")
  double* reals;            // [nReals][M_local] structure of arrays
  int* order;               // [M_local][s_max] SampledModel.currentSamplingOrder
  int* curpos;              // [M_local]        SampledModel.currentPosition
  unsigned long long* evals;// [M_local]
  unsigned long long* execs;// [M_local]
  int* err;                 // sticky error bits
  uint32_t* scan_counter;   // [R] scans done (Philox `scan`)
  uint32_t* swap_counter;   // [R] ParallelTempering.swapIndex
  uint32_t* batch_counter;  // [R] scan index inside the current run_scans batch
  PeerBox peers;            // multi-GPU exchange (peers.world == 0: off)
};

struct DevStats {     // all indexed [R][N] by temperature position unless noted
  long long* en_n; double* en_m1; double* en_m2;              // energies (SummaryStatistics)
  double* cov_mx; double* cov_my; double* cov_c; int* cov_n;  // CovarAccumulator
  long long* acc_n; double* acc_m1;                           // swapAcceptPrs
  double* ls_sum; long long* ls_n;                            // LogSumAccumulator
  int* path_flag;                                             // [R][N] by slot: Paths "newSample"
  unsigned long long* rejuv;                                  // [R]
  unsigned long long* total_evals;                            // [1]
  int* oos_seen;                                              // [R] sticky SampledModel.annealedOutOfSupportDetected
};

struct DevOut {       // per-scan outputs of the current batch (nullable)
  double* energy; double* logdens; int* noos; int* indicators; double* samples; int n_reals;
  int reversible;
};

// how the threads that share one chain synchronise and index themselves
template <class Fam> struct LogDensityInline;
template <class Fam> struct LogDensityCall;
struct CtaSync {    // a chain owns whole CTAs (a cluster of G of them)
  template <class Fam> using LogDensityOf = LogDensityInline<Fam>;
  BPT_D static void sync() { __syncthreads(); }
  BPT_D static int tid() { return threadIdx.x; }
  BPT_D static int nthr() { return blockDim.x; }
};
struct WarpSync {   // a chain owns one warp (fused replica kernel)
  template <class Fam> using LogDensityOf = LogDensityCall<Fam>;
  BPT_D static void sync() { __syncwarp(); }
  BPT_D static int tid() { return threadIdx.x & 31; }
  BPT_D static int nthr() { return 32; }
};

// RealSliceSampler.logDensity :156-161 over ExponentiatedFactor.logDensity :70-80: the sampler's view of the
// annealed density with its variable set to x.  Deliberately NOT inlined: the slice sampler calls it from
// seven places; one shared copy keeps the kernel's code (and instruction-cache footprint, which dominated the
// many-small-replicas kernel: ncu stall_no_instruction) small and gives the row loop its own registers.
template <class Fam>
struct LogDensityBase {
  const DevEngine& E; Fam& fam; GroupReducer& red;
  int k, aux; double beta;
  unsigned long long& evals; int& err;
  BPT_D double eval(double x) {
    evals++;
    double fx = fam.fixed(k, x, aux);
    if (fx != fx) { if (E.nan_neginf) fx = BPT_NEG_INF; else err |= ERR_NAN; }   // ExponentiatedFactor.java:26-29
    if (beta == 0.0 || fx == BPT_NEG_INF) return fx;
    double s; int c;
    fam.annealed(k, x, aux, red, s, c);
    double a = beta * s;
    if (c) a += E.anneal_support ? c * annealed_minus_infinity(beta) : BPT_NEG_INF;
    double r = fx + a;
    if (r != r) { if (E.nan_neginf) r = BPT_NEG_INF; else err |= ERR_NAN; }   // ExponentiatedFactor.java:26-29
    return r;
  }
};
// inlined at every call site: best for the big-data kernels (the row loop keeps its operands in registers)
template <class Fam>
struct LogDensityInline : LogDensityBase<Fam> {
  BPT_D double operator()(double x) { return this->eval(x); }
};
// one shared copy: best for the fused replica kernel, whose 8-32 warps run different parts of the sampler
// at the same time (inlined: ncu stall_no_instruction = 12 warps per issue; outlined: 2x faster)
template <class Fam>
struct LogDensityCall : LogDensityBase<Fam> {
  __device__ __noinline__ double operator()(double x) { return this->eval(x); }
};

// one thread: publish global slot g's record (and, for the chain at position 0, its reals) to every rank
BPT_D void peer_publish(const DevEngine& E, int g, int pos, double loglik, int noos, double fixed, unsigned long long evals,
                        const double* P, int n_reals) {
  const PeerBox& B = E.peers;
  if (B.world == 0) return;
  const size_t M = (size_t)E.R * E.N, par = B.xseq & 1u;
  for (int r = 0; r < B.world; r++) {
    MailHalf* rec = B.record[r] + (par * M + (size_t)g) * 2;
    peer_st(rec, loglik, B.xseq, noos);
    peer_st(rec + 1, fixed, B.xseq, (int)evals);
    if (pos == 0 && P) {
      const int rep = g / E.N;
      MailHalf* sm = B.sample[r] + (par * E.R + rep) * (size_t)B.n_reals;
      for (int v = 0; v < n_reals; v++) peer_st(sm + v, P[v], B.xseq, v);
    }
  }
}

// Fills the replicated table (and the position-0 sample buffer) from this rank's mailboxes once every rank's
// records of exchange `xseq` have arrived.  A record that does not arrive within ~20 s (a rank died or the host
// code is not SPMD) raises ERR_PEER_TIMEOUT instead of hanging.
static __global__ void gather_kernel(DevEngine E) {
  const PeerBox& B = E.peers;
  const size_t M = (size_t)E.R * E.N, par = B.xseq & 1u;
  const size_t n_rec = M, n_smp = (size_t)E.R * B.n_reals;
  const long long t0 = clock64();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_rec + n_smp; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_rec) {
      const MailHalf* rec = B.my_record + (par * M + i) * 2;
      MailHalf a, b;
      for (;;) {
        a = peer_ld(rec); b = peer_ld(rec + 1);
        if (a.seq == B.xseq && b.seq == B.xseq) break;
        __nanosleep(200);
        if (clock64() - t0 > 40LL * 1000 * 1000 * 1000) { atomicOr(E.err, ERR_PEER_TIMEOUT); return; }
      }
      double* t = E.table + i * 4;
      t[0] = a.v; t[1] = (double)a.i; t[2] = b.v; t[3] = (double)b.i;
    } else {
      const size_t k = i - n_rec;
      const MailHalf* sm = B.my_sample + par * n_smp + k;
      MailHalf a;
      for (;;) {
        a = peer_ld(sm);
        if (a.seq == B.xseq) break;
        __nanosleep(200);
        if (clock64() - t0 > 40LL * 1000 * 1000 * 1000) { atomicOr(E.err, ERR_PEER_TIMEOUT); return; }
      }
      B.sample_buf[k] = a.v;
    }
  }
}

// One chain's share of ParallelTempering.moveKernel for one scan: forward sample at beta = 0, else
// floor(nPasses * S) sampler executions in the persistent shuffled order.  P / order are the chain's
// parameter and sampler-order arrays (shared memory); every thread of the chain group runs this.
// families with latent integers declare kHasIntSlice; the branch is not compiled into the other kernels
template <class Fam, class = void> struct HasIntSlice { static constexpr bool value = false; };
template <class Fam> struct HasIntSlice<Fam, decltype((void)Fam::kHasIntSlice)> { static constexpr bool value = Fam::kHasIntSlice; };
// families whose latent integers move by CategoricalSampler / UniformSampler (IntSliceSampler with a fixed window)
template <class Fam, class = void> struct HasIntFixed { static constexpr bool value = false; };
template <class Fam> struct HasIntFixed<Fam, decltype((void)Fam::kHasIntFixed)> { static constexpr bool value = Fam::kHasIntFixed; };

template <class Fam, class Sync>
BPT_D void explore_chain(const DevEngine& E, Fam& fam, GroupReducer& red, double* P, int* order, int& curpos, int S,
                         int pos, double beta, uint32_t scan, uint32_t key1, int mode, unsigned long long& evals,
                         unsigned long long& execs, int& err) {
  const int tid = Sync::tid();
  if (mode == MODE_FORWARD_INIT) {
    Rng rng(E.key0, key1, P_INIT, (uint32_t)pos, 0u, 0u);
    fam.forward(rng);
  } else if (mode == MODE_EXPLORE) {
    if (beta == 0.0 && E.use_prior) {
      Rng rng(E.key0, key1, P_FORWARD, (uint32_t)pos, scan, 0u);
      fam.forward(rng);
    } else {
      // A NaN coordinate (only possible through set_state) makes every factor NaN: the reference throws at the first
      // evaluation (ExponentiatedFactor.java:26-29); here the slice sampler's shrinkage would never terminate.
      bool nan_state = false;
      for (int v = 0; v < Fam::n_reals(fam.d); v++) nan_state = nan_state || (P[v] != P[v]);
      int n_steps = E.steps_override > 0 ? E.steps_override : (int)floor(E.n_passes * S);
      if (nan_state) { err |= ERR_NAN; n_steps = 0; }
      for (int t = 0; t < n_steps; t++) {
        Sync::sync();
        if (curpos == -1) {   // Collections.shuffle(order, rnd); SampledModel.java:267-272
          if (tid == 0) {
            Rng sh(E.key0, key1, P_SHUFFLE, (uint32_t)pos, scan, (uint32_t)t);
            for (int i = S; i > 1; i--) {
              const int j = sh.next_int(i);
              const int tmp = order[i - 1]; order[i - 1] = order[j]; order[j] = tmp;
            }
          }
          Sync::sync();
          curpos = S - 1;
        }
        const int k = order[curpos--];
        fam.begin_execution();
        Rng rng(E.key0, key1, P_MOVE, (uint32_t)pos, scan, (uint32_t)t);
        int aux = 0;
        typename Sync::template LogDensityOf<Fam> lp{{E, fam, red, k, aux, beta, evals, err}};
        execs++;
        if (fam.sampler_type(k) == S_REAL_SLICE) {
          const double x0 = P[fam.sampler_var(k)];
          const double xn = real_slice(lp, rng, x0, false, 0.0, 0.0, err);
          fam.commit(k, xn, 0);
        } else if (HasIntSlice<Fam>::value && fam.sampler_type(k) == S_INT_SLICE) {
          // IntSliceSampler on a latent IntVar (stepping out, w = 10); the state is held as an exact double
          if constexpr (HasIntSlice<Fam>::value) {
            const int x0 = (int)P[fam.sampler_var(k)];
            const int xn = int_slice(lp, rng, x0, false, 0, 0, err);
            fam.commit(k, (double)xn, 0);
          }
        } else if (HasIntFixed<Fam>::value && fam.sampler_type(k) == S_INT_FIXED) {
          // CategoricalSampler.xtend:24-28 / UniformSampler.xtend:21-27: IntSliceSampler with a fixed window
          if constexpr (HasIntFixed<Fam>::value) {
            const int x0 = (int)P[fam.sampler_var(k)];
            const int xn = int_slice(lp, rng, x0, true, fam.int_lo(k), fam.int_hi(k), err);
            fam.commit(k, (double)xn, 0);
          }
        } else {   // SimplexSampler.xtend:21-28
          const int K = fam.simplex_dim();
          aux = rng.next_int(K);
          lp.aux = aux;
          const int base = fam.sampler_var(k);
          const int nx = aux == K - 1 ? 0 : aux + 1;
          const double x0 = P[base + aux];
          const double sum = P[base + aux] + P[base + nx];
          const double xn = real_slice(lp, rng, x0, true, 0.0, sum, err);
          fam.commit(k, xn, aux);
        }
      }
    }
  }
  Sync::sync();
}

// families may state how many CTAs must fit one SM (kMinBlocks): it bounds the registers per thread
template <class Fam, class = void> struct MinBlocksOf { static constexpr int value = 1; };
template <class Fam> struct MinBlocksOf<Fam, decltype((void)Fam::kMinBlocks)> { static constexpr int value = Fam::kMinBlocks; };

// NaN in the from-scratch sums (SampledModel.updateAll): an error (ExponentiatedFactor.java:26-29), or, under
// treatNaNAsNegativeInfinity, a factor equal to -inf: a fixed one makes the fixed sum -inf, an annealed one is out of
// support (families that sum their annealed factors in closed form cannot say how many were NaN: counted once)
BPT_D void nan_policy(const DevEngine& E, double& loglik, int& noos, double& logprior, int& err) {
  if (loglik == loglik && logprior == logprior) return;
  if (!E.nan_neginf) { err |= ERR_NAN; return; }
  if (logprior != logprior) logprior = BPT_NEG_INF;
  if (loglik != loglik) { loglik = 0.0; noos += 1; }
}

template <class Fam>
__global__ void __launch_bounds__(Fam::kThreads, MinBlocksOf<Fam>::value)
explore_kernel(DevEngine E, typename Fam::Data D, int G, int mode) {
  __shared__ ReduceSmem rs;
  __shared__ double P[kMaxParams + 4 * 8];
  __shared__ int order[kMaxSamplers];

  const int tid = threadIdx.x;
  const int l = blockIdx.x / G;
  const int crank = blockIdx.x % G;
  const int g = E.slot_lo + l;
  const int rep = g / E.N, s_in_rep = g - rep * E.N;
  const int pos = E.pos_of_slot[rep * E.N + s_in_rep];
  const double beta = E.beta_override == E.beta_override ? E.beta_override : E.ladder[rep * E.N + pos];
  const uint32_t scan = E.scan_counter[rep];
  const uint32_t key1 = E.key1 ^ (E.replica_offset + (uint32_t)rep);

  Fam fam;
  fam.init(D, l, P, G, crank, E.err);
  const int nR = Fam::n_reals(D), S = Fam::n_samplers(D);
  for (int v = tid; v < nR; v += blockDim.x) P[v] = E.reals[(size_t)v * E.M_local + l];
  for (int k = tid; k < S; k += blockDim.x) order[k] = E.order[(size_t)l * E.s_max + k];
  int curpos = E.curpos[l];
  Fam::tables_init();
  __syncthreads();

  GroupReducer red;
  red.init(&rs, G, crank);
  unsigned long long evals = 0, execs = 0;
  int err = 0;
  explore_chain<Fam, CtaSync>(E, fam, red, P, order, curpos, S, pos, beta, scan, key1, mode, evals, execs, err);

  double loglik, logprior; int noos;
  fam.full(red, loglik, noos, logprior);
  nan_policy(E, loglik, noos, logprior, err);
  __syncthreads();
  if (crank == 0) {
    for (int v = tid; v < nR; v += blockDim.x) E.reals[(size_t)v * E.M_local + l] = P[v];
    for (int k = tid; k < S; k += blockDim.x) E.order[(size_t)l * E.s_max + k] = order[k];
    if (tid == 0) {
      double* t = E.table + (size_t)g * 4;
      t[0] = loglik; t[1] = (double)noos; t[2] = logprior; t[3] = (double)evals;
      double* sd = E.send + (size_t)l * 4;
      sd[0] = loglik; sd[1] = (double)noos; sd[2] = logprior; sd[3] = (double)evals;
      E.curpos[l] = curpos;
      E.evals[l] += evals;
      E.execs[l] += execs;
      if (err) atomicOr(E.err, err);
      peer_publish(E, g, pos, loglik, noos, logprior, evals, P, nR);
    }
  }
}

// commons-math3 SummaryStatistics first/second moment recurrences
BPT_D void summary_add(long long& n, double& m1, double& m2, double x) {
  if (n == 0) { m1 = 0.0; m2 = 0.0; }
  n++;
  const double dev = x - m1, ndev = dev / (double)n;
  m1 += ndev;
  m2 += ((double)n - 1.0) * dev * ndev;
}

// energy accumulators + DEO swap pass + index process of one replica; called by every thread of a CTA
BPT_D void swap_replica(const DevEngine& E, const DevStats& St, const DevOut& O, int rep) {
  const int N = E.N, tid = threadIdx.x;
  const size_t base = (size_t)rep * N;
  const double* ladder = E.ladder + base;
  int* sop = E.slot_of_pos + base;
  int* pos_of = E.pos_of_slot + base;
  const uint32_t scan = E.scan_counter[rep];
  const uint32_t swap_index = E.swap_counter[rep];
  const uint32_t b = E.batch_counter[rep];
  const uint32_t key1 = E.key1 ^ (E.replica_offset + (uint32_t)rep);
  const size_t obase = ((size_t)b * E.R + rep) * N;

  // ---- A: accumulators of moveKernel :88-90 and the per-scan records (PT.java:116-143,360-372)
  unsigned long long ev = 0;
  for (int p = tid; p < N; p += blockDim.x) {
    const int s = sop[p];
    const size_t g = base + s;
    const double ll = E.table[g * 4 + 0], no = E.table[g * 4 + 1], fx = E.table[g * 4 + 2];
    ev += (unsigned long long)E.table[g * 4 + 3];
    if (no > 0 && E.anneal_support) atomicOr(&St.oos_seen[rep], 1);   // SampledModel.java:361-362,396-398
    const double after = -ll, before = -E.prev_loglik[g];
    E.prev_loglik[g] = ll;
    summary_add(St.en_n[base + p], St.en_m1[base + p], St.en_m2[base + p], after);
    {   // CovarAccumulator.add :7-13
      int n = St.cov_n[base + p] + 1;
      double mx = St.cov_mx[base + p], my = St.cov_my[base + p];
      const double dx = before - mx;
      mx += dx / n; my += (after - my) / n;
      St.cov_c[base + p] += dx * (after - my);
      St.cov_mx[base + p] = mx; St.cov_my[base + p] = my; St.cov_n[base + p] = n;
    }
    if (O.energy) O.energy[obase + p] = after;
    if (O.logdens) O.logdens[obase + p] = annealed_log_density(fx, ll, no, ladder[p], E.anneal_support);
    if (O.noos) O.noos[obase + p] = (int)no;
    if (O.indicators) O.indicators[obase + p] = 0;
    if (p == 0 && O.samples) {
      const int l = (int)g - E.slot_lo;
      if (E.peers.world) {       // multi-GPU: the position-0 reals every rank received with the exchange
        for (int v = 0; v < O.n_reals; v++)
          O.samples[((size_t)b * E.R + rep) * O.n_reals + v] = E.peers.sample_buf[(size_t)rep * O.n_reals + v];
      } else if (l >= 0 && l < E.M_local)
        for (int v = 0; v < O.n_reals; v++)
          O.samples[((size_t)b * E.R + rep) * O.n_reals + v] = E.reals[(size_t)v * E.M_local + l];
    }
  }
  if (ev) atomicAdd(St.total_evals, ev);
  __syncthreads();

  // ---- B: swapKernel :64-75
  if (N > 1) {
    int offset;
    if (O.reversible) { Rng r(E.key0, key1, P_SWAPDIR, 0u, scan, 0u); offset = r.next_int(2); }
    else offset = (int)(swap_index & 1u);
    const int n_pairs = (N - offset) / 2;
    for (int q = tid; q < n_pairs; q += blockDim.x) {
      const int i = offset + 2 * q, j = i + 1;
      const int si = sop[i], sj = sop[j];
      const double* ti = E.table + (base + si) * 4;
      const double* tj = E.table + (base + sj) * 4;
      const double bi = ladder[i], bj = ladder[j];
      // swapAcceptPr :107-124
      const double ss = annealed_log_density(tj[2], tj[0], tj[1], bi, E.anneal_support) -
                        annealed_log_density(tj[2], tj[0], tj[1], bj, E.anneal_support);
      St.ls_sum[base + i] = log_add(St.ls_sum[base + i], ss);
      St.ls_n[base + i] += 1;
      const double log_ratio = ss + annealed_log_density(ti[2], ti[0], ti[1], bj, E.anneal_support) -
                               annealed_log_density(ti[2], ti[0], ti[1], bi, E.anneal_support);
      double a = fmin(1.0, exp(log_ratio));
      if (log_ratio != log_ratio || a != a) a = 0.0;
      Rng r(E.key0, key1, P_SWAP, (uint32_t)i, scan, 0u);
      if (r.bernoulli(a)) {   // doSwap :155-164
        sop[i] = sj; sop[j] = si;
        pos_of[sj] = i; pos_of[si] = j;
        if (O.indicators) O.indicators[obase + i] = 1;
      }
      double m2_unused = 0.0;
      summary_add(St.acc_n[base + i], St.acc_m1[base + i], m2_unused, a);
    }
    __syncthreads();
    // ---- C: index process, Paths.xtend:40-50 (particle identity = slot)
    unsigned long long rj = 0;
    for (int p = tid; p < N; p += blockDim.x) {
      const int s = sop[p];
      int f = St.path_flag[base + s];
      if (p == 0 && f) { f = 0; rj++; }
      else if (p == N - 1) f = 1;
      St.path_flag[base + s] = f;
    }
    if (rj) atomicAdd(&St.rejuv[rep], rj);
  }
  __syncthreads();
  if (tid == 0) {
    E.scan_counter[rep] = scan + 1;
    if (N > 1) E.swap_counter[rep] = swap_index + 1;
    E.batch_counter[rep] = b + 1;
  }
}

static __global__ void swap_kernel(DevEngine E, DevStats St, DevOut O) { swap_replica(E, St, O, blockIdx.x); }

// Fused kernel for many small models (config C5: thousands of independent replicas, a few chains each):
// one CTA per replica, one WARP per chain (reductions are shuffles only), and the whole scan loop --
// explore, record, DEO swap, index process -- runs inside the kernel, so n_scans scans cost one launch and
// the swap pass costs two block barriers instead of two kernel launches.
// per-chain parameter / sampler-order slots of the fused kernel: 8 unless the family states kFusedParams
template <class Fam, class = void> struct FusedParamsOf { static constexpr int value = 8; };
template <class Fam> struct FusedParamsOf<Fam, decltype((void)Fam::kFusedParams)> { static constexpr int value = Fam::kFusedParams; };

template <class Fam>
__global__ void __launch_bounds__(1024) replica_kernel(DevEngine E, DevStats St, DevOut O, typename Fam::Data D, int n_scans) {
  __shared__ double Pall[32][FusedParamsOf<Fam>::value];
  __shared__ int order_all[32][FusedParamsOf<Fam>::value];
  const int rep = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = rep * E.N + warp;              // local slot == global slot (single-process path)
  double* P = Pall[warp];
  int* order = order_all[warp];
  const int nR = Fam::n_reals(D), S = Fam::n_samplers(D);
  const uint32_t key1 = E.key1 ^ (E.replica_offset + (uint32_t)rep);
  Fam fam;
  fam.init(D, l, P, 1, 0, E.err);
  fam.set_lanes(lane, 32);
  for (int v = lane; v < nR; v += 32) P[v] = E.reals[(size_t)v * E.M_local + l];
  for (int k = lane; k < S; k += 32) order[k] = E.order[(size_t)l * E.s_max + k];
  int curpos = E.curpos[l];
  __syncwarp();
  GroupReducer red;
  red.init_warp();
  unsigned long long evals_total = 0, execs = 0;
  int err = 0;
  for (int it = 0; it < n_scans; it++) {
    const int pos = E.pos_of_slot[l];
    const double beta = E.ladder[rep * E.N + pos];
    const uint32_t scan = E.scan_counter[rep];
    unsigned long long evals = 0;
    explore_chain<Fam, WarpSync>(E, fam, red, P, order, curpos, S, pos, beta, scan, key1, MODE_EXPLORE, evals, execs, err);
    double loglik, logprior; int noos;
    fam.full(red, loglik, noos, logprior);
    nan_policy(E, loglik, noos, logprior, err);
    for (int v = lane; v < nR; v += 32) E.reals[(size_t)v * E.M_local + l] = P[v];
    if (lane == 0) {
      double* t = E.table + (size_t)l * 4;
      t[0] = loglik; t[1] = (double)noos; t[2] = logprior; t[3] = (double)evals;
    }
    evals_total += evals;
    __syncthreads();
    swap_replica(E, St, O, rep);
    __syncthreads();
  }
  for (int k = lane; k < S; k += 32) E.order[(size_t)l * E.s_max + k] = order[k];
  if (lane == 0) {
    E.curpos[l] = curpos;
    E.evals[l] += evals_total;
    E.execs[l] += execs;
    if (err) atomicOr(E.err, err);
  }
}

// energy "before" of the next scan = current table (called after an evaluation-only pass)
static __global__ void prime_prev_kernel(DevEngine E, DevStats St) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E.R * E.N) {
    E.prev_loglik[i] = E.table[(size_t)i * 4];
    if (E.table[(size_t)i * 4 + 1] > 0 && E.anneal_support) atomicOr(&St.oos_seen[i / E.N], 1);
  }
}

// ParallelTempering.setAnnealingParameters :197-204 + Paths.initPaths :131-139
static __global__ void reset_round_kernel(DevEngine E, DevStats St) {
  const int rep = blockIdx.x, N = E.N;
  const size_t base = (size_t)rep * N;
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    St.en_n[base + p] = 0; St.en_m1[base + p] = 0.0; St.en_m2[base + p] = 0.0;
    St.cov_mx[base + p] = 0.0; St.cov_my[base + p] = 0.0; St.cov_c[base + p] = 0.0; St.cov_n[base + p] = 0;
    St.acc_n[base + p] = 0; St.acc_m1[base + p] = 0.0;
    St.ls_sum[base + p] = BPT_NEG_INF; St.ls_n[base + p] = 0;
    // initial visit of the particle sitting at position p
    const int s = E.slot_of_pos[base + p];
    St.path_flag[base + s] = (p == N - 1) ? 1 : 0;
  }
  if (threadIdx.x == 0) St.rejuv[rep] = 0;
}

}  // namespace bpt
