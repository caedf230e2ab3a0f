// pool.cuh -- the "pool" exploration kernel: every SM serves every chain.
//
// explore_kernel (kernels.cuh) gives each chain its own thread-block cluster.  For the big-data families that
// wastes the machine in three ways (ncu, round 1): clusters larger than two CTAs do not pack into the GPCs, so
// 64 chains x 2 CTAs use 128 of 148 SMs; every launch waits for its slowest chain (cold chains need ~25 % more
// evaluations per scan than hot ones); and every evaluation pays its reduction, cluster barrier and scalar
// slice-sampler logic while its SMs idle.  Here the roles are split instead:
//
//   * one persistent CTA per SM (cooperative launch).  Its WORKER warps own a fixed slab of the observations
//     (rows [lo_s, hi_s), multiples of 64 rows) for ALL chains;
//   * a chain's sampler (ParallelTempering.moveKernel -> SampledModel.posteriorSamplingScan ->
//     RealSliceSampler.execute, the same explore_chain()/slice.cuh code as explore_kernel) runs on one CONTROLLER
//     warp (all lanes redundantly, like the threads of a cluster in explore_kernel).  Each log-density evaluation
//     is posted as one 16-byte request record in global memory (L2);
//   * every SM picks the request up (poll, claim in shared memory), one warp evaluates the SM's slab of that chain
//     and publishes {partial sum, sequence number, out-of-support count} as one 16-byte record;
//   * the controller warp polls the chain's n_SM records (five per lane) and adds them in a fixed order.
//
// All chains' requests interleave on all SMs, so the load is balanced whatever the per-chain evaluation counts
// are, all SMs are used for any number of chains, and one chain's fixed costs (L2 round trips, reduction, scalar
// logic) overlap other chains' row work.  Sums are fixed-order (lane-sequential, xor-butterfly, ascending SM
// index): results are bit-reproducible and do not depend on which warp served a request.
//
// Accepted moves are not applied to the per-chain cache eta in a separate pass: the commit (eta += dprev * X[kprev])
// rides on the chain's next request, whose rows are read anyway.
//
// Spin waits exist only inside this one cooperative launch (co-residency is guaranteed) and every one of them
// gives up after kPoolTimeoutCycles through a global abort flag, which surfaces as BLANGPT_ECUDA, never as a hang.
#pragma once
#include "kernels.cuh"

namespace bpt {

#ifndef BLANGPT_POOL_BELL
#define BLANGPT_POOL_BELL 0          // idle workers sleep on an mbarrier phase (1) or poll shared memory with nanosleep (0)
#endif
#ifndef BLANGPT_POOL_THREADS
#define BLANGPT_POOL_THREADS 768
#endif
constexpr int kPoolThreads = BLANGPT_POOL_THREADS;   // 24 warps per CTA: controller warps (one per chain), one dispatcher, the rest workers
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kPoolMaxChains = 1024;      // chains per GPU served by the pool (size of the per-CTA request ring; power of two)
constexpr int kPoolMaxSlotsPerLane = 8;   // n_SM <= 256
constexpr long long kPoolTimeoutCycles = 12LL * 1000 * 1000 * 1000;   // ~6 s at 1.9 GHz

enum PoolReqType : int { PR_EVAL = 1, PR_FULL = 2 };
enum { PR_HAS_COMMIT = 1 << 12 };

// Everything that crosses SMs is a 16-byte record {double, sequence number, int} written and read with ONE
// aligned 16-byte access (a single transaction), so a reader that sees the sequence number it waits for also
// holds the payload: no separate flag, no counter, no atomics.
struct __align__(16) PoolHalf { double v; uint32_t seq; int i; };
// request of chain c: reqa[c] = {delta, seq, k | type << 8 | PR_HAS_COMMIT}, reqb[c] = {dprev, seq, kprev} (read only with
// PR_HAS_COMMIT).  The a records of all chains are contiguous: one CTA's dispatcher warp polls them all with a few loads.

struct PoolCtl {
  int C, n_cta, n_ctrl;  // local chains, WORKER CTAs (row slabs), controller CTAs (32 chains each); grid = n_cta + n_ctrl
  PoolHalf* reqa;        // [C]
  PoolHalf* reqb;        // [C]
  PoolHalf* slot;        // [C][n_cta] partial results {sum, seq, out-of-support count}
  uint32_t* finished;    // [1]   chains whose scan is complete
  int* abort_flag;       // [1]
  double* wbuf;          // [C][kMaxParams]  coefficients for PR_FULL
  unsigned long long* prof;  // [n_cta * kPoolWarps][8] cycle counters (null unless BLANGPT_POOL_PROFILE is set)
};

// ---- memory-model primitives (gpu scope) ----
BPT_D uint32_t pool_ld_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
BPT_D int pool_ld_s32(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
BPT_D double pool_ld_f64(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
BPT_D void pool_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
BPT_D PoolHalf pool_ld_half(const PoolHalf* p) {
  uint32_t a, b, c, d;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  PoolHalf h; h.v = __hiloint2double((int)b, (int)a); h.seq = c; h.i = (int)d;
  return h;
}
BPT_D void pool_st_half(PoolHalf* p, double v, uint32_t seq, int i) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"((uint32_t)__double2loint(v)),
               "r"((uint32_t)__double2hiint(v)), "r"(seq), "r"((uint32_t)i) : "memory");
}

// The doorbell (BLANGPT_POOL_BELL=1, off by default): idle worker warps sleep in hardware on an mbarrier phase instead
// of polling shared memory with nanosleep; the dispatcher arrives once per publication, a waiter names the phase by
// its parity.  ncu counts the polls as a sixth of all issued instructions, but the workers are busy ~90 % of the time
// and the polls mostly fill issue slots nobody wants: measured 3.975 M evals/s with the bell, 3.990 M without
// (profiles/ab_pool_r02.md), so polling stays the default.
BPT_D void bell_init(unsigned long long* b) {
  if (!BLANGPT_POOL_BELL) return;
  asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
BPT_D void bell_ring(unsigned long long* b) {
  if (!BLANGPT_POOL_BELL) return;
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared.b64 st, [%0];\n\t}" :: "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
// sleeps until phase `parity` of the bell has completed, or for the hardware's time limit (then false)
BPT_D bool bell_wait(unsigned long long* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
  return ok != 0;
}

// Requests reach the worker warps of a CTA through a ring in shared memory filled by the CTA's dispatcher warp, the
// only warp of the SM that polls global memory for requests (every warp polling saturated the L2 slices that hold
// the request records: 75 % of the L2's throughput went to polls).  entry.i = k | type << 8 | commit << 12 | chain << 16.
struct PoolQueue {
  unsigned int head;                              // next ticket (workers, shared-memory atomic)
  unsigned int tail;                              // entries published (dispatcher)
  int done;                                       // every chain has finished
  int lo, hi;                                     // this SM's rows (PF::slab), the same for every request
  unsigned int pubs;                              // publications so far (dispatcher); phase pubs of `bell` is the one in progress
  unsigned long long bell;                        // mbarrier (count 1): the dispatcher arrives once per publication
  PoolHalf entry[kPoolMaxChains];
};
struct PoolSmem {
  union {
    struct {                                        // worker CTAs
      PoolQueue q;
      uint32_t seen[kPoolMaxChains];                // dispatcher: newest sequence number queued, per chain
    };
    struct {                                        // controller CTAs, one chain per warp
      double P[kPoolWarps][kMaxParams + 4 * 8];     // chain parameters
      int order[kPoolWarps][kMaxSamplers];          // SampledModel.currentSamplingOrder
    };
  };
};

// Bank-conflict-free copies of the two lookup tables of logit_fast() (dynamic shared memory, 40 KB).  Lanes index
// the tables with unrelated j, so the compact tables cost 5-7 shared-memory wavefronts per lookup and the LSU data
// pipe, not the fp64 pipe, limited the kernel (ncu: l1tex data-pipe wavefronts 96 % of peak).  Here entry j is
// replicated across a 128-byte row, and a lane always reads its own column of the row:
//   log table: 256 rows x 8 copies of {1/c_j, -log(1/c_j)} (16 B): lanes l, l+8, l+16, l+24 share a column -> 4
//              wavefronts, the minimum for 32 x 16 B;
//   exp table:  64 rows x 16 copies of 2^(j/64) (8 B): lanes l, l+16 share a column -> 2 wavefronts, the minimum.
constexpr uint32_t kPoolLogTabBytes = 256u * 128u, kPoolExpTabBytes = 64u * 128u;
constexpr size_t kPoolDynSmemBytes = kPoolLogTabBytes + kPoolExpTabBytes;
BPT_D void pool_tables_init(unsigned char* dyn) {
  double2* lg = reinterpret_cast<double2*>(dyn);
  double* ex = reinterpret_cast<double*>(dyn + kPoolLogTabBytes);
  for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) lg[i] = kLgtLogTabC[i >> 3];
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) ex[i] = kLgtExpTabC[i >> 4];
}
// logit_fast<2>() on the replicated tables.  tl_b = shared-memory address of this lane's column of the log table minus
// 0x3ff00 rows (the biased exponent of u in (1, 2] is folded into the base), te_base = this lane's column of the exp
// table.  Index arithmetic as shift + multiply-add pairs (two integer instructions per lookup); results are
// bit-identical to logit_fast<2>().
// Returns logit_is_fast_bits(z0) && logit_is_fast_bits(z1), evaluated on words the arithmetic produces anyway:
// z > -12 on the high word of z - |z| (= 2 min(z, 0)), |z| < 700 (and not NaN) on the high word of z.
BPT_D bool logit_fast2_rep(double z0, double z1, double& s, uint32_t tl, uint32_t te_base) {
  const double z[2] = {z0, z1};
  double kd[2], r[2], p[2], te[2]; int m[2];
#pragma unroll
  for (int w = 0; w < 2; w++) {
    const double a = fabs(z[w]);
    kd[w] = fma(a, kLgtK[0], kLgtK[1]);     // magic add: low word = round(-a * 64 log2 e)
    m[w] = __double2loint(kd[w]);
    kd[w] -= kLgtK[1];
    uint32_t ea;
    asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 63;\n\tmad.lo.u32 %0, t, 128, %2;\n\t}" : "=r"(ea) : "r"(m[w]), "r"(te_base));
    asm("ld.shared.f64 %0, [%1];" : "=d"(te[w]) : "r"(ea));
    r[w] = fma(a, kLgtK[0], -kd[w]);        // f = -a * 64 log2 e - m, one rounding
    p[w] = kLgtExp2C[0];
  }
#pragma unroll
  for (int i = 1; i < 6; i++)
#pragma unroll
    for (int w = 0; w < 2; w++) p[w] = fma(p[w], r[w], kLgtExp2C[i]);
  double q[2], rr[2], lc[2];
#pragma unroll
  for (int w = 0; w < 2; w++) {
    const double tp = te[w] * p[w];
    int thi;                                  // high word + ((m >> 6) << 20): scale by 2^(m >> 6)
    asm("{\n\t.reg .s32 t;\n\tshr.s32 t, %1, 6;\n\tmad.lo.s32 %0, t, 1048576, %2;\n\t}" : "=r"(thi) : "r"(m[w]), "r"(__double2hiint(tp)));
    const double t = __hiloint2double(thi, __double2loint(tp));
    const double u = 1.0 + t;
    const unsigned j = min(255u, (unsigned)(__double2hiint(u) - 0x3ff00000) >> 12);
    double2 cj;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cj.x), "=d"(cj.y) : "r"(tl + (j << 7)));
    rr[w] = fma(u, cj.x, -1.0);
    lc[w] = cj.y;
    q[w] = kLgtLogC[0];
  }
#pragma unroll
  for (int i = 1; i < 4; i++)
#pragma unroll
    for (int w = 0; w < 2; w++) q[w] = fma(q[w], rr[w], kLgtLogC[i]);
  bool ok = true;
#pragma unroll
  for (int w = 0; w < 2; w++) {
    const double l = fma(rr[w], fma(rr[w], q[w], 1.0), lc[w]);    // log1p(exp(-|z|)) = log c_j + rr (1 + rr q)
    const double wz = z[w] - fabs(z[w]);                           // 2 min(z, 0)
    ok = ok && (unsigned)__double2hiint(wz) < 0xC0380000u && ((unsigned)__double2hiint(z[w]) & 0x7fffffffu) < 0x4085E000u;
    s += fma(0.5, wz, -l);                                         // min(z,0) - log1p(exp(-|z|))
  }
  return ok;
}

// integer form of logit_is_fast(): zs > -12 (to within one high-word step) and |zs| < 700; NaN and inf fail
BPT_D bool logit_is_fast_bits(double zs) {
  const int hi = __double2hiint(zs);
  return ((unsigned)hi & 0x7fffffffu) < (hi < 0 ? 0x40280000u : 0x4085E000u);
}

// ===========================================================================
// C2 logistic regression on the pool
// ===========================================================================
struct LogisticPool {
  using Data = LogisticFam::Data;

  // rows of one SM: units of 64 rows, spread evenly
  BPT_D static void slab(const Data& D, int cta, int n_cta, int& lo, int& hi) {
    const long long U = (D.n + 63) >> 6;
    lo = (int)(U * cta / n_cta) << 6;
    hi = min(D.n, (int)(U * (cta + 1) / n_cta) << 6);
  }

  // the rare literal path of two rows (some lane of the warp has a row outside the lean range): out of line, so
  // that the hot loop stays straight-line code; results by value, so that the caller's accumulators stay in registers
  struct SlowPair { double s; int c; };
  __device__ __noinline__ static SlowPair duo_slow(double s0, double za, double zb, const uint8_t* yp, int* err, uint32_t tab) {
    SlowPair r; r.s = s0; r.c = 0;
    if (logit_is_fast_bits(za) && logit_is_fast_bits(zb)) { const double z[2] = {za, zb}; logit_fast<2>(z, r.s, tab); }   // as the hot loop
    else { logit_term(za, yp, r.s, r.c, err); logit_term(zb, yp + 1, r.s, r.c, err); }
    return r;
  }
  template <bool FULL_WARP>
  BPT_D static void duo(double za, double zb, const uint8_t* yp, double& s, int& c, int* err, uint32_t tab) {
    const bool ok = logit_is_fast_bits(za) && logit_is_fast_bits(zb);
    const double z[2] = {za, zb};
    double t = s;
    logit_fast<2>(z, t, tab);
    if (__all_sync(FULL_WARP ? 0xffffffffu : __activemask(), ok)) { s = t; return; }
    if (ok) s = t;
    else { const SlowPair r = duo_slow(s, za, zb, yp, err, tab); s = r.s; c += r.c; }
  }

  // PR_EVAL on rows [lo, hi) of chain c: sum_i logit_term(eta_i + delta * X[i][k]); COMMIT first applies
  // eta_i += dprev * X[i][kprev] and stores it.  Two rows per lane per step (64 rows per warp step), the next
  // step's 16-byte loads issued before the current step's arithmetic; eight warps per scheduler hide the rest.
  // Lane-sequential sums; the caller reduces over lanes.
  template <bool COMMIT>
  BPT_D static void eval_rows(const Data& D, int c, int lo, int hi, int k, double delta, int kprev, double dprev, int lane,
                              uint32_t tab, int* err, double& acc, int& cnt) {
    double* eta = D.eta + (size_t)c * D.npad;
    const double* colk = D.xt + (size_t)k * D.npad;
    const double* colp = D.xt + (size_t)kprev * D.npad;
    const int n_pairs = (hi - lo) >> 1;
    double2* pe = reinterpret_cast<double2*>(eta + lo) + lane;
    const double2* px = reinterpret_cast<const double2*>(colk + lo) + lane;
    const double2* pp = reinterpret_cast<const double2*>(colp + lo) + lane;
    const uint8_t* py = D.y + lo + 2 * lane;
    // loads as volatile asm: issued where they are written (the compiler otherwise sinks the prefetch to the end of
    // the step to save registers, which exposes the whole L2 latency every step)
    extern __shared__ __align__(128) unsigned char pool_dyn[];
    uint32_t tl = (uint32_t)__cvta_generic_to_shared(pool_dyn) + ((uint32_t)lane & 7u) * 16u;
    uint32_t te = (uint32_t)__cvta_generic_to_shared(pool_dyn) + kPoolLogTabBytes + ((uint32_t)lane & 15u) * 8u;
    asm volatile("" : "+r"(tl), "+r"(te));   // opaque: kept in registers instead of being re-derived per use
    auto ldg2 = [](const double2* p) { return *p; };
    int j = 0;
    if (n_pairs >= 32) {
      double2 e = ldg2(pe), x = ldg2(px), q;
      if (COMMIT) q = ldg2(pp);
      // Blocks of up to 32 steps.  The inner loop is straight-line code with no branch but its own back edge and no
      // call (so its constants stay in registers): a step whose range vote fails keeps the accumulator unchanged and
      // sets its bit in `bad`; those steps (0.4 % on the bench) are redone out of line after the block.  The sum is
      // therefore "lean steps in order, then literal steps in order" per block: a fixed order.
      while (j + 32 <= n_pairs) {
        const int j0 = j;
        unsigned bad = 0u, bit = 1u;
        const int j_end = min(n_pairs - 31, j0 + 32 * 32);
        for (; j < j_end; j += 32, bit <<= 1) {
          if (COMMIT) { e.x = fma(dprev, q.x, e.x); e.y = fma(dprev, q.y, e.y); pe[j] = e; }
          const double za = fma(delta, x.x, e.x), zb = fma(delta, x.y, e.y);
          // e, x (and q) are dead from here on: the next step's rows are loaded into the same registers right away
          // and have the whole step's arithmetic to arrive
          if (j + 64 <= n_pairs) { e = ldg2(pe + j + 32); x = ldg2(px + j + 32); if (COMMIT) q = ldg2(pp + j + 32); }
          double t = acc;
          const bool ok = logit_fast2_rep(za, zb, t, tl, te);
          const bool all = __all_sync(0xffffffffu, ok);
          acc = all ? t : acc;
          bad |= all ? 0u : bit;
        }
        while (bad) {
          const int jb = j0 + 32 * (__ffs(bad) - 1);
          bad &= bad - 1u;
          const double2 eb = pe[jb], xb = px[jb];       // eta already holds the committed value
          const SlowPair r = duo_slow(acc, fma(delta, xb.x, eb.x), fma(delta, xb.y, eb.y), py + 2 * jb, err, tab);
          acc = r.s; cnt += r.c;
        }
      }
    }
    if (j + lane < n_pairs) {     // ragged rest (last SM only)
      double2 e = pe[j];
      const double2 x = px[j];
      if (COMMIT) { const double2 q = pp[j]; e.x = fma(dprev, q.x, e.x); e.y = fma(dprev, q.y, e.y); pe[j] = e; }
      duo<false>(fma(delta, x.x, e.x), fma(delta, x.y, e.y), py + 2 * j, acc, cnt, err, tab);
    }
    if (((hi - lo) & 1) && lane == 0) {   // odd last row of the data set
      const int i = hi - 1;
      double e = eta[i];
      if (COMMIT) { e = fma(dprev, colp[i], e); eta[i] = e; }
      logit_term(fma(delta, colk[i], e), D.y + i, acc, cnt, err);
    }
  }

  // PR_FULL: rebuild eta from the coefficients (sum_j w_j X[i][j], j ascending from 0.0, F/SpikedGLM.bl:29-33)
  // and sum every row term from scratch (SampledModel.updateAll)
  BPT_D static void full_rows(const Data& D, int c, int lo, int hi, const double* w, int lane, int* err, double& acc, int& cnt) {
    double* eta = D.eta + (size_t)c * D.npad;
    for (int i = lo + 2 * lane; i < hi; i += 64) {
      const bool two = i + 1 < hi;
      double a = 0.0, b = 0.0;
      for (int j = 0; j < D.p; j++) {
        const double wj = w[j];
        const double* col = D.xt + (size_t)j * D.npad;
        if (two) { const double2 xv = *reinterpret_cast<const double2*>(col + i); a = fma(wj, xv.x, a); b = fma(wj, xv.y, b); }
        else a = fma(wj, col[i], a);
      }
      eta[i] = a;
      logit_term(a, D.y + i, acc, cnt, err);
      if (two) { eta[i + 1] = b; logit_term(b, D.y + i + 1, acc, cnt, err); }
    }
  }

  // a worker warp serves request (a, b) of chain c on this SM's slab; returns true when the request stored to eta
  BPT_D static bool serve(const Data& D, const PoolCtl& pc, int c, int lo, int hi, const PoolHalf& a, const PoolHalf& b, int lane,
                          uint32_t tab, int* err, double& s, int& cnt) {
    s = 0.0; cnt = 0;
    const int type = (a.i >> 8) & 0xf, k = a.i & 0xff;
    if (type == PR_EVAL) {
      if (a.i & PR_HAS_COMMIT) { eval_rows<true>(D, c, lo, hi, k, a.v, b.i, b.v, lane, tab, err, s, cnt); return true; }
      eval_rows<false>(D, c, lo, hi, k, a.v, 0, 0.0, lane, tab, err, s, cnt);
      return false;
    }
    full_rows(D, c, lo, hi, pc.wbuf + (size_t)c * kMaxParams, lane, err, s, cnt);   // wbuf: written once per launch
    return true;
  }

  // ---- the controller warp's view of the family (the Fam interface of explore_chain); every lane runs the
  //      sampler redundantly, lane 0 posts, all lanes collect the partial sums ----
  struct Ctl {
    using Data = LogisticFam::Data;
    Data d; double* P; PoolCtl pc; int l, lane; uint32_t seq; int pend_k; double pend_delta; bool aborted; int* lerr;
    unsigned long long cyc_wait;
    BPT_D static int n_reals(const Data& d) { return d.p; }
    BPT_D static int n_samplers(const Data& d) { return d.p; }
    BPT_D void init(const Data& d_, int local_slot, double* P_, const PoolCtl& pc_, int* lerr_) {
      d = d_; P = P_; pc = pc_; l = local_slot; lane = threadIdx.x & 31; seq = 0; pend_k = 0; pend_delta = 0.0;
      aborted = false; lerr = lerr_; cyc_wait = 0;
    }
    BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
    BPT_D int sampler_var(int k) const { return k; }
    BPT_D int simplex_dim() const { return 0; }
    BPT_D void begin_execution() {}
    BPT_D double fixed(int, double x, int) const { return normal_f2(0.0, d.v0, x); }

    BPT_D void post(int type, int k, double delta) {
      seq++;
      if (lane == 0) {
        int word = k | (type << 8);
        if (pend_delta != 0.0) { pool_st_half(pc.reqb + l, pend_delta, seq, pend_k); word |= PR_HAS_COMMIT; pool_fence(); }
        pool_st_half(pc.reqa + l, delta, seq, word);
      }
      pend_delta = 0.0;
    }
    // all lanes: wait until every SM's partial of request `seq` has arrived, then add them in a fixed order
    BPT_D bool wait(double& s, int& c) {
      const long long t0 = clock64();
      const PoolHalf* slots = pc.slot + (size_t)l * pc.n_cta;
      PoolHalf h[kPoolMaxSlotsPerLane];
      unsigned spins = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < kPoolMaxSlotsPerLane; j++) {
          const int i = lane + 32 * j;
          if (i < pc.n_cta) { h[j] = pool_ld_half(slots + i); ok = ok && h[j].seq == seq; }
        }
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(20);
        if ((++spins & 255u) == 0) {
          int stop = 0;
          if (lane == 0) {
            stop = pool_ld_s32(pc.abort_flag);
            if (!stop && clock64() - t0 > kPoolTimeoutCycles) { atomicExch(pc.abort_flag, 1); stop = 1; }
          }
          if (__shfl_sync(0xffffffffu, stop, 0)) { aborted = true; break; }
        }
      }
      if (aborted) { *lerr |= ERR_POOL_TIMEOUT; s = 0.0; c = 1; return false; }
      double t = 0.0; int tc = 0;
#pragma unroll
      for (int j = 0; j < kPoolMaxSlotsPerLane; j++)
        if (lane + 32 * j < pc.n_cta) { t += h[j].v; tc += h[j].i; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, o); tc += __shfl_xor_sync(0xffffffffu, tc, o); }
      s = t; c = tc;
      cyc_wait += (unsigned long long)(clock64() - t0);
      return true;
    }
    BPT_D void annealed(int k, double x, int, GroupReducer&, double& s, int& c) {
      if (aborted) { s = 0.0; c = 1; return; }
      post(PR_EVAL, k, x - P[k]);
      wait(s, c);
    }
    BPT_D void commit(int k, double x, int) {
      const double delta = x - P[k];
      if (delta != 0.0) {
        if (pend_delta != 0.0 && !aborted) { double s; int c; post(PR_EVAL, k, 0.0); wait(s, c); }   // never in practice: an evaluation precedes every commit
        pend_k = k; pend_delta = delta;
      }
      __syncwarp();   // every lane has read P[k]
      if (lane == 0) P[k] = x;
      __syncwarp();
    }
    BPT_D void refresh() {}
    BPT_D void full(GroupReducer&, double& loglik, int& noos, double& logprior) {
      __syncwarp();
      for (int j = lane; j < d.p; j += 32) pc.wbuf[(size_t)l * kMaxParams + j] = P[j];
      pool_fence();
      __syncwarp();
      pend_delta = 0.0;     // eta is rebuilt from the coefficients: a pending commit is already in them
      loglik = 0.0; noos = 1;
      if (!aborted) { post(PR_FULL, 0, 0.0); wait(loglik, noos); }
      double lp = 0.0;
      const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
      for (int j = 0; j < d.p; j++) lp += f01 + normal_f2(0.0, d.v0, P[j]);
      logprior = lp;
    }
    BPT_D void forward(Rng& rng) {
      __syncwarp();
      for (int j = 0; j < d.p; j++) { const double w = gen_normal(rng, 0.0, d.v0); if (lane == 0) P[j] = w; }
      __syncwarp();
    }
  };
};

// ===========================================================================
// the kernel
// ===========================================================================
// the CTA's dispatcher warp: polls the request records of all chains, queues the new ones for the workers
BPT_D void pool_dispatcher(const PoolCtl& pc, PoolSmem& sm, int lane) {
  const unsigned full = 0xffffffffu;
  const int C = pc.C;
  volatile unsigned int* tail_p = &sm.q.tail;
  unsigned int tail = 0, pubs = 0;
  const long long t0 = clock64();
  unsigned spins = 0;
  for (;;) {
    const bool all_done = pool_ld_u32(pc.finished) >= (uint32_t)C;
    bool any = false;
    for (int base = 0; base < C; base += 64) {
      const int c0 = base + lane, c1 = base + 32 + lane;
      PoolHalf m0, m1;
      m0.seq = 0; m0.v = 0.0; m0.i = 0; m1 = m0;
      if (c0 < C) m0 = pool_ld_half(pc.reqa + c0);
      if (c1 < C) m1 = pool_ld_half(pc.reqa + c1);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const PoolHalf& mine = half ? m1 : m0;
        const int c = half ? c1 : c0;
        const bool fresh = c < C && mine.seq != 0 && mine.seq != sm.seen[c];
        const unsigned m = __ballot_sync(full, fresh);
        if (m) {
          if (fresh) {
            sm.seen[c] = mine.seq;
            PoolHalf e = mine;
            e.i = (mine.i & 0xffff) | (c << 16);
            sm.q.entry[(tail + __popc(m & ((1u << lane) - 1u))) & (kPoolMaxChains - 1)] = e;
          }
          tail += __popc(m);
          any = true;
        }
      }
    }
    if (any) {
      __syncwarp();
      __threadfence_block();
      if (lane == 0) {
        *tail_p = tail;
        __threadfence_block();
        bell_ring(&sm.q.bell);                                   // completes phase `pubs`
        *(volatile unsigned int*)&sm.q.pubs = ++pubs;
      }
      continue;
    }
    if (all_done) { __syncwarp(); if (lane == 0) { *(volatile int*)&sm.q.done = 1; __threadfence_block(); bell_ring(&sm.q.bell); } return; }
    __nanosleep(20);
    if ((++spins & 1023u) == 0) {
      int stop = 0;
      if (lane == 0) {
        stop = pool_ld_s32(pc.abort_flag);
        if (!stop && clock64() - t0 > 40 * kPoolTimeoutCycles) { atomicExch(pc.abort_flag, 1); stop = 1; }   // whole-launch limit
      }
      if (__shfl_sync(full, stop, 0)) { if (lane == 0) { *(volatile int*)&sm.q.done = 1; __threadfence_block(); bell_ring(&sm.q.bell); } return; }
    }
  }
}

template <class PF>
BPT_D void pool_worker(const typename PF::Data& D, const PoolCtl& pc, PoolSmem& sm, int warp, int lane, int cta, int* err) {
  const unsigned full = 0xffffffffu;
  const uint32_t tab = logit_table_address();
#ifdef BLANGPT_POOL_PROF      // cycle counters (tools/build_variant.sh <name> -DBLANGPT_POOL_PROF, run with BLANGPT_POOL_PROFILE=1)
  const bool prof = pc.prof != nullptr;
#else
  constexpr bool prof = false;
#endif
  const long long t_enter = prof ? clock64() : 0;
  unsigned long long n_served = 0, cyc_serve = 0;
  unsigned n_poll = 0;
  volatile unsigned int* tail_p = &sm.q.tail;
  volatile int* done_p = &sm.q.done;
  for (;;) {
    // ---- take a ticket, wait until the dispatcher has published that entry ----
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&sm.q.head, 1u);
    ticket = __shfl_sync(full, ticket, 0);
    bool stop = false;
    for (;;) {
      // order: publication count, then tail -- a publication that this pass misses completes the phase waited for
      const unsigned int pubs = *(volatile unsigned int*)&sm.q.pubs;
      if ((int)(*tail_p - ticket) > 0) break;
      if (prof) n_poll++;
      if (*done_p) { stop = true; break; }
      if (BLANGPT_POOL_BELL) bell_wait(&sm.q.bell, pubs & 1u);
      else __nanosleep(150);
    }
    if (stop) {
      if (prof && lane == 0) {
        unsigned long long* q = pc.prof + ((size_t)blockIdx.x * kPoolWarps + warp) * 8;
        q[0] += n_served; q[1] += cyc_serve; q[2] += (unsigned long long)(clock64() - t_enter); q[3] += n_poll;
      }
      return;
    }
    const long long t_s0 = prof ? clock64() : 0;
    PoolHalf a = sm.q.entry[ticket & (kPoolMaxChains - 1)];
    const int hit_c = (int)((unsigned)a.i >> 16);
    // No acquire fence here: everything read below is reached through the request (address dependency) and was
    // made visible at gpu scope before the request was posted -- eta by the release fence of the warp of THIS SM
    // that last stored it (same L1), wbuf and the b record by the controller's fence.
    PoolHalf b;
    b.v = 0.0; b.seq = a.seq; b.i = 0;
    if (a.i & PR_HAS_COMMIT) b = pool_ld_half(pc.reqb + hit_c);   // written (and fenced) before a
    double s; int cnt;
    const bool stored = PF::serve(D, pc, hit_c, *(volatile int*)&sm.q.lo, *(volatile int*)&sm.q.hi, a, b, lane, tab, err, s, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
    if (__any_sync(full, cnt != 0)) cnt = __reduce_add_sync(full, cnt);   // out-of-support rows: rare
    if (stored) { pool_fence(); __syncwarp(); }   // release this request's eta stores before the result becomes visible
    if (lane == 0) pool_st_half(pc.slot + (size_t)hit_c * pc.n_cta + cta, s, a.seq, cnt);
    if (prof) { n_served++; cyc_serve += (unsigned long long)(clock64() - t_s0); }
  }
}

// one chain's scan on a controller warp: explore_kernel's per-chain prologue, explore_chain(), full(), epilogue
template <class PF>
BPT_D void pool_controller(const DevEngine& E, const typename PF::Data& D, const PoolCtl& pc, PoolSmem& sm, int warp, int lane,
                           int l, int mode) {
  const int g = E.slot_lo + l;
  const int rep = g / E.N, s_in_rep = g - rep * E.N;
  const int pos = E.pos_of_slot[rep * E.N + s_in_rep];
  const double beta = E.beta_override == E.beta_override ? E.beta_override : E.ladder[rep * E.N + pos];
  const uint32_t scan = E.scan_counter[rep];
  const uint32_t key1 = E.key1 ^ (E.replica_offset + (uint32_t)rep);
  double* P = sm.P[warp];
  int* order = sm.order[warp];
  const int nR = PF::Ctl::n_reals(D), S = PF::Ctl::n_samplers(D);
  for (int v = lane; v < nR; v += 32) P[v] = E.reals[(size_t)v * E.M_local + l];
  for (int k = lane; k < S; k += 32) order[k] = E.order[(size_t)l * E.s_max + k];
  int curpos = E.curpos[l];
  __syncwarp();
  int err = 0;
  const long long t_c0 = clock64();
  typename PF::Ctl fam;
  fam.init(D, l, P, pc, &err);
  GroupReducer red;
  red.init_warp();
  unsigned long long evals = 0, execs = 0;
  explore_chain<typename PF::Ctl, WarpSync>(E, fam, red, P, order, curpos, S, pos, beta, scan, key1, mode, evals, execs, err);
  double loglik, logprior; int noos;
  fam.full(red, loglik, noos, logprior);
  nan_policy(E, loglik, noos, logprior, err);
  __syncwarp();
  for (int v = lane; v < nR; v += 32) E.reals[(size_t)v * E.M_local + l] = P[v];
  for (int k = lane; k < S; k += 32) E.order[(size_t)l * E.s_max + k] = order[k];
  if (lane == 0) {
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
  if (pc.prof && lane == 0) {
    unsigned long long* q = pc.prof + ((size_t)blockIdx.x * kPoolWarps + warp) * 8;
    q[4] += fam.seq; q[5] += fam.cyc_wait; q[6] += (unsigned long long)(clock64() - t_c0);
  }
  pool_fence();
  __syncwarp();
  if (lane == 0) atomicAdd(pc.finished, 1u);
}

template <class PF>
__global__ void __launch_bounds__(kPoolThreads, 1)
pool_kernel(DevEngine E, typename PF::Data D, PoolCtl pc, int mode) {
  __shared__ PoolSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((int)blockIdx.x >= pc.n_ctrl) for (int i = tid; i < pc.C; i += blockDim.x) sm.seen[i] = 0;
  if (tid == 0) {
    sm.q.head = 0; sm.q.tail = 0; sm.q.done = 0; sm.q.pubs = 0;
    if ((int)blockIdx.x >= pc.n_ctrl) {
      int lo, hi; PF::slab(D, blockIdx.x - pc.n_ctrl, pc.n_cta, lo, hi); sm.q.lo = lo; sm.q.hi = hi;
      bell_init(&sm.q.bell);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  logit_table_init();
  {
    extern __shared__ __align__(128) unsigned char pool_dyn[];
    if ((int)blockIdx.x >= pc.n_ctrl) pool_tables_init(pool_dyn);
  }
  __syncthreads();
  // The first n_ctrl CTAs only run samplers (one chain per warp), the others only serve rows: an SM that hosted
  // both was slower than its peers (one warp fewer, the sampler's code in its instruction cache), and every
  // evaluation waits for its slowest SM.
  if ((int)blockIdx.x < pc.n_ctrl) {
    const int l = blockIdx.x * kPoolWarps + warp;
    if (l < pc.C) pool_controller<PF>(E, D, pc, sm, warp, lane, l, mode);
    return;
  }
  const int cta = blockIdx.x - pc.n_ctrl;
  if (warp == kPoolWarps - 1) { pool_dispatcher(pc, sm, lane); return; }
  pool_worker<PF>(D, pc, sm, warp, lane, cta, E.err);
}

}  // namespace bpt
