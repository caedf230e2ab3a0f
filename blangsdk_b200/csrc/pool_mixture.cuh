// pool_mixture.cuh -- BASELINE config 2 (K-component Gaussian mixture, indicators marginalised) on the pool kernel.
//
// Same split of roles as pool.cuh: controller warps run the samplers (explore_chain / slice.cuh), every worker SM owns
// a slab of the observations for ALL chains.  What MixtureFam (families.cuh) does inside one CTA per chain is done here
// per slab:
//   PR_FIRST  first evaluation of a sampler execution: all K components per row; stores, per row, the sum R of the
//             components the sampler does not move (and, for the simplex sampler, the two moved components'
//             unweighted densities) in the chain's scratch arrays in HBM;
//   PR_EVAL   the other evaluations of the execution: S = R + (moved terms at the candidate): one exp + one log per row
//             (no exp for a simplex move), 8 (24) bytes per row streamed from HBM, prefetched two steps ahead;
//   PR_FULL   from-scratch sum over all components (SampledModel.updateAll).
// The per-evaluation component constants (mu_c, -0.5 log 2 pi var_c, -0.5 / var_c, pi_c) travel through wbuf; the
// observations of a slab (54 KB at 1 M observations / 148 SMs) stay in the SM's L1 / L2.
// Rows whose mixture sum leaves [1e-250, 1e250] (every component underflowing, invalid variances, NaN) are redone
// with the literal arithmetic of F/Multimodal.bl:12-14, exactly as MixtureFam does.
#pragma once
#include "pool.cuh"

namespace bpt {

enum { PR_FIRST = 3 };

template <int KT>
struct MixturePool {
  using Data = typename MixtureFam<KT>::Data;

  BPT_D static void slab(const Data& D, int cta, int n_cta, int& lo, int& hi) {
    const long long U = (D.n + 63) >> 6;
    lo = (int)(U * cta / n_cta) << 6;
    hi = min(D.n, (int)(U * (cta + 1) / n_cta) << 6);
  }
  struct Cst { double mu[KT], a[KT], b[KT], pi[KT]; };
  BPT_D static bool in_range(double S) {   // 1e-250 < S < 1e250 to within one high-word step; NaN / inf / <= 0 fail
    const unsigned hi = (unsigned)__double2hiint(S);
    return hi - 0x0C100000u < (0x73D00000u - 0x0C100000u);
  }
  // exp_tab() / log_tab() (families.cuh) on the pool's bank-conflict-free replicated tables (pool_tables_init): the
  // same arithmetic, bit for bit; only the shared-memory addresses differ.  te / tl = this lane's column of the exp /
  // log table.  exp_rep clamps its argument below at -700 on the high word (one integer min instead of fmax's compare
  // and selects; arguments in (-700.016, -700) keep their own value, 1e-304 either way).
  BPT_D static double exp_rep(double x, uint32_t te) {
    x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC085E000u), __double2loint(x));
    double kd = fma(x, -kLgtK[0], kLgtK[1]);       // low word = m = round(x * 64 log2 e) = 64 k + j
    const int m = __double2loint(kd);
    kd -= kLgtK[1];
    uint32_t ea;
    asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 63;\n\tmad.lo.u32 %0, t, 128, %2;\n\t}" : "=r"(ea) : "r"(m), "r"(te));
    double tv;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(ea));
    double r = fma(kd, kLgtK[2], x);
    r = fma(kd, kLgtK[3], r);
    double p = kLgtExpC[0];
#pragma unroll
    for (int i = 1; i < 6; i++) p = fma(p, r, kLgtExpC[i]);
    const double tp = tv * p;
    int thi;
    asm("{\n\t.reg .s32 t;\n\tshr.s32 t, %1, 6;\n\tmad.lo.s32 %0, t, 1048576, %2;\n\t}" : "=r"(thi) : "r"(m), "r"(__double2hiint(tp)));
    return __hiloint2double(thi, __double2loint(tp));
  }
  BPT_D static double log_rep(double a, uint32_t tl) {   // log(a) for normal positive a
    const int hi = __double2hiint(a);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));   // [1, 2)
    double2 cj;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cj.x), "=d"(cj.y) : "r"(tl + ((((uint32_t)hi >> 12) & 0xffu) << 7)));
    const double rr = fma(m, cj.x, -1.0);
    double q = kLgtLogC[0];
#pragma unroll
    for (int i = 1; i < 4; i++) q = fma(q, rr, kLgtLogC[i]);
    const double lm = fma(rr * rr, q, rr) + cj.y;
    return fma((double)e, 6.93147180559945309417e-01, lm);
  }
  BPT_D static void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

  // literal form: log( sum_k pi_k * exp(logN_k(y)) ) with library exp / log (handles -inf, 0, NaN)
  __device__ __noinline__ static void obs_slow(double yv, const Cst& C, double& s, int& c, int* err) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < KT; q++) { const double dd = C.mu[q] - yv; acc += C.pi[q] * exp(fma(C.b[q], dd * dd, C.a[q])); }
    acc_term(log(acc), s, c, err);
  }
  // the same from the request's constants in wbuf (the cached passes do not keep all K components in registers)
  struct RowSum { double s; int c; };      // results by value: an accumulator whose address escapes lives in local memory
  __device__ __noinline__ static RowSum obs_slow_wb(double yv, const double* wb, double s, int* err) {
    Cst C;
#pragma unroll
    for (int q = 0; q < KT; q++) {
      C.mu[q] = pool_ld_f64(wb + q); C.a[q] = pool_ld_f64(wb + KT + q); C.b[q] = pool_ld_f64(wb + 2 * KT + q); C.pi[q] = pool_ld_f64(wb + 3 * KT + q);
    }
    RowSum r; r.s = s; r.c = 0;
    obs_slow(yv, C, r.s, r.c, err);
    return r;
  }
  // all components of one row: R = sum of the unmoved ones, S = all, E1 / E2 = the moved components' densities
  BPT_D static void lean(double yv, const Cst& C, int j0, int j1, uint32_t te, double& R, double& S, double& E1, double& E2) {
    double M = 0.0; R = 0.0; E1 = 0.0; E2 = 0.0;
#pragma unroll
    for (int q0 = 0; q0 < KT; q0 += 4) {
      double arg[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const double dd = C.mu[q0 + q] - yv; arg[q] = fma(C.b[q0 + q], dd * dd, C.a[q0 + q]); }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double e = exp_rep(arg[q], te);     // clamps at -700
        const bool moved = (q0 + q == j0) || (q0 + q == j1);
        if (q0 + q == j0) E1 = e;
        if (q0 + q == j1) E2 = e;
        if (moved) M = fma(C.pi[q0 + q], e, M); else R = fma(C.pi[q0 + q], e, R);
      }
    }
    S = R + M;
  }

  // PR_FIRST (STORE = true) and PR_FULL (STORE = false) on rows [lo, hi) of chain c
  template <bool STORE>
  BPT_D static void first_rows(const Data& D, int c, int lo, int hi, const Cst& C, int j0, int j1, int lane, uint32_t tab, int* err,
                               double& acc, int& cnt) {
    double* rest = D.rest + (size_t)c * D.npad; double* e1 = D.e1 + (size_t)c * D.npad; double* e2 = D.e2 + (size_t)c * D.npad;
    const int n_pairs = (hi - lo) >> 1;
    const double2* py = reinterpret_cast<const double2*>(D.y + lo) + lane;
    extern __shared__ __align__(128) unsigned char pool_dyn[];
    uint32_t tl = (uint32_t)__cvta_generic_to_shared(pool_dyn) + ((uint32_t)lane & 7u) * 16u;
    uint32_t te = (uint32_t)__cvta_generic_to_shared(pool_dyn) + kPoolLogTabBytes + ((uint32_t)lane & 15u) * 8u;
    asm volatile("" : "+r"(tl), "+r"(te));
    // one row, literal fallback included; r_out = NaN marks a row the cached passes must redo literally
    auto one = [&](double yv, double& r_out, double& x1, double& x2) {
      double R, S;
      lean(yv, C, j0, j1, te, R, S, x1, x2);
      if (in_range(S)) { acc += log_rep(S, tl); r_out = R; }
      else { obs_slow(yv, C, acc, cnt, err); r_out = CUDART_NAN; }
    };
    auto store = [&](int pair, double2 r, double2 x1, double2 x2) {
      if (!STORE) return;
      const int i = lo + 2 * pair;
      *reinterpret_cast<double2*>(rest + i) = r;
      if (j1 >= 0) { *reinterpret_cast<double2*>(e1 + i) = x1; *reinterpret_cast<double2*>(e2 + i) = x2; }
    };
    int j = 0;
    while (j + 32 <= n_pairs) {          // blocks of up to 32 full-warp steps, literal steps deferred (see pool.cuh)
      const int j_block = j;
      unsigned bad = 0u, bit = 1u;
      const int j_end = min(n_pairs - 31, j_block + 32 * 32);
      for (; j < j_end; j += 32, bit <<= 1) {
        const double2 v = py[j];
        double2 r, x1, x2; double S0, S1;
        lean(v.x, C, j0, j1, te, r.x, S0, x1.x, x2.x); lean(v.y, C, j0, j1, te, r.y, S1, x1.y, x2.y);
        const double l0 = log_rep(S0, tl), l1 = log_rep(S1, tl);
        const bool all = __all_sync(0xffffffffu, in_range(S0) && in_range(S1));
        acc = all ? (acc + l0) + l1 : acc;
        bad |= all ? 0u : bit;
        store(j + lane, r, x1, x2);
      }
      while (bad) {
        const int jb = j_block + 32 * (__ffs(bad) - 1);
        bad &= bad - 1u;
        const double2 v = py[jb];
        double2 r, x1, x2;
        one(v.x, r.x, x1.x, x2.x); one(v.y, r.y, x1.y, x2.y);
        store(jb + lane, r, x1, x2);
      }
    }
    if (j + lane < n_pairs) {            // ragged rest (last SM only)
      const double2 v = py[j];
      double2 r, x1, x2;
      one(v.x, r.x, x1.x, x2.x); one(v.y, r.y, x1.y, x2.y);
      store(j + lane, r, x1, x2);
    }
    if (((hi - lo) & 1) && lane == 0) {  // odd last row of the data set
      const int i = hi - 1;
      double r, x1, x2;
      one(D.y[i], r, x1, x2);
      if (STORE) { rest[i] = r; if (j1 >= 0) { e1[i] = x1; e2[i] = x2; } }
    }
  }

  // PR_EVAL on rows [lo, hi): S_i = R_i + (moved terms at the candidate)
  template <bool SIMPLEX>
  BPT_D static void cached_rows(const Data& D, int c, int lo, int hi, const double* wb, int j0, int j1, int lane, uint32_t tab, int* err,
                                double& acc, int& cnt) {
    const double* rest = D.rest + (size_t)c * D.npad; const double* e1 = D.e1 + (size_t)c * D.npad; const double* e2 = D.e2 + (size_t)c * D.npad;
    const double mu0 = pool_ld_f64(wb + j0), a0 = pool_ld_f64(wb + KT + j0), b0 = pool_ld_f64(wb + 2 * KT + j0), pi0 = pool_ld_f64(wb + 3 * KT + j0);
    const double pi1 = SIMPLEX ? pool_ld_f64(wb + 3 * KT + j1) : 0.0;
    const int n_pairs = (hi - lo) >> 1;
    const double2* py = reinterpret_cast<const double2*>(D.y + lo) + lane;
    const double2* pr = reinterpret_cast<const double2*>(rest + lo) + lane;
    const double2* p1 = reinterpret_cast<const double2*>(e1 + lo) + lane;
    const double2* p2 = reinterpret_cast<const double2*>(e2 + lo) + lane;
    // opaque: the stream pointers stay in registers (the compiler otherwise re-derives them from the kernel
    // arguments inside the hot loop, a dozen instructions per step)
    asm volatile("" : "+l"(py), "+l"(pr));
    if (SIMPLEX) asm volatile("" : "+l"(p1), "+l"(p2));
    extern __shared__ __align__(128) unsigned char pool_dyn[];
    uint32_t tl = (uint32_t)__cvta_generic_to_shared(pool_dyn) + ((uint32_t)lane & 7u) * 16u;
    uint32_t te = (uint32_t)__cvta_generic_to_shared(pool_dyn) + kPoolLogTabBytes + ((uint32_t)lane & 15u) * 8u;
    asm volatile("" : "+r"(tl), "+r"(te));   // opaque: kept in registers instead of being re-derived per use
    auto value = [&](double yv, double R, double E1, double E2) {
      if (!SIMPLEX) { const double dd = mu0 - yv; return fma(pi0, exp_rep(fma(b0, dd * dd, a0), te), R); }
      return fma(pi0, E1, fma(pi1, E2, R));
    };
    // out-of-line step for the rare literal rows: same arithmetic as the lean path where the lean path applies
    auto one = [&](double yv, double R, double E1, double E2) {
      const double S = value(yv, R, E1, E2);
      if (in_range(S)) acc += log_rep(S, tl);
      else { const RowSum r = obs_slow_wb(yv, wb, acc, err); acc = r.s; cnt += r.c; }
    };
    int j = 0;
    if (n_pairs >= 32) {
      // The scratch arrays stream from HBM (no reuse inside an evaluation).  One lane in eight asks the L2 for the
      // 128-byte line kAhead steps ahead; the registers are refilled one step ahead, right after their last use
      // (the loads then find their line in the L2), so no register set is rotated.  y is cache-resident.
      constexpr int kAhead = 6;
      bool pf_lane = (lane & 7) == 0;
      { int t = pf_lane; asm volatile("" : "+r"(t)); pf_lane = t != 0; }   // opaque: not re-derived from %tid per step
      if (pf_lane) {
#pragma unroll
        for (int a = 1; a < kAhead; a++)
          if (32 * a + 32 <= n_pairs) { prefetch_l2(pr + 32 * a); if (SIMPLEX) { prefetch_l2(p1 + 32 * a); prefetch_l2(p2 + 32 * a); } }
      }
      double2 r0 = __ldcs(pr), v = py[0];
      double2 f0 = make_double2(0.0, 0.0), g0 = f0;
      if (SIMPLEX) { f0 = __ldcs(p1); g0 = __ldcs(p2); }
      while (j + 32 <= n_pairs) {
        const int j_block = j;
        unsigned bad = 0u, bit = 1u;
        const int j_end = min(n_pairs - 31, j_block + 32 * 32);
        for (; j < j_end; j += 32, bit <<= 1) {
          const double S0 = value(v.x, r0.x, f0.x, g0.x), S1 = value(v.y, r0.y, f0.y, g0.y);
          if (j + 64 <= n_pairs) {
            r0 = __ldcs(pr + j + 32); v = py[j + 32];
            if (SIMPLEX) { f0 = __ldcs(p1 + j + 32); g0 = __ldcs(p2 + j + 32); }
          }
          if (pf_lane && j + 32 * kAhead + 32 <= n_pairs) {
            prefetch_l2(pr + j + 32 * kAhead);
            if (SIMPLEX) { prefetch_l2(p1 + j + 32 * kAhead); prefetch_l2(p2 + j + 32 * kAhead); }
          }
          const double l0 = log_rep(S0, tl), l1 = log_rep(S1, tl);
          const bool all = __all_sync(0xffffffffu, in_range(S0) && in_range(S1));
          acc = all ? (acc + l0) + l1 : acc;
          bad |= all ? 0u : bit;
        }
        while (bad) {
          const int jb = j_block + 32 * (__ffs(bad) - 1);
          bad &= bad - 1u;
          const double2 vb = py[jb], rb = pr[jb];
          double2 fb = make_double2(0.0, 0.0), gb = fb;
          if (SIMPLEX) { fb = p1[jb]; gb = p2[jb]; }
          one(vb.x, rb.x, fb.x, gb.x); one(vb.y, rb.y, fb.y, gb.y);
        }
      }
    }
    if (j + lane < n_pairs) {
      const double2 vb = py[j], rb = pr[j];
      double2 fb = make_double2(0.0, 0.0), gb = fb;
      if (SIMPLEX) { fb = p1[j]; gb = p2[j]; }
      one(vb.x, rb.x, fb.x, gb.x); one(vb.y, rb.y, fb.y, gb.y);
    }
    if (((hi - lo) & 1) && lane == 0) {
      const int i = hi - 1;
      one(D.y[i], rest[i], SIMPLEX ? e1[i] : 0.0, SIMPLEX ? e2[i] : 0.0);
    }
  }

  BPT_D static bool serve(const Data& D, const PoolCtl& pc, int c, int lo, int hi, const PoolHalf& a, const PoolHalf&, int lane,
                          uint32_t tab, int* err, double& s, int& cnt) {
    s = 0.0; cnt = 0;
    const int type = (a.i >> 8) & 0xf, k = a.i & 0xff, aux = (a.i >> 13) & 7, K = D.K;
    const int j0 = k < K ? k : (k < 2 * K ? k - K : aux);
    const int j1 = k < 2 * K ? -1 : (aux == K - 1 ? 0 : aux + 1);
    // this evaluation's component constants, written by the controller before the request (fenced)
    const double* wb = pc.wbuf + (size_t)c * kMaxParams;
    if (type == PR_EVAL) {
      if (j1 < 0) cached_rows<false>(D, c, lo, hi, wb, j0, j1, lane, tab, err, s, cnt);
      else cached_rows<true>(D, c, lo, hi, wb, j0, j1, lane, tab, err, s, cnt);
      return false;
    }
    const double mine = lane < 4 * KT ? pool_ld_f64(wb + lane) : 0.0;     // lanes fetch one constant each
    Cst C;
#pragma unroll
    for (int q = 0; q < KT; q++) {
      C.mu[q] = __shfl_sync(0xffffffffu, mine, q); C.a[q] = __shfl_sync(0xffffffffu, mine, KT + q);
      C.b[q] = __shfl_sync(0xffffffffu, mine, 2 * KT + q); C.pi[q] = __shfl_sync(0xffffffffu, mine, 3 * KT + q);
    }
    if (type == PR_FIRST) { first_rows<true>(D, c, lo, hi, C, j0, j1, lane, tab, err, s, cnt); return true; }
    first_rows<false>(D, c, lo, hi, C, -1, -1, lane, tab, err, s, cnt);   // PR_FULL
    return false;
  }

  // ---- the controller warp's view of the family ----
  struct Ctl {
    using Data = typename MixtureFam<KT>::Data;
    Data d; double* P; PoolCtl pc; int l, lane; uint32_t seq; bool aborted, first; int* lerr;
    unsigned long long cyc_wait;
    BPT_D static int n_reals(const Data& d) { return 3 * d.K; }
    BPT_D static int n_samplers(const Data& d) { return 2 * d.K + 1; }
    BPT_D void init(const Data& d_, int local_slot, double* P_, const PoolCtl& pc_, int* lerr_) {
      d = d_; P = P_; pc = pc_; l = local_slot; lane = threadIdx.x & 31; seq = 0; aborted = false; first = true; lerr = lerr_; cyc_wait = 0;
    }
    BPT_D int sampler_type(int k) const { return k == 2 * d.K ? S_SIMPLEX : S_REAL_SLICE; }
    BPT_D int sampler_var(int k) const { return k; }
    BPT_D int simplex_dim() const { return d.K; }
    BPT_D void begin_execution() { first = true; }
    BPT_D void substituted(int k, double x, int aux, int c, double& mu, double& var, double& pi) const {   // MixtureFam::substituted
      const int K = d.K;
      mu = P[c]; var = P[K + c]; pi = P[2 * K + c];
      if (k < K) { if (c == k) mu = x; }
      else if (k < 2 * K) { if (c == k - K) var = x; }
      else {
        const int nx = aux == K - 1 ? 0 : aux + 1;
        if (c == aux) pi = x;
        else if (c == nx) pi = fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x);
      }
    }
    BPT_D double fixed(int k, double x, int aux) const {
      const int K = d.K;
      if (k < K) return normal_f2(d.m0, d.v0, x);
      if (k < 2 * K) return gamma_ab(d.gshape, d.grate, x);
      double sum = 0.0;   // Dirichlet.bl:13-22
      for (int c = 0; c < K; c++) {
        double mu, var, pi; substituted(k, x, aux, c, mu, var, pi);
        if (d.conc < 0.0) return BPT_NEG_INF;
        sum += (d.conc - 1.0) * log(pi);
      }
      return sum;
    }
    // component constants of this evaluation -> wbuf (lane c writes component c; padding components contribute 0)
    BPT_D void stage(int k, double x, int aux) {
      __syncwarp();
      if (lane < KT) {
        double mu = 0.0, var = 1.0, pi = 0.0;
        if (lane < d.K) substituted(k, x, aux, lane, mu, var, pi);
        double* w = pc.wbuf + (size_t)l * kMaxParams;
        w[lane] = mu;
        w[KT + lane] = var <= 0.0 ? BPT_NEG_INF : -0.5 * BPT_LOG_2PI + (-0.5 * log(var));
        w[2 * KT + lane] = var <= 0.0 ? 0.0 : -0.5 / var;
        w[3 * KT + lane] = pi;
      }
      pool_fence();
      __syncwarp();
    }
    BPT_D void post(int type, int k, int aux) {
      seq++;
      if (lane == 0) pool_st_half(pc.reqa + l, 0.0, seq, k | (type << 8) | (aux << 13));
    }
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
    BPT_D void annealed(int k, double x, int aux, GroupReducer&, double& s, int& c) {
      if (aborted) { s = 0.0; c = 1; return; }
      stage(k, x, aux);
      post(first ? PR_FIRST : PR_EVAL, k, aux);
      first = false;
      wait(s, c);
    }
    BPT_D void commit(int k, double x, int aux) {
      const int K = d.K;
      const int nx = aux == K - 1 ? 0 : aux + 1;
      const double comp = k == 2 * K ? fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x) : 0.0;
      __syncwarp();
      if (lane == 0) { if (k == 2 * K) { P[2 * K + aux] = x; P[2 * K + nx] = comp; } else P[k] = x; }
      __syncwarp();
    }
    BPT_D void refresh() {}
    BPT_D void full(GroupReducer&, double& loglik, int& noos, double& logprior) {
      const int K = d.K;
      loglik = 0.0; noos = 1;
      if (!aborted) { stage(0, P[0], 0); post(PR_FULL, 0, 0); wait(loglik, noos); }
      double lp = 0.0;
      for (int c = 0; c < K; c++) lp += (d.conc - 1.0) * log(P[2 * K + c]);
      lp += d.conc < 0.0 ? BPT_NEG_INF : (-K * lgamma(d.conc) + lgamma(K * d.conc));
      const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
      for (int c = 0; c < K; c++) {
        lp += f01 + normal_f2(d.m0, d.v0, P[c]);
        lp += gamma_ab(d.gshape, d.grate, P[K + c]) + gamma_cd(d.gshape, d.grate);
      }
      logprior = lp;
    }
    BPT_D void forward(Rng& rng) {   // MixtureFam::forward with warp instead of block barriers
      const int K = d.K;
      double g[KT]; double sum = 0.0;
      if (d.conc < 1.0) {
        double lse = BPT_NEG_INF;
        for (int c = 0; c < K; c++) {
          const double gv = gen_gamma(rng, d.conc + 1.0, 1.0);
          const double logu = log(rng.next_double());
          g[c] = log(gv) + logu / d.conc;
          lse = log_add(lse, g[c]);
        }
        for (int c = 0; c < K; c++) g[c] = exp(g[c] - lse);
      } else {
        for (int c = 0; c < K; c++) { g[c] = gen_gamma(rng, d.conc, 1e7); sum += g[c]; }
        for (int c = 0; c < K; c++) g[c] /= sum;
      }
      for (int c = 0; c < K; c++) { if (g[c] == 0.0) g[c] = 1e-300; else if (g[c] == 1.0) g[c] = 1.0 - 1e-16; }
      double mu[KT], var[KT];
      for (int c = 0; c < K; c++) { mu[c] = gen_normal(rng, d.m0, d.v0); var[c] = gen_gamma(rng, d.gshape, d.grate); }
      __syncwarp();
      if (lane == 0) for (int c = 0; c < K; c++) { P[c] = mu[c]; P[K + c] = var[c]; P[2 * K + c] = g[c]; }
      __syncwarp();
    }
  };
};

}  // namespace bpt
