// families.cuh -- the closed catalogue of model families the engine runs.
//
// Each family is what the Java shim lowers a recognised Blang model to
// (INTEGRATION.md).  A family answers four questions for one chain:
//   fixed(k, x)     sum of the FIXED (prior) factors connected to sampler k's
//                   variable, evaluated with that variable set to x
//   annealed(k, x)  sum of the finite pre-annealing values of the ANNEALED
//                   factors connected to sampler k + how many of them are -inf
//                   (a group reduction over the observations)
//   commit(k, x)    write the accepted value (and refresh per-observation caches)
//   full()          preAnnealedLogLikelihood, nOutOfSupport and the fixed sum over
//                   EVERY factor (SampledModel.updateAll, R/runtime/SampledModel.java:340-364)
// Factor bodies follow R/distributions/*.bl term by term;
// what is regrouped (e.g. N identical -0.5*log(variance) terms summed as one
// product) changes results only at rounding level and is inside the 1e-9 bar.
//
// Data layout in HBM: observations are streamed with 16-byte loads
// (double2 / paired rows); per-chain parameters live in shared memory (P[]).
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace bpt {

constexpr int kMaxParams = 64;    // latent reals per chain held in shared memory
constexpr int kMaxSamplers = 65;

enum SamplerType : int { S_REAL_SLICE = 0, S_SIMPLEX = 1 };

// rows [lo, hi) of this CTA inside the chain group; lo is even
struct RowRange { int lo, hi; };
BPT_D RowRange group_rows(int n, int G, int crank) {
  int chunk = (n + G - 1) / G;
  chunk = (chunk + 1) & ~1;
  RowRange r;
  r.lo = min(n, crank * chunk);
  r.hi = min(n, r.lo + chunk);
  return r;
}

BPT_D void acc_term(double t, double& s, int& c, int* err) {
  if (t > BPT_NEG_INF) s += t;
  else { c++; if (t != t) atomicOr(err, ERR_NAN); }
}

// Normal.bl:21-24 with constant mean/variance (a prior on x)
BPT_D double normal_f2(double mean, double variance, double x) {
  if (variance <= 0.0) return BPT_NEG_INF;
  const double d = mean - x;
  return -0.5 * (d * d) / variance;
}
// Gamma.bl:14-23, the two factors that read the realization
BPT_D double gamma_ab(double shape, double rate, double x) {
  if (shape <= 0.0 || rate <= 0.0) return BPT_NEG_INF;
  if (x <= 0.0) return BPT_NEG_INF;
  return (shape - 1.0) * log(x * rate) + (-x * rate);
}
// Gamma.bl:24-31, the two constant factors
BPT_D double gamma_cd(double shape, double rate) {
  if (shape <= 0.0 || rate <= 0.0) return BPT_NEG_INF;
  return -lgamma(shape) + log(rate);
}

// ===========================================================================
// C1  Normal with unknown mean and variance.  P = [mu, sigma2]
// ===========================================================================
struct NormalFam {
  struct Data {
    const double* y; int n;
    double m0, v0, rate;
  };
  static constexpr int kThreads = 128;
  Data d; double* P; RowRange rr; int* err;

  BPT_D static int n_reals(const Data&) { return 2; }
  BPT_D static int n_samplers(const Data&) { return 2; }
  BPT_D void init(const Data& d_, int, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; rr = group_rows(d.n, G, crank); err = err_;
  }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D double fixed(int k, double x, int) const {
    return k == 0 ? normal_f2(d.m0, d.v0, x) : gamma_ab(1.0, d.rate, x);   // Exponential.bl:11 -> Gamma(1.0, rate)
  }
  BPT_D double sum_sq(double mu) const {
    double s = 0.0;
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
      if (i + 1 < rr.hi) {
        const double2 v = *reinterpret_cast<const double2*>(d.y + i);
        const double a = mu - v.x, b = mu - v.y;
        s = fma(a, a, s); s = fma(b, b, s);
      } else { const double a = mu - d.y[i]; s = fma(a, a, s); }
    }
    return s;
  }
  // mu: the n F2 factors; sigma2: the n F1 and the n F2 factors (Normal.bl:17-24)
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double mu = k == 0 ? x : P[0], s2 = k == 1 ? x : P[1];
    if (s2 <= 0.0) { s = 0.0; c = (k == 0 ? 1 : 2) * d.n; return; }
    double q = sum_sq(mu); int z = 0;
    red.reduce(q, z);
    s = -0.5 * q / s2;
    if (k == 1) s += d.n * (-0.5 * log(s2));
    c = 0;
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    const double mu = P[0], s2 = P[1];
    loglik = d.n * (-0.5 * BPT_LOG_2PI);
    if (s2 <= 0.0) noos = 2 * d.n;
    else {
      double q = sum_sq(mu); int z = 0;
      red.reduce(q, z);
      loglik += d.n * (-0.5 * log(s2)) + (-0.5 * q / s2);
      noos = 0;
    }
    logprior = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0)) + normal_f2(d.m0, d.v0, mu) +
               gamma_ab(1.0, d.rate, s2) + gamma_cd(1.0, d.rate);
  }
  BPT_D void forward(Rng& rng) {   // declaration order: mu then sigma2
    const double mu = gen_normal(rng, d.m0, d.v0);
    const double s2 = gen_exponential(rng, d.rate);
    P[0] = mu; P[1] = s2;
  }
};

// ===========================================================================
// C2  Bayesian logistic regression.  P = w[0..p)
//     per-chain cache eta_i = x_i . w in HBM; a move of w_k needs only
//     column k of X:  eta_i(x) = eta_i + (x - w_k) * X[i][k]
// ===========================================================================
struct LogisticFam {
  struct Data {
    const double* xt;      // [p][npad] column-major copy of the design matrix
    const uint8_t* y;      // [npad]
    double* eta;           // [localSlots][npad]
    int n, npad, p;
    double v0;
  };
  static constexpr int kThreads = 512;
  Data d; double* P; double* eta; RowRange rr; int* err;

  BPT_D static int n_reals(const Data& d) { return d.p; }
  BPT_D static int n_samplers(const Data& d) { return d.p; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; eta = d.eta + (size_t)local_slot * d.npad; rr = group_rows(d.n, G, crank); err = err_;
  }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D double fixed(int, double x, int) const { return normal_f2(0.0, d.v0, x); }

  // Bernoulli.bl:11-14 -> Categorical.bl:12-15 with p = logistic(eta): the reference's naive form
  BPT_D void row_term(double z, int y, double& s, int& c) const {
    const double pr = 1.0 / (1.0 + exp(-z));
    const double pick = y ? pr : 1.0 - pr;
    acc_term(log(pick), s, c, err);
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double delta = x - P[k];
    const double* col = d.xt + (size_t)k * d.npad;
    double acc = 0.0; int cnt = 0;
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
      if (i + 1 < rr.hi) {
        const double2 e = *reinterpret_cast<const double2*>(eta + i);
        const double2 xv = *reinterpret_cast<const double2*>(col + i);
        const uchar2 yy = *reinterpret_cast<const uchar2*>(d.y + i);
        row_term(fma(delta, xv.x, e.x), yy.x, acc, cnt);
        row_term(fma(delta, xv.y, e.y), yy.y, acc, cnt);
      } else row_term(fma(delta, col[i], eta[i]), d.y[i], acc, cnt);
    }
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int) {
    const double delta = x - P[k];
    const double* col = d.xt + (size_t)k * d.npad;
    if (delta != 0.0)
      for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
        if (i + 1 < rr.hi) {
          double2 e = *reinterpret_cast<const double2*>(eta + i);
          const double2 xv = *reinterpret_cast<const double2*>(col + i);
          e.x = fma(delta, xv.x, e.x); e.y = fma(delta, xv.y, e.y);
          *reinterpret_cast<double2*>(eta + i) = e;
        } else eta[i] = fma(delta, col[i], eta[i]);
      }
    __syncthreads();   // every thread has read P[k] for its delta before anyone overwrites it
    P[k] = x;
  }
  // rebuild eta from scratch: sum_j w_j * X[i][j], j ascending from 0.0 (F/SpikedGLM.bl:29-33)
  BPT_D void refresh() {
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
      const bool two = i + 1 < rr.hi;
      double a = 0.0, b = 0.0;
      for (int j = 0; j < d.p; j++) {
        const double w = P[j];
        const double* col = d.xt + (size_t)j * d.npad;
        if (two) { const double2 xv = *reinterpret_cast<const double2*>(col + i); a = fma(w, xv.x, a); b = fma(w, xv.y, b); }
        else a = fma(w, col[i], a);
      }
      eta[i] = a; if (two) eta[i + 1] = b;
    }
  }
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) {
    refresh();
    double acc = 0.0; int cnt = 0;
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
      row_term(eta[i], d.y[i], acc, cnt);
      if (i + 1 < rr.hi) row_term(eta[i + 1], d.y[i + 1], acc, cnt);
    }
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    double lp = 0.0;
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int j = 0; j < d.p; j++) lp += f01 + normal_f2(0.0, d.v0, P[j]);
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {
    for (int j = 0; j < d.p; j++) { const double w = gen_normal(rng, 0.0, d.v0); P[j] = w; }
  }
};

// ===========================================================================
// C3  K-component Gaussian mixture, indicators marginalised.
//     P = [mu[K], var[K], pi[K]];  samplers: k<K mu_k, K<=k<2K var_k, 2K simplex
//     per observation: log( sum_k pi_k * exp(logN_k(y_i)) )   (naive form, F/Multimodal.bl:12-14)
// ===========================================================================
template <int KT>
struct MixtureFam {
  struct Data {
    const double* y; int n; int K;
    double m0, v0, gshape, grate, conc;
  };
  static constexpr int kThreads = 512;
  Data d; double* P; double* Cst; RowRange rr; int* err;   // Cst: [4][KT] per-evaluation component constants

  BPT_D static int n_reals(const Data& d) { return 3 * d.K; }
  BPT_D static int n_samplers(const Data& d) { return 2 * d.K + 1; }
  BPT_D void init(const Data& d_, int, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; Cst = P_ + kMaxParams; rr = group_rows(d.n, G, crank); err = err_;
  }
  BPT_D int sampler_type(int k) const { return k == 2 * d.K ? S_SIMPLEX : S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return d.K; }
  // candidate-substituted parameter of component c
  BPT_D void substituted(int k, double x, int aux, int c, double& mu, double& var, double& pi) const {
    const int K = d.K;
    mu = P[c]; var = P[K + c]; pi = P[2 * K + c];
    if (k < K) { if (c == k) mu = x; }
    else if (k < 2 * K) { if (c == k - K) var = x; }
    else {   // SimplexWritableVariable.set :26-30
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
  BPT_D void stage_constants(int k, double x, int aux) {
    __syncthreads();
    if (threadIdx.x < d.K) {
      const int c = threadIdx.x;
      double mu, var, pi; substituted(k, x, aux, c, mu, var, pi);
      Cst[0 * KT + c] = mu;
      Cst[1 * KT + c] = var <= 0.0 ? BPT_NEG_INF : -0.5 * BPT_LOG_2PI + (-0.5 * log(var));
      Cst[2 * KT + c] = var <= 0.0 ? 0.0 : -0.5 / var;
      Cst[3 * KT + c] = pi;
    }
    __syncthreads();
  }
  BPT_D void obs_term(double yv, const double* mu, const double* a, const double* b, const double* pi, double& s,
                      int& c) const {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < KT; q++)
      if (q < d.K) { const double dd = mu[q] - yv; acc += pi[q] * exp(fma(b[q], dd * dd, a[q])); }
    acc_term(log(acc), s, c, err);
  }
  BPT_D void sum_obs(double& s, int& c) const {
    double mu[KT], a[KT], b[KT], pi[KT];
#pragma unroll
    for (int q = 0; q < KT; q++) { mu[q] = Cst[q]; a[q] = Cst[KT + q]; b[q] = Cst[2 * KT + q]; pi[q] = Cst[3 * KT + q]; }
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
      if (i + 1 < rr.hi) {
        const double2 v = *reinterpret_cast<const double2*>(d.y + i);
        obs_term(v.x, mu, a, b, pi, s, c); obs_term(v.y, mu, a, b, pi, s, c);
      } else obs_term(d.y[i], mu, a, b, pi, s, c);
    }
  }
  BPT_D void annealed(int k, double x, int aux, GroupReducer& red, double& s, int& c) {
    stage_constants(k, x, aux);
    double acc = 0.0; int cnt = 0;
    sum_obs(acc, cnt);
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int aux) {
    const int K = d.K;
    if (k < 2 * K) { P[k] = x; return; }
    const int nx = aux == K - 1 ? 0 : aux + 1;
    const double comp = fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x);
    __syncthreads();
    P[2 * K + aux] = x; P[2 * K + nx] = comp;
    __syncthreads();
  }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) {
    const int K = d.K;
    stage_constants(0, P[0], 0);
    double acc = 0.0; int cnt = 0;
    sum_obs(acc, cnt);
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    double lp = 0.0;
    // F/MixtureModel.bl:14-20 order: Dirichlet, then per component Normal, Gamma
    for (int c = 0; c < K; c++) lp += (d.conc - 1.0) * log(P[2 * K + c]);
    lp += d.conc < 0.0 ? BPT_NEG_INF : (-K * lgamma(d.conc) + lgamma(K * d.conc));
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int c = 0; c < K; c++) {
      lp += f01 + normal_f2(d.m0, d.v0, P[c]);
      lp += gamma_ab(d.gshape, d.grate, P[K + c]) + gamma_cd(d.gshape, d.grate);
    }
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {
    const int K = d.K;
    double g[KT]; double sum = 0.0;
    // Generators.dirichletInPlace :233-272
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
    for (int c = 0; c < K; c++) {
      if (g[c] == 0.0) g[c] = 1e-300; else if (g[c] == 1.0) g[c] = 1.0 - 1e-16;
    }
    double mu[KT], var[KT];
    for (int c = 0; c < K; c++) { mu[c] = gen_normal(rng, d.m0, d.v0); var[c] = gen_gamma(rng, d.gshape, d.grate); }
    __syncthreads();
    for (int c = 0; c < K; c++) { P[c] = mu[c]; P[K + c] = var[c]; P[2 * K + c] = g[c]; }
    __syncthreads();
  }
};

// ===========================================================================
// C5  Beta-binomial.  P = [alpha, beta]; one factor per group (BetaBinomial.bl:17-30)
// ===========================================================================
struct BetaBinomialFam {
  struct Data {
    const int32_t* y; const int32_t* nt;   // [R][n]
    const double* cst;                     // [R][n]  lnGamma(n+1) - lnGamma(y+1) - lnGamma(n-y+1), -inf if unsupported
    int n; int chains_per_replica;
    double rate;
  };
  static constexpr int kThreads = 32;
  Data d; double* P; const int32_t* y; const int32_t* nt; const double* cst; RowRange rr; int* err;

  BPT_D static int n_reals(const Data&) { return 2; }
  BPT_D static int n_samplers(const Data&) { return 2; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; err = err_;
    const size_t rep = (size_t)(local_slot / d.chains_per_replica);
    y = d.y + rep * d.n; nt = d.nt + rep * d.n; cst = d.cst + rep * d.n;
    rr = group_rows(d.n, G, crank);
  }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D double fixed(int, double x, int) const { return gamma_ab(1.0, d.rate, x); }
  BPT_D void sum_groups(double a, double b, double& s, int& c) const {
    if (a <= 0.0 || b <= 0.0) {
      for (int g = rr.lo + threadIdx.x; g < rr.hi; g += blockDim.x) c++;
      return;
    }
    const double ab = lgamma(a + b) - lgamma(a) - lgamma(b);
    for (int g = rr.lo + threadIdx.x; g < rr.hi; g += blockDim.x) {
      const double yy = (double)y[g], n = (double)nt[g];
      const double t = cst[g] + lgamma(yy + a) + lgamma(n - yy + b) + ab - lgamma(n + a + b);
      acc_term(t, s, c, err);
    }
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double a = k == 0 ? x : P[0], b = k == 1 ? x : P[1];
    double acc = 0.0; int cnt = 0;
    sum_groups(a, b, acc, cnt);
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    double acc = 0.0; int cnt = 0;
    sum_groups(P[0], P[1], acc, cnt);
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    logprior = gamma_ab(1.0, d.rate, P[0]) + gamma_cd(1.0, d.rate) + gamma_ab(1.0, d.rate, P[1]) + gamma_cd(1.0, d.rate);
  }
  BPT_D void forward(Rng& rng) {
    const double a = gen_exponential(rng, d.rate);
    const double b = gen_exponential(rng, d.rate);
    P[0] = a; P[1] = b;
  }
};

}  // namespace bpt
