// common.cuh -- RNG, generators and annealing helpers shared by every kernel.
//
// Randomness replaces bayonet.distributions.Random (not in the reference tree):
// a stateless Philox4x32-10 stream, key = (seed, replica), counter =
// (draw, step, scan, position | purpose << 24).  Streams are indexed by
// temperature POSITION like ParallelTempering.parallelRandomStreams
// (R/engines/ParallelTempering.java:180), never by GPU rank, so a fixed seed
// gives the same chain regardless of how slots are sharded.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cmath>

namespace bpt {

#define BPT_NEG_INF (-HUGE_VAL)
#define BPT_POS_INF (HUGE_VAL)
#define BPT_HD __host__ __device__ __forceinline__
#define BPT_D __device__ __forceinline__

enum Purpose : uint32_t { P_MOVE = 0, P_SHUFFLE = 1, P_SWAP = 2, P_SWAPDIR = 3, P_FORWARD = 4, P_INIT = 5 };

enum ErrBits : int { ERR_NAN = 1, ERR_SLICE_DIVERGED = 2, ERR_INT_SLICE_EMPTY = 4, ERR_POOL_TIMEOUT = 8, ERR_PEER_TIMEOUT = 16 };

BPT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

BPT_HD double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

struct Rng {
  uint32_t k0, k1, stream, scan, step, draw;
  BPT_HD Rng(uint32_t key0, uint32_t key1, uint32_t purpose, uint32_t position, uint32_t scan_, uint32_t step_)
      : k0(key0), k1(key1), stream(position | (purpose << 24)), scan(scan_), step(step_), draw(0) {}
  BPT_HD void block(uint32_t w[4]) { philox4x32_10(draw++, step, scan, stream, k0, k1, w); }
  // Random.nextDouble(): U in [0,1)
  BPT_HD double next_double() { uint32_t w[4]; block(w); return u53(w[0], w[1]); }
  // Random.nextInt(n): 64-bit multiply-high
  BPT_HD int next_int(int n) {
    uint32_t w[4]; block(w);
    uint64_t x = ((uint64_t)w[0] << 32) | w[1];
#ifdef __CUDA_ARCH__
    return (int)__umul64hi(x, (uint64_t)n);
#else
    return (int)(((unsigned __int128)x * (uint64_t)n) >> 64);
#endif
  }
  // bayonet Random.nextBernoulli(p)
  BPT_HD bool bernoulli(double p) { return next_double() < p; }
  // Random.nextGaussian(): Box-Muller on the two doubles of one block
  BPT_HD double gaussian() {
    uint32_t w[4]; block(w);
    double u1 = 1.0 - u53(w[0], w[1]);
    double u2 = u53(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
  }
};

// ---- R/distributions/Generators.java -------------------------------------
BPT_HD double gen_unit_exponential(Rng& r) { return -log(1.0 - r.next_double()); }          // :185-188, U in (0,1]
BPT_HD double gen_exponential(Rng& r, double rate) { return gen_unit_exponential(r) / rate; }  // :191-194
BPT_HD double gen_uniform(Rng& r, double mn, double mx) { return mn + r.next_double() * (mx - mn); }  // :197-202
BPT_HD double gen_normal(Rng& r, double mean, double var) { return r.gaussian() * sqrt(var) + mean; }  // :287-290
// :134-140 -> commons-math3 GammaDistribution.sample (Marsaglia-Tsang / Ahrens-Dieter)
BPT_HD double gen_gamma(Rng& r, double shape, double rate) {
  double scale = 1.0 / rate, result;
  if (shape < 1.0) {
    for (;;) {
      double u = r.next_double();
      double bGS = 1.0 + shape / 2.718281828459045235360287;
      double p = bGS * u;
      if (p <= 1.0) {
        double x = pow(p, 1.0 / shape);
        double u2 = r.next_double();
        if (u2 > exp(-x)) continue;
        result = scale * x; break;
      } else {
        double x = -log((bGS - p) / shape);
        double u2 = r.next_double();
        if (u2 > pow(x, shape - 1.0)) continue;
        result = scale * x; break;
      }
    }
  } else {
    double d = shape - 0.333333333333333333;
    double c = 1.0 / (3.0 * sqrt(d));
    for (;;) {
      double x = r.gaussian();
      double v = (1.0 + c * x) * (1.0 + c * x) * (1.0 + c * x);
      if (v <= 0.0) continue;
      double x2 = x * x;
      double u = r.next_double();
      if (u < 1.0 - 0.0331 * x2 * x2) { result = scale * d * v; break; }
      if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) { result = scale * d * v; break; }
    }
  }
  if (result == 0.0) result = 1e-300;   // ZERO_PLUS_EPS :322
  return result;
}
// :299-309 -> commons-math3 PoissonDistribution.sample() for mean < 40: multiply uniforms until below exp(-mean)
BPT_HD int gen_poisson_small(Rng& r, double mean) {
  const double p = exp(-mean);
  long n = 0; double prod = 1.0;
  while (n < 1000 * mean) {
    prod *= r.next_double();
    if (prod >= p) n++; else return (int)n;
  }
  return (int)n;
}
// :143-152.  commons-math3's BetaDistribution.sample() is Cheng's BB/BC rejection sampler; the draw stream is
// unpinned against the JVM anyway, so this uses the two-gamma identity X / (X + Y) and keeps the end-point guards
BPT_HD double gen_beta(Rng& r, double alpha, double beta) {
  const double x = gen_gamma(r, alpha, 1.0);
  const double y = gen_gamma(r, beta, 1.0);
  double result = x / (x + y);
  if (result == 0.0) result = 1e-300;
  if (result == 1.0) result = 1.0 - 1e-16;   // ONE_MINUS_EPS :323
  return result;
}

// ---- R/mcmc/internals/ExponentiatedFactor.java:82-87 ----------------------
BPT_HD double annealed_minus_infinity(double exponent) {
  if (exponent >= 1.0) return BPT_NEG_INF;
  return -exponent * 1E100;
}
// SampledModel.logDensity() :161-176 from the three cached scalars
BPT_HD double annealed_log_density(double fixed, double loglik, double noos, double beta, int anneal_support) {
  if (!anneal_support && noos > 0) return BPT_NEG_INF;
  return 0.0 + fixed + beta * loglik + (noos == 0 ? 0.0 : noos * annealed_minus_infinity(beta));
}

// NumericalUtils.logAdd
BPT_HD double log_add(double a, double b) {
  if (a == BPT_NEG_INF) return b;
  if (b == BPT_NEG_INF) return a;
  double mx = a > b ? a : b, mn = a > b ? b : a;
  return mx + log1p(exp(mn - mx));
}

#define BPT_LOG_2PI 1.8378770664093454835606594728112

}  // namespace bpt
