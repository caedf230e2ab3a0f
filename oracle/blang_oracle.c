/*
 * blang_oracle.c -- CPU restatement (plain C, fp64) of blangSDK's NRPT hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see blang_oracle.h for who may use it and for
 * the parity status ("parity unpinned" w.r.t. the JVM; pinned on the
 * reference's own known answers).
 *
 * Structure follows the reference one-to-one so a reviewer can put the two
 * side by side:
 *   Factor / factor_eval      <- the logf blocks of R/distributions/ *.bl
 *   exponentiated_*           <- R/mcmc/internals/ExponentiatedFactor.java
 *   SampledModel (sm_*)       <- R/runtime/SampledModel.java
 *   real_slice / int_slice    <- R/mcmc/RealSliceSampler.java, IntSliceSampler.java
 *   simplex_/categorical_     <- R/mcmc/SimplexSampler.xtend, CategoricalSampler.xtend
 *   orc_pt_*                  <- R/engines/ParallelTempering.java, factories/PT.java
 *   spline / adapt            <- R/engines/internals/Spline.java, EngineStaticUtils.java
 *   paths                     <- R/engines/internals/ptanalysis/Paths.xtend
 *
 * Randomness: bayonet.distributions.Random (not in the tree) is replaced by a
 * stateless Philox4x32-10 counter stream keyed (seed, replica) with counter
 * (draw, step, scan, stream|purpose<<24).  Streams are indexed by temperature
 * POSITION like parallelRandomStreams (ParallelTempering.java:180).  The CUDA
 * engine uses the identical layout, so trajectories are comparable
 * draw-for-draw.
 */
#include "blang_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NEG_INF (-INFINITY)

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123)                     */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { P_MOVE = 0, P_SHUFFLE = 1, P_SWAP = 2, P_SWAPDIR = 3, P_FORWARD = 4, P_INIT = 5 };

typedef struct {
  uint32_t key[2];
  uint32_t stream, scan, step, draw;
} Rng;

static Rng rng_make(uint64_t seed, uint32_t replica, int purpose, uint32_t position,
                    uint32_t scan, uint32_t step) {
  Rng r;
  r.key[0] = (uint32_t)seed;
  r.key[1] = (uint32_t)(seed >> 32) ^ replica;
  r.stream = position | ((uint32_t)purpose << 24);
  r.scan = scan; r.step = step; r.draw = 0;
  return r;
}
static void rng_block(Rng* r, uint32_t out[4]) {
  uint32_t c[4] = { r->draw++, r->step, r->scan, r->stream };
  orc_philox4x32_10(c, r->key, out);
}
static double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
/* Random.nextDouble(): U in [0,1) */
static double rng_double(Rng* r) { uint32_t w[4]; rng_block(r, w); return u53(w[0], w[1]); }
/* Random.nextInt(n): 64-bit multiply-high (bias < 2^-60) */
static int rng_int(Rng* r, int n) {
  uint32_t w[4]; rng_block(r, w);
  uint64_t x = ((uint64_t)w[0] << 32) | w[1];
  return (int)(((unsigned __int128)x * (uint64_t)n) >> 64);
}
/* bayonet Random.nextBernoulli(p) : nextDouble() < p */
static int rng_bernoulli(Rng* r, double p) { return rng_double(r) < p; }
/* Random.nextGaussian(): Box-Muller on the two doubles of one block */
static double rng_gaussian(Rng* r) {
  uint32_t w[4]; rng_block(r, w);
  double u1 = 1.0 - u53(w[0], w[1]);   /* (0,1] */
  double u2 = u53(w[2], w[3]);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

/* ------------------------------------------------------------------ */
/* R/distributions/Generators.java                                     */
/* ------------------------------------------------------------------ */
#define ZERO_PLUS_EPS 1e-300            /* Generators.java:322 */
#define ONE_MINUS_EPS (1.0 - 1e-16)     /* Generators.java:323 */

/* Generators.unitRateExponential :185-188.  Draws -log(U), U in (0,1]
 * (reference: [0,1); the 2^-53 event U=0 is excluded). */
static double gen_unit_exponential(Rng* r) { return -log(1.0 - rng_double(r)); }
static double gen_exponential(Rng* r, double rate) { return gen_unit_exponential(r) / rate; } /* :191-194 */
/* Generators.uniform :197-202 (right exclusive) */
static double gen_uniform(Rng* r, double mn, double mx) { return mn + rng_double(r) * (mx - mn); }
/* Generators.discreteUniform :205-222 */
static int gen_discrete_uniform(Rng* r, int mn, int mx) { return rng_int(r, mx - mn) + mn; }
/* Generators.normal :287-290 */
static double gen_normal(Rng* r, double mean, double var) { return rng_gaussian(r) * sqrt(var) + mean; }
/* Generators.gamma :134-140 -> commons-math3 GammaDistribution.sample():
 * Marsaglia & Tsang (2000) for shape >= 1, Ahrens-Dieter (1974) GS for shape < 1. */
static double gen_gamma(Rng* r, double shape, double rate) {
  double scale = 1.0 / rate, result;
  if (shape < 1.0) {
    for (;;) {
      double u = rng_double(r);
      double bGS = 1.0 + shape / 2.718281828459045235360287;
      double p = bGS * u;
      if (p <= 1.0) {
        double x = pow(p, 1.0 / shape);
        double u2 = rng_double(r);
        if (u2 > exp(-x)) continue;
        result = scale * x; break;
      } else {
        double x = -log((bGS - p) / shape);
        double u2 = rng_double(r);
        if (u2 > pow(x, shape - 1.0)) continue;
        result = scale * x; break;
      }
    }
  } else {
    double d = shape - 0.333333333333333333;
    double c = 1.0 / (3.0 * sqrt(d));
    for (;;) {
      double x = rng_gaussian(r);
      double v = (1.0 + c * x) * (1.0 + c * x) * (1.0 + c * x);
      if (v <= 0.0) continue;
      double x2 = x * x;
      double u = rng_double(r);
      if (u < 1.0 - 0.0331 * x2 * x2) { result = scale * d * v; break; }
      if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) { result = scale * d * v; break; }
    }
  }
  if (result == 0.0) result = ZERO_PLUS_EPS;
  return result;
}
/* Generators.poisson :299-309 -> commons-math3 PoissonDistribution.sample() for mean < 40: multiply uniforms
 * until the product falls below exp(-mean). */
static int gen_poisson_small(Rng* r, double mean) {
  const double p = exp(-mean);
  long n = 0; double prod = 1.0;
  while (n < 1000 * mean) {
    prod *= rng_double(r);
    if (prod >= p) n++; else return (int)n;
  }
  return (int)n;
}
/* Generators.beta :143-152.  commons-math3's BetaDistribution.sample() is Cheng's BB/BC rejection sampler; the
 * draw stream is unpinned against the JVM anyway, so the restatement uses the two-gamma identity
 * X / (X + Y), X ~ Gamma(alpha, 1), Y ~ Gamma(beta, 1), and keeps the reference's end-point guards. */
static double gen_beta(Rng* r, double alpha, double beta) {
  double x = gen_gamma(r, alpha, 1.0);
  double y = gen_gamma(r, beta, 1.0);
  double result = x / (x + y);
  if (result == 0.0) result = ZERO_PLUS_EPS;
  if (result == 1.0) result = ONE_MINUS_EPS;
  return result;
}

/* ------------------------------------------------------------------ */
/* bayonet.math (third party, restated from published definitions)      */
/* ------------------------------------------------------------------ */
/* SpecialFunctions.lnGamma.  DOCUMENTED CHOICE: libm lgamma (accurate to a
 * few ulp).  bayonet's own source is not in the tree; its doc signature
 * "double lnGamma(double alpha)" (R/runtime/internals/doc/contents/
 * BuiltInFunctions.xtend:61) suggests the Pike & Hill series below, which
 * differs from lgamma by < 3e-10 absolute for x>0 (tests/test_oracle_pins.py
 * measures it).  That is inside the 1e-9 relative bar for every catalogue
 * density but it is NOT bit parity: "parity unpinned". */
static int g_lngamma_impl = 0;   /* 0 = libm lgamma (the documented choice), 1 = the Pike & Hill restatement */
double orc_lngamma_pike_hill(double alpha);
void orc_set_lngamma(int which) { g_lngamma_impl = which; }
double orc_lngamma(double x) { return g_lngamma_impl ? orc_lngamma_pike_hill(x) : lgamma(x); }
double orc_lngamma_pike_hill(double alpha) {
  double x = alpha, f = 0.0, z;
  if (x < 7.0) {
    f = 1.0; z = x - 1.0;
    while (++z < 7.0) f *= z;
    x = z; f = -log(f);
  }
  z = 1.0 / (x * x);
  return f + (x - 0.5) * log(x) - x + 0.918938533204673 +
         (((-0.000595238095238 * z + 0.000793650793651) * z - 0.002777777777778) * z +
          0.083333333333333) / x;
}
/* SpecialFunctions.logistic: 1/(1+exp(-x)) (naive form) */
double orc_logistic(double x) { return 1.0 / (1.0 + exp(-x)); }
/* NumericalUtils.logAdd: log(exp(a)+exp(b)), stable */
double orc_log_add(double a, double b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  double mx = a > b ? a : b, mn = a > b ? b : a;
  return mx + log1p(exp(mn - mx));
}
/* NumericalUtils.isClose(a,b,tol): |a-b| < tol */
static int is_close(double a, double b, double tol) { return fabs(a - b) < tol; }

/* ------------------------------------------------------------------ */
/* Factors: one struct per `logf` block instance                        */
/* ------------------------------------------------------------------ */
enum {
  F_CONST = 0,
  F_NORMAL_F1,        /* Normal.bl:17-20   variance = real i0                            */
  F_NORMAL_F2_LIK,    /* Normal.bl:21-24   mean = real i0, variance = real i1, x = y[i2] */
  F_NORMAL_F2_PRIOR,  /* Normal.bl:21-24   mean = c0, variance = c1, x = real i0         */
  F_GAMMA_A,          /* Gamma.bl:14-18    shape=c0 rate=c1 x = real i0                  */
  F_GAMMA_B,          /* Gamma.bl:19-23    rate=c1 x = real i0                           */
  F_LOGISTIC_ROW,     /* F/SpikedGLM.bl:26-35 + Bernoulli.bl:11-14 + Categorical.bl:12-15 */
  F_MIXTURE_OBS,      /* F/Multimodal.bl:12-14 generalised to K components               */
  F_DIRICHLET_A,      /* Dirichlet.bl:13-22 conc = c0 (symmetric), simplex = reals i0..i0+i1-1 */
  F_ISING_EDGE,       /* F/Ising.bl:15-24  spins i0,i1, coupling c0                      */
  F_BERN_NODE,        /* F/Ising.bl:27-29 -> Bernoulli.bl -> Categorical.bl; spin i0, p=c0 */
  F_BETABINOM,        /* BetaBinomial.bl:17-30 group i0, alpha = real i1, beta = real i2 */
  F_IID_LAW,          /* family 6: law i1 of the observation distribution, observation i0 (iid_law below) */
  F_LINREG_F1, F_LINREG_F2,   /* family 8: Normal.bl:17-24 with mean = sum_j w_j x[i0][j], variance = sigma * sigma */
  F_HBETA,            /* family 7: block i2 of Beta.bl:14-41 with LATENT alpha = real 0, beta = real 1, realization = real i0 */
  F_HBINOM,           /* family 7: block i1 of Binomial.bl:14-27, group i0, p = real 2 + i0 */
  F_POIS_INT,         /* family 9: block i1 of Poisson.bl:11-26 on the latent int i0, mean = c0 */
  F_HBINOM_LAT,       /* family 9: block i1 of Binomial.bl:14-27, group i0, trials = 1 + int i0, p = real 2 + i0 */
  F_BETA_A, F_BETA_B,                      /* Beta.bl:14-27 prior on real i0, alpha = c0, beta = c1  */
  F_UNIF_B,                                /* ContinuousUniform.bl:16-19 prior on real i0, [c0, c1]  */
  F_CAT_LOGP,         /* family 10: Categorical.bl:12-15 log(probabilities.get(realization)): realization = int i0, simplex = reals i1.. (K = i2) */
  F_MIXIND_F1,        /* family 10: Normal.bl:17-20 with variance = variances.get(indicator i0)                    */
  F_MIXIND_F2,        /* family 10: Normal.bl:21-24 with mean / variance of component indicator i0, x = y[i0]      */
  F_DUNIF_B,          /* family 11: DiscreteUniform.bl:16-20 support of int i0 in [c0, c1)                         */
  F_MANUAL_NORMAL,    /* family 12: F/CustomAnnealTest.bl: Normal::distribution(mu, c1).logDensity(x = c0) * beta, beta = AnnealingParameter */
  F_CP_F2             /* family 11: Normal.bl:21-24, mean = (i0 < changePoint ? real 0 : real 1), variance = c0, x = y[i0] */
};

typedef struct {
  int32_t kind, annealed;   /* annealed: 0 fixed, 1 exponentiated (automatic annealing), 2 "other annealed": reads the
                               AnnealingParameter itself (GraphAnalysis.java:249-262), never cached (SampledModel.java:55) */
  int32_t i0, i1, i2;
  double c0, c1;
} Factor;

enum { S_REAL_SLICE = 0, S_SIMPLEX = 1, S_CATEGORICAL = 2, S_INT_SLICE = 3, S_UNIFORM = 4 };

typedef struct {
  int32_t type;
  int32_t var;        /* real index (REAL_SLICE), first simplex real (SIMPLEX), int index (CATEGORICAL) */
  int32_t K;          /* simplex dimension / number of categories / maxExclusive (S_UNIFORM) */
  int32_t lo;         /* minInclusive (S_UNIFORM, UniformSampler.xtend:23-27) */
  int32_t n_fixed, n_ann, n_other;
  int32_t *fixed_idx, *ann_idx;   /* sampler2sparseUpdate{Fixed,Annealed}, SampledModel.java:67-69 */
  int32_t *other_idx;             /* connected custom-annealed factors: evaluated by the sampler, not cached */
} Sampler;

enum { FW_NORMAL = 0, FW_EXPONENTIAL, FW_GAMMA, FW_DIRICHLET_SYM, FW_BERNOULLI_INT, FW_BETA, FW_UNIFORM, FW_BETA_LATENT, FW_POISSON_INT,
       FW_CATEGORICAL_INT, FW_DISCRETE_UNIFORM_INT };
typedef struct { int32_t kind, var, K; double a, b; } Forward;

struct orc_model {
  int family;
  int n_reals, n_ints;
  int n_factors; Factor* factors;
  int n_fixed_idx, n_ann_idx; int32_t *fixed_idx, *ann_idx;  /* sparseUpdate{Fixed,Annealed}Indices */
  int n_other_idx; int32_t* other_idx;                        /* otherAnnealedFactors, SampledModel.java:81 */
  int n_samplers; Sampler* samplers;
  int n_forward; Forward* forward;     /* GraphAnalysis.createForwardSimulator order */
  /* data */
  int n_obs, p, K, N;
  double* y; double* x; int32_t* yi; int32_t* ntrials;
  double* init_reals; int32_t* init_ints;
};

/* --- factor bodies ---------------------------------------------------- */
#define LOG_2PI 1.8378770664093454835606594728112

/* ---- family 6: the laws of R/distributions/<Name>.bl, one C case per `logf` block, in file order ----
 * th = the distribution's parameters (reals, declaration order), y = realization, x = trials (Binomial).
 * IID_MASK[d][l] = which parameters the block's signature names (bit j = parameter j): the samplers of those
 * parameters see the factor; IID_READS_Y = whether it names the realization. */
#define IID_MAX_LAWS 4
static const int IID_NPARAMS[17] = {0, 1, 1, 2, 3, 1, 2, 1, 2, 2, 2, 2, 2, 2, 2, 1, 1};
static const int IID_NLAWS[17]   = {0, 3, 2, 2, 4, 1, 1, 4, 4, 3, 2, 3, 3, 3, 2, 1, 2};
static const int IID_MASK[17][IID_MAX_LAWS] = {
  {0}, {1, 1, 0}, {1, 0}, {1, 3}, {0, 4, 1, 7}, {1}, {3}, {1, 1, 0, 1}, {3, 2, 1, 2}, {3, 3, 3}, {3, 3},
  {2, 3, 3}, {1, 2, 3}, {3, 3, 3}, {3, 3}, {1}, {1, 1}};
static const int IID_READS_Y[17][IID_MAX_LAWS] __attribute__((unused)) = {   /* documentation of the .bl signatures; no code path needs it */
  {0}, {1, 0, 1}, {1, 1}, {1, 1}, {0, 0, 0, 1}, {1}, {1}, {1, 1, 0, 0}, {1, 1, 0, 0}, {0, 1, 1}, {0, 1},
  {0, 1, 1}, {0, 0, 1}, {0, 1, 1}, {0, 1}, {1}, {0, 1}};

static double gamma_law(int law, double shape, double rate, double y) {   /* Gamma.bl:14-31 */
  switch (law) {
    case 0: if (shape <= 0.0 || rate <= 0) return NEG_INF; if (y <= 0.0) return NEG_INF; return (shape - 1.0) * log(y * rate);
    case 1: if (rate <= 0) return NEG_INF; if (y <= 0.0) return NEG_INF; return -y * rate;
    case 2: if (shape <= 0.0) return NEG_INF; return -orc_lngamma(shape);
    default: if (rate <= 0.0) return NEG_INF; return log(rate);
  }
}
static double iid_law(int dist, int law, const double* th, double y, double x) {
  switch (dist) {
    case ORC_IID_POISSON: {                       /* Poisson.bl:11-26 */
      double mean = th[0];
      if (law == 0) { if (mean <= 0) return NEG_INF; if (y < 0) return NEG_INF; return y * log(mean); }
      if (law == 1) { if (mean <= 0) return NEG_INF; return -mean; }
      if (y < 0) return NEG_INF;
      return -orc_lngamma(y + 1);                 /* StaticUtils.logFactorial :156-158 */
    }
    case ORC_IID_BINOMIAL: {                      /* Binomial.bl:14-27, k = y, n = x */
      double p = th[0];
      if (law == 0) {
        if (p < 0.0 || p > 1.0) return NEG_INF;
        if (y < 0) return NEG_INF;
        if (x <= 0 || y > x) return NEG_INF;
        return y * log(p) + (x - y) * log(1.0 - p);
      }
      if (y < 0) return NEG_INF;
      if (x <= 0 || y > x) return NEG_INF;
      return orc_lngamma(x + 1) - orc_lngamma(y + 1) - orc_lngamma(x - y + 1);   /* logBinomial :160-162 */
    }
    case ORC_IID_NEGBINOMIAL: {                   /* NegativeBinomial.bl:14-28 */
      double r = th[0], p = th[1];
      if (law == 0) {
        if (r <= 0 || y < 0) return NEG_INF;
        double nn = y + r - 1.0;
        double result = orc_lngamma(nn + 1) - orc_lngamma(y + 1) - orc_lngamma(nn - y + 1);
        if (result != result) return NEG_INF;
        return result;
      }
      if (p <= 0.0 || p >= 1.0) return NEG_INF;
      if (r <= 0 || y < 0) return NEG_INF;
      return y * log(p) + r * log(1.0 - p);
    }
    case ORC_IID_STUDENTT: {                      /* StudentT.bl:13-33 */
      double nu = th[0], mu = th[1], sigma = th[2];
      if (law == 0) return -0.5 * log(M_PI);
      if (law == 1) { if (sigma <= 0.0) return NEG_INF; return -log(sigma); }
      if (law == 2) { if (nu <= 0.0) return NEG_INF; return orc_lngamma((nu + 1.0) / 2.0) - 0.5 * log(nu) - orc_lngamma(nu / 2.0); }
      if (nu <= 0.0 || sigma <= 0.0) return NEG_INF;
      return -((nu + 1.0) / 2.0) * log(1.0 + (1.0 / nu) * (1.0 / (pow(sigma, 2))) * pow((y - mu), 2));
    }
    case ORC_IID_GEOMETRIC: {                     /* Geometric.bl:10-16 */
      double p = th[0];
      if (p <= 0 || p >= 1) return NEG_INF;
      if (y < 0) return NEG_INF;
      return y * log(1 - p) + log(p);
    }
    case ORC_IID_LAPLACE: {                       /* Laplace.bl:12-17 */
      double loc = th[0], scale = th[1];
      if (scale <= 0) return NEG_INF;
      return -log(2 * scale) - fabs(y - loc) / scale;
    }
    case ORC_IID_EXPONENTIAL: return gamma_law(law, 1.0, th[0], y);   /* Exponential.bl:11 -> Gamma(1.0, rate) */
    case ORC_IID_GAMMA: return gamma_law(law, th[0], th[1], y);
    case ORC_IID_WEIBULL: {                       /* Weibull.bl:14-33 */
      double scale = th[0], shape = th[1];
      if (scale <= 0.0) return NEG_INF;
      if (shape <= 0.0) return NEG_INF;
      if (law == 0) return log(shape) - (shape * log(scale));
      if (y <= 0.0) return NEG_INF;
      if (law == 1) return (shape - 1) * log(y);
      return -pow((y / scale), shape);
    }
    case ORC_IID_GUMBEL: {                        /* Gumbel.bl:13-22 */
      double loc = th[0], scale = th[1];
      if (scale <= 0.0) return NEG_INF;
      if (law == 0) return -log(scale);
      return -exp((loc - y) / scale) + ((loc - y) / scale);
    }
    case ORC_IID_LOGISTICDIST: {                  /* Logistic.bl:13-26 */
      double loc = th[0], scale = th[1];
      if (scale <= 0.0) return NEG_INF;
      if (law == 0) return -log(scale);
      if (law == 1) return (loc - y) / scale;
      return -2 * log(1.0 + exp((loc - y) / scale));
    }
    case ORC_IID_HALFSTUDENTT: {                  /* HalfStudentT.bl:12-28 */
      double nu = th[0], sigma = th[1];
      if (law == 0) { if (nu <= 0.0) return NEG_INF; return log(2.0) + orc_lngamma((nu + 1) / 2.0) - orc_lngamma(nu / 2.0) - 0.5 * log(nu * M_PI); }
      if (law == 1) { if (sigma <= 0.0) return NEG_INF; return -log(sigma); }
      if (y < 0.0) return NEG_INF;
      if (sigma <= 0.0) return NEG_INF;
      if (nu <= 0.0) return NEG_INF;
      return -((nu + 1.0) / 2.0) * log(1.0 + 1.0 / nu * pow(y / sigma, 2));
    }
    case ORC_IID_LOGLOGISTIC: {                   /* LogLogistic.bl:14-33 */
      double scale = th[0], shape = th[1];
      if (law == 0) { if (scale <= 0.0) return NEG_INF; if (shape <= 0.0) return NEG_INF; return log(shape) - (shape * log(scale)); }
      if (y < 0.0) return NEG_INF;
      if (scale <= 0.0) return NEG_INF;
      if (shape <= 0.0) return NEG_INF;
      if (law == 1) return shape * log(y) - log(y);
      return -2 * log(1 + pow((y / scale), shape));
    }
    case ORC_IID_GOMPERTZ: {                      /* Gompertz.bl:13-25 */
      double shape = th[0], scale = th[1];
      if (law == 0) { if (shape <= 0.0) return NEG_INF; if (scale <= 0.0) return NEG_INF; return log(shape / scale); }
      if (y < 0.0) return NEG_INF;
      if (scale <= 0.0) return NEG_INF;
      if (shape <= 0.0) return NEG_INF;
      return (y / scale) - (shape * (exp(y / scale) - 1));
    }
    case ORC_IID_YULESIMON: {                     /* YuleSimon.bl:10-16, logBeta = StaticUtils :168-170 */
      double rho = th[0];
      if (rho <= 0) return NEG_INF;
      if (y < 0) return NEG_INF;
      return log(rho) + (orc_lngamma(1.0 + y) + orc_lngamma(rho + 1) - orc_lngamma(1.0 + y + (rho + 1)));
    }
    case ORC_IID_UNIFORM_MAX: {                   /* ContinuousUniform.bl:14-23 with min = 0.0, max latent (F/Doomsday.bl:9) */
      const double mn = 0.0, mx = th[0];
      if (law == 0) { if (mx - mn <= 0.0) return NEG_INF; return -log(mx - mn); }
      if (mn <= y && y <= mx) return 0.0;
      return NEG_INF;
    }
  }
  return NAN;
}

static double factor_eval(const orc_model* m, const Factor* f, const double* re, const int32_t* in) {
  switch (f->kind) {
    case F_CONST: return f->c0;
    case F_NORMAL_F1: {
      double v = re[f->i0];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * log(v);
    }
    case F_NORMAL_F2_LIK: {
      double mean = re[f->i0], v = re[f->i1], x = m->y[f->i2];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * pow(mean - x, 2) / v;
    }
    case F_NORMAL_F2_PRIOR: {
      double mean = f->c0, v = f->c1, x = re[f->i0];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * pow(mean - x, 2) / v;
    }
    case F_GAMMA_A: {
      double shape = f->c0, rate = f->c1, x = re[f->i0];
      if (shape <= 0.0 || rate <= 0) return NEG_INF;
      if (x <= 0.0) return NEG_INF;
      return (shape - 1.0) * log(x * rate);
    }
    case F_GAMMA_B: {
      double rate = f->c1, x = re[f->i0];
      if (rate <= 0) return NEG_INF;
      if (x <= 0.0) return NEG_INF;
      return -x * rate;
    }
    case F_LOGISTIC_ROW: {
      const double* row = m->x + (size_t)f->i0 * m->p;
      double sum = 0.0;
      for (int j = 0; j < m->p; j++) sum += re[j] * row[j];
      double prob = orc_logistic(sum);
      if (prob < 0.0 || prob > 1.0) return NEG_INF;        /* invalidParameter -> -inf (ExponentiatedFactor.java:32) */
      if (!(prob >= 0.0)) return NAN;
      double pick = (m->yi[f->i0] == 1) ? prob : 1.0 - prob; /* fixedSimplex(1-p,p).get(realization) */
      return log(pick);
    }
    case F_MIXTURE_OBS: {
      int K = m->K; double x = m->y[f->i0], acc = 0.0;
      const double *mu = re, *var = re + K, *pi = re + 2 * K;
      for (int k = 0; k < K; k++) {
        double ld;
        if (var[k] <= 0.0) ld = NEG_INF;
        else ld = -0.5 * LOG_2PI + (-0.5 * log(var[k])) + (-0.5 * pow(mu[k] - x, 2) / var[k]);
        acc += pi[k] * exp(ld);
      }
      return log(acc);
    }
    case F_DIRICHLET_A: {
      double sum = 0.0;
      for (int d = 0; d < f->i1; d++) {
        double conc = f->c0;
        if (conc < 0.0) return NEG_INF;
        sum += (conc - 1.0) * log(re[f->i0 + d]);
      }
      return sum;
    }
    case F_ISING_EDGE: {
      int a = in[f->i0], b = in[f->i1];
      if (a < 0 || a > 1 || b < 0 || b > 1) return NEG_INF;
      return f->c0 * (2 * a - 1) * (2 * b - 1);
    }
    case F_BERN_NODE: {
      int s = in[f->i0];
      if (s < 0 || s > 1) return NEG_INF;   /* reference: index out of bounds; never reached by the samplers */
      double prob = f->c0;
      if (prob < 0.0 || prob > 1.0) return NEG_INF;
      return log(s == 1 ? prob : 1.0 - prob);
    }
    case F_IID_LAW: return iid_law(m->K, f->i1, re, m->y[f->i0], m->x[f->i0]);
    case F_LINREG_F1: {
      double sigma = re[m->p], variance = sigma * sigma;
      if (variance <= 0.0) return NEG_INF;
      return -0.5 * log(variance);
    }
    case F_LINREG_F2: {
      const double* row = m->x + (size_t)f->i0 * m->p;
      double mean = 0.0;
      for (int j = 0; j < m->p; j++) mean += re[j] * row[j];
      double sigma = re[m->p], variance = sigma * sigma;
      if (variance <= 0.0) return NEG_INF;
      return -0.5 * pow(mean - m->y[f->i0], 2) / variance;
    }
    case F_POIS_INT: {
      const double th[1] = {f->c0};
      return iid_law(ORC_IID_POISSON, f->i1, th, (double)in[f->i0], 0.0);
    }
    case F_HBINOM_LAT: {
      const double th[1] = {re[2 + f->i0]};
      return iid_law(ORC_IID_BINOMIAL, f->i1, th, (double)m->yi[f->i0], (double)(1 + in[f->i0]));
    }
    case F_HBETA: {
      double alpha = re[0] + f->c0, beta = re[1] + f->c0, x = re[f->i0];   /* c0: 0, or the 0.5 of SimpleHierarchicalModel.bl */
      switch (f->i2) {
        case 0:
          if (x <= 0.0 || x >= 1.0) return NEG_INF;
          if (alpha <= 0.0) return NEG_INF;
          return (alpha - 1.0) * log(x);
        case 1:
          if (x <= 0.0 || x >= 1.0) return NEG_INF;
          if (beta <= 0.0) return NEG_INF;
          return (beta - 1.0) * log1p(-x);
        case 2:
          if (alpha <= 0.0) return NEG_INF;
          if (beta <= 0.0) return NEG_INF;
          return orc_lngamma(alpha + beta);
        case 3: if (alpha <= 0.0) return NEG_INF; return -orc_lngamma(alpha);
        default: if (beta <= 0.0) return NEG_INF; return -orc_lngamma(beta);
      }
    }
    case F_HBINOM: {
      const double th[1] = {re[2 + f->i0]};
      return iid_law(ORC_IID_BINOMIAL, f->i1, th, (double)m->yi[f->i0], (double)m->ntrials[f->i0]);
    }
    case F_BETA_A: {
      double alpha = f->c0, x = re[f->i0];
      if (x <= 0.0 || x >= 1.0) return NEG_INF;
      if (alpha <= 0.0) return NEG_INF;
      return (alpha - 1.0) * log(x);
    }
    case F_BETA_B: {
      double beta = f->c1, x = re[f->i0];
      if (x <= 0.0 || x >= 1.0) return NEG_INF;
      if (beta <= 0.0) return NEG_INF;
      return (beta - 1.0) * log1p(-x);
    }
    case F_UNIF_B: {
      double x = re[f->i0];
      if (f->c0 <= x && x <= f->c1) return 0.0;
      return NEG_INF;
    }
    case F_CAT_LOGP: {   /* Categorical.bl:12-15; an index outside the simplex throws in the reference (never reached: fixed window) */
      int z = in[f->i0];
      if (z < 0 || z >= f->i2) return NEG_INF;
      return log(re[f->i1 + z]);
    }
    case F_MIXIND_F1: {
      int K = m->K, z = in[f->i0];
      if (z < 0 || z >= K) return NEG_INF;
      double v = re[K + z];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * log(v);
    }
    case F_MIXIND_F2: {
      int K = m->K, z = in[f->i0];
      if (z < 0 || z >= K) return NEG_INF;
      double mean = re[z], v = re[K + z], x = m->y[f->i0];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * pow(mean - x, 2) / v;
    }
    case F_DUNIF_B: {
      int x = in[f->i0];
      if ((int)f->c0 <= x && x < (int)f->c1) return 0.0;
      return NEG_INF;
    }
    case F_MANUAL_NORMAL: {   /* Normal::distribution(mu, variance).logDensity(x): the three blocks of Normal.bl:14-24 in order */
      double mean = re[f->i0], v = f->c1, x = f->c0, sum = 0.0;
      sum += -0.5 * log(2.0 * M_PI);
      if (v <= 0.0) return NEG_INF;
      sum += -0.5 * log(v);
      sum += -0.5 * pow(mean - x, 2) / v;
      return sum;
    }
    case F_CP_F2: {
      double mean = f->i0 < in[0] ? re[0] : re[1], v = f->c0, x = m->y[f->i0];
      if (v <= 0.0) return NEG_INF;
      return -0.5 * pow(mean - x, 2) / v;
    }
    case F_BETABINOM: {
      double alpha = re[f->i1], beta = re[f->i2];
      double x = (double)m->yi[f->i0], n = (double)m->ntrials[f->i0];
      if (alpha <= 0.0 || beta <= 0.0) return NEG_INF;
      if (x < 0.0) return NEG_INF;
      if (n <= 0.0 || x > n) return NEG_INF;
      return orc_lngamma(n + 1) + orc_lngamma(x + alpha) + orc_lngamma(n - x + beta) +
             orc_lngamma(alpha + beta) - orc_lngamma(x + 1) - orc_lngamma(n - x + 1) -
             orc_lngamma(n + alpha + beta) - orc_lngamma(alpha) - orc_lngamma(beta);
    }
  }
  return NAN;
}

/* ------------------------------------------------------------------ */
/* model builders                                                       */
/* ------------------------------------------------------------------ */
typedef struct { int32_t* v; int n, cap; } IVec;
static void iv_push(IVec* a, int x) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 16; a->v = (int32_t*)realloc(a->v, sizeof(int32_t) * a->cap); }
  a->v[a->n++] = x;
}

typedef struct {
  orc_model* m;
  int fcap;
  IVec* s_fixed; IVec* s_ann; IVec* s_other;  /* per sampler */
} Builder;

static int b_factor(Builder* b, int kind, int annealed, int i0, int i1, int i2, double c0, double c1) {
  orc_model* m = b->m;
  if (m->n_factors == b->fcap) { b->fcap = b->fcap ? b->fcap * 2 : 64; m->factors = (Factor*)realloc(m->factors, sizeof(Factor) * b->fcap); }
  Factor* f = &m->factors[m->n_factors];
  f->kind = kind; f->annealed = annealed; f->i0 = i0; f->i1 = i1; f->i2 = i2; f->c0 = c0; f->c1 = c1;
  return m->n_factors++;
}
static void b_connect(Builder* b, int sampler, int factor) {
  const int a = b->m->factors[factor].annealed;
  iv_push(a == 2 ? &b->s_other[sampler] : (a ? &b->s_ann[sampler] : &b->s_fixed[sampler]), factor);
}
static Builder b_begin(int family, int n_reals, int n_ints, int n_samplers) {
  Builder b; memset(&b, 0, sizeof b);
  b.m = (orc_model*)calloc(1, sizeof(orc_model));
  b.m->family = family; b.m->n_reals = n_reals; b.m->n_ints = n_ints; b.m->n_samplers = n_samplers;
  b.m->samplers = (Sampler*)calloc(n_samplers > 0 ? n_samplers : 1, sizeof(Sampler));
  b.s_fixed = (IVec*)calloc(n_samplers > 0 ? n_samplers : 1, sizeof(IVec));
  b.s_ann = (IVec*)calloc(n_samplers > 0 ? n_samplers : 1, sizeof(IVec));
  b.s_other = (IVec*)calloc(n_samplers > 0 ? n_samplers : 1, sizeof(IVec));
  b.m->init_reals = (double*)calloc(n_reals > 0 ? n_reals : 1, sizeof(double));
  b.m->init_ints = (int32_t*)calloc(n_ints > 0 ? n_ints : 1, sizeof(int32_t));
  return b;
}
static orc_model* b_end(Builder* b) {
  orc_model* m = b->m;
  IVec fx = {0, 0, 0}, an = {0, 0, 0}, ot = {0, 0, 0};
  for (int i = 0; i < m->n_factors; i++) iv_push(m->factors[i].annealed == 2 ? &ot : (m->factors[i].annealed ? &an : &fx), i);
  m->fixed_idx = fx.v; m->n_fixed_idx = fx.n; m->ann_idx = an.v; m->n_ann_idx = an.n; m->other_idx = ot.v; m->n_other_idx = ot.n;
  for (int s = 0; s < m->n_samplers; s++) {
    m->samplers[s].fixed_idx = b->s_fixed[s].v; m->samplers[s].n_fixed = b->s_fixed[s].n;
    m->samplers[s].ann_idx = b->s_ann[s].v; m->samplers[s].n_ann = b->s_ann[s].n;
    m->samplers[s].other_idx = b->s_other[s].v; m->samplers[s].n_other = b->s_other[s].n;
  }
  free(b->s_fixed); free(b->s_ann); free(b->s_other);
  return m;
}
static void b_forward(orc_model* m, int kind, int var, int K, double a, double bb) {
  m->forward = (Forward*)realloc(m->forward, sizeof(Forward) * (m->n_forward + 1));
  Forward* f = &m->forward[m->n_forward++];
  f->kind = kind; f->var = var; f->K = K; f->a = a; f->b = bb;
}
/* `x ~ Normal(mean, var)` with constant parameters: three factors (Normal.bl:14-25) */
static void b_normal_prior(Builder* b, int sampler, int var, double mean, double variance) {
  b_factor(b, F_CONST, 0, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
  b_factor(b, F_CONST, 0, 0, 0, 0, variance <= 0.0 ? NEG_INF : -0.5 * log(variance), 0);
  b_connect(b, sampler, b_factor(b, F_NORMAL_F2_PRIOR, 0, var, 0, 0, mean, variance));
  b_forward(b->m, FW_NORMAL, var, 0, mean, variance);
}
/* `x ~ Gamma(shape, rate)` with constant parameters: four factors (Gamma.bl:14-31) */
static void b_gamma_prior(Builder* b, int sampler, int var, double shape, double rate, int as_exponential) {
  b_connect(b, sampler, b_factor(b, F_GAMMA_A, 0, var, 0, 0, shape, rate));
  b_connect(b, sampler, b_factor(b, F_GAMMA_B, 0, var, 0, 0, shape, rate));
  b_factor(b, F_CONST, 0, 0, 0, 0, shape <= 0.0 ? NEG_INF : -orc_lngamma(shape), 0);
  b_factor(b, F_CONST, 0, 0, 0, 0, rate <= 0.0 ? NEG_INF : log(rate), 0);
  b_forward(b->m, as_exponential ? FW_EXPONENTIAL : FW_GAMMA, var, 0, shape, rate);
}

/* `x ~ Beta(alpha, beta)` with constant parameters: five factors (Beta.bl:14-41) */
static void b_beta_prior(Builder* b, int sampler, int var, double alpha, double beta) {
  b_connect(b, sampler, b_factor(b, F_BETA_A, 0, var, 0, 0, alpha, beta));
  b_connect(b, sampler, b_factor(b, F_BETA_B, 0, var, 0, 0, alpha, beta));
  b_factor(b, F_CONST, 0, 0, 0, 0, (alpha <= 0.0 || beta <= 0.0) ? NEG_INF : orc_lngamma(alpha + beta), 0);
  b_factor(b, F_CONST, 0, 0, 0, 0, alpha <= 0.0 ? NEG_INF : -orc_lngamma(alpha), 0);
  b_factor(b, F_CONST, 0, 0, 0, 0, beta <= 0.0 ? NEG_INF : -orc_lngamma(beta), 0);
  b_forward(b->m, FW_BETA, var, 0, alpha, beta);
}
/* `x ~ ContinuousUniform(min, max)`: two factors (ContinuousUniform.bl:12-19) */
static void b_uniform_prior(Builder* b, int sampler, int var, double mn, double mx) {
  b_factor(b, F_CONST, 0, 0, 0, 0, (mx - mn <= 0.0) ? NEG_INF : -log(mx - mn), 0);
  b_connect(b, sampler, b_factor(b, F_UNIF_B, 0, var, 0, 0, mn, mx));
  b_forward(b->m, FW_UNIFORM, var, 0, mn, mx);
}

/* family 6: y_i | theta ~ Dist(theta) i.i.d., all observed (annealed); theta_j ~ prior_j.  One real slice sampler
 * per parameter, in declaration order.  Every law of the distribution is one factor PER OBSERVATION, also the
 * laws that read only parameters (Poisson's -mean, Student's -log sigma ...): that is what instantiating
 * `y.get(i) | theta ~ Dist(theta)` n times produces, and it is what nOutOfSupport counts. */
orc_model* orc_model_iid(int dist, const double* y, const double* trials, int n, const int32_t* prior_kind,
                         const double* prior_a, const double* prior_b) {
  if (dist < ORC_IID_POISSON || dist > ORC_IID_UNIFORM_MAX) return NULL;
  const int P = IID_NPARAMS[dist];
  Builder b = b_begin(ORC_FAM_IID, P, 0, P);
  orc_model* m = b.m;
  m->n_obs = n; m->K = dist;
  m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  m->x = (double*)calloc(n > 0 ? n : 1, sizeof(double));
  if (trials) memcpy(m->x, trials, sizeof(double) * n);
  for (int j = 0; j < P; j++) {
    m->samplers[j].type = S_REAL_SLICE; m->samplers[j].var = j;
    switch (prior_kind[j]) {
      case ORC_PRIOR_NORMAL: b_normal_prior(&b, j, j, prior_a[j], prior_b[j]); break;
      case ORC_PRIOR_GAMMA: b_gamma_prior(&b, j, j, prior_a[j], prior_b[j], 0); break;
      case ORC_PRIOR_EXPONENTIAL: b_gamma_prior(&b, j, j, 1.0, prior_a[j], 1); break;
      case ORC_PRIOR_BETA: b_beta_prior(&b, j, j, prior_a[j], prior_b[j]); break;
      default: b_uniform_prior(&b, j, j, prior_a[j], prior_b[j]); break;
    }
  }
  for (int i = 0; i < n; i++)
    for (int l = 0; l < IID_NLAWS[dist]; l++) {
      int fi = b_factor(&b, F_IID_LAW, 1, i, l, 0, 0, 0);
      for (int j = 0; j < P; j++) if (IID_MASK[dist][l] & (1 << j)) b_connect(&b, j, fi);
    }
  return b_end(&b);
}

/* family 8: F/LinRegression.bl generalised to p covariates:
 *   w_j ~ Normal(0, v0), sigma ~ ContinuousUniform(0, smax), y_i | w, sigma ~ Normal(sum_j w_j x_ij, sigma * sigma).
 * (the fixture is p = 2 with columns (x_i, 1): beta * x_i + alpha.)  reals = [w_0..w_{p-1}, sigma]. */
orc_model* orc_model_linreg(const double* x, const double* y, int n, int p, double v0, double smax) {
  Builder b = b_begin(ORC_FAM_LINREG, p + 1, 0, p + 1);
  orc_model* m = b.m;
  m->n_obs = n; m->p = p;
  m->x = (double*)malloc(sizeof(double) * (size_t)n * p + 8); memcpy(m->x, x, sizeof(double) * (size_t)n * p);
  m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  for (int j = 0; j <= p; j++) { m->samplers[j].type = S_REAL_SLICE; m->samplers[j].var = j; }
  for (int j = 0; j < p; j++) b_normal_prior(&b, j, j, 0.0, v0);
  b_uniform_prior(&b, p, p, 0.0, smax);
  for (int i = 0; i < n; i++) {
    b_factor(&b, F_CONST, 1, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
    b_connect(&b, p, b_factor(&b, F_LINREG_F1, 1, i, 0, 0, 0, 0));
    int f2 = b_factor(&b, F_LINREG_F2, 1, i, 0, 0, 0, 0);
    for (int j = 0; j <= p; j++) b_connect(&b, j, f2);
  }
  return b_end(&b);
}

/* family 7: F/HierarchicalModel.bl (the rocket-launch example of the Blang documentation), not marginalised:
 *   a ~ Exponential(rate_a), b ~ Exponential(rate_b), p_g | a, b ~ Beta(a, b), failures_g | p_g ~ Binomial(launches_g, p_g).
 * reals = [a, b, p_0 .. p_{G-1}], one real slice sampler each.  The five Beta blocks per group are FIXED factors
 * (p_g is latent), the two Binomial blocks per group are annealed (failures are observed). */
orc_model* orc_model_hierarchical(const int32_t* failures, const int32_t* launches, int G, double rate_a, double rate_b) {
  Builder b = b_begin(ORC_FAM_HIERARCHICAL, G + 2, 0, G + 2);
  orc_model* m = b.m; m->n_obs = G;
  m->yi = (int32_t*)malloc(sizeof(int32_t) * (G > 0 ? G : 1)); memcpy(m->yi, failures, sizeof(int32_t) * G);
  m->ntrials = (int32_t*)malloc(sizeof(int32_t) * (G > 0 ? G : 1)); memcpy(m->ntrials, launches, sizeof(int32_t) * G);
  for (int s = 0; s < G + 2; s++) { m->samplers[s].type = S_REAL_SLICE; m->samplers[s].var = s; }
  b_gamma_prior(&b, 0, 0, 1.0, rate_a, 1);
  b_gamma_prior(&b, 1, 1, 1.0, rate_b, 1);
  for (int g = 0; g < G; g++) {
    int fa = b_factor(&b, F_HBETA, 0, 2 + g, 0, 0, 0, 0);   b_connect(&b, 0, fa); b_connect(&b, 2 + g, fa);
    int fb = b_factor(&b, F_HBETA, 0, 2 + g, 0, 1, 0, 0);   b_connect(&b, 1, fb); b_connect(&b, 2 + g, fb);
    int fc = b_factor(&b, F_HBETA, 0, 2 + g, 0, 2, 0, 0);   b_connect(&b, 0, fc); b_connect(&b, 1, fc);
    b_connect(&b, 0, b_factor(&b, F_HBETA, 0, 2 + g, 0, 3, 0, 0));
    b_connect(&b, 1, b_factor(&b, F_HBETA, 0, 2 + g, 0, 4, 0, 0));
    b_forward(m, FW_BETA_LATENT, 2 + g, 0, 0, 0);
    b_connect(&b, 2 + g, b_factor(&b, F_HBINOM, 1, g, 0, 0, 0, 0));
    b_factor(&b, F_HBINOM, 1, g, 1, 0, 0, 0);
  }
  m->init_reals[0] = 1.0; m->init_reals[1] = 1.0;
  for (int g = 0; g < G; g++) m->init_reals[2 + g] = 0.5;
  return b_end(&b);
}

/* family 9: F/SimpleHierarchicalModel.bl -- family 7 with LATENT launch counts (IntSliceSampler with stepping out):
 *   a, b ~ Exponential(rate); p_g | a, b ~ Beta(a + 0.5, b + 0.5); L_g ~ Poisson(mean); failures_g ~ Binomial(1 + L_g, p_g).
 * reals = [a, b, p_0..p_{G-1}], ints = [L_0..L_{G-1}]; samplers: G + 2 real slice, then G int slice. */
orc_model* orc_model_rockets_latent(const int32_t* failures, int G, double rate_a, double rate_b, double pois_mean) {
  Builder b = b_begin(ORC_FAM_ROCKETS_LATENT, G + 2, G, 2 * G + 2);
  orc_model* m = b.m; m->n_obs = G;
  m->yi = (int32_t*)malloc(sizeof(int32_t) * (G > 0 ? G : 1)); memcpy(m->yi, failures, sizeof(int32_t) * G);
  for (int s = 0; s < G + 2; s++) { m->samplers[s].type = S_REAL_SLICE; m->samplers[s].var = s; }
  for (int g = 0; g < G; g++) { m->samplers[G + 2 + g].type = S_INT_SLICE; m->samplers[G + 2 + g].var = g; }
  b_gamma_prior(&b, 0, 0, 1.0, rate_a, 1);
  b_gamma_prior(&b, 1, 1, 1.0, rate_b, 1);
  for (int g = 0; g < G; g++) {
    const int sp = 2 + g, sl = G + 2 + g;
    int fa = b_factor(&b, F_HBETA, 0, 2 + g, 0, 0, 0.5, 0);   b_connect(&b, 0, fa); b_connect(&b, sp, fa);
    int fb = b_factor(&b, F_HBETA, 0, 2 + g, 0, 1, 0.5, 0);   b_connect(&b, 1, fb); b_connect(&b, sp, fb);
    int fc = b_factor(&b, F_HBETA, 0, 2 + g, 0, 2, 0.5, 0);   b_connect(&b, 0, fc); b_connect(&b, 1, fc);
    b_connect(&b, 0, b_factor(&b, F_HBETA, 0, 2 + g, 0, 3, 0.5, 0));
    b_connect(&b, 1, b_factor(&b, F_HBETA, 0, 2 + g, 0, 4, 0.5, 0));
    b_forward(m, FW_BETA_LATENT, 2 + g, 0, 0.5, 0);
    b_connect(&b, sl, b_factor(&b, F_POIS_INT, 0, g, 0, 0, pois_mean, 0));
    b_factor(&b, F_POIS_INT, 0, g, 1, 0, pois_mean, 0);
    b_connect(&b, sl, b_factor(&b, F_POIS_INT, 0, g, 2, 0, pois_mean, 0));
    b_forward(m, FW_POISSON_INT, g, 0, pois_mean, 0);
    int l0 = b_factor(&b, F_HBINOM_LAT, 1, g, 0, 0, 0, 0);     b_connect(&b, sp, l0); b_connect(&b, sl, l0);
    b_connect(&b, sl, b_factor(&b, F_HBINOM_LAT, 1, g, 1, 0, 0, 0));
  }
  m->init_reals[0] = 1.0; m->init_reals[1] = 1.0;
  for (int g = 0; g < G; g++) { m->init_reals[2 + g] = 0.5; m->init_ints[g] = 0; }
  return b_end(&b);
}

orc_model* orc_model_normal(const double* y, int n, double m0, double v0, double rate) {
  Builder b = b_begin(ORC_FAM_NORMAL, 2, 0, 2);
  orc_model* m = b.m;
  m->n_obs = n; m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  m->samplers[0].type = S_REAL_SLICE; m->samplers[0].var = 0;
  m->samplers[1].type = S_REAL_SLICE; m->samplers[1].var = 1;
  b_normal_prior(&b, 0, 0, m0, v0);
  b_gamma_prior(&b, 1, 1, 1.0, rate, 1);          /* Exponential.bl:11 -> Gamma(1.0, rate) */
  for (int i = 0; i < n; i++) {                   /* y_i | mu, s2 ~ Normal(mu, s2), all observed => annealed */
    b_factor(&b, F_CONST, 1, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
    b_connect(&b, 1, b_factor(&b, F_NORMAL_F1, 1, 1, 0, 0, 0, 0));
    int f2 = b_factor(&b, F_NORMAL_F2_LIK, 1, 0, 1, i, 0, 0);
    b_connect(&b, 0, f2); b_connect(&b, 1, f2);
  }
  m->init_reals[0] = 0.0; m->init_reals[1] = 1.0;
  return b_end(&b);
}

orc_model* orc_model_logistic(const double* x, const int32_t* y, int n, int p, double v0) {
  Builder b = b_begin(ORC_FAM_LOGISTIC, p, 0, p);
  orc_model* m = b.m;
  m->n_obs = n; m->p = p;
  m->x = (double*)malloc(sizeof(double) * (size_t)n * p + 8); memcpy(m->x, x, sizeof(double) * (size_t)n * p);
  m->yi = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1)); memcpy(m->yi, y, sizeof(int32_t) * n);
  for (int j = 0; j < p; j++) {
    m->samplers[j].type = S_REAL_SLICE; m->samplers[j].var = j;
    b_normal_prior(&b, j, j, 0.0, v0);
  }
  for (int i = 0; i < n; i++) {
    int f = b_factor(&b, F_LOGISTIC_ROW, 1, i, 0, 0, 0, 0);
    for (int j = 0; j < p; j++) b_connect(&b, j, f);
  }
  return b_end(&b);
}

/* reals layout: mu[K], var[K], pi[K]; samplers: k<K mu_k, K<=k<2K var_k, 2K simplex */
orc_model* orc_model_mixture(const double* y, int n, int K, double m0, double v0,
                             double gshape, double grate, double conc) {
  Builder b = b_begin(ORC_FAM_MIXTURE, 3 * K, 0, 2 * K + 1);
  orc_model* m = b.m;
  m->n_obs = n; m->K = K;
  m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  /* F/MixtureModel.bl:14-20 order: pi first, then per component mean, variance */
  m->samplers[2 * K].type = S_SIMPLEX; m->samplers[2 * K].var = 2 * K; m->samplers[2 * K].K = K;
  b_connect(&b, 2 * K, b_factor(&b, F_DIRICHLET_A, 0, 2 * K, K, 0, conc, 0));
  b_factor(&b, F_CONST, 0, 0, 0, 0, conc < 0 ? NEG_INF : (-K * orc_lngamma(conc) + orc_lngamma(K * conc)), 0); /* Dirichlet.bl:23-30 */
  b_forward(m, FW_DIRICHLET_SYM, 2 * K, K, conc, 0);
  for (int k = 0; k < K; k++) {
    m->samplers[k].type = S_REAL_SLICE; m->samplers[k].var = k;
    m->samplers[K + k].type = S_REAL_SLICE; m->samplers[K + k].var = K + k;
    b_normal_prior(&b, k, k, m0, v0);
    b_gamma_prior(&b, K + k, K + k, gshape, grate, 0);
  }
  for (int i = 0; i < n; i++) {
    int f = b_factor(&b, F_MIXTURE_OBS, 1, i, 0, 0, 0, 0);
    for (int s = 0; s < 2 * K + 1; s++) b_connect(&b, s, f);
  }
  for (int k = 0; k < K; k++) { m->init_reals[k] = 0.0; m->init_reals[K + k] = 1.0; m->init_reals[2 * K + k] = 1.0 / K; }
  return b_end(&b);
}

/* family 10: blang.examples.MixtureModel (F/MixtureModel.bl) with K components and the cluster indicators NOT
 * marginalised: pi ~ Dirichlet(conc), mu_k ~ Normal(m0, v0), var_k ~ Gamma(shape, rate),
 * z_i | pi ~ Categorical(pi) (latent: CategoricalSampler = IntSliceSampler with the fixed window [0, K),
 * R/mcmc/CategoricalSampler.xtend:24-28; with K > 10 the window test of IntSliceSampler.accept :115-132 runs),
 * y_i | z_i ~ Normal(means.get(z_i), variances.get(z_i)) (observed: annealed).
 * Connections follow the DSL's static dependencies: the Normal's mean / variance closures read the whole lists
 * `means` / `variances` and the indicator, so every mean's sampler sees every observation's f2, every variance's
 * sampler every f1 and f2, and indicator i's sampler its own Categorical factor, f1_i and f2_i; the simplex sampler
 * sees the Dirichlet factor and all n Categorical factors.
 * reals = [mu_0..mu_{K-1}, var_0.., pi_0..]; ints = z_0..z_{n-1}; samplers: mu, var, pi (simplex), then z_i. */
orc_model* orc_model_mixture_indicators(const double* y, int n, int K, double m0, double v0, double gshape, double grate, double conc) {
  Builder b = b_begin(ORC_FAM_MIXTURE_INDICATORS, 3 * K, n, 2 * K + 1 + n);
  orc_model* m = b.m;
  m->n_obs = n; m->K = K;
  m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  m->samplers[2 * K].type = S_SIMPLEX; m->samplers[2 * K].var = 2 * K; m->samplers[2 * K].K = K;
  b_connect(&b, 2 * K, b_factor(&b, F_DIRICHLET_A, 0, 2 * K, K, 0, conc, 0));
  b_factor(&b, F_CONST, 0, 0, 0, 0, conc < 0 ? NEG_INF : (-K * orc_lngamma(conc) + orc_lngamma(K * conc)), 0);
  b_forward(m, FW_DIRICHLET_SYM, 2 * K, K, conc, 0);
  for (int k = 0; k < K; k++) {
    m->samplers[k].type = S_REAL_SLICE; m->samplers[k].var = k;
    m->samplers[K + k].type = S_REAL_SLICE; m->samplers[K + k].var = K + k;
    b_normal_prior(&b, k, k, m0, v0);
    b_gamma_prior(&b, K + k, K + k, gshape, grate, 0);
  }
  for (int i = 0; i < n; i++) {
    const int sz = 2 * K + 1 + i;
    m->samplers[sz].type = S_CATEGORICAL; m->samplers[sz].var = i; m->samplers[sz].K = K;
    int fc = b_factor(&b, F_CAT_LOGP, 0, i, 2 * K, K, 0, 0);
    b_connect(&b, 2 * K, fc); b_connect(&b, sz, fc);
    b_forward(m, FW_CATEGORICAL_INT, i, K, 2 * K, 0);
    b_factor(&b, F_CONST, 1, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
    int f1 = b_factor(&b, F_MIXIND_F1, 1, i, 0, 0, 0, 0);
    int f2 = b_factor(&b, F_MIXIND_F2, 1, i, 0, 0, 0, 0);
    for (int k = 0; k < K; k++) { b_connect(&b, k, f2); b_connect(&b, K + k, f1); b_connect(&b, K + k, f2); }
    b_connect(&b, sz, f1); b_connect(&b, sz, f2);
  }
  for (int k = 0; k < K; k++) { m->init_reals[k] = 0.0; m->init_reals[K + k] = 1.0; m->init_reals[2 * K + k] = 1.0 / K; }
  return b_end(&b);
}

/* family 11: a change-point model, the smallest model with a DiscreteUniform latent (R/distributions/DiscreteUniform.bl,
 * sampled by UniformSampler = IntSliceSampler with the fixed window [minInclusive, maxExclusive), UniformSampler.xtend:21-27):
 *   mu_0, mu_1 ~ Normal(m0, v0); c ~ DiscreteUniform(0, n + 1); y_i | mu, c ~ Normal(i < c ? mu_0 : mu_1, s2) with s2 a constant.
 * The Normal's mean closure reads mu_0, mu_1 and c, so all three samplers see every f2; f0 and f1 depend on nothing latent.
 * reals = [mu_0, mu_1]; ints = [c]; samplers: mu_0, mu_1, c. */
orc_model* orc_model_changepoint(const double* y, int n, double m0, double v0, double s2) {
  Builder b = b_begin(ORC_FAM_CHANGEPOINT, 2, 1, 3);
  orc_model* m = b.m;
  m->n_obs = n;
  m->y = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->y, y, sizeof(double) * n);
  for (int k = 0; k < 2; k++) { m->samplers[k].type = S_REAL_SLICE; m->samplers[k].var = k; b_normal_prior(&b, k, k, m0, v0); }
  m->samplers[2].type = S_UNIFORM; m->samplers[2].var = 0; m->samplers[2].lo = 0; m->samplers[2].K = n + 1;
  b_factor(&b, F_CONST, 0, 0, 0, 0, -log((double)(n + 1)), 0);                       /* DiscreteUniform.bl:12-15 */
  b_connect(&b, 2, b_factor(&b, F_DUNIF_B, 0, 0, 0, 0, 0.0, (double)(n + 1)));      /* :16-20 */
  b_forward(m, FW_DISCRETE_UNIFORM_INT, 0, 0, 0.0, (double)(n + 1));
  for (int i = 0; i < n; i++) {
    b_factor(&b, F_CONST, 1, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
    b_factor(&b, F_CONST, 1, 0, 0, 0, s2 <= 0.0 ? NEG_INF : -0.5 * log(s2), 0);
    int f2 = b_factor(&b, F_CP_F2, 1, i, 0, 0, s2, 0);
    b_connect(&b, 0, f2); b_connect(&b, 1, f2); b_connect(&b, 2, f2);
  }
  m->init_reals[0] = 0.0; m->init_reals[1] = 0.0; m->init_ints[0] = 0;
  return b_end(&b);
}

/* family 12: F/CustomAnnealRef.bl (manual = 0) and F/CustomAnnealTest.bl (manual = 1), the pair behind
 * T/TestEndToEnd.xtend:169-186 ("automatic and manual annealing give the same beta = 1 logDensity to 1e-10"):
 *   mu ~ Normal(m0, v0);  x (fixed, observed) | mu ~ Normal(mu, s2)                      -- three exponentiated factors
 *   mu ~ Normal(m0, v0);  | x, mu, beta = new AnnealingParameter ~ LogPotential(Normal::distribution(mu, s2).logDensity(x) * beta)
 *                                                                                         -- one "other annealed" factor */
orc_model* orc_model_anneal_pair(double x, double m0, double v0, double s2, int manual) {
  Builder b = b_begin(ORC_FAM_ANNEAL_PAIR, 1, 0, 1);
  orc_model* m = b.m; m->n_obs = 1;
  m->y = (double*)malloc(sizeof(double)); m->y[0] = x;
  m->samplers[0].type = S_REAL_SLICE; m->samplers[0].var = 0;
  b_normal_prior(&b, 0, 0, m0, v0);
  if (manual) b_connect(&b, 0, b_factor(&b, F_MANUAL_NORMAL, 2, 0, 0, 0, x, s2));
  else {
    b_factor(&b, F_CONST, 1, 0, 0, 0, -0.5 * log(2.0 * M_PI), 0);
    b_factor(&b, F_CONST, 1, 0, 0, 0, s2 <= 0.0 ? NEG_INF : -0.5 * log(s2), 0);
    b_connect(&b, 0, b_factor(&b, F_NORMAL_F2_PRIOR + 0, 1, 0, 0, 0, x, s2));   /* Normal.bl:21-24 with mean = mu: (mu - x)^2 is symmetric */
  }
  m->init_reals[0] = 0.0;
  return b_end(&b);
}

/* F/Ising.bl + F/Functions.xtend:8-23 (edge order preserved) */
orc_model* orc_model_ising(int N, double coupling, double moment) {
  Builder b = b_begin(ORC_FAM_ISING, 0, N * N, N * N);
  orc_model* m = b.m; m->N = N;
  double p = orc_logistic(-2.0 * moment);
  for (int v = 0; v < N * N; v++) { m->samplers[v].type = S_CATEGORICAL; m->samplers[v].var = v; m->samplers[v].K = 2; }
#define EDGE(u, v) do { int f_ = b_factor(&b, F_ISING_EDGE, 1, (u), (v), 0, coupling, 0); b_connect(&b, (u), f_); if ((v) != (u)) b_connect(&b, (v), f_); } while (0)
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N - 1; j++) EDGE(N * i + j, N * i + j + 1);
    EDGE(N * i, N * i + N - 1);
  }
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < N - 1; i++) EDGE(N * i + j, N * (i + 1) + j);
    EDGE(j, N * (N - 1) + j);
  }
#undef EDGE
  for (int v = 0; v < N * N; v++) {
    b_connect(&b, v, b_factor(&b, F_BERN_NODE, 0, v, 0, 0, p, 0));
    b_forward(m, FW_BERNOULLI_INT, v, 0, p, 0);
  }
  return b_end(&b);
}

orc_model* orc_model_betabinomial(const int32_t* y, const int32_t* nt, int n, double rate) {
  Builder b = b_begin(ORC_FAM_BETABINOMIAL, 2, 0, 2);
  orc_model* m = b.m; m->n_obs = n;
  m->yi = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1)); memcpy(m->yi, y, sizeof(int32_t) * n);
  m->ntrials = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1)); memcpy(m->ntrials, nt, sizeof(int32_t) * n);
  for (int s = 0; s < 2; s++) { m->samplers[s].type = S_REAL_SLICE; m->samplers[s].var = s; b_gamma_prior(&b, s, s, 1.0, rate, 1); }
  for (int g = 0; g < n; g++) {
    int f = b_factor(&b, F_BETABINOM, 1, g, 0, 1, 0, 0);
    b_connect(&b, 0, f); b_connect(&b, 1, f);
  }
  m->init_reals[0] = 1.0; m->init_reals[1] = 1.0;
  return b_end(&b);
}

void orc_model_free(orc_model* m) {
  if (!m) return;
  for (int s = 0; s < m->n_samplers; s++) { free(m->samplers[s].fixed_idx); free(m->samplers[s].ann_idx); free(m->samplers[s].other_idx); }
  free(m->samplers); free(m->factors); free(m->fixed_idx); free(m->ann_idx); free(m->other_idx); free(m->forward);
  free(m->y); free(m->x); free(m->yi); free(m->ntrials); free(m->init_reals); free(m->init_ints);
  free(m);
}
int orc_model_n_reals(const orc_model* m) { return m->n_reals; }
int orc_model_n_ints(const orc_model* m) { return m->n_ints; }
int orc_model_n_samplers(const orc_model* m) { return m->n_samplers; }
int orc_model_n_factors(const orc_model* m) { return m->n_factors; }
double orc_model_factor_value(const orc_model* m, int f, const double* re, const int32_t* in) {
  return factor_eval(m, &m->factors[f], re, in);
}
int orc_model_factor_is_annealed(const orc_model* m, int f) { return m->factors[f].annealed; }

/* ------------------------------------------------------------------ */
/* ExponentiatedFactor (R/mcmc/internals/ExponentiatedFactor.java)      */
/* ------------------------------------------------------------------ */
static double annealed_minus_infinity(double exponent) {          /* :82-87 */
  if (exponent >= 1.0) return NEG_INF;
  return -exponent * 1E100;
}

/* ------------------------------------------------------------------ */
/* SampledModel (R/runtime/SampledModel.java)                           */
/* ------------------------------------------------------------------ */
typedef struct {
  const orc_model* m;
  double* reals; int32_t* ints;
  double exponent;                  /* _annealingExponent */
  int anneal_support;
  double* caches;
  double sum_pre, sum_fixed; int n_oos; int oos_detected;
  int32_t* order; int cur_pos;      /* currentSamplingOrder / currentPosition :85-86 */
  double cumulative_n_scans;        /* :256 */
  uint64_t n_evals, n_factor_evals;
  int error;                        /* NaN factor / slice diverged */
} SampledModel;

/* enclosedLogDensity :22-33 (NaN -> error flag; InvalidParameter is folded into the factor bodies) */
static double sm_enclosed(SampledModel* s, int fi) {
  double r = factor_eval(s->m, &s->m->factors[fi], s->reals, s->ints);
  if (s->m->factors[fi].annealed == 2) r = r * s->exponent;   /* `dist.logDensity(x) * beta`, beta = the AnnealingParameter */
  s->n_factor_evals++;
  if (isnan(r)) { s->error |= 1; return NEG_INF; }
  return r;
}
/* ExponentiatedFactor.logDensity :70-80; exponent==null (fixed factor) means 1.0 */
static double sm_factor_log_density(SampledModel* s, int fi) {
  double exp_value = s->m->factors[fi].annealed == 1 ? s->exponent : 1.0;   /* other-annealed: exponent == null, :70-80 */
  if (exp_value == 0.0) return 0.0;
  double e = sm_enclosed(s, fi);
  if (e == NEG_INF && s->anneal_support) return annealed_minus_infinity(exp_value);
  return exp_value * e;
}
/* updateAll :340-364 */
static void sm_update_all(SampledModel* s) {
  const orc_model* m = s->m;
  s->sum_fixed = 0.0;
  for (int k = 0; k < m->n_fixed_idx; k++) {
    int fi = m->fixed_idx[k];
    double c = sm_factor_log_density(s, fi);
    s->sum_fixed += c; s->caches[fi] = c;
  }
  s->sum_pre = 0.0; s->n_oos = 0;
  for (int k = 0; k < m->n_ann_idx; k++) {
    int fi = m->ann_idx[k];
    double c = sm_enclosed(s, fi);
    s->caches[fi] = c;
    if (c == NEG_INF) s->n_oos++; else s->sum_pre += c;
  }
  if (s->n_oos > 0 && s->anneal_support) s->oos_detected = 1;
}
/* update :366-405 */
static void sm_update(SampledModel* s, int sampler) {
  if (s->sum_pre == NEG_INF || s->sum_fixed == NEG_INF) { sm_update_all(s); return; }
  const Sampler* sp = &s->m->samplers[sampler];
  for (int k = 0; k < sp->n_fixed; k++) {
    int fi = sp->fixed_idx[k];
    double c = sm_factor_log_density(s, fi);
    s->sum_fixed += c - s->caches[fi];
    s->caches[fi] = c;
  }
  for (int k = 0; k < sp->n_ann; k++) {
    int fi = sp->ann_idx[k];
    double old = s->caches[fi];
    if (old == NEG_INF) s->n_oos--; else s->sum_pre -= old;
    double c = sm_enclosed(s, fi);
    s->caches[fi] = c;
    if (c == NEG_INF) { if (s->anneal_support) s->oos_detected = 1; s->n_oos++; }
    else s->sum_pre += c;
  }
}
/* sumOtherAnnealed() :219-225: custom-annealed factors are evaluated explicitly every time */
static double sm_sum_other_annealed(const SampledModel* s) {
  double sum = 0.0;
  for (int k = 0; k < s->m->n_other_idx; k++) sum += sm_factor_log_density((SampledModel*)s, s->m->other_idx[k]);
  return sum;
}
/* logDensity() :161-176 */
static double sm_log_density(const SampledModel* s) {
  double e = s->exponent;
  if (!s->anneal_support && s->n_oos > 0) return NEG_INF;
  return sm_sum_other_annealed(s) + s->sum_fixed + e * s->sum_pre +
         (s->n_oos == 0 ? 0.0 : s->n_oos * annealed_minus_infinity(e));
}
/* logDensity(beta) :190-197 */
static double sm_log_density_at(SampledModel* s, double beta) {
  double prev = s->exponent; s->exponent = beta;
  double r = sm_log_density(s);
  s->exponent = prev; return r;
}

static SampledModel* sm_new(const orc_model* m, int anneal_support) {
  SampledModel* s = (SampledModel*)calloc(1, sizeof(SampledModel));
  s->m = m; s->anneal_support = anneal_support; s->exponent = 1.0;
  s->reals = (double*)calloc(m->n_reals > 0 ? m->n_reals : 1, sizeof(double));
  s->ints = (int32_t*)calloc(m->n_ints > 0 ? m->n_ints : 1, sizeof(int32_t));
  memcpy(s->reals, m->init_reals, sizeof(double) * m->n_reals);
  memcpy(s->ints, m->init_ints, sizeof(int32_t) * m->n_ints);
  s->caches = (double*)calloc(m->n_factors > 0 ? m->n_factors : 1, sizeof(double));
  s->order = (int32_t*)malloc(sizeof(int32_t) * (m->n_samplers > 0 ? m->n_samplers : 1));
  for (int i = 0; i < m->n_samplers; i++) s->order[i] = i;
  s->cur_pos = -1;
  sm_update_all(s);
  return s;
}
static void sm_free(SampledModel* s) {
  if (!s) return;
  free(s->reals); free(s->ints); free(s->caches); free(s->order); free(s);
}

/* forwardSample :284-291 */
static void sm_forward_sample(SampledModel* s, Rng* r) {
  const orc_model* m = s->m;
  for (int k = 0; k < m->n_forward; k++) {
    const Forward* f = &m->forward[k];
    switch (f->kind) {
      case FW_NORMAL: s->reals[f->var] = gen_normal(r, f->a, f->b); break;
      case FW_EXPONENTIAL: s->reals[f->var] = gen_exponential(r, f->b); break;
      case FW_GAMMA: s->reals[f->var] = gen_gamma(r, f->a, f->b); break;
      case FW_BERNOULLI_INT: s->ints[f->var] = rng_bernoulli(r, f->a) ? 1 : 0; break;
      case FW_BETA: s->reals[f->var] = gen_beta(r, f->a, f->b); break;
      case FW_BETA_LATENT: s->reals[f->var] = gen_beta(r, s->reals[0] + f->a, s->reals[1] + f->a); break;   /* parameters drawn just before */
      case FW_POISSON_INT: s->ints[f->var] = gen_poisson_small(r, f->a); break;
      case FW_UNIFORM: s->reals[f->var] = f->a + rng_double(r) * (f->b - f->a); break;   /* Generators.uniform :197-202 */
      case FW_DISCRETE_UNIFORM_INT: s->ints[f->var] = gen_discrete_uniform(r, (int)f->a, (int)f->b); break;   /* :205-222 */
      case FW_CATEGORICAL_INT: {   /* Generators.categorical :155-160 -> Random.nextCategorical: first index whose cumulative sum exceeds U */
        const double* pi = s->reals + (int)f->a; int K = f->K;
        double u = rng_double(r), cum = 0.0; int z = K - 1;
        for (int d = 0; d < K; d++) { cum += pi[d]; if (u < cum) { z = d; break; } }
        s->ints[f->var] = z;
      } break;
      case FW_DIRICHLET_SYM: {   /* Generators.dirichletInPlace :233-272 */
        double* out = s->reals + f->var; int K = f->K; double conc = f->a;
        if (conc < 1.0) {
          double lse = NEG_INF;
          for (int d = 0; d < K; d++) {
            double g = gen_gamma(r, conc + 1.0, 1.0);
            double logu = log(rng_double(r));
            out[d] = log(g) + logu / conc;
            lse = orc_log_add(lse, out[d]);
          }
          for (int d = 0; d < K; d++) out[d] = exp(out[d] - lse);
        } else {
          double sum = 0.0;
          for (int d = 0; d < K; d++) { out[d] = gen_gamma(r, conc, 1e7); sum += out[d]; }
          for (int d = 0; d < K; d++) out[d] /= sum;
        }
        for (int d = 0; d < K; d++) {
          if (out[d] == 0.0) out[d] = ZERO_PLUS_EPS; else if (out[d] == 1.0) out[d] = ONE_MINUS_EPS;
        }
      } break;
    }
  }
  sm_update_all(s);
}

/* ------------------------------------------------------------------ */
/* samplers                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
  SampledModel* s; const Sampler* sp;
  int simplex_index;                 /* SimplexWritableVariable.index */
  double simplex_sum;
} SamplerCtx;

/* RealSliceSampler.logDensity :156-161 -- sum over connected (wrapped) factors */
static double sampler_log_density(SamplerCtx* c) {
  SampledModel* s = c->s; const Sampler* sp = c->sp;
  double sum = 0.0;
  for (int k = 0; k < sp->n_fixed; k++) sum += sm_factor_log_density(s, sp->fixed_idx[k]);
  for (int k = 0; k < sp->n_ann; k++) sum += sm_factor_log_density(s, sp->ann_idx[k]);
  for (int k = 0; k < sp->n_other; k++) sum += sm_factor_log_density(s, sp->other_idx[k]);
  s->n_evals++;
  return sum;
}
static double var_get(SamplerCtx* c) {
  if (c->sp->type == S_SIMPLEX) return c->s->reals[c->sp->var + c->simplex_index];
  return c->s->reals[c->sp->var];
}
/* WritableRealVar.set ; SimplexWritableVariable.set :26-30 */
static void var_set(SamplerCtx* c, double x) {
  if (c->sp->type == S_SIMPLEX) {
    int i = c->simplex_index, K = c->sp->K, nx = (i == K - 1) ? 0 : i + 1;
    double* p = c->s->reals + c->sp->var;
    double sum = p[i] + p[nx];
    double comp = fmax(0.0, sum - x);
    p[i] = x; p[nx] = comp;
  } else c->s->reals[c->sp->var] = x;
}
static double log_density_at(SamplerCtx* c, double x) { var_set(c, x); return sampler_log_density(c); }

/* RealSliceSampler.nextLogSliceHeight :143-148 */
static double next_log_slice_height(Rng* r, double log_density) {
  if (log_density == NEG_INF) return -1e100;
  return log_density - gen_unit_exponential(r);
}
/* RealSliceSampler.accept :121-141 */
static int real_accept(SamplerCtx* c, int fixed_window, double old_state, double new_state,
                       double log_slice_height, double left, double right) {
  if (fixed_window) return 1;
  int differ = 0;
  while (right - left > 1.1 * 1.0) {
    double middle = (left + right) / 2.0;
    if ((old_state < middle && new_state >= middle) || (old_state >= middle && new_state < middle)) differ = 1;
    if (new_state < middle) right = middle; else left = middle;
    if (differ && log_slice_height >= log_density_at(c, left) && log_slice_height >= log_density_at(c, right)) return 0;
  }
  return 1;
}
/* RealSliceSampler.execute :53-119 */
static void real_slice_execute(SamplerCtx* c, Rng* r, int fixed_window, double fw_left, double fw_right) {
  if (isnan(var_get(c))) { c->s->error |= 1; return; }   /* the reference throws at the first NaN factor, ExponentiatedFactor.java:26-29 */
  const double log_slice_height = next_log_slice_height(r, sampler_log_density(c));
  const double old_state = var_get(c);
  double left, right;
  if (fixed_window) { left = fw_left; right = fw_right; }
  else {
    left = old_state - 1.0 * rng_double(r);
    right = left + 1.0;
    int iter = 0;
    while (iter++ < 20 && (log_slice_height < log_density_at(c, left) || log_slice_height < log_density_at(c, right))) {
      if (rng_bernoulli(r, 0.5)) {
        left += -(right - left);
        if (left == NEG_INF) { c->s->error |= 2; var_set(c, old_state); return; }
      } else {
        right += right - left;
        if (right == INFINITY) { c->s->error |= 2; var_set(c, old_state); return; }
      }
    }
  }
  double sl = left, sr = right;
  for (;;) {
    const double new_state = gen_uniform(r, sl, sr);
    if (log_slice_height <= log_density_at(c, new_state) &&
        real_accept(c, fixed_window, old_state, new_state, log_slice_height, left, right)) {
      var_set(c, new_state);
      return;
    }
    if (new_state < old_state) sl = new_state; else sr = new_state;
    if (is_close(sl, sr, 1e-6)) { var_set(c, old_state); return; }
  }
}

/* IntSliceSampler.accept :115-132 */
static double int_log_density_at(SamplerCtx* c, int x) { c->s->ints[c->sp->var] = x; return sampler_log_density(c); }
static int int_accept(SamplerCtx* c, int old_state, int new_state, double h, int left, int right) {
  int differ = 0;
  while (right - left > 10) {
    int middle = (left + right) / 2;
    if ((old_state < middle && new_state >= middle) || (old_state >= middle && new_state < middle)) differ = 1;
    if (new_state < middle) right = middle; else left = middle;
    if (differ && h >= int_log_density_at(c, left) && h >= int_log_density_at(c, right - 1)) return 0;
  }
  return 1;
}
/* IntSliceSampler.execute :53-113 */
static void int_slice_execute(SamplerCtx* c, Rng* r, int fixed_window, int fw_left, int fw_right) {
  const double h = next_log_slice_height(r, sampler_log_density(c));
  const int old_state = c->s->ints[c->sp->var];
  int left, right;
  if (fixed_window) { left = fw_left; right = fw_right; }
  else {
    left = old_state - rng_int(r, 10);
    right = left + 10;
    int iter = 0;
    while (iter++ < 20 && (h < int_log_density_at(c, left) || h < int_log_density_at(c, right - 1))) {
      if (rng_bernoulli(r, 0.5)) left += -(right - left);
      else right += right - left;
    }
  }
  int sl = left, sr = right;
  for (;;) {
    if (sr <= sl) {   /* Generators.discreteUniform :205-211 throws "Invalid arguments": no point of the window is acceptable */
      c->s->error |= ORC_ERR_INT_SLICE_EMPTY;
      c->s->ints[c->sp->var] = old_state; return;
    }
    const int new_state = gen_discrete_uniform(r, sl, sr);
    if (h <= int_log_density_at(c, new_state) && int_accept(c, old_state, new_state, h, left, right)) {
      c->s->ints[c->sp->var] = new_state;
      return;
    }
    if (new_state < old_state) sl = new_state + 1; else sr = new_state;
  }
}

/* Sampler.execute dispatch */
static void sampler_execute(SampledModel* s, int k, Rng* r) {
  SamplerCtx c; c.s = s; c.sp = &s->m->samplers[k]; c.simplex_index = 0; c.simplex_sum = 0;
  switch (c.sp->type) {
    case S_REAL_SLICE: real_slice_execute(&c, r, 0, 0, 0); break;
    case S_SIMPLEX: {                                  /* SimplexSampler.xtend:21-28 */
      int K = c.sp->K;
      c.simplex_index = rng_int(r, K);
      int nx = (c.simplex_index == K - 1) ? 0 : c.simplex_index + 1;
      double sum = s->reals[c.sp->var + c.simplex_index] + s->reals[c.sp->var + nx];
      real_slice_execute(&c, r, 1, 0.0, sum);
    } break;
    case S_CATEGORICAL: int_slice_execute(&c, r, 1, 0, c.sp->K); break;   /* CategoricalSampler.xtend:24-28 */
    case S_INT_SLICE: int_slice_execute(&c, r, 0, 0, 0); break;             /* IntSliceSampler on a latent IntVar */
    case S_UNIFORM: int_slice_execute(&c, r, 1, c.sp->lo, c.sp->K); break;  /* UniformSampler.xtend:23-27 */
  }
}

/* posteriorSamplingStep(rand) :258-282 -- `step` tags the Philox counter */
static void sm_posterior_sampling_step(SampledModel* s, uint64_t seed, uint32_t replica,
                                       uint32_t position, uint32_t scan, uint32_t step) {
  int S = s->m->n_samplers;
  if (S == 0) return;
  if (s->cur_pos == -1) {
    /* Collections.shuffle(list, rnd): for i=size..2 swap(i-1, nextInt(i)) */
    Rng sh = rng_make(seed, replica, P_SHUFFLE, position, scan, step);
    for (int i = S; i > 1; i--) {
      int j = rng_int(&sh, i);
      int32_t t = s->order[i - 1]; s->order[i - 1] = s->order[j]; s->order[j] = t;
    }
    s->cur_pos = S - 1;
  }
  int k = s->order[s->cur_pos--];
  Rng r = rng_make(seed, replica, P_MOVE, position, scan, step);
  sampler_execute(s, k, &r);
  sm_update(s, k);
  s->cumulative_n_scans += 1.0 / S;
  if (s->cumulative_n_scans > 10.0) { sm_update_all(s); s->cumulative_n_scans = 0.0; }
}

/* ------------------------------------------------------------------ */
/* accumulators                                                         */
/* ------------------------------------------------------------------ */
/* commons-math3 SummaryStatistics (FirstMoment / SecondMoment recurrences) */
typedef struct { int64_t n; double m1, m2; double mn, mx; } Summary;
static void summary_reset(Summary* s) { s->n = 0; s->m1 = NAN; s->m2 = NAN; s->mn = NAN; s->mx = NAN; }
static void summary_add(Summary* s, double x) {
  if (s->n == 0) { s->m1 = 0.0; s->m2 = 0.0; }
  s->n++;
  double dev = x - s->m1, ndev = dev / (double)s->n;
  s->m1 += ndev;
  s->m2 += ((double)s->n - 1.0) * dev * ndev;
  if (s->n == 1 || x < s->mn) s->mn = x;
  if (s->n == 1 || x > s->mx) s->mx = x;
}
static double summary_mean(const Summary* s) { return s->n == 0 ? NAN : s->m1; }
static double summary_var(const Summary* s) {
  if (s->n == 0) return NAN;
  if (s->n == 1) return 0.0;
  return s->m2 / ((double)s->n - 1.0);
}
/* CovarAccumulator.java:7-16 */
typedef struct { double meanx, meany, C; int n; } Covar;
static void covar_add(Covar* c, double x, double y) {
  c->n += 1;
  double dx = x - c->meanx;
  c->meanx += dx / c->n;
  c->meany += (y - c->meany) / c->n;
  c->C += dx * (y - c->meany);
}
/* LogSumAccumulator.java:13-30 */
typedef struct { double log_sum; int64_t n; } LogSum;

/* ------------------------------------------------------------------ */
/* monotone cubic spline (R/engines/internals/Spline.java:129-203)      */
/* ------------------------------------------------------------------ */
void orc_monotone_spline_build(const double* x, const double* y, int n, double* m) {
  if (n < 2) { for (int i = 0; i < n; i++) m[i] = NAN; return; }   /* the reference throws: "at least two control points" (:130-133) */
  double* d = (double*)malloc(sizeof(double) * (n - 1));
  for (int i = 0; i < n - 1; i++) { double h = x[i + 1] - x[i]; d[i] = (y[i + 1] - y[i]) / h; }
  m[0] = d[0];
  for (int i = 1; i < n - 1; i++) m[i] = (d[i - 1] + d[i]) * 0.5;
  m[n - 1] = d[n - 2];
  for (int i = 0; i < n - 1; i++) {
    if (d[i] == 0.0) { m[i] = 0.0; m[i + 1] = 0.0; }
    else {
      double a = m[i] / d[i], b = m[i + 1] / d[i];
      double h = hypot(a, b);
      if (h > 3.0) { double t = 3.0 / h; m[i] *= t; m[i + 1] *= t; }
    }
  }
  free(d);
}
double orc_monotone_spline_eval(const double* mX, const double* mY, const double* mM, int n, double x) {
  if (isnan(x)) return x;
  if (x <= mX[0]) return mY[0];
  if (x >= mX[n - 1]) return mY[n - 1];
  int i = 0;
  while (x >= mX[i + 1]) { i += 1; if (x == mX[i]) return mY[i]; }
  double h = mX[i + 1] - mX[i];
  double t = (x - mX[i]) / h;
  return (mY[i] * (1 + 2 * t) + h * mM[i] * t) * (1 - t) * (1 - t) +
         (mY[i + 1] * (3 - 2 * t) + h * mM[i + 1] * (t - 1)) * t * t;
}

/* EngineStaticUtils.estimateCumulativeFunctionsFromIntensities :124-155 +
 * fixedSizeOptimalPartitionFromScheduleGenerator :157-168 */
static void cumulative_from_intensities(const double* betas_asc, const double* intens, int n,
                                        double nudge, double* ys) {
  ys[0] = 0.0;
  for (int i = 1; i < n; i++) ys[i] = ys[i - 1] + intens[i - 1] + nudge;   /* paddedCumulative :86-93 */
}
void orc_adapt_ladder(const double* betas_asc, const double* accept_asc, int n, double* out_asc,
                      double* cum_lambda_out) {
  double* intens = (double*)malloc(sizeof(double) * n);
  double* ys = (double*)malloc(sizeof(double) * n);
  double* mm = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n - 1; i++) intens[i] = 1.0 - accept_asc[i];     /* acceptProbabilitiesToIntensities :95-105 */
  if (cum_lambda_out) cumulative_from_intensities(betas_asc, intens, n, 0.0, cum_lambda_out);
  double nudge = 0.0;
  for (int attempt = 0; attempt < 2; attempt++) {
    cumulative_from_intensities(betas_asc, intens, n, nudge, ys);
    double norm = ys[n - 1]; int retry = 0;
    for (int i = 0; i < n; i++) {
      ys[i] /= norm;
      if (i > 0 && ys[i] == ys[i - 1]) { retry = 1; break; }
    }
    if (!retry) break;
    nudge = 1e-6;
  }
  /* schedule generator: spline through (ys, betas) */
  orc_monotone_spline_build(ys, betas_asc, n, mm);
  out_asc[0] = 0.0;
  for (int i = 1; i < n - 1; i++) out_asc[i] = orc_monotone_spline_eval(ys, betas_asc, mm, n, i / (n - 1.0));
  out_asc[n - 1] = 1.0;
  free(intens); free(ys); free(mm);
}

/* ------------------------------------------------------------------ */
/* ParallelTempering + PT                                                */
/* ------------------------------------------------------------------ */
struct orc_pt {
  const orc_model* m;
  orc_pt_config cfg;
  int n;
  SampledModel** states;            /* index = temperature position; 0 is beta=1 */
  double* betas;                    /* temperingParameters, descending */
  Summary *energies, *swap_accept;
  LogSum* log_sums;
  Covar* covars;
  uint32_t swap_index;              /* never reset (ParallelTempering.java:49,67) */
  uint32_t scan_index;
  /* in-memory replacement for swapIndicators.csv + Paths.xtend: per particle flag */
  int32_t* particle_of_pos;         /* identity of the particle at each position (its position at round start) */
  uint8_t* new_sample;              /* Paths.nRejuvenations state per particle */
  int64_t rejuvenations;
  int n_scans_in_round;
  uint64_t retired_evals, retired_factor_evals;
};

void orc_ladder_equally_spaced(int n, double* out) {
  if (n == 1) { out[0] = 1.0; return; }
  int k = 0;
  for (int i = n - 1; i >= 0; i--) out[k++] = ((double)i) / ((double)n - 1.0);
}

static void paths_visit(orc_pt* pt, int particle, int position) {   /* Paths.xtend:40-50 */
  if (position == 0 && pt->new_sample[particle]) { pt->new_sample[particle] = 0; pt->rejuvenations++; }
  else if (position == pt->n - 1) pt->new_sample[particle] = 1;
}
static void paths_reset(orc_pt* pt) {                               /* Paths.initPaths :131-139 */
  pt->rejuvenations = 0;
  for (int c = 0; c < pt->n; c++) { pt->particle_of_pos[c] = c; pt->new_sample[c] = 0; }
  for (int c = 0; c < pt->n; c++) paths_visit(pt, c, c);
}

static int cmp_desc(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x < y) - (x > y);
}
/* setAnnealingParameters :183-205 */
void orc_pt_set_ladder(orc_pt* pt, const double* betas, int n) {
  if (n != pt->n) return;
  memcpy(pt->betas, betas, sizeof(double) * n);
  qsort(pt->betas, n, sizeof(double), cmp_desc);
  for (int i = 0; i < n; i++) pt->states[i]->exponent = pt->betas[i];
  for (int i = 0; i < n; i++) { summary_reset(&pt->energies[i]); memset(&pt->covars[i], 0, sizeof(Covar)); }
  for (int i = 0; i < n - 1; i++) { summary_reset(&pt->swap_accept[i]); pt->log_sums[i].log_sum = NEG_INF; pt->log_sums[i].n = 0; }
  paths_reset(pt);
  pt->n_scans_in_round = 0;
}
void orc_pt_get_ladder(const orc_pt* pt, double* out) { memcpy(out, pt->betas, sizeof(double) * pt->n); }

orc_pt* orc_pt_create(const orc_model* m, const orc_pt_config* cfg) {
  orc_pt* pt = (orc_pt*)calloc(1, sizeof(orc_pt));
  pt->m = m; pt->cfg = *cfg; pt->n = cfg->n_chains;
  int n = pt->n;
  pt->states = (SampledModel**)calloc(n, sizeof(SampledModel*));
  for (int i = 0; i < n; i++) pt->states[i] = sm_new(m, cfg->anneal_support);   /* initStates :207-216 */
  pt->betas = (double*)calloc(n, sizeof(double));
  pt->energies = (Summary*)calloc(n, sizeof(Summary));
  pt->swap_accept = (Summary*)calloc(n, sizeof(Summary));
  pt->log_sums = (LogSum*)calloc(n, sizeof(LogSum));
  pt->covars = (Covar*)calloc(n, sizeof(Covar));
  pt->particle_of_pos = (int32_t*)calloc(n, sizeof(int32_t));
  pt->new_sample = (uint8_t*)calloc(n, 1);
  double* lad = (double*)malloc(sizeof(double) * n);
  orc_ladder_equally_spaced(n, lad);
  orc_pt_set_ladder(pt, lad, n);
  free(lad);
  return pt;
}
void orc_pt_free(orc_pt* pt) {
  if (!pt) return;
  for (int i = 0; i < pt->n; i++) sm_free(pt->states[i]);
  free(pt->states); free(pt->betas); free(pt->energies); free(pt->swap_accept);
  free(pt->log_sums); free(pt->covars); free(pt->particle_of_pos); free(pt->new_sample); free(pt);
}
void orc_pt_set_state(orc_pt* pt, int pos, const double* re, const int32_t* in) {
  SampledModel* s = pt->states[pos];
  if (re) memcpy(s->reals, re, sizeof(double) * pt->m->n_reals);
  if (in) memcpy(s->ints, in, sizeof(int32_t) * pt->m->n_ints);
  sm_update_all(s);
}
void orc_pt_get_state(const orc_pt* pt, int pos, double* re, int32_t* in) {
  const SampledModel* s = pt->states[pos];
  if (re) memcpy(re, s->reals, sizeof(double) * pt->m->n_reals);
  if (in) memcpy(in, s->ints, sizeof(int32_t) * pt->m->n_ints);
}
/* PT.informedInitialization :284-298 (COPIES / FORWARD; SCM is out of scope) */
void orc_pt_initialize(orc_pt* pt, int init_type) {
  if (init_type != ORC_INIT_FORWARD) return;
  for (int i = 0; i < pt->n; i++) {
    SampledModel* s = pt->states[i];
    double c = s->exponent; s->exponent = 0.0;
    Rng r = rng_make(pt->cfg.seed, pt->cfg.replica, P_INIT, (uint32_t)i, 0, 0);
    sm_forward_sample(s, &r);
    s->exponent = c;
  }
}
void orc_pt_densities(const orc_pt* pt, double* logdens, double* loglik, int32_t* noos, double* fixed) {
  for (int i = 0; i < pt->n; i++) {
    const SampledModel* s = pt->states[i];
    if (logdens) logdens[i] = sm_log_density(s);
    if (loglik) loglik[i] = s->sum_pre;
    if (noos) noos[i] = s->n_oos;
    if (fixed) fixed[i] = s->sum_fixed;
  }
}
double orc_pt_check_caches(const orc_pt* pt) {   /* SampledModel.check :178-188 */
  double worst = 0.0;
  for (int i = 0; i < pt->n; i++) {
    SampledModel* s = pt->states[i];
    double expected = sm_log_density(s), sum = 0.0;
    for (int f = 0; f < pt->m->n_factors; f++) sum += sm_factor_log_density(s, f);
    if (expected == NEG_INF && sum == NEG_INF) continue;
    double d = fabs(expected - sum) / fmax(1.0, fabs(sum));
    if (!(d <= worst)) worst = d;
  }
  return worst;
}

/* moveKernel :77-92.  The checkerboard order (Ising only) is the sweep the
 * CUDA engine uses: colour 0 then colour 1, each spin once per pass, Philox
 * step = pass*nSpins + spin.  Same invariant law as the reference's random
 * scan. */
static void move_one(orc_pt* pt, int c) {
  SampledModel* cur = pt->states[c];
  const orc_model* m = pt->m;
  double before = -cur->sum_pre;
  if (pt->betas[c] == 0.0 && pt->cfg.use_prior_samples) {
    Rng r = rng_make(pt->cfg.seed, pt->cfg.replica, P_FORWARD, (uint32_t)c, pt->scan_index, 0);
    sm_forward_sample(cur, &r);
  } else if (pt->cfg.sweep_order == ORC_ORDER_CHECKERBOARD && m->family == ORC_FAM_ISING) {
    int N = m->N, S = N * N;
    int n_pass = (int)floor(pt->cfg.n_passes_per_scan);
    for (int pass = 0; pass < n_pass; pass++)
      for (int colour = 0; colour < 2; colour++)
        for (int v = 0; v < S; v++) {
          if ((((v / N) + (v % N)) & 1) != colour) continue;
          Rng r = rng_make(pt->cfg.seed, pt->cfg.replica, P_MOVE, (uint32_t)c, pt->scan_index, (uint32_t)(pass * S + v));
          sampler_execute(cur, v, &r);
          sm_update(cur, v);
        }
  } else {
    int n_steps = (int)floor(pt->cfg.n_passes_per_scan * m->n_samplers);   /* posteriorSamplingScan :249-254 */
    for (int t = 0; t < n_steps; t++)
      sm_posterior_sampling_step(cur, pt->cfg.seed, pt->cfg.replica, (uint32_t)c, pt->scan_index, (uint32_t)t);
  }
  double after = -cur->sum_pre;
  summary_add(&pt->energies[c], after);
  covar_add(&pt->covars[c], before, after);
}
/* BriefParallel.process(nChains, nThreads, ...) :79 -- a pthread parallel-for with a shared counter */
typedef struct { orc_pt* pt; int next; pthread_mutex_t mu; } MoveJob;
static void* move_worker(void* arg) {
  MoveJob* job = (MoveJob*)arg;
  for (;;) {
    pthread_mutex_lock(&job->mu);
    int c = job->next++;
    pthread_mutex_unlock(&job->mu);
    if (c >= job->pt->n) break;
    move_one(job->pt, c);
  }
  return NULL;
}
void orc_pt_move_kernel(orc_pt* pt) {
  int nt = pt->cfg.n_threads;
  if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nt > pt->n) nt = pt->n;
  if (nt <= 1) { for (int c = 0; c < pt->n; c++) move_one(pt, c); return; }
  MoveJob job; job.pt = pt; job.next = 0; pthread_mutex_init(&job.mu, NULL);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, move_worker, &job);
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&job.mu);
  free(th);
}

/* swapAcceptPr :107-124 */
static double swap_accept_pr(orc_pt* pt, int i) {
  int j = i + 1;
  double ss = sm_log_density_at(pt->states[j], pt->betas[i]) - sm_log_density_at(pt->states[j], pt->betas[j]);
  pt->log_sums[i].log_sum = orc_log_add(pt->log_sums[i].log_sum, ss);
  pt->log_sums[i].n++;
  double log_ratio = ss + sm_log_density_at(pt->states[i], pt->betas[j]) - sm_log_density_at(pt->states[i], pt->betas[i]);
  double a = fmin(1.0, exp(log_ratio));
  if (isnan(log_ratio) || isnan(a)) a = 0.0;
  return a;
}
/* swapKernel :64-75, :99-105; doSwap :155-164 */
void orc_pt_swap_kernel(orc_pt* pt, int32_t* indicators_out) {
  int n = pt->n;
  int32_t* ind = (int32_t*)calloc(n, sizeof(int32_t));
  int offset;
  if (pt->cfg.reversible) {
    Rng r = rng_make(pt->cfg.seed, pt->cfg.replica, P_SWAPDIR, 0, pt->scan_index, 0);
    offset = rng_int(&r, 2);
  } else offset = (int)(pt->swap_index++ % 2);
  for (int k = 0; k < (n - offset) / 2; k++) {
    int i = offset + 2 * k;
    double a = swap_accept_pr(pt, i);
    Rng r = rng_make(pt->cfg.seed, pt->cfg.replica, P_SWAP, (uint32_t)i, pt->scan_index, 0);
    if (rng_bernoulli(&r, a)) {
      SampledModel* t = pt->states[i]; pt->states[i] = pt->states[i + 1]; pt->states[i + 1] = t;
      pt->states[i]->exponent = pt->betas[i]; pt->states[i + 1]->exponent = pt->betas[i + 1];
      int32_t q = pt->particle_of_pos[i]; pt->particle_of_pos[i] = pt->particle_of_pos[i + 1]; pt->particle_of_pos[i + 1] = q;
      ind[i] = 1;
    }
    summary_add(&pt->swap_accept[i], a);
  }
  for (int c = 0; c < n; c++) paths_visit(pt, pt->particle_of_pos[c], c);   /* Paths ctor :83-117 */
  if (indicators_out) memcpy(indicators_out, ind, sizeof(int32_t) * n);
  free(ind);
}
void orc_pt_scan(orc_pt* pt, double* energies_out, int32_t* indicators_out) {
  orc_pt_move_kernel(pt);
  if (energies_out) for (int c = 0; c < pt->n; c++) energies_out[c] = -pt->states[c]->sum_pre;
  if (pt->n > 1) orc_pt_swap_kernel(pt, indicators_out);
  else if (indicators_out) indicators_out[0] = 0;
  pt->scan_index++;
  pt->n_scans_in_round++;
}
uint64_t orc_pt_n_evals(const orc_pt* pt) { uint64_t t = 0; for (int i = 0; i < pt->n; i++) t += pt->states[i]->n_evals; return t; }
uint64_t orc_pt_n_factor_evals(const orc_pt* pt) { uint64_t t = 0; for (int i = 0; i < pt->n; i++) t += pt->states[i]->n_factor_evals; return t; }
uint64_t orc_pt_n_scans(const orc_pt* pt) { return pt->scan_index; }
int orc_pt_error(const orc_pt* pt) {
  int e = 0;
  for (int i = 0; i < pt->n; i++) e |= pt->states[i]->error;
  return e;
}

void orc_pt_round_stats(const orc_pt* pt, orc_round_stats* o) {
  int n = pt->n;
  o->n_chains = n; o->n_scans_in_round = pt->n_scans_in_round;
  double Lambda = 0.0; int oos = 0;
  for (int i = 0; i < n; i++) {
    o->energy_mean[i] = summary_mean(&pt->energies[i]);
    o->energy_var[i] = summary_var(&pt->energies[i]);
    o->energy_cov[i] = pt->covars[i].C / pt->covars[i].n;
    if (pt->states[i]->oos_detected) oos = 1;
  }
  for (int i = 0; i < n - 1; i++) {
    o->swap_accept_mean[i] = summary_mean(&pt->swap_accept[i]);
    o->log_sum[i] = pt->log_sums[i].log_sum; o->log_sum_n[i] = pt->log_sums[i].n;
    Lambda += 1.0 - o->swap_accept_mean[i];
  }
  o->Lambda = Lambda;
  o->round_trips = pt->rejuvenations;
  if (n == 1) o->log_z_stepping_stone = NAN;
  else {
    double r = 0.0;
    for (int c = 0; c < n - 1; c++) r += pt->log_sums[c].log_sum - log((double)pt->log_sums[c].n);
    o->log_z_stepping_stone = r;
  }
  if (oos) o->log_z_thermo = NAN;
  else {
    double sum = 0.0;
    for (int c = 0; c < n - 1; c++) sum += (pt->betas[c] - pt->betas[c + 1]) * (summary_mean(&pt->energies[c]) + summary_mean(&pt->energies[c + 1])) / 2.0;
    o->log_z_thermo = -sum;
  }
}

/* PT.adapt :148-171 (targetAccept branch not restated) */
void orc_pt_adapt(orc_pt* pt, double* knots) {
  int n = pt->n;
  double* asc = (double*)malloc(sizeof(double) * n);
  double* acc = (double*)malloc(sizeof(double) * n);
  double* out = (double*)malloc(sizeof(double) * n);
  double* cum = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; i++) asc[i] = pt->betas[n - 1 - i];
  for (int i = 0; i < n - 1; i++) {
    double r = summary_mean(&pt->swap_accept[i]);
    if (r == 1.0) r = 0.99999; else if (!isfinite(r)) r = 0.0;
    acc[n - 2 - i] = r;
  }
  orc_adapt_ladder(asc, acc, n, out, cum);
  if (knots) { memcpy(knots, asc, sizeof(double) * n); memcpy(knots + n, cum, sizeof(double) * n); }
  orc_pt_set_ladder(pt, out, n);
  free(asc); free(acc); free(out); free(cum);
}

/* PT.rounds :239-265 */
int orc_rounds(int n_scans, double f, int32_t* rn, int32_t* ra) {
  int32_t tn[64], ta[64]; int k = 0;
  int remaining = (int)(f * n_scans);
  tn[k] = n_scans - remaining; ta[k] = 0; k++;
  while (remaining > 0 && k < 64) {
    int next = (int)(f * remaining) - 1;
    tn[k] = remaining - next; ta[k] = 1; k++;
    remaining = next;
  }
  for (int i = 0; i < k; i++) { rn[i] = tn[k - 1 - i]; ra[i] = ta[k - 1 - i]; }
  return k;
}

/* PT.performInference :85-113 */
int orc_pt_perform_inference(orc_pt* pt, int n_scans, double adapt_fraction,
                             double* samples_out, int32_t* int_samples_out,
                             double* energies_out, int32_t* indicators_out, orc_run_summary* sum) {
  int n = pt->n;
  if (n == 1) adapt_fraction = 0.0;                       /* PT.java:235 */
  int32_t rn[64], ra[64];
  int n_rounds = orc_rounds(n_scans, adapt_fraction, rn, ra);
  int scan = 0;
  orc_round_stats st;
  st.swap_accept_mean = (double*)malloc(sizeof(double) * n); st.energy_mean = (double*)malloc(sizeof(double) * n);
  st.energy_var = (double*)malloc(sizeof(double) * n); st.energy_cov = (double*)malloc(sizeof(double) * n);
  st.log_sum = (double*)malloc(sizeof(double) * n); st.log_sum_n = (int64_t*)malloc(sizeof(int64_t) * n);
  if (sum) { memset(sum, 0, sizeof *sum); sum->n_rounds = n_rounds; }
  for (int r = 0; r < n_rounds; r++) {
    for (int s = 0; s < rn[r]; s++) {
      orc_pt_move_kernel(pt);
      if (energies_out) for (int c = 0; c < n; c++) energies_out[(size_t)scan * n + c] = -pt->states[c]->sum_pre;
      if (samples_out && pt->m->n_reals) memcpy(samples_out + (size_t)scan * pt->m->n_reals, pt->states[0]->reals, sizeof(double) * pt->m->n_reals);
      if (int_samples_out && pt->m->n_ints) memcpy(int_samples_out + (size_t)scan * pt->m->n_ints, pt->states[0]->ints, sizeof(int32_t) * pt->m->n_ints);
      if (n > 1) orc_pt_swap_kernel(pt, indicators_out ? indicators_out + (size_t)scan * n : NULL);
      pt->scan_index++; pt->n_scans_in_round++;
      scan++;
    }
    if (n > 1) {
      orc_pt_round_stats(pt, &st);
      if (sum) {
        sum->round_n_scans[r] = rn[r]; sum->round_is_adapt[r] = ra[r];
        sum->round_Lambda[r] = st.Lambda; sum->round_log_z[r] = st.log_z_stepping_stone; sum->round_trips[r] = st.round_trips;
        double mn = INFINITY, mean = 0;
        for (int i = 0; i < n - 1; i++) { if (st.swap_accept_mean[i] < mn) mn = st.swap_accept_mean[i]; mean += st.swap_accept_mean[i]; }
        sum->round_min_accept[r] = mn; sum->round_mean_accept[r] = mean / (n - 1);
      }
      /* the last round's adapt() is "instrumentation only" (PT.java:103-107): the caller
       * reads final statistics from `sum`; we still adapt to mirror the ladder change */
      if (r < n_rounds - 1) orc_pt_adapt(pt, NULL);
    } else if (sum) { sum->round_n_scans[r] = rn[r]; sum->round_is_adapt[r] = ra[r]; }
  }
  free(st.swap_accept_mean); free(st.energy_mean); free(st.energy_var); free(st.energy_cov); free(st.log_sum); free(st.log_sum_n);
  return scan;
}

/* ------------------------------------------------------------------ */
/* SCM initialisation of PT (PT.informedInitialization, PT.java:300-329;  */
/* AdaptiveJarzynski.java:91-233; AdaptiveTemperatureSchedule.java:48-78; */
/* EngineStaticUtils.relativeCESS :27-47).  bayonet's ParticlePopulation  */
/* and commons-math3's PegasusSolver are restated from their published    */
/* algorithms (neither is in the tree).                                   */
/* ------------------------------------------------------------------ */
enum { P_MASTER = 6 };

typedef struct { SampledModel** p; double* w; int n; double t; int as; double thr; int err; } ScmPop;

static int scm_ratio(SampledModel* s, double t, double next, double* out) {   /* SampledModel.logDensityRatio :199-206 */
  double num = sm_log_density_at(s, next), den = sm_log_density_at(s, t);
  if (num == NEG_INF && den == NEG_INF) return 0;
  *out = num - den;
  return 1;
}
static double scm_objective(ScmPop* q, double x) {
  double lnum = NEG_INF, lden = NEG_INF;
  for (int i = 0; i < q->n; i++) {
    double r;
    if (!scm_ratio(q->p[i], q->t, x, &r)) { q->err = 1; return NAN; }
    lnum = orc_log_add(lnum, log(q->w[i]) + 1.0 * r);
    lden = orc_log_add(lden, log(q->w[i]) + 2.0 * r);
  }
  return exp(2.0 * lnum - lden) - q->thr;
}
static double scm_pegasus(ScmPop* q, double x0, double x1, double atol, int max_eval) {
  const double ftol = 1e-15, rtol = 1e-14;
  double f0 = scm_objective(q, x0), f1 = scm_objective(q, x1);
  int evals = 2;
  if (f0 == 0.0) return x0;
  if (f1 == 0.0) return x1;
  for (;;) {
    double x = x1 - ((f1 * (x1 - x0)) / (f1 - f0));
    double fx = scm_objective(q, x);
    if (++evals > max_eval) return x;
    if (fx == 0.0) return x;
    if (f1 * fx < 0) { x0 = x1; f0 = f1; }
    else f0 *= f1 / (f1 + fx);
    x1 = x; f1 = fx;
    if (fabs(f1) <= ftol) return x1;
    if (fabs(x1 - x0) < fmax(rtol * fabs(x1), atol)) return x1;
  }
}
static void sm_copy(SampledModel* dst, const SampledModel* src) {   /* SampledModel.duplicate (deep clone) */
  const orc_model* m = src->m;
  memcpy(dst->reals, src->reals, sizeof(double) * m->n_reals);
  memcpy(dst->ints, src->ints, sizeof(int32_t) * m->n_ints);
  memcpy(dst->caches, src->caches, sizeof(double) * m->n_factors);
  memcpy(dst->order, src->order, sizeof(int32_t) * m->n_samplers);
  dst->cur_pos = src->cur_pos; dst->cumulative_n_scans = src->cumulative_n_scans;
  dst->sum_pre = src->sum_pre; dst->sum_fixed = src->sum_fixed; dst->n_oos = src->n_oos;
  dst->oos_detected = src->oos_detected; dst->exponent = src->exponent;
}

int orc_pt_initialize_scm(orc_pt* pt, int n_particles, double threshold, double ess_threshold,
                          double* log_norm_out, int32_t* n_temps_out, int32_t* n_resample_out) {
  const orc_model* m = pt->m;
  const int P = n_particles, N = pt->n;
  const uint64_t seed = pt->cfg.seed ^ 0x53434D2D696E6974ULL;
  orc_pt_initialize(pt, ORC_INIT_FORWARD);
  SampledModel** ps = (SampledModel**)calloc(P, sizeof(SampledModel*));
  SampledModel** tmp = (SampledModel**)calloc(P, sizeof(SampledModel*));
  double* w = (double*)malloc(sizeof(double) * P);
  double* lw = (double*)malloc(sizeof(double) * P);
  int* anc = (int*)malloc(sizeof(int) * P);
  for (int i = 0; i < P; i++) {                       /* AdaptiveJarzynski.sampleInitial :214-222 */
    ps[i] = sm_new(m, pt->cfg.anneal_support); tmp[i] = sm_new(m, pt->cfg.anneal_support);
    ps[i]->exponent = 0.0;
    Rng r = rng_make(seed, 0, P_INIT, (uint32_t)i, 0, 0);
    sm_forward_sample(ps[i], &r);
    w[i] = 1.0 / P;
  }
  ScmPop q; q.p = ps; q.w = w; q.n = P; q.t = 0.0; q.as = pt->cfg.anneal_support; q.thr = threshold; q.err = 0;
  double log_scaling = 0.0;
  uint32_t iter = 0, master = 0;
  int n_temps = 0, n_res = 0, status = 0;
  for (int li = N - 1; li >= 0 && !status; li--) {
    const double target = pt->betas[li];
    if (target == 0.0) continue;
    while (q.t < target) {
      double next;
      double at_one = scm_objective(&q, 1.0);
      if (q.err) { status = 1; break; }
      if (isnan(at_one)) next = fmax(q.t, 1e-10);
      else next = scm_objective(&q, target) >= 0 ? target : scm_pegasus(&q, q.t, target, 1e-16, 100);
      if (!(next > q.t)) next = target;
      int need = 0; double den = 0.0;
      for (int i = 0; i < P; i++) { if (w[i] == 0.0) need = 1; den += w[i] * w[i]; }
      double rel_ess = 1.0 / (den * P);
      if (need || (rel_ess < ess_threshold && next < target)) {     /* stratified resampling */
        double cum = w[0]; int j = 0;
        for (int i = 0; i < P; i++) {
          Rng r = rng_make(seed, 0, P_MASTER, 0, master++, 0);
          double u = (i + rng_double(&r)) / P;
          while (u > cum && j < P - 1) cum += w[++j];
          anc[i] = j;
        }
        for (int i = 0; i < P; i++) sm_copy(tmp[i], ps[anc[i]]);
        for (int i = 0; i < P; i++) { SampledModel* t = ps[i]; ps[i] = tmp[i]; tmp[i] = t; w[i] = 1.0 / P; }
        n_res++;
      }
      double lsum = NEG_INF;
      for (int i = 0; i < P; i++) {                                 /* propose :171-212 */
        double r;
        if (!scm_ratio(ps[i], q.t, next, &r)) { status = 1; break; }
        lw[i] = r + log(w[i]);
        lsum = orc_log_add(lsum, lw[i]);
      }
      if (status) break;
      for (int i = 0; i < P; i++) {                                 /* sampleNext :224-233: one sampler step */
        SampledModel* s = ps[i];
        s->exponent = next;
        if (pt->cfg.sweep_order == ORC_ORDER_CHECKERBOARD && m->family == ORC_FAM_ISING) {
          int Nn = m->N, S = Nn * Nn;
          for (int colour = 0; colour < 2; colour++)
            for (int v = 0; v < S; v++) {
              if ((((v / Nn) + (v % Nn)) & 1) != colour) continue;
              Rng r = rng_make(seed, 0, P_MOVE, (uint32_t)i, iter, (uint32_t)v);
              sampler_execute(s, v, &r);
              sm_update(s, v);
            }
        } else sm_posterior_sampling_step(s, seed, 0, (uint32_t)i, iter, 0);
      }
      for (int i = 0; i < P; i++) w[i] = exp(lw[i] - lsum);
      log_scaling += lsum;
      q.t = next; iter++; n_temps++;
    }
    if (status) break;
    Rng r = rng_make(seed, 0, P_MASTER, 0, master++, 0);
    double u = rng_double(&r), cum = 0.0; int pick = P - 1;
    for (int i = 0; i < P; i++) { cum += w[i]; if (u < cum) { pick = i; break; } }
    orc_pt_set_state(pt, li, ps[pick]->reals, ps[pick]->ints);
  }
  if (log_norm_out) *log_norm_out = log_scaling;
  if (n_temps_out) *n_temps_out = n_temps;
  if (n_resample_out) *n_resample_out = n_res;
  for (int i = 0; i < P; i++) { sm_free(ps[i]); sm_free(tmp[i]); }
  free(ps); free(tmp); free(w); free(lw); free(anc);
  return status;
}

/* one-shot from-scratch density */
void orc_model_log_density(const orc_model* m, const double* reals, const int32_t* ints,
                           double beta, int anneal_support, double out[4]) {
  SampledModel* s = sm_new(m, anneal_support);
  if (reals) memcpy(s->reals, reals, sizeof(double) * m->n_reals);
  if (ints) memcpy(s->ints, ints, sizeof(int32_t) * m->n_ints);
  s->exponent = beta;
  sm_update_all(s);
  out[0] = sm_log_density(s); out[1] = s->sum_pre; out[2] = (double)s->n_oos; out[3] = s->sum_fixed;
  sm_free(s);
}

/* factories/Exact.java:82-92 -- exhaustive enumeration, tiny lattices only */
double orc_ising_exact_log_z(int N, double coupling, double moment, double beta) {
  int S = N * N;
  if (S > 25) return NAN;
  orc_model* m = orc_model_ising(N, coupling, moment);
  int ne = 0; int32_t(*e)[2] = (int32_t(*)[2])malloc(sizeof(int32_t) * 2 * m->n_factors);
  for (int f = 0; f < m->n_factors; f++) if (m->factors[f].kind == F_ISING_EDGE) { e[ne][0] = m->factors[f].i0; e[ne][1] = m->factors[f].i1; ne++; }
  double p = orc_logistic(-2.0 * moment);
  double l1 = log(p), l0 = log(1.0 - p);
  double mx = NEG_INF, acc = 0.0;   /* streaming log-sum-exp */
  for (uint32_t st = 0; st < (1u << S); st++) {
    int agree = 0;
    for (int k = 0; k < ne; k++) { int a = (st >> e[k][0]) & 1, b = (st >> e[k][1]) & 1; agree += (a == b) ? 1 : -1; }
    int ones = __builtin_popcount(st);
    double ld = beta * coupling * agree + ones * l1 + (S - ones) * l0;
    if (ld > mx) { acc = acc * exp(mx - ld) + 1.0; mx = ld; } else acc += exp(ld - mx);
  }
  free(e); orc_model_free(m);
  return mx + log(acc);
}
