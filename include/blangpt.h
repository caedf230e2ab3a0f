/*
 * blangpt.h -- C ABI of libblangpt.so, the B200 (sm_100a) engine for Blang's
 * non-reversible parallel tempering (NRPT) scan.
 *
 * This is the drop-in boundary: exactly the entry points a Java shim
 * (`GpuPT implements PosteriorInferenceEngine`) binds through
 * JNI / Panama.  Plain pointers and sizes only; the library owns all device
 * memory behind the opaque handle; every host buffer is copied during the
 * call; no callbacks.  Not thread-safe: one host thread drives one handle
 * (the reference calls performInference once on the main thread,
 * R/runtime/Runner.xtend:238-240).  Every function returns 0 on success and a
 * BLANGPT_E* code otherwise; blangpt_last_error() gives the message the shim
 * turns into a RuntimeException (the reference's error convention,
 * Runner.xtend:164-168).  There is NO CPU fallback: without a CUDA device
 * blangpt_create fails.
 *
 * Citations: R/ = /root/reference/src/main/java/blang/.
 */
#ifndef BLANGPT_H
#define BLANGPT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLANGPT_VERSION 100

/* status codes */
enum {
  BLANGPT_OK = 0,
  BLANGPT_EINVAL = 1,       /* bad argument / unsupported option                         */
  BLANGPT_ECUDA = 2,        /* CUDA runtime error (message in last_error)                */
  BLANGPT_ESTATE = 3,       /* call order violated (e.g. run before set_model)           */
  BLANGPT_ENUMERIC = 4,     /* device-side condition that throws in Java: NaN factor
                               (ExponentiatedFactor.java:26-29), slice diverged to
                               infinity (RealSliceSampler.java:81-82,88-89), or no
                               acceptable integer in the slice window ("Invalid
                               arguments", Generators.java:205-211 via
                               IntSliceSampler.java:93-111)                              */
  BLANGPT_EUNSUPPORTED = 5  /* model / option outside the closed catalogue               */
};

/* model families: the closed catalogue the shim's check() recognises
 * (replaces per-factor virtual dispatch, R/runtime/SampledModel.java:258-282) */
enum {
  BLANGPT_FAM_NORMAL = 1,       /* data: y[n] f64.            hyper: m0, v0, rate        */
  BLANGPT_FAM_LOGISTIC = 2,     /* data: x[n*p] f64 row-major, y[n] i32. hyper: v0       */
  BLANGPT_FAM_MIXTURE = 3,      /* data: y[n] f64. hyper: K, m0, v0, gshape, grate, conc */
  BLANGPT_FAM_ISING = 4,        /* no data.       hyper: N, coupling, moment             */
  BLANGPT_FAM_BETABINOMIAL = 5, /* data: y[R*n] i32, ntrials[R*n] i32. hyper: rate       */
  BLANGPT_FAM_IID = 6,          /* y_i ~ Dist(theta) i.i.d., theta_j ~ prior_j.  data: y[n] f64 (+ trials[n] f64 for
                                   BINOMIAL).  hyper: dist, then (prior kind, a, b) for each parameter            */
  BLANGPT_FAM_HIERARCHICAL = 7, /* F/HierarchicalModel.bl, not marginalised: a, b ~ Exponential; p_g ~ Beta(a, b);
                                   failures_g ~ Binomial(launches_g, p_g).  data: failures[G] i32, launches[G] i32.
                                   hyper: rate_a, rate_b.  reals = a, b, p_0 .. p_{G-1}                            */
  BLANGPT_FAM_LINREG = 8,       /* F/LinRegression.bl with p covariates: w_j ~ Normal(0, v0), sigma ~ Uniform(0, smax),
                                   y_i ~ Normal(x_i . w, sigma^2).  data: x[n*p] f64 row-major, y[n] f64.
                                   hyper: v0, smax.  reals = w_0 .. w_{p-1}, sigma                                 */
  BLANGPT_FAM_ROCKETS_LATENT = 9,/* F/SimpleHierarchicalModel.bl: family 7 with latent launch counts L_g ~ Poisson(mean),
                                   p_g ~ Beta(a + 0.5, b + 0.5), failures_g ~ Binomial(1 + L_g, p_g); the L_g are updated
                                   by IntSliceSampler with stepping out.  data: failures[G] i32.  hyper: rate_a, rate_b,
                                   mean.  reals = a, b, p_0 .. p_{G-1}, L_0 .. L_{G-1} (the integers as exact doubles)   */
  BLANGPT_FAM_MIXTURE_INDICATORS = 10, /* blang.examples.MixtureModel (F/MixtureModel.bl) with K components and the cluster
                                   indicators NOT marginalised: z_i | pi ~ Categorical(pi) latent, moved by CategoricalSampler
                                   (R/mcmc/CategoricalSampler.xtend:24-28; K > 10 runs IntSliceSampler.accept :115-132).
                                   A small-model family: 3K + n <= 96, 2K + 1 + n <= 65.  data: y[n] f64.  hyper: K, m0, v0,
                                   gshape, grate, conc.  reals = mu[K], var[K], pi[K], z[n] (exact doubles)               */
  BLANGPT_FAM_ANNEAL_PAIR = 12, /* F/CustomAnnealRef.bl / F/CustomAnnealTest.bl (T/TestEndToEnd.xtend:169-186): mu ~ Normal(m0, v0),
                                   x | mu ~ Normal(mu, s2) annealed automatically (manual = 0) or through an AnnealingParameter
                                   (manual = 1, SampledModel.sumOtherAnnealed :219-225).  no data.  hyper: x, m0, v0, s2, manual */
  BLANGPT_FAM_CHANGEPOINT = 11  /* mu_0, mu_1 ~ Normal(m0, v0); c ~ DiscreteUniform(0, n + 1) moved by UniformSampler
                                   (R/mcmc/UniformSampler.xtend:21-27); y_i ~ Normal(i < c ? mu_0 : mu_1, s2).  data: y[n] f64.
                                   hyper: m0, v0, s2.  reals = mu_0, mu_1, c (exact double)                               */
};
/* family 6: observation laws (R/distributions/<Name>.bl) and their latent parameters, in this order */
enum {
  BLANGPT_IID_POISSON = 1,      /* mean                 */
  BLANGPT_IID_BINOMIAL = 2,     /* p                    */
  BLANGPT_IID_NEGBINOMIAL = 3,  /* r, p                 */
  BLANGPT_IID_STUDENTT = 4,     /* nu, mu, sigma        */
  BLANGPT_IID_GEOMETRIC = 5,    /* p                    */
  BLANGPT_IID_LAPLACE = 6,      /* location, scale      */
  BLANGPT_IID_EXPONENTIAL = 7,  /* rate                 */
  BLANGPT_IID_GAMMA = 8,        /* shape, rate          */
  BLANGPT_IID_WEIBULL = 9,      /* scale, shape         */
  BLANGPT_IID_GUMBEL = 10,      /* location, scale      */
  BLANGPT_IID_LOGISTIC = 11,    /* location, scale      (Logistic.bl, the distribution) */
  BLANGPT_IID_HALFSTUDENTT = 12,/* nu, sigma            */
  BLANGPT_IID_LOGLOGISTIC = 13, /* scale, shape         */
  BLANGPT_IID_GOMPERTZ = 14,    /* shape, scale         */
  BLANGPT_IID_YULESIMON = 15,   /* rho                  */
  BLANGPT_IID_UNIFORM_MAX = 16  /* max: y_i ~ ContinuousUniform(0.0, max), the likelihood of F/Doomsday.bl */
};
/* family 6: priors; (a, b) = (mean, variance) | (shape, rate) | (rate, unused) | (alpha, beta) | (min, max) */
enum {
  BLANGPT_PRIOR_NORMAL = 0, BLANGPT_PRIOR_GAMMA = 1, BLANGPT_PRIOR_EXPONENTIAL = 2, BLANGPT_PRIOR_BETA = 3,
  BLANGPT_PRIOR_UNIFORM = 4
};

/* PT.InitType (R/engines/internals/factories/PT.java:267) */
enum { BLANGPT_INIT_COPIES = 0, BLANGPT_INIT_FORWARD = 1, BLANGPT_INIT_SCM = 2 };

enum { BLANGPT_F64 = 0, BLANGPT_I32 = 1 };

typedef struct {
  const void* data;   /* host pointer */
  int64_t n_elem;
  int32_t dtype;      /* BLANGPT_F64 | BLANGPT_I32 */
  int32_t rows, cols; /* optional shape hint (row-major) */
} blangpt_array;

/* Engine options.  Field names follow the reference's @Arg names
 * (R/engines/ParallelTempering.java:25-40, PT.java:47-82, Runner.xtend:69-73). */
typedef struct {
  int32_t nChains;                 /* default 8                                        */
  int32_t nReplicas;               /* independent PT replicas in this handle (1)       */
  uint64_t random;                 /* seed, default 1                                   */
  double nPassesPerScan;           /* default 3                                        */
  int32_t usePriorSamples;         /* default 1                                        */
  int32_t reversible;              /* default 0                                        */
  int32_t annealSupport;           /* default 1                                        */
  int32_t treatNaNAsNegativeInfinity; /* default 0                                     */
  int32_t device;                  /* CUDA ordinal                                     */
  int32_t rank, worldSize;         /* chain sharding: this process owns a contiguous
                                      block of chain slots (nReplicas must be 1 when
                                      worldSize > 1)                                   */
  int32_t ctasPerChain;            /* 0 = auto; 1,2,4,8 = thread-block cluster size     */
  int64_t replicaOffset;           /* global index of local replica 0 (RNG keying)      */
  int32_t statisticRecordedMaxChainIndex; /* PT.java:76-78; default 100 (outputs only)  */
  int32_t reserved;
} blangpt_config;

typedef struct blangpt_handle blangpt_handle;

/* fills *cfg with the reference defaults */
void blangpt_default_config(blangpt_config* cfg);

int blangpt_create(const blangpt_config* cfg, blangpt_handle** out);
void blangpt_destroy(blangpt_handle* h);
const char* blangpt_last_error(const blangpt_handle* h);   /* h may be NULL: last create error */

/* Lower a recognised model (what GpuPT.setSampledModel does after walking the
 * SampledModel once, R/runtime/SampledModel.java:106-148).  Arrays in the
 * family's documented order. */
int blangpt_set_model(blangpt_handle* h, int32_t family, const blangpt_array* arrays, int32_t n_arrays,
                      const double* hyper, int32_t n_hyper);
int blangpt_n_reals(const blangpt_handle* h);      /* latent reals per chain, declaration order */
int blangpt_n_ints(const blangpt_handle* h);
int blangpt_n_samplers(const blangpt_handle* h);   /* SampledModel.nPosteriorSamplers :155-158  */
int blangpt_ctas_per_chain(const blangpt_handle* h); /* thread-block cluster size chosen for this model */
/* which exploration kernel serves this model: 0 = one thread-block cluster per chain (kernels.cuh),
 * 1 = the persistent pool kernel in which every SM serves every chain (pool.cuh; big-data families with
 * enough chains; BLANGPT_POOL=0/1 in the environment overrides the choice at set_model time) */
int blangpt_kernel_variant(const blangpt_handle* h);

/* Ladder, descending, betas[replica][position]; position 0 is beta = 1
 * (ParallelTempering.setAnnealingParameters :183-205: sorts, requires betas[0]==1,
 * resets every accumulator).  n == nChains broadcasts one ladder to all replicas. */
int blangpt_set_ladder(blangpt_handle* h, const double* betas_desc, int64_t n);
int blangpt_get_ladder(const blangpt_handle* h, double* out, int64_t n);

/* States are addressed by temperature POSITION like ParallelTempering.states[i].
 * reals[nReals], ints[nInts].  Only positions whose slot is local to this rank may be
 * written/read when worldSize > 1 (others return EINVAL). */
int blangpt_set_state(blangpt_handle* h, int64_t replica, int32_t position, const double* reals, const int32_t* ints);
int blangpt_get_state(blangpt_handle* h, int64_t replica, int32_t position, double* reals, int32_t* ints);
/* PT.informedInitialization :284-329.  COPIES keeps the states as set; FORWARD draws every chain from the
 * prior; SCM (the reference's default) runs an annealed SMC (R/engines/AdaptiveJarzynski.java:91-233,
 * schedules/AdaptiveTemperatureSchedule.java:48-78) with nParticles particles and hands one particle, drawn
 * by weight, to each ladder point. */
int blangpt_initialize(blangpt_handle* h, int32_t init_type);
/* PT.scmInit options (PT.java:68-72: nParticles 100, temperatureSchedule.threshold 0.9; resamplingESSThreshold 0.5) */
int blangpt_set_scm_options(blangpt_handle* h, int32_t nParticles, double threshold, double resamplingESSThreshold);
/* after an SCM initialisation: population.logNormEstimate() (PT.java:322-327), temperatures visited, resamplings */
int blangpt_scm_summary(const blangpt_handle* h, double* logNormEstimate, int32_t* nTemperatures, int32_t* nResamplings);

/* Parity hook: annealed log-densities of the CURRENT states, per position:
 * out_logdens = SampledModel.logDensity() at the chain's own beta (:161-176),
 * out_loglik = preAnnealedLogLikelihood (:209-212), out_noos = nOutOfSupport (:214-217),
 * out_fixed = sum of the fixed (prior) factors.  Arrays [nReplicas*nChains], nullable. */
int blangpt_log_density(blangpt_handle* h, double* out_logdens, double* out_loglik, int32_t* out_noos,
                        double* out_fixed);

/* Per-scan outputs of blangpt_run_scans (all nullable; sized for n_scans):
 * what PT.recordEnergyStatistics / recordSamples / swapAndRecordStatistics write
 * every scan (PT.java:116-143,360-385). */
typedef struct {
  double* energy;        /* [n_scans][nReplicas][nChains]  -preAnnealedLogLikelihood after the move */
  double* logDensity;    /* [n_scans][nReplicas][nChains]  allLogDensities                          */
  int32_t* nOutOfSupport;/* [n_scans][nReplicas][nChains]                                           */
  int32_t* swapIndicators;/*[n_scans][nReplicas][nChains]  1 at the lower index of a swapped pair   */
  double* samples;       /* [n_scans][nReplicas][nReals]   beta=1 chain, before the swap            */
  int32_t* intSamples;   /* [n_scans][nReplicas][nInts]                                             */
} blangpt_scan_outputs;

/* n_scans x { moveKernel ; record ; swapKernel }  (PT.java:93-102).  Single-GPU path. */
int blangpt_run_scans(blangpt_handle* h, int32_t n_scans, const blangpt_scan_outputs* out);

/* Round statistics (accumulators are per temperature position and per round,
 * ParallelTempering.java:197-204).  All arrays [nReplicas][nChains] (pairs use
 * the first nChains-1 entries), nullable. */
typedef struct {
  int64_t nScansInRound;
  double* swapAcceptMean;     /* swapAcceptPrs[i].getMean()                  */
  double* energyMean;         /* energies[c].getMean()                       */
  double* energyVariance;     /* energies[c].getVariance() (n-1)             */
  double* energyCovariance;   /* CovarAccumulator.sampleCovariance (C/n)     */
  double* logSum;             /* LogSumAccumulator.logSum                    */
  int64_t* logSumN;           /* LogSumAccumulator.numberOfTerms             */
  int64_t* roundTrips;        /* [nReplicas] Paths.nRejuvenations            */
  double* Lambda;             /* [nReplicas] PT.java:434                     */
  double* logZSteppingStone;  /* [nReplicas] ParallelTempering.java:143-153  */
  double* logZThermodynamic;  /* [nReplicas] :130-141, NaN if support annealed */
} blangpt_round_stats;
int blangpt_get_round_stats(blangpt_handle* h, blangpt_round_stats* out);

/* PT.adapt (PT.java:148-171): new ladder from the round's mean swap acceptance
 * (EngineStaticUtils.java:95-168, Spline.java:129-203), then resets the accumulators.
 * cum_lambda_xy (nullable): [nReplicas][2][nChains] knots (ascending beta, cumulative Lambda). */
int blangpt_adapt(blangpt_handle* h, double* cum_lambda_xy);
/* What blangpt_adapt would install, without installing it.  The reference adapts once more after the last round
 * "for instrumentation" (PT.java:103-107); here the final round's accumulators stay readable and this call returns
 * that ladder ([nReplicas][nChains], descending) and, if cum_lambda_xy is not NULL, the knots. */
int blangpt_adapt_preview(blangpt_handle* h, double* out_ladder, double* cum_lambda_xy);
/* reset accumulators without changing the ladder */
int blangpt_reset_round(blangpt_handle* h);

/* Initial ladders (R/engines/internals/ladders/): kind 0 EquallySpaced (:9-18), 1 Geometric (param =
 * annealingScaling, default 0.8), 2 Polynomial (param = power, default 3), 3 UserSpecified (user list, optional
 * spline generalisation to another nChains; FromAnotherExec = UserSpecified on the parsed annealingParameters.csv).
 * Pure host function.  out[nChains] descending.  EINVAL where the reference throws. */
int blangpt_make_ladder(int32_t kind, int32_t nChains, double param, const double* user, int32_t n_user,
                        int32_t allowSplineGeneralization, double* out);

/* Per-round ptanalysis outputs of PT.reportParallelTemperingDiagnostics (PT.java:387-505) and
 * reportLambdaFunctions (:192-210), computed from the current round's accumulators of one replica. */
typedef struct {
  int64_t nScansInRound;
  double swapLowest, swapAverage;                       /* swapSummaries                               */
  double energyCorrelationAverage, energyCorrelationHighest; /* energyExplCorrelationSummaries          */
  double Lambda;                                        /* globalLambda = sum_i (1 - mean accept_i)    */
  double inefficiency;                                  /* sum_i (1 - a_i) / a_i                       */
  double timeToFirstRestart, effectiveNScans;           /* timeToFirstRestart                          */
  int64_t temperedRestarts; double temperedRestartRate; /* actualTemperedRestarts (round trips)        */
  double asymptoticRoundTripCount, asymptoticRoundTripRate;       /* 1 / (2 + 2 Lambda)                */
  double nonAsymptoticRoundTripCount, nonAsymptoticRoundTripRate; /* 1 / (2 + 2 inefficiency)          */
  double logNormalizationSteppingStone, logNormalizationThermodynamic;
  double cumulativeLambda[99];                          /* Lambda(beta), beta = 0.01 .. 0.99           */
  double lambdaInstantaneous[99];                       /* lambda(beta) = dLambda/dbeta                */
} blangpt_diagnostics;
int blangpt_get_diagnostics(blangpt_handle* h, int64_t replica, blangpt_diagnostics* out);

/* PT.adapt with --engine.targetAccept (PT.java:156-163; EngineStaticUtils.targetAcceptancePartition :75-83,
 * fixedSizeOptimalPartition :176-198): the ladder, RESIZED to max(2, int(Lambda / (1 - targetAccept))) points by
 * Pegasus root finding on the cumulative-Lambda spline of the current round.  The caller then re-creates the
 * engine with *n_out chains, every chain a copy of the beta = 1 state, as the reference does. */
int blangpt_target_accept_ladder(blangpt_handle* h, int64_t replica, double targetAccept, double* out_desc,
                                 int32_t capacity, int32_t* n_out);

/* PT.rounds (PT.java:239-265).  Arrays sized >= 64.  Returns the number of rounds. */
int blangpt_rounds(int32_t nScans, double adaptFraction, int32_t* round_nscans, int32_t* round_is_adapt);

/* PT.performInference (PT.java:85-113) in one call; per-round diagnostics
 * (reportParallelTemperingDiagnostics :387-505) for replica 0..nReplicas-1.
 * BUFFER SIZES: the per-scan outputs are written for EVERY scan of EVERY round, i.e. for the SUM of the
 * blangpt_rounds(nScans, adaptFraction) round sizes, which can exceed nScans by one (the reference's own quirk:
 * 1000 scans at adaptFraction 0.5 run 2+4+8+16+31+63+126+251+500 = 1001 scans, PT.java:239-265).  Size `out` for
 * that sum, not for nScans.  `thinning` is accepted for symmetry with the reference's option and ignored: every scan
 * is recorded and the caller thins when it writes (PT.java:98-99 does `if (scanIndex % thinning == 0) recordSamples`). */
typedef struct {
  int32_t nRounds;
  int32_t roundNScans[64];
  int32_t roundIsAdapt[64];
  /* [64][nReplicas] arrays, caller-allocated, nullable */
  double* Lambda;
  double* logZ;
  int64_t* roundTrips;
  double* minAccept;
  double* meanAccept;
  double roundMillis[64];
} blangpt_inference_summary;
int blangpt_perform_inference(blangpt_handle* h, int32_t nScans, double adaptFraction, int32_t thinning,
                              const blangpt_scan_outputs* out, blangpt_inference_summary* summary);

/* Work counters since create: [0] log-density evaluations (one per sampler logDensity()
 * call, RealSliceSampler.java:156-161) [1] scans [2] kernel launches [3] sampler executions */
int blangpt_counters(blangpt_handle* h, uint64_t out[4]);

/* ---- split-phase calls for the multi-GPU path (one process per GPU) ----
 * scan = blangpt_explore -> all-gather of the local table segment (NCCL, by the
 * caller, on `stream`) -> blangpt_swap on the replicated table.  Every rank takes the
 * same swap decisions from the same Philox key, so no second message is needed. */
int blangpt_set_stream(blangpt_handle* h, void* cuda_stream);
/* evaluation-only pass over the local chains (after set_state / initialize): fills the
 * local table segment; the caller all-gathers it and then calls blangpt_prime once. */
int blangpt_evaluate(blangpt_handle* h);
int blangpt_prime(blangpt_handle* h);
int blangpt_explore(blangpt_handle* h);
/* which: 0 = full table [nChains][4] f64 (loglik, noos, logprior, evals),
 *        1 = local send segment [localChains][4]                              */
void* blangpt_table_ptr(blangpt_handle* h, int32_t which);
int64_t blangpt_local_chains(const blangpt_handle* h, int64_t* first_slot);
/* DEO swap pass on the replicated table.  `out` (nullable): this scan's energy / logDensity / nOutOfSupport /
 * swapIndicators [nReplicas][nChains] and samples [nReplicas][nReals]; samples are filled by the rank that owns the
 * replica's beta = 1 chain and are NaN elsewhere.  With `out` the call synchronises; with NULL it only enqueues. */
int blangpt_swap(blangpt_handle* h, const blangpt_scan_outputs* out_one_scan_or_null);
int blangpt_synchronize(blangpt_handle* h);

/* Multi-GPU exchange without a collective library (replaces the thread pool over chains of
 * R/engines/ParallelTempering.java:68,79 across GPUs).  Every rank owns a mailbox in its GPU's memory; once the
 * mailboxes are connected, the exploration kernels STORE each chain's 32-byte record (loglik, nOutOfSupport, fixed,
 * evaluations) and the beta = 1 sample into every rank's mailbox over NVLink, and blangpt_run_scans,
 * blangpt_perform_inference, blangpt_initialize (SCM included) and blangpt_log_density work on a sharded handle
 * (worldSize > 1) exactly as on one GPU, with no host code between the scans.  All ranks must make the same
 * sequence of calls.  One process per GPU: exchange the 64-byte handles of blangpt_peer_export by any means
 * (torch.distributed, MPI, a file) and pass them, in rank order, to blangpt_peer_connect.  One process driving
 * several GPUs (one JVM): pass the device pointers of blangpt_peer_mailbox to blangpt_peer_connect_pointers and
 * drive each handle from its own host thread. */
int blangpt_peer_export(blangpt_handle* h, void* out_handle64);
int blangpt_peer_connect(blangpt_handle* h, const void* handles /* worldSize x 64 bytes */, int32_t n);
void* blangpt_peer_mailbox(blangpt_handle* h);
int blangpt_peer_connect_pointers(blangpt_handle* h, void* const* mailboxes, const int32_t* devices, int32_t n);

/* Measured compute ceilings on `device` (no handle needed): which = 0 -> fp64 FMA TFLOP/s,
 * which = 1 -> logistic-regression row terms per second with operands in registers (the
 * compute-only bound of the C2 likelihood kernel). */
int blangpt_microbench(int32_t device, int32_t which, double* out_rate);
/* Accuracy hook: the per-row Bernoulli(logistic(z)) log-likelihood term exactly as the C2 kernels
 * evaluate it (F/SpikedGLM.bl:26-35 -> Bernoulli.bl:11-14 -> Categorical.bl:12-15); -inf = out of support. */
int blangpt_debug_logit_terms(int32_t device, const double* z, const int32_t* y, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif
