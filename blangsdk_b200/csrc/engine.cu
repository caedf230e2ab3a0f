// engine.cu -- host side of libblangpt.so: the C ABI of include/blangpt.h.
//
// Mirrors the reference's engine objects (same option names and semantics):
//   blangpt_config            <- ParallelTempering @Arg fields (R/engines/ParallelTempering.java:25-40)
//   blangpt_run_scans         <- the scan loop of PT.performInference (R/engines/internals/factories/PT.java:93-102)
//   blangpt_get_round_stats   <- reportParallelTemperingDiagnostics (PT.java:387-505)
//   blangpt_adapt             <- PT.adapt (PT.java:148-171)
//   blangpt_perform_inference <- PT.performInference (PT.java:85-113)
// There is no CPU path: every compute call launches the sm_100a kernels of kernels.cuh.
#include "../../include/blangpt.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <chrono>
#include <algorithm>
#include <cmath>

#include "kernels.cuh"
#include "ising.cuh"
#include "schedule.hpp"
#include "pool_host.hpp"

using namespace bpt;

static thread_local std::string g_create_error;

struct blangpt_handle {
  blangpt_config cfg;
  std::string error;
  int family = 0;
  int n_reals = 0, n_ints = 0, n_samplers = 0;
  int G = 1;                      // CTAs per chain
  bool logistic_l2ahead = false;  // eta cache > L2: use the kernel variant with L2 prefetch hints
  int logistic_threads = 0;       // CTA size of the C2 kernel: 0 = choose in set_model, or BLANGPT_LOGISTIC_THREADS=512|1024
  bool model_set = false, tables_valid = false;
  cudaStream_t stream = nullptr;  // nullptr = legacy default stream
  double* swap_out = nullptr; int* swap_out_i = nullptr;   // one scan's outputs of blangpt_swap (split-phase path)
  bool own_stream = false;
  bool l2_persist = false;       // launches carry an L2 access-policy window: persisting lines are released in destroy

  DevEngine E{};
  DevStats St{};
  // family data
  NormalFam::Data normal{};
  IidFam::Data iid{};
  HierFam::Data hier{};
  LinRegFam::Data linreg{};
  RocketsLatentFam::Data rockets{};
  LogisticFam::Data logistic{};
  MixtureFam<4>::Data mixture{};
  BetaBinomialFam::Data betabin{};
  MixIndFam::Data mixind{};
  ChangePointFam::Data changepoint{};
  AnnealPairFam::Data annealpair{};
  IsingData ising{};
  std::vector<void*> allocs;      // every device allocation (freed in destroy)
  // device-side staging of run_scans' per-scan outputs: kept across calls and grown on demand (a cudaMalloc /
  // cudaFree pair per output per call costs a device synchronisation each)
  struct OutBuf { void* p = nullptr; size_t bytes = 0; } outbuf[6];
  std::vector<double> ladder;     // host copy [R][N]
  uint64_t n_scans = 0, n_launches = 0;
  int64_t scans_in_round = 0;
  int64_t M_total = 0;
  // SCM initialisation (PT.scmInit defaults, PT.java:68-72,507-515)
  int scm_particles = 100; double scm_threshold = 0.9, scm_resampling_ess = 0.5;
  double scm_log_norm = std::nan(""); int scm_n_temperatures = 0, scm_n_resamplings = 0;
  // multi-GPU exchange (kernels.cuh PeerBox): this rank's mailbox allocation and the opened peer mappings
  void* mailbox = nullptr; size_t mailbox_bytes = 0; size_t mail_record_bytes = 0;
  std::vector<void*> peer_maps; bool peers_ready = false;
  bool one_view() const { return cfg.worldSize == 1 || peers_ready; }   // this process sees every chain's record
  PoolState* pool = nullptr;      // pool exploration kernel (pool.cuh): every SM serves every chain; null = not used
};

#define FAIL(h, code, ...) do { set_error((h), __VA_ARGS__); return (code); } while (0)
#define CUDA_OK(h, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    set_error((h), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return BLANGPT_ECUDA; } } while (0)

static void set_error(blangpt_handle* h, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->error = buf; else g_create_error = buf;
}

// Device memory of a handle comes from the device's stream-ordered pool (cudaMallocAsync on the handle's stream), with
// the pool told to keep what is freed: a create -> run -> destroy cycle then costs no device synchronisation and no
// trip to the OS per buffer (a plain cudaMalloc / cudaFree pair per buffer cost 10-25 ms per engine).  The mailbox of the
// multi-GPU exchange stays a cudaMalloc allocation (CUDA IPC needs one).
static void keep_freed_memory(int device) {
  static bool done[64] = {};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[device] = true;
}
static void dev_free(blangpt_handle* h, void* p) { if (p) cudaFreeAsync(p, h->stream); }

template <class T>
static int dev_alloc(blangpt_handle* h, T** out, size_t n, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  keep_freed_memory(h->cfg.device);
  CUDA_OK(h, cudaMallocAsync(&p, bytes, h->stream));
  if (zero) CUDA_OK(h, cudaMemsetAsync(p, 0, bytes, h->stream));
  h->allocs.push_back(p);
  *out = (T*)p;
  return 0;
}

extern "C" void blangpt_default_config(blangpt_config* c) {
  memset(c, 0, sizeof *c);
  c->nChains = 8; c->nReplicas = 1; c->random = 1; c->nPassesPerScan = 3.0;
  c->usePriorSamples = 1; c->reversible = 0; c->annealSupport = 1; c->treatNaNAsNegativeInfinity = 0;
  c->device = 0; c->rank = 0; c->worldSize = 1; c->ctasPerChain = 0; c->replicaOffset = 0;
  c->statisticRecordedMaxChainIndex = 100;
}

extern "C" const char* blangpt_last_error(const blangpt_handle* h) {
  return h ? h->error.c_str() : g_create_error.c_str();
}

extern "C" int blangpt_create(const blangpt_config* cfg, blangpt_handle** out) {
  if (!cfg || !out) { set_error(nullptr, "null argument"); return BLANGPT_EINVAL; }
  if (cfg->nChains < 1) { set_error(nullptr, "Number of tempering chains must be greater than zero."); return BLANGPT_EINVAL; }
  if (!(cfg->nPassesPerScan >= 0.0) || std::isinf(cfg->nPassesPerScan)) { set_error(nullptr, "nPassesPerScan must be a finite number >= 0"); return BLANGPT_EINVAL; }
  if (cfg->nReplicas < 1 || cfg->worldSize < 1 || cfg->rank < 0 || cfg->rank >= cfg->worldSize) { set_error(nullptr, "bad nReplicas/rank/worldSize"); return BLANGPT_EINVAL; }
  if (cfg->worldSize > 1 && cfg->nReplicas != 1) { set_error(nullptr, "chain sharding (worldSize > 1) requires nReplicas == 1; shard replicas with one handle per rank"); return BLANGPT_EINVAL; }
  if (cfg->worldSize > 1 && cfg->nChains % cfg->worldSize != 0) { set_error(nullptr, "nChains must be a multiple of worldSize"); return BLANGPT_EINVAL; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(nullptr, "no CUDA device: libblangpt has no CPU fallback (%s)", cudaGetErrorString(e));
    return BLANGPT_ECUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { set_error(nullptr, "device %d out of range", cfg->device); return BLANGPT_EINVAL; }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) { set_error(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e)); return BLANGPT_ECUDA; }
  int cc_major = 0, cc_minor = 0;      // two attribute reads; cudaGetDeviceProperties costs 3-14 ms per call
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device);
  cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, cfg->device);
  if (cc_major < 10) { set_error(nullptr, "device sm_%d%d is not Blackwell (kernels are built for sm_100a only)", cc_major, cc_minor); return BLANGPT_ECUDA; }
  blangpt_handle* h = new blangpt_handle();
  h->cfg = *cfg;
  if (const char* t = getenv("BLANGPT_LOGISTIC_THREADS")) h->logistic_threads = atoi(t) == 1024 ? 1024 : 512;
  h->M_total = (int64_t)cfg->nReplicas * cfg->nChains;
  *out = h;
  return BLANGPT_OK;
}

extern "C" void blangpt_destroy(blangpt_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : h->peer_maps) if (p) cudaIpcCloseMemHandle(p);
  if (h->mailbox) cudaFree(h->mailbox);
  for (void* p : h->allocs) dev_free(h, p);
  for (auto& b : h->outbuf) if (b.p) cudaFree(b.p);
  cudaStreamSynchronize(h->stream);
  if (h->l2_persist) cudaCtxResetPersistingL2Cache();
  pool_destroy(h->pool); h->pool = nullptr;
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int blangpt_n_reals(const blangpt_handle* h) { return h ? h->n_reals : -1; }
extern "C" int blangpt_n_ints(const blangpt_handle* h) { return h ? h->n_ints : -1; }
extern "C" int blangpt_n_samplers(const blangpt_handle* h) { return h ? h->n_samplers : -1; }
extern "C" int blangpt_ctas_per_chain(const blangpt_handle* h) { return h ? h->G : -1; }
extern "C" int blangpt_kernel_variant(const blangpt_handle* h) { return h ? (h->pool ? 1 : 0) : -1; }

// ---------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------
// families may size the CTA from the data (launch_threads(D) <= kThreads, the kernel's launch bound)
template <class Fam>
static auto threads_for(const typename Fam::Data& D, int) -> decltype(Fam::launch_threads(D)) { return Fam::launch_threads(D); }
template <class Fam>
static int threads_for(const typename Fam::Data&, long) { return Fam::kThreads; }

template <class Fam>
static int launch_explore_t(blangpt_handle* h, const typename Fam::Data& D, int mode) {
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)(h->E.M_local * h->G));
  lc.blockDim = dim3((unsigned)threads_for<Fam>(D, 0));
  lc.dynamicSmemBytes = 0;
  lc.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)h->G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = h->G > 1 ? 1 : 0;
  CUDA_OK(h, cudaLaunchKernelEx(&lc, explore_kernel<Fam>, h->E, D, h->G, mode));
  h->n_launches++;
  return 0;
}

static int launch_explore_local(blangpt_handle* h, int mode) {
  switch (h->family) {
    case BLANGPT_FAM_NORMAL: return launch_explore_t<NormalFam>(h, h->normal, mode);
    case BLANGPT_FAM_IID: return launch_explore_t<IidFam>(h, h->iid, mode);
    case BLANGPT_FAM_HIERARCHICAL: return launch_explore_t<HierFam>(h, h->hier, mode);
    case BLANGPT_FAM_LINREG: return launch_explore_t<LinRegFam>(h, h->linreg, mode);
    case BLANGPT_FAM_ROCKETS_LATENT: return launch_explore_t<RocketsLatentFam>(h, h->rockets, mode);
    case BLANGPT_FAM_LOGISTIC:
      if (h->pool) {   // persistent pool kernel: all SMs serve all chains (pool.cuh)
        const int prc = pool_launch_logistic(h->pool, h->E, h->logistic, mode, h->stream);
        if (prc != 0) FAIL(h, BLANGPT_ECUDA, "pool launch failed: %s", cudaGetErrorString((cudaError_t)prc));
        h->n_launches++;
        return 0;
      }
      if (h->logistic_l2ahead) {   // per-chain caches larger than the L2: they stream from HBM.  1024 threads per CTA
        // (8 warps per scheduler hide the latency: 313 vs 356 ms per scan at 1024 chains) plus L2 prefetch hints
        LogisticFamT<1024, 1, 3>::Data d2; memcpy(&d2, &h->logistic, sizeof d2);
        return launch_explore_t<LogisticFamT<1024, 1, 3>>(h, d2, mode);
      }
      if (h->logistic_threads == 1024) {
        LogisticFamT<1024>::Data d2; memcpy(&d2, &h->logistic, sizeof d2);
        return launch_explore_t<LogisticFamT<1024>>(h, d2, mode);
      }
      return launch_explore_t<LogisticFam>(h, h->logistic, mode);
    case BLANGPT_FAM_MIXTURE:
      if (h->pool) {   // persistent pool kernel (pool_mixture.cuh)
        const int prc = pool_launch_mixture(h->pool, h->E, h->mixture, mode, h->stream);
        if (prc != 0) FAIL(h, BLANGPT_ECUDA, "pool launch failed: %s", cudaGetErrorString((cudaError_t)prc));
        h->n_launches++;
        return 0;
      }
      if (h->mixture.K <= 4) return launch_explore_t<MixtureFam<4>>(h, h->mixture, mode);
      else { MixtureFam<8>::Data d8; memcpy(&d8, &h->mixture, sizeof d8); return launch_explore_t<MixtureFam<8>>(h, d8, mode); }
    case BLANGPT_FAM_BETABINOMIAL: return launch_explore_t<BetaBinomialFam>(h, h->betabin, mode);
    case BLANGPT_FAM_MIXTURE_INDICATORS: return launch_explore_t<MixIndFam>(h, h->mixind, mode);
    case BLANGPT_FAM_CHANGEPOINT: return launch_explore_t<ChangePointFam>(h, h->changepoint, mode);
    case BLANGPT_FAM_ANNEAL_PAIR: return launch_explore_t<AnnealPairFam>(h, h->annealpair, mode);
    case BLANGPT_FAM_ISING: {
      int rc = ising_launch(h->E, h->ising, mode, h->stream);
      if (rc != 0) FAIL(h, BLANGPT_ECUDA, "ising launch failed: %s", cudaGetErrorString((cudaError_t)rc));
      h->n_launches += 1;
      return 0;
    }
  }
  FAIL(h, BLANGPT_ESTATE, "no model set");
}

// One exploration pass of the local chains; with the multi-GPU exchange on, every chain's record is published to all
// ranks by the kernel itself and gather_kernel completes this rank's replicated table on the same stream.
static int launch_explore(blangpt_handle* h, int mode) {
  if (h->peers_ready) h->E.peers.xseq++;
  int rc = launch_explore_local(h, mode);
  if (rc || !h->peers_ready) return rc;
  const int64_t items = h->M_total + (int64_t)h->cfg.nReplicas * h->n_reals;
  gather_kernel<<<(unsigned)std::min<int64_t>(64, (items + 127) / 128), 128, 0, h->stream>>>(h->E);
  CUDA_OK(h, cudaGetLastError());
  h->n_launches++;
  return 0;
}

// the fused replica kernel applies to warp-capable families with small data, <= 32 chains, one process
static bool use_fused(const blangpt_handle* h) {
  if (h->cfg.worldSize != 1 || h->cfg.nChains > 32) return false;
  const char* env = getenv("BLANGPT_FUSED");
  if (env && atoi(env) == 0) return false;
  const bool forced = env && atoi(env) == 1;
  if (h->family == BLANGPT_FAM_BETABINOMIAL) return forced || h->cfg.nReplicas >= 8;
  if (h->family == BLANGPT_FAM_NORMAL) return h->normal.n <= 65536 && (forced || h->cfg.nReplicas >= 8);
  if (h->family == BLANGPT_FAM_IID) return h->iid.n <= 65536 && (forced || h->cfg.nReplicas >= 8);
  if (h->family == BLANGPT_FAM_HIERARCHICAL || h->family == BLANGPT_FAM_ROCKETS_LATENT) return forced || h->cfg.nReplicas >= 8;
  return false;
}

static int launch_swap(blangpt_handle* h, const DevOut& O) {
  const int threads = std::min(256, std::max(32, ((h->cfg.nChains + 31) / 32) * 32));
  swap_kernel<<<h->cfg.nReplicas, threads, 0, h->stream>>>(h->E, h->St, O);
  CUDA_OK(h, cudaGetLastError());
  h->n_launches++;
  return 0;
}

static int check_device_errors(blangpt_handle* h) {
  int err = 0;
  CUDA_OK(h, cudaMemcpyAsync(&err, h->E.err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  if (err) {   // reported once: the handle is usable again after the caller repairs the state (set_state / initialize)
    CUDA_OK(h, cudaMemsetAsync(h->E.err, 0, sizeof(int), h->stream));
    h->tables_valid = false;
  }
  if (err & ERR_SLICE_DIVERGED)
    FAIL(h, BLANGPT_ENUMERIC, "Slice diverged to infinity. Possible cause is that a variable has no distribution attached to it, i.e. the model is improper.");
  if (err & ERR_INT_SLICE_EMPTY)
    FAIL(h, BLANGPT_ENUMERIC, "Invalid arguments: the integer slice sampler found no point of non-zero density in its window (the state it started from has zero density; Generators.discreteUniform throws in the reference). Initialise with SCM or from a feasible state.");
  if (err & ERR_PEER_TIMEOUT)
    FAIL(h, BLANGPT_ECUDA, "multi-GPU exchange: a peer's chain records did not arrive (a rank failed, or the ranks did not make the same sequence of calls)");
  if (err & ERR_POOL_TIMEOUT)
    FAIL(h, BLANGPT_ECUDA, "pool kernel: a wait inside the persistent exploration kernel timed out (aborted instead of hanging)");
  if ((err & ERR_NAN) && !h->cfg.treatNaNAsNegativeInfinity)
    FAIL(h, BLANGPT_ENUMERIC, "Factors should not return NaN. Use NEGATIVE_INFINITY to forbid configurations. If you are sure this NaN is OK, you can use the option 'treatNaNAsNegativeInfinity'.");
  return 0;
}

static int ensure_tables(blangpt_handle* h) {
  if (h->tables_valid) return 0;
  int rc = launch_explore(h, MODE_EVAL);
  if (rc) return rc;
  if (h->one_view()) {
    prime_prev_kernel<<<(unsigned)((h->M_total + 255) / 256), 256, 0, h->stream>>>(h->E, h->St);
    CUDA_OK(h, cudaGetLastError());
    h->n_launches++;
  }
  h->tables_valid = h->one_view();
  return 0;
}

static int reset_round(blangpt_handle* h) {
  reset_round_kernel<<<h->cfg.nReplicas, 128, 0, h->stream>>>(h->E, h->St);
  CUDA_OK(h, cudaGetLastError());
  h->n_launches++;
  h->scans_in_round = 0;
  return 0;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
// row-major [n][p] -> column-major [p][npad]; rows with y == 0 are negated (sign folding, exact)
__global__ void transpose_kernel(const double* __restrict__ in, const uint8_t* __restrict__ y, double* __restrict__ out, int n, int p, int npad) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy, c = c0 + threadIdx.x;
    const double v = (r < n && c < p) ? in[(size_t)r * p + c] : 0.0;
    tile[dy][threadIdx.x] = (r < n && y[r] == 0) ? -v : v;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, r = r0 + threadIdx.x;
    if (c < p && r < npad) out[(size_t)c * npad + r] = tile[threadIdx.x][dy];
  }
}

// Initial slot <-> position map.  One process: identity.  Sharded (W ranks own contiguous blocks of slots): position p
// starts on slot (p % W) * (N / W) + p / W, so every rank holds an even spread of the ladder from the first scan on --
// colder chains need ~25 % more evaluations per scan, and with the identity map rank 0 would hold the whole cold half
// until the particles have mixed.  Results are keyed by position (Philox streams, statistics), so they do not depend
// on this map; the index process counts round trips of particles, whatever their labels.
__global__ void init_perm_kernel(DevEngine E, int s_max_used, int world) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E.R * E.N) {
    const int p = i % E.N, base = i - p;
    const int s = world > 1 ? (p % world) * (E.N / world) + p / world : p;
    E.slot_of_pos[base + p] = s; E.pos_of_slot[base + s] = p;
  }
  if (i < E.M_local) {
    E.curpos[i] = -1;
    for (int k = 0; k < s_max_used; k++) E.order[(size_t)i * E.s_max + k] = k;
  }
}

static int alloc_engine(blangpt_handle* h) {
  const blangpt_config& c = h->cfg;
  DevEngine& E = h->E;
  const int64_t M = h->M_total;
  E.N = c.nChains; E.R = c.nReplicas;
  E.M_local = (int)(M / c.worldSize);
  E.slot_lo = c.rank * E.M_local;
  E.key0 = (uint32_t)c.random; E.key1 = (uint32_t)(c.random >> 32);
  E.replica_offset = (uint32_t)c.replicaOffset;
  E.n_passes = c.nPassesPerScan; E.use_prior = c.usePriorSamples; E.anneal_support = c.annealSupport;
  E.nan_neginf = c.treatNaNAsNegativeInfinity;
  E.beta_override = std::nan(""); E.steps_override = 0;
  E.s_max = h->family == BLANGPT_FAM_ISING ? 1 : std::max(1, h->n_samplers);
  double* ladder; int rc;
  if ((rc = dev_alloc(h, &ladder, M))) return rc;
  E.ladder = ladder;
  if ((rc = dev_alloc(h, &E.pos_of_slot, M))) return rc;
  if ((rc = dev_alloc(h, &E.slot_of_pos, M))) return rc;
  if ((rc = dev_alloc(h, &E.table, M * 4))) return rc;
  if ((rc = dev_alloc(h, &E.send, (size_t)E.M_local * 4))) return rc;
  if ((rc = dev_alloc(h, &E.prev_loglik, M))) return rc;
  if ((rc = dev_alloc(h, &E.reals, (size_t)std::max(1, h->n_reals) * E.M_local))) return rc;
  if ((rc = dev_alloc(h, &E.order, (size_t)E.M_local * E.s_max))) return rc;
  if ((rc = dev_alloc(h, &E.curpos, E.M_local))) return rc;
  if ((rc = dev_alloc(h, &E.evals, E.M_local))) return rc;
  if ((rc = dev_alloc(h, &E.execs, E.M_local))) return rc;
  if ((rc = dev_alloc(h, &E.err, 1))) return rc;
  if ((rc = dev_alloc(h, &E.scan_counter, c.nReplicas))) return rc;
  if ((rc = dev_alloc(h, &E.swap_counter, c.nReplicas))) return rc;
  if ((rc = dev_alloc(h, &E.batch_counter, c.nReplicas))) return rc;
  DevStats& S = h->St;
  if ((rc = dev_alloc(h, &S.en_n, M))) return rc;
  if ((rc = dev_alloc(h, &S.en_m1, M))) return rc;
  if ((rc = dev_alloc(h, &S.en_m2, M))) return rc;
  if ((rc = dev_alloc(h, &S.cov_mx, M))) return rc;
  if ((rc = dev_alloc(h, &S.cov_my, M))) return rc;
  if ((rc = dev_alloc(h, &S.cov_c, M))) return rc;
  if ((rc = dev_alloc(h, &S.cov_n, M))) return rc;
  if ((rc = dev_alloc(h, &S.acc_n, M))) return rc;
  if ((rc = dev_alloc(h, &S.acc_m1, M))) return rc;
  if ((rc = dev_alloc(h, &S.ls_sum, M))) return rc;
  if ((rc = dev_alloc(h, &S.ls_n, M))) return rc;
  if ((rc = dev_alloc(h, &S.path_flag, M))) return rc;
  if ((rc = dev_alloc(h, &S.rejuv, c.nReplicas))) return rc;
  if ((rc = dev_alloc(h, &S.total_evals, 1))) return rc;
  if ((rc = dev_alloc(h, &S.oos_seen, c.nReplicas))) return rc;
  if (c.worldSize > 1) {   // this rank's mailbox for the multi-GPU exchange (its own allocation: it is exported through CUDA IPC)
    h->mail_record_bytes = (size_t)2 * M * 2 * sizeof(MailHalf);
    h->mailbox_bytes = h->mail_record_bytes + (size_t)2 * c.nReplicas * std::max(1, h->n_reals) * sizeof(MailHalf);
    CUDA_OK(h, cudaMalloc(&h->mailbox, h->mailbox_bytes));
    CUDA_OK(h, cudaMemsetAsync(h->mailbox, 0, h->mailbox_bytes, h->stream));
    if ((rc = dev_alloc(h, &E.peers.sample_buf, (size_t)c.nReplicas * std::max(1, h->n_reals)))) return rc;
  }
  init_perm_kernel<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(E, h->family == BLANGPT_FAM_ISING ? 0 : h->n_samplers, c.worldSize);
  CUDA_OK(h, cudaGetLastError());
  // default ladder: EquallySpaced
  std::vector<double> one = equally_spaced(c.nChains);
  h->ladder.resize(M);
  for (int r = 0; r < c.nReplicas; r++) std::copy(one.begin(), one.end(), h->ladder.begin() + (size_t)r * c.nChains);
  CUDA_OK(h, cudaMemcpyAsync((void*)E.ladder, h->ladder.data(), M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return reset_round(h);
}

// Cluster size per chain: one wave when the chains fit (widest cluster that keeps >= 8k rows per CTA),
// otherwise the smallest cluster whose last wave is >= 95 % full (chains are independent, so clusters of
// different chains need not be co-resident).
static int pick_ctas_per_chain(const blangpt_handle* h, int64_t rows) {
  if (h->cfg.ctasPerChain > 0) return h->cfg.ctasPerChain;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
  const int64_t local = h->M_total / h->cfg.worldSize;
  int G = 1;
  if (local <= sms) {
    while (G < 8 && local * (G * 2) <= sms && rows / (G * 2) >= 8192) G *= 2;
    return G;
  }
  for (int g = 1; g <= 8; g *= 2) {
    if (g > 1 && rows / g < 8192) break;
    const double waves = (double)(local * g) / sms;
    G = g;
    if (std::ceil(waves) / waves <= 1.05) break;
  }
  return G;
}

extern "C" int blangpt_set_model(blangpt_handle* h, int32_t family, const blangpt_array* a, int32_t na,
                                 const double* hyper, int32_t nh) {
  if (!h) return BLANGPT_EINVAL;
  if (h->model_set) FAIL(h, BLANGPT_ESTATE, "model already set");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const blangpt_config& c = h->cfg;
  h->family = family;
  int rc;
  int64_t rows = 1;
  switch (family) {
    case BLANGPT_FAM_NORMAL: {
      if (na != 1 || nh != 3 || a[0].dtype != BLANGPT_F64) FAIL(h, BLANGPT_EINVAL, "normal: expects y[f64], hyper (m0, v0, rate)");
      const int n = (int)a[0].n_elem;
      double* y; if ((rc = dev_alloc(h, &y, n + 2))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      h->normal = {y, n, hyper[0], hyper[1], hyper[2]};
      h->n_reals = 2; h->n_samplers = 2; rows = n;
    } break;
    case BLANGPT_FAM_ROCKETS_LATENT: {
      if (na != 1 || nh != 3 || a[0].dtype != BLANGPT_I32) FAIL(h, BLANGPT_EINVAL, "rockets_latent: expects failures[i32 G], hyper (rate_a, rate_b, pois_mean)");
      const int G = (int)a[0].n_elem;
      if (G < 1 || 2 * G + 2 > kMaxParams) FAIL(h, BLANGPT_EUNSUPPORTED, "rockets_latent: 1 <= groups <= %d", (kMaxParams - 2) / 2);
      if (!(hyper[2] > 0.0 && hyper[2] < 40.0)) FAIL(h, BLANGPT_EUNSUPPORTED, "rockets_latent: 0 < pois_mean < 40");
      int* y;
      if ((rc = dev_alloc(h, &y, G))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, G * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      h->rockets = {y, G, hyper[0], hyper[1], hyper[2]};
      h->n_reals = 2 * G + 2; h->n_samplers = 2 * G + 2; rows = 1;
    } break;
    case BLANGPT_FAM_LINREG: {
      if (na != 2 || nh != 2 || a[0].dtype != BLANGPT_F64 || a[1].dtype != BLANGPT_F64) FAIL(h, BLANGPT_EINVAL, "linreg: expects x[f64 n*p], y[f64 n], hyper (v0, smax)");
      const int n = (int)a[1].n_elem;
      if (n <= 0 || a[0].n_elem % n != 0) FAIL(h, BLANGPT_EINVAL, "linreg: x size is not a multiple of n");
      const int p = (int)(a[0].n_elem / n);
      if (p + 1 > kMaxParams) FAIL(h, BLANGPT_EUNSUPPORTED, "linreg: at most %d coefficients", kMaxParams - 1);
      const int npad = (n + 1) & ~1;
      double *xr, *xt, *yd, *res; uint8_t* ones;
      if ((rc = dev_alloc(h, &xr, (size_t)n * p, false))) return rc;
      if ((rc = dev_alloc(h, &xt, (size_t)npad * p))) return rc;
      if ((rc = dev_alloc(h, &yd, npad))) return rc;
      if ((rc = dev_alloc(h, &ones, npad, false))) return rc;
      const int64_t local = h->M_total / c.worldSize;
      if ((rc = dev_alloc(h, &res, (size_t)npad * local))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(xr, a[0].data, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaMemcpyAsync(yd, a[1].data, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaMemsetAsync(ones, 1, npad, h->stream));        // no sign folding: every row keeps its sign
      dim3 grid((n + 31) / 32, (p + 31) / 32), block(32, 8);
      transpose_kernel<<<grid, block, 0, h->stream>>>(xr, ones, xt, n, p, npad);
      CUDA_OK(h, cudaGetLastError());
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
      dev_free(h, xr); h->allocs.erase(std::find(h->allocs.begin(), h->allocs.end(), (void*)xr));
      dev_free(h, ones); h->allocs.erase(std::find(h->allocs.begin(), h->allocs.end(), (void*)ones));
      h->linreg = {xt, yd, res, n, npad, p, hyper[0], hyper[1]};
      h->n_reals = p + 1; h->n_samplers = p + 1; rows = n;
    } break;
    case BLANGPT_FAM_HIERARCHICAL: {
      if (na != 2 || nh != 2 || a[0].dtype != BLANGPT_I32 || a[1].dtype != BLANGPT_I32 || a[0].n_elem != a[1].n_elem)
        FAIL(h, BLANGPT_EINVAL, "hierarchical: expects failures[i32 G], launches[i32 G], hyper (rate_a, rate_b)");
      const int G = (int)a[0].n_elem;
      if (G < 1 || G + 2 > kMaxParams) FAIL(h, BLANGPT_EUNSUPPORTED, "hierarchical: 1 <= groups <= %d", kMaxParams - 2);
      int *y, *nt;
      if ((rc = dev_alloc(h, &y, G)) || (rc = dev_alloc(h, &nt, G))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, G * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaMemcpyAsync(nt, a[1].data, G * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      h->hier = {y, nt, G, hyper[0], hyper[1]};
      h->n_reals = G + 2; h->n_samplers = G + 2; rows = 1;
    } break;
    case BLANGPT_FAM_IID: {
      const int dist = nh >= 1 ? (int)hyper[0] : 0;
      if (dist < IID_POISSON || dist >= IID_N_DISTS) FAIL(h, BLANGPT_EUNSUPPORTED, "iid: unknown distribution code %d", dist);
      const int P = iid_shape(dist).n_params;
      if (nh != 1 + 3 * P) FAIL(h, BLANGPT_EINVAL, "iid: hyper must be (dist, then kind, a, b for each of the %d parameters)", P);
      const int want_arrays = dist == IID_BINOMIAL ? 2 : 1;
      if (na != want_arrays || a[0].dtype != BLANGPT_F64 || (na == 2 && (a[1].dtype != BLANGPT_F64 || a[1].n_elem != a[0].n_elem)))
        FAIL(h, BLANGPT_EINVAL, "iid: expects y[f64 n]%s", dist == IID_BINOMIAL ? ", trials[f64 n]" : "");
      const int n = (int)a[0].n_elem;
      double *y, *tr = nullptr;
      if ((rc = dev_alloc(h, &y, n + 2))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      if (na == 2) {
        if ((rc = dev_alloc(h, &tr, n + 2))) return rc;
        CUDA_OK(h, cudaMemcpyAsync(tr, a[1].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      }
      h->iid = IidFam::Data{};
      h->iid.y = y; h->iid.trials = tr; h->iid.n = n; h->iid.dist = dist; h->iid.n_params = P;
      for (int j = 0; j < P; j++) {
        const int kind = (int)hyper[1 + 3 * j];
        if (kind < PRIOR_NORMAL || kind > PRIOR_UNIFORM) FAIL(h, BLANGPT_EUNSUPPORTED, "iid: unknown prior code %d", kind);
        h->iid.prior_kind[j] = kind; h->iid.prior_a[j] = hyper[2 + 3 * j]; h->iid.prior_b[j] = hyper[3 + 3 * j];
      }
      h->n_reals = P; h->n_samplers = P; rows = n;
    } break;
    case BLANGPT_FAM_LOGISTIC: {
      if (na != 2 || nh != 1 || a[0].dtype != BLANGPT_F64 || a[1].dtype != BLANGPT_I32) FAIL(h, BLANGPT_EINVAL, "logistic: expects x[f64 n*p], y[i32 n], hyper (v0)");
      const int n = (int)a[1].n_elem;
      if (n <= 0 || a[0].n_elem % n != 0) FAIL(h, BLANGPT_EINVAL, "logistic: x size is not a multiple of n");
      const int p = (int)(a[0].n_elem / n);
      if (p > kMaxParams) FAIL(h, BLANGPT_EUNSUPPORTED, "logistic: at most %d coefficients", kMaxParams);
      const int npad = (n + 1) & ~1;
      double *xr, *xt, *eta; uint8_t* y8;
      if ((rc = dev_alloc(h, &xr, (size_t)n * p, false))) return rc;
      if ((rc = dev_alloc(h, &xt, (size_t)npad * p))) return rc;
      if ((rc = dev_alloc(h, &y8, npad))) return rc;
      const int64_t local = h->M_total / c.worldSize;
      if ((rc = dev_alloc(h, &eta, (size_t)npad * local))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(xr, a[0].data, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      std::vector<uint8_t> yb(npad, 0);
      const int32_t* yi = (const int32_t*)a[1].data;
      for (int i = 0; i < n; i++) yb[i] = yi[i] == 1 ? 1 : 0;
      CUDA_OK(h, cudaMemcpyAsync(y8, yb.data(), npad, cudaMemcpyHostToDevice, h->stream));
      dim3 grid((n + 31) / 32, (p + 31) / 32), block(32, 8);
      transpose_kernel<<<grid, block, 0, h->stream>>>(xr, y8, xt, n, p, npad);
      CUDA_OK(h, cudaGetLastError());
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
      dev_free(h, xr); h->allocs.erase(std::find(h->allocs.begin(), h->allocs.end(), (void*)xr));
      h->logistic = {xt, y8, eta, n, npad, p, hyper[0]};
      h->n_reals = p; h->n_samplers = p; rows = n;
    } break;
    case BLANGPT_FAM_MIXTURE: {
      if (na != 1 || nh != 6 || a[0].dtype != BLANGPT_F64) FAIL(h, BLANGPT_EINVAL, "mixture: expects y[f64], hyper (K, m0, v0, gshape, grate, conc)");
      const int n = (int)a[0].n_elem, K = (int)hyper[0];
      if (K < 2 || K > 8) FAIL(h, BLANGPT_EUNSUPPORTED, "mixture: 2 <= K <= 8");
      double* y; if ((rc = dev_alloc(h, &y, n + 2))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      const int npad = (n + 1) & ~1;
      const int64_t local = h->M_total / c.worldSize;
      double *rest, *e1, *e2;
      if ((rc = dev_alloc(h, &rest, (size_t)npad * local, false))) return rc;
      if ((rc = dev_alloc(h, &e1, (size_t)npad * local, false))) return rc;
      if ((rc = dev_alloc(h, &e2, (size_t)npad * local, false))) return rc;
      h->mixture = {y, n, K, hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], rest, e1, e2, npad};
      h->n_reals = 3 * K; h->n_samplers = 2 * K + 1; rows = n;
    } break;
    case BLANGPT_FAM_MIXTURE_INDICATORS: {
      if (na != 1 || nh != 6 || a[0].dtype != BLANGPT_F64) FAIL(h, BLANGPT_EINVAL, "mixture_indicators: expects y[f64], hyper (K, m0, v0, gshape, grate, conc)");
      const int n = (int)a[0].n_elem, K = (int)hyper[0];
      if (K < 2 || K > 16) FAIL(h, BLANGPT_EUNSUPPORTED, "mixture_indicators: 2 <= K <= 16");
      if (3 * K + n > kMaxParams + 32 || 2 * K + 1 + n > kMaxSamplers)
        FAIL(h, BLANGPT_EUNSUPPORTED, "mixture_indicators is a small-model family: 3K + n <= %d and 2K + 1 + n <= %d (marginalise the indicators -- family MIXTURE -- for large n)", kMaxParams + 32, kMaxSamplers);
      double* y; if ((rc = dev_alloc(h, &y, n + 2))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      h->mixind = {y, n, K, hyper[1], hyper[2], hyper[3], hyper[4], hyper[5]};
      h->n_reals = 3 * K + n; h->n_samplers = 2 * K + 1 + n; rows = 1;
    } break;
    case BLANGPT_FAM_ANNEAL_PAIR: {
      if (na != 0 || nh != 5) FAIL(h, BLANGPT_EINVAL, "anneal_pair: no data arrays, hyper (x, m0, v0, s2, manual)");
      h->annealpair = {hyper[0], hyper[1], hyper[2], hyper[3], hyper[4] != 0.0 ? 1 : 0};
      h->n_reals = 1; h->n_samplers = 1; rows = 1;
    } break;
    case BLANGPT_FAM_CHANGEPOINT: {
      if (na != 1 || nh != 3 || a[0].dtype != BLANGPT_F64) FAIL(h, BLANGPT_EINVAL, "changepoint: expects y[f64], hyper (m0, v0, s2)");
      const int n = (int)a[0].n_elem;
      double* y; if ((rc = dev_alloc(h, &y, n + 2))) return rc;
      CUDA_OK(h, cudaMemcpyAsync(y, a[0].data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      h->changepoint = {y, n, hyper[0], hyper[1], hyper[2]};
      h->n_reals = 3; h->n_samplers = 3; rows = n;
    } break;
    case BLANGPT_FAM_BETABINOMIAL: {
      if (na != 2 || nh != 1 || a[0].dtype != BLANGPT_I32 || a[1].dtype != BLANGPT_I32) FAIL(h, BLANGPT_EINVAL, "betabinomial: expects y[i32 R*n], ntrials[i32 R*n], hyper (rate)");
      if (a[0].n_elem % c.nReplicas != 0 || a[0].n_elem != a[1].n_elem) FAIL(h, BLANGPT_EINVAL, "betabinomial: data size must be nReplicas * n");
      const int n = (int)(a[0].n_elem / c.nReplicas);
      const size_t tot = (size_t)a[0].n_elem;
      int32_t *y, *nt; double* cst;
      if ((rc = dev_alloc(h, &y, tot))) return rc;
      if ((rc = dev_alloc(h, &nt, tot))) return rc;
      if ((rc = dev_alloc(h, &cst, tot))) return rc;
      std::vector<double> cs(tot);
      const int32_t *yh = (const int32_t*)a[0].data, *nh_ = (const int32_t*)a[1].data;
      for (size_t i = 0; i < tot; i++) {   // BetaBinomial.bl:19-20 support checks + the three data-only lnGamma terms
        const double x = yh[i], n_ = nh_[i];
        cs[i] = (x < 0.0 || n_ <= 0.0 || x > n_) ? -INFINITY : lgamma(n_ + 1) - lgamma(x + 1) - lgamma(n_ - x + 1);
      }
      CUDA_OK(h, cudaMemcpyAsync(y, yh, tot * 4, cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaMemcpyAsync(nt, nh_, tot * 4, cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaMemcpyAsync(cst, cs.data(), tot * 8, cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
      h->betabin = {y, nt, cst, n, c.nChains, hyper[0]};
      h->n_reals = 2; h->n_samplers = 2; rows = n;
    } break;
    case BLANGPT_FAM_ISING: {
      if (nh != 3) FAIL(h, BLANGPT_EINVAL, "ising: hyper (N, coupling, moment)");
      const int N = (int)hyper[0];
      if (N < 2 || (N & 1)) FAIL(h, BLANGPT_EUNSUPPORTED, "ising: the checkerboard sweep needs an even lattice side >= 2");
      // the reference runs floor(nPassesPerScan * N*N) single-spin steps (SampledModel.java:249-256); the checkerboard
      // kernel runs whole sweeps, so a fractional number of passes would silently be truncated
      if (c.nPassesPerScan != std::floor(c.nPassesPerScan)) FAIL(h, BLANGPT_EUNSUPPORTED, "ising: nPassesPerScan must be a whole number (checkerboard sweeps); got %g", c.nPassesPerScan);
      h->ising.N = N; h->ising.coupling = hyper[1]; h->ising.moment = hyper[2];
      h->ising.p1 = 1.0 / (1.0 + exp(2.0 * hyper[2]));   // logistic(-2*moment), F/Ising.bl:28
      h->n_reals = 0; h->n_ints = N * N; h->n_samplers = N * N; rows = 1;
      const int64_t local = h->M_total / c.worldSize;
      if ((rc = dev_alloc(h, &h->ising.spins, (size_t)local * N * N))) return rc;
    } break;
    default: FAIL(h, BLANGPT_EUNSUPPORTED, "unknown model family %d: the engine only runs the closed catalogue (no CPU fallback)", family);
  }
  h->G = (family == BLANGPT_FAM_ISING || family == BLANGPT_FAM_MIXTURE_INDICATORS) ? 1 : pick_ctas_per_chain(h, rows);
  // (256 threads x two CTAs per SM x twice the cluster, and 768-thread CTAs, were measured and dropped: DESIGN.md 5.1)
  if (family == BLANGPT_FAM_LOGISTIC) {
    // CTA shape: 512 threads when a cluster of CTAs serves the chain (C2: 2 x 512), 1024 when one CTA does (more
    // chains than half the SMs: 128 chains 53.8 -> 46.0 ms per scan, 1024 chains 356 -> 313 ms).  When the chains'
    // caches plus X exceed ~3/4 of the L2 they stream from HBM and the kernel also issues L2 prefetch hints three
    // steps ahead (128 chains 46.0 -> 41.5 ms, 1024 chains 313 -> 300 ms).
    if (h->logistic_threads == 0) h->logistic_threads = (h->G == 1 && rows >= 32768) ? 1024 : 512;
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, c.device);
    const size_t cache_bytes = (size_t)h->logistic.npad * (h->M_total / c.worldSize) * sizeof(double);
    const size_t x_bytes = (size_t)h->logistic.npad * h->logistic.p * sizeof(double);
    h->logistic_l2ahead = h->logistic_threads == 1024 && l2 > 0 && (cache_bytes + x_bytes) > (size_t)l2 / 4 * 3;
    if (const char* t = getenv("BLANGPT_L2AHEAD")) h->logistic_l2ahead = atoi(t) != 0;
    // Pool kernel (pool.cuh) when there are enough chains to keep every SM busy with other chains' rows while one
    // chain's request is in flight; an explicit ctasPerChain asks for the cluster kernel.  BLANGPT_POOL=0/1 overrides.
    // Measured on the C2 shape (profiles/configs_r02.jsonl), M evals/s pool | cluster: 64 chains 3.98 | 2.78, 128: 2.93 | 2.58,
    // 256: 2.78 | 2.72, 512: 2.60 | 3.13, 1024: 2.42 | 2.88 -- beyond 256 chains per GPU one CTA per chain keeps every
    // SM busy by itself and the controllers' SMs are better spent on rows.
    const int64_t local_chains = h->M_total / c.worldSize;
    bool want_pool = local_chains >= 24 && local_chains <= 256 && rows >= 16384 && c.ctasPerChain == 0;
    if (const char* t = getenv("BLANGPT_POOL")) want_pool = atoi(t) != 0;
    if (want_pool) {
      const int prc = pool_create(&h->pool, (int)local_chains, c.device, BLANGPT_FAM_LOGISTIC, 0);
      if (prc > 0) FAIL(h, BLANGPT_ECUDA, "pool setup failed: %s", cudaGetErrorString((cudaError_t)prc));
      // Caches + X just beyond the L2 (128 chains: 102 + 26 MB against 126 MB): plain LRU then misses on nearly every
      // cache row.  An access-policy window over the caches lets a fraction of their lines persist in the L2 and the
      // rest stream past them; the fraction leaves room for X (hot, shared by all chains).  Measured at 128 chains:
      // 2.88 M evals/s without, 3.34 / 3.46 / 3.44 / 2.97 M at ratios 0.4 / 0.6 / 0.65 / 0.9; at 256 chains (205 MB of
      // caches, beyond the 128 MB a window may span) nothing is gained, so the window is used only when it covers them.
      if (h->pool && l2 > 0 && cache_bytes + x_bytes > (size_t)l2 / 10 * 9) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
        double ratio = cache_bytes <= (size_t)max_window ? (0.66 * l2 - (double)x_bytes) / (double)cache_bytes : 0.0;
        ratio = std::min(ratio, (double)max_persist / (double)std::max<size_t>(1, cache_bytes));
        if (const char* t = getenv("BLANGPT_L2PERSIST")) ratio = atof(t);
        if (ratio > 0.0 && max_persist > 0 && max_window > 0) {
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
          const size_t win = std::min(cache_bytes, (size_t)max_window);
          pool_set_l2_window(h->pool, h->logistic.eta, win, (float)std::min(1.0, ratio));
          h->l2_persist = true;
          if (getenv("BLANGPT_VERBOSE")) fprintf(stderr, "[blangpt] L2 window: %zu MB of caches, hit ratio %.2f, persisting limit %d MB, max window %d MB\n", win >> 20, ratio, max_persist >> 20, max_window >> 20);
        }
      }
    }
  }
  if (family == BLANGPT_FAM_MIXTURE) {   // pool kernel under the same rule as the logistic family
    const int64_t local_chains = h->M_total / c.worldSize;
    bool want_pool = local_chains >= 24 && rows >= 16384 && c.ctasPerChain == 0;
    if (const char* t = getenv("BLANGPT_POOL")) want_pool = atoi(t) != 0;
    if (want_pool) {
      const int prc = pool_create(&h->pool, (int)local_chains, c.device, BLANGPT_FAM_MIXTURE, h->mixture.K);
      if (prc > 0) FAIL(h, BLANGPT_ECUDA, "pool setup failed: %s", cudaGetErrorString((cudaError_t)prc));
    }
  }
  if (h->G != 1 && h->G != 2 && h->G != 4 && h->G != 8) FAIL(h, BLANGPT_EINVAL, "ctasPerChain must be 1, 2, 4 or 8");
  if ((rc = alloc_engine(h))) return rc;
  // initial state: reals as in the oracle's prototypes (latent scalars start at their constructor value)
  std::vector<double> init((size_t)std::max(1, h->n_reals) * h->E.M_local, 0.0);
  auto fill = [&](int v, double x) { for (int l = 0; l < h->E.M_local; l++) init[(size_t)v * h->E.M_local + l] = x; };
  if (family == BLANGPT_FAM_NORMAL) { fill(0, 0.0); fill(1, 1.0); }
  if (family == BLANGPT_FAM_BETABINOMIAL) { fill(0, 1.0); fill(1, 1.0); }
  if (family == BLANGPT_FAM_HIERARCHICAL) { fill(0, 1.0); fill(1, 1.0); for (int g = 0; g < h->hier.G; g++) fill(2 + g, 0.5); }
  if (family == BLANGPT_FAM_ROCKETS_LATENT) { fill(0, 1.0); fill(1, 1.0); for (int g = 0; g < h->rockets.G; g++) fill(2 + g, 0.5); }
  if (family == BLANGPT_FAM_MIXTURE_INDICATORS) { const int K = h->mixind.K; for (int k = 0; k < K; k++) { fill(k, 0.0); fill(K + k, 1.0); fill(2 * K + k, 1.0 / K); } }
  if (family == BLANGPT_FAM_MIXTURE) { const int K = h->mixture.K; for (int k = 0; k < K; k++) { fill(k, 0.0); fill(K + k, 1.0); fill(2 * K + k, 1.0 / K); } }
  if (h->n_reals) CUDA_OK(h, cudaMemcpyAsync(h->E.reals, init.data(), init.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->model_set = true; h->tables_valid = false;
  return BLANGPT_OK;
}

// ---------------------------------------------------------------------------
// ladder / states
// ---------------------------------------------------------------------------
extern "C" int blangpt_set_ladder(blangpt_handle* h, const double* betas, int64_t n) {
  if (!h || !betas) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int N = h->cfg.nChains, R = h->cfg.nReplicas;
  if (n != N && n != (int64_t)N * R) FAIL(h, BLANGPT_EINVAL, "ladder length must be nChains or nReplicas*nChains");
  for (int r = 0; r < R; r++) {
    const double* src = betas + (n == N ? 0 : (size_t)r * N);
    std::vector<double> v(src, src + N);
    std::sort(v.begin(), v.end(), [](double a, double b) { return a > b; });
    // a lone chain may sit at any exponent (SampledModel.setExponent, as factories/Pigeons.java:56-64 uses it)
    if (v[0] != 1.0 && N > 1) FAIL(h, BLANGPT_EINVAL, "the ladder must start at exactly 1.0 (ParallelTempering.java:191-192)");
    if (N == 1 && !(v[0] >= 0.0 && v[0] <= 1.0)) FAIL(h, BLANGPT_EINVAL, "the exponent must be in [0,1]");
    std::copy(v.begin(), v.end(), h->ladder.begin() + (size_t)r * N);
  }
  CUDA_OK(h, cudaMemcpyAsync((void*)h->E.ladder, h->ladder.data(), h->ladder.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return reset_round(h);
}

extern "C" int blangpt_get_ladder(const blangpt_handle* h, double* out, int64_t n) {
  if (!h || !out) return BLANGPT_EINVAL;
  if (n != (int64_t)h->ladder.size() && n != h->cfg.nChains) return BLANGPT_EINVAL;
  memcpy(out, h->ladder.data(), n * sizeof(double));
  return BLANGPT_OK;
}

static int slot_of(blangpt_handle* h, int64_t replica, int position, int* local) {
  if (replica < 0 || replica >= h->cfg.nReplicas || position < 0 || position >= h->cfg.nChains) FAIL(h, BLANGPT_EINVAL, "replica/position out of range");
  int s = 0;
  CUDA_OK(h, cudaMemcpyAsync(&s, h->E.slot_of_pos + replica * h->cfg.nChains + position, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  const int64_t g = replica * h->cfg.nChains + s;
  const int64_t l = g - h->E.slot_lo;
  if (l < 0 || l >= h->E.M_local) FAIL(h, BLANGPT_EINVAL, "position %d lives on another rank", position);
  *local = (int)l;
  return 0;
}

extern "C" int blangpt_set_state(blangpt_handle* h, int64_t replica, int32_t position, const double* reals, const int32_t* ints) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int l, rc;
  if ((rc = slot_of(h, replica, position, &l))) return rc;
  if (reals && h->n_reals)
    CUDA_OK(h, cudaMemcpy2DAsync(h->E.reals + l, (size_t)h->E.M_local * sizeof(double), reals, sizeof(double), sizeof(double), h->n_reals, cudaMemcpyHostToDevice, h->stream));
  if (ints && h->n_ints) {
    std::vector<uint8_t> b(h->n_ints);
    for (int i = 0; i < h->n_ints; i++) {
      if (ints[i] < 0 || ints[i] > 1) FAIL(h, BLANGPT_EINVAL, "ising spins must be 0 or 1");
      b[i] = (uint8_t)ints[i];
    }
    CUDA_OK(h, cudaMemcpyAsync(h->ising.spins + (size_t)l * h->n_ints, b.data(), h->n_ints, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->tables_valid = false;
  return BLANGPT_OK;
}

extern "C" int blangpt_get_state(blangpt_handle* h, int64_t replica, int32_t position, double* reals, int32_t* ints) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int l, rc;
  if ((rc = slot_of(h, replica, position, &l))) return rc;
  if (reals && h->n_reals)
    CUDA_OK(h, cudaMemcpy2DAsync(reals, sizeof(double), h->E.reals + l, (size_t)h->E.M_local * sizeof(double), sizeof(double), h->n_reals, cudaMemcpyDeviceToHost, h->stream));
  std::vector<uint8_t> b;
  if (ints && h->n_ints) {
    b.resize(h->n_ints);
    CUDA_OK(h, cudaMemcpyAsync(b.data(), h->ising.spins + (size_t)l * h->n_ints, h->n_ints, cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  if (ints && h->n_ints) for (int i = 0; i < h->n_ints; i++) ints[i] = b[i];
  return BLANGPT_OK;
}

// ---------------------------------------------------------------------------
// SCM initialisation: PT's default `initialization` (PT.java:300-329).  An annealed SMC
// (AdaptiveJarzynski.getApproximation, R/engines/AdaptiveJarzynski.java:91-137) with nParticles particles
// walks the temperature up; at each ladder point one particle, drawn by weight, becomes that chain's state.
// Particles are the chain slots of a child engine that shares the parent's device data; every particle runs
// at the same exponent (beta_override) and makes ONE sampler step per temperature (steps_override = 1,
// AdaptiveJarzynski.sampleNext :224-233).  The control logic -- adaptive temperature by a Pegasus search on
// the relative conditional ESS (schedules/AdaptiveTemperatureSchedule.java:48-78, EngineStaticUtils.java:21-47),
// stratified resampling below ESS 0.5, weight bookkeeping -- is O(nParticles) host arithmetic on the
// (loglik, nOOS, fixed) table, as it is host code in the reference.
// ---------------------------------------------------------------------------
namespace {
enum { P_MASTER = 6 };
struct ScmParticle { double loglik, noos, fixed; };

double scm_density(const ScmParticle& p, double beta, int anneal_support) {
  return annealed_log_density(p.fixed, p.loglik, p.noos, beta, anneal_support);
}
// SampledModel.logDensityRatio :199-206
bool scm_ratio(const ScmParticle& p, double t, double next, int as, double* out) {
  const double num = scm_density(p, next, as), den = scm_density(p, t, as);
  if (num == -INFINITY && den == -INFINITY) return false;
  *out = num - den;
  return true;
}
// EngineStaticUtils.relativeCESS :27-47
double scm_relative_cess(const std::vector<ScmParticle>& ps, const std::vector<double>& w, double t, double next, int as, bool* ok) {
  double lnum = -INFINITY, lden = -INFINITY;
  for (size_t i = 0; i < ps.size(); i++) {
    double r;
    if (!scm_ratio(ps[i], t, next, as, &r)) { *ok = false; return std::nan(""); }
    lnum = log_add(lnum, log(w[i]) + 1.0 * r);
    lden = log_add(lden, log(w[i]) + 2.0 * r);
  }
  return exp(2.0 * lnum - lden);
}
}  // namespace

static int scm_read_table(blangpt_handle* c, std::vector<ScmParticle>& ps) {
  const int P = c->cfg.nChains;
  std::vector<double> tab((size_t)P * 4);
  CUDA_OK(c, cudaMemcpyAsync(tab.data(), c->E.table, tab.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  int rc = check_device_errors(c);
  if (rc) return rc;
  ps.resize(P);
  for (int i = 0; i < P; i++) ps[i] = {tab[i * 4 + 0], tab[i * 4 + 1], tab[i * 4 + 2]};
  return 0;
}

// new particle i = old particle ancestors[i] (ParticlePopulation.resample + deepCopyParticles :149-170), on the
// device: the particle states never visit the host, whatever the number of particles and the size of a state.
// dst[r * row_stride + i * rec + b] = src[r * row_stride + anc[i] * rec + b], elements of T; variable-major arrays
// are `rows` rows of P one-element records, particle-major arrays one row of P records of `rec` elements.
template <class T>
__global__ void scm_gather_kernel(T* __restrict__ dst, const T* __restrict__ src, const int* __restrict__ anc, int P, size_t rec, size_t rows) {
  const size_t per_row = (size_t)P * rec, total = per_row * rows;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / per_row, q = t - r * per_row, i = q / rec, b = q - i * rec;
    dst[t] = src[r * per_row + (size_t)anc[i] * rec + b];
  }
}
template <class T>
static int scm_gather(blangpt_handle* c, T* arr, const int* d_anc, size_t rec, size_t rows) {
  const size_t total = (size_t)c->cfg.nChains * rec * rows;
  if (!total) return 0;
  T* tmp = nullptr;
  CUDA_OK(c, cudaMallocAsync((void**)&tmp, total * sizeof(T), c->stream));
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  scm_gather_kernel<T><<<grid, 256, 0, c->stream>>>(tmp, arr, d_anc, c->cfg.nChains, rec, rows);
  c->n_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(arr, tmp, total * sizeof(T), cudaMemcpyDeviceToDevice, c->stream);
  cudaFreeAsync(tmp, c->stream);
  if (e != cudaSuccess) { set_error(c, "SCM resampling gather failed: %s", cudaGetErrorString(e)); return BLANGPT_ECUDA; }
  return 0;
}
static int scm_permute(blangpt_handle* c, const std::vector<int>& anc) {
  const int P = c->cfg.nChains;
  int* d_anc = nullptr;
  CUDA_OK(c, cudaMallocAsync((void**)&d_anc, (size_t)P * sizeof(int), c->stream));
  int rc = 0;
  cudaError_t e = cudaMemcpyAsync(d_anc, anc.data(), (size_t)P * sizeof(int), cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) { set_error(c, "SCM resampling upload failed: %s", cudaGetErrorString(e)); rc = BLANGPT_ECUDA; }
  if (!rc && c->n_reals) rc = scm_gather(c, c->E.reals, d_anc, 1, (size_t)c->n_reals);          // reals[v][particle]
  if (!rc && c->n_ints) rc = scm_gather(c, c->ising.spins, d_anc, (size_t)c->n_ints, 1);          // spins[particle][S]
  // the sampler order travels with the state (SampledModel.currentSamplingOrder is deep-cloned)
  if (!rc) rc = scm_gather(c, c->E.order, d_anc, (size_t)c->E.s_max, 1);
  if (!rc) rc = scm_gather(c, c->E.curpos, d_anc, 1, 1);
  // anc is pageable host memory: the copy above has completed on return only for the staging part; wait before it dies
  if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { set_error(c, "SCM resampling failed on the device"); rc = BLANGPT_ECUDA; }
  cudaFreeAsync(d_anc, c->stream);
  if (rc) return rc;
  return launch_explore(c, MODE_EVAL);   // rebuild per-particle caches and the table
}

static int scm_initialize(blangpt_handle* h) {
  if (h->cfg.nReplicas != 1) FAIL(h, BLANGPT_EUNSUPPORTED, "SCM initialisation is implemented for one replica per handle");
  const int N = h->cfg.nChains, P = h->scm_particles, as = h->cfg.annealSupport;
  // chains at beta == 0 (and every chain before it is overwritten) start from the prior, PT.java:311-315
  int rc = blangpt_initialize(h, BLANGPT_INIT_FORWARD);
  if (rc) return rc;
  // ---- child engine: P particles sharing the parent's device data ----
  blangpt_handle* c = new blangpt_handle();
  c->cfg = h->cfg; c->cfg.nChains = P; c->cfg.nReplicas = 1; c->cfg.usePriorSamples = 0;
  c->cfg.worldSize = 1; c->cfg.rank = 0;   // every rank runs the same particle system (same seed) and keeps its own chains
  c->cfg.random = h->cfg.random ^ 0x53434D2D696E6974ULL;   // its own Philox key ("SCM-init")
  c->M_total = P; c->family = h->family; c->stream = h->stream; c->logistic_threads = h->logistic_threads;
  c->n_reals = h->n_reals; c->n_ints = h->n_ints; c->n_samplers = h->n_samplers;
  c->normal = h->normal; c->iid = h->iid; c->hier = h->hier; c->linreg = h->linreg; c->rockets = h->rockets; c->logistic = h->logistic; c->mixture = h->mixture; c->betabin = h->betabin; c->ising = h->ising; c->mixind = h->mixind; c->changepoint = h->changepoint; c->annealpair = h->annealpair;
  c->betabin.chains_per_replica = P;
  auto fail = [&](int code) { for (void* p : c->allocs) dev_free(c, p); cudaStreamSynchronize(c->stream); if (code) h->error = c->error; delete c; return code; };
  // CUDA calls on the child go through fail(): the child's allocations are released and its message reaches the caller
#define SCM_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    set_error(c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return fail(BLANGPT_ECUDA); } } while (0)
  if (h->family == BLANGPT_FAM_LOGISTIC && (rc = dev_alloc(c, &c->logistic.eta, (size_t)c->logistic.npad * P))) return fail(rc);
  if (h->family == BLANGPT_FAM_LINREG && (rc = dev_alloc(c, &c->linreg.res, (size_t)c->linreg.npad * P))) return fail(rc);
  if (h->family == BLANGPT_FAM_ISING && (rc = dev_alloc(c, &c->ising.spins, (size_t)P * h->n_ints))) return fail(rc);
  if (h->family == BLANGPT_FAM_MIXTURE) {
    const size_t sz = (size_t)c->mixture.npad * P;
    if ((rc = dev_alloc(c, &c->mixture.rest, sz, false)) || (rc = dev_alloc(c, &c->mixture.e1, sz, false)) ||
        (rc = dev_alloc(c, &c->mixture.e2, sz, false))) return fail(rc);
  }
  int64_t rows = h->family == BLANGPT_FAM_NORMAL ? h->normal.n : h->family == BLANGPT_FAM_LINREG ? h->linreg.n : h->family == BLANGPT_FAM_IID ? h->iid.n : h->family == BLANGPT_FAM_LOGISTIC ? h->logistic.n
               : h->family == BLANGPT_FAM_MIXTURE ? h->mixture.n : h->family == BLANGPT_FAM_CHANGEPOINT ? h->changepoint.n : 1;
  c->G = h->family == BLANGPT_FAM_ISING || h->family == BLANGPT_FAM_BETABINOMIAL || h->family == BLANGPT_FAM_MIXTURE_INDICATORS ? 1 : pick_ctas_per_chain(c, rows);
  if ((rc = alloc_engine(c))) return fail(rc);
  c->model_set = true;
  if ((rc = launch_explore(c, MODE_FORWARD_INIT))) return fail(rc);   // AdaptiveJarzynski.sampleInitial :214-222
  std::vector<ScmParticle> ps;
  if ((rc = scm_read_table(c, ps))) return fail(rc);
  std::vector<double> w(P, 1.0 / P);
  double log_scaling = 0.0, temperature = 0.0;
  uint32_t iter = 0, master = 0;
  int n_stalled = 0;
  const uint32_t k0 = (uint32_t)c->cfg.random, k1 = (uint32_t)(c->cfg.random >> 32);
  auto master_double = [&]() { Rng r(k0, k1, P_MASTER, 0u, master++, 0u); return r.next_double(); };
  h->scm_n_temperatures = 0; h->scm_n_resamplings = 0;
  c->E.steps_override = 1;

  for (int li = N - 1; li >= 0; li--) {            // ladder ascending; position li
    const double target = h->ladder[li];
    if (target == 0.0) continue;
    while (temperature < target) {
      // --- AdaptiveTemperatureSchedule.nextTemperature :48-70 ---
      bool ok = true;
      auto objective = [&](double x) { return scm_relative_cess(ps, w, temperature, x, as, &ok) - h->scm_threshold; };
      double next;
      const double at_one = objective(1.0);
      if (!ok) return fail((set_error(c, "Invalid logDensity ratio (0/0): this could be caused by a generate(rand){..} block not faithful with its laws{..} block."), BLANGPT_ENUMERIC));
      if (at_one != at_one) {
        // every particle is out of support: the reference re-proposes at max(temperature, 1e-10) until one comes back
        // (AdaptiveTemperatureSchedule.java:54-60); bounded here instead of looping for ever
        next = std::max(temperature, 1e-10);
        if (++n_stalled > 100000) return fail((set_error(c, "SCM initialisation: every particle stayed out of support for 100000 consecutive temperatures"), BLANGPT_ENUMERIC));
      } else {
        n_stalled = 0;
        next = objective(target) >= 0 ? target : pegasus_solve(objective, temperature, target, 1e-16, 100);
        if (!(next > temperature)) next = target;   // numerical stall of the root search (never reached in the tests)
      }
      // --- resamplingNeeded :139-147 / resample :149-154 (stratified) ---
      bool need = false;
      double ess_den = 0.0;
      for (int i = 0; i < P; i++) { if (w[i] == 0.0) need = true; ess_den += w[i] * w[i]; }
      const double rel_ess = 1.0 / (ess_den * P);
      if (need || (rel_ess < h->scm_resampling_ess && next < target)) {
        std::vector<int> anc(P);
        double cum = w[0]; int j = 0;
        for (int i = 0; i < P; i++) {
          const double u = (i + master_double()) / P;
          while (u > cum && j < P - 1) cum += w[++j];
          anc[i] = j;
        }
        if ((rc = scm_permute(c, anc))) return fail(rc);
        if ((rc = scm_read_table(c, ps))) return fail(rc);
        std::fill(w.begin(), w.end(), 1.0 / P);
        h->scm_n_resamplings++;
      }
      // --- propose :171-212: weights from the CURRENT states, then one sampler step at the next temperature ---
      std::vector<double> lw(P);
      double lsum = -INFINITY;
      for (int i = 0; i < P; i++) {
        double r;
        if (!scm_ratio(ps[i], temperature, next, as, &r)) return fail((set_error(c, "Invalid logDensity ratio (0/0)"), BLANGPT_ENUMERIC));
        lw[i] = r + log(w[i]);
        lsum = log_add(lsum, lw[i]);
      }
      c->E.beta_override = next;
      SCM_OK(cudaMemcpyAsync(c->E.scan_counter, &iter, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
      SCM_OK(cudaStreamSynchronize(c->stream));
      if ((rc = launch_explore(c, MODE_EXPLORE))) return fail(rc);
      if ((rc = scm_read_table(c, ps))) return fail(rc);
      for (int i = 0; i < P; i++) w[i] = exp(lw[i] - lsum);   // ParticlePopulation.buildDestructivelyFromLogWeights
      log_scaling += lsum;
      temperature = next;
      iter++; h->scm_n_temperatures++;
    }
    // --- population.sample(random).duplicate() becomes the chain at this ladder point, PT.java:318-320 ---
    const double u = master_double();
    double cum = 0.0; int pick = P - 1;
    for (int i = 0; i < P; i++) { cum += w[i]; if (u < cum) { pick = i; break; } }
    std::vector<double> reals(std::max(1, h->n_reals)); std::vector<int32_t> ints(std::max(1, h->n_ints));
    if (h->n_reals) SCM_OK(cudaMemcpy2D(reals.data(), 8, c->E.reals + pick, (size_t)P * 8, 8, h->n_reals, cudaMemcpyDeviceToHost));
    if (h->n_ints) {
      std::vector<uint8_t> b(h->n_ints);
      SCM_OK(cudaMemcpy(b.data(), c->ising.spins + (size_t)pick * h->n_ints, b.size(), cudaMemcpyDeviceToHost));
      for (int q = 0; q < h->n_ints; q++) ints[q] = b[q];
    }
    int owner_local = 0;
    if (slot_of(h, 0, li, &owner_local) != 0) { h->error.clear(); continue; }   // that chain lives on another rank
    if ((rc = blangpt_set_state(h, 0, li, h->n_reals ? reals.data() : nullptr, h->n_ints ? ints.data() : nullptr))) return fail(rc);
  }
  h->scm_log_norm = log_scaling;
  h->n_launches += c->n_launches;
  fail(0);
#undef SCM_OK
  h->tables_valid = false;
  return ensure_tables(h);
}

extern "C" int blangpt_set_scm_options(blangpt_handle* h, int32_t nParticles, double threshold, double resamplingESSThreshold) {
  if (!h) return BLANGPT_EINVAL;
  if (nParticles < 2 || !(threshold > 0.0 && threshold < 1.0)) FAIL(h, BLANGPT_EINVAL, "The adaptive tempering threshold should be between 0 and 1 (exclusive)");
  h->scm_particles = nParticles; h->scm_threshold = threshold; h->scm_resampling_ess = resamplingESSThreshold;
  return BLANGPT_OK;
}
extern "C" int blangpt_scm_summary(const blangpt_handle* h, double* logNormEstimate, int32_t* nTemperatures, int32_t* nResamplings) {
  if (!h) return BLANGPT_EINVAL;
  if (logNormEstimate) *logNormEstimate = h->scm_log_norm;
  if (nTemperatures) *nTemperatures = h->scm_n_temperatures;
  if (nResamplings) *nResamplings = h->scm_n_resamplings;
  return BLANGPT_OK;
}

extern "C" int blangpt_initialize(blangpt_handle* h, int32_t init_type) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (init_type == BLANGPT_INIT_SCM) return scm_initialize(h);
  if (init_type == BLANGPT_INIT_COPIES) { h->tables_valid = false; return ensure_tables(h); }
  if (init_type != BLANGPT_INIT_FORWARD) FAIL(h, BLANGPT_EINVAL, "unknown initialization");
  int rc = launch_explore(h, MODE_FORWARD_INIT);
  if (rc) return rc;
  if (h->one_view()) {
    prime_prev_kernel<<<(unsigned)((h->M_total + 255) / 256), 256, 0, h->stream>>>(h->E, h->St);
    CUDA_OK(h, cudaGetLastError());
    h->n_launches++;
  }
  h->tables_valid = h->one_view();
  return check_device_errors(h);
}

extern "C" int blangpt_log_density(blangpt_handle* h, double* out_logdens, double* out_loglik, int32_t* out_noos, double* out_fixed) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (!h->one_view()) FAIL(h, BLANGPT_EUNSUPPORTED, "log_density with worldSize > 1 needs the peer exchange (blangpt_peer_export / blangpt_peer_connect), or gather the table yourself (blangpt_table_ptr)");
  h->tables_valid = false;
  int rc = ensure_tables(h);
  if (rc) return rc;
  const int64_t M = h->M_total; const int N = h->cfg.nChains;
  std::vector<double> tab(M * 4); std::vector<int> sop(M);
  CUDA_OK(h, cudaMemcpyAsync(tab.data(), h->E.table, M * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaMemcpyAsync(sop.data(), h->E.slot_of_pos, M * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = check_device_errors(h))) return rc;
  for (int64_t i = 0; i < M; i++) {
    const int64_t r = i / N; const int p = (int)(i % N);
    const double* t = &tab[(r * N + sop[i]) * 4];
    if (out_loglik) out_loglik[i] = t[0];
    if (out_noos) out_noos[i] = (int32_t)t[1];
    if (out_fixed) out_fixed[i] = t[2];
    if (out_logdens) out_logdens[i] = annealed_log_density(t[2], t[0], t[1], h->ladder[r * N + p], h->cfg.annealSupport);
  }
  return BLANGPT_OK;
}

// ---------------------------------------------------------------------------
// scans
// ---------------------------------------------------------------------------
extern "C" int blangpt_run_scans(blangpt_handle* h, int32_t n_scans, const blangpt_scan_outputs* out) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (!h->one_view()) FAIL(h, BLANGPT_EINVAL, "run_scans with worldSize > 1 needs the peer exchange (blangpt_peer_export / blangpt_peer_connect); without it drive the scan with blangpt_explore / blangpt_swap");
  if (h->cfg.worldSize > 1 && out && out->intSamples) FAIL(h, BLANGPT_EUNSUPPORTED, "intSamples are not exchanged between ranks");
  if (n_scans <= 0) return BLANGPT_OK;
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int rc;
  if ((rc = ensure_tables(h))) return rc;
  const int64_t M = h->M_total; const int R = h->cfg.nReplicas;
  DevOut O{}; O.n_reals = h->n_reals; O.reversible = h->cfg.reversible;
  int next_buf = 0;
  auto staging = [&](size_t bytes, void** out_p) -> int {      // the handle's next staging buffer, at least `bytes` long
    blangpt_handle::OutBuf& b = h->outbuf[next_buf++];
    bytes = std::max<size_t>(1, bytes);
    if (b.bytes < bytes) {
      if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
      cudaError_t e = cudaMalloc(&b.p, bytes);
      if (e != cudaSuccess) { b.p = nullptr; set_error(h, "cudaMalloc(outputs): %s", cudaGetErrorString(e)); return BLANGPT_ECUDA; }
      b.bytes = bytes;
    }
    cudaMemsetAsync(b.p, 0, bytes, h->stream);
    *out_p = b.p;
    return 0;
  };
  auto alloc_out = [&](auto** dptr, const void* hostptr, size_t n) -> int {
    if (!hostptr) { next_buf++; return 0; }
    using T = std::remove_pointer_t<std::remove_pointer_t<decltype(dptr)>>;
    void* p = nullptr;
    const int rc2 = staging(n * sizeof(T), &p);
    if (rc2) return rc2;
    *dptr = (T*)p; return 0;
  };
  uint8_t* d_ising_samples = nullptr;
  if (out) {
    if ((rc = alloc_out(&O.energy, out->energy, (size_t)n_scans * M))) return rc;
    if ((rc = alloc_out(&O.logdens, out->logDensity, (size_t)n_scans * M))) return rc;
    if ((rc = alloc_out(&O.noos, out->nOutOfSupport, (size_t)n_scans * M))) return rc;
    if ((rc = alloc_out(&O.indicators, out->swapIndicators, (size_t)n_scans * M))) return rc;
    if (h->n_reals) { if ((rc = alloc_out(&O.samples, out->samples, (size_t)n_scans * R * h->n_reals))) return rc; } else next_buf++;
    if (h->n_ints && out->intSamples) {
      void* p = nullptr;
      if ((rc = staging((size_t)n_scans * R * h->n_ints, &p))) return rc;
      d_ising_samples = (uint8_t*)p;
    }
  }
  CUDA_OK(h, cudaMemsetAsync(h->E.batch_counter, 0, sizeof(uint32_t) * R, h->stream));
  if (use_fused(h)) {   // many small replicas: the whole batch of scans in one launch, one warp per chain
    const int threads = 32 * h->cfg.nChains;
    if (h->family == BLANGPT_FAM_NORMAL) replica_kernel<NormalFam><<<R, threads, 0, h->stream>>>(h->E, h->St, O, h->normal, n_scans);
    else if (h->family == BLANGPT_FAM_IID) replica_kernel<IidFam><<<R, threads, 0, h->stream>>>(h->E, h->St, O, h->iid, n_scans);
    else if (h->family == BLANGPT_FAM_HIERARCHICAL) replica_kernel<HierFam><<<R, threads, 0, h->stream>>>(h->E, h->St, O, h->hier, n_scans);
    else if (h->family == BLANGPT_FAM_ROCKETS_LATENT) replica_kernel<RocketsLatentFam><<<R, threads, 0, h->stream>>>(h->E, h->St, O, h->rockets, n_scans);
    else replica_kernel<BetaBinomialFam><<<R, threads, 0, h->stream>>>(h->E, h->St, O, h->betabin, n_scans);
    CUDA_OK(h, cudaGetLastError());
    h->n_launches++;
  } else
  for (int s = 0; s < n_scans; s++) {
    if ((rc = launch_explore(h, MODE_EXPLORE))) return rc;
    if (d_ising_samples) {
      ising_record_target(h->E, h->ising, d_ising_samples, s, h->stream);
      h->n_launches++;
    }
    if ((rc = launch_swap(h, O))) return rc;
  }
  h->n_scans += n_scans; h->scans_in_round += n_scans;
  if (out) {
    auto back = [&](void* host, const void* dev, size_t bytes) { if (host && dev) cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream); };
    back(out->energy, O.energy, (size_t)n_scans * M * 8);
    back(out->logDensity, O.logdens, (size_t)n_scans * M * 8);
    back(out->nOutOfSupport, O.noos, (size_t)n_scans * M * 4);
    back(out->swapIndicators, O.indicators, (size_t)n_scans * M * 4);
    if (h->n_reals) back(out->samples, O.samples, (size_t)n_scans * R * h->n_reals * 8);
  }
  rc = check_device_errors(h);   // synchronises the stream
  if (!rc && d_ising_samples) {
    std::vector<uint8_t> b((size_t)n_scans * R * h->n_ints);
    cudaMemcpy(b.data(), d_ising_samples, b.size(), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < b.size(); i++) out->intSamples[i] = b[i];
  }
  cudaError_t e = cudaGetLastError();
  if (!rc && e != cudaSuccess) FAIL(h, BLANGPT_ECUDA, "run_scans: %s", cudaGetErrorString(e));
  return rc;
}

extern "C" int blangpt_get_round_stats(blangpt_handle* h, blangpt_round_stats* o) {
  if (!h || !o) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int64_t M = h->M_total; const int N = h->cfg.nChains, R = h->cfg.nReplicas;
  std::vector<long long> en_n(M), acc_n(M), ls_n(M); std::vector<double> en_m1(M), en_m2(M), cov_c(M), acc_m1(M), ls_sum(M);
  std::vector<int> cov_n(M), oos_seen(R); std::vector<unsigned long long> rej(R); std::vector<double> tab(M * 4);
  auto get = [&](void* dst, const void* src, size_t b) { return cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToHost, h->stream); };
  CUDA_OK(h, get(en_n.data(), h->St.en_n, M * 8)); CUDA_OK(h, get(en_m1.data(), h->St.en_m1, M * 8));
  CUDA_OK(h, get(en_m2.data(), h->St.en_m2, M * 8)); CUDA_OK(h, get(cov_c.data(), h->St.cov_c, M * 8));
  CUDA_OK(h, get(cov_n.data(), h->St.cov_n, M * 4)); CUDA_OK(h, get(acc_n.data(), h->St.acc_n, M * 8));
  CUDA_OK(h, get(acc_m1.data(), h->St.acc_m1, M * 8)); CUDA_OK(h, get(ls_sum.data(), h->St.ls_sum, M * 8));
  CUDA_OK(h, get(ls_n.data(), h->St.ls_n, M * 8)); CUDA_OK(h, get(rej.data(), h->St.rejuv, R * 8));
  CUDA_OK(h, get(tab.data(), h->E.table, M * 4 * 8));
  CUDA_OK(h, get(oos_seen.data(), h->St.oos_seen, R * 4));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  o->nScansInRound = h->scans_in_round;
  const double nan = std::nan("");
  for (int r = 0; r < R; r++) {
    double Lambda = 0.0, ss = 0.0, thermo = 0.0; bool oos = oos_seen[r] != 0;
    for (int p = 0; p < N; p++) {
      const size_t i = (size_t)r * N + p;
      const double mean = en_n[i] ? en_m1[i] : nan;
      if (o->energyMean) o->energyMean[i] = mean;
      if (o->energyVariance) o->energyVariance[i] = en_n[i] == 0 ? nan : (en_n[i] == 1 ? 0.0 : en_m2[i] / ((double)en_n[i] - 1.0));
      if (o->energyCovariance) o->energyCovariance[i] = cov_c[i] / cov_n[i];
      if (tab[i * 4 + 1] > 0 && h->cfg.annealSupport) oos = true;
      if (p < N - 1) {
        const double a = acc_n[i] ? acc_m1[i] : nan;
        if (o->swapAcceptMean) o->swapAcceptMean[i] = a;
        if (o->logSum) o->logSum[i] = ls_sum[i];
        if (o->logSumN) o->logSumN[i] = ls_n[i];
        Lambda += 1.0 - a;
        ss += ls_sum[i] - log((double)ls_n[i]);
        const double next_mean = en_n[i + 1] ? en_m1[i + 1] : nan;
        thermo += (h->ladder[i] - h->ladder[i + 1]) * (mean + next_mean) / 2.0;
      } else {
        if (o->swapAcceptMean) o->swapAcceptMean[i] = nan;
        if (o->logSum) o->logSum[i] = nan;
        if (o->logSumN) o->logSumN[i] = 0;
      }
    }
    if (o->roundTrips) o->roundTrips[r] = (int64_t)rej[r];
    if (o->Lambda) o->Lambda[r] = Lambda;
    if (o->logZSteppingStone) o->logZSteppingStone[r] = N == 1 ? nan : ss;
    if (o->logZThermodynamic) o->logZThermodynamic[r] = oos ? nan : -thermo;
  }
  return BLANGPT_OK;
}

extern "C" int blangpt_reset_round(blangpt_handle* h) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  return reset_round(h);
}

extern "C" int blangpt_adapt(blangpt_handle* h, double* cum_xy) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int64_t M = h->M_total; const int N = h->cfg.nChains, R = h->cfg.nReplicas;
  if (N < 2) return reset_round(h);
  std::vector<long long> acc_n(M); std::vector<double> acc_m1(M);
  CUDA_OK(h, cudaMemcpyAsync(acc_n.data(), h->St.acc_n, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaMemcpyAsync(acc_m1.data(), h->St.acc_m1, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < R; r++) {
    std::vector<double> betas(h->ladder.begin() + (size_t)r * N, h->ladder.begin() + (size_t)(r + 1) * N);
    std::vector<double> acc(N - 1);
    for (int i = 0; i < N - 1; i++) acc[i] = acc_n[(size_t)r * N + i] ? acc_m1[(size_t)r * N + i] : std::nan("");
    std::vector<double> cx, cy;
    std::vector<double> upd = adapt_ladder(betas, acc, &cx, &cy);
    for (int i = 0; i < N; i++) if (!(upd[i] >= 0.0 && upd[i] <= 1.0)) FAIL(h, BLANGPT_ENUMERIC, "adapt produced an invalid ladder");
    upd[0] = 1.0;
    std::copy(upd.begin(), upd.end(), h->ladder.begin() + (size_t)r * N);
    if (cum_xy) { std::copy(cx.begin(), cx.end(), cum_xy + (size_t)r * 2 * N); std::copy(cy.begin(), cy.end(), cum_xy + (size_t)r * 2 * N + N); }
  }
  CUDA_OK(h, cudaMemcpyAsync((void*)h->E.ladder, h->ladder.data(), M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return reset_round(h);
}

// The ladder (and cumulative-Lambda knots) blangpt_adapt WOULD install for the current round's statistics, without
// touching the engine: PT.adapt after the last round "for instrumentation" (PT.java:103-107) for callers that want it
// while the final round's accumulators stay readable.
extern "C" int blangpt_adapt_preview(blangpt_handle* h, double* out_ladder, double* cum_xy) {
  if (!h || !out_ladder) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int64_t M = h->M_total; const int N = h->cfg.nChains, R = h->cfg.nReplicas;
  if (N < 2) { std::copy(h->ladder.begin(), h->ladder.end(), out_ladder); return BLANGPT_OK; }
  std::vector<long long> acc_n(M); std::vector<double> acc_m1(M);
  CUDA_OK(h, cudaMemcpyAsync(acc_n.data(), h->St.acc_n, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaMemcpyAsync(acc_m1.data(), h->St.acc_m1, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < R; r++) {
    std::vector<double> betas(h->ladder.begin() + (size_t)r * N, h->ladder.begin() + (size_t)(r + 1) * N);
    std::vector<double> acc(N - 1);
    for (int i = 0; i < N - 1; i++) acc[i] = acc_n[(size_t)r * N + i] ? acc_m1[(size_t)r * N + i] : std::nan("");
    std::vector<double> cx, cy;
    std::vector<double> upd = adapt_ladder(betas, acc, &cx, &cy);
    upd[0] = 1.0;
    std::copy(upd.begin(), upd.end(), out_ladder + (size_t)r * N);
    if (cum_xy) { std::copy(cx.begin(), cx.end(), cum_xy + (size_t)r * 2 * N); std::copy(cy.begin(), cy.end(), cum_xy + (size_t)r * 2 * N + N); }
  }
  return BLANGPT_OK;
}

extern "C" int blangpt_rounds(int32_t nScans, double f, int32_t* rn, int32_t* ra) {
  if (f < 0.0 || f >= 1.0 || !rn || !ra) return -1;
  std::vector<Round> r = rounds(nScans, f);
  if (r.size() > 64) return -1;
  for (size_t i = 0; i < r.size(); i++) { rn[i] = r[i].n_scans; ra[i] = r[i].is_adapt ? 1 : 0; }
  return (int)r.size();
}

// PT.adapt's targetAccept branch (PT.java:156-163): the resized ladder for the current round's acceptance rates.
// The caller re-creates the engine with *n_out chains, all copies of the beta = 1 state (`initialize(states[0], random)`).
extern "C" int blangpt_target_accept_ladder(blangpt_handle* h, int64_t replica, double targetAccept, double* out_desc,
                                            int32_t capacity, int32_t* n_out) {
  if (!h || !out_desc || !n_out) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int N = h->cfg.nChains;
  if (N < 2 || replica < 0 || replica >= h->cfg.nReplicas) FAIL(h, BLANGPT_EINVAL, "bad replica / single chain");
  if (!(targetAccept >= 0.0 && targetAccept < 1.0)) FAIL(h, BLANGPT_EINVAL, "targetAccept must be in [0,1)");
  const int64_t M = h->M_total;
  std::vector<long long> acc_n(M); std::vector<double> acc_m1(M);
  CUDA_OK(h, cudaMemcpyAsync(acc_n.data(), h->St.acc_n, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaMemcpyAsync(acc_m1.data(), h->St.acc_m1, M * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  const size_t base = (size_t)replica * N;
  std::vector<double> accv(N - 1);
  for (int i = 0; i < N - 1; i++) accv[i] = acc_n[base + i] ? acc_m1[base + i] : std::nan("");
  std::vector<double> desc(h->ladder.begin() + base, h->ladder.begin() + base + N);
  std::vector<double> asc(desc.rbegin(), desc.rend());
  MonotoneSpline cum = cumulative_lambda_spline(asc, intensities_ascending(accv));
  std::vector<double> part = target_acceptance_partition(cum, targetAccept);
  if (part.empty()) FAIL(h, BLANGPT_EINVAL, "invalid targetAccept");
  *n_out = (int32_t)part.size();
  if ((int)part.size() > capacity) FAIL(h, BLANGPT_EINVAL, "output buffer too small for %d chains", (int)part.size());
  for (size_t i = 0; i < part.size(); i++) out_desc[i] = part[part.size() - 1 - i];
  out_desc[0] = 1.0;
  return BLANGPT_OK;
}

extern "C" int blangpt_make_ladder(int32_t kind, int32_t nChains, double param, const double* user, int32_t n_user,
                                   int32_t allowSplineGeneralization, double* out) {
  if (!out) return BLANGPT_EINVAL;
  std::vector<double> l = make_ladder(kind, nChains, param, user, n_user, allowSplineGeneralization != 0);
  if ((int)l.size() != nChains) return BLANGPT_EINVAL;
  std::copy(l.begin(), l.end(), out);
  return BLANGPT_OK;
}

// PT.reportParallelTemperingDiagnostics (PT.java:387-505) + reportLambdaFunctions (:192-210) for one replica
extern "C" int blangpt_get_diagnostics(blangpt_handle* h, int64_t replica, blangpt_diagnostics* d) {
  if (!h || !d) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  const int N = h->cfg.nChains, R = h->cfg.nReplicas; const int64_t M = h->M_total;
  if (replica < 0 || replica >= R) FAIL(h, BLANGPT_EINVAL, "replica out of range");
  if (N < 2) FAIL(h, BLANGPT_EINVAL, "PT diagnostics need at least two chains (PT.java:102)");
  std::vector<double> acc(M), emean(M), evar(M), ecov(M), Lam(R), lz(R), lzt(R); std::vector<int64_t> rt(R);
  blangpt_round_stats st{};
  st.swapAcceptMean = acc.data(); st.energyVariance = evar.data(); st.energyCovariance = ecov.data(); st.energyMean = emean.data();
  st.Lambda = Lam.data(); st.logZSteppingStone = lz.data(); st.logZThermodynamic = lzt.data(); st.roundTrips = rt.data();
  int rc = blangpt_get_round_stats(h, &st);
  if (rc) return rc;
  const size_t base = (size_t)replica * N;
  double lowest = INFINITY, sum = 0.0, ineff = 0.0;
  for (int i = 0; i < N - 1; i++) { const double a = acc[base + i]; lowest = std::min(lowest, a); sum += a; ineff += (1.0 - a) / a; }
  d->nScansInRound = st.nScansInRound;
  d->swapLowest = lowest; d->swapAverage = sum / (N - 1);
  double csum = 0.0, cmax = -INFINITY;
  for (int c = 0; c < N; c++) {   // :400-417 energy before/after correlation, clamped to [-1,1]
    double corr = ecov[base + c] / evar[base + c];
    if (corr < -1) corr = -1;
    if (corr > 1) corr = 1;
    csum += corr; if (corr > cmax || cmax != cmax) cmax = corr;
  }
  d->energyCorrelationAverage = csum / N; d->energyCorrelationHighest = cmax;
  d->Lambda = Lam[replica];
  d->inefficiency = ineff;
  d->timeToFirstRestart = 2.0 * N * (1.0 + ineff);
  d->effectiveNScans = std::max(0.0, (double)st.nScansInRound - d->timeToFirstRestart);
  const double corrected = d->effectiveNScans == 0 ? (double)st.nScansInRound : d->effectiveNScans;
  d->temperedRestarts = rt[replica];
  d->temperedRestartRate = (double)rt[replica] / corrected;
  d->asymptoticRoundTripRate = 1.0 / (2.0 + 2.0 * d->Lambda);
  d->asymptoticRoundTripCount = d->asymptoticRoundTripRate * corrected;
  d->nonAsymptoticRoundTripRate = 1.0 / (2.0 + 2.0 * ineff);
  d->nonAsymptoticRoundTripCount = d->nonAsymptoticRoundTripRate * corrected;
  d->logNormalizationSteppingStone = lz[replica];
  d->logNormalizationThermodynamic = lzt[replica];
  // cumulative Lambda(beta) and its derivative on the 99-point grid of reportLambdaFunctions
  std::vector<double> desc(h->ladder.begin() + base, h->ladder.begin() + base + N);
  std::vector<double> asc(desc.rbegin(), desc.rend());
  std::vector<double> accv(acc.begin() + base, acc.begin() + base + N - 1);
  MonotoneSpline cum = cumulative_lambda_spline(asc, intensities_ascending(accv));
  for (int i = 1; i < 100; i++) {
    const double beta = (double)i / 100.0;
    d->cumulativeLambda[i - 1] = cum(beta);
    d->lambdaInstantaneous[i - 1] = cum.derivative(beta);
  }
  return BLANGPT_OK;
}

extern "C" int blangpt_perform_inference(blangpt_handle* h, int32_t nScans, double adaptFraction, int32_t thinning,
                                         const blangpt_scan_outputs* out, blangpt_inference_summary* sum) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (adaptFraction < 0.0 || adaptFraction >= 1.0) FAIL(h, BLANGPT_EINVAL, "adaptFraction must be in [0,1)");
  (void)thinning;   // samples are recorded every scan; thinning is applied by the writer
  const int N = h->cfg.nChains, R = h->cfg.nReplicas; const int64_t M = h->M_total;
  const double f = N == 1 ? 0.0 : adaptFraction;   // PT.java:235
  std::vector<Round> rs = rounds(nScans, f);
  if (rs.size() > 64) FAIL(h, BLANGPT_EINVAL, "too many rounds");
  if (sum) { sum->nRounds = (int)rs.size(); }
  int64_t scan = 0; int rc;
  std::vector<double> acc(M), Lam(R), lz(R), lzt(R); std::vector<int64_t> rt(R);
  for (size_t r = 0; r < rs.size(); r++) {
    auto t0 = std::chrono::steady_clock::now();
    blangpt_scan_outputs o{};
    if (out) {
      if (out->energy) o.energy = out->energy + scan * M;
      if (out->logDensity) o.logDensity = out->logDensity + scan * M;
      if (out->nOutOfSupport) o.nOutOfSupport = out->nOutOfSupport + scan * M;
      if (out->swapIndicators) o.swapIndicators = out->swapIndicators + scan * M;
      if (out->samples) o.samples = out->samples + scan * R * h->n_reals;
      if (out->intSamples) o.intSamples = out->intSamples + scan * R * h->n_ints;
    }
    if ((rc = blangpt_run_scans(h, rs[r].n_scans, out ? &o : nullptr))) return rc;
    scan += rs[r].n_scans;
    if (sum) { sum->roundNScans[r] = rs[r].n_scans; sum->roundIsAdapt[r] = rs[r].is_adapt; }
    if (N > 1) {
      blangpt_round_stats st{};
      st.swapAcceptMean = acc.data(); st.Lambda = Lam.data(); st.logZSteppingStone = lz.data();
      st.logZThermodynamic = lzt.data(); st.roundTrips = rt.data();
      if ((rc = blangpt_get_round_stats(h, &st))) return rc;
      if (sum) for (int q = 0; q < R; q++) {
        if (sum->Lambda) sum->Lambda[r * R + q] = Lam[q];
        if (sum->logZ) sum->logZ[r * R + q] = lz[q];
        if (sum->roundTrips) sum->roundTrips[r * R + q] = rt[q];
        double mn = INFINITY, mean = 0.0;
        for (int i = 0; i < N - 1; i++) { const double a = acc[(size_t)q * N + i]; if (a < mn) mn = a; mean += a; }
        if (sum->minAccept) sum->minAccept[r * R + q] = mn;
        if (sum->meanAccept) sum->meanAccept[r * R + q] = mean / (N - 1);
      }
      // the reference also adapts after the last round "for instrumentation" (PT.java:103-107);
      // we keep the final round's accumulators readable instead
      if (r + 1 < rs.size() && (rc = blangpt_adapt(h, nullptr))) return rc;
    }
    if (sum) sum->roundMillis[r] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  return BLANGPT_OK;
}

extern "C" int blangpt_counters(blangpt_handle* h, uint64_t out[4]) {
  if (!h || !out) return BLANGPT_EINVAL;
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!h->model_set) return BLANGPT_OK;
  std::vector<unsigned long long> ev(h->E.M_local), ex(h->E.M_local);
  CUDA_OK(h, cudaMemcpyAsync(ev.data(), h->E.evals, ev.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaMemcpyAsync(ex.data(), h->E.execs, ex.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  for (auto v : ev) out[0] += v;
  for (auto v : ex) out[3] += v;
  out[1] = h->n_scans; out[2] = h->n_launches;
  return BLANGPT_OK;
}

// ---------------------------------------------------------------------------
// split-phase multi-GPU path
// ---------------------------------------------------------------------------
extern "C" int blangpt_set_stream(blangpt_handle* h, void* s) {
  if (!h) return BLANGPT_EINVAL;
  h->stream = (cudaStream_t)s;
  return BLANGPT_OK;
}
extern "C" int blangpt_evaluate(blangpt_handle* h) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  return launch_explore(h, MODE_EVAL);
}
extern "C" int blangpt_prime(blangpt_handle* h) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  prime_prev_kernel<<<(unsigned)((h->M_total + 255) / 256), 256, 0, h->stream>>>(h->E, h->St);
  CUDA_OK(h, cudaGetLastError());
  h->n_launches++;
  h->tables_valid = true;
  return BLANGPT_OK;
}
extern "C" int blangpt_explore(blangpt_handle* h) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (!h->tables_valid) FAIL(h, BLANGPT_ESTATE, "call blangpt_evaluate, gather the table and blangpt_prime before the first blangpt_explore");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  return launch_explore(h, MODE_EXPLORE);
}
extern "C" void* blangpt_table_ptr(blangpt_handle* h, int32_t which) {
  if (!h || !h->model_set) return nullptr;
  return which == 0 ? (void*)h->E.table : (void*)h->E.send;
}
extern "C" int64_t blangpt_local_chains(const blangpt_handle* h, int64_t* first) {
  if (!h || !h->model_set) return -1;
  if (first) *first = h->E.slot_lo;
  return h->E.M_local;
}
extern "C" int blangpt_swap(blangpt_handle* h, const blangpt_scan_outputs* out) {
  if (!h) return BLANGPT_EINVAL;
  if (!h->model_set) FAIL(h, BLANGPT_ESTATE, "set_model first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  CUDA_OK(h, cudaMemsetAsync(h->E.batch_counter, 0, sizeof(uint32_t) * h->cfg.nReplicas, h->stream));
  DevOut O{}; O.n_reals = h->n_reals; O.reversible = h->cfg.reversible;
  int rc;
  const size_t M = (size_t)h->M_total, R = (size_t)h->cfg.nReplicas;
  if (out) {   // this scan's per-chain outputs (one scan's worth of the run_scans layout); buffers are kept in the handle
    if (out->intSamples) FAIL(h, BLANGPT_EUNSUPPORTED, "blangpt_swap: intSamples are recorded by run_scans only");
    if (!h->swap_out) {
      const size_t doubles = 2 * M + R * (size_t)std::max(1, h->n_reals), ints = 2 * M;
      if ((rc = dev_alloc(h, &h->swap_out, doubles)) || (rc = dev_alloc(h, &h->swap_out_i, ints))) return rc;
    }
    if (out->energy) O.energy = h->swap_out;
    if (out->logDensity) O.logdens = h->swap_out + M;
    if (out->samples && h->n_reals) {
      O.samples = h->swap_out + 2 * M;     // filled by the rank that owns the beta = 1 chain of a replica; else NaN
      std::vector<double> nan(R * h->n_reals, std::nan(""));
      CUDA_OK(h, cudaMemcpyAsync(O.samples, nan.data(), nan.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    if (out->nOutOfSupport) O.noos = h->swap_out_i;
    if (out->swapIndicators) O.indicators = h->swap_out_i + M;
  }
  if ((rc = launch_swap(h, O))) return rc;
  h->n_scans++; h->scans_in_round++;
  if (out) {
    auto back = [&](void* host, const void* dev, size_t bytes) { if (host && dev) cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream); };
    back(out->energy, O.energy, M * 8);
    back(out->logDensity, O.logdens, M * 8);
    back(out->nOutOfSupport, O.noos, M * 4);
    back(out->swapIndicators, O.indicators, M * 4);
    back(out->samples, O.samples, R * h->n_reals * 8);
    return check_device_errors(h);   // synchronises the stream
  }
  return BLANGPT_OK;
}
extern "C" int blangpt_synchronize(blangpt_handle* h) {
  if (!h) return BLANGPT_EINVAL;
  return check_device_errors(h);
}

// ---------------------------------------------------------------------------
// multi-GPU exchange: connect the ranks' mailboxes (kernels.cuh PeerBox).  After this, run_scans / perform_inference /
// initialize / log_density work on a sharded handle exactly as on a single GPU: every exploration pass publishes the
// chains' records to all ranks over NVLink and no host code runs between the scans.
// ---------------------------------------------------------------------------
static int peer_finish(blangpt_handle* h, const std::vector<void*>& boxes) {
  PeerBox& B = h->E.peers;
  const int W = h->cfg.worldSize;
  for (int r = 0; r < W; r++) {
    B.record[r] = (MailHalf*)boxes[r];
    B.sample[r] = (MailHalf*)((char*)boxes[r] + h->mail_record_bytes);
  }
  B.my_record = B.record[h->cfg.rank]; B.my_sample = B.sample[h->cfg.rank];
  B.world = W; B.xseq = 0; B.n_reals = h->n_reals;   // 0 for Ising: no real-valued sample to exchange
  h->peers_ready = true; h->tables_valid = false;
  return BLANGPT_OK;
}
extern "C" int blangpt_peer_export(blangpt_handle* h, void* out_handle64) {
  if (!h || !out_handle64) return BLANGPT_EINVAL;
  if (!h->model_set || !h->mailbox) FAIL(h, BLANGPT_ESTATE, "peer_export: set_model on a handle with worldSize > 1 first");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  CUDA_OK(h, cudaIpcGetMemHandle(&hd, h->mailbox));
  memcpy(out_handle64, &hd, 64);
  return BLANGPT_OK;
}
extern "C" int blangpt_peer_connect(blangpt_handle* h, const void* handles, int32_t n) {
  if (!h || !handles) return BLANGPT_EINVAL;
  if (!h->model_set || !h->mailbox) FAIL(h, BLANGPT_ESTATE, "peer_connect: set_model on a handle with worldSize > 1 first");
  if (n != h->cfg.worldSize || n > kMaxPeers) FAIL(h, BLANGPT_EINVAL, "peer_connect: expected %d handles (at most %d ranks)", h->cfg.worldSize, kMaxPeers);
  if (h->peers_ready) FAIL(h, BLANGPT_ESTATE, "peers already connected");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  std::vector<void*> boxes(n, nullptr);
  for (int r = 0; r < n; r++) {
    if (r == h->cfg.rank) { boxes[r] = h->mailbox; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)handles + (size_t)r * 64, 64);
    void* p = nullptr;
    CUDA_OK(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_maps.push_back(p);
    boxes[r] = p;
  }
  return peer_finish(h, boxes);
}
// same-process variant (one host process driving several GPUs, e.g. one JVM): the mailboxes are plain device pointers
// (blangpt_peer_mailbox of each handle); peer access between the devices is enabled here
extern "C" void* blangpt_peer_mailbox(blangpt_handle* h) { return h ? h->mailbox : nullptr; }
extern "C" int blangpt_peer_connect_pointers(blangpt_handle* h, void* const* mailboxes, const int32_t* devices, int32_t n) {
  if (!h || !mailboxes || !devices) return BLANGPT_EINVAL;
  if (!h->model_set || !h->mailbox) FAIL(h, BLANGPT_ESTATE, "peer_connect: set_model on a handle with worldSize > 1 first");
  if (n != h->cfg.worldSize || n > kMaxPeers) FAIL(h, BLANGPT_EINVAL, "peer_connect: expected %d mailboxes (at most %d ranks)", h->cfg.worldSize, kMaxPeers);
  if (h->peers_ready) FAIL(h, BLANGPT_ESTATE, "peers already connected");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  std::vector<void*> boxes(n, nullptr);
  for (int r = 0; r < n; r++) {
    boxes[r] = r == h->cfg.rank ? h->mailbox : mailboxes[r];
    if (r != h->cfg.rank && devices[r] != h->cfg.device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FAIL(h, BLANGPT_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", devices[r], cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  return peer_finish(h, boxes);
}

// ---------------------------------------------------------------------------
// micro-benchmarks: the measured fp64 compute ceilings bench.py reports next to
// the HBM roofline (MEASURED_PEAKS.json has no fp64 figure)
// ---------------------------------------------------------------------------
__global__ void mb_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void mb_logit_kernel(double* out, int iters) {
  __shared__ int serr;
  double z0 = -3.0 + 9.0 * threadIdx.x / blockDim.x, z1 = 1.0 - z0, acc = 0.0; int cnt = 0;
  __shared__ uint8_t sy[2];
  logit_table_init();
  __syncthreads();
  for (int i = 0; i < iters; i++) {
    logit_pair<false>(z0, z1, sy, acc, cnt, &serr, logit_table_address());
    z0 += 1e-6; z1 -= 1e-6;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + cnt;
}
__global__ void mb_dfma_latency_kernel(double* out, int iters, int chains) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const double b = 0.999999, c = 1e-7;
  if (chains == 1) for (int i = 0; i < iters; i++) { a0 = fma(a0, b, c); }
  else if (chains == 2) for (int i = 0; i < iters; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); }
  else for (int i = 0; i < iters; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
// which: 0 = fp64 FMA TFLOP/s, 1 = logistic row terms per second (compute-only ceiling of the C2 kernel),
// 10+c = cycles per dependent DFMA with c chains per thread, one warp per SM sub-partition (latency probe)
extern "C" int blangpt_microbench(int32_t device, int32_t which, double* out_rate) {
  if (!out_rate) return BLANGPT_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return BLANGPT_ECUDA;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (which >= 10) {
    const int chains = which - 10, iters = 200000;
    double* d = nullptr;
    if (cudaMalloc(&d, 128 * sizeof(double)) != cudaSuccess) return BLANGPT_ECUDA;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mb_dfma_latency_kernel<<<1, 128>>>(d, 1000, chains);
    cudaEventRecord(e0); mb_dfma_latency_kernel<<<1, 128>>>(d, iters, chains); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    *out_rate = (double)ms * 1e-3 * khz * 1e3 / ((double)iters * chains);   // cycles per DFMA issued per warp
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return cudaGetLastError() == cudaSuccess ? BLANGPT_OK : BLANGPT_ECUDA;
  }
  const int threads = 512, blocks = sms * 2, iters = which == 0 ? 20000 : 2000;
  double* d = nullptr;
  if (cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)) != cudaSuccess) return BLANGPT_ECUDA;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    if (which == 0) mb_dfma_kernel<<<blocks, threads>>>(d, iters); else mb_logit_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (e != cudaSuccess) return BLANGPT_ECUDA;
  const double n = (double)threads * blocks * iters;
  *out_rate = which == 0 ? n * 8 * 2 / (best * 1e-3) / 1e12 : n * 2 / (best * 1e-3);
  return BLANGPT_OK;
}

// debugging / accuracy hook: the per-row Bernoulli(logistic(z)) term exactly as the C2 kernels evaluate it.
// out_term[i] = value (or -inf when the row is out of support)
__global__ void debug_logit_kernel(const double* z, const int* y, int n, double* out) {
  __shared__ int serr;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  logit_table_init();
  __syncthreads();
  if (i >= n) return;
  double s = 0.0; int c = 0;
  const uint8_t yb = (uint8_t)y[i];
  logit_term(y[i] ? z[i] : -z[i], &yb, s, c, &serr);
  out[i] = c ? -HUGE_VAL : s;
}
extern "C" int blangpt_debug_logit_terms(int32_t device, const double* z, const int32_t* y, int64_t n, double* out) {
  if (!z || !y || !out || n <= 0) return BLANGPT_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return BLANGPT_ECUDA;
  double *dz = nullptr, *dout = nullptr; int* dy = nullptr;
  if (cudaMalloc(&dz, n * 8) != cudaSuccess || cudaMalloc(&dout, n * 8) != cudaSuccess || cudaMalloc(&dy, n * 4) != cudaSuccess) return BLANGPT_ECUDA;
  cudaMemcpy(dz, z, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, y, n * 4, cudaMemcpyHostToDevice);
  debug_logit_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dz, dy, (int)n, dout);
  cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  cudaFree(dz); cudaFree(dy); cudaFree(dout);
  return e == cudaSuccess ? BLANGPT_OK : BLANGPT_ECUDA;
}
