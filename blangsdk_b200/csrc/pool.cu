// pool.cu -- host side of the pool exploration kernel (see pool.cuh for the design).
#include "pool_host.hpp"
#include "pool.cuh"
#include "pool_mixture.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstring>

namespace bpt {

struct PoolState {
  PoolCtl ctl{};
  void* control = nullptr;      // requests | result slots | finished | abort: zeroed before every launch
  size_t control_bytes = 0;
  void* results = nullptr;      // wbuf
  unsigned long long* prof = nullptr; int launches = 0;
  void* win_base = nullptr; size_t win_bytes = 0; float win_ratio = 0.f;   // L2 access-policy window (pool_set_l2_window)
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <class K>
static cudaError_t prepare_kernel(K kernel, int* per_sm) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolDynSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kernel, kPoolThreads, kPoolDynSmemBytes);
}

int pool_create(PoolState** out, int chains, int device, int family, int mixture_K) {
  *out = nullptr;
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return (int)e;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  int per_sm = 0;
  if (family == 3) e = mixture_K <= 4 ? prepare_kernel(pool_kernel<MixturePool<4>>, &per_sm) : prepare_kernel(pool_kernel<MixturePool<8>>, &per_sm);
  else e = prepare_kernel(pool_kernel<LogisticPool>, &per_sm);
  if (e != cudaSuccess) return (int)e;
  const int n_ctrl = (chains + kPoolWarps - 1) / kPoolWarps;   // controller CTAs: one chain per warp
  const int n_work = sms - n_ctrl;                             // worker CTAs: one slab of rows each
  if (!coop || per_sm < 1 || n_work < sms / 2 || chains > kPoolMaxChains || n_work > 32 * kPoolMaxSlotsPerLane) return 0;
  PoolState& ps = *new PoolState();
  const size_t C = (size_t)chains;
  // control block (zeroed before every launch): requests, result slots, finished counter, abort flag
  size_t off = 0;
  const size_t o_reqa = off; off += align_up(C * sizeof(PoolHalf), 256);
  const size_t o_reqb = off; off += align_up(C * sizeof(PoolHalf), 256);
  const size_t o_slot = off; off += C * n_work * sizeof(PoolHalf);
  const size_t o_fin = off; off += sizeof(uint32_t);
  const size_t o_abort = off; off += sizeof(int);
  ps.control_bytes = align_up(off, 256);
  if ((e = cudaMalloc(&ps.control, ps.control_bytes)) != cudaSuccess) { delete &ps; return (int)e; }
  const size_t ro = align_up(C * kMaxParams * sizeof(double), 256);
  if ((e = cudaMalloc(&ps.results, ro)) != cudaSuccess) { cudaFree(ps.control); delete &ps; return (int)e; }
  cudaMemset(ps.results, 0, ro);
  char* cb = (char*)ps.control;
  PoolCtl& c = ps.ctl;
  c.C = chains; c.n_cta = n_work; c.n_ctrl = n_ctrl;
  c.reqa = (PoolHalf*)(cb + o_reqa); c.reqb = (PoolHalf*)(cb + o_reqb); c.slot = (PoolHalf*)(cb + o_slot);
  c.finished = (uint32_t*)(cb + o_fin); c.abort_flag = (int*)(cb + o_abort);
  c.wbuf = (double*)ps.results;
  c.prof = nullptr;
  if (getenv("BLANGPT_POOL_PROFILE")) {
    cudaMalloc(&ps.prof, (size_t)sms * kPoolWarps * 8 * sizeof(unsigned long long));
    cudaMemset(ps.prof, 0, (size_t)sms * kPoolWarps * 8 * sizeof(unsigned long long));
    c.prof = ps.prof;
  }
  *out = &ps;
  return 0;
}

// BLANGPT_POOL_PROFILE: where the warps' time went, summed over this handle's launches (stderr)
static void pool_report(PoolState* ps) {
  const int W = (ps->ctl.n_cta + ps->ctl.n_ctrl) * kPoolWarps;
  std::vector<unsigned long long> h((size_t)W * 8);
  cudaMemcpy(h.data(), ps->prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double served = 0, cyc_serve = 0, cyc_total = 0, polls = 0, evals = 0, cyc_wait = 0, cyc_ctrl = 0; int n_ctrl = 0;
  for (int w = 0; w < W; w++) {
    const unsigned long long* q = &h[(size_t)w * 8];
    served += q[0]; cyc_serve += q[1]; cyc_total += q[2]; polls += q[3];
    if (q[4]) { evals += q[4]; cyc_wait += q[5]; cyc_ctrl += q[6]; n_ctrl++; }
  }
  fprintf(stderr, "[pool profile] launches %d  requests served %.0f  cycles/request %.0f  worker busy %.3f  polls/request %.2f\n",
          ps->launches, served, cyc_serve / std::max(1.0, served), cyc_serve / std::max(1.0, cyc_total), polls / std::max(1.0, served));
  fprintf(stderr, "[pool profile] controllers %d  evaluations %.0f  cycles/evaluation %.0f  of which waiting %.0f (%.3f)\n",
          n_ctrl, evals, cyc_ctrl / std::max(1.0, evals), cyc_wait / std::max(1.0, evals), cyc_wait / std::max(1.0, cyc_ctrl));
}

void pool_set_l2_window(PoolState* ps, void* base, size_t bytes, float hit_ratio) {
  if (!ps) return;
  ps->win_base = base; ps->win_bytes = bytes; ps->win_ratio = hit_ratio;
}

void pool_destroy(PoolState* ps) {
  if (!ps) return;
  if (ps->prof) { pool_report(ps); cudaFree(ps->prof); }
  if (ps->control) cudaFree(ps->control);
  if (ps->results) cudaFree(ps->results);
  delete ps;
}

template <class PF>
static int pool_launch_t(const PoolState* psp, const DevEngine& E, const typename PF::Data& D, int mode, cudaStream_t stream);

int pool_launch_logistic(const PoolState* psp, const DevEngine& E, const LogisticFam::Data& D, int mode, cudaStream_t stream) {
  return pool_launch_t<LogisticPool>(psp, E, D, mode, stream);
}
int pool_launch_mixture(const PoolState* psp, const DevEngine& E, const MixtureFam<4>::Data& D, int mode, cudaStream_t stream) {
  if (D.K <= 4) return pool_launch_t<MixturePool<4>>(psp, E, D, mode, stream);
  MixtureFam<8>::Data d8; memcpy(&d8, &D, sizeof d8);
  return pool_launch_t<MixturePool<8>>(psp, E, d8, mode, stream);
}

template <class PF>
static int pool_launch_t(const PoolState* psp, const DevEngine& E, const typename PF::Data& D, int mode, cudaStream_t stream) {
  const PoolState& ps = *psp;
  const_cast<PoolState*>(psp)->launches++;
  cudaError_t e = cudaMemsetAsync(ps.control, 0, ps.control_bytes, stream);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)(ps.ctl.n_cta + ps.ctl.n_ctrl));
  lc.blockDim = dim3((unsigned)kPoolThreads);
  lc.dynamicSmemBytes = kPoolDynSmemBytes;
  lc.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  lc.attrs = at; lc.numAttrs = 1;
  if (ps.win_base && ps.win_bytes && ps.win_ratio > 0.f) {
    at[1].id = cudaLaunchAttributeAccessPolicyWindow;
    at[1].val.accessPolicyWindow.base_ptr = ps.win_base;
    at[1].val.accessPolicyWindow.num_bytes = ps.win_bytes;
    at[1].val.accessPolicyWindow.hitRatio = ps.win_ratio;
    at[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    lc.numAttrs = 2;
  }
  e = cudaLaunchKernelEx(&lc, pool_kernel<PF>, E, D, ps.ctl, mode);
  return (int)e;
}

}  // namespace bpt
