// pool_host.hpp -- host interface of the pool exploration kernel (design: pool.cuh; compiled in pool.cu).
// Deliberately free of device code so that engine.cu does not recompile when the kernel changes.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace bpt {

struct PoolState;   // opaque: global-memory request / result blocks of one handle

// Allocates the pool for `chains` local chains on `device`.  Returns 0 and *out != nullptr on success; 0 and
// *out == nullptr when the pool cannot serve this shape (no cooperative launch, too many chains per SM);
// a cudaError_t (> 0) when an allocation fails.
int pool_create(PoolState** out, int chains, int device, int family /* BLANGPT_FAM_LOGISTIC | BLANGPT_FAM_MIXTURE */, int mixture_K);
void pool_destroy(PoolState* ps);
// L2 residency control for working sets beyond the L2: launches carry an access-policy window over [base, base + bytes)
// (the per-chain caches) in which a fraction `hit_ratio` of the lines persists and the rest streams
void pool_set_l2_window(PoolState* ps, void* base, size_t bytes, float hit_ratio);
// one launch = one exploration pass (mode: ExploreMode) of every local chain; returns a cudaError_t
int pool_launch_logistic(const PoolState* ps, const DevEngine& E, const LogisticFam::Data& D, int mode, cudaStream_t stream);
int pool_launch_mixture(const PoolState* ps, const DevEngine& E, const MixtureFam<4>::Data& D, int mode, cudaStream_t stream);

}  // namespace bpt
