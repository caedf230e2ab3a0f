// slice.cuh -- Neal (2003) slice samplers exactly as blang.mcmc runs them.
//
// Every thread of a chain group executes this control flow redundantly; the
// only collective part is `lp(x)`, which evaluates the sampler's log-density
// (sum over the variable's connected factors) with a group reduction.  All
// guards of the reference are kept:
//   * log slice height  = lp(x0) - Exp(1), or -1e100 when lp(x0) = -inf
//   * stepping out by doubling, w = 1, at most 20 rounds, short-circuit ||
//   * shrinkage with right-exclusive uniform, non-strict acceptance
//   * Neal's doubling-consistency test with the 1.1*w threshold
//   * the 1e-6 degenerate-interval exit restoring x0
// Reference: R/mcmc/RealSliceSampler.java:53-148, R/mcmc/IntSliceSampler.java:53-132.
#pragma once
#include "common.cuh"

namespace bpt {

// RealSliceSampler.nextLogSliceHeight :143-148
BPT_D double next_log_slice_height(Rng& rng, double log_density) {
  if (log_density == BPT_NEG_INF) return -1e100;
  return log_density - gen_unit_exponential(rng);
}

// RealSliceSampler.accept :121-141
template <class LP>
BPT_D bool real_accept(LP& lp, bool fixed_window, double old_state, double new_state, double h, double left,
                       double right) {
  if (fixed_window) return true;
  bool differ = false;
  while (right - left > 1.1 * 1.0) {
    const double middle = (left + right) / 2.0;
    if ((old_state < middle && new_state >= middle) || (old_state >= middle && new_state < middle)) differ = true;
    if (new_state < middle) right = middle; else left = middle;
    if (differ && h >= lp(left) && h >= lp(right)) return false;
  }
  return true;
}

// RealSliceSampler.execute :53-119.  Returns the new value of the coordinate.
template <class LP>
BPT_D double real_slice(LP& lp, Rng& rng, double old_state, bool fixed_window, double fw_left, double fw_right,
                        int& err) {
  const double h = next_log_slice_height(rng, lp(old_state));
  double left, right;
  if (fixed_window) { left = fw_left; right = fw_right; }
  else {
    left = old_state - 1.0 * rng.next_double();
    right = left + 1.0;
    int iter = 0;
    while (iter++ < 20 && (h < lp(left) || h < lp(right))) {
      if (rng.bernoulli(0.5)) {
        left += -(right - left);
        if (left == BPT_NEG_INF) { err |= ERR_SLICE_DIVERGED; return old_state; }
      } else {
        right += right - left;
        if (right == BPT_POS_INF) { err |= ERR_SLICE_DIVERGED; return old_state; }
      }
    }
  }
  double sl = left, sr = right;
  for (;;) {
    const double new_state = gen_uniform(rng, sl, sr);
    if (h <= lp(new_state) && real_accept(lp, fixed_window, old_state, new_state, h, left, right)) return new_state;
    if (new_state < old_state) sl = new_state; else sr = new_state;
    if (fabs(sl - sr) < 1e-6) return old_state;   // NumericalUtils.isClose(.,.,1e-6) :109-117
  }
}

// IntSliceSampler.accept :115-132
template <class LP>
BPT_D bool int_accept(LP& lp, int old_state, int new_state, double h, int left, int right) {
  bool differ = false;
  while (right - left > 10) {
    const int middle = (left + right) / 2;
    if ((old_state < middle && new_state >= middle) || (old_state >= middle && new_state < middle)) differ = true;
    if (new_state < middle) right = middle; else left = middle;
    if (differ && h >= lp(left) && h >= lp(right - 1)) return false;
  }
  return true;
}

// IntSliceSampler.execute :53-113
// When no integer of the window is acceptable -- not even the current state, i.e. its density is -inf and the
// stepping out did not reach the support -- the reference's shrinkage ends in Generators.discreteUniform throwing
// "Invalid arguments" (Generators.java:205-211): reported here as ERR_INT_SLICE_EMPTY, state unchanged.
template <class LP>
BPT_D int int_slice(LP& lp, Rng& rng, int old_state, bool fixed_window, int fw_left, int fw_right, int& err) {
  const double h = next_log_slice_height(rng, lp(old_state));
  int left, right;
  if (fixed_window) { left = fw_left; right = fw_right; }
  else {
    left = old_state - rng.next_int(10);
    right = left + 10;
    int iter = 0;
    while (iter++ < 20 && (h < lp(left) || h < lp(right - 1))) {
      if (rng.bernoulli(0.5)) left += -(right - left);
      else right += right - left;
    }
  }
  int sl = left, sr = right;
  for (;;) {
    if (sr <= sl) { err |= ERR_INT_SLICE_EMPTY; return old_state; }
    const int new_state = rng.next_int(sr - sl) + sl;   // Generators.discreteUniform :205-222
    if (h <= lp(new_state) && int_accept(lp, old_state, new_state, h, left, right)) return new_state;
    if (new_state < old_state) sl = new_state + 1; else sr = new_state;
  }
}

}  // namespace bpt
