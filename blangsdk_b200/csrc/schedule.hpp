// schedule.hpp -- host-side round logic of the PT engine (fp64, microseconds per round).
//
//   rounds()         PT.rounds                         R/engines/internals/factories/PT.java:239-265
//   MonotoneSpline   Spline.MonotoneCubicSpline        R/engines/internals/Spline.java:129-203
//   adapt_ladder()   PT.adapt + EngineStaticUtils      PT.java:148-171, EngineStaticUtils.java:86-168
//   equally_spaced() ladders/EquallySpaced.java:9-18
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace bpt {

struct Round { int n_scans; bool is_adapt; };

// last round = nScans - int(f*nScans) non-adaptive scans; adaptive rounds are built
// backwards with sizes rem - (int(f*rem) - 1).  n=1000, f=0.5 -> 2,4,8,16,31,63,126,251 | 500
// (1001 scans in total: the reference's own quirk, kept).
inline std::vector<Round> rounds(int n_scans, double adapt_fraction) {
  std::vector<Round> rev;
  int remaining = (int)(adapt_fraction * n_scans);
  rev.push_back({n_scans - remaining, false});
  while (remaining > 0) {
    const int next = (int)(adapt_fraction * remaining) - 1;
    rev.push_back({remaining - next, true});
    remaining = next;
  }
  std::reverse(rev.begin(), rev.end());
  return rev;
}

inline std::vector<double> equally_spaced(int n) {
  std::vector<double> out;
  if (n == 1) { out.push_back(1.0); return out; }
  for (int i = n - 1; i >= 0; i--) out.push_back(((double)i) / ((double)n - 1.0));
  return out;
}

// Fritsch-Carlson monotone cubic Hermite spline; tangents clamped when hypot(a,b) > 3
struct MonotoneSpline {
  std::vector<double> x, y, m;
  MonotoneSpline(const std::vector<double>& xs, const std::vector<double>& ys) : x(xs), y(ys), m(xs.size()) {
    const int n = (int)x.size();
    std::vector<double> d(n - 1);
    for (int i = 0; i < n - 1; i++) d[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    m[0] = d[0];
    for (int i = 1; i < n - 1; i++) m[i] = (d[i - 1] + d[i]) * 0.5;
    m[n - 1] = d[n - 2];
    for (int i = 0; i < n - 1; i++) {
      if (d[i] == 0.0) { m[i] = 0.0; m[i + 1] = 0.0; }
      else {
        const double a = m[i] / d[i], b = m[i + 1] / d[i];
        const double h = std::hypot(a, b);
        if (h > 3.0) { const double t = 3.0 / h; m[i] *= t; m[i + 1] *= t; }
      }
    }
  }
  double operator()(double at) const {
    const int n = (int)x.size();
    if (std::isnan(at)) return at;
    if (at <= x[0]) return y[0];
    if (at >= x[n - 1]) return y[n - 1];
    int i = 0;
    while (at >= x[i + 1]) { i += 1; if (at == x[i]) return y[i]; }
    const double h = x[i + 1] - x[i];
    const double t = (at - x[i]) / h;
    return (y[i] * (1 + 2 * t) + h * m[i] * t) * (1 - t) * (1 - t) +
           (y[i + 1] * (3 - 2 * t) + h * m[i + 1] * (t - 1)) * t * t;
  }
  // d/dx of the interpolant (R/engines/internals/SplineDerivatives.xtend:10-32): 0 outside the knots
  double derivative(double at) const {
    const int n = (int)x.size();
    if (std::isnan(at)) return at;
    if (at <= x[0] || at >= x[n - 1]) return 0.0;
    int i = 0;
    while (at >= x[i + 1]) i += 1;
    const double h = x[i + 1] - x[i];
    const double t = (at - x[i]) / h, dt = 1.0 / h;
    // y_i (1+2t)(1-t)^2 + h m_i t (1-t)^2 + y_{i+1} (3-2t) t^2 + h m_{i+1} (t-1) t^2
    const double a = y[i] * (2 * (1 - t) * (1 - t) - 2 * (1 + 2 * t) * (1 - t));
    const double b = h * m[i] * ((1 - t) * (1 - t) - 2 * t * (1 - t));
    const double c = y[i + 1] * (-2 * t * t + 2 * t * (3 - 2 * t));
    const double d = h * m[i + 1] * (t * t + 2 * t * (t - 1));
    return (a + b + c + d) * dt;
  }
};

// EngineStaticUtils.acceptProbabilitiesToIntensities :95-105 with PT.adapt's clamps (PT.java:152), reversed
// to ascending beta
inline std::vector<double> intensities_ascending(const std::vector<double>& accept_by_position) {
  const int m = (int)accept_by_position.size();
  std::vector<double> out(m);
  for (int i = 0; i < m; i++) {
    double a = accept_by_position[i];
    if (a == 1.0) a = 0.99999; else if (!std::isfinite(a)) a = 0.0;
    out[m - 1 - i] = 1.0 - a;
  }
  return out;
}

// paddedCumulative :86-93
inline std::vector<double> padded_cumulative(const std::vector<double>& intens, double nudge) {
  std::vector<double> ys(intens.size() + 1, 0.0);
  for (size_t i = 1; i < ys.size(); i++) ys[i] = ys[i - 1] + intens[i - 1] + nudge;
  return ys;
}

// estimateCumulativeLambdaFromIntensities :112-117: spline through (beta_k, Lambda_k)
inline MonotoneSpline cumulative_lambda_spline(const std::vector<double>& betas_asc, const std::vector<double>& intens) {
  return MonotoneSpline(betas_asc, padded_cumulative(intens, 0.0));
}

// estimateScheduleGeneratorFromIntensities :119-155 + fixedSizeOptimalPartitionFromScheduleGenerator :157-168:
// spline through (Lambda_k / Lambda_N, beta_k), evaluated at n_out equally spaced points, ends pinned to 0 and 1.
// Zero-width steps: retry once with a 1e-6 nudge (:144-151).  Returns the ladder in ASCENDING order.
inline std::vector<double> partition_from_intensities(const std::vector<double>& betas_asc, const std::vector<double>& intens,
                                                      int n_out) {
  const int n = (int)betas_asc.size();
  std::vector<double> ys;
  double nudge = 0.0;
  for (int attempt = 0; attempt < 2; attempt++) {
    ys = padded_cumulative(intens, nudge);
    const double norm = ys[n - 1];
    bool retry = false;
    for (int i = 0; i < n; i++) {
      ys[i] /= norm;
      if (i > 0 && ys[i] == ys[i - 1]) { retry = true; break; }
    }
    if (!retry) break;
    nudge = 1e-6;
  }
  MonotoneSpline generator(ys, betas_asc);
  std::vector<double> out(n_out);
  out[0] = 0.0;
  for (int i = 1; i < n_out - 1; i++) out[i] = generator(i / (n_out - 1.0));
  out[n_out - 1] = 1.0;
  return out;
}

// PT.adapt (PT.java:148-171, non-targetAccept branch): betas_desc[n] (position order) and mean swap acceptance per
// pair [n-1] (position order) -> new descending ladder; cum_x / cum_y receive the cumulative-Lambda knots.
inline std::vector<double> adapt_ladder(const std::vector<double>& betas_desc, const std::vector<double>& accept,
                                        std::vector<double>* cum_x, std::vector<double>* cum_y) {
  const int n = (int)betas_desc.size();
  std::vector<double> asc(betas_desc.rbegin(), betas_desc.rend());
  const std::vector<double> intens = intensities_ascending(accept);
  if (cum_x) *cum_x = asc;
  if (cum_y) *cum_y = padded_cumulative(intens, 0.0);
  std::vector<double> out_asc = partition_from_intensities(asc, intens, n);
  std::vector<double> desc(out_asc.rbegin(), out_asc.rend());
  std::sort(desc.begin(), desc.end(), [](double a, double b) { return a > b; });
  return desc;
}

// commons-math3 PegasusSolver (BaseSecantSolver.doSolve, ANY_SIDE), restated from its published algorithm:
// a bracketing secant method; `atol` is the solver's absolute accuracy, relative 1e-14, function value 1e-15
template <class F>
inline double pegasus_solve(F f, double x0, double x1, double atol, int max_eval) {
  const double ftol = 1e-15, rtol = 1e-14;
  double f0 = f(x0), f1 = f(x1);
  int evals = 2;
  if (f0 == 0.0) return x0;
  if (f1 == 0.0) return x1;
  for (;;) {
    const double x = x1 - ((f1 * (x1 - x0)) / (f1 - f0));
    const double fx = f(x);
    if (++evals > max_eval) return x;
    if (fx == 0.0) return x;
    if (f1 * fx < 0) { x0 = x1; f0 = f1; }
    else f0 *= f1 / (f1 + fx);
    x1 = x; f1 = fx;
    if (std::fabs(f1) <= ftol) return x1;
    if (std::fabs(x1 - x0) < std::max(rtol * std::fabs(x1), atol)) return x1;
  }
}

// EngineStaticUtils.targetAcceptancePartition :75-83 + fixedSizeOptimalPartition :176-198: resize the ladder so
// that every pair rejects with probability about 1 - targetAccept.  Returns the new ladder ASCENDING.
inline std::vector<double> target_acceptance_partition(const MonotoneSpline& cumulative_lambda, double target_accept) {
  std::vector<double> result;
  if (!(target_accept >= 0.0 && target_accept <= 1.0)) return result;
  const double Lambda = cumulative_lambda(1.0);
  const int n_grids = std::max(2, (int)(Lambda / (1.0 - target_accept)));
  double left = 0.0;
  for (int i = 0; i < n_grids - 1; i++) {
    const double y = Lambda * i / (n_grids - 1.0);
    const double lo = std::max(0.0, left - 0.1);     // relaxed bracket, :188-190
    const double point = pegasus_solve([&](double x) { return cumulative_lambda(x) - y; }, lo, 1.0, 1e-16, 10000);
    result.push_back(point);
    left = point;
  }
  result.push_back(1.0);
  std::sort(result.begin(), result.end());
  return result;
}

// ---- initial ladders (R/engines/internals/ladders/*.java) --------------------------------------------------
enum LadderKind { LADDER_EQUALLY_SPACED = 0, LADDER_GEOMETRIC = 1, LADDER_POLYNOMIAL = 2, LADDER_USER_SPECIFIED = 3 };

// returns an empty vector on the conditions under which the reference throws
inline std::vector<double> make_ladder(int kind, int n_chains, double param, const double* user, int n_user,
                                       bool allow_spline_generalization) {
  std::vector<double> out;
  if (n_chains < 1) return out;
  switch (kind) {
    case LADDER_EQUALLY_SPACED: return equally_spaced(n_chains);
    case LADDER_GEOMETRIC:                                   // Geometric.java:17-33, annealingScaling default 0.8
      if (param < 0.0 || param >= 1.0) return out;
      if (n_chains == 1) { out.push_back(1.0); return out; }
      for (int i = 0; i < n_chains - 1; i++) out.push_back(std::pow(param, i));
      out.push_back(0.0);
      return out;
    case LADDER_POLYNOMIAL:                                  // Polynomial.java:17-31, power default 3
      if (param < 1.0) return out;
      if (n_chains == 1) { out.push_back(1.0); return out; }
      for (int i = 0; i < n_chains; i++) out.push_back(std::pow(1.0 - (double)i / ((double)n_chains - 1.0), param));
      return out;
    case LADDER_USER_SPECIFIED: {                            // UserSpecified.java:24-83
      if (!user || n_user < 1) return out;
      std::vector<double> sorted(user, user + n_user);
      std::sort(sorted.begin(), sorted.end(), [](double a, double b) { return a > b; });
      if (n_user == n_chains) { if (sorted[0] != 1.0) return std::vector<double>(); return sorted; }
      if (!allow_spline_generalization) return out;
      if (n_chains == 1) { out.push_back(1.0); return out; }
      if (sorted[0] != 1.0 || sorted[n_user - 1] != 0.0 || n_user < 2) return out;
      std::vector<double> asc(sorted.rbegin(), sorted.rend());
      std::vector<double> uniform(n_user - 1, 0.5);
      std::vector<double> res = partition_from_intensities(asc, uniform, n_chains);
      std::sort(res.begin(), res.end(), [](double a, double b) { return a > b; });
      return res;
    }
  }
  return out;
}

}  // namespace bpt
