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
};

// betas_desc[n] (position order) and mean swap acceptance per pair [n-1] (position order)
// -> new descending ladder; cum_x / cum_y receive the cumulative-Lambda knots (ascending beta).
inline std::vector<double> adapt_ladder(const std::vector<double>& betas_desc, const std::vector<double>& accept,
                                        std::vector<double>* cum_x, std::vector<double>* cum_y) {
  const int n = (int)betas_desc.size();
  std::vector<double> asc(betas_desc.rbegin(), betas_desc.rend());
  std::vector<double> intens(n - 1);
  for (int i = 0; i < n - 1; i++) {
    double a = accept[i];
    if (a == 1.0) a = 0.99999; else if (!std::isfinite(a)) a = 0.0;   // PT.java:152
    intens[n - 2 - i] = 1.0 - a;                                       // reversed to ascending beta
  }
  auto cumulative = [&](double nudge) {
    std::vector<double> ys(n, 0.0);
    for (int i = 1; i < n; i++) ys[i] = ys[i - 1] + intens[i - 1] + nudge;
    return ys;
  };
  if (cum_x) *cum_x = asc;
  if (cum_y) *cum_y = cumulative(0.0);
  std::vector<double> ys;
  double nudge = 0.0;
  for (int attempt = 0; attempt < 2; attempt++) {   // zero-width steps: retry once with a 1e-6 nudge (:144-151)
    ys = cumulative(nudge);
    const double norm = ys[n - 1];
    bool retry = false;
    for (int i = 0; i < n; i++) {
      ys[i] /= norm;
      if (i > 0 && ys[i] == ys[i - 1]) { retry = true; break; }
    }
    if (!retry) break;
    nudge = 1e-6;
  }
  MonotoneSpline generator(ys, asc);   // schedule generator: normalised cumulative Lambda -> beta
  std::vector<double> out_asc(n);
  out_asc[0] = 0.0;
  for (int i = 1; i < n - 1; i++) out_asc[i] = generator(i / (n - 1.0));
  out_asc[n - 1] = 1.0;
  std::vector<double> desc(out_asc.rbegin(), out_asc.rend());
  std::sort(desc.begin(), desc.end(), [](double a, double b) { return a > b; });
  return desc;
}

}  // namespace bpt
