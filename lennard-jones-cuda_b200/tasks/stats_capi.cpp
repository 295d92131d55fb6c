// C entry points over sample_statistics.h so the parity tests can feed a series and compare every estimator
// with the reference's NumberStatistics / TimeAverage (oracle/ref_stats_shim.cpp wraps those the same way).
#include <cstring>
#include <string>
#include <vector>

#include "fluctuation_observers.h"
#include "sample_statistics.h"
#include "task_parameters.h"

extern "C" {
// out[12]: mean, mean error, variance, variance error, scaled variance, its error, skewness, its error,
//          kurtosis, its error, statistical inefficiency s, correlated mean error
void ljtasks_series_statistics(const double* x, int n, double* out) {
  ljtasks::CorrelatedAverage a;
  for (int i = 0; i < n; ++i) a.add(x[i]);
  out[0] = a.stats.mean();            out[1] = a.stats.mean_error();
  out[2] = a.stats.variance();        out[3] = a.stats.variance_error();
  out[4] = a.stats.scaled_variance(); out[5] = a.stats.scaled_variance_error();
  out[6] = a.stats.skewness();        out[7] = a.stats.skewness_error();
  out[8] = a.stats.kurtosis();        out[9] = a.stats.kurtosis_error();
  out[10] = a.inefficiency();         out[11] = a.mean_error();
}

// kind 0: run-fluctuations defaults, 1: run-isotherm defaults; path may be "" (defaults only).
// Returns the value of `key`; *found = 0 when the key is unknown after reading.
double ljtasks_parameter(const char* path, int kind, const char* key, int* found) {
  ljtasks::TaskParameters p = kind == 0 ? ljtasks::TaskParameters::fluctuations() : ljtasks::TaskParameters::isotherm();
  if (path && path[0]) p.read(path);
  if (found) *found = p.values.count(key) ? 1 : 0;
  return p[key];
}
// The stamped output prefix without its time stamp: "<prefix>.<stamp>.N..." -> returns the part after the stamp.
int ljtasks_prefix_tail(const char* path, int kind, char* out, int cap) {
  ljtasks::TaskParameters p = kind == 0 ? ljtasks::TaskParameters::fluctuations() : ljtasks::TaskParameters::isotherm();
  if (path && path[0]) p.read(path);
  const std::string full = p.stamped_prefix(kind == 0);
  const std::string tail = full.substr(p.output_prefix.size() + 1 + 20);   // ".dd-mm-YYYY-THH-MM-SS" is 21 chars
  std::strncpy(out, (p.output_prefix + "|" + tail).c_str(), cap - 1);
  out[cap - 1] = 0;
  return (int)tail.size();
}
int ljtasks_coordinate_fractions(double step, double* out, int cap) {
  const std::vector<double> g = ljtasks::coordinate_fractions(step);
  for (std::size_t i = 0; i < g.size() && (int)i < cap; ++i) out[i] = g[i];
  return (int)g.size();
}
int ljtasks_momentum_cuts(double T0, double step, double vfactor, double* out, int cap) {
  const std::vector<double> g = ljtasks::momentum_cuts(T0, step, vfactor);
  for (std::size_t i = 0; i < g.size() && (int)i < cap; ++i) out[i] = g[i];
  return (int)g.size();
}
}
