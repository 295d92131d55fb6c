// TEST INFRASTRUCTURE ONLY — exposes the reference task helpers that count particles in sub-volumes
// (/root/reference/src/tasks/run-fluctuations/include/run-fluctuations-aux.h:128-278, compiled unmodified)
// so that the device-side counters (ljmd_subvolume_counts, ...) can be checked against them bit for bit.
#include <limits>   // the reference header uses std::numeric_limits without including <limits>
#include "include/run-fluctuations-aux.h"

extern "C" {

int ljref_subsystem_batch(void* h, double alpha_step, int type, int* out, int cap)
{
  std::vector<int> c = RunFluctuationsFunctions::GetNSubsystemBatch(*static_cast<MDSystem*>(h), alpha_step, type);
  int n = (int)c.size(); if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) out[i] = c[i];
  return (int)c.size();
}

int ljref_velocity_batch(void* h, double vcut_max, double alpha_step, int type, int* out, int cap)
{
  std::vector<int> c = RunFluctuationsFunctions::GetNsubVzBatch(*static_cast<MDSystem*>(h), vcut_max, alpha_step, type);
  int n = (int)c.size(); if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) out[i] = c[i];
  return (int)c.size();
}

int ljref_subsystem(void* h, double fraction, int type)
{
  return RunFluctuationsFunctions::GetNSubsystem(*static_cast<MDSystem*>(h), fraction, type);
}

int ljref_velocity_subsystem(void* h, double vcut, int type)
{
  return RunFluctuationsFunctions::GetNsubVz(*static_cast<MDSystem*>(h), vcut, type);
}

// The reference's statistics of one series: NumberStatistics (src/extra/sample-moments/NumberStatistics.h) and
// TimeAverage (src/tasks/auxiliary/time-average-aux.h:27-67), as the task drivers use them.
// out[12]: mean, mean error, variance, variance error, scaled variance, its error, skewness, its error,
//          kurtosis, its error, statistical inefficiency s, correlated mean error
void ljref_series_statistics(const double* x, int n, double* out)
{
  TimeAverage a;
  for (int i = 0; i < n; ++i) a.AddObservation(x[i]);
  out[0] = a.stats.GetMean();            out[1] = a.stats.GetMeanError();
  out[2] = a.stats.GetVariance();        out[3] = a.stats.GetVarianceError();
  out[4] = a.stats.GetScaledVariance();  out[5] = a.stats.GetScaledVarianceError();
  out[6] = a.stats.GetSkewness();        out[7] = a.stats.GetSkewnessError();
  out[8] = a.stats.GetKurtosis();        out[9] = a.stats.GetKurtosisError();
  out[10] = a.GetS();                    out[11] = a.GetMeanError();
}

}  // extern "C"
