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

}  // extern "C"
