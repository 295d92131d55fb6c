// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Thin extern "C" shim around the UNMODIFIED reference CPU implementation
// (/root/reference/src/library/MDSystem.cpp, compiled in place by oracle/Makefile
// into oracle/_ref/libljmd_ref.so).  It exists so that tests/ and bench.py's
// cpu_baseline / --impl reference leg can drive the reference through ctypes.
//
// Reference init is time-seeded (thirdparty/MersenneTwister/MersenneTwister.h:244-262),
// so every parity run injects a snapshot through the public members
// (MDSystem.h:72-74) and then calls CalculateForces(); CalculateParameters();
// resetAveraging();  exactly as SURVEY.md §8c prescribes.
#include "MDSystem.h"
#include <cstring>

extern "C" {

// N, T0, rho, canonical, boundaryConditions as in MDSystem.h:10-41. useCUDA=false.
void* ljref_create(int N, double T0, double rho, int canonical, int bc)
{
  MDSystem::MDSystemConfiguration small;      // default N=128: cheap ctor
  MDSystem* s = new MDSystem(small);
  MDSystem::MDSystemConfiguration cfg;
  cfg.N = N; cfg.T0 = T0; cfg.rho = rho;
  cfg.canonical = canonical != 0;
  cfg.boundaryConditions = bc;
#ifdef LJREF_ALLOW_CUDA
  // libljmd_ref_legacy.so: the reference host layer built with -DUSE_CUDA_TOOLKIT on top of the product's
  // legacy C seam — its GPU branch (MDSystem.cpp:240-251) then runs the sm_100a kernels.
  cfg.useCUDA = true;
#else
  cfg.useCUDA = false;
#endif
  s->Reinitialize(cfg);                         // one O(N^2) evaluation
  return s;
}

void ljref_destroy(void* h) { delete static_cast<MDSystem*>(h); }

// Inject a snapshot (float[4N] each) and bring all derived state up to date.
void ljref_set_state(void* h, const float* pos, const float* vel)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  const size_t n = 4 * (size_t)s->m_config.N;
  std::memcpy(s->h_Pos, pos, n * sizeof(float));
  std::memcpy(s->h_Vel, vel, n * sizeof(float));
  s->CalculateForces();
  s->CalculateParameters();
  s->resetAveraging();
  s->t = 0.;
}

void ljref_get_state(void* h, float* pos, float* vel, float* frc)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  const size_t n = 4 * (size_t)s->m_config.N;
  if (pos) std::memcpy(pos, s->h_Pos, n * sizeof(float));
  if (vel) std::memcpy(vel, s->h_Vel, n * sizeof(float));
  if (frc) std::memcpy(frc, s->h_Force, n * sizeof(float));
}

void ljref_set_canonical(void* h, int c) { static_cast<MDSystem*>(h)->m_config.canonical = c != 0; }
void ljref_set_boundary(void* h, int bc) { static_cast<MDSystem*>(h)->setBoundaryCondition(bc); }
void ljref_set_T0(void* h, double T0) { static_cast<MDSystem*>(h)->m_config.T0 = T0; }

void ljref_integrate(void* h, double dt, int nsteps)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  for (int i = 0; i < nsteps; ++i) s->Integrate(dt);
}

// Only the force evaluation (for timing the O(N^2) loop alone).
void ljref_calculate_forces(void* h) { static_cast<MDSystem*>(h)->CalculateForces(); }

// out[0..11] = U, T, K, V, P, Pshear, t, L, av_U_tot, av_T_tot, av_p_tot, av_iters
void ljref_get_scalars(void* h, double* out)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  out[0] = s->U; out[1] = s->T; out[2] = s->K; out[3] = s->V; out[4] = s->P;
  out[5] = s->Pshear; out[6] = s->t; out[7] = s->L;
  out[8] = s->av_U_tot; out[9] = s->av_T_tot; out[10] = s->av_p_tot; out[11] = s->av_iters;
}

float ljref_get_rdf(void* h, int* out256)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  for (int i = 0; i < 256; ++i) out256[i] = s->NdNdr2[i];
  return s->rdf_dr2;
}

// RDF() curve: 256 (r, g) pairs (MDSystem.cpp:633-649).
int ljref_rdf_curve(void* h, double* r, double* g)
{
  SplineFunction f = static_cast<MDSystem*>(h)->RDF();
  for (size_t i = 0; i < f.vals.size(); ++i) { r[i] = f.vals[i].first; g[i] = f.vals[i].second; }
  return (int)f.vals.size();
}

// Velocity histogram (MDSystem.cpp:651-694).
void ljref_initvelo(void* h, double vmax, double step) { static_cast<MDSystem*>(h)->initvelo(vmax, step); }
void ljref_updatevelo(void* h) { static_cast<MDSystem*>(h)->updatevelo(); }
int ljref_getvelo(void* h, double* v, double* dens, int cap)
{
  SplineFunction f = static_cast<MDSystem*>(h)->getvelo();
  int n = (int)f.vals.size(); if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) { v[i] = f.vals[i].first; dens[i] = f.vals[i].second; }
  return n;
}

void ljref_renormalize_to_energy(void* h, double ust) { static_cast<MDSystem*>(h)->RenormalizeVelocitiesToEnergy(ust); }
void ljref_renormalize_velocities(void* h, int recalc) { static_cast<MDSystem*>(h)->RenormalizeVelocities(recalc != 0); }
void ljref_correct_total_momentum(void* h) { static_cast<MDSystem*>(h)->CorrectTotalMomentum(); }
// Overwrite h_Vel only (no re-evaluation): the state the velocity helpers act on.
void ljref_poke_velocities(void* h, const float* vel)
{
  MDSystem* s = static_cast<MDSystem*>(h);
  std::memcpy(s->h_Vel, vel, 4 * (size_t)s->m_config.N * sizeof(float));
}
double ljref_kinetic_temperature(void* h) { MDSystem* s = static_cast<MDSystem*>(h); return s->KineticTemperature(s->h_Vel); }

}  // extern "C"
