// Source-compatible stand-in for the reference's `MDSystem` (/root/reference/src/library/MDSystem.h:6-164),
// written on the C ABI of include/ljmd.h.  The reference's callers — the Qt GUI (src/gui) and the three
// tasks (src/tasks) — read and write these PUBLIC members directly and call these methods; compiling them
// against this header and linking libljmd_host + libljmd.so runs their MD step on the B200.
//
// Differences a maintainer should know (INTEGRATION.md has the full list):
//  * there is no CPU backend: `useCUDA` is accepted and ignored, construction fails loudly without a GPU;
//  * `NdNdr2` is refreshed when RDF() is called (the histogram is rebuilt lazily on the device from the
//    positions of the last force evaluation) instead of on every Integrate();
//  * `Pshear` is 0 unless LJMD_PSHEAR=1 is set (then every Integrate() also evaluates ljmd_get_pshear: the
//    reference's CPU-path formula, MDSystem.cpp:299,309,335,353, in one extra all-pairs pass; the reference's own
//    GPU path never sets it, MDSystem.cpp:240-251);
//  * initial conditions are sampled on the device (ljmd_init_state: the reference's lattice, seeded Philox
//    velocities; env LJMD_SEED, default time-seeded like the reference's MTRand; LJMD_HOST_INIT=1 keeps a host
//    sampler on std::mt19937_64), so `rangen` / `m_MaxwellGenerator` members are not exposed;
//  * `m_config.numGPUs` (new, defaulted) / env LJMD_DEVICES shard the system over several GPUs.
#ifndef mdsystem_h
#define mdsystem_h
// Callers of the reference header also get these through its includes and rely on them (printf, pow, ...).
#include <climits>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <vector>

#include "splinefunction.h"

struct ljmd_system;

class MDSystem {
 public:
  struct MDSystemConfiguration {
    int N;                   // number of particles
    double T0;               // initial / thermostat temperature
    double rho;              // number density
    bool canonical;          // TVN instead of EVN
    int boundaryConditions;  // 0 periodic, 1 hard wall, 2 none (expansion)
    bool useCUDA;            // accepted for compatibility; the GPU is always used
    int CUDABlockSize;       // accepted for compatibility; launch shapes are chosen by the library
    // New, defaulted, at the end so that existing aggregate-style code keeps compiling: GPUs of this
    // node the system is sharded over.  1: one device (LJMD_DEVICE, default 0).  > 1: devices 0..numGPUs-1.
    // 0: take the list from the environment variable LJMD_DEVICES ("0,1,2,3"), one device when it is unset.
    int numGPUs;
    MDSystemConfiguration()
        : N(128), T0(1.5), rho(0.2), canonical(false), boundaryConditions(0), useCUDA(false), CUDABlockSize(256),
          numGPUs(0) {}
  };

  MDSystemConfiguration m_config;

  // total energy, instant temperature, kinetic energy, potential energy, pressure
  double U, T, K, V, P;
  double Pshear;  // 0 unless LJMD_PSHEAR=1 (see above)
  bool CUDAInit;

  // host mirrors, float[4N] each, refreshed by every method that changes them
  float* h_Pos;
  float* h_Vel;
  float* h_Force;
  // kept for source compatibility; device memory is owned by the library handle
  float* d_Pos;
  float* d_Vel;
  float* d_Force;

  double t;  // current time
  double L;  // box edge

  SplineFunction curvelo;  // running mean of the speed distribution
  int veloIters;

  std::vector<int> NdNdr2;  // r^2 histogram of the last force evaluation (valid after RDF())
  float rdf_dr2;

  double momN, momN2, momN3, momN4;
  int momiters;

  double av_U_tot, av_T_tot, av_p_tot;
  int av_iters;

 public:
  MDSystem(const MDSystem::MDSystemConfiguration& config = MDSystem::MDSystemConfiguration());
  ~MDSystem(void);

  void Reinitialize(const MDSystem::MDSystemConfiguration& config);
  void ReallocateMemory();
  void SampleInitialConditions();
  void CorrectTotalMomentum();
  void CalculateForces();
  void CalculateForces(float* frc);
  double KineticTemperature(float* vel);
  double CalculateXi(float* frc, float* vel);
  void CalculateParameters();
  void RenormalizeVelocities(bool RecalculateTkin = false);
  void RenormalizeVelocitiesToEnergy(double ust);
  void ApplyBoundaryConditions();
  double getTime() { return t; }
  void Integrate(double dt);

  // Extension: n steps without refreshing the host mirrors in between (they are refreshed once at the end).
  void IntegrateMany(double dt, int nsteps);
  // Extension: the library handle behind this instance, for callers that go on with the C ABI of include/ljmd.h
  // (device-resident batches, observation trace).  The host mirrors are NOT refreshed by calls made on it.
  ljmd_system* handle() const { return m_sys; }

  float* getArray(int type);
  void setArray(int type, const float* data);

  void initvelo(double vmax = 12., double shag = 0.12);
  void updatevelo();
  SplineFunction getvelo();
  double Maxwell(double v);

  void resetAveraging();
  SplineFunction RDF(double rmax = 5., double shag = 0.1);
  void Fluctuations(double fraction);
  int fast_round(float x);

  void setHardwareMode(bool useCUDA);
  void setCanonical(bool canonical) { m_config.canonical = canonical; }
  void setPeriodicBoundaryCondition(bool periodic) { m_config.boundaryConditions = !periodic; }
  void setBoundaryCondition(int cond) { m_config.boundaryConditions = cond; }

 private:
  MDSystem(const MDSystem&);
  MDSystem& operator=(const MDSystem&);
  ljmd_system* m_sys;
  int m_velo_bins;
  double m_velo_step;
  bool m_host_vel_dirty;
  bool m_device_sampled;   // SampleInitialConditions ran on the device (state already evaluated there)  // h_Vel was edited on the host since the last upload
  void syncConfig();
  void pullScalars(bool accumulate);
  void check(int rc, const char* what);
};

#endif
