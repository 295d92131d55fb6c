// Compile-only probe of the MDSystem surface the reference's callers use (src/gui/mainwindow.cpp,
// glwidget.cpp, MDSystemGL.cpp; src/tasks/*).  tests/test_host.py compiles it with -fsyntax-only against the
// product's host/MDSystem.h and, where /root/reference exists, against the reference's own header: the same
// source must be valid for both.
#include "MDSystem.h"

double probe(bool gui) {
  MDSystem::MDSystemConfiguration config;
  config.N = 400; config.T0 = 1.4; config.rho = 0.05; config.canonical = true;
  config.boundaryConditions = 0; config.useCUDA = true; config.CUDABlockSize = 256;
  MDSystem syst(config);
  syst.Reinitialize(config);
  syst.RenormalizeVelocitiesToEnergy(1.708);
  syst.RenormalizeVelocities();
  syst.m_config.canonical = false;
  syst.setCanonical(true);
  syst.setBoundaryCondition(1);
  syst.setPeriodicBoundaryCondition(true);
  syst.setHardwareMode(true);
  syst.Integrate(0.004);
  syst.resetAveraging();
  syst.initvelo(12., 0.12);
  syst.updatevelo();
  SplineFunction velo = syst.getvelo();
  SplineFunction rdf = syst.RDF(5.0, 0.03);
  double acc = syst.Maxwell(1.0) + syst.getTime() + velo.f(1.0) + rdf.vals[0].second;
  acc += syst.U + syst.T + syst.K + syst.V + syst.P + syst.L + syst.t;
  acc += syst.av_U_tot + syst.av_T_tot + syst.av_p_tot + syst.av_iters;
  acc += syst.h_Pos[0] + syst.h_Vel[1] + syst.h_Force[2] + syst.m_config.N + syst.m_config.rho + syst.m_config.T0;
  acc += syst.NdNdr2[0] + syst.rdf_dr2 + syst.KineticTemperature(syst.h_Vel);
  if (gui) {
    float* p = syst.getArray(0);
    syst.setArray(0, p);
    syst.CalculateForces();
    syst.CalculateParameters();
    syst.ApplyBoundaryConditions();
    syst.CorrectTotalMomentum();
  }
  return acc;
}
