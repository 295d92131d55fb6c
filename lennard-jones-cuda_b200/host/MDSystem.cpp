// Host-side mirror of the reference's MDSystem API on top of the C ABI (include/ljmd.h).
// Each method cites the reference lines whose behaviour it reproduces
// (paths relative to /root/reference/src/library/).
#include "MDSystem.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <random>

#include "../../include/ljmd.h"

namespace {
const double kPi = 3.141592653589793238462643;

unsigned long long initial_seed() {
  const char* e = std::getenv("LJMD_SEED");
  if (e && *e) return std::strtoull(e, 0, 10);
  return (unsigned long long)std::time(0) * 2654435761ull + (unsigned long long)std::clock();
}
}  // namespace

void SplineFunction::loadFromFile(const char* file) {
  std::ifstream fin(file);
  vals.clear();
  double x, y;
  while (fin >> x >> y) vals.push_back(std::make_pair(x, y));
  std::sort(vals.begin(), vals.end());
}

void MDSystem::check(int rc, const char* what) {
  if (rc != LJMD_OK) {
    // the reference's checkCudaErrors prints and exits (helper_cuda.h); keep that contract for callers
    std::fprintf(stderr, "MDSystem: %s failed: %s\n", what, ljmd_last_error());
    std::exit(EXIT_FAILURE);
  }
}

MDSystem::MDSystem(const MDSystem::MDSystemConfiguration& config)   // MDSystem.cpp:55-64
    : U(0), T(0), K(0), V(0), P(0), Pshear(0), CUDAInit(false), h_Pos(0), h_Vel(0), h_Force(0), d_Pos(0), d_Vel(0),
      d_Force(0), t(0), L(0), veloIters(0), rdf_dr2(0.1f), momN(0), momN2(0), momN3(0), momN4(0), momiters(0),
      av_U_tot(0), av_T_tot(0), av_p_tot(0), av_iters(0), m_sys(0), m_velo_bins(101), m_velo_step(0.12),
      m_host_vel_dirty(false), m_device_sampled(false) {
  Reinitialize(config);
}

MDSystem::~MDSystem(void) {   // MDSystem.cpp:218-230
  delete[] h_Pos;
  delete[] h_Vel;
  delete[] h_Force;
  if (m_sys) ljmd_destroy(m_sys);
}

void MDSystem::Reinitialize(const MDSystem::MDSystemConfiguration& config) {   // MDSystem.cpp:66-114
  m_config = config;
  const int N = m_config.N;
  L = std::pow(N / m_config.rho, 1. / 3.);   // :70
  T = m_config.T0;                           // :72
  t = 0;
  delete[] h_Pos;
  delete[] h_Vel;
  delete[] h_Force;
  h_Pos = new float[4 * (size_t)N];
  h_Vel = new float[4 * (size_t)N];
  h_Force = new float[4 * (size_t)N];
  std::memset(h_Vel, 0, sizeof(float) * 4 * (size_t)N);
  std::memset(h_Force, 0, sizeof(float) * 4 * (size_t)N);
  rdf_dr2 = ljmd_rdf_dr2(N);                 // :93-95
  NdNdr2 = std::vector<int>(LJMD_RDF_BINS, 0);   // :97
  ReallocateMemory();                        // :99 — (re)creates the device handle (before the sampling: it runs there)
  m_device_sampled = false;
  SampleInitialConditions();                 // :90 — on the device: lattice, seeded velocities, forces, parameters
  if (!m_device_sampled) {                   // host sampling (LJMD_HOST_INIT=1): upload and evaluate, :101-103
    check(ljmd_set_state(m_sys, h_Pos, h_Vel), "ljmd_set_state");
    check(ljmd_get_state(m_sys, 0, 0, h_Force), "ljmd_get_state");
  }
  pullScalars(false);
  veloIters = 0;
  initvelo();                                // :107
  momN = momN2 = momN3 = momN4 = 0.;
  momiters = 0;
  av_U_tot = av_T_tot = av_p_tot = 0.;       // :112-113
  av_iters = 0;
}

void MDSystem::ReallocateMemory() {   // MDSystem.cpp:116-144: device buffers follow N
  if (m_sys) ljmd_destroy(m_sys);
  m_sys = 0;
  // device list: m_config.numGPUs > 1 -> devices 0..numGPUs-1; LJMD_DEVICES="0,1,.." (numGPUs == 0 only, so a
  // caller that asks for one device gets one); else the single device LJMD_DEVICE (default 0)
  std::vector<int> devs;
  if (m_config.numGPUs > 1) {
    for (int d = 0; d < m_config.numGPUs; ++d) devs.push_back(d);
  } else if (m_config.numGPUs == 0) {
    if (const char* e = std::getenv("LJMD_DEVICES")) {
      for (const char* p = e; *p;) {
        char* end = 0;
        const long d = std::strtol(p, &end, 10);
        if (end == p) break;
        devs.push_back((int)d);
        p = (*end == ',') ? end + 1 : end;
      }
    }
  }
  if (devs.size() > 1) {
    check(ljmd_create_multi(&m_sys, m_config.N, m_config.rho, m_config.T0, m_config.canonical ? 1 : 0,
                            m_config.boundaryConditions, rdf_dr2, devs.data(), (int)devs.size()),
          "ljmd_create_multi");
  } else {
    int dev = devs.empty() ? 0 : devs[0];
    if (devs.empty())
      if (const char* e = std::getenv("LJMD_DEVICE")) dev = std::atoi(e);
    check(ljmd_create(&m_sys, m_config.N, m_config.rho, m_config.T0, m_config.canonical ? 1 : 0,
                      m_config.boundaryConditions, rdf_dr2, dev),
          "ljmd_create");
  }
  CUDAInit = true;
  m_host_vel_dirty = false;
}

void MDSystem::SampleInitialConditions() {   // MDSystem.cpp:147-181
  // Device path (default): ljmd_init_state samples on the GPU(s) — the reference's lattice bit for bit, Philox
  // velocities keyed by LJMD_SEED (time-seeded like the reference when unset), momentum removed, T = T0 — and
  // evaluates forces and parameters; the host mirrors follow.  LJMD_HOST_INIT=1 keeps the serial host loop below.
  const char* hostinit = std::getenv("LJMD_HOST_INIT");
  if (m_sys && !(hostinit && hostinit[0] == '1')) {
    check(ljmd_set_T0(m_sys, m_config.T0), "ljmd_set_T0");
    check(ljmd_init_state(m_sys, initial_seed()), "ljmd_init_state");
    check(ljmd_get_state(m_sys, h_Pos, h_Vel, h_Force), "ljmd_get_state");
    pullScalars(false);
    m_host_vel_dirty = false;
    m_device_sampled = true;
    return;
  }
  const int N = m_config.N;
  const int Nsingle = (int)std::ceil(std::pow((double)N, 1. / 3.));
  const double dL = L / Nsingle;
  std::mt19937_64 gen(initial_seed());
  std::uniform_real_distribution<double> uni(0., 1.);
  // speeds from the Maxwell distribution by rejection on alpha = exp(-v/sqrt(2T)) (:36-53, :172)
  auto sample_speed = [&]() {
    for (;;) {
      double alpha = uni(gen);
      if (alpha <= 0.) continue;
      const double tl = -std::log(alpha);
      if (uni(gen) * 1.2 < tl * tl * std::exp(-tl * tl) / alpha) return std::sqrt(2. * T) * tl;
    }
  };
  for (int iN = 0; iN < N; ++iN) {
    const int i = 4 * iN;
    h_Pos[i] = (float)(((iN % Nsingle) + 0.5) * dL);                       // :161-167
    h_Pos[i + 1] = (float)((((iN / Nsingle) % Nsingle) + 0.5) * dL);
    h_Pos[i + 2] = (float)(((iN / (Nsingle * Nsingle)) + 0.5) * dL);
    h_Pos[i + 3] = (float)(L / 150.f);                                      // :168 (GL homogeneous w)
    const double v = sample_speed();
    const double cth = 2. * uni(gen) - 1., ph = 2. * kPi * uni(gen);        // :173-177 isotropic direction
    const double sth = std::sqrt(1. - cth * cth);
    h_Vel[i] = (float)(v * sth * std::cos(ph));
    h_Vel[i + 1] = (float)(v * sth * std::sin(ph));
    h_Vel[i + 2] = (float)(v * cth);
    h_Vel[i + 3] = 0.f;
  }
  CorrectTotalMomentum();        // :179
  RenormalizeVelocities(true);   // :180
}

void MDSystem::CorrectTotalMomentum() {   // MDSystem.cpp:183-216
  const int N = m_config.N;
  double px = 0., py = 0., pz = 0.;
  for (int i = 0; i < 4 * N; i += 4) { px += h_Vel[i]; py += h_Vel[i + 1]; pz += h_Vel[i + 2]; }
  const double totmass = N;   // the reference counts unit masses in a double (:186-192)
  for (int i = 0; i < 4 * N; i += 4) {   // float += double: the sum is formed in double and rounded once (:199-201)
    h_Vel[i] += -px / totmass;
    h_Vel[i + 1] += -py / totmass;
    h_Vel[i + 2] += -pz / totmass;
  }
  m_host_vel_dirty = true;
}

void MDSystem::syncConfig() {
  // callers flip these between steps, some by writing m_config directly (semiGCEfluctuations.cpp:58,66)
  check(ljmd_set_canonical(m_sys, m_config.canonical ? 1 : 0), "ljmd_set_canonical");
  check(ljmd_set_boundary(m_sys, m_config.boundaryConditions), "ljmd_set_boundary");
  check(ljmd_set_T0(m_sys, m_config.T0), "ljmd_set_T0");
}

void MDSystem::pullScalars(bool accumulate) {
  double s[LJMD_S_COUNT];
  check(ljmd_get_scalars(m_sys, s), "ljmd_get_scalars");
  U = s[LJMD_S_U]; T = s[LJMD_S_T]; K = s[LJMD_S_K]; V = s[LJMD_S_V]; P = s[LJMD_S_P];
  // P_xy: the reference computes it on its CPU path only and nothing reads it; an extra all-pairs pass per step,
  // so it is filled only on request (LJMD_PSHEAR=1), and is 0 otherwise — what DESIGN.md documents
  static const bool want_shear = std::getenv("LJMD_PSHEAR") && std::getenv("LJMD_PSHEAR")[0] == '1';
  if (want_shear) check(ljmd_get_pshear(m_sys, &Pshear), "ljmd_get_pshear");
  if (accumulate) {   // MDSystem.cpp:355-358
    av_iters++;
    av_U_tot += U;
    av_p_tot += P;
    av_T_tot += T;
  }
}

void MDSystem::CalculateForces() { CalculateForces(h_Force); }   // MDSystem.cpp:232-235

void MDSystem::CalculateForces(float* frc) {   // MDSystem.cpp:237-251 (GPU branch)
  syncConfig();
  check(ljmd_upload(m_sys, h_Pos, h_Vel), "ljmd_upload");
  check(ljmd_compute_forces(m_sys, 0), "ljmd_compute_forces");
  check(ljmd_get_state(m_sys, 0, 0, frc), "ljmd_get_state");
  double s[LJMD_S_COUNT];
  check(ljmd_get_scalars(m_sys, s), "ljmd_get_scalars");
  V = s[LJMD_S_V];
  P = s[LJMD_S_PVIRIAL];   // virial part only until CalculateParameters (:248)
}

double MDSystem::KineticTemperature(float* vel) {   // MDSystem.cpp:361-373
  double ret = 0.;
  for (int i = 0; i < 4 * m_config.N; i += 4) ret += (vel[i] * vel[i] + vel[i + 1] * vel[i + 1] + vel[i + 2] * vel[i + 2]);
  return ret * (1. / 3. / m_config.N);
}

double MDSystem::CalculateXi(float* frc, float* vel) {   // MDSystem.cpp:313-323
  double num = 0., den = 0.;
  for (int i = 0; i < 4 * m_config.N; i += 4) {
    num += vel[i] * frc[i] + vel[i + 1] * frc[i + 1] + vel[i + 2] * frc[i + 2];
    den += vel[i] * vel[i] + vel[i + 1] * vel[i + 1] + vel[i + 2] * vel[i + 2];
  }
  return num / den;
}

void MDSystem::CalculateParameters() {   // MDSystem.cpp:325-359 (GPU branch: V from the force.w column)
  const int N = m_config.N;
  K = 0.;
  V = 0.;
  for (int i = 0; i < 4 * N; i += 4) {
    K += (h_Vel[i] * h_Vel[i] + h_Vel[i + 1] * h_Vel[i + 1] + h_Vel[i + 2] * h_Vel[i + 2]) / 2.;
    V += h_Force[i + 3];
  }
  V *= 4 / 2;
  T = 2. * K / 3. / N;
  P += N * T;
  P /= (N / m_config.rho);
  U = K + V;
  av_iters++;
  av_U_tot += U;
  av_p_tot += P;
  av_T_tot += T;
}

void MDSystem::RenormalizeVelocities(bool RecalculateTkin) {   // MDSystem.cpp:375-389
  double Tkin = T;
  if (RecalculateTkin) Tkin = KineticTemperature(h_Vel);
  const double f = std::sqrt(m_config.T0 / Tkin);
  for (int i = 0; i < 4 * m_config.N; i += 4) { h_Vel[i] *= f; h_Vel[i + 1] *= f; h_Vel[i + 2] *= f; }
  K *= m_config.T0 / Tkin;
  T = m_config.T0;
  U = K + V;
  m_host_vel_dirty = true;
}

void MDSystem::RenormalizeVelocitiesToEnergy(double ust) {   // MDSystem.cpp:391-404
  const double Kold = K, Kdes = ust * m_config.N - V;
  const double f = std::sqrt(Kdes / Kold);
  for (int i = 0; i < 4 * m_config.N; i += 4) { h_Vel[i] *= f; h_Vel[i + 1] *= f; h_Vel[i + 2] *= f; }
  K = Kdes;
  U = K + V;
  m_host_vel_dirty = true;
}

void MDSystem::ApplyBoundaryConditions() {   // MDSystem.cpp:406-436 (host arrays; Integrate applies them on the device)
  const int bc = m_config.boundaryConditions;
  if (bc == 2) return;
  for (int i = 0; i < 4 * m_config.N; i += 4)
    for (int k = 0; k < 3; ++k) {
      if (bc == 0) {
        if (h_Pos[i + k] < 0.) h_Pos[i + k] += L;
        if (h_Pos[i + k] > L) h_Pos[i + k] -= L;
      } else {
        if (h_Pos[i + k] < 0. && h_Vel[i + k] < 0) h_Vel[i + k] = -h_Vel[i + k];
        if (h_Pos[i + k] > L && h_Vel[i + k] > 0) h_Vel[i + k] = -h_Vel[i + k];
      }
    }
  m_host_vel_dirty = true;
}

void MDSystem::Integrate(double dt) {   // MDSystem.cpp:438-583
  syncConfig();
  // host arrays are the caller-visible state: upload them, step, bring them back (the reference's GPU path
  // also crosses the bus every step, MDSystem.cpp:242-250)
  check(ljmd_integrate_host(m_sys, dt, h_Pos, h_Vel, h_Force), "ljmd_integrate_host");
  m_host_vel_dirty = false;
  pullScalars(true);
  t += dt;   // :582
}

void MDSystem::IntegrateMany(double dt, int nsteps) {
  if (nsteps <= 0) return;
  syncConfig();
  check(ljmd_upload(m_sys, h_Pos, h_Vel), "ljmd_upload");
  double s0[LJMD_S_COUNT], s1[LJMD_S_COUNT];
  check(ljmd_get_scalars(m_sys, s0), "ljmd_get_scalars");
  check(ljmd_step(m_sys, dt, nsteps, 0), "ljmd_step");
  check(ljmd_get_state(m_sys, h_Pos, h_Vel, h_Force), "ljmd_get_state");
  check(ljmd_get_scalars(m_sys, s1), "ljmd_get_scalars");
  pullScalars(false);
  // the device kept the per-step running sums (same order of additions as CalculateParameters)
  av_iters += (int)(s1[LJMD_S_AV_ITERS] - s0[LJMD_S_AV_ITERS]);
  av_U_tot += s1[LJMD_S_AV_U_TOT] - s0[LJMD_S_AV_U_TOT];
  av_T_tot += s1[LJMD_S_AV_T_TOT] - s0[LJMD_S_AV_T_TOT];
  av_p_tot += s1[LJMD_S_AV_P_TOT] - s0[LJMD_S_AV_P_TOT];
  t += dt * nsteps;
  m_host_vel_dirty = false;
}

float* MDSystem::getArray(int type) {   // MDSystem.cpp:586-613: device -> host mirror
  if (type == 1) { check(ljmd_get_state(m_sys, 0, h_Vel, 0), "ljmd_get_state"); return h_Vel; }
  check(ljmd_get_state(m_sys, h_Pos, 0, 0), "ljmd_get_state");
  return h_Pos;
}

void MDSystem::setArray(int type, const float* data) {   // MDSystem.cpp:615-630: host -> device
  if (type == 1) check(ljmd_upload(m_sys, 0, data), "ljmd_upload");
  else check(ljmd_upload(m_sys, data, 0), "ljmd_upload");
}

void MDSystem::initvelo(double vmax, double step) {   // MDSystem.cpp:651-669
  const int N = m_config.N;
  m_velo_bins = (int)(vmax / step) + 1;
  m_velo_step = step;
  std::vector<int> dens(m_velo_bins, 0);
  check(ljmd_upload(m_sys, 0, h_Vel), "ljmd_upload");
  check(ljmd_velocity_histogram(m_sys, step, m_velo_bins, dens.data()), "ljmd_velocity_histogram");
  curvelo = SplineFunction();
  for (int i = 0; i < m_velo_bins; ++i) curvelo.add_val(step * (0.5 + i), dens[i] / step / N);
  veloIters = 1;
}

SplineFunction MDSystem::getvelo() { return curvelo; }   // MDSystem.cpp:671-674

void MDSystem::updatevelo() {   // MDSystem.cpp:676-694
  const int N = m_config.N;
  const int maxind = (int)curvelo.vals.size();
  const double shag = curvelo.vals[1].first - curvelo.vals[0].first;
  std::vector<int> dens(maxind, 0);
  // h_Vel is a public member the reference reads here (:684): a caller may have edited it in place since the last
  // call, so it always goes up (16 B per particle, once per GUI frame)
  check(ljmd_upload(m_sys, 0, h_Vel), "ljmd_upload");
  m_host_vel_dirty = false;
  check(ljmd_velocity_histogram(m_sys, shag, maxind, dens.data()), "ljmd_velocity_histogram");
  for (int i = 0; i < maxind; ++i)
    curvelo.vals[i].second = (curvelo.vals[i].second * veloIters + dens[i] / shag / N) / (veloIters + 1);
  veloIters++;
}

double MDSystem::Maxwell(double v) {   // MDSystem.cpp:696-699
  return 4 * kPi * std::pow(1. / 2. / kPi / T, 3. / 2.) * v * v * std::exp(-v * v / 2. / T);
}

void MDSystem::resetAveraging() {   // MDSystem.cpp:701-705
  av_U_tot = av_T_tot = av_p_tot = 0.;
  av_iters = 0;
}

SplineFunction MDSystem::RDF(double, double) {   // MDSystem.cpp:633-649 (arguments ignored there too)
  const int N = m_config.N;
  check(ljmd_get_rdf(m_sys, &NdNdr2[0]), "ljmd_get_rdf");
  const double n0 = N / L / L / L;
  std::vector<double> x, y;
  for (std::size_t ir = 0; ir < NdNdr2.size(); ++ir) {
    const double r = std::sqrt((ir + 0.5) * rdf_dr2);
    x.push_back(r);
    y.push_back(NdNdr2[ir] / rdf_dr2 / 2. / kPi / r / n0 / static_cast<double>(N));
  }
  return SplineFunction(x, y);
}

void MDSystem::Fluctuations(double fraction) {   // MDSystem.cpp:708-730
  const int N = m_config.N;
  const double tsz = L * std::pow(fraction, 1. / 3.);
  long long tN = 0;
  for (int i = 0; i < 4 * N; i += 4) {
    bool in = true;
    for (int k = 0; k < 3; ++k) in = in && h_Pos[i + k] >= 0.5 * L - 0.5 * tsz && h_Pos[i + k] <= 0.5 * L + 0.5 * tsz;
    if (in) tN++;
  }
  momN += tN; momN2 += tN * tN; momN3 += tN * tN * tN; momN4 += tN * tN * tN * tN;
  momiters++;
  const double ev = 1. - 2. * kPi / 3. * N / L / L / L;
  std::cout << "Iteration: " << momiters << "\t" << "<N> = " << momN / momiters << "\t" << "w[N] = "
            << (momN2 / momiters - (momN / momiters) * (momN / momiters)) / (momN / momiters) << "\t" << "EV: "
            << ev * ev << "\t" << "Binom: " << 1. - fraction << "\n";
}

int MDSystem::fast_round(float x) {   // MDSystem.cpp:732-739
  return x > 0 ? static_cast<int>(x + 0.5f) : static_cast<int>(x - 0.5f);
}

void MDSystem::setHardwareMode(bool useCUDA) {   // MDSystem.cpp:741-745
  m_config.useCUDA = useCUDA;   // recorded; there is no CPU backend to switch to
}
