// run-fluctuations: particle-number fluctuations in coordinate- and momentum-space sub-volumes of a
// Lennard-Jones fluid, canonical (TVN) or microcanonical (EVN) ensemble.
//
// Same command line, parameter file, console table and output files as the reference's driver
// (/root/reference/src/tasks/run-fluctuations/run-fluctuations.cpp:10-202), but the production loop is
// device-resident: 1000 steps per ljmd_step call with the RDF histogram accumulated on the device and the
// per-step observables (five occupancy families, the alpha = 1/2 slab count, U, T, P, mean velocity) recorded
// by the observation trace — no h_Pos / h_Vel download per step (the reference's loop :120-135 reads both).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "MDSystem.h"
#include "ljmd.h"
#include "fluctuation_observers.h"
#include "task_parameters.h"

using namespace ljtasks;

static void must(int rc, const char* what) {
  if (rc != LJMD_OK) {
    std::fprintf(stderr, "%s failed: %s\n", what, ljmd_last_error());
    std::exit(1);
  }
}

// LJMD_TASK_TIMING=1: wall-clock stamps of the phases on stderr
static void stamp(const char* what) {
  static const bool on = std::getenv("LJMD_TASK_TIMING") != NULL;
  static const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  if (on) std::fprintf(stderr, "[%8.3f s] %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
}

// how many `t += dt` it takes until the loop condition of the reference stops holding
static long steps_until(double& t, double tend, double dt, long limit) {
  long n = 0;
  while (t < tend && n < limit) { t += dt; ++n; }
  return n;
}

int main(int argc, char* argv[]) {
  TaskParameters par = TaskParameters::fluctuations();
  if (argc > 1) {
    std::cout << "Reading parameters from file " << argv[1] << std::endl;
    par.read(argv[1]);
  }
  par.output_prefix = par.stamped_prefix(true);
  if (ljmd_device_count() == 0) {
    std::cerr << "Could not find a CUDA device! This build has no CPU path.\n";
    return 1;
  }
  par.print();
  stamp("start");

  const int N = (int)par.integer("N");
  const double dt = par["dt*"], teq = par["teq"], tfin = par["tfin"], dalpha = par["subvolume_spacing"];
  const int report_every = 1000;   // :137

  MDSystem::MDSystemConfiguration config;
  config.N = N;
  config.T0 = par["T*"];
  config.rho = par["rho*"];
  config.useCUDA = true;
  config.canonical = par.integer("canonical") != 0;
  MDSystem syst(config);
  syst.Reinitialize(config);
  if (!config.canonical) syst.RenormalizeVelocitiesToEnergy(par["u*"]);
  stamp("system constructed");

  // from here on everything runs on the library handle; the class instance only provided the initial state
  ljmd_system* h = syst.handle();
  must(ljmd_set_state(h, syst.h_Pos, syst.h_Vel), "ljmd_set_state");

  double t = 0.;
  {
    const long neq = steps_until(t, teq, dt, 2000000000L);   // :84-87
    must(ljmd_step(h, dt, (int)neq, 0), "ljmd_step (equilibration)");
  }
  stamp("equilibrated");

  const double vfactor = 3.0;   // MomentumFlucsAverage::Vfactor (:436)
  const double vcut_max = std::sqrt(config.T0) * vfactor;
  OccupancySeries occX(coordinate_fractions(dalpha)), occY(coordinate_fractions(dalpha)),
      occZ(coordinate_fractions(dalpha)), occCube(coordinate_fractions(dalpha)),
      occVz(momentum_cuts(config.T0, dalpha, vfactor));
  // counters: x, y, z slabs, cube, |vz|, and the alpha = 1/2 z slab of GetNSubsystem(syst, 0.5, 2) (:133)
  const int kinds[6] = {0, 1, 2, 3, 6, 2};
  const double steps[6] = {dalpha, dalpha, dalpha, dalpha, dalpha, 0.5};
  const double vcuts[6] = {0., 0., 0., 0., vcut_max, 0.};
  must(ljmd_trace_begin(h, 6, kinds, steps, vcuts, report_every), "ljmd_trace_begin");
  int row = 0;
  must(ljmd_trace_row_length(h, &row), "ljmd_trace_row_length");
  const std::size_t nc = occX.grid.size(), nv = occVz.grid.size();
  if ((std::size_t)row != 4 * nc + nv + 1) {
    std::fprintf(stderr, "trace row has %d counts, expected %zu\n", row, 4 * nc + nv + 1);
    return 1;
  }

  static const char* console_cols[] = {"t*", "u*", "T*", "Z", "<u*>", "<T*>", "<Z>", "<w>/(1-x)"};
  for (int c = 0; c < 8; ++c) std::cout << std::setw(15) << console_cols[c] << " ";
  std::cout << std::endl;
  std::ofstream fout((par.output_prefix + ".TimeDep.txt").c_str());
  static const char* file_cols[] = {"t*", "u*", "T*", "Z", "<u*>", "<T*>", "<Z>", "<wx>/(1-x)", "<wy>/(1-x)",
                                    "<wz>/(1-x)", "<vx>", "<vy>", "<vz>"};
  for (int c = 0; c < 13; ++c) fout << std::setw(15) << file_cols[c] << " ";
  fout << std::endl;

  must(ljmd_reset_averaging(h), "ljmd_reset_averaging");
  {
    long long dummy[256];
    int ns = 0;
    must(ljmd_get_rdf_accum(h, dummy, &ns, 1), "ljmd_get_rdf_accum");   // start the RDF mean at the production phase
  }
  std::vector<double> scal((std::size_t)report_every * LJMD_TRACE_SCALARS), mvel((std::size_t)report_every * 3);
  std::vector<long long> counts((std::size_t)report_every * row);
  double tot_N = 0., tot_N2 = 0.;
  long totIters = 0;
  const float dr2 = ljmd_rdf_dr2(N);
  while (t < tfin || tfin < 0.) {
    long chunk = report_every - totIters % report_every;
    if (tfin >= 0.) chunk = steps_until(t, tfin, dt, chunk);
    else t += dt * chunk;
    must(ljmd_step(h, dt, (int)chunk, 1), "ljmd_step");
    int got = 0;
    must(ljmd_trace_read(h, report_every, &got, scal.data(), counts.data(), mvel.data()), "ljmd_trace_read");
    stamp("batch stepped and trace read");
    for (int k = 0; k < got; ++k) {
      const long long* r = counts.data() + (std::size_t)k * row;
      occX.add_step(r);
      occY.add_step(r + nc);
      occZ.add_step(r + 2 * nc);
      occCube.add_step(r + 3 * nc);
      occVz.add_step(r + 4 * nc);
      const double half = (double)r[4 * nc + nv];
      tot_N += half;
      tot_N2 += half * half;
    }
    totIters += got;
    if (got == 0 || totIters % report_every != 0) continue;

    double sc[LJMD_S_COUNT];
    must(ljmd_get_scalars(h, sc), "ljmd_get_scalars");
    const double rho = config.rho;
    const double u = sc[LJMD_S_U] / N, T = sc[LJMD_S_T], Z = sc[LJMD_S_P] / (rho * T);
    const double uav = sc[LJMD_S_AV_U_TOT] / sc[LJMD_S_AV_ITERS] / N, Tav = sc[LJMD_S_AV_T_TOT] / sc[LJMD_S_AV_ITERS];
    const double Zav = sc[LJMD_S_AV_P_TOT] / sc[LJMD_S_AV_ITERS] / (rho * Tav);
    const double Nav = tot_N / totIters, N2av = tot_N2 / totIters;
    std::cout << std::setw(15) << t << " " << std::setw(15) << u << " " << std::setw(15) << T << " " << std::setw(15) << Z
              << " " << std::setw(15) << uav << " " << std::setw(15) << Tav << " " << std::setw(15) << Zav << " "
              << std::setw(15) << (N2av - Nav * Nav) / Nav / (1. - 0.5) << " " << std::endl;
    fout << std::setw(15) << t << " " << std::setw(15) << u << " " << std::setw(15) << T << " " << std::setw(15) << Z << " "
         << std::setw(15) << uav << " " << std::setw(15) << Tav << " " << std::setw(15) << Zav << " ";
    OccupancySeries* slabs[3] = {&occX, &occY, &occZ};
    for (int a = 0; a < 3; ++a) {
      const std::size_t mid = slabs[a]->grid.size() / 2;   // :163-168
      fout << std::setw(15) << slabs[a]->omega_running(mid) / (1. - slabs[a]->grid[mid]) << " ";
    }
    const double* mv = mvel.data() + (std::size_t)(got - 1) * 3;
    fout << std::setw(15) << mv[0] << " " << std::setw(15) << mv[1] << " " << std::setw(15) << mv[2] << " " << std::endl;
    fout.flush();

    long long hist[256];
    int samples = 0;
    must(ljmd_get_rdf_accum(h, hist, &samples, 0), "ljmd_get_rdf_accum");
    write_rdf_file(par.output_prefix + ".RDF.dat", hist, samples, N, sc[LJMD_S_L], dr2);
    occX.write_coordinate_file(par.output_prefix + ".flucsX.dat");
    occY.write_coordinate_file(par.output_prefix + ".flucsY.dat");
    occZ.write_coordinate_file(par.output_prefix + ".flucsZ.dat");
    occCube.write_coordinate_file(par.output_prefix + ".flucsCube.dat");
    occVz.write_momentum_file(par.output_prefix + ".flucsVz.dat", N);
    stamp("report written");
  }
  must(ljmd_trace_end(h), "ljmd_trace_end");
  return 0;
}
