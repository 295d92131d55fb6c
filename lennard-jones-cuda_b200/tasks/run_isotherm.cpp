// run-isotherm: equation of state along an isotherm in the canonical ensemble, one simulation per density.
//
// Same command line, parameter file, console table and `.dat` columns as the reference's driver
// (/root/reference/src/tasks/run-isotherm/run-isotherm.cpp:10-170).  The reference observes u* and p* before
// every Integrate (:103-106); here the steps run in device-resident batches of up to 1000 and the per-step
// scalars come back through the observation trace (ljmd_trace_*), so the series fed to the correlated
// averages is the same one, without a host round trip per step.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <vector>

#include "MDSystem.h"
#include "ljmd.h"
#include "sample_statistics.h"
#include "task_parameters.h"

using namespace ljtasks;

static void must(int rc, const char* what) {
  if (rc != LJMD_OK) {
    std::fprintf(stderr, "%s failed: %s\n", what, ljmd_last_error());
    std::exit(1);
  }
}

template <class Stream>
static void header(Stream& o, bool with_w) {
  static const char* cols[] = {"rho*", "t*", "T*", "<u*>", "d<u*>", "teq_u*", "<p*>", "d<p*>", "teq_p*", "<Z>", "d<Z>"};
  for (int c = 0; c < 11; ++c) o << std::setw(15) << cols[c] << " ";
  if (with_w) o << std::setw(15) << "w[N]" << " " << std::setw(15) << "dw[N]" << " ";
  o << std::endl;
}

template <class Stream>
static void row(Stream& o, double rho, double t, double T, CorrelatedAverage& u, CorrelatedAverage& p, double dt, double T0) {
  o << std::setw(15) << rho << " " << std::setw(15) << t << " " << std::setw(15) << T << " "
    << std::setw(15) << u.mean() << " " << std::setw(15) << u.mean_error() << " " << std::setw(15) << u.inefficiency() * dt << " "
    << std::setw(15) << p.mean() << " " << std::setw(15) << p.mean_error() << " " << std::setw(15) << p.inefficiency() * dt << " "
    << std::setw(15) << p.mean() / rho / T0 << " " << std::setw(15) << p.mean_error() / rho / T0 << " ";
}

int main(int argc, char* argv[]) {
  TaskParameters par = TaskParameters::isotherm();
  if (argc > 1) {
    std::cout << "Reading parameters from file " << argv[1] << std::endl;
    par.read(argv[1]);
  }
  par.output_prefix = par.stamped_prefix(false);
  if (ljmd_device_count() == 0) {
    std::cerr << "Could not find a CUDA device! This build has no CPU path.\n";
    return 1;
  }
  par.print();

  const int N = (int)par.integer("N");
  const double T0 = par["T*"], rhomin = par["rho*_min"], rhomax = par["rho*_max"], drho = par["drho*"];
  const double dt = par["dt*"], teq = par["teq"], tfin = par["tfin"];
  const int batch = 1000;

  header(std::cout, false);
  std::ofstream fout((par.output_prefix + ".dat").c_str());
  header(fout, true);

  double Pprev = 0., Ppreverr = 0.;
  std::vector<double> scal((std::size_t)batch * LJMD_TRACE_SCALARS);
  for (double rho = rhomin; rho <= rhomax; rho += drho) {   // :89
    std::cout << std::endl;
    MDSystem::MDSystemConfiguration config;
    config.N = N;
    config.T0 = T0;
    config.rho = rho;
    config.useCUDA = true;
    config.canonical = true;
    MDSystem syst(config);
    syst.Reinitialize(config);
    ljmd_system* h = syst.handle();
    must(ljmd_set_state(h, syst.h_Pos, syst.h_Vel), "ljmd_set_state");

    double t = 0.;
    long neq = 0;
    while (t < teq) { t += dt; ++neq; }   // :103-106
    must(ljmd_step(h, dt, (int)neq, 0), "ljmd_step (equilibration)");

    // number of production iterations of `while (t < tfin) { observe; Integrate; t += dt; }`
    long niter = 0;
    for (double tt = t; tt < tfin; tt += dt) ++niter;

    double sc[LJMD_S_COUNT];
    must(ljmd_get_scalars(h, sc), "ljmd_get_scalars");
    double U = sc[LJMD_S_U], P = sc[LJMD_S_P], T = sc[LJMD_S_T];
    must(ljmd_trace_begin(h, 0, NULL, NULL, NULL, batch), "ljmd_trace_begin");
    CorrelatedAverage uav, Pav;
    long done = 0;        // Integrate calls issued so far
    int have = 0, next = 0;   // rows of the current batch / next unread row
    for (long it = 0; it < niter; ++it) {
      uav.add(U / N);
      Pav.add(P);
      if ((it + 1) % 1000 == 0) {
        row(std::cout, rho, t, T, uav, Pav, dt, T0);
        std::cout << std::endl;
      }
      if (next == have) {   // run the next batch of steps
        const long chunk = std::min<long>(batch, niter - done);
        must(ljmd_step(h, dt, (int)chunk, 0), "ljmd_step");
        must(ljmd_trace_read(h, batch, &have, scal.data(), NULL, NULL), "ljmd_trace_read");
        done += chunk;
        next = 0;
      }
      const double* r = scal.data() + (std::size_t)next * LJMD_TRACE_SCALARS;   // t, U, T, P, ...
      U = r[1]; T = r[2]; P = r[3];
      ++next;
      t += dt;
    }
    must(ljmd_trace_end(h), "ljmd_trace_end");

    row(fout, rho, t - dt, T, uav, Pav, dt, T0);
    const double Pcur = Pav.mean(), Pcurerr = Pav.mean_error();
    const double tdrho = (rho == rhomin) ? rhomin : drho;   // :148-150
    const double w = T0 / ((Pcur - Pprev) / tdrho);
    const double werr = T0 * tdrho * std::sqrt(Pcurerr * Pcurerr + Ppreverr * Ppreverr) / std::fabs(Pcur - Pprev);
    fout << std::setw(15) << w << " " << std::setw(15) << werr << " " << std::endl;
    fout.flush();
    Pprev = Pcur;
    Ppreverr = Pcurerr;
  }
  return 0;
}
