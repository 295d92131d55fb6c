// semiGCEfluctuations: particle-number cumulants in centred cubic sub-volumes of a microcanonical system that
// was equilibrated canonically ("semi-grand-canonical" sampling): one observation every 200 steps.
//
// Same constants and console table as the reference's driver
// (/root/reference/src/tasks/semiGCEfluctuations/semiGCEfluctuations.cpp:27-106).  Optional arguments (an
// extension) override them:  semiGCEfluctuations [events [N [T* [rho*]]]].
// The 200 steps between observations stay on the device; the occupancies of all nested cubes come from one
// ljmd_subvolume_counts call (the reference downloads h_Pos and loops over it once per fraction, :7-25,:81-83).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "MDSystem.h"
#include "ljmd.h"
#include "sample_statistics.h"

using namespace ljtasks;

static void must(int rc, const char* what) {
  if (rc != LJMD_OK) {
    std::fprintf(stderr, "%s failed: %s\n", what, ljmd_last_error());
    std::exit(1);
  }
}

int main(int argc, char* argv[]) {
  const int nev = argc > 1 ? std::atoi(argv[1]) : 10000;
  const int N = argc > 2 ? std::atoi(argv[2]) : 512;
  const double Tst = argc > 3 ? std::atof(argv[3]) : 1.312;
  const double rhost = argc > 4 ? std::atof(argv[4]) : 0.316;
  const double dt = 0.005;
  const int iterspreeq = 10000, itersstep = 200;

  std::vector<double> fracs;
  for (double tfr = 0.05; tfr <= 1.; tfr += 0.05) fracs.push_back(tfr);   // :30-32
  std::vector<SampleStatistics> stats(fracs.size());

  if (ljmd_device_count() == 0) {
    std::fprintf(stderr, "Could not find a CUDA device! This build has no CPU path.\n");
    return 1;
  }
  MDSystem::MDSystemConfiguration config;
  config.N = N;
  config.T0 = Tst;
  config.rho = rhost;
  config.useCUDA = true;
  MDSystem syst(config);
  syst.Reinitialize(config);
  ljmd_system* h = syst.handle();
  must(ljmd_set_state(h, syst.h_Pos, syst.h_Vel), "ljmd_set_state");

  must(ljmd_set_canonical(h, 1), "ljmd_set_canonical");   // equilibration, :58-62
  must(ljmd_step(h, dt, iterspreeq, 0), "ljmd_step (equilibration)");
  must(ljmd_set_canonical(h, 0), "ljmd_set_canonical");   // production, :66

  std::vector<int> cum(256);
  for (int iN = 1; iN <= nev; ++iN) {
    must(ljmd_step(h, dt, itersstep, 0), "ljmd_step");
    int nb = 0;
    must(ljmd_subvolume_counts(h, 3, 0.05, cum.data(), (int)cum.size(), &nb), "ljmd_subvolume_counts");
    for (std::size_t i = 0; i < fracs.size(); ++i) {
      // nested cubes: entry i of the cumulative table is the cube of volume fraction fracs[i]; the device grid
      // stops below alpha = 1, the whole box holds all N
      const double n = (int)i < nb ? (double)cum[i] : (double)N;
      stats[i].add(n);
    }
    if (iN % 10 == 0) {
      for (std::size_t i = 0; i < fracs.size(); ++i) {
        std::printf("%15d %10lf +- %-10lf %10lf +- %-10lf %10lf +- %-10lf %10lf +- %-10lf\n", iN, stats[i].mean(),
                    stats[i].mean_error(), stats[i].scaled_variance(), stats[i].scaled_variance_error(),
                    stats[i].scaled_variance() / (1. - fracs[i]), stats[i].scaled_variance_error() / (1. - fracs[i]),
                    stats[i].skewness() / (1. - 2. * fracs[i]), stats[i].skewness_error() / (1. - 2. * fracs[i]));
      }
      std::printf("\n");
    }
  }
  return 0;
}
