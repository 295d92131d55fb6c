// Accumulators of the fluctuation run, fed from the device-side observation trace (include/ljmd.h,
// ljmd_trace_*) instead of from per-step host passes over h_Pos / h_Vel.  Output files keep the column layout
// of the reference's RDF_average / CoordFlucsAverage / MomentumFlucsAverage ::PrintToFile
// (/root/reference/src/tasks/run-fluctuations/include/run-fluctuations-aux.h:297-330, 379-421, 480-527).
#ifndef LJMD_TASKS_FLUCTUATION_OBSERVERS_H
#define LJMD_TASKS_FLUCTUATION_OBSERVERS_H
#include <cmath>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

#include "sample_statistics.h"

namespace ljtasks {

// The fraction grids, built by repeated addition exactly as the reference's loops do (the device side of
// ljmd_subvolume_counts restates the same loops, so both agree on the number of entries).
inline std::vector<double> coordinate_fractions(double step) {     // :345-353
  std::vector<double> a;
  for (double alpha = step; alpha <= 1. - 1.e-9; alpha += step) a.push_back(alpha);
  return a;
}
inline std::vector<double> momentum_cuts(double T0, double step, double vfactor) {   // :446-456
  std::vector<double> v;
  for (double alpha = step; alpha <= 1. + 1.e-9; alpha += step) v.push_back(std::sqrt(T0) * alpha * vfactor);
  return v;
}

// Occupancy statistics of a family of nested sub-volumes (one CorrelatedAverage per fraction).
class OccupancySeries {
 public:
  std::vector<double> grid;      // alpha (coordinate space) or vcut (momentum space)
  std::vector<double> sumN, sumN2;
  std::vector<CorrelatedAverage> series;
  long steps;
  explicit OccupancySeries(const std::vector<double>& g = std::vector<double>())
      : grid(g), sumN(g.size(), 0.), sumN2(g.size(), 0.), series(g.size()), steps(0) {}
  // one trace row: cumulative counts, one per grid entry
  void add_step(const long long* counts) {
    for (std::size_t k = 0; k < grid.size(); ++k) {
      const double n = (double)counts[k];
      sumN[k] += n;
      sumN2[k] += n * n;
      series[k].add(n);
    }
    ++steps;
  }
  // scaled variance at grid entry k from the plain running sums (what the time-dependence file shows)
  double omega_running(std::size_t k) const {
    const double m = sumN[k] / steps, m2 = sumN2[k] / steps;
    return (m2 - m * m) / m;
  }

  void write_coordinate_file(const std::string& path) {
    std::ofstream out(path.c_str());
    if (out.is_open()) {
      static const char* cols[] = {"alpha", "<N>", "w[N]", "w[N]/(1-alpha)", "error", "s"};
      for (int c = 0; c < 6; ++c) out << std::setw(15) << cols[c] << " ";
    }
    out << std::endl;
    for (std::size_t k = 0; k < grid.size(); ++k) {
      const double alpha = grid[k];
      const double w = series[k].stats.scaled_variance(), s = series[k].inefficiency();
      out << std::setw(15) << alpha << " " << std::setw(15) << series[k].mean() << " " << std::setw(15) << w << " "
          << std::setw(15) << w / (1. - alpha) << " "
          << std::setw(15) << series[k].stats.scaled_variance_error() * std::sqrt(s) / (1. - alpha) << " "
          << std::setw(15) << s << " " << std::endl;
    }
  }
  void write_momentum_file(const std::string& path, int Ntotal) {
    std::ofstream out(path.c_str());
    if (out.is_open()) {
      static const char* cols[] = {"vcut", "<N>", "alpha", "w[N]", "w[N]/(1-alpha)", "error", "s"};
      for (int c = 0; c < 7; ++c) out << std::setw(15) << cols[c] << " ";
    }
    out << std::endl;
    for (std::size_t k = 0; k < grid.size(); ++k) {
      const double mean = series[k].mean(), alpha = mean / Ntotal;
      const double w = series[k].stats.scaled_variance(), s = series[k].inefficiency();
      out << std::setw(15) << grid[k] << " " << std::setw(15) << mean << " " << std::setw(15) << alpha << " "
          << std::setw(15) << w << " " << std::setw(15) << w / (1. - alpha) << " "
          << std::setw(15) << series[k].stats.scaled_variance_error() * std::sqrt(s) / (1. - alpha) << " "
          << std::setw(15) << s << " " << std::endl;
    }
  }
};

// Mean radial distribution function from the device's accumulated r^2 histogram: g is linear in the counts, so
// the mean of the per-step curves (RDF_average::AddTimeStep, :303-313) is the curve of the mean histogram
// (MDSystem::RDF, /root/reference/src/library/MDSystem.cpp:633-649).
inline void write_rdf_file(const std::string& path, const long long* hist256, int samples, int N, double L, double dr2) {
  std::ofstream out(path.c_str());
  if (out.is_open()) out << std::setw(15) << "r*" << " " << std::setw(15) << "G(r)" << " ";
  out << std::endl;
  const double pi = 3.14159265358979323846, n0 = N / L / L / L;
  for (int k = 0; k < 256; ++k) {
    const double r = std::sqrt((k + 0.5) * dr2);
    const double g = samples > 0 ? ((double)hist256[k] / samples) / dr2 / 2. / pi / r / n0 / (double)N : 0.;
    out << std::setw(15) << r << " " << std::setw(15) << g << " " << std::endl;
  }
}

}  // namespace ljtasks
#endif
