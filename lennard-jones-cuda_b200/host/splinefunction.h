// Piecewise-linear table returned by MDSystem::RDF() / getvelo().
// Same public surface as the reference's helper type (/root/reference/src/library/splinefunction.h:7-79):
// callers read `vals` directly and call f(), add_val(), clear(), fill(), setConstant(), loadFromFile().
#ifndef SPLINEFUNCTION_H
#define SPLINEFUNCTION_H
#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

class SplineFunction {
 public:
  std::vector<std::pair<double, double> > vals;  // (x, y), kept sorted by x

  SplineFunction() {}
  SplineFunction(const std::vector<double>& x, const std::vector<double>& y) { fill(x, y); }

  void fill(const std::vector<double>& x, const std::vector<double>& y) {
    vals.clear();
    vals.reserve(x.size());
    for (std::size_t i = 0; i < x.size(); ++i) vals.push_back(std::make_pair(x[i], y[i]));
    std::sort(vals.begin(), vals.end());
  }
  void add_val(double x, double val) {
    // insertion keeps the order without re-sorting the whole table
    std::pair<double, double> e(x, val);
    vals.insert(std::upper_bound(vals.begin(), vals.end(), e), e);
  }
  // Linear interpolation between the bracketing points; linear extrapolation from the two end points.
  double f(double arg) const {
    const std::size_t n = vals.size();
    std::size_t hi = std::lower_bound(vals.begin(), vals.end(), std::make_pair(arg, 0.)) - vals.begin();
    if (hi == 0) hi = 1;
    if (hi == n) hi = n - 1;
    const std::pair<double, double>&a = vals[hi - 1], &b = vals[hi];
    return a.second + (arg - a.first) * (b.second - a.second) / (b.first - a.first);
  }
  double fsquare(double arg) const { const double r = f(arg); return r * r; }
  void clear() { vals.assign(2, std::make_pair(0., 0.)); vals[1].first = 1.; }
  void clearall() { vals.clear(); }
  void setConstant(double val) { vals.clear(); vals.push_back(std::make_pair(0., val)); vals.push_back(std::make_pair(1., val)); }
  void loadFromFile(const char* file);
};

#endif  // SPLINEFUNCTION_H
