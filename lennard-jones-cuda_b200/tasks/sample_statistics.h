// Sample statistics of one observable with delta-method standard errors: what the fluctuation tasks print.
//
// Covers the part of the reference's `SampleMoments::NumberStatistics`
// (/root/reference/src/extra/sample-moments/NumberStatistics.h) that its tasks call — AddObservation, GetMean,
// GetMeanError (:301-304), GetVariance[Error] (:307-310), GetScaledVariance[Error] (:313-316, :277-296),
// GetSkewness[Error] (:320-322), GetKurtosis[Error] (:325-327) — with the same estimators:
//   raw moment sums S_k = sum x^k (no mean shift), m_k = S_k / n, central moments by binomial expansion,
//   Cov(mu_r, mu_q) = (mu_{r+q} - mu_r mu_q + r q mu_2 mu_{r-1} mu_{q-1} - r mu_{r-1} mu_{q+1} - q mu_{r+1} mu_{q-1}) / n,
//   Cov(mu_r, m_1)  = (mu_{r+1} - r mu_2 mu_{r-1}) / n,
// and first-order error propagation for ratios.  Moments up to order 8 are kept (the kurtosis error needs mu_8).
#ifndef LJMD_TASKS_SAMPLE_STATISTICS_H
#define LJMD_TASKS_SAMPLE_STATISTICS_H
#include <cmath>
#include <cstdint>

namespace ljtasks {

class SampleStatistics {
 public:
  static const int kOrder = 8;
  SampleStatistics() { clear(); }
  void clear() {
    n_ = 0;
    for (int k = 0; k <= kOrder; ++k) sum_[k] = 0.;
    fresh_ = false;
  }
  void add(double x) {
    double p = 1.;
    for (int k = 0; k <= kOrder; ++k) { sum_[k] += p; p *= x; }
    ++n_;
    fresh_ = false;
  }
  int64_t count() const { return n_; }

  double mean() { refresh(); return m_[1]; }
  double mean_error() { refresh(); return std::sqrt((m_[2] - m_[1] * m_[1]) / (double)n_); }
  double central(int r) { refresh(); return mu_[r]; }
  double variance() { return central(2); }
  double variance_error() { return std::sqrt(cov_central(2, 2)); }
  // omega = mu_2 / mean
  double scaled_variance() { return variance() / mean(); }
  double scaled_variance_error() {
    const double a = central(2), b = mean();
    const double va = cov_central(2, 2), vb = mean_error() * mean_error(), cab = cov_central_mean(2);
    return std::fabs(a / b) * std::sqrt(va / a / a + vb / b / b - 2. * cab / a / b);
  }
  // kappa_3 / kappa_2 = mu_3 / mu_2
  double skewness() { return central(3) / central(2); }
  double skewness_error() {
    const double a = central(3), b = central(2);
    const double var = cov_central(3, 3) / (b * b) + a * a * cov_central(2, 2) / (b * b * b * b) -
                       2. * a * cov_central(3, 2) / (b * b * b);
    return std::sqrt(var);
  }
  // kappa_4 / kappa_2 = (mu_4 - 3 mu_2^2) / mu_2
  double kurtosis() { const double b = central(2); return (central(4) - 3. * b * b) / b; }
  double kurtosis_error() {
    const double b = central(2), c = central(4);
    const double g4 = 1. / b, g2 = -(3. * b * b + c) / (b * b);
    const double var = g4 * g4 * cov_central(4, 4) + g2 * g2 * cov_central(2, 2) + 2. * g4 * g2 * cov_central(4, 2);
    return std::sqrt(var);
  }

 private:
  int64_t n_;
  double sum_[kOrder + 1], m_[kOrder + 1], mu_[kOrder + 2];
  bool fresh_;

  static double binom(int n, int k) {
    double b = 1.;
    for (int i = 1; i <= k; ++i) b = b * (double)(n - k + i) / (double)i;
    return std::floor(b + 0.5);
  }
  void refresh() {
    if (fresh_) return;
    for (int k = 0; k <= kOrder; ++k) m_[k] = n_ > 0 ? sum_[k] / (double)n_ : 0.;
    mu_[0] = 1.; mu_[1] = 0.;
    for (int r = 2; r <= kOrder; ++r) {
      double acc = 0., pw = 1.;   // sum_i C(r,i) m_{r-i} (-m_1)^i
      for (int i = 0; i <= r; ++i) {
        acc += binom(r, i) * m_[r - i] * pw * ((i & 1) ? -1. : 1.);
        pw *= m_[1];
      }
      mu_[r] = acc;
    }
    mu_[kOrder + 1] = 0.;
    fresh_ = true;
  }
  double cov_central(int r, int q) {
    refresh();
    if (r <= 1 || q <= 1 || r + q > kOrder) return 0.;
    double c = mu_[r + q] - mu_[r] * mu_[q] + (double)(r * q) * mu_[2] * mu_[r - 1] * mu_[q - 1] -
               (double)r * mu_[r - 1] * mu_[q + 1] - (double)q * mu_[r + 1] * mu_[q - 1];
    return c / (double)n_;
  }
  double cov_central_mean(int r) {
    refresh();
    return (mu_[r + 1] - (double)r * mu_[2] * mu_[r - 1]) / (double)n_;
  }
};

// Time average of a correlated series with the statistical-inefficiency correction of the reference's
// `TimeAverage` (/root/reference/src/tasks/auxiliary/time-average-aux.h:27-67; Allen & Tildesley pp. 194-195):
// s = 2 / ln(var / C_1) with the lag-one autocovariance C_1 = <x_k x_{k+1}> - <x>^2, s = 1 when that is negative.
class CorrelatedAverage {
 public:
  SampleStatistics stats;
  CorrelatedAverage() : prev_(0.), lag1_(0.), n_(0) {}
  void add(double x) {
    stats.add(x);
    if (n_ > 0) lag1_ += prev_ * x;
    prev_ = x;
    ++n_;
  }
  double inefficiency() {
    const double mu = stats.mean();
    const double c1 = lag1_ / (double)(n_ - 1) - mu * mu;
    double s = 2. / std::log(stats.variance() / c1);
    if (s < 0.) s = 1.;
    return s;
  }
  double mean() { return stats.mean(); }
  double mean_error() { return stats.mean_error() * std::sqrt(inefficiency()); }
  int count() const { return n_; }

 private:
  double prev_, lag1_;
  int n_;
};

}  // namespace ljtasks
#endif
