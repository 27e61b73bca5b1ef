// rdk_host_math.cpp -- scalar host utilities of the C ABI.
//
// rdk_compute_gamma_cats replaces corax_compute_gamma_cats (reference call
// sites src/model.cpp:239,248,257,266): Yang (1994) equal-probability discrete
// Gamma; MEAN = conditional means through the regularised incomplete gamma at
// shape alpha+1, MEDIAN = quantile midpoints renormalised to mean 1.  Quantiles
// come from the chi-square percentage-point algorithm AS 91 (Best & Roberts
// 1975) with the normal deviate of AS 70 and the incomplete gamma integral of
// AS 32 (Bhattacharjee 1970), i.e. the routines libpll-2/coraxlib take from
// PAML.  Called O(1) times per optimiser step: host scalar code on purpose.
#include "../../include/rdk.h"

#include <cmath>
#include <cstdio>
#include <vector>

namespace {

void set_err(const char *m) {
  rdk_errno = RDK_ERROR_PARAM;
  snprintf(rdk_errmsg, 200, "%s", m);
}

// ln Gamma(x), x > 0: shift to x >= 7 then Stirling's series (Pike & Hill 1966)
double log_gamma(double x) {
  double shift = 0.0;
  if (x < 7.0) {
    double prod = 1.0;
    double z = x - 1.0;
    while (++z < 7.0) prod *= z;
    x = z;
    shift = -std::log(prod);
  }
  const double z = 1.0 / (x * x);
  const double series =
      (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
  return shift + (x - 0.5) * std::log(x) - x + .918938533204673 + series;
}

// regularised lower incomplete gamma P(shape, x); lg = ln Gamma(shape)
double reg_inc_gamma(double x, double shape, double lg) {
  const double tol = 1e-8, big = 1e30;
  if (x == 0) return 0;
  if (x < 0 || shape <= 0) return -1;
  const double front = std::exp(shape * std::log(x) - x - lg);
  if (!(x > 1 && x >= shape)) {
    // power series
    double sum = 1, term = 1, d = shape;
    do {
      d++;
      term *= x / d;
      sum += term;
    } while (term > tol);
    return sum * (front / shape);
  }
  // continued fraction
  double a = 1 - shape, b = a + x + 1, n = 0;
  double pn[6] = {1, x, x + 1, x * b, 0, 0};
  double cur = pn[2] / pn[3];
  for (;;) {
    a++;
    b += 2;
    n++;
    const double an = a * n;
    pn[4] = b * pn[2] - an * pn[0];
    pn[5] = b * pn[3] - an * pn[1];
    if (pn[5] != 0) {
      const double next = pn[4] / pn[5];
      const double diff = std::fabs(cur - next);
      if (diff <= tol && diff <= tol * next) break;
      cur = next;
    }
    for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
    if (std::fabs(pn[4]) >= big)
      for (int i = 0; i < 4; ++i) pn[i] /= big;
  }
  return 1 - front * cur;
}

// standard normal deviate for lower tail probability p (AS 70)
double normal_quantile(double p) {
  const double a0 = -.322232431088, a1 = -1, a2 = -.342242088547, a3 = -.0204231210245,
               a4 = -.453642210148e-4, b0 = .0993484626060, b1 = .588581570495,
               b2 = .531103462366, b3 = .103537752850, b4 = .0038560700634;
  const double tail = (p < 0.5 ? p : 1 - p);
  if (tail < 1e-20) return -9999;
  const double y = std::sqrt(std::log(1 / (tail * tail)));
  const double z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) /
                           ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return (p < 0.5 ? -z : z);
}

// chi-square quantile with v degrees of freedom (AS 91)
double chi2_quantile(double p, double v) {
  const double e = .5e-6, aa = .6931471805;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  const double g = log_gamma(v / 2);
  const double xx = v / 2, c = xx - 1;
  double       ch, a = 0, q = 0, p1 = 0, p2 = 0, t = 0, b = 0;
  if (v < -1.24 * std::log(p)) {
    ch = std::pow((p * xx * std::exp(g + xx * aa)), 1 / xx);
    if (ch - e < 0) return ch;
  } else if (v <= .32) {
    ch = 0.4;
    a = std::log(1 - p);
    do {
      q = ch;
      p1 = 1 + ch * (4.67 + ch);
      p2 = ch * (6.73 + ch * (6.66 + ch));
      t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - std::exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (std::fabs(q / ch - 1) - .01 > 0);
  } else {
    const double x = normal_quantile(p);
    p1 = 0.222222 / v;
    ch = v * std::pow((x * std::sqrt(p1) + 1 - p1), 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * std::log(1 - p) - c * std::log(.5 * ch) + g;
  }
  do {
    q = ch;
    p1 = .5 * ch;
    if ((t = reg_inc_gamma(p1, xx, g)) < 0) return -1;
    p2 = p - t;
    t = p2 * std::exp(xx * aa + g + p1 - c * std::log(ch));
    b = t / ch;
    a = 0.5 * t - b * c;
    const double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    const double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    const double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    const double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    const double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    const double s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (std::fabs(q / ch - 1) > e);
  return ch;
}

inline double gamma_quantile(double p, double shape, double rate) {
  return chi2_quantile(p, 2.0 * shape) / (2.0 * rate);
}

}  // namespace

extern "C" int rdk_compute_gamma_cats(double alpha, unsigned int categories, double *output_rates,
                                      int rates_mode) {
  if (alpha < 0.02) {
    set_err("Invalid alpha value (must be >= 0.02)");
    return RDK_FAILURE;
  }
  if (categories == 0) {
    set_err("Number of categories must be positive");
    return RDK_FAILURE;
  }
  if (categories == 1) {
    output_rates[0] = 1.0;
    return RDK_SUCCESS;
  }
  const double shape = alpha, rate = alpha;
  const double factor = shape / rate * categories;
  if (rates_mode == RDK_GAMMA_RATES_MEDIAN) {
    const double half_step = 1.0 / (2.0 * categories);
    double       total = 0.0;
    for (unsigned i = 0; i < categories; ++i)
      output_rates[i] = gamma_quantile((double)(i * 2 + 1) * half_step, shape, rate);
    for (unsigned i = 0; i < categories; ++i) total += output_rates[i];
    for (unsigned i = 0; i < categories; ++i) output_rates[i] /= (total / (double)categories);
    return RDK_SUCCESS;
  }
  if (rates_mode == RDK_GAMMA_RATES_MEAN) {
    std::vector<double> cut(categories);
    const double        lg1 = log_gamma(shape + 1);
    for (unsigned i = 0; i + 1 < categories; ++i)
      cut[i] = gamma_quantile((i + 1.0) / categories, shape, rate);
    for (unsigned i = 0; i + 1 < categories; ++i) cut[i] = reg_inc_gamma(cut[i] * rate, shape + 1, lg1);
    output_rates[0] = cut[0] * factor;
    output_rates[categories - 1] = (1 - cut[categories - 2]) * factor;
    for (unsigned i = 1; i + 1 < categories; ++i) output_rates[i] = (cut[i] - cut[i - 1]) * factor;
    return RDK_SUCCESS;
  }
  set_err("Unknown gamma rates mode");
  return RDK_FAILURE;
}
