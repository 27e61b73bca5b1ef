// optim_capi.cpp -- C face of the optimiser components in optim.hpp, so that they can be driven
// alone (tests/test_optim.py: analytic functions through ctypes callbacks).  Return 1 on success,
// 0 on failure with the message in rdh_last_error().
#include "optim.hpp"

#include <string>

void rdh_set_error(const std::string &s);

// value and slope of the caller's function at x
extern "C" typedef void (*rdh_slope_fn)(double x, double *value, double *slope, void *user);
// the caller's objective at x[0..n)
extern "C" typedef double (*rdh_objective_fn)(const double *x, int n, void *user);

// root of the slope between lo and hi (slopes of opposite sign there); out = {x, value, slope},
// *probes = evaluations spent after the two end points
extern "C" int rdh_optim_slope_root(rdh_slope_fn fn, void *user, double lo, double hi, double x_tolerance,
                                    double *out, unsigned *probes) {
  try {
    unsigned spent = 0;
    auto     sample = [&](double x) {
      rd::slope_sample_t s;
      s.x = x;
      fn(x, &s.value, &s.slope, user);
      return s;
    };
    rd::brent_options_t opt;
    opt.x_tolerance = x_tolerance;
    const auto a = sample(lo), b = sample(hi);
    const auto r = rd::slope_root_brent(a, b, opt, [&](double x) {
      ++spent;
      return sample(x);
    });
    out[0] = r.x;
    out[1] = r.value;
    out[2] = r.slope;
    if (probes) *probes = spent;
    return 1;
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return 0;
  }
}

// the caller's function at xs[0..n) -> out[0..n): one batch
extern "C" typedef void (*rdh_batch_fn)(const double *xs, int n, double *out, void *user);

// rd::unit_segment_search_t::argmax on [0, 1] (model_t::optimize_alpha without the tree); a NaN
// value that a decision consumes fails with "lh at root is not a number".  *batches = calls of fn.
extern "C" int rdh_optim_argmax_on_segment(rdh_batch_fn fn, void *user, double x_now, double atol,
                                           int look_ahead, double *best_x, unsigned *batches) {
  try {
    unsigned calls = 0;
    auto     search = rd::make_unit_segment_search(
        [&](const std::vector<double> &xs) {
          ++calls;
          std::vector<double> out(xs.size(), 0.0);
          fn(xs.data(), (int)xs.size(), out.data(), user);
          return out;
        },
        [](double v) {
          if (std::isnan(v)) throw std::runtime_error("lh at root is not a number: " + std::to_string(v));
        },
        look_ahead != 0);
    *best_x = search.argmax(x_now, atol, "test segment");
    if (batches) *batches = calls;
    return 1;
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return 0;
  }
}

// the forward-difference slope of rd::unit_segment_search_t at x; out = {value, slope}
extern "C" int rdh_optim_slope_on_segment(rdh_batch_fn fn, void *user, double x, double *out) {
  try {
    auto search = rd::make_unit_segment_search(
        [&](const std::vector<double> &xs) {
          std::vector<double> f(xs.size(), 0.0);
          fn(xs.data(), (int)xs.size(), f.data(), user);
          return f;
        },
        [](double v) {
          if (std::isnan(v)) throw std::runtime_error("lh at root is not a number: " + std::to_string(v));
        },
        true);
    const auto s = search.slope_at(x);
    out[0] = s.value;
    out[1] = s.slope;
    return 1;
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return 0;
  }
}

// box-constrained minimisation from x[0..n) (updated in place as minimize_in_box defines it);
// *f_end = objective at the last point, *evaluations = objective calls
extern "C" int rdh_optim_minimize_in_box(rdh_objective_fn fn, void *user, double *x, int n, double lower,
                                         double upper, double pgtol, double factr, double *f_end,
                                         unsigned *evaluations) {
  try {
    unsigned                    calls = 0;
    std::vector<double>         v(x, x + n);
    rd::box_minimizer_options_t opt;
    opt.lower = lower;
    opt.upper = upper;
    opt.pgtol = pgtol;
    opt.factr = factr;
    const double f = rd::minimize_in_box(v, opt, [&](const std::vector<double> &pt) {
      ++calls;
      return fn(pt.data(), (int)pt.size(), user);
    });
    for (int i = 0; i < n; ++i) x[i] = v[(size_t)i];
    if (f_end) *f_end = f;
    if (evaluations) *evaluations = calls;
    return 1;
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return 0;
  }
}
