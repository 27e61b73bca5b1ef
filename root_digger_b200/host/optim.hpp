// optim.hpp -- the two numerical optimisers model_t drives, as stand-alone components that know
// nothing about trees or partitions:
//
//   rd::slope_root_brent   root of a derivative on a bracket (Brent's method on the slope).  A
//                          "probe" is whatever the caller uses to evaluate (value, slope) at x.
//   rd::unit_segment_search_t
//                          the best position on a segment by the sign of the slope, on batched
//                          evaluations: model_t::optimize_alpha and compute_dlh without the tree
//                          (for model_t a batch is one fused root-only evaluation call on the GPU).
//   rd::lbfgsb_session_t   RAII face of the reverse-communication L-BFGS-B 3.0 routine (setulb),
//   rd::minimize_in_box    and a box-constrained minimiser with forward-difference gradients on
//                          top of it, the engine of model_t::optimize_params.
//
// Results contract.  RootDigger's answer (chosen branch, position on it, LWR ranking) is the end
// point of an optimiser TRAJECTORY, so these follow the decision rules of the reference step for
// step -- the same comparisons on the same floating-point expressions -- and the deviations from
// the textbook algorithms are kept and named (Q1..Q4 below).  tests/test_reference_sources.py
// holds the reference's own src/model.cpp (brents :606-676, bfgs_params :1430-1522) against this
// code bit for bit; tests/test_optim.py exercises the components alone on analytic functions.
#ifndef RD_HOST_OPTIM_HPP_
#define RD_HOST_OPTIM_HPP_

#include "lbfgsb_driver.hpp"

#include <cmath>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rd {

// ---------------------------------------------------------------------------------------------
// Brent on the slope
// ---------------------------------------------------------------------------------------------
// one evaluated abscissa: the function value and its derivative there
struct slope_sample_t {
  double x = 0.0;
  double value = 0.0;
  double slope = 0.0;
};

struct brent_options_t {
  double   x_tolerance = 1e-12;      // absolute tolerance on x (the reference's atol)
  double   slope_floor = 1e-12;      // |slope| at or below this counts as a root
  double   coincident = 1e-12;       // |a - c| below this: secant instead of inverse quadratic
  unsigned max_iterations = 64;
};

// Finds x in [lo.x, hi.x] where the slope changes sign, given samples at both ends whose slopes
// have opposite signs.  probe(x) -> slope_sample_t evaluates a new abscissa.  Returns the sample
// the iteration settled on (value included, so the caller needs no further evaluation).
//
//   Q1  The rank test that orders the current estimate `b` against the contra-point `c` is the
//       reference's: the two are exchanged when |slope(b)| < |slope(c)|, i.e. `b` ends up the
//       endpoint with the LARGER residual.  (Textbook Brent exchanges on the opposite test.)
//       The iterates, and with them the root position reported to the last bit, depend on it.
//   Q2  64 iterations, then failure -- never a silent best effort.
template <typename Probe>
slope_sample_t slope_root_brent(slope_sample_t lo, slope_sample_t hi, const brent_options_t &opt,
                                Probe &&probe) {
  if (!(lo.slope * hi.slope < 0)) throw std::runtime_error("Brents called with endpoints which don't bracket");
  constexpr double eps = std::numeric_limits<double>::epsilon();
  slope_sample_t   a = lo, b = hi, c = hi;
  double           step = b.x - a.x, prev_step = step;

  for (unsigned it = 0; it < opt.max_iterations; ++it) {
    if (b.slope * c.slope > 0.0) {  // root no longer between b and c: fall back to [a, b]
      c = a;
      step = prev_step = b.x - a.x;
    }
    if (std::fabs(b.slope) < std::fabs(c.slope)) {  // Q1
      a = b;
      b = c;
      c = a;
    }
    const double tol = 2.0 * std::fabs(b.x) * eps + 0.5 * opt.x_tolerance;
    const double half = 0.5 * (c.x - b.x);
    if (std::fabs(half) <= tol || std::fabs(b.slope) <= opt.slope_floor) return b;

    bool interpolated = false;
    if (std::fabs(prev_step) >= tol && std::fabs(a.slope) > std::fabs(b.slope)) {
      const double s = b.slope / a.slope;
      double       p, q;
      if (std::fabs(a.x - c.x) < opt.coincident) {  // two distinct points: secant
        p = 2.0 * half * s;
        q = 1.0 - s;
      } else {  // three: inverse quadratic interpolation
        const double qa = a.slope / c.slope, r = b.slope / c.slope;
        p = s * (2.0 * half * qa * (qa - r) - (b.x - a.x) * (r - 1.0));
        q = (qa - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      p = std::fabs(p);
      const double bound1 = 3.0 * half * q - std::fabs(half * q), bound2 = std::fabs(prev_step * q);
      if (2.0 * p < (bound1 < bound2 ? bound1 : bound2)) {  // accepted: inside the bracket, shrinking fast enough
        prev_step = step;
        step = p / q;
        interpolated = true;
      }
    }
    if (!interpolated) step = prev_step = half;  // bisection

    a = b;
    const double x_next = std::fabs(step) > tol ? b.x + step : b.x + (half >= 0.0 ? tol : -tol);
    b = probe(x_next);
  }
  throw std::runtime_error("Brents method failed to converge");  // Q2
}

// ---------------------------------------------------------------------------------------------
// the best position on a segment
// ---------------------------------------------------------------------------------------------
// Maximum of a function on [0, 1] located by the sign of its slope -- model_t::optimize_alpha and
// compute_dlh (reference src/model.cpp:679-794, 481-519) without the tree: `evaluate(xs)` returns
// the function at every abscissa of a batch (for model_t one fused engine call), `refuse(value)`
// throws for a value the caller cannot work with (NaN) and is applied to the values a decision
// actually CONSUMES, in consumption order -- a batch may hold evaluations made ahead of need.
//
//   slope      forward difference of step 1e-8, taken backwards where x + step would reach 1
//              (Appendix B-15); two evaluations that are both infinite give a flat slope
//   argmax     1. the value at the current position (only checked) and value + slope at 0 and 1
//                 -- one batch of five evaluations
//              2. an end whose slope is flat (|slope| < atol) ends the search: the better END wins;
//                 slopes of opposite sign bracket a stationary point -> Brent on the slope, and
//                 the better of that point and the better end wins
//              3. same sign at both ends: scan the dyadic grid 1/2; 1/4, 3/4; 1/8 ... 31/32 level
//                 by level for a slope of the other sign (one batch per level when looking ahead:
//                 the later points of a level are evaluated speculatively), refine on both sides
//                 of the first such point and take the best of the two refinements, the better
//                 end and the best flat grid point seen so far; without a turn: the best flat
//                 grid point, else the end the slope points to
template <typename Evaluate, typename Refuse>
class unit_segment_search_t {
public:
  struct raw_slope_t {
    double x, fx, fxh, sign;
  };

  unit_segment_search_t(Evaluate evaluate, Refuse refuse, bool look_ahead, double step = 1e-8)
      : _evaluate(std::move(evaluate)), _refuse(std::move(refuse)), _look_ahead(look_ahead), _step(step) {}

  // values at `values_at`, then the evaluation pairs of the slopes at `slopes_at`: ONE batch
  std::vector<raw_slope_t> probe(const std::vector<double> &values_at, const std::vector<double> &slopes_at,
                                 std::vector<double> *values = nullptr) {
    std::vector<double>      xs(values_at);
    std::vector<raw_slope_t> raw;
    raw.reserve(slopes_at.size());
    for (double x : slopes_at) {
      raw_slope_t r{x, 0.0, 0.0, 1.0};
      double      partner = x + _step;
      if (partner >= 1.0) {
        partner = x - _step;
        r.sign = -1.0;
      }
      xs.push_back(x);
      xs.push_back(partner);
      raw.push_back(r);
    }
    const std::vector<double> f = _evaluate(xs);
    if (f.size() != xs.size()) throw std::logic_error("unit_segment_search_t: evaluate() returned a batch of another size");
    if (values) values->assign(f.begin(), f.begin() + (std::ptrdiff_t)values_at.size());
    for (size_t i = 0; i < raw.size(); ++i) {
      raw[i].fx = f[values_at.size() + 2 * i];
      raw[i].fxh = f[values_at.size() + 2 * i + 1];
    }
    return raw;
  }

  slope_sample_t settle(const raw_slope_t &raw) const {
    _refuse(raw.fx);
    _refuse(raw.fxh);
    if (std::isinf(raw.fxh) && std::isinf(raw.fx)) return {raw.x, raw.fx, 0.0};
    const double slope = (raw.fxh - raw.fx) / _step;
    return {raw.x, raw.fx, slope * raw.sign};
  }

  slope_sample_t slope_at(double x) { return settle(probe({}, {x})[0]); }

  slope_sample_t refine_between(const slope_sample_t &lo, const slope_sample_t &hi, double atol) {
    brent_options_t opt;
    opt.x_tolerance = atol;
    return slope_root_brent(lo, hi, opt, [this](double x) { return slope_at(x); });
  }

  // `context` is appended to the message of the one failure that names the segment
  double argmax(double x_now, double atol, const std::string &context = std::string()) {
    slope_sample_t lo, hi;
    if (_look_ahead) {
      std::vector<double> now;
      const auto          ends = probe({x_now}, {0.0, 1.0}, &now);
      _refuse(now[0]);
      lo = settle(ends[0]);
      hi = settle(ends[1]);
    } else {
      _refuse(_evaluate(std::vector<double>{x_now})[0]);
      lo = slope_at(0.0);
      hi = slope_at(1.0);
    }
    if (std::isnan(lo.slope) || std::isnan(hi.slope))
      throw std::runtime_error("Initial derivatives failed when optimizing alpha: " + context);

    slope_sample_t best_end = lo.value >= hi.value ? lo : hi;
    if (std::fabs(lo.slope) < atol || std::fabs(hi.slope) < atol) return best_end.x;

    if ((lo.slope < 0.0 && hi.slope > 0.0) || (lo.slope > 0.0 && hi.slope < 0.0)) {
      const auto inner = refine_between(lo, hi, atol);
      return best_end.value > inner.value ? best_end.x : inner.x;
    }

    const bool     rising = lo.slope > 0.0 && hi.slope > 0.0;
    slope_sample_t best_flat{0.0, -std::numeric_limits<double>::infinity(), 0.0};
    bool           have_flat = false;
    for (size_t cells = 2; cells <= 32; cells *= 2) {
      std::vector<double> grid;
      for (size_t k = 1; k <= cells; k += 2) grid.push_back(1.0 / (double)cells * k);
      std::vector<raw_slope_t> level;
      if (_look_ahead) level = probe({}, grid);
      for (size_t j = 0; j < grid.size(); ++j) {
        const auto here = _look_ahead ? settle(level[j]) : slope_at(grid[j]);
        if (std::fabs(here.slope) < atol && best_flat.value < here.value) {
          best_flat = here;
          have_flat = true;
        }
        const bool turns = rising ? here.slope < 0.0 : here.slope > 0.0;
        if (!turns) continue;
        const auto left = refine_between(lo, here, atol);
        const auto right = refine_between(here, hi, atol);
        if (best_end.value < best_flat.value) best_end = best_flat;
        const auto &inner = left.value < right.value ? right : left;
        return best_end.value >= inner.value ? best_end.x : inner.x;
      }
    }
    if (have_flat) return best_flat.x;
    return rising ? 1.0 : 0.0;
  }

private:
  Evaluate _evaluate;
  Refuse   _refuse;
  bool     _look_ahead;
  double   _step;
};

template <typename Evaluate, typename Refuse>
unit_segment_search_t<Evaluate, Refuse> make_unit_segment_search(Evaluate evaluate, Refuse refuse, bool look_ahead) {
  return unit_segment_search_t<Evaluate, Refuse>(std::move(evaluate), std::move(refuse), look_ahead);
}

// ---------------------------------------------------------------------------------------------
// L-BFGS-B
// ---------------------------------------------------------------------------------------------
// One run of setulb with its workspace.  Reverse communication: advance() hands back what the
// routine wants next; the caller stores f (and g) through the references before calling again.
class lbfgsb_session_t {
public:
  enum class request_t { evaluate, new_iterate, finished };

  lbfgsb_session_t(std::vector<double> x0, double lower, double upper, int memory, double factr, double pgtol)
      : _n((int)x0.size()), _m(memory), _factr(factr), _pgtol(pgtol), _x(std::move(x0)),
        _lower((size_t)_n, lower), _upper((size_t)_n, upper), _bound_kind((size_t)_n, 2 /* both bounds */),
        _g((size_t)_n, 0.0),
        _wa((2 * (size_t)memory + 5) * (size_t)_n + 12 * (size_t)memory * ((size_t)memory + 1), 0.0),
        _iwa(3 * (size_t)_n, 0), _setulb(load_setulb()) {}

  std::vector<double> &x() { return _x; }
  std::vector<double> &gradient() { return _g; }
  double              &f() { return _f; }

  request_t advance() {
    _setulb(&_n, &_m, _x.data(), _lower.data(), _upper.data(), _bound_kind.data(), &_f, _g.data(), &_factr,
            &_pgtol, _wa.data(), _iwa.data(), &_task, &_iprint, &_csave, _lsave, _isave, _dsave);
    if (lbfgsb_is_fg(_task)) return request_t::evaluate;
    return _task == LBFGSB_NEW_X ? request_t::new_iterate : request_t::finished;
  }

private:
  int                 _n, _m;
  double              _factr, _pgtol, _f = 0.0;
  std::vector<double> _x, _lower, _upper;
  std::vector<int>    _bound_kind;
  std::vector<double> _g, _wa;
  std::vector<int>    _iwa;
  setulb_fn           _setulb;
  int                 _task = LBFGSB_START, _iprint = -1, _csave = 0;
  int                 _lsave[4] = {0, 0, 0, 0}, _isave[44] = {0};
  double              _dsave[29] = {0};
};

struct box_minimizer_options_t {
  double lower = 0.0, upper = 1.0;  // the same box for every coordinate
  double fd_step = 1e-4;            // forward difference: h = max(fd_step * |x_i|, fd_step)
  double pgtol = 1e-7, factr = 1e4;
  int    memory = 20;
  size_t max_rounds = 500;
};

// Minimises objective(x) over the box, starting from (and reporting through) `x`.
// objective(const std::vector<double>&) -> double installs the point and evaluates it; it is the
// only way this function touches the outside world.  Returns the objective at the last point.
//
//   Q3  The objective is evaluated after EVERY return of setulb, whatever it asked for, and once
//       more after the loop; the caller's side effects (the parameters left installed in the
//       partition) follow from that sequence.
//   Q4  `x` receives the final point when it is not worse than the start (f_start >= f_end),
//       otherwise it keeps the start -- but the LAST point stays installed either way.
template <typename Objective>
double minimize_in_box(std::vector<double> &x, const box_minimizer_options_t &opt, Objective &&objective) {
  lbfgsb_session_t run(x, opt.lower, opt.upper, opt.memory, opt.factr, opt.pgtol);
  const double     f_start = objective(x);
  run.f() = f_start;
  auto &pt = run.x();

  for (size_t round = 0; round < opt.max_rounds; ++round) {
    const auto want = run.advance();
    run.f() = objective(pt);  // Q3
    if (want == lbfgsb_session_t::request_t::finished) break;
    if (want == lbfgsb_session_t::request_t::new_iterate) continue;
    auto &g = run.gradient();
    for (size_t i = 0; i < pt.size(); ++i) {
      const double keep = pt[i];
      double       h = opt.fd_step * std::fabs(keep);
      if (h < opt.fd_step) h = opt.fd_step;
      pt[i] += h;
      const double f_h = objective(pt);
      if (!std::isfinite(f_h)) throw std::runtime_error("dlh is not finite");
      g[i] = (f_h - run.f()) / h;
      if (!std::isfinite(g[i])) throw std::runtime_error("gradient is not finite");
      pt[i] = keep;
    }
  }
  const double f_end = objective(pt);
  if (f_start >= f_end) x = pt;  // Q4
  return f_end;
}

}  // namespace rd
#endif
