"""TEST INFRASTRUCTURE -- CPU restatement of the reference's search for the root position on a
branch, with the likelihood replaced by any function f on [0, 1].

Restates, statement by statement and in the reference's evaluation order (one f call where the
reference makes one compute_lh_root call):

  compute_dlh      src/model.cpp:481-519   forward difference, step 1e-8, backwards at the upper end
  brents           src/model.cpp:606-676   Brent on the slope with the reference's own rank test
  optimize_alpha   src/model.cpp:679-794   ends, bracket, dyadic grid

Python floats are IEEE doubles and every expression below keeps the reference's association, so
the results are comparable BIT FOR BIT with root_digger_b200/host/optim.hpp
(rd::unit_segment_search_t, rd::slope_root_brent) driven with the same f -- which is how
tests/test_optim.py uses it: the product batches and restructures, this file does not.

Pinned against the reference itself: the reference's own src/model.cpp, compiled unchanged
(oracle/_ref), and the product agree bit for bit on real likelihood surfaces
(tests/test_reference_sources.py); this restatement extends that to surfaces chosen to reach the
branches real data rarely takes (same-sign ends with an interior turn, flat grid points, -inf
plateaus).  Only tests/ may import this module.
"""
import math
import sys

import numpy as np

EPSILON = 1e-8
DBL_EPSILON = sys.float_info.epsilon


def _div(a, b):
    """IEEE-754 division (C++ semantics: x / 0 is an infinity or NaN, never an exception)"""
    with np.errstate(all="ignore"):
        return float(np.float64(a) / np.float64(b))


class NotANumber(RuntimeError):
    """what compute_lh_root throws for a NaN log-likelihood (src/model.cpp:446-449)"""


class Trace:
    """f with the reference's NaN rule and a log of the abscissae evaluated, in order"""

    def __init__(self, f):
        self.f, self.xs = f, []

    def __call__(self, x):  # compute_lh_root(root at ratio x)
        self.xs.append(x)
        v = self.f(x)
        if math.isnan(v):
            raise NotANumber("lh at root is not a number: nan")
        return v


def compute_dlh(lh_root, x):
    """src/model.cpp:481-519 -> (lh, dlh)"""
    x_prime = x + EPSILON
    sign = 1.0
    if x_prime >= 1.0:
        x_prime = x - EPSILON
        sign = -1.0
    fx = lh_root(x)
    fxh = lh_root(x_prime)
    if math.isinf(fxh) and math.isinf(fx):
        return fx, 0.0
    dlh = (fxh - fx) / EPSILON
    return fx, dlh * sign


def brents(lh_root, beg, d_beg, end, d_end, atol):
    """src/model.cpp:606-676; beg / end are ratios, d_* = (lh, dlh) -> (ratio, lh)"""
    if not (d_beg[1] * d_end[1] < 0):
        raise RuntimeError("Brents called with endpoints which don't bracket")
    midpoint, d_midpoint = end, d_end
    d = e = end - beg
    for _ in range(64):
        if d_end[1] * d_midpoint[1] > 0.0:
            midpoint, d_midpoint = beg, d_beg
            d = e = end - beg
        if abs(d_end[1]) < abs(d_midpoint[1]):
            beg = end
            end = midpoint
            midpoint = beg
            d_beg = d_end
            d_end = d_midpoint
            d_midpoint = d_beg
        tol = 2.0 * abs(end) * DBL_EPSILON + 0.5 * atol
        e_tol = 0.5 * (midpoint - end)
        if abs(e_tol) <= tol or abs(d_end[1]) <= 1e-12:
            return end, d_end[0]
        if abs(e) >= tol and abs(d_beg[1]) > abs(d_end[1]):
            s = _div(d_end[1], d_beg[1])
            if abs(beg - midpoint) < 1e-12:
                p = 2.0 * e_tol * s
                q = 1.0 - s
            else:
                q = _div(d_beg[1], d_midpoint[1])
                r = _div(d_end[1], d_midpoint[1])
                p = s * (2.0 * e_tol * q * (q - r) - (end - beg) * (r - 1.0))
                q = (q - 1.0) * (r - 1.0) * (s - 1.0)
            if p > 0.0:
                q = -q
            p = abs(p)
            min1 = 3.0 * e_tol * q - abs(e_tol * q)
            min2 = abs(e * q)
            if 2.0 * p < (min1 if min1 < min2 else min2):
                e = d
                d = _div(p, q)
            else:
                d = e_tol
                e = d
        else:
            d = e_tol
            e = d
        beg = end
        d_beg = d_end
        if abs(d) > tol:
            end += d
        else:
            end += tol if e_tol >= 0.0 else -tol
        d_end = compute_dlh(lh_root, end)
    raise RuntimeError("Brents method failed to converge")


def optimize_alpha(f, x_now, atol):
    """src/model.cpp:679-794 -> (best ratio, [abscissae evaluated, in order])"""
    lh_root = f if isinstance(f, Trace) else Trace(f)
    lh_root(x_now)
    beg, end = 0.0, 1.0
    d_beg = compute_dlh(lh_root, beg)
    d_end = compute_dlh(lh_root, end)
    if math.isnan(d_beg[1]) or math.isnan(d_end[1]):
        raise RuntimeError("Initial derivatives failed when optimizing alpha")
    best_endpoint = beg if d_beg[0] >= d_end[0] else end
    lh_best_endpoint = d_beg if d_beg[0] >= d_end[0] else d_end
    if abs(d_beg[1]) < atol or abs(d_end[1]) < atol:
        return best_endpoint, lh_root.xs
    if (d_beg[1] < 0.0 and d_end[1] > 0.0) or (d_beg[1] > 0.0 and d_end[1] < 0.0):
        mid = brents(lh_root, beg, d_beg, end, d_end, atol)
        return (best_endpoint if lh_best_endpoint[0] > mid[1] else mid[0]), lh_root.xs

    beg_end_pos = d_beg[1] > 0.0 and d_end[1] > 0.0
    best_midpoint_lh = (-math.inf, 0.0)
    best_midpoint = None
    found_midpoint = False
    midpoints = 2
    while midpoints <= 32:
        for midpoint in range(1, midpoints + 1):
            if midpoint % 2 == 0:
                continue
            alpha = 1.0 / float(midpoints) * midpoint
            d_midpoint = compute_dlh(lh_root, alpha)
            if abs(d_midpoint[1]) < atol:
                if best_midpoint_lh[0] < d_midpoint[0]:
                    best_midpoint_lh = d_midpoint
                    best_midpoint = alpha
                    found_midpoint = True
            if (beg_end_pos and d_midpoint[1] < 0.0) or (not beg_end_pos and d_midpoint[1] > 0.0):
                r1 = brents(lh_root, beg, d_beg, alpha, d_midpoint, atol)
                r2 = brents(lh_root, alpha, d_midpoint, end, d_end, atol)
                if lh_best_endpoint[0] < best_midpoint_lh[0]:
                    lh_best_endpoint = best_midpoint_lh
                    best_endpoint = best_midpoint
                if r1[1] < r2[1]:
                    return (best_endpoint if lh_best_endpoint[0] >= r2[1] else r2[0]), lh_root.xs
                return (best_endpoint if lh_best_endpoint[0] >= r1[1] else r1[0]), lh_root.xs
        midpoints *= 2
    if found_midpoint:
        return best_midpoint, lh_root.xs
    return (end if beg_end_pos else beg), lh_root.xs
