"""Host optimisation driver: mirror of the reference's Nelder_mead.minimize
(lib/nelder_mead.ml:53-132), the caller that turns one likelihood evaluation into a workload
(BASELINE.json config 5).  Derivative-free simplex with alpha = 1, gamma = 2, rho = sigma = 0.5;
stops when the sample standard deviation of the simplex values (Gsl.Stats.sd) drops below `tol`
or after `maxit` iterations.  `f` is an opaque float-array -> float; with the GPU evaluator it is
"-log-likelihood(parameters)": upload O(branches) scalars, K1 + K2 + K3, read one double back.
"""
import math


def centroid(xs):
    """lib/nelder_mead.ml:40-47"""
    if len(xs) == 0:
        raise ValueError("Nelder_mead.centroid: empty array")
    n, d = len(xs), len(xs[0])
    out = []
    for i in range(d):
        acc = 0.0
        for x in xs:
            acc = acc + x[i]
        out.append(acc / float(n))
    return out


def update(c, alpha, towards):
    """lib/nelder_mead.ml:49-51: c + alpha (x - c)."""
    return [ci + alpha * (xi - ci) for ci, xi in zip(c, towards)]


def _sd(values):
    """Gsl.Stats.sd: sample standard deviation (n - 1 denominator)."""
    n = len(values)
    mean = 0.0
    for i, v in enumerate(values):            # GSL's running mean recurrence
        mean += (v - mean) / (i + 1)
    var = 0.0
    for i, v in enumerate(values):
        delta = v - mean
        var += (delta * delta - var) / (i + 1)
    return math.sqrt(var * (n / (n - 1.0)))


def _array_order(xs):
    """Utils.array_order ~compare:Float.compare (lib/utils.ml:78-81): ranks, ties by index."""
    return [i for _, i in sorted((x, i) for i, x in enumerate(xs))]


def _centroid_np(points):
    """centroid() on a [k][d] numpy array, coordinate sums accumulated in the same left-to-right order
    (np.cumsum is sequential), so the result equals the list version to the bit."""
    import numpy as np
    return np.cumsum(points, axis=0)[-1] / float(len(points))


def minimize_np(f, sample, tol=1e-8, maxit=100_000):
    """minimize() with the simplex held in one numpy array: the same algorithm, the same operation order in
    every coordinate (tests/test_nelder_mead_cpu.py compares the two to the bit) — for problems with hundreds of
    parameters (BASELINE config 5: 263), where the O(d^2) centroid per iteration is host time per evaluation."""
    import numpy as np
    alpha, gamma, rho, sigma = 1.0, 2.0, 0.5, 0.5
    x0 = np.asarray(sample(), dtype=np.float64)
    n = len(x0)
    if n == 0:
        raise ValueError("Nelder_mead.minimize: sample returns empty vectors")
    pts = np.empty((n + 1, n))
    for k in range(n + 1):
        y = np.asarray(sample(), dtype=np.float64)
        if len(y) != n:
            raise ValueError("Nelder_mead.minimize: sample returns vectors of varying lengths")
        pts[k] = y
    obj = [f(pts[k]) for k in range(n + 1)]
    upd = lambda c, a, x: c + a * (x - c)
    i = 0
    while True:
        ranks = _array_order(obj)
        best, second_worst, worst = ranks[0], ranks[n - 1], ranks[n]
        c = _centroid_np(pts[ranks[:n]])
        x_r = upd(c, -alpha, pts[worst])
        f_r = f(x_r)
        if not f_r < obj[best] and f_r < obj[second_worst]:
            pts[worst], obj[worst] = x_r, f_r
        elif f_r < obj[best]:
            x_e = upd(c, gamma, x_r)
            f_e = f(x_e)
            pts[worst] = x_e if f_e < f_r else x_r
            obj[worst] = min(f_r, f_e)
        else:
            if f_r < obj[worst]:
                x_c = upd(c, rho, x_r)
                f_c = f(x_c)
                accepted = f_c <= f_r
            else:
                x_c = upd(c, rho, pts[worst])
                f_c = f(x_c)
                accepted = f_c < obj[worst]
            if accepted:
                pts[worst], obj[worst] = x_c, f_c
            else:
                for k in range(n + 1):
                    if k != best:
                        pts[k] = upd(pts[best], sigma, pts[k])
                        obj[k] = f(pts[k])
        s = _sd(obj)
        if s < tol or i >= maxit:
            return obj[best], pts[best].copy(), i
        i += 1


def minimize(f, sample, tol=1e-8, maxit=100_000, debug=False):
    """Nelder_mead.minimize ?tol ?maxit ?debug ~f ~sample () -> (best value, best point, iterations)."""
    alpha, gamma, rho, sigma = 1.0, 2.0, 0.5, 0.5
    x0 = list(sample())
    n = len(x0)
    if n == 0:
        raise ValueError("Nelder_mead.minimize: sample returns empty vectors")

    def draw():
        y = list(sample())
        if len(y) != n:
            raise ValueError("Nelder_mead.minimize: sample returns vectors of varying lengths")
        return y
    points = [draw() for _ in range(n + 1)]
    obj = [f(p) for p in points]
    i = 0
    while True:
        ranks = _array_order(obj)
        best, second_worst, worst = ranks[0], ranks[n - 1], ranks[n]
        if debug:
            print("\n\nIteration %d: %f\nDelta: %g" % (i, obj[best], obj[worst] - obj[best]))
        c = centroid([points[r] for r in ranks[:n]])
        x_r = update(c, -alpha, points[worst])
        f_r = f(x_r)
        if debug:
            print("Candidate: %f" % f_r)
        if not f_r < obj[best] and f_r < obj[second_worst]:
            if debug:
                print("Reflection")
            points[worst], obj[worst] = x_r, f_r
        elif f_r < obj[best]:
            if debug:
                print("Expansion")
            x_e = update(c, gamma, x_r)
            f_e = f(x_e)
            points[worst] = x_e if f_e < f_r else x_r
            obj[worst] = min(f_r, f_e)
        else:
            if f_r < obj[worst]:                                   # outside contraction
                x_c = update(c, rho, x_r)
                f_c = f(x_c)
                accepted = f_c <= f_r
            else:                                                   # inside contraction
                x_c = update(c, rho, points[worst])
                f_c = f(x_c)
                accepted = f_c < obj[worst]
            if accepted:
                if debug:
                    print("Contraction, f_c = %f" % f_c)
                points[worst], obj[worst] = x_c, f_c
            else:
                if debug:
                    print("Shrink")
                for k in range(n + 1):
                    if k != best:
                        points[k] = update(points[best], sigma, points[k])
                        obj[k] = f(points[k])
        s = _sd(obj)
        if debug:
            print("Sigma: %f" % s)
        if s < tol or i >= maxit:
            return obj[best], points[best], i
        i += 1
