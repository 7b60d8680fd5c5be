"""Nelder_mead.minimize mirror: the reference's inline tests (lib/nelder_mead.ml:134-156)."""
import random

import pytest

from phylogenetics_b200 import nelder_mead as NM


def test_parabola():
    rng = random.Random(1)
    obj, _, _ = NM.minimize(lambda x: x[0] ** 2.0, lambda: [rng.uniform(-100.0, 100.0)], tol=1e-3)
    assert abs(obj) < 1e-3


def test_rosenbrock():
    rng = random.Random(2)
    f = lambda x: 100.0 * (x[1] - x[0] ** 2.0) ** 2.0 + (1.0 - x[0]) ** 2.0
    obj, x, _ = NM.minimize(f, lambda: [rng.uniform(-100.0, 100.0) for _ in range(2)])
    assert abs(obj) < 1e-3


def test_powell_quartic():
    rng = random.Random(3)
    f = lambda x: (x[0] + 10.0 * x[1]) ** 2.0 + 5.0 * (x[2] - x[3]) ** 2.0 + (x[1] - 2.0 * x[2]) ** 4.0 \
        + 10.0 * (x[0] - x[3]) ** 4.0
    obj, _, _ = NM.minimize(f, lambda: [rng.uniform(-100.0, 100.0) for _ in range(4)])
    assert abs(obj) < 1e-3


def test_argument_errors_and_helpers():
    with pytest.raises(ValueError):
        NM.minimize(lambda x: 0.0, lambda: [])
    with pytest.raises(ValueError):
        NM.centroid([])
    lens = iter([2, 2, 3, 2])
    with pytest.raises(ValueError):
        NM.minimize(lambda x: 0.0, lambda: [0.0] * next(lens))
    assert NM.centroid([[0.0, 2.0], [2.0, 4.0]]) == [1.0, 3.0]
    assert NM.update([1.0, 1.0], -1.0, [2.0, 3.0]) == [0.0, -1.0]
    assert abs(NM._sd([1.0, 2.0, 3.0, 4.0]) - 1.2909944487358056) < 1e-15
    assert NM._array_order([3.0, 1.0, 2.0, 1.0]) == [1, 3, 2, 0]
    _, _, its = NM.minimize(lambda x: x[0] ** 2, lambda: [5.0], maxit=3, tol=0.0)
    assert its == 3


def test_cpp_mirror_matches_the_python_mirror_run_by_run():
    """phylogenetics_b200/host/phylo_b200.hpp, namespace Nelder_mead: the reference's three inline tests from
    starting points drawn from one 32-bit LCG — the C++ program and this test replay the same stream, so the two
    mirrors must agree on the value reached and on the number of iterations, not only on `< 1e-3`."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "_build", "nelder_mead_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off",
                    "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "phylogenetics_b200", "host"),
                    os.path.join(root, "tests", "cpp", "nelder_mead_test.cpp"), "-o", exe,
                    "-L" + os.path.join(root, "phylogenetics_b200"), "-lphylo_b200",
                    "-Wl,-rpath,$ORIGIN/../../../phylogenetics_b200"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "all ok" in out.stdout and "FAIL" not in out.stdout, out.stdout + out.stderr
    got = {l.split()[1]: (float(l.split()[2]), int(l.split()[3])) for l in out.stdout.splitlines() if l.startswith("RESULT")}

    state = [12345]

    def nxt():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xFFFFFFFF
        return state[0] / 4294967296.0 * 200.0 - 100.0

    def draw(n):
        return lambda: [nxt() for _ in range(n)]

    def rosen(x):
        a, b = x[1] - x[0] * x[0], 1.0 - x[0]
        return 100.0 * (a * a) + b * b

    def powell(x):
        a, b, c, d = x[0] + 10.0 * x[1], x[2] - x[3], x[1] - 2.0 * x[2], x[0] - x[3]
        return a * a + 5.0 * (b * b) + (c * c) * (c * c) + 10.0 * ((d * d) * (d * d))

    want = {"parabola": NM.minimize(lambda x: x[0] * x[0], draw(1), tol=1e-3),
            "rosenbrock": NM.minimize(rosen, draw(2)),
            "powell": NM.minimize(powell, draw(4))}
    for name, (obj, _, its) in want.items():
        assert got[name][1] == its, (name, got[name], obj, its)
        assert got[name][0] == obj, (name, got[name], obj)


def test_numpy_simplex_is_the_same_algorithm_to_the_bit():
    """minimize_np (simplex in one array, for the 263-parameter problem of BASELINE config 5) against the list
    mirror of lib/nelder_mead.ml:53-132: same value, same point, same iteration count."""
    import numpy as np
    from phylogenetics_b200 import nelder_mead as nm

    def rosen(x):
        x = np.asarray(x)
        return float(np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
    for d in (2, 7, 23):
        r1, r2 = np.random.default_rng(d), np.random.default_rng(d)
        a = nm.minimize(rosen, lambda: list(r1.uniform(-2, 2, d)), tol=1e-10, maxit=2500)
        b = nm.minimize_np(rosen, lambda: r2.uniform(-2, 2, d), tol=1e-10, maxit=2500)
        assert a[0] == b[0] and a[2] == b[2] and np.array_equal(np.asarray(a[1]), b[1])
