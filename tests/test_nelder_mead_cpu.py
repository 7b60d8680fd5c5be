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
