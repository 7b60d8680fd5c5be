"""Oracle restatement of the reference's linear-algebra helpers used on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows lib/linear_algebra.ml of the reference.  Matrices are plain numpy
float64 arrays indexed [row, col]; the column-major Bigarray layout of the
reference is a storage detail that only matters at the C-ABI.
"""
import math

import numpy as np

# Pade coefficient tables, lib/linear_algebra.ml:321-330 and :357-363 (they are
# the standard Higham-2005 numerators, b_k = (2m-k)! m! / ((2m)! k! (m-k)!) scaled
# to integers).
_PADE = {
    3: [120., 60., 12., 1.],
    5: [30240., 15120., 3360., 420., 30., 1.],
    7: [17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.],
    9: [17643225600., 8821612800., 2075673600., 302702400., 30270240.,
        2162160., 110880., 3960., 90., 1.],
    13: [64764752532480000., 32382376266240000., 7771770303897600.,
         1187353796428800., 129060195264000., 10559470521600.,
         670442572800., 33522128640., 1323241920., 40840800., 960960.,
         16380., 182., 1.],
}


def norm1(x):
    """Matrix.norm1 = dlange 'O' (max column abs-sum), lib/linear_algebra.ml:211."""
    return float(np.max(np.sum(np.abs(x), axis=0)))


def expm(x):
    """Matrix.expm, lib/linear_algebra.ml:308-404.

    Pade 3/5/7/9 below the 1-norm thresholds of :320-330, otherwise Pade-13
    with scaling by 2^-ceil(log2(norm/5.4)) and repeated squaring (:351-401).
    No balancing (the reference has a TODO for it at :316).
    """
    x = np.asarray(x, dtype=np.float64)
    m, n = x.shape
    if m != n:
        raise ValueError("matrix not square")          # invalid_arg, :311
    if m == 1:
        return np.array([[math.exp(x[0, 0])]])           # :313-314
    ident = np.eye(m)
    nx = norm1(x)
    if nx <= 2.097847961257068:                           # :320
        if nx > 0.9504178996162932:
            c = _PADE[9]
        elif nx > 0.2539398330063230:
            c = _PADE[7]
        elif nx > 0.01495585217958292:
            c = _PADE[5]
        else:
            c = _PADE[3]
        x2 = x @ x
        p = ident.copy()
        u = c[1] * p
        v = c[0] * p
        for i in range(1, len(c) // 2):                   # :337-343
            p = p @ x2
            u = u + c[2 * i + 1] * p
            v = v + c[2 * i] * p
        u = x @ u
        return np.linalg.solve(v - u, v + u)              # gesv, :345-349
    s = math.log(nx / 5.4) / math.log(2.0)                # :352, log2 via log/log 2 (:305-306)
    t = math.ceil(s)
    if s > 0.0:
        x = x * (2.0 ** (-t))                             # :354
    c = _PADE[13]
    x2 = x @ x
    x4 = x2 @ x2
    x6 = x2 @ x4
    w = c[9] * x2 + c[11] * x4 + c[13] * x6               # :368-372
    w = x6 @ w
    w = w + c[1] * ident + c[3] * x2 + c[5] * x4 + c[7] * x6
    u = x @ w                                             # :378
    w = c[8] * x2 + c[10] * x4 + c[12] * x6               # :380-384
    w = x6 @ w
    v = w + c[0] * ident + c[2] * x2 + c[4] * x4 + c[6] * x6
    r = np.linalg.solve(v - u, v + u)                     # :392-394
    if s > 0.0:
        for _ in range(int(t)):                           # :397-400
            r = r @ r
    return r


def zero_eigen_vector(mat):
    """Matrix.zero_eigen_vector, lib/linear_algebra.ml:266-275.

    Least-squares solution (dgels) of [mat^T ; 1^T] x = [0 ; 1].
    """
    mat = np.asarray(mat, dtype=np.float64)
    n = mat.shape[1]
    if mat.shape[0] != n:
        raise ValueError("Expected square matrix")
    a = np.vstack([mat.T, np.ones((1, n))])
    b = np.zeros(n + 1)
    b[n] = 1.0
    sol, *_ = np.linalg.lstsq(a, b, rcond=None)
    return sol


def diagonalize(m):
    """Matrix.diagonalize = dsyevr on a symmetric matrix, lib/linear_algebra.ml:238-241.

    Returns (eigenvalues ascending, eigenvectors as columns), m = P diag(v) P^T.
    """
    v, p = np.linalg.eigh(np.asarray(m, dtype=np.float64))
    return v, p


def matrix_pow(x, k):
    """Matrix.pow, lib/linear_algebra.ml:277-291 (square-and-multiply)."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[0] != x.shape[1]:
        raise ValueError("non-square matrix")
    if k < 0:
        raise ValueError("negative power")
    if k == 0:
        return np.eye(x.shape[0])
    if k % 2 == 0:
        r = matrix_pow(x, k // 2)
        return r @ r
    r = matrix_pow(x, (k - 1) // 2)
    return x @ (r @ r)
