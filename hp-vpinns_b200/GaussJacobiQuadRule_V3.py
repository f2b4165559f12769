"""Quadrature set-up on the CPU (stays on the host per the north star): same names, arguments and return
conventions as the reference's Utilities/GaussJacobiQuadRule_V3.py (GJQ:24-61), so the reference drivers' line
``from GaussJacobiQuadRule_V3 import Jacobi, DJacobi, GaussLobattoJacobiWeights`` keeps working when this
directory is on sys.path.  Written independently on scipy's recurrence-based evaluators
(`eval_jacobi`, `roots_jacobi`) rather than on poly1d objects; pinned against tables produced by the
reference module in tests/golden/tables.npz.  Also hosts the test-function tables the engine consumes."""
import numpy as np
from scipy.special import eval_jacobi, gammaln, roots_jacobi


def Jacobi(n, a, b, x):
    """P_n^{(a,b)}(x) (GJQ:24-26)."""
    return eval_jacobi(n, a, b, np.asarray(x, dtype=np.float64))


def DJacobi(n, a, b, x, k):
    """k-th derivative of P_n^{(a,b)}: Gamma(a+b+n+1+k) / (2^k Gamma(a+b+n+1)) P_{n-k}^{(a+k,b+k)} (GJQ:30-33)."""
    x = np.asarray(x, dtype=np.float64)
    if k > n:
        return np.zeros_like(x)
    scale = np.exp(gammaln(a + b + n + 1 + k) - gammaln(a + b + n + 1)) / 2.0 ** k
    return scale * eval_jacobi(n - k, a + k, b + k, x)


def GaussJacobiWeights(Q, a, b):
    """Gauss-Jacobi nodes and weights (GJQ:38-40)."""
    X, W = roots_jacobi(Q, a, b)
    return [X, W]


def GaussLobattoJacobiWeights(Q, a, b):
    """Gauss-Lobatto-Jacobi rule with Q points including both end points (GJQ:46-61)."""
    X = roots_jacobi(Q - 2, a + 1, b + 1)[0]
    if a == 0 and b == 0:
        c = 2.0 / ((Q - 1) * Q)
        W = c / Jacobi(Q - 1, 0, 0, X) ** 2
        Wl = c / Jacobi(Q - 1, 0, 0, -1.0) ** 2
        Wr = c / Jacobi(Q - 1, 0, 0, 1.0) ** 2
    else:
        c = np.exp((a + b + 1) * np.log(2.0) + gammaln(a + Q) + gammaln(b + Q) - np.log(Q - 1) - gammaln(Q)
                   - gammaln(a + b + Q + 1))
        W = c / Jacobi(Q - 1, a, b, X) ** 2
        Wl = (b + 1) * c / Jacobi(Q - 1, a, b, -1.0) ** 2
        Wr = (a + 1) * c / Jacobi(Q - 1, a, b, 1.0) ** 2
    return [np.concatenate(([-1.0], X, [1.0])), np.concatenate(([Wl], W, [Wr]))]


def Test_fcn(N_test, x):
    """phi_n = P_{n+1} - P_{n-1}, n = 1..N (P1D:157-162, P2D:196-208, ADI:257-262).  Shape (N,) + x.shape."""
    x = np.asarray(x, dtype=np.float64)
    return np.asarray([Jacobi(n + 1, 0, 0, x) - Jacobi(n - 1, 0, 0, x) for n in range(1, N_test + 1)])


def dTest_fcn(N_test, x):
    """First and second derivative of phi_n w.r.t. the reference coordinate (P1D:164-183, P2D:210-229,
    ADI:265-284):  phi'_n = (n+2)/2 P_n^{(1,1)} - n/2 P_{n-2}^{(1,1)},
                   phi''_n = (n+2)(n+3)/4 P_{n-1}^{(2,2)} - n(n+1)/4 P_{n-3}^{(2,2)},
    a term with negative degree being absent."""
    x = np.asarray(x, dtype=np.float64)
    d1, d2 = [], []
    for n in range(1, N_test + 1):
        a = 0.5 * (n + 2) * Jacobi(n, 1, 1, x)
        if n >= 2:
            a = a - 0.5 * n * Jacobi(n - 2, 1, 1, x)
        b = 0.25 * (n + 2) * (n + 3) * Jacobi(n - 1, 2, 2, x)
        if n >= 3:
            b = b - 0.25 * n * (n + 1) * Jacobi(n - 3, 2, 2, x)
        d1.append(a)
        d2.append(b)
    return np.asarray(d1), np.asarray(d2)
