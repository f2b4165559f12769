"""TEST INFRASTRUCTURE ONLY -- stand-in for ``pyDOE.lhs`` (absent from this image; the reference calls it at
P1D:311, P2D:314-347, ADI:358-388,466-474).  Restates pyDOE 0.3.8's default ("classic") Latin-hypercube:
one uniform draw per stratum and an independent permutation of the strata per column, both taken from the
numpy GLOBAL RNG exactly as pyDOE does, so ``np.random.seed(1234)`` in the scripts governs it."""
import numpy as np


def lhs(n, samples=None, criterion=None, iterations=None):
    if samples is None:
        samples = n
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a = cut[:samples]
    b = cut[1:samples + 1]
    rdpoints = np.zeros_like(u)
    for j in range(n):
        rdpoints[:, j] = u[:, j] * (b - a) + a
    H = np.zeros_like(rdpoints)
    for j in range(n):
        order = np.random.permutation(range(samples))
        H[:, j] = rdpoints[order, j]
    return H
