"""Shared helpers for the test-suite (importable as ``helpers`` from any test)."""

import numpy as np


def unpack_tril(packed: np.ndarray, N: int, order: str = "C", fill=0.0) -> np.ndarray:
    """Rebuild an (N, N) matrix from a packed lower triangle (see make_golden.py)."""
    M = np.full((N, N), fill, dtype=packed.dtype, order=order)
    M[np.tril_indices(N)] = packed
    return M


def rel_fro(X: np.ndarray, Y: np.ndarray) -> float:
    """Relative Frobenius error over the lower triangle."""
    X = np.tril(np.asarray(X, dtype=np.float64))
    Y = np.tril(np.asarray(Y, dtype=np.float64))
    d = np.linalg.norm(Y)
    return float(np.linalg.norm(X - Y) / (d if d > 0 else 1.0))


def synth_factor(n: int, seed, dtype=np.float64) -> np.ndarray:
    """Directly synthesised Cholesky factor: any lower-triangular matrix
    with a positive diagonal is a valid factor."""
    g = np.random.default_rng(seed)
    L = np.tril(g.standard_normal((n, n)) / np.sqrt(n), -1)
    L[np.diag_indices(n)] = 1.0 + g.random(n)
    return L.astype(dtype)


def spd_factor(n: int, seed, dtype=np.float64):
    """(A, L): random SPD matrix with spectrum 1 + Gamma(10, 1) and a Haar-ish eigenbasis, like the
    reference's fixtures (tests/cholupdates/rank_1/conftest.py:27-51, utils.py:9-53)."""
    g = np.random.default_rng(seed)
    spectrum = np.sort(1.0 + g.gamma(10.0, 1.0, size=n))
    Q, R = np.linalg.qr(g.standard_normal((n, n)))
    Q = Q * np.sign(np.diag(R))
    A = (Q * spectrum) @ Q.T
    A = 0.5 * (A + A.T)
    L = np.linalg.cholesky(A)
    return A.astype(dtype), L.astype(dtype)


def downdate_vector(L: np.ndarray, seed, norm_p: float = None) -> np.ndarray:
    """v = L p with ||p||^2 in [0.2, 0.9] (utils.py:104-143) or a prescribed ||p||."""
    g = np.random.default_rng(seed)
    n = L.shape[0]
    u = g.standard_normal(n)
    u /= np.linalg.norm(u)
    if norm_p is None:
        norm_p = np.sqrt(g.uniform(0.2, 0.9))
    return (np.tril(L).astype(np.float64) @ (norm_p * u)).astype(L.dtype)


def rel_vec(x: np.ndarray, y: np.ndarray) -> float:
    """Relative 2-norm error of a vector against the oracle's."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    d = np.linalg.norm(y)
    return float(np.linalg.norm(x - y) / (d if d > 0 else 1.0))


def config1_inputs(n: int):
    """(L, v) of BASELINE.json configs[0]: the reference's
    ``random_spd_matrix(n, fast=True, rng=default_rng(0))`` (src/cholupdates/utils.py:78-90, restated
    because the reference package does not travel to the GPU box; pinned against the live generator by
    tests/test_oracle.py) + ``scipy.linalg.cholesky(lower=True)``; ``v ~ N(0, 10^2)`` from default_rng(1)."""
    import scipy.linalg

    g = np.random.default_rng(0)
    A = g.standard_normal((n, n))
    M = A @ A.T
    M += np.eye(n)
    D = np.sqrt(np.diag(M))
    M = D[:, None] * M * D[None, :]
    M = 0.5 * (M + M.T)
    L = scipy.linalg.cholesky(M, lower=True)
    v = np.random.default_rng(1).normal(scale=10.0, size=n)
    return L, v


def tol_for(dtype, n: int) -> float:
    """Parity tolerance (relative Frobenius error over the lower triangle): the north star's
    1e-12 * sqrt(n) for fp64; 1e-5 * sqrt(n) for fp32."""
    return (1e-12 if np.dtype(dtype) == np.float64 else 1e-5) * np.sqrt(max(n, 1))
