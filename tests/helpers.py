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
