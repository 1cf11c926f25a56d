"""On-device versions of the reference's data generators (``src/cholupdates/utils.py:9-143``).

The reference uses them for its tests and its benchmark notebook; here they make large SPD test
inputs (e.g. for the block-row-cyclic n = 65536 workload) without a 35 GB host array.  They follow the
reference's constructions with torch's CUDA generator, so the DISTRIBUTIONS match, not the random
streams (``tests/helpers.config1_inputs`` restates the NumPy stream bit for bit where that matters).
Library calls (QR, triangular solve) are fine here: this is test-input plumbing, not the hot path.
"""

from __future__ import annotations

from typing import Optional, Tuple


def _gen(device, generator):
    import torch

    if generator is not None:
        return generator
    g = torch.Generator(device=device)
    g.seed()
    return g


def random_spd_eigendecomposition(N: int, dtype=None, generator=None, device="cuda") -> Tuple["torch.Tensor", "torch.Tensor"]:
    """(spectrum, basis): spectrum ``1 + Gamma(10, 1)`` sorted ascending, eigenbasis Haar-distributed
    (reference ``utils.py:9-53``)."""
    import torch

    dtype = dtype or torch.float64
    g = _gen(device, generator)
    # Gamma(10, 1) = sum of 10 Exp(1) draws
    u = torch.rand((10, N), dtype=torch.float64, device=device, generator=g).clamp_min(1e-300)
    spectrum = (1.0 - torch.log(u).sum(dim=0)).sort().values.to(dtype)
    if N == 1:
        return spectrum, torch.ones((1, 1), dtype=dtype, device=device)
    Z = torch.randn((N, N), dtype=torch.float64, device=device, generator=g)
    Q, R = torch.linalg.qr(Z)
    Q = Q * torch.sign(torch.diagonal(R))  # Haar measure (Mezzadri 2007)
    if torch.linalg.det(Q) < 0:            # special orthogonal group, like scipy.stats.special_ortho_group
        Q[:, 0] = -Q[:, 0]
    return spectrum, Q.to(dtype)


def random_spd_matrix(N: int, fast: bool = False, dtype=None, generator=None, device="cuda"):
    """Random symmetric positive-definite matrix (reference ``utils.py:56-101``): ``fast`` = ``A A^T + I``
    with the reference's diagonal scaling, else a random eigendecomposition."""
    import torch

    dtype = dtype or torch.float64
    g = _gen(device, generator)
    if fast:
        A = torch.randn((N, N), dtype=dtype, device=device, generator=g)
        M = A @ A.T
        M += torch.eye(N, dtype=dtype, device=device)
        D = torch.sqrt(torch.diagonal(M))
        M = D[:, None] * M * D[None, :]
    else:
        spectrum, Q = random_spd_eigendecomposition(N, dtype=dtype, generator=g, device=device)
        M = (Q * spectrum) @ Q.T
    return 0.5 * (M + M.T)


def random_rank_1_downdate(L, generator=None):
    """A vector ``v`` such that ``L L^T - v v^T`` stays positive definite: uniformly random direction,
    ``||L^{-1} v||^2`` uniform in [0.2, 0.9] (reference ``utils.py:104-143``, with ONE norm per vector)."""
    import torch

    N = L.shape[0]
    g = _gen(L.device, generator)
    v_dir = torch.randn(N, dtype=L.dtype, device=L.device, generator=g)
    v_dir /= torch.linalg.norm(v_dir)
    p_dir = torch.linalg.solve_triangular(torch.tril(L), v_dir[:, None], upper=False)[:, 0]
    norm_sq = 0.2 + 0.7 * torch.rand((), dtype=L.dtype, device=L.device, generator=g)
    return torch.sqrt(norm_sq / torch.dot(p_dir, p_dir)) * v_dir
