"""Resident-factor handle: successive updates / downdates of ONE host factor without a PCIe round
trip of the matrix per call.

The reference's callers hold the factor in host memory and call ``update(L, v, overwrite_L=True)``
in a loop (GP / Kalman / quasi-Newton); the reference's own per-call O(N^2) cost next to the kernel
is the copy at ``src/cholupdates/rank_1/_seeger.py:272-276``.  On the GPU path the analogous cost is
2 x 8 N^2 / 2 bytes over PCIe per call (26 ms at N = 16384 against a 0.45 ms kernel).  This handle
uploads the lower triangle once, runs every update on the device copy (only ``v`` travels:
8 N bytes each way), and downloads the triangle when the caller asks for it::

    with ResidentFactor(L) as fac:          # L: numpy (N, N), C- or F-contiguous, float32 / float64
        for v in vectors:
            fac.update(v)                   # same checks / exceptions as rank_1.update
        fac.download(out=L)                 # lower triangle back into L (upper triangle untouched)

The device copy keeps the memory order of ``L``; the kernels never touch its strict upper triangle,
which is therefore neither uploaded nor downloaded.  Everything goes through the same C ABI entry
points as ``rank_1.update`` on CUDA tensors (``chub_update_*`` / ``chub_downdate_*`` on device
pointers); there is no CPU fallback.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .rank_1 import _seeger

_BLOCK = 1024  # rows (C order) / columns (F order) per transfer rectangle


class ResidentFactor:
    def __init__(self, L: np.ndarray, device=None):
        import torch

        _lib.require_device()
        if not isinstance(L, np.ndarray) or L.ndim != 2 or L.shape[0] != L.shape[1]:
            raise ValueError(
                f"The given Cholesky factor `L` is not a square matrix (given shape: "
                f"{getattr(L, 'shape', None)}).")
        if L.dtype not in (np.single, np.double):
            raise TypeError(
                f"The given Cholesky factor `L` does not have dtype `np.single` or `np.double` "
                f"(given dtype: {L.dtype.name})")
        if not (L.flags.c_contiguous or L.flags.f_contiguous):
            raise ValueError(
                "`L` is neither Fortran- nor C-contiguous, i.e. the memory layout of the matrix is "
                "neither column- nor row-major.")
        self.n = int(L.shape[0])
        self.dtype = L.dtype
        self.order = "F" if (L.flags.f_contiguous and not L.flags.c_contiguous) else "C"
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tdt = torch.float64 if L.dtype == np.double else torch.float32
        # row-major storage of L (C order) or of L^T (F order): the same bytes as the host array
        self._store = torch.empty((self.n, self.n), dtype=tdt, device=self.device)
        self._L = self._store if self.order == "C" else self._store.t()
        self._v = torch.empty(self.n, dtype=tdt, device=self.device)
        self._copy_triangle(L, to_device=True)

    # ---- transfers: only the lower triangle of L travels, as rectangles that end at the diagonal
    def _copy_triangle(self, L: np.ndarray, to_device: bool) -> None:
        import torch

        n = self.n
        host = L if self.order == "C" else L.T  # row-major view of the stored bytes
        for b0 in range(0, n, _BLOCK):
            b1 = min(n, b0 + _BLOCK)
            # C order: rows [b0, b1) x columns [0, b1).  F order (stored as L^T): rows of L^T are the
            # columns of L -> rows [b0, b1) x columns [b0, n) of the store hold L[b0:, b0:b1].
            sl = (slice(b0, b1), slice(0, b1)) if self.order == "C" else (slice(b0, b1), slice(b0, n))
            if to_device:
                self._store[sl].copy_(torch.from_numpy(host[sl]), non_blocking=False)
            else:
                blk = self._store[sl].cpu().numpy()
                if self.order == "C":
                    # the last rectangle row block contains entries above the diagonal: keep the host's
                    rows = np.arange(b0, b1)[:, None]
                    cols = np.arange(0, b1)[None, :]
                    mask = cols <= rows
                    host[sl][mask] = blk[mask]
                else:
                    rows = np.arange(b0, b1)[:, None]      # = column index j of L
                    cols = np.arange(b0, n)[None, :]        # = row index i of L
                    mask = cols >= rows
                    host[sl][mask] = blk[mask]

    # ---- the two operations (reference semantics, see rank_1._seeger)
    def _apply(self, op: str, v, check_diag: bool):
        import torch

        if not isinstance(v, np.ndarray) or v.ndim != 1 or v.shape[0] != self.n:
            raise ValueError(
                f"The shape of the given vector `v` is incompatible with the shape of the given Cholesky "
                f"factor `L`. Expected shape {(self.n,)} but got {getattr(v, 'shape', None)}.")
        if v.dtype != self.dtype:
            raise TypeError(f"`L` and `v` don't have the same dtype ({self.dtype.name} != {v.dtype.name}).")
        with torch.cuda.device(self.device):
            self._v.copy_(torch.from_numpy(np.ascontiguousarray(v)))
            _seeger._run(op, self._L, self._v, bool(check_diag))  # raises LinAlgError like rank_1.*
            return self._v.cpu().numpy()

    def update(self, v: np.ndarray, check_diag: bool = True) -> np.ndarray:
        """In-place rank-1 update of the resident factor.  Returns the scratch vector (what the
        reference leaves in ``v``)."""
        return self._apply("update", v, check_diag)

    def downdate(self, v: np.ndarray, check_diag: bool = True) -> np.ndarray:
        """In-place rank-1 downdate; raises ``numpy.linalg.LinAlgError`` (factor untouched) if the
        result is not positive definite."""
        return self._apply("downdate", v, check_diag)

    def download(self, out: np.ndarray = None) -> np.ndarray:
        """The lower triangle of the resident factor -> ``out`` (same shape, dtype and order as the
        uploaded array; its strict upper triangle is left alone).  ``out=None``: a new array with a
        zero upper triangle."""
        import torch

        if out is None:
            out = np.zeros((self.n, self.n), dtype=self.dtype, order=self.order)
        if out.shape != (self.n, self.n) or out.dtype != self.dtype:
            raise ValueError("`out` must have the shape and dtype of the uploaded factor.")
        out_order = "F" if (out.flags.f_contiguous and not out.flags.c_contiguous) else "C"
        if out_order != self.order or not (out.flags.c_contiguous or out.flags.f_contiguous):
            raise ValueError("`out` must have the memory order of the uploaded factor.")
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            self._copy_triangle(out, to_device=False)
        return out

    def device_tensor(self):
        """The device copy as a torch tensor in the order of the host array (zero-copy view); its
        strict upper triangle is uninitialised."""
        return self._L

    def close(self) -> None:
        self._store = self._L = self._v = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
