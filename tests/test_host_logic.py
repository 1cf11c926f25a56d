"""CPU-only tests of the host side: argument validation (order, types and messages of the
reference's checks), method / impl dispatch, the C ABI's exported symbols, and that the product
path fails loudly -- never falls back to a CPU implementation -- when no GPU is present."""

import ctypes
import os
import re

import numpy as np
import pytest

import cholupdates_b200
from cholupdates_b200 import _lib
from cholupdates_b200.rank_1 import _seeger

rank_1 = cholupdates_b200.rank_1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = [rank_1.update, rank_1.downdate]


def test_public_surface_matches_reference():
    # rank_1/__init__.py:39-49, _update.py:185, _seeger.py:299
    for name in ("update", "downdate", "update_seeger", "downdate_seeger",
                 "update_in_place", "downdate_in_place", "update_batched", "downdate_batched"):
        assert callable(getattr(rank_1, name))
    assert "seeger" in rank_1.update.available_methods
    assert "b200" in rank_1.downdate.available_methods
    assert rank_1.update_seeger.available_impls == ["b200"]


def test_c_abi_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "cholupdates_b200.h")).read()
    names = sorted(set(re.findall(r"\b(chub_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.chub_version() >= 100
    assert lib.chub_workspace_bytes(1, 16384) > 0
    assert isinstance(lib.chub_last_error(), bytes)


@pytest.mark.parametrize("fn", OPS)
@pytest.mark.parametrize("shape", [(3, 2), (3,), (1, 3, 3)])
def test_raise_on_invalid_cholesky_factor_shape(fn, shape):
    # test_update.py:88-97
    with pytest.raises(ValueError, match="not a square matrix"):
        fn(np.ones(shape), np.ones(shape[-1]))


@pytest.mark.parametrize("fn", OPS)
@pytest.mark.parametrize("shape", [(3, 2), (3, 1), (1, 3, 3)])
def test_raise_on_invalid_vector_shape(fn, shape):
    # test_update.py:100-107
    with pytest.raises(ValueError, match="incompatible"):
        fn(np.eye(3), np.ones(shape))


@pytest.mark.parametrize("fn", OPS)
def test_raise_on_vector_dimension_mismatch(fn):
    # test_update.py:110-126
    for n in (1, 2, 3, 5):
        with pytest.raises(ValueError):
            fn(np.eye(n), np.ones(n + 1))
        with pytest.raises(ValueError):
            fn(np.eye(n + 1), np.ones(n))


@pytest.mark.parametrize("fn", OPS)
def test_raise_on_zero_diagonal_numpy(fn):
    # _arg_validation.py:18-22; raised BEFORE the v-shape check and before any device work
    L = np.eye(4)
    L[2, 2] = 0.0
    with pytest.raises(np.linalg.LinAlgError, match="zeros on its diagonal"):
        fn(L, np.ones(4))
    with pytest.raises(np.linalg.LinAlgError):
        fn(L, np.ones(5))  # zero diagonal wins over the shape mismatch, as in the reference


@pytest.mark.parametrize("fn", OPS)
@pytest.mark.parametrize("L_dtype", [np.double, np.single, np.half, np.cdouble, np.intc])
@pytest.mark.parametrize("v_dtype", [np.double, np.single, np.half, np.cdouble, np.intc])
def test_raise_on_wrong_dtype(fn, L_dtype, v_dtype):
    # test_update_seeger.py:40-57
    if L_dtype == v_dtype and L_dtype in (np.double, np.single):
        return
    with pytest.raises(TypeError):
        fn(np.eye(3, dtype=L_dtype), np.ones(3, dtype=v_dtype))


@pytest.mark.parametrize("fn", OPS)
def test_raise_on_non_contiguous_inputs(fn):
    # test_update_seeger.py:22-37
    L = np.eye(6)[::2, ::2] + 0.0
    big = np.eye(6)
    with pytest.raises(ValueError, match="neither Fortran- nor C-contiguous"):
        fn(big[::2, ::2], np.ones(3))
    with pytest.raises(ValueError, match="not a contiguous array"):
        fn(L, np.ones(6)[::2])


@pytest.mark.parametrize("fn", OPS)
def test_unknown_method_and_impl(fn):
    # test_update.py:169-172, test_update_seeger.py:60-63
    with pytest.raises(NotImplementedError, match="Unknown method"):
        fn(np.eye(3), np.ones(3), method="nonexistent")
    with pytest.raises(NotImplementedError):
        fn(np.eye(3), np.ones(3), method="seeger", impl="nonexistent")
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        fn(np.eye(3), np.ones(3), method="seeger", impl="cython")


def test_product_path_never_imports_the_oracle():
    import sys

    pkg_dir = os.path.join(ROOT, "cholupdates_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "seeger_oracle" not in src and "oracle/" not in src and "build_ref" not in src, f
    assert "seeger_oracle" not in {m.split(".")[-1] for m in sys.modules if m.startswith("cholupdates_b200")}


def test_fails_loudly_without_a_gpu():
    lib = _lib.load()
    if lib.chub_device_count() > 0:
        pytest.skip("a CUDA device is present")
    for fn in OPS:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(np.eye(3), np.ones(3))
    st = ctypes.c_int32(0)
    L, v = np.eye(3), np.ones(3)
    rc = lib.chub_update_host_f64(L.ctypes.data, 3, 3, 0, v.ctypes.data, 0, ctypes.byref(st))
    assert rc != 0  # a CUDA error code, not a silently computed result
    np.testing.assert_array_equal(L, np.eye(3))


def test_layout_detection():
    assert _seeger._layout_of(np.zeros((4, 4), order="C")) == _lib.ROW_MAJOR
    assert _seeger._layout_of(np.zeros((4, 4), order="F")) == _lib.COL_MAJOR
    assert _seeger._layout_of(np.zeros((8, 8))[::2, ::2]) is None
    assert _seeger._layout_of(np.zeros((1, 1))) in (_lib.ROW_MAJOR, _lib.COL_MAJOR)


def test_rowcyclic_argument_checks_need_no_gpu():
    """Shape / dtype / contiguity checks of the distributed entry points fire before any device work."""
    import torch

    from cholupdates_b200.distributed import RowCyclicFactor

    class _NoBackend:
        pass

    n = 300
    A = torch.eye(n, dtype=torch.float64)
    fac = RowCyclicFactor(A, n, backend=_NoBackend())
    with pytest.raises(ValueError):
        RowCyclicFactor(torch.eye(n)[:100], n, backend=_NoBackend())  # wrong number of local rows
    with pytest.raises(ValueError):
        fac.update(torch.ones(n + 1, dtype=torch.float64))
    with pytest.raises(TypeError):
        fac.update(torch.ones(n, dtype=torch.float32))
    with pytest.raises(TypeError):
        fac.downdate(torch.ones(n, dtype=torch.float32))
    with pytest.raises(ValueError):
        fac.downdate(torch.ones(2 * n, dtype=torch.float64)[::2])
    with pytest.raises(ValueError):
        fac.update_many(torch.ones((2, n + 1), dtype=torch.float64))
    with pytest.raises(TypeError):
        fac.update_many(torch.ones((2, n), dtype=torch.float32))
    with pytest.raises(ValueError):
        fac.update_many(torch.ones((n, 2), dtype=torch.float64).t())
    assert fac.update_many(torch.ones((0, n), dtype=torch.float64)).shape == (0, n)


def test_wavefront_sizing_entry_points_need_no_gpu():
    """Slot / workspace / mailbox sizing of the wavefront mode is host arithmetic: consistent with the
    Python driver's own task bookkeeping (max tasks of one rank on an anti-diagonal)."""
    import ctypes

    lib = _lib.load()
    i64, ci = ctypes.c_int64, ctypes.c_int
    lib.chub_rcw_max_tasks.restype, lib.chub_rcw_max_tasks.argtypes = ci, [i64, i64, ci]
    lib.chub_rcw_workspace_bytes.restype, lib.chub_rcw_workspace_bytes.argtypes = ctypes.c_size_t, [i64, i64, ci]
    lib.chub_rcw_mailbox_bytes.restype, lib.chub_rcw_mailbox_bytes.argtypes = ctypes.c_size_t, [i64, i64, ci]
    pay = lib.chub_rcw_payload_doubles()
    assert pay == lib.chub_rcw_max_fuse() * (5 * 256 + 8)
    for n in (1, 255, 256, 257, 4096, 65536):
        n_panels = (n + 255) // 256
        for K in (1, 2, 7, 64, 1000):
            for G in (1, 2, 3, 8):
                # brute force: the most panels J = r (mod G) inside any window of min(K, panels) consecutive J
                L = min(K, n_panels)
                want = max(1, max(sum(1 for J in range(lo, lo + L) if J % G == r)
                                  for lo in range(0, n_panels - L + 1) for r in range(G)))
                got = lib.chub_rcw_max_tasks(n, K, G)
                assert got >= want and got == max(1, -(-L // G)), (n, K, G, got, want)
                n_pad = (n + 31) // 32 * 32
                assert lib.chub_rcw_workspace_bytes(n, K, G) >= 8 * (K * n_pad + K * 8)
                assert lib.chub_rcw_mailbox_bytes(n, K, G) >= 8 * 4 * G * got * pay



def test_strided_views_layout_and_leading_dimension():
    """`strided=True` (superset of the reference's contiguity check, _seeger.py:260-264): views with a unit
    inner stride map onto the C ABI's (layout, ld); everything else keeps raising ValueError."""
    from cholupdates_b200.rank_1 import _seeger

    big = np.zeros((12, 16))
    assert _seeger._layout_and_ld(big[:8, :8], False) is None
    assert _seeger._layout_and_ld(big[:8, :8], True) == (0, 16)
    assert _seeger._layout_and_ld(big.T[:8, :8], True) == (1, 16)
    assert _seeger._layout_and_ld(big[:8:2, :4], True) == (0, 32)
    assert _seeger._layout_and_ld(big[:8, :16:2], True) is None  # inner stride 2: not addressable
    assert _seeger._layout_and_ld(np.zeros((5, 5), order="F"), False) == (1, 5)
    err = _seeger._checks_after_diag(big[:8, :8], np.zeros(8), strided=False)
    assert isinstance(err, ValueError) and "neither Fortran- nor C-contiguous" in str(err)
    assert _seeger._checks_after_diag(big[:8, :8], np.zeros(8), strided=True) is None
