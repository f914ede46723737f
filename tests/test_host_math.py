"""CPU: host linear algebra of the C++ mirror (the reference calls LAPACK eigh: ED_DIAG.f90:194, ED_GF_NORMAL.f90:618)."""
import numpy as np
import pytest


@pytest.mark.parametrize("n", [1, 2, 3, 17, 100, 256])
def test_dense_eigh(edb, n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    A = A + A.T
    a = np.asfortranarray(A.copy())
    w = np.zeros(n)
    assert edb.lib().ed_host_eigh(n, a.ctypes.data_as(edb.dp), w.ctypes.data_as(edb.dp)) == 0
    assert np.allclose(w, np.linalg.eigvalsh(A), atol=1e-11, rtol=0)
    assert np.abs(A @ a - a * w).max() < 1e-11
    assert np.abs(a.T @ a - np.eye(n)).max() < 1e-12
    # the eigenvalue-only pre-pass of the LAPACK sectors: the same reduction and QL sweeps, no accumulation -> the same bits
    a2 = np.asfortranarray(A.copy())
    w2 = np.zeros(n)
    assert edb.lib().ed_host_eigvals(n, a2.ctypes.data_as(edb.dp), w2.ctypes.data_as(edb.dp)) == 0
    assert np.array_equal(a2, A) and np.array_equal(w2, w)


@pytest.mark.parametrize("n", [1, 2, 50, 200])
def test_tridiagonal_eigh_matches_oracle_tql2(edb, oracle, n):
    rng = np.random.default_rng(100 + n)
    d = rng.normal(size=n)
    e = np.zeros(n)
    e[1:] = rng.uniform(0.1, 1.0, size=n - 1)
    w = np.zeros(n)
    z = np.zeros((n, n), order="F")
    assert edb.lib().ed_host_eigh_tridiag(n, d.ctypes.data_as(edb.dp), e.ctypes.data_as(edb.dp),
                                          w.ctypes.data_as(edb.dp), z.ctypes.data_as(edb.dp)) == 0
    w2, z2 = oracle.tql2(d, e[1:])
    assert np.array_equal(w, w2)                       # same algorithm, same arithmetic
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    assert np.allclose(w, np.linalg.eigvalsh(T), atol=1e-12)
    assert np.allclose(np.abs(z[0]), np.abs(z2[0]), atol=1e-12)
