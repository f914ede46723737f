"""CPU ORACLE (test infrastructure, NOT product code) for the dmft-ed Lanczos hot path.

PARITY UNPINNED: the reference ships no golden vectors/tests for this path and cannot be built in this
image (no Fortran compiler; SciFortran/DMFT_Tools are un-vendored, un-pinned dependencies).  This module
restates the reference's algorithm on the CPU: the inner loops live in ed_oracle.c (literal C restatement),
the solver phases above them (ED_DIAG, ED_GF_NORMAL, ED_OBSERVABLES) are restated here with numpy.
Every function cites the reference file:line it follows.  It is pinned by independent invariants only
(tests/test_oracle_invariants.py) and by the probe anchors of BASELINE.md section 5.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/libed_oracle.so with gcc (oracle/Makefile)."""
    so = os.path.join(_HERE, "libed_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ed_oracle.c", "ed_oracle_mt.c", "ed_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libed_oracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    u64p = C.POINTER(C.c_uint64)
    i64p = C.POINTER(C.c_int64)
    L.ora_model_new.restype = C.c_void_p
    L.ora_model_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, dp, dp, dp]
    L.ora_model_free.argtypes = [C.c_void_p]
    L.ora_init_bath.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, dp]
    L.ora_binomial.restype = C.c_int64
    L.ora_binomial.argtypes = [C.c_int, C.c_int]
    L.ora_sector_dim.restype = C.c_int64
    L.ora_sector_dim.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ora_build_sector.restype = C.c_int64
    L.ora_build_sector.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_int]
    L.ora_binary_search.restype = C.c_int64
    L.ora_binary_search.argtypes = [u64p, C.c_int64, C.c_uint64]
    L.ora_c.argtypes = [C.c_int, C.c_uint64, u64p, dp]
    L.ora_cdg.argtypes = [C.c_int, C.c_uint64, u64p, dp]
    L.ora_direct_hxv.argtypes = [C.c_void_p, u64p, C.c_int64, dp, dp, C.c_int64, C.c_int64]
    L.ora_gather_hxv.argtypes = [C.c_void_p, u64p, C.c_int64, dp, dp, C.c_int64, C.c_int64]
    L.ora_stored_build.restype = C.c_int64
    L.ora_stored_build.argtypes = [C.c_void_p, u64p, C.c_int64, i64p, i64p, dp]
    L.ora_stored_hxv.argtypes = [C.c_int64, i64p, i64p, dp, dp, dp]
    L.ora_tql2.restype = C.c_int
    L.ora_tql2.argtypes = [C.c_int, dp, dp, dp]
    L.ora_lanc_tridiag.restype = C.c_int
    L.ora_lanc_tridiag.argtypes = [C.c_void_p, u64p, C.c_int64, dp, C.c_int, C.c_double, dp, dp]
    L.ora_lanc_gs.restype = C.c_int
    L.ora_lanc_gs.argtypes = [C.c_void_p, u64p, C.c_int64, dp, C.c_int, C.c_double, C.c_int, dp, dp, dp]
    L.ora_apply_op.restype = C.c_double
    L.ora_apply_op.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_int64, u64p, C.c_int64, dp, dp]
    L.ora_observables.argtypes = [C.c_int, C.c_int, u64p, C.c_int64, dp, C.c_double] + [dp] * 8
    L.ora_philox_normal.argtypes = [C.c_uint64, C.c_int64, C.c_int64, dp]
    L.ora_philox_uniform.argtypes = [C.c_uint64, C.c_int64, C.c_int64, dp]
    L.ora_map_entry.restype = C.c_uint64
    L.ora_map_entry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64]
    L.ora_window_hxv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, i64p, C.c_int64, dp]
    L.ora_num_threads.restype = C.c_int
    L.ora_gather_hxv_mt.restype = C.c_double
    L.ora_gather_hxv_mt.argtypes = [C.c_void_p, u64p, C.c_int64, dp, dp, C.c_int64, C.c_int64, C.c_int]
    L.ora_direct_hxv_timed.restype = C.c_double
    L.ora_direct_hxv_timed.argtypes = [C.c_void_p, u64p, C.c_int64, dp, dp, C.c_int64, C.c_int64]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


# ----------------------------------------------------------------------------------------------------------
# parameters (defaults = ED_INPUT_VARS.f90:121-196)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Params:
    Norb: int = 1
    Nbath: int = 6
    Nspin: int = 1
    uloc: tuple = (2.0, 0.0, 0.0)
    ust: float = 0.0
    jh: float = 0.0
    jx: float = 0.0
    jp: float = 0.0
    beta: float = 1000.0
    xmu: float = 0.0
    hfmode: bool = True
    Lmats: int = 5000
    Lreal: int = 5000
    wini: float = -5.0
    wfin: float = 5.0
    eps: float = 0.01
    gs_threshold: float = 1e-9
    hwband: float = 2.0
    lanc_method: str = "arpack"
    lanc_nstates_sector: int = 6
    lanc_nstates_total: int = 1
    lanc_ncv_factor: int = 3
    lanc_ncv_add: int = 5
    lanc_niter: int = 512
    lanc_ngfiter: int = 200
    lanc_tolerance: float = 1e-12
    lanc_dim_threshold: int = 256
    ed_twin: bool = False
    ed_sparse_H: bool = True
    chispin_flag: bool = False
    Ltau: int = 1000
    chidens_flag: bool = False

    @property
    def Ns(self):
        return (self.Nbath + 1) * self.Norb          # ED_SETUP.f90:99-101


def init_bath(p: Params) -> np.ndarray:
    """init_dmft_bath with noise 0 (ED_BATH/dmft_aux.f90:105-127) -> user bath vector [e..., v...]."""
    bath = np.zeros(2 * p.Nspin * p.Norb * p.Nbath)
    lib().ora_init_bath(p.Norb, p.Nbath, p.Nspin, p.hwband, _dp(bath))
    return bath


class Model:
    """Owns an ora_model (bath + Hloc + interaction parameters)."""

    def __init__(self, p: Params, bath: np.ndarray, hloc: np.ndarray | None = None):
        self.p = p
        self.bath = np.ascontiguousarray(bath, dtype=np.float64)
        n = p.Nspin * p.Nspin * p.Norb * p.Norb
        if hloc is None:
            hloc = np.zeros((p.Nspin, p.Nspin, p.Norb, p.Norb), dtype=np.complex128)
        self.hloc = np.asarray(hloc, dtype=np.complex128).reshape(p.Nspin, p.Nspin, p.Norb, p.Norb)
        flat = self.hloc.reshape(-1, order="F")       # Fortran column-major, as impHloc is stored
        self._hre = np.ascontiguousarray(flat.real)
        self._him = np.ascontiguousarray(flat.imag)
        assert self._hre.size == n
        ul = np.zeros(5)
        ul[: len(p.uloc)] = p.uloc
        self._ul = ul
        self.h = lib().ora_model_new(p.Norb, p.Nbath, p.Nspin, int(p.hfmode), _dp(ul), p.ust, p.jh, p.jx, p.jp,
                                     p.xmu, _dp(self.bath), _dp(self._hre), _dp(self._him))

    def __del__(self):
        try:
            if self.h:
                lib().ora_model_free(self.h)
                self.h = None
        except Exception:
            pass

    # bath accessors, 0-based (ispin, iorb, k)
    def e(self, ispin, iorb, k):
        p = self.p
        return self.bath[(ispin * p.Norb + iorb) * p.Nbath + k]

    def v(self, ispin, iorb, k):
        p = self.p
        return self.bath[p.Nspin * p.Norb * p.Nbath + (ispin * p.Norb + iorb) * p.Nbath + k]


# ----------------------------------------------------------------------------------------------------------
# thin wrappers over the C restatement
# ----------------------------------------------------------------------------------------------------------
def binomial(n, k):
    return int(lib().ora_binomial(n, k))


def sector_dim(Ns, nup, ndw):
    return int(lib().ora_sector_dim(Ns, nup, ndw))


def build_sector(Ns, nup, ndw, literal=False) -> np.ndarray:
    """ED_SETUP.f90:899-916 in 64-bit."""
    dim = sector_dim(Ns, nup, ndw)
    m = np.empty(dim, dtype=np.uint64)
    got = lib().ora_build_sector(Ns, nup, ndw, _u64p(m), int(literal))
    assert got == dim
    return m


def direct_hxv(model: Model, smap: np.ndarray, v: np.ndarray) -> np.ndarray:
    """directMatVec_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:21-92), complex in/out."""
    vin = np.ascontiguousarray(v, dtype=np.complex128)
    hv = np.zeros_like(vin)
    lib().ora_direct_hxv(model.h, _u64p(smap), smap.size, _dp(vin.view(np.float64)), _dp(hv.view(np.float64)),
                         0, smap.size)
    return hv


def gather_hxv(model: Model, smap: np.ndarray, v: np.ndarray, i0=0, i1=None) -> np.ndarray:
    vin = np.ascontiguousarray(v, dtype=np.complex128)
    hv = np.zeros_like(vin)
    i1 = smap.size if i1 is None else i1
    lib().ora_gather_hxv(model.h, _u64p(smap), smap.size, _dp(vin.view(np.float64)), _dp(hv.view(np.float64)),
                         i0, i1)
    return hv


def window_hxv(model: Model, nup: int, ndw: int, seed: int, rows) -> np.ndarray:
    """Re (H v)(i) for the reference indices `rows` with v = philox_uniform(seed): needs no map and no vector
    (ora_window_hxv), so it works at Ns=16/18 where the literal oracle does not fit the host."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.empty(rows.size)
    lib().ora_window_hxv(model.h, nup, ndw, seed, _i64p(rows), rows.size, _dp(out))
    return out


def map_entry(Ns, nup, ndw, i) -> int:
    return int(lib().ora_map_entry(Ns, nup, ndw, int(i)))


def stored_build(model: Model, smap: np.ndarray):
    """ed_buildH_c (ED_HAMILTONIAN_STORED_HxV.f90:28-113): (rowptr, cols, vals) in insertion order."""
    dim = smap.size
    rowptr = np.zeros(dim + 1, dtype=np.int64)
    nnz = lib().ora_stored_build(model.h, _u64p(smap), dim, _i64p(rowptr), None, None)
    cols = np.zeros(nnz, dtype=np.int64)
    vals = np.zeros(nnz, dtype=np.complex128)
    lib().ora_stored_build(model.h, _u64p(smap), dim, _i64p(rowptr), _i64p(cols), _dp(vals.view(np.float64)))
    return rowptr, cols, vals


def stored_hxv(rowptr, cols, vals, v):
    vin = np.ascontiguousarray(v, dtype=np.complex128)
    hv = np.zeros_like(vin)
    lib().ora_stored_hxv(rowptr.size - 1, _i64p(rowptr), _i64p(cols), _dp(vals.view(np.float64)),
                         _dp(vin.view(np.float64)), _dp(hv.view(np.float64)))
    return hv


def dense_h(model: Model, smap: np.ndarray) -> np.ndarray:
    """sp_dump_matrix of the stored form (ED_HAMILTONIAN_STORED_HxV.f90:100-110)."""
    rowptr, cols, vals = stored_build(model, smap)
    dim = smap.size
    H = np.zeros((dim, dim), dtype=np.complex128)
    for i in range(dim):
        sl = slice(rowptr[i], rowptr[i + 1])
        np.add.at(H[i], cols[sl], vals[sl])
    return H


def tql2(diag, sub):
    """tql2 (.repo/PLAIN_LANCZOS.f90:427-565): diag(1:n), sub(2:n) -> (eigenvalues, Z)."""
    n = len(diag)
    d = np.array(diag, dtype=np.float64)
    e = np.zeros(n)
    e[1:] = np.asarray(sub, dtype=np.float64)[: n - 1] if n > 1 else []
    z = np.eye(n, order="F")
    ierr = lib().ora_tql2(n, _dp(d), _dp(e), _dp(z))
    assert ierr == 0
    return d, z


def philox_normal(seed: int, n: int, i0: int = 0) -> np.ndarray:
    out = np.empty(n)
    lib().ora_philox_normal(seed, i0, n, _dp(out))
    return out


def philox_uniform(seed: int, n: int, i0: int = 0) -> np.ndarray:
    out = np.empty(n)
    lib().ora_philox_uniform(seed, i0, n, _dp(out))
    return out


def lanc_tridiag(model: Model, smap, vin, nitermax, threshold=1e-13):
    """sp_lanc_tridiag (ancestor .repo/PLAIN_LANCZOS.f90:154-180). Returns alanc[n], blanc[n] (blanc[0] unused
    = Fortran blanc(1)), nused.  vin is normalised in place like the reference."""
    v = np.ascontiguousarray(vin, dtype=np.complex128).copy()
    a = np.zeros(nitermax)
    b = np.zeros(nitermax)
    nused = lib().ora_lanc_tridiag(model.h, _u64p(smap), smap.size, _dp(v.view(np.float64)), nitermax, threshold,
                                   _dp(a), _dp(b))
    return a, b, nused


def lanc_gs(model: Model, smap, v0, nitermax, threshold, ncheck=10):
    """sp_lanc_eigh (ancestor .repo/PLAIN_LANCZOS.f90:286-385)."""
    v = np.ascontiguousarray(v0, dtype=np.complex128).copy()
    a = np.zeros(nitermax + 1)
    b = np.zeros(nitermax + 1)
    egs = C.c_double(0.0)
    nlanc = lib().ora_lanc_gs(model.h, _u64p(smap), smap.size, _dp(v.view(np.float64)), nitermax, threshold,
                              ncheck, C.byref(egs), _dp(a), _dp(b))
    return egs.value, v, nlanc, a[:nlanc], b[:nlanc]


def apply_op(Ns, isite, dagger, mapI, mapJ, gs):
    """Seed of a GF chain, ED_GF_NORMAL.f90:159-174 / 212-227 (isite 1-based level)."""
    g = np.ascontiguousarray(gs, dtype=np.complex128)
    vv = np.zeros(mapJ.size, dtype=np.complex128)
    n2 = lib().ora_apply_op(Ns, isite, int(dagger), _u64p(mapI), mapI.size, _u64p(mapJ), mapJ.size,
                            _dp(g.view(np.float64)), _dp(vv.view(np.float64)))
    return vv, float(n2)


# ----------------------------------------------------------------------------------------------------------
# solver phases (numpy restatement)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class State:
    e: float
    sector: int           # 1-based isector = nup*(Ns+1)+ndw+1   (ED_SETUP.f90:382-393)
    nup: int
    ndw: int
    vec: np.ndarray


@dataclass
class Result:
    states: list = field(default_factory=list)
    egs: float = 0.0
    zeta: float = 0.0
    eig_by_sector: dict = field(default_factory=dict)
    nlanc_by_sector: dict = field(default_factory=dict)
    chains: list = field(default_factory=list)     # (iorb, ispin, isign, istate, norm2, alfa, beta, nused)
    wm: np.ndarray | None = None
    wr: np.ndarray | None = None
    impGmats: np.ndarray | None = None
    impGreal: np.ndarray | None = None
    impSmats: np.ndarray | None = None
    impSreal: np.ndarray | None = None
    impG0mats: np.ndarray | None = None
    impG0real: np.ndarray | None = None
    dens: np.ndarray | None = None
    dens_up: np.ndarray | None = None
    dens_dw: np.ndarray | None = None
    docc: np.ndarray | None = None
    magz: np.ndarray | None = None
    sz2: np.ndarray | None = None
    n2: np.ndarray | None = None
    s2tot: float = 0.0
    vm: np.ndarray | None = None
    tau: np.ndarray | None = None
    spinChi_iv: np.ndarray | None = None
    spinChi_tau: np.ndarray | None = None
    spinChi_w: np.ndarray | None = None
    chi_chains: list = field(default_factory=list)
    densChi_iv: np.ndarray | None = None
    densChi_tau: np.ndarray | None = None
    densChi_w: np.ndarray | None = None
    densChi_tot_iv: np.ndarray | None = None
    densChi_tot_tau: np.ndarray | None = None
    densChi_tot_w: np.ndarray | None = None


def sector_index(Ns, nup, ndw):
    return nup * (Ns + 1) + ndw + 1                    # ED_SETUP.f90:382-387


def start_vector(dim: int, seed: int = 1234567) -> np.ndarray:
    """Deterministic Lanczos start vector shared by oracle and product: Philox uniforms in (-1,1), bit-identical
    on CPU and GPU (the reference draws random_number() uniforms, .repo/PLAIN_LANCZOS.f90:310-318; seed value
    from .repo/ARPACK_LANCZOS.f90:165)."""
    return philox_uniform(seed, dim).astype(np.complex128)


def ed_diag(model: Model, res: Result, sectors=None):
    """ed_diag_c (ED_DIAG.f90:49-251), T=0, ed_twin=F."""
    p = model.p
    Ns = p.Ns
    oldzero = 1000.0
    states: list[State] = []
    for nup in range(Ns + 1):
        for ndw in range(Ns + 1):
            isector = sector_index(Ns, nup, ndw)
            if sectors is not None and (nup, ndw) not in sectors:
                continue
            dim = sector_dim(Ns, nup, ndw)
            if p.lanc_method == "lanczos":                                   # :93-97
                neigen, nitermax = 1, min(dim, p.lanc_niter)
            else:                                                             # :88-92
                neigen = min(dim, min(dim, p.lanc_nstates_sector))
                nitermax = min(dim, p.lanc_niter)
            lanc_solve = True
            if neigen == dim:
                lanc_solve = False
            if dim <= max(p.lanc_dim_threshold, 1):
                lanc_solve = False
            smap = build_sector(Ns, nup, ndw)
            if lanc_solve:
                if p.lanc_method == "lanczos":
                    e0, vec, nl, _, _ = lanc_gs(model, smap, start_vector(dim), nitermax, p.lanc_tolerance)
                    res.nlanc_by_sector[(nup, ndw)] = nl
                    evals = np.array([e0])
                    evecs = vec[:, None]
                else:
                    # sp_eigh = ARPACK 'SA' (ED_DIAG.f90:149-166); scipy's eigsh wraps the same *saupd/*seupd
                    from scipy.sparse.linalg import LinearOperator, eigsh
                    op = LinearOperator((dim, dim), dtype=np.complex128,
                                        matvec=lambda x: direct_hxv(model, smap, x))
                    ncv = min(dim, p.lanc_ncv_factor * max(neigen, p.lanc_nstates_sector) + p.lanc_ncv_add)
                    evals, evecs = eigsh(op, k=neigen, which="SA", ncv=ncv, tol=p.lanc_tolerance,
                                         maxiter=max(nitermax, 10 * dim), v0=start_vector(dim).real)
                    order = np.argsort(evals)
                    evals, evecs = evals[order], evecs[:, order]
            else:
                H = dense_h(model, smap)                                      # :188-197
                w, Z = np.linalg.eigh(H)
                evals, evecs = w[:neigen], Z[:, :neigen]
            res.eig_by_sector[(nup, ndw)] = np.array(evals)
            for i in range(neigen):                                           # :224-235
                enemin = float(evals[i])
                if enemin < oldzero - 10.0 * p.gs_threshold:
                    oldzero = enemin
                    states = [State(enemin, isector, nup, ndw, np.array(evecs[:, i], dtype=np.complex128))]
                elif abs(enemin - oldzero) <= p.gs_threshold:
                    oldzero = min(oldzero, enemin)
                    _insert_state(states, State(enemin, isector, nup, ndw, np.array(evecs[:, i], dtype=np.complex128)))
    res.states = states
    res.egs = min(s.e for s in states)
    res.zeta = float(len(states))                                             # ED_DIAG.f90:410-411 (T=0)
    return res


def _insert_state(states, st):
    """es_insert_state_c (ED_EIGENSPACE.f90:169-218): ordered insert, before the first entry with e <= c%e."""
    pos = 0
    while pos < len(states) and not (st.e <= states[pos].e):
        pos += 1
    states.insert(pos, st)


def grids(p: Params):
    """allocate_grids (ED_AUX_FUNX.f90:449-461)."""
    wm = np.pi / p.beta * (2.0 * np.arange(1, p.Lmats + 1) - 1.0)
    wr = np.linspace(p.wini, p.wfin, p.Lreal)
    return wm, wr


def eigh_tridiag(alfa, beta):
    """eigh(diag, subdiag, Ev=Z) of ED_GF_NORMAL.f90:616-618 (LAPACK tridiagonal solver)."""
    n = len(alfa)
    T = np.diag(np.asarray(alfa, dtype=float))
    if n > 1:
        T += np.diag(np.asarray(beta[1:n], dtype=float), 1) + np.diag(np.asarray(beta[1:n], dtype=float), -1)
    return np.linalg.eigh(T)


def add_to_lanczos_gf(p: Params, res: Result, vnorm2, Ei, alanc, blanc, isign, iorb, jorb, ispin):
    """add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:580-632), T=0 branch."""
    pesoBZ = vnorm2 / res.zeta
    lam, Z = eigh_tridiag(alanc, blanc)
    for j in range(len(alanc)):
        de = lam[j] - Ei
        peso = pesoBZ * Z[0, j] * Z[0, j]
        res.impGmats[ispin, ispin, iorb, jorb, :] += peso / (1j * res.wm - isign * de)
        res.impGreal[ispin, ispin, iorb, jorb, :] += peso / ((res.wr + 1j * p.eps) - isign * de)


def build_gf(model: Model, res: Result):
    """build_gf_normal + lanc_build_gf_normal_c (ED_GF_NORMAL.f90:18-31, 116-260), bath_type=normal."""
    p = model.p
    Ns = p.Ns
    res.wm, res.wr = grids(p)
    shp = (p.Nspin, p.Nspin, p.Norb, p.Norb)
    res.impGmats = np.zeros(shp + (p.Lmats,), dtype=np.complex128)
    res.impGreal = np.zeros(shp + (p.Lreal,), dtype=np.complex128)
    for ispin in range(p.Nspin):
        for iorb in range(p.Norb):
            isite = (iorb + 1) if ispin == 0 else (iorb + 1 + Ns)              # impIndex, ED_SETUP.f90:443-446
            for istate, st in enumerate(res.states):
                mapI = build_sector(Ns, st.nup, st.ndw)
                for dagger, isign in ((1, 1), (0, -1)):
                    jup = st.nup + ((1 if dagger else -1) if ispin == 0 else 0)
                    jdw = st.ndw + ((1 if dagger else -1) if ispin == 1 else 0)
                    if not (0 <= jup <= Ns and 0 <= jdw <= Ns):                # getCDGsector/getCsector == 0
                        continue
                    mapJ = build_sector(Ns, jup, jdw)
                    vv, norm2 = apply_op(Ns, isite, dagger, mapI, mapJ, st.vec)
                    vv = vv / np.sqrt(norm2)
                    nlanc = min(mapJ.size, p.lanc_ngfiter)
                    alfa, beta, nused = lanc_tridiag(model, mapJ, vv, nlanc)
                    res.chains.append(dict(iorb=iorb, ispin=ispin, isign=isign, istate=istate, norm2=norm2,
                                           alfa=alfa, beta=beta, nused=nused, jup=jup, jdw=jdw))
                    add_to_lanczos_gf(p, res, norm2, st.e, alfa, beta, isign, iorb, iorb, ispin)
    return res


def apply_sz(Ns, Norb, iorb, smap, gs):
    """Seed of the spin-susceptibility chain (ED_GF_CHISPIN.f90:93-100; iorb = None: S_z^tot, :198-205):
    vvinit(m) = 1/2 (n_up - n_dw) gs(m) over the impurity level(s); returns the unnormalised vector."""
    words = np.asarray(smap, dtype=np.uint64)
    orbs = range(Norb) if iorb is None else [iorb]
    sgn = np.zeros(words.size)
    for a in orbs:
        sgn += ((words >> np.uint64(a)) & np.uint64(1)).astype(float) - ((words >> np.uint64(a + Ns)) & np.uint64(1)).astype(float)
    return 0.5 * sgn * gs


def add_to_lanczos_spinchi(p: Params, res: Result, vnorm, Ei, alanc, blanc, isign, iorb):
    """add_to_lanczos_spinChi (ED_GF_CHISPIN.f90:247-319), T = 0 (pesoBZ = 1)."""
    beta = p.beta
    pesoF = vnorm ** 2 / res.zeta
    lam, Z = eigh_tridiag(alanc, blanc)
    for j in range(len(alanc)):
        dE = lam[j] - Ei
        peso = pesoF * Z[0, j] * Z[0, j]
        ex = np.exp(-beta * dE)
        res.spinChi_iv[iorb, 0] += peso * beta if beta * dE < 1e-1 else peso * (1.0 - ex) / dE
        if isign == 1:
            res.spinChi_iv[iorb, 1:] += peso * (ex - 1.0) / (1j * res.vm[1:] - dE)
            res.spinChi_tau[iorb, :] += peso * np.exp(-res.tau * dE)
            res.spinChi_w[iorb, :] += peso * (ex - 1.0) / ((res.wr + 1j * p.eps) - dE)
        else:
            res.spinChi_iv[iorb, 1:] += peso * (1.0 - ex) / (1j * res.vm[1:] + dE)
            res.spinChi_tau[iorb, :] += peso * np.exp(-(beta - res.tau) * dE)
            res.spinChi_w[iorb, :] += peso * (1.0 - ex) / ((res.wr + 1j * p.eps) + dE)


def build_chi_spin(model: Model, res: Result):
    """buildChi_impurity / build_chi_spin (ED_GREENS_FUNCTIONS.f90:72-103, ED_GF_CHISPIN.f90:22-40): per kept state and
    orbital one Lanczos chain in the state's OWN sector seeded with S_z,a |gs> (:57-141), S_z^tot for Norb > 1 (:160-237).
    Quirk kept as is: the single-orbital routine passes the norm of the seed (:101), the total one its square (:206); both
    are squared again in add_to_lanczos_spinChi (:263); the final division by zeta_function (:36-38) comes on top of
    the one inside pesoF."""
    p = model.p
    Ns = p.Ns
    Ltau = max(int(p.beta), p.Ltau)                                            # ED_INPUT_VARS.f90:211
    res.vm = np.pi / p.beta * 2.0 * np.arange(0, p.Lmats + 1)                  # ED_AUX_FUNX.f90:452-458
    res.tau = np.linspace(0.0, p.beta, Ltau + 1)
    if res.wr is None:
        res.wm, res.wr = grids(p)
    res.spinChi_iv = np.zeros((p.Norb + 1, p.Lmats + 1), dtype=np.complex128)
    res.spinChi_tau = np.zeros((p.Norb + 1, Ltau + 1))
    res.spinChi_w = np.zeros((p.Norb + 1, p.Lreal), dtype=np.complex128)
    if not p.chispin_flag:
        return res
    chans = list(range(p.Norb)) + ([None] if p.Norb > 1 else [])
    for ic, iorb in enumerate(chans):
        for istate, st in enumerate(res.states):
            smap = build_sector(Ns, st.nup, st.ndw)
            vv = apply_sz(Ns, p.Norb, iorb, smap, st.vec)
            n2 = float(np.vdot(vv, vv).real)
            if n2 <= 0.0:
                continue
            nrm = np.sqrt(n2)
            nlanc = min(smap.size, p.lanc_ngfiter)
            alfa, beta, nused = lanc_tridiag(model, smap, vv / nrm, nlanc)
            res.chi_chains.append(dict(iorb=ic, istate=istate, norm=nrm, alfa=alfa, beta=beta, nused=nused))
            vnorm = n2 if iorb is None else nrm
            for isign in (1, -1):
                add_to_lanczos_spinchi(p, res, vnorm, st.e, alfa, beta, isign, ic)
    res.spinChi_tau /= res.zeta
    res.spinChi_w /= res.zeta
    res.spinChi_iv /= res.zeta
    return res


def apply_n(Ns, Norb, iorb, smap, gs):
    """Seed of the charge-susceptibility chain (ED_GF_CHIDENS.f90:126-133; iorb = None: total, :227-234):
    vvinit(m) = (n_up + n_dw) gs(m) over the impurity level(s)."""
    words = np.asarray(smap, dtype=np.uint64)
    orbs = range(Norb) if iorb is None else [iorb]
    sgn = np.zeros(words.size)
    for a in orbs:
        sgn += ((words >> np.uint64(a)) & np.uint64(1)).astype(float) + ((words >> np.uint64(a + Ns)) & np.uint64(1)).astype(float)
    return sgn * gs


def add_to_lanczos_denschi(p: Params, res: Result, vnorm2, Ei, alanc, blanc, isign, iv, tau_out, w_out):
    """add_to_lanczos_densChi / _tot (ED_GF_CHIDENS.f90:692-765, 876-948), T = 0; iv/tau_out/w_out are views of the channel.
    The sign of the static isign=+1 term (:727-731) is the reference's."""
    beta = p.beta
    pesoF = vnorm2 / res.zeta
    lam, Z = tql2(np.asarray(alanc, dtype=float), np.asarray(blanc, dtype=float)[1:])
    for j in range(len(alanc)):
        dE = lam[j] - Ei
        peso = pesoF * Z[0, j] * Z[0, j]
        ex = np.exp(-beta * dE)
        if isign == 1:
            iv[0] += -peso * beta if beta * dE < 1e-1 else peso * (ex - 1.0) / dE
            iv[1:] += peso * (ex - 1.0) / (1j * res.vm[1:] - dE)
            tau_out[:] += peso * np.exp(-res.tau * dE)
            w_out[:] += peso * (ex - 1.0) / ((res.wr + 1j * p.eps) - dE)
        else:
            iv[0] += peso * beta if beta * dE < 1e-1 else peso * (1.0 - ex) / dE
            iv[1:] += peso * (1.0 - ex) / (1j * res.vm[1:] + dE)
            tau_out[:] += peso * np.exp(-(beta - res.tau) * dE)
            w_out[:] += peso * (1.0 - ex) / ((res.wr + 1j * p.eps) + dE)


def build_chi_dens(model: Model, res: Result):
    """build_chi_dens (ED_GF_CHIDENS.f90:21-66), the channels with REAL seeds: densChi(a,a) (lanc_ed_build_densChi_diag_c
    :90-169) and densChi_tot for Norb > 1 (:191-269).  The inter-orbital / spin-mixed channels (:291-673) seed with
    (n_a + i n_b)|gs> (complex vectors) and are not restated: their entries stay zero."""
    p = model.p
    Ns = p.Ns
    Ltau = res.tau.size - 1
    res.densChi_iv = np.zeros((p.Norb, p.Norb, p.Lmats + 1), dtype=np.complex128)
    res.densChi_tau = np.zeros((p.Norb, p.Norb, Ltau + 1))
    res.densChi_w = np.zeros((p.Norb, p.Norb, p.Lreal), dtype=np.complex128)
    res.densChi_tot_iv = np.zeros(p.Lmats + 1, dtype=np.complex128)
    res.densChi_tot_tau = np.zeros(Ltau + 1)
    res.densChi_tot_w = np.zeros(p.Lreal, dtype=np.complex128)
    if not p.chidens_flag:
        return res
    chans = list(range(p.Norb)) + ([None] if p.Norb > 1 else [])
    for iorb in chans:
        for st in res.states:
            smap = build_sector(Ns, st.nup, st.ndw)
            vv = apply_n(Ns, p.Norb, iorb, smap, st.vec)
            n2 = float(np.vdot(vv, vv).real)
            if n2 <= 0.0:
                continue
            nlanc = min(smap.size, p.lanc_ngfiter)
            alfa, beta, nused = lanc_tridiag(model, smap, vv / np.sqrt(n2), nlanc)
            for isign in (1, -1):
                if iorb is None:
                    add_to_lanczos_denschi(p, res, n2, st.e, alfa, beta, isign, res.densChi_tot_iv, res.densChi_tot_tau, res.densChi_tot_w)
                else:
                    add_to_lanczos_denschi(p, res, n2, st.e, alfa, beta, isign, res.densChi_iv[iorb, iorb], res.densChi_tau[iorb, iorb],
                                           res.densChi_w[iorb, iorb])
    res.densChi_tau /= res.zeta                                                # :62-64 (not the total channel)
    res.densChi_w /= res.zeta
    res.densChi_iv /= res.zeta
    return res


def delta_bath(model: Model, x, ispin, iorb):
    """delta_bath_mats_main, normal/normal (ED_BATH_FUNCTIONS.f90:245-256)."""
    p = model.p
    eps = np.array([model.e(ispin, iorb, k) for k in range(p.Nbath)])
    vps = np.array([model.v(ispin, iorb, k) for k in range(p.Nbath)])
    return (vps[None, :] ** 2 / (x[:, None] - eps[None, :])).sum(axis=1)


def build_sigma(model: Model, res: Result):
    """build_sigma_normal (ED_GF_NORMAL.f90:656-694) with invg0_bath (ED_BATH_FUNCTIONS.f90:1784-1807)."""
    p = model.p
    shp = (p.Nspin, p.Nspin, p.Norb, p.Norb)
    res.impSmats = np.zeros(shp + (p.Lmats,), dtype=np.complex128)
    res.impSreal = np.zeros(shp + (p.Lreal,), dtype=np.complex128)
    res.impG0mats = np.zeros(shp + (p.Lmats,), dtype=np.complex128)
    res.impG0real = np.zeros(shp + (p.Lreal,), dtype=np.complex128)
    zm = 1j * res.wm
    zr = res.wr + 1j * p.eps
    for ispin in range(p.Nspin):
        for iorb in range(p.Norb):
            hl = model.hloc[ispin, ispin, iorb, iorb]
            for z, S, G, G0 in ((zm, res.impSmats, res.impGmats, res.impG0mats),
                                (zr, res.impSreal, res.impGreal, res.impG0real)):
                invg0 = z + p.xmu - hl - delta_bath(model, z, ispin, iorb)
                S[ispin, ispin, iorb, iorb, :] = invg0 - 1.0 / G[ispin, ispin, iorb, iorb, :]
                G0[ispin, ispin, iorb, iorb, :] = 1.0 / invg0
    return res


def observables(model: Model, res: Result):
    """observables_impurity core (ED_OBSERVABLES.f90:105-162), T=0."""
    p = model.p
    Ns = p.Ns
    n = p.Norb
    res.dens, res.dens_up, res.dens_dw, res.docc, res.magz = (np.zeros(n) for _ in range(5))
    res.sz2, res.n2 = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
    s2 = C.c_double(0.0)
    for st in res.states:
        peso = 1.0 / res.zeta
        smap = build_sector(Ns, st.nup, st.ndw)
        g = np.ascontiguousarray(st.vec, dtype=np.complex128)
        lib().ora_observables(Ns, n, _u64p(smap), smap.size, _dp(g.view(np.float64)), peso,
                              _dp(res.dens), _dp(res.dens_up), _dp(res.dens_dw), _dp(res.docc), _dp(res.magz),
                              _dp(res.sz2), _dp(res.n2), C.byref(s2))
    res.s2tot = s2.value
    return res


def ed_solve(p: Params, bath: np.ndarray, hloc=None, sectors=None) -> Result:
    """ed_solve_single (ED_MAIN.f90:253-282): diagonalize_impurity, buildgf_impurity, observables_impurity."""
    model = Model(p, bath, hloc)
    res = Result()
    ed_diag(model, res, sectors=sectors)
    build_gf(model, res)
    build_sigma(model, res)
    build_chi_spin(model, res)
    build_chi_dens(model, res)
    observables(model, res)
    res.model = model
    return res
