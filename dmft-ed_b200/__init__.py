"""dmft-ed_b200 -- B200-native Lanczos hot path of dmft-ed (ctypes view of libedgpu.so).

The product is the CUDA/C++ shared library `libedgpu.so` (C-ABI: include/edgpu.h, include/ed_b200.h).  This
module is a thin ctypes binding used by tests/ and bench.py; it holds no numerics of its own and has NO CPU
fallback: importing works without a GPU (so that symbol checks can run), every compute call needs one.

Import with  `importlib.import_module("dmft-ed_b200")`  (the directory name has a hyphen).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EDGPU_LIB_PATH") or os.path.join(_HERE, "libedgpu.so")   # override: kernel-variant experiments
_lib = None

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class EdgpuError(RuntimeError):
    pass


class edgpu_params(C.Structure):
    _fields_ = [("norb", C.c_int32), ("nbath", C.c_int32), ("nspin", C.c_int32), ("hfmode", C.c_int32),
                ("layout", C.c_int32), ("hxv_kernel", C.c_int32), ("reserved", C.c_int32 * 8)]


class ed_input(C.Structure):
    _fields_ = [("Norb", C.c_int32), ("Nbath", C.c_int32), ("Nspin", C.c_int32),
                ("uloc", C.c_double * 5), ("ust", C.c_double), ("jh", C.c_double), ("jx", C.c_double), ("jp", C.c_double),
                ("beta", C.c_double), ("xmu", C.c_double), ("hfmode", C.c_int32),
                ("Lmats", C.c_int32), ("Lreal", C.c_int32),
                ("wini", C.c_double), ("wfin", C.c_double), ("eps", C.c_double),
                ("gs_threshold", C.c_double), ("hwband", C.c_double),
                ("lanc_method", C.c_int32), ("lanc_nstates_sector", C.c_int32), ("lanc_nstates_total", C.c_int32),
                ("lanc_niter", C.c_int32), ("lanc_ngfiter", C.c_int32), ("lanc_tolerance", C.c_double),
                ("lanc_dim_threshold", C.c_int32), ("ed_twin", C.c_int32), ("ed_sparse_H", C.c_int32),
                ("ed_verbose", C.c_int32), ("gpu_layout", C.c_int32), ("gpu_hxv_kernel", C.c_int32),
                ("chispin_flag", C.c_int32), ("Ltau", C.c_int32), ("chidens_flag", C.c_int32), ("reserved", C.c_int32 * 5)]


# every symbol declared in include/edgpu.h and include/ed_b200.h (checked by tests/test_abi.py)
EDGPU_SYMBOLS = [
    "edgpu_init", "edgpu_finalize", "edgpu_bind_thread", "edgpu_sector_context", "edgpu_last_error", "edgpu_version", "edgpu_ns", "edgpu_set_hamiltonian",
    "edgpu_sector_build", "edgpu_sector_build_shard", "edgpu_sector_info", "edgpu_comm_unique_id", "edgpu_comm_init", "edgpu_comm_finalize", "edgpu_comm_info", "edgpu_comm_allreduce_host", "edgpu_vec_download_rows", "edgpu_sector_free", "edgpu_sector_dim", "edgpu_sector_map", "edgpu_sector_map_check",
    "edgpu_vec_alloc", "edgpu_vec_free", "edgpu_vec_upload", "edgpu_vec_download", "edgpu_vec_fill_normal", "edgpu_vec_fill_uniform",
    "edgpu_vec_copy", "edgpu_vec_dot", "edgpu_vec_scale", "edgpu_hxv", "edgpu_hxv_dev",
    "edgpu_sector_build_csr", "edgpu_sector_drop_csr", "edgpu_sector_csr_nnz", "edgpu_sector_csr_download",
    "edgpu_sector_dense", "edgpu_lanczos_gs", "edgpu_lanczos_eigs", "edgpu_lanczos_tridiag", "edgpu_apply_c", "edgpu_apply_sz", "edgpu_apply_n", "edgpu_observables",
    "edgpu_shard_ld", "edgpu_shard_hxv_dw", "edgpu_shard_hxv_up", "edgpu_shard_hxv_up_slabs", "edgpu_shard_perm",
    "edgpu_shard_hxv_up_peers", "edgpu_dev_alloc", "edgpu_dev_free", "edgpu_ipc_export", "edgpu_ipc_open", "edgpu_ipc_close", "edgpu_copy_async",
    "edgpu_bench_hxv", "edgpu_device_info", "edgpu_sync",
]
ED_SYMBOLS = [
    "ed_input_defaults", "ed_get_bath_dimension", "ed_init_solver", "ed_finalize_solver", "ed_comm_unique_id", "ed_set_comm", "ed_last_error", "ed_solve",
    "ed_get_sigma_matsubara", "ed_get_sigma_real", "ed_get_gimp_matsubara", "ed_get_gimp_real",
    "ed_get_g0imp_matsubara", "ed_get_g0imp_real", "ed_get_dens", "ed_get_dens_up", "ed_get_dens_dw", "ed_get_docc",
    "ed_get_mag", "ed_get_sz2_n2", "ed_get_grids", "ed_get_spinchi", "ed_get_denschi", "ed_get_state_count", "ed_get_state", "ed_get_state_vector",
    "ed_get_sector_energy", "ed_get_sector_nlanc", "ed_get_chain_count", "ed_get_chain", "ed_set_sectors_mask", "ed_get_timings",
    "ed_host_eigh", "ed_host_eigvals", "ed_host_eigh_tridiag",
]


def build(verbose: bool = False) -> str:
    """Compile libedgpu.so for sm_100a with nvcc (dmft-ed_b200/csrc/build.sh)."""
    out = subprocess.run(["bash", os.path.join(_HERE, "csrc", "build.sh")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libedgpu.so failed")
    return LIB_PATH


def lib():
    """Load libedgpu.so; fails loudly when the CUDA library has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EdgpuError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
                         "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.edgpu_last_error.restype = C.c_char_p
    L.edgpu_last_error.argtypes = [vp]
    L.edgpu_init.argtypes = [C.POINTER(edgpu_params), C.c_int, vp, C.POINTER(vp)]
    L.edgpu_finalize.argtypes = [vp]
    L.edgpu_ns.argtypes = [vp]
    L.edgpu_set_hamiltonian.argtypes = [vp, dp, C.c_int32, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.edgpu_sector_build.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.edgpu_sector_free.argtypes = [vp]
    L.edgpu_sector_build_shard.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.edgpu_sector_info.argtypes = [vp, i32p, i64p, i32p, i32p]
    L.edgpu_comm_unique_id.argtypes = [vp, C.c_char_p]
    L.edgpu_comm_init.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32]
    L.edgpu_comm_finalize.argtypes = [vp]
    L.edgpu_comm_info.argtypes = [vp, i32p, i32p]
    L.edgpu_comm_allreduce_host.argtypes = [vp, dp, C.c_int64, C.c_int32]
    L.ed_comm_unique_id.argtypes = [vp, C.c_char_p]
    L.ed_set_comm.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32]
    L.edgpu_vec_download_rows.argtypes = [vp, C.c_int64, C.c_int64, vp]
    L.edgpu_sector_dim.argtypes = [vp, i64p, i64p, i64p]
    L.edgpu_sector_map.argtypes = [vp, C.c_int64, C.c_int64, u64p]
    L.edgpu_sector_map_check.argtypes = [vp, u64p, i64p]
    L.edgpu_vec_alloc.argtypes = [vp, C.POINTER(vp)]
    L.edgpu_vec_free.argtypes = [vp]
    L.edgpu_vec_upload.argtypes = [vp, vp, C.c_int32]
    L.edgpu_vec_download.argtypes = [vp, vp, C.c_int32]
    L.edgpu_vec_fill_normal.argtypes = [vp, C.c_uint64]
    L.edgpu_vec_fill_uniform.argtypes = [vp, C.c_uint64]
    L.edgpu_vec_copy.argtypes = [vp, vp]
    L.edgpu_vec_dot.argtypes = [vp, vp, dp]
    L.edgpu_vec_scale.argtypes = [vp, C.c_double]
    L.edgpu_hxv.argtypes = [vp, C.c_int64, vp, vp]
    L.edgpu_hxv_dev.argtypes = [vp, vp, vp]
    L.edgpu_sector_build_csr.argtypes = [vp]
    L.edgpu_sector_drop_csr.argtypes = [vp]
    L.edgpu_sector_csr_nnz.argtypes = [vp, i64p]
    L.edgpu_sector_csr_download.argtypes = [vp, i64p, i64p, dp]
    L.edgpu_sector_dense.argtypes = [vp, dp]
    L.edgpu_lanczos_gs.argtypes = [vp, vp, C.c_int32, C.c_double, C.c_int32, dp, i32p, dp, dp]
    L.edgpu_lanczos_tridiag.argtypes = [vp, vp, C.c_int32, C.c_double, dp, dp, i32p]
    L.edgpu_lanczos_eigs.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_uint64, dp, C.POINTER(vp), i32p, i32p]
    L.edgpu_apply_c.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, dp]
    L.edgpu_apply_sz.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, dp]
    L.edgpu_apply_n.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, dp]
    L.edgpu_observables.argtypes = [vp, vp, C.c_double] + [dp] * 8
    L.edgpu_shard_ld.argtypes = [vp, i64p]
    L.edgpu_shard_hxv_dw.argtypes = [vp, C.c_int64, C.c_int64, vp, vp]
    L.edgpu_shard_hxv_up.argtypes = [vp, C.c_int64, C.c_int64, vp, vp, C.c_int32]
    L.edgpu_shard_hxv_up_slabs.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, i64p, i64p, vp, vp, C.c_int32]
    L.edgpu_shard_perm.argtypes = [vp, vp, vp]
    L.edgpu_shard_hxv_up_peers.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, i64p, i64p, C.POINTER(vp), i64p, C.POINTER(vp), C.c_int32]
    L.edgpu_copy_async.argtypes = [vp, vp, vp, C.c_int64, vp]
    L.edgpu_dev_alloc.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    L.edgpu_dev_free.argtypes = [vp, vp]
    L.edgpu_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.edgpu_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.edgpu_ipc_close.argtypes = [vp, vp]
    L.edgpu_bench_hxv.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, dp, i64p]
    L.edgpu_device_info.argtypes = [vp, i32p, i64p, i64p]
    L.edgpu_sync.argtypes = [vp]
    # host mirror
    L.ed_input_defaults.restype = None
    L.ed_input_defaults.argtypes = [C.POINTER(ed_input)]
    L.ed_get_bath_dimension.restype = C.c_int32
    L.ed_get_bath_dimension.argtypes = [C.POINTER(ed_input)]
    L.ed_init_solver.argtypes = [C.POINTER(ed_input), C.c_int, vp, dp, C.c_int32, dp, C.POINTER(vp)]
    L.ed_finalize_solver.argtypes = [vp]
    L.ed_last_error.restype = C.c_char_p
    L.ed_last_error.argtypes = [vp]
    L.ed_solve.argtypes = [vp, dp, C.c_int32, dp]
    for name in ("ed_get_sigma_matsubara", "ed_get_sigma_real", "ed_get_gimp_matsubara", "ed_get_gimp_real",
                 "ed_get_g0imp_matsubara", "ed_get_g0imp_real", "ed_get_dens", "ed_get_dens_up", "ed_get_dens_dw",
                 "ed_get_docc", "ed_get_mag"):
        getattr(L, name).argtypes = [vp, dp]
    L.ed_get_sz2_n2.argtypes = [vp, dp, dp, dp]
    L.ed_get_grids.argtypes = [vp, dp, dp]
    L.ed_get_spinchi.argtypes = [vp, dp, dp, dp, dp, dp, i32p]
    L.ed_get_denschi.argtypes = [vp, dp, dp, dp, dp, dp, dp]
    L.ed_get_state_count.argtypes = [vp, i32p, dp, dp]
    L.ed_get_state.argtypes = [vp, C.c_int32, dp, i32p, i32p]
    L.ed_get_state_vector.argtypes = [vp, C.c_int32, dp, C.c_int64]
    L.ed_get_sector_energy.argtypes = [vp, C.c_int32, C.c_int32, dp]
    L.ed_get_sector_nlanc.argtypes = [vp, C.c_int32, C.c_int32, i32p]
    L.ed_get_chain_count.argtypes = [vp, i32p]
    L.ed_get_chain.argtypes = [vp, C.c_int32, i32p, i32p, i32p, i32p, i32p, i32p, dp, dp, dp, C.c_int32]
    L.ed_set_sectors_mask.argtypes = [vp, i32p, C.c_int32]
    L.ed_get_timings.argtypes = [vp, dp]
    L.ed_host_eigh.argtypes = [C.c_int32, dp, dp]
    L.ed_host_eigvals.argtypes = [C.c_int32, dp, dp]
    L.ed_host_eigh_tridiag.argtypes = [C.c_int32, dp, dp, dp, dp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(dp)


# ----------------------------------------------------------------------------------------------------------
# object wrappers over the C-ABI (edgpu.h)
# ----------------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, norb, nbath, nspin=1, hfmode=True, device=-1, stream=None, layout=0, hxv_kernel=0, debug_flags=0):
        L = lib()
        p = edgpu_params(norb=norb, nbath=nbath, nspin=nspin, hfmode=int(hfmode), layout=layout, hxv_kernel=hxv_kernel)
        p.reserved[0] = debug_flags      # bit 0: 2-column down strips, bit 1: single-stage up pass (test hooks)
        h = C.c_void_p()
        if L.edgpu_init(C.byref(p), device, C.c_void_p(stream or 0), C.byref(h)) != 0:
            raise EdgpuError(L.edgpu_last_error(None).decode())
        self.h = h
        self.norb, self.nbath, self.nspin = norb, nbath, nspin
        self.ns = L.edgpu_ns(h)

    def check(self, rc):
        if rc != 0:
            raise EdgpuError(lib().edgpu_last_error(self.h).decode())

    def set_hamiltonian(self, bath, uloc, ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0, hloc=None):
        bath = np.ascontiguousarray(bath, dtype=np.float64)
        ul = np.zeros(5)
        ul[: len(uloc)] = uloc
        hp = None
        if hloc is not None:
            hl = np.asarray(hloc, dtype=np.complex128).reshape(self.nspin, self.nspin, self.norb, self.norb)
            self._hl = np.ascontiguousarray(hl.reshape(-1, order="F")).view(np.float64)
            hp = _p(self._hl)
        self.check(lib().edgpu_set_hamiltonian(self.h, _p(bath), bath.size, hp, _p(ul), ust, jh, jx, jp, xmu))

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(lib().edgpu_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank, nranks):
        self.check(lib().edgpu_comm_init(self.h, uid, rank, nranks))

    def sector_shard(self, nup, ndw, rank, nranks):
        return Sector(self, nup, ndw, rank, nranks)

    def sector(self, nup, ndw):
        return Sector(self, nup, ndw)

    def device_info(self):
        sm, l2, mem = C.c_int32(), C.c_int64(), C.c_int64()
        self.check(lib().edgpu_device_info(self.h, C.byref(sm), C.byref(l2), C.byref(mem)))
        return sm.value, l2.value, mem.value

    def sync(self):
        self.check(lib().edgpu_sync(self.h))

    def close(self):
        if self.h:
            lib().edgpu_finalize(self.h)
            self.h = None


class Sector:
    def __init__(self, ctx: Context, nup, ndw, rank=0, nranks=1):
        self.ctx = ctx
        h = C.c_void_p()
        if nranks > 1:
            ctx.check(lib().edgpu_sector_build_shard(ctx.h, nup, ndw, rank, nranks, C.byref(h)))
        else:
            ctx.check(lib().edgpu_sector_build(ctx.h, nup, ndw, C.byref(h)))
        self.h = h
        d, du, dd = C.c_int64(), C.c_int64(), C.c_int64()
        lib().edgpu_sector_dim(h, C.byref(d), C.byref(du), C.byref(dd))
        self.dim, self.dim_up, self.dim_dw = d.value, du.value, dd.value
        self.nup, self.ndw = nup, ndw

    def info(self):
        k, n, r, nr = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        self.ctx.check(lib().edgpu_sector_info(self.h, C.byref(k), C.byref(n), C.byref(r), C.byref(nr)))
        return {"layout_kind": k.value, "nalloc": n.value, "shard_rank": r.value, "shard_nranks": nr.value}

    def map(self, first=0, count=None):
        count = self.dim - first if count is None else count
        out = np.empty(count, dtype=np.uint64)
        self.ctx.check(lib().edgpu_sector_map(self.h, first, count, out.ctypes.data_as(u64p)))
        return out

    def map_check(self):
        cs, v = C.c_uint64(), C.c_int64()
        self.ctx.check(lib().edgpu_sector_map_check(self.h, C.byref(cs), C.byref(v)))
        return cs.value, v.value

    def vec(self, host=None):
        v = Vec(self)
        if host is not None:
            v.upload(host)
        return v

    def hxv_host(self, v):
        """edgpu_hxv: host complex in/out (the spHtimesV_cc parity hook)."""
        vin = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.empty_like(vin)
        self.ctx.check(lib().edgpu_hxv(self.h, vin.size, vin.ctypes.data, out.ctypes.data))
        return out

    def hxv(self, x, y):
        self.ctx.check(lib().edgpu_hxv_dev(self.h, x.h, y.h))

    def build_csr(self):
        self.ctx.check(lib().edgpu_sector_build_csr(self.h))

    def drop_csr(self):
        self.ctx.check(lib().edgpu_sector_drop_csr(self.h))

    def csr(self):
        n = C.c_int64()
        self.ctx.check(lib().edgpu_sector_csr_nnz(self.h, C.byref(n)))
        rp = np.zeros(self.dim + 1, dtype=np.int64)
        cols = np.zeros(n.value, dtype=np.int64)
        vals = np.zeros(n.value)
        self.ctx.check(lib().edgpu_sector_csr_download(self.h, rp.ctypes.data_as(i64p), cols.ctypes.data_as(i64p), _p(vals)))
        return rp, cols, vals

    def dense(self):
        H = np.zeros((self.dim, self.dim), order="F")
        self.ctx.check(lib().edgpu_sector_dense(self.h, _p(H)))
        return H

    def lanczos_gs(self, v0, nitermax, threshold=1e-12, ncheck=10):
        e0, nl = C.c_double(), C.c_int32()
        a = np.zeros(nitermax + 1)
        b = np.zeros(nitermax + 1)
        self.ctx.check(lib().edgpu_lanczos_gs(self.h, v0.h, nitermax, threshold, ncheck, C.byref(e0), C.byref(nl), _p(a), _p(b)))
        return e0.value, nl.value, a[: nl.value], b[: nl.value]

    def lanczos_eigs(self, neigen, ncv=None, tol=1e-12, maxrestart=300, seed=1234567):
        """edgpu_lanczos_eigs: (evals[neigen], [Vec...], nconv, nmatvec)"""
        ncv = ncv or 10 * neigen
        ev = np.zeros(neigen)
        hs = (C.c_void_p * neigen)()
        nc, nm = C.c_int32(), C.c_int32()
        self.ctx.check(lib().edgpu_lanczos_eigs(self.h, neigen, ncv, maxrestart, tol, seed, _p(ev), hs, C.byref(nc), C.byref(nm)))
        vecs = []
        for h in hs:
            v = Vec.__new__(Vec)
            v.s, v.h = self, C.c_void_p(h)
            vecs.append(v)
        return ev, vecs, nc.value, nm.value

    def lanczos_tridiag(self, v, nlanc, threshold=1e-13):
        a = np.zeros(nlanc)
        b = np.zeros(nlanc)
        nu = C.c_int32()
        self.ctx.check(lib().edgpu_lanczos_tridiag(self.h, v.h, nlanc, threshold, _p(a), _p(b), C.byref(nu)))
        return a, b, nu.value

    def observables(self, gs, peso=1.0):
        n = self.ctx.norb
        out = dict(dens=np.zeros(n), dens_up=np.zeros(n), dens_dw=np.zeros(n), docc=np.zeros(n), magz=np.zeros(n),
                   sz2=np.zeros((n, n), order="F"), n2=np.zeros((n, n), order="F"))
        s2 = C.c_double(0.0)
        self.ctx.check(lib().edgpu_observables(self.h, gs.h, peso, _p(out["dens"]), _p(out["dens_up"]), _p(out["dens_dw"]),
                                               _p(out["docc"]), _p(out["magz"]), _p(out["sz2"]), _p(out["n2"]), C.byref(s2)))
        out["s2tot"] = s2.value
        return out

    def bench_hxv(self, x, y, iters, flush_l2=False):
        ms, nl = C.c_double(), C.c_int64()
        self.ctx.check(lib().edgpu_bench_hxv(self.h, x.h, y.h, iters, int(flush_l2), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def free(self):
        if self.h:
            lib().edgpu_sector_free(self.h)
            self.h = None


class Vec:
    def __init__(self, sector: Sector):
        self.s = sector
        h = C.c_void_p()
        sector.ctx.check(lib().edgpu_vec_alloc(sector.h, C.byref(h)))
        self.h = h

    def upload(self, host):
        a = np.ascontiguousarray(host)
        if np.iscomplexobj(a):
            a = a.astype(np.complex128, copy=False)
            self.s.ctx.check(lib().edgpu_vec_upload(self.h, a.ctypes.data, 1))
        else:
            a = a.astype(np.float64, copy=False)
            self.s.ctx.check(lib().edgpu_vec_upload(self.h, a.ctypes.data, 0))
        return self

    def download(self, cplx=False):
        out = np.empty(self.s.dim, dtype=np.complex128 if cplx else np.float64)
        self.s.ctx.check(lib().edgpu_vec_download(self.h, out.ctypes.data, int(cplx)))
        return out

    def download_rows(self, rd0, rd1):
        out = np.empty((rd1 - rd0) * self.s.dim_up, dtype=np.float64)
        self.s.ctx.check(lib().edgpu_vec_download_rows(self.h, rd0, rd1, out.ctypes.data))
        return out

    def fill_normal(self, seed):
        self.s.ctx.check(lib().edgpu_vec_fill_normal(self.h, seed))
        return self

    def fill_uniform(self, seed):
        self.s.ctx.check(lib().edgpu_vec_fill_uniform(self.h, seed))
        return self

    def dot(self, other):
        out = C.c_double()
        self.s.ctx.check(lib().edgpu_vec_dot(self.h, other.h, C.byref(out)))
        return out.value

    def scale(self, alpha):
        self.s.ctx.check(lib().edgpu_vec_scale(self.h, alpha))

    def free(self):
        if self.h:
            lib().edgpu_vec_free(self.h)
            self.h = None


def apply_c(s_in: Sector, s_out: Sector, isite, dagger, vin: Vec, vout: Vec, normalise=True):
    n2 = C.c_double()
    s_in.ctx.check(lib().edgpu_apply_c(s_in.h, s_out.h, isite, int(dagger), vin.h, vout.h, int(normalise), C.byref(n2)))
    return n2.value


# ----------------------------------------------------------------------------------------------------------
# host mirror (ed_b200.h)
# ----------------------------------------------------------------------------------------------------------
def default_input(**kw) -> ed_input:
    inp = ed_input()
    lib().ed_input_defaults(C.byref(inp))
    for k, v in kw.items():
        if k == "uloc":
            for i, u in enumerate(v):
                inp.uloc[i] = u
        elif k == "lanc_method":
            inp.lanc_method = {"arpack": 0, "lanczos": 1}.get(v, v)
        elif k == "workers":                 # host worker threads of ed_solve (ed_input.reserved[2]; 0 = default 4)
            inp.reserved[2] = int(v)
        else:
            setattr(inp, k, v)
    return inp


class Solver:
    """ed_init_solver / ed_solve / ed_get_* (ED_MAIN.f90, ED_IO) through the C++ host mirror."""

    def __init__(self, inp: ed_input, hloc=None, device=-1, stream=None):
        L = lib()
        self.inp = inp
        self.nbath_len = L.ed_get_bath_dimension(C.byref(inp))
        self.bath = np.zeros(self.nbath_len)
        self._hl = None
        hp = None
        if hloc is not None:
            hl = np.asarray(hloc, dtype=np.complex128).reshape(inp.Nspin, inp.Nspin, inp.Norb, inp.Norb)
            self._hl = np.ascontiguousarray(hl.reshape(-1, order="F")).view(np.float64)
            hp = _p(self._hl)
        h = C.c_void_p()
        if L.ed_init_solver(C.byref(inp), device, C.c_void_p(stream or 0), _p(self.bath), self.nbath_len, hp, C.byref(h)) != 0:
            raise EdgpuError("ed_init_solver failed: " + L.edgpu_last_error(None).decode())
        self.h = h

    def check(self, rc):
        if rc != 0:
            raise EdgpuError(lib().ed_last_error(self.h).decode())

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(lib().ed_comm_unique_id(self.h, buf))
        return buf.raw

    def set_comm(self, uid: bytes, rank, nranks):
        """distributed ed_solve: sectors and ground-state chains dealt over the ranks (ED_MAIN.f90:598-636 analogue)"""
        self.check(lib().ed_set_comm(self.h, uid, rank, nranks))

    def set_sectors(self, pairs):
        a = np.ascontiguousarray(np.array(pairs, dtype=np.int32).reshape(-1))
        self.check(lib().ed_set_sectors_mask(self.h, a.ctypes.data_as(i32p), a.size // 2))

    def solve(self, bath=None):
        b = np.ascontiguousarray(self.bath if bath is None else bath, dtype=np.float64)
        self.check(lib().ed_solve(self.h, _p(b), b.size, None))

    def _cget(self, name, L):
        i = self.inp
        out = np.zeros((i.Nspin, i.Nspin, i.Norb, i.Norb, L), dtype=np.complex128, order="F")
        self.check(getattr(lib(), name)(self.h, out.ctypes.data_as(dp)))
        return out

    def sigma_matsubara(self):
        return self._cget("ed_get_sigma_matsubara", self.inp.Lmats)

    def sigma_real(self):
        return self._cget("ed_get_sigma_real", self.inp.Lreal)

    def gimp_matsubara(self):
        return self._cget("ed_get_gimp_matsubara", self.inp.Lmats)

    def gimp_real(self):
        return self._cget("ed_get_gimp_real", self.inp.Lreal)

    def g0imp_matsubara(self):
        return self._cget("ed_get_g0imp_matsubara", self.inp.Lmats)

    def _dget(self, name):
        out = np.zeros(self.inp.Norb)
        self.check(getattr(lib(), name)(self.h, _p(out)))
        return out

    def dens(self):
        return self._dget("ed_get_dens")

    def docc(self):
        return self._dget("ed_get_docc")

    def mag(self):
        return self._dget("ed_get_mag")

    def sz2_n2(self):
        n = self.inp.Norb
        a, b, s = np.zeros((n, n), order="F"), np.zeros((n, n), order="F"), C.c_double()
        self.check(lib().ed_get_sz2_n2(self.h, _p(a), _p(b), C.byref(s)))
        return a, b, s.value

    def spinchi(self):
        """(chi_iv[Norb+1, 0:Lmats], chi_tau[Norb+1, 0:Ltau], chi_w[Norb+1, Lreal], vm, tau) of build_chi_spin."""
        lt = C.c_int32()
        self.check(lib().ed_get_spinchi(self.h, None, None, None, None, None, C.byref(lt)))
        n1, lm, lr, ltau = self.inp.Norb + 1, self.inp.Lmats, self.inp.Lreal, lt.value
        iv = np.zeros((n1, lm + 1), dtype=np.complex128, order="F")
        ct = np.zeros((n1, ltau + 1), order="F")
        cw = np.zeros((n1, lr), dtype=np.complex128, order="F")
        vm, tau = np.zeros(lm + 1), np.zeros(ltau + 1)
        self.check(lib().ed_get_spinchi(self.h, C.cast(iv.ctypes.data, dp), _p(ct), C.cast(cw.ctypes.data, dp),
                                        _p(vm), _p(tau), C.byref(lt)))
        return iv, ct, cw, vm, tau

    def denschi(self):
        """(chi_iv[Norb,Norb,0:Lmats], chi_tau[Norb,Norb,0:Ltau], chi_w[Norb,Norb,Lreal], tot_iv, tot_tau, tot_w) of build_chi_dens
        (diagonal and total channels)."""
        lt = C.c_int32()
        self.check(lib().ed_get_spinchi(self.h, None, None, None, None, None, C.byref(lt)))
        n, lm, lr, ltau = self.inp.Norb, self.inp.Lmats, self.inp.Lreal, lt.value
        iv = np.zeros((n, n, lm + 1), dtype=np.complex128, order="F")
        ct = np.zeros((n, n, ltau + 1), order="F")
        cw = np.zeros((n, n, lr), dtype=np.complex128, order="F")
        tiv, tt, tw = np.zeros(lm + 1, dtype=np.complex128), np.zeros(ltau + 1), np.zeros(lr, dtype=np.complex128)
        self.check(lib().ed_get_denschi(self.h, C.cast(iv.ctypes.data, dp), _p(ct), C.cast(cw.ctypes.data, dp),
                                        C.cast(tiv.ctypes.data, dp), _p(tt), C.cast(tw.ctypes.data, dp)))
        return iv, ct, cw, tiv, tt, tw

    def states(self):
        n, z, e = C.c_int32(), C.c_double(), C.c_double()
        self.check(lib().ed_get_state_count(self.h, C.byref(n), C.byref(z), C.byref(e)))
        out = []
        for i in range(n.value):
            ee, nu, nd = C.c_double(), C.c_int32(), C.c_int32()
            lib().ed_get_state(self.h, i, C.byref(ee), C.byref(nu), C.byref(nd))
            out.append((ee.value, nu.value, nd.value))
        return out, z.value, e.value

    def sector_energy(self, nup, ndw):
        e = C.c_double()
        self.check(lib().ed_get_sector_energy(self.h, nup, ndw, C.byref(e)))
        return e.value

    def sector_nlanc(self, nup, ndw):
        n = C.c_int32()
        self.check(lib().ed_get_sector_nlanc(self.h, nup, ndw, C.byref(n)))
        return n.value

    def chains(self):
        n = C.c_int32()
        self.check(lib().ed_get_chain_count(self.h, C.byref(n)))
        out = []
        cap = max(int(self.inp.lanc_ngfiter), 1)
        for i in range(n.value):
            io, isp, isg, ist, nl, nu = (C.c_int32() for _ in range(6))
            n2 = C.c_double()
            a, b = np.zeros(cap), np.zeros(cap)
            lib().ed_get_chain(self.h, i, C.byref(io), C.byref(isp), C.byref(isg), C.byref(ist), C.byref(nl), C.byref(nu),
                               C.byref(n2), _p(a), _p(b), cap)
            out.append(dict(iorb=io.value, ispin=isp.value, isign=isg.value, istate=ist.value, nlanc=nl.value,
                            nused=nu.value, norm2=n2.value, alfa=a[: nl.value].copy(), beta=b[: nl.value].copy()))
        return out

    def timings(self):
        t = np.zeros(4)
        lib().ed_get_timings(self.h, _p(t))
        return dict(diag=t[0], gf=t[1], sigma=t[2], observables=t[3])

    def close(self):
        if self.h:
            lib().ed_finalize_solver(self.h)
            self.h = None
