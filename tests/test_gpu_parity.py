"""GPU (B200): parity of the CUDA path (through the C-ABI of libedgpu.so) against the CPU oracle.

Tolerances from BASELINE.json north_star: sector maps bit-exact; E0 1e-10 relative; Lanczos alpha/beta (GF
chains), densities, double occupancy 1e-9; G_imp / Sigma(iw) 1e-8.  H*v itself is compared at 1e-12.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HXV_TOL = 1e-12

CASES = {
    "cfg1": dict(Norb=1, Nbath=4),
    "nohf_mu": dict(Norb=1, Nbath=3, xmu=0.3, hfmode=False),
    "2orb_hund": dict(Norb=2, Nbath=2, uloc=(2.0, 1.5), ust=1.2, jh=0.3),
    "2orb_jxjp": dict(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, jx=0.25, jp=0.25),
    "nspin2": dict(Norb=1, Nbath=3, Nspin=2),
    "3orb_jxjp": dict(Norb=3, Nbath=1, uloc=(2.0, 2.0, 2.0), ust=1.0, jh=0.2, jx=0.2, jp=0.2),
}


def make(oracle, edb, case, seed=11, generic_bath=True, hloc_offdiag=False, hloc_diag=True, layout=0, hxv_kernel=0, debug_flags=0):
    kw = dict(lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64)
    kw.update(case)
    p = oracle.Params(**kw)
    rng = np.random.default_rng(seed)
    bath = oracle.init_bath(p)
    if generic_bath:
        bath = bath + 0.05 * rng.normal(size=bath.size)
    hloc = np.zeros((p.Nspin, p.Nspin, p.Norb, p.Norb), dtype=complex)
    for s in range(p.Nspin):
        for a in range(p.Norb):
            hloc[s, s, a, a] = 0.1 * (a + 1) * (1 if s == 0 else -1) if hloc_diag else 0.0
        if hloc_offdiag and p.Norb > 1:
            hloc[s, s, 0, 1] = hloc[s, s, 1, 0] = 0.15
    model = oracle.Model(p, bath, hloc)
    ctx = edb.Context(p.Norb, p.Nbath, p.Nspin, p.hfmode, layout=layout, hxv_kernel=hxv_kernel, debug_flags=debug_flags)
    ctx.set_hamiltonian(bath, p.uloc, p.ust, p.jh, p.jx, p.jp, p.xmu, hloc=hloc)
    return p, model, ctx, rng


def all_sectors(Ns):
    return [(a, b) for a in range(Ns + 1) for b in range(Ns + 1)]


# ---------------------------------------------------------------------------------------------- sector maps
@pytest.mark.parametrize("Norb,Nbath", [(1, 4), (2, 2), (1, 7)])
def test_sector_maps_bit_exact_all_sectors(oracle, edb, Norb, Nbath):
    p, model, ctx, _ = make(oracle, edb, dict(Norb=Norb, Nbath=Nbath))
    Ns = p.Ns
    secs = all_sectors(Ns) if Ns <= 6 else [(a, b) for a in range(0, Ns + 1, 2) for b in range(1, Ns + 1, 3)]
    for nup, ndw in secs:
        s = ctx.sector(nup, ndw)
        ref = oracle.build_sector(Ns, nup, ndw, literal=(Ns <= 6))
        assert s.dim == ref.size == oracle.sector_dim(Ns, nup, ndw)
        assert np.array_equal(s.map(), ref)
        cs, viol = s.map_check()
        idx = np.arange(ref.size, dtype=np.uint64)
        assert viol == 0 and cs == int((ref * (np.uint64(2) * idx + np.uint64(1))).sum(dtype=np.uint64))
        s.free()
    ctx.close()


def test_sector_map_cfg2_and_ranged_download(oracle, edb):
    p, model, ctx, _ = make(oracle, edb, dict(Norb=1, Nbath=9))
    s = ctx.sector(5, 5)
    ref = oracle.build_sector(10, 5, 5)
    assert s.dim == 63504 and np.array_equal(s.map(), ref)
    assert np.array_equal(s.map(1000, 777), ref[1000:1777])
    s.free()
    ctx.close()


def test_sector_map_full_size_properties(oracle, edb):
    """Ns=14 and Ns=16 half filling: dimension, ordering and popcounts checked on the device without
    materialising the map; head/tail windows compared bit-exactly with the oracle formula."""
    for Norb, Nbath, n, dim in [(2, 6, 7, 11778624), (2, 7, 8, 165636900)]:
        p, model, ctx, _ = make(oracle, edb, dict(Norb=Norb, Nbath=Nbath, uloc=(2.0, 2.0)))
        s = ctx.sector(n, n)
        assert s.dim == dim
        cs, viol = s.map_check()
        assert viol == 0
        Ns = p.Ns
        ups = np.array([w for w in range(1 << Ns) if bin(w).count("1") == n], dtype=np.uint64)
        assert ups.size == s.dim_up
        k = min(4096, int(s.dim_up))
        head = s.map(0, k)
        assert np.array_equal(head, (ups[:k] + (ups[0] << np.uint64(Ns))))
        tail = s.map(s.dim - k, k)
        assert np.array_equal(tail, ups[-k:] + (ups[-1] << np.uint64(Ns)))
        # checksum of checksums: sum_i map[i]*(2i+1) factorises over the two spin lists
        du = s.dim_up
        iu = np.arange(du, dtype=np.uint64)
        tot = np.uint64(0)
        with np.errstate(over="ignore"):
            su0, su1 = ups.sum(dtype=np.uint64), (ups * (np.uint64(2) * iu + np.uint64(1))).sum(dtype=np.uint64)
            s1 = (np.uint64(2) * iu + np.uint64(1)).sum(dtype=np.uint64)
            for rd in range(du):
                base = np.uint64(2 * rd * du)
                hi = ups[rd] << np.uint64(Ns)
                tot += su1 + base * su0 + hi * (s1 + base * np.uint64(du))
        assert cs == int(tot)
        s.free()
        ctx.close()


# ---------------------------------------------------------------------------------------------- H*v
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("hloc_offdiag", [False, True])
def test_hxv_matches_oracle_all_sectors(oracle, edb, name, hloc_offdiag):
    p, model, ctx, rng = make(oracle, edb, CASES[name], hloc_offdiag=hloc_offdiag)
    Ns = p.Ns
    for nup, ndw in all_sectors(Ns):
        smap = oracle.build_sector(Ns, nup, ndw)
        s = ctx.sector(nup, ndw)
        v = rng.normal(size=smap.size) + 1j * rng.normal(size=smap.size)
        ref = oracle.direct_hxv(model, smap, v)                      # literal scatter form, complex(8)
        got = s.hxv_host(v)                                          # spHtimesV_cc hook (host complex in/out)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < HXV_TOL * scale, (name, nup, ndw)
        # device-resident real path
        x, y = s.vec(v.real), s.vec()
        s.hxv(x, y)
        assert np.abs(y.download() - ref.real).max() < HXV_TOL * scale
        x.free(); y.free(); s.free()
    ctx.close()


def test_hxv_cfg2_half_filling_and_linearity(oracle, edb):
    p, model, ctx, rng = make(oracle, edb, dict(Norb=1, Nbath=9), generic_bath=False)
    smap = oracle.build_sector(10, 5, 5)
    s = ctx.sector(5, 5)
    v = oracle.philox_normal(20240607, smap.size)
    ref = oracle.direct_hxv(model, smap, v).real
    x, y, z = s.vec(v), s.vec(), s.vec()
    s.hxv(x, y)
    assert np.abs(y.download() - ref).max() < HXV_TOL * np.abs(ref).max()
    # device Philox fill agrees with the host generator (libm differences only)
    z.fill_normal(20240607)
    assert np.abs(z.download() - v).max() < 1e-12
    z.fill_uniform(1234567)
    assert np.array_equal(z.download(), oracle.philox_uniform(1234567, smap.size))      # bit-identical
    # symmetry <a|H b> = <H a|b>
    w = rng.normal(size=smap.size)
    a, ha = s.vec(w), s.vec()
    s.hxv(a, ha)
    assert abs(a.dot(y) - ha.dot(x)) < 1e-9 * abs(a.dot(y))
    for t in (x, y, z, a, ha):
        t.free()
    s.free()
    ctx.close()


def test_hxv_rejects_bad_input(edb, oracle):
    p, model, ctx, rng = make(oracle, edb, dict(Norb=1, Nbath=4))
    s = ctx.sector(2, 3)
    with pytest.raises(edb.EdgpuError, match="Nloc != dim"):
        s.hxv_host(np.zeros(s.dim + 1, dtype=complex))
    with pytest.raises(edb.EdgpuError):
        ctx.sector(6, 0)
    # the device vectors are real: a complex(8) vector keeps its real part only if the imaginary part is exactly zero
    z = rng.standard_normal(s.dim) + 0j
    v = s.vec().upload(z)
    assert np.array_equal(v.download(), z.real)
    z[s.dim // 2] += 1e-30j
    with pytest.raises(edb.EdgpuError, match="non-zero imaginary part"):
        v.upload(z)
    v.free()
    hl = np.zeros((1, 1, 1, 1), dtype=complex)
    hl[0, 0, 0, 0] = 1j
    with pytest.raises(edb.EdgpuError, match="complex impHloc"):
        ctx.set_hamiltonian(oracle.init_bath(p), p.uloc, hloc=hl)
    with pytest.raises(edb.EdgpuError, match="wrong bath dimensions"):
        ctx.set_hamiltonian(np.zeros(3), p.uloc)
    ctx.close()


# ---------------------------------------------------------------------------------------------- stored CSR
@pytest.mark.parametrize("name", ["cfg1", "2orb_jxjp", "nspin2"])
def test_csr_equals_reference_row_lists(oracle, edb, name):
    p, model, ctx, rng = make(oracle, edb, CASES[name], hloc_offdiag=True)
    Ns = p.Ns
    for nup, ndw in [(Ns // 2, Ns // 2), (Ns // 2 + 1, Ns // 2), (1, Ns - 1), (0, Ns)]:
        smap = oracle.build_sector(Ns, nup, ndw)
        rp0, c0, v0 = oracle.stored_build(model, smap)
        s = ctx.sector(nup, ndw)
        s.build_csr()
        rp, c, v = s.csr()
        assert np.array_equal(rp, rp0) and np.array_equal(c, c0)       # same entries in the same insertion order
        assert np.abs(v - v0.real).max() < 1e-13 and np.abs(v0.imag).max() == 0
        x = rng.normal(size=smap.size) + 1j * rng.normal(size=smap.size)
        ref = oracle.stored_hxv(rp0, c0, v0, x)
        assert np.abs(s.hxv_host(x) - ref).max() < HXV_TOL * max(1.0, np.abs(ref).max())
        s.drop_csr()
        s.free()
    ctx.close()


def test_dense_matrix_small_sector(oracle, edb):
    p, model, ctx, rng = make(oracle, edb, CASES["2orb_hund"])
    smap = oracle.build_sector(6, 3, 2)
    s = ctx.sector(3, 2)
    H0 = oracle.dense_h(model, smap)
    assert np.abs(s.dense() - H0.real).max() < 1e-13
    s.free()
    ctx.close()


# ---------------------------------------------------------------------------------------------- Lanczos
def test_lanczos_gs_cfg2(oracle, edb):
    p, model, ctx, rng = make(oracle, edb, dict(Norb=1, Nbath=9), generic_bath=False, hloc_diag=False)
    smap = oracle.build_sector(10, 5, 5)
    v0 = oracle.start_vector(smap.size)
    e_ref, vec_ref, nl_ref, a_ref, b_ref = oracle.lanc_gs(model, smap, v0, 512, 1e-12)
    s = ctx.sector(5, 5)
    v = s.vec(v0.real)
    e0, nl, a, b = s.lanczos_gs(v, 512, 1e-12)
    assert abs(e0 - e_ref) < 1e-10 * abs(e_ref)
    assert abs(e0 - (-11.341244826804)) < 1e-9                      # BASELINE.md anchor
    gs = v.download()
    assert abs(abs(gs @ vec_ref.real) - 1.0) < 1e-9
    assert abs(np.linalg.norm(gs) - 1.0) < 1e-12
    # the early Lanczos coefficients are comparable (same start vector)
    k = 20
    assert np.abs(a[:k] - a_ref[:k]).max() < 1e-9 and np.abs(b[1:k] - b_ref[1:k]).max() < 1e-9
    obs = s.observables(v)
    # a plain-Lanczos Ritz vector is only ~sqrt(1e-12) accurate, so analytic values hold to ~1e-7 ...
    assert abs(obs["dens"][0] - 1.0) < 1e-6 and abs(obs["docc"][0] - 0.162122572520) < 1e-6
    # ... while oracle and device, which run the same recurrence from the same start vector, agree to 1e-9
    up = (smap & np.uint64(1)).astype(float)
    dw = ((smap >> np.uint64(10)) & np.uint64(1)).astype(float)
    w = np.abs(vec_ref) ** 2
    assert abs(obs["dens"][0] - ((up + dw) * w).sum()) < 1e-9 and abs(obs["docc"][0] - (up * dw * w).sum()) < 1e-9
    v.free(); s.free(); ctx.close()


@pytest.mark.parametrize("name,sec,isite,dagger", [
    ("cfg1", (2, 3), 1, 1), ("cfg1", (2, 3), 1, 0), ("nspin2", (2, 2), 5, 1), ("nspin2", (2, 2), 5, 0),
    ("2orb_hund", (3, 3), 2, 1), ("2orb_hund", (3, 2), 8, 0),
])
def test_apply_c_and_tridiag_chain(oracle, edb, name, sec, isite, dagger):
    p, model, ctx, rng = make(oracle, edb, CASES[name])
    Ns = p.Ns
    nup, ndw = sec
    mapI = oracle.build_sector(Ns, nup, ndw)
    g = rng.normal(size=mapI.size)
    g /= np.linalg.norm(g)
    d = 1 if dagger else -1
    jup, jdw = (nup + d, ndw) if isite <= Ns else (nup, ndw + d)
    mapJ = oracle.build_sector(Ns, jup, jdw)
    vv_ref, n2_ref = oracle.apply_op(Ns, isite, dagger, mapI, mapJ, g)
    si, sj = ctx.sector(nup, ndw), ctx.sector(jup, jdw)
    vin, vout = si.vec(g), sj.vec()
    n2 = edb.apply_c(si, sj, isite, dagger, vin, vout, normalise=False)
    assert abs(n2 - n2_ref) < 1e-13
    assert np.array_equal(vout.download(), vv_ref.real)               # pure signed permutation: bit-exact
    nlanc = min(mapJ.size, 40)
    a_ref, b_ref, nu_ref = oracle.lanc_tridiag(model, mapJ, vv_ref / np.sqrt(n2_ref), nlanc)
    vout.scale(1.0 / np.sqrt(n2))
    a, b, nu = sj.lanczos_tridiag(vout, nlanc)
    assert nu == nu_ref
    # alpha/beta to 1e-9 over the leading part; without re-orthogonalisation the trailing coefficients of a chain
    # that exhausts a small sector amplify rounding noise, so the tail is compared through the
    # quantity it feeds: the pole sum of add_to_lanczos_gf_normal
    k = nlanc if mapJ.size >= 200 else min(nlanc, 12)
    assert np.abs(a[:k] - a_ref[:k]).max() < 1e-9 and np.abs(b[:k] - b_ref[:k]).max() < 1e-9
    lam, Z = oracle.eigh_tridiag(a, b)
    lam0, Z0 = oracle.eigh_tridiag(a_ref, b_ref)
    for zz in (1.0j, 3.0j, 0.5 + 1.0j):
        assert abs((Z[0] ** 2 / (zz - lam)).sum() - (Z0[0] ** 2 / (zz - lam0)).sum()) < 1e-8
    with pytest.raises(edb.EdgpuError):
        edb.apply_c(si, si, isite, dagger, vin, vin)
    for t in (vin, vout):
        t.free()
    si.free(); sj.free(); ctx.close()


def test_observables_match_oracle(oracle, edb):
    import ctypes as C
    p, model, ctx, rng = make(oracle, edb, CASES["2orb_hund"])
    smap = oracle.build_sector(6, 3, 2)
    g = rng.normal(size=smap.size)
    g /= np.linalg.norm(g)
    n = p.Norb
    ref = dict(dens=np.zeros(n), dens_up=np.zeros(n), dens_dw=np.zeros(n), docc=np.zeros(n), magz=np.zeros(n),
               sz2=np.zeros((n, n), order="F"), n2=np.zeros((n, n), order="F"))
    s2 = C.c_double(0.0)
    gc = g.astype(np.complex128)
    dp = C.POINTER(C.c_double)
    oracle.lib().ora_observables(p.Ns, n, smap.ctypes.data_as(C.POINTER(C.c_uint64)), smap.size,
                                 gc.view(np.float64).ctypes.data_as(dp), 0.5,
                                 *[ref[k].ctypes.data_as(dp) for k in ("dens", "dens_up", "dens_dw", "docc", "magz", "sz2", "n2")],
                                 C.byref(s2))
    s = ctx.sector(3, 2)
    v = s.vec(g)
    got = s.observables(v, peso=0.5)
    for k in ref:
        assert np.abs(got[k] - ref[k]).max() < 1e-12, k
    assert abs(got["s2tot"] - s2.value) < 1e-12
    v.free(); s.free(); ctx.close()
