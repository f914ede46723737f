"""Golden fixtures (tests/golden/oracle_golden.npz, written by tests/golden/make_golden.py).

CPU part: the oracle still reproduces its frozen outputs and the survey anchors.
GPU part: the CUDA path reproduces the same fixtures through the C-ABI (no CPU code involved at run time)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))


def lanczos_params(O, **kw):
    base = dict(lanc_method="lanczos", lanc_nstates_sector=1)
    base.update(kw)
    return O.Params(**base)


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_reproduces_golden_maps(oracle):
    offs = G["maps_ns5_offsets"]
    k = 0
    for nup in range(6):
        for ndw in range(6):
            assert np.array_equal(oracle.build_sector(5, nup, ndw), G["maps_ns5"][offs[k]:offs[k + 1]])
            k += 1


def test_oracle_reproduces_golden_hxv(oracle):
    p = lanczos_params(oracle, Norb=1, Nbath=4)
    m = oracle.Model(p, oracle.init_bath(p))
    out = oracle.direct_hxv(m, oracle.build_sector(5, 2, 3), G["hxv_cfg1_23_in"])
    assert np.array_equal(out, G["hxv_cfg1_23_out"])
    p2 = lanczos_params(oracle, Norb=2, Nbath=2, uloc=(2.0, 1.5), ust=1.2, jh=0.3, jx=0.2, jp=0.1)
    m2 = oracle.Model(p2, G["hxv_2orb_bath"])
    out2 = oracle.direct_hxv(m2, oracle.build_sector(6, 3, 3), G["hxv_2orb_33_in"])
    assert np.array_equal(out2, G["hxv_2orb_33_out"])


def test_golden_agrees_with_survey_anchors():
    assert abs(G["cfg1_egs"][0] - G["anchor_cfg1_e0"][0]) < 1e-10
    assert abs(G["cfg2_egs"][0] - G["anchor_cfg2_e012"][0]) < 1e-10
    assert abs(G["cfg2_dens_docc"][0] - G["anchor_cfg2_dens_docc"][0]) < 1e-6     # plain-Lanczos Ritz vector accuracy
    assert abs(G["cfg2_dens_docc"][1] - G["anchor_cfg2_dens_docc"][1]) < 1e-6
    assert abs(G["cfg1_dens_docc"][0] - 1.0) < 1e-10


def test_oracle_reproduces_golden_solve_cfg1(oracle):
    p = oracle.Params(Norb=1, Nbath=4, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0)
    r = oracle.ed_solve(p, oracle.init_bath(p))
    assert abs(r.egs - G["cfg1_egs"][0]) < 1e-12
    assert np.abs(r.impGmats[0, 0, 0, 0] - G["cfg1_gmats"]).max() < 1e-11
    assert np.abs(r.impSmats[0, 0, 0, 0] - G["cfg1_smats"]).max() < 1e-9


# ------------------------------------------------------------------------------------------------ GPU: product
@pytest.mark.gpu
@pytest.mark.parametrize("layout", [1, 2])
def test_gpu_reproduces_golden_maps_and_hxv(edb, oracle, layout):
    bath = oracle.init_bath(lanczos_params(oracle, Norb=1, Nbath=4))
    ctx = edb.Context(1, 4, layout=layout, hxv_kernel=layout)
    ctx.set_hamiltonian(bath, (2.0,))
    offs = G["maps_ns5_offsets"]
    k = 0
    for nup in range(6):
        for ndw in range(6):
            s = ctx.sector(nup, ndw)
            assert np.array_equal(s.map(), G["maps_ns5"][offs[k]:offs[k + 1]])      # bit-exact
            s.free()
            k += 1
    s = ctx.sector(2, 3)
    assert np.abs(s.hxv_host(G["hxv_cfg1_23_in"]) - G["hxv_cfg1_23_out"]).max() < 1e-12
    s.free()
    ctx.close()
    ctx = edb.Context(2, 2, layout=1)
    ctx.set_hamiltonian(G["hxv_2orb_bath"], (2.0, 1.5), 1.2, 0.3, 0.2, 0.1)
    s = ctx.sector(3, 3)
    assert np.abs(s.hxv_host(G["hxv_2orb_33_in"]) - G["hxv_2orb_33_out"]).max() < 1e-12
    s.free()
    ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_cfg2_probe(edb, oracle):
    bath = oracle.init_bath(lanczos_params(oracle, Norb=1, Nbath=9))
    for layout in (1, 2):
        ctx = edb.Context(1, 9, layout=layout, hxv_kernel=layout)
        ctx.set_hamiltonian(bath, (2.0,))
        s = ctx.sector(5, 5)
        x, y = s.vec().fill_normal(20240607), s.vec()
        s.hxv(x, y)
        hv = y.download()
        assert np.abs(hv[G["hxv_cfg2_55_probe_idx"]] - G["hxv_cfg2_55_probe_out"]).max() < 1e-11
        assert abs(np.linalg.norm(hv) - G["hxv_cfg2_55_norm"][0]) < 1e-9
        ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_solves(edb):
    inp = edb.default_input(Norb=1, Nbath=4, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0,
                            ed_sparse_H=0)
    sol = edb.Solver(inp)
    sol.solve()
    _, _, egs = sol.states()
    assert abs(egs - G["cfg1_egs"][0]) < 1e-10 * abs(egs)
    assert abs(sol.dens()[0] - G["cfg1_dens_docc"][0]) < 1e-9 and abs(sol.docc()[0] - G["cfg1_dens_docc"][1]) < 1e-9
    assert np.abs(sol.gimp_matsubara()[0, 0, 0, 0] - G["cfg1_gmats"]).max() < 1e-8
    assert np.abs(sol.sigma_matsubara()[0, 0, 0, 0] - G["cfg1_smats"]).max() < 1e-8
    sol.close()
    inp = edb.default_input(Norb=1, Nbath=9, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0,
                            lanc_ngfiter=60, ed_sparse_H=0)
    sol = edb.Solver(inp)
    sol.set_sectors([(5, 5), (4, 5), (5, 4), (6, 5), (5, 6), (4, 4), (6, 6)])
    sol.solve()
    _, _, egs = sol.states()
    assert abs(egs - G["cfg2_egs"][0]) < 1e-10 * abs(egs)
    assert abs(sol.dens()[0] - G["cfg2_dens_docc"][0]) < 1e-9 and abs(sol.docc()[0] - G["cfg2_dens_docc"][1]) < 1e-9
    assert np.abs(sol.gimp_matsubara()[0, 0, 0, 0] - G["cfg2_gmats"]).max() < 1e-8
    assert np.abs(sol.sigma_matsubara()[0, 0, 0, 0] - G["cfg2_smats"]).max() < 1e-8
    ch = sol.chains()[0]
    assert np.abs(ch["alfa"][:5] - G["cfg2_chain0_alfa"][:5]).max() < 1e-9
    assert np.abs(ch["beta"][:5] - G["cfg2_chain0_beta"][:5]).max() < 1e-9
    sol.close()


# ------------------------------------------------------------------------------------------------ spin susceptibility
def _chi_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_chi", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                                                                  "make_golden_chi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return (("cfg1", m.CFG1), ("two", m.TWO))


def test_oracle_reproduces_golden_chi(oracle):
    GC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden_chi.npz"))
    for name, kw in _chi_cases():
        p = oracle.Params(**kw)
        r = oracle.ed_solve(p, oracle.init_bath(p))
        assert np.abs(r.spinChi_tau - GC[name + "_chi_tau"]).max() < 1e-10
        assert np.abs(r.spinChi_iv - GC[name + "_chi_iv"]).max() < 1e-9 * max(1.0, np.abs(GC[name + "_chi_iv"]).max())
        assert np.abs(r.densChi_tau - GC[name + "_dchi_tau"]).max() < 1e-9 * max(1.0, np.abs(GC[name + "_dchi_tau"]).max())


@pytest.mark.gpu
def test_gpu_reproduces_golden_chi(edb):
    GC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden_chi.npz"))
    for name, kw in _chi_cases():
        kw = dict(kw)
        kw["chispin_flag"] = 1
        kw["chidens_flag"] = 1
        inp = edb.default_input(ed_sparse_H=0, **kw)
        sol = edb.Solver(inp)
        sol.solve()
        iv, ct, cw, vm, tau = sol.spinchi()
        assert np.abs(ct - GC[name + "_chi_tau"]).max() < 1e-8
        assert np.abs(iv - GC[name + "_chi_iv"]).max() < 1e-8 * max(1.0, np.abs(GC[name + "_chi_iv"]).max())
        div, dct, dcw, tiv, tt, tw = sol.denschi()
        assert np.abs(dct - GC[name + "_dchi_tau"]).max() < 1e-8 * max(1.0, np.abs(GC[name + "_dchi_tau"]).max())
        assert np.abs(tt - GC[name + "_dchi_tot_tau"]).max() < 1e-8 * max(1.0, np.abs(GC[name + "_dchi_tot_tau"]).max())
        sol.close()
