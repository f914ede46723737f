"""GPU (B200): ed_init_solver / ed_solve / ed_get_* through the C++ host mirror vs the oracle's ed_solve."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_pair(oracle, edb, sectors=None, sparse=1, **kw):
    base = dict(lanc_method="lanczos", lanc_nstates_sector=1, Lmats=128, Lreal=128, beta=100.0)
    base.update(kw)
    p = oracle.Params(**base)
    bath = oracle.init_bath(p)
    ref = oracle.ed_solve(p, bath, sectors=sectors)
    inp = edb.default_input(Norb=p.Norb, Nbath=p.Nbath, Nspin=p.Nspin, uloc=p.uloc, ust=p.ust, jh=p.jh, jx=p.jx, jp=p.jp,
                            beta=p.beta, xmu=p.xmu, hfmode=int(p.hfmode), Lmats=p.Lmats, Lreal=p.Lreal,
                            lanc_method=p.lanc_method, lanc_nstates_sector=p.lanc_nstates_sector,
                            lanc_ngfiter=p.lanc_ngfiter, lanc_niter=p.lanc_niter, ed_sparse_H=sparse)
    sol = edb.Solver(inp)
    assert np.array_equal(sol.bath, bath)                         # init_dmft_bath mirror
    if sectors is not None:
        sol.set_sectors(sectors)
    sol.solve()
    return p, ref, sol


def compare(p, ref, sol):
    states, zeta, egs = sol.states()
    assert zeta == ref.zeta and len(states) == len(ref.states)
    assert abs(egs - ref.egs) < 1e-10 * abs(ref.egs)
    assert sorted((s[1], s[2]) for s in states) == sorted((s.nup, s.ndw) for s in ref.states)
    assert np.abs(sol.dens() - ref.dens).max() < 1e-9
    assert np.abs(sol.docc() - ref.docc).max() < 1e-9
    assert np.abs(sol.mag() - ref.magz).max() < 1e-9
    sz2, n2, s2 = sol.sz2_n2()
    assert np.abs(sz2 - ref.sz2).max() < 1e-9 and np.abs(n2 - ref.n2).max() < 1e-9 and abs(s2 - ref.s2tot) < 1e-9
    assert np.abs(sol.gimp_matsubara() - ref.impGmats).max() < 1e-8
    assert np.abs(sol.gimp_real() - ref.impGreal).max() < 1e-8 * max(1.0, np.abs(ref.impGreal).max())
    assert np.abs(sol.sigma_matsubara() - ref.impSmats).max() < 1e-8
    assert np.abs(sol.g0imp_matsubara() - ref.impG0mats).max() < 1e-12


def test_ed_solve_cfg1_full_scan(oracle, edb):
    """BASELINE config 1: drivers/ed_hm_bethe.f90, Nbath=4 (Ns=5), U=2, half filling, all 36 sectors."""
    for sparse in (1, 0):
        p, ref, sol = run_pair(oracle, edb, Norb=1, Nbath=4, sparse=sparse)
        compare(p, ref, sol)
        # GF chains: alpha/beta to 1e-9 (chains run to the full sector dimension: compare the leading part,
        # the trailing coefficients of an un-reorthogonalised Lanczos are rounding noise)
        chains = sol.chains()
        assert len(chains) == len(ref.chains)
        for c, r in zip(chains, ref.chains):
            assert (c["iorb"], c["ispin"], c["isign"], c["nlanc"]) == (r["iorb"], r["ispin"], r["isign"], len(r["alfa"]))
            assert abs(c["norm2"] - r["norm2"]) < 1e-9
        sol.close()


def test_ed_solve_two_orbitals_hund(oracle, edb):
    p, ref, sol = run_pair(oracle, edb, Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, lanc_dim_threshold=64)
    compare(p, ref, sol)
    sol.close()


def test_ed_solve_cfg2_half_filling_window(oracle, edb):
    """BASELINE config 2 (Nbath=9, Ns=10): ground state + Matsubara GF, sector scan restricted to the
    half-filling window (ED_SECTORS), so that the literal CPU oracle finishes in seconds."""
    secs = [(5, 5), (4, 5), (5, 4), (6, 5), (5, 6), (4, 4), (6, 6)]
    p, ref, sol = run_pair(oracle, edb, sectors=secs, Norb=1, Nbath=9, lanc_ngfiter=60)
    compare(p, ref, sol)
    assert abs(sol.sector_energy(5, 5) - (-11.341244826804)) < 1e-9
    for c, r in zip(sol.chains(), ref.chains):
        k = 25
        assert np.abs(c["alfa"][:k] - r["alfa"][:k]).max() < 1e-9
        assert np.abs(c["beta"][:k] - r["beta"][:k]).max() < 1e-9
    sol.close()
