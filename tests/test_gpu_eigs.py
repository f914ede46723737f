"""GPU (B200): several eigenpairs per sector (edgpu_lanczos_eigs = sp_eigh / ARPACK at ED_DIAG.f90:149-166) and the
reference DEFAULT input lanc_method="arpack", lanc_nstates_sector=6 through ed_solve, against the oracle's eigsh branch
(scipy wraps the same ARPACK dsaupd/dseupd the reference calls through SciFortran)."""
import numpy as np
import pytest

from test_gpu_parity import make
from test_gpu_solver import compare, run_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", [0, 3])
def test_lanczos_eigs_lowest_pairs(oracle, edb, kernel):
    case = dict(Norb=2, Nbath=3, uloc=(2.0, 2.0), ust=1.2, jh=0.2)
    p, model, ctx, rng = make(oracle, edb, case, hxv_kernel=kernel, debug_flags=16 if kernel == 3 else 0)
    for sec in [(4, 4), (3, 5)]:
        smap = oracle.build_sector(p.Ns, *sec)
        H = np.zeros((smap.size, smap.size))
        eye = np.eye(smap.size)
        for j in range(smap.size):
            H[:, j] = oracle.direct_hxv(model, smap, eye[:, j]).real
        w = np.linalg.eigvalsh(H)
        s = ctx.sector(*sec)
        ev, vecs, nconv, nmv = s.lanczos_eigs(6, ncv=60, tol=1e-12)
        assert nconv == 6
        assert np.abs(ev - w[:6]).max() < 1e-10 * max(1.0, np.abs(w[:6]).max())
        V = np.stack([v.download() for v in vecs], axis=1)
        assert np.abs(V.T @ V - np.eye(6)).max() < 1e-10                       # orthonormal, also inside degenerate levels
        assert np.abs(H @ V - V * ev[None, :]).max() < 1e-8
        for v in vecs:
            v.free()
        s.free()
    ctx.close()


def test_ed_solve_reference_default_arpack_six_states(oracle, edb):
    """Orbital-degenerate two-band model away from half filling: the ground state is degenerate INSIDE a
    (nup,ndw) sector; the reference default (arpack, 6 states per sector) keeps every member (ED_DIAG.f90:224-235)."""
    kw = dict(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, xmu=-1.2, lanc_dim_threshold=32,
              lanc_method="arpack", lanc_nstates_sector=6)
    p, ref, sol = run_pair(oracle, edb, **kw)
    states, zeta, egs = sol.states()
    assert zeta == ref.zeta and len(states) == len(ref.states)
    compare(p, ref, sol, tol_obs=1e-8, tol_g=1e-8)
    # the same model with one state per sector would lose members whenever a sector holds a degenerate level
    per_sector = {}
    for st in ref.states:
        per_sector[(st.nup, st.ndw)] = per_sector.get((st.nup, st.ndw), 0) + 1
    sol.close()
    assert max(per_sector.values()) > 1                     # xmu=-1.2: (2,3) and (3,2) each hold a doubly degenerate ground level
    if True:
        kw1 = dict(kw, lanc_method="lanczos", lanc_nstates_sector=1)
        p1, ref1, sol1 = run_pair(oracle, edb, **kw1)
        assert len(sol1.states()[0]) < len(states)
        sol1.close()
