import sys, importlib
sys.path.insert(0, '.')
import numpy as np
edb = importlib.import_module("dmft-ed_b200")
from oracle import ed_oracle as O
p = O.Params(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, lanc_method="lanczos", lanc_nstates_sector=1)
bath = O.init_bath(p); model = O.Model(p, bath)
ctx = edb.Context(2, 2); ctx.set_hamiltonian(bath, p.uloc, p.ust, p.jh)
for sec in [(4, 2), (3, 3), (2, 4)]:
    smap = O.build_sector(6, *sec)
    H = O.dense_h(model, smap).real
    w, Z = np.linalg.eigh(H)
    v0 = O.start_vector(smap.size)
    e_ref, vec_ref, nl_ref, a_ref, b_ref = O.lanc_gs(model, smap, v0, min(512, smap.size), 1e-12)
    s = ctx.sector(*sec)
    v = s.vec().fill_uniform(1234567)
    assert np.array_equal(v.download(), v0.real)
    e0, nl, a, b = s.lanczos_gs(v, min(512, smap.size), 1e-12)
    gs = v.download()
    print(sec, "dim", smap.size, "exact", w[:3], "e", e0, e_ref, "nl", nl, nl_ref)
    print("  a diff max", np.abs(a[:min(nl, nl_ref)] - a_ref[:min(nl, nl_ref)]).max(), "b diff", np.abs(b[:min(nl, nl_ref)] - b_ref[:min(nl, nl_ref)]).max())
    sg = np.sign(gs @ vec_ref.real)
    print("  |dev-ora|", np.abs(gs * sg - vec_ref.real).max(), "|dev-exact|", min(np.abs(gs - Z[:, 0]).max(), np.abs(gs + Z[:, 0]).max()),
          "|ora-exact|", min(np.abs(vec_ref.real - Z[:, 0]).max(), np.abs(vec_ref.real + Z[:, 0]).max()))
print("---- CSR path")
for sec in [(4, 2), (2, 4)]:
    smap = O.build_sector(6, *sec)
    v0 = O.start_vector(smap.size)
    e_ref, vec_ref, nl_ref, a_ref, b_ref = O.lanc_gs(model, smap, v0, min(512, smap.size), 1e-12)
    s = ctx.sector(*sec); s.build_csr()
    v = s.vec().fill_uniform(1234567)
    e0, nl, a, b = s.lanczos_gs(v, min(512, smap.size), 1e-12)
    gs = v.download(); sg = np.sign(gs @ vec_ref.real)
    print(sec, "nl", nl, nl_ref, "a diff", np.abs(a[:min(nl, nl_ref)] - a_ref[:min(nl, nl_ref)]).max(), "|dev-ora|", np.abs(gs * sg - vec_ref.real).max())
print("---- solver, sparse 0/1")
ref = O.ed_solve(O.Params(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, lanc_dim_threshold=64, beta=100.), bath)
print("ref states", [(s.e, s.nup, s.ndw) for s in ref.states], ref.dens, ref.docc)
for sparse in (0, 1):
    inp = edb.default_input(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, lanc_dim_threshold=64, beta=100., ed_sparse_H=sparse)
    sol = edb.Solver(inp); sol.solve()
    print(sparse, sol.states(), sol.dens(), sol.docc(), np.abs(sol.gimp_matsubara() - ref.impGmats).max())
    for i, st in enumerate(ref.states):
        import ctypes as C
        vec = np.zeros(st.vec.size)
        edb.lib().ed_get_state_vector(sol.h, i, vec.ctypes.data_as(edb.dp), vec.size)
        print("   state", i, np.abs(np.abs(vec) - np.abs(st.vec.real)).max())
    sol.close()
