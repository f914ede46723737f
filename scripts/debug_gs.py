import sys, importlib
sys.path.insert(0, '.')
import numpy as np
edb = importlib.import_module("dmft-ed_b200")
from oracle import ed_oracle as O
p = O.Params(Norb=1, Nbath=9, lanc_method="lanczos", lanc_nstates_sector=1)
bath = O.init_bath(p); model = O.Model(p, bath)
ctx = edb.Context(1, 9); ctx.set_hamiltonian(bath, p.uloc)
smap = O.build_sector(10, 5, 5)
v0 = O.start_vector(smap.size)
e_ref, vec_ref, nl_ref, a_ref, b_ref = O.lanc_gs(model, smap, v0, 512, 1e-12)
s = ctx.sector(5, 5)
v = s.vec(v0.real)
e0, nl, a, b = s.lanczos_gs(v, 512, 1e-12)
gs = v.download()
print("e", e0, e_ref, "nl", nl, nl_ref)
print("a diff", np.abs(a[:min(nl,nl_ref)] - a_ref[:min(nl,nl_ref)]))
print("overlap", gs @ vec_ref.real, np.linalg.norm(gs))
r = O.direct_hxv(model, smap, gs).real - e0 * gs
print("resid dev", np.abs(r).max())
r = O.direct_hxv(model, smap, vec_ref).real - e_ref * vec_ref.real
print("resid ora", np.abs(r).max())
