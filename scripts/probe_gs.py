"""Ground-state Lanczos time per sector, star (2) vs fiber (3) kernels.  usage: probe_gs.py cfg3"""
import importlib
import sys
import time

sys.path.insert(0, ".")
import bench as B  # noqa: E402


def main():
    wl = sys.argv[1]
    edb = importlib.import_module("dmft-ed_b200")
    ctxs = {k: B.make_model_ctx(edb, wl, 0, None, kernel=k)[0] for k in (2, 3)}
    for sec in [(7, 7), (5, 6), (4, 4), (3, 6), (7, 9)]:
        for k in (2, 3):
            t0 = time.perf_counter()
            s = ctxs[k].sector(*sec)
            v = s.vec().fill_uniform(1234567)
            ctxs[k].sync()
            t1 = time.perf_counter()
            e0, nl, _, _ = s.lanczos_gs(v, 512)
            ctxs[k].sync()
            t2 = time.perf_counter()
            a, b, nu = s.lanczos_tridiag(v, 200)
            ctxs[k].sync()
            t3 = time.perf_counter()
            v.free(); s.free()
            t4 = time.perf_counter()
            print(f"{sec} kernel={k} dim={s.dim} build {1e3*(t1-t0):7.1f} ms  gs {1e3*(t2-t1):8.1f} ms ({nl} steps, {1e3*(t2-t1)/max(nl,1)/2:.3f} ms/step)  tridiag200 {1e3*(t3-t2):7.1f} ms  free {1e3*(t4-t3):6.1f} ms  e0={e0:.6f}", flush=True)


main()
