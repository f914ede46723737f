"""H*v time of every large sector of a workload's model for the star kernels (hxv_kernel=2), the fiber kernels (3) and the
sector-build time: finds where the automatic choice should switch.  usage: probe_sectors.py cfg3 [min_dim]"""
import importlib
import sys
import time

sys.path.insert(0, ".")
import bench as B  # noqa: E402


def main():
    wl = sys.argv[1]
    min_dim = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 19
    edb = importlib.import_module("dmft-ed_b200")
    Norb, Nbath, _, _, _ = B.WORKLOADS[wl]
    Ns = Norb * (Nbath + 1)
    import math
    secs = [(a, b) for a in range(Ns + 1) for b in range(a, Ns + 1) if math.comb(Ns, a) * math.comb(Ns, b) >= min_dim]
    ctxs = {k: B.make_model_ctx(edb, wl, 0, None, kernel=k)[0] for k in (2, 3)}
    print("sector dim  build_ms(star,fiber)  hxv_ms(star,fiber)")
    for sec in secs:
        row = []
        for k in (2, 3):
            t0 = time.perf_counter()
            s = ctxs[k].sector(*sec)
            x, y = s.vec().fill_uniform(1), s.vec()
            s.hxv(x, y)
            ctxs[k].sync()
            tb = (time.perf_counter() - t0) * 1e3
            ms, _ = s.bench_hxv(x, y, 10)
            row.append((tb, ms, s.dim))
            x.free(); y.free(); s.free()
        print(f"{sec} {row[0][2]:>10d}  {row[0][0]:8.2f} {row[1][0]:8.2f}   {row[0][1]:8.4f} {row[1][1]:8.4f}", flush=True)


main()
