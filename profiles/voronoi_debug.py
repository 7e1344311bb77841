import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import opossum_b200 as ob
from oracle import oracle as orc
ctx = ob.Context(0)
def run(x, y, e, nx, tag):
    ref, _ = orc.fluence_voronoi(orc.hits_from_xye(x, y, e), nx)
    h = ob.HitMap(ctx, len(x)); h.upload(x, y, e)
    got = ob.fluence_voronoi(ctx, [h], nx).map
    gn, rn = np.isnan(got), np.isnan(ref)
    ok = ~rn & ~gn
    err = np.abs(got[ok] - ref[ok]) / np.abs(ref[ok])
    print(tag, "nan mismatch", int((gn != rn).sum()), "finite", int(ok.sum()), "max err", err.max(), "median", np.median(err), "frac>1e-9", float((err > 1e-9).mean()))
    bad = np.argwhere(ok & (np.abs(got - ref) > 1e-6 * np.abs(ref)))
    print("   bad rows(y idx) range", bad[:, 0].min() if len(bad) else None, bad[:, 0].max() if len(bad) else None, "cols", bad[:, 1].min() if len(bad) else None, bad[:, 1].max() if len(bad) else None)
rng = np.random.default_rng(11)
n = 2000
e = 1e-6 * (1 + rng.random(n))
x, y = 3e-3 * rng.standard_normal(n), 0.2e-3 * rng.standard_normal(n)
run(x, y, e, 50, "aniso 15:1")
run(x, y * 15, e, 50, "iso gaussian")
run(x, y * 5, e, 50, "aniso 3:1")
u = rng.random(n); v = rng.random(n)
run(u * 6e-3, v * 0.4e-3, e, 50, "uniform box 15:1")
run(u * 6e-3, v * 6e-3, e, 50, "uniform box 1:1")
