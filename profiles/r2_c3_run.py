"""C3 (BASELINE configs[2]): ghost-focus analysis of the two-lens telescope, 10^6 seed rays, bounce orders 0-3; wall time of
N analyses (the first ones allocate the bundle pool and compile the specialised kernels), interactions and launches per analysis."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/profiles')
import opossum_b200 as ob, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ctx = ob.Context(0)
sc = scenes.to_product(ctx, scenes.c3_telescope())
src = scenes.source_to_product(scenes.c3_source(1_000_000)).generate(ctx)
ts = []
for k in range(N):
    if k == 3:
        ctx.jit_wait()  # the specialised kernels of the analysis' programs are compiled in the background: measure them, not the hand-over
        sc.ghostfocus(src, 3); ctx.synchronize()
    t = time.perf_counter(); st = sc.ghostfocus(src, 3); ctx.synchronize(); ts.append(time.perf_counter() - t)
warm = sorted(ts[3:])
print("ms per analysis (after 3 warm-up): min %.3f median %.3f max %.3f | first three %s | interactions %d launches %d -> %.3e interactions/s (median)"
      % (warm[0] * 1e3, warm[len(warm) // 2] * 1e3, warm[-1] * 1e3, ["%.1f" % (t * 1e3) for t in ts[:3]], st.interactions, st.launches,
         st.interactions / warm[len(warm) // 2]))
