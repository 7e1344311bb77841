import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/profiles')
import opossum_b200 as ob, scenes
ctx = ob.Context(0)
sc = scenes.to_product(ctx, scenes.c3_telescope())
src = scenes.source_to_product(scenes.c3_source(1_000_000)).generate(ctx)
for k in range(3):
    t = time.perf_counter(); st = sc.ghostfocus(src, 3); ctx.synchronize(); print(k, time.perf_counter() - t, st.interactions, st.launches)
