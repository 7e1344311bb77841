"""Diagnostic: where does the e2e step go?  (upload only / trace only / both, wall clock per step)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import opossum_b200 as ob, scenes
n = 10_000_000
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = ob.Context(0, stream=stream.cuda_stream)
    scene = scenes.to_product(ctx, scenes.c2_laser_chain())
    src = scenes.source_to_product(scenes.c2_source(n)).generate(ctx)
    pos_h = torch.empty((n, 3), dtype=torch.float64).pin_memory(); e_h = torch.empty(n, dtype=torch.float64).pin_memory()
    tmp = np.empty(n)
    for k, f in enumerate(("px", "py", "pz")):
        ctx.to_host(src.field_ptr(f), tmp); pos_h[:, k] = torch.from_numpy(tmp)
    ctx.to_host(src.field_ptr("e"), tmp); e_h.copy_(torch.from_numpy(tmp))
    bufs = [ob.Rays(ctx, n), ob.Rays(ctx, n)]
    wvl, w = [1054e-9], [1.0]
    def upload(r): r.new_collimated_with_spectrum(pos_h.data_ptr(), e_h.data_ptr(), n, wvl, w)
    def timeit(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
    k = [0]
    def up_only():
        r = bufs[k[0] % 2]; k[0] += 1; upload(r); len(r); r.spot_stats()
    print("upload + expand + sync (serial) ms", timeit(up_only))
    def up_async_only():
        r = bufs[k[0] % 2]; k[0] += 1; upload(r)
    print("upload enqueue only (host) ms", timeit(up_async_only))
    def trace_only():
        scene.raytrace(src, keep_source=True, sync=False); scene.node_fluence(9, download=False); scene.node_fluence(10, download=False); scene.node_spot(8)
    print("trace+readout ms", timeit(trace_only))
    upload(bufs[0])
    def both():
        cur, nxt = bufs[k[0] % 2], bufs[(k[0] + 1) % 2]; k[0] += 1
        t0 = time.perf_counter(); upload(nxt); t1 = time.perf_counter()
        scene.raytrace(cur, sync=False); t2 = time.perf_counter()
        scene.node_fluence(9, download=False); scene.node_fluence(10, download=False); scene.node_spot(8); t3 = time.perf_counter()
        both.t = [both.t[0] + t1 - t0, both.t[1] + t2 - t1, both.t[2] + t3 - t2]
    both.t = [0, 0, 0]
    k[0] = 0
    print("pipelined ms", timeit(both), "host split (upload, raytrace, readout) ms per step", [x / 13 * 1e3 for x in both.t])
    ctx.close()
