"""All five BASELINE.json configs at their full sizes on one B200: throughput and size-independent checks.

    python profiles/run_configs.py [c1 c2 c3 c4 c5 c5full] > profiles/r1_configs.jsonl

One JSON line per config: rays, interactions, wall time of the analysis (device-resident source, synchronised), the
interactions/s, and the properties that were checked (energy balance, symmetry, monotonic dispersion, ...).  The
per-ray parity of the same scenes against the oracle is in tests/ at reduced ray counts.
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402


def timed(fn, ctx, reps=3):
    fn()
    ctx.jit_wait()  # specialised kernels are compiled in the background: time the steady state
    fn()
    ctx.synchronize()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ctx.synchronize()
        best = min(best, time.perf_counter() - t)
    return out, best


def c1(ctx):
    sc = scenes.to_product(ctx, scenes.c1_single_lens())
    src = scenes.source_to_product(scenes.c1_source(100)).generate(ctx)
    st, dt = timed(lambda: sc.raytrace(src, keep_source=True), ctx)
    fd = sc.node_fluence(3, 100, 83)
    e_det = sc.node_spot(2).total_energy
    # energy balance: two Fresnel surfaces at (near) normal incidence transmit (1-R)^2 with R = ((n-1)/(n+1))^2
    r0 = ((1.5068 - 1) / (1.5068 + 1)) ** 2
    return {"config": "C1 single lens, 10^4 rays", "rays": 10_000, "interactions": st.interactions, "seconds": dt,
            "checks": {"detector_energy": e_det, "normal_incidence_estimate": (1 - r0) ** 2,
                       "energy_within_1e-3_of_estimate": bool(abs(e_det - (1 - r0) ** 2) < 1e-3),
                       "map_energy_equals_hit_energy": bool(abs(fd.total_energy - e_det) < 1e-9 * e_det)}}


def c2(ctx, n=10_000_000):
    sc = scenes.to_product(ctx, scenes.c2_laser_chain())
    src = scenes.source_to_product(scenes.c2_source(n)).generate(ctx)
    st, dt = timed(lambda: sc.raytrace(src, keep_source=True), ctx)
    st_strict = sc.raytrace(src, keep_source=True, strict=True)
    e_main, e_side = sc.node_fluence(9).total_energy, sc.node_fluence(10).total_energy
    st2 = sc.raytrace(src, keep_source=True)
    return {"config": "C2 laser chain, 10^7 rays", "rays": n, "interactions": st.interactions, "seconds": dt,
            "checks": {"repeatable_counts": bool(st2 == st), "strict_build_same_counts": bool(st_strict == st),
                       "valid_out": st.valid_out, "alive_out": st.alive_out,
                       # splitter 50/50, then the main arm loses 2 Fresnel surfaces and a 0.9 filter
                       "side_over_main_energy": e_side / e_main}}


def c3(ctx, n=1_000_000, max_bounces=3):
    sc = scenes.to_product(ctx, scenes.c3_telescope())
    src = scenes.source_to_product(scenes.c3_source(n)).generate(ctx)
    t = time.perf_counter()
    sc.ghostfocus(src, max_bounces)
    ctx.synchronize()
    dt_cold = time.perf_counter() - t  # first analysis on a fresh scene: every bundle / hit map is a cudaMalloc
    st, dt = timed(lambda: sc.ghostfocus(src, max_bounces), ctx, reps=2)  # buffers recycled from the scene's pool
    per_pass = []
    for p in range(st.passes):
        counts, energy, nb, nr = sc.gf_pass_tally(p, 6)
        per_pass.append({"bundles": nb, "rays": nr, "valid_per_bounce": counts.tolist(), "energy_per_bounce": energy.tolist()})
    # pass 0: the seed bundle crosses 2 x ConstantR(0.05) and 2 x Fresnel surfaces; everything that is not transmitted
    # went into the four reflected bundles, which re-enter in pass 1 -> order-1 energy of pass 1 <= 1 - transmitted
    e0 = per_pass[0]["energy_per_bounce"][0]
    return {"config": f"C3 ghost focus, 10^6 seeds, bounce orders 0..{max_bounces}", "rays": n, "interactions": st.interactions,
            "seconds": dt, "seconds_first_analysis": dt_cold, "launches": st.launches, "passes": per_pass,
            "checks": {"seed_bundle_transmission": e0, "transmission_upper_bound_0.95^2": 0.95 ** 2,
                       "transmitted_below_bound": bool(e0 < 0.95 ** 2), "all_seed_rays_valid": per_pass[0]["valid_per_bounce"][0] == n}}


def c4(ctx, side=10_000, nx=4096, ny=4096):
    sc = scenes.to_product(ctx, scenes.c4_fluence())
    src = scenes.source_to_product(scenes.c4_source(side)).generate(ctx)
    st, dt = timed(lambda: sc.raytrace(src, keep_source=True, snapshots=False), ctx, reps=2)
    t = time.perf_counter()
    fd = sc.node_fluence(2, nx, ny)
    dt_map = time.perf_counter() - t
    m = fd.map
    sym_x = float(np.max(np.abs(m - m[:, ::-1])) / fd.peak)
    sym_y = float(np.max(np.abs(m - m[::-1, :])) / fd.peak)
    hits = sc.node_hitmaps(2, 0)
    return {"config": "C4 4096x4096 fluence map, 10^8 rays", "rays": side * side, "interactions": st.interactions, "seconds": dt,
            "map_seconds_incl_download": dt_map,
            "checks": {"total_energy_J": fd.total_energy, "peak_J_m2": fd.peak, "mirror_symmetry_x_rel": sym_x, "mirror_symmetry_y_rel": sym_y,
                       "symmetric_within_1e-9": bool(max(sym_x, sym_y) < 1e-9), "hit_maps": len(hits)}}


def c5(ctx, n_pos, n_wvl=64, chunk_pos=1_562_500, nx=512, ny=512):
    """streamed: positions are processed in chunks of whole wavelength groups; pass 1 finds the map range, pass 2 bins"""
    import ctypes as C

    import torch

    from opossum_b200 import _ffi as F
    sc = scenes.to_product(ctx, scenes.c5_broadband())
    srcd = scenes.c5_source(n_pos, n_wvl)
    sd = scenes.source_to_product(srcd)
    rays = ob.Rays(ctx, min(chunk_pos, n_pos) * n_wvl)
    dev = torch.device("cuda", 0)
    rng = torch.full((4,), -float("inf"), dtype=torch.float64, device=dev)
    tmp = torch.zeros(4, dtype=torch.float64, device=dev)
    fmap = torch.zeros(nx * ny, dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L = F.lib()
    inter = 0
    t0 = time.perf_counter()
    for phase in (1, 2):
        for b in range(0, n_pos, chunk_pos):
            e = min(b + chunk_pos, n_pos)
            sd.generate(ctx, b, e, rays=rays)
            st = sc.raytrace(rays, snapshots=False)
            inter += st.interactions if phase == 1 else 0
            hm = sc.node_hitmaps(4, 0)
            arr = (C.c_void_p * len(hm))(*[h.h for h in hm])
            if phase == 1:
                ob.api._check(L.opm_hitmaps_bounding_box_async(arr, len(hm), -1, C.c_void_p(tmp.data_ptr())))
                ctx.synchronize()
                torch.maximum(rng, tmp, out=rng)
                torch.cuda.synchronize()
            else:
                ob.api._check(L.opm_fluence_binning_async(arr, len(hm), -1, nx, ny, C.c_void_p(rng.data_ptr()), int(b > 0),
                                                          C.c_void_p(fmap.data_ptr()), C.c_void_p(status.data_ptr())))
    ctx.synchronize()
    dt = time.perf_counter() - t0
    r = rng.tolist()
    lrtb = (-r[0], r[1], r[2], -r[3])
    fd = ob.fluence_peak_total(ctx, fmap.data_ptr(), nx, ny, lrtb)
    wv, w = scenes.spectrum_of(srcd)
    return {"config": f"C5 broadband, {n_pos} positions x {n_wvl} wavelengths, streamed in chunks of {chunk_pos} positions, two passes "
                      "(range, then Binning)", "rays": n_pos * n_wvl, "interactions": inter, "seconds": dt, "traced_interactions": 2 * inter,
            "checks": {"status_bits": int(status.item()), "map_total_energy_J": fd.total_energy, "source_energy_J": float(np.sum(w)),
                       "map_energy_below_source": bool(fd.total_energy < float(np.sum(w))), "lrtb": lrtb}}


def kde(ctx, n=20000, nx=100, ny=83):
    """the KDE fluence estimator (kde/mod.rs) on the C1 detector's hit map: O(n^2) bandwidth estimate
    (2*10^8 pair distances, sorted) + n x pixels Gaussian sum, against the oracle on the host cores"""
    from oracle import oracle as orc
    side = int(round(n ** 0.5))
    sc = scenes.to_product(ctx, scenes.c1_single_lens())
    sc.raytrace(scenes.source_to_product(scenes.c1_source(side)).generate(ctx))
    fd, dt = timed(lambda: sc.node_fluence(3, nx, ny, estimator="kde"), ctx, reps=2)
    x, y, z, e, b = sc.node_hitmaps(3, 0)[0].download()
    hits = orc.hits_from_xye(x, y, e)
    t = time.perf_counter()
    ref_map, _, ref_bw = orc.fluence_kde(hits, nx, ny, n_threads=orc.max_threads())
    dt_cpu = time.perf_counter() - t
    err = float(np.max(np.abs(fd.map - ref_map)) / ref_map.max())
    return {"config": f"KDE fluence estimator, {len(x)} hit points -> {nx}x{ny} map", "rays": len(x), "interactions": len(x) * nx * ny,
            "seconds": dt, "cpu_oracle_seconds": dt_cpu, "cpu_threads": orc.max_threads(),
            "checks": {"band_width": fd.band_width, "band_width_oracle": ref_bw, "max_abs_err_over_peak": err, "within_1e-9": bool(err < 1e-9)}}


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    ctx = ob.Context(0)
    table = {"c1": lambda: c1(ctx), "c2": lambda: c2(ctx), "c3": lambda: c3(ctx), "c4": lambda: c4(ctx),
             "c5": lambda: c5(ctx, 1_562_500), "c5full": lambda: c5(ctx, 100_000_000), "kde": lambda: kde(ctx)}
    for k in which:
        out = table[k]()
        out["interactions_per_s"] = out.get("traced_interactions", out["interactions"]) / out["seconds"]
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
