"""Large-map Binning kernels on the C4 hit map (BASELINE configs[3]: 10^8 hits -> 4096 x 4096): the warp-segmented kernel of
round 1 against the run-merging kernel in several launch shapes.  One trace, then every variant bins the SAME hit map into the
same grid; maps are compared bin by bin with the first variant.

    python profiles/c4_binning_probe.py [side] [nx]      # side^2 rays; prints one JSON line per variant
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200 import _ffi as F  # noqa: E402
from opossum_b200.api import _check  # noqa: E402

SIDE = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
NX = NY = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
REPS = 5


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    L = F.lib()
    with torch.cuda.stream(stream):
        ctx = ob.Context(0, stream=stream.cuda_stream)
        scene = scenes.to_product(ctx, scenes.c4_fluence())
        src = scenes.source_to_product(scenes.c4_source(SIDE)).generate(ctx)
        st = scene.raytrace(src, keep_source=True, snapshots=False)
        ctx.jit_wait()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
        e0.record(stream)
        for _ in range(REPS):
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
        e1.record(stream)
        torch.cuda.synchronize()
        trace_ms = e0.elapsed_time(e1) / REPS
        hm = scene.node_hitmaps(2, 0)
        arr = (C.c_void_p * len(hm))(*[h.h for h in hm])
        rng = torch.zeros(4, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        pk = torch.zeros(2, dtype=torch.float64, device=dev)
        vp = C.c_void_p
        _check(L.opm_hitmaps_bounding_box_async(arr, len(hm), -1, vp(rng.data_ptr())))
        torch.cuda.synchronize()
        print(json.dumps({"rays": SIDE * SIDE, "map": [NX, NY], "hits": st.hits, "trace_ms": trace_ms, "range": rng.tolist()}), flush=True)
        variants = [("warp", None)]
        if os.environ.get("PROBE_VARIANTS"):  # e.g. "5,0,1;8,0,1" (shape, CTAs per SM (0 = the shape's own), interleave)
            variants += [("runs", v) for v in os.environ["PROBE_VARIANTS"].split(";")]
        else:
            for shape in (0, 6, 8, 9, 10, 11, 12, 20, 21, 22, 23, 24, 25, 26, 27):
                variants.append(("runs", f"{shape},0,1"))
        ref = None
        for which, cfg in variants:
            os.environ["OPM_B200_BINNING_LARGE"] = which
            if cfg is not None:
                os.environ["OPM_B200_BINNING_RUNS"] = cfg
            m = torch.zeros(NX * NY, dtype=torch.float64, device=dev)

            def once():
                _check(L.opm_fluence_binning_async(arr, len(hm), -1, NX, NY, vp(rng.data_ptr()), 0, vp(m.data_ptr()), vp(status.data_ptr())))

            once()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(REPS):
                once()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / REPS
            out = {"kernel": which, "cfg": cfg, "zero_plus_bin_ms": ms, "hits_per_s": st.hits / (ms * 1e-3),
                   "gbs_28B_per_slot": SIDE * SIDE * 28 / (ms * 1e-3) / 1e9, "status": int(status.item())}
            if ref is None:
                ref = m.clone()
                out["total"] = float(m.sum())
            else:
                nz = ref > 1e-12 * ref.max()
                out["max_rel_vs_warp"] = float(((m - ref).abs()[nz] / ref[nz]).max())
                out["max_abs_below_floor"] = float((m - ref).abs()[~nz].max()) if (~nz).any() else 0.0
            print(json.dumps(out), flush=True)
            del m
        # peak / total of the last map
        m = ref
        _check(L.opm_fluence_peak_total_async(ctx.h, vp(m.data_ptr()), NX, NY, vp(rng.data_ptr()), vp(pk.data_ptr())))
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(REPS):
            _check(L.opm_fluence_peak_total_async(ctx.h, vp(m.data_ptr()), NX, NY, vp(rng.data_ptr()), vp(pk.data_ptr())))
        e1.record(stream)
        torch.cuda.synchronize()
        print(json.dumps({"peak_total_ms": e0.elapsed_time(e1) / REPS, "peak": pk.tolist()}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
