"""Binning splat of the two C2 fluence detectors (10^7 slots each, 100 x 83 map): time per `fluence_binning` call
(bounding box + splat + peak/total, CUDA events on the library's stream) for the accumulation variant chosen by
OPM_B200_BINNING_FIXED (unset/0 = f64 compare-and-swap adds in shared memory; 1.. = fixed-point sums from native 32-bit
shared-memory adds, CTA shapes 1024x1, 1024x2, 512x2, 512x3, 256x6), and the map itself for a comparison between variants.

    OPM_B200_BINNING_FIXED=1 python profiles/binning_probe.py out.npz
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200.api import fluence_binning  # noqa: E402


def main():
    n = int(os.environ.get("PROBE_RAYS", "10000000"))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = ob.Context(0, stream=stream.cuda_stream)
        sc = scenes.to_product(ctx, scenes.c2_laser_chain())
        src = scenes.source_to_product(scenes.c2_source(n)).generate(ctx)
        sc.raytrace(src, keep_source=True)
        out = {"variant": os.environ.get("OPM_B200_BINNING_FIXED", "0"), "rays": n}
        maps = {}
        for node in (9, 10):
            hm = sc.node_hitmaps(node)
            dev = ctx.device_alloc(100 * 83 * 8)
            fd = fluence_binning(ctx, hm, 100, 83, map_dev=dev)
            maps[str(node)] = fd.map
            for _ in range(3):
                fluence_binning(ctx, hm, 100, 83, map_dev=dev, download=False)
            reps = 20
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fluence_binning(ctx, hm, 100, 83, map_dev=dev, download=False)
            b.record(stream)
            b.synchronize()
            out[f"node{node}_ms_per_call"] = a.elapsed_time(b) / reps
            out[f"node{node}_peak"] = fd.peak
            out[f"node{node}_total"] = fd.total_energy
            ctx.device_free(dev)
        if len(sys.argv) > 1:
            np.savez(sys.argv[1], **maps)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
