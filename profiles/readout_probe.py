"""One C2 analysis (10^7 rays) followed by one detector read-out: the process `ncu` profiles for the read-out kernels
(hits_bbox, spot_pass1/2, binning_smem_fixed, peak_total).

    ncu --set full --clock-control none -k regex:'bbox|spot_pass|binning' -c 8 -o gpurun_out/readout python profiles/readout_probe.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200 import sharding  # noqa: E402

stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = ob.Context(0, stream=stream.cuda_stream)
    ctx.set_jit(0)
    sc = scenes.to_product(ctx, scenes.c2_laser_chain())
    src = scenes.source_to_product(scenes.c2_source(10_000_000)).generate(ctx)
    ro = sharding.DetectorReadout(ctx, sc, (9, 10), (8,), 100, 83, torch.device("cuda", 0))
    sc.raytrace(src, keep_source=True)
    rep = ro.run()
    print({k: (v["peak"], v["total_energy"]) if isinstance(v, dict) else v.n_valid for k, v in rep.items()})
