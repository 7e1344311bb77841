"""BASELINE.json configs[3] across GPUs: a 4096 x 4096 Binning fluence map from 10^8 rays SHARDED over the ranks (strong scaling:
the total ray count is fixed), detector read-out with the NCCL collectives of opossum_b200/sharding.py (all_gather of the map
range, all_reduce(sum) of the 128 MiB map).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/run_c4_multi.py

Rank 0 prints one JSON line: per-step time (trace + read-out + collectives, CUDA events, max over ranks), interactions/s, the
time of the map all-reduce alone, and the size-independent checks (mirror symmetry of the map, total energy, map range) that
must not depend on N.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200 import sharding  # noqa: E402

SIDE, NX, NY, STEPS = 10_000, 4096, 4096, 5


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = ob.Context(local, stream=stream.cuda_stream)
        scene = scenes.to_product(ctx, scenes.c4_fluence())
        sd = scenes.source_to_product(scenes.c4_source(SIDE))
        first, count, stride = sharding.shard_strided(SIDE * SIDE, rank, world)
        src = sd.generate_strided(ctx, first, count, stride)
        readout = sharding.DetectorReadout(ctx, scene, (2,), (), NX, NY, dev)

        def step():
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
            return readout.run(download_maps=False)

        st = scene.raytrace(src, keep_source=True, snapshots=False)
        rep = readout.run(download_maps=False)
        ctx.jit_wait()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(STEPS):
            rep = step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / STEPS
        m = readout.maps[0].view(NX, NY)  # column-major ny x nx: m[col, row]
        sym_x = float((m - m.flip(0)).abs().max() / rep[2]["peak"])
        sym_y = float((m - m.flip(1)).abs().max() / rep[2]["peak"])
        # the map all-reduce alone (this scales the map by the world size each time: the checks above come first)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(STEPS):
            sharding.allreduce_map(readout.maps)
        a1.record(stream)
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / STEPS
        t = torch.tensor([ms, ar_ms, float(st.interactions)], dtype=torch.float64, device=dev)
        tsum = t.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if rank == 0:
            inter = tsum[2].item()
            print(json.dumps({"config": "C4 4096x4096 fluence map, 10^8 rays sharded over the ranks (strong scaling)", "n_gpus": world,
                              "rays_total": SIDE * SIDE, "rays_per_gpu": count, "interactions": inter, "ms_per_step": t[0].item(),
                              "interactions_per_s": inter / (t[0].item() * 1e-3), "map_allreduce_ms": t[1].item() if world > 1 else 0.0,
                              "map_bytes": NX * NY * 8,
                              "checks": {"total_energy_J": rep[2]["total_energy"], "peak_J_m2": rep[2]["peak"], "lrtb": rep[2]["lrtb"],
                                         "mirror_symmetry_x_rel": sym_x, "mirror_symmetry_y_rel": sym_y,
                                         "symmetric_within_1e-9": bool(max(sym_x, sym_y) < 1e-9)}}))
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
