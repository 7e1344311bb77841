"""One rank's share of the C4 strong-scaling step at N = 8 on ONE GPU: the contiguous shard [0, 10^8 / 8) of the grid source,
trace + the peer-memory read-out with a one-rank peer group (same kernels and launch sequence as at N = 8; the meetings find
their own flag at once, the slice reduce reads no peer).  Shows where the read-out's time goes without cross-GPU waiting.

    python profiles/r2_c4_shard_probe.py [shards]      # JSON line; run it under `ncu --metrics gpu__time_duration.sum` for the launch list
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200 import _ffi as F  # noqa: E402
from opossum_b200 import sharding  # noqa: E402
from opossum_b200.api import _check  # noqa: E402

SHARDS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SIDE, NX, NY = 10_000, 4096, 4096
REPS = 10


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    L = F.lib()
    vp = C.c_void_p
    with torch.cuda.stream(stream):
        ctx = ob.Context(0, stream=stream.cuda_stream)
        scene = scenes.to_product(ctx, scenes.c4_fluence())
        n = SIDE * SIDE // SHARDS
        src = scenes.source_to_product(scenes.c4_source(SIDE)).generate(ctx, 0, n)
        scene.raytrace(src, keep_source=True, snapshots=False)
        ctx.jit_wait()
        grp = sharding.PeerGroup(ctx, NX * NY)
        hm = scene.node_hitmaps(2, 0)
        arr = (C.c_void_p * len(hm))(*[h.h for h in hm])
        rng = torch.zeros(4, dtype=torch.float64, device=dev)
        pk = torch.zeros(2, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)

        def trace():
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)

        def readout():
            _check(L.opm_peer_fluence_report_async(grp.h, arr, len(hm), -1, NX, NY, vp(rng.data_ptr()), vp(pk.data_ptr()), vp(status.data_ptr()), None))

        def timed(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(REPS):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / REPS

        t_trace = timed(trace)
        t_read = timed(readout)
        l0 = ctx.launch_count
        trace(); readout()
        launches = ctx.launch_count - l0
        t_both = timed(lambda: (trace(), readout()))
        print(json.dumps({"rays": n, "trace_ms": t_trace, "readout_ms": t_read, "step_ms": t_both, "launches_per_step": launches,
                          "peak": pk.tolist(), "range": rng.tolist()}))
        ctx.close()


if __name__ == "__main__":
    main()
