"""C4 at N = 1 (10^8 rays, 4096^2 map): step time of trace + read-out in consecutive passes, with and without the library's
event pair around every trace launch, and the SM clock / power during each pass."""
from __future__ import annotations

import json
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import opossum_b200 as ob  # noqa: E402
import scenes  # noqa: E402
from opossum_b200 import sharding  # noqa: E402

SIDE, NX, NY = 10_000, 4096, 4096


def smi():
    out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                          "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    return out


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = ob.Context(0, stream=stream.cuda_stream)
        scene = scenes.to_product(ctx, scenes.c4_fluence())
        src = scenes.source_to_product(scenes.c4_source(SIDE)).generate(ctx, 0, SIDE * SIDE)
        scene.raytrace(src, keep_source=True, snapshots=False)
        ro = sharding.DetectorReadout(ctx, scene, (2,), (), NX, NY, dev, peer=False, solo=True)
        ctx.jit_wait()

        def step():
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
            ro.launch(False, 0)

        def run(steps, timing):
            ctx.set_kernel_timing(timing)
            ctx.kernel_timing(reset=True)
            samples = []
            stop = threading.Event()
            th = threading.Thread(target=lambda: [samples.append(smi()) or time.sleep(0.01) for _ in iter(lambda: stop.is_set(), True)])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            th.start()
            e0.record(stream)
            for _ in range(steps):
                step()
            e1.record(stream)
            ro.collect(0)
            torch.cuda.synchronize()
            stop.set(); th.join()
            kms, _ = ctx.kernel_timing(reset=True)
            ctx.set_kernel_timing(False)
            return {"steps": steps, "events": timing, "ms_per_step": e0.elapsed_time(e1) / steps, "trace_ms": kms / steps if timing else None,
                    "smi": samples[len(samples) // 2] if samples else None}

        for _ in range(3):
            step()
        ro.collect(0)
        for steps, timing in ((10, False), (10, True), (10, False), (40, False), (10, True), (10, False)):
            print(json.dumps(run(steps, timing)), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
