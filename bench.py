#!/usr/bin/env python
"""bench.py -- ray-surface interactions/s of the ray-propagation hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1]): the multi-element laser chain -- lens, two 45 deg mirrors, wedge
(Sellmeier H-K9L), beam splitter with a detector on its second output, lens, filter, spot diagram, fluence detector --
traced by the ray-trace analyzer for 10^7 rays PER GPU (weak scaling: rank r owns source indices [r*n, (r+1)*n) of an
N*10^7-ray FibonacciEllipse / General2DGaussian source).  A step = one `RayTracingAnalyzer::analyze` of the scene over
the rank's shard + the detector read-out (two 100x83 Binning fluence maps, spot statistics); with N > 1 the fluence
maps and tallies are all-reduced over NCCL (min/max of the data-dependent map range first, then the maps).

  value     device-resident: the source bundle sits in HBM when the timed region starts (CUDA events, max over ranks)
  e2e       same step through the host-facing API: pinned host positions + energies (what the reference's position /
            energy distributions hand to `Rays::new_collimated_with_spectrum`) are copied to the device every step and
            the detector reports are read back to the host
  roofline  the fused trace kernel: algorithmic bytes / its CUDA-event launch time against the measured HBM copy peak;
            `fp64` says how far the same launches are from the measured DFMA peak (the fused kernel is FP64-pipe bound)
  cpu_baseline / --impl reference   the CPU oracle (a restatement of the reference's algorithm, oracle/) on the host cores

The oracle is imported only in the cpu_baseline / reference legs.  No CPU fallback: without the CUDA library or a
GPU the product arm raises.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "ray-surface interactions/s"
UNIT = "interactions/s"
RAY_BYTES = 93     # SoA bundle: 10 f64 + 3 u32 + 1 u8 (opossum_b200/csrc/opm_internal.h)
RAY_BYTES_INPLACE = 81  # in-place write: wavelength and id never change
HIT_BYTES = 36     # hit record: 4 f64 + u32
MAP_NX, MAP_NY = 100, 83  # nodes/fluence_detector/mod.rs:103


# ---------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md): sampled during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (one worker process per core, each tracing a contiguous shard of the sample)
# ---------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(k):
    import scenes
    src = _CPU["shards"][k]
    osc = _CPU.get("scene")
    if osc is None:
        osc = _CPU["scene"] = scenes.to_oracle(scenes.c2_laser_chain(), n_threads=1)
    t0 = time.perf_counter()
    osc.raytrace(src, record_all=False)
    for node in (9, 10):
        osc.node_fluence(node, MAP_NX, MAP_NY)
    osc.node_spot(8)
    return osc.interactions, time.perf_counter() - t0


class CpuArm:
    """The oracle's restatement of the reference's ray-trace analysis, sharded over `cores` worker processes."""

    def __init__(self, rays_per_core: int, cores: int | None = None):
        import multiprocessing as mp

        import numpy as np

        import scenes
        self.cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
        self.n = rays_per_core * self.cores
        src = scenes.source_to_oracle(scenes.c2_source(self.n))
        _CPU["shards"] = np.array_split(src, self.cores)
        _CPU.pop("scene", None)
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None
        self.sample = (f"{self.n} rays of the C2 source ({rays_per_core} per worker process x {self.cores}), full chain + detector read-out; "
                       "oracle port, per-surface passes over an AoS bundle like the reference")

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, range(self.cores), chunksize=1) if self.pool else [_cpu_worker(0)]
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res), dt

    def single_thread(self):
        inter, dt = _cpu_worker(0)
        return inter / dt

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.cpu_rays_per_core)
    for _ in range(args.warmup):
        arm.step()
    inter, dt = 0, 0.0
    for _ in range(args.steps):
        i, t = arm.step()
        inter += i
        dt += t
    value = inter / dt
    arm.close()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "rays_per_step": arm.n, "surfaces_per_ray": 13},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


WORKLOAD_NAME = "C2 multi-element laser chain (lens, 2 mirrors, Sellmeier wedge, beam splitter + side detector, lens, filter, " \
                "spot diagram, fluence detector), ray-trace analyzer, FibonacciEllipse/General2DGaussian source, 1054 nm"


# ---------------------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local: int) -> str:
    """Best effort: run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    (first touch) are local to the GPU's PCIe root and H2D copies do not cross the socket interconnect."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        cpus = (base / "local_cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if node >= 0 and ids:
            os.sched_setaffinity(0, ids)
            return f"numa node {node} ({len(ids)} cpus)"
        return "no numa information"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def run_product(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import opossum_b200 as ob
    import scenes
    from opossum_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: opossum_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    n = args.rays
    with torch.cuda.stream(stream):
        ctx = ob.Context(local, stream=stream.cuda_stream)
        scene = scenes.to_product(ctx, scenes.c2_laser_chain())
        sd = scenes.source_to_product(scenes.c2_source(n * world))
        # this rank's shard of the global source, generated on the device: interleaved indices {rank, rank + G, ...} so that
        # every GPU sees the same mix of clipped and surviving rays (a contiguous split leaves the outer, clipped part of the
        # Fibonacci index range to the last rank)
        first, count, stride = sharding.shard_strided(n * world, rank, world)
        src = sd.generate_strided(ctx, first, count, stride)


        readout = sharding.DetectorReadout(ctx, scene, (9, 10), (8,), MAP_NX, MAP_NY, dev)

        def report(rep):
            s = rep[8]
            out = {n: (rep[n]["peak"], rep[n]["total_energy"]) for n in (9, 10)}
            out["spot"] = [float(s.n_valid), s.total_energy, s.energy_centroid[0], s.energy_centroid[1], s.beam_radius_geo]
            return out

        def read_out(download: bool):
            """detector reports: FluenceDetector x2 (Binning 100x83, data-dependent range) + SpotDiagram, chained on the device
            (opossum_b200/sharding.py); with N > 1 the map ranges, maps and spot accumulators meet in two NCCL collectives"""
            return report(readout.run(download_maps=download))

        pipe = {"k": 0, "pending": None, "last": None}

        def step_resident():
            """one analysis + read-out; the host reads the reports of step k-1 after it has queued step k (two sets of pinned
            report buffers), so the device never waits for the host between steps"""
            scene.raytrace(src, keep_source=True, sync=False)
            slot = pipe["k"] % 2
            pipe["k"] += 1
            readout.launch(False, slot)
            prev, pipe["pending"] = pipe["pending"], slot
            if prev is not None:
                pipe["last"] = report(readout.collect(prev))

        def drain():
            if pipe["pending"] is not None:
                pipe["last"] = report(readout.collect(pipe["pending"]))
                pipe["pending"] = None
            return pipe["last"]

        # ---- e2e inputs: what the reference's distributions produce on the host (positions n x 3, energies n)
        pos_h = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        e_h = torch.empty(n, dtype=torch.float64).pin_memory()
        tmp = np.empty(n)
        for k, f in enumerate(("px", "py", "pz")):
            ctx.to_host(src.field_ptr(f), tmp)
            pos_h[:, k] = torch.from_numpy(tmp)
        ctx.to_host(src.field_ptr("e"), tmp)
        e_h.copy_(torch.from_numpy(tmp))
        e2e_rays = [ob.Rays(ctx, n), ob.Rays(ctx, n)]
        wvl, w = scenes.spectrum_of(scenes.c2_source(n))
        e2e_k = [0]

        def upload(rays):
            rays.new_collimated_with_spectrum(pos_h.data_ptr(), e_h.data_ptr(), n, wvl, w)

        def step_e2e():
            """host buffers -> reports.  Two bundles alternate: the upload of the NEXT step's inputs runs on the library's
            copy stream while this step traces, so every step still pays one full host->device copy inside the timed region."""
            cur, nxt = e2e_rays[e2e_k[0] % 2], e2e_rays[(e2e_k[0] + 1) % 2]
            e2e_k[0] += 1
            upload(nxt)
            scene.raytrace(cur, sync=False)
            return read_out(True)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- one checked run: interactions per step, buffer populations for the byte model, launches per step
        st = scene.raytrace(src, keep_source=True, sync=True)
        check = read_out(False)
        inter_per_step = st.interactions
        alive_main = st.alive_out
        split_rays = scene.node_rays(5)
        alive_side = split_rays.nr_of_rays(valid_only=False) if split_rays is not None else 0
        ctx.jit_wait()  # the specialised kernels of this scene are compiled in the background: measure the steady state
        step_resident()
        l0 = ctx.launch_count
        step_resident()
        launches_per_step = ctx.launch_count - l0
        steady = drain()
        for key in (9, 10, "spot"):  # the pipelined steps deliver the reports of the checked run
            assert np.allclose(np.asarray(steady[key], dtype=float), np.asarray(check[key], dtype=float), rtol=1e-9), (key, steady[key], check[key])

        # ---- value: device-resident
        for _ in range(args.warmup):
            step_resident()
        drain()
        ctx.set_kernel_timing(True)
        ctx.kernel_timing(reset=True)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step_resident()
        ev1.record(stream)
        drain()  # the reports of the last step
        barrier()
        ms = ev0.elapsed_time(ev1)
        kern_ms, kern_launches = ctx.kernel_timing(reset=True)
        ctx.set_kernel_timing(False)

        # ---- e2e: host buffers in, reports out
        upload(e2e_rays[0])
        for _ in range(max(args.warmup, 1)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        clk = clocks.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and e2e)

        t = torch.tensor([ms, e2e_ms, float(inter_per_step), kern_ms], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, e2e_ms, kern_ms = tmax[0].item(), tmax[1].item(), tmax[3].item()
            inter_total = tsum[2].item()
        else:
            inter_total = float(inter_per_step)

        fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
        jit = ctx.jit_stats()

    if rank == 0:
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            hbm_peak, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        # algorithmic bytes of the trace launches of ONE step on this rank (DESIGN.md "byte model"): every buffer the step
        # must read or write, once.  Read the source bundle (n x 93 B); for the rays still in the bundle write the main
        # output bundle, the spot diagram's copy of the bundle (93 B) and the fluence detector's hit record (36 B), and the
        # split-off bundle after its side chain with its hit record.  (The fluence detectors keep only hit maps: their
        # report reads nothing else, nodes/fluence_detector/mod.rs:91-135.)  When the side chain is its own launch
        # (2 trace launches per step) the split-off bundle is additionally written by the main launch and read back.
        launches_trace = kern_launches / args.steps
        bytes_step = (n * RAY_BYTES + alive_main * (2 * RAY_BYTES + HIT_BYTES) + alive_side * (RAY_BYTES + HIT_BYTES))
        if launches_trace > 1.5:
            bytes_step += alive_side * (RAY_BYTES + RAY_BYTES_INPLACE)
        kern_ms_step = kern_ms / args.steps
        achieved = bytes_step / (kern_ms_step * 1e-3) / 1e9
        unfused = inter_per_step * 146.0 / (kern_ms_step * 1e-3) / 1e9  # 146 B per stand-alone interaction
        flop_per_inter = 300.0  # SURVEY 8d estimate (FMA = 2 flop), sqrt/div counted as one
        h2d = n * 32 + 16 * len(wvl)
        d2h = 2 * MAP_NX * MAP_NY * 8 + 2 * 48 + 64
        # DRAM bytes of the same launch from the committed `ncu --set full` capture (never measured under the profiler here)
        traffic = None
        try:
            tr = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())
            same_bytes = abs(tr["algorithmic_bytes_per_launch"] - bytes_step / max(launches_trace, 1)) < 0.02 * tr["algorithmic_bytes_per_launch"]
            if jit["launches"] and tr["workload_rays_per_gpu"] == n and tr["launches_per_step"] == launches_trace and same_bytes:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": inter_total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "rays_per_gpu": n, "rays_total": n * world, "sharding": "source index i -> rank i mod N (interleaved)", "surfaces_per_ray": 13,
                       "interactions_per_step": int(inter_total), "fluence_map": f"{MAP_NX}x{MAP_NY} x2 (Binning, data-dependent range)",
                       "l2": f"inputs larger than L2: {n * RAY_BYTES / 1e6:.0f} MB source bundle per GPU vs 126 MB",
                       "host_pipeline": "the reports of step k are read on the host after step k+1 is queued (two pinned buffer sets); all K reports are copied to the host inside the timed region",
                       "collective": "2 NCCL collectives per step: all_gather(map ranges | spot sums), all_gather(fluence maps | spot radii), each combined in rank order by one kernel" if world > 1 else "none (1 GPU)"},
            "e2e": {"value": inter_total * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "pipeline": "two bundles: H2D of step k+1 (copy stream) overlaps trace + read-out of step k",
                    "host_binding": numa},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": ("opm_jit_trace (fused surface sequence, compiled at run time for this op program)" if jit["launches"] else
                                                   "opm::trace_kernel (fused surface sequence, interpreter)"),
                         "jit": jit, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_step / max(launches_trace, 1), "launches_per_step": launches_trace,
                         "kernel_ms_per_launch": kern_ms_step / max(launches_trace, 1), "kernel_share_of_step": kern_ms_step / (ms / args.steps),
                         "unfused_equivalent_gbs": unfused,
                         "fp64": {"peak_tflops_measured_dfma": fp64_peak, "flop_per_interaction_model": flop_per_inter,
                                  "achieved_tflops_model": inter_per_step * flop_per_inter / (kern_ms_step * 1e-3) / 1e12}},
            "check": {"fluence_peak_J_m2": {str(k): v[0] for k, v in check.items() if k != "spot"},
                      "detector_energy_J": {str(k): v[1] for k, v in check.items() if k != "spot"}, "spot_valid_rays": check["spot"][0]},
        }
    if world > 1:
        dist.barrier()
    # ---- cpu_baseline on rank 0 at N = 1 (test infrastructure timed beside the product; never on the product path)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm(args.cpu_rays_per_core)
            arm.step()
            best = 0.0
            for _ in range(2):
                i, dt = arm.step()
                best = max(best, i / dt)
            one = arm.single_thread()
            arm.close()
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.sample,
                                    "single_thread_value": one}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ctx.close()  # closes the scene and every bundle that lives on the context first
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--rays", type=int, default=10_000_000, help="rays per GPU (BASELINE.json configs[1]: 10^7)")
    ap.add_argument("--cpu-rays-per-core", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "product" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "product" and world != args.gpus:
        if args.gpus == 1 or world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
        raise SystemExit("multi-GPU runs are launched with `python -m torch.distributed.run --nproc-per-node N bench.py --gpus N`")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
