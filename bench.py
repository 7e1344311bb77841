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
SPOT_COPY_BYTES = 41  # spot-diagram copy of the bundle: position (3 f64) + energy + id + bounce count + valid flag (OPM_SCENE_LEAN_SPOT_DATA)
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


def _cpu_worker(arg):
    import scenes
    k, keep = arg
    src = _CPU["shards"][k]
    osc = _CPU.get("scene")
    if osc is None:
        osc = _CPU["scene"] = scenes.to_oracle(scenes.c2_laser_chain(), n_threads=1)
    t0 = time.perf_counter()
    out = osc.raytrace(src, record_all=False)
    maps = {node: osc.node_fluence(node, MAP_NX, MAP_NY) for node in (9, 10)}
    spot = osc.node_spot(8)
    dt = time.perf_counter() - t0
    if keep:  # the shard the product is compared with (parity check of the bench's own workload; never inside a timed step)
        return osc.interactions, dt, {"out": out, "side": osc.chain_out[10], "maps": maps, "spot": spot, "interactions": osc.interactions}
    return osc.interactions, dt, None


class CpuArm:
    """The oracle's restatement of the reference's ray-trace analysis over the bench's OWN source (C2, `n_rays` rays, the same
    FibonacciEllipse / General2DGaussian description the product generates on the device), sharded over `cores` worker
    processes by source index i -> worker i mod cores (rays are independent: any partition is one analysis)."""

    def __init__(self, n_rays: int, cores: int | None = None):
        import multiprocessing as mp

        import scenes
        self.cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
        self.n = n_rays
        src = scenes.source_to_oracle(scenes.c2_source(self.n))
        _CPU["shards"] = [src[k::self.cores].copy() for k in range(self.cores)]
        _CPU.pop("scene", None)
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None
        self.sample = (f"the bench's own C2 source, all {self.n} rays, source index i -> worker process i mod {self.cores}; full chain + "
                       "detector read-out per worker; oracle port, per-surface passes over an AoS bundle like the reference")

    def step(self):
        t0 = time.perf_counter()
        args = [(k, False) for k in range(self.cores)]
        res = self.pool.map(_cpu_worker, args, chunksize=1) if self.pool else [_cpu_worker(args[0])]
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res), dt

    def trace_kept(self, k=0):
        """results of worker k's shard (untimed): final bundles, both maps, spot statistics"""
        return (self.pool.apply(_cpu_worker, ((k, True),)) if self.pool else _cpu_worker((k, True)))[2]

    def single_thread(self):
        inter, dt, _ = _cpu_worker((0, False))
        return inter / dt

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.rays)
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
        "config": {"workload": WORKLOAD_NAME, "rays_per_gpu": arm.n, "rays_total": arm.n, "surfaces_per_ray": 13,
                   "fluence_map": f"{MAP_NX}x{MAP_NY} x2 (Binning, data-dependent range)"},
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
        return _bind_by_topo(local)
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def _bind_by_topo(local: int) -> str:
    """sysfs has no numa_node in this container: take the GPU's "CPU Affinity" column of `nvidia-smi topo -m`"""
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:  # noqa: BLE001
        return "no numa information"
    import re
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    lines = [ln for ln in out.splitlines() if ln.strip()]
    if not lines:
        return "no numa information"
    header = [h.strip() for h in lines[0].split("\t")]
    if "CPU Affinity" not in header:
        return "no numa information"
    col = header.index("CPU Affinity")
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = local
    if visible:
        try:
            phys = int(visible.split(",")[local])
        except (ValueError, IndexError):
            phys = local
    for ln in lines[1:]:
        cells = [c.strip() for c in ln.split("\t")]
        if cells and cells[0] == f"GPU{phys}" and len(cells) > col:
            ids = set()
            try:
                for part in cells[col].split(","):
                    a, _, b = part.partition("-")
                    ids.update(range(int(a), int(b or a) + 1))
            except ValueError:
                return "no numa information"
            allowed = os.sched_getaffinity(0)
            if ids and len(ids & allowed) and len(ids) < len(allowed):
                os.sched_setaffinity(0, ids & allowed)
                numa = cells[col + 1] if len(cells) > col + 1 else "?"
                return f"nvidia-smi topo: cpus {cells[col]} (numa {numa})"
            return f"nvidia-smi topo: one affinity domain ({cells[col]}), nothing to bind"
    return "no numa information"


def parity_vs_oracle(ctx, n, cores, kept):
    """Untimed check of the bench's own workload against the CPU oracle: the product traces the shard of the 10^7-ray source one
    CPU worker traced (source indices 0, cores, 2 cores, ...) and is compared ray by ray (integer state bit-exact, positions /
    directions / energies / OPL at 1e-9 relative), both fluence maps bin by bin on the oracle's grid (1e-9 relative for bins
    >= 1e-12 of the peak) and the spot statistics."""
    import numpy as np

    import opossum_b200 as ob
    import scenes
    sc = scenes.to_product(ctx, scenes.c2_laser_chain())
    count = (n + cores - 1) // cores
    rays = scenes.source_to_product(scenes.c2_source(n)).generate_strided(ctx, 0, count, cores)
    ctx.set_jit(2, 200000)  # the specialised kernel (the one the timed steps run), waiting for its compilation
    st = sc.raytrace(rays)
    out = {"rays_traced": count, "kernel": "opm_jit_trace" if ctx.jit_stats()["launches"] else "interpreter"}

    def rays_err(got, ref):
        if len(got) != len(ref):
            return {"len": [len(got), len(ref)]}
        e = {"int_state_equal": bool(all(np.array_equal(got[f], ref[f]) for f in ("id", "valid", "bounces", "refr")))}
        sp = np.maximum(np.linalg.norm(ref["pos"], axis=1), 1e-3)
        e["pos"] = float((np.linalg.norm(got["pos"] - ref["pos"], axis=1) / sp).max())
        e["dir"] = float(np.linalg.norm(got["dir"] - ref["dir"], axis=1).max())
        for f, floor in (("e", 1e-30), ("opl", 1e-12)):
            e[f] = float((np.abs(got[f] - ref[f]) / np.maximum(np.abs(ref[f]), floor)).max())
        return e

    out["main_bundle"] = rays_err(rays.to_array(), kept["out"])
    out["split_off_bundle"] = rays_err(sc.node_rays(5).to_array(), kept["side"])
    worst_bin = 0.0
    for node in (9, 10):
        fmap, lrtb, peak, tot = kept["maps"][node]
        fd = ob.fluence_binning(ctx, sc.node_hitmaps(node, 0), MAP_NX, MAP_NY, lrtb=lrtb)
        filled = np.abs(fmap) >= 1e-12 * np.abs(fmap).max()
        rel = float((np.abs(fd.map - fmap)[filled] / np.abs(fmap)[filled]).max())
        low = float(np.abs(fd.map - fmap)[~filled].max() / np.abs(fmap).max()) if (~filled).any() else 0.0
        out[f"map_{node}"] = {"worst_rel_bin": rel, "below_floor_abs_over_peak": low, "peak_rel": abs(fd.peak - peak) / peak,
                              "bbox_rel": float(max(abs(a - b) / abs(b) for a, b in zip(sc.node_fluence(node, MAP_NX, MAP_NY, download=False).lrtb, lrtb)))}
        worst_bin = max(worst_bin, rel, low * 1e3)
    s, r = sc.node_spot(8), kept["spot"]
    out["spot"] = {"n_valid_equal": bool(s.n_valid == r["n_valid"]), "total_energy": abs(s.total_energy - r["total_energy"]) / r["total_energy"],
                   "geo_radius": abs(s.beam_radius_geo - r["geo"]) / r["geo"], "rms_radius": abs(s.energy_beam_radius_rms - r["rms"]) / r["rms"]}
    rb = [out["main_bundle"], out["split_off_bundle"]]
    out["ok"] = bool(all("len" not in e and e["int_state_equal"] and max(e["pos"], e["dir"], e["e"], e["opl"]) <= 1e-9 for e in rb)
                     and worst_bin <= 1e-9 and out["spot"]["n_valid_equal"]
                     and max(out["spot"]["total_energy"], out["spot"]["geo_radius"], out["spot"]["rms_radius"]) <= 1e-9
                     and st.interactions == kept_interactions(kept))
    ctx.set_jit(1, 200000)
    return out


def kept_interactions(kept):
    return kept.get("interactions", 0)


C4_SIDE, C4_NX, C4_NY = 10_000, 4096, 4096


def run_c4_strong(args, ctx, dev, stream, rank, world):
    """BASELINE configs[3], the north-star scaling config: 10^8 rays IN TOTAL (grid 10^4 x 10^4) sharded over the ranks by
    contiguous source-index ranges, lens + fluence detector, one 4096 x 4096 Binning map of all hits.  A step = trace of the
    rank's shard + the detector read-out (common data-dependent range, windowed binning, column-slice merge over NVLink peer
    memory, peak / total on every rank).  Strong scaling: total work fixed.  With N > 1 rank 0 also runs the WHOLE job alone in
    the same process (untimed region of the others), which gives the speed-up basis on the same box and the parity reference
    for the merged map."""
    import torch
    import torch.distributed as dist

    import scenes
    from opossum_b200 import sharding
    steps = max(3, min(args.steps, 20))
    n_total = C4_SIDE * C4_SIDE

    def setup(first, count, peer, solo=False):
        scene = scenes.to_product(ctx, scenes.c4_fluence())
        src = scenes.source_to_product(scenes.c4_source(C4_SIDE)).generate(ctx, first, first + count)
        st = scene.raytrace(src, keep_source=True, snapshots=False)
        ro = sharding.DetectorReadout(ctx, scene, (2,), (), C4_NX, C4_NY, dev, peer=peer, solo=solo)
        return scene, src, ro, st

    def timed(scene, src, ro, barrier):
        def step():
            scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
            ro.launch(False, 0)
        for _ in range(3):
            step()
        ro.collect(0)
        # the trace kernel's share, measured in a pass of its own: the event pair around every trace launch stays out of the
        # timed steps (it costs them a few microseconds each)
        barrier()
        ctx.set_kernel_timing(True)
        ctx.kernel_timing(reset=True)
        for _ in range(steps):
            step()
        ro.collect(0)
        barrier()
        kms, _ = ctx.kernel_timing(reset=True)
        ctx.set_kernel_timing(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for _ in range(steps):
            step()
        host_ms = 1e3 * (time.perf_counter() - h0) / steps  # what the host needs to queue one step (must stay below the step)
        e1.record(stream)
        rep = ro.collect(0)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        timed.host_ms = host_ms
        return ms, kms / steps, rep[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    b, e = sharding.shard_range(n_total, rank, world)
    scene, src, ro, st = setup(b, e - b, None)
    ctx.jit_wait()
    ms, trace_ms, rep = timed(scene, src, ro, barrier)
    t = torch.tensor([ms, trace_ms, float(st.interactions)], dtype=torch.float64, device=dev)
    t_own = t.tolist()
    tsum = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, trace_ms, inter = t[0].item(), t[1].item(), tsum[2].item()
    out = {"config": "C4: 10^8 rays total (grid 10^4 x 10^4, contiguous source-index shards), lens + fluence detector, 4096 x 4096 Binning "
                     "map, data-dependent range", "scaling": "strong", "n_gpus": world, "rays_total": n_total, "rays_per_gpu": e - b,
           "steps": steps, "ms_per_step": ms, "interactions_per_s": inter / (ms * 1e-3), "trace_ms": trace_ms, "readout_ms": ms - trace_ms,
           "host_queue_ms_per_step": timed.host_ms,
           "readout": ("bounding box kept by the trace kernel, zero-fill + run-merging splat, peak/total" if world == 1 else
                       ("windowed splat + column-slice merge over NVLink peer memory, 3 flag meetings, no NCCL" if ro.peer is not None
                        else "NCCL all-reduce of the whole map (peer memory unavailable)")),
           "check": {"peak_J_m2": rep["peak"], "total_energy_J": rep["total_energy"], "lrtb": list(rep["lrtb"])}}
    if world > 1:
        # the collective part alone: three meetings of the ranks (what replaces all_gather + all_reduce + all_gather)
        if ro.peer is not None:
            blk = torch.zeros(4, dtype=torch.float64, device=dev)
            got = torch.zeros((world, 4), dtype=torch.float64, device=dev)
            for _ in range(3):
                ro.peer[0].allgather(blk, got)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(30):
                ro.peer[0].allgather(blk, got)
            e1.record(stream)
            barrier()
            tm = torch.tensor([e0.elapsed_time(e1) / 30], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            out["collective_ms"] = 3 * tm.item()
            out["meeting_us"] = 1e3 * tm.item()
            # where the read-out's time goes on this rank (diagnostic, one extra untimed step)
            import ctypes as C
            from opossum_b200 import _ffi as F
            ph = (C.c_float * 12)()
            F.lib().opm_peer_group_phase_times(ro.peer[0].h, 1, None)
            for _ in range(3):
                scene.raytrace(src, keep_source=True, snapshots=False, sync=False)
                ro.launch(False, 0)
            F.lib().opm_peer_group_phase_times(ro.peer[0].h, 0, ph)
            ro.collect(0)
            pt = torch.tensor(list(ph) + [t_own[1]], dtype=torch.float64, device=dev)
            allp = torch.zeros((world, 13), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allp.view(-1), pt)
            out["readout_phases_ms_per_rank"] = {k: [round(v, 4) for v in allp[:, j].tolist()]
                                                 for j, k in enumerate(("meeting1", "zero_window", "splat", "reduce_meetings23", "t_splat_start", "t_range_posted",
                                                                       "t_ranges_here", "t_reduce_start", "t_parts_binned", "t_last_cta_done",
                                                                       "t_peaks_here", "t_first_cta_done", "trace"))}
        # merged map on rank 0 (gathered from the column slices) for the parity check below
        merged = ro.run(download_maps=True)[2]
        barrier()
        del ro, src, scene
        if rank == 0:  # the whole job on one GPU of the same box: speed-up basis + parity reference
            scene1, src1, ro1, st1 = setup(0, n_total, False, solo=True)
            # the basis is the FASTER of two passes: the whole job on one GPU runs 7x longer per pass than a rank of the sharded
            # run and meets the board's power cap sooner (profiles/r2_summary.md); the conservative speed-up uses the better one
            ms1, trace1, rep1 = timed(scene1, src1, ro1, lambda: torch.cuda.synchronize())
            time.sleep(0.5)
            ms1b, trace1b, rep1 = timed(scene1, src1, ro1, lambda: torch.cuda.synchronize())
            n1_passes = [ms1, ms1b]
            if ms1b < ms1:
                ms1, trace1 = ms1b, trace1b
            ref = ro1.run(download_maps=True)[2]
            import numpy as np
            nz = ref["map"] > 0
            rel = float((np.abs(merged["map"] - ref["map"])[nz] / ref["map"][nz]).max())
            out["n1_same_box"] = {"ms_per_step": ms1, "trace_ms": trace1, "readout_ms": ms1 - trace1, "interactions_per_s": st1.interactions / (ms1 * 1e-3),
                                   "passes_ms_per_step": n1_passes}
            out["speedup_vs_n1_basis"] = ms1 / ms
            out["parity_vs_n1"] = {"map_worst_rel_bin": rel, "filled_bins_equal": bool(np.array_equal(nz, merged["map"] > 0)),
                                   "range_equal": bool(tuple(merged["lrtb"]) == tuple(ref["lrtb"])),
                                   "peak_rel": abs(merged["peak"] - ref["peak"]) / ref["peak"],
                                   "total_rel": abs(merged["total_energy"] - ref["total_energy"]) / ref["total_energy"],
                                   "interactions_equal": bool(int(inter) == st1.interactions),
                                   "ok": bool(rel <= 1e-12 and tuple(merged["lrtb"]) == tuple(ref["lrtb"]) and int(inter) == st1.interactions)}
            del ro1, src1, scene1
        barrier()
    else:
        out["collective_ms"] = 0.0
        out["speedup_vs_n1_basis"] = 1.0
    return out


C5_NPOS, C5_NWVL, C5_CHUNK, C5_NX = 100_000_000, 64, 1_562_500, 512


def run_c5(args, ctx, dev, stream, rank, world):
    """BASELINE configs[4]: broadband source (Gaussian spectrum, 64 lines over 950-1150 nm) through a Sellmeier lens and a prism pair,
    10^8 positions x 64 wavelengths = 6.4 * 10^9 rays IN TOTAL, position-major (rays.rs:284-289), sharded over the ranks by
    contiguous POSITION ranges (every rank owns whole 64-line groups) and streamed in chunks: the ray state is never resident in
    full.  Pass 1 traces every chunk for the detector's data-dependent map range, the ranks' ranges meet in one 4-double
    all-reduce(max); pass 2 traces again and bins into the common 512 x 512 grid; one all-reduce(sum) of the 2 MB map.  Strong
    scaling.  At N > 1 a reduced source (10^6 positions) is also run sharded and on rank 0 alone and the two maps are compared."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import opossum_b200 as ob
    import scenes
    from opossum_b200 import _ffi as F
    from opossum_b200 import sharding
    L = F.lib()
    nx = C5_NX
    sc = scenes.to_product(ctx, scenes.c5_broadband())
    rays = ob.Rays(ctx, C5_CHUNK * C5_NWVL)

    def run(n_pos, shard_world, shard_rank, reduce):
        sd = scenes.source_to_product(scenes.c5_source(n_pos, C5_NWVL))
        b0, e0 = sharding.shard_range(n_pos, shard_rank, shard_world)
        rng = torch.full((4,), -float("inf"), dtype=torch.float64, device=dev)
        tmp = torch.zeros(4, dtype=torch.float64, device=dev)
        fmap = torch.zeros(nx * nx, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if reduce:
            dist.barrier()
        torch.cuda.synchronize()
        inter0 = ctx.interactions_total()
        inter = 0
        t0 = time.perf_counter()
        for phase in (1, 2):
            first = True
            for b in range(b0, e0, C5_CHUNK):
                e = min(b + C5_CHUNK, e0)
                sd.generate(ctx, b, e, rays=rays)
                sc.raytrace(rays, snapshots=False, sync=False)
                hm = sc.node_hitmaps(4, 0)
                arr = (C.c_void_p * len(hm))(*[h.h for h in hm])
                if phase == 1:
                    ob.api._check(L.opm_hitmaps_bounding_box_async(arr, len(hm), -1, C.c_void_p(tmp.data_ptr())))
                    torch.maximum(rng, tmp, out=rng)
                else:
                    ob.api._check(L.opm_fluence_binning_async(arr, len(hm), -1, nx, nx, C.c_void_p(rng.data_ptr()), 0 if first else 1,
                                                              C.c_void_p(fmap.data_ptr()), C.c_void_p(status.data_ptr())))
                    first = False
            if phase == 1:
                inter = ctx.interactions_total() - inter0  # the launches raise no count on the host: read the device counter once
                if reduce:
                    dist.all_reduce(rng, op=dist.ReduceOp.MAX)  # (-l, r, t, -b): the common range is the element-wise maximum
        if reduce:
            dist.all_reduce(fmap, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        if reduce:
            dist.barrier()
        dt = time.perf_counter() - t0
        r = rng.tolist()
        return inter, dt, fmap, (-r[0], r[1], r[2], -r[3]), int(status.item())

    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream, device=dev)):
        run(2 * C5_CHUNK * world, world, rank, world > 1)  # warm-up: kernels compiled, buffers sized
        inter, dt, fmap, lrtb, status = run(C5_NPOS, world, rank, world > 1)
        t = torch.tensor([float(inter), dt], dtype=torch.float64, device=dev)
        if world > 1:
            tsum, tmax = t.clone(), t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            inter, dt = int(tsum[0].item()), tmax[1].item()
        fd = ob.fluence_peak_total(ctx, fmap.data_ptr(), nx, nx, lrtb)
        out = {"config": "C5: 10^8 positions x 64 wavelengths (Gaussian spectrum 950-1150 nm) = 6.4e9 rays total through a Sellmeier lens + prism pair, "
                         "position-major, contiguous position shards, streamed in chunks of 1.5625e6 positions x 64 lines; two passes "
                         "(range, then Binning into 512 x 512)", "scaling": "strong", "n_gpus": world, "rays_total": C5_NPOS * C5_NWVL,
               "interactions_counted": inter, "interactions_traced": 2 * inter, "seconds": dt,
               "interactions_per_s": inter / dt, "traced_interactions_per_s": 2 * inter / dt,
               "collectives": "all_reduce(max) of the 4-double range, all_reduce(sum) of the 2 MB map (NCCL)" if world > 1 else "none",
               "check": {"status_bits": status, "map_total_energy_J": fd.total_energy, "peak_J_m2": fd.peak, "lrtb": list(lrtb)}}
        if world > 1:  # parity of the sharded path on a reduced source: merged map vs rank 0 alone
            n_small = 1_000_000
            _, _, fm, lr, _ = run(n_small, world, rank, True)
            if rank == 0:
                _, _, fm1, lr1, _ = run(n_small, 1, 0, False)
                a, b = fm.cpu().numpy(), fm1.cpu().numpy()
                nz = b > 0
                out["parity_vs_one_gpu"] = {"positions": n_small, "range_equal": bool(tuple(lr) == tuple(lr1)),
                                            "filled_bins_equal": bool(np.array_equal(nz, a > 0)),
                                            "map_worst_rel_bin": float((np.abs(a - b)[nz] / b[nz]).max())}
                out["parity_vs_one_gpu"]["ok"] = bool(out["parity_vs_one_gpu"]["range_equal"] and out["parity_vs_one_gpu"]["filled_bins_equal"]
                                                      and out["parity_vs_one_gpu"]["map_worst_rel_bin"] <= 1e-9)
            dist.barrier()
    return out


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
            scene.raytrace(src, keep_source=True, sync=False, lean_spot_data=True)
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

        # ---- e2e inputs: what the reference's PositionDistribution produces on the host -- the points of this rank's shard, (x, y)
        # only: every position distribution of the reference has z = 0 -- and the EnergyDistribution as a descriptor: the energies
        # are a function of (x, y) (general_gaussian.rs:93-120) and are evaluated on the device, like the spectrum
        pos_h = torch.empty((n, 2), dtype=torch.float64).pin_memory()
        tmp = np.empty(n)
        for k, f in enumerate(("px", "py")):
            ctx.to_host(src.field_ptr(f), tmp)
            pos_h[:, k] = torch.from_numpy(tmp)
        e2e_rays = [ob.Rays(ctx, n), ob.Rays(ctx, n)]
        wvl, w = scenes.spectrum_of(scenes.c2_source(n))
        sd_e2e = scenes.source_to_product(scenes.c2_source(n * world))
        if world > 1:  # a shard cannot know the normalisation sum of the whole source: it is a property of the source, taken once
            en = scenes.c2_source(n * world)["energy"]
            sd_e2e.general_gaussian(en[2], en[3], en[4], en[5], en[6], norm=sd.energy_norm(ctx))
        e2e_k = [0]

        def upload(rays):
            rays.new_collimated_xy(pos_h.data_ptr(), n, sd_e2e)

        def queue_reports(p):
            """queue the read-out of the step just traced WITH its device->host copies (maps included) and read the reports of the
            step before on the host: like step_resident, the device never waits for the host between steps, and every report is
            still read inside the timed region (drain_reports before the closing barrier)"""
            slot = p["k"] % 2
            p["k"] += 1
            readout.launch(True, slot)
            prev, p["pending"] = p["pending"], slot
            if prev is not None:
                p["last"] = report(readout.collect(prev))

        def drain_reports(p):
            if p["pending"] is not None:
                p["last"] = report(readout.collect(p["pending"]))
                p["pending"] = None
            return p["last"]

        pipe_e2e = {"k": 0, "pending": None, "last": None}
        pipe_doc = {"k": 0, "pending": None, "last": None}

        def step_e2e():
            """host buffers -> reports.  Two bundles alternate: the upload of the NEXT step's inputs runs on the library's
            copy stream while this step traces, so every step still pays one full host->device copy inside the timed region;
            the reports (two maps + statistics) of a step are read on the host while the next step runs."""
            cur, nxt = e2e_rays[e2e_k[0] % 2], e2e_rays[(e2e_k[0] + 1) % 2]
            e2e_k[0] += 1
            upload(nxt)
            scene.raytrace(cur, sync=False, lean_spot_data=True)
            queue_reports(pipe_e2e)

        doc_rays = ob.Rays(ctx, n)

        def step_document():
            """document level, what `OpmDocument::analyze` does (opm_document.rs:270-289): source DESCRIPTOR in (position, energy
            and spectral distribution: a few hundred bytes) -> rays generated on the device -> trace -> reports on the host"""
            sd.generate_strided(ctx, first, count, stride, rays=doc_rays)
            scene.raytrace(doc_rays, sync=False, lean_spot_data=True)
            queue_reports(pipe_doc)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- one checked run: interactions per step, buffer populations for the byte model, launches per step
        st = scene.raytrace(src, keep_source=True, sync=True, lean_spot_data=True)
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
        drain_reports(pipe_e2e)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        e2e_check = drain_reports(pipe_e2e)  # the reports of the last step: all K reports are on the host before the clock stops
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        for key in (9, 10, "spot"):
            assert np.allclose(np.asarray(e2e_check[key], dtype=float), np.asarray(check[key], dtype=float), rtol=1e-9), (key, e2e_check[key], check[key])
        # ---- document-level e2e: descriptors in, reports out
        for _ in range(max(args.warmup, 1)):
            step_document()
        drain_reports(pipe_doc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_document()
        doc_check = drain_reports(pipe_doc)
        barrier()
        doc_ms = 1e3 * (time.perf_counter() - t0)
        for key in (9, 10, "spot"):
            assert np.allclose(np.asarray(doc_check[key], dtype=float), np.asarray(check[key], dtype=float), rtol=1e-9), (key, doc_check[key], check[key])
        clk = clocks.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and e2e)

        t = torch.tensor([ms, e2e_ms, float(inter_per_step), kern_ms, doc_ms], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, e2e_ms, kern_ms, doc_ms = tmax[0].item(), tmax[1].item(), tmax[3].item(), tmax[4].item()
            inter_total = tsum[2].item()
        else:
            inter_total = float(inter_per_step)

        fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
        jit = ctx.jit_stats()

        # ---- N > 1: the merged reports of the sharded run against ONE GPU tracing the whole global source (untimed)
        parity_n1 = None
        if world > 1:
            merged = readout.run(download_maps=True)
            barrier()
            if rank == 0:
                scene_g = scenes.to_product(ctx, scenes.c2_laser_chain())
                src_g = sd.generate(ctx)
                scene_g.raytrace(src_g, keep_source=True, lean_spot_data=True)
                solo = sharding.DetectorReadout(ctx, scene_g, (9, 10), (8,), MAP_NX, MAP_NY, dev, solo=True).run(download_maps=True)
                worst = 0.0
                for node in (9, 10):
                    a, b = merged[node]["map"], solo[node]["map"]
                    nz = b > 0
                    worst = max(worst, float((np.abs(a - b)[nz] / b[nz]).max()), 0.0 if np.array_equal(nz, a > 0) else 1.0)
                    worst = max(worst, 0.0 if tuple(merged[node]["lrtb"]) == tuple(solo[node]["lrtb"]) else 1.0)
                sa, sb = merged[8], solo[8]
                spot_rel = max(abs(sa.total_energy - sb.total_energy) / sb.total_energy, abs(sa.beam_radius_geo - sb.beam_radius_geo) / sb.beam_radius_geo,
                               abs(sa.energy_beam_radius_rms - sb.energy_beam_radius_rms) / sb.energy_beam_radius_rms)
                parity_n1 = {"what": f"reports of the {world}-GPU run vs the same {n * world}-ray source traced by GPU 0 alone",
                             "maps_worst_rel_bin": worst, "spot_worst_rel": spot_rel, "spot_n_valid_equal": bool(sa.n_valid == sb.n_valid),
                             "bounce_lvl_equal": bool(sa.bounce_lvl == sb.bounce_lvl),
                             "bar": "maps 1e-9 relative per bin (the north star's bar: the small-map splat sums in fixed point, each tap rounded "
                                    "by at most 2^-33 of itself, and the shards round different sets of taps), spot statistics 1e-12",
                             "ok": bool(worst <= 1e-9 and spot_rel <= 1e-12 and sa.n_valid == sb.n_valid and sa.bounce_lvl == sb.bounce_lvl)}
                del scene_g, src_g
            barrier()

        # ---- BASELINE configs[3]: the north-star strong-scaling config (10^8 rays total, 4096^2 map)
        c4 = None if args.no_c4 else run_c4_strong(args, ctx, dev, stream, rank, world)
        # ---- BASELINE configs[4]: broadband source, 10^8 positions x 64 wavelengths, streamed
        c5 = None if args.no_c5 else run_c5(args, ctx, dev, stream, rank, world)

    if rank == 0:
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            hbm_peak, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        # algorithmic bytes of the trace launches of ONE step on this rank (DESIGN.md "byte model"): every buffer the step
        # must read or write, once.  Read the source bundle (n x 93 B); for the rays still in the bundle write the main
        # output bundle, the spot diagram's copy of the bundle (41 B: position, energy, id, bounce count, valid flag) and the fluence detector's hit record (36 B), and the
        # split-off bundle after its side chain with its hit record.  (The fluence detectors keep only hit maps: their
        # report reads nothing else, nodes/fluence_detector/mod.rs:91-135.)  When the side chain is its own launch
        # (2 trace launches per step) the split-off bundle is additionally written by the main launch and read back.
        launches_trace = kern_launches / args.steps
        bytes_step = (n * RAY_BYTES + alive_main * (RAY_BYTES + SPOT_COPY_BYTES + HIT_BYTES) + alive_side * (RAY_BYTES + HIT_BYTES))
        if launches_trace > 1.5:
            bytes_step += alive_side * (RAY_BYTES + RAY_BYTES_INPLACE)
        kern_ms_step = kern_ms / args.steps
        achieved = bytes_step / (kern_ms_step * 1e-3) / 1e9
        unfused = inter_per_step * 146.0 / (kern_ms_step * 1e-3) / 1e9  # 146 B per stand-alone interaction
        h2d = n * 16 + 16 * len(wvl) + 256  # (x, y) of every position + spectrum + the energy distribution's descriptor
        d2h = 2 * MAP_NX * MAP_NY * 8 + 2 * 48 + 64
        # What only a profiler can count comes from the committed `ncu --set full` capture of the same launch (profiles/r2_traffic.json,
        # never measured under the profiler here): DRAM bytes, executed FP64 instructions, issue-slot and pipe utilisation.  It is
        # used only while it still describes this build's launch (same rays, launches per step and algorithmic bytes).
        traffic, ncu = None, None
        try:
            tr = json.loads((ROOT / "profiles" / "r2_traffic.json").read_text())["trace"]
            same_bytes = abs(tr["algorithmic_bytes_per_launch"] - bytes_step / max(launches_trace, 1)) < 0.02 * tr["algorithmic_bytes_per_launch"]
            if jit["launches"] and tr["workload_rays_per_gpu"] == n and tr["launches_per_step"] == launches_trace and same_bytes:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                ncu = tr
        except (OSError, ValueError, KeyError):
            pass
        # SURVEY 8(d)'s byte model: state S = 77 B per ray (9 f64 + u32 + u8), written without the read-only wavelength (69 B),
        # 32 B per hit record; the split-off bundle is written once (77 B); NO copy of the bundle for the spot diagram
        bytes_8d = n * 77 + alive_main * (69 + 32) + alive_side * (77 + 32)
        kern_s = kern_ms_step * 1e-3
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
                    "inputs": "the PositionDistribution's points as (x, y) in pinned host memory (z = 0 for every reference distribution); "
                              "EnergyDistribution, spectrum and isometry as descriptors evaluated on the device "
                              "(opm_rays_new_collimated_xy, rays.rs:264-297)",
                    "pipeline": "two bundles: H2D of step k+1 (copy stream) overlaps trace + read-out of step k; the reports of step k "
                                "(two maps + statistics, copied device->host every step) are read on the host while step k+1 runs, the "
                                "last ones before the clock stops",
                    "host_binding": numa,
                    "document_level": {"value": inter_total * args.steps / (doc_ms * 1e-3), "unit": UNIT, "ms_per_step": doc_ms / args.steps,
                                       "h2d_bytes_per_step": 512, "d2h_bytes_per_step": d2h,
                                       "what": "source descriptor in (opm_source_generate on the device) -> trace -> reports on the host, "
                                               "what OpmDocument::analyze does (opm_document.rs:270-289)"}},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk,
            "roofline": {"bound": "issue",
                         "bound_note": ("the fused kernel is bound by instruction issue (ncu: issue slots %.0f %% busy, FP64 pipe %.0f %%, DRAM %.0f %%); "
                                        % (ncu["issue_active_pct"], ncu["fp64_pipe_pct"], ncu["dram_throughput_pct"]) if ncu else
                                        "the fused kernel is bound by instruction issue (no ncu capture of this build's launch is committed); ") +
                                       "achieved / peak / frac below measure it against the HBM roof the north star names (bound: hbm there)",
                         "kernel": ("opm_jit_trace (fused surface sequence, compiled at run time for this op program)" if jit["launches"] else
                                    "opm::trace_kernel (fused surface sequence, interpreter)"),
                         "jit": jit, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "byte_model": "93 B ray state read, 93 B written per ray kept + 41 B spot-diagram copy (position, energy, id, bounce count, valid flag: what the report reads; round 1 and early round 2 wrote all 93 B) + 36 B per hit record (DESIGN.md section 4)",
                         "algorithmic_bytes_per_launch": bytes_step / max(launches_trace, 1), "launches_per_step": launches_trace,
                         "kernel_ms_per_launch": kern_ms_step / max(launches_trace, 1), "kernel_share_of_step": kern_ms_step / (ms / args.steps),
                         "survey_8d": {"bytes_per_launch": bytes_8d / max(launches_trace, 1), "achieved": bytes_8d / kern_s / 1e9,
                                       "frac": bytes_8d / kern_s / 1e9 / hbm_peak,
                                       "model": "SURVEY 8(d): 77 B read, 69 B written per ray kept, 77 B split-off bundle, 32 B per hit record, no spot-diagram copy"},
                         "traffic_over_algorithmic": (traffic / (bytes_step / max(launches_trace, 1))) if traffic else None,
                         "unfused_equivalent_gbs": unfused,
                         "issue": ({"issue_slots_busy_pct": ncu["issue_active_pct"], "fp64_pipe_pct": ncu["fp64_pipe_pct"], "dram_pct": ncu["dram_throughput_pct"],
                                    "warp_instructions_per_launch": ncu["warp_instructions"],
                                    "ms_at_full_issue_rate": ncu["gpu_time_s"] * 1e3 * ncu["issue_active_pct"] / 100.0,
                                    "source": "profiles/r2_traffic.json (ncu --set full of the same launch)"} if ncu else None),
                         "fp64": {"peak_tflops_measured_dfma": fp64_peak,
                                  "flop_per_launch_ncu": ncu["fp64_flop_per_launch"] if ncu else None,
                                  "achieved_tflops": (ncu["fp64_flop_per_launch"] / kern_s / 1e12) if ncu else None,
                                  "frac": (ncu["fp64_flop_per_launch"] / kern_s / 1e12 / fp64_peak) if (ncu and fp64_peak) else None,
                                  "how": "executed DFMA x 2 + DMUL + DADD thread instructions of one launch counted by ncu, divided by this run's kernel time"}},
            "check": {"fluence_peak_J_m2": {str(k): v[0] for k, v in check.items() if k != "spot"},
                      "detector_energy_J": {str(k): v[1] for k, v in check.items() if k != "spot"}, "spot_valid_rays": check["spot"][0]},
            "c4_strong": c4,
            "c5_broadband": c5,
        }
        if parity_n1 is not None:
            line["parity_vs_one_gpu"] = parity_n1
    if world > 1:
        dist.barrier()
    # ---- cpu_baseline on rank 0 at N = 1 (test infrastructure timed beside the product; never on the product path)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm(n)
            arm.step()
            best = 0.0
            for _ in range(2):
                i, dt = arm.step()
                best = max(best, i / dt)
            one = arm.single_thread()
            kept = arm.trace_kept(0)
            arm.close()
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.sample,
                                    "single_thread_value": one}
            # the same workload, product against oracle (the oracle as the checker)
            with torch.cuda.stream(stream):
                line["parity_vs_oracle"] = parity_vs_oracle(ctx, n, arm.cores, kept)
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
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 strong-scaling object")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 broadband object")
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
