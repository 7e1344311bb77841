"""profiles/r2_traffic.json "trace" entry from the raw page of an `ncu --set full` capture of the C2 trace launch:
    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python profiles/make_traffic.py raw.csv <algorithmic bytes per launch of the bench line> [label of the capture]
(bench.py reads the entry only while it still describes the build it measures: same rays, launches per step and algorithmic bytes)"""
import csv
import json
import sys
from pathlib import Path

raw, alg = sys.argv[1], float(sys.argv[2])
label = sys.argv[3] if len(sys.argv) > 3 else raw
rows = list(csv.reader(open(raw)))
h, v = rows[0], rows[2]


def m(name):
    return float(v[h.index(name)].replace(",", ""))


unit = rows[1]


def scaled(name):  # values the page prints with a unit prefix (ms, Gbyte ...)
    val, u = m(name), unit[h.index(name)].lower()
    f = {"": 1.0, "s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    return val * f.get(u, 1.0)


t = scaled("gpu__time_duration.sum")
cycles = m("smsp__cycles_elapsed.max") if "smsp__cycles_elapsed.max" in h else m("sm__cycles_elapsed.max")
per = lambda op: m(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cycles
dfma, dmul, dadd = per("dfma"), per("dmul"), per("dadd")
entry = {
    "kernel": "opm_jit_trace",
    "workload": "C2, 10^7 rays per GPU, the one trace launch of a bench step",
    "workload_rays_per_gpu": 10000000, "launches_per_step": 1, "gpu_time_s": t,
    "dram_bytes_read": scaled("dram__bytes_read.sum"), "dram_bytes_write": scaled("dram__bytes_write.sum"),
    "algorithmic_bytes_per_launch": alg,
    "warp_instructions": m("smsp__inst_executed.sum"),
    "issue_active_pct": m("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "fp64_pipe_pct": m("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "dram_throughput_pct": m("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "thread_dfma": dfma, "thread_dmul": dmul, "thread_dadd": dadd, "fp64_flop_per_launch": 2 * dfma + dmul + dadd,
    "registers": m("launch__registers_per_thread"), "warps_per_sm": m("sm__warps_active.avg.per_cycle_active"),
    "stall_long_scoreboard_per_issue": m("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    "stall_wait_per_issue": m("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    "lts_sectors_data_ecc": m("lts__t_sectors_data_ecc.sum"),
    "source": label,
}
p = Path(__file__).resolve().parent / "r2_traffic.json"
doc = json.loads(p.read_text()) if p.exists() else {}
doc["trace"] = entry
p.write_text(json.dumps(doc, indent=1))
print(json.dumps(entry, indent=1))
