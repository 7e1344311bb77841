#!/bin/bash
# accumulation variants of the shared-memory Binning kernel on the C2 detectors (profiles/binning_probe.py)
mkdir -p gpurun_out/binning
for v in 0 2 4 11 12 14 21 24 31 32 34 44; do
  OPM_B200_BINNING_FIXED=$v timeout 300 python profiles/binning_probe.py gpurun_out/binning/map_$v.npz 2>gpurun_out/binning/err_$v.log | tail -1
done > gpurun_out/binning/results.jsonl
python - <<'PY' >> gpurun_out/binning/results.jsonl
import numpy as np, json
ref = np.load("gpurun_out/binning/map_0.npz")
for v in (2, 4, 11, 12, 14, 21, 24, 31, 32, 34, 44):
    try:
        m = np.load(f"gpurun_out/binning/map_{v}.npz")
    except Exception as e:
        print(json.dumps({"variant": v, "error": str(e)})); continue
    out = {"variant": v}
    for k in ref.files:
        a, b = ref[k], m[k]
        out[f"node{k}_max_abs_diff_over_peak"] = float(np.abs(a - b).max() / a.max())
        nz = a > 1e-6 * a.max()
        out[f"node{k}_max_rel_diff_bins_above_1e-6_peak"] = float((np.abs(a - b)[nz] / a[nz]).max())
    print(json.dumps(out))
PY
cat gpurun_out/binning/results.jsonl
