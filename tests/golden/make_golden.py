"""Generates the golden vectors in this directory FROM THE ORACLE (oracle/opm_oracle.c + oracle/scene.py).

The reference is Rust and cannot be built or imported in this environment (no rustc/cargo, crates not vendored), so
these vectors do not come from the reference itself: they freeze the oracle -- which is pinned by the reference's own
known-answer tests in tests/test_oracle_kats.py -- so that (a) the oracle cannot drift unnoticed and (b) the CUDA path
can be checked on a box where only the fixtures travel.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {
    "c1_single_lens": (scenes.c1_single_lens, lambda: scenes.c1_source(24), {"spot": 2, "fluence": [3]}),
    "c2_laser_chain": (scenes.c2_laser_chain, lambda: scenes.c2_source(1500), {"spot": 8, "fluence": [9, 10]}),
    "c4_fluence": (scenes.c4_fluence, lambda: scenes.c4_source(30), {"spot": None, "fluence": [2]}),
    "c5_broadband": (scenes.c5_broadband, lambda: scenes.c5_source(120, 8), {"spot": None, "fluence": [4]}),
}
MAP = (40, 33)


def run_case(name):
    nodes_fn, src_fn, det = CASES[name]
    osc = scenes.to_oracle(nodes_fn())
    out = osc.raytrace(scenes.source_to_oracle(src_fn()))
    d = {"rays": out, "interactions": np.int64(osc.interactions)}
    for node in det["fluence"]:
        fmap, lrtb, peak, tot = osc.node_fluence(node, *MAP)
        d[f"map_{node}"], d[f"lrtb_{node}"], d[f"peak_{node}"], d[f"total_{node}"] = fmap, np.array(lrtb), peak, tot
    if det["spot"] is not None:
        s = osc.node_spot(det["spot"])
        d["spot"] = np.array([s["n_valid"], s["total_energy"], s["centroid"][0], s["centroid"][1], s["geo"], s["rms"]])
    return d


def run_ghost_focus(max_bounces=2, n=400):
    osc = scenes.to_oracle(scenes.c3_telescope())
    passes = osc.ghostfocus(scenes.source_to_oracle(scenes.c3_source(n)), max_bounces)
    d = {"interactions": np.int64(osc.interactions), "n_passes": np.int64(len(passes))}
    for p, bundles in enumerate(passes):
        d[f"pass{p}_n_bundles"] = np.int64(len(bundles))
        counts, energy = np.zeros(6, dtype=np.int64), np.zeros(6)
        for b in bundles:
            v = b.rays[b.rays["valid"] == 1]
            for k in np.unique(v["bounces"]):
                counts[k] += int((v["bounces"] == k).sum())
                energy[k] += float(v["e"][v["bounces"] == k].sum())
        d[f"pass{p}_counts"], d[f"pass{p}_energy"] = counts, energy
        d[f"pass{p}_n_rays"] = np.int64(sum(len(b.rays) for b in bundles))
    d["peaks"] = np.array([[osc.nodes[n_].s[s].peak_seen for s in (0, 1)] for n_ in range(4)])
    return d


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(OUT / f"{name}.npz", **run_case(name))
    np.savez_compressed(OUT / "c3_ghost_focus.npz", **run_ghost_focus())
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")
