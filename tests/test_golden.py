"""Golden vectors of tests/golden/ (generated from the oracle by tests/golden/make_golden.py, see its header):
CPU: the oracle still reproduces them exactly; GPU: the CUDA path matches them within the parity contract."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

import scenes
from helpers import RTOL, assert_map_match, assert_rays_match

GOLD = Path(__file__).resolve().parent / "golden"
spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(GOLD / f"{name}.npz")
    got = mg.run_case(name)
    assert set(gold.files) == set(got)
    for k in gold.files:
        a, b = gold[k], np.asarray(got[k])
        if a.dtype.fields:  # ray records: every field bit for bit
            for f in a.dtype.names:
                assert np.array_equal(a[f], b[f]), (name, k, f)
        else:
            assert np.array_equal(a, b, equal_nan=True), (name, k)


def test_oracle_reproduces_golden_ghost_focus():
    gold = np.load(GOLD / "c3_ghost_focus.npz")
    got = mg.run_ghost_focus()
    for k in gold.files:
        assert np.array_equal(gold[k], np.asarray(got[k]), equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_cuda_matches_golden(name, strict):
    import opossum_b200 as ob
    nodes_fn, src_fn, det = mg.CASES[name]
    gold = np.load(GOLD / f"{name}.npz")
    ctx = ob.Context(0)
    try:
        sc = scenes.to_product(ctx, nodes_fn())
        rays = scenes.source_to_product(src_fn()).generate(ctx)
        st = sc.raytrace(rays, strict=strict)
        assert st.interactions == int(gold["interactions"])
        assert_rays_match(rays.to_array(), gold["rays"], what=name)
        for node in det["fluence"]:
            fd = sc.node_fluence(node, *mg.MAP)
            assert np.allclose(fd.lrtb, gold[f"lrtb_{node}"], rtol=RTOL, atol=0.0)
            if tuple(fd.lrtb) == tuple(gold[f"lrtb_{node}"]):
                assert_map_match(fd.map, gold[f"map_{node}"])
            same = ob.fluence_binning(ctx, sc.node_hitmaps(node, 0), *mg.MAP, lrtb=tuple(gold[f"lrtb_{node}"]))
            assert_map_match(same.map, gold[f"map_{node}"], what=f"{name} node {node} on the golden grid")
            assert fd.total_energy == pytest.approx(float(gold[f"total_{node}"]), rel=RTOL)
        if det["spot"] is not None:
            s = sc.node_spot(det["spot"])
            g = gold["spot"]
            assert s.n_valid == int(g[0]) and s.total_energy == pytest.approx(g[1], rel=RTOL)
            assert np.allclose(s.energy_centroid[:2], g[2:4], rtol=RTOL, atol=1e-15)
            assert s.beam_radius_geo == pytest.approx(g[4], rel=RTOL) and s.energy_beam_radius_rms == pytest.approx(g[5], rel=RTOL)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_matches_golden_ghost_focus():
    import opossum_b200 as ob
    gold = np.load(GOLD / "c3_ghost_focus.npz")
    ctx = ob.Context(0)
    try:
        sc = scenes.to_product(ctx, scenes.c3_telescope())
        rays = scenes.source_to_product(scenes.c3_source(400)).generate(ctx)
        st = sc.ghostfocus(rays, 2)
        assert st.interactions == int(gold["interactions"]) and st.passes == int(gold["n_passes"])
        for p in range(int(gold["n_passes"])):
            counts, energy, nb, nr = sc.gf_pass_tally(p, 6)
            assert nb == int(gold[f"pass{p}_n_bundles"]) and nr == int(gold[f"pass{p}_n_rays"])
            assert np.array_equal(counts.astype(np.int64), gold[f"pass{p}_counts"])
            assert np.allclose(energy, gold[f"pass{p}_energy"], rtol=RTOL, atol=1e-30)
        for node in range(4):
            for surf in (0, 1):
                ref = gold["peaks"][node, surf]
                got = sc.node_peak_fluence(node, surf)
                assert (got == pytest.approx(ref, rel=RTOL)) if np.isfinite(ref) else not np.isfinite(got)
    finally:
        ctx.close()
