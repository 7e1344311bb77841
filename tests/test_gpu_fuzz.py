"""Randomised scenes (seeded, reproducible) through the scene layer: GPU against the oracle for chains the hand-written
tests do not spell out -- random node sequences, glasses, coatings, apertures, decentres and tilts, all three kernels
(interpreter, literal `strict` build, run-time specialised)."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from oracle import oracle as orc
from helpers import RTOL, assert_map_match, assert_rays_match

pytestmark = pytest.mark.gpu
MM, NM = scenes.MM, scenes.NM

GLASSES = [scenes.NK9L_1054, scenes.HK9L, scenes.HZF52, ("const", 1.7), ("conrady", 1.49, 4.0e-9, 1.2e-22, 300 * NM, 2500 * NM)]
COATS = [None, ("ideal_ar",), ("fresnel",), ("constant_r", 0.04), ("constant_r", 0.3)]


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def random_scene(seed: int):
    rng = np.random.default_rng(1000 + seed)
    pick = lambda seq: seq[int(rng.integers(len(seq)))]  # noqa: E731
    beam = float(rng.uniform(3.0, 9.0))  # mm, half size of the source

    def aperture(scale=1.0):
        kind = pick(["none", "none", "circle", "rect", "gaussian"])
        r = beam * scale * float(rng.uniform(0.55, 1.3))
        cx, cy = (float(rng.uniform(-0.8, 0.8)) for _ in range(2))
        if kind == "circle":
            return ("circle", r * MM, cx * MM, cy * MM)
        if kind == "rect":
            return ("rect", 2 * r * MM, 1.6 * r * MM, cx * MM, cy * MM)
        if kind == "gaussian":
            return ("gaussian", 1.5 * r * MM, 1.1 * r * MM, cx * MM, cy * MM)
        return None

    def align():
        if rng.random() < 0.5:
            return {}
        t = tuple(float(rng.uniform(-0.4, 0.4)) * MM for _ in range(2)) + (0.0,)
        e = tuple(float(rng.uniform(-0.01, 0.01)) for _ in range(2)) + (float(rng.uniform(-0.5, 0.5)),)
        return {"alignment": (t, e)}

    nodes, z = [scenes.node("source", **({"ap_out": aperture()} if rng.random() < 0.3 else {}))], 0.0
    has_splitter = False
    for _ in range(int(rng.integers(3, 8))):
        z += float(rng.uniform(15.0, 60.0))
        kind = pick(["lens", "lens", "cylindric_lens", "wedge", "dummy", "ideal_filter", "paraxial_surface", "beam_splitter"])
        kw = {k: v for k, v in (("ap_in", aperture()), ("ap_out", aperture(1.1))) if v is not None}
        kw.update(align())
        if kind in ("lens", "cylindric_lens"):
            r1 = pick([math.inf, 1.0, -1.0]) * float(rng.uniform(150.0, 900.0))
            r2 = pick([math.inf, 1.0, -1.0]) * float(rng.uniform(150.0, 900.0))
            p = (r1 * MM if math.isfinite(r1) else r1, r2 * MM if math.isfinite(r2) else r2, float(rng.uniform(3.0, 8.0)) * MM)
            nodes.append(scenes.node(kind, z=z * MM, p=p, index=pick(GLASSES), coat_in=pick(COATS), coat_out=pick(COATS), **kw))
        elif kind == "wedge":
            nodes.append(scenes.node("wedge", z=z * MM, p=(float(rng.uniform(3.0, 9.0)) * MM, math.radians(float(rng.uniform(-2.0, 2.0)))),
                                     index=pick(GLASSES), coat_in=pick(COATS), coat_out=pick(COATS), **kw))
        elif kind == "ideal_filter":
            nodes.append(scenes.node("ideal_filter", z=z * MM, p=(float(rng.uniform(0.2, 1.0)),), **kw))
        elif kind == "paraxial_surface":
            nodes.append(scenes.node("paraxial_surface", z=z * MM, p=(pick([1.0, -1.0]) * float(rng.uniform(200.0, 900.0)) * MM,), **kw))
        elif kind == "beam_splitter" and not has_splitter:
            has_splitter = True
            nodes.append(scenes.node("beam_splitter", z=z * MM, p=(float(rng.uniform(0.1, 0.9)),), **kw))
        else:
            nodes.append(scenes.node("dummy", z=z * MM, **kw))
    z += float(rng.uniform(20.0, 80.0))
    nodes.append(scenes.node("spot_diagram", z=z * MM))
    nodes.append(scenes.node("fluence_detector", z=(z + 5.0) * MM))
    if has_splitter:
        sp = next(i for i, n in enumerate(nodes) if n["kind"] == "beam_splitter")
        nodes.append(scenes.node("dummy", z=(nodes[sp]["t"][2] / MM + 10.0) * MM, parent=sp, parent_port=1, ap_in=aperture()) if rng.random() < 0.5
                     else scenes.node("ideal_filter", z=(nodes[sp]["t"][2] / MM + 10.0) * MM, parent=sp, parent_port=1, p=(0.5,)))
        nodes.append(scenes.node("fluence_detector", z=(nodes[sp]["t"][2] / MM + 30.0) * MM))
    n = int(rng.integers(1500, 4000))
    pos = pick([("fibonacci_ellipse", n, beam * MM, 0.8 * beam * MM), ("fibonacci_rectangle", n, 2 * beam * MM, 1.5 * beam * MM),
                ("grid", 50, 47, 2 * beam * MM, 1.7 * beam * MM), ("hexapolar", 12, beam * MM)])
    energy = pick([("uniform", 1.0), ("general_gaussian", 2.0, (0.3 * MM, -0.2 * MM), (0.6 * beam * MM, 0.5 * beam * MM), 2.0, 0.3, False)])
    wvl = sorted(float(w) for w in rng.uniform(500.0, 1600.0, int(rng.integers(1, 4))))
    w = rng.uniform(0.2, 1.0, len(wvl))
    src = {"pos": pos, "energy": energy, "wvl": [x * NM for x in wvl], "w": list(w / w.sum())}
    return nodes, src


@pytest.mark.parametrize("seed", range(16))
def test_random_scene_matches_oracle(ctx, seed):
    nodes, src = random_scene(seed)
    osc = scenes.to_oracle(nodes)
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    fluence_nodes = [i for i, n in enumerate(nodes) if n["kind"] == "fluence_detector"]
    spot = next(i for i, n in enumerate(nodes) if n["kind"] == "spot_diagram")
    for mode in ("fast", "strict", "jit"):
        ctx.set_jit(2 if mode == "jit" else 0, 1)
        try:
            psc = scenes.to_product(ctx, nodes)
            rays = scenes.source_to_product(src).generate(ctx)
            st = psc.raytrace(rays, strict=(mode == "strict"))
            what = f"seed {seed} {mode}"
            assert st.interactions == osc.interactions, what
            assert st.valid_out == orc.nr_of_rays(ref_out), what
            assert_rays_match(rays.to_array(), ref_out, what=what)
            if orc.nr_of_rays(osc.nodes[spot].stored) > 0:
                ref_spot, s = osc.node_spot(spot), psc.node_spot(spot)
                assert s.n_valid == ref_spot["n_valid"], what
                assert s.total_energy == pytest.approx(ref_spot["total_energy"], rel=RTOL), what
                assert s.beam_radius_geo == pytest.approx(ref_spot["geo"], rel=1e-8), what
            for node in fluence_nodes:
                if len(osc.nodes[node].s[0].merged_hits()) == 0:
                    with pytest.raises(ob.OpossumError):
                        psc.node_fluence(node)
                    continue
                fmap, lrtb, peak, tot = osc.node_fluence(node)
                fd = psc.node_fluence(node)
                assert np.allclose(fd.lrtb, lrtb, rtol=RTOL, atol=1e-18), what
                if fd.lrtb == lrtb:
                    assert_map_match(fd.map, fmap, what=what)
                # the product's hit map on the oracle's grid: bins one to one in every kernel build
                same = ob.fluence_binning(ctx, psc.node_hitmaps(node, 0), 100, 83, lrtb=lrtb)
                assert_map_match(same.map, fmap, what=what + " (oracle's grid)")
                assert fd.total_energy == pytest.approx(tot, rel=RTOL), what
            if any(n["kind"] == "beam_splitter" for n in nodes):  # the second output, traced through its side chain
                sp = [n["kind"] for n in nodes].index("beam_splitter")
                assert_rays_match(psc.node_rays(sp).to_array(), osc.chain_out[_side_start(nodes, sp)], what=what + " side arm")
        finally:
            ctx.set_jit(1, 200000)


def _side_start(nodes, sp):
    return next(i for i, n in enumerate(nodes) if n.get("parent") == sp and n.get("parent_port") == 1)
