"""Per-ray position history (`pos_hist`, ray.rs:70): the GPU log against the oracle's recording of the
reference's pushes (refract ray.rs:616, diffract :537, clipping apodize rays.rs:566), ray by ray, position by position."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import RTOL, aperture_pair, assert_rays_match, index_pair, iso_pair, make_bundle, mm, nm, node_sphere_pair, surface_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def assert_histories_match(got: list, ref_hist: orc.History, ref_rows, ref_rays=None, what="history"):
    """got[k] is the (len, 3) array of the k-th ray; ref_rows[k] the oracle row of the same ray"""
    assert len(got) == len(ref_rows), f"{what}: {len(got)} rays vs {len(ref_rows)}"
    for k, row in enumerate(ref_rows):
        ref = ref_hist.of(int(row), ref_rays)
        assert got[k].shape == ref.shape, f"{what}: ray {k} has {len(got[k])} positions, the oracle {len(ref)}"
        if len(ref):  # positions to 1e-9 relative to the size of the position vector
            assert np.allclose(got[k], ref, rtol=RTOL, atol=RTOL * np.abs(ref).max()), f"{what}: ray {k}\n{got[k]}\n{ref}"


def test_reference_kat_refract_pushes_the_old_position(ctx):
    """ray.rs:1236-1272: a ray from the origin refracted at a plane 10 mm ahead has pos_hist == [(0,0,0)]; the reflected
    clone of a refraction is handed out without a history by the bundle method (rays.rs:1014)"""
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), nm(1053.0), 1.0)
    plane = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(10.0))), coating=F.COAT_CONSTANT_R, r=0.2)
    r = ob.Rays.from_array(ctx, ray)
    r.enable_position_history(4)
    refl, st = r.refract_on_surface(plane[1], index_pair("const", 1.5)[1])
    assert st.children == 1 and r.position_history_levels == 1
    h = r.position_history(with_current=False)
    assert len(h) == 1 and np.array_equal(h[0], [[0.0, 0.0, 0.0]])
    hc = r.position_history(with_current=True)
    assert np.allclose(hc[0], [[0.0, 0.0, 0.0], [0.0, 0.0, mm(10.0)]], rtol=0, atol=1e-18)
    assert refl.position_history_levels == -1  # children of a refraction: cleared history


def test_ops_sequence_history_matches_oracle(ctx):
    """lens front (sphere, Fresnel) -> clipping aperture -> rear plane -> folding mirror (reflected rays continue, misses
    leave) -> far plane some rays miss (MissedSurfaceStrategy::Ignore) -> tight aperture; every push of every ray"""
    rays = make_bundle(3000, divergence=0.04, wvl=(nm(700.0), nm(1400.0)))
    rays["valid"][::17] = 0  # invalid rays are carried along and never push
    node = iso_pair((0.0, 0.0, mm(30.0)))
    front = node_sphere_pair(mm(120.0), node, coating=F.COAT_FRESNEL)
    rear = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(34.0))), coating=F.COAT_CONSTANT_R, r=0.02)
    mirror_iso = iso_pair((0.0, 0.0, mm(80.0)), (math.pi / 4, 0.0, 0.0))
    mirror = surface_pair(F.GEOM_PLANE, iso=mirror_iso, coating=F.COAT_CONSTANT_R, r=1.0)
    ball = node_sphere_pair(mm(-6.0), iso_pair((0.0, mm(60.0), mm(80.0)), (-math.pi / 2, 0.0, 0.0)))  # small sphere: most rays miss
    glass, vac = index_pair("sellmeier1", 6.14555251E-1, 6.56775017E-1, 1.02699346E+0, 1.45987884E-2, 2.87769588E-3,
                            1.07653051E+2, nm(300.0), nm(2000.0)), index_pair("const", 1.0)
    ap1, ap2 = aperture_pair("circle", mm(6.0), 0.0, 0.0), aperture_pair("rect", mm(8.0), mm(3.0), 0.0, 0.0)

    ref = np.ascontiguousarray(rays.copy())
    H = orc.History(len(ref))
    H.attach(ref)
    try:
        orc.rays_refract_on_surface(ref, front[0], glass[0])
        orc.rays_apodize(ref, ap1[0], node[0])
        orc.rays_refract_on_surface(ref, rear[0], vac[0])
        ch, _, _ = orc.rays_refract_on_surface(ref, mirror[0], None, refraction_intended=False)
        orc.History.detach()
        row = {int(i): k for k, i in enumerate(ref["id"])}
        rows = [row[int(i)] for i in ch["id"]]
        H2 = H.select(rows)
        ref2 = np.ascontiguousarray(ch)
        H2.attach(ref2)
        orc.rays_refract_on_surface(ref2, ball[0], glass[0], strategy=orc.MISS_IGNORE)
        orc.rays_apodize(ref2, ap2[0], mirror_iso[0])
    finally:
        orc.lib().orc_history_detach()

    g = ob.Rays.from_array(ctx, rays)
    g.enable_position_history(2)  # fewer levels than pushes: the log grows on demand
    prog = (ob.Program().refract(front[1], glass[1]).apodize(ap1[1], node[1]).refract(rear[1], vac[1]).mirror(mirror[1])
            .refract(ball[1], glass[1], strategy=F.MISS_IGNORE).apodize(ap2[1], mirror_iso[1]))
    prog.run(g)
    got_rays = g.to_array()
    assert_rays_match(got_rays, ref2, keyed=True, what="rays after the sequence")
    assert g.position_history_levels == 6
    order = np.argsort(ref2["id"], kind="stable")
    got = g.position_history(with_current=True)
    got = [got[k] for k in np.argsort(got_rays["id"], kind="stable")]
    assert_histories_match(got, H2, order, ref2, "ops sequence")
    lens = np.array([len(h) for h in got])
    assert lens.min() < lens.max()  # clipped, missing and surviving rays have histories of different lengths


def history_scene():
    """the C2 chain shape at small size with a ray-propagation visualizer at the end of both arms"""
    MM = scenes.MM
    return [
        scenes.node("source"),                                                                                           # 0
        scenes.node("lens", z=100 * MM, p=(1515.0 * MM, -1515.0 * MM, 5.0 * MM), index=scenes.HK9L, coat_in=("fresnel",),
                    coat_out=("fresnel",)),                                                                                # 1
        scenes.node("thin_mirror", z=300 * MM, euler=(math.pi / 4, 0.0, 0.0), ap_in=("circle", 7.5 * MM, 0.0, 0.0)),       # 2
        scenes.node("thin_mirror", z=300 * MM, xy=(0.0, 200 * MM), euler=(-3 * math.pi / 4, 0.0, 0.0)),                   # 3
        scenes.node("wedge", z=400 * MM, xy=(0.0, 200 * MM), p=(10 * MM, math.radians(1.0)), index=scenes.HK9L),          # 4
        scenes.node("beam_splitter", z=500 * MM, xy=(0.0, 200 * MM), p=(0.3,)),                                           # 5
        scenes.node("lens", z=600 * MM, xy=(0.0, 200 * MM), p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=scenes.HK9L,
                    ap_in=("circle", 5.0 * MM, 0.0, 0.0)),                                                                # 6
        scenes.node("ray_propagation_visualizer", z=800 * MM, xy=(0.0, 200 * MM)),                                        # 7
        scenes.node("dummy", z=600 * MM, xy=(0.0, 200 * MM), parent=5, parent_port=1, ap_in=("rect", 6 * MM, 9 * MM, 0.0, 0.0)),  # 8
        scenes.node("ray_propagation_visualizer", z=700 * MM, xy=(0.0, 200 * MM)),                                        # 9
    ]


def test_scene_history_matches_oracle(ctx):
    """RayTracingAnalyzer with OPM_SCENE_POSITION_HISTORY: the bundles the two RayPropagationVisualizer nodes keep carry the
    path of every ray (Rays::get_rays_position_history(true), rays.rs:1357-1383), main arm and split-off arm"""
    nodes, src = history_scene(), scenes.c2_source(2500)
    osc = scenes.to_oracle(nodes)
    ref_out = osc.raytrace(scenes.source_to_oracle(src), history=True)
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    st = psc.raytrace(rays, position_history=True)
    assert st.interactions == osc.interactions
    for node in (7, 9):
        ref_rays, ref_hist = osc.nodes[node].stored, osc.nodes[node].stored_hist
        stored = psc.node_rays(node)
        got_rays = stored.to_array()
        assert_rays_match(got_rays, ref_rays, keyed=True, what=f"node {node} bundle")
        got = stored.position_history(with_current=True)
        got = [got[k] for k in np.argsort(got_rays["id"], kind="stable")]
        assert_histories_match(got, ref_hist, np.argsort(ref_rays["id"], kind="stable"), ref_rays, f"node {node}")
        n_pos = np.array([len(h) for h in got])
        assert n_pos.max() >= 9 and (got_rays["valid"] != 1).any()  # clipped rays end at the aperture that stopped them
    # the traced bundle itself (main arm) keeps logging after the detector copy was taken
    got_rays = rays.to_array()
    got = rays.position_history(with_current=False)
    got = [got[k] for k in np.argsort(got_rays["id"], kind="stable")]
    assert_histories_match(got, osc.chain_hist[0], np.argsort(ref_out["id"], kind="stable"), None, "main arm output")


def test_history_is_off_by_default_and_cleared_by_new_rays(ctx):
    nodes, src = history_scene(), scenes.c2_source(800)
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    psc.raytrace(rays)
    assert rays.position_history_levels == -1 and psc.node_rays(7).position_history_levels == -1
    with pytest.raises(ob.OpossumError):
        rays.position_history()
    rays2 = scenes.source_to_product(src).generate(ctx)
    psc.raytrace(rays2, position_history=True, keep_source=True)
    assert rays2.position_history_levels == -1  # keep_source: the scene's working bundle logs, not the source
    out = psc.output_rays()
    lv = out.position_history_levels
    assert lv > 0 and psc.node_rays(7).position_history_levels > 0
    psc.raytrace(rays2, keep_source=True)      # same pooled buffers, history not asked for
    assert psc.output_rays().position_history_levels == -1 and psc.node_rays(7).position_history_levels == -1
    # a bundle that is refilled starts over (Ray::new has an empty pos_hist, ray.rs:126)
    b = ob.Rays.from_array(ctx, make_bundle(100))
    b.enable_position_history()
    plane = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(10.0))))
    b.refract_on_surface(plane[1], None, want_reflected=False)
    assert b.position_history_levels == 1
    b.upload(make_bundle(50))
    assert b.position_history_levels == 0 and all(len(h) == 0 for h in b.position_history(with_current=False))
    c = b.clone()
    assert c.position_history_levels == 0
    with pytest.raises(ob.OpossumError):
        b.compact()
