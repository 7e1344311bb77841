"""Node positioning from the optical axis (NodeGroup::connect_nodes(.., distance) + calc_node_positions,
nodes/node_group/analysis_raytrace.rs:92-212, optic_graph.rs:878-926) through the C ABI against the oracle: the node
isometries of a folded chain described by edge distances only, the analysis of the placed scene, the equivalence with
explicit isometries, and the error behaviour."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from oracle import oracle as orc
from helpers import RTOL, assert_map_match, assert_rays_match

pytestmark = pytest.mark.gpu
MM, NM = scenes.MM, scenes.NM


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def assert_isometries_match(psc, isos, what=""):
    for k, want in enumerate(isos):
        got = psc.node_isometry(k).matrices()
        ref = want.matrices()
        scale = max(1.0, float(np.abs(ref[1]).max()))
        for g, r, name in zip(got, ref, ("rot", "t", "inv rot", "inv t")):
            assert np.allclose(g, r, rtol=0, atol=RTOL * scale), f"{what} node {k} {name}: {g} vs {r}"


def test_isometry_new_from_view_matches_oracle():
    """utils/geom_transformation.rs:166-185"""
    rng = np.random.default_rng(5)
    for _ in range(50):
        p, d = rng.normal(size=3), rng.normal(size=3)
        up = orc.define_up_direction(d / np.linalg.norm(d))
        a, b = ob.Isometry.new_from_view(p, d, up).matrices(), orc.Isometry.from_view(p, d, up).matrices()
        for x, y in zip(a, b):
            assert np.allclose(x, y, rtol=0, atol=1e-14 * max(1.0, np.abs(p).max()))
    with pytest.raises(ob.OpossumError):
        ob.Isometry.new_from_view((0, 0, 0), (0, 0, 1), (0, 0, 2))  # up collinear with the view
    with pytest.raises(ob.OpossumError):
        ob.Isometry.new_from_view((0, 0, 0), (0, 0, 0), (0, 1, 0))


def test_folded_chain_positions_match_oracle(ctx):
    """two tilted mirrors, a wedge that deflects the axis, a beam splitter with a side arm: every node isometry, then the
    analysis of the placed scene"""
    nodes = scenes.folded_chain_by_distances()
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(1054 * NM)
    psc = scenes.to_product(ctx, nodes)
    src = {"pos": ("fibonacci_ellipse", 3000, 4 * MM, 3 * MM), "energy": ("uniform", 1.0), "wvl": [1054 * NM], "w": [1.0]}
    with pytest.raises(ob.OpossumError):  # analysed before it was placed: no effective isometry (analyzers/raytrace.rs:108-112)
        psc.raytrace(scenes.source_to_product(src).generate(ctx))
    with pytest.raises(ob.OpossumError):
        psc.node_isometry(3)
    psc.calc_node_positions(1054 * NM)
    assert_isometries_match(psc, isos)
    # the mirrors fold the path: the chain does not stay on the z axis
    t8 = psc.node_isometry(8).matrices()[1]
    assert abs(t8[0]) > 0.2 and abs(t8[1]) > 0.01
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    for mode in ("fast", "strict", "jit"):
        ctx.set_jit(2 if mode == "jit" else 0, 1)
        try:
            rays = scenes.source_to_product(src).generate(ctx)
            st = psc.raytrace(rays, strict=(mode == "strict"))
            assert st.interactions == osc.interactions, mode
            assert st.valid_out == orc.nr_of_rays(ref_out), mode
            assert_rays_match(rays.to_array(), ref_out, what=mode)
            for node in (9, 11):
                fmap, lrtb, peak, tot = osc.node_fluence(node)
                fd = psc.node_fluence(node)
                assert np.allclose(fd.lrtb, lrtb, rtol=1e-7, atol=1e-15), mode
                assert fd.total_energy == pytest.approx(tot, rel=1e-6), mode
            s, ref_spot = psc.node_spot(8), osc.node_spot(8)
            assert s.n_valid == ref_spot["n_valid"] and s.total_energy == pytest.approx(ref_spot["total_energy"], rel=RTOL)
        finally:
            ctx.set_jit(1, 200000)
    # positioning is idempotent: placed nodes are left untouched (analysis_raytrace.rs:168-170)
    psc.calc_node_positions(1054 * NM)
    assert_isometries_match(psc, isos, "second pass")


def test_distances_equal_explicit_isometries_on_axis(ctx):
    """a centred, unfolded chain: placing by distances gives the isometries one would write by hand (each node `distance`
    behind the last surface vertex of its predecessor) and the same analysis"""
    by_distance = [scenes.node("source"),
                   scenes.node("lens", distance=50 * MM, p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=scenes.NK9L_1054,
                               coat_in=("fresnel",), coat_out=("fresnel",)),
                   scenes.node("spot_diagram", distance=(200 - 2.79) * MM),
                   scenes.node("fluence_detector", distance=10 * MM)]
    a = scenes.to_product(ctx, by_distance)
    a.calc_node_positions(1000 * NM)
    b = scenes.to_product(ctx, scenes.c1_single_lens())
    for k in range(4):
        for x, y in zip(a.node_isometry(k).matrices(), b.node_isometry(k).matrices()):
            assert np.allclose(x, y, rtol=0, atol=1e-15)
    src = scenes.source_to_product(scenes.c1_source(60))
    ra, rb = src.generate(ctx), src.generate(ctx)
    sa, sb = a.raytrace(ra), b.raytrace(rb)
    assert sa == sb
    assert_rays_match(ra.to_array(), rb.to_array(), what="distance vs explicit")
    assert_map_match(a.node_fluence(3).map, b.node_fluence(3).map)


def test_positioning_with_a_paraxial_lens_off_axis_source(ctx):
    """the axis ray of a decentred, tilted source through a paraxial surface (prev_dir of ray.rs:450) and a mirror; strict build"""
    nodes = [scenes.node("source", xy=(2 * MM, 1 * MM), euler=(0.0, 0.05, 0.3)),
             scenes.node("paraxial_surface", distance=40 * MM, p=(150 * MM,), alignment=((1 * MM, 0.0, 0.0), (0.0, 0.0, 0.0))),
             scenes.node("thin_mirror", distance=70 * MM, alignment=((0.0, 0.0, 0.0), (0.0, math.radians(40.0), 0.0))),
             scenes.node("dummy", distance=30 * MM),
             scenes.node("spot_diagram", distance=25 * MM)]
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(800 * NM)
    psc = scenes.to_product(ctx, nodes)
    psc.calc_node_positions(800 * NM)
    assert_isometries_match(psc, isos)
    src = {"pos": ("grid", 30, 30, 6 * MM, 6 * MM), "energy": ("uniform", 1.0), "wvl": [800 * NM], "w": [1.0]}
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    rays = scenes.source_to_product(src).generate(ctx)
    psc.raytrace(rays, strict=True)
    assert_rays_match(rays.to_array(), ref_out, what="paraxial chain")


def test_positioning_errors(ctx):
    from opossum_b200 import scene as S
    sc = S.Scene(ctx)
    with pytest.raises(ob.OpossumError):  # the first node is the source and carries its own isometry
        sc.add_node(S.Dummy(), distance=0.1)
    sc.add_node(S.Source())
    with pytest.raises(ob.OpossumError):  # "propagation length must be finite" (ray.rs:414-418)
        sc.add_node(S.Dummy(), distance=math.inf)
    sc.add_node(S.Dummy(), distance=0.1)
    with pytest.raises(ob.OpossumError):
        sc.calc_node_positions(0.0)
    sc.calc_node_positions(1e-6)
    assert np.allclose(sc.node_isometry(1).matrices()[1], (0.0, 0.0, 0.1), rtol=0, atol=1e-16)


def test_opm_document_end_to_end(ctx):
    """a `.opm` file (tests/golden/opm/folded_chain_handwritten.opm: only the source has an isometry, edges carry distances)
    read by opossum_b200/opm_document.py, placed by calc_node_positions and analysed, against the oracle's scene built from the
    equivalent neutral description -- the mapping of the serialised properties, ports, alignment, source and edges end to end"""
    from pathlib import Path

    from opossum_b200 import opm_document as D
    doc = D.load_opm(Path(__file__).parent / "golden" / "opm" / "folded_chain_handwritten.opm")
    psc, sd, index = doc.build(ctx)
    psc.calc_node_positions(doc.axis_wavelength())
    nodes = [scenes.node("source", z=10 * MM, xy=(1 * MM, -2 * MM)),
             scenes.node("lens", distance=100 * MM, p=(515.0 * MM, -515.0 * MM, 5.0 * MM), index=scenes.HK9L, coat_in=("fresnel",),
                         coat_out=("fresnel",), ap_in=("circle", 3.5 * MM, 0.2 * MM, 0.0)),
             scenes.node("thin_mirror", distance=150 * MM, alignment=((0.0, 0.0, 0.0), (math.radians(22.5), 0.0, 0.0)),
                         coat_in=("constant_r", 0.98), coat_out=("constant_r", 0.98)),
             scenes.node("beam_splitter", distance=120 * MM, p=(0.3,)),
             scenes.node("ideal_filter", distance=50 * MM, p=(0.8,)),
             scenes.node("spot_diagram", distance=80 * MM),
             scenes.node("fluence_detector", distance=10 * MM),
             scenes.node("fluence_detector", distance=90 * MM, parent=3, parent_port=1)]
    src = {"pos": ("fibonacci_ellipse", 3000, 4 * MM, 3 * MM),
           "energy": ("general_gaussian", 2.0, (0.3 * MM, -0.2 * MM), (2.5 * MM, 2.0 * MM), 2.0, 0.3, False),
           "wvl": [1054 * NM, 527 * NM], "w": [0.7, 0.3]}
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(1054 * NM)
    assert_isometries_match(psc, isos, ".opm")
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    rays = sd.generate(ctx)
    st = psc.raytrace(rays, doc.raytrace_config())
    assert st.interactions == osc.interactions and st.valid_out == orc.nr_of_rays(ref_out)
    assert_rays_match(rays.to_array(), ref_out, what=".opm scene")
    for node in (6, 7):
        fmap, lrtb, peak, tot = osc.node_fluence(node)
        fd = psc.node_fluence(node)
        assert np.allclose(fd.lrtb, lrtb, rtol=1e-7, atol=1e-15) and fd.total_energy == pytest.approx(tot, rel=1e-6)
    assert index[doc.nodes[-1].uuid] == 7


def test_opm_document_polygon_and_stack_apertures_end_to_end(ctx):
    """the same file with the lens' circular input aperture replaced by (a) a BinaryPolygon as the reference serialises it -- points
    plus earcut's triangle indices (aperture.rs:247-251) -- and (b) a Stack of that polygon and a Gaussian: the document's analysis
    against the oracle's scene with the same triangles"""
    from pathlib import Path

    from opossum_b200 import opm_document as D
    text = (Path(__file__).parent / "golden" / "opm" / "folded_chain_handwritten.opm").read_text()
    circle = "BinaryCircle((radius: 0.0035, center: (0.0002, 0.0), aperture_type: Hole))"
    assert circle in text
    pts = [(-0.003, -0.002), (0.0025, -0.0028), (0.0032, 0.001), (0.0004, 0.0034), (-0.0028, 0.0015)]  # a convex pentagon around the axis
    fan = [(0, 1, 2), (0, 2, 3), (0, 3, 4)]
    poly = ("BinaryPolygon((points: [%s], aperture_type: Hole, triangle_indices: [%s]))"
            % (", ".join("(%r, %r)" % q for q in pts), ", ".join("[%d, %d, %d]" % t for t in fan)))
    tris = [[c for i in t for c in pts[i]] for t in fan]
    gauss = "Gaussian((sigma: (0.004, 0.003), center: (0.0, 0.0), aperture_type: Hole))"
    cases = {"polygon": (poly, ("polygon", tris, False)),
             "stack": ("Stack((apertures: [%s, %s], aperture_type: Hole))" % (gauss, poly),
                       ("stack", [orc.ap_gaussian(0.004, 0.003), orc.ap_polygon(tris)], False))}
    for name, (ron, neutral) in cases.items():
        doc = D.loads_opm(text.replace(circle, ron))
        psc, sd, _ = doc.build(ctx)
        psc.calc_node_positions(doc.axis_wavelength())
        nodes = [scenes.node("source", z=10 * MM, xy=(1 * MM, -2 * MM)),
                 scenes.node("lens", distance=100 * MM, p=(515.0 * MM, -515.0 * MM, 5.0 * MM), index=scenes.HK9L, coat_in=("fresnel",),
                             coat_out=("fresnel",), ap_in=neutral),
                 scenes.node("thin_mirror", distance=150 * MM, alignment=((0.0, 0.0, 0.0), (math.radians(22.5), 0.0, 0.0)),
                             coat_in=("constant_r", 0.98), coat_out=("constant_r", 0.98)),
                 scenes.node("beam_splitter", distance=120 * MM, p=(0.3,)),
                 scenes.node("ideal_filter", distance=50 * MM, p=(0.8,)),
                 scenes.node("spot_diagram", distance=80 * MM),
                 scenes.node("fluence_detector", distance=10 * MM),
                 scenes.node("fluence_detector", distance=90 * MM, parent=3, parent_port=1)]
        src = {"pos": ("fibonacci_ellipse", 3000, 4 * MM, 3 * MM),
               "energy": ("general_gaussian", 2.0, (0.3 * MM, -0.2 * MM), (2.5 * MM, 2.0 * MM), 2.0, 0.3, False),
               "wvl": [1054 * NM, 527 * NM], "w": [0.7, 0.3]}
        osc = scenes.to_oracle(nodes)
        osc.calc_node_positions(1054 * NM)
        ref_out = osc.raytrace(scenes.source_to_oracle(src))
        rays = sd.generate(ctx)
        st = psc.raytrace(rays, doc.raytrace_config())
        n_valid = orc.nr_of_rays(ref_out)
        assert 0 < n_valid < 6000, name  # the aperture clips some of the 3000 x 2 rays, not all (the mirror drops them from its output)
        assert st.interactions == osc.interactions and st.valid_out == n_valid, name
        assert_rays_match(rays.to_array(), ref_out, what=f".opm scene, {name} aperture")


@pytest.mark.parametrize("collimating", [False, True])
def test_off_axis_parabolic_mirror(ctx, collimating):
    """ParabolicMirror with an off-axis angle (nodes/parabolic_mirror.rs:286-320,388-421): a collimated beam onto a 45 deg off-axis
    parabola, detectors before and behind the focus, placed by distances; interpreter, strict and specialised kernels"""
    nodes = [scenes.node("source"),
             scenes.node("parabolic_mirror", distance=80 * MM, p=(250 * MM, math.radians(45.0), 0.3, 1.0, 1.0 if collimating else 0.0)),
             scenes.node("spot_diagram", distance=200 * MM),
             scenes.node("fluence_detector", distance=120 * MM)]
    src = {"pos": ("hexapolar", 14, 6 * MM), "energy": ("uniform", 1.0), "wvl": [800 * NM], "w": [1.0]}
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(800 * NM)
    psc = scenes.to_product(ctx, nodes)
    psc.calc_node_positions(800 * NM)
    assert_isometries_match(psc, isos, "oap")
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    assert orc.nr_of_rays(ref_out) == len(ref_out) > 500  # every ray is reflected and reaches the detectors
    for mode in ("fast", "strict", "jit"):
        ctx.set_jit(2 if mode == "jit" else 0, 1)
        try:
            rays = scenes.source_to_product(src).generate(ctx)
            st = psc.raytrace(rays, strict=(mode == "strict"))
            assert st.interactions == osc.interactions, mode
            assert_rays_match(rays.to_array(), ref_out, what=f"oap {mode}")
            s, ref_spot = psc.node_spot(2), osc.node_spot(2)
            assert s.n_valid == ref_spot["n_valid"] and s.beam_radius_geo == pytest.approx(ref_spot["geo"], rel=1e-8)
        finally:
            ctx.set_jit(1, 200000)


def test_align_like_node_at_distance(ctx):
    """OpticNode::align_like_node_at_distance (optic_node.rs:476-480, analysis_raytrace.rs:142-163): a detector behind a deflecting
    wedge that is not placed on the (deflected) axis but like the lens, 200 mm along the lens's z"""
    nodes = [scenes.node("source", xy=(0.5 * MM, 0.0), euler=(0.01, 0.0, 0.0)),
             scenes.node("lens", distance=50 * MM, p=(300 * MM, -300 * MM, 4 * MM), index=scenes.HK9L),
             scenes.node("wedge", distance=30 * MM, p=(6 * MM, math.radians(5.0)), index=scenes.HK9L),
             scenes.node("spot_diagram", distance=100 * MM, align_like=(1, 200 * MM)),
             scenes.node("fluence_detector", distance=20 * MM)]
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(1000 * NM)
    psc = scenes.to_product(ctx, nodes)
    psc.calc_node_positions(1000 * NM)
    assert_isometries_match(psc, isos, "align like")
    lens, spot = psc.node_isometry(1).matrices(), psc.node_isometry(3).matrices()
    assert np.allclose(spot[0], lens[0], rtol=0, atol=1e-15)                       # same orientation as the lens
    assert np.allclose(spot[1] - lens[1], 0.2 * lens[0].reshape(3, 3)[:, 2], rtol=0, atol=1e-15)  # 200 mm along the lens's z
    src = {"pos": ("grid", 25, 25, 5 * MM, 5 * MM), "energy": ("uniform", 1.0), "wvl": [1000 * NM], "w": [1.0]}
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    rays = scenes.source_to_product(src).generate(ctx)
    psc.raytrace(rays)
    assert_rays_match(rays.to_array(), ref_out, what="align like")


def test_energy_meter_node(ctx):
    """nodes/energy_meter.rs: a single-surface detector; report = total energy of the valid rays that reach it"""
    nodes = [scenes.node("source"),
             scenes.node("lens", distance=40 * MM, p=(200 * MM, -200 * MM, 5 * MM), index=scenes.HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),
             scenes.node("ideal_filter", distance=20 * MM, p=(0.6,), ap_in=("circle", 3.0 * MM, 0.5 * MM, 0.0)),
             scenes.node("energy_meter", distance=30 * MM)]
    src = {"pos": ("fibonacci_ellipse", 5000, 4 * MM, 4 * MM), "energy": ("uniform", 2.0), "wvl": [1054 * NM], "w": [1.0]}
    osc, psc = scenes.to_oracle(nodes), scenes.to_product(ctx, nodes)
    osc.calc_node_positions(1054 * NM)
    psc.calc_node_positions(1054 * NM)
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    rays = scenes.source_to_product(src).generate(ctx)
    psc.raytrace(rays)
    assert_rays_match(rays.to_array(), ref_out, what="energy meter chain")
    e = psc.node_energy(3)
    assert e == pytest.approx(osc.node_energy(3), rel=RTOL) and 0.3 < e < 1.2  # 2 J x 0.6 x (1 - R)^2, clipped by the aperture
    with pytest.raises(ob.OpossumError):
        psc.node_energy(1)  # a lens keeps no bundle
