"""End-to-end parity of the BASELINE.json configs (reduced ray counts) through the scene layer and the fused kernel:
product (C++ scene layer -> C ABI -> CUDA) vs the oracle's restatement of the reference's node / analyzer flow."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from oracle import oracle as orc
from helpers import RTOL, assert_hits_match, assert_map_match, assert_rays_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def run_both(ctx, nodes, src, **kw):
    osc = scenes.to_oracle(nodes)
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    st = psc.raytrace(rays, **kw)
    return osc, ref_out, psc, rays, st


def check_detector(osc, psc, node, nx=100, ny=83):
    """FluenceDetector report vs the oracle.  (1) The product's own data-dependent map range agrees with the oracle's to 1e-9.
    (2) The product's hit map binned into the ORACLE's grid (explicit range) is compared bin by bin at 1e-9 relative -- so every
    kernel build is checked one to one, also when its bounding box differs from the oracle's in the last bit.  (3) When the two
    boxes are bit-identical the product's own report is compared the same way."""
    fmap, lrtb, peak, tot = osc.node_fluence(node, nx, ny)
    fd = psc.node_fluence(node, nx, ny)
    assert np.allclose(fd.lrtb, lrtb, rtol=RTOL, atol=0.0)
    same = ob.fluence_binning(psc.ctx, psc.node_hitmaps(node, 0), nx, ny, lrtb=lrtb)
    assert_map_match(same.map, fmap, what=f"node {node} map on the oracle's grid")
    assert same.peak == pytest.approx(peak, rel=RTOL) and same.total_energy == pytest.approx(tot, rel=RTOL)
    if fd.lrtb == lrtb:
        assert_map_match(fd.map, fmap, what=f"node {node} map")
        assert fd.peak == pytest.approx(peak, rel=RTOL) and fd.total_energy == pytest.approx(tot, rel=RTOL)
    else:  # another grid (last-bit different range): the integral is grid independent, the peak moves with the bin edges
        assert fd.total_energy == pytest.approx(tot, rel=RTOL)
        assert fd.peak == pytest.approx(peak, rel=1e-6)


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_c1_single_lens(ctx, strict):
    """configs[0]: 10^4-ray grid, lens, spot diagram + 100x83 fluence map; full per-ray compare"""
    osc, ref_out, psc, rays, st = run_both(ctx, scenes.c1_single_lens(), scenes.c1_source(), strict=strict, record_all_hits=True)
    assert st.interactions == osc.interactions == 40000
    assert_rays_match(rays.to_array(), ref_out, what="C1 final bundle")
    ref_spot, s = osc.node_spot(2), psc.node_spot(2)
    assert s.n_valid == ref_spot["n_valid"] and s.total_energy == pytest.approx(ref_spot["total_energy"], rel=RTOL)
    assert np.allclose(s.energy_centroid[:2], ref_spot["centroid"][:2], rtol=RTOL, atol=1e-15)
    assert s.beam_radius_geo == pytest.approx(ref_spot["geo"], rel=RTOL)
    assert s.energy_beam_radius_rms == pytest.approx(ref_spot["rms"], rel=RTOL)
    check_detector(osc, psc, 3)
    if strict:  # axis-aligned scene + strict kernel: the map range is bit-identical to the oracle's
        assert psc.node_fluence(3).lrtb == osc.node_fluence(3)[1]
    for node in (1, 2, 3):  # hit maps of every surface (RECORD_ALL_HITS): lens front/rear, detectors
        for surf in (0, 1):
            ref_hits = osc.nodes[node].s[surf].merged_hits()
            maps = psc.node_hitmaps(node, surf)
            assert (len(maps) > 0) == (len(ref_hits) > 0)
            if maps:
                assert_hits_match(maps[0].download(), ref_hits)


def test_c2_specialised_kernel_spot_sums_and_lean_copy():
    """C2 through the run-time specialised kernel: the SpotDiagram's copy holds position / energy / valid flag only
    (OPM_SCENE_LEAN_SPOT_DATA) and its phase-1 sums come out of the trace launch (no first pass over the copy).  Every statistic
    against the oracle, and against the two-pass path on a full copy."""
    import os
    src = scenes.c2_source(20000)
    os.environ["OPM_B200_JIT_SPOT"] = "1"  # read when a context is created
    ctx = ob.Context(0)
    del os.environ["OPM_B200_JIT_SPOT"]
    osc = scenes.to_oracle(scenes.c2_laser_chain())
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    ref = osc.node_spot(8)
    stats = {}
    try:
        for mode in ("two_pass", "two_pass_lean", "fused_full", "fused_lean"):
            ctx.set_jit(2 if mode.startswith("fused") else 0, 1)
            psc = scenes.to_product(ctx, scenes.c2_laser_chain())
            rays = scenes.source_to_product(src).generate(ctx)
            n0 = ctx.launch_count
            st = psc.raytrace(rays, lean_spot_data=mode.endswith("lean"))
            assert st.interactions == osc.interactions
            assert_rays_match(rays.to_array(), ref_out, what=f"C2 main chain ({mode})")
            n1 = ctx.launch_count
            s = psc.node_spot(8)
            stats[mode] = (s, ctx.launch_count - n1)
            assert s.n_valid == ref["n_valid"] and s.n_total == len(ref_out) and s.bounce_lvl == 0
            assert s.total_energy == pytest.approx(ref["total_energy"], rel=RTOL)
            assert np.allclose(s.energy_centroid[:2], ref["centroid"][:2], rtol=RTOL, atol=1e-15)
            assert s.beam_radius_geo == pytest.approx(ref["geo"], rel=RTOL)
            assert s.energy_beam_radius_rms == pytest.approx(ref["rms"], rel=RTOL)
            if mode.endswith("lean"):  # the copy: positions, energies, flags, ids and bounce counts of the full copy
                full = osc.nodes[8].stored
                got = psc.node_rays(8).to_array()
                assert len(got) == len(full)
                for f in ("pos", "e", "valid", "id", "bounces"):
                    assert np.allclose(got[f], full[f], rtol=RTOL, atol=1e-300), f
                assert psc.node_rays(8).bounce_lvl() == 0
            check_detector(osc, psc, 9)
    finally:
        jit = ctx.jit_stats()
        psc = rays = None  # owned by the context
        import gc
        gc.collect()
        ctx.close()
    assert jit["compile_failures"] == 0
    a, b = stats["two_pass"][0], stats["fused_lean"][0]
    for f in ("centroid", "energy_centroid"):
        assert np.allclose(list(getattr(a, f)), list(getattr(b, f)), rtol=1e-12, atol=1e-18)
    assert a.beam_radius_rms == pytest.approx(b.beam_radius_rms, rel=1e-12)
    assert stats["fused_lean"][1] < stats["two_pass"][1]  # the pass over the copy that takes the sums is gone


@pytest.mark.parametrize("mode", ["fused", "per_surface", "strict"])
def test_c2_laser_chain(ctx, mode):
    """configs[1] at 20k rays: lens, two 45 deg mirrors (rays clipped at the first leave the bundle at the second), wedge
    (Sellmeier), beam splitter with a detector on the second output, lens, filter, spot diagram, fluence detector"""
    kw = {"fused": {}, "per_surface": {"per_surface": True}, "strict": {"strict": True}}[mode]
    osc, ref_out, psc, rays, st = run_both(ctx, scenes.c2_laser_chain(), scenes.c2_source(20000), **kw)
    assert st.interactions == osc.interactions
    got = rays.to_array()  # removed rays (dropped at the mirrors) are skipped by download
    assert len(got) == len(ref_out) == st.alive_out and st.valid_out == orc.nr_of_rays(ref_out)
    assert 0 < len(ref_out) < 20000
    assert_rays_match(got, ref_out, what="C2 main chain")
    split = psc.node_rays(5).to_array()  # the splitter's second output, traced in place through its side chain (node 10)
    assert_rays_match(split, osc.chain_out[10], what="C2 split-off bundle after the side chain")
    ref_spot, s = osc.node_spot(8), psc.node_spot(8)
    assert s.n_valid == ref_spot["n_valid"] and s.total_energy == pytest.approx(ref_spot["total_energy"], rel=RTOL)
    assert np.allclose(s.energy_centroid[:2], ref_spot["centroid"][:2], rtol=RTOL, atol=1e-15)
    assert s.beam_radius_geo == pytest.approx(ref_spot["geo"], rel=RTOL)
    check_detector(osc, psc, 9)
    check_detector(osc, psc, 10)
    if mode == "per_surface":
        assert ctx.launch_count > 40


def test_beam_splitter_with_two_inputs(ctx):
    """BeamSplitter::analyze_raytrace with both inputs (nodes/beam_splitter/mod.rs:158-293).  The reference's known answer
    (analysis_raytrace.rs:149-181): 1 J on input 1, 0.5 J on input 2, ratio 0.6 -> 0.8 J and 0.7 J out; then two bundles through a
    splitter with apertures, ray by ray against the oracle; one missing input."""
    MM, NM = scenes.MM, scenes.NM
    nodes = [scenes.node("source"), scenes.node("beam_splitter", z=10 * MM, p=(0.6,))]
    osc, psc = scenes.to_oracle(nodes), scenes.to_product(ctx, nodes)

    def one_ray(e):
        return orc.rays_collimated_with_spectrum(np.zeros((1, 3)), np.array([e]), [1053 * NM], [1.0])
    a, b = one_ray(1.0), one_ray(0.5)
    o1, o2 = psc.beam_splitter_two_inputs(1, ob.Rays.from_array(ctx, a), ob.Rays.from_array(ctx, b))
    assert o1.to_array()["e"].sum() == pytest.approx(0.8, abs=1e-15) and o2.to_array()["e"].sum() == pytest.approx(0.7, abs=1e-15)
    r1, r2 = osc.beam_splitter_two_inputs(1, a, b)
    assert r1["e"].sum() == pytest.approx(0.8, abs=1e-15) and r2["e"].sum() == pytest.approx(0.7, abs=1e-15)
    # bundles, apertures on both ports, a tilted splitter
    nodes = [scenes.node("source"),
             scenes.node("beam_splitter", z=40 * MM, euler=(0.3, 0.0, 0.0), p=(0.35,), ap_in=("circle", 6 * MM, 0.0, 0.0),
                         ap_out=("rect", 9 * MM, 7 * MM, 0.5 * MM, 0.0))]
    osc, psc = scenes.to_oracle(nodes), scenes.to_product(ctx, nodes)
    in1 = scenes.source_to_oracle(scenes.c3_source(700))
    in2 = scenes.source_to_oracle({"pos": ("grid", 20, 20, 11 * MM, 11 * MM), "energy": ("uniform", 0.4), "wvl": [1000 * NM], "w": [1.0]})
    r1, r2 = osc.beam_splitter_two_inputs(1, in1, in2)
    o1, o2 = psc.beam_splitter_two_inputs(1, ob.Rays.from_array(ctx, in1), ob.Rays.from_array(ctx, in2))
    assert_rays_match(o1.to_array(), r1, what="output 1")  # order: input 1's rays, then what input 2 split off (Rays::merge appends)
    assert_rays_match(o2.to_array(), r2, what="output 2")
    assert (r1["valid"] == 0).sum() > 10  # the apertures clip
    r1, r2 = osc.beam_splitter_two_inputs(1, None, in2)
    o1, o2 = psc.beam_splitter_two_inputs(1, None, ob.Rays.from_array(ctx, in2))
    assert_rays_match(o1.to_array(), r1, what="output 1, no input 1")
    assert_rays_match(o2.to_array(), r2, what="output 2, no input 1")


def test_c2_single_surface_nodes_and_limits(ctx):
    """dummy, paraxial surface, cylindric lens, parabolic mirror + the limits after every node
    (bounces > max, refractions >= max, node_group/analysis_raytrace.rs:20-27)"""
    MM = scenes.MM
    nodes = [scenes.node("source", ap_out=("rect", 18 * MM, 14 * MM, 0.0, 0.0)),
             scenes.node("dummy", z=20 * MM, ap_in=("circle", 9 * MM, 0.0, 0.0), ap_out=("gaussian", 6 * MM, 6 * MM, 0.0, 0.0)),
             scenes.node("cylindric_lens", z=40 * MM, p=(300 * MM, float("-inf"), 4 * MM), index=scenes.HK9L),
             scenes.node("paraxial_surface", z=80 * MM, p=(400 * MM,)),
             scenes.node("parabolic_mirror", z=200 * MM, p=(500 * MM,), alignment=((0.0, 0.0, 0.0), (0.01, 0.0, 0.0))),
             scenes.node("fluence_detector", z=100 * MM, euler=(3.14159, 0.0, 0.0))]
    src = {"pos": ("fibonacci_rectangle", 5000, 20 * MM, 16 * MM), "energy": ("uniform", 1.0), "wvl": [800 * scenes.NM], "w": [1.0]}
    for max_refr in (1000, 3):
        cfg_o = __import__("oracle.scene", fromlist=["x"]).RayTraceConfig(max_number_of_refractions=max_refr)
        osc = scenes.to_oracle(nodes)
        ref_out = osc.raytrace(scenes.source_to_oracle(src), cfg_o)
        psc = scenes.to_product(ctx, nodes)
        rays = scenes.source_to_product(src).generate(ctx)
        st = psc.raytrace(rays, ob.RayTraceConfig(max_number_of_refractions=max_refr))
        assert st.interactions == osc.interactions
        assert_rays_match(rays.to_array(), ref_out, what=f"single-surface chain max_refr={max_refr}")
        if max_refr == 1000:
            assert orc.nr_of_rays(ref_out) > 1000
            check_detector(osc, psc, 5, 64, 48)
        else:
            assert orc.nr_of_rays(ref_out) == 0  # cylinder front + rear + paraxial = 3 refractions -> `>= 3` invalidates


def test_c4_fluence_map(ctx):
    """configs[3] reduced: grid source, lens, detector, 512x512 Binning map with the data-dependent bounding box"""
    osc, ref_out, psc, rays, st = run_both(ctx, scenes.c4_fluence(), scenes.c4_source(300), strict=True, snapshots=False)
    assert st.interactions == osc.interactions == 3 * 90000
    check_detector(osc, psc, 2, 512, 512)
    assert psc.node_rays(2) is None  # OPM_SCENE_NO_SNAPSHOTS


def test_c5_broadband(ctx):
    """configs[4] reduced: 2000 positions x 16 wavelengths (position-major) through Sellmeier + Schott prisms"""
    src = scenes.c5_source(2000, 16)
    osc, ref_out, psc, rays, st = run_both(ctx, scenes.c5_broadband(), src)
    assert st.interactions == osc.interactions and len(ref_out) == 32000
    assert_rays_match(rays.to_array(), ref_out, what="C5")
    wv = np.unique(ref_out["wvl"])
    assert len(wv) == 16
    # dispersion is real: the mean detector position depends on wavelength
    y_by_w = [ref_out["pos"][ref_out["wvl"] == w][:, 1].mean() for w in wv]
    assert np.ptp(y_by_w) > 1e-6
    check_detector(osc, psc, 4, 128, 96)


@pytest.mark.parametrize("max_bounces", [1, 2, 3])
def test_c3_ghost_focus(ctx, max_bounces):
    """configs[2] reduced: 3000 seed rays, bounce orders up to 3 incl. the reference's cache re-injection at pass 3
    (SURVEY appendix C-6).  Per pass: number of bundles, per-bounce-order valid counts (bit-exact) and energies,
    every collected bundle ray by ray (keyed by seed id); per surface: peak fluence of the per-bundle 101x101 maps."""
    nodes = scenes.c3_telescope()
    src = scenes.c3_source(3000)
    osc = scenes.to_oracle(nodes)
    ref_passes = osc.ghostfocus(scenes.source_to_oracle(src), max_bounces)
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    st = psc.ghostfocus(rays, max_bounces)
    assert st.passes == max_bounces + 1 and st.interactions == osc.interactions
    for p, ref_bundles in enumerate(ref_passes):
        got = psc.gf_pass_bundles(p)
        assert len(got) == len(ref_bundles), f"pass {p}"
        counts, energy, nb, nr = psc.gf_pass_tally(p, 6)
        ref_counts = np.zeros(6, dtype=np.uint64)
        ref_energy = np.zeros(6)
        for g, r in zip(got, ref_bundles):
            assert_rays_match(g.to_array(), r.rays, keyed=True, what=f"pass {p} bundle {r.uuid}", e_floor=1e-25)
            v = r.rays[r.rays["valid"] == 1]
            for b in np.unique(v["bounces"]):
                ref_counts[b] += int((v["bounces"] == b).sum())
                ref_energy[b] += float(v["e"][v["bounces"] == b].sum())
        assert np.array_equal(counts, ref_counts), f"pass {p}"
        assert np.allclose(energy, ref_energy, rtol=RTOL, atol=1e-30)
        assert nr == sum(len(r.rays) for r in ref_bundles)
    _compare_surface_fluences(osc, psc, (0, 1, 2, 3))


def _compare_surface_fluences(osc, psc, nodes):
    """per surface: the largest per-bundle peak fluence and the critical-fluence records (peak, bounce level) of the LIDT check"""
    n_crit = 0
    for node in nodes:
        for surf in (0, 1):
            ref_peak = osc.nodes[node].s[surf].peak_seen
            got_peak = psc.node_peak_fluence(node, surf)
            if np.isfinite(ref_peak):
                assert got_peak == pytest.approx(ref_peak, rel=RTOL), (node, surf)
            else:
                assert not np.isfinite(got_peak)
            ref_crit = sorted((pk, b) for pk, b in osc.nodes[node].s[surf].critical.values())
            got_crit = sorted((pk, b) for _, pk, b in psc.node_critical_fluences(node, surf))
            assert len(ref_crit) == len(got_crit), (node, surf)
            for (a, ab), (b, bb) in zip(ref_crit, got_crit):
                assert ab == bb and b == pytest.approx(a, rel=RTOL)
            n_crit += len(ref_crit)
    return n_crit


def test_ghost_focus_fluence_check_sees_the_bundle_before_the_aperture(ctx):
    """The LIDT check of a surface reads the bundle's bounce level BETWEEN refract_on_surface and apodize
    (analyzers/raytrace.rs:118-133).  Here the output aperture of the second lens clips every ray that reaches it: after the
    surface pass those bundles hold no valid ray (bounce level 0), yet the reference has already evaluated their hits with
    the level they arrived with -- ghost bundles of level 2 included.  A low LIDT turns every evaluation into a record."""
    nodes = scenes.c3_telescope()
    nodes[2] = dict(nodes[2], ap_out=("circle", 1.0e-3, 50.0e-3, 0.0), lidt=1.0e-6)
    nodes[1] = dict(nodes[1], lidt=1.0e-6)
    src = scenes.c3_source(2000)
    osc = scenes.to_oracle(nodes)
    ref_passes = osc.ghostfocus(scenes.source_to_oracle(src), 2)
    psc = scenes.to_product(ctx, nodes)
    st = psc.ghostfocus(scenes.source_to_product(src).generate(ctx), 2)
    assert st.interactions == osc.interactions
    for p, ref_bundles in enumerate(ref_passes):
        got = psc.gf_pass_bundles(p)
        assert len(got) == len(ref_bundles)
        for g, r in zip(got, ref_bundles):
            assert_rays_match(g.to_array(), r.rays, keyed=True, what=f"pass {p} bundle {r.uuid}", e_floor=1e-25)
    assert _compare_surface_fluences(osc, psc, (0, 1, 2, 3)) > 4
    levels = {b for _, _, b in psc.node_critical_fluences(2, 1)}
    assert 2 in levels, levels  # a level-2 ghost bundle evaluated on the clipping surface


def test_ghost_focus_with_the_voronoi_estimator(ctx):
    """GhostFocusConfig::fluence_estimator = Voronoi, the reference's default (analyzers/ghostfocus.rs:61): the per-bundle LIDT check
    takes the peak of a 101 x 101 Voronoi estimate of the bundle's hits (rays_hit_map.rs:777-783)"""
    nodes = scenes.c3_telescope()
    nodes[1] = dict(nodes[1], lidt=1.0e-6)
    src = scenes.c3_source(400)
    osc = scenes.to_oracle(nodes)
    osc.ghostfocus(scenes.source_to_oracle(src), 1, estimator="voronoi")
    psc = scenes.to_product(ctx, nodes)
    st = psc.ghostfocus(scenes.source_to_product(src).generate(ctx), 1, estimator="voronoi")
    assert st.interactions == osc.interactions
    n_crit = 0
    for node in (0, 1, 2, 3):
        for surf in (0, 1):
            ref_peak, got_peak = osc.nodes[node].s[surf].peak_seen, psc.node_peak_fluence(node, surf)
            if np.isfinite(ref_peak):
                assert got_peak == pytest.approx(ref_peak, rel=1e-6), (node, surf)  # cell areas: differences of products of positions
            else:
                assert not np.isfinite(got_peak)
            ref_crit = sorted((pk, b) for pk, b in osc.nodes[node].s[surf].critical.values())
            got_crit = sorted((pk, b) for _, pk, b in psc.node_critical_fluences(node, surf))
            assert [b for _, b in ref_crit] == [b for _, b in got_crit]
            n_crit += len(ref_crit)
    assert n_crit >= 2
    osc.ghostfocus(scenes.source_to_oracle(src), 1)  # back to the Binning estimator for the tests that follow


def test_sharded_trace_equals_whole_source(ctx):
    """SURVEY 8e on one GPU: two shards of the C2 source traced separately (as two ranks would), detector results combined
    with the shard API (common map range + accumulate binning = all-reduce of maps; two-phase spot partials) must equal
    the whole-source run: counts bit-exact, map bins and statistics to 1e-12 (only the summation order differs)."""
    import ctypes as C

    from opossum_b200 import ffi as F
    from opossum_b200 import sharding
    nodes, n = scenes.c2_laser_chain(), 30000
    sd = scenes.source_to_product(scenes.c2_source(n))
    whole = scenes.to_product(ctx, nodes)
    st_whole = whole.raytrace(sd.generate(ctx))
    ref_fd = whole.node_fluence(9)
    ref_spot = whole.node_spot(8)
    shards, stats = [], []
    for g in range(2):
        b, e = sharding.shard_range(n, g, 2)
        sc = scenes.to_product(ctx, nodes)
        stats.append(sc.raytrace(sd.generate(ctx, b, e)))
        shards.append(sc)
    assert sum(s.interactions for s in stats) == st_whole.interactions
    assert sum(s.valid_out for s in stats) == st_whole.valid_out and sum(s.alive_out for s in stats) == st_whole.alive_out
    # fluence: global range = min/max of the shard ranges, both shards binned into the same device map
    boxes = [ob.hitmaps_bounding_box(sc.node_hitmaps(9, 0))[0] for sc in shards]
    lrtb = (min(b[0] for b in boxes), max(b[1] for b in boxes), max(b[2] for b in boxes), min(b[3] for b in boxes))
    assert lrtb == ref_fd.lrtb
    dev_map = ctx.device_alloc(100 * 83 * 8)
    for g, sc in enumerate(shards):
        fd = ob.fluence_binning(ctx, sc.node_hitmaps(9, 0), 100, 83, lrtb=lrtb, map_dev=dev_map, accumulate=g > 0, download=g == 1)
    ctx.device_free(dev_map)
    # the same hit points binned in two pieces: the fixed-point splat rounds every tap to 2^-33 of itself (unit chosen per CTA), so
    # a bin is reproducible to 2.4e-10 of its own value, not to the last bits
    assert_map_match(fd.map, ref_fd.map, rtol=2.4e-10, same_hits=True)
    assert fd.peak == pytest.approx(ref_fd.peak, rel=1e-12) and fd.total_energy == pytest.approx(ref_fd.total_energy, rel=1e-12)
    # spot diagram: phase-1 partials summed, phase 2 about the global centroids, radius accumulators combined
    parts = [sc.node_spot_partial(8) for sc in shards]
    glob = F.OpmSpotPartial()
    for j in range(7):
        glob.sum[j] = parts[0].sum[j] + parts[1].sum[j]
    glob.n_valid, glob.n_total = parts[0].n_valid + parts[1].n_valid, parts[0].n_total + parts[1].n_total
    glob.first_key = parts[0].first_key if parts[0].n_valid else parts[1].first_key
    radii = []
    for sc in shards:
        p = F.OpmSpotPartial.from_buffer_copy(glob)
        sc.node_spot_partial(8, reduced=p)
        radii.append(p)
    glob.max_d2_mm2 = max(p.max_d2_mm2 for p in radii)
    glob.sum_d2_mm2 = sum(p.sum_d2_mm2 for p in radii)
    glob.sum_ed2 = sum(p.sum_ed2 for p in radii)
    got = F.OpmSpotStats()
    F.lib().opm_spot_stats_from_partial(C.byref(glob), C.byref(got))
    assert got.n_valid == ref_spot.n_valid and got.n_total == ref_spot.n_total and got.bounce_lvl == ref_spot.bounce_lvl
    for f in ("total_energy", "beam_radius_geo", "beam_radius_rms", "energy_beam_radius_rms"):
        assert getattr(got, f) == pytest.approx(getattr(ref_spot, f), rel=1e-12), f
    assert np.allclose(got.energy_centroid[:], ref_spot.energy_centroid[:], rtol=1e-9, atol=1e-15)


def test_device_chained_readout_equals_stepwise_reports(ctx):
    """sharding.DetectorReadout (device-chained: bbox -> Binning -> peak/total and spot sums -> radii without host round
    trips) must give the same reports as the stepwise calls, which are checked against the oracle above."""
    import torch

    from opossum_b200 import sharding
    psc = scenes.to_product(ctx, scenes.c2_laser_chain())
    rays = scenes.source_to_product(scenes.c2_source(25000)).generate(ctx)
    psc.raytrace(rays)
    ref = {n: psc.node_fluence(n, 100, 83) for n in (9, 10)}
    ref_spot = psc.node_spot(8)
    ro = sharding.DetectorReadout(ctx, psc, (9, 10), (8,), 100, 83, torch.device("cuda", 0))
    for _ in range(2):  # buffers are reused across calls
        torch.cuda.synchronize()
        got = ro.run(download_maps=True)
        ctx.synchronize()
        for n in (9, 10):
            assert got[n]["lrtb"] == ref[n].lrtb
            assert np.allclose(got[n]["map"], ref[n].map, rtol=1e-12, atol=1e-12 * ref[n].peak)
            assert got[n]["peak"] == pytest.approx(ref[n].peak, rel=1e-12)
            assert got[n]["total_energy"] == pytest.approx(ref[n].total_energy, rel=1e-12)
        s = got[8]
        assert s.n_valid == ref_spot.n_valid and s.n_total == ref_spot.n_total and s.bounce_lvl == ref_spot.bounce_lvl
        for f in ("total_energy", "beam_radius_geo", "beam_radius_rms", "energy_beam_radius_rms"):
            assert getattr(s, f) == pytest.approx(getattr(ref_spot, f), rel=1e-12), f
        assert np.allclose(s.energy_centroid[:], ref_spot.energy_centroid[:], rtol=1e-12, atol=1e-18)


@pytest.mark.parametrize("world", [1, 2, 8])
def test_readout_combine_kernels_match_their_specification(ctx, world):
    """opm_readout_combine_ranges_sums / _maps_radii (what every rank runs on an all-gathered block) against the torch
    restatement that the world-size-2 gloo test exercises on CPU; rank order makes the sums reproducible bit for bit"""
    import ctypes as C

    import torch

    from opossum_b200 import ffi as F
    from opossum_b200 import sharding
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(world)
    nf, ns, maps_len = 2, 3, 2 * 100 * 83
    nA = 4 * nf + 11 * ns
    gA = torch.rand((world, nA), generator=g, dtype=torch.float64) * 10 - 3
    for k in range(ns):  # ids (o + 10) of the first valid ray: some ranks have none (-1), some tie
        o = 4 * nf + 11 * k
        gA[:, o + 10] = torch.tensor([(-1.0, 5.0, 5.0, 2.0, -1.0, 7.0, 2.0, 9.0)[(r + k) % 8] for r in range(world)])
        gA[:, o + 9] = torch.arange(world, dtype=torch.float64) + 10 * k
    gA[0, 1] = -math.inf  # a rank without hits contributes -inf to the ranges
    gB = torch.rand((world, maps_len + 3 * ns), generator=g, dtype=torch.float64)
    L, vp = F.lib(), C.c_void_p
    dA, dB = gA.to(dev), gB.to(dev)
    outA, outB = torch.zeros(nA, dtype=torch.float64, device=dev), torch.zeros(maps_len + 3 * ns, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    stream.synchronize()
    assert L.opm_readout_combine_ranges_sums(ctx.h, vp(dA.data_ptr()), world, nf, ns, vp(outA.data_ptr())) == 0
    assert L.opm_readout_combine_maps_radii(ctx.h, vp(dB.data_ptr()), world, maps_len, ns, vp(outB.data_ptr())) == 0
    ctx.synchronize()
    refA = sharding.combine_ranges_sums_reference(gA, nf, ns)
    refB = sharding.combine_maps_radii_reference(gB, maps_len, ns)
    assert torch.allclose(outA.cpu(), refA, rtol=1e-15, atol=0)
    assert torch.equal(outA.cpu()[: 4 * nf], refA[: 4 * nf])
    for k in range(ns):
        o = 4 * nf + 11 * k
        assert torch.equal(outA.cpu()[o + 9: o + 11], refA[o + 9: o + 11])  # the same rank's pair, ties to the lowest rank
    assert torch.allclose(outB.cpu(), refB, rtol=1e-14, atol=0)
    assert torch.equal(outB.cpu()[maps_len::3], refB[maps_len::3])


@pytest.fixture(scope="module")
def jit_ctx():
    c = ob.Context(0)
    c.set_jit(2, 1)  # every launch of the contracted build waits for and runs the kernel compiled for its program structure
    yield c
    c.close()


@pytest.mark.parametrize("case", ["c1", "c2", "c4", "c5"])
def test_specialised_kernel_matches_oracle(jit_ctx, case):
    """the run-time compiled straight-line kernel (opm_jit.cu) against the oracle on the BASELINE scenes"""
    nodes, src, dets, spot = {
        "c1": (scenes.c1_single_lens(), scenes.c1_source(60), [3], 2),
        "c2": (scenes.c2_laser_chain(), scenes.c2_source(20000), [9, 10], 8),
        "c4": (scenes.c4_fluence(), scenes.c4_source(150), [2], None),
        "c5": (scenes.c5_broadband(), scenes.c5_source(1500, 8), [4], None)}[case]
    before = jit_ctx.jit_stats()
    osc, ref_out, psc, rays, st = run_both(jit_ctx, nodes, src)
    after = jit_ctx.jit_stats()
    assert after["launches"] > before["launches"] and after["compile_failures"] == 0, after
    assert st.interactions == osc.interactions
    assert_rays_match(rays.to_array(), ref_out, what=f"{case} specialised kernel")
    for node in dets:
        check_detector(osc, psc, node)
    if spot is not None:
        ref_spot, s = osc.node_spot(spot), psc.node_spot(spot)
        assert s.n_valid == ref_spot["n_valid"] and s.total_energy == pytest.approx(ref_spot["total_energy"], rel=RTOL)
        assert s.beam_radius_geo == pytest.approx(ref_spot["geo"], rel=RTOL)


def test_specialised_kernel_ghost_focus(jit_ctx):
    """ghost focus (child emission with warp-aggregated appends, mirrors, limits) through the specialised kernels"""
    nodes, src = scenes.c3_telescope(), scenes.c3_source(2000)
    osc = scenes.to_oracle(nodes)
    ref_passes = osc.ghostfocus(scenes.source_to_oracle(src), 2)
    psc = scenes.to_product(jit_ctx, nodes)
    st = psc.ghostfocus(scenes.source_to_product(src).generate(jit_ctx), 2)
    assert st.interactions == osc.interactions and jit_ctx.jit_stats()["compile_failures"] == 0
    for p, ref_bundles in enumerate(ref_passes):
        got = psc.gf_pass_bundles(p)
        assert len(got) == len(ref_bundles)
        for g, r in zip(got, ref_bundles):
            assert_rays_match(g.to_array(), r.rays, keyed=True, what=f"pass {p} bundle {r.uuid}", e_floor=1e-25)


def test_fluence_detector_with_kde_estimator(ctx):
    """FluenceDetector with FluenceEstimator::KDE end to end: C1 scene at 1500 rays, detector map 40x33"""
    nodes, src = scenes.c1_single_lens(), scenes.c1_source(39)
    osc = scenes.to_oracle(nodes)
    osc.raytrace(scenes.source_to_oracle(src))
    ref_map, ref_lrtb, ref_bw = orc.fluence_kde(osc.nodes[3].s[0].merged_hits(), 40, 33)
    psc = scenes.to_product(ctx, nodes)
    psc.raytrace(scenes.source_to_product(src).generate(ctx))
    fd = psc.node_fluence(3, 40, 33, estimator="kde")
    assert fd.band_width == pytest.approx(ref_bw, rel=1e-9)
    assert np.allclose(fd.lrtb, ref_lrtb, rtol=1e-9, atol=1e-15)
    assert_map_match(fd.map, ref_map, rtol=1e-8)  # the map range itself carries the 1e-9 of the hit positions
    with pytest.raises(ValueError):
        psc.node_fluence(3, 40, 33, estimator="nearest")


def test_fluence_detector_with_voronoi_estimator(ctx):
    """FluenceDetector with its DEFAULT estimator (FluenceEstimator::Voronoi, nodes/fluence_detector/mod.rs:61) end to end: lens +
    detector, Fibonacci source; HitMap::calc_combined_fluence_with_voronoi of the detector's hits against the oracle.  The hit
    positions agree to 1e-9 relative, the cell areas -- differences of products of positions -- to a little less; map points ON the
    convex hull's boundary may fall to either side.  (100, 83) -- what the reference's own report asks for -- is refused: the
    reference panics on it (nx x nx estimates added to an nx x ny matrix)."""
    nodes = scenes.c1_single_lens()
    src = scenes.c3_source(1500)
    osc = scenes.to_oracle(nodes)
    osc.raytrace(scenes.source_to_oracle(src))
    hits = osc.nodes[3].s[0].merged_hits()
    ref_map, ref_lrtb = orc.fluence_voronoi_combined([hits[hits["bounce"] == b] for b in np.unique(hits["bounce"])], 48, 48)
    psc = scenes.to_product(ctx, nodes)
    psc.raytrace(scenes.source_to_product(src).generate(ctx))
    fd = psc.node_fluence(3, 48, 48, estimator="voronoi")
    assert np.allclose(fd.lrtb, ref_lrtb, rtol=1e-9, atol=1e-15)
    both = np.isfinite(fd.map) & np.isfinite(ref_map)
    assert (np.isfinite(fd.map) != np.isfinite(ref_map)).sum() <= 8 and both.sum() > 0.6 * ref_map.size
    assert np.max(np.abs(fd.map[both] - ref_map[both]) / ref_map[both]) < 1e-6
    assert fd.peak == pytest.approx(np.nanmax(ref_map), rel=1e-6)
    with pytest.raises(ob.OpossumError):
        psc.node_fluence(3, 100, 83, estimator="voronoi")


@pytest.mark.parametrize("jit", [0, 2], ids=["interpreter", "specialised"])
def test_reflective_grating_node(ctx, jit):
    """ReflectiveGrating (nodes/reflective_grating.rs): the reference's Littrow KATs through the scene layer, and a
    stretcher-like chain (lens, grating near Littrow, detector) against the oracle"""
    import math

    from opossum_b200 import scene as S
    MM, NM = scenes.MM, scenes.NM
    ctx.set_jit(jit, 1)
    try:
        for extra in (0.0, math.radians(1.0)):  # reflective_grating.rs:386-443
            tilt = S.littrow_angle(1740e3, -1, 1000 * NM) + extra
            nodes = [scenes.node("source"), scenes.node("reflective_grating", z=0.0, euler=(0.0, tilt, 0.0), p=(1740e3, -1))]
            src = {"pos": ("grid", 1, 1, 0.0, 0.0), "energy": ("uniform", 1.0), "wvl": [1000 * NM], "w": [1.0]}
            psc = scenes.to_product(ctx, nodes)
            rays = scenes.source_to_product(src).generate(ctx)
            st = psc.raytrace(rays)
            got = rays.to_array()
            assert st.valid_out == 1 and np.array_equal(got["pos"][0], [0.0, 0.0, 0.0])
            if extra == 0.0:
                expect = np.array([0.0, 0.0, -1.0])
            else:
                ia = math.asin(-1000e-9 * 1740000. / 2.) + extra
                da = math.asin(-1740000. * 1000e-9 - math.sin(ia))
                expect = np.array([math.sin(-ia + da), 0.0, -math.cos(-ia + da)])
            assert np.all(np.abs(got["dir"][0] - expect) <= 4e-15)
        tilt = S.littrow_angle(1740e3, -1, 1030 * NM) + math.radians(8.0)
        nodes = [scenes.node("source"),
                 scenes.node("lens", z=50 * MM, p=(800 * MM, -800 * MM, 4 * MM), index=scenes.HK9L),
                 scenes.node("reflective_grating", z=250 * MM, euler=(0.0, tilt, 0.0), p=(1740e3, -1), ap_in=("rect", 30 * MM, 20 * MM, 0.0, 0.0)),
                 scenes.node("fluence_detector", z=100 * MM, xy=(-40 * MM, 0.0), euler=(0.0, math.pi + 0.3, 0.0))]
        src = {"pos": ("fibonacci_ellipse", 3000, 6 * MM, 6 * MM), "energy": ("uniform", 1.0),
               "spectrum": ("gaussian", 1010 * NM, 1050 * NM, 9, 1030 * NM, 10 * NM, 1.0)}
        osc, ref_out, psc, rays, st = run_both(ctx, nodes, src)
        assert st.interactions == osc.interactions and orc.nr_of_rays(ref_out) > 20000
        assert_rays_match(rays.to_array(), ref_out, what="lens + grating + detector")
        wv = np.unique(ref_out["wvl"])
        x_by_w = [ref_out["pos"][(ref_out["wvl"] == w) & (ref_out["valid"] == 1)][:, 0].mean() for w in wv]
        assert np.all(np.diff(x_by_w) != 0.0) and np.ptp(x_by_w) > 1e-4  # angular dispersion reaches the detector
        check_detector(osc, psc, 3, 64, 48)
    finally:
        ctx.set_jit(1, 200000)
