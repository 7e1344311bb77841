"""GPU parity tests of the bundle operations: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Bar: counts / valid / bounce / refraction state bit-exact, per-ray pos/dir/E/OPL within 1e-9
relative (helpers.RTOL).  Also transcribes the reference's exact-equality KATs onto the GPU (strict kernel)."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import (HK9L, HZF52, RTOL, aperture_pair, append_pair, assert_hits_match, assert_rays_match, index_pair,
                     iso_pair, make_bundle, mm, nm, node_sphere_pair, surface_pair)

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps
Z = (0.0, 0.0, 1.0)


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def gpu_refract(ctx, rays_np, surf, idx, refraction_intended=True, strategy=F.MISS_STOP, strict=False):
    """one REFRACT op through the fused kernel; returns (rays, children, hits, stats) as numpy"""
    r = ob.Rays.from_array(ctx, rays_np)
    hm = ob.HitMap(ctx, max(len(rays_np), 1))
    ch = ob.Rays(ctx, max(len(rays_np), 1))
    st = ob.Program().refract(surf, idx, refraction_intended, strategy, hit_map=hm, children=ch).run(r, strict=strict)
    return r.to_array(), ch.to_array(), hm.download(), st


# ------------------------------------------------------------------ reference KATs on the GPU
def test_kat_refract_collimated(ctx):
    """ray.rs:1203-1290 on the GPU (strict kernel: exact equality like the reference's assert_eq!)"""
    _, s = surface_pair(F.GEOM_PLANE, iso=iso_pair(mm(0.0, 0.0, 10.0)), coating=F.COAT_CONSTANT_R, r=0.2)
    _, n15 = index_pair("const", 1.5)
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1054.0), 1.0)
    out, ch, hits, st = gpu_refract(ctx, ray, s, n15, strict=True)
    r, c = out[0], ch[0]
    assert tuple(r["pos"]) == mm(0.0, 0.0, 10.0) and r["n"] == 1.5 and tuple(r["dir"]) == Z
    assert r["opl"] == mm(10.0) and r["bounces"] == 0 and r["refr"] == 1 and r["e"] == (1.0 - 0.2) * 1.0
    assert tuple(c["pos"]) == mm(0.0, 0.0, 10.0) and c["n"] == 1.0 and tuple(c["dir"]) == (0.0, 0.0, -1.0)
    assert c["opl"] == mm(10.0) and c["bounces"] == 1 and c["refr"] == 0 and c["e"] == 0.2 * 1.0
    assert (hits[0][0], hits[1][0], hits[2][0], hits[3][0], hits[4][0]) == (0.0, 0.0, 0.0, 1.0, 0)
    assert (st.interactions, st.children, st.hits, st.missed) == (1, 1, 1, 0)
    for bad in (0.9, math.nan, math.inf):
        with pytest.raises(ob.OpossumError) as ei:
            gpu_refract(ctx, ray, s, index_pair("const", bad)[1])
        assert ei.value.code in (F.OPM_ERR_N2_INVALID, F.OPM_ERR_INDEX_MODEL)


def test_kat_refract_45deg_and_tir(ctx):
    """ray.rs:1347-1458: dir (0, 0.4714045207910317, 0.8819171036881969); total internal reflection"""
    _, s = surface_pair(F.GEOM_PLANE, iso=iso_pair(mm(0.0, 0.0, 10.0)))
    ray = orc.ray_new((0.0, 0.0, 0.0), (1.0, 0.0, 1.0), nm(1054.0), 1.0)
    out, _, _, _ = gpu_refract(ctx, ray, s, index_pair("const", 1.5)[1], strict=True)
    assert tuple(out[0]["pos"]) == mm(10.0, 0.0, 10.0)
    assert out[0]["dir"][0] == 0.4714045207910317 and abs(out[0]["dir"][1]) <= EPS
    assert abs(out[0]["dir"][2] - 0.8819171036881969) <= EPS
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 2.0, 1.0), nm(1054.0), 1.0)
    ray["n"] = 1.5
    out, ch, hits, st = gpu_refract(ctx, ray, s, index_pair("const", 1.0)[1], strict=True)
    assert len(ch) == 0 and len(hits[0]) == 0 and st.missed == 1
    assert tuple(out[0]["pos"]) == mm(0.0, 20.0, 10.0)
    t = np.array([0.0, 2.0, -1.0]) / math.sqrt(5.0)
    assert np.all(np.abs(out[0]["dir"] - t) <= EPS) and out[0]["bounces"] == 1 and out[0]["e"] == 1.0


def test_kat_miss_strategies(ctx):
    """ray.rs:1319-1346, rays.rs:2246-2266"""
    _, s = surface_pair(F.GEOM_PLANE, iso=iso_pair(mm(0.0, 0.0, 10.0)))
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), nm(1054.0), 1.0)
    for strat, valid in ((F.MISS_STOP, 0), (F.MISS_IGNORE, 1)):
        out, ch, _, st = gpu_refract(ctx, ray, s, index_pair("const", 1.5)[1], strategy=strat)
        assert len(ch) == 0 and st.missed == 1 and out[0]["valid"] == valid
        assert tuple(out[0]["pos"]) == (0.0, 0.0, 0.0) and out[0]["opl"] == 0.0 and out[0]["n"] == 1.0


def test_kat_sphere_and_fresnel(ctx):
    """surface/sphere.rs:213-402 (exact on-axis hits, tangent ray) and coatings/fresnel.rs:54-83 via the energy split"""
    for radius in (mm(10.0), mm(-10.0)):
        s = ob.OpticSurface.sphere_at_position(radius, (0.0, 0.0, 0.0))
        out, _, _, _ = gpu_refract(ctx, orc.ray_new(mm(0.0, 0.0, -5.0), Z, nm(1053.0), 1.0), s, None, strict=True)
        assert tuple(out[0]["pos"]) == (0.0, 0.0, 0.0)
        out, _, _, st = gpu_refract(ctx, orc.ray_new(mm(0.0, 0.0, 5.0), Z, nm(1053.0), 1.0), s, None, strict=True)
        assert st.missed == 1 and out[0]["valid"] == 0
    s = ob.OpticSurface.sphere_at_position(mm(1.0), mm(0.0, 0.0, 10.0))
    out, _, _, st = gpu_refract(ctx, orc.ray_new(mm(0.0, 1.0, 0.0), Z, nm(1053.0), 1.0), s, None, strict=True)
    assert tuple(out[0]["pos"]) == mm(0.0, 1.0, 11.0) and st.missed == 0  # disc == 0 exactly -> Roots::One
    out, _, _, st = gpu_refract(ctx, orc.ray_new(mm(0.0, 1.1, 0.0), Z, nm(1053.0), 1.0), s, None, strict=True)
    assert st.missed == 1
    fs = ob.OpticSurface.plane(ob.Isometry.new_along_z(mm(1.0)), coating=F.COAT_FRESNEL)
    out, ch, _, _ = gpu_refract(ctx, orc.ray_new((0, 0, 0), Z, nm(1000.0), 1.0), fs, index_pair("const", 1.5)[1], strict=True)
    assert abs(ch[0]["e"] - 0.04) <= 4 * EPS
    out, ch, _, _ = gpu_refract(ctx, orc.ray_new((0, 0, 0), (0.0, 1.0, 1.0), nm(1000.0), 1.0), fs, index_pair("const", 1.5)[1], strict=True)
    assert abs(ch[0]["e"] - 0.05023991101223595) <= 8 * EPS  # CUDA libm vs glibc: a few ulp in acos/sin/cos


# ------------------------------------------------------------------ seeded bundles vs the oracle
GEOMS = {
    "plane": lambda node: surface_pair(F.GEOM_PLANE, iso=node),
    "sphere_convex": lambda node: node_sphere_pair(mm(120.0), node),
    "sphere_concave": lambda node: node_sphere_pair(mm(-90.0), node),
    "cylinder": lambda node: surface_pair(F.GEOM_CYLINDER, mm(150.0), append_pair(node, iso_pair((0.0, 0.0, mm(150.0))))),
    "cylinder_neg": lambda node: surface_pair(F.GEOM_CYLINDER, mm(-150.0), append_pair(node, iso_pair((0.0, 0.0, mm(-150.0))))),
    "parabola": lambda node: surface_pair(F.GEOM_PARABOLA, mm(-200.0), node),
    "parabola_pos": lambda node: surface_pair(F.GEOM_PARABOLA, mm(200.0), node),
}
NODE_ISOS = {
    "axis": ((0.0, 0.0, mm(50.0)), (0.0, 0.0, 0.0)),
    "tilted": ((mm(0.7), mm(-0.4), mm(60.0)), (0.11, -0.07, 0.3)),
}
INDICES = {
    "none": None,
    "const": ("const", 1.5068),
    "sellmeier": ("sellmeier1", *HK9L, nm(500.0), nm(2000.0)),
    "schott": ("schott", *HZF52, nm(500.0), nm(2000.0)),
    "conrady": ("conrady", 1.45, 0.012, 0.0004, nm(500.0), nm(2000.0)),
}


def _with_coating(pair, coating, r=0.0):
    o, g = pair
    o.coating, o.coating_r = coating, r
    g.c.coating, g.c.coating_r = coating, r
    return o, g


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
@pytest.mark.parametrize("iso_name", list(NODE_ISOS))
@pytest.mark.parametrize("geom", list(GEOMS))
def test_refract_geometries(ctx, geom, iso_name, strict):
    """every surface solver x placement; Sellmeier glass, Fresnel coating (a4, a5, a7, a8 of SURVEY 8a)"""
    so, sg = _with_coating(GEOMS[geom](iso_pair(*NODE_ISOS[iso_name])), F.COAT_FRESNEL)
    io, ig = index_pair(*INDICES["sellmeier"])
    rays = make_bundle(4096)
    ref = rays.copy()
    ch_ref, hits_ref, st_ref = orc.rays_refract_on_surface(ref, so, io, True, orc.MISS_STOP)
    out, ch, hits, st = gpu_refract(ctx, rays, sg, ig, strict=strict)
    assert (st.interactions, st.children, st.hits) == (st_ref.n_interactions, st_ref.n_children, st_ref.n_hits)
    assert bool(st.missed) == bool(st_ref.rays_missed)
    assert st_ref.n_children > 0
    assert_rays_match(out, ref, what=f"{geom} refracted")
    assert_rays_match(ch, ch_ref, keyed=True, what=f"{geom} reflected")
    assert_hits_match(hits, hits_ref)  # hits are slot-indexed: already in ray order
    assert st.valid_out == orc.nr_of_rays(ref) and st.alive_out == len(ref)


@pytest.mark.parametrize("index", list(INDICES))
@pytest.mark.parametrize("coating", [(F.COAT_IDEAL_AR, 0.0), (F.COAT_CONSTANT_R, 0.05), (F.COAT_FRESNEL, 0.0)])
def test_refract_index_and_coating(ctx, index, coating):
    """all four index models + passive surface x all three coatings, leaving glass (n1 > n2 -> some TIR)"""
    so, sg = _with_coating(node_sphere_pair(mm(-40.0), iso_pair((0.0, 0.0, mm(30.0)), (0.02, 0.01, 0.0))), *coating)
    rays = make_bundle(4096, divergence=0.35)
    rays["n"] = 1.7  # rays travel inside glass
    io = ig = None
    if INDICES[index] is not None:
        io, ig = index_pair(*INDICES[index])
    ref = rays.copy()
    ch_ref, hits_ref, st_ref = orc.rays_refract_on_surface(ref, so, io, True, orc.MISS_IGNORE)
    out, ch, hits, st = gpu_refract(ctx, rays, sg, ig, strategy=F.MISS_IGNORE)
    assert (st.interactions, st.children, st.hits) == (st_ref.n_interactions, st_ref.n_children, st_ref.n_hits)
    assert_rays_match(out, ref, what="refracted")
    # Fresnel at n1 == n2 (passive surface) is exactly 0 analytically; off axis both sides return ~1e-33 of rounding
    # noise (coatings/fresnel.rs:21-33 cancels), so child energies are compared on the scale of the parent energy
    assert_rays_match(ch, ch_ref, keyed=True, what="reflected", e_floor=1e-15 * float(rays["e"].max()))
    assert_hits_match(hits, hits_ref)


def test_total_internal_reflection_bundle(ctx):
    """ray.rs:674-679: steep glass->vacuum incidence; part of the bundle is totally reflected (bounce +1, no child,
    no hit-map entry), the rest refracts"""
    so, sg = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(30.0)), (math.radians(33.0), 0.0, 0.0)))
    rays = make_bundle(4096, divergence=0.25)
    rays["n"] = 1.7
    io, ig = index_pair("const", 1.0)
    ref = rays.copy()
    ch_ref, hits_ref, st_ref = orc.rays_refract_on_surface(ref, so, io, True, orc.MISS_STOP)
    out, ch, hits, st = gpu_refract(ctx, rays, sg, ig)
    n_tir = int((ref["bounces"] == 1).sum())
    assert 0 < n_tir < len(rays), n_tir
    assert st.missed == n_tir and st.children == len(ch_ref) == len(rays) - n_tir
    assert_rays_match(out, ref, what="TIR bundle")
    assert_rays_match(ch, ch_ref, keyed=True, what="TIR children")
    assert_hits_match(hits, hits_ref)


def test_refract_errors_match_reference(ctx):
    """wavelength outside the model range aborts (refr_index_sellmeier1.rs:79-81); ConstantR 0.0 is an error at trace
    time (coatings/mod.rs:49, constant_r.rs:23)"""
    so, sg = surface_pair(F.GEOM_PLANE, iso=iso_pair((0, 0, mm(10.0))))
    io, ig = index_pair("sellmeier1", *HK9L, nm(500.0), nm(2000.0))
    rays = make_bundle(64, wvl=(nm(400.0), nm(2100.0)))
    with pytest.raises(orc.OracleError) as eo:
        orc.rays_refract_on_surface(rays.copy(), so, io)
    with pytest.raises(ob.OpossumError) as eg:
        gpu_refract(ctx, rays, sg, ig)
    assert eg.value.code == F.OPM_ERR_WVL_RANGE and eo.value.code == 2
    assert str(eg.value) == str(eo.value)
    so, sg = surface_pair(F.GEOM_PLANE, iso=iso_pair((0, 0, mm(10.0))), coating=F.COAT_CONSTANT_R, r=0.0)
    with pytest.raises(ob.OpossumError) as eg:
        gpu_refract(ctx, make_bundle(64), sg, index_pair("const", 1.5)[1])
    assert eg.value.code == F.OPM_ERR_COATING


def test_async_launch_errors_surface_at_synchronize(ctx):
    """a launch that does not read its statistics back (sync=False: throughput runs, OPM_SCENE_ASYNC) cannot return the error on
    which the reference aborts (rays.rs:987-991); the status bits stay on the device across later launches and come back from
    Context.synchronize -- once"""
    _, sg = surface_pair(F.GEOM_PLANE, iso=iso_pair((0, 0, mm(10.0))))
    _, ig = index_pair("sellmeier1", *HK9L, nm(500.0), nm(2000.0))
    bad = ob.Rays.from_array(ctx, make_bundle(64, wvl=(nm(400.0), nm(2100.0))))
    good = ob.Rays.from_array(ctx, make_bundle(64, wvl=(nm(600.0), nm(1500.0))))
    ctx.synchronize()
    assert ob.Program().refract(sg, ig).run(bad, sync=False) is None
    ob.Program().refract(sg, ig).run(good, sync=False)  # a clean launch afterwards clears nothing
    with pytest.raises(ob.OpossumError) as e:
        ctx.synchronize()
    assert e.value.code == F.OPM_ERR_WVL_RANGE
    ctx.synchronize()  # reported once
    ob.Program().refract(sg, ig).run(good, sync=False)
    ctx.check_errors()


def test_mirror_drops_and_keeps_order(ctx):
    """refraction_intended = false (nodes/thin_mirror.rs:209-250, SURVEY C-1): the reflected bundle continues,
    invalid and missing rays leave the bundle; compact and in-place outputs agree"""
    node = iso_pair((0.0, 0.0, mm(40.0)), (math.pi / 4, 0.0, 0.0))
    so, sg = _with_coating(surface_pair(F.GEOM_PLANE, iso=node), F.COAT_CONSTANT_R, 1.0)
    rays = make_bundle(4096, divergence=0.6)
    rays["valid"][::7] = 0
    rays["dir"][::11] *= -1.0  # these fly away from the mirror -> miss
    ref = rays.copy()
    ch_ref, hits_ref, st_ref = orc.rays_refract_on_surface(ref, so, None, False, orc.MISS_STOP)
    assert 0 < len(ch_ref) < len(rays)
    # in place: rays without a child are flagged removed and skipped by download
    r = ob.Rays.from_array(ctx, rays)
    st = ob.Program().mirror(sg).run(r)
    assert st.alive_out == len(ch_ref)
    assert_rays_match(r.to_array(), ch_ref, what="mirror in place")  # slot order == reference order
    assert r.nr_of_rays(False) == len(ch_ref)
    r.compact()
    assert len(r) == len(ch_ref)
    assert_rays_match(r.to_array(), ch_ref, what="mirror + stable compaction")
    # compacted by the trace kernel itself (warp-ballot aggregated append): same set, order keyed by id
    r2 = ob.Rays.from_array(ctx, rays)
    out = ob.Rays(ctx, len(rays))
    st2 = ob.Program().mirror(sg).run(r2, out=out, compact=True)
    assert len(out) == len(ch_ref) == st2.alive_out
    assert_rays_match(out.to_array(), ch_ref, keyed=True, what="mirror compact")


APERTURES = {
    "circle": ("circle", mm(5.0), mm(0.3), mm(-0.2)),
    "circle_obstruction": ("circle", mm(2.0), 0.0, 0.0),
    "rect": ("rect", mm(9.0), mm(6.0), mm(0.5), 0.0),
    "gaussian": ("gaussian", mm(4.0), mm(6.0), 0.0, mm(1.0)),
    "polygon": ("polygon", [[(-0.006, -0.006), (0.006, -0.006), (0.0, 0.007)], [(0.0, 0.007), (0.006, -0.006), (0.008, 0.008)]]),
}


@pytest.mark.parametrize("name", list(APERTURES) + ["stack", "none"])
def test_apodize(ctx, name):
    """rays.rs:557-573 + aperture.rs (circle/rect/polygon/gaussian/stack, hole and obstruction)"""
    if name == "stack":
        ao, ag = aperture_pair("stack", [aperture_pair("circle", mm(6.0)), aperture_pair("rect", mm(8.0), mm(8.0)),
                                         aperture_pair("gaussian", mm(5.0), mm(5.0))])
    elif name == "none":
        ao, ag = aperture_pair("none")
    else:
        kind, *a = APERTURES[name]
        ao, ag = aperture_pair(kind, *a, obstruction=name.endswith("obstruction"))
    io, ig = iso_pair((mm(0.2), 0.0, mm(3.0)), (0.05, -0.02, 0.4))
    rays = make_bundle(4096)
    rays["valid"][::9] = 0
    ref = rays.copy()
    inv_ref = orc.rays_apodize(ref, ao, io)
    r = ob.Rays.from_array(ctx, rays)
    inv = r.apodize(ag, ig)
    assert inv == inv_ref
    assert_rays_match(r.to_array(), ref, what=f"apodize {name}")
    if name in ("circle", "rect", "polygon"):
        assert 0 < orc.nr_of_rays(ref) < int(rays["valid"].sum())


def test_threshold_limits_filter_split(ctx):
    """rays.rs:1145-1162, 1386-1404, 1119-1133, 1253-1264 incl. their error conditions"""
    rays = make_bundle(2048)
    rays["bounces"] = np.arange(len(rays)) % 5
    rays["refr"] = np.arange(len(rays)) % 3
    rays["valid"][::13] = 0
    ref = rays.copy()
    r = ob.Rays.from_array(ctx, rays)
    thr = float(np.median(rays["e"]))
    orc.rays_threshold(ref, thr); r.invalidate_by_threshold_energy(thr)
    assert_rays_match(r.to_array(), ref, what="threshold")
    orc.rays_threshold(ref, -1.0); r.invalidate_by_threshold_energy(-1.0)  # negative: warning, unmodified
    assert_rays_match(r.to_array(), ref, what="negative threshold")
    for bad in (math.nan, math.inf):
        with pytest.raises(ob.OpossumError) as ei:
            r.invalidate_by_threshold_energy(bad)
        assert ei.value.code == F.OPM_ERR_THRESHOLD
    orc.rays_filter_bounces(ref, 3); r.filter_by_nr_of_bounces(3)
    assert_rays_match(r.to_array(), ref, what="bounces")
    orc.rays_filter_refractions(ref, 2); r.filter_by_nr_of_refractions(2)
    assert_rays_match(r.to_array(), ref, what="refractions")
    assert orc.nr_of_rays(ref) > 0
    orc.rays_filter_energy(ref, 0.3); r.filter_energy(0.3)
    assert_rays_match(r.to_array(), ref, what="filter")
    for bad in (-0.1, 1.1):
        with pytest.raises(ob.OpossumError):
            r.filter_energy(bad)
        with pytest.raises(ob.OpossumError) as ei:
            r.split(bad)
        assert ei.value.code == F.OPM_ERR_SPLIT_RATIO
    sp_ref = orc.rays_split(ref, 0.6)
    sp = r.split(0.6)
    assert_rays_match(r.to_array(), ref, what="split kept")
    assert_rays_match(sp.to_array(), sp_ref, what="split off")  # invalid rays are copied too (rays.rs:1260)
    assert len(sp) == len(rays)


def test_paraxial_and_transform(ctx):
    """rays.rs:943-959 / ray.rs:443-471 and Rays::transformed_by_iso"""
    io, ig = iso_pair((mm(0.1), mm(-0.3), mm(20.0)), (0.02, 0.03, -0.1))
    rays = make_bundle(2048, z=mm(20.0))
    rays["valid"][::5] = 0
    ref = rays.copy()
    r = ob.Rays.from_array(ctx, rays)
    orc.rays_refract_paraxial(ref, mm(150.0), io); r.refract_paraxial(mm(150.0), ig)
    assert_rays_match(r.to_array(), ref, what="paraxial")
    orc.rays_refract_paraxial(ref, mm(-80.0), io); r.refract_paraxial(mm(-80.0), ig)
    assert_rays_match(r.to_array(), ref, what="paraxial negative")
    for bad in (0.0, math.inf, math.nan):
        with pytest.raises(ob.OpossumError) as ei:
            r.refract_paraxial(bad, ig)
        assert ei.value.code == F.OPM_ERR_FOCAL_LENGTH
    for inverse in (False, True):
        orc.rays_transform(ref, io, inverse); r.transformed_by_iso(ig, inverse)
        assert_rays_match(r.to_array(), ref, what="transform")


def test_bundle_statistics(ctx):
    """rays.rs:521-701 + reference KATs rays.rs:2019-2106, 2696-2734"""
    rays = make_bundle(100_000, radius=mm(3.0))
    rays["pos"][:, 0] += mm(1.5)
    rays["valid"][::17] = 0
    rays["bounces"] = 2
    r = ob.Rays.from_array(ctx, rays)
    s = r.spot_stats()
    assert s.n_valid == orc.nr_of_rays(rays) and s.n_total == len(rays) and s.bounce_lvl == orc.bounce_lvl(rays) == 2
    assert s.total_energy == pytest.approx(orc.total_energy(rays), rel=RTOL)
    assert np.allclose(s.centroid[:], orc.centroid(rays), rtol=RTOL, atol=1e-15)
    assert np.allclose(s.energy_centroid[:], orc.energy_weighted_centroid(rays), rtol=RTOL, atol=1e-15)
    assert s.beam_radius_geo == pytest.approx(orc.beam_radius_geo(rays), rel=RTOL)
    assert s.beam_radius_rms == pytest.approx(orc.beam_radius_rms(rays), rel=RTOL)
    assert s.energy_beam_radius_rms == pytest.approx(orc.energy_weighted_beam_radius_rms(rays), rel=RTOL)
    io, ig = iso_pair((mm(1.0), 0.0, mm(5.0)), (0.1, 0.2, 0.0))  # detector frame (nodes/spot_diagram.rs:106-112)
    loc = rays.copy()
    orc.rays_transform(loc, io, inverse=True)
    s = r.spot_stats(ig)
    assert np.allclose(s.energy_centroid[:], orc.energy_weighted_centroid(loc), rtol=RTOL, atol=1e-15)
    assert s.beam_radius_geo == pytest.approx(orc.beam_radius_geo(loc), rel=RTOL)
    # reference KATs
    k = np.concatenate([orc.ray_new(mm(1.0, 2.0, 0.0), Z, nm(1053.0), 1.0), orc.ray_new(mm(2.0, 3.0, 0.0), Z, nm(1053.0), 1.0),
                        orc.ray_new(mm(2.0, 3.0, 0.0), Z, nm(1053.0), 1.0)])
    k[2]["valid"] = 0
    rk = ob.Rays.from_array(ctx, k)
    assert tuple(rk.centroid()) == mm(1.5, 2.5, 0.0) and rk.beam_radius_geo() == mm(math.sqrt(0.5))
    assert rk.total_energy() == 2.0 and rk.nr_of_rays() == 2 and rk.nr_of_rays(False) == 3
    k = np.concatenate([orc.ray_new(mm(-1.0, 0.0, 0.0), Z, nm(1054.0), 1.0), orc.ray_new(mm(1.0, 0.0, 0.0), Z, nm(1054.0), 0.5)])
    assert ob.Rays.from_array(ctx, k).energy_weighted_centroid()[0] / 1e-3 == pytest.approx(-1.0 / 3.0, rel=4 * EPS)
    empty = ob.Rays(ctx, 1)
    assert empty.centroid() is None and empty.beam_radius_geo() is None and empty.total_energy() == 0.0
    counts, energy = r.tally(4)
    assert counts[2] == orc.nr_of_rays(rays) and counts.sum() == counts[2]
    assert energy[2] == pytest.approx(orc.total_energy(rays), rel=RTOL)


def test_upload_download_roundtrip_and_soa(ctx):
    rays = make_bundle(1000)
    rays["valid"][::3] = 0
    r = ob.Rays.from_array(ctx, rays)
    back = r.to_array()
    assert back.tobytes() == rays.tobytes()
    c = r.clone(); c.add_rays(r)
    assert len(c) == 2000 and c.to_array()[1000:].tobytes() == rays.tobytes()
    soa = ob.Rays(ctx, 1000)
    pos = [np.ascontiguousarray(rays["pos"][:, k]) for k in range(3)]
    d = [np.ascontiguousarray(2.0 * rays["dir"][:, k]) for k in range(3)]  # Ray::new normalises (ray.rs:127)
    soa.upload_soa(pos, d, np.ascontiguousarray(rays["e"]), np.ascontiguousarray(rays["wvl"]))
    got = soa.to_array()
    fresh = rays.copy(); fresh["valid"] = 1
    assert_rays_match(got, fresh, what="soa upload")
    assert len(ob.Rays.from_array(ctx, rays[:0]).to_array()) == 0  # empty bundle


def test_new_collimated_with_spectrum_from_host_buffers(ctx):
    """Rays::new_collimated_with_spectrum (rays.rs:264-297) from host positions / energies: position-major, wavelength-minor,
    E = E_pos * w.  The copy runs on the library's copy stream and is expanded at first use; re-uploading into a bundle
    that was just traced (double buffering) must give the new contents."""
    xyz = np.ascontiguousarray(orc.fibonacci_ellipse(3000, mm(5.0), mm(4.0)))
    e = np.ascontiguousarray(orc.energy_general_gaussian(xyz, 2.0, (0.0, 0.0), (mm(3.0), mm(3.0)), 2.0, 0.0, False))
    wvl, w = [nm(900.0), nm(1000.0), nm(1100.0)], [0.2, 0.5, 0.3]
    ref = orc.rays_collimated_with_spectrum(xyz, e, wvl, w)
    bufs = [ob.Rays(ctx, 1), ob.Rays(ctx, 1)]
    surf = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(30.0))))
    for k in range(4):
        r = bufs[k % 2]
        scale = 1.0 + k
        ek = np.ascontiguousarray(e * scale)
        r.new_collimated_with_spectrum(xyz.ctypes.data, ek.ctypes.data, len(xyz), wvl, w)
        assert len(r) == len(ref) == 9000
        got = r.to_array()
        exp = ref.copy()
        exp["e"] *= scale
        assert_rays_match(got, exp, what=f"upload {k}")
        assert np.array_equal(got["wvl"], ref["wvl"]) and np.array_equal(got["pos"], ref["pos"])
        r.refract_on_surface(surf[1], None, want_reflected=False)  # leaves the bundle modified before its next upload
    with pytest.raises(ob.OpossumError):
        bufs[0].new_collimated_with_spectrum(xyz.ctypes.data, e.ctypes.data, len(xyz), [0.0], [1.0])  # Ray::new: wavelength must be > 0


@pytest.mark.parametrize("energy", ["gaussian", "gaussian_known_norm", "uniform"])
def test_new_collimated_xy_matches_oracle(ctx, energy):
    """opm_rays_new_collimated_xy: the position distribution's (x, y) from the host (16 B per position), z = 0, the energy
    distribution (general_gaussian.rs:93-120 / uniform.rs:37-41) evaluated on the device -- against the oracle's
    Rays::new_collimated_with_spectrum (rays.rs:264-297) of the same points, twice into the same bundle."""
    xyz = np.ascontiguousarray(orc.fibonacci_ellipse(3001, mm(5.0), mm(4.0)))
    assert np.all(xyz[:, 2] == 0.0)
    wvl, w = [nm(900.0), nm(1000.0), nm(1100.0)], [0.2, 0.5, 0.3]
    sd = ob.SourceDesc(len(xyz), wvl, w, total_energy=2.0)
    if energy == "uniform":
        e = orc.energy_uniform(len(xyz), 2.0)
    else:
        e = np.ascontiguousarray(orc.energy_general_gaussian(xyz, 2.0, (mm(0.3), 0.0), (mm(3.0), mm(2.5)), 2.0, 0.2, False))
        sd.general_gaussian((mm(0.3), 0.0), (mm(3.0), mm(2.5)), 2.0, 0.2, False)
        if energy == "gaussian_known_norm":  # the sum over the whole source, as a sharded caller passes it
            sd2 = ob.SourceDesc(len(xyz), wvl, w, total_energy=2.0).fibonacci_ellipse(mm(5.0), mm(4.0))
            sd2.general_gaussian((mm(0.3), 0.0), (mm(3.0), mm(2.5)), 2.0, 0.2, False)
            sd.general_gaussian((mm(0.3), 0.0), (mm(3.0), mm(2.5)), 2.0, 0.2, False, norm=sd2.energy_norm(ctx))
    ref = orc.rays_collimated_with_spectrum(xyz, e, wvl, w)
    xy = np.ascontiguousarray(xyz[:, :2])
    r = ob.Rays(ctx, 1)
    surf = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(30.0))))
    for k in range(2):
        r.new_collimated_xy(xy.ctypes.data, len(xy), sd)
        assert len(r) == len(ref) == 9003
        got = r.to_array()
        assert_rays_match(got, ref, what=f"xy upload {k}")
        assert np.array_equal(got["wvl"], ref["wvl"]) and np.array_equal(got["pos"], ref["pos"])
        assert abs(got["e"].sum() - 2.0) < 1e-12
        r.refract_on_surface(surf[1], None, want_reflected=False)
    # a shard of the positions with the whole source's normalisation = the matching slice of the whole bundle
    if energy == "gaussian_known_norm":
        half = np.ascontiguousarray(xy[1000:2000])
        r.new_collimated_xy(half.ctypes.data, len(half), sd)
        got = r.to_array()
        assert np.allclose(got["e"], ref["e"][3000:6000], rtol=1e-12, atol=0)


def test_source_generate_strided_is_a_slice_of_the_source(ctx):
    """opm_source_generate_strided: rank g of G gets positions g, g+G, ... of the same global source (ids stay global)"""
    sd = ob.SourceDesc(1001, [nm(1000.0), nm(1100.0)], [0.4, 0.6], total_energy=2.0).fibonacci_ellipse(mm(6.0), mm(5.0))
    sd.general_gaussian((0.0, 0.0), (mm(3.0), mm(3.0)), 3.0)
    # the normalisation sum is a parallel reduction (order-dependent in the last digit): fix it once for all shards
    sd.general_gaussian((0.0, 0.0), (mm(3.0), mm(3.0)), 3.0, norm=sd.energy_norm(ctx))
    full = sd.generate(ctx).to_array().reshape(1001, 2)
    for g, G in ((0, 3), (2, 3), (7, 8)):
        first, count, stride = g, len(range(g, 1001, G)), G
        got = sd.generate_strided(ctx, first, count, stride).to_array().reshape(count, 2)
        ref = full[g::G]
        for f in ("pos", "dir", "e", "wvl", "id", "valid"):
            assert np.array_equal(got[f], ref[f]), f
    with pytest.raises(ob.OpossumError):
        sd.generate_strided(ctx, 5, 1000, 3)


# ------------------------------------------------------------------ reflective grating
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
@pytest.mark.parametrize("order", [-1, 1, 0, -2])
def test_diffract_on_periodic_surface(ctx, order, strict):
    """Rays::diffract_on_periodic_surface (rays.rs:1058-1111) on a tilted, decentred grating: a divergent polychromatic
    bundle, supported and evanescent orders mixed, invalid rays carried along; diffracted clones, parents (zero energy) and
    counts against the oracle; the bundle-continues form (ReflectiveGrating node) against the clones"""
    dens = 600e3
    node = iso_pair((mm(0.3), mm(-0.2), mm(60.0)), (0.05, 0.4, 0.1))
    so, sg = surface_pair(F.GEOM_PLANE, iso=node)
    gvec = [2 * math.pi * dens * c for c in node[0].transform_vector((1.0, 0.0, 0.0))]
    rays = make_bundle(3000, divergence=0.25, wvl=(nm(500.0), nm(1600.0)))
    rays["valid"][::13] = 0
    ref = rays.copy()
    ch_ref, st_ref = orc.rays_diffract_on_periodic_surface(ref, so, orc.IndexModel.const(1.0), gvec, order, False)
    r = ob.Rays.from_array(ctx, rays)
    out = ob.Rays(ctx, len(rays))
    prog = ob.Program().diffract(sg, ob.RefractiveIndex.const(1.0), gvec, order, children=out, continue_as_child=False)
    st = prog.run(r, strict=strict)
    assert st.interactions == st_ref.n_interactions and st.children == len(ch_ref)
    if order == -2:
        assert 0 < len(ch_ref) < st_ref.n_interactions  # some wavelengths are evanescent in second order
    assert_rays_match(out.to_array(), ch_ref, keyed=True, what=f"diffracted order {order}", e_floor=1e-25)
    assert_rays_match(r.to_array(), ref, what=f"parents order {order}", e_floor=1e-25)
    r2 = ob.Rays.from_array(ctx, rays)
    st2 = ob.Program().diffract(sg, ob.RefractiveIndex.const(1.0), gvec, order).run(r2, strict=strict)
    assert st2.alive_out == len(ch_ref)
    assert_rays_match(r2.to_array(), ch_ref, what=f"bundle continues, order {order}", e_floor=1e-25)


def test_diffract_index_error(ctx):
    """the medium behind the grating is only validated: an index < 1 is already refused by the index model
    (refractive_index/mod.rs:54-58, called at rays.rs:1071) before ray.rs:489-493 sees it"""
    node = iso_pair((0.0, 0.0, mm(10.0)))
    so, sg = surface_pair(F.GEOM_PLANE, iso=node)
    r = ob.Rays.from_array(ctx, make_bundle(64))
    with pytest.raises(ob.OpossumError) as e:
        r.diffract_on_periodic_surface(sg, ob.RefractiveIndex.const(0.5), (1e6, 0.0, 0.0), -1)
    assert e.value.code == F.OPM_ERR_INDEX_MODEL


def test_fork_join_equals_split_and_separate_trace(ctx):
    """OPM_OP_FORK .. OPM_OP_JOIN (the split-off bundle of Rays::split, rays.rs:1253-1264, traced inside the same launch, its
    parent parked in shared memory) against OPM_OP_SPLIT followed by a separate launch over the split-off bundle, in both
    ways of running a program; main bundle, side bundle, hit records and counts must agree exactly"""
    node = iso_pair((0.0, 0.0, mm(30.0)))
    front = node_sphere_pair(mm(150.0), node, coating=F.COAT_FRESNEL)[1]
    side_surf = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(50.0)), (0.02, -0.01, 0.0)), coating=F.COAT_CONSTANT_R, r=0.3)[1]
    main_surf = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(70.0))))[1]
    glass = ob.RefractiveIndex.const(1.5)
    ap = ob.Aperture.circle(mm(6.0))
    rays = make_bundle(5000, divergence=0.05)
    rays["valid"][::9] = 0
    for jit in (0, 2):
        ctx.set_jit(jit, 1)
        # reference form: split, then the side chain as its own launch
        a = ob.Rays.from_array(ctx, rays)
        side_a = ob.Rays(ctx, len(rays))
        hm_a = ob.HitMap(ctx, len(rays))
        st_a = (ob.Program().refract(front, glass).threshold(1e-12).split(0.3, side_a).refract(main_surf, None).limits(10, 10)).run(a)
        st_a2 = (ob.Program().apodize(ap, ob.Isometry.identity()).threshold(1e-12).refract(side_surf, None, hit_map=hm_a).limits(10, 10)).run(side_a)
        # fused form
        b = ob.Rays.from_array(ctx, rays)
        side_b = ob.Rays(ctx, len(rays))
        hm_b = ob.HitMap(ctx, len(rays))
        st_b = (ob.Program().refract(front, glass).threshold(1e-12)
                .fork(0.3).apodize(ap, ob.Isometry.identity()).threshold(1e-12).refract(side_surf, None, hit_map=hm_b).limits(10, 10).join(side_b)
                .refract(main_surf, None).limits(10, 10)).run(b)
        assert st_b.interactions == st_a.interactions + st_a2.interactions and st_b.hits == st_a2.hits
        assert st_b.valid_out == st_a.valid_out and st_b.alive_out == st_a.alive_out and st_b.apodized == st_a2.apodized
        ga, gb = a.to_array(), b.to_array()
        sa, sb = side_a.to_array(), side_b.to_array()
        # the interpreter runs the same arithmetic in both forms: bit-identical.  The specialised kernel knows inside the fused program
        # that the side chain's rays already carry unit directions (RefractCt UNIT_DIR) and skips a renormalisation the separate
        # launch makes: last-bit differences
        same = np.array_equal if jit == 0 else (lambda p, q: np.allclose(p, q, rtol=1e-13, atol=1e-300))
        for f in ga.dtype.names:
            assert same(ga[f], gb[f]), f"main bundle field {f} (jit={jit})"
            assert same(sa[f], sb[f]), f"side bundle field {f} (jit={jit})"
        for u, w in zip(hm_a.download(), hm_b.download()):
            assert same(u, w)
    ctx.set_jit(1, 200000)
    with pytest.raises(ob.OpossumError):
        ob.Program().fork(0.5).refract(main_surf, None).run(ob.Rays.from_array(ctx, rays))  # FORK without JOIN
    with pytest.raises(ob.OpossumError):
        ob.Program().join(None).run(ob.Rays.from_array(ctx, rays))  # JOIN without FORK
