"""GPU parity of the fluence helper rays (FluenceRays, rays.rs:1470-1567): three satellite rays per ray carried through
refractions, mirrors and paraxial surfaces, FluenceHitPoints in the hit map, and the helper-ray fluence estimator
(rays_hit_map.rs:627-722) -- CUDA path through the C ABI against the oracle's restatement (oracle.helpers_* /
rays_refract_on_surface_w_helpers / fluence_helper_rays) on the same inputs, at the north star's 1e-9."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import RTOL, assert_rays_match, index_pair, iso_pair, make_bundle, mm, nm, node_sphere_pair, surface_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def assert_helpers_match(got: dict, ref: dict, what=""):
    """positions relative to the bundle's extent, directions absolute, indices exact in sign (validity), effective energies relative"""
    assert got["pos"].shape == ref["pos"].shape, (what, got["pos"].shape, ref["pos"].shape)
    if len(ref["eff"]) == 0:
        return
    scale = max(np.abs(ref["pos"]).max(), 1e-3)
    assert np.abs(got["pos"] - ref["pos"]).max() / scale <= RTOL, f"{what}: helper positions {np.abs(got['pos'] - ref['pos']).max() / scale:.3e}"
    assert np.abs(got["dir"] - ref["dir"]).max() <= RTOL, f"{what}: helper directions {np.abs(got['dir'] - ref['dir']).max():.3e}"
    assert np.array_equal(np.signbit(got["n"]), np.signbit(ref["n"])), f"{what}: helper validity"
    assert np.allclose(got["n"], ref["n"], rtol=RTOL, atol=0), f"{what}: helper indices"
    assert np.allclose(got["eff"], ref["eff"], rtol=RTOL, atol=1e-300), f"{what}: effective energies"


def helper_areas(h):
    ab, ac = h["pos"][:, 0] - h["pos"][:, 1], h["pos"][:, 0] - h["pos"][:, 2]
    return 0.5 * np.linalg.norm(np.cross(ab, ac), axis=1)


def test_attach_and_download(ctx):
    rays = make_bundle(500, divergence=0.03)
    fl = 1.0e4 * (1.0 + 0.5 * np.cos(np.arange(len(rays))))
    ref = orc.helpers_new(rays, fl)
    r = ob.Rays.from_array(ctx, rays)
    assert not r.has_fluence_helpers
    r.attach_fluence_helpers(fl)
    assert r.has_fluence_helpers
    got = r.fluence_helpers()
    assert_helpers_match(got, orc.helpers_as_product_layout(ref), "attach")
    assert np.allclose(got["eff"] / helper_areas(got), fl, rtol=1e-6)  # Ray::helper_ray_fluence: what went in (area wavelength^2, 1e-12 m^2 next to 1e-3 m positions)
    c = r.clone()  # Rays::clone clones the FluenceRays
    assert_helpers_match(c.fluence_helpers(), got, "clone")
    with pytest.raises(ob.OpossumError):
        ob.Rays.from_array(ctx, rays).attach_fluence_helpers(-1.0)
    with pytest.raises(ob.OpossumError):
        ob.Rays.from_array(ctx, rays).attach_fluence_helpers(math.nan)


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_lens_and_detector_with_helpers(ctx, strict):
    """a lens (two Fresnel spheres, Sellmeier-free constant glass) and a detector plane, in ONE fused launch: rays, helper rays,
    effective energies and the three surfaces' FluenceHitPoints against the oracle; the focused beam's fluence rises"""
    rays = make_bundle(600, radius=mm(6.0), divergence=0.0, wvl=(nm(1000.0), nm(1000.0)))
    fl = np.full(len(rays), 3.0e4)
    lens_iso = iso_pair((0.0, 0.0, mm(20.0)))
    front = node_sphere_pair(mm(80.0), lens_iso, coating=F.COAT_FRESNEL)
    rear_iso = (lens_iso[0].append(orc.Isometry.new((0, 0, mm(-80.0) + mm(3.0)))), lens_iso[1].append(ob.Isometry.new((0, 0, mm(-80.0) + mm(3.0)))))
    rear = surface_pair(F.GEOM_SPHERE, mm(-80.0), rear_iso, coating=F.COAT_FRESNEL)
    det = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(60.0)), (0.02, -0.01, 0.0)))
    glass, vac = index_pair("const", 1.5), index_pair("const", 1.0)
    # oracle
    ref = rays.copy()
    rh = orc.helpers_new(ref, fl)
    _, _, h0 = orc.rays_refract_on_surface_w_helpers(ref, rh, front[0], glass[0])
    _, _, h1 = orc.rays_refract_on_surface_w_helpers(ref, rh, rear[0], vac[0])
    _, _, h2 = orc.rays_refract_on_surface_w_helpers(ref, rh, det[0], None)
    # product
    r = ob.Rays.from_array(ctx, rays).attach_fluence_helpers(fl)
    hm = [ob.HitMap(ctx, len(rays)) for _ in range(3)]
    st = (ob.Program().refract(front[1], glass[1], hit_map=hm[0]).threshold(1e-12).refract(rear[1], vac[1], hit_map=hm[1])
          .refract(det[1], None, hit_map=hm[2]).limits(10, 10)).run(r, strict=strict)
    assert st.interactions == 3 * len(rays) and st.hits == len(h0) + len(h1) + len(h2)
    assert_rays_match(r.to_array(), ref, what="rays with helpers")
    assert_helpers_match(r.fluence_helpers(), orc.helpers_as_product_layout(rh), "lens + detector")
    for m, h in zip(hm, (h0, h1, h2)):
        x, y, z, e, b = m.download()
        assert len(e) == len(h)
        assert np.allclose(x, h["x"], rtol=RTOL, atol=1e-12) and np.allclose(y, h["y"], rtol=RTOL, atol=1e-12)
        # the fluence is an effective energy over the area of a triangle of 1 um side next to positions of 1e-2 m: 1e-9 of
        # the positions is 1e-5 of the area -- the bar of this quotient is what the subtraction leaves
        assert np.allclose(e, h["e"], rtol=1e-5)
    # What a surface records is the fluence of the helper triangle as it was at the PREVIOUS surface (the helpers are refracted after
    # the ray, ray.rs:655 / rays.rs:996): the detector's record is the lens' rear fluence times the Fresnel losses, the
    # convergence only shows in the helpers the bundle leaves with
    assert np.median(h2["e"]) < np.median(h0["e"])
    got = r.fluence_helpers()
    assert np.median(got["eff"] / helper_areas(got)) > 1.5 * np.median(h0["e"])


def test_mirror_keeps_helpers_behind(ctx):
    """a mirror's continuing bundle is the reflected clone: its helpers stay where they were and take the reflected directions
    (rays.rs:996-1010), the effective energy takes R"""
    rays = make_bundle(300, radius=mm(5.0), divergence=0.01, wvl=(nm(800.0), nm(800.0)))
    fl = np.full(len(rays), 1.0e4)
    mir = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(30.0)), (0.3, 0.0, 0.0)), coating=F.COAT_CONSTANT_R, r=0.9)
    ref = rays.copy()
    rh = orc.helpers_new(ref, fl)
    before = orc.helpers_as_product_layout(rh)
    ch, chh, hits = orc.rays_refract_on_surface_w_helpers(ref, rh, mir[0], None, refraction_intended=False)
    r = ob.Rays.from_array(ctx, rays).attach_fluence_helpers(fl)
    hm = ob.HitMap(ctx, len(rays))
    ob.Program().mirror(mir[1], hit_map=hm).run(r)
    assert_rays_match(r.to_array(), ch, what="mirror: continuing bundle")
    got = r.fluence_helpers()
    assert_helpers_match(got, orc.helpers_as_product_layout(chh), "mirror")
    assert np.array_equal(got["pos"], before["pos"]) and np.allclose(got["eff"], before["eff"] * 0.9, rtol=1e-14)
    assert np.all(got["dir"][:, :, 2] < 0.0)


def test_paraxial_surface_moves_the_helpers(ctx):
    rays = make_bundle(200, radius=mm(4.0), divergence=0.0, wvl=(nm(1053.0), nm(1053.0)))
    rays["valid"][::7] = 0  # the helpers of an invalid ray are refracted all the same (rays.rs:953-955)
    fl = np.full(len(rays), 5.0e3)
    iso = iso_pair((0.0, 0.0, mm(10.0)))
    plane = surface_pair(F.GEOM_PLANE, iso=iso)
    ref = rays.copy()
    rh = orc.helpers_new(ref, fl)
    orc.rays_refract_on_surface_w_helpers(ref, rh, plane[0], None)
    orc.rays_refract_paraxial_w_helpers(ref, rh, mm(100.0), iso[0])
    r = ob.Rays.from_array(ctx, rays).attach_fluence_helpers(fl)
    ob.Program().refract(plane[1], None).paraxial(mm(100.0), iso[1]).run(r)
    assert_rays_match(r.to_array(), ref, what="paraxial with helpers")
    assert_helpers_match(r.fluence_helpers(), orc.helpers_as_product_layout(rh), "paraxial")


def test_helper_rays_estimator_matches_oracle(ctx):
    """RaysHitMap::calc_fluence_with_helper_rays on a scattered hit set with varying fluences: the reference's known answer
    (hit_map/mod.rs:1030-1068) and a random set against the oracle, NaN pattern and values"""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    hm = ob.HitMap(ctx, 5)
    hm.upload(np.array([p[0] for p in p1]), np.array([p[1] for p in p1]), np.full(5, 1.0e4), bounce=np.ones(5, dtype=np.uint32))
    fd = ob.fluence_helper_rays(ctx, [hm], 51)
    assert fd.map[25, 25] == pytest.approx(10000.0, rel=1e-12)
    rng = np.random.default_rng(11)
    n = 400
    x, y = rng.normal(0.0, 2e-3, n), rng.normal(0.0, 1e-3, n)
    e = 1.0e4 * np.exp(-0.5 * ((x / 2e-3) ** 2 + (y / 1e-3) ** 2)) + 10.0
    hits = orc.hits_from_xye(x, y, e)
    ref, lrtb = orc.fluence_helper_rays(hits, 40)
    hm = ob.HitMap(ctx, n)
    hm.upload(x, y, e)
    fd = ob.fluence_helper_rays(ctx, [hm], 40)
    assert np.allclose(fd.lrtb, lrtb, rtol=1e-12)
    assert np.array_equal(np.isnan(fd.map), np.isnan(ref)) and not np.isnan(ref).any()  # the four far sites close the hull
    assert np.abs(fd.map - ref).max() / np.abs(ref).max() <= RTOL
    with pytest.raises(ob.OpossumError):
        two = ob.HitMap(ctx, 2)
        two.upload(np.zeros(2), np.ones(2), np.ones(2))
        ob.fluence_helper_rays(ctx, [two], 10)


def test_what_a_bundle_with_helpers_refuses(ctx):
    rays = make_bundle(100)
    plane = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(10.0))))[1]
    r = ob.Rays.from_array(ctx, rays).attach_fluence_helpers(1.0)
    with pytest.raises(ob.OpossumError):
        ob.Program().refract(plane, None).run(r, out=ob.Rays(ctx, 100), compact=True)
    with pytest.raises(ob.OpossumError):
        ob.Program().refract(plane, None).split(0.5, ob.Rays(ctx, 100)).run(r)
    with pytest.raises(ob.OpossumError):
        ob.Program().refract(plane, None, children=ob.Rays(ctx, 100)).run(r)


def test_scene_with_helper_rays_and_report(ctx):
    """the scene layer with a helper-ray bundle: source -> lens -> FluenceDetector, FluenceEstimator::HelperRays report
    (HitMap::calc_combined_fluence_with_helper_rays) against the oracle run surface by surface"""
    import scenes
    MM = scenes.MM
    nodes = [scenes.node("source"),
             scenes.node("lens", z=30 * MM, p=(120 * MM, -120 * MM, 4 * MM), index=("const", 1.5), coat_in=("fresnel",), coat_out=("fresnel",)),
             scenes.node("fluence_detector", z=60 * MM)]
    xyz = orc.hexapolar(6, 6e-3)
    n = len(xyz)
    fl = orc.fluence_general_gaussian(xyz[:, :2], 1.0, (0.0, 0.0), (3e-3, 3e-3), 0.0)
    rays = orc.rays_from_arrays(xyz, np.tile([0.0, 0.0, 1.0], (n, 1)), np.full(n, 1.0 / n), np.full(n, 1000e-9))
    psc = scenes.to_product(ctx, nodes)
    r = ob.Rays.from_array(ctx, rays).attach_fluence_helpers(fl)
    st = psc.raytrace(r)
    fd = psc.node_fluence(2, 31, 31, estimator="helper_rays")
    # oracle: the same surfaces one by one (lens front, lens rear, detector plane), then the estimator on the detector's records
    osc = scenes.to_oracle(nodes)
    ref = rays.copy()
    rh = orc.helpers_new(ref, fl)
    lens, det = osc.nodes[1], osc.nodes[2]
    glass = orc.IndexModel.const(1.5)
    orc.rays_refract_on_surface_w_helpers(ref, rh, lens.s[0].surf, glass)
    orc.rays_refract_on_surface_w_helpers(ref, rh, lens.s[1].surf, orc.IndexModel.const(1.0))
    _, _, hits = orc.rays_refract_on_surface_w_helpers(ref, rh, det.s[0].surf, None)
    assert st.interactions == 3 * n and len(hits) == n
    assert_rays_match(r.to_array(), ref, what="scene with helpers")
    assert_helpers_match(r.fluence_helpers(), orc.helpers_as_product_layout(rh), "scene")
    ref_map, lrtb = orc.fluence_helper_rays_combined([hits], 31, 31)
    assert np.allclose(fd.lrtb, lrtb, rtol=RTOL, atol=0)
    ok = np.isfinite(ref_map)
    assert np.array_equal(np.isfinite(fd.map), ok)
    assert np.abs(fd.map[ok] - ref_map[ok]).max() / np.abs(ref_map[ok]).max() <= 1e-5  # the records are quotients by 1e-12 m^2 areas, see above
    assert fd.peak == pytest.approx(np.nanmax(ref_map), rel=1e-5)


def test_new_collimated_w_fluence_helper(ctx):
    """Rays::new_collimated_w_fluence_helper (rays.rs:352-398): ray energies = Voronoi cell areas x fluences; the bundle's energy
    approaches the distribution's total (1 J) as the reference's docstring promises"""
    xyz = orc.hexapolar(7, 12e-3)
    fl = orc.fluence_general_gaussian(xyz[:, :2], 1.0, (0.0, 0.0), (2.5e-3, 2.5e-3), 0.0)
    ref, rh = orc.rays_new_collimated_w_fluence_helper(xyz, 1000e-9, fl)
    r = ob.Rays(ctx, len(xyz)).new_collimated_w_fluence_helper(xyz, 1000e-9, fl)
    got = r.to_array()
    assert_rays_match(got, ref, what="new_collimated_w_fluence_helper")
    assert_helpers_match(r.fluence_helpers(), orc.helpers_as_product_layout(rh), "new_collimated_w_fluence_helper")
    assert abs(got["e"].sum() - 1.0) < 0.05
