"""Known-answer tests that PIN THE ORACLE to the reference.

Every test is a transcription of one of the reference's own unit tests for the ray-propagation
path (file:line under /root/reference/opossum/src given in each docstring).  `assert_eq!` in the
reference is transcribed as exact equality, `assert_abs_diff_eq!` as |a-b| <= f64::EPSILON (its
default), `assert_relative_eq!` as relative f64::EPSILON unless the test states a tolerance.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

EPS = np.finfo(np.float64).eps


def mm(*v):
    """uom `millimeter!(x)`: value * 1e-3 (base unit metre)."""
    r = tuple(x * 1.0e-3 for x in v)
    return r[0] if len(r) == 1 else r


def nm(x):
    return x * 1.0e-9


def plane(z_mm=0.0, coating=orc.COAT_IDEAL_AR, r=0.0):
    return orc.Surface(geom=orc.GEOM_PLANE, coating=coating, coating_r=r, iso=orc.Isometry.new_along_z(mm(z_mm)))


def sphere_at(radius, pos):
    """Sphere::new_at_position (sphere.rs:30-41): iso = T(pos.x, pos.y, pos.z + radius)."""
    return orc.Surface(geom=orc.GEOM_SPHERE, param=radius,
                       iso=orc.Isometry.new((pos[0], pos[1], pos[2] + radius)))


Z = (0.0, 0.0, 1.0)


# ---------------------------------------------------------------- ray.rs
def test_ray_new():
    """ray.rs:897-927"""
    r = orc.ray_new(mm(1.0, 2.0, 3.0), (0.0, 0.0, 2.0), nm(1053.0), 1.0)[0]
    assert tuple(r["pos"]) == mm(1.0, 2.0, 3.0)
    assert tuple(r["dir"]) == Z
    assert r["opl"] == 0.0 and r["n"] == 1.0 and r["valid"] == 1 and r["bounces"] == 0 and r["refr"] == 0
    for bad_wvl in (0.0, nm(-10.0), math.nan, math.inf, -math.inf):
        with pytest.raises(orc.OracleError):
            orc.ray_new((0, 0, 0), Z, bad_wvl, 1.0)
    for bad_e in (-0.1, math.nan, math.inf):
        with pytest.raises(orc.OracleError):
            orc.ray_new((0, 0, 0), Z, nm(1053.0), bad_e)
    with pytest.raises(orc.OracleError):
        orc.ray_new((0, 0, 0), (0.0, 0.0, 0.0), nm(1053.0), 1.0)
    orc.ray_new((0, 0, 0), Z, nm(1053.0), 0.0)  # ray.rs:948 zero energy is ok


def test_refract_on_surface_collimated():
    """ray.rs:1203-1290: plane z=10 mm, n 1 -> 1.5, ConstantR 0.2"""
    s = plane(10.0, orc.COAT_CONSTANT_R, 0.2)
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1054.0), 1.0)
    for bad in (0.9, math.nan, math.inf):
        with pytest.raises(orc.OracleError):
            orc.ray_refract_on_surface(ray.copy(), s, bad)
    child, hit = orc.ray_refract_on_surface(ray, s, 1.5)
    r, c = ray[0], child[0]
    assert tuple(r["pos"]) == mm(0.0, 0.0, 10.0)
    assert r["n"] == 1.5
    assert tuple(r["dir"]) == Z
    assert r["opl"] == mm(10.0)
    assert r["bounces"] == 0 and r["refr"] == 1
    assert r["e"] == (1.0 - 0.2) * 1.0
    assert tuple(c["pos"]) == mm(0.0, 0.0, 10.0)
    assert c["n"] == 1.0
    assert tuple(c["dir"]) == (0.0, 0.0, -1.0)
    assert c["opl"] == mm(10.0)
    assert c["bounces"] == 1 and c["refr"] == 0
    assert c["e"] == 0.2 * 1.0
    # hit map entry: surface frame, INPUT energy, filed under the refracted ray's bounce count (ray.rs:640-654)
    assert (hit[0]["x"], hit[0]["y"], hit[0]["z"], hit[0]["e"], hit[0]["bounce"]) == (0.0, 0.0, 0.0, 1.0, 0)
    ray = orc.ray_new(mm(0.0, 1.0, 0.0), Z, nm(1054.0), 1.0)
    orc.ray_refract_on_surface(ray, s, 1.5)
    assert tuple(ray[0]["pos"]) == mm(0.0, 1.0, 10.0)
    assert tuple(ray[0]["dir"]) == Z
    assert ray[0]["opl"] == mm(10.0)
    assert ray[0]["bounces"] == 0 and ray[0]["refr"] == 1


def test_refract_on_surface_same_index():
    """ray.rs:1291-1318: passive surface, OPL = n * sqrt(2) * z"""
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 1.0, 1.0), nm(1054.0), 1.0)
    ray["n"] = 2.0
    orc.ray_refract_on_surface(ray, plane(10.0), None)
    d = np.array([0.0, 1.0, 1.0]) / math.sqrt(2.0)
    assert tuple(ray[0]["pos"]) == mm(0.0, 10.0, 10.0)
    assert ray[0]["dir"][0] == 0.0
    assert abs(ray[0]["dir"][1] - d[1]) <= EPS and abs(ray[0]["dir"][2] - d[2]) <= EPS
    assert abs(ray[0]["opl"] - 2.0 * math.sqrt(2.0) * mm(10.0)) <= EPS
    assert ray[0]["refr"] == 0  # n2 = None does not count as a refraction (ray.rs:636-638)


def test_refract_on_surface_non_intersecting():
    """ray.rs:1319-1346"""
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), nm(1054.0), 1.0)
    child, hit = orc.ray_refract_on_surface(ray, plane(10.0), 1.5, orc.MISS_STOP)
    r = ray[0]
    assert child is None and hit is None
    assert tuple(r["pos"]) == (0.0, 0.0, 0.0) and tuple(r["dir"]) == (0.0, 0.0, -1.0)
    assert r["n"] == 1.0 and r["opl"] == 0.0 and r["bounces"] == 0 and r["refr"] == 0
    assert r["valid"] == 0  # ray.rs:683 (Stop); not asserted by the reference test
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), nm(1054.0), 1.0)
    orc.ray_refract_on_surface(ray, plane(10.0), 1.5, orc.MISS_IGNORE)
    assert ray[0]["valid"] == 1


def test_refract_on_surface_non_collimated():
    """ray.rs:1347-1428: dir = (0, 0.4714045207910317, 0.8819171036881969) at 45 deg into n=1.5"""
    s = plane(10.0)
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 1.0, 1.0), nm(1054.0), 1.0)
    orc.ray_refract_on_surface(ray, s, 1.0)
    d = np.array([0.0, 1.0, 1.0]) / math.sqrt(2.0)
    assert tuple(ray[0]["pos"]) == mm(0.0, 10.0, 10.0)
    assert ray[0]["dir"][0] == 0.0
    assert abs(ray[0]["dir"][1] - d[1]) <= EPS and abs(ray[0]["dir"][2] - d[2]) <= EPS
    assert abs(ray[0]["opl"] - math.sqrt(2.0) * mm(10.0)) <= EPS
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 1.0, 1.0), nm(1054.0), 1.0)
    orc.ray_refract_on_surface(ray, s, 1.5)
    assert ray[0]["bounces"] == 0 and ray[0]["refr"] == 1
    assert tuple(ray[0]["pos"]) == mm(0.0, 10.0, 10.0)
    assert ray[0]["dir"][0] == 0.0
    assert abs(ray[0]["dir"][1] - 0.4714045207910317) <= EPS
    assert abs(ray[0]["dir"][2] - 0.8819171036881969) <= EPS
    ray = orc.ray_new((0.0, 0.0, 0.0), (1.0, 0.0, 1.0), nm(1054.0), 1.0)
    orc.ray_refract_on_surface(ray, s, 1.5)
    assert tuple(ray[0]["pos"]) == mm(10.0, 0.0, 10.0)
    assert ray[0]["dir"][0] == 0.4714045207910317  # assert_eq! in the reference (ray.rs:1425)
    assert abs(ray[0]["dir"][1]) <= EPS
    assert abs(ray[0]["dir"][2] - 0.8819171036881969) <= EPS


def test_refract_on_surface_total_reflection():
    """ray.rs:1429-1458"""
    ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 2.0, 1.0), nm(1054.0), 1.0)
    ray["n"] = 1.5
    child, hit = orc.ray_refract_on_surface(ray, plane(10.0), 1.0)
    assert child is None and hit is None
    assert tuple(ray[0]["pos"]) == mm(0.0, 20.0, 10.0)
    t = np.array([0.0, 2.0, -1.0]) / math.sqrt(5.0)
    assert np.all(np.abs(ray[0]["dir"] - t) <= EPS)
    assert ray[0]["bounces"] == 1 and ray[0]["e"] == 1.0 and ray[0]["n"] == 1.5


def test_filter_and_split():
    """ray.rs:1459-1473, 1503-1515"""
    ray = orc.ray_new(mm(0.0, 1.0, 0.0), Z, nm(1054.0), 1.0)
    orc.rays_filter_energy(ray, 0.3)
    assert ray[0]["e"] == 0.3
    for bad in (-0.1, 1.1):
        with pytest.raises(orc.OracleError):
            orc.rays_filter_energy(ray, bad)
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1054.0), 1.0)
    for bad in (1.1, -0.1):
        with pytest.raises(orc.OracleError):
            orc.rays_split(ray, bad)
    sp = orc.rays_split(ray, 0.1)
    assert ray[0]["e"] == 0.1 and sp[0]["e"] == 0.9
    assert tuple(ray[0]["pos"]) == tuple(sp[0]["pos"]) and tuple(ray[0]["dir"]) == tuple(sp[0]["dir"])


def test_refract_paraxial():
    """ray.rs:1142-1202 (refract_paraxial_collimated, refract_paraxial_recollimate)"""
    iso = orc.Isometry.identity()
    pos = mm(1.0, 2.0, 0.0)
    for f_mm, t in ((100.0, (-1.0, -2.0, 100.0)), (-100.0, (1.0, 2.0, 100.0))):
        ray = orc.ray_new(pos, Z, nm(1053.0), 1.0)
        orc.rays_refract_paraxial(ray, mm(f_mm), iso)
        t = np.array(t) / np.sqrt(np.dot(t, t))
        assert tuple(ray[0]["pos"]) == pos
        assert np.all(np.abs(ray[0]["dir"] - t) <= EPS)
        assert ray[0]["refr"] == 1
    ray = orc.ray_new(mm(0.0, 10.0, 0.0), Z, nm(1053.0), 1.0)
    orc.rays_refract_paraxial(ray, mm(10.0), iso)
    assert abs(ray[0]["opl"] / 1e-3 - (-1.0 * (math.sqrt(200.0) - 10.0))) <= 10 * EPS
    pos = mm(0.0, 100.0, 0.0)
    ray = orc.ray_new(pos, Z, nm(1053.0), 1.0)
    orc.rays_refract_paraxial(ray, mm(100.0), iso)
    t = np.array([0.0, -100.0, 100.0]) / math.sqrt(20000.0)
    assert tuple(ray[0]["pos"]) == pos and np.all(np.abs(ray[0]["dir"] - t) <= EPS)
    pos = mm(0.0, 100.0, 100.0)
    ray = orc.ray_new(pos, (0.0, 1.0, 1.0), nm(1053.0), 1.0)
    orc.rays_refract_paraxial(ray, mm(100.0), iso)
    assert tuple(ray[0]["pos"]) == pos and tuple(ray[0]["dir"]) == Z
    ray = orc.ray_new(pos, (0.0, -1.0, 1.0), nm(1053.0), 1.0)
    orc.rays_refract_paraxial(ray, mm(-100.0), iso)
    assert tuple(ray[0]["pos"]) == pos and tuple(ray[0]["dir"]) == Z
    with pytest.raises(orc.OracleError):
        orc.rays_refract_paraxial(ray, 0.0, iso)


def test_transformed_ray():
    """ray.rs:1588-1651"""
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1053.0), 1.0)
    r = ray.copy()
    orc.rays_transform(r, orc.Isometry.new_along_z(1.0))
    assert tuple(r[0]["pos"]) == (0.0, 0.0, 1.0) and tuple(r[0]["dir"]) == Z
    r = ray.copy()
    orc.rays_transform(r, orc.Isometry.new((0, 0, 0), (0.0, math.radians(90.0), 0.0)))
    assert np.allclose(r[0]["dir"], (1.0, 0.0, 0.0), rtol=0, atol=EPS)
    r = orc.ray_new((1.0, 1.0, 1.0), Z, nm(1053.0), 1.0)
    orc.rays_transform(r, orc.Isometry.new((0.0, 1.0, 0.0), (0.0, 0.0, math.radians(90.0))))
    assert abs(r[0]["pos"][0] + 1.0) <= 2 * EPS and r[0]["pos"][1] == pytest.approx(2.0, rel=EPS)
    assert r[0]["pos"][2] == pytest.approx(1.0, rel=EPS)
    r = ray.copy()
    orc.rays_transform(r, orc.Isometry.new_along_z(1.0), inverse=True)
    assert tuple(r[0]["pos"]) == (0.0, 0.0, -1.0)
    r = ray.copy()
    orc.rays_transform(r, orc.Isometry.new((0, 0, 0), (0.0, math.radians(90.0), 0.0)), inverse=True)
    assert np.allclose(r[0]["dir"], (-1.0, 0.0, 0.0), rtol=0, atol=EPS)
    r = orc.ray_new((-1.0, 2.0, 1.0), Z, nm(1053.0), 1.0)
    orc.rays_transform(r, orc.Isometry.new((0.0, 1.0, 0.0), (0.0, 0.0, math.radians(90.0))), inverse=True)
    assert abs(r[0]["pos"][0] - 1.0) <= 5 * EPS and r[0]["pos"][1] == pytest.approx(1.0, rel=EPS)


# ---------------------------------------------------------------- surfaces
def test_plane_kats():
    """surface/plane.rs:118-165"""
    s = plane(10.0)
    q, n = orc.intersect(s, (0, 0, 0), Z)
    assert tuple(q) == mm(0.0, 0.0, 10.0) and tuple(n) == (0.0, 0.0, -1.0)
    assert orc.intersect(plane(-10.0), (0, 0, 0), Z) is None
    q, n = orc.intersect(plane(0.0), (0, 0, 0), Z)
    assert tuple(q) == (0.0, 0.0, 0.0) and tuple(n) == (0.0, 0.0, -1.0)
    q, n = orc.intersect(s, mm(0.0, 1.0, 1.0), Z)
    assert tuple(q) == mm(0.0, 1.0, 10.0) and tuple(n) == (0.0, 0.0, -1.0)
    d = np.array([0.0, 1.0, 1.0]) / np.sqrt(2.0)  # Ray::new normalises
    q, n = orc.intersect(s, mm(0.0, 1.0, 0.0), tuple(orc.ray_new(mm(0.0, 1.0, 0.0), (0.0, 1.0, 1.0), nm(1053.0), 1.0)[0]["dir"]))
    assert tuple(q) == mm(0.0, 11.0, 10.0) and tuple(n) == (0.0, 0.0, -1.0)


@pytest.mark.parametrize("radius_mm", [10.0, -10.0])
def test_sphere_on_axis_forward(radius_mm):
    """surface/sphere.rs:213-276"""
    s = sphere_at(mm(radius_mm), (0.0, 0.0, 0.0))
    q, n = orc.intersect(s, mm(0.0, 0.0, -5.0), Z)
    assert tuple(q) == (0.0, 0.0, 0.0) and tuple(n) == (0.0, 0.0, -1.0)
    q, _ = orc.intersect(s, mm(0.0, 0.0, -15.0), Z)
    assert np.all(np.abs(q) <= EPS)
    assert orc.intersect(s, mm(0.0, 0.0, 5.0), Z) is None
    assert orc.intersect(s, mm(0.0, 0.0, 15.0), Z) is None


@pytest.mark.parametrize("radius_mm", [10.0, -10.0])
def test_sphere_on_axis_backward(radius_mm):
    """surface/sphere.rs:277-366"""
    s = sphere_at(mm(radius_mm), (0.0, 0.0, 0.0))
    back = (0.0, 0.0, -1.0)
    q, n = orc.intersect(s, mm(0.0, 0.0, 5.0), back)
    assert tuple(q) == (0.0, 0.0, 0.0) and tuple(n) == (0.0, 0.0, 1.0)
    q, _ = orc.intersect(s, mm(0.0, 0.0, 15.0), back)
    assert np.all(np.abs(q) <= EPS)
    assert orc.intersect(s, mm(0.0, 0.0, -5.0), back) is None
    assert orc.intersect(s, mm(0.0, 0.0, -15.0), back) is None


def test_sphere_collinear():
    """surface/sphere.rs:367-402: miss, and the tangent ray that needs disc == 0 exactly (Roots::One)"""
    s = sphere_at(mm(1.0), mm(0.0, 0.0, 10.0))
    assert orc.intersect(s, mm(0.0, 1.1, 0.0), Z) is None
    assert orc.intersect(s, mm(0.0, -1.1, 0.0), Z) is None
    q, n = orc.intersect(s, mm(0.0, 1.0, 0.0), Z)
    assert tuple(q) == mm(0.0, 1.0, 11.0) and tuple(n) == (0.0, 1.0, 0.0)
    q, n = orc.intersect(s, mm(0.0, -1.0, -1.0), Z)
    assert q[0] == 0.0 and abs(q[1] + 0.001) <= EPS and abs(q[2] - 0.011) <= 1000 * EPS
    assert np.all(np.abs(n - np.array([0.0, -1.0, 0.0])) <= EPS)


def test_parabola_kats():
    """surface/parabola.rs:140-235"""
    ident = orc.Isometry.identity()
    p = orc.Surface(geom=orc.GEOM_PARABOLA, param=1.0, iso=ident)
    q, n = orc.intersect(p, (-1.0, -1.0, -10.0), Z)
    assert tuple(q) == (-1.0, -1.0, 0.5) and tuple(n) == (-1.0, -1.0, -2.0)
    pn = orc.Surface(geom=orc.GEOM_PARABOLA, param=-1.0, iso=ident)
    d = orc.ray_new((0, 0, -1.0), (0.0, 1.0, 0.75), nm(1000.0), 1.0)[0]["dir"]
    assert orc.intersect(pn, (0.0, 0.0, -1.0), tuple(d)) is not None
    d = orc.ray_new((0, 0, 0), (0.0, 0.5, 2.0), nm(1000.0), 1.0)[0]["dir"]
    assert orc.intersect(p, (0.0, -0.5, -1.0), tuple(d)) is not None
    q, n = orc.intersect(p, (0.0, -1.0, 0.0), (0.0, 1.0, 0.0))
    assert tuple(q) == (0.0, 0.0, 0.0)
    assert tuple(n / np.sqrt((n * n).sum())) == (0.0, 0.0, -1.0)
    assert orc.intersect(p, (0.0, -1.0, -1.0), (0.0, 1.0, 0.0)) is None
    assert orc.intersect(p, (0.0, 0.0, -1.0), (0.0, 0.0, -1.0)) is None


def test_cylinder_kats():
    """surface/cylinder.rs:170-228 (Cylinder::new takes the isometry of the cylinder axis)"""
    s = orc.Surface(geom=orc.GEOM_CYLINDER, param=mm(1.0), iso=orc.Isometry.new_along_z(mm(10.0)))
    for y in (0.0, 1.0):
        q, n = orc.intersect(s, mm(0.0, y, 0.0), Z)
        assert abs(q[0]) <= EPS and abs(q[1] - mm(y)) <= EPS and abs(q[2] - 0.009) <= EPS
        assert np.all(np.abs(n - np.array([0.0, 0.0, -1.0])) <= EPS)
    sb = orc.Surface(geom=orc.GEOM_CYLINDER, param=mm(1.0), iso=orc.Isometry.new_along_z(mm(-10.0)))
    assert orc.intersect(sb, (0, 0, 0), Z) is None
    assert orc.intersect(s, mm(1.1, 0.0, 0.0), Z) is None
    q, n = orc.intersect(s, mm(1.0, 0.0, 0.0), Z)
    assert tuple(q) == mm(1.0, 0.0, 10.0) and tuple(n) == (1.0, 0.0, 0.0)
    q, n = orc.intersect(s, mm(-1.0, 0.0, 0.0), Z)
    assert q[1] == 0.0 and abs(q[0] + 0.001) <= EPS and abs(q[2] - 0.01) <= 1000 * EPS
    assert np.all(np.abs(n - np.array([-1.0, 0.0, 0.0])) <= EPS)


# ---------------------------------------------------------------- coatings / index
def test_fresnel_kats():
    """coatings/fresnel.rs:54-83"""
    nrm = (0.0, 0.0, -1.0)
    assert orc.fresnel(Z, nrm, 1.0, 1.0) == 0.0
    assert orc.fresnel(Z, nrm, 2.0, 2.0) == 0.0
    assert abs(orc.fresnel(Z, nrm, 1.0, 1.5) - 0.04) <= EPS
    # set_direction does not normalise (ray.rs:269-275)
    assert abs(orc.fresnel((0.0, 1.0, 1.0), nrm, 1.0, 1.5) - 0.05023991101223595) <= EPS


def test_constant_r_rejects_zero():
    """coatings/constant_r.rs:22-29 (`!is_normal` rejects 0.0) -- surfaces at trace time via coatings/mod.rs:49"""
    for bad in (0.0, -0.1, 1.1, math.nan, math.inf):
        ray = orc.ray_new((0, 0, 0), Z, nm(1000.0), 1.0)
        with pytest.raises(orc.OracleError):
            orc.ray_refract_on_surface(ray, plane(10.0, orc.COAT_CONSTANT_R, bad), 1.5)
    ray = orc.ray_new((0, 0, 0), Z, nm(1000.0), 1.0)
    orc.ray_refract_on_surface(ray, plane(10.0, orc.COAT_CONSTANT_R, 1.0), 1.5)
    assert ray[0]["e"] == 0.0


HK9L = (6.14555251E-1, 6.56775017E-1, 1.02699346E+0, 1.45987884E-2, 2.87769588E-3, 1.07653051E+2)


def test_index_models():
    """refr_index_sellmeier1.rs:234-251, refr_index_schott.rs:251-268, refr_index_conrady.rs:121-133"""
    lo, hi = nm(500.0), nm(2000.0)
    m = orc.IndexModel.sellmeier1(*HK9L, lo, hi)
    assert m.get_refractive_index(nm(1054.0)) == pytest.approx(1.5068, rel=1e-4)
    m2 = orc.IndexModel.schott(3.26760058E+000, -2.05384566E-002, 3.51507672E-002, 7.70151348E-003,
                               -9.08139817E-004, 7.52649555E-005, lo, hi)
    assert m2.get_refractive_index(nm(1054.0)) == pytest.approx(1.8116, rel=1e-4)
    m3 = orc.IndexModel.conrady(1.0, 1.0, 1.0, lo, hi)
    assert m3.get_refractive_index(nm(1054.0)) == pytest.approx(2.7806, rel=1e-4)
    for mod in (m, m2, m3):
        for bad in (nm(499.0), nm(2001.0)):
            with pytest.raises(orc.OracleError):
                mod.get_refractive_index(bad)
        mod.get_refractive_index(lo)          # Range is end-exclusive: start ok ...
        with pytest.raises(orc.OracleError):  # ... end is not
            mod.get_refractive_index(hi)
    assert orc.IndexModel.const(1.5).get_refractive_index(nm(1.0)) == 1.5
    with pytest.raises(orc.OracleError):      # refractive_index/mod.rs:54-58
        orc.IndexModel.conrady(0.5, 0.0, 0.0, lo, hi).get_refractive_index(nm(1000.0))


def test_roots_branches():
    """roots 0.0.8 branch structure: No / One / Two ascending / linear"""
    assert orc.find_roots_quadratic(1.0, 0.0, 1.0) == []
    assert orc.find_roots_quadratic(1.0, -2.0, 1.0) == [1.0]
    assert orc.find_roots_quadratic(1.0, 0.0, -1.0) == [-1.0, 1.0]
    assert orc.find_roots_quadratic(1.0, -3.0, 2.0) == [1.0, 2.0]
    assert orc.find_roots_quadratic(0.0, 2.0, -4.0) == [2.0]
    assert orc.find_roots_quadratic(0.0, 0.0, 1.0) == []
    assert orc.find_roots_quadratic(0.0, 0.0, 0.0) == [0.0]


# ---------------------------------------------------------------- apertures
def test_apertures():
    """aperture.rs doc example :26-33 and tests :470-640 (circle, rect, gaussian, stack, obstruction)"""
    c = orc.ap_circle(mm(1.0), *mm(1.0, 1.0))
    assert orc.apodization_factor(c, *mm(1.0, 1.0)) == 1.0 and orc.apodization_factor(c, 0.0, 0.0) == 0.0
    c0 = orc.ap_circle(1.0)
    assert orc.apodization_factor(c0, 1.0, 0.0) == 1.0 and orc.apodization_factor(c0, 0.0, 1.0) == 1.0  # `<=` on the rim
    assert orc.apodization_factor(c0, 1.0, 1.0) == 0.0
    co = orc.ap_circle(1.0, obstruction=True)
    assert orc.apodization_factor(co, 0.0, 0.0) == 0.0 and orc.apodization_factor(co, 2.0, 0.0) == 1.0
    r = orc.ap_rect(1.0, 2.0, 1.0, 1.0)
    assert orc.apodization_factor(r, 1.0, 1.0) == 1.0 and orc.apodization_factor(r, 0.5, 0.0) == 1.0  # corner, `<=`
    assert orc.apodization_factor(r, 0.0, 0.0) == 0.0 and orc.apodization_factor(r, 1.6, 1.0) == 0.0
    g = orc.ap_gaussian(1.0, 1.0)
    assert orc.apodization_factor(g, 0.0, 0.0) == 1.0
    assert orc.apodization_factor(g, 1.0, 0.0) == math.exp(-0.5)
    st = orc.ap_stack([orc.ap_circle(1.0), orc.ap_rect(1.0, 1.0)])
    assert orc.apodization_factor(st, 0.4, 0.4) == 1.0 and orc.apodization_factor(st, 0.9, 0.0) == 0.0
    tri = orc.ap_polygon([[(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)]])
    assert orc.apodization_factor(tri, 0.25, 0.25) == 1.0 and orc.apodization_factor(tri, 0.75, 0.75) == 0.0


# ---------------------------------------------------------------- bundles
def _bundle(items):
    rays = np.concatenate([orc.ray_new(p, Z, nm(1053.0), e) for p, e in items])
    return rays


def test_total_energy_centroid_radius():
    """rays.rs:2019-2106"""
    rays = _bundle([((0, 0, 0), 1.0), ((0, 0, 0), 1.0), ((0, 0, 0), 1.0)])
    rays[2]["valid"] = 0
    assert orc.total_energy(rays) == 2.0
    assert orc.total_energy(rays[:0]) == 0.0
    rays = _bundle([(mm(1.0, 2.0, 0.0), 1.0), (mm(2.0, 3.0, 0.0), 1.0), (mm(2.0, 3.0, 0.0), 1.0)])
    rays[2]["valid"] = 0
    assert orc.centroid(rays[:0]) is None
    assert tuple(orc.centroid(rays)) == mm(1.5, 2.5, 0.0)
    assert orc.beam_radius_geo(rays[:0]) is None
    assert orc.beam_radius_geo(rays) == mm(math.sqrt(0.5))
    rays = _bundle([(mm(1.0, 1.0, 0.0), 1.0)])
    assert orc.beam_radius_rms(rays) == 0.0
    rays = _bundle([(mm(1.0, 1.0, 0.0), 1.0), ((0, 0, 0), 1.0), (mm(1.0, 15.0, 0.0), 1.0)])
    rays[2]["valid"] = 0
    assert orc.beam_radius_rms(rays) == mm(math.sqrt(2.0) / 2.0)


def test_energy_weighted_centroid():
    """rays.rs:2696-2734"""
    rays = _bundle([(mm(-1.0, 0.0, 0.0), 1.0), (mm(1.0, 0.0, 0.0), 1.0)])
    assert tuple(orc.energy_weighted_centroid(rays)) == (0.0, 0.0, 0.0)
    rays = _bundle([(mm(-1.0, 0.0, 0.0), 1.0), (mm(1.0, 0.0, 0.0), 0.5)])
    c = orc.energy_weighted_centroid(rays)
    assert c[0] / 1e-3 == pytest.approx(-1.0 / 3.0, rel=EPS) and c[1] == 0.0 and c[2] == 0.0
    rays[1]["valid"] = 0
    assert orc.energy_weighted_centroid(rays)[0] / 1e-3 == pytest.approx(-1.0, rel=EPS)
    assert orc.energy_weighted_centroid(rays[:0]) is None
    # energy-weighted rms radius is unpinned by the reference (mutants.out/missed.txt rays.rs:694,698):
    rays = _bundle([(mm(-1.0, 0.0, 0.0), 1.0), (mm(1.0, 0.0, 0.0), 1.0)])
    assert orc.energy_weighted_beam_radius_rms(rays) == pytest.approx(1e-3, rel=1e-15)


def test_bundle_refract_missed_and_energy():
    """rays.rs:2246-2286"""
    default_surface = plane(0.0)
    vac = orc.IndexModel.const(1.0)
    rays = _bundle([(mm(0.0, 0.0, 1.0), 1.0)])
    children, hits, st = orc.rays_refract_on_surface(rays, default_surface, vac, True, orc.MISS_STOP)
    assert len(children) == 0 and st.rays_missed == 1
    rays = _bundle([(mm(0.0, 0.0, -1.0), 1.0)])
    s = plane(0.0, orc.COAT_CONSTANT_R, 0.2)
    children, hits, st = orc.rays_refract_on_surface(rays, s, vac, True, orc.MISS_STOP)
    assert orc.total_energy(rays) == 0.8 and orc.total_energy(children) == 0.2
    assert st.n_interactions == 1 and st.n_children == 1 and st.n_hits == 1


def test_threshold_and_apodize():
    """rays.rs:2307-2350"""
    empty = _bundle([((0, 0, 0), 1.0)])[:0]
    for bad in (math.nan, math.inf):
        with pytest.raises(orc.OracleError):
            orc.rays_threshold(empty, bad)
    orc.rays_threshold(empty, -0.1)
    rays = _bundle([((0, 0, 0), 1.0), ((0, 0, 0), 0.1)])
    orc.rays_threshold(rays, -0.1)
    assert orc.nr_of_rays(rays) == 2
    orc.rays_threshold(rays, 0.1)
    assert orc.nr_of_rays(rays) == 2
    orc.rays_threshold(rays, 0.5)
    assert orc.nr_of_rays(rays) == 1
    orc.rays_threshold(rays, 1.1)
    assert orc.nr_of_rays(rays) == 0
    rays = _bundle([((0, 0, 0), 1.0), (mm(1.0, 1.0, 0.0), 1.0)])
    inv = orc.rays_apodize(rays, orc.ap_circle(mm(0.5)), orc.Isometry.identity())
    assert inv and orc.total_energy(rays) == 1.0


def test_limits():
    """rays.rs:1386-1404: bounces `>` max, refractions `>=` max"""
    rays = _bundle([((0, 0, 0), 1.0)] * 3)
    rays["bounces"] = [0, 1, 2]
    orc.rays_filter_bounces(rays, 1)
    assert list(rays["valid"]) == [1, 1, 0]
    rays = _bundle([((0, 0, 0), 1.0)] * 3)
    rays["refr"] = [0, 1, 2]
    orc.rays_filter_refractions(rays, 1)
    assert list(rays["valid"]) == [1, 0, 0]


# ---------------------------------------------------------------- binning
def test_binning_kat():
    """surface/hit_map/mod.rs:1200-1240: 51x51, centre bin = 2601.0 (both point sets, merged)"""
    pos1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    h = orc.hits_from_xye([p[0] for p in pos1], [p[1] for p in pos1], [1.0] * 5)
    fmap, lrtb = orc.fluence_binning(h, 51, 51)
    assert fmap.shape == (51, 51)
    assert fmap[25, 25] == pytest.approx(2601.0, rel=EPS)
    assert lrtb == (-0.5, 0.5, 0.5, -0.5)
    r = math.sqrt(0.5)
    pos2 = [(r, 0.0), (0.0, r), (-r, 0.0), (0.0, -r), (0.0, 0.0)]
    h2 = orc.hits_from_xye([p[0] for p in pos1 + pos2], [p[1] for p in pos1 + pos2], [1.0] * 10)
    fmap2, _ = orc.fluence_binning(h2, 51, 51)
    # the reference asserts 2601.0 again although two points now sit on the centre: it only holds
    # because the merged map spans +-sqrt(0.5): bin_area doubles, (2 J)/(2/2601) = 2601
    assert fmap2[25, 25] == pytest.approx(2601.0, rel=4 * EPS)
    with pytest.raises(orc.OracleError):
        orc.fluence_binning(h[:0], 51, 51)
    assert orc.fluence_peak(fmap) == fmap.max()
    # corner hits drop their outside taps: sum(map)*bin_area == 5 J here because every hit sits on a node
    assert orc.fluence_total_energy(fmap, lrtb) == pytest.approx(5.0, rel=1e-14)


# ---------------------------------------------------------------- sources
def test_grid_kats():
    """position_distributions/grid.rs:119-143"""
    p = orc.grid(2, 2, mm(1.0), mm(1.0))
    assert [tuple(x) for x in p] == [mm(-0.5, -0.5, 0.0), mm(-0.5, 0.5, 0.0), mm(0.5, -0.5, 0.0), mm(0.5, 0.5, 0.0)]
    assert [tuple(x) for x in orc.grid(1, 1, mm(1.0), mm(1.0))] == [(0.0, 0.0, 0.0)]
    assert [tuple(x) for x in orc.grid(1, 2, mm(1.0), mm(1.0))] == [mm(0.0, -0.5, 0.0), mm(0.0, 0.5, 0.0)]


def test_hexapolar_counts():
    """position_distributions/hexapolar.rs:74-99"""
    assert len(orc.hexapolar(1, 0.0)) == 1 and len(orc.hexapolar(0, mm(1.0))) == 1
    for rings, n in ((1, 7), (2, 19), (6, 127), (7, 169), (255, 195841)):
        assert len(orc.hexapolar(rings, mm(1.0))) == n


def test_fibonacci_and_spectrum_properties():
    """fibonacci.rs:53-65,113-126; spectral_distribution/gaussian.rs:78-96 (weights Kahan-normalised to 1)"""
    p = orc.fibonacci_rectangle(1000, 2.0, 4.0)
    assert np.all(np.abs(p[:, 0]) <= 1.0) and np.all(np.abs(p[:, 1]) <= 2.0) and np.all(p[:, 2] == 0.0)
    assert tuple(p[0]) == (-1.0, -2.0, 0.0)
    p = orc.fibonacci_ellipse(1000, 1.0, 2.0)
    assert np.all((p[:, 0] / 1.0) ** 2 + (p[:, 1] / 2.0) ** 2 <= 1.0 + 1e-12)
    wvl, w = orc.spectrum_gaussian(nm(950.0), nm(1150.0), 64, nm(1053.0), nm(50.0))
    assert wvl[0] == nm(950.0) and abs(wvl[-1] - nm(1150.0)) < 1e-18 and abs(w.sum() - 1.0) < 1e-14
    e = orc.energy_general_gaussian(orc.fibonacci_ellipse(500, 0.01, 0.01), 1.0, sigma=(0.004, 0.004), power=5.0)
    assert abs(e.sum() - 1.0) < 1e-13 and np.all(e >= 0.0)
    rays = orc.rays_collimated_with_spectrum(orc.grid(2, 2, 1.0, 1.0), orc.energy_uniform(4, 1.0), wvl[:3], w[:3])
    assert len(rays) == 12 and list(rays["wvl"][:3]) == list(wvl[:3])  # position-major, wavelength-minor
    assert rays["e"][1] == 0.25 * w[1]


# ------------------------------------------------------------------ kernel density estimator
def _hits_xy(points, e=None):
    p = np.asarray(points, dtype=float)
    return orc.hits_from_xye(p[:, 0], p[:, 1], np.zeros(len(p)) if e is None else e)


def test_kde_point_distances_std_dev():
    """kde/mod.rs:178-209"""
    d, sd = orc.kde_point_distances(_hits_xy([(0.0, 0.0), (1e-3, 0.0)]))
    assert list(d) == [1e-3] and sd == 0.0
    d, sd = orc.kde_point_distances(_hits_xy([(0.0, 0.0), (1.0, 0.0), (-1.0, 0.0)]))
    assert list(d) == [1.0, 1.0, 2.0] and sd == math.sqrt(2.0 / 9.0)
    _, sd = orc.kde_point_distances(_hits_xy([(0.0, 0.0)]))
    assert math.isnan(sd)


def test_kde_distances_iqr():
    """kde/mod.rs:210-240 (incl. the Wikipedia example)"""
    assert math.isnan(orc.kde_distances_iqr([]))
    assert orc.kde_distances_iqr([1.0]) == 1.0
    assert orc.kde_distances_iqr([0.0, 1.0]) == 1.0
    assert orc.kde_distances_iqr([0.0, 1.0, 2.0]) == 2.0
    assert orc.kde_distances_iqr([0.0, 1.0, 2.0, 3.0]) == 2.5
    assert orc.kde_distances_iqr([25.0, 28.0, 4.0, 28.0, 19.0, 3.0, 9.0, 17.0, 29.0, 29.0]) == 28.0


def test_kde_bandwidth_estimate():
    """kde/mod.rs:241-262"""
    assert math.isnan(orc.kde_bandwidth_estimate(_hits_xy([(0.0, 0.0)])))
    assert orc.kde_bandwidth_estimate(_hits_xy([(0.0, 0.0), (1e-3, 0.0)])) == 0.5e-3
    bw = orc.kde_bandwidth_estimate(_hits_xy([(0.0, 0.0), (1e-3, 0.0), (-1e-3, 0.0)]))
    assert abs(bw - 0.00034057440111656337) < 1e-15  # assert_abs_diff_eq! default epsilon


def test_kde_gaussian_norm():
    """kde/gaussian.rs:40-58: a 1 J Gaussian of sigma 5 mm integrates to 1 J over a 200 mm square (2 %)"""
    h = _hits_xy([(0.0, 0.0)], e=[1.0])
    xs = np.linspace(-0.1, 0.1, 120)
    total = sum(orc.kde_value(h, 5e-3, x, y) for x in xs for y in xs) * (0.2 * 0.2 / (120 * 120))
    assert abs(total - 1.0) < 0.02


def test_kde_fluence_map_conserves_energy():
    """calc_fluence_with_kde (rays_hit_map.rs:575-613): the map over the 3-bandwidth margin holds (nearly) all the energy"""
    xyz = orc.hexapolar(3, 5e-3)  # the hit set of the reference's own `kde` criterion bench (benches/fluence_estimator.rs:12-24)
    assert len(xyz) == 37
    h = orc.hits_from_xye(xyz[:, 0], xyz[:, 1], np.full(37, 1.0 / 37))
    fmap, (l, r, t, b), bw = orc.fluence_kde(h, 30, 30)
    assert bw > 0 and l == xyz[:, 0].min() - 3 * bw and t == xyz[:, 1].max() + 3 * bw
    energy = fmap.sum() * ((r - l) / 30) * ((t - b) / 30)
    assert abs(energy - 1.0) < 0.05


# ------------------------------------------------------------------ reflective grating
def _grating(extra_tilt, wvl=1000e-9, dens=1740e3, order=-1):
    littrow = math.asin(order * wvl * dens / 2.)  # with_rot_from_littrow (nodes/reflective_grating.rs:107-128)
    iso = orc.Isometry.new((0.0, 0.0, 0.0), (0.0, littrow + extra_tilt, 0.0))
    surf = orc.Surface(geom=orc.GEOM_PLANE, coating=orc.COAT_IDEAL_AR, param=0.0, coating_r=0.0, iso=iso)
    gvec = [2 * math.pi * dens * c for c in iso.transform_vector((1.0, 0.0, 0.0))]
    return surf, gvec


def test_grating_littrow_and_one_degree_off():
    """nodes/reflective_grating.rs:386-443: 1740 lines/mm, order -1, 1000 nm ray along z from the origin"""
    for extra in (0.0, math.radians(1.0)):
        surf, gvec = _grating(extra)
        rays = orc.rays_from_arrays(np.zeros((1, 3)), np.array([[0.0, 0.0, 1.0]]), np.array([1.0]), np.array([1000e-9]))
        out, st = orc.rays_diffract_on_periodic_surface(rays, surf, orc.IndexModel.const(1.0), gvec, -1, False)
        assert len(out) == 1 and out["valid"][0] == 1 and np.array_equal(out["pos"][0], [0.0, 0.0, 0.0])
        assert out["bounces"][0] == 0 and out["e"][0] == 1.0 and rays["e"][0] == 0.0  # bounce taken back, energy moved to the clone
        if extra == 0.0:
            expect = np.array([0.0, 0.0, -1.0])
        else:
            input_angle = math.asin(-1000e-9 * 1740000. / 2.) + extra
            diffraction_angle = math.asin(-1740000. * 1000e-9 - math.sin(input_angle))
            expect = np.array([math.sin(-input_angle + diffraction_angle), 0.0, -math.cos(-input_angle + diffraction_angle)])
        assert np.all(np.abs(out["dir"][0] - expect) <= 1e-15 * np.maximum(1.0, np.abs(expect)))  # assert_relative_eq!(.., epsilon = 1e-15)


def test_grating_evanescent_order_and_miss():
    """ray.rs:541-551: an order whose in-plane k exceeds k0 invalidates the ray (no clone); a miss leaves it untouched"""
    surf, gvec = _grating(0.0)
    rays = orc.rays_from_arrays(np.zeros((2, 3)), np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]]), np.ones(2), np.full(2, 1000e-9))
    rays["pos"][1] = (0.0, 0.0, -1e-3)  # flies away from the grating plane
    before = rays.copy()
    out, st = orc.rays_diffract_on_periodic_surface(rays, surf, orc.IndexModel.const(1.0), gvec, -3, False)
    assert len(out) == 0 and st.rays_missed == 1
    assert rays["valid"][0] == 0 and rays["valid"][1] == 1
    assert np.array_equal(rays["pos"][1], before["pos"][1]) and rays["e"][1] == 1.0


# ---- position history ------------------------------------------------------------------------------
def test_position_history_refract_kat():
    """ray.rs:1236-1272: after refract_on_surface at a plane 10 mm ahead `ray.pos_hist == [(0,0,0)]` and the reflected
    clone carries the same history (it is cloned after the push, ray.rs:616-625)"""
    ray = np.ascontiguousarray(orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1053e-9, 1.0))
    plane = orc.Surface(geom=orc.GEOM_PLANE, coating=orc.COAT_CONSTANT_R, coating_r=0.2, param=0.0,
                        iso=orc.Isometry.new((0.0, 0.0, 10e-3), (0.0, 0.0, 0.0)))
    H = orc.History(1, cap=4)
    H.attach(ray)
    try:
        ch, _, st = orc.rays_refract_on_surface(ray, plane, orc.IndexModel.const(1.5))
    finally:
        orc.History.detach()
    assert st.n_children == 1 and H.len[0] == 1
    assert np.array_equal(H.of(0), [[0.0, 0.0, 0.0]])
    assert np.array_equal(H.of(0, ray), [[0.0, 0.0, 0.0], [0.0, 0.0, 10e-3]])  # position_history_with_current
    assert np.array_equal(ch["pos"][0], [0.0, 0.0, 10e-3])


def test_position_history_push_rules():
    """a miss pushes nothing (ray.rs:680-686), total reflection pushes (the surface was reached, :616), apodize pushes the
    position of exactly the rays it invalidates (rays.rs:566), invalid rays never push"""
    rays = np.ascontiguousarray(orc.rays_from_arrays(
        pos=[(0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 5e-3, 0)], direction=[(0, 0, 1), (0, 0, -1), (0, 0.8, 0.6), (0, 0, 1)],
        e=[1.0] * 4, wvl=[1e-6] * 4))
    rays["n"][2] = 1.5  # steep ray inside glass -> total internal reflection at the plane
    plane = orc.Surface(geom=orc.GEOM_PLANE, coating=orc.COAT_IDEAL_AR, coating_r=0.0, param=0.0,
                        iso=orc.Isometry.new((0.0, 0.0, 10e-3), (0.0, 0.0, 0.0)))
    H = orc.History(4, cap=4)
    H.attach(rays)
    try:
        orc.rays_refract_on_surface(rays, plane, orc.IndexModel.const(1.0), strategy=orc.MISS_IGNORE)
        assert list(H.len) == [1, 0, 1, 1] and rays["bounces"][2] == 1
        orc.rays_apodize(rays, orc.ap_circle(2e-3), orc.Isometry.identity())
        assert list(H.len) == [1, 0, 2, 2] and list(rays["valid"]) == [1, 1, 0, 0]
        assert np.array_equal(H.of(3), [[0.0, 5e-3, 0.0], [0.0, 5e-3, 10e-3]])
        orc.rays_refract_on_surface(rays, plane, orc.IndexModel.const(1.0), strategy=orc.MISS_IGNORE)
        assert list(H.len) == [2, 0, 2, 2]
    finally:
        orc.History.detach()


# ---- spectrum.rs / wavefront -------------------------------------------------------------------------
UM = 1e-6


def _prep():  # spectrum.rs:657-659: Spectrum::new(1.0..4.0 um, 0.5 um)
    return orc.spectrum_new(1.0 * UM, 4.0 * UM, 0.5 * UM)


def test_spectrum_new_and_single_peak_kats():
    """spectrum.rs:661-675 (new), :800-833 (set_single_peak*), :875-886 (total_energy*)"""
    lam, data = _prep()
    assert np.allclose(lam, [1.0, 1.5, 2.0, 2.5, 3.0, 3.5], rtol=0, atol=1e-15) and not data.any()
    for bad in ((1.0 * UM, 4.0 * UM, -0.5 * UM), (4.0 * UM, 1.0 * UM, 0.5 * UM), (-1.0 * UM, 4.0 * UM, 0.5 * UM)):
        with pytest.raises(orc.OracleError):
            orc.spectrum_new(*bad)
    orc.spectrum_add_single_peak(lam, data, 2.0 * UM, 1.0)
    assert data[2] == 2.0 and orc.spectrum_total_energy(lam, data) == 1.0
    orc.spectrum_add_single_peak(lam, data, 2.0 * UM, 1.0)
    assert data[2] == 4.0  # additive
    lam, data = _prep()
    orc.spectrum_add_single_peak(lam, data, 2.25 * UM, 1.0)
    assert data[2] == 1.0 and data[3] == 1.0  # interpolated
    lam, data = _prep()
    orc.spectrum_add_single_peak(lam, data, 2.0 * UM, 1.0)
    orc.spectrum_add_single_peak(lam, data, 2.25 * UM, 1.0)
    assert data[2] == 3.0 and data[3] == 1.0
    lam, data = _prep()
    orc.spectrum_add_single_peak(lam, data, 1.0 * UM, 1.0)
    assert data[0] == 2.0  # lower bound
    before = data.copy()
    orc.spectrum_add_single_peak(lam, data, 0.5 * UM, 1.0)  # outside: warning, unmodified
    orc.spectrum_add_single_peak(lam, data, 4.0 * UM, 1.0)
    assert np.array_equal(data, before)
    with pytest.raises(orc.OracleError):
        orc.spectrum_add_single_peak(lam, data, 1.5 * UM, -1.0)
    lam, data = orc.spectrum_new(1.0 * UM, 4.0 * UM, 1.0 * UM)
    orc.spectrum_add_single_peak(lam, data, 1.5 * UM, 1.0)
    assert orc.spectrum_total_energy(lam, data) == 1.0


def test_spectrum_from_laser_lines_kats():
    """spectrum.rs:726-756"""
    lam, data = orc.spectrum_from_laser_lines([1.0 * UM], [1.0], 1e-9)
    assert orc.spectrum_total_energy(lam, data) == 1.0
    assert abs(lam[0] - 1.0) < 1e-12 and abs(lam[1] - 1.001) < 1e-12 and abs(data[0] - 1000.0) < 1e-9 and data[1] == 0.0
    lam, data = orc.spectrum_from_laser_lines([1.0 * UM, 1.010 * UM], [1.0, 0.5], 1e-9)
    assert abs(orc.spectrum_total_energy(lam, data) - 1.5) < 1e-9
    assert abs(data[0] - 1000.0) < 1e-9 and data[1] == 0.0 and data[2] == 0.0 and abs(lam[10] - 1.010) < 1e-12
    assert abs(data[10] - 500.0) < 1e-9 and data[11] == 0.0 and abs(lam[11] - 1.011) < 1e-12


def test_spectrum_get_value_kat():
    """spectrum.rs:887-909"""
    lam, data = [1.0, 2.0, 3.0], [1.0, 2.0, 4.0]
    want = {0.9: None, 1.0: 1.0, 1.2: 1.2, 2.0: 2.0, 2.75: 3.5, 3.0: 4.0, 3.1: None}
    for w, v in want.items():
        got = orc.spectrum_get_value(lam, data, w * UM)
        assert (got is None) if v is None else abs(got - v) < 1e-12, (w, got, v)
    assert orc.spectrum_get_value([1.0], [1.0], 0.9 * UM) is None and orc.spectrum_get_value([1.0], [1.0], 1.0 * UM) == 1.0


def test_wavefront_statistics_kat():
    """nodes/wavefront.rs:353-368: two rays at the origin, one an extra wavelength along -> ptv 1.0, rms 0.5"""
    wvl = 1000e-9
    rays = np.ascontiguousarray(orc.rays_from_arrays(pos=[(0, 0, 0), (0, 0, wvl)], direction=[(0, 0, 1)] * 2, e=[1.0] * 2, wvl=[wvl] * 2))
    rays["opl"][1] = wvl  # Ray::propagate(wvl) in vacuum
    xyw = orc.rays_wavefront_error(rays, wvl, orc.Isometry.identity())
    assert xyw.shape == (2, 3) and xyw[0, 2] == 0.0 and abs(xyw[1, 2] + 1.0) < 1e-12
    ptv, rms = orc.wavefront_statistics(xyw)
    assert abs(ptv - 1.0) < 1e-12 and abs(rms - 0.5) < 1e-12
    with pytest.raises(orc.OracleError):
        orc.wavefront_statistics(np.zeros((0, 3)))


# ---------------------------------------------------------------------------------------------------------------
# node positioning helpers
# ---------------------------------------------------------------------------------------------------------------
def test_define_up_direction_kats():
    """ray.rs:1653-1755 define_up_direction_test (assert_relative_eq!; the last two with epsilon = 10 eps)"""
    unit = lambda v: np.asarray(v, dtype=float) / np.linalg.norm(v)  # noqa: E731  (Ray::new normalises the direction)
    x, y = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])
    cases = [((1, 0, 0), y, EPS), ((0, 1, 0), x, EPS), ((0, 0, 1), y, EPS), ((0, 1, 1), unit((0, 1, -1)), EPS), ((1, 0, 1), y, EPS),
             ((1, 1, 0), unit((-1, 1, 0)), EPS), ((-1, 0, 3), y, EPS), ((1, 1, 1), unit((-1, 2, -1)), 10 * EPS),
             ((-1, -1, -1), unit((-1, 2, -1)), 10 * EPS)]
    for d, want, tol in cases:
        got = orc.define_up_direction(unit(d))
        assert np.all(np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want))), (d, got, want)


def test_calc_new_up_direction_kat():
    """ray.rs:1757-1790 calc_new_up_direction_test: z -> y turns up = y into -z, y -> z turns it back, z -> x leaves y"""
    up = np.array([0.0, 1.0, 0.0])
    up = orc.calc_new_up_direction((0, 0, 1), (0, 1, 0), up)
    assert np.allclose(up, (0, 0, -1), rtol=0, atol=EPS)
    up = orc.calc_new_up_direction((0, 1, 0), (0, 0, 1), up)
    assert np.allclose(up, (0, 1, 0), rtol=0, atol=EPS)
    up = orc.calc_new_up_direction((0, 0, 1), (1, 0, 0), up)
    assert np.allclose(up, (0, 1, 0), rtol=0, atol=EPS)
    # unchanged direction (below 1000 eps): untouched, bit for bit
    up2 = orc.calc_new_up_direction((0, 0, 1), (0, 1e-14, 1), np.array([0.3, 0.4, 0.5]))
    assert np.array_equal(up2, (0.3, 0.4, 0.5))


def test_calc_node_positions_on_a_straight_chain():
    """nodes/node_group/analysis_raytrace.rs:92-212, optic_graph.rs:878-926: on an unfolded, centred chain a node sits `distance`
    behind the LAST surface vertex of its predecessor (the axis ray leaves a lens at its rear vertex), unrotated"""
    import scenes
    MM = scenes.MM
    nodes = [scenes.node("source"),
             scenes.node("lens", distance=50 * MM, p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=scenes.NK9L_1054),
             scenes.node("wedge", distance=30 * MM, p=(10 * MM, 0.0), index=scenes.HK9L),
             scenes.node("spot_diagram", distance=100 * MM),
             scenes.node("fluence_detector", distance=10 * MM)]
    osc = scenes.to_oracle(nodes)
    isos = osc.calc_node_positions(1000e-9)
    want_z = [0.0, 50 * MM, (50 + 2.79 + 30) * MM, (50 + 2.79 + 30 + 10 + 100) * MM, (50 + 2.79 + 30 + 10 + 100 + 10) * MM]
    for iso, z in zip(isos, want_z):
        rot, t, _, _ = iso.matrices()
        assert np.allclose(rot.reshape(3, 3), np.eye(3), rtol=0, atol=4 * EPS)
        assert np.allclose(t, (0.0, 0.0, z), rtol=4 * EPS, atol=1e-18)
    # and the placed scene traces like the one with explicit isometries
    explicit = [dict(n) for n in nodes]
    for n, z in zip(explicit, want_z):
        n.pop("distance", None)
        n["t"] = (0.0, 0.0, z)
    src = scenes.source_to_oracle(scenes.c1_source(20))
    a = osc.raytrace(src.copy())
    b = scenes.to_oracle(explicit).raytrace(src.copy())
    for f in ("pos", "dir", "e", "opl"):
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=1e-15), f
    assert np.array_equal(a["valid"], b["valid"])


# ---------------------------------------------------------------------------------------------------------------
# bundle-level tests of rays.rs that the per-ray transcriptions above do not cover
# ---------------------------------------------------------------------------------------------------------------
def test_bundle_refract_paraxial():
    """rays.rs:2108-2146: focal length 0 / NaN / +-inf are errors; an on-axis ray is untouched, a ray 1 mm off axis turns to
    normalize(0, -1, 100)"""
    iso = orc.Isometry.identity()
    empty = np.zeros(0, dtype=orc.RAY_DTYPE)
    for bad in (0.0, math.nan, math.inf, -math.inf):
        with pytest.raises(orc.OracleError):
            orc.rays_refract_paraxial(empty, mm(bad) if math.isfinite(bad) else bad, iso)
    orc.rays_refract_paraxial(empty, mm(100.0), iso)
    rays = np.concatenate([orc.ray_new(mm(0.0, 0.0, 0.0), Z, nm(1053.0), 1.0), orc.ray_new(mm(0.0, 1.0, 0.0), Z, nm(1053.0), 1.0)])
    before = rays.copy()
    orc.rays_refract_paraxial(rays, mm(100.0), iso)
    assert tuple(rays[0]["pos"]) == tuple(before[0]["pos"]) and tuple(rays[0]["dir"]) == tuple(before[0]["dir"])
    assert tuple(rays[1]["pos"]) == tuple(before[1]["pos"])
    want = np.array([0.0, -1.0, 100.0]) / math.sqrt(10001.0)
    assert np.all(np.abs(rays[1]["dir"] - want) <= EPS)


def test_bundle_refract_on_surface_same_index():
    """rays.rs:2214-2245: n2 = None keeps every ray's own index (1.5 and 2.0): the directions do not change"""
    rays = np.concatenate([orc.ray_new(mm(0.0, 0.0, -10.0), (0.0, 1.0, 1.0), nm(1053.0), 1.0),
                           orc.ray_new(mm(0.0, 1.0, -10.0), (0.0, 1.0, 1.0), nm(1053.0), 1.0)])
    rays["n"] = (1.5, 2.0)  # Ray::set_refractive_index
    orc.rays_refract_on_surface(rays, plane(), None, True, orc.MISS_STOP)
    want = np.array([0.0, 1.0, 1.0]) / math.sqrt(2.0)
    for r in rays:
        assert np.all(np.abs(r["dir"] - want) <= EPS)
    assert tuple(rays["n"]) == (1.5, 2.0)


def test_bundle_filter_energy():
    """rays.rs:2288-2306: Constant(0.5) on an empty bundle is fine, -0.1 and 1.1 are errors; a bundle of one ray ends up like the
    ray filtered on its own"""
    empty = np.zeros(0, dtype=orc.RAY_DTYPE)
    orc.rays_filter_energy(empty, 0.5)
    for bad in (-0.1, 1.1):
        with pytest.raises(orc.OracleError):
            orc.rays_filter_energy(empty, bad)
    ray = orc.ray_new(mm(0.0, 1.0, 0.0), Z, nm(1054.0), 1.0)
    rays = ray.copy()
    orc.rays_filter_energy(ray, 0.3)
    orc.rays_filter_energy(rays, 0.3)
    assert len(rays) == 1
    for f in ("pos", "dir"):
        assert tuple(rays[0][f]) == tuple(ray[0][f])
    assert rays[0]["wvl"] == ray[0]["wvl"] and rays[0]["e"] == ray[0]["e"] == 0.3


def test_bundle_split():
    """rays.rs:2429-2459: ratios outside [0, 1] are errors; 1 J + 2 J split with ratio 0.2 leaves 0.6 J and splits off 2.4 J"""
    rays = np.concatenate([orc.ray_new((0.0, 0.0, 0.0), Z, nm(1053.0), 1.0), orc.ray_new((0.0, 0.0, 0.0), Z, nm(1050.0), 2.0)])
    for bad in (1.1, -0.1):
        with pytest.raises(orc.OracleError):
            orc.rays_split(rays.copy(), bad)
    split = orc.rays_split(rays, 0.2)
    assert abs(orc.total_energy(rays) - 0.6) <= EPS
    assert abs(orc.total_energy(split) - 2.4) <= 10 * EPS
    # an invalid ray is copied unchanged and counts for nothing (rays.rs:1258-1262, total_energy over valid rays)
    rays = np.concatenate([orc.ray_new((0.0, 0.0, 0.0), Z, nm(1053.0), 1.0), orc.ray_new((0.0, 0.0, 0.0), Z, nm(1050.0), 2.0),
                           orc.ray_new((0.0, 0.0, 0.0), Z, nm(1053.0), 5.0)])
    rays["valid"][2] = 0
    split = orc.rays_split(rays, 0.2)
    assert abs(orc.total_energy(split) - 2.4) <= 10 * EPS and split[2]["e"] == 5.0 and split[2]["valid"] == 0


def test_bundle_to_spectrum_range_follows_the_valid_rays():
    """rays.rs:2352-2384 wavelength_range: the minimum / maximum wavelength of the VALID rays (1050..1053 nm with an invalid
    1000 nm ray in the bundle) -- here through Rays::to_spectrum, whose slot range starts from it (rays.rs:1199-1215)"""
    rays = np.concatenate([orc.ray_new((0.0, 0.0, 0.0), Z, nm(w), 1.0) for w in (1053.0, 1053.0, 1050.0, 1051.0, 1000.0)])
    rays["valid"][4] = 0
    lam, val = orc.rays_to_spectrum(rays, nm(0.5))
    assert lam[0] * 1e-6 >= nm(1050.0) - nm(2.0) and lam[-1] * 1e-6 <= nm(1053.0) + nm(2.0)
    assert lam[0] * 1e-6 > nm(1001.0)  # the invalid 1000 nm ray does not widen the range


def test_parabola_off_axis_isometry_kats():
    """nodes/parabolic_mirror.rs:901-951: the anchor isometry of a 1 m, 45 deg, collimating off-axis parabola with off-axis direction
    (0, 1) (assert_relative_eq!, epsilon = 3 eps), and the parent focal length 0.5 m of a 90 deg one"""
    from oracle import scene as OS
    iso, _ = OS.parabola_off_axis_isometry(1.0, math.radians(45.0), 0.0, 1.0, 1.0)
    rot, t, _, _ = iso.matrices()
    want_rot = np.array([[-0.0, 1.0, -0.0], [-0.7071067811865475, -0.0, -0.7071067811865476], [-0.7071067811865476, 0.0, 0.7071067811865475]])
    want_t = np.array([-0.0, -0.6035533905932736, -0.3964466094067264])
    assert np.all(np.abs(rot.reshape(3, 3) - want_rot) <= 3 * EPS) and np.all(np.abs(t - want_t) <= 3 * EPS)
    assert OS.parabola_off_axis_isometry(1.0, math.radians(90.0), 0.0, 1.0, 1.0)[1] == pytest.approx(0.5, rel=2 * EPS)
    ident, pf = OS.parabola_off_axis_isometry(2.0)  # on axis: no anchor offset, the parabola of the node itself
    assert np.array_equal(ident.matrices()[0], np.eye(3).ravel()) and np.array_equal(ident.matrices()[1], np.zeros(3)) and pf == 2.0


def test_fluence_data_peak_and_total_energy_kats():
    """nodes/fluence_detector/fluence_data.rs:278-300: peak of [[1, 2], [3, 4]] J/cm^2 is 4 J/cm^2; the total energy of
    [[NaN, 8], [8, 4]] J/m^2 over the unit square is 5 J (NaN bins are skipped, every bin covers a quarter)"""
    j_per_cm2 = 1.0e4
    assert orc.fluence_peak(np.array([[1.0, 2.0], [3.0, 4.0]]) * j_per_cm2) == 4.0 * j_per_cm2
    assert orc.fluence_total_energy(np.array([[math.nan, 8.0], [8.0, 4.0]]), (0.0, 1.0, 1.0, 0.0)) == 5.0


def test_hit_map_bounding_box_kat():
    """surface/hit_map/rays_hit_map.rs:1001-1060 calc_2d_bounding_box with margin 0 (what the Binning estimator uses,
    :337-342): (left, right, top, bottom) grows with every hit; no hits is an error"""
    with pytest.raises(orc.OracleError):
        orc.hits_bounding_box(orc.hits_from_xye([], [], []))
    pts = [(0.0, 0.0), (-1.0, 1.0), (-1.0, -1.0), (-1.0, 2.0), (1.0, 2.0)]
    want = [(0.0, 0.0, 0.0, 0.0), (-1.0, 0.0, 1.0, 0.0), (-1.0, 0.0, 1.0, -1.0), (-1.0, 0.0, 2.0, -1.0), (-1.0, 1.0, 2.0, -1.0)]
    for k in range(len(pts)):
        x, y = zip(*pts[: k + 1])
        assert orc.hits_bounding_box(orc.hits_from_xye(x, y, [1.0] * (k + 1))) == want[k]


def test_isometry_kats():
    """utils/geom_transformation.rs:680-800: new_along_z (translation (0, 0, 0.01), no rotation), its inverse (0, 0, -0.01),
    append_z (10 mm then 20 mm = 0.03 m), append_with_rot (T(0, 0, 10 mm) R_y(90 deg) then T(0, 0, 20 mm) maps the origin to
    (0.02, 0, 0.01)), rotation (Euler angles 10 / 20 / 30 deg come back out of the rotation matrix)"""
    i = orc.Isometry.new_along_z(mm(10.0))
    rot, t, rot_inv, t_inv = i.matrices()
    assert np.array_equal(rot, np.eye(3).ravel()) and tuple(t) == (0.0, 0.0, 0.01)
    assert np.array_equal(rot_inv, np.eye(3).ravel()) and tuple(t_inv) == (0.0, 0.0, -0.01)
    both = i.append(orc.Isometry.new_along_z(mm(20.0)))
    assert tuple(both.matrices()[1]) == (0.0, 0.0, 0.03)
    r = orc.Isometry.new(mm(0.0, 0.0, 10.0), (0.0, math.radians(90.0), 0.0)).append(orc.Isometry.new_along_z(mm(20.0)))
    p = r.transform_point((0.0, 0.0, 0.0))
    assert np.all(np.abs(p - (0.02, 0.0, 0.01)) <= EPS)
    m = orc.Isometry.new((0.0, 0.0, 0.0), tuple(math.radians(a) for a in (10.0, 20.0, 30.0))).matrices()[0].reshape(3, 3)
    roll, pitch, yaw = math.atan2(m[2, 1], m[2, 2]), -math.asin(m[2, 0]), math.atan2(m[1, 0], m[0, 0])  # nalgebra euler_angles()
    assert np.all(np.abs(np.array([roll, pitch, yaw]) - np.radians([10.0, 20.0, 30.0])) <= EPS)
    ident = orc.Isometry.identity().matrices()
    assert np.array_equal(ident[0], np.eye(3).ravel()) and not ident[1].any() and np.array_equal(ident[2], np.eye(3).ravel())


def _one_node_chain(kind, ray, **node_kw):
    """the ray through a source at the origin (identity, no aperture) and ONE node: what the reference's node tests do with
    `AnalysisRayTrace::analyze(&mut node, input, &RayTraceConfig::default())`"""
    import scenes
    osc = scenes.to_oracle([scenes.node("source"), scenes.node(kind, **node_kw)])
    return osc.raytrace(ray.copy())


def test_paraxial_surface_node_kats():
    """nodes/paraxial_surface.rs:316-458: analyze_geometric_ok (on axis: untouched, at z = 10 mm), test_shifted_x / _y (a lens
    decentred by its focal length sends the axis ray to normalize(1, 0, 1) / (0, 1, 1): assert_eq!), test_rotated_y / _x (45 deg
    tilts: position and direction with assert_relative_eq!)"""
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1000.0), 1.0)
    out = _one_node_chain("paraxial_surface", ray, z=mm(10.0), p=(mm(10.0),))
    assert orc.nr_of_rays(out) == 1 and tuple(out[0]["pos"]) == mm(0.0, 0.0, 10.0) and tuple(out[0]["dir"]) == Z
    s = 1.0 / math.sqrt(2.0)
    out = _one_node_chain("paraxial_surface", ray, z=mm(10.0), xy=(mm(10.0), 0.0), p=(mm(10.0),))
    want = np.array([1.0, 0.0, 1.0]) / np.linalg.norm([1.0, 0.0, 1.0])
    assert orc.nr_of_rays(out) == 1 and tuple(out[0]["pos"]) == mm(0.0, 0.0, 10.0) and np.array_equal(out[0]["dir"], want)
    out = _one_node_chain("paraxial_surface", ray, z=mm(10.0), xy=(0.0, mm(10.0)), p=(mm(10.0),))
    want = np.array([0.0, 1.0, 1.0]) / np.linalg.norm([0.0, 1.0, 1.0])
    assert tuple(out[0]["pos"]) == mm(0.0, 0.0, 10.0) and np.array_equal(out[0]["dir"], want)
    # tilted by 45 deg about x ("rotated_y"): the ray starts 10/sqrt(2) mm above the axis
    ray = orc.ray_new(mm(0.0, 10.0 / math.sqrt(2.0), 0.0), Z, nm(1000.0), 1.0)
    out = _one_node_chain("paraxial_surface", ray, z=mm(10.0), euler=(math.radians(45.0), 0.0, 0.0), p=(mm(10.0),))
    rel = lambda a, b: abs(a - b) <= EPS or abs(a - b) <= EPS * max(abs(a), abs(b))  # noqa: E731  approx::relative_eq defaults
    assert orc.nr_of_rays(out) == 1
    assert abs(out[0]["pos"][0]) <= EPS and rel(out[0]["pos"][1], 0.01 / math.sqrt(2.0)) and rel(out[0]["pos"][2], 0.01 / math.sqrt(2.0) + 0.01)
    assert np.all(np.abs(out[0]["dir"] - np.array([0.0, -1.0, 1.0]) * s) <= EPS)
    ray = orc.ray_new(mm(-10.0 / math.sqrt(2.0), 0.0, 0.0), Z, nm(1000.0), 1.0)
    out = _one_node_chain("paraxial_surface", ray, z=mm(10.0), euler=(0.0, math.radians(45.0), 0.0), p=(mm(10.0),))
    assert rel(out[0]["pos"][0], -0.01 / math.sqrt(2.0)) and abs(out[0]["pos"][1]) <= EPS and rel(out[0]["pos"][2], 0.01 / math.sqrt(2.0) + 0.01)
    assert np.all(np.abs(out[0]["dir"] - np.array([1.0, 0.0, 1.0]) * s) <= EPS)


def _hexapolar_bundle(radius_mm, rings, wvl_nm=1000.0, energy=1.0):
    """Rays::new_uniform_collimated(wavelength, energy, &Hexapolar::new(radius, rings)) (rays.rs:234-247)"""
    xyz = orc.hexapolar(rings, mm(radius_mm))
    return orc.rays_collimated_with_spectrum(xyz, orc.energy_uniform(len(xyz), energy), [nm(wvl_nm)], [1.0])


@pytest.mark.parametrize("kind", ["lens", "cylindric_lens"])
def test_lens_node_kats(kind):
    """nodes/lens/mod.rs:440-501 and nodes/cylindric_lens/mod.rs:359-430: analyze_flatflat -- a 10 mm plane-parallel plate of index
    2 at z = 10 mm leaves every ray of a hexapolar bundle along z with an optical path of exactly 30 mm (assert_eq!);
    analyze_biconvex -- a biconvex lens of index 1 is neutral (Lens: assert_eq!, CylindricLens: assert_relative_eq!)"""
    out = _one_node_chain(kind, _hexapolar_bundle(10.0, 3), z=mm(10.0), p=(math.inf, -math.inf, mm(10.0)), index=("const", 2.0))
    assert len(out) == 37 and orc.nr_of_rays(out) == 37
    for r in out:
        assert tuple(r["dir"]) == Z and r["opl"] == mm(30.0)
    out = _one_node_chain(kind, _hexapolar_bundle(10.0, 3), p=(mm(100.0), mm(-100.0), mm(10.0)), index=("const", 1.0))
    for r in out:
        if kind == "lens":
            assert tuple(r["dir"]) == Z
        else:
            assert abs(r["dir"][0]) <= EPS and abs(r["dir"][1]) <= EPS and abs(r["dir"][2] - 1.0) <= EPS


def test_thin_mirror_and_wedge_node_kats():
    """nodes/thin_mirror.rs:376-399 (a flat mirror at z = 10 mm sends the axis ray back: position (0, 0, 10 mm), direction
    (0, 0, -1)) and nodes/wedge/mod.rs:393-413 (the default wedge -- 10 mm thick, no wedge angle -- at z = 10 mm: the ray leaves
    at (0, 0, 20 mm) along z); all assert_eq!"""
    ray = orc.ray_new((0.0, 0.0, 0.0), Z, nm(1000.0), 1.0)  # Ray::origin_along_z
    out = _one_node_chain("thin_mirror", ray, z=mm(10.0))
    assert orc.nr_of_rays(out) == 1 and tuple(out[0]["pos"]) == mm(0.0, 0.0, 10.0) and tuple(out[0]["dir"]) == (0.0, 0.0, -1.0)
    out = _one_node_chain("wedge", ray, z=mm(10.0), p=(mm(10.0), 0.0), index=("const", 1.5))
    assert orc.nr_of_rays(out) == 1 and tuple(out[0]["pos"]) == mm(0.0, 0.0, 20.0) and tuple(out[0]["dir"]) == Z


def test_beam_splitter_and_filter_node_kats():
    """nodes/beam_splitter/analysis_raytrace.rs:122-147 analyze_one_input: Ratio(0.6) sends 0.6 J of a 1 J ray to the first output
    and 0.4 J to the second (assert_eq!); nodes/ideal_filter.rs:373-399 analyzer_geometric_fixed: a Constant(0.3) filter leaves
    0.3 J of a 1 J hexapolar bundle"""
    import scenes
    osc = scenes.to_oracle([scenes.node("source"), scenes.node("beam_splitter", p=(0.6,))])
    out = osc.raytrace(orc.ray_new((0.0, 0.0, 0.0), Z, nm(1053.0), 1.0))
    assert orc.total_energy(out) == 0.6 and orc.total_energy(osc.nodes[1].split) == 0.4
    out = _one_node_chain("ideal_filter", _hexapolar_bundle(5.0, 1, 1054.0), p=(0.3,))
    assert len(out) == 7 and abs(orc.total_energy(out) - 0.3) <= EPS


def test_bundle_constructor_kats():
    """rays.rs:1888-1927: new_collimated with a General2DGaussian energy distribution (hexapolar, 2 rings) carries 1 J in total;
    new_uniform_collimated: 19 rays, 1 J within 10 eps; a hexapolar source of radius 0 is ONE ray at the origin along z"""
    xyz = orc.hexapolar(2, mm(1.0))
    e = orc.energy_general_gaussian(xyz, 1.0, (0.0, 0.0), mm(1.0, 1.0), 1.0, 0.0, True)
    rays = orc.rays_collimated_with_spectrum(xyz, e, [nm(1054.0)], [1.0])
    assert abs(orc.total_energy(rays) - 1.0) <= EPS
    rays = orc.rays_collimated_with_spectrum(xyz, orc.energy_uniform(len(xyz), 1.0), [nm(1054.0)], [1.0])
    assert orc.nr_of_rays(rays) == 19 and abs(orc.total_energy(rays) - 1.0) < 10 * EPS
    xyz0 = orc.hexapolar(2, 0.0)
    rays = orc.rays_collimated_with_spectrum(xyz0, orc.energy_uniform(len(xyz0), 1.0), [nm(1054.0)], [1.0])
    assert orc.nr_of_rays(rays) == 1 and tuple(rays[0]["pos"]) == (0.0, 0.0, 0.0) and tuple(rays[0]["dir"]) == Z


def test_kde_fluence_end_to_end_kats():
    """surface/hit_map/mod.rs:1124-1169 calc_combined_fluence_with_kde: five 1 J hits (the corners of a unit square and its centre)
    on a 51 x 51 map give 5.474418964842738 J/m^2 in the centre bin; merged with five more (a square rotated by 45 deg and the
    centre again) 8.969644069111087 -- the whole estimator (pair distances, Silverman bandwidth, range with its 3-bandwidth
    margin, Gaussian sum) end to end, assert_relative_eq!"""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    r = math.sqrt(0.5)
    p2 = [(r, 0.0), (0.0, r), (-r, 0.0), (0.0, -r), (0.0, 0.0)]
    rel = lambda a, b: abs(a - b) <= EPS or abs(a - b) <= EPS * max(abs(a), abs(b))  # noqa: E731
    fmap, _, _ = orc.fluence_kde(orc.hits_from_xye([p[0] for p in p1], [p[1] for p in p1], [1.0] * 5), 51, 51)
    assert rel(fmap[25, 25], 5.474418964842738)
    both = p1 + p2  # HitMap::get_merged_rays_hit_map: the hit points of both bundles in insertion order
    fmap, _, _ = orc.fluence_kde(orc.hits_from_xye([p[0] for p in both], [p[1] for p in both], [1.0] * 10), 51, 51)
    assert rel(fmap[25, 25], 8.969644069111087)


def test_beam_splitter_two_inputs_kat():
    """nodes/beam_splitter/analysis_raytrace.rs:148-178 analyze_two_input: Ratio(0.6) with 1 J at input 1 and 0.5 J at input 2 gives
    0.8 J at the first output (0.6 * 1 + 0.4 * 0.5) and 0.7 J at the second (assert_abs_diff_eq!); no input, no output (:113-120)"""
    import scenes
    osc = scenes.to_oracle([scenes.node("source"), scenes.node("beam_splitter", p=(0.6,))])
    in1 = orc.ray_new(mm(0.0, 0.0, -10.0), Z, nm(1053.0), 1.0)
    in2 = orc.ray_new(mm(0.0, 0.0, -10.0), Z, nm(1053.0), 0.5)
    out1, out2 = osc.beam_splitter_two_inputs(1, in1, in2)
    assert abs(orc.total_energy(out1) - 0.8) <= EPS and abs(orc.total_energy(out2) - 0.7) <= EPS
    assert len(out1) == 2 and len(out2) == 2
    assert osc.beam_splitter_two_inputs(1, None, None) == (None, None)
    only1 = osc.beam_splitter_two_inputs(1, in1, None)
    assert orc.total_energy(only1[0]) == 0.6 and orc.total_energy(only1[1]) == 0.4


# ---------------------------------------------------------------------------------------------------------------
# Voronoi fluence estimator (surface/hit_map/mod.rs:942-998, utils/griddata.rs tests)
# ---------------------------------------------------------------------------------------------------------------
def test_kat_voronoi_combined_fluence_centre_bins():
    """hit_map/mod.rs:942-982 `calc_combined_fluence_with_voronoi`: five 1 J hits (unit square + centre) -> centre bin 2 J/m^2
    (the centre's cell is the diamond of area 0.5 m^2); a second bundle (diamond + centre) on top -> 4 J/m^2"""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    h1 = orc.hits_from_xye(np.array([p[0] for p in p1]), np.array([p[1] for p in p1]), np.ones(5))
    m, lrtb = orc.fluence_voronoi_combined([h1], 51, 51)
    assert m[25, 25] == pytest.approx(2.0, rel=1e-14) and lrtb == (-0.5, 0.5, 0.5, -0.5)
    r = math.sqrt(0.5)
    p2 = [(r, 0.0), (0.0, r), (-r, 0.0), (0.0, -r), (0.0, 0.0)]
    h2 = orc.hits_from_xye(np.array([p[0] for p in p2]), np.array([p[1] for p in p2]), np.ones(5))
    m, _ = orc.fluence_voronoi_combined([h1, h2], 51, 51)
    assert m[25, 25] == pytest.approx(4.0, rel=1e-14)


def test_kat_voronoi_too_few_points_and_collinear():
    """hit_map/mod.rs:984-998 (two hits: Err), griddata.rs `all_points_on_same_line_test` (collinear points: no triangulation)"""
    two = orc.hits_from_xye(np.array([-0.5, 0.0]), np.array([-0.5, 0.0]), np.ones(2))
    with pytest.raises(orc.OracleError):
        orc.fluence_voronoi(two, 51)
    line = orc.hits_from_xye(np.array([0.0, 1.0, 2.0]), np.array([0.0, 1.0, 2.0]), np.ones(3))
    with pytest.raises(orc.OracleError):
        orc.fluence_voronoi(line, 51)
    ok = orc.hits_from_xye(np.array([0.0, 1.0, 2.0, 2.0]), np.array([0.0, 1.0, 2.0, 3.0]), np.ones(4))
    orc.fluence_voronoi(ok, 11)


def test_voronoi_cells_tile_the_bounding_polygon_and_interpolation_reproduces_linear_data():
    """properties of the restatement (not reference KATs): the cells partition the bounding polygon, so sum(E_i / fluence_i) is its
    area; natural-neighbour coordinates reproduce a linear function of (x, y) exactly inside the hull"""
    rng = np.random.default_rng(5)
    n = 400
    x, y = rng.random(n) * 3.0 - 1.0, rng.random(n) * 2.0 + 0.5  # cm
    fl, area_hull = orc.voronoi_site_fluence(x, y, np.ones(n))
    total_cells = float(np.sum(1.0 / fl))
    assert area_hull < total_cells < 1.25 * area_hull  # the hull pushed outwards by half a mean triangle height
    # a rays hit map whose per-site fluence is linear in (x, y): E_i = f(x_i, y_i) * area_i
    f = 2.0 + 0.7 * x - 0.3 * y
    hits = orc.hits_from_xye(x * 1e-2, y * 1e-2, f / fl)
    m, lrtb = orc.fluence_voronoi(hits, 41)
    ax = np.linspace(x.min(), x.max(), 41); ay = np.linspace(y.min(), y.max(), 41)
    want = (2.0 + 0.7 * ax[None, :] - 0.3 * ay[:, None]) * 1e4
    inside = np.isfinite(m)
    assert inside.sum() > 0.8 * m.size
    assert np.allclose(m[inside], want[inside], rtol=1e-9)


# ------------------------------------------------------------------ fluence helper rays
def test_kat_helper_ray_fluence():
    """ray.rs:1819-1843: a ray with FluenceRays of 1 J/cm^2 reports that fluence (1e-8); ray.rs:1892-1920: four times the effective
    energy is four times the fluence; rays.rs:1484-1488: a negative or non-finite initial fluence is an error"""
    r = orc.ray_new((1.0, 1.0, 1.0), (0.0, 0.0, 1.0), 1000e-9, 1.0)
    h = orc.helpers_new(r, 1.0e4)
    assert abs(orc.helper_ray_fluence(h, 0) / 10000.0 - 1.0) <= 1e-8
    h["eff"] *= 4.0
    assert abs(orc.helper_ray_fluence(h, 0) / 40000.0 - 1.0) <= 1e-8
    for bad in (-1.0, math.nan, math.inf, -math.inf):
        with pytest.raises(ValueError):
            orc.helpers_new(r, bad)


def test_kat_new_collimated_w_fluence_helper_gaussian():
    """rays.rs:2846-2873: Hexapolar(15 mm, 5 rings) positions, General2DGaussian(1 J, sigma 2.5 mm) fluences: the first ray (the
    centre) carries the peak fluence 1 / (2 pi sigma^2) to 2 ulp"""
    xyz = orc.hexapolar(5, 15e-3)
    fl = orc.fluence_general_gaussian(xyz[:, :2], 1.0, (0.0, 0.0), (2.5e-3, 2.5e-3), 0.0)
    rays = orc.rays_from_arrays(xyz, np.tile([0.0, 0.0, 1.0], (len(xyz), 1)), np.ones(len(xyz)), np.full(len(xyz), 1000e-9))
    h = orc.helpers_new(rays, fl)
    assert tuple(xyz[0]) == (0.0, 0.0, 0.0)
    assert abs(orc.helper_ray_fluence(h, 0) * (2.0 * math.pi * 0.0025 * 0.0025) - 1.0) <= 1e-9  # the triangle's area carries 1e-10 of rounding
    assert abs(fl[0] * (2.0 * math.pi * 0.0025 * 0.0025) - 1.0) <= 2 * np.finfo(float).eps


def test_kat_helper_rays_estimator():
    """hit_map/mod.rs:1030-1068: five FluenceHitPoints of 1 J/cm^2 -> 10000 J/m^2 at the centre of a 51 x 51 map; a second bundle of
    five on the common bounding box doubles it.  :1072-1091: fewer than three points is an error."""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    s = math.sqrt(0.5)
    p2 = [(s, 0.0), (0.0, s), (-s, 0.0), (0.0, -s), (0.0, 0.0)]
    h1 = orc.hits_from_xye([p[0] for p in p1], [p[1] for p in p1], [1.0e4] * 5)
    h2 = orc.hits_from_xye([p[0] for p in p2], [p[1] for p in p2], [1.0e4] * 5)
    m, lrtb = orc.fluence_helper_rays_combined([h1], 51, 51)
    assert m[25, 25] == pytest.approx(10000.0, rel=1e-12)
    m, lrtb = orc.fluence_helper_rays_combined([h1, h2], 51, 51)
    assert m[25, 25] == pytest.approx(20000.0, rel=1e-12)
    with pytest.raises(Exception):
        orc.fluence_helper_rays(h1[:2], 50)


def test_helper_rays_follow_a_plane_refraction():
    """rays.rs:2165-2190 runs Rays::refract_on_surface on a helper-ray bundle (no assertion there).  Here with numbers: a collimated
    bundle through a plane into n = 1.5 keeps its helper triangles (area wavelength^2), so the recorded FluenceHitPoint is the initial
    fluence, and the effective energy drops by the Fresnel factor 1 - R = 0.96 at normal incidence"""
    xyz = orc.hexapolar(2, 5e-3)
    n = len(xyz)
    rays = orc.rays_from_arrays(xyz, np.tile([0.0, 0.0, 1.0], (n, 1)), np.ones(n), np.full(n, 1000e-9))
    h = orc.helpers_new(rays, 2.0e4)
    surf = orc.Surface(geom=orc.GEOM_PLANE, coating=orc.COAT_FRESNEL, param=0.0, coating_r=0.0, iso=orc.Isometry.new((0.0, 0.0, 0.01)))
    ch, chh, hits = orc.rays_refract_on_surface_w_helpers(rays, h, surf, orc.IndexModel.const(1.5))
    assert len(hits) == n and np.allclose(hits["e"], 2.0e4, rtol=1e-8)
    assert np.allclose(h["eff"], 2.0e4 * (1000e-9) ** 2 * 0.96, rtol=1e-12)
    assert np.allclose(chh["eff"], 2.0e4 * (1000e-9) ** 2 * 0.04, rtol=1e-12)
    assert np.allclose(h["rays"]["pos"][:, :, 2], 0.01, atol=1e-15) and np.allclose(chh["rays"]["pos"][:, :, 2], 0.0)  # the clone's helpers stay behind
    assert np.allclose(chh["rays"]["dir"][:, :, 2], -1.0)  # ... with the reflected direction
    for i in range(n):
        assert abs(orc.helper_ray_fluence(h, i) / (2.0e4 * 0.96) - 1.0) <= 1e-7
