"""GPU parity of the Voronoi fluence estimator (opossum_b200/csrc/opm_voronoi.cu; the reference's default,
nodes/fluence_detector/mod.rs:61, analyzers/ghostfocus.rs:61) against the oracle's brute-force restatement
(oracle/opm_voronoi.c): the reference's known answers (hit_map/mod.rs:942-998), seeded hit sets of several shapes (round,
anisotropic, gridded with collinear hull points, clustered), explicit ranges, the combination of several rays hit maps, errors."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import mm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def hitmap_from(ctx, x, y, e, bounce=None):
    h = ob.HitMap(ctx, max(len(x), 1))
    h.upload(x, y, e, bounce=bounce)
    return h


def assert_voronoi_match(got, ref, rtol=1e-9):
    """same NaN pattern (outside the convex hull), finite bins to 1e-9 relative"""
    assert got.shape == ref.shape
    gn, rn = np.isnan(got), np.isnan(ref)
    assert np.array_equal(gn, rn), f"{(gn != rn).sum()} bins differ in being inside the convex hull"
    ok = ~rn
    err = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert err.max() <= rtol, f"worst relative error {err.max():.3e}"


def test_voronoi_reference_kats(ctx):
    """hit_map/mod.rs:942-982: centre bin 2 J/m^2, then 4 J/m^2 with the second bundle on top; :984-998 too few points"""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    h1 = hitmap_from(ctx, [p[0] for p in p1], [p[1] for p in p1], [1.0] * 5)
    fd = ob.fluence_voronoi_combined(ctx, [[h1]], 51, 51)
    assert fd.map[25, 25] == pytest.approx(2.0, rel=1e-14) and fd.lrtb == (-0.5, 0.5, 0.5, -0.5)
    r = math.sqrt(0.5)
    p2 = [(r, 0.0), (0.0, r), (-r, 0.0), (0.0, -r), (0.0, 0.0)]
    h2 = hitmap_from(ctx, [p[0] for p in p2], [p[1] for p in p2], [1.0] * 5)
    fd = ob.fluence_voronoi_combined(ctx, [[h1], [h2]], 51, 51)
    assert fd.map[25, 25] == pytest.approx(4.0, rel=1e-14)
    ref, _ = orc.fluence_voronoi_combined([orc.hits_from_xye(np.array([p[0] for p in p1]), np.array([p[1] for p in p1]), np.ones(5)),
                                           orc.hits_from_xye(np.array([p[0] for p in p2]), np.array([p[1] for p in p2]), np.ones(5))], 51, 51)
    assert_voronoi_match(fd.map, ref, rtol=1e-12)
    with pytest.raises(ob.OpossumError) as ei:
        ob.fluence_voronoi(ctx, [hitmap_from(ctx, [-0.5, 0.0], [-0.5, 0.0], [1.0, 1.0])], 51)
    assert ei.value.code == F.OPM_ERR_VORONOI
    with pytest.raises(ob.OpossumError):  # collinear points: no triangulation (griddata.rs:281-283)
        ob.fluence_voronoi(ctx, [hitmap_from(ctx, [0.0, 1.0, 2.0], [0.0, 1.0, 2.0], [1.0] * 3)], 51)
    with pytest.raises(ValueError):       # the reference panics on nx != ny
        ob.fluence_voronoi_combined(ctx, [[h1]], 100, 83)


def _shape(name, n, rng):
    if name == "round":       # a Fibonacci disc like the bench's sources: hundreds of hull vertices
        xyz = orc.fibonacci_ellipse(n, mm(4.0), mm(4.0))
        return xyz[:, 0] + mm(0.2), xyz[:, 1] - mm(0.1)
    if name == "anisotropic":
        return mm(3.0) * rng.standard_normal(n), mm(0.2) * rng.standard_normal(n)
    if name == "grid":        # collinear points on every hull edge, ties everywhere
        s = int(math.sqrt(n))
        gx, gy = np.meshgrid(np.linspace(-mm(2.0), mm(2.0), s), np.linspace(-mm(1.0), mm(3.0), s), indexing="ij")
        return gx.ravel(), gy.ravel()
    cx, cy = mm(2.0) * rng.standard_normal(8), mm(2.0) * rng.standard_normal(8)  # clustered: cells of very different sizes
    k = rng.integers(0, 8, n)
    return cx[k] + mm(0.05) * rng.standard_normal(n), cy[k] + mm(0.05) * rng.standard_normal(n)


@pytest.mark.parametrize("shape,n,nx", [("round", 3000, 64), ("anisotropic", 2000, 50), ("grid", 1600, 61), ("clustered", 1500, 48),
                                        ("round", 40, 32)])
def test_voronoi_matches_oracle(ctx, shape, n, nx):
    rng = np.random.default_rng(11)
    x, y = _shape(shape, n, rng)
    e = 1e-6 * (1.0 + rng.random(len(x)))
    ref, ref_lrtb = orc.fluence_voronoi(orc.hits_from_xye(x, y, e), nx)
    fd = ob.fluence_voronoi(ctx, [hitmap_from(ctx, x, y, e)], nx)
    assert fd.lrtb == ref_lrtb
    assert_voronoi_match(fd.map, ref)
    assert fd.peak == pytest.approx(orc.fluence_peak(ref.T.ravel()), rel=1e-9)
    assert np.isfinite(ref).sum() > 0.3 * ref.size


def test_voronoi_explicit_ranges_bounce_levels_and_combination(ctx):
    """two bundles on one surface, one of them with hits of two bounce levels: three rays hit maps (hit_map/mod.rs:101-115), each
    estimated on the common bounding box and summed (:232-262)"""
    rng = np.random.default_rng(3)
    xa, ya = mm(2.0) * rng.standard_normal(900), mm(1.5) * rng.standard_normal(900)
    ea = 1e-6 * (1.0 + rng.random(900))
    ba = (rng.random(900) < 0.3).astype(np.uint32) * 2  # bounce level 0 or 2
    xb, yb = mm(0.8) * rng.standard_normal(500) + mm(0.5), mm(0.8) * rng.standard_normal(500)
    eb = 2e-6 * (1.0 + rng.random(500))
    ha, hb = hitmap_from(ctx, xa, ya, ea, bounce=ba), hitmap_from(ctx, xb, yb, eb)
    assert ob.hitmap_bounce_levels(ha) == [0, 2] and ob.hitmap_bounce_levels(hb) == [0]
    parts = [orc.hits_from_xye(xa[ba == 0], ya[ba == 0], ea[ba == 0]), orc.hits_from_xye(xa[ba == 2], ya[ba == 2], ea[ba == 2]),
             orc.hits_from_xye(xb, yb, eb)]
    ref, ref_lrtb = orc.fluence_voronoi_combined(parts, 40, 40)
    fd = ob.fluence_voronoi_combined(ctx, [[ha], [hb]], 40, 40)
    assert fd.lrtb == ref_lrtb
    assert_voronoi_match(fd.map, ref)
    # one rays hit map with explicit ranges narrower than the data
    rg = (-mm(1.0), mm(1.5), -mm(0.5), mm(1.0))
    ref1, lr1 = orc.fluence_voronoi(parts[2], 33, rg)
    fd1 = ob.fluence_voronoi(ctx, [hb], 33, ranges=rg)
    assert fd1.lrtb == lr1
    assert_voronoi_match(fd1.map, ref1)
