"""GPU parity of the detector and source kernels against the oracle: Binning fluence estimator
(rays_hit_map.rs:330-391; shared-memory privatised and global-atomic paths), bounding box, peak/total
(fluence_data.rs), on-device sources (position/energy/spectral distributions, rays.rs:264-297)."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import RTOL, assert_map_match, assert_rays_match, iso_pair, mm, nm

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


def seeded_hits(n, seed=3, spread=mm(4.0)):
    rng = np.random.default_rng(seed)
    x = spread * rng.standard_normal(n) * 0.4 + mm(0.3)
    y = spread * rng.standard_normal(n) * 0.25 - mm(0.1)
    e = 1e-6 * (1.0 + rng.random(n))
    return x, y, e


def test_binning_reference_kat(ctx):
    """surface/hit_map/mod.rs:1200-1240: 5 x 1 J on a 51x51 map -> centre bin 2601.0, again after merging 5 more"""
    p1 = [(-0.5, -0.5), (0.0, 0.0), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5)]
    h1 = hitmap_from(ctx, [p[0] for p in p1], [p[1] for p in p1], [1.0] * 5)
    fd = ob.fluence_binning(ctx, [h1], 51, 51)
    assert fd.map[25, 25] == pytest.approx(2601.0, rel=4e-16) and fd.lrtb == (-0.5, 0.5, 0.5, -0.5)
    assert fd.peak == fd.map.max() and fd.total_energy == pytest.approx(5.0, rel=1e-14)
    r = math.sqrt(0.5)
    p2 = [(r, 0.0), (0.0, r), (-r, 0.0), (0.0, -r), (0.0, 0.0)]
    h2 = hitmap_from(ctx, [p[0] for p in p2], [p[1] for p in p2], [1.0] * 5)
    fd = ob.fluence_binning(ctx, [h1, h2], 51, 51)  # merged hit maps (hit_map/mod.rs:161-175)
    assert fd.map[25, 25] == pytest.approx(2601.0, rel=1e-15)


@pytest.mark.parametrize("nx,ny", [(100, 83), (101, 101), (51, 7), (400, 300), (1024, 1024)])
def test_binning_auto_bbox_matches_oracle(ctx, nx, ny):
    """data-dependent bounding box (rays_hit_map.rs:522-562), the FluenceDetector's only mode (hit_map/mod.rs:369);
    100x83 / 101x101 use the shared-memory privatised map, the larger ones warp-run aggregated global atomics"""
    x, y, e = seeded_hits(200_000)
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), nx, ny)
    h = hitmap_from(ctx, x, y, e)
    box, n = ob.hitmaps_bounding_box([h])
    assert n == len(x) and box == ref_lrtb == (x.min(), x.max(), y.max(), y.min())
    fd = ob.fluence_binning(ctx, [h], nx, ny)
    assert fd.lrtb == ref_lrtb
    assert_map_match(fd.map, ref_map, same_hits=True)
    assert fd.peak == pytest.approx(orc.fluence_peak(ref_map), rel=RTOL)
    assert fd.total_energy == pytest.approx(orc.fluence_total_energy(ref_map, ref_lrtb), rel=RTOL)


def test_binning_coherent_runs(ctx):
    """grid-ordered hits: long runs of consecutive hits in the same bin exercise the warp segmented reduction"""
    g = orc.grid(600, 500, mm(10.0), mm(8.0))
    x, y = g[:, 0].copy(), g[:, 1].copy()
    e = 1e-6 * (1.0 + 0.3 * np.cos(1e3 * x) * np.sin(2e3 * y))
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), 300, 200)
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e)], 300, 200)
    assert_map_match(fd.map, ref_map, same_hits=True)
    assert fd.total_energy == pytest.approx(orc.fluence_total_energy(ref_map, ref_lrtb), rel=RTOL)


def test_binning_explicit_ranges_and_errors(ctx):
    """explicit ranges follow rays_hit_map.rs:337-339 literally: (left,right,top,bottom) = (r1.start,r1.end,r2.start,r2.end);
    hits right of / above the range index out of bounds (the reference panics) -> error; empty map -> error"""
    x, y, e = seeded_hits(50_000, spread=mm(1.0))
    lo, hi = mm(-6.0), mm(6.0)
    ranges = (lo, hi, hi, lo)  # ax2 given as top..bottom so that y indices run upwards
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), 64, 48, ranges)
    h = hitmap_from(ctx, x, y, e)
    fd = ob.fluence_binning(ctx, [h], 64, 48, ranges=ranges)
    assert fd.lrtb == ref_lrtb == ranges
    assert_map_match(fd.map, ref_map, same_hits=True)
    with pytest.raises(orc.OracleError):
        orc.fluence_binning(orc.hits_from_xye(x, y, e), 64, 48, (lo, mm(0.1), hi, lo))
    with pytest.raises(ob.OpossumError) as ei:
        ob.fluence_binning(ctx, [h], 64, 48, ranges=(lo, mm(0.1), hi, lo))
    assert ei.value.code == F.OPM_ERR_BIN_RANGE
    empty = hitmap_from(ctx, x, y, -np.ones_like(e))  # e < 0 marks "no hit recorded in this slot"
    with pytest.raises(ob.OpossumError) as ei:
        ob.fluence_binning(ctx, [empty], 10, 10)
    assert ei.value.code == F.OPM_ERR_EMPTY_HITMAP


def test_binning_bounce_filter_and_accumulate(ctx):
    """hit maps are keyed by bounce level (hit_map/mod.rs:101-115): filter one level; accumulate two passes into one map"""
    x, y, e = seeded_hits(80_000)
    b = (np.arange(len(x)) % 3).astype(np.uint32)
    h = hitmap_from(ctx, x, y, e, bounce=b)
    sel = b == 1
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x[sel], y[sel], e[sel]), 101, 101)
    fd = ob.fluence_binning(ctx, [h], 101, 101, bounce=1)
    assert fd.lrtb == ref_lrtb
    assert_map_match(fd.map, ref_map, same_hits=True)
    # two shards binned into one device map with a shared (externally reduced) bounding box: the multi-GPU pattern
    full_map, full_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), 200, 150)
    half = len(x) // 2
    ha, hb = hitmap_from(ctx, x[:half], y[:half], e[:half]), hitmap_from(ctx, x[half:], y[half:], e[half:])
    (la, _), (lb, _) = ob.hitmaps_bounding_box([ha]), ob.hitmaps_bounding_box([hb])
    box = (min(la[0], lb[0]), max(la[1], lb[1]), max(la[2], lb[2]), min(la[3], lb[3]))
    assert box == full_lrtb
    dev = ctx.device_alloc(200 * 150 * 8)
    ob.fluence_binning(ctx, [ha], 200, 150, lrtb=box, map_dev=dev, download=False)
    fd = ob.fluence_binning(ctx, [hb], 200, 150, lrtb=box, map_dev=dev, accumulate=True)
    ctx.device_free(dev)
    assert_map_match(fd.map, full_map, same_hits=True)


def test_sources_match_oracle(ctx):
    """rays.rs:264-297 with grid.rs:58-92, fibonacci.rs:53-65,113-126, uniform.rs:37-41, general_gaussian.rs:93-120,
    hexapolar.rs:38-56, spectral_distribution/gaussian.rs:78-96; position-major / wavelength-minor order; source isometry"""
    wvl, w = orc.spectrum_gaussian(nm(950.0), nm(1150.0), 8, nm(1053.0), nm(50.0))
    io, ig = iso_pair((mm(1.0), mm(-2.0), mm(5.0)), (0.01, -0.02, 0.3))
    cases = {
        "grid": (lambda: orc.grid(37, 29, mm(20.0), mm(15.0)), lambda s: s.grid(37, 29, mm(20.0), mm(15.0)), 37 * 29),
        "fib_rect": (lambda: orc.fibonacci_rectangle(1000, mm(20.0), mm(10.0)), lambda s: s.fibonacci_rectangle(mm(20.0), mm(10.0)), 1000),
        "fib_ellipse": (lambda: orc.fibonacci_ellipse(1000, mm(10.0), mm(7.0)), lambda s: s.fibonacci_ellipse(mm(10.0), mm(7.0)), 1000),
        "hexapolar": (lambda: orc.hexapolar(17, mm(8.0)), lambda s: s.hexapolar(17, mm(8.0)), 1 + 3 * 17 * 18),
        "hexapolar255": (lambda: orc.hexapolar(255, mm(8.0)), lambda s: s.hexapolar(255, mm(8.0)), 1 + 3 * 255 * 256),
    }
    for name, (opos, gset, npos) in cases.items():
        for gauss in (False, True):
            xyz = opos()
            e_pos = (orc.energy_general_gaussian(xyz, 2.0, (mm(0.5), 0.0), (mm(4.0), mm(3.0)), 5.0, 0.2, name == "grid")
                     if gauss else orc.energy_uniform(npos, 2.0))
            ref = orc.rays_collimated_with_spectrum(xyz, e_pos, wvl, w)
            orc.rays_transform(ref, io)
            sd = gset(ob.SourceDesc(npos, wvl, w, total_energy=2.0, iso=ig))
            if gauss:
                sd.general_gaussian((mm(0.5), 0.0), (mm(4.0), mm(3.0)), 5.0, 0.2, rectangular=(name == "grid"))
            got = sd.generate(ctx).to_array()
            assert_rays_match(got, ref, what=f"source {name} gauss={gauss}")
            # a shard [begin, end) of positions equals the same slice of the full source (multi-GPU sharding by index)
            b, e_ = npos // 3, 2 * npos // 3
            if gauss:
                sd.c.energy_norm = sd.energy_norm(ctx)
            part = sd.generate(ctx, b, e_).to_array()
            assert_rays_match(part, ref[b * len(wvl): e_ * len(wvl)], what=f"source shard {name}")


# ------------------------------------------------------------------ KDE estimator (kde/mod.rs)
def _kde_hits(ctx, xyz, e):
    hm = ob.HitMap(ctx, len(e))
    hm.upload(xyz[:, 0], xyz[:, 1], e)
    return hm, orc.hits_from_xye(xyz[:, 0], xyz[:, 1], e)


@pytest.mark.parametrize("case", ["hexapolar37", "fibonacci2000", "three_points", "explicit_ranges"])
def test_kde_fluence_matches_oracle(ctx, case):
    """calc_fluence_with_kde (rays_hit_map.rs:575-613): bandwidth (Silverman on all pair distances: std dev, 75 % quantile),
    map range (bounding box + 3 bandwidths) and the Gaussian sum per map point against the oracle.  hexapolar37 -> 30x30
    is the hit set of the reference's own `kde` criterion bench (benches/fluence_estimator.rs:12-24)."""
    ranges = None
    if case == "hexapolar37":
        xyz, nx, ny = orc.hexapolar(3, mm(5.0)), 30, 30
    elif case == "fibonacci2000":
        xyz, nx, ny = orc.fibonacci_ellipse(2000, mm(4.0), mm(3.0)), 64, 48
    elif case == "three_points":
        xyz, nx, ny = np.array([[0.0, 0.0, 0.0], [1e-3, 0.0, 0.0], [-1e-3, 0.0, 0.0]]), 20, 10
    else:
        xyz, nx, ny = orc.fibonacci_rectangle(500, mm(6.0), mm(6.0)), 40, 40
        ranges = (mm(-2.0), mm(2.5), mm(3.0), mm(-1.0))  # (left, right, top, bottom) as the reference reads them
    e = 1e-3 * (1.0 + 0.3 * np.cos(np.arange(len(xyz))))
    hm, hits = _kde_hits(ctx, xyz, e)
    ref_map, ref_lrtb, ref_bw = orc.fluence_kde(hits, nx, ny, ranges)
    fd, bw = ob.fluence_kde(ctx, [hm], nx, ny, ranges)
    assert bw == pytest.approx(ref_bw, rel=1e-12)
    if case == "three_points":
        assert abs(bw - 0.00034057440111656337) < 1e-15  # kde/mod.rs:255-261
    assert np.allclose(fd.lrtb, ref_lrtb, rtol=1e-12, atol=1e-18)
    assert_map_match(fd.map, ref_map, same_hits=True)
    assert fd.peak == pytest.approx(orc.fluence_peak(ref_map), rel=RTOL)


def test_kde_errors_and_bounce_filter(ctx):
    """fewer than two hit points or coincident points: 'bandwidth must be != 0.0 and finite' (kde/mod.rs:36-42,83-98);
    the bounce filter selects the hit points of one bounce level (hit_map/mod.rs:101-115)"""
    xyz = orc.fibonacci_ellipse(300, mm(3.0), mm(3.0))
    e = np.full(300, 1e-3)
    hm = ob.HitMap(ctx, 300)
    bounce = (np.arange(300) % 3).astype(np.uint32)
    hm.upload(xyz[:, 0], xyz[:, 1], e, bounce=bounce)
    sel = bounce == 1
    ref_map, _, ref_bw = orc.fluence_kde(orc.hits_from_xye(xyz[sel, 0], xyz[sel, 1], e[sel]), 32, 32)
    fd, bw = ob.fluence_kde(ctx, [hm], 32, 32, bounce=1)
    assert bw == pytest.approx(ref_bw, rel=1e-12)
    assert_map_match(fd.map, ref_map, same_hits=True)
    for pts in ([[0.0, 0.0, 0.0]], [[1e-3, 2e-3, 0.0], [1e-3, 2e-3, 0.0]]):
        p = np.array(pts)
        h1 = ob.HitMap(ctx, len(p))
        h1.upload(p[:, 0], p[:, 1], np.ones(len(p)))
        with pytest.raises(ob.OpossumError) as ei:
            ob.fluence_kde(ctx, [h1], 8, 8)
        assert ei.value.code == F.OPM_ERR_KDE_BANDWIDTH


@pytest.mark.parametrize("nx,ny", [(100, 83), (101, 101), (12, 9)])
def test_binning_fixed_point_wide_energy_range(ctx, nx, ny):
    """the CTA-private fixed-point sums (binning_smem_fixed_kernel): energies over 14 orders of magnitude in one map, many hits
    per bin, hits of zero energy -- every bin within 1e-9 (floor 1e-9 of the peak) of the oracle's f64 sums, total energy 1e-12"""
    rng = np.random.default_rng(11)
    n = 300_000
    x = mm(3.0) * rng.random(n) - mm(1.0)
    y = mm(2.0) * rng.random(n) + mm(0.5)
    e = 10.0 ** rng.uniform(-14.0, 0.0, n)
    e[::97] = 0.0
    e[5] = 3.0  # one hit far above the rest sets the unit of its CTA
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), nx, ny)
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e)], nx, ny)
    assert fd.lrtb == ref_lrtb
    assert_map_match(fd.map, ref_map, same_hits=True)
    assert fd.total_energy == pytest.approx(orc.fluence_total_energy(ref_map, ref_lrtb), rel=1e-12)
    # tiny energies only: the unit follows the largest energy of the CTA's slots, nothing is rounded away
    e2 = 1e-200 * (1.0 + rng.random(n))
    ref2, _ = orc.fluence_binning(orc.hits_from_xye(x, y, e2), nx, ny)
    assert_map_match(ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e2)], nx, ny).map, ref2, same_hits=True)


def test_binning_fixed_point_non_finite_energy(ctx):
    """an infinite energy cannot be scaled: those taps go to the map as they are (like f64 adds: inf, or NaN where a tap weight
    is 0), the finite bins stay exact"""
    x = np.array([0.0, 1.0, 2.0, 3.0, 4.0, 2.0]) * mm(1.0)
    y = np.array([0.0, 1.0, 2.0, 3.0, 4.0, 2.0]) * mm(1.0)
    e = np.array([1.0, 2.0, np.inf, 4.0, 5.0, 1.0])
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e)], 5, 5)
    assert not np.isfinite(fd.map[2, 2]) and np.isfinite(fd.map[0, 0]) and np.isfinite(fd.map[4, 4])
    area = (mm(4.0) / 5) ** 2
    assert fd.map[0, 0] == pytest.approx(1.0 / area, rel=1e-12) and fd.map[4, 4] == pytest.approx(5.0 / area, rel=1e-12)
    assert fd.peak == pytest.approx(5.0 / area, rel=1e-12)  # fluence_data.rs:45-68: peak over the finite bins


@pytest.mark.parametrize("order", ["rows_up", "rows_down", "with_holes"])
def test_binning_large_map_run_merging(ctx, order):
    """2048 x 2048 map (32 MiB: the run-merging kernel, opm_kernels.cu binning_runs_kernel) from grid-ordered hits, 2.2 hits per
    bin along the fast axis like BASELINE configs[3]: consecutive hits share a bin or step to the next row (up, or down when the
    order is reversed), columns change inside a thread's run, empty slots (e < 0) interrupt runs"""
    g = orc.grid(900, 4500, mm(12.0), mm(9.0))  # x outer, y inner: y runs fastest in slot order
    x = g[:, 0] * (1.0 + 0.02 * np.sin(300.0 * g[:, 1]))
    y = g[:, 1] + mm(0.002) * np.cos(900.0 * g[:, 0])
    e = 1e-9 * (1.0 + 0.5 * np.sin(1e3 * x) * np.cos(2e3 * y))
    if order == "rows_down":
        x, y, e = x[::-1].copy(), y[::-1].copy(), e[::-1].copy()
    e_up = e.copy()
    if order == "with_holes":
        e_up[::7] = -1.0
        e_up[1000:1100] = -1.0
    ok = e_up >= 0.0
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x[ok], y[ok], e[ok]), 2048, 2048)
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e_up)], 2048, 2048)
    assert fd.lrtb == ref_lrtb
    assert_map_match(fd.map, ref_map, what=f"2048^2 map, {order}", same_hits=True)
    assert fd.peak == pytest.approx(orc.fluence_peak(ref_map), rel=RTOL)
    assert fd.total_energy == pytest.approx(orc.fluence_total_energy(ref_map, ref_lrtb), rel=RTOL)


@pytest.mark.parametrize("nx,ny", [(100, 83), (600, 500)])
def test_binning_negative_taps_outside_an_explicit_range(ctx, nx, ny):
    """a hit left of / below an explicit range start has a negative fractional index: the reference (rays_hit_map.rs:352-376) adds
    the resulting NEGATIVE taps to column / row 0 and 1.  Both Binning kernels (fixed-point shared-memory splat for the small map,
    run-merging global splat for the large one) must add them too."""
    x, y, e = seeded_hits(30_000, spread=mm(1.0))
    left, bottom = float(np.quantile(x, 0.02)), float(np.quantile(y, 0.03))  # 2-3 % of the hits lie outside
    ranges = (left, float(x.max()), float(y.max()), bottom)
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), nx, ny, ranges)
    assert (ref_map < 0).any()
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e)], nx, ny, ranges=ranges)
    assert fd.lrtb == ref_lrtb
    scale = np.abs(ref_map).max()
    assert np.max(np.abs(fd.map - ref_map)) <= 1e-9 * scale  # sums of mixed sign: compare against the map's scale
    assert ((fd.map < 0) == (ref_map < 0)).mean() > 0.999


@pytest.mark.parametrize("nx,ny", [(101, 101), (400, 300)])
def test_binning_of_a_single_hit_follows_the_reference_arithmetic(ctx, nx, ny):
    """A hit map with ONE hit (a ghost bundle of which one ray reaches a detector): the bounding box has no extent, the
    fractional indices are 0 / 0 = NaN, `NaN as usize` is 0 (rays_hit_map.rs:352-357) and four bins around (0, 0) receive NaN;
    the peak over the finite bins is 0 (fluence_data.rs:45-68).  No error: the ghost-focus check records a peak of 0."""
    x, y, e = np.array([mm(1.5)]), np.array([mm(-0.5)]), np.array([1e-6])
    ref_map, ref_lrtb = orc.fluence_binning(orc.hits_from_xye(x, y, e), nx, ny)
    assert np.isnan(ref_map).sum() == 4 and orc.fluence_peak(ref_map) == 0.0
    fd = ob.fluence_binning(ctx, [hitmap_from(ctx, x, y, e)], nx, ny)
    assert fd.lrtb == ref_lrtb
    assert np.array_equal(np.isnan(fd.map), np.isnan(ref_map))
    assert np.array_equal(fd.map[~np.isnan(fd.map)], ref_map[~np.isnan(ref_map)])
    assert fd.peak == 0.0
