"""Edge cases of the GPU path against the oracle: empty and ragged bundles, a single ray, bundles without a valid ray,
programs longer than one launch, detectors that nothing hits, zero-energy rays, and the interpreter / run-time
specialised kernels against each other."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import RTOL, assert_rays_match, index_pair, iso_pair, make_bundle, mm, nm, node_sphere_pair, surface_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def lens_program(hm=None):
    node = iso_pair((0.0, 0.0, mm(30.0)))
    front = node_sphere_pair(mm(120.0), node, coating=F.COAT_FRESNEL)
    rear = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(34.0))), coating=F.COAT_CONSTANT_R, r=0.02)
    glass, vac = index_pair("sellmeier1", 6.14555251E-1, 6.56775017E-1, 1.02699346E+0, 1.45987884E-2, 2.87769588E-3,
                            1.07653051E+2, nm(300.0), nm(2000.0)), index_pair("const", 1.0)

    def oracle(rays):
        orc.rays_refract_on_surface(rays, front[0], glass[0])
        orc.rays_threshold(rays, 1e-12)
        _, hits, _ = orc.rays_refract_on_surface(rays, rear[0], vac[0])
        orc.rays_threshold(rays, 1e-12)
        return hits
    prog = (ob.Program().refract(front[1], glass[1]).threshold(1e-12).refract(rear[1], vac[1], hit_map=hm).threshold(1e-12))
    return oracle, prog


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 256, 257, 1000])
def test_ragged_bundle_sizes(ctx, n):
    """tile boundaries of the persistent kernel (256-ray tiles, 32-lane warps), the single ray and the empty bundle"""
    rays = make_bundle(max(n, 1), wvl=(nm(700.0), nm(1400.0)))[:n]
    hm = ob.HitMap(ctx, max(n, 1))
    oracle, prog = lens_program(hm)
    ref = rays.copy()
    hits = oracle(ref)
    r = ob.Rays.from_array(ctx, rays)
    assert len(r) == n
    st = prog.run(r)
    assert st.interactions == 2 * n and st.alive_out == n and st.hits == len(hits)
    assert_rays_match(r.to_array(), ref, what=f"n={n}")
    if n == 0:
        assert r.total_energy() == 0.0 and r.centroid() is None


def test_bundle_without_valid_rays_and_zero_energy_rays(ctx):
    """invalid rays are carried along untouched (rays.rs:987: only valid rays are refracted); zero-energy rays stay valid
    and still spawn (zero-energy) children unless a positive threshold removes them"""
    rays = make_bundle(500)
    rays["valid"][:] = 0
    oracle, prog = lens_program()
    r = ob.Rays.from_array(ctx, rays)
    st = prog.run(r)
    assert st.interactions == 0 and st.valid_out == 0 and st.alive_out == 500
    got = r.to_array()
    assert np.array_equal(got["pos"], rays["pos"]) and np.array_equal(got["opl"], rays["opl"])
    zero = make_bundle(300)
    zero["e"][::2] = 0.0
    surf = surface_pair(F.GEOM_PLANE, iso=iso_pair((0.0, 0.0, mm(20.0))), coating=F.COAT_CONSTANT_R, r=0.1)
    ref = zero.copy()
    ch_ref, _, _ = orc.rays_refract_on_surface(ref, surf[0], index_pair("const", 1.5)[0])
    g = ob.Rays.from_array(ctx, zero)
    ch, st = g.refract_on_surface(surf[1], index_pair("const", 1.5)[1])
    assert st.children == len(ch_ref) == 300  # the zero-energy half reflects zero-energy children
    assert_rays_match(ch.to_array(), ch_ref, keyed=True, what="children incl. zero energy")
    assert_rays_match(g.to_array(), ref, what="parents incl. zero energy")
    g.invalidate_by_threshold_energy(1e-15)
    orc.rays_threshold(ref, 1e-15)
    assert g.nr_of_rays() == orc.nr_of_rays(ref) == 150


def test_program_longer_than_one_launch(ctx):
    """a chain whose op list exceeds OPM_MAX_PROGRAM_OPS is split into several launches by the scene layer"""
    MM = scenes.MM
    nodes = [scenes.node("source")]
    for k in range(30):  # 30 dummies -> 30 x (refract, 2 apodize, thresholds, limits): far more than 64 ops
        # radii chosen off the source's own radii 9 mm * sqrt(i / 4000): a ray exactly ON an aperture edge is decided by the
        # last bit of its position (the `<=` of aperture.rs:158-180) and is not a parity case
        nodes.append(scenes.node("dummy", z=(10 + 5 * k) * MM, ap_in=("circle", (9.03 - 0.1 * k) * MM, 0.0, 0.0)))
    nodes.append(scenes.node("fluence_detector", z=200 * MM))
    src = {"pos": ("fibonacci_ellipse", 4000, 9 * MM, 9 * MM), "energy": ("uniform", 1.0), "wvl": [1000 * scenes.NM], "w": [1.0]}
    osc = scenes.to_oracle(nodes)
    ref = osc.raytrace(scenes.source_to_oracle(src))
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    l0 = ctx.launch_count
    st = psc.raytrace(rays)
    assert st.interactions == osc.interactions and st.valid_out == orc.nr_of_rays(ref) and 0 < st.valid_out < 4000
    assert ctx.launch_count - l0 >= 4  # at least two trace launches (each with its program fetch)
    assert_rays_match(rays.to_array(), ref, what="long chain")


def test_detector_that_nothing_hits(ctx):
    """every ray is clipped before the detector: Binning raises the reference's 'could not calculate bounding box'
    (rays_hit_map.rs:526-531), the spot statistics report no valid ray"""
    MM = scenes.MM
    nodes = [scenes.node("source"), scenes.node("dummy", z=10 * MM, ap_in=("circle", 1 * MM, 0.05, 0.0)),
             scenes.node("spot_diagram", z=20 * MM), scenes.node("fluence_detector", z=30 * MM)]
    src = {"pos": ("grid", 20, 20, 20 * MM, 20 * MM), "energy": ("uniform", 1.0), "wvl": [1000 * scenes.NM], "w": [1.0]}
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    st = psc.raytrace(rays)
    assert st.valid_out == 0 and st.alive_out == 400
    with pytest.raises(ob.OpossumError) as e:
        psc.node_fluence(3)
    assert e.value.code == F.OPM_ERR_EMPTY_HITMAP
    s = psc.node_spot(2)
    assert s.n_valid == 0 and s.n_total == 400 and np.isnan(s.beam_radius_geo)


def test_interpreter_and_specialised_kernels_agree(ctx):
    """the same program through the interpreter kernel and through the kernel compiled for it: identical counts and
    flags, values equal to rounding (both are the contracted build; only instruction scheduling differs)"""
    rays = make_bundle(5000, divergence=0.05)
    outs = []
    for jit in (0, 2):
        ctx.set_jit(jit, 1)
        hm = ob.HitMap(ctx, len(rays))
        _, prog = lens_program(hm)
        r = ob.Rays.from_array(ctx, rays)
        st = prog.run(r)
        outs.append((r.to_array(), hm.download(), st))
    ctx.set_jit(1, 200000)
    (a, ha, sa), (b, hb, sb) = outs
    assert sa == sb
    assert_rays_match(b, a, rtol=1e-13, what="specialised vs interpreter")
    assert np.array_equal(ha[4], hb[4]) and np.allclose(ha[0], hb[0], rtol=1e-13, atol=1e-18)
    assert ctx.jit_stats()["launches"] >= 1 and ctx.jit_stats()["compile_failures"] == 0


def test_wavelength_hint_is_only_a_hint(ctx):
    """run-time specialised kernel with the once-per-CTA index evaluation (RefractCt HINT): a bundle of three wavelengths under a
    hint that names one of them, none of them, and one outside the glass's range -- rays of the hinted wavelength take the CTA's
    value, the others evaluate the Sellmeier model themselves; the results are the oracle's either way"""
    import scenes
    MM, NM = scenes.MM, scenes.NM
    nodes = [scenes.node("source"),
             scenes.node("lens", z=30 * MM, p=(120 * MM, -150 * MM, 6 * MM), index=scenes.HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),
             scenes.node("wedge", z=80 * MM, p=(8 * MM, math.radians(4.0)), index=scenes.HZF52),
             scenes.node("spot_diagram", z=200 * MM)]
    src = {"pos": ("fibonacci_ellipse", 4000, 5 * MM, 4 * MM), "energy": ("uniform", 1.0), "wvl": [532 * NM, 1054 * NM, 1550 * NM], "w": [0.2, 0.5, 0.3]}
    osc = scenes.to_oracle(nodes)
    ref = osc.raytrace(scenes.source_to_oracle(src))
    psc = scenes.to_product(ctx, nodes)
    ctx.set_jit(2, 1)
    try:
        for hint in (0.0, 1054 * NM, 800 * NM, 5000 * NM):
            rays = scenes.source_to_product(src).generate(ctx)
            rays.set_wavelength_hint(hint)
            st = psc.raytrace(rays)
            assert st.interactions == osc.interactions, hint
            assert_rays_match(rays.to_array(), ref, what=f"hint {hint}")
    finally:
        ctx.set_jit(1, 200000)
    assert ctx.jit_stats()["compile_failures"] == 0


def test_bounding_box_kept_by_the_trace_kernel_equals_the_pass_over_the_map(ctx):
    """the run-time specialised kernel keeps the bounding box (and the hit count) of a hit map it records; the read-out then
    skips its own pass over the map: both must give the same bits, also for a detector most rays miss and for one nothing hits;
    an upload into the map drops the kept box"""
    import scenes
    MM, NM = scenes.MM, scenes.NM
    nodes = [scenes.node("source"),
             scenes.node("lens", z=30 * MM, p=(120 * MM, -150 * MM, 6 * MM), index=scenes.HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),
             scenes.node("fluence_detector", z=60 * MM, ap_in=("circle", 1.5 * MM, 0.4 * MM, -0.2 * MM)),
             scenes.node("dummy", z=90 * MM, ap_in=("circle", 0.1 * MM, 50 * MM, 0.0)),  # an aperture far off the beam: no ray survives
             scenes.node("fluence_detector", z=120 * MM)]
    src = {"pos": ("fibonacci_ellipse", 5000, 5 * MM, 4 * MM), "energy": ("uniform", 1.0), "wvl": [1054 * NM], "w": [1.0]}
    out = {}
    for mode in ("interpreter", "jit"):
        ctx.set_jit(2 if mode == "jit" else 0, 1)
        try:
            psc = scenes.to_product(ctx, nodes)
            psc.raytrace(scenes.source_to_product(src).generate(ctx))
            boxes = []
            for node in (2, 4):
                hm = psc.node_hitmaps(node)
                try:
                    boxes.append(ob.hitmaps_bounding_box(hm))
                except ob.OpossumError as e:
                    boxes.append(("error", e.code))
            out[mode] = boxes
            if mode == "jit":  # a host upload replaces the map's content: the kept box must not survive it
                h = psc.node_hitmaps(2)[0]
                h.upload(np.array([1.0, 2.0]), np.array([-3.0, 4.0]), np.array([1.0, 1.0]))
                assert ob.hitmaps_bounding_box([h]) == ((1.0, 2.0, 4.0, -3.0), 2)
        finally:
            ctx.set_jit(1, 200000)
    assert out["jit"] == out["interpreter"]
    assert out["jit"][0][1] > 0 and out["jit"][1][0] == "error"


def test_peak_checks_in_a_batch_equal_the_checks_one_at_a_time_and_the_oracle(ctx):
    """opm_fluence_peak_check_batch_async (what the ghost-focus analysis defers its per-bundle checks to): five hit maps of
    different sizes and bounce mixes -- one without a hit of its level, one with a single hit, one empty -- through the batch, through
    opm_fluence_peak_check_async one at a time, and through the oracle's 101 x 101 Binning map of the hits of that level
    (RaysHitMap::get_max_fluence, rays_hit_map.rs:777-798)"""
    import ctypes as C
    rng = np.random.default_rng(77)
    N = 101
    pad = (N * N + 7) // 8 * 8
    sizes, levels = [50_000, 1234, 700, 1, 0], [1, 0, 3, 2, 0]
    maps, data = [], []
    for n, lvl in zip(sizes, levels):
        x, y = rng.normal(0.0, 2e-3, n), rng.normal(1e-3, 1e-3, n)
        e = rng.uniform(0.0, 1.0, n)
        b = rng.integers(0, 3, n).astype(np.uint32) if n > 1 else np.full(n, lvl, dtype=np.uint32)
        e[rng.random(n) < 0.1] = -1.0  # empty slots
        hm = ob.HitMap(ctx, max(n, 1))
        if n:
            hm.upload(x, y, e, bounce=b)
        maps.append(hm)
        data.append((x, y, e, b))
    # a sixth check sums TWO hit maps (the same bundle id met the surface twice): maps 0 and 1 at bounce level 1
    groups = [[i] for i in range(len(maps))] + [[0, 1]]
    levels = levels + [1]
    k = len(groups)
    acc = ctx.device_alloc(64 * k)
    assert F.lib().opm_fluence_peak_check_acc_init(ctx.h, C.c_void_p(acc), k) == 0
    d_maps = ctx.device_alloc(pad * 8 * k)
    d_res = ctx.device_alloc(64 * k)
    ctx.device_memset(d_res, 0, 64 * k)
    flat = [maps[i].h for g in groups for i in g]
    arr = (C.c_void_p * len(flat))(*flat)
    per = (C.c_int32 * k)(*[len(g) for g in groups])
    slots = (C.c_void_p * k)(*[d_res + 64 * i for i in range(k)])
    filt = (C.c_int64 * k)(*levels)
    assert F.lib().opm_fluence_peak_check_batch_async(arr, per, k, filt, N, N, C.c_void_p(acc), C.c_void_p(d_maps), slots) == 0
    batch = np.zeros((k, 8))
    ctx.to_host(d_res, batch)
    # one at a time
    acc1, map1, res1 = ctx.device_alloc(64), ctx.device_alloc(pad * 8), ctx.device_alloc(64)
    assert F.lib().opm_fluence_peak_check_acc_init(ctx.h, C.c_void_p(acc1), 1) == 0
    for i in range(k):
        ctx.device_memset(res1, 0, 64)
        one = (C.c_void_p * len(groups[i]))(*[maps[j].h for j in groups[i]])
        assert F.lib().opm_fluence_peak_check_async(one, len(groups[i]), levels[i], N, N, C.c_void_p(acc1), C.c_void_p(map1), C.c_void_p(res1)) == 0
        single = np.zeros(8)
        ctx.to_host(res1, single)
        assert np.array_equal(batch[i, :4], single[:4]), i                      # the range: bit for bit
        assert batch[i, 6:7].view(np.uint32)[0] == single[6:7].view(np.uint32)[0], i
        x, y, e, b = (np.concatenate([data[j][f] for j in groups[i]]) for f in range(4))
        sel = (e >= 0.0) & (b == levels[i]) if len(e) else np.zeros(0, dtype=bool)
        if not sel.any():
            assert np.all(np.isneginf(batch[i, :4])), i                         # get_rays_hit_map(bounce, uuid) == None
            continue
        assert batch[i, 4] == pytest.approx(single[4], rel=1e-9, nan_ok=True) and batch[i, 5] == pytest.approx(single[5], rel=1e-9, nan_ok=True), i
        hits = orc.hits_from_xye(x[sel], y[sel], e[sel])
        if sel.sum() == 1:
            continue  # a range of zero width: no map to compare (the reference's NaN bin index lands in bin 0)
        ref_map, ref_lrtb = orc.fluence_binning(hits, N, N)
        assert np.allclose([-batch[i, 0], batch[i, 1], batch[i, 2], -batch[i, 3]], ref_lrtb, rtol=1e-15, atol=0), i
        assert batch[i, 4] == pytest.approx(np.nanmax(ref_map), rel=1e-9), i
    for p in (acc, d_maps, d_res, acc1, map1, res1):
        ctx.device_free(p)
