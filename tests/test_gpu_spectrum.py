"""Spectrum-driven detectors and operators against the oracle: Rays::to_spectrum / Spectrometer, the
FilterType::Spectrum and SplittingConfig::Spectrum forms of filter_energy / split, and the WaveFront detector."""
import math

import numpy as np
import pytest

import opossum_b200 as ob
import scenes
from opossum_b200 import ffi as F
from oracle import oracle as orc
from helpers import RTOL, assert_rays_match, iso_pair, make_bundle, mm, nm

pytestmark = pytest.mark.gpu
UM = 1e-6


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def broadband_bundle(n=20000, seed=3):
    rng = np.random.default_rng(seed)
    rays = make_bundle(n, wvl=(nm(950.0), nm(1150.0)), seed=seed)
    rays["e"] = rng.uniform(0.0, 2.0, n)
    rays["wvl"][::7] = nm(1000.0) + np.round(rng.uniform(0, 500, len(rays["wvl"][::7]))) * nm(0.2)  # many rays exactly on a slot
    rays["valid"][::11] = 0
    rays["opl"] = rng.uniform(0.2, 0.2001, n)
    return rays


def assert_values_match(got, ref, what):
    assert got.shape == ref.shape, what
    assert np.allclose(got, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max()), f"{what}: max dev {np.abs(got - ref).max()} of {np.abs(ref).max()}"


@pytest.mark.parametrize("res", [nm(0.2), nm(1.0), nm(13.0)])
def test_to_spectrum_matches_oracle(ctx, res):
    rays = broadband_bundle()
    lam_ref, val_ref = orc.rays_to_spectrum(rays, res)
    r = ob.Rays.from_array(ctx, rays)
    lam, val = r.to_spectrum(res)
    assert np.array_equal(lam, lam_ref)  # the slot wavelengths are the same fma(i, step, start)
    assert_values_match(val, val_ref, "spectrum values")
    # host helpers (Spectrum::total_energy / center_wavelength / get_value) against the oracle's on the same arrays
    assert ob.spectrum_total_energy(lam_ref, val_ref) == orc.spectrum_total_energy(lam_ref, val_ref)
    assert ob.spectrum_center_wavelength(lam_ref, val_ref) == orc.spectrum_center_wavelength(lam_ref, val_ref)
    if res <= nm(1.0):  # (a coarse grid can end before max + resolution: Spectrum::total_energy leaves the last slot out)
        assert abs(ob.spectrum_total_energy(lam, val) - orc.total_energy(rays)) < 1e-9 * orc.total_energy(rays)
    for w in (nm(900.0), lam_ref[0] * UM, nm(1000.1), nm(1049.93), lam_ref[-1] * UM, nm(1200.0)):
        assert ob.spectrum_get_value(lam_ref, val_ref, w) == orc.spectrum_get_value(lam_ref, val_ref, w)
    mn, mx, nv = r.wavelength_range()
    ok = rays["valid"] == 1
    assert (mn, mx, nv) == (rays["wvl"][ok].min(), rays["wvl"][ok].max(), int(ok.sum()))


def test_to_spectrum_of_shards_adds_up_and_edge_cases(ctx):
    """multi-GPU form: every shard bins on the common wavelength range, the values add"""
    rays = broadband_bundle(9000, seed=5)
    lam_ref, val_ref = orc.rays_to_spectrum(rays, nm(0.5))
    parts = [ob.Rays.from_array(ctx, rays[k::3]) for k in range(3)]
    rng = [p.wavelength_range() for p in parts]
    common = (min(r[0] for r in rng), max(r[1] for r in rng))
    total = None
    for p in parts:
        lam, val = p.to_spectrum(nm(0.5), wavelength_range=common)
        assert np.array_equal(lam, lam_ref)
        total = val if total is None else total + val
    assert_values_match(total, val_ref, "sum of the shards")
    # no valid ray: "ray bundle is empty - cannot create spectrum" (rays.rs:1216-1220)
    dead = rays[:50].copy()
    dead["valid"] = 0
    with pytest.raises(ob.OpossumError) as e:
        ob.Rays.from_array(ctx, dead).to_spectrum(nm(1.0))
    assert e.value.code == F.OPM_ERR_SPECTRUM
    with pytest.raises(orc.OracleError):
        orc.rays_to_spectrum(dead, nm(1.0))
    with pytest.raises(ob.OpossumError):
        parts[0].to_spectrum(-nm(1.0))
    # a single wavelength (laser line): 2 slots, everything in the first (spectrum.rs:726-734)
    line = make_bundle(100, wvl=(nm(1053.0), nm(1053.0)))
    lam, val = ob.Rays.from_array(ctx, line).to_spectrum(nm(1.0))
    lr, vr = orc.rays_to_spectrum(line, nm(1.0))
    assert np.array_equal(lam, lr) and len(lam) == 2 and val[1] == 0.0
    assert_values_match(val, vr, "laser line")


def filter_curve():
    lam = np.array([0.90, 0.98, 1.00, 1.05, 1.051, 1.12, 1.20])  # um, non-equidistant
    val = np.array([0.05, 0.20, 0.95, 0.90, 0.10, 0.50, 1.00])
    return lam, val


@pytest.mark.parametrize("jit", [0, 2])
def test_spectral_filter_and_split_match_oracle(ctx, jit):
    """FilterType::Spectrum (ray.rs:712-721) and SplittingConfig::Spectrum (ray.rs:752-772), stand-alone calls and inside a
    fused program (interpreter and run-time specialised kernel), invalid rays copied unchanged"""
    rays = broadband_bundle(6000, seed=9)
    lam, val = filter_curve()
    if (lam[-1] * UM) / 1e-6 == lam[-1]:
        rays["wvl"][5] = lam[-1] * UM   # exactly the last slot: Some(last value) (spectrum.rs:323-327)
    rays["wvl"][6] = lam[0] * UM        # exactly the first slot
    sp = ob.Spectrum(ctx, lam, val)
    ref = rays.copy()
    orc.rays_filter_energy_spectrum(ref, lam, val)
    ref_split = orc.rays_split_spectrum(ref, lam, val)
    ctx.set_jit(jit, 1)
    try:
        g = ob.Rays.from_array(ctx, rays)
        g.filter_energy_spectrum(sp)
        split = g.split_spectrum(sp)
        assert_rays_match(g.to_array(), ref, what="filter + split, main")
        assert_rays_match(split.to_array(), ref_split, what="filter + split, split-off")
        g2, out2 = ob.Rays.from_array(ctx, rays), ob.Rays(ctx, len(rays))
        ob.Program().filter(spectrum=sp).split(0.0, out2, spectrum=sp).run(g2)
        assert_rays_match(g2.to_array(), ref, what="program, main")
        assert_rays_match(out2.to_array(), ref_split, what="program, split-off")
        g3, out3 = ob.Rays.from_array(ctx, rays), ob.Rays(ctx, len(rays))
        ob.Program().filter(spectrum=sp).fork(0.0, spectrum=sp).threshold(0.0).join(out3).run(g3)
        assert_rays_match(g3.to_array(), ref, what="fork/join, main")
        assert_rays_match(out3.to_array(), ref_split, what="fork/join, split-off")
        # a wavelength outside the spectrum fails the call (ray.rs:716-720, :757-761)
        bad = rays.copy()
        bad["wvl"][17] = 1.25 * UM
        for fn in (lambda r: r.filter_energy_spectrum(sp), lambda r: r.split_spectrum(sp)):
            with pytest.raises(ob.OpossumError) as e:
                fn(ob.Rays.from_array(ctx, bad))
            assert e.value.code == F.OPM_ERR_SPECTRUM
        with pytest.raises(orc.OracleError):
            orc.rays_filter_energy_spectrum(bad.copy(), lam, val)
        # a spectrum that is no transmission spectrum: ratio outside [0, 1] (ray.rs:763-767)
        sp_bad = ob.Spectrum(ctx, lam, val * 3.0)
        with pytest.raises(ob.OpossumError) as e:
            ob.Rays.from_array(ctx, rays).split_spectrum(sp_bad)
        assert e.value.code == F.OPM_ERR_SPLIT_RATIO
    finally:
        ctx.set_jit(1, 200000)


def test_wavefront_kernels_match_oracle_on_the_same_rays(ctx):
    """Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802) + calc_wavefront_statistics (nodes/wavefront.rs:145-164)
    on identical inputs: rows in bundle order, reference ray = closest to the monitor axis (first on ties)"""
    rays = broadband_bundle(15000, seed=21)
    rays["wvl"] = nm(1053.0)
    mon = iso_pair((mm(0.3), mm(-0.2), mm(5.0)), (0.01, -0.02, 0.3))
    rays["pos"][10] = rays["pos"][4000] = (mm(0.3), mm(-0.2), mm(5.0))  # two rays at the monitor origin, different path lengths: the first wins
    rays["valid"][10] = rays["valid"][4000] = 1
    for wvl in (nm(1053.0), 0.0):
        w_ref = wvl if wvl > 0 else orc.spectrum_center_wavelength(*orc.rays_to_spectrum(rays, nm(1.0)))
        ref = orc.rays_wavefront_error(rays, w_ref, mon[0])
        ptv, rms = orc.wavefront_statistics(ref)
        st, xyw = ob.Rays.from_array(ctx, rays).wavefront_error(mon[1], wvl)
        assert st.n_rays == len(ref) == orc.nr_of_rays(rays)
        assert abs(st.wavelength_m - w_ref) <= 1e-12 * w_ref
        assert np.allclose(xyw[:, :2], ref[:, :2], rtol=1e-12, atol=1e-12 * np.abs(ref[:, :2]).max())
        assert np.allclose(xyw[:, 2], ref[:, 2], rtol=0, atol=1e-12 * np.abs(ref[:, 2]).max() + 1e-9 * abs(st.wavelength_m - w_ref) / w_ref)
        assert abs(st.ptv - ptv) <= 1e-12 * ptv and abs(st.rms - rms) <= 1e-10 * rms
    # on-axis tie: the reference ray is ray 10 (its own error is exactly 0), not ray 4000
    ok = np.flatnonzero(rays["valid"] == 1)
    assert xyw[np.searchsorted(ok, 10), 2] == 0.0 and xyw[np.searchsorted(ok, 4000), 2] != 0.0
    dead = rays[:20].copy()
    dead["valid"] = 0
    with pytest.raises(ob.OpossumError) as e:
        ob.Rays.from_array(ctx, dead).wavefront_error(mon[1], nm(1053.0))
    assert e.value.code == F.OPM_ERR_EMPTY_WAVEFRONT


def spectral_scene():
    """broadband source -> dichroic splitter (spectrum) -> lens -> spectral filter -> spectrometer -> wavefront monitor;
    second splitter output -> spectrometer"""
    MM = scenes.MM
    lam, val = filter_curve()
    edge = (np.array([0.90, 1.04, 1.06, 1.20]), np.array([0.02, 0.05, 0.95, 0.98]))
    return [
        scenes.node("source"),                                                                                   # 0
        scenes.node("beam_splitter", z=20 * MM, spectrum=edge),                                                  # 1
        scenes.node("lens", z=50 * MM, p=(300 * MM, -300 * MM, 5 * MM), index=scenes.HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),  # 2
        scenes.node("ideal_filter", z=80 * MM, spectrum=(lam, val)),                                             # 3
        scenes.node("spectrometer", z=100 * MM),                                                                 # 4
        scenes.node("wavefront", z=120 * MM),                                                                    # 5
        scenes.node("spectrometer", z=60 * MM, parent=1, parent_port=1),                                         # 6
    ]


def test_spectral_scene_matches_oracle(ctx):
    nodes = spectral_scene()
    src = scenes.c5_source(3000, 24)
    osc = scenes.to_oracle(nodes)
    ref_out = osc.raytrace(scenes.source_to_oracle(src))
    psc = scenes.to_product(ctx, nodes)
    rays = scenes.source_to_product(src).generate(ctx)
    st = psc.raytrace(rays)
    assert st.interactions == osc.interactions and st.valid_out == orc.nr_of_rays(ref_out)
    assert_rays_match(rays.to_array(), ref_out, what="main arm")
    for node in (4, 6):
        lam_ref, val_ref = osc.node_spectrum(node)
        lam, val = psc.node_spectrum(node)
        assert np.array_equal(lam, lam_ref)
        assert_values_match(val, val_ref, f"spectrometer node {node}")
    # the dichroic splitter sends the red half to the main arm, the blue half to the second output
    c_main = ob.spectrum_center_wavelength(*psc.node_spectrum(4))
    c_side = ob.spectrum_center_wavelength(*psc.node_spectrum(6))
    assert c_main > nm(1060.0) > nm(1040.0) > c_side
    ref = osc.node_wavefront(5)
    wst, xyw = psc.node_wavefront(5)
    assert wst.n_rays == len(ref["map"]) and abs(wst.wavelength_m - ref["wavelength"]) <= 1e-9 * ref["wavelength"]
    # the wavefront error is a difference of path lengths: the 1e-9 bar on the OPL (here ~0.12 m) bounds it to
    # 1e-9 * OPL / wavelength wavelengths
    opl_waves = np.abs(osc.nodes[5].stored["opl"]).max() / ref["wavelength"]
    assert np.allclose(xyw[:, :2], ref["map"][:, :2], rtol=RTOL, atol=RTOL * np.abs(ref["map"][:, :2]).max())
    assert np.abs(xyw[:, 2] - ref["map"][:, 2]).max() <= RTOL * opl_waves
    assert abs(wst.ptv - ref["ptv"]) <= 2 * RTOL * opl_waves and abs(wst.rms - ref["rms"]) <= 2 * RTOL * opl_waves
    assert wst.ptv > 1.0  # a lens in front of the monitor: many waves of defocus
