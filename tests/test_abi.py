"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/opossum_b200.h declares; enum values shared with the oracle agree; the host isometry math agrees
with the oracle's nalgebra restatement; without a GPU the library refuses loudly (no CPU fallback)."""
import ctypes as C
import math
import re
from pathlib import Path

import numpy as np
import pytest

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "opossum_b200.h").read_text()
    declared = set(re.findall(r"\b(opm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(F.EXPORTS), declared ^ set(F.EXPORTS)
    L = F.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.opm_abi_version() == 1


def test_struct_layouts():
    assert C.sizeof(F.OpmRay) == 96 == ob.RAY_DTYPE.itemsize == orc.RAY_DTYPE.itemsize
    assert C.sizeof(F.OpmIsometry) == 192 and C.sizeof(F.OpmSurface) == 216 and C.sizeof(F.OpmIndexModel) == 72
    assert C.sizeof(F.OpmOp) == 384 and C.sizeof(F.OpmTraceStats) == 72


def test_ctypes_mirrors_have_the_sizes_the_library_was_compiled_with():
    """every structure of include/*.h that the Python binding mirrors: sizeof on the C side == ctypes.sizeof"""
    L = F.lib()
    names = ["OpmRay", "OpmIso", "OpmIsometry", "OpmSurface", "OpmIndexModel", "OpmApertureItem", "OpmAperture", "OpmTraceStats",
             "OpmRefractOp", "OpmDiffractOp", "OpmFilterOp", "OpmSplitOp", "OpmForkOp", "OpmOp", "OpmSpotStats", "OpmSpotPartial",
             "OpmSourceDesc", "OpmRayTraceConfig", "OpmWavefrontStats", "OpmNodeDesc", "OpmCoatingDesc", "OpmGhostFocusConfig",
             "OpmGhostFocusStats", "OpmCriticalFluence"]
    for name in names:
        c_size = L.opm_abi_sizeof(name.encode())
        assert c_size > 0, name
        assert hasattr(F, name), f"no ctypes mirror of {name}"
        assert C.sizeof(getattr(F, name)) == c_size, f"{name}: ctypes {C.sizeof(getattr(F, name))} vs C {c_size}"
    assert L.opm_abi_sizeof(b"NoSuchStruct") == 0


def test_enums_match_oracle():
    assert (F.GEOM_PLANE, F.GEOM_SPHERE, F.GEOM_CYLINDER, F.GEOM_PARABOLA) == (
        orc.GEOM_PLANE, orc.GEOM_SPHERE, orc.GEOM_CYLINDER, orc.GEOM_PARABOLA)
    assert (F.COAT_IDEAL_AR, F.COAT_CONSTANT_R, F.COAT_FRESNEL) == (orc.COAT_IDEAL_AR, orc.COAT_CONSTANT_R, orc.COAT_FRESNEL)
    assert (F.MISS_STOP, F.MISS_IGNORE) == (orc.MISS_STOP, orc.MISS_IGNORE)


def test_no_gpu_fails_loudly():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(ob.OpossumError) as ei:
        ob.Context(0)
    assert ei.value.code == F.OPM_ERR_CUDA and "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("t,e", [((0, 0, 0), (0, 0, 0)), ((0.0, 0.0, 0.01), (0.0, 0.0, 0.0)),
                                 ((0.1, -0.2, 0.3), (0.3, -0.7, 1.1)), ((0, 0.05, 1.0), (math.pi / 4, 0, 0)),
                                 ((1, 2, 3), (0.0, math.radians(90.0), 0.0))])
def test_isometry_matches_oracle(t, e):
    """3x3 form (product) vs unit-quaternion form (oracle, nalgebra) agree to 1e-15; identity rotations exactly."""
    a, b = orc.Isometry.new(t, e), ob.Isometry.new(t, e)
    for x, y in zip(a.matrices(), b.matrices()):
        assert np.allclose(x, y, rtol=0, atol=4e-16)
    if e == (0, 0, 0) or e == (0.0, 0.0, 0.0):
        for x, y in zip(a.matrices(), b.matrices()):
            assert np.array_equal(x, y)
    t2, e2 = (0.0, 0.0, 0.2), (0.05, 0.0, -0.2)
    a2, b2 = a.append(orc.Isometry.new(t2, e2)), b.append(ob.Isometry.new(t2, e2))
    for x, y in zip(a2.matrices(), b2.matrices()):
        assert np.allclose(x, y, rtol=0, atol=1e-15)


def test_default_config():
    c = F.OpmRayTraceConfig()
    F.lib().opm_raytrace_config_default(C.byref(c))
    assert (c.min_energy_per_ray, c.max_number_of_bounces, c.max_number_of_refractions, c.missed_surface_strategy) == (
        1e-12, 1000, 1000, F.MISS_STOP)
    assert ob.RayTraceConfig() == ob.RayTraceConfig(1e-12, 1000, 1000, F.MISS_STOP)


def test_isometry_rejects_non_finite():
    with pytest.raises(ob.OpossumError):
        ob.Isometry.new((math.nan, 0, 0))
    with pytest.raises(ob.OpossumError):
        ob.Isometry.new((0, 0, 0), (0, math.inf, 0))


def test_spectrum_gaussian_host_helper_matches_the_oracle():
    """opm_spectrum_gaussian (spectral_distribution/gaussian.rs:78-96) is host math: same wavelengths and weights as the oracle"""
    import numpy as np

    import opossum_b200 as ob
    from oracle import oracle as orc
    for args in ((950e-9, 1150e-9, 64, 1053e-9, 50e-9, 1.0), (400e-9, 800e-9, 7, 600e-9, 120e-9, 2.5), (1e-6, 1e-6, 1, 1e-6, 1e-8, 1.0)):
        wv, w = ob.spectrum_gaussian(*args)
        rv, rw = orc.spectrum_gaussian(*args)
        assert np.array_equal(wv, rv)
        assert np.allclose(w, rw, rtol=1e-14, atol=0) and abs(w.sum() - 1.0) < 1e-14


def test_jit_source_compiles_without_a_gpu():
    """the run-time specialised kernel (opm_jit.cu): a program with every op kind, geometry and coating is generated and
    compiled by NVRTC for sm_100a against the embedded device headers -- the build check of the JIT path"""
    import ctypes as C
    buf = C.create_string_buffer(4096)
    n = F.lib().opm_jit_selftest(buf, 4096)
    assert n > 10_000, buf.value.decode()


def test_parabola_off_axis_isometry_host_helper():
    """opm_parabola_off_axis_isometry (pure host math, no device): the reference's KAT nodes/parabolic_mirror.rs:901-951 and the
    oracle's restatement"""
    import math

    import numpy as np

    import opossum_b200 as ob
    from oracle import scene as OS
    iso, pf = ob.Isometry.parabola_off_axis(1.0, math.radians(45.0), (0.0, 1.0), True)
    rot, t, _, _ = iso.matrices()
    want_rot = np.array([-0.0, 1.0, -0.0, -0.7071067811865475, -0.0, -0.7071067811865476, -0.7071067811865476, 0.0, 0.7071067811865475])
    want_t = np.array([-0.0, -0.6035533905932736, -0.3964466094067264])
    eps = np.finfo(float).eps
    assert np.all(np.abs(rot - want_rot) <= 3 * eps) and np.all(np.abs(t - want_t) <= 3 * eps)
    assert ob.Isometry.parabola_off_axis(1.0, math.radians(90.0), (0.0, 1.0), True)[1] == 0.5000000000000001
    for args in [(0.25, math.radians(30.0), (0.35, 8.35), False), (2.0, 0.0, (1.0, 0.0), True), (0.5, math.radians(-60.0), (-1.0, 0.3), True)]:
        a, pfa = ob.Isometry.parabola_off_axis(*args)
        b, pfb = OS.parabola_off_axis_isometry(args[0], args[1], args[2][0], args[2][1], args[3])
        assert pfa == pfb and all(np.allclose(x, y, rtol=0, atol=1e-15) for x, y in zip(a.matrices(), b.matrices()))
    # ParabolicMirror::check_attributes (nodes/parabolic_mirror.rs:231-259, tests :594-700)
    def refused(*a):
        try:
            ob.Isometry.parabola_off_axis(*a)
        except ob.OpossumError:
            return True
        return False
    assert all(refused(bad, 0.1) for bad in (0.0, math.inf, math.nan))
    assert all(refused(1.0, math.radians(a)) for a in (180.0, 190.0, -180.0, -190.0, math.nan, math.inf, -math.inf))
    assert not any(refused(1.0, math.radians(a)) for a in (45.0, -45.0, 0.0, 179.0))
    assert refused(1.0, 0.5, (math.nan, 1.0)) and refused(1.0, 0.5, (1e-17, 0.0)) and not refused(1.0, 0.5, (-1.0, -1.0))


def test_isometry_host_helpers_reference_kats():
    """opm_isometry_new / _append (pure host math): utils/geom_transformation.rs:680-800 -- non-finite translations and angles are
    errors, new_along_z, its inverse, append_z, append_with_rot, the Euler angles of Isometry::new(.., (10, 20, 30) deg)"""
    import math

    import numpy as np

    import opossum_b200 as ob
    for bad in (math.nan, math.inf, -math.inf):
        for k in range(3):
            v = [0.0, 0.0, 0.0]
            v[k] = bad
            for args in ((v, (0.0, 0.0, 0.0)), ((0.0, 0.0, 0.0), v)):
                try:
                    ob.Isometry.new(*args)
                    raise AssertionError("non-finite isometry accepted")
                except ob.OpossumError:
                    pass
    rot, t, rot_inv, t_inv = ob.Isometry.new_along_z(0.01).matrices()
    assert np.array_equal(rot, np.eye(3).ravel()) and tuple(t) == (0.0, 0.0, 0.01) and tuple(t_inv) == (0.0, 0.0, -0.01)
    assert tuple(ob.Isometry.new_along_z(0.01).append(ob.Isometry.new_along_z(0.02)).matrices()[1]) == (0.0, 0.0, 0.03)
    r = ob.Isometry.new((0.0, 0.0, 0.01), (0.0, math.radians(90.0), 0.0)).append(ob.Isometry.new_along_z(0.02))
    assert np.all(np.abs(r.matrices()[1] - (0.02, 0.0, 0.01)) <= np.finfo(float).eps)  # image of the origin
    m = ob.Isometry.new((0.0, 0.0, 0.0), tuple(math.radians(a) for a in (10.0, 20.0, 30.0))).matrices()[0].reshape(3, 3)
    got = np.array([math.atan2(m[2, 1], m[2, 2]), -math.asin(m[2, 0]), math.atan2(m[1, 0], m[0, 0])])
    assert np.all(np.abs(got - np.radians([10.0, 20.0, 30.0])) <= np.finfo(float).eps)
