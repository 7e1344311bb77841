"""ctypes binding of the CPU ORACLE (oracle/libopm_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package (opossum_b200) never
imports this module.

The arithmetic lives in opm_oracle.c (a C restatement of the reference, every function citing
the reference file:line).  This file only marshals arrays and mirrors the reference's *control
flow* that is O(#nodes): the node -> surface builders, `pass_through_surface`, the ray-trace
walk and the ghost-focus pass loop (see oracle/scene.py).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libopm_oracle.so"


def build(force: bool = False) -> Path:
    """Compile oracle/libopm_oracle.so with the committed Makefile (gcc only)."""
    srcs = [_HERE / "opm_oracle.c", _HERE / "opm_voronoi.c", _HERE / "opm_spectrum.c", _HERE / "opm_oracle.h"]
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(f.stat().st_mtime for f in srcs):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


# ---------------------------------------------------------------------------
# layouts
# ---------------------------------------------------------------------------
RAY_DTYPE = np.dtype(
    [("pos", "<f8", 3), ("dir", "<f8", 3), ("e", "<f8"), ("wvl", "<f8"), ("opl", "<f8"), ("n", "<f8"),
     ("bounces", "<u4"), ("refr", "<u4"), ("valid", "<u4"), ("id", "<u4")], align=True)
assert RAY_DTYPE.itemsize == 96
HIT_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("e", "<f8"), ("bounce", "<u4"), ("id", "<u4")],
                     align=True)
assert HIT_DTYPE.itemsize == 40


class Iso3(C.Structure):
    _fields_ = [("q", C.c_double * 4), ("t", C.c_double * 3)]


class Isometry(C.Structure):
    """Mirror of utils/geom_transformation.rs `Isometry` (transform + cached inverse)."""
    _fields_ = [("fwd", Iso3), ("inv", Iso3)]

    @staticmethod
    def identity() -> "Isometry":
        o = Isometry()
        lib().orc_iso_identity(C.byref(o))
        return o

    @staticmethod
    def new(translation, euler_xyz=(0.0, 0.0, 0.0)) -> "Isometry":
        """Isometry::new(translation [m], axes_angles [rad]) geom_transformation.rs:44-72."""
        o = Isometry()
        lib().orc_iso_new((C.c_double * 3)(*translation), (C.c_double * 3)(*euler_xyz), C.byref(o))
        return o

    @staticmethod
    def new_along_z(z: float) -> "Isometry":
        return Isometry.new((0.0, 0.0, z), (0.0, 0.0, 0.0))

    @staticmethod
    def from_view(p, direction, up) -> "Isometry":
        o = Isometry()
        lib().orc_iso_from_view((C.c_double * 3)(*p), (C.c_double * 3)(*direction), (C.c_double * 3)(*up), C.byref(o))
        return o

    @staticmethod
    def new_axis_angle(translation, axisangle) -> "Isometry":
        """nalgebra Isometry3::new(translation, axisangle) wrapped by Isometry::new_from_transform"""
        o = Isometry()
        lib().orc_iso_new_axis_angle((C.c_double * 3)(*translation), (C.c_double * 3)(*axisangle), C.byref(o))
        return o

    def append(self, rhs: "Isometry") -> "Isometry":
        o = Isometry()
        lib().orc_iso_append(C.byref(self), C.byref(rhs), C.byref(o))
        return o

    def _apply(self, fn, v):
        out = (C.c_double * 3)()
        fn(C.byref(self), (C.c_double * 3)(*v), out)
        return np.array(out[:])

    def transform_point(self, p):
        return self._apply(lib().orc_iso_transform_point, p)

    def inverse_transform_point(self, p):
        return self._apply(lib().orc_iso_inverse_transform_point, p)

    def transform_vector(self, v):
        return self._apply(lib().orc_iso_transform_vector, v)

    def inverse_transform_vector(self, v):
        return self._apply(lib().orc_iso_inverse_transform_vector, v)

    def matrices(self):
        """(rot_fwd[9], t_fwd[3], rot_inv[9], t_inv[3]) -- for cross-checking the product's 3x3 form."""
        rf, tf, ri, ti = (C.c_double * 9)(), (C.c_double * 3)(), (C.c_double * 9)(), (C.c_double * 3)()
        lib().orc_iso_to_matrix(C.byref(self.fwd), rf, tf)
        lib().orc_iso_to_matrix(C.byref(self.inv), ri, ti)
        return np.array(rf[:]), np.array(tf[:]), np.array(ri[:]), np.array(ti[:])


GEOM_PLANE, GEOM_SPHERE, GEOM_CYLINDER, GEOM_PARABOLA = 0, 1, 2, 3
COAT_IDEAL_AR, COAT_CONSTANT_R, COAT_FRESNEL = 0, 1, 2
IDX_CONST, IDX_SELLMEIER1, IDX_SCHOTT, IDX_CONRADY = 0, 1, 2, 3
MISS_STOP, MISS_IGNORE = 0, 1
AP_NONE, AP_CIRCLE, AP_RECT, AP_POLYGON, AP_GAUSSIAN, AP_STACK = 0, 1, 2, 3, 4, 5


class Surface(C.Structure):
    _fields_ = [("geom", C.c_int32), ("coating", C.c_int32), ("param", C.c_double), ("coating_r", C.c_double),
                ("iso", Isometry)]


class IndexModel(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double * 6), ("wvl_lo", C.c_double),
                ("wvl_hi", C.c_double)]

    @staticmethod
    def const(n: float) -> "IndexModel":
        m = IndexModel(type=IDX_CONST)
        m.c[0] = n
        return m

    @staticmethod
    def sellmeier1(k1, k2, k3, l1, l2, l3, lo, hi) -> "IndexModel":
        m = IndexModel(type=IDX_SELLMEIER1, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [k1, k2, k3, l1, l2, l3]
        return m

    @staticmethod
    def schott(a0, a1, a2, a3, a4, a5, lo, hi) -> "IndexModel":
        m = IndexModel(type=IDX_SCHOTT, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [a0, a1, a2, a3, a4, a5]
        return m

    @staticmethod
    def conrady(n0, a, b, lo, hi) -> "IndexModel":
        m = IndexModel(type=IDX_CONRADY, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [n0, a, b, 0.0, 0.0, 0.0]
        return m

    def get_refractive_index(self, wvl: float) -> float:
        out = C.c_double()
        _check(lib().orc_refractive_index(C.byref(self), wvl, C.byref(out)))
        return out.value


class Aperture(C.Structure):
    pass


Aperture._fields_ = [("type", C.c_int32), ("obstruction", C.c_int32), ("p", C.c_double * 4), ("n_sub", C.c_int32),
                     ("_pad", C.c_int32), ("tris", C.POINTER(C.c_double)), ("sub", C.POINTER(Aperture))]


def ap_none() -> Aperture:
    return Aperture(type=AP_NONE)


def ap_circle(r, cx=0.0, cy=0.0, obstruction=False) -> Aperture:
    a = Aperture(type=AP_CIRCLE, obstruction=int(obstruction))
    a.p[:] = [r, cx, cy, 0.0]
    return a


def ap_rect(w, h, cx=0.0, cy=0.0, obstruction=False) -> Aperture:
    a = Aperture(type=AP_RECT, obstruction=int(obstruction))
    a.p[:] = [w, h, cx, cy]
    return a


def ap_gaussian(sx, sy, cx=0.0, cy=0.0, obstruction=False) -> Aperture:
    a = Aperture(type=AP_GAUSSIAN, obstruction=int(obstruction))
    a.p[:] = [sx, sy, cx, cy]
    return a


def ap_polygon(triangles, obstruction=False) -> Aperture:
    """triangles: (k,3,2) array of already ear-cut triangles (the reference uses `earcutr`, construction only)."""
    t = np.ascontiguousarray(np.asarray(triangles, dtype=np.float64).reshape(-1, 6))
    a = Aperture(type=AP_POLYGON, obstruction=int(obstruction), n_sub=t.shape[0])
    a._keep = t
    a.tris = t.ctypes.data_as(C.POINTER(C.c_double))
    return a


def ap_stack(items, obstruction=False) -> Aperture:
    arr = (Aperture * len(items))(*items)
    a = Aperture(type=AP_STACK, obstruction=int(obstruction), n_sub=len(items))
    a._keep = (arr, items)
    a.sub = C.cast(arr, C.POINTER(Aperture))
    return a


class RefractStats(C.Structure):
    _fields_ = [("n_children", C.c_uint64), ("n_hits", C.c_uint64), ("n_interactions", C.c_uint64),
                ("rays_missed", C.c_int32), ("valid_rays_found", C.c_int32)]


class OracleError(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        super().__init__(lib().orc_strerror(code).decode())


def _check(rc: int):
    if rc != 0:
        raise OracleError(rc)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.orc_strerror.restype = C.c_char_p
    L.orc_strerror.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    for name in ("orc_iso_transform_point", "orc_iso_inverse_transform_point", "orc_iso_transform_vector",
                 "orc_iso_inverse_transform_vector"):
        getattr(L, name).argtypes = [C.POINTER(Isometry), dp, dp]
        getattr(L, name).restype = None
    L.orc_iso_identity.argtypes = [C.POINTER(Isometry)]
    L.orc_iso_new.argtypes = [dp, dp, C.POINTER(Isometry)]
    L.orc_iso_append.argtypes = [C.POINTER(Isometry)] * 3
    L.orc_iso_from_view.argtypes = [dp, dp, dp, C.POINTER(Isometry)]
    L.orc_iso_to_matrix.argtypes = [C.POINTER(Iso3), dp, dp]
    L.orc_find_roots_quadratic.argtypes = [C.c_double] * 3 + [dp]
    L.orc_find_roots_quadratic.restype = C.c_int
    L.orc_powi.argtypes = [C.c_double, C.c_int]
    L.orc_powi.restype = C.c_double
    L.orc_refractive_index.argtypes = [C.POINTER(IndexModel), C.c_double, dp]
    L.orc_fresnel_reflectivity.argtypes = [dp, dp, C.c_double, C.c_double]
    L.orc_fresnel_reflectivity.restype = C.c_double
    L.orc_apodization_factor.argtypes = [C.POINTER(Aperture), C.c_double, C.c_double]
    L.orc_apodization_factor.restype = C.c_double
    L.orc_intersect.argtypes = [C.POINTER(Surface), dp, dp, dp, dp]
    L.orc_intersect.restype = C.c_int
    L.orc_ray_new.argtypes = [dp, dp, C.c_double, C.c_double, vp]
    L.orc_ray_refract_on_surface.argtypes = [vp, C.POINTER(Surface), C.c_int, C.c_double, C.c_int, vp,
                                             C.POINTER(C.c_int), vp, C.POINTER(C.c_int)]
    L.orc_rays_refract_on_surface.argtypes = [vp, C.c_uint64, C.POINTER(Surface), C.POINTER(IndexModel), C.c_int,
                                              C.c_int, vp, vp, C.POINTER(RefractStats), C.c_int]
    L.orc_rays_apodize.argtypes = [vp, C.c_uint64, C.POINTER(Aperture), C.POINTER(Isometry), C.POINTER(C.c_int),
                                   C.c_int]
    L.orc_rays_threshold.argtypes = [vp, C.c_uint64, C.c_double, C.c_int]
    L.orc_rays_filter_bounces.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.orc_rays_filter_bounces.restype = None
    L.orc_rays_filter_refractions.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.orc_rays_filter_refractions.restype = None
    L.orc_rays_filter_energy.argtypes = [vp, C.c_uint64, C.c_double]
    L.orc_rays_split.argtypes = [vp, C.c_uint64, C.c_double, vp]
    L.orc_rays_refract_paraxial.argtypes = [vp, C.c_uint64, C.c_double, C.POINTER(Isometry)]
    L.orc_rays_transform.argtypes = [vp, C.c_uint64, C.POINTER(Isometry), C.c_int]
    L.orc_rays_transform.restype = None
    L.orc_rays_total_energy.argtypes = [vp, C.c_uint64]
    L.orc_rays_total_energy.restype = C.c_double
    L.orc_rays_count.argtypes = [vp, C.c_uint64, C.c_int]
    L.orc_rays_count.restype = C.c_uint64
    for name in ("orc_rays_centroid", "orc_rays_energy_weighted_centroid"):
        getattr(L, name).argtypes = [vp, C.c_uint64, dp]
    for name in ("orc_rays_beam_radius_geo", "orc_rays_beam_radius_rms", "orc_rays_energy_weighted_beam_radius_rms"):
        getattr(L, name).argtypes = [vp, C.c_uint64, dp]
    L.orc_rays_bounce_lvl.argtypes = [vp, C.c_uint64]
    L.orc_rays_bounce_lvl.restype = C.c_uint32
    L.orc_hits_bounding_box.argtypes = [vp, C.c_uint64, dp]
    L.orc_fluence_binning.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, dp, dp, dp]
    L.orc_rays_diffract_on_periodic_surface.argtypes = [vp, C.c_uint64, C.POINTER(Surface), C.POINTER(IndexModel), dp, C.c_int32, C.c_int, vp,
                                                         C.POINTER(RefractStats)]
    L.orc_kde_distances_iqr.argtypes = [dp, C.c_uint64]
    L.orc_kde_distances_iqr.restype = C.c_double
    L.orc_kde_point_distances.argtypes = [vp, C.c_uint64, dp, dp]
    L.orc_kde_bandwidth_estimate.argtypes = [vp, C.c_uint64]
    L.orc_kde_bandwidth_estimate.restype = C.c_double
    L.orc_kde_value.argtypes = [vp, C.c_uint64, C.c_double, C.c_double, C.c_double]
    L.orc_kde_value.restype = C.c_double
    L.orc_fluence_kde.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, dp, dp, dp, dp, C.c_int]
    L.orc_fluence_peak.argtypes = [dp, C.c_uint64]
    L.orc_fluence_peak.restype = C.c_double
    L.orc_fluence_total_energy.argtypes = [dp, C.c_uint64, C.c_uint64, dp]
    L.orc_fluence_total_energy.restype = C.c_double
    L.orc_grid.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, dp]
    L.orc_grid.restype = None
    L.orc_fibonacci_rectangle.argtypes = [C.c_uint64, C.c_double, C.c_double, dp]
    L.orc_fibonacci_rectangle.restype = None
    L.orc_fibonacci_ellipse.argtypes = [C.c_uint64, C.c_double, C.c_double, dp]
    L.orc_fibonacci_ellipse.restype = None
    L.orc_hexapolar_count.argtypes = [C.c_uint32]
    L.orc_hexapolar_count.restype = C.c_uint64
    L.orc_hexapolar.argtypes = [C.c_uint32, C.c_double, dp]
    L.orc_hexapolar.restype = None
    L.orc_energy_uniform.argtypes = [C.c_uint64, C.c_double, dp]
    L.orc_energy_uniform.restype = None
    L.orc_energy_general_gaussian.argtypes = [dp, C.c_uint64] + [C.c_double] * 7 + [C.c_int, dp]
    L.orc_energy_general_gaussian.restype = None
    L.orc_linspace.argtypes = [C.c_double, C.c_double, C.c_uint64, dp]
    L.orc_spectrum_gaussian.argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_double, C.c_double, C.c_double, dp, dp]
    L.orc_spectrum_gaussian.restype = None
    L.orc_rays_collimated_with_spectrum.argtypes = [dp, dp, C.c_uint64, dp, dp, C.c_uint64, vp]
    u64p = C.POINTER(C.c_uint64)
    L.orc_spectrum_new.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp, C.c_uint64, u64p]
    L.orc_spectrum_add_single_peak.argtypes = [dp, dp, C.c_uint64, C.c_double, C.c_double]
    L.orc_spectrum_from_laser_lines.argtypes = [dp, dp, C.c_uint64, C.c_double, dp, dp, C.c_uint64, u64p]
    L.orc_rays_to_spectrum.argtypes = [vp, C.c_uint64, C.c_double, dp, dp, C.c_uint64, u64p]
    L.orc_spectrum_total_energy.argtypes = [dp, dp, C.c_uint64]
    L.orc_spectrum_total_energy.restype = C.c_double
    L.orc_spectrum_center_wavelength.argtypes = [dp, dp, C.c_uint64]
    L.orc_spectrum_center_wavelength.restype = C.c_double
    L.orc_spectrum_get_value.argtypes = [dp, dp, C.c_uint64, C.c_double, dp]
    L.orc_spectrum_add_lorentzian_peak.argtypes = [dp, dp, C.c_uint64, C.c_double, C.c_double, C.c_double]
    L.orc_spectrum_scale_vertical.argtypes = [dp, C.c_uint64, C.c_double]
    L.orc_spectrum_calc_ratio.argtypes = [C.c_double] * 4
    L.orc_spectrum_calc_ratio.restype = C.c_double
    L.orc_spectrum_resample.argtypes = [dp, dp, C.c_uint64, dp, dp, C.c_uint64]
    L.orc_spectrum_resample.restype = None
    for name in ("orc_spectrum_filter", "orc_spectrum_add", "orc_spectrum_sub"):
        getattr(L, name).argtypes = [dp, dp, C.c_uint64, dp, dp, C.c_uint64]
    L.orc_spectrum_split_by_spectrum.argtypes = [dp, dp, C.c_uint64, dp, dp, C.c_uint64, dp]
    L.orc_spectrum_average_resolution.argtypes = [dp, C.c_uint64]
    L.orc_spectrum_average_resolution.restype = C.c_double
    L.orc_spectrum_merge.argtypes = [dp, dp, C.c_uint64, dp, dp, C.c_uint64, dp, dp, C.c_uint64, u64p]
    L.orc_spectrum_from_csv.argtypes = [C.c_char_p, dp, dp, C.c_uint64, u64p]
    L.orc_spectrum_generate_filter.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, dp, C.c_uint64, u64p]
    L.orc_rays_filter_energy_spectrum.argtypes = [vp, C.c_uint64, dp, dp, C.c_uint64]
    L.orc_rays_split_spectrum.argtypes = [vp, C.c_uint64, dp, dp, C.c_uint64, vp]
    L.orc_rays_wavefront_error.argtypes = [vp, C.c_uint64, C.c_double, C.POINTER(Isometry), dp, u64p]
    L.orc_wavefront_statistics.argtypes = [dp, C.c_uint64, dp, dp]
    L.orc_history_attach.argtypes = [vp, C.c_uint64, dp, vp, C.c_uint32]
    L.orc_history_attach.restype = None
    L.orc_history_detach.restype = C.c_int
    _lib = L
    return L


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ---------------------------------------------------------------------------
# thin functional API over arrays of RAY_DTYPE
# ---------------------------------------------------------------------------
def max_threads() -> int:
    return lib().orc_max_threads()


def find_roots_quadratic(a2, a1, a0):
    r = (C.c_double * 2)()
    n = lib().orc_find_roots_quadratic(a2, a1, a0, r)
    return list(r[:n])


def fresnel(direction, normal, n1, n2) -> float:
    return lib().orc_fresnel_reflectivity((C.c_double * 3)(*direction), (C.c_double * 3)(*normal), n1, n2)


def apodization_factor(ap: Aperture, x: float, y: float) -> float:
    return lib().orc_apodization_factor(C.byref(ap), x, y)


def intersect(surf: Surface, pos, direction):
    q, n = (C.c_double * 3)(), (C.c_double * 3)()
    hit = lib().orc_intersect(C.byref(surf), (C.c_double * 3)(*pos), (C.c_double * 3)(*direction), q, n)
    return (np.array(q[:]), np.array(n[:])) if hit else None


def ray_new(pos, direction, wvl, e) -> np.ndarray:
    """Ray::new (ray.rs:108-140) -> 1-element RAY_DTYPE array."""
    out = np.zeros(1, dtype=RAY_DTYPE)
    _check(lib().orc_ray_new((C.c_double * 3)(*pos), (C.c_double * 3)(*direction), wvl, e, _vp(out)))
    return out


def define_up_direction(direction) -> np.ndarray:
    """Ray::define_up_direction (ray.rs:818-838)"""
    up = (C.c_double * 3)()
    lib().orc_define_up_direction((C.c_double * 3)(*direction), up)
    return np.array(up[:])


def calc_new_up_direction(prev_dir, direction, up) -> np.ndarray:
    """Ray::calc_new_up_direction (ray.rs:842-860): the rotated up-direction"""
    u = (C.c_double * 3)(*up)
    lib().orc_calc_new_up_direction((C.c_double * 3)(*prev_dir), (C.c_double * 3)(*direction), u)
    return np.array(u[:])


def rays_from_arrays(pos, direction, e, wvl) -> np.ndarray:
    n = len(e)
    out = np.zeros(n, dtype=RAY_DTYPE)
    out["pos"] = pos
    d = np.asarray(direction, dtype=np.float64)
    out["dir"] = d / np.sqrt((d * d).sum(axis=-1, keepdims=True))
    out["e"], out["wvl"], out["n"], out["valid"] = e, wvl, 1.0, 1
    out["id"] = np.arange(n, dtype=np.uint32)
    return out


def ray_refract_on_surface(ray: np.ndarray, surf: Surface, n2, strategy=MISS_STOP):
    """Ray::refract_on_surface (ray.rs:580-688) on a 1-element array, in place. Returns (child|None, hit|None)."""
    child = np.zeros(1, dtype=RAY_DTYPE)
    hit = np.zeros(1, dtype=HIT_DTYPE)
    hc, hh = C.c_int(), C.c_int()
    _check(lib().orc_ray_refract_on_surface(_vp(ray), C.byref(surf), int(n2 is not None), float(n2 or 0.0), strategy,
                                            _vp(child), C.byref(hc), _vp(hit), C.byref(hh)))
    return (child if hc.value else None), (hit if hh.value else None)


def rays_refract_on_surface(rays: np.ndarray, surf: Surface, idx: IndexModel | None, refraction_intended=True,
                            strategy=MISS_STOP, n_threads=1, want_children=True, want_hits=True):
    """Rays::refract_on_surface (rays.rs:976-1047), in place. Returns (children, hits, stats); outputs that are not wanted
    are not materialised (None)."""
    n = len(rays)
    children = np.empty(max(n, 1), dtype=RAY_DTYPE) if want_children else None
    hits = np.empty(max(n, 1), dtype=HIT_DTYPE) if want_hits else None
    st = RefractStats()
    _check(lib().orc_rays_refract_on_surface(_vp(rays), n, C.byref(surf), C.byref(idx) if idx is not None else None,
                                             int(refraction_intended), strategy, _vp(children) if want_children else None,
                                             _vp(hits) if want_hits else None, C.byref(st), n_threads))
    return (children[: st.n_children].copy() if want_children else None, hits[: st.n_hits].copy() if want_hits else None, st)


def rays_apodize(rays, ap: Aperture, iso: Isometry, n_threads=1) -> bool:
    inv = C.c_int()
    _check(lib().orc_rays_apodize(_vp(rays), len(rays), C.byref(ap), C.byref(iso), C.byref(inv), n_threads))
    return bool(inv.value)


def rays_threshold(rays, min_energy: float, n_threads=1):
    _check(lib().orc_rays_threshold(_vp(rays), len(rays), min_energy, n_threads))


def rays_filter_bounces(rays, max_bounces: int):
    lib().orc_rays_filter_bounces(_vp(rays), len(rays), max_bounces)


def rays_filter_refractions(rays, max_refr: int):
    lib().orc_rays_filter_refractions(_vp(rays), len(rays), max_refr)


def rays_filter_energy(rays, t: float):
    _check(lib().orc_rays_filter_energy(_vp(rays), len(rays), t))


def rays_split(rays, ratio: float) -> np.ndarray:
    out = np.zeros(len(rays), dtype=RAY_DTYPE)
    _check(lib().orc_rays_split(_vp(rays), len(rays), ratio, _vp(out)))
    return out


def rays_refract_paraxial(rays, f: float, iso: Isometry):
    _check(lib().orc_rays_refract_paraxial(_vp(rays), len(rays), f, C.byref(iso)))


def rays_diffract_on_periodic_surface(rays, surf: Surface, idx: IndexModel, grating_vector, order: int, refraction_intended=False):
    """Rays::diffract_on_periodic_surface (rays.rs:1058-1111), in place -> (diffracted clones, stats)"""
    n = len(rays)
    children = np.empty(max(n, 1), dtype=RAY_DTYPE)
    g = (C.c_double * 3)(*[float(x) for x in grating_vector])
    st = RefractStats()
    _check(lib().orc_rays_diffract_on_periodic_surface(_vp(rays), n, C.byref(surf), C.byref(idx), g, int(order),
                                                       int(refraction_intended), _vp(children), C.byref(st)))
    return children[: st.n_children].copy(), st


class History:
    """`pos_hist` of every ray of one bundle (ray.rs:70): positions xyz[i, k], k < len[i], row i = ray i of the array the
    history is attached to.  attach() makes the C oracle record the reference's pushes (ray.rs:537,616; rays.rs:566) for
    the rays of that array until detach()."""

    def __init__(self, n: int, cap: int = 64):
        self.xyz = np.zeros((max(n, 1), cap, 3), dtype=np.float64)
        self.len = np.zeros(max(n, 1), dtype=np.uint32)
        self.n = n

    def attach(self, rays: np.ndarray):
        assert len(rays) == self.n and rays.flags["C_CONTIGUOUS"]
        lib().orc_history_attach(_vp(rays), len(rays), _dptr(self.xyz), _vp(self.len), self.xyz.shape[1])

    @staticmethod
    def detach():
        if lib().orc_history_detach():
            raise RuntimeError("oracle position history overflow: raise History(cap=...)")

    def select(self, rows) -> "History":
        """history of clones of the rays `rows` (a cloned Ray carries its pos_hist)"""
        rows = np.asarray(rows, dtype=np.int64)
        h = History(len(rows), self.xyz.shape[1])
        if len(rows):
            h.xyz[: len(rows)] = self.xyz[rows]
            h.len[: len(rows)] = self.len[rows]
        return h

    def copy(self) -> "History":
        return self.select(np.arange(self.n))

    def of(self, i: int, rays: np.ndarray | None = None) -> np.ndarray:
        """Ray::position_history (ray.rs:296-308), or with `rays` position_history_with_current (:313-327)"""
        h = self.xyz[i, : self.len[i]]
        return np.vstack([h, rays["pos"][i][None, :]]) if rays is not None else h.copy()


def rays_transform(rays, iso: Isometry, inverse=False):
    lib().orc_rays_transform(_vp(rays), len(rays), C.byref(iso), int(inverse))


def total_energy(rays) -> float:
    return lib().orc_rays_total_energy(_vp(rays), len(rays))


def nr_of_rays(rays, valid_only=True) -> int:
    return lib().orc_rays_count(_vp(rays), len(rays), int(valid_only))


def _opt3(fn, rays):
    c = (C.c_double * 3)()
    return np.array(c[:]) if fn(_vp(rays), len(rays), c) else None


def centroid(rays):
    return _opt3(lib().orc_rays_centroid, rays)


def energy_weighted_centroid(rays):
    return _opt3(lib().orc_rays_energy_weighted_centroid, rays)


def _opt1(fn, rays):
    c = C.c_double()
    return c.value if fn(_vp(rays), len(rays), C.byref(c)) else None


def beam_radius_geo(rays):
    return _opt1(lib().orc_rays_beam_radius_geo, rays)


def beam_radius_rms(rays):
    return _opt1(lib().orc_rays_beam_radius_rms, rays)


def energy_weighted_beam_radius_rms(rays):
    return _opt1(lib().orc_rays_energy_weighted_beam_radius_rms, rays)


def bounce_lvl(rays) -> int:
    return lib().orc_rays_bounce_lvl(_vp(rays), len(rays))


def hits_from_xye(x, y, e, z=None) -> np.ndarray:
    h = np.zeros(len(x), dtype=HIT_DTYPE)
    h["x"], h["y"], h["e"] = x, y, e
    if z is not None:
        h["z"] = z
    return h


def fluence_binning(hits: np.ndarray, nx: int, ny: int, ranges=None):
    """calc_fluence_with_binning (rays_hit_map.rs:330-391). Returns (map[ny,nx], (left,right,top,bottom))."""
    m = np.zeros(nx * ny, dtype=np.float64)
    lrtb = (C.c_double * 4)()
    rg = (C.c_double * 4)(*ranges) if ranges is not None else None
    _check(lib().orc_fluence_binning(_vp(hits), len(hits), nx, ny, rg, _dptr(m), lrtb))
    return m.reshape(nx, ny).T.copy(), tuple(lrtb[:])  # column-major ny x nx -> [row=y, col=x]


def voronoi_site_fluence(x_cm, y_cm, e):
    """per-site fluence E_i / area(Voronoi cell i) in J/cm^2 (rays_hit_map.rs:432-458); returns (fluence, hull area in cm^2)"""
    x = np.ascontiguousarray(x_cm, dtype=np.float64); y = np.ascontiguousarray(y_cm, dtype=np.float64)
    ee = np.ascontiguousarray(e, dtype=np.float64)
    out = np.zeros(len(x))
    area = C.c_double()
    _check(lib().orc_voronoi_site_fluence(_dptr(x), _dptr(y), _dptr(ee), len(x), _dptr(out), C.byref(area)))
    return out, area.value


def fluence_voronoi(hits: np.ndarray, nx: int, ranges=None):
    """RaysHitMap::calc_fluence_with_voronoi (rays_hit_map.rs:407-521) of ONE rays hit map.  ranges = (ax1.start, ax1.end,
    ax2.start, ax2.end) in metres or None.  Returns (map[nx, nx] indexed [y index, x index] in J/m^2, (left, right, top, bottom))."""
    m = np.zeros(nx * nx, dtype=np.float64)
    lrtb = (C.c_double * 4)()
    rg = (C.c_double * 4)(*ranges) if ranges is not None else None
    _check(lib().orc_fluence_voronoi(_vp(hits), len(hits), nx, rg, _dptr(m), lrtb))
    return m.reshape(nx, nx).T.copy(), tuple(lrtb[:])


def fluence_helper_rays(hits: np.ndarray, nx: int, ranges=None):
    """RaysHitMap::calc_fluence_with_helper_rays (rays_hit_map.rs:627-722) of ONE rays hit map whose `e` field holds FLUENCES in
    J/m^2 (FluenceHitPoint).  Returns like fluence_voronoi."""
    m = np.zeros(nx * nx, dtype=np.float64)
    lrtb = (C.c_double * 4)()
    rg = (C.c_double * 4)(*ranges) if ranges is not None else None
    _check(lib().orc_fluence_helper_rays(_vp(hits), len(hits), nx, rg, _dptr(m), lrtb))
    return m.reshape(nx, nx).T.copy(), tuple(lrtb[:])


def fluence_helper_rays_combined(hit_maps, nx: int, ny: int):
    """HitMap::calc_combined_fluence_with_helper_rays (hit_map/mod.rs:283-310): every rays hit map on the common bounding box, summed
    (nx x nx estimates onto an nx x ny matrix: the reference panics unless nx == ny)"""
    if nx != ny:
        raise ValueError("calc_combined_fluence_with_helper_rays: matrix dimensions mismatch (the reference panics unless nx == ny)")
    x0 = min(float(h["x"].min()) for h in hit_maps); x1 = max(float(h["x"].max()) for h in hit_maps)
    y0 = min(float(h["y"].min()) for h in hit_maps); y1 = max(float(h["y"].max()) for h in hit_maps)
    total = np.zeros((nx, nx))
    for h in hit_maps:
        m, _ = fluence_helper_rays(h, nx, (x0, x1, y0, y1))
        total += m
    return total, (x0, x1, y1, y0)


# ---------------------------------------------------------------------------
# fluence helper rays (FluenceRays, rays.rs:1470-1567): pure-Python loops over the single-ray functions -- small bundles only
# ---------------------------------------------------------------------------
def fluence_general_gaussian(xy, total_energy, mu=(0.0, 0.0), sigma=(1.0, 1.0), theta=0.0) -> np.ndarray:
    """fluence_distributions::General2DGaussian::apply (general_gaussian.rs:75-90, math_distribution_functions.rs:97-121): J/m^2"""
    peak = total_energy / (2. * math.pi * sigma[0] * sigma[1])
    st, ct = math.sin(theta), math.cos(theta)
    out = np.zeros(len(xy))
    for i, p in enumerate(np.asarray(xy, dtype=np.float64)):
        x_rot = math.fma(p[0] - mu[0], ct, -((p[1] - mu[1]) * st)) if hasattr(math, "fma") else (p[0] - mu[0]) * ct - (p[1] - mu[1]) * st
        y_rot = math.fma(p[1] - mu[1], ct, (p[0] - mu[0]) * st) if hasattr(math, "fma") else (p[1] - mu[1]) * ct + (p[0] - mu[0]) * st
        q = (x_rot / sigma[0]) * (x_rot / sigma[0]) + (y_rot / sigma[1]) ** 2
        out[i] = peak * math.exp(-0.5 * q)
    return out


def helpers_new(rays: np.ndarray, init_fluence) -> dict:
    """FluenceRays::new for every ray (rays.rs:1477-1516; Ray::new_collimated_w_fluence_helper, ray.rs:157-166).
    init_fluence in J/m^2.  Returns {"rays": (n, 3) RAY_DTYPE helper rays, "eff": (n,) effective energies}."""
    n = len(rays)
    fl = np.broadcast_to(np.asarray(init_fluence, dtype=np.float64), (n,))
    if np.any(~np.isfinite(fl)) or np.any(np.signbit(fl)):
        raise ValueError("Initial fluence of Fluence Rays must be non-negative and finite!")
    h = np.zeros((n, 3), dtype=RAY_DTYPE)
    eff = np.zeros(n)
    for i in range(n):
        area = rays["wvl"][i] * rays["wvl"][i]
        eff[i] = fl[i] * area
        radius = math.sqrt(4. * area / math.sqrt(27.))
        for k in range(3):
            ang = float(k) * 2. / 3. * math.pi
            pos = rays["pos"][i] + np.array([radius * math.cos(ang), radius * math.sin(ang), 0.])
            h[i, k] = ray_new(pos, rays["dir"][i], rays["wvl"][i], rays["e"][i])[0]
    return {"rays": h, "eff": eff}


def rays_new_collimated_w_fluence_helper(xyz, wvl: float, fluence):
    """Rays::new_collimated_w_fluence_helper (rays.rs:352-398): energy = Voronoi cell area x fluence (0 for a cell of fewer than three
    vertices); fluence in J/m^2.  Returns (rays, helpers)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = len(xyz)
    fl = np.broadcast_to(np.asarray(fluence, dtype=np.float64), (n,))
    inv_area, _ = voronoi_site_fluence(xyz[:, 0] / 1.0E-2, xyz[:, 1] / 1.0E-2, np.ones(n))  # 1 / area [1/cm^2]
    with np.errstate(divide="ignore", invalid="ignore"):
        area_m2 = (1.0 / inv_area) * 1.0E-4
    e = np.where(np.isfinite(area_m2), area_m2 * fl, 0.0)
    rays = rays_from_arrays(xyz, np.tile([0.0, 0.0, 1.0], (n, 1)), e, np.full(n, wvl))
    return rays, helpers_new(rays, fl)


def helper_ray_fluence(helpers: dict, i: int) -> float:
    """Ray::helper_ray_fluence (ray.rs:172-198) in J/m^2"""
    p = helpers["rays"]["pos"][i]
    ab, ac = p[0] - p[1], p[0] - p[2]
    c = (ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0])
    area = 0.5 * math.sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2])
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(helpers["eff"][i]) / np.float64(area))


def rays_refract_on_surface_w_helpers(rays: np.ndarray, helpers: dict, surf: Surface, idx: IndexModel | None, refraction_intended=True,
                                      strategy=MISS_STOP):
    """Rays::refract_on_surface for a bundle with FluenceRays (rays.rs:976-1047 with ray.rs:640-672), in place.  The main ray is
    refracted first: its hit point is a FluenceHitPoint with the fluence of the helper triangle as it still is, the effective
    energy takes 1 - R (R on the reflected clone); then the helpers of a ray that has a reflected clone are refracted, and the
    clone's helpers -- copies taken BEFORE that -- take the reflected directions (:996-1010).
    Returns (children, child_helpers, hits) with hits["e"] = fluence in J/m^2.  For a mirror (refraction_intended = False) the
    caller continues with the children, like nodes/thin_mirror.rs."""
    n = len(rays)
    children, hits, ch_rays, ch_eff = [], [], [], []
    for i in range(n):
        if rays["valid"][i] != 1:
            continue
        n2 = idx.get_refractive_index(float(rays["wvl"][i])) if idx is not None else None
        fluence = helper_ray_fluence(helpers, i)  # evaluated inside the main ray's refraction (ray.rs:655)
        shadow = rays[i:i + 1].copy()
        shadow["e"] = helpers["eff"][i]
        pre_helpers = helpers["rays"][i].copy()
        child, hit = ray_refract_on_surface(rays[i:i + 1], surf, n2, strategy)
        if child is None:
            continue
        if not math.isfinite(fluence) or fluence < 0 or (fluence == 0 and math.copysign(1.0, fluence) < 0):
            raise ValueError("FluenceHitPoint: the fluence must be finite and not negative")
        hit = hit.copy()
        hit["e"] = fluence
        hits.append(hit[0])
        sh_child, _ = ray_refract_on_surface(shadow, surf, n2, strategy)  # the same arithmetic on the effective energy
        helpers["eff"][i] = shadow["e"][0]
        refl_helpers = pre_helpers.copy()
        for k in range(3):
            hk = helpers["rays"][i, k:k + 1]
            hchild, _ = ray_refract_on_surface(hk, surf, n2, strategy)
            if hchild is not None:
                refl_helpers["dir"][k] = hchild["dir"][0]  # set_direction: stored as given (ray.rs set_direction)
        c = child.copy()
        if not refraction_intended:
            c["bounces"] -= 1  # reduce_bounce_counter (rays.rs:1016)
        children.append(c[0])
        ch_rays.append(refl_helpers)
        ch_eff.append(sh_child["e"][0])
    ch = np.array(children, dtype=RAY_DTYPE) if children else np.zeros(0, dtype=RAY_DTYPE)
    hh = np.array(hits, dtype=HIT_DTYPE) if hits else np.zeros(0, dtype=HIT_DTYPE)
    chh = {"rays": np.array(ch_rays, dtype=RAY_DTYPE).reshape(-1, 3) if ch_rays else np.zeros((0, 3), dtype=RAY_DTYPE),
           "eff": np.array(ch_eff, dtype=np.float64)}
    return ch, chh, hh


def rays_refract_paraxial_w_helpers(rays: np.ndarray, helpers: dict, f: float, iso: Isometry):
    """Rays::refract_paraxial (rays.rs:943-958): the ray if it is valid, and its VALID helpers whether it is or not"""
    rays_refract_paraxial(rays, f, iso)
    for i in range(len(rays)):
        rays_refract_paraxial(helpers["rays"][i], f, iso)


def helpers_as_product_layout(helpers: dict) -> dict:
    """the oracle's helper rays in the layout of opossum_b200.Rays.fluence_helpers()"""
    h = helpers["rays"]
    nn = np.where(h["valid"] == 1, h["n"], -h["n"]) if len(h) else np.zeros((0, 3))
    return {"pos": h["pos"].copy(), "dir": h["dir"].copy(), "n": nn, "eff": helpers["eff"].copy()}


def fluence_voronoi_combined(hit_maps, nx: int, ny: int):
    """HitMap::calc_combined_fluence_with_voronoi (hit_map/mod.rs:232-262): the bounding box of ALL rays hit maps
    (get_bounding_box, :178-203), one Voronoi estimate per rays hit map -- per (bounce level, bundle) -- on that common grid, summed
    (NaN outside a bundle's convex hull propagates).  The reference adds nx x nx estimates onto an nx x ny matrix: it panics
    unless nx == ny (so does this: ValueError)."""
    if nx != ny:
        raise ValueError("calc_combined_fluence_with_voronoi: matrix dimensions mismatch (the reference panics unless nx == ny)")
    x0 = min(float(h["x"].min()) for h in hit_maps); x1 = max(float(h["x"].max()) for h in hit_maps)
    y0 = min(float(h["y"].min()) for h in hit_maps); y1 = max(float(h["y"].max()) for h in hit_maps)
    total = np.zeros((nx, nx))
    for h in hit_maps:
        m, _ = fluence_voronoi(h, nx, (x0, x1, y0, y1))
        total += m
    return total, (x0, x1, y1, y0)


def kde_distances_iqr(values) -> float:
    v = np.ascontiguousarray(values, dtype=np.float64)
    return lib().orc_kde_distances_iqr(_dptr(v), len(v))


def kde_point_distances(hits: np.ndarray):
    """Kde::point_distances_std_dev (kde/mod.rs:44-64) -> (distances, std_dev)"""
    n = len(hits)
    d = np.zeros(max(n * (n - 1) // 2, 1))
    sd = C.c_double()
    lib().orc_kde_point_distances(_vp(hits), n, _dptr(d), C.byref(sd))
    return d[: n * (n - 1) // 2], sd.value


def kde_bandwidth_estimate(hits: np.ndarray) -> float:
    return lib().orc_kde_bandwidth_estimate(_vp(hits), len(hits))


def kde_value(hits: np.ndarray, band_width: float, px: float, py: float) -> float:
    return lib().orc_kde_value(_vp(hits), len(hits), band_width, px, py)


def fluence_kde(hits: np.ndarray, nx: int, ny: int, ranges=None, n_threads=1):
    """calc_fluence_with_kde (rays_hit_map.rs:575-613). Returns (map[ny,nx], (left,right,top,bottom), band width)."""
    m = np.zeros(nx * ny, dtype=np.float64)
    lrtb = (C.c_double * 4)()
    rg = (C.c_double * 4)(*ranges) if ranges is not None else None
    bw = C.c_double()
    _check(lib().orc_fluence_kde(_vp(hits), len(hits), nx, ny, rg, _dptr(m), lrtb, C.byref(bw), n_threads))
    return m.reshape(nx, ny).T.copy(), tuple(lrtb[:]), bw.value


def hits_bounding_box(hits: np.ndarray):
    """RaysHitMap::calc_2d_bounding_box(margin 0) (rays_hit_map.rs:522-562) -> (left, right, top, bottom)"""
    out = (C.c_double * 4)()
    _check(lib().orc_hits_bounding_box(_vp(hits), len(hits), out))
    return tuple(out[:])


def fluence_peak(fmap: np.ndarray) -> float:
    m = np.ascontiguousarray(fmap, dtype=np.float64).ravel()
    return lib().orc_fluence_peak(_dptr(m), m.size)


def fluence_total_energy(fmap: np.ndarray, lrtb) -> float:
    ny, nx = fmap.shape
    m = np.ascontiguousarray(fmap, dtype=np.float64).ravel()
    return lib().orc_fluence_total_energy(_dptr(m), nx, ny, (C.c_double * 4)(*lrtb))


# ---- sources -------------------------------------------------------------
def grid(nx, ny, side_x, side_y) -> np.ndarray:
    out = np.zeros((max(nx, 1) * max(ny, 1), 3))
    lib().orc_grid(nx, ny, side_x, side_y, _dptr(out))
    return out


def fibonacci_rectangle(n, sx, sy) -> np.ndarray:
    out = np.zeros((n, 3))
    lib().orc_fibonacci_rectangle(n, sx, sy, _dptr(out))
    return out


def fibonacci_ellipse(n, rx, ry) -> np.ndarray:
    out = np.zeros((n, 3))
    lib().orc_fibonacci_ellipse(n, rx, ry, _dptr(out))
    return out


def hexapolar(rings, radius) -> np.ndarray:
    out = np.zeros((lib().orc_hexapolar_count(rings), 3))
    lib().orc_hexapolar(rings, radius, _dptr(out))
    return out if radius != 0.0 else out[:1]


def energy_uniform(n, total) -> np.ndarray:
    out = np.zeros(n)
    lib().orc_energy_uniform(n, total, _dptr(out))
    return out


def energy_general_gaussian(xyz, total, mu=(0.0, 0.0), sigma=(1.0, 1.0), power=1.0, theta=0.0,
                            rectangular=False) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.zeros(len(xyz))
    lib().orc_energy_general_gaussian(_dptr(xyz), len(xyz), total, mu[0], mu[1], sigma[0], sigma[1], power, theta,
                                      int(rectangular), _dptr(out))
    return out


def linspace(start, end, num) -> np.ndarray:
    out = np.zeros(num)
    _check(lib().orc_linspace(start, end, num, _dptr(out)))
    return out


def spectrum_gaussian(lo, hi, num, mu, fwhm, power=1.0):
    wvl, w = np.zeros(num), np.zeros(num)
    lib().orc_spectrum_gaussian(lo, hi, num, mu, fwhm, power, _dptr(wvl), _dptr(w))
    return wvl, w


def rays_collimated_with_spectrum(xyz, e_pos, wvl, w) -> np.ndarray:
    """Rays::new_collimated_with_spectrum (rays.rs:264-297): position-major, wavelength-minor."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    e_pos = np.ascontiguousarray(e_pos, dtype=np.float64)
    wvl = np.ascontiguousarray(np.atleast_1d(wvl), dtype=np.float64)
    w = np.ascontiguousarray(np.atleast_1d(w), dtype=np.float64)
    out = np.zeros(len(xyz) * len(wvl), dtype=RAY_DTYPE)
    _check(lib().orc_rays_collimated_with_spectrum(_dptr(xyz), _dptr(e_pos), len(xyz), _dptr(wvl), _dptr(w),
                                                   len(wvl), _vp(out)))
    return out


# ---------------------------------------------------------------------------
# spectrum.rs / nodes/spectrometer.rs / nodes/wavefront.rs.  A Spectrum is (lambda [um], value [1/um]).
# ---------------------------------------------------------------------------
def spectrum_new(start_m: float, end_m: float, res_m: float):
    n = C.c_uint64()
    _check(lib().orc_spectrum_new(start_m, end_m, res_m, None, None, 0, C.byref(n)))
    lam, data = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    _check(lib().orc_spectrum_new(start_m, end_m, res_m, _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam[: n.value], data[: n.value]


def spectrum_add_single_peak(lam, data, wl_m: float, value: float):
    _check(lib().orc_spectrum_add_single_peak(_dptr(lam), _dptr(data), len(lam), wl_m, value))


def spectrum_from_laser_lines(wl_m, e, res_m: float):
    wl, e = np.ascontiguousarray(wl_m, dtype=np.float64), np.ascontiguousarray(e, dtype=np.float64)
    n = C.c_uint64()
    _check(lib().orc_spectrum_from_laser_lines(_dptr(wl), _dptr(e), len(wl), res_m, None, None, 0, C.byref(n)))
    lam, data = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    _check(lib().orc_spectrum_from_laser_lines(_dptr(wl), _dptr(e), len(wl), res_m, _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam[: n.value], data[: n.value]


def rays_to_spectrum(rays, res_m: float):
    """Rays::to_spectrum (rays.rs:1209-1223)"""
    n = C.c_uint64()
    _check(lib().orc_rays_to_spectrum(_vp(rays), len(rays), res_m, None, None, 0, C.byref(n)))
    lam, data = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    _check(lib().orc_rays_to_spectrum(_vp(rays), len(rays), res_m, _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam[: n.value], data[: n.value]


def spectrum_total_energy(lam, data) -> float:
    return lib().orc_spectrum_total_energy(_dptr(lam), _dptr(data), len(lam))


def spectrum_center_wavelength(lam, data) -> float:
    return lib().orc_spectrum_center_wavelength(_dptr(lam), _dptr(data), len(lam))


def spectrum_get_value(lam, data, wl_m: float):
    lam, data = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(data, dtype=np.float64)
    v = C.c_double()
    return v.value if lib().orc_spectrum_get_value(_dptr(lam), _dptr(data), len(lam), wl_m, C.byref(v)) else None


def _spec(lam, data):
    return np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(data, dtype=np.float64)


def spectrum_add_lorentzian_peak(lam, data, center_m: float, width_m: float, energy: float):
    """Spectrum::add_lorentzian_peak (spectrum.rs:249-283), in place"""
    _check(lib().orc_spectrum_add_lorentzian_peak(_dptr(lam), _dptr(data), len(lam), center_m, width_m, energy))


def spectrum_scale_vertical(data, factor: float):
    _check(lib().orc_spectrum_scale_vertical(_dptr(data), len(data), factor))


def spectrum_calc_ratio(bl, br, sl, sr) -> float:
    return lib().orc_spectrum_calc_ratio(bl, br, sl, sr)


def spectrum_resample(lam, data, src_lam, src_data):
    """Spectrum::resample (spectrum.rs:377-428): (lam, data) takes the values of the source spectrum, in place"""
    sl, sd = _spec(src_lam, src_data)
    lib().orc_spectrum_resample(_dptr(lam), _dptr(data), len(lam), _dptr(sl), _dptr(sd), len(sl))


def spectrum_filter(lam, data, f_lam, f_data):
    fl, fd = _spec(f_lam, f_data)
    _check(lib().orc_spectrum_filter(_dptr(lam), _dptr(data), len(lam), _dptr(fl), _dptr(fd), len(fl)))


def spectrum_split_by_spectrum(lam, data, f_lam, f_data) -> np.ndarray:
    fl, fd = _spec(f_lam, f_data)
    out = np.zeros(len(lam))
    _check(lib().orc_spectrum_split_by_spectrum(_dptr(lam), _dptr(data), len(lam), _dptr(fl), _dptr(fd), len(fl), _dptr(out)))
    return out


def spectrum_add(lam, data, o_lam, o_data):
    ol, od = _spec(o_lam, o_data)
    _check(lib().orc_spectrum_add(_dptr(lam), _dptr(data), len(lam), _dptr(ol), _dptr(od), len(ol)))


def spectrum_sub(lam, data, o_lam, o_data):
    ol, od = _spec(o_lam, o_data)
    _check(lib().orc_spectrum_sub(_dptr(lam), _dptr(data), len(lam), _dptr(ol), _dptr(od), len(ol)))


def spectrum_average_resolution(lam) -> float:
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    return lib().orc_spectrum_average_resolution(_dptr(lam), len(lam))


def spectrum_merge(l1, d1, l2, d2):
    """merge_spectra (spectrum.rs:622-650) of two present spectra"""
    l1, d1 = _spec(l1, d1)
    l2, d2 = _spec(l2, d2)
    n = C.c_uint64()
    _check(lib().orc_spectrum_merge(_dptr(l1), _dptr(d1), len(l1), _dptr(l2), _dptr(d2), len(l2), None, None, 0, C.byref(n)))
    lam, data = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    _check(lib().orc_spectrum_merge(_dptr(l1), _dptr(d1), len(l1), _dptr(l2), _dptr(d2), len(l2), _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam[: n.value], data[: n.value]


FILTER_KINDS = {"short_pass_step": 0, "short_pass_smooth": 1, "long_pass_step": 2, "long_pass_smooth": 3}


def spectrum_generate_filter(kind: str, start_m, end_m, res_m, cut_off_m, width_m=0.0):
    """generate_filter_spectrum (spectrum_helper.rs:124-190)"""
    k = FILTER_KINDS[kind]
    n = C.c_uint64()
    _check(lib().orc_spectrum_generate_filter(k, start_m, end_m, res_m, cut_off_m, width_m, None, None, 0, C.byref(n)))
    lam, data = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    _check(lib().orc_spectrum_generate_filter(k, start_m, end_m, res_m, cut_off_m, width_m, _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam[: n.value], data[: n.value]


def spectrum_from_csv(path):
    """Spectrum::from_csv (spectrum.rs:89-117)"""
    n = C.c_uint64()
    _check(lib().orc_spectrum_from_csv(str(path).encode(), None, None, 0, C.byref(n)))
    lam, data = np.zeros(n.value), np.zeros(n.value)
    _check(lib().orc_spectrum_from_csv(str(path).encode(), _dptr(lam), _dptr(data), n.value, C.byref(n)))
    return lam, data


def rays_filter_energy_spectrum(rays, lam, data):
    lam, data = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(data, dtype=np.float64)
    _check(lib().orc_rays_filter_energy_spectrum(_vp(rays), len(rays), _dptr(lam), _dptr(data), len(lam)))


def rays_split_spectrum(rays, lam, data) -> np.ndarray:
    lam, data = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(data, dtype=np.float64)
    out = np.zeros(len(rays), dtype=RAY_DTYPE)
    _check(lib().orc_rays_split_spectrum(_vp(rays), len(rays), _dptr(lam), _dptr(data), len(lam), _vp(out)))
    return out


def rays_wavefront_error(rays, wavelength_m: float, monitor_iso: Isometry) -> np.ndarray:
    """Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802): (n_valid, 3) rows x [mm], y [mm], error [wavelengths]"""
    out = np.zeros((max(len(rays), 1), 3))
    n = C.c_uint64()
    _check(lib().orc_rays_wavefront_error(_vp(rays), len(rays), wavelength_m, C.byref(monitor_iso), _dptr(out), C.byref(n)))
    return out[: n.value]


def wavefront_statistics(xyw: np.ndarray):
    """WaveFrontErrorMap::calc_wavefront_statistics (nodes/wavefront.rs:145-164) -> (ptv, rms)"""
    xyw = np.ascontiguousarray(xyw, dtype=np.float64)
    ptv, rms = C.c_double(), C.c_double()
    _check(lib().orc_wavefront_statistics(_dptr(xyw), len(xyw), C.byref(ptv), C.byref(rms)))
    return ptv.value, rms.value
