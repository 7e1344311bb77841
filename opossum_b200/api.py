"""Host-side mirror of the reference's ray API over the C ABI (include/opossum_b200.h).

Names follow the reference (opossum/src): `Rays` (rays.rs), `OpticSurface` (surface/optic_surface.rs),
`Isometry` (utils/geom_transformation.rs), `RefractiveIndex*` (refractive_index/), `Aperture`
(aperture.rs), `CoatingType` (coatings/mod.rs), `HitMap`/fluence (surface/hit_map/).  Every method is a
thin call into libopossum_b200.so -- all ray arithmetic runs in the CUDA kernels; nothing here
computes on rays with numpy.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _ffi as F

RAY_DTYPE = np.dtype(
    [("pos", "<f8", 3), ("dir", "<f8", 3), ("e", "<f8"), ("wvl", "<f8"), ("opl", "<f8"), ("n", "<f8"),
     ("bounces", "<u4"), ("refr", "<u4"), ("valid", "<u4"), ("id", "<u4")], align=True)
assert RAY_DTYPE.itemsize == C.sizeof(F.OpmRay) == 96


class OpossumError(RuntimeError):
    """Mirror of `OpossumError` (error.rs:9-28): carries the status code and the reference's message."""

    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(message)


def _check(rc: int):
    if rc != F.OPM_OK:
        msg = F.lib().opm_last_error().decode() or F.lib().opm_status_string(rc).decode()
        raise OpossumError(rc, msg)


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ---------------------------------------------------------------------------------------------
class Isometry:
    """utils/geom_transformation.rs `Isometry` (transform + cached inverse), 3x3 + t form."""

    def __init__(self, c: F.OpmIsometry | None = None):
        self.c = c if c is not None else F.OpmIsometry()
        if c is None:
            F.lib().opm_isometry_identity(C.byref(self.c))

    @staticmethod
    def identity() -> "Isometry":
        return Isometry()

    @staticmethod
    def new(translation, euler_xyz=(0.0, 0.0, 0.0)) -> "Isometry":
        """Isometry::new(translation [m], axes angles [rad], applied x -> y -> z) :44-72"""
        o = F.OpmIsometry()
        _check(F.lib().opm_isometry_new(_d3(translation), _d3(euler_xyz), C.byref(o)))
        return Isometry(o)

    @staticmethod
    def new_along_z(z: float) -> "Isometry":
        return Isometry.new((0.0, 0.0, z))

    @staticmethod
    def new_from_view(view_point, view_direction, up_direction) -> "Isometry":
        """Isometry::new_from_view :166-185 (origin at the view point, local z along the view direction)"""
        o = F.OpmIsometry()
        _check(F.lib().opm_isometry_from_view(_d3(view_point), _d3(view_direction), _d3(up_direction), C.byref(o)))
        return Isometry(o)

    @staticmethod
    def parabola_off_axis(focal_length, oa_angle, oa_direction=(1.0, 0.0), collimating=False):
        """ParabolicMirror::calc_off_axis_isometry / calc_parent_focal_length (nodes/parabolic_mirror.rs:286-320)
        -> (anchor isometry, parent focal length)"""
        o, pf = F.OpmIsometry(), C.c_double()
        _check(F.lib().opm_parabola_off_axis_isometry(float(focal_length), float(oa_angle), (C.c_double * 2)(*oa_direction),
                                                      int(bool(collimating)), C.byref(o), C.byref(pf)))
        return Isometry(o), pf.value

    def append(self, rhs: "Isometry") -> "Isometry":
        o = F.OpmIsometry()
        F.lib().opm_isometry_append(C.byref(self.c), C.byref(rhs.c), C.byref(o))
        return Isometry(o)

    def matrices(self):
        return (np.array(self.c.fwd.rot[:]), np.array(self.c.fwd.t[:]), np.array(self.c.inv.rot[:]),
                np.array(self.c.inv.t[:]))


class RefractiveIndex:
    """refractive_index/mod.rs `RefractiveIndexType`"""

    def __init__(self, c: F.OpmIndexModel):
        self.c = c

    @staticmethod
    def const(n: float) -> "RefractiveIndex":
        m = F.OpmIndexModel(type=F.IDX_CONST)
        m.c[0] = n
        return RefractiveIndex(m)

    @staticmethod
    def sellmeier1(k1, k2, k3, l1, l2, l3, lo, hi) -> "RefractiveIndex":
        m = F.OpmIndexModel(type=F.IDX_SELLMEIER1, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [k1, k2, k3, l1, l2, l3]
        return RefractiveIndex(m)

    @staticmethod
    def schott(a0, a1, a2, a3, a4, a5, lo, hi) -> "RefractiveIndex":
        m = F.OpmIndexModel(type=F.IDX_SCHOTT, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [a0, a1, a2, a3, a4, a5]
        return RefractiveIndex(m)

    @staticmethod
    def conrady(n0, a, b, lo, hi) -> "RefractiveIndex":
        m = F.OpmIndexModel(type=F.IDX_CONRADY, wvl_lo=lo, wvl_hi=hi)
        m.c[:] = [n0, a, b, 0.0, 0.0, 0.0]
        return RefractiveIndex(m)


def refr_index_vaccuum() -> RefractiveIndex:
    return RefractiveIndex.const(1.0)


class Aperture:
    """aperture.rs `Aperture`"""

    def __init__(self):
        self.c = F.OpmAperture()
        self._keep = None

    @staticmethod
    def none() -> "Aperture":
        return Aperture()

    @staticmethod
    def _single(t, p, obstruction) -> "Aperture":
        a = Aperture()
        a.c.n_items = 1
        a.c.items[0].type = t
        a.c.items[0].obstruction = int(obstruction)
        a.c.items[0].p[:] = p
        return a

    @staticmethod
    def circle(r, cx=0.0, cy=0.0, obstruction=False) -> "Aperture":
        return Aperture._single(F.AP_CIRCLE, [r, cx, cy, 0.0], obstruction)

    @staticmethod
    def rect(w, h, cx=0.0, cy=0.0, obstruction=False) -> "Aperture":
        return Aperture._single(F.AP_RECT, [w, h, cx, cy], obstruction)

    @staticmethod
    def gaussian(sx, sy, cx=0.0, cy=0.0, obstruction=False) -> "Aperture":
        return Aperture._single(F.AP_GAUSSIAN, [sx, sy, cx, cy], obstruction)

    @staticmethod
    def polygon(triangles, obstruction=False) -> "Aperture":
        t = np.ascontiguousarray(np.asarray(triangles, dtype=np.float64).reshape(-1, 6))
        a = Aperture._single(F.AP_POLYGON, [0.0] * 4, obstruction)
        a._keep = t
        a.c.n_tris = t.shape[0]
        a.c.tris = t.ctypes.data_as(C.POINTER(C.c_double))
        return a

    @staticmethod
    def stack(items, obstruction=False) -> "Aperture":
        if len(items) > F.OPM_MAX_APERTURE_ITEMS:
            raise OpossumError(F.OPM_ERR_ARG, "aperture stack too deep")
        a = Aperture()
        a.c.n_items = len(items)
        a.c.is_stack = 1
        a.c.stack_obstruction = int(obstruction)
        for k, it in enumerate(items):
            if it.c.n_items != 1 or it.c.is_stack:
                raise OpossumError(F.OPM_ERR_ARG, "only flat stacks of simple apertures are supported")
            a.c.items[k] = it.c.items[0]
            if it.c.n_tris:
                a.c.n_tris, a.c.tris, a._keep = it.c.n_tris, it.c.tris, it._keep
        return a


class OpticSurface:
    """surface/optic_surface.rs `OpticSurface`: geometry + isometry of the geo surface + coating."""

    def __init__(self, geom=F.GEOM_PLANE, param=0.0, iso: Isometry | None = None, coating=F.COAT_IDEAL_AR,
                 reflectivity=0.0):
        self.c = F.OpmSurface(geom=geom, coating=coating, geom_param=param, coating_r=reflectivity)
        self.c.iso = (iso or Isometry.identity()).c

    @staticmethod
    def plane(iso=None, **kw):
        return OpticSurface(F.GEOM_PLANE, 0.0, iso, **kw)

    @staticmethod
    def sphere(radius, iso, **kw):
        """`iso` is the isometry of the centre of curvature (node_iso o T(0,0,R))."""
        return OpticSurface(F.GEOM_SPHERE, radius, iso, **kw)

    @staticmethod
    def sphere_at_position(radius, pos, **kw):
        """Sphere::new_at_position (surface/sphere.rs:30-41)"""
        return OpticSurface(F.GEOM_SPHERE, radius, Isometry.new((pos[0], pos[1], pos[2] + radius)), **kw)

    @staticmethod
    def cylinder(radius, iso, **kw):
        return OpticSurface(F.GEOM_CYLINDER, radius, iso, **kw)

    @staticmethod
    def parabola(focal_length, iso, **kw):
        return OpticSurface(F.GEOM_PARABOLA, focal_length, iso, **kw)


@dataclass
class RayTraceConfig:
    """analyzers/raytrace.rs:302-323"""
    min_energy_per_ray: float = 1.0 * 1.0e-12
    max_number_of_bounces: int = 1000
    max_number_of_refractions: int = 1000
    missed_surface_strategy: int = F.MISS_STOP


# ---------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + stream.  Raises OpossumError(OPM_ERR_CUDA) when no B200 is visible (no CPU fallback)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        """stream: a cudaStream_t handle the library launches on (e.g. `torch.cuda.Stream().cuda_stream`); None = a private
        non-blocking stream.  0 is CUDA's default stream (what `torch.cuda.default_stream().cuda_stream` returns): it is passed
        on as cudaStreamLegacy, never silently replaced by a private stream."""
        h = C.c_void_p()
        if stream is not None and stream == 0:
            stream = 0x1  # cudaStreamLegacy: the NULL stream under an explicit handle
        _check(F.lib().opm_context_create(device, C.c_void_p(stream) if stream is not None else None, C.byref(h)))
        self.h = h
        self.device = device
        self._children = weakref.WeakSet()  # bundles / hit maps / scenes living on this context

    def _adopt(self, child):
        self._children.add(child)

    def synchronize(self):
        """waits for the stream; raises what asynchronous launches (sync=False) raised on the device meanwhile"""
        _check(F.lib().opm_context_synchronize(self.h))

    def check_errors(self, clear=True):
        """raise the error status asynchronous launches left on the device, without waiting for queued work"""
        _check(F.lib().opm_context_check_errors(self.h, int(clear)))

    @property
    def stream(self) -> int:
        return F.lib().opm_context_stream(self.h) or 0

    @property
    def launch_count(self) -> int:
        return F.lib().opm_context_launch_count(self.h)

    def close(self):
        """destroys the context after every handle that still lives on it (device buffers hold a pointer to it)"""
        if self.h:
            for child in sorted(self._children, key=lambda c: getattr(c, "_close_order", 1)):
                child.close()
            F.lib().opm_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_kernel_timing(self, enable: bool):
        """bracket every launch of the fused trace kernel with CUDA events on this context's stream"""
        _check(F.lib().opm_context_set_kernel_timing(self.h, int(enable)))

    def kernel_timing(self, reset=True):
        """(accumulated trace-kernel device time in ms, number of launches); blocks until the events have completed"""
        ms, n = C.c_double(), C.c_uint64()
        _check(F.lib().opm_context_kernel_timing(self.h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    def set_jit(self, enable, min_rays: int = 200000):
        """run-time specialised trace kernels (NVRTC, cached per program structure) for bundles of >= min_rays rays.
        enable: 0 off, 1 compile in the background and interpret meanwhile (default), 2 wait for the compiler"""
        _check(F.lib().opm_context_set_jit(self.h, int(enable), min_rays))

    def jit_wait(self):
        """block until no kernel compilation is in flight"""
        _check(F.lib().opm_context_jit_wait(self.h))

    def jit_stats(self) -> dict:
        a, b, c, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
        _check(F.lib().opm_context_jit_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(ms)))
        return {"kernels_compiled": a.value, "compile_failures": b.value, "launches": c.value, "compile_ms": ms.value}

    def measure_fp64_peak(self) -> float:
        """DFMA throughput of this GPU in TFLOP/s (microbenchmark kernel in the library)"""
        out = C.c_double()
        _check(F.lib().opm_measure_fp64_peak(self.h, C.byref(out)))
        return out.value

    # device buffers for fluence maps
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _check(F.lib().opm_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p: int):
        F.lib().opm_device_free(self.h, C.c_void_p(p))

    def interactions_total(self) -> int:
        """ray-surface interactions of all launches of the context so far (a device counter; synchronises)"""
        out = C.c_uint64()
        _check(F.lib().opm_context_interactions_total(self.h, C.byref(out)))
        return out.value

    def device_memset(self, dev: int, value: int, nbytes: int):
        _check(F.lib().opm_device_memset(self.h, C.c_void_p(dev), value, nbytes))

    def to_host(self, dev: int, arr: np.ndarray):
        _check(F.lib().opm_device_to_host(self.h, C.c_void_p(arr.ctypes.data), C.c_void_p(dev), arr.nbytes))

    def to_device(self, dev: int, arr: np.ndarray):
        _check(F.lib().opm_host_to_device(self.h, C.c_void_p(dev), C.c_void_p(arr.ctypes.data), arr.nbytes))


class HitMap:
    """Device hit-point record of one surface (rays_hit_map.rs:43-49): surface-frame xyz + input energy + bounce."""

    def __init__(self, ctx: Context, capacity: int):
        self.ctx = ctx
        h = C.c_void_p()
        _check(F.lib().opm_hitmap_create(ctx.h, capacity, C.byref(h)))
        self.h = h
        ctx._adopt(self)

    def __len__(self):
        return F.lib().opm_hitmap_len(self.h)

    def upload(self, x, y, e, z=None, bounce=None):
        x, y, e = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, e))
        dp = C.POINTER(C.c_double)
        zz = np.ascontiguousarray(z, dtype=np.float64) if z is not None else None
        bb = np.ascontiguousarray(bounce, dtype=np.uint32) if bounce is not None else None
        _check(F.lib().opm_hitmap_upload(self.h, x.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                         zz.ctypes.data_as(dp) if zz is not None else None, e.ctypes.data_as(dp),
                                         bb.ctypes.data_as(C.POINTER(C.c_uint32)) if bb is not None else None, len(x)))

    def download(self):
        n = len(self)
        x, y, z, e = (np.zeros(max(n, 1)) for _ in range(4))
        b = np.zeros(max(n, 1), dtype=np.uint32)
        dp = C.POINTER(C.c_double)
        k = C.c_uint64()
        _check(F.lib().opm_hitmap_download(self.h, x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp),
                                           e.ctypes.data_as(dp), b.ctypes.data_as(C.POINTER(C.c_uint32)), max(n, 1),
                                           C.byref(k)))
        k = k.value
        return x[:k], y[:k], z[:k], e[:k], b[:k]

    def close(self):
        if self.h:
            F.lib().opm_hitmap_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class TraceStats:
    interactions: int = 0
    children: int = 0
    hits: int = 0
    missed: int = 0
    apodized: int = 0
    valid_out: int = 0
    alive_out: int = 0

    @staticmethod
    def of(s: F.OpmTraceStats) -> "TraceStats":
        return TraceStats(s.interactions, s.children, s.hits, s.missed, s.apodized, s.valid_out, s.alive_out)


class Spectrum:
    """spectrum.rs `Spectrum` resident on the device (for the spectrum-driven filter / splitter ops): ascending wavelength
    slots in micrometers and their values.  total_energy / center_wavelength / get_value are the library's host helpers."""

    def __init__(self, ctx: Context, lambda_um, values):
        self.ctx = ctx
        self.lam = np.ascontiguousarray(lambda_um, dtype=np.float64)
        self.val = np.ascontiguousarray(values, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        h = C.c_void_p()
        _check(F.lib().opm_spectrum_create(ctx.h, self.lam.ctypes.data_as(dp), self.val.ctypes.data_as(dp), len(self.lam), C.byref(h)))
        self.h = h
        ctx._adopt(self)

    def close(self):
        if self.h:
            F.lib().opm_spectrum_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectrum_total_energy(lam, val) -> float:
    lam, val = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(val, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    return F.lib().opm_spectrum_total_energy(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam))


def spectrum_center_wavelength(lam, val) -> float:
    lam, val = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(val, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    return F.lib().opm_spectrum_center_wavelength(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam))


def spectrum_get_value(lam, val, wavelength: float):
    lam, val = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(val, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    v = C.c_double()
    return v.value if F.lib().opm_spectrum_get_value(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam), wavelength, C.byref(v)) else None


def _sp(lam, val):
    lam, val = np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(val, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    return lam, val, lam.ctypes.data_as(dp), val.ctypes.data_as(dp)


def _sp_sized(call):
    """the two-call protocol of the functions that create a spectrum: slot count first, then the arrays"""
    n = C.c_uint64()
    _check(call(None, None, 0, C.byref(n)))
    lam, val = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    dp = C.POINTER(C.c_double)
    _check(call(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), n.value, C.byref(n)))
    return lam[: n.value], val[: n.value]


def spectrum_new(start: float, end: float, resolution: float):
    """Spectrum::new (spectrum.rs:47-73): (slots [um], zeros)"""
    return _sp_sized(lambda l, v, cap, n: F.lib().opm_spectrum_new(start, end, resolution, l, v, cap, n))


def spectrum_add_single_peak(lam, val, wavelength: float, value: float):
    """Spectrum::add_single_peak (spectrum.rs:193-227), in place (val must be a contiguous float64 array)"""
    dp = C.POINTER(C.c_double)
    _check(F.lib().opm_spectrum_add_single_peak(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam), wavelength, value))


def spectrum_from_laser_lines(wavelengths, energies, resolution: float):
    """Spectrum::from_laser_lines (spectrum.rs:126-148)"""
    w, e, wp, ep = _sp(wavelengths, energies)
    return _sp_sized(lambda l, v, cap, n: F.lib().opm_spectrum_from_laser_lines(wp, ep, len(w), resolution, l, v, cap, n))


def spectrum_add_lorentzian_peak(lam, val, center: float, width: float, energy: float):
    """Spectrum::add_lorentzian_peak (spectrum.rs:249-283), in place"""
    dp = C.POINTER(C.c_double)
    _check(F.lib().opm_spectrum_add_lorentzian_peak(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam), center, width, energy))


def spectrum_scale_vertical(val, factor: float):
    """Spectrum::scale_vertical (spectrum.rs:355-367), in place"""
    _check(F.lib().opm_spectrum_scale_vertical(val.ctypes.data_as(C.POINTER(C.c_double)), len(val), factor))


def spectrum_resample(lam, val, src_lam, src_val):
    """Spectrum::resample (spectrum.rs:377-428): val takes the source spectrum's values on the slots lam, in place"""
    sl, sv, slp, svp = _sp(src_lam, src_val)
    dp = C.POINTER(C.c_double)
    _check(F.lib().opm_spectrum_resample(lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam), slp, svp, len(sl)))


def _spectrum_combine(what, lam, val, o_lam, o_val, split=None):
    ol, ov, olp, ovp = _sp(o_lam, o_val)
    dp = C.POINTER(C.c_double)
    _check(F.lib().opm_spectrum_combine(what, lam.ctypes.data_as(dp), val.ctypes.data_as(dp), len(lam), olp, ovp, len(ol),
                                        split.ctypes.data_as(dp) if split is not None else None))


def spectrum_filter(lam, val, f_lam, f_val):
    """Spectrum::filter (spectrum.rs:430-439), in place"""
    _spectrum_combine(0, lam, val, f_lam, f_val)


def spectrum_add(lam, val, o_lam, o_val):
    """Spectrum::add (spectrum.rs:483-492), in place"""
    _spectrum_combine(1, lam, val, o_lam, o_val)


def spectrum_sub(lam, val, o_lam, o_val):
    """Spectrum::sub (spectrum.rs:497-507), in place"""
    _spectrum_combine(2, lam, val, o_lam, o_val)


def spectrum_split_by_spectrum(lam, val, f_lam, f_val) -> np.ndarray:
    """Spectrum::split_by_spectrum (spectrum.rs:456-472): val keeps the transmitted part, the rest is returned"""
    out = np.zeros(len(lam))
    _spectrum_combine(3, lam, val, f_lam, f_val, out)
    return out


def spectrum_average_resolution(lam) -> float:
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    return F.lib().opm_spectrum_average_resolution(lam.ctypes.data_as(C.POINTER(C.c_double)), len(lam))


def spectrum_merge(l1, v1, l2, v2):
    """merge_spectra (spectrum.rs:622-650)"""
    a, b, ap, bp = _sp(l1, v1)
    c, d, cp, dpp = _sp(l2, v2)
    return _sp_sized(lambda l, v, cap, n: F.lib().opm_spectrum_merge(ap, bp, len(a), cp, dpp, len(c), l, v, cap, n))


FILTER_KINDS = {"short_pass_step": 0, "short_pass_smooth": 1, "long_pass_step": 2, "long_pass_smooth": 3}


def spectrum_generate_filter(kind: str, start: float, end: float, resolution: float, cut_off: float, width: float = 0.0):
    """generate_filter_spectrum (spectrum_helper.rs:124-190): (slots [um], transmission)"""
    k = FILTER_KINDS[kind]
    return _sp_sized(lambda l, v, cap, n: F.lib().opm_spectrum_generate_filter(k, start, end, resolution, cut_off, width, l, v, cap, n))


def spectrum_from_csv(path):
    """Spectrum::from_csv (spectrum.rs:89-117)"""
    return _sp_sized(lambda l, v, cap, n: F.lib().opm_spectrum_from_csv(str(path).encode(), l, v, cap, n))


class Rays:
    """rays.rs `Rays`: a bundle held as struct-of-arrays f64 buffers in HBM."""

    def __init__(self, ctx: Context, capacity: int = 1):
        self.ctx = ctx
        h = C.c_void_p()
        _check(F.lib().opm_rays_create(ctx.h, capacity, C.byref(h)))
        self.h = h
        ctx._adopt(self)

    # ---- construction / inspection
    @staticmethod
    def from_array(ctx: Context, rays: np.ndarray) -> "Rays":
        """Rays::from(Vec<Ray>)"""
        a = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        r = Rays(ctx, max(len(a), 1))
        _check(F.lib().opm_rays_upload(r.h, C.c_void_p(a.ctypes.data), len(a)))
        return r

    def upload(self, rays: np.ndarray):
        """replaces the bundle's rays (a position history, if kept, starts over)"""
        a = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        _check(F.lib().opm_rays_upload(self.h, C.c_void_p(a.ctypes.data), len(a)))

    def upload_soa(self, pos, direction, e, wvl):
        """SoA upload (one H2D copy per field, from pinned memory when the arrays are pinned)."""
        dp = C.POINTER(C.c_double)
        P = (dp * 3)(*[p.ctypes.data_as(dp) for p in pos])
        D = (dp * 3)(*[d.ctypes.data_as(dp) for d in direction])
        _check(F.lib().opm_rays_upload_soa(self.h, P, D, e.ctypes.data_as(dp), wvl.ctypes.data_as(dp), len(e)))

    def new_collimated_with_spectrum(self, pos_xyz_ptr: int, e_pos_ptr: int, n_pos: int, wvl, weights):
        """Rays::new_collimated_with_spectrum (rays.rs:264-297) from host buffers (ideally pinned): raw addresses of the
        n_pos x 3 positions and the n_pos energies; asynchronous on the context's stream."""
        wv = np.ascontiguousarray(wvl, dtype=np.float64)
        wg = np.ascontiguousarray(weights, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        _check(F.lib().opm_rays_new_collimated_with_spectrum(self.h, C.c_void_p(pos_xyz_ptr), C.c_void_p(e_pos_ptr), n_pos,
                                                             wv.ctypes.data_as(dp), wg.ctypes.data_as(dp), len(wv)))

    def new_collimated_xy(self, pos_xy_ptr: int, n_pos: int, source: "SourceDesc"):
        """Rays::new_collimated_with_spectrum from the position distribution's (x, y) alone (n_pos x 2 doubles, ideally pinned):
        z = 0, energies from the source's EnergyDistribution descriptor, spectrum and isometry applied on the device."""
        _check(F.lib().opm_rays_new_collimated_xy(self.h, C.c_void_p(pos_xy_ptr), n_pos, C.byref(source.c)))

    def to_array(self) -> np.ndarray:
        n = len(self)
        out = np.zeros(max(n, 1), dtype=RAY_DTYPE)
        k = C.c_uint64()
        _check(F.lib().opm_rays_download(self.h, C.c_void_p(out.ctypes.data), max(n, 1), C.byref(k)))
        return out[: k.value]

    def __len__(self):
        return F.lib().opm_rays_len(self.h)

    def attach_fluence_helpers(self, init_fluence) -> "Rays":
        """FluenceRays::new for every ray (rays.rs:1477-1516): three helper rays around it, effective energy = fluence * wavelength^2.
        init_fluence: J/m^2, one value per ray (or a scalar)."""
        fl = np.ascontiguousarray(np.broadcast_to(np.asarray(init_fluence, dtype=np.float64), (len(self),)))
        _check(F.lib().opm_rays_attach_fluence_helpers(self.h, fl.ctypes.data_as(C.POINTER(C.c_double))))
        return self

    def new_collimated_w_fluence_helper(self, pos_xyz, wavelength: float, fluence) -> "Rays":
        """Rays::new_collimated_w_fluence_helper (rays.rs:352-398): energy = Voronoi cell area x fluence [J/m^2], FluenceRays attached"""
        p = np.ascontiguousarray(pos_xyz, dtype=np.float64).reshape(-1, 3)
        fl = np.ascontiguousarray(np.broadcast_to(np.asarray(fluence, dtype=np.float64), (len(p),)))
        dp = C.POINTER(C.c_double)
        _check(F.lib().opm_rays_new_collimated_w_fluence_helper(self.h, p.ctypes.data_as(dp), len(p), wavelength, fl.ctypes.data_as(dp)))
        return self

    @property
    def has_fluence_helpers(self) -> bool:
        return bool(F.lib().opm_rays_has_fluence_helpers(self.h))

    def fluence_helpers(self) -> dict:
        """the FluenceRays of every ray in the order of to_array(): pos (n, 3, 3), dir (n, 3, 3), n (n, 3; negative = that helper is
        invalid), eff (n,)"""
        n = max(len(self), 1)
        out = np.zeros((n, 22))
        k = C.c_uint64()
        _check(F.lib().opm_rays_download_fluence_helpers(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(k)))
        out = out[: k.value]
        return {"pos": out[:, 0:9].reshape(-1, 3, 3).copy(), "dir": out[:, 9:18].reshape(-1, 3, 3).copy(), "n": out[:, 18:21].copy(),
                "eff": out[:, 21].copy()}

    def clone(self) -> "Rays":
        r = Rays(self.ctx, max(len(self), 1))
        _check(F.lib().opm_rays_copy(self.h, r.h))
        return r

    def add_rays(self, other: "Rays"):
        _check(F.lib().opm_rays_append(self.h, other.h))

    merge = add_rays

    def compact(self):
        _check(F.lib().opm_rays_compact(self.h))

    def clear(self):
        _check(F.lib().opm_rays_clear(self.h))

    # ---- spectra / wavefront (spectrum.rs, nodes/spectrometer.rs, nodes/wavefront.rs)
    def set_wavelength_hint(self, wavelength: float):
        """performance hint: (nearly) every ray carries this wavelength [m]; checked per ray on the device, 0 = none"""
        _check(F.lib().opm_rays_set_wavelength_hint(self.h, float(wavelength)))

    def wavelength_range(self):
        """(min, max, n_valid) of the valid rays' wavelengths"""
        mm = (C.c_double * 2)()
        nv = C.c_uint64()
        _check(F.lib().opm_rays_wavelength_range(self.h, mm, C.byref(nv)))
        return mm[0], mm[1], nv.value

    def to_spectrum(self, resolution: float, wavelength_range=None):
        """Rays::to_spectrum (rays.rs:1209-1223) -> (wavelength slots [um], values [1/um]); wavelength_range = (min, max) of
        all shards when the bundle is one shard of several (then the values of the shards add up)"""
        rng = (C.c_double * 2)(*wavelength_range) if wavelength_range is not None else None
        n = C.c_uint64()
        _check(F.lib().opm_rays_to_spectrum(self.h, resolution, rng, None, None, 0, C.byref(n)))
        lam, val = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
        dp = C.POINTER(C.c_double)
        _check(F.lib().opm_rays_to_spectrum(self.h, resolution, rng, lam.ctypes.data_as(dp), val.ctypes.data_as(dp), n.value, C.byref(n)))
        return lam[: n.value], val[: n.value]

    def filter_energy_spectrum(self, spectrum: "Spectrum"):
        """Rays::filter_energy(FilterType::Spectrum) (rays.rs:1119-1133, ray.rs:712-721)"""
        _check(F.lib().opm_rays_filter_energy_spectrum(self.h, spectrum.h))

    def split_spectrum(self, spectrum: "Spectrum") -> "Rays":
        """Rays::split(SplittingConfig::Spectrum) (rays.rs:1253-1264, ray.rs:752-772)"""
        out = Rays(self.ctx, max(len(self), 1))
        _check(F.lib().opm_rays_split_spectrum(self.h, spectrum.h, out.h))
        return out

    def wavefront_error(self, monitor_iso: "Isometry", wavelength: float = 0.0, want_map: bool = True):
        """Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802) + calc_wavefront_statistics (nodes/wavefront.rs:145-164)
        -> (OpmWavefrontStats, (n_valid, 3) array or None); wavelength <= 0: centre wavelength of to_spectrum(1 nm)"""
        st = F.OpmWavefrontStats()
        n = len(self)
        xyw = np.zeros((max(n, 1), 3)) if want_map else None
        _check(F.lib().opm_rays_wavefront_error(self.h, C.byref(monitor_iso.c), wavelength,
                                                xyw.ctypes.data_as(C.POINTER(C.c_double)) if want_map else None, max(n, 1), C.byref(st)))
        return st, (xyw[: st.n_rays] if want_map else None)

    # ---- per-ray position history (`pos_hist`, ray.rs:70)
    def enable_position_history(self, levels_hint: int = 16):
        """from now on every refract / diffract / clipping apodize call logs the positions the reference pushes"""
        _check(F.lib().opm_rays_enable_position_history(self.h, levels_hint))

    def disable_position_history(self):
        _check(F.lib().opm_rays_disable_position_history(self.h))

    def clear_position_history(self):
        """Ray::clear_pos_hist (ray.rs:225-227) for every ray"""
        _check(F.lib().opm_rays_clear_position_history(self.h))

    @property
    def position_history_levels(self) -> int:
        return F.lib().opm_rays_position_history_levels(self.h)

    def position_history(self, with_current: bool = True):
        """Rays::get_rays_position_history (rays.rs:1357-1383) without the wavelength grouping: one (len_i, 3) array per
        ray in bundle order -- Ray::position_history_with_current / position_history (ray.rs:296-327)"""
        levels = self.position_history_levels
        if levels < 0:
            raise OpossumError(F.OPM_ERR_ARG, "the bundle keeps no position history")
        n, cap = len(self), levels + 1
        xyz = np.zeros((max(n, 1), cap, 3), dtype=np.float64)
        ln = np.zeros(max(n, 1), dtype=np.uint32)
        k = C.c_uint64()
        _check(F.lib().opm_rays_position_history(self.h, 1 if with_current else 0, C.c_void_p(xyz.ctypes.data),
                                                 C.c_void_p(ln.ctypes.data), max(n, 1), cap, C.byref(k)))
        return [xyz[i, : ln[i]].copy() for i in range(k.value)]

    def field_ptr(self, name: str) -> int:
        return F.lib().opm_rays_field_ptr(self.h, name.encode()) or 0

    def close(self):
        if self.h:
            F.lib().opm_rays_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's bundle methods
    def refract_on_surface(self, surface: OpticSurface, refractive_index: RefractiveIndex | None,
                           refraction_intended: bool = True, missed_surface_strategy: int = F.MISS_STOP,
                           hit_map: HitMap | None = None, want_reflected: bool = True):
        """Rays::refract_on_surface (rays.rs:976-1047) -> (reflected Rays, TraceStats)."""
        reflected = Rays(self.ctx, max(len(self), 1)) if want_reflected else None
        st = F.OpmTraceStats()
        _check(F.lib().opm_rays_refract_on_surface(
            self.h, C.byref(surface.c), C.byref(refractive_index.c) if refractive_index is not None else None,
            int(refraction_intended), missed_surface_strategy, reflected.h if reflected is not None else None,
            hit_map.h if hit_map is not None else None, C.byref(st)))
        return reflected, TraceStats.of(st)

    def diffract_on_periodic_surface(self, surface: OpticSurface, refractive_index: RefractiveIndex, grating_vector,
                                     diffraction_order: int, refraction_intended: bool = False):
        """Rays::diffract_on_periodic_surface (rays.rs:1058-1111) -> (diffracted Rays, TraceStats)"""
        out = Rays(self.ctx, max(len(self), 1))
        st = F.OpmTraceStats()
        g = (C.c_double * 3)(*[float(x) for x in grating_vector])
        _check(F.lib().opm_rays_diffract_on_periodic_surface(self.h, C.byref(surface.c), C.byref(refractive_index.c), g,
                                                             int(diffraction_order), int(refraction_intended), out.h, C.byref(st)))
        return out, TraceStats.of(st)

    def apodize(self, aperture: Aperture, iso: Isometry) -> bool:
        inv = C.c_int32()
        _check(F.lib().opm_rays_apodize(self.h, C.byref(aperture.c), C.byref(iso.c), C.byref(inv)))
        return bool(inv.value)

    def invalidate_by_threshold_energy(self, min_energy_per_ray: float):
        _check(F.lib().opm_rays_invalidate_by_threshold_energy(self.h, min_energy_per_ray))

    def filter_by_nr_of_bounces(self, max_bounces: int):
        _check(F.lib().opm_rays_filter_by_nr_of_bounces(self.h, max_bounces))

    def filter_by_nr_of_refractions(self, max_refractions: int):
        _check(F.lib().opm_rays_filter_by_nr_of_refractions(self.h, max_refractions))

    def filter_energy(self, transmission: float):
        _check(F.lib().opm_rays_filter_energy(self.h, transmission))

    def split(self, ratio: float) -> "Rays":
        out = Rays(self.ctx, max(len(self), 1))
        _check(F.lib().opm_rays_split(self.h, ratio, out.h))
        return out

    def refract_paraxial(self, focal_length: float, iso: Isometry):
        _check(F.lib().opm_rays_refract_paraxial(self.h, focal_length, C.byref(iso.c)))

    def transformed_by_iso(self, iso: Isometry, inverse: bool = False):
        _check(F.lib().opm_rays_transform(self.h, C.byref(iso.c), int(inverse)))

    # ---- statistics
    def spot_stats(self, frame: Isometry | None = None) -> F.OpmSpotStats:
        s = F.OpmSpotStats()
        _check(F.lib().opm_rays_spot_stats(self.h, C.byref(frame.c) if frame is not None else None, C.byref(s)))
        return s

    def spot_partial(self, frame: Isometry | None = None, reduced: "F.OpmSpotPartial | None" = None) -> F.OpmSpotPartial:
        """two-phase statistics for sharded bundles: without `reduced` phase 1 (sums), with the all-reduced sums phase 2
        (radius accumulators about the global centroids); see sharding.spot_stats"""
        fr = C.byref(frame.c) if frame is not None else None
        if reduced is None:
            p = F.OpmSpotPartial()
            _check(F.lib().opm_rays_spot_partial_sums(self.h, fr, C.byref(p)))
            return p
        _check(F.lib().opm_rays_spot_partial_radii(self.h, fr, C.byref(reduced)))
        return reduced

    def total_energy(self) -> float:
        return self.spot_stats().total_energy

    def nr_of_rays(self, valid_only: bool = True) -> int:
        s = self.spot_stats()
        return s.n_valid if valid_only else s.n_total

    def centroid(self):
        s = self.spot_stats()
        return np.array(s.centroid[:]) if s.n_valid else None

    def energy_weighted_centroid(self):
        s = self.spot_stats()
        return np.array(s.energy_centroid[:]) if s.n_valid else None

    def beam_radius_geo(self):
        s = self.spot_stats()
        return s.beam_radius_geo if s.n_valid else None

    def beam_radius_rms(self):
        s = self.spot_stats()
        return s.beam_radius_rms if s.n_valid else None

    def energy_weighted_beam_radius_rms(self):
        s = self.spot_stats()
        return s.energy_beam_radius_rms if s.n_valid else None

    def bounce_lvl(self) -> int:
        out = C.c_uint32()
        _check(F.lib().opm_rays_bounce_lvl(self.h, C.byref(out)))
        return out.value

    def tally(self, n_orders: int = 8):
        counts = (C.c_uint64 * n_orders)()
        energy = (C.c_double * n_orders)()
        _check(F.lib().opm_rays_tally(self.h, n_orders, counts, energy))
        return np.array(counts[:], dtype=np.uint64), np.array(energy[:])


# ---------------------------------------------------------------------------------------------
class Program:
    """A fused surface sequence: ops run back to back in ONE kernel launch (opm_trace_sequence)."""

    def __init__(self):
        self.ops: list[F.OpmOp] = []
        self._keep: list = []

    def refract(self, surface: OpticSurface, index: RefractiveIndex | None, refraction_intended=True,
                strategy=F.MISS_STOP, hit_map: HitMap | None = None, children: Rays | None = None,
                continue_as_child=False) -> "Program":
        op = F.OpmOp(kind=F.OP_REFRACT)
        r = op.u.refract
        r.surface = surface.c
        if index is not None:
            r.index = index.c
            r.has_index = 1
        r.refraction_intended = int(refraction_intended)
        r.strategy = strategy
        r.flags = (F.REFRACT_EMIT_CHILDREN if children is not None else 0) | \
                  (F.REFRACT_CONTINUE_AS_CHILD if continue_as_child else 0)
        r.hits = hit_map.h if hit_map is not None else None
        r.children = children.h if children is not None else None
        self._keep += [hit_map, children]
        self.ops.append(op)
        return self

    def mirror(self, surface: OpticSurface, strategy=F.MISS_STOP, hit_map=None) -> "Program":
        """nodes/thin_mirror.rs:209-250: n2 = None, refraction_intended = false, the reflected bundle continues."""
        return self.refract(surface, None, refraction_intended=False, strategy=strategy, hit_map=hit_map,
                            continue_as_child=True)

    def diffract(self, surface: OpticSurface, index: RefractiveIndex, grating_vector, order: int, refraction_intended=False,
                 children: Rays | None = None, continue_as_child=True) -> "Program":
        """reflective grating (rays.rs:1058-1111); continue_as_child: the diffracted bundle continues (the node's behaviour)"""
        op = F.OpmOp(kind=F.OP_DIFFRACT)
        g = op.u.diffract
        g.surface, g.index = surface.c, index.c
        g.grating_vector[:] = [float(x) for x in grating_vector]
        g.order, g.refraction_intended = int(order), int(refraction_intended)
        g.flags = (F.REFRACT_EMIT_CHILDREN if children is not None else 0) | (F.REFRACT_CONTINUE_AS_CHILD if continue_as_child else 0)
        g.children = children.h if children is not None else None
        self._keep.append(children)
        self.ops.append(op)
        return self

    def apodize(self, aperture: Aperture, iso: Isometry) -> "Program":
        op = F.OpmOp(kind=F.OP_APODIZE)
        op.u.apodize.iso = iso.c
        op.u.apodize.aperture = aperture.c
        self._keep.append(aperture)
        self.ops.append(op)
        return self

    def threshold(self, min_energy: float) -> "Program":
        op = F.OpmOp(kind=F.OP_THRESHOLD)
        op.u.threshold.min_energy = min_energy
        self.ops.append(op)
        return self

    def limits(self, max_bounces: int, max_refractions: int | None) -> "Program":
        op = F.OpmOp(kind=F.OP_LIMITS)
        op.u.limits.max_bounces = max_bounces
        op.u.limits.max_refractions = max_refractions if max_refractions is not None else 0
        op.u.limits.check_refractions = int(max_refractions is not None)
        self.ops.append(op)
        return self

    def filter(self, transmission: float = 1.0, spectrum: "Spectrum | None" = None) -> "Program":
        """FilterType::Constant(transmission) or, with `spectrum`, FilterType::Spectrum (ray.rs:702-727)"""
        op = F.OpmOp(kind=F.OP_FILTER)
        op.u.filter.transmission = transmission
        op.u.filter.spectrum = spectrum.h if spectrum is not None else None
        self._keep.append(spectrum)
        self.ops.append(op)
        return self

    def paraxial(self, focal_length: float, iso: Isometry) -> "Program":
        op = F.OpmOp(kind=F.OP_PARAXIAL)
        op.u.paraxial.focal_length = focal_length
        op.u.paraxial.iso = iso.c
        self.ops.append(op)
        return self

    def transform(self, iso: Isometry, inverse=False) -> "Program":
        op = F.OpmOp(kind=F.OP_TRANSFORM)
        op.u.transform.iso = iso.c
        op.u.transform.inverse = int(inverse)
        self.ops.append(op)
        return self

    def split(self, ratio: float, split_out: Rays | None, spectrum: "Spectrum | None" = None) -> "Program":
        """SplittingConfig::Ratio(ratio) or, with `spectrum`, SplittingConfig::Spectrum (ray.rs:752-772)"""
        op = F.OpmOp(kind=F.OP_SPLIT)
        op.u.split.ratio = ratio
        op.u.split.split_out = split_out.h if split_out is not None else None
        op.u.split.spectrum = spectrum.h if spectrum is not None else None
        self._keep += [split_out, spectrum]
        self.ops.append(op)
        return self

    def fork(self, ratio: float, spectrum: "Spectrum | None" = None) -> "Program":
        """Rays::split whose split-off bundle (E * (1 - ratio)) runs the ops up to the matching join() in the same launch"""
        op = F.OpmOp(kind=F.OP_FORK)
        op.u.fork.ratio = ratio
        op.u.fork.spectrum = spectrum.h if spectrum is not None else None
        self._keep.append(spectrum)
        self.ops.append(op)
        return self

    def join(self, out: Rays | None) -> "Program":
        """stores the split-off bundle (slot-indexed) and resumes the main bundle"""
        op = F.OpmOp(kind=F.OP_JOIN)
        op.u.join.out = out.h if out is not None else None
        self._keep.append(out)
        self.ops.append(op)
        return self

    def snapshot(self, out: Rays, spot_only=False, spot_frame: "Isometry | None" = None, want_spot_sums=False) -> "Program":
        """copy of the bundle at this point; spot_only: position, energy and valid flag only; want_spot_sums: the launch takes the
        phase-1 sums of out.spot_stats(spot_frame) on the way (specialised kernel)"""
        op = F.OpmOp(kind=F.OP_SNAPSHOT)
        op.u.snapshot.out = out.h
        op.u.snapshot.fields = F.SNAPSHOT_SPOT if spot_only else F.SNAPSHOT_ALL
        op.u.snapshot.want_spot_sums = 1 if want_spot_sums else 0
        if spot_frame is not None:
            op.u.snapshot.spot_frame = C.addressof(spot_frame.c)
            self._keep.append(spot_frame)
        self._keep.append(out)
        self.ops.append(op)
        return self

    def run(self, rays: Rays, out: Rays | None = None, compact=False, strict=False, sync=True) -> TraceStats | None:
        arr = (F.OpmOp * max(len(self.ops), 1))(*self.ops)
        flags = (F.TRACE_COMPACT if compact else 0) | (F.TRACE_STRICT if strict else 0)
        st = F.OpmTraceStats()
        _check(F.lib().opm_trace_sequence(rays.h, arr, len(self.ops), flags, out.h if out is not None else None,
                                          C.byref(st) if sync else None))
        return TraceStats.of(st) if sync else None


# ---------------------------------------------------------------------------------------------
@dataclass
class FluenceData:
    """nodes/fluence_detector/fluence_data.rs `FluenceData` (Binning estimator)."""
    map: np.ndarray          # [ny, nx] (rows = y like nalgebra's DMatrix)
    lrtb: tuple              # left, right, top, bottom
    peak: float
    total_energy: float


def fluence_binning(ctx: Context, hit_maps, nx: int, ny: int, ranges=None, lrtb=None, bounce=-1,
                    map_dev: int | None = None, accumulate=False, download=True) -> FluenceData:
    """HitMap::calc_fluence_map(.., FluenceEstimator::Binning) (hit_map/mod.rs:363-379, rays_hit_map.rs:330-391)."""
    L = F.lib()
    own = map_dev is None
    if own:
        map_dev = ctx.device_alloc(nx * ny * 8)
    try:
        arr = (C.c_void_p * len(hit_maps))(*[h.h for h in hit_maps])
        rg = (C.c_double * 4)(*ranges) if ranges is not None else None
        ov = (C.c_double * 4)(*lrtb) if lrtb is not None else None
        out = (C.c_double * 4)()
        _check(L.opm_fluence_binning(arr, len(hit_maps), bounce, nx, ny, rg, ov, int(accumulate), C.c_void_p(map_dev), out))
        peak, tot = C.c_double(), C.c_double()
        _check(L.opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, ny, out, C.byref(peak), C.byref(tot)))
        m = None
        if download:
            flat = np.zeros(nx * ny)
            ctx.to_host(map_dev, flat)
            m = flat.reshape(nx, ny).T.copy()  # column-major ny x nx
        return FluenceData(m, tuple(out[:]), peak.value, tot.value)
    finally:
        if own:
            ctx.device_free(map_dev)


def fluence_kde(ctx: Context, hit_maps, nx: int, ny: int, ranges=None, bounce=-1):
    """HitMap::calc_fluence_map(.., FluenceEstimator::KDE) (rays_hit_map.rs:575-613, kde/mod.rs) -> (FluenceData, band width)"""
    L = F.lib()
    map_dev = ctx.device_alloc(nx * ny * 8)
    try:
        arr = (C.c_void_p * len(hit_maps))(*[h.h for h in hit_maps])
        rg = (C.c_double * 4)(*ranges) if ranges is not None else None
        out = (C.c_double * 4)()
        bw = C.c_double()
        _check(L.opm_fluence_kde(arr, len(hit_maps), bounce, nx, ny, rg, C.c_void_p(map_dev), out, C.byref(bw)))
        peak, tot = C.c_double(), C.c_double()
        _check(L.opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, ny, out, C.byref(peak), C.byref(tot)))
        flat = np.zeros(nx * ny)
        ctx.to_host(map_dev, flat)
        return FluenceData(flat.reshape(nx, ny).T.copy(), tuple(out[:]), peak.value, tot.value), bw.value
    finally:
        ctx.device_free(map_dev)


def hitmap_bounce_levels(hit_map) -> list:
    """the bounce counts under which the hits of a slot-indexed hit map are filed (hit_map/mod.rs:101-115), below 64"""
    m = C.c_uint64()
    _check(F.lib().opm_hitmap_bounce_levels(hit_map.h, C.byref(m)))
    return [b for b in range(64) if (m.value >> b) & 1]


def fluence_voronoi(ctx: Context, hit_maps, nx: int, ranges=None, bounce=-1, map_dev: int | None = None, accumulate=False) -> FluenceData:
    """RaysHitMap::calc_fluence_with_voronoi (rays_hit_map.rs:407-521) of ONE rays hit map (the hits passing `bounce`): nx x nx map
    (the reference gives both axes nr_of_points.0 points), ranges = (ax1.start, ax1.end, ax2.start, ax2.end) or None."""
    L = F.lib()
    own = map_dev is None
    if own:
        map_dev = ctx.device_alloc(nx * nx * 8)
    try:
        arr = (C.c_void_p * len(hit_maps))(*[h.h for h in hit_maps])
        rg = (C.c_double * 4)(*ranges) if ranges is not None else None
        out = (C.c_double * 4)()
        _check(L.opm_fluence_voronoi(arr, len(hit_maps), bounce, nx, rg, int(accumulate), C.c_void_p(map_dev), out))
        peak, tot = C.c_double(), C.c_double()
        _check(L.opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, nx, out, C.byref(peak), C.byref(tot)))
        flat = np.zeros(nx * nx)
        ctx.to_host(map_dev, flat)
        return FluenceData(flat.reshape(nx, nx).T.copy(), tuple(out[:]), peak.value, tot.value)
    finally:
        if own:
            ctx.device_free(map_dev)


def fluence_helper_rays(ctx: Context, hit_maps, nx: int, ranges=None, bounce=-1, map_dev: int | None = None, accumulate=False) -> FluenceData:
    """RaysHitMap::calc_fluence_with_helper_rays (rays_hit_map.rs:627-722) of ONE rays hit map whose records carry fluences (the
    trace of a bundle with fluence helper rays wrote them); arguments as fluence_voronoi."""
    L = F.lib()
    own = map_dev is None
    if own:
        map_dev = ctx.device_alloc(nx * nx * 8)
    try:
        arr = (C.c_void_p * len(hit_maps))(*[h.h for h in hit_maps])
        rg = (C.c_double * 4)(*ranges) if ranges is not None else None
        out = (C.c_double * 4)()
        _check(L.opm_fluence_helper_rays(arr, len(hit_maps), bounce, nx, rg, int(accumulate), C.c_void_p(map_dev), out))
        peak, tot = C.c_double(), C.c_double()
        _check(L.opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, nx, out, C.byref(peak), C.byref(tot)))
        flat = np.zeros(nx * nx)
        ctx.to_host(map_dev, flat)
        return FluenceData(flat.reshape(nx, nx).T.copy(), tuple(out[:]), peak.value, tot.value)
    finally:
        if own:
            ctx.device_free(map_dev)


def fluence_voronoi_combined(ctx: Context, bundles, nx: int, ny: int) -> FluenceData:
    """HitMap::calc_combined_fluence_with_voronoi (hit_map/mod.rs:232-262): `bundles` = one list of hit maps per ray bundle (uuid)
    of the surface; every (bundle, bounce level) is a rays hit map of its own, estimated on the common bounding box
    (get_bounding_box, :178-203) and summed.  The reference adds nx x nx estimates to an nx x ny matrix and panics unless
    nx == ny; so does this (ValueError)."""
    if nx != ny:
        raise ValueError("calc_combined_fluence_with_voronoi: matrix dimensions mismatch (the reference panics unless nx == ny)")
    every = [h for b in bundles for h in b]
    l, r, t, b_ = hitmaps_bounding_box(every)[0]
    map_dev = ctx.device_alloc(nx * nx * 8)
    try:
        ctx.device_memset(map_dev, 0, nx * nx * 8)
        for maps in bundles:
            levels = sorted({lv for h in maps for lv in hitmap_bounce_levels(h)})
            for lv in levels:
                fluence_voronoi(ctx, maps, nx, ranges=(l, r, b_, t), bounce=lv, map_dev=map_dev, accumulate=True)
        out = (C.c_double * 4)(l, r, t, b_)
        peak, tot = C.c_double(), C.c_double()
        _check(F.lib().opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, nx, out, C.byref(peak), C.byref(tot)))
        flat = np.zeros(nx * nx)
        ctx.to_host(map_dev, flat)
        return FluenceData(flat.reshape(nx, nx).T.copy(), (l, r, t, b_), peak.value, tot.value)
    finally:
        ctx.device_free(map_dev)


def fluence_peak_total(ctx: Context, map_dev: int, nx: int, ny: int, lrtb) -> FluenceData:
    """FluenceData::new peak (fluence_data.rs:45-68) and total_energy (:130-141) of a map that already sits on the device
    (e.g. after an all-reduce of per-GPU maps)."""
    box = (C.c_double * 4)(*lrtb)
    peak, tot = C.c_double(), C.c_double()
    _check(F.lib().opm_fluence_peak_total(ctx.h, C.c_void_p(map_dev), nx, ny, box, C.byref(peak), C.byref(tot)))
    return FluenceData(None, tuple(lrtb), peak.value, tot.value)


def hitmaps_bounding_box(hit_maps, bounce=-1):
    arr = (C.c_void_p * len(hit_maps))(*[h.h for h in hit_maps])
    out = (C.c_double * 4)()
    n = C.c_uint64()
    _check(F.lib().opm_hitmaps_bounding_box(arr, len(hit_maps), bounce, out, C.byref(n)))
    return tuple(out[:]), n.value


# ---------------------------------------------------------------------------------------------
def spectrum_gaussian(lo: float, hi: float, num: int, mu: float, fwhm: float, power: float = 1.0):
    """spectral_distribution/gaussian.rs:78-96: (wavelengths, normalised weights); host math in the library"""
    wv, w = np.zeros(num), np.zeros(num)
    dp = C.POINTER(C.c_double)
    _check(F.lib().opm_spectrum_gaussian(lo, hi, num, mu, fwhm, power, wv.ctypes.data_as(dp), w.ctypes.data_as(dp)))
    return wv, w


class SourceDesc:
    """Deterministic collimated source generated on the device (rays.rs:264-297)."""

    def __init__(self, n_positions: int, wavelengths, weights=None, total_energy=1.0, iso: Isometry | None = None):
        self.c = F.OpmSourceDesc()
        self.c.n_positions = n_positions
        self.c.total_energy = total_energy
        self._w = np.ascontiguousarray(np.atleast_1d(wavelengths), dtype=np.float64)
        self._g = np.ascontiguousarray(np.atleast_1d(weights if weights is not None else [1.0]), dtype=np.float64)
        self.c.n_wavelengths = len(self._w)
        self.c.wavelengths = self._w.ctypes.data_as(C.POINTER(C.c_double))
        self.c.weights = self._g.ctypes.data_as(C.POINTER(C.c_double))
        self.c.iso = (iso or Isometry.identity()).c
        self.c.power = 1.0
        self.c.sigma[:] = [1.0, 1.0]

    def grid(self, nx, ny, side_x, side_y):
        self.c.pos_dist, self.c.grid_nx, self.c.grid_ny, self.c.size_x, self.c.size_y = F.POS_GRID, nx, ny, side_x, side_y
        return self

    def hexapolar(self, rings: int, radius: float):
        """Hexapolar (position_distributions/hexapolar.rs): n_positions must be 1 + 3 rings (rings + 1)"""
        self.c.pos_dist, self.c.grid_nx, self.c.size_x = F.POS_HEXAPOLAR, rings, radius
        return self

    def fibonacci_rectangle(self, sx, sy):
        self.c.pos_dist, self.c.size_x, self.c.size_y = F.POS_FIBONACCI_RECT, sx, sy
        return self

    def fibonacci_ellipse(self, rx, ry):
        self.c.pos_dist, self.c.size_x, self.c.size_y = F.POS_FIBONACCI_ELLIPSE, rx, ry
        return self

    def general_gaussian(self, mu=(0.0, 0.0), sigma=(1.0, 1.0), power=1.0, theta=0.0, rectangular=False, norm=0.0):
        self.c.energy_dist = F.EDIST_GENERAL_GAUSSIAN
        self.c.mu[:], self.c.sigma[:] = list(mu), list(sigma)
        self.c.power, self.c.theta, self.c.rectangular, self.c.energy_norm = power, theta, int(rectangular), norm
        return self

    def energy_norm(self, ctx: Context, begin=0, end=None) -> float:
        out = C.c_double()
        _check(F.lib().opm_source_energy_norm(ctx.h, C.byref(self.c), begin, end if end is not None else self.c.n_positions,
                                              C.byref(out)))
        return out.value

    def generate_strided(self, ctx: Context, first: int, count: int, stride: int, rays: Rays | None = None) -> Rays:
        """positions first, first + stride, ...: the load-balanced shard (first = rank, stride = world size)"""
        r = rays if rays is not None else Rays(ctx, max(count * self.c.n_wavelengths, 1))
        _check(F.lib().opm_source_generate_strided(r.h, C.byref(self.c), first, count, stride))
        return r

    def generate(self, ctx: Context, begin=0, end=None, rays: Rays | None = None) -> Rays:
        end = end if end is not None else self.c.n_positions
        r = rays if rays is not None else Rays(ctx, max((end - begin) * self.c.n_wavelengths, 1))
        _check(F.lib().opm_source_generate(r.h, C.byref(self.c), begin, end))
        return r
