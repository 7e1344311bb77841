"""Python veneer over the C++ scene layer (include/opossum_b200_scene.h): node constructors named after the
reference's nodes (opossum/src/nodes/*) and the two analyzers.  Marshalling only -- the node -> surface
compilation and the analysis drivers live in opossum_b200/csrc/opm_scene.cpp."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _ffi as F
from .api import (Aperture, Context, FluenceData, HitMap, Isometry, Rays, RayTraceConfig, RefractiveIndex, TraceStats,
                  _check)


@dataclass
class CoatingType:
    """coatings/mod.rs `CoatingType`"""
    type: int = F.COAT_IDEAL_AR
    reflectivity: float = 0.0

    @staticmethod
    def ideal_ar():
        return CoatingType(F.COAT_IDEAL_AR)

    @staticmethod
    def constant_r(reflectivity: float):
        return CoatingType(F.COAT_CONSTANT_R, reflectivity)

    @staticmethod
    def fresnel():
        return CoatingType(F.COAT_FRESNEL)


class Node:
    """Common part of an optic node: isometry, alignment, coatings and apertures of input_1 / output_1."""

    def __init__(self, node_type: int, iso: Isometry | None = None):
        self.c = F.OpmNodeDesc()
        F.lib().opm_node_desc_init(C.byref(self.c), node_type)
        if iso is not None:
            self.c.iso = iso.c
        self._keep = []

    def with_alignment(self, decenter, tilt):
        """OpticNode::set_alignment (optic_node.rs:383-393)"""
        self.c.has_alignment = 1
        self.c.alignment = Isometry.new(decenter, tilt).c
        return self

    def with_coatings(self, input_1: CoatingType | None = None, output_1: CoatingType | None = None):
        if input_1 is not None:
            self.c.coating_in.type, self.c.coating_in.reflectivity = input_1.type, input_1.reflectivity
        if output_1 is not None:
            self.c.coating_out.type, self.c.coating_out.reflectivity = output_1.type, output_1.reflectivity
        return self

    def with_apertures(self, input_1: Aperture | None = None, output_1: Aperture | None = None):
        if input_1 is not None:
            self.c.aperture_in = input_1.c
            self._keep.append(input_1)
        if output_1 is not None:
            self.c.aperture_out = output_1.c
            self._keep.append(output_1)
        return self

    def with_lidt(self, lidt_j_per_m2: float):
        self.c.lidt = lidt_j_per_m2
        return self

    def with_spectrum(self, spectrum):
        """IdealFilter: FilterType::Spectrum; BeamSplitter: SplittingConfig::Spectrum (api.Spectrum, kept alive by the node)"""
        self._spectrum = spectrum
        self.c.spectrum = spectrum.h if spectrum is not None else None
        return self

    def fed_by(self, parent: int, port: int = 0):
        self.c.parent, self.c.parent_port = parent, port
        return self


def Source(iso=None):
    return Node(F.NODE_SOURCE, iso)


def Lens(front_curvature, rear_curvature, center_thickness, refractive_index: RefractiveIndex, iso=None):
    """Lens::new (nodes/lens/mod.rs:98-140); curvature +-inf = flat"""
    n = Node(F.NODE_LENS, iso)
    n.c.p[0], n.c.p[1], n.c.p[2] = front_curvature, rear_curvature, center_thickness
    n.c.index = refractive_index.c
    return n


def CylindricLens(front_curvature, rear_curvature, center_thickness, refractive_index: RefractiveIndex, iso=None):
    n = Node(F.NODE_CYLINDRIC_LENS, iso)
    n.c.p[0], n.c.p[1], n.c.p[2] = front_curvature, rear_curvature, center_thickness
    n.c.index = refractive_index.c
    return n


def Wedge(center_thickness, wedge_angle, refractive_index: RefractiveIndex, iso=None):
    n = Node(F.NODE_WEDGE, iso)
    n.c.p[0], n.c.p[1] = center_thickness, wedge_angle
    n.c.index = refractive_index.c
    return n


def ThinMirror(curvature=math.inf, iso=None):
    n = Node(F.NODE_THIN_MIRROR, iso)
    n.c.p[0] = curvature
    return n


def ParabolicMirror(focal_length, iso=None, oa_angle=0.0, oa_direction=(1.0, 0.0), collimating=False):
    """ParabolicMirror::new / new_with_off_axis (nodes/parabolic_mirror.rs:88-150)"""
    n = Node(F.NODE_PARABOLIC_MIRROR, iso)
    n.c.p[0], n.c.p[1], n.c.p[2], n.c.p[3], n.c.p[4] = focal_length, oa_angle, oa_direction[0], oa_direction[1], 1.0 if collimating else 0.0
    return n


def BeamSplitter(ratio, iso=None):
    n = Node(F.NODE_BEAM_SPLITTER, iso)
    n.c.p[0] = ratio
    return n


def Dummy(iso=None):
    return Node(F.NODE_DUMMY, iso)


def IdealFilter(transmission, iso=None):
    n = Node(F.NODE_IDEAL_FILTER, iso)
    n.c.p[0] = transmission
    return n


def ParaxialSurface(focal_length, iso=None):
    n = Node(F.NODE_PARAXIAL_SURFACE, iso)
    n.c.p[0] = focal_length
    return n


def ReflectiveGrating(line_density_per_m=1740.0e3, diffraction_order=-1, iso=None):
    """ReflectiveGrating::new (nodes/reflective_grating.rs:84-105): line density in 1/m, grating vector along the node's x"""
    if not (math.isfinite(line_density_per_m) and line_density_per_m > 0.0):
        raise ValueError("Only positive finite values are allowed for a grating line density")
    n = Node(F.NODE_REFLECTIVE_GRATING, iso)
    n.c.p[0], n.c.p[1] = line_density_per_m, float(diffraction_order)
    return n


def littrow_angle(line_density_per_m, diffraction_order, wavelength):
    """with_rot_from_littrow (nodes/reflective_grating.rs:107-128): tilt about y that puts `wavelength` in Littrow"""
    return math.asin(diffraction_order * wavelength * line_density_per_m / 2.0)


def SpotDiagram(iso=None):
    return Node(F.NODE_SPOT_DIAGRAM, iso)


def FluenceDetector(iso=None):
    return Node(F.NODE_FLUENCE_DETECTOR, iso)


def Spectrometer(iso=None):
    """nodes/spectrometer.rs: detector whose report is Rays::to_spectrum(0.2 nm) of the bundle it keeps (Scene.node_spectrum)"""
    return Node(F.NODE_SPECTROMETER, iso)


def WaveFront(iso=None):
    """nodes/wavefront.rs: detector whose report is the wavefront error map with PtV and RMS (Scene.node_wavefront)"""
    return Node(F.NODE_WAVEFRONT, iso)


def EnergyMeter(iso=None):
    """nodes/energy_meter.rs: a detector whose report is the total energy of the rays that reach it"""
    return Node(F.NODE_ENERGY_METER, iso)


def RayPropagationVisualizer(iso=None):
    """nodes/ray_propagation_visualizer.rs: a single-surface detector that keeps the bundle; after
    raytrace(position_history=True) its rays (Scene.node_rays) carry their paths (Rays.position_history)"""
    return Node(F.NODE_RAY_PROPAGATION_VISUALIZER, iso)


class _BorrowedRays(Rays):
    """A bundle owned by the scene (never destroyed from Python)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, C.c_void_p(handle)

    def close(self):
        self.h = None


class _BorrowedHitMap(HitMap):
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, C.c_void_p(handle)

    def close(self):
        self.h = None


class Scene:
    """A chain of optic nodes starting at a Source (NodeGroup of the reference, restricted to explicit isometries)."""
    _close_order = 0  # a closing context destroys scenes before stand-alone bundles

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        _check(F.lib().opm_scene_create(ctx.h, C.byref(h)))
        self.h = h
        self._nodes = []
        ctx._adopt(self)

    def set_ambient(self, index: RefractiveIndex):
        _check(F.lib().opm_scene_set_ambient(self.h, C.byref(index.c)))

    def add_node(self, node: Node, distance: float | None = None, align_like: tuple | None = None) -> int:
        """NodeGroup::add_node (+ connect_nodes to the node named by `fed_by`, default the previous one).  distance = None: the
        node sits at its own isometry; else it has none and `calc_node_positions` places it `distance` [m] behind its
        predecessor along the optical axis (NodeGroup::connect_nodes(.., distance))"""
        idx = C.c_int32()
        if align_like is not None:  # (node index, distance): OpticNode::align_like_node_at_distance (optic_node.rs:476-480)
            _check(F.lib().opm_scene_add_node_aligned_like(self.h, C.byref(node.c), int(align_like[0]), float(align_like[1]),
                                                           float(distance or 0.0), C.byref(idx)))
        elif distance is None:
            _check(F.lib().opm_scene_add_node(self.h, C.byref(node.c), C.byref(idx)))
        else:
            _check(F.lib().opm_scene_add_node_at_distance(self.h, C.byref(node.c), float(distance), C.byref(idx)))
        self._nodes.append(node)
        return idx.value

    def calc_node_positions(self, axis_wavelength: float, config: RayTraceConfig | None = None):
        """AnalysisRayTrace::calc_node_positions of the node group (nodes/node_group/analysis_raytrace.rs:92-212): places
        every node that was added with a distance; `axis_wavelength` [m] is the source's alignment wavelength (or the
        energy-weighted centre wavelength of its rays, rays.rs:1166-1175)"""
        cfg = None
        if config is not None:
            cfg = F.OpmRayTraceConfig()
            cfg.min_energy_per_ray, cfg.max_number_of_bounces = config.min_energy_per_ray, config.max_number_of_bounces
            cfg.max_number_of_refractions, cfg.missed_surface_strategy = config.max_number_of_refractions, config.missed_surface_strategy
        _check(F.lib().opm_scene_calc_node_positions(self.h, float(axis_wavelength), C.byref(cfg) if cfg is not None else None))

    def node_isometry(self, node: int) -> Isometry:
        """OpticNode::isometry()"""
        o = F.OpmIsometry()
        _check(F.lib().opm_scene_node_isometry(self.h, node, C.byref(o)))
        return Isometry(o)

    def raytrace(self, source_rays: Rays, config: RayTraceConfig | None = None, strict=False, per_surface=False,
                 record_all_hits=False, snapshots=True, sync=True, keep_source=False, position_history=False, keep_light_data=False,
                 lean_spot_data=False) -> TraceStats | None:
        """RayTracingAnalyzer::analyze (analyzers/raytrace.rs:36-49); traces `source_rays` in place.  lean_spot_data: SpotDiagram nodes
        keep only position, energy and valid flag of their copy of the bundle (what their report reads)."""
        cfg = F.OpmRayTraceConfig()
        c = config or RayTraceConfig()
        cfg.min_energy_per_ray, cfg.max_number_of_bounces = c.min_energy_per_ray, c.max_number_of_bounces
        cfg.max_number_of_refractions, cfg.missed_surface_strategy = c.max_number_of_refractions, c.missed_surface_strategy
        flags = ((F.SCENE_STRICT if strict else 0) | (F.SCENE_PER_SURFACE if per_surface else 0) |
                 (F.SCENE_RECORD_ALL_HITS if record_all_hits else 0) | (0 if snapshots else F.SCENE_NO_SNAPSHOTS) |
                 (0 if sync else F.SCENE_ASYNC) | (F.SCENE_KEEP_SOURCE if keep_source else 0) |
                 (F.SCENE_POSITION_HISTORY if position_history else 0) | (F.SCENE_KEEP_LIGHT_DATA if keep_light_data else 0) |
                 (F.SCENE_LEAN_SPOT_DATA if lean_spot_data else 0))
        st = F.OpmTraceStats()
        _check(F.lib().opm_scene_raytrace(self.h, source_rays.h, C.byref(cfg), flags, C.byref(st)))
        return TraceStats.of(st) if sync else None

    def output_rays(self) -> Rays | None:
        """output bundle of the main chain after raytrace(keep_source=True)"""
        h = F.lib().opm_scene_output_rays(self.h)
        return _BorrowedRays(self.ctx, h) if h else None

    def node_rays(self, node: int) -> Rays | None:
        h = F.lib().opm_scene_node_rays(self.h, node)
        return _BorrowedRays(self.ctx, h) if h else None

    def node_spot(self, node: int) -> F.OpmSpotStats:
        s = F.OpmSpotStats()
        _check(F.lib().opm_scene_node_spot(self.h, node, C.byref(s)))
        return s

    def node_spot_partial(self, node: int, reduced: "F.OpmSpotPartial | None" = None) -> F.OpmSpotPartial:
        """SpotDiagram report in two phases for a bundle sharded over GPUs (sharding.spot_stats drives it)"""
        p = reduced if reduced is not None else F.OpmSpotPartial()
        _check(F.lib().opm_scene_node_spot_partial(self.h, node, 1 if reduced is None else 2, C.byref(p)))
        return p

    def node_fluence(self, node: int, nx=100, ny=83, download=True, estimator="binning") -> FluenceData:
        """FluenceDetector::node_report (nodes/fluence_detector/mod.rs:91-135); estimator "binning", "kde" or "voronoi"
        (FluenceEstimator::Binning / ::KDE / ::Voronoi, the reference's default: nx x nx map, nx must equal ny)"""
        m = np.zeros(nx * ny) if download else None
        lrtb = (C.c_double * 4)()
        peak, tot = C.c_double(), C.c_double()
        if estimator == "kde":
            bw = C.c_double()
            _check(F.lib().opm_scene_node_fluence_kde(self.h, node, nx, ny, m.ctypes.data_as(C.POINTER(C.c_double)) if download else None,
                                                      lrtb, C.byref(peak), C.byref(tot), C.byref(bw)))
            fd = FluenceData(m.reshape(nx, ny).T.copy() if download else None, tuple(lrtb[:]), peak.value, tot.value)
            fd.band_width = bw.value
            return fd
        if estimator in ("voronoi", "helper_rays"):
            fn = F.lib().opm_scene_node_fluence_voronoi if estimator == "voronoi" else F.lib().opm_scene_node_fluence_helper_rays
            _check(fn(self.h, node, nx, ny, m.ctypes.data_as(C.POINTER(C.c_double)) if download else None, lrtb, C.byref(peak), C.byref(tot)))
            return FluenceData(m.reshape(nx, ny).T.copy() if download else None, tuple(lrtb[:]), peak.value, tot.value)
        if estimator != "binning":
            raise ValueError("estimator must be 'binning', 'kde', 'voronoi' or 'helper_rays'")
        _check(F.lib().opm_scene_node_fluence(self.h, node, nx, ny,
                                              m.ctypes.data_as(C.POINTER(C.c_double)) if download else None, lrtb,
                                              C.byref(peak), C.byref(tot)))
        return FluenceData(m.reshape(nx, ny).T.copy() if download else None, tuple(lrtb[:]), peak.value, tot.value)

    def node_energy(self, node: int) -> float:
        """EnergyMeter::node_report (nodes/energy_meter.rs:136-151)"""
        e = C.c_double()
        _check(F.lib().opm_scene_node_energy(self.h, node, C.byref(e)))
        return e.value

    def node_spectrum(self, node: int, resolution: float = 0.0, wavelength_range=None):
        """Spectrometer::node_report (nodes/spectrometer.rs:155-175) -> (slots [um], values [1/um])"""
        rng = (C.c_double * 2)(*wavelength_range) if wavelength_range is not None else None
        n = C.c_uint64()
        _check(F.lib().opm_scene_node_spectrum(self.h, node, resolution, rng, None, None, 0, C.byref(n)))
        lam, val = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
        dp = C.POINTER(C.c_double)
        _check(F.lib().opm_scene_node_spectrum(self.h, node, resolution, rng, lam.ctypes.data_as(dp), val.ctypes.data_as(dp), n.value,
                                               C.byref(n)))
        return lam[: n.value], val[: n.value]

    def node_wavefront(self, node: int, want_map: bool = True):
        """WaveFront::node_report (nodes/wavefront.rs:173-235) -> (OpmWavefrontStats, (n_valid, 3) rows x mm, y mm, error in wavelengths)"""
        st = F.OpmWavefrontStats()
        r = self.node_rays(node)
        n = len(r) if r is not None else 0
        xyw = np.zeros((max(n, 1), 3)) if want_map else None
        _check(F.lib().opm_scene_node_wavefront(self.h, node, xyw.ctypes.data_as(C.POINTER(C.c_double)) if want_map else None,
                                                max(n, 1), C.byref(st)))
        return st, (xyw[: st.n_rays] if want_map else None)

    def node_hitmaps(self, node: int, surface: int = 0):
        n = F.lib().opm_scene_node_hitmaps(self.h, node, surface, None, 0)
        arr = (C.c_void_p * max(n, 1))()
        F.lib().opm_scene_node_hitmaps(self.h, node, surface, arr, n)
        return [_BorrowedHitMap(self.ctx, arr[k]) for k in range(n)]

    def beam_splitter_two_inputs(self, node: int, in1: "Rays | None", in2: "Rays | None", cfg=None):
        """BeamSplitter::analyze_raytrace with both inputs (nodes/beam_splitter/mod.rs:158-293) -> (out1, out2)"""
        n = max(len(in1) if in1 is not None else 0, len(in2) if in2 is not None else 0, 1)
        out1, out2 = Rays(self.ctx, 2 * n), Rays(self.ctx, 2 * n)
        _check(F.lib().opm_scene_beam_splitter_two_inputs(self.h, node, in1.h if in1 is not None else None, in2.h if in2 is not None else None,
                                                          C.byref(cfg.c) if cfg is not None else None, out1.h, out2.h))
        return out1, out2

    # ---- ghost focus
    def ghostfocus(self, source_rays: Rays, max_bounces: int = 1, strict=False, estimator="binning") -> F.OpmGhostFocusStats:
        """GhostFocusAnalyzer::analyze (analyzers/ghostfocus.rs:82-122); estimator of the per-bundle LIDT check: "binning" or
        "voronoi" (the reference's default)"""
        cfg = F.OpmGhostFocusConfig(max_bounces=max_bounces, fluence_estimator={"binning": 0, "voronoi": 1}[estimator])
        st = F.OpmGhostFocusStats()
        _check(F.lib().opm_scene_ghostfocus(self.h, source_rays.h, C.byref(cfg), F.SCENE_STRICT if strict else 0, C.byref(st)))
        return st

    def gf_pass_tally(self, gf_pass: int, n_orders: int = 8):
        counts, energy = (C.c_uint64 * n_orders)(), (C.c_double * n_orders)()
        nb, nr = C.c_uint64(), C.c_uint64()
        _check(F.lib().opm_scene_gf_pass_tally(self.h, gf_pass, n_orders, counts, energy, C.byref(nb), C.byref(nr)))
        return np.array(counts[:], dtype=np.uint64), np.array(energy[:]), nb.value, nr.value

    def gf_pass_bundles(self, gf_pass: int):
        n = F.lib().opm_scene_gf_pass_bundles(self.h, gf_pass, None, 0)
        arr = (C.c_void_p * max(n, 1))()
        F.lib().opm_scene_gf_pass_bundles(self.h, gf_pass, arr, n)
        return [_BorrowedRays(self.ctx, arr[k]) for k in range(n)]

    def node_critical_fluences(self, node: int, surface: int = 0):
        n = F.lib().opm_scene_node_critical_fluences(self.h, node, surface, None, 0)
        arr = (F.OpmCriticalFluence * max(n, 1))()
        F.lib().opm_scene_node_critical_fluences(self.h, node, surface, arr, n)
        return [(arr[k].bundle_uuid, arr[k].peak_fluence, arr[k].bounce) for k in range(n)]

    def node_peak_fluence(self, node: int, surface: int = 0) -> float:
        p = C.c_double()
        _check(F.lib().opm_scene_node_peak_fluence(self.h, node, surface, C.byref(p)))
        return p.value

    def reset_data(self):
        _check(F.lib().opm_scene_reset_data(self.h))

    def close(self):
        if self.h:
            F.lib().opm_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
