"""ctypes declarations of the C ABI in include/opossum_b200.h (1:1, no logic).

The shared library is built in-tree by `opossum_b200/csrc/Makefile` (nvcc, sm_100a).  There is no
fallback: if the library is missing or no B200 is visible, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["OPM_B200_LIB"]) if os.environ.get("OPM_B200_LIB") else PKG / "lib" / "libopossum_b200.so"  # override: tuning builds

OPM_MAX_APERTURE_ITEMS = 4
OPM_MAX_PROGRAM_OPS = 128

# status codes
(OPM_OK, OPM_ERR_N2_INVALID, OPM_ERR_WVL_RANGE, OPM_ERR_INDEX_MODEL, OPM_ERR_COATING, OPM_ERR_APODIZE_FACTOR,
 OPM_ERR_THRESHOLD, OPM_ERR_HITPOINT, OPM_ERR_SPLIT_RATIO, OPM_ERR_FOCAL_LENGTH, OPM_ERR_EMPTY_HITMAP,
 OPM_ERR_BIN_RANGE, OPM_ERR_ARG, OPM_ERR_CAPACITY, OPM_ERR_CUDA, OPM_ERR_KDE_BANDWIDTH,
 OPM_ERR_SPECTRUM, OPM_ERR_EMPTY_WAVEFRONT, OPM_ERR_PEER_TIMEOUT, OPM_ERR_VORONOI) = range(20)

GEOM_PLANE, GEOM_SPHERE, GEOM_CYLINDER, GEOM_PARABOLA = range(4)
COAT_IDEAL_AR, COAT_CONSTANT_R, COAT_FRESNEL = range(3)
IDX_CONST, IDX_SELLMEIER1, IDX_SCHOTT, IDX_CONRADY = range(4)
MISS_STOP, MISS_IGNORE = 0, 1
AP_NONE, AP_CIRCLE, AP_RECT, AP_POLYGON, AP_GAUSSIAN = range(5)
(OP_REFRACT, OP_APODIZE, OP_THRESHOLD, OP_LIMITS, OP_FILTER, OP_PARAXIAL, OP_SPLIT, OP_SNAPSHOT, OP_TRANSFORM, OP_DIFFRACT, OP_FORK,
 OP_JOIN) = range(1, 13)
REFRACT_EMIT_CHILDREN, REFRACT_CONTINUE_AS_CHILD = 1, 2
TRACE_COMPACT, TRACE_STRICT = 1, 2
POS_GRID, POS_FIBONACCI_RECT, POS_FIBONACCI_ELLIPSE, POS_HEXAPOLAR = range(4)
EDIST_UNIFORM, EDIST_GENERAL_GAUSSIAN = 0, 1


class OpmRay(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("dir", C.c_double * 3), ("e", C.c_double), ("wvl", C.c_double),
                ("opl", C.c_double), ("n", C.c_double), ("bounces", C.c_uint32), ("refr", C.c_uint32),
                ("valid", C.c_uint32), ("id", C.c_uint32)]


class OpmIso(C.Structure):
    _fields_ = [("rot", C.c_double * 9), ("t", C.c_double * 3)]


class OpmIsometry(C.Structure):
    _fields_ = [("fwd", OpmIso), ("inv", OpmIso)]


class OpmSurface(C.Structure):
    _fields_ = [("geom", C.c_int32), ("coating", C.c_int32), ("geom_param", C.c_double), ("coating_r", C.c_double),
                ("iso", OpmIsometry)]


class OpmIndexModel(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double * 6), ("wvl_lo", C.c_double),
                ("wvl_hi", C.c_double)]


class OpmApertureItem(C.Structure):
    _fields_ = [("type", C.c_int32), ("obstruction", C.c_int32), ("p", C.c_double * 4)]


class OpmAperture(C.Structure):
    _fields_ = [("n_items", C.c_int32), ("is_stack", C.c_int32), ("stack_obstruction", C.c_int32),
                ("n_tris", C.c_int32), ("items", OpmApertureItem * OPM_MAX_APERTURE_ITEMS),
                ("tris", C.POINTER(C.c_double))]


class OpmRayTraceConfig(C.Structure):
    _fields_ = [("min_energy_per_ray", C.c_double), ("max_number_of_bounces", C.c_uint64),
                ("max_number_of_refractions", C.c_uint64), ("missed_surface_strategy", C.c_int32), ("_pad", C.c_int32)]


class OpmTraceStats(C.Structure):
    _fields_ = [("interactions", C.c_uint64), ("children", C.c_uint64), ("hits", C.c_uint64), ("missed", C.c_uint64),
                ("apodized", C.c_uint64), ("valid_out", C.c_uint64), ("alive_out", C.c_uint64),
                ("status_bits", C.c_uint32), ("_pad", C.c_uint32), ("first_valid_key", C.c_uint64)]


class OpmRefractOp(C.Structure):
    _fields_ = [("surface", OpmSurface), ("index", OpmIndexModel), ("has_index", C.c_int32),
                ("refraction_intended", C.c_int32), ("strategy", C.c_int32), ("flags", C.c_int32),
                ("hits", C.c_void_p), ("children", C.c_void_p)]


class OpmApodizeOp(C.Structure):
    _fields_ = [("iso", OpmIsometry), ("aperture", OpmAperture)]


class OpmThresholdOp(C.Structure):
    _fields_ = [("min_energy", C.c_double)]


class OpmLimitsOp(C.Structure):
    _fields_ = [("max_bounces", C.c_uint64), ("max_refractions", C.c_uint64), ("check_refractions", C.c_int32),
                ("_pad", C.c_int32)]


class OpmFilterOp(C.Structure):
    _fields_ = [("transmission", C.c_double), ("spectrum", C.c_void_p)]


class OpmParaxialOp(C.Structure):
    _fields_ = [("focal_length", C.c_double), ("iso", OpmIsometry)]


class OpmSplitOp(C.Structure):
    _fields_ = [("ratio", C.c_double), ("split_out", C.c_void_p), ("spectrum", C.c_void_p)]


class OpmSnapshotOp(C.Structure):
    _fields_ = [("out", C.c_void_p), ("fields", C.c_int32), ("want_spot_sums", C.c_int32), ("spot_frame", C.c_void_p)]


class OpmTransformOp(C.Structure):
    _fields_ = [("iso", OpmIsometry), ("inverse", C.c_int32), ("_pad", C.c_int32)]


class OpmDiffractOp(C.Structure):
    _fields_ = [("surface", OpmSurface), ("index", OpmIndexModel), ("grating_vector", C.c_double * 3), ("order", C.c_int32),
                ("refraction_intended", C.c_int32), ("flags", C.c_int32), ("_pad", C.c_int32), ("children", C.c_void_p)]


class OpmForkOp(C.Structure):
    _fields_ = [("ratio", C.c_double), ("spectrum", C.c_void_p)]


class OpmJoinOp(C.Structure):
    _fields_ = [("out", C.c_void_p)]


class _OpUnion(C.Union):
    _fields_ = [("refract", OpmRefractOp), ("apodize", OpmApodizeOp), ("threshold", OpmThresholdOp),
                ("limits", OpmLimitsOp), ("filter", OpmFilterOp), ("paraxial", OpmParaxialOp), ("split", OpmSplitOp),
                ("snapshot", OpmSnapshotOp), ("transform", OpmTransformOp), ("diffract", OpmDiffractOp), ("fork", OpmForkOp), ("join", OpmJoinOp)]


class OpmOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("u", _OpUnion)]


class OpmSpotStats(C.Structure):
    _fields_ = [("n_valid", C.c_uint64), ("n_total", C.c_uint64), ("total_energy", C.c_double),
                ("centroid", C.c_double * 3), ("energy_centroid", C.c_double * 3), ("beam_radius_geo", C.c_double),
                ("beam_radius_rms", C.c_double), ("energy_beam_radius_rms", C.c_double), ("bounce_lvl", C.c_uint32),
                ("_pad", C.c_uint32)]


class OpmSpotPartial(C.Structure):
    _fields_ = [("sum", C.c_double * 7), ("n_valid", C.c_uint64), ("n_total", C.c_uint64), ("first_key", C.c_uint64),
                ("max_d2_mm2", C.c_double), ("sum_d2_mm2", C.c_double), ("sum_ed2", C.c_double)]


class OpmSourceDesc(C.Structure):
    _fields_ = [("pos_dist", C.c_int32), ("energy_dist", C.c_int32), ("n_positions", C.c_uint64),
                ("grid_nx", C.c_uint64), ("grid_ny", C.c_uint64), ("size_x", C.c_double), ("size_y", C.c_double),
                ("total_energy", C.c_double), ("mu", C.c_double * 2), ("sigma", C.c_double * 2), ("power", C.c_double),
                ("theta", C.c_double), ("rectangular", C.c_int32), ("_pad", C.c_int32), ("energy_norm", C.c_double),
                ("n_wavelengths", C.c_uint32), ("_pad2", C.c_uint32), ("wavelengths", C.POINTER(C.c_double)),
                ("weights", C.POINTER(C.c_double)), ("iso", OpmIsometry)]


(NODE_SOURCE, NODE_LENS, NODE_CYLINDRIC_LENS, NODE_WEDGE, NODE_THIN_MIRROR, NODE_PARABOLIC_MIRROR, NODE_BEAM_SPLITTER,
 NODE_DUMMY, NODE_IDEAL_FILTER, NODE_PARAXIAL_SURFACE, NODE_SPOT_DIAGRAM, NODE_FLUENCE_DETECTOR, NODE_REFLECTIVE_GRATING,
 NODE_RAY_PROPAGATION_VISUALIZER, NODE_SPECTROMETER, NODE_WAVEFRONT, NODE_ENERGY_METER) = range(17)
SCENE_STRICT, SCENE_PER_SURFACE, SCENE_RECORD_ALL_HITS, SCENE_NO_SNAPSHOTS, SCENE_ASYNC, SCENE_KEEP_SOURCE = 2, 4, 8, 16, 32, 64
SCENE_POSITION_HISTORY, SCENE_KEEP_LIGHT_DATA, SCENE_LEAN_SPOT_DATA = 128, 256, 512
SNAPSHOT_ALL, SNAPSHOT_SPOT = 0, 1


class OpmCoatingDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("reflectivity", C.c_double)]


class OpmNodeDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("parent", C.c_int32), ("parent_port", C.c_int32), ("has_alignment", C.c_int32),
                ("iso", OpmIsometry), ("alignment", OpmIsometry), ("p", C.c_double * 6), ("index", OpmIndexModel),
                ("coating_in", OpmCoatingDesc), ("coating_out", OpmCoatingDesc), ("aperture_in", OpmAperture),
                ("aperture_out", OpmAperture), ("lidt", C.c_double), ("spectrum", C.c_void_p)]


class OpmWavefrontStats(C.Structure):
    _fields_ = [("wavelength_m", C.c_double), ("ptv", C.c_double), ("rms", C.c_double), ("n_rays", C.c_uint64)]


class OpmGhostFocusConfig(C.Structure):
    _fields_ = [("max_bounces", C.c_uint64), ("fluence_estimator", C.c_uint32), ("_pad", C.c_uint32)]


class OpmGhostFocusStats(C.Structure):
    _fields_ = [("interactions", C.c_uint64), ("launches", C.c_uint64), ("bundles", C.c_uint64), ("passes", C.c_uint32),
                ("_pad", C.c_uint32)]


class OpmCriticalFluence(C.Structure):
    _fields_ = [("bundle_uuid", C.c_uint64), ("peak_fluence", C.c_double), ("bounce", C.c_uint32), ("_pad", C.c_uint32)]


# every symbol include/opossum_b200_scene.h declares
SCENE_EXPORTS = [
    "opm_node_desc_init", "opm_scene_create", "opm_scene_destroy", "opm_scene_set_ambient", "opm_scene_add_node",
    "opm_scene_add_node_at_distance", "opm_scene_add_node_aligned_like", "opm_scene_calc_node_positions", "opm_scene_node_isometry",
    "opm_scene_node_count", "opm_scene_reset_data", "opm_scene_raytrace", "opm_scene_output_rays", "opm_scene_node_rays", "opm_scene_node_spot",
    "opm_scene_node_spot_partial", "opm_scene_node_spot_async",
    "opm_scene_node_fluence", "opm_scene_node_fluence_kde", "opm_scene_node_fluence_voronoi", "opm_scene_node_fluence_helper_rays", "opm_scene_beam_splitter_two_inputs", "opm_scene_node_energy", "opm_scene_node_spectrum", "opm_scene_node_wavefront", "opm_scene_node_hitmaps", "opm_scene_ghostfocus", "opm_scene_gf_pass_tally",
    "opm_scene_gf_pass_bundles", "opm_scene_node_critical_fluences", "opm_scene_node_peak_fluence",
]

# every symbol include/opossum_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "opm_abi_version", "opm_abi_sizeof", "opm_last_error", "opm_status_string", "opm_context_create", "opm_context_destroy",
    "opm_context_synchronize", "opm_context_check_errors", "opm_context_stream", "opm_context_launch_count", "opm_context_set_kernel_timing",
    "opm_context_kernel_timing", "opm_context_set_jit", "opm_context_jit_wait", "opm_context_jit_stats", "opm_jit_selftest", "opm_isometry_identity",
    "opm_isometry_new", "opm_isometry_append", "opm_isometry_from_view", "opm_parabola_off_axis_isometry", "opm_raytrace_config_default", "opm_rays_create", "opm_rays_destroy",
    "opm_rays_len", "opm_rays_capacity", "opm_rays_clear", "opm_rays_set_wavelength_hint", "opm_rays_reserve", "opm_rays_upload",
    "opm_rays_upload_soa", "opm_rays_new_collimated_with_spectrum", "opm_rays_download", "opm_rays_copy", "opm_rays_append", "opm_rays_compact",
    "opm_rays_enable_position_history", "opm_rays_disable_position_history", "opm_rays_clear_position_history",
    "opm_rays_position_history_levels", "opm_rays_position_history",
    "opm_rays_field_ptr", "opm_rays_refract_on_surface", "opm_rays_diffract_on_periodic_surface", "opm_rays_apodize",
    "opm_rays_invalidate_by_threshold_energy", "opm_rays_filter_by_nr_of_bounces",
    "opm_rays_filter_by_nr_of_refractions", "opm_rays_filter_energy", "opm_rays_split", "opm_rays_refract_paraxial",
    "opm_rays_transform", "opm_trace_sequence", "opm_rays_spot_stats", "opm_rays_spot_partial_sums",
    "opm_rays_spot_partial_radii", "opm_spot_stats_from_partial", "opm_rays_spot_sums_async", "opm_rays_spot_radii_async", "opm_rays_bounce_lvl", "opm_rays_tally", "opm_hitmap_create",
    "opm_hitmap_destroy", "opm_hitmap_capacity", "opm_hitmap_len", "opm_hitmap_upload", "opm_hitmap_download",
    "opm_hitmaps_bounding_box", "opm_fluence_binning", "opm_fluence_peak_total", "opm_hitmaps_bounding_box_async",
    "opm_fluence_binning_async", "opm_fluence_peak_total_async", "opm_fluence_kde",
    "opm_spectrum_create", "opm_spectrum_destroy", "opm_spectrum_total_energy", "opm_spectrum_center_wavelength", "opm_spectrum_get_value",
    "opm_spectrum_new", "opm_spectrum_add_single_peak", "opm_spectrum_from_laser_lines", "opm_spectrum_add_lorentzian_peak",
    "opm_spectrum_scale_vertical", "opm_spectrum_resample", "opm_spectrum_combine", "opm_spectrum_average_resolution", "opm_spectrum_merge",
    "opm_spectrum_from_csv", "opm_spectrum_generate_filter",
    "opm_rays_wavelength_range", "opm_rays_to_spectrum", "opm_rays_filter_energy_spectrum", "opm_rays_split_spectrum",
    "opm_rays_wavefront_error", "opm_readout_combine_ranges_sums", "opm_readout_combine_maps_radii", "opm_peer_group_create", "opm_peer_group_connect", "opm_peer_group_destroy",
    "opm_peer_allgather_async", "opm_peer_fluence_report_async", "opm_peer_fluence_gather_map", "opm_peer_group_phase_times", "opm_measure_fp64_peak", "opm_device_alloc",
    "opm_device_free", "opm_device_memset", "opm_device_to_host", "opm_host_to_device", "opm_host_alloc_pinned", "opm_host_free_pinned",
    "opm_spectrum_gaussian", "opm_source_generate", "opm_source_generate_strided", "opm_source_energy_norm", "opm_rays_new_collimated_xy", "opm_rays_refresh_lens", "opm_rays_expect_appends", "opm_context_interactions_total",
    "opm_context_last_first_valid_key_async", "opm_fluence_peak_check_acc_init", "opm_fluence_peak_check_async", "opm_fluence_peak_check_side_async", "opm_context_join_side_stream", "opm_fluence_peak_check_capture_key_async", "opm_fluence_peak_check_batch_async", "opm_fluence_peak_check_batch_side_async", "opm_fluence_voronoi", "opm_hitmap_bounce_levels", "opm_rays_attach_fluence_helpers", "opm_rays_has_fluence_helpers",
    "opm_rays_download_fluence_helpers", "opm_fluence_helper_rays", "opm_rays_new_collimated_w_fluence_helper", "opm_voronoi_cell_areas",
]

_lib = None


def build(force: bool = False) -> Path:
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    env = dict(os.environ)
    cmd = ["make", "-C", str(PKG / "csrc"), "-j4"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libopossum_b200.so.  Raises if it has not been built: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is the only implementation of this package)")
    L = C.CDLL(str(LIB_PATH))
    vp, dp, i32, u64, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_uint64, C.c_int64
    pp = C.POINTER(vp)
    sig = {
        "opm_abi_version": (i32, []),
        "opm_last_error": (C.c_char_p, []),
        "opm_status_string": (C.c_char_p, [i32]),
        "opm_context_create": (i32, [i32, vp, pp]),
        "opm_context_destroy": (None, [vp]),
        "opm_context_synchronize": (i32, [vp]),
        "opm_context_check_errors": (i32, [vp, i32]),
        "opm_context_stream": (vp, [vp]),
        "opm_context_launch_count": (u64, [vp]),
        "opm_context_set_kernel_timing": (i32, [vp, i32]),
        "opm_context_kernel_timing": (i32, [vp, dp, C.POINTER(u64), i32]),
        "opm_context_set_jit": (i32, [vp, i32, u64]),
        "opm_context_jit_wait": (i32, [vp]),
        "opm_context_jit_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), dp]),
        "opm_jit_selftest": (u64, [C.c_char_p, u64]),
        "opm_isometry_identity": (None, [C.POINTER(OpmIsometry)]),
        "opm_isometry_new": (i32, [dp, dp, C.POINTER(OpmIsometry)]),
        "opm_isometry_append": (None, [C.POINTER(OpmIsometry)] * 3),
        "opm_isometry_from_view": (i32, [dp, dp, dp, C.POINTER(OpmIsometry)]),
        "opm_parabola_off_axis_isometry": (i32, [C.c_double, C.c_double, dp, i32, C.POINTER(OpmIsometry), dp]),
        "opm_raytrace_config_default": (None, [C.POINTER(OpmRayTraceConfig)]),
        "opm_rays_create": (i32, [vp, u64, pp]),
        "opm_rays_destroy": (None, [vp]),
        "opm_rays_len": (u64, [vp]),
        "opm_rays_capacity": (u64, [vp]),
        "opm_abi_sizeof": (u64, [C.c_char_p]),
        "opm_rays_clear": (i32, [vp]),
        "opm_rays_set_wavelength_hint": (i32, [vp, C.c_double]),
        "opm_rays_reserve": (i32, [vp, u64]),
        "opm_rays_upload": (i32, [vp, vp, u64]),
        "opm_rays_upload_soa": (i32, [vp, C.POINTER(dp), C.POINTER(dp), dp, dp, u64]),
        "opm_rays_new_collimated_with_spectrum": (i32, [vp, vp, vp, u64, dp, dp, C.c_uint32]),
        "opm_rays_download": (i32, [vp, vp, u64, C.POINTER(u64)]),
        "opm_rays_copy": (i32, [vp, vp]),
        "opm_rays_append": (i32, [vp, vp]),
        "opm_rays_compact": (i32, [vp]),
        "opm_rays_enable_position_history": (i32, [vp, C.c_uint32]),
        "opm_rays_disable_position_history": (i32, [vp]),
        "opm_rays_clear_position_history": (i32, [vp]),
        "opm_rays_position_history_levels": (i32, [vp]),
        "opm_rays_position_history": (i32, [vp, i32, vp, vp, u64, C.c_uint32, C.POINTER(u64)]),
        "opm_rays_field_ptr": (vp, [vp, C.c_char_p]),
        "opm_rays_refract_on_surface": (i32, [vp, C.POINTER(OpmSurface), C.POINTER(OpmIndexModel), i32, i32, vp, vp,
                                              C.POINTER(OpmTraceStats)]),
        "opm_rays_diffract_on_periodic_surface": (i32, [vp, C.POINTER(OpmSurface), C.POINTER(OpmIndexModel), dp, i32, i32, vp,
                                                        C.POINTER(OpmTraceStats)]),
        "opm_rays_apodize": (i32, [vp, C.POINTER(OpmAperture), C.POINTER(OpmIsometry), C.POINTER(i32)]),
        "opm_rays_invalidate_by_threshold_energy": (i32, [vp, C.c_double]),
        "opm_rays_filter_by_nr_of_bounces": (i32, [vp, u64]),
        "opm_rays_filter_by_nr_of_refractions": (i32, [vp, u64]),
        "opm_rays_filter_energy": (i32, [vp, C.c_double]),
        "opm_rays_split": (i32, [vp, C.c_double, vp]),
        "opm_rays_refract_paraxial": (i32, [vp, C.c_double, C.POINTER(OpmIsometry)]),
        "opm_rays_transform": (i32, [vp, C.POINTER(OpmIsometry), i32]),
        "opm_trace_sequence": (i32, [vp, C.POINTER(OpmOp), i32, i32, vp, C.POINTER(OpmTraceStats)]),
        "opm_rays_spot_stats": (i32, [vp, C.POINTER(OpmIsometry), C.POINTER(OpmSpotStats)]),
        "opm_rays_spot_partial_sums": (i32, [vp, C.POINTER(OpmIsometry), C.POINTER(OpmSpotPartial)]),
        "opm_rays_spot_partial_radii": (i32, [vp, C.POINTER(OpmIsometry), C.POINTER(OpmSpotPartial)]),
        "opm_spot_stats_from_partial": (None, [C.POINTER(OpmSpotPartial), C.POINTER(OpmSpotStats)]),
        "opm_rays_spot_sums_async": (i32, [vp, C.POINTER(OpmIsometry), vp]),
        "opm_rays_spot_radii_async": (i32, [vp, C.POINTER(OpmIsometry), vp, vp]),
        "opm_hitmaps_bounding_box_async": (i32, [pp, i32, i64, vp]),
        "opm_fluence_binning_async": (i32, [pp, i32, i64, u64, u64, vp, i32, vp, vp]),
        "opm_fluence_peak_total_async": (i32, [vp, vp, u64, u64, vp, vp]),
        "opm_fluence_kde": (i32, [pp, i32, i64, u64, u64, dp, vp, dp, dp]),
        "opm_spectrum_create": (i32, [vp, dp, dp, u64, pp]),
        "opm_spectrum_destroy": (None, [vp]),
        "opm_spectrum_total_energy": (C.c_double, [dp, dp, u64]),
        "opm_spectrum_center_wavelength": (C.c_double, [dp, dp, u64]),
        "opm_spectrum_get_value": (i32, [dp, dp, u64, C.c_double, dp]),
        "opm_spectrum_new": (i32, [C.c_double, C.c_double, C.c_double, dp, dp, u64, C.POINTER(u64)]),
        "opm_spectrum_add_single_peak": (i32, [dp, dp, u64, C.c_double, C.c_double]),
        "opm_spectrum_from_laser_lines": (i32, [dp, dp, u64, C.c_double, dp, dp, u64, C.POINTER(u64)]),
        "opm_spectrum_add_lorentzian_peak": (i32, [dp, dp, u64, C.c_double, C.c_double, C.c_double]),
        "opm_spectrum_scale_vertical": (i32, [dp, u64, C.c_double]),
        "opm_spectrum_resample": (i32, [dp, dp, u64, dp, dp, u64]),
        "opm_spectrum_combine": (i32, [i32, dp, dp, u64, dp, dp, u64, dp]),
        "opm_spectrum_average_resolution": (C.c_double, [dp, u64]),
        "opm_spectrum_merge": (i32, [dp, dp, u64, dp, dp, u64, dp, dp, u64, C.POINTER(u64)]),
        "opm_spectrum_from_csv": (i32, [C.c_char_p, dp, dp, u64, C.POINTER(u64)]),
        "opm_spectrum_generate_filter": (i32, [i32] + [C.c_double] * 5 + [dp, dp, u64, C.POINTER(u64)]),
        "opm_rays_wavelength_range": (i32, [vp, dp, C.POINTER(u64)]),
        "opm_rays_to_spectrum": (i32, [vp, C.c_double, dp, dp, dp, u64, C.POINTER(u64)]),
        "opm_rays_filter_energy_spectrum": (i32, [vp, vp]),
        "opm_rays_split_spectrum": (i32, [vp, vp, vp]),
        "opm_readout_combine_ranges_sums": (i32, [vp, vp, i32, i32, i32, vp]),
        "opm_readout_combine_maps_radii": (i32, [vp, vp, i32, u64, i32, vp]),
        "opm_peer_group_create": (i32, [vp, i32, i32, u64, pp, vp]),
        "opm_peer_group_connect": (i32, [vp, vp]),
        "opm_peer_group_destroy": (None, [vp]),
        "opm_peer_allgather_async": (i32, [vp, vp, i32, vp]),
        "opm_peer_fluence_report_async": (i32, [vp, vp, i32, C.c_int64, u64, u64, vp, vp, vp, vp]),
        "opm_peer_fluence_gather_map": (i32, [vp, u64, u64, vp]),
        "opm_peer_group_phase_times": (i32, [vp, i32, vp]),
        "opm_rays_wavefront_error": (i32, [vp, C.POINTER(OpmIsometry), C.c_double, dp, u64, C.POINTER(OpmWavefrontStats)]),
        "opm_rays_bounce_lvl": (i32, [vp, C.POINTER(C.c_uint32)]),
        "opm_rays_tally": (i32, [vp, C.c_uint32, C.POINTER(u64), dp]),
        "opm_hitmap_create": (i32, [vp, u64, pp]),
        "opm_hitmap_destroy": (None, [vp]),
        "opm_hitmap_capacity": (u64, [vp]),
        "opm_hitmap_len": (u64, [vp]),
        "opm_hitmap_upload": (i32, [vp, dp, dp, dp, dp, C.POINTER(C.c_uint32), u64]),
        "opm_hitmap_download": (i32, [vp, dp, dp, dp, dp, C.POINTER(C.c_uint32), u64, C.POINTER(u64)]),
        "opm_hitmaps_bounding_box": (i32, [pp, i32, i64, dp, C.POINTER(u64)]),
        "opm_fluence_binning": (i32, [pp, i32, i64, u64, u64, dp, dp, i32, vp, dp]),
        "opm_fluence_peak_total": (i32, [vp, vp, u64, u64, dp, dp, dp]),
        "opm_measure_fp64_peak": (i32, [vp, dp]),
        "opm_device_alloc": (i32, [vp, u64, pp]),
        "opm_device_free": (None, [vp, vp]),
        "opm_device_memset": (i32, [vp, vp, i32, u64]),
        "opm_device_to_host": (i32, [vp, vp, vp, u64]),
        "opm_host_to_device": (i32, [vp, vp, vp, u64]),
        "opm_host_alloc_pinned": (i32, [u64, pp]),
        "opm_host_free_pinned": (None, [vp]),
        "opm_source_generate": (i32, [vp, C.POINTER(OpmSourceDesc), u64, u64]),
        "opm_source_generate_strided": (i32, [vp, C.POINTER(OpmSourceDesc), u64, u64, u64]),
        "opm_spectrum_gaussian": (i32, [C.c_double, C.c_double, u64, C.c_double, C.c_double, C.c_double, dp, dp]),
        "opm_source_energy_norm": (i32, [vp, C.POINTER(OpmSourceDesc), u64, u64, dp]),
        "opm_rays_new_collimated_xy": (i32, [vp, vp, u64, C.POINTER(OpmSourceDesc)]),
        "opm_fluence_voronoi": (i32, [vp, i32, C.c_int64, u64, dp, i32, vp, dp]),
        "opm_hitmap_bounce_levels": (i32, [vp, C.POINTER(u64)]),
        "opm_rays_attach_fluence_helpers": (i32, [vp, dp]),
        "opm_rays_has_fluence_helpers": (i32, [vp]),
        "opm_rays_download_fluence_helpers": (i32, [vp, dp, u64, C.POINTER(u64)]),
        "opm_fluence_helper_rays": (i32, [vp, i32, C.c_int64, u64, dp, i32, vp, dp]),
        "opm_rays_new_collimated_w_fluence_helper": (i32, [vp, dp, u64, C.c_double, dp]),
        "opm_voronoi_cell_areas": (i32, [vp, dp, dp, u64, dp]),
        "opm_rays_expect_appends": (i32, [vp, u64]),
        "opm_fluence_peak_check_side_async": (i32, [vp, i32, C.c_int64, u64, u64, vp, vp, vp]),
        "opm_context_join_side_stream": (i32, [vp]),
        "opm_fluence_peak_check_capture_key_async": (i32, [vp, vp, i32]),
        "opm_fluence_peak_check_batch_async": (i32, [vp, vp, i32, vp, u64, u64, vp, vp, vp]),
        "opm_fluence_peak_check_batch_side_async": (i32, [vp, vp, i32, vp, u64, u64, vp, vp, vp]),
    }
    sig.update({
        "opm_node_desc_init": (None, [C.POINTER(OpmNodeDesc), i32]),
        "opm_scene_create": (i32, [vp, pp]),
        "opm_scene_destroy": (None, [vp]),
        "opm_scene_set_ambient": (i32, [vp, C.POINTER(OpmIndexModel)]),
        "opm_scene_add_node": (i32, [vp, C.POINTER(OpmNodeDesc), C.POINTER(i32)]),
        "opm_scene_add_node_at_distance": (i32, [vp, C.POINTER(OpmNodeDesc), C.c_double, C.POINTER(i32)]),
        "opm_scene_node_energy": (i32, [vp, i32, dp]),
        "opm_scene_add_node_aligned_like": (i32, [vp, C.POINTER(OpmNodeDesc), i32, C.c_double, C.c_double, C.POINTER(i32)]),
        "opm_scene_calc_node_positions": (i32, [vp, C.c_double, C.POINTER(OpmRayTraceConfig)]),
        "opm_scene_node_isometry": (i32, [vp, i32, C.POINTER(OpmIsometry)]),
        "opm_scene_node_count": (i32, [vp]),
        "opm_scene_reset_data": (i32, [vp]),
        "opm_scene_raytrace": (i32, [vp, vp, C.POINTER(OpmRayTraceConfig), i32, C.POINTER(OpmTraceStats)]),
        "opm_scene_output_rays": (vp, [vp]),
        "opm_scene_node_rays": (vp, [vp, i32]),
        "opm_scene_node_spot": (i32, [vp, i32, C.POINTER(OpmSpotStats)]),
        "opm_scene_node_spot_partial": (i32, [vp, i32, i32, C.POINTER(OpmSpotPartial)]),
        "opm_scene_node_spot_async": (i32, [vp, i32, i32, vp, vp]),
        "opm_scene_node_fluence": (i32, [vp, i32, u64, u64, dp, dp, dp, dp]),
        "opm_scene_node_fluence_kde": (i32, [vp, i32, u64, u64, dp, dp, dp, dp, dp]),
        "opm_scene_node_fluence_voronoi": (i32, [vp, i32, u64, u64, dp, dp, dp, dp]),
        "opm_scene_node_fluence_helper_rays": (i32, [vp, i32, u64, u64, dp, dp, dp, dp]),
        "opm_scene_beam_splitter_two_inputs": (i32, [vp, i32, vp, vp, vp, vp, vp]),
        "opm_scene_node_spectrum": (i32, [vp, i32, C.c_double, dp, dp, dp, u64, C.POINTER(u64)]),
        "opm_scene_node_wavefront": (i32, [vp, i32, dp, u64, C.POINTER(OpmWavefrontStats)]),
        "opm_scene_node_hitmaps": (i32, [vp, i32, i32, pp, i32]),
        "opm_scene_ghostfocus": (i32, [vp, vp, C.POINTER(OpmGhostFocusConfig), i32, C.POINTER(OpmGhostFocusStats)]),
        "opm_scene_gf_pass_tally": (i32, [vp, C.c_uint32, C.c_uint32, C.POINTER(u64), dp, C.POINTER(u64), C.POINTER(u64)]),
        "opm_scene_gf_pass_bundles": (i32, [vp, C.c_uint32, pp, i32]),
        "opm_scene_node_critical_fluences": (i32, [vp, i32, i32, C.POINTER(OpmCriticalFluence), i32]),
        "opm_scene_node_peak_fluence": (i32, [vp, i32, i32, dp]),
    })
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
