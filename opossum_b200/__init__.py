"""opossum_b200 -- B200-native ray-propagation engine behind OPOSSUM's `Rays` API.

The package is a thin host layer over libopossum_b200.so (hand-written sm_100a CUDA kernels behind the C ABI
of include/opossum_b200.h).  There is no CPU implementation: importing works anywhere, but creating a
`Context` needs a B200 and raises otherwise.
"""
from . import _ffi as ffi  # noqa: F401
from .api import (Aperture, Context, FluenceData, HitMap, Isometry, OpossumError, OpticSurface, Program, RAY_DTYPE,  # noqa: F401
                  RayTraceConfig, Rays, RefractiveIndex, SourceDesc, TraceStats, fluence_binning, fluence_kde, fluence_peak_total,
                  fluence_voronoi, fluence_voronoi_combined, fluence_helper_rays, hitmap_bounce_levels, hitmaps_bounding_box, refr_index_vaccuum, spectrum_gaussian, Spectrum, spectrum_total_energy,
                  spectrum_center_wavelength, spectrum_get_value, spectrum_new, spectrum_add_single_peak, spectrum_from_laser_lines, spectrum_add_lorentzian_peak, spectrum_scale_vertical, spectrum_resample, spectrum_filter, spectrum_add, spectrum_sub, spectrum_split_by_spectrum, spectrum_average_resolution, spectrum_merge, spectrum_from_csv, spectrum_generate_filter)

__all__ = ["Aperture", "Context", "FluenceData", "HitMap", "Isometry", "OpossumError", "OpticSurface", "Program",
           "RAY_DTYPE", "RayTraceConfig", "Rays", "RefractiveIndex", "SourceDesc", "TraceStats", "fluence_binning",
           "fluence_kde", "fluence_peak_total", "fluence_voronoi", "fluence_voronoi_combined", "fluence_helper_rays", "hitmap_bounce_levels", "hitmaps_bounding_box", "refr_index_vaccuum", "spectrum_gaussian", "Spectrum",
           "spectrum_total_energy", "spectrum_center_wavelength", "spectrum_get_value", "spectrum_new", "spectrum_add_single_peak", "spectrum_from_laser_lines", "spectrum_add_lorentzian_peak", "spectrum_scale_vertical", "spectrum_resample", "spectrum_filter", "spectrum_add", "spectrum_sub", "spectrum_split_by_spectrum", "spectrum_average_resolution", "spectrum_merge", "spectrum_from_csv", "spectrum_generate_filter", "ffi"]
from .opm_document import OpmDocument, OpmDocumentError, load_opm, loads_opm  # noqa: E402,F401  (.opm scene files)
