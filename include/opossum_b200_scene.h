/*
 * opossum_b200_scene.h -- host-side mirror of the reference's node / analyzer layer for the
 * ray-propagation path, written in C++ above the C ABI of opossum_b200.h and exported with C linkage.
 *
 * What it replaces in the reference (opossum/src/):
 *   - the node -> surface compilers `OpticNode::update_surfaces` of Lens (nodes/lens/mod.rs:226-293),
 *     CylindricLens (nodes/cylindric_lens/mod.rs:146-214), Wedge (nodes/wedge/mod.rs:116-160), ThinMirror
 *     (nodes/thin_mirror.rs:121-155), ParabolicMirror on and off axis (nodes/parabolic_mirror.rs:286-320,388-421), and the flat
 *     single-surface nodes (optic_node.rs:62-80): BeamSplitter, Dummy, IdealFilter, ParaxialSurface, Source,
 *     SpotDiagram, FluenceDetector;
 *   - each node's `AnalysisRayTrace::analyze` / `AnalysisGhostFocus::analyze` and the shared
 *     `pass_through_surface` / `pass_through_detector_surface` (analyzers/raytrace.rs:96-207);
 *   - the drivers `RayTracingAnalyzer::analyze` (analyzers/raytrace.rs:36-49 with the walk of
 *     nodes/node_group/analysis_raytrace.rs:29-91) and `GhostFocusAnalyzer::analyze`
 *     (analyzers/ghostfocus.rs:82-122 with nodes/node_group/analysis_ghostfocus.rs:22-109).
 *
 *   - node positioning from the optical axis: `NodeGroup::connect_nodes(.., distance)` + `calc_node_positions`
 *     (nodes/node_group/analysis_raytrace.rs:92-212, nodes/node_group/optic_graph.rs:878-926).
 *
 * Scope limits (SURVEY.md 8f "next"): the scene is a chain of nodes (explicit isometries, or edge distances placed by
 * opm_scene_calc_node_positions); a beam splitter may feed a side chain from its second output; `.opm` scene files are read
 * by the host-side glue above this interface (opossum_b200/opm_document.py).  Fluence estimators on the device: Binning and KDE.
 */
#ifndef OPOSSUM_B200_SCENE_H
#define OPOSSUM_B200_SCENE_H

#include "opossum_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    OPM_NODE_SOURCE = 0,
    OPM_NODE_LENS = 1,
    OPM_NODE_CYLINDRIC_LENS = 2,
    OPM_NODE_WEDGE = 3,
    OPM_NODE_THIN_MIRROR = 4,
    OPM_NODE_PARABOLIC_MIRROR = 5,
    OPM_NODE_BEAM_SPLITTER = 6,
    OPM_NODE_DUMMY = 7,
    OPM_NODE_IDEAL_FILTER = 8,
    OPM_NODE_PARAXIAL_SURFACE = 9,
    OPM_NODE_SPOT_DIAGRAM = 10,
    OPM_NODE_FLUENCE_DETECTOR = 11,
    OPM_NODE_REFLECTIVE_GRATING = 12, /* nodes/reflective_grating.rs */
    OPM_NODE_RAY_PROPAGATION_VISUALIZER = 13, /* nodes/ray_propagation_visualizer.rs: a single-surface detector that keeps the bundle;
                                                 with OPM_SCENE_POSITION_HISTORY its rays carry their paths */
    OPM_NODE_SPECTROMETER = 14,  /* nodes/spectrometer.rs: detector, report = Rays::to_spectrum(0.2 nm) */
    OPM_NODE_WAVEFRONT = 15,     /* nodes/wavefront.rs: detector, report = wavefront error map, PtV, RMS */
    OPM_NODE_ENERGY_METER = 16   /* nodes/energy_meter.rs: detector, report = Rays::total_energy of the bundle it keeps */
} OpmNodeType;

typedef struct { int32_t type; int32_t _pad; double reflectivity; } OpmCoatingDesc; /* coatings/mod.rs:17-29 */

/* One optic node.  Geometry parameters p[]:
 *   LENS / CYLINDRIC_LENS : p[0] front curvature, p[1] rear curvature (+-INFINITY = flat), p[2] centre thickness
 *   WEDGE                 : p[0] centre thickness, p[1] wedge angle [rad]
 *   THIN_MIRROR           : p[0] curvature (INFINITY = flat)
 *   PARABOLIC_MIRROR      : p[0] focal length, p[1] off-axis angle [rad] (0 = on axis), p[2], p[3] off-axis direction (x, y) in the
 *                           local frame (0, 0 = the default (1, 0)), p[4] != 0: collimating (nodes/parabolic_mirror.rs:45-70)
 *   BEAM_SPLITTER         : p[0] splitting ratio (SplittingConfig::Ratio, ray.rs:36-46), or `spectrum` (SplittingConfig::Spectrum)
 *   IDEAL_FILTER          : p[0] transmission (FilterType::Constant), or `spectrum` (FilterType::Spectrum)
 *   PARAXIAL_SURFACE      : p[0] focal length
 *   REFLECTIVE_GRATING    : p[0] line density [1/m] (default 1740/mm), p[1] diffraction order (default -1)
 */
typedef struct {
    int32_t type;              /* OpmNodeType */
    int32_t parent;            /* index of the node whose output feeds this one; -1 = the previous node */
    int32_t parent_port;       /* 0 = main output, 1 = second output of a beam splitter */
    int32_t has_alignment;     /* local decenter/tilt appended to the node isometry (optic_node.rs:356-363) */
    OpmIsometry iso;           /* node isometry */
    OpmIsometry alignment;
    double p[6];
    OpmIndexModel index;       /* glass of LENS / CYLINDRIC_LENS / WEDGE */
    OpmCoatingDesc coating_in, coating_out; /* per optic surface (input_1 / output_1) */
    OpmAperture aperture_in, aperture_out;
    double lidt;               /* J/m^2; 0 = the default 1 J/cm^2 (surface/optic_surface.rs:55) */
    const OpmSpectrum* spectrum; /* IDEAL_FILTER / BEAM_SPLITTER: transmission spectrum (must outlive the scene); NULL = use p[0] */
} OpmNodeDesc;

/* defaults of the reference's node constructors: IdealAR coatings (mirrors: ConstantR 1.0), no apertures */
void opm_node_desc_init(OpmNodeDesc* d, int32_t type);

typedef struct OpmScene OpmScene;

OpmStatus opm_scene_create(OpmContext* ctx, OpmScene** out);
void opm_scene_destroy(OpmScene* scene);
/* SceneryResources::ambient_refr_index (optic_scenery_rsc.rs:8-20); default vacuum */
OpmStatus opm_scene_set_ambient(OpmScene* scene, const OpmIndexModel* ambient);
OpmStatus opm_scene_add_node(OpmScene* scene, const OpmNodeDesc* desc, int32_t* node_index);
int32_t opm_scene_node_count(const OpmScene* scene);

/* ---- node positioning from the optical axis
 * NodeGroup::add_node + connect_nodes(predecessor, port, node, "input_1", distance) for a node WITHOUT an isometry of its own
 * (desc->iso is ignored; desc->parent / parent_port name the predecessor's output as in opm_scene_add_node; the local
 * `alignment` still applies, optic_node.rs:356-363).  The node cannot be analysed before it has been placed. */
OpmStatus opm_scene_add_node_at_distance(OpmScene* scene, const OpmNodeDesc* desc, double distance_m, int32_t* node_index);
/* OpticNode::align_like_node_at_distance (optic_node.rs:476-480, nodes/node_group/analysis_raytrace.rs:142-163): the node is not placed
 * on the optical axis but gets the isometry of node `like_node` moved by `distance_m` along that node's local z; it is still fed by
 * desc->parent.  If `like_node` has no isometry when its turn comes, the node is placed on the axis at `axis_distance_m` instead. */
OpmStatus opm_scene_add_node_aligned_like(OpmScene* scene, const OpmNodeDesc* desc, int32_t like_node, double distance_m,
                                          double axis_distance_m, int32_t* node_index);
/* AnalysisRayTrace::calc_node_positions of the node group (nodes/node_group/analysis_raytrace.rs:92-212), which
 * RayTracingAnalyzer::analyze runs before the analysis proper (analyzers/raytrace.rs:44): ONE axis ray -- Rays::
 * get_optical_axis_ray = Ray::new_collimated(origin, axis_wavelength, 1 J) in the source's frame (nodes/source.rs:236-268,
 * rays.rs:1414-1421) -- is passed from node to node through each node's own analysis on the device (beam splitter: only its
 * input surface, both outputs carry the ray; beam_splitter/analysis_raytrace.rs:40-93); every node that has not been placed
 * gets Isometry::new_from_view(position of the ray after Ray::propagate(distance), its direction, up) (optic_graph.rs:
 * 878-926, ray.rs:408-434); `up` starts as Ray::define_up_direction of the source's axis ray (ray.rs:818-838) and is rotated by
 * each node's last direction change (Ray::calc_new_up_direction, ray.rs:842-860).  Nodes are walked in index order (the
 * reference: topological order).  Ends with reset_data like the reference.  config NULL = RayTraceConfig::default(). */
OpmStatus opm_scene_calc_node_positions(OpmScene* scene, double axis_wavelength_m, const OpmRayTraceConfig* config);
/* OpticNode::isometry(): the node isometry (without the local alignment); OPM_ERR_ARG while the node is not placed */
OpmStatus opm_scene_node_isometry(const OpmScene* scene, int32_t node, OpmIsometry* out);
/* OpticNode::reset_data: clears hit maps, ray caches and stored detector data (optic_node.rs:84-95) */
OpmStatus opm_scene_reset_data(OpmScene* scene);

enum {
    OPM_SCENE_STRICT = 2,           /* same value as OPM_TRACE_STRICT */
    OPM_SCENE_PER_SURFACE = 4,      /* one launch per reference call instead of one fused launch per chain */
    OPM_SCENE_RECORD_ALL_HITS = 8,  /* keep the hit map of every surface (the reference always does); default: FluenceDetector nodes only */
    OPM_SCENE_NO_SNAPSHOTS = 16,    /* detectors keep hit maps but no copy of the bundle (no spot statistics) */
    OPM_SCENE_ASYNC = 32,           /* do not synchronise / read counters back (throughput runs) */
    OPM_SCENE_KEEP_SOURCE = 64,     /* leave `source_rays` untouched: the main chain's output goes to opm_scene_output_rays() */
    OPM_SCENE_KEEP_LIGHT_DATA = 256, /* FluenceDetector nodes also keep a copy of the bundle (the reference's `light_data`); no report
                                       reads it (nodes/fluence_detector/mod.rs:91-135), so by default only their hit map is kept */
    OPM_SCENE_LEAN_SPOT_DATA = 512, /* SpotDiagram nodes keep only what their report reads of their copy of the bundle: position, energy,
                                       valid flag, id and bounce count per ray (OPM_SNAPSHOT_SPOT; nodes/spot_diagram.rs:101-163); the
                                       other arrays of opm_scene_node_rays() are undefined */
    OPM_SCENE_POSITION_HISTORY = 128 /* keep the per-ray position history (`pos_hist`): read it with opm_rays_position_history() from the
                                       bundles of opm_scene_node_rays() / opm_scene_output_rays() / the traced source bundle */
};

/* RayTracingAnalyzer::analyze.  `source_rays` is the bundle the Source node's light-data builder produced, in
 * the source's local frame; it is traced in place and holds the main chain's output afterwards. */
OpmStatus opm_scene_raytrace(OpmScene* scene, OpmRays* source_rays, const OpmRayTraceConfig* config, int32_t flags,
                             OpmTraceStats* stats);

/* output bundle of the main chain when OPM_SCENE_KEEP_SOURCE was given (owned by the scene), else NULL */
OpmRays* opm_scene_output_rays(OpmScene* scene);

/* detector read-out (node_report of SpotDiagram / FluenceDetector) */
OpmRays* opm_scene_node_rays(OpmScene* scene, int32_t node);           /* bundle stored at a detector / split-off bundle of a splitter */
OpmStatus opm_scene_node_spot(OpmScene* scene, int32_t node, OpmSpotStats* out); /* nodes/spot_diagram.rs:101-163 */
/* the same report in two phases for a bundle sharded over GPUs (see OpmSpotPartial): phase 1 = sums, phase 2 = radii */
OpmStatus opm_scene_node_spot_partial(OpmScene* scene, int32_t node, int32_t phase, OpmSpotPartial* inout);
/* device-chained form: phase 1 writes sums_dev[11], phase 2 reads sums_dev and writes radii_dev[3] (opm_rays_spot_*_async) */
OpmStatus opm_scene_node_spot_async(OpmScene* scene, int32_t node, int32_t phase, double* sums_dev, double* radii_dev);
/* nodes/fluence_detector/mod.rs:91-135 with FluenceEstimator::Binning: map_host (nx*ny, column-major ny x nx) may be NULL */
OpmStatus opm_scene_node_fluence(OpmScene* scene, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                 double* peak, double* total_energy);
/* the same report with FluenceEstimator::KDE (rays_hit_map.rs:575-613, kde/mod.rs); band_width = the estimated kernel width */
OpmStatus opm_scene_node_fluence_kde(OpmScene* scene, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                     double* peak, double* total_energy, double* band_width);
/* BeamSplitter::analyze_raytrace with BOTH inputs (nodes/beam_splitter/mod.rs:158-293, analysis_raytrace.rs:12-39; the reference's
 * test :149-181: 1 J and 0.5 J in at ratio 0.6 -> 0.8 J and 0.7 J out): out1 = what in1 keeps + what in2 splits off, out2 the other
 * way round, after the input surface / aperture and before the output aperture and the energy threshold.  The scene GRAPH is a
 * tree (one input per node); this is the node-level call for a caller that holds two bundles.  in1 or in2 may be NULL; the
 * inputs are not modified; out1 / out2 are overwritten.  config NULL = RayTraceConfig::default(). */
OpmStatus opm_scene_beam_splitter_two_inputs(OpmScene* scene, int32_t node, OpmRays* in1, OpmRays* in2, const OpmRayTraceConfig* config,
                                             OpmRays* out1, OpmRays* out2);
/* the same report with FluenceEstimator::Voronoi, the node's default (nodes/fluence_detector/mod.rs:61): HitMap::
 * calc_combined_fluence_with_voronoi (hit_map/mod.rs:232-262) -- one estimate per (bundle, bounce level) on the common bounding box,
 * summed; NaN outside a bundle's convex hull.  nx x nx map; nx != ny is refused (OPM_ERR_ARG) where the reference panics on the
 * dimensions of its matrix sum -- which is what its own (100, 83) report does. */
/* the same report with FluenceEstimator::HelperRays (HitMap::calc_combined_fluence_with_helper_rays, hit_map/mod.rs:283-310), for a
 * scene that traced a bundle with fluence helper rays (opm_rays_attach_fluence_helpers): the detector's hit maps then hold
 * FluenceHitPoints.  Such a bundle runs the whole chain in the interpreter kernel; beam splitters with a second output are refused. */
OpmStatus opm_scene_node_fluence_helper_rays(OpmScene* scene, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                             double* peak, double* total);
OpmStatus opm_scene_node_fluence_voronoi(OpmScene* scene, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                         double* peak, double* total_energy);
/* EnergyMeter::node_report (nodes/energy_meter.rs:136-151): total energy of the valid rays of the bundle the node keeps */
OpmStatus opm_scene_node_energy(OpmScene* scene, int32_t node, double* energy_J);
/* Spectrometer::node_report (nodes/spectrometer.rs:155-175): Rays::to_spectrum of the bundle the node keeps; resolution_m <= 0:
 * the node's 0.2 nm.  lambda_um == NULL: only *n_out.  Arguments as opm_rays_to_spectrum. */
OpmStatus opm_scene_node_spectrum(OpmScene* scene, int32_t node, double resolution_m, const double* range_m, double* lambda_um,
                                  double* values, uint64_t cap, uint64_t* n_out);
/* WaveFront::node_report (nodes/wavefront.rs:173-235): wavefront error map of the kept bundle in the frame of the node's
 * surface, at the centre wavelength of its spectrum, with PtV and RMS.  Arguments as opm_rays_wavefront_error. */
OpmStatus opm_scene_node_wavefront(OpmScene* scene, int32_t node, double* xyw_out, uint64_t cap, OpmWavefrontStats* out);
/* hit maps recorded on surface `surface` (0 = input_1, 1 = output_1) of a node: fills up to `capacity` handles */
int32_t opm_scene_node_hitmaps(OpmScene* scene, int32_t node, int32_t surface, OpmHitMap** out, int32_t capacity);

/* ---- ghost focus (analyzers/ghostfocus.rs) */
typedef enum { OPM_ESTIMATOR_BINNING = 0, OPM_ESTIMATOR_VORONOI = 1 } OpmFluenceEstimator;
typedef struct {
    uint64_t max_bounces;
    uint32_t fluence_estimator;  /* OpmFluenceEstimator of the per-bundle LIDT check (GhostFocusConfig::fluence_estimator; the
                                    reference's default is Voronoi, analyzers/ghostfocus.rs:61).  Binning chains on the device; a
                                    Voronoi check synchronises (its cells are prepared on the host). */
    uint32_t _pad;
} OpmGhostFocusConfig;
typedef struct {
    uint64_t interactions;
    uint64_t launches;
    uint64_t bundles;          /* ray bundles that reached a sink over all passes */
    uint32_t passes;
    uint32_t _pad;
} OpmGhostFocusStats;
OpmStatus opm_scene_ghostfocus(OpmScene* scene, OpmRays* source_rays, const OpmGhostFocusConfig* config, int32_t flags,
                               OpmGhostFocusStats* stats);
/* accumulated rays per pass (scenery.add_to_accumulated_rays, analyzers/ghostfocus.rs:117-119): per-bounce-order tallies
 * of the bundles collected in pass `pass` */
OpmStatus opm_scene_gf_pass_tally(OpmScene* scene, uint32_t pass, uint32_t n_orders, uint64_t* counts, double* energy,
                                  uint64_t* n_bundles, uint64_t* n_rays);
/* bundles collected in a pass (ray_collection): returns the number written */
int32_t opm_scene_gf_pass_bundles(OpmScene* scene, uint32_t pass, OpmRays** out, int32_t capacity);
/* critical fluences of a surface (hit_map/mod.rs:121-146): peak fluence [J/m^2] per offending bundle */
typedef struct { uint64_t bundle_uuid; double peak_fluence; uint32_t bounce; uint32_t _pad; } OpmCriticalFluence;
int32_t opm_scene_node_critical_fluences(OpmScene* scene, int32_t node, int32_t surface, OpmCriticalFluence* out,
                                         int32_t capacity);
/* peak of the per-bundle 101x101 Binning maps evaluated on a surface during the analysis (max over bundles) */
OpmStatus opm_scene_node_peak_fluence(OpmScene* scene, int32_t node, int32_t surface, double* peak);

#ifdef __cplusplus
}
#endif
#endif
