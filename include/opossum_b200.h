/*
 * opossum_b200.h -- C ABI of the B200-native ray-propagation engine.
 *
 * This is the drop-in boundary for OPOSSUM's ray-propagation hot path.  The reference
 * (opossum-labs/opossum v0.6.0, pure Rust) has no FFI: its extension point for this path is the
 * set of `Rays` bundle methods that every node's `analyze` calls through
 * `AnalysisRayTrace::pass_through_surface` (analyzers/raytrace.rs:96-140).  Each entry point
 * below names the reference interface it replaces (file:line under opossum/src/).  A Rust host
 * binds these with a plain `extern "C"` block; plain pointers and sizes
 * only, no C++/torch types.
 *
 * Conventions
 *  - SI base units like the reference's `uom` quantities: metres, joules, radians.
 *  - Every function returns an OpmStatus (0 = ok).  `opm_last_error()` gives the message of the
 *    last failure on the calling thread; messages mirror the reference's `OpossumError` strings.
 *  - Ownership: the caller owns every handle it creates and destroys it; the library never frees
 *    caller memory.  Handles belong to the context (device + stream) they were created on and
 *    are not thread-safe (the reference's analyzers are single-threaded per scenery).
 *  - There is NO CPU fallback.  Every compute entry point fails with OPM_ERR_CUDA when no
 *    sm_100 device / driver is present.
 */
#ifndef OPOSSUM_B200_H
#define OPOSSUM_B200_H

#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#else /* compiled at run time by NVRTC (the library's own kernels): no system headers there */
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OPM_ABI_VERSION 1
#define OPM_MAX_APERTURE_ITEMS 4
#define OPM_MAX_PROGRAM_OPS 128

typedef int32_t OpmStatus;
enum {
    OPM_OK = 0,
    OPM_ERR_N2_INVALID = 1,      /* "the refractive index must be >=1.0 and finite"            ray.rs:588-592 */
    OPM_ERR_WVL_RANGE = 2,       /* "wavelength outside valid range"                           refr_index_sellmeier1.rs:79-81 */
    OPM_ERR_INDEX_MODEL = 3,     /* "refractive index calculated by model is <1.0 or not finite" refractive_index/mod.rs:54-58 */
    OPM_ERR_COATING = 4,         /* "reflectivity must be within [0.0, 1.0] and finite."       coatings/constant_r.rs:23-27 */
    OPM_ERR_APODIZE_FACTOR = 5,  /* "transmission factor must be within (0.0..=1.0)"           ray.rs:705-709 */
    OPM_ERR_THRESHOLD = 6,       /* "threshold energy must be finite"                          rays.rs:1150-1154 */
    OPM_ERR_HITPOINT = 7,        /* hit point energy/position not finite                       rays_hit_map.rs:59-66 */
    OPM_ERR_SPLIT_RATIO = 8,     /* "splitting_ratio must be within [0.0;1.0]"                 ray.rs:763-767 */
    OPM_ERR_FOCAL_LENGTH = 9,    /* "focal length must be !=0.0 and finite"                    rays.rs:944-948 */
    OPM_ERR_EMPTY_HITMAP = 10,   /* "could not calculate bounding box"                         rays_hit_map.rs:526-531 */
    OPM_ERR_BIN_RANGE = 11,      /* hit outside an explicit map range (the reference panics)   rays_hit_map.rs:364-374 */
    OPM_ERR_ARG = 12,            /* invalid argument / handle */
    OPM_ERR_CAPACITY = 13,       /* an output bundle / hit map is too small */
    OPM_ERR_CUDA = 14,           /* CUDA runtime failure (including: no device) */
    OPM_ERR_KDE_BANDWIDTH = 15,  /* "bandwidth must be != 0.0 and finite"                      kde/mod.rs:36-42 */
    OPM_ERR_SPECTRUM = 16,       /* OpossumError::Spectrum; "wavelength of ray outside filter spectrum" ray.rs:716-720, :757-761;
                                    "ray bundle is empty - cannot create spectrum" rays.rs:1216-1220 */
    OPM_ERR_EMPTY_WAVEFRONT = 17,/* "Empty wavefront-data vector!"                             nodes/wavefront.rs:121-123 */
    OPM_ERR_PEER_TIMEOUT = 18,   /* multi-GPU read-out: a peer rank did not arrive at a meeting in time (no reference counterpart) */
    OPM_ERR_VORONOI = 19         /* Voronoi estimator: fewer than 3 hits, non-finite or collinear points  rays_hit_map.rs:416-420, griddata.rs:269-283 */
};

/* ------------------------------------------------------------------------------------------
 * Plain-old-data descriptors
 * ---------------------------------------------------------------------------------------- */

/* Host-side array-of-structs mirror of the parity fields of `struct Ray` (ray.rs:66-96).  Used only by
 * upload/download; on the device rays live as struct-of-arrays f64 buffers. */
typedef struct {
    double pos[3];
    double dir[3];
    double e;         /* energy, J */
    double wvl;       /* wavelength, m */
    double opl;       /* path_length, m */
    double n;         /* refractive_index of the medium the ray is in */
    uint32_t bounces; /* number_of_bounces */
    uint32_t refr;    /* number_of_refractions */
    uint32_t valid;   /* 0 invalid, 1 valid */
    uint32_t id;      /* index of the seed ray (not in the reference; keys unordered outputs) */
} OpmRay;             /* 96 bytes */

/* Rigid transform p' = rot * p + t (rot row-major).  `Isometry` (utils/geom_transformation.rs:27-31)
 * keeps the transform and its cached inverse; so do we. */
typedef struct { double rot[9]; double t[3]; } OpmIso;
typedef struct { OpmIso fwd; OpmIso inv; } OpmIsometry;

typedef enum { OPM_GEOM_PLANE = 0, OPM_GEOM_SPHERE = 1, OPM_GEOM_CYLINDER = 2, OPM_GEOM_PARABOLA = 3 } OpmGeom;
typedef enum { OPM_COAT_IDEAL_AR = 0, OPM_COAT_CONSTANT_R = 1, OPM_COAT_FRESNEL = 2 } OpmCoating;
typedef enum { OPM_IDX_CONST = 0, OPM_IDX_SELLMEIER1 = 1, OPM_IDX_SCHOTT = 2, OPM_IDX_CONRADY = 3 } OpmIndexType;
typedef enum { OPM_MISS_STOP = 0, OPM_MISS_IGNORE = 1 } OpmMissedSurfaceStrategy; /* analyzers/raytrace.rs:282-293 */
typedef enum { OPM_AP_NONE = 0, OPM_AP_CIRCLE = 1, OPM_AP_RECT = 2, OPM_AP_POLYGON = 3, OPM_AP_GAUSSIAN = 4 } OpmApertureType;

/* `OpticSurface` geometry + coating (surface/optic_surface.rs:29-43, surface/geo_surface.rs:19-47).
 * `iso` is the isometry of the GEO surface: for spheres/cylinders the frame of the centre of
 * curvature (node_iso o T(0,0,R), nodes/lens/mod.rs:238-247). */
typedef struct {
    int32_t geom;        /* OpmGeom */
    int32_t coating;     /* OpmCoating */
    double geom_param;   /* radius of curvature (sphere, cylinder) or focal length (parabola) */
    double coating_r;    /* ConstantR reflectivity */
    OpmIsometry iso;
} OpmSurface;

/* `RefractiveIndexType` (refractive_index/mod.rs:20-32) */
typedef struct {
    int32_t type;        /* OpmIndexType */
    int32_t _pad;
    double c[6];         /* Const: n | Sellmeier1: k1,k2,k3,l1,l2,l3 | Schott: a0..a5 | Conrady: n0,a,b */
    double wvl_lo, wvl_hi; /* valid range [lo, hi) in metres (ignored for Const) */
} OpmIndexModel;

/* `Aperture` (aperture.rs:66-84).  n_items = 0 is Aperture::None; is_stack = 1 multiplies the items
 * (aperture.rs:407-416).  One level of stacking, at most OPM_MAX_APERTURE_ITEMS members. */
typedef struct {
    int32_t type;        /* OpmApertureType */
    int32_t obstruction; /* ApertureType::Obstruction */
    double p[4];         /* circle: r,cx,cy | rect: w,h,cx,cy | gaussian: sigma_x,sigma_y,cx,cy */
} OpmApertureItem;
typedef struct {
    int32_t n_items;
    int32_t is_stack;
    int32_t stack_obstruction;
    int32_t n_tris;          /* polygon only: number of triangles in `tris` */
    OpmApertureItem items[OPM_MAX_APERTURE_ITEMS];
    const double* tris;      /* HOST pointer, n_tris*6 doubles (x1,y1,x2,y2,x3,y3), ear-cut by the caller */
} OpmAperture;

/* `RayTraceConfig` (analyzers/raytrace.rs:302-323) */
typedef struct {
    double min_energy_per_ray;       /* default 1e-12 J */
    uint64_t max_number_of_bounces;  /* default 1000 */
    uint64_t max_number_of_refractions; /* default 1000 */
    int32_t missed_surface_strategy; /* default OPM_MISS_STOP */
    int32_t _pad;
} OpmRayTraceConfig;

/* Counters of one launch.  `rays_missed` / `valid_rays_found` drive the reference's warnings
 * (rays.rs:1026-1031); `interactions` is the benchmark unit: one valid ray through one
 * `refract_on_surface`. */
typedef struct {
    uint64_t interactions;
    uint64_t children;       /* reflected rays emitted */
    uint64_t hits;           /* hit-map entries recorded */
    uint64_t missed;         /* valid rays that missed or were totally reflected */
    uint64_t apodized;       /* rays invalidated by an aperture (rays.rs:557-573 return value != 0) */
    uint64_t valid_out;      /* valid rays in the bundle after the call */
    uint64_t alive_out;      /* rays in the bundle after the call (valid or not) */
    uint32_t status_bits;    /* bit k set = OpmStatus k raised by some ray */
    uint32_t _pad;
    uint64_t first_valid_key; /* (id << 32 | bounces) of the valid ray with the lowest id after the call, ~0 if none:
                                 Rays::bounce_lvl (rays.rs:183-194) without a second pass over the bundle */
} OpmTraceStats;

typedef struct OpmContext OpmContext;
typedef struct OpmRays OpmRays;        /* device ray bundle, struct-of-arrays */
typedef struct OpmHitMap OpmHitMap;    /* device hit-point record of one surface pass (rays_hit_map.rs:43-49) */

/* ------------------------------------------------------------------------------------------
 * Fused surface-sequence program: one kernel launch runs all ops over every ray, ray state in
 * registers, op descriptors staged in shared memory.  One op = one `Rays` method call.
 * ---------------------------------------------------------------------------------------- */
typedef enum {
    OPM_OP_REFRACT = 1,    /* Rays::refract_on_surface            rays.rs:976-1047  */
    OPM_OP_APODIZE = 2,    /* Rays::apodize                       rays.rs:557-573   */
    OPM_OP_THRESHOLD = 3,  /* Rays::invalidate_by_threshold_energy rays.rs:1145-1162 */
    OPM_OP_LIMITS = 4,     /* filter_by_nr_of_bounces / _refractions rays.rs:1386-1404 */
    OPM_OP_FILTER = 5,     /* Rays::filter_energy(Constant)       rays.rs:1119-1133 */
    OPM_OP_PARAXIAL = 6,   /* Rays::refract_paraxial              rays.rs:943-959   */
    OPM_OP_SPLIT = 7,      /* Rays::split(Ratio)                  rays.rs:1253-1264 */
    OPM_OP_SNAPSHOT = 8,   /* copy of the bundle a detector node keeps (analyzers/raytrace.rs:186-205) */
    OPM_OP_TRANSFORM = 9,  /* Rays::transformed_by_iso (nodes/source.rs:205-210)                      */
    OPM_OP_DIFFRACT = 10,  /* Rays::diffract_on_periodic_surface  rays.rs:1058-1111 (reflective grating) */
    OPM_OP_FORK = 11,      /* Rays::split(Ratio) whose split-off bundle is traced in the SAME launch: the ops up to the matching
                              OPM_OP_JOIN run on the split-off rays (E*(1-ratio)) while the main bundle (E*ratio) is parked in
                              shared memory; the beam splitter's second output never takes a round trip through HBM */
    OPM_OP_JOIN = 12       /* stores the split-off bundle (slot-indexed, like OPM_OP_SPLIT's output) and resumes the main bundle */
} OpmOpKind;

enum {
    OPM_REFRACT_EMIT_CHILDREN = 1,   /* append reflected rays to `children` (warp-aggregated atomics) */
    OPM_REFRACT_CONTINUE_AS_CHILD = 2, /* mirror in a sequence: the reflected ray continues, rays without one leave the bundle
                                        (refraction_intended = false, nodes/thin_mirror.rs:209-250) */
    OPM_REFRACT_STAT_KEY = 4,        /* the launch statistic first_valid_key (Rays::bounce_lvl, rays.rs:183-194) is taken right after
                                        this op instead of at the end of the program: the ghost-focus analysis reads the bounce
                                        level between refract_on_surface and apodize (analyzers/raytrace.rs:118-133) */
    OPM_REFRACT_STAT_KEY2 = 8        /* a SECOND such statistic, taken right after this op: a launch that covers both surfaces of a
                                        lens yields the bundle's bounce level behind each of its two refractions
                                        (opm_fluence_peak_check_capture_key_async with slot 1) */
};

typedef struct {
    OpmSurface surface;
    OpmIndexModel index;          /* medium behind the surface; used when has_index != 0 */
    int32_t has_index;            /* 0 = `None`: passive surface, n2 = n of the ray (ray.rs:587) */
    int32_t refraction_intended;  /* rays.rs:979; 0 decrements the child's bounce counter (rays.rs:1016) */
    int32_t strategy;             /* OpmMissedSurfaceStrategy */
    int32_t flags;                /* OPM_REFRACT_* */
    OpmHitMap* hits;              /* optional: record (surface-frame xyz, input energy, bounce) per refracted ray */
    OpmRays* children;            /* optional: reflected bundle (appended) */
} OpmRefractOp;

/* Reflective grating: the diffracted (reflected) clone of every valid ray that hits the surface in a supported order goes
 * to `children` and / or continues in the bundle (flags OPM_REFRACT_*, as for mirrors); rays in an evanescent order are
 * invalidated; the grating vector is given in the WORLD frame, |grating_vector| = 2 pi * line density */
typedef struct {
    OpmSurface surface;
    OpmIndexModel index;          /* medium behind the surface: only validated (ray.rs:489-493) */
    double grating_vector[3];
    int32_t order;                /* diffraction order */
    int32_t refraction_intended;  /* 0 (ReflectiveGrating node) takes the bounce back from the clones (rays.rs:1081) */
    int32_t flags;                /* OPM_REFRACT_* */
    int32_t _pad;
    OpmRays* children;            /* optional: diffracted bundle (appended) */
} OpmDiffractOp;

typedef struct { OpmIsometry iso; OpmAperture aperture; } OpmApodizeOp; /* iso = effective_surface_iso (optic_node.rs:372-382) */
typedef struct { double min_energy; } OpmThresholdOp;
typedef struct { uint64_t max_bounces; uint64_t max_refractions; int32_t check_refractions; int32_t _pad; } OpmLimitsOp;
typedef struct OpmSpectrum OpmSpectrum;   /* a Spectrum resident on the device, see opm_spectrum_create() */
typedef struct { double transmission; const OpmSpectrum* spectrum; } OpmFilterOp;   /* spectrum != NULL: FilterType::Spectrum (ray.rs:712-721) */
typedef struct { double focal_length; OpmIsometry iso; } OpmParaxialOp;
/* split_out: slot-indexed copy with E*(1-ratio), NULL = discard it; spectrum != NULL: SplittingConfig::Spectrum (ray.rs:755-761) */
typedef struct { double ratio; OpmRays* split_out; const OpmSpectrum* spectrum; } OpmSplitOp;
typedef struct { double ratio; const OpmSpectrum* spectrum; } OpmForkOp;
typedef struct { OpmRays* out; } OpmJoinOp;                         /* out: the split-off bundle after its side chain; NULL = discard */
/* slot-indexed copy of the bundle at this point.  fields = OPM_SNAPSHOT_SPOT: only what a spot-diagram report and the bundle
 * statistics (opm_rays_spot_*, opm_rays_bounce_lvl, opm_rays_tally) read is copied -- position, energy, valid flag, id and bounce
 * count (nodes/spot_diagram.rs:101-163, rays.rs:183-194, 521-701): 41 of the 93 bytes of a ray; the other arrays of `out`
 * (direction, wavelength, path length, index, refraction count) are left as they were.
 * spot_frame: the frame the copy's statistics will be asked for in (opm_rays_spot_*, `frame_or_null` there; the detector's
 * effective isometry, nodes/spot_diagram.rs:106-112) or NULL for none / world.  With want_spot_sums != 0 a launch that runs the
 * specialised kernel takes the phase-1 sums of those statistics while it writes the copy (rays.rs:599-637); the next
 * opm_rays_spot_* call on `out` with the same frame then has no first pass to make. */
typedef enum { OPM_SNAPSHOT_ALL = 0, OPM_SNAPSHOT_SPOT = 1 } OpmSnapshotFields;
typedef struct { OpmRays* out; int32_t fields; int32_t want_spot_sums; const OpmIsometry* spot_frame; } OpmSnapshotOp;
typedef struct { OpmIsometry iso; int32_t inverse; int32_t _pad; } OpmTransformOp;

typedef struct {
    int32_t kind; /* OpmOpKind */
    int32_t _pad;
    union {
        OpmRefractOp refract;
        OpmApodizeOp apodize;
        OpmThresholdOp threshold;
        OpmLimitsOp limits;
        OpmFilterOp filter;
        OpmParaxialOp paraxial;
        OpmSplitOp split;
        OpmSnapshotOp snapshot;
        OpmTransformOp transform;
        OpmDiffractOp diffract;
        OpmForkOp fork;
        OpmJoinOp join;
    } u;
} OpmOp;

enum {
    OPM_TRACE_COMPACT = 1,  /* write survivors densely into `out` (warp-ballot aggregated append) instead of in place */
    OPM_TRACE_STRICT = 2    /* kernel variant compiled without FMA contraction: reference operation order */
};

/* ------------------------------------------------------------------------------------------
 * Entry points
 * ---------------------------------------------------------------------------------------- */
int32_t opm_abi_version(void);
/* sizeof() of a public structure as this library was compiled ("OpmRay", "OpmOp", "OpmSurface", "OpmAperture", "OpmTraceStats",
 * "OpmSpotStats", "OpmSourceDesc", "OpmWavefrontStats", "OpmNodeDesc", ...); 0 = unknown name.  Lets a binding check its mirror. */
uint64_t opm_abi_sizeof(const char* struct_name);
const char* opm_last_error(void);
const char* opm_status_string(OpmStatus s);

/* device: CUDA ordinal; stream: a cudaStream_t the caller owns, or NULL for a library-owned stream */
OpmStatus opm_context_create(int32_t device, void* stream, OpmContext** out);
void opm_context_destroy(OpmContext* ctx);
/* Waits for the context's stream.  Launches that do not read their statistics back (stats == NULL, OPM_SCENE_ASYNC) cannot return
 * the per-ray error conditions on which the reference aborts an analysis (`?` in rays.rs:987-991: n2 invalid, wavelength outside an
 * index model's range, non-finite hit point, apodisation factor outside (0, 1], full child bundle); their status bits are kept
 * on the device and come back HERE as the error of the lowest such bit (then cleared). */
OpmStatus opm_context_synchronize(OpmContext* ctx);
/* the same check without waiting for anything but the copy of the status word; clear != 0: forget the bits afterwards */
OpmStatus opm_context_check_errors(OpmContext* ctx, int32_t clear);
/* ray-surface interactions of all launches of the context so far (a device counter no launch resets); synchronises */
OpmStatus opm_context_interactions_total(OpmContext* ctx, uint64_t* out);
/* first_valid_key (OpmTraceStats) of the context's latest trace launch stored to dst_dev on the device, asynchronously */
OpmStatus opm_context_last_first_valid_key_async(OpmContext* ctx, uint64_t* dst_dev);
void* opm_context_stream(OpmContext* ctx);
/* number of kernels launched through this context so far (bench.py "gpu_launches") */
uint64_t opm_context_launch_count(const OpmContext* ctx);
/* measurement: while enabled every launch of the fused trace kernel is bracketed by a CUDA event pair on the
 * context's stream; opm_context_kernel_timing drains them (blocking) and returns the accumulated device time */
OpmStatus opm_context_set_kernel_timing(OpmContext* ctx, int32_t enable);
OpmStatus opm_context_kernel_timing(OpmContext* ctx, double* total_ms, uint64_t* launches, int32_t reset);

/* Run-time specialised kernels: for bundles of at least `min_rays` rays the contracted build compiles (NVRTC, cached per
 * program STRUCTURE: op kinds, geometries, coatings, index models, flags -- not the numbers) the straight-line kernel of
 * the op program instead of interpreting it.  The compiler runs on a worker thread: launches use the interpreter kernel
 * until the specialised one is ready (enable = 1) or wait for it (enable = 2); any failure falls back to the interpreter.
 * Defaults: 1, 200000 rays; environment OPM_B200_JIT=0|1|2, OPM_B200_JIT_MIN_RAYS=n.  opm_context_jit_wait blocks until no
 * compilation is in flight (benchmarks: measure the steady state). */
OpmStatus opm_context_set_jit(OpmContext* ctx, int32_t enable, uint64_t min_rays);
OpmStatus opm_context_jit_wait(OpmContext* ctx);
OpmStatus opm_context_jit_stats(const OpmContext* ctx, uint64_t* kernels_compiled, uint64_t* compile_failures, uint64_t* launches,
                                double* compile_ms);
/* build check that needs no GPU: compiles a program with every op kind to an sm_100a cubin; returns its size, 0 on failure
 * (reason in log) */
uint64_t opm_jit_selftest(char* log, uint64_t log_capacity);

/* host helpers mirroring utils/geom_transformation.rs (pure host math, no device needed) */
void opm_isometry_identity(OpmIsometry* out);
OpmStatus opm_isometry_new(const double translation[3], const double euler_xyz[3], OpmIsometry* out); /* :44-72 */
void opm_isometry_append(const OpmIsometry* a, const OpmIsometry* b, OpmIsometry* out);              /* :216-223 */
/* Isometry::new_from_view (:166-185): origin at `pos`, local z along `dir`, local y as close to `up` as orthogonality allows
 * (nalgebra Isometry3::face_towards(pos, pos + dir, up)); `up` must not be collinear with `dir` */
OpmStatus opm_isometry_from_view(const double pos[3], const double dir[3], const double up[3], OpmIsometry* out);
/* ParabolicMirror::calc_off_axis_isometry and calc_parent_focal_length (nodes/parabolic_mirror.rs:286-320): the anchor isometry of an
 * off-axis parabola (decentre, z shift, rotation of the off-axis direction about z; collimating: half a turn about the surface normal
 * at the off-axis point) and the focal length of its parent parabola.  dir_xy (0, 0) = the default direction (1, 0). */
OpmStatus opm_parabola_off_axis_isometry(double focal_length, double oa_angle_rad, const double dir_xy[2], int32_t collimating,
                                         OpmIsometry* out, double* parent_focal_length);
void opm_raytrace_config_default(OpmRayTraceConfig* out);                                             /* analyzers/raytrace.rs:315-322 */

/* ---- struct Rays (rays.rs:62-73) */
OpmStatus opm_rays_create(OpmContext* ctx, uint64_t capacity, OpmRays** out);
void opm_rays_destroy(OpmRays* rays);
uint64_t opm_rays_len(const OpmRays* rays);        /* slots in use (includes invalid and removed rays) */
/* opm_rays_len of several bundles that asynchronous launches appended children to, with one synchronisation for all of them */
OpmStatus opm_rays_refresh_lens(OpmRays* const* list, int32_t n);
/* The same without any synchronisation, for an EMPTY bundle that ONE launch emits children into (OPM_REFRACT_EMIT_CHILDREN): call
 * before that launch.  The launch then writes the reflected child of the ray in slot i to slot i of this bundle and marks the
 * slots without a child as removed rays -- which every consumer skips, like the rays a mirror drops -- so the bundle's extent is
 * the launch's (opm_rays_len, known at once), the children keep their parents' order (the reference's, rays.rs:1002-1030), and
 * the host never reads a device counter.  n_slots: capacity to reserve (the parent bundle's length). */
OpmStatus opm_rays_expect_appends(OpmRays* rays, uint64_t n_slots);
uint64_t opm_rays_capacity(const OpmRays* rays);
/* Fluence helper rays (FluenceRays, rays.rs:1470-1567; Ray::new_collimated_w_fluence_helper, ray.rs:157-166): every ray of the bundle
 * gets three satellite rays on a triangle of area wavelength^2 around it and the effective energy init_fluence * wavelength^2
 * (init_fluence_J_per_m2: HOST array of opm_rays_len values, finite and not negative).  From then on opm_trace_sequence and the
 * single-op calls carry them along: a refraction records the FLUENCE of the helper triangle as it is BEFORE the helpers reach the
 * surface in the hit map (a FluenceHitPoint, ray.rs:655-669) instead of the ray's energy, scales the effective energy by 1 - R
 * (by R on the reflected ray a mirror keeps, whose helpers stay where they were and only take their reflected directions,
 * rays.rs:996-1010), then refracts the helpers of every ray that was not missed or totally reflected; a paraxial surface
 * refracts the valid helpers of every ray (rays.rs:953-955).  Hit maps written this way are read with opm_fluence_helper_rays.
 * Bundles with helpers run the interpreter kernel; they cannot be compacted, keep a position history, emit children, be
 * diffracted or be split inside a fused program (opm_rays_copy clones the helpers, like Rays::clone).
 * opm_rays_download_fluence_helpers: 22 doubles per ray in the order of opm_rays_download -- helper positions [3][3], helper
 * directions [3][3], the helpers' refractive indices [3] (negative: that helper was set invalid), the effective energy. */
OpmStatus opm_rays_attach_fluence_helpers(OpmRays* rays, const double* init_fluence_J_per_m2);
/* Rays::new_collimated_w_fluence_helper (rays.rs:352-398): n rays along +z at pos_xyz (host, n x 3: the PositionDistribution's
 * points), energy = area of the point's Voronoi cell (create_voronoi_cells, utils/griddata.rs:265-345) x its fluence (host, n
 * values in J/m^2: the FluenceDistribution's), 0 for a cell of fewer than three vertices; each ray with its FluenceRays.
 * opm_voronoi_cell_areas: those cell areas [m^2] by themselves (NaN: fewer than three vertices). */
OpmStatus opm_rays_new_collimated_w_fluence_helper(OpmRays* rays, const double* pos_xyz, uint64_t n, double wavelength_m,
                                                   const double* fluence_J_per_m2);
OpmStatus opm_voronoi_cell_areas(OpmContext* ctx, const double* x_m, const double* y_m, uint64_t n, double* area_m2_out);
int32_t opm_rays_has_fluence_helpers(const OpmRays* rays);
OpmStatus opm_rays_download_fluence_helpers(const OpmRays* rays, double* out, uint64_t capacity_rays, uint64_t* n_out);
OpmStatus opm_rays_clear(OpmRays* rays);
/* Tells the library that (nearly) every ray of the bundle carries `wavelength_m` -- what a one-wavelength source produces
 * (opm_source_generate and opm_rays_new_collimated_with_spectrum set it themselves; copies, split-off bundles and children
 * inherit it; uploads clear it).  A performance hint only: the run-time specialised kernels then evaluate each index model
 * once per CTA for that wavelength and check every ray against it, rays of another wavelength evaluate the model themselves.
 * 0 = no hint. */
OpmStatus opm_rays_set_wavelength_hint(OpmRays* rays, double wavelength_m);
OpmStatus opm_rays_reserve(OpmRays* rays, uint64_t capacity);
OpmStatus opm_rays_upload(OpmRays* rays, const OpmRay* host, uint64_t n);                 /* Rays::from(Vec<Ray>) */
/* struct-of-arrays upload from (pinned) host memory; opl/n/bounces/refr/valid take their `Ray::new` defaults */
OpmStatus opm_rays_upload_soa(OpmRays* rays, const double* pos_xyz[3], const double* dir_xyz[3],
                              const double* e, const double* wvl, uint64_t n);
/* Rays::new_collimated_with_spectrum (rays.rs:264-297): pos_xyz = n_pos x 3 (the PositionDistribution's points), e_pos = n_pos
 * (the EnergyDistribution's energies), spectrum = (wavelength, weight) pairs; rays are expanded on the device, position-major */
OpmStatus opm_rays_new_collimated_with_spectrum(OpmRays* rays, const double* pos_xyz, const double* e_pos, uint64_t n_pos,
                                                const double* wavelengths, const double* weights, uint32_t n_wavelengths);
OpmStatus opm_rays_download(const OpmRays* rays, OpmRay* host, uint64_t capacity, uint64_t* n_out); /* skips removed rays */
OpmStatus opm_rays_copy(const OpmRays* src, OpmRays* dst);                                 /* Rays::clone */
OpmStatus opm_rays_append(OpmRays* dst, const OpmRays* src);                               /* Rays::add_rays / merge rays.rs:932,1268 */
OpmStatus opm_rays_compact(OpmRays* rays);                    /* drop removed rays, order preserving */
/* Per-ray position history (`pos_hist` ray.rs:70): once enabled, every call that pushes a position in the
 * reference logs it for the bundle's rays -- refract_on_surface (the position the ray leaves, ray.rs:616; also on total
 * reflection, not on a miss), diffract_on_periodic_surface (:537), apodize for the rays it invalidates (rays.rs:566).
 * Split-off and stored copies (OPM_OP_SPLIT / OPM_OP_SNAPSHOT, opm_rays_copy) inherit the history like a cloned Ray;
 * reflected children (refraction_intended) start without one (rays.rs:1014); uploading or generating rays clears it.
 * Bundles with a history run on the interpreter kernel, cannot be compacted and cannot use OPM_OP_FORK.
 * levels_hint = expected number of logging calls (the log grows on demand). */
OpmStatus opm_rays_enable_position_history(OpmRays* rays, uint32_t levels_hint);
OpmStatus opm_rays_disable_position_history(OpmRays* rays);
OpmStatus opm_rays_clear_position_history(OpmRays* rays);                 /* Ray::clear_pos_hist ray.rs:225-227 */
int32_t opm_rays_position_history_levels(const OpmRays* rays);            /* logging calls so far = max positions per ray; -1: disabled */
/* Ray::position_history / position_history_with_current (ray.rs:296-327) of every ray, in the order of opm_rays_download:
 * xyz_out[(ray * pos_cap + k) * 3 + c] for k < len_out[ray]; OPM_ERR_CAPACITY if a ray has more than pos_cap positions */
OpmStatus opm_rays_position_history(const OpmRays* rays, int32_t with_current, double* xyz_out, uint32_t* len_out,
                                    uint64_t ray_cap, uint32_t pos_cap, uint64_t* n_rays_out);
/* raw device pointer of one field ("px","py","pz","dx","dy","dz","e","wvl","opl","n","bounces","refr","id","valid") */
void* opm_rays_field_ptr(OpmRays* rays, const char* field);

/* Rays::refract_on_surface (rays.rs:976-1047) as a stand-alone call */
OpmStatus opm_rays_refract_on_surface(OpmRays* rays, const OpmSurface* surface, const OpmIndexModel* n2_or_null,
                                      int32_t refraction_intended, int32_t strategy,
                                      OpmRays* reflected_out_or_null, OpmHitMap* hits_or_null, OpmTraceStats* stats);
/* Rays::diffract_on_periodic_surface (rays.rs:1058-1111): in place; the diffracted clones are appended to `diffracted` */
OpmStatus opm_rays_diffract_on_periodic_surface(OpmRays* rays, const OpmSurface* surface, const OpmIndexModel* refractive_index,
                                                const double grating_vector[3], int32_t diffraction_order, int32_t refraction_intended,
                                                OpmRays* diffracted, OpmTraceStats* stats);
OpmStatus opm_rays_apodize(OpmRays* rays, const OpmAperture* aperture, const OpmIsometry* iso, int32_t* invalidated); /* rays.rs:557 */
OpmStatus opm_rays_invalidate_by_threshold_energy(OpmRays* rays, double min_energy_per_ray);   /* rays.rs:1145 */
OpmStatus opm_rays_filter_by_nr_of_bounces(OpmRays* rays, uint64_t max_bounces);               /* rays.rs:1396 */
OpmStatus opm_rays_filter_by_nr_of_refractions(OpmRays* rays, uint64_t max_refractions);       /* rays.rs:1386 */
OpmStatus opm_rays_filter_energy(OpmRays* rays, double transmission);                          /* rays.rs:1119 */
OpmStatus opm_rays_split(OpmRays* rays, double ratio, OpmRays* split_out);                     /* rays.rs:1253 */
OpmStatus opm_rays_refract_paraxial(OpmRays* rays, double focal_length, const OpmIsometry* iso); /* rays.rs:943 */
OpmStatus opm_rays_transform(OpmRays* rays, const OpmIsometry* iso, int32_t inverse);          /* Rays::transformed_by_iso */

/* fused sequence: replaces the per-node chain of the calls above (analyzers/raytrace.rs:96-140,
 * nodes/node_group/analysis_raytrace.rs:29-91).  `out` may be NULL or == rays for in-place operation. */
OpmStatus opm_trace_sequence(OpmRays* rays, const OpmOp* ops, int32_t n_ops, int32_t flags, OpmRays* out,
                             OpmTraceStats* stats);

/* ---- bundle statistics (rays.rs:521-701) */
typedef struct {
    uint64_t n_valid;           /* nr_of_rays(true)  rays.rs:535-540 */
    uint64_t n_total;           /* nr_of_rays(false) */
    double total_energy;        /* rays.rs:521-530 */
    double centroid[3];         /* rays.rs:599-612 */
    double energy_centroid[3];  /* rays.rs:618-637 */
    double beam_radius_geo;     /* rays.rs:642-660 */
    double beam_radius_rms;     /* rays.rs:666-684 */
    double energy_beam_radius_rms; /* rays.rs:690-701 */
    uint32_t bounce_lvl;        /* rays.rs:183-194 (bounce count of the valid ray with the lowest id) */
    uint32_t _pad;
} OpmSpotStats;
/* frame: NULL or the detector's effective surface isometry (positions are inverse-transformed into it,
 * nodes/spot_diagram.rs:106-112) */
OpmStatus opm_rays_spot_stats(const OpmRays* rays, const OpmIsometry* frame_or_null, OpmSpotStats* out);
/* The same statistics in two phases, for bundles sharded over several GPUs (rays.rs:599-701 need the GLOBAL centroids
 * before the radii): phase 1 fills the sums, the caller all-reduces them (sum[]: SUM, n_*: SUM, first_key: MIN), phase 2
 * fills the radius accumulators about the centroids those sums define (max_d2_mm2: MAX, sum_*: SUM), and
 * opm_spot_stats_from_partial turns the reduced accumulators into the report values. */
typedef struct {
    double sum[7];              /* x y z e*x e*y e*z e over valid rays */
    uint64_t n_valid, n_total;
    uint64_t first_key;         /* (id << 32 | bounces) of the first valid ray (lowest id); ~0 if there is none */
    double max_d2_mm2;          /* max squared distance [mm^2] from the (unweighted) centroid */
    double sum_d2_mm2;          /* sum of squared distances [mm^2] from the centroid */
    double sum_ed2;             /* sum of e * squared distance [m^2] from the energy-weighted centroid */
} OpmSpotPartial;
OpmStatus opm_rays_spot_partial_sums(const OpmRays* rays, const OpmIsometry* frame_or_null, OpmSpotPartial* out);
OpmStatus opm_rays_spot_partial_radii(const OpmRays* rays, const OpmIsometry* frame_or_null, OpmSpotPartial* inout);
void opm_spot_stats_from_partial(const OpmSpotPartial* reduced, OpmSpotStats* out);
/* Device-chained forms of the two phases: asynchronous on the context's stream, no host synchronisation; the
 * accumulators stay on the device in layouts an all-reduce can combine directly (multi-GPU: NCCL on the same buffers):
 *   sums_dev[11] = x y z e*x e*y e*z e | n_valid | n_total (all SUM) | bounce count and id of the first valid ray (-1: none;
 *                  across GPUs: the entry with the smallest id, ties to the lowest rank)
 *   radii_dev[3] = max d^2 [mm^2] (MAX) | sum d^2 [mm^2] | sum e*d^2 [m^2] (SUM)                                       */
OpmStatus opm_rays_spot_sums_async(const OpmRays* rays, const OpmIsometry* frame_or_null, double* sums_dev);
OpmStatus opm_rays_spot_radii_async(const OpmRays* rays, const OpmIsometry* frame_or_null, const double* sums_dev, double* radii_dev);
/* Rays::bounce_lvl (rays.rs:183-194): bounce count of the first valid ray (lowest id); 0 for none */
OpmStatus opm_rays_bounce_lvl(const OpmRays* rays, uint32_t* bounce_lvl);
/* per-bounce-order tallies: counts[b] valid rays and energy[b] their energy, b < n_orders (last bin = overflow) */
OpmStatus opm_rays_tally(const OpmRays* rays, uint32_t n_orders, uint64_t* counts, double* energy);

/* ---- hit map + Binning fluence (surface/hit_map/rays_hit_map.rs) */
OpmStatus opm_hitmap_create(OpmContext* ctx, uint64_t capacity, OpmHitMap** out);
void opm_hitmap_destroy(OpmHitMap* h);
uint64_t opm_hitmap_capacity(const OpmHitMap* h);
uint64_t opm_hitmap_len(const OpmHitMap* h);       /* slots written by the last pass (entries with e < 0 are empty) */
OpmStatus opm_hitmap_upload(OpmHitMap* h, const double* x, const double* y, const double* z, const double* e,
                            const uint32_t* bounce_or_null, uint64_t n);
/* packed download of the non-empty entries */
OpmStatus opm_hitmap_download(const OpmHitMap* h, double* x, double* y, double* z, double* e, uint32_t* bounce,
                              uint64_t capacity, uint64_t* n_out);
/* calc_2d_bounding_box (rays_hit_map.rs:522-562) over several hit maps (merged, hit_map/mod.rs:161-175);
 * bounce_filter < 0 = all bounces.  lrtb = left,right,top,bottom. */
OpmStatus opm_hitmaps_bounding_box(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, double lrtb[4],
                                   uint64_t* n_hits);
/* calc_fluence_with_binning (rays_hit_map.rs:330-391).  ranges: NULL (auto bounding box) or
 * {ax1.start, ax1.end, ax2.start, ax2.end}.  map_dev: DEVICE buffer of nx*ny doubles, column-major ny x nx like
 * nalgebra's DMatrix; zeroed by the call unless accumulate != 0.  lrtb_out receives the range used. */
OpmStatus opm_fluence_binning(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                              const double* ranges, const double* lrtb_override, int32_t accumulate, double* map_dev,
                              double lrtb_out[4]);
/* Device-chained forms (asynchronous, no host synchronisation).  The map range lives on the device as
 * range_dev[4] = (-left, right, top, -bottom): one element-wise MAX all-reduce merges the ranges of several GPUs and -inf
 * is the identity of a GPU without hits.  status_dev accumulates OpmStatus bits (OPM_ERR_BIN_RANGE).
 * bounce filter: >= 0 a bounce count, -1 all bounces, OPM_BOUNCE_OF_LAST_LAUNCH = Rays::bounce_lvl (rays.rs:183-194) of the
 * bundle the context's latest trace launch ran on -- its first_valid_key statistic, read ON THE DEVICE, so that trace -> fluence
 * check chains without the host (OpticSurface::evaluate_fluence_of_ray_bundle, optic_surface.rs:202-224). */
#define OPM_BOUNCE_OF_LAST_LAUNCH INT64_MIN
/* Peak fluence of ONE bundle's hits on a surface, the check the ghost-focus analysis makes per bundle and surface
 * (OpticSurface::evaluate_fluence_of_ray_bundle, optic_surface.rs:202-224 -> RaysHitMap::get_max_fluence, rays_hit_map.rs:777-798):
 * bounding box, Binning into a small nx x ny map (<= 160 KB), peak -- three launches, nothing uploaded, asynchronous.
 * acc_dev: OPM_PEAK_CHECK_ACC_BYTES of device memory prepared ONCE by opm_fluence_peak_check_acc_init (every check leaves it
 * ready for the next); map_dev: nx * ny doubles of scratch; result_dev: 8 doubles, of which [6] must be zero on entry:
 * [0..4) range (-left, right, top, -bottom; -inf = the bundle recorded no hit of that bounce level), [4] peak, [5] total,
 * [6] OpmStatus bits of the splat (u32), [7] first_valid_key of the launch when bounce_filter is OPM_BOUNCE_OF_LAST_LAUNCH (u64). */
#define OPM_PEAK_CHECK_ACC_BYTES 64
OpmStatus opm_fluence_peak_check_acc_init(OpmContext* ctx, void* acc_dev, uint64_t n);
OpmStatus opm_fluence_peak_check_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                       void* acc_dev, double* map_dev, double* result_dev);
/* The same check on the context's SIDE stream: ordered behind everything issued so far, concurrent with what the context is given
 * afterwards (the next bundle's trace launch), so a chain of trace -> check -> trace -> check keeps only the traces on the
 * critical path.  With OPM_BOUNCE_OF_LAST_LAUNCH the check takes its own copy of the launch statistic along (result_dev[7]).
 * acc_dev and map_dev must be used by side-stream checks only (they run one after the other).  opm_context_join_side_stream orders
 * the main stream behind the checks issued so far (before result_dev is read); opm_context_synchronize waits for both streams. */
OpmStatus opm_fluence_peak_check_side_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                            void* acc_dev, double* map_dev, double* result_dev);
OpmStatus opm_context_join_side_stream(OpmContext* ctx);
/* MANY such checks in three launches -- for a caller that can defer them, as the ghost-focus analysis can: the verdicts of
 * evaluate_fluence_of_ray_bundle steer no control flow and the hit maps stay until the surfaces are reset.  Check k takes
 * maps_per_check[k] (1 to 4; NULL: one each) consecutive hit maps of the flat list `maps` -- the maps of one bundle id on one surface,
 * summed into one fluence map like HitMap::get_rays_hit_map's --, accumulator k of acc_dev (n_checks x OPM_PEAK_CHECK_ACC_BYTES, prepared once by opm_fluence_peak_check_acc_init), map k of
 * maps_dev (n_checks maps of nx * ny doubles, each padded to a multiple of 8 doubles) and the result slot results_dev[k] (8 doubles
 * as above, [6] zero on entry).  bounce_filters: NULL = every check uses the key word [7] of its own result slot, which
 * opm_fluence_peak_check_capture_key_async put there right after the bundle's trace launch (Rays::bounce_lvl at the flagged op,
 * analyzers/raytrace.rs:118-133); else one filter per check (a level, -1, or OPM_BOUNCE_OF_LAST_LAUNCH for the captured key). */
/* capture: no copy of its own -- the statistic stays in place until the context's NEXT trace launch, whose prologue writes it to
 * result_dev[7] before it resets it; a batch launch, opm_context_join_side_stream and opm_context_synchronize deliver what is left */
OpmStatus opm_fluence_peak_check_capture_key_async(OpmContext* ctx, double* result_dev, int32_t slot /* 0: OPM_REFRACT_STAT_KEY (or the
                                                   end of the program), 1: OPM_REFRACT_STAT_KEY2 */);
OpmStatus opm_fluence_peak_check_batch_async(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks,
                                             const int64_t* bounce_filters, uint64_t nx, uint64_t ny, void* acc_dev, double* maps_dev,
                                             double* const* results_dev);
/* the batch on the context's side stream (see opm_fluence_peak_check_side_async): behind everything issued so far, beside what
 * follows; acc_dev / maps_dev then belong to side-stream batches only */
OpmStatus opm_fluence_peak_check_batch_side_async(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks,
                                                  const int64_t* bounce_filters, uint64_t nx, uint64_t ny, void* acc_dev, double* maps_dev,
                                                  double* const* results_dev);
OpmStatus opm_hitmaps_bounding_box_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter_or_minus1, double* range_dev);
OpmStatus opm_fluence_binning_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter_or_minus1, uint64_t nx, uint64_t ny,
                                    const double* range_dev, int32_t accumulate, double* map_dev, uint32_t* status_dev);
/* peak_total_dev[2] = peak fluence (J/m^2), total energy (J) */
OpmStatus opm_fluence_peak_total_async(OpmContext* ctx, const double* map_dev, uint64_t nx, uint64_t ny, const double* range_dev,
                                       double* peak_total_dev);
/* HitMap::calc_fluence_map(.., FluenceEstimator::KDE): RaysHitMap::calc_fluence_with_kde (rays_hit_map.rs:575-613) with
 * Kde::bandwidth_estimate / kde_2d (kde/mod.rs:83-134, Gaussian2D kde/gaussian.rs).  ranges: NULL (bounding box + 3 band
 * widths) or {r1.start, r1.end, r2.start, r2.end}, read as (left, right, top, bottom) like the reference does.  O(hits^2)
 * memory for the bandwidth estimate: at most 50000 hit points. */
OpmStatus opm_fluence_kde(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter_or_minus1, uint64_t nx, uint64_t ny,
                          const double* ranges_or_null, double* map_dev, double lrtb_out[4], double* band_width_out);
/* RaysHitMap::calc_fluence_with_voronoi (rays_hit_map.rs:407-521), the reference's DEFAULT estimator (fluence_detector/mod.rs:61,
 * analyzers/ghostfocus.rs:61), of ONE rays hit map (the hits that pass bounce_filter): per-hit Voronoi cells inside the convex
 * hull pushed outwards (create_voronoi_cells, griddata.rs:265-345), fluence = E / cell area, natural-neighbour interpolation
 * onto the grid (griddata.rs:387-436; NaN outside the convex hull).  BOTH axes take nx points (the reference uses
 * nr_of_points.0 for the second axis too, :470 and :489): map_dev holds nx * nx doubles, column-major (x index * nx + y index),
 * J/m^2.  ranges: NULL (the hits' limits) or {ax1.start, ax1.end, ax2.start, ax2.end} in metres.  accumulate != 0 adds to
 * map_dev, which is how HitMap::calc_combined_fluence_with_voronoi (hit_map/mod.rs:232-262) sums the estimates of the rays hit
 * maps of one surface on their common bounding box.  Synchronous. */
/* RaysHitMap::calc_fluence_with_helper_rays (rays_hit_map.rs:627-722) of ONE rays hit map whose records carry fluences (written by
 * the trace of a bundle with fluence helper rays): natural-neighbour interpolation of the recorded fluences, with voronator's four
 * helper sites around the bounding polygon at the value 0 (:700-701), onto nx x nx points; arguments as opm_fluence_voronoi.
 * HitMap::calc_combined_fluence_with_helper_rays (hit_map/mod.rs:283-310) is the sum over the rays hit maps of a surface on their
 * common bounding box (accumulate != 0). */
OpmStatus opm_fluence_helper_rays(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter_or_minus1, uint64_t nx,
                                  const double* ranges_or_null, int32_t accumulate, double* map_dev, double lrtb_out[4]);
OpmStatus opm_fluence_voronoi(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter_or_minus1, uint64_t nx,
                              const double* ranges_or_null, int32_t accumulate, double* map_dev, double lrtb_out[4]);
/* bit b of *mask is set when a recorded hit of the map has bounce count b (b < 64): the rays hit maps a slot-indexed hit map holds
 * (hit_map/mod.rs:101-115 files hits under their bounce count) */
OpmStatus opm_hitmap_bounce_levels(OpmHitMap* map, uint64_t* mask);
/* FluenceData::new peak (fluence_data.rs:45-68) and total_energy (:130-141) of a device map */
OpmStatus opm_fluence_peak_total(OpmContext* ctx, const double* map_dev, uint64_t nx, uint64_t ny, const double lrtb[4],
                                 double* peak, double* total_energy);
/* ---- spectra, spectrometer, wavefront (spectrum.rs, nodes/spectrometer.rs, nodes/wavefront.rs) ------------
 * A Spectrum (spectrum.rs:35-37) is a pair of arrays: ascending wavelength slots in micrometers and values in 1/um. */
/* device-resident copy for the spectrum-driven ops (OpmFilterOp / OpmSplitOp / OpmForkOp); n >= 1, slots ascending */
OpmStatus opm_spectrum_create(OpmContext* ctx, const double* lambda_um, const double* values, uint64_t n, OpmSpectrum** out);
void opm_spectrum_destroy(OpmSpectrum* spectrum);
/* host helpers on a spectrum's arrays: Spectrum::total_energy (:289-298, Kahan), center_wavelength (:304-316, metres),
 * get_value (:321-349; returns 1 = Some(*out), 0 = None) */
double opm_spectrum_total_energy(const double* lambda_um, const double* values, uint64_t n);
double opm_spectrum_center_wavelength(const double* lambda_um, const double* values, uint64_t n);
int32_t opm_spectrum_get_value(const double* lambda_um, const double* values, uint64_t n, double wavelength_m, double* out);
/* host-side construction and combination of spectra (no device work; errors = OpossumError::Spectrum -> OPM_ERR_SPECTRUM).
 * Arrays: lambda_um ascending slots, values per slot; a NULL lambda_um asks for the number of slots only (*n_out).
 *  opm_spectrum_new                Spectrum::new (spectrum.rs:47-73): slots start + i * resolution, values 0
 *  opm_spectrum_add_single_peak    Spectrum::add_single_peak (:193-227); a wavelength outside the range leaves the spectrum as it is
 *  opm_spectrum_from_laser_lines   Spectrum::from_laser_lines (:126-148)
 *  opm_spectrum_add_lorentzian_peak Spectrum::add_lorentzian_peak (:249-283)
 *  opm_spectrum_scale_vertical     Spectrum::scale_vertical (:355-367) = Spectrum::filter_with_type(FilterType::Constant) (:445-453)
 *  opm_spectrum_resample           Spectrum::resample (:377-428): the values of the source spectrum on the slots of (lambda_um, values)
 *  opm_spectrum_combine            Spectrum::filter (:430-439), ::add (:483-492), ::sub (:497-507), ::split_by_spectrum (:456-472; values keep
 *                                  the transmitted part, split_out[n] takes the rest): the other spectrum is resampled onto the slots first
 *  opm_spectrum_average_resolution Spectrum::average_resolution (:172-176), metres
 *  opm_spectrum_merge              merge_spectra (:622-650) of two present spectra
 *  opm_spectrum_from_csv           Spectrum::from_csv (:89-117): `wavelength [nm];value [%]` records */
enum { OPM_SPECTRUM_FILTER = 0, OPM_SPECTRUM_ADD = 1, OPM_SPECTRUM_SUB = 2, OPM_SPECTRUM_SPLIT = 3 };
OpmStatus opm_spectrum_new(double start_m, double end_m, double resolution_m, double* lambda_um, double* values, uint64_t cap, uint64_t* n_out);
OpmStatus opm_spectrum_add_single_peak(const double* lambda_um, double* values, uint64_t n, double wavelength_m, double value);
OpmStatus opm_spectrum_from_laser_lines(const double* wavelengths_m, const double* energies, uint64_t n_lines, double resolution_m,
                                        double* lambda_um, double* values, uint64_t cap, uint64_t* n_out);
OpmStatus opm_spectrum_add_lorentzian_peak(const double* lambda_um, double* values, uint64_t n, double center_m, double width_m, double energy);
OpmStatus opm_spectrum_scale_vertical(double* values, uint64_t n, double factor);
OpmStatus opm_spectrum_resample(const double* lambda_um, double* values, uint64_t n, const double* src_lambda_um, const double* src_values,
                                uint64_t n_src);
OpmStatus opm_spectrum_combine(int32_t what, const double* lambda_um, double* values, uint64_t n, const double* other_lambda_um,
                               const double* other_values, uint64_t n_other, double* split_out);
double opm_spectrum_average_resolution(const double* lambda_um, uint64_t n);
OpmStatus opm_spectrum_merge(const double* lambda1_um, const double* values1, uint64_t n1, const double* lambda2_um, const double* values2,
                             uint64_t n2, double* lambda_um, double* values, uint64_t cap, uint64_t* n_out);
OpmStatus opm_spectrum_from_csv(const char* path, double* lambda_um, double* values, uint64_t cap, uint64_t* n_out);
/* generate_filter_spectrum (spectrum_helper.rs:124-190): ideal short- / long-pass transmission curves (step, or a sine of the given
 * width around the cut-off) on the slots of Spectrum::new(start, end, resolution) */
enum { OPM_FILTER_SHORT_PASS_STEP = 0, OPM_FILTER_SHORT_PASS_SMOOTH = 1, OPM_FILTER_LONG_PASS_STEP = 2, OPM_FILTER_LONG_PASS_SMOOTH = 3 };
OpmStatus opm_spectrum_generate_filter(int32_t kind, double start_m, double end_m, double resolution_m, double cut_off_m, double width_m,
                                       double* lambda_um, double* values, uint64_t cap, uint64_t* n_out);
/* smallest / largest wavelength and number of the valid rays (multi-GPU: reduce with min / max / sum, then pass the common
 * range to opm_rays_to_spectrum on every rank and sum the values) */
OpmStatus opm_rays_wavelength_range(const OpmRays* rays, double min_max_m[2], uint64_t* n_valid);
/* Rays::to_spectrum (rays.rs:1209-1223) = Spectrum::from_laser_lines (spectrum.rs:126-148) of the valid rays: slots from
 * the smallest wavelength to the largest + 2 resolutions, every ray's energy spread over its two neighbouring slots
 * (add_single_peak, :193-227).  range_m: NULL (this bundle's own) or {min, max}.  lambda_um == NULL: only *n_out. */
OpmStatus opm_rays_to_spectrum(const OpmRays* rays, double resolution_m, const double* range_m, double* lambda_um, double* values,
                               uint64_t cap, uint64_t* n_out);
/* Rays::filter_energy(FilterType::Spectrum) (rays.rs:1119-1133, ray.rs:712-721), Rays::split(SplittingConfig::Spectrum)
 * (rays.rs:1253-1264, ray.rs:752-772): a ray whose wavelength is outside the spectrum fails the call (OPM_ERR_SPECTRUM) */
OpmStatus opm_rays_filter_energy_spectrum(OpmRays* rays, const OpmSpectrum* spectrum);
OpmStatus opm_rays_split_spectrum(OpmRays* rays, const OpmSpectrum* spectrum, OpmRays* split_out);
/* Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802) + WaveFrontErrorMap::calc_wavefront_statistics
 * (nodes/wavefront.rs:145-164): per valid ray (x [mm], y [mm] in the monitor frame, wavefront error in wavelengths relative
 * to the ray closest to the monitor axis), peak-to-valley and rms (about the mean).  wavelength_m <= 0: the centre wavelength
 * of to_spectrum(1 nm), what WaveFront::node_report uses (nodes/wavefront.rs:180, rays.rs:719-726).
 * xyw_out: NULL or cap rows of 3 doubles (bundle order). */
typedef struct { double wavelength_m, ptv, rms; uint64_t n_rays; } OpmWavefrontStats;
OpmStatus opm_rays_wavefront_error(const OpmRays* rays, const OpmIsometry* monitor_iso, double wavelength_m, double* xyw_out,
                                   uint64_t cap, OpmWavefrontStats* out);

/* ---- multi-GPU detector read-out (host side: opossum_b200/sharding.py) ----------------------------------------
 * Every rank all-gathers a small block, then combines the `world` rows in rank order with one launch (bit-identical on all ranks):
 *  block A = [4 x n_fluence map ranges as (-l, r, t, -b) | 11 x n_spot spot sums of opm_rays_spot_sums_async]: ranges by max,
 *            sums by sum, the first-valid-ray pair from the rank holding the smallest ray id;
 *  block B = [maps_len fluence bins | 3 x n_spot radius accumulators of opm_rays_spot_radii_async]: bins and sums by sum, max by max. */
OpmStatus opm_readout_combine_ranges_sums(OpmContext* ctx, const double* gathered_dev, int32_t world, int32_t n_fluence, int32_t n_spot,
                                          double* out_dev);
OpmStatus opm_readout_combine_maps_radii(OpmContext* ctx, const double* gathered_dev, int32_t world, uint64_t maps_len, int32_t n_spot,
                                         double* out_dev);

/* ---- multi-GPU read-out of a LARGE fluence map over NVLink peer memory (opossum_b200/csrc/opm_peer.cu) ---------------
 * The merge of the per-rank hit maps (hit_map/mod.rs:363-379) with the common data-dependent range (rays_hit_map.rs:522-562) when
 * the hits are sharded over the GPUs of one box, one process per GPU.  Instead of an all-reduce of the whole map, every rank bins
 * into the map columns its own hits span (its window), owns a slice of the columns and sums the overlapping windows of its peers
 * by reading their memory directly; the ranks synchronise through flags in each other's memory.  Results: range, peak and total
 * on every rank, the map distributed by column slices (gathered on demand).
 *   create  : allocates this rank's exchange segment (control block + two buffers of map_bins doubles) and returns its CUDA IPC
 *             handle (OPM_IPC_HANDLE_BYTES bytes);
 *   connect : `handles` = the handles of all ranks in rank order (exchanged by the launcher: torch.distributed, MPI, a file);
 *             maps every peer's segment into this process.  world == 1 needs no handles (NULL). */
#define OPM_IPC_HANDLE_BYTES 64
typedef struct OpmPeerGroup OpmPeerGroup;
OpmStatus opm_peer_group_create(OpmContext* ctx, int32_t rank, int32_t world, uint64_t map_bins, OpmPeerGroup** out,
                                uint8_t handle_out[OPM_IPC_HANDLE_BYTES]);
OpmStatus opm_peer_group_connect(OpmPeerGroup* group, const uint8_t* handles);
void opm_peer_group_destroy(OpmPeerGroup* group);
/* one meeting of the ranks: all-gather of n <= 32 doubles (out_dev: world x n), also a barrier; asynchronous on the stream.
 * A rank that waits longer than the group's time limit (default 20 s) gives up and raises OPM_ERR_PEER_TIMEOUT in the sticky
 * status (opm_context_synchronize) instead of hanging the GPU. */
OpmStatus opm_peer_allgather_async(OpmPeerGroup* group, const double* block_dev, int32_t n, double* out_dev);
/* FluenceDetector report (Binning) of hit maps sharded over the ranks: range_dev[4] = (-left, right, top, -bottom) and
 * peak_total_dev[2] on every rank; asynchronous.  Three meetings in three launches when the trace kernel kept the hit map's
 * bounding box: the kernel that zero-fills this rank's window carries meeting 1, the splat follows, the slice-reduce kernel
 * carries meetings 2 and 3.  status_dev accumulates OpmStatus bits like opm_fluence_binning_async.
 * host_report (may be NULL): 8 doubles of pinned, device-mapped host memory that the last kernel fills itself --
 * [0..4) range, [4] peak, [5] total, [6] *status_dev, [7] a sequence number that changes with every report, stored last -- so a
 * caller that only wants the report needs no device->host copy in the stream. */
OpmStatus opm_peer_fluence_report_async(OpmPeerGroup* group, OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx,
                                        uint64_t ny, double* range_dev, double* peak_total_dev, uint32_t* status_dev,
                                        double* host_report);
/* diagnostic: device time in ms of the four phases of the latest report (meeting 1 | zero-fill of the window | splat | slice
 * reduce with meetings 2 and 3) in out_ms[0..4), then eight time stamps taken inside the kernels, in ms after the start of the
 * splat kernel: [4] 0, [5] own range posted, [6] all ranges arrived, [7] slice reduce starts, [8] all parts binned, [9] last CTA
 * of the reduce done, [10] all slices' peaks arrived, [11] first CTA done with its items.  enable != 0 records the phase
 * boundaries of the reports that follow; out_ms may be NULL. */
OpmStatus opm_peer_group_phase_times(OpmPeerGroup* group, int32_t enable, float out_ms[12]);
/* the whole map (nx * ny, column-major ny x nx) of the LAST report into map_dev on this rank: every rank's column slice is copied
 * from its memory.  The ranks must not start another report before the copy has run (the caller's barrier). */
OpmStatus opm_peer_fluence_gather_map(OpmPeerGroup* group, uint64_t nx, uint64_t ny, double* map_dev);

/* measured FP64 FMA throughput of the device in TFLOP/s (FMA = 2 flop): roofline denominator of the fused trace kernel */
OpmStatus opm_measure_fp64_peak(OpmContext* ctx, double* tflops);
/* device buffers for maps (so callers need no CUDA runtime of their own) */
OpmStatus opm_device_alloc(OpmContext* ctx, uint64_t bytes, void** out);
void opm_device_free(OpmContext* ctx, void* p);
OpmStatus opm_device_memset(OpmContext* ctx, void* dev, int32_t value, uint64_t bytes);  /* asynchronous on the context's stream */
OpmStatus opm_device_to_host(OpmContext* ctx, void* host, const void* dev, uint64_t bytes);
OpmStatus opm_host_to_device(OpmContext* ctx, void* dev, const void* host, uint64_t bytes);
OpmStatus opm_host_alloc_pinned(uint64_t bytes, void** out);
void opm_host_free_pinned(void* p);

/* ---- on-device sources (rays.rs:264-297 with the generators of position_distributions/) */
typedef enum { OPM_POS_GRID = 0, OPM_POS_FIBONACCI_RECT = 1, OPM_POS_FIBONACCI_ELLIPSE = 2, OPM_POS_HEXAPOLAR = 3 } OpmPosDist;
typedef enum { OPM_EDIST_UNIFORM = 0, OPM_EDIST_GENERAL_GAUSSIAN = 1 } OpmEnergyDist;
typedef struct {
    int32_t pos_dist;         /* OpmPosDist */
    int32_t energy_dist;      /* OpmEnergyDist */
    uint64_t n_positions;     /* total positions of the whole source (all ranks) */
    uint64_t grid_nx, grid_ny;/* Grid: points in x (outer) and y (inner)  grid.rs:58-92; Hexapolar: grid_nx = number of rings
                                 (n_positions = 1 + 3 rings (rings + 1), hexapolar.rs:38-56) */
    double size_x, size_y;    /* Grid/FibonacciRect: side lengths; FibonacciEllipse: radii; Hexapolar: size_x = radius */
    double total_energy;      /* J */
    double mu[2], sigma[2], power, theta; int32_t rectangular; int32_t _pad; /* General2DGaussian general_gaussian.rs:93-120 */
    double energy_norm;       /* General2DGaussian: sum of the unnormalised weights over ALL positions; 0 = compute it */
    uint32_t n_wavelengths;   /* spectrum lines per position (position-major order, rays.rs:284-289) */
    uint32_t _pad2;
    const double* wavelengths;/* HOST, n_wavelengths, metres */
    const double* weights;    /* HOST, n_wavelengths, normalised */
    OpmIsometry iso;          /* source isometry applied to the bundle (nodes/source.rs:205-210) */
} OpmSourceDesc;
/* generates positions [pos_begin, pos_end) x all wavelengths into `rays` (ids = global ray index mod 2^32) */
/* Gaussian spectrum (spectral_distribution/gaussian.rs:78-96): num wavelengths over [lo, hi] and their normalised weights (host math) */
OpmStatus opm_spectrum_gaussian(double lo, double hi, uint64_t num, double mu, double fwhm, double power, double* wavelengths,
                                double* weights);
OpmStatus opm_source_generate(OpmRays* rays, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t pos_end);
/* positions first_pos, first_pos + stride, ... (n_pos of them): a load-balanced shard of the source for rank g of G is
 * (first_pos = g, stride = G) -- rays are independent, so any partition of the source index set is a valid sharding */
OpmStatus opm_source_generate_strided(OpmRays* rays, const OpmSourceDesc* desc, uint64_t first_pos, uint64_t n_pos, uint64_t stride);
/* Rays::new_collimated_with_spectrum (rays.rs:264-297) from the PositionDistribution's (x, y) alone: pos_xy = n_pos x 2 doubles in
 * (pinned) host memory -- every position distribution of the reference yields z = 0 (position_distributions/, every generator) -- while the
 * EnergyDistribution (uniform.rs:37-41, general_gaussian.rs:93-120: a function of x and y), the spectrum and the source isometry
 * are taken from `desc` and evaluated on the device: 16 bytes per position cross PCIe instead of 32.  Of `desc` the fields
 * energy_dist, total_energy, mu / sigma / power / theta / rectangular, n_positions (positions of the WHOLE source; 0 = n_pos),
 * energy_norm (0 = the given positions are the whole source: the sum is taken on the device), the spectrum and iso are used.
 * Asynchronous like opm_rays_new_collimated_with_spectrum (copy stream; the expansion is queued when the bundle is first used). */
OpmStatus opm_rays_new_collimated_xy(OpmRays* rays, const double* pos_xy, uint64_t n_pos, const OpmSourceDesc* desc);
/* sum of unnormalised General2DGaussian weights over [pos_begin,pos_end) (for multi-GPU: allreduce, then set energy_norm) */
OpmStatus opm_source_energy_norm(OpmContext* ctx, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t pos_end,
                                 double* partial_sum);

#ifdef __cplusplus
}
#endif
#endif /* OPOSSUM_B200_H */
