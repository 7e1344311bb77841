/*
 * opm_oracle.h -- CPU ORACLE for the OPOSSUM ray-propagation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may
 * load it.  The product (opossum_b200/) never links, imports or calls it.
 *
 * It is a plain-C, array-of-structs, f64 restatement of the reference's algorithm
 * (opossum-labs/opossum v0.6.0, Rust).  Every function cites the reference
 * file:line (relative to /root/reference/opossum/src) it follows.  The reference
 * cannot be compiled in this environment (no rustc/cargo, crates not vendored),
 * so this restatement is pinned by the reference's own known-answer tests
 * (tests/test_oracle_kats.py transcribes them with their file:line).
 *
 * Third-party arithmetic restated from the published algorithms (sources not in
 * /root/reference): roots 0.0.8 (find_roots_quadratic), nalgebra 0.33.2
 * (Isometry3 / UnitQuaternion / Vector3), uom 0.37 (unit factors), kahan 0.1.4,
 * libm::modf, compiler-rt __powidf2 (f64::powi).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp (see oracle/Makefile); `fma()`
 * appears exactly where the reference uses `mul_add`.
 */
#ifndef OPM_ORACLE_H
#define OPM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ray.rs:66-96 (fields needed for parity; pos_hist/helper rays are out of scope) */
typedef struct {
    double pos[3];
    double dir[3];
    double e;        /* J */
    double wvl;      /* m */
    double opl;      /* path_length, m */
    double n;        /* refractive_index of the current medium */
    uint32_t bounces;
    uint32_t refr;
    uint32_t valid;
    uint32_t id;     /* not in the reference: index of the seed ray, used to key comparisons */
} OrcRay;            /* 96 bytes */

/* nalgebra Isometry3<f64>: unit quaternion (w,i,j,k) + translation */
typedef struct { double q[4]; double t[3]; } OrcIso3;
/* utils/geom_transformation.rs:27-31: transform + cached inverse */
typedef struct { OrcIso3 fwd; OrcIso3 inv; } OrcIsometry;

enum { ORC_GEOM_PLANE = 0, ORC_GEOM_SPHERE = 1, ORC_GEOM_CYLINDER = 2, ORC_GEOM_PARABOLA = 3 };
enum { ORC_COAT_IDEAL_AR = 0, ORC_COAT_CONSTANT_R = 1, ORC_COAT_FRESNEL = 2 };
enum { ORC_IDX_CONST = 0, ORC_IDX_SELLMEIER1 = 1, ORC_IDX_SCHOTT = 2, ORC_IDX_CONRADY = 3 };
enum { ORC_MISS_STOP = 0, ORC_MISS_IGNORE = 1 };
enum { ORC_AP_NONE = 0, ORC_AP_CIRCLE = 1, ORC_AP_RECT = 2, ORC_AP_POLYGON = 3, ORC_AP_GAUSSIAN = 4, ORC_AP_STACK = 5 };

/* error codes (messages mirror error.rs / the call sites) */
enum {
    ORC_OK = 0,
    ORC_ERR_N2_INVALID = 1,        /* ray.rs:588-592 */
    ORC_ERR_WVL_RANGE = 2,         /* refr_index_sellmeier1.rs:79-81 etc. */
    ORC_ERR_INDEX_MODEL = 3,       /* refractive_index/mod.rs:54-58 */
    ORC_ERR_COATING = 4,           /* coatings/constant_r.rs:23-27 */
    ORC_ERR_APODIZE_FACTOR = 5,    /* ray.rs:705-709 */
    ORC_ERR_THRESHOLD = 6,         /* rays.rs:1150-1154 */
    ORC_ERR_HITPOINT = 7,          /* rays_hit_map.rs:59-66 */
    ORC_ERR_SPLIT_RATIO = 8,       /* ray.rs:763-767 */
    ORC_ERR_FOCAL_LENGTH = 9,      /* rays.rs:944-948 */
    ORC_ERR_EMPTY_HITMAP = 10,     /* rays_hit_map.rs:526-531 */
    ORC_ERR_ARG = 11,
    ORC_ERR_SPECTRUM = 12,         /* OpossumError::Spectrum (spectrum.rs) and the spectrum look-ups of ray.rs:712-721,755-761 */
    ORC_ERR_EMPTY_WAVEFRONT = 13   /* nodes/wavefront.rs:121-123,146-148 */
};

typedef struct {
    int32_t geom;
    int32_t coating;
    double param;        /* radius of curvature (sphere/cylinder) or focal length (parabola) */
    double coating_r;    /* ConstantR reflectivity */
    OrcIsometry iso;     /* isometry of the geo surface (for spheres/cylinders: node_iso o T(0,0,R)) */
} OrcSurface;

typedef struct {
    int32_t type;
    int32_t _pad;
    double c[6];         /* Const: n ; Sellmeier1: k1,k2,k3,l1,l2,l3 ; Schott: a0..a5 ; Conrady: n0,a,b */
    double wvl_lo, wvl_hi; /* m, [lo,hi) */
} OrcIndexModel;

typedef struct OrcAperture {
    int32_t type;
    int32_t obstruction;   /* ApertureType::Obstruction */
    double p[4];           /* circle: r,cx,cy ; rect: w,h,cx,cy ; gaussian: sx,sy,cx,cy */
    int32_t n_sub;         /* polygon: number of triangles ; stack: number of sub-apertures */
    int32_t _pad;
    const double* tris;    /* polygon: n_sub * 6 doubles (x1,y1,x2,y2,x3,y3) */
    const struct OrcAperture* sub; /* stack members */
} OrcAperture;

/* rays_hit_map.rs:43-49 + the bounce level it is filed under (hit_map/mod.rs:101-115) */
typedef struct { double x, y, z, e; uint32_t bounce; uint32_t id; } OrcHit;

const char* orc_strerror(int code);

/* ---- isometry (utils/geom_transformation.rs) */
void orc_iso_identity(OrcIsometry* out);
void orc_iso_new(const double t[3], const double euler_xyz[3], OrcIsometry* out);       /* :44-72 */
void orc_iso_append(const OrcIsometry* a, const OrcIsometry* b, OrcIsometry* out);      /* :216-223 */
void orc_iso_from_view(const double p[3], const double dir[3], const double up[3], OrcIsometry* out); /* :166-185 */
void orc_iso_new_axis_angle(const double t[3], const double axisangle[3], OrcIsometry* out);   /* nalgebra Isometry3::new */
void orc_parabola_off_axis_isometry(double focal_length, double oa_angle, double dir_x, double dir_y, int collimating,
                                    OrcIsometry* out, double* parent_focal_length);  /* nodes/parabolic_mirror.rs:286-320 */
void orc_iso_transform_point(const OrcIsometry* iso, const double p[3], double out[3]);         /* :255-259 */
void orc_define_up_direction(const double dir[3], double up[3]);                              /* ray.rs:818-838 */
void orc_calc_new_up_direction(const double prev_dir[3], const double dir[3], double up[3]); /* ray.rs:842-860 */
void orc_iso_inverse_transform_point(const OrcIsometry* iso, const double p[3], double out[3]); /* :280-284 */
void orc_iso_transform_vector(const OrcIsometry* iso, const double v[3], double out[3]);        /* :350-352 */
void orc_iso_inverse_transform_vector(const OrcIsometry* iso, const double v[3], double out[3]);/* :371-373 */
/* 3x3 row-major rotation + translation of the forward / inverse transforms (for cross-checks only) */
void orc_iso_to_matrix(const OrcIso3* iso, double rot[9], double t[3]);

/* ---- third-party arithmetic */
int orc_find_roots_quadratic(double a2, double a1, double a0, double roots[2]);  /* returns 0,1,2 */
double orc_powi(double a, int b);

/* ---- models */
int orc_refractive_index(const OrcIndexModel* m, double wvl, double* n_out);  /* refractive_index/mod.rs:41-60 */
double orc_fresnel_reflectivity(const double dir[3], const double normal[3], double n1, double n2); /* coatings/fresnel.rs:16-34 */
int orc_coating_reflectivity(const OrcSurface* s, const double dir[3], const double normal[3], double n1, double n2, double* r_out);
double orc_apodization_factor(const OrcAperture* ap, double x, double y);  /* aperture.rs:92-101 */

/* ---- surfaces: returns 1 on hit (point + un-normalised normal in world frame), 0 on miss */
int orc_intersect(const OrcSurface* s, const double pos[3], const double dir[3], double q[3], double nrm[3]); /* geo_surface.rs:22-32 */

/* ---- per-ray (ray.rs) */
int orc_ray_new(const double pos[3], const double dir[3], double wvl, double e, OrcRay* out);   /* :108-140 */
/* :580-688. has_n2: Option<f64>.  Returns ORC_OK/err; *child_out valid iff *has_child; *hit_out valid iff *has_hit */
int orc_ray_refract_on_surface(OrcRay* ray, const OrcSurface* s, int has_n2, double n2, int strategy,
                               OrcRay* child_out, int* has_child, OrcHit* hit_out, int* has_hit);
int orc_ray_refract_paraxial(OrcRay* ray, double focal_length, const OrcIsometry* iso);          /* :443-471 */
/* :482-556 reflective grating: gvec = grating vector in the world frame, |gvec| = 2 pi line density */
int orc_ray_diffract_on_periodic_surface(OrcRay* ray, const OrcSurface* s, double n2, const double gvec[3], int32_t order,
                                         OrcRay* child_out, int* has_child);

/* ---- bundles (rays.rs).  All loops are over `n` rays in order; OpenMP variants keep the order. */
typedef struct {
    uint64_t n_children;
    uint64_t n_hits;
    uint64_t n_interactions; /* valid rays entering the call */
    int32_t rays_missed;     /* "rays totally reflected or missed a surface" rays.rs:1026-1028 */
    int32_t valid_rays_found;/* rays.rs:1029-1031 */
} OrcRefractStats;

/* rays.rs:976-1047.  idx==NULL -> n2=None (passive).  children/hits must hold n entries (may be NULL to discard). */
int orc_rays_refract_on_surface(OrcRay* rays, uint64_t n, const OrcSurface* s, const OrcIndexModel* idx,
                                int refraction_intended, int strategy,
                                OrcRay* children, OrcHit* hits, OrcRefractStats* stats, int n_threads);
int orc_rays_apodize(OrcRay* rays, uint64_t n, const OrcAperture* ap, const OrcIsometry* iso, int* invalidated, int n_threads); /* :557-573 */
int orc_rays_threshold(OrcRay* rays, uint64_t n, double min_energy, int n_threads);   /* :1145-1162 */
void orc_rays_filter_bounces(OrcRay* rays, uint64_t n, uint64_t max_bounces);          /* :1396-1404 */
void orc_rays_filter_refractions(OrcRay* rays, uint64_t n, uint64_t max_refr);         /* :1386-1394 */
int orc_rays_filter_energy(OrcRay* rays, uint64_t n, double transmission);             /* :1119-1133 */
int orc_rays_split(OrcRay* rays, uint64_t n, double ratio, OrcRay* split_out);         /* :1253-1264 */
int orc_rays_refract_paraxial(OrcRay* rays, uint64_t n, double focal_length, const OrcIsometry* iso); /* :943-959 */
int orc_rays_diffract_on_periodic_surface(OrcRay* rays, uint64_t n, const OrcSurface* s, const OrcIndexModel* idx, const double gvec[3],
                                          int32_t order, int refraction_intended, OrcRay* children, OrcRefractStats* stats); /* :1058-1111 */
void orc_rays_transform(OrcRay* rays, uint64_t n, const OrcIsometry* iso, int inverse);/* ray.rs:787-804 */

/* position history (`pos_hist`, ray.rs:70): while attached, the pushes of ray.rs:537,616 and rays.rs:566 on rays of
 * [base, base + n) are appended to xyz[(i * cap + k) * 3 + c], k < len[i]; detach returns 1 if a list overflowed */
void orc_history_attach(const OrcRay* base, uint64_t n, double* xyz, uint32_t* len, uint32_t cap);
int orc_history_detach(void);

/* statistics (rays.rs:521-701) */
double orc_rays_total_energy(const OrcRay* rays, uint64_t n);                          /* Kahan, valid only */
uint64_t orc_rays_count(const OrcRay* rays, uint64_t n, int valid_only);
int orc_rays_centroid(const OrcRay* rays, uint64_t n, double c[3]);                    /* returns 0 if none */
int orc_rays_energy_weighted_centroid(const OrcRay* rays, uint64_t n, double c[3]);
int orc_rays_beam_radius_geo(const OrcRay* rays, uint64_t n, double* r_m);
int orc_rays_beam_radius_rms(const OrcRay* rays, uint64_t n, double* r_m);
int orc_rays_energy_weighted_beam_radius_rms(const OrcRay* rays, uint64_t n, double* r_m);
uint32_t orc_rays_bounce_lvl(const OrcRay* rays, uint64_t n);                          /* :183-194 */

/* ---- hit map / fluence (surface/hit_map/rays_hit_map.rs) */
int orc_hits_bounding_box(const OrcHit* hits, uint64_t n, double lrtb[4]);            /* :522-562 margin 0 */
/* :330-391.  ranges: NULL or {r1.start,r1.end,r2.start,r2.end}.  map: ny rows x nx cols, column-major. */
int orc_fluence_binning(const OrcHit* hits, uint64_t n, uint64_t nx, uint64_t ny, const double* ranges,
                        double* map, double lrtb_out[4]);
double orc_fluence_peak(const double* map, uint64_t n);                                /* fluence_data.rs:45-68 */
double orc_fluence_total_energy(const double* map, uint64_t nx, uint64_t ny, const double lrtb[4]); /* :130-141 */

/* ---- kernel density estimator (kde/mod.rs, kde/gaussian.rs): the KDE fluence estimator, SURVEY 8f-1 */
double orc_kde_distances_iqr(const double* values, uint64_t n);                       /* kde/mod.rs:65-81 */
int orc_kde_point_distances(const OrcHit* hits, uint64_t n, double* distances_or_null, double* std_dev); /* :44-64 */
double orc_kde_bandwidth_estimate(const OrcHit* hits, uint64_t n);                    /* :83-106 */
double orc_kde_value(const OrcHit* hits, uint64_t n, double band_width, double px, double py); /* :107-113, gaussian.rs:12-28 */
/* Voronoi fluence estimator (opm_voronoi.c): RaysHitMap::calc_fluence_with_voronoi, rays_hit_map.rs:407-521 */
int orc_voronoi_site_fluence(const double* x_cm, const double* y_cm, const double* e, uint64_t n, double* fluence, double* area_hull);
int orc_fluence_voronoi(const OrcHit* hits, uint64_t n, uint64_t nx, const double* ranges, double* map, double lrtb_out[4]);
/* RaysHitMap::calc_fluence_with_helper_rays, rays_hit_map.rs:627-722 (hits.e = fluence in J/m^2) */
int orc_fluence_helper_rays(const OrcHit* hits, uint64_t n, uint64_t nx, const double* ranges, double* map, double lrtb_out[4]);
int orc_fluence_kde(const OrcHit* hits, uint64_t n, uint64_t nx, uint64_t ny, const double* ranges, double* map, double lrtb_out[4],
                    double* band_width_out, int n_threads);                           /* rays_hit_map.rs:575-613 */

/* ---- sources */
void orc_grid(uint64_t nx, uint64_t ny, double side_x, double side_y, double* xyz);   /* position_distributions/grid.rs:58-92 */
void orc_fibonacci_rectangle(uint64_t n, double side_x, double side_y, double* xyz);   /* fibonacci.rs:53-65 */
void orc_fibonacci_ellipse(uint64_t n, double rx, double ry, double* xyz);             /* fibonacci.rs:113-126 */
uint64_t orc_hexapolar_count(uint32_t rings);
void orc_hexapolar(uint32_t rings, double radius, double* xyz);                        /* hexapolar.rs:38-56 */
void orc_energy_uniform(uint64_t n, double total, double* e);                          /* energy_distributions/uniform.rs:37-41 */
void orc_energy_general_gaussian(const double* xyz, uint64_t n, double total, double mux, double muy,
                                 double sx, double sy, double power, double theta, int rectangular, double* e); /* general_gaussian.rs:93-120 */
int orc_linspace(double start, double end, uint64_t num, double* out);                 /* utils/griddata.rs:211-231 */
void orc_spectrum_gaussian(double lo, double hi, uint64_t num, double mu, double fwhm, double power,
                           double* wvl, double* w);                                    /* spectral_distribution/gaussian.rs:78-96 */
/* rays.rs:264-297: position-major, wavelength-minor; E = E_pos * w */
int orc_rays_collimated_with_spectrum(const double* xyz, const double* e_pos, uint64_t n_pos,
                                      const double* wvl, const double* w, uint64_t n_wvl, OrcRay* out);

/* ---- spectrum.rs, nodes/spectrometer.rs, nodes/wavefront.rs: wavelengths of a Spectrum in micrometers */
int orc_spectrum_new(double start_m, double end_m, double res_m, double* lambda, double* data, uint64_t cap, uint64_t* n_out); /* :47-73 */
int orc_spectrum_add_single_peak(const double* lambda, double* data, uint64_t n, double wl_m, double value);                 /* :193-227 */
int orc_spectrum_from_laser_lines(const double* wl_m, const double* e, uint64_t n_lines, double res_m, double* lambda, double* data,
                                  uint64_t cap, uint64_t* n_out);                                                              /* :126-148 */
int orc_rays_to_spectrum(const OrcRay* rays, uint64_t n, double res_m, double* lambda, double* data, uint64_t cap, uint64_t* n_out); /* rays.rs:1209-1223 */
double orc_spectrum_total_energy(const double* lambda, const double* data, uint64_t n);                                       /* :289-298 */
double orc_spectrum_center_wavelength(const double* lambda, const double* data, uint64_t n);                                  /* :304-316, metres */
int orc_spectrum_get_value(const double* lambda, const double* data, uint64_t n, double wl_m, double* out);                   /* :321-349 */
/* opm_spectrum.c: the remaining Spectrum operations (line numbers: opossum/src/spectrum.rs) */
int orc_spectrum_add_lorentzian_peak(const double* lambda, double* data, uint64_t n, double center_m, double width_m, double energy); /* :249-283 */
int orc_spectrum_scale_vertical(double* data, uint64_t n, double factor);                                                     /* :355-367 */
double orc_spectrum_calc_ratio(double bucket_left, double bucket_right, double source_left, double source_right);             /* :588-606 */
void orc_spectrum_resample(const double* lambda, double* data, uint64_t n, const double* sl, const double* sd, uint64_t ns);  /* :377-428 */
int orc_spectrum_filter(const double* lambda, double* data, uint64_t n, const double* fl, const double* fd, uint64_t nf);     /* :430-439 */
int orc_spectrum_split_by_spectrum(const double* lambda, double* data, uint64_t n, const double* fl, const double* fd, uint64_t nf,
                                   double* split_out);                                                                        /* :456-472 */
int orc_spectrum_add(const double* lambda, double* data, uint64_t n, const double* ol, const double* od, uint64_t no);        /* :483-492 */
int orc_spectrum_sub(const double* lambda, double* data, uint64_t n, const double* ol, const double* od, uint64_t no);        /* :497-507 */
double orc_spectrum_average_resolution(const double* lambda, uint64_t n);                                                     /* :172-176, metres */
int orc_spectrum_merge(const double* l1, const double* d1, uint64_t n1, const double* l2, const double* d2, uint64_t n2, double* lambda,
                       double* data, uint64_t cap, uint64_t* n_out);                                                          /* :622-650 */
int orc_spectrum_from_csv(const char* path, double* lambda, double* data, uint64_t cap, uint64_t* n_out);                     /* :89-117 */
int orc_spectrum_generate_filter(int kind, double start_m, double end_m, double res_m, double cut_off_m, double width_m, double* lambda,
                                 double* data, uint64_t cap, uint64_t* n_out);                                                /* spectrum_helper.rs:124-190 */
int orc_rays_filter_energy_spectrum(OrcRay* rays, uint64_t n, const double* lambda, const double* data, uint64_t ns);         /* ray.rs:712-721 */
int orc_rays_split_spectrum(OrcRay* rays, uint64_t n, const double* lambda, const double* data, uint64_t ns, OrcRay* split_out); /* ray.rs:752-772 */
int orc_rays_wavefront_error(const OrcRay* rays, uint64_t n, double wavelength_m, const OrcIsometry* monitor_iso, double* xyw,
                             uint64_t* n_out);                                                                                 /* rays.rs:769-802 */
int orc_wavefront_statistics(const double* xyw, uint64_t n, double* ptv, double* rms);                                        /* nodes/wavefront.rs:145-164 */

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
