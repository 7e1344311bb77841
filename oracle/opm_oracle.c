/*
 * opm_oracle.c -- CPU ORACLE (test infrastructure only; see opm_oracle.h).
 *
 * Plain-C restatement of the reference's ray-propagation path.  Reference paths
 * are relative to /root/reference/opossum/src.  Arithmetic is written in the
 * reference's operation order; `fma()` appears exactly where the reference calls
 * `mul_add`.  Compile with -ffp-contract=off so nothing else is contracted.
 */
#include "opm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small vector helpers in nalgebra's operation order                         */
/* ------------------------------------------------------------------------- */

/* nalgebra Matrix::dot for 3x1: (a0*b0 + a1*b1) + a2*b2 */
static inline double dot3(const double a[3], const double b[3]) {
    double x = a[0] * b[0];
    double y = a[1] * b[1];
    double z = a[2] * b[2];
    return x + y + z;
}
/* nalgebra Matrix::cross for 3x1 */
static inline void cross3(const double a[3], const double b[3], double o[3]) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
/* norm_squared: 0 + dot(v,v); norm: sqrt(norm_squared) */
static inline double norm3(const double a[3]) { return sqrt(0.0 + dot3(a, a)); }
/* normalize(): self.unscale(self.norm()) => component-wise division */
static inline void normalize3(const double a[3], double o[3]) {
    double n = norm3(a);
    o[0] = a[0] / n; o[1] = a[1] / n; o[2] = a[2] / n;
}
/* Rust f64::signum: copysign(1, x), NaN -> NaN */
static inline double signum(double x) { return isnan(x) ? x : copysign(1.0, x); }

const char* orc_strerror(int code) {
    switch (code) {
        case ORC_OK: return "ok";
        case ORC_ERR_N2_INVALID: return "the refractive index must be >=1.0 and finite";
        case ORC_ERR_WVL_RANGE: return "wavelength outside valid range";
        case ORC_ERR_INDEX_MODEL: return "refractive index calculated by model is <1.0 or not finite";
        case ORC_ERR_COATING: return "reflectivity must be within [0.0, 1.0] and finite.";
        case ORC_ERR_APODIZE_FACTOR: return "transmission factor must be within (0.0..=1.0)";
        case ORC_ERR_THRESHOLD: return "threshold energy must be finite";
        case ORC_ERR_HITPOINT: return "hit point: value must be positive and finite / position must be finite";
        case ORC_ERR_SPLIT_RATIO: return "splitting_ratio must be within [0.0;1.0]";
        case ORC_ERR_FOCAL_LENGTH: return "focal length must be !=0.0 and finite";
        case ORC_ERR_EMPTY_HITMAP: return "could not calculate bounding box";
        case ORC_ERR_SPECTRUM: return "spectrum: wavelength outside the spectrum / no valid rays / invalid range or resolution";
        case ORC_ERR_EMPTY_WAVEFRONT: return "Empty wavefront-data vector!";
        default: return "invalid argument";
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* nalgebra 0.33.2: UnitQuaternion / Isometry3                                */
/* ------------------------------------------------------------------------- */

/* UnitQuaternion * Vector3 (nalgebra geometry/quaternion_ops.rs):
 *   t = q.vector().cross(v) * 2 ; cross = q.vector().cross(t) ; t * q.scalar() + cross + v */
static void quat_rotate(const double q[4], const double v[3], double o[3]) {
    const double qv[3] = { q[1], q[2], q[3] };
    double t[3], c[3];
    cross3(qv, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    cross3(qv, t, c);
    double r0 = t[0] * q[0] + c[0] + v[0];
    double r1 = t[1] * q[0] + c[1] + v[1];
    double r2 = t[2] * q[0] + c[2] + v[2];
    o[0] = r0; o[1] = r1; o[2] = r2;
}
/* Quaternion * Quaternion */
static void quat_mul(const double a[4], const double b[4], double o[4]) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double i = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double j = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double k = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = i; o[2] = j; o[3] = k;
}
/* Isometry3 * Point3 = rotation * p + translation */
static void iso3_point(const OrcIso3* s, const double p[3], double o[3]) {
    double r[3];
    quat_rotate(s->q, p, r);
    o[0] = r[0] + s->t[0]; o[1] = r[1] + s->t[1]; o[2] = r[2] + s->t[2];
}
static void iso3_vector(const OrcIso3* s, const double v[3], double o[3]) { quat_rotate(s->q, v, o); }
/* Isometry3::inverse: rotation.inverse(); translation = rotation_inv * (-translation) */
static void iso3_inverse(const OrcIso3* s, OrcIso3* o) {
    double qi[4] = { s->q[0], -s->q[1], -s->q[2], -s->q[3] };
    double nt[3] = { -s->t[0], -s->t[1], -s->t[2] };
    double t[3];
    quat_rotate(qi, nt, t);
    memcpy(o->q, qi, sizeof qi);
    memcpy(o->t, t, sizeof t);
}
/* Isometry3 * Isometry3 */
static void iso3_mul(const OrcIso3* a, const OrcIso3* b, OrcIso3* o) {
    double shift[3], q[4];
    quat_rotate(a->q, b->t, shift);
    quat_mul(a->q, b->q, q);
    o->t[0] = a->t[0] + shift[0]; o->t[1] = a->t[1] + shift[1]; o->t[2] = a->t[2] + shift[2];
    memcpy(o->q, q, sizeof q);
}
/* UnitQuaternion::from_rotation_matrix (row-major m[r*3+c]) */
static void quat_from_matrix(const double m[9], double q[4]) {
    double tr = m[0] + m[4] + m[8];
    if (tr > 0.0) {
        double denom = sqrt(tr + 1.0) * 2.0;
        q[0] = 0.25 * denom;
        q[1] = (m[7] - m[5]) / denom;
        q[2] = (m[2] - m[6]) / denom;
        q[3] = (m[3] - m[1]) / denom;
    } else if (m[0] > m[4] && m[0] > m[8]) {
        double denom = sqrt(1.0 + m[0] - m[4] - m[8]) * 2.0;
        q[0] = (m[7] - m[5]) / denom;
        q[1] = 0.25 * denom;
        q[2] = (m[1] + m[3]) / denom;
        q[3] = (m[2] + m[6]) / denom;
    } else if (m[4] > m[8]) {
        double denom = sqrt(1.0 + m[4] - m[0] - m[8]) * 2.0;
        q[0] = (m[2] - m[6]) / denom;
        q[1] = (m[1] + m[3]) / denom;
        q[2] = 0.25 * denom;
        q[3] = (m[5] + m[7]) / denom;
    } else {
        double denom = sqrt(1.0 + m[8] - m[0] - m[4]) * 2.0;
        q[0] = (m[3] - m[1]) / denom;
        q[1] = (m[2] + m[6]) / denom;
        q[2] = (m[5] + m[7]) / denom;
        q[3] = 0.25 * denom;
    }
}

void orc_iso_identity(OrcIsometry* out) {
    memset(out, 0, sizeof *out);
    out->fwd.q[0] = 1.0;
    out->inv.q[0] = 1.0;
}
static void iso_from_transform(const OrcIso3* t, OrcIsometry* out) { /* geom_transformation.rs:226-229 */
    out->fwd = *t;
    iso3_inverse(t, &out->inv);
}
/* geom_transformation.rs:44-72: Rotation3::from_euler_angles(roll=ax, pitch=ay, yaw=az) */
void orc_iso_new(const double t[3], const double e[3], OrcIsometry* out) {
    double sr = sin(e[0]), cr = cos(e[0]);
    double sp = sin(e[1]), cp = cos(e[1]);
    double sy = sin(e[2]), cy = cos(e[2]);
    double m[9] = {
        cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
        sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
        -sp,     cp * sr,                cp * cr
    };
    OrcIso3 tr;
    quat_from_matrix(m, tr.q);
    tr.t[0] = t[0]; tr.t[1] = t[1]; tr.t[2] = t[2];
    iso_from_transform(&tr, out);
}
void orc_iso_append(const OrcIsometry* a, const OrcIsometry* b, OrcIsometry* out) { /* :216-223 */
    OrcIso3 t;
    iso3_mul(&a->fwd, &b->fwd, &t);
    iso_from_transform(&t, out);
}
/* nalgebra 0.33.2 Isometry3::new(translation, axisangle): rotation by |axisangle| about axisangle / |axisangle|
 * (UnitQuaternion::from_scaled_axis: w = cos(angle / 2), (i, j, k) = axis * sin(angle / 2)); used by the off-axis parabola */
void orc_iso_new_axis_angle(const double t[3], const double axisangle[3], OrcIsometry* out) {
    OrcIso3 tr;
    const double angle = norm3(axisangle);
    if (angle == 0.0) { tr.q[0] = 1.0; tr.q[1] = tr.q[2] = tr.q[3] = 0.0; }
    else {
        const double h = 0.5 * angle, sn = sin(h);
        tr.q[0] = cos(h);
        tr.q[1] = axisangle[0] / angle * sn; tr.q[2] = axisangle[1] / angle * sn; tr.q[3] = axisangle[2] / angle * sn;
    }
    tr.t[0] = t[0]; tr.t[1] = t[1]; tr.t[2] = t[2];
    iso_from_transform(&tr, out);
}
/* ParabolicMirror::calc_off_axis_isometry / calc_parent_focal_length (nodes/parabolic_mirror.rs:286-320) */
void orc_parabola_off_axis_isometry(double focal_length, double oa_angle, double dir_x, double dir_y, int collimating,
                                    OrcIsometry* out, double* parent_focal_length) {
    const double tan_val = tan(oa_angle / 2.);
    const double parent_f = focal_length / fma(tan_val, tan_val, 1.);
    const double decenter_x = sin(oa_angle) * focal_length;
    const double z_shift = focal_length * (1. / fma(tan_val, tan_val, 1.) - cos(oa_angle));
    const double z_rot_angle = atan2(dir_y, dir_x);
    const double zero[3] = { 0, 0, 0 }, tr[3] = { decenter_x, 0., z_shift }, rz[3] = { 0., 0., z_rot_angle };
    OrcIsometry iso, rot_iso, tot;
    orc_iso_new(tr, zero, &iso);        /* Isometry::new_translation */
    orc_iso_new(zero, rz, &rot_iso);    /* Isometry::new_rotation */
    orc_iso_append(&rot_iso, &iso, &tot);
    if (collimating) {
        const double nv[3] = { decenter_x, 0., -2. * parent_f };
        double tn[3], axis[3], rot_v[3];
        orc_iso_transform_vector(&tot, nv, tn);
        normalize3(tn, axis);
        rot_v[0] = axis[0] * M_PI; rot_v[1] = axis[1] * M_PI; rot_v[2] = axis[2] * M_PI;
        OrcIsometry rot, t2;
        orc_iso_new_axis_angle(zero, rot_v, &rot);
        orc_iso_append(&rot, &tot, &t2);
        tot = t2;
    }
    *out = tot;
    if (parent_focal_length) *parent_focal_length = parent_f;
}
/* :166-185: Isometry3::face_towards(eye, eye+dir, up) */
void orc_iso_from_view(const double p[3], const double dir[3], const double up[3], OrcIsometry* out) {
    double target[3] = { p[0] + dir[0], p[1] + dir[1], p[2] + dir[2] };
    double d[3] = { target[0] - p[0], target[1] - p[1], target[2] - p[2] };
    double z[3], x[3], y[3], tmp[3];
    normalize3(d, z);
    cross3(up, z, tmp); normalize3(tmp, x);
    cross3(z, x, tmp);  normalize3(tmp, y);
    double m[9] = { x[0], y[0], z[0], x[1], y[1], z[1], x[2], y[2], z[2] };
    OrcIso3 tr;
    quat_from_matrix(m, tr.q);
    tr.t[0] = p[0]; tr.t[1] = p[1]; tr.t[2] = p[2];
    iso_from_transform(&tr, out);
}
/* ---- node positioning helpers (test infrastructure like everything here) */
/* Ray::define_up_direction (ray.rs:818-838) */
void orc_define_up_direction(const double dir_in[3], double up[3]) {
    const double X[3] = { 1, 0, 0 }, Y[3] = { 0, 1, 0 }, Z[3] = { 0, 0, 1 };
    double dir[3], c[3];
    normalize3(dir_in, dir);
    cross3(dir, X, c);
    const double nx = norm3(c);
    cross3(dir, Z, c);
    const double nz = norm3(c);
    if (nx < DBL_EPSILON || nz < DBL_EPSILON) { up[0] = 0; up[1] = 1; up[2] = 0; return; }
    cross3(dir, Y, c);
    if (norm3(c) < DBL_EPSILON) { up[0] = 1; up[1] = 0; up[2] = 0; return; }
    const double k = dot3(Y, dir);
    double proj[3] = { Y[0] - k * dir[0], Y[1] - k * dir[1], Y[2] - k * dir[2] };
    normalize3(proj, up);
}
/* Ray::calc_new_up_direction (ray.rs:842-860): up <- rotation_between(prev_dir, dir) * up when the direction changed by more
 * than 1000 eps.  nalgebra 0.33.2 Rotation3::rotation_between = scaled_rotation_between(a, b, 1): axis = unit(na x nb) (None
 * below the default epsilon), angle = acos(na . nb), Rotation3::from_axis_angle (Rodrigues matrix); parallel vectors give the
 * identity, anti-parallel ones None (the reference then leaves `up` alone). */
void orc_calc_new_up_direction(const double prev_dir[3], const double dir[3], double up[3]) {
    double d[3] = { prev_dir[0] - dir[0], prev_dir[1] - dir[1], prev_dir[2] - dir[2] };
    if (!(norm3(d) > DBL_EPSILON * 1000.)) return;  /* relative_ne!(norm, 0, epsilon = 1000 eps) */
    double na[3], nb[3], c[3];
    if (!(norm3(prev_dir) > 0.0) || !(norm3(dir) > 0.0)) return;  /* try_normalize(0) fails: identity */
    normalize3(prev_dir, na);
    normalize3(dir, nb);
    cross3(na, nb, c);
    const double cn = norm3(c);
    if (!(cn > DBL_EPSILON)) return;  /* zero angle: identity; pi: None */
    const double ux = c[0] / cn, uy = c[1] / cn, uz = c[2] / cn;
    const double angle = acos(dot3(na, nb));
    if (angle == 0.0) return;
    const double sn = sin(angle), cs = cos(angle), omc = 1.0 - cs;
    const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz;
    const double m[9] = { sqx + (1.0 - sqx) * cs,     ux * uy * omc - uz * sn, ux * uz * omc + uy * sn,
                          ux * uy * omc + uz * sn, sqy + (1.0 - sqy) * cs,     uy * uz * omc - ux * sn,
                          ux * uz * omc - uy * sn, uy * uz * omc + ux * sn, sqz + (1.0 - sqz) * cs };
    const double v[3] = { up[0], up[1], up[2] };
    for (int r = 0; r < 3; ++r) up[r] = m[3 * r] * v[0] + m[3 * r + 1] * v[1] + m[3 * r + 2] * v[2];
}
void orc_iso_transform_point(const OrcIsometry* iso, const double p[3], double o[3]) { iso3_point(&iso->fwd, p, o); }
void orc_iso_inverse_transform_point(const OrcIsometry* iso, const double p[3], double o[3]) { iso3_point(&iso->inv, p, o); }
void orc_iso_transform_vector(const OrcIsometry* iso, const double v[3], double o[3]) { iso3_vector(&iso->fwd, v, o); }
void orc_iso_inverse_transform_vector(const OrcIsometry* iso, const double v[3], double o[3]) { iso3_vector(&iso->inv, v, o); }

void orc_iso_to_matrix(const OrcIso3* iso, double rot[9], double t[3]) {
    for (int c = 0; c < 3; ++c) {
        double e[3] = { 0, 0, 0 }, r[3];
        e[c] = 1.0;
        quat_rotate(iso->q, e, r);
        rot[0 * 3 + c] = r[0]; rot[1 * 3 + c] = r[1]; rot[2 * 3 + c] = r[2];
    }
    t[0] = iso->t[0]; t[1] = iso->t[1]; t[2] = iso->t[2];
}

/* ------------------------------------------------------------------------- */
/* roots 0.0.8: find_roots_quadratic (restated from the published algorithm)  */
/* ------------------------------------------------------------------------- */
int orc_find_roots_quadratic(double a2, double a1, double a0, double r[2]) {
    if (a2 == 0.0) { /* find_roots_linear */
        if (a1 == 0.0) {
            if (a0 == 0.0) { r[0] = 0.0; return 1; }
            return 0;
        }
        r[0] = -a0 / a1;
        return 1;
    }
    double disc = a1 * a1 - 4.0 * a2 * a0;
    if (disc < 0.0) return 0;
    double a2x2 = 2.0 * a2;
    if (disc == 0.0) { r[0] = -a1 / a2x2; return 1; }
    double sq = sqrt(disc);
    double same, diff;
    if (a1 < 0.0) { same = -a1 + sq; diff = -a1 - sq; }
    else          { same = -a1 - sq; diff = -a1 + sq; }
    double x1, x2;
    if (fabs(same) > fabs(a2x2)) {
        double a0x2 = 2.0 * a0;
        if (fabs(diff) > fabs(a2x2)) { x1 = a0x2 / same; x2 = a0x2 / diff; }
        else                         { x1 = a0x2 / same; x2 = same / a2x2; }
    } else {
        x1 = diff / a2x2; x2 = same / a2x2;
    }
    if (x1 < x2) { r[0] = x1; r[1] = x2; } else { r[0] = x2; r[1] = x1; }
    return 2;
}

/* compiler-rt __powidf2 == Rust f64::powi */
double orc_powi(double a, int b) {
    const int recip = b < 0;
    double r = 1.0;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

/* ------------------------------------------------------------------------- */
/* refractive index models                                                    */
/* ------------------------------------------------------------------------- */
/* uom 0.37 (Cargo.lock:7828; not under /root/reference): Length::get::<micrometer>() = value * (1 / 1e-6) = value * 1e6 -- a
 * multiplication by the reciprocal of the unit's coefficient, not a division by it.  Pinned by the reference's own tests:
 * spectrum.rs:763-766 expects nanometer!(800.0).get::<micrometer>() == 0.8 exactly (800e-9 * 1e6 = 0.8, 800e-9 / 1e-6 =
 * 0.8000000000000002) and :757-761 expects nanometer!(380.0) -> 0.38, which rules out a division on the way in as well. */
static inline double to_micrometer(double m) { return m * 1.0E6; }

int orc_refractive_index(const OrcIndexModel* m, double wvl, double* n_out) {
    double n;
    if (m->type != ORC_IDX_CONST) { /* Range::contains: start <= x < end */
        if (!(m->wvl_lo <= wvl && wvl < m->wvl_hi)) return ORC_ERR_WVL_RANGE;
    }
    switch (m->type) {
        case ORC_IDX_CONST: /* refr_index_const.rs:42-44 */
            n = m->c[0];
            break;
        case ORC_IDX_SELLMEIER1: { /* refr_index_sellmeier1.rs:78-90 */
            double lambda = to_micrometer(wvl);
            double l_sq = lambda * lambda;
            n = sqrt(1.0 + m->c[0] * l_sq / (l_sq - m->c[3])
                         + m->c[1] * l_sq / (l_sq - m->c[4])
                         + m->c[2] * l_sq / (l_sq - m->c[5]));
            break;
        }
        case ORC_IDX_SCHOTT: { /* refr_index_schott.rs:73-92 */
            double l = to_micrometer(wvl);
            n = sqrt(fma(m->c[5], orc_powi(l, -8),
                     fma(m->c[4], orc_powi(l, -6),
                     fma(m->c[3], orc_powi(l, -4),
                     fma(m->c[2], orc_powi(l, -2),
                     fma(m->c[1], orc_powi(l, 2), m->c[0]))))));
            break;
        }
        case ORC_IDX_CONRADY: { /* refr_index_conrady.rs:57-64 */
            double l = to_micrometer(wvl);
            n = m->c[0] + (m->c[1] / l) + (m->c[2] / pow(l, 3.5));
            break;
        }
        default: return ORC_ERR_ARG;
    }
    if (n < 1.0 || !isfinite(n)) return ORC_ERR_INDEX_MODEL; /* refractive_index/mod.rs:54-58 */
    *n_out = n;
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* coatings                                                                   */
/* ------------------------------------------------------------------------- */
/* nalgebra Matrix::angle */
static double angle3(const double a[3], const double b[3]) {
    double prod = dot3(a, b);
    double n1 = norm3(a), n2 = norm3(b);
    if (n1 == 0.0 || n2 == 0.0) return 0.0;
    double cang = prod / (n1 * n2);
    if (cang < -1.0) cang = -1.0;
    if (cang > 1.0) cang = 1.0;
    return acos(cang);
}
/* coatings/fresnel.rs:16-34 */
double orc_fresnel_reflectivity(const double dir[3], const double normal[3], double n1, double n2) {
    double neg_n[3] = { -1.0 * normal[0], -1.0 * normal[1], -1.0 * normal[2] };
    double alpha = angle3(dir, neg_n);
    double s = sin(alpha);
    double beta = acos(sqrt(fma(orc_powi(n1, 2), -orc_powi(s, 2), orc_powi(n2, 2))) / n2);
    double r_s = fma(n1, cos(alpha), -(n2 * cos(beta))) / fma(n1, cos(alpha), n2 * cos(beta));
    double r_p = fma(n2, cos(alpha), -(n1 * cos(beta))) / fma(n2, cos(alpha), n1 * cos(beta));
    return fma(r_p, r_p, orc_powi(r_s, 2)) / 2.0;
}
/* coatings/mod.rs:37-57 */
int orc_coating_reflectivity(const OrcSurface* s, const double dir[3], const double normal[3],
                             double n1, double n2, double* r_out) {
    switch (s->coating) {
        case ORC_COAT_IDEAL_AR: *r_out = 0.0; return ORC_OK; /* ideal_ar.rs:13-22 */
        case ORC_COAT_CONSTANT_R: { /* constant_r.rs:22-29: !(0..=1).contains || !is_normal */
            double r = s->coating_r;
            if (!(r >= 0.0 && r <= 1.0) || fpclassify(r) != FP_NORMAL) return ORC_ERR_COATING;
            *r_out = r;
            return ORC_OK;
        }
        case ORC_COAT_FRESNEL: *r_out = orc_fresnel_reflectivity(dir, normal, n1, n2); return ORC_OK;
        default: return ORC_ERR_ARG;
    }
}

/* ------------------------------------------------------------------------- */
/* apertures (aperture.rs)                                                    */
/* ------------------------------------------------------------------------- */
double orc_apodization_factor(const OrcAperture* ap, double x, double y) {
    double t;
    switch (ap->type) {
        case ORC_AP_NONE: return 1.0;
        case ORC_AP_CIRCLE: { /* :158-180 */
            double px = x - ap->p[1], py = y - ap->p[2];
            t = (fma(py, py, orc_powi(px, 2)) <= orc_powi(ap->p[0], 2)) ? 1.0 : 0.0;
            break;
        }
        case ORC_AP_RECT: { /* :222-243 */
            double px = x - ap->p[2], py = y - ap->p[3];
            double qx = fabs(px) - ap->p[0] / 2.0;
            double qy = fabs(py) - ap->p[1] / 2.0;
            double mx = fmax(qx, 0.0), my = fmax(qy, 0.0);
            double sdf = sqrt(fma(mx, mx, orc_powi(my, 2))) + fmin(fmax(qx, qy), 0.0);
            t = (sdf <= 0.0) ? 1.0 : 0.0;
            break;
        }
        case ORC_AP_POLYGON: { /* :291-316 */
            int in = 0;
            for (int k = 0; k < ap->n_sub; ++k) {
                const double* tr = ap->tris + 6 * k;
                double p1x = tr[0], p1y = tr[1], p2x = tr[2], p2y = tr[3], p3x = tr[4], p3y = tr[5];
                double den = fma(p2y - p3y, p1x - p3x, (p3x - p2x) * (p1y - p3y));
                double a = fma(p2y - p3y, x - p3x, (p3x - p2x) * (y - p3y)) / den;
                double b = fma(p3y - p1y, x - p3x, (p1x - p3x) * (y - p3y)) / den;
                double c = 1.0 - a - b;
                if (a >= 0.0 && a <= 1.0 && b >= 0.0 && b <= 1.0 && c >= 0.0 && c <= 1.0) { in = 1; break; }
            }
            t = in ? 1.0 : 0.0;
            break;
        }
        case ORC_AP_GAUSSIAN: { /* :367-382 */
            double u = (x - ap->p[2]) / ap->p[0];
            double v = (y - ap->p[3]) / ap->p[1];
            t = exp(-0.5 * fma(u, u, orc_powi(v, 2)));
            break;
        }
        case ORC_AP_STACK: { /* :407-416 */
            t = 1.0;
            for (int k = 0; k < ap->n_sub; ++k) t *= orc_apodization_factor(&ap->sub[k], x, y);
            break;
        }
        default: return 1.0;
    }
    if (ap->obstruction) t = 1.0 - t;
    return t;
}

/* ------------------------------------------------------------------------- */
/* surfaces                                                                   */
/* ------------------------------------------------------------------------- */
/* shared root selection of sphere.rs:84-113 / cylinder.rs:72-101 */
static int pick_root_curved(int nroots, const double r[2], double radius, int back, double* t_out) {
    if (nroots == 0) return 0;
    if (nroots == 1) {
        if (r[0] >= 0.0) { *t_out = r[0]; return 1; }
        return 0;
    }
    double t;
    if (!signbit(radius)) t = back ? fmax(r[0], r[1]) : fmin(r[0], r[1]);
    else                  t = back ? fmin(r[0], r[1]) : fmax(r[0], r[1]);
    if (signbit(t)) return 0;
    *t_out = t;
    return 1;
}

static int intersect_local(const OrcSurface* s, const double p[3], const double d[3], double q[3], double nrm[3]) {
    switch (s->geom) {
        case ORC_GEOM_PLANE: { /* surface/plane.rs:43-62 */
            double pp[3] = { p[0], p[1], p[2] };
            if (pp[2] != 0.0) {
                double dz = -pp[2];
                if (signum(dz) != signum(d[2])) return 0;
                double len = dz / d[2];
                pp[0] += len * d[0]; pp[1] += len * d[1]; pp[2] += len * d[2];
            }
            q[0] = pp[0]; q[1] = pp[1]; q[2] = pp[2];
            nrm[0] = 0.0; nrm[1] = 0.0; nrm[2] = -1.0 * signum(d[2]);
            return 1;
        }
        case ORC_GEOM_SPHERE: { /* surface/sphere.rs:59-137 */
            double radius = s->param;
            int back = signbit(d[2]);
            double a = 0.0 + dot3(d, d);
            double b = 2.0 * fma(d[2], 0.0, dot3(d, p));
            double c = fma(radius, -radius, 0.0 + dot3(p, p));
            double r[2], t;
            int nr = orc_find_roots_quadratic(a, b, c, r);
            if (!pick_root_curved(nr, r, radius, back, &t)) return 0;
            q[0] = p[0] + t * d[0]; q[1] = p[1] + t * d[1]; q[2] = p[2] + t * d[2];
            normalize3(q, nrm);
            if (signbit(radius)) { if (!back) { nrm[0] *= -1.0; nrm[1] *= -1.0; nrm[2] *= -1.0; } }
            if (!signbit(radius) && back) { nrm[0] *= -1.0; nrm[1] *= -1.0; nrm[2] *= -1.0; }
            return 1;
        }
        case ORC_GEOM_CYLINDER: { /* surface/cylinder.rs:46-129 */
            double radius = s->param;
            int back = signbit(d[2]);
            double a = fma(d[0], d[0], d[2] * d[2]);
            double b = 2.0 * fma(d[0], p[0], d[2] * p[2]);
            double c = fma(radius, -radius, fma(p[0], p[0], p[2] * p[2]));
            double r[2], t;
            int nr = orc_find_roots_quadratic(a, b, c, r);
            if (!pick_root_curved(nr, r, radius, back, &t)) return 0;
            q[0] = p[0] + t * d[0]; q[1] = p[1] + t * d[1]; q[2] = p[2] + t * d[2];
            double n1[3];
            normalize3(q, n1);
            n1[1] = 0.0;
            normalize3(n1, nrm);
            if (signbit(radius)) { if (!back) { nrm[0] *= -1.0; nrm[1] *= -1.0; nrm[2] *= -1.0; } }
            if (!signbit(radius) && back) { nrm[0] *= -1.0; nrm[1] *= -1.0; nrm[2] *= -1.0; }
            return 1;
        }
        case ORC_GEOM_PARABOLA: { /* surface/parabola.rs:45-113 */
            double f = s->param;
            double a = fma(d[0], d[0], d[1] * d[1]);
            double b = 2.0 * fma(2.0 * f, -d[2], fma(p[0], d[0], p[1] * d[1]));
            double c = fma(4.0 * f, -p[2], fma(p[0], p[0], p[1] * p[1]));
            double r[2], t;
            int nr = orc_find_roots_quadratic(a, b, c, r);
            if (nr == 0) return 0;
            if (nr == 1) { if (r[0] >= 0.0) t = r[0]; else return 0; }
            else {
                t = signbit(f) ? fmax(r[0], r[1]) : fmin(r[0], r[1]);
                if (signbit(t)) return 0;
            }
            q[0] = p[0] + t * d[0]; q[1] = p[1] + t * d[1]; q[2] = p[2] + t * d[2];
            nrm[0] = q[0]; nrm[1] = q[1]; nrm[2] = -2.0 * f;
            return 1;
        }
        default: return 0;
    }
}

/* surface/geo_surface.rs:22-32 */
int orc_intersect(const OrcSurface* s, const double pos[3], const double dir[3], double q[3], double nrm[3]) {
    double p[3], d[3], ql[3], nl[3];
    iso3_point(&s->iso.inv, pos, p);
    iso3_vector(&s->iso.inv, dir, d);
    if (!intersect_local(s, p, d, ql, nl)) return 0;
    iso3_point(&s->iso.fwd, ql, q);
    iso3_vector(&s->iso.fwd, nl, nrm);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Ray (ray.rs)                                                               */
/* ------------------------------------------------------------------------- */
int orc_ray_new(const double pos[3], const double dir[3], double wvl, double e, OrcRay* out) { /* :108-140 */
    if (wvl == 0.0 || signbit(wvl) || !isfinite(wvl)) return ORC_ERR_ARG;
    if (signbit(e) || !isfinite(e)) return ORC_ERR_ARG;
    if (norm3(dir) == 0.0) return ORC_ERR_ARG;
    memset(out, 0, sizeof *out);
    memcpy(out->pos, pos, 3 * sizeof(double));
    normalize3(dir, out->dir);
    out->e = e; out->wvl = wvl; out->opl = 0.0; out->n = 1.0;
    out->bounces = 0; out->refr = 0; out->valid = 1; out->id = 0;
    return ORC_OK;
}

/* ray.rs:580-688 */
/* ---- position history (`pos_hist`, ray.rs:70).  While a sink is attached, every push of the reference
 * (refract ray.rs:616, diffract :537, apodize of a ray it invalidates rays.rs:566) on a ray of [base, base + n) is
 * appended to that ray's list xyz[(i * cap + k) * 3 + c], k < len[i].  One sink at a time (single-threaded callers). */
static struct { const OrcRay* base; uint64_t n; double* xyz; uint32_t* len; uint32_t cap; int overflow; } g_hist;
void orc_history_attach(const OrcRay* base, uint64_t n, double* xyz, uint32_t* len, uint32_t cap) {
    g_hist.base = base; g_hist.n = n; g_hist.xyz = xyz; g_hist.len = len; g_hist.cap = cap; g_hist.overflow = 0;
}
int orc_history_detach(void) {
    int ov = g_hist.overflow;
    memset(&g_hist, 0, sizeof g_hist);
    return ov;
}
static void hist_push(const OrcRay* ray, const double pos[3]) {
    if (!g_hist.base || ray < g_hist.base || ray >= g_hist.base + g_hist.n) return;
    uint64_t i = (uint64_t)(ray - g_hist.base);
    uint32_t k = g_hist.len[i];
    if (k >= g_hist.cap) { g_hist.overflow = 1; return; }
    double* d = g_hist.xyz + ((size_t)i * g_hist.cap + k) * 3;
    d[0] = pos[0]; d[1] = pos[1]; d[2] = pos[2];
    g_hist.len[i] = k + 1;
}

int orc_ray_refract_on_surface(OrcRay* ray, const OrcSurface* s, int has_n2, double n2, int strategy,
                               OrcRay* child_out, int* has_child, OrcHit* hit_out, int* has_hit) {
    *has_child = 0;
    *has_hit = 0;
    double n2v = has_n2 ? n2 : ray->n;
    if (n2v < 1.0 || !isfinite(n2v)) return ORC_ERR_N2_INVALID;
    double q[3], normal[3];
    if (orc_intersect(s, ray->pos, ray->dir, q, normal)) {
        double mu = ray->n / n2v;
        double s1[3], N[3], cr[3];
        normalize3(ray->dir, s1);
        normalize3(normal, N);
        cross3(N, s1, cr);
        double dis = fma(mu * mu, -dot3(cr, cr), 1.0);
        double k = 2.0 * dot3(s1, N);
        double refl[3] = { s1[0] - k * N[0], s1[1] - k * N[1], s1[2] - k * N[2] };
        double dp[3] = { ray->pos[0] - q[0], ray->pos[1] - q[1], ray->pos[2] - q[2] };
        ray->opl += ray->n * norm3(dp);
        hist_push(ray, ray->pos); /* ray.rs:616 */
        ray->pos[0] = q[0]; ray->pos[1] = q[1]; ray->pos[2] = q[2];
        if (!signbit(dis)) {
            OrcRay child = *ray;
            double R;
            int rc = orc_coating_reflectivity(s, ray->dir, normal, ray->n, n2v, &R);
            if (rc != ORC_OK) return rc;
            double E = ray->e;
            double ncr[3] = { -1.0 * cr[0], -1.0 * cr[1], -1.0 * cr[2] };
            double w[3];
            cross3(N, ncr, w);
            double root = sqrt(fma(mu * mu, -dot3(cr, cr), 1.0));
            ray->dir[0] = mu * w[0] - N[0] * root;
            ray->dir[1] = mu * w[1] - N[1] * root;
            ray->dir[2] = mu * w[2] - N[2] * root;
            ray->e = E * (1.0 - R);
            child.dir[0] = refl[0]; child.dir[1] = refl[1]; child.dir[2] = refl[2];
            child.e = E * R;
            child.bounces += 1;
            ray->n = n2v;
            if (has_n2) ray->refr += 1;
            /* hit map: surface-frame position + INPUT energy, filed under self.number_of_bounces (:640-654) */
            double hl[3];
            iso3_point(&s->iso.inv, q, hl);
            if (!isfinite(E) || signbit(E)) return ORC_ERR_HITPOINT;
            if (!isfinite(hl[0]) || !isfinite(hl[1]) || !isfinite(hl[2])) return ORC_ERR_HITPOINT;
            hit_out->x = hl[0]; hit_out->y = hl[1]; hit_out->z = hl[2]; hit_out->e = E;
            hit_out->bounce = ray->bounces; hit_out->id = ray->id;
            *has_hit = 1;
            *child_out = child;
            *has_child = 1;
        } else { /* total internal reflection */
            ray->bounces += 1;
            ray->dir[0] = refl[0]; ray->dir[1] = refl[1]; ray->dir[2] = refl[2];
        }
    } else {
        if (strategy == ORC_MISS_STOP) ray->valid = 0;
    }
    return ORC_OK;
}

/* Ray::diffract_on_periodic_surface (ray.rs:482-556): reflective grating, diffraction order m, grating vector G (world
 * frame, |G| = 2 pi line density).  On a supported order the ray moves to the hit point, takes the diffracted direction,
 * counts a bounce, *child_out is its clone and the ray itself is left with zero energy; an evanescent order invalidates
 * the ray; a miss leaves it untouched.  SURVEY 8f-3. */
int orc_ray_diffract_on_periodic_surface(OrcRay* ray, const OrcSurface* s, double n2, const double gvec[3], int32_t order,
                                         OrcRay* child_out, int* has_child) {
    *has_child = 0;
    if (n2 < 1.0 || !isfinite(n2)) return ORC_ERR_N2_INVALID;
    double q[3], normal[3];
    if (!orc_intersect(s, ray->pos, ray->dir, q, normal)) return ORC_OK;
    const double PI = 3.14159265358979323846264338327950288;
    double N[3];
    normalize3(normal, N);
    double ray_dir_norm = norm3(ray->dir);
    double k0_n = 2. * PI * ray->n / ray->wvl;
    double k_vec[3] = { ray->dir[0] * k0_n / ray_dir_norm, ray->dir[1] * k0_n / ray_dir_norm, ray->dir[2] * k0_n / ray_dir_norm };
    double kxn[3], k_para[3];
    cross3(k_vec, N, kxn);
    cross3(N, kxn, k_para);
    double kn = dot3(k_vec, N);
    double k_perp[3] = { N[0] * kn, N[1] * kn, N[2] * kn };
    double m = (double)order;
    double k_para_out[3] = { k_para[0] + m * gvec[0], k_para[1] + m * gvec[1], k_para[2] + m * gvec[2] };
    double kpo = norm3(k_para_out);
    double k_perp_norm_out = sqrt(fma(k0_n, k0_n, -(kpo * kpo)));  /* norm().powi(2) */
    double dp[3] = { ray->pos[0] - q[0], ray->pos[1] - q[1], ray->pos[2] - q[2] };
    ray->opl += ray->n * norm3(dp);
    double ql[3];
    iso3_point(&s->iso.inv, q, ql);
    ray->opl += m * norm3(gvec) / 2. / PI * ql[0] * ray->wvl;
    hist_push(ray, ray->pos); /* ray.rs:537 */
    ray->pos[0] = q[0]; ray->pos[1] = q[1]; ray->pos[2] = q[2];
    if (isfinite(k_perp_norm_out)) {
        double kpn[3];
        normalize3(k_perp, kpn);
        double out[3] = { -kpn[0] * k_perp_norm_out + k_para_out[0], -kpn[1] * k_perp_norm_out + k_para_out[1],
                          -kpn[2] * k_perp_norm_out + k_para_out[2] };
        normalize3(out, ray->dir);
        ray->bounces += 1;
        *child_out = *ray;
        *has_child = 1;
        ray->e = 0.;
    } else {
        ray->valid = 0;
    }
    return ORC_OK;
}

/* ray.rs:443-471 */
int orc_ray_refract_paraxial(OrcRay* ray, double f, const OrcIsometry* iso) {
    if (f == 0.0 || !isfinite(f)) return ORC_ERR_FOCAL_LENGTH;
    double p[3], d[3];
    iso3_point(&iso->inv, ray->pos, p);
    iso3_vector(&iso->inv, ray->dir, d);
    double power = 1.0 / f;
    double az = fabs(d[2]);
    d[0] /= az; d[1] /= az; d[2] /= az;
    d[0] -= power * p[0];
    d[1] -= power * p[1];
    double r_sq = fma(p[0], p[0], p[1] * p[1]);
    double f_sq = f * f;
    ray->opl -= sqrt(r_sq + f_sq) - fabs(f);
    ray->refr += 1;
    double pw[3], dw[3];
    iso3_point(&iso->fwd, p, pw);
    iso3_vector(&iso->fwd, d, dw);
    ray->pos[0] = pw[0]; ray->pos[1] = pw[1]; ray->pos[2] = pw[2];
    normalize3(dw, ray->dir);
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Rays (rays.rs)                                                             */
/* ------------------------------------------------------------------------- */
int orc_rays_refract_on_surface(OrcRay* rays, uint64_t n, const OrcSurface* s, const OrcIndexModel* idx,
                                int refraction_intended, int strategy,
                                OrcRay* children, OrcHit* hits, OrcRefractStats* stats, int n_threads) {
    uint8_t* flags = (uint8_t*)calloc(n ? n : 1, 1); /* bit0 child, bit1 hit, bit2 missed, bit3 valid-on-entry */
    OrcRay* ctmp = children ? children : (OrcRay*)malloc((n ? n : 1) * sizeof(OrcRay));
    OrcHit* htmp = hits ? hits : (OrcHit*)malloc((n ? n : 1) * sizeof(OrcHit));
    if (!flags || !ctmp || !htmp) { free(flags); if (!children) free(ctmp); if (!hits) free(htmp); return ORC_ERR_ARG; }
    int err = ORC_OK;
    uint64_t err_idx = UINT64_MAX;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) if (n_threads > 1)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        uint64_t i = (uint64_t)ii;
        OrcRay* ray = &rays[i];
        if (!ray->valid) continue;
        uint8_t f = 8;
        double n2 = 0.0;
        int rc = ORC_OK;
        if (idx) rc = orc_refractive_index(idx, ray->wvl, &n2);
        int hc = 0, hh = 0;
        if (rc == ORC_OK)
            rc = orc_ray_refract_on_surface(ray, s, idx != NULL, n2, strategy, &ctmp[i], &hc, &htmp[i], &hh);
        if (rc != ORC_OK) {
#pragma omp critical
            { if (i < err_idx) { err_idx = i; err = rc; } }
            flags[i] = f;
            continue;
        }
        if (hc) {
            f |= 1;
            if (!refraction_intended) ctmp[i].bounces -= 1; /* reduce_bounce_counter rays.rs:1016 */
        } else {
            f |= 4;
        }
        if (hh) f |= 2;
        flags[i] = f;
    }
    OrcRefractStats st;
    memset(&st, 0, sizeof st);
    /* order-preserving compaction (children are pushed in ray order, rays.rs:1019) */
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t f = flags[i];
        if (f & 8) { st.valid_rays_found = 1; st.n_interactions++; }
        if (f & 4) st.rays_missed = 1;
        if (f & 1) { if (st.n_children != i) ctmp[st.n_children] = ctmp[i]; st.n_children++; }
        if (f & 2) { if (st.n_hits != i) htmp[st.n_hits] = htmp[i]; st.n_hits++; }
    }
    if (stats) *stats = st;
    free(flags);
    if (!children) free(ctmp);
    if (!hits) free(htmp);
    return err;
}

/* rays.rs:557-573 */
int orc_rays_apodize(OrcRay* rays, uint64_t n, const OrcAperture* ap, const OrcIsometry* iso, int* invalidated, int n_threads) {
    int inv = 0, err = ORC_OK;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) if (n_threads > 1) reduction(|:inv) reduction(max:err)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        OrcRay* ray = &rays[ii];
        if (!ray->valid) continue;
        double p[3];
        iso3_point(&iso->inv, ray->pos, p);
        double f = orc_apodization_factor(ap, p[0], p[1]);
        if (f > 0.0) {
            if (!(f >= 0.0 && f <= 1.0)) { err = ORC_ERR_APODIZE_FACTOR; continue; } /* ray.rs:705-709 */
            ray->e *= f;
        } else {
            hist_push(ray, ray->pos); /* rays.rs:566 */
            ray->valid = 0;
            inv = 1;
        }
    }
    if (invalidated) *invalidated = inv;
    return err;
}

/* rays.rs:1145-1162 */
int orc_rays_threshold(OrcRay* rays, uint64_t n, double min_energy, int n_threads) {
    if (signbit(min_energy)) return ORC_OK; /* warning only, bundle unmodified */
    if (!isfinite(min_energy)) return ORC_ERR_THRESHOLD;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) if (n_threads > 1)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        if (rays[i].e < min_energy) rays[i].valid = 0;
    return ORC_OK;
}
void orc_rays_filter_bounces(OrcRay* rays, uint64_t n, uint64_t max_bounces) { /* :1396-1404 */
    for (uint64_t i = 0; i < n; ++i) if (rays[i].bounces > max_bounces) rays[i].valid = 0;
}
void orc_rays_filter_refractions(OrcRay* rays, uint64_t n, uint64_t max_refr) { /* :1386-1394 */
    for (uint64_t i = 0; i < n; ++i) if (rays[i].refr >= max_refr) rays[i].valid = 0;
}
int orc_rays_filter_energy(OrcRay* rays, uint64_t n, double t) { /* :1119-1133 + ray.rs:702-727 */
    if (!(t >= 0.0 && t <= 1.0)) return ORC_ERR_APODIZE_FACTOR;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) rays[i].e *= t;
    return ORC_OK;
}
int orc_rays_split(OrcRay* rays, uint64_t n, double ratio, OrcRay* out) { /* :1253-1264 + ray.rs:752-772 */
    for (uint64_t i = 0; i < n; ++i) {
        if (rays[i].valid) {
            if (!(ratio >= 0.0 && ratio <= 1.0)) return ORC_ERR_SPLIT_RATIO;
            out[i] = rays[i];
            rays[i].e *= ratio;
            out[i].e *= 1.0 - ratio;
        } else {
            out[i] = rays[i];
        }
    }
    return ORC_OK;
}
/* Rays::diffract_on_periodic_surface (rays.rs:1058-1111): valid rays only, n2 from the index model per ray; the diffracted
 * clones are returned in ray order (children must hold n entries); refraction_intended == 0 takes the bounce back from the
 * clones (reduce_bounce_counter), as the ReflectiveGrating node does */
int orc_rays_diffract_on_periodic_surface(OrcRay* rays, uint64_t n, const OrcSurface* s, const OrcIndexModel* idx, const double gvec[3],
                                          int32_t order, int refraction_intended, OrcRay* children, OrcRefractStats* stats) {
    OrcRefractStats st;
    memset(&st, 0, sizeof st);
    for (uint64_t i = 0; i < n; ++i) {
        if (!rays[i].valid) continue;
        double n2;
        int rc = orc_refractive_index(idx, rays[i].wvl, &n2);
        if (rc) return rc;
        OrcRay child;
        int has_child = 0;
        rc = orc_ray_diffract_on_periodic_surface(&rays[i], s, n2, gvec, order, &child, &has_child);
        if (rc) return rc;
        st.n_interactions++;
        if (has_child) {
            if (!refraction_intended) child.bounces -= 1;
            children[st.n_children++] = child;
        } else st.rays_missed = 1;
        st.valid_rays_found = 1;
    }
    if (stats) *stats = st;
    return ORC_OK;
}
int orc_rays_refract_paraxial(OrcRay* rays, uint64_t n, double f, const OrcIsometry* iso) { /* :943-959 */
    if (f == 0.0 || !isfinite(f)) return ORC_ERR_FOCAL_LENGTH;
    for (uint64_t i = 0; i < n; ++i)
        if (rays[i].valid) { int rc = orc_ray_refract_paraxial(&rays[i], f, iso); if (rc) return rc; }
    return ORC_OK;
}
void orc_rays_transform(OrcRay* rays, uint64_t n, const OrcIsometry* iso, int inverse) { /* ray.rs:787-804 */
    const OrcIso3* t = inverse ? &iso->inv : &iso->fwd;
    for (uint64_t i = 0; i < n; ++i) {
        double p[3], d[3];
        iso3_point(t, rays[i].pos, p);
        iso3_vector(t, rays[i].dir, d);
        memcpy(rays[i].pos, p, sizeof p);
        memcpy(rays[i].dir, d, sizeof d);
    }
}

/* kahan 0.1.4 KahanSum::add_assign */
typedef struct { double sum, err; } Kahan;
static inline void kahan_add(Kahan* k, double x) {
    double y = x - k->err;
    double s = k->sum + y;
    k->err = (s - k->sum) - y;
    k->sum = s;
}

double orc_rays_total_energy(const OrcRay* rays, uint64_t n) { /* rays.rs:521-530 */
    Kahan k = { 0.0, 0.0 };
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) kahan_add(&k, rays[i].e);
    return k.sum;
}
uint64_t orc_rays_count(const OrcRay* rays, uint64_t n, int valid_only) { /* :535-540 */
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid || !valid_only) c++;
    return c;
}
int orc_rays_centroid(const OrcRay* rays, uint64_t n, double c[3]) { /* :599-612 */
    double len = (double)orc_rays_count(rays, n, 1);
    if (len == 0.0) return 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) { s0 = s0 + rays[i].pos[0]; s1 = s1 + rays[i].pos[1]; s2 = s2 + rays[i].pos[2]; }
    c[0] = s0 / len; c[1] = s1 / len; c[2] = s2 / len;
    return 1;
}
int orc_rays_energy_weighted_centroid(const OrcRay* rays, uint64_t n, double c[3]) { /* :618-637 */
    if (orc_rays_count(rays, n, 1) == 0) return 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, se = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) {
        double e = rays[i].e;
        s0 = s0 + rays[i].pos[0] * e; s1 = s1 + rays[i].pos[1] * e; s2 = s2 + rays[i].pos[2] * e; se = se + e;
    }
    c[0] = s0 / se; c[1] = s1 / se; c[2] = s2 / se;
    return 1;
}
/* uom: get::<millimeter>() = value * 1e3 ; millimeter!(x) = x * 1e-3 */
int orc_rays_beam_radius_geo(const OrcRay* rays, uint64_t n, double* r_m) { /* :642-660 */
    double c[3];
    if (!orc_rays_centroid(rays, n, c)) return 0;
    double cx = c[0] * 1.0E3, cy = c[1] * 1.0E3;
    double max_dist = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) {
        double x = rays[i].pos[0] * 1.0E3, y = rays[i].pos[1] * 1.0E3;
        double dx = cx - x, dy = cy - y;
        double dist = sqrt(0.0 + (dx * dx + dy * dy));
        if (dist > max_dist) max_dist = dist;
    }
    *r_m = max_dist * 1.0E-3;
    return 1;
}
int orc_rays_beam_radius_rms(const OrcRay* rays, uint64_t n, double* r_m) { /* :666-684 */
    double c[3];
    if (!orc_rays_centroid(rays, n, c)) return 0;
    double cx = c[0] * 1.0E3, cy = c[1] * 1.0E3;
    double sum = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) {
        double x = rays[i].pos[0] * 1.0E3, y = rays[i].pos[1] * 1.0E3;
        double dx = cx - x, dy = cy - y;
        sum += orc_powi(sqrt(0.0 + (dx * dx + dy * dy)), 2);
    }
    sum /= (double)orc_rays_count(rays, n, 1);
    *r_m = sqrt(sum) * 1.0E-3;
    return 1;
}
int orc_rays_energy_weighted_beam_radius_rms(const OrcRay* rays, uint64_t n, double* r_m) { /* :690-701 */
    double c[3];
    if (!orc_rays_energy_weighted_centroid(rays, n, c)) return 0;
    double sum = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) {
        double dist = (c[0] - rays[i].pos[0]) * (c[0] - rays[i].pos[0]) + (c[1] - rays[i].pos[1]) * (c[1] - rays[i].pos[1]);
        sum += dist * rays[i].e;
    }
    sum /= orc_rays_total_energy(rays, n);
    *r_m = sqrt(sum);
    return 1;
}
uint32_t orc_rays_bounce_lvl(const OrcRay* rays, uint64_t n) { /* :183-194 */
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) return rays[i].bounces;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* hit map + Binning fluence                                                  */
/* ------------------------------------------------------------------------- */
int orc_hits_bounding_box(const OrcHit* h, uint64_t n, double lrtb[4]) { /* rays_hit_map.rs:522-562 */
    if (n == 0) return ORC_ERR_EMPTY_HITMAP;
    double left = h[0].x, right = h[0].x, top = h[0].y, bottom = h[0].y;
    for (uint64_t i = 0; i < n; ++i) {
        if (h[i].x < left) left = h[i].x;
        if (h[i].y < bottom) bottom = h[i].y;
        if (h[i].x > right) right = h[i].x;
        if (h[i].y > top) top = h[i].y;
    }
    lrtb[0] = left - 0.0; lrtb[1] = right + 0.0; lrtb[2] = top + 0.0; lrtb[3] = bottom - 0.0;
    return ORC_OK;
}
/* ---- kernel density estimator (kde/mod.rs, kde/gaussian.rs) ---------------------------------- */
static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}
/* Kde::distances_iqr (kde/mod.rs:65-81): despite its name the 75 % quantile of the values */
double orc_kde_distances_iqr(const double* values, uint64_t n) {
    if (n == 0) return NAN;
    double* v = (double*)malloc(n * sizeof(double));
    if (!v) return NAN;
    memcpy(v, values, n * sizeof(double));
    qsort(v, n, sizeof(double), cmp_double);
    double quart = 0.75 * (double)n, ip;
    double out;
    if (modf(quart, &ip) == 0.0) out = 0.5 * (v[(uint64_t)quart - 1] + v[(uint64_t)quart]);
    else out = v[(uint64_t)floor(quart)];
    free(v);
    return out;
}
/* Kde::point_distances_std_dev (kde/mod.rs:44-64): all pair distances (i < j, in that order) and their population std dev */
int orc_kde_point_distances(const OrcHit* h, uint64_t n, double* distances, double* std_dev) {
    uint64_t m = 0;
    double sum = 0.0;
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j) {
            double dx = h[j].x - h[i].x, dy = h[j].y - h[i].y;  /* math_utils::distance_2d_point: ((x2-x1)^2 + (y2-y1)^2).sqrt() */
            double d = sqrt(dx * dx + dy * dy);
            if (distances) distances[m] = d;
            ++m;
            sum += d;
        }
    double nr = (double)m, avg = sum / nr, dev = 0.0;
    m = 0;
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j) {
            double dx = h[j].x - h[i].x, dy = h[j].y - h[i].y;
            double d = distances ? distances[m] : sqrt(dx * dx + dy * dy);
            ++m;
            dev += (d - avg) * (d - avg);
        }
    *std_dev = sqrt(dev / nr);
    return ORC_OK;
}
/* Kde::bandwidth_estimate (kde/mod.rs:83-106): Silverman's rule of thumb on the pair distances */
double orc_kde_bandwidth_estimate(const OrcHit* h, uint64_t n) {
    if (n < 2) return NAN;
    if (n == 2) {
        double dx = h[1].x - h[0].x, dy = h[1].y - h[0].y;
        double bw = 0.5 * sqrt(dx * dx + dy * dy);
        return bw == 0.0 ? NAN : bw;
    }
    uint64_t m = n * (n - 1) / 2;
    double* d = (double*)malloc(m * sizeof(double));
    if (!d) return NAN;
    double sd;
    orc_kde_point_distances(h, n, d, &sd);
    double iqr = orc_kde_distances_iqr(d, m);
    free(d);
    double a = iqr / 1.34;
    double mn = sd < a ? sd : a;  /* uom Length::min == f64::min (NaN-ignoring) */
    if (isnan(sd)) mn = a;
    if (isnan(a)) mn = sd;
    return 0.9 * mn * pow((double)n, -0.2);
}
/* Gaussian2D::value summed over the hit map (kde/gaussian.rs:12-28, kde/mod.rs:107-113), in hit order */
double orc_kde_value(const OrcHit* h, uint64_t n, double band_width, double px, double py) {
    const double FRAC_1_2PI = 0.15915494309189535;
    double sigma_square = band_width * band_width, sum = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        double amplitude = h[i].e * FRAC_1_2PI / sigma_square;
        double dist_square = (px - h[i].x) * (px - h[i].x) + (py - h[i].y) * (py - h[i].y);
        sum += amplitude * exp(-dist_square / (2. * sigma_square));
    }
    return sum;
}
/* RaysHitMap::calc_fluence_with_kde (rays_hit_map.rs:575-613) + Kde::kde_2d (kde/mod.rs:114-134).  ranges: NULL or
 * {r1.start, r1.end, r2.start, r2.end} taken as (left, right, top, bottom) like the reference does.  map: ny rows x nx
 * columns, column-major.  Errors: bandwidth not normal -> ORC_ERR_ARG ("bandwidth must be != 0.0 and finite"). */
int orc_fluence_kde(const OrcHit* h, uint64_t n, uint64_t nx, uint64_t ny, const double* ranges, double* map, double lrtb_out[4],
                    double* band_width_out, int n_threads) {
    double bw = orc_kde_bandwidth_estimate(h, n);
    if (band_width_out) *band_width_out = bw;
    if (!isnormal(bw)) return ORC_ERR_ARG;
    double left, right, top, bottom;
    if (ranges) { left = ranges[0]; right = ranges[1]; top = ranges[2]; bottom = ranges[3]; }
    else {
        double b[4];
        int rc = orc_hits_bounding_box(h, n, b);
        if (rc) return rc;
        double margin = 3. * bw;
        left = b[0] - margin; right = b[1] + margin; bottom = b[3] - margin; top = b[2] + margin;
    }
    double dx = (right - left) / (double)nx, dy = (top - bottom) / (double)ny;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) if (n_threads > 1)
    for (int64_t col = 0; col < (int64_t)nx; ++col)
        for (uint64_t row = 0; row < ny; ++row)
            map[(uint64_t)col * ny + row] = orc_kde_value(h, n, bw, left + (double)col * dx, bottom + (double)row * dy);
    lrtb_out[0] = left; lrtb_out[1] = right; lrtb_out[2] = top; lrtb_out[3] = bottom;
    return ORC_OK;
}

/* Rust `f64 as usize`: saturating, NaN -> 0 */
static inline uint64_t f64_to_usize(double v) {
    if (isnan(v) || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)v;
}
/* rays_hit_map.rs:330-391 */
int orc_fluence_binning(const OrcHit* h, uint64_t n, uint64_t nx, uint64_t ny, const double* ranges,
                        double* map, double lrtb_out[4]) {
    double left, right, top, bottom;
    if (ranges) { left = ranges[0]; right = ranges[1]; top = ranges[2]; bottom = ranges[3]; } /* :337-339 (sic) */
    else {
        double b[4];
        int rc = orc_hits_bounding_box(h, n, b);
        if (rc) return rc;
        left = b[0]; right = b[1]; top = b[2]; bottom = b[3];
    }
    double bin_width = (right - left) / (double)nx;
    double bin_height = (top - bottom) / (double)ny;
    double bin_area = bin_width * bin_height;
    double width_step = (right - left) / (double)(nx - 1);
    double height_step = (top - bottom) / (double)(ny - 1);
    memset(map, 0, (size_t)(nx * ny) * sizeof(double));
    for (uint64_t i = 0; i < n; ++i) {
        double ipx, ipy;
        double fx = modf((h[i].x - left) / width_step, &ipx);
        double fy = modf((h[i].y - bottom) / height_step, &ipy);
        uint64_t xi = f64_to_usize(ipx), yi = f64_to_usize(ipy);
        double fluence = h[i].e / bin_area;
        double f_x = (1.0 - fx) * (1.0 - fy) * fluence;
        double f_nx = fx * (1.0 - fy) * fluence;
        double f_ny = (1.0 - fx) * fy * fluence;
        double f_xy = fx * fy * fluence;
        if (xi >= nx || yi >= ny) return ORC_ERR_ARG; /* index out of bounds: the reference panics */
        map[xi * ny + yi] += f_x;
        if (xi < nx - 1) {
            map[(xi + 1) * ny + yi] += f_nx;
            if (yi < ny - 1) map[(xi + 1) * ny + yi + 1] += f_xy;
        }
        if (yi < ny - 1) map[xi * ny + yi + 1] += f_ny;
    }
    if (lrtb_out) { lrtb_out[0] = left; lrtb_out[1] = right; lrtb_out[2] = top; lrtb_out[3] = bottom; }
    return ORC_OK;
}
double orc_fluence_peak(const double* map, uint64_t n) { /* fluence_data.rs:45-68 */
    double peak = -INFINITY;
    for (uint64_t i = 0; i < n; ++i) if (isfinite(map[i])) peak = fmax(peak, map[i]);
    return peak;
}
/* fluence_data.rs:130-141; x_range = left..right, y_range = bottom..top (rays_hit_map.rs:378-383) */
double orc_fluence_total_energy(const double* map, uint64_t nx, uint64_t ny, const double lrtb[4]) {
    double dx = (lrtb[1] - lrtb[0]) / (double)nx;
    double dy = (lrtb[2] - lrtb[3]) / (double)ny;
    double area = dx * dy;
    double energy = 0.0;
    for (uint64_t i = 0; i < nx * ny; ++i) if (!isnan(map[i])) energy += area * map[i];
    return energy;
}

/* ------------------------------------------------------------------------- */
/* sources                                                                    */
/* ------------------------------------------------------------------------- */
void orc_grid(uint64_t nx_in, uint64_t ny_in, double side_x, double side_y, double* xyz) { /* grid.rs:58-92 */
    uint64_t nx = nx_in < 1 ? 1 : nx_in, ny = ny_in < 1 ? 1 : ny_in;
    double dx = nx > 1 ? side_x / (double)(nx - 1) : 0.0;
    double dy = ny > 1 ? side_y / (double)(ny - 1) : 0.0;
    double ox = nx > 1 ? side_x / 2.0 : 0.0;
    double oy = ny > 1 ? side_y / 2.0 : 0.0;
    uint64_t k = 0;
    for (uint64_t ix = 0; ix < nx; ++ix)
        for (uint64_t iy = 0; iy < ny; ++iy) {
            xyz[3 * k + 0] = (double)ix * dx - ox;
            xyz[3 * k + 1] = (double)iy * dy - oy;
            xyz[3 * k + 2] = 0.0;
            ++k;
        }
}
static inline double fract(double x) { return x - trunc(x); }
void orc_fibonacci_rectangle(uint64_t n, double sx, double sy, double* xyz) { /* fibonacci.rs:53-65 */
    double golden = (1.0 + sqrt(5.0)) / 2.0;
    for (uint64_t i = 0; i < n; ++i) {
        double fi = (double)i;
        xyz[3 * i + 0] = sx * (fract(fi / golden) - 0.5);
        xyz[3 * i + 1] = sy * ((fi / (double)n) - 0.5);
        xyz[3 * i + 2] = 0.0;
    }
}
void orc_fibonacci_ellipse(uint64_t n, double rx, double ry, double* xyz) { /* fibonacci.rs:113-126 */
    double golden = (1.0 + sqrt(5.0)) / 2.0;
    for (uint64_t i = 0; i < n; ++i) {
        double a = 2.0 * M_PI * fract((double)i / golden);
        double sn = sin(a), cs = cos(a);
        double sr = sqrt((double)i / (double)n);
        xyz[3 * i + 0] = rx * sn * sr;
        xyz[3 * i + 1] = ry * cs * sr;
        xyz[3 * i + 2] = 0.0;
    }
}
uint64_t orc_hexapolar_count(uint32_t rings) { return 1 + 3ull * rings * (rings + 1); }
void orc_hexapolar(uint32_t rings, double radius, double* xyz) { /* hexapolar.rs:38-56 */
    uint64_t k = 0;
    xyz[0] = xyz[1] = xyz[2] = 0.0; k = 1;
    if (radius != 0.0) {
        double step = radius / (double)rings;
        for (uint32_t ring = 0; ring < rings; ++ring) {
            double r = (double)(ring + 1) * step;
            uint32_t ppr = 6 * (ring + 1);
            double astep = 2.0 * M_PI / (double)ppr;
            for (uint32_t p = 0; p < ppr; ++p) {
                double a = (double)p * astep;
                xyz[3 * k + 0] = r * sin(a);
                xyz[3 * k + 1] = r * cos(a);
                xyz[3 * k + 2] = 0.0;
                ++k;
            }
        }
    }
}
void orc_energy_uniform(uint64_t n, double total, double* e) { /* uniform.rs:37-41 */
    double per = total / (double)n;
    for (uint64_t i = 0; i < n; ++i) e[i] = per;
}
/* general_gaussian.rs:93-120 + utils/math_distribution_functions.rs:54-70,133-150 */
void orc_energy_general_gaussian(const double* xyz, uint64_t n, double total, double mux, double muy,
                                 double sx, double sy, double power, double theta, int rectangular, double* e) {
    double st = sin(theta), ct = cos(theta);
    Kahan k = { 0.0, 0.0 };
    for (uint64_t i = 0; i < n; ++i) {
        double px = xyz[3 * i], py = xyz[3 * i + 1];
        double x_rot = fma(px - mux, ct, -((py - muy) * st));
        double y_rot = fma(py - muy, ct, (px - mux) * st);
        double g;
        if (rectangular)
            g = exp(-pow(0.5 * orc_powi(x_rot / sx, 2), power) - pow(0.5 * orc_powi(y_rot / sy, 2), power));
        else
            g = exp(-pow(0.5 * fma(x_rot / sx, x_rot / sx, orc_powi(y_rot / sy, 2)), power));
        e[i] = g;
        kahan_add(&k, g);
    }
    for (uint64_t i = 0; i < n; ++i) e[i] = total * e[i] / k.sum;
}
int orc_linspace(double start, double end, uint64_t num, double* out) { /* griddata.rs:211-231 */
    if (!isfinite(start) || !isfinite(end)) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < num; ++i) out[i] = start;
    if (num < 2) return ORC_OK;
    double bin = (end - start) / (double)(num - 1);
    for (uint64_t i = 0; i < num; ++i) out[i] = out[i] + (double)i * bin;
    return ORC_OK;
}
/* spectral_distribution/gaussian.rs:78-96 + math_distribution_functions.rs:168-176 */
void orc_spectrum_gaussian(double lo, double hi, uint64_t num, double mu, double fwhm, double power,
                           double* wvl, double* w) {
    orc_linspace(lo, hi, num, wvl);
    double sigma = fwhm / (2.0 * sqrt(2.0 * pow(log(2.0), 1.0 / power)));
    Kahan k = { 0.0, 0.0 };
    for (uint64_t i = 0; i < num; ++i) {
        w[i] = exp(-pow(0.5 * orc_powi((wvl[i] - mu) / sigma, 2), power));
        kahan_add(&k, w[i]);
    }
    for (uint64_t i = 0; i < num; ++i) w[i] = w[i] / k.sum;
}
/* rays.rs:264-297 */
int orc_rays_collimated_with_spectrum(const double* xyz, const double* e_pos, uint64_t n_pos,
                                      const double* wvl, const double* w, uint64_t n_wvl, OrcRay* out) {
    const double z[3] = { 0.0, 0.0, 1.0 };
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_pos; ++i)
        for (uint64_t j = 0; j < n_wvl; ++j) {
            int rc = orc_ray_new(&xyz[3 * i], z, wvl[j], e_pos[i] * w[j], &out[k]);
            if (rc) return rc;
            out[k].id = (uint32_t)k;
            ++k;
        }
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * spectrum.rs: Spectrum = Vec<(wavelength in micrometers, value in 1/micrometers)>, here two arrays.
 * uom: micrometer!(x) = x * 1e-6, get::<micrometer>() = value * 1e6, get::<nanometer>() = value * (1.0 / 1e-9)
 * ------------------------------------------------------------------------------------------- */
/* Spectrum::new (spectrum.rs:47-73); lambda == NULL: only the number of slots is returned */
int orc_spectrum_new(double start_m, double end_m, double res_m, double* lambda, double* data, uint64_t cap, uint64_t* n_out) {
    if (res_m <= 0.0) return ORC_ERR_SPECTRUM;              /* "resolution must be positive" */
    if (start_m >= end_m) return ORC_ERR_SPECTRUM;          /* "wavelength range must be in ascending order and not empty" */
    if (signbit(start_m) || signbit(end_m)) return ORC_ERR_SPECTRUM;
    double cnt = round((end_m - start_m) / res_m);
    uint64_t n = cnt > 0.0 ? (uint64_t)cnt : 0;             /* f64_to_usize: `as usize` */
    double start = start_m * 1e6, step = res_m * 1e6;
    if (n_out) *n_out = n;
    if (!lambda) return ORC_OK;
    if (n > cap) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) { lambda[i] = fma((double)i, step, start); if (data) data[i] = 0.0; }
    return ORC_OK;
}
/* Spectrum::add_single_peak (spectrum.rs:193-227) */
int orc_spectrum_add_single_peak(const double* lambda, double* data, uint64_t n, double wl_m, double value) {
    if (n == 0) return ORC_ERR_SPECTRUM;
    double r0 = lambda[0] * 1e-6, r1 = lambda[n - 1] * 1e-6;     /* range(): micrometer!(first)..micrometer!(last) */
    if (!(r0 <= wl_m && wl_m < r1)) return ORC_OK;               /* warn!("peak wavelength is not in spectrum range. Spectrum unmodified.") */
    if (value < 0.0) return ORC_ERR_SPECTRUM;                    /* "energy must be positive" */
    double w = wl_m * 1e6;
    if (n < 2) return ORC_ERR_SPECTRUM;                          /* "spectrum size is too small" */
    uint64_t idx = n;
    for (uint64_t i = 0; i < n; ++i) if (lambda[i] >= w) { idx = i; break; }
    if (idx == n) return ORC_ERR_SPECTRUM;                       /* "insertion point not found" */
    if (idx == 0) {
        double delta = lambda[1] - lambda[0];
        data[0] += value / delta;
    } else {
        double lower = lambda[idx - 1], upper = lambda[idx];
        double delta = upper - lower;
        double epm = value / delta;
        double part = epm * (w - lower) / delta;
        data[idx] += part;
        data[idx - 1] += epm - part;
    }
    return ORC_OK;
}
/* Spectrum::from_laser_lines (spectrum.rs:126-148) */
int orc_spectrum_from_laser_lines(const double* wl_m, const double* e, uint64_t n_lines, double res_m, double* lambda, double* data,
                                  uint64_t cap, uint64_t* n_out) {
    if (n_lines == 0) return ORC_ERR_SPECTRUM;                   /* "no laser lines provided" */
    if (res_m <= 0.0) return ORC_ERR_SPECTRUM;
    double mn = wl_m[0], mx = wl_m[0];
    for (uint64_t i = 0; i < n_lines; ++i) { if (wl_m[i] < mn) mn = wl_m[i]; if (wl_m[i] > mx) mx = wl_m[i]; }
    uint64_t n = 0;
    int rc = orc_spectrum_new(mn, mx + 2.0 * res_m, res_m, lambda, data, cap, &n);
    if (n_out) *n_out = n;
    if (rc || !lambda) return rc;
    for (uint64_t i = 0; i < n_lines; ++i) {
        rc = orc_spectrum_add_single_peak(lambda, data, n, wl_m[i], e[i]);
        if (rc) return rc;
    }
    return ORC_OK;
}
/* Rays::to_spectrum (rays.rs:1209-1223): the valid rays as laser lines */
int orc_rays_to_spectrum(const OrcRay* rays, uint64_t n, double res_m, double* lambda, double* data, uint64_t cap, uint64_t* n_out) {
    uint64_t nv = orc_rays_count(rays, n, 1);
    if (nv == 0) return ORC_ERR_SPECTRUM;                        /* "ray bundle is empty - cannot create spectrum" */
    double* wl = (double*)malloc(nv * sizeof(double));
    double* e = (double*)malloc(nv * sizeof(double));
    if (!wl || !e) { free(wl); free(e); return ORC_ERR_ARG; }
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) if (rays[i].valid) { wl[k] = rays[i].wvl; e[k] = rays[i].e; ++k; }
    int rc = orc_spectrum_from_laser_lines(wl, e, nv, res_m, lambda, data, cap, n_out);
    free(wl); free(e);
    return rc;
}
/* Spectrum::total_energy (spectrum.rs:289-298) */
double orc_spectrum_total_energy(const double* lambda, const double* data, uint64_t n) {
    Kahan k = { 0.0, 0.0 };
    for (uint64_t i = 0; i + 1 < n; ++i) kahan_add(&k, (lambda[i + 1] - lambda[i]) * data[i]);
    return k.sum;
}
/* Spectrum::center_wavelength (spectrum.rs:304-316); returns metres */
double orc_spectrum_center_wavelength(const double* lambda, const double* data, uint64_t n) {
    double weighted_sum = 0.0, total_weight = 0.0;
    for (uint64_t i = 0; i + 1 < n; ++i) {
        double bin_width = lambda[i + 1] - lambda[i];
        double bin_weight = data[i] * bin_width;
        weighted_sum += lambda[i] * bin_weight;
        total_weight += bin_weight;
    }
    return (weighted_sum / total_weight) * 1e-6;
}
/* Spectrum::get_value (spectrum.rs:321-349): returns 1 = Some(*out), 0 = None */
int orc_spectrum_get_value(const double* lambda, const double* data, uint64_t n, double wl_m, double* out) {
    if (n == 0) return 0;
    double w = wl_m * 1e6;
    if (w == lambda[n - 1]) { *out = data[n - 1]; return 1; }
    double r0 = lambda[0] * 1e-6, r1 = lambda[n - 1] * 1e-6;
    if (!(r0 <= wl_m && wl_m < r1)) return 0;
    uint64_t idx = n;
    for (uint64_t i = 0; i < n; ++i) if (lambda[i] >= w) { idx = i; break; }
    if (idx == n) return 0;
    uint64_t l = idx == 0 ? 0 : idx - 1, r = idx == 0 ? 1 : idx;
    double ratio = (w - lambda[l]) / (lambda[r] - lambda[l]);
    *out = fma(data[l], 1.0 - ratio, data[r] * ratio);
    return 1;
}
/* Ray::filter_energy with FilterType::Spectrum (ray.rs:712-721) over the valid rays (rays.rs:1119-1133) */
int orc_rays_filter_energy_spectrum(OrcRay* rays, uint64_t n, const double* lambda, const double* data, uint64_t ns) {
    for (uint64_t i = 0; i < n; ++i) {
        if (!rays[i].valid) continue;
        double t;
        if (!orc_spectrum_get_value(lambda, data, ns, rays[i].wvl, &t)) return ORC_ERR_SPECTRUM; /* "wavelength of ray outside filter spectrum" */
        rays[i].e *= t;
    }
    return ORC_OK;
}
/* Ray::split with SplittingConfig::Spectrum (ray.rs:752-772) over the bundle (rays.rs:1253-1264: invalid rays are copied unchanged) */
int orc_rays_split_spectrum(OrcRay* rays, uint64_t n, const double* lambda, const double* data, uint64_t ns, OrcRay* split_out) {
    for (uint64_t i = 0; i < n; ++i) {
        split_out[i] = rays[i];
        if (!rays[i].valid) continue;
        double ratio;
        if (!orc_spectrum_get_value(lambda, data, ns, rays[i].wvl, &ratio)) return ORC_ERR_SPECTRUM; /* "ray splitting failed. wavelength outside given spectrum" */
        if (!(ratio >= 0.0 && ratio <= 1.0)) return ORC_ERR_SPLIT_RATIO;
        rays[i].e *= ratio;
        split_out[i].e *= 1.0 - ratio;
    }
    return ORC_OK;
}

/* Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802): rows (x mm, y mm, wavefront error in wavelengths) of the
 * valid rays in order, relative to the ray closest to the monitor axis (the first one on ties) */
int orc_rays_wavefront_error(const OrcRay* rays, uint64_t n, double wavelength_m, const OrcIsometry* monitor_iso, double* xyw,
                             uint64_t* n_out) {
    double wvl = wavelength_m * (1.0 / 1e-9);
    double min_radius = INFINITY, center = 0.0;
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (!rays[i].valid) continue;
        double p[3];
        iso3_point(&monitor_iso->inv, rays[i].pos, p);
        double x = p[0] * 1e3, y = p[1] * 1e3;
        xyw[3 * k] = x; xyw[3 * k + 1] = y;
        xyw[3 * k + 2] = -(rays[i].opl * (1.0 / 1e-9));
        double radius = fma(x, x, y * y);
        if (radius < min_radius) { min_radius = radius; center = xyw[3 * k + 2]; }
        ++k;
    }
    for (uint64_t j = 0; j < k; ++j) { xyw[3 * j + 2] -= center; xyw[3 * j + 2] /= wvl; }
    if (n_out) *n_out = k;
    return ORC_OK;
}
/* WaveFrontErrorMap::calc_wavefront_statistics (nodes/wavefront.rs:145-164): (ptv, rms) of column 3 (stride 3) */
int orc_wavefront_statistics(const double* xyw, uint64_t n, double* ptv, double* rms) {
    if (n == 0) return ORC_ERR_EMPTY_WAVEFRONT;                  /* "Empty wavefront-data vector!" */
    double mx = xyw[2], mn = xyw[2], sum = 0.0;
    for (uint64_t i = 0; i < n; ++i) { double v = xyw[3 * i + 2]; if (v > mx) mx = v; if (v < mn) mn = v; sum += v; }
    double avg = sum / (double)n;
    double s2 = 0.0;
    for (uint64_t i = 0; i < n; ++i) { double d = xyw[3 * i + 2] - avg; s2 += d * d; }
    *ptv = mx - mn;
    *rms = sqrt(s2 / (double)(int32_t)n);
    return ORC_OK;
}
