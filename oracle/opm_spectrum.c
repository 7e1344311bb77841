/* opm_spectrum.c -- CPU oracle, TEST INFRASTRUCTURE ONLY (see opm_oracle.h): the host-side operations on a Spectrum that the ray
 * path's callers use beside the ones in opm_oracle.c (opossum/src/spectrum.rs).  A Spectrum is Vec<(wavelength in micrometers,
 * value in 1/micrometer)>, here two arrays.  Every function follows the reference's statements in their order; the line
 * numbers are those of opossum/src/spectrum.rs.  Pinned by the reference's own tests (spectrum.rs:677-1030), transcribed in
 * tests/test_oracle_kats.py.
 *
 * Third-party code behind from_csv: the `csv` crate (Cargo.lock: csv 1.3.1 / csv-core), not under /root/reference.  Restated
 * here for the subset the reference configures (no headers, delimiter ';', default terminator CRLF / LF / CR, default
 * `flexible(false)`: every record must have as many fields as the first one -- the reference's spec_to_csv_test_03.csv fails
 * on exactly that; empty lines are skipped; a trailing line without a terminator is a record; quoting is not restated). */
#include "opm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Spectrum::add_lorentzian_peak (spectrum.rs:249-283); lorentz() :608-610 */
static double lorentz(double center, double width, double x) {
    return 0.5 / M_PI * width / fma(0.25 * width, width, (x - center) * (x - center));
}
int orc_spectrum_add_lorentzian_peak(const double* lambda, double* data, uint64_t n, double center_m, double width_m, double energy) {
    if (signbit(center_m)) return ORC_ERR_SPECTRUM; /* "center wavelength must be positive" */
    if (signbit(width_m)) return ORC_ERR_SPECTRUM;  /* "line width must be positive" */
    if (energy < 0.0) return ORC_ERR_SPECTRUM;      /* "energy must be positive" */
    double c = center_m * 1e6, w = width_m * 1e6;
    for (uint64_t i = 0; i < n; ++i) data[i] = fma(energy, lorentz(c, w, lambda[i]), data[i]);
    return ORC_OK;
}
/* Spectrum::scale_vertical (spectrum.rs:355-367) */
int orc_spectrum_scale_vertical(double* data, uint64_t n, double factor) {
    if (factor < 0.0) return ORC_ERR_SPECTRUM; /* "scaling factor mus be >= 0.0" */
    for (uint64_t i = 0; i < n; ++i) data[i] = data[i] * factor;
    return ORC_OK;
}
/* calc_ratio (spectrum.rs:588-606) */
double orc_spectrum_calc_ratio(double bucket_left, double bucket_right, double source_left, double source_right) {
    if (bucket_left < source_left && bucket_right > source_left && bucket_right < source_right)
        return (bucket_right - source_left) / (source_right - source_left);
    if (bucket_left <= source_left && bucket_right >= source_right) return 1.0;
    if (bucket_left >= source_left && bucket_right <= source_right) return (bucket_right - bucket_left) / (source_right - source_left);
    if (bucket_left > source_left && bucket_left < source_right && bucket_right > source_right)
        return (source_right - bucket_left) / (source_right - source_left);
    return 0.0;
}
/* Spectrum::resample (spectrum.rs:377-428): self = (lambda, data) takes the values of src = (sl, sd).  `windows(2)` iterators as
 * indices: window k of a list of m entries exists while k + 1 < m. */
void orc_spectrum_resample(const double* lambda, double* data, uint64_t n, const double* sl, const double* sd, uint64_t ns) {
    if (ns < 2) return;                                 /* src_it.next() is None */
    uint64_t src_next = 1;                              /* the next window src_it would yield */
    double src_lower = sl[0], src_upper = sl[1];
    uint64_t src_idx = 0;
    if (n < 2) return;                                  /* bucket_it.next() is None */
    uint64_t bucket_next = 1;
    double bucket_lower = lambda[0], bucket_upper = lambda[1];
    uint64_t bucket_idx = 0;
    data[bucket_idx] = 0.0;
    while (src_upper < bucket_lower) {
        if (src_next + 1 < ns) {
            src_lower = sl[src_next]; src_upper = sl[src_next + 1]; ++src_next;
            src_idx += 1;
        } else break;
    }
    for (;;) {
        double ratio = orc_spectrum_calc_ratio(bucket_lower, bucket_upper, src_lower, src_upper);
        double bucket_value = sd[src_idx] * ratio * (src_upper - src_lower) / (bucket_upper - bucket_lower);
        data[bucket_idx] += bucket_value;
        if (src_upper < bucket_upper) {
            if (src_next + 1 < ns) {
                src_lower = sl[src_next]; src_upper = sl[src_next + 1]; ++src_next;
                src_idx += 1;
                continue;
            }
            break;
        } else if (bucket_next + 1 < n) {
            bucket_lower = lambda[bucket_next]; bucket_upper = lambda[bucket_next + 1]; ++bucket_next;
            bucket_idx += 1;
            data[bucket_idx] = 0.0;
            continue;
        }
        break;
    }
}
/* `let mut resampled_spec = self.clone(); resampled_spec.resample(other);` -- the values self would get; caller frees */
static double* resampled_like_self(const double* lambda, const double* data, uint64_t n, const double* sl, const double* sd, uint64_t ns) {
    double* r = (double*)malloc((n ? n : 1) * sizeof(double));
    if (!r) return NULL;
    memcpy(r, data, n * sizeof(double));
    orc_spectrum_resample(lambda, r, n, sl, sd, ns);
    return r;
}
/* Spectrum::filter (spectrum.rs:430-439) */
int orc_spectrum_filter(const double* lambda, double* data, uint64_t n, const double* fl, const double* fd, uint64_t nf) {
    double* r = resampled_like_self(lambda, data, n, fl, fd, nf);
    if (!r) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) data[i] = data[i] * r[i];
    free(r);
    return ORC_OK;
}
/* Spectrum::split_by_spectrum (spectrum.rs:456-472): data keeps value * t, split_out gets value * (1 - t) */
int orc_spectrum_split_by_spectrum(const double* lambda, double* data, uint64_t n, const double* fl, const double* fd, uint64_t nf,
                                   double* split_out) {
    double* r = resampled_like_self(lambda, data, n, fl, fd, nf);
    if (!r) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) {
        double v = data[i];
        data[i] = v * r[i];
        split_out[i] = v * (1.0 - r[i]);
    }
    free(r);
    return ORC_OK;
}
/* Spectrum::add (spectrum.rs:483-492) / Spectrum::sub (:497-507; f64::clamp(0, |d|): NaN stays NaN) */
int orc_spectrum_add(const double* lambda, double* data, uint64_t n, const double* ol, const double* od, uint64_t no) {
    double* r = resampled_like_self(lambda, data, n, ol, od, no);
    if (!r) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) data[i] = data[i] + r[i];
    free(r);
    return ORC_OK;
}
int orc_spectrum_sub(const double* lambda, double* data, uint64_t n, const double* ol, const double* od, uint64_t no) {
    double* r = resampled_like_self(lambda, data, n, ol, od, no);
    if (!r) return ORC_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) {
        double d = data[i] - r[i], hi = fabs(d);
        data[i] = d < 0.0 ? 0.0 : (d > hi ? hi : d);
    }
    free(r);
    return ORC_OK;
}
/* Spectrum::average_resolution (spectrum.rs:172-176), metres */
double orc_spectrum_average_resolution(const double* lambda, uint64_t n) {
    double bandwidth = lambda[n - 1] * 1e-6 - lambda[0] * 1e-6;
    return bandwidth / ((double)n - 1.0);
}
/* merge_spectra (spectrum.rs:622-650) for two present spectra; lambda == NULL: only the number of slots */
int orc_spectrum_merge(const double* l1, const double* d1, uint64_t n1, const double* l2, const double* d2, uint64_t n2, double* lambda,
                       double* data, uint64_t cap, uint64_t* n_out) {
    if (n1 == 0 || n2 == 0) return ORC_ERR_ARG;
    double s1 = l1[0] * 1e-6, e1 = l1[n1 - 1] * 1e-6, s2 = l2[0] * 1e-6, e2 = l2[n2 - 1] * 1e-6;
    double minimum = fmin(s1, s2), maximum = fmax(e1, e2);
    double resolution = fmin(orc_spectrum_average_resolution(l1, n1), orc_spectrum_average_resolution(l2, n2));
    uint64_t n = 0;
    int rc = orc_spectrum_new(minimum, maximum, resolution, lambda, data, cap, &n);
    if (n_out) *n_out = n;
    if (rc || !lambda) return rc; /* the reference unwraps: an invalid range panics */
    orc_spectrum_resample(lambda, data, n, l1, d1, n1);
    return orc_spectrum_add(lambda, data, n, l2, d2, n2);
}

/* Spectrum::from_csv (spectrum.rs:89-117): two columns, ';', wavelength in nm, value in percent.  lambda == NULL: count only. */
static int parse_f64(const char* s, size_t len, double* out) { /* Rust's str::parse::<f64>: the whole field, no blanks */
    if (len == 0 || len > 255) return 0;
    char buf[256];
    memcpy(buf, s, len);
    buf[len] = 0;
    for (size_t i = 0; i < len; ++i) {
        char c = buf[i];
        int ok = (c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E';
        if (!ok) { /* "inf", "infinity", "nan" in any case are accepted by Rust as well */
            const char* t = buf + ((buf[0] == '+' || buf[0] == '-') ? 1 : 0);
            if (strcasecmp(t, "inf") && strcasecmp(t, "infinity") && strcasecmp(t, "nan")) return 0;
            break;
        }
    }
    char* end = NULL;
    double v = strtod(buf, &end);
    if (end != buf + len) return 0;
    *out = v;
    return 1;
}
int orc_spectrum_from_csv(const char* path, double* lambda, double* data, uint64_t cap, uint64_t* n_out) {
    FILE* f = fopen(path, "rb");
    if (!f) return ORC_ERR_SPECTRUM; /* File::open error */
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* txt = (char*)malloc((size_t)sz + 1);
    if (!txt) { fclose(f); return ORC_ERR_ARG; }
    size_t got = fread(txt, 1, (size_t)sz, f);
    fclose(f);
    txt[got] = 0;
    uint64_t n = 0;
    int rc = ORC_OK;
    long first_fields = -1;
    size_t pos = 0;
    while (pos < got && rc == ORC_OK) {
        size_t eol = pos;
        while (eol < got && txt[eol] != '\n' && txt[eol] != '\r') ++eol;
        size_t len = eol - pos;
        if (len > 0) { /* one record; csv skips empty lines */
            long fields = 1;
            size_t sep = len;
            for (size_t k = 0; k < len; ++k)
                if (txt[pos + k] == ';') { if (sep == len) sep = k; ++fields; }
            if (first_fields < 0) first_fields = fields;
            double l = 0.0, d = 0.0;
            if (fields != first_fields) rc = ORC_ERR_SPECTRUM;                         /* UnequalLengths */
            else if (!parse_f64(txt + pos, sep, &l)) rc = ORC_ERR_SPECTRUM;            /* record.get(0).parse */
            else if (fields < 2) rc = ORC_ERR_SPECTRUM;                                /* record.get(1).unwrap() panics in the reference */
            else {
                size_t s2 = sep + 1, e2 = s2;
                while (e2 < len && txt[pos + e2] != ';') ++e2;
                if (!parse_f64(txt + pos + s2, e2 - s2, &d)) rc = ORC_ERR_SPECTRUM;
            }
            if (rc == ORC_OK) {
                if (lambda && n < cap) { lambda[n] = l * 1.0E-3; data[n] = d * 0.01; }
                ++n;
            }
        }
        pos = eol;
        if (pos < got && txt[pos] == '\r') ++pos;
        if (pos < got && txt[pos] == '\n') ++pos;
    }
    free(txt);
    if (rc) return rc;
    if (n == 0) return ORC_ERR_SPECTRUM; /* "no csv data was found in file" */
    if (n_out) *n_out = n;
    if (lambda && n > cap) return ORC_ERR_ARG;
    return ORC_OK;
}

/* spectrum_helper.rs:124-190 generate_filter_spectrum: kind 0 ShortPassStep, 1 ShortPassSmooth, 2 LongPassStep, 3 LongPassSmooth.
 * lambda == NULL: only the number of slots.  (The smooth forms reach 0 / 1 at cut_off + width, not + width / 2: the reference's
 * code, :156 and :182, not its doc comment.) */
int orc_spectrum_generate_filter(int kind, double start_m, double end_m, double res_m, double cut_off_m, double width_m, double* lambda,
                                 double* data, uint64_t cap, uint64_t* n_out) {
    uint64_t n = 0;
    int rc = orc_spectrum_new(start_m, end_m, res_m, lambda, data, cap, &n);
    if (n_out) *n_out = n;
    if (rc) return rc;
    if (kind < 0 || kind > 3) return ORC_ERR_ARG;
    if (kind == 1 || kind == 3) { /* width.is_zero() || !width.is_normal() || width.is_sign_negative() */
        if (width_m == 0.0 || !isnormal(width_m) || signbit(width_m)) return ORC_ERR_SPECTRUM;
    }
    if (!lambda) return ORC_OK;
    double cut = cut_off_m * 1e6, w = width_m * 1e6;
    for (uint64_t i = 0; i < n; ++i) {
        double l = lambda[i];
        switch (kind) {
            case 0: data[i] = l < cut ? 1.0 : 0.0; break;
            case 2: data[i] = l > cut ? 1.0 : 0.0; break;
            default: {
                double lo = kind == 1 ? 1.0 : 0.0, hi = kind == 1 ? 0.0 : 1.0;
                if (l <= cut - w / 2.0) data[i] = lo;
                else if (l >= cut + w) data[i] = hi;
                else {
                    double angle = M_PI / w * (l - cut);
                    data[i] = kind == 1 ? fma(0.5, -sin(angle), 0.5) : fma(0.5, sin(angle), 0.5);
                }
            }
        }
    }
    return ORC_OK;
}
