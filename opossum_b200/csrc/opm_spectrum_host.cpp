// opm_spectrum_host.cpp -- the host-side operations on a Spectrum (opossum/src/spectrum.rs) that the callers of the ray path use
// around it: building a spectrum (slots, laser lines, Lorentzian lines, a transmission curve from a CSV file), combining two of them
// (resample, filter, split, add, sub, merge).  A Spectrum travels over the C ABI as two arrays -- ascending wavelength slots in
// micrometers and values in 1/um (spectrum.rs:35-37) -- which is also what opm_spectrum_create() uploads for the spectrum-driven
// filter / splitter ops and what opm_rays_to_spectrum() returns.  O(slots) host arithmetic, no device work.
//
// The reference's statements are followed in their order (the values are compared bit for bit with the oracle's restatement);
// line numbers are those of opossum/src/spectrum.rs.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/opossum_b200.h"

namespace opm {
OpmStatus host_fail(OpmStatus s, const char* msg);  // opm_api.cu: sets opm_last_error()
}
using opm::host_fail;

namespace {

// uom 0.37: Length::new::<micrometer>(x) = x * 1e-6 and get::<micrometer>() = value * (1 / 1e-6) = value * 1e6 -- a multiplication
// by the reciprocal, not a division: the reference's own test pins it (spectrum.rs:763-766: nanometer!(800.0).get::<micrometer>()
// == 0.8 exactly; 800e-9 / 1e-6 is 0.8000000000000002)
constexpr double kMicro = 1e-6, kPerMicro = 1.0 / 1e-6;

// `slice.windows(2)` over the wavelength slots: yields (lower, upper) of consecutive slots
class SlotWindows {
public:
    SlotWindows(const double* slots, uint64_t n) : s_(slots), n_(n) {}
    bool next() {
        if (k_ + 1 >= n_) return false;
        lower = s_[k_]; upper = s_[k_ + 1];
        ++k_;
        return true;
    }
    double lower = 0.0, upper = 0.0;

private:
    const double* s_;
    uint64_t n_, k_ = 0;
};

// calc_ratio (:588-606): the share of the source interval that falls into the bucket
double overlap_ratio(double b_lo, double b_hi, double s_lo, double s_hi) {
    const double s_len = s_hi - s_lo;
    if (b_lo < s_lo && b_hi > s_lo && b_hi < s_hi) return (b_hi - s_lo) / s_len;  // bucket is left partly outside source
    if (b_lo <= s_lo && b_hi >= s_hi) return 1.0;                                   // bucket contains source
    if (b_lo >= s_lo && b_hi <= s_hi) return (b_hi - b_lo) / s_len;                // bucket is part of source
    if (b_lo > s_lo && b_lo < s_hi && b_hi > s_hi) return (s_hi - b_lo) / s_len;    // bucket is right partly outside source
    return 0.0;
}

// Spectrum::resample (:377-428): the values of `src` mapped onto the slots of `dst`
void resample_into(const double* dst_slots, double* dst, uint64_t n, const double* src_slots, const double* src, uint64_t ns) {
    SlotWindows sw(src_slots, ns), bw(dst_slots, n);
    if (!sw.next()) return;
    if (!bw.next()) return;
    uint64_t si = 0, bi = 0;
    dst[bi] = 0.0;
    while (sw.upper < bw.lower) {  // source intervals left of the first bucket
        if (!sw.next()) break;
        ++si;
    }
    for (;;) {
        const double ratio = overlap_ratio(bw.lower, bw.upper, sw.lower, sw.upper);
        dst[bi] += src[si] * ratio * (sw.upper - sw.lower) / (bw.upper - bw.lower);
        if (sw.upper < bw.upper) {
            if (!sw.next()) break;
            ++si;
        } else {
            if (!bw.next()) break;
            dst[++bi] = 0.0;
        }
    }
}

// `self.clone().resample(other)`: the other spectrum's values on self's slots (slots of self the walk never reaches keep self's value)
std::vector<double> other_on_my_slots(const double* slots, const double* values, uint64_t n, const double* o_slots, const double* o_values,
                                      uint64_t no) {
    std::vector<double> r(values, values + n);
    resample_into(slots, r.data(), n, o_slots, o_values, no);
    return r;
}

bool bad(const void* a, const void* b) { return !a || !b; }

// Spectrum::new's checks and slot count (:47-63)
OpmStatus slot_count(double start_m, double end_m, double res_m, uint64_t* n) {
    if (res_m <= 0.0 || std::isnan(res_m)) return host_fail(OPM_ERR_SPECTRUM, "resolution must be positive");
    if (!(start_m < end_m)) return host_fail(OPM_ERR_SPECTRUM, "wavelength range must be in ascending order and not empty");
    if (std::signbit(start_m) || std::signbit(end_m)) return host_fail(OPM_ERR_SPECTRUM, "wavelength range limits must both be positive");
    const double cnt = std::round((end_m - start_m) / res_m);
    *n = cnt > 0.0 ? (cnt >= 18446744073709551615.0 ? UINT64_MAX : (uint64_t)cnt) : 0;  // `as usize` saturates
    return OPM_OK;
}

}  // namespace

// Spectrum::new (:47-73).  lambda_um == NULL: only *n_out
extern "C" OpmStatus opm_spectrum_new(double start_m, double end_m, double resolution_m, double* lambda_um, double* values, uint64_t cap,
                                      uint64_t* n_out) {
    uint64_t n = 0;
    if (OpmStatus st = slot_count(start_m, end_m, resolution_m, &n)) return st;
    if (n_out) *n_out = n;
    if (!lambda_um) return OPM_OK;
    if (!values || n > cap) return host_fail(OPM_ERR_ARG, "spectrum arrays too small");
    const double start = start_m * kPerMicro, step = resolution_m * kPerMicro;
    for (uint64_t i = 0; i < n; ++i) { lambda_um[i] = std::fma((double)i, step, start); values[i] = 0.0; }
    return OPM_OK;
}

// Spectrum::add_single_peak (:193-227): the energy of a resolution-limited line shared between its two neighbouring slots
extern "C" OpmStatus opm_spectrum_add_single_peak(const double* lambda_um, double* values, uint64_t n, double wavelength_m, double value) {
    if (bad(lambda_um, values) || n == 0) return host_fail(OPM_ERR_ARG, "NULL or empty spectrum");
    if (!(lambda_um[0] * kMicro <= wavelength_m && wavelength_m < lambda_um[n - 1] * kMicro)) return OPM_OK;  // warning, spectrum unmodified
    if (value < 0.0) return host_fail(OPM_ERR_SPECTRUM, "energy must be positive");
    if (n < 2) return host_fail(OPM_ERR_SPECTRUM, "spectrum size is too small");
    const double w = wavelength_m * kPerMicro;
    const double* at = std::find_if(lambda_um, lambda_um + n, [w](double l) { return l >= w; });
    if (at == lambda_um + n) return host_fail(OPM_ERR_SPECTRUM, "insertion point not found");
    const uint64_t idx = (uint64_t)(at - lambda_um);
    if (idx == 0) {
        values[0] += value / (lambda_um[1] - lambda_um[0]);
        return OPM_OK;
    }
    const double lower = lambda_um[idx - 1], delta = lambda_um[idx] - lower;
    const double per_um = value / delta;
    const double upper_part = per_um * (w - lower) / delta;
    values[idx] += upper_part;
    values[idx - 1] += per_um - upper_part;
    return OPM_OK;
}

// Spectrum::from_laser_lines (:126-148)
extern "C" OpmStatus opm_spectrum_from_laser_lines(const double* wavelengths_m, const double* energies, uint64_t n_lines, double resolution_m,
                                                   double* lambda_um, double* values, uint64_t cap, uint64_t* n_out) {
    if (n_lines == 0) return host_fail(OPM_ERR_SPECTRUM, "no laser lines provided");
    if (bad(wavelengths_m, energies)) return host_fail(OPM_ERR_ARG, "NULL argument");
    if (resolution_m <= 0.0 || std::isnan(resolution_m)) return host_fail(OPM_ERR_SPECTRUM, "resolution must be positive");
    const auto mm = std::minmax_element(wavelengths_m, wavelengths_m + n_lines);
    uint64_t n = 0;
    OpmStatus st = opm_spectrum_new(*mm.first, *mm.second + 2.0 * resolution_m, resolution_m, lambda_um, values, cap, &n);
    if (n_out) *n_out = n;
    if (st || !lambda_um) return st;
    for (uint64_t i = 0; i < n_lines && !st; ++i) st = opm_spectrum_add_single_peak(lambda_um, values, n, wavelengths_m[i], energies[i]);
    return st;
}

// Spectrum::add_lorentzian_peak (:249-283), lorentz (:608-610)
extern "C" OpmStatus opm_spectrum_add_lorentzian_peak(const double* lambda_um, double* values, uint64_t n, double center_m, double width_m,
                                                      double energy) {
    if (n && bad(lambda_um, values)) return host_fail(OPM_ERR_ARG, "NULL argument");
    if (std::signbit(center_m)) return host_fail(OPM_ERR_SPECTRUM, "center wavelength must be positive");
    if (std::signbit(width_m)) return host_fail(OPM_ERR_SPECTRUM, "line width must be positive");
    if (energy < 0.0) return host_fail(OPM_ERR_SPECTRUM, "energy must be positive");
    const double c = center_m * kPerMicro, w = width_m * kPerMicro;
    const double pi = 3.14159265358979323846264338327950288;
    for (uint64_t i = 0; i < n; ++i) {
        const double dx = lambda_um[i] - c;
        const double line = 0.5 / pi * w / std::fma(0.25 * w, w, dx * dx);
        values[i] = std::fma(energy, line, values[i]);
    }
    return OPM_OK;
}

// Spectrum::scale_vertical (:355-367)
extern "C" OpmStatus opm_spectrum_scale_vertical(double* values, uint64_t n, double factor) {
    if (n && !values) return host_fail(OPM_ERR_ARG, "NULL argument");
    if (factor < 0.0) return host_fail(OPM_ERR_SPECTRUM, "scaling factor mus be >= 0.0");
    std::transform(values, values + n, values, [factor](double v) { return v * factor; });
    return OPM_OK;
}

// Spectrum::resample (:377-428): (lambda_um, values) takes the values of the source spectrum; energy-conserving where the source
// lies inside the slots
extern "C" OpmStatus opm_spectrum_resample(const double* lambda_um, double* values, uint64_t n, const double* src_lambda_um,
                                           const double* src_values, uint64_t n_src) {
    if ((n && bad(lambda_um, values)) || (n_src && bad(src_lambda_um, src_values))) return host_fail(OPM_ERR_ARG, "NULL argument");
    resample_into(lambda_um, values, n, src_lambda_um, src_values, n_src);
    return OPM_OK;
}

// Spectrum::filter (:430-439), ::add (:483-492), ::sub (:497-507, negative differences clamped to 0), ::split_by_spectrum (:456-472)
extern "C" OpmStatus opm_spectrum_combine(int32_t what, const double* lambda_um, double* values, uint64_t n, const double* other_lambda_um,
                                          const double* other_values, uint64_t n_other, double* split_out) {
    if ((n && bad(lambda_um, values)) || (n_other && bad(other_lambda_um, other_values))) return host_fail(OPM_ERR_ARG, "NULL argument");
    if (what < OPM_SPECTRUM_FILTER || what > OPM_SPECTRUM_SPLIT || (what == OPM_SPECTRUM_SPLIT && n && !split_out))
        return host_fail(OPM_ERR_ARG, "unknown spectrum operation / no array for the split-off spectrum");
    const std::vector<double> o = other_on_my_slots(lambda_um, values, n, other_lambda_um, other_values, n_other);
    for (uint64_t i = 0; i < n; ++i) {
        const double v = values[i];
        switch (what) {
            case OPM_SPECTRUM_FILTER: values[i] = v * o[i]; break;
            case OPM_SPECTRUM_ADD: values[i] = v + o[i]; break;
            case OPM_SPECTRUM_SUB: {
                const double d = v - o[i];
                values[i] = d < 0.0 ? 0.0 : d;  // clamp(0, |d|); NaN stays NaN
                break;
            }
            default: values[i] = v * o[i]; split_out[i] = v * (1.0 - o[i]); break;
        }
    }
    return OPM_OK;
}

// Spectrum::average_resolution (:172-176), metres
extern "C" double opm_spectrum_average_resolution(const double* lambda_um, uint64_t n) {
    if (!lambda_um || n == 0) return std::nan("");
    return (lambda_um[n - 1] * kMicro - lambda_um[0] * kMicro) / ((double)n - 1.0);
}

// merge_spectra (:622-650) of two present spectra: slots spanning both at the finer average resolution, s1 resampled, s2 added
extern "C" OpmStatus opm_spectrum_merge(const double* l1, const double* v1, uint64_t n1, const double* l2, const double* v2, uint64_t n2,
                                        double* lambda_um, double* values, uint64_t cap, uint64_t* n_out) {
    if (bad(l1, v1) || bad(l2, v2) || n1 == 0 || n2 == 0) return host_fail(OPM_ERR_ARG, "NULL or empty spectrum");
    const double lo = std::fmin(l1[0] * kMicro, l2[0] * kMicro), hi = std::fmax(l1[n1 - 1] * kMicro, l2[n2 - 1] * kMicro);
    const double res = std::fmin(opm_spectrum_average_resolution(l1, n1), opm_spectrum_average_resolution(l2, n2));
    uint64_t n = 0;
    OpmStatus st = opm_spectrum_new(lo, hi, res, lambda_um, values, cap, &n);  // the reference unwraps (panics) where this fails
    if (n_out) *n_out = n;
    if (st || !lambda_um) return st;
    resample_into(lambda_um, values, n, l1, v1, n1);
    return opm_spectrum_combine(OPM_SPECTRUM_ADD, lambda_um, values, n, l2, v2, n2, nullptr);
}

// generate_filter_spectrum (spectrum_helper.rs:124-190): the transmission curve of an ideal short- / long-pass filter on the slots of
// Spectrum::new(range, resolution) -- what a caller hands to opm_spectrum_create for an IdealFilter with FilterType::Spectrum.  A cut-off
// outside the range is a warning in the reference, not an error.  The smooth forms follow the reference's code: full transmission /
// absorption up to cut_off - width / 2, the sine in between, and the other end at cut_off + width (:156, :182).
extern "C" OpmStatus opm_spectrum_generate_filter(int32_t kind, double start_m, double end_m, double resolution_m, double cut_off_m,
                                                  double width_m, double* lambda_um, double* values, uint64_t cap, uint64_t* n_out) {
    uint64_t n = 0;
    if (OpmStatus st = opm_spectrum_new(start_m, end_m, resolution_m, lambda_um, values, cap, &n)) return st;
    if (n_out) *n_out = n;
    if (kind < OPM_FILTER_SHORT_PASS_STEP || kind > OPM_FILTER_LONG_PASS_SMOOTH) return host_fail(OPM_ERR_ARG, "unknown filter type");
    const bool smooth = kind == OPM_FILTER_SHORT_PASS_SMOOTH || kind == OPM_FILTER_LONG_PASS_SMOOTH;
    if (smooth && (width_m == 0.0 || !std::isnormal(width_m) || std::signbit(width_m)))
        return host_fail(OPM_ERR_SPECTRUM, "width must be positive and finite");
    if (!lambda_um) return OPM_OK;
    const double cut = cut_off_m * kPerMicro, width = width_m * kPerMicro;
    const double pi = 3.14159265358979323846264338327950288;
    const bool short_pass = kind == OPM_FILTER_SHORT_PASS_STEP || kind == OPM_FILTER_SHORT_PASS_SMOOTH;
    const double pass = 1.0, block = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        const double l = lambda_um[i];
        if (!smooth) { values[i] = (short_pass ? l < cut : l > cut) ? pass : block; continue; }
        if (l <= cut - width / 2.0) values[i] = short_pass ? pass : block;
        else if (l >= cut + width) values[i] = short_pass ? block : pass;
        else {
            const double s = std::sin(pi / width * (l - cut));
            values[i] = std::fma(0.5, short_pass ? -s : s, 0.5);
        }
    }
    return OPM_OK;
}

// Spectrum::from_csv (:89-117): records of `wavelength in nm ; value in percent` (a Thorlabs transmission export).  The csv crate as
// the reference configures it: no header row, ';', empty lines skipped, every record as long as the first one; a field must parse
// as a whole (Rust's str::parse::<f64>: no blanks, no hexadecimal).
namespace {
bool parse_field(const std::string& f, double* out) {
    if (f.empty()) return false;
    std::string body = (f[0] == '+' || f[0] == '-') ? f.substr(1) : f;
    std::string low(body);
    std::transform(low.begin(), low.end(), low.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (low != "inf" && low != "infinity" && low != "nan" && f.find_first_not_of("0123456789+-.eE") != std::string::npos) return false;
    char* end = nullptr;
    const double v = std::strtod(f.c_str(), &end);
    if (end != f.c_str() + f.size()) return false;
    *out = v;
    return true;
}
}  // namespace
extern "C" OpmStatus opm_spectrum_from_csv(const char* path, double* lambda_um, double* values, uint64_t cap, uint64_t* n_out) {
    if (!path) return host_fail(OPM_ERR_ARG, "NULL path");
    std::ifstream in(path, std::ios::binary);
    if (!in) return host_fail(OPM_ERR_SPECTRUM, "csv file could not be opened");
    std::stringstream whole;
    whole << in.rdbuf();
    const std::string text = whole.str();
    std::vector<std::pair<double, double>> rows;
    size_t width = 0;
    for (size_t pos = 0; pos < text.size();) {
        const size_t eol = std::min(text.find_first_of("\r\n", pos), text.size());
        if (eol > pos) {
            std::vector<std::string> fields;
            for (size_t a = pos;;) {
                const size_t b = std::min(text.find(';', a), eol);
                fields.emplace_back(text, a, b - a);
                if (b >= eol) break;
                a = b + 1;
            }
            if (width == 0) width = fields.size();
            double l = 0.0, v = 0.0;
            if (fields.size() != width) return host_fail(OPM_ERR_SPECTRUM, "csv records of unequal lengths");
            if (!parse_field(fields[0], &l)) return host_fail(OPM_ERR_SPECTRUM, "invalid float literal in csv file");
            if (width < 2) return host_fail(OPM_ERR_SPECTRUM, "csv record without a value column");  // the reference panics here
            if (!parse_field(fields[1], &v)) return host_fail(OPM_ERR_SPECTRUM, "invalid float literal in csv file");
            rows.emplace_back(l * 1.0E-3, v * 0.01);  // nanometers -> micrometers, percent -> transmission
        }
        pos = eol;
        if (pos < text.size() && text[pos] == '\r') ++pos;
        if (pos < text.size() && text[pos] == '\n') ++pos;
    }
    if (rows.empty()) return host_fail(OPM_ERR_SPECTRUM, "no csv data was found in file");
    if (n_out) *n_out = rows.size();
    if (!lambda_um) return OPM_OK;
    if (!values || rows.size() > cap) return host_fail(OPM_ERR_ARG, "spectrum arrays too small");
    for (size_t i = 0; i < rows.size(); ++i) { lambda_um[i] = rows[i].first; values[i] = rows[i].second; }
    return OPM_OK;
}
