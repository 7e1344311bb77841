// opm_kernels.h -- launchers of the plumbing / detector / source kernels (opm_kernels.cu).
#pragma once
#include "opm_internal.h"

namespace opm {

struct SpotAccum {
    double sum[7];  // x y z e*x e*y e*z e  (valid rays)
    unsigned long long n_valid, n_total, first_key;  // first_key = min (id << 32 | bounces) over valid rays
    unsigned long long max_d2_key;                   // ordered key of max squared distance [mm^2] from the centroid
    double sum_d2, sum_ed2;
};
struct BBoxAccum { unsigned long long key[4]; unsigned long long count; };  // left,right,top,bottom as ordered keys
struct PeakAccum { unsigned long long peak_key; double total; };
struct PeakCheckAcc { BBoxAccum box; unsigned done, _pad; unsigned long long _pad2[2]; };  // 64 bytes (OPM_PEAK_CHECK_ACC_BYTES)
struct WvlAccum { unsigned long long min_key, max_key, count; };  // init: key(+inf), key(-inf), 0
// wavefront detector accumulators; init: min_r_key = key(+inf), first_slot = ~0, max_key = key(-inf), min_key = key(+inf), rest 0
struct WfAccum { unsigned long long min_r_key, first_slot, max_key, min_key, count; double sum, sum2; };
struct BinParams {
    unsigned long long nx, ny;
    double left, bottom, width_step, height_step, bin_area;
};
struct SourceParams {
    int32_t pos_dist, energy_dist, rectangular, _pad;
    uint64_t n_positions, grid_ny;
    uint32_t n_wavelengths, _pad2;
    double dist_x, dist_y, off_x, off_y;  // grid.rs:61-80
    double size_x, size_y, golden;
    double total_energy, e_uniform, energy_norm;
    double mu[2], sigma[2], power, sin_t, cos_t;
    DevIso iso;
};

// host-side inverse of the ordered-key encoding used by the atomicMin/Max reductions
inline double key_to_double(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    __builtin_memcpy(&d, &b, sizeof d);
    return d;
}
inline unsigned long long double_to_key(double d) {
    unsigned long long b;
    __builtin_memcpy(&b, &d, sizeof b);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

int k_aos_to_soa(const OpmRay* a, DevRays R, uint64_t off, uint64_t n, int sms, void* stream);
int k_soa_to_aos(DevRays R, OpmRay* a, uint64_t n, int sms, void* stream);
int k_soa_defaults(DevRays R, uint64_t n, int sms, void* stream);
int k_transform(DevRays R, DevIso I, uint64_t n, int sms, void* stream);
int k_spot_pass1(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc, int sms, void* stream);
int k_spot_pass2(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc, int sms, void* stream);
int k_bounce_lvl(DevRays R, uint64_t n, unsigned long long* first_key, int sms, void* stream);
int k_tally(DevRays R, uint64_t n, uint32_t n_orders, unsigned long long* counts, double* energy, int sms, void* stream);
int k_hits_bbox(DevHits H, uint64_t n, long long bounce_filter, BBoxAccum* acc, int sms, void* stream);
// device-resident chaining: the map range lives on the device as (-left, right, top, -bottom), see opm_kernels.cu
int k_bbox_finalize(const BBoxAccum* acc, double* range_dev, void* stream);
int k_peakcheck_acc_init(PeakCheckAcc* acc, uint64_t n, void* stream);
int k_peakcheck_bbox(DevHits H, uint64_t n, long long bounce_filter, PeakCheckAcc* acc, double* range_out, double* map, uint64_t bins, int first,
                     int last, int sms, void* stream);
int k_peakcheck_peak(const double* map, uint64_t nx, uint64_t ny, const unsigned long long* first_inv, double* result, void* stream);
// one deferred peak check: hit map, the bounce filter (>= 0, -1, or the negated address of the key word), its accumulator,
// result slot (range 4 | peak | total | status | key) and map
// (a check may sum several hit maps: the same bundle id met the surface more than once -- HitMap::get_rays_hit_map(bounce, uuid))
constexpr int kPeakCheckMaps = 4;
struct PeakCheckJob { DevHits H[kPeakCheckMaps]; uint64_t n[kPeakCheckMaps]; int n_maps, _pad; long long bounce_filter; PeakCheckAcc* acc; double* result; double* map; };
int k_peakcheck_batch(const PeakCheckJob* d_jobs, int n_jobs, int max_maps, const int* active, uint64_t max_n, uint64_t nx, uint64_t ny, int sms,
                      void* stream);
int k_binning(DevHits H, uint64_t n, long long bounce_filter, uint64_t nx, uint64_t ny, const double* range_dev, double* map,
              unsigned* status, int sms, void* stream);
int k_peak_total(const double* map, uint64_t nx, uint64_t ny, const double* range_dev, PeakAccum* acc, double* out2_dev, int sms, void* stream);
// kernel density estimator (kde/mod.rs)
int k_hits_pack(DevHits H, uint64_t n, long long bounce_filter, double* x, double* y, double* e, unsigned long long* count,
                unsigned long long cap, int sms, void* stream);
int k_kde_pair_distances(const double* x, const double* y, uint64_t n, double* dist, double* sum, double* dev, int sms, void* stream);
int k_sort_doubles(const double* in, double* out, uint64_t m, void* temp, size_t* temp_bytes, void* stream);
int k_kde_2d(const double* hx, const double* hy, const double* he, uint64_t n, double band_width, double left, double bottom, double dx,
             double dy, uint64_t nx, uint64_t ny, double* map, void* stream);
int k_spot_sums_store(const SpotAccum* acc, double* sums_dev, void* stream);
int k_spot_sums_load(const double* sums_dev, SpotAccum* acc, void* stream);
int k_spot_radii_store(const SpotAccum* acc, double* radii_dev, void* stream);
int k_source_norm(const SourceParams& S, uint64_t b, uint64_t e, double* out, int sms, void* stream);
int k_source_generate(const SourceParams& S, DevRays R, uint64_t pos_begin, uint64_t pos_stride, uint64_t n_rays, const double* wvl,
                      const double* wgt, const double* norm_dev, int sms, void* stream);
int k_expand_collimated(const double* xyz, const double* e_pos, const double* wvl, const double* wgt, uint32_t n_wvl, DevRays R,
                        uint64_t n_rays, int sms, void* stream);
int k_compact(DevRays S, DevRays D, uint64_t n, unsigned* counts, unsigned long long* offsets, unsigned long long* total, void* stream);
int k_history_gather(const HistParams& hp, uint32_t levels, DevRays R, uint64_t n, int with_current, uint32_t pos_cap, double* xyz,
                     uint32_t* len, int sms, void* stream);
int k_wavelength_range(DevRays R, uint64_t n, WvlAccum* acc, int sms, void* stream);
int k_spectrum_splat(DevRays R, uint64_t n, double start_um, double step_um, uint64_t nslots, double* data, unsigned* status, int sms,
                     void* stream);
int k_wavefront(DevRays R, uint64_t n, DevIso inv, double wvl_nm, WfAccum* acc, double* xyw, int sms, void* stream);
int k_readout_combine_a(const double* g, int w, int nf, int ns, double* out, void* stream);
int k_readout_combine_b(const double* g, int w, uint64_t maps_len, int ns, double* out, int sms, void* stream);
int k_fp64_peak(double* out, int iters, int blocks, void* stream);
int k_fetch_from_host(void* dst, const void* src_host, size_t bytes, void* stream);  // bytes % 8 == 0, src pinned + mapped
int k_set_u64(unsigned long long* p, unsigned long long v, void* stream);
// ---- Voronoi fluence estimator (opm_voronoi.cu)
struct VorGrid { double x0, y0, cs, inv_cs; int gx, gy; const int* cell_start; const int* cell_sites; };  // uniform grid over the sites
struct VorBound {  // the reference's bounding polygon (convex, m edges: a point and the OUTWARD normal of each), a rectangle around it,
    double x0, y0, x1, y1, cx, cy, r_in2;  // and the circle around (cx, cy) inscribed in it
    int m;
    const double *px, *py, *nx, *ny;
};
struct VorHull { int h; const double *x, *y, *z; };  // convex hull of the sites, counter-clockwise, with the sites' values
int k_voronoi_to_cm(int n, double* x, double* y, void* stream);
int k_voronoi_cells(int n, const double* sx, const double* sy, const double* e, const VorGrid& G, const VorBound& B, double* fluence,
                    unsigned* status, void* stream);
int k_voronoi_interp(int nx_pts, int ny_pts, double ax0, double ax_bin, double ay0, double ay_bin, int n, const double* sx, const double* sy,
                     const double* sz, const VorGrid& G, const VorHull& H, double big, double* map, unsigned* status, void* stream);
int k_voronoi_accumulate(uint64_t bins, double* map, const double* part, int sms, void* stream);
#ifdef OPM_VORONOI_HOST  // host bookkeeping of the estimator (needs <vector>)
void vor_convex_hull(const double* x, const double* y, int n, std::vector<int>& hull, int* on_boundary);
bool vor_all_on_same_line(const double* x, const double* y, int n);
void vor_bounding_polygon(const double* x, const double* y, const std::vector<int>& hull, int n_triangles, std::vector<double>& px,
                          std::vector<double>& py, double* cx_out, double* cy_out);
void vor_build_grid(const double* x, const double* y, int n, VorGrid& G, std::vector<int>& cell_start, std::vector<int>& cell_sites);
#endif
int k_hits_bounce_mask(DevHits H, uint64_t n, unsigned long long* mask, int sms, void* stream);
int k_copy_rays(DevRays S, uint64_t soff, DevRays D, uint64_t doff, uint64_t n, int sms, void* stream);
int k_first_key_store(const unsigned long long* first_inv, unsigned long long* out, void* stream);
int k_expand_collimated_xy(const double* xy, uint64_t n_pos, const SourceParams& S, double* norm_dev, const double* wvl, const double* wgt,
                           uint32_t n_wvl, DevRays R, uint64_t n_rays, int sms, void* stream);
constexpr int kPrologueTails = 8;
struct ProloguePlan {
    unsigned long long* dst; const unsigned long long* src_host; unsigned words;   // op descriptors (pinned + mapped source)
    unsigned long long* stats; unsigned stats_words;                               // zeroed
    unsigned long long* key_dst[2];  // NULL, or where the first-valid-ray keys of the PREVIOUS launch go before the statistics are zeroed
    BBoxAccum* acc[2];                                                             // reset to the identity (NULL: none)
    SpotAccum* spot[2];                                                            // reset to the identity (NULL: none)
    unsigned long long* tail[kPrologueTails]; unsigned long long tail_val[kPrologueTails]; int n_tails;
};
int k_trace_prologue(const ProloguePlan& P, void* stream);
int k_fill_u8(uint8_t* p, uint8_t v, uint64_t n, int sms, void* stream);
// FluenceRays::new for every slot (rays.rs:1477-1516); base[21 * stride + slot] holds the initial fluence on entry; cs = cos, sin x 3
int k_helpers_init(DevRays R, uint64_t n, double* base, uint64_t stride, const double cs[6], int sms, void* stream);

}  // namespace opm
