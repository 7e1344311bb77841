// opm_internal.h -- device-side POD layouts shared by the kernels and the C-ABI host code.
// Not part of the public boundary (include/opossum_b200.h is).
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

#include "../../include/opossum_b200.h"

namespace opm {

// valid byte of a ray slot
enum : uint8_t { RAY_INVALID = 0, RAY_VALID = 1, RAY_REMOVED = 2 };

// Struct-of-arrays view of a ray bundle in HBM: 10 f64 + 3 u32 + 1 u8 = 93 B per ray.
struct DevRays {
    double *px, *py, *pz, *dx, *dy, *dz, *e, *wvl, *opl, *n;
    uint32_t *bounces, *refr, *id;
    uint8_t* valid;
};
constexpr int kRayBytes = 10 * 8 + 3 * 4 + 1;

// Slot-indexed hit record of one surface pass: 4 f64 + u32 = 36 B per slot, e < 0 marks "no hit".
struct DevHits {
    double *x, *y, *z, *e;
    uint32_t* bounce;
};

struct DevIso { double r[9]; double t[3]; };

struct DevApItem { int32_t type, obstruction; double p[4]; };
struct DevAperture {
    int32_t n_items, is_stack, stack_obstruction, n_tris;
    DevApItem items[OPM_MAX_APERTURE_ITEMS];
    const double* tris;  // device
};

// internal bit of DevRefract::flags (the public OPM_REFRACT_* flags use the low bits): the reflected child of the ray in slot i
// goes to slot i of the children bundle, slots without a child become removed rays -- no append counter, no ordering left to
// the scheduler (opm_rays_expect_appends)
constexpr int32_t kRefractChildrenBySlot = 1 << 16;
enum : int32_t { OPK_REFRACT = 1, OPK_APODIZE, OPK_THRESHOLD, OPK_LIMITS, OPK_FILTER, OPK_PARAXIAL, OPK_SPLIT, OPK_SNAPSHOT, OPK_TRANSFORM,
                 OPK_DIFFRACT, OPK_FORK, OPK_JOIN };

struct DevRefract {
    int32_t geom, coating, has_index, refraction_intended, strategy, flags, index_type, _pad;
    double param, coating_r;
    double inv_abs_param;            // 1 / |param| (sphere / cylinder normals in the contracted build)
    DevIso fwd, inv;
    double ic[6];
    double wlo, whi;
    double hint_wvl;                 // > 0: the bundle is (believed to be) monochromatic at this wavelength, see RefractCt HINT
    DevHits hits;                    // hits.e == nullptr: not recorded
    unsigned long long* hit_acc;     // bounding-box accumulator of the hit map (BBoxAccum: 4 ordered keys + count) the specialised
                                     // kernel keeps while it records; nullptr: none
    DevRays children;                // children.px == nullptr: not emitted
    unsigned long long* child_tail;  // device counter = current length of the children bundle
    uint64_t child_cap;
};
// Rays::diffract_on_periodic_surface (rays.rs:1058-1111): the surface part (geometry, index model for the n2 check, children,
// flags) is a DevRefract; gvec = grating vector in the world frame, gnorm = |gvec|, order = diffraction order
struct DevDiffract { DevRefract surf; double gvec[3]; double gnorm; double order; };
struct DevApodize { DevIso inv; DevAperture ap; };
struct DevThreshold { double min_energy; };
// the counters of a ray are 32-bit: the limits travel saturated to 32 bits (bounces > 2^32 - 1 is never true; a refraction limit
// above 2^32 - 1 clears check_refr instead), so each test is one 32-bit compare
struct DevLimits { uint32_t max_bounces, max_refr; int32_t check_refr, _pad; };
inline uint32_t limit_sat32(uint64_t v) { return v > 0xffffffffull ? 0xffffffffu : (uint32_t)v; }
// a Spectrum (spectrum.rs:35-37) on the device: ascending wavelength slots [um] and their values; lam == nullptr: none
struct DevSpectrum { const double* lam; const double* val; uint64_t n; };
struct DevFilter { double t; DevSpectrum s; };   // s.lam != nullptr: FilterType::Spectrum (ray.rs:712-721)
struct DevParaxial { double f; DevIso fwd, inv; };
struct DevSplit { double ratio; DevRays out; DevSpectrum s; };  // s.lam != nullptr: SplittingConfig::Spectrum (ray.rs:755-761)
// fields: OPM_SNAPSHOT_ALL (every array of the bundle) or OPM_SNAPSHOT_SPOT (position, energy, valid flag: what a spot-diagram report
// reads).  acc != nullptr (specialised kernel only): the launch also takes the phase-1 sums of the spot statistics of the copy
// (SpotAccum, rays.rs:599-637) with the positions in the frame `inv` (has_frame) -- no separate pass over the copy
struct DevSnapshot { DevRays out; int32_t fields, has_frame; DevIso inv; void* acc; };
constexpr int kSpotSmemBytes = 7 * 8 + 8;  // per thread: running sums x y z ex ey ez e + first (id, bounces) key
struct DevFork { double ratio; DevSpectrum s; };
struct DevJoin { DevRays out; };
constexpr int kStashWords = 21;  // parked main-bundle ray of a FORK: pos, dir, e, opl, n (18 words) + bounces, refr, flags
struct DevTransform { DevIso iso; };

// Every op carries an epilogue: the energy threshold (rays.rs:1145-1162) and the bounce / refraction limits
// (rays.rs:1386-1404) that the reference applies after nearly every surface are folded into the op they follow, so
// a surface costs one dispatch instead of three.
enum : int32_t { EPI_THRESHOLD = 1, EPI_LIMITS = 2, EPI_CHECK_REFR = 4 };
struct alignas(16) DevOp {
    int32_t kind, epi_flags;
    double epi_min_energy;
    uint32_t epi_max_bounces, epi_max_refr;  // saturated, see DevLimits
    uint64_t _epi_pad;
    union {
        DevRefract refract;
        DevApodize apodize;
        DevThreshold threshold;
        DevLimits limits;
        DevFilter filter;
        DevParaxial paraxial;
        DevSplit split;
        DevSnapshot snapshot;
        DevTransform transform;
        DevDiffract diffract;
        DevFork fork;
        DevJoin join;
    } u;
};
static_assert(sizeof(DevOp) % 16 == 0, "DevOp must be a multiple of 16 B for the bulk copy into shared memory");

// Optional per-ray position history (`pos_hist`, ray.rs:70): a level-major log beside the bundle,
// base[(level * 3 + c) * stride + slot], NaN = the ray has no entry at this level.  Every op that pushes in the
// reference (refract :616, diffract :537, apodize of a ray it invalidates rays.rs:566) owns one level per launch.
struct HistParams {
    double* base;
    uint64_t stride;
    short level[OPM_MAX_PROGRAM_OPS];  // per device op, -1 = logs nothing
};

// Fluence helper rays (FluenceRays, rays.rs:1470-1567): three satellite rays per ray and the effective energy of the
// area element they span, field-major beside the bundle: base[f * stride + slot], f = 0..8 helper positions (helper k, component
// c at 3 k + c), 9..17 helper directions, 18..20 the helpers' refractive indices (NEGATIVE: that helper is invalid, ray.rs
// set_invalid after a missed surface), 21 the effective energy.  in == out for a launch that leaves the rays where they were.
constexpr int kHelperFields = 22;
struct HelperParams { const double* in; double* out; uint64_t in_stride, out_stride; };

// device mirror of OpmTraceStats (all counters are atomically accumulated)
struct DevStats {
    unsigned long long interactions, children, hits, missed, apodized, valid_out, alive_out;
    unsigned int status_bits, _pad;
    unsigned long long first_inv;  // max over valid rays of ~(id << 32 | bounces); 0 = no valid ray (zero-initialised)
    unsigned long long first_inv2; // the same right after the op flagged OPM_REFRACT_STAT_KEY2 (0: no such op, or no valid ray there)
    // Everything above is zeroed before every launch.  The word below is not: status bits raised by ANY launch since the host last
    // looked (asynchronous launches do not read `status_bits` back -- their errors surface at opm_context_synchronize /
    // opm_context_check_errors, like the reference's `?` aborts the analysis, rays.rs:987-991)
    unsigned int sticky_bits, _pad2;
    unsigned long long total_interactions;  // interactions of ALL launches of the context (ghost focus reads it once per analysis)
};
constexpr unsigned long long kDevStatsZeroBytes = 10 * 8;  // the per-launch part

struct TraceLaunch {
    const DevOp* ops;  // device
    int32_t n_ops;
    DevRays in, out;
    uint64_t n;
    uint64_t out_cap;
    unsigned long long* out_tail;  // compact mode: device counter of `out`
    DevStats* stats;
    int32_t compact;
    int32_t has_fork;  // the program parks rays in shared memory: kStashWords * 4 bytes per thread
    int32_t grid, block;
    void* stream;
    const HistParams* hist = nullptr;  // host pointer; non-NULL: the interpreter variant that logs positions (never compact)
    const HelperParams* helpers = nullptr;  // host pointer; non-NULL: the interpreter variant that carries fluence helper rays
};

// launchers implemented in opm_trace.cu (compiled twice: contracted "fast" and -fmad=false "strict")
int launch_trace_fast(const TraceLaunch& L);
int launch_trace_strict(const TraceLaunch& L);
int trace_block_fast();
int trace_block_strict();
int trace_occupancy_fast(int block, size_t smem);
int trace_occupancy_strict(int block, size_t smem);

// run-time specialised trace kernels (opm_jit.cu)
struct JitKernel;
struct JitCache;
JitCache* jit_cache_create();
void jit_cache_destroy(JitCache* c);
void jit_cache_stats(const JitCache* c, uint64_t* compiled, uint64_t* failures, uint64_t* launches, double* compile_ms);
void jit_cache_wait(JitCache* c);
const JitKernel* jit_get(JitCache* cache, const DevOp* ops, int n_ops, bool compact, int block, int min_ctas, bool wait,
                         const char** why);
int jit_occupancy(const JitKernel* k);
void* jit_const_ptr(const JitKernel* k);  // the kernel's __constant__ descriptor table
int jit_launch(JitCache* cache, const JitKernel* k, const TraceLaunch& L, const void* h_ops_pinned);
size_t jit_selftest(char* why, size_t cap);  // compiles a program with every op kind to a cubin (no GPU needed); 0 = failed

}  // namespace opm
