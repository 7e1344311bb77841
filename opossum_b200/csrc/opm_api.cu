// opm_api.cu -- implementation of the C ABI declared in include/opossum_b200.h.
//
// Host-side bookkeeping only: handle lifetimes, descriptor validation (same error conditions as the
// reference's OpmResult paths), conversion of public op descriptors into the device program, launches.
// All arithmetic on rays happens in the CUDA kernels (opm_trace.cu, opm_kernels.cu); there is no CPU path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>

#include "opm_internal.h"
#define OPM_VORONOI_HOST
#include "opm_kernels.h"
#include "opm_peer.h"
#include "../../include/opossum_b200_scene.h"

using namespace opm;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct OpmContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sms = 148;
    uint64_t launches = 0;
    DevStats* d_stats = nullptr;
    DevStats* h_stats = nullptr;  // pinned
    DevOp* d_ops = nullptr;
    // pinned staging ring for op programs: launches stay asynchronous (a pageable source would make every
    // cudaMemcpyAsync a stream synchronisation point)
    static constexpr int kRing = 8;
    DevOp* h_ops_ring = nullptr;
    cudaEvent_t ring_ev[kRing] = {};
    int ring_pos = 0;
    // small host -> device parameter blocks (accumulator initial values): pinned + mapped ring read by a fetch kernel,
    // so they never queue behind a bulk upload on the H2D copy engine
    static constexpr int kPutSlots = 32;
    static constexpr size_t kPutBytes = 256;
    double* d_srcspec = nullptr;  // spectrum table of the source being generated (opm_source_generate, <= 16 lines)
    unsigned char* h_put_ring = nullptr;
    cudaEvent_t put_ev[kPutSlots] = {};
    int put_pos = 0;
    // run-time specialised trace kernels: used by the contracted build for bundles of at least jit_min_rays rays
    JitCache* jit = nullptr;
    bool jit_enabled = true;
    bool jit_sync = false;  // wait for the compiler instead of interpreting until the kernel is ready
    uint64_t jit_min_rays = 200000;
    bool jit_fuse_spot = false;  // OPM_B200_JIT_SPOT=1: OPM_OP_SNAPSHOT want_spot_sums is honoured (see run_program)
    bool jit_warned = false;
    // CTA shape of the specialised kernels (tunable: OPM_B200_JIT_BLOCK / _MIN_CTAS).  One CTA of 32 warps per SM at 64 registers
    // per thread measured best on the C2 chain (profiles/r1_summary.md, CTA-shape sweep): the warps of one CTA walk the 70 KB of
    // straight-line code together and share its instruction-cache lines.
    // CTA shape of the specialised kernels.  jit_block 0 = by program: one CTA of 1024 threads per SM (64 registers), or of 896 (72
    // registers) for a program with a FORK, whose parked-ray bookkeeping is what makes the 64-register build spill -- measured on
    // C2 (FORK): 1.012 ms at 1024, 0.985 ms at 896 (960: 1.030, 832: 1.042); on C4 / C5 (no FORK) 1024 is 3-5 % ahead of 896
    int jit_block = 0, jit_min_ctas = 1;
    cudaStream_t copy_stream = nullptr;  // host uploads that are expanded on the device run here, overlapping the main stream
    // side computations nobody waits for until opm_context_join_side_stream / opm_context_synchronize (the ghost-focus analysis'
    // per-bundle fluence checks): created on first use
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ev = nullptr;
    bool side_busy = false;
    // job list of opm_fluence_peak_check_batch_async: pinned staging + device copy, reused once jobs_ev says the copy has drained
    // opm_fluence_peak_check_capture_key_async: where the latest launch's first-valid-ray keys are still to be copied -- the next
    // trace launch's prologue takes them along (it is what resets them), flush_pending_keys() copies what is left
    unsigned long long* pending_key[2] = {nullptr, nullptr};
    static constexpr int kJobRing = 4;  // batches whose job lists may be in flight before the host has to wait for the oldest copy
    PeakCheckJob* h_jobs = nullptr;
    PeakCheckJob* d_jobs = nullptr;
    int jobs_cap = 0, jobs_pos = 0;
    cudaEvent_t jobs_ev[kJobRing] = {};
    // optional per-launch timing of the fused trace kernel (CUDA events on the launching stream)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_free;
    double timing_ms = 0.0;
    uint64_t timing_launches = 0;
    unsigned char* d_scratch = nullptr;  // small accumulators
    unsigned char* h_scratch = nullptr;  // pinned mirror
    static constexpr size_t kScratch = 4096;
};

struct OpmRays {
    OpmContext* ctx = nullptr;
    uint64_t cap = 0, len = 0;
    void* base = nullptr;
    DevRays v{};
    unsigned long long* d_tail = nullptr;
    bool len_dirty = false;  // the device tail counter is ahead of `len` (asynchronous append)
    // > 0: every ray of the bundle is believed to carry this wavelength (one-wavelength sources).  A hint for the specialised
    // kernels (RefractCt HINT), which check it per ray: a stale value costs time, not correctness.
    double mono_wvl = 0.0;
    // host upload in flight (opm_rays_new_collimated_with_spectrum): positions / energies / spectrum are copied into
    // this bundle's staging buffer on the context's copy stream; the expansion into the SoA arrays is enqueued on the
    // main stream the first time the bundle is used, so the copy overlaps whatever the main stream is doing until then
    unsigned char* d_stage = nullptr;
    size_t stage_bytes = 0;
    double* h_spec = nullptr;  // pinned copy of the spectrum table (a pageable source would make the copy synchronous)
    size_t h_spec_bytes = 0;
    cudaEvent_t ev_uploaded = nullptr, ev_expanded = nullptr;
    bool pending = false;
    int pend_kind = 0;         // 0: positions (n x 3) + energies; 1: xy only, energies from pend_src on the device
    SourceParams pend_src{};
    uint64_t pend_n_pos = 0;
    uint32_t pend_n_wvl = 0;
    // optional per-ray position history (`pos_hist`, ray.rs:70): level-major log, see HistParams.
    // Levels [0, hist_next) are in use; a slot's NaN rows are levels at which that ray logged nothing.
    double* d_hist = nullptr;
    uint32_t hist_levels = 0, hist_next = 0;
    uint64_t hist_stride = 0;
    // phase-1 accumulators of the spot statistics (SpotAccum, 128 bytes behind the tail counter).  spot_acc_valid: the launch that
    // wrote this bundle as an OPM_OP_SNAPSHOT copy also took the sums (specialised kernel), in the frame spot_inv (spot_has_frame)
    // opm_rays_expect_appends: the extent a launch that appends to this bundle leaves it with (slots it does not fill are removed
    // rays); 0 = none expected, the device tail counter decides (len_dirty)
    // fluence helper rays (FluenceRays, rays.rs:1470-1567): kHelperFields arrays of helper_stride doubles, see HelperParams; NULL: none
    double* d_helpers = nullptr;
    uint64_t helper_stride = 0;
    uint64_t expected_len = 0;
    bool sealed = false;  // took its appends under opm_rays_expect_appends: the device counter no longer describes it
    SpotAccum* d_spot_acc = nullptr;
    bool spot_acc_valid = false, spot_has_frame = false;
    DevIso spot_inv{};
};

struct OpmSpectrum {  // spectrum.rs:35-37 on the device: [lam(n) | val(n)] in one allocation
    OpmContext* ctx = nullptr;
    uint64_t n = 0;
    double* d = nullptr;
};

struct OpmHitMap {
    OpmContext* ctx = nullptr;
    uint64_t cap = 0, len = 0;
    void* base = nullptr;
    DevHits v{};
    // bounding box of the recorded hits, kept by the specialised trace kernel while it writes the map (BBoxAccum); valid while
    // the map holds exactly what that launch recorded
    unsigned long long* d_acc = nullptr;
    bool acc_valid = false;
};

static OpmStatus hist_resize(OpmRays* r, uint32_t levels, uint64_t stride);
static OpmStatus hist_clear(OpmRays* r);

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static OpmStatus fail(OpmStatus s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return s;
}
static OpmStatus flush_pending_keys(OpmContext* c);
namespace opm {
// for the host-only translation units of the library (opm_spectrum_host.cpp)
OpmStatus host_fail(OpmStatus s, const char* msg) { return fail(s, "%s", msg); }
}  // namespace opm
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(OPM_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), \
                                           __FILE__, __LINE__, cudaGetErrorString(_e));                  \
    } while (0)
#define KL(expr)                                                                                         \
    do {                                                                                                 \
        int _e = (expr);                                                                                 \
        if (_e != 0) return fail(OPM_ERR_CUDA, "kernel launch failed: %s at %s:%d",                      \
                                 cudaGetErrorString((cudaError_t)_e), __FILE__, __LINE__);               \
    } while (0)

// device scratch layout: [0,1024) accumulators, [1024,1028) status, [2048,2080) range, [2112,2192) spot sums,
// [2240,2264) spot radii, [2304,2320) peak/total
static double* scratch_range(OpmContext* c) { return (double*)(c->d_scratch + 2048); }
static double* scratch_sums(OpmContext* c) { return (double*)(c->d_scratch + 2112); }
static double* scratch_radii(OpmContext* c) { return (double*)(c->d_scratch + 2240); }
static double* scratch_peak(OpmContext* c) { return (double*)(c->d_scratch + 2304); }

// asynchronous host -> device transfer of a small parameter block (bytes <= kPutBytes, multiple of 8) on the main stream
static OpmStatus dev_put(OpmContext* c, void* dst, const void* src, size_t bytes) {
    if (bytes > OpmContext::kPutBytes || (bytes & 7)) return fail(OPM_ERR_ARG, "dev_put: bad size %zu", bytes);
    const int slot = c->put_pos;
    c->put_pos = (c->put_pos + 1) % OpmContext::kPutSlots;
    CU(cudaEventSynchronize(c->put_ev[slot]));  // the fetch that last used this slot has run
    unsigned char* h = c->h_put_ring + (size_t)slot * OpmContext::kPutBytes;
    memcpy(h, src, bytes);
    KL(k_fetch_from_host(dst, h, bytes, c->stream));
    c->launches++;
    CU(cudaEventRecord(c->put_ev[slot], c->stream));
    return OPM_OK;
}

extern "C" const char* opm_status_string(OpmStatus s) {
    switch (s) {
        case OPM_OK: return "ok";
        case OPM_ERR_N2_INVALID: return "the refractive index must be >=1.0 and finite";
        case OPM_ERR_WVL_RANGE: return "wavelength outside valid range";
        case OPM_ERR_INDEX_MODEL: return "refractive index calculated by model is <1.0 or not finite";
        case OPM_ERR_COATING: return "reflectivity must be within [0.0, 1.0] and finite.";
        case OPM_ERR_APODIZE_FACTOR: return "transmission factor must be within (0.0..=1.0)";
        case OPM_ERR_THRESHOLD: return "threshold energy must be finite";
        case OPM_ERR_HITPOINT: return "hit point: value must be positive and finite, position must be finite";
        case OPM_ERR_SPLIT_RATIO: return "splitting_ratio must be within [0.0;1.0]";
        case OPM_ERR_FOCAL_LENGTH: return "focal length must be !=0.0 and finite";
        case OPM_ERR_EMPTY_HITMAP: return "could not calculate bounding box";
        case OPM_ERR_BIN_RANGE: return "hit point outside the given fluence map range";
        case OPM_ERR_ARG: return "invalid argument";
        case OPM_ERR_CAPACITY: return "output bundle or hit map too small";
        case OPM_ERR_CUDA: return "CUDA failure";
        case OPM_ERR_KDE_BANDWIDTH: return "bandwidth must be != 0.0 and finite";
        case OPM_ERR_SPECTRUM: return "spectrum: wavelength outside the spectrum, no valid rays, or invalid range / resolution";
        case OPM_ERR_EMPTY_WAVEFRONT: return "Empty wavefront-data vector!";
        case OPM_ERR_PEER_TIMEOUT: return "multi-GPU read-out: a peer rank did not arrive at a meeting in time";
        case OPM_ERR_VORONOI: return "Voronoi diagram for fluence estimation could not be created";
        default: return "unknown status";
    }
}
extern "C" const char* opm_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t opm_abi_version(void) { return OPM_ABI_VERSION; }
extern "C" uint64_t opm_abi_sizeof(const char* name) {
    if (!name) return 0;
#define OPM_SZ(T) if (!strcmp(name, #T)) return sizeof(T);
    OPM_SZ(OpmRay) OPM_SZ(OpmIso) OPM_SZ(OpmIsometry) OPM_SZ(OpmSurface) OPM_SZ(OpmIndexModel) OPM_SZ(OpmApertureItem) OPM_SZ(OpmAperture)
    OPM_SZ(OpmTraceStats) OPM_SZ(OpmRefractOp) OPM_SZ(OpmDiffractOp) OPM_SZ(OpmFilterOp) OPM_SZ(OpmSplitOp) OPM_SZ(OpmForkOp) OPM_SZ(OpmOp)
    OPM_SZ(OpmSpotStats) OPM_SZ(OpmSpotPartial) OPM_SZ(OpmSourceDesc) OPM_SZ(OpmRayTraceConfig) OPM_SZ(OpmWavefrontStats)
    OPM_SZ(OpmNodeDesc) OPM_SZ(OpmCoatingDesc) OPM_SZ(OpmGhostFocusConfig) OPM_SZ(OpmGhostFocusStats) OPM_SZ(OpmCriticalFluence)
#undef OPM_SZ
    return 0;
}

static OpmStatus status_from_bits(uint32_t bits) {
    if (!bits) return OPM_OK;
    for (int k = 1; k < 32; ++k)
        if (bits & (1u << k)) return fail((OpmStatus)k, "%s", opm_status_string((OpmStatus)k));
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_context_create(int32_t device, void* stream, OpmContext** out) {
    if (!out) return fail(OPM_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(OPM_ERR_CUDA, "no CUDA device available (%s): opossum_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(OPM_ERR_ARG, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(OPM_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    OpmContext* c = new OpmContext();
    c->device = device;
    c->sms = prop.multiProcessorCount;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    c->jit = jit_cache_create();
    if (const char* e = getenv("OPM_B200_JIT")) { c->jit_enabled = atoi(e) != 0; c->jit_sync = atoi(e) == 2; }
    if (const char* e = getenv("OPM_B200_JIT_MIN_RAYS")) c->jit_min_rays = strtoull(e, nullptr, 10);
    if (const char* e = getenv("OPM_B200_JIT_SPOT")) c->jit_fuse_spot = e[0] == '1';
    if (const char* e = getenv("OPM_B200_JIT_BLOCK")) { int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) c->jit_block = v; }
    if (const char* e = getenv("OPM_B200_JIT_MIN_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 16) c->jit_min_ctas = v; }
    CU(cudaMallocHost(&c->h_put_ring, OpmContext::kPutSlots * OpmContext::kPutBytes));
    for (int k = 0; k < OpmContext::kPutSlots; ++k) CU(cudaEventCreateWithFlags(&c->put_ev[k], cudaEventDisableTiming));
    CU(cudaMalloc(&c->d_stats, sizeof(DevStats)));
    CU(cudaMemset(c->d_stats, 0, sizeof(DevStats)));
    CU(cudaMallocHost(&c->h_stats, sizeof(DevStats)));
    CU(cudaMalloc(&c->d_ops, sizeof(DevOp) * OPM_MAX_PROGRAM_OPS * OpmContext::kRing));
    CU(cudaMallocHost(&c->h_ops_ring, sizeof(DevOp) * OPM_MAX_PROGRAM_OPS * OpmContext::kRing));
    for (int k = 0; k < OpmContext::kRing; ++k) CU(cudaEventCreateWithFlags(&c->ring_ev[k], cudaEventDisableTiming));
    CU(cudaMalloc(&c->d_scratch, OpmContext::kScratch));
    CU(cudaMallocHost(&c->h_scratch, OpmContext::kScratch));
    *out = c;
    return OPM_OK;
}
extern "C" void opm_context_destroy(OpmContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < OpmContext::kRing; ++k) if (c->ring_ev[k]) cudaEventDestroy(c->ring_ev[k]);
    cudaFreeHost(c->h_ops_ring);
    for (auto* v : {&c->timing_events, &c->timing_free})
        for (auto& p : *v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); cudaEventDestroy(c->side_ev); }
    for (int k = 0; k < OpmContext::kJobRing; ++k) if (c->jobs_ev[k]) cudaEventDestroy(c->jobs_ev[k]);
    cudaFreeHost(c->h_jobs); cudaFree(c->d_jobs);
    jit_cache_destroy(c->jit);
    for (int k = 0; k < OpmContext::kPutSlots; ++k) if (c->put_ev[k]) cudaEventDestroy(c->put_ev[k]);
    cudaFreeHost(c->h_put_ring);
    cudaFree(c->d_stats); cudaFreeHost(c->h_stats); cudaFree(c->d_ops); cudaFree(c->d_scratch); cudaFreeHost(c->h_scratch); cudaFree(c->d_srcspec);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}
// status bits raised by asynchronous launches since the last look (DevStats::sticky_bits); clear: forget them afterwards
extern "C" OpmStatus opm_context_check_errors(OpmContext* c, int32_t clear) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    CU(cudaSetDevice(c->device));
    unsigned* h = (unsigned*)(c->h_scratch + 1032);
    CU(cudaMemcpyAsync(h, &c->d_stats->sticky_bits, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const unsigned bits = *h;
    if (bits && clear) CU(cudaMemsetAsync(&c->d_stats->sticky_bits, 0, 4, c->stream));
    return status_from_bits(bits);
}
// Waits for the context's stream and reports what asynchronous launches raised meanwhile: a wavelength outside an index
// model's range, an invalid n2, a non-finite hit point, a full child bundle, a bad apodisation factor -- conditions on which the
// reference aborts the analysis (`?` in rays.rs:987-991) -- come back here as the error of the first such bit (then cleared).
extern "C" OpmStatus opm_context_synchronize(OpmContext* c) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    CU(cudaSetDevice(c->device));
    if (OpmStatus s = flush_pending_keys(c)) return s;
    if (c->side_busy) { CU(cudaStreamSynchronize(c->side_stream)); c->side_busy = false; }
    CU(cudaStreamSynchronize(c->stream));
    return opm_context_check_errors(c, 1);
}
// interactions of all launches of the context so far (a device counter no launch resets); synchronises the stream
extern "C" OpmStatus opm_context_interactions_total(OpmContext* c, uint64_t* out) {
    if (!c || !out) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    unsigned long long* h = (unsigned long long*)(c->h_scratch + 1040);
    CU(cudaMemcpyAsync(h, &c->d_stats->total_interactions, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = *h;
    return OPM_OK;
}
// first_valid_key of the context's latest trace launch, copied on the device (asynchronous): ~0 when the bundle had no valid ray
extern "C" OpmStatus opm_context_last_first_valid_key_async(OpmContext* c, uint64_t* dst_dev) {
    if (!c || !dst_dev) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    KL(k_first_key_store(&c->d_stats->first_inv, (unsigned long long*)dst_dev, c->stream)); c->launches++;
    return OPM_OK;
}
// ---- run-time specialised kernels (opm_jit.cu)
extern "C" uint64_t opm_jit_selftest(char* log, uint64_t cap) { return jit_selftest(log, cap); }
extern "C" OpmStatus opm_context_set_jit(OpmContext* c, int32_t enable, uint64_t min_rays) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    c->jit_enabled = enable != 0;
    c->jit_sync = enable == 2;
    c->jit_min_rays = min_rays;
    return OPM_OK;
}
extern "C" OpmStatus opm_context_jit_wait(OpmContext* c) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    jit_cache_wait(c->jit);
    return OPM_OK;
}
extern "C" OpmStatus opm_context_jit_stats(const OpmContext* c, uint64_t* kernels_compiled, uint64_t* compile_failures,
                                           uint64_t* launches, double* compile_ms) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    jit_cache_stats(c->jit, kernels_compiled, compile_failures, launches, compile_ms);
    return OPM_OK;
}
extern "C" void* opm_context_stream(OpmContext* c) { return c ? (void*)c->stream : nullptr; }
// per-launch device time of the fused trace kernel: an event pair around every launch while enabled
extern "C" OpmStatus opm_context_set_kernel_timing(OpmContext* c, int32_t enable) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    c->timing = enable != 0;
    return OPM_OK;
}
extern "C" OpmStatus opm_context_kernel_timing(OpmContext* c, double* total_ms, uint64_t* launches, int32_t reset) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    CU(cudaSetDevice(c->device));
    for (auto& p : c->timing_events) {
        CU(cudaEventSynchronize(p.second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p.first, p.second));
        c->timing_ms += ms;
        c->timing_launches++;
        c->timing_free.push_back(p);
    }
    c->timing_events.clear();
    if (total_ms) *total_ms = c->timing_ms;
    if (launches) *launches = c->timing_launches;
    if (reset) { c->timing_ms = 0.0; c->timing_launches = 0; }
    return OPM_OK;
}
extern "C" uint64_t opm_context_launch_count(const OpmContext* c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------------------------
// isometry (utils/geom_transformation.rs) -- host math in 3x3 + t form
// ---------------------------------------------------------------------------------------------
static void iso_invert(const OpmIso& f, OpmIso& inv) {  // (R^T, -R^T t)  geom_transformation.rs:226-229
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) inv.rot[r * 3 + c] = f.rot[c * 3 + r];
    for (int r = 0; r < 3; ++r)
        inv.t[r] = -(inv.rot[r * 3 + 0] * f.t[0] + inv.rot[r * 3 + 1] * f.t[1] + inv.rot[r * 3 + 2] * f.t[2]);
}
extern "C" void opm_isometry_identity(OpmIsometry* o) {
    memset(o, 0, sizeof *o);
    o->fwd.rot[0] = o->fwd.rot[4] = o->fwd.rot[8] = 1.0;
    o->inv.rot[0] = o->inv.rot[4] = o->inv.rot[8] = 1.0;
}
// Isometry::new: rotation applied in the order x -> y -> z, i.e. R = Rz(az) * Ry(ay) * Rx(ax)
// (nalgebra Rotation3::from_euler_angles(roll, pitch, yaw), geom_transformation.rs:44-72)
extern "C" OpmStatus opm_isometry_new(const double t[3], const double e[3], OpmIsometry* o) {
    for (int k = 0; k < 3; ++k) {
        if (!std::isfinite(t[k])) return fail(OPM_ERR_ARG, "translation coordinates must be finite");
        if (!std::isfinite(e[k])) return fail(OPM_ERR_ARG, "axis angles must be finite");
    }
    double sr = sin(e[0]), cr = cos(e[0]), sp = sin(e[1]), cp = cos(e[1]), sy = sin(e[2]), cy = cos(e[2]);
    double m[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                   sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                   -sp,     cp * sr,                cp * cr};
    memcpy(o->fwd.rot, m, sizeof m);
    memcpy(o->fwd.t, t, 3 * sizeof(double));
    iso_invert(o->fwd, o->inv);
    return OPM_OK;
}
extern "C" void opm_isometry_append(const OpmIsometry* a, const OpmIsometry* b, OpmIsometry* o) {  // a o b: b first (:216-223)
    OpmIso f;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            f.rot[r * 3 + c] = a->fwd.rot[r * 3 + 0] * b->fwd.rot[0 * 3 + c] + a->fwd.rot[r * 3 + 1] * b->fwd.rot[1 * 3 + c] +
                               a->fwd.rot[r * 3 + 2] * b->fwd.rot[2 * 3 + c];
    for (int r = 0; r < 3; ++r)
        f.t[r] = a->fwd.rot[r * 3 + 0] * b->fwd.t[0] + a->fwd.rot[r * 3 + 1] * b->fwd.t[1] + a->fwd.rot[r * 3 + 2] * b->fwd.t[2] + a->fwd.t[r];
    o->fwd = f;
    iso_invert(o->fwd, o->inv);
}
// Isometry::new_from_view (geom_transformation.rs:166-185) = Isometry3::face_towards(eye, eye + dir, up): rotation columns
// x = unit(up x z), y = unit(z x x), z = unit((eye + dir) - eye); nalgebra keeps the rotation as a unit quaternion, here it stays
// the matrix (the round trip through the quaternion only adds rounding)
extern "C" OpmStatus opm_isometry_from_view(const double p[3], const double dir[3], const double up[3], OpmIsometry* o) {
    if (!p || !dir || !up || !o) return fail(OPM_ERR_ARG, "null argument");
    double z[3], x[3], y[3];
    for (int k = 0; k < 3; ++k) z[k] = (p[k] + dir[k]) - p[k];
    auto unit = [](double v[3]) {
        const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int k = 0; k < 3; ++k) v[k] /= n;
        return n;
    };
    auto cross = [](const double a[3], const double b[3], double c[3]) {
        c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
    };
    if (!(unit(z) > 0.0)) return fail(OPM_ERR_ARG, "view direction must not be zero");
    cross(up, z, x);
    if (!(unit(x) > 0.0)) return fail(OPM_ERR_ARG, "up direction must not be collinear to the view direction");
    cross(z, x, y);
    unit(y);
    for (int r = 0; r < 3; ++r) {
        o->fwd.rot[3 * r + 0] = x[r]; o->fwd.rot[3 * r + 1] = y[r]; o->fwd.rot[3 * r + 2] = z[r];
        o->fwd.t[r] = p[r];
        if (!std::isfinite(p[r]) || !std::isfinite(x[r]) || !std::isfinite(y[r]) || !std::isfinite(z[r])) return fail(OPM_ERR_ARG, "view must be finite");
    }
    iso_invert(o->fwd, o->inv);
    return OPM_OK;
}
// ParabolicMirror::calc_off_axis_isometry / calc_parent_focal_length (nodes/parabolic_mirror.rs:286-320)
extern "C" OpmStatus opm_parabola_off_axis_isometry(double f, double oa_angle, const double dir_xy[2], int32_t collimating, OpmIsometry* out,
                                                    double* parent_focal_length) {
    if (!out) return fail(OPM_ERR_ARG, "NULL argument");
    if (!std::isnormal(f)) return fail(OPM_ERR_ARG, "focal length must not be 0.0 and finite");
    // ParabolicMirror::check_attributes (nodes/parabolic_mirror.rs:231-259)
    if (!std::isfinite(oa_angle)) return fail(OPM_ERR_ARG, "off-axis angle and finite");
    if (std::fabs(oa_angle) >= M_PI) return fail(OPM_ERR_ARG, "off-axis angle must be smaller than 180 deg");
    double dx = dir_xy ? dir_xy[0] : 1.0, dy = dir_xy ? dir_xy[1] : 0.0;
    if (dx == 0.0 && dy == 0.0) dx = 1.0;  // descriptor default: (1, 0)
    if (!std::isfinite(dx) || !std::isfinite(dy) || std::sqrt(dx * dx + dy * dy) < 2.220446049250313e-16)
        return fail(OPM_ERR_ARG, "off-axis direction values must be finite and the vector norm non-zero");
    const double tan_val = tan(oa_angle / 2.);
    const double parent_f = f / fma(tan_val, tan_val, 1.);
    const double decenter_x = sin(oa_angle) * f;
    const double z_shift = f * (1. / fma(tan_val, tan_val, 1.) - cos(oa_angle));
    const double z_rot = atan2(dy, dx);
    OpmIsometry iso, rot_iso, tot;
    const double tr[3] = {decenter_x, 0., z_shift}, zero[3] = {0., 0., 0.}, rz[3] = {0., 0., z_rot};
    OpmStatus s = opm_isometry_new(tr, zero, &iso);
    if (!s) s = opm_isometry_new(zero, rz, &rot_iso);
    if (s) return s;
    opm_isometry_append(&rot_iso, &iso, &tot);
    if (collimating) {  // half a turn about the transformed normal vector (Isometry3::new(0, unit * pi)): R = 2 a a^T - 1
        const double nv[3] = {decenter_x, 0., -2. * parent_f};
        double a[3];
        for (int r = 0; r < 3; ++r) a[r] = tot.fwd.rot[3 * r] * nv[0] + tot.fwd.rot[3 * r + 1] * nv[1] + tot.fwd.rot[3 * r + 2] * nv[2];
        const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        for (int r = 0; r < 3; ++r) a[r] /= n;
        OpmIsometry rot, t2;
        opm_isometry_identity(&rot);
        for (int r = 0; r < 3; ++r)
            for (int cidx = 0; cidx < 3; ++cidx) rot.fwd.rot[3 * r + cidx] = 2. * a[r] * a[cidx] - (r == cidx ? 1. : 0.);
        iso_invert(rot.fwd, rot.inv);
        opm_isometry_append(&rot, &tot, &t2);
        tot = t2;
    }
    *out = tot;
    if (parent_focal_length) *parent_focal_length = parent_f;
    return OPM_OK;
}
extern "C" void opm_raytrace_config_default(OpmRayTraceConfig* c) {  // analyzers/raytrace.rs:315-322
    c->min_energy_per_ray = 1.0 * 1.0E-12;
    c->max_number_of_bounces = 1000;
    c->max_number_of_refractions = 1000;
    c->missed_surface_strategy = OPM_MISS_STOP;
    c->_pad = 0;
}

// ---------------------------------------------------------------------------------------------
// rays storage
// ---------------------------------------------------------------------------------------------

static_assert(sizeof(SpotAccum) <= 128, "SpotAccum slot of a bundle");
static size_t rays_layout(uint64_t cap, void* base, DevRays* v, unsigned long long** tail, SpotAccum** spot) {
    size_t off = 0;
    unsigned char* b = (unsigned char*)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? b + o : nullptr; };
    double** d[10] = {&v->px, &v->py, &v->pz, &v->dx, &v->dy, &v->dz, &v->e, &v->wvl, &v->opl, &v->n};
    for (auto p : d) *p = (double*)take(cap * 8);
    v->bounces = (uint32_t*)take(cap * 4);
    v->refr = (uint32_t*)take(cap * 4);
    v->id = (uint32_t*)take(cap * 4);
    v->valid = (uint8_t*)take(cap);
    *tail = (unsigned long long*)take(8);
    if (spot) *spot = (SpotAccum*)take(128);
    return off;
}

static OpmStatus rays_alloc(OpmRays* r, uint64_t cap) {
    DevRays tmp; unsigned long long* t; SpotAccum* sp;
    size_t bytes = rays_layout(cap, nullptr, &tmp, &t, &sp);
    void* base = nullptr;
    CU(cudaMalloc(&base, bytes));
    r->base = base; r->cap = cap;
    rays_layout(cap, base, &r->v, &r->d_tail, &r->d_spot_acc);
    r->spot_acc_valid = false;
    CU(cudaMemsetAsync(r->d_tail, 0, 8, r->ctx->stream));
    return OPM_OK;
}

extern "C" OpmStatus opm_rays_create(OpmContext* ctx, uint64_t capacity, OpmRays** out) {
    if (!ctx || !out) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    OpmRays* r = new OpmRays();
    r->ctx = ctx;
    OpmStatus s = rays_alloc(r, capacity ? capacity : 1);
    if (s != OPM_OK) { delete r; return s; }
    *out = r;
    return OPM_OK;
}
extern "C" void opm_rays_destroy(OpmRays* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    if (r->d_stage) cudaStreamSynchronize(r->ctx->copy_stream);
    cudaStreamSynchronize(r->ctx->stream);
    cudaFree(r->base);
    cudaFree(r->d_hist);
    cudaFree(r->d_helpers);
    cudaFree(r->d_stage);
    cudaFreeHost(r->h_spec);
    if (r->ev_uploaded) cudaEventDestroy(r->ev_uploaded);
    if (r->ev_expanded) cudaEventDestroy(r->ev_expanded);
    delete r;
}
// enqueues the expansion of a pending host upload on the main stream (after the copy has landed)
static OpmStatus rays_materialize(OpmRays* r) {
    if (!r->pending) return OPM_OK;
    OpmContext* c = r->ctx;
    r->pending = false;
    const uint64_t n_pos = r->pend_n_pos, n_rays = n_pos * r->pend_n_wvl;
    double* d_xyz = (double*)r->d_stage;
    double* d_e = d_xyz + 3 * n_pos;
    double* d_w = (double*)(r->d_stage + align_up(n_pos * 32, 256));
    CU(cudaStreamWaitEvent(c->stream, r->ev_uploaded, 0));
    if (r->pend_kind == 1) {  // [xy n_pos x 2 | norm (8 bytes) ... | spectrum]
        double* d_xy = (double*)r->d_stage;
        double* d_norm = d_xy + 2 * n_pos;
        double* d_spec = (double*)(r->d_stage + align_up(n_pos * 16 + 8, 256));
        KL(k_expand_collimated_xy(d_xy, n_pos, r->pend_src, d_norm, d_spec, d_spec + r->pend_n_wvl, r->pend_n_wvl, r->v, n_rays, c->sms, c->stream));
        c->launches += 1 + ((r->pend_src.energy_dist == OPM_EDIST_GENERAL_GAUSSIAN && !(r->pend_src.energy_norm > 0.0)) ? 1 : 0);
    } else if (n_rays) { KL(k_expand_collimated(d_xyz, d_e, d_w, d_w + r->pend_n_wvl, r->pend_n_wvl, r->v, n_rays, c->sms, c->stream)); c->launches++; }
    CU(cudaEventRecord(r->ev_expanded, c->stream));
    KL(k_set_u64(r->d_tail, n_rays, c->stream)); c->launches++;
    return OPM_OK;
}
static OpmStatus rays_refresh_len(OpmRays* r) {
    if (r->pending) { OpmStatus s = rays_materialize(r); if (s) return s; }
    if (!r->len_dirty) return OPM_OK;
    unsigned long long t = 0;
    CU(cudaMemcpyAsync(&t, r->d_tail, 8, cudaMemcpyDeviceToHost, r->ctx->stream));
    CU(cudaStreamSynchronize(r->ctx->stream));
    r->len = t > r->cap ? r->cap : t;
    r->len_dirty = false;
    return OPM_OK;
}
// lengths of several bundles that launches appended to (ghost-focus children), with ONE synchronisation
extern "C" OpmStatus opm_rays_refresh_lens(OpmRays* const* list, int32_t n) {
    if (n <= 0) return OPM_OK;
    if (!list) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = nullptr;
    std::vector<OpmRays*> dirty;
    for (int k = 0; k < n; ++k) {
        OpmRays* r = list[k];
        if (!r) return fail(OPM_ERR_ARG, "NULL bundle");
        if (r->pending) { OpmStatus s = rays_materialize(r); if (s) return s; }
        if (!r->len_dirty) continue;
        if (c && r->ctx != c) return fail(OPM_ERR_ARG, "bundles of different contexts");
        c = r->ctx;
        dirty.push_back(r);
    }
    if (dirty.empty()) return OPM_OK;
    CU(cudaSetDevice(c->device));
    constexpr size_t kBatch = 256;  // 2 KB of the pinned scratch area at a time
    for (size_t b = 0; b < dirty.size(); b += kBatch) {
        const size_t m = std::min(kBatch, dirty.size() - b);
        unsigned long long* h = (unsigned long long*)(c->h_scratch + 2048);
        for (size_t k = 0; k < m; ++k) CU(cudaMemcpyAsync(h + k, dirty[b + k]->d_tail, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (size_t k = 0; k < m; ++k) {
            OpmRays* r = dirty[b + k];
            r->len = h[k] > r->cap ? r->cap : h[k];
            r->len_dirty = false;
        }
    }
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_expect_appends(OpmRays* r, uint64_t n_slots) {
    if (!r) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    if (r->pending) { OpmStatus s = rays_materialize(r); if (s) return s; }
    if (r->len != 0 || r->len_dirty) return fail(OPM_ERR_ARG, "opm_rays_expect_appends needs an empty bundle");
    OpmStatus s = opm_rays_reserve(r, n_slots ? n_slots : 1);
    if (s) return s;
    r->expected_len = n_slots ? n_slots : 1;
    r->spot_acc_valid = false;
    return OPM_OK;
}
extern "C" uint64_t opm_rays_len(const OpmRays* r) {
    if (!r) return 0;
    rays_refresh_len(const_cast<OpmRays*>(r));
    return r->len;
}
extern "C" uint64_t opm_rays_capacity(const OpmRays* r) { return r ? r->cap : 0; }
extern "C" OpmStatus opm_rays_set_wavelength_hint(OpmRays* r, double wavelength) {
    if (!r || !(wavelength >= 0.0) || !std::isfinite(wavelength)) return fail(OPM_ERR_ARG, "wavelength hint must be >= 0 and finite");
    r->mono_wvl = wavelength;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_clear(OpmRays* r) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    r->len = 0; r->len_dirty = false; r->pending = false; r->mono_wvl = 0.0; r->spot_acc_valid = false; r->expected_len = 0; r->sealed = false;
    if (r->d_hist) { OpmStatus hs = hist_clear(r); if (hs) return hs; }  // new rays start without a history (ray.rs:126)
    if (r->d_helpers) { CU(cudaStreamSynchronize(r->ctx->stream)); cudaFree(r->d_helpers); r->d_helpers = nullptr; r->helper_stride = 0; }
    CU(cudaMemsetAsync(r->d_tail, 0, 8, r->ctx->stream));
    return OPM_OK;
}
// ---- fluence helper rays ------------------------------------------------------------------------
static OpmStatus helpers_alloc(OpmRays* r, uint64_t stride) {
    if (stride < 1) stride = 1;
    if (r->d_helpers && r->helper_stride >= stride) return OPM_OK;
    if (r->d_helpers) { CU(cudaStreamSynchronize(r->ctx->stream)); cudaFree(r->d_helpers); r->d_helpers = nullptr; }
    CU(cudaMalloc(&r->d_helpers, (size_t)kHelperFields * stride * sizeof(double)));
    r->helper_stride = stride;
    return OPM_OK;
}
static OpmStatus helpers_copy(const OpmRays* src, OpmRays* dst, uint64_t n) {
    OpmStatus s = helpers_alloc(dst, n);
    if (s) return s;
    if (n) CU(cudaMemcpy2DAsync(dst->d_helpers, dst->helper_stride * 8, src->d_helpers, src->helper_stride * 8, n * 8, kHelperFields,
                                cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_attach_fluence_helpers(OpmRays* r, const double* init_fluence) {
    if (!r || !init_fluence) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    const uint64_t n = r->len;
    for (uint64_t i = 0; i < n; ++i)  // FluenceRays::new (rays.rs:1484-1488)
        if (!std::isfinite(init_fluence[i]) || std::signbit(init_fluence[i]))
            return fail(OPM_ERR_ARG, "Initial fluence of Fluence Rays must be non-negative and finite!");
    s = helpers_alloc(r, n);
    if (s) return s;
    if (n == 0) return OPM_OK;
    double* eff = r->d_helpers + 21 * r->helper_stride;  // the fluences travel in the slot of the effective energy they become
    CU(cudaMemcpyAsync(eff, init_fluence, n * 8, cudaMemcpyHostToDevice, c->stream));
    double cs[6];
    for (int k = 0; k < 3; ++k) {  // rays.rs:1495-1499: usize_to_f64(i) * 2. / 3. * PI
        const double a = (double)k * 2. / 3. * 3.14159265358979323846264338327950288;
        cs[2 * k] = std::cos(a); cs[2 * k + 1] = std::sin(a);
    }
    KL(k_helpers_init(r->v, n, r->d_helpers, r->helper_stride, cs, c->sms, c->stream)); c->launches++;
    CU(cudaStreamSynchronize(c->stream));  // init_fluence is the caller's (pageable) memory
    return OPM_OK;
}
extern "C" int32_t opm_rays_has_fluence_helpers(const OpmRays* r) { return r && r->d_helpers ? 1 : 0; }
extern "C" OpmStatus opm_rays_download_fluence_helpers(const OpmRays* rc, double* out, uint64_t capacity, uint64_t* n_out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !out || !n_out) return fail(OPM_ERR_ARG, "NULL argument");
    if (!r->d_helpers) return fail(OPM_ERR_ARG, "the bundle carries no fluence helper rays (opm_rays_attach_fluence_helpers)");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    const uint64_t n = r->len;
    *n_out = 0;
    if (n == 0) return OPM_OK;
    std::vector<double> tmp((size_t)kHelperFields * n);
    std::vector<uint8_t> valid(n);
    CU(cudaMemcpy2DAsync(tmp.data(), n * 8, r->d_helpers, r->helper_stride * 8, n * 8, kHelperFields, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(valid.data(), r->v.valid, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (valid[i] == RAY_REMOVED) continue;  // the order and the rays of opm_rays_download
        if (k >= capacity) return fail(OPM_ERR_CAPACITY, "download buffer too small");
        for (int f = 0; f < kHelperFields; ++f) out[k * kHelperFields + f] = tmp[(size_t)f * n + i];
        ++k;
    }
    *n_out = k;
    return OPM_OK;
}
static OpmStatus copy_fields(const DevRays& s, uint64_t soff, const DevRays& d, uint64_t doff, uint64_t n, cudaStream_t st) {
    if (n == 0) return OPM_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    KL(k_copy_rays(s, soff, d, doff, n, sms, st));  // one launch for the fourteen arrays
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_reserve(OpmRays* r, uint64_t capacity) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    if (capacity <= r->cap) return OPM_OK;
    CU(cudaSetDevice(r->ctx->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    OpmRays old = *r;
    s = rays_alloc(r, capacity);
    if (s) { *r = old; return s; }
    s = copy_fields(old.v, 0, r->v, 0, old.len, r->ctx->stream);
    if (s) return s;
    KL(k_set_u64(r->d_tail, r->len, r->ctx->stream));
    CU(cudaStreamSynchronize(r->ctx->stream));
    cudaFree(old.base);
    if (r->d_hist && r->hist_stride < capacity) return hist_resize(r, r->hist_levels, capacity);
    return OPM_OK;
}
static OpmStatus rays_set_len(OpmRays* r, uint64_t n) {
    r->len = n; r->len_dirty = false;
    KL(k_set_u64(r->d_tail, n, r->ctx->stream));  // value travels as a kernel argument: no host staging, fully asynchronous
    r->ctx->launches++;
    return OPM_OK;
}

// ---- position history ---------------------------------------------------------------------------
// (re)allocates the log for `levels` levels of `stride` slots; used levels are carried over, everything else reads NaN
static OpmStatus hist_resize(OpmRays* r, uint32_t levels, uint64_t stride) {
    if (levels < 1) levels = 1;
    if (stride < r->cap) stride = r->cap;
    if (stride < 1) stride = 1;
    OpmContext* c = r->ctx;
    double* nb = nullptr;
    const size_t bytes = (size_t)levels * 3 * stride * sizeof(double);
    CU(cudaMalloc(&nb, bytes));
    CU(cudaMemsetAsync(nb, 0xFF, bytes, c->stream));  // all-ones = a quiet NaN
    if (r->d_hist) {
        const uint32_t keep = r->hist_next < levels ? r->hist_next : levels;
        const uint64_t w = r->hist_stride < stride ? r->hist_stride : stride;
        if (keep && w)
            CU(cudaMemcpy2DAsync(nb, stride * sizeof(double), r->d_hist, r->hist_stride * sizeof(double), w * sizeof(double),
                                 (size_t)keep * 3, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(r->d_hist);
        if (r->hist_next > levels) r->hist_next = levels;
    }
    r->d_hist = nb; r->hist_levels = levels; r->hist_stride = stride;
    return OPM_OK;
}
// Ray::clear_pos_hist (ray.rs:225-227) for the whole bundle
static OpmStatus hist_clear(OpmRays* r) {
    if (r->d_hist && r->hist_next)
        CU(cudaMemsetAsync(r->d_hist, 0xFF, (size_t)r->hist_next * 3 * r->hist_stride * sizeof(double), r->ctx->stream));
    r->hist_next = 0;
    return OPM_OK;
}
// dst's history := the first `levels` levels of src's (a clone of the rays carries their history, ray.rs:625 / :752-772)
static OpmStatus hist_copy(const OpmRays* src, OpmRays* dst, uint32_t levels, uint64_t n) {
    OpmStatus s = OPM_OK;
    if (!dst->d_hist || dst->hist_levels < levels || dst->hist_stride < n) {
        dst->hist_next = 0;  // nothing worth carrying over
        s = hist_resize(dst, levels > dst->hist_levels ? levels : dst->hist_levels, n > dst->hist_stride ? n : dst->hist_stride);
    } else s = hist_clear(dst);
    if (s) return s;
    if (src && src->d_hist && levels && n) {
        if (levels > src->hist_next) levels = src->hist_next;
        if (levels)
            CU(cudaMemcpy2DAsync(dst->d_hist, dst->hist_stride * sizeof(double), src->d_hist, src->hist_stride * sizeof(double),
                                 n * sizeof(double), (size_t)levels * 3, cudaMemcpyDeviceToDevice, dst->ctx->stream));
    }
    dst->hist_next = (src && src->d_hist) ? levels : 0;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_enable_position_history(OpmRays* r, uint32_t levels_hint) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    CU(cudaSetDevice(r->ctx->device));
    if (r->d_hist) return OPM_OK;
    return hist_resize(r, levels_hint ? levels_hint : 16, r->cap);
}
extern "C" OpmStatus opm_rays_disable_position_history(OpmRays* r) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    if (!r->d_hist) return OPM_OK;
    CU(cudaSetDevice(r->ctx->device));
    CU(cudaStreamSynchronize(r->ctx->stream));
    cudaFree(r->d_hist);
    r->d_hist = nullptr; r->hist_levels = r->hist_next = 0; r->hist_stride = 0;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_clear_position_history(OpmRays* r) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    CU(cudaSetDevice(r->ctx->device));
    return hist_clear(r);
}
extern "C" int32_t opm_rays_position_history_levels(const OpmRays* r) { return (r && r->d_hist) ? (int32_t)r->hist_next : -1; }
// Ray::position_history / position_history_with_current (ray.rs:296-327) of every ray of the bundle, in the order of
// opm_rays_download: xyz_out[(ray * pos_cap + k) * 3 + c], len_out[ray] positions each
extern "C" OpmStatus opm_rays_position_history(const OpmRays* rc, int32_t with_current, double* xyz_out, uint32_t* len_out,
                                               uint64_t ray_cap, uint32_t pos_cap, uint64_t* n_out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !xyz_out || !len_out || !n_out) return fail(OPM_ERR_ARG, "NULL argument");
    if (!r->d_hist) return fail(OPM_ERR_ARG, "the bundle keeps no position history (opm_rays_enable_position_history)");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_materialize(r);
    if (!s) s = rays_refresh_len(r);
    if (s) return s;
    *n_out = 0;
    const uint64_t n = r->len;
    if (n == 0) return OPM_OK;
    if (pos_cap == 0) return fail(OPM_ERR_CAPACITY, "position capacity is 0");
    double* d_xyz = nullptr;
    uint32_t* d_len = nullptr;
    CU(cudaMalloc(&d_xyz, (size_t)n * pos_cap * 3 * sizeof(double)));
    CU(cudaMalloc(&d_len, (size_t)n * sizeof(uint32_t)));
    HistParams hp;
    hp.base = r->d_hist; hp.stride = r->hist_stride;
    int rc2 = k_history_gather(hp, r->hist_next, r->v, n, with_current ? 1 : 0, pos_cap, d_xyz, d_len, c->sms, c->stream);
    c->launches++;
    std::vector<double> h_xyz((size_t)n * pos_cap * 3);
    std::vector<uint32_t> h_len(n);
    cudaError_t e = rc2 ? (cudaError_t)rc2 : cudaMemcpyAsync(h_xyz.data(), d_xyz, h_xyz.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_len.data(), d_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_xyz); cudaFree(d_len);
    if (e != cudaSuccess) return fail(OPM_ERR_CUDA, "position history read-out failed: %s", cudaGetErrorString(e));
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (h_len[i] == 0xFFFFFFFFu) continue;  // removed slot (not part of the bundle)
        if (k >= ray_cap) return fail(OPM_ERR_CAPACITY, "position history buffer holds %llu rays, the bundle has more", (unsigned long long)ray_cap);
        if (h_len[i] > pos_cap) return fail(OPM_ERR_CAPACITY, "a ray has %u positions, the buffer holds %u per ray", h_len[i], pos_cap);
        memcpy(xyz_out + k * pos_cap * 3, h_xyz.data() + i * pos_cap * 3, (size_t)pos_cap * 3 * sizeof(double));
        len_out[k++] = h_len[i];
    }
    *n_out = k;
    return OPM_OK;
}

extern "C" OpmStatus opm_rays_upload(OpmRays* r, const OpmRay* host, uint64_t n) {
    if (!r || (!host && n)) return fail(OPM_ERR_ARG, "NULL argument");
    r->pending = false; r->mono_wvl = 0.0; r->spot_acc_valid = false;
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    if (r->d_hist) { OpmStatus hs = hist_clear(r); if (hs) return hs; }  // new rays start without a history (ray.rs:126)
    OpmStatus s = opm_rays_reserve(r, n);
    if (s) return s;
    if (n) {
        OpmRay* d_aos = nullptr;
        CU(cudaMalloc(&d_aos, n * sizeof(OpmRay)));
        CU(cudaMemcpyAsync(d_aos, host, n * sizeof(OpmRay), cudaMemcpyHostToDevice, c->stream));
        KL(k_aos_to_soa(d_aos, r->v, 0, n, c->sms, c->stream)); c->launches++;
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(d_aos);
    }
    return rays_set_len(r, n);
}
extern "C" OpmStatus opm_rays_upload_soa(OpmRays* r, const double* pos[3], const double* dir[3], const double* e,
                                         const double* wvl, uint64_t n) {
    if (!r || !pos || !dir || !e || !wvl) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    r->pending = false; r->mono_wvl = 0.0; r->spot_acc_valid = false;
    if (r->d_hist) { OpmStatus hs = hist_clear(r); if (hs) return hs; }  // new rays start without a history (ray.rs:126)
    OpmStatus s = opm_rays_reserve(r, n);
    if (s) return s;
    if (n) {
        double* dst[8] = {r->v.px, r->v.py, r->v.pz, r->v.dx, r->v.dy, r->v.dz, r->v.e, r->v.wvl};
        const double* src[8] = {pos[0], pos[1], pos[2], dir[0], dir[1], dir[2], e, wvl};
        for (int k = 0; k < 8; ++k) CU(cudaMemcpyAsync(dst[k], src[k], n * 8, cudaMemcpyHostToDevice, c->stream));
        KL(k_soa_defaults(r->v, n, c->sms, c->stream)); c->launches++;
    }
    return rays_set_len(r, n);
}
// Rays::new_collimated_with_spectrum (rays.rs:264-297): the host passes what the reference's position / energy /
// spectral distribution generators return (32 B per position + the spectrum table); the bundle is expanded on the device
extern "C" OpmStatus opm_rays_new_collimated_with_spectrum(OpmRays* r, const double* pos_xyz, const double* e_pos, uint64_t n_pos,
                                                            const double* wvl, const double* weights, uint32_t n_wvl) {
    if (!r || !pos_xyz || !e_pos || !wvl || !weights || n_wvl == 0) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    for (uint32_t k = 0; k < n_wvl; ++k)  // Ray::new (ray.rs:114-116)
        if (!(wvl[k] > 0.0) || !std::isfinite(wvl[k])) return fail(OPM_ERR_ARG, "wavelength must be >0");
    const uint64_t n_rays = n_pos * n_wvl;
    r->pending = false;  // an earlier upload that was never used is superseded
    r->mono_wvl = n_wvl == 1 ? wvl[0] : 0.0; r->spot_acc_valid = false;
    if (r->d_hist) { OpmStatus hs = hist_clear(r); if (hs) return hs; }  // new rays start without a history (ray.rs:126)
    OpmStatus s = opm_rays_reserve(r, n_rays);
    if (s) return s;
    const size_t need = align_up(n_pos * 32, 256) + (size_t)n_wvl * 16;
    if (need > r->stage_bytes) {
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(r->d_stage);
        r->d_stage = nullptr; r->stage_bytes = 0;
        CU(cudaMalloc(&r->d_stage, need));
        r->stage_bytes = need;
    }
    if (!r->ev_uploaded) {
        CU(cudaEventCreateWithFlags(&r->ev_uploaded, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&r->ev_expanded, cudaEventDisableTiming));
        CU(cudaEventRecord(r->ev_expanded, c->stream));
        CU(cudaEventRecord(r->ev_uploaded, c->copy_stream));
    }
    double* d_xyz = (double*)r->d_stage;
    double* d_e = d_xyz + 3 * n_pos;
    double* d_w = (double*)(r->d_stage + align_up(n_pos * 32, 256));
    // the staging buffer may still be read by the expansion of the previous upload; the bundle's arrays are only
    // touched on the main stream, which orders the expansion after everything that used them before
    CU(cudaEventSynchronize(r->ev_uploaded));  // the previous upload of THIS bundle has left the pinned spectrum table
    if ((size_t)n_wvl * 16 > r->h_spec_bytes) {
        cudaFreeHost(r->h_spec);
        r->h_spec = nullptr; r->h_spec_bytes = 0;
        CU(cudaMallocHost(&r->h_spec, (size_t)n_wvl * 16));
        r->h_spec_bytes = (size_t)n_wvl * 16;
    }
    memcpy(r->h_spec, wvl, (size_t)n_wvl * 8);
    memcpy(r->h_spec + n_wvl, weights, (size_t)n_wvl * 8);
    CU(cudaStreamWaitEvent(c->copy_stream, r->ev_expanded, 0));
    CU(cudaMemcpyAsync(d_xyz, pos_xyz, n_pos * 24, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d_e, e_pos, n_pos * 8, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d_w, r->h_spec, (size_t)n_wvl * 16, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(r->ev_uploaded, c->copy_stream));
    r->pending = true; r->pend_kind = 0; r->pend_n_pos = n_pos; r->pend_n_wvl = n_wvl;
    r->len = n_rays; r->len_dirty = false;  // the device-side length is set when the expansion is enqueued
    return OPM_OK;
}
static OpmStatus source_params(const OpmSourceDesc* d, SourceParams& S);
// The same bundle from the PositionDistribution's (x, y) alone (16 bytes per position over PCIe instead of 32): z = 0 for every
// position distribution of the reference, and the EnergyDistribution described by `desc` (energy_dist, total_energy, the Gaussian's
// parameters; n_positions = positions of the WHOLE source for Uniform; energy_norm = sum of the unnormalised Gaussian weights over
// the whole source, 0 = these positions are the whole source: sum them on the device) is evaluated on the device, like the
// spectrum and the source isometry of `desc`.  pos_dist / grid_* / size_* of `desc` are not used.
extern "C" OpmStatus opm_rays_new_collimated_xy(OpmRays* r, const double* pos_xy, uint64_t n_pos, const OpmSourceDesc* desc) {
    if (!r || !pos_xy || !desc || !desc->wavelengths || desc->n_wavelengths == 0) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    const uint32_t n_wvl = desc->n_wavelengths;
    for (uint32_t k = 0; k < n_wvl; ++k)  // Ray::new (ray.rs:114-116)
        if (!(desc->wavelengths[k] > 0.0) || !std::isfinite(desc->wavelengths[k])) return fail(OPM_ERR_ARG, "wavelength must be >0");
    OpmSourceDesc d2 = *desc;
    d2.pos_dist = OPM_POS_FIBONACCI_RECT;  // the generator is not used here: skip its consistency checks
    if (d2.n_positions == 0) d2.n_positions = n_pos;
    SourceParams S;
    OpmStatus s = source_params(&d2, S);
    if (s) return s;
    const uint64_t n_rays = n_pos * n_wvl;
    r->pending = false;
    r->mono_wvl = n_wvl == 1 ? desc->wavelengths[0] : 0.0; r->spot_acc_valid = false;
    if (r->d_hist) { OpmStatus hs = hist_clear(r); if (hs) return hs; }
    s = opm_rays_reserve(r, n_rays);
    if (s) return s;
    const size_t need = align_up(n_pos * 16 + 8, 256) + (size_t)n_wvl * 16;
    if (need > r->stage_bytes) {
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(r->d_stage);
        r->d_stage = nullptr; r->stage_bytes = 0;
        CU(cudaMalloc(&r->d_stage, need));
        r->stage_bytes = need;
    }
    if (!r->ev_uploaded) {
        CU(cudaEventCreateWithFlags(&r->ev_uploaded, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&r->ev_expanded, cudaEventDisableTiming));
        CU(cudaEventRecord(r->ev_expanded, c->stream));
        CU(cudaEventRecord(r->ev_uploaded, c->copy_stream));
    }
    double* d_xy = (double*)r->d_stage;
    double* d_spec = (double*)(r->d_stage + align_up(n_pos * 16 + 8, 256));
    CU(cudaEventSynchronize(r->ev_uploaded));
    if ((size_t)n_wvl * 16 > r->h_spec_bytes) {
        cudaFreeHost(r->h_spec);
        r->h_spec = nullptr; r->h_spec_bytes = 0;
        CU(cudaMallocHost(&r->h_spec, (size_t)n_wvl * 16));
        r->h_spec_bytes = (size_t)n_wvl * 16;
    }
    memcpy(r->h_spec, desc->wavelengths, (size_t)n_wvl * 8);
    if (desc->weights) memcpy(r->h_spec + n_wvl, desc->weights, (size_t)n_wvl * 8);
    else for (uint32_t k = 0; k < n_wvl; ++k) r->h_spec[n_wvl + k] = 1.0;
    CU(cudaStreamWaitEvent(c->copy_stream, r->ev_expanded, 0));
    CU(cudaMemcpyAsync(d_xy, pos_xy, n_pos * 16, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d_spec, r->h_spec, (size_t)n_wvl * 16, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(r->ev_uploaded, c->copy_stream));
    r->pending = true; r->pend_kind = 1; r->pend_src = S; r->pend_n_pos = n_pos; r->pend_n_wvl = n_wvl;
    r->len = n_rays; r->len_dirty = false;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_download(const OpmRays* rc, OpmRay* host, uint64_t capacity, uint64_t* n_out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !n_out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    uint64_t n = r->len;
    *n_out = 0;
    if (n == 0) return OPM_OK;
    OpmRay* d_aos = nullptr;
    CU(cudaMalloc(&d_aos, n * sizeof(OpmRay)));
    KL(k_soa_to_aos(r->v, d_aos, n, c->sms, c->stream)); c->launches++;
    std::vector<OpmRay> tmp(n);
    CU(cudaMemcpyAsync(tmp.data(), d_aos, n * sizeof(OpmRay), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_aos);
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (tmp[i].valid == RAY_REMOVED) continue;
        if (k >= capacity) return fail(OPM_ERR_CAPACITY, "download buffer too small");
        host[k++] = tmp[i];
    }
    *n_out = k;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_copy(const OpmRays* src, OpmRays* dst) {
    if (!src || !dst) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(dst->ctx->device));
    OpmStatus s = rays_refresh_len(const_cast<OpmRays*>(src));
    if (s) return s;
    s = opm_rays_reserve(dst, src->len);
    if (s) return s;
    s = copy_fields(src->v, 0, dst->v, 0, src->len, dst->ctx->stream);
    if (s) return s;
    if (src->d_hist) { s = hist_copy(src, dst, src->hist_next, src->len); if (s) return s; }
    else if (dst->d_hist) { s = hist_clear(dst); if (s) return s; }
    if (src->d_helpers) { s = helpers_copy(src, dst, src->len); if (s) return s; }  // Rays::clone clones the FluenceRays
    else if (dst->d_helpers) { CU(cudaStreamSynchronize(dst->ctx->stream)); cudaFree(dst->d_helpers); dst->d_helpers = nullptr; dst->helper_stride = 0; }
    dst->mono_wvl = src->mono_wvl; dst->spot_acc_valid = false;
    return rays_set_len(dst, src->len);
}
extern "C" OpmStatus opm_rays_append(OpmRays* dst, const OpmRays* src) {
    if (!src || !dst) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(dst->ctx->device));
    OpmStatus s = rays_refresh_len(const_cast<OpmRays*>(src));
    if (s) return s;
    s = rays_refresh_len(dst);
    if (s) return s;
    if (dst->d_hist) return fail(OPM_ERR_ARG, "a bundle with a position history cannot take appended rays");
    s = opm_rays_reserve(dst, dst->len + src->len);
    if (s) return s;
    s = copy_fields(src->v, 0, dst->v, dst->len, src->len, dst->ctx->stream);
    if (s) return s;
    if (dst->len == 0) dst->mono_wvl = src->mono_wvl; else if (dst->mono_wvl != src->mono_wvl) dst->mono_wvl = 0.0; dst->spot_acc_valid = false;
    return rays_set_len(dst, dst->len + src->len);
}
extern "C" OpmStatus opm_rays_compact(OpmRays* r) {
    if (!r) return fail(OPM_ERR_ARG, "rays is NULL");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    uint64_t n = r->len;
    if (n == 0) return OPM_OK;
    if (r->d_hist) return fail(OPM_ERR_ARG, "a bundle with a position history cannot be compacted (the log is indexed by slot)");
    uint64_t nb = (n + 1023) / 1024;
    unsigned* counts = nullptr; unsigned long long* offsets = nullptr;
    CU(cudaMalloc(&counts, nb * sizeof(unsigned)));
    CU(cudaMalloc(&offsets, (nb + 1) * sizeof(unsigned long long)));
    OpmRays old = *r;
    s = rays_alloc(r, old.cap);
    if (s) { *r = old; cudaFree(counts); cudaFree(offsets); return s; }
    int rc = k_compact(old.v, r->v, n, counts, offsets, r->d_tail, c->stream);
    c->launches += 3;
    unsigned long long total = 0;
    cudaMemcpyAsync(&total, r->d_tail, 8, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(counts); cudaFree(offsets); cudaFree(old.base);
    if (rc != 0 || e != cudaSuccess) return fail(OPM_ERR_CUDA, "compaction failed: %s", cudaGetErrorString(e));
    r->len = total; r->len_dirty = false;
    return OPM_OK;
}
extern "C" void* opm_rays_field_ptr(OpmRays* r, const char* f) {
    if (!r || !f) return nullptr;
    if (rays_refresh_len(r)) return nullptr;
    struct { const char* n; void* p; } tab[] = {
        {"px", r->v.px}, {"py", r->v.py}, {"pz", r->v.pz}, {"dx", r->v.dx}, {"dy", r->v.dy}, {"dz", r->v.dz},
        {"e", r->v.e}, {"wvl", r->v.wvl}, {"opl", r->v.opl}, {"n", r->v.n}, {"bounces", r->v.bounces},
        {"refr", r->v.refr}, {"id", r->v.id}, {"valid", r->v.valid}};
    for (auto& t : tab) if (!strcmp(t.n, f)) return t.p;
    return nullptr;
}

// ---------------------------------------------------------------------------------------------
// hit maps
// ---------------------------------------------------------------------------------------------
static size_t hits_layout(uint64_t cap, void* base, DevHits* v) {
    size_t off = 0;
    unsigned char* b = (unsigned char*)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? b + o : nullptr; };
    v->x = (double*)take(cap * 8); v->y = (double*)take(cap * 8); v->z = (double*)take(cap * 8); v->e = (double*)take(cap * 8);
    v->bounce = (uint32_t*)take(cap * 4);
    return off;  // the caller appends 64 bytes for the bounding-box accumulator
}
extern "C" OpmStatus opm_hitmap_create(OpmContext* ctx, uint64_t capacity, OpmHitMap** out) {
    if (!ctx || !out) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    OpmHitMap* h = new OpmHitMap();
    h->ctx = ctx; h->cap = capacity ? capacity : 1;
    DevHits tmp;
    size_t bytes = hits_layout(h->cap, nullptr, &tmp);
    cudaError_t e = cudaMalloc(&h->base, bytes + 64);
    if (e != cudaSuccess) { delete h; return fail(OPM_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    hits_layout(h->cap, h->base, &h->v);
    h->d_acc = (unsigned long long*)((unsigned char*)h->base + bytes);
    *out = h;
    return OPM_OK;
}
extern "C" void opm_hitmap_destroy(OpmHitMap* h) {
    if (!h) return;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    cudaFree(h->base);
    delete h;
}
extern "C" uint64_t opm_hitmap_capacity(const OpmHitMap* h) { return h ? h->cap : 0; }
extern "C" uint64_t opm_hitmap_len(const OpmHitMap* h) { return h ? h->len : 0; }
extern "C" OpmStatus opm_hitmap_upload(OpmHitMap* h, const double* x, const double* y, const double* z, const double* e,
                                       const uint32_t* bounce, uint64_t n) {
    if (!h || !x || !y || !e) return fail(OPM_ERR_ARG, "NULL argument");
    if (n > h->cap) return fail(OPM_ERR_CAPACITY, "hit map capacity %llu < %llu", (unsigned long long)h->cap, (unsigned long long)n);
    h->acc_valid = false;
    cudaStream_t st = h->ctx->stream;
    CU(cudaSetDevice(h->ctx->device));
    CU(cudaMemcpyAsync(h->v.x, x, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->v.y, y, n * 8, cudaMemcpyHostToDevice, st));
    if (z) CU(cudaMemcpyAsync(h->v.z, z, n * 8, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(h->v.z, 0, n * 8, st));
    CU(cudaMemcpyAsync(h->v.e, e, n * 8, cudaMemcpyHostToDevice, st));
    if (bounce) CU(cudaMemcpyAsync(h->v.bounce, bounce, n * 4, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(h->v.bounce, 0, n * 4, st));
    CU(cudaStreamSynchronize(st));
    h->len = n;
    return OPM_OK;
}
extern "C" OpmStatus opm_hitmap_download(const OpmHitMap* h, double* x, double* y, double* z, double* e, uint32_t* bounce,
                                         uint64_t capacity, uint64_t* n_out) {
    if (!h || !n_out) return fail(OPM_ERR_ARG, "NULL argument");
    cudaStream_t st = h->ctx->stream;
    CU(cudaSetDevice(h->ctx->device));
    uint64_t n = h->len;
    std::vector<double> tx(n), ty(n), tz(n), te(n);
    std::vector<uint32_t> tb(n);
    if (n) {
        CU(cudaMemcpyAsync(tx.data(), h->v.x, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ty.data(), h->v.y, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(tz.data(), h->v.z, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(te.data(), h->v.e, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(tb.data(), h->v.bounce, n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (!(te[i] >= 0.0)) continue;
        if (k >= capacity) return fail(OPM_ERR_CAPACITY, "download buffer too small");
        if (x) x[k] = tx[i];
        if (y) y[k] = ty[i];
        if (z) z[k] = tz[i];
        if (e) e[k] = te[i];
        if (bounce) bounce[k] = tb[i];
        ++k;
    }
    *n_out = k;
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// program compilation: public op descriptors -> device ops
// ---------------------------------------------------------------------------------------------
static void to_dev_iso(const OpmIso& s, DevIso& d) { memcpy(d.r, s.rot, sizeof d.r); memcpy(d.t, s.t, sizeof d.t); }

struct Compiled {
    std::vector<DevOp> ops;
    std::vector<void*> temp_dev;        // polygon triangle uploads, freed after the launch
    std::vector<OpmRays*> appended;     // bundles that receive appended rays (children)
    std::vector<OpmRays*> slot_outputs; // split / snapshot outputs (len = n)
    std::vector<std::pair<int, OpmRays*>> slot_output_ops;  // the same with the index of their device op (position history)
    std::vector<OpmHitMap*> hitmaps;
    std::vector<std::pair<int, OpmHitMap*>> hitmap_ops;  // the same with the index of the device op that records into them
    std::vector<std::pair<int, OpmRays*>> spot_ops;      // snapshot ops that want the phase-1 spot sums of their copy
    int fork_depth = 0;                 // inside an OPM_OP_FORK .. OPM_OP_JOIN range
    bool has_fork = false;
};

static OpmStatus compile_aperture(OpmContext* c, const OpmAperture& a, DevAperture& d, Compiled& out) {
    memset(&d, 0, sizeof d);
    if (a.n_items < 0 || a.n_items > OPM_MAX_APERTURE_ITEMS) return fail(OPM_ERR_ARG, "aperture n_items out of range");
    d.n_items = a.n_items; d.is_stack = a.is_stack; d.stack_obstruction = a.stack_obstruction; d.n_tris = 0;
    for (int k = 0; k < a.n_items; ++k) {
        d.items[k].type = a.items[k].type; d.items[k].obstruction = a.items[k].obstruction;
        memcpy(d.items[k].p, a.items[k].p, sizeof d.items[k].p);
        if (a.items[k].type == OPM_AP_POLYGON) {
            if (a.n_tris <= 0 || !a.tris) return fail(OPM_ERR_ARG, "polygon aperture without triangles");
            double* dt = nullptr;
            CU(cudaMalloc(&dt, (size_t)a.n_tris * 6 * sizeof(double)));
            out.temp_dev.push_back(dt);
            CU(cudaMemcpyAsync(dt, a.tris, (size_t)a.n_tris * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            d.tris = dt; d.n_tris = a.n_tris;
        }
    }
    return OPM_OK;
}

static DevSpectrum dev_spectrum(const OpmSpectrum* sp) {
    DevSpectrum d;
    d.lam = sp->d; d.val = sp->d + sp->n; d.n = sp->n;
    return d;
}

static OpmStatus compile_ops(OpmContext* c, const OpmOp* ops, int n_ops, uint64_t n_rays, Compiled& out) {
    if (n_ops < 0 || n_ops > OPM_MAX_PROGRAM_OPS) return fail(OPM_ERR_ARG, "program has %d ops (max %d)", n_ops, OPM_MAX_PROGRAM_OPS);
    for (int k = 0; k < n_ops; ++k) {
        const OpmOp& p = ops[k];
        DevOp d;
        memset(&d, 0, sizeof d);
        switch (p.kind) {
            case OPM_OP_REFRACT: {
                const OpmRefractOp& r = p.u.refract;
                DevRefract& o = d.u.refract;
                d.kind = OPK_REFRACT;
                if (r.surface.geom < OPM_GEOM_PLANE || r.surface.geom > OPM_GEOM_PARABOLA) return fail(OPM_ERR_ARG, "unknown surface geometry %d", r.surface.geom);
                if (r.surface.geom != OPM_GEOM_PLANE) {  // sphere.rs:48-57, cylinder.rs:30-37, parabola.rs:31-42: is_normal
                    if (!std::isnormal(r.surface.geom_param)) return fail(OPM_ERR_ARG, "radius of curvature / focal length must be != 0.0 and finite");
                }
                o.geom = r.surface.geom; o.coating = r.surface.coating; o.param = r.surface.geom_param; o.coating_r = r.surface.coating_r;
                o.inv_abs_param = r.surface.geom != OPM_GEOM_PLANE ? 1.0 / std::fabs(r.surface.geom_param) : 0.0;
                to_dev_iso(r.surface.iso.fwd, o.fwd); to_dev_iso(r.surface.iso.inv, o.inv);
                o.has_index = r.has_index ? 1 : 0; o.refraction_intended = r.refraction_intended ? 1 : 0;
                o.strategy = r.strategy; o.flags = r.flags & 0xffff; o.index_type = r.index.type;
                memcpy(o.ic, r.index.c, sizeof o.ic); o.wlo = r.index.wvl_lo; o.whi = r.index.wvl_hi;
                if (r.hits) {
                    if (r.hits->cap < n_rays) return fail(OPM_ERR_CAPACITY, "hit map capacity %llu < %llu rays", (unsigned long long)r.hits->cap, (unsigned long long)n_rays);
                    o.hits = r.hits->v;
                    out.hitmaps.push_back(r.hits);
                    out.hitmap_ops.push_back({(int)out.ops.size(), r.hits});
                }
                if (r.flags & OPM_REFRACT_EMIT_CHILDREN) {
                    if (!r.children) return fail(OPM_ERR_ARG, "OPM_REFRACT_EMIT_CHILDREN without a children bundle");
                    if (r.children->sealed) return fail(OPM_ERR_ARG, "the children bundle took its appends under opm_rays_expect_appends: clear it first");
                    OpmStatus s = rays_refresh_len(r.children);
                    if (s) return s;
                    s = opm_rays_reserve(r.children, r.children->len + n_rays);
                    if (s) return s;
                    o.children = r.children->v; o.child_tail = r.children->d_tail; o.child_cap = r.children->cap;
                    if (r.children->expected_len) {  // opm_rays_expect_appends: child of slot i -> slot i, the extent is this launch's
                        if (r.children->len != 0) return fail(OPM_ERR_ARG, "opm_rays_expect_appends needs an empty children bundle");
                        o.flags |= kRefractChildrenBySlot;
                        r.children->expected_len = n_rays ? n_rays : 1;
                    }
                    out.appended.push_back(r.children);
                }
                break;
            }
            case OPM_OP_APODIZE: {
                d.kind = OPK_APODIZE;
                to_dev_iso(p.u.apodize.iso.inv, d.u.apodize.inv);
                OpmStatus s = compile_aperture(c, p.u.apodize.aperture, d.u.apodize.ap, out);
                if (s) return s;
                if (d.u.apodize.ap.n_items == 0) continue;  // Aperture::None: factor 1.0 for every ray
                break;
            }
            case OPM_OP_THRESHOLD: {  // rays.rs:1145-1154
                double t = p.u.threshold.min_energy;
                if (std::signbit(t)) continue;  // "negative threshold energy given. Ray bundle unmodified."
                if (!std::isfinite(t)) return fail(OPM_ERR_THRESHOLD, "%s", opm_status_string(OPM_ERR_THRESHOLD));
                if (!out.ops.empty() && out.ops.back().epi_flags == 0) {  // fold into the epilogue of the op it follows
                    out.ops.back().epi_flags = EPI_THRESHOLD; out.ops.back().epi_min_energy = t;
                    continue;
                }
                d.kind = OPK_THRESHOLD; d.u.threshold.min_energy = t;
                break;
            }
            case OPM_OP_LIMITS:
                if (!out.ops.empty() && !(out.ops.back().epi_flags & EPI_LIMITS)) {  // the epilogue runs threshold, then limits
                    DevOp& b = out.ops.back();
                    b.epi_flags |= EPI_LIMITS | ((p.u.limits.check_refractions && p.u.limits.max_refractions <= 0xffffffffull) ? EPI_CHECK_REFR : 0);
                    b.epi_max_bounces = limit_sat32(p.u.limits.max_bounces); b.epi_max_refr = limit_sat32(p.u.limits.max_refractions);
                    continue;
                }
                d.kind = OPK_LIMITS; d.u.limits.max_bounces = limit_sat32(p.u.limits.max_bounces); d.u.limits.max_refr = limit_sat32(p.u.limits.max_refractions);
                d.u.limits.check_refr = (p.u.limits.check_refractions && p.u.limits.max_refractions <= 0xffffffffull) ? 1 : 0;
                break;
            case OPM_OP_FILTER: {  // rays.rs:1120-1126
                double t = p.u.filter.transmission;
                d.kind = OPK_FILTER;
                if (p.u.filter.spectrum) {  // FilterType::Spectrum: the transmission is looked up per ray (ray.rs:712-721)
                    if (p.u.filter.spectrum->ctx != c) return fail(OPM_ERR_ARG, "spectrum belongs to another context");
                    d.u.filter.s = dev_spectrum(p.u.filter.spectrum);
                    d.u.filter.t = 1.0;
                    break;
                }
                if (!(t >= 0.0 && t <= 1.0)) return fail(OPM_ERR_APODIZE_FACTOR, "transmission value must be in the range [0.0;1.0]");
                d.u.filter.t = t;
                break;
            }
            case OPM_OP_PARAXIAL: {  // rays.rs:944-948
                double f = p.u.paraxial.focal_length;
                if (f == 0.0 || !std::isfinite(f)) return fail(OPM_ERR_FOCAL_LENGTH, "%s", opm_status_string(OPM_ERR_FOCAL_LENGTH));
                d.kind = OPK_PARAXIAL; d.u.paraxial.f = f;
                to_dev_iso(p.u.paraxial.iso.fwd, d.u.paraxial.fwd); to_dev_iso(p.u.paraxial.iso.inv, d.u.paraxial.inv);
                break;
            }
            case OPM_OP_SPLIT: {  // ray.rs:763-767
                double ratio = p.u.split.ratio;
                if (p.u.split.spectrum) {  // SplittingConfig::Spectrum (ray.rs:755-761)
                    if (p.u.split.spectrum->ctx != c) return fail(OPM_ERR_ARG, "spectrum belongs to another context");
                    d.u.split.s = dev_spectrum(p.u.split.spectrum);
                    ratio = 1.0;
                }
                if (!(ratio >= 0.0 && ratio <= 1.0)) return fail(OPM_ERR_SPLIT_RATIO, "%s", opm_status_string(OPM_ERR_SPLIT_RATIO));
                d.kind = OPK_SPLIT; d.u.split.ratio = ratio;
                if (p.u.split.split_out) {
                    OpmStatus s = opm_rays_reserve(p.u.split.split_out, n_rays);
                    if (s) return s;
                    d.u.split.out = p.u.split.split_out->v;
                    out.slot_outputs.push_back(p.u.split.split_out);
                    out.slot_output_ops.push_back({(int)out.ops.size(), p.u.split.split_out});
                }
                break;
            }
            case OPM_OP_FORK: {  // ray.rs:763-767
                double ratio = p.u.fork.ratio;
                if (p.u.fork.spectrum) {
                    if (p.u.fork.spectrum->ctx != c) return fail(OPM_ERR_ARG, "spectrum belongs to another context");
                    d.u.fork.s = dev_spectrum(p.u.fork.spectrum);
                    ratio = 1.0;
                }
                if (!(ratio >= 0.0 && ratio <= 1.0)) return fail(OPM_ERR_SPLIT_RATIO, "%s", opm_status_string(OPM_ERR_SPLIT_RATIO));
                if (out.fork_depth != 0) return fail(OPM_ERR_ARG, "OPM_OP_FORK inside a fork: one level only");
                out.fork_depth = 1; out.has_fork = true;
                d.kind = OPK_FORK; d.u.fork.ratio = ratio;
                break;
            }
            case OPM_OP_JOIN: {
                if (out.fork_depth != 1) return fail(OPM_ERR_ARG, "OPM_OP_JOIN without a matching OPM_OP_FORK");
                out.fork_depth = 0;
                d.kind = OPK_JOIN;
                if (p.u.join.out) {
                    OpmStatus s = opm_rays_reserve(p.u.join.out, n_rays);
                    if (s) return s;
                    d.u.join.out = p.u.join.out->v;
                    out.slot_outputs.push_back(p.u.join.out);
                }
                break;
            }
            case OPM_OP_SNAPSHOT: {
                if (!p.u.snapshot.out) return fail(OPM_ERR_ARG, "snapshot without output bundle");
                OpmStatus s = opm_rays_reserve(p.u.snapshot.out, n_rays);
                if (s) return s;
                if (p.u.snapshot.fields != OPM_SNAPSHOT_ALL && p.u.snapshot.fields != OPM_SNAPSHOT_SPOT)
                    return fail(OPM_ERR_ARG, "unknown snapshot field set %d", p.u.snapshot.fields);
                d.kind = OPK_SNAPSHOT; d.u.snapshot.out = p.u.snapshot.out->v;
                d.u.snapshot.fields = p.u.snapshot.fields;
                d.u.snapshot.has_frame = p.u.snapshot.spot_frame ? 1 : 0;
                if (p.u.snapshot.spot_frame) to_dev_iso(p.u.snapshot.spot_frame->inv, d.u.snapshot.inv);
                d.u.snapshot.acc = nullptr;  // set by the launcher when the specialised kernel runs
                if (p.u.snapshot.want_spot_sums) out.spot_ops.push_back({(int)out.ops.size(), p.u.snapshot.out});
                out.slot_outputs.push_back(p.u.snapshot.out);
                out.slot_output_ops.push_back({(int)out.ops.size(), p.u.snapshot.out});
                break;
            }
            case OPM_OP_DIFFRACT: {
                const OpmDiffractOp& g = p.u.diffract;
                DevDiffract& o = d.u.diffract;
                d.kind = OPK_DIFFRACT;
                if (g.surface.geom < OPM_GEOM_PLANE || g.surface.geom > OPM_GEOM_PARABOLA) return fail(OPM_ERR_ARG, "unknown surface geometry %d", g.surface.geom);
                if (g.surface.geom != OPM_GEOM_PLANE && !std::isnormal(g.surface.geom_param)) return fail(OPM_ERR_ARG, "radius of curvature / focal length must be != 0.0 and finite");
                o.surf.geom = g.surface.geom; o.surf.coating = g.surface.coating; o.surf.param = g.surface.geom_param; o.surf.coating_r = g.surface.coating_r;
                o.surf.inv_abs_param = g.surface.geom != OPM_GEOM_PLANE ? 1.0 / std::fabs(g.surface.geom_param) : 0.0;
                to_dev_iso(g.surface.iso.fwd, o.surf.fwd); to_dev_iso(g.surface.iso.inv, o.surf.inv);
                o.surf.has_index = 1; o.surf.refraction_intended = g.refraction_intended ? 1 : 0; o.surf.strategy = OPM_MISS_IGNORE;
                o.surf.flags = g.flags; o.surf.index_type = g.index.type;
                memcpy(o.surf.ic, g.index.c, sizeof o.surf.ic); o.surf.wlo = g.index.wvl_lo; o.surf.whi = g.index.wvl_hi;
                for (int j = 0; j < 3; ++j) o.gvec[j] = g.grating_vector[j];
                o.gnorm = std::sqrt(0.0 + (g.grating_vector[0] * g.grating_vector[0] + g.grating_vector[1] * g.grating_vector[1]) + g.grating_vector[2] * g.grating_vector[2]);
                o.order = (double)g.order;
                if (g.flags & OPM_REFRACT_EMIT_CHILDREN) {
                    if (!g.children) return fail(OPM_ERR_ARG, "OPM_REFRACT_EMIT_CHILDREN without a children bundle");
                    OpmStatus s = rays_refresh_len(g.children);
                    if (s) return s;
                    s = opm_rays_reserve(g.children, g.children->len + n_rays);
                    if (s) return s;
                    if (g.children->expected_len || g.children->sealed) return fail(OPM_ERR_ARG, "opm_rays_expect_appends is for OPM_OP_REFRACT children");
                    o.surf.children = g.children->v; o.surf.child_tail = g.children->d_tail; o.surf.child_cap = g.children->cap;
                    out.appended.push_back(g.children);
                }
                break;
            }
            case OPM_OP_TRANSFORM: {
                d.kind = OPK_TRANSFORM;
                to_dev_iso(p.u.transform.inverse ? p.u.transform.iso.inv : p.u.transform.iso.fwd, d.u.transform.iso);
                const DevIso& I = d.u.transform.iso;  // the identity moves nothing (rotation by the unit matrix and adding 0.0 are exact)
                const bool ident = I.r[0] == 1.0 && I.r[4] == 1.0 && I.r[8] == 1.0 && I.r[1] == 0.0 && I.r[2] == 0.0 && I.r[3] == 0.0 &&
                                   I.r[5] == 0.0 && I.r[6] == 0.0 && I.r[7] == 0.0 && I.t[0] == 0.0 && I.t[1] == 0.0 && I.t[2] == 0.0;
                if (ident) continue;
                break;
            }
            default: return fail(OPM_ERR_ARG, "unknown op kind %d", p.kind);
        }
        out.ops.push_back(d);
    }
    if (out.fork_depth != 0) return fail(OPM_ERR_ARG, "OPM_OP_FORK without a matching OPM_OP_JOIN in the same program");
    return OPM_OK;
}

static void stats_to_public(const DevStats& d, OpmTraceStats* s) {
    s->interactions = d.interactions; s->children = d.children; s->hits = d.hits; s->missed = d.missed;
    s->apodized = d.apodized; s->valid_out = d.valid_out; s->alive_out = d.alive_out; s->status_bits = d.status_bits; s->_pad = 0;
    s->first_valid_key = d.first_inv ? ~d.first_inv : ~0ull;
}

// The one place that launches the fused kernel.  stats == NULL: fully asynchronous (no host sync).
static OpmStatus run_program(OpmRays* rays, const OpmOp* ops, int n_ops, int flags, OpmRays* out, OpmTraceStats* stats) {
    if (!rays || (!ops && n_ops)) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = rays->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(rays);
    if (s) return s;
    const uint64_t n = rays->len;
    const bool compact = (flags & OPM_TRACE_COMPACT) != 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (compact && (!out || out == rays)) return fail(OPM_ERR_ARG, "OPM_TRACE_COMPACT needs a distinct output bundle");
    OpmRays* dst = (out && out != rays) ? out : rays;

    Compiled prog;
    s = compile_ops(c, ops, n_ops, n, prog);
    if (!s && rays->mono_wvl > 0.0)  // one-wavelength bundle: the specialised kernels evaluate each index model once per CTA
        for (DevOp& d : prog.ops)
            if (d.kind == OPK_REFRACT && d.u.refract.has_index) d.u.refract.hint_wvl = rays->mono_wvl;
    auto free_temps = [&]() { for (void* p : prog.temp_dev) cudaFree(p); };
    if (s) { free_temps(); return s; }

    if (compact) {
        s = rays_refresh_len(dst);
        if (!s) s = opm_rays_reserve(dst, dst->len + n);
        if (s) { free_temps(); return s; }
    } else if (dst != rays) {
        s = opm_rays_reserve(dst, n);
        if (s) { free_temps(); return s; }
    }
    // fluence helper rays: the interpreter variant that carries them through refractions and paraxial surfaces
    HelperParams help{};
    const bool use_helpers = rays->d_helpers != nullptr;
    if (use_helpers) {
        auto no = [&](const char* what) { free_temps(); return fail(OPM_ERR_ARG, "a bundle with fluence helper rays cannot %s", what); };
        if (compact) return no("be compacted by a trace");
        if (rays->d_hist || dst->d_hist) return no("keep a position history");
        if (n > rays->helper_stride) return no("have grown since the helpers were attached");
        for (const DevOp& d : prog.ops) {
            if (d.kind == OPK_SPLIT || d.kind == OPK_FORK || d.kind == OPK_JOIN) return no("be split in a fused program (copy the bundle and use opm_rays_split)");
            if (d.kind == OPK_DIFFRACT) return no("be diffracted (the reference moves no helper ray there)");
            if (d.kind == OPK_REFRACT && d.u.refract.children.px) return no("emit reflected children");
        }
        if (dst != rays) { s = helpers_alloc(dst, n); if (s) { free_temps(); return s; } }
        help.in = rays->d_helpers; help.in_stride = rays->helper_stride;
        help.out = dst->d_helpers; help.out_stride = dst->helper_stride;
    } else if (dst != rays && dst->d_helpers) {
        cudaStreamSynchronize(c->stream); cudaFree(dst->d_helpers); dst->d_helpers = nullptr; dst->helper_stride = 0;
    }
    // position history: the bundle the program leaves its rays in logs one level per pushing op
    HistParams hp;
    const bool use_hist = rays->d_hist != nullptr || dst->d_hist != nullptr;
    std::vector<uint32_t> level_before;  // per device op: levels in use when it starts
    if (use_hist) {
        if (compact) { free_temps(); return fail(OPM_ERR_ARG, "a bundle with a position history cannot be compacted by a trace"); }
        if (prog.has_fork) { free_temps(); return fail(OPM_ERR_ARG, "OPM_OP_FORK needs bundles without a position history (use OPM_OP_SPLIT)"); }
        if (dst != rays) { s = hist_copy(rays, dst, rays->d_hist ? rays->hist_next : 0, n); if (s) { free_temps(); return s; } }
        uint32_t need = dst->hist_next;
        for (const DevOp& d : prog.ops) need += (d.kind == OPK_REFRACT || d.kind == OPK_DIFFRACT || d.kind == OPK_APODIZE) ? 1u : 0u;
        if (need > dst->hist_levels || dst->hist_stride < n) {
            s = hist_resize(dst, need > 2 * dst->hist_levels ? need : 2 * dst->hist_levels, n);
            if (s) { free_temps(); return s; }
        }
        uint32_t lv = dst->hist_next;
        for (size_t k = 0; k < prog.ops.size(); ++k) {
            const int kind = prog.ops[k].kind;
            level_before.push_back(lv);
            hp.level[k] = (kind == OPK_REFRACT || kind == OPK_DIFFRACT || kind == OPK_APODIZE) ? (short)lv++ : (short)-1;
        }
        hp.base = dst->d_hist; hp.stride = dst->hist_stride;
        dst->hist_next = lv;
    }
    std::vector<OpmRays*> planned;  // bundles whose length the prologue launch sets
    auto is_planned = [&](OpmRays* r) { for (OpmRays* q : planned) if (q == r) return true; return false; };
    // a program whose ops were all elided (Aperture::None, negative threshold) still has to honour out/compact
    if (n > 0 && (!prog.ops.empty() || dst != rays)) {
        if (prog.ops.empty()) {  // pure copy: keep the kernel path uniform with a no-op limits check
            DevOp d; memset(&d, 0, sizeof d);
            d.kind = OPK_LIMITS; d.u.limits.max_bounces = 0xffffffffu; d.u.limits.max_refr = 0xffffffffu; d.u.limits.check_refr = 0;
            prog.ops.push_back(d);
        }
        const int slot = c->ring_pos;
        c->ring_pos = (c->ring_pos + 1) % OpmContext::kRing;
        DevOp* h_slot = c->h_ops_ring + (size_t)slot * OPM_MAX_PROGRAM_OPS;
        DevOp* d_slot = c->d_ops + (size_t)slot * OPM_MAX_PROGRAM_OPS;
        CU(cudaEventSynchronize(c->ring_ev[slot]));  // the copy that last used this pinned slot has drained
        const bool strict = (flags & OPM_TRACE_STRICT) != 0;
        // hit maps that exactly one op of this launch records into get their bounding box from the specialised kernel (if that is
        // the kernel that runs: the accumulator pointer is a structural fact of its source)
        const bool jit_candidate = !strict && !use_hist && !use_helpers && c->jit_enabled && n >= c->jit_min_rays && prog.ops.size() * sizeof(DevOp) <= 60 * 1024;
        static const bool fuse_bbox = [] { const char* e = getenv("OPM_B200_JIT_BBOX"); return !(e && e[0] == '0'); }();  // tuning switch
        const int jit_block = c->jit_block ? c->jit_block : (prog.has_fork ? 896 : 1024);
        std::vector<OpmHitMap*> fused;
        if (jit_candidate && fuse_bbox)
            for (auto& ho : prog.hitmap_ops) {
                int uses = 0;
                for (auto& other : prog.hitmap_ops) uses += other.second == ho.second ? 1 : 0;
                if (uses != 1 || prog.ops[ho.first].kind != OPK_REFRACT) continue;
                if (fused.size() == 2) break;  // 36 bytes of shared memory per thread and map, next to the parked rays of a FORK
                prog.ops[ho.first].u.refract.hit_acc = ho.second->d_acc;
                fused.push_back(ho.second);
            }
        // snapshot copies whose phase-1 spot sums the specialised kernel takes on the way (kSpotSmemBytes of shared memory per thread
        // each, next to the parked rays of a FORK and the running boxes)
        // Off unless OPM_B200_JIT_SPOT=1: measured on C2 (profiles/r2_summary.md) the sums cost the trace kernel 0.056 ms (0.12 ms when the
        // running boxes already fill the shared memory and the L1 shrinks to its minimum) against the 0.085 ms pass they replace
        const bool fuse_spot = c->jit_fuse_spot;  // read from the environment when the context was created
        std::vector<std::pair<int, OpmRays*>> spot_fused;
        if (jit_candidate && fuse_spot) {
            size_t smem_need = (prog.has_fork ? (size_t)jit_block * kStashWords * 4 + 128 : 0) + fused.size() * (size_t)jit_block * 36;
            for (auto& so : prog.spot_ops) {
                if (spot_fused.size() == 2 || smem_need + (size_t)jit_block * kSpotSmemBytes > 222 * 1024) break;  // 227 KB per CTA, ~5 KB of it static
                bool twice = false;
                for (auto& q : spot_fused) twice |= q.second == so.second;
                if (twice) continue;
                smem_need += (size_t)jit_block * kSpotSmemBytes;
                prog.ops[so.first].u.snapshot.acc = so.second->d_spot_acc;
                spot_fused.push_back(so);
            }
        }
        memcpy(h_slot, prog.ops.data(), prog.ops.size() * sizeof(DevOp));
        int block = strict ? trace_block_strict() : trace_block_fast();
        // large bundles run the kernel compiled for the structure of this program (opm_jit.cu); otherwise, and whenever
        // that kernel cannot be had, the interpreter
        const JitKernel* jk = nullptr;
        if (jit_candidate) {  // __constant__ holds 64 KB
            const char* why = nullptr;
            jk = jit_get(c->jit, prog.ops.data(), (int)prog.ops.size(), compact, jit_block, c->jit_min_ctas, c->jit_sync, &why);
            if (jk) block = jit_block;
            if (!jk && why && !c->jit_warned) {  // why == NULL: still compiling, interpret meanwhile
                c->jit_warned = true;
                fprintf(stderr, "opossum_b200: run-time kernel specialisation unavailable, using the interpreter kernel (%s)\n", why ? why : "?");
            }
        }
        // one prologue launch: descriptors to the device, statistics zeroed, bounding-box accumulators reset, output lengths set
        ProloguePlan pp;
        memset(&pp, 0, sizeof pp);
        pp.dst = (unsigned long long*)(jk ? jit_const_ptr(jk) : (void*)d_slot);
        pp.src_host = (const unsigned long long*)h_slot;
        pp.words = (unsigned)(prog.ops.size() * sizeof(DevOp) / 8);
        pp.stats = (unsigned long long*)c->d_stats; pp.stats_words = (unsigned)(kDevStatsZeroBytes / 8);  // not the sticky status word
        for (int k = 0; k < 2; ++k) { pp.key_dst[k] = c->pending_key[k]; c->pending_key[k] = nullptr; }
        for (OpmHitMap* h : prog.hitmaps) h->acc_valid = false;
        if (jk)
            for (size_t k = 0; k < fused.size(); ++k) { pp.acc[k] = (BBoxAccum*)fused[k]->d_acc; fused[k]->acc_valid = true; }
        for (OpmRays* r : prog.slot_outputs) r->spot_acc_valid = false;
        rays->spot_acc_valid = false; dst->spot_acc_valid = false;
        if (jk)
            for (size_t k = 0; k < spot_fused.size(); ++k) {  // reset by the prologue launch, filled by the kernel
                OpmRays* r = spot_fused[k].second;
                const DevSnapshot& sn = prog.ops[spot_fused[k].first].u.snapshot;
                pp.spot[k] = r->d_spot_acc;
                r->spot_acc_valid = true; r->spot_has_frame = sn.has_frame != 0; r->spot_inv = sn.inv;
            }
        // slot-indexed outputs keep the input's length; the kernel does not touch their tail counters (only compacting stores and
        // child emission do), so the prologue can set them
        auto plan_len = [&](OpmRays* r) {
            if (pp.n_tails < kPrologueTails) { pp.tail[pp.n_tails] = r->d_tail; pp.tail_val[pp.n_tails] = n; pp.n_tails++; r->len = n; r->len_dirty = false; return true; }
            return false;
        };
        for (OpmRays* r : prog.slot_outputs) if (plan_len(r)) planned.push_back(r);
        if (!compact && dst != rays && plan_len(dst)) planned.push_back(dst);
        KL(k_trace_prologue(pp, c->stream));
        c->launches++;
        TraceLaunch L;
        L.ops = d_slot; L.n_ops = (int)prog.ops.size();
        L.in = rays->v; L.out = dst->v; L.n = n; L.out_cap = dst->cap; L.out_tail = dst->d_tail;
        L.stats = c->d_stats; L.compact = compact ? 1 : 0; L.block = block; L.stream = c->stream;
        L.has_fork = prog.has_fork ? 1 : 0;
        L.hist = use_hist ? &hp : nullptr;
        L.helpers = use_helpers ? &help : nullptr;
        size_t smem = prog.ops.size() * sizeof(DevOp) + (prog.has_fork ? (size_t)block * kStashWords * 4 : 0);
        int occ = jk ? jit_occupancy(jk) : strict ? trace_occupancy_strict(L.block, smem) : trace_occupancy_fast(L.block, smem);
        uint64_t tiles = (n + L.block - 1) / L.block;
        uint64_t maxg = (uint64_t)c->sms * occ;
        L.grid = (int)(tiles < maxg ? tiles : maxg);
        std::pair<cudaEvent_t, cudaEvent_t> tev{nullptr, nullptr};
        if (c->timing) {
            if (!c->timing_free.empty()) { tev = c->timing_free.back(); c->timing_free.pop_back(); }
            else { CU(cudaEventCreate(&tev.first)); CU(cudaEventCreate(&tev.second)); }
            CU(cudaEventRecord(tev.first, c->stream));
        }
        int rc = jk ? jit_launch(c->jit, jk, L, nullptr) : strict ? launch_trace_strict(L) : launch_trace_fast(L);
        c->launches += 1;
        CU(cudaEventRecord(c->ring_ev[slot], c->stream));  // the pinned slot is free once the fetch kernel has read it
        if (c->timing) { CU(cudaEventRecord(tev.second, c->stream)); c->timing_events.push_back(tev); }
        if (rc != 0) { free_temps(); return fail(OPM_ERR_CUDA, "trace kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc)); }
    }
    // a split-off or stored copy of the rays carries the history they had at that op (clone, ray.rs:752-772)
    if (use_hist)
        for (auto& so : prog.slot_output_ops) {
            const uint32_t lv = (size_t)so.first < level_before.size() ? level_before[so.first] : dst->hist_next;
            s = hist_copy(dst, so.second, lv, n);
            if (s) { free_temps(); return s; }
        }
    // host-side lengths
    for (OpmHitMap* h : prog.hitmaps) h->len = n;
    for (OpmRays* r : prog.slot_outputs) {
        if (!is_planned(r)) { s = rays_set_len(r, n); if (s) { free_temps(); return s; } }
        r->mono_wvl = rays->mono_wvl;
    }
    for (OpmRays* r : prog.appended) {  // children carry their parents' wavelengths
        if (r->len == 0 && !r->len_dirty) r->mono_wvl = rays->mono_wvl; else if (r->mono_wvl != rays->mono_wvl) r->mono_wvl = 0.0; r->spot_acc_valid = false;
        if (r->expected_len) {  // opm_rays_expect_appends: the extent is known, unfilled slots are removed rays -- no read-back, ever
            r->len = r->expected_len; r->expected_len = 0; r->len_dirty = false; r->sealed = true;
        } else r->len_dirty = true;
    }
    if (dst != rays) dst->mono_wvl = rays->mono_wvl;
    if (compact) dst->len_dirty = true;
    else if (dst != rays && !is_planned(dst)) { s = rays_set_len(dst, n); if (s) { free_temps(); return s; } }

    if (stats || !prog.temp_dev.empty()) {
        if (n > 0 && !prog.ops.empty())
            CU(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
        else memset(c->h_stats, 0, sizeof(DevStats));
        cudaError_t e = cudaStreamSynchronize(c->stream);
        free_temps();
        if (e != cudaSuccess) return fail(OPM_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e));
        for (OpmRays* r : prog.appended) if (r->len_dirty) { s = rays_refresh_len(r); if (s) return s; }
        if (compact) { s = rays_refresh_len(dst); if (s) return s; }
        if (stats) {
            stats_to_public(*c->h_stats, stats);
            // reported right here: not again at the next opm_context_synchronize
            if (stats->status_bits) CU(cudaMemsetAsync(&c->d_stats->sticky_bits, 0, 4, c->stream));
            return status_from_bits(stats->status_bits);
        }
    }
    return OPM_OK;
}

extern "C" OpmStatus opm_trace_sequence(OpmRays* rays, const OpmOp* ops, int32_t n_ops, int32_t flags, OpmRays* out,
                                        OpmTraceStats* stats) {
    return run_program(rays, ops, n_ops, flags, out, stats);
}

// ---------------------------------------------------------------------------------------------
// stand-alone `Rays` methods: 1-op programs through the same kernel
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_rays_refract_on_surface(OpmRays* rays, const OpmSurface* surface, const OpmIndexModel* n2,
                                                 int32_t refraction_intended, int32_t strategy, OpmRays* reflected,
                                                 OpmHitMap* hits, OpmTraceStats* stats) {
    if (!rays || !surface) return fail(OPM_ERR_ARG, "NULL argument");
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_REFRACT;
    op.u.refract.surface = *surface;
    if (n2) { op.u.refract.index = *n2; op.u.refract.has_index = 1; }
    op.u.refract.refraction_intended = refraction_intended;
    op.u.refract.strategy = strategy;
    op.u.refract.flags = reflected ? OPM_REFRACT_EMIT_CHILDREN : 0;
    op.u.refract.hits = hits;
    op.u.refract.children = reflected;
    OpmTraceStats local;
    return run_program(rays, &op, 1, 0, nullptr, stats ? stats : &local);
}
extern "C" OpmStatus opm_rays_diffract_on_periodic_surface(OpmRays* rays, const OpmSurface* surface, const OpmIndexModel* index,
                                                           const double gvec[3], int32_t order, int32_t refraction_intended,
                                                           OpmRays* diffracted, OpmTraceStats* stats) {
    if (!rays || !surface || !index || !gvec) return fail(OPM_ERR_ARG, "NULL argument");
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_DIFFRACT;
    op.u.diffract.surface = *surface;
    op.u.diffract.index = *index;
    for (int j = 0; j < 3; ++j) op.u.diffract.grating_vector[j] = gvec[j];
    op.u.diffract.order = order;
    op.u.diffract.refraction_intended = refraction_intended;
    op.u.diffract.flags = diffracted ? OPM_REFRACT_EMIT_CHILDREN : 0;
    op.u.diffract.children = diffracted;
    OpmTraceStats local;
    return run_program(rays, &op, 1, 0, nullptr, stats ? stats : &local);
}
extern "C" OpmStatus opm_rays_apodize(OpmRays* rays, const OpmAperture* ap, const OpmIsometry* iso, int32_t* invalidated) {
    if (!rays || !ap || !iso) return fail(OPM_ERR_ARG, "NULL argument");
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_APODIZE; op.u.apodize.iso = *iso; op.u.apodize.aperture = *ap;
    OpmTraceStats st;
    OpmStatus s = run_program(rays, &op, 1, 0, nullptr, &st);
    if (invalidated) *invalidated = st.apodized ? 1 : 0;
    return s;
}
static OpmStatus run_simple(OpmRays* rays, OpmOp& op) {
    OpmTraceStats st;
    return run_program(rays, &op, 1, 0, nullptr, &st);
}
extern "C" OpmStatus opm_rays_invalidate_by_threshold_energy(OpmRays* rays, double min_energy) {
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_THRESHOLD; op.u.threshold.min_energy = min_energy;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_filter_by_nr_of_bounces(OpmRays* rays, uint64_t max_bounces) {
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_LIMITS; op.u.limits.max_bounces = max_bounces; op.u.limits.check_refractions = 0;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_filter_by_nr_of_refractions(OpmRays* rays, uint64_t max_refr) {
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_LIMITS; op.u.limits.max_bounces = ~0ull; op.u.limits.max_refractions = max_refr; op.u.limits.check_refractions = 1;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_filter_energy(OpmRays* rays, double t) {
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_FILTER; op.u.filter.transmission = t;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_split(OpmRays* rays, double ratio, OpmRays* split_out) {
    if (!split_out) return fail(OPM_ERR_ARG, "split without output bundle");
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_SPLIT; op.u.split.ratio = ratio; op.u.split.split_out = split_out;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_filter_energy_spectrum(OpmRays* rays, const OpmSpectrum* spectrum) {
    if (!spectrum) return fail(OPM_ERR_ARG, "NULL spectrum");
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_FILTER; op.u.filter.spectrum = spectrum;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_split_spectrum(OpmRays* rays, const OpmSpectrum* spectrum, OpmRays* split_out) {
    if (!spectrum) return fail(OPM_ERR_ARG, "NULL spectrum");
    if (!split_out) return fail(OPM_ERR_ARG, "split without output bundle");
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_SPLIT; op.u.split.split_out = split_out; op.u.split.spectrum = spectrum;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_refract_paraxial(OpmRays* rays, double f, const OpmIsometry* iso) {
    if (!iso) return fail(OPM_ERR_ARG, "NULL argument");
    OpmOp op; memset(&op, 0, sizeof op);
    op.kind = OPM_OP_PARAXIAL; op.u.paraxial.focal_length = f; op.u.paraxial.iso = *iso;
    return run_simple(rays, op);
}
extern "C" OpmStatus opm_rays_transform(OpmRays* rays, const OpmIsometry* iso, int32_t inverse) {
    if (!rays || !iso) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = rays->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(rays);
    if (s) return s;
    if (rays->len == 0) return OPM_OK;
    DevIso I;
    to_dev_iso(inverse ? iso->inv : iso->fwd, I);
    KL(k_transform(rays->v, I, rays->len, c->sms, c->stream)); c->launches++;
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU detector read-out: combine the blocks an all-gather collected
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_readout_combine_ranges_sums(OpmContext* c, const double* gathered_dev, int32_t world, int32_t n_fluence,
                                                     int32_t n_spot, double* out_dev) {
    if (!c || !gathered_dev || !out_dev || world < 1 || n_fluence < 0 || n_spot < 0) return fail(OPM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    KL(k_readout_combine_a(gathered_dev, world, n_fluence, n_spot, out_dev, c->stream)); c->launches++;
    return OPM_OK;
}
extern "C" OpmStatus opm_readout_combine_maps_radii(OpmContext* c, const double* gathered_dev, int32_t world, uint64_t maps_len,
                                                    int32_t n_spot, double* out_dev) {
    if (!c || !gathered_dev || !out_dev || world < 1 || n_spot < 0) return fail(OPM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    KL(k_readout_combine_b(gathered_dev, world, maps_len, n_spot, out_dev, c->sms, c->stream)); c->launches++;
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU read-out over NVLink peer memory (opm_peer.cu)
// ---------------------------------------------------------------------------------------------
static OpmStatus check_same_ctx(OpmHitMap* const* maps, int n_maps, OpmContext** ctx);
struct OpmPeerGroup {
    OpmContext* ctx = nullptr;
    int rank = 0, world = 1;
    uint64_t map_bins = 0;
    unsigned char* seg = nullptr;           // this rank's segment: [PeerCtrl | part | final]
    void* peer_seg[kMaxPeers] = {};         // every rank's segment as mapped here (own: seg)
    bool connected = false;
    PeerView view{};
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    cudaEvent_t tev[5] = {};                // diagnostic: phase boundaries of the latest report (opm_peer_group_phase_times)
    bool timing = false;
    double2* d_cta_peaks = nullptr;         // (peak, total) of every CTA of the slice reduce
    double* d_work = nullptr;               // [own range 4 | all ranges world x 4 | range 4 | ... | last word: count of finished slice-reduce CTAs]
};
constexpr int kPeerWorkDoubles = 16 + 8 * kMaxPeers + 16;  // ... | 16 diagnostic time stamps at the end
static size_t peer_seg_bytes(uint64_t bins) { return kPeerCtrlBytes + 2 * align_up(bins * sizeof(double), 256); }
static double* peer_final(const OpmPeerGroup* g, int r) {
    return (double*)((unsigned char*)g->peer_seg[r] + kPeerCtrlBytes + align_up(g->map_bins * sizeof(double), 256));
}
extern "C" OpmStatus opm_peer_group_create(OpmContext* c, int32_t rank, int32_t world, uint64_t map_bins, OpmPeerGroup** out,
                                           uint8_t handle_out[OPM_IPC_HANDLE_BYTES]) {
    if (!c || !out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || map_bins == 0)
        return fail(OPM_ERR_ARG, "peer group: rank %d of %d (max %d ranks), %llu bins", rank, world, kMaxPeers, (unsigned long long)map_bins);
    static_assert(sizeof(cudaIpcMemHandle_t) == OPM_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    CU(cudaSetDevice(c->device));
    OpmPeerGroup* g = new OpmPeerGroup();
    g->ctx = c; g->rank = rank; g->world = world; g->map_bins = map_bins;
    if (const char* e = getenv("OPM_B200_PEER_TIMEOUT_MS")) g->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
    cudaError_t e = cudaMalloc(&g->seg, peer_seg_bytes(map_bins));
    if (e == cudaSuccess) e = cudaMemset(g->seg, 0, kPeerCtrlBytes);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_work, (size_t)kPeerWorkDoubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(g->d_work, 0, (size_t)kPeerWorkDoubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_cta_peaks, (size_t)c->sms * kPeerReduceCtasPerSm * sizeof(double2));
    if (e == cudaSuccess && world > 1 && handle_out) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out, g->seg);
    if (e != cudaSuccess) {
        cudaFree(g->seg); cudaFree(g->d_work); cudaFree(g->d_cta_peaks); delete g;
        return fail(OPM_ERR_CUDA, "peer group: %s", cudaGetErrorString(e));
    }
    g->peer_seg[rank] = g->seg;
    if (world == 1) { g->connected = true; }
    *out = g;
    return OPM_OK;
}
static void peer_build_view(OpmPeerGroup* g) {
    memset(&g->view, 0, sizeof g->view);
    g->view.rank = g->rank; g->view.world = g->world;
    for (int r = 0; r < g->world; ++r) {
        g->view.ctrl[r] = (PeerCtrl*)g->peer_seg[r];
        g->view.part[r] = (double*)((unsigned char*)g->peer_seg[r] + kPeerCtrlBytes);
    }
}
extern "C" OpmStatus opm_peer_group_connect(OpmPeerGroup* g, const uint8_t* handles) {
    if (!g) return fail(OPM_ERR_ARG, "peer group is NULL");
    CU(cudaSetDevice(g->ctx->device));
    if (g->world > 1 && !g->connected) {
        if (!handles) return fail(OPM_ERR_ARG, "peer group: handles of all ranks needed");
        for (int r = 0; r < g->world; ++r) {
            if (r == g->rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)r * OPM_IPC_HANDLE_BYTES, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(&g->peer_seg[r], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(OPM_ERR_CUDA, "peer group: cannot map the segment of rank %d (%s) -- no peer access between the GPUs?", r,
                            cudaGetErrorString(e));
            }
        }
        g->connected = true;
    }
    peer_build_view(g);
    return OPM_OK;
}
extern "C" void opm_peer_group_destroy(OpmPeerGroup* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    for (int r = 0; r < g->world; ++r)
        if (r != g->rank && g->peer_seg[r]) cudaIpcCloseMemHandle(g->peer_seg[r]);
    cudaFree(g->seg); cudaFree(g->d_work); cudaFree(g->d_cta_peaks);
    if (g->timing) for (auto& e : g->tev) cudaEventDestroy(e);
    delete g;
}
static OpmStatus peer_meet(OpmPeerGroup* g, const double* block, int n, double* out, int mode, double* out2) {
    OpmContext* c = g->ctx;
    if (!g->connected) return fail(OPM_ERR_ARG, "peer group is not connected");
    if (g->view.world != g->world) peer_build_view(g);
    KL(k_peer_meet(g->view, ++g->seq, block, n, out, mode, out2, &c->d_stats->sticky_bits, g->timeout_ns, nullptr, nullptr, c->stream));
    c->launches++;
    return OPM_OK;
}
extern "C" OpmStatus opm_peer_allgather_async(OpmPeerGroup* g, const double* block_dev, int32_t n, double* out_dev) {
    if (!g || !block_dev || !out_dev || n < 1 || n > kPeerBlock) return fail(OPM_ERR_ARG, "peer all-gather: 1..%d doubles", kPeerBlock);
    CU(cudaSetDevice(g->ctx->device));
    return peer_meet(g, block_dev, n, out_dev, 0, nullptr);
}
extern "C" OpmStatus opm_peer_fluence_report_async(OpmPeerGroup* g, OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx,
                                                   uint64_t ny, double* range_dev, double* peak_total_dev, uint32_t* status_dev,
                                                   double* host_report) {
    if (!g || !range_dev || !peak_total_dev || !status_dev || nx < 2 || ny < 2 || nx * ny > g->map_bins) return fail(OPM_ERR_ARG, "peer fluence report: bad map arguments");
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (c != g->ctx) return fail(OPM_ERR_ARG, "hit maps and peer group live on different contexts");
    CU(cudaSetDevice(c->device));
    if (g->view.world != g->world) peer_build_view(g);
    double* own = g->d_work;                         // 4
    double* all = own + 4;                           // world x 4
    double* range = all + 4 * kMaxPeers;             // 4
    unsigned* done = (unsigned*)(g->d_work + 8 + 4 * kMaxPeers);          // finished CTAs of the slice reduce; left at 0 by every read-out
    double2* cta_peaks = g->d_cta_peaks;
    if (!g->connected) return fail(OPM_ERR_ARG, "peer group is not connected");
    // meeting 1: every rank's own range, the common range.  When the trace kernel kept the bounding box while it recorded the hits
    // the meeting kernel reads that accumulator itself
    auto mark = [&](int k) { if (g->timing) cudaEventRecord(g->tev[k], c->stream); };
    mark(0);
    double* part = g->view.part[g->rank];
    // meeting 1: every rank's own range, the common range.  When the trace kernel kept the bounding box while it recorded the hits,
    // the kernel that zero-fills the window carries the meeting at its head; else bounding-box pass, meeting kernel, zero-fill
    PeerHook hook;
    memset(&hook, 0, sizeof hook);
    if (bounce_filter < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
    if (n_maps == 1 && bounce_filter == -1 && maps[0]->acc_valid && maps[0]->len) {
        hook.enabled = 1; hook.V = g->view; hook.seq = ++g->seq; hook.timeout_ns = g->timeout_ns;
        hook.acc = (const BBoxAccum*)maps[0]->d_acc; hook.own_out = own; hook.all_out = all; hook.range_out = range;
        hook.status = &c->d_stats->sticky_bits;
        hook.stamps = g->timing ? (unsigned long long*)(g->d_work + kPeerWorkDoubles - 16) : nullptr;
        mark(1);
    } else {
        s = opm_hitmaps_bounding_box_async(maps, n_maps, bounce_filter, own);
        if (s) return s;
        s = peer_meet(g, own, 4, all, 1, range);
        if (s) return s;
        mark(1);
    }
    KL(k_peer_zero_window(part, nx, ny, own, range, hook, c->sms, c->stream)); c->launches++;
    mark(2);
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) { KL(k_binning(maps[k]->v, maps[k]->len, bounce_filter, nx, ny, range, part, status_dev, c->sms, c->stream)); c->launches++; }
    mark(3);
    // meetings 2 ("every part is binned") and 3 (peak and total of the whole map) are the head and the tail of the slice reduce
    const unsigned long long seq2 = ++g->seq;
    ++g->seq;
    KL(k_peer_reduce_slice(g->view, nx, ny, all, range, peer_final(g, g->rank), cta_peaks, done, seq2, peak_total_dev, range_dev,
                           &c->d_stats->sticky_bits, g->timeout_ns, g->timing ? (unsigned long long*)(g->d_work + kPeerWorkDoubles - 16) : nullptr,
                           status_dev, host_report, c->sms, c->stream));
    c->launches++;
    mark(4);
    return OPM_OK;
}
// Diagnostic: device time of the phases of the latest opm_peer_fluence_report_async (meeting 1 | zero-fill of the window | splat |
// slice reduce with meetings 2 and 3), in milliseconds.  enable != 0 switches the event records on for the following reports;
// out_ms == NULL only switches.  Synchronises the stream.
extern "C" OpmStatus opm_peer_group_phase_times(OpmPeerGroup* g, int32_t enable, float out_ms[12]) {
    if (!g) return fail(OPM_ERR_ARG, "peer group is NULL");
    CU(cudaSetDevice(g->ctx->device));
    if (out_ms) {
        if (!g->timing) return fail(OPM_ERR_ARG, "phase timing was not enabled");
        CU(cudaStreamSynchronize(g->ctx->stream));
        for (int k = 0; k < 4; ++k) CU(cudaEventElapsedTime(&out_ms[k], g->tev[k], g->tev[k + 1]));
        unsigned long long st[8];  // in-kernel time stamps (globaltimer, ns) relative to the start of the splat kernel
        CU(cudaMemcpy(st, g->d_work + kPeerWorkDoubles - 16, sizeof st, cudaMemcpyDeviceToHost));
        for (int k = 0; k < 8; ++k) out_ms[4 + k] = (float)(((double)st[k] - (double)st[0]) * 1e-6);
    }
    if (enable && !g->timing) { for (auto& e : g->tev) CU(cudaEventCreate(&e)); g->timing = true; }
    return OPM_OK;
}
extern "C" OpmStatus opm_peer_fluence_gather_map(OpmPeerGroup* g, uint64_t nx, uint64_t ny, double* map_dev) {
    if (!g || !map_dev || nx * ny > g->map_bins) return fail(OPM_ERR_ARG, "peer gather: bad arguments");
    OpmContext* c = g->ctx;
    CU(cudaSetDevice(c->device));
    for (int r = 0; r < g->world; ++r) {
        const uint64_t c0 = (uint64_t)r * nx / g->world, c1 = (uint64_t)(r + 1) * nx / g->world;
        if (c1 > c0) CU(cudaMemcpyAsync(map_dev + c0 * ny, peer_final(g, r) + c0 * ny, (c1 - c0) * ny * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// spectra, spectrometer, wavefront (spectrum.rs, nodes/spectrometer.rs, nodes/wavefront.rs)
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_spectrum_create(OpmContext* c, const double* lambda_um, const double* values, uint64_t n, OpmSpectrum** out) {
    if (!c || !lambda_um || !values || !out || n == 0) return fail(OPM_ERR_ARG, "NULL argument or empty spectrum");
    for (uint64_t i = 1; i < n; ++i)
        if (!(lambda_um[i] > lambda_um[i - 1])) return fail(OPM_ERR_SPECTRUM, "spectrum wavelengths must be ascending");
    CU(cudaSetDevice(c->device));
    OpmSpectrum* sp = new OpmSpectrum();
    sp->ctx = c; sp->n = n;
    cudaError_t e = cudaMalloc(&sp->d, 2 * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(sp->d, lambda_um, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sp->d + n, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // the caller's arrays may go away
    if (e != cudaSuccess) { cudaFree(sp->d); delete sp; return fail(OPM_ERR_CUDA, "spectrum upload failed: %s", cudaGetErrorString(e)); }
    *out = sp;
    return OPM_OK;
}
extern "C" void opm_spectrum_destroy(OpmSpectrum* sp) {
    if (!sp) return;
    cudaSetDevice(sp->ctx->device);
    cudaStreamSynchronize(sp->ctx->stream);
    cudaFree(sp->d);
    delete sp;
}
// Spectrum::total_energy (spectrum.rs:289-298): Kahan sum (kahan 0.1.4) of slot width x value
extern "C" double opm_spectrum_total_energy(const double* lambda_um, const double* values, uint64_t n) {
    double sum = 0.0, err = 0.0;
    for (uint64_t i = 0; i + 1 < n; ++i) {
        const double x = (lambda_um[i + 1] - lambda_um[i]) * values[i];
        const double y = x - err, t = sum + y;
        err = (t - sum) - y;
        sum = t;
    }
    return sum;
}
// Spectrum::center_wavelength (spectrum.rs:304-316), returned in metres (micrometer!(x) = x * 1e-6)
extern "C" double opm_spectrum_center_wavelength(const double* lambda_um, const double* values, uint64_t n) {
    double weighted_sum = 0.0, total_weight = 0.0;
    for (uint64_t i = 0; i + 1 < n; ++i) {
        const double bin_width = lambda_um[i + 1] - lambda_um[i];
        const double bin_weight = values[i] * bin_width;
        weighted_sum += lambda_um[i] * bin_weight;
        total_weight += bin_weight;
    }
    return (weighted_sum / total_weight) * 1e-6;
}
// Spectrum::get_value (spectrum.rs:321-349)
extern "C" int32_t opm_spectrum_get_value(const double* lam, const double* val, uint64_t n, double wl_m, double* out) {
    if (!lam || !val || !out || n == 0) return 0;
    const double w = wl_m * 1e6;
    if (w == lam[n - 1]) { *out = val[n - 1]; return 1; }
    if (!(lam[0] * 1e-6 <= wl_m && wl_m < lam[n - 1] * 1e-6)) return 0;
    uint64_t idx = n;
    for (uint64_t i = 0; i < n; ++i) if (lam[i] >= w) { idx = i; break; }
    if (idx == n) return 0;
    const uint64_t l = idx == 0 ? 0 : idx - 1, r = idx == 0 ? 1 : idx;
    const double ratio = (w - lam[l]) / (lam[r] - lam[l]);
    *out = std::fma(val[l], 1.0 - ratio, val[r] * ratio);
    return 1;
}
extern "C" OpmStatus opm_rays_wavelength_range(const OpmRays* rc, double min_max_m[2], uint64_t* n_valid) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !min_max_m) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_materialize(r);
    if (!s) s = rays_refresh_len(r);
    if (s) return s;
    WvlAccum init;
    init.min_key = double_to_key(INFINITY); init.max_key = double_to_key(-INFINITY); init.count = 0;
    WvlAccum* d_acc = (WvlAccum*)(c->d_scratch + 3200);
    s = dev_put(c, d_acc, &init, sizeof init);
    if (s) return s;
    if (r->len) { KL(k_wavelength_range(r->v, r->len, d_acc, c->sms, c->stream)); c->launches++; }
    WvlAccum* h = (WvlAccum*)(c->h_scratch + 3200);
    CU(cudaMemcpyAsync(h, d_acc, sizeof *h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    min_max_m[0] = key_to_double(h->min_key); min_max_m[1] = key_to_double(h->max_key);
    if (n_valid) *n_valid = h->count;
    return OPM_OK;
}
// Spectrum::new (spectrum.rs:47-73): slot count and the two numbers every slot wavelength derives from
static OpmStatus spectrum_grid(double start_m, double end_m, double res_m, uint64_t* n, double* start_um, double* step_um) {
    if (!(res_m > 0.0)) return fail(OPM_ERR_SPECTRUM, "resolution must be positive");
    if (start_m >= end_m) return fail(OPM_ERR_SPECTRUM, "wavelength range must be in ascending order and not empty");
    if (std::signbit(start_m) || std::signbit(end_m)) return fail(OPM_ERR_SPECTRUM, "wavelength range limits must both be positive");
    const double cnt = std::round((end_m - start_m) / res_m);
    *n = cnt > 0.0 ? (uint64_t)cnt : 0;  // f64_to_usize (utils/math_utils.rs:38-43)
    *start_um = start_m * 1e6; *step_um = res_m * 1e6;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_to_spectrum(const OpmRays* rc, double resolution_m, const double* range_m, double* lambda_um, double* values,
                                          uint64_t cap, uint64_t* n_out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !n_out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    double mm[2];
    if (range_m) { mm[0] = range_m[0]; mm[1] = range_m[1]; }
    else {
        uint64_t nv = 0;
        OpmStatus s = opm_rays_wavelength_range(r, mm, &nv);
        if (s) return s;
        if (nv == 0) return fail(OPM_ERR_SPECTRUM, "ray bundle is empty - cannot create spectrum");  // rays.rs:1216-1220
    }
    if (!(resolution_m > 0.0)) return fail(OPM_ERR_SPECTRUM, "resolution must be positive");
    uint64_t n = 0;
    double start = 0.0, step = 0.0;
    OpmStatus s = spectrum_grid(mm[0], mm[1] + 2.0 * resolution_m, resolution_m, &n, &start, &step);  // spectrum.rs:143
    if (s) return s;
    *n_out = n;
    if (!lambda_um) return OPM_OK;
    if (!values || cap < n) return fail(OPM_ERR_CAPACITY, "spectrum has %llu slots, the buffers hold %llu", (unsigned long long)n, (unsigned long long)cap);
    for (uint64_t i = 0; i < n; ++i) lambda_um[i] = std::fma((double)i, step, start);
    if (n == 0) return OPM_OK;
    CU(cudaSetDevice(c->device));
    s = rays_materialize(r);
    if (!s) s = rays_refresh_len(r);
    if (s) return s;
    double* d_data = nullptr;
    CU(cudaMalloc(&d_data, n * sizeof(double)));
    unsigned* d_status = (unsigned*)(c->d_scratch + 1024);
    cudaError_t e = cudaMemsetAsync(d_data, 0, n * sizeof(double), c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, 4, c->stream);
    if (e == cudaSuccess && r->len) { e = (cudaError_t)k_spectrum_splat(r->v, r->len, start, step, n, d_data, d_status, c->sms, c->stream); c->launches++; }
    unsigned* h_status = (unsigned*)(c->h_scratch + 1024);
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, d_data, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_status, d_status, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_data);
    if (e != cudaSuccess) return fail(OPM_ERR_CUDA, "spectrum kernel failed: %s", cudaGetErrorString(e));
    return status_from_bits(*h_status);
}
extern "C" OpmStatus opm_rays_wavefront_error(const OpmRays* rc, const OpmIsometry* monitor_iso, double wavelength_m, double* xyw_out,
                                              uint64_t cap, OpmWavefrontStats* out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !monitor_iso || !out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    memset(out, 0, sizeof *out);
    if (!(wavelength_m > 0.0)) {  // WaveFront::node_report: centre wavelength of to_spectrum(1 nm) (rays.rs:719-726)
        uint64_t ns = 0;
        OpmStatus s = opm_rays_to_spectrum(r, 1.0e-9, nullptr, nullptr, nullptr, 0, &ns);
        if (s) return s;
        std::vector<double> lam(ns ? ns : 1), val(ns ? ns : 1);
        s = opm_rays_to_spectrum(r, 1.0e-9, nullptr, lam.data(), val.data(), ns, &ns);
        if (s) return s;
        wavelength_m = opm_spectrum_center_wavelength(lam.data(), val.data(), ns);
    }
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_materialize(r);
    if (!s) s = rays_refresh_len(r);
    if (s) return s;
    const uint64_t n = r->len;
    WfAccum init;
    memset(&init, 0, sizeof init);
    init.min_r_key = double_to_key(INFINITY); init.first_slot = ~0ull;
    init.max_key = double_to_key(-INFINITY); init.min_key = double_to_key(INFINITY);
    WfAccum* d_acc = (WfAccum*)(c->d_scratch + 3200);
    s = dev_put(c, d_acc, &init, sizeof init);
    if (s) return s;
    double* d_xyw = nullptr;
    if (xyw_out && n) CU(cudaMalloc(&d_xyw, n * 3 * sizeof(double)));
    DevIso inv;
    to_dev_iso(monitor_iso->inv, inv);
    cudaError_t e = cudaSuccess;
    if (n) { e = (cudaError_t)k_wavefront(r->v, n, inv, wavelength_m * (1.0 / 1e-9), d_acc, d_xyw, c->sms, c->stream); c->launches += 4; }
    WfAccum* h = (WfAccum*)(c->h_scratch + 3200);
    std::vector<double> tmp;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_acc, sizeof *h, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && d_xyw) { tmp.resize(n * 3); e = cudaMemcpyAsync(tmp.data(), d_xyw, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_xyw);
    if (e != cudaSuccess) return fail(OPM_ERR_CUDA, "wavefront kernels failed: %s", cudaGetErrorString(e));
    out->wavelength_m = wavelength_m;
    out->n_rays = h->count;
    if (h->count == 0) return fail(OPM_ERR_EMPTY_WAVEFRONT, "%s", opm_status_string(OPM_ERR_EMPTY_WAVEFRONT));  // nodes/wavefront.rs:121-123
    out->ptv = key_to_double(h->max_key) - key_to_double(h->min_key);
    out->rms = std::sqrt(h->sum2 / (double)(int32_t)h->count);
    if (xyw_out) {
        uint64_t k = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (std::isnan(tmp[3 * i])) continue;  // not a valid ray
            if (k >= cap) return fail(OPM_ERR_CAPACITY, "wavefront buffer holds %llu rows, the bundle has %llu valid rays", (unsigned long long)cap, (unsigned long long)h->count);
            memcpy(xyw_out + 3 * k, tmp.data() + 3 * i, 3 * sizeof(double));
            ++k;
        }
    }
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------------
// Two-phase form (multi-GPU: the partial accumulators are all-reduced between the phases):
// phase 1 = sums over valid rays, phase 2 = radii about the centroids the (global) sums define.  The device-chained
// forms keep both on the device in reduce-friendly double layouts (see opm_kernels.cu) and never synchronise.
extern "C" OpmStatus opm_rays_spot_sums_async(const OpmRays* rc, const OpmIsometry* frame, double* sums_dev) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !sums_dev) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    {   // the launch that wrote this copy took the sums on the way (OPM_OP_SNAPSHOT want_spot_sums), in this very frame?
        DevIso want;
        memset(&want, 0, sizeof want);
        if (frame) to_dev_iso(frame->inv, want);
        if (r->spot_acc_valid && r->spot_has_frame == (frame != nullptr) && (!frame || memcmp(&want, &r->spot_inv, sizeof want) == 0)) {
            KL(k_spot_sums_store(r->d_spot_acc, sums_dev, c->stream)); c->launches++;
            return OPM_OK;
        }
    }
    SpotAccum init;
    memset(&init, 0, sizeof init);
    init.first_key = ~0ull;
    init.max_d2_key = double_to_key(0.0);
    SpotAccum* d_acc = (SpotAccum*)c->d_scratch;
    s = dev_put(c, d_acc, &init, sizeof init);
    if (s) return s;
    DevIso inv;
    memset(&inv, 0, sizeof inv);
    if (frame) to_dev_iso(frame->inv, inv);
    if (r->len) { KL(k_spot_pass1(r->v, r->len, frame ? 1 : 0, inv, d_acc, c->sms, c->stream)); c->launches++; }
    KL(k_spot_sums_store(d_acc, sums_dev, c->stream)); c->launches++;
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_spot_radii_async(const OpmRays* rc, const OpmIsometry* frame, const double* sums_dev, double* radii_dev) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !sums_dev || !radii_dev) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    SpotAccum* d_acc = (SpotAccum*)c->d_scratch;
    KL(k_spot_sums_load(sums_dev, d_acc, c->stream)); c->launches++;
    DevIso inv;
    memset(&inv, 0, sizeof inv);
    if (frame) to_dev_iso(frame->inv, inv);
    if (r->len) { KL(k_spot_pass2(r->v, r->len, frame ? 1 : 0, inv, d_acc, c->sms, c->stream)); c->launches++; }
    KL(k_spot_radii_store(d_acc, radii_dev, c->stream)); c->launches++;
    return OPM_OK;
}
static void partial_from_sums(const double* h, OpmSpotPartial* p) {
    memset(p, 0, sizeof *p);
    memcpy(p->sum, h, sizeof p->sum);
    p->n_valid = (uint64_t)h[7]; p->n_total = (uint64_t)h[8];
    p->first_key = h[9] < 0.0 ? ~0ull : (((uint64_t)h[10] << 32) | (uint64_t)h[9]);
}
static void sums_from_partial(const OpmSpotPartial* p, double* h) {
    memcpy(h, p->sum, sizeof p->sum);
    h[7] = (double)p->n_valid; h[8] = (double)p->n_total;
    const bool none = p->first_key == ~0ull || p->n_valid == 0;
    h[9] = none ? -1.0 : (double)(p->first_key & 0xffffffffull);
    h[10] = none ? -1.0 : (double)(p->first_key >> 32);
}
extern "C" OpmStatus opm_rays_spot_partial_sums(const OpmRays* rays, const OpmIsometry* frame, OpmSpotPartial* out) {
    if (!rays || !out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = rays->ctx;
    OpmStatus s = opm_rays_spot_sums_async(rays, frame, scratch_sums(c));
    if (s) return s;
    double* h = (double*)(c->h_scratch + 2112);
    CU(cudaMemcpyAsync(h, scratch_sums(c), 88, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    partial_from_sums(h, out);
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_spot_partial_radii(const OpmRays* rays, const OpmIsometry* frame, OpmSpotPartial* inout) {
    if (!rays || !inout) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = rays->ctx;
    CU(cudaSetDevice(c->device));
    double enc[11];
    sums_from_partial(inout, enc);
    OpmStatus s = dev_put(c, scratch_sums(c), enc, sizeof enc);
    if (!s) s = opm_rays_spot_radii_async(rays, frame, scratch_sums(c), scratch_radii(c));
    if (s) return s;
    double* h = (double*)(c->h_scratch + 2240);
    CU(cudaMemcpyAsync(h, scratch_radii(c), 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    inout->max_d2_mm2 = h[0]; inout->sum_d2_mm2 = h[1]; inout->sum_ed2 = h[2];
    return OPM_OK;
}
extern "C" void opm_spot_stats_from_partial(const OpmSpotPartial* a, OpmSpotStats* out) {
    memset(out, 0, sizeof *out);
    out->n_valid = a->n_valid; out->n_total = a->n_total; out->total_energy = a->sum[6];
    const double nan = std::nan("");
    if (a->n_valid == 0) {  // the reference returns None
        for (int k = 0; k < 3; ++k) out->centroid[k] = out->energy_centroid[k] = nan;
        out->beam_radius_geo = out->beam_radius_rms = out->energy_beam_radius_rms = nan;
        return;
    }
    double len = (double)a->n_valid;
    for (int k = 0; k < 3; ++k) { out->centroid[k] = a->sum[k] / len; out->energy_centroid[k] = a->sum[3 + k] / a->sum[6]; }
    out->beam_radius_geo = std::sqrt(a->max_d2_mm2) * 1.0E-3;  // millimeter!(max_dist)
    out->beam_radius_rms = std::sqrt(a->sum_d2_mm2 / len) * 1.0E-3;
    out->energy_beam_radius_rms = std::sqrt(a->sum_ed2 / a->sum[6]);
    out->bounce_lvl = (uint32_t)(a->first_key & 0xffffffffull);
}
extern "C" OpmStatus opm_rays_spot_stats(const OpmRays* rc, const OpmIsometry* frame, OpmSpotStats* out) {
    if (!rc || !out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = rc->ctx;
    OpmStatus s = opm_rays_spot_sums_async(rc, frame, scratch_sums(c));          // both phases chained on the device,
    if (!s) s = opm_rays_spot_radii_async(rc, frame, scratch_sums(c), scratch_radii(c));  // one synchronisation
    if (s) return s;
    double* h = (double*)(c->h_scratch + 2112);
    CU(cudaMemcpyAsync(h, scratch_sums(c), 80 + 48 + 24, cudaMemcpyDeviceToHost, c->stream));  // sums .. radii are contiguous in the scratch
    CU(cudaStreamSynchronize(c->stream));
    OpmSpotPartial p;
    partial_from_sums(h, &p);
    const double* hr = (const double*)(c->h_scratch + 2240);
    p.max_d2_mm2 = hr[0]; p.sum_d2_mm2 = hr[1]; p.sum_ed2 = hr[2];
    opm_spot_stats_from_partial(&p, out);
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_bounce_lvl(const OpmRays* rc, uint32_t* out) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    *out = 0;
    if (r->len == 0) return OPM_OK;
    unsigned long long* d_key = (unsigned long long*)c->d_scratch;
    CU(cudaMemsetAsync(d_key, 0xff, 8, c->stream));
    KL(k_bounce_lvl(r->v, r->len, d_key, c->sms, c->stream)); c->launches++;
    CU(cudaMemcpyAsync(c->h_scratch, d_key, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long key = *(unsigned long long*)c->h_scratch;
    if (key != ~0ull) *out = (uint32_t)(key & 0xffffffffull);
    return OPM_OK;
}
extern "C" OpmStatus opm_rays_tally(const OpmRays* rc, uint32_t n_orders, uint64_t* counts, double* energy) {
    OpmRays* r = const_cast<OpmRays*>(rc);
    if (!r || !counts || !energy || n_orders == 0 || n_orders > 128) return fail(OPM_ERR_ARG, "bad argument (n_orders in 1..128)");
    OpmContext* c = r->ctx;
    CU(cudaSetDevice(c->device));
    OpmStatus s = rays_refresh_len(r);
    if (s) return s;
    double* d_e = (double*)c->d_scratch;
    unsigned long long* d_c = (unsigned long long*)(c->d_scratch + n_orders * 8);
    CU(cudaMemsetAsync(c->d_scratch, 0, n_orders * 16, c->stream));
    if (r->len) { KL(k_tally(r->v, r->len, n_orders, d_c, d_e, c->sms, c->stream)); c->launches++; }
    CU(cudaMemcpyAsync(c->h_scratch, c->d_scratch, n_orders * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(energy, c->h_scratch, n_orders * 8);
    memcpy(counts, c->h_scratch + n_orders * 8, n_orders * 8);
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// fluence
// ---------------------------------------------------------------------------------------------
static OpmStatus check_same_ctx(OpmHitMap* const* maps, int n_maps, OpmContext** ctx) {
    if (!maps || n_maps <= 0) return fail(OPM_ERR_ARG, "no hit maps given");
    *ctx = maps[0]->ctx;
    for (int k = 0; k < n_maps; ++k)
        if (!maps[k] || maps[k]->ctx != *ctx) return fail(OPM_ERR_ARG, "hit maps must share one context");
    return OPM_OK;
}
// ---- device-chained forms: asynchronous on the context's stream, results stay on the device ------------------
// scratch layout: [0,1024) accumulators, [1024,1028) status, [2048,2080) range, [2112,2192) spot sums, [2240,2264) radii,
// [2304,2320) peak/total

// used_acc: where the accumulator of this call lives on the device (the context's scratch, or the hit map's own)
static OpmStatus bounding_box_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, double* range_dev, const BBoxAccum** used_acc);
extern "C" OpmStatus opm_hitmaps_bounding_box_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, double* range_dev) {
    return bounding_box_async(maps, n_maps, bounce_filter, range_dev, nullptr);
}
static OpmStatus bounding_box_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, double* range_dev, const BBoxAccum** used_acc) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (!range_dev) return fail(OPM_ERR_ARG, "range_dev is NULL");
    CU(cudaSetDevice(c->device));
    if (bounce_filter == OPM_BOUNCE_OF_LAST_LAUNCH) bounce_filter = -(int64_t)(uintptr_t)&c->d_stats->first_inv;  // read on the device
    else if (bounce_filter < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
    if (used_acc) *used_acc = (const BBoxAccum*)c->d_scratch;
    if (n_maps == 1 && bounce_filter == -1 && maps[0]->acc_valid && maps[0]->len) {  // the trace kernel kept the box while it recorded
        if (used_acc) *used_acc = (const BBoxAccum*)maps[0]->d_acc;
        KL(k_bbox_finalize((const BBoxAccum*)maps[0]->d_acc, range_dev, c->stream)); c->launches++;
        return OPM_OK;
    }
    BBoxAccum init;
    init.key[0] = ~0ull; init.key[1] = 0; init.key[2] = 0; init.key[3] = ~0ull; init.count = 0;
    BBoxAccum* d_acc = (BBoxAccum*)c->d_scratch;
    { OpmStatus ps = dev_put(c, d_acc, &init, sizeof init); if (ps) return ps; }
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) { KL(k_hits_bbox(maps[k]->v, maps[k]->len, bounce_filter, d_acc, c->sms, c->stream)); c->launches++; }
    KL(k_bbox_finalize(d_acc, range_dev, c->stream)); c->launches++;
    return OPM_OK;
}
extern "C" OpmStatus opm_fluence_binning_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                               const double* range_dev, int32_t accumulate, double* map_dev, uint32_t* status_dev) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (!map_dev || !range_dev || !status_dev || nx == 0 || ny == 0) return fail(OPM_ERR_ARG, "bad map arguments");
    CU(cudaSetDevice(c->device));
    if (bounce_filter == OPM_BOUNCE_OF_LAST_LAUNCH) bounce_filter = -(int64_t)(uintptr_t)&c->d_stats->first_inv;  // read on the device
    else if (bounce_filter < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
    if (!accumulate) CU(cudaMemsetAsync(map_dev, 0, nx * ny * sizeof(double), c->stream));
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) { KL(k_binning(maps[k]->v, maps[k]->len, bounce_filter, nx, ny, range_dev, map_dev, status_dev, c->sms, c->stream)); c->launches++; }
    return OPM_OK;
}
// ---- per-bundle peak-fluence check of the ghost-focus analysis: three launches, no parameter uploads (opm_kernels.cu)
extern "C" OpmStatus opm_fluence_peak_check_acc_init(OpmContext* c, void* acc_dev, uint64_t n) {
    if (!c || !acc_dev) return fail(OPM_ERR_ARG, "NULL argument");
    static_assert(sizeof(PeakCheckAcc) == OPM_PEAK_CHECK_ACC_BYTES, "accumulator size");
    CU(cudaSetDevice(c->device));
    if (n) { KL(k_peakcheck_acc_init((PeakCheckAcc*)acc_dev, n, c->stream)); c->launches++; }
    return OPM_OK;
}
static OpmStatus peak_check_impl(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny, void* acc_dev,
                                 double* map_dev, double* result_dev, bool side);
extern "C" OpmStatus opm_fluence_peak_check_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                                  void* acc_dev, double* map_dev, double* result_dev) {
    return peak_check_impl(maps, n_maps, bounce_filter, nx, ny, acc_dev, map_dev, result_dev, false);
}
extern "C" OpmStatus opm_fluence_peak_check_side_async(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                                       void* acc_dev, double* map_dev, double* result_dev) {
    return peak_check_impl(maps, n_maps, bounce_filter, nx, ny, acc_dev, map_dev, result_dev, true);
}
extern "C" OpmStatus opm_context_join_side_stream(OpmContext* c) {
    if (!c) return fail(OPM_ERR_ARG, "context is NULL");
    CU(cudaSetDevice(c->device));
    if (OpmStatus s = flush_pending_keys(c)) return s;
    if (!c->side_busy) return OPM_OK;
    CU(cudaEventRecord(c->side_ev, c->side_stream));
    CU(cudaStreamWaitEvent(c->stream, c->side_ev, 0));
    c->side_busy = false;
    return OPM_OK;
}
static OpmStatus peak_check_impl(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny, void* acc_dev,
                                 double* map_dev, double* result_dev, bool side) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (!acc_dev || !map_dev || !result_dev || nx < 2 || ny < 2 || nx * ny * sizeof(double) > 160 * 1024)
        return fail(OPM_ERR_ARG, "peak check: bad arguments (maps up to 160 KB)");
    CU(cudaSetDevice(c->device));
    const unsigned long long* key_src = nullptr;
    if (bounce_filter == OPM_BOUNCE_OF_LAST_LAUNCH) {
        key_src = &c->d_stats->first_inv;
        if (side) {  // the next trace launch resets the statistic: the check takes its own copy along (the result slot's key word)
            unsigned long long* own = reinterpret_cast<unsigned long long*>(result_dev) + 7;
            CU(cudaMemcpyAsync(own, key_src, 8, cudaMemcpyDeviceToDevice, c->stream));
            key_src = own;
        }
        bounce_filter = -(int64_t)(uintptr_t)key_src;
    } else if (bounce_filter < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
    cudaStream_t st = c->stream;
    if (side) {  // behind everything the main stream holds so far, beside what it gets from now on
        if (!c->side_stream) { CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&c->side_ev, cudaEventDisableTiming)); }
        CU(cudaEventRecord(c->side_ev, c->stream));
        CU(cudaStreamWaitEvent(c->side_stream, c->side_ev, 0));
        st = c->side_stream;
        c->side_busy = true;
    }
    int last = -1;
    for (int k = 0; k < n_maps; ++k) if (maps[k]->len) last = k;
    bool first = true;
    for (int k = 0; k < n_maps; ++k) {
        if (!maps[k]->len && !(last < 0 && k == n_maps - 1)) continue;
        // (a check without any recorded slot still runs one launch: it zero-fills the map and writes the empty range)
        KL(k_peakcheck_bbox(maps[k]->v, maps[k]->len, bounce_filter, (PeakCheckAcc*)acc_dev, result_dev, map_dev, nx * ny, first ? 1 : 0,
                            (k == last || last < 0) ? 1 : 0, c->sms, st));
        c->launches++;
        first = false;
    }
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) { KL(k_binning(maps[k]->v, maps[k]->len, bounce_filter, nx, ny, result_dev, map_dev, (unsigned*)(result_dev + 6), c->sms, st)); c->launches++; }
    KL(k_peakcheck_peak(map_dev, nx, ny, key_src, result_dev, st)); c->launches++;
    return OPM_OK;
}
// the bundle's bounce level for a deferred check: the latest launch's first_valid_key statistic, copied into the key word of the
// check's result slot in stream order (the next trace launch resets the statistic)
static OpmStatus flush_pending_keys(OpmContext* c) {
    for (int k = 0; k < 2; ++k) {
        if (!c->pending_key[k]) continue;
        const unsigned long long* src = k ? &c->d_stats->first_inv2 : &c->d_stats->first_inv;
        CU(cudaMemcpyAsync(c->pending_key[k], src, 8, cudaMemcpyDeviceToDevice, c->stream));
        c->pending_key[k] = nullptr;
    }
    return OPM_OK;
}
extern "C" OpmStatus opm_fluence_peak_check_capture_key_async(OpmContext* c, double* result_dev, int32_t slot) {
    if (!c || !result_dev || slot < 0 || slot > 1) return fail(OPM_ERR_ARG, "NULL argument / key slot not 0 or 1");
    CU(cudaSetDevice(c->device));
    if (c->pending_key[slot]) { OpmStatus s = flush_pending_keys(c); if (s) return s; }  // one destination per key and launch
    // no copy now: the statistic stays where it is until the next trace launch, whose prologue delivers it before resetting it
    c->pending_key[slot] = reinterpret_cast<unsigned long long*>(result_dev) + 7;
    return OPM_OK;
}
static OpmStatus peak_check_batch_impl(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks, const int64_t* bounce_filters,
                                       uint64_t nx, uint64_t ny, void* acc_dev, double* maps_dev, double* const* results_dev, bool side);
extern "C" OpmStatus opm_fluence_peak_check_batch_async(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks,
                                                        const int64_t* bounce_filters, uint64_t nx, uint64_t ny, void* acc_dev, double* maps_dev,
                                                        double* const* results_dev) {
    return peak_check_batch_impl(maps, maps_per_check, n_checks, bounce_filters, nx, ny, acc_dev, maps_dev, results_dev, false);
}
extern "C" OpmStatus opm_fluence_peak_check_batch_side_async(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks,
                                                             const int64_t* bounce_filters, uint64_t nx, uint64_t ny, void* acc_dev,
                                                             double* maps_dev, double* const* results_dev) {
    return peak_check_batch_impl(maps, maps_per_check, n_checks, bounce_filters, nx, ny, acc_dev, maps_dev, results_dev, true);
}
static OpmStatus peak_check_batch_impl(OpmHitMap* const* maps, const int32_t* maps_per_check, int32_t n_checks, const int64_t* bounce_filters,
                                       uint64_t nx, uint64_t ny, void* acc_dev, double* maps_dev, double* const* results_dev, bool side) {
    if (n_checks == 0) return OPM_OK;
    if (n_checks < 0) return fail(OPM_ERR_ARG, "negative number of checks");
    int64_t total_maps = 0;
    int max_maps = 1;
    for (int k = 0; k < n_checks; ++k) {
        const int m = maps_per_check ? maps_per_check[k] : 1;
        if (m < 1 || m > kPeakCheckMaps) return fail(OPM_ERR_ARG, "peak check batch: a check takes 1 to %d hit maps", kPeakCheckMaps);
        total_maps += m;
        if (m > max_maps) max_maps = m;
    }
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, (int32_t)total_maps, &c);
    if (s) return s;
    if (!acc_dev || !maps_dev || !results_dev || nx < 2 || ny < 2 || nx * ny * sizeof(double) > 160 * 1024 || n_checks > 65535)
        return fail(OPM_ERR_ARG, "peak check batch: bad arguments (maps up to 160 KB, at most 65535 checks)");
    CU(cudaSetDevice(c->device));
    if (n_checks > c->jobs_cap) {
        if (c->jobs_cap) { CU(cudaStreamSynchronize(c->stream)); if (c->side_stream) CU(cudaStreamSynchronize(c->side_stream)); }  // lists in flight
        cudaFreeHost(c->h_jobs); cudaFree(c->d_jobs);
        c->h_jobs = nullptr; c->d_jobs = nullptr; c->jobs_cap = 0;
        const int cap = n_checks < 256 ? 256 : n_checks;
        CU(cudaMallocHost(&c->h_jobs, sizeof(PeakCheckJob) * cap * OpmContext::kJobRing));
        CU(cudaMalloc(&c->d_jobs, sizeof(PeakCheckJob) * cap * OpmContext::kJobRing));
        for (int k = 0; k < OpmContext::kJobRing; ++k) if (!c->jobs_ev[k]) CU(cudaEventCreateWithFlags(&c->jobs_ev[k], cudaEventDisableTiming));
        c->jobs_cap = cap;
    }
    const int ring = c->jobs_pos;
    c->jobs_pos = (c->jobs_pos + 1) % OpmContext::kJobRing;
    CU(cudaEventSynchronize(c->jobs_ev[ring]));  // the batch that last used this slot of the ring has finished with its lists
    PeakCheckJob* const h_list = c->h_jobs + (size_t)ring * c->jobs_cap;
    PeakCheckJob* const d_list = c->d_jobs + (size_t)ring * c->jobs_cap;
    const size_t map_stride = (nx * ny + 7) / 8 * 8;
    uint64_t max_n = 0;
    int64_t next_map = 0;
    for (int k = 0; k < n_checks; ++k) {
        if (!results_dev[k]) return fail(OPM_ERR_ARG, "peak check batch: NULL result slot");
        PeakCheckJob& j = h_list[k];
        memset(&j, 0, sizeof j);
        j.n_maps = maps_per_check ? maps_per_check[k] : 1;
        for (int m = 0; m < j.n_maps; ++m, ++next_map) {
            j.H[m] = maps[next_map]->v; j.n[m] = maps[next_map]->len;
            if (j.n[m] > max_n) max_n = j.n[m];
        }
        const int64_t f = bounce_filters ? bounce_filters[k] : OPM_BOUNCE_OF_LAST_LAUNCH;
        if (f == OPM_BOUNCE_OF_LAST_LAUNCH) j.bounce_filter = -(long long)(uintptr_t)(reinterpret_cast<unsigned long long*>(results_dev[k]) + 7);
        else if (f < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
        else j.bounce_filter = f;
        j.acc = (PeakCheckAcc*)acc_dev + k; j.result = results_dev[k]; j.map = maps_dev + (size_t)k * map_stride;
    }
    s = flush_pending_keys(c);
    if (s) return s;
    cudaStream_t st = c->stream;
    if (side) {  // behind everything the main stream holds so far (the traces that recorded the hits, the captured keys), beside what follows
        if (!c->side_stream) { CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&c->side_ev, cudaEventDisableTiming)); }
        CU(cudaEventRecord(c->side_ev, c->stream));
        CU(cudaStreamWaitEvent(c->side_stream, c->side_ev, 0));
        st = c->side_stream;
        c->side_busy = true;
    }
    CU(cudaMemcpyAsync(d_list, h_list, sizeof(PeakCheckJob) * n_checks, cudaMemcpyHostToDevice, st));
    int active[kPeakCheckMaps] = {};
    for (int k = 0; k < n_checks; ++k)
        for (int m = 0; m < h_list[k].n_maps; ++m) active[m]++;
    KL(k_peakcheck_batch(d_list, n_checks, max_maps, active, max_n, nx, ny, c->sms, st));
    CU(cudaEventRecord(c->jobs_ev[ring], st));  // both lists of this slot are free once the batch's last kernel has run
    c->launches += 2 * max_maps + 1;
    return OPM_OK;
}
extern "C" OpmStatus opm_fluence_peak_total_async(OpmContext* c, const double* map_dev, uint64_t nx, uint64_t ny, const double* range_dev,
                                                  double* peak_total_dev) {
    if (!c || !map_dev || !range_dev || !peak_total_dev) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    PeakAccum init;
    init.peak_key = double_to_key(-INFINITY); init.total = 0.0;
    PeakAccum* d_acc = (PeakAccum*)c->d_scratch;
    { OpmStatus ps = dev_put(c, d_acc, &init, sizeof init); if (ps) return ps; }
    KL(k_peak_total(map_dev, nx, ny, range_dev, d_acc, peak_total_dev, c->sms, c->stream)); c->launches += 2;
    return OPM_OK;
}

// ---- synchronous forms (one host synchronisation each) ----------------------------------------------------------
static OpmStatus range_to_host(OpmContext* c, const double* range_dev, double lrtb[4], const unsigned* status_dev, unsigned* status) {
    double* h = (double*)(c->h_scratch + 2048);
    CU(cudaMemcpyAsync(h, range_dev, 32, cudaMemcpyDeviceToHost, c->stream));
    unsigned* hs = (unsigned*)(c->h_scratch + 1024);
    if (status_dev) CU(cudaMemcpyAsync(hs, status_dev, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    lrtb[0] = -h[0]; lrtb[1] = h[1]; lrtb[2] = h[2]; lrtb[3] = -h[3];
    if (status) *status = status_dev ? *hs : 0u;
    return OPM_OK;
}
extern "C" OpmStatus opm_hitmaps_bounding_box(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, double lrtb[4],
                                              uint64_t* n_hits) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    const BBoxAccum* d_used = nullptr;
    s = bounding_box_async(maps, n_maps, bounce_filter, scratch_range(c), &d_used);
    if (s) return s;
    CU(cudaMemcpyAsync(c->h_scratch, d_used, sizeof(BBoxAccum), cudaMemcpyDeviceToHost, c->stream));
    double box[4];
    s = range_to_host(c, scratch_range(c), box, nullptr, nullptr);
    if (s) return s;
    const BBoxAccum& a = *(BBoxAccum*)c->h_scratch;
    if (n_hits) *n_hits = a.count;
    if (a.count == 0) return fail(OPM_ERR_EMPTY_HITMAP, "%s", opm_status_string(OPM_ERR_EMPTY_HITMAP));
    memcpy(lrtb, box, sizeof box);
    return OPM_OK;
}
extern "C" OpmStatus opm_fluence_binning(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                         const double* ranges, const double* lrtb_override, int32_t accumulate, double* map_dev,
                                         double lrtb_out[4]) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (!map_dev || nx == 0 || ny == 0) return fail(OPM_ERR_ARG, "bad map arguments");
    CU(cudaSetDevice(c->device));
    const double* given = ranges ? ranges : lrtb_override;  // rays_hit_map.rs:337-339: (left, right, top, bottom) (sic)
    if (given) {
        double enc[4] = {-given[0], given[1], given[2], -given[3]};
        s = dev_put(c, scratch_range(c), enc, sizeof enc);
    } else {
        s = opm_hitmaps_bounding_box_async(maps, n_maps, bounce_filter, scratch_range(c));
    }
    if (s) return s;
    unsigned* d_status = (unsigned*)(c->d_scratch + 1024);
    CU(cudaMemsetAsync(d_status, 0, 4, c->stream));
    s = opm_fluence_binning_async(maps, n_maps, bounce_filter, nx, ny, scratch_range(c), accumulate, map_dev, d_status);
    if (s) return s;
    double box[4];
    unsigned status = 0;
    s = range_to_host(c, scratch_range(c), box, d_status, &status);
    if (s) return s;
    if (!given && !(std::isfinite(box[0]) && std::isfinite(box[1]) && std::isfinite(box[2]) && std::isfinite(box[3])))
        return fail(OPM_ERR_EMPTY_HITMAP, "%s", opm_status_string(OPM_ERR_EMPTY_HITMAP));
    if (lrtb_out) memcpy(lrtb_out, box, sizeof box);
    return status_from_bits(status);
}
extern "C" OpmStatus opm_fluence_peak_total(OpmContext* c, const double* map_dev, uint64_t nx, uint64_t ny, const double lrtb[4],
                                            double* peak, double* total) {
    if (!c || !map_dev || !lrtb) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    double enc[4] = {-lrtb[0], lrtb[1], lrtb[2], -lrtb[3]};
    OpmStatus s = dev_put(c, scratch_range(c), enc, sizeof enc);
    if (!s) s = opm_fluence_peak_total_async(c, map_dev, nx, ny, scratch_range(c), scratch_peak(c));
    if (s) return s;
    double* h = (double*)(c->h_scratch + 2304);
    CU(cudaMemcpyAsync(h, scratch_peak(c), 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (peak) *peak = h[0];
    if (total) *total = h[1];
    return OPM_OK;
}

// ---------------------------------------------------------------------------------------------
// KDE fluence estimator: RaysHitMap::calc_fluence_with_kde (rays_hit_map.rs:575-613), Kde (kde/mod.rs) -- SURVEY 8f-1
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_fluence_kde(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, uint64_t ny,
                                     const double* ranges, double* map_dev, double lrtb_out[4], double* band_width_out) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if (!map_dev || nx == 0 || ny == 0) return fail(OPM_ERR_ARG, "bad map arguments");
    CU(cudaSetDevice(c->device));
    uint64_t slots = 0;
    for (int k = 0; k < n_maps; ++k) slots += maps[k]->len;
    // dense hit list
    double *hx = nullptr, *hy = nullptr, *he = nullptr, *dist = nullptr, *sorted = nullptr;
    void* temp = nullptr;
    auto cleanup = [&]() { cudaFree(hx); cudaFree(dist); cudaFree(sorted); cudaFree(temp); };
    CU(cudaMalloc(&hx, (slots ? slots : 1) * 3 * sizeof(double)));
    hy = hx + (slots ? slots : 1); he = hy + (slots ? slots : 1);
    unsigned long long* d_count = (unsigned long long*)(c->d_scratch + 3072);  // count | sum of distances | sum of squared deviations
    double* d_sums = (double*)(d_count + 1);
    cudaMemsetAsync(d_count, 0, 24, c->stream);
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) {
            int rc = k_hits_pack(maps[k]->v, maps[k]->len, bounce_filter, hx, hy, he, d_count, slots, c->sms, c->stream);
            c->launches++;
            if (rc) { cleanup(); return fail(OPM_ERR_CUDA, "hit packing failed"); }
        }
    unsigned long long n = 0;
    cudaMemcpyAsync(&n, d_count, 8, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cleanup(); return fail(OPM_ERR_CUDA, "hit packing failed"); }
    // Kde::bandwidth_estimate (kde/mod.rs:83-106)
    const double nan = std::nan("");
    double bw = nan;
    std::vector<double> first(4, 0.0);
    if (n == 2) {
        cudaMemcpyAsync(first.data(), hx, 16, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(first.data() + 2, hy, 16, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        double dx = first[1] - first[0], dy = first[3] - first[2];
        bw = 0.5 * std::sqrt(dx * dx + dy * dy);
        if (bw == 0.0) bw = nan;
    } else if (n > 2) {
        if (n > 50000) { cleanup(); return fail(OPM_ERR_CAPACITY, "KDE bandwidth estimate over %llu hit points needs %llu pair distances", n, n * (n - 1) / 2); }
        const uint64_t m = n * (n - 1) / 2;
        size_t temp_bytes = 0;
        k_sort_doubles(nullptr, nullptr, m, nullptr, &temp_bytes, c->stream);
        if (cudaMalloc(&dist, m * 8) != cudaSuccess || cudaMalloc(&sorted, m * 8) != cudaSuccess || cudaMalloc(&temp, temp_bytes ? temp_bytes : 8) != cudaSuccess) {
            cleanup(); cudaGetLastError();
            return fail(OPM_ERR_CUDA, "out of device memory for the KDE pair distances");
        }
        int rc = k_kde_pair_distances(hx, hy, n, dist, d_sums, d_sums + 1, c->sms, c->stream);
        if (!rc) rc = k_sort_doubles(dist, sorted, m, temp, &temp_bytes, c->stream);
        c->launches += 3;
        if (rc) { cleanup(); return fail(OPM_ERR_CUDA, "KDE bandwidth kernels failed: %s", cudaGetErrorString((cudaError_t)rc)); }
        // Kde::distances_iqr (kde/mod.rs:65-81)
        const double quart = 0.75 * (double)m;
        double ip;
        const bool whole = std::modf(quart, &ip) == 0.0;
        const uint64_t q = whole ? (uint64_t)quart - 1 : (uint64_t)std::floor(quart);
        double qv[2] = {0.0, 0.0}, sums[2] = {0.0, 0.0};
        cudaMemcpyAsync(qv, sorted + q, whole ? 16 : 8, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(sums, d_sums, 16, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cleanup(); return fail(OPM_ERR_CUDA, "KDE bandwidth kernels failed"); }
        const double iqr = whole ? 0.5 * (qv[0] + qv[1]) : qv[0];
        const double std_dev = std::sqrt(sums[1] / (double)m);
        bw = 0.9 * std::fmin(std_dev, iqr / 1.34) * std::pow((double)n, -0.2);  // Silverman's rule of thumb
    }
    if (band_width_out) *band_width_out = bw;
    if (!std::isnormal(bw)) { cleanup(); return fail(OPM_ERR_KDE_BANDWIDTH, "%s", opm_status_string(OPM_ERR_KDE_BANDWIDTH)); }
    // map range: given ranges (left, right, top, bottom like the reference reads them) or bounding box + 3 bandwidths
    double l, r, t, b;
    if (ranges) { l = ranges[0]; r = ranges[1]; t = ranges[2]; b = ranges[3]; }
    else {
        double box[4];
        s = opm_hitmaps_bounding_box(maps, n_maps, bounce_filter, box, nullptr);
        if (s) { cleanup(); return s; }
        const double margin = 3. * bw;
        l = box[0] - margin; r = box[1] + margin; b = box[3] - margin; t = box[2] + margin;
    }
    const double dx = (r - l) / (double)nx, dy = (t - b) / (double)ny;
    int rc = k_kde_2d(hx, hy, he, n, bw, l, b, dx, dy, nx, ny, map_dev, c->stream);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (rc || e != cudaSuccess) return fail(OPM_ERR_CUDA, "KDE map kernel failed");
    if (lrtb_out) { lrtb_out[0] = l; lrtb_out[1] = r; lrtb_out[2] = t; lrtb_out[3] = b; }
    return OPM_OK;
}

// ---- Voronoi fluence estimator (opm_voronoi.cu): RaysHitMap::calc_fluence_with_voronoi, rays_hit_map.rs:407-521
extern "C" OpmStatus opm_hitmap_bounce_levels(OpmHitMap* map, uint64_t* mask) {
    if (!map || !mask) return fail(OPM_ERR_ARG, "NULL argument");
    OpmContext* c = map->ctx;
    CU(cudaSetDevice(c->device));
    unsigned long long* d = (unsigned long long*)(c->d_scratch + 3072);
    CU(cudaMemsetAsync(d, 0, 8, c->stream));
    if (map->len) { KL(k_hits_bounce_mask(map->v, map->len, d, c->sms, c->stream)); c->launches++; }
    unsigned long long* h = (unsigned long long*)(c->h_scratch + 3072);
    CU(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *mask = *h;
    return OPM_OK;
}
struct VorSites { std::vector<double> e, z; };  // sorted sites: what went in as `e`, and e / cell area [cm^2]
static OpmStatus fluence_voronoi_impl(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, const double* ranges,
                                      int32_t accumulate, double* map_dev, double lrtb_out[4], bool helper_rays, VorSites* cells_only = nullptr);
// create_voronoi_cells (utils/griddata.rs:265-345) + calc_closed_poly_area of every cell, as Rays::new_collimated_w_fluence_helper
// uses them (rays.rs:367-389): the area [m^2] of each point's Voronoi cell inside the pushed-out convex hull; NaN = fewer than
// three vertices
extern "C" OpmStatus opm_voronoi_cell_areas(OpmContext* c, const double* x_m, const double* y_m, uint64_t n, double* area_m2_out) {
    if (!c || !x_m || !y_m || !area_m2_out) return fail(OPM_ERR_ARG, "NULL argument");
    OpmHitMap* hm = nullptr;
    OpmStatus s = opm_hitmap_create(c, n ? n : 1, &hm);
    if (s) return s;
    std::vector<double> tag(n);
    for (uint64_t i = 0; i < n; ++i) tag[i] = (double)(i + 1);  // the sites come back sorted: the "energy" carries the index
    s = opm_hitmap_upload(hm, x_m, y_m, nullptr, tag.data(), nullptr, n);
    VorSites sites;
    OpmHitMap* maps[1] = {hm};
    if (!s) s = fluence_voronoi_impl(maps, 1, -1, 2, nullptr, 0, nullptr, nullptr, false, &sites);
    opm_hitmap_destroy(hm);
    if (s) return s;
    if (sites.e.size() != n) return fail(OPM_ERR_VORONOI, "Voronoi cells: %zu sites for %llu points", sites.e.size(), (unsigned long long)n);
    for (uint64_t k = 0; k < n; ++k) {
        const uint64_t i = (uint64_t)sites.e[k] - 1;
        area_m2_out[i] = (sites.e[k] / sites.z[k]) * 1.0E-4;  // cm^2 -> m^2
    }
    return OPM_OK;
}
// Rays::new_collimated_w_fluence_helper (rays.rs:352-398): rays along +z at the given positions, energy = Voronoi cell area x
// fluence (0 for a cell of fewer than three vertices), each with its FluenceRays
extern "C" OpmStatus opm_rays_new_collimated_w_fluence_helper(OpmRays* rays, const double* pos_xyz, uint64_t n, double wavelength_m,
                                                              const double* fluence) {
    if (!rays || !pos_xyz || !fluence || n == 0) return fail(OPM_ERR_ARG, "NULL argument");
    std::vector<double> x(n), y(n), z(n), dx(n, 0.0), dy(n, 0.0), dz(n, 1.0), e(n), w(n, wavelength_m), area(n);
    for (uint64_t i = 0; i < n; ++i) { x[i] = pos_xyz[3 * i]; y[i] = pos_xyz[3 * i + 1]; z[i] = pos_xyz[3 * i + 2]; }
    OpmStatus s = opm_voronoi_cell_areas(rays->ctx, x.data(), y.data(), n, area.data());
    if (s) return s;
    for (uint64_t i = 0; i < n; ++i) e[i] = std::isfinite(area[i]) ? area[i] * fluence[i] : 0.0;
    const double* P[3] = {x.data(), y.data(), z.data()};
    const double* D[3] = {dx.data(), dy.data(), dz.data()};
    s = opm_rays_upload_soa(rays, P, D, e.data(), w.data(), n);
    if (s) return s;
    CU(cudaStreamSynchronize(rays->ctx->stream));  // the staging vectors go out of scope
    return opm_rays_attach_fluence_helpers(rays, fluence);
}
extern "C" OpmStatus opm_fluence_voronoi(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, const double* ranges,
                                         int32_t accumulate, double* map_dev, double lrtb_out[4]) {
    return fluence_voronoi_impl(maps, n_maps, bounce_filter, nx, ranges, accumulate, map_dev, lrtb_out, false);
}
extern "C" OpmStatus opm_fluence_helper_rays(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, const double* ranges,
                                             int32_t accumulate, double* map_dev, double lrtb_out[4]) {
    return fluence_voronoi_impl(maps, n_maps, bounce_filter, nx, ranges, accumulate, map_dev, lrtb_out, true);
}
// helper_rays: RaysHitMap::calc_fluence_with_helper_rays (rays_hit_map.rs:627-722) -- the hit records carry FLUENCES (the trace of a
// bundle with fluence helper rays wrote them), the sites of the diagram take them as they are, voronator's four helper sites around
// the bounding polygon take the value 0 (:700-701) and ALL of them go into the natural-neighbour interpolation
static OpmStatus fluence_voronoi_impl(OpmHitMap* const* maps, int32_t n_maps, int64_t bounce_filter, uint64_t nx, const double* ranges,
                                      int32_t accumulate, double* map_dev, double lrtb_out[4], bool helper_rays, VorSites* cells_only) {
    OpmContext* c;
    OpmStatus s = check_same_ctx(maps, n_maps, &c);
    if (s) return s;
    if ((!map_dev && !cells_only) || nx == 0 || nx > 46340) return fail(OPM_ERR_ARG, "bad map arguments");
    if (bounce_filter < -1) return fail(OPM_ERR_ARG, "bad bounce filter");
    CU(cudaSetDevice(c->device));
    uint64_t slots = 0;
    for (int k = 0; k < n_maps; ++k) slots += maps[k]->len;
    if (slots > 0x7fffffffull) return fail(OPM_ERR_CAPACITY, "Voronoi estimator: too many hit slots");
    std::vector<void*> dev;  // everything allocated here
    auto cleanup = [&]() { cudaStreamSynchronize(c->stream); for (void* p : dev) cudaFree(p); };
    auto alloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return nullptr; } dev.push_back(p); return p; };
    // dense hit list
    const size_t cap = slots ? slots : 1;
    double* hx = (double*)alloc(cap * 4 * sizeof(double));
    if (!hx) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); }
    double *hy = hx + cap, *he = hy + cap, *fl = he + cap;
    unsigned long long* d_count = (unsigned long long*)(c->d_scratch + 3072);
    unsigned* d_status = (unsigned*)(c->d_scratch + 3072 + 8);
    cudaMemsetAsync(d_count, 0, 16, c->stream);
    for (int k = 0; k < n_maps; ++k)
        if (maps[k]->len) {
            int rc = k_hits_pack(maps[k]->v, maps[k]->len, bounce_filter, hx, hy, he, d_count, slots, c->sms, c->stream);
            c->launches++;
            if (rc) { cleanup(); return fail(OPM_ERR_CUDA, "hit packing failed"); }
        }
    unsigned long long n64 = 0;
    cudaMemcpyAsync(&n64, d_count, 8, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cleanup(); return fail(OPM_ERR_CUDA, "hit packing failed"); }
    const int n = (int)n64;
    if (n < 3) { cleanup(); return fail(OPM_ERR_VORONOI, "Too few points (<3) on hitmap to calculate fluence!"); }
    // the warp-aggregated packing does not keep the slot order; the estimate is a function of the SET of hits, but the last bits of
    // the sums are not: sort the hits by (x, y, e) on the host so that equal hit sets give bit-equal maps
    KL(k_voronoi_to_cm(n, hx, hy, c->stream)); c->launches++;
    std::vector<double> x(n), y(n), e(n);
    CU(cudaMemcpyAsync(x.data(), hx, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(y.data(), hy, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(e.data(), he, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    {
        std::vector<int> ord(n);
        for (int i = 0; i < n; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](int a, int b) { return x[a] != x[b] ? x[a] < x[b] : (y[a] != y[b] ? y[a] < y[b] : e[a] < e[b]); });
        std::vector<double> t(n);
        for (int i = 0; i < n; ++i) t[i] = x[ord[i]]; x.swap(t);
        for (int i = 0; i < n; ++i) t[i] = y[ord[i]]; y.swap(t);
        for (int i = 0; i < n; ++i) t[i] = e[ord[i]]; e.swap(t);
        CU(cudaMemcpyAsync(hx, x.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(hy, y.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(he, e.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    }
    for (int i = 0; i < n; ++i)
        if (!std::isfinite(x[i]) || !std::isfinite(y[i])) { cleanup(); return fail(OPM_ERR_VORONOI, "Coordinate values for voronoi-diagram generation must be finite!"); }
    if (vor_all_on_same_line(x.data(), y.data(), n)) { cleanup(); return fail(OPM_ERR_VORONOI, "All passed points lie on the same line! Triangulation not possible!"); }
    // convex hull, bounding polygon (create_voronoi_cells, griddata.rs:284-343), grid
    std::vector<int> hull;
    int on_b = 0;
    vor_convex_hull(x.data(), y.data(), n, hull, &on_b);
    const int h = (int)hull.size();
    if (h < 3) { cleanup(); return fail(OPM_ERR_VORONOI, "Could not create voronoi diagram!"); }
    std::vector<double> px, py;
    double pcx = 0.0, pcy = 0.0;
    vor_bounding_polygon(x.data(), y.data(), hull, 2 * n - 2 - on_b, px, py, &pcx, &pcy);
    const int m = (int)px.size();
    std::vector<double> pol(4 * (size_t)m);  // px | py | outward normal x | y
    double area2 = 0.0;
    for (int i = 0; i < m; ++i) { const int j = (i + 1) % m; area2 += px[i] * py[j] - py[i] * px[j]; }
    const double orient = area2 >= 0.0 ? 1.0 : -1.0;  // counter-clockwise: the interior is on the left of every edge
    VorBound B;
    memset(&B, 0, sizeof B);
    B.x0 = B.x1 = px[0]; B.y0 = B.y1 = py[0];
    double r_in = INFINITY;
    for (int i = 0; i < m; ++i) {
        const int j = (i + 1) % m;
        const double ex = px[j] - px[i], ey = py[j] - py[i];
        const double nxo = orient * ey, nyo = -orient * ex;  // outward normal of a counter-clockwise edge (ex, ey) is (ey, -ex)
        pol[i] = px[i]; pol[m + i] = py[i]; pol[2 * m + i] = nxo; pol[3 * m + i] = nyo;
        const double len = std::sqrt(nxo * nxo + nyo * nyo);
        if (len > 0.0) r_in = std::min(r_in, ((px[i] - pcx) * nxo + (py[i] - pcy) * nyo) / len);
        B.x0 = std::min(B.x0, px[i]); B.x1 = std::max(B.x1, px[i]); B.y0 = std::min(B.y0, py[i]); B.y1 = std::max(B.y1, py[i]);
    }
    r_in = std::isfinite(r_in) && r_in > 0.0 ? r_in * (1.0 - 1e-9) : 0.0;
    B.cx = pcx; B.cy = pcy; B.r_in2 = r_in * r_in; B.m = m;
    VorGrid G;
    std::vector<int> cell_start, cell_sites;
    vor_build_grid(x.data(), y.data(), n, G, cell_start, cell_sites);
    double* d_pol = (double*)alloc(pol.size() * 8);
    int* d_cells = (int*)alloc((cell_start.size() + cell_sites.size()) * 4);
    if (!d_pol || !d_cells) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); }
    CU(cudaMemcpyAsync(d_pol, pol.data(), pol.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_cells, cell_start.data(), cell_start.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_cells + cell_start.size(), cell_sites.data(), cell_sites.size() * 4, cudaMemcpyHostToDevice, c->stream));
    B.px = d_pol; B.py = d_pol + m; B.nx = d_pol + 2 * m; B.ny = d_pol + 3 * m;
    G.cell_start = d_cells; G.cell_sites = d_cells + cell_start.size();
    std::vector<double> z(n);
    unsigned st_bits = 0;
    if (!helper_rays) {
        // cells -> per-site fluence
        KL(k_voronoi_cells(n, hx, hy, he, G, B, fl, d_status, c->stream)); c->launches++;
        CU(cudaMemcpyAsync(z.data(), fl, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&st_bits, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (st_bits) { cleanup(); return fail(OPM_ERR_CAPACITY, "Voronoi estimator: a cell has more than 64 vertices"); }
    } else {
        for (int i = 0; i < n; ++i) z[i] = e[i] / 1.0E4;  // get::<joule_per_square_centimeter>() of the SI value
    }
    if (cells_only) { cells_only->e = e; cells_only->z = z; cleanup(); return OPM_OK; }
    // interpolation over the sites with a finite value (griddata.rs:412): normally all of them
    int nf = 0;
    for (int i = 0; i < n; ++i) nf += std::isfinite(z[i]) ? 1 : 0;
    const double *ix = hx, *iy = hy, *iz = fl;
    std::vector<double> fx, fy, fz;
    if (nf != n || helper_rays) {
        for (int i = 0; i < n; ++i) if (std::isfinite(z[i])) { fx.push_back(x[i]); fy.push_back(y[i]); fz.push_back(z[i]); }
        if (helper_rays) {
            // voronator 0.2 VoronoiDiagram::with_bounding_polygon appends four helper sites to the point list (restated from the
            // crate's source as remembered -- it is not under /root/reference): left, right, below and above the box of the
            // bounding polygon, one box width / height away.  `sites.len() - hit_points.len()` of rays_hit_map.rs:701 is this 4.
            const double w = B.x1 - B.x0, hgt = B.y1 - B.y0;
            const double ex[4] = {B.x0 - w, B.x1 + w, B.x0 + w / 2.0, B.x0 + w / 2.0};
            const double ey[4] = {B.y0 + hgt / 2.0, B.y0 + hgt / 2.0, B.y0 - hgt, B.y1 + hgt};
            for (int k = 0; k < 4; ++k) { fx.push_back(ex[k]); fy.push_back(ey[k]); fz.push_back(0.0); }
            nf = (int)fx.size();
        }
        double* d_f = (double*)alloc((size_t)std::max(nf, 1) * 3 * 8);
        if (!d_f) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); }
        if (nf) {
            CU(cudaMemcpyAsync(d_f, fx.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_f + nf, fy.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_f + 2 * (size_t)nf, fz.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream));
        }
        ix = d_f; iy = d_f + nf; iz = d_f + 2 * (size_t)nf;
        hull.clear();
        if (nf >= 3) {
            vor_convex_hull(fx.data(), fy.data(), nf, hull, nullptr);
            vor_build_grid(fx.data(), fy.data(), nf, G, cell_start, cell_sites);
            int* d_c2 = (int*)alloc((cell_start.size() + cell_sites.size()) * 4);
            if (!d_c2) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); }
            CU(cudaMemcpyAsync(d_c2, cell_start.data(), cell_start.size() * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_c2 + cell_start.size(), cell_sites.data(), cell_sites.size() * 4, cudaMemcpyHostToDevice, c->stream));
            G.cell_start = d_c2; G.cell_sites = d_c2 + cell_start.size();
        }
    }
    const std::vector<double>& sxh = nf != n ? fx : x;
    const std::vector<double>& syh = nf != n ? fy : y;
    const std::vector<double>& szh = nf != n ? fz : z;
    const int hh = (int)hull.size();
    std::vector<double> hv(3 * (size_t)std::max(hh, 1));
    for (int i = 0; i < hh; ++i) { hv[i] = sxh[hull[i]]; hv[hh + i] = syh[hull[i]]; hv[2 * hh + i] = szh[hull[i]]; }
    double* d_hull = (double*)alloc(hv.size() * 8);
    if (!d_hull) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); }
    CU(cudaMemcpyAsync(d_hull, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice, c->stream));
    VorHull H{hh >= 3 ? hh : 0, d_hull, d_hull + hh, d_hull + 2 * hh};
    // axes (rays_hit_map.rs:459-497): given ranges or the hits' limits, in centimetres; nx points on BOTH axes
    double x0, x1, y0, y1;
    if (ranges) { x0 = ranges[0] / 1.0E-2; x1 = ranges[1] / 1.0E-2; y0 = ranges[2] / 1.0E-2; y1 = ranges[3] / 1.0E-2; }
    else {
        x0 = y0 = INFINITY; x1 = y1 = -INFINITY;
        for (int i = 0; i < n; ++i) { x0 = std::fmin(x0, x[i]); x1 = std::fmax(x1, x[i]); y0 = std::fmin(y0, y[i]); y1 = std::fmax(y1, y[i]); }
    }
    if (!std::isfinite(x0) || !std::isfinite(x1) || !std::isfinite(y0) || !std::isfinite(y1)) { cleanup(); return fail(OPM_ERR_VORONOI, "cannot construct voronoi cells with non-finite axes bounds!"); }
    const double bin_x = nx > 1 ? (x1 - x0) / (double)(nx - 1) : 0.0, bin_y = nx > 1 ? (y1 - y0) / (double)(nx - 1) : 0.0;
    double ext = 0.0;
    for (int i = 0; i < (int)sxh.size(); ++i) ext = std::fmax(ext, std::fmax(std::fabs(sxh[i]), std::fabs(syh[i])));
    const double big = 1.0e3 * (std::fmax(x1 - x0, y1 - y0) + ext + 1.0);
    double* target = map_dev;
    if (accumulate) { target = (double*)alloc((size_t)nx * nx * 8); if (!target) { cleanup(); return fail(OPM_ERR_CUDA, "out of device memory"); } }
    CU(cudaMemsetAsync(d_status, 0, 4, c->stream));
    KL(k_voronoi_interp((int)nx, (int)nx, x0, bin_x, y0, bin_y, nf, ix, iy, iz, G, H, big, target, d_status, c->stream)); c->launches++;
    if (accumulate) { KL(k_voronoi_accumulate(nx * nx, map_dev, target, c->sms, c->stream)); c->launches++; }
    CU(cudaMemcpyAsync(&st_bits, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    cudaError_t ce = cudaStreamSynchronize(c->stream);
    cleanup();
    if (ce != cudaSuccess) return fail(OPM_ERR_CUDA, "Voronoi kernels failed: %s", cudaGetErrorString(ce));
    if (st_bits) return fail(OPM_ERR_CAPACITY, "Voronoi estimator: a natural-neighbour cell exceeds the polygon capacity");
    if (lrtb_out) {
        if (ranges) { lrtb_out[0] = ranges[0]; lrtb_out[1] = ranges[1]; lrtb_out[2] = ranges[3]; lrtb_out[3] = ranges[2]; }
        else { lrtb_out[0] = x0 * 1.0E-2; lrtb_out[1] = x1 * 1.0E-2; lrtb_out[2] = y1 * 1.0E-2; lrtb_out[3] = y0 * 1.0E-2; }
    }
    return OPM_OK;
}

// measured FP64 FMA throughput of this device (TFLOP/s, FMA = 2 flop): the roofline denominator of the fused kernel
extern "C" OpmStatus opm_measure_fp64_peak(OpmContext* c, double* tflops) {
    if (!c || !tflops) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const int iters = 1 << 15, blocks = c->sms * 8;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(a, c->stream));
        KL(k_fp64_peak((double*)c->d_scratch, iters, blocks, c->stream)); c->launches++;
        CU(cudaEventRecord(b, c->stream));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        double tf = (double)blocks * 256.0 * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    return OPM_OK;
}
extern "C" OpmStatus opm_device_alloc(OpmContext* c, uint64_t bytes, void** out) {
    if (!c || !out) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return OPM_OK;
}
extern "C" void opm_device_free(OpmContext* c, void* p) {
    if (!c || !p) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(p);
}
extern "C" OpmStatus opm_device_memset(OpmContext* c, void* dev, int32_t value, uint64_t bytes) {  // asynchronous
    if (!c || !dev) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(dev, value, bytes, c->stream));
    return OPM_OK;
}
extern "C" OpmStatus opm_device_to_host(OpmContext* c, void* host, const void* dev, uint64_t bytes) {
    if (!c) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return OPM_OK;
}
extern "C" OpmStatus opm_host_to_device(OpmContext* c, void* dev, const void* host, uint64_t bytes) {
    if (!c) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return OPM_OK;
}
extern "C" OpmStatus opm_host_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return fail(OPM_ERR_ARG, "NULL argument");
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return OPM_OK;
}
extern "C" void opm_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------------
// sources
// ---------------------------------------------------------------------------------------------
// spectral_distribution/gaussian.rs:78-96 (+ linspace utils/griddata.rs:211-231, gaussian() math_distribution_functions.rs:168-176):
// `num` wavelengths equally spaced over [lo, hi], generalised-Gaussian weights normalised (Kahan sum) to 1.  Host math.
extern "C" OpmStatus opm_spectrum_gaussian(double lo, double hi, uint64_t num, double mu, double fwhm, double power, double* wvl,
                                           double* weights) {
    if (!wvl || !weights || num == 0) return fail(OPM_ERR_ARG, "NULL argument");
    if (!std::isfinite(lo) || !std::isfinite(hi)) return fail(OPM_ERR_ARG, "wavelength range must be finite");
    const double bin = num > 1 ? (hi - lo) / (double)(num - 1) : 0.0;
    for (uint64_t i = 0; i < num; ++i) wvl[i] = lo + (double)i * bin;
    const double sigma = fwhm / (2.0 * std::sqrt(2.0 * std::pow(std::log(2.0), 1.0 / power)));
    double sum = 0.0, comp = 0.0;  // kahan 0.1.4
    for (uint64_t i = 0; i < num; ++i) {
        const double u = (wvl[i] - mu) / sigma;
        weights[i] = std::exp(-std::pow(0.5 * (u * u), power));
        const double y = weights[i] - comp, t = sum + y;
        comp = (t - sum) - y;
        sum = t;
    }
    for (uint64_t i = 0; i < num; ++i) weights[i] = weights[i] / sum;
    return OPM_OK;
}

static OpmStatus source_params(const OpmSourceDesc* d, SourceParams& S) {
    memset((void*)&S, 0, sizeof S);
    if (!d) return fail(OPM_ERR_ARG, "desc is NULL");
    if (d->n_positions == 0) return fail(OPM_ERR_ARG, "source has no positions");
    S.pos_dist = d->pos_dist; S.energy_dist = d->energy_dist; S.rectangular = d->rectangular;
    S.n_positions = d->n_positions; S.n_wavelengths = d->n_wavelengths ? d->n_wavelengths : 1;
    S.size_x = d->size_x; S.size_y = d->size_y;
    S.golden = (1.0 + std::sqrt(5.0)) / 2.0;  // f64::midpoint(1., sqrt(5.)) fibonacci.rs:56
    if (d->pos_dist == OPM_POS_GRID) {  // grid.rs:58-80
        uint64_t nx = d->grid_nx < 1 ? 1 : d->grid_nx, ny = d->grid_ny < 1 ? 1 : d->grid_ny;
        if (nx * ny != d->n_positions) return fail(OPM_ERR_ARG, "grid_nx*grid_ny != n_positions");
        S.grid_ny = ny;
        S.dist_x = nx > 1 ? d->size_x / (double)(nx - 1) : 0.0;
        S.dist_y = ny > 1 ? d->size_y / (double)(ny - 1) : 0.0;
        S.off_x = nx > 1 ? d->size_x / 2.0 : 0.0;
        S.off_y = ny > 1 ? d->size_y / 2.0 : 0.0;
    }
    if (d->pos_dist == OPM_POS_HEXAPOLAR) {  // hexapolar.rs:24-36,38-56
        const uint64_t rings = d->grid_nx;
        if (rings > 255) return fail(OPM_ERR_ARG, "hexapolar: nr_of_rings is a u8");
        if (std::signbit(d->size_x) || !std::isfinite(d->size_x)) return fail(OPM_ERR_ARG, "radius must be positive and finite");
        const uint64_t expect = d->size_x == 0.0 ? 1 : 1 + 3 * rings * (rings + 1);
        if (expect != d->n_positions) return fail(OPM_ERR_ARG, "hexapolar: n_positions must be 1 + 3 rings (rings + 1)");
        S.dist_x = rings ? d->size_x / (double)rings : 0.0;  // radius_step
    }
    S.total_energy = d->total_energy;
    S.e_uniform = d->total_energy / (double)d->n_positions;  // uniform.rs:39
    S.energy_norm = d->energy_norm;
    S.mu[0] = d->mu[0]; S.mu[1] = d->mu[1]; S.sigma[0] = d->sigma[0]; S.sigma[1] = d->sigma[1]; S.power = d->power;
    S.sin_t = std::sin(d->theta); S.cos_t = std::cos(d->theta);
    to_dev_iso(d->iso.fwd, S.iso);
    return OPM_OK;
}
extern "C" OpmStatus opm_source_energy_norm(OpmContext* c, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t pos_end, double* partial) {
    if (!c || !partial) return fail(OPM_ERR_ARG, "NULL argument");
    SourceParams S;
    OpmStatus s = source_params(desc, S);
    if (s) return s;
    if (pos_end > desc->n_positions || pos_begin > pos_end) return fail(OPM_ERR_ARG, "position range out of bounds");
    CU(cudaSetDevice(c->device));
    double* d_sum = (double*)c->d_scratch;
    CU(cudaMemsetAsync(d_sum, 0, 8, c->stream));
    if (pos_end > pos_begin) { KL(k_source_norm(S, pos_begin, pos_end, d_sum, c->sms, c->stream)); c->launches++; }
    CU(cudaMemcpyAsync(c->h_scratch, d_sum, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *partial = *(double*)c->h_scratch;
    return OPM_OK;
}
static OpmStatus source_generate(OpmRays* rays, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t n_pos, uint64_t stride);
extern "C" OpmStatus opm_source_generate(OpmRays* rays, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t pos_end) {
    if (pos_begin > pos_end) return fail(OPM_ERR_ARG, "position range out of bounds");
    return source_generate(rays, desc, pos_begin, pos_end - pos_begin, 1);
}
extern "C" OpmStatus opm_source_generate_strided(OpmRays* rays, const OpmSourceDesc* desc, uint64_t first_pos, uint64_t n_pos,
                                                 uint64_t stride) {
    if (stride == 0) return fail(OPM_ERR_ARG, "stride must be >= 1");
    return source_generate(rays, desc, first_pos, n_pos, stride);
}
static OpmStatus source_generate(OpmRays* rays, const OpmSourceDesc* desc, uint64_t pos_begin, uint64_t n_pos, uint64_t stride) {
    if (!rays) return fail(OPM_ERR_ARG, "rays is NULL");
    OpmContext* c = rays->ctx;
    SourceParams S;
    OpmStatus s = source_params(desc, S);
    if (s) return s;
    if (n_pos && pos_begin + (n_pos - 1) * stride >= desc->n_positions) return fail(OPM_ERR_ARG, "position range out of bounds");
    CU(cudaSetDevice(c->device));
    const bool need_norm = S.energy_dist == OPM_EDIST_GENERAL_GAUSSIAN && !(S.energy_norm > 0.0);
    const double* hw = desc->n_wavelengths ? desc->wavelengths : nullptr;
    const double* hg = desc->n_wavelengths ? desc->weights : nullptr;
    if (!hw) return fail(OPM_ERR_ARG, "source needs at least one wavelength");
    rays->pending = false;
    rays->mono_wvl = (desc->n_wavelengths == 1 && desc->wavelengths) ? desc->wavelengths[0] : 0.0; rays->spot_acc_valid = false;
    if (rays->d_hist) { OpmStatus hs = hist_clear(rays); if (hs) return hs; }  // new rays start without a history (ray.rs:126)
    uint64_t n_rays = n_pos * S.n_wavelengths;
    s = opm_rays_reserve(rays, n_rays);
    if (s) return s;
    if ((size_t)S.n_wavelengths * 16 <= OpmContext::kPutBytes) {
        // fully asynchronous: the spectrum table travels through the small-parameter ring, the normalisation sum of a Gaussian
        // energy distribution (general_gaussian.rs:107-114, taken on every apply) stays on the device
        if (!c->d_srcspec) CU(cudaMalloc(&c->d_srcspec, OpmContext::kPutBytes));
        double tab[OpmContext::kPutBytes / 8];
        for (uint32_t k = 0; k < S.n_wavelengths; ++k) { tab[k] = hw[k]; tab[S.n_wavelengths + k] = hg ? hg[k] : 1.0; }
        s = dev_put(c, c->d_srcspec, tab, (size_t)S.n_wavelengths * 16);
        if (s) return s;
        double* d_norm = (double*)(c->d_scratch + 512);
        if (need_norm) {
            CU(cudaMemsetAsync(d_norm, 0, 8, c->stream));
            KL(k_source_norm(S, 0, desc->n_positions, d_norm, c->sms, c->stream)); c->launches++;
        }
        if (n_rays) {
            KL(k_source_generate(S, rays->v, pos_begin, stride, n_rays, c->d_srcspec, c->d_srcspec + S.n_wavelengths, need_norm ? d_norm : nullptr,
                                 c->sms, c->stream));
            c->launches++;
        }
        return rays_set_len(rays, n_rays);
    }
    if (need_norm) {
        double norm = 0.0;
        s = opm_source_energy_norm(c, desc, 0, desc->n_positions, &norm);
        if (s) return s;
        S.energy_norm = norm;
    }
    double one = 1.0, *d_w = nullptr;
    CU(cudaMalloc(&d_w, (size_t)S.n_wavelengths * 16));
    CU(cudaMemcpyAsync(d_w, hw, (size_t)S.n_wavelengths * 8, cudaMemcpyHostToDevice, c->stream));
    if (hg) CU(cudaMemcpyAsync(d_w + S.n_wavelengths, hg, (size_t)S.n_wavelengths * 8, cudaMemcpyHostToDevice, c->stream));
    else CU(cudaMemcpyAsync(d_w + S.n_wavelengths, &one, 8, cudaMemcpyHostToDevice, c->stream));
    if (n_rays) { KL(k_source_generate(S, rays->v, pos_begin, stride, n_rays, d_w, d_w + S.n_wavelengths, nullptr, c->sms, c->stream)); c->launches++; }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_w);
    if (e != cudaSuccess) return fail(OPM_ERR_CUDA, "source generation failed: %s", cudaGetErrorString(e));
    return rays_set_len(rays, n_rays);
}
