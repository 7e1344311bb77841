// opm_jit.cu -- run-time specialisation of the fused surface-sequence kernel.
//
// The interpreter kernel (opm_trace.cu) pays a dispatch, a set of structural branches (which geometry? which coating?
// index model? hit record? epilogue flags?) and shared-memory descriptor loads for every op of every ray.  For a bundle
// that is large enough to amortise a compilation, the library instead generates the straight-line kernel of THIS
// program: the op bodies of opm_trace_ops.cuh called back to back with the structural facts as template arguments
// (RefractCt / ApertureCt / EpilogueCt), and the descriptors in __constant__ memory at compile-time indices, so the
// numbers (isometries, radii, coefficients) become constant-bank operands of the FP64 instructions.  Only the STRUCTURE
// is baked in: scenes that differ in numbers alone (parameter sweeps, tolerancing) reuse the kernel.
//
// NVRTC is loaded lazily with dlopen and the cubin through the runtime's cudaLibrary API, so the library has no link-time
// dependency on libnvrtc / libcuda; if either step fails the caller falls back to the interpreter.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "opm_internal.h"
#include "opm_jit_headers.inc"
#include "opm_kernels.h"

namespace opm {

// NVRTC runs on a worker thread (state COMPILING); the thread that owns the CUDA context loads the finished cubin the next
// time the program is looked up.  Until then, and after a failure, launches use the interpreter kernel.
enum JitState : int { JIT_COMPILING = 0, JIT_CUBIN_READY = 1, JIT_LOADED = 2, JIT_FAILED = 3 };
struct JitKernel {
    std::atomic<int> state{JIT_COMPILING};
    std::thread worker;
    std::vector<char> cubin;
    double compile_ms = 0.0;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    void* d_const = nullptr;  // device address of the kernel's __constant__ op array
    size_t const_bytes = 0;
    int n_ops = 0;
    int occupancy = 1;
    bool has_fork = false;
    int stash_words = 0;  // shared-memory words per thread for the parked rays of the program's FORKs
    int bbox_slots = 0;  // hit maps whose bounding box the kernel keeps (36 bytes of shared memory per thread each)
    int spot_slots = 0;  // bundle copies whose phase-1 spot sums the kernel takes (kSpotSmemBytes per thread each)
    size_t smem = 0;  // dynamic shared memory per CTA
    std::string why;
};

struct JitCache {
    std::map<std::string, std::unique_ptr<JitKernel>> kernels;
    uint64_t compiled = 0, failures = 0, launches = 0;
    double compile_ms = 0.0;
};

namespace {

struct Nvrtc {
    void* h = nullptr;
    bool tried = false;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
};
Nvrtc g_nvrtc;

bool nvrtc_load(std::string* why) {
    Nvrtc& N = g_nvrtc;
    if (N.tried) { if (!N.h && why) *why = "libnvrtc not available"; return N.h != nullptr; }
    N.tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) { N.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (N.h) break; }
    if (!N.h) { if (why) *why = "libnvrtc not available"; return false; }
#define OPM_SYM(field, name) *(void**)(&N.field) = dlsym(N.h, name); if (!N.field) { dlclose(N.h); N.h = nullptr; if (why) *why = "libnvrtc lacks " name; return false; }
    OPM_SYM(CreateProgram, "nvrtcCreateProgram")
    OPM_SYM(CompileProgram, "nvrtcCompileProgram")
    OPM_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    OPM_SYM(GetCUBIN, "nvrtcGetCUBIN")
    OPM_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    OPM_SYM(GetProgramLog, "nvrtcGetProgramLog")
    OPM_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef OPM_SYM
    return true;
}

void appendf(std::string& s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    s += buf;
}

// rotation exactly the identity: the isometry is a pure translation, a structural fact the kernel is specialised for
bool rot_is_identity(const DevIso& I) {
    for (int j = 0; j < 9; ++j)
        if (I.r[j] != ((j % 4 == 0) ? 1.0 : 0.0)) return false;
    return true;
}
bool jit_rot_id_enabled() {
    static const bool on = [] { const char* e = getenv("OPM_B200_JIT_ROT_ID"); return !(e && e[0] == '0'); }();  // tuning switch
    return on;
}

// ops whose index model is evaluated once per CTA for the bundle's common wavelength (RefractCt HINT): refractions into a
// wavelength-dependent medium of a bundle the launcher believes to be monochromatic
bool op_has_hint(const DevOp& op) {
    static const bool on = [] { const char* e = getenv("OPM_B200_JIT_MONO"); return !(e && e[0] == '0'); }();  // tuning switch
    return on && op.kind == OPK_REFRACT && op.u.refract.has_index && op.u.refract.index_type != OPM_IDX_CONST && op.u.refract.hint_wvl > 0.0;
}
bool op_has_bbox(const DevOp& op) { return op.kind == OPK_REFRACT && op.u.refract.hits.e && op.u.refract.hit_acc; }
static int program_bbox_slots(const DevOp* ops, int n_ops) {
    int n = 0;
    for (int k = 0; k < n_ops; ++k) n += op_has_bbox(ops[k]) ? 1 : 0;
    return n;
}
bool op_has_spot(const DevOp& op) { return op.kind == OPK_SNAPSHOT && op.u.snapshot.acc; }
static int program_spot_slots(const DevOp* ops, int n_ops) {
    int n = 0;
    for (int k = 0; k < n_ops; ++k) n += op_has_spot(ops[k]) ? 1 : 0;
    return n;
}
static bool jit_unit_dir_enabled() {
    static const bool on = [] { const char* e = getenv("OPM_B200_JIT_UNIT_DIR"); return !(e && e[0] == '0'); }();  // tuning switch
    return on;
}
// does this op leave every ray that is still valid behind it with a unit direction (up to rounding)?  A refraction / reflection
// (ray.rs:606-631) whose missed rays stop, and the paraxial surface (ray.rs:469), do; isometries and the energy ops keep what they
// find.  (A surface whose missed rays carry on, and a grating, pass such rays through untouched: they establish nothing.)
static void unit_dir_after(const DevOp& op, bool& unit) {
    if ((op.kind == OPK_REFRACT && op.u.refract.strategy == OPM_MISS_STOP) || op.kind == OPK_PARAXIAL) unit = true;
}
std::string refract_view(const DevRefract& o, int k, int hint_slot, bool unit_dir = false) {  // hint_slot < 0: no hint
    const int rot_id = jit_rot_id_enabled() && rot_is_identity(o.inv) && rot_is_identity(o.fwd) ? 1 : 0;
    std::string s;
    appendf(s, "RefractCt<%d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d>(opm_jit_ops[%d].u.refract", o.geom, o.coating, o.has_index ? 1 : 0, o.index_type,
            o.strategy, o.flags, o.refraction_intended ? 1 : 0, o.hits.e ? 1 : 0, rot_id, hint_slot >= 0 ? 1 : 0, unit_dir ? 1 : 0, k);
    if (hint_slot >= 0) appendf(s, ", s_hint + %d", 3 * hint_slot);
    s += ")";
    return s;
}

// one line of kernel body per op; the same text is the structural cache key
std::string op_line(const DevOp& op, int k, int hint_slot = -1, bool unit_dir = false, int park = 63) {
    std::string s;
    switch (op.kind) {
        case OPK_REFRACT: {
            const DevRefract& o = op.u.refract;
            if (op_has_bbox(op)) {  // the kernel keeps the bounding box of the hits it records here (bbox_reg_*, opm_trace_ops.cuh)
                s += "        { HitOut ho; exec_refract(" + refract_view(o, k, hint_slot, unit_dir) + ", x, &ho);";
                appendf(s, " bbox_smem_add(bb_%d, ho); }\n", k);
            } else {
                s += "        exec_refract(" + refract_view(o, k, hint_slot, unit_dir) + ", x);\n";
            }
            break;
        }
        case OPK_APODIZE: {
            const DevAperture& a = op.u.apodize.ap;
            unsigned to = 0;
            for (int j = 0; j < a.n_items && j < OPM_MAX_APERTURE_ITEMS; ++j)
                to |= ((unsigned)(a.items[j].type & 7) | (a.items[j].obstruction ? 8u : 0u)) << (4 * j);
            appendf(s, "        exec_apodize<%s>(ApertureCt<%d, %d, %d, 0x%xu>(opm_jit_ops[%d].u.apodize.ap), opm_jit_ops[%d].u.apodize.inv, x);\n",
                    jit_rot_id_enabled() && rot_is_identity(op.u.apodize.inv) ? "true" : "false", a.n_items, a.is_stack ? 1 : 0,
                    a.stack_obstruction ? 1 : 0, to, k, k);
            break;
        }
        case OPK_THRESHOLD: appendf(s, "        exec_threshold(opm_jit_ops[%d].u.threshold, x);\n", k); break;
        case OPK_LIMITS: appendf(s, "        exec_limits(opm_jit_ops[%d].u.limits, x);\n", k); break;
        case OPK_FILTER: appendf(s, "        exec_filter<%s>(opm_jit_ops[%d].u.filter, x);\n", op.u.filter.s.lam ? "true" : "false", k); break;
        case OPK_PARAXIAL: appendf(s, "        exec_paraxial(opm_jit_ops[%d].u.paraxial, x);\n", k); break;
        case OPK_SPLIT: appendf(s, "        exec_split<%s>(opm_jit_ops[%d].u.split, x);\n", op.u.split.s.lam ? "true" : "false", k); break;
        case OPK_SNAPSHOT:
            appendf(s, "        exec_snapshot_store<%s>(opm_jit_ops[%d].u.snapshot, x);\n", op.u.snapshot.fields == OPM_SNAPSHOT_SPOT ? "true" : "false", k);
            if (op_has_spot(op))
                appendf(s, "        spot_smem_add<%s, %s>(opm_jit_ops[%d].u.snapshot, sp_%d, s_spotn_%d, x);\n", op.u.snapshot.has_frame ? "true" : "false",
                        jit_rot_id_enabled() && rot_is_identity(op.u.snapshot.inv) ? "true" : "false", k, k, k);
            break;
        case OPK_TRANSFORM: appendf(s, "        exec_transform(opm_jit_ops[%d].u.transform, x);\n", k); break;
        case OPK_FORK: appendf(s, "        exec_fork<%s, %d>(opm_jit_ops[%d].u.fork, x, opm_jit_stash);\n", op.u.fork.s.lam ? "true" : "false", park, k); break;
        case OPK_JOIN: appendf(s, "        exec_join<%d>(opm_jit_ops[%d].u.join, x, opm_jit_stash);\n", park, k); break;
        case OPK_DIFFRACT: {
            const DevRefract& o = op.u.diffract.surf;
            appendf(s, "        exec_diffract(RefractCt<%d, %d, %d, %d, %d, %d, %d, %d>(opm_jit_ops[%d].u.diffract.surf), opm_jit_ops[%d].u.diffract, x);\n",
                    o.geom, o.coating, o.has_index ? 1 : 0, o.index_type, o.strategy, o.flags, o.refraction_intended ? 1 : 0, 0, k, k);
            break;
        }
        default: appendf(s, "        /* op %d: unknown kind %d */\n", k, op.kind); break;
    }
    if (op.epi_flags) appendf(s, "        exec_epilogue(EpilogueCt<%d>(), opm_jit_ops[%d], x);\n", op.epi_flags, k);
    return s;
}

// staged tile loads (tile_load_staged, opm_trace_ops.cuh): OPM_B200_JIT_PIPE=1.  Off by default: measured on the C2 chain the
// kernel takes 1.954 ms with them and 1.955 ms without (profiles/r1_summary.md) -- the other resident CTAs already hide the
// load latency of a tile.
static bool jit_pipe_enabled() {
    static const bool on = [] { const char* e = getenv("OPM_B200_JIT_PIPE"); return e && e[0] == '1'; }();
    return on;
}
static size_t jit_stage_bytes(int block) { return ((size_t)block * kRayBytes + 127) & ~(size_t)127; }
static bool program_has_fork(const DevOp* ops, int n_ops) {
    for (int k = 0; k < n_ops; ++k) if (ops[k].kind == OPK_FORK) return true;
    return false;
}
// dynamic shared memory of a specialised kernel: [parked rays of a FORK | stage 0 | stage 1]
// shared-memory words per thread for the parked parent rays of the program's FORKs (0: no FORK)
static size_t jit_stash_bytes(int stash_words, int block) { return stash_words ? (((size_t)block * stash_words * 4 + 127) & ~(size_t)127) : 0; }
// STASH_* bits of opm_trace_ops.cuh (the generator does not include the device header)
enum : int { JS_POS = 1, JS_DIR = 2, JS_OPL = 4, JS_N = 8, JS_BOUNCES = 16, JS_REFR = 32, JS_ALL = 63 };
static int js_words(int park) {
    return 3 + ((park & JS_POS) ? 6 : 0) + ((park & JS_DIR) ? 6 : 0) + ((park & JS_OPL) ? 2 : 0) + ((park & JS_N) ? 2 : 0) + ((park & JS_BOUNCES) ? 1 : 0) +
           ((park & JS_REFR) ? 1 : 0);
}
static bool jit_unit_dir_enabled();
static void unit_dir_after(const DevOp& op, bool& unit);
// unit[k]: every ray that reaches op k has passed an op that leaves a unit direction
static std::vector<char> program_unit_dirs(const DevOp* ops, int n_ops) {
    std::vector<char> unit(n_ops, 0);
    bool u = false;
    for (int k = 0; k < n_ops; ++k) { unit[k] = u && jit_unit_dir_enabled(); unit_dir_after(ops[k], u); }
    return unit;
}
// which fields of the parked parent the side chain of the FORK at op `f` can change (conservative; energy and flags are always
// parked): the bits exec_fork / exec_join are instantiated with
static int fork_park_mask(const DevOp* ops, int n_ops, int f, const std::vector<char>& unit) {
    static const bool on = [] { const char* e = getenv("OPM_B200_JIT_PARK"); return !(e && e[0] == '0'); }();  // tuning switch
    if (!on) return JS_ALL;
    int m = 0;
    for (int k = f + 1; k < n_ops && ops[k].kind != OPK_JOIN; ++k) {
        const DevOp& d = ops[k];
        switch (d.kind) {
            case OPK_REFRACT: {
                const DevRefract& o = d.u.refract;
                m |= JS_POS | JS_OPL;
                if (o.has_index || (o.flags & OPM_REFRACT_CONTINUE_AS_CHILD)) m |= JS_DIR | JS_N | JS_BOUNCES | JS_REFR;
                else if (!(o.geom == OPM_GEOM_PLANE && unit[k])) m |= JS_DIR;  // a passive plane hands a unit direction back bit for bit
                break;
            }
            case OPK_PARAXIAL: m |= JS_POS | JS_DIR | JS_OPL | JS_REFR; break;
            case OPK_TRANSFORM: m |= JS_POS | JS_DIR; break;
            case OPK_APODIZE: case OPK_THRESHOLD: case OPK_LIMITS: case OPK_FILTER: case OPK_SNAPSHOT: case OPK_SPLIT: break;  // energy, flags
            default: return JS_ALL;  // gratings, anything new
        }
    }
    return m;
}
static int program_stash_words(const DevOp* ops, int n_ops) {
    const std::vector<char> unit = program_unit_dirs(ops, n_ops);
    int w = 0;
    for (int k = 0; k < n_ops; ++k)
        if (ops[k].kind == OPK_FORK) w = std::max(w, js_words(fork_park_mask(ops, n_ops, k, unit)));
    return w;
}
static int program_bbox_slots(const DevOp* ops, int n_ops);
// [parked rays of a FORK | tile stages | running bounding boxes of the recorded hit maps | running spot sums of the bundle copies]
static size_t jit_bbox_offset(int stash_words, int block) {
    return jit_stash_bytes(stash_words, block) + (jit_pipe_enabled() && block % 16 == 0 ? 2 * jit_stage_bytes(block) : 0);
}
size_t jit_dynamic_smem(int stash_words, int block, int bbox_slots, int spot_slots) {
    return jit_bbox_offset(stash_words, block) + (size_t)bbox_slots * block * 36 + (size_t)spot_slots * block * kSpotSmemBytes;
}

// Bulk L2 prefetch of the CTA's next tile and its flags loaded one tile ahead (tile_prefetch / TileAhead, opm_trace_ops.cuh): only
// for programs whose tiles keep a warp busy for long compared with a round trip to the HBM -- five or more surfaces.  Measured
// (profiles/r2n_gpu_run.sh, r2o_gpu_run.sh): C2 (13 surfaces, bound by instruction issue) 0.969 -> 0.915 ms per launch; C4
// (3 surfaces, at the HBM roof already) 3.74 -> 5.31 ms with the prefetch -- the warps reach the next tile while its prefetch is
// still on the way and the lines are read twice -- and 3.74 -> 4.22 ms with the flags alone.  OPM_B200_JIT_PREFETCH=0|1 forces both.
static bool program_wants_prefetch(const DevOp* ops, int n_ops) {
    static const int forced = [] { const char* e = getenv("OPM_B200_JIT_PREFETCH"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    if (forced >= 0) return forced != 0;
    int surfaces = 0;
    for (int k = 0; k < n_ops; ++k) surfaces += ops[k].kind == OPK_REFRACT || ops[k].kind == OPK_PARAXIAL || ops[k].kind == OPK_DIFFRACT;
    return surfaces >= 5;
}

std::string make_source(const DevOp* ops, int n_ops, bool compact, int block, int min_ctas) {
    const bool pipe = jit_pipe_enabled() && block % 16 == 0;
    const int stash_w = program_stash_words(ops, n_ops);
    const size_t stash = jit_stash_bytes(stash_w, block);
    // Body of the tile loop.  Rays that are (or become) invalid leave the straight-line code for a second copy of the program --
    // the same op lines under the knowledge `x.valid == false`, which the compiler folds down to the ops' side effects for such a
    // ray (empty hit records, detector copies, FORK / JOIN bookkeeping).  The live copy therefore has no `if (x.valid)` diamonds:
    // no register moves where their two sides would meet again, no convergence barriers.  The two copies meet at the end of the
    // tile, and behind a JOIN, where the parked (possibly valid) parent of a dead split-off ray returns to the live copy.
    static const bool split = [] { const char* e = getenv("OPM_B200_JIT_SPLIT"); return !(e && e[0] == '0'); }();  // tuning switch
    std::string body, hint_init;
    int n_hints = 0;
    std::vector<int> hint_slot(n_ops, -1);
    for (int k = 0; k < n_ops; ++k)
        if (op_has_hint(ops[k])) {
            hint_slot[k] = n_hints++;
            appendf(hint_init, "    if (threadIdx.x == %d) {\n        const DevRefract& o = opm_jit_ops[%d].u.refract;\n        double n2 = 1.0;\n",
                    hint_slot[k] % block, k);
            hint_init += "        const int st = refractive_index(" + refract_view(ops[k].u.refract, k, -1) + ", o.hint_wvl, n2);\n";
            appendf(hint_init, "        s_hint[%d] = st == OPM_OK ? o.hint_wvl : __longlong_as_double(0x7ff8000000000000ll);\n"
                               "        s_hint[%d] = n2; s_hint[%d] = rcp_fast(n2);\n    }\n",
                    3 * hint_slot[k], 3 * hint_slot[k] + 1, 3 * hint_slot[k] + 2);
        }
    const std::vector<char> unit = program_unit_dirs(ops, n_ops);
    std::vector<int> park(n_ops, JS_ALL);  // FORK and its JOIN: the fields the side chain can change
    for (int k = 0, open = -1; k < n_ops; ++k) {
        if (ops[k].kind == OPK_FORK) { open = k; park[k] = fork_park_mask(ops, n_ops, k, unit); }
        else if (ops[k].kind == OPK_JOIN && open >= 0) { park[k] = park[open]; open = -1; }
    }
    if (!split) {
        for (int k = 0; k < n_ops; ++k) body += op_line(ops[k], k, hint_slot[k], unit[k] != 0, park[k]);
    } else {
        std::string dead;
        for (int k = 0; k < n_ops; ++k) {
            const std::string line = op_line(ops[k], k, hint_slot[k], unit[k] != 0, park[k]);
            if (k > 0 && ops[k - 1].kind == OPK_JOIN) appendf(body, "      live_%d:\n", k);
            appendf(body, "        if (!x.valid) goto dead_%d;\n        {\n", k);
            body += line;
            body += "        }\n";
            appendf(dead, "      dead_%d:\n        x.valid = false;\n        {\n", k);
            dead += line;
            dead += "        }\n";
            if (ops[k].kind == OPK_JOIN && k + 1 < n_ops) appendf(dead, "        if (x.valid) goto live_%d;\n", k + 1);
        }
        body += "        goto tile_done;\n";
        body += dead;
        body += "      tile_done:;\n";
    }
    std::string s;
    appendf(s, "#define OPM_JIT_BLOCK %d\n", block);
    if (!program_wants_prefetch(ops, n_ops)) s += "#define OPM_TILE_NO_BULK_PREFETCH 1\n#define OPM_TILE_NO_FLAG_AHEAD 1\n";
    bool key2 = false;  // a second first-valid-ray key (OPM_REFRACT_STAT_KEY2): compiled in only for programs that ask for it
    for (int k = 0; k < n_ops; ++k) key2 |= ops[k].kind == OPK_REFRACT && (ops[k].u.refract.flags & OPM_REFRACT_STAT_KEY2) != 0;
    if (key2) s += "#define OPM_STAT_KEY2 1\n";
    s += "#include \"opm_trace_ops.cuh\"\nnamespace opm {\n";
    appendf(s, "extern \"C\" __constant__ DevOp opm_jit_ops[%d];\n", n_ops);
    appendf(s, "extern \"C\" __global__ void __launch_bounds__(%d, %d)\n", block, min_ctas);
    s += "opm_jit_trace(DevRays in, DevRays out, unsigned long long n, unsigned long long out_cap, unsigned long long* out_tail,\n"
         "              DevStats* stats) {\n"
         "    extern __shared__ __align__(128) unsigned char opm_jit_smem[];  // [parked rays of a FORK | tile stages], sized by the launch\n"
         "    unsigned* opm_jit_stash = reinterpret_cast<unsigned*>(opm_jit_smem);\n"
         "    __shared__ __align__(16) cta_count_t s_cnt[kCntWords];\n"
         "    __shared__ unsigned s_status;\n"
         "    __shared__ unsigned long long s_first_inv;\n"
         "#ifdef OPM_STAT_KEY2\n"
         "    __shared__ unsigned long long s_first_inv2;\n"
         "    if (threadIdx.x == 0) s_first_inv2 = 0ull;\n"
         "#define OPM_JIT_KEY2_PTR (&s_first_inv2)\n"
         "#define OPM_JIT_KEY2_VAL s_first_inv2\n"
         "#else\n"
         "#define OPM_JIT_KEY2_PTR ((unsigned long long*)0)\n"
         "#define OPM_JIT_KEY2_VAL 0ull\n"
         "#endif\n"
         "    cta_count_init(s_cnt);\n"
         "    if (threadIdx.x == 0) { s_status = 0u; s_first_inv = 0ull; }\n";
    if (n_hints) {
        appendf(s, "    __shared__ double s_hint[%d];  // per hinted op: wavelength | n2 | 1 / n2\n", 3 * n_hints);
        s += hint_init;
    }
    {
        size_t sp0 = jit_bbox_offset(stash_w, block) + (size_t)program_bbox_slots(ops, n_ops) * block * 36;
        sp0 = (sp0 + 7) & ~(size_t)7;
        for (int k = 0; k < n_ops; ++k)
            if (op_has_spot(ops[k])) {
                appendf(s, "    unsigned char* const sp_%d = opm_jit_smem + %zu;\n    __shared__ unsigned s_spotn_%d[2];\n    spot_smem_init(sp_%d, s_spotn_%d);\n",
                        k, sp0, k, k, k);
                sp0 += (size_t)block * kSpotSmemBytes;
            }
    }
    s += "    __syncthreads();\n"
         "    const unsigned tiles = (unsigned)((n + OPM_BDIM - 1) / OPM_BDIM);\n";
    if (pipe) {
        s += "    __shared__ unsigned long long s_bar[2];\n"
             "    TilePipe pipe;\n";
        appendf(s, "    tile_pipe_init(pipe, opm_jit_smem + %zu, s_bar, in, n);\n", stash);
    }
    {
        const size_t bb0 = jit_bbox_offset(stash_w, block);
        int j = 0;
        for (int k = 0; k < n_ops; ++k)
            if (op_has_bbox(ops[k])) {
                appendf(s, "    unsigned char* const bb_%d = opm_jit_smem + %zu;\n    bbox_smem_init(bb_%d);\n", k, bb0 + (size_t)j * block * 36, k);
                ++j;
            }
    }
    s += "    unsigned warp_best = 0xffffffffu;\n"
         "    TileAhead ahead;\n"
         "    tile_ahead_init(ahead, in, (uint64_t)blockIdx.x * OPM_BDIM + threadIdx.x, n);\n"
         "    for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {\n"
         "        RayCtx x;\n"
         "        uint8_t vb;\n";
    s += pipe ? "        tile_load_staged(pipe, in, tile, n, x, vb);\n"
              : "        tile_load(in, (uint64_t)tile * OPM_BDIM + threadIdx.x, n, x, vb, ahead, (uint64_t)(tile + gridDim.x) * OPM_BDIM + threadIdx.x);\n"
                "        tile_prefetch(in, (uint64_t)(tile + OPM_TILE_PREFETCH_DIST * gridDim.x) * OPM_BDIM + threadIdx.x, n);\n";
    s += body;
    appendf(s, "        tile_store<%s>(in, out, out_cap, out_tail, x, vb);\n", compact ? "true" : "false");
    s += "        tile_count(x, s_cnt, &s_status, &s_first_inv, warp_best, OPM_JIT_KEY2_PTR);\n"
         "    }\n"
         "    __syncthreads();\n";
    for (int k = 0; k < n_ops; ++k)
        if (op_has_bbox(ops[k])) appendf(s, "    bbox_reg_flush(bbox_smem_get(bb_%d), opm_jit_ops[%d].u.refract.hit_acc);\n", k, k);
    for (int k = 0; k < n_ops; ++k)
        if (op_has_spot(ops[k])) appendf(s, "    spot_smem_flush(sp_%d, s_spotn_%d, opm_jit_ops[%d].u.snapshot.acc);\n", k, k, k);
    s += "    cta_flush(stats, s_cnt, s_status, s_first_inv, OPM_JIT_KEY2_VAL);\n"
         "}\n}  // namespace opm\n";
    return s;
}

// source -> sm_100a cubin (needs no GPU)
bool compile_cubin(const std::string& src, std::vector<char>& cubin, std::string& why) {
    Nvrtc& N = g_nvrtc;
    nvrtcProgram prog = nullptr;
    if (N.CreateProgram(&prog, src.c_str(), "opm_jit_trace.cu", kJitHeaderCount, kJitHeaderSources, kJitHeaderNames) != NVRTC_SUCCESS) {
        why = "nvrtcCreateProgram failed";
        return false;
    }
    // -default-device: the public header declares the (host) entry points without execution-space annotations
    std::vector<std::string> extra;  // tuning: OPM_B200_JIT_DEFINES="-DOPM_STORE_CG -D..." is passed through to the compiler
    if (const char* e = getenv("OPM_B200_JIT_DEFINES")) {
        std::string all(e);
        size_t pos = 0;
        while (pos < all.size()) {
            size_t sp = all.find(' ', pos);
            if (sp == std::string::npos) sp = all.size();
            if (sp > pos) extra.push_back(all.substr(pos, sp - pos));
            pos = sp + 1;
        }
    }
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    for (auto& x : extra) opts.push_back(x.c_str());
    nvrtcResult rc = N.CompileProgram(prog, (int)opts.size(), opts.data());
    if (rc != NVRTC_SUCCESS) {
        size_t ls = 0;
        N.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) N.GetProgramLog(prog, &log[0]);
        why = "nvrtc compilation failed: " + log.substr(0, 3000);
        N.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    N.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    N.GetCUBIN(prog, cubin.data());
    N.DestroyProgram(&prog);
    return true;
}

// cubin -> kernel handle + address of its constant op array (on the thread that owns the CUDA context)
bool load(JitKernel& K, int n_ops, int block) {
    cudaError_t e = cudaLibraryLoadData(&K.lib, K.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { K.why = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); cudaGetLastError(); return false; }
    e = cudaLibraryGetKernel(&K.kernel, K.lib, "opm_jit_trace");
    if (e == cudaSuccess) e = cudaLibraryGetGlobal(&K.d_const, &K.const_bytes, K.lib, "opm_jit_ops");
    if (e != cudaSuccess) { K.why = std::string("cudaLibraryGetKernel/Global: ") + cudaGetErrorString(e); cudaGetLastError(); return false; }
    if (K.const_bytes < (size_t)n_ops * sizeof(DevOp)) { K.why = "constant array smaller than the program"; return false; }
    int nb = 0;
    K.smem = jit_dynamic_smem(K.stash_words, block, K.bbox_slots, K.spot_slots) + 8;
    if (const char* e = getenv("OPM_B200_JIT_SMEM_PAD_KB")) K.smem += (size_t)atoi(e) * 1024;  // probe: what the L1 share of the SM's memory is worth
    if (K.smem > 48 * 1024 &&
        cudaFuncSetAttribute((const void*)K.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K.smem) != cudaSuccess) {
        K.why = "dynamic shared memory of the specialised kernel not available"; cudaGetLastError(); return false;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)K.kernel, block, K.smem) != cudaSuccess || nb < 1) { nb = 1; cudaGetLastError(); }
    K.occupancy = nb;
    K.n_ops = n_ops;
    K.cubin.clear();
    K.cubin.shrink_to_fit();
    return true;
}

}  // namespace

// A process may end while a compilation is still in flight (a context that was never destroyed, an exception on the
// caller's side): the exit handler below waits for the workers, so that NVRTC is not torn down under a running compile.
static std::mutex g_inflight_mu;
static std::condition_variable g_inflight_cv;
static int g_inflight = 0;
static void wait_for_workers_at_exit() {
    std::unique_lock<std::mutex> lk(g_inflight_mu);
    g_inflight_cv.wait(lk, [] { return g_inflight == 0; });
}
static void worker_begin() {
    static std::once_flag once;  // registered after NVRTC was loaded, so it runs before NVRTC's own exit handlers
    std::call_once(once, [] { atexit(wait_for_workers_at_exit); });
    std::lock_guard<std::mutex> lk(g_inflight_mu);
    ++g_inflight;
}
static void worker_end() {
    std::lock_guard<std::mutex> lk(g_inflight_mu);
    if (--g_inflight == 0) g_inflight_cv.notify_all();
}

JitCache* jit_cache_create() { return new JitCache(); }
void jit_cache_destroy(JitCache* c) {
    if (!c) return;
    for (auto& kv : c->kernels) {
        if (kv.second->worker.joinable()) kv.second->worker.join();
        if (kv.second->lib) cudaLibraryUnload(kv.second->lib);
    }
    delete c;
}
void jit_cache_stats(const JitCache* c, uint64_t* compiled, uint64_t* failures, uint64_t* launches, double* compile_ms) {
    if (compiled) *compiled = c ? c->compiled : 0;
    if (failures) *failures = c ? c->failures : 0;
    if (launches) *launches = c ? c->launches : 0;
    if (compile_ms) *compile_ms = c ? c->compile_ms : 0.0;
}
// blocks until no compilation is in flight (benchmarks and tests: measure / check the specialised kernels, not the hand-over)
void jit_cache_wait(JitCache* c) {
    if (!c) return;
    for (auto& kv : c->kernels) if (kv.second->worker.joinable()) kv.second->worker.join();
}

// wait = false: returns nullptr while the kernel of this program is still being compiled (the caller interprets meanwhile)
const JitKernel* jit_get(JitCache* cache, const DevOp* ops, int n_ops, bool compact, int block, int min_ctas, bool wait,
                         const char** why) {
    static thread_local std::string s_why;
    if (why) *why = nullptr;
    if (!cache || n_ops <= 0) { if (why) *why = "no program"; return nullptr; }
    std::string src = make_source(ops, n_ops, compact, block, min_ctas);
    if (const char* dump = getenv("OPM_B200_JIT_DUMP")) {  // diagnosis: the generated source, to compile offline with nvcc -Xptxas -v
        // one file per distinct program: <dump>_<ops>_<hash>.cu
        char name[1024];
        snprintf(name, sizeof name, "%s_%d_%08zx.cu", dump, n_ops, std::hash<std::string>()(src) & 0xffffffffu);
        if (FILE* f = fopen(name, "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
    }
    auto it = cache->kernels.find(src);
    if (it == cache->kernels.end()) {
        it = cache->kernels.emplace(std::move(src), std::unique_ptr<JitKernel>(new JitKernel())).first;
        JitKernel* k = it->second.get();
        k->has_fork = program_has_fork(ops, n_ops);
        k->stash_words = program_stash_words(ops, n_ops);
        k->bbox_slots = program_bbox_slots(ops, n_ops);
        k->spot_slots = program_spot_slots(ops, n_ops);
        if (!nvrtc_load(&k->why)) {
            k->state = JIT_FAILED;
            cache->failures++;
        } else {
            const std::string* text = &it->first;  // map nodes are stable: the worker reads the key in place
            worker_begin();
            k->worker = std::thread([k, text]() {
                const auto t0 = std::chrono::steady_clock::now();
                const bool ok = compile_cubin(*text, k->cubin, k->why);
                k->compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                k->state.store(ok ? JIT_CUBIN_READY : JIT_FAILED, std::memory_order_release);
                worker_end();
            });
        }
    }
    JitKernel& K = *it->second;
    int st = K.state.load(std::memory_order_acquire);
    if (st == JIT_COMPILING && wait) {
        if (K.worker.joinable()) K.worker.join();
        st = K.state.load(std::memory_order_acquire);
    }
    if (st == JIT_CUBIN_READY) {
        if (K.worker.joinable()) K.worker.join();
        cache->compile_ms += K.compile_ms;
        if (load(K, n_ops, block)) { K.state = JIT_LOADED; cache->compiled++; }
        else { K.state = JIT_FAILED; cache->failures++; }
        st = K.state.load();
    } else if (st == JIT_FAILED && K.worker.joinable()) {
        K.worker.join();
        cache->failures++;
        cache->compile_ms += K.compile_ms;
    }
    if (st == JIT_LOADED) return &K;
    if (st == JIT_FAILED) { s_why = K.why; if (why) *why = s_why.c_str(); }
    return nullptr;
}

int jit_occupancy(const JitKernel* K) { return K ? K->occupancy : 1; }
void* jit_const_ptr(const JitKernel* K) { return K ? (void*)K->d_const : nullptr; }

// Build check without a GPU: a program with one op of every kind, every geometry and coating is turned into source and
// compiled to an sm_100a cubin.  Returns the cubin size, or 0 with the reason in `why`.
static size_t jit_selftest_impl(std::string& why) {
    if (!nvrtc_load(&why)) return 0;
    std::vector<DevOp> ops;
    auto push = [&](int kind) { DevOp d; memset(&d, 0, sizeof d); d.kind = kind; ops.push_back(d); return &ops.back(); };
    for (int geom = OPM_GEOM_PLANE; geom <= OPM_GEOM_PARABOLA; ++geom) {
        DevOp* d = push(OPK_REFRACT);
        d->u.refract.geom = geom; d->u.refract.coating = geom % 3; d->u.refract.has_index = 1; d->u.refract.index_type = geom;
        d->u.refract.refraction_intended = 1; d->u.refract.hits.e = (double*)8; d->epi_flags = EPI_THRESHOLD | EPI_LIMITS | EPI_CHECK_REFR;
        if (geom == OPM_GEOM_PLANE) d->u.refract.flags = OPM_REFRACT_STAT_KEY;       // the two first-valid-ray keys of a fused lens launch
        if (geom == OPM_GEOM_SPHERE) d->u.refract.flags = OPM_REFRACT_STAT_KEY2;
    }
    {
        DevOp* d = push(OPK_REFRACT);  // mirror that also emits its children
        d->u.refract.flags = OPM_REFRACT_EMIT_CHILDREN | OPM_REFRACT_CONTINUE_AS_CHILD; d->u.refract.coating = OPM_COAT_CONSTANT_R;
    }
    {
        DevOp* d = push(OPK_APODIZE);
        d->u.apodize.ap.n_items = 4; d->u.apodize.ap.is_stack = 1;
        for (int j = 0; j < 4; ++j) d->u.apodize.ap.items[j].type = OPM_AP_CIRCLE + j;
    }
    for (int kind : {OPK_THRESHOLD, OPK_LIMITS, OPK_FILTER, OPK_PARAXIAL, OPK_SPLIT, OPK_SNAPSHOT, OPK_TRANSFORM, OPK_DIFFRACT, OPK_FORK, OPK_JOIN}) push(kind);
    {
        DevOp* d = push(OPK_SNAPSHOT);  // spot-diagram copy with the phase-1 sums taken on the way
        d->u.snapshot.fields = OPM_SNAPSHOT_SPOT; d->u.snapshot.has_frame = 1; d->u.snapshot.acc = (void*)8;
    }
    std::vector<char> cubin;
    for (bool compact : {false, true}) {
        std::string src = make_source(ops.data(), (int)ops.size(), compact, 256, 2);
        if (!compile_cubin(src, cubin, why)) return 0;
    }
    return cubin.size();
}
size_t jit_selftest(char* why, size_t cap) {
    std::string w;
    size_t n = jit_selftest_impl(w);
    if (why && cap) { snprintf(why, cap, "%s", w.c_str()); }
    return n;
}

// h_ops: the pinned, mapped copy of the program (the descriptors' numbers and pointers of THIS launch)
int jit_launch(JitCache* cache, const JitKernel* K, const TraceLaunch& L, const void* h_ops) {
    cudaStream_t st = (cudaStream_t)L.stream;
    if (h_ops) {  // NULL: the caller's prologue launch has already brought the descriptors to jit_const_ptr()
        int rc = k_fetch_from_host(K->d_const, h_ops, (size_t)L.n_ops * sizeof(DevOp), L.stream);
        if (rc) return rc;
    }
    DevRays in = L.in, out = L.out;
    unsigned long long n = L.n, cap = L.out_cap;
    unsigned long long* tail = L.out_tail;
    DevStats* stats = L.stats;
    void* args[] = {&in, &out, &n, &cap, &tail, &stats};
    cudaError_t e = cudaLaunchKernel((const void*)K->kernel, dim3(L.grid), dim3(L.block), args, K->smem, st);
    if (e != cudaSuccess) return (int)e;
    cache->launches++;
    return (int)cudaGetLastError();
}

}  // namespace opm
