// opm_trace_ops.cuh -- the per-op bodies of the fused surface-sequence kernel.
//
// One function per `Rays` method of the reference (file:line cited), operating on a ray held in registers (RayCtx).
// Shared by the two kernels that run an op program:
//   * the interpreter (opm_trace.cu): ops staged in shared memory, dispatched by a warp-uniform switch, structural facts
//     read at run time (RefractRt / ApertureRt / EpilogueRt views);
//   * the run-time-compiled kernel of a fixed program (opm_jit.cpp): the same functions called back to back with the
//     structural facts as template arguments (*Ct views), descriptors read from constant memory.
#pragma once
#include "opm_device.cuh"

namespace opm {

// CTA size: a compile-time constant in the run-time specialised kernel (the generator defines OPM_JIT_BLOCK: the parked rays of
// a FORK and the running bounding boxes are then addressed as base + immediate), blockDim.x in the interpreter
#ifdef OPM_JIT_BLOCK
#define OPM_BDIM ((unsigned)(OPM_JIT_BLOCK))
#else
#define OPM_BDIM (blockDim.x)
#endif

// fire-and-forget add of ONE lane to a shared-memory counter.  Written as PTX: for an atomicAdd in divergent code the compiler
// emits the warp-aggregated form (vote, leader election, popc, multiply: 12 instructions) although the caller has already
// reduced over the warp and elected lane 0
__device__ __forceinline__ void red_shared_add(unsigned* p, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// warp-aggregated reservation of `want ? 1 : 0` slots per lane; returns this lane's slot.  Aggregates over the lanes that arrive
// here together (__activemask): the live and the dead copy of a specialised program call this from different places, each group
// of lanes then reserves for itself
__device__ __forceinline__ unsigned long long warp_reserve(bool want, unsigned long long* tail) {
    const unsigned active = __activemask();
    unsigned mask = __ballot_sync(active, want);
    if (mask == 0) return 0;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(tail, (unsigned long long)__popc(mask));
    base = __shfl_sync(active, base, leader);
    return base + __popc(mask & lanemask_lt());
}

// cache operator of the bundle / hit-record stores (tuning: -DOPM_STORE_PLAIN / -DOPM_STORE_CG through OPM_B200_JIT_DEFINES)
#if defined(OPM_STORE_PLAIN)
#define OPM_ST(ptr, v) (*(ptr) = (v))
#elif defined(OPM_STORE_CG)
#define OPM_ST(ptr, v) __stcg((ptr), (v))
#else
#define OPM_ST(ptr, v) __stcs((ptr), (v))
#endif

__device__ __forceinline__ void load_ray(const DevRays& R, uint64_t i, RayState& r) {
    r.pos = v3(__ldcs(R.px + i), __ldcs(R.py + i), __ldcs(R.pz + i));
    r.dir = v3(__ldcs(R.dx + i), __ldcs(R.dy + i), __ldcs(R.dz + i));
    r.e = __ldcs(R.e + i); r.wvl = __ldcs(R.wvl + i); r.opl = __ldcs(R.opl + i); r.n = __ldcs(R.n + i);
    r.bounces = __ldcs(R.bounces + i); r.refr = __ldcs(R.refr + i); r.id = __ldcs(R.id + i);
}
template <bool FULL>
__device__ __forceinline__ void store_ray(const DevRays& R, uint64_t i, const RayState& r, uint8_t valid) {
    OPM_ST(R.px + i, r.pos.x); OPM_ST(R.py + i, r.pos.y); OPM_ST(R.pz + i, r.pos.z);
    OPM_ST(R.dx + i, r.dir.x); OPM_ST(R.dy + i, r.dir.y); OPM_ST(R.dz + i, r.dir.z);
    OPM_ST(R.e + i, r.e); OPM_ST(R.opl + i, r.opl); OPM_ST(R.n + i, r.n);
#ifndef OPM_DIAG_NO_U32  // diagnosis builds only (profiles/store_variants.sh): which stores cause the ECC read traffic?
    OPM_ST(R.bounces + i, r.bounces); OPM_ST(R.refr + i, r.refr);
#endif
    if (FULL) {  // wavelength and id never change in place
        OPM_ST(R.wvl + i, r.wvl);
#ifndef OPM_DIAG_NO_U32
        OPM_ST(R.id + i, r.id);
#endif
    }
#ifndef OPM_DIAG_NO_U8
    R.valid[i] = valid;
#endif
}

// Per-tile event counters, two 16-bit fields per register (a warp's sum of a field is at most 32 lanes x 64 ops):
// a = interactions | hits << 16, b = missed | children << 16.  Each tile ends with one REDUX per register and one
// shared-memory atomic per field from lane 0; the CTA flushes its totals to the global statistics once.
struct Counters {
    unsigned a, b, apodized, status;
    // OPM_REFRACT_STAT_KEY: (id, bounces) of this ray right after the flagged refract op when it was valid there (kid = 0xffffffff:
    // not valid); kid = kKeyUnset: no flagged op ran, the launch statistic `first_valid_key` looks at the end of the program
    unsigned kid, kb;
    // OPM_REFRACT_STAT_KEY2: the same pair right after a second flagged op (kid2 = kKeyUnset: no such op ran).  Compiled in where
    // OPM_STAT_KEY2 is defined: the interpreter, and the specialised kernels of programs that hold such an op -- the others carry
    // nothing of it
#ifdef OPM_STAT_KEY2
    unsigned kid2, kb2;
#endif
};
#ifdef OPM_STAT_KEY2
#define OPM_KEY2_RESET(c) do { (c).kid2 = kKeyUnset; (c).kb2 = 0; } while (0)
#else
#define OPM_KEY2_RESET(c) do { } while (0)
#endif
constexpr unsigned kKeyUnset = 0xfffffffeu;
enum { CNT_INTERACTIONS = 0, CNT_CHILDREN, CNT_HITS, CNT_MISSED, CNT_APODIZED, CNT_VALID_OUT, CNT_ALIVE_OUT, CNT_N };

// one ray of the current tile
struct RayCtx {
    RayState r;
    uint64_t i;      // slot in the bundle
    bool in_range;   // i < n
    bool alive;      // part of the bundle (not removed)
    bool valid;      // Ray::valid
    // some slot of this warp's 32 held a ray when the tile was loaded: every slot-indexed store of the tile then writes ALL its
    // in-range lanes (whatever the registers hold for a removed ray, flagged RAY_REMOVED / e < 0), so that every 32-byte sector
    // reaches the L2 complete.  Sectors with holes are read back from DRAM before they can be written: on C2, where 8.6 % of the
    // rays leave the bundle at scattered slots, that was 0.64 GB of reads per launch (ncu, profiles/r2_summary.md)
    bool warp_live;
    Counters c;
};

// folded threshold + limits (rays.rs:1145-1162, 1386-1404): EPI_* flags
struct EpilogueRt {
    int f;
    __device__ __forceinline__ explicit EpilogueRt(int f_) : f(f_) {}
    __device__ __forceinline__ int flags() const { return f; }
};
template <int FLAGS>
struct EpilogueCt {
    __device__ __forceinline__ static constexpr int flags() { return FLAGS; }
};
template <class E>
__device__ __forceinline__ void exec_epilogue(const E& e, const DevOp& op, RayCtx& x) {
    const int f = e.flags();
    if (f == 0 || !x.alive) return;
    if ((f & EPI_THRESHOLD) && x.r.e < op.epi_min_energy) x.valid = false;
    if (f & EPI_LIMITS) {
        if (x.r.bounces > op.epi_max_bounces) x.valid = false;
        if ((f & EPI_CHECK_REFR) && x.r.refr >= op.epi_max_refr) x.valid = false;
    }
}

// what happens to the reflected clone of a refract / diffract op: appended to the children bundle (ghost focus) and / or
// taking the parent's place in the bundle (mirrors, gratings: rays without a clone leave the bundle)
template <class V>
__device__ __forceinline__ void finish_children(const V& v, const RefractResult& res, RayCtx& x) {
    const DevRefract& o = v.d();
    if ((v.flags() & OPM_REFRACT_EMIT_CHILDREN) && (v.flags() & kRefractChildrenBySlot)) {
        if (x.in_range) {
            if (res.has_child || x.warp_live) {  // whole sectors: a slot without a child is a removed ray
                RayState child = x.r;
                if (res.has_child) { child_into(res, child); x.c.b += 1u << 16; }
                store_ray<true>(o.children, x.i, child, res.has_child ? RAY_VALID : RAY_REMOVED);
            } else o.children.valid[x.i] = RAY_REMOVED;
        }
    } else if (v.flags() & OPM_REFRACT_EMIT_CHILDREN) {
        unsigned long long slot = warp_reserve(res.has_child, o.child_tail);
        if (res.has_child) {
            if (slot < o.child_cap) {
                RayState child = x.r;
                child_into(res, child);
                store_ray<true>(o.children, slot, child, RAY_VALID);
                x.c.b += 1u << 16;
            } else x.c.status |= 1u << OPM_ERR_CAPACITY;
        }
    }
    if (v.flags() & OPM_REFRACT_CONTINUE_AS_CHILD) {  // mirror: only reflected rays stay in the bundle
        if (x.alive) {
            if (res.has_child) child_into(res, x.r);
            else { x.alive = false; x.valid = false; }
        }
    }
}

// Rays::refract_on_surface (rays.rs:976-1047) for one ray + hit record + child emission / mirror continuation.  Returns
// true when the ray reached the surface, i.e. when the reference pushes the old position to `pos_hist` (ray.rs:616).
// the hit exec_refract recorded for this ray, for the bounding box the specialised kernel keeps next to the hit map
struct HitOut { double x, y; bool ok; };
template <class V>
__device__ __forceinline__ bool exec_refract(const V& v, RayCtx& x, HitOut* ho = nullptr) {
    const DevRefract& o = v.d();
    RefractResult res;
    res.has_child = false; res.moved = false; res.hit_e = -1.0; res.hit_bounce = 0; res.hit_local = v3(0, 0, 0);
    if (x.valid) {
        x.c.a += 1u;
        res = refract_on_surface(v, x.r);
        if (res.status != OPM_OK) { x.c.status |= 1u << res.status; x.valid = false; res.has_child = false; res.hit_e = -1.0; }
        else {
            if (res.missed) x.c.b += 1u;
            if (res.invalidated) x.valid = false;
        }
    }
    if (ho) ho->ok = false;
    if (v.has_hits() && x.in_range) {  // slot-indexed hit record, e < 0 = empty
        OPM_ST(o.hits.e + x.i, res.hit_e);
        if (res.hit_e >= 0.0 || x.warp_live) {  // full sectors (RayCtx::warp_live): the fields of a slot with e < 0 mean nothing
            OPM_ST(o.hits.x + x.i, res.hit_local.x); OPM_ST(o.hits.y + x.i, res.hit_local.y);
            OPM_ST(o.hits.z + x.i, res.hit_local.z); OPM_ST(o.hits.bounce + x.i, res.hit_bounce);
        }
        if (res.hit_e >= 0.0) {
            x.c.a += 1u << 16;
            if (ho) { ho->ok = true; ho->x = res.hit_local.x; ho->y = res.hit_local.y; }
        }
    }
    // Rays::bounce_lvl as the ghost-focus analysis reads it: between refract_on_surface and apodize (analyzers/raytrace.rs:118-133)
    if (v.flags() & OPM_REFRACT_STAT_KEY) { x.c.kid = (x.valid && x.alive) ? x.r.id : 0xffffffffu; x.c.kb = x.r.bounces; }
#ifdef OPM_STAT_KEY2
    if (v.flags() & OPM_REFRACT_STAT_KEY2) { x.c.kid2 = (x.valid && x.alive) ? x.r.id : 0xffffffffu; x.c.kb2 = x.r.bounces; }
#endif
    finish_children(v, res, x);
    return res.moved;
}

// Bounding box of the hits a launch records (rays_hit_map.rs:522-562), kept by the specialised kernel while it writes the hit
// map, so that the detector read-out needs no pass of its own over the map: per-thread running extrema across the tiles, one
// reduction per CTA at the end of the kernel, ordered-key atomics into the hit map's accumulator (layout of BBoxAccum).
// The running extrema live in dynamic shared memory, field-major per thread ([x0 | x1 | y0 | y1][blockDim] doubles, then
// [n][blockDim] counts: 36 bytes per thread and hit map) -- registers held across the tile loop cost the kernel more than the
// separate pass they replace.
struct BBoxReg { double x0, x1, y0, y1; unsigned n; };
__device__ __forceinline__ size_t bbox_slot_bytes() { return (size_t)OPM_BDIM * 36; }
__device__ __forceinline__ void bbox_smem_init(unsigned char* slot) {
    double* d = reinterpret_cast<double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
    d[t] = 1.7976931348623157e308; d[B + t] = -1.7976931348623157e308; d[2 * B + t] = 1.7976931348623157e308; d[3 * B + t] = -1.7976931348623157e308;
    reinterpret_cast<unsigned*>(d + 4 * B)[t] = 0u;
}
__device__ __forceinline__ void bbox_smem_add(unsigned char* slot, const HitOut& h) {
    if (!h.ok) return;
    double* d = reinterpret_cast<double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
    // compare, store only on change: after a thread's first tiles its extrema rarely move
    if (h.x < d[t]) d[t] = h.x;
    if (h.x > d[B + t]) d[B + t] = h.x;
    if (h.y < d[2 * B + t]) d[2 * B + t] = h.y;
    if (h.y > d[3 * B + t]) d[3 * B + t] = h.y;
    reinterpret_cast<unsigned*>(d + 4 * B)[t] += 1u;
}
__device__ __forceinline__ BBoxReg bbox_smem_get(const unsigned char* slot) {
    const double* d = reinterpret_cast<const double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
    BBoxReg b;
    b.x0 = d[t]; b.x1 = d[B + t]; b.y0 = d[2 * B + t]; b.y1 = d[3 * B + t]; b.n = reinterpret_cast<const unsigned*>(d + 4 * B)[t];
    return b;
}
__device__ __forceinline__ unsigned long long ordered_key(double d) {  // monotonic double -> uint64 (opm_kernels.cu dkey)
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
// all threads of the CTA; ends with a barrier, so calls may follow one another
__device__ __forceinline__ void bbox_reg_flush(BBoxReg b, unsigned long long* acc) {
    __shared__ double sh[32][4];
    __shared__ unsigned shn[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b.x0 = fmin(b.x0, __shfl_xor_sync(0xffffffffu, b.x0, o)); b.x1 = fmax(b.x1, __shfl_xor_sync(0xffffffffu, b.x1, o));
        b.y0 = fmin(b.y0, __shfl_xor_sync(0xffffffffu, b.y0, o)); b.y1 = fmax(b.y1, __shfl_xor_sync(0xffffffffu, b.y1, o));
        b.n += __shfl_xor_sync(0xffffffffu, b.n, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[w][0] = b.x0; sh[w][1] = b.x1; sh[w][2] = b.y0; sh[w][3] = b.y1; shn[w] = b.n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long n = 0;
        for (int k = 0; k < (int)((OPM_BDIM + 31) >> 5); ++k) {
            if (!shn[k]) continue;
            n += shn[k];
            b.x0 = fmin(b.x0, sh[k][0]); b.x1 = fmax(b.x1, sh[k][1]); b.y0 = fmin(b.y0, sh[k][2]); b.y1 = fmax(b.y1, sh[k][3]);
        }
        if (n) {
            atomicMin(&acc[0], ordered_key(b.x0)); atomicMax(&acc[1], ordered_key(b.x1));
            atomicMax(&acc[2], ordered_key(b.y1)); atomicMin(&acc[3], ordered_key(b.y0));
            atomicAdd(&acc[4], n);
        }
    }
    __syncthreads();
}

// Rays::diffract_on_periodic_surface (rays.rs:1058-1111) for one ray: reflective grating; returns like
// exec_refract (ray.rs:537)
template <class V>
__device__ __forceinline__ bool exec_diffract(const V& v, const DevDiffract& o, RayCtx& x) {
    RefractResult res;
    res.has_child = false; res.moved = false;
    if (x.valid) {
        x.c.a += 1u;
        res = diffract_on_periodic_surface(v, o, x.r);
        if (res.status != OPM_OK) { x.c.status |= 1u << res.status; x.valid = false; res.has_child = false; }
        else {
            if (res.missed) x.c.b += 1u;
            if (res.invalidated) x.valid = false;
        }
    }
    finish_children(v, res, x);
    return res.moved;
}

// Rays::apodize (rays.rs:557-573); returns true for a ray this call invalidated (its position goes to `pos_hist`, :566)
// ROT_ID: `inv` is a pure translation (see RefractCt)
template <bool ROT_ID = false, class A>
__device__ __forceinline__ bool exec_apodize(const A& av, const DevIso& inv, RayCtx& x) {
    if (!x.valid) return false;
    double lx = ROT_ID ? x.r.pos.x + inv.t[0] : inv.r[0] * x.r.pos.x + inv.r[1] * x.r.pos.y + inv.r[2] * x.r.pos.z + inv.t[0];
    double ly = ROT_ID ? x.r.pos.y + inv.t[1] : inv.r[3] * x.r.pos.x + inv.r[4] * x.r.pos.y + inv.r[5] * x.r.pos.z + inv.t[1];
    double f = aperture_factor(av, lx, ly);
    if (f > 0.0) {
        if (!(f <= 1.0)) x.c.status |= 1u << OPM_ERR_APODIZE_FACTOR;
        else x.r.e *= f;
    } else { x.valid = false; x.c.apodized += 1u; return true; }
    return false;
}

// rays.rs:1145-1162 (all rays; a no-op for already invalid ones)
__device__ __forceinline__ void exec_threshold(const DevThreshold& o, RayCtx& x) {
    if (x.alive && x.r.e < o.min_energy) x.valid = false;
}
// rays.rs:1386-1404
__device__ __forceinline__ void exec_limits(const DevLimits& o, RayCtx& x) {
    if (!x.alive) return;
    if (x.r.bounces > o.max_bounces) x.valid = false;
    if (o.check_refr && x.r.refr >= o.max_refr) x.valid = false;
}
// the per-ray splitting ratio / transmission of a spectrum-driven op (ray.rs:712-721, 755-761); false = the wavelength is
// outside the spectrum (the reference returns an error: the ray raises OPM_ERR_SPECTRUM)
// SPECTRAL = false: the program is known to use the fixed factor (run-time specialised kernel), the look-up is compiled out
template <bool SPECTRAL = true>
__device__ __forceinline__ bool spectral_factor(const DevSpectrum& s, double fixed, RayCtx& x, double& f) {
    f = fixed;
    if (!SPECTRAL || s.lam == nullptr) return true;
    if (spectrum_get_value(s, x.r.wvl, f)) return true;
    x.c.status |= 1u << OPM_ERR_SPECTRUM;
    x.valid = false;
    return false;
}
// rays.rs:1119-1133
template <bool SPECTRAL = true>
__device__ __forceinline__ void exec_filter(const DevFilter& o, RayCtx& x) {
    if (!x.valid) return;
    double t;
    if (spectral_factor<SPECTRAL>(o.s, o.t, x, t)) x.r.e *= t;
}
// rays.rs:943-959
__device__ __forceinline__ void exec_paraxial(const DevParaxial& o, RayCtx& x) {
    if (x.valid) refract_paraxial(o, x.r);
}
// rays.rs:1253-1264, ray.rs:752-772
template <bool SPECTRAL = true>
__device__ __forceinline__ void exec_split(const DevSplit& o, RayCtx& x) {
    if (!x.in_range) return;
    if (x.alive) {
        RayState s = x.r;
        const bool was_valid = x.valid;
        if (x.valid) {
            double ratio;
            if (spectral_factor<SPECTRAL>(o.s, o.ratio, x, ratio)) {
                if (SPECTRAL && o.s.lam != nullptr && !(ratio >= 0.0 && ratio <= 1.0)) { x.c.status |= 1u << OPM_ERR_SPLIT_RATIO; x.valid = false; }
                else { x.r.e *= ratio; s.e *= 1.0 - ratio; }
            }
        }
        if (o.out.px != nullptr) store_ray<true>(o.out, x.i, s, was_valid && x.valid ? RAY_VALID : RAY_INVALID);
    } else if (o.out.px != nullptr) {
        if (x.warp_live) store_ray<true>(o.out, x.i, x.r, RAY_REMOVED);  // full sectors
        else o.out.valid[x.i] = RAY_REMOVED;
    }
}
// OPM_OP_FORK / OPM_OP_JOIN: Rays::split (rays.rs:1253-1264, ray.rs:752-772) whose split-off bundle runs its side chain in the
// same launch.  The main-bundle ray (E * ratio) is parked in shared memory, one 32-bit word per field and thread
// (field-major, so consecutive lanes hit consecutive banks); the ops up to the JOIN see the split-off ray (E * (1 - ratio)).
__device__ __forceinline__ void stash_put(unsigned* stash, int f, double v) {
    stash[(2 * f) * OPM_BDIM + threadIdx.x] = (unsigned)__double2loint(v);
    stash[(2 * f + 1) * OPM_BDIM + threadIdx.x] = (unsigned)__double2hiint(v);
}
__device__ __forceinline__ double stash_get(const unsigned* stash, int f) {
    return __hiloint2double((int)stash[(2 * f + 1) * OPM_BDIM + threadIdx.x], (int)stash[(2 * f) * OPM_BDIM + threadIdx.x]);
}
// PARK: which fields of the parent the side chain can change (STASH_*; energy and flags are always parked).  The interpreter parks
// everything; the specialised kernel parks what the ops between its FORK and JOIN can touch, the rest of the parent is still in
// the registers when the JOIN comes (fewer shared-memory words per thread: more L1, fewer stores and loads).
enum : int { STASH_POS = 1, STASH_DIR = 2, STASH_OPL = 4, STASH_N = 8, STASH_BOUNCES = 16, STASH_REFR = 32, STASH_ALL = 63 };
__host__ __device__ constexpr int stash_words(int park) {
    return 3 + ((park & STASH_POS) ? 6 : 0) + ((park & STASH_DIR) ? 6 : 0) + ((park & STASH_OPL) ? 2 : 0) + ((park & STASH_N) ? 2 : 0) +
           ((park & STASH_BOUNCES) ? 1 : 0) + ((park & STASH_REFR) ? 1 : 0);
}
// word offset of a field's slot: energy 0-1, flags 2, then the parked fields in the order of their bits
__host__ __device__ constexpr int stash_off(int park, int field) { return stash_words(park & (field - 1)); }
__device__ __forceinline__ void stash_putw(unsigned* stash, int w, double v) {
    stash[w * OPM_BDIM + threadIdx.x] = (unsigned)__double2loint(v);
    stash[(w + 1) * OPM_BDIM + threadIdx.x] = (unsigned)__double2hiint(v);
}
__device__ __forceinline__ double stash_getw(const unsigned* stash, int w) {
    return __hiloint2double((int)stash[(w + 1) * OPM_BDIM + threadIdx.x], (int)stash[w * OPM_BDIM + threadIdx.x]);
}
template <bool SPECTRAL = true, int PARK = STASH_ALL>
__device__ __forceinline__ void exec_fork(const DevFork& o, RayCtx& x, unsigned* stash) {
    double ratio = o.ratio;
    if (SPECTRAL && x.valid && spectral_factor<true>(o.s, o.ratio, x, ratio) && o.s.lam != nullptr && !(ratio >= 0.0 && ratio <= 1.0)) {
        x.c.status |= 1u << OPM_ERR_SPLIT_RATIO; x.valid = false;
    }
    const double e_main = x.valid ? x.r.e * ratio : x.r.e;  // invalid rays are copied unchanged (rays.rs:1258-1262)
    stash_putw(stash, 0, e_main);
    stash[2 * OPM_BDIM + threadIdx.x] = (x.alive ? 1u : 0u) | (x.valid ? 2u : 0u);
    if (PARK & STASH_POS) { const int w = stash_off(PARK, STASH_POS); stash_putw(stash, w, x.r.pos.x); stash_putw(stash, w + 2, x.r.pos.y); stash_putw(stash, w + 4, x.r.pos.z); }
    if (PARK & STASH_DIR) { const int w = stash_off(PARK, STASH_DIR); stash_putw(stash, w, x.r.dir.x); stash_putw(stash, w + 2, x.r.dir.y); stash_putw(stash, w + 4, x.r.dir.z); }
    if (PARK & STASH_OPL) stash_putw(stash, stash_off(PARK, STASH_OPL), x.r.opl);
    if (PARK & STASH_N) stash_putw(stash, stash_off(PARK, STASH_N), x.r.n);
    if (PARK & STASH_BOUNCES) stash[stash_off(PARK, STASH_BOUNCES) * OPM_BDIM + threadIdx.x] = x.r.bounces;
    if (PARK & STASH_REFR) stash[stash_off(PARK, STASH_REFR) * OPM_BDIM + threadIdx.x] = x.r.refr;
    if (x.valid) x.r.e *= 1.0 - ratio;
}
template <int PARK = STASH_ALL>
__device__ __forceinline__ void exec_join(const DevJoin& o, RayCtx& x, const unsigned* stash) {
    if (x.in_range && o.out.px != nullptr) {
        if (x.alive || x.warp_live) store_ray<true>(o.out, x.i, x.r, !x.alive ? RAY_REMOVED : x.valid ? RAY_VALID : RAY_INVALID);
        else o.out.valid[x.i] = RAY_REMOVED;
    }
    x.r.e = stash_getw(stash, 0);
    const unsigned fl = stash[2 * OPM_BDIM + threadIdx.x];
    if (PARK & STASH_POS) { const int w = stash_off(PARK, STASH_POS); x.r.pos = v3(stash_getw(stash, w), stash_getw(stash, w + 2), stash_getw(stash, w + 4)); }
    if (PARK & STASH_DIR) { const int w = stash_off(PARK, STASH_DIR); x.r.dir = v3(stash_getw(stash, w), stash_getw(stash, w + 2), stash_getw(stash, w + 4)); }
    if (PARK & STASH_OPL) x.r.opl = stash_getw(stash, stash_off(PARK, STASH_OPL));
    if (PARK & STASH_N) x.r.n = stash_getw(stash, stash_off(PARK, STASH_N));
    if (PARK & STASH_BOUNCES) x.r.bounces = stash[stash_off(PARK, STASH_BOUNCES) * OPM_BDIM + threadIdx.x];
    if (PARK & STASH_REFR) x.r.refr = stash[stash_off(PARK, STASH_REFR) * OPM_BDIM + threadIdx.x];
    x.alive = (fl & 1u) != 0; x.valid = (fl & 2u) != 0;
}
// copy of the bundle a detector node keeps (analyzers/raytrace.rs:186-205).  LEAN: only what a spot-diagram report and the bundle
// statistics read (nodes/spot_diagram.rs:101-163, rays.rs:521-701: positions, energies, which rays are valid, id and bounce count)
template <bool LEAN = false>
__device__ __forceinline__ void exec_snapshot_store(const DevSnapshot& o, RayCtx& x) {
    if (!x.in_range) return;
    if (x.alive || x.warp_live) {  // a removed ray: full sectors, flagged RAY_REMOVED
        const uint8_t flag = !x.alive ? RAY_REMOVED : x.valid ? RAY_VALID : RAY_INVALID;
        if (LEAN) {
            OPM_ST(o.out.px + x.i, x.r.pos.x); OPM_ST(o.out.py + x.i, x.r.pos.y); OPM_ST(o.out.pz + x.i, x.r.pos.z);
            OPM_ST(o.out.e + x.i, x.r.e);
            OPM_ST(o.out.id + x.i, x.r.id); OPM_ST(o.out.bounces + x.i, x.r.bounces);  // Rays::bounce_lvl of the copy (rays.rs:183-194)
            o.out.valid[x.i] = flag;
        } else store_ray<true>(o.out, x.i, x.r, flag);
    } else o.out.valid[x.i] = RAY_REMOVED;
}
__device__ __forceinline__ void exec_snapshot(const DevSnapshot& o, RayCtx& x) {  // interpreter: structural facts read at run time
    if (o.fields == OPM_SNAPSHOT_SPOT) exec_snapshot_store<true>(o, x);
    else exec_snapshot_store<false>(o, x);
}
// Phase-1 sums of the spot statistics (rays.rs:599-637: centroid and energy-weighted centroid of the valid rays, their number,
// the first valid ray) taken while the copy is written.  Running sums per thread in dynamic shared memory, field-major
// ([x | y | z | ex | ey | ez | e][OPM_BDIM] doubles, then [OPM_BDIM] keys): one reduction per CTA at the end of the kernel, one
// set of atomics into the SpotAccum of the copy.  The ray counts go through two CTA counters (one ballot per warp and tile).
__device__ __forceinline__ void spot_smem_init(unsigned char* slot, unsigned* s_n) {
    double* d = reinterpret_cast<double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 7; ++j) d[j * B + t] = 0.0;
    reinterpret_cast<unsigned long long*>(d + 7 * B)[t] = ~0ull;
    if (t < 2) s_n[t] = 0u;
}
// HAS_FRAME / ROT_ID: positions go through the isometry `inv` / which is a pure translation
template <bool HAS_FRAME, bool ROT_ID>
__device__ __forceinline__ void spot_smem_add(const DevSnapshot& o, unsigned char* slot, unsigned* s_n, const RayCtx& x) {
    const bool v = x.in_range && x.alive && x.valid;
    const unsigned am = __activemask();  // the live and the dead copy of a specialised program call this separately
    const unsigned mv = __ballot_sync(am, v), ma = __ballot_sync(am, x.in_range && x.alive);
    if ((threadIdx.x & 31u) == (unsigned)(__ffs(am) - 1)) {
        if (mv) red_shared_add(&s_n[0], (unsigned)__popc(mv));
        if (ma) red_shared_add(&s_n[1], (unsigned)__popc(ma));
    }
    if (!v) return;
    V3 p = x.r.pos;
    if (HAS_FRAME) p = ROT_ID ? v3(p.x + o.inv.t[0], p.y + o.inv.t[1], p.z + o.inv.t[2]) : iso_point(o.inv, p);
    double* d = reinterpret_cast<double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
    const double e = x.r.e;
    d[t] += p.x; d[B + t] += p.y; d[2 * B + t] += p.z;
    d[3 * B + t] = fma(p.x, e, d[3 * B + t]); d[4 * B + t] = fma(p.y, e, d[4 * B + t]); d[5 * B + t] = fma(p.z, e, d[5 * B + t]);
    d[6 * B + t] += e;
    unsigned long long* k = reinterpret_cast<unsigned long long*>(d + 7 * B) + t;
    const unsigned long long key = ((unsigned long long)x.r.id << 32) | x.r.bounces;
    if (key < *k) *k = key;
}
// all threads of the CTA, after the tile loop and a barrier; ends with a barrier.  SpotAccum layout: sum[7] | n_valid | n_total |
// first_key (opm_kernels.h)
__device__ __forceinline__ void spot_smem_flush(const unsigned char* slot, const unsigned* s_n, void* acc_) {
    __shared__ double sh[32][7];
    __shared__ unsigned long long shk[32];
    const double* d = reinterpret_cast<const double*>(slot);
    const unsigned B = OPM_BDIM, t = threadIdx.x;
    double s[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) s[j] = d[j * B + t];
    unsigned long long key = reinterpret_cast<const unsigned long long*>(d + 7 * B)[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 7; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    const int w = t >> 5, l = t & 31;
    if (l == 0) {
#pragma unroll
        for (int j = 0; j < 7; ++j) sh[w][j] = s[j];
        shk[w] = key;
    }
    __syncthreads();
    if (t == 0 && s_n[1]) {
        double* sum = reinterpret_cast<double*>(acc_);
        unsigned long long* cnt = reinterpret_cast<unsigned long long*>(sum + 7);
        for (int k = 1; k < (int)((B + 31) >> 5); ++k) {
#pragma unroll
            for (int j = 0; j < 7; ++j) s[j] += sh[k][j];
            key = shk[k] < key ? shk[k] : key;
        }
        if (s_n[0]) {
#pragma unroll
            for (int j = 0; j < 7; ++j) atomicAdd(&sum[j], s[j]);
            atomicAdd(&cnt[0], (unsigned long long)s_n[0]);
            atomicMin(&cnt[2], key);
        }
        atomicAdd(&cnt[1], (unsigned long long)s_n[1]);
    }
    __syncthreads();
}
// ray.rs:787-804
__device__ __forceinline__ void exec_transform(const DevTransform& o, RayCtx& x) {
    if (!x.alive) return;
    x.r.pos = iso_point(o.iso, x.r.pos);
    x.r.dir = iso_vector(o.iso, x.r.dir);
}

// ---- fluence helper rays -------------------------------------------------------------------------
// FluenceRays of one ray (rays.rs:1470-1516) in registers / local memory for the length of a program
struct HelperState {
    V3 pos[3], dir[3];
    double n[3];   // negative: the helper has been set invalid (a missed surface under MissedSurfaceStrategy::Stop)
    double eff;    // FluenceRays::effective_energy
};
__device__ __forceinline__ void helpers_load(const HelperParams& H, uint64_t i, HelperState& h) {
    const double* b = H.in + i;
    const uint64_t s = H.in_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        h.pos[k] = v3(b[(3 * k) * s], b[(3 * k + 1) * s], b[(3 * k + 2) * s]);
        h.dir[k] = v3(b[(9 + 3 * k) * s], b[(9 + 3 * k + 1) * s], b[(9 + 3 * k + 2) * s]);
        h.n[k] = b[(18 + k) * s];
    }
    h.eff = b[21 * s];
}
__device__ __forceinline__ void helpers_store(const HelperParams& H, uint64_t i, const HelperState& h) {
    double* b = H.out + i;
    const uint64_t s = H.out_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b[(3 * k) * s] = h.pos[k].x; b[(3 * k + 1) * s] = h.pos[k].y; b[(3 * k + 2) * s] = h.pos[k].z;
        b[(9 + 3 * k) * s] = h.dir[k].x; b[(9 + 3 * k + 1) * s] = h.dir[k].y; b[(9 + 3 * k + 2) * s] = h.dir[k].z;
        b[(18 + k) * s] = h.n[k];
    }
    b[21 * s] = h.eff;
}
// Ray::helper_ray_fluence (ray.rs:172-198): effective energy over the area of the triangle the three helpers span
__device__ __forceinline__ double helper_ray_fluence(const HelperState& h) {
    const V3 ab = v3(h.pos[0].x - h.pos[1].x, h.pos[0].y - h.pos[1].y, h.pos[0].z - h.pos[1].z);
    const V3 ac = v3(h.pos[0].x - h.pos[2].x, h.pos[0].y - h.pos[2].y, h.pos[0].z - h.pos[2].z);
    const double cx = ab.y * ac.z - ab.z * ac.y, cy = ab.z * ac.x - ab.x * ac.z, cz = ab.x * ac.y - ab.y * ac.x;
    const double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
    return h.eff / area;
}
// Rays::refract_on_surface for a ray with FluenceRays (rays.rs:987-1012 with ray.rs:640-672).  In the reference's order: the main
// ray is refracted first -- its hit point carries the fluence of the helper triangle AS IT STILL IS (the helpers have not reached
// this surface yet), the effective energy takes the factor 1 - R (R for the reflected clone) -- then, if there is a reflected
// clone, each helper is refracted on the same surface.  The clone keeps the helpers where they were and only takes their
// reflected directions (rays.rs:1000-1010): that is what a mirror's continuing bundle carries.
// Returns like exec_refract.
template <class V>
__device__ __forceinline__ bool exec_refract_helpers(const V& v, RayCtx& x, HelperState& h) {
    const DevRefract& o = v.d();
    RefractResult res;
    res.has_child = false; res.moved = false; res.hit_e = -1.0; res.hit_bounce = 0; res.hit_local = v3(0, 0, 0);
    double fluence = 0.0;
    if (x.valid) {
        x.c.a += 1u;
        fluence = helper_ray_fluence(h);
        RayState shadow = x.r;  // the same refraction with the effective energy in the place of the ray's: eff (1 - R) and eff R
        shadow.e = h.eff;
        res = refract_on_surface(v, x.r);
        if (res.status != OPM_OK) { x.c.status |= 1u << res.status; x.valid = false; res.has_child = false; res.hit_e = -1.0; }
        else {
            if (res.missed) x.c.b += 1u;
            if (res.invalidated) x.valid = false;
        }
        if (res.has_child) {
            // FluenceHitPoint::new (hit_point.rs): the fluence must be finite and not negative
            if (!isfinite(fluence) || signbit_(fluence)) { x.c.status |= 1u << OPM_ERR_HITPOINT; x.valid = false; res.has_child = false; res.hit_e = -1.0; }
        }
        if (res.has_child) {
            res.hit_e = fluence;
            const RefractResult rs = refract_on_surface(v, shadow);
            const bool mirror = (v.flags() & OPM_REFRACT_CONTINUE_AS_CHILD) != 0;
            h.eff = mirror ? rs.child_e : shadow.e;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                RayState hr = x.r;  // wavelength of the ray; the other fields below
                hr.pos = h.pos[k]; hr.dir = h.dir[k]; hr.n = fabs(h.n[k]); hr.e = 1.0; hr.opl = 0.0; hr.bounces = 0; hr.refr = 0;
                const RefractResult rh = refract_on_surface(v, hr);
                // (a helper records no hit point: a hit-point error cannot come from it; the index / coating errors are the ray's own)
                if (!mirror) {
                    h.pos[k] = hr.pos; h.dir[k] = hr.dir;
                    const bool inval = h.n[k] < 0.0 || (rh.status == OPM_OK && rh.invalidated);
                    h.n[k] = inval ? -hr.n : hr.n;
                } else if (rh.status == OPM_OK && rh.has_child) h.dir[k] = rh.child_dir;
            }
        }
    }
    if (v.has_hits() && x.in_range) {  // slot-indexed hit record, e < 0 = empty; here e is the FLUENCE of a FluenceHitPoint
        OPM_ST(o.hits.e + x.i, res.hit_e);
        if (res.hit_e >= 0.0 || x.warp_live) {
            OPM_ST(o.hits.x + x.i, res.hit_local.x); OPM_ST(o.hits.y + x.i, res.hit_local.y);
            OPM_ST(o.hits.z + x.i, res.hit_local.z); OPM_ST(o.hits.bounce + x.i, res.hit_bounce);
        }
        if (res.hit_e >= 0.0) x.c.a += 1u << 16;
    }
    if (v.flags() & OPM_REFRACT_STAT_KEY) { x.c.kid = (x.valid && x.alive) ? x.r.id : 0xffffffffu; x.c.kb = x.r.bounces; }
#ifdef OPM_STAT_KEY2
    if (v.flags() & OPM_REFRACT_STAT_KEY2) { x.c.kid2 = (x.valid && x.alive) ? x.r.id : 0xffffffffu; x.c.kb2 = x.r.bounces; }
#endif
    finish_children(v, res, x);
    return res.moved;
}
// Rays::refract_paraxial (rays.rs:943-958): the ray if it is valid, and its valid helpers whether it is or not
__device__ __forceinline__ void exec_paraxial_helpers(const DevParaxial& o, RayCtx& x, HelperState& h) {
    if (x.valid) refract_paraxial(o, x.r);
    if (!x.alive) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (h.n[k] < 0.0) continue;
        RayState hr = x.r;
        hr.pos = h.pos[k]; hr.dir = h.dir[k]; hr.opl = 0.0; hr.refr = 0;
        refract_paraxial(o, hr);
        h.pos[k] = hr.pos; h.dir[k] = hr.dir;
    }
}

// ---- tile prologue / epilogue shared by both kernels -----------------------------------------------------------
__device__ __forceinline__ void tile_load(const DevRays& in, uint64_t i, uint64_t n, RayCtx& x, uint8_t& vb) {
    x.i = i;
    x.in_range = i < n;
    x.c.a = 0; x.c.b = 0; x.c.apodized = 0; x.c.status = 0; x.c.kid = kKeyUnset; x.c.kb = 0; OPM_KEY2_RESET(x.c);
    x.r.pos = v3(0, 0, 0); x.r.dir = v3(0, 0, 1); x.r.e = 0; x.r.wvl = 0; x.r.opl = 0; x.r.n = 1; x.r.bounces = 0; x.r.refr = 0; x.r.id = 0;
    vb = RAY_REMOVED;
    if (x.in_range) {
#ifndef OPM_TILE_INDEPENDENT_LOADS
        vb = in.valid[i];
        if (vb != RAY_REMOVED) load_ray(in, i, x.r);
#else   // probe (profiles/tile_load_variants.sh): the 13 field loads do not wait for the flag -- one round trip per tile instead
        // of two.  Measured on C2: 1.335 ms against 1.318 ms; the other resident warps already cover the second round trip.
        vb = in.valid[i];
        load_ray(in, i, x.r);
        if (vb == RAY_REMOVED) {
            x.r.pos = v3(0, 0, 0); x.r.dir = v3(0, 0, 1); x.r.e = 0; x.r.wvl = 0; x.r.opl = 0; x.r.n = 1; x.r.bounces = 0; x.r.refr = 0; x.r.id = 0;
        }
#endif
    }
    x.alive = vb != RAY_REMOVED;
    x.valid = vb == RAY_VALID;
    x.warp_live = __any_sync(0xffffffffu, x.alive);
}
// The flag of the ray this thread takes in the CTA's NEXT tile is loaded one tile ahead: the tile prologue then waits for one
// round trip (the 13 field loads) instead of two (flag, then fields) and still loads nothing for a slot that holds no ray.
// One register; nothing else writes the flag of a slot before its own thread gets to it.  Specialised kernel only; measured on
// C2 (profiles/r2m_tile_variants.sh): 0.969 -> 0.959 ms alone, 0.916 -> 0.915 ms beside the bulk prefetch.  -DOPM_TILE_NO_FLAG_AHEAD
// restores the dependent flag load.
struct TileAhead { uint8_t vb; };
__device__ __forceinline__ void tile_ahead_init(TileAhead& a, const DevRays& in, uint64_t i, uint64_t n) {
    a.vb = RAY_REMOVED;
#ifndef OPM_TILE_NO_FLAG_AHEAD
    if (i < n) a.vb = in.valid[i];
#endif
}
__device__ __forceinline__ void tile_load(const DevRays& in, uint64_t i, uint64_t n, RayCtx& x, uint8_t& vb, TileAhead& a, uint64_t i_next) {
#ifdef OPM_TILE_NO_FLAG_AHEAD
    tile_load(in, i, n, x, vb);
#else
    x.i = i;
    x.in_range = i < n;
    x.c.a = 0; x.c.b = 0; x.c.apodized = 0; x.c.status = 0; x.c.kid = kKeyUnset; x.c.kb = 0; OPM_KEY2_RESET(x.c);
    x.r.pos = v3(0, 0, 0); x.r.dir = v3(0, 0, 1); x.r.e = 0; x.r.wvl = 0; x.r.opl = 0; x.r.n = 1; x.r.bounces = 0; x.r.refr = 0; x.r.id = 0;
    vb = a.vb;  // RAY_REMOVED beyond the bundle
    if (vb != RAY_REMOVED) load_ray(in, i, x.r);
    a.vb = RAY_REMOVED;
    if (i_next < n) a.vb = in.valid[i_next];
    x.alive = vb != RAY_REMOVED;
    x.valid = vb == RAY_VALID;
    x.warp_live = __any_sync(0xffffffffu, x.alive);
#endif
}
// probe (-DOPM_TILE_PREFETCH, profiles/tile_load_variants.sh): asks the L2 for the lines of the tile this CTA takes next (one
// request per 128-byte line).  Measured on C2: 1.359 ms against 1.318 ms -- left off.
// Specialised kernel (default; -DOPM_TILE_NO_BULK_PREFETCH turns it off): the same request as 14 bulk prefetches
// (cp.async.bulk.prefetch.L2 -> UBLKPF, one per field segment of the tile) from ONE thread of the CTA instead of two instructions
// per field from every warp: when the warps get to the tile, their loads hit the L2 instead of waiting for the HBM.  Measured on
// C2 (profiles/r2m_tile_variants.sh): 0.969 -> 0.916 ms per launch (per-warp prefetches: 0.929 ms).
#ifndef OPM_TILE_PREFETCH_DIST
#define OPM_TILE_PREFETCH_DIST 1  // tiles of this CTA ahead
#endif
__device__ __forceinline__ void tile_prefetch(const DevRays& in, uint64_t i, uint64_t n) {
#if !defined(OPM_TILE_NO_BULK_PREFETCH) && defined(OPM_JIT_BLOCK) && defined(OPM_TILE_BULK_PREFETCH_PER_WARP)
    // probe: every warp asks for its own 32 rays of the next tile (14 requests of 256 / 128 / 32 bytes from lane 0) -- the request
    // then leads the warp's own loads by exactly one tile of that warp, however far the warps of the CTA have drifted apart.
    // Measured on C2 (profiles/r2y_variants.txt): 0.958 ms against 0.915 ms for the 14 requests per CTA -- left off.
    if ((threadIdx.x & 31u) == 0u && i + 32 <= n) {
        const double* f64[10] = {in.px, in.py, in.pz, in.dx, in.dy, in.dz, in.e, in.wvl, in.opl, in.n};
#pragma unroll
        for (int f = 0; f < 10; ++f) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(f64[f] + i), "r"(256));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.bounces + i), "r"(128));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.refr + i), "r"(128));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.id + i), "r"(128));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.valid + i), "r"(32));
    }
#elif !defined(OPM_TILE_NO_BULK_PREFETCH) && defined(OPM_JIT_BLOCK)
    if (threadIdx.x == 0 && i + OPM_JIT_BLOCK <= n) {  // i = first ray of the next tile; whole tiles only (16-byte granules)
        static_assert(OPM_JIT_BLOCK % 16 == 0, "bulk prefetch takes multiples of 16 bytes");
        const double* f64[10] = {in.px, in.py, in.pz, in.dx, in.dy, in.dz, in.e, in.wvl, in.opl, in.n};
#pragma unroll
        for (int f = 0; f < 10; ++f) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(f64[f] + i), "r"(OPM_JIT_BLOCK * 8));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.bounces + i), "r"(OPM_JIT_BLOCK * 4));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.refr + i), "r"(OPM_JIT_BLOCK * 4));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.id + i), "r"(OPM_JIT_BLOCK * 4));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.valid + i), "r"(OPM_JIT_BLOCK));
    }
#endif
#ifdef OPM_TILE_PREFETCH
    if (i >= n) return;
    const unsigned lane = threadIdx.x & 31u;
    if ((lane & 15u) == 0u) {
        const double* f64[10] = {in.px, in.py, in.pz, in.dx, in.dy, in.dz, in.e, in.wvl, in.opl, in.n};
#pragma unroll
        for (int f = 0; f < 10; ++f) asm volatile("prefetch.global.L2 [%0];" ::"l"(f64[f] + i));
        if (lane == 0u) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(in.bounces + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(in.refr + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(in.id + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(in.valid + i));
        }
    }
#endif
}

// ---- staged tile loads (run-time specialised kernel) ---------------------------------------------------------------
// One elected thread copies whole tiles of the bundle -- 14 contiguous field segments of OPM_BDIM rays -- into shared
// memory with TMA bulk copies (cp.async.bulk -> UBLKCP) two tiles ahead of the compute; the CTA's warps find their ray
// state on chip when they get to the tile instead of waiting on 14 global loads each.  Two stages, one mbarrier per
// stage; a stage is re-armed after a block barrier that follows the last read of it.  A trailing partial tile takes the
// direct path (tile_load).
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {  // try_wait suspends the thread in hardware until the phase completes or a time limit passes
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
constexpr int kStageBytesPerRay = 10 * 8 + 3 * 4 + 1;  // = kRayBytes, laid out field-major in a stage
struct TilePipe {
    unsigned char* stage[2];
    unsigned long long* bar;  // [2]
    unsigned iter;            // tiles of this CTA consumed so far through the stages
    unsigned tiles;           // number of tiles of the launch
    unsigned full_tiles;      // tiles [0, full_tiles) hold OPM_BDIM rays
};
// thread 0: arm the stage's barrier and start the 14 copies of one full tile
__device__ __forceinline__ void tile_issue(const DevRays& in, unsigned tile, unsigned char* st, unsigned long long* bar) {
    const unsigned B = OPM_BDIM;
    const size_t off = (size_t)tile * B;
    mbar_expect_tx(bar, B * kStageBytesPerRay);
    const double* f64[10] = {in.px, in.py, in.pz, in.dx, in.dy, in.dz, in.e, in.wvl, in.opl, in.n};
#pragma unroll
    for (int f = 0; f < 10; ++f) bulk_g2s(st + (size_t)f * B * 8, f64[f] + off, B * 8, bar);
    const uint32_t* u32[3] = {in.bounces, in.refr, in.id};
#pragma unroll
    for (int f = 0; f < 3; ++f) bulk_g2s(st + (size_t)B * 80 + (size_t)f * B * 4, u32[f] + off, B * 4, bar);
    bulk_g2s(st + (size_t)B * 92, in.valid + off, B, bar);
}
__device__ __forceinline__ void tile_pipe_init(TilePipe& p, unsigned char* smem, unsigned long long* bar, const DevRays& in, uint64_t n) {
    const unsigned B = OPM_BDIM;
    const unsigned stage_bytes = (B * kStageBytesPerRay + 127u) & ~127u;
    p.stage[0] = smem; p.stage[1] = smem + stage_bytes;
    p.bar = bar; p.iter = 0;
    p.tiles = (unsigned)((n + B - 1) / B);
    p.full_tiles = (unsigned)(n / B);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t0 = blockIdx.x, t1 = blockIdx.x + gridDim.x;
        if (t0 < p.full_tiles) tile_issue(in, t0, p.stage[0], &bar[0]);
        if (t1 < p.full_tiles) tile_issue(in, t1, p.stage[1], &bar[1]);
    }
}
__device__ __forceinline__ void tile_load_staged(TilePipe& p, const DevRays& in, unsigned tile, uint64_t n, RayCtx& x, uint8_t& vb) {
    const uint64_t i = (uint64_t)tile * OPM_BDIM + threadIdx.x;
    if (tile >= p.full_tiles) { tile_load(in, i, n, x, vb); return; }  // the trailing partial tile
    const unsigned s = p.iter & 1u, B = OPM_BDIM, t = threadIdx.x;
    unsigned char* st = p.stage[s];
    mbar_wait(&p.bar[s], (p.iter >> 1) & 1u);
    x.i = i; x.in_range = true;
    x.c.a = 0; x.c.b = 0; x.c.apodized = 0; x.c.status = 0; x.c.kid = kKeyUnset; x.c.kb = 0; OPM_KEY2_RESET(x.c);
    const double* d = reinterpret_cast<const double*>(st);
    x.r.pos = v3(d[t], d[B + t], d[2 * B + t]);
    x.r.dir = v3(d[3 * B + t], d[4 * B + t], d[5 * B + t]);
    x.r.e = d[6 * B + t]; x.r.wvl = d[7 * B + t]; x.r.opl = d[8 * B + t]; x.r.n = d[9 * B + t];
    const uint32_t* u = reinterpret_cast<const uint32_t*>(st + (size_t)B * 80);
    x.r.bounces = u[t]; x.r.refr = u[B + t]; x.r.id = u[2 * B + t];
    vb = st[(size_t)B * 92 + t];
    if (vb == RAY_REMOVED) {  // what tile_load leaves for a slot that is not part of the bundle
        x.r.pos = v3(0, 0, 0); x.r.dir = v3(0, 0, 1); x.r.e = 0; x.r.wvl = 0; x.r.opl = 0; x.r.n = 1; x.r.bounces = 0; x.r.refr = 0; x.r.id = 0;
    }
    x.alive = vb != RAY_REMOVED;
    x.valid = vb == RAY_VALID;
    x.warp_live = __any_sync(0xffffffffu, x.alive);
    __syncthreads();  // every thread has its ray in registers: the stage can take the tile after next
    if (t == 0) {
        const unsigned next = tile + 2u * gridDim.x;
        if (next < p.full_tiles) tile_issue(in, next, st, &p.bar[s]);
    }
    p.iter++;
}

template <bool COMPACT>
__device__ __forceinline__ void tile_store(const DevRays& in, const DevRays& out, uint64_t out_cap, unsigned long long* out_tail,
                                           RayCtx& x, uint8_t vb) {
    if (COMPACT) {
        unsigned long long slot = warp_reserve(x.alive, out_tail);
        if (x.alive) {
            if (slot < out_cap) store_ray<true>(out, slot, x.r, x.valid ? RAY_VALID : RAY_INVALID);
            else x.c.status |= 1u << OPM_ERR_CAPACITY;
        }
    } else if (x.in_range) {
        if (x.alive || x.warp_live) {  // a removed ray: full sectors, flagged RAY_REMOVED (RayCtx::warp_live)
            const uint8_t flag = !x.alive ? RAY_REMOVED : x.valid ? RAY_VALID : RAY_INVALID;
            if (out.px == in.px) store_ray<false>(out, x.i, x.r, flag);
            else store_ray<true>(out, x.i, x.r, flag);
        } else if (vb != RAY_REMOVED || out.px != in.px) {
            out.valid[x.i] = RAY_REMOVED;
        }
    }
}

// tile counters: one warp reduction per register, then lane k adds field k to THIS WARP's row of the CTA's counter table in
// shared memory -- a plain load / add / store of 7 lanes, no atomics (an atomic add from divergent code is compiled into the
// warp-aggregated form, a dozen instructions per counter).  32-bit counters: a warp sees at most rays / warps x ops
// interactions, below 2^32 for any bundle that fits the device.  cta_flush sums the rows.
typedef unsigned cta_count_t;
constexpr int kCntStride = 8;                  // CNT_N fields per row, padded
constexpr int kCntWords = 32 * kCntStride;     // one row per warp of the largest CTA
__device__ __forceinline__ void cta_count_init(cta_count_t* s_cnt) {
    for (unsigned k = threadIdx.x; k < (unsigned)kCntWords; k += blockDim.x) s_cnt[k] = 0u;
}
// warp_best: the smallest ray id this warp has offered for the launch's first valid ray so far (0xffffffff before its first tile);
// a tile none of whose rays has an id at or below it cannot change the statistic and skips the two warp minima
__device__ __forceinline__ void tile_count(const RayCtx& x, cta_count_t* s_cnt, unsigned* s_status, unsigned long long* s_first_inv,
                                           unsigned& warp_best, unsigned long long* s_first_inv2) {
#ifdef OPM_STAT_KEY2
    if (x.c.kid2 != kKeyUnset) {  // warp-uniform
        if (__any_sync(0xffffffffu, x.c.kid2 != 0xffffffffu)) {
            const unsigned idmin = __reduce_min_sync(0xffffffffu, x.c.kid2);
            const unsigned bmin = __reduce_min_sync(0xffffffffu, x.c.kid2 == idmin ? x.c.kb2 : 0xffffffffu);
            if ((threadIdx.x & 31) == 0) {
                const unsigned long long inv = ~(((unsigned long long)idmin << 32) | bmin);
                if (inv > *(volatile unsigned long long*)s_first_inv2) atomicMax(s_first_inv2, inv);
            }
        }
    }
#else
    (void)s_first_inv2;
#endif
    const unsigned ra = __reduce_add_sync(0xffffffffu, x.c.a), rb = __reduce_add_sync(0xffffffffu, x.c.b);
    const unsigned rp = __reduce_add_sync(0xffffffffu, x.c.apodized);
    const bool v = x.valid && x.alive;
    const unsigned live = __ballot_sync(0xffffffffu, x.alive), ok = __ballot_sync(0xffffffffu, v);
    const unsigned st = __reduce_or_sync(0xffffffffu, x.c.status);
    const bool lane0 = (threadIdx.x & 31) == 0;
    const bool cap = x.c.kid != kKeyUnset;  // warp-uniform: the ops of a program are
    const unsigned kid = cap ? x.c.kid : (v ? x.r.id : 0xffffffffu), kb = cap ? x.c.kb : x.r.bounces;
    if (__any_sync(0xffffffffu, kid != 0xffffffffu && kid <= warp_best)) {  // first valid ray of the warp = lowest (id, bounces): two 32-bit warp minima
        const unsigned idmin = __reduce_min_sync(0xffffffffu, kid);
        warp_best = idmin;
        const unsigned bmin = __reduce_min_sync(0xffffffffu, (kid != 0xffffffffu && kid == idmin) ? kb : 0xffffffffu);
        if (lane0) {
            const unsigned long long inv = ~(((unsigned long long)idmin << 32) | bmin);  // largest for the first ray
            if (inv > *(volatile unsigned long long*)s_first_inv) atomicMax(s_first_inv, inv);  // rarely after the CTA's first tiles
        }
    }
    if (lane0) {  // the warp's own row (32 bytes, CNT_* order): two vector loads, seven adds, two vector stores
        uint4* row = reinterpret_cast<uint4*>(s_cnt + (threadIdx.x >> 5) * kCntStride);
        uint4 a = row[0], b = row[1];
        a.x += ra & 0xffffu; a.y += rb >> 16; a.z += ra >> 16; a.w += rb & 0xffffu;
        b.x += rp; b.y += (unsigned)__popc(ok); b.z += (unsigned)__popc(live);
        row[0] = a; row[1] = b;
        if (st) atomicOr(s_status, st);
    }
}
static_assert(CNT_INTERACTIONS == 0 && CNT_CHILDREN == 1 && CNT_HITS == 2 && CNT_MISSED == 3 && CNT_APODIZED == 4 && CNT_VALID_OUT == 5 &&
              CNT_ALIVE_OUT == 6 && kCntStride == 8, "tile_count writes the counter row as two uint4");

// CTA totals -> global statistics (call after __syncthreads())
__device__ __forceinline__ void cta_flush(DevStats* stats, const cta_count_t* s_cnt, unsigned s_status, unsigned long long s_first_inv,
                                          unsigned long long s_first_inv2) {
#ifdef OPM_STAT_KEY2
    if (threadIdx.x == 0 && s_first_inv2) atomicMax(&stats->first_inv2, s_first_inv2);
#else
    (void)s_first_inv2;
#endif
    if (threadIdx.x < CNT_N) {
        unsigned long long total = 0;
        for (int w = 0; w < 32; ++w) total += s_cnt[w * kCntStride + threadIdx.x];
        if (total) {
            // DevStats starts with the counters in the order interactions, children, hits, missed, apodized, valid_out, alive_out = CNT_*
            atomicAdd(&stats->interactions + threadIdx.x, total);
            if (threadIdx.x == CNT_INTERACTIONS) atomicAdd(&stats->total_interactions, total);
        }
    }
    if (threadIdx.x == 0 && s_status) { atomicOr(&stats->status_bits, s_status); atomicOr(&stats->sticky_bits, s_status); }
    if (threadIdx.x == 0 && s_first_inv) atomicMax(&stats->first_inv, s_first_inv);
}

}  // namespace opm
