// opm_trace.cu -- the fused surface-sequence kernel (sm_100a).
//
// One launch runs a whole program of `Rays` operations (refract_on_surface, apodize,
// invalidate_by_threshold_energy, filter limits, filter_energy, refract_paraxial, split, detector
// snapshot) over every ray of a bundle:
//   * ray state (10 f64 + 3 u32 + 1 u8 = 93 B) is read once from the struct-of-arrays bundle with
//     coalesced streaming loads, lives in registers across all ops and is written once;
//   * the op descriptors (surface isometries, index model, coating, aperture ...) are staged in shared
//     memory with ONE bulk async copy (TMA, cp.async.bulk -> UBLKCP) per CTA and read as warp-uniform
//     broadcasts; the CTA is persistent and walks ray tiles with a grid stride;
//   * reflected child rays (ghost focus) are appended to the child bundle with warp-aggregated atomics
//     (one atomicAdd per warp from the ballot of lanes that have a child);
//   * in compact mode the surviving rays are appended to the output bundle the same way, so rays that
//     left the bundle (mirror misses) cost no bandwidth downstream.
// No tensor cores: the path is FP64-pipe / HBM bound, not a contraction.
//
// Compiled twice into the same library: -DOPM_STRICT -fmad=false (reference operation order) and the
// default contracted build.  Reference semantics: rays.rs:976-1047, ray.rs:580-688,
// analyzers/raytrace.rs:96-140.
#include <cuda/barrier>
#include <cuda_runtime.h>

#include "opm_device.cuh"

#ifdef OPM_STRICT
#define OPM_VARIANT(name) name##_strict
#else
#define OPM_VARIANT(name) name##_fast
#endif

// CTA shape of the fused kernel: threads per CTA and the CTAs per SM the register allocation is bounded for
#ifndef OPM_TRACE_BLOCK
#define OPM_TRACE_BLOCK 256
#endif
#ifndef OPM_TRACE_MIN_CTAS
#define OPM_TRACE_MIN_CTAS 2
#endif

namespace opm {
namespace {

using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// warp-aggregated reservation of `want ? 1 : 0` slots per lane; returns this lane's slot
__device__ __forceinline__ unsigned long long warp_reserve(bool want, unsigned long long* tail) {
    unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return 0;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(tail, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & lanemask_lt());
}

__device__ __forceinline__ void load_ray(const DevRays& R, uint64_t i, RayState& r) {
    r.pos = v3(__ldcs(R.px + i), __ldcs(R.py + i), __ldcs(R.pz + i));
    r.dir = v3(__ldcs(R.dx + i), __ldcs(R.dy + i), __ldcs(R.dz + i));
    r.e = __ldcs(R.e + i); r.wvl = __ldcs(R.wvl + i); r.opl = __ldcs(R.opl + i); r.n = __ldcs(R.n + i);
    r.bounces = __ldcs(R.bounces + i); r.refr = __ldcs(R.refr + i); r.id = __ldcs(R.id + i);
}
template <bool FULL>
__device__ __forceinline__ void store_ray(const DevRays& R, uint64_t i, const RayState& r, uint8_t valid) {
    __stcs(R.px + i, r.pos.x); __stcs(R.py + i, r.pos.y); __stcs(R.pz + i, r.pos.z);
    __stcs(R.dx + i, r.dir.x); __stcs(R.dy + i, r.dir.y); __stcs(R.dz + i, r.dir.z);
    __stcs(R.e + i, r.e); __stcs(R.opl + i, r.opl); __stcs(R.n + i, r.n);
    __stcs(R.bounces + i, r.bounces); __stcs(R.refr + i, r.refr);
    if (FULL) { __stcs(R.wvl + i, r.wvl); __stcs(R.id + i, r.id); }  // wavelength and id never change in place
    R.valid[i] = valid;
}

// Per-tile event counters, two 16-bit fields per register (a warp's sum of a field is at most 32 lanes x 64 ops):
// a = interactions | hits << 16, b = missed | children << 16.  Each tile ends with one REDUX per register and one
// shared-memory atomic per field from lane 0; the CTA flushes its totals to the global statistics once.
struct Counters {
    unsigned a, b, apodized, status;
};
enum { CNT_INTERACTIONS = 0, CNT_CHILDREN, CNT_HITS, CNT_MISSED, CNT_APODIZED, CNT_VALID_OUT, CNT_ALIVE_OUT, CNT_N };

template <bool COMPACT>
__global__ void __launch_bounds__(OPM_TRACE_BLOCK, OPM_TRACE_MIN_CTAS)
trace_kernel(const DevOp* __restrict__ g_ops, int n_ops, DevRays in, DevRays out, uint64_t n, uint64_t out_cap,
             unsigned long long* out_tail, DevStats* stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DevOp* ops = reinterpret_cast<DevOp*>(smem_raw);
    __shared__ barrier_t bar;

    // ---- stage the program in shared memory: one TMA bulk copy per CTA
    const unsigned bytes = (unsigned)(n_ops * sizeof(DevOp));
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cuda::device::experimental::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier_t::arrival_token token;
    if (threadIdx.x == 0) {
        cuda::device::experimental::cp_async_bulk_global_to_shared(ops, g_ops, bytes, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));

    __shared__ unsigned long long s_cnt[CNT_N];
    __shared__ unsigned s_status;
    __shared__ unsigned long long s_first_inv;  // max of ~(id << 32 | bounces) over the valid rays this CTA wrote
    if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) { s_status = 0u; s_first_inv = 0ull; }
    __syncthreads();

    const unsigned tiles = (unsigned)((n + blockDim.x - 1) / blockDim.x);
    for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        Counters c = {0, 0, 0, 0};
        const uint64_t i = (uint64_t)tile * blockDim.x + threadIdx.x;
        const bool in_range = i < n;
        RayState r;
        r.pos = v3(0, 0, 0); r.dir = v3(0, 0, 1); r.e = 0; r.wvl = 0; r.opl = 0; r.n = 1; r.bounces = 0; r.refr = 0; r.id = 0;
        uint8_t vb = RAY_REMOVED;
        if (in_range) {
            vb = in.valid[i];
            if (vb != RAY_REMOVED) load_ray(in, i, r);
        }
        bool alive = vb != RAY_REMOVED;  // part of the bundle
        bool valid = vb == RAY_VALID;

        for (int k = 0; k < n_ops; ++k) {
            const DevOp& op = ops[k];
            switch (op.kind) {  // warp-uniform
                case OPK_REFRACT: {
                    const DevRefract& o = op.u.refract;
                    RefractResult res;
                    res.has_child = false; res.hit_e = -1.0; res.hit_bounce = 0; res.hit_local = v3(0, 0, 0);
                    if (valid) {
                        c.a += 1u;
                        res = refract_on_surface(o, r);
                        if (res.status != OPM_OK) { c.status |= 1u << res.status; valid = false; res.has_child = false; res.hit_e = -1.0; }
                        else {
                            if (res.missed) c.b += 1u;
                            if (res.invalidated) valid = false;
                        }
                    }
                    if (o.hits.e != nullptr && in_range) {  // slot-indexed hit record, e < 0 = empty
                        __stcs(o.hits.e + i, res.hit_e);
                        if (res.hit_e >= 0.0) {
                            __stcs(o.hits.x + i, res.hit_local.x); __stcs(o.hits.y + i, res.hit_local.y);
                            __stcs(o.hits.z + i, res.hit_local.z); __stcs(o.hits.bounce + i, res.hit_bounce);
                            c.a += 1u << 16;
                        }
                    }
                    if (o.flags & OPM_REFRACT_EMIT_CHILDREN) {
                        unsigned long long slot = warp_reserve(res.has_child, o.child_tail);
                        if (res.has_child) {
                            if (slot < o.child_cap) {
                                RayState child = r;
                                child_into(res, child);
                                store_ray<true>(o.children, slot, child, RAY_VALID);
                                c.b += 1u << 16;
                            } else c.status |= 1u << OPM_ERR_CAPACITY;
                        }
                    }
                    if (o.flags & OPM_REFRACT_CONTINUE_AS_CHILD) {  // mirror: only reflected rays stay in the bundle
                        if (alive) {
                            if (res.has_child) child_into(res, r);
                            else { alive = false; valid = false; }
                        }
                    }
                    break;
                }
                case OPK_APODIZE: {  // rays.rs:557-573
                    if (valid) {
                        const DevApodize& o = op.u.apodize;
                        double x = o.inv.r[0] * r.pos.x + o.inv.r[1] * r.pos.y + o.inv.r[2] * r.pos.z + o.inv.t[0];
                        double y = o.inv.r[3] * r.pos.x + o.inv.r[4] * r.pos.y + o.inv.r[5] * r.pos.z + o.inv.t[1];
                        double f = aperture_factor(o.ap, x, y);
                        if (f > 0.0) {
                            if (!(f <= 1.0)) c.status |= 1u << OPM_ERR_APODIZE_FACTOR;
                            else r.e *= f;
                        } else { valid = false; c.apodized += 1u; }
                    }
                    break;
                }
                case OPK_THRESHOLD:  // rays.rs:1145-1162 (all rays; a no-op for already invalid ones)
                    if (alive && r.e < op.u.threshold.min_energy) valid = false;
                    break;
                case OPK_LIMITS: {  // rays.rs:1386-1404
                    const DevLimits& o = op.u.limits;
                    if (alive) {
                        if ((uint64_t)r.bounces > o.max_bounces) valid = false;
                        if (o.check_refr && (uint64_t)r.refr >= o.max_refr) valid = false;
                    }
                    break;
                }
                case OPK_FILTER:  // rays.rs:1119-1133
                    if (valid) r.e *= op.u.filter.t;
                    break;
                case OPK_PARAXIAL:  // rays.rs:943-959
                    if (valid) refract_paraxial(op.u.paraxial, r);
                    break;
                case OPK_SPLIT: {  // rays.rs:1253-1264, ray.rs:752-772
                    const DevSplit& o = op.u.split;
                    if (in_range) {
                        if (alive) {
                            RayState s = r;
                            if (valid) { r.e *= o.ratio; s.e *= 1.0 - o.ratio; }
                            if (o.out.px != nullptr) store_ray<true>(o.out, i, s, valid ? RAY_VALID : RAY_INVALID);
                        } else if (o.out.px != nullptr) {
                            o.out.valid[i] = RAY_REMOVED;
                        }
                    }
                    break;
                }
                case OPK_SNAPSHOT: {
                    const DevSnapshot& o = op.u.snapshot;
                    if (in_range) {
                        if (alive) store_ray<true>(o.out, i, r, valid ? RAY_VALID : RAY_INVALID);
                        else o.out.valid[i] = RAY_REMOVED;
                    }
                    break;
                }
                case OPK_TRANSFORM:  // ray.rs:787-804
                    if (alive) {
                        r.pos = iso_point(op.u.transform.iso, r.pos);
                        r.dir = iso_vector(op.u.transform.iso, r.dir);
                    }
                    break;
                default: break;
            }
            if (op.epi_flags && alive) {  // folded threshold + limits (rays.rs:1145-1162, 1386-1404)
                const int f = op.epi_flags;
                if ((f & EPI_THRESHOLD) && r.e < op.epi_min_energy) valid = false;
                if (f & EPI_LIMITS) {
                    if ((uint64_t)r.bounces > op.epi_max_bounces) valid = false;
                    if ((f & EPI_CHECK_REFR) && (uint64_t)r.refr >= op.epi_max_refr) valid = false;
                }
            }
        }

        if (COMPACT) {
            unsigned long long slot = warp_reserve(alive, out_tail);
            if (alive) {
                if (slot < out_cap) store_ray<true>(out, slot, r, valid ? RAY_VALID : RAY_INVALID);
                else c.status |= 1u << OPM_ERR_CAPACITY;
            }
        } else if (in_range) {
            if (alive) {
                if (out.px == in.px) store_ray<false>(out, i, r, valid ? RAY_VALID : RAY_INVALID);
                else store_ray<true>(out, i, r, valid ? RAY_VALID : RAY_INVALID);
            } else if (vb != RAY_REMOVED || out.px != in.px) {
                out.valid[i] = RAY_REMOVED;
            }
        }
        // ---- tile counters: one warp reduction per register, lane 0 adds the fields to the CTA totals
        {
            const unsigned ra = __reduce_add_sync(0xffffffffu, c.a), rb = __reduce_add_sync(0xffffffffu, c.b);
            const unsigned rp = __reduce_add_sync(0xffffffffu, c.apodized);
            const unsigned live = __ballot_sync(0xffffffffu, alive), ok = __ballot_sync(0xffffffffu, valid && alive);
            const unsigned st = __reduce_or_sync(0xffffffffu, c.status);
            if (ok) {  // first valid ray of the warp = lowest id: ~key is largest there
                unsigned long long inv = (valid && alive) ? ~(((unsigned long long)r.id << 32) | r.bounces) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, inv, o); inv = t > inv ? t : inv; }
                if ((threadIdx.x & 31) == 0) atomicMax(&s_first_inv, inv);
            }
            if ((threadIdx.x & 31) == 0) {
                if (ra & 0xffffu) atomicAdd(&s_cnt[CNT_INTERACTIONS], (unsigned long long)(ra & 0xffffu));
                if (ra >> 16) atomicAdd(&s_cnt[CNT_HITS], (unsigned long long)(ra >> 16));
                if (rb & 0xffffu) atomicAdd(&s_cnt[CNT_MISSED], (unsigned long long)(rb & 0xffffu));
                if (rb >> 16) atomicAdd(&s_cnt[CNT_CHILDREN], (unsigned long long)(rb >> 16));
                if (rp) atomicAdd(&s_cnt[CNT_APODIZED], (unsigned long long)rp);
                if (live) atomicAdd(&s_cnt[CNT_ALIVE_OUT], (unsigned long long)__popc(live));
                if (ok) atomicAdd(&s_cnt[CNT_VALID_OUT], (unsigned long long)__popc(ok));
                if (st) atomicOr(&s_status, st);
            }
        }
    }

    // ---- CTA totals -> global statistics
    __syncthreads();
    if (threadIdx.x < CNT_N && s_cnt[threadIdx.x]) {
        unsigned long long* dst[CNT_N] = {&stats->interactions, &stats->children, &stats->hits, &stats->missed,
                                          &stats->apodized, &stats->valid_out, &stats->alive_out};
        atomicAdd(dst[threadIdx.x], s_cnt[threadIdx.x]);
    }
    if (threadIdx.x == 0 && s_status) atomicOr(&stats->status_bits, s_status);
    if (threadIdx.x == 0 && s_first_inv) atomicMax(&stats->first_inv, s_first_inv);
}

}  // namespace

int OPM_VARIANT(launch_trace)(const TraceLaunch& L) {
    size_t smem = (size_t)L.n_ops * sizeof(DevOp);
    cudaStream_t st = (cudaStream_t)L.stream;
    cudaError_t err;
    if (L.compact) {
        err = cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        trace_kernel<true><<<L.grid, L.block, smem, st>>>(L.ops, L.n_ops, L.in, L.out, L.n, L.out_cap, L.out_tail, L.stats);
    } else {
        err = cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        trace_kernel<false><<<L.grid, L.block, smem, st>>>(L.ops, L.n_ops, L.in, L.out, L.n, L.out_cap, L.out_tail, L.stats);
    }
    return (int)cudaGetLastError();
}

int OPM_VARIANT(trace_block)() { return OPM_TRACE_BLOCK; }

int OPM_VARIANT(trace_occupancy)(int block, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<false>, block, smem) != cudaSuccess) return 1;
    return nb < 1 ? 1 : nb;
}

}  // namespace opm
