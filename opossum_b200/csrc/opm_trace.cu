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

#define OPM_STAT_KEY2 1  // the interpreter runs any program: OPM_REFRACT_STAT_KEY2 included
#include "opm_trace_ops.cuh"

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

__device__ __forceinline__ void hist_push(const HistParams& hp, int level, uint64_t slot, const V3& p) {
    double* row = hp.base + (size_t)level * 3 * hp.stride + slot;
    row[0] = p.x; row[hp.stride] = p.y; row[2 * hp.stride] = p.z;
}

template <bool COMPACT, bool HIST, bool HELP = false>
__device__ __forceinline__ void trace_body(const DevOp* __restrict__ g_ops, int n_ops, const DevRays& in, const DevRays& out, uint64_t n,
                                           uint64_t out_cap, unsigned long long* out_tail, DevStats* stats, const HistParams& hp,
                                           const HelperParams& hlp = HelperParams()) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DevOp* ops = reinterpret_cast<DevOp*>(smem_raw);
    unsigned* stash = reinterpret_cast<unsigned*>(smem_raw + (size_t)n_ops * sizeof(DevOp));  // only sized when the program forks
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

    __shared__ __align__(16) cta_count_t s_cnt[kCntWords];
    __shared__ unsigned s_status;
    __shared__ unsigned long long s_first_inv;  // max of ~(id << 32 | bounces) over the valid rays this CTA wrote
    __shared__ unsigned long long s_first_inv2; // the same right after the op flagged OPM_REFRACT_STAT_KEY2
    cta_count_init(s_cnt);
    if (threadIdx.x == 0) { s_status = 0u; s_first_inv = 0ull; s_first_inv2 = 0ull; }
    __syncthreads();

    const unsigned tiles = (unsigned)((n + blockDim.x - 1) / blockDim.x);
    unsigned warp_best = 0xffffffffu;
    for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        RayCtx x;
        uint8_t vb;
        tile_load(in, (uint64_t)tile * blockDim.x + threadIdx.x, n, x, vb);
        tile_prefetch(in, (uint64_t)(tile + gridDim.x) * blockDim.x + threadIdx.x, n);
        HelperState hs;
        if (HELP && x.in_range && x.alive) helpers_load(hlp, x.i, hs);
        for (int k = 0; k < n_ops; ++k) {
            const DevOp& op = ops[k];
            switch (op.kind) {  // warp-uniform
                case OPK_REFRACT: {
                    const V3 p0 = x.r.pos;
                    const bool moved = HELP ? exec_refract_helpers(RefractRt(op.u.refract), x, hs) : exec_refract(RefractRt(op.u.refract), x);
                    if (HIST && moved && hp.level[k] >= 0) hist_push(hp, hp.level[k], x.i, p0);
                    break;
                }
                case OPK_APODIZE: {
                    const bool clipped = exec_apodize(ApertureRt(op.u.apodize.ap), op.u.apodize.inv, x);
                    if (HIST && clipped && hp.level[k] >= 0) hist_push(hp, hp.level[k], x.i, x.r.pos);
                    break;
                }
                case OPK_THRESHOLD: exec_threshold(op.u.threshold, x); break;
                case OPK_LIMITS: exec_limits(op.u.limits, x); break;
                case OPK_FILTER: exec_filter(op.u.filter, x); break;
                case OPK_PARAXIAL:
                    if (HELP) exec_paraxial_helpers(op.u.paraxial, x, hs);
                    else exec_paraxial(op.u.paraxial, x);
                    break;
                case OPK_SPLIT: exec_split(op.u.split, x); break;
                case OPK_SNAPSHOT: exec_snapshot(op.u.snapshot, x); break;
                case OPK_TRANSFORM: exec_transform(op.u.transform, x); break;
                case OPK_DIFFRACT: {
                    const V3 p0 = x.r.pos;
                    const bool moved = exec_diffract(RefractRt(op.u.diffract.surf), op.u.diffract, x);
                    if (HIST && moved && hp.level[k] >= 0) hist_push(hp, hp.level[k], x.i, p0);
                    break;
                }
                case OPK_FORK: exec_fork(op.u.fork, x, stash); break;
                case OPK_JOIN: exec_join(op.u.join, x, stash); break;
                default: break;
            }
            exec_epilogue(EpilogueRt(op.epi_flags), op, x);
        }
        tile_store<COMPACT>(in, out, out_cap, out_tail, x, vb);
        if (HELP && x.in_range && x.alive) helpers_store(hlp, x.i, hs);
        tile_count(x, s_cnt, &s_status, &s_first_inv, warp_best, &s_first_inv2);
    }
    __syncthreads();
    cta_flush(stats, s_cnt, s_status, s_first_inv, s_first_inv2);
}

template <bool COMPACT>
__global__ void __launch_bounds__(OPM_TRACE_BLOCK, OPM_TRACE_MIN_CTAS)
trace_kernel(const DevOp* __restrict__ g_ops, int n_ops, DevRays in, DevRays out, uint64_t n, uint64_t out_cap,
             unsigned long long* out_tail, DevStats* stats) {
    HistParams none;  // never read: HIST = false
    trace_body<COMPACT, false>(g_ops, n_ops, in, out, n, out_cap, out_tail, stats, none);
}
// the same program with the per-ray position log: a separate kernel, so that the common one carries no trace of it
__global__ void __launch_bounds__(OPM_TRACE_BLOCK, OPM_TRACE_MIN_CTAS)
trace_history_kernel(const DevOp* __restrict__ g_ops, int n_ops, DevRays in, DevRays out, uint64_t n, DevStats* stats,
                     const __grid_constant__ HistParams hp) {
    trace_body<false, true>(g_ops, n_ops, in, out, n, 0, nullptr, stats, hp);
}

// the same program for a bundle that carries fluence helper rays: a separate kernel as well
__global__ void __launch_bounds__(OPM_TRACE_BLOCK, 1)
trace_helpers_kernel(const DevOp* __restrict__ g_ops, int n_ops, DevRays in, DevRays out, uint64_t n, DevStats* stats,
                     const HelperParams hlp) {
    HistParams none;
    trace_body<false, false, true>(g_ops, n_ops, in, out, n, 0, nullptr, stats, none, hlp);
}

}  // namespace

int OPM_VARIANT(launch_trace)(const TraceLaunch& L) {
    size_t smem = (size_t)L.n_ops * sizeof(DevOp) + (L.has_fork ? (size_t)L.block * kStashWords * 4 : 0);
    cudaStream_t st = (cudaStream_t)L.stream;
    cudaError_t err;
    if (L.helpers) {
        if (L.compact || L.hist) return (int)cudaErrorInvalidValue;
        err = cudaFuncSetAttribute(trace_helpers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        trace_helpers_kernel<<<L.grid, L.block, smem, st>>>(L.ops, L.n_ops, L.in, L.out, L.n, L.stats, *L.helpers);
    } else if (L.hist) {
        if (L.compact) return (int)cudaErrorInvalidValue;
        err = cudaFuncSetAttribute(trace_history_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        trace_history_kernel<<<L.grid, L.block, smem, st>>>(L.ops, L.n_ops, L.in, L.out, L.n, L.stats, *L.hist);
    } else if (L.compact) {
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
