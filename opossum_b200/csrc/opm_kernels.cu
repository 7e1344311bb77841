// opm_kernels.cu -- bundle plumbing, detector and source kernels (sm_100a).
//
//   * AoS<->SoA staging for upload/download, defaults of Ray::new (ray.rs:108-140)
//   * Rays::transformed_by_iso
//   * spot-diagram statistics (rays.rs:521-701, nodes/spot_diagram.rs:101-163): two fused passes
//   * per-bounce-order tallies
//   * hit-map bounding box (rays_hit_map.rs:522-562) and the Binning fluence estimator
//     (rays_hit_map.rs:330-391): shared-memory privatised map when it fits, otherwise warp-run
//     aggregated global f64 atomics
//   * FluenceData peak / total energy (fluence_data.rs:45-68,130-141)
//   * on-device sources (position_distributions/{grid,fibonacci}.rs, energy_distributions/, rays.rs:264-297)
//   * order-preserving compaction of removed rays (ballot + popc ranks, block offsets from a scan)
// All kernels are HBM-bound streaming kernels: coalesced SoA accesses, persistent grids sized from the
// SM count, reductions through warp shuffles before any atomic.
#include <cuda_runtime.h>

#include <cfloat>

#include "opm_trace_ops.cuh"  // includes opm_device.cuh; the mbarrier / bulk-copy helpers
#include <cstdlib>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "opm_kernels.h"

namespace opm {

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// monotonic double <-> uint64 key so atomicMin/atomicMax work on doubles of any sign
__device__ __forceinline__ unsigned long long dkey(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// AoS <-> SoA
// ---------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const OpmRay* __restrict__ a, DevRays R, uint64_t off, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        OpmRay r = a[i];
        uint64_t j = off + i;
        R.px[j] = r.pos[0]; R.py[j] = r.pos[1]; R.pz[j] = r.pos[2];
        R.dx[j] = r.dir[0]; R.dy[j] = r.dir[1]; R.dz[j] = r.dir[2];
        R.e[j] = r.e; R.wvl[j] = r.wvl; R.opl[j] = r.opl; R.n[j] = r.n;
        R.bounces[j] = r.bounces; R.refr[j] = r.refr; R.id[j] = r.id;
        R.valid[j] = r.valid ? RAY_VALID : RAY_INVALID;
    }
}
__global__ void soa_to_aos_kernel(DevRays R, OpmRay* __restrict__ a, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        OpmRay r;
        r.pos[0] = R.px[i]; r.pos[1] = R.py[i]; r.pos[2] = R.pz[i];
        r.dir[0] = R.dx[i]; r.dir[1] = R.dy[i]; r.dir[2] = R.dz[i];
        r.e = R.e[i]; r.wvl = R.wvl[i]; r.opl = R.opl[i]; r.n = R.n[i];
        r.bounces = R.bounces[i]; r.refr = R.refr[i]; r.id = R.id[i];
        r.valid = R.valid[i];  // 2 = removed: dropped by the host side of download
        a[i] = r;
    }
}
// Ray::new defaults after a struct-of-arrays upload: normalised direction, opl 0, n 1, counters 0, valid
__global__ void soa_defaults_kernel(DevRays R, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        V3 d = v3(R.dx[i], R.dy[i], R.dz[i]);
        double nn = sqrt(dot(d, d));
        R.dx[i] = d.x / nn; R.dy[i] = d.y / nn; R.dz[i] = d.z / nn;
        R.opl[i] = 0.0; R.n[i] = 1.0; R.bounces[i] = 0; R.refr[i] = 0; R.id[i] = (uint32_t)i; R.valid[i] = RAY_VALID;
    }
}
__global__ void transform_kernel(DevRays R, DevIso I, uint64_t n) {  // ray.rs:787-804
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] == RAY_REMOVED) continue;
        V3 p = iso_point(I, v3(R.px[i], R.py[i], R.pz[i]));
        V3 d = iso_vector(I, v3(R.dx[i], R.dy[i], R.dz[i]));
        R.px[i] = p.x; R.py[i] = p.y; R.pz[i] = p.z;
        R.dx[i] = d.x; R.dy[i] = d.y; R.dz[i] = d.z;
    }
}

// ---------------------------------------------------------------------------------------------
// spot statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spot_pass1_kernel(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc) {
    double s[7] = {0, 0, 0, 0, 0, 0, 0};  // x y z ex ey ez e
    unsigned long long nv = 0, nt = 0, first = ~0ull;
    // four independent slots per thread and iteration: the loads of a batch are in flight together
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        uint8_t v[4];
        double px[4], py[4], pz[4], e[4];
        uint32_t id[4], bn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i = base + u * stride;
            const bool in = i < n;
            v[u] = in ? R.valid[i] : (uint8_t)RAY_REMOVED;
            const bool ld = in;  // not `valid`: the loads would wait for the flag (two round trips); most slots are valid anyway
            px[u] = ld ? __ldcs(R.px + i) : 0.0; py[u] = ld ? __ldcs(R.py + i) : 0.0; pz[u] = ld ? __ldcs(R.pz + i) : 0.0;
            e[u] = ld ? __ldcs(R.e + i) : 0.0; id[u] = ld ? __ldcs(R.id + i) : 0u; bn[u] = ld ? __ldcs(R.bounces + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u] == RAY_REMOVED) continue;
            nt++;
            if (v[u] != RAY_VALID) continue;
            nv++;
            V3 p = v3(px[u], py[u], pz[u]);
            if (use_frame) p = iso_point(inv, p);
            s[0] += p.x; s[1] += p.y; s[2] += p.z;
            s[3] += p.x * e[u]; s[4] += p.y * e[u]; s[5] += p.z * e[u]; s[6] += e[u];
            unsigned long long key = ((unsigned long long)id[u] << 32) | bn[u];
            first = key < first ? key : first;
        }
    }
    __shared__ double sh[8][7];
    __shared__ unsigned long long shn[8][3];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 7; ++j) s[j] = warp_sum(s[j]);
    nv = warp_sum_u64(nv); nt = warp_sum_u64(nt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, first, o); first = t < first ? t : first; }
    if (l == 0) { for (int j = 0; j < 7; ++j) sh[w][j] = s[j]; shn[w][0] = nv; shn[w][1] = nt; shn[w][2] = first; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[7] = {0, 0, 0, 0, 0, 0, 0};
        unsigned long long a = 0, b = 0, f = ~0ull;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
            for (int j = 0; j < 7; ++j) t[j] += sh[k][j];
            a += shn[k][0]; b += shn[k][1]; f = shn[k][2] < f ? shn[k][2] : f;
        }
        for (int j = 0; j < 7; ++j) atomicAdd(&acc->sum[j], t[j]);
        atomicAdd(&acc->n_valid, a); atomicAdd(&acc->n_total, b); atomicMin(&acc->first_key, f);
    }
}
__global__ void __launch_bounds__(256) spot_pass2_kernel(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc) {
    const double nvalid = (double)acc->n_valid;
    // rays.rs:599-612 / 618-637: centroid = sum / len ; energy centroid = sum(e*p) / sum(e)
    const double cx = acc->sum[0] / nvalid, cy = acc->sum[1] / nvalid;
    const double ex = acc->sum[3] / acc->sum[6], ey = acc->sum[4] / acc->sum[6];
    const double cxm = cx * 1.0E3, cym = cy * 1.0E3;  // get::<millimeter>()
    double mx = 0.0, s2 = 0.0, se = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        bool ok[4];
        double px[4], py[4], pz[4], e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i = base + u * stride;
            const bool in = i < n;  // loads do not wait for the flag
            ok[u] = in && R.valid[i] == RAY_VALID;
            px[u] = in ? __ldcs(R.px + i) : 0.0; py[u] = in ? __ldcs(R.py + i) : 0.0; pz[u] = in ? __ldcs(R.pz + i) : 0.0;
            e[u] = in ? __ldcs(R.e + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            V3 p = v3(px[u], py[u], pz[u]);
            if (use_frame) p = iso_point(inv, p);
            double dx = cxm - p.x * 1.0E3, dy = cym - p.y * 1.0E3;
            double d2 = dx * dx + dy * dy;
            mx = fmax(mx, d2);     // max of sqrt == sqrt of max (rays.rs:642-660)
            s2 += d2;              // rays.rs:666-684: distance(..).powi(2)
            double ux = ex - p.x, uy = ey - p.y;
            se += (ux * ux + uy * uy) * e[u];  // rays.rs:690-701
        }
    }
    __shared__ double sh[8][3];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    mx = warp_max(mx); s2 = warp_sum(s2); se = warp_sum(se);
    if (l == 0) { sh[w][0] = mx; sh[w][1] = s2; sh[w][2] = se; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a = fmax(a, sh[k][0]); b += sh[k][1]; c += sh[k][2]; }
        atomicMax(&acc->max_d2_key, dkey(a));
        atomicAdd(&acc->sum_d2, b); atomicAdd(&acc->sum_ed2, c);
    }
}
__global__ void __launch_bounds__(256) bounce_lvl_kernel(DevRays R, uint64_t n, unsigned long long* first_key) {
    unsigned long long first = ~0ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        unsigned long long key = ((unsigned long long)R.id[i] << 32) | R.bounces[i];
        first = key < first ? key : first;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, first, o); first = t < first ? t : first; }
    if ((threadIdx.x & 31) == 0 && first != ~0ull) atomicMin(first_key, first);
}
__global__ void __launch_bounds__(256) tally_kernel(DevRays R, uint64_t n, uint32_t n_orders, unsigned long long* counts, double* energy) {
    extern __shared__ unsigned char tsm[];
    double* se = reinterpret_cast<double*>(tsm);
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(se + n_orders);
    for (uint32_t k = threadIdx.x; k < n_orders; k += blockDim.x) { se[k] = 0.0; sc[k] = 0; }
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        uint32_t b = R.bounces[i];
        if (b >= n_orders) b = n_orders - 1;
        atomicAdd(&se[b], R.e[i]);
        atomicAdd(&sc[b], 1ull);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n_orders; k += blockDim.x)
        if (sc[k]) { atomicAdd(&energy[k], se[k]); atomicAdd(&counts[k], sc[k]); }
}

// ---------------------------------------------------------------------------------------------
// hit map: bounding box and Binning
// ---------------------------------------------------------------------------------------------
// bounce filter of the hit-map kernels: >= 0 a bounce count, -1 none, < -1 the NEGATED device address of a launch statistic
// `first_inv` (DevStats): the bounce level of the bundle the context's latest trace launch ran on, read here on the device, so
// that the ghost-focus analysis chains trace -> per-bundle fluence check without a host round trip (Rays::bounce_lvl,
// rays.rs:183-194: the bounce count of the first valid ray, 0 when there is none)
__device__ __forceinline__ long long resolve_bounce_filter(long long f) {
    if (f >= -1) return f;
    const unsigned long long inv = *reinterpret_cast<const unsigned long long*>(-f);
    return inv ? (long long)(~inv & 0xffffffffull) : 0ll;
}
__device__ __forceinline__ void hits_bbox_body(const DevHits& H, uint64_t n, long long bounce_filter, BBoxAccum* acc) {
    double xmin = DBL_MAX, xmax = -DBL_MAX, ymin = DBL_MAX, ymax = -DBL_MAX;
    unsigned long long cnt = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        double e[4], x[4], y[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four independent slots: their loads are in flight together
            const uint64_t i = base + u * stride;
            const bool in = i < n;
            e[u] = in ? __ldcs(H.e + i) : -1.0;
            // x and y do not wait for e (one round trip instead of two; empty slots cost their bytes)
            x[u] = in ? __ldcs(H.x + i) : 0.0; y[u] = in ? __ldcs(H.y + i) : 0.0;
            ok[u] = e[u] >= 0.0;
            if (ok[u] && bounce_filter >= 0) ok[u] = (long long)H.bounce[i] == bounce_filter;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            xmin = fmin(xmin, x[u]); xmax = fmax(xmax, x[u]); ymin = fmin(ymin, y[u]); ymax = fmax(ymax, y[u]);
            cnt++;
        }
    }
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax); cnt = warp_sum_u64(cnt);
    // one set of atomics per CTA, not per warp: ten thousand atomics on the same five addresses are a serial tail in the L2
    // (66 -> 46 us for 10^7 slots)
    __shared__ double sh[8][4];
    __shared__ unsigned long long shc[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[w][0] = xmin; sh[w][1] = xmax; sh[w][2] = ymin; sh[w][3] = ymax; shc[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
            if (!shc[k]) continue;
            c += shc[k];
            xmin = fmin(xmin, sh[k][0]); xmax = fmax(xmax, sh[k][1]); ymin = fmin(ymin, sh[k][2]); ymax = fmax(ymax, sh[k][3]);
        }
        if (c) {
            atomicMin(&acc->key[0], dkey(xmin)); atomicMax(&acc->key[1], dkey(xmax));
            atomicMax(&acc->key[2], dkey(ymax)); atomicMin(&acc->key[3], dkey(ymin));
            atomicAdd(&acc->count, c);
        }
    }
}
__global__ void __launch_bounds__(256) hits_bbox_kernel(DevHits H, uint64_t n, long long bounce_filter, BBoxAccum* acc) {
    hits_bbox_body(H, n, resolve_bounce_filter(bounce_filter), acc);
}

// Rust `f64 as usize` (saturating, NaN -> 0)
__device__ __forceinline__ unsigned long long f64_to_usize(double v) {
    if (!(v > 0.0)) return 0ull;
    if (v >= 18446744073709551615.0) return ~0ull;
    return (unsigned long long)v;
}

struct BinTaps { unsigned long long base; double w00, w10, w01, w11; bool ok, has_x, has_y; };

__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// Map ranges travel on the device as (-left, right, top, -bottom): one element-wise MAX merges the ranges of several
// GPUs, and -inf is the identity of a GPU without hits.
__global__ void bbox_finalize_kernel(const BBoxAccum* acc, double* range) {
    if (acc->count == 0) { range[0] = range[1] = range[2] = range[3] = -INFINITY; return; }
    range[0] = -dkey_inv(acc->key[0]); range[1] = dkey_inv(acc->key[1]);
    range[2] = dkey_inv(acc->key[2]); range[3] = -dkey_inv(acc->key[3]);
}
// rays_hit_map.rs:343-350 from a device-resident range (same operations as the host formula)
__device__ __forceinline__ bool bin_params_from_range(unsigned long long nx, unsigned long long ny, const double* range, BinParams& P) {
    double l = -range[0], r = range[1], t = range[2], b = -range[3];
    P.nx = nx; P.ny = ny; P.left = l; P.bottom = b;
    double bin_width = (r - l) / (double)nx;
    double bin_height = (t - b) / (double)ny;
    P.bin_area = bin_width * bin_height;
    P.width_step = (r - l) / (double)(nx - 1);
    P.height_step = (t - b) / (double)(ny - 1);
    return isfinite(l) && isfinite(r) && isfinite(t) && isfinite(b);
}

// rays_hit_map.rs:352-376 for one hit
__device__ __forceinline__ BinTaps bin_taps(const BinParams& P, double x, double y, double e, unsigned* status) {
    BinTaps t;
    double ipx, ipy;
    double fx = modf((x - P.left) / P.width_step, &ipx);
    double fy = modf((y - P.bottom) / P.height_step, &ipy);
    unsigned long long xi = f64_to_usize(ipx), yi = f64_to_usize(ipy);
    double fl = e / P.bin_area;
    t.w00 = (1.0 - fx) * (1.0 - fy) * fl;
    t.w10 = fx * (1.0 - fy) * fl;
    t.w01 = (1.0 - fx) * fy * fl;
    t.w11 = fx * fy * fl;
    t.ok = (xi < P.nx) && (yi < P.ny);
    if (!t.ok) *status |= 1u << OPM_ERR_BIN_RANGE;  // the reference panics (index out of bounds)
    t.has_x = xi < P.nx - 1;
    t.has_y = yi < P.ny - 1;
    t.base = xi * P.ny + yi;  // column-major ny x nx
    return t;
}

// x / d from the correctly rounded reciprocal of a loop-invariant d: one multiplication and one residual correction
// (Markstein); equal to the IEEE quotient except in rare last-bit cases, where it is one ulp off
__device__ __forceinline__ double div_by_invariant(double x, double d, double inv) {
    const double q = x * inv;
    return fma(fma(-q, d, x), inv, q);
}
struct BinInv { double width_step, height_step, bin_area; };
// bin_taps for the shared-memory kernel (the map has fewer than 2^31 bins): the three divisions by loop-invariant numbers go
// through div_by_invariant; modf(q) = (q - trunc(q), trunc(q)) for finite q (one FRND instead of the library routine); Rust's
// saturating `f64 as usize` is cvt.rzi.u64.f64 behind a NaN / sign guard; indices in 32 bits
struct BinTaps32 { unsigned base; double w00, w10, w01, w11; bool ok, has_x, has_y; };
__device__ __forceinline__ BinTaps32 bin_taps_inv(const BinParams& P, const BinInv& I, double x, double y, double e, unsigned* status) {
    BinTaps32 t;
    const double qx = div_by_invariant(x - P.left, P.width_step, I.width_step);
    const double qy = div_by_invariant(y - P.bottom, P.height_step, I.height_step);
    const double ipx = trunc(qx), ipy = trunc(qy);
    const double fx = qx - ipx, fy = qy - ipy;
    // Rust's `f64 as usize`: NaN and negatives -> 0, too large -> MAX.  The conversion instruction saturates the same way but turns
    // NaN into 2^63: a NaN index (a hit on a range of zero width: 0 / 0) must land in bin 0 like the reference's, not raise
    // OPM_ERR_BIN_RANGE
    const unsigned long long xi = ipx > 0.0 ? __double2ull_rz(ipx) : 0ull, yi = ipy > 0.0 ? __double2ull_rz(ipy) : 0ull;
    const double fl = div_by_invariant(e, P.bin_area, I.bin_area);
    const double gx = 1.0 - fx, gy = 1.0 - fy;
    t.w00 = gx * gy * fl;
    t.w10 = fx * gy * fl;
    t.w01 = gx * fy * fl;
    t.w11 = fx * fy * fl;
    t.ok = (xi < P.nx) && (yi < P.ny);
    if (!t.ok) *status |= 1u << OPM_ERR_BIN_RANGE;
    const unsigned nx = (unsigned)P.nx, ny = (unsigned)P.ny, ux = (unsigned)xi, uy = (unsigned)yi;
    t.has_x = ux + 1u < nx;
    t.has_y = uy + 1u < ny;
    t.base = ux * ny + uy;  // column-major ny x nx
    return t;
}

// map fits in shared memory: privatise the whole map per CTA, flush non-zero bins with global atomics
__global__ void __launch_bounds__(1024) binning_smem_kernel(DevHits H, uint64_t n, long long bounce_filter, unsigned long long nx,
                                                            unsigned long long ny, const double* __restrict__ range,
                                                            double* __restrict__ map, unsigned* status_out) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    extern __shared__ double smap[];
    BinParams P;
    if (!bin_params_from_range(nx, ny, range, P)) return;  // no hit on any GPU: nothing to bin
    const unsigned long long bins = P.nx * P.ny;
    for (unsigned long long k = threadIdx.x; k < bins; k += blockDim.x) smap[k] = 0.0;
    __syncthreads();
    unsigned status = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double e = __ldcs(H.e + i);
        if (!(e >= 0.0)) continue;
        if (bounce_filter >= 0 && (long long)H.bounce[i] != bounce_filter) continue;
        BinTaps t = bin_taps(P, __ldcs(H.x + i), __ldcs(H.y + i), e, &status);
        if (!t.ok) continue;
        atomicAdd(&smap[t.base], t.w00);
        if (t.has_x) {
            atomicAdd(&smap[t.base + P.ny], t.w10);
            if (t.has_y) atomicAdd(&smap[t.base + P.ny + 1], t.w11);
        }
        if (t.has_y) atomicAdd(&smap[t.base + 1], t.w01);
    }
    __syncthreads();
    for (unsigned long long k = threadIdx.x; k < bins; k += blockDim.x) {
        double v = smap[k];
        if (v != 0.0) atomicAdd(&map[k], v);
    }
    if (status) atomicOr(status_out, status);
}

// The same splat with the CTA-private map held as 64-bit FIXED-POINT sums split over two 32-bit words per bin. Shared-memory
// f64 (and u64) adds are compare-and-swap loops on sm_100a (ATOMS.CAST.SPIN.64); 32-bit integer adds are native (ATOMS.ADD), so a
// tap costs two fire-and-forget adds: the low word (its returned old value tells whether the add wrapped) and the high word plus
// that carry. The unit is a power of two chosen per CTA from the largest energy among the CTA's slots so that no bin can exceed
// 2^62 units: a tap is rounded to 2^-62 x (slots per CTA) of the largest possible tap (<= 2^-45 relative to it for 10^5 slots per
// CTA), and the CTA's sums do not depend on the order of the adds. Taps, indices and the flush into the global f64 map are
// those of the kernel above.
template <int BLOCK, int UNROLL, bool FASTDIV>
__device__ __forceinline__ void binning_smem_fixed_body(const DevHits& H, uint64_t n, long long bounce_filter, unsigned long long nx,
                                                        unsigned long long ny, const double* __restrict__ range, double* __restrict__ map,
                                                        unsigned* status_out) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    extern __shared__ unsigned sfix[];
    __shared__ unsigned long long s_emax_key;
    BinParams P;
    if (!bin_params_from_range(nx, ny, range, P)) return;
    const unsigned long long bins = P.nx * P.ny;
    unsigned* lo = sfix;
    unsigned* hi = sfix + bins;
    for (unsigned long long k = threadIdx.x; k < 2 * bins; k += BLOCK) sfix[k] = 0u;
    if (threadIdx.x == 0) s_emax_key = 0ull;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * BLOCK;
    const uint64_t first = blockIdx.x * (uint64_t)BLOCK + threadIdx.x;
    const uint64_t cta_first = blockIdx.x * (uint64_t)BLOCK;
    const uint64_t cta_slots = cta_first < n ? (n - cta_first + stride - 1) / stride * BLOCK : 0;  // upper bound, CTA-uniform
    double emax = 0.0;
    for (uint64_t i = first; i < n; i += stride) emax = fmax(emax, H.e[i]);  // NaN / negative (= no hit) lose against 0
    emax = warp_max(emax);
    if ((threadIdx.x & 31) == 0 && emax > 0.0) atomicMax(&s_emax_key, (unsigned long long)__double_as_longlong(emax));  // >= 0: bits are ordered
    __syncthreads();
    const double flmax = __longlong_as_double((long long)s_emax_key) / P.bin_area;
    // every bin of this CTA stays below flmax * (slots of the CTA) < 2^ex; one unit = 2^(ex-62)
    int ex = 0;
    (void)frexp(flmax * (double)cta_slots, &ex);
    ex = max(-900, min(ex, 1000));
    const double inv_unit = ldexp(1.0, 62 - ex), unit = ldexp(1.0, ex - 62);
    unsigned status = 0;
    const BinInv I{1.0 / P.width_step, 1.0 / P.height_step, 1.0 / P.bin_area};
    const bool scaled = flmax > 0.0 && isfinite(flmax);
    {
        auto add = [&](unsigned b, double w) {
            if (w == 0.0) return;
            // NaN / infinite taps (a NaN hit position, an infinite energy) go to the global map as they are, like the f64 kernel's
            if (!scaled || !isfinite(w)) { atomicAdd(&map[b], w); return; }
            const unsigned long long q = __double2ull_rn(w * inv_unit);
            // Taps below 2^32 units keep fewer than 32 significant bits in fixed point: they (and negative taps -- a hit left of or
            // below an explicit range, which the reference adds as they are -- whose conversion saturates to 0) go to the global f64
            // map directly.  Every fixed-point tap is therefore rounded by at most 2^-33 of itself, so a bin -- a sum of
            // non-negative taps -- is within 1.2e-10 of the f64 sum RELATIVE TO ITS OWN VALUE, however small it is against the
            // map's peak.  Such taps are rare (~1e-3 of all taps for a beam with a 100:1 energy range).
            if (q < (1ull << 32)) { atomicAdd(&map[b], w); return; }
            const unsigned l = (unsigned)q;
            unsigned h = (unsigned)(q >> 32);
            if (l) { const unsigned old = atomicAdd(&lo[b], l); h += (unsigned)(old + l < old); }
            if (h) atomicAdd(&hi[b], h);
        };
        auto splat = [&](uint64_t i, double x, double y, double e) {
            if (!(e >= 0.0)) return;
            if (bounce_filter >= 0 && (long long)H.bounce[i] != bounce_filter) return;
            if (FASTDIV) {
                const BinTaps32 t = bin_taps_inv(P, I, x, y, e, &status);
                if (!t.ok) return;
                const unsigned ny32 = (unsigned)P.ny;
                add(t.base, t.w00);
                if (t.has_x) {
                    add(t.base + ny32, t.w10);
                    if (t.has_y) add(t.base + ny32 + 1u, t.w11);
                }
                if (t.has_y) add(t.base + 1u, t.w01);
                return;
            }
            BinTaps t = bin_taps(P, x, y, e, &status);
            if (!t.ok) return;
            add(t.base, t.w00);
            if (t.has_x) {
                add(t.base + P.ny, t.w10);
                if (t.has_y) add(t.base + P.ny + 1, t.w11);
            }
            if (t.has_y) add(t.base + 1, t.w01);
        };
        // the loads of UNROLL slots are issued before the first splat (atomics order memory operations: the compiler does not
        // move the next slot's loads above them)
        uint64_t i = first;
        for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
            double x[UNROLL], y[UNROLL], e[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) { e[u] = H.e[i + u * stride]; x[u] = __ldcs(H.x + i + u * stride); y[u] = __ldcs(H.y + i + u * stride); }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) splat(i + u * stride, x[u], y[u], e[u]);
        }
        for (; i < n; i += stride) splat(i, __ldcs(H.x + i), __ldcs(H.y + i), H.e[i]);
    }
    __syncthreads();
    for (unsigned long long k = threadIdx.x; k < bins; k += BLOCK) {
        const unsigned long long q = ((unsigned long long)hi[k] << 32) | lo[k];
        if (q) atomicAdd(&map[k], (double)q * unit);
    }
    if (status) atomicOr(status_out, status);
}
template <int BLOCK, int MIN_CTAS, int UNROLL, bool FASTDIV>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) binning_smem_fixed_kernel(DevHits H, uint64_t n, long long bounce_filter,
                                                                             unsigned long long nx, unsigned long long ny,
                                                                             const double* __restrict__ range, double* __restrict__ map,
                                                                             unsigned* status_out) {
    binning_smem_fixed_body<BLOCK, UNROLL, FASTDIV>(H, n, bounce_filter, nx, ny, range, map, status_out);
}

// large maps: consecutive lanes that fall into the same bin (a run) are summed with a segmented warp
// scan first, the run head issues the global f64 atomics
__global__ void __launch_bounds__(256) binning_global_kernel(DevHits H, uint64_t n, long long bounce_filter, unsigned long long nx,
                                                             unsigned long long ny, const double* __restrict__ range,
                                                             double* __restrict__ map, unsigned* status_out) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    BinParams P;
    if (!bin_params_from_range(nx, ny, range, P)) return;
    unsigned status = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) / 32 * 32;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
        BinTaps t;
        t.ok = false; t.base = ~0ull; t.w00 = t.w10 = t.w01 = t.w11 = 0.0; t.has_x = t.has_y = false;
        if (i < n) {
            double e = __ldcs(H.e + i);
            if (e >= 0.0 && !(bounce_filter >= 0 && (long long)H.bounce[i] != bounce_filter))
                t = bin_taps(P, __ldcs(H.x + i), __ldcs(H.y + i), e, &status);
        }
        const int lane = threadIdx.x & 31;
        // runs of consecutive lanes that hit the same base bin: run id = number of run heads at or below this lane
        const unsigned long long key = t.ok ? t.base : (~0ull - (unsigned long long)lane);  // invalid lanes never join a run
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev != key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int rid = __popc(heads & (0xffffffffu >> (31 - lane)));
        // segmented suffix sum (Hillis-Steele): run ids are monotonic, so equal ids at distance o mean one run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int r2 = __shfl_down_sync(0xffffffffu, rid, o);
            double a = __shfl_down_sync(0xffffffffu, t.w00, o), b = __shfl_down_sync(0xffffffffu, t.w10, o);
            double c = __shfl_down_sync(0xffffffffu, t.w01, o), d = __shfl_down_sync(0xffffffffu, t.w11, o);
            if (lane + o < 32 && r2 == rid) { t.w00 += a; t.w10 += b; t.w01 += c; t.w11 += d; }
        }
        if (t.ok && head) {
            atomicAdd(&map[t.base], t.w00);
            if (t.has_x) {
                atomicAdd(&map[t.base + P.ny], t.w10);
                if (t.has_y) atomicAdd(&map[t.base + P.ny + 1], t.w11);
            }
            if (t.has_y) atomicAdd(&map[t.base + 1], t.w01);
        }
    }
    if (status) atomicOr(status_out, status);
}

// Large maps, second form: every THREAD walks RUN consecutive slots and keeps the four taps of the bin it is in in registers.
// Slots are in source order, and the sources that feed a large map are ordered (grid: y runs fastest), so consecutive hits fall
// into the same bin or the one above / below it.  A step to the neighbouring row keeps the two taps the bins share
// ((x, y+1) and (x+1, y+1) of the old bin are (x, y) and (x+1, y) of the new one) and flushes the other two; anything else flushes
// all four.  A 10^8-hit C4 map (2.4 hits per bin along y) goes from 4 global f64 reductions per hit to about one, with no
// shuffles.  Taps the reference does not add (x + 1 == nx, y + 1 == ny) are held as exact zeros and zeros are never flushed,
// so the index arithmetic needs no edge cases.  Hits without locality cost what the un-merged splat costs.
// A CTA owns a contiguous range of slots (not a grid-stride set): concurrently running CTAs then work on different parts of the
// map and the reductions of one bin are not spread over a thousand CTAs at the same moment (the L2 serialises per address).
template <int RUN>
struct RunLoad {
    double x[RUN], y[RUN], e[RUN];
};
template <int RUN>
__device__ __forceinline__ void run_load(const DevHits& H, uint64_t s0, uint64_t n, RunLoad<RUN>& L) {
    if (s0 + RUN <= n) {  // the arrays are 256-byte aligned and s0 is a multiple of RUN: 16-byte loads
#pragma unroll
        for (int u = 0; u < RUN; u += 2) {
            // streaming (evict-first) loads: the hit records pass through the L2 once, the map's lines should stay
            const double2 ev = __ldcs(reinterpret_cast<const double2*>(H.e + s0 + u));
            const double2 xv = __ldcs(reinterpret_cast<const double2*>(H.x + s0 + u));
            const double2 yv = __ldcs(reinterpret_cast<const double2*>(H.y + s0 + u));
            L.e[u] = ev.x; L.e[u + 1] = ev.y; L.x[u] = xv.x; L.x[u + 1] = xv.y; L.y[u] = yv.x; L.y[u + 1] = yv.y;
        }
    } else {
#pragma unroll
        for (int u = 0; u < RUN; ++u) {
            const bool in = s0 + u < n;
            L.e[u] = in ? H.e[s0 + u] : -1.0; L.x[u] = in ? H.x[s0 + u] : 0.0; L.y[u] = in ? H.y[s0 + u] : 0.0;
        }
    }
}
// PREFETCH: the loads of the CTA's next chunk are issued before the current one is splatted (a second register set instead of more
// resident warps)
template <int BLOCK, int MIN_CTAS, int RUN, bool PREFETCH = false>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) binning_runs_kernel(DevHits H, uint64_t n, long long bounce_filter, unsigned long long nx,
                                                                       unsigned long long ny, const double* __restrict__ range,
                                                                       double* __restrict__ map, unsigned* status_out, int interleave) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    static_assert(RUN % 2 == 0, "slots are loaded in pairs");
    BinParams P;
    if (!bin_params_from_range(nx, ny, range, P)) return;
    const BinInv I{1.0 / P.width_step, 1.0 / P.height_step, 1.0 / P.bin_area};
    const unsigned ny32 = (unsigned)P.ny;
    unsigned status = 0;
    const uint64_t chunk = (uint64_t)BLOCK * RUN;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const uint64_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    uint64_t c0, c1, cstep;
    if (interleave) { c0 = blockIdx.x; c1 = n_chunks; cstep = gridDim.x; }
    else { c0 = blockIdx.x * per; c1 = c0 + per < n_chunks ? c0 + per : n_chunks; cstep = 1; }
    RunLoad<RUN> N;
    if (PREFETCH && c0 < c1) {
        const uint64_t s = c0 * chunk + (uint64_t)threadIdx.x * RUN;
        if (s < n) run_load(H, s, n, N);
    }
    for (uint64_t c = c0; c < c1; c += cstep) {
        const uint64_t s0 = c * chunk + (uint64_t)threadIdx.x * RUN;
        RunLoad<RUN> L;
        if (PREFETCH) {
            L = N;
            const uint64_t s1 = (c + cstep) * chunk + (uint64_t)threadIdx.x * RUN;
            if (c + cstep < c1 && s1 < n) run_load(H, s1, n, N);
            if (s0 >= n) continue;
        } else {
            if (s0 >= n) continue;
            run_load(H, s0, n, L);
        }
        bool open = false;
        unsigned base = 0;
        double a00 = 0.0, a10 = 0.0, a01 = 0.0, a11 = 0.0;
        auto red = [&](unsigned b, double v) { if (v != 0.0) atomicAdd(&map[b], v); };
#pragma unroll
        for (int u = 0; u < RUN; ++u) {
            if (!(L.e[u] >= 0.0)) continue;
            if (bounce_filter >= 0 && (long long)H.bounce[s0 + u] != bounce_filter) continue;
            BinTaps32 t = bin_taps_inv(P, I, L.x[u], L.y[u], L.e[u], &status);
            if (!t.ok) continue;
            if (!t.has_x) { t.w10 = 0.0; t.w11 = 0.0; }
            if (!t.has_y) { t.w01 = 0.0; t.w11 = 0.0; }
            if (open && t.base == base) { a00 += t.w00; a10 += t.w10; a01 += t.w01; a11 += t.w11; continue; }
            if (open && t.base == base + 1u) {         // one row up (or the first bin of the next column: then a01 = a11 = 0)
                red(base, a00); red(base + ny32, a10);
                a00 = a01 + t.w00; a10 = a11 + t.w10; a01 = t.w01; a11 = t.w11;
            } else if (open && t.base + 1u == base) {  // one row down
                red(base + 1u, a01); red(base + ny32 + 1u, a11);
                a01 = a00 + t.w01; a11 = a10 + t.w11; a00 = t.w00; a10 = t.w10;
            } else {
                if (open) { red(base, a00); red(base + ny32, a10); red(base + 1u, a01); red(base + ny32 + 1u, a11); }
                a00 = t.w00; a10 = t.w10; a01 = t.w01; a11 = t.w11;
                open = true;
            }
            base = t.base;
        }
        if (open) { red(base, a00); red(base + ny32, a10); red(base + 1u, a01); red(base + ny32 + 1u, a11); }
    }
    if (status) atomicOr(status_out, status);
}

// The same splat with the hit records staged through shared memory by TMA bulk copies (cp.async.bulk, one elected thread):
// STAGES tiles of BLOCK * RUN slots (x | y | e) per CTA, one mbarrier per stage, so the HBM reads of the next tiles are in flight
// while the warps splat the current one -- the direct-load form above waits on its loads (ncu: 9.4 warps per issue stalled on
// the long scoreboard at half occupancy).  A stage is re-armed after the block barrier that follows its last read.
template <int BLOCK, int MIN_CTAS, int RUN, int STAGES>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) binning_runs_tma_kernel(DevHits H, uint64_t n, long long bounce_filter, unsigned long long nx,
                                                                           unsigned long long ny, const double* __restrict__ range,
                                                                           double* __restrict__ map, unsigned* status_out) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    extern __shared__ __align__(128) unsigned char tsm[];
    __shared__ unsigned long long bar[STAGES];
    BinParams P;
    if (!bin_params_from_range(nx, ny, range, P)) return;
    const BinInv I{1.0 / P.width_step, 1.0 / P.height_step, 1.0 / P.bin_area};
    const unsigned ny32 = (unsigned)P.ny;
    unsigned status = 0;
    constexpr unsigned kTile = BLOCK * RUN;            // slots per tile
    constexpr unsigned kArr = kTile * 8;               // bytes of one field of a tile
    constexpr unsigned kStage = 3 * kArr;
    const uint64_t n_full = n / kTile;                 // tiles [0, n_full) are complete: they travel by bulk copy
    const uint64_t n_tiles = (n + kTile - 1) / kTile;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t tile, int s) {  // thread 0
        unsigned char* st = tsm + (size_t)s * kStage;
        mbar_expect_tx(&bar[s], kStage);
        bulk_g2s(st, H.x + tile * kTile, kArr, &bar[s]);
        bulk_g2s(st + kArr, H.y + tile * kTile, kArr, &bar[s]);
        bulk_g2s(st + 2 * kArr, H.e + tile * kTile, kArr, &bar[s]);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES - 1; ++s) {
            const uint64_t t = blockIdx.x + (uint64_t)s * gridDim.x;
            if (t < n_full) issue(t, s);
        }
    unsigned k = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
        const uint64_t s0 = tile * kTile + (uint64_t)threadIdx.x * RUN;
        RunLoad<RUN> L;
        if (tile < n_full) {
            const int s = k % STAGES;
            mbar_wait(&bar[s], (k / STAGES) & 1u);
            const double2* sx = reinterpret_cast<const double2*>(tsm + (size_t)s * kStage) + (size_t)threadIdx.x * (RUN / 2);
#pragma unroll
            for (int u = 0; u < RUN; u += 2) {
                const double2 xv = sx[u / 2], yv = sx[kTile / 2 + u / 2], ev = sx[kTile + u / 2];
                L.x[u] = xv.x; L.x[u + 1] = xv.y; L.y[u] = yv.x; L.y[u + 1] = yv.y; L.e[u] = ev.x; L.e[u + 1] = ev.y;
            }
        } else {
            if (s0 < n) run_load(H, s0, n, L);
        }
        __syncthreads();  // every thread holds its slots in registers: the stage of the PREVIOUS iteration's successor can be re-armed
        if (threadIdx.x == 0) {
            const uint64_t nt = tile + (uint64_t)(STAGES - 1) * gridDim.x;
            if (nt < n_full) issue(nt, (k + STAGES - 1) % STAGES);
        }
        if (s0 >= n) continue;
        bool open = false;
        unsigned base = 0;
        double a00 = 0.0, a10 = 0.0, a01 = 0.0, a11 = 0.0;
        auto red = [&](unsigned b, double v) { if (v != 0.0) atomicAdd(&map[b], v); };
#pragma unroll
        for (int u = 0; u < RUN; ++u) {
            if (s0 + u >= n || !(L.e[u] >= 0.0)) continue;
            if (bounce_filter >= 0 && (long long)H.bounce[s0 + u] != bounce_filter) continue;
            BinTaps32 t = bin_taps_inv(P, I, L.x[u], L.y[u], L.e[u], &status);
            if (!t.ok) continue;
            if (!t.has_x) { t.w10 = 0.0; t.w11 = 0.0; }
            if (!t.has_y) { t.w01 = 0.0; t.w11 = 0.0; }
            if (open && t.base == base) { a00 += t.w00; a10 += t.w10; a01 += t.w01; a11 += t.w11; continue; }
            if (open && t.base == base + 1u) {
                red(base, a00); red(base + ny32, a10);
                a00 = a01 + t.w00; a10 = a11 + t.w10; a01 = t.w01; a11 = t.w11;
            } else if (open && t.base + 1u == base) {
                red(base + 1u, a01); red(base + ny32 + 1u, a11);
                a01 = a00 + t.w01; a11 = a10 + t.w11; a00 = t.w00; a10 = t.w10;
            } else {
                if (open) { red(base, a00); red(base + ny32, a10); red(base + 1u, a01); red(base + ny32 + 1u, a11); }
                a00 = t.w00; a10 = t.w10; a01 = t.w01; a11 = t.w11;
                open = true;
            }
            base = t.base;
        }
        if (open) { red(base, a00); red(base + ny32, a10); red(base + 1u, a01); red(base + ny32 + 1u, a11); }
    }
    if (status) atomicOr(status_out, status);
}

// fluence_data.rs:45-68 (peak over finite bins) and :130-141 (sum over non-NaN bins)
__global__ void __launch_bounds__(256) peak_total_kernel(const double* __restrict__ map, uint64_t nx, uint64_t ny,
                                                         const double* __restrict__ range, PeakAccum* acc) {
    const uint64_t bins = nx * ny;
    // fluence_data.rs:131-133, x_range = left..right, y_range = bottom..top
    const double dx = (range[1] - (-range[0])) / (double)nx, dy = (range[2] - (-range[3])) / (double)ny;
    const double area = dx * dy;
    double peak = -INFINITY, tot = 0.0;
    // four independent 16-byte loads per thread and iteration (a 4096 x 4096 map is 128 MiB: this pass is pure HBM streaming)
    const uint64_t head = (reinterpret_cast<uintptr_t>(map) & 8u) ? 1 : 0;  // maps packed behind one another may start at 8 mod 16
    const uint64_t pairs = (bins - (bins ? head : 0)) / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    const double2* __restrict__ m2 = reinterpret_cast<const double2*>(map + head);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < pairs; i += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i + u * stride < pairs ? __ldcs(m2 + i + u * stride) : make_double2(NAN, NAN);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (isfinite(v[u].x)) peak = fmax(peak, v[u].x);
            if (!isnan(v[u].x)) tot += area * v[u].x;
            if (isfinite(v[u].y)) peak = fmax(peak, v[u].y);
            if (!isnan(v[u].y)) tot += area * v[u].y;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && bins) {  // the bins in front of / behind the 16-byte pairs
        for (int k = 0; k < 2; ++k) {
            const uint64_t j = k == 0 ? 0 : head + 2 * pairs;
            if ((k == 0 && !head) || j >= bins) continue;
            const double v = map[j];
            if (isfinite(v)) peak = fmax(peak, v);
            if (!isnan(v)) tot += area * v;
        }
    }
    peak = warp_max(peak); tot = warp_sum(tot);
    if ((threadIdx.x & 31) == 0) { atomicMax(&acc->peak_key, dkey(peak)); atomicAdd(&acc->total, tot); }
}

// ---- per-bundle peak-fluence check of the ghost-focus analysis (OpticSurface::evaluate_fluence_of_ray_bundle,
// optic_surface.rs:202-224; RaysHitMap::get_max_fluence, rays_hit_map.rs:777-798) in three launches without parameter uploads:
//   1. bounding box of the bundle's hits; the CTA that finishes last turns it into the range (-l, r, t, -b) and restores the
//      accumulator's identity for the next check; the first launch of a check also zero-fills the small map;
//   2. the Binning splat (k_binning);
//   3. one CTA: peak and total of the map, the bounce level the check used (the launch's first_valid_key).
// result[8] = range 4 | peak | total | status bits (u32, accumulated by the splat) | first_valid_key (u64)
__device__ __forceinline__ void peakcheck_bbox_body(const DevHits& H, uint64_t n, long long bounce_filter, PeakCheckAcc* acc,
                                                    double* __restrict__ range_out, double* __restrict__ map, uint64_t bins, int first,
                                                    int last) {
    if (first)
        for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < bins; k += (uint64_t)gridDim.x * blockDim.x) map[k] = 0.0;
    hits_bbox_body(H, n, resolve_bounce_filter(bounce_filter), &acc->box);
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(&acc->done, 1u) != gridDim.x - 1) return;
    __threadfence();
    const unsigned long long k0 = atomicMin(&acc->box.key[0], ~0ull), k1 = atomicMax(&acc->box.key[1], 0ull);
    const unsigned long long k2 = atomicMax(&acc->box.key[2], 0ull), k3 = atomicMin(&acc->box.key[3], ~0ull);
    const unsigned long long cnt = atomicAdd(&acc->box.count, 0ull);
    if (cnt == 0) { range_out[0] = range_out[1] = range_out[2] = range_out[3] = -INFINITY; }
    else { range_out[0] = -dkey_inv(k0); range_out[1] = dkey_inv(k1); range_out[2] = dkey_inv(k2); range_out[3] = -dkey_inv(k3); }
    acc->box.key[0] = ~0ull; acc->box.key[1] = 0ull; acc->box.key[2] = 0ull; acc->box.key[3] = ~0ull; acc->box.count = 0ull; acc->done = 0u;
}
__global__ void __launch_bounds__(256) peakcheck_bbox_kernel(DevHits H, uint64_t n, long long bounce_filter, PeakCheckAcc* acc,
                                                             double* __restrict__ range_out, double* __restrict__ map, uint64_t bins,
                                                             int first, int last) {
    peakcheck_bbox_body(H, n, bounce_filter, acc, range_out, map, bins, first, last);
}
__device__ __forceinline__ void peakcheck_peak_body(const double* __restrict__ map, uint64_t nx, uint64_t ny,
                                                    const unsigned long long* first_inv, double* __restrict__ result) {
    const double* range = result;
    const uint64_t bins = nx * ny;
    const double dx = (range[1] - (-range[0])) / (double)nx, dy = (range[2] - (-range[3])) / (double)ny;
    const double area = dx * dy;
    double peak = -INFINITY, tot = 0.0;
    for (uint64_t k = threadIdx.x; k < bins; k += blockDim.x) {
        const double v = map[k];
        if (isfinite(v)) peak = fmax(peak, v);
        if (!isnan(v)) tot += area * v;
    }
    peak = warp_max(peak); tot = warp_sum(tot);
    __shared__ double sp[32], st[32];
    if ((threadIdx.x & 31) == 0) { sp[threadIdx.x >> 5] = peak; st[threadIdx.x >> 5] = tot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { peak = fmax(peak, sp[k]); tot += st[k]; }
        result[4] = peak; result[5] = tot;
        const unsigned long long inv = first_inv ? *first_inv : 0ull;
        reinterpret_cast<unsigned long long*>(result)[7] = inv ? ~inv : ~0ull;
    }
}
__global__ void __launch_bounds__(1024) peakcheck_peak_kernel(const double* __restrict__ map, uint64_t nx, uint64_t ny,
                                                              const unsigned long long* first_inv, double* __restrict__ result) {
    peakcheck_peak_body(map, nx, ny, first_inv, result);
}
// The same three steps for MANY checks at once (the ghost-focus analysis defers its per-bundle checks: their verdicts steer no
// control flow and the hit maps stay): blockIdx.y (bbox, splat) / blockIdx.x (peak) picks the check, every check has its own
// accumulator, map and result slot; the bounce level of check k is the key word of its result slot (captured right after the
// bundle's launch).  Three launches for up to a few hundred checks instead of three latency-bound launches each.
// `mi`: which of the check's hit maps this launch takes (a batch runs max(n_maps) bounding-box launches, then as many splats: the
// first bounding-box launch zero-fills the map, a check's last one finalises its range, every splat adds to the map)
__global__ void __launch_bounds__(256) peakcheck_bbox_batch_kernel(const PeakCheckJob* __restrict__ jobs, uint64_t bins, int mi) {
    const PeakCheckJob* j = jobs + blockIdx.y;
    const int n_maps = j->n_maps;
    if (mi >= n_maps) return;
    const DevHits H = j->H[mi];  // by value: the body's loops must not go back to the job list for the array pointers
    const uint64_t n = j->n[mi];
    const long long filter = j->bounce_filter;
    PeakCheckAcc* const acc = j->acc;
    double* const result = j->result;
    double* const map = j->map;
    peakcheck_bbox_body(H, n, filter, acc, result, map, bins, mi == 0, mi == n_maps - 1);
}
template <int BLOCK, int MIN_CTAS, int UNROLL, bool FASTDIV>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) binning_fixed_batch_kernel(const PeakCheckJob* __restrict__ jobs, unsigned long long nx,
                                                                              unsigned long long ny, int mi) {
    const PeakCheckJob* j = jobs + blockIdx.y;
    if (mi >= j->n_maps || j->n[mi] == 0) return;  // a hit map without recorded slots has no splat
    // the check's arguments staged in shared memory: at 32 registers per thread (two CTAs of 1024 threads per SM) five array pointers
    // held by value are spilled and reloaded from local memory in the splat loop; the single-check kernel has them in the constant bank
    __shared__ DevHits sH;
    __shared__ uint64_t sn;
    __shared__ long long sfilter;
    __shared__ double* sresult;
    __shared__ double* smap;
    if (threadIdx.x == 0) { sH = j->H[mi]; sn = j->n[mi]; sfilter = j->bounce_filter; sresult = j->result; smap = j->map; }
    __syncthreads();
    binning_smem_fixed_body<BLOCK, UNROLL, FASTDIV>(sH, sn, sfilter, nx, ny, sresult, smap, reinterpret_cast<unsigned*>(sresult + 6));
}
__global__ void __launch_bounds__(1024) peakcheck_peak_batch_kernel(const PeakCheckJob* __restrict__ jobs, uint64_t nx, uint64_t ny) {
    const PeakCheckJob* j = jobs + blockIdx.x;
    double* const result = j->result;
    const double* const map = j->map;
    peakcheck_peak_body(map, nx, ny, reinterpret_cast<const unsigned long long*>(result) + 7, result);
}
__global__ void peakcheck_acc_init_kernel(PeakCheckAcc* acc, uint64_t n) {
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        acc[k].box.key[0] = ~0ull; acc[k].box.key[1] = 0ull; acc[k].box.key[2] = 0ull; acc[k].box.key[3] = ~0ull; acc[k].box.count = 0ull;
        acc[k].done = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// kernel density estimator (kde/mod.rs, kde/gaussian.rs): the KDE fluence estimator
// ---------------------------------------------------------------------------------------------
// dense (x, y, e) list of the recorded hits of a slot-indexed hit map (warp-aggregated append; order is not kept)
__global__ void __launch_bounds__(256) hits_pack_kernel(DevHits H, uint64_t n, long long bounce_filter, double* __restrict__ x,
                                                        double* __restrict__ y, double* __restrict__ e, unsigned long long* count,
                                                        unsigned long long cap) {
    bounce_filter = resolve_bounce_filter(bounce_filter);
    const uint64_t n_round = (n + 31) / 32 * 32;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        bool ok = false;
        double ev = -1.0;
        if (i < n) {
            ev = H.e[i];
            ok = ev >= 0.0 && !(bounce_filter >= 0 && (long long)H.bounce[i] != bounce_filter);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (!mask) continue;
        const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok) {
            unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
            if (slot < cap) { x[slot] = H.x[i]; y[slot] = H.y[i]; e[slot] = ev; }
        }
    }
}
// Kde::point_distances_std_dev, first half (kde/mod.rs:44-56): every pair distance (i < j) at its place in the
// reference's order, and their sum
__global__ void __launch_bounds__(256) kde_pair_distance_kernel(const double* __restrict__ x, const double* __restrict__ y, uint64_t n,
                                                                double* __restrict__ dist, double* sum) {
    const uint64_t i = blockIdx.y;
    const double xi = x[i], yi = y[i];
    const uint64_t base = i * (2 * n - i - 1) / 2;
    double s = 0.0;
    for (uint64_t j = i + 1 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        double dx = x[j] - xi, dy = y[j] - yi;  // math_utils::distance_2d_point
        double d = sqrt(dx * dx + dy * dy);
        dist[base + (j - i - 1)] = d;
        s += d;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(sum, s);
}
// second half (kde/mod.rs:57-63): sum of squared deviations from the mean distance
__global__ void __launch_bounds__(256) kde_deviation_kernel(const double* __restrict__ dist, uint64_t m, const double* __restrict__ sum,
                                                            double* dev) {
    const double avg = *sum / (double)m;
    double s = 0.0;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        double d = dist[k] - avg;
        s += d * d;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(dev, s);
}
// Kde::kde_2d (kde/mod.rs:114-134): one thread per map point, the hit list streamed through shared memory, summed in hit
// order like the reference's iterator sum.  map: column-major ny x nx.
__global__ void __launch_bounds__(256) kde_2d_kernel(const double* __restrict__ hx, const double* __restrict__ hy,
                                                     const double* __restrict__ he, uint64_t n, double band_width, double left,
                                                     double bottom, double dx, double dy, uint64_t nx, uint64_t ny,
                                                     double* __restrict__ map) {
    __shared__ double sx[256], sy[256], sa[256];
    const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const bool in = k < nx * ny;
    const uint64_t col = in ? k / ny : 0, row = in ? k % ny : 0;
    const double px = left + (double)col * dx, py = bottom + (double)row * dy;
    const double FRAC_1_2PI = 0.15915494309189535;
    const double sigma_square = band_width * band_width;
    double sum = 0.0;
    for (uint64_t t0 = 0; t0 < n; t0 += 256) {
        const uint64_t j = t0 + threadIdx.x;
        __syncthreads();
        if (j < n) { sx[threadIdx.x] = hx[j]; sy[threadIdx.x] = hy[j]; sa[threadIdx.x] = he[j] * FRAC_1_2PI / sigma_square; }
        __syncthreads();
        const int cnt = (int)((n - t0) < 256 ? (n - t0) : 256);
        for (int q = 0; q < cnt; ++q) {  // Gaussian2D::value (kde/gaussian.rs:24-27)
            double dist_square = (px - sx[q]) * (px - sx[q]) + (py - sy[q]) * (py - sy[q]);
            sum += sa[q] * exp(-dist_square / (2. * sigma_square));
        }
    }
    if (in) map[k] = sum;
}

// ---------------------------------------------------------------------------------------------
// sources
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fract(double x) { return x - trunc(x); }
__device__ __forceinline__ void source_position(const SourceParams& S, uint64_t ip, double& x, double& y) {
    switch (S.pos_dist) {
        case OPM_POS_GRID: {  // position_distributions/grid.rs:58-92 (x outer, y inner)
            uint64_t ix = ip / S.grid_ny, iy = ip % S.grid_ny;
            x = (double)ix * S.dist_x - S.off_x;
            y = (double)iy * S.dist_y - S.off_y;
            break;
        }
        case OPM_POS_FIBONACCI_RECT: {  // fibonacci.rs:53-65
            double fi = (double)ip;
            x = S.size_x * (fract(fi / S.golden) - 0.5);
            y = S.size_y * ((fi / (double)S.n_positions) - 0.5);
            break;
        }
        case OPM_POS_HEXAPOLAR: {  // hexapolar.rs:38-56: centre, then ring r (1-based) with 6 r points at radius r * R / rings
            x = 0.0; y = 0.0;
            if (ip == 0 || S.size_x == 0.0) break;
            uint64_t ring = (uint64_t)((sqrt(12.0 * (double)(ip - 1) + 9.0) - 3.0) / 6.0);  // 0-based: 1 + 3 ring (ring + 1) <= ip
            while (1 + 3 * ring * (ring + 1) > ip) --ring;
            while (1 + 3 * (ring + 1) * (ring + 2) <= ip) ++ring;
            const uint64_t point_nr = ip - (1 + 3 * ring * (ring + 1));
            const double radius = (double)(ring + 1) * S.dist_x;  // dist_x = radius_step
            const double angle_step = 2.0 * 3.14159265358979323846 / (double)(6 * (ring + 1));
            double sn, cs;
            sincos((double)point_nr * angle_step, &sn, &cs);
            x = radius * sn; y = radius * cs;
            break;
        }
        default: {  // fibonacci.rs:113-126
            double sn, cs;
            sincos(2.0 * 3.14159265358979323846 * fract((double)ip / S.golden), &sn, &cs);
            double sr = sqrt((double)ip / (double)S.n_positions);
            x = S.size_x * sn * sr;
            y = S.size_y * cs * sr;
            break;
        }
    }
}
// utils/math_distribution_functions.rs:54-70,133-150
__device__ __forceinline__ double general_gaussian_weight(const SourceParams& S, double px, double py) {
    double x_rot = fma(px - S.mu[0], S.cos_t, -((py - S.mu[1]) * S.sin_t));
    double y_rot = fma(py - S.mu[1], S.cos_t, (px - S.mu[0]) * S.sin_t);
    double ax = x_rot / S.sigma[0], ay = y_rot / S.sigma[1];
    if (S.rectangular) return exp(-pow(0.5 * (ax * ax), S.power) - pow(0.5 * (ay * ay), S.power));
    return exp(-pow(0.5 * fma(ax, ax, ay * ay), S.power));
}
__global__ void __launch_bounds__(256) source_norm_kernel(SourceParams S, uint64_t pos_begin, uint64_t pos_end, double* out) {
    double s = 0.0;
    for (uint64_t ip = pos_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; ip < pos_end; ip += (uint64_t)gridDim.x * blockDim.x) {
        double x, y;
        source_position(S, ip, x, y);
        s += general_gaussian_weight(S, x, y);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
// rays.rs:264-297 (position-major, wavelength-minor) + nodes/source.rs:205-210 (source isometry)
__global__ void __launch_bounds__(256) source_generate_kernel(SourceParams S, DevRays R, uint64_t pos_begin, uint64_t pos_stride,
                                                              uint64_t n_rays, const double* __restrict__ wvl,
                                                              const double* __restrict__ wgt, const double* __restrict__ norm_dev) {
    const V3 zdir = iso_vector(S.iso, v3(0.0, 0.0, 1.0));
    if (norm_dev) S.energy_norm = *norm_dev;  // the normalisation sum was taken on the device just before (source_norm_kernel)
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_rays; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ip = pos_begin + (k / S.n_wavelengths) * pos_stride;
        uint32_t iw = (uint32_t)(k % S.n_wavelengths);
        double x, y;
        source_position(S, ip, x, y);
        double e_pos = S.energy_dist == OPM_EDIST_UNIFORM ? S.e_uniform
                                                          : S.total_energy * general_gaussian_weight(S, x, y) / S.energy_norm;
        V3 p = iso_point(S.iso, v3(x, y, 0.0));
        R.px[k] = p.x; R.py[k] = p.y; R.pz[k] = p.z;
        R.dx[k] = zdir.x; R.dy[k] = zdir.y; R.dz[k] = zdir.z;
        R.e[k] = e_pos * wgt[iw]; R.wvl[k] = wvl[iw]; R.opl[k] = 0.0; R.n[k] = 1.0;
        R.bounces[k] = 0; R.refr[k] = 0; R.id[k] = (uint32_t)(ip * S.n_wavelengths + iw); R.valid[k] = RAY_VALID;
    }
}

// Rays::new_collimated_with_spectrum (rays.rs:264-297) from host-generated positions and energies: position-major,
// wavelength-minor, direction +z, E = E_pos * w
__global__ void __launch_bounds__(256) expand_collimated_kernel(const double* __restrict__ xyz, const double* __restrict__ e_pos,
                                                                const double* __restrict__ wvl, const double* __restrict__ wgt,
                                                                uint32_t n_wvl, DevRays R, uint64_t n_rays) {
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_rays; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ip = k / n_wvl;
        uint32_t iw = (uint32_t)(k % n_wvl);
        R.px[k] = xyz[3 * ip]; R.py[k] = xyz[3 * ip + 1]; R.pz[k] = xyz[3 * ip + 2];
        R.dx[k] = 0.0; R.dy[k] = 0.0; R.dz[k] = 1.0;
        R.e[k] = e_pos[ip] * wgt[iw]; R.wvl[k] = wvl[iw]; R.opl[k] = 0.0; R.n[k] = 1.0;
        R.bounces[k] = 0; R.refr[k] = 0; R.id[k] = (uint32_t)k; R.valid[k] = RAY_VALID;
    }
}

// The same from the PositionDistribution's points alone: every position distribution of the reference yields z = 0
// (position_distributions/*.rs) and the EnergyDistribution is a function of (x, y) (uniform.rs:37-41, general_gaussian.rs:93-120),
// so the host hands over 16 bytes per position and the energies are evaluated here.  norm: sum of the unnormalised Gaussian
// weights over all positions (S.energy_norm if the caller knows it, else *norm_dev from xy_norm_kernel).
__global__ void __launch_bounds__(256) xy_norm_kernel(const double2* __restrict__ xy, uint64_t n_pos, SourceParams S, double* out) {
    double s = 0.0;
    for (uint64_t ip = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; ip < n_pos; ip += (uint64_t)gridDim.x * blockDim.x) {
        const double2 p = xy[ip];
        s += general_gaussian_weight(S, p.x, p.y);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
__global__ void __launch_bounds__(256) expand_collimated_xy_kernel(const double2* __restrict__ xy, SourceParams S, const double* __restrict__ norm_dev,
                                                                   const double* __restrict__ wvl, const double* __restrict__ wgt,
                                                                   uint32_t n_wvl, DevRays R, uint64_t n_rays) {
    const V3 zdir = iso_vector(S.iso, v3(0.0, 0.0, 1.0));
    const double norm = S.energy_norm > 0.0 ? S.energy_norm : (norm_dev ? *norm_dev : 1.0);
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_rays; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t ip = k / n_wvl;
        const uint32_t iw = (uint32_t)(k % n_wvl);
        const double2 q = xy[ip];
        const double e_pos = S.energy_dist == OPM_EDIST_UNIFORM ? S.e_uniform : S.total_energy * general_gaussian_weight(S, q.x, q.y) / norm;
        const V3 p = iso_point(S.iso, v3(q.x, q.y, 0.0));
        R.px[k] = p.x; R.py[k] = p.y; R.pz[k] = p.z;
        R.dx[k] = zdir.x; R.dy[k] = zdir.y; R.dz[k] = zdir.z;
        R.e[k] = e_pos * wgt[iw]; R.wvl[k] = wvl[iw]; R.opl[k] = 0.0; R.n[k] = 1.0;
        R.bounces[k] = 0; R.refr[k] = 0; R.id[k] = (uint32_t)k; R.valid[k] = RAY_VALID;
    }
}

// ---------------------------------------------------------------------------------------------
// order-preserving compaction of removed rays
// ---------------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;
__global__ void __launch_bounds__(kCompactBlock) compact_count_kernel(const uint8_t* __restrict__ valid, uint64_t n, unsigned* counts) {
    uint64_t i = blockIdx.x * (uint64_t)kCompactBlock + threadIdx.x;
    bool keep = i < n && valid[i] != RAY_REMOVED;
    int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned)c;
}
// single-CTA exclusive scan of the block counts (nb is ~n/1024, tiny next to the ray traffic)
__global__ void __launch_bounds__(1024) compact_scan_kernel(const unsigned* __restrict__ counts, unsigned long long* offsets, uint64_t nb,
                                                            unsigned long long* total) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nb; base += 1024) {
        uint64_t i = base + threadIdx.x;
        unsigned long long v = i < nb ? counts[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = wsum[threadIdx.x], y = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, y, o); if (threadIdx.x >= o) y += t; }
            wsum[threadIdx.x] = y - w;
        }
        __syncthreads();
        unsigned long long incl = x + wsum[threadIdx.x >> 5] + carry;
        if (i < nb) offsets[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(kCompactBlock) compact_scatter_kernel(DevRays S, DevRays D, uint64_t n, const unsigned long long* __restrict__ offsets) {
    __shared__ unsigned wcount[32];
    uint64_t i = blockIdx.x * (uint64_t)kCompactBlock + threadIdx.x;
    bool keep = i < n && S.valid[i] != RAY_REMOVED;
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) wcount[w] = __popc(mask);
    __syncthreads();
    if (w == 0) {
        unsigned c = wcount[l], x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += t; }
        wcount[l] = x - c;
    }
    __syncthreads();
    if (keep) {
        uint64_t j = offsets[blockIdx.x] + wcount[w] + __popc(mask & ((1u << l) - 1u));
        D.px[j] = S.px[i]; D.py[j] = S.py[i]; D.pz[j] = S.pz[i];
        D.dx[j] = S.dx[i]; D.dy[j] = S.dy[i]; D.dz[j] = S.dz[i];
        D.e[j] = S.e[i]; D.wvl[j] = S.wvl[i]; D.opl[j] = S.opl[i]; D.n[j] = S.n[i];
        D.bounces[j] = S.bounces[i]; D.refr[j] = S.refr[i]; D.id[j] = S.id[i]; D.valid[j] = S.valid[i];
    }
}
__global__ void fill_u8_kernel(uint8_t* p, uint8_t v, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
// FP64 pipe peak: 8 independent DFMA chains per thread (roofline denominator of the FP64-bound fused kernel)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // keeps the chains alive
}
__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }
// all fourteen arrays of a bundle in one launch (Rays::clone, growth of a bundle, append): n slots from S[soff..] to D[doff..]
__global__ void __launch_bounds__(256) copy_rays_kernel(DevRays S, uint64_t soff, DevRays D, uint64_t doff, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = soff + i, b = doff + i;
        const double px = __ldcs(S.px + a), py = __ldcs(S.py + a), pz = __ldcs(S.pz + a), dx = __ldcs(S.dx + a), dy = __ldcs(S.dy + a);
        const double dz = __ldcs(S.dz + a), e = __ldcs(S.e + a), w = __ldcs(S.wvl + a), o = __ldcs(S.opl + a), nn = __ldcs(S.n + a);
        const uint32_t bo = __ldcs(S.bounces + a), rf = __ldcs(S.refr + a), id = __ldcs(S.id + a);
        const uint8_t v = S.valid[a];
        __stcs(D.px + b, px); __stcs(D.py + b, py); __stcs(D.pz + b, pz); __stcs(D.dx + b, dx); __stcs(D.dy + b, dy);
        __stcs(D.dz + b, dz); __stcs(D.e + b, e); __stcs(D.wvl + b, w); __stcs(D.opl + b, o); __stcs(D.n + b, nn);
        __stcs(D.bounces + b, bo); __stcs(D.refr + b, rf); __stcs(D.id + b, id);
        D.valid[b] = v;
    }
}

// small host -> device transfers that must not queue behind bulk uploads on the copy engine: the source is pinned host
// memory mapped into the device address space (unified addressing), read over PCIe by one CTA
__global__ void __launch_bounds__(256) fetch_from_host_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src_host,
                                                              unsigned words) {
    for (unsigned k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src_host[k];
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline int grid_for(uint64_t n, int block, int max_blocks) {
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (uint64_t)max_blocks) g = max_blocks;
    return (int)g;
}
#define ST ((cudaStream_t)stream)

int k_aos_to_soa(const OpmRay* a, DevRays R, uint64_t off, uint64_t n, int sms, void* stream) {
    aos_to_soa_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(a, R, off, n);
    return (int)cudaGetLastError();
}
int k_soa_to_aos(DevRays R, OpmRay* a, uint64_t n, int sms, void* stream) {
    soa_to_aos_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(R, a, n);
    return (int)cudaGetLastError();
}
int k_soa_defaults(DevRays R, uint64_t n, int sms, void* stream) {
    soa_defaults_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(R, n);
    return (int)cudaGetLastError();
}
int k_transform(DevRays R, DevIso I, uint64_t n, int sms, void* stream) {
    transform_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(R, I, n);
    return (int)cudaGetLastError();
}
int k_spot_pass1(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc, int sms, void* stream) {
    spot_pass1_kernel<<<grid_for(n, 256 * 4, sms * 8), 256, 0, ST>>>(R, n, use_frame, inv, acc);
    return (int)cudaGetLastError();
}
int k_spot_pass2(DevRays R, uint64_t n, int use_frame, DevIso inv, SpotAccum* acc, int sms, void* stream) {
    spot_pass2_kernel<<<grid_for(n, 256 * 4, sms * 8), 256, 0, ST>>>(R, n, use_frame, inv, acc);
    return (int)cudaGetLastError();
}
int k_bounce_lvl(DevRays R, uint64_t n, unsigned long long* first_key, int sms, void* stream) {
    bounce_lvl_kernel<<<grid_for(n, 256, sms * 4), 256, 0, ST>>>(R, n, first_key);
    return (int)cudaGetLastError();
}
int k_tally(DevRays R, uint64_t n, uint32_t n_orders, unsigned long long* counts, double* energy, int sms, void* stream) {
    tally_kernel<<<grid_for(n, 256, sms * 2), 256, n_orders * 16, ST>>>(R, n, n_orders, counts, energy);
    return (int)cudaGetLastError();
}
int k_hits_bbox(DevHits H, uint64_t n, long long bounce_filter, BBoxAccum* acc, int sms, void* stream) {
    hits_bbox_kernel<<<grid_for(n, 256 * 4, sms * 8), 256, 0, ST>>>(H, n, bounce_filter, acc);
    return (int)cudaGetLastError();
}
int k_peakcheck_acc_init(PeakCheckAcc* acc, uint64_t n, void* stream) {
    peakcheck_acc_init_kernel<<<grid_for(n, 256, 64), 256, 0, ST>>>(acc, n);
    return (int)cudaGetLastError();
}
int k_peakcheck_bbox(DevHits H, uint64_t n, long long bounce_filter, PeakCheckAcc* acc, double* range_out, double* map, uint64_t bins, int first,
                     int last, int sms, void* stream) {
    peakcheck_bbox_kernel<<<grid_for(n, 256 * 4, sms * 8), 256, 0, ST>>>(H, n, bounce_filter, acc, range_out, map, bins, first, last);
    return (int)cudaGetLastError();
}
// n_jobs checks in three launches; the CTAs of a launch are shared out among the checks (a few waves in total: the splat's
// CTA-private map is flushed once per CTA, so a check gets no more CTAs than keep the SMs busy)
// active[mi]: checks that have a hit map number mi (the others' CTAs of that launch return at once, so the launch's CTAs are shared
// out among the active ones: a second-map launch with two active checks gives each of them half the GPU, not a sixteenth)
int k_peakcheck_batch(const PeakCheckJob* d_jobs, int n_jobs, int max_maps, const int* active, uint64_t max_n, uint64_t nx, uint64_t ny, int sms,
                      void* stream) {
    if (n_jobs <= 0) return 0;
    if (max_maps < 1 || max_maps > kPeakCheckMaps) return (int)cudaErrorInvalidValue;
    const uint64_t bins = nx * ny;
    const size_t bytes = (size_t)bins * sizeof(double);
    if (bytes > 160 * 1024 || n_jobs > 65535) return (int)cudaErrorInvalidValue;
    auto share = [&](int total, int most, int among) {
        if (among < 1) among = 1;
        int g = (total + among - 1) / among;
        return g < 1 ? 1 : (g > most ? most : g);
    };
    for (int mi = 0; mi < max_maps; ++mi) {
        const int g_bbox = share(sms * 8, grid_for(max_n, 256 * 4, sms * 8), active ? active[mi] : n_jobs);
        peakcheck_bbox_batch_kernel<<<dim3(g_bbox, n_jobs), 256, 0, ST>>>(d_jobs, bins, mi);
    }
    const int ctas = (2 * (bytes + 1024) <= 227 * 1024) ? 2 : 1;
    cudaError_t e = cudaFuncSetAttribute(binning_fixed_batch_kernel<1024, 2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    for (int mi = 0; mi < max_maps; ++mi) {
        const int g_bin = share(sms * ctas * 2, grid_for(max_n, 1024, sms * ctas), active ? active[mi] : n_jobs);
        binning_fixed_batch_kernel<1024, 2, 2, true><<<dim3(g_bin, n_jobs), 1024, bytes, ST>>>(d_jobs, nx, ny, mi);
    }
    peakcheck_peak_batch_kernel<<<n_jobs, 1024, 0, ST>>>(d_jobs, nx, ny);
    return (int)cudaGetLastError();
}
int k_peakcheck_peak(const double* map, uint64_t nx, uint64_t ny, const unsigned long long* first_inv, double* result, void* stream) {
    peakcheck_peak_kernel<<<1, 1024, 0, ST>>>(map, nx, ny, first_inv, result);
    return (int)cudaGetLastError();
}
int k_bbox_finalize(const BBoxAccum* acc, double* range_dev, void* stream) {
    bbox_finalize_kernel<<<1, 1, 0, ST>>>(acc, range_dev);
    return (int)cudaGetLastError();
}
int k_binning(DevHits H, uint64_t n, long long bounce_filter, uint64_t nx, uint64_t ny, const double* range_dev, double* map,
              unsigned* status, int sms, void* stream) {
    size_t bytes = (size_t)(nx * ny) * sizeof(double);
    static const bool force_global = getenv("OPM_B200_BINNING_GLOBAL") != nullptr;  // tuning switch
    // OPM_B200_BINNING_FIXED=0: the f64 compare-and-swap kernel (tuning switch; profiles/binning_variants.sh)
    static const bool fixed = [] { const char* e = getenv("OPM_B200_BINNING_FIXED"); return e ? atoi(e) != 0 : true; }();
    if (bytes <= 160 * 1024 && !force_global && fixed) {
        // two CTAs of 1024 threads per SM when two private maps fit in its shared memory (all 64 warps of the SM hide the
        // latency of the adds), else one; measured on the C2 detectors (10^7 slots, 100 x 83): 200 -> 85 us per map
        const int ctas = (2 * (bytes + 1024) <= 227 * 1024) ? 2 : 1;
        cudaError_t e = cudaFuncSetAttribute(binning_smem_fixed_kernel<1024, 2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        binning_smem_fixed_kernel<1024, 2, 2, true><<<grid_for(n, 1024, sms * ctas), 1024, bytes, ST>>>(H, n, bounce_filter, nx, ny, range_dev,
                                                                                                      map, status);
    } else if (bytes <= 160 * 1024 && !force_global) {
        cudaError_t e = cudaFuncSetAttribute(binning_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        // the privatised map takes most of an SM's shared memory: one CTA per SM, as many warps as it can hold
        binning_smem_kernel<<<grid_for(n, 1024, sms), 1024, bytes, ST>>>(H, n, bounce_filter, nx, ny, range_dev, map, status);
    } else {
        // tuning switches (profiles/c4_binning_probe.py): OPM_B200_BINNING_LARGE=warp selects the warp-segmented kernel,
        // OPM_B200_BINNING_RUNS="shape,ctas_per_sm,interleave" the launch shape of the run-merging one
        const char* which = getenv("OPM_B200_BINNING_LARGE");
        if ((which && !strcmp(which, "warp")) || nx * ny >= (1ull << 31)) {
            binning_global_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(H, n, bounce_filter, nx, ny, range_dev, map, status);
            return (int)cudaGetLastError();
        }
        int shape = 0, ctas = 0, interleave = 1;
        if (const char* cfg = getenv("OPM_B200_BINNING_RUNS")) sscanf(cfg, "%d,%d,%d", &shape, &ctas, &interleave);
#define OPM_RUNS(BLOCK, MINC, RUN, PF)                                                                                               \
    {                                                                                                                                \
        const int per_sm = ctas > 0 ? ctas : MINC;                                                                                   \
        binning_runs_kernel<BLOCK, MINC, RUN, PF><<<grid_for((n + RUN - 1) / RUN, BLOCK, sms * per_sm), BLOCK, 0, ST>>>(             \
            H, n, bounce_filter, nx, ny, range_dev, map, status, interleave);                                                        \
    }
#define OPM_RUNS_TMA(BLOCK, MINC, RUN, STAGES)                                                                                       \
    {                                                                                                                                \
        const int per_sm = ctas > 0 ? ctas : MINC;                                                                                   \
        const size_t smem = (size_t)STAGES * 3 * BLOCK * RUN * 8;                                                                    \
        auto kern = binning_runs_tma_kernel<BLOCK, MINC, RUN, STAGES>;                                                               \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
        if (e != cudaSuccess) return (int)e;                                                                                         \
        kern<<<grid_for((n + RUN - 1) / RUN, BLOCK, sms * per_sm), BLOCK, smem, ST>>>(H, n, bounce_filter, nx, ny, range_dev, map,   \
                                                                                      status);                                     \
    }
        // measured on the C4 hit map (10^8 hits -> 4096 x 4096, profiles/r2_summary.md): 0 -> 0.709 ms (zero-fill + splat), 1 -> 0.739,
        // 2 -> 0.765, 3 -> 0.720 (staged by TMA: no better -- the splat is bound by the SMs' rate of global reductions, not by the
        // latency of its loads), the warp-segmented kernel of round 1 -> 1.261 ms
        switch (shape) {
            case 1: OPM_RUNS(256, 4, 4, false) break;
            case 2: OPM_RUNS(256, 2, 8, false) break;
            case 3: OPM_RUNS_TMA(512, 2, 4, 2) break;
            default: OPM_RUNS(256, 3, 4, true) break;
        }
#undef OPM_RUNS
#undef OPM_RUNS_TMA
    }
    return (int)cudaGetLastError();
}
// Ray::position_history / position_history_with_current (ray.rs:296-327): the logged positions of every slot in level
// order (NaN rows = levels at which the ray logged nothing), ray-major; len = 0xFFFFFFFF marks a removed slot
__global__ void __launch_bounds__(256) history_gather_kernel(const double* __restrict__ base, uint64_t stride, uint32_t levels, DevRays R,
                                                             uint64_t n, int with_current, uint32_t pos_cap, double* __restrict__ xyz,
                                                             uint32_t* __restrict__ len) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] == RAY_REMOVED) { len[i] = 0xFFFFFFFFu; continue; }
        uint32_t k = 0;
        double* out = xyz + i * (uint64_t)pos_cap * 3;
        for (uint32_t l = 0; l < levels; ++l) {
            const double* row = base + (size_t)l * 3 * stride + i;
            const double x = row[0];
            if (isnan(x)) continue;
            if (k < pos_cap) { out[3 * k] = x; out[3 * k + 1] = row[stride]; out[3 * k + 2] = row[2 * stride]; }
            ++k;
        }
        if (with_current) {
            if (k < pos_cap) { out[3 * k] = R.px[i]; out[3 * k + 1] = R.py[i]; out[3 * k + 2] = R.pz[i]; }
            ++k;
        }
        len[i] = k;
    }
}
int k_history_gather(const HistParams& hp, uint32_t levels, DevRays R, uint64_t n, int with_current, uint32_t pos_cap, double* xyz,
                     uint32_t* len, int sms, void* stream) {
    history_gather_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(hp.base, hp.stride, levels, R, n, with_current, pos_cap, xyz, len);
    return (int)cudaGetLastError();
}
__global__ void peak_finalize_kernel(const PeakAccum* acc, double* out2) { out2[0] = dkey_inv(acc->peak_key); out2[1] = acc->total; }
// bit b set when a recorded hit has bounce count b (< 64)
__global__ void __launch_bounds__(256) hits_bounce_mask_kernel(DevHits H, uint64_t n, unsigned long long* mask) {
    unsigned long long m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (H.e[i] >= 0.0 && H.bounce[i] < 64u) m |= 1ull << H.bounce[i];
    m |= __shfl_xor_sync(0xffffffffu, m, 16); m |= __shfl_xor_sync(0xffffffffu, m, 8); m |= __shfl_xor_sync(0xffffffffu, m, 4);
    m |= __shfl_xor_sync(0xffffffffu, m, 2); m |= __shfl_xor_sync(0xffffffffu, m, 1);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(mask, m);
}
int k_hits_bounce_mask(DevHits H, uint64_t n, unsigned long long* mask, int sms, void* stream) {
    hits_bounce_mask_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(H, n, mask);
    return (int)cudaGetLastError();
}
int k_hits_pack(DevHits H, uint64_t n, long long bounce_filter, double* x, double* y, double* e, unsigned long long* count,
                unsigned long long cap, int sms, void* stream) {
    hits_pack_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(H, n, bounce_filter, x, y, e, count, cap);
    return (int)cudaGetLastError();
}
int k_kde_pair_distances(const double* x, const double* y, uint64_t n, double* dist, double* sum, double* dev, int sms, void* stream) {
    if (n < 2) return 0;
    const uint64_t m = n * (n - 1) / 2;
    unsigned gx = (unsigned)((n + 255) / 256);
    if (gx > 64) gx = 64;
    kde_pair_distance_kernel<<<dim3(gx, (unsigned)(n - 1)), 256, 0, ST>>>(x, y, n, dist, sum);
    kde_deviation_kernel<<<grid_for(m, 256, sms * 8), 256, 0, ST>>>(dist, m, sum, dev);
    return (int)cudaGetLastError();
}
// ascending sort of the pair distances (the 75 % quantile of kde/mod.rs:65-81 is read from the sorted list)
int k_sort_doubles(const double* in, double* out, uint64_t m, void* temp, size_t* temp_bytes, void* stream) {
    if (m > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    return (int)cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, in, out, (int)m, 0, 64, ST);
}
int k_kde_2d(const double* hx, const double* hy, const double* he, uint64_t n, double band_width, double left, double bottom, double dx,
             double dy, uint64_t nx, uint64_t ny, double* map, void* stream) {
    const uint64_t pts = nx * ny;
    kde_2d_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, ST>>>(hx, hy, he, n, band_width, left, bottom, dx, dy, nx, ny, map);
    return (int)cudaGetLastError();
}
int k_peak_total(const double* map, uint64_t nx, uint64_t ny, const double* range_dev, PeakAccum* acc, double* out2_dev, int sms, void* stream) {
    peak_total_kernel<<<grid_for((nx * ny + 7) / 8, 256, sms * 8), 256, 0, ST>>>(map, nx, ny, range_dev, acc);
    if (out2_dev) peak_finalize_kernel<<<1, 1, 0, ST>>>(acc, out2_dev);
    return (int)cudaGetLastError();
}
// spot statistics: accumulator <-> the reduce-friendly double layouts
//   sums[11]  = x y z e*x e*y e*z e | n_valid | n_total | bounce count and id of the first valid ray (-1: none)
//   radii[3]  = max d^2 [mm^2] | sum d^2 [mm^2] | sum e*d^2 [m^2]
__global__ void spot_sums_store_kernel(const SpotAccum* acc, double* sums) {
    for (int j = 0; j < 7; ++j) sums[j] = acc->sum[j];
    sums[7] = (double)acc->n_valid; sums[8] = (double)acc->n_total;
    sums[9] = acc->n_valid ? (double)(acc->first_key & 0xffffffffull) : -1.0;
    sums[10] = acc->n_valid ? (double)(acc->first_key >> 32) : -1.0;
}
__global__ void spot_sums_load_kernel(const double* sums, SpotAccum* acc) {
    for (int j = 0; j < 7; ++j) acc->sum[j] = sums[j];
    acc->n_valid = (unsigned long long)sums[7]; acc->n_total = (unsigned long long)sums[8];
    acc->first_key = sums[9] < 0.0 ? ~0ull : (((unsigned long long)sums[10] << 32) | (unsigned long long)sums[9]);
    acc->max_d2_key = dkey(0.0); acc->sum_d2 = 0.0; acc->sum_ed2 = 0.0;
}
__global__ void spot_radii_store_kernel(const SpotAccum* acc, double* radii) {
    radii[0] = dkey_inv(acc->max_d2_key); radii[1] = acc->sum_d2; radii[2] = acc->sum_ed2;
}
int k_spot_sums_store(const SpotAccum* acc, double* sums_dev, void* stream) {
    spot_sums_store_kernel<<<1, 1, 0, ST>>>(acc, sums_dev);
    return (int)cudaGetLastError();
}
int k_spot_sums_load(const double* sums_dev, SpotAccum* acc, void* stream) {
    spot_sums_load_kernel<<<1, 1, 0, ST>>>(sums_dev, acc);
    return (int)cudaGetLastError();
}
int k_spot_radii_store(const SpotAccum* acc, double* radii_dev, void* stream) {
    spot_radii_store_kernel<<<1, 1, 0, ST>>>(acc, radii_dev);
    return (int)cudaGetLastError();
}
int k_source_norm(const SourceParams& S, uint64_t b, uint64_t e, double* out, int sms, void* stream) {
    source_norm_kernel<<<grid_for(e - b, 256, sms * 4), 256, 0, ST>>>(S, b, e, out);
    return (int)cudaGetLastError();
}
int k_source_generate(const SourceParams& S, DevRays R, uint64_t pos_begin, uint64_t pos_stride, uint64_t n_rays, const double* wvl,
                      const double* wgt, const double* norm_dev, int sms, void* stream) {
    source_generate_kernel<<<grid_for(n_rays, 256, sms * 8), 256, 0, ST>>>(S, R, pos_begin, pos_stride, n_rays, wvl, wgt, norm_dev);
    return (int)cudaGetLastError();
}
int k_expand_collimated(const double* xyz, const double* e_pos, const double* wvl, const double* wgt, uint32_t n_wvl, DevRays R,
                        uint64_t n_rays, int sms, void* stream) {
    expand_collimated_kernel<<<grid_for(n_rays, 256, sms * 8), 256, 0, ST>>>(xyz, e_pos, wvl, wgt, n_wvl, R, n_rays);
    return (int)cudaGetLastError();
}
int k_expand_collimated_xy(const double* xy, uint64_t n_pos, const SourceParams& S, double* norm_dev, const double* wvl, const double* wgt,
                           uint32_t n_wvl, DevRays R, uint64_t n_rays, int sms, void* stream) {
    const bool need_norm = S.energy_dist == OPM_EDIST_GENERAL_GAUSSIAN && !(S.energy_norm > 0.0);
    if (need_norm) {
        cudaError_t e = cudaMemsetAsync(norm_dev, 0, 8, ST);
        if (e != cudaSuccess) return (int)e;
        if (n_pos) xy_norm_kernel<<<grid_for(n_pos, 256, sms * 8), 256, 0, ST>>>((const double2*)xy, n_pos, S, norm_dev);
    }
    if (n_rays)
        expand_collimated_xy_kernel<<<grid_for(n_rays, 256, sms * 8), 256, 0, ST>>>((const double2*)xy, S, need_norm ? norm_dev : nullptr, wvl, wgt,
                                                                                   n_wvl, R, n_rays);
    return (int)cudaGetLastError();
}
int k_compact(DevRays S, DevRays D, uint64_t n, unsigned* counts, unsigned long long* offsets, unsigned long long* total, void* stream) {
    uint64_t nb = (n + kCompactBlock - 1) / kCompactBlock;
    if (nb == 0) return 0;
    compact_count_kernel<<<(unsigned)nb, kCompactBlock, 0, ST>>>(S.valid, n, counts);
    compact_scan_kernel<<<1, 1024, 0, ST>>>(counts, offsets, nb, total);
    compact_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, ST>>>(S, D, n, offsets);
    return (int)cudaGetLastError();
}
int k_fp64_peak(double* out, int iters, int blocks, void* stream) {
    fp64_peak_kernel<<<blocks, 256, 0, ST>>>(out, iters, 1.0);
    return (int)cudaGetLastError();
}
int k_fetch_from_host(void* dst, const void* src_host, size_t bytes, void* stream) {
    fetch_from_host_kernel<<<1, 256, 0, ST>>>((unsigned long long*)dst, (const unsigned long long*)src_host, (unsigned)(bytes / 8));
    return (int)cudaGetLastError();
}
// Everything a trace launch needs on the device before it starts, in ONE small launch instead of four: the op descriptors of
// this launch (pinned host -> device, over PCIe by the CTA itself), the per-launch statistics zeroed, the bounding-box
// accumulators of the hit maps the kernel fills reset to their identity, and the lengths of the slot-indexed output bundles.
__global__ void __launch_bounds__(256) trace_prologue_kernel(ProloguePlan P) {
    for (unsigned k = threadIdx.x; k < P.words; k += blockDim.x) P.dst[k] = P.src_host[k];
    if (threadIdx.x < P.stats_words) {
        // deferred opm_fluence_peak_check_capture_key_async: the previous launch's keys leave with the launch that resets them
        static_assert(offsetof(DevStats, first_inv) == 8 * 8 && offsetof(DevStats, first_inv2) == 9 * 8, "key words of DevStats");
        if (threadIdx.x == 8 && P.key_dst[0]) *P.key_dst[0] = P.stats[8];
        if (threadIdx.x == 9 && P.key_dst[1]) *P.key_dst[1] = P.stats[9];
        P.stats[threadIdx.x] = 0ull;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 5 * 2) {
        const unsigned a = (threadIdx.x - 32) / 5, j = (threadIdx.x - 32) % 5;
        if (P.acc[a]) {
            if (j < 4) P.acc[a]->key[j] = (j == 0 || j == 3) ? ~0ull : 0ull;
            else P.acc[a]->count = 0ull;
        }
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + (unsigned)P.n_tails) *P.tail[threadIdx.x - 64] = P.tail_val[threadIdx.x - 64];
    if (threadIdx.x >= 96 && threadIdx.x < 98 && P.spot[threadIdx.x - 96]) {
        SpotAccum* a = P.spot[threadIdx.x - 96];
        for (int j = 0; j < 7; ++j) a->sum[j] = 0.0;
        a->n_valid = 0ull; a->n_total = 0ull; a->first_key = ~0ull;
        a->max_d2_key = dkey(0.0); a->sum_d2 = 0.0; a->sum_ed2 = 0.0;
    }
}
int k_trace_prologue(const ProloguePlan& P, void* stream) {
    trace_prologue_kernel<<<1, 256, 0, ST>>>(P);
    return (int)cudaGetLastError();
}
// DevStats::first_inv (max of ~(id << 32 | bounces) over the valid rays, 0 = none) -> the public first_valid_key
__global__ void first_key_store_kernel(const unsigned long long* first_inv, unsigned long long* out) {
    const unsigned long long inv = *first_inv;
    *out = inv ? ~inv : ~0ull;
}
int k_first_key_store(const unsigned long long* first_inv, unsigned long long* out, void* stream) {
    first_key_store_kernel<<<1, 1, 0, ST>>>(first_inv, out);
    return (int)cudaGetLastError();
}
int k_copy_rays(DevRays S, uint64_t soff, DevRays D, uint64_t doff, uint64_t n, int sms, void* stream) {
    copy_rays_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(S, soff, D, doff, n);
    return (int)cudaGetLastError();
}
int k_set_u64(unsigned long long* p, unsigned long long v, void* stream) {
    set_u64_kernel<<<1, 1, 0, ST>>>(p, v);
    return (int)cudaGetLastError();
}
// FluenceRays::new (rays.rs:1477-1516): three helper rays on a triangle of area wavelength^2 around the ray (outer radius
// sqrt(4 area / sqrt 27)), its direction (normalised by Ray::new, ray.rs:127), index 1; effective energy = fluence * area
struct HelperTrig { double c[6]; };
__global__ void __launch_bounds__(256) helpers_init_kernel(DevRays R, uint64_t n, double* __restrict__ base, uint64_t stride, HelperTrig T) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double wvl = R.wvl[i];
        const double area = wvl * wvl;
        const double fluence = base[21 * stride + i];
        const double radius = sqrt(4. * area / sqrt(27.));
        const V3 d = normalize(v3(R.dx[i], R.dy[i], R.dz[i]));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            base[(3 * k) * stride + i] = R.px[i] + radius * T.c[2 * k];
            base[(3 * k + 1) * stride + i] = R.py[i] + radius * T.c[2 * k + 1];
            base[(3 * k + 2) * stride + i] = R.pz[i] + 0.;
            base[(9 + 3 * k) * stride + i] = d.x; base[(9 + 3 * k + 1) * stride + i] = d.y; base[(9 + 3 * k + 2) * stride + i] = d.z;
            base[(18 + k) * stride + i] = 1.0;
        }
        base[21 * stride + i] = fluence * area;
    }
}
int k_helpers_init(DevRays R, uint64_t n, double* base, uint64_t stride, const double cs[6], int sms, void* stream) {
    HelperTrig T;
    for (int k = 0; k < 6; ++k) T.c[k] = cs[k];
    helpers_init_kernel<<<grid_for(n, 256, sms * 4), 256, 0, ST>>>(R, n, base, stride, T);
    return (int)cudaGetLastError();
}
int k_fill_u8(uint8_t* p, uint8_t v, uint64_t n, int sms, void* stream) {
    fill_u8_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(p, v, n);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// spectrometer and wavefront detectors (spectrum.rs, nodes/spectrometer.rs, nodes/wavefront.rs)
// ---------------------------------------------------------------------------------------------
// smallest / largest wavelength and number of the valid rays (Spectrum::from_laser_lines, spectrum.rs:133-142)
__global__ void __launch_bounds__(256) wavelength_range_kernel(DevRays R, uint64_t n, WvlAccum* acc) {
    double mn = INFINITY, mx = -INFINITY;
    unsigned long long cnt = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        const double w = R.wvl[i];
        mn = fmin(mn, w); mx = fmax(mx, w); cnt++;
    }
    mn = warp_min(mn); mx = warp_max(mx); cnt = warp_sum_u64(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMin(&acc->min_key, dkey(mn)); atomicMax(&acc->max_key, dkey(mx)); atomicAdd(&acc->count, cnt);
    }
}
// Spectrum::add_single_peak (spectrum.rs:193-227) for every valid ray: slot i holds wavelength fma(i, step, start) [um]; the
// ray's energy per micrometer goes to the two slots around its wavelength.  Slots privatised per CTA in shared memory.
__global__ void __launch_bounds__(256) spectrum_splat_kernel(DevRays R, uint64_t n, double start, double step, unsigned long long nslots,
                                                             double* __restrict__ data, int use_smem, unsigned* status_out) {
    extern __shared__ double sdata[];
    if (use_smem) {
        for (unsigned long long k = threadIdx.x; k < nslots; k += blockDim.x) sdata[k] = 0.0;
        __syncthreads();
    }
    double* dst = use_smem ? sdata : data;
    const double r0 = start * 1e-6;                                    // range(): micrometer!(first)..micrometer!(last)
    const double r1 = fma((double)(nslots - 1), step, start) * 1e-6;
    unsigned status = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        const double wl = R.wvl[i], value = R.e[i];
        if (!(r0 <= wl && wl < r1)) continue;                          // "peak wavelength is not in spectrum range. Spectrum unmodified."
        if (value < 0.0 || nslots < 2) { status |= 1u << OPM_ERR_SPECTRUM; continue; }
        const double w = wl * 1e6;
        // first slot with lambda >= w: estimate, then settle with the slots' own arithmetic
        double g = ceil((w - start) / step);
        long long idx = g > 0.0 ? (g < (double)nslots ? (long long)g : (long long)nslots) : 0;
        while (idx > 0 && fma((double)(idx - 1), step, start) >= w) --idx;
        while (idx < (long long)nslots && fma((double)idx, step, start) < w) ++idx;
        if (idx >= (long long)nslots) { status |= 1u << OPM_ERR_SPECTRUM; continue; }  // "insertion point not found"
        if (idx == 0) {
            const double delta = fma(1.0, step, start) - start;
            atomicAdd(&dst[0], value / delta);
        } else {
            const double lower = fma((double)(idx - 1), step, start), upper = fma((double)idx, step, start);
            const double delta = upper - lower;
            const double epm = value / delta;
            const double part = epm * (w - lower) / delta;
            atomicAdd(&dst[idx], part);
            atomicAdd(&dst[idx - 1], epm - part);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (unsigned long long k = threadIdx.x; k < nslots; k += blockDim.x) {
            const double v = sdata[k];
            if (v != 0.0) atomicAdd(&data[k], v);
        }
    }
    if (status) atomicOr(status_out, status);
}

// Rays::wavefront_error_at_pos_in_units_of_wvl (rays.rs:769-802).  x, y in mm in the monitor frame; the reference ray is
// the valid ray with the smallest x^2 + y^2, the first one in bundle order on ties.
__device__ __forceinline__ void wf_xy(const DevRays& R, uint64_t i, const DevIso& inv, double& x, double& y) {
    const V3 p = iso_point(inv, v3(R.px[i], R.py[i], R.pz[i]));
    x = p.x * 1e3; y = p.y * 1e3;
}
__global__ void __launch_bounds__(256) wavefront_center1_kernel(DevRays R, uint64_t n, DevIso inv, WfAccum* acc) {
    double mn = INFINITY;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        double x, y;
        wf_xy(R, i, inv, x, y);
        mn = fmin(mn, fma(x, x, y * y));
    }
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0 && mn < INFINITY) atomicMin(&acc->min_r_key, dkey(mn));
}
__global__ void __launch_bounds__(256) wavefront_center2_kernel(DevRays R, uint64_t n, DevIso inv, WfAccum* acc) {
    const unsigned long long key = acc->min_r_key;
    unsigned long long first = ~0ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (R.valid[i] != RAY_VALID) continue;
        double x, y;
        wf_xy(R, i, inv, x, y);
        if (dkey(fma(x, x, y * y)) == key && i < first) first = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, first, o); first = t < first ? t : first; }
    if ((threadIdx.x & 31) == 0 && first != ~0ull) atomicMin(&acc->first_slot, first);
}
// pass 3: error of every valid ray (optionally written slot-indexed, NaN = not a valid ray), max / min / sum / count;
// pass 4 (second = 1): sum of squared deviations from the mean of pass 3
__global__ void __launch_bounds__(256) wavefront_stats_kernel(DevRays R, uint64_t n, DevIso inv, double wvl_nm, WfAccum* acc,
                                                              double* __restrict__ xyw, int second) {
    const unsigned long long fs = acc->first_slot;
    const double center = fs != ~0ull ? -(R.opl[fs] * (1.0 / 1e-9)) : 0.0;
    const double avg = second ? acc->sum / (double)acc->count : 0.0;
    double mx = -INFINITY, mn = INFINITY, sum = 0.0;
    unsigned long long cnt = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool ok = R.valid[i] == RAY_VALID;
        double x = NAN, y = NAN, wf = NAN;
        if (ok) {
            wf_xy(R, i, inv, x, y);
            wf = -(R.opl[i] * (1.0 / 1e-9));
            wf -= center;
            wf /= wvl_nm;
            if (second) { const double d = wf - avg; sum += d * d; }
            else { mx = fmax(mx, wf); mn = fmin(mn, wf); sum += wf; cnt++; }
        }
        if (!second && xyw) { xyw[3 * i] = x; xyw[3 * i + 1] = y; xyw[3 * i + 2] = wf; }
    }
    sum = warp_sum(sum);
    if (second) {
        if ((threadIdx.x & 31) == 0) atomicAdd(&acc->sum2, sum);
        return;
    }
    mx = warp_max(mx); mn = warp_min(mn); cnt = warp_sum_u64(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMax(&acc->max_key, dkey(mx)); atomicMin(&acc->min_key, dkey(mn)); atomicAdd(&acc->sum, sum); atomicAdd(&acc->count, cnt);
    }
}

int k_wavelength_range(DevRays R, uint64_t n, WvlAccum* acc, int sms, void* stream) {
    wavelength_range_kernel<<<grid_for(n, 256, sms * 8), 256, 0, ST>>>(R, n, acc);
    return (int)cudaGetLastError();
}
int k_spectrum_splat(DevRays R, uint64_t n, double start_um, double step_um, uint64_t nslots, double* data, unsigned* status, int sms,
                     void* stream) {
    const size_t bytes = (size_t)nslots * sizeof(double);
    const int use_smem = bytes <= 96 * 1024;
    if (use_smem) {
        cudaError_t e = cudaFuncSetAttribute(spectrum_splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    spectrum_splat_kernel<<<grid_for(n, 256, sms * 2), 256, use_smem ? bytes : 0, ST>>>(R, n, start_um, step_um, nslots, data, use_smem, status);
    return (int)cudaGetLastError();
}
int k_wavefront(DevRays R, uint64_t n, DevIso inv, double wvl_nm, WfAccum* acc, double* xyw, int sms, void* stream) {
    const int g = grid_for(n, 256, sms * 8);
    wavefront_center1_kernel<<<g, 256, 0, ST>>>(R, n, inv, acc);
    wavefront_center2_kernel<<<g, 256, 0, ST>>>(R, n, inv, acc);
    wavefront_stats_kernel<<<g, 256, 0, ST>>>(R, n, inv, wvl_nm, acc, xyw, 0);
    wavefront_stats_kernel<<<g, 256, 0, ST>>>(R, n, inv, wvl_nm, acc, nullptr, 1);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// multi-GPU detector read-out: what one all-gather collected from `w` ranks, combined in rank order (every rank
// gets bit-identical results).  Row r of `g` is rank r's block.
// ---------------------------------------------------------------------------------------------
// block A = [4 x nf map ranges as (-l, r, t, -b) | 11 x ns spot sums]: ranges by max; sums[0..8] by sum; sums[9..10] =
// (bounces, id) of the first valid ray = the rank with the smallest id >= 0, the lowest rank on ties (rays.rs:183-194)
__global__ void readout_combine_a_kernel(const double* __restrict__ g, int w, int nf, int ns, double* __restrict__ out) {
    const int nA = 4 * nf + 11 * ns;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    if (i < 4 * nf) {
        double m = g[i];
        for (int r = 1; r < w; ++r) m = fmax(m, g[(size_t)r * nA + i]);
        out[i] = m;
        return;
    }
    const int k = (i - 4 * nf) / 11, c = (i - 4 * nf) % 11, o = 4 * nf + 11 * k;
    if (c < 9) {
        double s = 0.0;
        for (int r = 0; r < w; ++r) s += g[(size_t)r * nA + i];
        out[i] = s;
    } else {
        int first = 0;
        double best = INFINITY;
        for (int r = 0; r < w; ++r) {
            const double id = g[(size_t)r * nA + o + 10];
            if (id >= 0.0 && id < best) { best = id; first = r; }
        }
        out[i] = g[(size_t)first * nA + i];
    }
}
// block B = [maps_len map bins | 3 x ns spot radius accumulators (max d^2, sum d^2, sum E d^2)]: bins and sums by sum, max by max
__global__ void __launch_bounds__(256) readout_combine_b_kernel(const double* __restrict__ g, int w, uint64_t maps_len, int ns,
                                                                double* __restrict__ out) {
    const uint64_t len = maps_len + 3ull * ns;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool is_max = i >= maps_len && (i - maps_len) % 3 == 0;
        double a = g[i];
        for (int r = 1; r < w; ++r) { const double v = g[(size_t)r * len + i]; a = is_max ? fmax(a, v) : a + v; }
        out[i] = a;
    }
}
int k_readout_combine_a(const double* g, int w, int nf, int ns, double* out, void* stream) {
    const int nA = 4 * nf + 11 * ns;
    if (nA <= 0) return 0;
    readout_combine_a_kernel<<<(nA + 127) / 128, 128, 0, ST>>>(g, w, nf, ns, out);
    return (int)cudaGetLastError();
}
int k_readout_combine_b(const double* g, int w, uint64_t maps_len, int ns, double* out, int sms, void* stream) {
    const uint64_t len = maps_len + 3ull * ns;
    if (len == 0) return 0;
    readout_combine_b_kernel<<<grid_for(len, 256, sms * 8), 256, 0, ST>>>(g, w, maps_len, ns, out);
    return (int)cudaGetLastError();
}

}  // namespace opm
