// opm_peer.cu -- the exchange step of a sharded detector read-out over NVLink peer memory, sm_100a.
//
// Rays shard by source index, one process per GPU; the only exchange of the path is the read-out of a detector whose hits are
// spread over the ranks (hit_map/mod.rs:363-379 merges hit maps, rays_hit_map.rs:522-562 gives the common range).  For a
// large fluence map an all-reduce moves the whole map through every GPU although (a) each rank's hits cover a part of the
// map only -- a contiguous index range of an ordered source lands in a band of map columns -- and (b) nobody needs the sum on
// every GPU: the report is a peak, a total and ONE copy of the map.  Here every rank
//   1. bins its hits into its own full-size buffer `part`, touching (and zeroing) only the columns of its WINDOW -- the map
//      columns its own bounding box spans inside the common range, known on the device, never on the host;
//   2. owns a slice of the map columns and sums, in rank order, the parts of all ranks whose windows reach into its slice,
//      reading the peers' buffers directly (CUDA IPC mappings, loads over NVLink), peak and total of the slice in the same pass;
//   3. the per-slice (peak, total) pairs meet in a 16-byte exchange.
// The ranks meet three times per read-out (map ranges; "my part is binned"; peak / total).  A meeting: a rank stores its block
// into every peer's mailbox, raises its flag there (release, system scope) and spins on its own flags (acquire) -- no NCCL
// launch, no host.  The first meeting is one small kernel (which also turns the trace kernel's bounding-box accumulator into
// the rank's range), the other two are the head and the tail of the slice-reduce kernel: four launches per read-out.  Mailboxes and flags alternate between two parities: a rank can be at most one meeting
// ahead of another.  Every spin has a time limit and ends in an error status instead of a hung GPU.
#include <cuda_runtime.h>

#include <cfloat>

#include "opm_kernels.h"
#include "opm_peer.h"

namespace opm {

// One meeting of the ranks: all-gather of a block of n <= kPeerBlock doubles, which is also a barrier.
//   mode 0: out[r * n + k] = block of rank r
//   mode 1 (map ranges, blocks of 4 as (-l, r, t, -b)): the same, and out2[0..4) = element-wise max over the ranks
//   mode 2 (peak key | total): out2[0] = max over ranks of the peak, out2[1] = sum of the totals in rank order
//   acc != NULL (mode 1): the block is this rank's own range, taken from the bounding-box accumulator the trace kernel kept
//   (what bbox_finalize_kernel computes) and also stored to `own_out` -- no separate launch for it
__global__ void __launch_bounds__(kPeerThreads) peer_meet_kernel(PeerView V, unsigned long long seq, const double* __restrict__ block, int n,
                                                                 double* __restrict__ out, int mode, double* __restrict__ out2,
                                                                 unsigned* status, unsigned long long timeout_ns, const BBoxAccum* acc,
                                                                 double* __restrict__ own_out) {
    const int t = threadIdx.x, par = (int)(seq & 1ull);
    PeerCtrl* mine = V.ctrl[V.rank];
    __shared__ double sblock[kPeerBlock];
    if (t < n) {
        double v;
        if (acc) {
            v = peer_own_range(acc, t);
            own_out[t] = v;
        } else v = block[t];
        sblock[t] = v;
    }
    __syncthreads();
    for (int i = t; i < V.world * kPeerBlock; i += kPeerThreads) {
        const int peer = i / kPeerBlock, k = i % kPeerBlock;
        if (k < n) *(volatile double*)&V.ctrl[peer]->mail[par][V.rank][k] = sblock[k];
    }
    __threadfence_system();
    __syncthreads();
    if (t < V.world) st_release_sys(&V.ctrl[t]->flag[par][V.rank], seq);
    if (t < V.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[par][t]) < seq) {
            if (global_timer_ns() - t0 > timeout_ns) { atomicOr(status, 1u << OPM_ERR_PEER_TIMEOUT); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int i = t; i < V.world * n; i += kPeerThreads) out[i] = *(volatile const double*)&mine->mail[par][i / n][i % n];
    if (mode == 1 && t < 4) {
        double m = *(volatile const double*)&mine->mail[par][0][t];
        for (int r = 1; r < V.world; ++r) m = fmax(m, *(volatile const double*)&mine->mail[par][r][t]);
        out2[t] = m;
    }
    if (mode == 2 && t == 0) {
        double pk = -INFINITY, tot = 0.0;
        for (int r = 0; r < V.world; ++r) {
            pk = fmax(pk, *(volatile const double*)&mine->mail[par][r][0]);
            tot += *(volatile const double*)&mine->mail[par][r][1];
        }
        out2[0] = pk; out2[1] = tot;
    }
}

// first and last map column a rank's own hits can touch, from its own bounding box inside the common range.  The column of a hit is
// trunc((x - left) / width_step) (bin_taps, opm_kernels.cu) and a hit also feeds the column to its right; one more column on
// either side covers a last-bit difference between this quotient and the splat's reciprocal form.
__device__ __forceinline__ void peer_window(const double* own, const double* range, unsigned long long nx, long long& c0, long long& c1) {
    const double l = -range[0], r = range[1];
    const double width_step = (r - l) / (double)(nx - 1);
    const double ol = -own[0], orr = own[1];
    if (!(isfinite(ol) && isfinite(orr)) || !(isfinite(l) && isfinite(r))) { c0 = 0; c1 = -1; return; }  // no hits: empty window
    double a = floor((ol - l) / width_step) - 1.0, b = floor((orr - l) / width_step) + 2.0;
    if (!(a >= 0.0)) a = 0.0;                       // also NaN (a one-column range: width_step = 0)
    if (!(b <= (double)(nx - 1))) b = (double)(nx - 1);
    c0 = (long long)a; c1 = (long long)b;
}

// zero the columns of this rank's window in its own part.  More than a fill: the window's lines are in the L2 when the splat's
// reductions arrive (measured: without it, after the trace has streamed 2.6 GB through the L2, the same splat takes 130 instead of
// 95 us).  With hook.enabled the kernel carries MEETING 1 at its head (peer_hook_meet): one launch less, and no one-CTA kernel
// on the critical path.
__global__ void __launch_bounds__(256) peer_zero_window_kernel(double* __restrict__ part, unsigned long long nx, unsigned long long ny,
                                                               const double* __restrict__ own, const double* __restrict__ range,
                                                               const PeerHook hook) {
    __shared__ double s_range[4], s_own[4];
    if (hook.enabled) {
        peer_hook_meet(hook, s_range);
        if (threadIdx.x < 4) s_own[threadIdx.x] = peer_own_range(hook.acc, threadIdx.x);
        __syncthreads();
        own = s_own; range = s_range;
    }
    long long c0, c1;
    peer_window(own, range, nx, c0, c1);
    if (c1 < c0) return;
    const unsigned long long first = (unsigned long long)c0 * ny, cnt = (unsigned long long)(c1 - c0 + 1) * ny;
    double* p = part + first;
    if (((first | cnt) & 1ull) == 0) {
        double2* p2 = reinterpret_cast<double2*>(p);
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt / 2; i += (unsigned long long)gridDim.x * blockDim.x)
            p2[i] = make_double2(0.0, 0.0);
    } else {
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt; i += (unsigned long long)gridDim.x * blockDim.x)
            p[i] = 0.0;
    }
}

// slice owner: sum of the parts whose windows contain the column, in rank order; peak / total of the slice (fluence_data.rs:45-68,
// 130-141).  Work item = (column of the slice, block of kSliceRows rows); a thread takes pairs of rows (16-byte accesses).
// Reads of the parts are system-scope relaxed loads (a peer's part is never served from a stale L1 line), issued for up to four
// ranks before the first sum needs a value -- the loads of the edge columns cross NVLink.  Every CTA leaves its (peak, total) in
// its own slot; the CTA that finishes last folds the slots in CTA order (no same-address atomics, a run-to-run stable total).
constexpr int kSliceRows = 1024;
__device__ __forceinline__ double2 ld_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// The kernel carries two of the three meetings itself.  MEETING 2 ("my part is binned") at its start: the first CTA raises this
// rank's flag at every peer -- the splat before it in the stream has completed, the system-scope fence orders its stores before the
// flag -- and every CTA waits for the flags of all ranks before it reads a peer's part.  MEETING 3 (peak | total of the slices) at
// its end: the CTA that finishes last brings the slice's (peak, total) to every peer, waits for theirs and writes the report's
// peak and total; it is also what keeps a fast rank from binning into its part for the next read-out while a peer still sums it.
__global__ void __launch_bounds__(256) peer_reduce_slice_kernel(PeerView V, unsigned long long nx, unsigned long long ny,
                                                                const double* __restrict__ all_ranges, const double* __restrict__ range,
                                                                double* __restrict__ final_map, double2* __restrict__ cta_peaks, unsigned* done,
                                                                unsigned long long seq2, double* __restrict__ peak_total_out,
                                                                double* __restrict__ range_out, unsigned* status, unsigned long long timeout_ns,
                                                                unsigned long long* stamps, const unsigned* bin_status,
                                                                double* __restrict__ host_report) {
    // stamps (diagnostic, may be NULL): globaltimer at [3] kernel start, [4] all parts binned (first CTA), [5] last CTA done,
    // [6] all slices' peaks here, [7] first CTA done with its items
    if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[3] = global_timer_ns();
    __shared__ long long w0[kMaxPeers], w1[kMaxPeers];
    __shared__ unsigned s_last;
    PeerCtrl* mine = V.ctrl[V.rank];
    const int par2 = (int)(seq2 & 1ull), par3 = par2 ^ 1;
    const unsigned long long seq3 = seq2 + 1;
    if (threadIdx.x < V.world) peer_window(all_ranges + 4 * threadIdx.x, range, nx, w0[threadIdx.x], w1[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x < V.world) st_release_sys(&V.ctrl[threadIdx.x]->flag[par2][V.rank], seq2);
    peer_wait_all(V, par2, seq2, status, timeout_ns);
    __syncthreads();
    if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[4] = global_timer_ns();
    const unsigned long long s0 = (unsigned long long)V.rank * nx / V.world, s1 = (unsigned long long)(V.rank + 1) * nx / V.world;
    const unsigned long long row_blocks = (ny + kSliceRows - 1) / kSliceRows;
    const unsigned long long items = (s1 - s0) * row_blocks;
    const double dx = (range[1] - (-range[0])) / (double)nx, dy = (range[2] - (-range[3])) / (double)ny;
    const double area = dx * dy;
    const bool pairs = (ny & 1ull) == 0;  // every column starts 16-byte aligned
    double peak = -INFINITY, tot = 0.0;
    auto tally = [&](double v) {
        if (isfinite(v)) peak = fmax(peak, v);
        if (!isnan(v)) tot += area * v;
    };
    for (unsigned long long it = blockIdx.x; it < items; it += gridDim.x) {
        const unsigned long long col = s0 + it / row_blocks, rb = it % row_blocks;
        unsigned mask = 0;
        for (int r = 0; r < V.world; ++r)
            if ((long long)col >= w0[r] && (long long)col <= w1[r]) mask |= 1u << r;
        const unsigned long long row_end = (rb + 1) * kSliceRows < ny ? (rb + 1) * kSliceRows : ny;
        if (pairs) {
            constexpr int kPairs = kSliceRows / 2 / 256;  // pairs of rows per thread and item
            unsigned long long idx[kPairs];
            double2 v[kPairs];
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                const unsigned long long row = rb * kSliceRows + 2ull * (threadIdx.x + 256u * q);
                idx[q] = row < row_end ? col * ny + row : ~0ull;
                v[q] = make_double2(0.0, 0.0);
            }
            unsigned m = mask;
            while (m) {  // up to four ranks at a time: all their loads are in flight before the first sum
                int rk[4];
                double2 got[4][kPairs];
#pragma unroll
                for (int j = 0; j < 4; ++j) { rk[j] = m ? __ffs(m) - 1 : -1; m &= m - 1; }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < kPairs; ++q)
                        if (rk[j] >= 0 && idx[q] != ~0ull) got[j][q] = ld_sys_f64x2(V.part[rk[j]] + idx[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < kPairs; ++q)
                        if (rk[j] >= 0 && idx[q] != ~0ull) { v[q].x += got[j][q].x; v[q].y += got[j][q].y; }
            }
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
                if (idx[q] != ~0ull) {
                    *reinterpret_cast<double2*>(final_map + idx[q]) = v[q];
                    tally(v[q].x); tally(v[q].y);
                }
        } else {
            for (unsigned long long row = rb * kSliceRows + threadIdx.x; row < row_end; row += blockDim.x) {
                const unsigned long long idx = col * ny + row;
                double v = 0.0;
                for (int r = 0; r < V.world; ++r)
                    if (mask & (1u << r)) v += ld_sys_f64(V.part[r] + idx);
                final_map[idx] = v;
                tally(v);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { peak = fmax(peak, __shfl_xor_sync(0xffffffffu, peak, o)); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    __shared__ double sp[8], st[8];
    if ((threadIdx.x & 31) == 0) { sp[threadIdx.x >> 5] = peak; st[threadIdx.x >> 5] = tot; }
    if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[7] = global_timer_ns();
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { peak = fmax(peak, sp[k]); tot += st[k]; }
        cta_peaks[blockIdx.x] = make_double2(peak, tot);
        __threadfence();
        s_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- the last CTA of this rank: peak / total of the slice, then meeting 3
    __threadfence();
    if (stamps && threadIdx.x == 0) stamps[5] = global_timer_ns();
    peak = -INFINITY; tot = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
        const double2 c = __ldcg(&cta_peaks[k]);
        peak = fmax(peak, c.x); tot += c.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { peak = fmax(peak, __shfl_xor_sync(0xffffffffu, peak, o)); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sp[threadIdx.x >> 5] = peak; st[threadIdx.x >> 5] = tot; }
    __syncthreads();
    if (threadIdx.x < V.world) {
        double my_peak = sp[0], my_total = st[0];
        for (int k = 1; k < 8; ++k) { my_peak = fmax(my_peak, sp[k]); my_total += st[k]; }
        volatile double* m = V.ctrl[threadIdx.x]->mail[par3][V.rank];
        m[0] = my_peak; m[1] = my_total;
        st_release_sys(&V.ctrl[threadIdx.x]->flag[par3][V.rank], seq3);
    }
    peer_wait_all(V, par3, seq3, status, timeout_ns);
    __syncthreads();
    if (threadIdx.x == 0) {
        double pk = -INFINITY, total = 0.0;
        for (int r = 0; r < V.world; ++r) {
            pk = fmax(pk, *(volatile const double*)&mine->mail[par3][r][0]);
            total += *(volatile const double*)&mine->mail[par3][r][1];
        }
        peak_total_out[0] = pk; peak_total_out[1] = total;
        if (host_report) {  // the report straight into the caller's pinned host memory: no device->host copy in the stream
            volatile double* h = host_report;
            for (int k = 0; k < 4; ++k) h[k] = range[k];
            h[4] = pk; h[5] = total; h[6] = (double)(bin_status ? *bin_status : 0u);
            __threadfence_system();
            h[7] = (double)seq3;
        }
        if (stamps) stamps[6] = global_timer_ns();
        *done = 0u;  // for the next read-out (stream-ordered)
    }
    if (threadIdx.x < 4) range_out[threadIdx.x] = range[threadIdx.x];
}

#define ST ((cudaStream_t)stream)
int k_peer_meet(const PeerView& V, unsigned long long seq, const double* block, int n, double* out, int mode, double* out2, unsigned* status,
                unsigned long long timeout_ns, const BBoxAccum* acc, double* own_out, void* stream) {
    peer_meet_kernel<<<1, kPeerThreads, 0, ST>>>(V, seq, block, n, out, mode, out2, status, timeout_ns, acc, own_out);
    return (int)cudaGetLastError();
}
int k_peer_zero_window(double* part, uint64_t nx, uint64_t ny, const double* own, const double* range, const PeerHook& hook, int sms,
                       void* stream) {
    peer_zero_window_kernel<<<sms * 4, 256, 0, ST>>>(part, nx, ny, own, range, hook);
    return (int)cudaGetLastError();
}
int k_peer_reduce_slice(const PeerView& V, uint64_t nx, uint64_t ny, const double* all_ranges, const double* range, double* final_map,
                        double2* cta_peaks, unsigned* done, unsigned long long seq2, double* peak_total_out, double* range_out, unsigned* status,
                        unsigned long long timeout_ns, unsigned long long* stamps, const unsigned* bin_status, double* host_report, int sms,
                        void* stream) {
    peer_reduce_slice_kernel<<<sms * kPeerReduceCtasPerSm, 256, 0, ST>>>(V, nx, ny, all_ranges, range, final_map, cta_peaks, done, seq2, peak_total_out,
                                                                        range_out, status, timeout_ns, stamps, bin_status, host_report);
    return (int)cudaGetLastError();
}

}  // namespace opm
