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
// The ranks meet three times per read-out (map ranges; "my part is binned"; peak / total).  Each meeting is one small kernel:
// a rank stores its block into every peer's mailbox, raises its flag there (release, system scope) and spins on its own flags
// (acquire) -- no NCCL launch, no host.  Mailboxes and flags alternate between two parities: a rank can be at most one meeting
// ahead of another.  Every spin has a time limit and ends in an error status instead of a hung GPU.
#include <cuda_runtime.h>

#include <cfloat>

#include "opm_kernels.h"
#include "opm_peer.h"

namespace opm {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long peer_dkey(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double peer_dkey_inv(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// One meeting of the ranks: all-gather of a block of n <= kPeerBlock doubles, which is also a barrier.
//   mode 0: out[r * n + k] = block of rank r
//   mode 1 (map ranges, blocks of 4 as (-l, r, t, -b)): the same, and out2[0..4) = element-wise max over the ranks
//   mode 2 (peak key | total): out2[0] = max over ranks of the peak, out2[1] = sum of the totals in rank order
__global__ void __launch_bounds__(kPeerThreads) peer_meet_kernel(PeerView V, unsigned long long seq, const double* __restrict__ block, int n,
                                                                 double* __restrict__ out, int mode, double* __restrict__ out2,
                                                                 unsigned* status, unsigned long long timeout_ns) {
    const int t = threadIdx.x, par = (int)(seq & 1ull);
    PeerCtrl* mine = V.ctrl[V.rank];
    for (int i = t; i < V.world * kPeerBlock; i += kPeerThreads) {
        const int peer = i / kPeerBlock, k = i % kPeerBlock;
        if (k < n) *(volatile double*)&V.ctrl[peer]->mail[par][V.rank][k] = block[k];
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

// zero the columns of this rank's window in its own part
__global__ void __launch_bounds__(256) peer_zero_window_kernel(double* __restrict__ part, unsigned long long nx, unsigned long long ny,
                                                               const double* __restrict__ own, const double* __restrict__ range) {
    long long c0, c1;
    peer_window(own, range, nx, c0, c1);
    if (c1 < c0) return;
    const unsigned long long first = (unsigned long long)c0 * ny, cnt = (unsigned long long)(c1 - c0 + 1) * ny;
    double* p = part + first;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt; i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = 0.0;
}

// slice owner: sum of the parts whose windows contain the column, in rank order; peak / total of the slice (fluence_data.rs:45-68,
// 130-141).  Work item = (column of the slice, block of kRows rows).  Peer parts are read with ld.volatile: peer data is cached
// in the L1 only, and the same addresses held other values one read-out ago.
constexpr int kSliceRows = 1024;
__global__ void __launch_bounds__(256) peer_reduce_slice_kernel(PeerView V, unsigned long long nx, unsigned long long ny,
                                                                const double* __restrict__ all_ranges, const double* __restrict__ range,
                                                                double* __restrict__ final_map, PeakAccum* acc) {
    __shared__ long long w0[kMaxPeers], w1[kMaxPeers];
    if (threadIdx.x < V.world) peer_window(all_ranges + 4 * threadIdx.x, range, nx, w0[threadIdx.x], w1[threadIdx.x]);
    __syncthreads();
    const unsigned long long s0 = (unsigned long long)V.rank * nx / V.world, s1 = (unsigned long long)(V.rank + 1) * nx / V.world;
    const unsigned long long row_blocks = (ny + kSliceRows - 1) / kSliceRows;
    const unsigned long long items = (s1 - s0) * row_blocks;
    const double dx = (range[1] - (-range[0])) / (double)nx, dy = (range[2] - (-range[3])) / (double)ny;
    const double area = dx * dy;
    double peak = -INFINITY, tot = 0.0;
    for (unsigned long long it = blockIdx.x; it < items; it += gridDim.x) {
        const unsigned long long col = s0 + it / row_blocks, rb = it % row_blocks;
        for (unsigned long long row = rb * kSliceRows + threadIdx.x; row < ny && row < (rb + 1) * kSliceRows; row += blockDim.x) {
            const unsigned long long idx = col * ny + row;
            double v = 0.0;
            for (int r = 0; r < V.world; ++r)
                if ((long long)col >= w0[r] && (long long)col <= w1[r]) v += *(volatile const double*)(V.part[r] + idx);
            final_map[idx] = v;
            if (isfinite(v)) peak = fmax(peak, v);
            if (!isnan(v)) tot += area * v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { peak = fmax(peak, __shfl_xor_sync(0xffffffffu, peak, o)); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    __shared__ double sp[8], st[8];
    if ((threadIdx.x & 31) == 0) { sp[threadIdx.x >> 5] = peak; st[threadIdx.x >> 5] = tot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { peak = fmax(peak, sp[k]); tot += st[k]; }
        atomicMax(&acc->peak_key, peer_dkey(peak));
        atomicAdd(&acc->total, tot);
    }
}
__global__ void peer_peak_block_kernel(const PeakAccum* acc, double* block) { block[0] = peer_dkey_inv(acc->peak_key); block[1] = acc->total; }
__global__ void peer_peak_init_kernel(PeakAccum* acc) { acc->peak_key = peer_dkey(-INFINITY); acc->total = 0.0; }

#define ST ((cudaStream_t)stream)
int k_peer_meet(const PeerView& V, unsigned long long seq, const double* block, int n, double* out, int mode, double* out2, unsigned* status,
                unsigned long long timeout_ns, void* stream) {
    peer_meet_kernel<<<1, kPeerThreads, 0, ST>>>(V, seq, block, n, out, mode, out2, status, timeout_ns);
    return (int)cudaGetLastError();
}
int k_peer_zero_window(double* part, uint64_t nx, uint64_t ny, const double* own, const double* range, int sms, void* stream) {
    peer_zero_window_kernel<<<sms * 4, 256, 0, ST>>>(part, nx, ny, own, range);
    return (int)cudaGetLastError();
}
int k_peer_reduce_slice(const PeerView& V, uint64_t nx, uint64_t ny, const double* all_ranges, const double* range, double* final_map,
                        PeakAccum* acc, double* peak_block, int sms, void* stream) {
    peer_peak_init_kernel<<<1, 1, 0, ST>>>(acc);
    peer_reduce_slice_kernel<<<sms * 8, 256, 0, ST>>>(V, nx, ny, all_ranges, range, final_map, acc);
    peer_peak_block_kernel<<<1, 1, 0, ST>>>(acc, peak_block);
    return (int)cudaGetLastError();
}

}  // namespace opm
