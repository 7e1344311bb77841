// opm_peer.h -- device-side layout of the peer-memory exchange (opm_peer.cu) shared with the C-ABI host code.
#pragma once
#include "opm_internal.h"
#include "opm_kernels.h"

namespace opm {

constexpr int kMaxPeers = 16;     // ranks of one box
constexpr int kPeerBlock = 32;    // doubles per rank and meeting
constexpr int kPeerThreads = 256;
constexpr int kPeerReduceCtasPerSm = 4;  // grid of the slice reduce; every CTA owns a (peak, total) slot

// head of every rank's exchange segment (cudaMalloc'ed, exported with cudaIpcGetMemHandle): flags and mailboxes, two parities
struct PeerCtrl {
    unsigned long long flag[2][kMaxPeers];          // flag[p][r]: sequence number of rank r's latest meeting of parity p
    double mail[2][kMaxPeers][kPeerBlock];          // mail[p][r]: the block rank r brought to that meeting
};
constexpr size_t kPeerCtrlBytes = (sizeof(PeerCtrl) + 255) / 256 * 256;

// the segments of all ranks as mapped into THIS process
struct PeerView {
    PeerCtrl* ctrl[kMaxPeers];
    double* part[kMaxPeers];  // the map-sized buffer behind the control block
    int rank, world;
};

// MEETING 1 carried by the kernel that zero-fills the rank's window (peer_zero_window_kernel): the first CTA turns the bounding
// box the trace kernel kept (acc) into this rank's range and brings it to every peer; every CTA waits for the ranges of all
// ranks and takes the element-wise maximum as the common range.  enabled == 0: the ranges are already in memory.
struct PeerHook {
    int enabled;
    PeerView V;
    unsigned long long seq, timeout_ns;
    const BBoxAccum* acc;
    double* own_out;   // [4]        this rank's range
    double* all_out;   // [world][4] every rank's range (windows of the slice reduce)
    double* range_out; // [4]        the common range
    unsigned* status;
    unsigned long long* stamps;  // diagnostic (may be NULL): globaltimer at [0] kernel start, [1] range posted, [2] all ranges here (first CTA)
};

#ifdef __CUDACC__
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
// this rank's range (-l, r, t, -b) element t from the accumulator of ordered keys (bbox_finalize_kernel)
__device__ __forceinline__ double peer_own_range(const BBoxAccum* acc, int t) {
    if (acc->count == 0) return -INFINITY;
    const double d = peer_dkey_inv(acc->key[t]);
    return (t == 0 || t == 3) ? -d : d;
}
// waits until every rank has raised flag[par][.] of this rank's control block to seq (threads 0 .. world-1 of the CTA)
__device__ __forceinline__ void peer_wait_all(const PeerView& V, int par, unsigned long long seq, unsigned* status, unsigned long long timeout_ns) {
    if ((int)threadIdx.x < V.world) {
        const PeerCtrl* mine = V.ctrl[V.rank];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[par][threadIdx.x]) < seq) {
            if (global_timer_ns() - t0 > timeout_ns) { atomicOr(status, 1u << OPM_ERR_PEER_TIMEOUT); break; }
            __nanosleep(64);
        }
    }
}
// the head of the zero-fill kernel when it carries meeting 1; s_range: 4 doubles of shared memory that receive the common range
__device__ __forceinline__ void peer_hook_meet(const PeerHook& K, double* s_range) {
    const PeerView& V = K.V;
    const int t = threadIdx.x, par = (int)(K.seq & 1ull);
    PeerCtrl* mine = V.ctrl[V.rank];
    if (K.stamps && blockIdx.x == 0 && t == 0) K.stamps[0] = global_timer_ns();
    if (blockIdx.x == 0) {
        if (t < 4) { const double v = peer_own_range(K.acc, t); K.own_out[t] = v; s_range[t] = v; }
        __syncthreads();
        if (t < V.world) {  // mail and flag of one peer by ONE thread: the release orders this thread's own stores before the flag
            volatile double* m = V.ctrl[t]->mail[par][V.rank];
            m[0] = s_range[0]; m[1] = s_range[1]; m[2] = s_range[2]; m[3] = s_range[3];
            st_release_sys(&V.ctrl[t]->flag[par][V.rank], K.seq);
        }
        if (K.stamps && t == 0) K.stamps[1] = global_timer_ns();
    }
    peer_wait_all(V, par, K.seq, K.status, K.timeout_ns);
    __syncthreads();
    if (K.stamps && blockIdx.x == 0 && t == 0) K.stamps[2] = global_timer_ns();
    if (t < 4) {
        double m = *(volatile const double*)&mine->mail[par][0][t];
        for (int r = 1; r < V.world; ++r) m = fmax(m, *(volatile const double*)&mine->mail[par][r][t]);
        s_range[t] = m;
        if (blockIdx.x == 0) K.range_out[t] = m;
    }
    if (blockIdx.x == 0)
        for (int i = t; i < V.world * 4; i += blockDim.x) K.all_out[i] = *(volatile const double*)&mine->mail[par][i / 4][i % 4];
    __syncthreads();
}
#endif

int k_peer_meet(const PeerView& V, unsigned long long seq, const double* block, int n, double* out, int mode, double* out2, unsigned* status,
                unsigned long long timeout_ns, const BBoxAccum* acc, double* own_out, void* stream);
int k_peer_reduce_slice(const PeerView& V, uint64_t nx, uint64_t ny, const double* all_ranges, const double* range, double* final_map,
                        double2* cta_peaks, unsigned* done, unsigned long long seq2, double* peak_total_out, double* range_out, unsigned* status,
                        unsigned long long timeout_ns, unsigned long long* stamps, const unsigned* bin_status, double* host_report, int sms,
                        void* stream);
int k_peer_zero_window(double* part, uint64_t nx, uint64_t ny, const double* own, const double* range, const PeerHook& hook, int sms,
                       void* stream);

}  // namespace opm
