// opm_peer.h -- device-side layout of the peer-memory exchange (opm_peer.cu) shared with the C-ABI host code.
#pragma once
#include "opm_internal.h"
#include "opm_kernels.h"

namespace opm {

constexpr int kMaxPeers = 16;     // ranks of one box
constexpr int kPeerBlock = 32;    // doubles per rank and meeting
constexpr int kPeerThreads = 256;

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

int k_peer_meet(const PeerView& V, unsigned long long seq, const double* block, int n, double* out, int mode, double* out2, unsigned* status,
                unsigned long long timeout_ns, void* stream);
int k_peer_zero_window(double* part, uint64_t nx, uint64_t ny, const double* own, const double* range, int sms, void* stream);
int k_peer_reduce_slice(const PeerView& V, uint64_t nx, uint64_t ny, const double* all_ranges, const double* range, double* final_map,
                        PeakAccum* acc, double* peak_block, int sms, void* stream);

}  // namespace opm
