"""Multi-GPU layer of the ray path: rays shard by source index, one process per GPU, and the only exchange is the
all-reduce of detector results.

    rank g owns source positions [g*N/G, (g+1)*N/G)        -- whole wavelength groups stay on one GPU (rays.rs:284-289)
    fluence map   : all-reduce(min/max) of the data-dependent map range (rays_hit_map.rs:522-562), every rank bins its
                    hits into the SAME grid, all-reduce(sum) of the maps (hit_map/mod.rs:363-379 merges hit maps the same way)
    spot diagram  : all-reduce of the phase-1 sums, then of the phase-2 radius accumulators (rays.rs:599-701)
    tallies       : all-reduce(sum) of per-bounce-order counts (int64) and energies (f64)

Everything here is host logic over `torch.distributed` (NCCL on the GPUs, gloo in the CPU tests); the per-ray work is
done by the caller's kernels.  Reductions of a handful of scalars travel as one small tensor.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when no process group is initialised"""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_strided(n_positions: int, rank: int, world_size: int) -> tuple[int, int, int]:
    """(first, count, stride) of the interleaved shard {rank, rank + G, rank + 2G, ...}: every rank sees the same mix of
    rays (load balance when apertures clip a contiguous part of the index range); rays are independent, so any
    partition of the source index set is a valid sharding (SURVEY 8e uses contiguous ranges)"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return rank, (n_positions - rank + world_size - 1) // world_size, world_size


def shard_range(n_positions: int, rank: int, world_size: int) -> tuple[int, int]:
    """contiguous index range [g*N/G, (g+1)*N/G) of the source POSITIONS owned by `rank`"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return (rank * n_positions) // world_size, ((rank + 1) * n_positions) // world_size


def _reduce(values, op, device, dtype=torch.float64):
    t = torch.tensor(list(values), dtype=dtype, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=op)
    return t.tolist()


def global_map_range(lrtb, device="cpu"):
    """(left, right, top, bottom) over all ranks from each rank's own bounding box; a rank without hits passes None.
    Raises like the reference (`could not calculate bounding box`) when no rank has a hit."""
    inf = math.inf
    l, r, t, b = lrtb if lrtb is not None else (inf, -inf, -inf, inf)
    nl, r, t, nb = _reduce([-l, r, t, -b], dist.ReduceOp.MAX, device)
    if not (math.isfinite(nl) and math.isfinite(r) and math.isfinite(t) and math.isfinite(nb)):
        raise ValueError("could not calculate bounding box")
    return (-nl, r, t, -nb)


def allreduce_map(fluence_map: torch.Tensor) -> torch.Tensor:
    """sum of the per-rank maps (all binned on the same global grid), in place"""
    if world()[1] > 1:
        dist.all_reduce(fluence_map, op=dist.ReduceOp.SUM)
    return fluence_map


def allreduce_tally(counts, energies, device="cpu"):
    """per-bounce-order valid-ray counts (exact, int64) and energies (f64) over all ranks"""
    c = _reduce([int(x) for x in counts], dist.ReduceOp.SUM, device, torch.int64)
    e = _reduce([float(x) for x in energies], dist.ReduceOp.SUM, device)
    return c, e


def reduce_spot_sums(p, device="cpu"):
    """phase 1 -> global sums (fields of an OpmSpotPartial-like object are updated in place)"""
    sums = _reduce(list(p.sum), dist.ReduceOp.SUM, device)
    n = _reduce([int(p.n_valid), int(p.n_total)], dist.ReduceOp.SUM, device, torch.int64)
    # bounce level = bounce count of the first valid ray (rays.rs:183-194).  Ray ids are GLOBAL source indices, so the first
    # valid ray of the whole bundle is the one with the smallest id on any rank, whatever the partition (contiguous or
    # interleaved); ties go to the lowest rank.  Same rule as the device path (opm_readout_combine_ranges_sums).
    rank = world()[0]
    big = (1 << 62)
    has = int(p.n_valid) > 0 and int(p.first_key) != 0xFFFFFFFFFFFFFFFF
    my_id = (int(p.first_key) >> 32) if has else big
    min_id = -_reduce([-my_id], dist.ReduceOp.MAX, device, torch.int64)[0]
    cand = ((rank << 32) | (int(p.first_key) & 0xFFFFFFFF)) if (has and my_id == min_id) else big
    cand = -_reduce([-cand], dist.ReduceOp.MAX, device, torch.int64)[0]
    for j in range(7):
        p.sum[j] = sums[j]
    p.n_valid, p.n_total = n
    p.first_key = 0xFFFFFFFFFFFFFFFF if min_id >= big else ((min_id << 32) | (cand & 0xFFFFFFFF))
    return p


def reduce_spot_radii(p, device="cpu"):
    """phase 2 -> global radius accumulators"""
    p.max_d2_mm2 = _reduce([p.max_d2_mm2], dist.ReduceOp.MAX, device)[0]
    p.sum_d2_mm2, p.sum_ed2 = _reduce([p.sum_d2_mm2, p.sum_ed2], dist.ReduceOp.SUM, device)
    return p


# ---------------------------------------------------------------------------------------------------------------
# device-chained detector read-out (GPU): every phase stays on the device, the ranks meet in three small collectives and
# the host synchronises once, when it reads the reports
# ---------------------------------------------------------------------------------------------------------------
def combine_ranges_sums_reference(gA: torch.Tensor, nf: int, ns: int) -> torch.Tensor:
    """what opm_readout_combine_ranges_sums computes, in torch ops (CPU or GPU tensors): the specification the kernel is
    tested against and the form the world-size-2 gloo test runs"""
    out = torch.empty(gA.shape[1], dtype=gA.dtype, device=gA.device)
    if nf:
        out[: 4 * nf] = gA[:, : 4 * nf].max(dim=0).values
    for k in range(ns):
        o = 4 * nf + 11 * k
        out[o: o + 9] = gA[:, o: o + 9].sum(dim=0)
        # bounce level = bounce count of the first valid ray (rays.rs:183-194): smallest id, ties to the lowest rank
        ids = gA[:, o + 10]
        first = int(torch.argmin(torch.where(ids < 0, torch.full_like(ids, math.inf), ids)))
        out[o + 9: o + 11] = gA[first, o + 9: o + 11]
    return out


def combine_maps_radii_reference(gB: torch.Tensor, maps_len: int, ns: int) -> torch.Tensor:
    """what opm_readout_combine_maps_radii computes, in torch ops"""
    out = gB.sum(dim=0)
    if ns:
        out[maps_len::3] = gB[:, maps_len::3].max(dim=0).values
    return out


class PeerGroup:
    """This rank's side of the NVLink peer-memory exchange of a large fluence map (include/opossum_b200.h `OpmPeerGroup`,
    opossum_b200/csrc/opm_peer.cu): allocates the rank's segment in the library, exchanges the CUDA IPC handles of all ranks
    through `torch.distributed` (a host-side all_gather_object, once) and maps the peers' segments."""

    def __init__(self, ctx, map_bins: int):
        import ctypes as C

        from . import _ffi as F
        from .api import _check
        self.ctx, self.map_bins = ctx, int(map_bins)
        self.rank, self.world = world()
        h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _check(F.lib().opm_peer_group_create(ctx.h, self.rank, self.world, self.map_bins, C.byref(h), handle))
        self.h = h
        ctx._adopt(self)
        self._close_order = 0  # before the context's bundles
        if self.world > 1:
            mine = bytes(handle)
            every = [None] * self.world
            dist.all_gather_object(every, mine)
            blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(every))
            _check(F.lib().opm_peer_group_connect(self.h, blob))
        else:
            _check(F.lib().opm_peer_group_connect(self.h, None))

    def allgather(self, block: torch.Tensor, out: torch.Tensor):
        """one meeting: `block` (n <= 32 float64 on the device) of every rank -> out (world x n); also a barrier of the ranks"""
        import ctypes as C

        from . import _ffi as F
        from .api import _check
        _check(F.lib().opm_peer_allgather_async(self.h, C.c_void_p(block.data_ptr()), block.numel(), C.c_void_p(out.data_ptr())))

    def close(self):
        from . import _ffi as F
        if getattr(self, "h", None):
            F.lib().opm_peer_group_destroy(self.h)
            self.h = None


class DetectorReadout:
    """Reports of the detectors of a scene whose rays are sharded over the ranks: Binning fluence maps
    (`FluenceDetector::node_report`, nodes/fluence_detector/mod.rs:91-135) and spot-diagram statistics
    (`SpotDiagram::node_report`, nodes/spot_diagram.rs:101-163).

        phase A  per detector: bounding box of its hits -> range (-l, r, t, -b); spot sums                 [local, async]
        gather 1 all_gather of (ranges | sums), combined in rank order by one kernel                          [1 collective]
        phase B  per detector: Binning into the common grid; spot radii about the global centroids           [local, async]
        gather 2 all_gather of (maps | radius accumulators), summed in rank order by one kernel               [1 collective]
                 (maps too large to gather -- more than FUSED_LIMIT doubles over all ranks -- are all-reduced instead)
        phase C  peak / total of every (global) map                                                          [local, async]
        read     one device->host copy of the scalars (and the maps if asked), one synchronisation

    `run()` does all of it; `launch(slot)` / `collect(slot)` split it at the synchronisation so that a caller that analyses
    one scene after another (a sweep, a tolerancing run) can queue the next trace before it reads the previous reports:
    two sets of pinned host buffers, one CUDA event each, the device buffers are reused in stream order.
    """
    FUSED_LIMIT = 8 * 1024 * 1024  # doubles in the gathered (maps | radii) buffer

    def __init__(self, ctx, scene, fluence_nodes, spot_nodes, nx: int, ny: int, device, peer=None, solo=False):
        import ctypes as C

        from . import _ffi as F
        self.C, self.F = C, F
        self.ctx, self.scene, self.nx, self.ny = ctx, scene, nx, ny
        self.fl, self.sp = list(fluence_nodes), list(spot_nodes)
        nf, ns = len(self.fl), len(self.sp)
        self.nA = 4 * nf + 11 * ns                       # ranges | spot sums
        self.nB = 3 * ns                                 # spot radii
        self.maps_len = max(nf, 1) * nx * ny
        self.buf = torch.zeros(self.nA + 2 * nf, dtype=torch.float64, device=device)        # A | peak, total per map
        self.mb = torch.zeros(self.maps_len + self.nB, dtype=torch.float64, device=device)  # maps | B
        self.maps = self.mb[: self.maps_len].view(max(nf, 1), nx * ny)
        self.status = torch.zeros(max(nf, 1), dtype=torch.int32, device=device)
        self.host = [{"buf": torch.zeros(self.buf.numel(), dtype=torch.float64).pin_memory(),
                      "B": torch.zeros(max(self.nB, 1), dtype=torch.float64).pin_memory(),
                      "status": torch.zeros(self.status.numel(), dtype=torch.int32).pin_memory(),
                      "report": torch.zeros(8 * max(nf, 1), dtype=torch.float64).pin_memory(), "lean": False,
                      "maps": None, "event": None, "with_maps": False} for _ in range(2)]
        w = 1 if solo else world()[1]  # solo: this rank reads its own detectors out without meeting the others
        self.w = w
        self.fused = w > 1 and self.mb.numel() * w <= self.FUSED_LIMIT and peer is not True
        # maps too large to gather: column-sliced merge over NVLink peer memory (PeerGroup), one group per detector; the NCCL
        # all-reduce of the whole map remains as the fallback when the GPUs cannot map each other's memory
        self.peer = None
        if w > 1 and not self.fused and nf and peer is not False:
            try:
                self.peer = [PeerGroup(ctx, nx * ny) for _ in range(nf)]
            except Exception as e:  # noqa: BLE001  (no peer access, IPC not permitted in this container, ...)
                if peer is True:
                    raise
                import warnings
                warnings.warn(f"peer-memory read-out unavailable ({e}); all-reducing whole fluence maps over NCCL instead")
            flag = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32,
                                device=device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks or none
            if int(flag.item()) == 0:
                self.peer = None
        self.gA = torch.zeros((w, self.nA), dtype=torch.float64, device=device) if w > 1 else None
        self.gMB = torch.zeros((w, self.mb.numel()), dtype=torch.float64, device=device) if self.fused else None
        self.gB = torch.zeros((w, max(self.nB, 1)), dtype=torch.float64, device=device) if (w > 1 and not self.fused) else None

        ctx._adopt(self)       # closed with (and before) the context: the pinned report buffers were last used on ITS stream, and
        self._close_order = -1  # torch's host allocator records an event on that stream when it takes them back

    def close(self):
        if getattr(self, "host", None) is None:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for g in (self.peer or []):
            g.close()
        self.peer = None
        self.host = None
        self.buf = self.mb = self.maps = self.status = self.gA = self.gMB = self.gB = None

    def _ptr(self, t: torch.Tensor, offset: int = 0) -> int:
        return t.data_ptr() + 8 * offset

    def _hit_arrays(self, node):
        hm = self.scene.node_hitmaps(node, 0)
        return (self.C.c_void_p * len(hm))(*[h.h for h in hm]), len(hm)

    def run(self, download_maps: bool = False) -> dict:
        self.launch(download_maps, 0)
        return self.collect(0)

    def launch(self, download_maps: bool = False, slot: int = 0):
        """queue the whole read-out and its device->host copies on the current stream; nothing waits"""
        # the library's kernels run on the context's stream; the torch side of the chain (collectives, status reset, pinned
        # copies, the completion event) must be ordered with them: everything below is issued on that very stream
        with torch.cuda.stream(torch.cuda.ExternalStream(self.ctx.stream, device=self.buf.device)):
            self._launch(download_maps, slot)

    def _launch(self, download_maps: bool, slot: int):
        L, C, F = self.F.lib(), self.C, self.F
        from .api import _check
        w = self.w
        nf, ns, nx, ny = len(self.fl), len(self.sp), self.nx, self.ny
        hits = [self._hit_arrays(n) for n in self.fl]
        vp = C.c_void_p
        peer = self.peer is not None
        # ---- phase A
        if not peer:
            for j, (arr, n) in enumerate(hits):
                _check(L.opm_hitmaps_bounding_box_async(arr, n, -1, vp(self._ptr(self.buf, 4 * j))))
        for k, node in enumerate(self.sp):
            _check(L.opm_scene_node_spot_async(self.scene.h, node, 1, vp(self._ptr(self.buf, 4 * nf + 11 * k)), None))
        if w > 1 and self.nA and (ns or not peer):
            dist.all_gather_into_tensor(self.gA.view(-1), self.buf[: self.nA])
            _check(L.opm_readout_combine_ranges_sums(self.ctx.h, vp(self.gA.data_ptr()), w, nf, ns, vp(self.buf.data_ptr())))
        # ---- phase B
        H = self.host[slot]
        # peer path with fluence detectors only and no map download: the last kernel of the report writes range, peak, total and
        # status straight into this slot's pinned host block -- no status reset, no device->host copies in the stream
        H["lean"] = lean = bool(peer and not ns and not download_maps)
        if not lean:
            self.status.zero_()
        if peer:  # range, windowed binning, column-slice merge, peak / total: three flag meetings over NVLink, no NCCL
            for j, (arr, n) in enumerate(hits):
                _check(L.opm_peer_fluence_report_async(self.peer[j].h, arr, n, -1, nx, ny, vp(self._ptr(self.buf, 4 * j)),
                                                       vp(self._ptr(self.buf, self.nA + 2 * j)), vp(self.status.data_ptr() + 4 * j),
                                                       vp(H["report"].data_ptr() + 64 * j) if lean else None))
        for j, (arr, n) in enumerate(hits if not peer else []):
            _check(L.opm_fluence_binning_async(arr, n, -1, nx, ny, vp(self._ptr(self.buf, 4 * j)), 0,
                                               vp(self.maps[j].data_ptr()), vp(self.status.data_ptr() + 4 * j)))
        for k, node in enumerate(self.sp):
            _check(L.opm_scene_node_spot_async(self.scene.h, node, 2, vp(self._ptr(self.buf, 4 * nf + 11 * k)),
                                               vp(self._ptr(self.mb, self.maps_len + 3 * k))))
        if self.fused:
            dist.all_gather_into_tensor(self.gMB.view(-1), self.mb)
            _check(L.opm_readout_combine_maps_radii(self.ctx.h, vp(self.gMB.data_ptr()), w, self.maps_len, ns, vp(self.mb.data_ptr())))
        elif w > 1:
            if nf and not peer:
                dist.all_reduce(self.maps, op=dist.ReduceOp.SUM)
            if ns:
                dist.all_gather_into_tensor(self.gB.view(-1), self.mb[self.maps_len:])
                _check(L.opm_readout_combine_maps_radii(self.ctx.h, vp(self.gB.data_ptr()), w, 0, ns, vp(self._ptr(self.mb, self.maps_len))))
        # ---- phase C
        for j in range(nf if not peer else 0):
            _check(L.opm_fluence_peak_total_async(self.ctx.h, vp(self.maps[j].data_ptr()), nx, ny, vp(self._ptr(self.buf, 4 * j)),
                                                  vp(self._ptr(self.buf, self.nA + 2 * j))))
        # ---- read
        if not lean:
            H["buf"].copy_(self.buf, non_blocking=True)
            if ns:
                H["B"].copy_(self.mb[self.maps_len:], non_blocking=True)
            H["status"].copy_(self.status, non_blocking=True)
        H["with_maps"] = bool(download_maps and nf)
        if H["with_maps"] and peer:  # every rank collects the column slices of all ranks (peer copies)
            for j in range(nf):
                _check(L.opm_peer_fluence_gather_map(self.peer[j].h, nx, ny, vp(self.maps[j].data_ptr())))
        if H["with_maps"]:
            if H["maps"] is None:
                H["maps"] = torch.zeros_like(self.maps, device="cpu").pin_memory()
            H["maps"].copy_(self.maps, non_blocking=True)
        if H["event"] is None:
            H["event"] = torch.cuda.Event()
        H["event"].record(torch.cuda.current_stream())

    def collect(self, slot: int = 0) -> dict:
        """wait for the read-out queued in `slot` and build the reports"""
        L, C, F = self.F.lib(), self.C, self.F
        nf, nx, ny = len(self.fl), self.nx, self.ny
        H = self.host[slot]
        H["event"].synchronize()
        h, download_maps = H["buf"], H["with_maps"]
        if download_maps and self.peer is not None:
            dist.barrier()  # no rank starts its next read-out while a peer still copies its slice
        out = {}
        for j, node in enumerate(self.fl):
            if H["lean"]:
                rep = H["report"][8 * j: 8 * j + 8].tolist()
                r, pk, st = rep[0:4], rep[4:6], int(rep[6])
                if st:
                    self.status.zero_()  # sticky on this path: reported once
            else:
                r = h[4 * j: 4 * j + 4].tolist()
                pk = h[self.nA + 2 * j: self.nA + 2 * j + 2].tolist()
                st = int(H["status"][j])
            lrtb = (-r[0], r[1], r[2], -r[3])
            if not all(math.isfinite(v) for v in lrtb):
                raise ValueError("could not calculate bounding box")
            if st:
                raise RuntimeError(f"fluence binning raised status bits {st:#x}")
            out[node] = {"lrtb": lrtb, "peak": pk[0], "total_energy": pk[1],
                         "map": H["maps"][j].numpy().reshape(nx, ny).T.copy() if download_maps else None}
        for k, node in enumerate(self.sp):
            s = h[4 * nf + 11 * k: 4 * nf + 11 * k + 11].tolist()
            rr = H["B"][3 * k: 3 * k + 3].tolist()
            p = F.OpmSpotPartial()
            for i in range(7):
                p.sum[i] = s[i]
            p.n_valid, p.n_total = int(s[7]), int(s[8])
            p.first_key = 0xFFFFFFFFFFFFFFFF if s[9] < 0 else ((int(s[10]) << 32) | int(s[9]))
            p.max_d2_mm2, p.sum_d2_mm2, p.sum_ed2 = rr
            st = F.OpmSpotStats()
            L.opm_spot_stats_from_partial(C.byref(p), C.byref(st))
            out[node] = st
        return out


# ---------------------------------------------------------------------------------------------------------------
# drivers over the product API (GPU): used by bench.py and by applications that shard a scene over the GPUs of a box
# ---------------------------------------------------------------------------------------------------------------
def fluence_map(ctx, hit_maps, nx: int, ny: int, map_tensor: torch.Tensor):
    """HitMap::calc_fluence_map(Binning) of hit maps sharded over the ranks, result replicated on every rank.
    `map_tensor`: nx*ny float64 CUDA tensor that receives the (all-reduced) map.  Returns FluenceData (map on the device)."""
    from . import api
    dev = map_tensor.device
    try:
        lrtb, _ = api.hitmaps_bounding_box(hit_maps)
    except api.OpossumError as e:
        if e.code != api.F.OPM_ERR_EMPTY_HITMAP or world()[1] == 1:
            raise
        lrtb = None
    if world()[1] > 1:
        lrtb = global_map_range(lrtb, dev)
    fd = api.fluence_binning(ctx, hit_maps, nx, ny, lrtb=lrtb, map_dev=map_tensor.data_ptr(), download=False)
    if world()[1] > 1:
        allreduce_map(map_tensor)
        fd = api.fluence_peak_total(ctx, map_tensor.data_ptr(), nx, ny, lrtb)
    return fd


def spot_stats(scene, node: int, device):
    """SpotDiagram::node_report of a bundle sharded over the ranks (exact two-phase reduction)"""
    from . import _ffi as F
    import ctypes as C
    p = scene.node_spot_partial(node)
    if world()[1] > 1:
        reduce_spot_sums(p, device)
    scene.node_spot_partial(node, reduced=p)
    if world()[1] > 1:
        reduce_spot_radii(p, device)
    out = F.OpmSpotStats()
    F.lib().opm_spot_stats_from_partial(C.byref(p), C.byref(out))
    return out
