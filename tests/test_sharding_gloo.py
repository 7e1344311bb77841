"""Multi-process (world_size 2, gloo, CPU) test of the sharding layer (opossum_b200/sharding.py):
rays shard by source index, each rank traces its shard (here with the CPU oracle standing in for the kernels), and the
detector results are combined with the module's all-reduces.  The combined fluence map, map range, tallies and
spot-diagram statistics must equal the single-process result over the whole source."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
from opossum_b200 import ffi as F
from opossum_b200 import sharding
from oracle import oracle as orc

N_RAYS = 6000
NX, NY = 64, 48


def _trace(src):
    osc = scenes.to_oracle(scenes.c1_single_lens())
    out = osc.raytrace(src)
    return osc, out


def _partial_sums(rays, iso) -> F.OpmSpotPartial:
    """CPU stand-in for opm_rays_spot_partial_sums (test infrastructure)"""
    loc = rays.copy()
    orc.rays_transform(loc, iso, inverse=True)
    v = loc[loc["valid"] == 1]
    p = F.OpmSpotPartial()
    s = [v["pos"][:, 0].sum(), v["pos"][:, 1].sum(), v["pos"][:, 2].sum(), (v["pos"][:, 0] * v["e"]).sum(),
         (v["pos"][:, 1] * v["e"]).sum(), (v["pos"][:, 2] * v["e"]).sum(), v["e"].sum()]
    for j in range(7):
        p.sum[j] = s[j]
    p.n_valid, p.n_total = len(v), len(loc)
    p.first_key = (int(v["id"][0]) << 32 | int(v["bounces"][0])) if len(v) else 0xFFFFFFFFFFFFFFFF
    return p


def _partial_radii(rays, iso, p):
    loc = rays.copy()
    orc.rays_transform(loc, iso, inverse=True)
    v = loc[loc["valid"] == 1]
    cx, cy = p.sum[0] / p.n_valid, p.sum[1] / p.n_valid
    ex, ey = p.sum[3] / p.sum[6], p.sum[4] / p.sum[6]
    d2 = (cx / 1e-3 - v["pos"][:, 0] / 1e-3) ** 2 + (cy / 1e-3 - v["pos"][:, 1] / 1e-3) ** 2
    p.max_d2_mm2 = float(d2.max()) if len(v) else 0.0
    p.sum_d2_mm2 = float(d2.sum())
    p.sum_ed2 = float((((ex - v["pos"][:, 0]) ** 2 + (ey - v["pos"][:, 1]) ** 2) * v["e"]).sum())


def _worker(rank, world_size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        src_all = scenes.source_to_oracle(scenes.c1_source(int(round(N_RAYS ** 0.5))))
        b, e = sharding.shard_range(len(src_all), rank, world_size)
        osc, out = _trace(src_all[b:e].copy())
        hits = osc.nodes[3].s[0].merged_hits()
        _, lrtb_local = orc.fluence_binning(hits, NX, NY)
        lrtb = sharding.global_map_range(lrtb_local)
        fmap, _ = orc.fluence_binning(hits, NX, NY, ranges=lrtb)
        t = torch.from_numpy(np.ascontiguousarray(fmap))
        sharding.allreduce_map(t)
        counts = [int(((out["valid"] == 1) & (out["bounces"] == k)).sum()) for k in range(3)]
        energy = [float(out["e"][(out["valid"] == 1) & (out["bounces"] == k)].sum()) for k in range(3)]
        counts, energy = sharding.allreduce_tally(counts, energy)
        iso = osc.nodes[2].s[0].eff_iso
        stored = osc.nodes[2].stored
        p = _partial_sums(stored, iso)
        sharding.reduce_spot_sums(p)
        _partial_radii(stored, iso, p)
        sharding.reduce_spot_radii(p)
        st = F.OpmSpotStats()
        F.lib().opm_spot_stats_from_partial(C.byref(p), C.byref(st))
        # the gathered-block form DetectorReadout uses on the GPU (one all-gather per phase, rows combined in rank order):
        # block A = [map range as (-l, r, t, -b) | 11 spot sums], block B = [map bins | 3 radius accumulators]
        pl = _partial_sums(stored, iso)
        first = (-1.0, -1.0) if pl.first_key == 0xFFFFFFFFFFFFFFFF else (float(pl.first_key & 0xFFFFFFFF), float(pl.first_key >> 32))
        a = torch.tensor([-lrtb_local[0], lrtb_local[1], lrtb_local[2], -lrtb_local[3]] + list(pl.sum) +
                         [float(pl.n_valid), float(pl.n_total), *first], dtype=torch.float64)
        ga = [torch.zeros_like(a) for _ in range(world_size)]
        dist.all_gather(ga, a)
        A = sharding.combine_ranges_sums_reference(torch.stack(ga), 1, 1)
        pg = F.OpmSpotPartial()
        for j in range(7):
            pg.sum[j] = float(A[4 + j])
        pg.n_valid, pg.n_total = int(A[11]), int(A[12])
        _partial_radii(stored, iso, pg)
        fmap_g, _ = orc.fluence_binning(hits, NX, NY, ranges=(-float(A[0]), float(A[1]), float(A[2]), -float(A[3])))
        b_blk = torch.cat([torch.from_numpy(np.ascontiguousarray(fmap_g)).reshape(-1),
                           torch.tensor([pg.max_d2_mm2, pg.sum_d2_mm2, pg.sum_ed2], dtype=torch.float64)])
        gb = [torch.zeros_like(b_blk) for _ in range(world_size)]
        dist.all_gather(gb, b_blk)
        B = sharding.combine_maps_radii_reference(torch.stack(gb), NX * NY, 1)
        blocks_ok = (torch.equal(B[: NX * NY].reshape(t.shape), t) and float(B[NX * NY]) == p.max_d2_mm2 and
                     abs(float(B[NX * NY + 1]) - p.sum_d2_mm2) <= 1e-12 * p.sum_d2_mm2 and int(A[13]) == 0 and int(A[14]) == 0 and
                     (-float(A[0]), float(A[1]), float(A[2]), -float(A[3])) == tuple(lrtb))
        # interleaved shards (sharding.shard_strided): the first valid ray of the whole bundle can sit on any rank -- rank 0's first
        # valid ray is id 4 (5 bounces), rank 1's is id 1 (2 bounces): the bundle's bounce level is 2 on both paths
        ps = F.OpmSpotPartial()
        ps.n_valid = ps.n_total = 1
        ps.first_key = ((4 << 32) | 5) if rank == 0 else ((1 << 32) | 2)
        sharding.reduce_spot_sums(ps)
        rows = torch.zeros((2, 11), dtype=torch.float64)
        rows[0, 7], rows[0, 9], rows[0, 10] = 1, 5, 4
        rows[1, 7], rows[1, 9], rows[1, 10] = 1, 2, 1
        As = sharding.combine_ranges_sums_reference(rows, 0, 1)
        strided_ok = int(ps.first_key) == ((1 << 32) | 2) and int(As[9]) == 2 and int(As[10]) == 1
        if rank == 0:
            q.put({"strided_ok": bool(strided_ok), "lrtb": lrtb, "map": t.numpy().copy(), "counts": counts, "energy": energy, "n_valid": st.n_valid,
                   "total_energy": st.total_energy, "centroid": list(st.energy_centroid), "geo": st.beam_radius_geo,
                   "rms": st.energy_beam_radius_rms, "bounce_lvl": st.bounce_lvl, "shard": (b, e), "blocks_ok": bool(blocks_ok)})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_every_index_once():
    for n, g in ((10, 3), (10_000_000, 8), (7, 8), (0, 2)):
        rs = [sharding.shard_range(n, r, g) for r in range(g)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    with pytest.raises(ValueError):
        sharding.shard_range(10, 3, 3)


def test_global_map_range_single_process():
    assert sharding.global_map_range((-1.0, 2.0, 3.0, -4.0)) == (-1.0, 2.0, 3.0, -4.0)
    with pytest.raises(ValueError):
        sharding.global_map_range(None)


@pytest.mark.timeout(300)
def test_two_rank_reduction_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process over the whole source
    src_all = scenes.source_to_oracle(scenes.c1_source(int(round(N_RAYS ** 0.5))))
    osc, out = _trace(src_all)
    fmap, lrtb = orc.fluence_binning(osc.nodes[3].s[0].merged_hits(), NX, NY)
    assert got["shard"] == (0, len(src_all) // 2)
    assert got["strided_ok"]  # one rule for the bounce level on both multi-GPU paths: the smallest global ray id
    assert got["blocks_ok"]  # the gathered-block combination gives the same map, range, radii and first ray as the all-reduces
    assert got["lrtb"] == pytest.approx(lrtb, rel=0, abs=0)  # min/max are exact
    peak = fmap.max()
    assert np.all(np.abs(got["map"] - fmap) <= 1e-12 * np.maximum(np.abs(fmap), peak))  # only the summation order differs
    assert got["counts"] == [int(((out["valid"] == 1) & (out["bounces"] == k)).sum()) for k in range(3)]
    assert got["energy"][0] == pytest.approx(float(out["e"][out["valid"] == 1].sum()), rel=1e-12)
    ref = osc.node_spot(2)
    assert got["n_valid"] == ref["n_valid"] and got["bounce_lvl"] == 0
    assert got["total_energy"] == pytest.approx(ref["total_energy"], rel=1e-12)
    assert np.allclose(got["centroid"][:2], ref["centroid"][:2], rtol=1e-9, atol=1e-15)
    assert got["geo"] == pytest.approx(ref["geo"], rel=1e-9)
    assert got["rms"] == pytest.approx(ref["rms"], rel=1e-9)


def test_shard_strided_partitions_every_index_once():
    for n, g in ((10, 3), (10_000_000, 8), (7, 8), (0, 2)):
        seen = []
        for r in range(g):
            first, count, stride = sharding.shard_strided(n, r, g)
            seen += [first + k * stride for k in range(count)] if n < 1000 else []
            assert count == len(range(r, n, g))
        if n < 1000:
            assert sorted(seen) == list(range(n))
