"""Shared builders for the parity tests: the same neutral description -> oracle objects and product objects."""
from __future__ import annotations

import math

import numpy as np

import opossum_b200 as ob
from opossum_b200 import ffi as F
from oracle import oracle as orc

RTOL = 1e-9  # BASELINE.json north_star: per-ray positions, directions, energies and OPL within 1e-9 relative

HK9L = (6.14555251E-1, 6.56775017E-1, 1.02699346E+0, 1.45987884E-2, 2.87769588E-3, 1.07653051E+2)
HZF52 = (3.26760058E+000, -2.05384566E-002, 3.51507672E-002, 7.70151348E-003, -9.08139817E-004, 7.52649555E-005)


def mm(*v):
    r = tuple(x * 1.0e-3 for x in v)
    return r[0] if len(r) == 1 else r


def nm(x):
    return x * 1.0e-9


def iso_pair(translation=(0.0, 0.0, 0.0), euler=(0.0, 0.0, 0.0)):
    return orc.Isometry.new(translation, euler), ob.Isometry.new(translation, euler)


def append_pair(a, b):
    return a[0].append(b[0]), a[1].append(b[1])


def surface_pair(geom, param=0.0, iso=None, coating=0, r=0.0):
    """geom/coating ints are shared enum values (checked in test_abi)."""
    iso = iso or iso_pair()
    return (orc.Surface(geom=geom, coating=coating, param=param, coating_r=r, iso=iso[0]),
            ob.OpticSurface(geom, param, iso[1], coating, r))


def node_sphere_pair(radius, node_iso, **kw):
    """Sphere as a lens/mirror node builds it: geo iso = node_iso o T(0,0,R) (nodes/lens/mod.rs:238-247)."""
    return surface_pair(F.GEOM_SPHERE, radius, append_pair(node_iso, iso_pair((0.0, 0.0, radius))), **kw)


def index_pair(kind, *a):
    return getattr(orc.IndexModel, kind)(*a), getattr(ob.RefractiveIndex, kind)(*a)


def aperture_pair(kind, *a, **kw):
    if kind == "none":
        return orc.ap_none(), ob.Aperture.none()
    if kind == "stack":
        items = a[0]
        return orc.ap_stack([i[0] for i in items], **kw), ob.Aperture.stack([i[1] for i in items], **kw)
    return getattr(orc, "ap_" + kind)(*a, **kw), getattr(ob.Aperture, kind)(*a, **kw)


def make_bundle(n=4096, radius=mm(8.0), z=0.0, divergence=0.02, seed=7, wvl=(nm(600.0), nm(1500.0))):
    """Deterministic, boundary-avoiding bundle: Fibonacci ellipse positions, smoothly varying directions/energies."""
    i = np.arange(n, dtype=np.float64)
    xyz = orc.fibonacci_ellipse(n, radius, radius)
    xyz[:, 2] = z
    rng = np.random.default_rng(seed)
    d = np.stack([divergence * (xyz[:, 0] / radius) + 0.003 * rng.standard_normal(n),
                  divergence * (xyz[:, 1] / radius) + 0.003 * rng.standard_normal(n), np.ones(n)], axis=1)
    e = 1.0e-3 * (1.0 + 0.5 * np.sin(0.37 * i))
    w = wvl[0] + (wvl[1] - wvl[0]) * ((i * 0.6180339887498949) % 1.0)
    return orc.rays_from_arrays(xyz, d, e, w)


def sort_by_id(a: np.ndarray) -> np.ndarray:
    return a[np.argsort(a["id"], kind="stable")]


def assert_rays_match(gpu: np.ndarray, ref: np.ndarray, rtol=RTOL, keyed=False, what="rays", e_floor=1e-30):
    """Counts and integer state bit-exact; pos/dir/e/opl/n within rtol (vector norms for pos/dir).
    e_floor: energies below it are rounding noise of a cancelling formula (Fresnel at equal indices) and are
    compared against the floor instead of their own magnitude."""
    assert len(gpu) == len(ref), f"{what}: bundle length {len(gpu)} != {len(ref)}"
    if len(ref) == 0:
        return
    if keyed:
        gpu, ref = sort_by_id(gpu), sort_by_id(ref)
    for f in ("id", "valid", "bounces", "refr"):
        assert np.array_equal(gpu[f], ref[f]), f"{what}: field {f} differs at {np.flatnonzero(gpu[f] != ref[f])[:5]}"
    scale_p = np.maximum(np.linalg.norm(ref["pos"], axis=1), 1e-3)
    dp = np.linalg.norm(gpu["pos"] - ref["pos"], axis=1) / scale_p
    assert dp.max() <= rtol, f"{what}: pos rel err {dp.max():.3e} at {dp.argmax()}"
    dd = np.linalg.norm(gpu["dir"] - ref["dir"], axis=1) / np.maximum(np.linalg.norm(ref["dir"], axis=1), 1e-300)
    assert dd.max() <= rtol, f"{what}: dir rel err {dd.max():.3e} at {dd.argmax()}"
    for f, floor in (("e", e_floor), ("opl", 1e-12), ("n", 1e-30), ("wvl", 1e-30)):
        err = np.abs(gpu[f] - ref[f]) / np.maximum(np.abs(ref[f]), floor)
        assert err.max() <= rtol, f"{what}: {f} rel err {err.max():.3e} at {err.argmax()}"


def assert_hits_match(gpu_xyzeb, ref_hits: np.ndarray, rtol=RTOL):
    x, y, z, e, b = gpu_xyzeb
    assert len(x) == len(ref_hits), f"hit count {len(x)} != {len(ref_hits)}"
    if len(x) == 0:
        return
    assert np.array_equal(b, ref_hits["bounce"])
    p = np.stack([x, y, z], axis=1)
    q = np.stack([ref_hits["x"], ref_hits["y"], ref_hits["z"]], axis=1)
    scale = np.maximum(np.linalg.norm(q, axis=1), 1e-3)
    assert (np.linalg.norm(p - q, axis=1) / scale).max() <= rtol
    assert (np.abs(e - ref_hits["e"]) / np.maximum(np.abs(ref_hits["e"]), 1e-30)).max() <= rtol


def _max3x3(a):
    """maximum over the 3 x 3 neighbourhood of every bin"""
    p = np.pad(a, 1, mode="constant", constant_values=0.0)
    out = a.copy()
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            out = np.maximum(out, p[dy:dy + a.shape[0], dx:dx + a.shape[1]])
    return out


def assert_map_match(gpu_map, ref_map, rtol=RTOL, floor=1e-12, what="fluence map", same_hits=False):
    """BASELINE north_star: fluence maps within 1e-9 RELATIVE per bin.

    same_hits=True (both sides binned the SAME hit points: the Binning kernels on their own): every bin the reference fills to
    at least `floor` x peak (1e-12) is compared against ITS OWN value; a bin below that has to stay below it in the product.

    same_hits=False (the hits come from two traces: positions agree to ~1e-15 of the path length, not bit by bit): a hit that moves
    by d bin widths moves d x its fluence between ADJACENT bins (cloud-in-cell weights are linear in the position), so a bin that
    holds 1e-7 of what its neighbours hold -- the far corner of one tap -- inherits their position noise.  The scale of a bin is
    therefore the largest reference value in its 3 x 3 neighbourhood: |difference| <= rtol x that.  For the bins of a filled map
    this is the per-bin relative bar (a bin and its neighbours are of one magnitude); it is never weaker than rtol x peak.

    Returns the worst error relative to the scale used."""
    assert gpu_map.shape == ref_map.shape
    if ref_map.size == 0:
        return 0.0
    nan_g, nan_r = np.isnan(gpu_map), np.isnan(ref_map)
    assert np.array_equal(nan_g, nan_r), f"{what}: NaN bins differ"
    same_inf = np.isinf(ref_map) & (gpu_map == ref_map)
    skip = nan_r | same_inf
    g, r = np.where(skip, 0.0, gpu_map), np.where(skip, 0.0, ref_map)
    peak = np.max(np.abs(r[np.isfinite(r)])) if np.isfinite(r).any() else 0.0
    if peak == 0.0:
        assert np.array_equal(g, r), f"{what}: the reference map is empty, the product's is not"
        return 0.0
    scale = np.abs(r) if same_hits else _max3x3(np.where(np.isfinite(r), np.abs(r), 0.0))
    filled = scale >= floor * peak
    low = np.abs(g - r)[~filled]
    assert low.size == 0 or low.max() <= floor * peak, (f"{what}: a bin the reference leaves below {floor:g} of the peak holds "
                                                        f"{low.max() / peak:.3e} of the peak in the product")
    if not filled.any():
        return 0.0
    rel = np.where(filled, np.abs(g - r) / np.where(filled, scale, 1.0), 0.0)
    worst = float(rel.max())
    k = np.unravel_index(int(rel.argmax()), rel.shape)
    assert worst <= rtol, (f"{what}: worst relative bin error {worst:.3e} at {k} (ref {r[k]:.6e} = {r[k] / peak:.2e} of the peak, "
                           f"got {g[k]:.6e}, scale {scale[k]:.6e}); {int((rel > rtol).sum())} of {int(filled.sum())} filled bins above {rtol:g}")
    return worst
