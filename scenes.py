"""Benchmark / parity scenes of BASELINE.json `configs`, as neutral descriptions (plain dicts).

Harness code shared by tests/, bench.py and __graft_entry__.py -- not part of the product package.  The same
description is turned into (a) the product's scene through the C ABI (`to_product`) and (b) the CPU oracle's scene
(`to_oracle`, test infrastructure).  Nodes carry explicit isometries, or -- `distance=` -- the edge distance from their
predecessor, to be placed by `calc_node_positions`; ambient medium = vacuum.  Glass data: H-K9L / H-ZF52 as shipped in the reference's examples
(opossum/examples/hhts/hhts.rs:38-64, refr_index_sellmeier1.rs:237-243).
"""
from __future__ import annotations

import math

MM = 1.0e-3
NM = 1.0e-9
HK9L = ("sellmeier1", 6.14555251E-1, 6.56775017E-1, 1.02699346E+0, 1.45987884E-2, 2.87769588E-3, 1.07653051E+2, 300 * NM, 2000 * NM)
HZF52 = ("schott", 3.26760058E+000, -2.05384566E-002, 3.51507672E-002, 7.70151348E-003, -9.08139817E-004, 7.52649555E-005,
         300 * NM, 2000 * NM)
NK9L_1054 = ("const", 1.5068)


def node(kind, z=0.0, xy=(0.0, 0.0), euler=(0.0, 0.0, 0.0), **kw):
    d = {"kind": kind, "t": (xy[0], xy[1], z), "euler": tuple(euler)}
    d.update(kw)
    return d


# ---------------------------------------------------------------------------------------------------------------
def c1_single_lens():
    """configs[0]: collimated source, lens, spot diagram + fluence detector (lens data: examples/lens_test.rs:28-34)."""
    return [
        node("source"),
        node("lens", z=50 * MM, p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=NK9L_1054, coat_in=("fresnel",), coat_out=("fresnel",)),
        node("spot_diagram", z=250 * MM),
        node("fluence_detector", z=260 * MM),
    ]


def c1_source(n_side=100):
    return {"pos": ("grid", n_side, n_side, 20 * MM, 20 * MM), "energy": ("uniform", 1.0), "wvl": [1000 * NM], "w": [1.0]}


def c2_laser_chain():
    """configs[1]: lens -> 45 deg mirror -> 45 deg mirror -> wedge -> beam splitter -> lens -> filter -> spot diagram ->
    fluence detector, plus a fluence detector on the splitter's second output.  13 surface passes per seed ray."""
    return [
        node("source"),                                                                                             # 0
        node("lens", z=100 * MM, p=(515.0 * MM, -515.0 * MM, 5.0 * MM), index=HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),  # 1
        node("thin_mirror", z=300 * MM, euler=(math.pi / 4, 0.0, 0.0), ap_in=("circle", 7.5 * MM, 0.0, 0.0)),          # 2: +z -> +y
        node("thin_mirror", z=300 * MM, xy=(0.0, 200 * MM), euler=(-3 * math.pi / 4, 0.0, 0.0)),                      # 3: +y -> +z
        node("wedge", z=400 * MM, xy=(0.0, 200 * MM), p=(10 * MM, math.radians(1.0)), index=HK9L,
             coat_in=("constant_r", 0.01), coat_out=("constant_r", 0.01)),                                          # 4
        node("beam_splitter", z=500 * MM, xy=(0.0, 200 * MM), p=(0.5,)),                                             # 5
        node("lens", z=600 * MM, xy=(0.0, 200 * MM), p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=HK9L,
             coat_in=("fresnel",), coat_out=("fresnel",), ap_in=("circle", 12.5 * MM, 0.0, 200 * MM * 0.0)),         # 6
        node("ideal_filter", z=700 * MM, xy=(0.0, 200 * MM), p=(0.9,)),                                              # 7
        node("spot_diagram", z=800 * MM, xy=(0.0, 200 * MM)),                                                        # 8
        node("fluence_detector", z=810 * MM, xy=(0.0, 200 * MM)),                                                    # 9
        node("fluence_detector", z=700 * MM, xy=(0.0, 200 * MM), parent=5, parent_port=1),                           # 10: second output
    ]


def c2_source(n):
    """FibonacciEllipse r = 10 mm, General2DGaussian power 5 (examples/ghost_focus.rs:26-33 style), 1054 nm, 1 J."""
    return {"pos": ("fibonacci_ellipse", n, 10 * MM, 10 * MM),
            "energy": ("general_gaussian", 1.0, (0.0, 0.0), (6 * MM, 6 * MM), 5.0, 0.0, False), "wvl": [1054 * NM], "w": [1.0]}


def c3_telescope():
    """configs[2]: two-lens Kepler telescope (4 refracting surfaces, ConstantR 0.05 / Fresnel as in
    examples/ghost_focus.rs:40-66) + detector, ghost-focus analysis."""
    return [
        node("source"),
        node("lens", z=100 * MM, p=(100 * MM, -100 * MM, 8 * MM), index=NK9L_1054, coat_in=("constant_r", 0.05), coat_out=("constant_r", 0.05)),
        node("lens", z=350 * MM, p=(150 * MM, -150 * MM, 8 * MM), index=NK9L_1054, coat_in=("fresnel",), coat_out=("fresnel",)),
        node("spot_diagram", z=500 * MM),
    ]


def c3_source(n):
    return {"pos": ("fibonacci_ellipse", n, 8 * MM, 8 * MM), "energy": ("uniform", 1.0), "wvl": [1053 * NM], "w": [1.0]}


def c4_fluence():
    """configs[3]: lens + fluence detector, 4096 x 4096 Binning map."""
    return [
        node("source"),
        node("lens", z=50 * MM, p=(205.55 * MM, -205.55 * MM, 2.79 * MM), index=NK9L_1054, coat_in=("fresnel",), coat_out=("fresnel",)),
        node("fluence_detector", z=120 * MM),
    ]


def c4_source(n_side):
    return {"pos": ("grid", n_side, n_side, 20 * MM, 20 * MM), "energy": ("uniform", 1.0), "wvl": [1000 * NM], "w": [1.0]}


def c5_broadband():
    """configs[4]: Sellmeier lens + a 10 deg prism (wedge) pair as in examples/prism_dispersion.rs:23-31,47-75."""
    return [
        node("source"),
        node("lens", z=50 * MM, p=(300 * MM, -300 * MM, 5 * MM), index=HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),
        node("wedge", z=100 * MM, p=(10 * MM, math.radians(10.0)), index=HK9L),
        node("wedge", z=150 * MM, p=(10 * MM, math.radians(-10.0)), index=HZF52),
        node("fluence_detector", z=300 * MM),
    ]


def c5_source(n_pos, n_wvl=64):
    return {"pos": ("fibonacci_rectangle", n_pos, 10 * MM, 10 * MM), "energy": ("uniform", 1.0),
            "spectrum": ("gaussian", 950 * NM, 1150 * NM, n_wvl, 1053 * NM, 50 * NM, 1.0)}


def folded_chain_by_distances():
    """A folded beam path described the way the reference's examples do (examples/folded_telescope.rs, laser_system.rs): only
    the source has an isometry, every other node hangs `distance` behind its predecessor and is placed by
    `calc_node_positions`; the two mirrors carry their tilt as a local alignment, the wedge deflects the axis."""
    return [
        node("source", z=10 * MM, xy=(1 * MM, -2 * MM), euler=(0.02, -0.03, 0.0)),
        node("lens", distance=100 * MM, p=(515.0 * MM, -515.0 * MM, 5.0 * MM), index=HK9L, coat_in=("fresnel",), coat_out=("fresnel",)),
        node("thin_mirror", distance=150 * MM, alignment=((0.0, 0.0, 0.0), (math.radians(22.5), 0.0, 0.0))),
        node("thin_mirror", distance=120 * MM, alignment=((0.0, 0.0, 0.0), (0.0, math.radians(-30.0), 0.0))),
        node("wedge", distance=80 * MM, p=(8 * MM, math.radians(3.0)), index=HK9L),
        node("beam_splitter", distance=60 * MM, p=(0.3,)),
        node("paraxial_surface", distance=50 * MM, p=(400 * MM,)),
        node("ideal_filter", distance=20 * MM, p=(0.8,)),
        node("spot_diagram", distance=150 * MM),
        node("fluence_detector", distance=10 * MM),
        node("lens", distance=40 * MM, p=(300 * MM, math.inf, 4 * MM), index=NK9L_1054, parent=5, parent_port=1),
        node("fluence_detector", distance=90 * MM),
    ]


# ---------------------------------------------------------------------------------------------------------------
KIND_IDS = {"source": 0, "lens": 1, "cylindric_lens": 2, "wedge": 3, "thin_mirror": 4, "parabolic_mirror": 5, "beam_splitter": 6,
            "dummy": 7, "ideal_filter": 8, "paraxial_surface": 9, "spot_diagram": 10, "fluence_detector": 11, "reflective_grating": 12,
            "ray_propagation_visualizer": 13, "spectrometer": 14, "wavefront": 15, "energy_meter": 16}


def to_product(ctx, nodes):
    """neutral description -> opossum_b200.scene.Scene (through the C ABI)"""
    import opossum_b200 as ob
    from opossum_b200 import scene as S

    def ap(a):
        return getattr(ob.Aperture, a[0])(*a[1:]) if a else None

    def coat(c):
        return getattr(S.CoatingType, c[0])(*c[1:]) if c else None

    sc = S.Scene(ctx)
    for d in nodes:
        n = S.Node(KIND_IDS[d["kind"]], ob.Isometry.new(d["t"], d["euler"]))
        for k, v in enumerate(d.get("p", ())):
            n.c.p[k] = v
        if "index" in d:
            n.c.index = getattr(ob.RefractiveIndex, d["index"][0])(*d["index"][1:]).c
        n.with_coatings(coat(d.get("coat_in")), coat(d.get("coat_out")))
        n.with_apertures(ap(d.get("ap_in")), ap(d.get("ap_out")))
        if "alignment" in d:
            n.with_alignment(*d["alignment"])
        if "lidt" in d:
            n.with_lidt(d["lidt"])
        if "spectrum" in d:  # (wavelength slots [um], values): FilterType::Spectrum / SplittingConfig::Spectrum
            n.with_spectrum(ob.Spectrum(ctx, *d["spectrum"]))
        n.fed_by(d.get("parent", -1), d.get("parent_port", 0))
        sc.add_node(n, distance=d.get("distance"), align_like=d.get("align_like"))
    return sc


def to_oracle(nodes, n_threads=1):
    """neutral description -> oracle.scene.OScene (CPU restatement; test infrastructure)"""
    from oracle import oracle as orc
    from oracle import scene as OS
    coat_ids = {"ideal_ar": orc.COAT_IDEAL_AR, "constant_r": orc.COAT_CONSTANT_R, "fresnel": orc.COAT_FRESNEL}

    def ap(a):
        return getattr(orc, "ap_" + a[0])(*a[1:]) if a else None

    def coat(c):
        return (coat_ids[c[0]], c[1] if len(c) > 1 else 0.0) if c else None

    out = []
    for d in nodes:
        idx = getattr(orc.IndexModel, d["index"][0])(*d["index"][1:]) if "index" in d else orc.IndexModel.const(1.5)
        align = orc.Isometry.new(*d["alignment"]) if "alignment" in d else None
        out.append(OS.ONode(KIND_IDS[d["kind"]], orc.Isometry.new(d["t"], d["euler"]), d.get("p", ()), idx, coat(d.get("coat_in")),
                            coat(d.get("coat_out")), ap(d.get("ap_in")), ap(d.get("ap_out")), align, d.get("lidt", 1.0e4),
                            d.get("parent", -1), d.get("parent_port", 0), spectrum=d.get("spectrum"), distance=d.get("distance"),
                            align_like=d.get("align_like")))
    return OS.OScene(out, n_threads=n_threads)


def spectrum_of(src, backend="product"):
    """(wavelengths, weights) of a source description; Gaussian spectra (spectral_distribution/gaussian.rs:78-96) come
    from the library's host helper for the product and from the oracle's own restatement for the oracle"""
    if "spectrum" in src:
        _, lo, hi, n, mu, fwhm, power = src["spectrum"]
        if backend == "oracle":
            from oracle import oracle as orc
            return orc.spectrum_gaussian(lo, hi, n, mu, fwhm, power)
        import opossum_b200 as ob
        return ob.spectrum_gaussian(lo, hi, n, mu, fwhm, power)
    import numpy as np
    return np.asarray(src["wvl"], dtype=float), np.asarray(src["w"], dtype=float)


def source_to_product(src, wvl=None, w=None):
    """source description -> opossum_b200.SourceDesc (generated on the device)"""
    import opossum_b200 as ob
    if wvl is None:
        wvl, w = spectrum_of(src)
    pos = src["pos"]
    n_pos = pos[1] * pos[2] if pos[0] == "grid" else (1 + 3 * pos[1] * (pos[1] + 1) if pos[0] == "hexapolar" else pos[1])
    en = src["energy"]
    sd = ob.SourceDesc(n_pos, wvl, w, total_energy=en[1])
    if pos[0] == "grid":
        sd.grid(pos[1], pos[2], pos[3], pos[4])
    elif pos[0] == "fibonacci_rectangle":
        sd.fibonacci_rectangle(pos[2], pos[3])
    elif pos[0] == "hexapolar":
        sd.hexapolar(pos[1], pos[2])
    else:
        sd.fibonacci_ellipse(pos[2], pos[3])
    if en[0] == "general_gaussian":
        sd.general_gaussian(en[2], en[3], en[4], en[5], en[6])
    return sd


def source_to_oracle(src, wvl=None, w=None):
    """source description -> oracle ray array (Rays::new_collimated_with_spectrum, rays.rs:264-297)"""
    from oracle import oracle as orc
    if wvl is None:
        wvl, w = spectrum_of(src, backend="oracle")
    pos = src["pos"]
    if pos[0] == "grid":
        xyz = orc.grid(pos[1], pos[2], pos[3], pos[4])
    elif pos[0] == "fibonacci_rectangle":
        xyz = orc.fibonacci_rectangle(pos[1], pos[2], pos[3])
    elif pos[0] == "hexapolar":
        xyz = orc.hexapolar(pos[1], pos[2])
    else:
        xyz = orc.fibonacci_ellipse(pos[1], pos[2], pos[3])
    en = src["energy"]
    e = orc.energy_uniform(len(xyz), en[1]) if en[0] == "uniform" else orc.energy_general_gaussian(xyz, en[1], en[2], en[3], en[4], en[5], en[6])
    return orc.rays_collimated_with_spectrum(xyz, e, wvl, w)
