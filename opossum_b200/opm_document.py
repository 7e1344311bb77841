"""`.opm` scene files: the reference's `OpmDocument` (opm_document.rs:70-81) as serde writes it in RON, read
into the scene layer of this package.

    doc = load_opm("scene.opm")                    # parse + interpret
    scene, source, index = doc.build(ctx)          # opossum_b200.scene.Scene, SourceDesc | None, {uuid: node index}
    scene.calc_node_positions(doc.axis_wavelength())
    rays = source.generate(ctx); scene.raytrace(rays, doc.raytrace_config())

Host-side glue above the C ABI, no ray arithmetic: a RON subset parser (structs `(a: 1, ..)`, tuples, lists, maps, enum
variants `Name(..)`, strings, numbers, booleans, `Some`/`None`), then the mapping of the serialised structures onto
`OpmNodeDesc` / `OpmSourceDesc`.  What the mapping follows in the reference (opossum/src/):
  nodes/node_attr.rs:26-49 (`NodeAttr`: node_type, name, ports, uuid, lidt, props, isometry, inverted, alignment),
  surface/optic_surface.rs:29-43 (ports: anchor_point_iso, aperture, coating, lidt), properties/proptype.rs:59-130,
  nodes/node_group/optic_graph.rs:1040-1075 (graph: nodes, edges (src uuid, src port, target uuid, target port, distance)),
  the property names each node registers (nodes/lens/mod.rs:58-80, wedge/mod.rs:54-70, thin_mirror.rs:58, beam_splitter/mod.rs:52,
  ideal_filter.rs:58, paraxial_surface.rs, parabolic_mirror.rs, reflective_grating.rs:63-71, source.rs:57-71),
  lightdata/ray_data_builder.rs:20-31 (`Collimated { pos_dist, energy_dist, spect_dist }`) with the distribution structs of
  position_distributions/, energy_distributions/, spectral_distribution/.
uom quantities are serialised as plain numbers in SI base units, nalgebra points as tuples, `Isometry` as its transform
(rotation quaternion (i, j, k, w) + translation).  Scope: a chain of nodes with at most side chains behind beam splitters, one
source, the node types of `include/opossum_b200_scene.h`; nested `NodeGroup`s are taken apart (`_flatten_graph`); inverted nodes and
groups, reference nodes, `Raw` / `PointSrc` light data and spectrum-valued filters are rejected with `OpmDocumentError`.  Pinned by the reference's fixture files
(opossum/files_for_testing/opm/, read by tests/test_opm_document.py where they lie).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

__all__ = ["parse_ron", "load_opm", "loads_opm", "OpmDocument", "OpmDocumentError"]


class OpmDocumentError(ValueError):
    """OpossumError::OpmDocument / parse errors"""


# ---------------------------------------------------------------------------------------------------------------
# RON subset
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Variant:
    """enum variant or named struct: `Name`, `Name(a, b)` (args) or `Name(x: 1)` (fields)"""
    name: str
    args: tuple = ()
    fields: dict | None = None

    def single(self):
        if self.fields is not None or len(self.args) != 1:
            raise OpmDocumentError(f"variant {self.name} does not wrap a single value")
        return self.args[0]


_TOKEN = re.compile(r"""\s*(?:(//[^\n]*)|("(?:[^"\\]|\\.)*")|([A-Za-z_][A-Za-z0-9_]*)|([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|NaN))|([(){}\[\]:,]))""")


def _tokens(text: str):
    pos, out = 0, []
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return out
            raise OpmDocumentError(f"unexpected character at offset {pos}: {text[pos:pos + 20]!r}")
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2) is not None:
            out.append(("str", bytes(m.group(2)[1:-1], "utf-8").decode("unicode_escape")))
        elif m.group(3) is not None:
            out.append(("id", m.group(3)))
        elif m.group(4) is not None:
            t = m.group(4)
            if t in ("inf", "+inf"):
                out.append(("num", math.inf))
            elif t == "-inf":
                out.append(("num", -math.inf))
            elif t.endswith("NaN"):
                out.append(("num", math.nan))
            else:
                out.append(("num", int(t) if re.fullmatch(r"[-+]?\d+", t) else float(t)))
        else:
            out.append(("p", m.group(5)))


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def take(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise OpmDocumentError(f"expected {val or kind}, found {tok[1]!r} (token {self.i})")
        self.i += 1
        return tok

    def value(self):
        kind, v = self.peek()
        if kind in ("str", "num"):
            self.i += 1
            return v
        if kind == "id":
            self.i += 1
            if v == "true":
                return True
            if v == "false":
                return False
            if v in ("inf", "NaN"):
                return math.inf if v == "inf" else math.nan
            if self.peek() == ("p", "("):
                args, fields = self.parens()
                if v == "Some":
                    return args[0]
                return Variant(v, args, fields)
            return None if v == "None" else Variant(v)
        if (kind, v) == ("p", "("):
            args, fields = self.parens()
            return fields if fields is not None else args
        if (kind, v) == ("p", "["):
            self.i += 1
            out = []
            while self.peek() != ("p", "]"):
                out.append(self.value())
                self.separator("]")
            self.i += 1
            return out
        if (kind, v) == ("p", "{"):
            self.i += 1
            out = {}
            while self.peek() != ("p", "}"):
                k = self.value()
                self.take("p", ":")
                out[k] = self.value()
                self.separator("}")
            self.i += 1
            return out
        raise OpmDocumentError(f"unexpected token {v!r} (token {self.i})")

    def parens(self):
        """( a: 1, b: 2 ) -> ((), {fields});  ( 1, 2 ) -> ((1, 2), None)"""
        self.take("p", "(")
        named = self.peek()[0] == "id" and self.peek(1) == ("p", ":")
        args, fields = [], {} if named else None
        while self.peek() != ("p", ")"):
            if named:
                k = self.take("id")[1]
                self.take("p", ":")
                fields[k] = self.value()
            else:
                args.append(self.value())
            self.separator(")")
        self.i += 1
        return tuple(args), fields

    def separator(self, close):
        """a comma, or the closing bracket right away"""
        if self.peek() == ("p", ","):
            self.i += 1
        elif self.peek() != ("p", close):
            raise OpmDocumentError(f"expected ',' or '{close}', found {self.peek()[1]!r} (token {self.i})")


def parse_ron(text: str):
    """structs -> dict, tuples -> tuple, lists -> list, maps -> dict, `Name(..)` -> Variant, `Some(x)` -> x, `None` -> None"""
    p = _Parser(_tokens(text))
    if p.peek()[0] == "eof":
        raise OpmDocumentError("empty document")
    v = p.value()
    if p.peek()[0] != "eof":
        raise OpmDocumentError("trailing characters after the document")
    return v


# ---------------------------------------------------------------------------------------------------------------
# interpretation
# ---------------------------------------------------------------------------------------------------------------
NODE_TYPES = {"source": 0, "lens": 1, "cylindric lens": 2, "wedge": 3, "mirror": 4, "parabolic mirror": 5, "beam splitter": 6,
              "dummy": 7, "ideal filter": 8, "paraxial surface": 9, "spot diagram": 10, "fluence detector": 11,
              "reflective grating": 12, "ray propagation": 13, "spectrometer": 14, "wavefront monitor": 15,
              "energy meter": 16}


def _quat_to_euler_free(rot, trans):
    """(i, j, k, w) + translation -> row-major 3x3 rotation, translation"""
    i, j, k, w = (float(c) for c in rot)
    n = math.sqrt(i * i + j * j + k * k + w * w)
    if not n > 0.0:
        raise OpmDocumentError("zero quaternion in an isometry")
    i, j, k, w = i / n, j / n, k / n, w / n
    r = (1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w),
         2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w),
         2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j))
    return r, tuple(float(c) for c in trans)


def _isometry(v):
    """serialised `Isometry` (utils/geom_transformation.rs:27-31; `inverse` is rebuilt) -> (rotation 3x3, translation) | None"""
    if v is None:
        return None
    tr = v["transform"]
    return _quat_to_euler_free(tr["rotation"], tr["translation"])


def _index(v):
    """RefractiveIndexType (refractive_index/mod.rs:22-31) -> ("const", n) | ("sellmeier1", k1..l3, lo, hi) | ..."""
    s = v.single()
    if v.name == "Const":
        return ("const", float(s["refractive_index"]))
    rng = s["wvl_range"]
    lo, hi = float(rng["start"]), float(rng["end"])
    if v.name == "Sellmeier1":
        return ("sellmeier1",) + tuple(float(s[k]) for k in ("k1", "k2", "k3", "l1", "l2", "l3")) + (lo, hi)
    if v.name == "Schott":
        return ("schott",) + tuple(float(s[k]) for k in ("a0", "a1", "a2", "a3", "a4", "a5")) + (lo, hi)
    if v.name == "Conrady":
        return ("conrady", float(s["n0"]), float(s["a"]), float(s["b"]), lo, hi)
    raise OpmDocumentError(f"unknown refractive index model {v.name}")


def _coating(v):
    """CoatingType (coatings/mod.rs:18-29)"""
    if v.name == "IdealAR":
        return ("ideal_ar",)
    if v.name == "Fresnel":
        return ("fresnel",)
    if v.name == "ConstantR":
        return ("constant_r", float(v.fields["reflectivity"]))
    raise OpmDocumentError(f"unknown coating {v.name}")


def _aperture(v):
    """Aperture (aperture.rs:66-84).  A BinaryPolygon is serialised with its earcut triangulation (PolygonConfig: points,
    aperture_type, triangle_indices; aperture.rs:247-251), which is what `in_polygon` tests a point against (:289-316) -- the
    triangles are read, nothing is triangulated here.  A Stack (StackConfig, :385-416) is read when it is a flat list of simple
    apertures with at most one polygon, which is what the device aperture holds."""
    if v is None or v.name == "None":
        return None
    c = v.single()
    obstruction = c["aperture_type"].name == "Obstruction"
    if v.name == "BinaryPolygon":
        pts = [(float(q[0]), float(q[1])) for q in c["points"]]
        tris = []
        for tri in c["triangle_indices"]:
            idx = [int(i) for i in tri]
            if len(idx) != 3 or any(i < 0 or i >= len(pts) for i in idx):
                raise OpmDocumentError("polygon aperture: a triangle does not name three of its points")
            tris.append([coord for i in idx for coord in pts[i]])
        if not tris:
            raise OpmDocumentError("polygon aperture without triangles")
        return ("polygon", tris, obstruction)
    if v.name == "Stack":
        items = [_aperture(a) for a in c["apertures"]]
        if any(a is None or a[0] == "stack" for a in items) or sum(a[0] == "polygon" for a in items) > 1:
            raise OpmDocumentError("aperture stack: only flat stacks of simple apertures with at most one polygon are supported")
        return ("stack", items, obstruction)
    cx, cy = (float(t) for t in c["center"])
    if v.name == "BinaryCircle":
        return ("circle", float(c["radius"]), cx, cy, obstruction)
    if v.name == "BinaryRectangle":
        return ("rect", float(c["width"]), float(c["height"]), cx, cy, obstruction)
    if v.name == "Gaussian":
        return ("gaussian", float(c["sigma"][0]), float(c["sigma"][1]), cx, cy, obstruction)
    raise OpmDocumentError(f"aperture {v.name} is not supported")


def _prop(props: dict, name: str, kind: str, default=None):
    v = props.get(name)
    if v is None:
        return default
    if not isinstance(v, Variant) or v.name != kind:
        raise OpmDocumentError(f"property {name!r}: expected {kind}, found {getattr(v, 'name', v)!r}")
    return v.single() if v.args or v.fields is not None else None


@dataclass
class OpmSource:
    pos: tuple
    energy: tuple
    wavelengths: list
    weights: list
    spectrum: tuple | None = None  # ("gaussian", lo, hi, n, mu, fwhm, power)


def _source(builder) -> OpmSource:
    """LightDataBuilder::Geometric(RayDataBuilder::Collimated { .. }) (lightdata/ray_data_builder.rs:20-31)"""
    if builder is None:
        raise OpmDocumentError("source has empty light data builder")
    if builder.name != "Geometric":
        raise OpmDocumentError(f"light data builder {builder.name} is not ray data")
    rb = builder.single()
    if rb.name != "Collimated":
        raise OpmDocumentError(f"ray data builder {rb.name} is not supported (only Collimated)")
    pd, ed, sd = rb.fields["pos_dist"], rb.fields["energy_dist"], rb.fields["spect_dist"]
    p = pd.single()
    if pd.name == "Grid":
        pos = ("grid", int(p["nr_of_points"][0]), int(p["nr_of_points"][1]), float(p["side_length"][0]), float(p["side_length"][1]))
    elif pd.name == "Hexapolar":
        pos = ("hexapolar", int(p["nr_of_rings"]), float(p["radius"]))
    elif pd.name == "FibonacciRectangle":
        pos = ("fibonacci_rectangle", int(p["nr_of_rays"]), float(p["side_length_x"]), float(p["side_length_y"]))
    elif pd.name == "FibonacciEllipse":
        pos = ("fibonacci_ellipse", int(p["nr_of_rays"]), float(p["radius_x"]), float(p["radius_y"]))
    else:
        raise OpmDocumentError(f"position distribution {pd.name} is not generated on the device")
    e = ed.single()
    if ed.name == "Uniform":
        energy = ("uniform", float(e["total_energy"]))
    elif ed.name == "General2DGaussian":
        energy = ("general_gaussian", float(e["total_energy"]), tuple(float(c) for c in e["mu_xy"]), tuple(float(c) for c in e["sigma_xy"]),
                  float(e["power"]), float(e["theta"]), bool(e["rectangular"]))
    else:
        raise OpmDocumentError(f"energy distribution {ed.name} is not supported")
    s = sd.single()
    if sd.name == "LaserLines":
        wl = [float(l[0]) for l in s["lines"]]
        wt = [float(l[1]) for l in s["lines"]]
        tot = sum(wt)
        return OpmSource(pos, energy, wl, [w / tot for w in wt])  # spectral_distribution/laser_lines.rs: weights normalised
    if sd.name == "Gaussian":
        lo, hi = (float(c) for c in s["wvl_range"])
        return OpmSource(pos, energy, [], [], ("gaussian", lo, hi, int(s["num_points"]), float(s["mu"]), float(s["fwhm"]), float(s["power"])))
    raise OpmDocumentError(f"spectral distribution {sd.name} is not supported")


@dataclass
class OpmNode:
    uuid: str
    node_type: str
    name: str
    kind: int
    p: tuple = ()
    index: tuple | None = None
    coatings: tuple = (None, None)
    apertures: tuple = (None, None)
    lidt: float = 1.0e4
    isometry: tuple | None = None   # (rotation 3x3, translation) of NodeAttr::isometry, None = placed by calc_node_positions
    alignment: tuple | None = None
    inverted: bool = False
    source: OpmSource | None = None
    alignment_wavelength: float | None = None
    fluence_estimator: str | None = None
    align_like: tuple | None = None   # NodeAttr::align_like_node_at_distance: (uuid, distance [m])


def _node(entry: dict) -> OpmNode:
    a = entry["attributes"] if "attributes" in entry else entry["node_attr"]
    t = a["node_type"]
    if t not in NODE_TYPES:
        raise OpmDocumentError(f"node type {t!r} is not supported by the GPU path")
    props = a.get("props") or {}
    n = OpmNode(uuid=a["uuid"], node_type=t, name=a.get("name", t), kind=NODE_TYPES[t], lidt=float(a.get("lidt", 1.0e4)),
                isometry=_isometry(a.get("isometry")), alignment=_isometry(a.get("alignment")), inverted=bool(a.get("inverted", False)))
    al = a.get("align_like_node_at_distance")
    if al is not None:
        n.align_like = (str(al[0]), float(al[1]))
    if n.inverted:
        raise OpmDocumentError(f"node {n.name!r} is inverted: not supported")
    ports = a.get("ports") or {}

    def port(group, name):
        s = (ports.get(group) or {}).get(name)
        return (None, None) if s is None else (_coating(s["coating"]), _aperture(s.get("aperture")))
    names = {"beam splitter": ("input_1", "out1_trans1_refl2")}.get(t, ("input_1", "output_1"))
    cin, ain = port("inputs", names[0])
    cout, aout = port("outputs", names[1])
    n.coatings, n.apertures = (cin, cout), (ain, aout)
    inf = math.inf
    if t in ("lens", "cylindric lens"):
        n.p = (_prop(props, "front curvature", "Length", 0.5), _prop(props, "rear curvature", "Length", -0.5),
               _prop(props, "center thickness", "Length", 0.01))
        n.index = _index(props["refractive index"].single()) if "refractive index" in props else ("const", 1.5)
    elif t == "wedge":
        n.p = (_prop(props, "center thickness", "Length", 0.01), _prop(props, "wedge", "Angle", 0.0))
        n.index = _index(props["refractive index"].single()) if "refractive index" in props else ("const", 1.5)
    elif t == "mirror":
        n.p = (_prop(props, "curvature", "Length", inf),)
    elif t == "parabolic mirror":  # nodes/parabolic_mirror.rs:45-70
        d = _prop(props, "oa direction", "Vec2", (1.0, 0.0))
        n.p = (_prop(props, "focal length", "Length", 1.0), _prop(props, "oa angle", "Angle", 0.0), float(d[0]), float(d[1]),
               1.0 if _prop(props, "collimating", "Bool", False) else 0.0)
    elif t == "beam splitter":
        cfg = props.get("splitter config")
        ratio = 0.5
        if cfg is not None:
            sc = cfg.single()
            if sc.name != "Ratio":
                raise OpmDocumentError("spectrum-driven beam splitters are not read from .opm files")
            ratio = float(sc.single())
        n.p = (ratio,)
    elif t == "ideal filter":
        ft = props.get("filter type")
        tr = 1.0
        if ft is not None:
            f = ft.single()
            if f.name != "Constant":
                raise OpmDocumentError("spectrum-driven filters are not read from .opm files")
            tr = float(f.single())
        n.p = (tr,)
    elif t == "paraxial surface":
        n.p = (_prop(props, "focal length", "Length", 0.01),)
    elif t == "reflective grating":
        n.p = (_prop(props, "line density", "LinearDensity", 1740.0e3), float(_prop(props, "diffraction order", "I32", -1)))
    elif t == "source":
        ld = props.get("light data")
        b = ld.single() if isinstance(ld, Variant) and ld.args else None  # LightDataBuilder(None): a source without light
        n.source = _source(b) if b is not None else None
        n.alignment_wavelength = _prop(props, "alignment wavelength", "LengthOption", None)
        iso = props.get("light data iso")
        if isinstance(iso, Variant) and iso.args and iso.single() is not None:
            raise OpmDocumentError("a separate isometry of the emitted light field is not supported")
    elif t == "fluence detector":
        fe = props.get("fluence estimator")
        n.fluence_estimator = fe.single().name if isinstance(fe, Variant) and fe.args else "Voronoi"  # fluence_detector/mod.rs:62
    return n


@dataclass
class OpmDocument:
    version: str
    name: str
    nodes: list = field(default_factory=list)
    edges: list = field(default_factory=list)   # (src uuid, src port, target uuid, target port, distance [m])
    ambient: tuple = ("const", 1.0)
    analyzers: list = field(default_factory=list)  # (kind, fields)

    # ---- analysis parameters
    def raytrace_config(self):
        """the first RayTrace analyzer's RayTraceConfig (analyzers/raytrace.rs:302-323), else the default"""
        from .api import RayTraceConfig
        from . import _ffi as F
        cfg = RayTraceConfig()
        for kind, f in self.analyzers:
            if kind == "RayTrace" and f:
                cfg.min_energy_per_ray = float(f["min_energy_per_ray"])
                cfg.max_number_of_bounces = int(f["max_number_of_bounces"])
                cfg.max_number_of_refractions = int(f["max_number_of_refractions"])
                cfg.missed_surface_strategy = F.MISS_STOP if f["missed_surface_strategy"].name == "Stop" else F.MISS_IGNORE
                break
        return cfg

    def source_node(self):
        src = [n for n in self.nodes if n.node_type == "source"]
        if len(src) != 1:
            raise OpmDocumentError(f"the GPU path takes exactly one source, the document has {len(src)}")
        return src[0]

    def axis_wavelength(self) -> float:
        """the source's alignment wavelength, else the energy-weighted centre wavelength of its rays (nodes/source.rs:244-256,
        rays.rs:1166-1175; the rays' energies factor into position x spectrum, so the spectrum's weights suffice)"""
        s = self.source_node()
        if s.alignment_wavelength:
            return float(s.alignment_wavelength)
        wl, w = self._spectrum(s.source)
        return float(sum(a * b for a, b in zip(wl, w)) / sum(w))

    @staticmethod
    def _spectrum(src: OpmSource):
        if src.spectrum is not None:
            from .api import spectrum_gaussian
            _, lo, hi, n, mu, fwhm, power = src.spectrum
            wl, w = spectrum_gaussian(lo, hi, n, mu, fwhm, power)
            return list(wl), list(w)
        return src.wavelengths, src.weights

    # ---- ordering: a chain from the source, side chains behind the second output of beam splitters
    def ordered(self):
        """[(node, parent index | -1, parent port, distance | None)] in an order the scene layer accepts (parents first)"""
        by_id = {n.uuid: n for n in self.nodes}
        out_edges: dict[str, list] = {}
        has_in = set()
        for (s, sp, t, tp, d) in self.edges:
            if s not in by_id or t not in by_id:
                raise OpmDocumentError("edge names an unknown node")
            if tp not in ("input_1",):
                raise OpmDocumentError(f"target port {tp!r} is not supported (second inputs of beam splitters)")
            out_edges.setdefault(s, []).append((sp, t, float(d)))
            has_in.add(t)
        src = self.source_node()
        order, index = [], {}

        def walk(uuid, parent, port, dist):
            n = by_id[uuid]
            if uuid in index:
                raise OpmDocumentError(f"node {n.name!r} has more than one incoming edge")
            index[uuid] = len(order)
            order.append((n, parent, port, dist))
            outs = out_edges.get(uuid, [])
            main = [e for e in outs if e[0] in ("output_1", "out1_trans1_refl2")]
            side = [e for e in outs if e[0] == "out2_trans2_refl1"]
            if len(main) > 1 or len(side) > 1 or len(main) + len(side) != len(outs):
                raise OpmDocumentError(f"node {n.name!r}: unsupported fan-out")
            me = index[uuid]
            # the main chain first, so that `parent = -1` (previous node) stays valid for it
            chain = [(main[0], 0)] if main else []
            pending = [(side[0], 1)] if side else []
            for (sp, t, d), p in chain:
                walk(t, me, p, d)
            for (sp, t, d), p in pending:
                walk(t, me, p, d)
        walk(src.uuid, -1, 0, None)
        unreachable = [n.name for n in self.nodes if n.uuid not in index]
        if unreachable:
            raise OpmDocumentError(f"nodes not connected to the source: {unreachable}")
        return order, index

    # ---- the scene of this package
    def build(self, ctx):
        """-> (Scene, SourceDesc | None, {uuid: node index}); nodes without an isometry are added with their edge distance and
        wait for `scene.calc_node_positions(doc.axis_wavelength())`"""
        from . import _ffi as F
        from .api import Aperture, Isometry, RefractiveIndex, SourceDesc
        from .scene import CoatingType, Node, Scene
        order, index = self.ordered()
        sc = Scene(ctx)
        sc.set_ambient(getattr(RefractiveIndex, self.ambient[0])(*self.ambient[1:]))

        def iso_of(rt):
            o = F.OpmIsometry()
            r, t = rt
            for k in range(9):
                o.fwd.rot[k] = r[k]
                o.inv.rot[k] = r[3 * (k % 3) + k // 3]
            for k in range(3):
                o.fwd.t[k] = t[k]
                o.inv.t[k] = -sum(r[3 * j + k] * t[j] for j in range(3))
            return Isometry(o)

        def ap(a):
            if a is None:
                return None
            if a[0] == "stack":
                return Aperture.stack([ap(item) for item in a[1]], obstruction=a[-1])
            return getattr(Aperture, a[0])(*a[1:-1], obstruction=a[-1])

        def coat(c):
            return None if c is None else getattr(CoatingType, c[0])(*c[1:])

        for k, (n, parent, port, dist) in enumerate(order):
            node = Node(n.kind, iso_of(n.isometry) if n.isometry is not None else None)
            for j, v in enumerate(n.p):
                node.c.p[j] = float(v)
            if n.index is not None:
                node.c.index = getattr(RefractiveIndex, n.index[0])(*n.index[1:]).c
            node.with_coatings(coat(n.coatings[0]), coat(n.coatings[1]))
            node.with_apertures(ap(n.apertures[0]), ap(n.apertures[1]))
            if n.alignment is not None:
                node.c.has_alignment = 1
                node.c.alignment = iso_of(n.alignment).c
            node.with_lidt(n.lidt)
            if k > 0:
                node.fed_by(parent, port)
            placed = n.isometry is not None or k == 0  # the source without an isometry sits at the origin
            like = None
            if not placed and n.align_like is not None:
                if n.align_like[0] not in index or index[n.align_like[0]] >= k:
                    raise OpmDocumentError(f"node {n.name!r} is aligned like a node that comes after it")
                like = (index[n.align_like[0]], n.align_like[1])
            sc.add_node(node, distance=None if placed else dist, align_like=like)
        src = self.source_node().source
        sd = None
        if src is not None:
            wl, w = self._spectrum(src)
            pos, en = src.pos, src.energy
            n_pos = pos[1] * pos[2] if pos[0] == "grid" else (1 + 3 * pos[1] * (pos[1] + 1) if pos[0] == "hexapolar" else pos[1])
            sd = SourceDesc(n_pos, wl, w, total_energy=en[1])
            if pos[0] == "grid":
                sd.grid(pos[1], pos[2], pos[3], pos[4])
            elif pos[0] == "hexapolar":
                sd.hexapolar(pos[1], pos[2])
            elif pos[0] == "fibonacci_rectangle":
                sd.fibonacci_rectangle(pos[2], pos[3])
            else:
                sd.fibonacci_ellipse(pos[2], pos[3])
            if en[0] == "general_gaussian":
                sd.general_gaussian(en[2], en[3], en[4], en[5], en[6])
        return sc, sd, index


def _port_map(v) -> dict:
    """PortMap (port_map.rs:12): a newtype around {external port name: (node uuid, internal port name)}"""
    if v is None:
        return {}
    m = v[0] if isinstance(v, tuple) and len(v) == 1 else v
    if not isinstance(m, dict):
        raise OpmDocumentError("port map of a group is not a map")
    return {str(k): (str(t[0]), str(t[1])) for k, t in m.items()}


def _flatten_graph(graph: dict, doc: "OpmDocument"):
    """The nodes and edges of an OpticGraph (nodes/node_group/optic_graph.rs:1040-1071) into the document's flat lists.  A nested
    NodeGroup (an entry with a `graph` of its own, optic_ref.rs:76-99) is taken apart: its nodes join the list, and an edge that ends
    at one of its external ports is re-attached to the inner node and port the group's input_map / output_map names -- which is what
    the reference's analysis does with the light that arrives there (node_group/analysis_raytrace.rs:29-91: `get_incoming` hands the
    external port's light to the mapped inner node, the mapped output ports carry the sub-graph's results out) and what its node
    positioning does with the edge's distance (`set_external_distances`, :96-98; optic_graph.rs:878-926: the inner node is placed that
    distance behind the predecessor OUTSIDE the group).  The group node's own isometry places nothing.  Deviation: the reference starts
    every group's positioning with the up direction +y (:102), the flat scene carries the up direction on from the outside; the two
    differ only when the beam has been turned about its own axis before it enters the group.  Refused: inverted groups, reference
    nodes."""
    groups: dict[str, tuple[dict, dict]] = {}

    def walk(g: dict):
        for entry in g.get("nodes") or []:
            a = entry.get("attributes") or entry.get("node_attr") or {}
            t = a.get("node_type")
            if t == "reference":
                raise OpmDocumentError("node type 'reference' (a node that stands for another node) is not supported")
            if t == "group":
                if bool(a.get("inverted", False)):
                    raise OpmDocumentError(f"group {a.get('name', '')!r} is inverted: not supported")
                sub = entry.get("graph")
                if not isinstance(sub, dict):
                    raise OpmDocumentError(f"group {a.get('name', '')!r} has no graph")
                groups[str(a["uuid"])] = (_port_map(sub.get("input_map")), _port_map(sub.get("output_map")))
                walk(sub)
                continue
            doc.nodes.append(_node(entry))
        for e in g.get("edges") or []:
            doc.edges.append((str(e[0]), str(e[1]), str(e[2]), str(e[3]), float(e[4])))

    walk(graph)
    if not groups:
        return
    resolved = []
    for (s, sp, t, tp, d) in doc.edges:
        for _ in range(64):  # a port of a group may itself be mapped to a port of a group nested in it
            if s not in groups:
                break
            if sp not in groups[s][1]:
                raise OpmDocumentError(f"edge leaves a group at port {sp!r}, which its output map does not name")
            s, sp = groups[s][1][sp]
        for _ in range(64):
            if t not in groups:
                break
            if tp not in groups[t][0]:
                raise OpmDocumentError(f"edge enters a group at port {tp!r}, which its input map does not name")
            t, tp = groups[t][0][tp]
        if s in groups or t in groups:
            raise OpmDocumentError("port maps of nested groups do not end at a node")
        resolved.append((s, sp, t, tp, d))
    doc.edges[:] = resolved


def loads_opm(text: str) -> OpmDocument:
    """OpmDocument::from_string (opm_document.rs:116-121)"""
    try:
        root = parse_ron(text)
    except OpmDocumentError as e:
        raise OpmDocumentError(f"parsing of model failed: {e}") from None
    if not isinstance(root, dict) or "opm_file_version" not in root:
        raise OpmDocumentError("parsing of model failed: not an OpmDocument")
    scenery = root.get("scenery") or {}
    attr = scenery.get("node_attr") or {}
    graph = scenery.get("graph") or {}
    doc = OpmDocument(version=str(root["opm_file_version"]), name=attr.get("name", ""))
    _flatten_graph(graph, doc)
    g = root.get("global") or {}
    if "ambient_refr_index" in g:
        doc.ambient = _index(g["ambient_refr_index"])
    for a in (root.get("analyzers") or {}).values():
        t = a["analyzer_type"]
        doc.analyzers.append((t.name, t.single() if t.args else None))
    return doc


def load_opm(path) -> OpmDocument:
    """OpmDocument::from_file (opm_document.rs:97-110)"""
    with open(path, "r", encoding="utf-8") as f:
        return loads_opm(f.read())
