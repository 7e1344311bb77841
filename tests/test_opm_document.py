"""`.opm` reader (opossum_b200/opm_document.py) on the CPU: the RON subset parser and the interpretation of the
serialised `OpmDocument`, pinned by the reference's own fixture files -- opossum/files_for_testing/opm/opticscenery.opm,
optic_ref.opm, incorrect_opm.opm, empty.opm, what the tests of opm_document.rs load.  They are read where they lie under
/root/reference (this container; the files are not copied into the repository, the tests that need them are skipped where the
reference is absent) -- plus a hand-written document (tests/golden/opm/) that follows the same serde rules for the node types
the fixtures do not contain."""
import math
import re
from pathlib import Path

import pytest

from opossum_b200 import opm_document as D

OPM = Path(__file__).parent / "golden" / "opm"
REF = Path("/root/reference/opossum/files_for_testing/opm")
needs_reference = pytest.mark.skipif(not REF.is_dir(), reason="the reference's fixture files are not available here")


def test_ron_subset():
    v = D.parse_ron('( a: 1, b: -2.5e-3, c: "x\\"y", d: [1, 2,], e: { "k": Some(3), "n": None }, f: V((p: inf)), g: (1, true), h: Unit, )')
    assert v["a"] == 1 and v["b"] == -2.5e-3 and v["c"] == 'x"y' and v["d"] == [1, 2]
    assert v["e"] == {"k": 3, "n": None} and v["g"] == (1, True)
    assert v["f"].name == "V" and v["f"].single() == {"p": math.inf} and v["h"] == D.Variant("Unit")
    for bad in ("", "( a: 1", "( a: 1 ) x", "( a 1 )", "@"):
        with pytest.raises(D.OpmDocumentError):
            D.parse_ron(bad)


@needs_reference
def test_reference_fixture_opticscenery():
    """files_for_testing/opm/opticscenery.opm: two dummy nodes, one edge of distance 0, vacuum, a RayTrace analyzer with the
    default configuration"""
    doc = D.load_opm(REF / "opticscenery.opm")
    assert doc.version == "0" and doc.name == "OpticScenery demo"
    assert [(n.name, n.node_type) for n in doc.nodes] == [("dummy1", "dummy"), ("dummy2", "dummy")]
    assert doc.nodes[0].uuid == "484d82d5-656a-450a-a92d-ba67a77ef27e" and doc.nodes[1].uuid == "d5230f4f-45e7-4c74-9b26-032158d26076"
    for n in doc.nodes:
        assert n.coatings == (("ideal_ar",), ("ideal_ar",)) and n.apertures == (None, None) and n.lidt == 10000.0
        assert n.isometry is None and n.alignment is None and not n.inverted
    assert doc.edges == [("484d82d5-656a-450a-a92d-ba67a77ef27e", "output_1", "d5230f4f-45e7-4c74-9b26-032158d26076", "input_1", 0.0)]
    assert doc.ambient == ("const", 1.0)
    cfg = doc.raytrace_config()
    assert (cfg.min_energy_per_ray, cfg.max_number_of_bounces, cfg.max_number_of_refractions, cfg.missed_surface_strategy) == (1e-12, 1000, 1000, 0)
    with pytest.raises(D.OpmDocumentError):  # no source in this scenery
        doc.source_node()


@needs_reference
def test_reference_fixture_optic_ref_and_broken_files():
    """optic_ref.opm is one serialised node (`OpticRef`), incorrect_opm.opm and empty.opm are the reference's negative cases
    ("parsing of model failed", opm_document.rs:116-121)"""
    n = D._node(D.parse_ron((REF / "optic_ref.opm").read_text()))
    assert (n.name, n.node_type, n.uuid, n.lidt) == ("test123", "dummy", "98248e7f-dc4c-4131-8710-f3d5be2ff087", 10000.0)
    for bad in ("incorrect_opm.opm", "empty.opm"):
        with pytest.raises(D.OpmDocumentError, match="parsing of model failed"):
            D.load_opm(REF / bad)


def test_handwritten_document_interpretation():
    doc = D.load_opm(OPM / "folded_chain_handwritten.opm")
    by = {n.name: n for n in doc.nodes}
    assert by["l1"].p == (0.515, -0.515, 0.005) and by["l1"].index[0] == "sellmeier1" and by["l1"].index[-2:] == (3e-7, 2e-6)
    assert by["l1"].coatings == (("fresnel",), ("fresnel",)) and by["l1"].apertures[0] == ("circle", 0.0035, 0.0002, 0.0, False)
    assert by["m1"].p == (math.inf,) and by["m1"].coatings[0] == ("constant_r", 0.98)
    rot, tr = by["m1"].alignment
    c, s = math.cos(math.radians(22.5)), math.sin(math.radians(22.5))
    assert tr == (0.0, 0.0, 0.0) and rot == pytest.approx((1, 0, 0, 0, c, -s, 0, s, c), abs=1e-15)
    assert by["bs"].p == (0.3,) and by["f"].p == (0.8,) and by["fd"].fluence_estimator == "Binning"
    assert by["src"].isometry[1] == (0.001, -0.002, 0.01) and by["src"].alignment_wavelength == 1.054e-6
    s = by["src"].source
    assert s.pos == ("fibonacci_ellipse", 3000, 0.004, 0.003) and s.energy[:2] == ("general_gaussian", 2.0)
    assert s.wavelengths == [1.054e-6, 5.27e-7] and s.weights == pytest.approx([0.7, 0.3])
    order, index = doc.ordered()
    assert [(n.name, p, port, d) for n, p, port, d in order] == [
        ("src", -1, 0, None), ("l1", 0, 0, 0.1), ("m1", 1, 0, 0.15), ("bs", 2, 0, 0.12), ("f", 3, 0, 0.05), ("spot", 4, 0, 0.08),
        ("fd", 5, 0, 0.01), ("fd side", 3, 1, 0.09)]
    assert doc.axis_wavelength() == 1.054e-6


def test_unsupported_content_is_refused():
    text = (OPM / "folded_chain_handwritten.opm").read_text()
    with pytest.raises(D.OpmDocumentError, match="not supported"):
        D.loads_opm(text.replace('node_type: "mirror"', 'node_type: "reference"'))
    with pytest.raises(D.OpmDocumentError, match="inverted"):
        D.loads_opm(text.replace('name: "f",', 'name: "f", ').replace("inverted: false,\n                    ),\n                ),\n                (\n                    attributes: (\n                        node_type: \"spot diagram\"", "inverted: true,\n                    ),\n                ),\n                (\n                    attributes: (\n                        node_type: \"spot diagram\""))
    with pytest.raises(D.OpmDocumentError):
        D.loads_opm(text.replace('"input_1", 0.09', '"input_2", 0.09')).ordered()


def _nested_variant(text: str, inner_again: bool = False) -> str:
    """the handwritten document with the ideal filter and the spot diagram moved into a nested NodeGroup (optic_ref.rs:76-99: an
    entry with `attributes` and a `graph` of its own), external ports "in" / "out" mapped onto them; with `inner_again` the filter
    sits in a group inside that group"""
    lines = text.split("\n")
    starts = [i for i, l in enumerate(lines) if l == "                ("]
    ends = [i for i, l in enumerate(lines) if l == "                ),"]
    assert len(starts) == len(ends) == 8
    f, spot = (lines[starts[k]: ends[k] + 1] for k in (4, 5))
    U = "00000000-0000-4000-8000-0000000000"

    def group(uuid, name, nodes, edges, imap, omap):
        head = ['(', '    attributes: (', '        node_type: "group",', f'        name: "{name}",', '        ports: (inputs: {}, outputs: {}),',
                f'        uuid: "{uuid}",', '        lidt: 10000.0,', '        props: {},', '        inverted: false,', '    ),', '    graph: (',
                '        nodes: [']
        tail = ['        ],', '        edges: [' + ", ".join(edges) + '],', '        input_map: ({' + imap + '}),', '        output_map: ({' + omap + '}),',
                '    ),', '),']
        return head + ["        " + l.strip() for l in nodes] + tail

    f_id, spot_id = U + "05", U + "06"
    if inner_again:
        f_part = group(U + "bb", "innermost", f, [], f'"a": ("{f_id}", "input_1")', f'"b": ("{f_id}", "output_1")')
        first, first_in, first_out = U + "bb", "a", "b"
    else:
        f_part, first, first_in, first_out = f, f_id, "input_1", "output_1"
    g = group(U + "aa", "inner", f_part + spot, [f'("{first}", "{first_out}", "{spot_id}", "input_1", 0.08)'],
              f'"in": ("{first}", "{first_in}")', f'"out": ("{spot_id}", "output_1")')
    out = lines[: starts[4]] + ["                " + l for l in g] + lines[ends[5] + 1:]
    t = "\n".join(out)
    t = t.replace(f'"out1_trans1_refl2", "{f_id}", "input_1", 0.05', f'"out1_trans1_refl2", "{U}aa", "in", 0.05')
    t = t.replace(f'                ("{f_id}", "output_1", "{spot_id}", "input_1", 0.08),\n', "")
    t = t.replace(f'("{spot_id}", "output_1", "{U}07", "input_1", 0.01)', f'("{U}aa", "out", "{U}07", "input_1", 0.01)')
    return t


def test_nested_groups_are_taken_apart():
    """a NodeGroup inside the scenery (and one inside that): the flattened document is the flat one -- same nodes, same chain, same
    distances; inverted groups and unmapped ports are refused"""
    text = (OPM / "folded_chain_handwritten.opm").read_text()
    flat = D.loads_opm(text)
    want = [(n.name, p, port, d) for n, p, port, d in flat.ordered()[0]]
    for again in (False, True):
        nested_text = _nested_variant(text, again)
        assert nested_text.count('node_type: "group"') == (3 if again else 2)
        doc = D.loads_opm(nested_text)
        assert sorted(n.uuid for n in doc.nodes) == sorted(n.uuid for n in flat.nodes)
        assert [(n.name, p, port, d) for n, p, port, d in doc.ordered()[0]] == want
        by, by_flat = {n.uuid: n for n in doc.nodes}, {n.uuid: n for n in flat.nodes}
        assert all(by[u] == by_flat[u] for u in by)
    nested_text = _nested_variant(text)
    with pytest.raises(D.OpmDocumentError, match="input map"):
        D.loads_opm(nested_text.replace('"in": (', '"elsewhere": ('))
    with pytest.raises(D.OpmDocumentError, match="inverted"):
        D.loads_opm(re.sub(r'(name: "inner",.*?props: \{\},\s*inverted: )false', r'\1true', nested_text, count=1, flags=re.S))


PORT = '(anchor_point_iso: (transform: (rotation: (0.0, 0.0, 0.0, 1.0), translation: (0.0, 0.0, 0.0))), %s coating: %s, lidt: 20000.0)'


def _node_text(node_type, props, coat_in="IdealAR", ap_in="", outs=("output_1",), extra=""):
    outs_text = ", ".join('"%s": %s' % (o, PORT % ("", "IdealAR")) for o in outs)
    return ('(attributes: (node_type: "%s", name: "n", ports: (inputs: {"input_1": %s}, outputs: {%s}), '
            'uuid: "00000000-0000-4000-8000-000000000001", lidt: 10000.0, props: {%s}, %s inverted: false))'
            % (node_type, PORT % (ap_in, coat_in), outs_text, props, extra))


def test_property_and_port_mapping_of_every_node_type():
    """the serialised forms of the properties each node registers (see the module docstring for the reference lines) and of the
    port fields (coating variants, the aperture configs of aperture.rs:123-340)"""
    n = D._node(D.parse_ron(_node_text("wedge", '"center thickness": Length(0.008), "wedge": Angle(0.05), "refractive index": '
                                       'RefractiveIndex(Schott((a0: 3.2, a1: -0.02, a2: 0.03, a3: 0.007, a4: -0.0009, a5: 0.00007, '
                                       'wvl_range: (start: 0.0000003, end: 0.000002))))',
                                       coat_in="ConstantR(reflectivity: 0.25)",
                                       ap_in="aperture: BinaryRectangle((width: 0.01, height: 0.02, center: (0.001, -0.001), aperture_type: Obstruction)),")))
    assert n.kind == 3 and n.p == (0.008, 0.05) and n.index == ("schott", 3.2, -0.02, 0.03, 0.007, -0.0009, 0.00007, 3e-7, 2e-6)
    assert n.coatings[0] == ("constant_r", 0.25) and n.apertures[0] == ("rect", 0.01, 0.02, 0.001, -0.001, True)
    n = D._node(D.parse_ron(_node_text("cylindric lens", '"front curvature": Length(inf), "rear curvature": Length(-0.2), "center thickness": '
                                       'Length(0.004), "refractive index": RefractiveIndex(Conrady((n0: 1.5, a: 0.000000004, b: 0.0, '
                                       'wvl_range: (start: 0.0000003, end: 0.0000025))))',
                                       ap_in="aperture: Gaussian((sigma: (0.002, 0.003), center: (0.0, 0.0), aperture_type: Hole)),")))
    assert n.kind == 2 and n.p == (math.inf, -0.2, 0.004) and n.index == ("conrady", 1.5, 4e-9, 0.0, 3e-7, 2.5e-6)
    assert n.apertures[0] == ("gaussian", 0.002, 0.003, 0.0, 0.0, False) and n.coatings[1] == ("ideal_ar",)
    n = D._node(D.parse_ron(_node_text("parabolic mirror", '"focal length": Length(0.25), "oa angle": Angle(0.7853981633974483), "collimating": '
                                       'Bool(true), "oa direction": Vec2((0.0, 1.0))')))
    assert n.kind == 5 and n.p == (0.25, 0.7853981633974483, 0.0, 1.0, 1.0)
    n = D._node(D.parse_ron(_node_text("paraxial surface", '"focal length": Length(0.1)')))
    assert n.kind == 9 and n.p == (0.1,)
    n = D._node(D.parse_ron(_node_text("reflective grating", '"line density": LinearDensity(1200000.0), "diffraction order": I32(-2)')))
    assert n.kind == 12 and n.p == (1200000.0, -2.0)
    n = D._node(D.parse_ron(_node_text("energy meter", '"meter type": Metertype(IdealEnergyMeter)')))
    assert n.kind == 16
    n = D._node(D.parse_ron(_node_text("fluence detector", "")))
    assert n.kind == 11 and n.fluence_estimator == "Voronoi"  # the reference's default estimator (fluence_detector/mod.rs:62)
    n = D._node(D.parse_ron(_node_text("dummy", "", extra='align_like_node_at_distance: Some(("00000000-0000-4000-8000-000000000009", 0.3)),')))
    assert n.align_like == ("00000000-0000-4000-8000-000000000009", 0.3)
    # PolygonConfig travels with its triangulation (aperture.rs:247-251); StackConfig as a list of apertures (:385-390)
    poly = ("BinaryPolygon((points: [(0.0, 0.0), (0.01, 0.0), (0.01, 0.01), (0.0, 0.01)], aperture_type: Hole, "
            "triangle_indices: [[3, 0, 1], [1, 2, 3]]))")
    n = D._node(D.parse_ron(_node_text("dummy", "", ap_in="aperture: %s," % poly)))
    assert n.apertures[0] == ("polygon", [[0.0, 0.01, 0.0, 0.0, 0.01, 0.0], [0.01, 0.0, 0.01, 0.01, 0.0, 0.01]], False)
    n = D._node(D.parse_ron(_node_text("dummy", "", ap_in="aperture: Stack((apertures: [BinaryCircle((radius: 0.02, center: (0.0, 0.0), "
                                       "aperture_type: Hole)), %s], aperture_type: Obstruction))," % poly)))
    assert n.apertures[0][0] == "stack" and n.apertures[0][-1] is True
    assert [a[0] for a in n.apertures[0][1]] == ["circle", "polygon"]
    with pytest.raises(D.OpmDocumentError, match="triangle"):
        D._node(D.parse_ron(_node_text("dummy", "", ap_in="aperture: %s," % poly.replace("[1, 2, 3]", "[1, 2, 7]"))))
    with pytest.raises(D.OpmDocumentError, match="flat stacks"):
        D._node(D.parse_ron(_node_text("dummy", "", ap_in="aperture: Stack((apertures: [%s, %s], aperture_type: Hole))," % (poly, poly))))
    for bad in ('"splitter config": SplitterType(Spectrum(()))',):
        with pytest.raises(D.OpmDocumentError):
            D._node(D.parse_ron(_node_text("beam splitter", bad, outs=("out1_trans1_refl2", "out2_trans2_refl1"))))


def test_source_light_data_mapping():
    """RayDataBuilder::Collimated with every distribution the device generates (position_distributions/, energy_distributions/,
    spectral_distribution/)"""
    def src(pos, energy, spec):
        return D._node(D.parse_ron(_node_text("source", '"light data": LightDataBuilder(Some(Geometric(Collimated(pos_dist: %s, energy_dist: %s, '
                                              'spect_dist: %s)))), "alignment wavelength": LengthOption(None)' % (pos, energy, spec))))
    uni, line = "Uniform((total_energy: 1.5))", "LaserLines((lines: [(0.000001, 2.0)]))"
    n = src("Grid((nr_of_points: (7, 5), side_length: (0.01, 0.02)))", uni, line)
    assert n.source.pos == ("grid", 7, 5, 0.01, 0.02) and n.source.energy == ("uniform", 1.5) and n.alignment_wavelength is None
    assert n.source.wavelengths == [1e-6] and n.source.weights == [1.0]  # weights are normalised
    assert src("Hexapolar((nr_of_rings: 4, radius: 0.005))", uni, line).source.pos == ("hexapolar", 4, 0.005)
    assert src("FibonacciRectangle((nr_of_rays: 100, side_length_x: 0.01, side_length_y: 0.02))", uni, line).source.pos == ("fibonacci_rectangle", 100, 0.01, 0.02)
    g = src("FibonacciEllipse((nr_of_rays: 10, radius_x: 0.001, radius_y: 0.002))", uni,
            "Gaussian((wvl_range: (0.00000095, 0.00000115), num_points: 64, mu: 0.000001053, fwhm: 0.00000005, power: 1.0))").source
    assert g.spectrum == ("gaussian", 9.5e-7, 1.15e-6, 64, 1.053e-6, 5e-8, 1.0)
    for bad_pos in ("Random((nr_of_points: 10, side_length_x: 0.01, side_length_y: 0.01))", "Sobol((nr_of_rays: 10))"):
        with pytest.raises(D.OpmDocumentError, match="not generated on the device"):
            src(bad_pos, uni, line)
    empty = D._node(D.parse_ron(_node_text("source", '"light data": LightDataBuilder(None)')))
    assert empty.source is None


def _to_ron(v, indent=0):
    """a serialiser for the subset, laid out like ron::ser::PrettyConfig does (one field per line, trailing commas)"""
    pad, inner = "    " * indent, "    " * (indent + 1)
    if v is None:
        return "None"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return "inf" if v == math.inf else "-inf" if v == -math.inf else repr(v)
    if isinstance(v, str):
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if isinstance(v, list):
        return "[\n" + "".join(inner + _to_ron(x, indent + 1) + ",\n" for x in v) + pad + "]" if v else "[]"
    if isinstance(v, tuple):
        return "(" + ", ".join(_to_ron(x, indent) for x in v) + ")"
    if isinstance(v, D.Variant):
        if v.fields is not None:
            return v.name + "(\n" + "".join(inner + k + ": " + _to_ron(x, indent + 1) + ",\n" for k, x in v.fields.items()) + pad + ")"
        return v.name + ("(" + ", ".join(_to_ron(x, indent) for x in v.args) + ")" if v.args else "")
    if isinstance(v, dict) and v.get("__map__"):
        items = [(k, x) for k, x in v.items() if k != "__map__"]
        return "{\n" + "".join(inner + _to_ron(k, indent + 1) + ": " + _to_ron(x, indent + 1) + ",\n" for k, x in items) + pad + "}"
    return "(\n" + "".join(inner + k + ": " + _to_ron(x, indent + 1) + ",\n" for k, x in v.items()) + pad + ")"


def test_ron_round_trip_property():
    """random nested documents survive serialise -> parse (hypothesis; finite numbers, identifiers that are no keywords)"""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    ident = st.from_regex(r"[a-z_][a-z0-9_]{0,8}", fullmatch=True).filter(lambda s: s not in ("true", "false", "inf", "e"))
    upper = st.from_regex(r"[A-Z][A-Za-z0-9]{0,8}", fullmatch=True).filter(lambda s: s not in ("Some", "None", "NaN"))
    leaf = st.one_of(st.booleans(), st.integers(-10**9, 10**9), st.floats(allow_nan=False, allow_infinity=True, width=64),
                     st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), max_size=12))

    def extend(children):
        struct = st.dictionaries(ident, children, min_size=1, max_size=4)
        return st.one_of(st.lists(children, max_size=3), struct, st.tuples(children, children),
                         st.dictionaries(st.text(alphabet="abc xyz", min_size=1, max_size=6), children, min_size=1, max_size=3)
                         .map(lambda d: {**d, "__map__": True}),
                         st.builds(lambda n, a: D.Variant(n, (a,)), upper, children), st.builds(lambda n, f: D.Variant(n, (), f), upper, struct),
                         st.builds(D.Variant, upper))
    docs = st.recursive(leaf, extend, max_leaves=25)

    def strip(v):  # what the parser returns for a map: a plain dict
        if isinstance(v, dict):
            return {k: strip(x) for k, x in v.items() if k != "__map__"}
        if isinstance(v, list):
            return [strip(x) for x in v]
        if isinstance(v, tuple):
            return tuple(strip(x) for x in v)
        if isinstance(v, D.Variant):
            return D.Variant(v.name, tuple(strip(x) for x in v.args), None if v.fields is None else strip(v.fields))
        return v

    @settings(max_examples=200, deadline=None)
    @given(docs)
    def check(doc):
        assert D.parse_ron(_to_ron(doc)) == strip(doc)
    check()
