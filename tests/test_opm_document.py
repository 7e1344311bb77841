"""`.opm` reader (opossum_b200/opm_document.py) on the CPU: the RON subset parser and the interpretation of the
serialised `OpmDocument`, pinned by the reference's own fixture files (copies of opossum/files_for_testing/opm/ in
tests/golden/opm/: opticscenery.opm, optic_ref.opm, incorrect_opm.opm, empty.opm -- what opm_document.rs:300-420 tests load),
plus a hand-written document that follows the same serde rules for the node types the fixtures do not contain."""
import math
from pathlib import Path

import pytest

from opossum_b200 import opm_document as D

OPM = Path(__file__).parent / "golden" / "opm"


def test_ron_subset():
    v = D.parse_ron('( a: 1, b: -2.5e-3, c: "x\\"y", d: [1, 2,], e: { "k": Some(3), "n": None }, f: V((p: inf)), g: (1, true), h: Unit, )')
    assert v["a"] == 1 and v["b"] == -2.5e-3 and v["c"] == 'x"y' and v["d"] == [1, 2]
    assert v["e"] == {"k": 3, "n": None} and v["g"] == (1, True)
    assert v["f"].name == "V" and v["f"].single() == {"p": math.inf} and v["h"] == D.Variant("Unit")
    for bad in ("", "( a: 1", "( a: 1 ) x", "( a 1 )", "@"):
        with pytest.raises(D.OpmDocumentError):
            D.parse_ron(bad)


def test_reference_fixture_opticscenery():
    """files_for_testing/opm/opticscenery.opm: two dummy nodes, one edge of distance 0, vacuum, a RayTrace analyzer with the
    default configuration"""
    doc = D.load_opm(OPM / "opticscenery.opm")
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


def test_reference_fixture_optic_ref_and_broken_files():
    """optic_ref.opm is one serialised node (`OpticRef`), incorrect_opm.opm and empty.opm are the reference's negative cases
    ("parsing of model failed", opm_document.rs:116-121)"""
    n = D._node(D.parse_ron((OPM / "optic_ref.opm").read_text()))
    assert (n.name, n.node_type, n.uuid, n.lidt) == ("test123", "dummy", "98248e7f-dc4c-4131-8710-f3d5be2ff087", 10000.0)
    for bad in ("incorrect_opm.opm", "empty.opm"):
        with pytest.raises(D.OpmDocumentError, match="parsing of model failed"):
            D.load_opm(OPM / bad)


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
