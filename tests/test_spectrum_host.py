"""Host-side Spectrum operations (opossum/src/spectrum.rs): the reference's own tests as known answers for the oracle
(oracle/opm_spectrum.c) AND for the library's host helpers (opossum_b200/csrc/opm_spectrum_host.cpp, through the C ABI), then the
two against each other bit for bit on seeded spectra.  No device work: runs without a GPU."""
from pathlib import Path

import numpy as np
import pytest

import opossum_b200 as ob
from oracle import oracle as orc

UM, NM = 1e-6, 1e-9
REF_CSV = Path("/root/reference/opossum/files_for_testing/spectrum")


class _Orc:
    new = staticmethod(orc.spectrum_new)
    add_single_peak = staticmethod(orc.spectrum_add_single_peak)
    from_laser_lines = staticmethod(orc.spectrum_from_laser_lines)
    add_lorentzian_peak = staticmethod(orc.spectrum_add_lorentzian_peak)
    scale_vertical = staticmethod(orc.spectrum_scale_vertical)
    resample = staticmethod(orc.spectrum_resample)
    filter = staticmethod(orc.spectrum_filter)
    add = staticmethod(orc.spectrum_add)
    sub = staticmethod(orc.spectrum_sub)
    split = staticmethod(orc.spectrum_split_by_spectrum)
    average_resolution = staticmethod(orc.spectrum_average_resolution)
    merge = staticmethod(orc.spectrum_merge)
    from_csv = staticmethod(orc.spectrum_from_csv)
    generate_filter = staticmethod(orc.spectrum_generate_filter)
    get_value = staticmethod(orc.spectrum_get_value)
    total_energy = staticmethod(orc.spectrum_total_energy)
    error = orc.OracleError if hasattr(orc, "OracleError") else Exception


class _Lib:
    new = staticmethod(ob.spectrum_new)
    add_single_peak = staticmethod(ob.spectrum_add_single_peak)
    from_laser_lines = staticmethod(ob.spectrum_from_laser_lines)
    add_lorentzian_peak = staticmethod(ob.spectrum_add_lorentzian_peak)
    scale_vertical = staticmethod(ob.spectrum_scale_vertical)
    resample = staticmethod(ob.spectrum_resample)
    filter = staticmethod(ob.spectrum_filter)
    add = staticmethod(ob.spectrum_add)
    sub = staticmethod(ob.spectrum_sub)
    split = staticmethod(ob.spectrum_split_by_spectrum)
    average_resolution = staticmethod(ob.spectrum_average_resolution)
    merge = staticmethod(ob.spectrum_merge)
    from_csv = staticmethod(ob.spectrum_from_csv)
    generate_filter = staticmethod(ob.spectrum_generate_filter)
    get_value = staticmethod(ob.spectrum_get_value)
    total_energy = staticmethod(ob.spectrum_total_energy)
    error = Exception


IMPLS = [pytest.param(_Orc, id="oracle"), pytest.param(_Lib, id="library")]


def _prep(S):  # spectrum.rs:657-659
    return S.new(1.0 * UM, 4.0 * UM, 0.5 * UM)


@pytest.mark.parametrize("S", IMPLS)
def test_new_range_resolution_kats(S):
    """spectrum.rs:661-675 (new), :758-799 (helper spectra, wrong parameters, range, estimate_resolution)"""
    lam, val = _prep(S)
    assert list(lam) == [1.0, 1.5, 2.0, 2.5, 3.0, 3.5] and not val.any()
    assert S.average_resolution(lam) / UM == 0.5
    vis, _ = S.new(380.0 * NM, 750.0 * NM, 0.1 * NM)  # create_visible_spec
    assert vis[0] == 0.38 and abs(vis[-1] - 0.7499) < 1e-12
    nir, _ = S.new(800.0 * NM, 2500.0 * NM, 0.1 * NM)  # create_nir_spec
    assert nir[0] == 0.8
    for bad in ((1.0 * UM, 4.0 * UM, -0.5 * UM), (4.0 * UM, 1.0 * UM, 0.5 * UM), (-1.0 * UM, 4.0 * UM, 0.5 * UM)):
        with pytest.raises(S.error):
            S.new(*bad)


@pytest.mark.parametrize("S", IMPLS)
def test_single_peak_and_laser_lines_kats(S):
    """spectrum.rs:726-757 (from_laser_lines_*), :800-843 (set_single_peak*), :875-886 (total_energy*), :935-938 (he_ne_spectrum)"""
    lam, val = _prep(S)
    S.add_single_peak(lam, val, 2.0 * UM, 1.0)
    assert val[2] == 2.0 and S.total_energy(lam, val) == 1.0
    lam, val = _prep(S)
    S.add_single_peak(lam, val, 2.25 * UM, 1.0)
    assert val[2] == 1.0 and val[3] == 1.0
    S.add_single_peak(lam, val, 0.5 * UM, 1.0)  # outside the range: warning, unmodified
    S.add_single_peak(lam, val, 4.0 * UM, 1.0)
    assert list(val) == [0.0, 0.0, 1.0, 1.0, 0.0, 0.0]
    with pytest.raises(S.error):
        S.add_single_peak(lam, val, 1.5 * UM, -1.0)
    lam, val = S.new(1.0 * UM, 4.0 * UM, 1.0 * UM)
    S.add_single_peak(lam, val, 1.5 * UM, 1.0)
    assert S.total_energy(lam, val) == 1.0
    lam, val = S.new(380.0 * NM, 750.0 * NM, 0.1 * NM)  # create_he_ne_spec(1.0)
    S.add_single_peak(lam, val, 632.816 * NM, 1.0)
    assert S.total_energy(lam, val) == 1.0
    lam, val = S.from_laser_lines([1000.0 * NM], [1.0], 1.0 * NM)  # :726-734
    assert len(lam) == 2 and S.total_energy(lam, val) == 1.0
    lam, val = S.from_laser_lines([1000.0 * NM, 1005.0 * NM], [1.0, 0.5], 1.0 * NM)  # :736-756
    assert len(lam) == 7 and abs(S.total_energy(lam, val) - 1.5) < 1e-12
    with pytest.raises(S.error):
        S.from_laser_lines([], [], 1.0 * NM)
    with pytest.raises(S.error):
        S.from_laser_lines([1000.0 * NM], [1.0], 0.0)


@pytest.mark.parametrize("S", IMPLS)
def test_lorentzian_and_scale_kats(S):
    """spectrum.rs:850-873 (add_lorentzian*), :768-774 (nd_glass_spec), :911-944 (scale_vertical*)"""
    lam, val = S.new(1.0 * UM, 50.0 * UM, 0.1 * UM)
    S.add_lorentzian_peak(lam, val, 25.0 * UM, 0.5 * UM, 2.0)
    assert abs(S.total_energy(lam, val) - 2.0) < 0.1
    lam, val = _prep(S)
    for bad in ((-5.0 * UM, 0.5 * UM, 2.0), (2.0 * UM, -0.5 * UM, 2.0), (2.0 * UM, 0.5 * UM, -2.0)):
        with pytest.raises(S.error):
            S.add_lorentzian_peak(lam, val, *bad)
    lam, val = S.new(800.0 * NM, 2500.0 * NM, 0.1 * NM)  # create_nd_glass_spec
    S.add_lorentzian_peak(lam, val, 1054.0 * NM, 0.5 * NM, 1.0)
    assert lam[0] == 0.8 and 0.9 < S.total_energy(lam, val) < 1.0
    lam, val = S.new(1.0 * UM, 5.0 * UM, 1.0 * UM)
    S.add_single_peak(lam, val, 2.5 * UM, 1.0)
    S.scale_vertical(val, 0.5)
    assert list(val) == [0.0, 0.25, 0.25, 0.0]
    with pytest.raises(S.error):
        S.scale_vertical(val, -0.5)
    a_l, a = S.new(380.0 * NM, 750.0 * NM, 0.1 * NM)  # scale_vertical2: he_ne(1.0) * 0.6 == he_ne(0.6)
    b_l, b = S.new(380.0 * NM, 750.0 * NM, 0.1 * NM)
    S.add_single_peak(a_l, a, 632.816 * NM, 1.0)
    S.add_single_peak(b_l, b, 632.816 * NM, 0.6)
    S.scale_vertical(a, 0.6)
    assert S.total_energy(a_l, a) == S.total_energy(b_l, b)


def test_calc_ratio_kats():
    """spectrum.rs:946-958"""
    f = orc.spectrum_calc_ratio
    cases = [((1, 2, 3, 4), 0.0), ((1, 4, 2, 3), 1.0), ((2, 3, 0, 4), 0.25), ((0, 1, 0, 2), 0.5), ((1, 2, 0, 2), 0.5),
             ((0, 2, 1, 3), 0.5), ((0, 2, 1, 2), 1.0), ((2, 4, 1, 3), 0.5), ((1, 4, 1, 3), 1.0), ((1, 2, 1, 2), 1.0)]
    for args, want in cases:
        assert f(*map(float, args)) == want, args


@pytest.mark.parametrize("S", IMPLS)
def test_resample_kats(S):
    """spectrum.rs:959-1016"""
    l1, v1 = S.new(1.0 * UM, 5.0 * UM, 1.0 * UM)
    l2, v2 = S.new(1.0 * UM, 5.0 * UM, 1.0 * UM)
    S.add_single_peak(l2, v2, 2.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert np.array_equal(v1, v2) and S.total_energy(l1, v1) == S.total_energy(l2, v2)
    l1, v1 = S.new(1.0 * UM, 5.0 * UM, 1.0 * UM)  # resample_delete_old_data
    S.add_single_peak(l1, v1, 3.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert np.array_equal(v1, v2)
    l1, v1 = S.new(1.0 * UM, 5.0 * UM, 0.5 * UM)  # resample_interp
    l2, v2 = S.new(1.0 * UM, 6.0 * UM, 1.0 * UM)
    S.add_single_peak(l2, v2, 2.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert S.total_energy(l1, v1) == S.total_energy(l2, v2)
    assert np.allclose(v1, [0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0], rtol=0, atol=2.3e-16)
    l1, v1 = S.new(1.0 * UM, 5.0 * UM, 1.0 * UM)  # resample_interp2
    l2, v2 = S.new(1.0 * UM, 6.0 * UM, 0.5 * UM)
    S.add_single_peak(l2, v2, 2.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert list(v1) == [0.0, 1.0, 0.0, 0.0] and S.total_energy(l1, v1) == S.total_energy(l2, v2)
    l1, v1 = S.new(1.0 * UM, 4.0 * UM, 1.0 * UM)  # resample_right_outside
    l2, v2 = S.new(4.0 * UM, 6.0 * UM, 1.0 * UM)
    S.add_single_peak(l2, v2, 4.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert list(v1) == [0.0, 0.0, 0.0] and S.total_energy(l1, v1) == 0.0
    l1, v1 = S.new(4.0 * UM, 6.0 * UM, 1.0 * UM)  # resample_left_outside
    l2, v2 = S.new(1.0 * UM, 4.0 * UM, 1.0 * UM)
    S.add_single_peak(l2, v2, 2.0 * UM, 1.0)
    S.resample(l1, v1, l2, v2)
    assert list(v1) == [0.0, 0.0] and S.total_energy(l1, v1) == 0.0


@pytest.mark.parametrize("S", IMPLS)
def test_add_sub_kats(S):
    """spectrum.rs:1017-1034"""
    for op, want in ((S.add, [0.0, 1.0, 1.5, 0.5, 0.0, 0.0]), (S.sub, [0.0, 1.0, 0.5, 0.0, 0.0, 0.0])):
        l, v = _prep(S)
        S.add_single_peak(l, v, 1.75 * UM, 1.0)
        l2, v2 = _prep(S)
        S.add_single_peak(l2, v2, 2.25 * UM, 0.5)
        op(l, v, l2, v2)
        assert list(v) == want


CSV_OK = "\n500;5.00E+01\n501;4.981E+01\n502;4.982E+01\n503;4.984E+01\n504;4.996E+01\n505;5.010E+01\n"


@pytest.mark.parametrize("S", IMPLS)
def test_from_csv_kats(S, tmp_path):
    """spectrum.rs:677-724 on files with the content of the reference's fixtures (files_for_testing/spectrum/spec_to_csv_test_0*.csv)"""
    files = {"ok": CSV_OK, "empty": "\n\n", "short": "500;5.010E+01\n501\n", "text": "500;5.010E+01\n501;ABC", "crlf": CSV_OK.replace("\n", "\r\n")}
    for k, txt in files.items():
        (tmp_path / f"{k}.csv").write_text(txt)
    for k in ("ok", "crlf"):
        lam, val = S.from_csv(tmp_path / f"{k}.csv")
        assert np.allclose(lam, [0.500, 0.501, 0.502, 0.503, 0.504, 0.505], rtol=0, atol=2.3e-16)
        assert np.allclose(val, [0.5, 0.4981, 0.4982, 0.4984, 0.4996, 0.5010], rtol=0, atol=2.3e-16)
    for k in ("empty", "short", "text", "missing"):
        with pytest.raises(S.error):
            S.from_csv(tmp_path / f"{k}.csv")


@pytest.mark.skipif(not REF_CSV.exists(), reason="the reference's fixture files are not on this machine")
def test_from_csv_on_the_reference_fixtures():
    for S in (_Orc, _Lib):
        lam, val = S.from_csv(REF_CSV / "spec_to_csv_test_01.csv")
        assert np.allclose(val, [0.5, 0.4981, 0.4982, 0.4984, 0.4996, 0.5010], rtol=0, atol=2.3e-16)
        for k in (2, 3, 4):
            with pytest.raises(S.error):
                S.from_csv(REF_CSV / f"spec_to_csv_test_0{k}.csv")
    big_o, big_l = _Orc.from_csv(REF_CSV / "NE03B.csv"), _Lib.from_csv(REF_CSV / "NE03B.csv")  # a Thorlabs transmission export
    assert len(big_o[0]) > 100 and np.array_equal(big_o[0], big_l[0]) and np.array_equal(big_o[1], big_l[1])


def _random_spectrum(rng, irregular):
    n = int(rng.integers(2, 40))
    lo = rng.uniform(0.3, 2.0)
    if irregular:
        lam = lo + np.cumsum(rng.uniform(0.01, 0.3, n))
    else:
        lam = lo + np.arange(n) * rng.uniform(0.01, 0.3)
    return np.ascontiguousarray(lam), np.ascontiguousarray(rng.uniform(0.0, 3.0, n))


def test_library_matches_oracle_bit_for_bit():
    """seeded spectra with overlapping, nested, disjoint and irregular slot lists through every operation: identical bits"""
    rng = np.random.default_rng(20260)
    for case in range(300):
        l1, v1 = _random_spectrum(rng, case % 3 == 0)
        l2, v2 = _random_spectrum(rng, case % 5 == 0)
        for op_o, op_l in ((orc.spectrum_resample, ob.spectrum_resample), (orc.spectrum_filter, ob.spectrum_filter),
                           (orc.spectrum_add, ob.spectrum_add), (orc.spectrum_sub, ob.spectrum_sub)):
            a, b = v1.copy(), v1.copy()
            op_o(l1, a, l2, v2)
            op_l(l1, b, l2, v2)
            assert np.array_equal(a, b, equal_nan=True), (case, op_o.__name__)
        a, b = v1.copy(), v1.copy()
        sa, sb = orc.spectrum_split_by_spectrum(l1, a, l2, v2), ob.spectrum_split_by_spectrum(l1, b, l2, v2)
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(sa, sb, equal_nan=True)
        a, b = v1.copy(), v1.copy()
        c, w, e = rng.uniform(0.2, 3.0) * UM, rng.uniform(0.0, 0.2) * UM, rng.uniform(0.0, 5.0)
        orc.spectrum_add_lorentzian_peak(l1, a, c, w, e)
        ob.spectrum_add_lorentzian_peak(l1, b, c, w, e)
        assert np.array_equal(a, b, equal_nan=True)
        orc.spectrum_add_single_peak(l1, a, c, e)
        ob.spectrum_add_single_peak(l1, b, c, e)
        assert np.array_equal(a, b, equal_nan=True)
        assert orc.spectrum_average_resolution(l1) == ob.spectrum_average_resolution(l1)
        mo, ml = orc.spectrum_merge(l1, v1, l2, v2), ob.spectrum_merge(l1, v1, l2, v2)
        assert np.array_equal(mo[0], ml[0]) and np.array_equal(mo[1], ml[1], equal_nan=True)
        wl = rng.uniform(0.3, 3.0, int(rng.integers(1, 6))) * UM
        en = rng.uniform(0.0, 2.0, len(wl))
        res = rng.uniform(0.5, 20.0) * NM
        fo, fl = orc.spectrum_from_laser_lines(wl, en, res), ob.spectrum_from_laser_lines(wl, en, res)
        assert np.array_equal(fo[0], fl[0]) and np.array_equal(fo[1], fl[1])


def test_merge_spans_both_spectra_at_the_finer_resolution():
    """merge_spectra (spectrum.rs:622-650): no test in the reference -- checked against its definition"""
    l1, v1 = ob.spectrum_new(1.0 * UM, 2.0 * UM, 0.1 * UM)
    l2, v2 = ob.spectrum_new(1.5 * UM, 4.0 * UM, 0.5 * UM)
    ob.spectrum_add_single_peak(l1, v1, 1.25 * UM, 1.0)
    ob.spectrum_add_single_peak(l2, v2, 3.0 * UM, 2.0)
    lam, val = ob.spectrum_merge(l1, v1, l2, v2)
    assert lam[0] == 1.0 and abs(lam[-1] - 3.4) < 1e-9 and abs((lam[1] - lam[0]) - 0.1) < 1e-12
    # range() ends at the last SLOT (3.5 um), Spectrum::new leaves out the end point: the merged buckets stop at 3.4 um and the
    # 0.1 um of the second spectrum's 4 / um line beyond it (0.4 of its 2.0) is dropped -- the reference's behaviour
    assert abs(ob.spectrum_total_energy(lam, val) - 2.6) < 1e-9
    lo, vo = orc.spectrum_merge(l1, v1, l2, v2)
    assert np.array_equal(lo, lam) and np.array_equal(vo, val)


@pytest.mark.parametrize("S", IMPLS)
def test_generate_filter_spectrum_kats(S):
    """spectrum_helper.rs:203-366: short / long pass, step and smooth"""
    S.generate_filter("short_pass_step", 1.0 * UM, 5.0 * UM, 1.0 * UM, 7.0 * UM)  # cut-off outside the range: a warning, not an error
    lam, val = S.generate_filter("short_pass_step", 1.0 * UM, 5.0 * UM, 1.0 * UM, 3.0 * UM)
    assert [S.get_value(lam, val, w * UM) for w in (1.0, 2.0, 3.0, 4.0)] == [1.0, 1.0, 0.0, 0.0]
    lam, val = S.generate_filter("long_pass_step", 1.0 * UM, 5.0 * UM, 1.0 * UM, 3.0 * UM)
    assert [S.get_value(lam, val, w * UM) for w in (1.0, 2.0, 3.0, 4.0)] == [0.0, 0.0, 0.0, 1.0]
    for kind in ("short_pass_smooth", "long_pass_smooth"):
        for bad in (0.0, -1.0 * UM):
            with pytest.raises(S.error):
                S.generate_filter(kind, 1.0 * UM, 5.0 * UM, 0.5 * UM, 3.0 * UM, bad)
    lam, val = S.generate_filter("short_pass_smooth", 1.0 * UM, 5.0 * UM, 0.5 * UM, 3.0 * UM, 1.0 * UM)
    assert [S.get_value(lam, val, w * UM) for w in (1.0, 2.0, 2.5, 3.0)] == [1.0, 1.0, 1.0, 0.5]
    assert S.get_value(lam, val, 4.0 * UM) == 0.0 and abs(S.get_value(lam, val, 3.5 * UM)) < 1e-15  # 0.5 - 0.5 sin(pi / 2)
    lam, val = S.generate_filter("long_pass_smooth", 1.0 * UM, 5.0 * UM, 0.5 * UM, 3.0 * UM, 1.0 * UM)
    assert [S.get_value(lam, val, w * UM) for w in (1.0, 2.0, 2.5, 3.0)] == [0.0, 0.0, 0.0, 0.5]
    assert S.get_value(lam, val, 4.0 * UM) == 1.0 and abs(S.get_value(lam, val, 3.5 * UM) - 1.0) < 1e-15


def test_generate_filter_library_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(100):
        kind = list(orc.FILTER_KINDS)[int(rng.integers(0, 4))]
        a = rng.uniform(0.3, 1.0) * UM
        args = (a, a + rng.uniform(0.2, 2.0) * UM, rng.uniform(0.5, 30.0) * NM, rng.uniform(0.2, 3.2) * UM, rng.uniform(1.0, 400.0) * NM)
        lo, vo = orc.spectrum_generate_filter(kind, *args)
        ll, vl = ob.spectrum_generate_filter(kind, *args)
        assert np.array_equal(lo, ll) and np.array_equal(vo, vl), (kind, args)
