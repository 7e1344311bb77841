"""BASELINE.json configs at their FULL sizes on the GPU, checked through size-independent properties (the per-ray
parity of the same scenes against the oracle runs at reduced ray counts in test_gpu_scenes.py / test_golden.py):
repeatable and build-independent counts, energy balances, mirror symmetry of the 4096x4096 map, ghost-focus bounce
bookkeeping.  Harness: profiles/run_configs.py (also used to produce profiles/r1_configs.jsonl)."""
import importlib.util
from pathlib import Path

import pytest

import opossum_b200 as ob

pytestmark = pytest.mark.gpu

spec = importlib.util.spec_from_file_location("run_configs", Path(__file__).resolve().parents[1] / "profiles" / "run_configs.py")
rc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rc)


@pytest.fixture(scope="module")
def ctx():
    c = ob.Context(0)
    yield c
    c.close()


def test_c1_single_lens_energy_balance(ctx):
    out = rc.c1(ctx)
    assert out["interactions"] == 40000
    assert out["checks"]["energy_within_1e-3_of_estimate"] and out["checks"]["map_energy_equals_hit_energy"]


def test_c2_laser_chain_1e7_rays(ctx):
    out = rc.c2(ctx)
    ck = out["checks"]
    assert out["interactions"] == 121426460  # 10^7 seeds through 13 surface passes minus the rays the first mirror clips
    assert ck["repeatable_counts"] and ck["strict_build_same_counts"]
    assert ck["valid_out"] == ck["alive_out"] == 9142646
    # the split-off arm skips two Fresnel surfaces (T ~ 0.92) and the 0.9 filter of the main arm: 1 / (0.92 * 0.9)
    assert ck["side_over_main_energy"] == pytest.approx(1.0 / (0.92 * 0.9), rel=5e-3)


def test_c3_ghost_focus_1e6_seeds_orders_0_to_3(ctx):
    out = rc.c3(ctx)
    p = out["passes"]
    assert len(p) == 4 and out["checks"]["all_seed_rays_valid"] and out["checks"]["transmitted_below_bound"]
    assert p[0]["bundles"] == 1 and p[1]["bundles"] == 4 and p[1]["valid_per_bounce"][1] == 4_000_000  # one child per surface
    # energy bookkeeping of pass 0: transmitted seed energy + the four reflected bundles = source energy (no aperture)
    assert p[0]["energy_per_bounce"][0] + p[1]["energy_per_bounce"][1] == pytest.approx(1.0, abs=1e-2)
    assert p[0]["energy_per_bounce"][0] + p[1]["energy_per_bounce"][1] <= 1.0 + 1e-12
    for q in p:  # bounce orders above max_bounces never survive the limits (analysis_ghostfocus.rs:14-20)
        assert q["valid_per_bounce"][4] == 0 and q["valid_per_bounce"][5] == 0


def test_c4_fluence_4096_map_1e8_rays(ctx):
    out = rc.c4(ctx)
    ck = out["checks"]
    assert out["interactions"] == 300_000_000
    assert ck["symmetric_within_1e-9"], ck  # symmetric grid source + centred lens: the map is mirror symmetric in x and y
    assert ck["total_energy_J"] == pytest.approx(0.91992, rel=1e-4)  # (1 - R)^2 of two Fresnel surfaces near normal incidence


def test_c5_broadband_streamed_1e8_rays(ctx):
    out = rc.c5(ctx, 1_562_500)
    ck = out["checks"]
    assert out["rays"] == 100_000_000 and ck["status_bits"] == 0 and ck["map_energy_below_source"]
    l, r, t, b = ck["lrtb"]
    assert l < r and b < t
