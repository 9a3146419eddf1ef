"""GPU parity on subsamples of the exact BASELINE.json workloads (configs 2-5): tests/config_parity.py does the work
(same generator and seeds as tools/run_config.py, same STATS lines on both sides, funnel + hits + strings + scores)."""
import pytest

from tests import config_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,units,profiles", [(2, 2048, 24), (3, 1536, 20), (4, 160, 20), (5, 2048, 24)])
def test_config_subsample_matches_oracle(config, units, profiles):
    rep = config_parity.verify_config(config, n_units=units, n_sel=profiles)
    assert rep["ok"] and rep["strings_identical"] and rep["full_profile_set_hits_identical"]
    assert rep["funnel_gpu"] == rep["funnel_oracle"]
    assert rep["domains"] >= 10, rep          # the survivor stages were exercised
    assert rep["max_score_diff_bits"] <= 0.01 and rep["max_lnP_rel_diff"] <= 1e-3
    if config == 3:
        assert max(rep["profile_widths_compared"]) > 2047   # config 3 as SURVEY 8(d) defines it: unclipped DNA profiles


def test_tmem_and_smem_msv_paths_agree_at_bench_size():
    rep = config_parity.verify_msv_impl(n_reads=262144)
    assert rep["ok"] and rep["funnel"]["n_past_msv"] > 1000
