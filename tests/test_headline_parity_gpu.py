"""GPU: parity of BOTH arithmetic modes on the benchmarked configuration itself (BASELINE.json configs[1]: 512x1024 view, 2 source
panoramas, 64 coarse + 64 fine samples, fine_depth_use_all False) — 2048 random rays of the view against the CPU oracle with the
seed-0 weights bench.py times.  Same code path as the `parity` object of the bench line (bench.py:parity_check), bounds stated there
(PARITY_BOUNDS): fp32 |a-e| <= 1e-4|e| + atol on every non-seam ray; bf16 rtol 1e-2 on the natural scale of each quantity."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def test_headline_config_parity_fp32_and_bf16():
    import bench
    import panogrf_b200 as pg
    _, _, out, idx, weights = bench.oracle_run(2048)
    par = bench.parity_check(torch, pg, torch.device("cuda:0"), out, idx, weights, bench.cfg_dict())
    print(par)
    assert par["rays"] == 2048 and par["ok"], par
    assert par["fp32_bad_frac"] <= bench.PARITY_BOUNDS["fp32_bad_frac"]
    assert par["fp32_p99_err_over_tol"] < 1.0                       # 99 % of the fp32 outputs inside rtol 1e-4
    assert par["bf16_bad_frac"] <= bench.PARITY_BOUNDS["bf16_bad_frac"]
    assert par["bf16_max_frac_of_range"] <= bench.PARITY_BOUNDS["bf16_max_frac_of_range"]
