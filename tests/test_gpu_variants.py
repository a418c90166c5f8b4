"""Every variant of the tcgen05 scan kernel (cluster size 1/2/4, merged double stage on/off, one or two
accumulators) against the oracle, whatever the planner would pick by default for the test shapes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

VARIANTS = {
    "single CTAs": {"KEMR_MMA_PAIR": "0"},
    "CTA pairs": {"KEMR_MMA_CL": "2"},
    "CTA pairs, one gallery chunk per stage": {"KEMR_MMA_CL": "2", "KEMR_MMA_NO_DS": "1"},
    "CTA pairs, one accumulator per gallery": {"KEMR_MMA_CL": "2", "KEMR_MMA_NO_MERGE": "1"},
    "quad clusters (TMA multicast)": {"KEMR_MMA_CL": "4"},
    "quad clusters, one gallery chunk per stage": {"KEMR_MMA_CL": "4", "KEMR_MMA_NO_DS": "1"},
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_scan_variant_matches_oracle(name):
    if not torch.cuda.is_available() or torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs a cc 10.x device")
    env = dict(os.environ, **VARIANTS[name])
    r = subprocess.run([sys.executable, os.path.join(HERE, "run_variant_check.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, f"{name}:\n{r.stdout}\n{r.stderr[-2000:]}"
