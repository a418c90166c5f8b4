import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_set():
    """Inputs and dense outputs of the unmodified reference on the small fixture set."""
    from knowledge_enhanced_multimodal_retrieval_b200 import synth
    z = np.load(os.path.join(GOLDEN, "small_set.npz"))
    with open(os.path.join(GOLDEN, "small_set_kg.json")) as f:
        kg = json.load(f)
    d = {k: z[k] for k in z.files}
    for name in ("query", "image", "target", "sq_query", "sq_image", "sq_target"):
        d[name] = synth.bf16_bits_to_f32(d[name + "_bits"])
    d.update(kg)
    return d


# hypothesis: the same examples on every run (the driver's CPU suite must not depend on a random seed); explore with
# `HYPOTHESIS_PROFILE=explore pytest ...` when hunting for counterexamples
try:
    from hypothesis import settings as _hyp_settings
    _hyp_settings.register_profile("repeatable", derandomize=True, deadline=None)
    _hyp_settings.register_profile("explore", deadline=None, max_examples=5000)
    _hyp_settings.load_profile(os.environ.get("HYPOTHESIS_PROFILE", "repeatable"))
except ImportError:                      # pragma: no cover
    pass
