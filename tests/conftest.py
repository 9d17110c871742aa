import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and the built libclover_b200.so")
    config.addinivalue_line("markers", "slow: full-size CPU oracle case (tens of seconds)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


PARITY_LOG = os.path.join(ROOT, "gpurun_out", "parity_errors.json")


def record_parity(name, data):
    """Achieved errors of a parity test, merged into gpurun_out/parity_errors.json (copied to profiles/ per round) so the
    distance to the north-star tolerances is on record, not just pass / fail."""
    import json
    os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
    try:
        allr = json.load(open(PARITY_LOG))
    except Exception:
        allr = {}
    allr[name] = data
    json.dump(allr, open(PARITY_LOG, "w"), indent=1, sort_keys=True)
