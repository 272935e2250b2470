import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


def pytest_sessionfinish(session, exitstatus):
    """GPU parity runs: dump the per-check counts (converged / capped / flagged / disagreeing solves) so the
    scope of every parity claim is on record (gpurun_out/parity_counts.json; copied to profiles/ per round)."""
    try:
        mod = sys.modules.get("test_gpu_parity") or sys.modules.get("tests.test_gpu_parity")
        counts = getattr(mod, "PARITY_COUNTS", None)
        if counts:
            import json
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "parity_counts.json"), "w") as f:
                json.dump(counts, f, indent=1)
    except Exception:
        pass
