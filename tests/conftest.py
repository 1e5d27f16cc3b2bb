import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_pack():
    import rrtmgp_b200 as R
    return R.synthetic.make_lut_pack(seed=11, dims=R.synthetic.SMALL_DIMS)


@pytest.fixture(scope="session")
def real_pack():
    import rrtmgp_b200 as R
    return R.synthetic.make_lut_pack(seed=7)


def pytest_sessionfinish(session, exitstatus):
    """GPU parity ledger: every engine-vs-oracle comparison of the run with the achieved
    error, the threshold and which bar bound.  Written beside the profiles and into gpurun_out/ (the only
    directory that travels back from the GPU box)."""
    import json
    try:
        from helpers import LEDGER
    except Exception:
        return
    if not LEDGER:
        return
    doc = {"source": "python -m pytest tests -m gpu (tests/helpers.py gate_f32 / gate_f64)",
           "thresholds": "test/float32_consistency.jl:53-62: LW 1e-3, SW 3e-2 clear / 1.2e-1 cloudy W/m2; Float64 1e-9 relative",
           "rows": LEDGER,
           "summary": {"comparisons": len(LEDGER), "failed": sum(1 for r in LEDGER if not r["passed"]),
                       "f32_rows_relaxed": [f'{r["test"]}:{r["key"]}' for r in LEDGER if r.get("columns_excused_by_f32_oracle")]}}
    for d in ("profiles", "gpurun_out"):
        try:
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, "parity_ledger.json"), "w") as f:
                json.dump(doc, f, indent=1)
        except OSError:
            pass
