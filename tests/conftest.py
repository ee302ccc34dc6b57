import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "gradabm-june_b200", ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


def pytest_sessionfinish(session, exitstatus):
    """Parity bookkeeping the assertions alone do not show (how many gradient components needed the fall-back
    criteria, near-tie counts, largest gaps): written next to the other GPU-run artefacts."""
    import json
    try:
        import helpers
    except Exception:  # noqa: BLE001
        return
    if not helpers._REPORT:
        return
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_report.json", "w") as f:
            json.dump(helpers._REPORT, f, indent=1, default=float)
    except OSError:
        pass
