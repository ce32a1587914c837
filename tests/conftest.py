import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_terminal_summary(terminalreporter):
    """Rows of the match lists that differ from the oracle, by class (tests/parity_utils.py): `near_tie` is the class
    BASELINE's north star allows (top-2 gap below 1e-3), `near_threshold` its extension (confidence within 2e-3 of the
    threshold); `unexplained` must be 0.  Also written to gpurun_out/parity_report.json when that directory exists."""
    import collections
    import json
    import os
    from tests import parity_utils
    if not parity_utils.PARITY_LOG:
        return
    agg = collections.OrderedDict()
    for e in parity_utils.PARITY_LOG:
        a = agg.setdefault(e["test"], dict.fromkeys(("calls", "rows_compared", "identical", "near_tie", "near_threshold", "unexplained"), 0))
        a["calls"] += 1
        for k in ("rows_compared", "identical", "near_tie", "near_threshold", "unexplained"):
            a[k] += e[k]
    terminalreporter.write_line("match-list parity against the oracle (rows):")
    for t, a in agg.items():
        terminalreporter.write_line(f"  {t}: {a['rows_compared']} compared, {a['identical']} identical, {a['near_tie']} near-tie, "
                                    f"{a['near_threshold']} near-threshold, {a['unexplained']} unexplained ({a['calls']} lists)")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(agg, f, indent=1)
