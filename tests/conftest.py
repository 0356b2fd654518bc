import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "causal-gen_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- parity report: every parity test records its MEASURED deviation next to the tolerance it asserts; the table is
# printed in the terminal summary (visible under -q) and written to gpurun_out/parity_report.txt
_PARITY_ROWS = []


def parity_report(test: str, quantity: str, measured: float, tol: float, note: str = ""):
    _PARITY_ROWS.append((test, quantity, float(measured), float(tol), note))


def pytest_terminal_summary(terminalreporter):
    if not _PARITY_ROWS:
        return
    lines = ["%-58s %-30s %12s %12s  %s" % ("test", "quantity", "measured", "tolerance", "note")]
    for t, q, m, tol, note in _PARITY_ROWS:
        lines.append("%-58s %-30s %12.4g %12.4g  %s" % (t[:58], q[:30], m, tol, note))
    terminalreporter.write_sep("=", "parity report (measured deviation vs asserted tolerance)")
    for ln in lines:
        terminalreporter.write_line(ln)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.txt"), "a") as f:
            f.write("\n".join(lines) + "\n")
