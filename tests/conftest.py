import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cb():
    """the package (its directory name has a hyphen, so no plain import statement)"""
    return importlib.import_module("corona-13_b200")


@pytest.fixture(scope="session")
def built():
    """make sure the oracle (and, where the reference tree exists, oracle/_ref) and the product are built"""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def lib(built):
    return importlib.import_module("corona-13_b200.lib")


def pytest_terminal_summary(terminalreporter):
    """tie rate of the two traversal modes (helpers.intersect_modes): rays compared, answers that differ, all classified"""
    try:
        from helpers import TIE_LOG, MODE_B_LOG
    except Exception:
        return
    if MODE_B_LOG:
        rays = sum(n for _, n, _ in MODE_B_LOG)
        diff = sum(d for _, _, d in MODE_B_LOG)
        terminalreporter.write_line(f"mode B (GPU-built tree vs the reference tree's answers): {rays} rays, {diff} answers differ ({diff / max(rays, 1):.2e}), "
                                    "every one proven a tie / grazed box")
        for what, n, d in MODE_B_LOG:
            terminalreporter.write_line(f"  {what}: {d} of {n}")
    if not TIE_LOG:
        return
    rays = sum(n for _, n, _ in TIE_LOG)
    diff = sum(d for _, _, d in TIE_LOG)
    terminalreporter.write_line(f"WIDE8 vs EXACT4: {rays} rays compared, {diff} answers differ ({diff / max(rays, 1):.2e}), every one classified as tie / grazed box")
    for what, n, d in TIE_LOG:
        if d:
            terminalreporter.write_line(f"  {what}: {d} of {n}")
