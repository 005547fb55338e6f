"""pytest configuration: registers the ``gpu`` marker; puts the repo root on sys.path."""

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


import pytest


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a 10-minute ceiling when pytest-timeout is installed (thread method: the process exits even
    when it is stuck inside a CUDA call), so a kernel that never returns cannot hold a GPU box until the job's limit.
    The whole GPU suite takes about a minute."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))
