import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# Parity numbers the tests measure (native vs fp32 oracle, fp16-eager vs fp32, ...) are collected here and printed in the terminal
# summary, so that a plain `pytest -q` run (no -s) carries them in its tail.
_REPORT = []


def report(line: str):
    _REPORT.append(line)
    print(line)


def pytest_terminal_summary(terminalreporter):
    if _REPORT:
        terminalreporter.write_sep("-", "measured parity")
        for line in _REPORT:
            terminalreporter.write_line(line)
