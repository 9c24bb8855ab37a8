"""Shared test plumbing.

* registers the ``gpu`` marker (tests that need a B200; everything else runs on CPU);
* puts the repo root on sys.path (for ``oracle``) and the product package directory
  ``normalizing-flow-with-diffusion-prior-model_b200/`` on sys.path so that the drop-in
  ``normalizing_flow`` package mirror is importable under the reference's own name.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):        # tests/: helper modules (split_pairs)
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _optional_eager_training_chains(monkeypatch):
    """The captured training chains (normalizing_flow/_train.py) are the default; NFDPM_TEST_TRAIN_GRAPHS=0 runs EVERY test
    with the eager kernel chain instead (both must stay green)."""
    if os.environ.get("NFDPM_TEST_TRAIN_GRAPHS", "1") == "0":
        monkeypatch.setenv("NFDPM_TRAIN_GRAPHS", "0")
