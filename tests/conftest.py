"""Shared pytest configuration: the `gpu` marker, golden-fixture loader, repo root on sys.path."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


FLOAT_CASES = [
    ("raw", "full"), ("euclidean", "full"), ("mahalanobis", "full"),
    ("gnn", "full"), ("gnn", "reduced"), ("msn", "full"), ("msn", "reduced"),
]


def state_from_golden(g, kind="euclidean"):
    from oracle import sknnr_oracle as orc

    return orc.FittedState(
        kind=kind,
        fit_Z=g["state_fit_Z"],
        y=g["state_y"],
        center=g.get("state_center"),
        scale=g.get("state_scale"),
        proj=g.get("state_proj"),
    )


def yaimpute_weights(d):
    return 1.0 / (1.0 + d)
