"""Shared helpers for the tests: golden loading, synthetic inputs, digests."""
from __future__ import annotations

import hashlib
from functools import lru_cache
from pathlib import Path

import numpy as np

from radar_point_cloud_tracking_b200 import synthetic as syn
from tests.golden.specs import CLUSTER3D_SPEC, DBSCAN_CASES, PIPE_SPEC, SWEEP_CASES, SWEEP_SPEC  # noqa: F401

GOLDEN = Path(__file__).resolve().parent / "golden"


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@lru_cache(maxsize=None)
def golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


@lru_cache(maxsize=None)
def pipe_inputs():
    spec = syn.SweepSpec(**PIPE_SPEC)
    return spec, syn.synth_echo(spec)


def canonical_partition_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Same partition up to label permutation (noise must match exactly)."""
    if a.shape != b.shape or ((a < 0) != (b < 0)).any():
        return False
    m = a >= 0
    pairs = np.unique(np.stack([a[m], b[m]], axis=1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))
