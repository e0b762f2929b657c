"""The parallel formulation of the forest construction (asynchronous exact FH rounds + reservation-based min-size merge,
as implemented by stereomatch_b200/csrc/forest.cu) against the sequential reference semantics: the numpy model of the
GPU algorithm must reproduce the oracle's edge mask edge for edge (the oracle is pinned to the reference's own
segment_graph / merge loop by tests/test_oracle_vs_ref.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "models"))
from forest_model import build_forest_model  # noqa: E402

from stereomatch_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.mark.parametrize("W,H,seed,c,ms,nat,low,high", [
    (96, 64, 7, 5000.0, 200, 0, 64, 256),
    (96, 64, 7, 300.0, 20, 0, 16, 64),        # tight band: many ingest events
    (120, 80, 3, 5000.0, 200, 1, 64, 256),    # natural-looking image: big flat (w = 0) regions, long equal-weight chains
    (64, 64, 1, 50.0, 5, 0, 1 << 30, 1 << 30),  # everything live from the start
    (50, 30, 10, 1e12, 2, 0, 32, 128),        # c -> inf: Kruskal MST
    (33, 1, 11, 100.0, 2, 0, 8, 16),
])
def test_async_fh_model_matches_sequential(oracle, W, H, seed, c, ms, nat, low, high):
    L, _, _ = (synth.make_natural_pair if nat else synth.make_pair)(W, H, 16, seed=seed)
    F = oracle.forest(L, c=c, min_size=ms)
    mask, _, _, info = build_forest_model(np.asarray(F.ew), W, H, c, ms, band_low=low, band_high=high)
    assert np.array_equal(mask, np.asarray(F.mask))
    assert info["rounds_fh"] > 0
