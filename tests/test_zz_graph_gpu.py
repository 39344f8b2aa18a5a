"""CUDA-graph replay of the training step (M1.train_step after its warm-up) against the eager launch sequence
from the same model state and Philox step: tools/check_graph.py as a test."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_graph_replay_equals_eager_step(ctx):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'tools'))
    import check_graph
    prev = os.environ.get("M1_CUDA_GRAPH")
    try:
        check_graph.main()
    finally:
        if prev is None:
            os.environ.pop("M1_CUDA_GRAPH", None)
        else:
            os.environ["M1_CUDA_GRAPH"] = prev
