"""The committed golden outputs of one tiny animated scene (tests/golden/oracle_frame_tiny.npz, written by
tests/golden/make_golden_frame.py from the CPU oracle). CPU suite: the oracle still reproduces them, byte for byte. GPU
suite: the CUDA path, through the C-ABI, reproduces the committed bytes as well — lists, attributes, counters, light map,
cube maps + depths, composited frame, TAA image and RGBA8 back buffer, with a plain and a work-graph frame in the sequence."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "oracle_frame_tiny.npz")


def _check(got):
    want = np.load(GOLD)
    assert sorted(want.files) == sorted(got)
    for k in want.files:
        assert np.array_equal(np.asarray(got[k]), want[k]), k
    assert want["counters"][1] > 0 and len(want["cube_volumes"]) > 0


def test_oracle_reproduces_golden_frame(oracle_lib):
    from make_golden_frame import KW, render
    from oracle_binding import OracleCaster
    _check(render(OracleCaster(filter_model=1, **KW)))


@pytest.mark.gpu
def test_product_reproduces_golden_frame():
    from make_golden_frame import KW, render
    from multivolumes_b200 import MultiRayCaster
    _check(render(MultiRayCaster(**KW)))
