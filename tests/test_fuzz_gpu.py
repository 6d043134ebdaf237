"""Differential fuzzing (tools/fuzz_parity.py): random scenes from the coordinate classes that make the reference's
arithmetic branch — exact grid lines, exact negative integers, +-0, denormals, 2e7, repeated points, type values without
a shader arm — random frame sizes and matrices; every tap and the frame bit for bit against the oracle, in every sort /
coverage / walk mode and on the opt-in flags. A bounded sample here; the tool runs any seed range (250 seeds: all identical)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_parity as F  # noqa: E402


def test_fuzz_scenes_are_deterministic_and_in_the_loader_domain():
    import numpy as np
    a, rows_a, W, H = F.random_scene(17)
    b, rows_b, _, _ = F.random_scene(17)
    assert np.array_equal(a.pos.view(np.uint32), b.pos.view(np.uint32)) and np.array_equal(rows_a, rows_b)
    assert a.fill_rule.max() <= 1 and all((int(t) & 7) <= 4 for t in a.curve_type)


@pytest.mark.gpu
@pytest.mark.parametrize("first", [0, 120, 1000, 5200, 6080])  # (seeds 0 and 124: a curve type without a shader arm is drawn at the origin, far from its control points — exact bands must not cull it; seeds 5206 and 6084: a line with a denormal y extent has a NaN crossing parameter, whose bits are the GPU's)
def test_gpu_fuzz_against_the_oracle(first):
    for seed in range(first, first + 12):
        F.check(seed, False)


@pytest.mark.gpu
def test_gpu_survives_input_outside_the_reference_domain():
    """Non-finite and beyond-int-range coordinates: the reference's conversions are undefined there, so nothing is compared;
    the frames must come back (tools/fuzz_parity.py --survive under compute-sanitizer: 0 errors)."""
    for seed in range(2000, 2010):
        F.survive(seed)
