"""CPU: the counter-based sequence simulator's NumPy twin (the device kernel K-s produces the same bytes; checked
in tests/test_gpu_chain.py).  Reference behaviour it stands in for: simulate_sequences (src/_sim_seq.py:431-504) --
stationary frequencies at the tips and JC path distances are recovered (the reference's own simulator tests,
tests/test_sim_seq.py:232-250, check the same two properties)."""
import math

import numpy as np

from phynetpy_b200 import simulate as S
from phynetpy_b200.tree import FlatTree


def test_counter_generator_is_a_function_of_its_key_only():
    u1 = S.counter_uniform(5, 3, np.arange(1000), 7)
    u2 = S.counter_uniform(5, 3, np.arange(500, 1000), 7)
    assert np.array_equal(u1[500:], u2)
    assert 0.0 <= u1.min() and u1.max() < 1.0 and abs(u1.mean() - 0.5) < 0.03
    assert not np.array_equal(u1, S.counter_uniform(5, 4, np.arange(1000), 7))
    assert not np.array_equal(u1, S.counter_uniform(6, 3, np.arange(1000), 7))
    assert not np.array_equal(u1, S.counter_uniform(5, 3, np.arange(1000), 8))


def test_twin_recovers_stationary_frequencies_and_jc_distance():
    t = 0.15
    two = FlatTree([2, 2, -1], [t, t, float("nan")], ["a", "b", None])
    aln = S.simulate_alignments_counter([two], 200_000, [0.25] * 4, [1.0] * 6, 11, [0])[0]
    p = float((aln[0] != aln[1]).mean())
    d = -0.75 * math.log(1.0 - 4.0 * p / 3.0)          # JC69 distance estimate of the path a-b (2t)
    assert abs(d - 2 * t) < 0.01
    rng = np.random.default_rng(3)
    par, bl = S.random_trees(6, 10, rng)
    alns = S.simulate_alignments_counter(S.flat_trees(par, bl), 20_000, S.DEFAULT_PI, S.DEFAULT_RATES, 4, list(range(6)))
    allb = np.concatenate([a.ravel() for a in alns])
    freq = np.array([(allb == c).mean() for c in b"ACGT"])
    assert np.allclose(freq, S.DEFAULT_PI, atol=0.01)
    again = S.simulate_alignments_counter(S.flat_trees(par, bl)[2:3], 20_000, S.DEFAULT_PI, S.DEFAULT_RATES, 4, [2])[0]
    assert np.array_equal(again, alns[2])
