"""Delay pattern: oracle and the package's Pattern API against fixtures produced by the imported reference
(tests/golden/make_golden.py) and the reference docstring example (codebooks_patterns.py:307-316)."""
import os

import numpy as np
import pytest
import torch

from oracle import patterns_oracle
from voicecraft_b200.codebooks_patterns import DelayedPatternProvider


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "patterns.npz"))
    i = 0
    while f"z{i}" in g:
        yield {k: g[f"{k}{i}"] for k in ("z", "values", "indexes", "mask", "rvalues", "rindexes", "rmask")}
        i += 1


def test_docstring_example():
    z = torch.tensor([[[1, 2, 3, 4]] * 3])
    v, idx, m = DelayedPatternProvider(3).get_pattern(4).build_pattern_sequence(z, -7)
    exp = [[-7, 1, 2, 3, 4, -7, -7], [-7, -7, 1, 2, 3, 4, -7], [-7, -7, -7, 1, 2, 3, 4]]
    assert v[0].tolist() == exp
    assert np.array_equal(patterns_oracle.build_pattern_sequence(z.numpy(), -7)[0][0], np.array(exp))


def test_oracle_matches_reference_goldens(golden_dir):
    n = 0
    for c in _cases(golden_dir):
        v, i, m = patterns_oracle.build_pattern_sequence(c["z"], 2048)
        assert np.array_equal(v, c["values"]) and np.array_equal(i, c["indexes"]) and np.array_equal(m, c["mask"])
        assert np.array_equal(patterns_oracle.delay_closed_form(c["z"], 2048), c["values"])
        rv, ri, rm = patterns_oracle.revert_pattern_sequence(v, 2048, c["z"].shape[2])
        assert np.array_equal(rv, c["rvalues"]) and np.array_equal(ri, c["rindexes"]) and np.array_equal(rm, c["rmask"])
        n += 1
    assert n >= 6


def test_package_pattern_matches_goldens(golden_dir):
    for c in _cases(golden_dir):
        B, K, T = c["z"].shape
        pat = DelayedPatternProvider(n_q=K).get_pattern(T)
        v, i, m = pat.build_pattern_sequence(torch.from_numpy(c["z"]), 2048)
        assert np.array_equal(v.numpy(), c["values"])
        assert np.array_equal(i.numpy(), c["indexes"])
        assert np.array_equal(m.numpy(), c["mask"])
        rv, ri, rm = pat.revert_pattern_sequence(v, 2048)
        assert np.array_equal(rv.numpy(), c["rvalues"]) and np.array_equal(ri.numpy(), c["rindexes"])
        assert np.array_equal(rm.numpy(), c["rmask"])
        assert pat.max_delay == K - 1 and pat.num_sequence_steps == T + K - 1


def test_roundtrip_property_large():
    """build -> revert is the identity on the valid region at BASELINE sizes (16 s, K=8)."""
    rng = np.random.RandomState(1)
    z = torch.from_numpy(rng.randint(0, 2048, size=(3, 8, 800)).astype(np.int64))
    pat = DelayedPatternProvider(8).get_pattern(800)
    v, _, _ = pat.build_pattern_sequence(z, 2048)
    assert v.shape == (3, 8, 808)
    rv, _, rm = pat.revert_pattern_sequence(v, 2048)
    assert rm.all() and torch.equal(rv, z)


def test_nondefault_delays_and_flatten():
    rng = np.random.RandomState(2)
    z = rng.randint(0, 100, size=(1, 3, 7)).astype(np.int64)
    for kw in (dict(delays=[0, 0, 2]), dict(flatten_first=2), dict(empty_initial=2), dict(delays=[0, 1, 1], flatten_first=1)):
        v, i, m = DelayedPatternProvider(3, **kw).get_pattern(7).build_pattern_sequence(torch.from_numpy(z), -1)
        ov, oi, om = patterns_oracle.build_pattern_sequence(z, -1, **kw)
        assert np.array_equal(v.numpy(), ov) and np.array_equal(i.numpy(), oi) and np.array_equal(m.numpy(), om), kw


def test_random_shapes_package_equals_oracle_and_roundtrips():
    """Property check over random (B, K, T) including the degenerate ones (T = 1, K = 1) and non-default options:
    the package's vectorised index tables equal the oracle's step-by-step restatement, and revert(build(z)) == z."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 8), st.integers(1, 33), st.integers(0, 3), st.integers(0, 2), st.integers(0, 10 ** 6))
    def check(B, K, T, flatten_first, empty_initial, seed):
        rng = np.random.RandomState(seed)
        z = rng.randint(0, 2048, size=(B, K, T)).astype(np.int64)
        kw = dict(flatten_first=flatten_first, empty_initial=empty_initial)
        pat = DelayedPatternProvider(K, **kw).get_pattern(T)
        v, i, m = pat.build_pattern_sequence(torch.from_numpy(z), 2048)
        ov, oi, om = patterns_oracle.build_pattern_sequence(z, 2048, **kw)
        assert np.array_equal(v.numpy(), ov) and np.array_equal(i.numpy(), oi) and np.array_equal(m.numpy(), om)
        rv, _, rm = pat.revert_pattern_sequence(v, 2048)
        assert rv.shape == (B, K, T)
        assert bool(rm.all()) and np.array_equal(rv.numpy(), z)

    check()


def test_empty_sequence_raises_like_the_reference():
    """T = 0: the reference indexes an empty tensor and raises IndexError (codebooks_patterns.py:163-170); so does the package."""
    z = torch.zeros(1, 4, 0, dtype=torch.int64)
    with pytest.raises(IndexError):
        DelayedPatternProvider(4).get_pattern(0).build_pattern_sequence(z, 2048)

@pytest.mark.gpu
def test_device_delay_kernel_matches_oracle():
    rng = np.random.RandomState(3)
    for (B, K, T) in [(1, 4, 150), (2, 8, 800), (3, 4, 1), (1, 1, 5)]:
        z = rng.randint(0, 2048, size=(B, K, T)).astype(np.int64)
        v, _, _ = DelayedPatternProvider(K).get_pattern(T).build_pattern_sequence(torch.from_numpy(z).cuda(), 2048)
        assert v.is_cuda and np.array_equal(v.cpu().numpy(), patterns_oracle.delay_closed_form(z, 2048))
