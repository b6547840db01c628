"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the delayed codebook pattern.

Follows reference models/codebooks_patterns.py:
  * DelayedPatternProvider.get_pattern            :336-352  (layout construction)
  * Pattern._build_pattern_sequence_scatter_indexes :117-149 (index / mask tables)
  * Pattern.build_pattern_sequence                 :151-176  (gather with special token)
  * Pattern._build_reverted_sequence_scatter_indexes :178-218, revert_pattern_sequence :220-245

Pinned against the imported reference by tests/golden/make_golden.py
(fixtures tests/golden/patterns_*.npz) and the docstring example :307-316.
"""
import numpy as np


def delayed_layout(timesteps, n_q, delays=None, flatten_first=0, empty_initial=0):
    """Return the layout as a list (one entry per sequence step) of lists of (t, q).

    reference: codebooks_patterns.py:336-352.
    """
    if delays is None:
        delays = list(range(n_q))
    layout = [[]]
    layout += [[] for _ in range(empty_initial)]
    if flatten_first:
        for t in range(min(timesteps, flatten_first)):
            for q in range(n_q):
                layout.append([(t, q)])
    for t in range(flatten_first, timesteps + max(delays)):
        step = []
        for q, d in enumerate(delays):
            tq = t - d
            if tq >= flatten_first:
                step.append((tq, q))
        layout.append(step)
    return layout


def max_delay(layout, timesteps):
    """reference: codebooks_patterns.py:84-90."""
    m = 0
    for step in layout[1:]:
        for (t, _q) in step:
            m = max(m, t + 1)
    return m - timesteps


def build_indexes(layout, timesteps, n_q, keep_only_valid_steps=False):
    """indexes[K,S] into flattened z (+ special slot K*T) and validity mask.

    reference: codebooks_patterns.py:117-149.
    """
    if keep_only_valid_steps:
        layout = layout[: len(layout) - max_delay(layout, timesteps)]
    S = len(layout)
    indexes = np.full((n_q, S), n_q * timesteps, dtype=np.int64)
    mask = np.zeros((n_q, S), dtype=bool)
    for s, step in enumerate(layout):
        for (t, q) in step:
            if t < timesteps:
                indexes[q, s] = t + q * timesteps
                mask[q, s] = True
    return indexes, mask


def build_pattern_sequence(z, special_token, n_q=None, delays=None, flatten_first=0,
                           empty_initial=0, keep_only_valid_steps=False):
    """z [B,K,T] int -> (values [B,K,S], indexes [K,S], mask [K,S]).

    reference: codebooks_patterns.py:151-176.
    """
    z = np.asarray(z)
    B, K, T = z.shape
    layout = delayed_layout(T, K, delays, flatten_first, empty_initial)
    indexes, mask = build_indexes(layout, T, K, keep_only_valid_steps)
    flat = np.concatenate([z.reshape(B, -1), np.full((B, 1), special_token, dtype=z.dtype)], axis=1)
    values = flat[:, indexes.reshape(-1)].reshape(B, K, indexes.shape[-1])
    return values, indexes, mask


def build_reverted_indexes(layout, timesteps, n_q, sequence_steps, keep_only_valid_steps=False,
                           is_model_output=False):
    """reference: codebooks_patterns.py:178-218."""
    if keep_only_valid_steps:
        layout = layout[: len(layout) - max_delay(layout, timesteps)]
    assert sequence_steps <= len(layout)
    if is_model_output:
        layout = layout[1:]
    indexes = np.full((n_q, timesteps), n_q * sequence_steps, dtype=np.int64)
    mask = np.zeros((n_q, timesteps), dtype=bool)
    for s, step in enumerate(layout):
        if s < sequence_steps:
            for (t, q) in step:
                if t < timesteps:
                    indexes[q, t] = s + q * sequence_steps
                    mask[q, t] = True
    return indexes, mask


def revert_pattern_sequence(s, special_token, timesteps, delays=None, flatten_first=0,
                            empty_initial=0, keep_only_valid_steps=False):
    """s [B,K,S] -> (values [B,K,T], indexes [K,T], mask [K,T]).

    reference: codebooks_patterns.py:220-245.
    """
    s = np.asarray(s)
    B, K, S = s.shape
    layout = delayed_layout(timesteps, K, delays, flatten_first, empty_initial)
    indexes, mask = build_reverted_indexes(layout, timesteps, K, S, keep_only_valid_steps)
    flat = np.concatenate([s.reshape(B, -1), np.full((B, 1), special_token, dtype=s.dtype)], axis=1)
    values = flat[:, indexes.reshape(-1)].reshape(B, K, indexes.shape[-1])
    return values, indexes, mask


def delay_closed_form(z, special_token):
    """Closed form of the default delay pattern used by the CUDA path:
    values[b,k,s] = z[b,k,s-1-k] if 0 <= s-1-k < T else special, S = T + K.
    (Derived from :336-352 with delays=range(K); checked against build_pattern_sequence.)
    """
    z = np.asarray(z)
    B, K, T = z.shape
    out = np.full((B, K, T + K), special_token, dtype=z.dtype)
    for k in range(K):
        out[:, k, 1 + k: 1 + k + T] = z[:, k, :]
    return out
