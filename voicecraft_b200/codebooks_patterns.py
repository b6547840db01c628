"""Delayed codebook pattern -- drop-in for the reference's ``models/codebooks_patterns.py`` API surface
that the hot path uses: ``DelayedPatternProvider(n_q, delays, flatten_first, empty_initial).get_pattern(T)``
and ``Pattern.build_pattern_sequence / revert_pattern_sequence / revert_pattern_logits``
(reference codebooks_patterns.py:21-266, 302-352).

Design: the reference materialises a Python list-of-lists layout and fills numpy index tables with a
double loop per new T (:142-146).  Here the layout is kept as two dense integer tables
(``t_of[q, s]`` = source timestep or -1) built with vectorised numpy arithmetic, and the gather itself
runs either as a torch index op or -- for CUDA tensors with the default delays -- in the
``vcb_delay_pattern`` kernel of libvcb200.so (integer gather, coalesced along the sequence).
"""
from functools import lru_cache
import typing as tp

import numpy as np
import torch


class Pattern:
    """Index tables of one interleaving pattern for ``timesteps`` frames and ``n_q`` codebooks.

    ``t_of[q, s]`` is the original timestep that codebook ``q`` shows at sequence step ``s`` (-1: none).
    Step 0 is always empty (it carries the special token), like the reference's ``layout[0] == []``.
    """

    def __init__(self, t_of: np.ndarray, timesteps: int, n_q: int, default_delays: bool = False):
        assert t_of.shape[0] == n_q and t_of.shape[1] >= 1 and (t_of[:, 0] < 0).all()
        self.t_of = t_of
        self.timesteps = timesteps
        self.n_q = n_q
        self._default_delays = default_delays
        self._check()

    # -- the reference's list-of-coordinates view, for API compatibility (codebooks_patterns.py:45-47)
    @property
    def layout(self):
        out = []
        for s in range(self.t_of.shape[1]):
            out.append([(int(self.t_of[q, s]), q) for q in range(self.n_q) if self.t_of[q, s] >= 0])
        return out

    def _check(self):
        """Per codebook the timesteps must be non-decreasing along the sequence (:58-77)."""
        for q in range(self.n_q):
            ts = self.t_of[q][self.t_of[q] >= 0]
            assert (np.diff(ts) >= 0).all(), f"past timesteps found for codebook {q}"

    @property
    def num_sequence_steps(self):
        return self.t_of.shape[1] - 1

    @property
    def max_delay(self):
        return int(self.t_of[:, 1:].max(initial=-1) + 1) - self.timesteps if self.t_of.shape[1] > 1 else -self.timesteps

    def _tables(self, timesteps, keep_only_valid_steps):
        t_of = self.t_of[:, : self.t_of.shape[1] - self.max_delay] if keep_only_valid_steps else self.t_of
        valid = (t_of >= 0) & (t_of < timesteps)
        q_idx = np.arange(self.n_q)[:, None]
        indexes = np.where(valid, t_of + q_idx * timesteps, self.n_q * timesteps).astype(np.int64)
        return indexes, valid

    def build_pattern_sequence(self, z: torch.Tensor, special_token: int, keep_only_valid_steps: bool = False):
        """z [B,K,T] -> (values [B,K,S], indexes [K,S], mask [K,S])   (:151-176)."""
        B, K, T = z.shape
        assert K == self.n_q, f"invalid number of codebooks for the sequence and the pattern: {K} != {self.n_q}"
        assert T <= self.timesteps, "invalid number of timesteps used to build the sequence from the pattern"
        indexes, mask = self._tables(T, keep_only_valid_steps)
        indexes_t = torch.from_numpy(indexes).to(z.device)
        mask_t = torch.from_numpy(mask).to(z.device)
        if (z.is_cuda and self._default_delays and not keep_only_valid_steps and z.dtype == torch.int64
                and T == self.timesteps):
            from . import _lib
            lib = _lib.load()
            z = z.contiguous()
            values = torch.empty(B, K, T + K, dtype=torch.int64, device=z.device)
            with torch.cuda.device(z.device):
                _lib.check(lib.vcb_delay_pattern(z.data_ptr(), values.data_ptr(), B, K, T, int(special_token),
                                                 torch.cuda.current_stream().cuda_stream))
            return values, indexes_t, mask_t
        flat = torch.cat([z.reshape(B, -1), torch.full_like(z[:, :1, 0], special_token)], dim=1)
        values = flat[:, indexes_t.view(-1)].view(B, K, indexes_t.shape[-1])
        return values, indexes_t, mask_t

    def _reverted_tables(self, sequence_steps, keep_only_valid_steps=False, is_model_output=False):
        """(:178-218)"""
        t_of = self.t_of[:, : self.t_of.shape[1] - self.max_delay] if keep_only_valid_steps else self.t_of
        assert sequence_steps <= t_of.shape[1], \
            f"sequence to revert is longer than the defined pattern: {sequence_steps} > {t_of.shape[1]}"
        if is_model_output:
            t_of = t_of[:, 1:]
        T = self.timesteps
        indexes = np.full((self.n_q, T), self.n_q * sequence_steps, dtype=np.int64)
        mask = np.zeros((self.n_q, T), dtype=bool)
        S = min(sequence_steps, t_of.shape[1])
        for q in range(self.n_q):
            s = np.nonzero((t_of[q, :S] >= 0) & (t_of[q, :S] < T))[0]
            indexes[q, t_of[q, s]] = s + q * sequence_steps      # later steps overwrite earlier ones, as in the loop
            mask[q, t_of[q, s]] = True
        return indexes, mask

    def revert_pattern_sequence(self, s: torch.Tensor, special_token: int, keep_only_valid_steps: bool = False):
        """s [B,K,S] -> (values [B,K,T], indexes [K,T], mask [K,T])   (:220-245)."""
        B, K, S = s.shape
        indexes, mask = self._reverted_tables(S, keep_only_valid_steps, is_model_output=False)
        indexes_t = torch.from_numpy(indexes).to(s.device)
        flat = torch.cat([s.reshape(B, -1), torch.full_like(s[:, :1, 0], special_token)], dim=1)
        values = flat[:, indexes_t.view(-1)].view(B, K, indexes_t.shape[-1])
        return values, indexes_t, torch.from_numpy(mask).to(s.device)

    def revert_pattern_logits(self, logits: torch.Tensor, special_token: float, keep_only_valid_steps: bool = False):
        """logits [B,card,K,S] -> (values [B,card,K,T], indexes, mask)   (:247-266)."""
        B, card, K, S = logits.shape
        indexes, mask = self._reverted_tables(S, keep_only_valid_steps, is_model_output=True)
        indexes_t = torch.from_numpy(indexes).to(logits.device)
        flat = logits.reshape(B, card, -1)
        flat = torch.cat([flat, torch.zeros_like(flat[:, :, :1]) + special_token], dim=-1)
        values = flat[:, :, indexes_t.view(-1)].view(B, card, K, indexes_t.shape[-1])
        return values, indexes_t, torch.from_numpy(mask).to(logits.device)


class DelayedPatternProvider:
    """Codebook q is delayed by ``delays[q]`` steps (default q), reference :302-352.

    >>> DelayedPatternProvider(3).get_pattern(4).build_pattern_sequence(z, S)   # z rows = [1,2,3,4]
    [[S, 1, 2, 3, 4, S, S], [S, S, 1, 2, 3, 4, S], [S, S, S, 1, 2, 3, 4]]
    """

    def __init__(self, n_q: int, delays: tp.Optional[tp.List[int]] = None, flatten_first: int = 0,
                 empty_initial: int = 0):
        assert n_q > 0
        self.n_q = n_q
        self._default = delays is None and flatten_first == 0 and empty_initial == 0
        if delays is None:
            delays = list(range(n_q))
        self.delays = delays
        self.flatten_first = flatten_first
        self.empty_initial = empty_initial
        assert len(self.delays) == self.n_q
        assert sorted(self.delays) == self.delays
        self.get_pattern = lru_cache(100)(self.get_pattern)  # type: ignore

    def get_pattern(self, timesteps: int) -> Pattern:
        K, ff = self.n_q, self.flatten_first
        max_delay = max(self.delays)
        n_flat = min(timesteps, ff) * K if ff else 0
        n_body = max(0, timesteps + max_delay - ff)
        S = 1 + self.empty_initial + n_flat + n_body
        t_of = np.full((K, S), -1, dtype=np.int64)
        base = 1 + self.empty_initial
        if ff:
            for t in range(min(timesteps, ff)):          # one codebook per step for the flattened prefix
                for q in range(K):
                    t_of[q, base + t * K + q] = t
        body0 = base + n_flat
        steps = np.arange(ff, ff + n_body)               # the reference's loop variable t
        for q, d in enumerate(self.delays):
            tq = steps - d
            t_of[q, body0: body0 + n_body] = np.where(tq >= ff, tq, -1)
        return Pattern(t_of, n_q=K, timesteps=timesteps, default_delays=self._default)
