"""Host-side integer logic of the drop-in VoiceCraft class that needs no GPU: the speech-editing prompt layout (segment
order, placeholder columns, per-segment delay pattern, cut; reference voicecraft.py:239-320, 615-683) and the un-delay
of sampled rows (:1126-1137), against the oracle that is pinned to the reference by the edit fixtures."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import lm_oracle
from voicecraft_b200 import synthetic
from voicecraft_b200.voicecraft import VoiceCraft

_MODELS = {}


def _pair(variant):
    if variant not in _MODELS:
        over = dict(eos=-1, n_special=3, reduced_eog=0) if variant.endswith("noeos") else {}
        if variant.startswith("k8"):
            over["n_codebooks"] = 8
        cfg = synthetic.make_config("tiny", **over)
        sd = synthetic.make_state_dict(cfg, seed=1)
        m = VoiceCraft(cfg)
        m.load_state_dict(sd)
        _MODELS[variant] = (cfg, m.eval(), lm_oracle.OracleLM(cfg, sd))
    return _MODELS[variant]


@st.composite
def _spans(draw):
    T = draw(st.integers(24, 90))
    M = draw(st.integers(1, 3))
    cuts = sorted(draw(st.lists(st.integers(1, T - 1), min_size=2 * M, max_size=2 * M, unique=True)))
    spans = [(cuts[2 * i], cuts[2 * i + 1]) for i in range(M)]
    return T, spans


@settings(max_examples=40, deadline=None)
@given(_spans(), st.sampled_from(["k4", "k4_noeos", "k8_noeos"]), st.integers(0, 10 ** 6))
def test_edit_prompt_layout_matches_oracle(ts, variant, seed):
    T, spans = ts
    cfg, model, oracle = _pair(variant)
    K = cfg.n_codebooks
    y = torch.from_numpy(np.random.RandomState(seed).randint(0, 2048, size=(K, T)).astype(np.int64))
    tok, mask_rows, more, non_mask = model._edit_prompt(y, spans)
    o_tok, o_pos, o_val, o_more, o_non_mask = oracle.edit_prompt(y, spans)
    assert tok.shape == (o_tok.shape[1], K)
    assert torch.equal(tok.t(), o_tok)
    rows = mask_rows.tolist()
    assert [i for i, v in enumerate(rows) if v >= 0] == list(o_pos)
    assert [v for v in rows if v >= 0] == list(o_val)
    assert list(more) == list(o_more) and list(non_mask) == list(o_non_mask)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 8), st.integers(0, 40), st.integers(0, 10 ** 6))
def test_undelay_matches_oracle(K, G, seed):
    """rows [G+K, K] sampled in delayed order -> [K, G] codes (generation always ends with the K-step EOG cascade)."""
    cfg, _, oracle = _pair("k4")
    rows = np.random.RandomState(seed).randint(0, 2048, size=(G + K, K)).astype(np.int64)
    got = VoiceCraft._undelay(rows, K)
    saved = oracle.c.n_codebooks
    try:
        oracle.c.n_codebooks = K
        ref = oracle._undelay([torch.from_numpy(r) for r in rows]).numpy()
    finally:
        oracle.c.n_codebooks = saved
    assert got.shape == (K, G) and np.array_equal(got, ref)
