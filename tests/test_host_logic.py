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


def test_weight_norm_folding_matches_torch():
    """audiocraft checkpoints store weight-normalised convolutions as (weight_g, weight_v); the engine wants plain weights."""
    from voicecraft_b200.tokenizer import fold_weight_norm
    torch.manual_seed(0)
    for mod in (torch.nn.Conv1d(6, 10, 3), torch.nn.ConvTranspose1d(6, 10, 4, stride=2)):
        wn = torch.nn.utils.weight_norm(mod)              # old-style parametrisation: the checkpoint's key layout
        with torch.no_grad():
            wn.weight_g.mul_(torch.rand_like(wn.weight_g) + 0.5)
        x = torch.randn(2, 6, 9)
        ref = wn(x)
        w = fold_weight_norm(wn.weight_g.detach(), wn.weight_v.detach())
        plain = type(mod)(6, 10, mod.kernel_size[0], stride=mod.stride[0])
        with torch.no_grad():
            plain.weight.copy_(w)
            plain.bias.copy_(wn.bias)
        assert torch.allclose(plain(x), ref, atol=1e-6)


def test_audiocraft_checkpoint_key_mapping_default_config():
    """decoder.model.{i} indices of the SEANet decoder (conv 0, LSTM 1, per ratio [ELU, convtr, resblock], [ELU, conv]):
    for the 4-ratio / 1-residual-layer config the transposed convs sit at 3, 6, 9, 12 and the output conv at 15, as in the
    published EnCodec checkpoints.  Every checkpoint tensor must be consumed, every engine tensor produced."""
    from oracle import encodec_oracle as eo
    from voicecraft_b200.tokenizer import state_dict_from_audiocraft
    cfg = eo.default_config()
    shapes = eo.weight_shapes(cfg)
    g = torch.Generator().manual_seed(0)
    sd = {}

    def wn(prefix, shape):
        sd[prefix + ".weight_g"] = torch.rand(shape[0], *([1] * (len(shape) - 1)), generator=g) + 0.5
        sd[prefix + ".weight_v"] = torch.randn(*shape, generator=g)
        sd[prefix + ".bias"] = torch.randn(shape[0] if "convtr" not in prefix else shape[1], generator=g)
    for q in range(cfg.n_q):
        sd[f"quantizer.vq.layers.{q}._codebook.embed"] = torch.randn(*shapes[f"vq.{q}.embed"], generator=g)
    wn("decoder.model.0.conv.conv", shapes["dec.conv_in.weight"])
    for l in range(cfg.lstm):
        for part in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            sd[f"decoder.model.1.lstm.{part}_l{l}"] = torch.randn(*shapes[f"dec.lstm.{part}_l{l}"], generator=g)
    for i, (ct, rb) in enumerate([(3, 4), (6, 7), (9, 10), (12, 13)]):
        wn(f"decoder.model.{ct}.convtr.convtr", shapes[f"dec.up{i}.convtr.weight"])
        wn(f"decoder.model.{rb}.block.1.conv.conv", shapes[f"dec.up{i}.res0.conv1.weight"])
        wn(f"decoder.model.{rb}.block.3.conv.conv", shapes[f"dec.up{i}.res0.conv2.weight"])
        wn(f"decoder.model.{rb}.shortcut.conv.conv", shapes[f"dec.up{i}.res0.shortcut.weight"])
    wn("decoder.model.15.conv.conv", shapes["dec.conv_out.weight"])

    class Recorder(dict):
        def __init__(self, d):
            super().__init__(d)
            self.used = set()

        def __getitem__(self, k):
            self.used.add(k)
            return super().__getitem__(k)
    rec = Recorder(sd)
    out = state_dict_from_audiocraft(rec, cfg)
    assert set(out) == set(shapes)
    for k, shp in shapes.items():
        assert tuple(out[k].shape) == tuple(shp), k
    assert rec.used == set(sd), sorted(set(sd) - rec.used)
