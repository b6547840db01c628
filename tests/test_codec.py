"""EnCodec decode: oracle vs the fixtures produced from the transformers twin (CPU), CUDA path vs oracle (-m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import encodec_oracle as eo

CASES = {
    "small_causal_reflect": (dict(n_filters=8, dimension=32, bins=64, lstm=2), 1),
    "small_noncausal_trueskip": (dict(n_filters=8, dimension=32, bins=64, lstm=1, causal=False, true_skip=True), 2),
    "small_constpad": (dict(n_filters=8, dimension=32, bins=64, lstm=1, pad_mode="constant"), 3),
    "mid_default": (dict(n_filters=16, dimension=64, bins=256), 4),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_twin_fixture(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "codec.npz"))
    over, seed = CASES[name]
    cfg = eo.default_config(**over)
    sd = eo.make_state_dict(cfg, seed=seed)
    wav = eo.decode(cfg, sd, torch.from_numpy(g[f"{name}.codes"]))
    ref = g[f"{name}.wav"]
    assert wav.shape == ref.shape
    assert np.abs(wav.numpy() - ref).max() < 2e-5


def test_hop_and_shapes():
    cfg = eo.default_config(n_filters=4, dimension=16, bins=32, lstm=0)
    sd = eo.make_state_dict(cfg, seed=0)
    for T in (1, 7, 8):
        codes = torch.randint(0, 32, (2, 4, T))
        assert eo.decode(cfg, sd, codes).shape == (2, 1, 320 * T)


def _gpu_tok(cfg, sd):
    from voicecraft_b200.tokenizer import AudioTokenizer
    return AudioTokenizer(device="cuda:0", config=cfg, state_dict=sd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_decode_matches_fixture(name, golden_dir):
    """Tolerance: fp32 kernels vs fp32 twin, |err| <= 2e-4 on O(1) waveforms (different summation order only)."""
    g = np.load(os.path.join(golden_dir, "codec.npz"))
    over, seed = CASES[name]
    cfg = eo.default_config(**over)
    tok = _gpu_tok(cfg, eo.make_state_dict(cfg, seed=seed))
    codes = torch.from_numpy(g[f"{name}.codes"])
    wav = tok.decode_codes(codes.cuda()).cpu().numpy()
    ref = g[f"{name}.wav"]
    assert wav.shape == ref.shape
    assert np.abs(wav - ref).max() < 2e-4, np.abs(wav - ref).max()
    # reference call signature: one utterance, [(codes, None)]
    one = tok.decode([(codes[:1].cuda(), None)])
    assert one.shape == (1, 1, ref.shape[-1]) and np.abs(one.cpu().numpy() - ref[:1]).max() < 2e-4


@pytest.mark.gpu
def test_cuda_decode_full_size_vs_oracle():
    """Real codec shape (n_filters 64, 4x2048, LSTM 2) on 2 x 1 s of tokens; odd T exercises ragged tiles."""
    cfg = eo.default_config()
    sd = eo.make_state_dict(cfg, seed=7)
    codes = torch.randint(0, 2048, (2, 4, 53), generator=torch.Generator().manual_seed(5))
    ref = eo.decode(cfg, sd, codes).numpy()
    wav = _gpu_tok(cfg, sd).decode_codes(codes.cuda()).cpu().numpy()
    assert wav.shape == ref.shape == (2, 1, 53 * 320)
    err = np.abs(wav - ref).max()
    assert err < 5e-4 * max(1.0, np.abs(ref).max()), err


@pytest.mark.gpu
def test_cuda_decode_batch_independence():
    """Property at larger sizes: decoding a batch equals decoding each row alone (bit-exact, same kernels)."""
    cfg = eo.default_config(n_filters=16, dimension=64, bins=256)
    tok = _gpu_tok(cfg, eo.make_state_dict(cfg, seed=9))
    codes = torch.randint(0, 256, (19, 4, 40), generator=torch.Generator().manual_seed(6)).cuda()
    full = tok.decode_codes(codes)
    for b in (0, 7, 18):
        assert torch.equal(full[b:b + 1], tok.decode_codes(codes[b:b + 1]))
