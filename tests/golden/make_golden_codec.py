"""Pin oracle/encodec_oracle.py against transformers' EncodecModel (structural twin of audiocraft's EnCodec, the
reference's un-vendored dependency) and write small input/output fixtures.  Build container only.
    python tests/golden/make_golden_codec.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import encodec_oracle as eo  # noqa: E402

CASES = {
    # name: (config overrides, B, T, seed)
    "small_causal_reflect": (dict(n_filters=8, dimension=32, bins=64, lstm=2), 2, 24, 1),
    "small_noncausal_trueskip": (dict(n_filters=8, dimension=32, bins=64, lstm=1, causal=False, true_skip=True), 1, 17, 2),
    "small_constpad": (dict(n_filters=8, dimension=32, bins=64, lstm=1, pad_mode="constant"), 3, 9, 3),
    "mid_default": (dict(n_filters=16, dimension=64, bins=256), 1, 50, 4),
}


def main():
    torch.set_num_threads(8)
    out = {}
    for name, (over, B, T, seed) in CASES.items():
        cfg = eo.default_config(**over)
        sd = eo.make_state_dict(cfg, seed=seed)
        g = torch.Generator().manual_seed(100 + seed)
        codes = torch.randint(0, cfg.bins, (B, cfg.n_q, T), generator=g)
        wav = eo.decode(cfg, sd, codes)
        hf = eo.to_hf_model(cfg, sd)
        with torch.no_grad():
            ref = hf.decode(codes.unsqueeze(0), [None])[0]
        assert ref.shape == wav.shape == (B, 1, T * int(np.prod(cfg.ratios))), (ref.shape, wav.shape)
        err = float((ref - wav).abs().max())
        print(f"{name}: shape {tuple(wav.shape)} max|oracle - transformers twin| = {err:.3g}, rms {float(ref.pow(2).mean().sqrt()):.3g}")
        assert err < 2e-5
        out[f"{name}.codes"] = codes.numpy()
        out[f"{name}.wav"] = ref.numpy()
    np.savez_compressed(os.path.join(HERE, "codec.npz"), **out)
    print("codec goldens written")

    # ---- encode direction: wav -> codes through the twin's encoder + RVQ ------------------
    enc = {}
    for name, (over, B, T, seed) in CASES.items():
        cfg = eo.default_config(**over)
        sd = eo.make_state_dict(cfg, seed=seed, encoder=True)
        hop = int(np.prod(cfg.ratios))
        g = torch.Generator().manual_seed(200 + seed)
        N = T * hop - (37 if name == "mid_default" else 0)          # one ragged length: exercises the extra right padding
        wav = torch.randn(B, cfg.channels, N, generator=g) * 0.3
        z = eo.encode_latent(cfg, sd, wav)
        codes, gaps = eo.rvq_encode(cfg, sd, z, return_gaps=True)
        hf = eo.to_hf_model(cfg, sd)
        with torch.no_grad():
            zr = hf.encoder(wav)
            ref = hf.encode(wav, bandwidth=hf.config.target_bandwidths[0]).audio_codes[0]
        err = float((zr - z).abs().max())
        same = float((ref == codes).float().mean())
        # a code may differ from the twin's only where the two nearest codes are (numerically) equidistant
        bad = (ref != codes)
        assert err < 2e-5 and ref.shape == codes.shape
        assert (not bad.any()) or float(gaps[bad].max()) < 1e-4, (same, float(gaps[bad].max()))
        print(f"{name}: wav {tuple(wav.shape)} -> codes {tuple(codes.shape)}, max|latent - twin| = {err:.3g}, codes equal {same:.4f}, "
              f"min decision gap {float(gaps.min()):.3g}")
        enc[f"{name}.wav"] = wav.numpy()
        enc[f"{name}.latent"] = zr.numpy()
        enc[f"{name}.codes"] = ref.numpy()
        enc[f"{name}.gaps"] = gaps.numpy()
    np.savez_compressed(os.path.join(HERE, "codec_encode.npz"), **enc)
    print("codec encode goldens written")


if __name__ == "__main__":
    main()
