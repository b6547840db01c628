"""BASELINE.json configs[0] ("330M random-init TTS decode, 3 s prompt -> 5 s gen, batch=1 on CPU"; SURVEY.md section 8d
config 1) run through the UNMODIFIED reference on this container's CPU cores, timed, and pinned as a fixture:

    323M stand-in (d=1024, L=24, H=16 -> head_dim 64, K=4, n_special=4), weights seed 0 (end tokens suppressed: only the
    reference's length cap ends generation), x = 40 phoneme ids (cap 400 frames = 8 s), prompt 150 frames (3 s)
    -> 250 generated frames + the K-step end cascade; top_k=40, top_p=1, temperature=1, stop_repetition=3, kvcache=1,
    torch.manual_seed(1).

Writes tests/golden/lm_cfg1_330m.npz (inputs, the reference's `res`, a thin logit trace) and lm_cfg1_330m.json (timing of
the reference and of the oracle port on the same cores).  The oracle must reproduce the reference's tokens exactly.
Run in the build container only:  python tests/golden/make_golden_cfg1.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
SILENCE = [1388, 1898, 131]
KW = dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)


def checkpoint():
    from voicecraft_b200 import synthetic
    cfg = synthetic.make_config("330M")
    sd = synthetic.make_state_dict(cfg, seed=0)
    for k in range(cfg.n_codebooks):
        sd[f"predict_layer.{k}.2.bias"][cfg.eos] = -1e4
        sd[f"predict_layer.{k}.2.bias"][cfg.eog] = -1e4
    return cfg, sd


def main():
    threads = int(os.environ.get("CFG1_THREADS", "8"))
    torch.set_num_threads(threads)
    from make_golden import import_reference, ref_model
    from oracle import lm_oracle
    from voicecraft_b200 import synthetic
    voicecraft, _ = import_reference()
    cfg, sd = checkpoint()
    x, xl, y = synthetic.synthetic_utterance(cfg, 100, 40, 150)
    model = ref_model(voicecraft, cfg, sd)
    trace = []
    orig = voicecraft.topk_sampling

    def spy(logits, **k):
        trace.append(logits.detach().clone())
        return orig(logits, **k)
    voicecraft.topk_sampling = spy
    torch.manual_seed(1)
    t0 = time.perf_counter()
    res, gen = model.inference_tts(x, xl, y, silence_tokens=SILENCE, kvcache=1, **KW)
    t_ref = time.perf_counter() - t0
    voicecraft.topk_sampling = orig
    oracle = lm_oracle.OracleLM(cfg, sd)
    torch.manual_seed(1)
    t0 = time.perf_counter()
    ores, ogen = oracle.inference_tts(x, xl, y, silence_tokens=SILENCE, **KW)
    t_or = time.perf_counter() - t0
    assert torch.equal(res, ores), "oracle tokens differ from the reference on config 1"
    G = int(gen.shape[-1])
    K = cfg.n_codebooks
    keep = sorted(set([0, 1, len(trace) // 2, len(trace) - 1]))
    np.savez_compressed(os.path.join(HERE, "lm_cfg1_330m.npz"), x=x.numpy(), x_lens=xl.numpy(), y=y.numpy(),
                        res=res.numpy().astype(np.int16), trace_steps=np.array(keep),
                        trace_logits=np.stack([trace[i].reshape(-1, trace[i].shape[-1]).numpy() for i in keep]),
                        n_steps=np.array(len(trace)))
    meta = dict(config="323M stand-in d=1024 L=24 H=16 (head_dim 64) K=4", text_len=40, prompt=150, generated_frames=G,
                sampling_steps=len(trace), seed=1, kw=KW, threads=threads,
                reference_cpu=dict(seconds=t_ref, codec_tokens_per_s=G * K / t_ref, rtf_x=G / 50.0 / t_ref,
                                   note="unmodified reference inference_tts, fp32, incl. prefill"),
                oracle_port_cpu=dict(seconds=t_or, codec_tokens_per_s=G * K / t_or, rtf_x=G / 50.0 / t_or))
    with open(os.path.join(HERE, "lm_cfg1_330m.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
