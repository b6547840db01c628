"""Generate golden fixtures by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

For every case it (1) runs the unmodified reference model on CPU fp32 with synthetic weights
(voicecraft_b200.synthetic.make_state_dict) and a fixed torch seed, (2) runs oracle/lm_oracle.py on the
same inputs/seed and REQUIRES identical token ids (and identical logits where traced), and (3) writes the
inputs + reference outputs to tests/golden/*.npz.  The fixtures are what tests/ compare the oracle
(-m "not gpu") and the CUDA path (-m gpu) against.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    """The only missing import on the LM path is torchmetrics (voicecraft.py:10): stub it."""
    sys.path.insert(0, REF)
    tm = types.ModuleType("torchmetrics")
    tmc = types.ModuleType("torchmetrics.classification")

    class MulticlassAccuracy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    tmc.MulticlassAccuracy = MulticlassAccuracy
    tm.classification = tmc
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.classification"] = tmc
    from models import voicecraft, codebooks_patterns
    return voicecraft, codebooks_patterns


def ref_model(voicecraft, cfg, sd):
    from argparse import Namespace
    m = voicecraft.VoiceCraft(Namespace(**vars(cfg)))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("accuracy_metrics") for k in missing), missing
    return m.eval()


CASES = [
    # name, cfg name, cfg overrides, kind, params
    dict(name="tts_topk40", cfg="tiny", kind="tts", text_len=4, prompt=20, seed=11,
         kw=dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="tts_greedy", cfg="tiny", kind="tts", text_len=5, prompt=17, seed=12,
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="tts_topp", cfg="tiny", kind="tts", text_len=5, prompt=12, seed=13,
         kw=dict(top_k=0, top_p=0.8, temperature=0.9, stop_repetition=2), silence_bias=True),
    dict(name="tts_eos_bias", cfg="tiny", kind="tts", text_len=8, prompt=10, seed=14, eos_bias=7.0,
         kw=dict(top_k=50, top_p=0.95, temperature=1.0, stop_repetition=3)),
    dict(name="tts_k8_noeos", cfg="tiny", over=dict(n_codebooks=8, eos=-1, n_special=3, reduced_eog=0), kind="tts",
         text_len=4, prompt=15, seed=15, kw=dict(top_k=30, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="tts_small", cfg="small", kind="tts", text_len=6, prompt=30, seed=16,
         kw=dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="batch3", cfg="tiny", kind="batch", text_len=4, prompt=14, seed=17, batch_size=3,
         kw=dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="batch4_eos", cfg="tiny", kind="batch", text_len=6, prompt=9, seed=18, batch_size=4, eos_bias=6.0,
         kw=dict(top_k=0, top_p=0.9, temperature=1.0, stop_repetition=3)),
    dict(name="edit1", cfg="tiny", kind="edit", text_len=8, prompt=40, seed=19, spans=[(10, 18)],
         kw=dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=-1)),
    dict(name="edit2", cfg="tiny", kind="edit", text_len=9, prompt=50, seed=20, spans=[(8, 14), (30, 37)],
         kw=dict(top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2), eog_bias=3.0),
    dict(name="edit3_noeos", cfg="tiny", over=dict(eos=-1, n_special=3, reduced_eog=0), kind="edit", text_len=9,
         prompt=45, seed=21, spans=[(5, 9), (20, 24), (33, 40)],
         kw=dict(top_k=20, top_p=1.0, temperature=1.0, stop_repetition=-1), eog_bias=3.0),
]

# second batch (sampler corner cases, K = 8 in the batch / edit loops); generated with --only so the first batch's files
# stay byte-identical
CASES += [
    dict(name="tts_topk_all", cfg="tiny", kind="tts", text_len=4, prompt=11, seed=31,
         kw=dict(top_k=5000, top_p=1.0, temperature=1.0, stop_repetition=3)),          # top_k > vocabulary
    dict(name="tts_temp03", cfg="tiny", kind="tts", text_len=5, prompt=13, seed=32,
         kw=dict(top_k=0, top_p=1.0, temperature=0.3, stop_repetition=3)),             # temperature only
    dict(name="tts_topp_tiny", cfg="tiny", kind="tts", text_len=4, prompt=10, seed=33,
         kw=dict(top_k=0, top_p=0.05, temperature=1.0, stop_repetition=3)),            # keeps min_tokens_to_keep = 1
    dict(name="tts_norep_sil", cfg="tiny", kind="tts", text_len=5, prompt=12, seed=34, silence_bias=True,
         kw=dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=-1)),           # penalty disabled, silence repeats
    dict(name="batch2_k8", cfg="tiny", over=dict(n_codebooks=8, eos=-1, n_special=3, reduced_eog=0), kind="batch",
         text_len=4, prompt=12, seed=35, batch_size=2, kw=dict(top_k=25, top_p=1.0, temperature=1.0, stop_repetition=3)),
    dict(name="edit1_k8", cfg="tiny", over=dict(n_codebooks=8, eos=-1, n_special=3, reduced_eog=0), kind="edit",
         text_len=7, prompt=36, seed=36, spans=[(6, 13)], eog_bias=3.0,
         kw=dict(top_k=30, top_p=1.0, temperature=1.0, stop_repetition=-1)),
]

SILENCE = [1388, 1898, 131]


def build_case(case):
    from voicecraft_b200 import synthetic
    cfg = synthetic.make_config(case["cfg"], **case.get("over", {}))
    sd = synthetic.make_state_dict(cfg, seed=case["seed"])
    if case.get("eos_bias"):      # make the end token likely so natural termination is exercised
        end = cfg.eos if cfg.eos > 0 else cfg.eog
        sd["predict_layer.0.2.bias"][end] += case["eos_bias"]
    if case.get("eog_bias"):
        sd["predict_layer.0.2.bias"][cfg.eog] += case["eog_bias"]
    if case.get("silence_bias"):  # make a silence token repeat so the repetition penalty fires
        sd["predict_layer.0.2.bias"][SILENCE[0]] += 9.0
    x, x_lens, y = synthetic.synthetic_utterance(cfg, 1000 + case["seed"], case["text_len"], case["prompt"])
    return cfg, sd, x, x_lens, y


def main():
    torch.set_num_threads(8)
    voicecraft, cbp = import_reference()
    from oracle import lm_oracle, patterns_oracle
    only = None
    if "--only" in sys.argv:          # regenerate just these LM cases and merge them into lm_cases.json
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))

    # ---- pattern goldens (codebooks_patterns.py docstring :307-316 plus random cases) -------------
    pat = {}
    rng = np.random.RandomState(0)
    for i, (K, T) in enumerate([] if only else [(3, 4), (4, 6), (4, 1), (8, 13), (4, 150), (1, 5)]):
        z = rng.randint(0, 2048, size=(2, K, T)).astype(np.int64)
        prov = cbp.DelayedPatternProvider(n_q=K)
        vals, idx, msk = prov.get_pattern(T).build_pattern_sequence(torch.from_numpy(z), 2048, False)
        ov, oi, om = patterns_oracle.build_pattern_sequence(z, 2048)
        assert np.array_equal(vals.numpy(), ov) and np.array_equal(idx.numpy(), oi) and np.array_equal(msk.numpy(), om)
        assert np.array_equal(patterns_oracle.delay_closed_form(z, 2048), ov)
        rv, ri, rm = prov.get_pattern(T).revert_pattern_sequence(vals, 2048)
        orv, ori, orm = patterns_oracle.revert_pattern_sequence(ov, 2048, T)
        assert np.array_equal(rv.numpy(), orv) and np.array_equal(ri.numpy(), ori) and np.array_equal(rm.numpy(), orm)
        pat[f"z{i}"], pat[f"values{i}"], pat[f"indexes{i}"], pat[f"mask{i}"] = z, ov, oi, om
        pat[f"rvalues{i}"], pat[f"rindexes{i}"], pat[f"rmask{i}"] = orv, ori, orm
    if not only:
        np.savez_compressed(os.path.join(HERE, "patterns.npz"), **pat)
        print("patterns ok")

    # ---- LM goldens -----------------------------------------------------------------------------
    meta = {}
    if only:
        with open(os.path.join(HERE, "lm_cases.json")) as f:
            meta = json.load(f)
    for case in CASES:
        if only and case["name"] not in only:
            continue
        cfg, sd, x, x_lens, y = build_case(case)
        model = ref_model(voicecraft, cfg, sd)
        oracle = lm_oracle.OracleLM(cfg, sd)
        kw = dict(case["kw"], silence_tokens=SILENCE, kvcache=1)
        out = {}
        # hook the reference's logits: wrap predict_layer[K-1] is awkward; trace via topk_sampling instead
        trace = []
        orig = voicecraft.topk_sampling

        def spy(logits, **k):
            trace.append(logits.detach().clone())
            return orig(logits, **k)
        voicecraft.topk_sampling = spy
        torch.manual_seed(case["seed"])
        if case["kind"] == "tts":
            res, gen = model.inference_tts(x, x_lens, y, **kw)
        elif case["kind"] == "batch":
            res, gen = model.inference_tts_batch(x, x_lens, y, batch_size=case["batch_size"], **kw)
        else:
            mi = torch.tensor(case["spans"], dtype=torch.long).unsqueeze(0)
            res = model.inference(x, x_lens, y, mi, **kw)
            gen = None
        voicecraft.topk_sampling = orig

        # kvcache=0 invariant of the reference
        if case["kind"] in ("tts", "edit") and case["name"] in ("tts_topk40", "edit2"):
            torch.manual_seed(case["seed"])
            kw0 = dict(kw, kvcache=0)
            if case["kind"] == "tts":
                res0, _ = model.inference_tts(x, x_lens, y, **kw0)
            else:
                res0 = model.inference(x, x_lens, y, mi, **kw0)
            out["kvcache0_equal"] = np.array(int(torch.equal(res0, res)))

        # oracle must reproduce the reference exactly (same seed -> same CPU noise stream)
        otrace = []
        orig_sr = lm_oracle.sample_rows

        def spy2(logits, *a, **k):
            otrace.append(logits.detach().clone())
            return orig_sr(logits, *a, **k)
        lm_oracle.sample_rows = spy2
        torch.manual_seed(case["seed"])
        if case["kind"] == "tts":
            ores, ogen = oracle.inference_tts(x, x_lens, y, **kw)
        elif case["kind"] == "batch":
            ores, ogen = oracle.inference_tts_batch(x, x_lens, y, batch_size=case["batch_size"], **kw)
        else:
            ores = oracle.inference(x, x_lens, y, mi, **kw)
            ogen = None
        lm_oracle.sample_rows = orig_sr
        assert torch.equal(ores, res), f"{case['name']}: oracle tokens differ from reference"
        if gen is not None:
            assert torch.equal(ogen, gen)
        assert len(trace) == len(otrace)
        maxdiff = max(float((a - b).abs().max()) for a, b in zip(trace, otrace))
        print(f"{case['name']}: steps={len(trace)} res={tuple(res.shape)} oracle==reference, max|dlogit|={maxdiff:.3g}")
        # B=1 flows are bit-identical; with B>1 ATen's CPU linear takes an input-layout dependent path
        # (addmm vs matmul+add_) and MKL results move by ~2e-6 -- same tokens required regardless.
        assert maxdiff <= 2e-5, "oracle logits must match the reference on CPU"
        out["oracle_max_dlogit"] = np.array(maxdiff)

        out.update(x=x.numpy(), x_lens=x_lens.numpy(), y=y.numpy(), res=res.numpy())
        if gen is not None:
            out["gen"] = gen.numpy()
        if case["kind"] == "edit":
            out["mask_interval"] = mi.numpy()
        # keep a thin logit trace (steps 0, 1, mid, last) for tolerance checks of the CUDA path
        keep = sorted(set([0, 1, len(trace) // 2, len(trace) - 1]))
        out["trace_steps"] = np.array(keep)
        out["trace_logits"] = np.stack([trace[i].reshape(-1, trace[i].shape[-1]).numpy() for i in keep])
        out["n_steps"] = np.array(len(trace))
        np.savez_compressed(os.path.join(HERE, f"lm_{case['name']}.npz"), **out)
        meta[case["name"]] = {k: v for k, v in case.items() if k not in ("name",)}
    with open(os.path.join(HERE, "lm_cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("all LM goldens written")


if __name__ == "__main__":
    main()
