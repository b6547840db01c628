"""Experiment: the B=32 decode step as TWO independent PDL chains (two engines, 16 utterances each, two CUDA streams) instead of
one chain of 32 rows.  The GEMMs of a step are latency-bound (0.33 of the HBM roofline), so two chains may overlap: one chain's
attention streams KV while the other's GEMM waits on its pipeline.  Prints ms per 32-utterance step for both arrangements.
usage: exp_dual_chain.py [steps] [chains]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voicecraft_b200.voicecraft import VoiceCraft

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sys.argv = [sys.argv[0]]
args = bench.parse()
cfg, sd = bench.make_model(args)
B = 32
dev = torch.device("cuda:0")
utts = bench.make_utterances(args, cfg, range(B))
cap = args.text_len * (cfg.encodec_sr // 5)
S_total = cap - (args.prompt + 1) - 2
start = max(3, (S_total - steps) // 2)
kw = dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)


def run(nch):
    per = B // nch
    models, sessions, streams = [], [], []
    for c in range(nch):
        m = VoiceCraft(cfg)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        m.configure_engine(max_slots=per, max_seq_len=(args.text_len + cap + 64 + 255) // 256 * 256, kv_dtype="bf16", max_new_tokens=cap + 64)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            u = utts[c * per:(c + 1) * per]
            sess = m.open_tts_session([x[0].to(dev) for x in u], [x[2].to(dev) for x in u], seeds=[1 + c * per + i for i in range(per)], **kw)
            sess.sample()
        models.append(m); sessions.append(sess); streams.append(s)
    for _ in range(start):
        for sess in sessions:
            sess.step()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(nch)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(nch)]
    t0 = time.perf_counter()
    for c in range(nch):
        e0[c].record(streams[c])
    for _ in range(steps):
        for sess in sessions:
            sess.step()
    for c in range(nch):
        e1[c].record(streams[c])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    dev_ms = max(e0[c].elapsed_time(e1[c]) for c in range(nch))
    span = max(e0[0].elapsed_time(e1[c]) for c in range(nch))
    for sess in sessions:
        st = sess.poll()
        assert all(x.n_steps == 1 + start + steps for x in st)
        sess.close()
    return {"chains": nch, "rows_per_chain": per, "ms_per_32utt_step": span / steps, "max_chain_ms_per_step": dev_ms / steps,
            "wall_ms_per_step": wall / steps, "codec_tok_s": B * cfg.n_codebooks * steps / (span * 1e-3), "ctx_start": args.text_len + args.prompt + 1 + start}


out = [run(1), run(chains)]
if chains != 4:
    out.append(run(4))
print(json.dumps(out))
