"""BASELINE.json configs[0] on the GPU path next to the reference's own CPU timing (tests/golden/lm_cfg1_330m.json):
323M stand-in (d=1024, L=24, H=16 -> head_dim 64), one utterance, 3 s prompt -> 5 s generated, top-k 40.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from voicecraft_b200 import synthetic
from voicecraft_b200.voicecraft import VoiceCraft
g = np.load(os.path.join(gu.GOLDEN, "lm_cfg1_330m.npz"))
meta = json.load(open(os.path.join(gu.GOLDEN, "lm_cfg1_330m.json")))
cfg = synthetic.make_config("330M")
sd = gu.suppress_end_tokens(cfg, synthetic.make_state_dict(cfg, seed=0))
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(kv_dtype="bf16", max_slots=1, max_seq_len=512)
x, xl, y = (torch.from_numpy(g[k]).cuda() for k in ("x", "x_lens", "y"))
kw = dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)
times = []
for rep in range(4):
    torch.manual_seed(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, gen = m.inference_tts(x, xl, y, **kw)
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
G, K = int(gen.shape[-1]), cfg.n_codebooks
dt = sorted(times[1:])[1]
print(json.dumps({"workload": "BASELINE configs[0]: 323M stand-in (head_dim 64), B=1, 3 s prompt -> 5 s generated, top-k 40, bf16 KV",
                  "generated_frames": G, "seconds_median_of_3": dt, "seconds_first_call": times[0], "codec_tokens_per_s": G * K / dt,
                  "rtf_x": G / 50.0 / dt, "reference_cpu_8_threads_build_container": meta["reference_cpu"],
                  "oracle_port_cpu_8_threads_build_container": meta["oracle_port_cpu"]}))
