"""Speech-editing infill throughput (BASELINE.json configs[2] shape on one GPU): 830M, B=16 independent utterances,
T=800 frames each, 160 phonemes, one masked span [300,400); end tokens suppressed so every span runs to the reference's
length cap (y_len > 10 * x_len).  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voicecraft_b200 import synthetic
from voicecraft_b200.voicecraft import VoiceCraft
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
class A: model = "830M"; batch = B; codebooks = 4; text_len = 160; prompt = 800
cfg, sd, utts = bench.make_model_inputs(A)
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(max_slots=B, max_seq_len=2048, max_new_tokens=1200)
xs = [u[0] for u in utts]; ys = [u[2] for u in utts]
spans = [torch.tensor([[[300, 400]]]) for _ in range(B)]
kw = dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=-1)
torch.manual_seed(0)
torch.cuda.synchronize(); t0 = time.perf_counter()
sess = m.open_edit_session([x.cuda() for x in xs], [y.cuda() for y in ys], spans, **kw)
torch.cuda.synchronize(); t_prefill = time.perf_counter() - t0
sess.sample()
for _ in range(5): sess.step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 400
e0.record()
for _ in range(n): sess.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
st = sess.poll()
print(json.dumps({"workload": f"830M speech-editing infill, B={B}, T=800, Lx=160, span [300,400)", "prefill_tokens": B * (160 + 711),
                  "prefill_s": t_prefill, "decode_ms_per_step": ms, "codec_tokens_per_s": B * 4 / ms * 1e3,
                  "ctx_during_timing": [160 + 711 + 6, 160 + 711 + 6 + n], "done": [int(s.done) for s in st][:4]}))
sess.close()
