"""(needs a library built with the device-timeline marks: `make -C voicecraft_b200/csrc clean all TIMELINE=1`)
Per-CTA device timeline of one decode step (830M, B=32): for every GEMM / attention launch the spread of CTA start,
griddepcontrol.wait return and exit times -- shows launch skew, stragglers and the real inter-kernel gaps."""
import ctypes as C, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voicecraft_b200 import _lib
from voicecraft_b200.voicecraft import VoiceCraft
class A: model="830M"; batch=32; codebooks=4; text_len=80; prompt=150
cfg, sd, utts = bench.make_model_inputs(A)
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(max_slots=32, max_seq_len=1024, max_new_tokens=900)
sess = m.open_tts_session([u[0].cuda() for u in utts], [u[2].cuda() for u in utts], top_k=40)
sess.sample()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300): sess.step()
torch.cuda.synchronize()
lib = _lib.load()
lib.vcb_timeline(2, None, 0, None)
for _ in range(2): sess.step()
buf = (C.c_uint64 * (2 * 65536))(); n = C.c_int32()
lib.vcb_timeline(0, buf, 65536, C.byref(n))
print("records", n.value)
per = collections.defaultdict(list)
for i in range(n.value):
    tag, t = buf[2*i], buf[2*i+1]
    if tag & 0x8000: per[tag & 0x7fff].append((t, tag >> 16))
t0 = min(t for v in per.values() for t, _ in v)
launches = collections.defaultdict(list)          # base tag -> list of {cta: t}
for tag, v in per.items():
    v.sort(); cur = {}
    for t, cta in v:
        if cta in cur: launches[tag].append(cur); cur = {}
        cur[cta] = (t - t0) / 1000.0
    if cur: launches[tag].append(cur)
modes = ["qkv", "resid", "act", "logits"]
rows = []
for mode in range(4):
    S, W, E = launches.get(0x100 + mode, []), launches.get(0x110 + mode, []), launches.get(0x130 + mode, [])
    for i in range(min(len(S), len(W), len(E))):
        rows.append(("gemm." + modes[mode], S[i], W[i], E[i]))
S, W, E = launches.get(0x300, []), launches.get(0x310, []), launches.get(0x330, [])
for i in range(min(len(S), len(W), len(E))): rows.append(("attn", S[i], W[i], E[i]))
rows.sort(key=lambda r: min(r[2].values()))
prev_end = None
print(f"{'kernel':12s} {'ctas':>4s} {'start min':>9s} {'start max':>9s} | {'wait min':>8s} {'wait max':>8s} | {'end min':>8s} {'end p50':>8s} {'end max':>8s} | gap(prev end max -> wait min)  busy(wait min -> end max)")
for name, s, w, e in rows[: 110]:
    es = sorted(e.values())
    gap = (min(w.values()) - prev_end) if prev_end is not None else 0.0
    print(f"{name:12s} {len(s):4d} {min(s.values()):9.2f} {max(s.values()):9.2f} | {min(w.values()):8.2f} {max(w.values()):8.2f} | "
          f"{es[0]:8.2f} {es[len(es)//2]:8.2f} {es[-1]:8.2f} | {gap:6.2f}   {es[-1]-min(w.values()):6.2f}")
    prev_end = es[-1]
