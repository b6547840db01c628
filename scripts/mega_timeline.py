"""Per-phase device timeline of the persistent decode-step kernel at the bench shape (830M, B=32), over ALL CTAs.
Events per (cta, phase): 0 dep (B producer saw the previous phase complete), 1 acc (first accumulator of the phase ready),
2 epi (this CTA finished its share of a tile's rows: reduce + epilogue + count), 8 all contributors of the tile have arrived, 3 end (this CTA's last segment handed over), 7 ring producer issued
the phase's last item; GEMM phases: 4 B producer issued its last activation tile, 5 first MMA issued, 6 last MMA issued;
attention phases: 4 dependency seen, 5 loop end, 6 flag published.
usage: python scripts/mega_timeline.py [steps_before] [kv]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from voicecraft_b200 import _lib
from voicecraft_b200.voicecraft import VoiceCraft
steps_before = int(sys.argv[1]) if len(sys.argv) > 1 else 300
class A: model = "830M"; batch = 32; codebooks = 4; text_len = 80; prompt = 150; workload = "tts"
cfg, sd = bench.make_model(A)
utts = bench.make_utterances(A, cfg, range(32))
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(max_slots=32, max_seq_len=1024, max_new_tokens=900, kv_dtype=sys.argv[2] if len(sys.argv) > 2 else "bf16")
sess = m.open_tts_session([u[0].cuda() for u in utts], [u[2].cuda() for u in utts], seeds=list(range(1, 33)), top_k=40)
lib = _lib.load()
sess.sample()
for _ in range(steps_before): sess.step()
nph = C.c_int32()
_lib.check(lib.vcb_debug_mega_timeline(sess.eng, None, 0, C.byref(nph)))
G = int(lib.vcb_counter(sess.eng, b"mega_grid"))
for _ in range(3): sess.step()
n = G * nph.value * 16
buf = (C.c_uint64 * n)()
_lib.check(lib.vcb_debug_mega_timeline(sess.eng, buf, n, C.byref(nph)))
t = np.frombuffer(buf, dtype=np.uint64).reshape(G, nph.value, 16).astype(np.float64)
t[t == 0] = np.nan
dur = t[:, :, 12:15].copy() / 1e3          # accumulated waits (us): 12 workers on full barriers, 13 producer on empty slots, 14 producer on the in-flight cap
t[:, :, 12:15] = np.nan
t0 = np.nanmin(t[:, 0, :])
t = (t - t0) / 1e3
names = ["qkv", "attn", "out", "ffn1", "ffn2"]
print("ctx", 231 + steps_before + 3, "grid", G, "phases", nph.value, " -- microseconds since the step's first record; min / median / max over CTAs")
f = lambda a: "   -  /   -  /   -  " if np.all(np.isnan(a)) else "%6.1f/%6.1f/%6.1f" % (np.nanmin(a), np.nanmedian(a), np.nanmax(a))
for p in range(nph.value):
    nm = names[p % 5] if p < nph.value - 2 else ("h1" if p == nph.value - 2 else "h2")
    if p >= 12 and p < nph.value - 7: continue
    if nm == "attn":
        print(f"{p:3d} {nm:5s} dep {f(t[:, p, 4])}  loop_end {f(t[:, p, 5])}  flag {f(t[:, p, 6])}  prod {f(t[:, p, 7])}  | waits: workers on data {f(dur[:, p, 0])}  producer on free slots {f(dur[:, p, 1])}  on flight cap {f(dur[:, p, 2])}")
    else:
        print(f"        hand-over (us): first accumulator -> all contributors of the tile arrived {f(t[:, p, 8] - t[:, p, 1])}   -> my rows reduced, epilogue done, counted {f(t[:, p, 2] - t[:, p, 8])}")
        print(f"{p:3d} {nm:5s} dep {f(t[:, p, 0])}  mma0 {f(t[:, p, 5])}  Blast {f(t[:, p, 4])}  mmaN {f(t[:, p, 6])}  acc {f(t[:, p, 1])}  end {f(t[:, p, 3])}  epi {f(t[:, p, 2])}  prod {f(t[:, p, 7])}")
print("kernel span: %.1f us" % np.nanmax(t))
# per-layer summary over the middle layers
per = []
for l in range(2, nph.value // 5 - 1):
    per.append(np.nanmin(t[:, 5 * (l + 1), 0]) - np.nanmin(t[:, 5 * l, 0]))
print("per-layer (qkv dep -> next qkv dep): mean %.1f us over %d layers" % (np.mean(per), len(per)))
sess.close()
