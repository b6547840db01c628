"""Per-phase device timeline of the persistent decode-step kernel at the bench shape (830M, B=32): for the first layers,
when each phase's dependency resolved, when its first accumulator was ready, when this CTA finished, for CTA 0 and the last CTA.
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
for _ in range(3): sess.step()
n = 2 * nph.value * 8
buf = (C.c_uint64 * n)()
_lib.check(lib.vcb_debug_mega_timeline(sess.eng, buf, n, C.byref(nph)))
t = np.frombuffer(buf, dtype=np.uint64).reshape(2, nph.value, 8).astype(np.int64)
t0 = t[t > 0].min()
names = ["qkv", "attn", "out", "ffn1", "ffn2"]
print("ctx", 231 + steps_before + 3, "phases", nph.value, "(us since the first record; ev: dep=B producer saw the previous phase done, acc=first accumulator ready,")
print("  epi=this CTA ran a tile epilogue (last), end=this CTA's work of the phase done, prod=ring producer issued the phase's last item)")
for c in range(2):
    print("CTA", "first" if c == 0 else "last")
    prev_end = None
    for p in range(min(nph.value, 22)):
        r = t[c, p]
        us = lambda v: "%8.2f" % ((v - t0) / 1e3) if v > 0 else "    -   "
        nm = names[p % 5] if p < nph.value - 2 else ("h1" if p == nph.value - 2 else "h2")
        if nm == "attn":
            print(f"  {p:3d} {nm:5s} dep {us(r[4])} loop_end {us(r[5])} flag {us(r[6])} prod {us(r[7])}")
        else:
            print(f"  {p:3d} {nm:5s} dep {us(r[0])} acc {us(r[1])} epi {us(r[2])} end {us(r[3])} prod {us(r[7])}")
    last = t[c][t[c] > 0].max()
    print("  step kernel span (first..last record): %.1f us" % ((last - t0) / 1e3))
sess.close()
