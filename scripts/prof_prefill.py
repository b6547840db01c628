"""Per-kernel-class device time of the chunked prefill (830M, B=32 x 231 tokens) using the engine's profile mode."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voicecraft_b200 import _lib
from voicecraft_b200.voicecraft import VoiceCraft
class A: model = "830M"; batch = 32; codebooks = 4; text_len = 80; prompt = 150
cfg, sd, utts = bench.make_model_inputs(A)
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(max_slots=32, max_seq_len=1024, max_new_tokens=900)
lib = _lib.load()
xs = [u[0].cuda() for u in utts]; ys = [u[2].cuda() for u in utts]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sess = m.open_tts_session(xs, ys, top_k=40)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"prefill wall {1e3*(t1-t0):.1f} ms for {32*231} tokens")
    sess.close()
eng = m._engine()
lib.vcb_set_option(eng, b"profile", 1)
sess = m.open_tts_session(xs, ys, top_k=40)
msb = (C.c_double * 6)(); cnt = (C.c_int64 * 6)()
lib.vcb_profile_read(eng, msb, cnt, 6)
lib.vcb_set_option(eng, b"profile", 0)
names = ["gemm", "attention", "ln/gather", "-", "sampler", "misc"]
for i in range(6):
    if cnt[i]: print(f"{names[i]:10s} {msb[i]:8.2f} ms  {cnt[i]:6d} launches  avg {1e3*msb[i]/cnt[i]:7.1f} us")
sess.close()
