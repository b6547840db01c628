"""Profiling target: 830M, B=32 decode.  Runs prefill + WARM decode steps, then N steps inside cudaProfilerStart/Stop
(use ncu --profile-from-start off).  usage: prof_decode.py [warm_steps] [profiled_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voicecraft_b200.voicecraft import VoiceCraft
class A: model = "830M"; batch = 32; codebooks = 4; text_len = 80; prompt = 150
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nprof = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg, sd, utts = bench.make_model_inputs(A)
m = VoiceCraft(cfg); m.load_state_dict(sd); m = m.cuda().eval()
m.configure_engine(max_slots=32, max_seq_len=1024, max_new_tokens=900)
sess = m.open_tts_session([u[0].cuda() for u in utts], [u[2].cuda() for u in utts], top_k=40)
sess.sample()
for _ in range(warm): sess.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(nprof): sess.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ctx", 231 + warm, "done")
