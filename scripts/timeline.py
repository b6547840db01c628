"""(needs a library built with the device-timeline marks: `make -C voicecraft_b200/csrc clean all TIMELINE=1`)
Dump a device-side timeline of a few decode steps (830M, B=32) -- which kernels overlap under PDL."""
import ctypes as C, os, sys
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
lib.vcb_timeline(1, None, 0, None)
for _ in range(3): sess.step()
buf = (C.c_uint64 * (2 * 65536))(); n = C.c_int32()
lib.vcb_timeline(0, buf, 65536, C.byref(n))
recs = sorted(((buf[2*i+1], buf[2*i]) for i in range(n.value)))
t0 = recs[0][0]
names = {0x100:"gemm.start",0x110:"gemm.waited",0x120:"gemm.acc_ready",0x130:"gemm.end",0x140:"gemm.dsmem_sent",0x150:"gemm.cluster_synced",0x160:"gemm.epi_done",0x200:"ln.start",0x210:"ln.waited",0x230:"ln.end",
         0x300:"attn.start",0x310:"attn.waited",0x330:"attn.end",0x400:"samp.start",0x410:"samp.waited",0x430:"samp.finish_slot_done"}
modes = ["qkv","resid","act","logits"]
def name(tag):
    base = tag & ~0xf if (tag & 0xf00) == 0x100 else tag
    nm = names.get(base, hex(tag))
    if (tag & 0xf00) == 0x100: nm += "." + modes[tag & 0xf]
    return nm
for t, tag in recs[: 120]:
    print(f"{(t - t0)/1000:10.2f} us  {name(tag)}")
# the step boundary: last layers -> heads -> sampler -> next step's first kernels
idx = [i for i, (t, tag) in enumerate(recs) if tag == 0x400]
if idx:
    print("---- around the first sampler ----")
    for t, tag in recs[max(0, idx[0] - 40): idx[0] + 25]:
        print(f"{(t - t0)/1000:10.2f} us  {name(tag)}")
