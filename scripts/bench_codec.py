"""EnCodec SEANet decode micro-benchmark (BASELINE.json configs[3] shape: K=4, 16 s = 800 frames per utterance).
usage: bench_codec.py [B] [T]   -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encodec_oracle as eo
from voicecraft_b200 import _lib
from voicecraft_b200.tokenizer import AudioTokenizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
cfg = eo.default_config()
tok = AudioTokenizer(device="cuda:0", config=cfg, state_dict=eo.make_state_dict(cfg, seed=0))
codes = torch.randint(0, 2048, (B, 4, T), generator=torch.Generator().manual_seed(0)).cuda()
tok.decode_codes(codes[:1, :, :50]); torch.cuda.synchronize()
for _ in range(2): tok.decode_codes(codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n): wav = tok.decode_codes(codes)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = _lib.load().enc_counter(tok._engine(), b"flops_per_frame") * B * T
print(json.dumps({"workload": f"EnCodec SEANet decode B={B} x {T} frames (K=4, 16 kHz)", "ms": ms, "utt_per_s": B / ms * 1e3,
                  "audio_s_per_s": B * T / 50 / ms * 1e3, "codec_tokens_per_s": B * T * 4 / ms * 1e3,
                  "tflops_fp32": fl / ms / 1e9, "flops": fl, "tc_path": bool(_lib.load().enc_counter(tok._engine(), b"tc_enabled"))}))
