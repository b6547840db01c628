"""One EnCodec decode under the CUDA profiler API (for `ncu --profile-from-start off`): usage prof_codec.py [B] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encodec_oracle as eo
from voicecraft_b200.tokenizer import AudioTokenizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = eo.default_config()
tok = AudioTokenizer(device="cuda:0", config=cfg, state_dict=eo.make_state_dict(cfg, seed=0))
codes = torch.randint(0, 2048, (B, 4, T), generator=torch.Generator().manual_seed(0)).cuda()
tok.decode_codes(codes); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tok.decode_codes(codes); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", B, T)
