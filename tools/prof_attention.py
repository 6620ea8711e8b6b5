"""Run KD once per shape (for ncu captures): python tools/prof_attention.py [L C heads]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import ops

L, C, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (10241, 320, 8)
B = 2
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, L, C), generator=g, device="cuda").half()
ws = [(torch.randn((C, C), generator=g, device="cuda") / C ** 0.5).half() for _ in range(4)]
wqkv = torch.cat(ws[:3], 0).contiguous()
bo = torch.zeros(C, device="cuda").half()
for _ in range(3):
    y = ops.attention(x, wqkv, ws[3], bo, H, (C // H) ** -0.5)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
