"""One launch of each KA variant at the benchmark's level-1 shape, for an `ncu --set full` capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import ops  # noqa: E402

a = torch.randn(2, 49152, 320, device="cuda").half()
b = torch.randn(2, 16384, 320, device="cuda").half()
for v in (None, "1"):
    if v:
        ops.KA_VARIANT = "pair"
    else:
        ops.KA_VARIANT = "cta"
    for _ in range(2):
        ops.sim_argmax(a, b, False)
    torch.cuda.synchronize()
