"""Small invocations of every kernel family, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import ops, patch  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
# projections with ragged tiles (staged row stores), with and without bias
for (M, N, K) in [(333, 320, 320), (1000, 960, 320), (130, 1920, 640)]:
    a = torch.randn((M, K), generator=g, device="cuda").half()
    w = torch.randn((N, K), generator=g, device="cuda").half()
    b = torch.randn((N,), generator=g, device="cuda").half()
    y = ops.linear(a, w, b)
    assert torch.isfinite(y).all()
# attention: one plain shape, one whose last wave is split over key ranges
for (B, L, C, H) in [(1, 333, 320, 8), (2, 2561, 640, 8)]:
    x = torch.randn((B, L, C), generator=g, device="cuda").half()
    ws = [(torch.randn((C, C), generator=g, device="cuda") / C ** 0.5).half() for _ in range(4)]
    y = ops.attention(x, torch.cat(ws[:3], 0).contiguous(), ws[3], torch.zeros(C, device="cuda").half(), H, (C // H) ** -0.5)
    assert torch.isfinite(y).all()
# the merge plan (K0, KA with the filtered epilogue, sort, maps, KC, KE), both KA builds
info = {"size": (16, 16), "args": dict(max_downsample=2, batch_size=2, align_batch=False, merge_global=False,
                                        global_merge_ratio=0.8, local_merge_ratio=0.9, global_rand=0.5, target_stride=4)}
x = torch.randn((2 * 8, 256, 320), generator=g, device="cuda").half()
for variant in ("cta", "pair"):
    ops.KA_VARIANT = variant
    mod = SimpleNamespace(generator=torch.Generator(device="cuda").manual_seed(1), global_tokens=None)
    plan = patch.build_merge_plan(mod, x, info)
    out = plan.unmerge_add(plan.merged_tokens, x)
    assert torch.isfinite(out).all()
torch.cuda.synchronize()
print("sanitize_small: ok")
