"""Sweep the key-range split count of the last attention wave (VTM_FA_SPLITS override) at the benchmark shapes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import ops  # noqa: E402


def run(B, L, C, H, splits_list, rounds=5, inner=10):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((B, L, C), generator=g, device="cuda").half()
    ws = [(torch.randn((C, C), generator=g, device="cuda") / C ** 0.5).half() for _ in range(4)]
    wqkv = torch.cat(ws[:3], 0).contiguous()
    bo = torch.zeros(C, device="cuda").half()
    res = {sp: [] for sp in splits_list}
    for r in range(rounds + 1):
        for sp in splits_list:
            if sp == 0:
                os.environ.pop("VTM_FA_SPLITS", None)
            else:
                os.environ["VTM_FA_SPLITS"] = str(sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(inner):
                ops.attention(x, wqkv, ws[3], bo, H, (C // H) ** -0.5)
            e1.record()
            torch.cuda.synchronize()
            if r:
                res[sp].append(e0.elapsed_time(e1) / inner)
    print(json.dumps({"shape": [B, L, C, H],
                      "ms_by_splits(0=auto)": {sp: round(sorted(v)[len(v) // 2], 4) for sp, v in res.items()}}), flush=True)


if __name__ == "__main__":
    run(2, 2561, 640, 8, [0, 1, 2, 3, 4, 5, 6, 7, 8])
    run(2, 10241, 320, 8, [0, 1, 2])
    run(2, 5325, 640, 8, [0, 1, 2, 3, 4])     # C3-like ds2
    run(2, 21300, 320, 8, [0, 1, 2, 4, 8])    # longer ds1
