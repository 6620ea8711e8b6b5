"""SURVEY §8(d) metric 2: merge ms/call — one compute_merge (all levels) + merge + unmerge(+residual) for one
transformer block, attention excluded — for this library and, beside it on the same GPU, for the reference's own
GPU formulation restated in plain torch ops (normalise, index split, `a @ b^T` into a materialised score matrix,
`max`, `argsort`, gathers / cat; zero-initialised scatter unmerge), written from SURVEY.md App. A.  The torch
restatement exists for timing only; on exact-arithmetic inputs its merged tokens are checked against ours.

  python tools/merge_call_bench.py            # C2 ds1 and ds2 shapes
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import patch  # noqa: E402


# ------------------------------------------------------------------ the reference's GPU path in torch ops
def torch_level(x, F, unm_pre, target_stride, randf, ratio):
    B, N, C = x.shape
    tnum = (N - unm_pre) // F
    stride = min(target_stride, F)
    frame = torch.arange(N - unm_pre, device=x.device) // tnum
    is_dst = (frame % stride) == randf
    pos = torch.arange(unm_pre, N, device=x.device)
    a_idx = pos[~is_dst]
    b_idx = torch.cat([pos[is_dst], torch.arange(unm_pre, device=x.device)])
    metric = x / x.norm(dim=-1, keepdim=True)
    a, b = metric[:, a_idx], metric[:, b_idx]
    scores = a @ b.transpose(-1, -2)                       # [B, Ns, Nd] fp16, materialised
    node_max, node_idx = scores.max(dim=-1)
    edge = node_max.argsort(dim=-1, descending=True, stable=True)
    Ns = a_idx.numel()
    r = min(Ns, int(Ns * ratio))
    unm_idx, src_idx = edge[:, r:], edge[:, :r]
    dst_idx = node_idx.gather(-1, src_idx)

    def merge(t):
        src, dst = t[:, a_idx], t[:, b_idx]
        unm = src.gather(1, unm_idx[..., None].expand(-1, -1, C))
        return torch.cat([unm, dst], dim=1)

    def unmerge(t):
        ul = unm_idx.shape[1]
        unm, dst = t[:, :ul], t[:, ul:]
        out = torch.zeros((B, N, C), device=t.device, dtype=t.dtype)
        out[:, b_idx] = dst
        out.scatter_(1, a_idx[unm_idx][..., None].expand(-1, -1, C), unm)
        out.scatter_(1, a_idx[src_idx][..., None].expand(-1, -1, C), dst.gather(1, dst_idx[..., None].expand(-1, -1, C)))
        return out

    return merge, unmerge, Ns - r


def torch_merge_call(x, F, T, ratio, randfs, resid):
    """x [B, F*T, C] (norm1 output), two-level local merge as compute_merge does, then unmerge + residual."""
    ops_m, ops_u = [], []
    cur, unm, curF, lvl = x, 0, F, 0
    while curF > 1:
        m, u, unm_num = torch_level(cur, curF, unm, 4, randfs[lvl], ratio)
        cur = m(cur)
        ops_m.append(m)
        ops_u.append(u)
        unm += unm_num
        curF = (cur.shape[1] - unm) // T
        lvl += 1
    merged = cur
    y = merged                                            # (attention excluded)
    for u in reversed(ops_u):
        y = u(y)
    return merged, y + resid


# ------------------------------------------------------------------ ours
class _Block(torch.nn.Module):
    pass


def ours_merge_call(module, info, x_frames, resid_frames):
    plan = patch.build_merge_plan(module, x_frames, info)
    return plan.merged_tokens, plan.unmerge_add(plan.merged_tokens, resid_frames), plan


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def run(F, hw, C, ratio, downsample, label):
    B = 2
    T = (hw // downsample) ** 2
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(123)
    base = torch.randn((B, 1, T, C), generator=g, device=dev)
    x = (base + 0.1 * torch.randn((B, F, T, C), generator=g, device=dev)).reshape(B * F, T, C).half()   # video-like
    module = _Block()
    module.generator = torch.Generator(device=dev).manual_seed(7)
    info = {"size": (hw, hw), "args": dict(max_downsample=2, batch_size=B, align_batch=False, merge_global=False,
                                            global_merge_ratio=0.8, local_merge_ratio=ratio, global_rand=0.5,
                                            target_stride=4)}
    # same draws for both: read ours back once
    _, _, plan = ours_merge_call(module, info, x, x)
    randfs = [int(r) for r in plan.randf]
    xj = x.reshape(B, F * T, C)
    ms_ours = timeit(lambda: ours_merge_call(module, info, x, x))
    ms_torch = timeit(lambda: torch_merge_call(xj, F, T, ratio, randfs, xj), iters=5, warm=2)
    print(json.dumps({"metric": "merge ms/call (compute_merge + merge + unmerge + residual, attention excluded)",
                      "shape": label, "B": B, "F": F, "T": T, "C": C, "ratio": ratio,
                      "merged_len": int(plan.merged_tokens.shape[1]),
                      "ours_ms": round(ms_ours, 3), "torch_ops_reference_path_ms": round(ms_torch, 3),
                      "speedup": round(ms_torch / ms_ours, 2)}), flush=True)


def check_exact():
    """On exact-arithmetic inputs the torch restatement and the library agree bit for bit (merged tokens, output)."""
    B, F, hw, C = 2, 8, 16, 128
    T = hw * hw
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.zeros((B * F, T, C), device=dev)
    cols = torch.rand((B * F, T, C), generator=g, device=dev).argsort(-1)[..., :64]
    vals = (torch.randint(0, 2, (B * F, T, 64), generator=g, device=dev).float() * 2 - 1) * 0.125
    x.scatter_(-1, cols, vals)
    x = x.half()
    module = _Block()
    module.generator = torch.Generator(device=dev).manual_seed(3)
    info = {"size": (hw, hw), "args": dict(max_downsample=2, batch_size=B, align_batch=False, merge_global=False,
                                            global_merge_ratio=0.8, local_merge_ratio=0.9, global_rand=0.5,
                                            target_stride=4)}
    merged, out, plan = ours_merge_call(module, info, x, x)
    xj = x.reshape(B, F * T, C)
    m2, o2 = torch_merge_call(xj, F, T, 0.9, [int(r) for r in plan.randf], xj)
    ok = bool(torch.equal(merged, m2) and torch.equal(out.reshape(B, F * T, C), o2))
    print(json.dumps({"check": "torch restatement == library on exact-arithmetic input", "ok": ok}), flush=True)
    return ok


if __name__ == "__main__":
    if not check_exact():
        sys.exit(1)
    run(16, 64, 320, 0.9, 1, "C2 ds1")
    run(16, 64, 640, 0.9, 2, "C2 ds2")
