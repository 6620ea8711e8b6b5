"""The reference's own formulation of the hot path, restated in plain torch ops — A TIMING COMPARATOR, not product code.

bench.py runs this beside the CUDA path on the same GPU in the same process (SURVEY §8d "reference timed beside it
(1)": the comparator behind north_star's ">= 2x the reference's own GPU attention path"), and on the host cores in fp32
as the CPU baseline.  /root/reference cannot travel to the GPU box and needs diffusers for anything above
vidtome/merge.py, so the path is written out here from SURVEY.md App. A with the same op sequence the reference
launches: boolean-mask index split, `metric / metric.norm()`, two gathers, a MATERIALISED
`a @ b^T` score matrix, `max`, `argsort`, gather / cat merges, zero-initialised scatter unmerges, attention through the
module's own forward (torch SDPA), residual add.  Local merging only (BASELINE config 2).

On exact-arithmetic inputs its merged tokens and outputs are bit-identical to the library's (checked by
`check_exact`, which bench.py calls before timing anything).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch


def torch_level(x: torch.Tensor, F: int, unm_pre: int, target_stride: int, randf: int, ratio: float):
    """One bipartite_soft_matching_randframe call (vidtome/merge.py:41-159) in torch ops; returns (merge, unmerge, unm_num)."""
    B, N, C = x.shape
    tnum = (N - unm_pre) // F
    stride = min(target_stride, F)
    frame = torch.arange(N - unm_pre, device=x.device) // tnum
    is_dst = (frame % stride) == randf
    pos = torch.arange(unm_pre, N, device=x.device)
    a_idx = pos[~is_dst]                                   # boolean-mask indexing: a host sync, as in merge.py:63-64
    b_idx = torch.cat([pos[is_dst], torch.arange(unm_pre, device=x.device)])
    metric = x / x.norm(dim=-1, keepdim=True)
    a, b = metric[:, a_idx], metric[:, b_idx]
    scores = a @ b.transpose(-1, -2)                       # [B, Ns, Nd], materialised (3.2 GB at C2 ds1 level 1)
    node_max, node_idx = scores.max(dim=-1)
    edge = node_max.argsort(dim=-1, descending=True, stable=True)
    Ns = a_idx.numel()
    r = min(Ns, int(Ns * ratio))
    unm_idx, src_idx = edge[:, r:], edge[:, :r]
    dst_idx = node_idx.gather(-1, src_idx)

    def merge(t):
        src, dst = t[:, a_idx], t[:, b_idx]
        unm = src.gather(1, unm_idx[..., None].expand(-1, -1, t.shape[-1]))
        return torch.cat([unm, dst], dim=1)

    def unmerge(t):
        c = t.shape[-1]
        ul = unm_idx.shape[1]
        unm, dst = t[:, :ul], t[:, ul:]
        out = torch.zeros((B, N, c), device=t.device, dtype=t.dtype)
        out[:, b_idx] = dst
        out.scatter_(1, a_idx[unm_idx][..., None].expand(-1, -1, c), unm)
        out.scatter_(1, a_idx[src_idx][..., None].expand(-1, -1, c), dst.gather(1, dst_idx[..., None].expand(-1, -1, c)))
        return out

    return merge, unmerge, Ns - r


def torch_compute_merge(x: torch.Tensor, F: int, T: int, ratio: float, randfs: List[int], target_stride: int = 4):
    """The level loop of compute_merge (vidtome/patch.py:37-56) on joined tokens x [B, F*T, C].
    Returns (merged tokens, unmerge closure)."""
    ops_u = []
    cur, unm, curF, lvl = x, 0, F, 0
    while curF > 1:
        m, u, unm_num = torch_level(cur, curF, unm, target_stride, randfs[lvl], ratio)
        cur = m(cur)
        ops_u.append(u)
        unm += unm_num
        curF = (cur.shape[1] - unm) // T
        lvl += 1

    def unmerge(y):
        for u in reversed(ops_u):
            y = u(y)
        return y
    return cur, unmerge


def torch_merge_call(x: torch.Tensor, F: int, T: int, ratio: float, randfs: List[int], resid: torch.Tensor):
    """merge ms/call unit: compute_merge + merge + unmerge + residual, attention excluded."""
    merged, unmerge = torch_compute_merge(x, F, T, ratio, randfs)
    return merged, unmerge(merged) + resid


class _RefPathBlockMixin:
    """forward of the reference's ToMeBlock (vidtome/patch.py:128-201) with the torch-op merge above."""

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        info = self._ref_info
        norm_hidden_states = self.norm1(hidden_states)
        h, w = info["size"]
        T = hidden_states.shape[1]
        downsample = int(math.ceil(math.sqrt((h * w) // T)))
        unmerge = None
        if downsample <= info["max_downsample"]:
            B = info["batch_size"]
            F = hidden_states.shape[0] // B
            joined = norm_hidden_states.reshape(B, F * T, -1)
            gen = self._ref_generator

            def draw(stride):                                # merge.py:56-57, read back to the host like the reference
                return int(torch.randint(0, stride, (1,), generator=gen, device=gen.device))
            merged, unmerge = _compute_merge_with_draws(joined, F, T, info["ratio"], draw, info["target_stride"])
            norm_hidden_states = merged
        attn_output = self.attn1(norm_hidden_states)
        if unmerge is not None:
            attn_output = unmerge(attn_output).reshape(hidden_states.shape)
        hidden_states = attn_output + hidden_states
        if getattr(self, "attn2", None) is not None:
            hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        if getattr(self, "ff", None) is not None:
            hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


def _compute_merge_with_draws(x, F, T, ratio, draw, target_stride):
    """torch_compute_merge with one `draw(stride)` per level (vidtome/patch.py:44-54)."""
    ops_u = []
    cur, unm, curF = x, 0, F
    while curF > 1:
        rf = draw(min(target_stride, curF))
        m, u, unm_num = torch_level(cur, curF, unm, target_stride, rf, ratio)
        cur = m(cur)
        ops_u.append(u)
        unm += unm_num
        curF = (cur.shape[1] - unm) // T

    def unmerge(y):
        for u in reversed(ops_u):
            y = u(y)
        return y
    return cur, unmerge


def apply_reference_path(net: torch.nn.Module, ratio: float = 0.9, batch_size: int = 2, max_downsample: int = 2,
                         target_stride: int = 4, seed: int = 123):
    """Swap every block named BasicTransformerBlock for the torch-op restatement of the reference's patched block."""
    info = {"size": None, "ratio": ratio, "batch_size": batch_size, "max_downsample": max_downsample,
            "target_stride": target_stride}
    dev = next(net.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)

    def pre(module, args):
        info["size"] = (args[0].shape[2], args[0].shape[3])
    handle = net.register_forward_pre_hook(pre)
    for m in net.modules():
        if type(m).__name__ == "BasicTransformerBlock":
            m.__class__ = type("RefPathBlock", (_RefPathBlockMixin, m.__class__), {"_orig": m.__class__})
            m._ref_info = info
            m._ref_generator = gen
    net._ref_hook = handle
    return net


def remove_reference_path(net: torch.nn.Module):
    for m in net.modules():
        if type(m).__name__ == "RefPathBlock":
            m.__class__ = type(m)._orig
    if hasattr(net, "_ref_hook"):
        net._ref_hook.remove()
        del net._ref_hook
    return net


def check_exact(device) -> bool:
    """On exact-arithmetic inputs the restatement and the library agree bit for bit (merged tokens and output)."""
    from types import SimpleNamespace
    from vidtome_b200 import patch
    B, F, hw, C = 2, 8, 16, 128
    T = hw * hw
    g = torch.Generator(device=device).manual_seed(1)
    x = torch.zeros((B * F, T, C), device=device)
    cols = torch.rand((B * F, T, C), generator=g, device=device).argsort(-1)[..., :64]
    vals = (torch.randint(0, 2, (B * F, T, 64), generator=g, device=device).float() * 2 - 1) * 0.125
    x.scatter_(-1, cols, vals)
    x = x.half()
    module = SimpleNamespace(generator=torch.Generator(device=device).manual_seed(3), global_tokens=None)
    info = {"size": (hw, hw), "args": dict(max_downsample=2, batch_size=B, align_batch=False, merge_global=False,
                                            global_merge_ratio=0.8, local_merge_ratio=0.9, global_rand=0.5,
                                            target_stride=4)}
    plan = patch.build_merge_plan(module, x, info)
    out = plan.unmerge_add(plan.merged_tokens, x)
    xj = x.reshape(B, F * T, C)
    m2, o2 = torch_merge_call(xj, F, T, 0.9, [int(r) for r in plan.randf], xj)
    return bool(torch.equal(plan.merged_tokens, m2) and torch.equal(out.reshape(B, F * T, C), o2))
