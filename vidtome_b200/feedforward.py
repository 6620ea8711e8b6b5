"""Host side of the block's feed-forward on tcgen05 (vidtome/patch.py:187-199).

`ToMeBlock.forward` uses this for a stock GEGLU feed-forward — diffusers' `FeedForward` with
`net = [GEGLU(dim, 4 dim), Dropout, Linear(4 dim, dim)]`, or the skeleton's stand-in with the same two Linear layers —
preceded by a plain LayerNorm; anything customised (other activations, LoRA-wrapped layers, hooks, AdaNorm) goes through
the modules themselves exactly as the reference does.  Three launches: LayerNorm (row kernel), GEGLU projection
(tcgen05 GEMM with the gate fused into its epilogue: the [M, 8 dim] projection is never written), output projection with
bias and residual fused into its epilogue.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def _plain_linear(m) -> bool:
    return type(m) is torch.nn.Linear and not (m._forward_hooks or m._forward_pre_hooks)


def geglu_parts(ff: torch.nn.Module) -> Optional[Tuple[torch.nn.Linear, torch.nn.Linear]]:
    """(projection Linear dim -> 2 inner, output Linear inner -> dim) of a stock GEGLU feed-forward, else None."""
    if ff is None or "forward" in vars(ff) or ff._forward_hooks or ff._forward_pre_hooks:
        return None
    proj = out = None
    net = getattr(ff, "net", None)
    if isinstance(net, (torch.nn.ModuleList, torch.nn.Sequential)) and len(net) >= 2:
        act = net[0]
        if type(act).__name__ != "GEGLU" or not hasattr(act, "proj") or act._forward_hooks or act._forward_pre_hooks:
            return None
        if getattr(act, "approximate", "none") not in ("none", None):
            return None
        proj = act.proj
        rest = list(net[1:])
        linears = [m for m in rest if isinstance(m, torch.nn.Linear)]
        drops = [m for m in rest if isinstance(m, torch.nn.Dropout)]
        if len(linears) != 1 or len(linears) + len(drops) != len(rest):
            return None
        if any(m.p > 0 and m.training for m in drops):
            return None
        out = linears[0]
    elif type(ff).__name__ == "GEGLUFeedForward" and hasattr(ff, "proj") and hasattr(ff, "out"):
        proj, out = ff.proj, ff.out
    else:
        return None
    if not (_plain_linear(proj) and _plain_linear(out)):
        return None
    inner, dim = out.weight.shape[1], out.weight.shape[0]
    if proj.weight.shape != (2 * inner, dim) or inner % 32 != 0 or dim % 8 != 0:
        return None
    if proj.weight.dtype != torch.float16 or not proj.weight.is_cuda:
        return None
    return proj, out


def _packed(ff: torch.nn.Module, proj: torch.nn.Linear, out: torch.nn.Linear):
    ws = (proj.weight, proj.bias, out.weight, out.bias)
    tag = tuple((w.data_ptr(), w._version) for w in ws if w is not None)
    cache = getattr(ff, "_vtm_packed", None)
    if cache is None or cache[0] != tag:
        w_il, b_il = ops.interleave_geglu(proj.weight.detach(), None if proj.bias is None else proj.bias.detach())
        cache = (tag, w_il, b_il, out.weight.detach().contiguous(), None if out.bias is None else out.bias.detach().contiguous())
        ff._vtm_packed = cache
    return cache[1:]


def feed_forward_residual(ff: torch.nn.Module, parts, ln, hidden_states: torch.Tensor) -> torch.Tensor:
    """hidden_states + ff(LayerNorm(hidden_states)) for a stock GEGLU feed-forward (patch.py:187-199).
    `ln` = (weight, bias, eps) of norm3; hidden_states [(B F), T, C] fp16 CUDA."""
    proj, out = parts
    w_il, b_il, w_o, b_o = _packed(ff, proj, out)
    shape = hidden_states.shape
    h2 = hidden_states.contiguous().view(-1, shape[-1])
    n3 = ops.layer_norm(h2, ln)
    u = ops.linear_geglu(n3, w_il, b_il)
    return ops.linear_residual(u, w_o, b_o, h2).view(shape)
