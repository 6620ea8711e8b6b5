"""Host side of KD: run a diffusers-style `Attention` module's self-attention on merged tokens through
the CUDA kernel (`vtm_attention`) instead of the module's own forward.

`ToMeBlock.forward` uses this only for a stock attention module (see patch._plain_attention_module);
anything customised (PnP's replaced forward, attention masks, cross_attention_kwargs, added KV
projections) goes through the module itself, exactly as the reference does at patch.py:157-162.
"""
from __future__ import annotations

import torch

from . import _lib, ops

# Switched on once the tcgen05 attention kernel is built into the library.
ENABLED = True


def _packed_weights(attn: torch.nn.Module, like: torch.Tensor):
    """[3C, C] fp16 (Wq | Wk | Wv), Wo [C, C], bo [C] — cached on the module, refreshed when a weight
    tensor is replaced or modified in place."""
    ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias)
    tag = tuple((w.data_ptr(), w._version) for w in ws if w is not None) + (like.device,)
    cache = getattr(attn, "_vtm_packed", None)
    if cache is None or cache[0] != tag:
        w_qkv = torch.cat([ws[0], ws[1], ws[2]], dim=0).to(device=like.device, dtype=torch.float16).contiguous()
        w_o = ws[3].to(device=like.device, dtype=torch.float16).contiguous()
        b_o = None if ws[4] is None else ws[4].to(device=like.device, dtype=torch.float16).contiguous()
        cache = (tag, w_qkv, w_o, b_o)
        attn._vtm_packed = cache
    return cache[1:]


def self_attention(attn: torch.nn.Module, x: torch.Tensor, shared_qk: bool = False) -> torch.Tensor:
    """softmax(q k^T * scale) v with q,k,v = to_q/to_k/to_v(x), then to_out[0] (utils/pnp_utils.py:47-95).
    shared_qk: PnP injection — q and k of sample 0 for every sample (utils/pnp_utils.py:57-68,87-91)."""
    w_qkv, w_o, b_o = _packed_weights(attn, x)
    heads = int(attn.heads)
    scale = float(getattr(attn, "scale", (x.shape[-1] // heads) ** -0.5))
    return ops.attention(x, w_qkv, w_o, b_o, heads, scale, shared_qk=shared_qk)


def cross_attention_eligible(attn: torch.nn.Module, ctx: torch.Tensor) -> bool:
    """True if `attn` is a stock diffusers-style cross-attention (to_q [C, C], to_k / to_v [C, Cctx] without bias,
    to_out = [Linear, Dropout], default processor, no hooks / overrides / options that alter the math)."""
    from .patch import _PLAIN_PROCESSORS
    if attn is None or "forward" in vars(attn) or not isinstance(ctx, torch.Tensor) or ctx.dim() != 3:
        return False
    need = ("to_q", "to_k", "to_v", "to_out", "heads")
    if not all(hasattr(attn, n) for n in need):
        return False
    to_out = attn.to_out
    if not isinstance(to_out, (torch.nn.ModuleList, torch.nn.Sequential)) or len(to_out) < 1:
        return False
    linears = (attn.to_q, attn.to_k, attn.to_v, to_out[0])
    if any(type(m) is not torch.nn.Linear for m in linears) or any(m.bias is not None for m in linears[:3]):
        return False
    for m in (attn,) + linears + tuple(to_out[1:]):
        if m._forward_hooks or m._forward_pre_hooks:
            return False
    for extra in to_out[1:]:
        if not isinstance(extra, torch.nn.Dropout) or (extra.p > 0 and extra.training):
            return False
    proc = getattr(attn, "processor", None)
    if proc is not None and type(proc).__name__ not in _PLAIN_PROCESSORS:
        return False
    for name, bad in (("group_norm", None), ("spatial_norm", None), ("norm_cross", None), ("added_kv_proj_dim", None)):
        if getattr(attn, name, None) not in (None, False):
            return False
    if getattr(attn, "residual_connection", False) or getattr(attn, "rescale_output_factor", 1.0) != 1.0:
        return False
    C = attn.to_q.weight.shape[1]
    Cctx = attn.to_k.weight.shape[1]
    if attn.to_q.weight.shape != (C, C) or attn.to_k.weight.shape != (C, Cctx) or attn.to_v.weight.shape != (C, Cctx):
        return False
    if to_out[0].weight.shape != (C, C) or ctx.shape[-1] != Cctx or Cctx % 8 != 0:
        return False
    if attn.to_q.weight.dtype != torch.float16 or not attn.to_q.weight.is_cuda or ctx.dtype != torch.float16:
        return False
    d = C // int(attn.heads)
    return d * int(attn.heads) == C and d % 8 == 0 and d <= 128


def cross_attention_residual(attn: torch.nn.Module, x: torch.Tensor, ctx: torch.Tensor, resid: torch.Tensor) -> torch.Tensor:
    """attn2(x, encoder_hidden_states=ctx) + resid (vidtome/patch.py:171-185) through vtm_cross_attention."""
    ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias)
    tag = tuple((w.data_ptr(), w._version) for w in ws if w is not None)
    cache = getattr(attn, "_vtm_packed_x", None)
    if cache is None or cache[0] != tag:
        w_kv = torch.cat([ws[1], ws[2]], dim=0).detach().contiguous()
        cache = (tag, ws[0].detach().contiguous(), w_kv, ws[3].detach().contiguous(),
                 None if ws[4] is None else ws[4].detach().contiguous())
        attn._vtm_packed_x = cache
    _, w_q, w_kv, w_o, b_o = cache
    heads = int(attn.heads)
    scale = float(getattr(attn, "scale", (x.shape[-1] // heads) ** -0.5))
    return ops.cross_attention(x.contiguous(), ctx.contiguous(), w_q, w_kv, w_o, b_o, heads, scale, resid=resid.contiguous())
