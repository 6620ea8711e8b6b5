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
