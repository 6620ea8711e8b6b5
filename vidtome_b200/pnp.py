"""PnP self-attention on merged tokens.

The reference's PnP control (utils/pnp_utils.py:39-106, `register_attention_control`) replaces `attn1.forward` of the
decoder blocks by a closure that — while the current timestep `attn.t` is in the injection schedule — computes the
attention map from the SOURCE sample only (`q[:B/num_inputs]`, `k[:B/num_inputs]`) and applies it to the values of
all `num_inputs` samples (`attn.repeat(num_inputs, 1, 1)`).  With `align_batch=True` (generate.py:97-98: `use_pnp or
cfg`) the merged token sets of the samples are aligned, so this is well defined on merged tokens.

`register_attention_control(model, injection_schedule, num_inputs)` here has the reference's name and arguments.  It
installs (i) an equivalent torch forward for every call that does not come through a merged ToMeBlock, and (ii) a marker
(`attn._vtm_pnp`) that lets ToMeBlock run the same computation through KD (`vtm_attention_ex`, flag
VTM_ATTN_SHARED_QK: every sample reads the queries and keys of sample 0 and its own values).  `register_time(model, t)`
sets `attn.t` on every marked module.

The decoder-block addressing of the reference (`model.unet.up_blocks[res].attentions[block].transformer_blocks[0].attn1`
for res/block in {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}) is used when the model has that hierarchy; any other
model (the skeleton) must pass the attention modules explicitly through `modules=`.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch

RES_DICT = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}       # utils/pnp_utils.py:97 (blocks 4 - 11 of the decoder)


def _targets(model, modules: Optional[Iterable[torch.nn.Module]]):
    if modules is not None:
        return list(modules)
    unet = model.unet if hasattr(model, "unet") else model
    out = []
    for res, blocks in RES_DICT.items():
        for block in blocks:
            out.append(unet.up_blocks[res].attentions[block].transformer_blocks[0].attn1)
    return out


def injection_active(attn: torch.nn.Module) -> bool:
    """utils/pnp_utils.py:57-58: `self.injection_schedule is not None and (self.t in self.injection_schedule or self.t == 1000)`."""
    sched = getattr(attn, "injection_schedule", None)
    t = getattr(attn, "t", None)
    return sched is not None and t is not None and (t in sched or t == 1000)


def _torch_forward(attn: torch.nn.Module, num_inputs: int):
    """The reference's replaced forward (utils/pnp_utils.py:47-95) in torch ops; used outside merged blocks."""
    to_out = attn.to_out[0] if isinstance(attn.to_out, (torch.nn.ModuleList, torch.nn.Sequential)) else attn.to_out

    def heads_first(t, h):
        b, n, c = t.shape
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3)

    def forward(x, encoder_hidden_states=None, attention_mask=None, **kwargs):
        h = attn.heads
        is_cross = encoder_hidden_states is not None
        ctx = encoder_hidden_states if is_cross else x
        q, k, v = attn.to_q(x), attn.to_k(ctx), attn.to_v(ctx)
        inject = not is_cross and injection_active(attn)
        if inject:
            src = q.shape[0] // num_inputs
            q, k = q[:src], k[:src]
        sim = torch.einsum("bhid,bhjd->bhij", heads_first(q, h), heads_first(k, h)) * attn.scale
        if attention_mask is not None:
            mask = attention_mask.reshape(x.shape[0], -1)[:, None, None, :]
            sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
        p = sim.softmax(dim=-1)
        if inject:
            p = p.repeat(num_inputs, 1, 1, 1)
        out = torch.einsum("bhij,bhjd->bhid", p, heads_first(v, h))
        out = out.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)
        return to_out(out)

    forward._vtm_pnp_forward = True
    return forward


def register_attention_control(model, injection_schedule: Optional[Sequence[int]], num_inputs: int,
                               modules: Optional[Iterable[torch.nn.Module]] = None):
    """Same call as the reference's (utils/pnp_utils.py:39): mark the decoder self-attention modules for source-sample
    Q/K injection during the timesteps in `injection_schedule`."""
    for attn in _targets(model, modules):
        attn.forward = _torch_forward(attn, num_inputs)
        attn.injection_schedule = injection_schedule
        attn._vtm_pnp = int(num_inputs)
    return model


def register_time(model, t: int, modules: Optional[Iterable[torch.nn.Module]] = None):
    """utils/pnp_utils.py:12-37 for the self-attention modules: the current timestep, read by the injection test."""
    if modules is None:
        root = model.unet if hasattr(model, "unet") else model
        modules = [m for m in root.modules() if hasattr(m, "_vtm_pnp")]
    for attn in modules:
        attn.t = t
    return model
