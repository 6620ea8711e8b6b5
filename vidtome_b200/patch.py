"""Patch layer — the drop-in boundary.  Same API as the reference's vidtome/patch.py:
`apply_patch / remove_patch / update_patch / collect_from_patch`, the `_tome_info` dict, the
`module.generator` / `module.global_tokens` attributes and the `ToMeBlock` class swap (patch.py:206-387).

What differs from the reference is inside `compute_merge` and the self-attention section of
`ToMeBlock.forward` (patch.py:14-91, :139-169):
  * every matching level runs as CUDA kernels through the C-ABI (merge.match_level);
  * the per-level merge/unmerge closures are never chained: their int32 index maps are composed on the
    device (KB2), so the merged tokens are produced by ONE row gather of the block input (KC) and the
    unmerge + split_frame + residual add is ONE gather-add pass (KE);
  * `module.global_tokens` stays in HBM (the reference parks it on the CPU, patch.py:80,82).
The RNG protocol is unchanged: one `torch.randint` per local level and one `torch.rand` per block when
global tokens exist, drawn from `module.generator` on the generator's device (merge.py:56-57, patch.py:62).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple, Type

import torch

from . import merge, ops
from ._lib import VtmSplit
from .utils import (draw_randf, draw_scalar, func_warper, init_generator, isinstance_str, join_frame, join_warper,
                    split_frame, split_warper)


@dataclass
class MergePlan:
    """Composed result of compute_merge for one block call."""
    fsize: int
    N0: int                               # F * T tokens per sample before merging
    merged_tokens: torch.Tensor           # [B, L, C] fp16, input of attn1
    pi: torch.Tensor                      # [B'|1, N0] int32: out[b, p] = y[b, pi[b, p]]
    levels: List[merge.LevelMatch] = field(default_factory=list)
    randf: List[Any] = field(default_factory=list)      # ints (eager) or 1-element int32 CUDA tensors (graph capture)
    coin: Optional[float] = None
    mu: Optional[torch.Tensor] = None     # [B'|1, L_local] int32: local merged token i = table row mu[i] (before a global stage)

    def unmerge(self, y: torch.Tensor, **kwarg) -> torch.Tensor:
        """u_a of the reference (patch.py:85,168): all unmerges + split_frame, one gather pass."""
        out = ops.unmerge_add(y.contiguous(), self.pi, None)
        return split_frame(out, self.fsize)

    def unmerge_add(self, y: torch.Tensor, hidden_states: torch.Tensor) -> torch.Tensor:
        """patch.py:168-169 fused: u_a(attn_output) + hidden_states in one pass (KE)."""
        resid = join_frame(hidden_states.contiguous(), self.fsize)
        out = ops.unmerge_add(y.contiguous(), self.pi, resid)
        return split_frame(out, self.fsize)


# Fuse a plain nn.LayerNorm norm1 into the K0 / KC kernels (its output is then never materialised).
FUSE_LAYERNORM = True

# Run a stock GEGLU feed-forward (norm3 + ff + residual, patch.py:187-199) through the tcgen05 kernels (feedforward.py).
FUSE_FEED_FORWARD = True
# Run a stock cross-attention (norm2 + attn2 + residual, patch.py:171-185) through the CUDA library (attention.py).
FUSE_CROSS_ATTENTION = True

# How `merge_global=True` obtains the global token set when several ranks each hold one chunk:
#   "recurrence" — the reference's semantics (patch.py:59-82): whatever the previous chunk processed by THIS
#                  process left in module.global_tokens;
#   "allgather"  — chunk-per-GPU variant (option A): one NCCL all-gather of the local merged tokens
#                  per merged block; rank k matches against the tokens of rank (k-1) mod G (dist.py);
#   "p2p"        — the same exchange fused into the kernel that produces the tokens: KC stores each merged row into
#                  the peer-mapped buffer of its consumer, rank k+1, over NVLink (one barrier, no collective pass);
#   "p2p_all"    — as "p2p" but into every rank's buffer (all-gather semantics; 7x the bytes on 8 GPUs).
GLOBAL_EXCHANGE = "recurrence"


def _fusable_layer_norm(norm: torch.nn.Module, x: torch.Tensor):
    """(weight, bias, eps) if `norm` is a stock affine nn.LayerNorm over the channel dim in fp16, else None."""
    if not FUSE_LAYERNORM or type(norm) is not torch.nn.LayerNorm or not norm.elementwise_affine:
        return None
    if tuple(norm.normalized_shape) != (x.shape[-1],) or norm.weight.dtype != torch.float16 or not norm.weight.is_cuda:
        return None
    return (norm.weight.contiguous(), None if norm.bias is None else norm.bias.contiguous(), float(norm.eps))


def build_merge_plan(module: torch.nn.Module, x: torch.Tensor, tome_info: Dict[str, Any],
                     ln=None) -> Optional[MergePlan]:
    """The body of compute_merge (vidtome/patch.py:14-91).  Returns None when the block is not merged
    (downsample > max_downsample, patch.py:27,86-88).  With `ln = (weight, bias, eps)`, `x` is the RAW
    hidden state and norm1 is applied inside the kernels that read it (K0, KC)."""
    original_h, original_w = tome_info["size"]
    original_tokens = original_h * original_w
    downsample = int(math.ceil(math.sqrt(original_tokens // x.shape[1])))    # patch.py:15-17
    args = tome_info["args"]
    if downsample > args["max_downsample"]:
        return None
    merge._check_metric(x)
    generator = module.generator
    fsize = x.shape[0] // args["batch_size"]                                 # patch.py:23
    tsize = x.shape[1]                                                       # patch.py:24
    align = bool(args["align_batch"])

    table = join_frame(x.contiguous(), fsize)                                # patch.py:37 (view)
    B, N0, C = table.shape
    mu: Optional[torch.Tensor] = None      # position -> row of `table` (None = identity)
    pi: Optional[torch.Tensor] = None      # level-0 position -> current position (None = identity)
    L, unm, curF = N0, 0, fsize
    levels: List[merge.LevelMatch] = []
    randfs: List[int] = []
    ratio = args["local_merge_ratio"]
    while curF > 1:                                                          # patch.py:44
        if ratio <= 0:
            unm += (L - unm) // curF                                         # merge.py:45-46 (no draw)
        else:
            stride = min(args["target_stride"], curF)                        # merge.py:55
            randf = draw_randf(generator, stride, curF)                      # merge.py:56-57
            randfs.append(randf)
            split = VtmSplit.local(L, unm, curF, args["target_stride"], randf)
            m = merge.match_level(table, mu, split, ratio, align, ln)        # patch.py:45-46
            mu, pi = ops.compose_maps(split, m.r, m.keys, m.edge, m.rank, mu, pi, 0, N0)
            levels.append(m)
            unm += m.unm_num                                                 # patch.py:47
            L = mu.shape[1]                                                  # patch.py:50
        curF = (L - unm) // tsize                                            # patch.py:54

    if mu is None:   # nothing merged (single frame or ratio <= 0): identity maps
        mu = torch.arange(N0, dtype=torch.int32, device=x.device)[None]
        pi = mu
    coin = None
    if args["merge_global"]:                                                 # patch.py:59
        g = getattr(module, "global_tokens", None)
        exchanged = False
        local_tokens = None
        if GLOBAL_EXCHANGE in ("allgather", "p2p", "p2p_all"):
            from . import dist as _dist
            if _dist.world() > 1:
                # "KF" = the merge gather together with the exchange (bench.py brackets it with CUDA events): bytes that
                # cross NVLink per rank = (G - 1) * B * L * C * 2 in either mode
                fan = 1 if GLOBAL_EXCHANGE == "p2p" else _dist.world() - 1
                nvl = 2.0 * fan * B * mu.shape[1] * C
                with ops._Timed("KF", 0.0, nvl):
                    if GLOBAL_EXCHANGE in ("p2p", "p2p_all"):
                        # KC stores the merged tokens straight into the consumer's ("p2p": rank k+1) or every rank's
                        # ("p2p_all") symmetric buffer: no separate collective pass
                        local_tokens, g = _dist.exchange_fused(table, mu, ln=ln, everyone=GLOBAL_EXCHANGE == "p2p_all")
                    else:
                        local_tokens = ops.gather_rows(table, mu, ln=ln)
                        g = _dist.exchange_global_tokens(local_tokens)      # the one collective of this path
                exchanged = True
        if local_tokens is None:
            local_tokens = ops.gather_rows(table, mu, ln=ln)                 # merged local tokens [B, L, C]
        if g is not None:                                                    # patch.py:60
            coin = float(draw_scalar(generator, lambda: torch.rand(
                1, generator=generator, device=generator.device)))                            # patch.py:62
            g = g.to(local_tokens).contiguous()                              # patch.py:65,70
            Lg = g.shape[1]
            if coin > args["global_rand"]:
                src_len, tokens, off = L, torch.cat([local_tokens, g], dim=1), 0          # patch.py:63-66
            else:
                src_len, tokens, off = Lg, torch.cat([g, local_tokens], dim=1), Lg        # patch.py:68-71
            gratio = args["global_merge_ratio"]
            if gratio <= 0:
                # merge.py:364-365 returns two values and patch.py:73 unpacks three
                raise ValueError("not enough values to unpack (expected 3, got 2)")
            split = VtmSplit.prefix(L + Lg, src_len)
            m = merge.match_level(tokens, None, split, gratio, align)        # patch.py:73-74
            levels.append(m)
            # tau: position in `tokens` -> position in the merged sequence (the 2s unmerge as a map)
            mu_g, tau = ops.compose_maps(split, m.r, m.keys, m.edge, m.rank, None, None, 0, L + Lg)
            merged_tokens = ops.gather_rows(tokens, mu_g)                    # patch.py:75
            # patch.py:80: global_tokens <- u(merged_tokens), the local partition after unmerging;
            # kept on the device instead of .cpu()
            tau_local = tau[:, off:off + L].contiguous()
            if not exchanged:
                module.global_tokens = ops.unmerge_add(merged_tokens, tau_local, None)
            # compose with the local unmerge: pi_total[p] = tau[off + pi[p]]
            if pi.shape[0] != tau.shape[0]:
                pi = pi.expand(tau.shape[0], -1)
            pi = torch.gather(tau, 1, (pi.long() + off)).to(torch.int32).contiguous()
        else:
            merged_tokens = local_tokens
            module.global_tokens = local_tokens.detach().clone()            # patch.py:82
    else:
        merged_tokens = ops.gather_rows(table, mu, ln=ln)                    # patch.py:50,56 composed
    return MergePlan(fsize=fsize, N0=N0, merged_tokens=merged_tokens, pi=pi, levels=levels, randf=randfs,
                     coin=coin, mu=mu)


def compute_merge(module: torch.nn.Module, x: torch.Tensor, tome_info: Dict[str, Any]) -> Tuple[Callable, ...]:
    """Reference signature and return triple (patch.py:14,91): (merge op, unmerge op, merged tokens).
    The merge op is returned for interface parity only — like the reference's diffusers block
    (patch.py:149-152) nothing here calls it: the merged tokens are already the third element."""
    plan = build_merge_plan(module, x, tome_info)
    if plan is None:
        return merge.do_nothing, merge.do_nothing, x                         # patch.py:86-88
    merged = plan.merged_tokens

    def m(_x: torch.Tensor, **kwarg) -> torch.Tensor:
        return merged
    def u(y: torch.Tensor, **kwarg) -> torch.Tensor:
        return plan.unmerge(y)
    m.plan = u.plan = plan
    return m, u, merged


# attention processors whose result is plain softmax(q k^T * scale) v (what KD computes)
_PLAIN_PROCESSORS = ("AttnProcessor", "AttnProcessor2_0", "XFormersAttnProcessor")


def _plain_attention_module(attn: torch.nn.Module) -> bool:
    """True only if replacing `attn.forward` by KD cannot change the result: `attn` is a stock diffusers-style
    Attention whose projections are bare `torch.nn.Linear` (exact type: LoRA-compatible or PEFT-wrapped layers
    carry extra terms and still expose `.weight`), to_q/to_k/to_v without bias, to_out = [Linear, Dropout(inactive)],
    no instance-level forward override (PnP installs one, utils/pnp_utils.py:99-101), no forward hooks, a default
    attention processor (or none), and none of the Attention options that alter the math (group / spatial norm,
    residual connection, output rescale, added KV projections).  Anything else goes through `self.attn1(...)`
    exactly as the reference does (patch.py:157-162), e.g. after `pipe.load_lora_weights` (generate.py:93-94)."""
    override = vars(attn).get("forward")
    if override is not None and not getattr(override, "_vtm_pnp_forward", False):
        return False                       # someone else's replaced forward (e.g. the reference's own PnP closure)
    need = ("to_q", "to_k", "to_v", "to_out", "heads")
    if not all(hasattr(attn, n) for n in need):
        return False
    to_out = attn.to_out
    if not isinstance(to_out, (torch.nn.ModuleList, torch.nn.Sequential)) or len(to_out) < 1:
        return False
    linears = (attn.to_q, attn.to_k, attn.to_v, to_out[0])
    if any(type(m) is not torch.nn.Linear for m in linears):
        return False
    if any(m.bias is not None for m in linears[:3]):
        return False
    for m in (attn,) + linears + tuple(to_out[1:]):
        if m._forward_hooks or m._forward_pre_hooks or getattr(m, "_forward_hooks_with_kwargs", None):
            return False
    for extra in to_out[1:]:
        if not isinstance(extra, torch.nn.Dropout) or (extra.p > 0 and extra.training):
            return False
    proc = getattr(attn, "processor", None)
    if proc is not None and type(proc).__name__ not in _PLAIN_PROCESSORS:
        return False
    if getattr(attn, "group_norm", None) is not None or getattr(attn, "spatial_norm", None) is not None:
        return False
    if getattr(attn, "norm_cross", None) or getattr(attn, "residual_connection", False):
        return False
    if getattr(attn, "rescale_output_factor", 1.0) != 1.0:
        return False
    if getattr(attn, "added_kv_proj_dim", None) is not None or getattr(attn, "add_k_proj", None) is not None:
        return False
    C = attn.to_q.weight.shape[1]
    if attn.to_q.weight.shape[0] != C or attn.to_k.weight.shape != attn.to_q.weight.shape:
        return False                       # inner dim != query dim or cross-attention shaped K/V
    if attn.to_v.weight.shape != attn.to_q.weight.shape or to_out[0].weight.shape != (C, C):
        return False
    head_dim = C // int(attn.heads)
    return head_dim * int(attn.heads) == C and head_dim % 8 == 0 and head_dim <= 128   # KD's supported range


def make_tome_block(block_class: Type[torch.nn.Module]) -> Type[torch.nn.Module]:
    """CompVis-LDM variant (patch.py:94-116).  The reference's version unpacks six values from
    compute_merge's three (patch.py:105 vs :91) and cannot run; we refuse at patch time instead."""
    raise NotImplementedError(
        "vidtome_b200: the LDM (non-diffusers) ToMeBlock of the reference is not runnable upstream "
        "(patch.py:105 unpacks 6 values from a 3-tuple); only diffusers models are supported")


def make_diffusers_tome_block(block_class: Type[torch.nn.Module]) -> Type[torch.nn.Module]:
    """Patched class for a diffusers BasicTransformerBlock (patch.py:119-203)."""

    class ToMeBlock(block_class):
        # the original class, restored by remove_patch
        _parent = block_class

        def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None,
                    encoder_attention_mask=None, timestep=None, cross_attention_kwargs=None,
                    class_labels=None) -> torch.Tensor:
            plan = None
            if self.use_ada_layer_norm:                                       # patch.py:139-146
                norm_hidden_states = self.norm1(hidden_states, timestep)
            elif self.use_ada_layer_norm_zero:
                norm_hidden_states, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.norm1(
                    hidden_states, timestep, class_labels, hidden_dtype=hidden_states.dtype)
            else:
                ln = _fusable_layer_norm(self.norm1, hidden_states) if hidden_states.is_cuda else None
                if ln is not None:
                    # norm1 fused into the merge kernels: only the kept rows are ever normalised
                    plan = build_merge_plan(self, hidden_states, self._tome_info, ln=ln)
                norm_hidden_states = None if plan is not None else self.norm1(hidden_states)

            if plan is None:
                plan = build_merge_plan(self, norm_hidden_states, self._tome_info)  # patch.py:149-150
            if plan is not None:
                norm_hidden_states = plan.merged_tokens

            # 1. Self-Attention (patch.py:154-162)
            cross_attention_kwargs = cross_attention_kwargs if cross_attention_kwargs is not None else {}
            only_cross = getattr(self, "only_cross_attention", False)
            from . import attention as _attention
            if (plan is not None and _attention.ENABLED and not only_cross and attention_mask is None
                    and not cross_attention_kwargs and _plain_attention_module(self.attn1)):
                # KD; with PnP control registered (pnp.py) and the timestep inside the injection schedule, the attention
                # map of the source sample is applied to every sample's values (utils/pnp_utils.py:57-68,87-91)
                shared = False
                if getattr(self.attn1, "_vtm_pnp", 0):
                    from . import pnp as _pnp
                    shared = _pnp.injection_active(self.attn1)
                    if shared and norm_hidden_states.shape[0] != self.attn1._vtm_pnp:
                        raise RuntimeError("PnP injection needs one merged token set per input (batch_size == num_inputs)")
                attn_output = _attention.self_attention(self.attn1, norm_hidden_states, shared_qk=shared)  # KD
            else:
                attn_output = self.attn1(
                    norm_hidden_states,
                    encoder_hidden_states=encoder_hidden_states if only_cross else None,
                    attention_mask=attention_mask, **cross_attention_kwargs)
            if self.use_ada_layer_norm_zero:
                attn_output = gate_msa.unsqueeze(1) * attn_output            # patch.py:164-165

            # Unmerge + residual (patch.py:168-169), fused into one pass
            if plan is not None:
                hidden_states = plan.unmerge_add(attn_output, hidden_states)
            else:
                hidden_states = attn_output + hidden_states

            attn2 = getattr(self, "attn2", None)
            if (attn2 is not None and FUSE_CROSS_ATTENTION and hidden_states.is_cuda and not self.use_ada_layer_norm
                    and encoder_attention_mask is None and not cross_attention_kwargs
                    and _attention.cross_attention_eligible(attn2, encoder_hidden_states)
                    and encoder_hidden_states.shape[0] == hidden_states.shape[0]):
                ln2 = _fusable_layer_norm(self.norm2, hidden_states)
                if ln2 is not None:
                    # norm2 -> q / kv projections (head-major) -> flash attention over the context -> out projection
                    # with bias and residual fused (patch.py:171-185), all in the CUDA library
                    shape = hidden_states.shape
                    n2 = ops.layer_norm(hidden_states.contiguous().view(-1, shape[-1]), ln2).view(shape)
                    hidden_states = _attention.cross_attention_residual(attn2, n2, encoder_hidden_states, hidden_states)
                    attn2 = None
            if attn2 is not None:                                            # patch.py:171-185
                norm_hidden_states = (self.norm2(hidden_states, timestep) if self.use_ada_layer_norm
                                      else self.norm2(hidden_states))
                attn_output = self.attn2(norm_hidden_states, encoder_hidden_states=encoder_hidden_states,
                                         attention_mask=encoder_attention_mask, **cross_attention_kwargs)
                hidden_states = attn_output + hidden_states

            # 3. Feed-forward (patch.py:187-199).  A block without `ff` (the hot-path skeleton used by
            # bench.py) stops after the self-attention section.
            ff = getattr(self, "ff", None)
            if ff is not None and FUSE_FEED_FORWARD and hidden_states.is_cuda and not self.use_ada_layer_norm_zero:
                from . import feedforward as _ff
                parts = _ff.geglu_parts(ff)
                ln3 = _fusable_layer_norm(self.norm3, hidden_states) if parts is not None else None
                if ln3 is not None:
                    # norm3 -> GEGLU projection (gate fused) -> output projection (+ bias + residual), tcgen05
                    return _ff.feed_forward_residual(ff, parts, ln3, hidden_states)
            if ff is not None:
                norm_hidden_states = self.norm3(hidden_states)
                if self.use_ada_layer_norm_zero:
                    norm_hidden_states = norm_hidden_states * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
                ff_output = self.ff(norm_hidden_states)
                if self.use_ada_layer_norm_zero:
                    ff_output = gate_mlp.unsqueeze(1) * ff_output
                hidden_states = ff_output + hidden_states
            return hidden_states

    return ToMeBlock


def hook_tome_model(model: torch.nn.Module):
    """Forward pre-hook recording the latent size (patch.py:206-212).  The UNet must be called with the
    latent as positional argument 0, as generate.py:275 does."""
    def hook(module, args):
        module._tome_info["size"] = (args[0].shape[2], args[0].shape[3])
        return None

    model._tome_info["hooks"].append(model.register_forward_pre_hook(hook))


def hook_tome_module(module: torch.nn.Module):
    """Forward pre-hook that gives the block its own `module.generator`, forked from the default RNG of the
    input's device the first time the block runs (and again if the block moves to another device), so that all
    blocks draw the same random target frames within one pass (patch.py:215-231)."""
    def hook(module, args):
        device = args[0].device
        current = getattr(module, "generator", None)
        if current is None:
            module.generator = init_generator(device)
        elif current.device != device:
            module.generator = init_generator(device, fallback=current)
        return None

    module._tome_info["hooks"].append(module.register_forward_pre_hook(hook))


def apply_patch(
        model: torch.nn.Module,
        local_merge_ratio: float = 0.9,
        merge_global: bool = False,
        global_merge_ratio=0.8,
        max_downsample: int = 2,
        seed: int = 123,
        batch_size: int = 2,
        include_control: bool = False,
        align_batch: bool = False,
        target_stride: int = 4,
        global_rand=0.5):
    """Patch a Stable-Diffusion model with VidToMe (signature, defaults and semantics of
    vidtome/patch.py:234-334).

     - model: a diffusers pipeline (uses `.unet`, and `.controlnet` when it is a
       StableDiffusionControlNetPipeline and include_control) or a bare diffusers UNet; detected by class
       NAME (DiffusionPipeline / ModelMixin), blocks by the name BasicTransformerBlock.
     - local_merge_ratio: fraction of src tokens merged inside a frame chunk (0.9 -> 1.3/4.0 kept for 4 frames).
     - merge_global / global_merge_ratio / global_rand: inter-chunk (global) token merging.
     - max_downsample: merge only in layers with at most this downsampling (1, 2, 4 or 8).
     - seed: stored, unused (as in the reference).   - batch_size: number of video chunks per pass (2 = CFG, 3 = PnP).
     - align_batch: share one matching across the batch (PnP).   - target_stride: one target frame per this many frames.
    """
    # start from an unpatched model: a second apply_patch must not stack hooks or subclasses
    remove_patch(model)

    is_diffusers = isinstance_str(model, "DiffusionPipeline") or isinstance_str(model, "ModelMixin")

    if not is_diffusers:
        if not hasattr(model, "model") or not hasattr(model.model, "diffusion_model"):
            raise RuntimeError("Provided model was not a Stable Diffusion / Latent Diffusion model, as expected.")
        diffusion_model = model.model.diffusion_model
    else:
        # a pipeline carries its UNet as `.unet`; a bare UNet is used as is
        diffusion_model = model.unet if hasattr(model, "unet") else model

    if isinstance_str(model, "StableDiffusionControlNetPipeline") and include_control:
        diffusion_models = [diffusion_model, model.controlnet]
    else:
        diffusion_models = [diffusion_model]

    from . import _lib
    _lib.load()   # fail at patch time, loudly, if the CUDA library is missing

    for diffusion_model in diffusion_models:
        diffusion_model._tome_info = {
            "size": None,
            "hooks": [],
            "args": {
                "max_downsample": max_downsample,
                "generator": None,
                "seed": seed,
                "batch_size": batch_size,
                "align_batch": align_batch,
                "merge_global": merge_global,
                "global_merge_ratio": global_merge_ratio,
                "local_merge_ratio": local_merge_ratio,
                "global_rand": global_rand,
                "target_stride": target_stride
            }
        }
        hook_tome_model(diffusion_model)

        for name, module in diffusion_model.named_modules():
            if isinstance_str(module, "BasicTransformerBlock"):
                make_tome_block_fn = make_diffusers_tome_block if is_diffusers else make_tome_block
                module.__class__ = make_tome_block_fn(module.__class__)
                module._tome_info = diffusion_model._tome_info
                hook_tome_module(module)

                # blocks of diffusers releases that predate the ada-norm flags: read as "plain LayerNorm"
                if not hasattr(module, "use_ada_layer_norm_zero") and is_diffusers:
                    module.use_ada_layer_norm = False
                    module.use_ada_layer_norm_zero = False

    return model


def _roots_of(model: torch.nn.Module, controlnet_owner: torch.nn.Module) -> List[torch.nn.Module]:
    """[UNet (or the model itself)] + [controlnet_owner.controlnet] when that attribute exists.

    The three lifecycle calls of the reference differ in WHERE they look for `.controlnet`:
      * remove_patch rebinds `model` to the UNet first (patch.py:341-344), so it looks on the UNet — a ControlNet
        patched through `include_control=True` is therefore not reached by remove_patch;
      * update_patch / collect_from_patch keep the object they were given (patch.py:361-364, :376-378), so on a
        pipeline they DO reach `pipe.controlnet`.
    Both behaviours are reproduced; the caller passes the object the reference would test."""
    root = model.unet if hasattr(model, "unet") else model
    roots = [root]
    if hasattr(controlnet_owner, "controlnet"):
        roots.append(controlnet_owner.controlnet)
    return roots


def remove_patch(model: torch.nn.Module):
    """Undo apply_patch if the model is patched: drop the hooks and restore the block classes
    (patch.py:337-355).  Returns what the reference returns: the last module it walked."""
    unet = model.unet if hasattr(model, "unet") else model
    roots = _roots_of(model, unet)                       # `.controlnet` is looked up on the UNet here
    for root in roots:
        for _, module in root.named_modules():
            info = getattr(module, "_tome_info", None)
            if info is not None:
                for handle in info["hooks"]:
                    handle.remove()
                info["hooks"].clear()
            if module.__class__.__name__ == "ToMeBlock":
                module.__class__ = module._parent
    return roots[-1]


def update_patch(model: torch.nn.Module, **kwargs):
    """Set attributes on every module that carries `_tome_info` (the UNet, each patched block and — on a
    pipeline that has one — the ControlNet and its blocks), e.g. `update_patch(pipe, global_tokens=None)`
    after each denoising step (patch.py:358-370, generate.py:233-236).  Like the reference, the return
    value is the last root walked (its loop variable shadows `model`), not the object passed in."""
    roots = _roots_of(model, model)                      # `.controlnet` is looked up on the object passed in
    for root in roots:
        for _, module in root.named_modules():
            if hasattr(module, "_tome_info"):
                for key, value in kwargs.items():
                    setattr(module, key, value)
    return roots[-1]


def collect_from_patch(model: torch.nn.Module, attr="tome"):
    """{qualified module name: getattr(module, attr)} for every module that has `attr` (patch.py:373-387);
    a pipeline's ControlNet is included, later roots overwrite equal names exactly as the reference's dict does."""
    found = dict()
    for root in _roots_of(model, model):
        for name, module in root.named_modules():
            if hasattr(module, attr):
                found[name] = getattr(module, attr)
    return found
