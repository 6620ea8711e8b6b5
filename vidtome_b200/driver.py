"""Minimal denoising driver around a patched UNet — the loop the metric "denoising steps/sec" counts.

Restates, on synthetic latents, the parts of the reference's Generator that sit around the hot path:
  get_chunks      generate.py:172-203   (frame chunking; order seq / rand / mix for global merging)
  ddim_sample     generate.py:205-224   (per step: every chunk's pred_noise, then the DDIM update)
  pred_noise      generate.py:238-279   (CFG batch cat([x, x]), UNet forward, guidance combine)
  pred_next_x     generate.py:281-311   (closed-form DDIM step)
  post_iter       generate.py:233-236   (reset global tokens after each step)
Text conditioning, VAE, ControlNet, PnP injection and file IO are outside the hot path and not restated.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import patch


def ddim_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                        beta_end: float = 0.012) -> torch.Tensor:
    """Stable Diffusion's "scaled_linear" schedule (what DDIMScheduler.from_pretrained gives for SD1.5/2.1)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(n_timesteps: int, num_train_timesteps: int = 1000) -> List[int]:
    """DDIMScheduler.set_timesteps with 'leading' spacing and steps_offset=1 (SD's scheduler config)."""
    ratio = num_train_timesteps // n_timesteps
    return [int(t) + 1 for t in (np.arange(0, n_timesteps) * ratio).round()[::-1]]


class ChunkedDenoiser:
    """Chunked classifier-free-guidance DDIM sampling on a (patched) UNet, generate.py:205-311."""

    def __init__(self, unet: torch.nn.Module, n_timesteps: int = 50, chunk_size: int = 16,
                 guidance_scale: float = 7.5, merge_global: bool = False, chunk_ord: str = "seq",
                 perm_div: float = 3.0, randomize_chunks: bool = False, cond: Optional[torch.Tensor] = None,
                 cuda_graph: bool = False, graph_warmup: int = 2, shard: bool = False):
        """`cuda_graph=True`: after `graph_warmup` eager steps, the noise prediction of a whole step (every chunk:
        CFG batch, UNet forward with the merge path, guidance combine) is captured once into a CUDA graph and
        replayed; per step the host then issues one graph launch plus the DDIM update instead of ~250 kernel
        launches.  The frame draws of the local matcher stay on the device (utils.draw_randf) and come from the
        blocks' generators, which are registered with the graph, so replays produce the same sequence of draws —
        and bit-identical latents — as eager stepping.  Needs static shapes and chunk order: local merging only
        (`merge_global=False`), `randomize_chunks=False`, frame counts divisible by the target stride."""
        self.unet = unet
        # shard=True: this process is one rank of a torch.distributed job and takes chunks i % world == rank of every
        # step (dist.shard_chunks); with merge_global and a collective exchange mode the ranks are checked for
        # lock-step once per step (dist.check_step_lockstep)
        self.shard = bool(shard)
        self.cuda_graph = bool(cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self._graph = None          # (CUDAGraph, static x, static timestep, static noise, shape)
        self._eager_steps = 0
        if self.cuda_graph and (merge_global or randomize_chunks):
            raise ValueError("cuda_graph=True needs merge_global=False and randomize_chunks=False (static work)")
        self.chunk_size = chunk_size
        self.guidance_scale = guidance_scale
        self.merge_global = merge_global
        # generate.py:86-89: "mix-#" means a partial permutation of len/# chunks ("mix" alone: # = 3)
        if "mix" in chunk_ord:
            perm_div = float(chunk_ord.split("-")[-1]) if "-" in chunk_ord else (perm_div if perm_div else 3.0)
            chunk_ord = "mix"
        self.chunk_ord, self.perm_div = chunk_ord, perm_div
        self.randomize_chunks = randomize_chunks
        self.cond = cond                      # [2, 77, D] (uncond, cond) text embeddings, may be None
        self.alphas_cumprod = ddim_alphas_cumprod()
        self.final_alpha_cumprod = self.alphas_cumprod[0]     # set_alpha_to_one=False in SD's config
        self.timesteps = ddim_timesteps(n_timesteps)

    # generate.py:172-203
    def get_chunks(self, flen: int) -> List[torch.Tensor]:
        x_index = torch.arange(flen)
        # The first chunk has a random length in the reference; benchmarks pin it to chunk_size
        rand_first = (np.random.randint(0, self.chunk_size) + 1) if self.randomize_chunks else self.chunk_size
        rest = x_index[rand_first:].split(self.chunk_size, dim=0)
        chunks = [x_index[:rand_first]] + (list(rest) if len(rest) and len(rest[0]) > 0 else [])
        if self.randomize_chunks and np.random.rand() > 0.5:
            chunks = chunks[::-1]
        if not self.merge_global:
            return chunks
        if self.chunk_ord == "rand":
            order = torch.randperm(len(chunks)).tolist()
        elif self.chunk_ord == "mix":
            randord = torch.randperm(len(chunks)).tolist()
            rand_len = int(len(randord) / self.perm_div)
            seqord = sorted(randord[rand_len:])
            if rand_len > 0:
                randord = randord[:rand_len]
                if abs(seqord[-1] - randord[-1]) < abs(seqord[0] - randord[-1]):
                    seqord = seqord[::-1]
                order = randord + seqord
            else:
                order = seqord
        else:
            order = list(range(len(chunks)))
        return [chunks[i] for i in order]

    # generate.py:238-279
    @torch.no_grad()
    def pred_noise(self, x: torch.Tensor, t: int) -> torch.Tensor:
        flen = len(x)
        text = None if self.cond is None else self.cond.repeat_interleave(flen, dim=0)
        latent_model_input = torch.cat([x, x])                                # classifier-free guidance
        eps = self.unet(latent_model_input, t, encoder_hidden_states=text).sample
        noise_pred_uncond, noise_pred_cond = eps.chunk(2)
        return noise_pred_uncond + self.guidance_scale * (noise_pred_cond - noise_pred_uncond)

    # generate.py:281-311 (inversion=False branch)
    def pred_next_x(self, x: torch.Tensor, eps: torch.Tensor, i: int) -> torch.Tensor:
        t = self.timesteps[i]
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = (self.alphas_cumprod[self.timesteps[i + 1]] if i < len(self.timesteps) - 1
                             else self.final_alpha_cumprod)
        mu, sigma = float(alpha_prod_t ** 0.5), float((1 - alpha_prod_t) ** 0.5)
        mu_prev, sigma_prev = float(alpha_prod_t_prev ** 0.5), float((1 - alpha_prod_t_prev) ** 0.5)
        pred_x0 = (x - sigma * eps) / mu
        return mu_prev * pred_x0 + sigma_prev * eps

    def _step_chunks(self, flen: int) -> List[torch.Tensor]:
        chunks = self.get_chunks(flen)
        if not self.shard:
            return chunks
        from . import dist as _dist
        chunks = _dist.shard_chunks(chunks)
        if self.merge_global and patch.GLOBAL_EXCHANGE in ("allgather", "p2p", "p2p_all") and _dist.world() > 1:
            uniform = _dist.check_step_lockstep([len(c) for c in chunks])
            if not uniform and patch.GLOBAL_EXCHANGE in ("p2p", "p2p_all"):
                raise RuntimeError("GLOBAL_EXCHANGE='p2p' needs equally long chunks on every rank (fixed chunking); "
                                   "use 'allgather' for ragged chunks")
        return chunks

    def _all_noises(self, x: torch.Tensor, t) -> torch.Tensor:
        noises = torch.zeros_like(x)
        for chunk in self._step_chunks(len(x)):
            # chunks are contiguous frame ranges (possibly visited in another order): slice instead of indexing with
            # a host tensor, which would cost a synchronous host->device copy per chunk per step
            lo, hi = int(chunk[0]), int(chunk[-1]) + 1
            noises[lo:hi] = self.pred_noise(x[lo:hi], t)
        return noises

    def _block_generators(self) -> List[torch.Generator]:
        seen, gens = set(), []
        for m in self.unet.modules():
            g = getattr(m, "generator", None)
            if isinstance(g, torch.Generator) and g.device.type == "cuda" and id(g) not in seen:
                seen.add(id(g))
                gens.append(g)
        return gens

    def _capture(self, x: torch.Tensor, t: int):
        """Capture `_all_noises` for inputs shaped like `x` (called after the eager warm-up steps, so that every
        lazily created piece of state — generators forked by the hook, cached packed weights, the library handle,
        allocator pools — already exists)."""
        gx = x.clone()
        gt = torch.tensor(int(t), device=x.device, dtype=torch.long)      # the UNet's timestep input, set per replay
        graph = torch.cuda.CUDAGraph()
        for g in self._block_generators():
            graph.register_generator_state(g)
        torch.cuda.synchronize(x.device)
        with torch.cuda.graph(graph):
            gn = self._all_noises(gx, gt)
        self._graph = (graph, gx, gt, gn, tuple(x.shape))

    # one iteration of generate.py:211-224
    @torch.no_grad()
    def step(self, x: torch.Tensor, i: int) -> torch.Tensor:
        t = self.timesteps[i]
        if self.cuda_graph and x.is_cuda and self._eager_steps >= self.graph_warmup:
            if self._graph is None or self._graph[4] != tuple(x.shape):
                self._capture(x, t)
            graph, gx, gt, gn, _ = self._graph
            gx.copy_(x, non_blocking=True)
            gt.fill_(int(t))
            graph.replay()
            noises = gn
        else:
            self._eager_steps += 1
            noises = self._all_noises(x, t)
        x = self.pred_next_x(x, noises, i)
        if self.merge_global:
            patch.update_patch(self.unet, global_tokens=None)                # generate.py:233-236
        return x

    @torch.no_grad()
    def sample(self, x: torch.Tensor) -> torch.Tensor:
        for i in range(len(self.timesteps)):
            x = self.step(x, i)
        return x
