"""Generate tests/golden/*.npz by running the REFERENCE itself (lixirui142/VidToMe, /root/reference).

Run in the build container only (`python tests/make_golden.py`); the fixtures are committed because
/root/reference does not exist on the GPU box.  Nothing in the product or the test suite imports the
reference at test time.

What is recorded, per case: the seeded inputs, every random draw the reference made (torch.randint at
merge.py:56-57, torch.rand at patch.py:62 — captured by wrapping the torch functions while the reference
runs), and the reference's outputs.  The reference's `argsort(descending=True)` (merge.py:98,113,402,417)
is forced to stable=True while generating, because the default CPU argsort is not stable and its tie
order is arbitrary, whereas torch's CUDA radix sort — the reference's real deployment — is stable; the
fraction of positions where the unpatched CPU argsort would have differed is stored as `unstable_frac`
for the record.

Input families:
  exact   fp16 rows with 64 entries of +-2^-k: every normalised value and every dot product is exactly
          representable, so scores (with their heavy ties) are bit-identical on any device in any
          accumulation order -> indices must match bit for bit everywhere;
  fp32    iid N(0,1) in fp32: ties are absent;
  video   base + 0.1*noise in fp16 (the survey's "video-like" family): compared tie-tolerantly.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "golden")
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import vidtome as ref  # noqa: E402  (the reference package)
from vidtome import merge as ref_merge, patch as ref_patch  # noqa: E402


# ------------------------------------------------------------------------------------------------
class Recorder:
    """Wraps torch.randint / torch.rand / Tensor.argsort while the reference runs."""

    def __init__(self, stable: bool = True):
        self.randint, self.rand = [], []
        self.stable = stable
        self.unstable_positions = 0
        self.sorted_positions = 0

    def __enter__(self):
        self._ri, self._r, self._as = torch.randint, torch.rand, torch.Tensor.argsort
        rec = self

        def randint(*a, **k):
            out = rec._ri(*a, **k)
            rec.randint.append(int(out.item()))
            return out

        def rand(*a, **k):
            out = rec._r(*a, **k)
            rec.rand.append(float(out.item()))
            return out

        def argsort(t, *a, **k):
            plain = rec._as(t, *a, **k)
            k2 = dict(k)
            k2["stable"] = True
            st = rec._as(t, *a, **k2)
            rec.unstable_positions += int((plain != st).sum())
            rec.sorted_positions += plain.numel()
            return st if rec.stable else plain

        torch.randint, torch.rand, torch.Tensor.argsort = randint, rand, argsort
        return self

    def __exit__(self, *exc):
        torch.randint, torch.rand, torch.Tensor.argsort = self._ri, self._r, self._as

    @property
    def unstable_frac(self):
        return self.unstable_positions / max(1, self.sorted_positions)


def exact_tokens(rng, shape, nnz=64, val=0.25):
    B, N, C = shape
    x = np.zeros(shape, dtype=np.float16)
    for b in range(B):
        for i in range(N):
            cols = rng.choice(C, size=nnz, replace=False)
            x[b, i, cols] = rng.choice([-val, val], size=nnz)
    return x


def exact_video_tokens(rng, B, F, T, C, nnz=64, val=0.25, flips=6):
    """Exact-arithmetic tokens with video structure: frame f's token t is the base token t with a few
    sign flips, so the best match of a src token is (usually) the same token in the dst frame, and many
    scores tie exactly."""
    x = np.zeros((B, F, T, C), dtype=np.float16)
    for b in range(B):
        base = exact_tokens(rng, (1, T, C), nnz, val)[0]
        for f in range(F):
            fr = base.copy()
            for t in range(T):
                nz = np.nonzero(fr[t])[0]
                sel = rng.choice(nz, size=rng.integers(0, flips + 1), replace=False)
                fr[t, sel] = -fr[t, sel]
            x[b, f] = fr
    return x.reshape(B, F * T, C)


def video_tokens(rng, B, F, T, C, dtype):
    base = rng.standard_normal((B, 1, T, C))
    x = base + 0.1 * rng.standard_normal((B, F, T, C))
    return x.reshape(B, F * T, C).astype(dtype)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def t2n(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
def case_randframe(name, x, F, ratio, unm_pre, align, seed, target_stride=4):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xt = torch.from_numpy(x)
    with Recorder() as rec:
        m, u, ret = ref_merge.bipartite_soft_matching_randframe(xt, F, ratio, unm_pre, g, target_stride, align)
        merged = m(xt)
        unmerged = u(merged)
    cl = {k: v for k, v in zip(m.__code__.co_freevars, [c.cell_contents for c in m.__closure__])}
    save(name, x=x, F=F, ratio=ratio, unm_pre=unm_pre, align=int(align), target_stride=target_stride,
         randf=np.array(rec.randint), unm_num=ret["unm_num"],
         unm_idx=t2n(cl["unm_idx"])[..., 0], src_idx=t2n(cl["src_idx"])[..., 0], dst_idx=t2n(cl["dst_idx"])[..., 0],
         merged=t2n(merged), unmerged=t2n(unmerged), unstable_frac=rec.unstable_frac)


def case_2s(name, x, src_len, ratio, align, chunk):
    xt = torch.from_numpy(x)
    with Recorder() as rec:
        m, u, ret = ref_merge.bipartite_soft_matching_2s(xt, src_len, ratio, align, unmerge_chunk=chunk)
        merged = m(xt)
        unmerged = u(merged)
    cl = {k: v for k, v in zip(m.__code__.co_freevars, [c.cell_contents for c in m.__closure__])}
    save(name, x=x, src_len=src_len, ratio=ratio, align=int(align), chunk=chunk, unm_num=ret["unm_num"],
         unm_idx=t2n(cl["unm_idx"])[..., 0], src_idx=t2n(cl["src_idx"])[..., 0], dst_idx=t2n(cl["dst_idx"])[..., 0],
         merged=t2n(merged), unmerged=t2n(unmerged), unstable_frac=rec.unstable_frac)


def case_compute_merge(name, chunks, size, batch_size, seed, **args):
    """Runs reference compute_merge (patch.py:14-91) over a sequence of chunks on ONE module, so the
    global-token recurrence (patch.py:59-82) is exercised.  `chunks`: list of [(B F), T, C] arrays."""
    module = SimpleNamespace(generator=torch.Generator(device="cpu").manual_seed(seed), global_tokens=None)
    info = {"size": size, "hooks": [], "args": dict(
        max_downsample=2, generator=None, seed=123, batch_size=batch_size, align_batch=False, merge_global=False,
        global_merge_ratio=0.8, local_merge_ratio=0.9, global_rand=0.5, target_stride=4)}
    info["args"].update(args)
    out = {"n_chunks": len(chunks), "size": np.array(size), "batch_size": batch_size}
    out.update({f"arg_{k}": v for k, v in info["args"].items() if k not in ("generator",)})
    unstable = []
    for i, x in enumerate(chunks):
        xt = torch.from_numpy(x)
        with Recorder() as rec:
            m, u, merged = ref_patch.compute_merge(module, xt, info)
            # unmerge the merged tokens themselves: a pure row gather of the input
            back = u(merged)
        unstable.append(rec.unstable_frac)
        out[f"x{i}"] = x
        out[f"randint{i}"] = np.array(rec.randint, dtype=np.int64)
        out[f"rand{i}"] = np.array(rec.rand, dtype=np.float64)
        out[f"merged{i}"] = t2n(merged)
        out[f"back{i}"] = t2n(back)
        if module.global_tokens is not None:
            out[f"global{i}"] = t2n(module.global_tokens)
    out["unstable_frac"] = np.array(unstable)
    save(name, **out)


def case_block(name, hidden, size, batch_size, seed, dim, heads, **args):
    """Reference-patched BasicTransformerBlock (full block: self-attn section + cross-attn + FF) in fp16
    on the CPU: vidtome.apply_patch on a one-block skeleton UNet, forward through the UNet pre-hook."""
    from vidtome_b200.skeleton import BasicTransformerBlock, ModelMixin

    class OneBlock(ModelMixin):
        def __init__(self):
            super().__init__()
            self.block = BasicTransformerBlock(dim, heads, cross_attention_dim=32, hot_path_only=False)

        def forward(self, latent, hidden, ctx):
            return self.block(hidden, encoder_hidden_states=ctx)

    torch.manual_seed(seed)
    net = OneBlock().half().eval()
    ref.apply_patch(net, batch_size=batch_size, **args)
    torch.manual_seed(seed + 1)   # default RNG state that the block forks into module.generator
    h = torch.from_numpy(hidden)
    ctx = torch.randn(h.shape[0], 7, 32).half()
    latent = torch.zeros(h.shape[0], 4, size[0], size[1])
    with Recorder() as rec, torch.no_grad():
        out = net(latent, h, ctx)
        # the self-attention section alone (what the hot path replaces): recompute with the same draws
    sd = {"sd_" + k: t2n(v) for k, v in net.state_dict().items()}
    save(name, hidden=hidden, ctx=t2n(ctx), size=np.array(size), batch_size=batch_size, dim=dim, heads=heads,
         randint=np.array(rec.randint, dtype=np.int64), out=t2n(out), unstable_frac=rec.unstable_frac,
         **{f"arg_{k}": v for k, v in args.items()}, **sd)
    ref.remove_patch(net)


def main():
    rng = np.random.default_rng(123)
    torch.manual_seed(123)
    # ---- L0: local matcher
    case_randframe("randframe_exact_f4", exact_video_tokens(rng, 2, 4, 48, 128), 4, 0.9, 0, False, 1)
    case_randframe("randframe_exact_f4_align", exact_video_tokens(rng, 3, 4, 48, 128), 4, 0.9, 0, True, 2)
    xe = np.concatenate([exact_tokens(rng, (2, 29, 128)), exact_video_tokens(rng, 2, 5, 40, 128)], axis=1)
    case_randframe("randframe_exact_f5_unm29", xe, 5, 0.75, 29, False, 3)          # ragged F, carried tokens
    case_randframe("randframe_exact_f2", exact_video_tokens(rng, 2, 2, 64, 128), 2, 0.5, 0, False, 4)
    case_randframe("randframe_fp32_f4", rng.standard_normal((2, 4 * 40, 64)).astype(np.float32), 4, 0.9, 0, False, 5)
    case_randframe("randframe_fp32_f4_align", rng.standard_normal((2, 4 * 40, 64)).astype(np.float32), 4, 0.9, 0, True, 6)
    case_randframe("randframe_video_f16", video_tokens(rng, 2, 16, 32, 64, np.float16), 16, 0.9, 0, False, 7)
    # ---- L0: global matcher
    xg = exact_tokens(rng, (2, 150, 128))
    case_2s("2s_exact_chunk0", xg, 70, 0.8, False, 0)
    case_2s("2s_exact_chunk1", xg, 70, 0.8, False, 1)
    case_2s("2s_exact_align", exact_tokens(rng, (3, 120, 128)), 60, 0.8, True, 0)
    case_2s("2s_fp32", rng.standard_normal((2, 130, 64)).astype(np.float32), 65, 0.8, False, 1)
    # ---- L1: compute_merge (levels 16 -> 4 -> 1; 8 -> 2 -> 1; global recurrence over 3 chunks)
    B = 2
    def chunk(F, T, C):
        return exact_video_tokens(rng, B, F, T, C).reshape(B * F, T, C)
    case_compute_merge("compute_merge_exact_f16", [chunk(16, 16, 128)], (4, 4), B, 11)
    case_compute_merge("compute_merge_exact_f8_align", [chunk(8, 16, 128)], (4, 4), B, 12, align_batch=True)
    case_compute_merge("compute_merge_exact_f6", [chunk(6, 16, 128)], (4, 4), B, 13)          # F % stride != 0
    case_compute_merge("compute_merge_exact_global", [chunk(4, 16, 128), chunk(4, 16, 128), chunk(2, 16, 128),
                                                      chunk(4, 16, 128)],
                       (4, 4), B, 14, merge_global=True)
    case_compute_merge("compute_merge_exact_global_align", [chunk(4, 16, 128), chunk(4, 16, 128), chunk(4, 16, 128)],
                       (4, 4), B, 15, merge_global=True, align_batch=True)
    case_compute_merge("compute_merge_skip_ds4", [chunk(4, 4, 128)], (8, 8), B, 16)          # downsample 4 > 2: no merge
    # ---- block level (reference ToMeBlock, fp16 CPU)
    # (later additions are generated AFTER the block fixtures below, from their own generator, so that the fixtures
    #  above stay byte-identical when this script is re-run)
    hid = video_tokens(rng, 2, 4, 64, 128, np.float16).reshape(8, 64, 128)
    case_block("block_ratio1", hid, (8, 8), 2, 21, 128, 2, local_merge_ratio=1.0)
    case_block("block_ratio09", hid, (8, 8), 2, 22, 128, 2, local_merge_ratio=0.9)
    # ---- additions: the default PnP configuration (3 samples, align_batch, configs/default.yaml:25,56) with global
    #      merging over two chunks, and a low local ratio over three levels' worth of frames
    rng2 = np.random.default_rng(321)
    def chunk3(F, T, C):
        return exact_video_tokens(rng2, 3, F, T, C).reshape(3 * F, T, C)
    case_compute_merge("compute_merge_exact_pnp_b3", [chunk3(4, 16, 128), chunk3(4, 16, 128)], (4, 4), 3, 31,
                       merge_global=True, align_batch=True)
    case_compute_merge("compute_merge_exact_ratio05_f8",
                       [exact_video_tokens(rng2, 2, 8, 16, 128).reshape(16, 16, 128)], (4, 4), 2, 32,
                       local_merge_ratio=0.5)
    # ---- oracle-only additions (replayed by tests/test_oracle_golden.py): odd frame count, stride 2, every src token
    #      merged (ratio 1.0), and both branches of the global coin forced (global_rand 0.0 / 1.0)
    def chunk2(F, T, C):
        return exact_video_tokens(rng2, 2, F, T, C).reshape(2 * F, T, C)
    case_compute_merge("compute_merge_exact_f5", [chunk2(5, 16, 128)], (4, 4), 2, 33)
    case_compute_merge("compute_merge_exact_stride2_f8", [chunk2(8, 16, 128)], (4, 4), 2, 34, target_stride=2)
    case_compute_merge("compute_merge_exact_ratio1_f4", [chunk2(4, 16, 128)], (4, 4), 2, 35, local_merge_ratio=1.0)
    case_compute_merge("compute_merge_exact_global_rand0", [chunk2(4, 16, 128), chunk2(4, 16, 128)], (4, 4), 2, 36,
                       merge_global=True, global_rand=0.0)
    case_compute_merge("compute_merge_exact_global_rand1", [chunk2(4, 16, 128), chunk2(4, 16, 128)], (4, 4), 2, 37,
                       merge_global=True, global_rand=1.0)


if __name__ == "__main__":
    main()
