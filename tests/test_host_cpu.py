"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the pure-host entry points
agree with the oracle, the patch API behaves like the reference's (class swap, hooks, attribute
broadcast, error conventions), and the product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import vidtome_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from vidtome_b200 import _lib
    return _lib.load()


def test_library_exports_every_symbol_declared_in_header():
    header = open(os.path.join(ROOT, "include", "vidtome_b200.h")).read()
    declared = set(re.findall(r"\b(vtm_[a-z0-9_]+)\s*\(", header))
    declared.discard("vtm_key_score")  # mentioned in a comment only
    from vidtome_b200 import _lib
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().vtm_version() == 100


def test_library_has_no_driver_link_dependency():
    """The library must load on a machine without libcuda (this container has none)."""
    import subprocess
    from vidtome_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


def test_error_strings():
    lib = _lib()
    assert lib.vtm_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert len(lib.vtm_error_string(code)) > 3


@pytest.mark.parametrize("N,unm_pre,F,stride", [(4 * 96, 0, 4, 4), (5 * 64 + 37, 37, 5, 4), (6 * 10, 0, 6, 4),
                                               (2 * 50 + 11, 11, 2, 4), (3 * 7, 0, 3, 4), (16 * 8, 0, 16, 4)])
def test_split_counts_match_oracle(N, unm_pre, F, stride):
    from vidtome_b200 import ops
    from vidtome_b200._lib import VtmSplit
    for randf in range(min(stride, F)):
        a_idx, b_idx, _, _ = O.split_indices_randframe(N, F, unm_pre, stride, randf)
        assert ops.split_counts(VtmSplit.local(N, unm_pre, F, stride, randf)) == (len(a_idx), len(b_idx))
    assert ops.split_counts(VtmSplit.prefix(100, 37)) == (37, 63)


def test_split_rejects_inconsistent_descriptor():
    from vidtome_b200._lib import VtmSplit
    lib = _lib()
    bad = VtmSplit(0, 100, 0, 3, 33, 3, 0, 0)          # 3 * 33 != 100
    assert lib.vtm_split_counts(ctypes.byref(bad), None, None) == -3
    bad = VtmSplit(0, 99, 0, 3, 33, 3, 3, 0)           # randf out of range
    assert lib.vtm_split_counts(ctypes.byref(bad), None, None) == -3
    bad = VtmSplit(1, 10, 0, 0, 0, 0, 0, 11)           # src_len > N
    assert lib.vtm_split_counts(ctypes.byref(bad), None, None) == -3


def test_merge_count_matches_python_truncation():
    from vidtome_b200 import ops
    rng = np.random.default_rng(0)
    for ns in [0, 1, 7, 3072, 12288, 49152, 5325, 55296, 9216] + list(rng.integers(1, 100000, 50)):
        for ratio in (0.9, 0.8, 0.95, 0.5, 1.0, 1.7, 0.3333333, 1e-9):
            assert ops.merge_count(int(ns), ratio) == O.merge_count(int(ns), ratio)


def test_compute_entry_points_validate_arguments_without_a_device():
    lib = _lib()
    assert lib.vtm_sim_argmax(None, None, 1, 1, 1, 8, 0, None, None) == -1
    assert lib.vtm_sim_argmax(1, 1, 1, 16, 16, 12, 0, 1, None) == -2          # C % 8 != 0
    assert lib.vtm_gather_rows(None, 0, None, 0, 1, 1, 8, None, 0, None) == -1
    assert lib.vtm_topr_sort(1, 1, 16, 1, 1, 1, 0, None) == -4                # workspace too small
    assert lib.vtm_linear_f16(1, 1, None, 16, 12, 64, 1, 12, None) == -2
    from vidtome_b200._lib import VtmSplit
    sp = VtmSplit.prefix(20, 8)
    assert lib.vtm_merge_reduce(1, 0, ctypes.byref(sp), 4, 1, 1, 1, 2, 16, 9, 1, 1, 1 << 30, None) == -6   # unknown mode
    assert lib.vtm_merge_reduce(1, 0, ctypes.byref(sp), 4, 1, 1, 1, 2, 16, 1, 1, 1, 16, None) == -4        # workspace
    assert lib.vtm_merge_reduce_workspace_bytes(2, 12, 16) >= 2 * 12 * 16 * 8 + 2 * 12 * 4
    # KC with peer destinations: at most 8 peers, and a peer list is required when n_peers > 0
    assert lib.vtm_gather_rows_peers(1, 0, None, 0, 1, 4, 8, None, None, 0.0, 1, 0, None, 1, None) == -1
    arr = (ctypes.c_void_p * 9)(*[ctypes.c_void_p(1)] * 9)
    assert lib.vtm_gather_rows_peers(1, 0, None, 0, 1, 4, 8, None, None, 0.0, 1, 0, arr, 9, None) == -2


# --------------------------------------------------------------------------- patch API (no CUDA needed)
def _skeleton():
    from vidtome_b200.skeleton import make_skeleton
    return make_skeleton("tiny", device="cpu", dtype=torch.float32)


def test_package_exports_match_reference():
    import vidtome_b200
    assert vidtome_b200.__all__ == ["merge", "patch", "apply_patch", "remove_patch", "update_patch",
                                    "collect_from_patch"]
    import inspect
    sig = inspect.signature(vidtome_b200.apply_patch)
    assert list(sig.parameters) == ["model", "local_merge_ratio", "merge_global", "global_merge_ratio",
                                    "max_downsample", "seed", "batch_size", "include_control", "align_batch",
                                    "target_stride", "global_rand"]
    d = {k: v.default for k, v in sig.parameters.items() if k != "model"}
    assert d == dict(local_merge_ratio=0.9, merge_global=False, global_merge_ratio=0.8, max_downsample=2,
                     seed=123, batch_size=2, include_control=False, align_batch=False, target_stride=4,
                     global_rand=0.5)


def test_apply_and_remove_patch_swap_classes_and_hooks():
    import vidtome_b200
    net = _skeleton()
    out = vidtome_b200.apply_patch(net, local_merge_ratio=0.8, merge_global=True, batch_size=3)
    assert out is net
    blocks = [m for m in net.modules() if type(m).__name__ == "ToMeBlock"]
    assert len(blocks) == 4
    info = net._tome_info
    assert info["size"] is None and len(info["hooks"]) == 1 + 4
    assert info["args"] == dict(max_downsample=2, generator=None, seed=123, batch_size=3, align_batch=False,
                                merge_global=True, global_merge_ratio=0.8, local_merge_ratio=0.8,
                                global_rand=0.5, target_stride=4)
    assert all(b._tome_info is info for b in blocks)
    assert all(type(b)._parent.__name__ == "BasicTransformerBlock" for b in blocks)
    # update_patch reaches the UNet and every block (patch.py:358-370)
    vidtome_b200.update_patch(net, global_tokens=None, foo=3)
    assert net.foo == 3 and all(b.foo == 3 and b.global_tokens is None for b in blocks)
    got = vidtome_b200.collect_from_patch(net, attr="foo")
    assert set(got.values()) == {3} and len(got) == 5 and "" in got
    # re-applying first removes the old patch (patch.py:277)
    vidtome_b200.apply_patch(net)
    assert len(net._tome_info["hooks"]) == 5
    vidtome_b200.remove_patch(net)
    assert not [m for m in net.modules() if type(m).__name__ == "ToMeBlock"]
    assert net._tome_info["hooks"] == []


def test_apply_patch_rejects_unsupported_model():
    import vidtome_b200
    with pytest.raises(RuntimeError, match="not a Stable Diffusion"):
        vidtome_b200.apply_patch(torch.nn.Linear(2, 2))


def test_pipeline_like_object_with_unet_attribute():
    import vidtome_b200

    class DiffusionPipeline:           # matched by NAME (patch.py:279)
        def __init__(self, unet):
            self.unet = unet

    pipe = DiffusionPipeline(_skeleton())
    vidtome_b200.apply_patch(pipe)
    assert hasattr(pipe.unet, "_tome_info")
    assert vidtome_b200.remove_patch(pipe) is pipe.unet


def test_no_cpu_fallback():
    """CPU tensors must be refused loudly — never computed some other way."""
    import vidtome_b200
    from vidtome_b200 import merge
    x = torch.randn(2, 64, 16)
    g = torch.Generator().manual_seed(0)
    with pytest.raises(RuntimeError, match="CUDA"):
        merge.bipartite_soft_matching_randframe(x.half(), 4, 0.9, 0, g)
    with pytest.raises(RuntimeError, match="CUDA"):
        merge.bipartite_soft_matching_2s(x.half(), 32, 0.8, False)
    net = _skeleton()
    vidtome_b200.apply_patch(net)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.randn(8, 4, 8, 8), 0)
    # ratio <= 0 keeps the reference's identity behaviour and arity (merge.py:45-46, :364-365)
    m, u, ret = merge.bipartite_soft_matching_randframe(x, 4, 0.0, 0, g)
    assert m(x) is x and ret == {"unm_num": 16}
    assert len(merge.bipartite_soft_matching_2s(x, 32, 0.0, False)) == 2


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vidtome_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "vidtome_oracle" not in src and "import oracle" not in src, f


def test_generators_stay_in_lockstep_hooks():
    """hook_tome_module forks the default RNG state into module.generator on first forward
    (patch.py:215-231); all blocks get the same state."""
    import vidtome_b200
    from vidtome_b200 import patch
    net = _skeleton()
    vidtome_b200.apply_patch(net, max_downsample=0)    # nothing merges -> runs on CPU through plain attention
    torch.manual_seed(5)
    net(torch.randn(8, 4, 8, 8), 0)
    states = [b.generator.get_state() for b in net.blocks]
    assert all(torch.equal(states[0], s) for s in states[1:])
    assert net._tome_info["size"] == (8, 8)


def test_markstein_division_used_by_k0_is_exact():
    """K0 divides a row by its fp16-rounded norm with q0 = RN(x r), rem = x - q0 n (FMA), q = RN(q0 + rem r),
    r = RN(1/n).  This equals IEEE x / n for every fp16 numerator and fp16 normal norm (so the kernel keeps
    torch's `metric / metric.norm()` semantics, merge.py:84, at one division per row)."""
    x = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16).astype(np.float32)
    rng = np.random.default_rng(0)
    norms = np.unique(rng.integers(0x0400, 0x7BFF, size=600).astype(np.uint16)).view(np.float16).astype(np.float32)
    for n in norms:
        n32 = np.float32(n)
        r = np.float32(1.0) / n32
        q0 = (x * r).astype(np.float32)
        rem = (x.astype(np.float64) - q0.astype(np.float64) * np.float64(n32)).astype(np.float32)
        q = (q0.astype(np.float64) + rem.astype(np.float64) * np.float64(r)).astype(np.float32)
        assert np.array_equal(q, x / n32)


def test_device_resident_randf_descriptor_validation():
    """vtm_split_t.randf_dev (CUDA-graph capture): counts must not depend on the draw, i.e. F % stride == 0."""
    from vidtome_b200._lib import VtmSplit
    lib = _lib()
    ns, nd = ctypes.c_int32(), ctypes.c_int32()
    fake = ctypes.c_void_p(0x1000)                      # never dereferenced by the host helper
    ok = VtmSplit(0, 8 * 64, 0, 8, 64, 4, 3, 0, fake)  # `randf` (3) is ignored when randf_dev is set
    assert lib.vtm_split_counts(ctypes.byref(ok), ctypes.byref(ns), ctypes.byref(nd)) == 0
    assert (ns.value, nd.value) == (6 * 64, 2 * 64)
    bad = VtmSplit(0, 6 * 64, 0, 6, 64, 4, 0, 0, fake)  # 6 frames, stride 4: 1 or 2 dst frames depending on randf
    assert lib.vtm_split_counts(ctypes.byref(bad), ctypes.byref(ns), ctypes.byref(nd)) == -3
    with pytest.raises(RuntimeError, match="int32 CUDA tensor"):
        VtmSplit.local(8 * 64, 0, 8, 4, torch.zeros(1, dtype=torch.int32))      # CPU tensor


def test_draw_randf_replays_the_reference_draw_on_cpu():
    """Eagerly, draw_randf is exactly the reference's `torch.randint(0, stride, [1], generator=...)` (merge.py:56-57)."""
    from vidtome_b200 import utils
    g1 = torch.Generator().manual_seed(9)
    g2 = torch.Generator().manual_seed(9)
    want = [int(torch.randint(0, s, torch.Size([1]), generator=g1)) for s in (4, 4, 2, 4, 3)]
    got = [utils.draw_randf(g2, s, 8) for s in (4, 4, 2, 4, 3)]
    assert got == want and all(isinstance(v, int) for v in got)


def test_cuda_graph_mode_rejects_dynamic_work():
    from vidtome_b200.driver import ChunkedDenoiser
    net = _skeleton()
    with pytest.raises(ValueError, match="cuda_graph"):
        ChunkedDenoiser(net, merge_global=True, cuda_graph=True)
    with pytest.raises(ValueError, match="cuda_graph"):
        ChunkedDenoiser(net, randomize_chunks=True, cuda_graph=True)
    den = ChunkedDenoiser(net, cuda_graph=True)        # CPU tensors simply keep stepping eagerly
    assert den._graph is None



# --------------------------------------------------------------------------- round 2: lifecycle on a ControlNet pipeline
def test_lifecycle_on_controlnet_pipeline_matches_reference_trace():
    """apply / update / collect / remove on a fake StableDiffusionControlNetPipeline (two skeleton nets) reach the
    same modules and return the same objects as the reference did (tests/golden/lifecycle_controlnet.json, written by
    tests/make_golden_r02.py from vidtome/patch.py:292-295,337-387): update_patch / collect_from_patch look for
    `.controlnet` on the object passed in, remove_patch on the UNet."""
    import json
    import sys
    import vidtome_b200
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "lifecycle_controlnet.json")))

    def factory():
        from vidtome_b200.skeleton import make_skeleton

        class DiffusionPipeline:
            pass

        class StableDiffusionControlNetPipeline(DiffusionPipeline):
            def __init__(self):
                self.unet = make_skeleton("tiny", device="cpu", dtype=torch.float32, seed=1)
                self.controlnet = make_skeleton("tiny", device="cpu", dtype=torch.float32, seed=2)
        return StableDiffusionControlNetPipeline()

    # the same tracing code that produced the fixture, minus the reference import
    src = open(os.path.join(ROOT, "tests", "make_golden_r02.py")).read()
    ns = {}
    start, end = src.index("def lifecycle_trace"), src.index("def case_mean")
    exec(src[start:end], ns)
    got = ns["lifecycle_trace"](vidtome_b200, factory)
    assert got == want


def test_update_patch_resets_controlnet_global_tokens():
    """with include_control=True and merge_global, update_patch(pipe, global_tokens=None) must reach the
    ControlNet blocks too, otherwise last step's global tokens leak into the next step."""
    import vidtome_b200
    from vidtome_b200.skeleton import make_skeleton

    class StableDiffusionControlNetPipeline:
        def __init__(self):
            self.unet = make_skeleton("tiny", device="cpu", dtype=torch.float32)
            self.controlnet = make_skeleton("tiny", device="cpu", dtype=torch.float32)

    class DiffusionPipeline(StableDiffusionControlNetPipeline):
        pass
    pipe = DiffusionPipeline()
    vidtome_b200.apply_patch(pipe, include_control=True, merge_global=True)
    for net in (pipe.unet, pipe.controlnet):
        for b in net.blocks:
            b.global_tokens = torch.zeros(1)
    vidtome_b200.update_patch(pipe, global_tokens=None)
    assert all(b.global_tokens is None for net in (pipe.unet, pipe.controlnet) for b in net.blocks)


def test_kd_fast_path_only_for_stock_attention_modules():
    """KD reads to_q/to_k/to_v/to_out[0].weight directly, so it may replace attn1.forward only when that
    cannot change the result.  LoRA-style wrappers (still exposing .weight), custom processors, hooks, instance
    forward overrides (PnP) and the Attention options that alter the math must all fall back to self.attn1(...)."""
    from vidtome_b200 import patch
    from vidtome_b200.skeleton import Attention

    def fresh():
        return Attention(64, 2, 32)
    assert patch._plain_attention_module(fresh())

    class LoRACompatibleLinear(torch.nn.Linear):          # diffusers' wrapper: a Linear subclass with an extra term
        def forward(self, x):
            return super().forward(x) + 1.0
    a = fresh(); a.to_q = LoRACompatibleLinear(64, 64, bias=False)
    assert not patch._plain_attention_module(a)

    class PeftLike(torch.nn.Module):                      # PEFT: wraps the base layer, forwards .weight / .bias
        def __init__(self, base):
            super().__init__(); self.base_layer = base
        weight = property(lambda self: self.base_layer.weight)
        bias = property(lambda self: self.base_layer.bias)
        def forward(self, x):
            return self.base_layer(x) * 2
    a = fresh(); a.to_out[0] = PeftLike(a.to_out[0])
    assert not patch._plain_attention_module(a)

    class LoRAAttnProcessor:
        pass

    class AttnProcessor2_0:
        pass
    a = fresh(); a.processor = LoRAAttnProcessor()
    assert not patch._plain_attention_module(a)
    a = fresh(); a.processor = AttnProcessor2_0()
    assert patch._plain_attention_module(a)
    a = fresh(); a.register_forward_hook(lambda m, i, o: o)
    assert not patch._plain_attention_module(a)
    a = fresh(); a.to_v.register_forward_pre_hook(lambda m, i: None)
    assert not patch._plain_attention_module(a)
    a = fresh(); a.forward = lambda *args, **kw: None     # PnP's injected forward (utils/pnp_utils.py:99-101)
    assert not patch._plain_attention_module(a)
    for attr, val in (("residual_connection", True), ("rescale_output_factor", 2.0), ("group_norm", torch.nn.GroupNorm(1, 64)),
                      ("added_kv_proj_dim", 8)):
        a = fresh(); setattr(a, attr, val)
        assert not patch._plain_attention_module(a), attr
    a = fresh(); a.to_out[1] = torch.nn.Dropout(0.5); a.train()
    assert not patch._plain_attention_module(a)
    a = fresh(); a.to_k = torch.nn.Linear(64, 64, bias=True)
    assert not patch._plain_attention_module(a)


def test_pnp_torch_forward_equals_reference_fixture_on_cpu():
    """pnp.register_attention_control installs a forward that IS the reference's PnP forward (utils/pnp_utils.py:47-95):
    bit-identical to the fixture the reference produced (tests/make_golden_r02.py --pnp), injection on and off."""
    from vidtome_b200 import patch, pnp
    from vidtome_b200.skeleton import Attention
    g = np.load(os.path.join(ROOT, "tests", "golden", "pnp_attention_b3.npz"))
    attn = Attention(int(g["dim"]), int(g["heads"]), int(g["dim"]) // int(g["heads"])).half()
    attn.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")})
    pnp.register_attention_control(None, [981, 961], 3, modules=[attn])
    x = torch.from_numpy(g["x"])
    for t, key in ((981, "out_inject"), (961, "out_inject"), (1000, "out_inject"), (500, "out_plain")):
        pnp.register_time(None, t, modules=[attn])
        with torch.no_grad():
            assert torch.equal(attn(x), torch.from_numpy(g[key])), (t, key)
    # our own override keeps the module eligible for KD; a foreign override does not
    assert patch._plain_attention_module(attn)
    attn.forward = lambda *a, **k: None
    assert not patch._plain_attention_module(attn)


# --------------------------------------------------------------------------- callers
def test_get_chunks_matches_reference_for_every_order_mode():
    """driver.ChunkedDenoiser.get_chunks against `Generator.get_chunks` itself (generate.py:172-203; its source is
    extracted from the reference file and run by tests/make_golden_r02.py --callers): same numpy / torch RNG seeds ->
    the same chunk lists for seq, rand, mix and mix-# orders, with and without global merging, short first chunks."""
    import json
    from vidtome_b200.driver import ChunkedDenoiser
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "callers_get_chunks.json")))
    assert len(cases) >= 9
    for c in cases:
        den = ChunkedDenoiser(_skeleton(), chunk_size=c["chunk_size"], merge_global=c["merge_global"],
                              chunk_ord=c["chunk_ord"], randomize_chunks=True)
        np.random.seed(c["seed"])
        torch.manual_seed(c["seed"])
        got = [[int(v) for v in ch] for ch in den.get_chunks(c["flen"])]
        assert got == c["chunks"], c


def test_ddim_update_matches_reference_pred_next_x():
    """driver.pred_next_x against `Generator.pred_next_x` (generate.py:281-311, extracted the same way) on SD's
    scaled-linear schedule: first, middle and last steps."""
    from vidtome_b200.driver import ChunkedDenoiser
    g = np.load(os.path.join(ROOT, "tests", "golden", "callers_pred_next_x.npz"))
    den = ChunkedDenoiser(_skeleton(), n_timesteps=50)
    x, eps = torch.from_numpy(g["x"]), torch.from_numpy(g["eps"])
    for i in (0, 1, 25, 48, 49):
        want = torch.from_numpy(g[f"i{i}"]).float()
        got = den.pred_next_x(x, eps, i).float()
        assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-3, i
