"""CUDA-graph stepping (ChunkedDenoiser(cuda_graph=True)) must be indistinguishable from eager stepping: the frame
draws of the local matcher stay on the device during capture (vtm_split_t.randf_dev) and come from the blocks'
registered generators, so replays continue the same random sequence and produce bit-identical latents."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cuda_graph: bool, steps: int, frames: int = 8, size: int = 16, name: str = "tiny"):
    import vidtome_b200
    from vidtome_b200.driver import ChunkedDenoiser
    from vidtome_b200.skeleton import make_skeleton
    torch.manual_seed(77)
    net = make_skeleton(name, device="cuda")
    vidtome_b200.apply_patch(net, local_merge_ratio=0.9, batch_size=2)
    den = ChunkedDenoiser(net, n_timesteps=50, chunk_size=frames, cuda_graph=cuda_graph, graph_warmup=2)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((frames, 4, size, size), generator=g, device="cuda", dtype=torch.float16)
    xs, draws = [], []
    for i in range(steps):
        x = den.step(x, i)
        xs.append(x.clone())
    torch.cuda.synchronize()
    return xs, den


def test_graph_replay_is_bit_identical_to_eager():
    eager, _ = _run(False, 6)
    graph, den = _run(True, 6)
    assert den._graph is not None, "the graph was never captured"
    for i, (a, b) in enumerate(zip(eager, graph)):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b), f"step {i}: max |diff| = {(a.float() - b.float()).abs().max().item()}"
    # the draws differ from step to step (a frozen draw would also pass an eager/graph comparison of one step)
    assert not torch.equal(graph[2], graph[3])


def test_graph_draws_follow_the_generator():
    """Replays must draw fresh randf values: over many steps all `stride` values occur at level 1."""
    import vidtome_b200
    from vidtome_b200 import patch, utils
    from vidtome_b200.skeleton import make_skeleton
    torch.manual_seed(3)
    gen = torch.Generator(device="cuda").manual_seed(11)
    ref = [int(torch.randint(0, 4, (1,), generator=gen, device="cuda")) for _ in range(12)]
    gen.manual_seed(11)
    graph = torch.cuda.CUDAGraph()
    graph.register_generator_state(gen)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        t = utils.draw_randf(gen, 4, 8)
        assert isinstance(t, torch.Tensor) and t.dtype == torch.int32
    got = []
    for _ in range(12):
        graph.replay()
        got.append(int(t.item()))
    assert got == ref and len(set(got)) > 1


def test_capture_rejects_draw_dependent_token_counts():
    from vidtome_b200 import utils
    gen = torch.Generator(device="cuda").manual_seed(1)
    graph = torch.cuda.CUDAGraph()
    graph.register_generator_state(gen)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="divisible"):
        with torch.cuda.graph(graph):
            utils.draw_randf(gen, 4, 6)


def test_device_resident_randf_matches_host_randf():
    """K0 and the map composition with vtm_split_t.randf_dev == the same calls with the host value."""
    from vidtome_b200 import merge, ops
    from vidtome_b200._lib import VtmSplit
    g = torch.Generator(device="cuda").manual_seed(2)
    B, F, T, C = 2, 8, 64, 64
    x = torch.randn((B, F * T, C), generator=g, device="cuda").half()
    for randf in range(4):
        host = VtmSplit.local(F * T, 0, F, 4, randf)
        dev = VtmSplit.local(F * T, 0, F, 4, torch.tensor([randf], device="cuda", dtype=torch.int32))
        assert ops.split_counts(host) == ops.split_counts(dev)
        a0, b0 = ops.normalize_split(x, None, host)
        a1, b1 = ops.normalize_split(x, None, dev)
        assert torch.equal(a0, a1) and torch.equal(b0, b1)
        m0 = merge.match_level(x, None, host, 0.9, False, None)
        m1 = merge.match_level(x, None, dev, 0.9, False, None)
        mu0, pi0 = ops.compose_maps(host, m0.r, m0.keys, m0.edge, m0.rank, None, None, 0, F * T)
        mu1, pi1 = ops.compose_maps(dev, m1.r, m1.keys, m1.edge, m1.rank, None, None, 0, F * T)
        assert torch.equal(mu0, mu1) and torch.equal(pi0, pi1)
    with pytest.raises(RuntimeError):          # 6 frames, stride 4: counts would depend on the draw
        ops.split_counts(VtmSplit.local(6 * T, 0, 6, 4, torch.tensor([1], device="cuda", dtype=torch.int32)))


def test_two_gpu_global_exchange_bit_exact():
    """§8e on hardware: chunk-per-GPU global merging on 2 ranks (NCCL): the all-gather exchange and the exchange fused
    into the merge gather (peer stores over NVLink) both reproduce, bit for bit, the single-GPU computation fed the
    same global tokens (tools/check_dist_gpu.py).  Needs two GPUs on the box; skipped otherwise."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(root, "tools", "check_dist_gpu.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "allgather bit-exact=True" in r.stdout and "fused-p2p bit-exact=True" in r.stdout
