"""torchrun check of the chunk-per-GPU global-merge exchange (NCCL):
  torchrun --nproc-per-node 2 tools/check_dist_gpu.py
Every rank merges its own chunk locally, all-gathers the merged tokens, and matches against the tokens of rank
k-1.  The result is compared, bit for bit, with the same computation done on one GPU from the all-gathered
inputs (the reference's `_2s` semantics fed the exchanged global tokens)."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import dist as vd, patch  # noqa: E402


def info(merge_global):
    return {"size": (32, 32), "hooks": [], "args": dict(max_downsample=2, generator=None, seed=123, batch_size=2,
            align_batch=False, merge_global=merge_global, global_merge_ratio=0.8, local_merge_ratio=0.9,
            global_rand=0.5, target_stride=4)}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, F, T, C = 2, 4, 1024, 320
    xs = []
    for k in range(world):          # every rank can rebuild every chunk (seeded by chunk id)
        g = torch.Generator(device="cuda").manual_seed(1000 + k)
        base = torch.randn((B, 1, T, C), generator=g, device="cuda")
        xs.append((base + 0.1 * torch.randn((B, F, T, C), generator=g, device="cuda")).half().reshape(B * F, T, C))
    gen = lambda: torch.Generator(device="cuda").manual_seed(7)

    patch.GLOBAL_EXCHANGE = "allgather"
    mod = SimpleNamespace(generator=gen(), global_tokens=None)
    plan = patch.build_merge_plan(mod, xs[rank], info(True))

    # the same exchange fused into the merge gather (peer stores over NVLink + one barrier), three rounds so that both
    # alternating buffers and their reuse are exercised
    p2p_ok = True
    for mode in ("p2p", "p2p_all"):
        patch.GLOBAL_EXCHANGE = mode
        for _ in range(3):
            mod2 = SimpleNamespace(generator=gen(), global_tokens=None)
            plan2 = patch.build_merge_plan(mod2, xs[rank], info(True))
            p2p_ok &= torch.equal(plan2.merged_tokens, plan.merged_tokens) and torch.equal(plan2.pi, plan.pi)

    def timed(mode, n=20):
        patch.GLOBAL_EXCHANGE = mode
        for _ in range(3):
            patch.build_merge_plan(SimpleNamespace(generator=gen(), global_tokens=None), xs[rank], info(True))
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            patch.build_merge_plan(SimpleNamespace(generator=gen(), global_tokens=None), xs[rank], info(True))
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_ag, ms_p2p = timed("allgather"), timed("p2p")

    # single-GPU restatement from the same inputs: local merge of chunk k-1 -> its tokens are this rank's global set
    patch.GLOBAL_EXCHANGE = "recurrence"
    prev = (rank - 1) % world
    m_prev = SimpleNamespace(generator=gen(), global_tokens=None)
    glob = patch.build_merge_plan(m_prev, xs[prev], info(False)).merged_tokens
    m_self = SimpleNamespace(generator=gen(), global_tokens=glob)
    ref = patch.build_merge_plan(m_self, xs[rank], info(True))
    ok = torch.equal(plan.merged_tokens, ref.merged_tokens) and torch.equal(plan.pi, ref.pi)
    flag = torch.tensor([int(ok), int(p2p_ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist global exchange world={world} merged={tuple(plan.merged_tokens.shape)} "
              f"allgather bit-exact={bool(flag[0].item())} fused-p2p bit-exact={bool(flag[1].item())} "
              f"merge-plan ms/call (max over ranks): allgather {ms_ag:.3f}, fused p2p {ms_p2p:.3f}")
    dist.destroy_process_group()
    sys.exit(0 if bool(flag.min().item()) else 1)


if __name__ == "__main__":
    main()
