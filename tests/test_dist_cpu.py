"""N>1 host logic on CPU: world_size-2 gloo processes exercise chunk sharding, the ragged token all-gather and
the global-token exchange rule of vidtome_b200/dist.py (the NCCL path on the GPU box runs the same code)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vidtome_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidtome_b200 import dist as vd
    try:
        # chunk sharding: 5 chunks over 2 ranks
        chunks = [torch.arange(i * 4, i * 4 + 4) for i in range(5)]
        mine = vd.shard_chunks(chunks)
        assert [int(c[0]) // 4 for c in mine] == ([0, 2, 4] if rank == 0 else [1, 3])
        # ragged all-gather: rank k holds L = 7 + 3k tokens
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn((2, 7 + 3 * rank, 16), generator=g).half()
        got = vd.all_gather_tokens(local)
        assert [t.shape[1] for t in got] == [7, 10]
        assert torch.equal(got[rank], local)
        other = torch.randn((2, 7 + 3 * (1 - rank), 16), generator=torch.Generator().manual_seed(100 + 1 - rank)).half()
        assert torch.equal(got[1 - rank], other)
        # exchange rule: rank k matches against rank (k-1) mod G
        glob = vd.exchange_global_tokens(local)
        assert torch.equal(glob, other)
        # same seed on every rank -> same forked generator draws
        vd.seed_all_ranks(7)
        from vidtome_b200.utils import init_generator
        gen = init_generator(torch.device("cpu"))
        draw = torch.randint(0, 4, (1,), generator=gen)
        draws = [torch.zeros_like(draw) for _ in range(world)]
        dist.all_gather(draws, draw)
        assert all(int(d) == int(draw) for d in draws)
        # lock-step check: equal chunk counts pass and report uniform / ragged lengths; unequal counts
        # raise on EVERY rank instead of leaving one of them waiting in a collective
        assert vd.check_step_lockstep([4, 4]) is True
        assert vd.check_step_lockstep([4, 4] if rank == 0 else [4, 2]) is False
        try:
            vd.check_step_lockstep([4, 4, 4] if rank == 0 else [4, 4])
            raise AssertionError("unequal chunk counts must raise")
        except RuntimeError as e:
            assert "different numbers of chunks" in str(e)
        np.save(os.path.join(out_dir, f"glob{rank}.npy"), glob.numpy())
        np.save(os.path.join(out_dir, f"local{rank}.npy"), local.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # the oracle's global stage fed the exchanged tokens (the parity definition of DESIGN.md §8)
    for rank in range(2):
        local = np.load(tmp_path / f"local{rank}.npy")
        glob = np.load(tmp_path / f"glob{rank}.npy")
        tokens = np.concatenate([local, glob], axis=1)
        m = O.bipartite_soft_matching_2s(tokens, local.shape[1], 0.8, False, unmerge_chunk=0)
        merged = m.merge(tokens)
        assert merged.shape[1] == tokens.shape[1] - m.r
        assert m.unmerge(merged).shape[1] == local.shape[1]


def test_single_process_defaults():
    from vidtome_b200 import dist as vd
    assert vd.world() == 1 and vd.rank() == 0
    t = torch.zeros(1, 3, 8)
    assert vd.all_gather_tokens(t)[0] is t
    assert vd.exchange_global_tokens(t) is None
    assert vd.shard_chunks([1, 2, 3]) == [1, 2, 3]
