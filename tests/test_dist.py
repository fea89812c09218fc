"""Sharding of the equity path over ranks, checked on CPU with gloo (world size 2).

The per-rank worker is a CPU stand-in for the GPU kernel -- the executable sampler specification scored by the oracle --
so what is tested here is the host-side algebra of neuron_poker_b200.dist: query blocks need no collective and gather
back in order; trial ranges are combined by one all-reduce; both reproduce the unsharded counts bit for bit because the
Philox counters carry global (query, trial) numbers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuron_poker_b200 import dist as npk_dist

QUERIES = [([51, 47], [0, 21, 46], 6), ([12, 13], [], 2), ([30, 31], [1, 2, 3, 4], 3), ([8, 40], [5, 6, 7, 9, 10], 4),
           ([20, 25], [11, 14, 15], 2)]
TRIALS, SEED = 48, 2026


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model_counts(first_query, queries, trials, trial_offset):
    import oracle
    import sampler_model
    wins, ties = [], []
    for i, (h, b, p) in enumerate(queries):
        m = sampler_model.run_model(oracle, "uniform", SEED, first_query + i, h, b, p, trials, trial_offset=trial_offset)
        wins.append(m["wins"])
        ties.append(m["ties"])
    return torch.tensor(wins, dtype=torch.int64), torch.tensor(ties, dtype=torch.int64)


def _worker(rank, world, port, by, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hole = [q[0] for q in QUERIES]
    board = [q[1] for q in QUERIES]
    npl = [q[2] for q in QUERIES]

    def run(h, b, p, t, trial_offset, first_query):
        return _model_counts(first_query, list(zip(h, b, p)), t, trial_offset)

    wins, ties = npk_dist.sharded_equity(hole, board, npl, TRIALS, seed_value=SEED, by=by, run=run)
    ret[rank] = (wins.tolist(), ties.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("by", ["query", "trial"])
def test_sharded_counts_equal_unsharded(by):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), by, ret), nprocs=world, join=True)
    ref_w, ref_t = _model_counts(0, QUERIES, TRIALS, 0)
    for r in range(world):
        assert ret[r] == (ref_w.tolist(), ref_t.tolist()), (by, r)


def test_shard_arithmetic():
    for n in (0, 1, 5, 169, 4096, 10000):
        for world in (1, 2, 3, 8):
            spans = [npk_dist.query_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sum(c for _, c in (npk_dist.trial_shard(n, r, world) for r in range(world))) == n


# ---- the same algebra on real GPUs over NCCL (skipped on boxes with fewer than two GPUs) -------------------------------
def _nccl_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["NPK_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(1)
    Q = 37
    cards = np.stack([rng.permutation(52)[:5] for _ in range(Q)]).astype(np.uint8)
    hole = torch.as_tensor(cards[:, :2].copy()).cuda()
    board = torch.full((Q, 5), 255, dtype=torch.uint8)
    board[:, :3] = torch.as_tensor(cards[:, 2:5].copy())
    board = board.cuda()
    npl = torch.full((Q,), 6, dtype=torch.uint8).cuda()
    out = {}
    for mode in ("uniform", "reference"):
        for by in ("query", "trial"):
            w, t = npk_dist.sharded_equity(hole, board, npl, 5001, seed_value=SEED, deal_mode=mode, by=by, uniform_shape=(6, 3))
            out[(mode, by)] = (w.cpu().tolist(), t.cpu().tolist())
        # the fused path: counters exchanged over NVLink peer memory inside the kernel, several steps in a row (both
        # parities of the exchange buffer, ranks running ahead of each other), and the NCCL variant of the same job
        for red in ("peer", "nccl"):
            job = npk_dist.TrialShardedJob(hole, board, npl, (6, 3), rank, world, deal_mode=mode, reduction=red)
            steps = []
            for i in range(5):
                tot = job.step(5001 + i, SEED + i)
                steps.append(tot.cpu().tolist())
            job.check()
            job.close()
            out[(mode, red)] = steps
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_sharded_counts_equal_unsharded():
    """Two ranks, two GPUs: query blocks (all-gather) and trial ranges (all-reduce of the counters) both reproduce the
    single-GPU counts bit for bit, in both dealing modes (5,001 trials: an odd split, so trial pairs straddle ranks)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import neuron_poker_b200 as npk
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    rng = np.random.default_rng(1)
    Q = 37
    cards = np.stack([rng.permutation(52)[:5] for _ in range(Q)]).astype(np.uint8)
    board = np.full((Q, 5), 255, dtype=np.uint8)
    board[:, :3] = cards[:, 2:5]
    for mode in ("uniform", "reference"):
        one = npk.get_equity_batch(cards[:, :2].copy(), board, np.full(Q, 6, dtype=np.uint8), 5001, seed_value=SEED,
                                   deal_mode=mode, uniform_shape=(6, 3))
        want = (one["wins"].cpu().tolist(), one["ties"].cpu().tolist())
        for r in range(world):
            for by in ("query", "trial"):
                assert ret[r][(mode, by)] == want, (mode, by, r)
        for i in range(5):
            one = npk.get_equity_batch(cards[:, :2].copy(), board, np.full(Q, 6, dtype=np.uint8), 5001 + i, seed_value=SEED + i,
                                       deal_mode=mode, uniform_shape=(6, 3))
            want = [one["wins"].cpu().tolist(), one["ties"].cpu().tolist()]
            for r in range(world):
                for red in ("peer", "nccl"):
                    assert ret[r][(mode, red)][i] == want, (mode, red, r, i)
