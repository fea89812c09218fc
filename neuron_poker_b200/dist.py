"""Multi-GPU sharding of the equity path: one process per GPU, torch.distributed for the plumbing.

Every (query, trial) is independent and keyed by the Philox counter (seed; trial, query), so any partition gives
bit-identical totals.  Two ways to shard:
  * by QUERY  (Q >= world size): contiguous query blocks per rank, NO data-path collective; results are gathered.
  * by TRIAL  (few queries, many trials): rank r runs trials [r*T/W, (r+1)*T/W) of every query via trial_offset and
    one all-reduce(sum) of the [Q,2] (wins, ties) counter tensor combines them (a few KB over NVLink with NCCL).
The kernel accumulates straight into the tensor that is all-reduced: there is no pack/unpack step.
"""
import numpy as np


def query_shard(n_queries, rank, world):
    """[begin, end) of the contiguous query block owned by `rank`."""
    base, extra = divmod(int(n_queries), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def trial_shard(n_trials, rank, world):
    """(trial_offset, count) of the trial range owned by `rank`."""
    begin, end = query_shard(n_trials, rank, world)
    return begin, end - begin


def allreduce_counts(counts, group=None):
    """Sum a counter tensor over ranks in place (int64 view of the device's u64 counters; NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_query_blocks(local, n_queries, group=None):
    """All-gather per-rank blocks of a query-sharded result into the full [Q, ...] tensor (ragged last blocks padded)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    width = -(-int(n_queries) // world)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        b, e = query_shard(n_queries, r, world)
        out.append(p[: e - b])
    return torch.cat(out, 0)


def sharded_equity(hole, board, n_players, trials, seed_value=0, deal_mode="uniform", by="query", run=None,
                   uniform_shape=None, group=None):
    """Run get_equity_batch across the ranks of the default process group and return full [Q] wins / ties on every rank.

    `run(hole, board, n_players, trials, trial_offset, first_query) -> (wins, ties)` is the per-rank worker; by default
    it is the GPU path.  (The CPU tests of this module pass a stub worker so the sharding algebra is checked with gloo.)
    """
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Q = len(hole)
    if run is None:
        from .equity import get_equity_batch

        def run(h, b, p, t, trial_offset, first_query):
            # a query shard is numbered from `first_query` in the Philox counter (query_offset), a trial shard from
            # `trial_offset`: the shard reproduces exactly the numbers the whole job would have produced for it
            out = get_equity_batch(h, b, p, t, seed_value=seed_value, deal_mode=deal_mode, trial_offset=trial_offset,
                                   uniform_shape=uniform_shape, validate=False, query_offset=first_query)
            return out["wins"], out["ties"]

    if by == "query":
        b, e = query_shard(Q, rank, world)
        wins, ties = run(hole[b:e], board[b:e], n_players[b:e], trials, 0, b)
        return gather_query_blocks(wins, Q, group), gather_query_blocks(ties, Q, group)
    off, cnt = trial_shard(trials, rank, world)
    wins, ties = run(hole, board, n_players, cnt, off, 0)
    both = torch.stack([wins, ties], 1).contiguous()
    allreduce_counts(both, group)
    return both[:, 0], both[:, 1]
