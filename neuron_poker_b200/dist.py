"""Multi-GPU sharding of the equity path: one process per GPU, torch.distributed for the plumbing.

Every (query, trial) is independent and keyed by the Philox counter (seed; trial, query), so any partition gives
bit-identical totals.  Two ways to shard:
  * by QUERY  (Q >= world size): contiguous query blocks per rank, NO data-path collective; results are gathered.
  * by TRIAL  (few queries, many trials): rank r runs trials [r*T/W, (r+1)*T/W) of every query via trial_offset and
    one all-reduce(sum) of the [Q,2] (wins, ties) counter tensor combines them (a few KB over NVLink with NCCL).
The kernel accumulates straight into the tensor that is all-reduced: there is no pack/unpack step.
TrialShardedJob goes one step further on NVLink-connected GPUs: the kernel's last warp exchanges the counters with the other
ranks through peer-mapped memory and sums them, so a step is ONE kernel launch and nothing else (no NCCL call, no memset).
"""
import ctypes

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


class TrialShardedJob(object):
    """A uniform-shape batch whose trials are split over the ranks of the default process group (cfg 4: 169 classes x
    1,000,000 trials over 2/4/8 GPUs).  `step(trials, seed)` runs this rank's trial range and returns the [2,Q] int64
    tensor (wins row, ties row) summed over all ranks, asynchronously on the current CUDA stream.

    reduction="peer" (default when every rank has its own GPU on one node): the count exchange happens inside the
    Monte-Carlo kernel over NVLink peer memory (npk_equity_batch_sharded, include/npk.h); the CUDA IPC handles are exchanged
    once with torch.distributed.  reduction="nccl": kernel + one NCCL all-reduce of the [2,Q] tensor."""

    def __init__(self, hole, board, n_players, uniform_shape, rank, world, deal_mode="uniform", reduction=None,
                 query_offset=0, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        from .equity import _DEAL
        self.torch, self.rank, self.world, self.group = torch, int(rank), int(world), group
        self.players, self.known = int(uniform_shape[0]), int(uniform_shape[1])
        self.deal_mode, self._deal = deal_mode, _DEAL[deal_mode]
        self.query_offset = int(query_offset)
        self.load(hole, board, n_players)
        self.device = self.hole.device
        self.Q = self.hole.shape[0]
        self.L = _lib.ensure_init(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.totals = torch.zeros((2, self.Q), dtype=torch.int64, device=self.device)
        self.reduction = reduction or ("peer" if self.world > 1 else "local")
        self.handle = None
        if self.reduction == "peer":
            h = ctypes.c_void_p()
            mine = np.zeros(64, dtype=np.uint8)
            with torch.cuda.device(self.device):
                _lib.check(self.L.npk_peer_create(self.rank, self.world, 2 * self.Q, ctypes.byref(h),
                                                  mine.ctypes.data_as(ctypes.c_void_p)))
                self.handle = h
                every = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
                dist.all_gather(every, torch.as_tensor(mine).to(self.device), group=group)
                allh = np.ascontiguousarray(torch.stack(every).cpu().numpy())
                _lib.check(self.L.npk_peer_connect(self.handle, allh.ctypes.data_as(ctypes.c_void_p)))
                torch.cuda.synchronize(self.device)
                dist.barrier(group=group)            # every rank has mapped every buffer before the first step
        self.launches_per_step = 1 if self.reduction == "peer" else (4 if self.reduction == "nccl" else 3)

    def load(self, hole, board, n_players):
        """Replace the queries (CUDA uint8 tensors [Q,2], [Q,5], [Q]; same Q)."""
        self.hole, self.board, self.n_players = hole.contiguous(), board.contiguous(), n_players.contiguous()

    def describe(self):
        if self.reduction == "peer":
            return ("trial ranges per rank; counters exchanged by the kernel's last warp over NVLink peer memory (%d x %d B "
                    "pushed per rank and step), summed in the same kernel: 1 launch per step, no NCCL on the step path"
                    % (self.world, 16 * self.Q))
        if self.reduction == "nccl":
            return "trial ranges per rank + NCCL all-reduce of the [2,Q] counters (zero, kernel, all-reduce per step)"
        return "one rank: the whole trial range"

    def step(self, trials, seed_value):
        torch = self.torch
        from . import _lib
        if self.reduction == "peer":
            with torch.cuda.device(self.device):
                _lib.check(self.L.npk_equity_batch_sharded(
                    self.handle, self.hole.data_ptr(), self.board.data_ptr(), self.n_players.data_ptr(), self.Q, int(trials),
                    self.players, self.known, ctypes.c_uint64(int(seed_value) & (2**64 - 1)), self.query_offset, self._deal,
                    self.totals.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
            return self.totals
        from .equity import get_equity_batch
        off, cnt = trial_shard(trials, self.rank, self.world)
        self.totals.zero_()
        get_equity_batch(self.hole, self.board, self.n_players, cnt, seed_value=seed_value, deal_mode=self.deal_mode,
                         trial_offset=off, uniform_shape=(self.players, self.known), validate=False,
                         out={"wins": self.totals[0], "ties": self.totals[1]}, query_offset=self.query_offset,
                         device=self.device)
        if self.reduction == "nccl":
            allreduce_counts(self.totals, self.group)
        return self.totals

    def check(self):
        """Synchronise and raise if a peer's counters did not arrive in an earlier step."""
        if self.reduction == "peer":
            from . import _lib
            err = ctypes.c_int(0)
            _lib.check(self.L.npk_peer_error(self.handle, ctypes.byref(err)))
            if err.value:
                raise RuntimeError("a rank's counters did not arrive within the kernel's time-out: the ranks' calls diverged")

    def close(self):
        if self.handle is not None:
            self.L.npk_peer_destroy(self.handle)
            self.handle = None
