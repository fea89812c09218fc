"""One launch of each kernel that is not the cfg3 headline, for `ncu -k regex:NAME -c 1`:
rank7_kernel (16 M hands), enum_kernel (1,024 turn + 1,024 river spots), equity_ranges_kernel (4,096 queries x 1,000 trials,
30 % range, reference dealer), the holdem kernels (65,536 tables, 3 self-play steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import neuron_poker_b200 as npk
from neuron_poker_b200.holdem import EquityAgents, HoldemTables

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
if what in ("rank7", "all"):
    hands = torch.rand(1 << 20, 52, generator=g).argsort(1)[:, :7].to(torch.uint8).to(dev).repeat(16, 1).contiguous()
    for _ in range(3):
        npk.rank7(hands)
if what in ("enum", "all"):
    Q = 1024
    cards = torch.rand(2 * Q, 52, generator=g).argsort(1)[:, :7].to(torch.uint8)
    board = cards[:, 2:7].clone()
    board[:Q, 4] = 255
    for _ in range(3):
        npk.enumerate_equity(cards[:, :2].contiguous().to(dev), board.to(dev))
if what in ("ranges", "all"):
    Q = 4096
    cards = torch.rand(Q, 52, generator=g).argsort(1)[:, :5].to(torch.uint8)
    board = torch.full((Q, 5), 255, dtype=torch.uint8)
    board[:, :3] = cards[:, 2:5]
    npl = torch.full((Q,), 6, dtype=torch.uint8)
    for _ in range(3):
        npk.get_equity_ranges_batch(cards[:, :2].contiguous(), board, npl, 1000, opponent_range=0.3, deal_mode="reference",
                                    validate=False, passes=len(sys.argv) > 2 and sys.argv[2] == "generic")
if what in ("holdem", "all"):
    tb = HoldemTables(65536, n_players=6, seed=7, autoplay=[1] * 6, device=dev)
    agents = EquityAgents.equity_vs_random()
    for _ in range(12):
        tb.selfplay_step(agents, runs=1000, deal_mode="uniform")
torch.cuda.synchronize()
print("done", what)
