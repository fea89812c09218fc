"""GPU timeline of a cfg5 self-play step: CUDA events recorded between the phases on the stream (no synchronisation inside the
step), averaged over 100 steps -- each interval is kernel time plus whatever gap precedes the next launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuron_poker_b200.holdem import EquityAgents, HoldemTables
from neuron_poker_b200.equity import get_equity_batch
mode = sys.argv[1] if len(sys.argv) > 1 else "uniform"
N, runs = 65536, 1000
tb = HoldemTables(N, n_players=6, seed=7, autoplay=[1] * 6)
agents = EquityAgents.equity_vs_random()
for _ in range(40):
    tb.selfplay_step(agents, runs=runs, deal_mode=mode)
torch.cuda.synchronize()
both = torch.zeros((2, N), dtype=torch.int64, device="cuda")
out = {"wins": both[0], "ties": both[1]}
K = 100
names = ["queries", "zero", "montecarlo", "decide", "step"]
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(K)]
for i in range(K):
    ev[i][0].record()
    hole, board, npl, active = tb.queries(); ev[i][1].record()
    both.zero_(); ev[i][2].record()
    get_equity_batch(hole, board, npl, runs, seed_value=1000 + i, deal_mode=mode, validate=False, out=out); ev[i][3].record()
    actions = tb.decide(agents, wins=out["wins"], ties=out["ties"], runs=runs); ev[i][4].record()
    tb.step(actions, restart_finished=True); ev[i][5].record()
torch.cuda.synchronize()
tot = ev[0][0].elapsed_time(ev[K - 1][5]) / K
print(mode, "step %.1f us:" % (1e3 * tot), "  ".join("%s %.1f" % (n, 1e3 * sum(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(K)) / K)
                                                   for j, n in enumerate(names)),
      " between steps %.1f" % (1e3 * sum(ev[i][5].elapsed_time(ev[i + 1][0]) for i in range(K - 1)) / (K - 1)), flush=True)
