"""Time the phases of a cfg5 self-play step (CUDA events, synchronised between phases; averages over 50 steps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuron_poker_b200.holdem import EquityAgents, HoldemTables
from neuron_poker_b200.equity import get_equity_batch
mode = sys.argv[1] if len(sys.argv) > 1 else "uniform"
N, runs = 65536, 1000
tb = HoldemTables(N, n_players=6, seed=7, autoplay=[1] * 6)
agents = EquityAgents.equity_vs_random()
for _ in range(30):
    tb.selfplay_step(agents, runs=runs, deal_mode=mode)
torch.cuda.synchronize()
out = {"wins": torch.zeros(N, dtype=torch.int64, device="cuda"), "ties": torch.zeros(N, dtype=torch.int64, device="cuda")}
acc = {}
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return r
K = 50
for i in range(K):
    hole, board, npl, active = timed("queries", tb.queries)
    timed("zero", lambda: (out["wins"].zero_(), out["ties"].zero_()))
    timed("montecarlo", lambda: get_equity_batch(hole, board, npl, runs, seed_value=1000 + i, deal_mode=mode, validate=False, out=out))
    actions = timed("decide", lambda: tb.decide(agents, wins=out["wins"], ties=out["ties"], runs=runs))
    timed("step", lambda: tb.step(actions, restart_finished=True))
evals = float((npl.double() * active.double()).sum()) * runs
print(mode, {k: "%.3f ms" % (1e3 * v / K) for k, v in acc.items()}, "evals/step %.3g" % evals,
      "MC alone: %.1f G evals/s" % (evals / (acc["montecarlo"] / K) / 1e9), flush=True)
shapes = {}
hb = (board != 255).sum(1); 
for p in range(1, 7):
    for b in (0, 3, 4, 5):
        c = int(((npl == p) & (hb == b) & (active == 1)).sum())
        if c: shapes[(p, b)] = c
print("shape counts (players, known board cards):", shapes)
