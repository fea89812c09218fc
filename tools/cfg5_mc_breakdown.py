"""Where the Monte-Carlo phase of a cfg5 step goes: the mixed call against the sum of its per-shape kernels."""
import sys, os, time
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
hole, board, npl, active = [x.clone() for x in tb.queries()]
torch.cuda.synchronize()
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
t_mixed = timeit(lambda: get_equity_batch(hole, board, npl, runs, seed_value=1, deal_mode=mode, validate=False))
print("mixed call: %.3f ms" % (1e3 * t_mixed))
hb = (board != 255).sum(1)
tot = 0.0
for p in range(1, 7):
    for b in (0, 3, 4, 5):
        m = (npl == p) & (hb == b)
        c = int(m.sum())
        if not c: continue
        h, bd, n_ = hole[m].contiguous(), board[m].contiguous(), npl[m].contiguous()
        t = timeit(lambda: get_equity_batch(h, bd, n_, runs, seed_value=1, deal_mode=mode, validate=False, uniform_shape=(p, b)))
        evals = c * runs * p
        print("  players %d board %d: %6d queries  %.3f ms  %.0f G evals/s" % (p, b, c, 1e3 * t, evals / t / 1e9))
        tot += t
print("sum of per-shape calls: %.3f ms" % (1e3 * tot))
