"""Host cost of enqueueing one cfg5 self-play step (Python + ctypes + CUDA launches) against its GPU time: the loop is timed on
the host while the GPU queue is kept short (a synchronisation every step) and while it runs ahead (no synchronisation)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuron_poker_b200.holdem import EquityAgents, HoldemTables
N, runs = 65536, 1000
tb = HoldemTables(N, n_players=6, seed=7, autoplay=[1] * 6)
agents = EquityAgents.equity_vs_random()
for _ in range(40):
    tb.selfplay_step(agents, runs=runs, deal_mode="uniform")
torch.cuda.synchronize()
K = 100
t0 = time.perf_counter()
for _ in range(K):
    tb.selfplay_step(agents, runs=runs, deal_mode="uniform")
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("enqueue %.1f us/step (host only), %.1f us/step until the GPU has finished" % (1e6 * t_enq / K, 1e6 * t_all / K), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    tb.selfplay_step(agents, runs=runs, deal_mode="uniform")
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
