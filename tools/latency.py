"""Latency of one blocking get_equity call (10,000 runs, 6 players, flop), with and without the one-query fast path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuron_poker_b200 as npk
for mode, fn in (("reference", lambda: npk.get_equity({"AS", "KS"}, {"2C", "7D", "KH"}, 6, 10000)),
                 ("uniform", lambda: npk.montecarlo({"AS", "KS"}, {"2C", "7D", "KH"}, 6, 10000))):
    for _ in range(50): fn()
    t0 = time.perf_counter(); n = 2000
    vals = [fn() for _ in range(n)]
    dt = time.perf_counter() - t0
    print(mode, "%.1f us/call  %.0f calls/s  mean equity %.4f" % (1e6 * dt / n, n / dt, sum(vals) / n), "single path" if not os.environ.get("NPK_NO_SINGLE_PATH") else "general path", flush=True)
