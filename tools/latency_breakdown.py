"""Where the microseconds of one blocking get_equity call go: the Python wrapper, the ctypes crossing, launch + kernel +
result hand-over (npk_equity_one called in a tight loop from ctypes with constant arguments)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuron_poker_b200 as npk
from neuron_poker_b200 import _lib, equity

L = _lib.ensure_init(0)
out = (ctypes.c_uint64 * 12)()
p_out = ctypes.addressof(out)
packed = equity._pack_query({"AS", "KS"}, {"2C", "7D", "KH"})
for trials in (10000, 1000, 64):
    for _ in range(200):
        L.npk_equity_one(packed, 6, trials, 1, 1, 0, p_out)
    n = 3000
    t0 = time.perf_counter()
    for i in range(n):
        L.npk_equity_one(packed, 6, trials, i, 1, 0, p_out)
    dt = time.perf_counter() - t0
    print("npk_equity_one via ctypes, %5d trials: %.2f us/call" % (trials, 1e6 * dt / n))
n = 3000
t0 = time.perf_counter()
for i in range(n):
    equity._pack_query({"AS", "KS"}, {"2C", "7D", "KH"})
print("_pack_query: %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for i in range(n):
    equity._next_seed()
print("_next_seed: %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
for _ in range(100):
    npk.get_equity({"AS", "KS"}, {"2C", "7D", "KH"}, 6, 10000)
t0 = time.perf_counter()
for i in range(n):
    npk.get_equity({"AS", "KS"}, {"2C", "7D", "KH"}, 6, 10000)
print("get_equity: %.2f us/call" % (1e6 * (time.perf_counter() - t0) / n))
