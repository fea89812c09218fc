"""The reference's own PYTHON implementations, run unmodified from baseline/_ref/ -- TEST INFRASTRUCTURE / CPU baseline.

`__graft_entry__.build()` copies the handful of reference files this path consists of (tools/montecarlo_python.py,
tools/hand_evaluator.py, tools/montecarlo_numpy2.py, tools/helper.py, gym_env/, agents/agent_consider_equity.py,
agents/agent_random.py) from /root/reference into the git-ignored baseline/_ref/ (it travels to the GPU box like
oracle/_ref/, it is never committed).  Nothing under neuron_poker_b200/ imports this module; bench.py uses it for the
`cpu_baseline` legs (Python run_montecarlo with timeout=+inf, get_equity as shipped, numpy2 flagged incorrect)
and tests/test_reference_env.py drives the real gym_env.env.HoldemTable with the GPU drop-in installed.

gymnasium, matplotlib and pyglet are absent from this image and irrelevant to the path: stub modules stand in for them
(the same stubs tests/golden/make_env_golden.py uses; nothing of the reference is modified).
"""
import contextlib
import io
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
FILES = ["tools/__init__.py", "tools/montecarlo_python.py", "tools/hand_evaluator.py", "tools/montecarlo_numpy2.py",
         "tools/helper.py", "gym_env/__init__.py", "gym_env/env.py", "gym_env/cycle.py", "gym_env/enums.py",
         "gym_env/rendering.py", "agents/__init__.py", "agents/agent_consider_equity.py", "agents/agent_random.py"]


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def install_copy(reference="/root/reference"):
    """Copy the reference's files for this path into baseline/_ref/ (build container only)."""
    import shutil
    for f in FILES:
        dst = os.path.join(REF_DIR, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, f), dst)
    return REF_DIR


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Env(object):
    def __init__(self):
        pass


class _Space(object):
    def __init__(self, *a, **kw):
        pass


_loaded = {}


def load(with_env=False):
    """Import the reference modules from baseline/_ref.  Returns a dict of modules."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    if "mc" not in _loaded:
        import tools.montecarlo_python as mc
        import tools.hand_evaluator as he
        import tools.montecarlo_numpy2 as np2
        assert os.path.dirname(os.path.abspath(mc.__file__)).startswith(REF_DIR), mc.__file__
        _loaded.update(mc=mc, he=he, np2=np2)
    if with_env and "env" not in _loaded:
        for missing in ("gymnasium", "matplotlib", "pyglet"):
            if missing not in sys.modules:
                try:
                    __import__(missing)
                except Exception:
                    if missing == "gymnasium":
                        sp = _stub("gymnasium.spaces", Discrete=_Space, Box=_Space)
                        reg = _stub("gymnasium.envs.registration", register=lambda **kw: None)
                        _stub("gymnasium", Env=_Env, spaces=sp, envs=_stub("gymnasium.envs", registration=reg))
                    elif missing == "matplotlib":
                        _stub("matplotlib", pyplot=_stub("matplotlib.pyplot"))
                    else:
                        _stub("pyglet")
        import gym_env.env as env
        import gym_env.enums as enums
        import agents.agent_consider_equity as ace
        import agents.agent_random as ar
        _loaded.update(env=env, enums=enums, agent_consider_equity=ace, agent_random=ar)
    return _loaded


RANKS, SUITS = "23456789TJQKA", "CDHS"


def _s(c):
    return RANKS[int(c) >> 2] + SUITS[int(c) & 3]


def python_run_montecarlo(hole, board, players, trials):
    """MonteCarlo.run_montecarlo with the 1-second cut-off disabled (timeout = +inf): all `trials` run.
    Returns (equity, seconds, trials actually run)."""
    mc = load()["mc"]
    sim = mc.MonteCarlo()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run_montecarlo([[_s(c) for c in hole]], [_s(c) for c in board if int(c) != 255], int(players), 1,
                           maxRuns=int(trials), timeout=time.time() + 1e9, ghost_cards='', opponent_range=1)
    return float(sim.equity), time.perf_counter() - t0, int(sim.runs)


def python_get_equity_as_shipped(hole, board, players, trials):
    """get_equity exactly as the env calls it (montecarlo_python.py:401-406): stops after 1 s of wall clock.
    Returns (equity, seconds, trials actually run)."""
    mc = load()["mc"]
    sim = mc.MonteCarlo()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run_montecarlo([[_s(c) for c in hole]], [_s(c) for c in board if int(c) != 255], int(players), 1,
                           maxRuns=int(trials), timeout=time.time() + 1, ghost_cards='', opponent_range=1)
    return float(sim.equity), time.perf_counter() - t0, int(sim.runs)


def numpy2_montecarlo(hole, board, players, trials):
    """numpy_montecarlo(my_cards, table, iterations, player_amount) (montecarlo_numpy2.py:333-346); its post-flop results
    are wrong upstream (all of its tests are skipped): timed, never compared.  Returns (equity, seconds, trials)."""
    np2 = load()["np2"]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        eq = np2.numpy_montecarlo([[_s(c) for c in hole]], [_s(c) for c in board if int(c) != 255],
                                  int(trials), int(players))
    return float(eq) / 100.0, time.perf_counter() - t0, int(trials)


def _job(args):
    kind, hole, board, players, trials = args
    fn = {"python": python_run_montecarlo, "numpy2": numpy2_montecarlo, "shipped": python_get_equity_as_shipped}[kind]
    return fn(hole, board, players, trials)


def fan_out(kind, queries, trials, processes):
    """`processes` worker processes, one query at a time each (the reference has no parallelism of its own: this is the
    whole-host figure of SURVEY 8d).  queries: list of (hole, board, players).  Returns (results, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")            # the parent may hold a CUDA context: never fork it
    jobs = [(kind, [int(c) for c in h], [int(c) for c in b], int(p), int(trials)) for h, b, p in queries]
    with ctx.Pool(processes) as pool:
        pool.map(_job, jobs[:processes])                      # warm the workers (imports)
        t0 = time.perf_counter()
        res = pool.map(_job, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return res, dt
