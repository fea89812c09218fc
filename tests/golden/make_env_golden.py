#!/usr/bin/env python3
"""Golden traces of the UNMODIFIED reference environment (gym_env/env.py, gym_env/cycle.py) for the vectorised
HoldemTable.  Run in the build container only (needs /root/reference):

    python tests/golden/make_env_golden.py

How the reference is driven:
  * gymnasium, matplotlib and pyglet are absent here and irrelevant to the state machine: three stub modules are put
    into sys.modules before `gym_env.env` is imported (nothing of the reference is modified or copied);
  * the env deals with np.random.randint(0, len(deck)) (env.py:680, 686); during a game that function is replaced by
    the Philox stream libnpk's tables use (draw k of table t = hi32(word_k * len), include/npk_holdem.h), so both deal
    the same cards;
  * get_equity is replaced by a constant (the traces are about the state machine; equity only feeds agents);
  * every seat is a non-autoplay player (the reference's own tests/test_gym_env.py::PlayerForTest pattern) and the
    actions come from a seeded policy: mostly a random legal move (all-ins down-weighted), sometimes an illegal one.

Files written:
  env_traces.npz     per game: its configuration, the action sequence and the table state after reset and after
                     every step (stage, seats, stacks, pots, cards, legal moves, cycle counters, reward, done, and the
                     observation vector array_everything with the stub equity 0.5)
  agent_cases.json   agents/agent_consider_equity.py::Player.action on random (equity, legal moves, thresholds)
"""
import contextlib
import io
import json
import os
import random
import sys
import types

import numpy as np

REF = os.environ.get("NPK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))          # tests/ (sampler_model.philox4x32_10)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Env:
    def __init__(self):
        pass


class _Space:
    def __init__(self, *a, **kw):
        pass


_spaces = _mod("gymnasium.spaces", Discrete=_Space, Box=_Space)
_reg = _mod("gymnasium.envs.registration", register=lambda **kw: None)
_mod("gymnasium", Env=_Env, spaces=_spaces, envs=_mod("gymnasium.envs", registration=_reg))
_mod("matplotlib", pyplot=_mod("matplotlib.pyplot"))
_mod("pyglet")

from gym_env.env import HoldemTable  # noqa: E402
from gym_env.enums import Action, Stage  # noqa: E402
import tools.montecarlo_python as _mp  # noqa: E402
from sampler_model import philox4x32_10, MASK  # noqa: E402

_mp.get_equity = lambda *a: 0.5
RANKS, SUITS = "23456789TJQKA", "CDHS"
SEED = 20261018


def cid(card):
    return 4 * RANKS.index(card[0]) + SUITS.index(card[1])


class Dealer:
    """np.random.randint replacement: draw k of table t from the Philox stream of include/npk_holdem.h."""

    def __init__(self, seed, table):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.table, self.k = table, 0

    def randint(self, low, high=None, *a, **kw):
        assert low == 0 and high is not None
        w = philox4x32_10(((self.k >> 2) & MASK, (self.k >> 34) & MASK, self.table & MASK, 0xD0000000), self.key)[self.k & 3]
        self.k += 1
        return (w * high) >> 32


class Seat:
    def __init__(self):
        self.name = "t"


FIELDS_SEAT = ["stack", "player_pots", "player_max_win", "can_still", "out_of_cash", "folder", "alive"]


def snapshot(env, dealer, reward):
    n = len(env.players)
    cp = env.current_player
    cyc = env.player_cycle
    cards = np.full((10, 2), 255, dtype=np.uint8)
    for i, p in enumerate(env.players):
        for j, c in enumerate(p.cards or []):
            cards[i, j] = cid(c)
    tc = np.full(5, 255, dtype=np.uint8)
    for j, c in enumerate(env.table_cards or []):
        tc[j] = cid(c)
    seat = np.zeros((7, 10))
    seat[0, :n] = [p.stack for p in env.players]
    seat[1, :n] = env.player_pots
    seat[2, :n] = env.player_max_win
    seat[3, :n] = cyc.can_still_make_moves_in_this_hand
    seat[4, :n] = cyc.out_of_cash_but_contributed
    seat[5, :n] = cyc.folder
    seat[6, :n] = cyc.alive
    legal = 0
    for a in env.legal_moves or []:
        legal |= 1 << a.value
    scalars = [env.stage.value, cp.seat if hasattr(cp, "seat") else -1, env.dealer_pos, int(env.done),
               -1 if env.winner_ix is None else int(env.winner_ix), legal, cyc.idx, cyc.step_counter,
               -1 if cyc.last_raiser is None else cyc.last_raiser, cyc.checkers, cyc.max_steps_total or 0,
               len(env.deck or []), dealer.k, cyc.dealer_idx]
    money = [env.community_pot, env.current_round_pot, env.min_call, float(reward)]
    obs = np.asarray(env.array_everything, dtype=np.float64)           # the observation vector (env.py:266-270)
    return np.array(scalars, dtype=np.int64), np.array(money, dtype=np.float64), seat, cards, tc, obs


def play(game, n_players, stacks, sb, bb, max_raises, max_steps, illegal_rate):
    dealer = Dealer(SEED, game)
    rng = random.Random(1000 + game)
    real_randint = np.random.randint
    np.random.randint = dealer.randint
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            env = HoldemTable(initial_stacks=stacks, small_blind=sb, big_blind=bb, funds_plot=False,
                              max_raises_per_player_round=max_raises)
            for _ in range(n_players):
                env.add_player(Seat())
            env.reset()
            snaps = [snapshot(env, dealer, 0)]
            actions = []
            while not env.done and len(actions) < max_steps:
                if not env.legal_moves:
                    break                                   # stuck in SHOWDOWN at hand start (reference defect)
                if rng.random() < illegal_rate:
                    a = rng.choice([x for x in list(Action)[:8] if x not in env.legal_moves] or env.legal_moves)
                else:                                       # all-ins are rare enough for games to last many hands
                    w = [0.06 if x == Action.ALL_IN else 0.4 if x == Action.RAISE_2POT else 1.0 for x in env.legal_moves]
                    a = rng.choices(env.legal_moves, weights=w)[0]
                _, reward, _, _, _ = env.step(a)
                actions.append(a.value)
                snaps.append(snapshot(env, dealer, reward))
    finally:
        np.random.randint = real_randint
    return actions, snaps


def build_traces():
    configs = []
    g = 0
    for n_players, stacks, sb, bb, max_raises, games in (
            (6, 100, 1, 2, 2, 10), (2, 100, 1, 2, 2, 8), (3, 30, 1, 2, 2, 6), (6, 10, 1, 2, 2, 6),
            (4, 100, 1, 2, 1, 4), (2, 100000, 1, 2, 3, 4), (9, 50, 5, 10, 2, 4), (6, 7, 1, 2, 3, 4), (2, 3, 1, 2, 2, 4)):
        configs.append((n_players, stacks, sb, bb, max_raises, list(range(g, g + games))))
        g += games
    out = {"seed": np.array([SEED], dtype=np.int64)}
    meta = []
    for ci, (n_players, stacks, sb, bb, max_raises, games) in enumerate(configs):
        for game in games:
            actions, snaps = play(game, n_players, stacks, sb, bb, max_raises, max_steps=400,
                                  illegal_rate=0.08 if game % 2 else 0.0)
            out["g%d_actions" % game] = np.array(actions, dtype=np.int8)
            for name, j in (("scalars", 0), ("money", 1), ("seat", 2), ("cards", 3), ("table_cards", 4), ("obs", 5)):
                out["g%d_%s" % (game, name)] = np.stack([s[j] for s in snaps])
            meta.append({"game": game, "config": ci, "n_players": n_players, "initial_stacks": stacks, "small_blind": sb,
                         "big_blind": bb, "max_raises": max_raises, "steps": len(actions),
                         "finished": bool(snaps[-1][0][3])})
            print(" game", game, "players", n_players, "steps", len(actions), "done", bool(snaps[-1][0][3]),
                  "hands", "-", "final stacks", [float(x) for x in snaps[-1][2][0][:n_players]])
    out["meta"] = np.frombuffer(json.dumps({
        "scalars": ["stage", "current_player", "dealer_pos", "done", "winner_ix", "legal_moves", "idx", "step_counter",
                    "last_raiser", "checkers", "max_steps_total", "n_deck", "rng_counter", "dealer_idx"],
        "money": ["community_pot", "current_round_pot", "min_call", "reward"], "seat": FIELDS_SEAT,
        "games": meta}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "env_traces.npz"), **out)
    print("env_traces.npz written:", sum(m["steps"] for m in meta), "steps in", len(meta), "games")


class AutoSeat:
    """Deterministic autoplay agent for the autoplay traces: the k-th call picks legal move (5k + 3) mod #moves among the
    legal moves sorted by value, all-ins excluded unless nothing else is legal (tests/test_holdem.py holds the same class)."""
    autoplay = True

    def __init__(self):
        self.name, self.k = "auto", 0

    def action(self, legal_moves, observation, info, funds_history):
        moves = sorted((m for m in legal_moves if m != Action.ALL_IN), key=lambda m: m.value) or list(legal_moves)
        self.k += 1
        return moves[(5 * self.k + 3) % len(moves)]


def build_autoplay_traces():
    """Games that mix autoplay agents with externally driven seats (gym_env/env.py:170-200): what step() returns -- reward,
    done -- and who is to move, after every call.  Table t of seed SEED + 1000 + game deals from the Philox stream of table 0."""
    games = []
    for game, (pattern, stacks) in enumerate([("EAA", 30), ("AEA", 100), ("AAEA", 20), ("EAEAAA", 50), ("AE", 40), ("EA", 12),
                                              ("AAAE", 100), ("AEAEA", 8)]):
        seed = SEED + 1000 + game
        dealer = Dealer(seed, 0)
        rng = random.Random(2000 + game)
        real_randint = np.random.randint
        np.random.randint = dealer.randint
        calls = []
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                env = HoldemTable(initial_stacks=stacks, small_blind=1, big_blind=2, funds_plot=False)
                for ch in pattern:
                    env.add_player(AutoSeat() if ch == "A" else Seat())
                env.reset()
                while not env.done and len(calls) < 300 and env.legal_moves:
                    if hasattr(env.current_player.agent_obj, "autoplay"):
                        a = None
                        _, reward, done, _, _ = env.step(Action.FOLD)      # ignored: an autoplay seat is to move
                    else:
                        w = [0.05 if x == Action.ALL_IN else 1.0 for x in env.legal_moves]
                        a = rng.choices(env.legal_moves, weights=w)[0]
                        _, reward, done, _, _ = env.step(a)
                    cp = env.current_player
                    calls.append({"action": None if a is None else a.value, "reward": float(reward), "done": bool(done),
                                  "current_player": cp.seat if hasattr(cp, "seat") else -1,
                                  "stacks": [float(p.stack) for p in env.players], "stage": env.stage.value,
                                  "rng_counter": dealer.k})
        finally:
            np.random.randint = real_randint
        games.append({"seed": seed, "pattern": pattern, "initial_stacks": stacks, "calls": calls})
        print(" autoplay game", game, pattern, "calls", len(calls), "done", calls[-1]["done"] if calls else None,
              "nonzero rewards", sum(1 for c in calls if c["reward"]))
    with open(os.path.join(HERE, "autoplay_traces.json"), "w") as f:
        json.dump({"source": "gym_env/env.py::HoldemTable.step with autoplay agents (unmodified reference)", "games": games}, f)
    print("autoplay_traces.json written")


def build_agent_cases():
    from agents.agent_consider_equity import Player
    rng = random.Random(5)
    cases = []
    for _ in range(3000):
        call, bet = round(rng.uniform(0, 1), 2), round(rng.uniform(-1, 1), 2)
        eq = rng.choice([rng.random(), round(rng.random(), 1), bet + 0.1, bet + 0.2, bet, call, bet - 0.1])
        legal = [a for a in list(Action)[:8] if rng.random() < 0.5]
        act = Player(min_call_equity=call, min_bet_equity=bet).action(legal, None, {"player_data": {"equity_to_river_alive": eq}}, None)
        cases.append([call, bet, eq, sum(1 << a.value for a in legal), act.value])
    with open(os.path.join(HERE, "agent_cases.json"), "w") as f:
        json.dump({"source": "agents/agent_consider_equity.py::Player.action", "columns":
                   ["min_call_equity", "min_bet_equity", "equity", "legal_moves_mask", "action"], "cases": cases}, f)
    print("agent_cases.json written")


if __name__ == "__main__":
    if "--autoplay-only" not in sys.argv:
        build_traces()
        build_agent_cases()
    build_autoplay_traces()
