"""The vectorised HoldemTable (include/npk_holdem.h, neuron_poker_b200/holdem.py) against the reference environment.

  * golden traces recorded from the UNMODIFIED reference env (tests/golden/make_env_golden.py): 50 games, 6,244 steps,
    2..9 players, short and deep stacks, 1..3 raises per round, with illegal moves mixed in -- every field of the table
    state must match after every step, money bit for bit;
  * the reference's own environment tests (reference tests/test_gym_env.py), restated line for line on tables of the
    batch API;
  * agents/agent_consider_equity.py decisions on 3,000 recorded cases.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_table_struct_matches_the_header():
    """The numpy mirror of struct NpkHoldemTable has the size the library reports (no GPU needed)."""
    from neuron_poker_b200 import _lib, holdem
    L = _lib.lib()
    L.npk_holdem_table_bytes.restype = __import__("ctypes").c_int64
    assert L.npk_holdem_table_bytes() == holdem.TABLE_DTYPE.itemsize
    header = open(os.path.join(os.path.dirname(HERE), "include", "npk_holdem.h")).read()
    for a in holdem.Action:
        assert "NPK_%s = %d" % (a.name, a.value) in header
    for s in holdem.Stage:
        assert "NPK_%s = %d" % (s.name, s.value) in header


@pytest.fixture(scope="module")
def traces():
    z = np.load(os.path.join(HERE, "golden", "env_traces.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _compare(game, step, st, z, meta, n):
    sc = dict(zip(meta["scalars"], z["g%d_scalars" % game][step]))
    mo = dict(zip(meta["money"], z["g%d_money" % game][step]))
    seat = z["g%d_seat" % game][step]
    where = "game %d step %d" % (game, step)
    for k in ("stage", "dealer_pos", "done", "winner_ix", "legal_moves", "idx", "step_counter", "last_raiser", "checkers",
              "max_steps_total", "n_deck", "rng_counter", "dealer_idx"):
        if sc["done"] and k in ("legal_moves", "n_deck"):
            continue
        assert int(st[k]) == int(sc[k]), (where, k, int(st[k]), int(sc[k]))
    # once the game is over the reference shows the winner as "current player" (env.py:246-247)
    assert int(st["current_player"]) == int(sc["current_player"]), (where, "current_player")
    for k in ("community_pot", "current_round_pot", "min_call", "reward"):
        assert float(st[k]) == float(mo[k]), (where, k, float(st[k]), float(mo[k]))
    for j, k in enumerate(meta["seat"]):
        assert (np.asarray(st[k][:n], dtype=np.float64) == seat[j][:n]).all(), (where, k, st[k][:n], seat[j][:n])
    assert (st["cards"][:n] == z["g%d_cards" % game][step][:n]).all(), (where, "cards")
    assert (st["table_cards"] == z["g%d_table_cards" % game][step]).all(), (where, "table_cards")


def _compare_obs(game, step, got, z, n, done):
    """The observation vector array_everything (env.py:266-270), bit for bit (nan == nan).  Once the game is over the
    reference's legal-move flags are the stale ones of the finished step (env.py:233-234 skips the refresh): skipped."""
    want = z["g%d_obs" % game][step]
    assert got.shape == want.shape == (22 + 51 * n,)
    keep = np.ones(len(want), dtype=bool)
    if done:
        lm = 4 + n + n + 4 + 2 + n + 2
        keep[lm:lm + 10] = False
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same[keep].all(), ("game %d step %d" % (game, step), np.nonzero(~same & keep)[0][:8], got[~same & keep][:8],
                              want[~same & keep][:8])


@pytest.mark.gpu
def test_tables_replay_the_reference_traces(cuda_device, traces):
    """Every game of env_traces.npz, all games of one configuration stepped together as one batch: the table state and
    the observation vector after reset and after every step."""
    import torch
    from neuron_poker_b200.holdem import HoldemTables
    z, meta = traces
    seed = int(z["seed"][0])
    games = meta["games"]
    total = 0
    for ci in sorted({g["config"] for g in games}):
        group = [g for g in games if g["config"] == ci]
        g0 = group[0]
        assert [g["game"] for g in group] == list(range(g0["game"], g0["game"] + len(group)))
        tb = HoldemTables(len(group), n_players=g0["n_players"], initial_stacks=g0["initial_stacks"],
                          small_blind=g0["small_blind"], big_blind=g0["big_blind"],
                          max_raises_per_player_round=g0["max_raises"], seed=seed, table_offset=g0["game"])
        tb.enable_observations()
        half = torch.full((len(group),), 0.5, dtype=torch.float64, device=tb.device)     # the traces' stub equity
        st = tb.state()
        obs = tb.observe(half).cpu().numpy()
        for i, g in enumerate(group):
            _compare(g["game"], 0, st[i], z, meta, g["n_players"])
            _compare_obs(g["game"], 0, obs[i], z, g["n_players"], False)
        for step in range(max(g["steps"] for g in group)):
            acts = np.array([z["g%d_actions" % g["game"]][step] if step < g["steps"] else -1 for g in group], dtype=np.int8)
            tb.step(acts)
            st = tb.state()
            obs = tb.observe(half).cpu().numpy()
            for i, g in enumerate(group):
                if step < g["steps"]:
                    assert st[i]["error"] == 0
                    _compare(g["game"], step + 1, st[i], z, meta, g["n_players"])
                    _compare_obs(g["game"], step + 1, obs[i], z, g["n_players"], bool(st[i]["done"]))
                    total += 1
    assert total == sum(g["steps"] for g in games) >= 6000


# ---- the reference's tests/test_gym_env.py on a one-table batch ---------------------------------------------------------------
class _Env(object):
    """One table of a batch with the attribute names the reference tests read."""

    def __init__(self, n_players, initial_stacks=100, small_blind=1, big_blind=2, max_raises_per_player_round=2):
        from neuron_poker_b200.holdem import HoldemTables
        self.tb = HoldemTables(1, n_players=n_players, initial_stacks=initial_stacks, small_blind=small_blind,
                               big_blind=big_blind, max_raises_per_player_round=max_raises_per_player_round, seed=5)

    def step(self, action):
        self.tb.step([int(action)])

    @property
    def s(self):
        return self.tb.state()[0]

    stage = property(lambda self: self.s["stage"])
    seat = property(lambda self: int(self.s["current_player"]))
    legal_moves = property(lambda self: self.tb.legal_moves_list(self.s["legal_moves"]))

    def stack(self, i):
        return float(self.s["stack"][i])

    def set_stack(self, i, v):
        import torch
        st = self.tb.state().copy()
        st[0]["stack"][i] = v
        self.tb.buf.copy_(torch.from_numpy(st.view(np.uint8)))


@pytest.mark.gpu
def test_reference_env_tests(cuda_device):
    from neuron_poker_b200.holdem import Action, Stage
    # test_basic_actions_with_6_player (reference tests/test_gym_env.py:26-52)
    env = _Env(6)
    assert (env.s["cards"][0] < 52).all() and env.seat == 3
    for a in (Action.CALL, Action.FOLD, Action.FOLD, Action.FOLD, Action.CALL):
        env.step(a)
    assert env.seat == 2
    assert [env.stack(i) for i in (3, 4, 5, 0, 1, 2)] == [98, 100, 100, 100, 98, 98]
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_POT)
    assert env.s["cycle_round_number"]
    env.step(Action.FOLD)
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP
    env.step(Action.CHECK)

    # test_no_player_raise_big_blind_do_last_action_in_round (:55-71): the blind "actions" are illegal moves
    env = _Env(2)
    env.step(Action.SMALL_BLIND); env.step(Action.BIG_BLIND); env.step(Action.CALL)
    assert env.stage == Stage.PREFLOP
    env.step(Action.CHECK)
    assert env.stage == Stage.FLOP

    # test_one_player_raise3bb_one_call_this_call_is_last_action_in_round (:74-88)
    env = _Env(2)
    env.step(Action.SMALL_BLIND); env.step(Action.BIG_BLIND); env.step(Action.RAISE_3BB)
    assert env.stage == Stage.PREFLOP
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP

    # test_raise_to_3_times_big_blind_after_big_blind_bet (:91-101)
    env = _Env(2)
    assert env.s["player_pots"][0] == 2
    env.step(Action.CALL); env.step(Action.RAISE_3BB)
    assert env.s["player_pots"][0] == 6

    # test_raise_to_3_times_big_blind_is_not_possible_with_not_enough_remaining_stack (:104-110)
    env = _Env(4, initial_stacks=2)
    env.step(Action.CALL)
    assert Action.RAISE_3BB not in env.legal_moves

    # test_raise_to_3_times_big_blind_is_possible_with_enough_remaining_stack (:113-125)
    env = _Env(2)
    env.set_stack(0, 4)
    env.step(Action.CALL)
    assert Action.RAISE_3BB in env.legal_moves
    env.step(Action.RAISE_3BB)
    assert env.stack(0) == 0

    # test_base_actions_6_players_check_legal_moves_and_stages (:163-193)
    env = _Env(6)
    for _ in range(4):
        env.step(Action.CALL)
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_HALF_POT)
    assert len(env.legal_moves) > 2 and env.stage == Stage.PREFLOP
    env.step(Action.RAISE_HALF_POT)
    assert env.stage == Stage.PREFLOP and len(env.legal_moves) > 2
    for _ in range(5):
        env.step(Action.CALL)
    assert env.stage == Stage.FLOP and env.seat == 1
    env.step(Action.CHECK)
    assert env.stage == Stage.FLOP and env.seat == 2
    env.step(Action.RAISE_HALF_POT); env.step(Action.FOLD)

    # test_unlimited_raising_preflop (:258-270)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=3)
    env.step(Action.CALL); env.step(Action.RAISE_POT); env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP

    # test_end_preflop_on_call (:273-282)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=3)
    env.step(Action.CALL); env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP

    # test_preflop_call_after_max_raises (:285-314)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=2)
    env.step(Action.CALL); env.step(Action.RAISE_POT); env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    env.step(Action.RAISE_POT)
    assert env.stage == Stage.PREFLOP
    assert env.legal_moves == [Action.FOLD, Action.CALL]          # the reference lists [CALL, FOLD]
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP
    env.step(Action.RAISE_POT); env.step(Action.CALL)
    assert env.stage == Stage.TURN
    env.step(Action.RAISE_POT); env.step(Action.CALL)
    assert env.stage == Stage.RIVER
    env.step(Action.RAISE_POT); env.step(Action.CALL)

    # test_one_max_raise_per_player (:341-345)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=1)
    assert env.stage == Stage.PREFLOP

    # test_headsup_bb_starts_flop_bb_ends_preflop (:348-360)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=2)
    assert env.stage == Stage.PREFLOP and env.seat == 1
    env.step(Action.CALL)
    assert env.stage == Stage.PREFLOP and env.seat == 0
    env.step(Action.CHECK)
    assert env.stage == Stage.FLOP and env.seat == 0

    # test_headsup_bb_starts_flop_sb_ends_preflop (:363-377)
    env = _Env(2, initial_stacks=100000, max_raises_per_player_round=2)
    assert env.stage == Stage.PREFLOP and env.seat == 1
    env.step(Action.CALL)
    assert env.stage == Stage.PREFLOP and env.seat == 0
    env.step(Action.RAISE_3BB)
    assert env.stage == Stage.PREFLOP and env.seat == 1
    env.step(Action.CALL)
    assert env.stage == Stage.FLOP and env.seat == 0


@pytest.mark.gpu
def test_equity_agent_decisions_match_the_reference(cuda_device):
    """agents/agent_consider_equity.py::Player.action on 3,000 recorded (thresholds, equity, legal moves) cases: the
    decide kernel reads the legal moves from the table state, so each case is planted into a table record."""
    import torch
    from neuron_poker_b200.holdem import EquityAgents, HoldemTables
    with open(os.path.join(HERE, "golden", "agent_cases.json")) as f:
        cases = json.load(f)["cases"]
    n = len(cases)
    tb = HoldemTables(n, n_players=2, seed=1)
    st = tb.state().copy()
    st["legal_moves"] = np.array([c[3] for c in cases], dtype=np.uint32)
    st["current_player"] = 0
    tb.buf.copy_(torch.from_numpy(st.view(np.uint8)))
    eq = torch.tensor([c[2] for c in cases], dtype=torch.float64, device=tb.device)
    got = np.full(n, -9, dtype=np.int64)
    # thresholds are per seat, not per table: group the cases by (min_call, min_bet)
    thr = np.array([(c[0], c[1]) for c in cases])
    for call, bet in sorted({(c[0], c[1]) for c in cases}):
        acts = tb.decide(EquityAgents(2).equity(0, call, bet), equity=eq).cpu().numpy()
        sel = (thr[:, 0] == call) & (thr[:, 1] == bet)
        got[sel] = acts[sel]
    want = np.array([c[4] if c[3] else -1 for c in cases])
    assert (got == want).all(), np.nonzero(got != want)[0][:10]


@pytest.mark.gpu
def test_selfplay_conserves_chips_and_finishes_games(cuda_device):
    """4,096 six-max tables of equity and random agents (main.py:136-150) playing 300 actions each with real Monte-Carlo
    equity: no table reports an error, chips are conserved on every table, games end and restart."""
    from neuron_poker_b200.holdem import EquityAgents, HoldemTables
    tb = HoldemTables(4096, n_players=6, seed=11, autoplay=[1] * 6)
    agents = EquityAgents.equity_vs_random()
    finished = 0
    for _ in range(300):
        tb.selfplay_step(agents, runs=200, deal_mode="uniform", restart_finished=False)
        st = tb.state()
        assert (st["error"] == 0).all()
        live = st["done"] == 0
        chips = st["stack"][:, :6].sum(1) + st["player_max_win"][:, :6].sum(1)
        assert np.allclose(chips[live], 600.0), chips[live][~np.isclose(chips[live], 600.0)][:5]
        finished += int((st["done"] != 0).sum())
        tb.reset_done()
    assert finished > 0
    assert (tb.state()["hands_played"] > 0).all()


@pytest.mark.gpu
def test_single_table_front_end_runs_the_reference_agent_protocol(cuda_device):
    """holdem.HoldemTable: the reference's env front end for one table.  Agents written against the reference's protocol
    (an `autoplay` attribute and action(legal_moves, observation, info, funds_history), agents/agent_consider_equity.py)
    drive a whole game; a seat without `autoplay` is driven through step(action) like the reference's tests do."""
    from neuron_poker_b200.holdem import Action, HoldemTable, Stage

    class EquityPlayer:                                   # agents/agent_consider_equity.py:21-58, same decisions
        def __init__(self, name, min_call_equity, min_bet_equity):
            self.name, self.min_call_equity, self.min_bet_equity, self.autoplay = name, min_call_equity, min_bet_equity, True
            self.seen = 0

        def action(self, action_space, observation, info, funds_history):
            eq = info["player_data"]["equity_to_river_alive"]
            assert 0.0 <= eq <= 1.0 and len(observation) == 22 + 51 * 3
            self.seen += 1
            if eq > self.min_bet_equity + .2 and Action.ALL_IN in action_space:
                return Action.ALL_IN
            if eq > self.min_bet_equity + .1 and Action.RAISE_2POT in action_space:
                return Action.RAISE_2POT
            if eq > self.min_bet_equity and Action.RAISE_POT in action_space:
                return Action.RAISE_POT
            if eq > self.min_bet_equity - .1 and Action.RAISE_HALF_POT in action_space:
                return Action.RAISE_HALF_POT
            if eq > self.min_call_equity and Action.CALL in action_space:
                return Action.CALL
            return Action.CHECK if Action.CHECK in action_space else Action.FOLD

    env = HoldemTable(initial_stacks=20, seed=3)
    agents = [EquityPlayer("a", .5, .7), EquityPlayer("b", .2, .9), EquityPlayer("c", .4, .6)]
    for a in agents:
        env.add_player(a)
    env.reset()                                            # all seats autoplay: reset() plays the whole game (env.py:165-166)
    assert env.done and sum(a.seen for a in agents) > 3
    assert abs(sum(p.stack for p in env.players) - 60) < 1e-9 and len(env.funds_history) >= 2

    class Human:                                          # reference tests/test_gym_env.py::PlayerForTest
        name = "h"

    env = HoldemTable(seed=4)
    for _ in range(6):
        env.add_player(Human())
    obs, info = env.reset()
    assert env.current_player.seat == 3 and env.stage == Stage.PREFLOP and len(obs) == 328
    obs, reward, done, truncated, info = env.step(Action.CALL)
    assert env.current_player.seat == 4 and env.players[3].stack == 98 and not done and truncated is False
    assert 0.0 <= info["player_data"]["equity_to_river_alive"] <= 1.0
    env.step(Action.CHECK)                                # illegal here: costs -1, nothing else happens
    assert env.reward == -1 and env.current_player.seat == 4


@pytest.mark.gpu
def test_autoplay_rewards_follow_the_reference(cuda_device):
    """tests/golden/autoplay_traces.json: eight games of the UNMODIFIED reference env mixing autoplay agents and externally
    driven seats (make_env_golden.py::build_autoplay_traces).  holdem.HoldemTable returns the same reward and done flag from every
    step() call, hands the move to the same seat and ends with the same stacks -- in particular the reward an autoplay sequence
    reports for the seat that was to move when step() was entered (gym_env/env.py:178-188, :282-306)."""
    import json
    from neuron_poker_b200.holdem import Action, HoldemTable

    class AutoSeat:                                        # the same deterministic agent as in make_env_golden.py
        autoplay = True

        def __init__(self):
            self.name, self.k = "auto", 0

        def action(self, legal_moves, observation, info, funds_history):
            moves = sorted((m for m in legal_moves if m != Action.ALL_IN), key=lambda m: m.value) or list(legal_moves)
            self.k += 1
            return moves[(5 * self.k + 3) % len(moves)]

    class Human:
        name = "h"

    with open(os.path.join(os.path.dirname(__file__), "golden", "autoplay_traces.json")) as f:
        games = json.load(f)["games"]
    checked = 0
    for g in games:
        env = HoldemTable(initial_stacks=g["initial_stacks"], small_blind=1, big_blind=2, seed=g["seed"])
        env.get_equity = lambda *a: 0.5                    # the traces were recorded with this constant
        for ch in g["pattern"]:
            env.add_player(AutoSeat() if ch == "A" else Human())
        env.reset()
        for i, c in enumerate(g["calls"]):
            a = Action.FOLD if c["action"] is None else Action(c["action"])
            _, reward, done, _, _ = env.step(a)
            cp = env.current_player.seat if env.current_player is not None else -1
            where = (g["pattern"], i, c)
            assert reward == c["reward"] and done == c["done"], (where, reward, done)
            assert [p.stack for p in env.players] == c["stacks"], (where, [p.stack for p in env.players])
            if not done:
                assert cp == c["current_player"] and env.stage.value == c["stage"], (where, cp, env.stage)
            checked += 1
    assert checked > 250
