"""Opponent ranges, hero ranges and ghost cards (reference tools/montecarlo_python.py:24-112, :136-181,
:206-208; reference tests test_montecarlo20/21, tests/test_montecarlo_python.py:215-232).

CPU part: the oracle's range-aware restatement against seeded runs of the unmodified reference, and the host-side range
tables against the reference's own ranking.  GPU part (-m gpu): libnpk's range kernels against the sampler specification
(bit-exact) and against the oracle (3 sigma)."""
import numpy as np
import pytest

import oracle
import sampler_model
from neuron_poker_b200 import ranges


def _opp_classes(run):
    r = run["opponent_range"]
    return set(r) if isinstance(r, list) else ranges.allowed_classes(r)


def test_preflop_order_is_the_references(golden_preflop):
    assert ranges.PREFLOP_ORDER == golden_preflop["order"]
    assert len({ranges.class_index(n) for n in ranges.PREFLOP_ORDER}) == 169
    for x, names in golden_preflop["allowed"].items():
        assert sorted(ranges.allowed_classes(float(x))) == names, x
    assert len(ranges.allowed_classes(0.004)) == 169          # take_top == 0 -> [-0:] is the whole list


def test_class_numbering_agrees_everywhere():
    for c1 in range(52):
        for c2 in range(52):
            if c1 != c2:
                k = oracle.hand_class(c1, c2)
                assert k == ranges.class_of_cards(c1, c2) == sampler_model.hand_class(c1, c2)
                name = oracle.RANKS[c1 >> 2] + oracle.RANKS[c2 >> 2] + ("" if c1 >> 2 == c2 >> 2 else
                                                                        "S" if (c1 & 3) == (c2 & 3) else "O")
                assert ranges.class_index(name) == k == oracle.class_index(name)
    assert ranges.class_index("AAO") is None and ranges.class_index("AK") is None and ranges.class_index("1KS") is None


def test_oracle_reproduces_reference_range_runs(golden_ranges):
    """wins, passes, win types and the RNG stream position, bit for bit, under np.random.seed(s)."""
    assert len(golden_ranges["runs"]) == 16
    for run in golden_ranges["runs"]:
        hero_classes = set(run["hero"]) if run["hero_is_range"] else None
        got = oracle.mc_reference_ranges(None if run["hero_is_range"] else run["hero"], run["board"], run["players"],
                                         run["runs"], run["seed"], _opp_classes(run), hero_classes=hero_classes,
                                         ghost=run["ghost"] or None)
        assert got["wins"] == run["wins"], run["name"]
        assert got["passes"] == run["passes"], run["name"]
        assert got["win_types"] == run["win_types"], run["name"]
        assert got["next_randint"] == run["next_randint_0_1000000"], run["name"]


def test_pair_list_specification_deals_the_reference_distribution_exactly():
    """The pair-list sampler (sampler_model.deal_ranges_fast) against the reference's literal loop, by exact enumeration on one
    opponent draw: for every state (set of unseen cards) tried, the accepted (entry, orientation) outcomes map one-to-one onto
    the (i1, i2) the reference accepts, and deal the same two cards."""
    import itertools
    import random
    rnd = random.Random(5)
    mask = ranges.opponent_mask(0.3)
    for _ in range(6):
        deck = sorted(rnd.sample(range(52), rnd.randint(12, 20)))
        n = len(deck)
        # the reference: i1 in [0,n), i2 in [0,n-1), i1 != i2, class(deck[i1], deck[i2]) allowed; pop(i1), pop(i2)
        ref = []
        for i1, i2 in itertools.product(range(n), range(n - 1)):
            if i1 != i2 and sampler_model._allowed(mask, deck[i1], deck[i2]):
                d = list(deck)
                ref.append((d.pop(i1), d.pop(i2)))
        # the pair list: unordered allowed pairs, both orientations, sb != max, successor rule
        fast = []
        pairs = [(a, b) for b in deck for a in deck if a < b and sampler_model._allowed(mask, a, b)]
        for a, b in pairs:
            for sa, sb in ((a, b), (b, a)):
                if sb == max(deck):
                    continue
                fast.append((sa, min(c for c in deck if c > sb) if sb > sa else sb))
        assert sorted(ref) == sorted(fast) and len(ref) > 0


def test_sampler_specification_with_ranges_converges_to_the_oracle():
    """The Philox-driven specification of libnpk's range dealers (tests/sampler_model.py) and the oracle's MT19937-driven
    restatement sample the same distributions: equities agree within 3 sigma of their combined standard error."""
    cases = [("reference", ["KS", "KC"], None, ["3D", "9H", "AS", "7S", "QH"], 3, ranges.allowed_classes(0.25), None),
             ("reference", None, {"AKO", "AA"}, ["3D", "9H", "AS", "7S", "QH"], 3, ranges.allowed_classes(0.25), None),
             ("uniform", ["AS", "KS"], None, ["2C", "7D", "KH"], 3, ranges.allowed_classes(0.5), ["AH", "AD"]),
             ("uniform", None, {"QQ", "AKS", "78S"}, [], 2, ranges.allowed_classes(0.3), None)]
    T = 6000
    for mode, hero, hero_cls, board, players, opp, ghost in cases:
        h = [oracle.card_id(c) for c in hero] if hero else None
        b = [oracle.card_id(c) for c in board]
        g = [oracle.card_id(c) for c in ghost] if ghost else None
        m = sampler_model.run_model(oracle, mode, 99, 0, h, b, players, T, opp_mask=ranges.mask_from_classes(opp),
                                    hero_mask=ranges.mask_from_classes(hero_cls) if hero_cls else None, ghost=g)
        p_model = (m["wins"] + m["ties"]) / T
        if mode == "reference":
            o = oracle.mc_reference_ranges(hero, board, players, 40000, 5, opp, hero_classes=hero_cls, ghost=ghost)
            p_or = o["wins"] / 40000
            passes_or = o["passes"] / 40000
            assert abs(m["passes"] / T - passes_or) < 0.05 * passes_or + 0.05
        else:
            w, t, _ = oracle.mc_uniform_ranges(hero, board, players, 40000, 5, opp, hero_classes=hero_cls, ghost=ghost)
            p_or = (w + t) / 40000
        se = (p_or * (1 - p_or) * (1 / T + 1 / 40000)) ** 0.5
        assert abs(p_model - p_or) < 3 * se + 1e-9, (mode, hero, hero_cls, p_model, p_or, se)
        f = sampler_model.run_model(oracle, mode, 199, 0, h, b, players, T, opp_mask=ranges.mask_from_classes(opp),
                                    hero_mask=ranges.mask_from_classes(hero_cls) if hero_cls else None, ghost=g, fast=True)
        p_fast = (f["wins"] + f["ties"]) / T
        assert abs(p_fast - p_or) < 3 * se + 1e-9, ("pair list", mode, hero, hero_cls, p_fast, p_or, se)


# ---- GPU: libnpk's range kernels ---------------------------------------------------------------------------------------
def _ids(cards):
    return [oracle.card_id(c) for c in cards]


RANGE_CASES = [
    # mode, hero cards, hero classes, board, players, opponent range, ghost
    ("reference", ["KS", "KC"], None, ["3D", "9H", "AS", "7S", "QH"], 3, 0.25, None),
    ("reference", None, {"AKO", "AA"}, ["3D", "9H", "AS", "7S", "QH"], 3, 0.25, None),
    ("reference", ["AS", "KS"], None, ["2C", "7D", "KH"], 4, 0.5, ["AH", "AD"]),
    ("reference", ["QS", "QH"], None, [], 2, 0.1, None),
    ("reference", ["9D", "9C"], None, ["2C", "7D", "KH", "TS"], 6, {"AKS", "KAO", "QQ", "JJ", "9TS", "T9O", "23O"}, None),
    ("uniform", ["KS", "KC"], None, ["3D", "9H", "AS", "7S", "QH"], 3, 0.25, None),
    ("uniform", None, {"QQ", "AKS", "78S"}, [], 2, 0.3, None),
    ("uniform", ["AS", "KS"], None, ["2C", "7D", "KH"], 5, 0.5, ["AH", "AD"]),
    ("uniform", ["5H", "5D"], None, ["5C", "KD", "2S", "2D"], 10, 1, None),
]


@pytest.mark.gpu
def test_range_kernels_match_the_sampler_specification(cuda_device):
    """wins, ties, passes and win types of equity_ranges_kernel equal the Python specification scored by the oracle,
    bit for bit, in both dealing modes, with hero ranges and ghost cards, at a trial offset."""
    import neuron_poker_b200 as npk
    T, off = 96, 1000
    for i, (mode, hero, hero_cls, board, players, opp, ghost) in enumerate(RANGE_CASES):
        b = _ids(board)
        out = npk.get_equity_ranges_batch(
            None if hero is None else np.array([_ids(hero)], dtype=np.uint8),
            np.array([b + [255] * (5 - len(b))], dtype=np.uint8), np.array([players], dtype=np.uint8), T,
            opponent_range=opp, hero_range=hero_cls, ghost=None if ghost is None else np.array([_ids(ghost)], dtype=np.uint8),
            seed_value=77 + i, deal_mode=mode, trial_offset=off, query_offset=5, win_types=True, passes=True)
        m = sampler_model.run_model(oracle, mode, 77 + i, 5, _ids(hero) if hero else None, b, players, T, trial_offset=off,
                                    opp_mask=ranges.opponent_mask(opp),
                                    hero_mask=ranges.mask_from_classes(hero_cls) if hero_cls else None,
                                    ghost=_ids(ghost) if ghost else None)
        got = (int(out["wins"][0]), int(out["ties"][0]), int(out["passes"][0]), [int(x) for x in out["win_types"][0]])
        assert got == (m["wins"], m["ties"], m["passes"], m["win_types"]), (mode, hero, hero_cls, got, m)


@pytest.mark.gpu
def test_pair_list_range_kernel_matches_its_specification(cuda_device):
    """Without `passes` the range entry point runs equity_ranges_fast_kernel (pair-list sampler, csrc/npk_ranges.cu): wins, ties
    and win types equal its Python specification (sampler_model.deal_ranges_fast) bit for bit in both dealing modes, with hero
    ranges and ghost cards, at a trial offset."""
    import neuron_poker_b200 as npk
    T, off = 96, 1000
    for i, (mode, hero, hero_cls, board, players, opp, ghost) in enumerate(RANGE_CASES):
        b = _ids(board)
        out = npk.get_equity_ranges_batch(
            None if hero is None else np.array([_ids(hero)], dtype=np.uint8),
            np.array([b + [255] * (5 - len(b))], dtype=np.uint8), np.array([players], dtype=np.uint8), T,
            opponent_range=opp, hero_range=hero_cls, ghost=None if ghost is None else np.array([_ids(ghost)], dtype=np.uint8),
            seed_value=177 + i, deal_mode=mode, trial_offset=off, query_offset=5, win_types=True, passes=False)
        m = sampler_model.run_model(oracle, mode, 177 + i, 5, _ids(hero) if hero else None, b, players, T, trial_offset=off,
                                    opp_mask=ranges.opponent_mask(opp),
                                    hero_mask=ranges.mask_from_classes(hero_cls) if hero_cls else None,
                                    ghost=_ids(ghost) if ghost else None, fast=True)
        got = (int(out["wins"][0]), int(out["ties"][0]), [int(x) for x in out["win_types"][0]])
        assert got == (m["wins"], m["ties"], m["win_types"]), (mode, hero, hero_cls, got, m)


@pytest.mark.gpu
def test_pair_list_range_kernel_within_3_sigma_of_the_oracle_and_of_the_generic_kernel(cuda_device):
    """The pair-list sampler deals the reference's distribution: 4 M trials against 400 k oracle trials (the MT19937-driven
    restatement of the reference's loop) and against the generic kernel that plays that loop literally."""
    import neuron_poker_b200 as npk
    T, TO = 4_000_000, 400_000
    for i, (mode, hero, hero_cls, board, players, opp, ghost) in enumerate(RANGE_CASES[:8]):
        b = _ids(board)
        args = (None if hero is None else np.array([_ids(hero)], dtype=np.uint8),
                np.array([b + [255] * (5 - len(b))], dtype=np.uint8), np.array([players], dtype=np.uint8), T)
        kw = dict(opponent_range=opp, hero_range=hero_cls, ghost=None if ghost is None else np.array([_ids(ghost)], dtype=np.uint8),
                  deal_mode=mode)
        fast = npk.get_equity_ranges_batch(*args, seed_value=4321 + i, passes=False, **kw)
        slow = npk.get_equity_ranges_batch(*args, seed_value=8765 + i, passes=True, **kw)
        p_fast = (int(fast["wins"][0]) + int(fast["ties"][0])) / T
        p_slow = (int(slow["wins"][0]) + int(slow["ties"][0])) / T
        opp_cls = opp if isinstance(opp, set) else ranges.allowed_classes(opp)
        if mode == "reference":
            o = oracle.mc_reference_ranges(hero, board, players, TO, 21 + i, opp_cls, hero_classes=hero_cls, ghost=ghost)
            p_or = o["wins"] / TO
        else:
            w, t, a = oracle.mc_uniform_ranges(hero, board, players, TO, 21 + i, opp_cls, hero_classes=hero_cls, ghost=ghost)
            p_or = (w + t) / TO
        se = (p_or * (1 - p_or) * (1 / T + 1 / TO)) ** 0.5
        assert abs(p_fast - p_or) < 3 * se + 1e-9, (mode, hero, hero_cls, p_fast, p_or, se)
        se2 = (max(p_slow * (1 - p_slow), 1e-9) * 2 / T) ** 0.5
        assert abs(p_fast - p_slow) < 3.5 * se2 + 1e-9, (mode, hero, hero_cls, p_fast, p_slow, se2)


@pytest.mark.gpu
def test_full_range_reference_mode_agrees_with_the_plain_reference_dealer(cuda_device):
    """With every class allowed, a fixed hero and no ghost cards the range kernel (index-based dealer on the ordered
    list) and equity_refdeal_kernel (shuffle + rejection of the two excluded outcomes) sample the same distribution:
    equities within 3 sigma of each other on every query of a mixed batch, mean `passes` per trial equal to 1 %."""
    import neuron_poker_b200 as npk
    rng = np.random.default_rng(3)
    Q, T = 24, 1_000_000
    cards = np.stack([rng.permutation(52)[:7] for _ in range(Q)]).astype(np.uint8)
    hole, board = cards[:, :2].copy(), cards[:, 2:7].copy()
    npl = rng.integers(2, 11, Q).astype(np.uint8)
    for q in range(Q):
        board[q, [0, 3, 4, 5][q % 4]:] = 255
    a = npk.get_equity_batch(hole, board, npl, T, seed_value=9, deal_mode="reference", passes=True)
    b = npk.get_equity_ranges_batch(hole, board, npl, T, opponent_range=1, seed_value=10, deal_mode="reference", passes=True)
    pa = ((a["wins"] + a["ties"]).double() / T).cpu().numpy()
    pb = ((b["wins"] + b["ties"]).double() / T).cpu().numpy()
    se = np.sqrt(np.maximum(pa * (1 - pa), 1e-9) * 2 / T)
    assert (np.abs(pa - pb) < 3.5 * se + 1e-9).all(), np.max(np.abs(pa - pb) / se)
    ra, rb = a["passes"].double().cpu().numpy(), b["passes"].double().cpu().numpy()
    assert (np.abs(ra - rb) < 0.01 * rb).all()


@pytest.mark.gpu
def test_range_kernels_within_3_sigma_of_the_oracle(cuda_device):
    """4 M GPU trials against 400 k oracle trials (MT19937-driven restatement of the reference dealer with ranges, resp.
    its unbiased counterpart): |p_gpu - p_oracle| < 3 * combined standard error; mean attempts per trial agree."""
    import neuron_poker_b200 as npk
    T, TO = 4_000_000, 400_000
    for i, (mode, hero, hero_cls, board, players, opp, ghost) in enumerate(RANGE_CASES[:8]):
        b = _ids(board)
        out = npk.get_equity_ranges_batch(
            None if hero is None else np.array([_ids(hero)], dtype=np.uint8),
            np.array([b + [255] * (5 - len(b))], dtype=np.uint8), np.array([players], dtype=np.uint8), T,
            opponent_range=opp, hero_range=hero_cls, ghost=None if ghost is None else np.array([_ids(ghost)], dtype=np.uint8),
            seed_value=1234 + i, deal_mode=mode, passes=True)
        p_gpu = (int(out["wins"][0]) + int(out["ties"][0])) / T
        opp_cls = opp if isinstance(opp, set) else ranges.allowed_classes(opp)
        if mode == "reference":
            o = oracle.mc_reference_ranges(hero, board, players, TO, 11 + i, opp_cls, hero_classes=hero_cls, ghost=ghost)
            p_or, att = o["wins"] / TO, o["passes"] / TO
        else:
            w, t, a = oracle.mc_uniform_ranges(hero, board, players, TO, 11 + i, opp_cls, hero_classes=hero_cls, ghost=ghost)
            p_or, att = (w + t) / TO, a / TO
        se = (p_or * (1 - p_or) * (1 / T + 1 / TO)) ** 0.5
        assert abs(p_gpu - p_or) < 3 * se + 1e-9, (mode, hero, hero_cls, p_gpu, p_or, se)
        assert abs(int(out["passes"][0]) / T - att) < 0.02 * att, (mode, int(out["passes"][0]) / T, att)


@pytest.mark.gpu
def test_reference_range_tests_20_and_21(cuda_device):
    """reference tests/test_montecarlo_python.py:215-232 through the MonteCarlo mirror, with the reference's own
    acceptance rule (|mean of 5 runs - expected| < 3 points, stdev < 3, win types sum to the equity)."""
    import neuron_poker_b200 as npk
    npk.seed(2024)
    sim = npk.MonteCarlo()
    for my_cards, expected in (([['KS', 'KC']], 12.8), ([{'AKO', 'AA'}], 77.8)):
        res = []
        for _ in range(5):
            sim.run_montecarlo(my_cards, ['3D', '9H', 'AS', '7S', 'QH'], 3, 1, maxRuns=15000, timeout=0, ghost_cards='',
                               opponent_range=0.25)
            res.append(sim.equity * 100)
            assert abs(sum(sim.winnerCardTypeList.values()) - sim.equity) < 1e-4
            assert sim.runs == 15000 and sim.passes >= 2 * 15000
        assert abs(np.mean(res) - expected) < 3 and np.std(res) < 3, (my_cards, res)
    npk.seed(None)


@pytest.mark.gpu
def test_unsatisfiable_range_is_an_error_not_a_hang(cuda_device):
    import neuron_poker_b200 as npk
    from neuron_poker_b200._lib import NpkError
    with pytest.raises(NpkError) as ei:        # no ace is left for an opponent who must hold AA
        npk.equity_counts_ranges(['AS', 'AH'], ['AD', 'AC', '2S'], 2, 5000, opponent_range={'AA'})
    assert ei.value.code == -6
    with pytest.raises(NpkError) as ei:        # spellings the reference's test can never produce
        npk.equity_counts_ranges(['AS', 'AH'], [], 2, 100, opponent_range={'AAO', 'AK'})
    assert ei.value.code == -6
    with pytest.raises(NpkError) as ei:        # ghost card that is also on the board: list.index fails in the reference
        npk.equity_counts_ranges(['AS', 'AH'], ['2C', '7D', 'KH'], 2, 100, opponent_range=1, ghost_cards=['KH', '3S'])
    assert ei.value.code == -5


# ---- several known hands in player_card_list (montecarlo_python.py:132-163) ---------------------------------------------------
def _golden_known():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "mc_known_seeded.json")) as f:
        return json.load(f)


def test_oracle_reproduces_reference_runs_with_known_hands():
    """player_card_list = [hero, known hands...]: wins, passes, win types and the RNG stream position of ten seeded runs of the
    unmodified reference (tests/golden/make_golden.py --only known), bit for bit."""
    runs = _golden_known()["runs"]
    assert len(runs) == 10
    for run in runs:
        got = oracle.mc_reference_ranges(run["hero"], run["board"], run["players"], run["runs"], run["seed"], _opp_classes(run),
                                         ghost=run["ghost"] or None, known=run["known"])
        assert got["wins"] == run["wins"], run["name"]
        assert got["passes"] == run["passes"], run["name"]
        assert got["win_types"] == run["win_types"], run["name"]
        assert got["next_randint"] == run["next_randint_0_1000000"], run["name"]


def test_sampler_specifications_with_known_hands_converge_to_the_oracle():
    """Both range specifications (literal attempt loop and pair list) with known opponent hands agree with the oracle's
    restatement of the reference at 3.5 sigma on 6,000 trials each."""
    T = 6000
    for i, run in enumerate(_golden_known()["runs"][::2]):
        hero, board = _ids(run["hero"]), _ids(run["board"])
        known = [_ids(h) for h in run["known"]]
        ghost = _ids(run["ghost"]) if run["ghost"] else None
        cls = _opp_classes(run)
        o = oracle.mc_reference_ranges(run["hero"], run["board"], run["players"], 60000, 3 + i, cls, ghost=run["ghost"] or None,
                                       known=run["known"])
        p_or = o["wins"] / 60000
        for fast in (False, True):
            m = sampler_model.run_model(oracle, "reference", 50 + i, 0, hero, board, run["players"], T,
                                        opp_mask=ranges.opponent_mask(cls), ghost=ghost, fast=fast, known=known)
            p = (m["wins"] + m["ties"]) / T
            se = (p_or * (1 - p_or) * (1 / T + 1 / 60000)) ** 0.5
            assert abs(p - p_or) < 3.5 * se + 1e-9, (run["name"], fast, p, p_or, se)


@pytest.mark.gpu
def test_known_hands_kernels_match_their_specifications_and_the_oracle(cuda_device):
    """Opponents whose cards are known (npk_equity_ranges_known_batch): both range kernels equal their specifications bit for
    bit in both dealing modes; at 2 M trials they agree with the oracle (reference / unbiased restatement) at 3 sigma."""
    import neuron_poker_b200 as npk
    T, off = 96, 500
    for i, run in enumerate(_golden_known()["runs"][::2]):
        hero, board = _ids(run["hero"]), _ids(run["board"])
        known = [_ids(h) for h in run["known"]]
        ghost = _ids(run["ghost"]) if run["ghost"] else None
        cls = _opp_classes(run)
        args = (np.array([hero], dtype=np.uint8), np.array([board + [255] * (5 - len(board))], dtype=np.uint8),
                np.array([run["players"]], dtype=np.uint8))
        kw = dict(opponent_range=cls, ghost=None if ghost is None else np.array([ghost], dtype=np.uint8),
                  known_opponents=np.array([known], dtype=np.uint8))
        for mode in ("reference", "uniform"):
            for fast in (False, True):
                out = npk.get_equity_ranges_batch(*args, T, seed_value=31 + i, deal_mode=mode, trial_offset=off, query_offset=2,
                                                  win_types=True, passes=not fast, **kw)
                m = sampler_model.run_model(oracle, mode, 31 + i, 2, hero, board, run["players"], T, trial_offset=off,
                                            opp_mask=ranges.opponent_mask(cls), ghost=ghost, fast=fast, known=known)
                got = (int(out["wins"][0]), int(out["ties"][0]), [int(x) for x in out["win_types"][0]])
                assert got == (m["wins"], m["ties"], m["win_types"]), (run["name"], mode, fast, got, m)
                if not fast:
                    assert int(out["passes"][0]) == m["passes"], (run["name"], mode)
            big, TO = 2_000_000, 200_000
            out = npk.get_equity_ranges_batch(*args, big, seed_value=900 + i, deal_mode=mode, **kw)
            p_gpu = (int(out["wins"][0]) + int(out["ties"][0])) / big
            if mode == "reference":
                p_or = oracle.mc_reference_ranges(run["hero"], run["board"], run["players"], TO, 17 + i, cls,
                                                  ghost=run["ghost"] or None, known=run["known"])["wins"] / TO
            else:
                w, t, _ = oracle.mc_uniform_ranges(run["hero"], run["board"], run["players"], TO, 17 + i, cls,
                                                   ghost=run["ghost"] or None, known=run["known"])
                p_or = (w + t) / TO
            se = (p_or * (1 - p_or) * (1 / big + 1 / TO)) ** 0.5
            assert abs(p_gpu - p_or) < 3 * se + 1e-9, (run["name"], mode, p_gpu, p_or, se)


@pytest.mark.gpu
def test_run_montecarlo_with_several_known_hands(cuda_device):
    """MonteCarlo.run_montecarlo([hero, known hand, ...], ...) -- the reference's call form -- against its own seeded runs at
    3.5 sigma; error behaviour of the combinations that are not supported."""
    import neuron_poker_b200 as npk
    from neuron_poker_b200._lib import NpkError
    npk.seed(99)
    sim = npk.MonteCarlo()
    for run in _golden_known()["runs"][::2]:
        rng = set(run["opponent_range"]) if isinstance(run["opponent_range"], list) else run["opponent_range"]
        sim.run_montecarlo([run["hero"]] + run["known"], run["board"], run["players"], 1, maxRuns=200000, timeout=0,
                           ghost_cards=run["ghost"] or '', opponent_range=rng)
        p_ref = run["wins"] / run["runs"]
        se = (max(p_ref * (1 - p_ref), 1e-4) * (1 / 200000 + 1 / run["runs"])) ** 0.5
        assert abs(sim.equity - p_ref) < 3.5 * se, (run["name"], sim.equity, p_ref, se)
        assert sim.runs == 200000 and abs(sum(sim.winnerCardTypeList.values()) - sim.equity) < 1e-9
    npk.seed(None)
    with pytest.raises(NotImplementedError):
        sim.run_montecarlo([{'AKO', 'AA'}, ['QH', 'QD']], [], 3, 1, maxRuns=100, timeout=0, ghost_cards='')
    with pytest.raises(NpkError) as ei:
        sim.run_montecarlo([['AS', 'KS'], ['AS', 'QD']], [], 3, 1, maxRuns=100, timeout=0, ghost_cards='')      # a card twice
    assert ei.value.code == -5
    with pytest.raises(ValueError):
        sim.run_montecarlo([['AS', 'KS'], ['QH', 'QD'], ['JH', 'JD']], [], 2, 1, maxRuns=100, timeout=0, ghost_cards='')
