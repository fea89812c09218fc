"""The CPU oracle (oracle/) against the golden fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests then compare the kernels with it."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import CAT_NAMES


def test_class_list_matches_reference(golden_tables):
    classes = oracle.class_tuples()
    assert len(classes) == 5034
    sha = hashlib.sha256("\n".join(repr(t) for t in classes).encode()).hexdigest()
    assert sha == "e17e1892eea55478fa6462a276ac1d932d748efc91d8c680a9a8d41fe63bedbe"
    enc = golden_tables["classes"]
    for i in (0, 406, 407, 1893, 3215, 3224, 4657, 4735, 4919, 4997, 5033):
        score = oracle.CAT_SCORES[enc[i, 0]]
        ranks = tuple(int(r) for r in enc[i, 1:] if r != -2)
        assert classes[i] == (score, ranks)


def test_known_answer_showdowns(golden_cases):
    """reference tests/test_evaluator.py:9-133 (cases 3 and 6b hold duplicate physical cards)."""
    assert len(golden_cases["known"]) == 14
    for case in golden_cases["known"]:
        tuples = [oracle.calc_score(h) for h in case["hands"]]
        for t, exp, ty in zip(tuples, case["tuples"], case["types"]):
            assert t[0] == tuple(exp[0]) and t[1] == tuple(exp[1]) and t[2] == ty, case["name"]
        vals = [t[:2] for t in tuples]
        assert vals.index(max(vals)) == case["winner"], case["name"]


def test_nonflush_and_flush_tables(golden_tables):
    hist = golden_tables["hist"]
    cards = np.zeros((len(hist), 7), dtype=np.uint8)
    for i, h in enumerate(hist):
        k = 0
        for r in range(13):
            for _ in range(h[r]):
                cards[i, k] = 4 * r + (k & 3)
                k += 1
    assert (oracle.rank7_batch(cards) == golden_tables["nonflush"]).all()
    fl = golden_tables["flush"]
    for mask in range(8192):
        if fl[mask] == 0xFFFF:
            continue
        c = [4 * r for r in range(13) if mask >> r & 1]
        pad = 0
        while len(c) < 7:
            c.append(4 * pad + 1)
            pad += 1
        assert oracle.rank7(c) == fl[mask]


def test_seeded_hands_and_showdowns(golden_cases):
    hands = golden_cases["random_hands"] + golden_cases["rare_hands"]
    cards = np.array([h["cards"] for h in hands], dtype=np.uint8)
    assert (oracle.rank7_batch(cards) == np.array([h["rank_id"] for h in hands])).all()
    for h in hands[::97]:
        assert CAT_NAMES.index(oracle.calc_score(h["cards"])[2]) == h["type"]
    for sd in golden_cases["showdowns"]:
        w, ty = oracle.get_winner(sd["holes"], sd["board"])
        assert (w, CAT_NAMES.index(ty)) == (sd["winner"], sd["type"])


def test_mc_loop_bit_exact_under_numpy_seed(golden_mc):
    """run_montecarlo under np.random.seed: wins, passes, win types and the RNG position must all match."""
    assert len(golden_mc["runs"]) == 20
    for r in golden_mc["runs"]:
        o = oracle.mc_reference(r["hero"], r["board"], r["players"], r["runs"], r["seed"])
        assert o["wins"] == r["wins"] and o["passes"] == r["passes"], r["name"]
        assert o["win_types"] == r["win_types"], r["name"]
        assert o["next_randint"] == r["next_randint_0_1000000"], r["name"]


def test_enumeration_goldens(golden_enum):
    n = 0
    for s in golden_enum["spots"]:
        if s["players"] == 2 and "uniform" in s and len(s["board"]) >= 3:
            assert list(oracle.enum_headsup(s["hero"], s["board"])) == s["uniform"], s["name"]
            n += 1
        if s["players"] == 3 and "uniform" in s:
            assert list(oracle.enum_river_multi(s["hero"], s["board"], 3)) == s["uniform"], s["name"]
            n += 1
        if "reference_mode" in s:
            assert list(oracle.enum_reference_headsup(s["hero"], s["board"])) == s["reference_mode"], s["name"]
    assert n >= 14
    for s in golden_enum["random_spots"]:
        assert list(oracle.enum_headsup(s["hero"], s["board"])) == s["uniform"]


def test_uniform_port_agrees_with_exact_enumeration(golden_enum):
    for s in golden_enum["spots"]:
        if s["name"] not in ("t1", "t16", "t12"):
            continue
        w, t, l = s["uniform"]
        p = (w + t) / (w + t + l)
        runs = 40000
        ow, ot = oracle.mc_uniform(s["hero"], s["board"], 2, runs, 7)
        sigma = (p * (1 - p) / runs) ** 0.5
        assert abs((ow + ot) / runs - p) < 4 * sigma + 1e-9, s["name"]


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (reference sources absent)")
def test_reference_cpp_agrees_with_restatement(golden_cases):
    """The reference's own C++ calc_score (Montecarlo.cpp:53-237) against the C restatement, incl. both quirks."""
    for h in (golden_cases["random_hands"][:300] + golden_cases["rare_hands"][:300]):
        cards = [oracle.card_str(c) for c in h["cards"]]
        s, r, ty = oracle.ref_calc_score(cards)
        assert (s, r, ty) == oracle.calc_score(cards)
    for case in golden_cases["known"]:
        if case["name"] in ("3", "6b"):      # std::set<std::string> cannot hold the duplicate cards of these cases
            continue
        assert oracle.ref_eval_best_hand(case["hands"]) == (1 if case["winner"] == 0 else 0)
    eq = oracle.ref_montecarlo(["3H", "3S"], ["8S", "4S", "QH", "8C", "4H"], 2, 20000)
    assert abs(eq - 399 / 990) < 4 * (0.403 * 0.597 / 20000) ** 0.5


def test_exhaustive_colex_enumeration():
    """All C(52,7) hands: the tabulated oracle equals the statement-by-statement one on a slice, the hand-type census is
    the well-known 7-card frequency table, and the per-chunk checksums match the committed golden file."""
    import json
    import os
    slow_r, slow_s = oracle.colex_range(77_000_000, 60_000, fast=False)
    fast_r, fast_s = oracle.colex_range(77_000_000, 60_000, fast=True)
    assert (slow_r == fast_r).all() and (slow_s == fast_s).all()
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "colex_checksums.json")))
    cs = oracle.colex_checksums(gold["chunk"], threads=4)
    assert cs.tolist() == gold["chunks"]
    tot = cs.sum(0).tolist()
    assert tot == gold["total"] and sum(tot[2:]) == 133784560
    assert tot[2:] == [23294460, 58627800, 31433400, 6461620, 6180020, 4047644, 3473184, 224848, 41584]
