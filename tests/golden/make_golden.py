#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py            # everything (~3-4 min, single core)
    python tests/golden/make_golden.py --quick    # skips the two 1,070,190-matchup flop spots

What it imports from the reference (never copied, only called):
  * tools/hand_evaluator.py  `_calc_score` (:27-119), `eval_best_hand` (:20-24), `get_winner` (:9-17)
  * tools/montecarlo_python.py `MonteCarlo.run_montecarlo` (:191-252)

Files written (all small, committed):
  eval_tables.npz      class list (5,034 tuples), rank_id per rank histogram (49,205) and per flush mask (8,192)
  eval_cases.json      the 14 known-answer showdowns of tests/test_evaluator.py + seeded random hands/showdowns
  enum_golden.json     exact (win, tie, lose) enumeration of the reference's Monte-Carlo test spots,
                       in UNIFORM dealing and, where enumerable, the exact expectation of the REFERENCE dealer
  mc_seeded.json       run_montecarlo under np.random.seed(s): wins / passes / win-type counts per spot, used to pin
                       the oracle's MT19937 + legacy randint + dealing restatement bit-exactly
  preflop_order.json   the reference's ranking of the 169 starting-hand classes and the allowed sets of some fractions
  mc_ranges_seeded.json seeded run_montecarlo runs with opponent ranges, hero ranges and ghost cards
  mc_known_seeded.json  seeded run_montecarlo runs with several known hands in player_card_list (:132-163)
"""
import argparse
import hashlib
import itertools
import json
import os
import random
import sys
import time
from collections import Counter

import numpy as np

REF = os.environ.get("NPK_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from tools.hand_evaluator import _calc_score, eval_best_hand, get_winner  # noqa: E402
from tools import montecarlo_python  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RANKS = "23456789TJQKA"
SUITS = "CDHS"
DECK = [r + s for r in RANKS for s in SUITS]  # montecarlo_python.py:114-119 order: id = 4*rank + suit
CATS = [(1,), (2, 1, 1), (2, 2, 1), (3, 1), (3, 1, 2), (3, 1, 3), (3, 2), (4,), (5,)]
CAT_NAMES = ["HighCard", "Pair", "TwoPair", "ThreeOfAKind", "Straight", "Flush", "FullHouse", "FoufOfAKind",
             "StraightFlush"]


def histograms():
    """All 49,205 seven-card rank histograms (c_2..c_A), each <=4, ascending lexicographic tuple order."""
    out = []

    def rec(prefix, left):
        if len(prefix) == 12:
            if left <= 4:
                out.append(tuple(prefix) + (left,))
            return
        for c in range(0, min(4, left) + 1):
            rec(prefix + [c], left - c)

    rec([], 7)
    out.sort()
    return out


def nonflush_hand(hist):
    """A 7-card hand with this rank histogram and no suit holding >=5 cards (suits assigned round-robin)."""
    hand = []
    k = 0
    for r, c in enumerate(hist):
        for j in range(c):
            # round-robin over suits across the 7 cards -> max suit count is 2; within a rank suits stay distinct
            hand.append(RANKS[r] + SUITS[(k + j) % 4])
        k += c
    assert len(set(hand)) == 7, hand
    assert max(Counter(s for _, s in hand).values()) < 5
    return hand


def flush_hand(mask):
    """A 7-card hand whose suit-C cards are exactly the ranks in `mask` (5..7 bits), padded with low off-suit cards."""
    ranks = [r for r in range(13) if mask >> r & 1]
    hand = [RANKS[r] + "C" for r in ranks]
    pads = [RANKS[r] + s for r in range(13) for s in "DHS"]
    i = 0
    while len(hand) < 7:
        hand.append(pads[i])
        i += 3  # different ranks -> 2D, 3D: never adds a third suit count >= 5
    return hand


def tuple_of(hand):
    score, ranks, _ = _calc_score(hand)
    return (tuple(score), tuple(ranks))


def build_tables():
    hists = histograms()
    assert len(hists) == 49205
    nf = [tuple_of(nonflush_hand(h)) for h in hists]
    fl = {}
    for mask in range(8192):
        if 5 <= bin(mask).count("1") <= 7:
            fl[mask] = tuple_of(flush_hand(mask))
    assert len(fl) == 4719
    classes = sorted(set(nf) | set(fl.values()))
    assert len(classes) == 5034, len(classes)
    cid = {t: i for i, t in enumerate(classes)}
    nf_ids = np.array([cid[t] for t in nf], dtype=np.uint16)
    fl_ids = np.full(8192, 0xFFFF, dtype=np.uint16)
    for m, t in fl.items():
        fl_ids[m] = cid[t]
    # class tuples as a fixed-width int8 array: [cat_index, r0..r7] padded with -2
    enc = np.full((len(classes), 9), -2, dtype=np.int8)
    for i, (score, ranks) in enumerate(classes):
        enc[i, 0] = CATS.index(score)
        enc[i, 1:1 + len(ranks)] = ranks
    sha_classes = hashlib.sha256("\n".join(repr(t) for t in classes).encode()).hexdigest()
    sha_nf = hashlib.sha256(nf_ids.astype("<u2").tobytes()).hexdigest()
    sha_fl = hashlib.sha256(fl_ids.astype("<u2").tobytes()).hexdigest()
    # SURVEY.md A.1-13 checksums (measured in the survey session with the same reference)
    assert sha_classes == "e17e1892eea55478fa6462a276ac1d932d748efc91d8c680a9a8d41fe63bedbe", sha_classes
    assert sha_nf == "74878c4519e21e50d6fbbfb2c6827a226693806381a6f8c1c802217c8b4643c8", sha_nf
    assert sha_fl == "c96a3fdd598345c9041c96a00ea63f61ca2ddba3b90a2b2ab718d4a90abec030", sha_fl
    np.savez_compressed(os.path.join(HERE, "eval_tables.npz"), classes=enc, nonflush=nf_ids, flush=fl_ids,
                        hist=np.array(hists, dtype=np.uint8))
    print("eval_tables.npz: classes", len(classes), "sha ok")
    return cid


# tests/test_evaluator.py:9-133 -- (hands, expected winner index); cases 3 and 6b hold duplicate physical cards
KNOWN = [
    ("1", [['3H', '3S', '4H', '4S', '8S', '8C', 'QH'], ['KH', '6C', '4H', '4S', '8S', '8C', 'QH']], 1),
    ("2", [['8H', '8D', 'QH', '7H', '9H', 'JH', 'TH'], ['KH', '6C', 'QH', '7H', '9H', 'JH', 'TH']], 1),
    ("3", [['AS', 'KS', 'TS', '9S', '7S', '2H', '2H'], ['AS', 'KS', 'TS', '9S', '8S', '2H', '2H']], 1),
    ("4", [['8S', 'TS', '8H', 'KS', '9S', 'TH', 'KH'], ['TD', '7S', '8H', 'KS', '9S', 'TH', 'KH']], 0),
    ("5", [['2D', '2H', 'AS', 'AD', 'AH', '8S', '7H'], ['7C', '7S', '7H', 'AD', 'AS', '8S', '8H']], 0),
    ("6", [['7C', '7S', '7H', 'AD', 'KS', '5S', '8H'], ['2D', '3H', 'AS', '4D', '5H', '8S', '7H']], 1),
    ("6b", [['7C', '7C', 'AC', 'AC', '8C', '8S', '7H'], ['2C', '3C', '4C', '5C', '6C', '8S', 'KH']], 1),
    ("7", [['AC', 'JS', 'AS', '2D', '5H', '3S', '3H'], ['QD', 'JD', 'TS', '9D', '6H', '8S', 'KH'],
           ['2D', '3D', '4S', '5D', '6H', '8S', 'KH']], 1),
    ("8", [['7C', '5S', '3S', 'JD', '8H', '2S', 'KH'], ['AD', '3D', '4S', '5D', '9H', '8S', 'KH']], 1),
    ("9", [['2C', '2D', '4S', '4D', '4H', '8S', 'KH'], ['7C', '7S', '7D', '7H', '8H', '8S', 'JH']], 1),
    ("10", [['7C', '5S', '3S', 'JD', '8H', '2S', 'KH'], ['AD', '3D', '3S', '5D', '9H', '8S', 'KH']], 1),
    ("11", [['7H', '7S', '3S', 'JD', '8H', '2S', 'KH'], ['7D', '3D', '3S', '7C', '9H', '8S', 'KH']], 1),
    ("12", [['AS', '8H', 'TS', 'JH', '3H', '2H', 'AH'], ['QD', 'QH', 'TS', 'JH', '3H', '2H', 'AH']], 1),
    ("13", [['9S', '7H', 'KS', 'KH', 'AH', 'AS', 'AC'], ['8D', '2H', 'KS', 'KH', 'AH', 'AS', 'AC']], 0),
]


def build_eval_cases(cid):
    known = []
    for name, hands, expected in KNOWN:
        best, htype = eval_best_hand(hands)
        assert best == hands[expected], name
        tuples = [_calc_score(h) for h in hands]
        known.append({"name": name, "hands": hands, "winner": expected, "winner_type": htype,
                      "types": [t[2] for t in tuples],
                      "tuples": [[list(t[0]), list(t[1])] for t in tuples]})
    rng = random.Random(20261018)
    hands7 = []
    for _ in range(20000):
        h = rng.sample(DECK, 7)
        score, ranks, htype = _calc_score(h)
        hands7.append({"cards": [DECK.index(c) for c in h], "rank_id": cid[(tuple(score), tuple(ranks))],
                       "type": CAT_NAMES.index(htype)})
    # hands biased towards the rare categories (flush-heavy / rank-heavy decks) so quirks get exercised
    rare = []
    for _ in range(6000):
        if rng.random() < 0.5:
            s = rng.choice(SUITS)
            pool = [r + s for r in RANKS] + rng.sample(DECK, 6)
        else:
            rs = rng.sample(RANKS, 3)
            pool = [r + s for r in rs for s in SUITS] + rng.sample(DECK, 4)
        pool = list(dict.fromkeys(pool))
        h = rng.sample(pool, 7)
        score, ranks, htype = _calc_score(h)
        rare.append({"cards": [DECK.index(c) for c in h], "rank_id": cid[(tuple(score), tuple(ranks))],
                     "type": CAT_NAMES.index(htype)})
    showdowns = []
    for _ in range(3000):
        n = rng.randint(2, 9)
        cards = rng.sample(DECK, 5 + 2 * n)
        board = cards[:5]
        holes = [cards[5 + 2 * i: 7 + 2 * i] for i in range(n)]
        ix, htype = get_winner(holes, board)
        showdowns.append({"board": [DECK.index(c) for c in board],
                          "holes": [[DECK.index(c) for c in h] for h in holes],
                          "winner": ix, "type": CAT_NAMES.index(htype)})
    with open(os.path.join(HERE, "eval_cases.json"), "w") as f:
        json.dump({"source": "tests/test_evaluator.py:9-133 + tools/hand_evaluator.py via make_golden.py",
                   "known": known, "random_hands": hands7, "rare_hands": rare, "showdowns": showdowns}, f)
    print("eval_cases.json:", len(known), "known,", len(hands7) + len(rare), "hands,", len(showdowns), "showdowns")


# tests/test_montecarlo_python.py spots: (name, line, hero, board, players, expected percent)
SPOTS = [
    ("t1", 44, ['3H', '3S'], ['8S', '4S', 'QH', '8C', '4H'], 2, 40.2),
    ("t2", 53, ['8H', '8D'], ['QH', '7H', '9H', 'JH', 'TH'], 2, 95.6),
    ("t3", 62, ['AS', 'KS'], [], 3, 51.8),
    ("t4", 71, ['AS', 'KS'], [], 2, 67.7),
    ("t5", 80, ['8S', 'TS'], ['8H', 'KS', '9S', 'TH', 'KH'], 2, 77.4),
    ("t6", 89, ['8S', 'TS'], ['2S', '3S', '4S', 'KS', 'AS'], 2, 87.0),
    ("t7", 98, ['8S', '2S'], ['5S', '3S', '4S', 'KS', 'AS'], 2, 100.0),
    ("t8", 107, ['8S', 'TS'], [], 5, 25.5),
    ("t8b", 116, ['2C', 'QS'], [], 2, 49.6),
    ("t9", 125, ['7H', '7S'], ['7C', '8C', '8S', 'AC', 'AH'], 2, 83.0),
    ("t10", 134, ['3S', 'QH'], ['2C', '5H', '7C'], 2, 33.1),
    ("t11", 143, ['5C', 'JS'], [], 4, 23.0),
    ("t12", 152, ['TC', 'TH'], ['4D', 'QD', 'KC'], 2, 67.08),
    ("t13", 161, ['JH', 'QS'], ['5C', 'JD', 'AS', 'KS', 'QD'], 2, 77.0),
    ("t14", 170, ['2H', '8S'], ['AC', 'AD', 'AS', 'KS', 'KD'], 2, 95.0),
    ("t15", 179, ['KD', 'KS'], ['4D', '6S', '9C', '9S', 'TC'], 2, 88.0),
    ("t16", 188, ['5H', 'KD'], ['KH', 'JS', '2C', 'QS'], 2, 79.2),
    ("t17", 197, ['JD', 'JS'], ['8C', 'TC', 'JC', '5H', 'QC'], 3, 26.1),
    ("t19", 206, ['TD', '7D'], ['8D', 'QD', '7C', '5D', '6D'], 2, 87.0),
]


def val(hand):
    return _calc_score(hand)[:2]


def enum_uniform_hu(hero, board):
    """Exact heads-up enumeration with uniform dealing: all board completions x all opponent pairs."""
    rest = [c for c in DECK if c not in hero and c not in board]
    win = tie = lose = 0
    for extra in itertools.combinations(rest, 5 - len(board)):
        full = board + list(extra)
        hv = val(hero + full)
        rest2 = [c for c in rest if c not in extra]
        for opp in itertools.combinations(rest2, 2):
            ov = val(list(opp) + full)
            if hv > ov:
                win += 1
            elif hv == ov:
                tie += 1
            else:
                lose += 1
    return win, tie, lose


def enum_uniform_3way_river(hero, board):
    """Ordered pairs of disjoint opponent hands on a complete board."""
    rest = [c for c in DECK if c not in hero and c not in board]
    hv = val(hero + board)
    pairs = list(itertools.combinations(rest, 2))
    vals = {p: val(list(p) + board) for p in pairs}
    win = tie = lose = 0
    for p1 in pairs:
        for p2 in pairs:
            if set(p1) & set(p2):
                continue
            best = max(vals[p1], vals[p2])
            if hv > best:
                win += 1
            elif hv == best:
                tie += 1
            else:
                lose += 1
    return win, tie, lose


def enum_reference_hu(hero, board):
    """Exact expectation of the reference's own dealer (montecarlo_python.py:165-189), heads-up, board size 4 or 5.

    Opponent: every index pair (i1 in [0,n), i2 in [0,n-1), i1 != i2) is equally likely; c1 = deck.pop(i1),
    c2 = deck.pop(i2) on the shortened list. Then each missing board card: j uniform in [0, len-1) -> never the last.
    Returns (numerator of hero>=opp, denominator).
    """
    deck = [c for c in DECK if c not in board]
    deck = [c for c in deck if c not in hero]  # hero cards popped after the board cards (:126-163)
    n = len(deck)
    num = den = 0
    for i1 in range(n):
        for i2 in range(n - 1):
            if i1 == i2:
                continue
            d = list(deck)
            c1 = d.pop(i1)
            c2 = d.pop(i2)
            if len(board) == 5:
                hv, ov = val(hero + board), val([c1, c2] + board)
                num += hv >= ov
                den += 1
            else:
                assert len(board) == 4
                for j in range(len(d) - 1):
                    full = board + [d[j]]
                    num += val(hero + full) >= val([c1, c2] + full)
                    den += 1
    return num, den


def build_enum(quick):
    out = {"source": "exact enumeration with tools/hand_evaluator.py::_calc_score via make_golden.py", "spots": []}
    for name, line, hero, board, players, expected in SPOTS:
        rec = {"name": name, "ref_line": line, "hero": hero, "board": board, "players": players,
               "test_expected_pct": expected}
        t0 = time.time()
        if players == 2 and len(board) >= 4:
            rec["uniform"] = list(enum_uniform_hu(hero, board))
            rec["reference_mode"] = list(enum_reference_hu(hero, board))
        elif players == 2 and len(board) == 3:
            if not quick:
                rec["uniform"] = list(enum_uniform_hu(hero, board))
        elif players == 3 and len(board) == 5:
            rec["uniform"] = list(enum_uniform_3way_river(hero, board))
        out["spots"].append(rec)
        print(" enum", name, rec.get("uniform"), rec.get("reference_mode"), "%.1fs" % (time.time() - t0))
    # random river / turn spots (cfg 2 style), heads-up
    rng = random.Random(7)
    rnd = []
    for k in range(24):
        nb = 5 if k < 16 else 4
        cards = rng.sample(DECK, 2 + nb)
        rnd.append({"hero": cards[:2], "board": cards[2:], "uniform": list(enum_uniform_hu(cards[:2], cards[2:]))})
    out["random_spots"] = rnd
    if quick and os.path.exists(os.path.join(HERE, "enum_golden.json")):
        old = json.load(open(os.path.join(HERE, "enum_golden.json")))
        keep = {s["name"]: s for s in old["spots"]}
        for s in out["spots"]:
            if "uniform" not in s and "uniform" in keep.get(s["name"], {}):
                s["uniform"] = keep[s["name"]]["uniform"]
    with open(os.path.join(HERE, "enum_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("enum_golden.json written")


def build_mc_seeded():
    """Seeded runs of the reference loop with the 1-second cut-off disabled (timeout far in the future)."""
    out = {"source": "tools/montecarlo_python.py::MonteCarlo.run_montecarlo under np.random.seed(seed), "
                     "timeout=+1e9 s, opponent_range=1", "numpy": np.__version__, "runs": []}
    cases = [
        ("t4", ['AS', 'KS'], [], 2, 3000), ("t3", ['AS', 'KS'], [], 3, 2000), ("t8", ['8S', 'TS'], [], 5, 1500),
        ("t1", ['3H', '3S'], ['8S', '4S', 'QH', '8C', '4H'], 2, 3000),
        ("t16", ['5H', 'KD'], ['KH', 'JS', '2C', 'QS'], 2, 3000),
        ("t10", ['3S', 'QH'], ['2C', '5H', '7C'], 2, 3000),
        ("flop6", ['AS', 'KS'], ['2C', '7D', 'KH'], 6, 2000),
        ("pre9", ['AS', 'KS'], [], 9, 1200), ("pre10", ['7D', '7C'], [], 10, 800),
        ("river4", ['JD', 'JS'], ['8C', 'TC', 'JC', '5H', 'QC'], 4, 1500),
    ]
    for name, hero, board, players, runs in cases:
        for seed in (1, 12345):
            np.random.seed(seed)
            mc = montecarlo_python.MonteCarlo()
            mc.run_montecarlo([list(hero)], list(board), players, 1, maxRuns=runs, timeout=time.time() + 1e9,
                              ghost_cards='', opponent_range=1)
            wins = int(round(mc.equity * mc.runs))
            types = {k: int(round(v * mc.runs)) for k, v in mc.winnerCardTypeList.items()}
            assert sum(types.values()) == wins
            out["runs"].append({"name": name, "hero": hero, "board": board, "players": players, "seed": seed,
                                "runs": mc.runs, "wins": wins, "passes": int(mc.passes), "win_types": types,
                                "next_randint_0_1000000": int(np.random.randint(0, 1000000))})
        print(" mc", name, out["runs"][-1]["wins"], "/", runs)
    with open(os.path.join(HERE, "mc_seeded.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mc_seeded.json written")


def build_ranges():
    """Ranges: the reference's ranking of the 169 starting-hand classes (get_opponent_allowed_cards_list), the allowed
    sets of a few fractions, and seeded run_montecarlo runs with an opponent range, a hero RANGE (player_card_list[0]
    is a set) and ghost cards (reference tests test_montecarlo20/21 use range 0.25 and the hero set {'AKO','AA'})."""
    import operator
    mc = montecarlo_python.MonteCarlo()
    mc.get_opponent_allowed_cards_list(1)
    order = [k for k, _ in sorted(mc.preflop_equities.items(), key=operator.itemgetter(1))]
    fractions = {str(x): sorted(mc.get_opponent_allowed_cards_list(x)) for x in (0.25, 0.5, 0.1, 0.01, 0.004, 1, 1.5, 0.999)}
    with open(os.path.join(HERE, "preflop_order.json"), "w") as f:
        json.dump({"source": "tools/montecarlo_python.py:37-110, keys of preflop_equities sorted ascending by value",
                   "order": order, "allowed": fractions}, f, indent=1)
    out = {"source": "MonteCarlo.run_montecarlo under np.random.seed(seed), timeout=+1e9 s, with opponent_range / hero "
                     "range / ghost cards", "numpy": np.__version__, "runs": []}
    cases = [
        # name, hero (list of cards or set of classes), board, players, runs, opponent_range, ghost
        ("t20", ['KS', 'KC'], ['3D', '9H', 'AS', '7S', 'QH'], 3, 1500, 0.25, ''),
        ("t21", {'AKO', 'AA'}, ['3D', '9H', 'AS', '7S', 'QH'], 3, 800, 0.25, ''),
        ("pre2_r10", ['QS', 'QH'], [], 2, 1500, 0.1, ''),
        ("flop4_r50", ['AS', 'KS'], ['2C', '7D', 'KH'], 4, 1000, 0.5, ''),
        ("turn3_set", ['9D', '9C'], ['2C', '7D', 'KH', 'TS'], 3, 1000, {'AKS', 'KAO', 'QQ', 'JJ', '9TS', 'T9O'}, ''),
        ("flop3_ghost", ['AS', 'KS'], ['2C', '7D', 'KH'], 3, 1000, 1, ['AH', 'AD']),
        ("hero_set_pre", {'AKS', 'QQ', '78S'}, [], 2, 600, 0.5, ''),
        ("tiny_range", ['5H', '5D'], ['5C', 'KD', '2S'], 2, 800, 0.004, ''),
    ]
    for name, hero, board, players, runs, rng, ghost in cases:
        for seed in (3, 777):
            np.random.seed(seed)
            m = montecarlo_python.MonteCarlo()
            first = set(hero) if isinstance(hero, set) else list(hero)
            m.run_montecarlo([first], list(board), players, 1, maxRuns=runs, timeout=time.time() + 1e9,
                             ghost_cards=ghost, opponent_range=rng)
            wins = int(round(m.equity * m.runs))
            types = {k: int(round(v * m.runs)) for k, v in m.winnerCardTypeList.items()}
            assert sum(types.values()) == wins
            out["runs"].append({"name": name, "hero": sorted(hero) if isinstance(hero, set) else hero,
                                "hero_is_range": isinstance(hero, set), "board": board, "players": players,
                                "seed": seed, "runs": m.runs, "opponent_range": sorted(rng) if isinstance(rng, set) else rng,
                                "ghost": list(ghost) if ghost else [], "wins": wins, "passes": int(m.passes),
                                "win_types": types, "next_randint_0_1000000": int(np.random.randint(0, 1000000))})
        print(" ranges", name, out["runs"][-1]["wins"], "/", runs, "passes", out["runs"][-1]["passes"])
    with open(os.path.join(HERE, "mc_ranges_seeded.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mc_ranges_seeded.json written")


def build_known_hands():
    """Several known hands in player_card_list (montecarlo_python.py:132-163: the reference's provision for bots sharing a
    table): the hero's hand first, then hands of opponents whose cards are known; the other opponents are dealt at random."""
    out = {"source": "MonteCarlo.run_montecarlo under np.random.seed(seed), timeout=+1e9 s, player_card_list = [hero, known "
                     "hands...]", "numpy": np.__version__, "runs": []}
    cases = [
        # name, hero, known hands, board, players, runs, opponent_range, ghost
        ("known1_flop4", ['AS', 'KS'], [['QH', 'QD']], ['2C', '7D', 'KH'], 4, 1000, 1, ''),
        ("known2_pre5_r50", ['TS', 'TH'], [['AD', 'KD'], ['7C', '2D']], [], 5, 800, 0.5, ''),
        ("known1_turn2", ['9D', '9C'], [['AH', 'KC']], ['2C', '7D', 'KH', 'TS'], 2, 1200, 1, ''),
        ("known1_ghost_set", ['AS', 'KS'], [['JH', 'JD']], ['2C', '7D', 'KH'], 3, 1000, {'AKS', 'KAO', 'QQ', '9TS', 'T9O', '22'},
         ['AH', 'AD']),
        ("known3_river6", ['5H', '5D'], [['AD', 'KD'], ['QC', 'QS'], ['8H', '9H']], ['5C', 'KS', '2S', 'TH', 'JH'], 6, 800, 0.3, ''),
    ]
    for name, hero, known, board, players, runs, rng, ghost in cases:
        for seed in (5, 4242):
            np.random.seed(seed)
            m = montecarlo_python.MonteCarlo()
            m.run_montecarlo([list(hero)] + [list(k) for k in known], list(board), players, 1, maxRuns=runs,
                             timeout=time.time() + 1e9, ghost_cards=ghost, opponent_range=rng)
            wins = int(round(m.equity * m.runs))
            types = {k: int(round(v * m.runs)) for k, v in m.winnerCardTypeList.items()}
            assert sum(types.values()) == wins
            out["runs"].append({"name": name, "hero": hero, "known": known, "board": board, "players": players, "seed": seed,
                                "runs": m.runs, "opponent_range": sorted(rng) if isinstance(rng, set) else rng,
                                "ghost": list(ghost) if ghost else [], "wins": wins, "passes": int(m.passes),
                                "win_types": types, "next_randint_0_1000000": int(np.random.randint(0, 1000000))})
        print(" known hands", name, out["runs"][-1]["wins"], "/", runs, "passes", out["runs"][-1]["passes"])
    with open(os.path.join(HERE, "mc_known_seeded.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mc_known_seeded.json written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None
    cid = build_tables()
    if not only or "cases" in only:
        build_eval_cases(cid)
    if not only or "mc" in only:
        build_mc_seeded()
    if not only or "enum" in only:
        build_enum(a.quick)
    if not only or "ranges" in only:
        build_ranges()
    if not only or "known" in only:
        build_known_hands()


if __name__ == "__main__":
    main()
