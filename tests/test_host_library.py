"""CPU-side checks of the product: libnpk.so loads and exports the whole C ABI, the host-built lookup tables equal the
golden tables generated from the reference, the Python boundary mirrors the reference's argument handling, and every
compute entry point fails loudly when no GPU is present (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import neuron_poker_b200 as npk
from neuron_poker_b200 import _lib, cards
import sampler_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_every_declared_symbol_is_exported():
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include")))
                     if h.endswith(".h"))
    names = set(re.findall(r"\b(npk_[a-z0-9_]+)\s*\(", header))
    assert {"npk_init", "npk_equity_batch", "npk_equity_host", "npk_rank7_batch", "npk_enum_batch",
            "npk_showdown_batch", "npk_rank7_colex", "npk_equity_ranges_batch", "npk_holdem_init", "npk_holdem_step",
            "npk_holdem_queries", "npk_holdem_decide"} <= names
    L = _lib.lib()
    for n in sorted(names):
        assert hasattr(L, n), n


def test_host_tables_equal_reference_tables(golden_tables):
    t = npk.host_tables()
    assert list(t["type_start"]) == [0, 407, 1877, 2640, 3215, 3225, 4502, 4658, 4736, 5034]
    gold_flush = golden_tables["flush"]
    valid = gold_flush != 0xFFFF
    assert (t["flush"][valid] == gold_flush[valid]).all()
    # every rank histogram through descriptor sums + row displacement (same arithmetic as the device code)
    hist = golden_tables["hist"].astype(np.uint32)
    desc_key = (t["desc"][::4] >> 9).astype(np.uint64)                 # mixed rank key of each rank (suit C cards)
    mk = (hist.astype(np.uint64) * desc_key[None, :]).sum(1) & ((1 << 23) - 1)
    ids = t["value"][t["rowoff"][(mk >> 10).astype(np.int64)].astype(np.int64) + (mk & 1023).astype(np.int64)]
    assert (ids == golden_tables["nonflush"]).all()
    assert len(t["value"]) * 2 + 2 * 8192 * 2 < 140 * 1024              # fits one CTA's shared memory with room for decks
    # class keys: type in the high word, same census as the reference
    types = (t["class_keys"] >> np.uint64(32)).astype(int)
    assert np.bincount(types, minlength=9).tolist() == [407, 1470, 763, 575, 10, 1277, 156, 78, 298]
    enc = golden_tables["classes"]
    assert (types == enc[:, 0]).all()
    for i in range(0, 5034, 7):
        ranks = [int(r) for r in enc[i, 1:] if r != -2]
        key = int(t["class_keys"][i]) & 0xFFFFFFFF
        got = [((key >> (4 * (7 - j))) & 15) - 2 for j in range(8) if (key >> (4 * (7 - j))) & 15]
        assert got == ranks, i


def test_host_rank7_against_golden_hands(golden_cases):
    hands = golden_cases["random_hands"] + golden_cases["rare_hands"]
    c = np.array([h["cards"] for h in hands], dtype=np.uint8)
    assert (npk.host_rank7(c) == np.array([h["rank_id"] for h in hands])).all()


def test_card_notation_and_errors():
    assert cards.DECK[:5] == ["2C", "2D", "2H", "2S", "3C"] and cards.DECK[-1] == "AS"     # montecarlo_python.py:114-119
    assert npk.card_id("AS") == 51 and npk.card_str(0) == "2C"
    hole, board = cards.encode_query({"AS", "KS"}, {"2C", "7D", "KH"})
    assert sorted(hole.tolist()) == [47, 51] and board.tolist()[3:] == [255, 255] and len(board) == 5
    with pytest.raises(ValueError):
        cards.encode_query(["AS", "KS"], ["XX"])                        # reference: list.index raises ValueError
    with pytest.raises(ValueError):
        cards.encode_query(["AS", "AS"], [])
    with pytest.raises(IndexError):
        npk.get_equity({"AS", "KS"}, set(), 0, 10)                      # reference: IndexError for players == 0


def test_philox_model_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    f = sampler_model.philox4x32_10
    assert f((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_model_is_consistent_with_enumeration(golden_enum):
    """The specified uniform sampler (Philox + Fisher-Yates) converges to the exact equity; the specified reference
    sampler converges to the exact expectation of the reference's biased dealer."""
    import oracle
    spot = next(s for s in golden_enum["spots"] if s["name"] == "t16")
    hole, board = [npk.card_id(c) for c in spot["hero"]], [npk.card_id(c) for c in spot["board"]]
    runs = 6000
    w, t, l = spot["uniform"]
    p = (w + t) / (w + t + l)
    r = sampler_model.run_model(oracle, "uniform", 11, 0, hole, board, 2, runs)
    assert abs((r["wins"] + r["ties"]) / runs - p) < 4 * (p * (1 - p) / runs) ** 0.5
    num, den = spot["reference_mode"]
    p = num / den
    r = sampler_model.run_model(oracle, "reference", 11, 0, hole, board, 2, runs)
    assert abs((r["wins"] + r["ties"]) / runs - p) < 4 * (p * (1 - p) / runs) ** 0.5
    assert sum(r["win_types"]) == r["wins"] + r["ties"]


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(_lib.NpkError) as e:
        npk.get_equity({"AS", "KS"}, set(), 2, 100)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(_lib.NpkError):
        npk.montecarlo({"AS", "KS"}, {"null"}, 2, 100)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under neuron_poker_b200/ may reference it."""
    pkg = os.path.join(ROOT, "neuron_poker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_reference_dealer_as_a_distribution_over_cards_is_exact():
    """the reference's index-based dealer (montecarlo_python.py:165-189) and the card-based rule that
    equity_refdeal_kernel implements -- uniform ordered pairs of distinct unseen cards minus the pairs whose second card
    is the successor of the first, board cards uniform over the unseen cards minus their maximum -- are the SAME
    distribution.  Exact enumeration with fractions on a nine-card deck: two opponents, then two board cards."""
    from fractions import Fraction
    from collections import defaultdict
    deck = [3, 7, 8, 20, 21, 22, 40, 50, 51]

    def reference(deck):
        out = defaultdict(Fraction)

        def opponents(r, k, prob, acc):
            if k == 0:
                return board(r, 2, prob, acc)
            n = len(r)
            valid = [(i1, i2) for i1 in range(n) for i2 in range(n - 1) if i1 != i2]
            for i1, i2 in valid:
                r2 = list(r)
                c1 = r2.pop(i1)
                c2 = r2.pop(i2)
                opponents(r2, k - 1, prob / len(valid), acc + ((c1, c2),))

        def board(r, k, prob, acc):
            if k == 0:
                out[acc] += prob
                return
            for j in range(len(r) - 1):                      # randint(0, len - 1): never the last element
                r2 = list(r)
                b = r2.pop(j)
                board(r2, k - 1, prob / (len(r) - 1), acc + (b,))

        opponents(list(deck), 2, Fraction(1), ())
        return out

    def card_rule(deck):
        out = defaultdict(Fraction)

        def opponents(avail, k, prob, acc):
            if k == 0:
                return board(avail, 2, prob, acc)
            pairs = []
            for c1 in avail:
                above = [c for c in avail if c > c1]
                succ = min(above) if above else None
                pairs += [(c1, c2) for c2 in avail if c2 != c1 and c2 != succ]
            for c1, c2 in pairs:
                opponents(avail - {c1, c2}, k - 1, prob / len(pairs), acc + ((c1, c2),))

        def board(avail, k, prob, acc):
            if k == 0:
                out[acc] += prob
                return
            cands = sorted(avail - {max(avail)})
            for b in cands:
                board(avail - {b}, k - 1, prob / len(cands), acc + (b,))

        opponents(frozenset(deck), 2, Fraction(1), ())
        return out

    a, b = reference(deck), card_rule(deck)
    assert sum(a.values()) == 1 and sum(b.values()) == 1
    assert a == b and len(a) > 10000


def test_reference_dealer_without_retries_is_the_reference_distribution():
    """The rejection-free form equity_refdeal_kernel uses (tests/sampler_model.py::deal_reference): (a, b) uniform in
    [0,n-1)^2 -> (i1, i2) = (a + (a >= b), b) hits every (i1, i2) the reference's retry loop accepts
    (montecarlo_python.py:169-172: i1 in [0,n), i2 in [0,n-1), i1 != i2) exactly once; and the backward Lehmer sweep
    yields the cards list.pop would."""
    import random
    import sampler_model as sm
    for n in range(2, 12):
        accepted = sorted((i1, i2) for i1 in range(n) for i2 in range(n - 1) if i1 != i2)
        mapped = sorted((a + (1 if a >= b else 0), b) for a in range(n - 1) for b in range(n - 1))
        assert mapped == accepted, n
    rnd = random.Random(3)
    for _ in range(300):
        n, d = rnd.randint(24, 50), rnd.randint(0, 23)
        raw = [rnd.randrange(n - k) for k in range(d)]
        lst = list(range(n))
        assert [lst.pop(i) for i in raw] == sm.lehmer_slots(raw)


def test_c_binding_of_the_one_query_call_packs_like_the_python_path():
    """neuron_poker_b200/_npkfast.so (csrc/npk_pyfast.c) is a second way into npk_equity_one: against a recording stand-in for
    that entry point (no GPU needed) it passes the same packed query, player count, trial count, dealing mode and seed as
    equity.equity_counts would, and hands everything that is not the common case back to Python (None)."""
    import ctypes
    import importlib.util
    from neuron_poker_b200 import _build, _lib
    from neuron_poker_b200.equity import _pack_query
    path = _build.build_fast()
    spec = importlib.util.spec_from_file_location("_npkfast", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.bind(0)
    assert mod.equity({"AS", "KS"}, set(), 2, 100, 1, lambda: 0.5) is None            # not bound
    seen = []
    proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                             ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64))

    def fake(packed, players, trials, seed, mode, want, out):
        seen.append((packed, players, trials, seed, mode, want))
        out[0], out[1] = trials // 4, trials // 2
        return 0 if players != 7 else -5

    cb = proto(fake)
    try:
        mod.bind(ctypes.cast(cb, ctypes.c_void_p).value)
        _check_c_binding(mod, seen, _pack_query)
    finally:          # the shared object has one binding per process: give it back to the real library if that is loaded
        mod.bind(ctypes.cast(_lib._lib.npk_equity_one, ctypes.c_void_p).value if _lib._lib is not None else 0)


def _check_c_binding(mod, seen, _pack_query):
    cases = [({"AS", "KS"}, {"2C", "7D", "KH"}, 6), (["3H", "3S"], ("8S", "4S", "QH", "8C", "4H"), 2), (("TD", "7D"), [], 10),
             ({"QC", "QD"}, ["2C", "7D", "KH", "9S"], np.int64(3))]
    for hole, board, players in cases:
        seen.clear()
        e = mod.equity(hole, board, players, 1000, 1, lambda: 0.25)
        assert e == 0.75
        assert seen == [(_pack_query(hole, board), int(players), 1000, 2 ** 51, 1, 0)]
    for bad in [({"AS"}, set(), 2, 100), ({"AS", "KS", "QS"}, set(), 2, 100), ({"AS", "Kx"}, set(), 2, 100),
                ({"AS", "KS"}, {"null"}, 2, 100), ({"AS", "KS"}, set(), 0, 100), ({"AS", "KS"}, set(), 11, 100),
                ({"AS", "KS"}, set(), 2, 0), ({"AS", "KS"}, set(), 2.5, 100), ({"AS", "KS"}, set(), 7, 100),
                ("ASKS", set(), 2, 100), ({"AS", "KS"}, {"2C", "3C", "4C", "5C", "6C", "7C"}, 2, 100)]:
        assert mod.equity(bad[0], bad[1], bad[2], bad[3], 0, lambda: 0.5) is None, bad
