"""ctypes front-end of oracle/liboracle.so (C restatement) and oracle/_ref/libnpk_ref.so (the reference's own C++).

TEST INFRASTRUCTURE ONLY.  Card ids are 4*rank + suit (ranks "23456789TJQKA", suits "CDHS"), the deck order of
reference tools/montecarlo_python.py:114-119.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RANKS = "23456789TJQKA"
SUITS = "CDHS"
CAT_NAMES = ["HighCard", "Pair", "TwoPair", "ThreeOfAKind", "Straight", "Flush", "FullHouse", "FoufOfAKind",
             "StraightFlush"]
CAT_SCORES = [(1,), (2, 1, 1), (2, 2, 1), (3, 1), (3, 1, 2), (3, 1, 3), (3, 2), (4,), (5,)]

_lib = None
_ref = None


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference sources are present)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "npk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", os.path.join(HERE, "liboracle.so")])
    ref_so = os.path.join(HERE, "_ref", "libnpk_ref.so")
    if os.path.exists("/root/reference/tools/montecarlo_cpp/Montecarlo.cpp") and (force or not os.path.exists(ref_so)):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(os.path.join(HERE, "liboracle.so"))
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_calc_score_flat.argtypes = [u8p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.oracle_key_cards.argtypes = [u8p, ctypes.c_int]
        L.oracle_key_cards.restype = ctypes.c_uint64
        L.oracle_build_classes.restype = ctypes.c_int
        L.oracle_class_keys.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
        L.oracle_rank7.argtypes = [u8p]
        L.oracle_rank7_batch.argtypes = [u8p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint16)]
        L.oracle_rank7_batch.restype = None
        L.oracle_type7.argtypes = [u8p]
        L.oracle_rank7_fast.argtypes = [u8p]
        L.oracle_fast_init.restype = None
        L.oracle_colex_range.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint16), i64p, ctypes.c_int]
        L.oracle_get_winner.argtypes = [u8p, ctypes.c_int, u8p, ctypes.POINTER(ctypes.c_int)]
        for name in ("oracle_mc_reference", "oracle_mc_uniform"):
            getattr(L, name).argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, i64p]
        u64p = ctypes.POINTER(ctypes.c_uint64)
        for name in ("oracle_mc_reference_ranges", "oracle_mc_uniform_ranges"):
            getattr(L, name).argtypes = [u8p, u64p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32,
                                         u64p, u8p, u8p, ctypes.c_int, i64p]
        L.oracle_hand_class.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_enum_headsup.argtypes = [u8p, u8p, ctypes.c_int, i64p]
        L.oracle_enum_river_multi.argtypes = [u8p, u8p, ctypes.c_int, i64p]
        L.oracle_enum_reference_headsup.argtypes = [u8p, u8p, ctypes.c_int, i64p]
        _lib = L
    return _lib


def card_id(card):
    return 4 * RANKS.index(card[0]) + SUITS.index(card[1])


def card_str(cid):
    return RANKS[cid >> 2] + SUITS[cid & 3]


def ids(cards):
    return np.array([c if isinstance(c, (int, np.integer)) else card_id(c) for c in cards], dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def calc_score(cards):
    """(score tuple, card_ranks tuple, hand type name) exactly like hand_evaluator._calc_score."""
    a = ids(cards)
    out = (ctypes.c_int * 20)()
    rc = lib().oracle_calc_score_flat(_p(a), len(a), out)
    if rc:
        raise Exception("Card Type error!")
    ns, nr = out[1], out[10]
    return tuple(out[2:2 + ns]), tuple(out[11:11 + nr]), CAT_NAMES[out[0]]


def class_tuples():
    """The 5,034 (score, card_ranks) tuples in ascending order, decoded from the oracle's 64-bit keys."""
    n = lib().oracle_build_classes()
    keys = (ctypes.c_uint64 * n)()
    lib().oracle_class_keys(keys, n)
    out = []
    for k in keys:
        ranks = []
        for i in range(8):
            v = (k >> (4 * (7 - i))) & 15
            if v:
                ranks.append(v - 2)
        score = []
        for i in range(7):
            v = (k >> (32 + 3 * (6 - i))) & 7
            if v:
                score.append(v - 1)
        out.append((tuple(score), tuple(ranks)))
    return out


def rank_of_tuple(score, ranks):
    """rank id of a (score, card_ranks) tuple"""
    return class_tuples().index((tuple(score), tuple(ranks)))


def rank7(cards):
    a = ids(cards)
    assert len(a) == 7
    return lib().oracle_rank7(_p(a))


def type7(cards):
    """hand type index 0..8 of a 7-card hand (position in CAT_NAMES)"""
    a = ids(cards)
    return lib().oracle_type7(_p(a))


def rank7_batch(cards):
    a = np.ascontiguousarray(cards, dtype=np.uint8).reshape(-1, 7)
    out = np.empty(len(a), dtype=np.uint16)
    lib().oracle_rank7_batch(_p(a), len(a), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    return out


N_HANDS_7 = 133784560


def colex_range(first, count, want_ranks=True, fast=True):
    """(ranks uint16[count] or None, sums int64[11]) of hands first..first+count-1 in colexicographic order.
    sums = [sum of ids, position-weighted sum, census of the 9 hand types]."""
    ranks = np.empty(count, dtype=np.uint16) if want_ranks else None
    sums = np.zeros(11, dtype=np.int64)
    rc = lib().oracle_colex_range(int(first), int(count),
                                  ranks.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)) if want_ranks else None,
                                  sums.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 1 if fast else 0)
    if rc:
        raise ValueError("range outside C(52,7)")
    return ranks, sums


def colex_checksums(chunk, threads=8):
    """Per-chunk sums over ALL C(52,7) hands, computed on `threads` threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    lib().oracle_fast_init()
    starts = list(range(0, N_HANDS_7, chunk))
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda s: colex_range(s, min(chunk, N_HANDS_7 - s), want_ranks=False)[1], starts))
    return np.stack(res)


def get_winner(holes, board):
    h = np.ascontiguousarray([ids(x) for x in holes], dtype=np.uint8)
    b = ids(board)
    t = ctypes.c_int(0)
    w = lib().oracle_get_winner(_p(h), len(h), _p(b), ctypes.byref(t))
    return w, CAT_NAMES[t.value]


def mc_reference(hero, board, players, runs, seed):
    """Reference dealer + evaluator under np.random.seed(seed): dict(wins, passes, win_types, next_randint)."""
    h, b = ids(hero), ids(board)
    out = (ctypes.c_int64 * 12)()
    rc = lib().oracle_mc_reference(_p(h), _p(b) if len(b) else None, len(b), int(players), int(runs), int(seed), out)
    if rc:
        raise ValueError("oracle_mc_reference rc=%d" % rc)
    return {"wins": out[0], "passes": out[1], "win_types": {CAT_NAMES[i]: out[2 + i] for i in range(9) if out[2 + i]},
            "next_randint": out[11]}


def mc_uniform(hero, board, players, runs, seed):
    h, b = ids(hero), ids(board)
    out = (ctypes.c_int64 * 2)()
    rc = lib().oracle_mc_uniform(_p(h), _p(b) if len(b) else None, len(b), int(players), int(runs), int(seed), out)
    if rc:
        raise ValueError("oracle_mc_uniform rc=%d" % rc)
    return out[0], out[1]


def hand_class(c1, c2):
    return lib().oracle_hand_class(int(c1), int(c2))


def class_index(name):
    """Class number (suited hi*13+lo, offsuit/pairs lo*13+hi) of a spelling like 'KAS', 'AKO', 'QQ'; None if the
    reference's range test (montecarlo_python.py:24-34) could never produce that string."""
    if len(name) == 2:
        return RANKS.index(name[0]) * 14 if name[0] == name[1] and name[0] in RANKS else None
    if len(name) == 3 and name[0] in RANKS and name[1] in RANKS and name[0] != name[1] and name[2] in "SO":
        a, b = RANKS.index(name[0]), RANKS.index(name[1])
        hi, lo = max(a, b), min(a, b)
        return hi * 13 + lo if name[2] == "S" else lo * 13 + hi
    return None


def class_mask(names):
    m = [0, 0, 0]
    for n in names:
        k = class_index(n)
        if k is not None:
            m[k >> 6] |= 1 << (k & 63)
    return (ctypes.c_uint64 * 3)(*m)


def _ranges_call(fn, hero, hero_classes, board, players, runs, seed, opp_classes, ghost, known=()):
    b = ids(board)
    h = ids(hero) if hero_classes is None else None
    g = ids(ghost) if ghost else None
    k = ids([c for hand in known for c in hand]) if known else None
    out = (ctypes.c_int64 * 12)()
    rc = fn(_p(h) if h is not None else None, class_mask(hero_classes) if hero_classes is not None else None,
            _p(b) if len(b) else None, len(b), int(players), int(runs), int(seed), class_mask(opp_classes),
            _p(g) if g is not None else None, _p(k) if k is not None else None, len(known), out)
    if rc:
        raise ValueError("oracle ranges call rc=%d" % rc)
    return out


def mc_reference_ranges(hero, board, players, runs, seed, opp_classes, hero_classes=None, ghost=None, known=()):
    """run_montecarlo with an opponent range (set of class spellings), optionally a hero range, ghost cards and further
    known hands (`known`: the entries of player_card_list after the hero's), under np.random.seed(seed):
    dict(wins, passes, win_types, next_randint).  `hero` is ignored when hero_classes is given."""
    out = _ranges_call(lib().oracle_mc_reference_ranges, hero, hero_classes, board, players, runs, seed, opp_classes, ghost,
                       known)
    return {"wins": out[0], "passes": out[1], "win_types": {CAT_NAMES[i]: out[2 + i] for i in range(9) if out[2 + i]},
            "next_randint": out[11]}


def mc_uniform_ranges(hero, board, players, runs, seed, opp_classes, hero_classes=None, ghost=None, known=()):
    """Unbiased dealing with ranges: (wins_strict, ties, attempts)."""
    out = _ranges_call(lib().oracle_mc_uniform_ranges, hero, hero_classes, board, players, runs, seed, opp_classes, ghost,
                       known)
    return out[0], out[1], out[2]


def enum_headsup(hero, board):
    h, b = ids(hero), ids(board)
    out = (ctypes.c_int64 * 3)()
    lib().oracle_enum_headsup(_p(h), _p(b) if len(b) else None, len(b), out)
    return tuple(out)


def enum_river_multi(hero, board, players):
    h, b = ids(hero), ids(board)
    out = (ctypes.c_int64 * 3)()
    rc = lib().oracle_enum_river_multi(_p(h), _p(b), int(players), out)
    if rc:
        raise ValueError("unsupported")
    return tuple(out)


def enum_reference_headsup(hero, board):
    h, b = ids(hero), ids(board)
    out = (ctypes.c_int64 * 2)()
    rc = lib().oracle_enum_reference_headsup(_p(h), _p(b), len(b), out)
    if rc:
        raise ValueError("unsupported")
    return tuple(out)


# ---- the reference's own C++ (oracle/_ref), when it was built in the container -------------------------------------
def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libnpk_ref.so"))


def ref():
    global _ref
    if _ref is None:
        R = ctypes.CDLL(os.path.join(HERE, "_ref", "libnpk_ref.so"))
        R.ref_montecarlo.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        R.ref_montecarlo.restype = ctypes.c_double
        R.ref_eval_best_hand.argtypes = [ctypes.c_char_p]
        R.ref_calc_score.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                     ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p]
        R.ref_montecarlo_batch.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                           ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_double)]
        R.ref_montecarlo_batch.restype = None
        _ref = R
    return _ref


def _strs(cards):
    return " ".join(c if isinstance(c, str) else card_str(int(c)) for c in cards).encode()


def ref_montecarlo(hero, board, players, iterations):
    return ref().ref_montecarlo(_strs(hero), _strs(board), int(players), int(iterations))


def ref_calc_score(cards):
    sc = (ctypes.c_int * 8)()
    rk = (ctypes.c_int * 9)()
    ns, nr = ctypes.c_int(0), ctypes.c_int(0)
    ty = ctypes.create_string_buffer(32)
    rc = ref().ref_calc_score(_strs(cards), sc, ctypes.byref(ns), rk, ctypes.byref(nr), ty)
    if rc:
        raise RuntimeError("Card Type error!")
    return tuple(sc[:ns.value]), tuple(rk[:nr.value]), ty.value.decode()


def ref_eval_best_hand(hands):
    return ref().ref_eval_best_hand(b";".join(_strs(h) for h in hands))


def ref_montecarlo_batch(heroes, boards, players, iterations, threads):
    n = len(heroes)
    hs = (ctypes.c_char_p * n)(*[_strs(h) for h in heroes])
    bs = (ctypes.c_char_p * n)(*[_strs(b) for b in boards])
    ps = (ctypes.c_int * n)(*[int(p) for p in players])
    out = (ctypes.c_double * n)()
    ref().ref_montecarlo_batch(hs, bs, ps, n, int(iterations), int(threads), out)
    return list(out)
