"""Starting-hand classes and opponent ranges (reference tools/montecarlo_python.py:24-34, :36-112).

A class is an unordered pair of ranks plus suitedness.  The reference spells it as two rank characters followed by 'S'
(suited), 'O' (offsuit) or nothing (pairs), and tests both spellings of a drawn hand against the allowed set
(get_two_short_notation, :24-34), so 'AKS' and 'KAS' are the same class, 'AA' is a pair and 'AAO' / 'AK' never match
anything.  libnpk numbers the classes 0..168:  suited hi*13+lo,  offsuit and pairs lo*13+hi  (rank indices in
"23456789TJQKA", hi >= lo) and takes ranges as 169-bit masks (three uint64 words).

PREFLOP_ORDER is the reference's own ranking of the 169 classes: the keys of MonteCarlo.preflop_equities in ascending
order of equity as `sorted(..., key=itemgetter(1))` leaves them (:105; ties keep dict order).  It is data derived from the
reference (tests/golden/make_golden.py regenerates it into tests/golden/preflop_order.json and a test compares both).
"""
import numpy as np

from .cards import CARD_RANKS_ORIGINAL

PREFLOP_ORDER = (
    "23O 24O 26O 34O 25O 27O 23S 36O 35O 37O 38O 24S 27S 28O 25S 26S 46O 47O 34S 45O 29O 48O 37S 35S 36S "
    "39O 56O 28S 57O 49O 38S 2TO 58O 45S 46S 47S 67O 3TO 48S 56S 29S 39S 68O 59O 57S 4TO 49S 2JO 69O 5TO "
    "2TS 78O 58S 6TO 3TS 79O 67S 3JO 59S 68S 4JO 2QO 2JS 69S 3QO 4TS 6JO 5JO 5TS 7TO 89O 6TS 78S 3JS 8TO "
    "7JO 4JS 79S 7TS 2QS 22 4QO 6JS 5QO 5JS 2KO 3KO 89S 8JO 9TO 6QO 4QS 3QS 7QO 7JS 4KO 8TS 5KO 9JO 5QS "
    "2KS 33 9TS 8QO 6QS 8JS 7QS 6KO 4KS 9QO 3KS 2AO 8KO TJO 7KO 3AO 8QS 5KS 9JS 44 6KS TQO 2AS 9QS 7KS "
    "9KO JQO 4AO 6AO 5AO TJS 8KS 7AO 3AS TQS TKO 4AS 9AO JQS JKO 9KS 8AO 55 6AS QKO 5AS 7AS 8AS TKS TAO "
    "66 QKS 9AS JKS JAO TAS QAO 77 KAO JAS QAS KAS 88 99 TT JJ QQ KK AA "
).split()

N_CLASSES = 169


def class_index(name):
    """Class number of a spelling the reference's range test can match, else None."""
    if not isinstance(name, str):
        return None
    r = CARD_RANKS_ORIGINAL
    if len(name) == 2:
        if name[0] == name[1] and name[0] in r:
            k = r.index(name[0])
            return k * 13 + k
        return None
    if len(name) == 3 and name[0] in r and name[1] in r and name[0] != name[1] and name[2] in "SO":
        a, b = r.index(name[0]), r.index(name[1])
        hi, lo = max(a, b), min(a, b)
        return hi * 13 + lo if name[2] == "S" else lo * 13 + hi
    return None


def class_of_cards(c1, c2):
    """Class number of two card ids (4*rank + suit)."""
    r1, r2 = int(c1) >> 2, int(c2) >> 2
    hi, lo = max(r1, r2), min(r1, r2)
    return hi * 13 + lo if (int(c1) & 3) == (int(c2) & 3) else lo * 13 + hi


def mask_from_classes(names):
    """169-bit mask (numpy uint64[3]) of the class spellings in `names`; spellings that can never match are ignored,
    exactly as a set lookup in the reference would never hit them."""
    m = [0, 0, 0]
    for n in names:
        k = class_index(n)
        if k is not None:
            m[k >> 6] |= 1 << (k & 63)
    return np.array(m, dtype=np.uint64)


def allowed_classes(opponent_range):
    """get_opponent_allowed_cards_list (:36-112): the top int(169 * range) classes of PREFLOP_ORDER.  Python slicing
    quirk kept: a range below 1/169 gives take_top == 0 and `[-0:]` is the WHOLE list."""
    take_top = int(len(PREFLOP_ORDER) * opponent_range)
    return set(PREFLOP_ORDER[-take_top:])


def opponent_mask(opponent_range):
    """Mask for run_montecarlo's `opponent_range` argument (:194-199): a number selects the top fraction of
    PREFLOP_ORDER, a set is taken as the allowed spellings themselves."""
    if isinstance(opponent_range, (set, frozenset)):
        return mask_from_classes(opponent_range)
    return mask_from_classes(allowed_classes(opponent_range))


FULL_MASK = mask_from_classes(PREFLOP_ORDER)
