"""Executable specification of libnpk's two samplers, in plain Python (test infrastructure).

The GPU kernels deal cards from a Philox4x32-10 stream keyed by (seed; trial, query).  This module restates that
procedure card for card, and scores the resulting hands with the ORACLE evaluator (oracle/), so a GPU run can be
checked for bit-identical win / tie / pass / win-type counts -- not only statistically.

  uniform   (npk_kernels.cu equity_uniform_kernel): partial Fisher-Yates over the unseen cards in ascending card id;
            draw k takes index hi32(x * (N-k)) where x is a fresh Philox word for even k and the low product word of
            the previous draw for odd k; the hole left by a draw is filled with the last live element.  Two consecutive
            trials share their Philox blocks (see deal_uniform).
  reference (equity_refdeal_kernel): the Python reference's dealer, montecarlo_python.py:165-189 -- pops at random
            indices of the ORDERED list of unseen cards, sampled without its retry loop; see deal_reference().
  ranges    (equity_ranges_kernel): the generic index-based dealer with class masks, see deal_ranges().
"""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
REFERENCE_BLOCK0 = 0x80000000


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return [c0, c1, c2, c3]


def deal_uniform(seed, query, trial, hole, board, players):
    """equity_uniform_kernel.  Returns (opponent hands [[c1,c2],...], full board [5]).

    A trial needs NW = ceil(D/2) words (two draws per word).  Trials are generated in pairs: pair P = trial >> 1 draws
    ceil(2*NW/4) Philox blocks with counter (P_lo, P_hi, query, block); trial 2P reads words [0, NW), trial 2P+1 words
    [NW, 2*NW)."""
    known = set(hole) | set(board)
    deck = [c for c in range(52) if c not in known]
    n = len(deck)
    nopp = players - 1
    d = 2 * nopp + (5 - len(board))
    nw = (d + 1) // 2
    nblk = (2 * nw + 3) // 4
    key = (seed & MASK, (seed >> 32) & MASK)
    pair, half = trial >> 1, trial & 1
    w = []
    for b in range(nblk):
        w += philox4x32_10((pair & MASK, (pair >> 32) & MASK, query & MASK, b), key)
    w = w[half * nw:(half + 1) * nw]
    out, rem = [], 0
    for k in range(d):
        x = rem if k & 1 else w[k >> 1]
        prod = x * (n - k)
        idx, rem = prod >> 32, prod & MASK
        out.append(deck[idx])
        deck[idx] = deck[n - 1 - k]
    opp = [out[2 * i:2 * i + 2] for i in range(nopp)]
    return opp, list(board) + out[2 * nopp:]


class _Words:
    def __init__(self, seed, query, trial):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.ctr = (trial & MASK, (trial >> 32) & MASK, query & MASK)
        self.blk, self.buf = REFERENCE_BLOCK0, []

    def next(self):
        if not self.buf:
            self.buf = philox4x32_10(self.ctr + (self.blk & MASK,), self.key)
            self.blk += 1
        return self.buf.pop(0)


PASSES_BLOCK0 = 0x40000000
MAX_ATTEMPTS = 1 << 16


def lehmer_slots(raw):
    """Canonical slots of a sequence of pops: raw[k] indexes the ordered list shortened by pops 0..k-1.  Restated the
    way the kernel does it (backward sweep: un-popping pop k shifts every later pop >= raw[k] up by one)."""
    s = list(raw)
    for k in range(len(s) - 2, -1, -1):
        for j in range(k + 1, len(s)):
            if s[j] >= raw[k]:
                s[j] += 1
    return s


def deal_reference(seed, query, trial, hole, board, players):
    """equity_refdeal_kernel.  Returns (opponent hands, full board, passes).

    The reference's dealer (montecarlo_python.py:165-189) pops from the ORDERED list of unseen cards: an opponent takes
    i1 in [0,n), i2 in [0,n-1), retried while i1 == i2 (:169-172), then pop(i1), pop(i2) (:178-179); a board card pops
    index j in [0, n-1) (:188).  The accepted (i1, i2) are uniform over (n-1)^2 pairs, and (a, b) in [0,n-1)^2 ->
    (a + (a >= b), b) is a bijection onto them: no retry.  Word o of the trial gives opponent o's a = hi32(w*(n-1)),
    b = hi32(lo32(w*(n-1))*(n-1)); a board word serves two board cards (index = hi32(x*(n-1)), x = the word, then the low
    product word).  Trials come in pairs sharing Philox blocks (counter (pair, query, 0x80000000 + block)), trial 2P reads
    words [0, NWR), trial 2P+1 words [NWR, 2*NWR), NWR = opponents + ceil(board cards to come / 2).
    `passes` is drawn from the law of the reference's attempt counter: per opponent 1 + a geometric number of failures
    (an attempt fails with probability 1/n), from the pair's block(s) at counter 0x40000000 + block, word
    (trial & 1) * opponents + o: failures = number of times x < floor((2^32-1)/n) holds along x -> x*n."""
    known = set(hole) | set(board)
    deck = [c for c in range(52) if c not in known]
    n0 = len(deck)
    nopp, nb = players - 1, 5 - len(board)
    nwr = nopp + (nb + 1) // 2
    nblk = (2 * nwr + 3) // 4
    key = (seed & MASK, (seed >> 32) & MASK)
    pair, half = trial >> 1, trial & 1
    w = []
    for b in range(nblk):
        w += philox4x32_10((pair & MASK, (pair >> 32) & MASK, query & MASK, (REFERENCE_BLOCK0 + b) & MASK), key)
    w = w[half * nwr:(half + 1) * nwr]
    raw = []
    for o in range(nopp):
        m = n0 - 2 * o - 1
        prod = w[o] * m
        a, b = prod >> 32, ((prod & MASK) * m) >> 32
        raw += [a + (1 if a >= b else 0), b]
    rem = 0
    for c in range(nb):
        m = n0 - 2 * nopp - c - 1
        x = rem if c & 1 else w[nopp + (c >> 1)]
        prod = x * m
        raw.append(prod >> 32)
        rem = prod & MASK
    cards = [deck[s] for s in lehmer_slots(raw)]
    # cross-check against the literal list semantics of the reference
    lst, lit = list(deck), []
    for i in raw:
        lit.append(lst.pop(i))
    assert lit == cards
    passes = 0
    if nopp:
        pw = []
        for b in range((2 * nopp + 3) // 4):
            pw += philox4x32_10((pair & MASK, (pair >> 32) & MASK, query & MASK, (PASSES_BLOCK0 + b) & MASK), key)
        for o in range(nopp):
            n = n0 - 2 * o
            x, tries = pw[half * nopp + o], 1
            while x < MASK // n and tries < MAX_ATTEMPTS:
                tries += 1
                x = (x * n) & MASK
            passes += tries
    opp = [cards[2 * i:2 * i + 2] for i in range(nopp)]
    return opp, list(board) + cards[2 * nopp:], passes


RANGES_BLOCK0 = {"reference": 0x80000000, "uniform": 0xC0000000}


def hand_class(c1, c2):
    """Starting-hand class number: suited hi*13+lo, offsuit and pairs lo*13+hi (rank indices, hi >= lo)."""
    r1, r2 = c1 >> 2, c2 >> 2
    hi, lo = max(r1, r2), min(r1, r2)
    return hi * 13 + lo if (c1 & 3) == (c2 & 3) else lo * 13 + hi


def _allowed(mask, c1, c2):
    k = hand_class(c1, c2)
    return (int(mask[k >> 6]) >> (k & 63)) & 1


def deal_ranges(mode, seed, query, trial, hole, board, players, opp_mask, hero_mask=None, ghost=None, known=()):
    """equity_ranges_kernel: dealing with an opponent range (169-bit mask, three 64-bit words), optionally a hero drawn
    from a range (`hole` ignored) and ghost cards removed from the deck.  Returns (hero, opponents, full board, passes).

    One Philox word per attempt: i1 = hi32(w*n), i2 = hi32(lo32(w*n)*(n-1)) on the ordered list of unseen cards.
      reference: the Python reference, montecarlo_python.py:136-181 -- retry while i1 == i2 or the class of
                 (deck[i1], deck[i2]) is not allowed (both read BEFORE popping); a hero keeps exactly those two cards;
                 an opponent gets deck.pop(i1) then deck.pop(i2) from the shortened list; board card j = hi32(w*(n-1)).
      uniform:   c1 = deck[i1], c2 = (deck without c1)[i2], retry while their class is not allowed; board j = hi32(w*n).
    """
    # `known`: hands of opponents whose cards are known (montecarlo_python.py:132-163): out of the deck, part of the showdown
    seen = set(board) | (set(ghost) if ghost else set()) | (set(hole) if hero_mask is None else set()) | {c for h in known for c in h}
    deck = [c for c in range(52) if c not in seen]
    ws = _Words(seed, query, trial)
    ws.blk = RANGES_BLOCK0[mode]
    passes = 0

    def draw(mask, is_hero):
        nonlocal passes
        n = len(deck)
        for _ in range(MAX_ATTEMPTS):
            passes += 1
            prod = ws.next() * n
            i1, i2 = prod >> 32, ((prod & MASK) * (n - 1)) >> 32
            if mode == "reference":
                if i1 == i2 or not _allowed(mask, deck[i1], deck[i2]):
                    continue
                if is_hero:
                    c1, c2 = deck[i1], deck[i2]
                    deck.remove(c1)
                    deck.remove(c2)
                else:
                    c1 = deck.pop(i1)
                    c2 = deck.pop(i2)
                return [c1, c2]
            c1 = deck[i1]
            c2 = (deck[:i1] + deck[i1 + 1:])[i2]
            if not _allowed(mask, c1, c2):
                continue
            deck.remove(c1)
            deck.remove(c2)
            return [c1, c2]
        raise RuntimeError("range cannot be satisfied")

    hero = list(hole) if hero_mask is None else draw(hero_mask, True)
    opp = [list(h) for h in known] + [draw(opp_mask, False) for _ in range(players - 1 - len(known))]
    full = list(board)
    while len(full) < 5:
        n = len(deck)
        j = (ws.next() * (n - 1 if mode == "reference" else n)) >> 32
        full.append(deck.pop(j))
    return hero, opp, full, passes


FAST_BLOCK0 = {"reference": 0xA0000000, "uniform": 0xE0000000}


def deal_ranges_fast(mode, seed, query, trial, hole, board, players, opp_mask, hero_mask=None, ghost=None, known=()):
    """equity_ranges_fast_kernel (csrc/npk_ranges.cu): the same distribution of dealt cards as deal_ranges without the
    attempt loop over the whole deck.  Per query the unordered pairs (a < b) of initially unseen cards whose class is
    allowed are listed in ascending pair number b*(b-1)/2 + a; a draw takes entry hi32(w * len) and swaps it when bit 31 of
    the low product word is set -> (sa, sb); it is redone when sa or sb has been dealt in this trial or (reference) sb is
    the highest unseen card (the reference's i2 never reaches the last list element).  reference: hero keeps (sa, sb); an
    opponent gets sa and, when sb > sa, the successor of sb among the unseen cards (pop(i1) shifted the list), else sb.
    Board card: index hi32(w * (n-1)) (reference) / hi32(w * n) (uniform) of the ordered unseen cards.  One Philox word per
    attempt / board card, blocks from 0xA0000000 (reference) / 0xE0000000 (uniform).  Returns (hero, opponents, board)."""
    seen = set(board) | (set(ghost) if ghost else set()) | (set(hole) if hero_mask is None else set()) | {c for h in known for c in h}
    deck0 = [c for c in range(52) if c not in seen]
    avail = set(deck0)
    ws = _Words(seed, query, trial)
    ws.blk = FAST_BLOCK0[mode]

    def pair_list(mask):
        return [(a, b) for b in range(1, 52) for a in range(b) if a in avail and b in avail and _allowed(mask, a, b)]

    opp_list = pair_list(opp_mask) if players - 1 - len(known) > 0 else []
    hero_list = pair_list(hero_mask) if hero_mask is not None else []

    def draw(lst, is_hero):
        if not lst:
            raise RuntimeError("range cannot be satisfied")
        for _ in range(MAX_ATTEMPTS):
            prod = ws.next() * len(lst)
            sa, sb = lst[prod >> 32]
            if (prod & MASK) >> 31:
                sa, sb = sb, sa
            if sa not in avail or sb not in avail:
                continue
            if mode == "reference" and sb == max(avail):
                continue
            c1, c2 = sa, sb
            if mode == "reference" and not is_hero and sb > sa:
                c2 = min(c for c in avail if c > sb)
            avail.discard(c1)
            avail.discard(c2)
            return [c1, c2]
        raise RuntimeError("range cannot be satisfied")

    hero = list(hole) if hero_mask is None else draw(hero_list, True)
    opp = [list(h) for h in known] + [draw(opp_list, False) for _ in range(players - 1 - len(known))]
    full = list(board)
    while len(full) < 5:
        n = len(avail)
        j = (ws.next() * (n - 1 if mode == "reference" else n)) >> 32
        c = sorted(avail)[j]
        avail.discard(c)
        full.append(c)
    return hero, opp, full


def run_model(oracle, mode, seed, query, hole, board, players, trials, trial_offset=0, opp_mask=None, hero_mask=None,
              ghost=None, fast=False, known=()):
    """dict(wins, ties, passes, win_types[9]) of the modelled sampler scored by the oracle's rank ids.  With `opp_mask`
    the range dealers (deal_ranges) are modelled, otherwise the plain ones."""
    wins = ties = passes = 0
    types = [0] * 9
    for t in range(trial_offset, trial_offset + trials):
        if opp_mask is not None:
            if fast:
                hole_t, opp, full = deal_ranges_fast(mode, seed, query, t, hole, board, players, opp_mask, hero_mask, ghost, known)
            else:
                hole_t, opp, full, p = deal_ranges(mode, seed, query, t, hole, board, players, opp_mask, hero_mask, ghost, known)
                passes += p
            hv = oracle.rank7(list(hole_t) + full)
            best = max([oracle.rank7(o + full) for o in opp], default=-1)
            if hv > best:
                wins += 1
            elif hv == best:
                ties += 1
            if hv >= best:
                types[oracle.type7(list(hole_t) + full)] += 1
            continue
        if mode == "uniform":
            opp, full = deal_uniform(seed, query, t, hole, board, players)
        else:
            opp, full, p = deal_reference(seed, query, t, hole, board, players)
            passes += p
        hv = oracle.rank7(list(hole) + full)
        best = max([oracle.rank7(o + full) for o in opp], default=-1)
        if hv > best:
            wins += 1
        elif hv == best:
            ties += 1
        if hv >= best:
            types[oracle.type7(list(hole) + full)] += 1
    return {"wins": wins, "ties": ties, "passes": passes, "win_types": types}
