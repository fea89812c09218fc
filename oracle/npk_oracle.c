/*
 * npk_oracle.c -- CPU restatement of neuron_poker's Monte-Carlo equity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under neuron_poker_b200/ imports, links or executes this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against fixtures produced by running the
 * unmodified Python reference (tests/golden/make_golden.py):
 *   - oracle_calc_score / oracle_rank7   vs eval_tables.npz (5,034 classes, sha256 of SURVEY A.1-13) and
 *                                          eval_cases.json (14 known-answer cases of tests/test_evaluator.py,
 *                                          26,000 seeded hands, 3,000 showdowns)
 *   - oracle_mc_reference                vs mc_seeded.json: bit-identical win counts, pass counts, win-type counts
 *                                          and RNG stream position under np.random.seed(s)
 *   - oracle_enum_*                      vs enum_golden.json (exact enumeration with the reference evaluator)
 *
 * Card id = 4*rank + suit, ranks "23456789TJQKA", suits "CDHS": the order of MonteCarlo.create_card_deck
 * (reference tools/montecarlo_python.py:114-119).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NCLASS_MAX 8192

/* ------------------------------------------------------------------------------------------------------------------
 * _calc_score  (reference tools/hand_evaluator.py:27-119), restated statement by statement.
 * `score` and `card_ranks` are Python tuples there; here: small arrays with explicit lengths.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int score[8];
    int nscore;
    int ranks[9];
    int nranks;
    int hand_type; /* 0 HighCard 1 Pair 2 TwoPair 3 ThreeOfAKind 4 Straight 5 Flush 6 FullHouse 7 FoufOfAKind 8 SF */
} oracle_score_t;

static void sort_desc(int *a, int n)
{
    for (int i = 1; i < n; i++) {
        int v = a[i], j = i - 1;
        while (j >= 0 && a[j] < v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

static int contains(const int *a, int n, int v)
{
    for (int i = 0; i < n; i++) if (a[i] == v) return 1;
    return 0;
}

/* rcounts + "zip(*sorted((cnt, rank) ...)[::-1])"  (:29-30): distinct ranks ordered by (count, rank) descending */
static void count_sorted(const uint8_t *cards, int n, int *score, int *ranks, int *len)
{
    int cnt[13] = {0};
    for (int i = 0; i < n; i++) cnt[cards[i] >> 2]++;
    int m = 0;
    for (int c = 7; c >= 1; c--)
        for (int r = 12; r >= 0; r--)
            if (cnt[r] == c) { score[m] = c; ranks[m] = r; m++; }
    *len = m;
}

static int tuple_eq(const int *a, int na, const int *b, int nb)
{
    if (na != nb) return 0;
    for (int i = 0; i < na; i++) if (a[i] != b[i]) return 0;
    return 1;
}

static int prefix_eq(const int *a, int na, const int *b, int nb)
{   /* a[0:nb] == b  with Python slice semantics (a shorter slice never equals b) */
    if (na < nb) return 0;
    for (int i = 0; i < nb; i++) if (a[i] != b[i]) return 0;
    return 1;
}

int oracle_calc_score(const uint8_t *cards, int n, oracle_score_t *out)
{
    int score[9], ranks[9], ns, nr;
    count_sorted(cards, n, score, ranks, &ns);
    nr = ns;

    static const int T22111[] = {2, 2, 1, 1, 1}, T211111[] = {2, 1, 1, 1, 1, 1};
    static const int T32[] = {3, 2}, T33[] = {3, 3}, T2221[] = {2, 2, 2, 1};
    int potential_threeofakind = score[0] == 3;                       /* :32 */
    int potential_twopair = tuple_eq(score, ns, T22111, 5);           /* :33 */
    int potential_pair = tuple_eq(score, ns, T211111, 6);             /* :34 */

    if (prefix_eq(score, ns, T32, 2) || prefix_eq(score, ns, T33, 2)) {          /* :36-38 full house */
        nr = 2;
        score[0] = 3; score[1] = 2; ns = 2;
    } else if (prefix_eq(score, ns, T2221, 4)) {                                  /* :39-42 three pair -> two pair */
        int kicker = ranks[2] > ranks[3] ? ranks[2] : ranks[3];
        ranks[2] = kicker; nr = 3;
        score[0] = 2; score[1] = 2; score[2] = 1; ns = 3;
    } else if (score[0] == 4) {                                                   /* :43-46 four of a kind (quirk) */
        int s[9];
        memcpy(s, ranks, sizeof(int) * nr);
        sort_desc(s, nr);
        ranks[0] = s[0]; ranks[1] = s[1]; nr = 2;
        score[0] = 4; ns = 1;
    } else if (ns >= 5) {                                                         /* :47-83 */
        int straight = 0, flush;
        int s[9], n_s;
        if (contains(ranks, nr, 12)) ranks[nr++] = -1;                            /* :49-50 */
        memcpy(s, ranks, sizeof(int) * nr); n_s = nr;
        sort_desc(s, n_s);                                                        /* :51 */
        for (int i = 0; i < n_s - 4; i++) {                                       /* :52-58 */
            straight = (s[i] - s[i + 4]) == 4;
            if (straight) {
                for (int k = 0; k < 5; k++) ranks[k] = s[i + k];
                nr = 5;
                break;
            }
        }
        int sc[4] = {0};                                                          /* :61-62 */
        for (int i = 0; i < n; i++) sc[cards[i] & 3]++;
        int mx = 0;
        for (int k = 0; k < 4; k++) if (sc[k] > mx) mx = sc[k];
        flush = mx >= 5;
        if (flush) {
            int fs = 0;
            for (fs = 0; fs < 4; fs++) if (sc[fs] >= 5) break;                    /* :64-66 first suit in CDHS order */
            uint8_t fh[7]; int nf = 0;
            for (int i = 0; i < n; i++) if ((cards[i] & 3) == fs) fh[nf++] = cards[i];   /* :68 */
            int fscore[9];
            count_sorted(fh, nf, fscore, ranks, &nr);                             /* :69-70 */
            sort_desc(ranks, nr);                                                 /* :71-72 */
            if (contains(ranks, nr, 12) && !contains(ranks, nr, -1)) ranks[nr++] = -1;   /* :75-76 */
            for (int i = 0; i < nr - 4; i++) {                                    /* :77-80; `straight` carries over */
                straight = (ranks[i] - ranks[i + 4]) == 4;
                if (straight) break;
            }
        }
        /* :83  score = ([(1,), (3,1,2)], [(3,1,3), (5,)])[flush][straight] */
        if (!flush && !straight) { score[0] = 1; ns = 1; }
        else if (!flush && straight) { score[0] = 3; score[1] = 1; score[2] = 2; ns = 3; }
        else if (flush && !straight) { score[0] = 3; score[1] = 1; score[2] = 3; ns = 3; }
        else { score[0] = 5; ns = 1; }
    }

    if (ns == 1 && score[0] == 1 && potential_threeofakind) { score[0] = 3; score[1] = 1; ns = 2; }        /* :85-86 */
    else if (ns == 1 && score[0] == 1 && potential_twopair) { score[0] = 2; score[1] = 2; score[2] = 1; ns = 3; }
    else if (ns == 1 && score[0] == 1 && potential_pair) { score[0] = 2; score[1] = 1; score[2] = 1; ns = 3; }

    int ht;
    static const int T313[] = {3, 1, 3}, T312[] = {3, 1, 2}, T31[] = {3, 1}, T22[] = {2, 2};
    if (score[0] == 5) ht = 8;                                                    /* :92-93 SF: not truncated */
    else if (score[0] == 4) ht = 7;
    else if (prefix_eq(score, ns, T32, 2)) ht = 6;
    else if (prefix_eq(score, ns, T313, 3)) { ht = 5; if (nr > 5) nr = 5; }
    else if (prefix_eq(score, ns, T312, 3)) { ht = 4; if (nr > 5) nr = 5; }
    else if (prefix_eq(score, ns, T31, 2)) { ht = 3; if (nr > 3) nr = 3; }
    else if (prefix_eq(score, ns, T22, 2)) { ht = 2; if (nr > 3) nr = 3; }
    else if (score[0] == 2) { ht = 1; if (nr > 4) nr = 4; }
    else if (score[0] == 1) { ht = 0; if (nr > 5) nr = 5; }
    else return -1;                                                               /* :116-117 'Card Type error!' */

    memcpy(out->score, score, sizeof(int) * ns); out->nscore = ns;
    memcpy(out->ranks, ranks, sizeof(int) * nr); out->nranks = nr;
    out->hand_type = ht;
    return 0;
}

/* Order-preserving 64-bit key of the Python tuple (score, card_ranks): tuples compare lexicographically and a proper
 * prefix is smaller, so absent entries encode as 0 and present ones as value+1 (score) / value+2 (ranks, -1 -> 1). */
uint64_t oracle_key(const oracle_score_t *s)
{
    uint64_t k = 0;
    for (int i = 0; i < 7; i++) k = (k << 3) | (uint64_t)(i < s->nscore ? s->score[i] + 1 : 0);
    for (int i = 0; i < 8; i++) k = (k << 4) | (uint64_t)(i < s->nranks ? s->ranks[i] + 2 : 0);
    return k;
}

uint64_t oracle_key_cards(const uint8_t *cards, int n)
{
    oracle_score_t s;
    if (oracle_calc_score(cards, n, &s)) return 0;
    return oracle_key(&s);
}

/* Flat export for ctypes: out = [hand_type, nscore, score[0..7], nranks, ranks[0..8]] (20 ints) */
int oracle_calc_score_flat(const uint8_t *cards, int n, int *out)
{
    oracle_score_t s;
    memset(&s, 0, sizeof s);
    int rc = oracle_calc_score(cards, n, &s);
    if (rc) return rc;
    out[0] = s.hand_type; out[1] = s.nscore;
    for (int i = 0; i < 8; i++) out[2 + i] = i < s.nscore ? s.score[i] : 0;
    out[10] = s.nranks;
    for (int i = 0; i < 9; i++) out[11 + i] = i < s.nranks ? s.ranks[i] : -2;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * rank_id: index of (score, card_ranks) in the ascending list of the 5,034 distinct 7-card values.
 * The class list is rebuilt here from the oracle's own calc_score over the whole key space (49,205 rank histograms
 * with a non-flush suit assignment + 4,719 flush masks), exactly the way make_golden.py does with the reference.
 * ---------------------------------------------------------------------------------------------------------------- */
static uint64_t g_classes[NCLASS_MAX];
static int g_nclasses = 0;

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

static void hist_rec(int r, int left, int *h, uint64_t *keys, int *nk)
{
    if (r == 12) {
        if (left > 4) return;
        h[12] = left;
        uint8_t cards[7]; int k = 0;
        for (int q = 0; q < 13; q++)
            for (int j = 0; j < h[q]; j++) { cards[k] = (uint8_t)(4 * q + ((k) & 3)); k++; }
        keys[(*nk)++] = oracle_key_cards(cards, 7);
        return;
    }
    for (int c = 0; c <= 4 && c <= left; c++) { h[r] = c; hist_rec(r + 1, left - c, h, keys, nk); }
}

int oracle_build_classes(void)
{
    if (g_nclasses) return g_nclasses;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * 60000);
    int nk = 0, h[13];
    hist_rec(0, 7, h, keys, &nk);
    for (int mask = 0; mask < 8192; mask++) {
        int pc = __builtin_popcount(mask);
        if (pc < 5 || pc > 7) continue;
        uint8_t cards[7]; int k = 0;
        for (int r = 0; r < 13; r++) if (mask >> r & 1) cards[k++] = (uint8_t)(4 * r);   /* suit C */
        for (int r = 0; k < 7; r++) cards[k++] = (uint8_t)(4 * r + 1);                   /* 2D, 3D pads */
        keys[nk++] = oracle_key_cards(cards, 7);
    }
    qsort(keys, nk, sizeof(uint64_t), cmp_u64);
    int m = 0;
    for (int i = 0; i < nk; i++) if (i == 0 || keys[i] != keys[i - 1]) g_classes[m++] = keys[i];
    free(keys);
    g_nclasses = m;
    return m;
}

int oracle_class_keys(uint64_t *out, int cap)
{
    int n = oracle_build_classes();
    for (int i = 0; i < n && i < cap; i++) out[i] = g_classes[i];
    return n;
}

int oracle_rank_of_key(uint64_t key)
{
    int lo = 0, hi = oracle_build_classes() - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (g_classes[mid] == key) return mid;
        if (g_classes[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

int oracle_rank7(const uint8_t *cards) { return oracle_rank_of_key(oracle_key_cards(cards, 7)); }

void oracle_rank7_batch(const uint8_t *cards, int64_t n, uint16_t *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = (uint16_t)oracle_rank7(cards + 7 * i);
}

/* hand type (0..8) of a 7-card hand */
int oracle_type7(const uint8_t *cards)
{
    oracle_score_t s;
    if (oracle_calc_score(cards, 7, &s)) return -1;
    return s.hand_type;
}

/* eval_best_hand / get_winner (hand_evaluator.py:9-24): stable descending sort -> first index among equal bests */
int oracle_get_winner(const uint8_t *holes /* [n][2] */, int n, const uint8_t *board /* [5] */, int *type_out)
{
    int best = -1; uint64_t bk = 0; int bt = 0;
    for (int p = 0; p < n; p++) {
        uint8_t c[7] = {holes[2 * p], holes[2 * p + 1], board[0], board[1], board[2], board[3], board[4]};
        oracle_score_t s;
        oracle_calc_score(c, 7, &s);
        uint64_t k = oracle_key(&s);
        if (best < 0 || k > bk) { best = p; bk = k; bt = s.hand_type; }
    }
    if (type_out) *type_out = bt;
    return best;
}

/* ------------------------------------------------------------------------------------------------------------------
 * numpy legacy RandomState: MT19937 (init_genrand seeding as np.random.seed(int)) and randint(low, high) for
 * ranges < 2^32 = masked rejection on 32-bit outputs (numpy/random/src/distributions: buffered_bounded_masked_uint32).
 * The reference draws every index with np.random.randint (montecarlo_python.py:169-170, 188).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int pos; } oracle_mt_t;

void oracle_mt_seed(oracle_mt_t *s, uint32_t seed)
{
    for (int i = 0; i < 624; i++) {
        s->mt[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    s->pos = 624;
}

static uint32_t mt_next(oracle_mt_t *s)
{
    if (s->pos >= 624) {
        uint32_t *mt = s->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->pos = 0;
    }
    uint32_t y = s->mt[s->pos++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

/* np.random.randint(low, high): uniform on [low, high) */
int64_t oracle_randint(oracle_mt_t *s, int64_t low, int64_t high)
{
    uint64_t rng = (uint64_t)(high - 1 - low);
    if (rng == 0) return low;
    uint32_t mask = (uint32_t)rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = mt_next(s) & mask; } while (v > rng);
    return low + (int64_t)v;
}

/* ------------------------------------------------------------------------------------------------------------------
 * MonteCarlo.run_montecarlo with opponent_range=1, ghost_cards='' and the wall-clock cut-off disabled
 * (reference tools/montecarlo_python.py:191-252; dealing :121-189).  REFERENCE dealing, bit-exact under a seed.
 *   out[0]=wins  out[1]=passes  out[2..10]=win-type counts (index = hand_type)  out[11]=next randint(0,1000000)
 * ---------------------------------------------------------------------------------------------------------------- */
static int list_pop(uint8_t *deck, int *n, int idx)
{
    int v = deck[idx];
    memmove(deck + idx, deck + idx + 1, (size_t)(*n - idx - 1));
    (*n)--;
    return v;
}

static int list_index(const uint8_t *deck, int n, int card)
{
    for (int i = 0; i < n; i++) if (deck[i] == card) return i;
    return -1;
}

int oracle_mc_reference(const uint8_t *hero, const uint8_t *board, int nboard, int players, int64_t runs,
                        uint32_t seed, int64_t *out)
{
    oracle_mt_t rng;
    oracle_mt_seed(&rng, seed);
    oracle_build_classes();
    int64_t wins = 0, passes = 0, types[9] = {0};
    if (players < 1) return -1;                                  /* reference: IndexError on hands[winner] */
    for (int64_t m = 0; m < runs; m++) {
        uint8_t deck[52]; int n = 52;                            /* :212 copy(OriginalDeck) */
        for (int i = 0; i < 52; i++) deck[i] = (uint8_t)i;
        uint8_t hole[10][2]; uint8_t table[5]; int nt = 0;
        for (int i = 0; i < nboard; i++) {                       /* :126-128 ValueError if not in deck */
            int ix = list_index(deck, n, board[i]);
            if (ix < 0) return -2;
            table[nt++] = (uint8_t)list_pop(deck, &n, ix);
        }
        hole[0][0] = hero[0]; hole[0][1] = hero[1];              /* :150-152 */
        for (int k = 0; k < 2; k++) {                            /* :154-161 failures swallowed */
            int ix = list_index(deck, n, hero[k]);
            if (ix >= 0) list_pop(deck, &n, ix);
        }
        for (int p = 1; p < players; p++) {                      /* :165-181 */
            int64_t i1, i2;
            for (;;) {
                passes++;
                i1 = oracle_randint(&rng, 0, n);
                i2 = oracle_randint(&rng, 0, n - 1);
                if (i1 != i2) break;                             /* range test always passes with opponent_range=1 */
            }
            hole[p][0] = (uint8_t)list_pop(deck, &n, (int)i1);
            hole[p][1] = (uint8_t)list_pop(deck, &n, (int)i2);
        }
        while (nt < 5)                                           /* :185-189 never the last card */
            table[nt++] = (uint8_t)list_pop(deck, &n, (int)oracle_randint(&rng, 0, n - 1));
        int wt;
        int winner = oracle_get_winner(&hole[0][0], players, table, &wt);   /* :222-223 */
        if (winner == 0) { wins++; types[wt]++; }                /* :226-231 */
    }
    out[0] = wins; out[1] = passes;
    for (int i = 0; i < 9; i++) out[2 + i] = types[i];
    out[11] = oracle_randint(&rng, 0, 1000000);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Ranges (reference tools/montecarlo_python.py:24-34 get_two_short_notation, :36-112 allowed list, :136-148 hero range,
 * :165-181 opponent range test, :206-208 ghost cards).
 * A starting-hand class is an unordered rank pair plus suitedness; the reference spells it rank chars + 'S' / 'O' / ''
 * (pairs) and tests both spellings, so membership only depends on the class.  Class number used by every checker and
 * by libnpk:  suited hi*13+lo,  offsuit and pairs lo*13+hi  (hi >= lo rank indices 0..12)  -> 169-bit masks uint64[3].
 * ---------------------------------------------------------------------------------------------------------------- */
int oracle_hand_class(int c1, int c2)
{
    int r1 = c1 >> 2, r2 = c2 >> 2, hi = r1 > r2 ? r1 : r2, lo = r1 > r2 ? r2 : r1;
    return ((c1 & 3) == (c2 & 3)) ? hi * 13 + lo : lo * 13 + hi;
}

static int class_allowed(const uint64_t *mask, int c1, int c2)
{
    int k = oracle_hand_class(c1, c2);
    return (int)(mask[k >> 6] >> (k & 63) & 1u);
}

#define ORACLE_MAX_ATTEMPTS 100000000LL   /* the reference would spin forever on an unsatisfiable range */

/* MonteCarlo.run_montecarlo with an opponent range, optionally a hero RANGE (player_card_list[0] is a set, :136-148)
 * and ghost cards (:206-208).  REFERENCE dealing, bit-exact under np.random.seed(seed).  Quirks kept:
 *   - the range test looks at deck[i1], deck[i2] BEFORE anything is popped (:173-174) while the opponent receives
 *     deck.pop(i1) and then deck.pop(i2) from the shortened list (:178-179): for i2 >= i1 the tested second card and
 *     the dealt one differ;
 *   - a hero drawn from a range keeps exactly the tested cards (:146-148) and is removed by value (:154-161).
 * out layout as oracle_mc_reference.  Returns -3 when a draw exceeds ORACLE_MAX_ATTEMPTS. */
int oracle_mc_reference_ranges(const uint8_t *hero, const uint64_t *hero_mask, const uint8_t *board, int nboard,
                               int players, int64_t runs, uint32_t seed, const uint64_t *opp_mask, const uint8_t *ghost,
                               const uint8_t *known, int n_known, int64_t *out)
{
    oracle_mt_t rng;
    oracle_mt_seed(&rng, seed);
    oracle_build_classes();
    int64_t wins = 0, passes = 0, types[9] = {0};
    if (players < 1) return -1;
    uint8_t original[52]; int n0 = 52;
    for (int i = 0; i < 52; i++) original[i] = (uint8_t)i;
    if (ghost) {                                                 /* :206-208 */
        for (int k = 0; k < 2; k++) {
            int ix = list_index(original, n0, ghost[k]);
            if (ix < 0) return -2;
            list_pop(original, &n0, ix);
        }
    }
    for (int64_t m = 0; m < runs; m++) {
        uint8_t deck[52]; int n = n0;
        memcpy(deck, original, 52);
        uint8_t hole[10][2]; uint8_t table[5]; int nt = 0;
        for (int i = 0; i < nboard; i++) {
            int ix = list_index(deck, n, board[i]);
            if (ix < 0) return -2;
            table[nt++] = (uint8_t)list_pop(deck, &n, ix);
        }
        if (hero_mask) {                                         /* :136-148 */
            int64_t i1, i2, tries = 0;
            for (;;) {
                passes++;
                if (++tries > ORACLE_MAX_ATTEMPTS) return -3;
                i1 = oracle_randint(&rng, 0, n);
                i2 = oracle_randint(&rng, 0, n - 1);
                if (i1 != i2 && class_allowed(hero_mask, deck[i1], deck[i2])) break;
            }
            hole[0][0] = deck[i1]; hole[0][1] = deck[i2];
        } else {
            hole[0][0] = hero[0]; hole[0][1] = hero[1];
        }
        for (int k = 0; k < 2; k++) {
            int ix = list_index(deck, n, hole[0][k]);
            if (ix >= 0) list_pop(deck, &n, ix);
        }
        /* further entries of player_card_list: hands that are known (:132-163), removed by value like the hero's
         * (a failed removal is swallowed, :154-161) */
        for (int f = 0; f < n_known; f++) {
            for (int k = 0; k < 2; k++) {
                hole[1 + f][k] = known[2 * f + k];
                int ix = list_index(deck, n, known[2 * f + k]);
                if (ix >= 0) list_pop(deck, &n, ix);
            }
        }
        for (int p = 1 + n_known; p < players; p++) {            /* :165-181 */
            int64_t i1, i2, tries = 0;
            for (;;) {
                passes++;
                if (++tries > ORACLE_MAX_ATTEMPTS) return -3;
                i1 = oracle_randint(&rng, 0, n);
                i2 = oracle_randint(&rng, 0, n - 1);
                if (i1 != i2 && class_allowed(opp_mask, deck[i1], deck[i2])) break;
            }
            hole[p][0] = (uint8_t)list_pop(deck, &n, (int)i1);
            hole[p][1] = (uint8_t)list_pop(deck, &n, (int)i2);
        }
        while (nt < 5)
            table[nt++] = (uint8_t)list_pop(deck, &n, (int)oracle_randint(&rng, 0, n - 1));
        int wt;
        int winner = oracle_get_winner(&hole[0][0], players, table, &wt);
        if (winner == 0) { wins++; types[wt]++; }
    }
    out[0] = wins; out[1] = passes;
    for (int i = 0; i < 9; i++) out[2 + i] = types[i];
    out[11] = oracle_randint(&rng, 0, 1000000);
    return 0;
}

/* The unbiased counterpart used to check libnpk's UNIFORM mode with ranges (the C++ sibling has no ranges): hero (if
 * drawn from a range) and every opponent receive two distinct cards drawn uniformly from the remaining deck, redrawn
 * until their class is allowed; the board is uniform.  out[0]=wins_strict out[1]=ties out[2]=attempts. */
int oracle_mc_uniform_ranges(const uint8_t *hero, const uint64_t *hero_mask, const uint8_t *board, int nboard,
                             int players, int64_t runs, uint32_t seed, const uint64_t *opp_mask, const uint8_t *ghost,
                             const uint8_t *known_cards, int n_known, int64_t *out)
{
    oracle_mt_t rng;
    oracle_mt_seed(&rng, seed);
    oracle_build_classes();
    int64_t wins = 0, ties = 0, attempts = 0;
    for (int64_t m = 0; m < runs; m++) {
        uint8_t deck[52]; int n = 0;
        for (int c = 0; c < 52; c++) {
            int known = 0;
            if (!hero_mask) known |= (c == hero[0] || c == hero[1]);
            if (ghost) known |= (c == ghost[0] || c == ghost[1]);
            for (int i = 0; i < nboard; i++) known |= (c == board[i]);
            for (int f = 0; f < 2 * n_known; f++) known |= (c == known_cards[f]);
            if (!known) deck[n++] = (uint8_t)c;
        }
        uint8_t hole[10][2];
        int first = hero_mask ? 0 : 1;
        if (!hero_mask) { hole[0][0] = hero[0]; hole[0][1] = hero[1]; }
        for (int f = 0; f < n_known; f++) { hole[1 + f][0] = known_cards[2 * f]; hole[1 + f][1] = known_cards[2 * f + 1]; }
        for (int p = first; p < players; p++) {
            if (p >= 1 && p <= n_known) continue;                /* a known hand */
            const uint64_t *mk = p == 0 ? hero_mask : opp_mask;
            int64_t tries = 0;
            int i1, i2;
            for (;;) {
                attempts++;
                if (++tries > ORACLE_MAX_ATTEMPTS) return -3;
                i1 = (int)oracle_randint(&rng, 0, n);
                i2 = (int)oracle_randint(&rng, 0, n - 1);
                if (i2 >= i1) i2++;
                if (class_allowed(mk, deck[i1], deck[i2])) break;
            }
            hole[p][0] = deck[i1]; hole[p][1] = deck[i2];
            int a = i1 > i2 ? i1 : i2, b = i1 > i2 ? i2 : i1;
            list_pop(deck, &n, a); list_pop(deck, &n, b);
        }
        uint8_t table[5]; int nt = 0;
        for (int i = 0; i < nboard; i++) table[nt++] = board[i];
        while (nt < 5) table[nt++] = (uint8_t)list_pop(deck, &n, (int)oracle_randint(&rng, 0, n));
        uint8_t h7[7] = {hole[0][0], hole[0][1], table[0], table[1], table[2], table[3], table[4]};
        int hv = oracle_rank7(h7), best = -1;
        for (int p = 1; p < players; p++) {
            h7[0] = hole[p][0]; h7[1] = hole[p][1];
            int v = oracle_rank7(h7);
            if (v > best) best = v;
        }
        if (hv > best) wins++; else if (hv == best) ties++;
    }
    out[0] = wins; out[1] = ties; out[2] = attempts;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Uniform dealing (the C++ sibling's semantics, tools/montecarlo_cpp/Montecarlo.cpp:240-259, 293-312: shuffle the
 * remaining cards, deal consecutively to opponents then board; ties count as wins).  The sibling seeds a fresh
 * mt19937_64 from std::random_device every trial, so no stream can be matched -- statistical parity only; this port
 * uses a Fisher-Yates partial shuffle driven by the MT19937 above.  out[0]=wins_strict out[1]=ties.
 * ---------------------------------------------------------------------------------------------------------------- */
int oracle_mc_uniform(const uint8_t *hero, const uint8_t *board, int nboard, int players, int64_t runs,
                      uint32_t seed, int64_t *out)
{
    oracle_mt_t rng;
    oracle_mt_seed(&rng, seed);
    oracle_build_classes();
    uint8_t base[52]; int nb = 0;
    if (nboard < 3) nboard = 0;                                  /* Montecarlo.cpp:242-243 */
    for (int c = 0; c < 52; c++) {
        int known = (c == hero[0] || c == hero[1]);
        for (int i = 0; i < nboard; i++) known |= (c == board[i]);
        if (!known) base[nb++] = (uint8_t)c;
    }
    int need = 2 * (players - 1) + (5 - nboard);
    if (need > nb) return -1;
    int64_t wins = 0, ties = 0;
    for (int64_t m = 0; m < runs; m++) {
        uint8_t d[52];
        memcpy(d, base, (size_t)nb);
        for (int k = 0; k < need; k++) {
            int j = k + (int)oracle_randint(&rng, 0, nb - k);
            uint8_t t = d[k]; d[k] = d[j]; d[j] = t;
        }
        uint8_t table[5]; int nt = 0, k = 2 * (players - 1);
        for (int i = 0; i < nboard; i++) table[nt++] = board[i];
        while (nt < 5) table[nt++] = d[k++];
        uint8_t h7[7] = {hero[0], hero[1], table[0], table[1], table[2], table[3], table[4]};
        int hv = oracle_rank7(h7), best = -1;
        for (int p = 1; p < players; p++) {
            h7[0] = d[2 * (p - 1)]; h7[1] = d[2 * (p - 1) + 1];
            int v = oracle_rank7(h7);
            if (v > best) best = v;
        }
        if (hv > best) wins++; else if (hv == best) ties++;
    }
    out[0] = wins; out[1] = ties;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Exact heads-up enumeration (uniform dealing): every completion of the board x every opponent pair.
 * out = {win, tie, lose} from the hero's point of view.  nboard in {3,4,5} finishes in well under a second;
 * nboard == 0 is 2.1e9 matchups -- callers use it only from long-running checks.
 * ---------------------------------------------------------------------------------------------------------------- */
static void enum_rec(uint8_t *rest, int nrest, int start, uint8_t *table, int nt, const uint8_t *hero, int64_t *out)
{
    if (nt == 5) {
        uint8_t h7[7] = {hero[0], hero[1], table[0], table[1], table[2], table[3], table[4]};
        int hv = oracle_rank7(h7);
        int vals[52];
        for (int i = 0; i < nrest; i++) vals[i] = -1;
        for (int a = 0; a < nrest; a++) {
            int ua = 0;
            for (int t = 0; t < 5; t++) ua |= rest[a] == table[t];
            if (ua) continue;
            for (int b = a + 1; b < nrest; b++) {
                int ub = 0;
                for (int t = 0; t < 5; t++) ub |= rest[b] == table[t];
                if (ub) continue;
                h7[0] = rest[a]; h7[1] = rest[b];
                int ov = oracle_rank7(h7);
                if (hv > ov) out[0]++; else if (hv == ov) out[1]++; else out[2]++;
            }
        }
        return;
    }
    for (int i = start; i < nrest; i++) {
        table[nt] = rest[i];
        enum_rec(rest, nrest, i + 1, table, nt + 1, hero, out);
    }
}

int oracle_enum_headsup(const uint8_t *hero, const uint8_t *board, int nboard, int64_t *out)
{
    oracle_build_classes();
    uint8_t rest[52]; int nrest = 0; uint8_t table[5];
    for (int c = 0; c < 52; c++) {
        int known = (c == hero[0] || c == hero[1]);
        for (int i = 0; i < nboard; i++) known |= (c == board[i]);
        if (!known) rest[nrest++] = (uint8_t)c;
    }
    for (int i = 0; i < nboard; i++) table[i] = board[i];
    out[0] = out[1] = out[2] = 0;
    enum_rec(rest, nrest, 0, table, nboard, hero, out);
    return 0;
}

/* Ordered tuples of (players-1) disjoint opponent hands on a complete board (test 17 for players=3). */
int oracle_enum_river_multi(const uint8_t *hero, const uint8_t *board, int players, int64_t *out)
{
    oracle_build_classes();
    uint8_t rest[52]; int nrest = 0;
    for (int c = 0; c < 52; c++) {
        int known = (c == hero[0] || c == hero[1]);
        for (int i = 0; i < 5; i++) known |= (c == board[i]);
        if (!known) rest[nrest++] = (uint8_t)c;
    }
    uint8_t h7[7] = {hero[0], hero[1], board[0], board[1], board[2], board[3], board[4]};
    int hv = oracle_rank7(h7);
    out[0] = out[1] = out[2] = 0;
    if (players == 2) return oracle_enum_headsup(hero, board, 5, out);
    if (players != 3) return -1;
    int np = 0; static int pa[1326], pb[1326], pv[1326];
    for (int a = 0; a < nrest; a++) for (int b = a + 1; b < nrest; b++) {
        h7[0] = rest[a]; h7[1] = rest[b];
        pa[np] = a; pb[np] = b; pv[np] = oracle_rank7(h7); np++;
    }
    for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) {
        if (pa[i] == pa[j] || pa[i] == pb[j] || pb[i] == pa[j] || pb[i] == pb[j]) continue;
        int best = pv[i] > pv[j] ? pv[i] : pv[j];
        if (hv > best) out[0]++; else if (hv == best) out[1]++; else out[2]++;
    }
    return 0;
}

/* Exact expectation of the REFERENCE dealer, heads-up, board of 4 or 5 cards:
 * every (i1 in [0,n), i2 in [0,n-1), i1 != i2) equally likely; missing board card uniform over all but the last.
 * out = {numerator of hero >= opponent, denominator}. */
int oracle_enum_reference_headsup(const uint8_t *hero, const uint8_t *board, int nboard, int64_t *out)
{
    oracle_build_classes();
    if (nboard != 4 && nboard != 5) return -1;
    uint8_t base[52]; int n = 0;
    for (int c = 0; c < 52; c++) {
        int known = (c == hero[0] || c == hero[1]);
        for (int i = 0; i < nboard; i++) known |= (c == board[i]);
        if (!known) base[n++] = (uint8_t)c;
    }
    int64_t num = 0, den = 0;
    for (int i1 = 0; i1 < n; i1++) for (int i2 = 0; i2 < n - 1; i2++) {
        if (i1 == i2) continue;
        uint8_t d[52]; int m = n;
        memcpy(d, base, (size_t)n);
        int c1 = list_pop(d, &m, i1), c2 = list_pop(d, &m, i2);
        uint8_t h[7] = {hero[0], hero[1], board[0], board[1], board[2], board[3], 0};
        uint8_t o[7] = {(uint8_t)c1, (uint8_t)c2, board[0], board[1], board[2], board[3], 0};
        if (nboard == 5) {
            h[6] = o[6] = board[4];
            num += oracle_rank7(h) >= oracle_rank7(o); den++;
        } else {
            for (int j = 0; j < m - 1; j++) {
                h[6] = o[6] = d[j];
                num += oracle_rank7(h) >= oracle_rank7(o); den++;
            }
        }
    }
    out[0] = num; out[1] = den;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Exhaustive enumeration support: all C(52,7) = 133,784,560 hands in colexicographic order (c0 < ... < c6,
 * index = sum_i C(c_i, i+1)), the order of libnpk's npk_rank7_colex.
 * oracle_rank7_fast() is the SAME function as oracle_rank7() tabulated: the rank id of every rank histogram and of
 * every flush mask is computed once with oracle_calc_score (above) and then looked up; tests/test_oracle.py checks
 * fast == slow on seeded hands.
 * ---------------------------------------------------------------------------------------------------------------- */
#define FAST_HASH_SIZE (1 << 17)
static uint32_t g_fast_code[FAST_HASH_SIZE];
static uint16_t g_fast_id[FAST_HASH_SIZE];
static uint16_t g_fast_flush[8192];
static int g_fast_ready = 0;
static const uint32_t POW5[13] = {1, 5, 25, 125, 625, 3125, 15625, 78125, 390625, 1953125, 9765625, 48828125, 244140625};

static void fast_rec(int r, int left, int *h)
{
    if (r == 12) {
        if (left > 4) return;
        h[12] = left;
        uint8_t cards[7]; int k = 0; uint32_t code = 0;
        for (int q = 0; q < 13; q++) {
            code += (uint32_t)h[q] * POW5[q];
            for (int j = 0; j < h[q]; j++) { cards[k] = (uint8_t)(4 * q + (k & 3)); k++; }
        }
        uint32_t slot = (code * 2654435761u) >> 15;
        while (g_fast_code[slot] != 0xFFFFFFFFu) slot = (slot + 1) & (FAST_HASH_SIZE - 1);
        g_fast_code[slot] = code;
        g_fast_id[slot] = (uint16_t)oracle_rank7(cards);
        return;
    }
    for (int c = 0; c <= 4 && c <= left; c++) { h[r] = c; fast_rec(r + 1, left - c, h); }
}

void oracle_fast_init(void)
{
    if (g_fast_ready) return;
    oracle_build_classes();
    memset(g_fast_code, 0xFF, sizeof g_fast_code);
    int h[13];
    fast_rec(0, 7, h);
    for (int mask = 0; mask < 8192; mask++) {
        int pc = __builtin_popcount(mask);
        g_fast_flush[mask] = 0xFFFF;
        if (pc < 5 || pc > 7) continue;
        uint8_t cards[7]; int k = 0;
        for (int r = 0; r < 13; r++) if (mask >> r & 1) cards[k++] = (uint8_t)(4 * r);
        for (int r = 0; k < 7; r++) cards[k++] = (uint8_t)(4 * r + 1);
        g_fast_flush[mask] = (uint16_t)oracle_rank7(cards);
    }
    g_fast_ready = 1;
}

int oracle_rank7_fast(const uint8_t *cards)
{
    uint32_t code = 0, sm[4] = {0, 0, 0, 0};
    int sc[4] = {0, 0, 0, 0};
    for (int i = 0; i < 7; i++) {
        int r = cards[i] >> 2, s = cards[i] & 3;
        code += POW5[r]; sc[s]++; sm[s] |= 1u << r;
    }
    for (int s = 0; s < 4; s++) if (sc[s] >= 5) return g_fast_flush[sm[s]];
    uint32_t slot = (code * 2654435761u) >> 15;
    while (g_fast_code[slot] != code) slot = (slot + 1) & (FAST_HASH_SIZE - 1);
    return g_fast_id[slot];
}

static int64_t binom64(int n, int k)
{
    if (k < 0 || k > n) return 0;
    int64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

/* ranks of hands first..first+count-1 (colex).  Any of out_ranks / sums may be NULL.
 * sums[0] += sum of rank ids, sums[1] += sum of rank_id * ((index mod 65521) + 1), sums[2..10] += hand-type census. */
int oracle_colex_range(int64_t first, int64_t count, uint16_t *out_ranks, int64_t *sums, int use_fast)
{
    if (use_fast) oracle_fast_init(); else oracle_build_classes();
    if (first < 0 || count < 0 || first + count > 133784560LL) return -1;
    static const int type_start[10] = {0, 407, 1877, 2640, 3215, 3225, 4502, 4658, 4736, 5034};
    uint8_t c[7];
    int64_t r = first;
    int hi = 51;
    for (int k = 7; k >= 1; k--) {
        int x = hi;
        while (binom64(x, k) > r) x--;
        r -= binom64(x, k);
        c[k - 1] = (uint8_t)x;
        hi = x - 1;
    }
    for (int64_t i = 0; i < count; i++) {
        int id = use_fast ? oracle_rank7_fast(c) : oracle_rank7(c);
        if (out_ranks) out_ranks[i] = (uint16_t)id;
        if (sums) {
            sums[0] += id;
            sums[1] += (int64_t)id * (((first + i) % 65521) + 1);
            int ty = 0;
            for (int t = 1; t < 9; t++) ty += id >= type_start[t];
            sums[2 + ty]++;
        }
        /* next combination in colex order: increment the lowest position that can move */
        int j = 0;
        while (j < 6 && c[j] + 1 == c[j + 1]) { c[j] = (uint8_t)j; j++; }
        c[j]++;
    }
    return 0;
}
