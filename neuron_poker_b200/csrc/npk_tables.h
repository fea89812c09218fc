// npk_tables.h -- host-side construction of the 7-card rank lookup tables staged into shared memory by the kernels.
//
// The value being tabulated is the reference's hand ordering, tools/hand_evaluator.py:27-119 (`_calc_score`), including
// its two non-standard rules (four-of-a-kind keyed by the two highest ranks present, :43-46; straight flush keyed by
// ALL ranks of the flush suit, :68-80, :92-93).  rank_id = index of the hand's (score, card_ranks) tuple in the
// ascending list of the 5,034 distinct 7-card tuples, so rank ids compare exactly like the tuples.
#pragma once
#include <cstdint>
#include <vector>

namespace npk {

constexpr int kNumRanks = 13;
constexpr int kNumClasses = 5034;
constexpr int kNumHistograms = 49205;
constexpr int kRowShift = 10;            // row = mk >> kRowShift, column = mk & (2^kRowShift - 1), mk = mixed key
constexpr int kMixBits = 23;             // mixed keys live modulo 2^23 (> largest plain key sum 7,825,759)
constexpr uint32_t kMixMul = 0x9E3779B1u; // odd multiplier: mk = (kMixMul * plain key) mod 2^23 scatters the rows
constexpr int kDescShift = 9;            // card descriptor = (mixed rank key << 9) | (16*suit + (12 - rank))
constexpr int kFlushTableSize = 8192;    // indexed by the 13-bit rank mask of the flush suit

// additive rank keys: key(hand) = sum over the 7 cards of kRankKey[rank]; distinct for distinct rank histograms
// (found by tools/find_rank_keys.cpp; verified again in build_tables()).  The tables are indexed by the MIXED key
// mk = (kMixMul * key) mod 2^23 = sum of mixed_rank_key(rank) mod 2^23: still additive per card, still injective
// (kMixMul is odd), but neighbouring hands no longer share a row, so row displacement packs 49,205 keys into
// 49,770 slots instead of > 100,000.
extern const uint32_t kRankKey[kNumRanks];
inline uint32_t mixed_rank_key(int rank) { return (kMixMul * kRankKey[rank]) & ((1u << kMixBits) - 1u); }
// 32-bit card descriptor for card id = 4*rank + suit (reference deck order, montecarlo_python.py:114-119)
// The low 6 bits hold the suit (bits 4-5) and the REVERSED rank 12 - rank (bits 0-3): the device turns a card into its
// contribution to the rank mask of suit fs with one XOR-AND and one clamped shift, 0x1000 >> ((d ^ (fs << 4)) & 63),
// which is 1 << rank when the suit matches and 0 otherwise.
inline uint32_t card_desc(int card) { return (mixed_rank_key(card >> 2) << kDescShift) | (uint32_t)(16 * (card & 3) + 12 - (card >> 2)); }

// first rank_id of each hand type, in the reference's type order:
// HighCard, Pair, TwoPair, ThreeOfAKind, Straight, Flush, FullHouse, FoufOfAKind, StraightFlush, (end)
struct Tables {
    std::vector<uint16_t> value;      // row-displaced rank ids: value[row_offset[mk >> kRowShift] + (mk & mask)]
    std::vector<uint16_t> row_offset; // one entry per row
    std::vector<uint16_t> flush;      // rank id by flush-suit rank mask; 0 where popcount is not 5..7 (never read)
    std::vector<uint64_t> class_key;  // the 5,034 order keys, ascending (index = rank_id)
    uint16_t type_start[10];
    uint32_t max_key;                 // largest plain (unmixed) key sum
};

// Reference hand value of a non-flush 7-card rank histogram / of a flush-suit rank mask as an order-preserving
// 64-bit key: bits 32.. = hand type (0..8), bits 0..31 = up to eight card_ranks entries (4 bits each, value+2,
// most significant first, 0 = absent).
uint64_t order_key_from_histogram(const uint8_t hist[kNumRanks]);
uint64_t order_key_from_flush_mask(uint32_t mask);

// Builds everything; returns an empty string on success or a description of the failed self-check.
const char* build_tables(Tables& out);

// Host evaluation through the built tables (same arithmetic as the device code): used by the table self-check and by
// the CPU-side tests of the table builder; the product's compute path is the CUDA kernels.
uint16_t host_rank7(const Tables& t, const uint8_t cards[7]);

}  // namespace npk
