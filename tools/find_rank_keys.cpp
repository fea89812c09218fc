// find_rank_keys.cpp -- offline search for the 13 additive rank keys used by the sm_100a evaluator tables.
//
// Goal: integers k[0..12] (k[0] = 0, ascending) such that the sum of the keys of any multiset of j <= 7 ranks with at
// most 4 copies per rank is unique among the multisets of the same size j.  Then  key(hand) = sum_i k[rank(card_i)]
// identifies the 7-card rank histogram with one integer add per card, and the sum stays below 2^23.
// Greedy: each new key is the smallest integer that keeps all same-size sums distinct.
//
//   g++ -O2 -o /tmp/find_rank_keys tools/find_rank_keys.cpp && /tmp/find_rank_keys
//
// The result is pasted into neuron_poker_b200/csrc/npk_tables.cpp (kRankKey) where npk_init re-verifies injectivity.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

int main()
{
    const int NR = 13, MAXC = 7;
    std::vector<std::vector<uint32_t>> S(MAXC + 1);   // S[j] = sums of j-card multisets over the ranks placed so far
    S[0].push_back(0);
    std::vector<uint32_t> key;
    std::vector<uint32_t> stamp(1u << 25, 0);
    uint32_t tick = 0;
    uint32_t cand = 0;
    for (int r = 0; r < NR; r++) {
        for (;; cand++) {
            bool ok = true;
            for (int j = 1; j <= MAXC && ok; j++) {
                ++tick;
                for (int c = 0; c <= 4 && c <= j && ok; c++) {
                    for (uint32_t s : S[j - c]) {
                        uint32_t v = s + c * cand;
                        if (v >= stamp.size()) { ok = false; break; }
                        if (stamp[v] == tick) { ok = false; break; }
                        stamp[v] = tick;
                    }
                }
            }
            if (ok) break;
        }
        key.push_back(cand);
        std::vector<std::vector<uint32_t>> T(MAXC + 1);
        for (int j = 0; j <= MAXC; j++)
            for (int c = 0; c <= 4 && c <= j; c++)
                for (uint32_t s : S[j - c]) T[j].push_back(s + c * cand);
        S.swap(T);
        std::printf("rank %2d key %8u   |S7| = %zu\n", r, cand, S[7].size());
        std::fflush(stdout);
        cand++;
    }
    uint32_t mx = *std::max_element(S[7].begin(), S[7].end());
    std::printf("keys = {");
    for (int r = 0; r < NR; r++) std::printf("%u%s", key[r], r + 1 < NR ? ", " : "}\n");
    std::printf("7-card sums: %zu distinct, max %u (%.2f bits)\n", S[7].size(), mx, __builtin_log2((double)mx));
    return 0;
}
