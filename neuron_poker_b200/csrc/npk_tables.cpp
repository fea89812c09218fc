// npk_tables.cpp -- see npk_tables.h.  Plain C++ (no CUDA), compiled into libnpk.so.
#include "npk_tables.h"

#include <algorithm>
#include <cstring>
#include <map>

namespace npk {

const uint32_t kRankKey[kNumRanks] = {0, 1, 5, 22, 98, 453, 2031, 8698, 22854, 83661, 262349, 636345, 1479181};

static uint64_t pack(int type, const int* ranks, int n)
{
    uint64_t k = (uint64_t)type << 32;
    for (int i = 0; i < n; i++) k |= (uint64_t)(ranks[i] + 2) << (4 * (7 - i));
    return k;
}

// Highest top rank of a 5-run inside a 13-bit rank-presence mask, ace also playing low; -1 if none.
// (hand_evaluator.py:49-58: "-1" is appended when an ace is present, then the first i with s[i]-s[i+4]==4.)
static int straight_top(uint32_t present)
{
    uint32_t m = (present << 1) | ((present >> 12) & 1u);   // bit 0 = ace-as-one, bit r+1 = rank r
    for (int top = 12; top >= 3; top--) {
        uint32_t run = 0x1Fu << (top - 3);                   // bits (top-3)..(top+1) of m = ranks top-4..top
        if ((m & run) == run) return top;
    }
    return -1;
}

uint64_t order_key_from_histogram(const uint8_t h[kNumRanks])
{
    int quad = -1, trips[3], ntr = 0, pairs[4], npr = 0, singles[8], nsg = 0, present_desc[8], npres = 0;
    uint32_t present = 0;
    for (int r = 12; r >= 0; r--) {
        if (!h[r]) continue;
        present |= 1u << r;
        present_desc[npres++] = r;
        if (h[r] == 4) { if (quad < 0) quad = r; }
        else if (h[r] == 3) trips[ntr++] = r;
        else if (h[r] == 2) pairs[npr++] = r;
        else singles[nsg++] = r;
    }
    int rk[8];
    if (quad >= 0) {                                   // :43-46  two highest ranks PRESENT, whichever is the quad
        rk[0] = present_desc[0]; rk[1] = present_desc[1];
        return pack(7, rk, 2);
    }
    if (ntr >= 1 && (ntr >= 2 || npr >= 1)) {          // :36-38  (3,2..) or (3,3..)
        rk[0] = trips[0]; rk[1] = ntr >= 2 ? trips[1] : pairs[0];
        return pack(6, rk, 2);
    }
    if (npr == 3) {                                    // :39-42  three pair -> two pair, kicker = max(third pair, single)
        rk[0] = pairs[0]; rk[1] = pairs[1]; rk[2] = std::max(pairs[2], singles[0]);
        return pack(2, rk, 3);
    }
    int top = straight_top(present);                   // :47-58  at least five distinct ranks from here on
    if (top >= 0) {
        for (int i = 0; i < 5; i++) rk[i] = top - i;   // wheel: 3,2,1,0,-1
        return pack(4, rk, 5);
    }
    if (ntr == 1) { rk[0] = trips[0]; rk[1] = singles[0]; rk[2] = singles[1]; return pack(3, rk, 3); }
    if (npr == 2) { rk[0] = pairs[0]; rk[1] = pairs[1]; rk[2] = singles[0]; return pack(2, rk, 3); }
    if (npr == 1) { rk[0] = pairs[0]; rk[1] = singles[0]; rk[2] = singles[1]; rk[3] = singles[2]; return pack(1, rk, 4); }
    for (int i = 0; i < 5; i++) rk[i] = singles[i];
    return pack(0, rk, 5);
}

uint64_t order_key_from_flush_mask(uint32_t mask)
{
    int rk[8], n = 0;
    for (int r = 12; r >= 0; r--) if (mask >> r & 1u) rk[n++] = r;
    if (straight_top(mask) >= 0) {                     // :68-80, :92-93  ALL flush-suit ranks, "-1" with an ace
        if (mask >> 12 & 1u) rk[n++] = -1;
        return pack(8, rk, n);
    }
    return pack(5, rk, 5);                             // :98-100 top five
}

static void enumerate_histograms(std::vector<std::vector<uint8_t>>& out)
{
    uint8_t h[kNumRanks];
    struct Rec {
        static void go(int r, int left, uint8_t* h, std::vector<std::vector<uint8_t>>& out)
        {
            if (r == kNumRanks - 1) {
                if (left > 4) return;
                h[r] = (uint8_t)left;
                out.emplace_back(h, h + kNumRanks);
                return;
            }
            for (int c = 0; c <= 4 && c <= left; c++) { h[r] = (uint8_t)c; go(r + 1, left - c, h, out); }
        }
    };
    Rec::go(0, 7, h, out);
}

const char* build_tables(Tables& t)
{
    std::vector<std::vector<uint8_t>> hists;
    enumerate_histograms(hists);
    if ((int)hists.size() != kNumHistograms) return "histogram enumeration size";

    std::vector<uint64_t> hkey(hists.size());
    std::vector<uint32_t> hsum(hists.size());
    std::vector<uint64_t> all;
    for (size_t i = 0; i < hists.size(); i++) {
        hkey[i] = order_key_from_histogram(hists[i].data());
        uint32_t s = 0;
        for (int r = 0; r < kNumRanks; r++) s += hists[i][r] * kRankKey[r];
        hsum[i] = s;
        all.push_back(hkey[i]);
    }
    std::vector<uint64_t> fkey(kFlushTableSize, 0);
    for (uint32_t m = 0; m < (uint32_t)kFlushTableSize; m++) {
        int pc = __builtin_popcount(m);
        if (pc < 5 || pc > 7) continue;
        fkey[m] = order_key_from_flush_mask(m);
        all.push_back(fkey[m]);
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    if ((int)all.size() != kNumClasses) return "class count is not 5034";
    t.class_key = all;
    auto id_of = [&](uint64_t k) { return (uint16_t)(std::lower_bound(all.begin(), all.end(), k) - all.begin()); };
    for (int ty = 0; ty <= 9; ty++)
        t.type_start[ty] = (uint16_t)(std::lower_bound(all.begin(), all.end(), (uint64_t)ty << 32) - all.begin());

    t.flush.assign(kFlushTableSize, 0);
    for (uint32_t m = 0; m < (uint32_t)kFlushTableSize; m++) if (fkey[m]) t.flush[m] = id_of(fkey[m]);

    // injectivity of the additive key over the 49,205 histograms
    {
        std::vector<uint32_t> s = hsum;
        std::sort(s.begin(), s.end());
        if (std::adjacent_find(s.begin(), s.end()) != s.end()) return "rank keys are not injective";
        t.max_key = s.back();
    }

    if (t.max_key >> kMixBits) return "plain key sums exceed the mixing modulus";
    for (auto& s : hsum) s = (s * kMixMul) & ((1u << kMixBits) - 1u);   // from here on: mixed keys

    // row displacement: rows = mk >> kRowShift, densest rows first, first offset where every column is free (or
    // already holds the same rank id).
    const uint32_t W = 1u << kRowShift, nrows = 1u << (kMixBits - kRowShift);
    std::vector<std::vector<std::pair<uint16_t, uint16_t>>> rows(nrows);
    for (size_t i = 0; i < hists.size(); i++)
        rows[hsum[i] >> kRowShift].push_back({(uint16_t)(hsum[i] & (W - 1)), id_of(hkey[i])});
    for (auto& row : rows) std::sort(row.begin(), row.end());
    std::vector<uint32_t> order(nrows);
    for (uint32_t r = 0; r < nrows; r++) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return rows[a].size() > rows[b].size(); });
    const uint32_t cap = 65536;
    std::vector<int32_t> tab(cap + W, -1);
    t.row_offset.assign(nrows, 0);
    uint32_t used = 0, first_free = 0;
    for (uint32_t r : order) {
        auto& row = rows[r];
        if (row.empty()) continue;
        while (first_free < cap && tab[first_free] >= 0) first_free++;
        uint32_t start = first_free > row.front().first ? first_free - row.front().first : 0;  // columns ascend
        uint32_t o = start;
        for (;; o++) {
            if (o + W > cap) return "row displacement does not fit 65536 slots";
            bool ok = true;
            for (auto& cv : row) {
                int32_t cur = tab[o + cv.first];
                if (cur >= 0 && cur != (int32_t)cv.second) { ok = false; break; }
            }
            if (ok) break;
        }
        for (auto& cv : row) tab[o + cv.first] = cv.second;
        t.row_offset[r] = (uint16_t)o;
        used = std::max(used, o + row.back().first + 1);
    }
    t.value.assign(used, 0);
    for (uint32_t i = 0; i < used; i++) t.value[i] = tab[i] >= 0 ? (uint16_t)tab[i] : 0;

    // self-check: every histogram and every flush mask reads back its own rank id
    for (size_t i = 0; i < hists.size(); i++) {
        uint32_t k = hsum[i];
        if (t.value[t.row_offset[k >> kRowShift] + (k & (W - 1))] != id_of(hkey[i])) return "row displacement read-back";
    }
    return "";
}

uint16_t host_rank7(const Tables& t, const uint8_t cards[7])
{
    uint32_t total = 0, suit_mask[4] = {0, 0, 0, 0};
    int suit_cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < 7; i++) {
        int r = cards[i] >> 2, s = cards[i] & 3;
        total += card_desc(cards[i]);                  // wraps mod 2^32 exactly like the device adds
        suit_cnt[s]++;
        suit_mask[s] |= 1u << r;
    }
    for (int s = 0; s < 4; s++)
        if (suit_cnt[s] >= 5) return t.flush[suit_mask[s]];
    uint32_t mk = total >> kDescShift;
    return t.value[t.row_offset[mk >> kRowShift] + (mk & ((1u << kRowShift) - 1))];
}

}  // namespace npk

#ifdef NPK_TABLES_MAIN
#include <cstdio>
#include <chrono>
int main()
{
    npk::Tables t;
    auto t0 = std::chrono::steady_clock::now();
    const char* err = npk::build_tables(t);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("err='%s' value=%zu rows=%zu max_key=%u build %.1f ms\n", err, t.value.size(), t.row_offset.size(),
                t.max_key, ms);
    for (int i = 0; i < 10; i++) std::printf("%u ", t.type_start[i]);
    std::printf("\n");
    return 0;
}
#endif
