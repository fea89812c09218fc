// ref_shim.cpp -- C-ABI doorway into the UNMODIFIED reference C++ equity calculator.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/README.md).  This file contains no reference code: it includes the
// reference header from where it lies (-I/root/reference/tools/montecarlo_cpp) and is linked against the reference's
// own Montecarlo.cpp compiled in place by oracle/Makefile.  It replaces the 23-line pybind11/cppimport module
// (reference tools/montecarlo_cpp/pymontecarlo.cpp:21-23) with plain `extern "C"` entry points so the built library
// (oracle/_ref/libnpk_ref.so) can travel to the GPU box, where neither /root/reference nor cppimport exist.
//
//   ref_montecarlo        -> montecarlo(set, set, int, int)            Montecarlo.cpp:240-259
//   ref_calc_score        -> calc_score(CardsWithTableCombined)         Montecarlo.cpp:53-237
//   ref_eval_best_hand    -> eval_best_hand(vector<...>)                Montecarlo.cpp:14-35
//   ref_montecarlo_batch  -> the same montecarlo(), fanned out over queries on std::thread workers
#include <tuple>
#include <set>
#include <string>
#include <vector>
#include <sstream>
#include <thread>
#include <atomic>
#include <cstring>
#include <stdexcept>
#include "Montecarlo.h"

static std::set<std::string> split_cards(const char* s)
{
    std::set<std::string> out;
    std::istringstream iss(s ? s : "");
    std::string tok;
    while (iss >> tok) out.insert(tok);
    return out;
}

extern "C" {

// cards are whitespace-separated two-character strings, e.g. "AS KS"; an empty board is "" (or "null" like the
// reference tests pass, both have size < 3 and are cleared by Montecarlo.cpp:242-243)
double ref_montecarlo(const char* my_cards, const char* cards_on_table, int players, int iterations)
{
    try {
        return montecarlo(split_cards(my_cards), split_cards(cards_on_table), players, iterations);
    } catch (const std::exception&) {
        return -1.0;
    }
}

// out_score[8], out_ranks[9] (unused slots untouched), returns 0 or -1 ("Card Type error!")
int ref_calc_score(const char* cards, int* out_score, int* n_score, int* out_ranks, int* n_ranks, char* out_type)
{
    try {
        auto res = calc_score(split_cards(cards));
        const auto& sc = std::get<0>(res);
        const auto& rk = std::get<1>(res);
        *n_score = (int)sc.size();
        *n_ranks = (int)rk.size();
        for (size_t i = 0; i < sc.size() && i < 8; i++) out_score[i] = sc[i];
        for (size_t i = 0; i < rk.size() && i < 9; i++) out_ranks[i] = rk[i];
        std::strncpy(out_type, std::get<2>(res).c_str(), 31);
        out_type[31] = 0;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// hands separated by ';' -- returns 1 if the first hand is (one of) the best, 0 otherwise, -1 on error
int ref_eval_best_hand(const char* hands)
{
    try {
        std::vector<CardsWithTableCombined> all;
        std::string s(hands), part;
        std::istringstream iss(s);
        while (std::getline(iss, part, ';')) all.push_back(split_cards(part.c_str()));
        return eval_best_hand(all) ? 1 : 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// n queries; my_cards[i], boards[i] as above.  Work is handed out query by query to `threads` workers.
void ref_montecarlo_batch(const char* const* my_cards, const char* const* boards, const int* players, int n,
                          int iterations, int threads, double* out)
{
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) return;
            out[i] = ref_montecarlo(my_cards[i], boards[i], players[i], iterations);
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

}  // extern "C"
