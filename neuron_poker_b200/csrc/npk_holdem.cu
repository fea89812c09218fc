// npk_holdem.cu -- the vectorised HoldemTable: N independent tables, one thread per table (include/npk_holdem.h).
//
// This is sequential per-table bookkeeping (a few hundred integer / double operations per action); the GPU's job is to
// keep 65,536 of them resident next to the equity kernels so that a self-play step never leaves the device.  Every
// function below names the reference method it follows (gym_env/env.py, gym_env/cycle.py); the reference's recursion
// _next_player -> _end_round -> _initiate_round -> _next_player is unrolled into the loop in next_player().
#include <cuda_runtime.h>

#include "../../include/npk.h"
#include "../../include/npk_holdem.h"
#include "npk_device.cuh"
#include "npk_holdem_launch.h"

namespace npk {

typedef NpkHoldemTable T;

struct HoldemCtx {
    DeviceTables tab;
    uint32_t seed_lo, seed_hi;
    uint32_t table_id;       // global table number (Philox counter word 2)
};

// ---- cards -----------------------------------------------------------------------------------------------------------
// np.random.randint(0, len(deck)) (env.py:680, 686): draw k of a table = hi32(word * len), word = Philox word k of the
// stream (counter = (k / 4, 0, table, 0xD0000000), key = seed).
__device__ __forceinline__ uint32_t deal_word(const HoldemCtx& c, uint64_t k)
{
    uint32_t w[4];
    philox4x32_10((uint32_t)(k >> 2), (uint32_t)(k >> 34), c.table_id, 0xD0000000u, c.seed_lo, c.seed_hi, w);
    return w[k & 3];
}

// k-th (0-based) set bit of a 52-bit mask, by halving on popcounts
__device__ __forceinline__ int nth_set_bit(uint64_t m, int k)
{
    uint32_t w = (uint32_t)m;
    int base = 0;
    const int c = __popc(w);
    if (k >= c) { k -= c; w = (uint32_t)(m >> 32); base = 32; }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t low = (1u << s) - 1u;
        const int cl = __popc(w & low);
        if (k >= cl) { k -= cl; w >>= s; base += s; } else { w &= low; }
    }
    return base;
}

// deck.pop(np.random.randint(0, len(deck))) (env.py:680, 686) on the ordered deck = remove the j-th remaining card
__device__ uint8_t pop_random_card(T& t, const HoldemCtx& c)
{
    const uint32_t n = (uint32_t)t.n_deck;
    const uint32_t j = __umulhi(deal_word(c, t.rng_counter++), n);
    const int card = nth_set_bit(t.deck_mask, (int)j);
    t.deck_mask &= ~(1ull << card);
    t.n_deck = (int32_t)n - 1;
    return (uint8_t)card;
}

// hand value through the lookup tables in global memory (tools/hand_evaluator.py:27-119 ordering)
__device__ uint32_t eval7_global(const DeviceTables& tab, const uint8_t cards[7])
{
    uint32_t total = 0, cnt[4] = {0, 0, 0, 0}, mask[4] = {0, 0, 0, 0};
    for (int i = 0; i < 7; i++) {
        const int c = cards[i];
        total += tab.desc[c];
        cnt[c & 3]++;
        mask[c & 3] |= 1u << (c >> 2);
    }
    const uint32_t mk = total >> kDevDescShift;
    uint32_t v = tab.value[tab.rowoff[mk >> kRowBits] + (mk & kColMask)];
    for (int s = 0; s < 4; s++)
        if (cnt[s] >= 5) v = tab.flush[mask[s]];
    return v;
}

// ---- PlayerCycle (gym_env/cycle.py) ------------------------------------------------------------------------------------
__device__ __forceinline__ void update_alive(T& t)                       // cycle.py:155-158
{
    for (int i = 0; i < t.n_players; i++) t.alive[i] = t.can_still[i] | t.out_of_cash[i];
}

__device__ __forceinline__ int sum_alive(const T& t)
{
    int s = 0;
    for (int i = 0; i < t.n_players; i++) s += t.alive[i];
    return s;
}

__device__ void cycle_new_hand_reset(T& t)                               // cycle.py:39-45
{
    t.idx = 0;                                                           // start_idx
    for (int i = 0; i < t.n_players; i++) { t.can_still[i] = 1; t.out_of_cash[i] = 0; t.folder[i] = 0; }
    t.step_counter = 0;
}

__device__ void cycle_new_street_reset(T& t)                             // cycle.py:47-55
{
    t.step_counter = 0;
    t.cycle_round_number = 0;
    t.idx = t.dealer_idx;
    t.last_raiser_step = t.n_players;
    t.checkers = 0;
    t.max_remaining_steps_without_raising = t.n_players - 1;
    t.last_raiser = -1;
}

__device__ void cycle_init(T& t)                                         // cycle.py:13-37 as built by env.reset (:160-163)
{
    t.max_steps_total = 0;
    t.last_raiser_step = 0;
    t.max_steps_after_raiser = (t.max_raises - 1) * t.n_players - 1;
    t.max_steps_after_big_blind = t.n_players;
    t.last_raiser = -1;
    t.cycle_round_number = 0;
    t.dealer_idx = -1;
    for (int i = 0; i < t.n_players; i++) t.alive[i] = 1;
    cycle_new_hand_reset(t);
    t.checkers = 0;
    t.max_remaining_steps_without_raising = t.n_players;
}

// next_player (cycle.py:57-103): the seat that acts next, or -1 where the reference returns False
__device__ int cycle_next_player(T& t)
{
    const int n = t.n_players;
    int movers = 0;
    for (int i = 0; i < n; i++) movers += t.can_still[i] | t.out_of_cash[i];
    if (movers < 2) return -1;
    t.idx += 1;
    t.step_counter += 1;
    t.idx %= n;
    if (t.step_counter > n) t.cycle_round_number += 1;
    if (t.max_steps_total && t.step_counter > t.max_steps_total) return -1;
    if (t.last_raiser > 0) {                                             // `if self.last_raiser:` -- None and 0 are falsy
        if (t.step_counter > t.last_raiser + t.max_remaining_steps_without_raising) return -1;
        if (t.max_steps_after_raiser && t.step_counter > t.max_steps_after_raiser + t.last_raiser) return -1;
    } else if (t.max_steps_after_raiser && t.step_counter > t.max_steps_after_big_blind + 2) {
        return -1;
    }
    if (t.checkers == sum_alive(t)) return -1;
    for (int guard = 0;; guard++) {
        if (t.can_still[t.idx]) break;
        t.idx += 1;
        t.step_counter += 1;
        t.idx %= n;
        if (t.max_steps_total && t.step_counter >= t.max_steps_total) return -1;
        if (guard > 1024) { t.error = 1; return -1; }                    // the reference would spin forever
    }
    update_alive(t);
    return t.idx;
}

__device__ void cycle_next_dealer(T& t)                                  // cycle.py:105-117
{
    const int n = t.n_players;
    t.dealer_idx += 1;
    t.dealer_idx %= n;
    for (int guard = 0; !t.can_still[t.dealer_idx]; guard++) {
        t.dealer_idx += 1;
        t.dealer_idx %= n;
        if (guard > 64) { t.error = 1; break; }
    }
    t.dealer_pos = t.dealer_idx;
}

__device__ __forceinline__ void deactivate(T& t, int i)                   // cycle.py:123-131
{
    if (!t.can_still[i]) t.error = 1;                                    // assert "Already deactivated"
    t.can_still[i] = 0;
}

// ---- HoldemTable ---------------------------------------------------------------------------------------------------------
// _process_decision (env.py:308-398) for the current player
__device__ void process_decision(T& t, int action)
{
    const int seat = t.current_player;
    if (seat < 0) { t.error = 1; return; }                               // AttributeError on `False.seat`
    if (action == NPK_FOLD) {
        deactivate(t, t.idx);
        t.folder[t.idx] = 1;
    } else {
        const double pot = t.community_pot + t.current_round_pot;
        double contribution = 0.0;
        bool raise_kind = false;
        switch (action) {
            case NPK_CALL: contribution = fmin(t.min_call - t.player_pots[seat], t.stack[seat]); break;
            case NPK_CHECK: contribution = 0.0; t.checkers += 1; break;
            case NPK_RAISE_3BB: contribution = 3 * t.big_blind - t.player_pots[seat]; raise_kind = true; break;
            case NPK_RAISE_HALF_POT: contribution = pot / 2; raise_kind = true; break;
            case NPK_RAISE_POT: contribution = pot; raise_kind = true; break;
            case NPK_RAISE_2POT: contribution = pot * 2; raise_kind = true; break;
            case NPK_ALL_IN: contribution = t.stack[seat]; raise_kind = true; break;
            case NPK_SMALL_BLIND: contribution = fmin(t.small_blind, t.stack[seat]); break;
            case NPK_BIG_BLIND:
                contribution = fmin(t.big_blind, t.stack[seat]);
                t.last_raiser_step = t.step_counter + t.n_players;       // mark_bb, cycle.py:145-148
                t.max_steps_total = t.step_counter + t.n_players * t.max_raises + 2;
                break;
            default: t.error = 1; return;
        }
        if (raise_kind && t.stage >= 0 && t.stage <= 3) t.num_raises[seat][t.stage] += 1;
        const bool blind = action == NPK_SMALL_BLIND || action == NPK_BIG_BLIND;
        if (contribution > t.min_call && !blind) t.last_raiser = t.step_counter;      // mark_raiser
        t.stack[seat] -= contribution;
        t.player_pots[seat] += contribution;
        t.current_round_pot += contribution;
        t.last_player_pot = t.player_pots[seat];
        if (t.stack[seat] == 0 && contribution > 0) {                    // mark_out_of_cash_but_contributed
            t.out_of_cash[t.idx] = 1;
            deactivate(t, t.idx);
        }
        t.min_call = fmax(t.min_call, contribution);
        t.player_max_win[seat] += contribution;
        if (t.stage_data && t.stage >= 0 && t.stage <= 3) {              // StageData, env.py:381-391
            double* sd = t.stage_data + t.stage * 6 * NPK_MAX_SEATS;     // rnd = stage.value + 0 (env.py:382)
            const int pos = t.idx;
            const double unit = t.big_blind * 100;
            sd[0 * NPK_MAX_SEATS + pos] = action == NPK_CALL ? 1.0 : 0.0;
            sd[1 * NPK_MAX_SEATS + pos] = (action == NPK_RAISE_2POT || action == NPK_RAISE_HALF_POT || action == NPK_RAISE_POT) ? 1.0 : 0.0;
            sd[2 * NPK_MAX_SEATS + pos] = t.min_call / unit;
            sd[3 * NPK_MAX_SEATS + pos] += contribution / unit;
            sd[4 * NPK_MAX_SEATS + pos] = t.stack[seat] / unit;
            sd[5 * NPK_MAX_SEATS + pos] = t.community_pot / unit;
        }
    }
    update_alive(t);
}

__device__ void clean_up_pots(T& t)                                      // env.py:559-562
{
    t.community_pot += t.current_round_pot;
    t.current_round_pot = 0;
    for (int i = 0; i < t.n_players; i++) t.player_pots[i] = 0;
}

__device__ void end_round(T& t, const HoldemCtx& c)                      // env.py:537-557 (+ _close_round :660-664)
{
    double s = 0;
    for (int i = 0; i < t.n_players; i++) s += t.player_pots[i];
    t.community_pot += s;
    for (int i = 0; i < t.n_players; i++) t.player_pots[i] = 0;
    int deal = 0;
    if (t.stage == NPK_PREFLOP) { t.stage = NPK_FLOP; deal = 3; }
    else if (t.stage == NPK_FLOP) { t.stage = NPK_TURN; deal = 1; }
    else if (t.stage == NPK_TURN) { t.stage = NPK_RIVER; deal = 1; }
    else if (t.stage == NPK_RIVER) { t.stage = NPK_SHOWDOWN; }
    for (int k = 0; k < deal && t.n_table_cards < 5; k++) t.table_cards[t.n_table_cards++] = pop_random_card(t, c);
    clean_up_pots(t);
}

// the street-independent head of _initiate_round (env.py:489-498)
__device__ void reset_round_state(T& t)
{
    t.min_call = 0;
    cycle_new_street_reset(t);
    if (t.stage != NPK_PREFLOP && t.n_players == 2) t.idx += 1;          // heads-up: advance by one after the flop
}

// _next_player (env.py:611-627) including the _end_round / _initiate_round chain it may start
__device__ void next_player(T& t, const HoldemCtx& c)
{
    for (int guard = 0; guard < 8; guard++) {
        const int cp = cycle_next_player(t);
        t.current_player = cp;
        if (cp >= 0) return;
        if (sum_alive(t) < 2) { t.stage = NPK_END_HIDDEN; return; }
        end_round(t, c);
        reset_round_state(t);                                            // _initiate_round for the new street
        if (t.stage == NPK_SHOWDOWN) return;
        if (t.stage > NPK_RIVER) { t.error = 1; return; }                // RuntimeError() in the reference
        t.max_steps_total = t.n_players * t.max_raises;
    }
    t.error = 1;
}

__device__ uint32_t legal_moves(const T& t)                              // env.py:629-658
{
    if (t.stage == NPK_SHOWDOWN) return 0;
    const int seat = t.current_player;
    if (seat < 0) return 0;
    uint32_t m = 0;
    double mx = t.player_pots[0];
    for (int i = 1; i < t.n_players; i++) mx = fmax(mx, t.player_pots[i]);
    if (t.player_pots[seat] == mx) m |= 1u << NPK_CHECK;
    else m |= (1u << NPK_CALL) | (1u << NPK_FOLD);
    if (t.stage < 0 || t.stage > 3) return m;                            // KeyError in the reference; never reached
    if (t.num_raises[seat][t.stage] < t.max_raises) {
        const double pot = t.community_pot + t.current_round_pot, st = t.stack[seat];
        if (st >= 3 * t.big_blind - t.player_pots[seat]) m |= 1u << NPK_RAISE_3BB;
        if (st >= pot / 2 && pot / 2 >= t.min_call) m |= 1u << NPK_RAISE_HALF_POT;
        if (st >= pot && pot >= t.min_call) m |= 1u << NPK_RAISE_POT;
        if (st >= pot * 2 && pot * 2 >= t.min_call) m |= 1u << NPK_RAISE_2POT;
        if (st > 0) m |= 1u << NPK_ALL_IN;
    }
    return m;
}

__device__ int get_winner(T& t, const HoldemCtx& c)                      // env.py:573-590, hand_evaluator.py:9-17
{
    int cnt = 0, only = -1;
    for (int i = 0; i < t.n_players; i++)
        if ((t.can_still[i] | t.out_of_cash[i]) && !t.folder[i]) { cnt++; if (only < 0) only = i; }
    if (cnt == 1) return only;
    if (cnt == 0 || t.stage != NPK_SHOWDOWN || t.n_table_cards != 5) { t.error = 1; return only < 0 ? 0 : only; }
    int best = -1;
    uint32_t bv = 0;
    for (int i = 0; i < t.n_players; i++) {
        if (!((t.can_still[i] | t.out_of_cash[i]) && !t.folder[i])) continue;
        const uint8_t h[7] = {t.cards[i][0], t.cards[i][1], t.table_cards[0], t.table_cards[1], t.table_cards[2],
                              t.table_cards[3], t.table_cards[4]};
        const uint32_t v = eval7_global(c.tab, h);
        if (best < 0 || v > bv) { best = i; bv = v; }                    // first index among the best (stable sort)
    }
    return best;
}

__device__ void award_winner(T& t, int w)                                // env.py:592-605
{
    const double m = t.player_max_win[w];
    double total = 0, all = 0;
    for (int i = 0; i < t.n_players; i++) { total += fmin(m, t.player_max_win[i]); all += t.player_max_win[i]; }
    t.stack[w] += total;
    t.winner_ix = w;
    if (total < all)
        for (int i = 0; i < t.n_players; i++) t.stack[i] += fmax(0.0, t.player_max_win[i] - m);
}

__device__ void end_hand(T& t, const HoldemCtx& c)                       // env.py:564-571
{
    clean_up_pots(t);
    t.winner_ix = get_winner(t, c);
    award_winner(t, t.winner_ix);
    t.hands_played += 1;
}

__device__ bool check_game_over(T& t)                                    // env.py:445-467
{
    cycle_new_hand_reset(t);
    int remaining = 0;
    for (int i = 0; i < t.n_players; i++) {
        if (t.stack[i] > 0) remaining++;
        else deactivate(t, i);
    }
    if (remaining < 2) { t.done = 1; return true; }
    if (t.stack[0] == 0) { t.done = 1; return true; }                    // "Early termination: Agent lost all its money"
    return false;
}

__device__ void start_new_hand(T& t, const HoldemCtx& c)                 // env.py:400-438
{
    for (int i = 0; i < t.n_players; i++) { t.funds_prev[i] = t.funds_last[i]; t.funds_last[i] = t.stack[i]; }
    t.funds_rows += 1;                                                   // _save_funds_history
    for (int i = 0; i < t.n_players; i++)
        for (int s = 0; s < 4; s++) t.num_raises[i][s] = 0;
    if (check_game_over(t)) return;
    if (t.stage_data)                                                    // fresh StageData objects (env.py:405)
        for (int k = 0; k < NPK_STAGE_DATA_DOUBLES; k++) t.stage_data[k] = 0.0;
    t.n_table_cards = 0;
    for (int i = 0; i < 5; i++) t.table_cards[i] = 0xFF;
    t.deck_mask = (1ull << 52) - 1ull;                                   // _create_card_deck: id = 4*rank + suit
    t.n_deck = 52;
    t.stage = NPK_PREFLOP;
    t.community_pot = 0;
    t.current_round_pot = 0;
    t.last_player_pot = 0;
    for (int i = 0; i < t.n_players; i++) {
        t.player_pots[i] = 0; t.player_max_win[i] = 0; t.first_action[i] = 1;
        t.cards[i][0] = 0xFF; t.cards[i][1] = 0xFF;
    }
    cycle_next_dealer(t);
    for (int i = 0; i < t.n_players; i++) {                              // _distribute_cards
        if (t.stack[i] <= 0) continue;
        t.cards[i][0] = pop_random_card(t, c);
        t.cards[i][1] = pop_random_card(t, c);
    }
    // _initiate_round, PREFLOP branch (env.py:504-513)
    reset_round_state(t);
    t.max_steps_total = t.n_players * t.max_raises + 2;
    next_player(t, c);
    process_decision(t, NPK_SMALL_BLIND);
    next_player(t, c);
    process_decision(t, NPK_BIG_BLIND);
    next_player(t, c);
}

// what _get_environment leaves behind (env.py:232-278): the current player falls back to the winner when nobody is to
// act, and the legal moves are those of that player
__device__ void observe(T& t)
{
    if (t.current_player < 0) {
        if (t.winner_ix >= 0) t.current_player = t.winner_ix;
        else t.error = 1;                                                // players[None] -> TypeError
    }
    t.legal_moves = legal_moves(t);
}

__device__ void table_reset(T& t, const HoldemCtx& c)                    // env.py:138-168
{
    t.done = 0;
    t.reward = 0;
    t.funds_rows = 0;
    for (int i = 0; i < t.n_players; i++) {
        t.stack[i] = t.initial_stacks; t.first_action[i] = 1; t.funds_prev[i] = 0; t.funds_last[i] = 0;
    }
    t.dealer_pos = 0;
    cycle_init(t);
    start_new_hand(t, c);
    observe(t);
}

__global__ void holdem_init_kernel(T* tables, long long n, HoldemInit cfg, HoldemCtx c0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T t;
    memset(&t, 0, sizeof(T));
    t.n_players = cfg.n_players;
    t.max_raises = cfg.max_raises;
    t.small_blind = cfg.small_blind; t.big_blind = cfg.big_blind; t.initial_stacks = cfg.initial_stacks;
    for (int s = 0; s < NPK_MAX_SEATS; s++) {
        t.autoplay[s] = cfg.autoplay[s];
        t.cards[s][0] = 0xFF; t.cards[s][1] = 0xFF;
    }
    for (int k = 0; k < 5; k++) t.table_cards[k] = 0xFF;
    t.winner_ix = -1;
    t.current_player = -1;
    t.last_raiser = -1;
    t.current_round_pot = 9;                                             // env.py:117 (overwritten by the first hand)
    HoldemCtx c = c0;
    c.table_id = c0.table_id + (uint32_t)i;
    table_reset(t, c);
    tables[i] = t;
}

__global__ void holdem_reset_done_kernel(T* tables, long long n, HoldemCtx c0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !tables[i].done) return;
    T& t = tables[i];                       // in place: only the fields a reset touches travel
    HoldemCtx c = c0;
    c.table_id = c0.table_id + (uint32_t)i;
    t.error = 0;
    table_reset(t, c);
}

// HoldemTable.step for a player that is not an autoplay agent (env.py:170-200)
// The tables of a block are staged in shared memory (coalesced 8-byte copies in and out): the state machine is a long
// chain of dependent small accesses in 32 different control-flow paths per warp, which is slow against L2 and cheap
// against shared memory; HBM sees each table exactly once in each direction.
constexpr int kStepThreads = 64;                                         // 64 x 760 B = 48,640 B of shared memory
static_assert(sizeof(T) % 8 == 0, "NpkHoldemTable is copied in 8-byte units");
static_assert(kStepThreads * sizeof(T) <= 48 * 1024, "the staged tables must fit the default shared-memory limit");

__device__ void step_one_table(T& t, int action, const HoldemCtx& c, int restart_finished, double* reward_out)
{
    if (t.done || t.error) { if (reward_out) *reward_out = 0; return; }
    t.reward = 0;
    t.acting_agent = t.idx;
    const uint32_t legal = legal_moves(t);
    if (action > NPK_ALL_IN || !(legal >> action & 1u)) {
        t.reward = -1;                                                   // _illegal_move (env.py:222-226)
    } else {
        process_decision(t, action);                                     // _execute_step (env.py:210-220)
        next_player(t, c);
        if (t.stage == NPK_END_HIDDEN || t.stage == NPK_SHOWDOWN) {
            end_hand(t, c);
            start_new_hand(t, c);
        }
        observe(t);
        const int a = t.acting_agent;
        if (t.first_action[a] || t.done) {                               // env.py:195-197, _calculate_reward :280-306
            t.first_action[a] = 0;
            if (t.done) {
                const double won = (t.winner_ix >= 0 && t.autoplay[t.winner_ix]) ? -1.0 : 1.0;
                t.reward = t.initial_stacks * t.n_players * won;
            } else if (t.funds_rows > 1) {
                t.reward = t.funds_last[a] - t.funds_prev[a];
            }
        }
    }
    if (reward_out) *reward_out = t.reward;
    if (restart_finished && t.done) { t.error = 0; table_reset(t, c); }   // a new env.reset() for a finished game
}

// HoldemTable.step for a player that is not an autoplay agent (env.py:170-200)
__global__ void __launch_bounds__(kStepThreads) holdem_step_kernel(T* tables, long long n, const int8_t* __restrict__ actions,
                                                                   double* __restrict__ rewards, HoldemCtx c0,
                                                                   int restart_finished)
{
    extern __shared__ __align__(16) unsigned long long s_words[];
    T* s_tab = reinterpret_cast<T*>(s_words);
    const long long first = (long long)blockIdx.x * kStepThreads;
    const long long count = min((long long)kStepThreads, n - first);
    if (count <= 0) return;
    constexpr int kWordsPerTable = (int)(sizeof(T) / 8);
    const unsigned long long* g_words = reinterpret_cast<const unsigned long long*>(tables + first);
    const int total_words = (int)count * kWordsPerTable;
    // the block's tables travel as ONE bulk copy each way (cp.async.bulk, the TMA 1-D path): 45 KB in a single instruction
    // instead of 89 dependent 8-byte loads per thread -- the first version of this kernel spent its time waiting for them
    // (98 us per step for 65,536 tables, long-scoreboard stalls 46 %, profiles/r02_ncu_holdem_step.txt)
    __shared__ uint64_t s_bar;
    const uint32_t bytes = (uint32_t)total_words * 8u;
    const bool bulk = (bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(g_words) & 15u) == 0;
    if (bulk) {
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, bytes);
            bulk_g2s(s_words, g_words, bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        for (int w = threadIdx.x; w < total_words; w += kStepThreads) s_words[w] = g_words[w];
        __syncthreads();
    }
    const long long i = first + threadIdx.x;
    bool touched = false;
    if (threadIdx.x < count) {
        const int action = actions[i];
        if (action >= 0) {
            HoldemCtx c = c0;
            c.table_id = c0.table_id + (uint32_t)i;
            step_one_table(s_tab[threadIdx.x], action, c, restart_finished, rewards ? rewards + i : nullptr);
            touched = true;
        }
    }
    // write back only if some table of the block was stepped
    if (!__syncthreads_or(touched)) return;
    unsigned long long* o_words = reinterpret_cast<unsigned long long*>(tables + first);
    if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the threads' shared-memory writes -> the copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o_words), "r"(smem_u32(s_words)), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the read
        }
    } else {
        for (int w = threadIdx.x; w < total_words; w += kStepThreads) o_words[w] = s_words[w];
    }
}

// Attach the StageData array.  The first hand has already been dealt by init, so the blinds it posted are replayed into
// the fresh block: small blind = seat after the dealer, big blind the one after (only the players who were dealt in).
__global__ void holdem_attach_kernel(T* tables, long long n, double* stage_data)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T& t = tables[i];
    t.stage_data = stage_data ? stage_data + i * NPK_STAGE_DATA_DOUBLES : nullptr;
    if (!stage_data) return;
    double* sd = t.stage_data;
    for (int k = 0; k < NPK_STAGE_DATA_DOUBLES; k++) sd[k] = 0.0;
    if (t.done || t.stage != NPK_PREFLOP || t.hands_played != 0 || t.funds_rows != 1) return;
    // Freshly reset table: exactly two _process_decision calls (the blinds) have happened.  The reference recorded,
    // for each of them, the values right after the contribution (env.py:383-391).
    const double unit = t.big_blind * 100;
    int seat = t.dealer_idx;
    double running_min_call = 0;
    for (int b = 0; b < 2; b++) {
        do { seat = (seat + 1) % t.n_players; } while (t.player_pots[seat] == 0 && t.stack[seat] == t.initial_stacks);
        const double c = t.player_pots[seat];
        running_min_call = fmax(running_min_call, c);
        sd[2 * NPK_MAX_SEATS + seat] = running_min_call / unit;
        sd[3 * NPK_MAX_SEATS + seat] = c / unit;
        sd[4 * NPK_MAX_SEATS + seat] = t.stack[seat] / unit;
        sd[5 * NPK_MAX_SEATS + seat] = 0.0;
    }
}

// array_everything (env.py:232-270)
__global__ void holdem_observe_kernel(const T* __restrict__ tables, long long n, const double* __restrict__ equity,
                                      double* __restrict__ obs)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T& t = tables[i];
    const int np = t.n_players;
    const int width = 22 + 51 * np;
    double* o = obs + i * width;
    const double unit = t.big_blind * 100;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    int k = 0;
    o[k++] = (double)t.current_player;                                   // PlayerData.position = current_player.seat
    o[k++] = equity ? equity[i] : nan;
    o[k++] = nan; o[k++] = nan;                                          // calculate_equity=False (env.py:256-259)
    for (int s = 0; s < np; s++) o[k++] = t.stack[s] / unit;
    for (int s = 0; s < np; s++) o[k++] = 0.0;                           // current_player_position: never set
    const int st = t.stage < 3 ? t.stage : 3;
    for (int s = 0; s < 4; s++) o[k++] = s == st ? 1.0 : 0.0;
    o[k++] = t.community_pot / unit;
    o[k++] = t.current_round_pot / unit;
    for (int s = 0; s < np; s++) o[k++] = 0.0;                           // active_players: never set
    o[k++] = t.big_blind;
    o[k++] = t.small_blind;
    for (int a = 0; a < 10; a++) o[k++] = (t.legal_moves >> a & 1u) ? 1.0 : 0.0;
    for (int r = 0; r < 8; r++)
        for (int f = 0; f < 6; f++)
            for (int s = 0; s < np; s++)
                o[k++] = (r < 4 && t.stage_data) ? t.stage_data[(r * 6 + f) * NPK_MAX_SEATS + s] : 0.0;
}

__global__ void holdem_queries_kernel(const T* __restrict__ tables, long long n, uint8_t* __restrict__ hole,
                                      uint8_t* __restrict__ board, uint8_t* __restrict__ n_players, uint8_t* __restrict__ active)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T& t = tables[i];
    const int seat = t.current_player;
    bool ok = !t.done && !t.error && seat >= 0 && t.stage <= NPK_RIVER && t.legal_moves != 0 &&
              t.cards[seat][0] < 52 && t.cards[seat][1] < 52;
    int np = 0;
    for (int s = 0; s < t.n_players; s++) np += t.alive[s];
    if (np < 1 || np > 10) ok = false;
    if (ok) {
        hole[2 * i] = t.cards[seat][0]; hole[2 * i + 1] = t.cards[seat][1];
        for (int k = 0; k < 5; k++) board[5 * i + k] = k < t.n_table_cards ? t.table_cards[k] : 0xFF;
        n_players[i] = (uint8_t)np;
    } else {                                                             // harmless one-player query
        hole[2 * i] = 0; hole[2 * i + 1] = 1;
        for (int k = 0; k < 5; k++) board[5 * i + k] = 0xFF;
        n_players[i] = 1;
    }
    if (active) active[i] = ok ? 1 : 0;
}

__global__ void holdem_decide_kernel(const T* __restrict__ tables, long long n, const unsigned long long* __restrict__ wins,
                                     const unsigned long long* __restrict__ ties, long long runs,
                                     const double* __restrict__ equity, HoldemAgents ag, HoldemCtx c0,
                                     unsigned long long decision_counter, int8_t* __restrict__ actions)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T& t = tables[i];
    const int seat = t.current_player;
    const uint32_t legal = t.legal_moves;
    if (t.done || t.error || seat < 0 || legal == 0) { actions[i] = -1; return; }
    int action;
    if (ag.kind[seat] == NPK_AGENT_RANDOM) {                             // agents/agent_random.py:25-28
        const uint32_t allowed = legal & ((1u << NPK_FOLD) | (1u << NPK_CHECK) | (1u << NPK_CALL) | (1u << NPK_RAISE_POT) |
                                          (1u << NPK_RAISE_HALF_POT) | (1u << NPK_RAISE_2POT));
        const int cnt = __popc(allowed);
        uint32_t w[4];
        philox4x32_10((uint32_t)decision_counter, (uint32_t)(decision_counter >> 32), c0.table_id + (uint32_t)i, 0xA0000000u,
                      c0.seed_lo, c0.seed_hi, w);
        int pick = (int)__umulhi(w[0], (uint32_t)cnt);
        action = NPK_FOLD;
        for (int a = 0; a < 8; a++)
            if (allowed >> a & 1u) { if (pick == 0) { action = a; break; } pick--; }
    } else {                                                             // agents/agent_consider_equity.py:21-58
        const double eq = equity ? equity[i] : (double)(wins[i] + ties[i]) / (double)runs;
        const double bet = ag.min_bet_equity[seat], call = ag.min_call_equity[seat];
        const double inc1 = 0.1, inc2 = 0.2;
        if (eq > bet + inc2 && (legal >> NPK_ALL_IN & 1u)) action = NPK_ALL_IN;
        else if (eq > bet + inc1 && (legal >> NPK_RAISE_2POT & 1u)) action = NPK_RAISE_2POT;
        else if (eq > bet && (legal >> NPK_RAISE_POT & 1u)) action = NPK_RAISE_POT;
        else if (eq > bet - inc1 && (legal >> NPK_RAISE_HALF_POT & 1u)) action = NPK_RAISE_HALF_POT;
        else if (eq > call && (legal >> NPK_CALL & 1u)) action = NPK_CALL;
        else if (legal >> NPK_CHECK & 1u) action = NPK_CHECK;
        else action = NPK_FOLD;
    }
    actions[i] = (int8_t)action;
}

// ---- launchers ---------------------------------------------------------------------------------------------------------------
static HoldemCtx make_ctx(const DeviceTables& tab, uint64_t seed, long long table_offset)
{
    HoldemCtx c;
    c.tab = tab;
    c.seed_lo = (uint32_t)seed; c.seed_hi = (uint32_t)(seed >> 32);
    c.table_id = (uint32_t)table_offset;
    return c;
}

constexpr int kHoldemThreads = 128;
static int blocks_for(long long n) { return (int)((n + kHoldemThreads - 1) / kHoldemThreads); }

cudaError_t launch_holdem_init(const DeviceTables& tab, void* tables, long long n, const HoldemInit& cfg, uint64_t seed,
                               long long table_offset, cudaStream_t s)
{
    holdem_init_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(static_cast<T*>(tables), n, cfg, make_ctx(tab, seed, table_offset));
    return cudaGetLastError();
}

cudaError_t launch_holdem_reset_done(const DeviceTables& tab, void* tables, long long n, uint64_t seed, long long table_offset,
                                     cudaStream_t s)
{
    holdem_reset_done_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(static_cast<T*>(tables), n, make_ctx(tab, seed, table_offset));
    return cudaGetLastError();
}

cudaError_t launch_holdem_step(const DeviceTables& tab, void* tables, long long n, const int8_t* actions, double* rewards,
                               uint64_t seed, long long table_offset, int restart_finished, cudaStream_t s)
{
    const int blocks = (int)((n + kStepThreads - 1) / kStepThreads);
    holdem_step_kernel<<<blocks, kStepThreads, kStepThreads * sizeof(T), s>>>(static_cast<T*>(tables), n, actions, rewards,
                                                                             make_ctx(tab, seed, table_offset), restart_finished);
    return cudaGetLastError();
}

cudaError_t launch_holdem_attach(void* tables, long long n, double* stage_data, cudaStream_t s)
{
    holdem_attach_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(static_cast<T*>(tables), n, stage_data);
    return cudaGetLastError();
}

cudaError_t launch_holdem_observe(const void* tables, long long n, const double* equity, double* obs, cudaStream_t s)
{
    holdem_observe_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(static_cast<const T*>(tables), n, equity, obs);
    return cudaGetLastError();
}

cudaError_t launch_holdem_queries(const void* tables, long long n, uint8_t* hole, uint8_t* board, uint8_t* n_players,
                                  uint8_t* active, cudaStream_t s)
{
    holdem_queries_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(static_cast<const T*>(tables), n, hole, board, n_players, active);
    return cudaGetLastError();
}

cudaError_t launch_holdem_decide(const DeviceTables& tab, const void* tables, long long n, const uint64_t* wins,
                                 const uint64_t* ties, long long runs, const double* equity, const HoldemAgents& ag,
                                 uint64_t seed, unsigned long long decision_counter, long long table_offset, int8_t* actions,
                                 cudaStream_t s)
{
    holdem_decide_kernel<<<blocks_for(n), kHoldemThreads, 0, s>>>(
        static_cast<const T*>(tables), n, reinterpret_cast<const unsigned long long*>(wins),
        reinterpret_cast<const unsigned long long*>(ties), runs, equity, ag, make_ctx(tab, seed, table_offset), decision_counter,
        actions);
    return cudaGetLastError();
}

}  // namespace npk
