/*
 * npk_holdem.h -- C ABI of the vectorised HoldemTable: N independent poker tables advanced in
 * lock step on the GPU, one thread per table, for self-play loops in which every action asks get_equity.
 *
 * What it replaces in the reference (paths relative to the reference repository root):
 *   npk_holdem_init      HoldemTable.__init__ + add_player x n + reset()             gym_env/env.py:67-168, 526-535
 *   npk_holdem_step      HoldemTable.step for an externally supplied action           gym_env/env.py:170-220
 *                        (_process_decision :308-398, _next_player :611-627, _end_round :537-557, _initiate_round
 *                        :489-524, _start_new_hand :400-438, _check_game_over :445-467, _end_hand :564-571,
 *                        _get_winner :573-590, _award_winner :592-605, _get_legal_moves :629-658) and the whole of
 *                        PlayerCycle                                                  gym_env/cycle.py:10-167
 *   card dealing         _create_card_deck / _distribute_cards / _distribute_cards_to_table   gym_env/env.py:667-688
 *                        (uniform: deck.pop(randint(0, len(deck))))
 *   npk_holdem_queries   the get_equity call of _get_environment                      gym_env/env.py:249-264
 *   npk_holdem_decide    agents/agent_consider_equity.py:21-58 and agents/agent_random.py:19-29
 *
 *   npk_holdem_observe   the observation vector `array_everything` of _get_environment     gym_env/env.py:232-270
 *                        (PlayerData, CommunityData, StageData: env.py:24-63; StageData updates :383-391)
 *
 * Not mirrored (outside the hot path): rendering, the pandas funds history beyond the two rows the reward reads, logging.
 *
 * State is an array of NpkHoldemTable structs in DEVICE memory owned by the caller (a torch tensor of
 * npk_holdem_table_bytes() * N bytes); a host copy of it is a plain C struct array (numpy structured dtype in
 * neuron_poker_b200/holdem.py).  Money is double: the reference computes in Python numbers and half-pot raises produce
 * halves.  All calls are asynchronous on the caller's stream and return 0 or a negative NPK_ERR_* code (npk.h).
 */
#ifndef NPK_HOLDEM_H
#define NPK_HOLDEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPK_MAX_SEATS 10

/* gym_env/enums.py */
enum { NPK_FOLD = 0, NPK_CHECK = 1, NPK_CALL = 2, NPK_RAISE_3BB = 3, NPK_RAISE_HALF_POT = 4, NPK_RAISE_POT = 5,
       NPK_RAISE_2POT = 6, NPK_ALL_IN = 7, NPK_SMALL_BLIND = 8, NPK_BIG_BLIND = 9 };
enum { NPK_PREFLOP = 0, NPK_FLOP = 1, NPK_TURN = 2, NPK_RIVER = 3, NPK_END_HIDDEN = 4, NPK_SHOWDOWN = 5 };
/* agent kinds for npk_holdem_decide */
enum { NPK_AGENT_EQUITY = 0, NPK_AGENT_RANDOM = 1 };

typedef struct NpkHoldemTable {
    /* players (PlayerShell, env.py:753-777) and pots (env.py:408-414) */
    double stack[NPK_MAX_SEATS];
    double player_pots[NPK_MAX_SEATS];
    double player_max_win[NPK_MAX_SEATS];
    double funds_prev[NPK_MAX_SEATS];   /* second-to-last row of funds_history */
    double funds_last[NPK_MAX_SEATS];   /* last row of funds_history (stacks when the current hand started) */
    double community_pot, current_round_pot, min_call, last_player_pot;
    double reward;                      /* reward of the last npk_holdem_step (env.py:280-306; -1 for an illegal move) */
    double small_blind, big_blind, initial_stacks;
    uint64_t rng_counter;               /* cards drawn since init: draw k uses Philox word k of this table's stream */
    double* stage_data;                 /* this table's [4][6][NPK_MAX_SEATS] StageData block, or NULL (npk_holdem_attach_stage_data) */
    uint64_t deck_mask;                 /* remaining deck: bit c set <=> card c still in it.  The reference's deck is an
                                           ordered list that only ever loses elements, so deck.pop(j) is "the j-th set bit" */
    /* PlayerCycle (cycle.py:13-37) */
    int32_t idx, dealer_idx, step_counter, cycle_round_number, max_steps_total /* 0 = None */, last_raiser_step,
        max_steps_after_raiser, max_steps_after_big_blind, last_raiser /* -1 = None */, checkers,
        max_remaining_steps_without_raising;
    /* table */
    int32_t stage, current_player /* seat, -1 = none (False) */, winner_ix /* -1 = None */, dealer_pos, done, funds_rows,
        n_players, max_raises, n_table_cards, n_deck, acting_agent, hands_played;
    int32_t error;                      /* 1: the reference would have raised here (AttributeError / AssertionError) */
    uint32_t legal_moves;               /* bit a set <=> Action a is in env.legal_moves */
    uint8_t can_still[NPK_MAX_SEATS];   /* can_still_make_moves_in_this_hand */
    uint8_t out_of_cash[NPK_MAX_SEATS]; /* out_of_cash_but_contributed */
    uint8_t folder[NPK_MAX_SEATS];
    uint8_t alive[NPK_MAX_SEATS];
    uint8_t first_action[NPK_MAX_SEATS];
    uint8_t autoplay[NPK_MAX_SEATS];    /* seat is an autoplay agent (only the sign of the final reward reads it) */
    uint8_t num_raises[NPK_MAX_SEATS][4];
    uint8_t cards[NPK_MAX_SEATS][2];    /* 0xFF = no card */
    uint8_t table_cards[5];
    uint8_t reserved[1];
} NpkHoldemTable;

int64_t npk_holdem_table_bytes(void);

/* Create N tables of n_players seats each and deal the first hand (HoldemTable(...); add_player x n; reset()).
 * autoplay [n_players] host bytes or NULL (all zero).  Table t draws its cards from the Philox4x32-10 stream keyed by
 * `seed` with counter (draw/4, table_offset + t, 0xD0000000): card = deck.pop(hi32(word * len(deck))). */
int npk_holdem_init(void* tables, int64_t N, int n_players, double initial_stacks, double small_blind, double big_blind,
                    int max_raises_per_player_round, const uint8_t* autoplay, uint64_t seed, int64_t table_offset,
                    void* stream);

/* env.step(action) for every table: actions [N] int8 on the device, a negative action leaves the table untouched.
 * An action that is not legal costs reward -1 and changes nothing else (env.py:222-226); finished (done) tables and
 * tables stuck in SHOWDOWN at hand start (reference defect) are left untouched.  rewards [N] double or NULL.
 * restart_finished != 0: a table whose game ends in this step is reset at once (like npk_holdem_reset_done), after its
 * reward has been recorded. */
int npk_holdem_step(void* tables, int64_t N, const int8_t* actions, double* rewards, uint64_t seed, int64_t table_offset,
                    int restart_finished, void* stream);

/* StageData bookkeeping for the observation vector (optional; the equity agents do not read it).  stage_data is a device
 * array [N][4 streets][6 fields][NPK_MAX_SEATS] of double, fields in the order of StageData (env.py:40-50): calls,
 * raises, min_call_at_action, contribution, stack_at_action, community_pot_at_action.  Attach it right after
 * npk_holdem_init (it is cleared, and the blinds already posted are recorded): npk_holdem_step then keeps it up to
 * date (cleared when a hand starts, written by _process_decision, env.py:383-391).  Pass NULL to detach. */
#define NPK_STAGE_DATA_DOUBLES (4 * 6 * NPK_MAX_SEATS)
int npk_holdem_attach_stage_data(void* tables, int64_t N, double* stage_data, void* stream);

/* Observation length for n_players seats: 22 + 51 * n_players (328 for six players). */
int64_t npk_holdem_observation_size(int n_players);
/* array_everything (env.py:266-270) for every table: obs [N, 22 + 51*n] double =
 *   PlayerData   position, equity_to_river_alive, equity_to_river_2plr (nan), equity_to_river_3plr (nan), stack[n]
 *   CommunityData current_player_position[n] (never set by the reference: zeros), stage one-hot[4], community_pot,
 *                current_round_pot, active_players[n] (zeros), big_blind, small_blind, legal_moves[10]
 *   StageData x 8 calls[n] raises[n] min_call_at_action[n] contribution[n] stack_at_action[n] community_pot_at_action[n]
 *                (the reference only ever writes entries 0..3: its round_number_in_street stays 0, env.py:382)
 * money divided by big_blind * 100 like the reference.  equity [N] double or NULL (then nan).  Needs attached stage data. */
int npk_holdem_observe(const void* tables, int64_t N, const double* equity, double* obs, void* stream);

/* The get_equity arguments of _get_environment for every table: hole [N,2] = cards of the current player (of the
 * winner once the game is over), board [N,5] with 0xFF padding, n_players [N] = sum(player_cycle.alive).
 * Tables that cannot act (done or stuck) get a valid dummy query (players = 1) and active [N] = 0. */
int npk_holdem_queries(const void* tables, int64_t N, uint8_t* hole, uint8_t* board, uint8_t* n_players, uint8_t* active,
                       void* stream);

/* Agent decisions.  wins, ties [N] u64 and `runs` give equity = (wins + ties) / runs (ties count as wins); or pass
 * equity [N] double directly (then wins/ties may be NULL).  Per seat (HOST arrays of n_players entries): agent_kind,
 * min_call_equity, min_bet_equity.  NPK_AGENT_EQUITY is agents/agent_consider_equity.py:21-58; NPK_AGENT_RANDOM picks
 * uniformly among the legal moves in {FOLD, CHECK, CALL, RAISE_POT, RAISE_HALF_POT, RAISE_2POT} (agent_random.py:25-28;
 * Philox word keyed by seed, table, decision_counter).  actions [N] int8: -1 where the table cannot act. */
int npk_holdem_decide(const void* tables, int64_t N, const uint64_t* wins, const uint64_t* ties, int64_t runs,
                      const double* equity, const uint8_t* agent_kind, const double* min_call_equity,
                      const double* min_bet_equity, uint64_t seed, int64_t decision_counter, int64_t table_offset,
                      int8_t* actions, void* stream);

/* Restart every finished table (done != 0) with fresh stacks, like a new env.reset(); others are left alone. */
int npk_holdem_reset_done(void* tables, int64_t N, uint64_t seed, int64_t table_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPK_HOLDEM_H */
