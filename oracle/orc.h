/*
 * oracle/orc.h -- CPU ORACLE for the die-e hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithms (alibasaran/die-e,
 * Rust).  Each function cites the reference file:line it follows.  Nothing in the
 * product path (die_e_b200/, csrc/, include/) may include, link or call anything
 * under oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * Parity pin status:
 *   - env step, legal-move generation, sequence order, dedup, win test, is_collectible:
 *     PINNED by the reference's own tests (tests/backgammon_test.rs, transcribed into
 *     tests/golden/ref_backgammon_kats.py by tests/golden/transcribe_reference_tests.py).
 *   - action codec: round trips PINNED by tests/encoding_test.rs.
 *   - TicTacToe: PINNED by tests/tictactoe_test.rs.
 *   - UCB/PUCT numerics, visit counts, rollouts, Dirichlet, sampling, as_tensor, net:
 *     PARITY UNPINNED by the reference (it has no tests for them and cannot be built
 *     here: no Rust toolchain).  For those the oracle, reviewed line by line against the
 *     cited source, is the pin.
 *
 * Randomness: the reference uses rand::thread_rng() everywhere (not seedable).  The
 * oracle replaces every draw site by the injected Philox4x32-10 stream contract that
 * include/diee.h documents (same contract as the CUDA path, separately implemented).
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data layouts (same bytes as include/diee.h, separately declared) ---- */

/* Board = ([i8;24], (u8,u8) hit, (u8,u8) collected)   backgammon_logic.rs:10 */
typedef struct {
    int8_t pts[24];
    uint8_t bar[2]; /* .0 = player -1, .1 = player +1 */
    uint8_t off[2];
} orc_board;

/* Backgammon { board, roll, player, is_second_play }   backgammon_logic.rs:54-60 */
typedef struct {
    orc_board b;
    uint8_t roll[2];
    int8_t player;
    uint8_t second;
} orc_bg_state; /* 32 bytes */

#define ORC_NONE (-2) /* absent sub-move marker */
/* Actions = Vec<(i8,i8)>, length 0..2 on the path   backgammon_logic.rs:12,263 */
typedef struct {
    int8_t from1, to1, from2, to2;
} orc_move;

/* generic sequence for the tree-level KAT functions (any dice-list length <= 4) */
#define ORC_MAX_SEQ 4
typedef struct {
    int8_t from[ORC_MAX_SEQ];
    int8_t to[ORC_MAX_SEQ];
    int32_t len;
} orc_seq;

/* ActionNode { value, children }  backgammon_logic.rs:16-19; siblings are contiguous */
typedef struct {
    int8_t from, to;
    int8_t die;
    int8_t pad;
    int32_t first_child;
    int32_t n_children;
} orc_action_node;

#define ORC_NO_WINNER 2
#define ORC_MAX_MOVES 512

/* ---- Philox4x32-10 stream contract ---- */
enum {
    ORC_STREAM_INIT = 0,    /* first roll of a game                       c0 = 0            */
    ORC_STREAM_GAME = 1,    /* per-ply block of a played game / playout   c0 = ply          */
    ORC_STREAM_ROLLOUT = 2, /* per-ply block of an MCTS rollout           c0 = ply, c3 = epoch<<16|sim */
    ORC_STREAM_EXPAND = 3,  /* dice frozen into a new tree node           c0 = node idx, c3 = epoch */
    ORC_STREAM_DIRICHLET = 4,
    ORC_STREAM_SAMPLE = 5
};
void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
static inline uint8_t orc_die(uint32_t w) { return (uint8_t)(1u + (uint32_t)(((uint64_t)w * 6u) >> 32)); }
static inline uint32_t orc_index(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* ---- backgammon env ---- */
void orc_bg_new(orc_bg_state *s);
void orc_bg_next_state(orc_board *b, const int8_t *from, const int8_t *to, int n, int player);
int orc_bg_is_collectible(const orc_board *b, int player);
int orc_bg_check_win(const orc_board *b, int player);
int orc_bg_check_winner(const orc_bg_state *s);
/* trees: returns number of roots (roots are pool[0..n_roots)), total nodes in *n_pool; <0 on overflow */
int orc_bg_normal_moves(const uint8_t *dice, int nd, const orc_board *b, int player,
                        orc_action_node *pool, int cap, int *n_pool);
int orc_bg_entry_moves(const uint8_t *dice, int nd, const orc_board *b, int player,
                       orc_action_node *pool, int cap, int *n_pool);
int orc_bg_action_trees(const uint8_t *dice, int nd, const orc_board *b, int player,
                        orc_action_node *pool, int cap, int *n_pool);
int orc_bg_extract_sequences_node(const orc_action_node *pool, int node, orc_seq *out, int cap);
int orc_bg_extract_sequences_list(const orc_action_node *pool, int n_roots, orc_seq *out, int cap);
int orc_bg_remove_duplicate_states(const orc_board *b, const orc_seq *in, int n, int player, orc_seq *out);
int orc_bg_valid_moves(const orc_bg_state *s, orc_move *out, int cap);
void orc_bg_apply_move(orc_bg_state *s, orc_move m, uint8_t die0, uint8_t die1);
void orc_bg_skip_turn(orc_bg_state *s, uint8_t die0, uint8_t die1);
uint32_t orc_bg_encode(const orc_bg_state *s, orc_move m);
orc_move orc_bg_decode(const orc_bg_state *s, uint32_t action);
int orc_bg_as_tensor(const orc_bg_state *s, float *out144);

/* one random ply with an injected Philox block (w[0],w[1] dice after, w[2] choice) */
int orc_bg_random_ply(orc_bg_state *s, const uint32_t w[4]);
/* C2: full random playout; returns winner or 0 at the cap; *plies = plies played */
int orc_bg_playout(orc_bg_state *s, uint64_t seed, uint32_t game_id, int round_limit, int32_t *plies);

/* ---- tictactoe env (tictactoe/mod.rs) ---- */
typedef struct {
    int8_t board[9];
    int8_t player;
    uint8_t pad[6];
} orc_ttt_state; /* 16 bytes */
void orc_ttt_new(orc_ttt_state *s);
int orc_ttt_valid_moves(const orc_ttt_state *s, uint8_t *out);
void orc_ttt_apply_move(orc_ttt_state *s, uint8_t m);
void orc_ttt_skip_turn(orc_ttt_state *s);
int orc_ttt_check_winner(const orc_ttt_state *s);
void orc_ttt_as_tensor(const orc_ttt_state *s, float *out27);

/* ---- pure MCTS (mcts/simple_mcts.rs, mcts/node.rs) ---- */
typedef struct {
    uint32_t iterations;
    float c;
    uint32_t simulate_round_limit;
    float dirichlet_alpha;
    float dirichlet_epsilon;
    uint32_t mode_flags;
} orc_mcts_cfg;
#define ORC_MODE_ROLLOUT_CHECK_CURRENT 1u /* Q5: test the rolled-out state, not the start state */
#define ORC_MODE_PASS_CHILD 2u            /* Q6: no-move node gets one EMPTY_MOVE child via skip_turn */

#define ORC_OK 0
#define ORC_ERR_NO_MOVES_PANIC (-3)
#define ORC_ERR_OVERFLOW (-4)

typedef struct {
    int32_t parent;
    float visits, value;
    orc_move action; /* bg; for ttt action.from1 = cell */
    int32_t n_moves; /* legal moves of the node's state */
    int32_t n_untried;
} orc_node_stats;

float orc_ln_f32(float x); /* (float)log((double)x): the contract's ln */
int orc_mcts_search_bg(const orc_bg_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                       uint32_t game_id, uint32_t epoch, orc_move *best, orc_node_stats *nodes_out,
                       orc_bg_state *states_out, int32_t *n_nodes_out);
int orc_mcts_search_bg_ex(const orc_bg_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                          uint32_t game_id, uint32_t epoch, orc_move *best, orc_node_stats *nodes_out,
                          orc_bg_state *states_out, int32_t *n_nodes_out, orc_bg_state *rollout_finals_out);
int orc_mcts_search_ttt(const orc_ttt_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                        uint32_t game_id, uint32_t epoch, uint8_t *best, orc_node_stats *nodes_out,
                        orc_ttt_state *states_out, int32_t *n_nodes_out);

/* ---- AlphaZero search + self-play (mcts/alpha_mcts.rs, alphazero/alpha_parallel.rs) ---- */
typedef void (*orc_eval_fn)(const orc_bg_state *states, int n, float *policy /*[n*1352]*/, float *value /*[n]*/, void *user);
typedef struct {
    int32_t parent, first_child, n_children;
    float visits, value, prior;
    orc_move action;
    orc_bg_state state;
} orc_anode; /* 60 bytes */
typedef struct {
    orc_bg_state state;
    uint32_t game_id;
    uint16_t ply;
    int8_t outcome;
    uint8_t pad;
    uint16_t n_pi;
    uint16_t pad2;
    uint32_t pi_offset;
} orc_traj_record; /* 48 bytes */
void orc_dirichlet(uint64_t seed, uint32_t epoch, float alpha, int n, float *out);
int orc_alpha_mcts_parallel(const orc_bg_state *states, int n, const uint32_t *game_ids, const orc_mcts_cfg *cfg,
                            uint64_t seed, uint32_t epoch, orc_eval_fn eval, void *user, int max_nodes,
                            orc_anode *nodes_out, int32_t *n_nodes_out, int32_t *status_out);
int orc_root_pi(const orc_anode *nodes, float temperature_inv, uint16_t *ids_out, float *pi_out);
int orc_weighted_select(const uint16_t *ids, const float *pi, int n, uint64_t seed, uint32_t game_id, uint32_t ply);
int orc_self_play(int n_games, const orc_mcts_cfg *cfg, float temperature, uint64_t seed, uint32_t first_game_id,
                  orc_eval_fn eval, void *user, int max_nodes, orc_traj_record *rec_out, int rec_cap, uint16_t *pi_ids_out,
                  float *pi_vals_out, int pi_cap, int *n_rec_out, int *n_pi_out, int *n_waves_out);

/* ---- thread-pool batch drivers (the reference's rayon par_iter over games, versus.rs:303-316) ---- */
int orc_mcts_search_bg_batch(const orc_bg_state *states, int n, const int8_t *players, const orc_mcts_cfg *cfg,
                             uint64_t seed, uint32_t first_game_id, uint32_t epoch, orc_move *best, int32_t *status,
                             int nthreads);
int orc_bg_playout_batch(const orc_bg_state *states, int n, uint64_t seed, uint32_t first_game_id, int round_limit,
                         int8_t *winners, int32_t *plies, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
