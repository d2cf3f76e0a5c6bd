/*
 * diee.h -- C ABI of libdiee_cuda.so: the B200-native (sm_100a) engine for the die-e hot path
 * (batched MCTS over thousands of concurrent backgammon games).
 *
 * The reference (alibasaran/die-e) is safe Rust with NO FFI/plugin interface.
 * The seam this library sits behind is therefore the Rust API itself; every entry point below
 * names the reference item it replaces (file:line under the reference's src/).  A Rust `-sys`
 * crate binding this header, and the safe wrapper re-exposing `LearnableGame`, `mct_search`,
 * `alpha_mcts_parallel`, `self_play_parallel`, is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int32 status: DIEE_OK or a negative DIEE_ERR_*; diee_last_error()
 *     gives the message.  Nothing unwinds across the boundary (the reference panics instead).
 *   - plain pointers and sizes only.  Functions without a suffix take HOST buffers and do the
 *     host<->device copies themselves (synchronous on return).  Functions ending in `_dev`
 *     take DEVICE pointers, are stream-ordered on the context's stream and return without
 *     synchronising: results are valid after diee_sync().
 *   - one diee_ctx per GPU; calls on one ctx must be serialised by the caller.
 *   - there is NO CPU fallback: every compute entry point fails with DIEE_ERR_CUDA when no
 *     CUDA device is usable.
 *
 * Injected random stream (part of the ABI; the reference uses rand::thread_rng(), which cannot
 * be seeded -- SURVEY.md Appendix C).  Every draw is a word of one Philox4x32-10 block with
 *     key     = (seed_lo, seed_hi)
 *     counter = (c0 = draw index, c1 = global game id, c2 = stream kind, c3 = aux)
 *   die face from word w:      1 + ((uint64)w * 6 >> 32)
 *   uniform index in [0,n):    (uint64)w * n >> 32
 *   block words: w0,w1 = the two dice (roll.0, roll.1)   w2 = uniform move choice   w3 = spare
 *   streams:
 *     DIEE_STREAM_INIT     c0 = 0           first roll of game c1          (alpha_parallel.rs:105-108)
 *     DIEE_STREAM_GAME     c0 = ply         ply of a played game/playout: w2 picks the move of
 *                                           ply c0, w0,w1 are the dice rolled after it
 *                                           (versus.rs:307-316, backgammon_logic.rs:176-196)
 *     DIEE_STREAM_ROLLOUT  c0 = ply         same, inside Node::simulate (node.rs:176-196);
 *                                           c3 = epoch<<16 | simulation index
 *     DIEE_STREAM_EXPAND   c0 = node index  dice frozen into a new tree node (node.rs:124-125,
 *                                           158-172); c3 = epoch
 *     DIEE_STREAM_DIRICHLET, DIEE_STREAM_SAMPLE  see diee_alpha_* below
 *   `epoch` distinguishes successive searches of one game (e.g. its move number).
 */
#ifndef DIEE_H
#define DIEE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIEE_OK 0
#define DIEE_ERR_INVALID (-1)        /* bad argument */
#define DIEE_ERR_CUDA (-2)           /* CUDA runtime error / no device */
#define DIEE_ERR_NO_MOVES_PANIC (-3) /* the reference would panic here (node.rs:119-121, quirk Q6) */
#define DIEE_ERR_OVERFLOW (-4)       /* move buffer / node pool exhausted */
#define DIEE_ERR_NOT_ROLLED (-5)     /* roll == (0,0): reference asserts (backgammon_logic.rs:199,404) */

#define DIEE_MAX_MOVES 256 /* legal-move capacity per state (observed max 123 unique) */
#define DIEE_NONE (-2)     /* absent sub-move */
#define DIEE_ACTION_SPACE 1352

enum { DIEE_GAME_BACKGAMMON = 0, DIEE_GAME_TICTACTOE = 1 };
enum {
    DIEE_STREAM_INIT = 0,
    DIEE_STREAM_GAME = 1,
    DIEE_STREAM_ROLLOUT = 2,
    DIEE_STREAM_EXPAND = 3,
    DIEE_STREAM_DIRICHLET = 4,
    DIEE_STREAM_SAMPLE = 5
};

/* `Backgammon` state, packed to 32 bytes (backgammon_logic.rs:10,54-60).  pts: signed counts,
 * player -1 negative (moves toward 0), player +1 positive (moves toward 23); bar/off index 0
 * belongs to player -1, index 1 to player +1.  `id` is not part of the packed state. */
typedef struct {
    int8_t pts[24];
    uint8_t bar[2];
    uint8_t off[2];
    uint8_t roll[2];
    int8_t player;
    uint8_t second; /* is_second_play */
} diee_bg_state;

/* `Actions = Vec<(i8,i8)>` of length 0..2 (backgammon_logic.rs:12,263): to == -1 collects,
 * from == -1 enters from the bar; an absent sub-move is (DIEE_NONE, DIEE_NONE);
 * EMPTY_MOVE (backgammon_logic.rs:72) is all DIEE_NONE. */
typedef struct {
    int8_t from1, to1, from2, to2;
} diee_move;

/* `TicTacToe` (tictactoe/mod.rs:6-13), 16 bytes */
typedef struct {
    int8_t board[9];
    int8_t player;
    uint8_t pad[6];
} diee_ttt_state;

/* `MctsConfig` (lib.rs:33-40) + mode flags */
typedef struct {
    uint32_t iterations;
    float c;
    uint32_t simulate_round_limit;
    float dirichlet_alpha;
    float dirichlet_epsilon;
    uint32_t mode_flags;
} diee_mcts_cfg;
/* default (0) = reference-exact.  Quirks of the reference exposed as modes: */
#define DIEE_MODE_ROLLOUT_CHECK_CURRENT 1u /* rollout tests the rolled-out state (node.rs:181 tests the start state) */
#define DIEE_MODE_PASS_CHILD 2u            /* a no-move node gets one EMPTY_MOVE child instead of the panic */

/* one node of the SoA pool as read back for inspection (mcts/node.rs:9-19) */
typedef struct {
    int32_t parent; /* -1 = root */
    float visits, value;
    diee_move action; /* tictactoe: from1 = cell */
    int32_t n_moves;  /* legal moves of the node's state (expandable_moves at creation) */
    int32_t n_untried;
} diee_node;

/* per-game work counters of one search (feed the roofline arithmetic in bench.py) */
typedef struct {
    uint64_t rollout_plies;   /* plies executed inside Node::simulate */
    uint32_t select_levels;   /* select_ucb calls (tree levels descended) */
    uint32_t select_children; /* children scored over all select_ucb calls */
    uint32_t expansions;      /* nodes created */
    uint32_t terminal_leaves; /* iterations that ended on a terminal leaf */
} diee_search_stats;

typedef struct diee_ctx diee_ctx;

/* ---- context ---- */
int32_t diee_ctx_create(int32_t device, diee_ctx **out);
int32_t diee_ctx_destroy(diee_ctx *ctx);
/* use an existing CUDA stream (cudaStream_t) for all launches of this ctx; NULL = ctx's own */
int32_t diee_ctx_set_stream(diee_ctx *ctx, void *cuda_stream);
int32_t diee_sync(diee_ctx *ctx);
const char *diee_last_error(const diee_ctx *ctx);
const char *diee_version(void);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t diee_launch_count(const diee_ctx *ctx);
/* raw device memory owned by the library (so hosts without a CUDA allocator can use *_dev) */
int32_t diee_dev_alloc(diee_ctx *ctx, uint64_t bytes, void **dptr_out);
int32_t diee_dev_free(diee_ctx *ctx, void *dptr);
int32_t diee_dev_upload(diee_ctx *ctx, void *dptr, const void *host, uint64_t bytes);
int32_t diee_dev_download(diee_ctx *ctx, void *host, const void *dptr, uint64_t bytes);

/* ---- the injected stream, host side (same function the kernels evaluate) ---- */
void diee_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);

/* ---- backgammon env ----
 * get_valid_moves (backgammon_logic.rs:403-414): ordered, de-duplicated legal plays of each
 * state.  moves_out[i*DIEE_MAX_MOVES + k], counts_out[i]; ids_out (nullable) = encode() of each
 * play (backgammon_logic.rs:262-359).  counts_out[i] = DIEE_ERR_NOT_ROLLED for an unrolled state. */
int32_t diee_bg_valid_moves(diee_ctx *ctx, const diee_bg_state *states, int32_t n, diee_move *moves_out,
                            int32_t *counts_out, uint16_t *ids_out);
int32_t diee_bg_valid_moves_dev(diee_ctx *ctx, const diee_bg_state *states, int32_t n, diee_move *moves_out,
                                int32_t *counts_out, uint16_t *ids_out);
/* apply_move / skip_turn (backgammon_logic.rs:176-196) with the dice of the next roll injected:
 * moves[i] == EMPTY_MOVE -> skip_turn.  next_rolls[2*i..] are used only when the turn passes. */
int32_t diee_bg_apply_moves(diee_ctx *ctx, diee_bg_state *states, const diee_move *moves,
                            const uint8_t *next_rolls, int32_t n);
int32_t diee_bg_apply_moves_dev(diee_ctx *ctx, diee_bg_state *states, const diee_move *moves,
                                const uint8_t *next_rolls, int32_t n);
/* random-vs-random playout of whole games (Agent::Random both sides, versus.rs:160-268,307-316),
 * fused in one launch: game i uses stream DIEE_STREAM_GAME of game id first_game_id+i; stops at a
 * winner or after round_limit plies.  winners_out: -1/+1, 0 at the cap.  finals_out nullable.
 * The _dev form needs its state arrays 16-byte aligned (DIEE_ERR_INVALID otherwise). */
int32_t diee_bg_playout(diee_ctx *ctx, const diee_bg_state *starts, int32_t n, uint64_t seed,
                        uint32_t first_game_id, int32_t round_limit, int8_t *winners_out,
                        int32_t *plies_out, diee_bg_state *finals_out);
int32_t diee_bg_playout_dev(diee_ctx *ctx, const diee_bg_state *starts, int32_t n, uint64_t seed,
                            uint32_t first_game_id, int32_t round_limit, int8_t *winners_out,
                            int32_t *plies_out, diee_bg_state *finals_out);
/* encode / decode one play per state (backgammon_logic.rs:262-401) */
int32_t diee_bg_encode_moves(diee_ctx *ctx, const diee_bg_state *states, const diee_move *moves, int32_t n,
                             uint16_t *ids_out);
int32_t diee_bg_decode_moves(diee_ctx *ctx, const diee_bg_state *states, const uint16_t *ids, int32_t n,
                             diee_move *moves_out);
/* as_tensor (backgammon_logic.rs:198-252): out[i] = f32 [6,4,6] */
int32_t diee_bg_encode_states(diee_ctx *ctx, const diee_bg_state *states, int32_t n, float *out);
int32_t diee_bg_encode_states_dev(diee_ctx *ctx, const diee_bg_state *states, int32_t n, float *out);

/* ---- pure MCTS: mct_search (mcts/simple_mcts.rs:10-39), one independent search per state ----
 * states: diee_bg_state[n] or diee_ttt_state[n] by game_kind; players[i] = the player the value
 * is counted for (versus.rs:305 passes game.get_player()).  Game i has global id first_game_id+i.
 * best_moves_out: diee_move[n] (backgammon) or uint8[n] (tictactoe; 10 = EMPTY_MOVE).
 * status_out[i]: DIEE_OK / DIEE_ERR_NO_MOVES_PANIC / DIEE_ERR_OVERFLOW per game.
 * Optional pool read-back (all nullable): nodes_out[n*(iterations+1)], node_states_out (same
 * count, state type by game_kind), n_nodes_out[n] (game i's live entries are the first n_nodes_out[i]
 * of its iterations+1; the rest are unspecified -- 0 live entries for a root that is already won);
 * stats_out[n] = work counters;
 * rollout_finals_out[n*iterations] = the state each simulation's rollout ended in (all-zero where
 * the iteration ran no rollout), which makes the rollouts themselves checkable against the oracle.
 * With reference-exact rollouts (no DIEE_MODE_ROLLOUT_CHECK_CURRENT) a rollout cannot influence the
 * tree (quirk Q5), so the search runs as two launches: the tree kernel, then every game's
 * rollouts concurrently -- same stream coordinates, same plies, same results. */
int32_t diee_mcts_search(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n,
                         const int8_t *players, const diee_mcts_cfg *cfg, uint64_t seed,
                         uint32_t first_game_id, uint32_t epoch, void *best_moves_out,
                         int32_t *status_out, diee_node *nodes_out, void *node_states_out,
                         int32_t *n_nodes_out, diee_search_stats *stats_out, void *rollout_finals_out);
/* device-resident form: states/players/best_moves/status are device pointers; the node pool
 * lives in the ctx (HBM) and is reused between calls.  stats_dev nullable (diee_search_stats[n]). */
int32_t diee_mcts_search_dev(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n,
                             const int8_t *players, const diee_mcts_cfg *cfg, uint64_t seed,
                             uint32_t first_game_id, uint32_t epoch, void *best_moves_out,
                             int32_t *status_out, diee_search_stats *stats_dev);
/* CUDA-event durations of the two kernels of the last reference-exact backgammon search on this context
 * (tree kernel | all rollouts), for the roofline line of bench.py.  Waits for that search to finish. */
int32_t diee_search_timing(diee_ctx *ctx, float *tree_ms, float *rollout_ms);
/* plies the rollouts of the last backgammon search on this context actually PLAYED.  (A reference-exact rollout runs
 * `simulate_round_limit` plies whatever happens -- quirk Q5 --, but once both sides have collected all 15 checkers the rest
 * are forced passes and the final state is written in closed form: those plies are not executed and not counted.) */
int32_t diee_search_work(diee_ctx *ctx, uint64_t *rollout_plies_played);

/* ---- policy/value net: ResNet (alphazero/nnet.rs:57-155), inference only ----
 * Architecture (backgammon): conv3x3(6->F)+BN+ReLU, `blocks` x [conv+BN+ReLU+conv+BN+add+ReLU],
 * policy head conv3x3(F->32)+BN+ReLU+flatten+Linear(768->1352)+softmax, value head
 * conv3x3(F->3)+BN+ReLU+flatten+Linear(72->1)+tanh; BatchNorm in eval mode (running stats, eps 1e-5).
 * diee_net_create takes the 22 + 12*blocks tensors as host f32 arrays in the reference's registration
 * order (nnet.rs:62-98, ResBlock::new :37-44), one layer after the other:
 *   conv:  weight [co,ci,3,3], bias [co]         batch_norm: weight, bias, running_mean, running_var
 *   linear: weight [out,in], bias [out]
 *   init conv, init bn | per block: conv1, conv2, bn1, bn2 | policy conv, bn, linear | value conv, bn, linear
 * (die_e_b200/nnet.py reads them out of a tch VarStore `.ot` file.)  BatchNorm is folded into the
 * convolutions at load time; the convolutions run on the tensor cores (tcgen05), see diee_net_set_precision.  forward = forward_t (nnet.rs:120-133): policy_out f32 [n,1352] (softmaxed),
 * value_out f32 [n] (tanh).  States are the packed 32-byte states; as_tensor is fused in. */
typedef struct diee_net diee_net;
int32_t diee_net_create(diee_ctx *ctx, int32_t game_kind, const float *const *tensors, const int64_t *numels,
                        int32_t n_tensors, diee_net **out);
/* Arithmetic of the forward pass (nnet.rs:120-133 computes in fp32, lib.rs:20).
 * DIEE_NET_BF16: bf16 operands on the tensor cores, fp32 accumulation -- the fast path, NOT inside the reference's
 *   tolerance through 39 layers (4e-2 on the value head); searches run with it are not comparable to the reference's.
 * DIEE_NET_SPLIT3 (default): the tensor-core mode INSIDE the fp32 tolerance.  Every activation and weight is carried as three
 *   bf16 planes and each product is six MMAs; the leading plane is a 7-bit integer digit under a per-board / per-channel
 *   power-of-two unit, so the sum of the full-size products is exact in the fp32 accumulator whatever the hardware
 *   rounds, and the truncating accumulation only touches terms 2^-7 and smaller (csrc/net_kernels.cu).  Error against
 *   fp64 within a small multiple of an fp32 forward's own (tests/test_gpu_net.py).
 * DIEE_NET_FP32: fp32 FMAs (round to nearest) on the CUDA cores, the reference's own arithmetic; differs from it only
 *   by summation order. */
#define DIEE_NET_BF16 0
#define DIEE_NET_SPLIT3 1
#define DIEE_NET_FP32 2
int32_t diee_net_set_precision(diee_ctx *ctx, diee_net *net, int32_t precision);
int32_t diee_net_destroy(diee_ctx *ctx, diee_net *net);
int64_t diee_net_param_count(const diee_net *net);
int32_t diee_net_forward(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, float *policy_out,
                         float *value_out);
int32_t diee_net_forward_dev(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, float *policy_out,
                             float *value_out);

/* ---- AlphaZero search: alpha_mcts_parallel (mcts/alpha_mcts.rs:91-202) ----
 * Lock-step search over n games with the net evaluated once per iteration for the whole batch.
 * game_ids[i] = global id of game i (keys its random streams); epoch = the wave number (keys the
 * shared Dirichlet vector and the expansion dice).  Returns, per game, the root's children in
 * legal-move order: action ids (encode), moves and visit counts -- what get_prob_tensor_parallel
 * (mcts/utils.rs:42-58) reads.  max_nodes = node slab per game (0 = 1 + (iterations+1)*128: no observed position has more than 123 legal plays);
 * status_out[i] = DIEE_ERR_OVERFLOW if it was exhausted.  nodes_out (nullable, [n*max_nodes]) dumps the pool.
 * Contract where the reference leaves arithmetic to libtorch/rand: masked-policy sums are sequential
 * f32 in legal-move order; games are slots in the order given; diee_dirichlet() is the noise vector. */
typedef struct {
    int32_t parent, first_child, n_children;
    float visits, value, prior;
    diee_move action;
    diee_bg_state state;
} diee_anode;
/* one MemoryFragment (alphazero/alphazero.rs:69-73) in packed form: state instead of the [1,6,4,6] tensor,
 * pi^(1/T) as (action id, weight) pairs pi_ids/pi_vals[pi_offset .. pi_offset+n_pi) */
typedef struct {
    diee_bg_state state;
    uint32_t game_id;
    uint16_t ply;
    int8_t outcome;
    uint8_t pad;
    uint16_t n_pi;
    uint16_t pad2;
    uint32_t pi_offset;
} diee_traj_record;
int32_t diee_dirichlet(uint64_t seed, uint32_t epoch, float alpha, int32_t n, float *out);
int32_t diee_alpha_search(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                          const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, uint16_t *root_ids_out,
                          diee_move *root_moves_out, float *root_visits_out, int32_t *root_counts_out, int32_t *status_out,
                          diee_anode *nodes_out, int32_t *n_nodes_out);
int32_t diee_alpha_search_dev(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                              const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, uint16_t *root_ids_out,
                              diee_move *root_moves_out, float *root_visits_out, int32_t *root_counts_out, int32_t *status_out);
/* self_play_parallel (alphazero/alpha_parallel.rs:101-231): n_games games from the opening position in
 * lock-step until every game has a winner or hit cfg->simulate_round_limit; records are appended in the
 * reference's emission order (quirk Q10 included).  Game i has id first_game_id + i. */
int32_t diee_selfplay_run(diee_ctx *ctx, diee_net *net, int32_t n_games, const diee_mcts_cfg *cfg, float temperature,
                          uint64_t seed, uint32_t first_game_id, int32_t max_nodes, diee_traj_record *rec_out, int32_t rec_cap,
                          uint16_t *pi_ids_out, float *pi_vals_out, int32_t pi_cap, int32_t *n_rec_out, int32_t *n_pi_out,
                          int32_t *n_waves_out);
/* The same driver with options (all zero / NULL = diee_selfplay_run).  max_waves > 0 time-boxes the run: after that
 * many game-move waves the games still running emit what they have recorded with outcome 0 (a bounded sample of the
 * same work for benchmarks; not a reference behaviour).  The other fields select the NON-PARITY throughput modes of
 * DIEE_SP_REFILL (needs target_games or max_waves) and leaves_per_game > 1 (diee_alpha_search_vl).  With
 * all of them zero the run is the reference's self_play_parallel, record for record. */
typedef struct {
    uint32_t flags;          /* DIEE_SP_* */
    int32_t max_waves;       /* 0 = until every game has ended */
    int32_t leaves_per_game; /* 0/1 = one leaf per game and iteration, as the reference (alpha_mcts.rs:149-201) */
    int32_t target_games;    /* DIEE_SP_REFILL: stop once this many games have finished */
    float virtual_loss;      /* leaves_per_game > 1 */
    uint32_t pad;
} diee_selfplay_opts;
#define DIEE_SP_REFILL 1u /* a finished game's slot starts a new game, so the forward batch stays at n_games */
typedef struct {
    int32_t waves, games_finished, games_cut, pad;
    uint64_t game_moves; /* searches run: one per live game and wave */
} diee_selfplay_report;
int32_t diee_selfplay_run_ex(diee_ctx *ctx, diee_net *net, int32_t n_games, const diee_mcts_cfg *cfg, float temperature,
                             uint64_t seed, uint32_t first_game_id, int32_t max_nodes, const diee_selfplay_opts *opts,
                             diee_traj_record *rec_out, int32_t rec_cap, uint16_t *pi_ids_out, float *pi_vals_out, int32_t pi_cap,
                             int32_t *n_rec_out, int32_t *n_pi_out, int32_t *n_waves_out, diee_selfplay_report *report_out);
/* NON-PARITY search: up to leaves_per_game leaves per game and step, each descent leaving a virtual loss
 * (visits + 1, value - virtual_loss) on its path; cfg->iterations / leaves_per_game forwards of n x leaves_per_game
 * boards instead of cfg->iterations forwards of n.  Same outputs as diee_alpha_search.  The reference selects one leaf per
 * game and iteration (alpha_mcts.rs:149-201); with leaves_per_game = 1, virtual_loss = 0 and no terminal leaf in reach
 * this IS that search (tests/test_gpu_alpha.py). */
int32_t diee_alpha_search_vl(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                             const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, int32_t leaves_per_game,
                             float virtual_loss, uint16_t *root_ids_out, diee_move *root_moves_out, float *root_visits_out,
                             int32_t *root_counts_out, int32_t *status_out);
uint64_t diee_net_eval_count(const diee_ctx *ctx);

/* ---- the arena: versus::play (src/versus.rs:160-268) with the games resident on the device ----
 * n_games games from the opening position (those of the second half start with skip_turn, versus.rs:172-174), streams
 * keyed by the game's index (DIEE_STREAM_INIT / _GAME / search streams with epoch = round).  One diee_arena_round = one
 * pass of the reference's loop: partition by the side to move, each side's agent picks an action for every one of its
 * games (Agent::Mcts = mct_search, Agent::Random = uniform over the legal moves), apply_move / skip_turn, winner and
 * round-limit tests, retirement.  Nothing but the 5-word summary crosses the bus per round:
 * summary_out = {games retired so far, wins of player 1 (the -1 side), wins of player 2, draws, 0}.  The arena is over
 * when summary_out[0] == n_games.  diee_arena_read gives the per-game results (all nullable). */
typedef struct diee_arena diee_arena;
enum { DIEE_AGENT_RANDOM = 0, DIEE_AGENT_MCTS = 1 };
int32_t diee_arena_create(diee_ctx *ctx, int32_t n_games, uint64_t seed, int32_t round_limit, diee_arena **out);
int32_t diee_arena_round(diee_ctx *ctx, diee_arena *arena, int32_t agent_p1, int32_t agent_p2, const diee_mcts_cfg *cfg,
                         int32_t *summary_out);
int32_t diee_arena_read(diee_ctx *ctx, diee_arena *arena, diee_bg_state *states_out, int8_t *winners_out, int32_t *rounds_out);
int32_t diee_arena_destroy(diee_ctx *ctx, diee_arena *arena);

/* ---- multi-GPU: the one exchange step of the path ----
 * Games never interact (the reference runs them as independent rayon tasks, versus.rs:303-316): each GPU plays
 * its own shard of game ids and no collective sits on the data path.  What is exchanged is the OUTPUT: finished
 * self-play trajectories are all-gathered into every rank's replay buffer (the cumulative `memory` of
 * alpha_parallel.rs:53), and a promoted model's weights are broadcast.  One rank per context / GPU, NCCL over
 * NVLink; NCCL is bound at run time (libnccl.so.2), DIEE_ERR_CUDA if it is not there.
 *   diee_comm_unique_id: rank 0 draws the rendezvous id and ships it to the other ranks by any means.
 *   diee_traj_allgather: every rank contributes its records (host buffers); all receive all of them, rank-major,
 *     with pi_offset rebased into the concatenated pi arrays.  *n_rec_out / *n_pi_out are the totals (set even when
 *     they exceed the capacities and DIEE_ERR_OVERFLOW is returned).
 *   diee_net_broadcast: the tensor list of diee_net_create, in place: on return every rank holds root's values. */
#define DIEE_COMM_ID_BYTES 128
int32_t diee_comm_unique_id(uint8_t *id_out);
int32_t diee_comm_init(diee_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t *id);
int32_t diee_comm_destroy(diee_ctx *ctx);
int32_t diee_traj_allgather(diee_ctx *ctx, const diee_traj_record *rec, int32_t n_rec, const uint16_t *pi_ids, const float *pi_vals,
                            int32_t n_pi, diee_traj_record *rec_out, int32_t rec_cap, uint16_t *pi_ids_out, float *pi_vals_out,
                            int32_t pi_cap, int32_t *n_rec_out, int32_t *n_pi_out);
int32_t diee_net_broadcast(diee_ctx *ctx, float *const *tensors, const int64_t *numels, int32_t n_tensors, int32_t root);

#ifdef __cplusplus
}
#endif
#endif
