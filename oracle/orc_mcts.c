/*
 * oracle/orc_mcts.c -- CPU ORACLE (test infrastructure only; see orc.h).
 * Restates src/mcts/simple_mcts.rs, src/mcts/node.rs, src/mcts/node_store.rs of
 * alibasaran/die-e (pure MCTS: UCB1 select, expand-last-untried, capped random rollout,
 * un-flipped backprop, most-visited answer), generically over the two games.
 *
 * Parity unpinned by the reference (no test pins visits/values/rollouts): this file,
 * reviewed against the cited lines, is the pin.  f32 arithmetic is IEEE single with no
 * contraction (build with -ffp-contract=off); ln is (float)log((double)x).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

float orc_ln_f32(float x) { return (float)log((double)x); }

typedef struct {
    int state_size;
    int deterministic;
    int (*valid_moves)(const void *s, orc_move *out);
    void (*apply)(void *s, orc_move m, uint8_t d0, uint8_t d1);
    void (*skip)(void *s, uint8_t d0, uint8_t d1);
    int (*winner)(const void *s);
} game_vt;

/* ---- adapters ---- */
static int bg_valid(const void *s, orc_move *out) { return orc_bg_valid_moves((const orc_bg_state *)s, out, ORC_MAX_MOVES); }
static void bg_apply(void *s, orc_move m, uint8_t d0, uint8_t d1) { orc_bg_apply_move((orc_bg_state *)s, m, d0, d1); }
static void bg_skip(void *s, uint8_t d0, uint8_t d1) { orc_bg_skip_turn((orc_bg_state *)s, d0, d1); }
static int bg_winner(const void *s) { return orc_bg_check_winner((const orc_bg_state *)s); }
static const game_vt BG = {sizeof(orc_bg_state), 0, bg_valid, bg_apply, bg_skip, bg_winner};

static int ttt_valid(const void *s, orc_move *out) {
    uint8_t cells[9];
    int n = orc_ttt_valid_moves((const orc_ttt_state *)s, cells);
    for (int i = 0; i < n; ++i) out[i] = (orc_move){(int8_t)cells[i], ORC_NONE, ORC_NONE, ORC_NONE};
    return n;
}
static void ttt_apply(void *s, orc_move m, uint8_t d0, uint8_t d1) { (void)d0; (void)d1; orc_ttt_apply_move((orc_ttt_state *)s, (uint8_t)m.from1); }
static void ttt_skip(void *s, uint8_t d0, uint8_t d1) { (void)d0; (void)d1; orc_ttt_skip_turn((orc_ttt_state *)s); }
static int ttt_winner(const void *s) { return orc_ttt_check_winner((const orc_ttt_state *)s); }
static const game_vt TTT = {sizeof(orc_ttt_state), 1, ttt_valid, ttt_apply, ttt_skip, ttt_winner};

/* Node<T>  node.rs:9-19 */
typedef struct {
    unsigned char state[32];
    int parent;
    int *children;
    int n_children;
    float visits, value;
    orc_move action;
    orc_move *moves; /* expandable_moves (untried = first n_untried entries; popped from the end) */
    int n_moves, n_untried;
} node_t;

typedef struct {
    const game_vt *g;
    node_t *nodes;
    int n, cap;
    uint32_t mode;
} store_t;

static const orc_move EMPTY = {ORC_NONE, ORC_NONE, ORC_NONE, ORC_NONE};

/* NodeStore::add_node -> Node::new  node_store.rs:34-45, node.rs:43-62 (moves computed eagerly) */
static int add_node(store_t *st, const void *state, int parent, orc_move action) {
    if (st->n >= st->cap) return -1;
    node_t *nd = &st->nodes[st->n];
    memset(nd, 0, sizeof *nd);
    memcpy(nd->state, state, (size_t)st->g->state_size);
    nd->parent = parent;
    nd->action = action;
    orc_move tmp[ORC_MAX_MOVES];
    int n = st->g->valid_moves(state, tmp);
    if (n < 0) n = 0;
    if (n == 0 && (st->mode & ORC_MODE_PASS_CHILD)) { tmp[0] = EMPTY; n = 1; } /* Q6 PASS_CHILD */
    nd->moves = (orc_move *)malloc(sizeof(orc_move) * (size_t)(n ? n : 1));
    memcpy(nd->moves, tmp, sizeof(orc_move) * (size_t)n);
    nd->children = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    nd->n_moves = n;
    nd->n_untried = n;
    return st->n++;
}

/* Node::ucb  node.rs:86-96 (c INSIDE the sqrt; strict f32 evaluation order) */
static float ucb(const store_t *st, const node_t *nd, float c) {
    if (nd->parent < 0) return INFINITY;
    const node_t *p = &st->nodes[nd->parent];
    volatile float exploitation = nd->value / nd->visits;
    volatile float t = c * orc_ln_f32(p->visits);
    t = t / nd->visits;
    volatile float exploration = sqrtf(t);
    return exploitation + exploration;
}

/* select_ucb  simple_mcts.rs:41-52.  Iterator::max_by folds keeping the LATER element unless
 * the earlier compares Greater; partial_cmp None (NaN) counts as Equal. */
static int select_ucb(const store_t *st, int idx, float c) {
    const node_t *nd = &st->nodes[idx];
    int best = nd->children[0];
    float bs = ucb(st, &st->nodes[best], c);
    for (int i = 1; i < nd->n_children; ++i) {
        int ch = nd->children[i];
        float s = ucb(st, &st->nodes[ch], c);
        if (!(bs > s)) { best = ch; bs = s; }
    }
    return best;
}

/* select_leaf_node  simple_mcts.rs:88-94 */
static int select_leaf(const store_t *st, int idx, float c) {
    for (;;) {
        const node_t *nd = &st->nodes[idx];
        if (nd->n_children == 0 || nd->n_untried != 0) return idx;
        idx = select_ucb(st, idx, c);
    }
}

/* backpropagate  simple_mcts.rs:96-103 (no sign flip, Q7) */
static void backprop(store_t *st, int idx, float result) {
    while (idx >= 0) {
        st->nodes[idx].visits += 1.0f;
        st->nodes[idx].value += result;
        idx = st->nodes[idx].parent;
    }
}

static float outcome(int winner, int player) { /* simple_mcts.rs:26-28, node.rs:182-184 */
    if (winner == player) return 1.0f;
    if (winner == -player) return -1.0f;
    return 0.0f;
}

/* Node::simulate  node.rs:176-196.  Default (reference-exact) tests the winner of the START
 * state each iteration (Q5); ORC_MODE_ROLLOUT_CHECK_CURRENT tests the rolled-out state. */
static float simulate(const store_t *st, const node_t *nd, int player, uint32_t limit, uint64_t seed,
                      uint32_t game_id, uint32_t c3, uint64_t *plies_acc, unsigned char *final_out) {
    unsigned char curbuf[32];
    unsigned char *cur = final_out ? final_out : curbuf; /* the rolled-out state is observable for parity tests */
    memcpy(cur, nd->state, 32);
    orc_move mv[ORC_MAX_MOVES];
    for (uint32_t k = 0; k < limit; ++k) {
        int w = st->g->winner((st->mode & ORC_MODE_ROLLOUT_CHECK_CURRENT) ? (const void *)cur : (const void *)nd->state);
        if (w != ORC_NO_WINNER) return outcome(w, player);
        uint32_t blk[4];
        orc_philox(seed, k, game_id, ORC_STREAM_ROLLOUT, c3, blk);
        int n = st->g->valid_moves(cur, mv);
        uint8_t d0 = orc_die(blk[0]), d1 = orc_die(blk[1]);
        if (n > 0) st->g->apply(cur, mv[orc_index(blk[2], (uint32_t)n)], d0, d1);
        else st->g->skip(cur, d0, d1);
        if (plies_acc) *plies_acc += 1;
    }
    return 0.0f;
}

static int search(const game_vt *g, const void *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                  uint32_t game_id, uint32_t epoch, orc_move *best, orc_node_stats *nodes_out,
                  void *states_out, int32_t *n_nodes_out, void *rollout_finals_out) {
    *best = EMPTY;
    if (n_nodes_out) *n_nodes_out = 0;
    if (g->winner(root) != ORC_NO_WINNER) return ORC_OK; /* simple_mcts.rs:12-14 */
    store_t st;
    st.g = g;
    st.cap = (int)cfg->iterations + 1;
    st.nodes = (node_t *)calloc((size_t)st.cap, sizeof(node_t));
    st.n = 0;
    st.mode = cfg->mode_flags;
    int rc = ORC_OK;
    add_node(&st, root, -1, EMPTY);
    for (uint32_t it = 0; it < cfg->iterations; ++it) { /* :20-37 */
        int sel = select_leaf(&st, 0, cfg->c);
        node_t *nd = &st.nodes[sel];
        int w = g->winner(nd->state);
        if (w != ORC_NO_WINNER) {
            backprop(&st, sel, outcome(w, player));
            continue;
        }
        /* Node::expand  node.rs:118-137 */
        if (nd->n_untried == 0) { rc = ORC_ERR_NO_MOVES_PANIC; break; } /* Q6 panic */
        orc_move a = nd->moves[--nd->n_untried]; /* pop() = last */
        unsigned char next[32];
        memcpy(next, nd->state, 32);
        uint32_t blk[4];
        orc_philox(seed, (uint32_t)st.n, game_id, ORC_STREAM_EXPAND, epoch, blk);
        uint8_t d0 = orc_die(blk[0]), d1 = orc_die(blk[1]);
        if (a.from1 == ORC_NONE && !g->deterministic) g->skip(next, d0, d1); /* PASS_CHILD */
        else if (a.from1 == ORC_NONE) g->skip(next, 0, 0);
        else g->apply(next, a, d0, d1);
        int child = add_node(&st, next, sel, a);
        nd = &st.nodes[sel];
        nd->children[nd->n_children++] = child;
        unsigned char fin[32];
        memset(fin, 0, sizeof fin);
        float v = simulate(&st, &st.nodes[child], player, cfg->simulate_round_limit, seed, game_id,
                           (epoch << 16) | (it & 0xFFFFu), NULL, fin);
        if (rollout_finals_out) memcpy((unsigned char *)rollout_finals_out + (size_t)it * (size_t)g->state_size, fin, (size_t)g->state_size);
        backprop(&st, child, v);
    }
    if (rc == ORC_OK) { /* select_most_visits  :71-86 */
        const node_t *r = &st.nodes[0];
        if (r->n_children > 0) {
            int b = r->children[0];
            for (int i = 1; i < r->n_children; ++i) {
                int ch = r->children[i];
                if (!(st.nodes[b].visits > st.nodes[ch].visits)) b = ch;
            }
            *best = st.nodes[b].action;
        }
    }
    for (int i = 0; i < st.n; ++i) {
        if (nodes_out) {
            nodes_out[i].parent = st.nodes[i].parent;
            nodes_out[i].visits = st.nodes[i].visits;
            nodes_out[i].value = st.nodes[i].value;
            nodes_out[i].action = st.nodes[i].action;
            nodes_out[i].n_moves = st.nodes[i].n_moves;
            nodes_out[i].n_untried = st.nodes[i].n_untried;
        }
        if (states_out) memcpy((unsigned char *)states_out + (size_t)i * (size_t)g->state_size, st.nodes[i].state, (size_t)g->state_size);
        free(st.nodes[i].moves);
        free(st.nodes[i].children);
    }
    if (n_nodes_out) *n_nodes_out = st.n;
    free(st.nodes);
    return rc;
}

/* mct_search  simple_mcts.rs:10-39 */
int orc_mcts_search_bg(const orc_bg_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                       uint32_t game_id, uint32_t epoch, orc_move *best, orc_node_stats *nodes_out,
                       orc_bg_state *states_out, int32_t *n_nodes_out) {
    return search(&BG, root, player, cfg, seed, game_id, epoch, best, nodes_out, states_out, n_nodes_out, NULL);
}

/* same, also returning the state each simulation's rollout ended in (iterations entries; zero where no rollout ran) */
int orc_mcts_search_bg_ex(const orc_bg_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                          uint32_t game_id, uint32_t epoch, orc_move *best, orc_node_stats *nodes_out,
                          orc_bg_state *states_out, int32_t *n_nodes_out, orc_bg_state *rollout_finals_out) {
    if (rollout_finals_out) memset(rollout_finals_out, 0, sizeof(orc_bg_state) * cfg->iterations);
    return search(&BG, root, player, cfg, seed, game_id, epoch, best, nodes_out, states_out, n_nodes_out, rollout_finals_out);
}

int orc_mcts_search_ttt(const orc_ttt_state *root, int player, const orc_mcts_cfg *cfg, uint64_t seed,
                        uint32_t game_id, uint32_t epoch, uint8_t *best, orc_node_stats *nodes_out,
                        orc_ttt_state *states_out, int32_t *n_nodes_out) {
    orc_move b;
    int rc = search(&TTT, root, player, cfg, seed, game_id, epoch, &b, nodes_out, states_out, n_nodes_out, NULL);
    *best = b.from1 == ORC_NONE ? 10 : (uint8_t)b.from1; /* EMPTY_MOVE = 10  tictactoe/mod.rs:18 */
    return rc;
}
