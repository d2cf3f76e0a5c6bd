/*
 * oracle/orc_alpha.c -- CPU ORACLE (test infrastructure only; see orc.h).
 * Restates the AlphaZero search and self-play driver of alibasaran/die-e:
 *   src/mcts/alpha_mcts.rs   alpha_select_leaf_node :14-20, select_alpha :22-33, alpha_mcts_parallel :91-202
 *   src/mcts/node.rs         alpha_ucb :98-112, alpha_expand_tensor :157-174
 *   src/mcts/utils.rs        get_prob_tensor_parallel :42-58, turn_policy_to_probs_tensor(_parallel) :60-84
 *   src/mcts/noise.rs        apply_dirichlet :27-34
 *   src/alphazero/alpha_parallel.rs  self_play_parallel :101-231
 *   src/alphazero/alphazero.rs       weighted_select_tensor_idx :129-137
 * including the quirks Q8-Q12.  The net is a callback, so a test can inject
 * the same net outputs into the oracle and into the CUDA path ("visit counts given identical
 * net outputs").  PARITY UNPINNED by the reference (it has no test for any of this): this file,
 * reviewed against the cited lines, is the pin.
 *
 * Contract choices where the reference leaves the arithmetic to libtorch / rand (documented in
 * include/diee.h): masked-policy sums and visit sums are sequential f32 in legal-move / child order;
 * games are processed in ascending index order; Dirichlet(alpha) = normalised Gamma(alpha,1) draws
 * (Marsaglia-Tsang with the alpha<1 boost) in double on the DIRICHLET Philox stream; the categorical
 * draw is a 53-bit uniform times the f64 total against f64 cumulative weights in action-id order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc.h"

#define A_SPACE 1352

/* ---------------- Dirichlet noise (noise.rs:27-34; rand_distr 0.4 Dirichlet = normalised Gammas) ---------------- */
static double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 21) ^ ((uint64_t)lo >> 11)) & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0); }

static double gamma_draw(uint64_t seed, uint32_t epoch, uint32_t comp, double alpha) {
    /* Marsaglia & Tsang (2000); for alpha < 1: Gamma(alpha) = Gamma(alpha+1) * U^(1/alpha) */
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t attempt = 0;; ++attempt) {
        uint32_t w[4], w2[4];
        orc_philox(seed, 2 * attempt, epoch, ORC_STREAM_DIRICHLET, comp, w);
        orc_philox(seed, 2 * attempt + 1, epoch, ORC_STREAM_DIRICHLET, comp, w2);
        double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
        if (u1 <= 0.0) u1 = 1.0 / 9007199254740992.0;
        const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2); /* Box-Muller */
        const double v0 = 1.0 + c * x;
        if (v0 <= 0.0) continue;
        const double v = v0 * v0 * v0;
        double u = u53(w2[0], w2[1]);
        if (u <= 0.0) u = 1.0 / 9007199254740992.0;
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
            double g = d * v;
            if (alpha < 1.0) {
                double ub = u53(w2[2], w2[3]);
                if (ub <= 0.0) ub = 1.0 / 9007199254740992.0;
                g *= pow(ub, 1.0 / alpha);
            }
            return g;
        }
    }
}

void orc_dirichlet(uint64_t seed, uint32_t epoch, float alpha, int n, float *out) {
    double *g = (double *)malloc(sizeof(double) * (size_t)n);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) { g[i] = gamma_draw(seed, epoch, (uint32_t)i, (double)alpha); sum += g[i]; }
    for (int i = 0; i < n; ++i) out[i] = (float)(g[i] / sum);
    free(g);
}

/* ---------------- the arena: one slab per game, children contiguous ---------------- */
typedef struct {
    orc_anode *nodes; /* [cap] */
    int n, cap;
    int overflow;
} tree_t;

static int is_terminal(const orc_bg_state *s) { return orc_bg_check_winner(s) != ORC_NO_WINNER; }

/* Node::alpha_ucb node.rs:98-112 in strict f32 */
static float alpha_ucb(const tree_t *t, int idx, float c) {
    const orc_anode *nd = &t->nodes[idx];
    volatile float q = nd->visits == 0.0f ? 0.0f : nd->value / nd->visits;
    if (nd->parent < 0) return INFINITY;
    const orc_anode *p = &t->nodes[nd->parent];
    volatile float s = sqrtf(p->visits);
    volatile float den = nd->visits + 1.0f;
    volatile float r = s / den;
    volatile float e = c * r;
    volatile float ep = e * nd->prior;
    return q + ep;
}

/* alpha_select_leaf_node :14-20 with select_alpha :22-33 (max_by keeps the LAST maximum) */
static int alpha_select_leaf(const tree_t *t, float c) {
    int idx = 0;
    for (;;) {
        const orc_anode *nd = &t->nodes[idx];
        if (nd->n_children == 0) return idx;
        int best = nd->first_child;
        float bs = alpha_ucb(t, best, c);
        for (int k = 1; k < nd->n_children; ++k) {
            const int ch = nd->first_child + k;
            const float s = alpha_ucb(t, ch, c);
            if (!(bs > s)) { best = ch; bs = s; }
        }
        idx = best;
    }
}

static void backprop(tree_t *t, int idx, float v) { /* simple_mcts.rs:96-103 */
    while (idx >= 0) {
        t->nodes[idx].visits += 1.0f;
        t->nodes[idx].value += v;
        idx = t->nodes[idx].parent;
    }
}

/* alpha_expand_tensor node.rs:157-174 with the masked, renormalised policy of utils.rs:60-84.
 * `row` is the (possibly Dirichlet-mixed) policy row of this node. */
static void expand(tree_t *t, int idx, const float *row, uint64_t seed, uint32_t gid, uint32_t epoch) {
    orc_anode *nd = &t->nodes[idx];
    if (nd->n_children != 0) return; /* expandable_moves already drained */
    orc_move mv[ORC_MAX_MOVES];
    int n = orc_bg_valid_moves(&nd->state, mv, ORC_MAX_MOVES);
    if (n <= 0) return; /* Q12: nothing to expand, the leaf is re-selected every iteration */
    if (t->n + n > t->cap) { t->overflow = 1; return; }
    float sel[ORC_MAX_MOVES];
    volatile float sum = 0.0f;
    for (int k = 0; k < n; ++k) {
        sel[k] = row[orc_bg_encode(&nd->state, mv[k])];
        sum = sum + sel[k]; /* sequential f32 in legal-move order (contract) */
    }
    const int first = t->n;
    nd->first_child = first;
    nd->n_children = n;
    for (int k = 0; k < n; ++k) {
        orc_anode *ch = &t->nodes[first + k];
        memset(ch, 0, sizeof *ch);
        ch->parent = idx;
        ch->first_child = -1;
        ch->action = mv[k];
        ch->prior = sel[k] / sum;
        ch->state = nd->state;
        uint32_t w[4];
        orc_philox(seed, (uint32_t)(first + k), gid, ORC_STREAM_EXPAND, epoch, w); /* dice frozen in the child, Q13 */
        orc_bg_apply_move(&ch->state, mv[k], orc_die(w[0]), orc_die(w[1]));
    }
    t->n += n;
}

/* alpha_mcts_parallel  alpha_mcts.rs:91-202 */
int orc_alpha_mcts_parallel(const orc_bg_state *states, int n, const uint32_t *game_ids, const orc_mcts_cfg *cfg,
                            uint64_t seed, uint32_t epoch, orc_eval_fn eval, void *user, int max_nodes,
                            orc_anode *nodes_out, int32_t *n_nodes_out, int32_t *status_out) {
    if (n <= 0) return ORC_OK;
    tree_t *tr = (tree_t *)calloc((size_t)n, sizeof(tree_t));
    float *policy = (float *)malloc(sizeof(float) * (size_t)n * A_SPACE);
    float *value = (float *)malloc(sizeof(float) * (size_t)n);
    orc_bg_state *batch = (orc_bg_state *)malloc(sizeof(orc_bg_state) * (size_t)n);
    float *dir = (float *)malloc(sizeof(float) * A_SPACE);
    int *sel_game = (int *)calloc((size_t)n, sizeof(int)); /* selected_nodes_idxs = vec![0; n]  (:142): arena node 0 */
    int *sel_node = (int *)calloc((size_t)n, sizeof(int)); /* = game 0's root, for EVERY game (Q9) */
    for (int g = 0; g < n; ++g) {
        tr[g].nodes = nodes_out + (size_t)g * (size_t)max_nodes;
        tr[g].cap = max_nodes;
        tr[g].n = 1;
        orc_anode *r = &tr[g].nodes[0];
        memset(r, 0, sizeof *r);
        r->parent = -1;
        r->first_child = -1;
        r->action = (orc_move){ORC_NONE, ORC_NONE, ORC_NONE, ORC_NONE};
        r->state = states[g];
    }
    /* root phase :97-127: forward_policy, shared Dirichlet sample mixed BEFORE masking (Q11) */
    eval(states, n, policy, value, user);
    orc_dirichlet(seed, epoch, cfg->dirichlet_alpha, A_SPACE, dir);
    {
        volatile float a = 1.0f - cfg->dirichlet_epsilon;
        for (int g = 0; g < n; ++g) {
            float *row = policy + (size_t)g * A_SPACE;
            for (int j = 0; j < A_SPACE; ++j) {
                volatile float x = a * row[j];
                volatile float y = cfg->dirichlet_epsilon * dir[j];
                row[j] = x + y;
            }
            tr[g].nodes[0].visits = 1.0f; /* :123 */
            expand(&tr[g], 0, row, seed, game_ids[g], epoch);
        }
    }
    for (uint32_t it = 0; it < cfg->iterations; ++it) { /* :149-201 */
        int any = 0;
        for (int g = 0; g < n; ++g) {
            const int leaf = alpha_select_leaf(&tr[g], cfg->c);
            const int w = orc_bg_check_winner(&tr[g].nodes[leaf].state);
            if (w != ORC_NO_WINNER) {
                const int rp = tr[g].nodes[0].state.player; /* value w.r.t. the ROOT player :157-163 */
                backprop(&tr[g], leaf, w == rp ? 1.0f : (w == -rp ? -1.0f : 0.0f));
            } else {
                any = 1;
                sel_game[g] = g;
                sel_node[g] = leaf;
            }
        }
        if (!any) continue; /* :171-173 */
        for (int g = 0; g < n; ++g) batch[g] = tr[sel_game[g]].nodes[sel_node[g]].state; /* stale entries included (Q9) */
        eval(batch, n, policy, value, user);
        for (int g = 0; g < n; ++g) { /* :192-200, in slot order */
            tree_t *t = &tr[sel_game[g]];
            expand(t, sel_node[g], policy + (size_t)g * A_SPACE, seed, game_ids[sel_game[g]], epoch);
            backprop(t, sel_node[g], value[g]);
        }
    }
    for (int g = 0; g < n; ++g) {
        n_nodes_out[g] = tr[g].n;
        status_out[g] = tr[g].overflow ? ORC_ERR_OVERFLOW : ORC_OK;
    }
    (void)is_terminal;
    free(tr); free(policy); free(value); free(batch); free(dir); free(sel_game); free(sel_node);
    return ORC_OK;
}

/* get_prob_tensor_parallel utils.rs:42-58 + pow_(1/T) alpha_parallel.rs:164-166, sparse over the root's children.
 * ids_out/pi_out: per child in child order.  Returns the number of children. */
int orc_root_pi(const orc_anode *nodes, float temperature_inv, uint16_t *ids_out, float *pi_out) {
    const orc_anode *r = &nodes[0];
    volatile float sum = 0.0f;
    for (int k = 0; k < r->n_children; ++k) sum = sum + nodes[r->first_child + k].visits;
    for (int k = 0; k < r->n_children; ++k) {
        const orc_anode *ch = &nodes[r->first_child + k];
        ids_out[k] = (uint16_t)orc_bg_encode(&r->state, ch->action);
        volatile float p = ch->visits / sum;
        pi_out[k] = powf(p, temperature_inv);
    }
    return r->n_children;
}

/* weighted_select_tensor_idx alphazero.rs:129-137: categorical draw over the dense [1352] weights
 * (f64 cumulative sums in action-id order; first index whose cumulative weight exceeds the draw) */
int orc_weighted_select(const uint16_t *ids, const float *pi, int n, uint64_t seed, uint32_t game_id, uint32_t ply) {
    double dense[A_SPACE];
    memset(dense, 0, sizeof dense);
    for (int k = 0; k < n; ++k) dense[ids[k]] = (double)pi[k];
    double total = 0.0;
    for (int j = 0; j < A_SPACE; ++j) total += dense[j];
    uint32_t w[4];
    orc_philox(seed, ply, game_id, ORC_STREAM_SAMPLE, 0, w);
    const double chosen = u53(w[0], w[1]) * total;
    double cum = 0.0;
    int last = -1;
    for (int j = 0; j < A_SPACE; ++j) {
        if (dense[j] == 0.0) continue;
        cum += dense[j];
        last = j;
        if (cum > chosen) return j;
    }
    return last;
}

/* self_play_parallel  alpha_parallel.rs:101-231.  Records are appended in emission order (a game
 * capped by the round limit AND finished in the same pass is emitted twice, Q10). */
int orc_self_play(int n_games, const orc_mcts_cfg *cfg, float temperature, uint64_t seed, uint32_t first_game_id,
                  orc_eval_fn eval, void *user, int max_nodes, orc_traj_record *rec_out, int rec_cap, uint16_t *pi_ids_out,
                  float *pi_vals_out, int pi_cap, int *n_rec_out, int *n_pi_out, int *n_waves_out) {
    typedef struct { orc_bg_state st; int8_t player; int pi_off, pi_n; int ply; } mem_t;
    orc_bg_state *st = (orc_bg_state *)malloc(sizeof(orc_bg_state) * (size_t)n_games);
    int *n_rounds = (int *)calloc((size_t)n_games, sizeof(int));
    int *alive = (int *)malloc(sizeof(int) * (size_t)n_games);
    mem_t **mem = (mem_t **)calloc((size_t)n_games, sizeof(mem_t *));
    int *mem_n = (int *)calloc((size_t)n_games, sizeof(int)), *mem_cap = (int *)calloc((size_t)n_games, sizeof(int));
    /* scratch pool of every pi ever recorded (indexed by mem_t.pi_off) */
    int sp_cap = 1 << 16, sp_n = 0;
    uint16_t *sp_ids = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)sp_cap);
    float *sp_vals = (float *)malloc(sizeof(float) * (size_t)sp_cap);
    int n_rec = 0, n_pi = 0, rc = ORC_OK, waves = 0;
    const float tinv = (float)(1.0 / (double)temperature);
    for (int g = 0; g < n_games; ++g) { /* :103-111 */
        orc_bg_new(&st[g]);
        uint32_t w[4];
        orc_philox(seed, 0, first_game_id + (uint32_t)g, ORC_STREAM_INIT, 0, w);
        st[g].roll[0] = orc_die(w[0]);
        st[g].roll[1] = orc_die(w[1]);
        alive[g] = 1;
    }
    orc_bg_state *live = (orc_bg_state *)malloc(sizeof(orc_bg_state) * (size_t)n_games);
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_games);
    int *idx_of = (int *)malloc(sizeof(int) * (size_t)n_games);
    orc_anode *nodes = (orc_anode *)malloc(sizeof(orc_anode) * (size_t)n_games * (size_t)max_nodes);
    int32_t *nn = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_games), *status = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_games);
#define EMIT(G, RELABEL, WINNER)                                                                              \
    for (int m_ = 0; m_ < mem_n[G]; ++m_) {                                                                   \
        const mem_t *mm = &mem[G][m_];                                                                        \
        if (n_rec >= rec_cap || n_pi + mm->pi_n > pi_cap) { rc = ORC_ERR_OVERFLOW; break; }                   \
        orc_traj_record *r_ = &rec_out[n_rec++];                                                              \
        r_->state = mm->st; r_->game_id = first_game_id + (uint32_t)(G); r_->ply = (uint16_t)mm->ply;         \
        r_->outcome = (int8_t)((RELABEL) ? ((WINNER) == mm->player ? 1 : ((WINNER) == -mm->player ? -1 : 0)) : 0); \
        r_->pad = 0; r_->n_pi = (uint16_t)mm->pi_n; r_->pad2 = 0; r_->pi_offset = (uint32_t)n_pi;             \
        memcpy(pi_ids_out + n_pi, sp_ids + mm->pi_off, sizeof(uint16_t) * (size_t)mm->pi_n);                  \
        memcpy(pi_vals_out + n_pi, sp_vals + mm->pi_off, sizeof(float) * (size_t)mm->pi_n);                   \
        n_pi += mm->pi_n;                                                                                     \
    }
    for (;;) {
        int nl = 0;
        for (int g = 0; g < n_games; ++g)
            if (alive[g]) { live[nl] = st[g]; ids[nl] = first_game_id + (uint32_t)g; idx_of[nl] = g; ++nl; }
        if (nl == 0 || rc != ORC_OK) break;
        orc_alpha_mcts_parallel(live, nl, ids, cfg, seed, (uint32_t)waves, eval, user, max_nodes, nodes, nn, status);
        ++waves;
        for (int pi = 0; pi < nl; ++pi) { /* :171-223 */
            const int g = idx_of[pi];
            if (status[pi] != ORC_OK) { rc = status[pi]; break; }
            const orc_anode *tn = nodes + (size_t)pi * (size_t)max_nodes;
            uint16_t cid[ORC_MAX_MOVES];
            float cpi[ORC_MAX_MOVES];
            const int nc = orc_root_pi(tn, tinv, cid, cpi);
            if (n_rounds[g] >= (int)cfg->simulate_round_limit) { /* :172-180, NO continue (Q10) */
                EMIT(g, 0, 0)
                alive[g] = 0;
            }
            double dsum = 0.0;
            for (int k = 0; k < nc; ++k) dsum += (double)cpi[k];
            uint32_t w[4];
            orc_philox(seed, (uint32_t)n_rounds[g], first_game_id + (uint32_t)g, ORC_STREAM_GAME, 0, w);
            if (nc == 0 || !(dsum != 0.0)) { /* :183-189 forced pass */
                n_rounds[g] += 1;
                orc_bg_skip_turn(&st[g], orc_die(w[0]), orc_die(w[1]));
                continue;
            }
            const int a = orc_weighted_select(cid, cpi, nc, seed, first_game_id + (uint32_t)g, (uint32_t)n_rounds[g]);
            if (mem_n[g] == mem_cap[g]) { mem_cap[g] = mem_cap[g] ? 2 * mem_cap[g] : 64; mem[g] = (mem_t *)realloc(mem[g], sizeof(mem_t) * (size_t)mem_cap[g]); }
            if (sp_n + nc > sp_cap) { while (sp_n + nc > sp_cap) sp_cap *= 2; sp_ids = (uint16_t *)realloc(sp_ids, sizeof(uint16_t) * (size_t)sp_cap); sp_vals = (float *)realloc(sp_vals, sizeof(float) * (size_t)sp_cap); }
            mem_t *mm = &mem[g][mem_n[g]++]; /* :195-199 */
            mm->st = st[g]; mm->player = st[g].player; mm->pi_off = sp_n; mm->pi_n = nc; mm->ply = n_rounds[g];
            memcpy(sp_ids + sp_n, cid, sizeof(uint16_t) * (size_t)nc);
            memcpy(sp_vals + sp_n, cpi, sizeof(float) * (size_t)nc);
            sp_n += nc;
            const orc_move mv = orc_bg_decode(&st[g], (uint32_t)a); /* :202-210 */
            orc_bg_apply_move(&st[g], mv, orc_die(w[0]), orc_die(w[1]));
            n_rounds[g] += 1;
            const int win = orc_bg_check_winner(&st[g]);
            if (win != ORC_NO_WINNER) { /* :215-223 */
                EMIT(g, 1, win)
                alive[g] = 0;
            }
        }
    }
#undef EMIT
    *n_rec_out = n_rec; *n_pi_out = n_pi;
    if (n_waves_out) *n_waves_out = waves;
    for (int g = 0; g < n_games; ++g) free(mem[g]);
    free(st); free(n_rounds); free(alive); free(mem); free(mem_n); free(mem_cap); free(sp_ids); free(sp_vals);
    free(live); free(ids); free(idx_of); free(nodes); free(nn); free(status);
    return rc;
}
