/*
 * oracle/orc_backgammon.c -- CPU ORACLE (test infrastructure only; see orc.h).
 * Restates src/backgammon/backgammon_logic.rs of alibasaran/die-e.  Every function
 * cites the reference lines it follows.  The structure deliberately mirrors the
 * reference's (candidate list -> sort -> dedup -> recursive tree -> DFS flatten ->
 * first-wins dedup by resulting board) so that move ORDER is reproduced, not just the set.
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>

/* ---- Philox4x32-10 (Salmon et al. 2011), the injected stream ---- */
void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* backgammon_logic.rs:80-94 */
void orc_bg_new(orc_bg_state *s) {
    static const int8_t init[24] = {2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2};
    memset(s, 0, sizeof *s);
    memcpy(s->b.pts, init, 24);
    s->player = -1;
}

/* get_next_state  backgammon_logic.rs:467-517 (no legality checks, arms in source order) */
void orc_bg_next_state(orc_board *b, const int8_t *from, const int8_t *to, int n, int player) {
    for (int i = 0; i < n; ++i) {
        int f = from[i], t = to[i];
        if (t == -1) { /* :470-480 collecting */
            b->pts[f] -= (int8_t)player;
            if (player == -1) b->off[0] += 1; else b->off[1] += 1;
            continue;
        }
        if (f == -1) { /* :482-500 from the bar */
            if (b->pts[t] == -player) {
                b->pts[t] = (int8_t)player;
                if (player == -1) { b->bar[1] += 1; b->bar[0] -= 1; }
                else { b->bar[0] += 1; b->bar[1] -= 1; }
            } else if (player == -1) {
                b->pts[t] -= 1; b->bar[0] -= 1;
            } else {
                b->pts[t] += 1; b->bar[1] -= 1;
            }
        } else if (b->pts[t] == -player) { /* :501-509 hit */
            b->pts[t] = (int8_t)player;
            b->pts[f] -= (int8_t)player;
            if (player == -1) b->bar[1] += 1; else b->bar[0] += 1;
        } else { /* :510-514 */
            b->pts[t] += (int8_t)player;
            b->pts[f] -= (int8_t)player;
        }
    }
}

/* is_collectible  backgammon_logic.rs:638-659 */
int orc_bg_is_collectible(const orc_board *b, int player) {
    if (player == -1) {
        if (b->bar[0] != 0) return 0;
        for (int i = 6; i <= 23; ++i) if (b->pts[i] < 0) return 0;
    } else if (player == 1) {
        if (b->bar[1] != 0) return 0;
        for (int i = 0; i <= 17; ++i) if (b->pts[i] > 0) return 0;
    }
    return 1;
}

/* check_win  :519-525 */
int orc_bg_check_win(const orc_board *b, int player) {
    return player == -1 ? b->off[0] == 15 : b->off[1] == 15;
}

/* check_winner -> check_win_without_player  :106-108, :527-534 */
int orc_bg_check_winner(const orc_bg_state *s) {
    if (s->b.off[0] == 15) return -1;
    if (s->b.off[1] == 15) return 1;
    return ORC_NO_WINNER;
}

typedef struct { int8_t m, from, to; } cand_t;

static int cand_cmp(const void *pa, const void *pb) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->m != b->m) return a->m < b->m ? -1 : 1;
    if (a->from != b->from) return a->from < b->from ? -1 : 1;
    if (a->to != b->to) return a->to < b->to ? -1 : 1;
    return 0;
}

/* sort_unstable + dedup  :619-620, :685-686 */
static int cand_sort_dedup(cand_t *c, int n) {
    qsort(c, (size_t)n, sizeof(cand_t), cand_cmp);
    int w = 0;
    for (int i = 0; i < n; ++i)
        if (w == 0 || cand_cmp(&c[w - 1], &c[i]) != 0) c[w++] = c[i];
    return w;
}

static int trees_rec(const uint8_t *dice, int nd, const orc_board *b, int player,
                     orc_action_node *pool, int cap, int *n_pool, int *first_out);

/* shared tail of get_normal_moves / get_entry_moves: one node per candidate, children by
 * recursion with the used die removed  (:622-634, :688-700, :705-720) */
static int build_nodes(const cand_t *c, int nc, const uint8_t *dice, int nd, const orc_board *b,
                       int player, orc_action_node *pool, int cap, int *n_pool, int *first_out) {
    int first = *n_pool;
    if (first + nc > cap) return -1;
    *n_pool += nc;
    *first_out = first;
    for (int i = 0; i < nc; ++i) {
        pool[first + i].from = c[i].from;
        pool[first + i].to = c[i].to;
        pool[first + i].die = c[i].m;
        pool[first + i].pad = 0;
        pool[first + i].first_child = -1;
        pool[first + i].n_children = 0;
    }
    for (int i = 0; i < nc; ++i) {
        /* _get_children_of_node_action :705-720 */
        orc_board nb = *b;
        orc_bg_next_state(&nb, &c[i].from, &c[i].to, 1, player);
        uint8_t nd2[ORC_MAX_SEQ];
        int k = 0, removed = 0;
        for (int j = 0; j < nd; ++j) {
            if (!removed && dice[j] == (uint8_t)c[i].m) { removed = 1; continue; } /* first occurrence */
            nd2[k++] = dice[j];
        }
        int fc = -1;
        int nch = trees_rec(nd2, k, &nb, player, pool, cap, n_pool, &fc);
        if (nch < 0) return -1;
        pool[first + i].first_child = nch ? fc : -1;
        pool[first + i].n_children = nch;
    }
    return nc;
}

/* get_normal_moves  :555-636 */
static int normal_rec(const uint8_t *dice, int nd, const orc_board *b, int player,
                      orc_action_node *pool, int cap, int *n_pool, int *first_out) {
    cand_t c[4 * 32];
    int nc = 0;
    const int8_t *board = b->pts;
    if (player == -1 && orc_bg_is_collectible(b, player)) { /* :562-580 */
        for (int d = 0; d < nd; ++d) {
            int m = (int8_t)dice[d];
            int point = m - 1;
            if (board[point] < 0) c[nc++] = (cand_t){(int8_t)m, (int8_t)point, -1};
            for (int i = point - 1; i >= 0; --i) {
                int left_sum = 0;
                for (int j = i + 1; j < 6; ++j) left_sum += board[j];
                if (board[i] < 0 && (int8_t)left_sum >= 0) {
                    c[nc++] = (cand_t){(int8_t)m, (int8_t)i, -1};
                    break;
                }
            }
        }
    } else if (player == 1 && orc_bg_is_collectible(b, player)) { /* :581-598 */
        for (int d = 0; d < nd; ++d) {
            int m = (int8_t)dice[d];
            int point = 24 - m;
            if (board[point] > 0) c[nc++] = (cand_t){(int8_t)m, (int8_t)point, -1};
            for (int i = point; i <= 23; ++i) {
                int left_sum = 0;
                for (int j = 18; j < i; ++j) left_sum += board[j];
                if (board[i] > 0 && (int8_t)left_sum <= 0) {
                    c[nc++] = (cand_t){(int8_t)m, (int8_t)i, -1};
                    break;
                }
            }
        }
    }
    for (int d = 0; d < nd; ++d) { /* :600-617 */
        int m = (int8_t)dice[d];
        for (int point = 0; point < 24; ++point) {
            int n = board[point];
            if (player == -1 && n <= player && point - m >= 0 && board[point - m] <= 1)
                c[nc++] = (cand_t){(int8_t)m, (int8_t)point, (int8_t)(point - m)};
            else if (player == 1 && n >= player && point + m <= 23 && board[point + m] >= -1)
                c[nc++] = (cand_t){(int8_t)m, (int8_t)point, (int8_t)(point + m)};
        }
    }
    nc = cand_sort_dedup(c, nc);
    return build_nodes(c, nc, dice, nd, b, player, pool, cap, n_pool, first_out);
}

/* get_entry_moves  :662-703 */
static int entry_rec(const uint8_t *dice, int nd, const orc_board *b, int player,
                     orc_action_node *pool, int cap, int *n_pool, int *first_out) {
    cand_t c[8];
    int nc = 0;
    const int8_t *board = b->pts;
    if (player == -1) {
        for (int d = 0; d < nd; ++d) {
            int m = (int8_t)dice[d];
            int point = 24 - m;
            if (board[point] < 2) c[nc++] = (cand_t){(int8_t)m, -1, (int8_t)point};
        }
    } else if (player == 1) {
        for (int d = 0; d < nd; ++d) {
            int m = (int8_t)dice[d];
            int point = m - 1;
            if (board[point] > -2) c[nc++] = (cand_t){(int8_t)m, -1, (int8_t)point};
        }
    }
    nc = cand_sort_dedup(c, nc);
    return build_nodes(c, nc, dice, nd, b, player, pool, cap, n_pool, first_out);
}

/* _get_action_trees  :544-552 */
static int trees_rec(const uint8_t *dice, int nd, const orc_board *b, int player,
                     orc_action_node *pool, int cap, int *n_pool, int *first_out) {
    int hit = player == -1 ? b->bar[0] : b->bar[1]; /* get_pieces_hit :536-542 */
    if (hit > 0) return entry_rec(dice, nd, b, player, pool, cap, n_pool, first_out);
    return normal_rec(dice, nd, b, player, pool, cap, n_pool, first_out);
}

int orc_bg_normal_moves(const uint8_t *dice, int nd, const orc_board *b, int player,
                        orc_action_node *pool, int cap, int *n_pool) {
    int first = 0; *n_pool = 0;
    return normal_rec(dice, nd, b, player, pool, cap, n_pool, &first);
}
int orc_bg_entry_moves(const uint8_t *dice, int nd, const orc_board *b, int player,
                       orc_action_node *pool, int cap, int *n_pool) {
    int first = 0; *n_pool = 0;
    return entry_rec(dice, nd, b, player, pool, cap, n_pool, &first);
}
int orc_bg_action_trees(const uint8_t *dice, int nd, const orc_board *b, int player,
                        orc_action_node *pool, int cap, int *n_pool) {
    int first = 0; *n_pool = 0;
    return trees_rec(dice, nd, b, player, pool, cap, n_pool, &first);
}

/* extract_sequences_helper  :734-750  (DFS pre-order, root-to-leaf paths) */
static int extract_rec(const orc_action_node *pool, int node, orc_seq cur, orc_seq *out, int cap, int n) {
    if (cur.len >= ORC_MAX_SEQ) return -1;
    cur.from[cur.len] = pool[node].from;
    cur.to[cur.len] = pool[node].to;
    cur.len += 1;
    if (pool[node].n_children == 0) {
        if (n >= cap) return -1;
        out[n++] = cur;
        return n;
    }
    for (int i = 0; i < pool[node].n_children; ++i) {
        n = extract_rec(pool, pool[node].first_child + i, cur, out, cap, n);
        if (n < 0) return -1;
    }
    return n;
}

/* extract_sequences_node  :730-732 */
int orc_bg_extract_sequences_node(const orc_action_node *pool, int node, orc_seq *out, int cap) {
    orc_seq cur;
    memset(&cur, 0, sizeof cur);
    return extract_rec(pool, node, cur, out, cap, 0);
}

/* extract_sequences_list  :722-728 (roots are pool[0..n_roots)) */
int orc_bg_extract_sequences_list(const orc_action_node *pool, int n_roots, orc_seq *out, int cap) {
    int n = 0;
    orc_seq cur;
    memset(&cur, 0, sizeof cur);
    for (int r = 0; r < n_roots; ++r) {
        n = extract_rec(pool, r, cur, out, cap, n);
        if (n < 0) return -1;
    }
    return n;
}

/* remove_duplicate_states  :753-774 (first sequence per distinct resulting board wins) */
int orc_bg_remove_duplicate_states(const orc_board *b, const orc_seq *in, int n, int player, orc_seq *out) {
    orc_board *seen = (orc_board *)malloc(sizeof(orc_board) * (size_t)(n > 0 ? n : 1));
    int ns = 0, w = 0;
    for (int i = 0; i < n; ++i) {
        orc_board cur = *b;
        for (int k = 0; k < in[i].len; ++k) orc_bg_next_state(&cur, &in[i].from[k], &in[i].to[k], 1, player);
        int dup = 0;
        for (int j = 0; j < ns; ++j)
            if (memcmp(&seen[j], &cur, sizeof cur) == 0) { dup = 1; break; }
        if (!dup) { seen[ns++] = cur; out[w++] = in[i]; }
    }
    free(seen);
    return w;
}

/* get_valid_moves  :403-414 */
int orc_bg_valid_moves(const orc_bg_state *s, orc_move *out, int cap) {
    if (s->roll[0] == 0 && s->roll[1] == 0) return -2; /* assert :404 */
    uint8_t dice[2];
    if (s->roll[0] > s->roll[1]) { dice[0] = s->roll[0]; dice[1] = s->roll[1]; } /* :406-409 */
    else { dice[0] = s->roll[1]; dice[1] = s->roll[0]; }
    enum { POOL = 2048 };
    orc_action_node pool[POOL];
    int n_pool = 0;
    int n_roots = orc_bg_action_trees(dice, 2, &s->b, s->player, pool, POOL, &n_pool);
    if (n_roots < 0) return -1;
    orc_seq seqs[ORC_MAX_MOVES], uniq[ORC_MAX_MOVES];
    int n = orc_bg_extract_sequences_list(pool, n_roots, seqs, ORC_MAX_MOVES);
    if (n < 0) return -1;
    int u = orc_bg_remove_duplicate_states(&s->b, seqs, n, s->player, uniq);
    if (u > cap) return -1;
    for (int i = 0; i < u; ++i) {
        out[i].from1 = uniq[i].from[0]; out[i].to1 = uniq[i].to[0];
        if (uniq[i].len > 1) { out[i].from2 = uniq[i].from[1]; out[i].to2 = uniq[i].to[1]; }
        else { out[i].from2 = ORC_NONE; out[i].to2 = ORC_NONE; }
    }
    return u;
}

static int move_len(orc_move m) { return m.from1 == ORC_NONE ? 0 : (m.from2 == ORC_NONE ? 1 : 2); }

/* apply_move  :176-186; roll_die :100-104 with the dice injected */
void orc_bg_apply_move(orc_bg_state *s, orc_move m, uint8_t die0, uint8_t die1) {
    int8_t f[2] = {m.from1, m.from2}, t[2] = {m.to1, m.to2};
    orc_bg_next_state(&s->b, f, t, move_len(m), s->player);
    if (s->roll[0] == s->roll[1] && !s->second) {
        s->second = 1;
    } else {
        s->second = 0;
        s->player = (int8_t)(-s->player);
        s->roll[0] = die0; s->roll[1] = die1;
    }
}

/* skip_turn  :192-196 */
void orc_bg_skip_turn(orc_bg_state *s, uint8_t die0, uint8_t die1) {
    s->second = 0;
    s->player = (int8_t)(-s->player);
    s->roll[0] = die0; s->roll[1] = die1;
}

/* encode  :262-359 (match arms in source order) */
static uint8_t min_roll_of(int f, int t) { /* :277-285 */
    if (f == -1 && t < 6) return (uint8_t)(t + 1);
    if (f == -1 && t > 17) return (uint8_t)(24 - t);
    if (t == -1 && f < 6) return (uint8_t)(f + 1);
    if (t == -1 && f > 17) return (uint8_t)(24 - f);
    int d = f - t;
    return (uint8_t)(d < 0 ? -d : d);
}

uint32_t orc_bg_encode(const orc_bg_state *s, orc_move m) {
    int n = move_len(m);
    if (n == 0) return 1351; /* :266-268 */
    uint8_t low_roll = s->roll[0] > s->roll[1] ? s->roll[1] : s->roll[0]; /* :272 */
    int low_first = 0, low_second = 0;
    uint8_t mr[2];
    int f[2] = {m.from1, m.from2}, t[2] = {m.to1, m.to2};
    mr[0] = min_roll_of(f[0], t[0]);
    mr[1] = n > 1 ? min_roll_of(f[1], t[1]) : 0; /* :288 */
    uint32_t sum = 0;
    for (int i = 0; i < n; ++i) { /* :299-349 */
        uint32_t mul = i == 0 ? 1u : 26u;
        int flag = 0, set = 0;
        if (f[i] == -1 && t[i] < 6) { sum += mul * 24u; flag = (uint8_t)(t[i] + 1) == low_roll; set = 1; }
        else if (f[i] == -1 && t[i] > 17) { sum += mul * 24u; flag = (uint8_t)(24 - t[i]) == low_roll; set = 1; }
        else if (t[i] == -1 && f[i] < 6) { sum += mul * (uint32_t)f[i]; }
        else if (t[i] == -1 && f[i] > 17) { sum += mul * (uint32_t)f[i]; }
        else { sum += mul * (uint32_t)f[i]; flag = mr[i] == low_roll; set = 1; }
        if (set) { if (i == 0) low_first = flag; else low_second = flag; }
    }
    if (n == 1) { low_first = 0; sum += 26u * 25u; } /* :352 */
    int high_first; /* :355 */
    if (low_first) high_first = 0;
    else if (low_second) high_first = 1;
    else if (mr[1] != 0) high_first = mr[0] >= mr[1];
    else high_first = mr[0] > low_roll;
    return high_first ? sum : sum + 676u; /* :358 */
}

/* decode  :361-401 */
orc_move orc_bg_decode(const orc_bg_state *s, uint32_t action) {
    orc_move r = {ORC_NONE, ORC_NONE, ORC_NONE, ORC_NONE};
    if (action == 1351) return r;
    int player = s->player;
    int high_first = action < 676;
    uint32_t x = high_first ? action : action - 676;
    int from1 = (int)(x % 26), from2 = (int)(x / 26);
    int single = from2 == 25;
    int hi = s->roll[0] > s->roll[1] ? s->roll[0] : s->roll[1];
    int lo = s->roll[0] > s->roll[1] ? s->roll[1] : s->roll[0];
    if (from1 == 24 && player == 1) from1 = -1; /* :384-385 */
    if (from2 == 24 && player == 1) from2 = -1;
    int to1, to2;
    if (high_first) { to1 = from1 + hi * player; to2 = from2 + lo * player; }
    else { to1 = from1 + lo * player; to2 = from2 + hi * player; }
    if (to1 >= 24 || to1 <= -1) to1 = -1; /* :395-398 */
    if (to2 >= 24 || to2 <= -1) to2 = -1;
    if (from1 == 24) from1 = -1;
    if (from2 == 24) from2 = -1;
    r.from1 = (int8_t)from1; r.to1 = (int8_t)to1;
    if (!single) { r.from2 = (int8_t)from2; r.to2 = (int8_t)to2; }
    return r;
}

/* as_tensor  :198-252 -> [1,6,4,6] f32, point p at (h=p/6, w=p%6), raw integers */
int orc_bg_as_tensor(const orc_bg_state *s, float *o) {
    if (s->roll[0] == 0 && s->roll[1] == 0) return -2; /* assert :199 */
    for (int p = 0; p < 24; ++p) {
        int top = p < 12;
        o[0 * 24 + p] = (float)s->b.pts[p];
        o[1 * 24 + p] = (float)s->player;
        o[2 * 24 + p] = (float)(top ? s->b.bar[0] : s->b.bar[1]);
        o[3 * 24 + p] = (float)(top ? s->b.off[0] : s->b.off[1]);
        o[4 * 24 + p] = (float)(top ? s->roll[0] : s->roll[1]);
        o[5 * 24 + p] = s->second ? 1.0f : 0.0f;
    }
    return 0;
}

/* one ply of the random policy used by Agent::Random (versus.rs:307-316) and by
 * Node::simulate's loop body (node.rs:186-193): choose uniformly, apply, else skip */
int orc_bg_random_ply(orc_bg_state *s, const uint32_t w[4]) {
    orc_move mv[ORC_MAX_MOVES];
    int n = orc_bg_valid_moves(s, mv, ORC_MAX_MOVES);
    if (n < 0) return n;
    uint8_t d0 = orc_die(w[0]), d1 = orc_die(w[1]);
    if (n > 0) orc_bg_apply_move(s, mv[orc_index(w[2], (uint32_t)n)], d0, d1);
    else orc_bg_skip_turn(s, d0, d1);
    return n;
}

/* C2: random-vs-random playout from s (already rolled).  Stops at a winner or after
 * round_limit plies (versus.rs:231-235 semantics: winner test after each applied ply). */
int orc_bg_playout(orc_bg_state *s, uint64_t seed, uint32_t game_id, int round_limit, int32_t *plies) {
    int p = 0;
    int w = orc_bg_check_winner(s);
    while (w == ORC_NO_WINNER && p < round_limit) {
        uint32_t blk[4];
        orc_philox(seed, (uint32_t)p, game_id, ORC_STREAM_GAME, 0, blk);
        orc_bg_random_ply(s, blk);
        ++p;
        w = orc_bg_check_winner(s);
    }
    *plies = p;
    return w == ORC_NO_WINNER ? 0 : w;
}
