// tests/lane_harness.cpp -- TEST ONLY.  Compiles the product's lane-per-game engine
// (die_e_b200/csrc/bg_lane.cuh) for the host and checks it against the CPU oracle (oracle/liborc.so):
//   * for every position: U == len(get_valid_moves) and play k == the oracle's k-th move, all k;
//   * l_step == apply_move / skip_turn; l_load / l_store round trip;
//   * whole Philox-driven playouts == orc_bg_playout (winner, plies, final state).
// Positions: (a) every ply of random playouts from the opening, (b) synthetic random boards that
// reach corners natural play rarely visits (opposing checkers inside the home board while bearing
// off, many checkers on the bar, crowded points).
//
// usage: lane_harness <n_playouts> <n_synthetic> <seed>      exit 0 = all equal
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../die_e_b200/csrc/bg_lane.cuh"
#include "../die_e_b200/csrc/bg_pb_table.h"
#include "../oracle/orc.h"

using namespace diee::lane;

static uint64_t rng_state;
static uint32_t rnd() {  // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
static int rint_(int n) { return (int)(rnd() % (uint32_t)n); }

static void print_state(const orc_bg_state &s) {
    fprintf(stderr, "pts:");
    for (int i = 0; i < 24; ++i) fprintf(stderr, " %d", s.b.pts[i]);
    fprintf(stderr, "  bar %d %d off %d %d roll %d %d player %d second %d\n", s.b.bar[0], s.b.bar[1], s.b.off[0], s.b.off[1],
            s.roll[0], s.roll[1], s.player, s.second);
}

static long long n_checked = 0, n_moves_checked = 0, n_boregime = 0, n_bo_opp_home = 0, n_doubles = 0, n_bar = 0;
static int max_moves = 0;
static long long n_closed = 0, n_pb = 0;
static long long n_bo_outside = 0, n_bo_outside3 = 0;
static std::vector<uint32_t> pb_index;
static std::vector<uint16_t> pb_plays;

static bool check_position(const orc_bg_state &s) {
    uint32_t w[8];
    memcpy(w, &s, 32);
    LaneBoard g;
    l_load(g, w);
    uint32_t w2[8];
    l_store(g, w2);
    if (memcmp(w, w2, 32) != 0) {
        fprintf(stderr, "load/store round trip differs\n");
        print_state(s);
        return false;
    }
    static orc_move mv[ORC_MAX_MOVES];
    const int n = orc_bg_valid_moves(&s, mv, ORC_MAX_MOVES);
    uint32_t scr[L_SCRATCH];
    LaneGen gen;
    l_movegen(g, gen, scr, 1);
    ++n_checked;
    {   // coverage bookkeeping: positions in the bear-off regime, and those with opposing checkers in the home board
        const uint32_t o123 = g.own[1] | g.own[2] | g.own[3];
        const uint32_t own1 = g.own[0] | o123, outside = own1 & ~0x3Fu;
        const bool bo = g.bar_own == 0 && (outside & (outside - 1u)) == 0 && (outside & ~(g.own[0] & ~o123)) == 0 && own1 != 0;
        if (bo) {   // the walk as the lane kernels instantiate it (bear-off regime taken for granted)
            uint32_t scr2[L_SCRATCH];
            LaneGen gen2;
            l_movegen_walk_t<true>(g, gen2, scr2, 1);
            if (gen2.U != n) { fprintf(stderr, "bear-off walk count differs: %d vs %d\n", gen2.U, n); print_state(s); return false; }
            for (int k = 0; k < n; ++k) {
                const LanePlay pl = l_pick_walk(gen2, scr2, 1, k);
                uint32_t o;
                memcpy(&o, &mv[k], 4);
                if (l_play_to_seq(pl, g.player) != o) { fprintf(stderr, "bear-off walk play %d differs\n", k); print_state(s); return false; }
            }
        }
        if (bo && n > 0 && outside) {   // one lone checker outside the home board: how many outside sources a play can have
            ++n_bo_outside;
            const int hi2 = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo2 = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
            int p = 0;
            while (!((outside >> p) & 1u)) ++p;
            if (p - hi2 >= 6 && lo2 != hi2) ++n_bo_outside3;
        }
        if (bo && n > 0) {
            ++n_boregime;
            if ((g.opp[0] | g.opp[1] | g.opp[2] | g.opp[3]) & 0x3Fu) ++n_bo_opp_home;
        }
        if (g.roll0 == g.roll1 && n > 0) ++n_doubles;
        if (g.bar_own > 0 && n > 0) ++n_bar;
        if (n > max_moves) max_moves = n;
        if (gen.closed && n > 0) ++n_closed;
        {   // the fused count + select the kernels call for contact play
            const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
            LaneMasks mm;
            if (g.bar_own > 0) {
                l_closed_applies(g, mm, lo, hi);
                LanePlay pl;
                if (l_contact_select(g, mm, lo, hi, -1, 0u, pl) != n) { fprintf(stderr, "bar_select count differs\n"); print_state(s); return false; }
                for (int k = 0; k < n; ++k) {
                    l_contact_select(g, mm, lo, hi, k, 0u, pl);
                    uint32_t o;
                    memcpy(&o, &mv[k], 4);
                    if (l_play_to_seq(pl, g.player) != o) { fprintf(stderr, "bar_select play %d differs\n", k); print_state(s); return false; }
                }
            } else if (l_closed_applies(g, mm, lo, hi)) {
                LanePlay pl;
                if (l_contact_select(g, mm, lo, hi, -1, 0u, pl) != n) { fprintf(stderr, "closed_select count differs\n"); print_state(s); return false; }
                for (int k = 0; k < n; ++k) {
                    l_contact_select(g, mm, lo, hi, k, 0u, pl);
                    uint32_t o;
                    memcpy(&o, &mv[k], 4);
                    if (l_play_to_seq(pl, g.player) != o) { fprintf(stderr, "closed_select play %d differs\n", k); print_state(s); return false; }
                }
            }
        }
        {   // ... and asked from the end (k = -3 - j: play n-1-j, none when j >= n), as the tree kernel does when it counts a node
            const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
            LaneMasks mm;
            const bool closed = l_closed_applies(g, mm, lo, hi);
            if (g.bar_own > 0 || closed) {
                for (int j = 0; j < 5; ++j) {
                    LanePlay pl;
                    if (l_contact_select(g, mm, lo, hi, -3 - j, 0u, pl) != n) { fprintf(stderr, "select from the end: count differs\n"); print_state(s); return false; }
                    if (j >= n) {
                        if (pl.n != 0) { fprintf(stderr, "select from the end: a play past the first one\n"); print_state(s); return false; }
                        continue;
                    }
                    uint32_t o;
                    memcpy(&o, &mv[n - 1 - j], 4);
                    if (l_play_to_seq(pl, g.player) != o) { fprintf(stderr, "select from the end: play %d differs\n", j); print_state(s); return false; }
                }
            }
        }
        if (l_pure_bearoff(g) && n > 0) ++n_pb;
        if (l_pure_bearoff(g)) {  // the play table the kernels use for these positions
            const uint32_t e = pb_index[l_pb_key(g)];
            if ((int)(e & 255u) != n) { fprintf(stderr, "table count differs: %u oracle %d\n", e & 255u, n); print_state(s); return false; }
            for (int k = 0; k < n; ++k) {
                const uint32_t q = l_play_to_seq(l_pb_unpack(pb_plays[(e >> 8) + k]), g.player);
                uint32_t o;
                memcpy(&o, &mv[k], 4);
                if (q != o) { fprintf(stderr, "table play %d differs\n", k); print_state(s); return false; }
            }
        }
    }
    if (gen.U != n) {
        fprintf(stderr, "count differs: lane %d oracle %d\n", gen.U, n);
        print_state(s);
        for (int k = 0; k < n; ++k) fprintf(stderr, "  oracle %d: (%d,%d) (%d,%d)\n", k, mv[k].from1, mv[k].to1, mv[k].from2, mv[k].to2);
        for (int k = 0; k < gen.U; ++k) {
            const uint32_t q = l_play_to_seq(l_pick(g, gen, scr, 1, k), g.player);
            fprintf(stderr, "  lane   %d: (%d,%d) (%d,%d)\n", k, (int8_t)q, (int8_t)(q >> 8), (int8_t)(q >> 16), (int8_t)(q >> 24));
        }
        return false;
    }
    for (int k = 0; k < n; ++k) {
        const LanePlay pl = l_pick(g, gen, scr, 1, k);
        const uint32_t q = l_play_to_seq(pl, g.player);
        uint32_t o;
        memcpy(&o, &mv[k], 4);
        if (q != o) {
            fprintf(stderr, "move %d of %d differs: lane (%d,%d) (%d,%d) oracle (%d,%d) (%d,%d)\n", k, n, (int8_t)q, (int8_t)(q >> 8),
                    (int8_t)(q >> 16), (int8_t)(q >> 24), mv[k].from1, mv[k].to1, mv[k].from2, mv[k].to2);
            print_state(s);
            for (int j = 0; j < n; ++j) fprintf(stderr, "  oracle %d: (%d,%d) (%d,%d)\n", j, mv[j].from1, mv[j].to1, mv[j].from2, mv[j].to2);
            return false;
        }
        ++n_moves_checked;
        // the successor state
        LaneBoard h = g;
        const int d0 = 1 + rint_(6), d1 = 1 + rint_(6);
        l_step(h, pl, d0, d1);
        orc_bg_state t = s;
        orc_bg_apply_move(&t, mv[k], (uint8_t)d0, (uint8_t)d1);
        uint32_t wa[8], wb[8];
        l_store(h, wa);
        memcpy(wb, &t, 32);
        if (memcmp(wa, wb, 32) != 0) {
            fprintf(stderr, "successor differs after move %d\n", k);
            print_state(s);
            print_state(t);
            orc_bg_state u;
            memcpy(&u, wa, 32);
            print_state(u);
            return false;
        }
    }
    if (n == 0) {
        LaneBoard h = g;
        LanePlay pl;
        pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
        l_step(h, pl, 3, 4);
        orc_bg_state t = s;
        orc_bg_skip_turn(&t, 3, 4);
        uint32_t wa[8];
        l_store(h, wa);
        if (memcmp(wa, &t, 32) != 0) {
            fprintf(stderr, "skip_turn differs\n");
            print_state(s);
            return false;
        }
    }
    return true;
}

// a random board: checkers of both sides thrown on random points (never sharing a point), the rest on
// the bar or collected.  `home_bias` crowds the movers into their home boards to reach bear-off corners.
static void synthetic(orc_bg_state &s, int home_bias) {
    memset(&s, 0, sizeof s);
    s.player = rint_(2) ? 1 : -1;
    for (int side = 0; side < 2; ++side) {
        const int sign = side == 0 ? -1 : 1;
        int left = 15;
        const int npts = 1 + rint_(home_bias ? 5 : 9);
        for (int q = 0; q < npts && left > 0; ++q) {
            int pt;
            if (home_bias && rint_(8) != 0) {
                const int h = rint_(home_bias == 2 && rint_(3) == 0 ? 9 : 6);
                pt = side == 0 ? h : 23 - h;
            } else {
                pt = rint_(24);
            }
            if (s.b.pts[pt] * sign < 0) continue;  // held by the other side
            int c = 1 + rint_(rint_(4) == 0 ? 5 : 2);
            if (c > left) c = left;
            if (s.b.pts[pt] * sign + c > 15) continue;
            s.b.pts[pt] = (int8_t)(s.b.pts[pt] + sign * c);
            left -= c;
        }
        int bar = 0;
        if (rint_(home_bias ? 12 : 4) == 0) bar = 1 + rint_(3);
        if (bar > left) bar = left;
        s.b.bar[side] = (uint8_t)bar;
        s.b.off[side] = (uint8_t)(left - bar);
    }
    s.roll[0] = (uint8_t)(1 + rint_(6));
    s.roll[1] = (uint8_t)(1 + rint_(6));
    s.second = (uint8_t)(s.roll[0] == s.roll[1] ? rint_(2) : 0);
}

int main(int argc, char **argv) {
    const int n_playouts = argc > 1 ? atoi(argv[1]) : 200;
    const long long n_syn = argc > 2 ? atoll(argv[2]) : 200000;
    rng_state = argc > 3 ? strtoull(argv[3], nullptr, 0) : 1;
    diee::pb_build_table(pb_index, pb_plays);

    // (a) every ply of oracle-driven playouts, plus the lane engine playing the same games on its own
    for (int gm = 0; gm < n_playouts; ++gm) {
        orc_bg_state s;
        orc_bg_new(&s);
        uint32_t blk[4];
        orc_philox(7, 0, (uint32_t)gm, ORC_STREAM_INIT, 0, blk);
        s.roll[0] = orc_die(blk[0]);
        s.roll[1] = orc_die(blk[1]);
        const orc_bg_state start = s;
        int p = 0;
        while (orc_bg_check_winner(&s) == ORC_NO_WINNER && p < 400) {
            if (!check_position(s)) return 1;
            orc_philox(7, (uint32_t)p, (uint32_t)gm, ORC_STREAM_GAME, 0, blk);
            orc_bg_random_ply(&s, blk);
            ++p;
        }
        // the same game played by the lane engine alone
        uint32_t w[8];
        memcpy(w, &start, 32);
        LaneBoard g;
        l_load(g, w);
        int q = 0;
        uint32_t scr[L_SCRATCH];
        while (l_winner(g) == 0 && q < 400) {
            uint32_t o[4];
            l_philox(7u, 0u, (uint32_t)q, (uint32_t)gm, DIEE_STREAM_GAME, 0u, o);
            LaneGen gen;
            l_movegen(g, gen, scr, 1);
            LanePlay pl;
            pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
            if (gen.U > 0) pl = l_pick(g, gen, scr, 1, (int)l_index(o[2], (uint32_t)gen.U));
            l_step(g, pl, l_die(o[0]), l_die(o[1]));
            ++q;
        }
        l_store(g, w);
        if (q != p || memcmp(w, &s, 32) != 0) {
            fprintf(stderr, "playout %d differs: lane plies %d oracle %d\n", gm, q, p);
            return 1;
        }
        // reference-exact rollouts keep stepping after the game is over: make sure that agrees too
        for (int extra = 0; extra < 40; ++extra) {
            if (!check_position(s)) return 1;
            orc_philox(9, (uint32_t)extra, (uint32_t)gm, ORC_STREAM_ROLLOUT, 0, blk);
            orc_bg_random_ply(&s, blk);
        }
    }
    // (b) synthetic boards
    for (long long i = 0; i < n_syn; ++i) {
        orc_bg_state s;
        synthetic(s, (int)(i % 3));
        if (!check_position(s)) return 1;
    }
    printf("lane engine == oracle on %lld positions, %lld plays (bear-off regime %lld, of which opposing checkers in the home board %lld; "
           "a lone checker outside %lld, with three outside sources %lld; doubles %lld; from the bar %lld; counted in closed form %lld; pure bear-off form %lld; most plays in one position %d)\n",
           n_checked, n_moves_checked, n_boregime, n_bo_opp_home, n_bo_outside, n_bo_outside3, n_doubles, n_bar, n_closed, n_pb, max_moves);
    return 0;
}
