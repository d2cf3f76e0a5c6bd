// bg_lane.cuh -- lane-per-game backgammon engine for the rollout / playout kernels (sm_100a).
//
// bg_device.cuh gives one WARP to a game (needed where one game must finish fast: the tree
// kernel).  Rollouts are the opposite regime: games x iterations independent plies streams, each
// strictly sequential, so the throughput form is one LANE per game with the whole board in
// registers as bit planes and no cross-lane traffic at all.
//
// What is computed (reference: alibasaran/die-e src/backgammon/backgammon_logic.rs): the same
// get_valid_moves :403-414 as bg_device.cuh -- candidates per die (:555-617, :662-682) sorted by
// (die, from, to) (:619-620), recursion with the used die removed (:705-720), DFS flatten
// (:722-750), first-wins dedup by resulting board (:753-774) -- but a rollout only needs the
// NUMBER of distinct plays U and the k-th of them (node.rs:186-190), so the list is never built:
//
//   * the board is kept mover-relative ("canonical": the side to move runs toward bit 0, its home
//     board is bits 0..5, it enters from the bar on 24 - die) as 4 + 4 bit planes of the per-point
//     checker counts, so own>=1 / own==1 / opponent>=2 / opponent==1 are two or three logic ops and
//     a turn change is a bit reversal of the planes;
//   * roots are visited in the reference's order (die ascending, then `from` ascending in REAL
//     coordinates = descending canonical bits for player +1); each root's children come from one
//     shift-and-mask (plus the bear-off scan when the side can bear off within this play);
//   * two plays give the same board iff their canonical (removals, arrivals) are equal.  A play is
//     (point moved by the low die, point moved by the high die); distinct pairs give distinct
//     boards except (a) the same checker moving on through a point it did not hit / another checker
//     refilling the vacated point -- both collapse to one net move F -> a, deduplicated in four
//     25-bit register masks indexed by F -- and (b) both sub-moves bearing off with the dice swapped.
//     Outside the bear-off regime this makes "is this child new?" a closed-form mask per root;
//     inside it the children are tested one by one against a per-lane bit matrix;
//   * per root the mask of NEW children is parked in per-lane scratch (shared memory, lane-strided,
//     conflict-free); after U is known the k-th play is found by a popcount walk over that scratch.
//
// The file compiles for the host too (tests/lane_harness.cpp checks it exhaustively against the
// CPU oracle without a GPU); on the device everything is inlined into the kernels of lane_kernels.cu.
#pragma once
#include <stdint.h>

#include "../../include/diee.h"

#if defined(__CUDACC__)
#define LANE_HD __host__ __device__ __forceinline__
#else
#define LANE_HD static inline
#endif

namespace diee {
namespace lane {

constexpr uint32_t L_M24 = 0x00FFFFFFu;
constexpr int L_BAR = 24;  // source index of a bar entry
constexpr int L_OFF = 25;  // arrival index of a collected checker
constexpr int L_SLOTS = 32;               // root slots: 2 die orders x (<= 15 occupied points | the bar)
constexpr int L_ROWS = 25;                // pair matrix rows (bear-off regime only)
constexpr int L_SCRATCH = L_SLOTS + L_ROWS;  // u32 words of scratch per lane
constexpr uint32_t L_SINGLE = 0x80000000u;   // new-mask flag: the root itself is a (new) one-move play

LANE_HD int l_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
LANE_HD int l_low(uint32_t x) {  // index of the lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)(x & (0u - x)));  // one find-leading-one on the isolated bit (ffs would be brev + flo)
#else
    return __builtin_ctz(x);
#endif
}
LANE_HD int l_high(uint32_t x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
LANE_HD uint32_t l_rev24(uint32_t x) {  // mirror the 24 points
#if defined(__CUDA_ARCH__)
    return __brev(x) >> 8;
#else
    uint32_t r = 0;
    for (int i = 0; i < 24; ++i) r |= ((x >> i) & 1u) << (23 - i);
    return r;
#endif
}

LANE_HD uint32_t l_rev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
#endif
}

// ---------------- Philox4x32-10 (same stream contract as include/diee.h) ----------------
LANE_HD void l_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
LANE_HD int l_die(uint32_t w) { return 1 + (int)(((uint64_t)w * 6u) >> 32); }
LANE_HD uint32_t l_index(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

// ---------------- the board of one game, in one lane's registers ----------------
struct LaneBoard {
    uint32_t own[4], opp[4];  // bit planes of the per-point counts, mover-relative
    int bar_own, bar_opp, off_own, off_opp;
    int roll0, roll1, player, second;
};

LANE_HD void l_inc(uint32_t p[4], uint32_t bit) {  // count at `bit` += 1 (ripple carry through the planes)
    uint32_t c = bit, t;
    t = p[0] & c; p[0] ^= c; c = t;
    t = p[1] & c; p[1] ^= c; c = t;
    t = p[2] & c; p[2] ^= c; c = t;
    p[3] ^= c;
}
LANE_HD void l_dec(uint32_t p[4], uint32_t bit) {  // count at `bit` -= 1 (ripple borrow)
    uint32_t c = bit, t;
    t = ~p[0] & c; p[0] ^= c; c = t;
    t = ~p[1] & c; p[1] ^= c; c = t;
    t = ~p[2] & c; p[2] ^= c; c = t;
    p[3] ^= c;
}

// packed 32-byte state (real coordinates) -> planes.  w[8] = the state as eight little-endian words.
LANE_HD void l_load(LaneBoard &g, const uint32_t w[8]) {
    uint32_t neg[4] = {0, 0, 0, 0}, pos[4] = {0, 0, 0, 0};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 24; ++i) {
        const int v = (int)(signed char)(w[i >> 2] >> (8 * (i & 3)));
        const uint32_t a = (uint32_t)(v < 0 ? -v : v);
        const uint32_t nm = v < 0 ? 1u : 0u, pm = v > 0 ? 1u : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) {
            neg[k] |= (((a >> k) & 1u) & nm) << i;
            pos[k] |= (((a >> k) & 1u) & pm) << i;
        }
    }
    const int bar0 = (int)(w[6] & 0xFF), bar1 = (int)((w[6] >> 8) & 0xFF);
    const int off0 = (int)((w[6] >> 16) & 0xFF), off1 = (int)((w[6] >> 24) & 0xFF);
    g.roll0 = (int)(w[7] & 0xFF);
    g.roll1 = (int)((w[7] >> 8) & 0xFF);
    g.player = (int)(signed char)(w[7] >> 16);
    g.second = (int)((w[7] >> 24) & 0xFF);
    if (g.player < 0) {
        for (int k = 0; k < 4; ++k) { g.own[k] = neg[k]; g.opp[k] = pos[k]; }
        g.bar_own = bar0; g.bar_opp = bar1; g.off_own = off0; g.off_opp = off1;
    } else {
        for (int k = 0; k < 4; ++k) { g.own[k] = l_rev24(pos[k]); g.opp[k] = l_rev24(neg[k]); }
        g.bar_own = bar1; g.bar_opp = bar0; g.off_own = off1; g.off_opp = off0;
    }
}

LANE_HD void l_store(const LaneBoard &g, uint32_t w[8]) {
    uint32_t neg[4], pos[4];
    int bar0, bar1, off0, off1;
    if (g.player < 0) {
        for (int k = 0; k < 4; ++k) { neg[k] = g.own[k]; pos[k] = g.opp[k]; }
        bar0 = g.bar_own; bar1 = g.bar_opp; off0 = g.off_own; off1 = g.off_opp;
    } else {
        for (int k = 0; k < 4; ++k) { pos[k] = l_rev24(g.own[k]); neg[k] = l_rev24(g.opp[k]); }
        bar1 = g.bar_own; bar0 = g.bar_opp; off1 = g.off_own; off0 = g.off_opp;
    }
    for (int q = 0; q < 6; ++q) w[q] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 24; ++i) {
        int n = 0, p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) {
            n |= (int)((neg[k] >> i) & 1u) << k;
            p |= (int)((pos[k] >> i) & 1u) << k;
        }
        const int v = p - n;
        w[i >> 2] |= (uint32_t)(v & 0xFF) << (8 * (i & 3));
    }
    w[6] = (uint32_t)bar0 | ((uint32_t)bar1 << 8) | ((uint32_t)off0 << 16) | ((uint32_t)off1 << 24);
    w[7] = (uint32_t)g.roll0 | ((uint32_t)g.roll1 << 8) | ((uint32_t)(g.player & 0xFF) << 16) | ((uint32_t)g.second << 24);
}

// check_winner  backgammon_logic.rs:527-534  (0 = none): -1 is asked first
LANE_HD int l_winner(const LaneBoard &g) {
    const int off_m = g.player < 0 ? g.off_own : g.off_opp;  // player -1's collected
    const int off_p = g.player < 0 ? g.off_opp : g.off_own;
    return off_m == 15 ? -1 : (off_p == 15 ? 1 : 0);
}

// one sub-move of get_next_state :467-517 in canonical coordinates (x: 0..23 | L_BAR, t: 0..23 | L_OFF)
LANE_HD void l_apply_sub(LaneBoard &g, int x, int t) {
    if (x == L_BAR) g.bar_own -= 1; else l_dec(g.own, 1u << x);
    if (t == L_OFF) { g.off_own += 1; return; }
    const uint32_t tb = 1u << t;
    if ((g.opp[0] & ~(g.opp[1] | g.opp[2] | g.opp[3])) & tb) {  // a single opposing checker is hit
        g.opp[0] &= ~tb;
        g.bar_opp += 1;
    }
    l_inc(g.own, tb);
}

// the turn passes: mirror the board, swap the sides, take the next roll (:182-184, :192-196)
LANE_HD void l_pass_turn(LaneBoard &g, int d0, int d1) {
    for (int k = 0; k < 4; ++k) {
        const uint32_t o = l_rev24(g.own[k]), p = l_rev24(g.opp[k]);
        g.own[k] = p;
        g.opp[k] = o;
    }
    int t = g.bar_own; g.bar_own = g.bar_opp; g.bar_opp = t;
    t = g.off_own; g.off_own = g.off_opp; g.off_opp = t;
    g.player = -g.player;
    g.second = 0;
    g.roll0 = d0;
    g.roll1 = d1;
}

// a chosen play in canonical coordinates; n = 0 (forced pass), 1 or 2 sub-moves
struct LanePlay {
    int n, x1, t1, x2, t2;
};

// apply_move :176-186 / skip_turn :192-196 with the next roll injected
LANE_HD void l_step(LaneBoard &g, const LanePlay &pl, int d0, int d1) {
    if (pl.n > 0) {
        l_apply_sub(g, pl.x1, pl.t1);
        if (pl.n > 1) l_apply_sub(g, pl.x2, pl.t2);
        if (g.roll0 == g.roll1 && !g.second) { g.second = 1; return; }
    }
    l_pass_turn(g, d0, d1);
}

// canonical play -> diee_move bytes (real coordinates)
LANE_HD uint32_t l_play_to_seq(const LanePlay &pl, int player) {
    if (pl.n == 0) return 0xFEFEFEFEu;
    int f[2] = {DIEE_NONE, DIEE_NONE}, t[2] = {DIEE_NONE, DIEE_NONE};
    const int xs[2] = {pl.x1, pl.x2}, ts[2] = {pl.t1, pl.t2};
    for (int i = 0; i < pl.n; ++i) {
        f[i] = xs[i] == L_BAR ? -1 : (player < 0 ? xs[i] : 23 - xs[i]);
        t[i] = ts[i] == L_OFF ? -1 : (player < 0 ? ts[i] : 23 - ts[i]);
    }
    return (uint32_t)(f[0] & 0xFF) | ((uint32_t)(t[0] & 0xFF) << 8) | ((uint32_t)(f[1] & 0xFF) << 16) | ((uint32_t)(t[1] & 0xFF) << 24);
}

// ---------------- candidate generation for one die (canonical) ----------------
// own1: points with an own checker; freem: points not held by >= 2 opposing checkers; H: the six
// home-board counts (own minus opponent, +16 bias, byte h = point h), only read when every own
// checker is home.  Returns the candidate sources (bit 24 = bar entry).  :544-703
LANE_HD uint32_t l_cands(int m, uint32_t own1, uint32_t freem, int bar, uint64_t H, bool opp_home, bool isplus) {
    if (bar > 0) return ((freem >> (24 - m)) & 1u) << 24;  // :545-548 -> :668-682
    uint32_t cand = (freem << m) & own1 & L_M24;             // :600-617
    if ((own1 & ~0x3Fu) == 0) {                               // is_collectible :638-659 (bar == 0 here)
        // bit h of cm: own checker on home point h AND the signed sum of the higher home points
        // shows no own surplus (:571-578 / :588-595; opposing checkers can cancel own ones, quirk Q3)
        uint32_t cm;
        if (!opp_home) {
            cm = own1 ? (1u << l_high(own1)) : 0u;  // nothing to cancel against: only the highest own point qualifies
        } else {
            cm = 0;
            int suf = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int h = 5; h >= 0; --h) {
                const int val = (int)((H >> (8 * h)) & 0xFF) - 16;
                if (val >= 1 && suf <= 0) cm |= 1u << h;
                suf += val;
            }
        }
        const uint32_t ex = own1 & (1u << (m - 1));          // the exact point (:565-568, :584-587)
        const int hmax = isplus ? m - 1 : m - 2;             // player +1 scans from the exact point, -1 from below it
        const uint32_t fb = cm & ((1u << (hmax + 1)) - 1u);
        const uint32_t fbbit = fb ? (1u << l_high(fb)) : 0u;  // first hit of the downward scan
        cand |= ex | fbbit;
    }
    return cand;
}

LANE_HD int l_to(int x, int m) { return x == L_BAR ? 24 - m : (x >= m ? x - m : L_OFF); }
LANE_HD int l_take(uint32_t mask, bool isplus) { return isplus ? l_high(mask) : l_low(mask); }  // next in `from` order

// ---------------- get_valid_moves, counted ----------------
struct LaneGen {
    uint32_t R0, R1;  // root sources per die order (order 0 = low die first; R1 = 0 for doubles)
    int lo, hi;       // the dice, low and high
    int U;            // number of distinct plays
    int N0;           // closed form, two different dice: distinct plays that start with the low die
    bool isplus;
    bool closed;      // counted in closed form (no scratch was written)
};

LANE_HD bool l_test_and_set(uint32_t &m, uint32_t bit) {  // true iff the bit was clear
    const bool fresh = !(m & bit);
    m |= bit;
    return fresh;
}

// Counts the distinct plays of `g` root by root and parks, per root in reference order, the mask of its
// NEW children in scr[slot * stride].  scr needs L_SCRATCH words (lane-strided).  Exact in every regime;
// used where the closed form below does not apply (checkers on the bar, bearing off).
// BO_ONLY: the caller guarantees the bear-off regime (bar empty, at most one lone checker outside the home board) --
// the lane kernels' walk path, which only ever sees such positions; the other cases then compile away.
template <bool BO_ONLY>
LANE_HD void l_movegen_walk_t(const LaneBoard &g, LaneGen &gen, uint32_t *scr, int stride) {
    const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;  // :406-409
    const bool dbl = hi == lo;
    const bool isplus = g.player > 0;
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3], p123 = g.opp[1] | g.opp[2] | g.opp[3];
    const uint32_t own1 = g.own[0] | o123;
    const uint32_t own_single = g.own[0] & ~o123;
    const uint32_t oppblot = g.opp[0] & ~p123;
    const uint32_t freem = ~p123 & L_M24;
    const int bar = BO_ONLY ? 0 : g.bar_own;
    gen.isplus = isplus;
    gen.closed = false;
    gen.U = 0;
    gen.N0 = 0;
    gen.R0 = gen.R1 = 0;
    gen.lo = lo;
    gen.hi = hi;
    if (own1 == 0 && bar == 0) return;  // everything is borne off

    // the side can bear off within this play: bar empty and at most one checker outside the home board
    const uint32_t outside = own1 & ~0x3Fu;
    const bool bo_regime = BO_ONLY || (bar == 0 && (outside & (outside - 1u)) == 0 && (outside & ~own_single) == 0);
    uint64_t H = 0;
    const bool opp_home = ((g.opp[0] | p123) & 0x3Fu) != 0;  // opposing checkers inside the mover's home board
    if (bo_regime && opp_home) {
        const uint32_t mag[4] = {g.own[0] | g.opp[0], g.own[1] | g.opp[1], g.own[2] | g.opp[2], g.own[3] | g.opp[3]};
        const uint32_t oppany = g.opp[0] | p123;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int h = 0; h < 6; ++h) {
            const int a = (int)(((mag[0] >> h) & 1u) | (((mag[1] >> h) & 1u) << 1) | (((mag[2] >> h) & 1u) << 2) | (((mag[3] >> h) & 1u) << 3));
            const int val = ((oppany >> h) & 1u) ? -a : a;
            H |= (uint64_t)(uint32_t)(val + 16) << (8 * h);
        }
    }

    uint32_t seen_lo = 0, seen_hi = 0, seen_sum = 0, seen_off = 0;  // net single moves F -> a seen so far: a = F-lo, F-hi, F-lo-hi, OFF
    uint32_t rowvalid = 0;             // rows of the pair matrix written during this call
    const uint32_t ownsrc = own1 | (bar > 0 ? 1u << L_BAR : 0u);
    uint32_t *rows = scr + L_SLOTS * stride;
    int U = 0, slot = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 2; ++b) {
        if (b == 1 && dbl) break;
        const int m1 = b == 0 ? lo : hi, m2 = b == 0 ? hi : lo;
        const uint32_t R = l_cands(m1, own1, freem, bar, H, opp_home, isplus);
        if (b == 0) gen.R0 = R; else gen.R1 = R;
        uint32_t r = R;
        while (r) {
            const int x = l_take(r, isplus);
            r &= ~(1u << x);
            const bool frombar = !BO_ONLY && x == L_BAR;
            const int t1 = l_to(x, m1);
            // the board after the first sub-move, as masks
            uint32_t own1p = own1;
            if (!frombar && ((own_single >> x) & 1u)) own1p &= ~(1u << x);
            bool hit1 = false;
            if (t1 < 24) { own1p |= 1u << t1; hit1 = (oppblot >> t1) & 1u; }
            uint64_t Hp = H;
            if (bo_regime && opp_home) {
                if (x < 6) Hp -= 1ull << (8 * x);
                if (t1 < 6) Hp += (uint64_t)(hit1 ? 2 : 1) << (8 * t1);
            }
            const uint32_t C = l_cands(m2, own1p, freem, bar - (frombar ? 1 : 0), Hp, opp_home, isplus);
            uint32_t newm = 0;
            if (C == 0) {  // a one-move play (leaf root): net move x -> t1
                const uint32_t bit = 1u << x;
                const bool fresh = t1 == L_OFF ? l_test_and_set(seen_off, bit)
                                               : (b == 0 ? l_test_and_set(seen_lo, bit) : l_test_and_set(seen_hi, bit));
                if (fresh) newm = L_SINGLE;
            } else if (!bo_regime) {
                // closed form.  The only children that collapse to a net single move x' -> x'-lo-hi:
                const uint32_t runbit = (t1 < 24 && !hit1) ? ((1u << t1) & C) : 0u;         // the same checker moves on
                const int ly = x + m2;
                const uint32_t leapbit = (!frombar && ly < 24) ? ((1u << ly) & C) : 0u;      // a checker refills x
                if (runbit && l_test_and_set(seen_sum, 1u << x)) newm |= runbit;
                if (leapbit && l_test_and_set(seen_sum, 1u << ly)) newm |= leapbit;
                const uint32_t plain = C & ~(runbit | leapbit);
                uint32_t newp;
                if (dbl) {
                    // {x,y} was already produced from root y iff y holds an own checker and comes before x
                    const uint32_t upto = (2u << x) - 1u;  // bits 0..x
                    const uint32_t later = isplus ? upto : ~(upto >> 1);
                    newp = bar > 0 ? plain : (plain & (later | ~own1));
                } else if (b == 0) {
                    newp = plain;
                } else {
                    // (child by low die, root by high die) was produced in order 0 iff the child's source
                    // already held an own checker there (with one checker on the bar order 0 starts elsewhere)
                    newp = bar == 1 ? plain : (plain & ~ownsrc);
                }
                newm |= newp;
            } else {
                // bear-off regime: test every child's canonical net effect
                uint32_t c = C;
                while (c) {
                    const int y = l_take(c, isplus);
                    c &= ~(1u << y);
                    const int t2 = l_to(y, m2);
                    bool fresh;
                    if ((t1 == y && t1 < 24 && !hit1) || (x == t2 && x < 24)) {
                        const int F = (t1 == y && t1 < 24 && !hit1) ? x : y;
                        const int a = (t1 == y && t1 < 24 && !hit1) ? t2 : t1;
                        fresh = a == L_OFF ? l_test_and_set(seen_off, 1u << F) : l_test_and_set(seen_sum, 1u << F);
                    } else {
                        int xl = b == 0 ? x : y, yh = b == 0 ? y : x;  // (moved by the low die, by the high die)
                        if (dbl || (t1 == L_OFF && t2 == L_OFF)) {       // the dice are interchangeable
                            const int mn = xl < yh ? xl : yh, mx = xl < yh ? yh : xl;
                            xl = mn; yh = mx;
                        }
                        uint32_t row = ((rowvalid >> xl) & 1u) ? rows[xl * stride] : 0u;
                        fresh = !((row >> yh) & 1u);
                        row |= 1u << yh;
                        rows[xl * stride] = row;
                        rowvalid |= 1u << xl;
                    }
                    if (fresh) newm |= 1u << y;
                }
            }
            scr[slot * stride] = newm;
            ++slot;
            U += l_popc(newm);
        }
    }
    gen.U = U;
}
LANE_HD void l_movegen_walk(const LaneBoard &g, LaneGen &gen, uint32_t *scr, int stride) { l_movegen_walk_t<false>(g, gen, scr, stride); }

// the k-th distinct play (0 <= k < gen.U) in reference order, from the masks l_movegen_walk parked
LANE_HD LanePlay l_pick_walk(const LaneGen &gen, const uint32_t *scr, int stride, int k) {
    LanePlay pl;
    pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
    int slot = 0, acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 2; ++b) {
        uint32_t r = b == 0 ? gen.R0 : gen.R1;
        const int m1 = b == 0 ? gen.lo : gen.hi, m2 = b == 0 ? gen.hi : gen.lo;
        while (r) {
            const int x = l_take(r, gen.isplus);
            r &= ~(1u << x);
            const uint32_t newm = scr[slot * stride];
            ++slot;
            const int c = l_popc(newm);
            if (k < acc + c) {
                pl.x1 = x;
                pl.t1 = l_to(x, m1);
                if (newm & L_SINGLE) { pl.n = 1; return pl; }
                uint32_t w = gen.isplus ? l_rev32(newm) : newm;
                for (int j = k - acc; j > 0; --j) w &= w - 1u;
                const int pos = l_low(w);
                pl.n = 2;
                pl.x2 = gen.isplus ? 31 - pos : pos;
                pl.t2 = l_to(pl.x2, m2);
                return pl;
            }
            acc += c;
        }
    }
    return pl;
}

// ---------------- closed form: contact play, nothing on the bar, no bearing off ----------------
// Here every candidate is a plain move x -> x-d, so with A(d) = points whose d-target is on the board and
// not blocked, the sources of die d are S(d) = A(d) & own, and the board after x -> t differs from the
// board before only at x (vacated iff it held one checker) and t.  Hence root x's children are
//     S(other die)  minus x if x held one checker  plus t if t can move on and held no own checker,
// and which of them are NEW follows from masks alone (file header: only the "same checker moves on" child
// t and the "another checker refills x" child x+d2 collapse to a net single move F -> F-lo-hi, F = x resp.
// x+d2; every other child that starts from an own point repeats a play of an earlier root -- in the other
// die order, or for doubles the mirrored pair -- and a child that starts from a point the mover did not
// hold (t after hitting a blot there) is always new).  So the count needs no loop over roots at all and
// the k-th play is found by a popcount walk that recomputes one root's mask.
struct LaneMasks {
    uint32_t own1, single, blot;  // own >= 1, own == 1, opponent == 1
    uint32_t A_lo, A_hi;          // points that may move the low / high die
};

LANE_HD bool l_closed_applies(const LaneBoard &g, LaneMasks &m, int lo, int hi) {
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3], p123 = g.opp[1] | g.opp[2] | g.opp[3];
    m.own1 = g.own[0] | o123;
    m.single = g.own[0] & ~o123;
    m.blot = g.opp[0] & ~p123;
    const uint32_t freem = ~p123 & L_M24;
    m.A_lo = (freem << lo) & L_M24;
    m.A_hi = (freem << hi) & L_M24;
    const uint32_t outside = m.own1 & ~0x3Fu;
    const bool bo_regime = (outside & (outside - 1u)) == 0 && (outside & ~m.single) == 0;
    return g.bar_own == 0 && !bo_regime;
}

// two different dice
struct LaneTwo {
    uint32_t R0, R1;       // roots: sources of the low die, of the high die
    uint32_t G0, L, Z0;    // order 0 per root: gains a child (t moves on), loses a child (x itself), has none
    uint32_t D0;           // F whose net move F -> F-lo-hi is produced twice in order 0
    uint32_t NR1, NL1, Z1; // order 1 per root: new run-on child, new refill child (as root masks), no child
    int nR0, nR1;
};
LANE_HD void l_two(const LaneMasks &m, int lo, int hi, bool isplus, LaneTwo &t) {
    t.R0 = m.A_lo & m.own1;
    t.R1 = m.A_hi & m.own1;
    t.nR0 = l_popc(t.R0);
    t.nR1 = l_popc(t.R1);
    t.G0 = t.R0 & ((m.A_hi & ~m.own1) << lo);
    t.L = t.R0 & t.R1 & m.single;
    const uint32_t RUN0 = t.R0 & ((m.A_hi & ~m.blot) << lo);   // F = x: x -> x-lo -> x-lo-hi without a hit on the way
    const uint32_t LEAP0 = m.own1 & (t.R0 << hi) & L_M24;      // F = x+hi refills x
    t.D0 = RUN0 & LEAP0;
    t.Z0 = t.nR1 == 0 ? (t.R0 & ~t.G0) : (t.nR1 == 1 ? (t.L & ~t.G0) : 0u);
    const uint32_t G1 = t.R1 & ((m.A_lo & ~m.own1) << hi);
    const uint32_t RUN1 = t.R1 & ((m.A_lo & ~m.blot) << hi);
    const uint32_t HIT1 = t.R1 & ((m.A_lo & m.blot) << hi);    // runs on through a blot it hit: a new board
    const uint32_t LEAP1 = m.own1 & (t.R1 << lo) & L_M24;
    const uint32_t fresh = ~(RUN0 | LEAP0);
    // F produced by both a run-on (root F) and a refill (root F-lo): the root that comes first keeps it
    const uint32_t newrun = isplus ? (RUN1 & fresh) : (RUN1 & fresh & ~LEAP1);
    const uint32_t newleapF = isplus ? (LEAP1 & fresh & ~RUN1) : (LEAP1 & fresh);
    t.NR1 = HIT1 | newrun;
    t.NL1 = newleapF >> lo;
    t.Z1 = t.nR0 == 0 ? (t.R1 & ~G1) : (t.nR0 == 1 ? (t.L & ~G1) : 0u);
}

// doubles
struct LaneDbl {
    uint32_t R, G, Lx, Z, HIT, NEWRUN, NEWLEAPF;
    int nR;
};
LANE_HD void l_dbl(const LaneMasks &m, int d, bool isplus, LaneDbl &t) {
    t.R = m.A_lo & m.own1;
    t.nR = l_popc(t.R);
    t.G = t.R & ((m.A_lo & ~m.own1) << d);
    t.Lx = t.R & m.single;
    const uint32_t RUN = t.R & ((m.A_lo & ~m.blot) << d);
    t.HIT = t.R & ((m.A_lo & m.blot) << d);
    const uint32_t LEAP = m.own1 & (t.R << d) & L_M24;
    t.NEWRUN = isplus ? RUN : (RUN & ~LEAP);
    t.NEWLEAPF = isplus ? (LEAP & ~RUN) : LEAP;
    t.Z = t.nR == 1 ? (t.Lx & ~t.G) : 0u;
}

// mask of the NEW children of doubles root x (L_SINGLE: the root alone is the play)
LANE_HD uint32_t l_dbl_newmask(const LaneMasks &m, const LaneDbl &t, int d, bool isplus, int x) {
    const uint32_t xb = 1u << x;
    if (t.Z & xb) return L_SINGLE;
    const uint32_t upto = (xb << 1) - 1u;
    // an own-point child repeats an earlier root's pair unless it comes at or after x in root order
    uint32_t nm = t.R & (isplus ? upto : ~(upto >> 1));
    nm &= ~(xb & m.single);
    const uint32_t tb = x >= d ? (xb >> d) : 0u, lb = (xb << d) & L_M24;
    nm &= ~(tb | lb);  // the run-on and the refill child are judged on their own
    if ((t.HIT | t.NEWRUN) & xb) nm |= tb;
    nm |= t.NEWLEAPF & lb;
    return nm;
}

LANE_HD uint32_t l_two_newmask0(const LaneMasks &m, const LaneTwo &t, int lo, int hi, bool isplus, int x) {
    const uint32_t xb = 1u << x;
    if (t.Z0 & xb) return L_SINGLE;
    uint32_t nm = t.R1 & ~(xb & m.single);
    if (t.G0 & xb) nm |= xb >> lo;
    // F -> F-lo-hi produced twice: the later root's copy is the duplicate
    if (isplus) nm &= ~(((xb << hi) & L_M24) & t.D0);
    else if (t.D0 & xb) nm &= ~(xb >> lo);
    return nm;
}
LANE_HD uint32_t l_two_newmask1(const LaneTwo &t, int lo, int hi, int y) {
    const uint32_t yb = 1u << y;
    if (t.Z1 & yb) return L_SINGLE;
    uint32_t nm = 0;
    if (t.NR1 & yb) nm |= yb >> hi;
    if (t.NL1 & yb) nm |= yb << lo;
    return nm;
}

// Counts the distinct plays of `g` in closed form; the caller has checked l_closed_applies.
LANE_HD void l_movegen_closed(const LaneBoard &g, const LaneMasks &m, int lo, int hi, LaneGen &gen) {
    const bool isplus = g.player > 0;
    gen.isplus = isplus;
    gen.closed = true;
    gen.lo = lo;
    gen.hi = hi;
    gen.R1 = 0;
    gen.N0 = 0;
    if (lo == hi) {
        LaneDbl t;
        l_dbl(m, lo, isplus, t);
        gen.R0 = t.R;
        gen.U = t.nR * (t.nR + 1) / 2 - l_popc(t.R & (t.R >> lo)) - l_popc(t.Lx) + l_popc(t.HIT) + l_popc(t.NEWRUN | t.NEWLEAPF) + l_popc(t.Z);
    } else {
        LaneTwo t;
        l_two(m, lo, hi, isplus, t);
        gen.R0 = t.R0;
        gen.R1 = t.R1;
        gen.N0 = t.nR0 * t.nR1 - l_popc(t.L) + l_popc(t.G0) - l_popc(t.D0) + l_popc(t.Z0);
        gen.U = gen.N0 + l_popc(t.NR1) + l_popc(t.NL1) + l_popc(t.Z1);
    }
}

LANE_HD LanePlay l_play_from(int x, int m1, int m2, uint32_t nm, int j, bool isplus) {
    LanePlay pl;
    pl.x1 = x;
    pl.t1 = l_to(x, m1);
    pl.x2 = pl.t2 = 0;
    if (nm & L_SINGLE) { pl.n = 1; return pl; }
    // the j-th child in `from` order: player +1 counts from the top, so count from the bottom of the mirror image
    uint32_t w = isplus ? l_rev32(nm) : nm;
    for (; j > 0; --j) w &= w - 1u;
    const int pos = l_low(w);
    pl.n = 2;
    pl.x2 = isplus ? 31 - pos : pos;
    pl.t2 = l_to(pl.x2, m2);
    return pl;
}

// ---------------- closed form: checkers on the bar ----------------
// Only entries are legal while the bar is occupied (:545-548), so there is one root per die order.
// Two or more on the bar: both sub-moves are entries and the two orders give the same board.  Exactly one:
// after it has entered on e = 24 - die, the other die is played from S(other) plus e itself; every such
// play is new (the other order starts with a different entry) except "enter and move on with the same
// checker", which both orders produce when neither entry point holds a blot.
struct LaneBarMasks {
    uint32_t C0, N1;   // order 0: children of the low-die entry; order 1: NEW children of the high-die entry
    bool fl, fh;       // the low / high die can enter
    bool single1;      // order 1 is a one-move play
    int n0, n1;
};
LANE_HD void l_bar(const LaneBoard &g, const LaneMasks &m, int lo, int hi, LaneBarMasks &t) {
    const uint32_t freem = ~(g.opp[1] | g.opp[2] | g.opp[3]) & L_M24;
    const int el = 24 - lo, eh = 24 - hi;
    t.fl = (freem >> el) & 1u;
    t.fh = (freem >> eh) & 1u;
    t.C0 = t.N1 = 0;
    t.single1 = false;
    if (g.bar_own >= 2) {  // enter twice: (lo, hi), or whichever single entry is possible
        t.n0 = t.fl ? 1 : 0;
        t.n1 = (!t.fl && t.fh) ? 1 : 0;
        if (lo == hi) t.n1 = 0;
        return;
    }
    t.C0 = t.fl ? (m.A_hi & (m.own1 | (1u << el))) : 0u;
    t.n0 = t.fl ? (t.C0 ? l_popc(t.C0) : 1) : 0;
    t.n1 = 0;
    if (lo != hi && t.fh) {
        const uint32_t C1 = m.A_lo & (m.own1 | (1u << eh));
        const bool run0 = t.fl && ((m.A_hi & ~m.blot) >> el & 1u);
        const bool run1 = (m.A_lo & ~m.blot) >> eh & 1u;
        t.N1 = C1 & ~((run0 && run1) ? (1u << eh) : 0u);
        t.single1 = C1 == 0;
        t.n1 = t.single1 ? 1 : l_popc(t.N1);
    }
}
LANE_HD void l_movegen_bar(const LaneBoard &g, const LaneMasks &m, int lo, int hi, LaneGen &gen) {
    LaneBarMasks t;
    l_bar(g, m, lo, hi, t);
    gen.isplus = g.player > 0;
    gen.closed = true;
    gen.lo = lo;
    gen.hi = hi;
    gen.R0 = gen.R1 = 0;
    gen.N0 = t.n0;
    gen.U = t.n0 + t.n1;
}
LANE_HD LanePlay l_pick_bar(const LaneBoard &g, const LaneMasks &m, int lo, int hi, int k) {
    LaneBarMasks t;
    l_bar(g, m, lo, hi, t);
    const bool isplus = g.player > 0;
    LanePlay pl;
    pl.x1 = L_BAR;
    pl.x2 = pl.t2 = 0;
    if (g.bar_own >= 2) {
        const bool both = t.fl && t.fh;
        pl.t1 = t.fl ? 24 - lo : 24 - hi;
        pl.n = both ? 2 : 1;
        if (both) { pl.x2 = L_BAR; pl.t2 = 24 - hi; }
        return pl;
    }
    if (k < t.n0) {
        pl.t1 = 24 - lo;
        return l_play_from(L_BAR, lo, hi, t.C0 ? t.C0 : L_SINGLE, k, isplus);
    }
    return l_play_from(L_BAR, hi, lo, t.single1 ? L_SINGLE : t.N1, k - t.n0, isplus);
}

// ---------------- pure bear-off: every own checker in the home board, no opposing checker there ----------------
// Six points, no hits, no blocks: the sources of die d are the own points >= d-1 (d-1 itself collects, the higher
// ones move down), or, when all checkers sit below d-1, the highest own point (it collects with the bigger die).
// The walk over roots is then twelve predicated steps on 6-bit masks; plays are told apart in registers:
// net single moves F -> a in four 6-bit masks (as everywhere), plays whose two sub-moves are interchangeable
// (doubles; both collect) in a symmetric 6x6 bit matrix, all others in a 6x6 matrix indexed
// (point moved by the low die, point moved by the high die).  Same scratch output as l_movegen_walk.
LANE_HD bool l_pure_bearoff(const LaneBoard &g) {
    const uint32_t own1 = g.own[0] | g.own[1] | g.own[2] | g.own[3];
    const uint32_t oppany = g.opp[0] | g.opp[1] | g.opp[2] | g.opp[3];
    return g.bar_own == 0 && own1 != 0 && (own1 & ~0x3Fu) == 0 && (oppany & 0x3Fu) == 0;
}
LANE_HD uint32_t l_pb_sources(int d, uint32_t own) {
    const uint32_t upper = own & ~((1u << (d - 1)) - 1u);
    return upper ? upper : (own ? (1u << l_high(own)) : 0u);
}
LANE_HD void l_movegen_pb(const LaneBoard &g, LaneGen &gen, uint32_t *scr, int stride) {
    const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
    const bool dbl = hi == lo;
    const bool isplus = g.player > 0;
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3];
    const uint32_t own = (g.own[0] | o123) & 0x3Fu;
    const uint32_t single = g.own[0] & ~o123 & 0x3Fu;
    gen.isplus = isplus;
    gen.closed = false;
    gen.N0 = 0;
    gen.lo = lo;
    gen.hi = hi;
    gen.R0 = gen.R1 = 0;
    uint32_t seen_lo = 0, seen_hi = 0, seen_sum = 0, seen_off = 0;
    uint64_t seen_any = 0, seen_ord = 0;  // bit 6*a + b
    int U = 0, slot = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 2; ++b) {
        if (b == 1 && dbl) break;
        const int m1 = b == 0 ? lo : hi, m2 = b == 0 ? hi : lo;
        const uint32_t R = l_pb_sources(m1, own);
        if (b == 0) gen.R0 = R; else gen.R1 = R;
        uint32_t r = R;
        while (r) {
            const int x = l_take(r, isplus);
            const uint32_t xb = 1u << x;
            r &= ~xb;
            const bool off1 = x < m1;                       // the first sub-move collects
            const uint32_t t1b = off1 ? 0u : (xb >> m1);
            const uint32_t ownp = (own & ~(single & xb)) | t1b;
            const uint32_t C = l_pb_sources(m2, ownp);
            uint32_t newm = 0;
            if (C == 0) {
                const bool fresh = off1 ? l_test_and_set(seen_off, xb) : (b == 0 ? l_test_and_set(seen_lo, xb) : l_test_and_set(seen_hi, xb));
                if (fresh) newm = L_SINGLE;
            } else {
                const uint32_t runbit = t1b & C;                                   // the same checker moves on: F = x
                const uint32_t refbit = (xb << m2) & 0x3Fu & C;                    // a checker refills x: F = x + m2
                if (runbit) {
                    const bool off2 = (x - m1) < m2;
                    if (off2 ? l_test_and_set(seen_off, xb) : l_test_and_set(seen_sum, xb)) newm |= runbit;
                }
                if (refbit) {
                    if (off1 ? l_test_and_set(seen_off, refbit) : l_test_and_set(seen_sum, refbit)) newm |= refbit;
                }
                const uint32_t rest = C & ~(runbit | refbit);
                // interchangeable sub-moves: doubles, or both collect (second collects iff its point < m2)
                const uint32_t anym = dbl ? rest : (off1 ? (rest & ((1u << m2) - 1u)) : 0u);
                const uint32_t ordm = rest & ~anym;
                if (anym) {
                    const uint32_t row = (uint32_t)(seen_any >> (6 * x)) & 0x3Fu;
                    newm |= anym & ~row;
                    seen_any |= (uint64_t)anym << (6 * x);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int y = 0; y < 6; ++y)
                        if ((anym >> y) & 1u) seen_any |= 1ull << (6 * y + x);
                }
                if (ordm) {
                    if (b == 0) {
                        newm |= ordm;
                        seen_ord |= (uint64_t)ordm << (6 * x);
                    } else {  // (child by the low die, this root by the high die): seen iff row[child] has bit x
                        uint32_t dup = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                        for (int y = 0; y < 6; ++y) dup |= (uint32_t)((seen_ord >> (6 * y + x)) & 1ull) << y;
                        newm |= ordm & ~dup;
                    }
                }
            }
            scr[slot * stride] = newm;
            ++slot;
            U += l_popc(newm);
        }
    }
    gen.U = U;
}

// ---------------- pure bear-off by table (bg_pb_table.h builds it, once per context) ----------------
// key = side, dice, occupied home points, home points holding exactly one checker; index[key] = first << 8 | count;
// a play is packed n | x1 << 2 | t1 << 5 | x2 << 8 | t2 << 11 (points 0..5, 7 = collected).
LANE_HD int l_pb_key(const LaneBoard &g) {
    const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3];
    const uint32_t occ = (g.own[0] | o123) & 0x3Fu, single = g.own[0] & ~o123 & 0x3Fu;
    return (((g.player > 0 ? 21 : 0) + hi * (hi - 1) / 2 + (lo - 1)) * 64 + (int)occ) * 64 + (int)single;
}
LANE_HD LanePlay l_pb_unpack(uint32_t w) {
    LanePlay pl;
    pl.n = (int)(w & 3u);
    const int x1 = (int)((w >> 2) & 7u), t1 = (int)((w >> 5) & 7u), x2 = (int)((w >> 8) & 7u), t2 = (int)((w >> 11) & 7u);
    pl.x1 = x1; pl.t1 = t1 == 7 ? L_OFF : t1;
    pl.x2 = x2; pl.t2 = t2 == 7 ? L_OFF : t2;
    return pl;
}

// Counts the distinct plays of `g`.  Contact play is counted in closed form; otherwise root by root.
// (l_movegen / l_pick are the two-pass form -- count, then pick -- of every regime.  The kernels call the fused
// l_contact_select, the bear-off table and l_movegen_walk_t<true> directly; the two-pass form stays as what
// tests/lane_harness.cpp checks them against, position by position, and as the generator of the table.)
LANE_HD void l_movegen(const LaneBoard &g, LaneGen &gen, uint32_t *scr, int stride) {
    const int hi = g.roll0 > g.roll1 ? g.roll0 : g.roll1, lo = g.roll0 > g.roll1 ? g.roll1 : g.roll0;
    LaneMasks m;
    if (l_closed_applies(g, m, lo, hi)) l_movegen_closed(g, m, lo, hi, gen);
    else if (g.bar_own > 0) l_movegen_bar(g, m, lo, hi, gen);
    else if (l_pure_bearoff(g)) l_movegen_pb(g, gen, scr, stride);
    else l_movegen_walk(g, gen, scr, stride);
}

// Root walk of the closed forms.  Root number i (in `from` order) contributes
//     base - (tri ? i : 0) + [in plus0] + [in plus1] - [in minus0] - [in minus1]
// plays (base = children every root has, tri = the triangular correction of doubles, the masks = the roots that
// gain or lose one play), so finding the root that holds play k needs a few bit tests per root; that root's
// mask of new children is built once, afterwards.
struct LaneCum {
    uint32_t R, plus0, plus1, minus0, minus1;
    int base;
    bool tri;
};
// -> point of the root holding play k; j = the play's index within that root
LANE_HD int l_find_root(const LaneCum &c, int k, bool isplus, int &j) {
    uint32_t r = c.R, xb = 0;
    int acc = 0, rank = 0;
    while (r) {
        xb = isplus ? (0x80000000u >> (31 - l_high(r))) : (r & (0u - r));
        r ^= xb;
        const int cnt = c.base - (c.tri ? rank : 0) + ((c.plus0 & xb) ? 1 : 0) + ((c.plus1 & xb) ? 1 : 0) -
                        ((c.minus0 & xb) ? 1 : 0) - ((c.minus1 & xb) ? 1 : 0);
        if (k < acc + cnt) break;
        acc += cnt;
        ++rank;
    }
    j = k - acc;
    return l_high(xb);
}

// the k-th distinct play (0 <= k < gen.U) in reference order
LANE_HD LanePlay l_pick(const LaneBoard &g, const LaneGen &gen, const uint32_t *scr, int stride, int k) {
    if (!gen.closed) return l_pick_walk(gen, scr, stride, k);
    const bool isplus = gen.isplus;
    const int lo = gen.lo, hi = gen.hi;
    LaneMasks m;
    l_closed_applies(g, m, lo, hi);
    if (g.bar_own > 0) return l_pick_bar(g, m, lo, hi, k);
        LaneCum c;
    int j;
    if (lo == hi) {
        LaneDbl t;
        l_dbl(m, lo, isplus, t);
        c.R = t.R; c.base = t.nR; c.tri = true;
        c.plus0 = t.HIT | t.NEWRUN | t.Z;                       // disjoint: a root without children neither runs on nor hits
        c.plus1 = (t.NEWLEAPF >> lo) & t.R;                     // the root whose refill child is new
        c.minus0 = t.Lx;                                        // a lone checker cannot be moved twice
        c.minus1 = t.R & (isplus ? (t.R << lo) : (t.R >> lo));  // the run-on / refill partner is judged on its own
        const int x = l_find_root(c, k, isplus, j);
        return l_play_from(x, lo, lo, l_dbl_newmask(m, t, lo, isplus, x), j, isplus);
    }
    LaneTwo t;
    l_two(m, lo, hi, isplus, t);
    c.tri = false;
    if (k < gen.N0) {
        c.R = t.R0; c.base = t.nR1;
        c.plus0 = t.G0 | t.Z0;                                  // disjoint
        c.plus1 = 0;
        c.minus0 = t.L;
        c.minus1 = isplus ? ((t.D0 >> hi) & t.R0) : t.D0;       // the root holding the later copy of a repeated net move
        const int x = l_find_root(c, k, isplus, j);
        return l_play_from(x, lo, hi, l_two_newmask0(m, t, lo, hi, isplus, x), j, isplus);
    }
    c.R = t.R1; c.base = 0;
    c.plus0 = t.NR1 | t.Z1;                                     // disjoint
    c.plus1 = t.NL1;
    c.minus0 = c.minus1 = 0;
    const int y = l_find_root(c, k - gen.N0, isplus, j);
    return l_play_from(y, hi, lo, l_two_newmask1(t, lo, hi, y), j, isplus);
}

// Count and selection in one go for contact play, bar entries included (the masks are built once): returns U;
// k >= 0 asks for the k-th play, k == -1 for the count only, k == -2 for the play at index_of(w, U) -- a rollout's
// uniform choice --, k <= -3 for the play -k-2 from the end (-3: the last one, which Node::expand pops first).  pl.n == 0 on return when there is no play to report.
LANE_HD int l_contact_select(const LaneBoard &g, const LaneMasks &m, int lo, int hi, int k, uint32_t w, LanePlay &pl) {
    const bool isplus = g.player > 0;
    pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
    // The die orders differ only in WHICH masks drive the root walk and the new-children mask, so each branch just
    // picks its masks; the walk and the child pick then run once, for all lanes of the warp together.
    int x = -1, m1 = lo, m2 = hi, j = 0, U;
    uint32_t nm = 0;
    if (g.bar_own > 0) {
        LaneBarMasks t;
        l_bar(g, m, lo, hi, t);
        U = t.n0 + t.n1;
        if (k == -2 && U > 0) k = (int)l_index(w, (uint32_t)U);
        else if (k < -2) k = k + U + 2 >= 0 ? k + U + 2 : -1;
        if (k >= 0 && k < U) {
            x = L_BAR;
            if (g.bar_own >= 2) {  // two entries (the second one is the bar's own bit), or the one that is possible
                if (t.fl) nm = t.fh ? (1u << L_BAR) : L_SINGLE;
                else { m1 = hi; m2 = lo; nm = L_SINGLE; }
            } else if (k < t.n0) {
                nm = t.C0 ? t.C0 : L_SINGLE; j = k;
            } else {
                m1 = hi; m2 = lo; nm = t.single1 ? L_SINGLE : t.N1; j = k - t.n0;
            }
        }
    } else {
        LaneCum c;
    uint32_t aZ, a1, a2, a3;  // per mode: roots that are plays on their own, and the three masks of the new-children rule
    int mode;
    if (lo == hi) {
        LaneDbl t;
        l_dbl(m, lo, isplus, t);
        U = t.nR * (t.nR + 1) / 2 - l_popc(t.R & (t.R >> lo)) - l_popc(t.Lx) + l_popc(t.HIT) + l_popc(t.NEWRUN | t.NEWLEAPF) + l_popc(t.Z);
        if (k == -2 && U > 0) k = (int)l_index(w, (uint32_t)U);
        else if (k < -2) k = k + U + 2 >= 0 ? k + U + 2 : -1;
        mode = 0;
        c.R = t.R; c.base = t.nR; c.tri = true;
        c.plus0 = t.HIT | t.NEWRUN | t.Z;                       // disjoint: a root without children neither runs on nor hits
        c.plus1 = (t.NEWLEAPF >> lo) & t.R;                     // the root whose refill child is new
        c.minus0 = t.Lx;                                        // a lone checker cannot be moved twice
        c.minus1 = t.R & (isplus ? (t.R << lo) : (t.R >> lo));  // the run-on / refill partner is judged on its own
        aZ = t.Z; a1 = t.R; a2 = t.HIT | t.NEWRUN; a3 = t.NEWLEAPF;
    } else {
        LaneTwo t;
        l_two(m, lo, hi, isplus, t);
        const int N0 = t.nR0 * t.nR1 - l_popc(t.L) + l_popc(t.G0) - l_popc(t.D0) + l_popc(t.Z0);
        U = N0 + l_popc(t.NR1) + l_popc(t.NL1) + l_popc(t.Z1);
        if (k == -2 && U > 0) k = (int)l_index(w, (uint32_t)U);
        else if (k < -2) k = k + U + 2 >= 0 ? k + U + 2 : -1;
        c.tri = false;
        if (k < N0) {
            mode = 1;
            c.R = t.R0; c.base = t.nR1;
            c.plus0 = t.G0 | t.Z0;                              // disjoint
            c.plus1 = 0;
            c.minus0 = t.L;
            c.minus1 = isplus ? ((t.D0 >> hi) & t.R0) : t.D0;   // the root holding the later copy of a repeated net move
            aZ = t.Z0; a1 = t.R1; a2 = t.G0; a3 = t.D0;
        } else {
            mode = 2;
            k -= N0; U -= N0;                                   // (undone below)
            m1 = hi; m2 = lo;
            c.R = t.R1; c.base = 0;
            c.plus0 = t.NR1 | t.Z1;                             // disjoint
            c.plus1 = t.NL1;
            c.minus0 = c.minus1 = 0;
            aZ = t.Z1; a1 = 0; a2 = t.NR1; a3 = t.NL1;
            if (k >= U) k = -1;
            U += N0;
        }
    }
    if (k >= 0 && (mode == 2 || k < U)) {
    x = l_find_root(c, k, isplus, j);
    const uint32_t xb = 1u << x;
    if (aZ & xb) {
        nm = L_SINGLE;
    } else if (mode == 0) {
        // doubles: an own-point child repeats an earlier root's pair unless it comes at or after x in root order; the
        // run-on and the refill child are judged on their own
        const uint32_t upto = (xb << 1) - 1u;
        nm = a1 & (isplus ? upto : ~(upto >> 1));
        nm &= ~(xb & m.single);
        const uint32_t tb = x >= lo ? (xb >> lo) : 0u, lb = (xb << lo) & L_M24;
        nm &= ~(tb | lb);
        if (a2 & xb) nm |= tb;
        nm |= a3 & lb;
    } else if (mode == 1) {
        nm = a1 & ~(xb & m.single);
        if (a2 & xb) nm |= xb >> lo;
        // F -> F-lo-hi produced twice: the later root's copy is the duplicate
        if (isplus) nm &= ~(((xb << hi) & L_M24) & a3);
        else if (a3 & xb) nm &= ~(xb >> lo);
    } else {
        nm = 0;
        if (a2 & xb) nm |= xb >> hi;
        if (a3 & xb) nm |= xb << lo;
    }
    }
    }
    if (x >= 0) pl = l_play_from(x, m1, m2, nm, j, isplus);
    return U;
}

}  // namespace lane
}  // namespace diee
