// bg_device.cuh -- warp-per-game backgammon primitives for sm_100a.
//
// One warp owns one game.  Lane l < 24 holds point l in a register; the eight scalar bytes of
// the packed 32-byte state (bar, off, roll, player, second) are warp-uniform registers.  The
// legal-move list is built in a per-warp shared-memory slab (RAW_CAP u32 sequences + KEPT_CAP
// dedup keys); the board itself never leaves registers during a playout, so a fused playout
// touches HBM once per game on the way in and once on the way out.
//
// What is computed (reference: alibasaran/die-e src/backgammon/backgammon_logic.rs):
//   get_valid_moves :403-414 = candidate generation per die (:555-617 normal + bear-off,
//   :662-682 bar entry), sort by (die, from, to) + dedup (:619-620), recursion with the used
//   die removed (:705-720), DFS root-to-leaf flatten (:722-750), first-wins dedup by resulting
//   board (:753-774).  The reference builds a heap tree; here the same ORDERED list is produced
//   without a tree:
//     * a (die, from) pair has at most one candidate, so "sorted by (die, from, to)" is "by die,
//       then by lane": lane f owns the root candidates that start on point f (lane 24 = bar);
//     * each root lane derives its children from bitmasks of the board after its first
//       sub-move (own>=1, own==1, opponent>=2, opponent==1) and, when bearing off, from the six
//       home-board counts packed in one 64-bit register;
//     * offsets of every root's run of sequences come from one warp prefix scan, so the list
//       is written in exactly the reference's DFS order;
//     * two sequences give the same board iff their canonical (removals, arrivals) multisets
//       are equal (a pass-through point cancels unless it hit a blot) -- a 20-bit exact key,
//       no hashing -- and first-wins dedup is __match_any_sync per 32-chunk plus a scan of the
//       kept keys of earlier chunks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diee.h"

namespace diee {

constexpr int RAW_CAP = 512;   // raw (pre-dedup) sequences per state; observed max 254
constexpr int KEPT_CAP = 256;  // == DIEE_MAX_MOVES
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t M24 = 0x00FFFFFFu;
constexpr uint32_t SEQ_NO_SECOND = 0xFEFE0000u;  // (from2,to2) = (DIEE_NONE, DIEE_NONE)
constexpr uint32_t SEQ_EMPTY = 0xFEFEFEFEu;

struct WarpSlab {  // per-warp shared memory
    uint32_t raw[RAW_CAP];
    uint32_t kept[KEPT_CAP];
};

// ---------------- Philox4x32-10 ----------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                        uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
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
__host__ __device__ __forceinline__ int die_of(uint32_t w) { return 1 + (int)(((uint64_t)w * 6u) >> 32); }
__host__ __device__ __forceinline__ uint32_t index_of(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

// 32 consecutive blocks of one stream, one per lane; refilled every 32 draws.
struct PhiloxLanes {
    uint32_t w0, w1, w2, w3;
    __device__ __forceinline__ void fill(uint64_t seed, uint32_t base, uint32_t c1, uint32_t c2, uint32_t c3, int lane) {
        uint32_t o[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), base + (uint32_t)lane, c1, c2, c3, o);
        w0 = o[0]; w1 = o[1]; w2 = o[2]; w3 = o[3];
    }
};

// ---------------- game state in a warp ----------------
struct BgWarp {
    int v;  // lane < 24: pts[lane]; other lanes 0
    int bar0, bar1, off0, off1, roll0, roll1, player, second;  // warp-uniform
};

__device__ __forceinline__ void bg_load(BgWarp &g, const diee_bg_state *s, int lane) {
    int b = ((const unsigned char *)s)[lane];
    g.v = lane < 24 ? (int)(signed char)b : 0;
    g.bar0 = __shfl_sync(FULL, b, 24);
    g.bar1 = __shfl_sync(FULL, b, 25);
    g.off0 = __shfl_sync(FULL, b, 26);
    g.off1 = __shfl_sync(FULL, b, 27);
    g.roll0 = __shfl_sync(FULL, b, 28);
    g.roll1 = __shfl_sync(FULL, b, 29);
    g.player = (int)(signed char)__shfl_sync(FULL, b, 30);
    g.second = __shfl_sync(FULL, b, 31);
}

__device__ __forceinline__ void bg_store(const BgWarp &g, diee_bg_state *s, int lane) {
    int b = g.v;
    b = lane == 24 ? g.bar0 : b;
    b = lane == 25 ? g.bar1 : b;
    b = lane == 26 ? g.off0 : b;
    b = lane == 27 ? g.off1 : b;
    b = lane == 28 ? g.roll0 : b;
    b = lane == 29 ? g.roll1 : b;
    b = lane == 30 ? g.player : b;
    b = lane == 31 ? g.second : b;
    ((unsigned char *)s)[lane] = (unsigned char)b;
}

// check_winner  backgammon_logic.rs:527-534  (0 = none here)
__device__ __forceinline__ int bg_winner(const BgWarp &g) { return g.off0 == 15 ? -1 : (g.off1 == 15 ? 1 : 0); }

// one sub-move of get_next_state  backgammon_logic.rs:467-517 (arms in source order)
__device__ __forceinline__ void bg_apply_sub(BgWarp &g, int f, int t, int lane) {
    const int p = g.player;
    if (t == -1) {  // collecting
        if (lane == f) g.v -= p;
        if (p < 0) g.off0 += 1; else g.off1 += 1;
        return;
    }
    const int tv = __shfl_sync(FULL, g.v, t);
    const bool hit = tv == -p;
    if (f == -1) {  // from the bar
        if (lane == t) g.v = hit ? p : g.v + p;
        if (p < 0) { g.bar0 -= 1; g.bar1 += hit; } else { g.bar1 -= 1; g.bar0 += hit; }
    } else {
        if (lane == t) g.v = hit ? p : g.v + p;
        if (lane == f) g.v -= p;
        if (hit) { if (p < 0) g.bar1 += 1; else g.bar0 += 1; }
    }
}

__device__ __forceinline__ void bg_apply_board(BgWarp &g, uint32_t seq, int lane) {
    const int f1 = (int)(signed char)(seq), t1 = (int)(signed char)(seq >> 8);
    const int f2 = (int)(signed char)(seq >> 16), t2 = (int)(signed char)(seq >> 24);
    if (f1 != DIEE_NONE) bg_apply_sub(g, f1, t1, lane);
    if (f2 != DIEE_NONE) bg_apply_sub(g, f2, t2, lane);
}

// apply_move :176-186 / skip_turn :192-196 with the next roll injected; seq == SEQ_EMPTY skips
__device__ __forceinline__ void bg_step(BgWarp &g, uint32_t seq, int d0, int d1, int lane) {
    if (seq != SEQ_EMPTY) {
        bg_apply_board(g, seq, lane);
        if (g.roll0 == g.roll1 && !g.second) { g.second = 1; return; }
    }
    g.second = 0;
    g.player = -g.player;
    g.roll0 = d0;
    g.roll1 = d1;
}

// ---------------- candidate generation for one die on one (possibly per-lane) board ----------------
// own1: points with >= 1 own checker; freem: points NOT blocked by >= 2 opposing checkers;
// H: home-board counts (own-relative, +16 bias, byte h = distance-from-off h); bar = own bar count.
// Returns cand (bit f = a candidate from point f, bit 24 = bar entry) and bo (subset that collects).
__device__ __forceinline__ void gen_cands(int p, int m, uint32_t own1, uint32_t freem, int bar, uint64_t H,
                                          uint32_t home, uint32_t &cand, uint32_t &bo) {
    bo = 0;
    if (bar > 0) {  // _get_action_trees :545-548 -> get_entry_moves :668-682
        const int e = p < 0 ? 24 - m : m - 1;
        cand = ((freem >> e) & 1u) ? (1u << 24) : 0u;
        return;
    }
    // :600-617  own checker on pt, target on the board and not blocked
    uint32_t src = (p < 0 ? (freem << m) : (freem >> m)) & own1 & M24;
    if ((own1 & ~home) == 0) {  // is_collectible :638-659 (bar == 0 here)
        // bit h of cm: own checker on home point h AND the signed sum of the higher home points
        // shows no own surplus  (:571-578 for player -1, :588-595 for player +1; quirk Q3)
        uint32_t cm = 0;
        int suf = 0;
#pragma unroll
        for (int h = 5; h >= 0; --h) {
            const int val = (int)((H >> (8 * h)) & 0xFF) - 16;
            if (val >= 1 && suf <= 0) cm |= 1u << h;
            suf += val;
        }
        const uint32_t ownh = p < 0 ? (own1 & 0x3Fu) : ((__brev(own1) >> 8) & 0x3Fu);
        const uint32_t ex = ownh & (1u << (m - 1));              // exact point (:565-568, :584-587)
        const int hmax = p < 0 ? m - 2 : m - 1;                  // -1 scans below pt, +1 scans from pt
        const uint32_t fb = cm & ((1u << (hmax + 1)) - 1u);
        const uint32_t fbbit = fb ? (0x80000000u >> __clz(fb)) : 0u;  // first hit of the downward scan
        const uint32_t boh = ex | fbbit;
        bo = p < 0 ? boh : (__brev(boh) >> 8);
    }
    cand = src | bo;
}

// canonical (removals, arrivals) key of a sequence: equal keys <=> equal resulting boards
__device__ __forceinline__ uint32_t seq_key(uint32_t s, uint32_t oppblot) {
    const int f1 = (int)(signed char)(s), t1 = (int)(signed char)(s >> 8);
    const int f2 = (int)(signed char)(s >> 16), t2 = (int)(signed char)(s >> 24);
    uint32_t r1 = f1 < 0 ? 24u : (uint32_t)f1;
    uint32_t a1 = t1 < 0 ? 25u : (uint32_t)t1;
    uint32_t r2 = 31u, a2 = 31u;
    if (f2 != DIEE_NONE) {
        r2 = f2 < 0 ? 24u : (uint32_t)f2;
        a2 = t2 < 0 ? 25u : (uint32_t)t2;
    }
    const bool hit1 = t1 >= 0 && ((oppblot >> t1) & 1u);
    if (a1 == r2 && a1 < 24u && !hit1) { a1 = 31u; r2 = 31u; }   // same checker moves on, no hit there
    else if (r1 == a2 && r1 < 24u) { r1 = 31u; a2 = 31u; }       // another checker refills the source
    return min(r1, r2) | (max(r1, r2) << 5) | (min(a1, a2) << 10) | (max(a1, a2) << 15);
}

// ---------------- get_valid_moves for the warp's game ----------------
// On return slab.raw[0..U) holds the ordered unique plays (diee_move byte layout) and U is
// returned (warp-uniform).  overflow is set if the raw or unique list did not fit.
__device__ __forceinline__ int bg_movegen(const BgWarp &g, WarpSlab &slab, int lane, bool &overflow) {
    const int p = g.player;
    const int hi = max(g.roll0, g.roll1), lo = min(g.roll0, g.roll1);  // :406-409
    const bool dbl = hi == lo;
    const int pv = g.v * p;  // own-relative count
    const uint32_t own1 = __ballot_sync(FULL, pv >= 1) & M24;
    const uint32_t own_single = __ballot_sync(FULL, pv == 1) & M24;
    const uint32_t oppblk = __ballot_sync(FULL, pv <= -2) & M24;
    const uint32_t oppblot = __ballot_sync(FULL, pv == -1) & M24;
    const uint32_t freem = ~oppblk & M24;
    const int bar_own = p < 0 ? g.bar0 : g.bar1;
    if (own1 == 0 && bar_own == 0) return 0;  // nothing left to move (the side has borne everything off)
    const uint32_t home = p < 0 ? 0x3Fu : 0xFC0000u;
    const uint32_t outside = own1 & ~home;

    // home-board counts, only when a bear-off is reachable within this play
    uint64_t H = 0;
    const bool need_home = bar_own == 0 && __popc(outside) <= 1;
    if (need_home) {
#pragma unroll
        for (int h = 0; h < 6; ++h) {
            const int val = __shfl_sync(FULL, pv, p < 0 ? h : 23 - h);
            H |= (uint64_t)(uint32_t)(val + 16) << (8 * h);
        }
    }

    uint32_t nseq_packed = 0;   // this lane's sequence counts: order 0 | order 1 << 16
    uint32_t C[2] = {0, 0}, BO2[2] = {0, 0};
    int T1[2] = {0, 0};
    bool isroot[2] = {false, false};
    const int norders = dbl ? 1 : 2;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        if (o >= norders) break;
        // sort key is (die, from, to): the LOW die's roots come first (:619)
        const int m1 = o == 0 ? lo : hi, m2 = o == 0 ? hi : lo;
        uint32_t R, BO1;
        gen_cands(p, m1, own1, freem, bar_own, H, home, R, BO1);
        if (lane <= 24 && ((R >> lane) & 1u)) {
            isroot[o] = true;
            const bool frombar = lane == 24;
            const int t1 = frombar ? (p < 0 ? 24 - m1 : m1 - 1) : (((BO1 >> lane) & 1u) ? -1 : lane + p * m1);
            T1[o] = t1;
            // board after the first sub-move, as masks
            uint32_t own1p = own1;
            if (!frombar && ((own_single >> lane) & 1u)) own1p &= ~(1u << lane);
            bool hit1 = false;
            if (t1 >= 0) { own1p |= 1u << t1; hit1 = (oppblot >> t1) & 1u; }
            uint64_t Hp = H;
            if (need_home) {
                if (!frombar && ((home >> lane) & 1u)) Hp -= 1ull << (8 * (p < 0 ? lane : 23 - lane));
                if (t1 >= 0 && ((home >> t1) & 1u)) Hp += (uint64_t)(hit1 ? 2 : 1) << (8 * (p < 0 ? t1 : 23 - t1));
            }
            gen_cands(p, m2, own1p, freem, bar_own - (frombar ? 1 : 0), Hp, home, C[o], BO2[o]);
            const uint32_t n = max(1, __popc(C[o]));
            nseq_packed |= n << (16 * o);
        }
    }

    // exclusive prefix over lanes (ascending `from`; the bar lane is alone when it is a root)
    uint32_t incl = nseq_packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += up;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - nseq_packed;
    const int N0 = total & 0xFFFFu, N = N0 + (int)(total >> 16);
    if (N == 0) return 0;
    if (N > RAW_CAP) overflow = true;

    // emit this lane's runs in DFS order (:722-750)
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        if (o >= norders) break;
        if (isroot[o]) {
            const int m2 = o == 0 ? hi : lo;
            int base = o == 0 ? (int)(excl & 0xFFFFu) : N0 + (int)(excl >> 16);
            const int f1 = lane == 24 ? -1 : lane;
            const uint32_t s1 = (uint32_t)(f1 & 0xFF) | ((uint32_t)(T1[o] & 0xFF) << 8);
            uint32_t c = C[o];
            if (c == 0) {
                if (base < RAW_CAP) slab.raw[base] = s1 | SEQ_NO_SECOND;
            } else {
                while (c) {
                    const int b = __ffs(c) - 1;
                    c &= c - 1;
                    int f2, t2;
                    if (b == 24) { f2 = -1; t2 = p < 0 ? 24 - m2 : m2 - 1; }
                    else { f2 = b; t2 = ((BO2[o] >> b) & 1u) ? -1 : b + p * m2; }
                    if (base < RAW_CAP) slab.raw[base] = s1 | ((uint32_t)(f2 & 0xFF) << 16) | ((uint32_t)(t2 & 0xFF) << 24);
                    ++base;
                }
            }
        }
    }
    __syncwarp();

    // first-wins dedup by resulting board (:753-774), compacting in place
    const int Nc = min(N, RAW_CAP);
    int U = 0;
    for (int c0 = 0; c0 < Nc; c0 += 32) {
        const int idx = c0 + lane;
        const bool valid = idx < Nc;
        const uint32_t s = valid ? slab.raw[idx] : 0u;
        const uint32_t key = valid ? seq_key(s, oppblot) : (0x80000000u | (uint32_t)lane);
        const uint32_t peers = __match_any_sync(FULL, key);
        bool keep = valid && (__ffs(peers) - 1 == lane);
        for (int e = 0; e < U; ++e) keep = keep && (slab.kept[e] != key);  // earlier chunks
        const uint32_t km = __ballot_sync(FULL, keep);
        const int pos = U + __popc(km & ((1u << lane) - 1u));
        __syncwarp();
        if (keep) {
            if (pos < KEPT_CAP) { slab.raw[pos] = s; slab.kept[pos] = key; }
            else overflow = true;
        }
        U += __popc(km);
        __syncwarp();
    }
    return min(U, KEPT_CAP);
}

// ---------------- action-id codec (backgammon_logic.rs:262-401) ----------------
__device__ __forceinline__ int min_roll_of(int f, int t) {  // :277-285, arms in source order
    if (f == -1 && t < 6) return t + 1;
    if (f == -1 && t > 17) return 24 - t;
    if (t == -1 && f < 6) return f + 1;
    if (t == -1 && f > 17) return 24 - f;
    return abs(f - t);
}

__device__ __forceinline__ uint32_t bg_encode_move(int roll0, int roll1, uint32_t seq) {
    if (seq == SEQ_EMPTY) return 1351u;  // :266-268
    const int f[2] = {(int)(signed char)(seq), (int)(signed char)(seq >> 16)};
    const int t[2] = {(int)(signed char)(seq >> 8), (int)(signed char)(seq >> 24)};
    const int n = f[1] == DIEE_NONE ? 1 : 2;
    const int low = min(roll0, roll1);
    int mr[2];
    mr[0] = min_roll_of(f[0], t[0]);
    mr[1] = n > 1 ? min_roll_of(f[1], t[1]) : 0;
    bool low_first = false, low_second = false;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // :299-349
        if (i >= n) break;
        const uint32_t mul = i == 0 ? 1u : 26u;
        bool flag = false, set = false;
        if (f[i] == -1 && t[i] < 6) { sum += mul * 24u; flag = (t[i] + 1) == low; set = true; }
        else if (f[i] == -1 && t[i] > 17) { sum += mul * 24u; flag = (24 - t[i]) == low; set = true; }
        else if (t[i] == -1 && f[i] < 6) { sum += mul * (uint32_t)f[i]; }
        else if (t[i] == -1 && f[i] > 17) { sum += mul * (uint32_t)f[i]; }
        else { sum += mul * (uint32_t)f[i]; flag = mr[i] == low; set = true; }
        if (set) { if (i == 0) low_first = flag; else low_second = flag; }
    }
    if (n == 1) { low_first = false; sum += 26u * 25u; }  // :352
    bool high_first;                                       // :355
    if (low_first) high_first = false;
    else if (low_second) high_first = true;
    else if (mr[1] != 0) high_first = mr[0] >= mr[1];
    else high_first = mr[0] > low;
    return high_first ? sum : sum + 676u;  // :358
}

__device__ __forceinline__ uint32_t bg_decode_move(int roll0, int roll1, int player, uint32_t action) {  // :361-401
    if (action == 1351u) return SEQ_EMPTY;
    const bool high_first = action < 676u;
    const uint32_t x = high_first ? action : action - 676u;
    int from1 = (int)(x % 26u), from2 = (int)(x / 26u);
    const bool single = from2 == 25;
    const int hi = max(roll0, roll1), lo = min(roll0, roll1);
    if (from1 == 24 && player == 1) from1 = -1;
    if (from2 == 24 && player == 1) from2 = -1;
    int to1 = high_first ? from1 + hi * player : from1 + lo * player;
    int to2 = high_first ? from2 + lo * player : from2 + hi * player;
    if (to1 >= 24 || to1 <= -1) to1 = -1;
    if (to2 >= 24 || to2 <= -1) to2 = -1;
    if (from1 == 24) from1 = -1;
    if (from2 == 24) from2 = -1;
    uint32_t s = (uint32_t)(from1 & 0xFF) | ((uint32_t)(to1 & 0xFF) << 8);
    if (single) return s | SEQ_NO_SECOND;
    return s | ((uint32_t)(from2 & 0xFF) << 16) | ((uint32_t)(to2 & 0xFF) << 24);
}

}  // namespace diee
