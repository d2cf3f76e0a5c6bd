// lane_pack.cuh -- what the lane kernels (lane_kernels.cu) and the persistent check-current search (mcts_kernels.cu) share:
// a game's board between its packed 32-byte state and a lane's registers, the code path its next ply needs, and the
// shared-memory layout of the packed kernels (games resident in shared memory, queued by the kind of their next ply).
#pragma once
#include "bg_lane.cuh"

namespace diee {

using namespace lane;

__device__ __forceinline__ void lane_load_state(LaneBoard &g, const diee_bg_state *s) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(s));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(s) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    l_load(g, w);
}
__device__ __forceinline__ void lane_store_state(const LaneBoard &g, diee_bg_state *s) {
    uint32_t w[8];
    l_store(g, w);
    reinterpret_cast<uint4 *>(s)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(s)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// The code path the next ply of a game needs.  A warp runs ONE path per step, for all its lanes that
// wait for that path (see lane_run_kernel).
// PATH_CLOSED: the distinct plays are counted in closed form (contact play, entering from the bar, or
// nothing to move); PATH_WALK: the side can bear off within the play, counted root by root.
// PATH_STORE: the game is over (or the rollout has nothing left to play): write its result, once, in one place.
enum { PATH_DONE = 0, PATH_CLOSED, PATH_WALK, PATH_STORE, PATH_COUNT };

__device__ __forceinline__ int lane_path(const LaneBoard &g) {
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3];
    const uint32_t own1 = g.own[0] | o123;
    if (g.bar_own > 0 || own1 == 0) return PATH_CLOSED;
    const uint32_t outside = own1 & ~0x3Fu;
    if ((outside & (outside - 1u)) == 0 && (outside & ~(g.own[0] & ~o123)) == 0) {
        // bearing off.  Every checker home and no opposing checker there: the play comes out of the table (cheap, so it
        // rides with the closed path); one checker still outside, or contact inside the home board: the walk.
        return (outside == 0 && ((g.opp[0] | g.opp[1] | g.opp[2] | g.opp[3]) & 0x3Fu) == 0) ? PATH_CLOSED : PATH_WALK;
    }
    return PATH_CLOSED;
}

#ifndef DIEE_PK_S
#define DIEE_PK_S 512
#endif
#ifndef DIEE_PK_PATIENCE
#define DIEE_PK_PATIENCE 4
#endif
#ifndef DIEE_PK_T
#define DIEE_PK_T 256
#endif
#ifndef DIEE_PK_MINB
#define DIEE_PK_MINB 3
#endif
constexpr int PK_T = DIEE_PK_T;  // threads per CTA
constexpr int PK_S = DIEE_PK_S;  // resident games per CTA
constexpr int PK_RING = PK_S <= 512 ? 512 : 1024;  // ring size per queue (a power of two >= PK_S: a game is in one queue at most)
constexpr int PK_AREAS = 3;      // warps that may run the bear-off walk at a time (one scratch area each)
constexpr int PK_PATIENCE = DIEE_PK_PATIENCE;  // polls without a full queue before a warp takes a partial one
constexpr int PK_WORDS = 13;
enum { PC_TWO = 0, PC_DBL, PC_BAR, PC_TABLE, PC_WALK, PC_TURN, PC_LISTS, PC_DEAD = PC_LISTS };
constexpr uint32_t PK_HAS_GAME = 1u << 24;
constexpr unsigned PK_EMPTY = 0xFFFFFFFFu;

struct PackSmem {
    uint32_t st[PK_WORDS][PK_S];        // own[4], opp[4], misc, ply, item, game id, counter word 3
    uint32_t scr[PK_AREAS][L_SCRATCH][32];
    unsigned ring[PC_LISTS][PK_RING];     // slot numbers, or PK_EMPTY; every cell is a one-entry mailbox (ring_put / ring_take)
    unsigned head[8], tail[8];
    int area_lock[4];
    int n_dead;
    int n_avail;  // games waiting in the queues (what an idle warp polls)
    int drain;    // the job has no more items: take what there is, at once
};

// A ring cell is a mailbox: a producer puts a slot number into an EMPTY cell (compare-and-swap), a consumer takes whatever
// is there and leaves EMPTY behind (exchange).  Positions are handed out by `tail` (atomic add) and `head` (compare-and-swap
// over what `tail` shows), so a cell's consumer may arrive before its producer has written (it waits), and -- when a
// consumer is slow to read what it claimed while the other warps recycle games through the same ring -- the producer of
// position p + PK_RING may arrive before the consumer of position p has taken its entry (it waits, too).  Entries of one
// ring are interchangeable (slots whose next ply is of that kind), so which of two waiting consumers gets which entry does
// not matter; every entry put is taken exactly once (tests/ring_model.cpp runs the same protocol with host threads).
__device__ __forceinline__ void ring_put(unsigned *cell, unsigned slot) {
    unsigned spins = 0;
    while (atomicCAS(cell, PK_EMPTY, slot) != PK_EMPTY)
        if (++spins > (1u << 28)) __trap();
}
__device__ __forceinline__ unsigned ring_take(unsigned *cell) {
    unsigned v, spins = 0;
    while ((v = atomicExch(cell, PK_EMPTY)) == PK_EMPTY)  // reserved, not written yet (a lost entry must not hang the device)
        if (++spins > (1u << 28)) __trap();
    return v;
}

__device__ __forceinline__ int pack_class(const LaneBoard &g) {
    if (lane_path(g) == PATH_WALK) return PC_WALK;
    if (g.bar_own > 0) return PC_BAR;
    const uint32_t own1 = g.own[0] | g.own[1] | g.own[2] | g.own[3];
    if (own1 != 0 && (own1 & ~0x3Fu) == 0) return PC_TABLE;  // (lane_path: every checker home, no opposing checker there)
    return g.roll0 == g.roll1 ? PC_DBL : PC_TWO;
}
__device__ __forceinline__ uint32_t pack_misc(const LaneBoard &g) {
    return (uint32_t)g.bar_own | ((uint32_t)g.bar_opp << 4) | ((uint32_t)g.off_own << 8) | ((uint32_t)g.off_opp << 12) |
           ((uint32_t)g.roll0 << 16) | ((uint32_t)g.roll1 << 19) | ((uint32_t)g.second << 22) | ((g.player > 0 ? 1u : 0u) << 23) | PK_HAS_GAME;
}
__device__ __forceinline__ void unpack_misc(LaneBoard &g, uint32_t m) {
    g.bar_own = (int)(m & 15u); g.bar_opp = (int)((m >> 4) & 15u); g.off_own = (int)((m >> 8) & 15u); g.off_opp = (int)((m >> 12) & 15u);
    g.roll0 = (int)((m >> 16) & 7u); g.roll1 = (int)((m >> 19) & 7u); g.second = (int)((m >> 22) & 1u); g.player = (m >> 23) & 1u ? 1 : -1;
}

}  // namespace diee
