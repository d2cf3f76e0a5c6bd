// bg_pb_table.h -- the pure bear-off play table (host side).
//
// When every checker of the side to move is in its home board and no opposing checker is (the lone side bearing
// off -- three quarters of all bear-off positions a rollout meets), get_valid_moves depends on very little: which
// of the six home points are empty, hold one checker or hold more, the two dice, and the side (it fixes the order
// `from` ascends in).  That is 64 x 64 x 21 x 2 keys, 30,618 of them reachable, with at most a few dozen plays each,
// so the ordered list of distinct plays is tabulated once per context: the kernels then replace the whole move
// generation of such a ply by two loads.  The table is filled by the lane engine itself (bg_lane.cuh compiled for
// the host -- the code tests/lane_harness.cpp checks against the oracle), so it cannot disagree with the walk.
#pragma once
#include <stdint.h>

#include <vector>

#include "bg_lane.cuh"

namespace diee {

constexpr int PB_KEYS = 2 * 21 * 64 * 64;

// a play, packed: n | x1 << 2 | t1 << 5 | x2 << 8 | t2 << 11   (points 0..5, 7 = collected)
static inline uint16_t pb_pack(const lane::LanePlay &p) {
    auto pt = [](int v) { return (uint16_t)(v == lane::L_OFF ? 7 : v); };
    uint16_t w = (uint16_t)p.n | (uint16_t)(pt(p.x1) << 2) | (uint16_t)(pt(p.t1) << 5);
    if (p.n > 1) w |= (uint16_t)(pt(p.x2) << 8) | (uint16_t)(pt(p.t2) << 11);
    return w;
}

// index[key] = first play << 8 | number of plays;  plays = all lists back to back
static inline void pb_build_table(std::vector<uint32_t> &index, std::vector<uint16_t> &plays) {
    index.assign(PB_KEYS, 0u);
    plays.clear();
    uint32_t scr[lane::L_SCRATCH];
    for (int side = 0; side < 2; ++side)
        for (int hi = 1; hi <= 6; ++hi)
            for (int lo = 1; lo <= hi; ++lo)
                for (uint32_t occ = 1; occ < 64; ++occ)
                    for (uint32_t single = 0; single < 64; ++single) {
                        if (single & ~occ) continue;
                        lane::LaneBoard g;
                        for (int k = 0; k < 4; ++k) g.own[k] = g.opp[k] = 0;
                        g.own[0] = single;          // one checker
                        g.own[1] = occ & ~single;   // two stand for "more than one"
                        g.bar_own = g.bar_opp = g.off_own = g.off_opp = 0;
                        g.roll0 = hi; g.roll1 = lo; g.player = side ? 1 : -1; g.second = 0;
                        lane::LaneGen gen;
                        lane::l_movegen_walk(g, gen, scr, 1);
                        const uint32_t first = (uint32_t)plays.size();
                        for (int k = 0; k < gen.U; ++k) plays.push_back(pb_pack(lane::l_pick_walk(gen, scr, 1, k)));
                        index[lane::l_pb_key(g)] = (first << 8) | (uint32_t)gen.U;
                    }
}

}  // namespace diee
