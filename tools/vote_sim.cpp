// tools/vote_sim.cpp -- DESIGN AID, not part of the product or of the tests.
//
// A host-side model of how lane_run_kernel<true> schedules the rollouts of one search, to compare scheduling
// designs without a GPU.  The rollouts themselves are the real thing (the product's lane engine, compiled for the
// host, on the Philox rollout stream): every rollout becomes the sequence of code paths its plies need.  The warp
// model then replays the kernel's vote (most waiting lanes wins between the closed path and the bear-off walk,
// results stored in batches, up to `reps` plies per vote while a lane stays on the voted path) and charges every
// executed ply-step a cost: a shared part plus one part per sub-case present among the lanes that take the step
// (the warp runs them one after the other).  Costs are warp instructions read off profiles/ (v8).
//
//   g++ -O2 -std=c++17 -o /tmp/vote_sim tools/vote_sim.cpp && /tmp/vote_sim [games] [iterations]
//
// Output: steps per path, lanes per step and modelled instructions per played ply for
//   A  the kernel as it is (one game per lane),
//   B  the bear-off walk folded into the closed path at a given cost (what a cheaper generator would buy),
//   C  two games per lane (a lane takes a step with whichever of its games waits for the voted path),
//   D  one vote path per sub-case (measured on the GPU: 1.9x slower -- the model's sanity check),
//   F  a CTA that re-packs its games by sub-case every ply (full, nearly homogeneous warps at the price of an exchange).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../die_e_b200/csrc/bg_lane.cuh"
#include "../die_e_b200/csrc/bg_pb_table.h"

using namespace diee::lane;

enum Sub : uint8_t { S_TWO = 0, S_DBL, S_BAR, S_PB, S_PASS, S_WALK, S_WALK1, S_COUNT };  // S_WALK: every checker home; S_WALK1: one outside
static const char *SUB_NAME[S_COUNT] = {"two dice", "doubles", "bar", "bear-off table", "pass", "bear-off walk (all home)", "bear-off walk (one outside)"};
// warp instructions per executed sub-case and per step (profiles/r01_lane_run_rollouts_v8_ncu_full_summary.txt, rounded)
static int COST[S_COUNT] = {230, 210, 100, 40, 20, 1000, 1000};  // (the walk: 22 % of the instructions at ~14 % of the ply steps)
static int COST_SHARED = 270;  // Philox, apply, turn change, next path, loop
static int COST_VOTE = 150, COST_STORE = 300;  // (vote: calibrated on the measured plies-per-vote sweep, 8 best)

static Sub classify(const LaneBoard &g) {
    const uint32_t o123 = g.own[1] | g.own[2] | g.own[3];
    const uint32_t own1 = g.own[0] | o123;
    if (g.bar_own > 0) return S_BAR;
    if (own1 == 0) return S_PASS;
    const uint32_t outside = own1 & ~0x3Fu;
    if ((outside & (outside - 1u)) == 0 && (outside & ~(g.own[0] & ~o123)) == 0)
        return outside != 0 ? S_WALK1 : (((g.opp[0] | g.opp[1] | g.opp[2] | g.opp[3]) & 0x3Fu) == 0 ? S_PB : S_WALK);
    return g.roll0 == g.roll1 ? S_DBL : S_TWO;
}

// the opening advanced `adv` plies by random play (bench.py's synthetic inputs)
static void start_state(LaneBoard &g, int game, int adv) {
    static const int8_t OPEN[24] = {2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2};
    uint8_t st[32];
    memset(st, 0, 32);
    memcpy(st, OPEN, 24);
    uint32_t o[4];
    l_philox(0xD1EEu, 0u, 0u, (uint32_t)game, DIEE_STREAM_INIT, 0u, o);
    st[28] = (uint8_t)l_die(o[0]); st[29] = (uint8_t)l_die(o[1]); st[30] = (uint8_t)-1; st[31] = 0;
    uint32_t w[8];
    memcpy(w, st, 32);
    l_load(g, w);
    uint32_t scr[L_SCRATCH];
    for (int q = 0; q < adv && l_winner(g) == 0; ++q) {
        l_philox(0xD1EEu, 0u, (uint32_t)q, (uint32_t)game, DIEE_STREAM_GAME, 0u, o);
        LaneGen gen;
        l_movegen(g, gen, scr, 1);
        LanePlay pl;
        pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
        if (gen.U > 0) pl = l_pick(g, gen, scr, 1, (int)l_index(o[2], (uint32_t)gen.U));
        l_step(g, pl, l_die(o[0]), l_die(o[1]));
    }
}

struct Result {
    long long steps[3] = {0, 0, 0}, lanes[3] = {0, 0, 0}, plies = 0, exec = 0, exec_lanes = 0, cost = 0;
    long long chain = 0;  // the longest dependent chain of any warp (or CTA, design F): what bounds a job that fits one wave
};

// D: one vote path per sub-case (what DESIGN.md section 4 reports as measured: 1.19 -> 2.21 ms)
static Result simulate_fine(const std::vector<std::vector<uint8_t>> &seqs, int reps, int store_min) {
    Result r;
    const int n = (int)seqs.size();
    for (int base = 0; base < n; base += 32) {
        int item[32], pos[32], live = 0;
        for (int l = 0; l < 32; ++l) { item[l] = base + l < n ? base + l : -1; pos[l] = 0; if (item[l] >= 0) ++live; }
        auto need = [&](int l) -> int { return item[l] < 0 ? -1 : (pos[l] >= (int)seqs[item[l]].size() ? (int)S_COUNT : (int)seqs[item[l]][pos[l]]); };
        while (live > 0) {
            r.cost += COST_VOTE + 10 * S_COUNT;
            int cnt[S_COUNT + 1] = {0};
            for (int l = 0; l < 32; ++l) { const int nd = need(l); if (nd >= 0) ++cnt[nd]; }
            int best = 0;
            for (int c = 1; c < S_COUNT; ++c) if (cnt[c] > cnt[best]) best = c;
            if (cnt[best] == 0 || cnt[S_COUNT] >= store_min) best = S_COUNT;
            const int slot = best == S_COUNT ? 2 : ((best == S_WALK || best == S_WALK1) ? 1 : 0);
            ++r.steps[slot];
            r.lanes[slot] += cnt[best];
            if (best == S_COUNT) {
                r.cost += COST_STORE;
                for (int l = 0; l < 32; ++l) if (need(l) == S_COUNT) { item[l] = -1; --live; }
                continue;
            }
            bool in[32];
            for (int l = 0; l < 32; ++l) in[l] = need(l) == best;
            for (int rep = 0; rep < reps; ++rep) {
                int active = 0;
                for (int l = 0; l < 32; ++l) {
                    if (!in[l]) continue;
                    if (need(l) != best) { in[l] = false; continue; }
                    ++pos[l];
                    ++active;
                }
                if (active == 0) break;
                ++r.exec;
                r.exec_lanes += active;
                r.plies += active;
                r.cost += COST_SHARED + COST[best];
            }
        }
    }
    return r;
}

// seqs[i] = sub-case of every played ply of rollout i.  games_per_lane rollouts share a lane.
// `waves` > 1: the job is `waves` times what fits the machine and lanes are refilled from the queue (warps are
// simulated one after the other, each drawing from its own share of the queue -- the order in which real warps draw
// does not matter for the averages).
// fold_walk: 0 = both kinds of walk are the voted walk path, 1 = the all-home kind runs on the closed path at walk_cost
// (the one-outside kind stays voted), 2 = both on the closed path
static Result simulate(const std::vector<std::vector<uint8_t>> &seqs, int games_per_lane, int fold_walk, int walk_cost, int reps,
                       int store_min, int waves = 1, double walk_weight = 1.0) {
    Result r;
    const int per_warp = 32 * games_per_lane;
    const int n = (int)seqs.size();
    const int share = per_warp * waves;  // items one warp works through
    for (int base = 0; base < n; base += share) {
        struct Slot { int item, pos; };
        std::vector<Slot> slot(per_warp);
        int live = 0;
        int next = base + per_warp;
        const int end = std::min(n, base + share);
        for (int i = 0; i < per_warp; ++i) { slot[i] = {base + i < end ? base + i : -1, 0}; if (slot[i].item >= 0) ++live; }
        const long long cost_at_start = r.cost;
        auto need = [&](const Slot &s) -> int {  // 0 closed, 1 walk, 2 store, -1 nothing
            if (s.item < 0) return -1;
            const auto &q = seqs[s.item];
            if (s.pos >= (int)q.size()) return 2;
            return (q[s.pos] == S_WALK1 && fold_walk != 2) || (q[s.pos] == S_WALK && !fold_walk) ? 1 : 0;
        };
        while (live > 0) {
            r.cost += COST_VOTE;
            // per lane: does any of its games wait for path p?
            int cnt[3] = {0, 0, 0};
            for (int l = 0; l < 32; ++l) {
                bool has[3] = {false, false, false};
                for (int k = 0; k < games_per_lane; ++k) { const int nd = need(slot[l * games_per_lane + k]); if (nd >= 0) has[nd] = true; }
                for (int p = 0; p < 3; ++p) cnt[p] += has[p];
            }
            int best = (double)cnt[0] >= walk_weight * cnt[1] ? 0 : 1;  // walk_weight < 1: the walk must gather more lanes to win
            if (cnt[best] == 0) best = 2;
            if (cnt[2] >= store_min && cnt[2] > 0) best = 2;
            ++r.steps[best];
            r.lanes[best] += cnt[best];
            if (best == 2) {
                r.cost += COST_STORE;
                for (auto &s : slot)
                    if (need(s) == 2) {
                        if (next < end) { s.item = next++; s.pos = 0; }  // refill from the queue
                        else { s.item = -1; --live; }
                    }
                continue;
            }
            // the lanes that take the step, each with ONE of its games, for up to `reps` plies
            std::vector<int> cur(32, -1);
            for (int l = 0; l < 32; ++l)
                for (int k = 0; k < games_per_lane; ++k) if (need(slot[l * games_per_lane + k]) == best) { cur[l] = l * games_per_lane + k; break; }
            for (int rep = 0; rep < reps; ++rep) {
                bool present[S_COUNT] = {false};
                int active = 0;
                for (int l = 0; l < 32; ++l) {
                    if (cur[l] < 0) continue;
                    Slot &s = slot[cur[l]];
                    if (need(s) != best) { cur[l] = -1; continue; }
                    present[seqs[s.item][s.pos]] = true;
                    ++s.pos;
                    ++active;
                }
                if (active == 0) break;
                ++r.exec;
                r.exec_lanes += active;
                r.plies += active;
                r.cost += COST_SHARED;
                for (int c = 0; c < S_COUNT; ++c)
                    if (present[c]) r.cost += ((c == S_WALK && fold_walk) || (c == S_WALK1 && fold_walk == 2)) ? walk_cost : COST[c];
            }
        }
        r.chain = std::max(r.chain, r.cost - cost_at_start);
    }
    return r;
}

// F: a CTA of W warps re-packs its 32*W games every ply: sorted by sub-case, consecutive 32-game chunks go to the warps,
// so a warp is full and sees one or two sub-cases.  Cost per warp and ply: shared + its chunk's sub-cases + `exchange`
// (state through shared memory, counting sort, two block barriers).  Finished games are replaced from the queue
// (`waves` shares) at COST_STORE per 32.
static Result simulate_sorted(const std::vector<std::vector<uint8_t>> &seqs, int W, int exchange, int waves) {
    Result r;
    const int per_cta = 32 * W, n = (int)seqs.size(), share = per_cta * waves;
    for (int base = 0; base < n; base += share) {
        std::vector<int> item(per_cta), pos(per_cta, 0);
        int next = base + per_cta;
        const int end = std::min(n, base + share);
        int live = 0;
        for (int i = 0; i < per_cta; ++i) { item[i] = base + i < end ? base + i : -1; if (item[i] >= 0) ++live; }
        long long stored = 0, chain = 0;
        while (live > 0) {
            // finished games leave (and are replaced)
            for (int i = 0; i < per_cta; ++i)
                if (item[i] >= 0 && pos[i] >= (int)seqs[item[i]].size()) {
                    ++stored;
                    if (next < end) { item[i] = next++; pos[i] = 0; }
                    else { item[i] = -1; --live; }
                }
            if (live == 0) break;
            // sort the live games by sub-case
            std::vector<std::pair<int, int>> order;  // (sub-case, slot)
            for (int i = 0; i < per_cta; ++i)
                if (item[i] >= 0 && pos[i] < (int)seqs[item[i]].size()) order.push_back({seqs[item[i]][pos[i]], i});
            if (order.empty()) continue;
            std::sort(order.begin(), order.end());
            long long slowest = 0;
            for (size_t c0 = 0; c0 < order.size(); c0 += 32) {
                bool present[S_COUNT] = {false};
                const size_t c1 = std::min(order.size(), c0 + 32);
                for (size_t j = c0; j < c1; ++j) { present[order[j].first] = true; ++pos[order[j].second]; }
                ++r.exec;
                r.exec_lanes += (long long)(c1 - c0);
                r.plies += (long long)(c1 - c0);
                long long wc = COST_SHARED + exchange;
                for (int c = 0; c < S_COUNT; ++c) if (present[c]) wc += COST[c];
                r.cost += wc;
                slowest = std::max(slowest, wc);
            }
            chain += slowest;  // the warps of a CTA meet at a barrier every ply
        }
        r.cost += stored * COST_STORE / 32;
        r.chain = std::max(r.chain, chain);
    }
    return r;
}

static void report(const char *name, const Result &r) {
    printf("%-58s steps closed/walk/store %8lld %8lld %7lld  lanes/step %5.1f %5.1f %5.1f | ply steps %8lld at %5.1f lanes | %6.1f instr per played ply | longest chain %7lld\n",
           name, r.steps[0], r.steps[1], r.steps[2], r.steps[0] ? (double)r.lanes[0] / r.steps[0] : 0.0,
           r.steps[1] ? (double)r.lanes[1] / r.steps[1] : 0.0, r.steps[2] ? (double)r.lanes[2] / r.steps[2] : 0.0, r.exec,
           r.exec ? (double)r.exec_lanes / r.exec : 0.0, r.plies ? (double)r.cost / r.plies : 0.0, r.chain);
}

int main(int argc, char **argv) {
    const int games = argc > 1 ? atoi(argv[1]) : 128, iterations = argc > 2 ? atoi(argv[2]) : 100;
    std::vector<uint32_t> pb_index;
    std::vector<uint16_t> pb_plays;
    diee::pb_build_table(pb_index, pb_plays);
    std::vector<std::vector<uint8_t>> seqs;
    long long hist[S_COUNT] = {0};
    uint32_t scr[L_SCRATCH];
    for (int gm = 0; gm < games; ++gm) {
        LaneBoard root;
        start_state(root, gm, 10 * (gm % 9));
        for (int it = 0; it < iterations; ++it) {
            LaneBoard g = root;
            std::vector<uint8_t> q;
            for (uint32_t k = 0; k < 400 && !(g.off_own == 15 && g.off_opp == 15); ++k) {
                uint32_t o[4];
                l_philox(0xD1EEu, 0u, k, (uint32_t)gm, DIEE_STREAM_ROLLOUT, (uint32_t)it, o);
                const Sub sub = classify(g);
                q.push_back((uint8_t)sub);
                ++hist[sub];
                LaneGen gen;
                l_movegen(g, gen, scr, 1);
                LanePlay pl;
                pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
                if (gen.U > 0) pl = l_pick(g, gen, scr, 1, (int)l_index(o[2], (uint32_t)gen.U));
                l_step(g, pl, l_die(o[0]), l_die(o[1]));
            }
            seqs.push_back(std::move(q));
        }
    }
    long long total = 0;
    for (int c = 0; c < S_COUNT; ++c) total += hist[c];
    printf("%d games x %d rollouts, %lld played plies (%.1f per rollout):", games, iterations, total, (double)total / seqs.size());
    for (int c = 0; c < S_COUNT; ++c) printf("  %s %.1f%%", SUB_NAME[c], 100.0 * hist[c] / total);
    printf("\n");
    report("A  as built (one game per lane, walk voted separately)", simulate(seqs, 1, 0, COST[S_WALK], 8, 8));
    report("B  walk folded into the closed path, at its own cost", simulate(seqs, 1, 2, COST[S_WALK], 8, 8));
    report("B' ... at a third of its cost", simulate(seqs, 1, 2, COST[S_WALK] / 3, 8, 8));
    report("B\" ... at table cost", simulate(seqs, 1, 2, COST[S_PB], 8, 8));
    report("E  only the all-home walk folded, at table cost (GPU: -4 %)", simulate(seqs, 1, 1, COST[S_PB], 8, 8));
    report("C  two games per lane", simulate(seqs, 2, 0, COST[S_WALK], 8, 16));
    report("C' four games per lane", simulate(seqs, 4, 0, COST[S_WALK], 8, 32));
    report("D  one vote path per sub-case", simulate_fine(seqs, 8, 8));
    for (int reps : {1, 4, 8, 16, 64}) {
        char name[64];
        snprintf(name, sizeof name, "A  with %d plies per vote (GPU: 8 is best)", reps);
        report(name, simulate(seqs, 1, 0, COST[S_WALK], reps, 8));
    }
    for (double ww : {0.25, 0.5, 0.75, 1.5, 3.0}) {
        char name[64];
        snprintf(name, sizeof name, "A  walk lanes weighted %.2f in the vote", ww);
        report(name, simulate(seqs, 1, 0, COST[S_WALK], 8, 8, 1, ww));
    }
    for (int fill : {16, 8}) {  // thinner warps: only `fill` of the 32 lanes get a rollout (shorter chains, more warps)
        std::vector<std::vector<uint8_t>> thin;
        for (size_t i = 0; i < seqs.size(); i += (size_t)fill) {
            for (int l = 0; l < 32; ++l) thin.push_back(l < fill && i + l < seqs.size() ? seqs[i + l] : std::vector<uint8_t>());
        }
        char name[80];
        snprintf(name, sizeof name, "G  %d rollouts per warp (plies counted incl. empty lanes: x%d warps)", fill, 32 / fill);
        report(name, simulate(thin, 1, 0, COST[S_WALK], 8, 8));
    }
    for (int W : {2, 4, 8}) {
        char name[80];
        snprintf(name, sizeof name, "F  CTA of %d warps re-packed by sub-case every ply (+150)", W);
        report(name, simulate_sorted(seqs, W, 150, 1));
    }
    report("F  ... 4 warps, exchange at 300", simulate_sorted(seqs, 4, 300, 1));
    printf("-- a job six times the machine, lanes refilled from the queue (store batch 24) --\n");
    report("A  as built", simulate(seqs, 1, 0, COST[S_WALK], 8, 24, 6));
    report("B' walk folded into the closed path at a third of its cost", simulate(seqs, 1, 2, COST[S_WALK] / 3, 8, 24, 6));
    report("B\" ... at table cost", simulate(seqs, 1, 2, COST[S_PB], 8, 24, 6));
    report("E  only the all-home walk folded, at table cost (GPU: -11 %)", simulate(seqs, 1, 1, COST[S_PB], 8, 24, 6));
    for (double ww : {0.25, 0.5, 0.75, 1.5}) {
        char name[64];
        snprintf(name, sizeof name, "A  walk lanes weighted %.2f in the vote", ww);
        report(name, simulate(seqs, 1, 0, COST[S_WALK], 8, 24, 6, ww));
    }
    for (int W : {2, 4, 8}) {
        char name[80];
        snprintf(name, sizeof name, "F  CTA of %d warps re-packed by sub-case every ply (+150)", W);
        report(name, simulate_sorted(seqs, W, 150, 6));
    }
    report("F  ... 4 warps, exchange at 300", simulate_sorted(seqs, 4, 300, 6));
    report("C  two games per lane", simulate(seqs, 2, 0, COST[S_WALK], 8, 24, 3));
    report("C' four games per lane", simulate(seqs, 4, 0, COST[S_WALK], 8, 24, 2));
    return 0;
}
