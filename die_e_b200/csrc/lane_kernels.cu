// lane_kernels.cu -- one LANE per game: the throughput kernels of the path.
//
//   lane_run_kernel<LANE_PLAYOUT>   C2: whole random-vs-random games
//   lane_run_kernel<LANE_ROLLOUT>   T4: every deferred Node::simulate of a split search (node.rs:176-196),
//                                   one lane per (game, iteration); reference-exact rollouts (quirk Q5)
//   lane_run_kernel<LANE_ROLLOUT_CC> T4 with DIEE_MODE_ROLLOUT_CHECK_CURRENT: the rollouts of ONE iteration of a lock-step
//                                   search, one lane per game; the rollout stops at a winner and its result goes back
//                                   to the tree kernel of the next iteration
//
// A ply is get_valid_moves -> uniform choice -> apply_move | skip_turn.  The board lives in the lane's
// registers as bit planes (bg_lane.cuh).  Contact plies touch no memory at all (closed forms); pure bear-off
// plies read two words of the 1 MB play table (L2 / L1 resident); the remaining bear-off plies use the lane's
// own column of a shared-memory scratch (new-children masks per root), laid out [word][lane] so that a warp's
// accesses never conflict.  HBM: 32 B in and 32 B (+5 B) out per game / per rollout, nothing in between.
// Randomness: one Philox4x32-10 block per ply, counter = (ply, game id, stream, epoch<<16|iteration)
// -- the contract of include/diee.h, identical to the warp-per-game kernels and to the oracle.
#include <climits>
#include <cstdlib>

#include "bg_lane.cuh"
#include "lane_pack.cuh"
#include "launchers.h"

namespace diee {

using namespace lane;

constexpr int LANE_CTA = 64;

#ifdef DIEE_LANE_STATS
__device__ unsigned long long g_lane_stats[16];
__device__ unsigned long long g_lane_stats2[16];  // [p] = steps executed on path p, [8+p] = lanes advanced
#endif

enum { LANE_PLAYOUT = 0, LANE_ROLLOUT = 1, LANE_ROLLOUT_CC = 2 };

struct LaneJob {
    // PLAYOUT (C2): item = game index; ROLLOUT (T4): item = game * iterations + iteration
    long long n_items;
    uint32_t iterations;  // ROLLOUT only
    uint32_t it_begin, it_count;  // ROLLOUT: the slice of iterations this launch covers (items = games x it_count)
    uint32_t n_games;             // ROLLOUT: games of the job
    int game_minor;               // ROLLOUT: item = iteration * n_games + game (consecutive lanes = different games) instead of game-major
    uint32_t limit;
    uint64_t seed;
    uint32_t first_game_id, epoch;
    int reps;                      // scheduling: plies a lane may play on one vote while it stays on the same path
    int store_min;                 // scheduling: lanes that must wait to write a result / take an item before it happens
    int lag_weight;                // scheduling: how much one step of waiting counts against one more waiting lane
    unsigned long long *next_item; // job queue head (zeroed before the launch); next_item[1] counts the plies PLAYED
    PbTable pb;                    // pure bear-off play table
    const diee_bg_state *states;   // PLAYOUT: starts[n]; ROLLOUT: node pool states
    const int32_t *sim_node;       // ROLLOUT: node each simulation rolls out from, or -1
    diee_bg_state *finals;         // PLAYOUT: nullable
    int8_t *winners;               // PLAYOUT
    int32_t *plies;                // PLAYOUT
    // ROLLOUT_CC (lock-step search): item = game (first_item + item), one iteration per launch
    uint32_t first_item;           // first game of this launch's group
    const int8_t *players;         // the player each game's search counts the value for
    float *results;                // [game]: Node::simulate's return value (node.rs:182-185,195)
    diee_search_stats *stats;      // [game]: rollout_plies += plies played (nullable)
    int pack_slots;                // lane_pack_kernel: resident games per CTA (<= PK_S; fewer when the job is smaller than the machine)
};

// One lane per item (a game, or one rollout of a search), whole job in one launch.
//
// Lanes of a warp are independent games, so nothing forces them to play ply k together, or even to
// work on the same item for the same time.  The kernel is persistent: a lane that has finished its item
// takes the next one from the job queue, so short games do not leave lanes empty while a long one ends.
// Each lane knows which code path its next ply needs (pass / two dice in contact / doubles in contact /
// entering from the bar / bearing off).  Every step the warp votes: the path with the most waiting lanes
// wins, lanes that have waited long counting extra so that nobody starves; the warp executes THAT path
// once, for the lanes waiting on it, and they move on to their next ply.  Divergence between code paths
// is thereby turned into batching.
#ifndef DIEE_LANE_MIN_BLOCKS
#define DIEE_LANE_MIN_BLOCKS 1
#endif
template <int MODE>
__global__ void __launch_bounds__(LANE_CTA)
lane_run_kernel(LaneJob job) {
    constexpr bool ROLLOUT = MODE != LANE_PLAYOUT;   // plays on the ROLLOUT stream from a node of the pool
    constexpr bool CC = MODE == LANE_ROLLOUT_CC;     // stops at a winner (of the rolled-out state) and reports the result
    __shared__ uint32_t scratch[L_SCRATCH][LANE_CTA];
    uint32_t *scr = &scratch[0][threadIdx.x];
    const int lane = threadIdx.x & 31;
    const uint32_t stream = ROLLOUT ? DIEE_STREAM_ROLLOUT : DIEE_STREAM_GAME;
    LaneBoard g;
    long long item = -1;
    uint32_t gid = 0, c3 = 0, k = 0;
    int need = PATH_DONE;
    uint32_t age = 0;  // steps this lane has waited for its path
    bool queue_open = true;

    for (;;) {
        // ---- idle lanes take the next items of the job (one atomic per warp) ----
        // (in batches, like the stores: store_min lanes, or everybody)
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, need == PATH_DONE);
        if (queue_open && (__popc(idle) >= job.store_min || idle == 0xFFFFFFFFu)) {
            long long first = 0;
            if (lane == 0) first = (long long)atomicAdd(job.next_item, (unsigned long long)__popc(idle));
            first = __shfl_sync(0xFFFFFFFFu, first, 0);
            if (first + __popc(idle) >= job.n_items) queue_open = false;
            if (need == PATH_DONE) {
                item = first + __popc(idle & ((1u << lane) - 1u));
                if (item >= job.n_items) item = -1;
                if (item >= 0) {
                    k = 0;
                    if (CC) {
                        const uint32_t gm = job.first_item + (uint32_t)item;
                        item = (long long)gm * job.iterations + job.it_begin;
                        const int node = job.sim_node[item];
                        if (node >= 0) {  // node < 0: nothing was deferred for this game in this iteration
                            lane_load_state(g, job.states + (size_t)gm * (job.iterations + 1) + node);
                            gid = job.first_game_id + gm; c3 = (job.epoch << 16) | (job.it_begin & 0xFFFFu);
                            need = (l_winner(g) != 0 || job.limit == 0) ? PATH_STORE : lane_path(g);
                        }
                    } else if (ROLLOUT) {
                        // Which rollouts share a warp: game-major = 32 rollouts of one game (same phase of the game, mostly the same
                        // code path); game-minor = 32 different games (the same work spread evenly over the warps).
                        uint32_t gm, it;
                        if (job.game_minor) {
                            const uint32_t q = (uint32_t)(item / job.n_games);
                            gm = (uint32_t)(item - (long long)q * job.n_games);
                            it = job.it_begin + q;
                        } else {
                            gm = (uint32_t)(item / job.it_count);
                            it = job.it_begin + (uint32_t)(item - (long long)gm * job.it_count);
                        }
                        item = (long long)gm * job.iterations + it;  // from here on: the (game, iteration) pair
                        const int node = job.sim_node[item];
                        if (node >= 0 && job.limit > 0) {  // node < 0: the iteration ended on a terminal leaf, no rollout
                            lane_load_state(g, job.states + (size_t)gm * (job.iterations + 1) + node);
                            gid = job.first_game_id + gm; c3 = (job.epoch << 16) | (it & 0xFFFFu);
                            need = (g.off_own == 15 && g.off_opp == 15) ? PATH_STORE : lane_path(g);
                        }
                    } else {
                        lane_load_state(g, job.states + item);
                        gid = job.first_game_id + (uint32_t)item;
                        need = (l_winner(g) != 0 || job.limit == 0) ? PATH_STORE : lane_path(g);
                    }
                }
            }
            continue;  // vote with the newcomers included (or take more, if some of them needed no ply at all)
        }

        // ---- vote ----
        int best = PATH_DONE, best_score = INT_MIN;
#pragma unroll
        for (int p = PATH_CLOSED; p <= PATH_WALK; ++p) {
            const uint32_t waiting = __ballot_sync(0xFFFFFFFFu, need == p);
            if (waiting) {
                int score = __popc(waiting);
                if (job.lag_weight) score += job.lag_weight * (int)__reduce_max_sync(0xFFFFFFFFu, need == p ? age : 0u);
                if (score > best_score) { best_score = score; best = p; }
            }
        }
        {   // results are written in batches: when enough of the warp waits to, or nothing else is left to do
            const uint32_t storing = __ballot_sync(0xFFFFFFFFu, need == PATH_STORE);
            if (storing && (best == PATH_DONE || __popc(storing) >= job.store_min)) best = PATH_STORE;
        }
        if (best == PATH_DONE) break;
#ifdef DIEE_LANE_STATS
        {   // [p] steps on path p, [8+p] lanes advanced; [4] votes, [5..7],[12] lanes idle / waiting for closed / walk / store at the vote
            const uint32_t adv = __ballot_sync(0xFFFFFFFFu, need == best);
            const uint32_t n0 = __ballot_sync(0xFFFFFFFFu, need == PATH_DONE), n1 = __ballot_sync(0xFFFFFFFFu, need == PATH_CLOSED);
            const uint32_t n2 = __ballot_sync(0xFFFFFFFFu, need == PATH_WALK), n3 = __ballot_sync(0xFFFFFFFFu, need == PATH_STORE);
            if (lane == 0) {
                atomicAdd(&g_lane_stats[best], 1ull); atomicAdd(&g_lane_stats[8 + best], (unsigned long long)__popc(adv));
                atomicAdd(&g_lane_stats[4], 1ull); atomicAdd(&g_lane_stats[5], (unsigned long long)__popc(n0));
                atomicAdd(&g_lane_stats[6], (unsigned long long)__popc(n1)); atomicAdd(&g_lane_stats[7], (unsigned long long)__popc(n2));
                atomicAdd(&g_lane_stats[12], (unsigned long long)__popc(n3));
            }
        }
#endif
        if (need != best) { ++age; continue; }
        age = 0;
        if (best == PATH_STORE) {
            {   // plies this job actually played (the closed-form tail below is not work): one atomic per storing batch
                const uint32_t act = __activemask();
                const uint32_t sum = __reduce_add_sync(act, k);
                if (lane == (int)(__ffs(act) - 1)) atomicAdd(job.next_item + 1, (unsigned long long)sum);
            }
            if (CC) {
                // node.rs:181-185 on the rolled-out state: a winner met BEFORE ply `limit` decides; at the cap the
                // result is 0 whatever the last ply did (the loop ends without another test)
                const uint32_t gm = (uint32_t)(item / job.iterations);
                const int w = k < job.limit ? l_winner(g) : 0, pl = job.players[gm];
                job.results[gm] = w == 0 ? 0.f : (w == pl ? 1.f : (w == -pl ? -1.f : 0.f));
                if (job.stats) job.stats[gm].rollout_plies += k;
                lane_store_state(g, job.finals + item);
            } else if (ROLLOUT) {
                if (k < job.limit) {
                    // both sides have collected everything: the remaining plies are skip_turns, i.e. the side to
                    // move alternates and the dice shown at the end are those of the last ply
                    uint32_t o[4];
                    l_philox((uint32_t)job.seed, (uint32_t)(job.seed >> 32), job.limit - 1u, gid, stream, c3, o);
                    if ((job.limit - k) & 1u) l_pass_turn(g, 0, 0);
                    g.second = 0; g.roll0 = l_die(o[0]); g.roll1 = l_die(o[1]);
                }
                lane_store_state(g, job.finals + item);
            } else {
                job.winners[item] = (int8_t)l_winner(g);  // versus.rs:231-235: the game stopped at its winner, or at the cap
                job.plies[item] = (int32_t)k;
                if (job.finals) lane_store_state(g, job.finals + item);
            }
            need = PATH_DONE;
            continue;
        }

        // ---- plies on path `best` (warp-uniform) for the lanes that wait for it: a lane keeps going while its next ply
        // needs the same path again (most do), up to job.reps plies per vote ----
        for (int rep = 0; rep < job.reps; ++rep) {
#ifdef DIEE_LANE_STATS
        { const uint32_t act = __activemask(); if (lane == (int)(__ffs(act) - 1)) { atomicAdd(&g_lane_stats[13], 1ull); atomicAdd(&g_lane_stats[14], (unsigned long long)__popc(act)); } }  // plies executed: warp-level, lanes
#endif
        uint32_t o[4];
        l_philox((uint32_t)job.seed, (uint32_t)(job.seed >> 32), k, gid, stream, c3, o);
        LanePlay pl;
        pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
        if (best == PATH_CLOSED) {
            const int hi = max(g.roll0, g.roll1), lo = min(g.roll0, g.roll1);
            LaneMasks m;
            l_closed_applies(g, m, lo, hi);
            if (g.bar_own == 0 && m.own1 != 0 && (m.own1 & ~0x3Fu) == 0) {
                // pure bear-off (lane_path sends only those here with every checker home): two loads
                const uint32_t e = __ldg(job.pb.index + l_pb_key(g));
                const uint32_t U = e & 255u;
                if (U > 0) pl = l_pb_unpack(__ldg(job.pb.plays + (e >> 8) + l_index(o[2], U)));
            } else {
                if (m.own1 != 0 || g.bar_own > 0) l_contact_select(g, m, lo, hi, -2, o[2], pl);
            }
        } else {
            LaneGen gen;
            l_movegen_walk_t<true>(g, gen, scr, LANE_CTA);
            if (gen.U > 0) pl = l_pick_walk(gen, scr, LANE_CTA, (int)l_index(o[2], (uint32_t)gen.U));
        }
        l_step(g, pl, l_die(o[0]), l_die(o[1]));
        ++k;

        // ---- what next ----
        if (ROLLOUT && !CC) need = (k == job.limit || (g.off_own == 15 && g.off_opp == 15)) ? PATH_STORE : lane_path(g);
        else need = (k == job.limit || l_winner(g) != 0) ? PATH_STORE : lane_path(g);
        if (need != best) break;
        }
    }
}

// every deferred rollout plays exactly `limit` plies (Node::simulate tests its START state, quirk Q5)
__global__ void bg_rollout_count_kernel(int n_games, uint32_t iterations, uint32_t limit, const int32_t *__restrict__ sim_node,
                                        diee_search_stats *__restrict__ stats_out) {
    const int gm = blockIdx.x * blockDim.x + threadIdx.x;
    if (gm >= n_games) return;
    unsigned long long c = 0;
    for (uint32_t it = 0; it < iterations; ++it) c += sim_node[(size_t)gm * iterations + it] >= 0 ? limit : 0u;
    stats_out[gm].rollout_plies += c;
}

// ---------------- the packed form: games queue up by the code their next ply needs ----------------
// lane_run_kernel keeps a game in the lane that took it, so a warp is as full as its vote: 11-16 of 32 lanes take part in a
// step and 6.8 are active per instruction.  For jobs much larger than the machine (the queue refills every lane anyway, the
// longest dependent chain does not matter) the games live in SHARED memory instead -- bit planes, ply counter and stream
// coordinates, 13 words per game, [word][slot] -- and wait in one of six queues: two dice | doubles | bar entries | bear-off
// table | bear-off walk | turnover (write the result, take the next item of the job).  A warp takes 32 games off the
// longest queue, plays ONE ply of each -- 32 lanes in the same code -- and appends every game to the queue of its next ply.
// There is no block barrier: the queues are rings of slot numbers whose cells are one-entry mailboxes (the producer adds
// to `tail`, writes the game, then puts the slot number into its cell; a consumer moves `head` by compare-and-swap and
// takes its cells, waiting where nothing has been put yet -- lane_pack.cuh ring_put / ring_take).  512 resident games
// per 8 warps guarantee a full queue somewhere while the job lasts (256 in flight leaves 256 waiting in six queues);
// when nothing is full for a few polls -- the tail of the job -- a warp takes what there is (1 / 2 / 4 / 16 polls:
// 3.70 / 3.70 / 3.70 / 3.92 ms for the rollouts of 8,192 games).  The first form of this kernel sorted the whole CTA
// between two barriers every ply: 35 instead of 66 warp instructions per ply at 19.3 lanes per instruction, but 8 of 12
// stalled warps sat at the barrier behind the slowest
// kind and the kernel was only 7 % faster; hence the queues.
// The plies themselves are the same device functions as above, and a game's dice and choices are keyed by (game id, ply),
// so who plays a ply changes nothing in what is played.
template <int MODE, bool ONE_WAVE>
__global__ void __launch_bounds__(PK_T, ONE_WAVE ? 2 : DIEE_PK_MINB)  // (the one-wave form runs two CTAs per SM, see launch_lane_job)
lane_pack_kernel(LaneJob job) {
    static_assert(MODE == LANE_PLAYOUT || MODE == LANE_ROLLOUT, "the lock-step rollouts are one wave: lane_run_kernel");
    constexpr bool ROLLOUT = MODE == LANE_ROLLOUT;
    extern __shared__ __align__(16) unsigned char pack_smem_raw[];
    PackSmem &sm = *reinterpret_cast<PackSmem *>(pack_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t stream = ROLLOUT ? DIEE_STREAM_ROLLOUT : DIEE_STREAM_GAME;
    // every slot starts in the turnover queue without a game: "take one"
    for (int s = tid; s < PK_S; s += PK_T) sm.st[8][s] = 0;
    for (int i = tid; i < PC_LISTS * PK_RING; i += PK_T) (&sm.ring[0][0])[i] = PK_EMPTY;
    __syncthreads();
    const int my_slots = job.pack_slots;  // (the others stay dead: a job smaller than the machine is spread over all CTAs)
    for (int s = tid; s < my_slots; s += PK_T) sm.ring[PC_TURN][s] = (unsigned)s;
    if (tid < 8) { sm.head[tid] = 0; sm.tail[tid] = tid == PC_TURN ? (unsigned)my_slots : 0u; }
    if (tid < 4) sm.area_lock[tid] = 0;
    if (tid == 0) {
        sm.n_dead = PK_S - my_slots;
        sm.n_avail = my_slots;
        sm.drain = (long long)gridDim.x * my_slots >= job.n_items ? 1 : 0;  // one wave: every item is resident from the start
    }
    __syncthreads();
    volatile unsigned *vhead = sm.head, *vtail = sm.tail;
    volatile int *vlock = sm.area_lock;
    volatile int *vdead = &sm.n_dead, *vavail = &sm.n_avail, *vdrain = &sm.drain;
#ifdef DIEE_LANE_STATS
    unsigned long long st_batches = 0, st_lanes = 0, st_polls = 0, st_kind[PC_LISTS] = {0, 0, 0, 0, 0, 0};
    unsigned long long st_cyc[PC_LISTS] = {0, 0, 0, 0, 0, 0}, st_nb[PC_LISTS] = {0, 0, 0, 0, 0, 0};
    unsigned st_kmax = 0;
    unsigned long long st_same = 0, st_fam = 0;  // plies whose next ply is of the same kind / of another closed-form kind
#endif
    int polls = 0;
    for (;;) {
        // one lane reads the CTA's counters and the warp agrees on them: what follows must be decided warp-uniformly
        unsigned ctl = 0;
        if (lane == 0) ctl = (unsigned)*vdead | ((*vavail > 0 ? 1u : 0u) << 16) | ((*vdrain ? 1u : 0u) << 17);
        ctl = __shfl_sync(0xFFFFFFFFu, ctl, 0);
        const bool draining = (ctl >> 17) & 1u;
        if ((ctl & 0xFFFFu) >= (unsigned)PK_S) break;
        if (!((ctl >> 16) & 1u)) {  // nothing waits anywhere: a cheap poll (idle warps share the issue slots with the busy ones)
            ++polls;
            __nanosleep(draining ? 40u : 200u);
            continue;
        }
        // ---- the longest queue ----
        unsigned cnt = 0;
        if (lane < PC_LISTS) {
            const unsigned h0 = vhead[lane];  // head before tail: head only grows and never passes tail
            cnt = vtail[lane] - h0;
            if (lane == PC_WALK && cnt && vlock[0] && vlock[1] && vlock[2]) cnt = 0;  // no scratch area free
            if (cnt > 1023u) cnt = 1023u;
        }
        const unsigned best = __reduce_max_sync(0xFFFFFFFFu, (cnt << 3) | (unsigned)(lane & 7));
        const int c = (int)(best & 7u);
        const int avail = (int)(best >> 3);
        const int take = avail >= 32 ? 32 : ((polls >= PK_PATIENCE || draining) ? avail : 0);
        if (take == 0) {
            ++polls;
#ifdef DIEE_LANE_STATS
            ++st_polls;
#endif
            __nanosleep(polls <= PK_PATIENCE ? 100u : 400u);  // (nothing at all to take: the other warps need the issue slots)
            continue;
        }
        int ok = 0, area = -1;
        unsigned h = 0;
        if (lane == 0) {
            if (c == PC_WALK)
                for (int a = 0; a < PK_AREAS && area < 0; ++a)
                    if (atomicCAS(&sm.area_lock[a], 0, 1) == 0) area = a;
            if (c != PC_WALK || area >= 0) {
                h = vhead[c];
                const unsigned t = vtail[c];
                if ((int)(t - h) >= take) ok = atomicCAS(&sm.head[c], h, h + (unsigned)take) == h;
                if (ok) atomicSub(&sm.n_avail, take);
                if (!ok && area >= 0) { atomicExch(&sm.area_lock[area], 0); area = -1; }
            }
        }
        ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
        if (!ok) continue;  // somebody else was faster: look again
        h = __shfl_sync(0xFFFFFFFFu, h, 0);
        area = __shfl_sync(0xFFFFFFFFu, area, 0);
        polls = 0;
        const bool act = lane < take;
        int slot = 0;
        if (act) {
            slot = (int)ring_take(&sm.ring[c][(h + (unsigned)lane) & (PK_RING - 1)]);
        }
        __threadfence_block();
        __syncwarp();
#ifdef DIEE_LANE_STATS
        ++st_batches; st_lanes += take; st_kind[c] += take;
        const long long t_batch = clock64();
#endif
        const uint32_t act_mask = __ballot_sync(0xFFFFFFFFu, act);
        int newc = -1;
        if (act) {
            LaneBoard g;
            uint32_t k, item, gid, c3;
            const uint32_t misc = sm.st[8][slot];
            bool keep = true;  // a game goes back into the slot
            if (c == PC_TURN) {
                k = sm.st[9][slot]; item = sm.st[10][slot]; gid = sm.st[11][slot]; c3 = sm.st[12][slot];
                const bool has = misc & PK_HAS_GAME;
#ifdef DIEE_LANE_STATS
                if (has && k > st_kmax) st_kmax = k;
#endif
                {   // plies this job actually played (the closed-form tail below is not work)
                    const uint32_t sum = __reduce_add_sync(act_mask, has ? k : 0u);
                    if (lane == 0 && sum) atomicAdd(job.next_item + 1, (unsigned long long)sum);
                }
                if (has) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) { g.own[w] = sm.st[w][slot]; g.opp[w] = sm.st[4 + w][slot]; }
                    unpack_misc(g, misc);
                    if (ROLLOUT) {
                        if (k < job.limit) {  // both sides have collected everything: the rest are skip_turns (see lane_run_kernel)
                            uint32_t o[4];
                            l_philox((uint32_t)job.seed, (uint32_t)(job.seed >> 32), job.limit - 1u, gid, stream, c3, o);
                            if ((job.limit - k) & 1u) l_pass_turn(g, 0, 0);
                            g.second = 0; g.roll0 = l_die(o[0]); g.roll1 = l_die(o[1]);
                        }
                        lane_store_state(g, job.finals + item);
                    } else {
                        job.winners[item] = (int8_t)l_winner(g);
                        job.plies[item] = (int32_t)k;
                        if (job.finals) lane_store_state(g, job.finals + item);
                    }
                }
                // the next items of the job
                long long first = 0;
                if (lane == 0) first = (long long)atomicAdd(job.next_item, (unsigned long long)take);
                first = __shfl_sync(act_mask, first, 0);
                const long long it_l = first + lane;
                keep = false;
                newc = PC_TURN;
                k = 0;
                if (it_l >= job.n_items) {
                    newc = PC_DEAD;
                    sm.drain = 1;
                } else if (ROLLOUT) {
                    uint32_t gm, it;
                    if (job.game_minor) {
                        const uint32_t q = (uint32_t)(it_l / job.n_games);
                        gm = (uint32_t)(it_l - (long long)q * job.n_games);
                        it = job.it_begin + q;
                    } else {
                        gm = (uint32_t)(it_l / job.it_count);
                        it = job.it_begin + (uint32_t)(it_l - (long long)gm * job.it_count);
                    }
                    item = gm * job.iterations + it;  // the (game, iteration) pair
                    const int node = job.sim_node[item];
                    if (node >= 0 && job.limit > 0) {  // node < 0: the iteration ended on a terminal leaf, no rollout
                        lane_load_state(g, job.states + (size_t)gm * (job.iterations + 1) + node);
                        gid = job.first_game_id + gm; c3 = (job.epoch << 16) | (it & 0xFFFFu);
                        keep = true;
                        newc = (g.off_own == 15 && g.off_opp == 15) ? PC_TURN : pack_class(g);
                    }
                } else {
                    item = (uint32_t)it_l;
                    lane_load_state(g, job.states + item);
                    gid = job.first_game_id + item; c3 = 0;
                    keep = true;
                    newc = (l_winner(g) != 0 || job.limit == 0) ? PC_TURN : pack_class(g);
                }
                if (keep) { sm.st[10][slot] = item; sm.st[11][slot] = gid; sm.st[12][slot] = c3; }
                else sm.st[8][slot] = 0;
            } else {
                // ---- one ply ----
#pragma unroll
                for (int w = 0; w < 4; ++w) { g.own[w] = sm.st[w][slot]; g.opp[w] = sm.st[4 + w][slot]; }
                unpack_misc(g, misc);
                k = sm.st[9][slot]; gid = sm.st[11][slot]; c3 = sm.st[12][slot];
                // A job that is resident from the start (one wave: nothing to refill with, its time is its longest games) is
                // about latency, not about how full the warps are: there a game stays with its lane while its next ply runs the
                // same code -- any closed-form kind after a closed-form kind, a walk after a walk -- up to job.reps plies per
                // visit, as in lane_run_kernel (forced packed run of the 1,024-game rollouts: 1.34 -> 1.18 ms).  In the tail of
                // a many-wave job it costs lanes (16.4 instead of 17.8 per instruction) and 1-2 % of the time: one ply per visit.
                int reps = ONE_WAVE ? job.reps : 1;
                for (;;) {
                    uint32_t o[4];
                    l_philox((uint32_t)job.seed, (uint32_t)(job.seed >> 32), k, gid, stream, c3, o);
                    LanePlay pl;
                    pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
                    if (c != PC_WALK) {
                        const int hi = max(g.roll0, g.roll1), lo = min(g.roll0, g.roll1);
                        LaneMasks m;
                        l_closed_applies(g, m, lo, hi);
                        if (g.bar_own == 0 && m.own1 != 0 && (m.own1 & ~0x3Fu) == 0) {
                            const uint32_t e = __ldg(job.pb.index + l_pb_key(g));
                            const uint32_t U = e & 255u;
                            if (U > 0) pl = l_pb_unpack(__ldg(job.pb.plays + (e >> 8) + l_index(o[2], U)));
                        } else {
                            if (m.own1 != 0 || g.bar_own > 0) l_contact_select(g, m, lo, hi, -2, o[2], pl);
                        }
                    } else {
                        LaneGen gen;
                        uint32_t *scr = &sm.scr[area][0][lane];
                        l_movegen_walk_t<true>(g, gen, scr, 32);
                        if (gen.U > 0) pl = l_pick_walk(gen, scr, 32, (int)l_index(o[2], (uint32_t)gen.U));
                    }
                    l_step(g, pl, l_die(o[0]), l_die(o[1]));
                    ++k;
                    const bool over = ROLLOUT ? (k == job.limit || (g.off_own == 15 && g.off_opp == 15)) : (k == job.limit || l_winner(g) != 0);
                    newc = over ? PC_TURN : pack_class(g);
#ifdef DIEE_LANE_STATS
                    if (newc == c) ++st_same; else if (c != PC_WALK && newc < PC_WALK) ++st_fam;
#endif
                    if (!ONE_WAVE || --reps <= 0 || (c == PC_WALK ? newc != PC_WALK : newc >= PC_WALK)) break;
                }
            }
            if (keep) {
#pragma unroll
                for (int w = 0; w < 4; ++w) { sm.st[w][slot] = g.own[w]; sm.st[4 + w][slot] = g.opp[w]; }
                sm.st[8][slot] = pack_misc(g);
                sm.st[9][slot] = k;
            }
        }
        __threadfence_block();  // the games before their ring entries
        __syncwarp();
#ifdef DIEE_LANE_STATS
        st_cyc[c] += (unsigned long long)(clock64() - t_batch); ++st_nb[c];
#endif
        if (area >= 0 && lane == 0) atomicExch(&sm.area_lock[area], 0);
        {   // ---- append every game to the queue of its next ply (one shared atomic per kind present) ----
            const uint32_t same = __match_any_sync(0xFFFFFFFFu, newc);
            const int leader = __ffs(same) - 1;
            unsigned base = 0;
            if (lane == leader) {
                if (newc >= 0 && newc < PC_LISTS) { base = atomicAdd(&sm.tail[newc], (unsigned)__popc(same)); atomicAdd(&sm.n_avail, __popc(same)); }
                else if (newc == PC_DEAD) atomicAdd(&sm.n_dead, __popc(same));
            }
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (newc >= 0 && newc < PC_LISTS) {
                ring_put(&sm.ring[newc][(base + (unsigned)__popc(same & ((1u << lane) - 1u))) & (PK_RING - 1)], (unsigned)slot);
            }
        }
    }
#ifdef DIEE_LANE_STATS
    // [0] batches, [1] games in them, [2] polls without a batch, [8 + kind] games played per kind
    if (lane == 0) {
        atomicAdd(&g_lane_stats[0], st_batches); atomicAdd(&g_lane_stats[1], st_lanes); atomicAdd(&g_lane_stats[2], st_polls);
        for (int q = 0; q < PC_LISTS; ++q) atomicAdd(&g_lane_stats[8 + q], st_kind[q]);
    }
    // (second block, read with diee_debug_lane_stats2: cycles and batches per kind, warp 0 lane 0 of every CTA)
    if (tid == 0)
        for (int q = 0; q < PC_LISTS; ++q) { atomicAdd(&g_lane_stats2[q], st_cyc[q]); atomicAdd(&g_lane_stats2[8 + q], st_nb[q]); }
    atomicMax(&g_lane_stats[15], (unsigned long long)st_kmax);
    atomicAdd(&g_lane_stats2[6], st_same); atomicAdd(&g_lane_stats2[7], st_fam);
#endif
}

template <int MODE>
static cudaError_t launch_lane_job(cudaStream_t st, LaneJob job, int *launches) {
    // tuning knobs, read once: waiting-time weight of the vote and resident CTAs (x 2 warps) per SM.  Measured on B200 with
    // 102,400 rollouts: 10 CTAs of 64 lanes per SM and weight 0 are best.
    static int sms = 0, lag_weight = 0, blocks_per_sm = 10, force_bps = 0, force_store_min = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        if (const char *e = getenv("DIEE_LANE_LAG")) lag_weight = atoi(e);
        if (const char *e = getenv("DIEE_LANE_BLOCKS_PER_SM")) blocks_per_sm = atoi(e) > 0 ? atoi(e) : 10;
        if (const char *e = getenv("DIEE_LANE_FORCE_BPS")) force_bps = atoi(e);
        if (const char *e = getenv("DIEE_LANE_STORE_MIN")) force_store_min = atoi(e);
    }
    job.lag_weight = lag_weight;
    static const int reps_env = getenv("DIEE_LANE_REPS") ? atoi(getenv("DIEE_LANE_REPS")) : 0;
    job.reps = reps_env > 0 ? reps_env : 8;  // measured: 1 / 2 / 4 / 8 / 16 / 64 plies per vote -> 1.50 / 1.48 / 1.46 / 1.44 / 1.50 / 1.69 ms (C3 rollouts)
    // one item per lane as long as the job fits ~10 CTAs per SM (the C2 / C3 sizes: fewer, fuller warps); a bigger job
    // runs at full occupancy and lanes are refilled from the queue (measured at 819,200 rollouts: 77.7 M simulations/s
    // with 10 CTAs per SM, 82.6 M with 16)
    const bool refilled = job.n_items > (long long)sms * blocks_per_sm * LANE_CTA * 5 / 4;
    const int bps = force_bps > 0 ? force_bps : (refilled ? 16 : blocks_per_sm);
    // results are written / items taken once this many lanes wait (a quarter of the warp; three quarters when the queue
    // keeps every lane busy anyway: 81.8 M -> 85.2 M simulations/s at 8,192 games)
    job.store_min = force_store_min > 0 ? force_store_min : (refilled ? 24 : 8);
    long long blocks = (job.n_items + LANE_CTA - 1) / LANE_CTA;
    if (blocks > (long long)sms * bps) blocks = (long long)sms * bps;
    // (the head only: next_item[1], the played-plies counter, is zeroed once per search / playout call by the caller)
    cudaError_t e = cudaMemsetAsync(job.next_item, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if constexpr (MODE != LANE_ROLLOUT_CC) {
        // The packed form, for jobs from ~640 items per SM on.  Measured on B200, rollouts of a 100-iteration search,
        // lane-resident / packed: 512 games 0.93 / 0.91 ms (stays lane-resident), 1,024 games 1.19 / 1.05, 2,048 games
        // 1.98 / 1.59, 4,096 3.39 / 2.29, 8,192 6.08 / 3.70, 32,768 22.6 / 12.4, 65,536 44.7 / 24.0 ms.
        // A job that fits two CTAs per SM (up to 1,024 items per SM) runs as ONE wave on two CTAs per SM -- fewer, fuller
        // CTAs (346 games each for the 1,024-game headline: 1.05 ms against 1.20 ms on three CTAs per SM of 231 games); bigger
        // jobs take three CTAs per SM with 512 resident games each and refill from the job queue.
        // DIEE_LANE_PACK=0 keeps the lane-resident kernel, =2 forces the packed one for every job size (which is how the
        // tests reach it with small batches); DIEE_PACK_MIN = items per SM from which the default picks it.
        const int pack_env = getenv("DIEE_LANE_PACK") ? atoi(getenv("DIEE_LANE_PACK")) : 1;
        const int pack_bps = getenv("DIEE_PACK_BPS") ? atoi(getenv("DIEE_PACK_BPS")) : 0;  // 0 = by job size
        const int pack_min = getenv("DIEE_PACK_MIN") ? atoi(getenv("DIEE_PACK_MIN")) : (MODE == LANE_PLAYOUT ? 400 : 640);  // (65,536 playouts: 0.936 ms lane-resident, 0.895 ms packed)
        const bool fits = job.n_items < (1ll << 31) && (MODE == LANE_PLAYOUT || (long long)job.n_games * job.iterations < (1ll << 31));
        if (fits && (pack_env == 2 || (pack_env == 1 && job.n_items >= (long long)sms * pack_min))) {
            const int bps = pack_bps > 0 ? pack_bps : (job.n_items <= 2ll * sms * PK_S ? 2 : 3);
            long long pb = (job.n_items + 31) / 32;
            if (pb > (long long)sms * bps) pb = (long long)sms * bps;
            long long per_cta = ((job.n_items + pb - 1) / pb + 31) / 32 * 32;
            job.pack_slots = (int)(per_cta < PK_S ? per_cta : PK_S);
            // (every launch: the attribute is per device, and a process may hold contexts on several)
            // resident from the start AND on two CTAs per SM: the kernel's ONE_WAVE form (several plies per visit; compiled for
            // two CTAs per SM).  A one-wave job on three CTAs per SM (1,024-1,536 items per SM) runs the other instance.
            if (bps == 2 && pb * job.pack_slots >= job.n_items) {
                if ((e = cudaFuncSetAttribute(lane_pack_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PackSmem))) != cudaSuccess) return e;
                lane_pack_kernel<MODE, true><<<(unsigned)pb, PK_T, sizeof(PackSmem), st>>>(job);
            } else {
                if ((e = cudaFuncSetAttribute(lane_pack_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PackSmem))) != cudaSuccess) return e;
                lane_pack_kernel<MODE, false><<<(unsigned)pb, PK_T, sizeof(PackSmem), st>>>(job);
            }
            if (launches) *launches += 1;
            return cudaGetLastError();
        }
    }
    lane_run_kernel<MODE><<<(unsigned)blocks, LANE_CTA, 0, st>>>(job);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_bg_playout(cudaStream_t st, const diee_bg_state *starts, int n, uint64_t seed, uint32_t first_game_id,
                              int round_limit, int8_t *winners_out, int32_t *plies_out, diee_bg_state *finals_out,
                              unsigned long long *queue_head, const PbTable &pb, int *launches) {
    if (n <= 0) return cudaSuccess;
    LaneJob job{};
    job.next_item = queue_head;
    job.pb = pb;
    job.n_items = n; job.limit = (uint32_t)round_limit; job.seed = seed; job.first_game_id = first_game_id;
    job.states = starts; job.finals = finals_out; job.winners = winners_out; job.plies = plies_out;
    return launch_lane_job<LANE_PLAYOUT>(st, job, launches);
}

cudaError_t launch_bg_rollouts(cudaStream_t st, int n_games, const diee_mcts_cfg &cfg, uint32_t it_begin, uint32_t it_end,
                               uint64_t seed, uint32_t first_game_id, uint32_t epoch, const PoolPtrs &pp,
                               unsigned long long *queue_head, int *launches) {
    const long long items = (long long)n_games * (it_end - it_begin);
    if (items <= 0 || cfg.simulate_round_limit == 0) return cudaSuccess;
    LaneJob job{};
    job.n_items = items; job.iterations = cfg.iterations; job.it_begin = it_begin; job.it_count = it_end - it_begin;
    job.n_games = (uint32_t)n_games;
    // (measured: game-minor order changes nothing at 1,024 games -- 1.21 vs 1.19 ms -- and costs 10 % at 8,192 games, where
    // lanes of one game share their code path; game-major stays the default)
    static const int order_env = getenv("DIEE_LANE_GAME_MINOR") ? atoi(getenv("DIEE_LANE_GAME_MINOR")) : 0;
    job.game_minor = order_env;
    job.limit = cfg.simulate_round_limit; job.seed = seed; job.first_game_id = first_game_id; job.epoch = epoch;
    job.states = static_cast<const diee_bg_state *>(pp.states); job.sim_node = pp.sim_node;
    job.finals = static_cast<diee_bg_state *>(pp.finals);
    job.next_item = queue_head;
    job.pb = pp.pb;
    return launch_lane_job<LANE_ROLLOUT>(st, job, launches);
}

// the rollouts of iteration `it` of a lock-step search (DIEE_MODE_ROLLOUT_CHECK_CURRENT) for games [g0, g0 + n_games)
cudaError_t launch_bg_rollouts_cc(cudaStream_t st, int g0, int n_games, const diee_mcts_cfg &cfg, uint32_t it, uint64_t seed,
                                  uint32_t first_game_id, uint32_t epoch, const PoolPtrs &pp, const int8_t *players, float *results,
                                  diee_search_stats *stats, unsigned long long *queue_head, int *launches) {
    if (n_games <= 0 || cfg.simulate_round_limit == 0) return cudaSuccess;
    LaneJob job{};
    job.n_items = n_games; job.iterations = cfg.iterations; job.it_begin = it; job.it_count = 1;
    job.limit = cfg.simulate_round_limit; job.seed = seed; job.first_game_id = first_game_id; job.epoch = epoch;
    job.states = static_cast<const diee_bg_state *>(pp.states); job.sim_node = pp.sim_node;
    job.finals = static_cast<diee_bg_state *>(pp.finals);
    job.next_item = queue_head;
    job.pb = pp.pb;
    job.first_item = (uint32_t)g0; job.players = players; job.results = results; job.stats = stats;
    return launch_lane_job<LANE_ROLLOUT_CC>(st, job, launches);
}

cudaError_t launch_bg_rollout_count(cudaStream_t st, int n_games, const diee_mcts_cfg &cfg, const PoolPtrs &pp,
                                    diee_search_stats *stats_out, int *launches) {
    if (!stats_out || n_games <= 0) return cudaSuccess;
    bg_rollout_count_kernel<<<(n_games + 127) / 128, 128, 0, st>>>(n_games, cfg.iterations, cfg.simulate_round_limit, pp.sim_node, stats_out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace diee
#ifdef DIEE_LANE_STATS
extern "C" int diee_debug_lane_stats2(unsigned long long *out16, int reset) {
    cudaMemcpyFromSymbol(out16, diee::g_lane_stats2, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(diee::g_lane_stats2, z, sizeof z); }
    return 0;
}
extern "C" int diee_debug_lane_stats(unsigned long long *out16, int reset) {
    cudaMemcpyFromSymbol(out16, diee::g_lane_stats, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(diee::g_lane_stats, z, sizeof z); }
    return 0;
}
#endif
