// Host model of the queue protocol of lane_pack_kernel (die_e_b200/csrc/lane_kernels.cu): W worker threads ("warps")
// share S slots ("resident games") that wait in K rings of slot numbers.  A producer reserves positions with an atomic
// add on `tail`, then PUTS each entry into its cell (compare-and-swap EMPTY -> slot, waiting while the cell is still full);
// a consumer moves `head` by compare-and-swap over what `tail` shows and TAKES its cells (exchange with EMPTY, waiting
// while a cell is still empty).  A cell is a one-entry mailbox: the first form of the protocol wrote and read cells with
// plain stores and loads, and this model caught the flaw -- a consumer that is slow to read what it claimed while the
// other workers recycle slots through the same ring gets its entry overwritten by the producer of position p + ring size.
// The model checks the invariants the kernel relies on: a slot is held by at most one worker at a time, no slot is lost
// or duplicated, and every slot makes its PLIES steps.
// ring_model ... 1 as 6th argument runs the FIRST form (expected to fail under preemption).
//   usage: ring_model <workers> <slots> <kinds> <plies> <seed>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static constexpr uint16_t EMPTY = 0xFFFF;
static constexpr int BATCH = 32;

struct Ring {
    std::vector<std::atomic<uint16_t>> e;
    std::atomic<unsigned> head{0}, tail{0};
    explicit Ring(int n) : e(n) { for (auto &x : e) x.store(EMPTY); }
};

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 8, S = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 6;
    const int PLIES = argc > 4 ? atoi(argv[4]) : 2000;
    const unsigned seed = argc > 5 ? (unsigned)atoi(argv[5]) : 1u;
    const bool first_form = argc > 6 && atoi(argv[6]) != 0;
    int ring_size = 1;
    while (ring_size < S) ring_size <<= 1;
    std::vector<Ring *> rings;
    for (int k = 0; k < K; ++k) rings.push_back(new Ring(ring_size));
    std::vector<std::atomic<int>> holder(S);      // which worker holds the slot (-1: it waits in a ring)
    std::vector<int> plies(S, 0);                 // steps the slot has made (written only by its holder)
    std::vector<uint32_t> state(S);               // the slot's "game": decides the kind of its next ply
    std::atomic<int> n_done{0}, n_avail{S}, errors{0};
    for (int s = 0; s < S; ++s) {
        holder[s].store(-1);
        state[s] = seed * 2654435761u + (uint32_t)s * 40503u;
        rings[0]->e[s].store((uint16_t)s);        // everybody starts in ring 0 (the kernel: the turnover queue)
    }
    rings[0]->tail.store((unsigned)S);
    auto worker = [&](int w) {
        int polls = 0;
        while (n_done.load() < S && errors.load() == 0) {
            if (n_avail.load() <= 0) { std::this_thread::yield(); continue; }
            // the longest ring
            int c = -1, avail = 0;
            for (int k = 0; k < K; ++k) {
                const unsigned h = rings[k]->head.load();     // head before tail
                const int a = (int)(rings[k]->tail.load() - h);
                if (a > avail) { avail = a; c = k; }
            }
            const int take = avail >= BATCH ? BATCH : (polls >= 4 ? avail : 0);
            if (take <= 0) { ++polls; std::this_thread::yield(); continue; }
            Ring &r = *rings[c];
            unsigned h = r.head.load();
            if ((int)(r.tail.load() - h) < take || !r.head.compare_exchange_strong(h, h + (unsigned)take)) continue;
            n_avail.fetch_sub(take);
            polls = 0;
            int slots[BATCH], next[BATCH];
            for (int i = 0; i < take; ++i) {
                std::atomic<uint16_t> &e = r.e[(h + (unsigned)i) & (unsigned)(ring_size - 1)];
                uint16_t v;
                if (first_form) {
                    while ((v = e.load(std::memory_order_acquire)) == EMPTY) std::this_thread::yield();  // reserved, not written yet
                    e.store(EMPTY, std::memory_order_relaxed);
                } else {
                    while ((v = e.exchange(EMPTY, std::memory_order_acq_rel)) == EMPTY) std::this_thread::yield();
                }
                slots[i] = v;
                int expect = -1;
                if (!holder[v].compare_exchange_strong(expect, w)) { errors.fetch_add(1); fprintf(stderr, "slot %d taken twice\n", (int)v); }
            }
            // one "ply" of each game: advance its state, classify it again
            for (int i = 0; i < take; ++i) {
                const int s = slots[i];
                state[s] = state[s] * 1664525u + 1013904223u;
                ++plies[s];
                next[i] = plies[s] >= PLIES ? -1 : (int)((state[s] >> 16) % (uint32_t)K);
            }
            // append every game to the ring of its next ply (one atomic per kind present), or retire it
            for (int k = -1; k < K; ++k) {
                int cnt = 0;
                for (int i = 0; i < take; ++i) cnt += next[i] == k;
                if (!cnt) continue;
                if (k < 0) {
                    for (int i = 0; i < take; ++i) if (next[i] < 0) holder[slots[i]].store(-2);
                    n_done.fetch_add(cnt);
                    continue;
                }
                const unsigned base = rings[k]->tail.fetch_add((unsigned)cnt);
                n_avail.fetch_add(cnt);
                int j = 0;
                for (int i = 0; i < take; ++i)
                    if (next[i] == k) {
                        holder[slots[i]].store(-1);   // released BEFORE the entry is published
                        std::atomic<uint16_t> &e = rings[k]->e[(base + (unsigned)j++) & (unsigned)(ring_size - 1)];
                        if (first_form) {
                            if (e.load() != EMPTY) { errors.fetch_add(1); fprintf(stderr, "ring %d wrapped onto a live entry\n", k); }
                            e.store((uint16_t)slots[i], std::memory_order_release);
                        } else {
                            uint16_t expect = EMPTY;
                            while (!e.compare_exchange_weak(expect, (uint16_t)slots[i], std::memory_order_acq_rel)) { expect = EMPTY; std::this_thread::yield(); }
                        }
                    }
            }
        }
    };
    std::vector<std::thread> th;
    for (int w = 0; w < W; ++w) th.emplace_back(worker, w);
    for (auto &t : th) t.join();
    long long total = 0;
    for (int s = 0; s < S; ++s) {
        total += plies[s];
        if (plies[s] != PLIES || holder[s].load() != -2) { errors.fetch_add(1); fprintf(stderr, "slot %d: %d plies, holder %d\n", s, plies[s], holder[s].load()); break; }
    }
    for (int k = 0; k < K; ++k)
        if (rings[k]->head.load() != rings[k]->tail.load()) { errors.fetch_add(1); fprintf(stderr, "ring %d not drained\n", k); }
    if (errors.load()) return 1;
    printf("ring protocol ok: %d workers, %d slots, %d rings, %lld plies\n", W, S, K, total);
    return 0;
}
