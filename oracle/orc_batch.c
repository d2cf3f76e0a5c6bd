/*
 * oracle/orc_batch.c -- CPU ORACLE (test infrastructure only; see orc.h).
 * The reference's parallel structure for the CPU path: one task per game over a thread pool
 * (games.par_iter().map(mct_search), src/versus.rs:303-306; rayon pool sized in src/main.rs:100-110).
 * Used by bench.py's cpu_baseline / --impl reference legs to time the oracle on the host cores.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "orc.h"

typedef struct {
    int kind; /* 0 = mcts bg, 1 = playout bg */
    const orc_bg_state *states;
    int n;
    const int8_t *players;
    const orc_mcts_cfg *cfg;
    uint64_t seed;
    uint32_t first_game_id, epoch;
    int round_limit;
    orc_move *best;
    int32_t *status;
    int8_t *winners;
    int32_t *plies;
    atomic_int next;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (;;) {
        int i = atomic_fetch_add(&j->next, 1);
        if (i >= j->n) break;
        if (j->kind == 0) {
            int32_t nn = 0;
            j->status[i] = orc_mcts_search_bg(&j->states[i], j->players[i], j->cfg, j->seed, j->first_game_id + (uint32_t)i,
                                              j->epoch, &j->best[i], NULL, NULL, &nn);
        } else {
            orc_bg_state s = j->states[i];
            j->winners[i] = (int8_t)orc_bg_playout(&s, j->seed, j->first_game_id + (uint32_t)i, j->round_limit, &j->plies[i]);
        }
    }
    return NULL;
}

static void run(job_t *j, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    atomic_store(&j->next, 0);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, j);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
}

int orc_mcts_search_bg_batch(const orc_bg_state *states, int n, const int8_t *players, const orc_mcts_cfg *cfg,
                             uint64_t seed, uint32_t first_game_id, uint32_t epoch, orc_move *best, int32_t *status,
                             int nthreads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.kind = 0; j.states = states; j.n = n; j.players = players; j.cfg = cfg; j.seed = seed;
    j.first_game_id = first_game_id; j.epoch = epoch; j.best = best; j.status = status;
    run(&j, nthreads);
    return 0;
}

int orc_bg_playout_batch(const orc_bg_state *states, int n, uint64_t seed, uint32_t first_game_id, int round_limit,
                         int8_t *winners, int32_t *plies, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.kind = 1; j.states = states; j.n = n; j.seed = seed; j.first_game_id = first_game_id;
    j.round_limit = round_limit; j.winners = winners; j.plies = plies;
    run(&j, nthreads);
    return 0;
}
