// alpha.cu -- host side of the AlphaZero search and the self-play driver,
// behind the extern "C" entry points diee_alpha_search / diee_selfplay_run / diee_dirichlet.
// Reference: src/mcts/alpha_mcts.rs:91-202, src/mcts/noise.rs:27-34, src/mcts/utils.rs:42-58,
// src/alphazero/alpha_parallel.rs:101-231, src/alphazero/alphazero.rs:69-73,129-137.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "alpha_launch.h"
#include "bg_device.cuh"
#include "ctx.h"
#include "launchers.h"

using namespace diee;

// ---------------- Dirichlet(alpha * 1_n) on the DIRICHLET stream (noise.rs:27-34) ----------------
// rand_distr's Dirichlet is a vector of Gamma(alpha, 1) draws divided by their sum.  Contract
// (include/diee.h): Marsaglia-Tsang Gamma in double, boosted by U^(1/alpha) for alpha < 1; component i
// draws from Philox blocks (c0 = 2*attempt, 2*attempt+1; c1 = epoch; c2 = DIRICHLET; c3 = i).
static inline double unit53(uint32_t hi, uint32_t lo) {
    const uint64_t m = (((uint64_t)hi << 21) ^ ((uint64_t)lo >> 11)) & ((1ull << 53) - 1);
    const double u = (double)m * (1.0 / 9007199254740992.0);
    return u;
}
static inline double positive(double u) { return u <= 0.0 ? 1.0 / 9007199254740992.0 : u; }

static double gamma_variate(uint64_t seed, uint32_t epoch, uint32_t comp, double alpha) {
    const double shape = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / std::sqrt(9.0 * d);
    for (uint32_t attempt = 0;; ++attempt) {
        uint32_t a[4], b[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 2 * attempt, epoch, DIEE_STREAM_DIRICHLET, comp, a);
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 2 * attempt + 1, epoch, DIEE_STREAM_DIRICHLET, comp, b);
        const double u1 = positive(unit53(a[0], a[1])), u2 = unit53(a[2], a[3]);
        const double x = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
        const double t = 1.0 + c * x;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        const double u = positive(unit53(b[0], b[1]));
        if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) {
            double g = d * v;
            if (alpha < 1.0) g *= std::pow(positive(unit53(b[2], b[3])), 1.0 / alpha);
            return g;
        }
    }
}

static void dirichlet_sample(uint64_t seed, uint32_t epoch, float alpha, int n, float *out) {
    std::vector<double> g((size_t)n);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) { g[i] = gamma_variate(seed, epoch, (uint32_t)i, (double)alpha); sum += g[i]; }
    for (int i = 0; i < n; ++i) out[i] = (float)(g[i] / sum);
}

static int default_max_nodes(const diee_mcts_cfg *cfg) { return 1 + ((int)cfg->iterations + 1) * 128; }

// one alpha_mcts_parallel over device-resident states; fills the ctx arena
// K > 1 (with virtual loss vl) is the NON-PARITY throughput mode of alpha_kernels.cu; K <= 1 is the reference's search
static int32_t alpha_search_device(diee_ctx *ctx, diee_net *net, const diee_bg_state *d_states, int n, const uint32_t *d_ids,
                                   const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int max_nodes, AlphaPool &P,
                                   int K = 0, float vl = 0.f) {
    const size_t total = (size_t)n * (size_t)max_nodes;
    const bool vl_path = K > 1 || K == -1;  // -1: the virtual-loss kernels with ONE leaf per step (tests)
    if (K == -1) K = 1;
    const size_t rows = (size_t)n * (size_t)(vl_path ? K : 1);  // rows of one batched forward
    RESERVE(ctx->a_state, sizeof(diee_bg_state) * total);
    RESERVE(ctx->a_parent, 4 * total);
    RESERVE(ctx->a_first, 4 * total);
    RESERVE(ctx->a_nchild, 4 * total);
    RESERVE(ctx->a_visits, 4 * total);
    RESERVE(ctx->a_value, 4 * total);
    RESERVE(ctx->a_prior, 4 * total);
    RESERVE(ctx->a_action, 4 * total);
    RESERVE(ctx->a_nnodes, 4 * (size_t)n);
    RESERVE(ctx->a_selg, 4 * rows);
    RESERVE(ctx->a_seln, 4 * rows);
    RESERVE(ctx->a_status, 4 * (size_t)n);
    RESERVE(ctx->a_any, 4 * (size_t)cfg->iterations + 4);
    RESERVE(ctx->a_batch, sizeof(diee_bg_state) * rows);
    RESERVE(ctx->a_policy, sizeof(float) * DIEE_ACTION_SPACE * rows);
    RESERVE(ctx->a_valueout, sizeof(float) * rows);
    RESERVE(ctx->a_dir, sizeof(float) * DIEE_ACTION_SPACE);
    P.max_nodes = max_nodes;
    P.state = ctx->a_state.p;
    P.parent = (int32_t *)ctx->a_parent.p; P.first = (int32_t *)ctx->a_first.p; P.nchild = (int32_t *)ctx->a_nchild.p;
    P.visits = (float *)ctx->a_visits.p; P.value = (float *)ctx->a_value.p; P.prior = (float *)ctx->a_prior.p;
    P.action = (uint32_t *)ctx->a_action.p; P.n_nodes = (int32_t *)ctx->a_nnodes.p;
    P.sel_game = (int32_t *)ctx->a_selg.p; P.sel_node = (int32_t *)ctx->a_seln.p; P.status = (int32_t *)ctx->a_status.p;
    P.any_selected = (int32_t *)ctx->a_any.p; P.batch = (diee_bg_state *)ctx->a_batch.p;
    P.policy = (float *)ctx->a_policy.p; P.value_out = (float *)ctx->a_valueout.p; P.dirichlet = (const float *)ctx->a_dir.p;
    cudaStream_t st = ctx->stream;
    // one Dirichlet vector of length 1352 for the whole batch (noise.rs:27-34, Q11)
    CU(cudaStreamSynchronize(st));  // dir_host may still be in flight from the previous wave
    ctx->dir_host.resize(DIEE_ACTION_SPACE);
    dirichlet_sample(seed, epoch, cfg->dirichlet_alpha, DIEE_ACTION_SPACE, ctx->dir_host.data());
    CU(cudaMemcpyAsync(ctx->a_dir.p, ctx->dir_host.data(), sizeof(float) * DIEE_ACTION_SPACE, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->a_any.p, 0, 4 * (size_t)cfg->iterations + 4, st));
    // root phase: forward_policy on the roots (:97-104)
    int32_t rc = diee_net_forward_dev(ctx, net, d_states, n, P.policy, P.value_out);
    if (rc != DIEE_OK) return rc;
    ctx->net_evals += (uint64_t)n;
    CU(launch_alpha_root(st, P, d_states, d_ids, n, *cfg, seed, epoch));
    ctx->launches += 1;
    if (vl_path) {  // non-parity: iterations / K steps of up to K leaves per game
        for (int left = (int)cfg->iterations; left > 0; left -= K) {
            const int budget = left < K ? left : K;
            CU(launch_alpha_select_vl(st, P, n, *cfg, K, vl, budget));
            rc = diee_net_forward_dev(ctx, net, P.batch, (int32_t)rows, P.policy, P.value_out);
            if (rc != DIEE_OK) return rc;
            ctx->net_evals += (uint64_t)n * (uint64_t)budget;
            CU(launch_alpha_expand_vl(st, P, d_ids, n, *cfg, seed, epoch, K, vl));
            ctx->launches += 2;
        }
        return DIEE_OK;
    }
    for (uint32_t it = 0; it < cfg->iterations; ++it) {  // :149-201
        CU(launch_alpha_select(st, P, n, *cfg, (int)it));
        rc = diee_net_forward_dev(ctx, net, P.batch, n, P.policy, P.value_out);  // forward_t on all N slots (:186)
        if (rc != DIEE_OK) return rc;
        ctx->net_evals += (uint64_t)n;
        CU(launch_alpha_expand(st, P, d_ids, n, *cfg, seed, epoch, (int)it));
        ctx->launches += 2;
    }
    return DIEE_OK;
}

static int32_t check_alpha_args(diee_ctx *ctx, diee_net *net, int n, const diee_mcts_cfg *cfg, uint32_t epoch) {
    if (!ctx) return DIEE_ERR_INVALID;
    if (!net || !cfg || n < 0) return fail(ctx, DIEE_ERR_INVALID, "alpha_search: bad argument");
    if (cfg->iterations == 0 || cfg->iterations > 65534u) return fail(ctx, DIEE_ERR_INVALID, "alpha_search: iterations must be in 1..65534");
    // Dirichlet::new(&vec![alpha; n]).unwrap() panics for alpha <= 0 (noise.rs:29); here: an error, never a hang in the
    // Gamma rejection loop or a silent NaN prior.  epsilon mixes two distributions, so it must lie in [0, 1].
    if (!(cfg->dirichlet_alpha > 0.f) || !std::isfinite(cfg->dirichlet_alpha))
        return fail(ctx, DIEE_ERR_INVALID, "alpha_search: dirichlet_alpha must be a finite number > 0");
    if (!(cfg->dirichlet_epsilon >= 0.f && cfg->dirichlet_epsilon <= 1.f))
        return fail(ctx, DIEE_ERR_INVALID, "alpha_search: dirichlet_epsilon must be in [0, 1]");
    (void)epoch;
    return DIEE_OK;
}

extern "C" {

int32_t diee_dirichlet(uint64_t seed, uint32_t epoch, float alpha, int32_t n, float *out) {
    if (!out || n <= 0 || !(alpha > 0.f) || !std::isfinite(alpha)) return DIEE_ERR_INVALID;
    dirichlet_sample(seed, epoch, alpha, n, out);
    return DIEE_OK;
}

int32_t diee_alpha_search(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                          const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, uint16_t *root_ids_out,
                          diee_move *root_moves_out, float *root_visits_out, int32_t *root_counts_out, int32_t *status_out,
                          diee_anode *nodes_out, int32_t *n_nodes_out) {
    int32_t rc = check_alpha_args(ctx, net, n, cfg, epoch);
    if (rc != DIEE_OK) return rc;
    if (n == 0) return DIEE_OK;
    if (!states || !game_ids || !root_ids_out || !root_visits_out || !root_counts_out || !status_out)
        return fail(ctx, DIEE_ERR_INVALID, "alpha_search: bad argument");
    for (int i = 0; i < n; ++i)
        if (states[i].roll[0] == 0 && states[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "alpha_search: state %d has not been rolled", i);
    if (max_nodes <= 0) max_nodes = default_max_nodes(cfg);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    RESERVE(ctx->a_states_in, sizeof(diee_bg_state) * (size_t)n);
    RESERVE(ctx->a_ids_in, 4 * (size_t)n);
    RESERVE(ctx->a_root_ids, 2 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_moves, 4 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_visits, 4 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_counts, 4 * (size_t)n);
    CU(cudaMemcpyAsync(ctx->a_states_in.p, states, sizeof(diee_bg_state) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->a_ids_in.p, game_ids, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    AlphaPool P;
    rc = alpha_search_device(ctx, net, (const diee_bg_state *)ctx->a_states_in.p, n, (const uint32_t *)ctx->a_ids_in.p, cfg, seed, epoch,
                             max_nodes, P);
    if (rc != DIEE_OK) return rc;
    CU(launch_alpha_root_out(st, P, n, (uint16_t *)ctx->a_root_ids.p, (uint32_t *)ctx->a_root_moves.p, (float *)ctx->a_root_visits.p,
                             (int32_t *)ctx->a_root_counts.p));
    ctx->launches += 1;
    const size_t nm = (size_t)n * DIEE_MAX_MOVES;
    CU(cudaMemcpyAsync(root_ids_out, ctx->a_root_ids.p, 2 * nm, cudaMemcpyDeviceToHost, st));
    if (root_moves_out) CU(cudaMemcpyAsync(root_moves_out, ctx->a_root_moves.p, 4 * nm, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(root_visits_out, ctx->a_root_visits.p, 4 * nm, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(root_counts_out, ctx->a_root_counts.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status_out, ctx->a_status.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (n_nodes_out) CU(cudaMemcpyAsync(n_nodes_out, ctx->a_nnodes.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (nodes_out) {
        const size_t total = (size_t)n * (size_t)max_nodes;
        std::vector<int32_t> parent(total), first(total), nchild(total);
        std::vector<float> visits(total), value(total), prior(total);
        std::vector<uint32_t> action(total);
        std::vector<diee_bg_state> state(total);
        CU(cudaMemcpyAsync(parent.data(), P.parent, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(first.data(), P.first, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(nchild.data(), P.nchild, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(visits.data(), P.visits, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(value.data(), P.value, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(prior.data(), P.prior, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(action.data(), P.action, 4 * total, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(state.data(), P.state, sizeof(diee_bg_state) * total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t i = 0; i < total; ++i) {
            diee_anode &a = nodes_out[i];
            a.parent = parent[i]; a.first_child = first[i]; a.n_children = nchild[i];
            a.visits = visits[i]; a.value = value[i]; a.prior = prior[i];
            memcpy(&a.action, &action[i], 4);
            a.state = state[i];
        }
    }
    CU(cudaStreamSynchronize(st));
    return DIEE_OK;
}

// device-resident form: all pointers are device pointers, results are valid after diee_sync()
int32_t diee_alpha_search_dev(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                              const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, uint16_t *root_ids_out,
                              diee_move *root_moves_out, float *root_visits_out, int32_t *root_counts_out, int32_t *status_out) {
    int32_t rc = check_alpha_args(ctx, net, n, cfg, epoch);
    if (rc != DIEE_OK) return rc;
    if (n == 0) return DIEE_OK;
    if (!states || !game_ids || !root_ids_out || !root_moves_out || !root_visits_out || !root_counts_out || !status_out)
        return fail(ctx, DIEE_ERR_INVALID, "alpha_search_dev: bad argument");
    if (max_nodes <= 0) max_nodes = default_max_nodes(cfg);
    CU(cudaSetDevice(ctx->device));
    AlphaPool P;
    rc = alpha_search_device(ctx, net, states, n, game_ids, cfg, seed, epoch, max_nodes, P);
    if (rc != DIEE_OK) return rc;
    CU(launch_alpha_root_out(ctx->stream, P, n, root_ids_out, (uint32_t *)root_moves_out, root_visits_out, root_counts_out));
    CU(cudaMemcpyAsync(status_out, ctx->a_status.p, 4 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += 1;
    return DIEE_OK;
}

// non-parity search: up to `leaves_per_game` leaves per game and step with virtual loss (host buffers)
int32_t diee_alpha_search_vl(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, const uint32_t *game_ids,
                             const diee_mcts_cfg *cfg, uint64_t seed, uint32_t epoch, int32_t max_nodes, int32_t leaves_per_game,
                             float virtual_loss, uint16_t *root_ids_out, diee_move *root_moves_out, float *root_visits_out,
                             int32_t *root_counts_out, int32_t *status_out) {
    int32_t rc = check_alpha_args(ctx, net, n, cfg, epoch);
    if (rc != DIEE_OK) return rc;
    if (n == 0) return DIEE_OK;
    if (!states || !game_ids || !root_ids_out || !root_visits_out || !root_counts_out || !status_out || leaves_per_game < 1 ||
        leaves_per_game > 64 || !(virtual_loss >= 0.f))
        return fail(ctx, DIEE_ERR_INVALID, "alpha_search_vl: bad argument");
    for (int i = 0; i < n; ++i)
        if (states[i].roll[0] == 0 && states[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "alpha_search_vl: state %d has not been rolled", i);
    if (max_nodes <= 0) max_nodes = default_max_nodes(cfg);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    RESERVE(ctx->a_states_in, sizeof(diee_bg_state) * (size_t)n);
    RESERVE(ctx->a_ids_in, 4 * (size_t)n);
    RESERVE(ctx->a_root_ids, 2 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_moves, 4 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_visits, 4 * (size_t)n * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_counts, 4 * (size_t)n);
    CU(cudaMemcpyAsync(ctx->a_states_in.p, states, sizeof(diee_bg_state) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->a_ids_in.p, game_ids, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    AlphaPool P;
    // K = 1 still runs the virtual-loss kernels (K passed as 2 would change the search): force the vl path with K >= 1
    rc = alpha_search_device(ctx, net, (const diee_bg_state *)ctx->a_states_in.p, n, (const uint32_t *)ctx->a_ids_in.p, cfg, seed, epoch,
                             max_nodes, P, leaves_per_game == 1 ? -1 : leaves_per_game, virtual_loss);
    if (rc != DIEE_OK) return rc;
    CU(launch_alpha_root_out(st, P, n, (uint16_t *)ctx->a_root_ids.p, (uint32_t *)ctx->a_root_moves.p, (float *)ctx->a_root_visits.p,
                             (int32_t *)ctx->a_root_counts.p));
    ctx->launches += 1;
    const size_t nm = (size_t)n * DIEE_MAX_MOVES;
    CU(cudaMemcpyAsync(root_ids_out, ctx->a_root_ids.p, 2 * nm, cudaMemcpyDeviceToHost, st));
    if (root_moves_out) CU(cudaMemcpyAsync(root_moves_out, ctx->a_root_moves.p, 4 * nm, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(root_visits_out, ctx->a_root_visits.p, 4 * nm, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(root_counts_out, ctx->a_root_counts.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status_out, ctx->a_status.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return DIEE_OK;
}

uint64_t diee_net_eval_count(const diee_ctx *ctx) { return ctx ? ctx->net_evals : 0; }

// self_play_parallel (alpha_parallel.rs:101-231).  Live games stay resident on the device between
// waves; per wave the host reads back only the root children (action, visits), forms pi = visits /
// sum, pi^(1/T) (utils.rs:42-58, alpha_parallel.rs:164-166), samples the move
// (weighted_select_tensor_idx, alphazero.rs:129-137), records the MemoryFragment and sends the chosen
// moves + next rolls back for the apply kernel.
int32_t diee_selfplay_run(diee_ctx *ctx, diee_net *net, int32_t n_games, const diee_mcts_cfg *cfg, float temperature,
                          uint64_t seed, uint32_t first_game_id, int32_t max_nodes, diee_traj_record *rec_out, int32_t rec_cap,
                          uint16_t *pi_ids_out, float *pi_vals_out, int32_t pi_cap, int32_t *n_rec_out, int32_t *n_pi_out,
                          int32_t *n_waves_out) {
    return diee_selfplay_run_ex(ctx, net, n_games, cfg, temperature, seed, first_game_id, max_nodes, nullptr, rec_out, rec_cap,
                                pi_ids_out, pi_vals_out, pi_cap, n_rec_out, n_pi_out, n_waves_out, nullptr);
}

int32_t diee_selfplay_run_ex(diee_ctx *ctx, diee_net *net, int32_t n_games, const diee_mcts_cfg *cfg, float temperature,
                             uint64_t seed, uint32_t first_game_id, int32_t max_nodes, const diee_selfplay_opts *opts,
                             diee_traj_record *rec_out, int32_t rec_cap, uint16_t *pi_ids_out, float *pi_vals_out, int32_t pi_cap,
                             int32_t *n_rec_out, int32_t *n_pi_out, int32_t *n_waves_out, diee_selfplay_report *report_out) {
    int32_t rc = check_alpha_args(ctx, net, n_games, cfg, 0);
    if (rc != DIEE_OK) return rc;
    diee_selfplay_opts o{};
    if (opts) o = *opts;
    const bool refill = (o.flags & DIEE_SP_REFILL) != 0;
    if (o.max_waves < 0 || (o.flags & ~DIEE_SP_REFILL) != 0 || o.leaves_per_game < 0 || o.leaves_per_game > 64 || o.target_games < 0 ||
        !(o.virtual_loss >= 0.f) || (refill && o.target_games == 0 && o.max_waves == 0) || (!refill && o.target_games != 0))
        return fail(ctx, DIEE_ERR_INVALID, "selfplay_run_ex: bad option (a refilled run needs target_games or max_waves)");
    const int K = o.leaves_per_game > 1 ? o.leaves_per_game : 0;  // > 1: the non-parity search of alpha_kernels.cu
    diee_selfplay_report rep{};
    if (!rec_out || !pi_ids_out || !pi_vals_out || !n_rec_out || !n_pi_out || !(temperature > 0.f))
        return fail(ctx, DIEE_ERR_INVALID, "selfplay_run: bad argument");
    *n_rec_out = 0; *n_pi_out = 0;
    if (n_waves_out) *n_waves_out = 0;
    if (n_games == 0) return DIEE_OK;
    if (max_nodes <= 0) max_nodes = default_max_nodes(cfg);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    struct Mem { diee_bg_state st; int8_t player; int pi_off, pi_n, ply; };
    const int N = n_games;
    std::vector<diee_bg_state> state((size_t)N);
    std::vector<int> n_rounds((size_t)N, 0);
    std::vector<char> alive((size_t)N, 1);
    std::vector<std::vector<Mem>> mem((size_t)N);
    std::vector<uint16_t> sp_ids;
    std::vector<float> sp_vals;
    const float tinv = (float)(1.0 / (double)temperature);
    std::vector<uint32_t> gid((size_t)N);  // the game a slot plays (refilled runs give a finished game's slot a new one)
    uint32_t next_gid = first_game_id + (uint32_t)N;
    auto new_game = [&](int g, uint32_t id) {  // T::new() + roll_die (:103-111)
        static const int8_t opening[24] = {2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2};
        gid[g] = id;
        memset(&state[g], 0, sizeof(diee_bg_state));
        memcpy(state[g].pts, opening, 24);
        state[g].player = -1;
        uint32_t w[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 0, id, DIEE_STREAM_INIT, 0, w);
        state[g].roll[0] = (uint8_t)die_of(w[0]);
        state[g].roll[1] = (uint8_t)die_of(w[1]);
        n_rounds[g] = 0;
        mem[g].clear();
        alive[g] = 1;
    };
    for (int g = 0; g < N; ++g) new_game(g, first_game_id + (uint32_t)g);
    RESERVE(ctx->a_states_in, sizeof(diee_bg_state) * (size_t)N);
    RESERVE(ctx->a_ids_in, 4 * (size_t)N);
    RESERVE(ctx->a_root_ids, 2 * (size_t)N * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_moves, 4 * (size_t)N * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_visits, 4 * (size_t)N * DIEE_MAX_MOVES);
    RESERVE(ctx->a_root_counts, 4 * (size_t)N);
    RESERVE(ctx->a_moves_in, 4 * (size_t)N);
    RESERVE(ctx->a_rolls_in, 2 * (size_t)N);
    std::vector<diee_bg_state> live((size_t)N);
    std::vector<uint32_t> ids((size_t)N);
    std::vector<int> idx_of((size_t)N);
    std::vector<uint16_t> r_ids((size_t)N * DIEE_MAX_MOVES);
    std::vector<uint32_t> r_moves((size_t)N * DIEE_MAX_MOVES), chosen((size_t)N);
    std::vector<float> r_visits((size_t)N * DIEE_MAX_MOVES);
    std::vector<int32_t> r_counts((size_t)N), status((size_t)N);
    std::vector<uint8_t> rolls((size_t)N * 2);
    int n_rec = 0, n_pi = 0, waves = 0;
    std::vector<int> cap_count((size_t)N);  // memories a round-capped game emits this pass (-1 = not capped)
    auto emit = [&](int g, bool relabel, int winner, int count) -> bool {
        for (int mi = 0; mi < count; ++mi) {
            const Mem &m = mem[g][(size_t)mi];
            if (n_rec >= rec_cap || n_pi + m.pi_n > pi_cap) return false;
            diee_traj_record &r = rec_out[n_rec++];
            r.state = m.st; r.game_id = gid[g]; r.ply = (uint16_t)m.ply;
            r.outcome = (int8_t)(relabel ? (winner == m.player ? 1 : (winner == -m.player ? -1 : 0)) : 0);
            r.pad = 0; r.n_pi = (uint16_t)m.pi_n; r.pad2 = 0; r.pi_offset = (uint32_t)n_pi;
            memcpy(pi_ids_out + n_pi, sp_ids.data() + m.pi_off, sizeof(uint16_t) * (size_t)m.pi_n);
            memcpy(pi_vals_out + n_pi, sp_vals.data() + m.pi_off, sizeof(float) * (size_t)m.pi_n);
            n_pi += m.pi_n;
        }
        return true;
    };
    for (;;) {  // while !states.is_empty() (:129)
        int nl = 0;
        for (int g = 0; g < N; ++g)
            if (alive[g]) { live[nl] = state[g]; ids[nl] = gid[g]; idx_of[nl] = g; ++nl; }
        if (nl == 0) break;
        if ((o.max_waves > 0 && waves >= o.max_waves) || (refill && o.target_games > 0 && rep.games_finished >= o.target_games)) {
            // time box: the games still running hand over what they have recorded so far, outcome 0 (not a reference
            // behaviour: a bounded sample of the same work for benchmarks)
            for (int g = 0; g < N; ++g)
                if (alive[g]) {
                    if (!emit(g, false, 0, (int)mem[g].size())) return fail(ctx, DIEE_ERR_OVERFLOW, "selfplay_run: record buffers too small");
                    rep.games_cut += 1;
                }
            break;
        }
        rep.game_moves += (uint64_t)nl;
        CU(cudaMemcpyAsync(ctx->a_states_in.p, live.data(), sizeof(diee_bg_state) * (size_t)nl, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->a_ids_in.p, ids.data(), 4 * (size_t)nl, cudaMemcpyHostToDevice, st));
        AlphaPool P;
        rc = alpha_search_device(ctx, net, (const diee_bg_state *)ctx->a_states_in.p, nl, (const uint32_t *)ctx->a_ids_in.p, cfg, seed,
                                 (uint32_t)waves, max_nodes, P, K, o.virtual_loss);  // fresh tree every game-move (:137)
        if (rc != DIEE_OK) return rc;
        ++waves;
        CU(launch_alpha_root_out(st, P, nl, (uint16_t *)ctx->a_root_ids.p, (uint32_t *)ctx->a_root_moves.p, (float *)ctx->a_root_visits.p,
                                 (int32_t *)ctx->a_root_counts.p));
        ctx->launches += 1;
        const size_t nm = (size_t)nl * DIEE_MAX_MOVES;
        CU(cudaMemcpyAsync(r_ids.data(), ctx->a_root_ids.p, 2 * nm, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(r_moves.data(), ctx->a_root_moves.p, 4 * nm, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(r_visits.data(), ctx->a_root_visits.p, 4 * nm, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(r_counts.data(), ctx->a_root_counts.p, 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(status.data(), ctx->a_status.p, 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int pi = 0; pi < nl; ++pi) {  // :171-223
            const int g = idx_of[pi];
            if (status[pi] != DIEE_OK) return fail(ctx, status[pi], "selfplay_run: search of game %d failed (node pool exhausted?)", g);
            const int nc = r_counts[pi];
            if (nc > DIEE_MAX_MOVES) return fail(ctx, DIEE_ERR_OVERFLOW, "selfplay_run: more than %d root children", DIEE_MAX_MOVES);
            const uint16_t *cid = &r_ids[(size_t)pi * DIEE_MAX_MOVES];
            const float *vis = &r_visits[(size_t)pi * DIEE_MAX_MOVES];
            float cpi[DIEE_MAX_MOVES];
            volatile float vsum = 0.f;
            for (int k = 0; k < nc; ++k) vsum = vsum + vis[k];  // sequential f32, child order (contract)
            for (int k = 0; k < nc; ++k) {
                volatile float p = vis[k] / vsum;
                cpi[k] = powf(p, tinv);  // pow_(1/T), no renormalisation (:164-166)
            }
            cap_count[pi] = -1;
            if (n_rounds[g] >= (int)cfg->simulate_round_limit) {  // :172-180, falls through (Q10)
                cap_count[pi] = (int)mem[g].size();  // emitted below, in the reference's per-game order
                alive[g] = 0;
            }
            uint32_t w[4];
            philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)n_rounds[g], gid[g], DIEE_STREAM_GAME, 0, w);
            rolls[2 * pi] = (uint8_t)die_of(w[0]);
            rolls[2 * pi + 1] = (uint8_t)die_of(w[1]);
            double dsum = 0.0;
            for (int k = 0; k < nc; ++k) dsum += (double)cpi[k];
            if (nc == 0 || !(dsum != 0.0)) {  // forced pass (:183-189)
                n_rounds[g] += 1;
                chosen[pi] = SEQ_EMPTY;
                continue;
            }
            // weighted_select_tensor_idx (alphazero.rs:129-137): f64 cumulative weights in action-id order
            double dense[DIEE_ACTION_SPACE];
            memset(dense, 0, sizeof dense);
            for (int k = 0; k < nc; ++k) dense[cid[k]] = (double)cpi[k];
            double total = 0.0;
            for (int j = 0; j < DIEE_ACTION_SPACE; ++j) total += dense[j];
            uint32_t sw[4];
            philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)n_rounds[g], gid[g], DIEE_STREAM_SAMPLE, 0, sw);
            const double pick = unit53(sw[0], sw[1]) * total;
            double cum = 0.0;
            int action = -1, lastnz = -1;
            for (int j = 0; j < DIEE_ACTION_SPACE && action < 0; ++j) {
                if (dense[j] == 0.0) continue;
                cum += dense[j];
                lastnz = j;
                if (cum > pick) action = j;
            }
            if (action < 0) action = lastnz;
            Mem m;  // MemoryFragment (:195-199)
            m.st = state[g]; m.player = state[g].player; m.pi_off = (int)sp_ids.size(); m.pi_n = nc; m.ply = n_rounds[g];
            sp_ids.insert(sp_ids.end(), cid, cid + nc);
            sp_vals.insert(sp_vals.end(), cpi, cpi + nc);
            mem[g].push_back(m);
            int kc = -1;  // decode(selected) must be one of the legal moves (:202-209)
            for (int k = 0; k < nc; ++k)
                if (cid[k] == (uint16_t)action) { kc = k; break; }
            if (kc < 0) return fail(ctx, DIEE_ERR_INVALID, "selfplay_run: sampled action %d is not a root child", action);
            chosen[pi] = r_moves[(size_t)pi * DIEE_MAX_MOVES + kc];
            n_rounds[g] += 1;
        }
        // apply_move / skip_turn for every live game on the device (:186,:210), then read the states back
        CU(cudaMemcpyAsync(ctx->a_moves_in.p, chosen.data(), 4 * (size_t)nl, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->a_rolls_in.p, rolls.data(), 2 * (size_t)nl, cudaMemcpyHostToDevice, st));
        CU(launch_bg_apply(st, (diee_bg_state *)ctx->a_states_in.p, (const diee_move *)ctx->a_moves_in.p, (const uint8_t *)ctx->a_rolls_in.p, nl));
        ctx->launches += 1;
        CU(cudaMemcpyAsync(live.data(), ctx->a_states_in.p, sizeof(diee_bg_state) * (size_t)nl, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int pi = 0; pi < nl; ++pi) {
            const int g = idx_of[pi];
            state[g] = live[pi];
            if (cap_count[pi] >= 0) {
                if (!emit(g, false, 0, cap_count[pi])) return fail(ctx, DIEE_ERR_OVERFLOW, "selfplay_run: record buffers too small");
                rep.games_finished += 1;  // a game = played to a winner or to the round cap
            }
            // (the pass branch `continue`s before the winner test; a capped game still plays this move and, if it wins
            // with it, is emitted a second time -- quirk Q10, kept in the reference mode only)
            if (chosen[pi] != SEQ_EMPTY && (alive[g] || !refill)) {
                const int win = state[g].off[0] == 15 ? -1 : (state[g].off[1] == 15 ? 1 : 0);
                if (win != 0) {  // :215-223
                    if (!emit(g, true, win, (int)mem[g].size())) return fail(ctx, DIEE_ERR_OVERFLOW, "selfplay_run: record buffers too small");
                    if (alive[g]) rep.games_finished += 1;
                    alive[g] = 0;
                }
            }
            // non-parity: the slot of a finished game starts the next game, so the forward batch stays at n_games
            if (refill && !alive[g] && !(o.target_games > 0 && rep.games_finished >= o.target_games)) new_game(g, next_gid++);
        }
    }
    *n_rec_out = n_rec; *n_pi_out = n_pi;
    if (n_waves_out) *n_waves_out = waves;
    rep.waves = waves;
    if (report_out) *report_out = rep;
    return DIEE_OK;
}

}  // extern "C"
