//! The AlphaZero side of the path: `ResNet` (`src/alphazero/nnet.rs:57-155`, inference only), `alpha_mcts_parallel`
//! (`src/mcts/alpha_mcts.rs:91-202`), `AlphaZero::self_play_parallel` (`src/alphazero/alpha_parallel.rs:101-231`) and
//! `MemoryFragment` (`src/alphazero/alphazero.rs:69-73`).  Training (`alphazero.rs:202-261`) stays in tch.
use crate::backgammon::{Actions, Backgammon};
use crate::base::LearnableGame;
use crate::ctx::{Ctx, DieeError};
use crate::mcts::MctsConfig;
use diee_sys as sys;
use std::ptr::null_mut;

pub struct ResNet<'a> {
    ctx: &'a Ctx,
    pub(crate) raw: *mut sys::diee_net,
}

impl<'a> ResNet<'a> {
    /// The 22 + 12 * blocks tensors in the reference's registration order (`nnet.rs:62-98`, `ResBlock::new :37-44`), as
    /// host f32 slices -- what `VarStore::variables()` holds after `vs.load(path)` (`nnet.rs:109-118`); BatchNorm is
    /// folded at load time.  (With the `tch` feature: `ResNet::from_path`.)
    pub fn from_tensors(ctx: &'a Ctx, tensors: &[&[f32]]) -> Result<Self, DieeError> {
        let ptrs: Vec<*const f32> = tensors.iter().map(|t| t.as_ptr()).collect();
        let numels: Vec<i64> = tensors.iter().map(|t| t.len() as i64).collect();
        let mut raw: *mut sys::diee_net = null_mut();
        ctx.check(unsafe { sys::diee_net_create(ctx.raw, sys::DIEE_GAME_BACKGAMMON, ptrs.as_ptr(), numels.as_ptr(), ptrs.len() as i32, &mut raw) })?;
        Ok(ResNet { ctx, raw })
    }

    /// `DIEE_NET_SPLIT3` (default: tensor cores, inside the fp32 tolerance), `DIEE_NET_FP32`, or `DIEE_NET_BF16` (fast, not parity)
    pub fn set_precision(&self, precision: i32) -> Result<(), DieeError> {
        self.ctx.check(unsafe { sys::diee_net_set_precision(self.ctx.raw, self.raw, precision) })
    }

    /// `forward_t` (`nnet.rs:120-133`): (softmaxed policy `[n][1352]`, tanh value `[n]`); `as_tensor` is fused in
    pub fn forward_t(&self, states: &[Backgammon]) -> Result<(Vec<f32>, Vec<f32>), DieeError> {
        let s: Vec<sys::diee_bg_state> = states.iter().map(Into::into).collect();
        let mut policy = vec![0f32; s.len() * sys::DIEE_ACTION_SPACE];
        let mut value = vec![0f32; s.len()];
        self.ctx.check(unsafe { sys::diee_net_forward(self.ctx.raw, self.raw, s.as_ptr(), s.len() as i32, policy.as_mut_ptr(), value.as_mut_ptr()) })?;
        Ok((policy, value))
    }

    /// `forward_policy` (`nnet.rs:150-155`)
    pub fn forward_policy(&self, states: &[Backgammon]) -> Result<Vec<f32>, DieeError> {
        Ok(self.forward_t(states)?.0)
    }
}

impl Drop for ResNet<'_> {
    fn drop(&mut self) {
        unsafe { sys::diee_net_destroy(self.ctx.raw, self.raw) };
    }
}

/// What the callers of `alpha_mcts_parallel` read from the store: each root's children `{action_taken, visits}`
/// (`src/mcts/utils.rs:42-58`), in legal-move order.  The node store itself stays in HBM.
#[derive(Debug, Clone)]
pub struct RootChildren {
    pub action_ids: Vec<u16>,
    pub moves: Vec<Actions>,
    pub visits: Vec<f32>,
}

/// `alpha_mcts_parallel(store, states, net, cfg, pb)` (`alpha_mcts.rs:91`): lock-step search over all states, the net
/// evaluated once per iteration for the whole batch; `epoch` = the wave number (keys the shared Dirichlet vector).
pub fn alpha_mcts_parallel(ctx: &Ctx, states: &[Backgammon], net: &ResNet, cfg: &MctsConfig, seed: u64, epoch: u32)
                           -> Result<Vec<RootChildren>, DieeError> {
    let n = states.len();
    let s: Vec<sys::diee_bg_state> = states.iter().map(Into::into).collect();
    let ids: Vec<u32> = states.iter().map(|g| g.id as u32).collect();
    let none = sys::diee_move { from1: sys::DIEE_NONE, to1: sys::DIEE_NONE, from2: sys::DIEE_NONE, to2: sys::DIEE_NONE };
    let mut r_ids = vec![0u16; n * sys::DIEE_MAX_MOVES];
    let mut r_moves = vec![none; n * sys::DIEE_MAX_MOVES];
    let mut r_vis = vec![0f32; n * sys::DIEE_MAX_MOVES];
    let mut r_cnt = vec![0i32; n];
    let mut status = vec![0i32; n];
    let c: sys::diee_mcts_cfg = cfg.into();
    ctx.check(unsafe {
        sys::diee_alpha_search(ctx.raw, net.raw, s.as_ptr(), n as i32, ids.as_ptr(), &c, seed, epoch, 0, r_ids.as_mut_ptr(),
                               r_moves.as_mut_ptr(), r_vis.as_mut_ptr(), r_cnt.as_mut_ptr(), status.as_mut_ptr(), null_mut(), null_mut())
    })?;
    let mut out = Vec::with_capacity(n);
    for g in 0..n {
        if status[g] != sys::DIEE_OK {
            return Err(DieeError { code: status[g], message: format!("search of game {} failed (node pool exhausted?)", g) });
        }
        let (b, k) = (g * sys::DIEE_MAX_MOVES, r_cnt[g] as usize);
        out.push(RootChildren {
            action_ids: r_ids[b..b + k].to_vec(),
            moves: r_moves[b..b + k].iter().map(Backgammon::move_to_actions).collect(),
            visits: r_vis[b..b + k].to_vec(),
        });
    }
    Ok(out)
}

/// `MemoryFragment` (`alphazero.rs:69-73`) with `ps` dense (`[1352]`) and `state` = the `[6,4,6]` planes
#[derive(Debug, Clone)]
pub struct MemoryFragment {
    pub outcome: i8,
    pub ps: Vec<f32>,
    pub state: Vec<f32>,
}

pub struct AlphaZeroConfig {
    pub temperature: f64,
    pub num_self_play_batches: usize,
}

pub struct AlphaZero<'a> {
    pub ctx: &'a Ctx,
    pub model: ResNet<'a>,
    pub config: AlphaZeroConfig,
    pub mcts_config: MctsConfig,
    pub seed: u64,
}

impl<'a> AlphaZero<'a> {
    /// `self_play_parallel` (`alpha_parallel.rs:101-231`): `num_self_play_batches` games in lock-step to a winner or the
    /// round limit; records come back in the reference's emission order and are expanded to `MemoryFragment`s here.
    pub fn self_play_parallel(&self, first_game_id: u32) -> Result<Vec<MemoryFragment>, DieeError> {
        let n = self.config.num_self_play_batches;
        let limit = self.mcts_config.simulate_round_limit;
        let rec_cap = n * (2 * limit + 4);
        let pi_cap = rec_cap * 48;
        let zero_state = sys::diee_bg_state { pts: [0; 24], bar: [0; 2], off: [0; 2], roll: [0; 2], player: 0, second: 0 };
        let zero = sys::diee_traj_record { state: zero_state, game_id: 0, ply: 0, outcome: 0, pad: 0, n_pi: 0, pad2: 0, pi_offset: 0 };
        let mut rec = vec![zero; rec_cap];
        let mut pi_ids = vec![0u16; pi_cap];
        let mut pi_vals = vec![0f32; pi_cap];
        let (mut n_rec, mut n_pi, mut n_waves) = (0i32, 0i32, 0i32);
        let c: sys::diee_mcts_cfg = (&self.mcts_config).into();
        self.ctx.check(unsafe {
            sys::diee_selfplay_run(self.ctx.raw, self.model.raw, n as i32, &c, self.config.temperature as f32, self.seed, first_game_id, 0,
                                   rec.as_mut_ptr(), rec_cap as i32, pi_ids.as_mut_ptr(), pi_vals.as_mut_ptr(), pi_cap as i32,
                                   &mut n_rec, &mut n_pi, &mut n_waves)
        })?;
        rec.truncate(n_rec as usize);
        // as_tensor of every recorded state in one call
        let states: Vec<sys::diee_bg_state> = rec.iter().map(|r| r.state).collect();
        let mut planes = vec![0f32; states.len() * 144];
        self.ctx.check(unsafe { sys::diee_bg_encode_states(self.ctx.raw, states.as_ptr(), states.len() as i32, planes.as_mut_ptr()) })?;
        Ok(rec.iter().enumerate().map(|(i, r)| {
            let mut ps = vec![0f32; sys::DIEE_ACTION_SPACE];
            let (o, k) = (r.pi_offset as usize, r.n_pi as usize);
            for j in 0..k {
                ps[pi_ids[o + j] as usize] = pi_vals[o + j];
            }
            MemoryFragment { outcome: r.outcome, ps, state: planes[i * 144..(i + 1) * 144].to_vec() }
        }).collect())
    }

    /// `weighted_select_tensor_idx` (`alphazero.rs:129-137`) on the injected SAMPLE stream: first index whose f64
    /// cumulative weight exceeds `u * sum`
    pub fn weighted_select_idx(pi: &[f32], seed: u64, game_id: u32, ply: u32) -> usize {
        let w = crate::ctx::philox(seed, ply, game_id, sys::DIEE_STREAM_SAMPLE as u32, 0);
        let m = ((((w[0] as u64) << 21) ^ ((w[1] as u64) >> 11)) & ((1u64 << 53) - 1)) as f64 / 9007199254740992.0;
        let total: f64 = pi.iter().map(|&p| p as f64).sum();
        let pick = m * total;
        let (mut cum, mut last) = (0.0f64, 0usize);
        for (j, &p) in pi.iter().enumerate() {
            if p == 0.0 { continue; }
            cum += p as f64;
            last = j;
            if cum > pick { return j; }
        }
        last
    }
}

/// silence "unused" for the trait import when the crate is built without callers of the generic helpers
#[allow(dead_code)]
fn _assert_game<T: LearnableGame>() {}
