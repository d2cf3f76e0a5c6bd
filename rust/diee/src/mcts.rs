//! Pure MCTS: `MctsConfig` (`src/lib.rs:33-52`) and `mct_search` (`src/mcts/simple_mcts.rs:10-39`).
//!
//! The reference calls `mct_search` once per game from a rayon `par_iter` (`src/versus.rs:303-306`); here the whole
//! slice is one `diee_mcts_search` call (`mct_search_batch`), and `mct_search` is the one-game form of it.
use crate::backgammon::{Actions, Backgammon};
use crate::ctx::{Ctx, DieeError};
use crate::tictactoe::TicTacToe;
use diee_sys as sys;
use std::ptr::null_mut;

#[derive(Debug, Clone)]
pub struct MctsConfig {
    pub iterations: usize,
    pub c: f32,
    pub simulate_round_limit: usize,
    pub dirichlet_alpha: f32,
    pub dirichlet_epsilon: f32,
    /// `DIEE_MODE_*`: 0 = reference-exact (incl. the panic on a no-move node, which surfaces as `NoMovesPanic`)
    pub mode_flags: u32,
}

impl From<&MctsConfig> for sys::diee_mcts_cfg {
    fn from(c: &MctsConfig) -> Self {
        sys::diee_mcts_cfg {
            iterations: c.iterations as u32,
            c: c.c,
            simulate_round_limit: c.simulate_round_limit as u32,
            dirichlet_alpha: c.dirichlet_alpha,
            dirichlet_epsilon: c.dirichlet_epsilon,
            mode_flags: c.mode_flags,
        }
    }
}

/// Where a game sits in the injected random stream: `(seed, first game id of the slice, epoch = search number)`.
#[derive(Debug, Clone, Copy)]
pub struct StreamKey {
    pub seed: u64,
    pub first_game_id: u32,
    pub epoch: u32,
}

/// `mct_search` for a slice of backgammon games: one launch for all of them.  `players[i]` is the player the value is
/// counted for (`versus.rs:305` passes `game.get_player()`).  A game whose search hits the reference's panic
/// (`node.rs:119-121`) panics here too unless `DIEE_MODE_PASS_CHILD` is set.
pub fn mct_search_batch(ctx: &Ctx, games: &[Backgammon], cfg: &MctsConfig, key: StreamKey) -> Result<Vec<Actions>, DieeError> {
    let states: Vec<sys::diee_bg_state> = games.iter().map(Into::into).collect();
    let players: Vec<i8> = games.iter().map(|g| g.player).collect();
    let none = sys::diee_move { from1: sys::DIEE_NONE, to1: sys::DIEE_NONE, from2: sys::DIEE_NONE, to2: sys::DIEE_NONE };
    let mut best = vec![none; games.len()];
    let mut status = vec![0i32; games.len()];
    let c: sys::diee_mcts_cfg = cfg.into();
    ctx.check(unsafe {
        sys::diee_mcts_search(ctx.raw, sys::DIEE_GAME_BACKGAMMON, states.as_ptr().cast(), games.len() as i32, players.as_ptr(), &c,
                              key.seed, key.first_game_id, key.epoch, best.as_mut_ptr().cast(), status.as_mut_ptr(), null_mut(),
                              null_mut(), null_mut(), null_mut(), null_mut())
    })?;
    for (i, &st) in status.iter().enumerate() {
        if st == sys::DIEE_ERR_NO_MOVES_PANIC {
            panic!("expand() called on node with no expandable moves (game {})", i); // node.rs:119-121
        }
        if st != sys::DIEE_OK {
            return Err(DieeError { code: st, message: format!("search of game {} failed", i) });
        }
    }
    Ok(best.iter().map(Backgammon::move_to_actions).collect())
}

/// `mct_search(state, player, &cfg)` (`simple_mcts.rs:10`) for one backgammon game.
pub fn mct_search(ctx: &Ctx, state: Backgammon, player: i8, cfg: &MctsConfig, seed: u64, epoch: u32) -> Actions {
    let mut g = state;
    g.player = state.player; // the value is counted for `player`; the reference always passes state.get_player()
    debug_assert_eq!(player, state.player);
    let key = StreamKey { seed, first_game_id: state.id as u32, epoch };
    mct_search_batch(ctx, &[g], cfg, key).expect("diee_mcts_search").remove(0)
}

/// `mct_search` for tic-tac-toe games (BASELINE configs[0]); a move is the cell index, 10 = EMPTY_MOVE.
pub fn mct_search_ttt_batch(ctx: &Ctx, games: &[TicTacToe], cfg: &MctsConfig, key: StreamKey) -> Result<Vec<u8>, DieeError> {
    let states: Vec<sys::diee_ttt_state> = games.iter().map(Into::into).collect();
    let players: Vec<i8> = games.iter().map(|g| g.player).collect();
    let mut best = vec![10u8; games.len()];
    let mut status = vec![0i32; games.len()];
    let c: sys::diee_mcts_cfg = cfg.into();
    ctx.check(unsafe {
        sys::diee_mcts_search(ctx.raw, sys::DIEE_GAME_TICTACTOE, states.as_ptr().cast(), games.len() as i32, players.as_ptr(), &c,
                              key.seed, key.first_game_id, key.epoch, best.as_mut_ptr().cast(), status.as_mut_ptr(), null_mut(),
                              null_mut(), null_mut(), null_mut(), null_mut())
    })?;
    Ok(best)
}
