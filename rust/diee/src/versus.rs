//! The arena (`src/versus.rs:160-318`): `play` runs 400 concurrent games, each round partitions them by side to move and
//! asks each side's agent for every game's action in ONE batched call (`get_actions_for_player`, `:270-318`).
use crate::alphazero::{alpha_mcts_parallel, AlphaZero, ResNet};
use crate::backgammon::{Actions, Backgammon};
use crate::base::LearnableGame;
use crate::ctx::{index_of, philox, Ctx, DieeError};
use crate::mcts::{mct_search_batch, MctsConfig, StreamKey};
use diee_sys as sys;

#[derive(Debug, Clone, Copy, PartialEq)]
pub enum Agent { Random, Mcts, Model }

pub struct Player<'a> {
    pub player_type: Agent,
    pub model: Option<ResNet<'a>>,
}

#[derive(Debug, Default, Clone)]
pub struct PlayResult {
    pub player1_wins: usize,
    pub player2_wins: usize,
    pub draws: usize,
    pub n_games: usize,
}

pub const N_GAMES: usize = 400; // versus.rs:168
pub const ROUND_LIMIT: usize = 400; // versus.rs:169

/// `get_actions_for_player` (`versus.rs:270-318`): the actions of one side for all of its games, in one call per agent kind.
pub fn get_actions_for_player(ctx: &Ctx, player: &Player, games: &[Backgammon], cfg: &MctsConfig, temp: f64, seed: u64, round: u32)
                              -> Result<Vec<Actions>, DieeError> {
    match player.player_type {
        Agent::Mcts => mct_search_batch(ctx, games, cfg, StreamKey { seed, first_game_id: 0, epoch: round }), // :303-306
        Agent::Random => {
            // valid_moves.choose() (:307-316): word 2 of the game's GAME-stream block for this round
            Ok(games.iter().map(|g| {
                let moves = g.get_valid_moves();
                if moves.is_empty() { return Backgammon::EMPTY_MOVE; }
                let w = philox(seed, round, g.id as u32, sys::DIEE_STREAM_GAME as u32, 0);
                moves[index_of(w[2], moves.len() as u32) as usize].clone()
            }).collect())
        }
        Agent::Model => {
            // alpha_mcts_parallel + pi^(1/T) + weighted sample (:276-302)
            let net = player.model.as_ref().expect("Agent::Model needs a model");
            let roots = alpha_mcts_parallel(ctx, games, net, cfg, seed, round)?;
            Ok(roots.iter().zip(games).map(|(r, g)| {
                let sum: f32 = r.visits.iter().sum();
                if r.moves.is_empty() || sum == 0.0 { return Backgammon::EMPTY_MOVE; }
                let mut pi = vec![0f32; sys::DIEE_ACTION_SPACE];
                for (id, v) in r.action_ids.iter().zip(&r.visits) {
                    pi[*id as usize] = (v / sum).powf((1.0 / temp) as f32);
                }
                let a = AlphaZero::weighted_select_idx(&pi, seed, g.id as u32, round);
                let k = r.action_ids.iter().position(|&id| id as usize == a).expect("sampled action is a root child");
                r.moves[k].clone()
            }).collect())
        }
    }
}

/// `play::<Backgammon>(player1, player2, &cfg, temp)` (`versus.rs:160-268`).  Player 1 moves first in the first half of
/// the games, player 2 in the second half (`:172-174`).
pub fn play(ctx: &Ctx, player1: &Player, player2: &Player, cfg: &MctsConfig, temp: f64, seed: u64) -> Result<PlayResult, DieeError> {
    let mut games: Vec<(usize, Backgammon, usize)> = (0..N_GAMES).map(|i| {
        let mut g = Backgammon::new();
        g.set_id(i);
        g.seed = seed;
        if i >= N_GAMES / 2 { g.skip_turn(); } // :172-174
        g.roll_die();
        (i, g, 0usize)
    }).collect();
    let mut res = PlayResult { n_games: N_GAMES, ..Default::default() };
    let mut round = 0u32;
    while !games.is_empty() {
        // partition by the side to move (:195-196); player 1 is the -1 side in games 0..200 and the +1 side after
        let side = |&(i, g, _): &(usize, Backgammon, usize)| (g.player == -1) == (i < N_GAMES / 2);
        let (p1, p2): (Vec<_>, Vec<_>) = games.iter().cloned().partition(side);
        let mut next = Vec::with_capacity(games.len());
        for (player, part) in [(player1, p1), (player2, p2)] {
            if part.is_empty() { continue; }
            let states: Vec<Backgammon> = part.iter().map(|t| t.1).collect();
            let actions = get_actions_for_player(ctx, player, &states, cfg, temp, seed, round)?;
            for ((i, mut g, rounds), a) in part.into_iter().zip(actions) {
                if a == Backgammon::EMPTY_MOVE { g.skip_turn(); } else { g.apply_move(&a); } // :223-229
                match g.check_winner() {
                    Some(w) => { if (w == -1) == (i < N_GAMES / 2) { res.player1_wins += 1 } else { res.player2_wins += 1 } }
                    None if rounds + 1 >= ROUND_LIMIT => res.draws += 1, // :240-248
                    None => next.push((i, g, rounds + 1)),
                }
            }
        }
        games = next;
        round += 1;
    }
    Ok(res)
}
