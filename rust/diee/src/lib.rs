//! `diee` -- die-e's Game/agent API on the B200 engine (libdiee_cuda.so through `diee-sys`).
//!
//! Module map = the reference's (`src/lib.rs:6-11`): `base` (trait `LearnableGame`), `backgammon`, `tictactoe`,
//! `mcts` (`MctsConfig`, `mct_search`), `alphazero` (`ResNet`, `alpha_mcts_parallel`, `AlphaZero::self_play_parallel`,
//! `MemoryFragment`), `versus` (`play`).  UNVERIFIED by rustc (no toolchain where this was written).
pub mod alphazero;
pub mod backgammon;
pub mod base;
pub mod ctx;
pub mod mcts;
pub mod tictactoe;
pub mod versus;

pub use ctx::{Ctx, DieeError};
pub use mcts::MctsConfig;
