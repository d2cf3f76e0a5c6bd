//! `trait LearnableGame` -- the reference's game API (`src/base.rs:8-51`), kept method for method.
//! `as_tensor` returns the `[C, H, W]` planes as a flat `Vec<f32>` (row-major, the layout of the reference's
//! `[1, C, H, W]` tensor); with the `tch` feature `as_tch_tensor` wraps it.
use serde::{de::DeserializeOwned, Serialize};
use std::fmt::Debug;

pub trait LearnableGame: Clone + Debug + DeserializeOwned + Serialize + Send + Sync + Copy {
    type Move: Clone + Debug + DeserializeOwned + Serialize + Send + Sync + PartialEq;
    const EMPTY_MOVE: Self::Move;
    const IS_DETERMINISTIC: bool;
    const ACTION_SPACE_SIZE: i64;
    const N_INPUT_CHANNELS: i64;
    const CONV_OUTPUT_SIZE: i64;
    const N_FILTERS: i64;
    const N_RES_BLOCKS: i64;

    fn new() -> Self;
    fn name() -> String;
    fn get_valid_moves(&self) -> Vec<Self::Move>;
    fn apply_move(&mut self, action: &Self::Move);
    fn roll_die(&mut self) -> (u8, u8) {
        if Self::IS_DETERMINISTIC {
            panic!("roll_die called on deterministic game!")
        }
        unimplemented!("You should implement roll_die for non-deterministic games!")
    }
    fn skip_turn(&mut self);
    fn get_player(&self) -> i8;
    fn check_winner(&self) -> Option<i8>;
    fn as_tensor(&self) -> Vec<f32>;
    fn decode(&self, action: u32) -> Self::Move;
    fn encode(&self, action: &Self::Move) -> u32;
    fn get_id(&self) -> usize;
    fn set_id(&mut self, new_id: usize);
    fn to_pretty_str(&self) -> String;

    #[cfg(feature = "tch")]
    fn as_tch_tensor(&self) -> tch::Tensor {
        let (c, s) = (Self::N_INPUT_CHANNELS, Self::CONV_OUTPUT_SIZE);
        // backgammon: [1, 6, 4, 6]; tic-tac-toe: [1, 3, 3, 3]
        let (h, w) = if s == 24 { (4, 6) } else { (3, 3) };
        tch::Tensor::from_slice(&self.as_tensor()).view([1, c, h, w])
    }
}
