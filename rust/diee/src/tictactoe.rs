//! `TicTacToe` (`src/tictactoe/mod.rs:6-117`).  The env itself is nine cells, so the trait methods are plain host code
//! (the same few lines as the reference's); what runs on the GPU is the SEARCH over it (`mcts::mct_search`,
//! `DIEE_GAME_TICTACTOE`), which evaluates the same rules in `csrc/mcts_kernels.cu::TttGame`.
use crate::base::LearnableGame;
use diee_sys as sys;
use serde::{Deserialize, Serialize};

#[derive(Clone, Copy, Serialize, Deserialize, Debug)]
pub struct TicTacToe {
    pub(crate) player: i8,
    pub board: [i8; 9],
    pub(crate) id: usize,
}

impl From<&TicTacToe> for sys::diee_ttt_state {
    fn from(t: &TicTacToe) -> Self {
        sys::diee_ttt_state { board: t.board, player: t.player, pad: [0; 6] }
    }
}

impl LearnableGame for TicTacToe {
    type Move = u8;
    const EMPTY_MOVE: Self::Move = 10;
    const ACTION_SPACE_SIZE: i64 = 9;
    const CONV_OUTPUT_SIZE: i64 = 9;
    const N_INPUT_CHANNELS: i64 = 3;
    const N_FILTERS: i64 = 64;
    const N_RES_BLOCKS: i64 = 4;
    const IS_DETERMINISTIC: bool = true;

    fn new() -> Self {
        TicTacToe { player: -1, board: [0; 9], id: 0 }
    }

    fn name() -> String {
        String::from("tictactoe")
    }

    fn get_valid_moves(&self) -> Vec<Self::Move> {
        (0u8..9).filter(|&i| self.board[i as usize] == 0).collect() // empty cells ascending (:36-44)
    }

    fn apply_move(&mut self, action: &Self::Move) {
        self.board[*action as usize] = self.player;
        self.player *= -1
    }

    fn skip_turn(&mut self) {
        self.player *= -1
    }

    fn get_player(&self) -> i8 {
        self.player
    }

    fn check_winner(&self) -> Option<i8> {
        // rows, columns, diagonals; a full board without a line is a draw = Some(0) (:59-79)
        const LINES: [[usize; 3]; 8] = [[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8], [0, 4, 8], [2, 4, 6]];
        for l in LINES.iter() {
            let v = self.board[l[0]];
            if v != 0 && v == self.board[l[1]] && v == self.board[l[2]] {
                return Some(v);
            }
        }
        if self.board.iter().all(|&v| v != 0) { Some(0) } else { None }
    }

    fn as_tensor(&self) -> Vec<f32> {
        // [1, 3, 3, 3]: planes (== -1, == 0, == 1) (:81-92)
        let mut out = vec![0f32; 27];
        for (plane, want) in [-1i8, 0, 1].iter().enumerate() {
            for i in 0..9 {
                out[plane * 9 + i] = (self.board[i] == *want) as u8 as f32;
            }
        }
        out
    }

    fn decode(&self, action: u32) -> Self::Move {
        action as u8
    }

    fn encode(&self, action: &Self::Move) -> u32 {
        *action as u32
    }

    fn get_id(&self) -> usize {
        self.id
    }

    fn set_id(&mut self, new_id: usize) {
        self.id = new_id
    }

    fn to_pretty_str(&self) -> String {
        let c = |v: i8| match v { -1 => 'X', 1 => 'O', _ => '.' };
        (0..3).map(|r| (0..3).map(|k| c(self.board[r * 3 + k])).collect::<String>()).collect::<Vec<_>>().join("\n")
    }
}
