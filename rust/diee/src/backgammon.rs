//! `Backgammon` (`src/backgammon/backgammon_logic.rs:54-60`) over the engine's env kernels.
//!
//! The struct keeps the reference's fields and serde shape (`{"board":[[24 x i8],[h0,h1],[c0,c1]],"roll":[a,b],
//! "player":+-1,"is_second_play":bool,"id":n}`), so saved games stay readable both ways.  Every rule is evaluated by
//! the library (bit-exact with the reference's `get_valid_moves` order, `tests/test_gpu_env.py`).  Dice come from the
//! game's own slot of the injected stream: `(seed, id)` key it and `draws` counts the rolls made so far.
use crate::base::LearnableGame;
use crate::ctx::{die_of, philox, Ctx};
use diee_sys as sys;
use serde::{Deserialize, Serialize};
use std::cell::RefCell;

pub type Board = ([i8; 24], (u8, u8), (u8, u8));
pub type Actions = Vec<(i8, i8)>;

thread_local! {
    /// the context single-state trait calls run on (the batched entry points take a `&Ctx` explicitly)
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

fn with_ctx<R>(f: impl FnOnce(&Ctx) -> R) -> R {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            *c = Some(Ctx::new(0).expect("no usable CUDA device (there is no CPU fallback)"));
        }
        f(c.as_ref().unwrap())
    })
}

#[derive(Debug, Clone, Copy, Serialize, Deserialize, PartialEq)]
pub struct Backgammon {
    pub board: Board,
    pub roll: (u8, u8),
    pub player: i8,
    pub is_second_play: bool,
    pub id: usize,
    #[serde(default)]
    pub seed: u64,
    #[serde(default)]
    pub draws: u32,
}

impl From<&Backgammon> for sys::diee_bg_state {
    fn from(b: &Backgammon) -> Self {
        sys::diee_bg_state {
            pts: b.board.0,
            bar: [b.board.1 .0, b.board.1 .1],
            off: [b.board.2 .0, b.board.2 .1],
            roll: [b.roll.0, b.roll.1],
            player: b.player,
            second: b.is_second_play as u8,
        }
    }
}

impl Backgammon {
    pub(crate) fn absorb(&mut self, s: &sys::diee_bg_state) {
        self.board = (s.pts, (s.bar[0], s.bar[1]), (s.off[0], s.off[1]));
        self.roll = (s.roll[0], s.roll[1]);
        self.player = s.player;
        self.is_second_play = s.second != 0;
    }

    fn next_roll(&mut self) -> (u8, u8) {
        // stream GAME, c0 = draw index (include/diee.h); the very first roll of a game is stream INIT
        let w = if self.draws == 0 {
            philox(self.seed, 0, self.id as u32, sys::DIEE_STREAM_INIT as u32, 0)
        } else {
            philox(self.seed, self.draws - 1, self.id as u32, sys::DIEE_STREAM_GAME as u32, 0)
        };
        self.draws += 1;
        (die_of(w[0]), die_of(w[1]))
    }

    pub fn actions_to_move(a: &Actions) -> sys::diee_move {
        let get = |i: usize| a.get(i).copied().unwrap_or((sys::DIEE_NONE, sys::DIEE_NONE));
        let (f1, t1) = get(0);
        let (f2, t2) = get(1);
        sys::diee_move { from1: f1, to1: t1, from2: f2, to2: t2 }
    }

    pub fn move_to_actions(m: &sys::diee_move) -> Actions {
        let mut v = Vec::with_capacity(2);
        if m.from1 != sys::DIEE_NONE || m.to1 != sys::DIEE_NONE {
            v.push((m.from1, m.to1));
        }
        if m.from2 != sys::DIEE_NONE || m.to2 != sys::DIEE_NONE {
            v.push((m.from2, m.to2));
        }
        v
    }

    fn apply_with(&mut self, mv: sys::diee_move) {
        let (d0, d1) = {
            // apply_move rolls only when the turn passes (backgammon_logic.rs:179-185); the kernel ignores the roll otherwise
            let passes = !(self.roll.0 == self.roll.1 && !self.is_second_play) || Self::move_to_actions(&mv).is_empty();
            if passes { self.next_roll() } else { (0, 0) }
        };
        let mut s: sys::diee_bg_state = (&*self).into();
        let rolls = [d0, d1];
        with_ctx(|c| c.check(unsafe { sys::diee_bg_apply_moves(c.raw, &mut s, &mv, rolls.as_ptr(), 1) }).expect("diee_bg_apply_moves"));
        self.absorb(&s);
    }
}

impl LearnableGame for Backgammon {
    type Move = Actions;
    const EMPTY_MOVE: Self::Move = vec![];
    const IS_DETERMINISTIC: bool = false;
    const ACTION_SPACE_SIZE: i64 = 1352;
    const N_INPUT_CHANNELS: i64 = 6;
    const CONV_OUTPUT_SIZE: i64 = 24;
    const N_FILTERS: i64 = 256;
    const N_RES_BLOCKS: i64 = 19;

    fn new() -> Self {
        // backgammon_logic.rs:80-94
        let pts: [i8; 24] = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2];
        Backgammon { board: (pts, (0, 0), (0, 0)), roll: (0, 0), player: -1, is_second_play: false, id: 0, seed: 0, draws: 0 }
    }

    fn name() -> String {
        "backgammon".to_string()
    }

    fn get_valid_moves(&self) -> Vec<Self::Move> {
        assert!(self.roll != (0, 0), "die has not been rolled!"); // backgammon_logic.rs:404
        let s: sys::diee_bg_state = self.into();
        let mut moves = vec![sys::diee_move { from1: 0, to1: 0, from2: 0, to2: 0 }; sys::DIEE_MAX_MOVES];
        let mut count = 0i32;
        with_ctx(|c| {
            c.check(unsafe { sys::diee_bg_valid_moves(c.raw, &s, 1, moves.as_mut_ptr(), &mut count, std::ptr::null_mut()) })
                .expect("diee_bg_valid_moves")
        });
        moves[..count as usize].iter().map(Self::move_to_actions).collect()
    }

    fn apply_move(&mut self, action: &Self::Move) {
        self.apply_with(Self::actions_to_move(action));
    }

    fn roll_die(&mut self) -> (u8, u8) {
        self.roll = self.next_roll();
        self.roll
    }

    fn skip_turn(&mut self) {
        self.apply_with(sys::diee_move { from1: sys::DIEE_NONE, to1: sys::DIEE_NONE, from2: sys::DIEE_NONE, to2: sys::DIEE_NONE });
    }

    fn get_player(&self) -> i8 {
        self.player
    }

    fn check_winner(&self) -> Option<i8> {
        // backgammon_logic.rs:527-534: no gammon / backgammon
        if self.board.2 .0 == 15 {
            Some(-1)
        } else if self.board.2 .1 == 15 {
            Some(1)
        } else {
            None
        }
    }

    fn as_tensor(&self) -> Vec<f32> {
        assert!(self.roll != (0, 0)); // backgammon_logic.rs:199
        let s: sys::diee_bg_state = self.into();
        let mut out = vec![0f32; 144];
        with_ctx(|c| c.check(unsafe { sys::diee_bg_encode_states(c.raw, &s, 1, out.as_mut_ptr()) }).expect("diee_bg_encode_states"));
        out
    }

    fn decode(&self, action: u32) -> Self::Move {
        let s: sys::diee_bg_state = self.into();
        let id = action as u16;
        let mut mv = sys::diee_move { from1: 0, to1: 0, from2: 0, to2: 0 };
        with_ctx(|c| c.check(unsafe { sys::diee_bg_decode_moves(c.raw, &s, &id, 1, &mut mv) }).expect("diee_bg_decode_moves"));
        Self::move_to_actions(&mv)
    }

    fn encode(&self, action: &Self::Move) -> u32 {
        assert!(action.len() <= 2); // backgammon_logic.rs:263
        let s: sys::diee_bg_state = self.into();
        let mv = Self::actions_to_move(action);
        let mut id = 0u16;
        with_ctx(|c| c.check(unsafe { sys::diee_bg_encode_moves(c.raw, &s, &mv, 1, &mut id) }).expect("diee_bg_encode_moves"));
        id as u32
    }

    fn get_id(&self) -> usize {
        self.id
    }

    fn set_id(&mut self, new_id: usize) {
        self.id = new_id;
    }

    fn to_pretty_str(&self) -> String {
        // the reference draws an ASCII board (backgammon_logic.rs:110-174, not on the hot path); a compact form here
        format!("player {} roll {:?} second {} bar {:?} off {:?}\n{:?}", self.player, self.roll, self.is_second_play,
                self.board.1, self.board.2, self.board.0)
    }
}
