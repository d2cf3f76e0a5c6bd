"""die_e_b200 -- B200-native engine for the die-e hot path (batched MCTS over backgammon games).

Host-side mirror of the reference's game/agent API (alibasaran/die-e src/base.rs, src/backgammon,
src/tictactoe, src/mcts) over the C ABI of libdiee_cuda.so (include/diee.h).  Every compute call
runs a hand-written sm_100a kernel; there is no CPU fallback.
"""
from . import _ffi
from ._ffi import Context, DieeError, default_context
from .backgammon import Backgammon
from .tictactoe import TicTacToe
from .mcts import MctsConfig, mct_search, mct_search_batch
from .nnet import ResNet
from .alphazero import AlphaZero, AlphaZeroConfig, MemoryFragment, alpha_mcts_parallel, memory_to_arrays

__all__ = ["Context", "DieeError", "default_context", "Backgammon", "TicTacToe", "MctsConfig", "mct_search",
           "mct_search_batch", "ResNet", "AlphaZero", "AlphaZeroConfig", "MemoryFragment", "alpha_mcts_parallel",
           "memory_to_arrays", "_ffi"]
