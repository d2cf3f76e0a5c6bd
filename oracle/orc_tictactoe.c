/*
 * oracle/orc_tictactoe.c -- CPU ORACLE (test infrastructure only; see orc.h).
 * Restates src/tictactoe/mod.rs of alibasaran/die-e.
 */
#include "orc.h"
#include <string.h>

/* TicTacToe::new  tictactoe/mod.rs:28-30 */
void orc_ttt_new(orc_ttt_state *s) {
    memset(s, 0, sizeof *s);
    s->player = -1;
}

/* get_valid_moves  :36-44 (empty cells ascending) */
int orc_ttt_valid_moves(const orc_ttt_state *s, uint8_t *out) {
    int n = 0;
    for (int i = 0; i < 9; ++i)
        if (s->board[i] == 0) out[n++] = (uint8_t)i;
    return n;
}

/* apply_move  :46-49 */
void orc_ttt_apply_move(orc_ttt_state *s, uint8_t m) {
    s->board[m] = s->player;
    s->player = (int8_t)(-s->player);
}

/* skip_turn  :51-53 */
void orc_ttt_skip_turn(orc_ttt_state *s) { s->player = (int8_t)(-s->player); }

/* check_winner  :59-79 (first matching line in table order; full board -> draw 0) */
int orc_ttt_check_winner(const orc_ttt_state *s) {
    static const int8_t L[8][3] = {{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {0, 3, 6},
                                   {1, 4, 7}, {2, 5, 8}, {0, 4, 8}, {2, 4, 6}};
    for (int k = 0; k < 8; ++k) {
        int a = s->board[L[k][0]], b = s->board[L[k][1]], c = s->board[L[k][2]];
        if (a != 0 && a == b && b == c) return a;
    }
    for (int i = 0; i < 9; ++i)
        if (s->board[i] == 0) return ORC_NO_WINNER;
    return 0;
}

/* as_tensor  :81-92 -> [1,3,3,3] planes (== -1, == 0, == 1) */
void orc_ttt_as_tensor(const orc_ttt_state *s, float *o) {
    for (int i = 0; i < 9; ++i) {
        o[0 * 9 + i] = s->board[i] == -1 ? 1.0f : 0.0f;
        o[1 * 9 + i] = s->board[i] == 0 ? 1.0f : 0.0f;
        o[2 * 9 + i] = s->board[i] == 1 ? 1.0f : 0.0f;
    }
}
