/*
 * qzb200.h -- C ABI of libqzb200.so, the B200 (sm_100a) Quoridor self-play engine.
 *
 * The reference (cryer/AlphaZero_Quoridor) has no FFI layer: its boundary is the duck-typed Python
 * API of quoridor.py / mcts.py / pure_mcts.py / policy_value_net.py.  Each entry point below names
 * the reference method(s) it replaces (file:line into the reference); the Python mirror of that API
 * (alphazero_quoridor_b200/{quoridor,mcts,pure_mcts}.py) is a thin ctypes shim over these symbols and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer into caller-owned memory (a torch CUDA tensor's data_ptr()),
 *     unless the parameter is documented as host memory.  The library never allocates, frees or
 *     retains a pointer past the call and keeps no global mutable state.
 *   - Every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*;
 *     NULL = the legacy default stream).  Re-entrant; safe from several host threads on different streams.
 *   - Return value: 0 = ok; negative = argument error (QZ_E_*); positive = the cudaError_t of the launch.
 *     Nothing throws.  qz_last_error_string() returns a thread-local description of the last failure.
 *   - There is no CPU fallback: on a machine without a CUDA device every compute call returns a
 *     positive cudaError_t.
 *
 * Data layout
 *   qz_state (24 B, 8-byte aligned), one per game:
 *       u64 H     horizontal walls, bit ix = r*8+c        (quoridor.py:49-53, value +1)
 *       u64 V     vertical walls                            (value -1)
 *       u64 meta  byte0 P1 tile (int8; 81..89 after an off-board winning jump)   quoridor.py:34-37
 *                 byte1 P2 tile (int8; -9..-1 likewise)
 *                 byte2 walls left P1, byte3 walls left P2                         quoridor.py:55-56
 *                 byte4 mover (1|2)                                                quoridor.py:27
 *                 byte5 flags: bit0 done, bits1-2 winner, bit3 stalemate, bit4 truncated, bit5 illegal
 *                 bytes6-7 ply counter
 *   legal mask: 3 x u64 per game, bit a of the 192-bit little-endian word = action a is legal
 *       actions 0..11 pawn (N,S,E,W,NN,SS,EE,WW,NE,NW,SE,SW), 12..75 horizontal wall at intersection
 *       a-12, 76..139 vertical wall at intersection a-76                           quoridor.py:12,39-43,154
 *       The reference's actions() ORDER (pawn ids ascending, then H(ix),V(ix) interleaved by ix,
 *       quoridor.py:157,420-430) is a pure function of the mask; MCTS children are stored in that order.
 */
#ifndef QZB200_H
#define QZB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QZ_ABI_VERSION 1

#define QZ_N_ACTIONS 140
#define QZ_N_PLANES 26
#define QZ_STATE_ELEMS (26 * 9 * 9)

/* argument errors */
#define QZ_E_NULL (-1)     /* a required pointer is NULL */
#define QZ_E_RANGE (-2)    /* a size / enum argument is out of range */
#define QZ_E_ALIGN (-3)    /* a pointer is not aligned for its element type */

/* qz_env_encode dtype / layout */
#define QZ_DTYPE_F32 0
#define QZ_DTYPE_BF16 1
#define QZ_DTYPE_F16 2
#define QZ_LAYOUT_NCHW 0   /* [n,26,9,9] contiguous, the reference's layout (quoridor.py:106-131) */
#define QZ_LAYOUT_NHWC 1   /* [n,9,9,C] (torch channels_last storage), C = channel stride >= 26 */

typedef struct qz_state {
    uint64_t H, V, meta;
} qz_state;

/* ABI version of the loaded library (== QZ_ABI_VERSION). Host-only, no device needed. */
int qz_version(void);

/* Thread-local text for the last non-zero return on this thread ("" if none). */
const char *qz_last_error_string(void);

/* Number of SMs of the current device (grid sizing for callers); negative/positive error as above. */
int qz_device_sm_count(int *out_sm_count);

/* Quoridor.__init__/reset (quoridor.py:9-56): n initial states. */
int qz_env_reset(qz_state *states, int64_t n, void *stream);

/*
 * Quoridor.step (quoridor.py:159-186) with has_a_winner (:193-202), _handle_pawn_action (:217-243),
 * _handle_wall_action (:246-257) and rotate_players (:260-269) for n games at once.
 *   actions[i]   action id 0..139 for game i; a negative id leaves game i untouched.
 *   legal_mask3  NULL = the reference's safe=False behaviour (apply unchecked).  Otherwise the masks from
 *                qz_env_legal_mask for the SAME states: an action outside the mask is not applied and the
 *                state's "illegal" flag is set (the shim raises ValueError, quoridor.py:167-169).
 *   done[i]      (nullable) 1 if game i is over after the call -- the reference's return value.
 * Finished games are left untouched (done stays 1).
 */
int qz_env_step(qz_state *states, const int32_t *actions, const uint64_t *legal_mask3, uint8_t *done, int64_t n,
                void *stream);

/*
 * Quoridor.actions (quoridor.py:138-157): pawn moves (_valid_pawn_actions :272-353) plus, while the mover
 * has walls left, the 128-candidate wall sweep (_valid_wall_actions :420-430, _validate_* :432-461,
 * _blocks_path :463-477, _bfs_to_goal :479-528).  One warp per game.  Finished games get an all-zero mask;
 * an all-zero mask on a live game is a stalemate (the reference returns [] and crashes downstream).
 */
int qz_env_legal_mask(const qz_state *states, uint64_t *mask3, int64_t n, void *stream);

/*
 * Quoridor.state (quoridor.py:58-131): the 26 x 9 x 9 planes of each game, written in `dtype`
 * straight into the policy-value net's input buffer.
 *   layout NCHW: out is [n][26][9][9].   layout NHWC: out is [n][9][9][c_stride], channels >= 26 zeroed.
 * Games whose pawn has left the board (terminal; the reference raises IndexError) are encoded with
 * that pawn plane empty.
 */
int qz_env_encode(const qz_state *states, void *out, int dtype, int layout, int c_stride, int64_t n, void *stream);

/*
 * pure_mcts.MCTS._evaluate_rollout (pure_mcts.py:86-108) with rollout_policy_fn (:7-10): n_rollouts uniform-
 * random playouts of at most limit-1 plies, one thread per rollout (persistent kernel, finished threads
 * claim the next rollout).  Also BASELINE config 1 (random legal play from reset()).
 *   states / n_states    start positions.  Rollout r starts from states[state_index[r]] or, when
 *                        state_index is NULL, from states[r / per_state].
 *   seed, rid_base, rids Philox4x32-10 key and per-rollout counter id: rids[r] or rid_base + r.  The draw
 *                        procedure is specified in csrc/qz_sample.cuh; a rollout's outcome depends only on
 *                        (start state, seed, rid), never on the launch shape or the GPU count.
 *   limit                the reference's `limit` (1000): at most limit-1 plies are played.
 *   result[r]            +1 / -1 from the STARTING mover's point of view, 0 if nobody won (:104-108).
 *   plies[r]             (nullable) plies played.     final_states[r]  (nullable) where the rollout ended.
 *   workspace            >= 8 bytes of device scratch, zeroed by the call.
 */
int qz_rollout(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
               int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
               int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QZB200_H */
