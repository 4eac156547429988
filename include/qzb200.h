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

#define QZ_ABI_VERSION 2

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

/* Kernel launches are asynchronous: the int every entry point returns only reports argument and LAUNCH errors.  A fault
 * inside a kernel (an illegal address in a persistent kernel, say) would otherwise surface at some later, unrelated
 * call.  qz_stream_check waits for `stream` and returns the device's error state: 0, or the positive cudaError_t with
 * its text in qz_last_error_string().  Callers place it where they synchronise anyway (BatchedMCTS.drain(check=True),
 * the tests); with the environment variable QZ_SYNC_CHECK set, EVERY entry point does it after launching (debugging). */
int qz_stream_check(void *stream);

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

/* The same for the positions whose flags[i] & flag_bits != 0 only (the others' masks are left untouched): the lazily
 * expanded search sweeps just the leaves that came back to a node (flags = qz_tree.leaf_flags, QZ_LEAF_NEEDS_MASK).
 * key / group (nullable / 0): positions are grouped in runs of `group` (the leaves of one game); a flagged position whose
 * key equals that of an earlier flagged position of its group (the same tree node, key = qz_tree.leaf_node) is not swept
 * but marked with dup_bits in flags (qz_mcts_extend builds a node's block from the first of them; QZ_LEAF_DUPLICATE). */
int qz_env_legal_mask_flagged(const qz_state *states, uint8_t *flags, int flag_bits, uint64_t *mask3, int64_t n,
                              const int32_t *key, int32_t group, int dup_bits, void *stream);

/*
 * The reference's random policy over the FULL legal set (pure_mcts.rollout_policy_fn, pure_mcts.py:7-10: argmax of
 * iid uniforms over actions() == one uniform legal action): actions[i] = the k-th set bit of game i's legal mask,
 * k = (word * popcount) >> 32, word = Philox4x32-10(key = seed, counter = (game_id[i] or i, ply >> 2, 0x7000))[ply & 3].
 * -1 for finished games and games with an empty mask.  With qz_env_legal_mask and qz_env_step this is BASELINE
 * config 1 literally: "env step + valid-action mask" for every ply of a random game.
 */
int qz_env_sample_legal(const qz_state *states, const uint64_t *mask3, uint64_t seed, const int64_t *game_id,
                        int32_t *actions, int64_t n, void *stream);

/*
 * BASELINE config 1 in two launches: every game of `states` is played to its end (or until its ply counter reaches
 * max_plies) by uniform-random legal moves, the FULL legal set being computed on every ply (Quoridor.actions +
 * the pick of qz_env_sample_legal + Quoridor.step) -- the same games, ply for ply, as looping the three calls.
 * A stalemated game gets its "stalemate" flag.  The ply counters in states[i].meta give the plies played.
 */
int qz_env_random_play(qz_state *states, uint64_t seed, const int64_t *game_id, int32_t max_plies, int64_t n,
                       void *stream);

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
 *   workspace            qz_rollout_workspace_bytes(n_rollouts) bytes of device scratch (8-byte aligned).  The first
 *                        64 64-bit words are work / list counters (zeroed by the call) except word 1, which
 *                        ACCUMULATES the plies played by every call (caller zeroes / reads it), i.e. the env-step
 *                        count without a per-rollout reduction; then one qz_state per rollout parks the position
 *                        between the phases (wall kernel, stuck kernel, qz_rollout_pawn_passes(limit) pawn passes),
 *                        the list of ejected rollouts and two lists of rollouts suspended between pawn passes.
 *   flags                QZ_ROLLOUT_DEFER_STUCK: the few rollouts ejected from the wall phase ("stuck": walls in hand
 *                        but next to no legal placement, hundreds of serial plies) are NOT finished by this call;
 *                        their result is QZ_ROLLOUT_PENDING (-128) until qz_rollout_finish, called later -- typically on
 *                        another stream while the caller does other work -- with the SAME arguments and buffers, has
 *                        run.  Results are identical either way.
 */
#define QZ_ROLLOUT_DEFER_STUCK 1
#define QZ_ROLLOUT_PENDING (-128)
int64_t qz_rollout_workspace_bytes(int64_t n_rollouts);
/* kernel launches of the pawn phase of one qz_rollout / qz_rollout_finish call (a rollout plays a bounded slice
 * of plies per pass -- 128, 128, 128, 256, then the rest -- and the survivors are re-packed into full warps in between) */
int32_t qz_rollout_pawn_passes(int32_t limit);
int qz_rollout(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
               int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
               int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, int32_t flags, void *stream);
int qz_rollout_finish(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
                      int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
                      int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Batched PUCT MCTS (mcts.py, pure_mcts.py).  n_games trees live in flat, caller-owned device arrays.
 *
 * qz_tree is a HOST struct of device pointers and sizes, passed by pointer and copied by the call.
 * Game g owns the slots [g*node_cap, (g+1)*node_cap) of five arrays.  Children are created LAZILY: TreeNode.expand
 * (mcts.py:27-35) writes a 3-slot block header (the node's 140-bit legal mask, children in existence, capacity, legal
 * count) followed by room for child slots, and a child gets its slot the first time the descent picks it -- a child
 * never visited has Q = 0 and n = 0, so the reference's first-max over all children (mcts.py:42) is reproduced exactly
 * from the existing children plus ONE candidate (the unvisited child with the largest c_puct*P, lowest actions() rank
 * on ties; with uniform priors simply the next action in actions() order).  See csrc/qz_mcts.cu for the encoding.
 *   child slot : prior f32 (TreeNode._P), visits i32 (_n_visits), q f64 (_Q), child_base i32 (-1 = is_leaf(), >= 0 = its
 *                block, <= -2 = moved to slot -2 - x), node_meta u32 = action | actions() rank << 8 | in-flight << 16
 *   per game   : root (slot index, always 0 after init/reroot), n_nodes (bump allocator), root_state;
 *                with stored (non-uniform) priors: prior_pool, pool_cap floats per game -- an expanded node keeps one
 *                float per legal action, in actions() order -- and n_pool, the floats in use
 *   per leaf   : (n_games * leaves_per_game entries, refilled by every select) leaf_node, leaf_state,
 *                path[max_depth] + path_len (root..leaf slot indices), leaf_flags (QZ_LEAF_*)
 * Sizing: a playout adds at most one child slot (amortised <= 4 with block doubling) and one block (3 + 4 slots), the
 * root's block has room for all its children: node_cap >= 160 + 12 * playouts kept per game is ample
 * (~0.3 MB per game for 1000 playouts; the eager layout of ABI 1 needed 3.4 MB).
 */
typedef struct qz_tree {
    int64_t n_games;
    int32_t node_cap;
    int32_t max_depth;
    int32_t leaves_per_game;
    int32_t pool_cap;           /* floats of prior_pool per game (0 = uniform priors only) */
    float *prior;
    int32_t *visits;
    double *q;
    int32_t *child_base;
    uint32_t *node_meta;
    int32_t *root;
    int32_t *n_nodes;
    qz_state *root_state;
    int32_t *leaf_node;
    qz_state *leaf_state;
    int32_t *path;
    int32_t *path_len;
    uint8_t *leaf_flags;
    float *prior_pool;          /* [n_games * pool_cap] or NULL */
    int32_t *n_pool;            /* [n_games] or NULL */
} qz_tree;

#define QZ_LEAF_TERMINAL 0x01       /* the game is over at the leaf (never sent to the evaluator's result) */
#define QZ_LEAF_DEPTH_OVERFLOW 0x02 /* descent stopped at max_depth */
#define QZ_LEAF_ARENA_OVERFLOW 0x04 /* children did not fit node_cap; leaf left unexpanded */
#define QZ_LEAF_DUPLICATE 0x08      /* another leaf of the same wave expanded this node first */
#define QZ_LEAF_INACTIVE 0x10       /* slot k >= k_leaves of this wave */
#define QZ_LEAF_PENDING 0x20        /* value was QZ_ROLLOUT_PENDING: expanded, backup owed (qz_mcts_backup_pending) */
#define QZ_LEAF_NEEDS_MASK 0x40     /* lazy expansion: the descent stopped at a visited node without a block (qz_mcts_extend) */

/* MCTS.__init__ (mcts.py:89-100) / update_with_move(-1) (:150-151): fresh root (prior 1.0) for every game
 * (or only those with select[g] != 0); root_states (nullable) are copied into tree->root_state. */
int qz_mcts_init(const qz_tree *tree, const qz_state *root_states, const uint8_t *select, void *stream);

/* The descent of MCTS._playout (mcts.py:107-113; pure_mcts.py:68-73): for each game, k_leaves (<=
 * leaves_per_game) descents from the root by first-max of Q + c_puct*P*sqrt(N_parent)/(1+n) (TreeNode.select /
 * get_value, mcts.py:37-42,64-70), replaying Quoridor.step from root_state.  Fills the per-leaf arrays.
 * uniform_prior != 0: priors are 1/len(children) in float64 (pure_mcts.py:13-16) instead of the stored f32.
 * With k_leaves > 1 later descents see a virtual loss on earlier paths (deviation; k_leaves = 1 is exact).
 * A child picked for the first time gets its slot here (overflow_count, nullable, counts the slots that did not fit).
 * lazy_expand != 0 (uniform priors only): see qz_mcts_extend below.  With leaves_per_game > 1 the root's per-child terms
 * are kept as reciprocals in shared memory (one fused multiply-add per child instead of two float64 divisions); the
 * one-leaf-per-wave search keeps the reference's literal arithmetic. */
int qz_mcts_select(const qz_tree *tree, double c_puct, int uniform_prior, int k_leaves, int lazy_expand,
                   int32_t *overflow_count, void *stream);

/* Lazy expansion (uniform priors only; lazy_expand != 0 above and mask3 == NULL below).  pure_mcts.py:75-83 expands a leaf
 * at its first visit, but nothing reads the children until a later playout comes back to the node -- and most leaves of a
 * pure-MCTS search are never visited again.  With lazy_expand the descent stops at a visited node that has no block yet
 * (QZ_LEAF_NEEDS_MASK) and this call computes that node's legal actions (Quoridor.actions, quoridor.py:138-157), builds
 * the block and takes the descent one more PUCT level (mcts.py:37-42), leaving leaf_node / leaf_state / path as the
 * eager form would.  Same visit counts as the eager form at k_leaves = 1.  mask3 (nullable): the legal masks of the flagged
 * leaves computed beforehand by qz_env_legal_mask_flagged(leaf_state, leaf_flags, QZ_LEAF_NEEDS_MASK, ...) -- a warp per
 * leaf instead of this kernel's warp per game; NULL = computed here.  root_only != 0: no descent; only gives a root that
 * was evaluated but never revisited (n_playout = 1) its block, so that the statistics see its children. */
int qz_mcts_extend(const qz_tree *tree, double c_puct, const uint64_t *mask3, int root_only, int32_t *overflow_count,
                   void *stream);

/* The rest of MCTS._playout (mcts.py:117-127): for every leaf of the last select, unless terminal, expand
 * with (action, prior) over the legal actions in actions() order (TreeNode.expand, mcts.py:27-35; priors
 * [n*K,140] are read at the legal actions only and NOT renormalised, policy_value_net.py:162; NULL = uniform;
 * mask3 == NULL with priors == NULL = lazy expansion: nothing is expanded here, see qz_mcts_extend)
 * and back up -leaf_value along the path with a sign flip per level (update_recursive, mcts.py:44-62).
 * Exactly one of value_f32 / value_f64 / value_i8 is the evaluator's value for the side to move.  Terminal
 * leaves use +1 (the reference's inverted sign, mcts.py:125) or -1 when fix_terminal_sign != 0. */
int qz_mcts_expand_backup(const qz_tree *tree, const uint64_t *mask3, const float *priors, const float *value_f32,
                          const double *value_f64, const int8_t *value_i8, double c_puct, int fix_terminal_sign,
                          int32_t *overflow_count, void *stream);

/* Deferred half of qz_mcts_expand_backup: a leaf whose value_i8 was QZ_ROLLOUT_PENDING (its rollout is being
 * finished by qz_rollout_finish) was expanded but not backed up and keeps its in-flight marks; this call backs up
 * every such leaf of the tree's CURRENT leaf arrays with the now final value_i8 (update_recursive, mcts.py:44-62). */
int qz_mcts_backup_pending(const qz_tree *tree, const int8_t *value_i8, int fix_terminal_sign, void *stream);

/* get_move_probs (mcts.py:141-144): per game, root-child visit counts / Q scattered by action id into
 * [n,140] arrays (nullable each), softmax(1/temp*log(visits+1e-10)) in float64, and the root's own visits. */
int qz_mcts_root_stats(const qz_tree *tree, double temp, int32_t *visits_out, double *q_out, double *probs_out,
                       int32_t *root_n_out, void *stream);

/* MCTSPlayer.choose_action (mcts.py:172-196) / pure_mcts get_move (:115).  mode 0: first max of visits;
 * mode 1: sample from probs (:185); mode 2: sample from (1-noise_eps)*probs + noise_eps*Dirichlet(dir_alpha)
 * (:181 uses 0.75/0.25 and 0.3).  Philox streams keyed by (seed, game_id[g] or g, ply).  -1 = no children. */
int qz_mcts_choose(const qz_tree *tree, int mode, double temp, double noise_eps, double dir_alpha, uint64_t seed,
                   const int64_t *game_id, int32_t *moves_out, void *stream);

/* MCTS.update_with_move (mcts.py:146-151): dst := the subtree of src under moves[g] (statistics kept,
 * compacted breadth-first) or a fresh root if the move is unknown / negative.  apply_move != 0 also advances
 * root_state by the move (Quoridor.step).  src and dst must be distinct arenas of equal shape. */
int qz_mcts_reroot(const qz_tree *src, const qz_tree *dst, const int32_t *moves, int apply_move, int32_t *overflow_count,
                   void *stream);

/* `node._children` of ONE node (mcts.py:19-25) for inspection: every legal action in actions() order, children that
 * never got a slot included.  out_i (int32 [4 + 4*140]): [0] children, [1] the node's visits, [2] its slot; per rank r:
 * [4+4r] action, [5+4r] slot or -1, [6+4r] visits, [7+4r] in-flight marks.  out_d (f64 [2 + 2*140]): [0] the node's Q,
 * [1] its prior; per rank r: [2+2r] Q, [3+2r] prior.  One tiny launch; the host-side TreeNode view uses it. */
int qz_mcts_node_children(const qz_tree *tree, int64_t game, int32_t node, int32_t *out_i, double *out_d, void *stream);

/* Deterministic policy-value stubs for parity testing (tests/golden/stubs.py: kind 1 = S1 uniform, 2 = S2 hash,
 * 3 = S3 hash/8) evaluated on the device: priors f32 [n,140] (0 at illegal actions), values f64 [n]. */
int qz_stub_eval(const qz_state *states, const uint64_t *mask3, int kind, float *priors, double *values, int64_t n,
                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QZB200_H */
