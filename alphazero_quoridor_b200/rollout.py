"""Random rollouts on the GPU: pure_mcts.MCTS._evaluate_rollout (pure_mcts.py:86-108), batched.

`rollout()` launches the persistent one-thread-per-rollout kernel (csrc/qz_rollout.cu) through the C ABI.
"""
import torch

from . import _lib


def workspace_words(n_rollouts):
    """int64 words of scratch `rollout` needs for n_rollouts (word 1 accumulates the plies played)."""
    return (_lib.load().qz_rollout_workspace_bytes(int(n_rollouts)) + 7) // 8


DEFER_STUCK = 1
PENDING = -128


def rollout(states, per_state=1, seed=0, rid_base=0, rids=None, state_index=None, limit=1000,
            return_plies=True, return_final=False, workspace=None, result=None, defer_stuck=False, finish=False):
    """Run uniform-random playouts of at most limit-1 plies.

    states       int64 [n_states,3] CUDA tensor of qz_state rows.
    per_state    rollouts per start state when `state_index` is None (rollout r starts from states[r // per_state]).
    state_index  optional int32 [n_rollouts]: explicit start state of each rollout.
    seed, rid_base / rids  Philox key and per-rollout stream ids (uint64 as int64 tensor for `rids`).
    defer_stuck  leave the few "stuck" rollouts (see csrc/qz_rollout.cu) unfinished: their result is PENDING (-128)
                 until the same call is repeated with finish=True (same tensors, typically on another stream).
    Returns (result int8 [n_rollouts], plies int32 [n_rollouts] | None, final_states int64 [n_rollouts,3] | None).
    """
    _lib.require_cuda()
    lib = _lib.load()
    dev = states.device
    assert states.is_cuda and states.dtype == torch.int64 and states.dim() == 2 and states.shape[1] == 3
    n_states = states.shape[0]
    if state_index is not None:
        state_index = state_index.to(device=dev, dtype=torch.int32).contiguous()
        n_roll = state_index.numel()
    else:
        n_roll = n_states * int(per_state)
    if rids is not None:
        rids = rids.to(device=dev, dtype=torch.int64).contiguous()
        assert rids.numel() == n_roll
    if result is None:
        result = torch.empty((n_roll,), dtype=torch.int8, device=dev)
    assert result.dtype == torch.int8 and result.numel() == n_roll
    plies = torch.empty((n_roll,), dtype=torch.int32, device=dev) if return_plies else None
    final = torch.empty((n_roll, 3), dtype=torch.int64, device=dev) if return_final else None
    need = (lib.qz_rollout_workspace_bytes(n_roll) + 7) // 8
    if workspace is None:
        workspace = torch.zeros((need,), dtype=torch.int64, device=dev)
    assert workspace.dtype == torch.int64 and workspace.numel() >= need, "workspace too small: see workspace_words()"
    assert states.is_contiguous()
    args = [_lib.ptr(states), n_states, _lib.ptr(state_index), int(per_state), n_roll, int(seed) & ((1 << 64) - 1),
            int(rid_base) & ((1 << 64) - 1), _lib.ptr(rids), int(limit), _lib.ptr(result), _lib.ptr(plies),
            _lib.ptr(final), _lib.ptr(workspace)]
    passes = lib.qz_rollout_pawn_passes(int(limit))
    with torch.cuda.device(dev):
        if finish:       # stuck kernel + pawn passes
            _lib.check(lib.qz_rollout_finish(*args, _lib.stream_ptr(dev)), "qz_rollout_finish", launches=1 + passes)
        else:            # wall kernel (+ stuck kernel unless deferred) + pawn passes
            _lib.check(lib.qz_rollout(*args, DEFER_STUCK if defer_stuck else 0, _lib.stream_ptr(dev)), "qz_rollout",
                       launches=(1 if defer_stuck else 2) + passes)
    return result, plies, final
