"""Quoridor environment: the reference's `Quoridor` surface (quoridor.py:5-610) over the CUDA kernels.

Two classes:

* `BatchedQuoridor` -- n games resident in HBM as 24-byte bitboard states; `reset/step/legal_mask/encode`
  are one kernel launch each (libqzb200.so, include/qzb200.h).  This is the hot path.
* `Quoridor` -- the drop-in, single-game mirror of the reference class (same constructor, methods,
  attributes, return types and action ordering).  It is a batch-of-1 view: the Python attributes the
  reference exposes (`_positions`, `_intersections`, ...) stay authoritative on the host exactly as in the
  reference (callers such as game.py and tests assign to them), are packed into a 24-byte state before each
  kernel call and unpacked after it.  All rules arithmetic happens on the GPU; there is no CPU fallback.

Deliberate deviations from the reference: nothing is printed from `step`; a finished game is
left untouched by `step`; `actions()` on a finished game returns [].
"""
import copy

import numpy as np
import torch

from . import _lib

M64 = (1 << 64) - 1

FLAG_DONE, FLAG_STALEMATE, FLAG_TRUNCATED, FLAG_ILLEGAL = 0x01, 0x08, 0x10, 0x20


# ------------------------------------------------------------------------------------------------ host packing
def _to_i64(x):
    x &= M64
    return x - (1 << 64) if x >= (1 << 63) else x


def pack_meta(p1, p2, w1, w2, cur, flags=0, ply=0):
    return ((p1 & 0xFF) | ((p2 & 0xFF) << 8) | ((w1 & 0xFF) << 16) | ((w2 & 0xFF) << 24) | ((cur & 0xFF) << 32)
            | ((flags & 0xFF) << 40) | ((ply & 0xFFFF) << 48))


def unpack_meta(m):
    m &= M64

    def i8(x):
        return x - 256 if x > 127 else x
    return dict(p1=i8(m & 0xFF), p2=i8((m >> 8) & 0xFF), w1=(m >> 16) & 0xFF, w2=(m >> 24) & 0xFF,
                cur=(m >> 32) & 0xFF, flags=(m >> 40) & 0xFF, ply=(m >> 48) & 0xFFFF)


def pack_state(H, V, p1, p2, w1, w2, cur, flags=0, ply=0):
    """-> three Python ints suitable for an int64 tensor row."""
    return [_to_i64(H), _to_i64(V), _to_i64(pack_meta(p1, p2, w1, w2, cur, flags, ply))]


def mask_to_actions(mask3):
    """140-bit legal mask -> list in the reference's actions() order (quoridor.py:157,420-430):
    pawn ids ascending, then H(ix), V(ix) interleaved by intersection."""
    bits = (int(mask3[0]) & M64) | ((int(mask3[1]) & M64) << 64) | ((int(mask3[2]) & M64) << 128)
    out = [a for a in range(12) if (bits >> a) & 1]
    walls = bits >> 12
    if walls:
        for ix in range(64):
            if (walls >> ix) & 1:
                out.append(12 + ix)
            if (walls >> (64 + ix)) & 1:
                out.append(76 + ix)
    return out


# ------------------------------------------------------------------------------------------------ batched env
class BatchedQuoridor:
    """n independent games on one GPU.  States live in `self.states` (int64 [n,3] = qz_state[n])."""

    def __init__(self, n_games, device=None, states=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = int(n_games)
        if states is not None:
            assert states.shape == (self.n, 3) and states.dtype == torch.int64
            self.states = states.to(self.device).contiguous()
        else:
            self.states = torch.empty((self.n, 3), dtype=torch.int64, device=self.device)
            self.reset()

    def _stream(self):
        return _lib.stream_ptr(self.device)

    def reset(self):
        """Quoridor.reset for every game (quoridor.py:26-56)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_env_reset(_lib.ptr(self.states), self.n, self._stream()), "qz_env_reset")
        return self

    def legal_mask(self, out=None):
        """Quoridor.actions for every game as 140-bit masks, int64 [n,3] (quoridor.py:138-157)."""
        if out is None:
            out = torch.empty((self.n, 3), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_env_legal_mask(_lib.ptr(self.states), _lib.ptr(out), self.n, self._stream()),
                       "qz_env_legal_mask")
        return out

    def step(self, actions, legal_mask=None, done=None):
        """Quoridor.step for every game (quoridor.py:159-186).  `actions` int32 [n] (negative = skip).
        With `legal_mask` given (safe=True semantics) illegal actions are rejected and flagged.
        Returns the uint8 [n] done vector."""
        actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        assert actions.shape == (self.n,)
        if done is None:
            done = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_env_step(_lib.ptr(self.states), _lib.ptr(actions), _lib.ptr(legal_mask),
                                            _lib.ptr(done), self.n, self._stream()), "qz_env_step")
        return done

    def sample_actions(self, legal_mask, seed=0, game_id=None, out=None):
        """One uniformly random legal action per game from its full legal mask (pure_mcts.py:7-10), int32 [n];
        -1 for finished / stalemated games.  Philox stream keyed by (seed, game_id[i] or i, ply)."""
        if out is None:
            out = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_env_sample_legal(_lib.ptr(self.states), _lib.ptr(legal_mask), int(seed) & M64,
                                                    _lib.ptr(game_id), _lib.ptr(out), self.n, self._stream()),
                       "qz_env_sample_legal")
        return out

    def random_play(self, seed=0, max_plies=3000, game_id=None, fused=True, check_every=16):
        """BASELINE config 1: uniform-random legal play to terminal with the FULL legal set computed every ply.
        fused=True: two launches (qz_env_random_play); fused=False: the legal_mask -> sample_actions -> step loop,
        which plays exactly the same games.  Returns the number of plies played per game (int64 [n])."""
        if fused:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.qz_env_random_play(_lib.ptr(self.states), int(seed) & M64, _lib.ptr(game_id),
                                                       int(max_plies), self.n, self._stream()), "qz_env_random_play")
            return (self.states[:, 2] >> 48) & 0xFFFF
        mask = torch.empty((self.n, 3), dtype=torch.int64, device=self.device)
        acts = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        done = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        for ply in range(max_plies):
            self.legal_mask(out=mask)
            self.sample_actions(mask, seed=seed, game_id=game_id, out=acts)
            self.step(acts, done=done)
            if ply % check_every == check_every - 1 and bool(((done != 0) | (acts < 0)).all()):
                break
        return (self.states[:, 2] >> 48) & 0xFFFF

    def encode(self, out=None, dtype=torch.float32, channels_last=False, c_stride=26):
        """Quoridor.state for every game (quoridor.py:58-131), written as `dtype` into `out`.
        NCHW: out [n,26,9,9] contiguous.  channels_last: out is a [n,c_stride,9,9] tensor in
        torch.channels_last memory format (physical [n,9,9,c_stride]); channels >= 26 are zero."""
        if out is None:
            if channels_last:
                out = torch.empty((self.n, c_stride, 9, 9), dtype=dtype, device=self.device,
                                  memory_format=torch.channels_last)
            else:
                out = torch.empty((self.n, 26, 9, 9), dtype=dtype, device=self.device)
        code = _lib.DTYPE_CODE[out.dtype]
        if channels_last:
            assert out.is_contiguous(memory_format=torch.channels_last) and out.shape[1] == c_stride
        else:
            assert out.is_contiguous() and tuple(out.shape[1:]) == (26, 9, 9)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qz_env_encode(_lib.ptr(self.states), _lib.c_void_p_of(out),
                                              code, _lib.LAYOUT_NHWC if channels_last else _lib.LAYOUT_NCHW,
                                              c_stride, self.n, self._stream()), "qz_env_encode")
        return out

    # ---- host-side views (tests / debugging; each one synchronises) ----
    def host_states(self):
        rows = self.states.cpu().numpy().astype(np.int64)
        out = []
        for H, V, m in rows:
            d = unpack_meta(int(m))
            d["H"], d["V"] = int(H) & M64, int(V) & M64
            d["done"] = bool(d["flags"] & FLAG_DONE)
            d["winner"] = (d["flags"] >> 1) & 3
            out.append(d)
        return out

    def legal_lists(self):
        m = self.legal_mask().cpu().numpy()
        return [mask_to_actions(row) for row in m]


# ------------------------------------------------------------------------------------------------ drop-in class
class Quoridor(object):
    """Drop-in for the reference `Quoridor` (quoridor.py:5-610)."""

    HORIZONTAL = 1
    VERTICAL = -1

    def __init__(self, safe=False, device=None):
        self.safe = safe
        self.action_space = 140
        self.n_players = 2
        self.players = [1, 2]
        self._device = device
        self._dev = None          # lazily created BatchedQuoridor(1): scratch for kernel calls
        self.reset()

    # quoridor.py:18-20
    def load(self, p1, p2):
        self.player1 = p1
        self.player2 = p2

    def get_current_player(self):
        return self.current_player

    # quoridor.py:26-56
    def reset(self):
        self.current_player = 1
        self.last_player = -1
        self.tiles = np.zeros(81)
        self._positions = {1: 4, 2: 76}
        self._DIRECTIONS = {'N': 0, 'S': 1, 'E': 2, 'W': 3, 'NN': 4, 'SS': 5, 'EE': 6, 'WW': 7,
                            'NE': 8, 'NW': 9, 'SE': 10, 'SW': 11}
        self.N_DIRECTIONS = 12
        self.N_TILES = 81
        self.N_ROWS = 9
        self.N_INTERSECTIONS = 64
        self._intersections = np.zeros(64)
        self._player1_walls_remaining = 10
        self._player2_walls_remaining = 10
        self._ply = 0

    # ---- host <-> device state ----
    def _masks(self):
        H = V = 0
        for ix in range(64):
            w = self._intersections[ix]
            if w == 1:
                H |= 1 << ix
            elif w == -1:
                V |= 1 << ix
        return H, V

    def packed(self):
        """The game as a qz_state row (3 x int64)."""
        H, V = self._masks()
        over, winner = self.has_a_winner()
        flags = (FLAG_DONE | (winner << 1)) if over else 0
        return pack_state(H, V, int(self._positions[1]), int(self._positions[2]),
                          int(self._player1_walls_remaining), int(self._player2_walls_remaining),
                          int(self.current_player), flags, self._ply & 0xFFFF)

    def _upload(self):
        if self._dev is None:
            self._dev = BatchedQuoridor(1, device=self._device)
        row = torch.tensor([self.packed()], dtype=torch.int64)
        self._dev.states.copy_(row, non_blocking=False)
        return self._dev

    def _download(self):
        H, V, m = [int(x) for x in self._dev.states.cpu().numpy()[0]]
        d = unpack_meta(m)
        H &= M64
        V &= M64
        for ix in range(64):
            self._intersections[ix] = 1 if (H >> ix) & 1 else (-1 if (V >> ix) & 1 else 0)
        self._positions[1], self._positions[2] = d["p1"], d["p2"]
        self._player1_walls_remaining, self._player2_walls_remaining = d["w1"], d["w2"]
        self._ply = d["ply"]
        return d

    # quoridor.py:58-131
    def state(self):
        for p in (self._positions[1], self._positions[2]):
            if p > 80 or p < -81:
                raise IndexError("index %d is out of bounds for axis 0 with size 81" % p)
        dev = self._upload()
        return dev.encode(dtype=torch.float32)[0].cpu().numpy().astype(np.float64)

    def load_state(self, state):
        """quoridor.py:133-136 is a TODO stub in the reference; kept as such."""
        current_player = state[-1] == np.zeros([9, 9])  # noqa: F841

    # quoridor.py:138-157
    def actions(self):
        if self.has_a_winner()[0]:
            return []
        dev = self._upload()
        return mask_to_actions(dev.legal_mask().cpu().numpy()[0])

    # quoridor.py:159-186
    def step(self, action):
        player = self.current_player
        if self.has_a_winner()[0]:
            return True
        dev = self._upload()
        mask = dev.legal_mask()
        self.valid_actions = mask_to_actions(mask.cpu().numpy()[0])
        if self.safe and action not in self.valid_actions:
            raise ValueError("Invalid Action: {action}".format(action=action))
        if not (0 <= int(action) < 140):
            raise ValueError("Invalid Pawn Action: {action}".format(action=action))
        act = torch.tensor([int(action)], dtype=torch.int32, device=dev.device)
        done = bool(dev.step(act).cpu().numpy()[0])
        d = self._download()
        if not done:
            self.current_player = d["cur"]
            self.last_player = player
        return done

    def game_end(self):
        pass

    # quoridor.py:193-202
    def has_a_winner(self):
        game_over = False
        winner = None
        if self._positions[2] < 9:
            winner = 2
            game_over = True
        elif self._positions[1] > 71:
            winner = 1
            game_over = True
        return game_over, winner

    def _get_rewards(self):
        """quoridor.py:205-214, literally (unused by the reference): (-1, 1) when P1 has won; when P2 has won the
        reference's `rewards, done = (1, -1)` at :208 unpacks the tuple, so the call returns (1, -1) itself."""
        if self._positions[2] < 9:
            return 1, -1
        if self._positions[1] > 71:
            return (-1, 1), True
        return (0, 0), False

    # quoridor.py:260-269
    def rotate_players(self):
        if self.current_player == 1:
            self.current_player = 2
            self.last_player = 1
        else:
            self.current_player = 1
            self.last_player = 2

    # quoridor.py:530-531
    def add_wall(self, wall, orientation):
        self._intersections[wall] = orientation

    # quoridor.py:533-568
    def print_board(self):
        p1r, p1c = divmod(self._positions[1], 9)
        p2r, p2c = divmod(self._positions[2], 9)
        grid = [['{:4}'.format('-') for _ in range(9)] for _ in range(9)]
        grid[p1r][p1c] = '{:4}'.format('X')
        grid[p2r][p2c] = '{:4}'.format('O')
        rows = self._intersections.reshape([8, 8])
        ir = 7
        for i in range(8, -1, -1):
            print(''.join(grid[i]))
            if ir >= 0:
                print('{:2}'.format(''), end='')
                for j in rows[ir, :]:
                    print('{:4}'.format('h' if j == 1 else ('v' if j == -1 else '')), end='')
                ir -= 1
                print()

    # quoridor.py:570-571 -- returns a NEW game, not a copy (reference behaviour)
    def clone(self):
        return Quoridor()

    def __deepcopy__(self, memo):
        new = Quoridor.__new__(Quoridor)
        for k, v in self.__dict__.items():
            if k == "_dev":
                new._dev = None
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    # quoridor.py:573-610
    def start_self_play(self, player, is_shown=0, temp=1e-3):
        self.reset()
        states, mcts_probs, current_players = [], [], []
        while True:
            move, move_probs = player.choose_action(self, temp=temp, return_prob=1)
            states.append(self.state())
            mcts_probs.append(move_probs)
            current_players.append(self.current_player)
            self.step(move)
            end, winner = self.has_a_winner()
            if end:
                winners_z = np.zeros(len(current_players))
                if winner != -1:
                    winners_z[np.array(current_players) == winner] = 1.0
                    winners_z[np.array(current_players) != winner] = -1.0
                player.reset_player()
                if is_shown:
                    print("Game end. Winner is player:", winner)
                return winner, zip(states, mcts_probs, winners_z)
