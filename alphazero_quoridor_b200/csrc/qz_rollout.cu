// qz_rollout.cu -- K4: random rollouts to terminal, one thread per rollout, persistent with refill.
// Replaces pure_mcts.MCTS._evaluate_rollout (pure_mcts.py:86-108) + rollout_policy_fn (:7-10); also serves
// BASELINE config 1 (uniform-random legal play from reset()).
//
// Work distribution: rollouts have very uneven lengths (tens to ~1000 plies), so a thread that finishes
// immediately claims the next unstarted rollout from a global counter (warp-aggregated atomicAdd) and the
// warp stays converged at the top of one "advance every live lane by one ply" loop.  A rollout is cut in two
// PHASES run by two kernels -- plies while walls remain (flood fills, ~2000 instructions each) and pawn-only
// plies (~140 instructions each) -- so that the lanes of a warp always execute the same kind of ply; the
// first capture (profiles/r1a_rollout_ncu_full.txt) of the single-kernel form showed 4.2 active lanes per
// instruction because a warp mixed both kinds.  No tensor cores: the game lives in registers (the pawn phase
// adds a byte-per-tile table in shared memory) and HBM sees 24 B in, 48 B through `mid`, 1..29 B out per ROLLOUT.
#include <stdlib.h>

#include "qz_common.cuh"
#include "qz_sample.cuh"
#include "qz_warp.cuh"

struct QzRolloutArgs {
    const qz_state *states;        // [n_states]
    const int32_t *state_index;    // nullable: rollout r starts from states[state_index[r]] (else r / per_state)
    const uint64_t *rids;          // nullable: RNG id of rollout r (else rid_base + r)
    uint64_t rid_base;
    uint64_t seed;
    int64_t n_rollouts;
    int32_t per_state;
    int32_t limit;
    int8_t *result;                // [n_rollouts] +1/-1/0 from the starting mover's view
    int32_t *plies;                // nullable [n_rollouts]
    qz_state *final_states;        // nullable [n_rollouts]
    unsigned long long *counter;   // workspace header, QZ_WS_COUNTERS words: [0] wall-phase work counter, [1] cumulative
                                   // plies (never zeroed), [3] number of ejected ("stuck") rollouts, [4] stuck-phase work
                                   // counter, [8 + 2j] work counter of pawn pass j, [9 + 2j] rollouts surviving pass j
    qz_state *mid;                 // [n_rollouts] state when a rollout leaves a phase / is suspended
    int32_t *stuck_list;           // [n_rollouts] rollouts ejected from the wall phase
    // one pawn-phase pass (set by qz_pawn_passes)
    const int32_t *list_in;        // nullable: the rollouts of this pass (else all of them)
    const unsigned long long *count_in;   // nullable: how many (else n_rollouts)
    int32_t *list_out;             // rollouts suspended by this pass, *count_out of them
    unsigned long long *count_out;
    unsigned long long *work;      // work counter of this pass
    int32_t slice;                 // plies a rollout may play per pass
};

// Warp-aggregated claim of the next unstarted rollout for every idle lane; returns -1 when none is left.
__device__ __forceinline__ int64_t qz_claim(unsigned long long *counter, bool want, int64_t total) {
    const int lane = threadIdx.x & 31;
    const unsigned wmask = __ballot_sync(QZ_FULL_MASK, want);
    if (!wmask) return -1;
    unsigned long long base = 0;
    const int leader = __ffs(wmask) - 1;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(wmask));
    base = __shfl_sync(QZ_FULL_MASK, base, leader);
    if (!want) return -1;
    const int64_t cand = (int64_t)base + __popc(wmask & ((1u << lane) - 1u));
    return cand < total ? cand : -1;
}

__device__ __forceinline__ int64_t qz_start_index(const QzRolloutArgs &a, int64_t r) {
    return a.state_index ? (int64_t)__ldg(a.state_index + r) : r / a.per_state;
}

// ---- phase 1: the wall phase ------------------------------------------------------------------------------
// While either player still owns a wall almost every ply draws a wall and runs the two flood fills of the
// path check.  Every lane of every warp of this kernel is in that phase, so the expensive code is executed
// by full warps; a rollout leaves the kernel as soon as no wall is left (or the game / the limit ends) and its
// state is parked in `mid` for phase 2.
//
// A rollout whose mover keeps drawing walls that fail the path check ("stuck": walls in hand, next to no
// legal placement) is under 1 % of the rollouts but cost 60 % of all path checks and, worse, a serial tail:
// hundreds of plies of up to 40 rejected draws each on ONE lane (profiles/r1b_rollout_wall_ncu_full.txt: SMs
// active 15 % of the kernel's duration).  Such a rollout is therefore EJECTED from this kernel after
// QZ_MAX_REJECTS failed draws in one ply and finished by qz_rollout_stuck_kernel, where a whole warp evaluates
// the draw attempts of a ply in parallel.
#ifndef QZ_MAX_REJECTS
#define QZ_MAX_REJECTS 2u
#endif

// The unit of work of one loop iteration is ONE DRAW ATTEMPT per lane (qz_sample.cuh), not one ply: a lane whose drawn
// wall failed the path check retries in the next iteration while its neighbours, whose draws were accepted, already play
// their next ply.  With a ply as the unit, a warp waited for its slowest lane -- and with 32 lanes some lane redraws on
// four plies out of five (profiles/r2j_wave_ncu_full.txt: 16 active lanes per instruction).  The attempt sequence, the
// "already known to block" sets and the ejection rule are exactly those of qz_sample_action_capped.
__global__ void __launch_bounds__(128, 4) qz_rollout_wall_kernel(QzRolloutArgs a) {
    QzState s;
    QzRng rng;
    int64_t r = -1;
    int steps = 0;
    uint32_t j = 0, n_bad = 0;                                          // attempt of the current ply, walls found to block
    uint64_t bad_h = 0, bad_v = 0;
    bool exhausted = false;
    s.H = s.V = s.meta = 0;
    rng = qz_rng_init(0, 0);
    for (;;) {
        const bool want = (r < 0) && !exhausted;
        const int64_t got = qz_claim(a.counter, want, a.n_rollouts);
        if (want) {
            if (got >= 0) {
                r = got;
                s = qz_load_state(a.states + qz_start_index(a, r));
                rng = qz_rng_init(a.seed, a.rids ? __ldg(a.rids + r) : a.rid_base + (uint64_t)r);
                steps = 0;
                j = 0; n_bad = 0; bad_h = bad_v = 0;
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(QZ_FULL_MASK, r < 0)) break;
        if (r >= 0) {
            bool leave = qz_done(s.meta) || steps >= a.limit - 1 || (qz_w1(s.meta) + qz_w2(s.meta)) == 0;
            if (!leave) {
                const QzPawnCtx c = qz_ctx_build(s.H, s.V);
                const uint32_t pmask = qz_mover_pawn_moves_ctx(c, s.meta);
                uint64_t hc = 0, vc = 0;
                if (qz_mover_walls(s.meta) > 0) { hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V); }
                const int npawn = qz_popc32(pmask), nh = qz_popc64(hc), nv = qz_popc64(vc);
                const uint32_t M = (uint32_t)(npawn + nh + nv);
                bool accept = false, stale = M == 0, eject = false;
                int act = -1;
                if (!stale) {
                    const uint32_t word = qz_attempt_word(rng, (uint32_t)steps, j);
                    act = qz_superset_action(pmask, hc, vc, npawn, nh, (int)qz_mulhi32(word, M));
                    if (act < 12) {
                        accept = true;                                  // a pawn move of the superset is legal as it stands
                    } else {
                        const bool vert = act >= 76;
                        const int ix = vert ? act - 76 : act - 12;
                        const uint64_t bit = 1ull << ix;
                        if (!((vert ? bad_v : bad_h) & bit)) {          // (drawn again: still illegal, next attempt)
                            const QzSweep w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta));
                            if (qz_wall_keeps_paths(w, ix, vert)) {
                                accept = true;
                            } else {
                                if (vert) bad_v |= bit; else bad_h |= bit;
                                n_bad++;
                                if (npawn == 0 && n_bad == M) stale = true;            // every candidate is a blocking wall
                                else if (n_bad > QZ_MAX_REJECTS) eject = true;
                            }
                        }
                    }
                }
                if (accept) {
                    s = qz_apply(s, act);
                    steps++;
                    j = 0; n_bad = 0; bad_h = bad_v = 0;
                } else if (eject) {                                     // stuck: hand over to qz_rollout_stuck_kernel
                    const unsigned long long k = atomicAdd(a.counter + 3, 1ull);
                    a.stuck_list[k] = (int32_t)r;
                    s.meta |= (uint64_t)QZ_FLAG_PENDING << 40;
                    leave = true;
                } else if (stale) {
                    s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40;
                    leave = true;
                } else {
                    j++;
                }
            }
            if (leave) {
                qz_store_state(a.mid + r, s);
                r = -1;
            }
        }
    }
}

// ---- phase 1b: stuck rollouts, one WARP per rollout -----------------------------------------------------------
// Every lane tracks the same state redundantly (no broadcasts).  The draw sequence of qz_sample.cuh is made of
// independent attempts, so the warp evaluates 32 of them at once: lane i decodes attempt 32*round + i.  A pawn
// move is legal as it stands, hence only the wall attempts BEFORE the first pawn attempt of the round can matter;
// exactly those need a verdict and the earliest legal attempt wins -- the same action the per-lane path would take.
//
// Legality memo.  Whether a wall keeps both paths (quoridor.py:463-477) is a pure function of (H, V, p1, p2), and a
// stuck rollout plays hundreds of plies between two wall placements while the two pawns wander over a few dozen tile
// pairs (host simulation: 270 wall-holding plies per wall configuration on 60 distinct pairs, 78 % of the plies
// revisit a pair).  Per wall configuration ("epoch") the warp keeps a small table in shared memory, keyed by
// (p1, p2): the candidates KNOWN legal and KNOWN blocking.  A ply first looks its pair up; a path check (two flood
// fills, ~1000 instructions of latency for the whole warp) only runs when an attempt that can still win is unknown,
// and then ALL 32 lanes check something: the lanes without an unknown attempt of their own take further unknown
// candidates of the pair, so two or three flood rounds settle a pair for good.  Flood rounds per wall-holding ply:
// 1.50 -> 0.49.  Everything else a ply needs is per-epoch data parked in shared memory: the byte-per-tile move table
// (qz_tile_table, as in the pawn kernel), the corner masks, and the ordered wall part of the superset.
#ifndef QZ_STUCK_THREADS
#define QZ_STUCK_THREADS 32
#endif
#define QZ_MEMO_SLOTS 128              // per warp; 2-way: a pair lives in slot h or h+1

struct QzStuckSmem {                   // one per warp
    uint32_t ctx[QZ_CTX_WORDS];        // corner masks of the epoch's walls (words 0-11 = the four move masks)
    uint32_t tile[QZ_TILE_TABLE_WORDS];   // byte t = move / corner info of tile t (qz_tile_table)
    uint8_t wall_tab[128];             // the epoch's precheck-passing walls in superset order (H by ix, then V by ix)
    uint32_t key[QZ_MEMO_SLOTS];       // epoch << 14 | p1 << 7 | p2 (0 = never used)
    uint64_t kl_h[QZ_MEMO_SLOTS], kl_v[QZ_MEMO_SLOTS];   // known legal
    uint64_t ki_h[QZ_MEMO_SLOTS], ki_v[QZ_MEMO_SLOTS];   // known blocking
};

// MIN_BLOCKS = resident one-warp blocks per SM the register budget is cut for (16 -> 128 registers, 20 -> 96, 24 -> 80):
// the pass is latency-bound with one warp per rollout, so residency is throughput as long as the spills stay small.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(QZ_STUCK_THREADS, MIN_BLOCKS) qz_rollout_stuck_kernel(QzRolloutArgs a) {
    __shared__ QzStuckSmem s_all[QZ_STUCK_THREADS / 32];
    QzStuckSmem &sm = s_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint8_t *tile_bytes = reinterpret_cast<const uint8_t *>(sm.tile);
    for (int i = lane; i < QZ_MEMO_SLOTS; i += 32) sm.key[i] = 0;
    __syncwarp();
    uint32_t epoch = 0;                                                  // bumped whenever the walls change
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(a.counter + 4, 1ull);
        k = __shfl_sync(QZ_FULL_MASK, k, 0);
        if (k >= a.counter[3]) break;
        const int64_t r = a.stuck_list[k];
        QzState s = qz_load_state(a.mid + r);
        s.meta &= ~((uint64_t)QZ_FLAG_PENDING << 40);
        const uint64_t m0 = __ldg(reinterpret_cast<const uint64_t *>(a.states + qz_start_index(a, r)) + 2);
        int steps = (int)qz_ply(s.meta) - (int)qz_ply(m0);
        QzRng rng = qz_rng_init(a.seed, a.rids ? __ldg(a.rids + r) : a.rid_base + (uint64_t)r);
        uint64_t ep_h = 0, ep_v = 0, hc = 0, vc = 0;
        int nh = 0, nv = 0;
        bool ep_valid = false;
        for (;;) {
            if (qz_done(s.meta) || steps >= a.limit - 1 || (qz_w1(s.meta) + qz_w2(s.meta)) == 0) break;
            if (!(ep_valid && ep_h == s.H && ep_v == s.V)) {
                // new wall configuration: park its masks, move table and wall candidates; old memo entries die with
                // the epoch number
                const QzPawnCtx c = qz_ctx_build(s.H, s.V);
                hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V);
                nh = qz_popc64(hc); nv = qz_popc64(vc);
                __syncwarp();
                if (lane == 0) { qz_ctx_store(c, sm.ctx); qz_tile_table(c, sm.tile, 1); }
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int ix = lane + 32 * q;
                    if ((hc >> ix) & 1ull) sm.wall_tab[__popcll(hc & ((1ull << ix) - 1ull))] = (uint8_t)(12 + ix);
                    if ((vc >> ix) & 1ull) sm.wall_tab[nh + __popcll(vc & ((1ull << ix) - 1ull))] = (uint8_t)(76 + ix);
                }
                __syncwarp();
                ep_h = s.H; ep_v = s.V; ep_valid = true;
                epoch = (epoch + 1) & 0x3FFFFu;
                if (epoch == 0) {                                        // tag wrapped: forget everything
                    for (int i = lane; i < QZ_MEMO_SLOTS; i += 32) sm.key[i] = 0;
                    __syncwarp();
                    epoch = 1;
                }
            }
            const int cur = qz_cur(s.meta), p1 = qz_p1(s.meta), p2 = qz_p2(s.meta);
            const int L = cur == 1 ? p1 : p2, O = cur == 1 ? p2 : p1;
            const uint32_t iL = tile_bytes[L], iO = tile_bytes[O];
            uint32_t hO = 0;
            if (qz_pawn_contact(iL, L, O) & 0xCu) {                      // east / west contact: the opponent's H corners matter
                const uint32_t *hm = sm.ctx + 24 + (O >> 5);
                const int sh = O & 31;
                hO = ((hm[0] >> sh) & 1u) | (((hm[3] >> sh) & 1u) << 1) | (((hm[6] >> sh) & 1u) << 2) | (((hm[9] >> sh) & 1u) << 3);
            }
            const uint32_t pawn = qz_pawn_moves_info(iL, iO, hO, L, O, cur);     // quoridor.py:272-353
            const int npawn = qz_popc32(pawn);
            const bool has_walls = qz_mover_walls(s.meta) > 0;
            int act = -1;
            if (!has_walls) {                                            // only pawn moves: attempt 0 is legal
                if (npawn) act = qz_nth_bit64((uint64_t)pawn, (int)qz_mulhi32(qz_attempt_word(rng, (uint32_t)steps, 0), (uint32_t)npawn));
            } else if (npawn + nh + nv != 0) {
                const uint32_t M = (uint32_t)(npawn + nh + nv);
                // ---- memo lookup (all lanes compute the same slot; lane 0 writes) ----
                const uint32_t key = (epoch << 14) | ((uint32_t)p1 << 7) | (uint32_t)p2;
                int slot = (p1 * 5 + p2 * 11) & (QZ_MEMO_SLOTS - 1);
                const int alt = (slot + 1) & (QZ_MEMO_SLOTS - 1);
                const uint32_t k0 = sm.key[slot], k1 = sm.key[alt];
                uint64_t klh = 0, klv = 0, kih = 0, kiv = 0;
                if (k0 == key) {
                    klh = sm.kl_h[slot]; klv = sm.kl_v[slot]; kih = sm.ki_h[slot]; kiv = sm.ki_v[slot];
                } else if (k1 == key) {
                    slot = alt;
                    klh = sm.kl_h[slot]; klv = sm.kl_v[slot]; kih = sm.ki_h[slot]; kiv = sm.ki_v[slot];
                } else {
                    if ((k0 >> 14) == epoch && (k1 >> 14) != epoch) slot = alt;      // prefer a slot of a dead epoch
                    __syncwarp();
                    if (lane == 0) { sm.key[slot] = key; sm.kl_h[slot] = 0; sm.kl_v[slot] = 0; sm.ki_h[slot] = 0; sm.ki_v[slot] = 0; }
                }
                __syncwarp();
                bool dirty = false, prepared = false;
                QzSweep w;
                for (uint32_t round = 0; act < 0; round++) {
                    const uint32_t word = qz_attempt_word(rng, (uint32_t)steps, round * 32 + lane);
                    const int kk = (int)qz_mulhi32(word, M);
                    const int cand = kk < npawn ? qz_nth_bit64((uint64_t)pawn, kk) : (int)sm.wall_tab[kk - npawn];
                    const unsigned pawn_lanes = __ballot_sync(QZ_FULL_MASK, cand < 12);
                    const int first_pawn = pawn_lanes ? __ffs(pawn_lanes) - 1 : 32;
                    const bool vert = cand >= 76;
                    const int ix = vert ? cand - 76 : cand - 12;        // garbage on pawn lanes, masked by `mine`
                    const uint64_t bit = 1ull << (ix & 63);
                    const bool mine = lane < first_pawn;                 // a wall attempt that can still win the draw
                    const unsigned known_ok = __ballot_sync(QZ_FULL_MASK, mine && ((vert ? klv : klh) & bit));
                    const int limit = known_ok ? __ffs(known_ok) - 1 : first_pawn;   // attempts after a sure winner are moot
                    const bool need = lane < limit && !((vert ? kiv : kih) & bit);  // unknown and still able to win
                    const unsigned need_lanes = __ballot_sync(QZ_FULL_MASK, need);
                    if (need_lanes) {
                        // one flood round for the whole warp: the `need` lanes check their own attempt, every other
                        // lane takes one more unknown candidate of this pair
                        uint64_t uh = hc & ~(klh | kih), uv = vc & ~(klv | kiv);
                        uh &= ~qz_warp_or64(need && !vert ? bit : 0ull);
                        uv &= ~qz_warp_or64(need && vert ? bit : 0ull);
                        const int nuh = qz_popc64(uh), nu = nuh + qz_popc64(uv);
                        const int rank = __popc(~need_lanes & ((1u << lane) - 1u));
                        bool chk = need, cvert = vert;
                        int cix = ix & 63;
                        if (!need && rank < nu) {
                            chk = true;
                            cvert = rank >= nuh;
                            cix = cvert ? qz_nth_bit64(uv, rank - nuh) : qz_nth_bit64(uh, rank);
                        }
                        if (!prepared) {
                            w.H = s.H; w.V = s.V; w.p1 = p1; w.p2 = p2;
                            w.dirs.n = bb_make(sm.ctx[0], sm.ctx[1], sm.ctx[2]); w.dirs.s = bb_make(sm.ctx[3], sm.ctx[4], sm.ctx[5]);
                            w.dirs.e = bb_make(sm.ctx[6], sm.ctx[7], sm.ctx[8]); w.dirs.w = bb_make(sm.ctx[9], sm.ctx[10], sm.ctx[11]);
                            prepared = true;
                        }
                        const bool ok = chk && qz_wall_keeps_paths(w, cix, cvert);
                        const uint64_t cbit = 1ull << cix;
                        klh |= qz_warp_or64(chk && ok && !cvert ? cbit : 0ull);
                        klv |= qz_warp_or64(chk && ok && cvert ? cbit : 0ull);
                        kih |= qz_warp_or64(chk && !ok && !cvert ? cbit : 0ull);
                        kiv |= qz_warp_or64(chk && !ok && cvert ? cbit : 0ull);
                        dirty = true;
                    }
                    const unsigned ok_lanes = __ballot_sync(QZ_FULL_MASK, mine && ((vert ? klv : klh) & bit));
                    const int winner = ok_lanes ? __ffs(ok_lanes) - 1 : first_pawn;      // earliest legal attempt
                    if (winner < 32) { act = __shfl_sync(QZ_FULL_MASK, cand, winner); break; }
                    // all 32 attempts were blocking walls; with no pawn move at all, stop once every candidate is known
                    // to block (stalemate; the reference would return [] and crash)
                    if (npawn == 0 && kih == hc && kiv == vc) break;
                }
                if (dirty) {
                    __syncwarp();
                    if (lane == 0) { sm.kl_h[slot] = klh; sm.kl_v[slot] = klv; sm.ki_h[slot] = kih; sm.ki_v[slot] = kiv; }
                    __syncwarp();
                }
            }
            if (act < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; break; }
            s = qz_apply(s, act);
            steps++;
        }
        if (lane == 0) qz_store_state(a.mid + r, s);
    }
}

// ---- phase 2: the pawn phase ------------------------------------------------------------------------------
// No wall can be placed any more, so everything the move rules read about a tile is constant for the rest of the
// rollout.  When a lane takes a rollout it builds the twelve corner masks once and transposes eight of them into a
// byte-per-tile table in shared memory (qz_tile_table; the four "corner is horizontal" masks, only read on an
// east / west contact with the opponent, are parked as they are).  A ply is then: ONE table byte for the tile
// the pawn just moved to (the other pawn's byte is carried over from the previous ply), a handful of bit
// operations, one Philox word, the k-th set bit, an add -- about 90 instructions, whatever tiles the lanes stand
// on -- instead of the ~350 of the generic state machine (profiles/r1g_rollout_ncu_full.txt: this kernel was 56 %
// of all rollout instructions, 17.7 active lanes per instruction because refills ran one lane at a time).
// Refills are batched: idle lanes wait until QZ_PAWN_REFILL of them can set up together.
#define QZ_PAWN_THREADS 128
#define QZ_PAWN_REFILL 8

struct QzPawnSmem {
    uint32_t tile[QZ_PAWN_THREADS * QZ_TILE_TABLE_WORDS];   // thread-major: byte t of thread i at i*84 + t
    uint32_t hmask[12 * QZ_PAWN_THREADS];                   // word-major: neH, nwH, seH, swH x 3 words
    int32_t delta[12];
};

__device__ __forceinline__ int8_t qz_pawn_result(int winner, int player0) {       // pure_mcts.py:104-108
    return (int8_t)(winner == 0 ? 0 : (winner == player0 ? 1 : -1));
}

__global__ void __launch_bounds__(QZ_PAWN_THREADS, 8) qz_rollout_pawn_kernel(QzRolloutArgs a) {
    __shared__ QzPawnSmem sm;
    const int tid = threadIdx.x;
    if (tid < 12) sm.delta[tid] = qz_delta(tid);
    __syncthreads();
    const uint8_t *my_tile = reinterpret_cast<const uint8_t *>(sm.tile + tid * QZ_TILE_TABLE_WORDS);
    const int64_t total = a.count_in ? (int64_t)*a.count_in : a.n_rollouts;
    // Philox blocks: `cur` serves plies 4q..4q+3, `nxt` is block q+1.  Lanes cross block boundaries on different
    // iterations, so computing a block on demand would run the 10 rounds with a quarter of the lanes on almost
    // every iteration; instead every busy lane refreshes `nxt` on every fourth iteration of the (warp-uniform)
    // loop -- a lane plays one ply per iteration, hence crosses exactly one boundary in between.
    QzPhilox4 cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    uint64_t rid = 0;
    int64_t r = -1;
    int L = 0, O = 0, mover = 1, steps = 0, steps0 = 0, player0 = 0, left = 0;
    uint32_t iL = 0, iO = 0;
    unsigned long long my_plies = 0;
    bool exhausted = false;
    for (unsigned iter = 0;; iter++) {
        const bool want = (r < 0) && !exhausted;
        const unsigned want_lanes = __ballot_sync(QZ_FULL_MASK, want);
        const unsigned busy_lanes = __ballot_sync(QZ_FULL_MASK, r >= 0);
        if ((iter & 3u) == 0 && r >= 0) nxt = qz_philox(a.seed, rid, ((uint32_t)steps >> 2) + 1u, 0);
        if (want_lanes && (busy_lanes == 0 || __popc(want_lanes) >= QZ_PAWN_REFILL)) {
            // claim until every wanting lane holds a LIVE pawn-phase rollout (or the work is gone); rollouts that
            // already ended in the wall phase, or are parked for the stuck kernel, are settled on the spot
            bool need = want, fresh = false;
            uint64_t H = 0, V = 0;
            while (__any_sync(QZ_FULL_MASK, need)) {
                int64_t got = qz_claim(a.work, need, total);
                if (!need) continue;
                if (got >= 0 && a.list_in) got = a.list_in[got];
                if (got < 0) { exhausted = true; need = false; continue; }
                const QzState s = qz_load_state(a.mid + got);
                const uint64_t m0 = __ldg(reinterpret_cast<const uint64_t *>(a.states + qz_start_index(a, got)) + 2);
                const unsigned fl = qz_flags(s.meta);
                const int st = (int)qz_ply(s.meta) - (int)qz_ply(m0);
                if (fl & QZ_FLAG_PENDING) {                  // deferred: finished later by qz_rollout_finish
                    a.result[got] = (int8_t)-128;
                    continue;
                }
                if ((fl & (QZ_FLAG_DONE | QZ_FLAG_STALEMATE)) || st >= a.limit - 1 || (qz_w1(s.meta) + qz_w2(s.meta)) != 0 ||
                    !qz_on_board(s.meta)) {                  // walls left here only if phase 1 ended the rollout
                    a.result[got] = qz_pawn_result(qz_winner(s.meta), qz_cur(m0));
                    if (a.plies) a.plies[got] = st;
                    if (a.final_states) qz_store_state(a.final_states + got, s);
                    my_plies += (unsigned long long)st;
                    continue;
                }
                r = got; need = false; fresh = true;
                H = s.H; V = s.V;
                mover = qz_cur(s.meta);
                L = mover == 1 ? qz_p1(s.meta) : qz_p2(s.meta);
                O = mover == 1 ? qz_p2(s.meta) : qz_p1(s.meta);
                steps = steps0 = st;
                left = a.slice;
                player0 = qz_cur(m0);
                rid = a.rids ? __ldg(a.rids + got) : a.rid_base + (uint64_t)got;
            }
            if (fresh) {
                cur = qz_philox(a.seed, rid, (uint32_t)steps >> 2, 0);
                nxt = qz_philox(a.seed, rid, ((uint32_t)steps >> 2) + 1u, 0);
                const QzPawnCtx c = qz_ctx_build(H, V);
                qz_tile_table(c, sm.tile + tid * QZ_TILE_TABLE_WORDS, 1);
                uint32_t *hm = sm.hmask + tid;
                hm[0 * QZ_PAWN_THREADS] = c.neH.w0; hm[1 * QZ_PAWN_THREADS] = c.neH.w1; hm[2 * QZ_PAWN_THREADS] = c.neH.w2;
                hm[3 * QZ_PAWN_THREADS] = c.nwH.w0; hm[4 * QZ_PAWN_THREADS] = c.nwH.w1; hm[5 * QZ_PAWN_THREADS] = c.nwH.w2;
                hm[6 * QZ_PAWN_THREADS] = c.seH.w0; hm[7 * QZ_PAWN_THREADS] = c.seH.w1; hm[8 * QZ_PAWN_THREADS] = c.seH.w2;
                hm[9 * QZ_PAWN_THREADS] = c.swH.w0; hm[10 * QZ_PAWN_THREADS] = c.swH.w1; hm[11 * QZ_PAWN_THREADS] = c.swH.w2;
                iL = my_tile[L];
                iO = my_tile[O];
            }
        } else if (busy_lanes == 0) {
            break;
        }
        if (r >= 0) {
            // one ply (quoridor.py:146,159-186 with no wall left; pure_mcts.py:7-10,97-103)
            uint32_t hO = 0;
            if (qz_pawn_contact(iL, L, O) & 0xCu) {          // east / west contact: the opponent's H corners matter
                const uint32_t *hm = sm.hmask + (O >> 5) * QZ_PAWN_THREADS + tid;
                const int sh = O & 31;
                hO = ((hm[0] >> sh) & 1u) | (((hm[3 * QZ_PAWN_THREADS] >> sh) & 1u) << 1) |
                     (((hm[6 * QZ_PAWN_THREADS] >> sh) & 1u) << 2) | (((hm[9 * QZ_PAWN_THREADS] >> sh) & 1u) << 3);
            }
            uint32_t pm = qz_pawn_moves_info(iL, iO, hO, L, O, mover);
            const int np = __popc(pm);
            int winner = 0;
            unsigned add_flags = 0;
            bool finished = false;
            if (np == 0) {
                add_flags = QZ_FLAG_STALEMATE;
                finished = true;
            } else {
                const uint32_t word = qz_philox_word(cur, steps & 3);           // attempt 0 of ply `steps` (qz_sample.cuh)
                int k = (int)__umulhi(word, (uint32_t)np);                       // drop the k lowest moves
                if (k > 0) pm &= pm - 1;
                if (k > 1) pm &= pm - 1;
                if (k > 2) pm &= pm - 1;
                for (k -= 3; k > 0; k--) pm &= pm - 1;
                L += sm.delta[__ffs(pm) - 1];                                    // quoridor.py:217-243
                steps++;
                if ((steps & 3) == 0) cur = nxt;
                if (mover == 1 ? L > 71 : L < 9) {            // :193-202; the mover is not rotated on a win (:176-181)
                    winner = mover;
                    add_flags = QZ_FLAG_DONE | ((unsigned)winner << QZ_FLAG_WINNER_SHIFT);
                    finished = true;
                } else {
                    const int t = L; L = O; O = t;
                    const uint32_t moved = my_tile[O];
                    iL = iO; iO = moved;
                    mover = 3 - mover;
                    finished = steps >= a.limit - 1;
                    if (!finished && --left == 0) {
                        // end of this pass's slice: park the rollout; the next pass resumes it among the other survivors,
                        // packed into full warps again (only the meta word changes in the pawn phase)
                        uint64_t *mw = reinterpret_cast<uint64_t *>(a.mid + r) + 2;
                        const uint64_t m = *mw;
                        const unsigned ply = qz_ply(m) + (unsigned)(steps - steps0);
                        *mw = qz_pack_meta(mover == 1 ? L : O, mover == 1 ? O : L, 0, 0, mover, qz_flags(m),
                                           ply < 0xFFFFu ? ply : 0xFFFFu);
                        a.list_out[atomicAdd(a.count_out, 1ull)] = (int32_t)r;
                        r = -1;
                    }
                }
            }
            if (finished && r >= 0) {
                a.result[r] = qz_pawn_result(winner, player0);
                if (a.plies) a.plies[r] = steps;
                if (a.final_states) {
                    QzState s = qz_load_state(a.mid + r);
                    const unsigned ply = qz_ply(s.meta) + (unsigned)(steps - steps0);
                    s.meta = qz_pack_meta(mover == 1 ? L : O, mover == 1 ? O : L, 0, 0, mover, qz_flags(s.meta) | add_flags,
                                          ply < 0xFFFFu ? ply : 0xFFFFu);
                    qz_store_state(a.final_states + r, s);
                }
                my_plies += (unsigned long long)steps;
                r = -1;
            }
        }
    }
    // cumulative env-step counter (workspace word 1; never zeroed by the library)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) my_plies += __shfl_xor_sync(QZ_FULL_MASK, my_plies, off);
    if ((tid & 31) == 0 && my_plies) atomicAdd(a.counter + 1, my_plies);
}

// workspace: header | mid[n] | stuck_list[n] | two survivor lists [n] (ping-pong between pawn passes)
#define QZ_WS_COUNTERS 64
#define QZ_WS_HEADER_BYTES (QZ_WS_COUNTERS * 8)
#define QZ_PAWN_SLICE 128
#define QZ_PAWN_MAX_PASSES ((QZ_WS_COUNTERS - 8) / 2)
static int64_t qz_list_bytes(int64_t n) { return ((n * 4 + 7) / 8) * 8; }

extern "C" int64_t qz_rollout_workspace_bytes(int64_t n_rollouts) {
    const int64_t n = n_rollouts > 0 ? n_rollouts : 0;
    return QZ_WS_HEADER_BYTES + n * (int64_t)sizeof(qz_state) + 3 * qz_list_bytes(n);
}

static int qz_persistent_blocks(const void *kernel, int64_t n_items, int items_per_block, int threads = 128) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
    int64_t blocks = (int64_t)sms * per_sm;                 // one resident wave: a multiple of the SM count
    const int64_t needed = (n_items + items_per_block - 1) / items_per_block;
    return (int)(blocks < needed ? blocks : needed);
}

// The pawn phase as a sequence of passes.  A rollout plays at most slices[j] plies in pass j and is then parked; the next
// pass picks the survivors up from a list, so they sit in full warps again instead of keeping a few lanes of
// every warp alive for up to `limit` iterations (rollout lengths are heavy-tailed: mean ~300 plies, cap 1000).
// The number of survivors is only known on the device: every pass launches a resident grid whose blocks exit at
// once when there is nothing to claim.  `finish` = the deferred pass over stuck_list.
// Slice schedule: three slices of QZ_PAWN_SLICE plies, where re-packing the survivors pays (most rollouts end there), then
// one of twice that, then whatever is left in a single pass -- the late passes hold a few per cent of the rollouts and
// are bound by the latency of one ply, so every further launch only added its ramp (profiles/inst_table_summary.txt:
// passes 4-8 of the uniform schedule ran at 11-20 % of the issue rate).  Returns the number of passes; slices[j] = plies
// a rollout may play in pass j.
static int qz_pawn_schedule(int64_t limit, int32_t *slices) {
    int64_t left = limit > 1 ? limit - 1 : 1;                           // a rollout never plays more plies than this
    int n = 0;
    const int64_t plan[4] = {QZ_PAWN_SLICE, QZ_PAWN_SLICE, QZ_PAWN_SLICE, 2 * QZ_PAWN_SLICE};
    while (left > 0 && n < 4) {
        const int64_t sl = plan[n] < left ? plan[n] : left;
        slices[n++] = (int32_t)sl;
        left -= sl;
    }
    if (left > 0) slices[n++] = (int32_t)(left < 0x7FFFFFFF ? left : 0x7FFFFFFF);
    return n;
}

extern "C" int32_t qz_rollout_pawn_passes(int32_t limit) {
    int32_t slices[8];
    return qz_pawn_schedule(limit, slices);
}

static int qz_pawn_passes(QzRolloutArgs a, bool finish, cudaStream_t st, const char *what) {
    int32_t slices[8];
    const int passes = qz_pawn_schedule(a.limit, slices);
    int32_t *lists[2] = {a.stuck_list + qz_list_bytes(a.n_rollouts) / 4, a.stuck_list + 2 * (qz_list_bytes(a.n_rollouts) / 4)};
    int blocks = qz_persistent_blocks((const void *)qz_rollout_pawn_kernel, a.n_rollouts, QZ_PAWN_THREADS);
    if (finish) blocks = blocks / 4 > 0 ? blocks / 4 : 1;              // the deferred list is well under 1 % of the rollouts
    for (int j = 0; j < passes; j++) {
        a.slice = slices[j];
        a.work = a.counter + 8 + 2 * j;
        a.count_out = a.counter + 9 + 2 * j;
        a.list_out = lists[j & 1];
        if (j == 0) {
            a.list_in = finish ? a.stuck_list : nullptr;
            a.count_in = finish ? a.counter + 3 : nullptr;
        } else {
            a.list_in = lists[(j - 1) & 1];
            a.count_in = a.counter + 9 + 2 * (j - 1);
        }
        qz_rollout_pawn_kernel<<<blocks, QZ_PAWN_THREADS, 0, st>>>(a);
        const int rc = qz_check_launch(what);
        if (rc) return rc;
    }
    return 0;
}

// QZ_STUCK_OCC (environment, read once): 16 | 20 | 24 resident warps per SM for the stuck pass -- an A/B knob; the default
// is the measured best (see profiles/).
#ifndef QZ_STUCK_OCC_DEFAULT
#define QZ_STUCK_OCC_DEFAULT 16
#endif
static int qz_launch_stuck(const QzRolloutArgs &a, cudaStream_t st, const char *what) {
    static const int occ = [] { const char *e = getenv("QZ_STUCK_OCC"); return e ? atoi(e) : QZ_STUCK_OCC_DEFAULT; }();
    const int per_block = QZ_STUCK_THREADS / 32;
    if (occ >= 24)
        qz_rollout_stuck_kernel<24><<<qz_persistent_blocks((const void *)qz_rollout_stuck_kernel<24>, a.n_rollouts, per_block, QZ_STUCK_THREADS), QZ_STUCK_THREADS, 0, st>>>(a);
    else if (occ >= 20)
        qz_rollout_stuck_kernel<20><<<qz_persistent_blocks((const void *)qz_rollout_stuck_kernel<20>, a.n_rollouts, per_block, QZ_STUCK_THREADS), QZ_STUCK_THREADS, 0, st>>>(a);
    else
        qz_rollout_stuck_kernel<16><<<qz_persistent_blocks((const void *)qz_rollout_stuck_kernel<16>, a.n_rollouts, per_block, QZ_STUCK_THREADS), QZ_STUCK_THREADS, 0, st>>>(a);
    return qz_check_launch(what);
}

static int qz_rollout_args(QzRolloutArgs &a, const qz_state *states, int64_t n_states, const int32_t *state_index,
                           int32_t per_state, int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids,
                           int32_t limit, int8_t *result, int32_t *plies, qz_state *final_states, void *workspace,
                           const char *fn) {
    if (!(n_rollouts >= 0 && n_states >= 0 && limit >= 1)) return qz_fail(QZ_E_RANGE, "%s: bad sizes", fn);
    if (n_rollouts == 0) return 0;
    if (!states || !result || !workspace) return qz_fail(QZ_E_NULL, "%s: states / result / workspace is NULL", fn);
    if (((uintptr_t)states | (uintptr_t)workspace | (uintptr_t)final_states) % 8)
        return qz_fail(QZ_E_ALIGN, "%s: states / workspace / final_states not 8-byte aligned", fn);
    if (state_index == nullptr && !(per_state >= 1 && n_rollouts <= n_states * (int64_t)per_state))
        return qz_fail(QZ_E_RANGE, "%s: per_state does not cover n_rollouts", fn);
    a.states = states; a.state_index = state_index; a.rids = rids; a.rid_base = rid_base; a.seed = seed;
    a.n_rollouts = n_rollouts; a.per_state = per_state > 0 ? per_state : 1; a.limit = limit;
    a.result = result; a.plies = plies; a.final_states = final_states;
    a.counter = (unsigned long long *)workspace;
    a.mid = (qz_state *)((char *)workspace + QZ_WS_HEADER_BYTES);
    a.stuck_list = (int32_t *)((char *)workspace + QZ_WS_HEADER_BYTES + n_rollouts * (int64_t)sizeof(qz_state));
    a.list_in = nullptr; a.count_in = nullptr; a.list_out = nullptr; a.count_out = nullptr; a.work = nullptr; a.slice = 0;
    return 0;
}

extern "C" int qz_rollout(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
                          int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
                          int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, int32_t flags,
                          void *stream) {
    QzRolloutArgs a;
    int rc = qz_rollout_args(a, states, n_states, state_index, per_state, n_rollouts, seed, rid_base, rids, limit, result,
                             plies, final_states, workspace, "qz_rollout");
    if (rc || n_rollouts == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(a.counter, 0, 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(a.counter + 2, 0, (QZ_WS_COUNTERS - 2) * 8, st);
    if (e != cudaSuccess) return qz_fail((int)e, "qz_rollout: memset: %s", cudaGetErrorString(e));
    qz_rollout_wall_kernel<<<qz_persistent_blocks((const void *)qz_rollout_wall_kernel, n_rollouts, 128), 128, 0, st>>>(a);
    rc = qz_check_launch("qz_rollout (wall phase)");
    if (rc) return rc;
    if (!(flags & QZ_ROLLOUT_DEFER_STUCK)) {
        // the number of ejected rollouts is only known on the device: launch a resident grid, blocks exit when the list is empty
        rc = qz_launch_stuck(a, st, "qz_rollout (stuck phase)");
        if (rc) return rc;
    }
    return qz_pawn_passes(a, false, st, "qz_rollout (pawn phase)");
}

// The deferred pass of qz_rollout(..., QZ_ROLLOUT_DEFER_STUCK): same arguments and buffers.
extern "C" int qz_rollout_finish(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
                                 int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
                                 int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, void *stream) {
    QzRolloutArgs a;
    int rc = qz_rollout_args(a, states, n_states, state_index, per_state, n_rollouts, seed, rid_base, rids, limit, result,
                             plies, final_states, workspace, "qz_rollout_finish");
    if (rc || n_rollouts == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(a.counter + 4, 0, 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(a.counter + 8, 0, (QZ_WS_COUNTERS - 8) * 8, st);
    if (e != cudaSuccess) return qz_fail((int)e, "qz_rollout_finish: memset: %s", cudaGetErrorString(e));
    rc = qz_launch_stuck(a, st, "qz_rollout_finish (stuck phase)");
    if (rc) return rc;
    return qz_pawn_passes(a, true, st, "qz_rollout_finish (pawn phase)");
}
