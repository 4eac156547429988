// Host simulation behind the memoised stuck-rollout kernel (csrc/qz_rollout.cu): plays random rollouts with
// the product's own rules header and draw procedure, ejects the "stuck" ones exactly as the wall-phase kernel does, and counts for
// them (a) how often a wall-holding ply revisits a pawn pair (p1, p2) within one wall configuration, (b) flood rounds and
// candidate checks per ply of the plain warp-per-rollout scheme against a per-pair known-legal / known-blocking memo whose
// idle lanes fill further unknown candidates during a round.  Build and run (CPU only):
//   g++ -O2 -std=c++17 -Wno-unknown-pragmas -Ialphazero_quoridor_b200/csrc -o /tmp/stuck_memo_sim tools/stuck_memo_sim.cpp && /tmp/stuck_memo_sim 30000 12
// Output quoted in DESIGN.md: pair hit rate 0.775, flood rounds 1.50 -> 0.49 and candidate checks 24.3 -> 5.3 per wall-holding ply.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>
#include "qz_rules.cuh"
#include "qz_philox.cuh"
#include "qz_sample.cuh"

struct Memo { uint64_t kl_h=0, kl_v=0, ki_h=0, ki_v=0; };

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 20000;
    int pre = argc > 2 ? atoi(argv[2]) : 12;      // random plies before the rollout starts (leaf depth proxy)
    uint64_t seed = 12345;
    long long total_rollouts = 0, stuck_rollouts = 0;
    long long plies_all = 0, plies_stuck = 0, exp_plies = 0;        // exp = mover has walls (needs rounds)
    long long rounds_now = 0, rounds_memo = 0, checks_now = 0, checks_memo = 0, pair_hits = 0, full_known = 0;
    long long epochs = 0, fill_checks = 0; long long distinct_pairs = 0;
    std::vector<int> stuck_len;
    for (int r = 0; r < n; r++) {
        QzState s = qz_initial_state();
        QzRng rng0 = qz_rng_init(seed, 1000000 + r);
        bool ok = true;
        for (int i = 0; i < pre; i++) { int a = qz_sample_action(s, rng0, i); if (a < 0 || qz_done(s.meta)) { ok = false; break; } s = qz_apply(s, a); }
        if (!ok || qz_done(s.meta)) continue;
        total_rollouts++;
        QzRng rng = qz_rng_init(seed, r);
        int steps = 0; bool stuck = false;
        std::map<int, Memo> memo; uint64_t eh = ~0ull, ev = ~0ull;
        long long my_stuck_plies = 0;
        while (!qz_done(s.meta) && steps < 999) {
            if ((qz_w1(s.meta) + qz_w2(s.meta)) == 0 && !stuck) break;   // pawn phase: not interesting
            if ((qz_w1(s.meta) + qz_w2(s.meta)) == 0) break;
            int act;
            if (!stuck) {
                act = qz_sample_action_capped(s, rng, steps, 2);
                if (act == -2) { stuck = true; stuck_rollouts++; continue; }
            } else {
                // replicate the warp kernel: rounds of 32 attempts
                const QzPawnCtx c = qz_ctx_build(s.H, s.V);
                const uint32_t pawn = qz_mover_pawn_moves_ctx(c, s.meta);
                const bool has_walls = qz_mover_walls(s.meta) > 0;
                uint64_t hc = has_walls ? qz_hcand(s.H, s.V) : 0, vc = has_walls ? qz_vcand(s.H, s.V) : 0;
                int npawn = qz_popc32(pawn), nh = qz_popc64(hc), nv = qz_popc64(vc);
                uint32_t M = npawn + nh + nv;
                act = -1;
                my_stuck_plies++;
                if (M && !has_walls) act = qz_nth_bit64(pawn, qz_mulhi32(qz_attempt_word(rng, steps, 0), M));
                else if (M) {
                    exp_plies++;
                    if (eh != s.H || ev != s.V) { memo.clear(); eh = s.H; ev = s.V; epochs++; }
                    int key = qz_p1(s.meta) * 128 + qz_p2(s.meta);
                    bool seen = memo.count(key);
                    if (seen) pair_hits++; else distinct_pairs++;
                    Memo &mm = memo[key];
                    QzSweep w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta));
                    uint64_t bad_h = 0, bad_v = 0;
                    bool any_round_memo = false;
                    for (uint32_t round = 0; act < 0; round++) {
                        int cand[32]; int first_pawn = 32;
                        for (int l = 0; l < 32; l++) {
                            uint32_t word = qz_attempt_word(rng, steps, round * 32 + l);
                            cand[l] = qz_superset_action(pawn, hc, vc, npawn, nh, qz_mulhi32(word, M));
                            if (cand[l] < 12 && first_pawn == 32) first_pawn = l;
                        }
                        int winner = first_pawn; bool now_any = false, memo_any = false; int uniq_unknown = 0;
                        bool okl[32];
                        for (int l = 0; l < first_pawn; l++) {
                            bool vert = cand[l] >= 76; int ix = vert ? cand[l] - 76 : cand[l] - 12; uint64_t bit = 1ull << ix;
                            okl[l] = false;
                            bool legal = qz_wall_keeps_paths(w, ix, vert);
                            if (!((vert ? bad_v : bad_h) & bit)) { now_any = true; checks_now++; }
                            bool known = ((vert ? mm.kl_v | mm.ki_v : mm.kl_h | mm.ki_h) & bit) != 0;
                            if (!known) { memo_any = true; checks_memo++; uniq_unknown++; if (legal) { if (vert) mm.kl_v |= bit; else mm.kl_h |= bit; } else { if (vert) mm.ki_v |= bit; else mm.ki_h |= bit; } }
                            okl[l] = legal;
                        }
                        if (now_any) rounds_now++;
                        if (memo_any) {
                            rounds_memo++; any_round_memo = true;
                            // opportunistic fill: idle lanes check other unknown candidates during the same flood
                            int idle = 32 - uniq_unknown;
                            for (int ix = 0; ix < 64 && idle > 0; ix++) {
                                for (int vert = 0; vert < 2 && idle > 0; vert++) {
                                    uint64_t bit = 1ull << ix;
                                    if (!(((vert ? vc : hc)) & bit)) continue;
                                    if ((vert ? mm.kl_v | mm.ki_v : mm.kl_h | mm.ki_h) & bit) continue;
                                    bool legal = qz_wall_keeps_paths(w, ix, vert);
                                    if (legal) { if (vert) mm.kl_v |= bit; else mm.kl_h |= bit; } else { if (vert) mm.ki_v |= bit; else mm.ki_h |= bit; }
                                    idle--; fill_checks++;
                                }
                            }
                        }
                        for (int l = 0; l < first_pawn; l++) if (okl[l]) { winner = l; break; }
                        if (winner < 32) { act = cand[winner]; break; }
                        for (int l = 0; l < 32; l++) { bool vert = cand[l] >= 76; int ix = vert ? cand[l]-76 : cand[l]-12; if (vert) bad_v |= 1ull<<ix; else bad_h |= 1ull<<ix; }
                        if (npawn == 0 && bad_h == hc && bad_v == vc) break;
                    }
                    if (!any_round_memo) full_known++;
                }
            }
            if (act < 0) break;
            s = qz_apply(s, act); steps++;
        }
        plies_all += steps; plies_stuck += my_stuck_plies;
        if (stuck) stuck_len.push_back((int)my_stuck_plies);
    }
    printf("rollouts %lld stuck %lld (%.3f%%)\n", total_rollouts, stuck_rollouts, 100.0 * stuck_rollouts / total_rollouts);
    printf("stuck plies %lld (per stuck rollout %.1f), expensive plies %lld, epochs %lld (exp plies/epoch %.1f)\n", plies_stuck,
           (double)plies_stuck / (stuck_rollouts ? stuck_rollouts : 1), exp_plies, epochs, (double)exp_plies / (epochs ? epochs : 1));
    printf("pair hit rate %.3f (distinct pairs/epoch %.1f)\n", (double)pair_hits / exp_plies, (double)distinct_pairs / (epochs ? epochs : 1));
    printf("rounds with a flood: now %lld (%.2f/exp ply) memo %lld (%.2f/exp ply); plies with no flood at all under memo %.3f\n", rounds_now,
           (double)rounds_now / exp_plies, rounds_memo, (double)rounds_memo / exp_plies, (double)full_known / exp_plies);
    printf("candidate checks: now %lld (%.1f/exp ply) memo %lld (%.1f/exp ply)\n", checks_now, (double)checks_now / exp_plies, checks_memo,
           (double)checks_memo / exp_plies);
    printf("fill checks %lld\n", fill_checks);
    return 0;
}
