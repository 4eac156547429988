/*
 * quoridor_oracle.c -- TEST INFRASTRUCTURE ONLY (the CPU oracle).
 *
 * A plain-C restatement of the reference's self-play path (cryer/AlphaZero_Quoridor:
 * quoridor.py, mcts.py, pure_mcts.py), written literally -- one C function per reference
 * method, same control flow, same data shapes (a 64-entry intersection array holding
 * +1/-1/0, a FIFO breadth-first search that calls the pawn-move generator per dequeued tile,
 * dict-ordered children, first-max tie breaks).  It deliberately shares NO code and NO data
 * layout with the product (alphazero_quoridor_b200/csrc uses bitboards + flood fill), so
 * agreement between the two is evidence, not tautology.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product never links, imports or calls anything in oracle/.
 *
 * Parity pin: the reference ships no tests, so this file is pinned against
 * fixtures produced by running the UNMODIFIED Python reference in the build container
 * (tests/golden/gen_golden.py -> tests/golden/ *.json[.gz]); tests/test_oracle_golden.py checks
 * every fixture.  Each function cites the reference file:line it follows.
 *
 * Undefined-in-the-reference cases: calling actions()/state() on a state
 * where a pawn has left the board raises IndexError (P1) or silently wraps (P2) in Python.
 * Here any out-of-range intersection read sets `oq_index_error` and yields 0; callers treat
 * such states as "reference undefined".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define OQ_H 1
#define OQ_V (-1)

typedef struct {
    int positions[3];           /* [1], [2]      quoridor.py:34-37 */
    int8_t intersections[64];   /* +1 H, -1 V    quoridor.py:49-53 */
    int walls_remaining[3];     /* [1], [2]      quoridor.py:55-56 */
    int current_player;         /*               quoridor.py:27    */
    int last_player;            /*               quoridor.py:28    */
} oq_game;

static _Thread_local int oq_index_error = 0;

int oq_sizeof_game(void) { return (int)sizeof(oq_game); }
int oq_get_index_error(void) { return oq_index_error; }
void oq_clear_index_error(void) { oq_index_error = 0; }

/* quoridor.py:26-56 */
void oq_reset(oq_game *g) {
    memset(g, 0, sizeof(*g));
    g->current_player = 1;
    g->last_player = -1;
    g->positions[1] = 4;
    g->positions[2] = 76;
    g->walls_remaining[1] = 10;
    g->walls_remaining[2] = 10;
}

/* numpy indexing with Python's negative wrap; out of range -> IndexError flag */
static int ix_at(const int8_t *ix, int i) {
    if (i < 0) i += 64;
    if (i < 0 || i >= 64) { oq_index_error = 1; return 0; }
    return ix[i];
}

static int py_floordiv9(int a) { return (a >= 0) ? a / 9 : -((-a + 8) / 9); }
static int py_mod9(int a) { int m = a % 9; return m < 0 ? m + 9 : m; }

enum { C_NW = 0, C_NE = 1, C_SE = 2, C_SW = 3 };

/* quoridor.py:356-418 -- literal, branch for branch (including the row-0 aliasing at :388,:392) */
static void oq_get_intersections(const int8_t *ix, int t, int out[4]) {
    int row = py_floordiv9(t);
    int n_border = t > 71;
    int e_border = py_mod9(t) == 8;
    int s_border = t < 9;
    int w_border = py_mod9(t) == 0;
    int nw = 0, ne = 0, se = 0, sw = 0;
    if (n_border) {
        ne = 1;
        if (w_border) {
            nw = -1; sw = -1;
            se = ix_at(ix, (t - 9) - (row - 1));
        } else if (e_border) {
            nw = 1; se = -1;
            sw = ix_at(ix, (t - 9) - (row - 1) - 1);
        } else {
            nw = 1;
            sw = ix_at(ix, (t - 9) - (row - 1) - 1);
            se = ix_at(ix, (t - 9) - (row - 1));
        }
    } else if (s_border) {
        sw = 1;
        if (w_border) {
            nw = -1; se = 1;
            ne = ix_at(ix, t - row);
        } else if (e_border) {
            se = -1; ne = -1;
            nw = ne = ix_at(ix, t - row - 1);
        } else {
            se = 1;
            ne = ix_at(ix, t - row);
            nw = ne = ix_at(ix, t - row - 1);
        }
    } else if (w_border) {
        nw = -1; sw = -1;
        ne = ix_at(ix, t - row);
        se = ix_at(ix, (t - 9) - (row - 1));
    } else if (e_border) {
        ne = -1; se = -1;
        nw = ix_at(ix, t - row - 1);
        sw = ix_at(ix, (t - 9) - (row - 1) - 1);
    } else {
        ne = ix_at(ix, t - row);
        nw = ix_at(ix, t - row - 1);
        sw = ix_at(ix, (t - 9) - (row - 1) - 1);
        se = ix_at(ix, (t - 9) - (row - 1));
    }
    out[C_NW] = nw; out[C_NE] = ne; out[C_SE] = se; out[C_SW] = sw;
}

/* quoridor.py:272-353.  Returns the count; ids appended in the reference's order. */
int oq_valid_pawn_actions(const int8_t *walls, int location, int opponent_loc, int player, int *valid) {
    int nv = 0;
    int opponent_north = location == opponent_loc - 9;
    int opponent_south = location == opponent_loc + 9;
    int opponent_east = location == opponent_loc - 1;
    int opponent_west = location == opponent_loc + 1;
    int current_row = py_floordiv9(location);
    int I[4], O[4];
    oq_get_intersections(walls, location, I);
    int n = I[C_NW] != OQ_H && I[C_NE] != OQ_H && !opponent_north;
    int s = I[C_SW] != OQ_H && I[C_SE] != OQ_H && !opponent_south;
    int e = I[C_NE] != OQ_V && I[C_SE] != OQ_V && !opponent_east;
    int w = I[C_NW] != OQ_V && I[C_SW] != OQ_V && !opponent_west;
    if (n || (player == 1 && current_row == 8)) valid[nv++] = 0;
    if (s || (player == 2 && current_row == 0)) valid[nv++] = 1;
    if (e) valid[nv++] = 2;
    if (w) valid[nv++] = 3;
    if (opponent_north && I[C_NE] != OQ_H && I[C_NW] != OQ_H) {
        oq_get_intersections(walls, opponent_loc, O);
        if ((O[C_NW] != OQ_H && O[C_NE] != OQ_H) || (current_row == 7 && player == 1)) valid[nv++] = 4;
        if (O[C_NE] != OQ_V && I[C_NE] != OQ_V) valid[nv++] = 8;
        if (O[C_NW] != OQ_V && I[C_NW] != OQ_V) valid[nv++] = 9;
    } else if (opponent_south && I[C_SE] != OQ_H && I[C_SW] != OQ_H) {
        oq_get_intersections(walls, opponent_loc, O);
        if ((O[C_SW] != OQ_H && O[C_SE] != OQ_H) || (current_row == 1 && player == 2)) valid[nv++] = 5;
        if (O[C_SE] != OQ_V && I[C_SE] != OQ_V) valid[nv++] = 10;
        if (O[C_SW] != OQ_V && I[C_SW] != OQ_V) valid[nv++] = 11;
    } else if (opponent_east && I[C_SE] != OQ_V && I[C_NE] != OQ_V) {
        oq_get_intersections(walls, opponent_loc, O);
        if (O[C_SE] != OQ_V && O[C_NE] != OQ_V) valid[nv++] = 6;
        if (O[C_NE] != OQ_H) valid[nv++] = 8;
        if (O[C_SE] != OQ_H) valid[nv++] = 10;
    } else if (opponent_west && I[C_SW] != OQ_V && I[C_NW] != OQ_V) {
        oq_get_intersections(walls, opponent_loc, O);
        if (O[C_NW] != OQ_V && O[C_SW] != OQ_V) valid[nv++] = 7;
        if (O[C_NW] != OQ_H) valid[nv++] = 9;
        if (O[C_SW] != OQ_H) valid[nv++] = 11;
    }
    return nv;
}

/* offsets of quoridor.py:217-243 and :493-516 */
static const int OQ_DELTA[12] = {9, -9, 1, -1, 18, -18, 2, -2, 10, 8, -8, -10};

/* quoridor.py:479-528.  `visited` is a membership map over the positions the reference's list can
 * hold (-32..127 covers every offset reachable from an on-board tile). */
static int oq_bfs_to_goal(const int8_t *intersections, int target_row, int player_position,
                          int opponent_position, int player) {
    unsigned char visited[192];
    int queue[256];
    int qh = 0, qt = 0;
    int target_visited = 0;
    memset(visited, 0, sizeof(visited));
    queue[qt++] = player_position;
    while (!target_visited && qh < qt) {
        int cur = queue[qh++];
        int dirs[12];
        int nd = oq_valid_pawn_actions(intersections, cur, opponent_position, player, dirs);
        for (int k = 0; k < nd; k++) {
            int np_ = cur + OQ_DELTA[dirs[k]];
            int new_row = py_floordiv9(np_);
            if (new_row == target_row) {
                target_visited = 1;
            } else if (!visited[np_ + 32]) {
                visited[np_ + 32] = 1;
                if (new_row != 9 && new_row != -1) queue[qt++] = np_;
            }
        }
    }
    return target_visited;
}

static _Thread_local long long oq_path_checks = 0;   /* statistics for tests / tuning */
long long oq_get_path_checks(void) { return oq_path_checks; }

/* quoridor.py:463-477 -- both searches always run */
static int oq_blocks_path(const oq_game *g, int wall_location, int orientation) {
    oq_path_checks++;
    int8_t ix[64];
    memcpy(ix, g->intersections, 64);
    ix[wall_location] = (int8_t)orientation;
    int p1 = oq_bfs_to_goal(ix, 8, g->positions[1], g->positions[2], 1);
    int p2 = oq_bfs_to_goal(ix, 0, g->positions[2], g->positions[1], 2);
    return !(p1 && p2);
}

/* quoridor.py:432-446 */
static int oq_validate_horizontal(const oq_game *g, int ix) {
    int column = ix % 8;
    if (g->intersections[ix] != 0) return 0;
    if (column != 0 && g->intersections[ix - 1] == 1) return 0;
    if (column != 7 && g->intersections[ix + 1] == 1) return 0;
    return !oq_blocks_path(g, ix, OQ_H);
}

/* quoridor.py:448-461 */
static int oq_validate_vertical(const oq_game *g, int ix) {
    int row = ix / 8;
    if (g->intersections[ix] != 0) return 0;
    if (row != 0 && g->intersections[ix - 8] == -1) return 0;
    if (row != 7 && g->intersections[ix + 8] == -1) return 0;
    return !oq_blocks_path(g, ix, OQ_V);
}

/* the cheap half of :432-461 only (no path search) -- used by the sampled-legality rollout */
static int oq_precheck(const oq_game *g, int ix, int orientation) {
    if (g->intersections[ix] != 0) return 0;
    if (orientation == OQ_H) {
        int column = ix % 8;
        if (column != 0 && g->intersections[ix - 1] == 1) return 0;
        if (column != 7 && g->intersections[ix + 1] == 1) return 0;
    } else {
        int row = ix / 8;
        if (row != 0 && g->intersections[ix - 8] == -1) return 0;
        if (row != 7 && g->intersections[ix + 8] == -1) return 0;
    }
    return 1;
}

/* quoridor.py:138-157 (+ :420-430).  Ordered: pawn ids ascending-as-appended, then H(ix),V(ix)... */
int oq_actions(const oq_game *g, int *out) {
    int player = g->current_player;
    int opponent = player == 2 ? 1 : 2;
    int n = oq_valid_pawn_actions(g->intersections, g->positions[player], g->positions[opponent], player, out);
    if (g->walls_remaining[player] > 0) {
        for (int ix = 0; ix < 64; ix++) {
            if (oq_validate_horizontal(g, ix)) out[n++] = ix + 12;
            if (oq_validate_vertical(g, ix)) out[n++] = ix + 64 + 12;
        }
    }
    return n;
}

/* quoridor.py:193-202 */
int oq_has_a_winner(const oq_game *g, int *winner) {
    if (g->positions[2] < 9) { *winner = 2; return 1; }
    if (g->positions[1] > 71) { *winner = 1; return 1; }
    *winner = 0;
    return 0;
}

/* quoridor.py:159-186 (the recomputation of valid_actions at :165 has no effect on results and is
 * omitted; `safe` checking is done by the caller).  Returns done. */
int oq_step(oq_game *g, int action) {
    int player = g->current_player;
    if (action < 12) {
        g->positions[player] += OQ_DELTA[action];              /* :217-243 */
    } else {
        int a = action - 12;                                     /* :246-257 */
        if (a < 64) g->intersections[a] = 1; else g->intersections[a - 64] = -1;
        g->walls_remaining[player] -= 1;
    }
    int winner;
    if (oq_has_a_winner(g, &winner)) return 1;
    g->last_player = g->current_player;                          /* :260-269 */
    g->current_player = g->current_player == 1 ? 2 : 1;
    return 0;
}

/* quoridor.py:58-131 -> out[26*81] */
int oq_state(const oq_game *g, double *out) {
    memset(out, 0, sizeof(double) * 26 * 81);
    int cur = g->current_player, opp = cur == 1 ? 2 : 1;
    int pc = g->positions[cur], po = g->positions[opp];
    /* Python negative indexing on the 81-vector; beyond that the reference raises IndexError */
    if (pc < 0) pc += 81;
    if (po < 0) po += 81;
    if (pc < 0 || pc > 80 || po < 0 || po > 80) { oq_index_error = 1; return -1; }
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) {
            int v = g->intersections[r * 8 + c];
            out[(v == 0 ? 0 : (v == -1 ? 1 : 2)) * 81 + r * 9 + c] = 1.0;  /* planes 0,1,2; 8x8 padded to 9x9 */
        }
    out[3 * 81 + pc] = 1.0;
    out[4 * 81 + po] = 1.0;
    int wc = g->walls_remaining[cur] - 1, wo = g->walls_remaining[opp] - 1;   /* :79-80, index -1 wraps to 9 */
    if (wc < 0) wc += 10;
    if (wo < 0) wo += 10;
    for (int i = 0; i < 81; i++) {
        out[(5 + wc) * 81 + i] = 1.0;
        out[(15 + wo) * 81 + i] = 1.0;
        out[25 * 81 + i] = cur == 1 ? 0.0 : 1.0;
    }
    return 0;
}

/* ---------------------------------------------------------------- u64-mask <-> game helpers */
void oq_set_position(oq_game *g, uint64_t H, uint64_t V, int p1, int p2, int w1, int w2, int cur) {
    oq_reset(g);
    for (int i = 0; i < 64; i++) {
        if ((H >> i) & 1) g->intersections[i] = 1;
        else if ((V >> i) & 1) g->intersections[i] = -1;
    }
    g->positions[1] = p1; g->positions[2] = p2;
    g->walls_remaining[1] = w1; g->walls_remaining[2] = w2;
    g->current_player = cur;
    g->last_player = cur == 1 ? 2 : 1;
}

void oq_get_position(const oq_game *g, uint64_t *H, uint64_t *V, int *meta) {
    uint64_t h = 0, v = 0;
    for (int i = 0; i < 64; i++) {
        if (g->intersections[i] == 1) h |= 1ull << i;
        else if (g->intersections[i] == -1) v |= 1ull << i;
    }
    *H = h; *V = v;
    meta[0] = g->positions[1]; meta[1] = g->positions[2];
    meta[2] = g->walls_remaining[1]; meta[3] = g->walls_remaining[2];
    meta[4] = g->current_player;
}

/* 140-bit legal mask in the product's format (3 x u64, bit a = action a) for easy comparison */
void oq_legal_mask(const oq_game *g, uint64_t *mask3) {
    int acts[140];
    int n = oq_actions(g, acts);
    mask3[0] = mask3[1] = mask3[2] = 0;
    for (int i = 0; i < n; i++) mask3[acts[i] >> 6] |= 1ull << (acts[i] & 63);
}

/* ---------------------------------------------------------------- deterministic stubs
 * tests/golden/stubs.py, bit for bit. */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

uint64_t oq_state_key(const oq_game *g) {
    uint64_t H, V; int m[5];
    oq_get_position(g, &H, &V, m);
    uint64_t meta = (uint64_t)(m[0] & 0xFF) | ((uint64_t)(m[1] & 0xFF) << 8) | ((uint64_t)m[2] << 16) |
                    ((uint64_t)m[3] << 24) | ((uint64_t)m[4] << 32);
    return splitmix64(H ^ splitmix64(V ^ splitmix64(meta)));
}

static float s2_prior(uint64_t key, int a) {
    uint64_t h = splitmix64(key + (uint64_t)a * 0x9E3779B97F4A7C15ull);
    return (float)((h >> 40) + 1) * 0x1p-30f;
}

static double s2_value(uint64_t key) {
    uint64_t h = splitmix64(key ^ 0xABCDEFull);
    return (double)(h >> 40) / 8388608.0 - 1.0;
}

/* kind 1 = S1 uniform, 2 = S2 hash, 3 = S3 (S2 priors, value/8).  Fills priors[n_acts] (float32) for the ordered legal list. */
static double stub_eval(int kind, const oq_game *g, const int *acts, int n, float *priors) {
    if (kind == 1) {
        float p = 1.0f / (float)(n > 0 ? n : 1);
        for (int i = 0; i < n; i++) priors[i] = p;
        return 0.0;
    }
    uint64_t key = oq_state_key(g);
    for (int i = 0; i < n; i++) priors[i] = s2_prior(key, acts[i]);
    return kind == 3 ? s2_value(key) / 8.0 : s2_value(key);
}

/* ---------------------------------------------------------------- MCTS (mcts.py / pure_mcts.py) */
typedef struct oq_node {
    struct oq_node *parent;
    struct oq_node **children;     /* insertion-ordered "dict"  mcts.py:21 */
    int *child_actions;
    int n_children;
    int n_visits;                  /* mcts.py:22 */
    double Q, u;                   /* mcts.py:23-24 */
    double P;                      /* mcts.py:25; holds a float32 value when is_f32_prior */
} oq_node;

typedef struct {
    oq_node *root;
    double c_puct;
    int n_playout;
    int stub_kind;        /* 1,2 = AlphaZero MCTS with stub; 0 = pure MCTS (uniform float64 priors + rollout) */
    int fix_terminal_sign;/* 0 = reference behaviour (inverted), 1 = corrected */
    uint64_t rng_seed;    /* pure MCTS rollouts */
    uint64_t rollout_counter;
    long long env_steps;  /* statistics */
} oq_mcts;

static oq_node *node_new(oq_node *parent, double prior) {
    oq_node *n = (oq_node *)calloc(1, sizeof(oq_node));
    n->parent = parent;
    n->P = prior;
    return n;
}

static void node_free(oq_node *n) {
    if (!n) return;
    for (int i = 0; i < n->n_children; i++) node_free(n->children[i]);
    free(n->children);
    free(n->child_actions);
    free(n);
}

/* mcts.py:27-35 */
static void node_expand(oq_node *n, const int *acts, const double *priors, int cnt) {
    n->children = (oq_node **)malloc(sizeof(oq_node *) * (size_t)(cnt > 0 ? cnt : 1));
    n->child_actions = (int *)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
    for (int i = 0; i < cnt; i++) {
        n->child_actions[i] = acts[i];
        n->children[i] = node_new(n, priors[i]);
    }
    n->n_children = cnt;
}

/* mcts.py:64-70.  With a float32 prior numpy evaluates c_puct*P in float32 (weak Python scalar)
 * and promotes to float64 at np.sqrt(int); with a float64 prior
 * (pure_mcts.py:15) everything is float64. */
static double node_get_value(oq_node *n, double c_puct, int f32_prior) {
    double cp;
    if (f32_prior) cp = (double)((float)c_puct * (float)n->P);
    else cp = c_puct * n->P;
    n->u = cp * sqrt((double)n->parent->n_visits) / (double)(1 + n->n_visits);
    return n->Q + n->u;
}

/* mcts.py:37-42: max() keeps the first maximal item */
static int node_select(oq_node *n, double c_puct, int f32_prior) {
    int best = 0;
    double bv = node_get_value(n->children[0], c_puct, f32_prior);
    for (int i = 1; i < n->n_children; i++) {
        double v = node_get_value(n->children[i], c_puct, f32_prior);
        if (v > bv) { bv = v; best = i; }
    }
    return best;
}

/* mcts.py:44-62 */
static void node_update_recursive(oq_node *n, double leaf_value) {
    if (n->parent) node_update_recursive(n->parent, -leaf_value);
    n->n_visits += 1;
    n->Q += 1.0 * (leaf_value - n->Q) / (double)n->n_visits;
}

oq_mcts *oq_mcts_new(int stub_kind, double c_puct, int n_playout) {
    oq_mcts *t = (oq_mcts *)calloc(1, sizeof(oq_mcts));
    t->root = node_new(NULL, 1.0);
    t->c_puct = c_puct;
    t->n_playout = n_playout;
    t->stub_kind = stub_kind;
    return t;
}

void oq_mcts_free(oq_mcts *t) { if (t) { node_free(t->root); free(t); } }
void oq_mcts_set_fix_terminal_sign(oq_mcts *t, int f) { t->fix_terminal_sign = f; }
void oq_mcts_set_seed(oq_mcts *t, uint64_t seed) { t->rng_seed = seed; t->rollout_counter = 0; }
void oq_mcts_set_rollout_counter(oq_mcts *t, uint64_t c) { t->rollout_counter = c; }
long long oq_mcts_env_steps(const oq_mcts *t) { return t->env_steps; }

/* mcts.py:103-127 */
static void mcts_playout(oq_mcts *t, oq_game *game) {
    oq_node *node = t->root;
    while (node->n_children != 0) {                       /* is_leaf, :76 */
        int k = node_select(node, t->c_puct, 1);
        int a = node->child_actions[k];
        node = node->children[k];
        oq_step(game, a);
        t->env_steps++;
    }
    int winner;
    int end = oq_has_a_winner(game, &winner);
    double leaf_value;
    if (!end) {
        int acts[140]; float pri[140]; double prid[140];
        int n = oq_actions(game, acts);                   /* policy_value_fn, policy_value_net.py:150 */
        leaf_value = stub_eval(t->stub_kind, game, acts, n, pri);
        for (int i = 0; i < n; i++) prid[i] = (double)pri[i];
        node_expand(node, acts, prid, n);
    } else {
        leaf_value = (winner == game->current_player) ? 1.0 : -1.0;   /* :125 -- always +1 */
        if (t->fix_terminal_sign) leaf_value = -leaf_value;
    }
    node_update_recursive(node, -leaf_value);
}

/* ---- counter-based RNG shared with the device rollout kernel (Philox4x32-10) ---- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oq_philox(uint64_t seed, uint64_t rid, uint32_t c2, uint32_t c3, uint32_t *out4) {
    uint32_t ctr[4] = {(uint32_t)rid, (uint32_t)(rid >> 32), c2, c3};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, out4);
}

/*
 * One uniformly random legal action (pure_mcts.py:7-10 + :99: argmax of iid U(0,1) over the legal list == a
 * uniform pick), by the draw procedure the product specifies (csrc/qz_sample.cuh) restated on top of the
 * LITERAL rules: the legal list is actions() (every wall verified by the reference's BFS); the ordered superset
 * S is the legal pawn moves followed by the walls passing the prechecks of quoridor.py:432-461 (H by ix, then V);
 *   attempt 0 of ply t uses word (t & 3)       of Philox(ctr = (rid_lo, rid_hi, t >> 2, 0));
 *   attempt j >= 1     uses word ((j - 1) & 3) of Philox(ctr = (rid_lo, rid_hi, t, 1 + ((j - 1) >> 2)));
 *   the attempt draws S[(word * M) >> 32] (with replacement) and stops if that action is in actions().
 * Returns the action, or -1 when nothing is legal (stalemate).
 */
int oq_sample_action(const oq_game *g, uint64_t seed, uint64_t rid, uint32_t ply) {
    int player = g->current_player, opp = player == 1 ? 2 : 1;
    int legal[140], is_legal[140];
    int nl = oq_actions(g, legal);
    memset(is_legal, 0, sizeof(is_legal));
    for (int i = 0; i < nl; i++) is_legal[legal[i]] = 1;
    int S[140], M = 0;
    int pawn[12];
    int np_ = oq_valid_pawn_actions(g->intersections, g->positions[player], g->positions[opp], player, pawn);
    for (int i = 0; i < np_; i++) S[M++] = pawn[i];
    if (g->walls_remaining[player] > 0) {
        for (int ix = 0; ix < 64; ix++) if (oq_precheck(g, ix, OQ_H)) S[M++] = 12 + ix;
        for (int ix = 0; ix < 64; ix++) if (oq_precheck(g, ix, OQ_V)) S[M++] = 76 + ix;
    }
    if (M == 0 || nl == 0) return -1;
    for (uint32_t j = 0;; j++) {
        uint32_t w[4], word;
        if (j == 0) { oq_philox(seed, rid, ply >> 2, 0, w); word = w[ply & 3]; }
        else { oq_philox(seed, rid, ply, 1u + ((j - 1u) >> 2), w); word = w[(j - 1u) & 3u]; }
        int a = S[(int)(((uint64_t)word * (uint32_t)M) >> 32)];
        if (is_legal[a]) return a;
    }
}

/* pure_mcts.py:86-108: <= limit-1 random plies; result from the starting mover's point of view.
 * Returns +1/-1, or 0 when no winner (limit reached or stalemate).  *plies = steps taken. */
int oq_rollout(oq_game *game, uint64_t seed, uint64_t rid, int limit, int *plies) {
    int player = game->current_player;
    int winner = 0, steps = 0;
    for (int i = 0; i < limit; i++) {
        if (oq_has_a_winner(game, &winner)) break;
        if (i == limit - 1) break;
        int a = oq_sample_action(game, seed, rid, (uint32_t)i);
        if (a < 0) break;                                /* stalemate: reference would raise */
        oq_step(game, a);
        steps++;
    }
    if (plies) *plies = steps;
    if (winner == 0) return 0;
    return winner == player ? 1 : -1;
}

/* pure_mcts.py:86-108 restated LITERALLY: every ply builds the full legal list with actions() (the whole
 * 128-candidate sweep while the mover has walls) and picks uniformly from it (rollout_policy_fn, :7-10).
 * This is the form the CPU baseline times; results are distributed exactly like oq_rollout's. */
int oq_rollout_literal(oq_game *game, uint64_t seed, uint64_t rid, int limit, int *plies) {
    int player = game->current_player;
    int winner = 0, steps = 0;
    for (int i = 0; i < limit; i++) {
        if (oq_has_a_winner(game, &winner)) break;
        if (i == limit - 1) break;
        int acts[140];
        int n = oq_actions(game, acts);
        if (n == 0) break;
        uint32_t w[4];
        oq_philox(seed, rid, (uint32_t)i >> 2, 0, w);
        int a = acts[(int)(((uint64_t)w[i & 3] * (uint32_t)n) >> 32)];
        oq_step(game, a);
        steps++;
    }
    if (plies) *plies = steps;
    if (winner == 0) return 0;
    return winner == player ? 1 : -1;
}

static int oq_literal_rollouts = 0;
void oq_set_literal_rollouts(int on) { oq_literal_rollouts = on; }

/* pure_mcts.py:66-83 */
static void pure_playout(oq_mcts *t, oq_game *game) {
    oq_node *node = t->root;
    while (node->n_children != 0) {
        int k = node_select(node, t->c_puct, 0);
        int a = node->child_actions[k];
        node = node->children[k];
        oq_step(game, a);
        t->env_steps++;
    }
    int winner;
    int end = oq_has_a_winner(game, &winner);
    if (!end) {
        int acts[140]; double pri[140];
        int n = oq_actions(game, acts);                   /* pure_mcts.py:13-16 */
        for (int i = 0; i < n; i++) pri[i] = 1.0 / (double)n;
        node_expand(node, acts, pri, n);
    }
    int plies = 0;
    double leaf_value = oq_literal_rollouts
        ? (double)oq_rollout_literal(game, t->rng_seed, t->rollout_counter++, 1000, &plies)
        : (double)oq_rollout(game, t->rng_seed, t->rollout_counter++, 1000, &plies);
    t->env_steps += plies;
    if (end && t->fix_terminal_sign) leaf_value = -leaf_value;
    node_update_recursive(node, -leaf_value);
}

/* mcts.py:129-144 / pure_mcts.py:110-115: run n_playout playouts from `game` (copied per playout,
 * the deepcopy of mcts.py:136).  Outputs children in insertion order.  Returns n_children. */
int oq_mcts_run(oq_mcts *t, const oq_game *game, int *acts, int *visits, double *qs) {
    for (int n = 0; n < t->n_playout; n++) {
        oq_game copy = *game;
        if (t->stub_kind == 0) pure_playout(t, &copy); else mcts_playout(t, &copy);
    }
    oq_node *r = t->root;
    for (int i = 0; i < r->n_children; i++) {
        acts[i] = r->child_actions[i];
        visits[i] = r->children[i]->n_visits;
        if (qs) qs[i] = r->children[i]->Q;
    }
    return r->n_children;
}

void oq_mcts_root_stats(const oq_mcts *t, int *n_visits, double *q) {
    *n_visits = t->root->n_visits; *q = t->root->Q;
}

/* mcts.py:6-9,141-144: softmax(1/temp * log(visits + 1e-10)) */
void oq_visits_to_probs(const int *visits, int n, double temp, double *probs) {
    double mx = -INFINITY, sum = 0.0;
    for (int i = 0; i < n; i++) {
        probs[i] = 1.0 / temp * log((double)visits[i] + 1e-10);
        if (probs[i] > mx) mx = probs[i];
    }
    for (int i = 0; i < n; i++) { probs[i] = exp(probs[i] - mx); sum += probs[i]; }
    for (int i = 0; i < n; i++) probs[i] /= sum;
}

/* mcts.py:146-151 */
void oq_mcts_update_with_move(oq_mcts *t, int last_move) {
    oq_node *r = t->root;
    for (int i = 0; i < r->n_children; i++) {
        if (r->child_actions[i] == last_move) {
            oq_node *keep = r->children[i];
            r->children[i] = NULL;
            keep->parent = NULL;
            node_free(r);
            t->root = keep;
            return;
        }
    }
    node_free(r);
    t->root = node_new(NULL, 1.0);
}

/* Uniform-random legal play with the FULL legal list every ply (BASELINE config 1) and the product's pick rule
 * (qz_env_sample_legal): the legal actions sorted by action id, k = (word * n) >> 32,
 * word = Philox(ctr = (game id, ply >> 2, 0x7000))[ply & 3].  Plays g to the end (or cap plies); returns plies played. */
int oq_random_game(oq_game *g, uint64_t seed, uint64_t game_id, int cap) {
    int plies = 0;
    while (plies < cap) {
        int winner;
        if (oq_has_a_winner(g, &winner)) break;
        int acts[140], present[140];
        int n = oq_actions(g, acts);
        if (n == 0) break;
        memset(present, 0, sizeof(present));
        for (int i = 0; i < n; i++) present[acts[i]] = 1;
        uint32_t w[4];
        oq_philox(seed, game_id, (uint32_t)plies >> 2, 0x7000u, w);
        int k = (int)(((uint64_t)w[plies & 3] * (uint32_t)n) >> 32);
        int a = -1;
        for (int id = 0; id < 140; id++) if (present[id] && k-- == 0) { a = id; break; }
        oq_step(g, a);
        plies++;
    }
    return plies;
}

/* ---------------------------------------------------------------- CPU-baseline drivers (bench.py)
 * A minimal pthread parallel-for with dynamic (atomic counter) scheduling: one work item per game /
 * position, all host threads the caller asks for. */
typedef void (*oq_item_fn)(int i, void *ctx, long long *acc);

typedef struct {
    oq_item_fn fn; void *ctx; int n; int next; long long acc[4]; pthread_mutex_t mu;
} oq_pf;

static void *oq_pf_worker(void *p) {
    oq_pf *pf = (oq_pf *)p;
    long long local[4] = {0, 0, 0, 0};
    for (;;) {
        int i = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
        if (i >= pf->n) break;
        pf->fn(i, pf->ctx, local);
    }
    pthread_mutex_lock(&pf->mu);
    for (int k = 0; k < 4; k++) pf->acc[k] += local[k];
    pthread_mutex_unlock(&pf->mu);
    return NULL;
}

int oq_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void oq_parallel_for(int n, int n_threads, oq_item_fn fn, void *ctx, long long acc[4]) {
    oq_pf pf;
    memset(&pf, 0, sizeof(pf));
    pf.fn = fn; pf.ctx = ctx; pf.n = n;
    pthread_mutex_init(&pf.mu, NULL);
    if (n_threads <= 0) n_threads = oq_max_threads();
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&th[started], NULL, oq_pf_worker, &pf) == 0) started++;
    oq_pf_worker(&pf);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&pf.mu);
    for (int k = 0; k < 4; k++) acc[k] = pf.acc[k];
}

/* BASELINE config 1: uniform-random legal play from reset() to terminal (cap plies), each ply =
 * actions() (full 128-candidate sweep while the mover has walls) + step().  A splitmix64 stream
 * picks the index; the choice of RNG does not affect the cost. */
typedef struct { uint64_t seed; int cap; long long *plies; int *winner; } rg_ctx;

static void rg_item(int gi, void *vctx, long long *acc) {
    rg_ctx *c = (rg_ctx *)vctx;
    oq_game g;
    oq_reset(&g);
    uint64_t s = splitmix64(c->seed ^ (uint64_t)gi * 0x9E3779B97F4A7C15ull);
    int plies = 0, winner = 0;
    while (plies < c->cap) {
        int acts[140];
        int n = oq_actions(&g, acts);
        if (n == 0) break;
        s = splitmix64(s);
        int a = acts[(int)(((s >> 32) * (uint64_t)n) >> 32)];
        plies++;
        if (oq_step(&g, a)) break;
    }
    oq_has_a_winner(&g, &winner);
    if (c->plies) c->plies[gi] = plies;
    if (c->winner) c->winner[gi] = winner;
    acc[0] += plies;
}

long long oq_bench_random_games(int n_games, uint64_t seed, int cap, int n_threads, long long *out_plies,
                                int *out_winner) {
    rg_ctx c = {seed, cap, out_plies, out_winner};
    long long acc[4];
    oq_parallel_for(n_games, n_threads, rg_item, &c, acc);
    return acc[0];
}

typedef struct {
    const uint64_t *H, *V; const int *meta5; uint64_t *mask3; int n_playout; double c_puct; uint64_t seed;
    int stub_kind; int *moves; int *visits;
} pos_ctx;

static void load_pos(const pos_ctx *c, int i, oq_game *g) {
    const int *m = c->meta5 + 5 * i;
    oq_set_position(g, c->H[i], c->V[i], m[0], m[1], m[2], m[3], m[4]);
}

/* BASELINE config 5: full legal sweep on given positions.  Returns total legal actions (checksum). */
static void sw_item(int i, void *vctx, long long *acc) {
    pos_ctx *c = (pos_ctx *)vctx;
    oq_game g;
    load_pos(c, i, &g);
    uint64_t mk[3];
    oq_legal_mask(&g, mk);
    if (c->mask3) { c->mask3[3 * i] = mk[0]; c->mask3[3 * i + 1] = mk[1]; c->mask3[3 * i + 2] = mk[2]; }
    acc[0] += __builtin_popcountll(mk[0]) + __builtin_popcountll(mk[1]) + __builtin_popcountll(mk[2]);
}

long long oq_bench_sweeps(const uint64_t *H, const uint64_t *V, const int *meta5, int n, int n_threads,
                          uint64_t *mask3_out) {
    pos_ctx c = {H, V, meta5, mask3_out, 0, 0.0, 0, 0, NULL, NULL};
    long long acc[4];
    oq_parallel_for(n, n_threads, sw_item, &c, acc);
    return acc[0];
}

/* BASELINE config 2: pure-MCTS move decisions (pure_mcts.py:138-143) for n_games independent positions,
 * n_playout playouts each with a fresh tree.  Returns total env steps (tree descent + rollout plies);
 * moves_out[i] = first-max-visits move. */
static void pm_item(int i, void *vctx, long long *acc) {
    pos_ctx *c = (pos_ctx *)vctx;
    oq_game g;
    load_pos(c, i, &g);
    oq_mcts *t = oq_mcts_new(0, c->c_puct, c->n_playout);
    oq_mcts_set_seed(t, splitmix64(c->seed + (uint64_t)i));
    int acts[140], visits[140];
    int n = oq_mcts_run(t, &g, acts, visits, NULL);
    int best = 0;
    for (int k = 1; k < n; k++) if (visits[k] > visits[best]) best = k;   /* pure_mcts.py:115 */
    if (c->moves) c->moves[i] = n > 0 ? acts[best] : -1;
    acc[0] += t->env_steps;
    acc[1] += c->n_playout;
    oq_mcts_free(t);
}

long long oq_bench_pure_mcts(const uint64_t *H, const uint64_t *V, const int *meta5, int n_games,
                             int n_playout, double c_puct, uint64_t seed, int n_threads, int *moves_out,
                             long long *playouts_out) {
    pos_ctx c = {H, V, meta5, NULL, n_playout, c_puct, seed, 0, moves_out, NULL};
    long long acc[4];
    oq_parallel_for(n_games, n_threads, pm_item, &c, acc);
    if (playouts_out) *playouts_out = acc[1];
    return acc[0];
}

/* BASELINE config 3: AlphaZero-MCTS simulations with the deterministic stub as the evaluator. */
static void sm_item(int i, void *vctx, long long *acc) {
    pos_ctx *c = (pos_ctx *)vctx;
    oq_game g;
    load_pos(c, i, &g);
    oq_mcts *t = oq_mcts_new(c->stub_kind, c->c_puct, c->n_playout);
    int acts[140], visits[140];
    int n = oq_mcts_run(t, &g, acts, visits, NULL);
    if (c->visits) {
        for (int k = 0; k < 140; k++) c->visits[140 * i + k] = 0;
        for (int k = 0; k < n; k++) c->visits[140 * i + acts[k]] = visits[k];
    }
    acc[0] += c->n_playout;
    oq_mcts_free(t);
}

long long oq_bench_stub_mcts(const uint64_t *H, const uint64_t *V, const int *meta5, int n_games,
                             int n_playout, double c_puct, int stub_kind, int n_threads, int *visits_out) {
    pos_ctx c = {H, V, meta5, NULL, n_playout, c_puct, 0, stub_kind, NULL, visits_out};
    long long acc[4];
    oq_parallel_for(n_games, n_threads, sm_item, &c, acc);
    return acc[0];
}
