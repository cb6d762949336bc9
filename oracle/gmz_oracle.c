/*
 * oracle/gmz_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's Gumbel-MCTS self-play path
 * (Datou/Datou-gomoku-muzero).  It is the checker for the CUDA engine and the
 * "port" CPU baseline of bench.py; nothing under datou_gomoku_muzero_b200/
 * may import, link or call it.  Each function cites the reference file:line
 * it restates.  Parity of this restatement is PINNED against fixtures produced
 * by importing the unmodified Python reference (tests/golden/make_golden.py ->
 * tests/golden/ npz files; checked by tests/test_oracle_golden.py).
 *
 * Arithmetic notes: the harness evaluator hands the
 * reference Python floats, so value_sum / Q / minmax / scores are IEEE double
 * with one rounding per operation.  Build with -ffp-contract=off so no FMA is
 * formed; the CUDA side is built with -fmad=false for the same reason.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_WORDS 8 /* 64-bit words per bitboard: boards up to 22x22 */

typedef struct {
    int32_t board_size;      /* config.BOARD_SIZE        config.py:18 */
    int32_t n_in_row;        /* config.N_IN_ROW          config.py:19 */
    int32_t num_simulations; /* config.NUM_SIMULATIONS   config.py:22 */
    int32_t num_top_actions; /* config.NUM_TOP_ACTIONS   config.py:23 */
    int32_t mode;            /* 0 = AlphaZeroMCTS, 1 = MuZeroMCTS (config.py:25) */
    int32_t eval_kind;       /* 0 = E0 hash evaluator, 1 = constant (tests/test_mcts_logic.py:60-80), 3 = harness callback (AlphaZero mode) */
    int32_t logit_div;       /* E0: > 0 quantised logits (k-32)/logit_div, k in 0..63; 0 = dense 24-bit logits / values */
    int32_t accum_dtype;     /* 0: the evaluator hands Python floats -> float64 arithmetic (upstream test mock);
                              * 1: it hands np.float32 scalars (workers.py:355,368) -> float32 value_sum / Q /
                              *    MinMaxStats under NumPy >= 2 */
    double c_visit;          /* config.py:31 */
    double c_scale;          /* config.py:32 */
    double minmax_delta;     /* config.py:33 */
    double discount;         /* config.py:34 */
    double const_value;      /* eval_kind 1 */
    double const_reward;     /* eval_kind 1, MuZero mode */
    uint64_t eval_seed;      /* E0 */
} orc_config;

/* ------------------------------------------------------------------ */
/* E0: the fixed deterministic evaluator (DESIGN.md "E0").  Pure integer
 * hashing of the observation, so Python / C / CUDA agree bit for bit.  */
/* ------------------------------------------------------------------ */
#define E0_GOLD 0x9E3779B97F4A7C15ULL
#define E0_CV 0xD1B54A32D192ED03ULL
#define E0_CA 0x8CB92BA72F3D8DD7ULL

#define E0_K1 0xBF58476D1CE4E5B9ULL
#define E0_K2 0x94D049BB133111EBULL

static inline uint64_t mix64(uint64_t z)
{
    z ^= z >> 30; z *= E0_K1;
    z ^= z >> 27; z *= E0_K2;
    z ^= z >> 31;
    return z;
}
static inline uint64_t rot32(uint64_t z) { return (z << 32) | (z >> 32); }

/* hash of the observation planes game.py:12-17 builds: own / opp / last move.  The per-word terms
 * are combined with XOR, so the words can be hashed in any order (in parallel on the GPU). */
static uint64_t e0_hash_obs(const orc_config *c, const int8_t *board, int player, int last_move)
{
    const int A = c->board_size * c->board_size, nw = (A + 63) / 64;
    uint64_t own[ORC_MAX_WORDS] = {0}, opp[ORC_MAX_WORDS] = {0};
    for (int a = 0; a < A; ++a) {
        if (board[a] == player) own[a >> 6] |= 1ULL << (a & 63);
        else if (board[a] == -player) opp[a >> 6] |= 1ULL << (a & 63);
    }
    const uint64_t h0 = mix64(c->eval_seed ^ E0_GOLD);
    uint64_t acc = 0;
    for (int w = 0; w < nw; ++w) {
        acc ^= ((own[w] ^ h0) + (uint64_t)(2 * w + 1) * E0_GOLD) * E0_K1;            /* one odd multiply per plane word; */
        acc ^= rot32(((opp[w] ^ h0) + (uint64_t)(2 * w + 2) * E0_GOLD) * E0_K2);     /* the final mix64 avalanches       */
    }
    return mix64(acc + (uint64_t)(int64_t)(last_move + 1) * E0_CV);
}

/* per-action 32-bit hash (one multiply; only the high bits are used) */
static inline uint32_t e0_action_hash(uint32_t s, int a)
{
    uint32_t x = s + (uint32_t)(a + 1) * 0x9E3779B1u;
    x ^= x >> 16; x *= 0x7FEB352Du;
    return x;
}

static void e0_heads(const orc_config *c, uint64_t h, float *logits, double *value)
{
    const int A = c->board_size * c->board_size;
    const uint32_t s = (uint32_t)h ^ (uint32_t)(h >> 32);
    const int vk = (int)((h >> 40) & 0xFFFFFF);
    if (c->logit_div > 0) {
        for (int a = 0; a < A; ++a) logits[a] = (float)((int)(e0_action_hash(s, a) >> 26) - 32) / (float)c->logit_div;
        *value = (double)(vk % 33 - 16) / 16.0;
    } else {   /* dense: 24 random mantissa bits, no quantisation */
        for (int a = 0; a < A; ++a) logits[a] = (float)((int)(e0_action_hash(s, a) >> 8) - (1 << 23)) * 0x1p-21f;
        *value = (double)(vk - (1 << 23)) * 0x1p-23;
    }
}

static uint64_t e0_child_hidden(uint64_t h_parent, int action)
{
    return mix64(h_parent + (uint64_t)(action + 1) * E0_CA);
}

static double e0_reward(const orc_config *c, uint64_t h)
{
    const int rk = (int)((h >> 16) & 0xFFFFFF);
    return c->logit_div > 0 ? (double)(rk % 5 - 2) / 16.0 : (double)(rk - (1 << 23)) * 0x1p-25;
}

/* eval_kind 3: an evaluator supplied by the test harness (e.g. a table of what a real network returned
 * for each observation): called with the position the reference would build its observation from. */
typedef void (*orc_eval_fn)(const int8_t *board, int32_t board_size, int32_t player, int32_t last_move,
                            float *logits, double *value);
static orc_eval_fn g_eval_cb = 0;
void orc_set_eval_callback(orc_eval_fn fn) { g_eval_cb = fn; }

/* exported so tests can check the Python and CUDA evaluators against it */
void orc_e0_initial(const orc_config *c, const int8_t *board, int player, int last_move,
                    float *logits, double *value, uint64_t *hidden)
{
    if (c->eval_kind == 3 && g_eval_cb) {
        g_eval_cb(board, c->board_size, player, last_move, logits, value);
        *hidden = 3;
        return;
    }
    if (c->eval_kind == 1) {
        const int A = c->board_size * c->board_size;
        for (int a = 0; a < A; ++a) logits[a] = 0.0f;
        *value = c->const_value; *hidden = 1;
        return;
    }
    uint64_t h = e0_hash_obs(c, board, player, last_move);
    e0_heads(c, h, logits, value);
    *hidden = h;
}

void orc_e0_recurrent(const orc_config *c, uint64_t h_parent, int action,
                      float *logits, double *value, double *reward, uint64_t *hidden)
{
    if (c->eval_kind == 1) {
        const int A = c->board_size * c->board_size;
        for (int a = 0; a < A; ++a) logits[a] = 0.0f;
        *value = c->const_value; *reward = c->const_reward; *hidden = 2;
        return;
    }
    uint64_t h = e0_child_hidden(h_parent, action);
    e0_heads(c, h, logits, value);
    *reward = e0_reward(c, h);
    *hidden = h;
}

/* ------------------------------------------------------------------ */
/* utils.MinMaxStats  (utils.py:6-25)                                  */
/* ------------------------------------------------------------------ */
typedef struct { double maximum, minimum, delta; int f32; } MinMax;

static void mm_init(MinMax *m, double delta, int f32) { m->maximum = -INFINITY; m->minimum = INFINITY; m->delta = delta; m->f32 = f32; }
static void mm_update(MinMax *m, double v)
{ /* utils.py:12-14 */
    if (v > m->maximum) m->maximum = v;
    if (v < m->minimum) m->minimum = v;
}
static double mm_normalize(const MinMax *m, double v)
{ /* utils.py:16-25 */
    if (m->maximum > m->minimum) {
        /* float32 mode: maximum / minimum are np.float32, so the denominator (max - min + delta) is
         * float32 arithmetic with delta rounded to float32; `v` arrives as np.float64 -- the element of
         * the float64 array np.array([get_qsa(a) ...]) (mcts.py:142: Python 0.0 for the unvisited
         * actions mixed with np.float32 promotes the array) -- so the numerator and the division are
         * float64.  (An array with NO Python float, i.e. every one of the A actions visited, would
         * stay float32: see `all_visited` in transformed_qs.) */
        double den = m->f32 ? (double)(((float)m->maximum - (float)m->minimum) + (float)m->delta)
                            : m->maximum - m->minimum + m->delta;
        double n = (v - m->minimum) / den;
        double lo = n < 1.0 ? n : 1.0;     /* min(1.0, normalized) */
        return lo > 0.0 ? lo : 0.0;        /* max(0.0, ...)        */
    }
    return 0.0;
}

/* ------------------------------------------------------------------ */
/* mcts.Node (mcts.py:14-44)                                           */
/* ------------------------------------------------------------------ */
typedef struct {
    int action, parent, visit_count, expanded;
    double value_sum, reward;
    float *logits;    /* [A] once expanded */
    int *children;    /* [A] -> node index or -1, allocated on first get_child */
    uint64_t hidden;  /* evaluator-defined hidden state handle */
} Node;

typedef struct {
    const orc_config *cfg;
    int A;
    Node *nodes; int n_nodes, cap;
    MinMax mm;
    const uint8_t *valid;     /* [A] root-valid mask, fixed for the whole search (mcts.py:213) */
    int f32;                  /* cfg->accum_dtype == 1: value_sum / value / Q carried as np.float32 (stored widened) */
    int all_visited;          /* float32 mode: a node with all A children visited was scored (dtype case not restated) */
    /* per-search engine state that the reference keeps on self (mcts.py:159,221,226) */
    int *sel; int n_sel;      /* selected_children_actions */
    const double *gumbel;
    int current_phase, current_num_top_actions; double used_visit_num; int visit_num_for_next_phase;
} Tree;

static int node_new(Tree *t, int action, int parent)
{
    if (t->n_nodes == t->cap) { t->cap *= 2; t->nodes = (Node *)realloc(t->nodes, sizeof(Node) * (size_t)t->cap); }
    Node *n = &t->nodes[t->n_nodes];
    memset(n, 0, sizeof(*n));
    n->action = action; n->parent = parent;
    return t->n_nodes++;
}
static int node_get_child(Tree *t, int node, int action)
{ /* mcts.py:27-30 */
    if (!t->nodes[node].children) {
        int *ch = (int *)malloc(sizeof(int) * (size_t)t->A);
        for (int a = 0; a < t->A; ++a) ch[a] = -1;
        t->nodes[node].children = ch;
    }
    if (t->nodes[node].children[action] < 0) {
        int id = node_new(t, action, node);          /* may realloc nodes */
        t->nodes[node].children[action] = id;
    }
    return t->nodes[node].children[action];
}
static int node_child_or_neg(const Tree *t, int node, int action)
{
    const Node *n = &t->nodes[node];
    return n->children ? n->children[action] : -1;
}
static double node_get_value(const Tree *t, const Node *n)
{ /* mcts.py:32-33; float32 mode: np.float32 / int -> float32 division */
    if (n->visit_count <= 0) return 0.0;
    return t->f32 ? (double)((float)n->value_sum / (float)n->visit_count) : n->value_sum / (double)n->visit_count;
}
static double node_get_qsa(const Tree *t, int node, int action)
{ /* mcts.py:35-38; float32 mode: reward + float32(discount) * value, both operations in float32 */
    int c = node_child_or_neg(t, node, action);
    if (c >= 0 && t->nodes[c].visit_count > 0) {
        if (t->f32) return (double)((float)t->nodes[c].reward + (float)t->cfg->discount * (float)node_get_value(t, &t->nodes[c]));
        return t->nodes[c].reward + t->cfg->discount * node_get_value(t, &t->nodes[c]);
    }
    return 0.0;
}
static void node_expand(Tree *t, int node, const float *logits, uint64_t hidden, double reward)
{ /* mcts.py:24-25 */
    Node *n = &t->nodes[node];
    if (!n->logits) n->logits = (float *)malloc(sizeof(float) * (size_t)t->A);
    memcpy(n->logits, logits, sizeof(float) * (size_t)t->A);
    n->hidden = hidden; n->reward = reward; n->expanded = 1;
}

static double clip1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); } /* np.clip(v,-1,1) */

/* GumbelMCTSBase._backpropagate for one leaf (mcts.py:119-138) */
static void backpropagate(Tree *t, int leaf, double value)
{
    value = clip1(value);
    int node = leaf;
    while (node >= 0) {
        Node *n = &t->nodes[node];
        if (t->f32) n->value_sum = (double)((float)n->value_sum + (float)value);   /* int 0 + np.float32 -> np.float32 */
        else n->value_sum += value;
        n->visit_count += 1;
        if (n->parent >= 0) mm_update(&t->mm, node_get_qsa(t, n->parent, n->action));
        if (t->f32) value = (double)((float)n->reward + (float)t->cfg->discount * (float)value);
        else value = n->reward + t->cfg->discount * value;
        value = clip1(value);
        node = n->parent;
    }
}

/* _get_transformed_completed_Qs (mcts.py:141-149) */
static void transformed_qs(const Tree *t, int node, double *out)
{
    const orc_config *c = t->cfg;
    int max_child_visit = 0, n_visited = 0;
    for (int a = 0; a < t->A; ++a) {
        int ch = node_child_or_neg(t, node, a);
        if (ch >= 0 && t->nodes[ch].visit_count > max_child_visit) max_child_visit = t->nodes[ch].visit_count;
        if (ch >= 0 && t->nodes[ch].visit_count > 0) ++n_visited;
    }
    /* float32 mode with every action visited: the reference's q array would be float32 and the rest of
     * the transform / softmax follow in float32 (not restated; reported through out_counts[4]) */
    if (t->f32 && n_visited == t->A) ((Tree *)t)->all_visited = 1;
    const double scale = (c->c_visit + (double)max_child_visit) * c->c_scale;
    for (int a = 0; a < t->A; ++a)
        out[a] = scale * mm_normalize(&t->mm, node_get_qsa(t, node, a));
}

/* _get_improved_policy (mcts.py:151-156): torch.softmax(float64) over the
 * root-valid actions of logits + sigma; ATen's CPU kernel computes
 * exp(x - max) * (1 / sum). */
static void improved_policy(const Tree *t, int node, const double *tq, double *out)
{
    const Node *n = &t->nodes[node];
    double mx = -INFINITY;
    for (int a = 0; a < t->A; ++a) {
        if (t->valid[a]) { out[a] = (double)n->logits[a] + tq[a]; if (out[a] > mx) mx = out[a]; }
        else out[a] = -INFINITY;
    }
    double sum = 0.0;
    for (int a = 0; a < t->A; ++a) { out[a] = t->valid[a] ? exp(out[a] - mx) : 0.0; sum += out[a]; }
    const double inv = 1.0 / sum;
    for (int a = 0; a < t->A; ++a) out[a] = out[a] * inv;
}

/* _select_action (mcts.py:95-117) */
static int select_action(Tree *t, int node, double *tq, double *pol)
{
    if (t->nodes[node].parent < 0) { /* root: first least-visited survivor, strict < */
        int best = -1; double min_visits = INFINITY;
        for (int i = 0; i < t->n_sel; ++i) {
            int ch = node_child_or_neg(t, node, t->sel[i]);
            int v = ch >= 0 ? t->nodes[ch].visit_count : 0;
            if ((double)v < min_visits) { min_visits = (double)v; best = t->sel[i]; }
        }
        return best;
    }
    transformed_qs(t, node, tq);
    improved_policy(t, node, tq, pol);
    long total = 0;
    for (int a = 0; a < t->A; ++a) { int ch = node_child_or_neg(t, node, a); if (ch >= 0) total += t->nodes[ch].visit_count; }
    int best = -1; double best_score = -INFINITY;
    for (int a = 0; a < t->A; ++a) {
        if (!t->valid[a]) continue;
        int ch = node_child_or_neg(t, node, a);
        double nv = ch >= 0 ? (double)t->nodes[ch].visit_count : 0.0;
        double score = pol[a] - nv / (double)(1 + total);
        if (best < 0 || score > best_score) { best_score = score; best = a; }   /* np.argmax: first max */
    }
    return best;
}

/* _select_leaf (mcts.py:88-93) */
static int select_leaf(Tree *t, double *tq, double *pol, int *depth_out)
{
    int node = 0, depth = 0;
    while (t->nodes[node].expanded) {
        int a = select_action(t, node, tq, pol);
        node = node_get_child(t, node, a);
        ++depth;
    }
    *depth_out = depth;
    return node;
}

/* _initialize_sequential_halving_schedule (mcts.py:158-164) */
static void halving_init(Tree *t)
{
    const int n = t->cfg->num_simulations, m = t->cfg->num_top_actions;
    t->current_phase = 0; t->current_num_top_actions = m; t->used_visit_num = 0.0;
    if (m <= 1 || log2((double)m) <= 0) t->visit_num_for_next_phase = n;
    else {
        double v = floor((double)n / (log2((double)m) * (double)m)) * (double)m;
        if ((double)n < v) v = (double)n;
        t->visit_num_for_next_phase = (int)v;
    }
}
/* _ready_for_next_gumbel_phase (mcts.py:166-181) */
static int halving_ready(Tree *t, int sim_idx)
{
    if (sim_idx < t->visit_num_for_next_phase) return 0;
    t->current_phase += 1;
    t->current_num_top_actions /= 2;
    if (t->current_num_top_actions < 1) return 0;
    const int n = t->cfg->num_simulations, m = t->cfg->num_top_actions, cm = t->current_num_top_actions;
    double extra;
    if (cm <= 1 || log2((double)m) <= 0) extra = (double)n - t->used_visit_num;
    else extra = floor((double)n / (log2((double)m) * (double)cm)) * (double)cm;
    t->used_visit_num += extra;
    int nx = t->visit_num_for_next_phase + (int)extra;
    t->visit_num_for_next_phase = nx < n ? nx : n;
    return 1;
}
/* _sequential_halving (mcts.py:183-185): stable descending sort, keep first m */
static void sequential_halving(Tree *t, double *tq)
{
    transformed_qs(t, 0, tq);
    double sc[1024]; int idx[1024];
    const int k = t->n_sel;
    for (int i = 0; i < k; ++i) {
        int a = t->sel[i];
        sc[i] = (t->gumbel[a] + (double)t->nodes[0].logits[a]) + tq[a];
        idx[i] = a;
    }
    for (int i = 1; i < k; ++i) { /* insertion sort, descending, stable */
        double s = sc[i]; int a = idx[i]; int j = i - 1;
        while (j >= 0 && sc[j] < s) { sc[j + 1] = sc[j]; idx[j + 1] = idx[j]; --j; }
        sc[j + 1] = s; idx[j + 1] = a;
    }
    int keep = t->current_num_top_actions < k ? t->current_num_top_actions : k;
    for (int i = 0; i < keep; ++i) t->sel[i] = idx[i];
    t->n_sel = keep;
}

/* Iteration order of a CPython set built by inserting `keys` (ascending
 * non-negative ints, hash(k) == k) one by one -- the order
 * `max(visit_counts, key=visit_counts.get)` scans in at mcts.py:274-275.
 * Restates Objects/setobject.c set_add_entry / set_table_resize /
 * set_insert_clean (LINEAR_PROBES 9, PERTURB_SHIFT 5), CPython >= 3.7. */
void orc_pyset_order(const int32_t *keys, int n, int32_t *out)
{
    size_t mask = 7, fill = 0;
    int32_t *table = (int32_t *)malloc(sizeof(int32_t) * 8);
    for (size_t i = 0; i < 8; ++i) table[i] = -1;
    for (int k = 0; k < n; ++k) {
        const size_t hash = (size_t)keys[k];
        size_t i = hash & mask, perturb = hash, e;
        for (;;) {
            e = i;
            int probes = (i + 9 <= mask) ? 9 : 0, found = 0;
            do { if (table[e] < 0) { found = 1; break; } ++e; } while (probes--);
            if (found) break;
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
        table[e] = keys[k]; ++fill;
        if (fill * 5 >= mask * 3) { /* set_table_resize(used*4) */
            size_t minused = fill > 50000 ? fill * 2 : fill * 4, newsize = 8;
            while (newsize <= minused) newsize <<= 1;
            int32_t *nt = (int32_t *)malloc(sizeof(int32_t) * newsize);
            for (size_t j = 0; j < newsize; ++j) nt[j] = -1;
            const size_t nmask = newsize - 1;
            for (size_t j = 0; j <= mask; ++j) {
                if (table[j] < 0) continue;
                const size_t h2 = (size_t)table[j];
                size_t ii = h2 & nmask, pp = h2, ee;
                for (;;) {
                    ee = ii;
                    if (nt[ee] < 0) break;
                    int ok = 0;
                    if (ii + 9 <= nmask) { for (int q = 0; q < 9; ++q) { ++ee; if (nt[ee] < 0) { ok = 1; break; } } }
                    if (ok) break;
                    pp >>= 5;
                    ii = (ii * 5 + 1 + pp) & nmask;
                }
                nt[ee] = table[j];
            }
            free(table); table = nt; mask = nmask;
        }
    }
    int o = 0;
    for (size_t j = 0; j <= mask; ++j) if (table[j] >= 0) out[o++] = table[j];
    free(table);
}

static void tree_free(Tree *t)
{
    for (int i = 0; i < t->n_nodes; ++i) { free(t->nodes[i].logits); free(t->nodes[i].children); }
    free(t->nodes); free(t->sel);
}

/* game.py:20-23 do_move on a flat board (overwrites, no legality check) */
static void do_move(int8_t *board, int *player, int *last_move, int *move_count, int a)
{
    board[a] = (int8_t)*player; *last_move = a; *player = -*player; *move_count += 1;
}

/*
 * AlphaZeroMCTS.search (mcts.py:197-280) / MuZeroMCTS.search (mcts.py:288-362).
 * Returns 0, or 1 for the sentinel "(zeros, 0.0, -1)" of mcts.py:214-215.
 * Optional outputs (may be NULL): out_visits[A] root-child visit counts,
 * out_leaf_actions/out_leaf_depths[num_simulations] one entry per evaluation
 * after the root, out_counts[5] = {sim_count, n_evals, n_nodes, max_depth, all_visited (float32 mode)},
 * out_minmax[2] = {minimum, maximum}.
 */
int orc_search(const orc_config *cfg, const int8_t *board, int player, int last_move, int move_count,
               const double *gumbel, double *out_policy, double *out_value, int32_t *out_action,
               int32_t *out_visits, int32_t *out_leaf_actions, int32_t *out_leaf_depths,
               int32_t *out_counts, double *out_minmax)
{
    const int A = cfg->board_size * cfg->board_size, S = cfg->num_simulations;
    (void)move_count;
    uint8_t *valid = (uint8_t *)malloc((size_t)A);
    int n_valid = 0;
    for (int a = 0; a < A; ++a) { valid[a] = board[a] == 0; n_valid += valid[a]; }
    for (int a = 0; a < A; ++a) out_policy[a] = 0.0;
    *out_value = 0.0; *out_action = -1;
    if (out_visits) memset(out_visits, 0, sizeof(int32_t) * (size_t)A);
    if (out_counts) memset(out_counts, 0, sizeof(int32_t) * 5);
    if (n_valid == 0) { free(valid); return 1; }

    Tree t; memset(&t, 0, sizeof(t));
    t.cfg = cfg; t.A = A; t.cap = S + 4; t.nodes = (Node *)malloc(sizeof(Node) * (size_t)t.cap);
    t.valid = valid; t.gumbel = gumbel; t.f32 = cfg->accum_dtype == 1; mm_init(&t.mm, cfg->minmax_delta, t.f32);
    float *logits = (float *)malloc(sizeof(float) * (size_t)A);
    double *tq = (double *)malloc(sizeof(double) * (size_t)A), *pol = (double *)malloc(sizeof(double) * (size_t)A);
    int8_t *tmp = (int8_t *)malloc((size_t)A);
    int *hist = (int *)malloc(sizeof(int) * (size_t)(S + 4));
    double value; uint64_t hidden;

    /* root (mcts.py:203-218) */
    const int root = node_new(&t, -1, -1);
    orc_e0_initial(cfg, board, player, last_move, logits, &value, &hidden);
    node_expand(&t, root, logits, hidden, 0.0);
    backpropagate(&t, root, value);

    /* Gumbel top-k (mcts.py:220-226): sort (score, action) tuples descending */
    halving_init(&t);
    {
        double *sc = (double *)malloc(sizeof(double) * (size_t)n_valid); int *ac = (int *)malloc(sizeof(int) * (size_t)n_valid);
        int k = 0;
        for (int a = 0; a < A; ++a) if (valid[a]) { sc[k] = gumbel[a] + (double)t.nodes[root].logits[a]; ac[k] = a; ++k; }
        for (int i = 1; i < k; ++i) { /* descending by (score, action) */
            double s = sc[i]; int a = ac[i]; int j = i - 1;
            while (j >= 0 && (sc[j] < s || (sc[j] == s && ac[j] < a))) { sc[j + 1] = sc[j]; ac[j + 1] = ac[j]; --j; }
            sc[j + 1] = s; ac[j + 1] = a;
        }
        t.n_sel = k < t.current_num_top_actions ? k : t.current_num_top_actions;
        t.sel = (int *)malloc(sizeof(int) * (size_t)(t.n_sel > 0 ? t.n_sel : 1));
        for (int i = 0; i < t.n_sel; ++i) t.sel[i] = ac[i];
        free(sc); free(ac);
    }

    int sim_count = 1, n_evals = 0, max_depth = 0;
    while (sim_count < S) {
        int depth = 0;
        if (cfg->mode == 0) {
            /* AlphaZero rollout on the real board (mcts.py:229-264) */
            int leaf = select_leaf(&t, tq, pol, &depth);
            int n_hist = 0;
            for (int nd = leaf; t.nodes[nd].parent >= 0; nd = t.nodes[nd].parent) hist[n_hist++] = t.nodes[nd].action;
            memcpy(tmp, board, (size_t)A);
            int p = player, lm = -1, mc = move_count;   /* temp_game.last_move starts as None (game.py:10) */
            for (int i = n_hist - 1; i >= 0; --i) do_move(tmp, &p, &lm, &mc, hist[i]);
            orc_e0_initial(cfg, tmp, p, lm, logits, &value, &hidden);
            node_expand(&t, leaf, logits, hidden, 0.0);
            backpropagate(&t, leaf, value);
            if (out_leaf_actions) out_leaf_actions[n_evals] = t.nodes[leaf].action;
            if (out_leaf_depths) out_leaf_depths[n_evals] = depth;
            ++n_evals; sim_count += 1;
        } else {
            /* MuZero batch (mcts.py:320-346): len(selected) selections with no
             * stat change in between all reach the same leaf; one recurrent
             * evaluation, then len(selected) sequential backups. */
            const int k = t.n_sel;
            if (k == 0) break;
            int leaf = -1;
            for (int i = 0; i < k; ++i) leaf = select_leaf(&t, tq, pol, &depth);
            double reward;
            const int parent = t.nodes[leaf].parent;
            orc_e0_recurrent(cfg, t.nodes[parent].hidden, t.nodes[leaf].action, logits, &value, &reward, &hidden);
            for (int i = 0; i < k; ++i) node_expand(&t, leaf, logits, hidden, reward);
            for (int i = 0; i < k; ++i) backpropagate(&t, leaf, value);
            if (out_leaf_actions) out_leaf_actions[n_evals] = t.nodes[leaf].action;
            if (out_leaf_depths) out_leaf_depths[n_evals] = depth;
            ++n_evals; sim_count += k;
        }
        if (depth > max_depth) max_depth = depth;
        if (halving_ready(&t, sim_count)) sequential_halving(&t, tq);
    }

    /* decision (mcts.py:271-280) */
    transformed_qs(&t, root, tq);
    improved_policy(&t, root, tq, out_policy);
    *out_value = node_get_value(&t, &t.nodes[root]);
    {
        int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_valid), *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_valid);
        int k = 0;
        for (int a = 0; a < A; ++a) if (valid[a]) keys[k++] = a;
        orc_pyset_order(keys, n_valid, order);
        int best = -1, best_n = -1;
        for (int i = 0; i < n_valid; ++i) {
            int ch = node_child_or_neg(&t, root, order[i]);
            int v = ch >= 0 ? t.nodes[ch].visit_count : 0;
            if (v > best_n) { best_n = v; best = order[i]; }
        }
        *out_action = best;
        free(keys); free(order);
    }
    if (out_visits)
        for (int a = 0; a < A; ++a) { int ch = node_child_or_neg(&t, root, a); out_visits[a] = ch >= 0 ? t.nodes[ch].visit_count : 0; }
    if (out_counts) { out_counts[0] = sim_count; out_counts[1] = n_evals; out_counts[2] = t.n_nodes; out_counts[3] = max_depth; out_counts[4] = t.all_visited; }
    if (out_minmax) { out_minmax[0] = t.mm.minimum; out_minmax[1] = t.mm.maximum; }

    tree_free(&t); free(valid); free(logits); free(tq); free(pol); free(tmp); free(hist);
    return 0;
}

/* Batch of independent searches, one per game, spread over host threads
 * (the reference's NUM_WORKERS process parallelism, main.py:100-101). */
int orc_search_batch(const orc_config *cfg, int n_games, const int8_t *boards, const int8_t *players,
                     const int32_t *last_moves, const int32_t *move_counts, const double *gumbel,
                     double *out_policy, double *out_value, int32_t *out_action, int32_t *out_visits,
                     int n_threads)
{
    const int A = cfg->board_size * cfg->board_size;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int g = 0; g < n_games; ++g) {
        orc_search(cfg, boards + (size_t)g * A, players[g], last_moves[g], move_counts[g], gumbel + (size_t)g * A,
                   out_policy + (size_t)g * A, out_value + g, out_action + g,
                   out_visits ? out_visits + (size_t)g * A : NULL, NULL, NULL, NULL, NULL);
    }
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* game.GomokuGame                                                     */
/* ------------------------------------------------------------------ */
/* check_win (game.py:25-58): from (r,c), four directions, count up to
 * n_in_row+1 stones each way; win if count >= n_in_row (overlines win). */
int orc_check_win(const int8_t *board, int N, int n_in_row, int r, int c)
{
    const int player = board[r * N + c];
    if (player == 0) return 0;
    static const int D[4][2] = {{0, 1}, {1, 0}, {1, 1}, {1, -1}};
    for (int d = 0; d < 4; ++d) {
        int count = 1;
        for (int i = 1; i < n_in_row + 2; ++i) {
            int nr = r + i * D[d][0], nc = c + i * D[d][1];
            if (nr >= 0 && nr < N && nc >= 0 && nc < N && board[nr * N + nc] == player) ++count; else break;
        }
        for (int i = 1; i < n_in_row + 2; ++i) {
            int nr = r - i * D[d][0], nc = c - i * D[d][1];
            if (nr >= 0 && nr < N && nc >= 0 && nc < N && board[nr * N + nc] == player) ++count; else break;
        }
        if (count >= n_in_row) return 1;
    }
    return 0;
}
/* get_game_ended (game.py:60-63): stone colour on a win, 0 on a draw, 2 = None */
int orc_game_ended(const int8_t *board, int N, int n_in_row, int last_move, int move_count)
{
    if (last_move >= 0 && orc_check_win(board, N, n_in_row, last_move / N, last_move % N)) return board[last_move];
    if (move_count >= N * N) return 0;
    return 2;
}

/* Self-play game loop (workers.py:162-189): search -> record -> do_move ->
 * get_game_ended; gumbel[max_moves][A] supplies each search's noise in order.
 * Outputs per move: actions, root values, policies [T][A], boards [T][A];
 * returns T (number of moves), *winner = +-1 / 0. */
int orc_selfplay_game(const orc_config *cfg, const double *gumbel, int max_moves,
                      int32_t *out_actions, double *out_values, double *out_policies, int8_t *out_boards,
                      int32_t *winner)
{
    const int N = cfg->board_size, A = N * N;
    int8_t *board = (int8_t *)calloc((size_t)A, 1);
    int player = 1, last_move = -1, move_count = 0, T = 0;
    *winner = 2;
    while (T < max_moves) {
        double value; int32_t action;
        int rc = orc_search(cfg, board, player, last_move, move_count, gumbel + (size_t)T * A,
                            out_policies + (size_t)T * A, &value, &action, NULL, NULL, NULL, NULL, NULL);
        if (rc != 0 || action < 0) break;
        memcpy(out_boards + (size_t)T * A, board, (size_t)A);
        out_actions[T] = action; out_values[T] = value; ++T;
        do_move(board, &player, &last_move, &move_count, action);
        int w = orc_game_ended(board, N, cfg->n_in_row, last_move, move_count);
        if (w != 2) { *winner = w; break; }
    }
    free(board);
    return T;
}

/* final_rewards (workers.py:183-187) */
void orc_final_rewards(int T, int winner, float *out)
{
    for (int i = 0; i < T; ++i) out[i] = 0.0f;
    if (winner != 0 && T > 0) {
        out[T - 1] = 1.0f;
        if (T > 1) out[T - 2] = -1.0f;
        for (int i = T - 3; i >= 0; --i) out[i] = -out[i + 2];
    }
}

/* compute_n_step_returns as called from self-play (workers.py:144-152,205):
 * rewards is a list of Python floats (double arithmetic for the reward sum),
 * values -> float32 array, bootstrap = float32(v) * float32(discount**n)
 * (NumPy>=2 weak-scalar promotion), float + float32 -> float32 add. */
void orc_n_step_returns(const double *rewards, const double *values, int T, int n_values,
                        double discount, int n_steps, float *out)
{
    for (int t = T - 1; t >= 0; --t) {
        const int bi = t + n_steps;
        double acc = 0.0; int first = 1; /* sum() starts from int 0: 0 + x == x exactly */
        for (int i = 0; i < n_steps; ++i) {
            if (t + i < T) {
                double term = pow(discount, (double)i) * rewards[t + i];
                acc = first ? (0.0 + term) : acc + term; first = 0;
            }
        }
        if (bi < n_values) {
            float b = (float)values[bi] * (float)pow(discount, (double)n_steps);
            out[t] = (float)acc + b;
        } else {
            out[t] = (float)(acc + 0.0);
        }
    }
}

/* ------------------------------------------------------------------ */
/* replay_buffer.SumTree / InMemoryReplayBuffer (replay_buffer.py:4-106) */
/* ------------------------------------------------------------------ */
/* SumTree.update + _propagate (replay_buffer.py:11-19) */
void orc_sumtree_update(double *tree, int64_t tree_idx, double priority)
{
    double change = priority - tree[tree_idx];
    tree[tree_idx] = priority;
    int64_t idx = tree_idx;
    while (idx != 0) { idx = (idx - 1) / 2; tree[idx] += change; }
}
/* SumTree.get_leaf (replay_buffer.py:27-38) */
int64_t orc_sumtree_get_leaf(const double *tree, int64_t capacity, double value)
{
    const int64_t len = 2 * capacity - 1;
    int64_t parent = 0;
    for (;;) {
        int64_t left = 2 * parent + 1, right = left + 1;
        if (left >= len) return parent;
        if (value <= tree[left]) parent = left;
        else { value -= tree[left]; parent = right; }
    }
}
/* InMemoryReplayBuffer.sample, PER branch (replay_buffer.py:60-86); u01[B]
 * are the uniform doubles np.random.uniform would have drawn. */
void orc_per_sample(const double *tree, int64_t capacity, int64_t count, int B, const double *u01, double beta,
                    int64_t *out_tree_idx, double *out_priority, float *out_weights)
{
    const double total = tree[0], segment = total / (double)B;
    float max_w = -INFINITY;
    for (int i = 0; i < B; ++i) {
        double lo = segment * (double)i, hi = segment * (double)(i + 1);
        double s = lo + (hi - lo) * u01[i];
        int64_t idx = orc_sumtree_get_leaf(tree, capacity, s);
        double p = tree[idx];
        out_tree_idx[i] = idx; out_priority[i] = p;
        double prob = p / total;
        out_weights[i] = (float)pow((double)count * prob, -beta);
        if (out_weights[i] > max_w) max_w = out_weights[i];
    }
    if (max_w > 0) for (int i = 0; i < B; ++i) out_weights[i] /= max_w;
}
/* InMemoryReplayBuffer.update_priorities (replay_buffer.py:98-103) */
double orc_per_update(double *tree, int B, const int64_t *tree_idx, const double *td_abs_plus_eps, double max_priority)
{
    for (int i = 0; i < B; ++i) {
        if (td_abs_plus_eps[i] > max_priority) max_priority = td_abs_plus_eps[i];
        orc_sumtree_update(tree, tree_idx[i], td_abs_plus_eps[i]);
    }
    return max_priority;
}
