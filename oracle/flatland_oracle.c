/* flatland_oracle.c — TEST INFRASTRUCTURE ONLY (see flatland_oracle.h).
 *
 * Serial restatement of the reference algorithm, one function per reference function, kept
 * deliberately close to the reference's own structure (dense predicted_pos[t][agent] tables,
 * linear scans, an explicit motion graph) so that it is an independent check of the
 * re-designed CUDA path.  All citations are relative to /root/reference.
 *
 * Floating point: every float expression below mirrors one C++ float expression of
 * flatland_cutils (compiled for x86-64 SSE, FLT_EVAL_METHOD 0); build with -ffp-contract=off.
 */
#define _GNU_SOURCE
#include "flatland_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

enum { WAITING = 0, READY = 1, MAL_OFF = 2, MOVING = 3, STOPPED = 4, MALFUNCTION = 5, DONE = 6 };
enum { A_NOTHING = 0, A_LEFT = 1, A_FORWARD = 2, A_RIGHT = 3, A_STOP = 4 };
static const int DR[4] = {-1, 0, 1, 0};
static const int DC[4] = {0, 1, 0, -1};
#define NPRED (FO_PRED_DEPTH + 1) /* rows 0..500 used by the tree (treeobs.cpp:50-65) */

struct FoEnv {
    int H, W, N, T;
    uint16_t *grid;
    /* static per agent (agent_utils.py:58-88) */
    int *init_r, *init_c, *init_dir, *tgt_r, *tgt_c, *earliest, *latest, *slot;
    double *speed;
    /* dynamic per agent */
    int *r, *c, *dir, *state, *old_r, *old_c, *old_dir, *ctr, *mal, *nmal, *saved, *arrival, *sig_mal;
    uint8_t *done;
    int elapsed, done_all;
    /* distance map per unique target (distance_map.py:71-79 computes once per unique target) */
    int n_slots;
    float *dm; /* [n_slots][H][W][4], INFINITY = unreachable */
    /* observation builder state */
    uint8_t *deadlocked; /* sticky (deadlock_checker.cpp) */
    /* scratch for observation */
    int *ppos, *pdir; /* [NPRED][N] */
};

static int on_map(int s) { return s == MOVING || s == STOPPED || s == MALFUNCTION; }
static int off_map(int s) { return s == WAITING || s == READY || s == MAL_OFF; }
static int popc16(unsigned v) { return __builtin_popcount(v & 0xFFFFu); }
static int cell(const FoEnv *e, int r, int c) { return e->grid[r * e->W + c]; }
/* grid4.py:66-87 / tool.h:337-352: nibble of orientation o, bit order N,E,S,W msb->lsb */
static int nib(const FoEnv *e, int r, int c, int o) { return (cell(e, r, c) >> ((3 - o) * 4)) & 0xF; }
static int tbit(int nibble, int d) { return (nibble >> (3 - d)) & 1; }
static int in_bounds(const FoEnv *e, int r, int c) { return r >= 0 && c >= 0 && r < e->H && c < e->W; }
static float dmv(const FoEnv *e, int slot, int r, int c, int d) {
    return e->dm[(((size_t)slot * e->H + r) * e->W + c) * 4 + d];
}

/* ------------------------------------------------------------------------------------------- */
/* DistanceMap._distance_map_walker / _get_and_update_neighbors (distance_map.py:81-160)        */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int r, c, o; float d; } QNode;

static int dm_neighbors(FoEnv *e, float *m, int r, int c, float cur, int enforce, QNode *out) {
    int n = 0;
    for (int nd = 0; nd < 4; nd++) {
        if (enforce >= 0 && nd != (enforce + 2) % 4) continue;
        int rr = r + DR[nd], cc = c + DC[nd];
        if (!in_bounds(e, rr, cc)) continue;
        int want = (nd + 2) % 4; /* movement from the neighbour into (r,c) */
        for (int o = 0; o < 4; o++) {
            if (!tbit(nib(e, rr, cc, o), want)) continue;
            float *p = &m[((size_t)rr * e->W + cc) * 4 + o];
            float nv = *p < cur + 1 ? *p : cur + 1;
            out[n].r = rr; out[n].c = cc; out[n].o = o; out[n].d = nv; n++;
            *p = nv;
        }
    }
    return n;
}

static void dm_walk(FoEnv *e, int slot, int tr, int tc) {
    size_t cells = (size_t)e->H * e->W;
    float *m = e->dm + (size_t)slot * cells * 4;
    for (size_t i = 0; i < cells * 4; i++) m[i] = INFINITY;
    for (int o = 0; o < 4; o++) m[((size_t)tr * e->W + tc) * 4 + o] = 0;
    uint8_t *visited = calloc(cells * 4, 1);
    for (int o = 0; o < 4; o++) visited[((size_t)tr * e->W + tc) * 4 + o] = 1;
    size_t cap = 1024, head = 0, tail = 0;
    QNode *q = malloc(cap * sizeof(QNode));
    QNode tmp[16];
    int k = dm_neighbors(e, m, tr, tc, 0, -1, tmp);
    for (int i = 0; i < k; i++) q[tail++] = tmp[i];
    while (head < tail) {
        QNode nd = q[head++];
        size_t id = ((size_t)nd.r * e->W + nd.c) * 4 + nd.o;
        if (visited[id]) continue;
        visited[id] = 1;
        k = dm_neighbors(e, m, nd.r, nd.c, nd.d, nd.o, tmp);
        if (tail + k > cap) { cap *= 2; q = realloc(q, cap * sizeof(QNode)); }
        for (int i = 0; i < k; i++) q[tail++] = tmp[i];
    }
    free(q);
    free(visited);
}

/* DistanceMap._compute (distance_map.py:57-79): one walk per unique target, in agent order. */
static void dm_compute(FoEnv *e) {
    int ns = 0;
    int *sr = malloc(sizeof(int) * e->N), *sc = malloc(sizeof(int) * e->N);
    for (int i = 0; i < e->N; i++) {
        int s = -1;
        for (int j = 0; j < ns; j++) if (sr[j] == e->tgt_r[i] && sc[j] == e->tgt_c[i]) { s = j; break; }
        if (s < 0) { s = ns; sr[ns] = e->tgt_r[i]; sc[ns] = e->tgt_c[i]; ns++; }
        e->slot[i] = s;
    }
    e->n_slots = ns;
    free(e->dm);
    e->dm = malloc(sizeof(float) * (size_t)ns * e->H * e->W * 4);
    for (int s = 0; s < ns; s++) dm_walk(e, s, sr[s], sc[s]);
    free(sr); free(sc);
}

/* ------------------------------------------------------------------------------------------- */
/* construction / reset                                                                         */
/* ------------------------------------------------------------------------------------------- */
#define IALLOC(n) ((int *)calloc((size_t)(n), sizeof(int)))

FoEnv *fo_create(int H, int W, int N, int T, const uint16_t *grid, const int16_t *init_pos,
                 const uint8_t *init_dir, const int16_t *target, const double *speed,
                 const int32_t *earliest, const int32_t *latest) {
    FoEnv *e = calloc(1, sizeof(FoEnv));
    e->H = H; e->W = W; e->N = N; e->T = T;
    e->grid = malloc(sizeof(uint16_t) * (size_t)H * W);
    memcpy(e->grid, grid, sizeof(uint16_t) * (size_t)H * W);
    e->init_r = IALLOC(N); e->init_c = IALLOC(N); e->init_dir = IALLOC(N); e->tgt_r = IALLOC(N);
    e->tgt_c = IALLOC(N); e->earliest = IALLOC(N); e->latest = IALLOC(N); e->slot = IALLOC(N);
    e->speed = calloc(N, sizeof(double));
    e->r = IALLOC(N); e->c = IALLOC(N); e->dir = IALLOC(N); e->state = IALLOC(N); e->old_r = IALLOC(N);
    e->old_c = IALLOC(N); e->old_dir = IALLOC(N); e->ctr = IALLOC(N); e->mal = IALLOC(N);
    e->nmal = IALLOC(N); e->saved = IALLOC(N); e->arrival = IALLOC(N); e->sig_mal = IALLOC(N);
    e->done = calloc(N, 1);
    e->deadlocked = calloc(N, 1);
    e->ppos = IALLOC((size_t)NPRED * N); e->pdir = IALLOC((size_t)NPRED * N);
    for (int i = 0; i < N; i++) {
        e->init_r[i] = init_pos[2 * i]; e->init_c[i] = init_pos[2 * i + 1]; e->init_dir[i] = init_dir[i];
        e->tgt_r[i] = target[2 * i]; e->tgt_c[i] = target[2 * i + 1];
        e->speed[i] = speed[i]; e->earliest[i] = earliest[i]; e->latest[i] = latest[i];
        e->arrival[i] = -1;   /* EnvAgent.arrival_time: attrib default None, set once per agent OBJECT (see reset_dynamic) */
    }
    return e;
}

void fo_free(FoEnv *e) {
    if (!e) return;
    free(e->grid); free(e->init_r); free(e->init_c); free(e->init_dir); free(e->tgt_r); free(e->tgt_c);
    free(e->earliest); free(e->latest); free(e->slot); free(e->speed); free(e->r); free(e->c); free(e->dir);
    free(e->state); free(e->old_r); free(e->old_c); free(e->old_dir); free(e->ctr); free(e->mal);
    free(e->nmal); free(e->saved); free(e->arrival); free(e->sig_mal); free(e->done); free(e->deadlocked);
    free(e->ppos); free(e->pdir); free(e->dm);
    free(e);
}

static void reset_dynamic(FoEnv *e) {
    /* EnvAgent.reset (agent_utils.py:90-105).  It does NOT touch arrival_time: the attribute starts as None when the
     * agent object is made (EnvAgent.from_line at a regenerating reset) and survives reset(False, False).  A train that
     * arrived in the previous episode therefore keeps its old arrival_time, and when it reaches its target again
     * handle_done_state (rail_env.py:493-499) does nothing: it stays on its target cell in state DONE. */
    for (int i = 0; i < e->N; i++) {
        e->r[i] = e->c[i] = -1; e->dir[i] = e->init_dir[i];
        e->old_r[i] = e->old_c[i] = -1; e->old_dir[i] = -1;
        e->state[i] = WAITING; e->ctr[i] = 0; e->mal[i] = 0; e->nmal[i] = 0; e->saved[i] = 0;
        e->sig_mal[i] = 0; e->done[i] = 0; e->deadlocked[i] = 0;
    }
    e->elapsed = 0; e->done_all = 0;
}

void fo_reset(FoEnv *e) {
    if (!e->dm) dm_compute(e); /* the map is static for a given world */
    reset_dynamic(e);
    /* TreeObsForRailEnv::reset -> AgentsLoader::update runs the deadlock checker once before the
     * first get_many (treeobs.cpp:22-28); with every agent off-map it changes nothing. */
}

/* ------------------------------------------------------------------------------------------- */
/* action preprocessing (rail_env.py:425-446, action_preprocessing.py, transition_utils.py)     */
/* ------------------------------------------------------------------------------------------- */
/* check_action (transition_utils.py:6-44): returns new direction, *valid = -1 (None) / 0 / 1 */
static int check_action(const FoEnv *e, int a, int r, int c, int d, int *valid) {
    int nb = nib(e, r, c, d), nt = popc16(nb), nd = d;
    *valid = -1;
    if (a == A_LEFT) { nd = d - 1; if (nt <= 1) *valid = 0; }
    else if (a == A_RIGHT) { nd = d + 1; if (nt <= 1) *valid = 0; }
    nd = ((nd % 4) + 4) % 4;
    if (a == A_FORWARD && nt == 1) {
        nd = tbit(nb, 0) ? 0 : tbit(nb, 1) ? 1 : tbit(nb, 2) ? 2 : 3; /* fast_argmax */
        *valid = 1;
    }
    return nd;
}

/* check_valid_action (transition_utils.py:47-82) */
static int check_valid_action(const FoEnv *e, int a, int r, int c, int d) {
    int valid, nd = check_action(e, a, r, c, d, &valid);
    int rr = r + DR[nd], cc = c + DC[nd];
    int cell_ok = in_bounds(e, rr, cc) && cell(e, rr, cc) > 0;
    if (valid < 0) valid = tbit(nib(e, r, c, d), nd);
    return cell_ok && valid;
}

static int preprocess_action(const FoEnv *e, int i, int raw) {
    int a = (raw >= 0 && raw <= 4) ? raw : A_NOTHING;            /* process_illegal_action */
    if (a == A_NOTHING) {                                         /* process_do_nothing */
        if (e->state[i] == MOVING) a = A_FORWARD;
        else if (e->saved[i]) a = e->saved[i];
    }
    if (e->state[i] == WAITING) a = A_NOTHING;                    /* preprocess_action_when_waiting */
    int r = e->r[i], c = e->c[i], d = e->dir[i];
    if (r < 0) { r = e->init_r[i]; c = e->init_c[i]; d = e->init_dir[i]; }
    if ((a == A_LEFT || a == A_RIGHT) && !check_valid_action(e, a, r, c, d)) a = A_FORWARD;
    if (a >= A_LEFT && a <= A_RIGHT && !check_valid_action(e, a, r, c, d)) a = A_STOP;
    return a;
}

/* ------------------------------------------------------------------------------------------- */
/* MotionCheck (agent_chains.py): explicit graph, same passes and node order as the reference   */
/* ------------------------------------------------------------------------------------------- */
enum { COL_NONE = 0, COL_RED = 1, COL_PURPLE = 2, COL_OTHER = 3 };

/* The motion graph: nodes are cells in insertion order; every agent contributes one edge
 * cur -> next.  Several agents can stand on one cell (the reference puts a train on the map without
 * asking MotionCheck when MALFUNCTION_OFF_MAP meets a STOP action, state_machine.py:41-42), so a node
 * can have several out-edges; its "agent" attribute is the last agent added (agent_chains.py:33). */
typedef struct {
    int n_nodes, n_agents;
    int *nr, *nc;      /* node key (row, col) in insertion order */
    int *agent;        /* agent attribute of a node, -1 if none */
    int *first_succ;   /* first out-edge added to the node (what G.successors(n).__next__() yields) */
    int *color;
    int *cur_node, *nxt_node; /* per agent */
} MGraph;

static int mg_node(MGraph *g, int r, int c) {
    for (int i = 0; i < g->n_nodes; i++) if (g->nr[i] == r && g->nc[i] == c) return i;
    int i = g->n_nodes++;
    g->nr[i] = r; g->nc[i] = c; g->agent[i] = -1; g->first_succ[i] = -1; g->color[i] = COL_NONE;
    return i;
}

static int mg_edge(const MGraph *g, int u, int v) {
    for (int k = 0; k < g->n_agents; k++) if (g->cur_node[k] == u && g->nxt_node[k] == v) return 1;
    return 0;
}

/* marks every node from which `v` is reachable (v included): dfs over the reversed graph */
static void mg_reverse_closure(const MGraph *g, int v, uint8_t *mark) {
    int *stack = malloc(sizeof(int) * (g->n_nodes + 1)), sp = 0;
    uint8_t *seen = calloc(g->n_nodes, 1);
    stack[sp++] = v; seen[v] = 1;
    while (sp) {
        int u = stack[--sp];
        mark[u] = 1;
        for (int k = 0; k < g->n_agents; k++) {
            int w = g->cur_node[k];
            if (g->nxt_node[k] == u && !seen[w]) { seen[w] = 1; stack[sp++] = w; }
        }
    }
    free(stack); free(seen);
}

static void mg_block_preds(MGraph *g, int v, int color) { /* block_preds, agent_chains.py:125-149 */
    uint8_t *mark = calloc(g->n_nodes, 1);
    mg_reverse_closure(g, v, mark);
    for (int u = 0; u < g->n_nodes; u++) if (mark[u] && g->color[u] != color) g->color[u] = color;
    free(mark);
}

void fo_motion_check(int n, const int16_t *cur, const int16_t *nxt, uint8_t *can_move) {
    MGraph g;
    int cap = 2 * n + 2;
    g.n_nodes = 0; g.n_agents = n;
    g.nr = IALLOC(cap); g.nc = IALLOC(cap); g.agent = IALLOC(cap); g.first_succ = IALLOC(cap); g.color = IALLOC(cap);
    g.cur_node = IALLOC(n); g.nxt_node = IALLOC(n);
    for (int i = 0; i < n; i++) { g.cur_node[i] = g.nxt_node[i] = -1; }
    for (int i = 0; i < n; i++) {                       /* addAgent, agent_chains.py:19-37 */
        int u = mg_node(&g, cur[2 * i], cur[2 * i + 1]);
        g.agent[u] = i;
        int v = mg_node(&g, nxt[2 * i], nxt[2 * i + 1]);
        if (g.first_succ[u] < 0) g.first_succ[u] = v;
        g.cur_node[i] = u; g.nxt_node[i] = v;
    }
    int nn = g.n_nodes;
    uint8_t *stops = calloc(nn, 1), *swaps = calloc(nn, 1), *blocked = calloc(nn, 1);
    for (int k = 0; k < n; k++) if (g.cur_node[k] == g.nxt_node[k]) stops[g.cur_node[k]] = 1;   /* find_stops2: self loops */
    for (int k = 0; k < n; k++) {                                               /* find_swaps: 2-cycles */
        int u = g.cur_node[k], v = g.nxt_node[k];
        if (u != v && mg_edge(&g, v, u)) { swaps[u] = 1; swaps[v] = 1; }
    }
    for (int u = 0; u < nn; u++) if (swaps[u]) mg_block_preds(&g, u, COL_PURPLE);
    for (int u = 0; u < nn; u++) if (stops[u]) mg_reverse_closure(&g, u, blocked); /* find_stop_preds */
    int *preds = IALLOC(nn + 1);
    for (int v = 0; v < nn; v++) {                      /* G.pred.items() in node insertion order */
        int np = 0;                                     /* predecessors of v in node insertion order, each once */
        for (int k = 0; k < n; k++) {
            if (g.nxt_node[k] != v) continue;
            int w = g.cur_node[k], at = np;
            while (at > 0 && preds[at - 1] >= w) at--;
            if (at < np && preds[at] == w) continue;    /* several trains on one cell heading the same way: one edge */
            for (int q = np; q > at; q--) preds[q] = preds[q - 1];
            preds[at] = w; np++;
        }
        if (blocked[v]) {
            g.color[v] = COL_RED;
        } else if (np > 1) {
            if (g.color[v] == COL_RED || g.color[v] == COL_PURPLE) continue;
            g.color[v] = COL_OTHER; /* blue / magenta */
            int win = -1;
            for (int k = 0; k < np; k++) if (win < 0 || g.agent[preds[k]] < g.agent[win]) win = preds[k];
            for (int k = 0; k < np; k++) if (preds[k] != win) mg_block_preds(&g, preds[k], COL_RED);
        }
    }
    for (int i = 0; i < n; i++) {                       /* check_motion, agent_chains.py:204-236 */
        int u = g.cur_node[i];
        if (g.color[u] == COL_RED || g.color[u] == COL_PURPLE) can_move[i] = 0;
        else can_move[i] = (g.first_succ[u] != u);
    }
    free(preds); free(stops); free(swaps); free(blocked);
    free(g.nr); free(g.nc); free(g.agent); free(g.first_succ); free(g.color); free(g.cur_node); free(g.nxt_node);
}

/* ------------------------------------------------------------------------------------------- */
/* state machine (step_utils/state_machine.py:12-80)                                            */
/* ------------------------------------------------------------------------------------------- */
static int fsm(int s, int in_mal, int mal_done, int edr, int stop, int valid_move, int reached, int conflict) {
    switch (s) {
    case WAITING: return in_mal ? MAL_OFF : edr ? READY : WAITING;
    case READY: return in_mal ? MAL_OFF : valid_move ? MOVING : READY;
    case MAL_OFF:
        if (!mal_done) return MAL_OFF;
        if (!edr) return WAITING;
        return valid_move ? MOVING : stop ? STOPPED : READY;
    case MOVING: return in_mal ? MALFUNCTION : reached ? DONE : (stop || conflict) ? STOPPED : MOVING;
    case STOPPED: return in_mal ? MALFUNCTION : valid_move ? MOVING : STOPPED;
    case MALFUNCTION: return !mal_done ? MALFUNCTION : valid_move ? MOVING : STOPPED;
    default: return DONE;
    }
}

/* greedy descent length used by the end reward: len(get_shortest_paths(...)) or 0 when None
 * (rail_env_shortest_paths.py:203-274, agent_utils.py:129-136) */
static int greedy_moves(const FoEnv *e, int r, int c, int d, int out_r[3], int out_c[3], int out_d[3]);

static int shortest_path_len(const FoEnv *e, int i) {
    int r, c, d = e->dir[i];
    if (off_map(e->state[i])) { r = e->init_r[i]; c = e->init_c[i]; }
    else if (on_map(e->state[i])) { r = e->r[i]; c = e->c[i]; }
    else { r = e->tgt_r[i]; c = e->tgt_c[i]; }
    int len = 0;
    float distance = INFINITY;
    while (!(r == e->tgt_r[i] && c == e->tgt_c[i])) {
        int mr[3], mc[3], md[3], k = greedy_moves(e, r, c, d, mr, mc, md), best = -1;
        for (int j = 0; j < k; j++) {
            float v = dmv(e, e->slot[i], mr[j], mc[j], md[j]);
            if (v < distance) { best = j; distance = v; }
        }
        len++;
        if (best < 0) return 0; /* path is None */
        r = mr[best]; c = mc[best]; d = md[best];
    }
    return len + 1;
}

/* ------------------------------------------------------------------------------------------- */
/* RailEnv.step (rail_env.py:501-634)                                                           */
/* ------------------------------------------------------------------------------------------- */
int fo_step(FoEnv *e, const uint8_t *actions, const uint8_t *sched, int32_t *rewards, uint8_t *dones) {
    int N = e->N;
    e->elapsed += 1;
    if (e->done_all) return -1;
    int *pa = IALLOC(N), *nr = IALLOC(N), *nc = IALLOC(N), *nd = IALLOC(N);
    int16_t *cur = malloc(sizeof(int16_t) * 2 * N), *nxt = malloc(sizeof(int16_t) * 2 * N);
    uint8_t *can_move = malloc(N);
    for (int i = 0; i < N; i++) rewards[i] = 0;
    for (int i = 0; i < N; i++) {                                   /* loop A, :519-569 */
        e->old_r[i] = e->r[i]; e->old_c[i] = e->c[i]; e->old_dir[i] = e->dir[i];
        if (e->mal[i] == 0 && sched[i] > 0) { e->mal[i] = sched[i]; e->nmal[i] += 1; } /* malfunction_handler.py:35-42 */
        int raw = actions[i] == FO_ACTION_ABSENT ? A_NOTHING : actions[i];
        int a = preprocess_action(e, i, raw);
        if (a >= A_LEFT && a <= A_RIGHT && !e->saved[i] && e->state[i] != DONE) e->saved[i] = a; /* action_saver.py:16-24 */
        int max_count = (int)(1 / e->speed[i]) - 1;                 /* speed_counter.py:40-42 */
        int upd = (e->ctr[i] == max_count) && !(e->mal[i] > 0) && a != A_STOP;
        if (e->r[i] < 0 && e->state[i] != DONE && a == A_STOP) e->saved[i] = 0;   /* :540-542 */
        if (e->state[i] == DONE) { nr[i] = e->r[i]; nc[i] = e->c[i]; nd[i] = e->dir[i]; }
        else if (e->r[i] < 0 && e->saved[i]) { nr[i] = e->init_r[i]; nc[i] = e->init_c[i]; nd[i] = e->init_dir[i]; }
        else if (e->saved[i] && upd) {
            int v, d2 = check_action(e, e->saved[i], e->r[i], e->c[i], e->dir[i], &v); /* env_utils.py:26-43 */
            nr[i] = e->r[i] + DR[d2]; nc[i] = e->c[i] + DC[d2]; nd[i] = d2;
            a = e->saved[i];
        } else { nr[i] = e->r[i]; nc[i] = e->c[i]; nd[i] = e->dir[i]; }
        pa[i] = a;
        cur[2 * i] = e->r[i] < 0 ? -1 : e->r[i]; cur[2 * i + 1] = e->r[i] < 0 ? i : e->c[i];
        nxt[2 * i] = nr[i] < 0 ? -1 : nr[i];     nxt[2 * i + 1] = nr[i] < 0 ? i : nc[i];
    }
    fo_motion_check(N, cur, nxt, can_move);                          /* :572 */
    int all_done = 1;
    for (int i = 0; i < N; i++) {                                   /* loop B, :574-627 */
        int max_count = (int)(1 / e->speed[i]) - 1;
        int exit_ = e->ctr[i] == max_count;
        int allowed = e->mal[i] > 0 ? 0 : can_move[i];
        allowed = allowed || (e->state[i] == STOPPED && !exit_);
        int a = pa[i];
        int in_mal = e->mal[i] > 0, mal_done = e->mal[i] == 0;
        int edr = e->elapsed >= e->earliest[i];
        int stop = a == A_STOP, vm = (a >= A_LEFT && a <= A_RIGHT) && allowed;
        int reached = e->r[i] >= 0 && e->r[i] == e->tgt_r[i] && e->c[i] == e->tgt_c[i];
        int conflict = !allowed && exit_;
        e->sig_mal[i] = in_mal;
        int prev = e->state[i];
        int ns = fsm(prev, in_mal, mal_done, edr, stop, vm, reached, conflict);
        e->state[i] = ns;
        allowed = allowed && ns != DONE;
        if (on_map(ns)) {
            if (off_map(prev)) { e->r[i] = e->init_r[i]; e->c[i] = e->init_c[i]; e->dir[i] = e->init_dir[i]; }
            else if (allowed && exit_) {
                e->r[i] = nr[i]; e->c[i] = nc[i]; e->dir[i] = nd[i];
                if (e->r[i] == e->tgt_r[i] && e->c[i] == e->tgt_c[i]) e->state[i] = DONE; /* update_if_reached */
            }
        }
        if (e->state[i] == DONE && e->arrival[i] < 0) {             /* handle_done_state :493-499 */
            e->arrival[i] = e->elapsed; e->done[i] = 1; e->r[i] = e->c[i] = -1;
        }
        all_done &= e->state[i] == DONE;
        if (e->state[i] == MOVING && e->old_r[i] >= 0) e->ctr[i] = (e->ctr[i] + 1) % (max_count + 1);
        if (e->mal[i] > 0) e->mal[i] -= 1;
        if (e->ctr[i] == 0 && e->r[i] >= 0) e->saved[i] = 0;       /* :626-627 */
    }
    if (all_done || e->elapsed >= e->T) {                           /* end_of_episode_update :476-491 */
        for (int i = 0; i < N; i++) {
            int rew = 0;
            if (e->state[i] == DONE) {
                int dlt = e->latest[i] - e->arrival[i];
                rew = dlt < 0 ? dlt : 0;
            } else {
                int len = shortest_path_len(e, i);
                int tt = (int)ceil((double)len / e->speed[i]);      /* agent_utils.py:129-136 */
                if (off_map(e->state[i])) rew = -1 * tt;
                if (on_map(e->state[i])) rew = (e->latest[i] - e->elapsed) - tt;
            }
            rewards[i] += rew;
            e->done[i] = 1;
        }
        e->done_all = 1;
    }
    for (int i = 0; i < N; i++) dones[i] = e->done[i];
    dones[N] = (uint8_t)e->done_all;
    free(pa); free(nr); free(nc); free(nd); free(cur); free(nxt); free(can_move);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* observation: loader (loader.cpp), deadlocks, predictions, tree, features                     */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int vr, vc;            /* agent_virtual_position (loader.cpp:72-81) */
    int mal01, nmal01;     /* read through py::bool_ (loader.cpp:38-41) */
    float speed, max_count_f;
    float dist_target, initial_dist;
    int trans;             /* full 16-bit cell (0 off-map) */
    int ct[4];             /* cell_transitions of (position, direction) */
    int road_type;
    int va[5];
} AView;

/* rotate_transition (tool.h:300-335) restated with integer ops:
 * inside each orientation block the 4 bits rotate right by k, then the blocks rotate right by k */
static int rotate_transition(int t, int deg) {
    int k = deg / 90, v = 0;
    for (int o = 0; o < 4; o++) {
        int b = (t >> ((3 - o) * 4)) & 0xF;
        b = ((b >> k) | (b << (4 - k))) & 0xF;
        v |= b << ((3 - o) * 4);
    }
    return ((v & ((1 << (k * 4)) - 1)) << ((4 - k) * 4)) | (v >> (k * 4));
}

static const int ROAD_TYPES[11] = { /* loader.cpp:123-134 */
    0x0000, 0x8020, 0x9220, 0x8421, 0x9621, 0xCC33, 0x5202, 0x2000, 0x4002, 0x1200, 0xC022};

static void make_view(const FoEnv *e, int i, AView *v) {
    int s = e->state[i];
    if (off_map(s)) { v->vr = e->init_r[i]; v->vc = e->init_c[i]; }
    else if (on_map(s)) { v->vr = e->r[i]; v->vc = e->c[i]; }
    else { v->vr = e->tgt_r[i]; v->vc = e->tgt_c[i]; }
    v->mal01 = e->mal[i] != 0; v->nmal01 = e->nmal[i] != 0;
    v->speed = (float)e->speed[i];
    v->max_count_f = (float)((int)(1 / e->speed[i]) - 1);
    v->initial_dist = dmv(e, e->slot[i], e->init_r[i], e->init_c[i], e->init_dir[i]); /* loader.cpp:163-179 */
    if (s == DONE) v->dist_target = 0;
    else if (off_map(s)) v->dist_target = v->initial_dist;
    else v->dist_target = dmv(e, e->slot[i], e->r[i], e->c[i], e->dir[i]);
    v->road_type = 0;                                               /* update_transitions, loader.cpp:122-161 */
    if (e->r[i] < 0) { v->trans = 0; v->ct[0] = v->ct[1] = v->ct[2] = v->ct[3] = 0; }
    else {
        v->trans = cell(e, e->r[i], e->c[i]);
        int nb = nib(e, e->r[i], e->c[i], e->dir[i]);
        for (int d = 0; d < 4; d++) v->ct[d] = tbit(nb, d);
        int found = 0;
        for (int rot = 0; rot < 4 && !found; rot++) {
            int t = rot ? rotate_transition(v->trans, 90 * rot) : v->trans;
            for (int k = 0; k < 11; k++) if (ROAD_TYPES[k] == t) { v->road_type = k; found = 1; break; }
        }
    }
    for (int k = 0; k < 5; k++) v->va[k] = 0;                      /* loader.cpp:273-312 */
    if (s == MOVING || s == STOPPED) {
        if (e->ctr[i] == 0) {
            int nb = nib(e, e->r[i], e->c[i], e->dir[i]);
            int branch_next = 0, cnt = 0;
            for (int a = A_LEFT; a <= A_RIGHT; a++) {
                int nd2 = (e->dir[i] + a - 2 + 4) % 4;
                v->va[a] = tbit(nb, nd2);
                if (v->va[a]) {
                    cnt++;
                    if (popc16(cell(e, e->r[i] + DR[nd2], e->c[i] + DC[nd2])) > 2) branch_next = 1;
                }
            }
            if (popc16(cell(e, e->r[i], e->c[i])) > 2 || (cnt == 1 && branch_next)) v->va[A_STOP] = 1;
        } else v->va[A_NOTHING] = 1;
    } else if (s == READY) { v->va[A_FORWARD] = 1; v->va[A_STOP] = 1; }
    else v->va[A_NOTHING] = 1;
}

/* DeadlockChecker (deadlock_checker.cpp:11-110) */
typedef struct { const FoEnv *e; const AView *v; int *apos; int *checked; int *dep; int *ndep; uint8_t *dl; } DL;

static int dl_check_blocked(DL *x, int h) {
    const FoEnv *e = x->e;
    x->checked[h] = 1;
    for (int d = 0; d < 4; d++) {
        if (!x->v[h].ct[d]) continue;
        int rr = e->r[h] + DR[d], cc = e->c[h] + DC[d];
        int opp = in_bounds(e, rr, cc) ? x->apos[rr * e->W + cc] : -1;
        if (opp == -1) { x->checked[h] = 2; return 0; }
        if (x->dl[opp]) continue;
        if (x->checked[opp] == 0) dl_check_blocked(x, opp);
        if (x->checked[opp] == 2 && !x->dl[opp]) { x->checked[h] = 2; return 0; }
        x->dep[h * 4 + x->ndep[h]++] = opp;
    }
    if (x->ndep[h] == 0) {
        x->checked[h] = 2;
        if (x->v[h].ct[0] + x->v[h].ct[1] + x->v[h].ct[2] + x->v[h].ct[3] == 0) return 0;
        x->dl[h] = 1;
        return 1;
    }
    return 0;
}

static void update_deadlocks(FoEnv *e, const AView *v) {
    int N = e->N;
    DL x; x.e = e; x.v = v; x.dl = e->deadlocked;
    x.apos = malloc(sizeof(int) * (size_t)e->H * e->W);
    for (int k = 0; k < e->H * e->W; k++) x.apos[k] = -1;
    x.checked = IALLOC(N); x.dep = IALLOC(4 * N); x.ndep = IALLOC(N);
    for (int i = 0; i < N; i++) if (on_map(e->state[i])) x.apos[e->r[i] * e->W + e->c[i]] = i;
    for (int i = 0; i < N; i++)
        if (on_map(e->state[i]) && !x.dl[i] && !x.checked[i]) dl_check_blocked(&x, i);
    int any = 1;                                                    /* _fix_deps */
    while (any) {
        any = 0;
        for (int h = 0; h < N; h++) {
            if (x.checked[h] != 1) continue;
            int cnt = 0;
            for (int k = 0; k < x.ndep[h]; k++) {
                int o = x.dep[h * 4 + k];
                if (x.checked[o] == 2) {
                    if (x.dl[o]) cnt++;
                    else { x.checked[h] = 2; any = 1; }
                }
            }
            if (cnt == x.ndep[h]) { x.checked[h] = 2; x.dl[h] = 1; any = 1; }
        }
    }
    for (int h = 0; h < N; h++) if (x.checked[h] == 1) { x.dl[h] = 1; x.checked[h] = 2; }
    free(x.apos); free(x.checked); free(x.dep); free(x.ndep);
}

/* get_valid_move_actions_ (predictions.cpp:13-76); result order = std::set order (L,F,R) */
static int greedy_moves(const FoEnv *e, int r, int c, int d, int out_r[3], int out_c[3], int out_d[3]) {
    int nb = nib(e, r, c, d), nt = popc16(nb), k = 0;
    if (popc16(cell(e, r, c)) == 1) {                                /* is_dead_end */
        int ex = (d + 2) % 4;
        if (tbit(nb, ex)) { out_r[k] = r + DR[ex]; out_c[k] = c + DC[ex]; out_d[k] = ex; k++; }
        return k;
    }
    (void)nt; /* both remaining branches enumerate d-1, d, d+1; only the action label differs */
    for (int t = -1; t <= 1; t++) {
        int nd2 = (d + t + 4) % 4;
        if (tbit(nb, nd2)) { out_r[k] = r + DR[nd2]; out_c[k] = c + DC[nd2]; out_d[k] = nd2; k++; }
    }
    return k;
}

/* get_shortest_paths + ShortestPathPredictorForRailEnv::get (predictions.cpp:78-235) for one agent:
 * fills ppos[t][i] = c*W + r and pdir[t][i] for t = 0..500 (treeobs.cpp:50-65, tool.h:368-401) */
static void predict_agent(FoEnv *e, int i, const AView *v) {
    int N = e->N, W = e->W;
    int cap = FO_PRED_DEPTH + 2;
    int *pr = IALLOC(cap), *pc = IALLOC(cap), *pd = IALLOC(cap), len = 0;
    int r = v->vr, c = v->vc, d = e->dir[i], depth = 0;
    float distance = INFINITY;
    int stuck = 0;
    while (depth < FO_PRED_DEPTH) {
        int mr[3], mc[3], md[3], k = greedy_moves(e, r, c, d, mr, mc, md), best = -1;
        for (int j = 0; j < k; j++) {
            float nv = dmv(e, e->slot[i], mr[j], mc[j], md[j]);
            if (nv < distance) { best = j; distance = nv; }
        }
        pr[len] = r; pc[len] = c; pd[len] = d; len++;
        depth++;
        if (best < 0) { stuck = 1; break; }
        r = mr[best]; c = mc[best]; d = md[best];
    }
    if (!stuck) { pr[len] = r; pc[len] = c; pd[len] = d; len++; }
    int tpc = (int)(1 / v->speed);                                   /* predictions.cpp:184 */
    int cr = v->vr, cc = v->vc, cd = e->dir[i], next = 1;           /* path[0] popped */
    e->ppos[0 * N + i] = cc * W + cr; e->pdir[0 * N + i] = cd;
    for (int idx = 0; idx < FO_PRED_DEPTH + 1; idx++) {
        int at_target = (cr == e->tgt_r[i] && cc == e->tgt_c[i]);
        if (!(at_target || next >= len)) {
            if (idx % tpc == 0) { cr = pr[next]; cc = pc[next]; cd = pd[next]; next++; }
        }
        if (idx + 1 < NPRED) { e->ppos[(idx + 1) * N + i] = cc * W + cr; e->pdir[(idx + 1) * N + i] = cd; }
    }
    free(pr); free(pc); free(pd);
}

typedef struct { int r, c, dir, ad, parent; float tot; int is_null; } QCell;

typedef struct {
    const FoEnv *e; const AView *views;
    /* occupancy maps (treeobs.cpp:67-92), keyed by cell */
    uint8_t *has_agent; int *occ_dir; float *occ_speed; int *occ_mal; int *rtd; uint8_t *has_rtd;
    int *stamp; int *stamp_now; /* per-branch visited set (treeobs.cpp:306) as a generation-stamped array */
} ObsCtx;

static void scale_node(const float *in, float *out, float max_dist, int n_agents) { /* treeobs.cpp:111-152 */
    for (int k = 0; k < 7; k++) out[k] = in[k] != INFINITY ? in[k] / (float)max_dist : -1;
    out[7] = in[7] != -1 ? in[7] / (float)n_agents : -1;
    out[8] = in[8] != -1 ? in[8] / (float)n_agents : -1;
    out[9] = in[9] != -1 ? in[9] / (float)n_agents : -1;
    out[10] = in[10] != -1 ? (float)in[10] : -1;
    out[11] = in[11] != -1 ? in[11] / (float)n_agents : -1;
}

/* _explore_branch (treeobs.cpp:258-610).  Returns 0 ok / -1 on "wrong cell type". */
static int explore_branch(const ObsCtx *x, int h, int idx_node, QCell *queue, int *qh, int *qt,
                          float *node, QCell *popped) {
    const FoEnv *e = x->e;
    int N = e->N, W = e->W;
    static const float NULL_NODE[12] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY,
                                        -1, -1, -1, -1, -1};
    if (*qh == *qt) {
        memcpy(node, NULL_NODE, sizeof(NULL_NODE));
        popped->r = popped->c = -1; popped->dir = -1; popped->ad = -1; popped->parent = -2; popped->is_null = 1;
        return 0;
    }
    QCell cur = queue[(*qh)++];
    *popped = cur;
    if (cur.is_null) { memcpy(node, NULL_NODE, sizeof(NULL_NODE)); return 0; }
    int r = cur.r, c = cur.c, d = cur.dir;
    float tot = cur.tot;
    int last_switch = 0, last_dead_end = 0, last_terminal = 0, last_target = 0;
    float time_per_cell = 1.0 / x->views[h].speed;                    /* double division, rounded to float */
    float own_target = INFINITY, other_agent = INFINITY, other_target = INFINITY, conflict = INFINITY,
          unusable = INFINITY, min_speed = 1.0;
    int same = 0, opp = 0, malf = 0, rtd_n = 0;
    int gen = ++*x->stamp_now;
    int exploring = 1, ct[4];
    while (exploring) {
        int id = r * W + c;
        if (x->has_agent[id]) {                                        /* :322-360 */
            if (tot < other_agent) other_agent = tot;
            if (x->occ_mal[id] > malf) malf = x->occ_mal[id];
            rtd_n += x->has_rtd[id] ? x->rtd[id] : 0;
            if (x->occ_dir[id] == d) {
                same += 1;
                if (x->occ_speed[id] < min_speed) min_speed = x->occ_speed[id];
            } else opp += 1;
        }
        int nb = nib(e, r, c, d);
        for (int k = 0; k < 4; k++) ct[k] = tbit(nb, k);
        int tb = cell(e, r, c), total = popc16(tb), crossing = tb == 0x8421;
        int pt = (int)((int)tot * time_per_cell);                      /* :378 */
        if (pt < NPRED) {
            int ipos = c * W + r;
            if (tot < NPRED) {
                int pre = pt - 1 < 0 ? 0 : pt - 1, post = NPRED - 1 < pt + 1 ? NPRED - 1 : pt + 1;
                int steps[3] = {pt, pre, post};
                for (int si = 0; si < 3; si++) {                       /* first matching time slice wins */
                    const int *row = e->ppos + (size_t)steps[si] * N;
                    int hit = 0;
                    for (int j = 0; j < N; j++) if (j != h && row[j] == ipos) { hit = 1; break; }
                    if (!hit) continue;
                    for (int ca = 0; ca < N; ca++) {
                        if (row[ca] != ipos) continue;
                        int pd = e->pdir[(size_t)pt * N + ca];         /* always row predicted_time */
                        if (d != pd && ct[(pd + 2) % 4] == 1 && tot < conflict) conflict = tot;
                        if (e->state[ca] == DONE && tot < conflict) conflict = tot;
                    }
                    break;
                }
            }
        }
        /* location_has_target is never filled (treeobs.cpp:72) -> other_target stays inf */
        int is_target = (r == e->tgt_r[h] && c == e->tgt_c[h]);
        if (is_target && tot < own_target) own_target = tot;
        if (x->stamp[(size_t)id * 4 + d] == gen) { last_terminal = 1; break; }
        x->stamp[(size_t)id * 4 + d] = gen;
        if (is_target) { last_target = 1; break; }
        if (crossing) total = 2;
        int num = ct[0] + ct[1] + ct[2] + ct[3];
        exploring = 0;
        if (total > 2 && 2 > num && tot < unusable) unusable = tot;
        if (num == 1) {
            if (total == 1) last_dead_end = 1;
            if (!last_dead_end) {
                exploring = 1;
                d = ct[0] ? 0 : ct[1] ? 1 : ct[2] ? 2 : 3;
                r += DR[d]; c += DC[d];
                tot += 1;
            }
        } else if (num > 0) { last_switch = 1; break; }
        else return -1;
    }
    float dnb, dmin;
    if (last_target) { dnb = tot; dmin = 0; }
    else if (last_terminal) { dnb = INFINITY; dmin = dmv(e, e->slot[h], r, c, d); }
    else { dnb = tot; dmin = dmv(e, e->slot[h], r, c, d); }
    node[0] = own_target; node[1] = other_target; node[2] = other_agent; node[3] = conflict;
    node[4] = unusable; node[5] = dnb; node[6] = dmin; node[7] = (float)same; node[8] = (float)opp;
    node[9] = (float)malf; node[10] = min_speed; node[11] = (float)rtd_n;
    int nb2 = nib(e, r, c, d);
    for (int ad = -1; ad <= 1; ad++) {                                /* :583-608 */
        int bd = (d + 4 + ad) % 4, rb = (bd + 2) % 4;
        QCell q; q.ad = ad; q.parent = idx_node; q.tot = tot + 1; q.is_null = 0;
        if (last_dead_end && tbit(nb2, rb)) { q.r = r + DR[rb]; q.c = c + DC[rb]; q.dir = rb; }
        else if (last_switch && tbit(nb2, bd)) { q.r = r + DR[bd]; q.c = c + DC[bd]; q.dir = bd; }
        else { q.r = q.c = -1; q.dir = bd; q.is_null = 1; }
        queue[(*qt)++] = q;
    }
    return 0;
}

/* calculate_evaluation_orders (tool.h:468-524) */
static void evaluation_orders(const int32_t *adj, int n_edges, int tree_size, int32_t *node_order, int32_t *edge_order) {
    uint8_t uneval[FO_MAX_NODES + 1] = {0};
    int count = 0;
    for (int k = 0; k < tree_size; k++) node_order[k] = 0;
    for (int k = 0; k < n_edges; k++) {
        int p = adj[3 * k], ch = adj[3 * k + 1];
        if (p != -2 && !uneval[p]) { uneval[p] = 1; count++; }
        if (ch != -2 && !uneval[ch]) { uneval[ch] = 1; count++; }
    }
    for (int k = count; k < tree_size; k++) node_order[k] = -2;
    int remaining = count, order = 0;
    while (remaining > 0) {
        uint8_t unready[FO_MAX_NODES + 1] = {0};
        for (int k = 0; k < n_edges; k++) {
            int p = adj[3 * k], ch = adj[3 * k + 1];
            if (ch != -2 && uneval[ch] && p >= 0) unready[p] = 1;
        }
        for (int nidx = 0; nidx < tree_size; nidx++)
            if (uneval[nidx] && !unready[nidx]) { node_order[nidx] = order; uneval[nidx] = 0; remaining--; }
        order++;
    }
    for (int k = 0; k < n_edges; k++) {
        int p = adj[3 * k];
        edge_order[k] = p < 0 ? -2 : node_order[p];
    }
}

/* TreeObsForRailEnv::get (treeobs.cpp:154-256) */
static int tree_for_agent(const ObsCtx *x, int h, float *forest, int32_t *adj, int32_t *node_order, int32_t *edge_order) {
    const FoEnv *e = x->e;
    const AView *v = &x->views[h];
    int nb = nib(e, v->vr, v->vc, e->dir[h]), num = popc16(nb);
    float root[12] = {0, 0, 0, 0, 0, 0, v->dist_target, 0, 0, (float)v->nmal01, v->speed, 0};
    scale_node(root, forest, (float)e->T, e->N);
    int orientation = e->dir[h];
    if (num == 1) orientation = tbit(nb, 0) ? 0 : tbit(nb, 1) ? 1 : tbit(nb, 2) ? 2 : 3;
    QCell queue[3 + 3 * FO_MAX_NODES];
    int qh = 0, qt = 0;
    for (int ad = -1; ad <= 1; ad++) {
        int bd = (orientation + ad + 4) % 4;
        QCell q; q.dir = bd; q.ad = ad; q.parent = 0; q.tot = 1.0; q.is_null = 0;
        if (tbit(nb, bd)) { q.r = v->vr + DR[bd]; q.c = v->vc + DC[bd]; }
        else { q.r = q.c = -1; q.is_null = 1; }
        queue[qt++] = q;
    }
    for (int n = 1; n < FO_MAX_NODES; n++) {
        float raw[12];
        QCell popped;
        if (explore_branch(x, h, n, queue, &qh, &qt, raw, &popped)) return -1;
        scale_node(raw, forest + n * 12, (float)e->T, e->N);
        int idx = n, ad = popped.ad;
        if (popped.parent == -2) { idx = -2; ad = -2; }
        adj[3 * (n - 1)] = popped.parent; adj[3 * (n - 1) + 1] = idx; adj[3 * (n - 1) + 2] = ad;
    }
    evaluation_orders(adj, FO_MAX_NODES - 1, FO_MAX_NODES, node_order, edge_order);
    return 0;
}

/* AgentAttrParser::get_features (feature_parser.cpp:3-98) */
static void agent_features(const FoEnv *e, int h, const AView *v, float *out) {
    int k = 0;
#define ONEHOT(len, pos) do { for (int q_ = 0; q_ < (len); q_++) out[k++] = (q_ == (pos)) ? 1.0f : 0.0f; } while (0)
    int s = e->state[h];
    ONEHOT(7, s);
    ONEHOT(11, v->road_type);
    ONEHOT(10, v->nmal01);
    ONEHOT(4, e->init_dir[h]);
    ONEHOT(4, e->dir[h]);
    ONEHOT(4, e->old_dir[h] < 0 ? e->dir[h] : e->old_dir[h]);
#undef ONEHOT
    int max_count = (int)(1 / e->speed[h]) - 1;
    out[k++] = (float)(s == MOVING);
    out[k++] = (float)e->deadlocked[h];
    out[k++] = (float)e->sig_mal[h];
    out[k++] = (float)(e->mal[h] == 0);
    out[k++] = (float)(e->ctr[h] == 0);
    out[k++] = (float)(e->ctr[h] == max_count);
    out[k++] = (float)(s == MALFUNCTION || s == MAL_OFF);
    out[k++] = (float)off_map(s);
    out[k++] = (float)on_map(s);
    for (int b = 15; b >= 0; b--) out[k++] = (float)((v->trans >> b) & 1);
    for (int a = 0; a < 5; a++) out[k++] = (float)v->va[a];
    float max_dist_target = (float)((e->H + e->W) * 8);
    float T = (float)e->T;
    float agent_handle = (float)h / (float)e->N;
    float curr_step = (float)e->elapsed / T;
    float earliest = (float)e->earliest[h] / T;
    float latest = (float)e->latest[h] / T;
    float arrival = (float)e->arrival[h] / T;
    float before_late = latest - curr_step;
    float dist_target = v->dist_target == INFINITY ? 8.0f : v->dist_target / max_dist_target;
    float anticipative = before_late < dist_target ? before_late : dist_target;
    float smc = v->max_count_f / 10;
    float speed = v->speed / 1.0f;
    float sctr = (float)e->ctr[h] / 10;
    float malc = (float)v->mal01 / 10;
    float idist = v->initial_dist == INFINITY ? 8.0f : v->initial_dist / max_dist_target;
    out[k++] = agent_handle; out[k++] = curr_step; out[k++] = earliest; out[k++] = latest; out[k++] = arrival;
    out[k++] = before_late; out[k++] = dist_target; out[k++] = anticipative; out[k++] = smc; out[k++] = speed;
    out[k++] = sctr; out[k++] = malc; out[k++] = idist;
}

int fo_obs(FoEnv *e, float *attr, float *forest, int32_t *adjacency, int32_t *node_order,
           int32_t *edge_order, uint8_t *valid_actions, float *dist_target, uint8_t *deadlocked) {
    int N = e->N, rc = 0;
    size_t ncells = (size_t)e->H * e->W;
    AView *views = malloc(sizeof(AView) * N);
    for (int i = 0; i < N; i++) make_view(e, i, &views[i]);           /* AgentsLoader::update */
    update_deadlocks(e, views);
    for (int i = 0; i < N; i++) predict_agent(e, i, &views[i]);
    ObsCtx x; x.e = e; x.views = views;
    x.has_agent = calloc(ncells, 1); x.occ_dir = IALLOC(ncells); x.occ_speed = calloc(ncells, sizeof(float));
    x.occ_mal = IALLOC(ncells); x.rtd = IALLOC(ncells); x.has_rtd = calloc(ncells, 1);
    int stamp_now = 0;
    x.stamp = IALLOC(ncells * 4); x.stamp_now = &stamp_now;
    for (int i = 0; i < N; i++) {                                      /* treeobs.cpp:74-92 */
        if (!off_map(e->state[i]) && e->r[i] != -1) {
            int id = e->r[i] * e->W + e->c[i];
            x.has_agent[id] = 1; x.occ_dir[id] = e->dir[i]; x.occ_speed[id] = views[i].speed;
            x.occ_mal[id] = views[i].mal01;
        }
        if (off_map(e->state[i])) {
            int id = e->init_r[i] * e->W + e->init_c[i];
            if (x.has_rtd[id]) x.rtd[id] += 1; else { x.has_rtd[id] = 1; x.rtd[id] = 0; }
        }
    }
    float tf[FO_MAX_NODES * 12];
    int32_t ta[(FO_MAX_NODES - 1) * 3], tn[FO_MAX_NODES], te[FO_MAX_NODES - 1];
    for (int h = 0; h < N; h++) {
        if (tree_for_agent(&x, h, tf, ta, tn, te)) { rc = -1; break; }
        if (forest) memcpy(forest + (size_t)h * FO_MAX_NODES * 12, tf, sizeof(tf));
        if (adjacency) memcpy(adjacency + (size_t)h * (FO_MAX_NODES - 1) * 3, ta, sizeof(ta));
        if (node_order) memcpy(node_order + (size_t)h * FO_MAX_NODES, tn, sizeof(tn));
        if (edge_order) memcpy(edge_order + (size_t)h * (FO_MAX_NODES - 1), te, sizeof(te));
    }
    for (int h = 0; h < N; h++) {
        if (attr) agent_features(e, h, &views[h], attr + (size_t)h * FO_ATTR_F);
        if (valid_actions) for (int a = 0; a < 5; a++) valid_actions[h * 5 + a] = (uint8_t)views[h].va[a];
        if (dist_target) dist_target[h] = views[h].dist_target;
        if (deadlocked) deadlocked[h] = e->deadlocked[h];
    }
    free(x.has_agent); free(x.occ_dir); free(x.occ_speed); free(x.occ_mal); free(x.rtd); free(x.has_rtd);
    free(x.stamp);
    free(views);
    return rc;
}

/* ------------------------------------------------------------------------------------------- */
/* read-back                                                                                    */
/* ------------------------------------------------------------------------------------------- */
void fo_get_state(const FoEnv *e, int16_t *pos, uint8_t *dir, uint8_t *state, uint8_t *ctr, uint8_t *mal,
                  uint16_t *nmal, uint8_t *saved, int32_t *arrival, int16_t *old_pos, int8_t *old_dir,
                  uint8_t *sig_mal) {
    for (int i = 0; i < e->N; i++) {
        if (pos) { pos[2 * i] = (int16_t)e->r[i]; pos[2 * i + 1] = (int16_t)e->c[i]; }
        if (dir) dir[i] = (uint8_t)e->dir[i];
        if (state) state[i] = (uint8_t)e->state[i];
        if (ctr) ctr[i] = (uint8_t)e->ctr[i];
        if (mal) mal[i] = (uint8_t)e->mal[i];
        if (nmal) nmal[i] = (uint16_t)e->nmal[i];
        if (saved) saved[i] = (uint8_t)e->saved[i];
        if (arrival) arrival[i] = e->arrival[i];
        if (old_pos) { old_pos[2 * i] = (int16_t)e->old_r[i]; old_pos[2 * i + 1] = (int16_t)e->old_c[i]; }
        if (old_dir) old_dir[i] = (int8_t)e->old_dir[i];
        if (sig_mal) sig_mal[i] = (uint8_t)e->sig_mal[i];
    }
}
int fo_elapsed(const FoEnv *e) { return e->elapsed; }
int fo_done_all(const FoEnv *e) { return e->done_all; }
int fo_num_targets(const FoEnv *e) { return e->n_slots; }
int fo_target_slot(const FoEnv *e, int agent) { return e->slot[agent]; }
void fo_get_dist_u16(const FoEnv *e, int slot, uint16_t *out) {
    size_t n = (size_t)e->H * e->W * 4;
    const float *m = e->dm + (size_t)slot * n;
    for (size_t i = 0; i < n; i++) out[i] = m[i] == INFINITY ? 65535 : (uint16_t)m[i];
}

/* ------------------------------------------------------------------------------------------- */
/* CPU baseline driver (bench.py cpu_baseline leg only)                                         */
/* ------------------------------------------------------------------------------------------- */
typedef struct { FoEnv **envs; int lo, hi, n_steps; uint32_t seed; long long agent_steps; } Work;

static uint32_t xorshift(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }

static void *bench_worker(void *arg) {
    Work *w = arg;
    for (int k = w->lo; k < w->hi; k++) {
        FoEnv *e = w->envs[k];
        int N = e->N;
        uint8_t *act = malloc(N), *sched = calloc(N, 1), *dones = malloc(N + 1);
        int32_t *rew = malloc(sizeof(int32_t) * N);
        float *attr = malloc(sizeof(float) * N * FO_ATTR_F), *forest = malloc(sizeof(float) * N * FO_MAX_NODES * 12);
        int32_t *adj = malloc(sizeof(int32_t) * N * 90), *no = malloc(sizeof(int32_t) * N * 31), *eo = malloc(sizeof(int32_t) * N * 30);
        uint8_t *va = malloc(N * 5);
        uint32_t s = w->seed * 2654435761u + (uint32_t)k * 40503u + 1u;
        for (int t = 0; t < w->n_steps; t++) {
            if (e->done_all) fo_reset(e);
            for (int i = 0; i < N; i++) act[i] = (uint8_t)(xorshift(&s) % 5);
            fo_step(e, act, sched, rew, dones);
            fo_obs(e, attr, forest, adj, no, eo, va, NULL, NULL);
            w->agent_steps += N;
        }
        free(act); free(sched); free(dones); free(rew); free(attr); free(forest); free(adj); free(no); free(eo); free(va);
    }
    return NULL;
}

long long fo_bench_run(FoEnv **envs, int n_envs, int n_steps, int n_threads, uint32_t seed) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_envs) n_threads = n_envs;
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    Work *w = calloc(n_threads, sizeof(Work));
    for (int t = 0; t < n_threads; t++) {
        w[t].envs = envs; w[t].lo = (int)((long long)n_envs * t / n_threads);
        w[t].hi = (int)((long long)n_envs * (t + 1) / n_threads);
        w[t].n_steps = n_steps; w[t].seed = seed;
        pthread_create(&th[t], NULL, bench_worker, &w[t]);
    }
    long long total = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); total += w[t].agent_steps; }
    free(th); free(w);
    return total;
}
