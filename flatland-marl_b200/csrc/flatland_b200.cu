// flatland_b200.cu — hand-written sm_100a kernels + the C ABI declared in include/flatland_b200.h.
//
// Four kernels on the hot path, all integer / byte work bound by dependent-load latency and HBM/L2,
// no dense contraction anywhere (so no tensor cores):
//   k_step   one CTA per environment, one thread per agent: RailEnv.step up to the observation
//            (rail_env.py:501-632): action preprocessing, MotionCheck as a least fixpoint resolved
//            with shared-memory broadcasts, state machine, counters, end-of-episode rewards.
//   k_prep   one CTA per environment: loader view + valid actions (loader.cpp:221-327), the serial
//            sticky deadlock checker on a spare thread (deadlock_checker.cpp), greedy shortest-path
//            predictions (predictions.cpp) turned into an inverse index  cell id -> occupancy
//            intervals  by a counting sort in shared memory.
//   k_tree   one warp per agent: the 31-node branch tree (treeobs.cpp:154-610), 83-float agent
//            attributes (feature_parser.cpp) and evaluation orders (tool.h:468-524), staged in shared
//            memory and written as coalesced vector stores in the policy's input layout.
//   k_bfs    one CTA per (environment, unique target): DistanceMap (distance_map.py:57-160) as a
//            level-synchronous pull BFS with the whole map in shared memory.  Reset-time only.
// Reference citations are relative to the reference repository root.  No fast-math: every float
// expression mirrors one float expression of the reference.
#include "flatland_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <vector>

#define DEVI __device__ __forceinline__

namespace {

enum : int { WAITING = 0, READY = 1, MAL_OFF = 2, MOVING = 3, STOPPED = 4, MALFUNCTION = 5, DONE = 6 };
enum : int { A_NOTHING = 0, A_LEFT = 1, A_FORWARD = 2, A_RIGHT = 3, A_STOP = 4 };
constexpr int NPRED = FL_PRED_DEPTH + 1;  // prediction rows 0..500 (treeobs.cpp:50-65)
constexpr int TREE_WARPS = 8;

std::atomic<uint64_t> g_launches{0};

// Optional per-kernel timing (fl_profile_*): every launch is bracketed by CUDA events recorded on the
// launching stream; fl_profile_collect turns them into per-kernel totals.  Off by default.
enum KernelId : int { K_BFS = 0, K_RESET, K_STEP, K_PREP, K_TREE, K_COUNT };
const char *const kKernelNames[K_COUNT] = {"k_bfs", "k_reset", "k_step", "k_prep", "k_tree"};
struct ProfRec { int id; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_pending;
std::vector<cudaEvent_t> g_prof_free;
double g_prof_ms[K_COUNT];
uint64_t g_prof_n[K_COUNT];

struct LaunchScope {  // RAII: counts the launch and, when profiling, records the event pair around it
    int id; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool on;
    static cudaEvent_t get() {
        if (!g_prof_free.empty()) { cudaEvent_t e = g_prof_free.back(); g_prof_free.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    LaunchScope(int id_, cudaStream_t st_) : id(id_), st(st_) {
        g_launches++;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        on = g_prof_on;
        if (on) { a = get(); b = get(); cudaEventRecord(a, st); }
    }
    ~LaunchScope() {
        if (!on) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_pending.push_back({id, a, b});
    }
};

DEVI bool on_map(int s) { return s >= MOVING && s <= MALFUNCTION; }
DEVI bool off_map(int s) { return s <= MAL_OFF; }
DEVI int d_row(int d) { return (d == 2) - (d == 0); }
DEVI int d_col(int d) { return (d == 1) - (d == 3); }
// grid4.py:66-87 / tool.h:337-352: 4-bit nibble of orientation o, bit order N,E,S,W msb->lsb
DEVI int nibble(unsigned cell, int o) { return (cell >> ((3 - o) * 4)) & 0xF; }
DEVI int tbit(int nb, int d) { return (nb >> (3 - d)) & 1; }
DEVI int first_dir(int nb) { return __clz(nb) - 28; }  // first set bit in N,E,S,W order, nb in 1..15

// ---------------------------------------------------------------------------------------------
// action preprocessing (transition_utils.py:6-82)
// ---------------------------------------------------------------------------------------------
DEVI int check_action(const uint16_t *__restrict__ g, int W, int a, int r, int c, int d, int &valid) {
    const int nb = nibble(__ldg(g + r * W + c), d), nt = __popc(nb);
    int nd = d;
    valid = -1;
    if (a == A_LEFT) { nd = d - 1; if (nt <= 1) valid = 0; }
    else if (a == A_RIGHT) { nd = d + 1; if (nt <= 1) valid = 0; }
    nd &= 3;
    if (a == A_FORWARD && nt == 1) { nd = first_dir(nb); valid = 1; }
    return nd;
}

DEVI bool check_valid_action(const uint16_t *__restrict__ g, int H, int W, int a, int r, int c, int d) {
    int valid;
    const int nd = check_action(g, W, a, r, c, d, valid);
    const int rr = r + d_row(nd), cc = c + d_col(nd);
    const bool cell_ok = rr >= 0 && cc >= 0 && rr < H && cc < W && __ldg(g + rr * W + cc) > 0;
    if (valid < 0) valid = tbit(nibble(__ldg(g + r * W + c), d), nd);
    return cell_ok && valid;
}

// step_utils/state_machine.py:12-80
DEVI int fsm(int s, bool in_mal, bool mal_done, bool edr, bool stop, bool valid_move, bool reached, bool conflict) {
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

// EnvAgent.reset for every agent of env e + cleared maps (agent_utils.py:90-105, rail_env.py:335-344,
// treeobs.cpp:22-28).  Called by all threads of a CTA.
DEVI void reset_env(const FlBatch &b, int e, bool rewind_schedule) {
    const int N = (int)b.N, HW = (int)(b.H * b.W);
    uint32_t *ci = b.cellinfo + (size_t)e * HW;
    for (int k = threadIdx.x; k < HW; k += blockDim.x) ci[k] = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const size_t ea = (size_t)e * N + i;
        b.rc[2 * ea] = -1; b.rc[2 * ea + 1] = -1;
        b.old_rc[2 * ea] = -1; b.old_rc[2 * ea + 1] = -1;
        b.dir[ea] = b.init_dir[ea]; b.old_dir[ea] = 255;
        b.state[ea] = WAITING; b.ctr[ea] = 0; b.mal[ea] = 0; b.saved[ea] = 0; b.sig_mal[ea] = 0;
        b.deadlocked[ea] = 0; b.done[ea] = 0; b.nmal[ea] = 0; b.arrival[ea] = -1;
        b.occ_cell[ea] = -1;
    }
    if (threadIdx.x == 0) {
        b.elapsed[e] = 0; b.done_all[e] = 0;
        if (rewind_schedule) { b.sched_pos[e] = 0; b.status[e] = 0; }
    }
}

__global__ void k_reset(FlBatch b, const uint8_t *__restrict__ mask) {
    const int e = blockIdx.x;
    if (mask && !mask[e]) return;
    reset_env(b, e, true);
}

// ---------------------------------------------------------------------------------------------
// k_step: RailEnv.step (rail_env.py:501-632)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_step(FlBatch b, const uint8_t *__restrict__ actions, int32_t *__restrict__ rewards,
       uint8_t *__restrict__ dones, uint32_t flags) {
    const int e = blockIdx.x, N = (int)b.N, H = (int)b.H, W = (int)b.W, HW = H * W;
    const int i = threadIdx.x;
    const bool act = i < N;
    extern __shared__ int sm[];
    int *s_cur = sm, *s_nxt = sm + N, *s_rep = sm + 2 * N, *s_blk = sm + 3 * N;
    const size_t ea = (size_t)e * N + (act ? i : 0);
    const uint16_t *__restrict__ g = b.grid + (size_t)e * HW;

    const bool was_done = b.done_all[e] != 0;
    const int elapsed = b.elapsed[e] + 1;
    const int srow = b.sched_pos[e] % (int)b.S;
    __syncthreads();
    if (was_done) {  // rail_env.py:508-509 raises; here: status bit, or in-place reset
        if (flags & FL_FLAG_AUTO_RESET) {
            reset_env(b, e, false);
            if (i == 0) atomicOr(&b.status[e], FL_ST_AUTO_RESET);
            if (act) { rewards[ea] = 0; dones[(size_t)e * (N + 1) + i] = 0; }
            if (i == 0) dones[(size_t)e * (N + 1) + N] = 0;
        } else {
            if (i == 0) atomicOr(&b.status[e], FL_ST_STEP_AFTER_DONE);
            if (act) { rewards[ea] = 0; dones[(size_t)e * (N + 1) + i] = 1; }
            if (i == 0) dones[(size_t)e * (N + 1) + N] = 1;
        }
        return;
    }

    // ---- loop A (rail_env.py:519-569): independent per agent -------------------------------
    int r = -1, c = -1, d = 0, st = DONE, ctr = 0, mal = 0, saved = 0, nmal = 0, arrival = 0;
    int ir = 0, ic = 0, idir = 0, tr = -1, tc = -1, maxc = 0;
    int old_r = -1, old_c = -1, old_d = 0, a = 0, nr = -1, nc = -1, nd = 0, cur_id = -1 - i, nxt_id = -1 - i;
    if (act) {
        const short2 p = reinterpret_cast<const short2 *>(b.rc)[ea];
        r = p.x; c = p.y; d = b.dir[ea]; st = b.state[ea]; ctr = b.ctr[ea]; mal = b.mal[ea];
        saved = b.saved[ea]; nmal = b.nmal[ea]; arrival = b.arrival[ea];
        const short2 ip = reinterpret_cast<const short2 *>(b.init_rc)[ea];
        const short2 tp = reinterpret_cast<const short2 *>(b.tgt_rc)[ea];
        ir = ip.x; ic = ip.y; idir = b.init_dir[ea]; tr = tp.x; tc = tp.y; maxc = b.max_count[ea];
        old_r = r; old_c = c; old_d = d;
        const int sv = b.sched[((size_t)e * b.S + srow) * N + i];  // malfunction_handler.py:35-42
        if (mal == 0 && sv > 0) { mal = sv; nmal += 1; }
        const int raw = actions[ea];
        a = raw <= 4 ? raw : A_NOTHING;                             // action_preprocessing.py:7-21
        if (a == A_NOTHING) { if (st == MOVING) a = A_FORWARD; else if (saved) a = saved; }
        if (st == WAITING) a = A_NOTHING;
        int pr = r, pc = c, pd = d;
        if (r < 0) { pr = ir; pc = ic; pd = idir; }
        if ((a == A_LEFT || a == A_RIGHT) && !check_valid_action(g, H, W, a, pr, pc, pd)) a = A_FORWARD;
        if (a >= A_LEFT && a <= A_RIGHT && !check_valid_action(g, H, W, a, pr, pc, pd)) a = A_STOP;
        if (a >= A_LEFT && a <= A_RIGHT && !saved && st != DONE) saved = a;     // action_saver.py:16-24
        const bool upd = ctr == maxc && mal == 0 && a != A_STOP;                // rail_env.py:535-537
        if (r < 0 && st != DONE && a == A_STOP) saved = 0;                      // rail_env.py:540-542
        if (st == DONE) { nr = r; nc = c; nd = d; }
        else if (r < 0 && saved) { nr = ir; nc = ic; nd = idir; }
        else if (saved && upd) {
            int v;
            nd = check_action(g, W, saved, r, c, d, v);                         // env_utils.py:26-43
            nr = r + d_row(nd); nc = c + d_col(nd);
            a = saved;
        } else { nr = r; nc = c; nd = d; }
        cur_id = r < 0 ? -1 - i : r * W + c;       // agent_chains.py:28-33: off-map = private node
        nxt_id = nr < 0 ? -1 - i : nr * W + nc;
        s_cur[i] = cur_id; s_nxt[i] = nxt_id; s_blk[i] = 0;
    }
    __syncthreads();

    // ---- MotionCheck (agent_chains.py:151-236) as a least fixpoint over CELL NODES:
    //        blocked(X) = some train on X stays | swaps | loses a contended cell | heads for a blocked node
    //      Several trains can share a cell (MALFUNCTION_OFF_MAP + STOP enters the map unchecked,
    //      state_machine.py:41-42); a node's verdict is shared by all of them and its "agent" attribute is
    //      the last one added, i.e. the highest handle (agent_chains.py:33) — its representative here.
    int rep_cur = i, rep_nxt = -1;
    bool sw = false;
    if (act) {
        for (int k = 0; k < N; k++) {              // shared-memory broadcasts, no bank conflicts
            const int ck = s_cur[k], nk = s_nxt[k];
            if (ck == cur_id) rep_cur = max(rep_cur, k);
            if (ck == nxt_id) { rep_nxt = max(rep_nxt, k); if (nk == cur_id && nxt_id != cur_id) sw = true; }
        }
        s_rep[i] = rep_cur;
    }
    __syncthreads();
    if (act) {
        bool loser = false;                        // another node wants my target and its agent index is lower
        if (nxt_id != cur_id)
            for (int k = 0; k < N; k++) {
                const int ck = s_cur[k];
                if (s_nxt[k] == nxt_id && ck != cur_id && ck != nxt_id && s_rep[k] < rep_cur) loser = true;
            }
        if (nxt_id == cur_id || sw || loser) s_blk[rep_cur] = 1;
    }
    __syncthreads();
    while (true) {                                 // propagate along chains until nothing changes
        int changed = 0;
        if (act && nxt_id != cur_id && rep_nxt >= 0 && !s_blk[rep_cur] && s_blk[rep_nxt]) { s_blk[rep_cur] = 1; changed = 1; }
        if (!__syncthreads_or(changed)) break;
    }
    const bool blocked = act ? s_blk[rep_cur] != 0 : true;

    // ---- loop B (rail_env.py:574-627) --------------------------------------------------------
    if (act) {
        const bool exit_ = ctr == maxc;
        bool allowed = (mal > 0 ? false : !blocked) || (st == STOPPED && !exit_);
        const bool in_mal = mal > 0, mal_done = mal == 0, edr = elapsed >= b.earliest[ea];
        const bool stop_given = a == A_STOP, vm = (a >= A_LEFT && a <= A_RIGHT) && allowed;
        const bool reached = r >= 0 && r == tr && c == tc;
        const bool conflict = !allowed && exit_;
        const int prev = st;
        st = fsm(prev, in_mal, mal_done, edr, stop_given, vm, reached, conflict);
        allowed = allowed && st != DONE;
        if (on_map(st)) {
            if (off_map(prev)) { r = ir; c = ic; d = idir; }
            else if (allowed && exit_) {
                r = nr; c = nc; d = nd;
                if (r == tr && c == tc) st = DONE;                              // update_if_reached
            }
        }
        if (st == DONE && arrival < 0) { arrival = elapsed; b.done[ea] = 1; r = -1; c = -1; }  // :493-499
        if (st == MOVING && old_r >= 0) ctr = (ctr + 1) % (maxc + 1);           // speed_counter.py:10-14
        b.sig_mal[ea] = in_mal;
        if (mal > 0) mal -= 1;
        if (ctr == 0 && r >= 0) saved = 0;                                      // rail_env.py:626-627
    }
    const int all_done = __syncthreads_and(!act || st == DONE);
    // ---- end of episode (rail_env.py:476-491, 397-423; agent_utils.py:129-147) ---------------
    const bool ended = all_done || elapsed >= b.max_steps[e];
    if (act) {
        int rew = 0;
        if (ended) {
            if (st == DONE) rew = min(b.latest[ea] - arrival, 0);
            else {
                // len(shortest path) = dist + 1 when reachable, 0 (path None) otherwise
                const int qr = r < 0 ? ir : r, qc = r < 0 ? ic : c;
                const unsigned dv = b.dist[(((size_t)e * b.n_slots + b.slot[ea]) * HW + qr * W + qc) * 4 + d];
                const int len = dv == FL_DIST_INF ? 0 : (int)dv + 1;
                const int tt = (int)ceil((double)len / b.speed[ea]);
                rew = off_map(st) ? -tt : (b.latest[ea] - elapsed) - tt;
            }
            b.done[ea] = 1;
            // episode statistics (eval_env.py:81-94 final_metric): arrivals and total reward
            unsigned long long *stt = reinterpret_cast<unsigned long long *>(b.stats + (size_t)e * 4);
            if (st == DONE) atomicAdd(stt + 1, 1ull);
            if (rew) atomicAdd(stt + 2, (unsigned long long)(long long)rew);
        }
        rewards[ea] = rew;
        dones[(size_t)e * (N + 1) + i] = ended ? 1 : b.done[ea];
        reinterpret_cast<short2 *>(b.rc)[ea] = make_short2((short)r, (short)c);
        reinterpret_cast<short2 *>(b.old_rc)[ea] = make_short2((short)old_r, (short)old_c);
        b.dir[ea] = d; b.old_dir[ea] = old_d; b.state[ea] = st; b.ctr[ea] = ctr; b.mal[ea] = mal;
        b.saved[ea] = saved; b.nmal[ea] = nmal; b.arrival[ea] = arrival;
    }
    if (i == 0) {
        b.elapsed[e] = elapsed;
        b.sched_pos[e] = b.sched_pos[e] + 1;
        if (ended) { b.done_all[e] = 1; atomicAdd(reinterpret_cast<unsigned long long *>(b.stats + (size_t)e * 4), 1ull); }
        atomicAdd(reinterpret_cast<unsigned long long *>(b.stats + (size_t)e * 4 + 3), (unsigned long long)N);
        dones[(size_t)e * (N + 1) + N] = ended ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// k_prep: loader view, deadlocks, shortest-path predictions -> inverse index
// ---------------------------------------------------------------------------------------------
// get_valid_move_actions_ (predictions.cpp:13-76), result in std::set order L,F,R
DEVI int greedy_moves(unsigned cell, int d, int out_d[3]) {
    const int nb = nibble(cell, d);
    int k = 0;
    if (__popc(cell) == 1) {                      // dead end: only way is back
        const int ex = (d + 2) & 3;
        if (tbit(nb, ex)) out_d[k++] = ex;
        return k;
    }
#pragma unroll
    for (int t = -1; t <= 1; t++) {
        const int nd = (d + t) & 3;
        if (tbit(nb, nd)) out_d[k++] = nd;
    }
    return k;
}

// Predicted occupancy of one agent as intervals per path element (predictions.cpp:78-235 + the
// transpose in treeobs.cpp:50-65).  Path element k >= 1 is occupied for prediction rows
// [1+(k-1)*tpc, k*tpc], the last element until row 500, element 0 for row 0 only (or all rows when
// the path has a single element).  The reference stops advancing once the cell equals the target.
// Emit(key, t0, t1, dir_here, dir_prev, dir_next) is called once per occupied element.
template <class Emit>
DEVI void walk_prediction(const uint16_t *__restrict__ g, const uint16_t *__restrict__ dm, int W,
                          int vr, int vc, int dir, int tr, int tc, int tpc, Emit emit) {
    int r = vr, c = vc, d = dir, k = 0;
    unsigned best_dist = FL_DIST_INF;
    int pr = r, pc = c, pd = d, ppd = d;          // pending (previous) element and the one before it
    bool have_prev = false;
    while (true) {
        // element k = (r, c, d) is known here; emit element k-1 now that its successor is known
        if (have_prev) {
            const int kk = k - 1;
            const int t0 = kk == 0 ? 0 : 1 + (kk - 1) * tpc;
            const int t1 = kk == 0 ? 0 : kk * tpc;
            if (t0 < NPRED) emit(pc * W + pr, t0, min(t1, NPRED - 1), pd, ppd, d);
            else return;
        }
        bool last = (r == tr && c == tc) || k >= FL_PRED_DEPTH;   // at target, or 500 greedy steps done
        int nr = r, nc = c, ndir = d;
        if (!last) {
            int md[3];
            const int n = greedy_moves(__ldg(g + r * W + c), d, md);
            int best = -1;
            for (int j = 0; j < n; j++) {
                const int rr = r + d_row(md[j]), cc = c + d_col(md[j]);
                const unsigned v = __ldg(dm + ((size_t)(rr * W + cc)) * 4 + md[j]);
                if (v < best_dist) { best = j; best_dist = v; nr = rr; nc = cc; ndir = md[j]; }
            }
            if (best < 0) last = true;             // rail disconnected: path ends here
        }
        if (last) {
            const int t0 = k == 0 ? 0 : 1 + (k - 1) * tpc;
            if (t0 < NPRED) emit(c * W + r, t0, NPRED - 1, d, pd, d);
            return;
        }
        ppd = pd; pr = r; pc = c; pd = d; have_prev = true;
        r = nr; c = nc; d = ndir; k++;
    }
}

DEVI uint64_t pack_entry(int agent, int t0, int t1, int dh, int dp, int dn, int done) {
    return (uint64_t)agent | ((uint64_t)t0 << 10) | ((uint64_t)t1 << 19) | ((uint64_t)dh << 28) |
           ((uint64_t)dp << 30) | ((uint64_t)dn << 32) | ((uint64_t)done << 34);
}

// loader.cpp:273-312: valid-action mask
DEVI void valid_actions_of(const uint16_t *__restrict__ g, int W, int st, int ctr, int r, int c, int d, uint8_t va[5]) {
    va[0] = va[1] = va[2] = va[3] = va[4] = 0;
    if (st == MOVING || st == STOPPED) {
        if (ctr == 0) {
            const unsigned cell = __ldg(g + r * W + c);
            const int nb = nibble(cell, d);
            int cnt = 0;
            bool branch_next = false;
            for (int a = A_LEFT; a <= A_RIGHT; a++) {
                const int nd = (d + a - 2) & 3;
                va[a] = tbit(nb, nd);
                if (va[a]) {
                    cnt++;
                    if (__popc(__ldg(g + (r + d_row(nd)) * W + c + d_col(nd))) > 2) branch_next = true;
                }
            }
            if (__popc(cell) > 2 || (cnt == 1 && branch_next)) va[A_STOP] = 1;
        } else va[A_NOTHING] = 1;
    } else if (st == READY) { va[A_FORWARD] = 1; va[A_STOP] = 1; }
    else va[A_NOTHING] = 1;
}

// Serial, order-dependent and sticky: an exact restatement of DeadlockChecker::update_deadlocks /
// _check_blocked / _fix_deps (deadlock_checker.cpp:11-110) with the recursion turned into an explicit
// stack.  Runs on one spare thread per environment while the other threads walk the predictions.
struct DeadlockScratch {
    uint8_t *checked, *ndep, *dl, *ct, *stk_d, *stk_phase;
    uint16_t *dep, *stk_h, *stk_opp;
    const int *cellid;
};

DEVI void update_deadlocks(const DeadlockScratch &x, const uint32_t *ci, int N, int H, int W) {
    for (int a0 = 0; a0 < N; a0++) {
        if (x.cellid[a0] < 0 || x.dl[a0] || x.checked[a0]) continue;
        int sp = 0;
        x.stk_h[0] = a0; x.stk_d[0] = 0; x.stk_phase[0] = 0; x.checked[a0] = 1; sp = 1;
        while (sp > 0) {
            const int f = sp - 1, h = x.stk_h[f];
            const int hr = x.cellid[h] / W, hc = x.cellid[h] % W;
            bool popped = false, pushed = false;
            while (x.stk_d[f] < 4) {
                const int dd = x.stk_d[f];
                int opp;
                if (x.stk_phase[f] == 1) { opp = x.stk_opp[f]; x.stk_phase[f] = 0; }
                else {
                    if (!tbit(x.ct[h], dd)) { x.stk_d[f]++; continue; }
                    const int rr = hr + d_row(dd), cc = hc + d_col(dd);
                    opp = -1;
                    if (rr >= 0 && cc >= 0 && rr < H && cc < W) {
                        opp = (int)(ci[rr * W + cc] >> 21) - 1;
                    }
                    if (opp < 0) { x.checked[h] = 2; popped = true; break; }           // road is free
                    if (x.dl[opp]) { x.stk_d[f]++; continue; }                          // road is blocked
                    if (x.checked[opp] == 0) {                                          // recurse
                        x.stk_phase[f] = 1; x.stk_opp[f] = (uint16_t)opp;
                        x.stk_h[sp] = (uint16_t)opp; x.stk_d[sp] = 0; x.stk_phase[sp] = 0; x.checked[opp] = 1; sp++;
                        pushed = true;
                        break;
                    }
                }
                if (x.checked[opp] == 2 && !x.dl[opp]) { x.checked[h] = 2; popped = true; break; }  // may become free
                x.dep[h * 4 + x.ndep[h]] = (uint16_t)opp; x.ndep[h]++;
                x.stk_d[f]++;
            }
            if (pushed) continue;
            if (!popped && x.ndep[h] == 0) {
                x.checked[h] = 2;
                if (x.ct[h] != 0) x.dl[h] = 1;
            }
            sp--;
        }
    }
    bool any = true;                                                                    // _fix_deps
    while (any) {
        any = false;
        for (int h = 0; h < N; h++) {
            if (x.checked[h] != 1) continue;
            int cnt = 0;
            for (int k = 0; k < x.ndep[h]; k++) {
                const int o = x.dep[h * 4 + k];
                if (x.checked[o] == 2) {
                    if (x.dl[o]) cnt++;
                    else { x.checked[h] = 2; any = true; }
                }
            }
            if (cnt == x.ndep[h]) { x.checked[h] = 2; x.dl[h] = 1; any = true; }
        }
    }
    for (int h = 0; h < N; h++) if (x.checked[h] == 1) { x.dl[h] = 1; x.checked[h] = 2; }
}

__global__ void __launch_bounds__(1024)
k_prep(FlBatch b, uint8_t *__restrict__ valid_actions, float *__restrict__ dist_target) {
    const int e = blockIdx.x, N = (int)b.N, H = (int)b.H, W = (int)b.W, HW = H * W;
    const int K = W * W + H;                       // key space of the reference's cell id c*W + r
    const int i = threadIdx.x, nt = blockDim.x;
    const bool act = i < N;
    extern __shared__ __align__(16) unsigned char smraw[];
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smraw);             // [K+1]
    uint32_t *s_part = s_cnt + (K + 1);                                // [nt] scan partials
    int *s_cellid = reinterpret_cast<int *>(s_part + nt);              // [N]
    uint16_t *s_dep = reinterpret_cast<uint16_t *>(s_cellid + N);      // [4N]
    uint16_t *s_stk_h = s_dep + 4 * N, *s_stk_opp = s_stk_h + N;       // [N] each
    uint8_t *s_checked = reinterpret_cast<uint8_t *>(s_stk_opp + N);   // [N] x6
    uint8_t *s_ndep = s_checked + N, *s_dl = s_ndep + N, *s_ct = s_dl + N, *s_stk_d = s_ct + N,
            *s_stk_phase = s_stk_d + N;
    int *s_initcell = reinterpret_cast<int *>(s_stk_phase + N + ((4 - (6 * N) % 4) % 4));  // [N]
    const uint16_t *__restrict__ g = b.grid + (size_t)e * HW;
    uint32_t *ci = b.cellinfo + (size_t)e * HW;
    const size_t ea = (size_t)e * N + (act ? i : 0);

    for (int k = i; k <= K; k += nt) s_cnt[k] = 0;
    if (act) { const int oc = b.occ_cell[ea]; if (oc >= 0) ci[oc] = 0; }   // forget the previous step's occupancy
    int r = -1, c = -1, d = 0, st = DONE, vr = 0, vc = 0, tr = 0, tc = 0, tpc = 1;
    const uint16_t *dm = nullptr;
    if (act) {
        const short2 p = reinterpret_cast<const short2 *>(b.rc)[ea];
        r = p.x; c = p.y; d = b.dir[ea]; st = b.state[ea];
        const short2 ip = reinterpret_cast<const short2 *>(b.init_rc)[ea];
        const short2 tp = reinterpret_cast<const short2 *>(b.tgt_rc)[ea];
        tr = tp.x; tc = tp.y;
        if (off_map(st)) { vr = ip.x; vc = ip.y; } else if (on_map(st)) { vr = r; vc = c; } else { vr = tr; vc = tc; }
        dm = b.dist + ((size_t)e * b.n_slots + b.slot[ea]) * HW * 4;
        tpc = (int)(1.0f / (float)b.speed[ea]);                        // predictions.cpp:184
        s_cellid[i] = on_map(st) ? r * W + c : -1;
        s_initcell[i] = off_map(st) ? ip.x * W + ip.y : -1;
        s_ct[i] = on_map(st) ? (uint8_t)nibble(__ldg(g + r * W + c), d) : 0;
        s_dl[i] = b.deadlocked[ea]; s_checked[i] = 0; s_ndep[i] = 0;
        uint8_t va[5];
        valid_actions_of(g, W, st, b.ctr[ea], r, c, d, va);
        for (int k = 0; k < 5; k++) valid_actions[ea * 5 + k] = va[k];
        float dt;                                                       // loader.cpp:163-179
        if (st == DONE) dt = 0.0f;
        else {
            const unsigned dv = off_map(st) ? __ldg(dm + ((size_t)(ip.x * W + ip.y)) * 4 + b.init_dir[ea])
                                            : __ldg(dm + ((size_t)(r * W + c)) * 4 + d);
            dt = dv == FL_DIST_INF ? INFINITY : (float)dv;
        }
        dist_target[ea] = dt;
    }
    __syncthreads();
    // occupancy map (treeobs.cpp:67-92, deadlock_checker.cpp:15-20): per cell the HIGHEST handle standing on
    // it (std::map assignment in handle order = last writer), its direction and malfunction flag, and the
    // number of off-map trains whose initial cell it is.  handle+1 sits in the top bits so atomicMax picks it.
    if (act) {
        int cellid = s_cellid[i];
        if (cellid >= 0) {
            int cnt = 0;
            for (int k = 0; k < N; k++) cnt += s_initcell[k] == cellid;
            atomicMax(&ci[cellid], ((uint32_t)(i + 1) << 21) | ((uint32_t)cnt << 11) | ((uint32_t)d << 9) |
                                       ((uint32_t)(b.mal[ea] != 0) << 8));
        }
        b.occ_cell[ea] = cellid;
    }
    __syncthreads();
    if (i == N) {                                  // the spare thread (block has at least N+1 threads)
        DeadlockScratch x{s_checked, s_ndep, s_dl, s_ct, s_stk_d, s_stk_phase, s_dep, s_stk_h, s_stk_opp, s_cellid};
        update_deadlocks(x, ci, N, H, W);
    }
    const int done_flag = st == DONE;
    if (act)                                       // pass 1: count entries per cell id
        walk_prediction(g, dm, W, vr, vc, d, tr, tc, tpc,
                        [&](int key, int, int, int, int, int) { atomicAdd(&s_cnt[key], 1u); });
    __syncthreads();
    if (act) b.deadlocked[ea] = s_dl[i];
    // exclusive scan of s_cnt[0..K] (K+1 values; the last becomes the total)
    const int per = (K + 1 + nt - 1) / nt, lo = min(i * per, K + 1), hi = min(lo + per, K + 1);
    uint32_t sum = 0;
    for (int k = lo; k < hi; k++) sum += s_cnt[k];
    s_part[i] = sum;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {       // Hillis-Steele inclusive scan of the partials
        const uint32_t v = i >= off ? s_part[i - off] : 0;
        __syncthreads();
        s_part[i] += v;
        __syncthreads();
    }
    uint32_t run = s_part[i] - sum;
    uint32_t *ks = b.key_start + (size_t)e * (K + 1);
    for (int k = lo; k < hi; k++) { const uint32_t v = s_cnt[k]; s_cnt[k] = run; ks[k] = run; run += v; }
    __syncthreads();
    if (act) {                                     // pass 2: scatter entries
        uint64_t *ent = b.entries + (size_t)e * b.ent_cap;
        walk_prediction(g, dm, W, vr, vc, d, tr, tc, tpc,
                        [&](int key, int t0, int t1, int dh, int dp, int dn) {
                            const uint32_t pos = atomicAdd(&s_cnt[key], 1u);
                            ent[pos] = pack_entry(i, t0, t1, dh, dp, dn, done_flag);
                        });
    }
}

// ---------------------------------------------------------------------------------------------
// k_tree: one warp per agent
// ---------------------------------------------------------------------------------------------
// rotate_transition (tool.h:300-335): rotate the 4 bits inside every orientation block right by k,
// then rotate the four blocks right by k
DEVI int rotate_transition(int t, int k) {
    int v = 0;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        int bl = (t >> ((3 - o) * 4)) & 0xF;
        bl = ((bl >> k) | (bl << (4 - k))) & 0xF;
        v |= bl << ((3 - o) * 4);
    }
    return (((v & ((1 << (k * 4)) - 1)) << ((4 - k) * 4)) | (v >> (k * 4))) & 0xFFFF;
}

__constant__ int c_road_types[11] = {0x0000, 0x8020, 0x9220, 0x8421, 0x9621, 0xCC33,
                                     0x5202, 0x2000, 0x4002, 0x1200, 0xC022};  // loader.cpp:123-134

DEVI float scale_dist(float v, float T) { return v != INFINITY ? v / T : -1.0f; }  // treeobs.cpp:111-152

struct alignas(16) TreeSmem {  // per warp
    float forest[FL_MAX_NODES * FL_NODE_F];
    int adj[(FL_MAX_NODES - 1) * 3];
    int norder[FL_MAX_NODES];
    int eorder[FL_MAX_NODES - 1];
    int q_rc[FL_MAX_NODES - 1];    // packed (r, c)
    int q_meta[FL_MAX_NODES - 1];  // dir | (ad+1)<<2 | null<<4 | parent<<8
    int q_tot[FL_MAX_NODES - 1];
};

__global__ void __launch_bounds__(TREE_WARPS * 32)
k_tree(FlBatch b, const uint8_t *__restrict__ valid_actions, const float *__restrict__ dist_target,
       float *__restrict__ out_attr, float *__restrict__ out_forest, int32_t *__restrict__ out_adj,
       int32_t *__restrict__ out_norder, int32_t *__restrict__ out_eorder, int bitmap_words) {
    const int N = (int)b.N, H = (int)b.H, W = (int)b.W, HW = H * W, K = W * W + H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ga = (long long)blockIdx.x * TREE_WARPS + warp;
    extern __shared__ __align__(16) unsigned char smraw[];
    TreeSmem *ts = reinterpret_cast<TreeSmem *>(smraw) + warp;
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(smraw + sizeof(TreeSmem) * TREE_WARPS) + (size_t)warp * bitmap_words;
    if (ga >= (long long)b.E * N) return;
    const int e = (int)(ga / N), h = (int)(ga % N);
    const size_t ea = (size_t)ga;
    const uint16_t *__restrict__ g = b.grid + (size_t)e * HW;
    const uint32_t *__restrict__ ci = b.cellinfo + (size_t)e * HW;
    const uint32_t *__restrict__ ks = b.key_start + (size_t)e * (K + 1);
    const uint64_t *__restrict__ ent = b.entries + (size_t)e * b.ent_cap;
    const uint16_t *__restrict__ dm = b.dist + ((size_t)e * b.n_slots + b.slot[ea]) * HW * 4;

    // ---- agent view (loader.cpp:8-179) -------------------------------------------------------
    const short2 p = reinterpret_cast<const short2 *>(b.rc)[ea];
    const short2 ip = reinterpret_cast<const short2 *>(b.init_rc)[ea];
    const short2 tp = reinterpret_cast<const short2 *>(b.tgt_rc)[ea];
    const int r0 = p.x, c0 = p.y, dir = b.dir[ea], st = b.state[ea], tr = tp.x, tc = tp.y;
    int vr, vc;
    if (off_map(st)) { vr = ip.x; vc = ip.y; } else if (on_map(st)) { vr = r0; vc = c0; } else { vr = tr; vc = tc; }
    const float speed = (float)b.speed[ea];
    const float T = (float)b.max_steps[e], Nf = (float)N;
    const int nmal01 = b.nmal[ea] != 0, mal01 = b.mal[ea] != 0;
    const float dtgt = dist_target[ea];
    const float tpc_f = (float)(1.0 / (double)speed);                   // treeobs.cpp:304

    for (int k = lane; k < bitmap_words; k += 32) bitmap[k] = 0;

    // ---- root node (treeobs.cpp:171-190) -----------------------------------------------------
    if (lane < FL_NODE_F) {
        float v = 0.0f;
        if (lane == 6) v = scale_dist(dtgt, T);
        else if (lane == 9) v = (float)nmal01 / Nf;
        else if (lane == 10) v = speed;
        ts->forest[lane] = v;
    }
    int qt = 0, qh = 0;                            // pushes so far, pops so far (only 30 pops ever happen)
    {
        const int nb = nibble(__ldg(g + vr * W + vc), dir);
        int orientation = dir;
        if (__popc(nb) == 1) orientation = first_dir(nb);
        for (int ad = -1; ad <= 1; ad++) {
            const int bd = (orientation + ad) & 3;
            const bool real = tbit(nb, bd);
            if (lane == 0) {
                ts->q_rc[qt] = real ? (((vr + d_row(bd)) & 0xFFFF) | ((vc + d_col(bd)) << 16)) : -1;
                ts->q_meta[qt] = bd | ((ad + 1) << 2) | ((real ? 0 : 1) << 4) | (0 << 8);
                ts->q_tot[qt] = 1;
            }
            qt++;
        }
    }
    __syncwarp();

    // ---- 30 nodes in FIFO order (treeobs.cpp:223-250, _explore_branch :258-610) ---------------
    for (int n = 1; n < FL_MAX_NODES; n++) {
        float *node = ts->forest + n * FL_NODE_F;
        if (qh == qt) {                            // queue empty: padding row
            if (lane < FL_NODE_F) node[lane] = -1.0f;
            if (lane < 3) ts->adj[(n - 1) * 3 + lane] = -2;
            continue;
        }
        const int qrc = ts->q_rc[qh], qm = ts->q_meta[qh];
        float tot = (float)ts->q_tot[qh];
        qh++;
        const int parent = qm >> 8, ad = ((qm >> 2) & 3) - 1;
        if (lane == 0) { ts->adj[(n - 1) * 3] = parent; ts->adj[(n - 1) * 3 + 1] = n; ts->adj[(n - 1) * 3 + 2] = ad; }
        if ((qm >> 4) & 1) {                       // null child: a valid leaf with all features -1
            if (lane < FL_NODE_F) node[lane] = -1.0f;
            continue;
        }
        int r = (int)(short)(qrc & 0xFFFF), c = qrc >> 16, d = qm & 3;
        float own = INFINITY, other_agent = INFINITY, conflict = INFINITY, unusable = INFINITY, min_speed = 1.0f;
        int same = 0, opp = 0, malf = 0, rtdn = 0;
        int kind = 0;                              // 1 switch, 2 dead end, 3 terminal (cycle), 4 target
        while (true) {
            const int cell = r * W + c;
            const uint32_t cinfo = ci[cell];
            const unsigned gc = __ldg(g + cell);
            if (cinfo) {                           // treeobs.cpp:322-360 (the observer itself counts too)
                if (tot < other_agent) other_agent = tot;
                malf = max(malf, (int)((cinfo >> 8) & 1u));
                const int cnt = (int)((cinfo >> 11) & 1023u);
                rtdn += cnt ? cnt - 1 : 0;
                if ((int)((cinfo >> 9) & 3u) == d) {
                    same++;
                    min_speed = fminf(min_speed, (float)b.speed[(size_t)e * N + (cinfo >> 21) - 1]);
                } else opp++;
            }
            const int nb = nibble(gc, d);
            int total = __popc(gc);
            const bool crossing = gc == 0x8421u;
            const int pt = (int)__fmul_rn((float)(int)tot, tpc_f);     // treeobs.cpp:378
            if (pt < NPRED && tot < (float)NPRED) {                     // treeobs.cpp:379-465
                const int key = c * W + r;
                const uint32_t s0 = ks[key], s1 = ks[key + 1];
                const int pre = max(0, pt - 1), post = min(NPRED - 1, pt + 1);
                unsigned acc = 0;
                for (uint32_t base = s0; base < s1; base += 32) {
                    const uint32_t idx = base + lane;
                    unsigned f = 0;
                    if (idx < s1) {
                        const uint64_t en = ent[idx];
                        const int ag = (int)(en & 1023), t0 = (int)((en >> 10) & 511), t1 = (int)((en >> 19) & 511);
                        const int dh = (int)((en >> 28) & 3), dp = (int)((en >> 30) & 3), dn = (int)((en >> 32) & 3);
                        const bool done = (en >> 34) & 1;
                        const bool in_cur = t0 <= pt && pt <= t1, in_pre = t0 <= pre && pre <= t1,
                                   in_post = t0 <= post && post <= t1;
                        const int pdir = pt < t0 ? dp : (pt > t1 ? dn : dh);  // always the direction at row pt
                        const bool cf = (d != pdir && tbit(nb, (pdir + 2) & 3)) || done;
                        const bool other = ag != h;
                        f = (in_cur && other ? 1u : 0u) | (in_pre && other ? 2u : 0u) | (in_post && other ? 4u : 0u) |
                            (in_cur && cf ? 8u : 0u) | (in_pre && cf ? 16u : 0u) | (in_post && cf ? 32u : 0u);
                    }
                    acc |= __reduce_or_sync(0xFFFFFFFFu, f);
                }
                const bool cf = (acc & 1u) ? (acc & 8u) : (acc & 2u) ? (acc & 16u) : (acc & 4u) ? (acc & 32u) : false;
                if (cf && tot < conflict) conflict = tot;
            }
            const bool is_target = r == tr && c == tc;
            if (is_target && tot < own) own = tot;
            const int bit = cell * 4 + d;          // per-branch visited set (treeobs.cpp:476-481)
            const uint32_t word = bitmap[bit >> 5];
            __syncwarp();                          // every lane has read before any lane stores
            if (word & (1u << (bit & 31))) { kind = 3; break; }
            bitmap[bit >> 5] = word | (1u << (bit & 31));  // every lane stores the same value
            if (is_target) { kind = 4; break; }
            if (crossing) total = 2;
            const int num = __popc(nb);
            if (total > 2 && num < 2 && tot < unusable) unusable = tot;
            if (num == 1) {
                if (total == 1) { kind = 2; break; }
                d = first_dir(nb); r += d_row(d); c += d_col(d); tot += 1.0f;
            } else if (num > 1) { kind = 1; break; }
            else {                                  // treeobs.cpp:527-535 throws; report and stop here
                if (lane == 0) atomicOr(&b.status[e], FL_ST_BAD_CELL);
                kind = 3; break;
            }
        }
        __syncwarp();
        for (int k = lane; k < bitmap_words; k += 32) bitmap[k] = 0;
        float dnb, dmin;
        if (kind == 4) { dnb = tot; dmin = 0.0f; }
        else {
            const unsigned dv = __ldg(dm + ((size_t)(r * W + c)) * 4 + d);
            dmin = dv == FL_DIST_INF ? INFINITY : (float)dv;
            dnb = kind == 3 ? INFINITY : tot;
        }
        if (lane < FL_NODE_F) {                    // scale_node (treeobs.cpp:111-152)
            float v;
            switch (lane) {
            case 0: v = scale_dist(own, T); break;
            case 1: v = -1.0f; break;              // location_has_target is never filled (treeobs.cpp:72)
            case 2: v = scale_dist(other_agent, T); break;
            case 3: v = scale_dist(conflict, T); break;
            case 4: v = scale_dist(unusable, T); break;
            case 5: v = scale_dist(dnb, T); break;
            case 6: v = scale_dist(dmin, T); break;
            case 7: v = (float)same / Nf; break;
            case 8: v = (float)opp / Nf; break;
            case 9: v = (float)malf / Nf; break;
            case 10: v = min_speed; break;
            default: v = (float)rtdn / Nf; break;
            }
            node[lane] = v;
        }
        const int nb2 = nibble(__ldg(g + r * W + c), d);
        for (int a2 = -1; a2 <= 1; a2++) {          // children in order L, F, R (treeobs.cpp:583-608)
            const int bd = (d + a2) & 3, rb = (bd + 2) & 3;
            int cd = bd, real = 0;
            if (kind == 2 && tbit(nb2, rb)) { cd = rb; real = 1; }
            else if (kind == 1 && tbit(nb2, bd)) { cd = bd; real = 1; }
            if (qt < FL_MAX_NODES - 1 && lane == 0) {
                ts->q_rc[qt] = real ? (((r + d_row(cd)) & 0xFFFF) | ((c + d_col(cd)) << 16)) : -1;
                ts->q_meta[qt] = cd | ((a2 + 1) << 2) | ((real ? 0 : 1) << 4) | (n << 8);
                ts->q_tot[qt] = (int)tot + 1;
            }
            qt++;
        }
        __syncwarp();
    }
    __syncwarp();

    // ---- evaluation orders (tool.h:468-524): node_order = height above the leaves ------------
    if (lane == 0) {
        int count = 1;
        for (int k = 0; k < FL_MAX_NODES - 1; k++) if (ts->adj[k * 3] != -2) count++;
        for (int k = 0; k < FL_MAX_NODES; k++) ts->norder[k] = k < count ? 0 : -2;
        for (int k = FL_MAX_NODES - 2; k >= 0; k--) {
            const int pa = ts->adj[k * 3], ch = ts->adj[k * 3 + 1];
            if (pa >= 0) ts->norder[pa] = max(ts->norder[pa], ts->norder[ch] + 1);
        }
        for (int k = 0; k < FL_MAX_NODES - 1; k++) {
            const int pa = ts->adj[k * 3];
            ts->eorder[k] = pa < 0 ? -2 : ts->norder[pa];
        }
    }
    __syncwarp();

    // ---- coalesced vector stores in the policy layout ----------------------------------------
    {
        float4 *dst = reinterpret_cast<float4 *>(out_forest + ea * (FL_MAX_NODES * FL_NODE_F));
        const float4 *src = reinterpret_cast<const float4 *>(ts->forest);
        for (int k = lane; k < FL_MAX_NODES * FL_NODE_F / 4; k += 32) dst[k] = src[k];
        int2 *da = reinterpret_cast<int2 *>(out_adj + ea * ((FL_MAX_NODES - 1) * 3));
        const int2 *sa = reinterpret_cast<const int2 *>(ts->adj);
        for (int k = lane; k < (FL_MAX_NODES - 1) * 3 / 2; k += 32) da[k] = sa[k];
        if (lane < FL_MAX_NODES) out_norder[ea * FL_MAX_NODES + lane] = ts->norder[lane];
        if (lane < FL_MAX_NODES - 1) out_eorder[ea * (FL_MAX_NODES - 1) + lane] = ts->eorder[lane];
    }

    // ---- agent attributes (feature_parser.cpp:3-98) ------------------------------------------
    {
        const int trans = r0 < 0 ? 0 : (int)__ldg(g + r0 * W + c0);
        int road = 0;                               // loader.cpp:122-161: first rotation that is in the table
        if (r0 >= 0) {
            int rot[4] = {trans, rotate_transition(trans, 1), rotate_transition(trans, 2), rotate_transition(trans, 3)};
            bool found = false;
            for (int q = 0; q < 4 && !found; q++)
                for (int k = 0; k < 11; k++) if (c_road_types[k] == rot[q]) { road = k; found = true; break; }
        }
        const int idir = b.init_dir[ea], od_raw = b.old_dir[ea], od = od_raw == 255 ? dir : od_raw;
        const int ctr = b.ctr[ea], maxc = b.max_count[ea], elapsed = b.elapsed[e];
        const float max_dist = (float)((H + W) * 8);
        const float agent_handle = (float)h / Nf;
        const float curr_step = (float)elapsed / T;
        const float earliest = (float)b.earliest[ea] / T, latest = (float)b.latest[ea] / T;
        const float arrival = (float)b.arrival[ea] / T;
        const float before_late = __fsub_rn(latest, curr_step);
        const float dist_f = dtgt == INFINITY ? 8.0f : dtgt / max_dist;
        const float anticipative = before_late < dist_f ? before_late : dist_f;
        const unsigned idv = __ldg(dm + ((size_t)(ip.x * W + ip.y)) * 4 + idir);
        const float idist = idv == FL_DIST_INF ? 8.0f : (float)idv / max_dist;
        float *dst = out_attr + ea * FL_ATTR_F;
        for (int k = lane; k < FL_ATTR_F; k += 32) {
            float v;
            if (k < 7) v = k == st;
            else if (k < 18) v = (k - 7) == road;
            else if (k < 28) v = (k - 18) == nmal01;
            else if (k < 32) v = (k - 28) == idir;
            else if (k < 36) v = (k - 32) == dir;
            else if (k < 40) v = (k - 36) == od;
            else if (k < 49) {
                switch (k - 40) {
                case 0: v = st == MOVING; break;
                case 1: v = b.deadlocked[ea] != 0; break;
                case 2: v = b.sig_mal[ea] != 0; break;
                case 3: v = b.mal[ea] == 0; break;
                case 4: v = ctr == 0; break;
                case 5: v = ctr == maxc; break;
                case 6: v = st == MALFUNCTION || st == MAL_OFF; break;
                case 7: v = off_map(st); break;
                default: v = on_map(st); break;
                }
            } else if (k < 65) v = (trans >> (15 - (k - 49))) & 1;
            else if (k < 70) v = valid_actions[ea * 5 + (k - 65)];
            else {
                switch (k - 70) {
                case 0: v = agent_handle; break;
                case 1: v = curr_step; break;
                case 2: v = earliest; break;
                case 3: v = latest; break;
                case 4: v = arrival; break;
                case 5: v = before_late; break;
                case 6: v = dist_f; break;
                case 7: v = anticipative; break;
                case 8: v = (float)maxc / 10.0f; break;
                case 9: v = speed / 1.0f; break;
                case 10: v = (float)ctr / 10.0f; break;
                case 11: v = (float)mal01 / 10.0f; break;
                default: v = idist; break;
                }
            }
            dst[k] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_bfs: DistanceMap (distance_map.py:57-160) as a level-synchronous pull BFS over (cell, orientation)
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(1024) k_bfs(FlBatch b) {
    const int H = (int)b.H, W = (int)b.W, HW = H * W, ns = (int)b.n_slots;
    const int e = blockIdx.x / ns, s = blockIdx.x % ns;
    extern __shared__ __align__(16) unsigned char smraw[];
    uint16_t *out = b.dist + ((size_t)e * ns + s) * HW * 4;
    uint16_t *dd = SMEM ? reinterpret_cast<uint16_t *>(smraw) : out;
    const uint16_t *__restrict__ g = b.grid + (size_t)e * HW;
    const int tr = b.slot_rc[((size_t)e * ns + s) * 2], tc = b.slot_rc[((size_t)e * ns + s) * 2 + 1];
    for (int k = threadIdx.x; k < HW * 4; k += blockDim.x) dd[k] = FL_DIST_INF;
    __syncthreads();
    if (tr >= 0) {
        if (threadIdx.x < 4) dd[(tr * W + tc) * 4 + threadIdx.x] = 0;
        __syncthreads();
        for (int level = 1; level < 0xFFFF; level++) {
            int changed = 0;
            for (int cell = threadIdx.x; cell < HW; cell += blockDim.x) {
                const unsigned gc = __ldg(g + cell);
                if (!gc) continue;
                const int r = cell / W, c = cell - r * W;
                for (int o = 0; o < 4; o++) {
                    if (dd[cell * 4 + o] != FL_DIST_INF) continue;
                    const int nb = nibble(gc, o);
                    for (int m = 0; m < 4; m++) {
                        if (!tbit(nb, m)) continue;
                        const int rr = r + d_row(m), cc = c + d_col(m);
                        if (rr < 0 || cc < 0 || rr >= H || cc >= W) continue;
                        if (dd[(rr * W + cc) * 4 + m] == level - 1) { dd[cell * 4 + o] = (uint16_t)level; changed = 1; break; }
                    }
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k < HW * 4; k += blockDim.x) out[k] = dd[k];
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

int check_batch(const FlBatch *b) {
    if (!b || b->E <= 0 || b->N <= 0 || b->H <= 0 || b->W <= 0 || b->n_slots <= 0 || b->S <= 0) return FL_ERR_BAD_ARG;
    if (b->N >= FL_MAX_AGENTS) return FL_ERR_TOO_MANY_AGENTS;
    if (b->ent_cap < b->N * (int64_t)NPRED) return FL_ERR_BAD_ARG;
    if (b->H >= 32768 || b->W >= 32768 || b->H * b->W > (1 << 20)) return FL_ERR_BAD_ARG;
    return FL_OK;
}

size_t prep_smem_bytes(const FlBatch *b, int nt) {
    const size_t K = (size_t)(b->W * b->W + b->H), N = (size_t)b->N;
    return (K + 1) * 4 + (size_t)nt * 4 + N * 4 + 4 * N * 2 + 2 * N * 2 + 6 * N + 4 + N * 4 + 16;
}

int finish(cudaError_t launch_err) { return launch_err == cudaSuccess ? FL_OK : (int)launch_err; }

}  // namespace

extern "C" {

int fl_abi_version(void) { return FL_ABI_VERSION; }
size_t fl_batch_sizeof(void) { return sizeof(FlBatch); }
uint64_t fl_launch_count(void) { return g_launches.load(); }

int fl_profile_num_kernels(void) { return K_COUNT; }
const char *fl_profile_kernel_name(int k) { return k >= 0 && k < K_COUNT ? kKernelNames[k] : ""; }

void fl_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
}

int fl_profile_collect(double *ms_out, uint64_t *launches_out, int reset) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (const ProfRec &r : g_prof_pending) {
        cudaError_t err = cudaEventSynchronize(r.b);
        if (err != cudaSuccess) return (int)err;
        float ms = 0.0f;
        err = cudaEventElapsedTime(&ms, r.a, r.b);
        if (err != cudaSuccess) return (int)err;
        g_prof_ms[r.id] += ms; g_prof_n[r.id] += 1;
        g_prof_free.push_back(r.a); g_prof_free.push_back(r.b);
    }
    g_prof_pending.clear();
    for (int k = 0; k < K_COUNT; k++) {
        if (ms_out) ms_out[k] = g_prof_ms[k];
        if (launches_out) launches_out[k] = g_prof_n[k];
        if (reset) { g_prof_ms[k] = 0.0; g_prof_n[k] = 0; }
    }
    return FL_OK;
}

const char *fl_error_string(int code) {
    switch (code) {
    case FL_OK: return "ok";
    case FL_ERR_BAD_ARG: return "flatland_b200: bad argument (null batch, non-positive dimension, ent_cap too small)";
    case FL_ERR_TOO_MANY_AGENTS: return "flatland_b200: more than FL_MAX_AGENTS-1 agents per environment";
    case FL_ERR_SMEM: return "flatland_b200: configuration needs more shared memory than one SM has";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "flatland_b200: unknown error";
    }
}

int fl_distance_map(const FlBatch *b, void *stream) {
    if (int rc = check_batch(b)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = (int)(b->H * b->W);
    const size_t smem = (size_t)HW * 4 * sizeof(uint16_t);
    const int nt = HW >= 4096 ? 1024 : 256;
    const unsigned grid = (unsigned)(b->E * b->n_slots);
    if (smem <= 227 * 1024) {
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(k_bfs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
        }
        LaunchScope ls(K_BFS, st);
        k_bfs<true><<<grid, nt, smem, st>>>(*b);
    } else {
        LaunchScope ls(K_BFS, st);
        k_bfs<false><<<grid, nt, 0, st>>>(*b);
    }
    return finish(cudaGetLastError());
}

int fl_reset(const FlBatch *b, const uint8_t *d_env_mask, void *stream) {
    if (int rc = check_batch(b)) return rc;
    {
        LaunchScope ls(K_RESET, (cudaStream_t)stream);
        k_reset<<<(unsigned)b->E, 128, 0, (cudaStream_t)stream>>>(*b, d_env_mask);
    }
    return finish(cudaGetLastError());
}

int fl_step(const FlBatch *b, const uint8_t *d_actions, int32_t *d_rewards, uint8_t *d_dones, uint32_t flags,
            void *stream) {
    if (int rc = check_batch(b)) return rc;
    if (!d_actions || !d_rewards || !d_dones) return FL_ERR_BAD_ARG;
    const int nt = round_up((int)b->N, 32);
    {
        LaunchScope ls(K_STEP, (cudaStream_t)stream);
        k_step<<<(unsigned)b->E, nt, 4 * b->N * sizeof(int), (cudaStream_t)stream>>>(*b, d_actions, d_rewards, d_dones, flags);
    }
    return finish(cudaGetLastError());
}

int fl_observe(const FlBatch *b, float *d_agent_attr, float *d_forest, int32_t *d_adjacency, int32_t *d_node_order,
               int32_t *d_edge_order, uint8_t *d_valid_actions, float *d_dist_target, void *stream) {
    if (int rc = check_batch(b)) return rc;
    if (!d_agent_attr || !d_forest || !d_adjacency || !d_node_order || !d_edge_order || !d_valid_actions || !d_dist_target)
        return FL_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int nt = round_up((int)b->N + 1, 32);   // one spare thread runs the deadlock checker
        const size_t smem = prep_smem_bytes(b, nt);
        if (smem > 227 * 1024) return FL_ERR_SMEM;
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
        }
        {
            LaunchScope ls(K_PREP, st);
            k_prep<<<(unsigned)b->E, nt, smem, st>>>(*b, d_valid_actions, d_dist_target);
        }
        if (cudaError_t err = cudaGetLastError()) return (int)err;
    }
    {
        const int words = (int)((b->H * b->W * 4 + 31) / 32);
        const size_t smem = sizeof(TreeSmem) * TREE_WARPS + (size_t)words * 4 * TREE_WARPS;
        if (smem > 227 * 1024) return FL_ERR_SMEM;
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(k_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
        }
        const long long agents = b->E * b->N;
        const unsigned grid = (unsigned)((agents + TREE_WARPS - 1) / TREE_WARPS);
        LaunchScope ls(K_TREE, st);
        k_tree<<<grid, TREE_WARPS * 32, smem, st>>>(*b, d_valid_actions, d_dist_target, d_agent_attr, d_forest,
                                                     d_adjacency, d_node_order, d_edge_order, words);
    }
    return finish(cudaGetLastError());
}

int fl_step_observe_host(const FlBatch *b, const uint8_t *h_actions, uint8_t *d_actions, const FlObsBuffers *d_out,
                         const FlObsBuffers *h_out, uint32_t flags, void *stream) {
    if (int rc = check_batch(b)) return rc;
    if (!h_actions || !d_actions || !d_out || !h_out) return FL_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t EN = (size_t)(b->E * b->N);
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, EN, cudaMemcpyHostToDevice, st);
    if (err != cudaSuccess) return (int)err;
    if (int rc = fl_step(b, d_actions, d_out->rewards, d_out->dones, flags, stream)) return rc;
    if (int rc = fl_observe(b, d_out->agent_attr, d_out->forest, d_out->adjacency, d_out->node_order, d_out->edge_order,
                            d_out->valid_actions, d_out->dist_target, stream))
        return rc;
#define FL_D2H(field, bytes)                                                                              \
    if (h_out->field) {                                                                                   \
        err = cudaMemcpyAsync(h_out->field, d_out->field, (bytes), cudaMemcpyDeviceToHost, st);           \
        if (err != cudaSuccess) return (int)err;                                                          \
    }
    FL_D2H(rewards, EN * 4)
    FL_D2H(dones, (size_t)b->E * (b->N + 1))
    FL_D2H(agent_attr, EN * FL_ATTR_F * 4)
    FL_D2H(forest, EN * FL_MAX_NODES * FL_NODE_F * 4)
    FL_D2H(adjacency, EN * (FL_MAX_NODES - 1) * 3 * 4)
    FL_D2H(node_order, EN * FL_MAX_NODES * 4)
    FL_D2H(edge_order, EN * (FL_MAX_NODES - 1) * 4)
    FL_D2H(valid_actions, EN * 5)
    FL_D2H(dist_target, EN * 4)
#undef FL_D2H
    return FL_OK;
}

}  // extern "C"
