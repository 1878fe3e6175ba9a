// step.cuh — k_reset and k_step: RailEnv.reset tail and RailEnv.step (rail_env.py:335-347, 501-632).
#pragma once
#include "common.cuh"

namespace {

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
// treeobs.cpp:22-28).  Called by all threads of a CTA.  keep_arrival: EnvAgent.reset does not touch arrival_time, so
// after reset(False, False) a train that arrived in the previous episode still carries its old arrival_time (and
// handle_done_state, rail_env.py:493-499, then leaves it standing on its target cell when it arrives again); fresh
// agent objects (a regenerating reset, an upload) start with arrival_time None.
DEVI void reset_env(const FlBatch &b, int e, bool rewind_schedule, bool keep_arrival) {
    const int N = (int)b.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const size_t ea = (size_t)e * N + i;
        b.rc[2 * ea] = -1; b.rc[2 * ea + 1] = -1;
        b.old_rc[2 * ea] = -1; b.old_rc[2 * ea + 1] = -1;
        b.dir[ea] = b.init_dir[ea]; b.old_dir[ea] = 255;
        b.state[ea] = WAITING; b.ctr[ea] = 0; b.mal[ea] = 0; b.saved[ea] = 0; b.sig_mal[ea] = 0;
        b.deadlocked[ea] = 0; b.done[ea] = 0; b.nmal[ea] = 0;
        if (!keep_arrival) b.arrival[ea] = -1;
    }
    if (threadIdx.x == 0) {
        b.elapsed[e] = 0; b.done_all[e] = 0;
        if (rewind_schedule) { b.sched_pos[e] = 0; b.status[e] = 0; }
    }
}

__global__ void k_reset(FlBatch b, const uint8_t *__restrict__ mask, uint32_t flags) {
    const int e = blockIdx.x;
    if (mask && !mask[e]) return;
    reset_env(b, e, !(flags & FL_RESET_KEEP_SCHEDULE), (flags & FL_RESET_KEEP_ARRIVAL) != 0);
    if (b.tree_cache)                               // a reset may follow an upload: drop the cached tree structures
        for (int i = threadIdx.x; i < (int)b.N; i += blockDim.x) b.tree_cache[((size_t)e * b.N + i) * FL_TREE_CACHE_WORDS + 31] = 0u;
    if (b.path_cache)
        for (int i = threadIdx.x; i < (int)b.N; i += blockDim.x) b.path_cache[((size_t)e * b.N + i) * b.pc_stride] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// k_step: RailEnv.step (rail_env.py:501-632)
// ---------------------------------------------------------------------------------------------
// map_cells > 0: MotionCheck looks its neighbours up in per-rail-cell tables (O(N)) instead of comparing every pair of agents
// (O(N^2): 40 us at 425 agents); map_cells = rail-cell capacity of the tables, which follow the four per-agent arrays.
__global__ void __launch_bounds__(1024)
k_step(FlBatch b, const uint8_t *__restrict__ actions, int32_t *__restrict__ rewards,
       uint8_t *__restrict__ dones, uint32_t flags, int map_cells) {
    const int e = blockIdx.x, N = (int)b.N, H = (int)b.H, W = (int)b.W, HW = H * W;
    const int i = threadIdx.x;
    const bool act = i < N;
    extern __shared__ int sm[];
    int *s_cur = sm, *s_nxt = sm + N, *s_rep = sm + 2 * N, *s_blk = sm + 3 * N;
    int *m_rep = sm + 4 * N, *m_want = m_rep + map_cells, *m_out = m_want + map_cells;   // per rail cell, see MotionCheck below
    const uint16_t *__restrict__ ridx = b.ridx + (size_t)e * b.ridx_stride;
    const int R_env = map_cells > 0 ? (int)(b.walk_total[(size_t)e * 4] >> 2) : 0;       // 0: the rail index is not built
    if (R_env > 0)
        for (int k = i; k < R_env; k += blockDim.x) { m_rep[k] = 0; m_want[k] = 0x7fffffff; m_out[k] = 0; }
    const size_t ea = (size_t)e * N + (act ? i : 0);
    const uint16_t *__restrict__ g = b.grid + (size_t)e * b.grid_stride;

    const bool was_done = b.done_all[e] != 0;
    const int elapsed = b.elapsed[e] + 1;
    const int srow = b.sched_pos[e] % (int)b.S;
    __syncthreads();
    if (was_done) {  // rail_env.py:508-509 raises; here: status bit, or in-place reset
        if (flags & FL_FLAG_AUTO_RESET) {
            reset_env(b, e, false, !(flags & FL_FLAG_FRESH_AGENTS));
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
    // rail indices of the two cells (-1: private node); a cell outside the rail (an invalid world) sends the environment to
    // the pairwise comparison, which needs no tables
    int ri_cur = -1, ri_nxt = -1;
    bool off_rail = false;
    if (act && R_env > 0) {
        if (r >= 0) { const unsigned v = ridx[r * W + c]; ri_cur = (int)v; off_rail |= v >= (unsigned)R_env; }
        if (nr >= 0) {
            const unsigned v = (nr < H && nc >= 0 && nc < W) ? (unsigned)ridx[nr * W + nc] : 0xFFFFu;
            ri_nxt = (int)v; off_rail |= v >= (unsigned)R_env;
        } else if (r >= 0) off_rail = true;        // a move off the top edge of the grid

    }
    const bool use_map = R_env > 0 && !__syncthreads_or(off_rail);           // (also orders the writes above)
    if (!use_map) __syncthreads();

    // ---- MotionCheck (agent_chains.py:151-236) as a least fixpoint over CELL NODES:
    //        blocked(X) = some train on X stays | swaps | loses a contended cell | heads for a blocked node
    //      Several trains can share a cell (MALFUNCTION_OFF_MAP + STOP enters the map unchecked,
    //      state_machine.py:41-42); a node's verdict is shared by all of them and its "agent" attribute is
    //      the last one added, i.e. the highest handle (agent_chains.py:33) — its representative here.
    int rep_cur = i, rep_nxt = -1;
    bool sw = false;
    if (use_map) {
        // The same three facts from tables per rail cell.  m_rep[X]: highest handle standing on X (+1).  m_out[X]: bit m = a
        // train on X wants to leave in direction m.  m_want[X]: lowest representative among the nodes that want to move onto X.
        //   representative of my node / of my target: m_rep of the cell (an off-map train is its own node);
        //   swap: a train on my target wants to leave it towards me, i.e. in the direction opposite to mine;
        //   loser: somebody with a lower representative wants my target — trains of my own node share my representative and a
        //   train that stays on the target does not "want" it, so the two exclusions of the pairwise test hold by construction.
        if (act && ri_cur >= 0) atomicMax(&m_rep[ri_cur], i + 1);
        __syncthreads();
        if (act) {
            if (ri_cur >= 0) rep_cur = m_rep[ri_cur] - 1;
            if (nxt_id != cur_id) {
                atomicMin(&m_want[ri_nxt], rep_cur);
                if (ri_cur >= 0) atomicOr(&m_out[ri_cur], 1 << nd);
            }
            s_rep[i] = rep_cur;
        }
        __syncthreads();
        if (act) {
            rep_nxt = ri_nxt >= 0 ? m_rep[ri_nxt] - 1 : i;
            bool loser = false;
            if (nxt_id != cur_id) {
                sw = ri_cur >= 0 && ((m_out[ri_nxt] >> (nd ^ 2)) & 1);
                loser = m_want[ri_nxt] < rep_cur;
            }
            if (nxt_id == cur_id || sw || loser) s_blk[rep_cur] = 1;
        }
        __syncthreads();
    } else {
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
    }
    // (racecheck warns about the loop below: s_blk flags only ever go 0 -> 1 and the loop runs to the fixpoint, so reading a
    // flag while another thread sets it only decides in which round the reader follows)
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
                const unsigned dv = b.dist[(size_t)e * b.dist_stride + ((size_t)b.slot[ea] * HW + qr * W + qc) * 4 + d];
                const int len = dv == FL_DIST_INF ? 0 : (int)dv + 1;
                const int tt = (int)ceil((double)len / b.speed[ea]);
                rew = off_map(st) ? -tt : (b.latest[ea] - elapsed) - tt;
            }
            b.done[ea] = 1;
            // episode statistics (eval_env.py:81-94 final_metric): total reward, and "arrivals" with the reference's own
            // predicate `a.position is None and a.state != READY_TO_DEPART` — which also counts trains that never left
            // (WAITING, MALFUNCTION_OFF_MAP) and does not count a train standing on its target in state DONE (the
            // second-episode case described at reset_env)
            unsigned long long *stt = reinterpret_cast<unsigned long long *>(b.stats + (size_t)e * 4);
            if (r < 0 && st != READY) atomicAdd(stt + 1, 1ull);
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


}  // namespace
