// walks.cuh — k_walks: static branch-walk tables, built once per world at reset.
//
// The tree observation explores the rail from a (cell, direction) state to the next switch, dead end or
// revisited state (treeobs.cpp:258-610, _explore_branch).  Which cells such a walk passes is a property of
// the rail alone — only what it meets on them (trains, predicted conflicts, the observer's own target) is
// dynamic.  k_walks therefore lists, for every rail state, the states its walk visits, so that k_observe can
// spread the cells of one walk over the lanes of a lane group instead of chasing them one by one:
//   ridx[cell]     rail index of a cell (0xFFFF = no rail); state id sid = 4 * ridx[cell] + direction
//   srec[sid]      row | col << 10 | dir << 20 | transitions nibble of (cell, dir) << 22 | "unusable switch here" << 26
//   wrec[sid]      16-byte record of the walk from sid:
//                  x  offset of the walk in wlist
//                  y  steps (the walk visits steps + 1 states) | kind << 16 | number of target hits << 20; kind:
//                     1 = ends on a switch, 2 = dead end, 3 = the last state revisits an earlier one (rail cycle,
//                     treeobs.cpp:476-481), 0 = ends on a cell without transitions (treeobs.cpp:527-535 throws)
//                  z  child 0 | child 1 << 16      state ids the children (left, forward, right; treeobs.cpp:583-608) of the
//                  w  child 2 | k_unusable << 16   node ending this walk start from, 0xFFFF = null child (always for walks
//                                                  ending in a cycle or a bad cell); k_unusable = first step < steps that
//                                                  stands on an unusable switch (treeobs.cpp:497-513), 0xFFFF = none
//   whoff[sid]     offset of the walk's target hits in whits (valid when the record counts any)
//   whits[...]     step | slot << 16 for every state of the walk standing on the target cell of a unique-target slot, in
//                  walk order: the observer's own target ends its walk early (treeobs.cpp:467-475, 483-489)
//   wlist[...]     the visited states, walk after walk: state id | transitions nibble of the state << 16
//   kcls[rail]     the reference keys predicted positions by c * W + r (treeobs.cpp:50-65, 379-465), which is not unique when
//                  H > W: cells (r, c) and (r + W, c - 1) share a key and conflict with each other's predictions.  kcls maps
//                  a rail cell to the lowest rail index of its key class (itself when H <= W)
//   sdist[slot][sid] the distance map of fl_distance_map indexed by state id
//   gtab[slot][sid]  how the shortest path to the slot's target continues after the walk from sid: next state | last step of
//                  the walk on the path << 16 | direction of the state at that step << 30
// FILL = false only measures (states, list length, hits) so that the host can size wlist and whits.
#pragma once
#include "common.cuh"

namespace {

constexpr int WK_SWITCH = 1, WK_DEADEND = 2, WK_CYCLE = 3, WK_BAD = 0;

// one step of a branch walk ignoring everything dynamic (treeobs.cpp:476-539); returns false when the walk
// ends on this state and sets kind
DEVI bool static_succ(const uint16_t *__restrict__ g, int H, int W, int &r, int &c, int &d, int &kind) {
    const unsigned gc = g[r * W + c];
    const int nb = nibble(gc, d), num = __popc(nb);
    int total = __popc(gc);
    if (gc == 0x8421u) total = 2;                    // diamond crossing
    if (num == 1) {
        if (total == 1) { kind = WK_DEADEND; return false; }
        const int nd = first_dir(nb), rr = r + d_row(nd), cc = c + d_col(nd);
        if (rr < 0 || cc < 0 || rr >= H || cc >= W || !g[rr * W + cc]) { kind = WK_BAD; return false; }  // rail into nothing
        d = nd; r = rr; c = cc;
        return true;
    }
    kind = num > 1 ? WK_SWITCH : WK_BAD;
    return false;
}

// steps and kind of the walk from (r, c, d); bound = number of states of the world (a longer walk must cycle)
DEVI int static_walk_len(const uint16_t *__restrict__ g, int H, int W, int r, int c, int d, int bound, int &kind) {
    const int r0 = r, c0 = c, d0 = d;
    int steps = 0;
    while (static_succ(g, H, W, r, c, d, kind)) {
        if (++steps > bound) {                       // rail cycle: Brent for the index of the first revisit
            int tr_ = r0, tc_ = c0, td = d0, hr = r0, hc = c0, hd = d0, power = 1, lam = 1, k2;
            static_succ(g, H, W, hr, hc, hd, k2);
            while (tr_ != hr || tc_ != hc || td != hd) {
                if (power == lam) { tr_ = hr; tc_ = hc; td = hd; power *= 2; lam = 0; }
                static_succ(g, H, W, hr, hc, hd, k2);
                lam++;
            }
            tr_ = hr = r0; tc_ = hc = c0; td = hd = d0;
            for (int k = 0; k < lam; k++) static_succ(g, H, W, hr, hc, hd, k2);
            int mu = 0;
            while (tr_ != hr || tc_ != hc || td != hd) {
                static_succ(g, H, W, tr_, tc_, td, k2);
                static_succ(g, H, W, hr, hc, hd, k2);
                mu++;
            }
            kind = WK_CYCLE;
            return mu + lam;
        }
    }
    return steps;
}

template <int NT>
DEVI uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_part, uint32_t &total) {
    const int tid = threadIdx.x;
    s_part[tid] = v;
    __syncthreads();
    for (int off = 1; off < NT; off <<= 1) {
        const uint32_t x = tid >= off ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += x;
        __syncthreads();
    }
    total = s_part[NT - 1];
    const uint32_t ex = s_part[tid] - v;
    __syncthreads();
    return ex;
}

// slot of the target standing on `cell`, -1 = none.  tbits: one bit per cell, set when some slot's target is there.
DEVI int target_slot_at(const FlBatch &b, int e, const uint32_t *tbits, int cell, int W) {
    if (!((tbits[cell >> 5] >> (cell & 31)) & 1u)) return -1;
    const int16_t *rc = b.slot_rc + (size_t)e * b.n_slots * 2;
    const int r = cell / W, c = cell - r * W;
    for (int s = 0; s < (int)b.n_slots; s++) if (rc[2 * s] == r && rc[2 * s + 1] == c) return s;
    return -1;
}

template <bool FILL, int NT>
__global__ void __launch_bounds__(NT) k_walks(FlBatch b, const int32_t *__restrict__ env_ids) {
    const int e = env_ids ? env_ids[blockIdx.x] : (int)blockIdx.x, H = (int)b.H, W = (int)b.W, HW = H * W, tid = threadIdx.x;
    __shared__ uint32_t s_part[NT];
    extern __shared__ uint32_t s_tbits[];            // (HW + 31) / 32 words: cells holding the target of a slot
    const uint16_t *__restrict__ g = b.grid + (size_t)e * b.grid_stride;
    uint16_t *ridx = b.ridx + (size_t)e * b.ridx_stride;
    int32_t *tot = b.walk_total + (size_t)e * 4;

    // new tables: every cached tree structure of the environment is stale (observe.cuh; lane 31's first word is the key)
    if (FILL && b.tree_cache)
        for (int i = tid; i < (int)b.N; i += NT) b.tree_cache[((size_t)e * b.N + i) * FL_TREE_CACHE_WORDS + 31] = 0u;
    if (FILL && b.path_cache)
        for (int i = tid; i < (int)b.N; i += NT) b.path_cache[((size_t)e * b.N + i) * b.pc_stride] = 0ull;
    for (int k = tid; k < (HW + 31) / 32; k += NT) s_tbits[k] = 0;
    __syncthreads();
    for (int s = tid; s < (int)b.n_slots; s += NT) {
        const int r = b.slot_rc[((size_t)e * b.n_slots + s) * 2], c = b.slot_rc[((size_t)e * b.n_slots + s) * 2 + 1];
        if (r >= 0 && c >= 0 && r < H && c < W) atomicOr(&s_tbits[(r * W + c) >> 5], 1u << ((r * W + c) & 31));
    }
    // rail indices in cell order
    const int per = (HW + NT - 1) / NT, lo = min(tid * per, HW), hi = min(lo + per, HW);
    uint32_t cnt = 0;
    for (int k = lo; k < hi; k++) cnt += g[k] != 0;
    uint32_t n_rail;
    uint32_t run = block_exclusive_scan<NT>(cnt, s_part, n_rail);
    const int S = (int)n_rail * 4;
    if (!FILL) {
        // list length = sum over states of (steps + 1); hits = states of all walks standing on a slot's target
        uint32_t len = 0, nh = 0;
        for (int cs = tid; cs < HW * 4; cs += NT) {
            const int cell = cs >> 2, d0 = cs & 3;
            if (!g[cell]) continue;
            int kind, r = cell / W, c = cell % W, d = d0;
            const int steps = static_walk_len(g, H, W, r, c, d, S, kind);
            len += (uint32_t)steps + 1u;
            for (int k = 0; k <= steps; k++) {
                nh += (s_tbits[(r * W + c) >> 5] >> ((r * W + c) & 31)) & 1u;
                if (k < steps) static_succ(g, H, W, r, c, d, kind);
            }
        }
        uint32_t total, total_h;
        block_exclusive_scan<NT>(len, s_part, total);
        block_exclusive_scan<NT>(nh, s_part, total_h);
        if (tid == 0) { tot[0] = S; tot[1] = (int)total; tot[2] = (int)total_h; tot[3] = 0; }
        return;
    }
    uint32_t *srec = b.srec + (size_t)e * b.state_stride;
    uint4 *wrec = reinterpret_cast<uint4 *>(b.wrec) + (size_t)e * b.state_stride;
    uint32_t *whoff = b.whoff + (size_t)e * b.state_stride;
    uint32_t *wlist = b.wlist + (size_t)e * b.wlist_stride;
    uint32_t *whits = b.whits + (size_t)e * b.whits_stride;
    for (int k = lo; k < hi; k++) {
        const unsigned gc = g[k];
        if (!gc) { ridx[k] = 0xFFFF; continue; }
        ridx[k] = (uint16_t)run;
        int total = __popc(gc);
        if (gc == 0x8421u) total = 2;
        for (int d = 0; d < 4; d++) {
            const int nb = nibble(gc, d);
            srec[run * 4 + d] = (uint32_t)(k / W) | ((uint32_t)(k % W) << 10) | ((uint32_t)d << 20) | ((uint32_t)nb << 22) |
                                ((uint32_t)(total > 2 && __popc(nb) < 2) << 26);
        }
        run++;
    }
    for (int k = HW + tid; k < (int)b.ridx_stride; k += NT) ridx[k] = 0xFFFF;
    __syncthreads();
    uint16_t *kcls = b.kcls + (size_t)e * b.state_stride;
    for (int k = lo; k < hi; k++) {
        const unsigned ri = ridx[k];
        if (ri == 0xFFFF) continue;
        const int r = k / W, c = k - r * W;
        unsigned cls = ri;
        for (int m = -(r / W); m < 0; m++) {             // aliases in front of this cell in cell order: (r + m W, c - m)
            const int rr = r + m * W, cc = c - m;
            if (cc < W && ridx[rr * W + cc] != 0xFFFF) { cls = ridx[rr * W + cc]; break; }
        }
        kcls[ri] = (uint16_t)cls;
    }
    for (int k = (int)n_rail + tid; k < (int)b.state_stride; k += NT) kcls[k] = 0xFFFF;
    // walk lengths, offsets, lists: thread t owns the contiguous states [slo, shi)
    const int sper = (S + NT - 1) / NT, slo = min(tid * sper, S), shi = min(slo + sper, S);
    uint32_t len = 0, nh = 0;
    for (int sid = slo; sid < shi; sid++) {
        const uint32_t rec = srec[sid];
        int r = (int)(rec & 1023), c = (int)((rec >> 10) & 1023), d = (int)((rec >> 20) & 3);
        int kind = WK_BAD;
        const int steps = static_walk_len(g, H, W, r, c, d, S, kind);
        uint32_t h = 0;
        for (int k = 0; k <= steps; k++) {
            h += (s_tbits[(r * W + c) >> 5] >> ((r * W + c) & 31)) & 1u;
            if (k < steps) { int k2; static_succ(g, H, W, r, c, d, k2); }
        }
        wrec[sid].y = (uint32_t)min(steps, 0xFFFF) | ((uint32_t)kind << 16) | (min(h, 255u) << 20);
        len += (uint32_t)steps + 1u;
        nh += h;
    }
    uint32_t total, total_h;
    uint32_t off = block_exclusive_scan<NT>(len, s_part, total);
    uint32_t hoff = block_exclusive_scan<NT>(nh, s_part, total_h);
    for (int sid = slo; sid < shi; sid++) {
        const uint32_t rec = srec[sid];
        int r = (int)(rec & 1023), c = (int)((rec >> 10) & 1023), d = (int)((rec >> 20) & 3), kind;
        const uint32_t y = wrec[sid].y;
        const int steps = (int)(y & 0xFFFFu), wk = (int)((y >> 16) & 15u);
        unsigned kunus = 0xFFFF;
        const uint32_t w_off = off;
        whoff[sid] = hoff;
        for (int k = 0; k <= steps; k++) {
            const int cell = r * W + c;
            const unsigned ri = ridx[cell];
            if (off < (uint32_t)b.wlist_stride) wlist[off] = (ri * 4 + d) | (((srec[ri * 4 + d] >> 22) & 15u) << 16);
            off++;
            if (k < steps && kunus == 0xFFFF && ((srec[ri * 4 + d] >> 26) & 1u)) kunus = (unsigned)k;
            const int ts = target_slot_at(b, e, s_tbits, cell, W);
            if (ts >= 0) { if (hoff < (uint32_t)b.whits_stride) whits[hoff] = (uint32_t)k | ((uint32_t)ts << 16); hoff++; }
            if (k < steps) static_succ(g, H, W, r, c, d, kind);
        }
        // children of the node this walk ends in, from its last state (r, c, d)
        const int enb = nibble(g[r * W + c], d);
        unsigned ch[3];
        for (int a2 = -1; a2 <= 1; a2++) {
            const int bd = (d + a2) & 3, rb = (bd + 2) & 3;
            int cd = -1;
            if (wk == WK_DEADEND && tbit(enb, rb)) cd = rb;
            else if (wk == WK_SWITCH && tbit(enb, bd)) cd = bd;
            unsigned cs = 0xFFFF;
            if (cd >= 0) {
                const int rr = r + d_row(cd), cc = c + d_col(cd);
                if (rr >= 0 && cc >= 0 && rr < H && cc < W && ridx[rr * W + cc] != 0xFFFF) cs = ridx[rr * W + cc] * 4u + (unsigned)cd;
            }
            ch[a2 + 1] = cs;
        }
        wrec[sid] = make_uint4(w_off, y, ch[0] | (ch[1] << 16), ch[2] | (kunus << 16));
    }
    for (int sid = S + tid; sid < (int)b.state_stride; sid += NT) {
        srec[sid] = 0; whoff[sid] = 0;
        wrec[sid] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
    // the distance map by state id (the observation never needs it off the rail)
    __syncthreads();
    {
        uint16_t *sd = b.sdist + (size_t)e * b.n_slots * b.state_stride;
        const uint16_t *dist = b.dist + (size_t)e * b.dist_stride;
        const int ss = (int)b.state_stride, n = (int)b.n_slots * ss;
        for (int idx = tid; idx < n; idx += NT) {
            const int s = idx / ss, sid = idx - s * ss;
            unsigned v = FL_DIST_INF;
            if (sid < S) {
                const uint32_t rec = srec[sid];
                v = dist[((size_t)s * HW + (size_t)((rec & 1023) * W + ((rec >> 10) & 1023))) * 4 + ((rec >> 20) & 3)];
            }
            sd[idx] = (uint16_t)v;
        }
    }
    // where the shortest path to each slot's target continues after the walk from a state (predictions.cpp:13-144): the walk
    // is followed to its end or to the slot's target; at a switch or dead end the child state with the lowest distance is
    // taken, the first of equals in the order left, forward, right
    __syncthreads();
    __shared__ int s_long;
    if (tid == 0) s_long = 0;
    __syncthreads();
    {
        uint32_t *gt = b.gtab + (size_t)e * b.n_slots * b.state_stride;
        const uint16_t *sd = b.sdist + (size_t)e * b.n_slots * b.state_stride;
        const int ss = (int)b.state_stride, n = (int)b.n_slots * ss;
        for (int idx = tid; idx < n; idx += NT) {
            const int s = idx / ss, sid = idx - s * ss;
            uint32_t g = 0xFFFFu;
            if (sid < S) {
                const uint4 w = wrec[sid];
                const int L = (int)(w.y & 0xFFFFu), kind = (int)((w.y >> 16) & 15u), nh = (int)((w.y >> 20) & 255u);
                int kend = L;
                bool hit = false;
                for (int q = 0; q < nh; q++) {
                    const uint32_t hv = whits[whoff[sid] + q];
                    if ((int)(hv >> 16) == s) { kend = (int)(hv & 0xFFFFu); hit = true; break; }
                }
                unsigned nxt = 0xFFFFu;
                if (!hit && (kind == WK_SWITCH || kind == WK_DEADEND)) {
                    const unsigned ch[3] = {w.z & 0xFFFFu, w.z >> 16, w.w & 0xFFFFu};
                    unsigned best = FL_DIST_INF;
                    for (int j = 0; j < 3; j++)
                        if (ch[j] != 0xFFFFu) { const unsigned v = sd[(size_t)s * ss + ch[j]]; if (v < best) { best = v; nxt = ch[j]; } }
                }
                if (kend > 0x3FFF) { s_long = 1; kend = 0x3FFF; }
                const unsigned edir = wlist[w.x + kend] & 3u;
                g = nxt | ((uint32_t)kend << 16) | (edir << 30);
            }
            gt[idx] = g;
        }
    }
    __syncthreads();
    if (tid == 0) { tot[0] = S; tot[1] = (int)total; tot[2] = (int)total_h; tot[3] = s_long; }
}

}  // namespace
