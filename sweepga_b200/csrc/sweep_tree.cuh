// sweep_tree.cuh — the n = 1 plane sweep on DEEP PILES (repeat arrays, centromeres), position-parallel and exact.
//
// Replaces the per-event walk of the active set (src/plane_sweep_exact.rs:197-259 `mark_good`, :306-351): the reference touches
// every active mapping at every event position, O(n * depth); a 1 M-mapping pile 5000 deep took the warp-per-group kernel of round 1
// (k_sweep_groups: rank-ordered active array, insertion per event) 20 s.
//
// With n = 1 the rule is (k_sweep_flat1's header): after the events of position p the active set is S(p) = { k : start_k <= p < end_k };
// item m is GOOD iff it is the best of S(p) — (score desc, start asc, index asc) — at some event position p in [start_m, end_m), and
// FLAGGED iff at some such p the best is another item b with overlap(m, b) > threshold.  So all that is needed is best(p) for every
// event position p of the group:
//   positions   the distinct starts and ends of the group's items, sorted: P[0 .. M) (one radix sort of 2 n events); an item covers
//               the index range [l, r) = [index of its start, index of its end)
//   ranks       the items in (score desc, start asc, index asc) order: one stable radix sort by the score key of the items, which
//               already are in (start, index) order
//   tree        a segment tree over P (k_pile_paint): every item writes its rank with atomicMin into the O(log M) nodes that tile
//               [l, r); best(p) = the smallest rank on the path from leaf p to the root (k_pile_best)
//   runs        stretches of P with one best item (head-flag scan): in a pile the best changes when a better item begins or the best
//               ends, i.e. a run spans hundreds of event positions
//   verdict     k_pile_verdict, thread per item: the runs that intersect [l, r) — usually one to three — decide good / flagged.
// Work O(n log n), no sequential walk.  Groups of k_sweep_flat1 whose neighbour scans get long are sent here (n = 1); other limits
// keep the warp-per-group walk.
#pragma once
#include "filter_kernels.cuh"

namespace swg {

// big groups -> their items: x-th pile item = sorted position bitem[x] of pile group bgrp[x]
__global__ void __launch_bounds__(256) k_pile_expand(const u32 *__restrict__ big_list, const u32 *__restrict__ boff, u32 n_big, u32 n_items,
                                                     const u32 *__restrict__ gstart, u32 *__restrict__ bitem, u32 *__restrict__ bgrp) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_items) return;
    u32 lo = 0, hi = n_big; // last b with boff[b] <= x
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (boff[mid] <= x) lo = mid; else hi = mid;
    }
    bitem[x] = gstart[big_list[lo]] + (x - boff[lo]);
    bgrp[x] = lo;
}

// two events per item: key = (pile group << 32) | position, payload = 2 x (begin) / 2 x + 1 (end); zero-length items get none
// (they are inserted and removed at one position and never evaluated: plane_sweep_exact.rs:316-331) — their keys sort last
__global__ void __launch_bounds__(256) k_pile_events(const u32 *__restrict__ bitem, const u32 *__restrict__ bgrp, const SweepItem *__restrict__ sdata,
                                                     u32 n_items, u64 *__restrict__ ek, u32 *__restrict__ ev, u64 *__restrict__ rk, u32 *__restrict__ rv) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_items) return;
    const SweepItem d = sdata[bitem[x]];
    const bool live = d.end > d.start;
    const u64 g = (u64)bgrp[x] << 32;
    ek[2 * x] = live ? (g | d.start) : NONE64;
    ek[2 * x + 1] = live ? (g | d.end) : NONE64;
    ev[2 * x] = 2 * x;
    ev[2 * x + 1] = 2 * x + 1;
    rk[x] = d.skey;
    rv[x] = x;
}

// segment tree over `size` leaves (power of two), nodes 1 .. 2 size - 1, all NONE32 on entry: item x paints rank[x] on [l, r)
__global__ void __launch_bounds__(256) k_pile_paint(const u32 *__restrict__ upos, const u32 *__restrict__ rank_of, u32 n_items, u32 size,
                                                    u32 *__restrict__ tree) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_items) return;
    u32 l = upos[2 * x], r = upos[2 * x + 1];
    if (l == NONE32 || l >= r) return;
    const u32 rk = rank_of[x];
    for (l += size, r += size; l < r; l >>= 1, r >>= 1) {
        if (l & 1) atomicMin(&tree[l++], rk);
        if (r & 1) atomicMin(&tree[--r], rk);
    }
}
__global__ void __launch_bounds__(256) k_pile_best(const u32 *__restrict__ tree, u32 n_pos, u32 size, u32 *__restrict__ best) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    u32 v = NONE32;
    for (u32 x = i + size; x >= 1; x >>= 1) v = min(v, tree[x]);
    best[i] = v;
}

// thread per pile item: the runs of constant best(p) that intersect its index range decide
__global__ void __launch_bounds__(256) k_pile_verdict(const u32 *__restrict__ bitem, const u32 *__restrict__ sitem, const SweepItem *__restrict__ sdata,
                                                      const u32 *__restrict__ upos, const u32 *__restrict__ rank_of,
                                                      const u32 *__restrict__ item_of_rank, const u32 *__restrict__ run_id,
                                                      const u32 *__restrict__ run_first, const u32 *__restrict__ run_best,
                                                      const u32 *__restrict__ n_runs_ptr, u32 n_items, double thr, u8 *__restrict__ keep,
                                                      u64 *__restrict__ ctr) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    u32 near = 0;
    if (x < n_items) {
        const u32 u = bitem[x];
        const SweepItem me = sdata[u];
        const u32 l = upos[2 * x], r = upos[2 * x + 1];
        bool good = false, flag = false;
        if (me.end > me.start && l != NONE32) {
            const u32 mine = rank_of[x], n_runs = *n_runs_ptr;
            const bool want_flag = thr < 1.0;
            for (u32 j = run_id[l]; j < n_runs && run_first[j] < r; j++) {
                const u32 b = run_best[j];
                if (b == mine) good = true;
                else if (b != NONE32) {
                    const SweepItem o = sdata[bitem[item_of_rank[b]]];
                    near += near_tie(o.skey, o.start, o.end, me.skey, me.start, me.end) ? 1u : 0u;
                    if (want_flag && overlaps_more_than(me.start, me.end, o.start, o.end, thr)) { flag = true; break; }
                }
                if (good && !want_flag) break;
            }
        }
        keep[sitem[u]] = (good && !flag) ? 1 : 0;
    }
    const int slots[1] = {C_NEAR_TIES};
    const u32 vals[1] = {near};
    block_count_add<1>(ctr, slots, vals);
}

} // namespace swg
