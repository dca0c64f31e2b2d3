// scan.cuh — device-wide exclusive prefix sum (u32) with fused producer/consumer functors.
//
// Used for every stream compaction on the path (group lists, chain lists, anchor lists):
// the producer computes the 0/1 (or small count) value of element i on the fly, the consumer
// receives (i, exclusive_prefix, value).  ONE launch: tiles publish their aggregate, a warp resolves
// the tile's exclusive prefix by decoupled look-back (32 predecessors per round); inputs are read once.
// HBM bound.
#pragma once
#include "common.cuh"

namespace swg {

#ifndef SWG_SC_THREADS
#define SWG_SC_THREADS 256
#endif
#ifndef SWG_SC_ITEMS
#define SWG_SC_ITEMS 8
#endif
constexpr int SC_THREADS = SWG_SC_THREADS;
constexpr int SC_ITEMS = SWG_SC_ITEMS;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ u32 block_exclusive_scan_256(u32 v, u32 *ws /*[SC_THREADS / 32]*/, u32 &block_total) {
    u32 x = v;
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
        u32 t = ws[w];
        if (w < (int)warp) base += t;
        tot += t;
    }
    block_total = tot;
    __syncthreads();
    return base + x - v;
}

// ---- single-pass variant: decoupled look-back over tile aggregates (one launch, producer inputs read once) ----------
constexpr u64 SC_FLAG_AGG = 1ull << 62, SC_FLAG_INCL = 2ull << 62, SC_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <class In, class Out>
__global__ void __launch_bounds__(SC_THREADS) sc_onepass_kernel(In in, Out out, u32 n, u64 *status, u32 *tile_counter, u32 *total_out) {
    __shared__ u32 ws[SC_THREADS / 32];
    __shared__ u32 s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u); // in-order tile ids
    __syncthreads();
    const u32 tile = s_tile;
    const u32 lane = threadIdx.x & 31;
    // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile (measured faster than a
    // warp-striped layout with one warp scan per item: 430 vs 529 us for the six scans of a 20 M step)
    const u32 base = tile * SC_TILE + threadIdx.x * SC_ITEMS;
    u32 v[SC_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        u32 i = base + k;
        v[k] = i < n ? in(i) : 0;
        s += v[k];
    }
    u32 tot;
    const u32 ex_local = block_exclusive_scan_256(s, ws, tot);
    if (threadIdx.x < 32) { // warp 0 resolves the tile's exclusive prefix, 32 predecessors per round
        if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? SC_FLAG_INCL : SC_FLAG_AGG) | tot);
        u32 excl = 0;
        if (tile != 0) {
            i64 t = (i64)tile - 1;
            while (true) {
                const i64 mine = t - lane;
                const u64 w = mine >= 0 ? ld_relaxed_u64(&status[mine]) : SC_FLAG_INCL;
                const u32 incl = __ballot_sync(0xFFFFFFFFu, (w & SC_FLAG_INCL) != 0);
                const u32 ready = __ballot_sync(0xFFFFFFFFu, (w & (SC_FLAG_INCL | SC_FLAG_AGG)) != 0);
                const u32 upto = incl ? (u32)(__ffs(incl) - 1) : 31u;     // nearest predecessor holding an inclusive prefix
                const u32 need = upto == 31 ? 0xFFFFFFFFu : ((2u << upto) - 1);
                if ((ready & need) != need) continue;                     // somebody in range has not published yet: poll again
                const u32 part = __reduce_add_sync(0xFFFFFFFFu, lane <= upto ? (u32)(w & SC_VAL_MASK) : 0u);
                excl += part;
                if (incl) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_INCL | (u64)(excl + tot));
        }
        if (lane == 0) {
            s_excl = excl;
            if ((u64)(tile + 1) * SC_TILE >= n) *total_out = excl + tot;
        }
    }
    __syncthreads();
    u32 ex = s_excl + ex_local;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        u32 i = base + k;
        if (i < n) out(i, ex, v[k]);
        ex += v[k];
    }
}

// ---- 0/1 flags: the same one-pass scan with a STRIPED tile (element k*256 + t of the tile belongs to thread t): every load
// and store of the functors is coalesced, and the in-tile scan is eight ballots + one 64-entry scan instead of a shuffle scan
// per thread.  in(i) -> bool; out(i, exclusive_prefix, flag).  Used for every stream compaction and head-flag numbering.
template <class In, class Out>
__global__ void __launch_bounds__(SC_THREADS) sc_flags_kernel(In in, Out out, u32 n, u64 *status, u32 *tile_counter, u32 *total_out) {
    static_assert(SC_THREADS == 256 && SC_ITEMS == 8, "striped flag scan: 8 warps x 8 items");
    __shared__ u32 s_cnt[64]; // [item][warp] counts, then exclusive prefixes in element order
    __shared__ u32 s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u); // in-order tile ids
    __syncthreads();
    const u32 tile = s_tile;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 base = tile * SC_TILE;
    u32 m[SC_ITEMS];
    u32 fbits = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        const u32 i = base + k * SC_THREADS + threadIdx.x;
        const bool f = i < n && in(i);
        fbits |= (f ? 1u : 0u) << k;
        m[k] = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) s_cnt[k * 8 + warp] = __popc(m[k]);
    }
    __syncthreads();
    if (threadIdx.x < 32) { // 64 counts -> exclusive prefixes (two per lane), tile total, look-back
        const u32 a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
        u32 x = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (u32)o) x += t;
        }
        const u32 tot = __shfl_sync(0xFFFFFFFFu, x, 31);
        s_cnt[2 * lane] = x - a - b;
        s_cnt[2 * lane + 1] = x - b;
        if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? SC_FLAG_INCL : SC_FLAG_AGG) | tot);
        u32 excl = 0;
        if (tile != 0) {
            i64 t = (i64)tile - 1;
            while (true) {
                const i64 mine = t - lane;
                const u64 w = mine >= 0 ? ld_relaxed_u64(&status[mine]) : SC_FLAG_INCL;
                const u32 incl = __ballot_sync(0xFFFFFFFFu, (w & SC_FLAG_INCL) != 0);
                const u32 ready = __ballot_sync(0xFFFFFFFFu, (w & (SC_FLAG_INCL | SC_FLAG_AGG)) != 0);
                const u32 upto = incl ? (u32)(__ffs(incl) - 1) : 31u;
                const u32 need = upto == 31 ? 0xFFFFFFFFu : ((2u << upto) - 1);
                if ((ready & need) != need) continue;
                excl += __reduce_add_sync(0xFFFFFFFFu, lane <= upto ? (u32)(w & SC_VAL_MASK) : 0u);
                if (incl) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_INCL | (u64)(excl + tot));
        }
        if (lane == 0) {
            s_excl = excl;
            if ((u64)(tile + 1) * SC_TILE >= n) *total_out = excl + tot;
        }
    }
    __syncthreads();
    const u32 lt = (1u << lane) - 1;
    const u32 tile_excl = s_excl;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        const u32 i = base + k * SC_THREADS + threadIdx.x;
        if (i < n) out(i, tile_excl + s_cnt[k * 8 + warp] + __popc(m[k] & lt), (fbits >> k) & 1u);
    }
}
template <class In, class Out>
static inline void scan_flags(In in, Out out, u32 n, u32 *temp, u32 *total_out, cudaStream_t st, LaunchCounter &lc) {
    if (n == 0) {
        SWG_CUDA(cudaMemsetAsync(total_out, 0, sizeof(u32), st));
        return;
    }
    u32 nb = cdiv(n, SC_TILE);
    SWG_CUDA(cudaMemsetAsync(temp, 0, sizeof(u32) * (2 * (size_t)nb + 4), st));
    u64 *status = reinterpret_cast<u64 *>(temp);
    u32 *counter = temp + 2 * (size_t)nb + 2;
    sc_flags_kernel<<<nb, SC_THREADS, 0, st>>>(in, out, n, status, counter, total_out);
    lc.n += 1;
    SWG_CUDA(cudaGetLastError());
}

// ---- segmented inclusive max-scan (u32), one pass ---------------------------------------------------------------------------
// out(i, m): m = max of val(j) over the j <= i of i's segment (head(i): i starts a segment).  The running maximum of the interval
// ends of a group in start order: "does anything at or left of k reach position p" is one load (the halo test of the plane sweeps).
// Same tiling and decoupled look-back as sc_onepass_kernel; a predecessor tile that holds a segment head ends the look-back like
// one that holds an inclusive value.  Tile status: flags | has_head << 32 | max.
template <class Head, class Val, class Out>
__global__ void __launch_bounds__(SC_THREADS) sc_segmax_kernel(Head head, Val val, Out out, u32 n, u64 *status, u32 *tile_counter) {
    __shared__ u32 wv[SC_THREADS / 32], wf[SC_THREADS / 32];
    __shared__ u32 s_tile, s_carry;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 base = tile * SC_TILE + threadIdx.x * SC_ITEMS;
    u32 v[SC_ITEMS];
    u32 fbits = 0, run = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        const u32 i = base + k;
        if (i < n) {
            const u32 x = val(i);
            if (head(i)) { run = x; fbits |= 1u << k; } else run = max(run, x);
        }
        v[k] = run;
    }
    // warp-level segmented inclusive scan of the thread aggregates (a: max since the last head, f: a head seen)
    u32 a = run, f = fbits ? 1u : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 ta = __shfl_up_sync(0xFFFFFFFFu, a, o), tf = __shfl_up_sync(0xFFFFFFFFu, f, o);
        if (lane >= (u32)o) { if (!f) a = max(a, ta); f |= tf; }
    }
    u32 ea = __shfl_up_sync(0xFFFFFFFFu, a, 1), ef = __shfl_up_sync(0xFFFFFFFFu, f, 1); // exclusive, inside the warp
    if (lane == 0) { ea = 0; ef = 0; }
    if (lane == 31) { wv[warp] = a; wf[warp] = f; }
    __syncthreads();
    u32 ca = 0, cf = 0, ta_ = 0, tf_ = 0; // carry of the warps before this one; aggregate of the whole tile
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
        const u32 x = wv[w], h = wf[w];
        if (w < (int)warp) { if (h) { ca = x; cf = 1; } else ca = max(ca, x); }
        if (h) { ta_ = x; tf_ = 1; } else ta_ = max(ta_, x);
    }
    const u32 xa = ef ? ea : max(ca, ea), xf = ef | cf; // exclusive carry of this thread inside the tile
    if (threadIdx.x < 32) {
        if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? SC_FLAG_INCL : SC_FLAG_AGG) | ((u64)tf_ << 32) | ta_);
        u32 excl = 0;
        if (tile != 0) {
            i64 t = (i64)tile - 1;
            while (true) {
                const i64 mine = t - lane;
                const u64 w = mine >= 0 ? ld_relaxed_u64(&status[mine]) : SC_FLAG_INCL;
                const u32 term = __ballot_sync(0xFFFFFFFFu, (w & SC_FLAG_INCL) != 0 || ((w & SC_FLAG_AGG) != 0 && ((w >> 32) & 1)));
                const u32 ready = __ballot_sync(0xFFFFFFFFu, (w & (SC_FLAG_INCL | SC_FLAG_AGG)) != 0);
                const u32 upto = term ? (u32)(__ffs(term) - 1) : 31u;
                const u32 need = upto == 31 ? 0xFFFFFFFFu : ((2u << upto) - 1);
                if ((ready & need) != need) continue;
                excl = max(excl, __reduce_max_sync(0xFFFFFFFFu, lane <= upto ? (u32)w : 0u));
                if (term) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_INCL | (tf_ ? ta_ : max(excl, ta_)));
        }
        if (lane == 0) s_carry = excl;
    }
    __syncthreads();
    const u32 carry = xf ? xa : max(xa, s_carry); // max over the segment's items before this thread's first item
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        const u32 i = base + k;
        if (i < n) out(i, (fbits & ((2u << k) - 1)) ? v[k] : max(v[k], carry));
    }
}
template <class Head, class Val, class Out>
static inline void scan_segmax(Head head, Val val, Out out, u32 n, u32 *temp, cudaStream_t st, LaunchCounter &lc) {
    if (n == 0) return;
    u32 nb = cdiv(n, SC_TILE);
    SWG_CUDA(cudaMemsetAsync(temp, 0, sizeof(u32) * (2 * (size_t)nb + 4), st));
    u64 *status = reinterpret_cast<u64 *>(temp);
    u32 *counter = temp + 2 * (size_t)nb + 2;
    sc_segmax_kernel<<<nb, SC_THREADS, 0, st>>>(head, val, out, n, status, counter);
    lc.n += 1;
    SWG_CUDA(cudaGetLastError());
}

// ---- unordered compaction: the elements with pred(i) appended to list[] in no particular order (one counter atomic per
// 1024-element block).  For lists whose consumers treat every entry independently (candidate lists): ~3x cheaper than the
// ordered scan above.  *counter must be zero on entry.
template <class Pred, class Out>
__global__ void __launch_bounds__(256) k_compact_unordered(Pred pred, Out out, u32 n, u32 *counter) {
    __shared__ u32 s_w[8][4];
    __shared__ u32 s_base;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 base = blockIdx.x * 1024;
    bool f[4];
    u32 m[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 i = base + k * 256 + threadIdx.x;
        f[k] = i < n && pred(i);
        m[k] = __ballot_sync(0xFFFFFFFFu, f[k]);
        if (lane == 0) s_w[warp][k] = __popc(m[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int k = 0; k < 4; k++)
            for (int w = 0; w < 8; w++) { const u32 t = s_w[w][k]; s_w[w][k] = tot; tot += t; }
        s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const u32 lt = (1u << lane) - 1;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (f[k]) out(base + k * 256 + threadIdx.x, s_base + s_w[warp][k] + __popc(m[k] & lt));
}
// out(i, slot) is called once per selected element with its slot in [0, *counter)
template <class Pred, class Out> static inline void compact_unordered(Pred pred, Out out, u32 n, u32 *counter, cudaStream_t st, LaunchCounter &lc) {
    SWG_CUDA(cudaMemsetAsync(counter, 0, sizeof(u32), st));
    if (n == 0) return;
    k_compact_unordered<<<cdiv(n, 1024), 256, 0, st>>>(pred, out, n, counter);
    lc.n += 1;
}

// temp: scan_temp_u32(n) u32 (tile status words + tile counter); total_out: device u32
template <class In, class Out>
static inline void scan_apply(In in, Out out, u32 n, u32 *temp, u32 *total_out, cudaStream_t st, LaunchCounter &lc) {
    if (n == 0) {
        SWG_CUDA(cudaMemsetAsync(total_out, 0, sizeof(u32), st));
        return;
    }
    u32 nb = cdiv(n, SC_TILE);
    SWG_CUDA(cudaMemsetAsync(temp, 0, sizeof(u32) * (2 * (size_t)nb + 4), st));
    u64 *status = reinterpret_cast<u64 *>(temp);
    u32 *counter = temp + 2 * (size_t)nb + 2;
    sc_onepass_kernel<<<nb, SC_THREADS, 0, st>>>(in, out, n, status, counter, total_out);
    lc.n += 1;
    SWG_CUDA(cudaGetLastError());
}
static inline size_t scan_temp_u32(u32 n) { return 2 * (size_t)cdiv(n, SC_TILE) + 4; }

} // namespace swg
