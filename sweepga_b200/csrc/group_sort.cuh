// group_sort.cuh — the record sort of the chaining stage as a counting sort by group + an ordering step inside every group.
//
// Replaces, like radix_sort.cuh, the (query, target, strand) IndexMap + stable sort_by_key(query_start) of
// src/paf_filter.rs:761-777 — the same result (groups in (query, target, strand) order, every group ordered by
// (query_start, input index)), produced with far fewer passes over the records when no group is huge:
//
//   runs     (scan_flags in the caller) a run = maximal stretch of consecutive input records with the same group key; every
//            record learns its run, every run its first record.  Aligner output is grouped, so a group is one run or a few.
//   count    k_gs_run_base: the run's slot range inside its group from an atomicAdd on a dense table indexed by
//            (query * n_seq + target) * 2 + strand.  Runs of one group handled by the same warp share one atomic and take
//            their ranges in input order; across warps the ranges follow the arrival of the atomics.
//   scan     k_gs_scan: one pass over the table: start of every non-empty group, dense group numbers, largest group
//   scatter  k_gs_scatter: record -> sort word (query_start << ib) | index at start(group) + base(run) + offset in run, and
//            the group index of that position
//   emit     k_gs_emit: one flat pass writes the final form of every position and marks the groups in which some word is
//            smaller than its predecessor.  Input that is sorted by position inside its groups ends here unless the runs of
//            a group arrived out of order.
//   order    marked groups only: k_gs_groups lists them by size (pairs are finished there); k_gs_warp (warp per group,
//            <= 128 records: registers), k_gs_mid (warp per group, <= 1024: shared memory), k_gs_cta (CTA per group,
//            <= 8192) merge the group's ascending pieces (or run a bitonic network over the sort words when there are many)
//            and overwrite the group's final words.
//   Final form = the layout of the packed LSD sort, (group key << ib) | index, plus the group index of every position.
//
// A one-sweep LSD pass is bound by the SM (warp ranking + shared-memory staging) at ~0.15 ms per 20 M records and
// the 53-bit key needs seven of them; this path touches every record three times.  When some group is larger than
// GS_CTA_MAX (configs[4]'s pile) or the table would not fit (n_seq > ~5800) the caller uses the LSD sort instead.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace swg {

constexpr u32 GS_WARP_MAX = 128;        // largest group one warp orders in registers (4 words per lane)
constexpr u32 GS_MID_MAX = 1024;        // largest group one warp orders in its 8 KB of shared memory
constexpr u32 GS_CTA_MAX = 8192;        // largest group one CTA orders in shared memory (64 KB)
constexpr u64 GS_MAX_TABLE = 1ull << 26; // table entries (256 MB)
constexpr int GS_CTA_THREADS = 256;

// Table entry of a group key, and back.  dense: the key is the bit-packed ((query << sb | target) << 1 | strand) of the chaining
// sort and the table is indexed by (query * n_seq + target) * 2 + strand (2 * n_seq^2 entries instead of 2^(2 sb + 1)); otherwise
// the key itself indexes the table (the sweeps' (sequence, partner genome) keys: a few hundred thousand entries).
struct GsIndex {
    int dense, sb;
    u32 n_seq;
    __device__ __forceinline__ u32 operator()(u32 grp) const {
        if (!dense) return grp;
        return ((grp >> (sb + 1)) * n_seq + ((grp >> 1) & ((1u << sb) - 1))) * 2 + (grp & 1);
    }
    __device__ __forceinline__ u32 key_of(u32 entry) const {
        if (!dense) return entry;
        const u32 pair = entry >> 1, q = pair / n_seq, t = pair - q * n_seq;
        return (((q << sb) | t) << 1) | (entry & 1);
    }
};

// ---- count: slot ranges of the runs ------------------------------------------------------------------------------------------
// keys[i] = (group key << shift) | query_start; excluded records carry the group key `dead` (their runs are skipped).
// A warp owns GS_RUN_CHUNK consecutive runs and walks them 32 at a time, in order.  Lanes whose runs belong to one group (a '+'
// group interrupted by records of other groups) issue ONE atomic and split the range in lane order, and the next 32 runs are
// taken up only after that atomic has returned: the runs of a group that lie inside one chunk get their ranges in input order.
// Runs of one group in different chunks race; k_gs_emit finds the groups this disorders and the ordering kernels repair them.
constexpr u32 GS_RUN_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_gs_run_base(const u32 *__restrict__ n_runs_ptr, const u32 *__restrict__ run_start, u32 n,
                                                     const u64 *__restrict__ keys, int shift, GsIndex ix, u32 dead, u32 *__restrict__ table,
                                                     u32 *__restrict__ run_base) {
    const u32 n_runs = *n_runs_ptr;
    const u32 lane = threadIdx.x & 31, full = 0xFFFFFFFFu;
    const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
    const u32 n_chunks = (n_runs + GS_RUN_CHUNK - 1) / GS_RUN_CHUNK;
    for (u32 chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += n_warps) {
        const u32 c0 = chunk * GS_RUN_CHUNK, c1 = min(c0 + GS_RUN_CHUNK, n_runs);
        // this lane's run of the first round; the loads of the next round are issued before the atomic of the current one
        u32 r = c0 + lane;
        u32 i0 = r < c1 ? run_start[r] : 0, i1 = r < c1 ? (r + 1 < n_runs ? run_start[r + 1] : n) : 0;
        u64 k0 = r < c1 ? keys[i0] : 0;
        for (u32 r0 = c0; r0 < c1; r0 += 32) {
            const u32 rn = r0 + 32 + lane;
            const u32 ni0 = rn < c1 ? run_start[rn] : 0, ni1 = rn < c1 ? (rn + 1 < n_runs ? run_start[rn + 1] : n) : 0;
            const u64 nk0 = rn < c1 ? keys[ni0] : 0;
            u32 gi = NONE32, len = 0;
            if (r < c1) {
                const u32 grp = (u32)(k0 >> shift);
                if (grp != dead) { gi = ix(grp); len = i1 - i0; }
            }
            const u32 peers = __match_any_sync(full, gi);
            u32 before = 0, total = len;
            if (__any_sync(full, peers != (1u << lane))) { // some group has several runs among these 32
                total = 0;
#pragma unroll 8
                for (u32 l = 0; l < 32; l++) {
                    const u32 v = __shfl_sync(full, len, l);
                    if ((peers >> l) & 1) { total += v; if (l < lane) before += v; }
                }
            }
            const u32 leader = (u32)__ffs(peers) - 1;
            u32 base = 0;
            if (lane == leader && gi != NONE32) base = atomicAdd(&table[gi], total);
            base = __shfl_sync(full, base, leader); // (waits for the atomic: the next round's atomics are issued after it)
            if (gi != NONE32) run_base[r] = base + before;
            r = rn; i0 = ni0; i1 = ni1; k0 = nk0;
        }
    }
}

// ---- scan of the table ---------------------------------------------------------------------------------------------------
// One pass, decoupled look-back over tiles (the scheme of scan.cuh) on a packed pair (non-empty groups before << 31 | records
// before).  Non-empty entry g: table[g] = start + 1 (0 stays "empty"), dtab[g] = dense, gstart[dense] = start, gkey[dense] = bit-packed key.
// out[0] = number of groups, out[1] = number of records, out[2] = largest group; gstart[n_groups] = number of records.
__global__ void __launch_bounds__(SC_THREADS) k_gs_scan(u32 *__restrict__ table, u32 *__restrict__ dtab, u32 n_entries, GsIndex ix,
                                                        u32 *__restrict__ gstart, u32 *__restrict__ gkey, u64 *status, u32 *tile_counter, u32 *out) {
    __shared__ u64 ws[SC_THREADS / 32];
    __shared__ u32 s_tile, s_max;
    __shared__ u64 s_excl;
    if (threadIdx.x == 0) { s_tile = atomicAdd(tile_counter, 1u); s_max = 0; }
    __syncthreads();
    const u32 tile = s_tile;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 base = tile * SC_TILE + threadIdx.x * SC_ITEMS;
    u32 v[SC_ITEMS];
    u64 s = 0;
    u32 mx = 0;
    if (base + SC_ITEMS <= n_entries) { // n_entries is even and the tile start a multiple of 8: two 16 B loads
        const uint4 a = *reinterpret_cast<const uint4 *>(table + base), b = *reinterpret_cast<const uint4 *>(table + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SC_ITEMS; k++) v[k] = base + k < n_entries ? table[base + k] : 0;
    }
    static_assert(SC_ITEMS == 8, "k_gs_scan loads eight entries per thread");
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        s += (u64)v[k] + (v[k] ? (1ull << 31) : 0);
        mx = max(mx, v[k]);
    }
    // block-wide exclusive scan of s
    u64 x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (u32)o) x += t;
    }
    if (lane == 31) ws[warp] = x;
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if (lane == 0 && mx) atomicMax(&s_max, mx);
    __syncthreads();
    u64 wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
        const u64 t = ws[w];
        if (w < (int)warp) wbase += t;
        tot += t;
    }
    const u64 ex_local = wbase + x - s;
    if (threadIdx.x < 32) {
        if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? SC_FLAG_INCL : SC_FLAG_AGG) | tot);
        u64 excl = 0;
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
                u64 part = lane <= upto ? (w & SC_VAL_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                excl += part;
                if (incl) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_INCL | (excl + tot));
        }
        if (lane == 0) {
            s_excl = excl;
            if (s_max) atomicMax(&out[2], s_max);
            if ((u64)(tile + 1) * SC_TILE >= n_entries) {
                const u64 all = excl + tot;
                const u32 ng = (u32)(all >> 31), nr = (u32)(all & 0x7FFFFFFFu);
                out[0] = ng;
                out[1] = nr;
                gstart[ng] = nr;
            }
        }
    }
    __syncthreads();
    u64 ex = s_excl + ex_local;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        if (v[k]) {
            const u32 g = base + k;
            const u32 start = (u32)(ex & 0x7FFFFFFFu), dense = (u32)(ex >> 31);
            table[g] = start + 1;
            dtab[g] = dense;
            gstart[dense] = start;
            gkey[dense] = ix.key_of(g);
            ex += (u64)v[k] + (1ull << 31);
        }
    }
}

// ---- scatter ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gs_scatter(const u64 *__restrict__ keys, const u32 *__restrict__ run_of, const u32 *__restrict__ run_start,
                                                    const u32 *__restrict__ run_base, u32 n, int shift, GsIndex ix, u32 dead, int ib,
                                                    const u32 *__restrict__ table, const u32 *__restrict__ dtab,
                                                    u64 *__restrict__ words, u32 *__restrict__ gid) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = keys[i];
    const u32 grp = (u32)(k >> shift);
    if (grp == dead) return;
    const u32 qs = (u32)(k & ((1ull << shift) - 1));
    const u32 r = run_of[i];
    const u32 gi = ix(grp);
    const u32 pos = table[gi] - 1 + run_base[r] + (i - run_start[r]);
    words[pos] = ((u64)qs << ib) | i;
    gid[pos] = dtab[gi];
}

// ---- emit: final form of every position, groups that are not in order ------------------------------------------------------
__global__ void __launch_bounds__(256) k_gs_emit(const u64 *__restrict__ words, const u32 *__restrict__ gid, const u32 *__restrict__ gkey,
                                                 const u32 *__restrict__ n_rec_ptr, int ib, u64 *__restrict__ out, u8 *__restrict__ unsorted,
                                                 u32 *__restrict__ sec /* optional: the secondary key of every position */) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= *n_rec_ptr) return;
    const u64 w = words[p];
    const u32 d = gid[p];
    out[p] = ((u64)gkey[d] << ib) | (w & ((1ull << ib) - 1));
    if (sec) sec[p] = (u32)(w >> ib);
    if (p > 0 && gid[p - 1] == d && words[p - 1] > w) unsorted[d] = 1;
}

// ---- order inside the groups --------------------------------------------------------------------------------------------
// thread per group: lists the groups k_gs_emit marked by size (pairs are finished here)
__global__ void __launch_bounds__(256) k_gs_groups(u32 n_groups, const u32 *__restrict__ gstart, const u32 *__restrict__ gkey,
                                                   int ib, const u8 *__restrict__ unsorted, const u64 *__restrict__ words, u64 *__restrict__ out,
                                                   u32 *__restrict__ sec,
                                                   u32 *__restrict__ list_warp, u32 *__restrict__ list_mid, u32 *__restrict__ list_cta,
                                                   u32 *__restrict__ list_ctr /*[3]*/) {
    const u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    u32 cls = 0; // 1: warp list, 2: mid list, 3: CTA list
    if (d < n_groups) {
        const u32 grp = gkey[d];
        if (unsorted[d]) {
            const u32 s = gstart[d], size = gstart[d + 1] - s;
            if (size == 2) {
                const u64 hi = (u64)grp << ib, mask = (1ull << ib) - 1;
                u64 a = words[s], b = words[s + 1];
                if (b < a) { const u64 t = a; a = b; b = t; }
                out[s] = hi | (a & mask);
                out[s + 1] = hi | (b & mask);
                if (sec) { sec[s] = (u32)(a >> ib); sec[s + 1] = (u32)(b >> ib); }
            } else cls = size <= GS_WARP_MAX ? 1 : size <= GS_MID_MAX ? 2 : 3;
        }
    }
    const u32 full = 0xFFFFFFFFu, lt = lanemask_lt();
#pragma unroll
    for (u32 c = 1; c <= 3; c++) {
        const u32 m = __ballot_sync(full, cls == c);
        if (m == 0) continue;
        u32 base = 0;
        if (lane_id() == (u32)__ffs(m) - 1) base = atomicAdd(&list_ctr[c - 1], (u32)__popc(m));
        base = __shfl_sync(full, base, __ffs(m) - 1);
        if (cls == c) (c == 1 ? list_warp : c == 2 ? list_mid : list_cta)[base + __popc(m & lt)] = d;
    }
}

// A group that is not in order usually is a few ascending pieces: its runs took their slot ranges out of input order, or an
// ordered block is followed by some stray records.  Such a group is MERGED: the final position of a word is its offset inside its
// own piece plus, for every other piece, the number of words below it there (binary search; one comparison when the piece lies
// wholly below or above the word).  The sort words are distinct (they end in the record index), so the positions are a permutation.
constexpr u32 GS_MERGE_MAX = 16; // pieces; beyond that the sorting network is cheaper
// One warp: finds the pieces of sm[0, size) and writes their starts to pstart[0 .. n], pstart[n] = size.  Returns n, or 0 if
// there are more than GS_MERGE_MAX.
__device__ __forceinline__ u32 gs_find_pieces_warp(const u64 *sm, u32 size, u32 *pstart, u32 lane) {
    const u32 full = 0xFFFFFFFFu;
    u32 n_pieces = 0;
    for (u32 base = 0; base < size; base += 32) {
        const u32 p = base + lane;
        const bool head = p < size && (p == 0 || sm[p - 1] > sm[p]);
        const u32 m = __ballot_sync(full, head);
        const u32 cnt = (u32)__popc(m);
        if (n_pieces + cnt > GS_MERGE_MAX) return 0;
        if (head) pstart[n_pieces + __popc(m & ((1u << lane) - 1))] = p;
        n_pieces += cnt;
    }
    if (lane == 0) pstart[n_pieces] = size;
    return n_pieces;
}
// The threads tid, tid + nthreads, ... of a warp or CTA (after a barrier that makes pstart visible) write the merged group.
__device__ __forceinline__ void gs_merge_pieces(const u64 *sm, u32 size, const u32 *pstart, u32 n_pieces, u64 *__restrict__ dst, u64 hi,
                                                u64 mask, u32 tid, u32 nthreads, u32 *__restrict__ sec /* or NULL */, int ib) {
    for (u32 e = tid; e < size; e += nthreads) {
        const u64 x = sm[e];
        u32 rank = 0;
        for (u32 j = 0; j < n_pieces; j++) {
            const u32 a = pstart[j], b = pstart[j + 1];
            if (e >= a && e < b) rank += e - a;
            else if (sm[b - 1] < x) rank += b - a;
            else if (sm[a] < x) {
                u32 lo = a + 1, up = b - 1; // sm[a] < x < sm[b - 1]
                while (lo < up) {
                    const u32 mid = (lo + up) >> 1;
                    if (sm[mid] < x) lo = mid + 1; else up = mid;
                }
                rank += lo - a;
            }
        }
        dst[rank] = hi | (x & mask);
        if (sec) sec[rank] = (u32)(x >> ib);
    }
}

// compare-exchange network over 32 * R words, word index = r * 32 + lane, ascending
template <int R> __device__ __forceinline__ void gs_bitonic_warp(u64 (&x)[R], u32 lane) {
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if ((r & jr) == 0) {
                        const bool up = ((r * 32) & k) == 0; // k >= 64: bit k of the index lies in r
                        const u64 a = x[r], b = x[r | jr];
                        const bool sw = up ? a > b : a < b;
                        x[r] = sw ? b : a;
                        x[r | jr] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const u64 o = __shfl_xor_sync(0xFFFFFFFFu, x[r], j);
                    const bool up = (((u32)(r * 32) | lane) & (u32)k) == 0;
                    const bool lower = (lane & (u32)j) == 0;
                    x[r] = (up == lower) ? (x[r] < o ? x[r] : o) : (x[r] > o ? x[r] : o);
                }
            }
        }
    }
}
template <int R>
__device__ __forceinline__ void gs_warp_group(const u64 *__restrict__ words, u64 *__restrict__ out, u32 s, u32 size, u64 hi, u64 mask, u32 lane,
                                              u64 *sm /*[32 * R], this warp's*/, u32 *pst, u32 *__restrict__ sec, int ib) {
    u64 x[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        x[r] = (u32)(r * 32) + lane < size ? words[s + r * 32 + lane] : NONE64;
        sm[r * 32 + lane] = x[r];
    }
    __syncwarp();
    // a few ascending pieces (the usual shape: an ordered block and some stray records): merge them; else the network
    const u32 n_pieces = gs_find_pieces_warp(sm, size, pst, lane);
    __syncwarp();
    if (n_pieces) {
        gs_merge_pieces(sm, size, pst, n_pieces, out + s, hi, mask, lane, 32, sec ? sec + s : nullptr, ib);
        __syncwarp();
        return;
    }
    gs_bitonic_warp<R>(x, lane);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const u32 p = (u32)(r * 32) + lane;
        if (p < size) {
            out[s + p] = hi | (x[r] & mask);
            if (sec) sec[s + p] = (u32)(x[r] >> ib);
        }
    }
}
// warp per listed group (3 .. GS_WARP_MAX records), warps take groups from a counter
__global__ void __launch_bounds__(256) k_gs_warp(const u32 *__restrict__ list, const u32 *__restrict__ n_list_ptr, u32 *work_ctr,
                                                 const u32 *__restrict__ gstart, const u32 *__restrict__ gkey, int ib, const u64 *__restrict__ words,
                                                 u64 *__restrict__ out, u32 *__restrict__ sec) {
    __shared__ u64 s_sm[8][GS_WARP_MAX];
    __shared__ u32 s_pst[8][GS_MERGE_MAX + 2];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n_list = *n_list_ptr;
    const u64 mask = (1ull << ib) - 1;
    // static assignment (warp w takes entries w, w + #warps, ...): groups of one size class cost about the same, and a work counter
    // would be one same-address atomic per group (~1 ns each, serialised: 0.2 ms for the 190 k groups of a primary sweep's sort)
    const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_list; w += n_warps) {
        const u32 d = list[w];
        const u32 s = gstart[d], size = gstart[d + 1] - s;
        const u64 hi = (u64)gkey[d] << ib;
        if (size <= 32) gs_warp_group<1>(words, out, s, size, hi, mask, lane, s_sm[warp], s_pst[warp], sec, ib);
        else if (size <= 64) gs_warp_group<2>(words, out, s, size, hi, mask, lane, s_sm[warp], s_pst[warp], sec, ib);
        else gs_warp_group<4>(words, out, s, size, hi, mask, lane, s_sm[warp], s_pst[warp], sec, ib);
    }
}
static_assert(GS_WARP_MAX == 128, "gs_warp_group<4> holds 128 words");

// warp per listed group (GS_WARP_MAX + 1 .. GS_MID_MAX records): the group in the warp's own 8 KB of shared memory; no block
// barrier anywhere, the eight warps of a CTA work on eight groups
__global__ void __launch_bounds__(256) k_gs_mid(const u32 *__restrict__ list, const u32 *__restrict__ n_list_ptr, u32 *work_ctr,
                                                const u32 *__restrict__ gstart, const u32 *__restrict__ gkey, int ib, const u64 *__restrict__ words,
                                                u64 *__restrict__ out, u32 *__restrict__ sec) {
    extern __shared__ __align__(16) unsigned char gs_smem_raw[];
    const u32 lane = threadIdx.x & 31;
    u64 *sm = reinterpret_cast<u64 *>(gs_smem_raw) + (threadIdx.x >> 5) * GS_MID_MAX;
    __shared__ u32 s_pst[8][GS_MERGE_MAX + 2];
    u32 *pst = s_pst[threadIdx.x >> 5];
    const u32 n_list = *n_list_ptr;
    const u64 mask = (1ull << ib) - 1;
    // static assignment (warp w takes entries w, w + #warps, ...): groups of one size class cost about the same, and a work counter
    // would be one same-address atomic per group (~1 ns each, serialised: 0.2 ms for the 190 k groups of a primary sweep's sort)
    const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_list; w += n_warps) {
        const u32 d = list[w];
        const u32 s = gstart[d], size = gstart[d + 1] - s;
        const u64 hi = (u64)gkey[d] << ib;
        u32 np = 256;
        while (np < size) np <<= 1;
        __syncwarp();
        for (u32 p = lane; p < np; p += 32) sm[p] = p < size ? words[s + p] : NONE64;
        __syncwarp();
        const u32 n_pieces = gs_find_pieces_warp(sm, size, pst, lane);
        __syncwarp();
        if (n_pieces) { // a few ascending pieces: merge
            gs_merge_pieces(sm, size, pst, n_pieces, out + s, hi, mask, lane, 32, sec ? sec + s : nullptr, ib);
            continue;
        }
        for (u32 k = 2; k <= np; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = lane; t < np / 2; t += 32) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), o = i | j;
                    const u64 a = sm[i], b = sm[o];
                    const bool up = (i & k) == 0;
                    if (up ? a > b : a < b) { sm[i] = b; sm[o] = a; }
                }
                __syncwarp();
            }
        }
        for (u32 p = lane; p < size; p += 32) {
            out[s + p] = hi | (sm[p] & mask);
            if (sec) sec[s + p] = (u32)(sm[p] >> ib);
        }
    }
}

// CTA per listed group (GS_MID_MAX + 1 .. GS_CTA_MAX records): the group in shared memory, in-order check, bitonic network
__global__ void __launch_bounds__(GS_CTA_THREADS) k_gs_cta(const u32 *__restrict__ list, const u32 *__restrict__ n_list_ptr, u32 *work_ctr,
                                                           const u32 *__restrict__ gstart, const u32 *__restrict__ gkey, int ib,
                                                           const u64 *__restrict__ words, u64 *__restrict__ out, u32 *__restrict__ sec) {
    extern __shared__ __align__(16) unsigned char gs_smem_raw[];
    u64 *sm = reinterpret_cast<u64 *>(gs_smem_raw);
    __shared__ u32 s_w, s_done, s_pst[GS_MERGE_MAX + 2];
    const u32 n_list = *n_list_ptr;
    const u64 mask = (1ull << ib) - 1;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_w = atomicAdd(work_ctr, 1u);
        __syncthreads();
        const u32 w = s_w;
        if (w >= n_list) return;
        const u32 d = list[w];
        const u32 s = gstart[d], size = gstart[d + 1] - s;
        const u64 hi = (u64)gkey[d] << ib;
        u32 np = 256;
        while (np < size) np <<= 1;
        for (u32 p = threadIdx.x; p < np; p += GS_CTA_THREADS) sm[p] = p < size ? words[s + p] : NONE64;
        __syncthreads();
        if (threadIdx.x < 32) { // a few ascending pieces: merge
            const u32 np_ = gs_find_pieces_warp(sm, size, s_pst, threadIdx.x);
            if (threadIdx.x == 0) s_done = np_;
        }
        __syncthreads();
        if (s_done) {
            gs_merge_pieces(sm, size, s_pst, s_done, out + s, hi, mask, threadIdx.x, GS_CTA_THREADS, sec ? sec + s : nullptr, ib);
            continue;
        }
        for (u32 k = 2; k <= np; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = threadIdx.x; t < np / 2; t += GS_CTA_THREADS) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), o = i | j;
                    const u64 a = sm[i], b = sm[o];
                    const bool up = (i & k) == 0;
                    if (up ? a > b : a < b) { sm[i] = b; sm[o] = a; }
                }
                __syncthreads();
            }
        }
        for (u32 p = threadIdx.x; p < size; p += GS_CTA_THREADS) {
            out[s + p] = hi | (sm[p] & mask);
            if (sec) sec[s + p] = (u32)(sm[p] >> ib);
        }
    }
}

static inline void gs_init_device() {
    SWG_CUDA(cudaFuncSetAttribute(k_gs_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GS_CTA_MAX * sizeof(u64))));
    SWG_CUDA(cudaFuncSetAttribute(k_gs_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * GS_MID_MAX * sizeof(u64))));
}

} // namespace swg
