// radix_sort.cuh — stable LSD radix sort of (u64 key, u32 payload) pairs for sm_100a.
//
// Replaces the reference's per-(query,target) HashMap/IndexMap grouping + sort_by_key
// (src/grouped_mappings.rs:33-39, src/paf_filter.rs:761-777, 1037-1100, 625-634):
// the group id sits in the high key bits, the position in the low bits, so ONE
// sort produces every group contiguous and position-ordered ("segmented" by
// construction), and stability keeps input order for ties exactly like
// Vec::sort_by_key.
//
// Structure (one-sweep): one histogram kernel counts every 8-bit digit of every
// pass in a single read of the keys; each pass is then ONE kernel that ranks a
// 6144-pair tile with warp match_any digit histograms, resolves the tile's
// global digit offsets by decoupled look-back over earlier tiles, stages the
// tile in shared memory in digit order and writes it out coalesced.
// HBM traffic per pass: 12 B read + 12 B write per pair (+ 8 B once for the
// histogram).  Bound: HBM.
#pragma once
#include "common.cuh"

namespace swg {

#ifndef SWG_RS_BITS
#define SWG_RS_BITS 8
#endif
constexpr int RS_BITS = SWG_RS_BITS; // digit width; 9 needs SWG_RS_THREADS >= 512 (one thread per digit in the scans)
constexpr int RS_RADIX = 1 << RS_BITS;
#ifndef SWG_RS_THREADS
#define SWG_RS_THREADS 384
#endif
#ifndef SWG_RS_ITEMS
#define SWG_RS_ITEMS 14
#endif
constexpr int RS_THREADS = SWG_RS_THREADS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = SWG_RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 5376 pairs (384 threads x 14)
static_assert(RS_RADIX <= RS_THREADS, "one thread per digit: the CTA must have at least 2^RS_BITS threads");
constexpr int RS_MAX_PASSES = 8;
#ifndef SWG_RS_LOOKBACK
#define SWG_RS_LOOKBACK 8
#endif
#ifndef SWG_RS_RANK
#define SWG_RS_RANK 2
#endif
constexpr int RS_LOOKBACK = SWG_RS_LOOKBACK; // predecessor tiles inspected per look-back round (independent loads in flight)

constexpr u32 RS_FLAG_AGG = 1u << 30;  // tile aggregate available
constexpr u32 RS_FLAG_INCL = 2u << 30; // inclusive prefix available
constexpr u32 RS_VAL_MASK = (1u << 30) - 1;

// ---- histogram of all passes in one read --------------------------------------------------
// gap_cb != 0: the keys are in the "gap layout" (high field << 32) | (low field < 2^gap_cb) — written before the width of the
// low field was known — and are squeezed to (high << gap_cb) | low on the fly (rs_squeeze); the sort then behaves as if the
// contiguous key had been stored.
__device__ __forceinline__ u64 rs_squeeze(u64 k, int gap_cb) { return ((k >> 32) << gap_cb) | (u64)(u32)k; }
__global__ void __launch_bounds__(512) rs_histogram_kernel(const u64 *__restrict__ keys, u32 n, int begin_bit, int passes,
                                                           u32 *__restrict__ hist /*[passes][256]*/, int gap_cb = 0) {
    __shared__ u32 sh[RS_MAX_PASSES * RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    // 128-bit loads: two keys per thread per step
    const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys);
    u32 n2 = n >> 1;
    const u32 lane = threadIdx.x & 31;
    // warp-uniform trip count (the body uses full-mask warp votes)
    for (u32 wbase = (blockIdx.x * blockDim.x + threadIdx.x) - lane; wbase < n2; wbase += gridDim.x * blockDim.x) {
        const u32 i = wbase + lane;
        const bool ok = i < n2;
        ulonglong2 v = ok ? k2[i] : make_ulonglong2(0, 0);
        if (gap_cb) { v.x = rs_squeeze(v.x, gap_cb); v.y = rs_squeeze(v.y, gap_cb); }
        u64 a = v.x >> begin_bit, b = v.y >> begin_bit;
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; p++) {
            if (p < passes) {
                u32 da = (u32)(a >> (p * RS_BITS)) & (RS_RADIX - 1);
                u32 db = (u32)(b >> (p * RS_BITS)) & (RS_RADIX - 1);
                // high-order digits of grouped input are usually the same across the warp: one atomic for all 64 keys
                const u32 d0 = __shfl_sync(0xFFFFFFFFu, da, 0);
                if (__all_sync(0xFFFFFFFFu, ok && da == d0 && db == d0)) {
                    if (lane == 0) atomicAdd(&sh[p * RS_RADIX + d0], 64u);
                } else if (ok) {
                    if (da == db) atomicAdd(&sh[p * RS_RADIX + da], 2u);
                    else { atomicAdd(&sh[p * RS_RADIX + da], 1u); atomicAdd(&sh[p * RS_RADIX + db], 1u); }
                }
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        u64 a = (gap_cb ? rs_squeeze(keys[n - 1], gap_cb) : keys[n - 1]) >> begin_bit;
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p * RS_RADIX + ((u32)(a >> (p * RS_BITS)) & (RS_RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of each pass's 256 bins, in place (one block per pass, 256 threads)
__global__ void rs_scan_hist_kernel(u32 *__restrict__ hist) {
    __shared__ u32 ws[RS_RADIX / 32];
    u32 *h = hist + blockIdx.x * RS_RADIX;
    u32 v = h[threadIdx.x];
    u32 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if ((threadIdx.x & 31) >= o) x += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    u32 base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += ws[w];
    h[threadIdx.x] = base + x - v;
}

// ---- one pass ---------------------------------------------------------------------------
// Three formats of one element:
//   RS_PAIRS   (u64 key, u32 payload) in, the same out                       12 B read + 12 B written
//   RS_PACK    (u64 key, u32 payload) in, ONE u64 out                        12 B read +  8 B written
//   RS_PACKED  one u64 in, one u64 out                                        8 B read +  8 B written
// An LSD sort never looks at a digit again once its pass is done, and the callers of the packed sort read the payload and
// the HIGH key bits of the result only (group id; positions are re-gathered from the records).  So as soon as the key
// bits still to be sorted plus the payload bits fit one word — packed = ((key >> c0) << ib) | payload after c0 low key
// bits have been consumed — the remaining passes move 8 B per element instead of 12 B.  RS_PACK is the pass that consumes
// key bits [c0 - 8, c0) and writes the packed word (its own digit is gone from the word, so the digit is staged as a byte).
enum { RS_PAIRS = 0, RS_PACK = 1, RS_PACKED = 2 };

template <int MODE> struct RsSmemT {
    u32 warp_hist[RS_WARPS][RS_RADIX]; // 16 KB: per-warp digit counts, later exclusive-over-warps offsets
    u64 stage_k[RS_TILE];              // 48 KB
    u32 stage_v[MODE == RS_PAIRS ? RS_TILE : 1]; // 24 KB (pairs only)
    u8 stage_d[MODE == RS_PACK ? RS_TILE : 4];  // the digit of the pass that packs
    u32 digit_excl[RS_RADIX];          // tile-local exclusive digit offsets
    u32 global_base[RS_RADIX];         // global output index of staged position 0 of each digit run
    u32 wsum[RS_WARPS];
    u32 tile;
};
typedef RsSmemT<RS_PAIRS> RsSmem;

// shift: bit position of this pass's digit in the INPUT word (RS_PACKED: inside the packed word).
// RS_PACK: pack_drop = c0 (low key bits dropped), pack_ib = payload bits.
template <bool FULL, int MODE>
__device__ __forceinline__ void rs_onesweep_tile(RsSmemT<MODE> &s, const u64 *__restrict__ keys_in, u64 *__restrict__ keys_out,
                                                 const u32 *__restrict__ vals_in, u32 *__restrict__ vals_out, u32 n, int shift,
                                                 int pack_drop, int pack_ib, int pack_gap,
                                                 const u32 *__restrict__ digit_base, u32 *tile_status, u32 tile) {
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile_base = (u64)tile * RS_TILE;
    const u64 warp_base = tile_base + (u64)warp * (32 * RS_ITEMS);
    const u32 warp_n = FULL ? 32 * RS_ITEMS : (u32)min((u64)(32 * RS_ITEMS), n > warp_base ? (u64)n - warp_base : (u64)0);

    u64 key[RS_ITEMS];
    u32 val[MODE == RS_PACKED ? 1 : RS_ITEMS];
    u32 rnk[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool ok = FULL || (u32)(i * 32 + lane) < warp_n;
        key[i] = ok ? keys_in[warp_base + i * 32 + lane] : NONE64;
    }
    // warp-level multi-split ranking (stable: items ascending, lanes ascending): for every item the lanes of the warp holding
    // the same digit ("peers"), the lowest of them as leader; the leader bumps the warp's digit count, everyone takes
    // count-before + its rank inside the run.  Three ways to find the peers (SWG_RS_RANK):
    //   0  eight ballots, one per digit bit (pure ALU work: ~58 SASS instructions per item as compiled — round-2 ncu: the
    //      pass was ALU-pipe bound on exactly this)
    //   1  MATCH.ANY
    //   2  shared-memory match masks: MATCH.ALL first (a warp holding ONE digit — every high-order pass over grouped input —
    //      needs nothing else); otherwise every lane ORs its lane bit into mask[digit] (one shared atomic), reads the mask
    //      back, and the leader clears it.  The masks overlay the staging area, which is idle during the ranking and zeroed
    //      at tile start.
    const u32 lt = lanemask_lt();
#if SWG_RS_RANK == 2
    if (FULL) {
        u32 *match = reinterpret_cast<u32 *>(s.stage_k) + warp * RS_RADIX;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const u32 d = (u32)(key[i] >> shift) & (RS_RADIX - 1);
            int uniform;
            __match_all_sync(0xFFFFFFFFu, d, &uniform);
            u32 peers = 0xFFFFFFFFu;
            if (!uniform) {
                atomicOr(&match[d], 1u << lane);
                __syncwarp();
                peers = match[d];
                __syncwarp();
            }
            const u32 leader = (u32)__ffs(peers) - 1;
            u32 prev = 0;
            if (lane == leader) {
                prev = s.warp_hist[warp][d];
                s.warp_hist[warp][d] = prev + (u32)__popc(peers);
                if (!uniform) match[d] = 0;
            }
            prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
            rnk[i] = prev + (u32)__popc(peers & lt);
            __syncwarp();
        }
    } else
#endif
    {
    // pass A: all the peer masks back to back (independent); pass B: the sequential warp-histogram update.
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool ok = FULL || (u32)(i * 32 + lane) < warp_n;
        const u32 d = (u32)(key[i] >> shift) & (RS_RADIX - 1);
        u32 peers = FULL ? 0xFFFFFFFFu : __ballot_sync(0xFFFFFFFFu, ok);
#if SWG_RS_RANK == 1
        peers &= __match_any_sync(0xFFFFFFFFu, d);
#else
#pragma unroll
        for (int b = 0; b < RS_BITS; b++) {
            const bool bit = (d & (1u << b)) != 0;
            const u32 m = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? m : ~m;
        }
#endif
        if (!FULL && !ok) peers = 1u << lane;
        // leader (5 bits) | run length (6 bits) | rank inside the run (5 bits)
        rnk[i] = (u32)(__ffs(peers) - 1) | ((u32)__popc(peers) << 8) | ((u32)__popc(peers & lt) << 16);
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool ok = FULL || (u32)(i * 32 + lane) < warp_n;
        const u32 d = (u32)(key[i] >> shift) & (RS_RADIX - 1);
        const u32 leader = rnk[i] & 31;
        u32 prev = 0;
        if (lane == leader && ok) {
            prev = s.warp_hist[warp][d];
            s.warp_hist[warp][d] = prev + ((rnk[i] >> 8) & 63);
        }
        prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
        rnk[i] = prev + (rnk[i] >> 16);
        __syncwarp();
    }
    }
    // payloads are only needed for staging: issue their loads now so the latency hides behind the scans/barriers
    if (MODE != RS_PACKED) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const bool ok = FULL || (u32)(i * 32 + lane) < warp_n;
            val[MODE == RS_PACKED ? 0 : i] = ok ? vals_in[warp_base + i * 32 + lane] : 0u;
        }
    }
    __syncthreads();

    // per digit: exclusive scan over warps, tile count
    u32 count = 0;
    u32 *my_status = tile_status + (u64)tile * RS_RADIX + tid;
    if (tid < RS_RADIX) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            u32 t = s.warp_hist[w][tid];
            s.warp_hist[w][tid] = count;
            count += t;
        }
        // publish the tile aggregate as early as possible: successors' look-backs can pass over this tile
        st_volatile_u32(my_status, (tile == 0 ? RS_FLAG_INCL : RS_FLAG_AGG) | count);
    }
    // exclusive scan over digits (tile-local run offsets)
    {
        u32 x = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) s.wsum[warp] = x;
        __syncthreads();
        if (tid < RS_RADIX) {
            u32 base = 0;
            for (u32 w = 0; w < warp; w++) base += s.wsum[w];
            s.digit_excl[tid] = base + x - count;
        }
    }
    __syncthreads();

    // stage in digit order (needs only tile-local offsets; hides the latency of the predecessors' publishing)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool ok = FULL || (u32)(i * 32 + lane) < warp_n;
        if (ok) {
            u32 d = (u32)(key[i] >> shift) & (RS_RADIX - 1);
            u32 pos = s.digit_excl[d] + s.warp_hist[warp][d] + rnk[i];
            if (MODE == RS_PAIRS) {
                s.stage_k[pos] = key[i];
                s.stage_v[pos] = val[MODE == RS_PACKED ? 0 : i];
            } else if (MODE == RS_PACK) {
                const u64 kk = pack_gap ? rs_squeeze(key[i], pack_gap) : key[i]; // (digits consumed so far lie below the gap)
                s.stage_k[pos] = ((kk >> pack_drop) << pack_ib) | (u64)val[MODE == RS_PACKED ? 0 : i];
                s.stage_d[pos] = (u8)d;
            } else {
                s.stage_k[pos] = key[i];
            }
        }
    }

    // decoupled look-back, one thread per digit, RS_LOOKBACK predecessors in flight per round
    if (tid < RS_RADIX) {
        u32 excl = 0;
        if (tile != 0) {
            i64 t = (i64)tile - 1;
            bool done = false;
            while (!done) {
                u32 v[RS_LOOKBACK];
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK; k++)
                    v[k] = (t - k >= 0) ? ld_volatile_u32(tile_status + (u64)(t - k) * RS_RADIX + tid) : RS_FLAG_INCL;
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK; k++) {
                    if (done) break;
                    if (v[k] & RS_FLAG_INCL) { excl += v[k] & RS_VAL_MASK; done = true; }
                    else if (v[k] & RS_FLAG_AGG) { excl += v[k] & RS_VAL_MASK; t--; }
                    else break; // not published yet: poll again from this tile
                }
            }
            st_volatile_u32(my_status, RS_FLAG_INCL | (excl + count));
        }
        s.global_base[tid] = digit_base[tid] + excl - s.digit_excl[tid];
    }
    __syncthreads();
    const u32 tile_n = FULL ? RS_TILE : (u32)min((u64)RS_TILE, (u64)n - tile_base);
    if (FULL) {
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const u32 j = it * RS_THREADS + tid;
            const u64 k = s.stage_k[j];
            const u32 d = MODE == RS_PACK ? (u32)s.stage_d[MODE == RS_PACK ? j : 0] : ((u32)(k >> shift) & (RS_RADIX - 1));
            const u32 g = s.global_base[d] + j;
            keys_out[g] = k;
            if (MODE == RS_PAIRS) vals_out[g] = s.stage_v[MODE == RS_PAIRS ? j : 0];
        }
    } else {
        for (u32 j = tid; j < tile_n; j += RS_THREADS) {
            const u64 k = s.stage_k[j];
            const u32 d = MODE == RS_PACK ? (u32)s.stage_d[MODE == RS_PACK ? j : 0] : ((u32)(k >> shift) & (RS_RADIX - 1));
            const u32 g = s.global_base[d] + j;
            keys_out[g] = k;
            if (MODE == RS_PAIRS) vals_out[g] = s.stage_v[MODE == RS_PAIRS ? j : 0];
        }
    }
}

#ifndef SWG_RS_MINBLOCKS
#define SWG_RS_MINBLOCKS 2
#endif
#ifndef SWG_RS_MINBLOCKS_PACKED
#define SWG_RS_MINBLOCKS_PACKED 2
#endif
template <int MODE>
__global__ void __launch_bounds__(RS_THREADS, MODE == RS_PACKED ? SWG_RS_MINBLOCKS_PACKED : SWG_RS_MINBLOCKS)
rs_onesweep_kernel(const u64 *__restrict__ keys_in, u64 *__restrict__ keys_out, const u32 *__restrict__ vals_in,
                   u32 *__restrict__ vals_out, u32 n, int shift, int pack_drop, int pack_ib, int pack_gap, const u32 *__restrict__ digit_base,
                   u32 *tile_status /*[tiles][256]*/, u32 *tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmemT<MODE> &s = *reinterpret_cast<RsSmemT<MODE> *>(smem_raw);
    const u32 tid = threadIdx.x;
    if (tid == 0) s.tile = atomicAdd(tile_counter, 1u); // in-order tile ids: look-back never waits on an unscheduled tile
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s.warp_hist[0][0])[i] = 0;
#if SWG_RS_RANK == 2
    static_assert(sizeof(u64) * RS_TILE >= sizeof(u32) * RS_WARPS * RS_RADIX, "the match masks overlay the staging area");
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) reinterpret_cast<u32 *>(s.stage_k)[i] = 0;
#endif
    __syncthreads();
    const u32 tile = s.tile;
    if ((u64)(tile + 1) * RS_TILE <= (u64)n)
        rs_onesweep_tile<true, MODE>(s, keys_in, keys_out, vals_in, vals_out, n, shift, pack_drop, pack_ib, pack_gap, digit_base, tile_status, tile);
    else
        rs_onesweep_tile<false, MODE>(s, keys_in, keys_out, vals_in, vals_out, n, shift, pack_drop, pack_ib, pack_gap, digit_base, tile_status, tile);
}

// ---- host driver --------------------------------------------------------------------------
// The one-sweep tile needs more dynamic shared memory than the 48 KB default; the opt-in is a per-DEVICE function
// attribute, so every context sets it for its own device (swg_create, after cudaSetDevice).
static inline void rs_init_device() {
    SWG_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<RS_PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmemT<RS_PAIRS>)));
    SWG_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<RS_PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmemT<RS_PACK>)));
    SWG_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<RS_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmemT<RS_PACKED>)));
}

struct RadixSortPlan {
    u32 n = 0;
    int begin_bit = 0, passes = 0;
    u32 tiles = 0;
    size_t temp_bytes = 0; // hist + counters + tile status (zeroed by the driver)
};

static inline RadixSortPlan rs_plan(u32 n, int begin_bit, int end_bit) {
    RadixSortPlan p;
    p.n = n;
    p.begin_bit = begin_bit;
    int bits = end_bit - begin_bit;
    if (bits < 1) bits = 1;
    p.passes = (bits + RS_BITS - 1) / RS_BITS;
    if (p.passes > RS_MAX_PASSES) p.passes = RS_MAX_PASSES;
    p.tiles = cdiv(n, RS_TILE);
    p.temp_bytes = sizeof(u32) * ((size_t)RS_MAX_PASSES * RS_RADIX + 16 + (size_t)p.passes * p.tiles * RS_RADIX);
    return p;
}

// Sorts by key bits [begin_bit, begin_bit + 8*passes).  On return (keys, vals) point at the
// sorted data and (keys_alt, vals_alt) at the scratch (the pointers are swapped as needed).
static inline void rs_sort_pairs(const RadixSortPlan &p, u64 *&keys, u64 *&keys_alt, u32 *&vals, u32 *&vals_alt, void *temp,
                                 cudaStream_t st, int sm_count, LaunchCounter &lc, cudaEvent_t ev_begin = nullptr,
                                 cudaEvent_t ev_end = nullptr) {
    if (p.n == 0) return;
    u32 *hist = (u32 *)temp;
    u32 *counters = hist + RS_MAX_PASSES * RS_RADIX;
    u32 *status = counters + 16;
    SWG_CUDA(cudaMemsetAsync(temp, 0, p.temp_bytes, st));
    u32 hgrid = (u32)min((u64)sm_count * 4, (u64)cdiv(p.n / 2 + 1, 512));
    rs_histogram_kernel<<<hgrid, 512, 0, st>>>(keys, p.n, p.begin_bit, p.passes, hist);
    rs_scan_hist_kernel<<<p.passes, RS_RADIX, 0, st>>>(hist);
    lc.n += 2;
    if (ev_begin) SWG_CUDA(cudaEventRecord(ev_begin, st));
    for (int pass = 0; pass < p.passes; pass++) {
        rs_onesweep_kernel<RS_PAIRS><<<p.tiles, RS_THREADS, sizeof(RsSmem), st>>>(keys, keys_alt, vals, vals_alt, p.n,
                                                                                  p.begin_bit + pass * RS_BITS, 0, 0, 0, hist + pass * RS_RADIX,
                                                                                  status + (size_t)pass * p.tiles * RS_RADIX, counters + pass);
        lc.n += 1;
        u64 *tk = keys; keys = keys_alt; keys_alt = tk;
        u32 *tv = vals; vals = vals_alt; vals_alt = tv;
    }
    if (ev_end) SWG_CUDA(cudaEventRecord(ev_end, st));
    SWG_CUDA(cudaGetLastError());
}

// ---- packed sort ---------------------------------------------------------------------------
// Sort (key, payload) pairs by key bits [0, key_bits) (begin bit 0) where the caller only needs, of the result, the payload
// and the key bits from `c0` upwards: returns ONE array of words ((key >> c0) << ib) | payload in sorted order.
// c0 = the smallest multiple of the digit width with key_bits - c0 + ib <= 64 (0: the words are packed by the first pass).
struct PackedSort {
    const u64 *packed = nullptr; // sorted words
    int c0 = 0, ib = 0;          // payload = word & ((1 << ib) - 1); key >> c0 = word >> ib
    int passes = 0, pack_pass = -1;
    int timed_passes = 0, timed_bytes_per_pair = 16; // what [ev_begin, ev_end] bracket: the packed-word passes (all passes if there is none)
};
static inline int rs_packed_c0(int key_bits, int ib) {
    int c0 = 0;
    while (key_bits - c0 + ib > 64) c0 += RS_BITS;
    return c0;
}
// keys/keys_alt: n u64 each; vals/vals_alt: n u32 each (vals_alt is unused when the first pass packs).  Needs
// payload < 2^ib and key_bits <= RS_MAX_PASSES * RS_BITS.  ev[2k], ev[2k+1] (optional): events around pass k.
static inline PackedSort rs_sort_packed(u32 n, int key_bits, int ib, u64 *keys, u64 *keys_alt, u32 *vals, u32 *vals_alt, void *temp,
                                        const RadixSortPlan &p, cudaStream_t st, int sm_count, LaunchCounter &lc,
                                        cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr, int gap_cb = 0,
                                        cudaEvent_t *pass_ev = nullptr /* tuning: events around every pass */) {
    PackedSort r;
    r.ib = ib;
    r.c0 = rs_packed_c0(key_bits, ib);
    r.passes = p.passes;
    if (n == 0) return r;
    u32 *hist = (u32 *)temp;
    u32 *counters = hist + RS_MAX_PASSES * RS_RADIX;
    u32 *status = counters + 16;
    SWG_CUDA(cudaMemsetAsync(temp, 0, p.temp_bytes, st));
    u32 hgrid = (u32)min((u64)sm_count * 4, (u64)cdiv(p.n / 2 + 1, 512));
    rs_histogram_kernel<<<hgrid, 512, 0, st>>>(keys, p.n, 0, p.passes, hist, gap_cb);
    rs_scan_hist_kernel<<<p.passes, RS_RADIX, 0, st>>>(hist);
    lc.n += 2;
    // passes [0, pk) move pairs, pass pk packs, the rest move packed words.  c0 == 0: a pass "-1" would pack, so the first
    // pass packs without dropping anything and keeps its own digit in the word (read back from it at write-out: RS_PACK
    // stages the digit byte either way).
    const int pk = r.c0 == 0 ? 0 : r.c0 / RS_BITS - 1;
    r.pack_pass = pk;
    r.timed_passes = p.passes - pk - 1;
    if (r.timed_passes <= 0) { r.timed_passes = p.passes; r.timed_bytes_per_pair = pk == 0 ? 20 : 24; }
    const int first_timed = r.timed_passes == p.passes ? 0 : pk + 1;
    for (int pass = 0; pass < p.passes; pass++) {
        if (pass == first_timed && ev_begin) SWG_CUDA(cudaEventRecord(ev_begin, st));
        if (pass_ev) SWG_CUDA(cudaEventRecord(pass_ev[2 * pass], st));
        const u32 *db = hist + pass * RS_RADIX;
        u32 *stt = status + (size_t)pass * p.tiles * RS_RADIX;
        if (pass < pk) {
            rs_onesweep_kernel<RS_PAIRS><<<p.tiles, RS_THREADS, sizeof(RsSmemT<RS_PAIRS>), st>>>(keys, keys_alt, vals, vals_alt, n, pass * RS_BITS, 0, 0, 0,
                                                                                              db, stt, counters + pass);
            u32 *tv = vals; vals = vals_alt; vals_alt = tv;
        } else if (pass == pk) {
            rs_onesweep_kernel<RS_PACK><<<p.tiles, RS_THREADS, sizeof(RsSmemT<RS_PACK>), st>>>(keys, keys_alt, vals, nullptr, n, pass * RS_BITS, r.c0, ib,
                                                                                            gap_cb, db, stt, counters + pass);
        } else {
            rs_onesweep_kernel<RS_PACKED><<<p.tiles, RS_THREADS, sizeof(RsSmemT<RS_PACKED>), st>>>(keys, keys_alt, nullptr, nullptr, n,
                                                                                                ib + pass * RS_BITS - r.c0, 0, 0, 0, db, stt, counters + pass);
        }
        lc.n += 1;
        if (pass_ev) SWG_CUDA(cudaEventRecord(pass_ev[2 * pass + 1], st));
        u64 *tk = keys; keys = keys_alt; keys_alt = tk;
    }
    if (ev_end) SWG_CUDA(cudaEventRecord(ev_end, st));
    SWG_CUDA(cudaGetLastError());
    r.packed = keys;
    return r;
}

} // namespace swg
