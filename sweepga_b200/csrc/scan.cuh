// scan.cuh — device-wide exclusive prefix sum (u32) with fused producer/consumer functors.
//
// Used for every stream compaction on the path (group lists, chain lists, anchor lists):
// the producer computes the 0/1 (or small count) value of element i on the fly, the consumer
// receives (i, exclusive_prefix, value).  Three launches: per-block reduce, single-block scan
// of the block sums, per-block scan+apply.  Reads the producer's inputs twice; HBM bound.
#pragma once
#include "common.cuh"

namespace swg {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ u32 block_exclusive_scan_256(u32 v, u32 *ws /*[8]*/, u32 &block_total) {
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

template <class In> __global__ void __launch_bounds__(SC_THREADS) sc_reduce_kernel(In in, u32 n, u32 *__restrict__ block_sums) {
    __shared__ u32 ws[8];
    u32 base = blockIdx.x * SC_TILE;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        u32 i = base + k * SC_THREADS + threadIdx.x;
        if (i < n) s += in(i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        block_sums[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_sums in place; writes the grand total to *total
__global__ void __launch_bounds__(SC_THREADS) sc_scan_sums_kernel(u32 *__restrict__ block_sums, u32 nblocks, u32 *__restrict__ total) {
    __shared__ u32 ws[8];
    u32 carry = 0;
    for (u32 base = 0; base < nblocks; base += SC_THREADS) {
        u32 i = base + threadIdx.x;
        u32 v = i < nblocks ? block_sums[i] : 0;
        u32 tot;
        u32 ex = block_exclusive_scan_256(v, ws, tot);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(SC_THREADS) sc_apply_kernel(In in, Out out, u32 n, const u32 *__restrict__ block_sums) {
    __shared__ u32 ws[8];
    // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile
    u32 base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    u32 v[SC_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        u32 i = base + k;
        v[k] = i < n ? in(i) : 0;
        s += v[k];
    }
    u32 tot;
    u32 ex = block_exclusive_scan_256(s, ws, tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        u32 i = base + k;
        if (i < n) out(i, ex, v[k]);
        ex += v[k];
    }
}

// temp: cdiv(n, SC_TILE) + 1 u32 (block sums); total_out: device u32
template <class In, class Out>
static inline void scan_apply(In in, Out out, u32 n, u32 *block_sums, u32 *total_out, cudaStream_t st, LaunchCounter &lc) {
    if (n == 0) {
        SWG_CUDA(cudaMemsetAsync(total_out, 0, sizeof(u32), st));
        return;
    }
    u32 nb = cdiv(n, SC_TILE);
    sc_reduce_kernel<<<nb, SC_THREADS, 0, st>>>(in, n, block_sums);
    sc_scan_sums_kernel<<<1, SC_THREADS, 0, st>>>(block_sums, nb, total_out);
    sc_apply_kernel<<<nb, SC_THREADS, 0, st>>>(in, out, n, block_sums);
    lc.n += 3;
    SWG_CUDA(cudaGetLastError());
}
static inline size_t scan_temp_u32(u32 n) { return (size_t)cdiv(n, SC_TILE) + 1; }

} // namespace swg
