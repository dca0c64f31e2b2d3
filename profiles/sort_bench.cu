// sort_bench.cu — tuning aid (NOT product code): times the passes of the record sort (csrc/radix_sort.cuh, rs_sort_packed) alone,
// so that a kernel variant compiles in seconds instead of rebuilding the whole library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I sweepga_b200/csrc [-DSWG_RS_RANK=.. -DSWG_RS_THREADS=.. ...]
//        profiles/sort_bench.cu -o build/sort_bench_<variant>
// Two key distributions over 53 key bits (12 + 12 + 1 + 28, the PanSN layout of DESIGN section 3), n = 20 M, payload = index:
//   random   every bit random (worst case for the warp ranking: 32 distinct digits per warp in every pass)
//   grouped  records arrive grouped by genome pair and ordered by position inside it, like the bench tables: the high-order
//            passes see warps holding one or two digits
// Prints per-pass CUDA-event times and checks the result (stable order of the full keys).
#include "radix_sort.cuh"
#include <vector>
#include <algorithm>
#include <cstring>
using namespace swg;

__global__ void k_make(u64 *k, u32 *v, u32 n, int mode) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 x = (u64)i * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    u64 key;
    if (mode == 0) key = x & ((1ull << 53) - 1);
    else {
        const u32 per = 2497;                 // records per genome pair
        u32 p = i / per, j = i % per;
        u32 hq = p / 89 % 90, ht = p % 89;    // two haplotypes
        u32 c = (u32)((u64)j * 24 / per);     // chromosome, ascending inside the pair
        u32 inter = (x >> 40) % 10 == 0;      // 10 % inter-chromosomal
        u32 c2 = inter ? (u32)((x >> 44) % 24) : c;
        u32 qid = hq * 24 + c, tid = ht * 24 + c2;
        u32 strand = ((x >> 50) & 31) == 0;
        u32 jj = j - (u32)(((u64)c * per + 23) / 24); // position inside the chromosome block
        u32 qs = jj * 2400000u / 105 + (u32)(x & 0xFFFF);
        key = ((u64)qid << 41) | ((u64)tid << 29) | ((u64)strand << 28) | (qs & 0xFFFFFFF);
    }
    k[i] = key;
    v[i] = i;
}
__global__ void k_check(const u64 *w, const u64 *src, u32 n, int ib, int c0, unsigned long long *bad) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 mask = (1ull << ib) - 1;
    const u32 a0 = (u32)(w[i] & mask);
    bool b = a0 >= n || (w[i] >> ib) != (src[a0 < n ? a0 : 0] >> c0);
    if (!b && i + 1 < n) {
        const u32 a1 = (u32)(w[i + 1] & mask);
        if (a1 < n) b = src[a0] > src[a1] || (src[a0] == src[a1] && a0 > a1);
    }
    if (b) atomicAdd(bad, 1ull);
}

int main(int argc, char **argv) {
    const u32 n = argc > 1 ? (u32)atol(argv[1]) : 20000000u;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int kb = 53, ib = bits_for(n - 1);
    cudaStream_t st;
    cudaStreamCreate(&st);
    rs_init_device();
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    u64 *src, *k, *k2; u32 *v, *v2; unsigned long long *bad;
    cudaMalloc(&src, 8ull * n); cudaMalloc(&k, 8ull * n); cudaMalloc(&k2, 8ull * n);
    cudaMalloc(&v, 4ull * n); cudaMalloc(&v2, 4ull * n); cudaMalloc(&bad, 8);
    RadixSortPlan p = rs_plan(n, 0, kb);
    void *tmp; cudaMalloc(&tmp, p.temp_bytes);
    LaunchCounter lc;
    const int NEV = 2 * (RS_MAX_PASSES + 2);
    cudaEvent_t ev[NEV];
    for (auto &e : ev) cudaEventCreate(&e);
    for (int mode = 0; mode < 2; mode++) {
        std::vector<double> per_pass(p.passes, 0.0);
        double whole = 0;
        PackedSort ps;
        for (int r = -1; r < reps; r++) {
            k_make<<<cdiv(n, 256), 256, 0, st>>>(src, v, n, mode);
            cudaMemcpyAsync(k, src, 8ull * n, cudaMemcpyDeviceToDevice, st);
            cudaEventRecord(ev[0], st);
            ps = rs_sort_packed(n, kb, ib, k, k2, v, v2, tmp, p, st, sm, lc, nullptr, nullptr, 0, ev + 2);
            cudaEventRecord(ev[1], st);
            cudaStreamSynchronize(st);
            if (r < 0) continue;
            float ms;
            cudaEventElapsedTime(&ms, ev[0], ev[1]); whole += ms;
            for (int q = 0; q < p.passes; q++) { cudaEventElapsedTime(&ms, ev[2 + 2 * q], ev[3 + 2 * q]); per_pass[q] += ms; }
        }
        cudaMemsetAsync(bad, 0, 8, st);
        k_check<<<cdiv(n, 256), 256, 0, st>>>(ps.packed, src, n, ib, ps.c0, bad);
        unsigned long long hb = 1;
        cudaMemcpyAsync(&hb, bad, 8, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaError_t e = cudaGetLastError();
        printf("%-8s n=%u passes=%d pack_pass=%d whole=%.3f ms ok=%d%s |", mode ? "grouped" : "random", n, p.passes, ps.pack_pass, whole / reps, hb == 0,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        double packed = 0; int np = 0;
        for (int q = 0; q < p.passes; q++) {
            printf(" %.4f", per_pass[q] / reps);
            if (q > ps.pack_pass) { packed += per_pass[q] / reps; np++; }
        }
        if (np) printf(" | packed pass mean %.4f ms = %.0f GB/s", packed / np, 16.0 * n / (packed / np) / 1e6);
        printf("\n");
    }
    return 0;
}
