// cub_yardstick.cu — YARDSTICK ONLY, never linked into the product: times cub::DeviceRadixSort on the same shapes as the
// record sort of the filter (20 M pairs, 53 key bits), so that rs_onesweep_kernel has a number to be compared with on the same box.
//   (a) SortPairs<u64, u32>(begin_bit 0, end_bit 53)         — the pairs sort the product used in round 1
//   (b) SortKeys<u64>(begin_bit 25, end_bit 62)              — what the packed sort moves after packing (8 B words)
// Prints ms per sort and the per-pass equivalent (CUB's one-sweep uses 8-bit digits: ceil(bits / 8) passes).
// Build + run on the GPU box: nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/cub_yardstick.cu -o /tmp/cuby && /tmp/cuby
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void fill(uint64_t *k, uint32_t *v, uint32_t n, int bits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    k[i] = bits >= 64 ? x : (x & ((1ull << bits) - 1));
    v[i] = i;
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 20000000u;
    const int reps = 10;
    uint64_t *k0, *k1, *src;
    uint32_t *v0, *v1;
    CK(cudaMalloc(&src, n * 8ull)); CK(cudaMalloc(&k0, n * 8ull)); CK(cudaMalloc(&k1, n * 8ull));
    CK(cudaMalloc(&v0, n * 4ull)); CK(cudaMalloc(&v1, n * 4ull));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; mode++) {
        const int bits = mode == 0 ? 53 : 62, begin = mode == 0 ? 0 : 25;
        fill<<<(n + 255) / 256, 256>>>(src, v0, n, bits);
        size_t tb = 0;
        cub::DoubleBuffer<uint64_t> dk(k0, k1);
        cub::DoubleBuffer<uint32_t> dv(v0, v1);
        if (mode == 0) CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)n, begin, bits));
        else CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, (int)n, begin, bits));
        void *tmp;
        CK(cudaMalloc(&tmp, tb));
        double total = 0;
        for (int r = 0; r < reps + 2; r++) {
            CK(cudaMemcpy(k0, src, n * 8ull, cudaMemcpyDeviceToDevice));
            cub::DoubleBuffer<uint64_t> a(k0, k1);
            cub::DoubleBuffer<uint32_t> b(v0, v1);
            CK(cudaEventRecord(e0));
            if (mode == 0) CK(cub::DeviceRadixSort::SortPairs(tmp, tb, a, b, (int)n, begin, bits));
            else CK(cub::DeviceRadixSort::SortKeys(tmp, tb, a, (int)n, begin, bits));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) total += ms;
        }
        const int passes = (bits - begin + 7) / 8;
        const double ms = total / reps, per = ms / passes, bytes = mode == 0 ? 24.0 : 16.0;
        printf("%s n=%u bits=[%d,%d): %.4f ms per sort (histogram included), %d passes -> %.4f ms per pass = %.0f GB/s algorithmic (%.0f B per element per pass)\n",
               mode == 0 ? "cub::DeviceRadixSort::SortPairs<u64,u32>" : "cub::DeviceRadixSort::SortKeys<u64>", n, begin, bits, ms, passes, per,
               n * bytes / (per / 1e3) / 1e9, bytes);
        CK(cudaFree(tmp));
    }
    return 0;
}
