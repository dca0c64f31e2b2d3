// glibc_log.cuh — ln(x) with the operations of glibc's log() in the same order, so that device scores carry the bits
// the reference's `f64::ln` produces on the host.
//
// Why: the plane sweeps rank by `identity * ln(query span)` (reference src/plane_sweep_exact.rs:68-76) and the chain
// filter compares `sum_matches / (sum_block + ln(gap))` with a threshold (src/paf_filter.rs:902-913).  Rust's f64::ln is
// the platform libm's log.  glibc's log (>= 2.28: Szabolcs Nagy's table-driven routine) is accurate to ~0.52 ulp but not
// correctly rounded, CUDA's log() to 1 ulp: they differ in the last bit for a few inputs in 10^5, enough to flip a
// near-tie.  Porting the routine itself removes the difference instead of bounding it.
//
// What is ported: the main path of the FMA build (`__log_fma`, selected by glibc's ifunc on every x86_64 CPU with
// FMA + AVX2) as found in Ubuntu GLIBC 2.39's libm.so.6: z = mantissa re-centred on [0.6875, 1.375), table lookup
// (1/c, log c) by the top 7 mantissa bits, r = fma(z, 1/c, -1), a degree-5 polynomial in r, result assembled as
// hi + (lo + poly).  Every fma below is a single-rounding fused operation exactly where the compiled routine has one
// (read off its instruction sequence), every other operation is an individually rounded add / multiply.
// Arguments on this path are positive integers (spans, gaps); x == 1 returns +0 like the routine.  Arguments in the
// routine's separate near-1 interval (0.9375 .. 1.0647), subnormal, non-finite or non-positive ones — none of which an
// integer >= 2 can be — go to CUDA's log().
//
// The constants come from the host libm (tools/gen_glibc_log_table.py -> glibc_log_table.h).  swg_create() checks the
// port against the running host's log() on a fixed set of arguments and records the verdict in the context
// (`log_matches_host`): when it holds, device and host scores are the same bits and no ranking can differ; when it does
// not (another libm), the near-tie audit of the sweeps triggers the exact re-rank with host-computed scores.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "glibc_log_table.h"

namespace swg {

struct __align__(16) GlogEntry { double invc, logc; };

#ifdef __CUDACC__
// in global memory, read through L1 as one 16-byte load: the index differs from lane to lane, and a __constant__ table serialises
// a warp's 32 distinct addresses (the score pass of a 20 M table took 0.34 ms with it)
__device__ const GlogEntry d_glog_tab[128] = {SWG_GLOG_TABLE};
#endif
static const GlogEntry h_glog_tab[128] = {SWG_GLOG_TABLE};

__host__ __device__ inline double glibc_log(double x) {
    uint64_t ix;
#ifdef __CUDA_ARCH__
    ix = (uint64_t)__double_as_longlong(x);
#else
    memcpy(&ix, &x, 8);
#endif
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const uint32_t top = (uint32_t)(ix >> 48);
    const bool near1 = ix - 0x3fee000000000000ull < 0x000308ffffffffffull + 1; // [1 - 2^-4, 1 + 0x1.09p-4)
    if (near1 || top - 0x0010u >= 0x7ff0u - 0x0010u) return log(x);           // near 1, subnormal, zero, negative, inf, nan
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    double z;
#ifdef __CUDA_ARCH__
    z = __longlong_as_double((long long)iz);
    const double2 ce = __ldg(reinterpret_cast<const double2 *>(d_glog_tab) + i);
    const double invc = ce.x, logc = ce.y;
#define SWG_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SWG_ADD(a, b) __dadd_rn((a), (b))
#define SWG_MUL(a, b) __dmul_rn((a), (b))
#else
    memcpy(&z, &iz, 8);
    const double invc = h_glog_tab[i].invc, logc = h_glog_tab[i].logc;
#define SWG_FMA(a, b, c) fma((a), (b), (c))
#define SWG_ADD(a, b) ((a) + (b))
#define SWG_MUL(a, b) ((a) * (b))
#endif
    const double kd = (double)k;
    const double w = SWG_FMA(kd, SWG_GLOG_LN2HI, logc);       // kd * Ln2hi + logc
    const double r = SWG_FMA(z, invc, -1.0);                  // z / c - 1
    const double p12 = SWG_FMA(r, SWG_GLOG_A2, SWG_GLOG_A1);  // A1 + r * A2
    const double hi = SWG_ADD(r, w);
    const double r2 = SWG_MUL(r, r);
    double lo = SWG_ADD(SWG_ADD(w, -hi), r);                  // (w - hi) + r
    lo = SWG_FMA(kd, SWG_GLOG_LN2LO, lo);
    const double r3 = SWG_MUL(r, r2);
    const double p34 = SWG_FMA(r, SWG_GLOG_A4, SWG_GLOG_A3);  // A3 + r * A4
    const double t = SWG_FMA(r2, SWG_GLOG_A0, lo);            // lo + r2 * A0
    const double q = SWG_FMA(p34, r2, p12);                   // (A1 + r A2) + r2 (A3 + r A4)
    const double y = SWG_FMA(r3, q, t);
    return SWG_ADD(y, hi);
#undef SWG_FMA
#undef SWG_ADD
#undef SWG_MUL
}

} // namespace swg
