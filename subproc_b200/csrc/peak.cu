// peak.cu -- INT32 ALU-pipe micro-benchmark: the measured denominator of the integer roofline.
//
// MEASURED_PEAKS.json carries HBM and tensor peaks only; the step kernels are bound by the
// integer ALU pipe (LOP3 / SHF / IADD3), so bench.py measures that peak on the box it runs on.
// Every thread advances 8 independent chains; one round of one chain is 2 funnel shifts (SHF)
// and 2 three-input logic ops (LOP3) = 4 ALU-pipe instructions, i.e. 32 lane-ops per thread
// per round.  cuobjdump -sass shows the loop body 1:1 (profiles/sass_counts_r01.txt).
#include "common.cuh"

namespace {

constexpr int kChains = 8;

__global__ void int32_peak_kernel(uint32_t *sink, int iters)
{
    uint32_t a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) {
        a[c] = threadIdx.x * 2654435761u + c;
        b[c] = blockIdx.x * 40503u + 7u * c + 1u;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < kChains; c++) {
            const uint32_t r = __funnelshift_l(a[c], b[c], 7);          // SHF
            a[c] = (a[c] & r) ^ b[c];                                    // LOP3
            const uint32_t q = __funnelshift_r(b[c], a[c], 9);          // SHF
            b[c] = (b[c] | q) ^ a[c];                                    // LOP3
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) acc ^= a[c] ^ b[c];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// Both integer-capable pipes at once: per chain and round 2 LOP3 (ALU pipe) + 2 IMAD (FMA pipe), the
// multiplier coming from a kernel argument so that ptxas cannot strength-reduce it.  This is the
// ceiling for code that balances its integer work over the two pipes (what fastboard.cuh aims at).
__global__ void int32_dual_peak_kernel(uint32_t *sink, int iters, uint32_t mul)
{
    uint32_t a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) {
        a[c] = threadIdx.x * 2654435761u + c;
        b[c] = blockIdx.x * 40503u + 7u * c + 1u;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < kChains; c++) {
            a[c] = (a[c] & b[c]) ^ 0x9E3779B9u;                           // LOP3
            b[c] = b[c] * mul + a[c];                                     // IMAD
            a[c] = (a[c] | b[c]) ^ 0x7F4A7C15u;                           // LOP3
            b[c] = b[c] * mul + a[c];                                     // IMAD
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) acc ^= a[c] ^ b[c];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace

extern "C" int othello_int32_dual_peak_kernel(uint32_t *sink, int blocks, int threads, int iters, void *stream)
{
    OB_CHECK_ARGS(sink && blocks > 0 && threads > 0 && threads <= 1024 && iters >= 0);
    int32_dual_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters, 2654435769u);
    return ob_launch_status();
}

extern "C" int othello_int32_peak_kernel(uint32_t *sink, int blocks, int threads, int iters, void *stream)
{
    OB_CHECK_ARGS(sink && blocks > 0 && threads > 0 && threads <= 1024 && iters >= 0);
    int32_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
    return ob_launch_status();
}
