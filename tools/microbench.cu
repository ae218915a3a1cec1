// tools/microbench.cu -- per-instruction throughput of the integer pipes on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Each kernel runs ITERS rounds of 8 independent dependent-chains of one instruction kind (inline
// PTX so ptxas cannot strength-reduce across kinds); prints lane-ops per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
constexpr int ITERS = 4096;

enum Op { LOP3, SHF, IMAD, IMADWIDE, IMADHI, IADD3, POPC, BREV, PRMT, SEL, MIX_LOP_IMAD, MIX_LOP_WIDE, MIX_LOP_HI,
          MIX_SHF_IMAD, IADD64, MIX_LOP_POPC, FLO, IMADSHL, MIX3, NOPS };
const char *NAMES[] = {"lop3", "shf.l.wrap", "mad.lo.u32", "mad.wide.u32", "mul.hi.u32", "add3(iadd3)", "popc", "brev",
                       "prmt", "selp", "lop3+mad.lo 1:1", "lop3+mad.wide 1:1", "lop3+mul.hi 1:1", "shf+mad.lo 1:1",
                       "add.u64 (2 instr)", "lop3+popc 3:1", "clz(flo)", "shl via mul (imad.shl)", "lop3:imad:wide 2:1:1"};

template <int OP> __global__ void k(uint32_t *out, uint32_t seed)
{
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = threadIdx.x * 2654435761u + c + seed; b[c] = a[c] ^ 0x9E3779B9u; w[c] = a[c]; }
    const uint32_t m1 = seed | 0x0F0F0F0Fu, m2 = seed * 3 + 128;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1)); }
            if (OP == SHF) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[c]) : "r"(b[c])); }
            if (OP == IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m2), "r"(b[c])); }
            if (OP == IMADWIDE) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(b[c]), "r"(m2)); }
            if (OP == IMADHI) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(m2)); }
            if (OP == IADD3) { asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[c]) : "r"(b[c]), "r"(m1)); }
            if (OP == POPC) { asm volatile("popc.b32 %0, %0;" : "+r"(a[c])); a[c] += b[c]; }
            if (OP == BREV) { asm volatile("brev.b32 %0, %0;" : "+r"(a[c])); }
            if (OP == PRMT) { asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(a[c]) : "r"(b[c])); }
            if (OP == SEL) { asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a[c]) : "r"(b[c]), "r"(m1)); }
            if (OP == MIX_LOP_IMAD) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m2), "r"(m1));
            }
            if (OP == MIX_LOP_WIDE) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(b[c]), "r"(m2));
            }
            if (OP == MIX_LOP_HI) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m2));
            }
            if (OP == MIX_SHF_IMAD) {
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[c]) : "r"(m1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m2), "r"(m1));
            }
            if (OP == IADD64) { asm volatile("add.u64 %0, %0, %1;" : "+l"(w[c]) : "l"((uint64_t)b[c] << 20 | m1)); }
            if (OP == MIX_LOP_POPC) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(m2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(a[c]) : "r"(b[c]), "r"(m2));
                uint32_t p; asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(a[c])); b[c] ^= p;
            }
            if (OP == FLO) { asm volatile("clz.b32 %0, %0;" : "+r"(a[c])); a[c] ^= b[c]; }
            if (OP == IMADSHL) { asm volatile("mul.lo.u32 %0, %0, 128;" : "+r"(a[c])); a[c] |= 1; }
            if (OP == MIX3) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m2), "r"(m1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(a[c]) : "r"(b[c]), "r"(m2));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m2));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= a[c] ^ b[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// PTX-level operations per chain per iteration (what we count as "ops")
const int OPS_PER[] = {1, 1, 1, 1, 1, 2, 2, 1, 1, 2, 2, 2, 2, 2, 1, 5, 2, 2, 4};

template <int OP> void run(uint32_t *d, int sms, double ghz)
{
    const int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, r + 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double ops = (double)OPS_PER[OP] * CHAINS * ITERS * blocks * threads;
    printf("%-26s %8.3f ms  %7.2f Tops/s  %6.1f lane-ops/clk/SM (at %.3f GHz)\n", NAMES[OP], best, ops / best / 1e9,
           ops / (best * 1e-3) / (ghz * 1e9) / sms, ghz);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    uint32_t *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    const int s = p.multiProcessorCount;
    run<LOP3>(d, s, ghz); run<SHF>(d, s, ghz); run<IMAD>(d, s, ghz); run<IMADWIDE>(d, s, ghz); run<IMADHI>(d, s, ghz);
    run<IADD3>(d, s, ghz); run<POPC>(d, s, ghz); run<BREV>(d, s, ghz); run<PRMT>(d, s, ghz); run<SEL>(d, s, ghz);
    run<MIX_LOP_IMAD>(d, s, ghz); run<MIX_LOP_WIDE>(d, s, ghz); run<MIX_LOP_HI>(d, s, ghz); run<MIX_SHF_IMAD>(d, s, ghz);
    run<IADD64>(d, s, ghz); run<MIX_LOP_POPC>(d, s, ghz); run<FLO>(d, s, ghz); run<IMADSHL>(d, s, ghz); run<MIX3>(d, s, ghz);
    return 0;
}
