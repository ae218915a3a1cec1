// common.cuh -- launch helpers shared by the translation units of libothello_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/othello_b200.h"
#include "bitboard.cuh"

#define OB_CHECK_ARGS(cond) do { if (!(cond)) return OTHELLO_E_INVALID; } while (0)

// Surface the launch error (if any) of the kernel just enqueued, as the C ABI's return code.
static inline int ob_launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

#define OB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return (int)e_; } while (0)

static inline unsigned ob_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
