// learn_acc.cuh -- layout of the learner's exact integer accumulators (include/othello_b200.h:
// othello_learn_accumulate) and their conversion to the statistics the solver reads; shared by learn.cu
// (accumulate, othello_learn_stats) and learn_solve.cu (othello_learn_refit).
#pragma once
#include "common.cuh"

namespace obl {

constexpr int kX = 10;                       // regressors incl. intercept
constexpr int kFp = 10;                      // fp64 sums per shard: Xty[0..8] (Xty[9] is identically 0), sum y^2
constexpr int kPairs = kX * (kX + 1) / 2;    // upper triangle of XtX
constexpr int kFpBase = 56;                  // acc[shard][kFpBase + 2 k], [.. + 1] = high, low word of fp sum k
constexpr double kFixScale = 1099511627776.0;            // 2^40
static_assert(kFpBase >= kPairs && kFpBase + 2 * kFp <= OTHELLO_ACC, "accumulator layout");

__host__ __device__ constexpr int pair_index(int i, int j) { return i * kX - i * (i - 1) / 2 + (j - i); }

// entry k of a shard's statistics row (XtX[10][10], Xty[10], n, sum y^2) from its accumulators: the correctly
// rounded image of the exact integer sum
__device__ __forceinline__ double stat_from_acc(const long long *a, int k)
{
    if (k < kX * kX) {
        const int i = k / kX, j = k % kX;
        return (double)a[i <= j ? pair_index(i, j) : pair_index(j, i)];
    }
    if (k == 110) return (double)a[pair_index(kX - 1, kX - 1)];          // n = sum of intercept * intercept
    const int q = k == 111 ? kFp - 1 : k - kX * kX;                      // sum y^2 | Xty[q]
    if (q == kX - 1 && k != 111) return 0.0;                             // Xty[intercept] cancels exactly
    long long hi = a[kFpBase + 2 * q];
    unsigned long long lo = (unsigned long long)a[kFpBase + 2 * q + 1];
    hi += (long long)(lo >> 32);                                         // normalise: exact integer arithmetic
    lo &= 0xffffffffull;
    return ((double)hi * 4294967296.0 + (double)lo) * (1.0 / kFixScale); // one rounding
}

}  // namespace obl
