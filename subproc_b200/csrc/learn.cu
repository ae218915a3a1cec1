// learn.cu -- sufficient statistics of the per-phase linear regression, straight from the
// trajectories the playout kernel left in HBM.
//
// Replaces, on the data-parallel learning path, the reference's per-position Redis round trips
// (__update_state_for_a_book / __update_state_map, progress_position_moves_learn.py:37-62) and
// the four pyres fitting jobs that re-read those values (fit_parameter, :160-184): every
// recorded position contributes, for both sides ('O' = Black, 'X' = White, :44-47), the sample
//     x = (mobility, a..h, 1)            counts() features 1..9 (+ intercept column)
//     y = (own - opp final discs) * 0.9 ** (last_turn - turn)               (:40-42,55)
// to the normal equations of its disc-count shard (:112-113).  The 4 x 112 doubles are the only
// thing ranks exchange (one NCCL all-reduce); the 10x10 solves are done by the host.
//
// Reads 16 B per position (coalesced rows of the SoA trajectory); XtX is accumulated in integers
// (exact), Xty / sum y^2 in fp64.
#include "common.cuh"
#include "fastboard.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 256;
constexpr int kX = 10;                       // regressors incl. intercept
constexpr int kPairs = kX * (kX + 1) / 2;    // upper triangle of XtX
constexpr int kF = kX + 2;                   // Xty[10], n, sum y^2
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ constexpr int pair_index(int i, int j) { return i * kX - i * (i - 1) / 2 + (j - i); }

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Every lane keeps the statistics of the shard it is currently seeing in REGISTERS (55 integer
// products on the FMA pipe + 12 fp64 sums per position) and spills them to the CTA's shared-memory
// totals only when its shard changes -- tiles are walked in ply order, so that is ~4 times per lane --
// and once at the end through warp reductions.
struct LaneAcc {
    unsigned xtx[kPairs];
    double f[kF];
    int shard;
};

__device__ __forceinline__ void lane_clear(LaneAcc &a, int shard)
{
#pragma unroll
    for (int p = 0; p < kPairs; p++) a.xtx[p] = 0u;
#pragma unroll
    for (int k = 0; k < kF; k++) a.f[k] = 0.0;
    a.shard = shard;
}

// a single lane hands its partial sums to the CTA totals (rare: shard change inside a lane)
__device__ __forceinline__ void lane_spill(LaneAcc &a, unsigned long long (*s_xtx)[kPairs], double (*s_f)[kF])
{
    if (a.shard >= 0) {
#pragma unroll
        for (int p = 0; p < kPairs; p++)
            if (a.xtx[p]) atomicAdd(&s_xtx[a.shard][p], (unsigned long long)a.xtx[p]);
#pragma unroll
        for (int k = 0; k < kF; k++)
            if (a.f[k] != 0.0) atomicAdd(&s_f[a.shard][k], a.f[k]);
    }
}

__global__ void __launch_bounds__(kThreads) learn_kernel(const u64 *__restrict__ traj_black,
                                                         const u64 *__restrict__ traj_white,
                                                         const int32_t *__restrict__ nplies,
                                                         const u64 *__restrict__ final_black,
                                                         const u64 *__restrict__ final_white, int64_t n_games,
                                                         int64_t stride, int t_max, const double *__restrict__ decay,
                                                         double *__restrict__ stats)
{
    __shared__ unsigned long long s_xtx[OTHELLO_PHASES][kPairs];
    __shared__ double s_f[OTHELLO_PHASES][kF];
    for (int i = threadIdx.x; i < OTHELLO_PHASES * kPairs; i += kThreads) (&s_xtx[0][0])[i] = 0ull;
    for (int i = threadIdx.x; i < OTHELLO_PHASES * kF; i += kThreads) (&s_f[0][0])[i] = 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t tiles_per_row = (n_games + kThreads - 1) / kThreads;
    const int64_t tiles = tiles_per_row * (int64_t)(t_max + 1);
    // contiguous, ply-major ranges of tiles per CTA: the shard changes a handful of times per lane
    const int64_t first = tiles * blockIdx.x / gridDim.x, last = tiles * (blockIdx.x + 1) / gridDim.x;
    LaneAcc acc;
    lane_clear(acc, -1);
    for (int64_t tile = first; tile < last; tile++) {
        const int t = (int)(tile / tiles_per_row);
        const int64_t g = (tile % tiles_per_row) * kThreads + threadIdx.x;
        int len = -1;
        if (g < n_games) len = nplies[g];
        if (t > len || len > t_max) continue;                 // positions 0..nplies are recorded; truncated games are skipped
        const u64 b = traj_black[(int64_t)t * stride + g], w = traj_white[(int64_t)t * stride + g];
        const int shard = phase_row(__popcll(b | w));
        if (shard != acc.shard) { lane_spill(acc, s_xtx, s_f); lane_clear(acc, shard); }
        int xb[kX], xw[kX];
        const u64 br = obf::rev64(b), wr = obf::rev64(w);
        xb[0] = __popcll(obf::legal_moves(b, w, br, wr));
        xw[0] = __popcll(obf::legal_moves(w, b, wr, br));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            xb[1 + k] = __popcll(b & kClassMask[k]);
            xw[1 + k] = __popcll(w & kClassMask[k]);
        }
        xb[9] = xw[9] = 1;
        const int value = __popcll(final_black[g]) - __popcll(final_white[g]);       // value_for_black (:40-42)
        const double y = (double)value * decay[len - t];                             // * l ** turn_left (:55)
        // XtX: both sides at once, exact integer sums
#pragma unroll
        for (int i = 0; i < kX; i++) {
#pragma unroll
            for (int j = i; j < kX; j++) acc.xtx[pair_index(i, j)] += (unsigned)(xb[i] * xb[j] + xw[i] * xw[j]);
        }
        // Xty: White's target is the negative of Black's (value_for_white, :42)
#pragma unroll
        for (int i = 0; i < kX; i++) acc.f[i] += (double)(xb[i] - xw[i]) * y;
        acc.f[kX] += 2.0;
        acc.f[kX + 1] += 2.0 * y * y;
    }
    // end of the CTA's range: reduce over the warp per shard, one atomic per value per warp
#pragma unroll 1
    for (int s = 0; s < OTHELLO_PHASES; s++) {
        const bool in = acc.shard == s;
        if (!__any_sync(kFull, in)) continue;
#pragma unroll
        for (int p = 0; p < kPairs; p++) {
            // 32 lanes x (tiles per CTA) x 2 * 64 * 64 stays far below 2^32
            const unsigned r = __reduce_add_sync(kFull, in ? acc.xtx[p] : 0u);
            if (lane == (p & 31) && r) atomicAdd(&s_xtx[s][p], (unsigned long long)r);
        }
#pragma unroll
        for (int k = 0; k < kF; k++) {
            const double r = warp_sum_f64(in ? acc.f[k] : 0.0);
            if (lane == k && r != 0.0) atomicAdd(&s_f[s][k], r);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OTHELLO_PHASES * OTHELLO_STATS; i += kThreads) {
        const int s = i / OTHELLO_STATS, k = i % OTHELLO_STATS;
        double v;
        if (k < kX * kX) {
            const int a = k / kX, b = k % kX;
            v = (double)s_xtx[s][a <= b ? pair_index(a, b) : pair_index(b, a)];
        } else {
            v = s_f[s][k - kX * kX];
        }
        if (v != 0.0) atomicAdd(&stats[i], v);
    }
}

}  // namespace

extern "C" int othello_learn_accumulate(const uint64_t *traj_black, const uint64_t *traj_white, const int32_t *nplies,
                                        const uint64_t *final_black, const uint64_t *final_white, int64_t n_games,
                                        int64_t stride, int32_t t_max, const double *decay, double *stats, void *stream)
{
    OB_CHECK_ARGS(n_games >= 0 && t_max >= 0 && stats && decay);
    if (n_games == 0) return 0;
    OB_CHECK_ARGS(traj_black && traj_white && nplies && final_black && final_white && stride >= n_games);
    const int64_t tiles = ((n_games + kThreads - 1) / kThreads) * (int64_t)(t_max + 1);
    int dev = 0, sms = 148;
    OB_CUDA(cudaGetDevice(&dev));
    OB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (int64_t)sms * 4;
    const unsigned blocks = (unsigned)(tiles < want ? tiles : want);
    learn_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>((const u64 *)traj_black, (const u64 *)traj_white, nplies,
                                                               (const u64 *)final_black, (const u64 *)final_white,
                                                               n_games, stride, t_max, decay, stats);
    return ob_launch_status();
}
