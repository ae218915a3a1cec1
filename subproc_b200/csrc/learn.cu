// learn.cu -- sufficient statistics of the per-phase linear regression, straight from the
// trajectories the game kernels left in HBM.
//
// Replaces, on the data-parallel learning path, the reference's per-position Redis round trips
// (__update_state_for_a_book / __update_state_map, progress_position_moves_learn.py:37-62) and
// the four pyres fitting jobs that re-read those values (fit_parameter, :160-184): every
// recorded position contributes, for both sides ('O' = Black, 'X' = White, :44-47), the sample
//     x = (mobility, a..h, 1)            counts() features 1..9 (+ intercept column)
//     y = (own - opp final discs) * 0.9 ** (last_turn - turn)               (:40-42,55)
// to the normal equations of its disc-count shard (:112-113).
//
// The statistics are accumulated as INTEGERS, so that they are identical -- bit for bit -- however
// the games are split over launches, CTAs, ranks or GPUs, and the ranks' all-reduce is an exact
// integer sum:
//   * X^T X (and the sample count, its intercept x intercept entry) is a Gram matrix of small
//     integers: a genuine dense contraction, so it runs on the tensor cores -- per warp and ply one
//     16 x 16 x 64 int8 product (32 positions x 2 sides; mma.sync m16n8k32 s8, int32 accumulators);
//   * X^T y and sum y^2 are fp64.  One lane owns one game and sums its positions in ply order, per
//     shard and per block of plies -- a fixed order, whatever the launch geometry -- and hands every
//     such partial sum over as a 2^-40 fixed-point integer (two int64 words).
// othello_learn_stats turns the integers into the [4][112] doubles the solver reads.
//
// Work split: a unit of work = 32 games (one per lane) x one block of kPlyBlock = 8 plies; the units are
// dealt round robin to the warps of a persistent grid.  A warp reads coalesced 256-byte rows of the SoA
// trajectory (16 B per position, the next ply's rows are requested before the current ply is worked on).
// The blocks are fixed multiples of 8 plies, so the fp64 partial sums do not depend on the launch geometry.
#include "common.cuh"
#include "fastboard.cuh"
#include "learn_acc.cuh"

using namespace ob;
using namespace obl;

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = 32 * kWarps;
constexpr int kPlyBlock = 8;                 // plies per unit of work (fixed: part of the definition of the fp64 sums)
constexpr int kGames = 32;                   // games per CTA (one per lane)
constexpr unsigned kFull = 0xffffffffu;

// D[16x8] += A[16x32] * B[32x8], signed 8-bit operands, 32-bit accumulators (tensor cores)
__device__ __forceinline__ void mma_s8(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Gram {                                // the warp's 16 x 16 accumulator: columns 0..7 and 8..15
    int lo[4], hi[4];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < 4; i++) lo[i] = hi[i] = 0;
    }
};

// hand the upper triangle of the warp's accumulator to the CTA totals of `shard`; every needed entry
// is held by exactly one lane (fragment layout of m16n8k32: c0/c1 = row lane/4, columns 2 * (lane % 4) + {0, 1};
// c2/c3 = row lane/4 + 8)
__device__ __forceinline__ void gram_flush(const Gram &g, unsigned long long (*s_xtx)[kPairs], int shard, int lane)
{
    const int r = lane >> 2, c = 2 * (lane & 3);
    if (r <= c && g.lo[0]) atomicAdd(&s_xtx[shard][pair_index(r, c)], (unsigned long long)g.lo[0]);
    if (r <= c + 1 && g.lo[1]) atomicAdd(&s_xtx[shard][pair_index(r, c + 1)], (unsigned long long)g.lo[1]);
    if ((lane & 3) == 0) {                   // columns 8 and 9
        if (g.hi[0]) atomicAdd(&s_xtx[shard][pair_index(r, 8)], (unsigned long long)g.hi[0]);
        if (g.hi[1]) atomicAdd(&s_xtx[shard][pair_index(r, 9)], (unsigned long long)g.hi[1]);
        if (r == 0 && g.hi[2]) atomicAdd(&s_xtx[shard][pair_index(8, 8)], (unsigned long long)g.hi[2]);
        if (r <= 1 && g.hi[3]) atomicAdd(&s_xtx[shard][pair_index(8 + r, 9)], (unsigned long long)g.hi[3]);
    }
}

// The lanes' fp64 partial sums of one (game, ply block, shard) -> 2^-40 fixed point, exact integer adds from
// here on.  Every game's integer q is split into its signed high word and unsigned low 32 bits, and the two
// are summed separately all the way (warp, CTA, grid, ranks): the pair of totals is then the same whatever
// the grouping.  The 32 integers of a warp are added through the warp-reduce instruction in 16-bit limbs (no
// carries to lose), then lane k adds value k to the CTA totals: two shared atomics per value per warp.
__device__ __forceinline__ void fp_flush(const double (&f)[kFp], bool in, unsigned long long (*s_fp)[2 * kFp], int shard, int lane)
{
    long long hi = 0;
    unsigned long long lo = 0;
#pragma unroll
    for (int k = 0; k < kFp; k++) {
        const long long q = in ? __double2ll_rn(f[k] * kFixScale) : 0ll;              // |q| < 2^60
        const unsigned l0 = __reduce_add_sync(kFull, (unsigned)(q & 0xffff));
        const unsigned l1 = __reduce_add_sync(kFull, (unsigned)((q >> 16) & 0xffff));
        const unsigned l2 = __reduce_add_sync(kFull, (unsigned)((q >> 32) & 0xffff));
        const int l3 = __reduce_add_sync(kFull, (int)(q >> 48));                       // signed top limb
        if (lane == k) {
            lo = (unsigned long long)l0 + ((unsigned long long)l1 << 16);              // = sum of the lanes' low words
            hi = (long long)l2 + ((long long)l3 << 16);                                // = sum of the lanes' high words
        }
    }
    if (lane < kFp) {
        if (hi) atomicAdd(&s_fp[shard][2 * lane], (unsigned long long)hi);
        if (lo) atomicAdd(&s_fp[shard][2 * lane + 1], lo);
    }
}

// Persistent warps: the units of work -- (group of 32 games, block of 8 plies), group-major inside a
// block index -- are dealt round robin to all warps of the grid, so every warp walks from the opening to the
// endgame (its shard, and with it the Gram accumulator, changes three times) and all warps run out of work
// together.
__global__ void __launch_bounds__(kThreads, 3) learn_kernel(const u64 *__restrict__ traj_black,
                                                            const u64 *__restrict__ traj_white,
                                                            const int32_t *__restrict__ nplies,
                                                            const u64 *__restrict__ final_black,
                                                            const u64 *__restrict__ final_white, int64_t n_games,
                                                            int64_t stride, int t_max, const double *__restrict__ decay,
                                                            unsigned long long *__restrict__ acc)
{
    __shared__ unsigned long long s_xtx[OTHELLO_PHASES][kPairs];
    __shared__ unsigned long long s_fp[OTHELLO_PHASES][2 * kFp];
    // feature bytes of the warp's 32 positions, [side][feature][position]: the K-major operand of the Gram product
    __shared__ unsigned stage[kWarps][2][kX][8];
    for (int i = threadIdx.x; i < OTHELLO_PHASES * kPairs; i += kThreads) (&s_xtx[0][0])[i] = 0ull;
    for (int i = threadIdx.x; i < OTHELLO_PHASES * 2 * kFp; i += kThreads) (&s_fp[0][0])[i] = 0ull;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_groups = (n_games + kGames - 1) / kGames;
    const int64_t n_units = n_groups * (int64_t)((t_max + kPlyBlock) / kPlyBlock);       // blocks cover plies 0..t_max
    const int64_t n_warps = (int64_t)gridDim.x * kWarps;

    Gram gram;
    gram.clear();
    int cur = -1;                                             // shard of the warp's Gram accumulator
    int since_flush = 0;
    unsigned char *stage_b = (unsigned char *)&stage[warp][0][0][0];
    const int fr = lane >> 2, fc = lane & 3;                  // fragment row / column group of this lane

    for (int64_t u = (int64_t)blockIdx.x * kWarps + warp; u < n_units; u += n_warps) {
        const int t0 = (int)(u / n_groups) * kPlyBlock;
        const int64_t g = (u % n_groups) * kGames + lane;
        int len = -1;
        if (g < n_games) {
            len = nplies[g];
            if (len > t_max) len = -1;                        // truncated games are skipped
        }
        const int t1 = min(t0 + kPlyBlock, min(t_max, __reduce_max_sync(kFull, len)) + 1);   // positions 0..nplies are recorded
        if (t0 >= t1) continue;
        double value = 0.0;
        if (t0 <= len) value = (double)(__popcll(final_black[g]) - __popcll(final_white[g]));   // value_for_black (:40-42)
        const u64 *pb = traj_black + g, *pw = traj_white + g;
        double f[kFp];
#pragma unroll
        for (int k = 0; k < kFp; k++) f[k] = 0.0;
        int mine = -1;                                        // shard of the lane's fp64 partial sums
        u64 nb = 0, nw = 0;
        if (t0 <= len) { nb = pb[(int64_t)t0 * stride]; nw = pw[(int64_t)t0 * stride]; }
        for (int t = t0; t < t1; t++) {
            const bool live = t <= len;
            const u64 b = nb, w = nw;
            if (t + 1 < t1 && t + 1 <= len) { nb = pb[(int64_t)(t + 1) * stride]; nw = pw[(int64_t)(t + 1) * stride]; }
            int x[2][kX - 1];
#pragma unroll
            for (int k = 0; k < kX - 1; k++) x[0][k] = x[1][k] = 0;
            int shard = -1;
            if (live) {
                shard = phase_row(__popcll(b | w));
                obf::mobility_both(b, w, x[0][0], x[1][0]);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    x[0][1 + k] = class_count(b, k);
                    x[1][1 + k] = class_count(w, k);
                }
            }
            // a lane's shard changes at most once inside a block (8 plies < 16 discs): close its partial sums;
            // the warp does it together (lanes that do not change contribute nothing)
            const bool change = live && mine >= 0 && shard != mine;
            if (__any_sync(kFull, change)) {
                for (unsigned todo = __ballot_sync(kFull, change); todo;) {
                    const int s = __shfl_sync(kFull, mine, __ffs(todo) - 1);
                    const bool in = change && mine == s;
                    todo &= ~__ballot_sync(kFull, in);
                    fp_flush(f, in, s_fp, s, lane);
                    if (in) {
#pragma unroll
                        for (int k = 0; k < kFp; k++) f[k] = 0.0;
                    }
                }
            }
            if (live) {
                mine = shard;
                const double y = value * __ldg(decay + (len - t));                        // * l ** turn_left (:55)
                // White's target is the negative of Black's (value_for_white, :42); the intercept terms cancel
#pragma unroll
                for (int k = 0; k < kX - 1; k++) f[k] += (double)(x[0][k] - x[1][k]) * y;
                f[kFp - 1] += 2.0 * (y * y);
            }
            // X^T X of the warp's positions, shard by shard (one shard unless games with passes straddle a boundary)
            const unsigned alive = __ballot_sync(kFull, live);
            for (unsigned todo = alive; todo;) {
                const int s = __shfl_sync(kFull, shard, __ffs(todo) - 1);
                const unsigned grp = __ballot_sync(kFull, live && shard == s);
                todo &= ~grp;
                if (s != cur || since_flush >= 4096) {        // (int32 accumulators: 4096 plies x 32 x 2 x 64^2 < 2^31)
                    if (cur >= 0) gram_flush(gram, s_xtx, cur, lane);
                    gram.clear();
                    cur = s;
                    since_flush = 0;
                }
                since_flush++;
                const bool in = (grp >> lane) & 1u;           // (x is all zero for lanes that are not live)
                const bool whole = grp == alive;              // the usual case: nothing to mask
#pragma unroll
                for (int side = 0; side < 2; side++) {
#pragma unroll
                    for (int k = 0; k < kX - 1; k++)
                        stage_b[(side * kX + k) * 32 + lane] = (unsigned char)((whole || in) ? x[side][k] : 0);
                    stage_b[(side * kX + kX - 1) * 32 + lane] = in ? 1 : 0;
                }
                __syncwarp();
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const unsigned a0 = stage[warp][side][fr][fc], a2 = stage[warp][side][fr][4 + fc];
                    const unsigned a1 = fr < 2 ? stage[warp][side][8 + fr][fc] : 0u;
                    const unsigned a3 = fr < 2 ? stage[warp][side][8 + fr][4 + fc] : 0u;
                    mma_s8(gram.lo, a0, a1, a2, a3, a0, a2);       // columns = features 0..7
                    mma_s8(gram.hi, a0, a1, a2, a3, a1, a3);       // columns = features 8, 9
                }
                __syncwarp();
            }
        }
        // the end of a ply block closes the lanes' fp64 partial sums: (game, block, shard) is the unit that is
        // rounded to fixed point, whatever warp, CTA or GPU works on it
        for (unsigned todo = __ballot_sync(kFull, mine >= 0); todo;) {
            const int s = __shfl_sync(kFull, mine, __ffs(todo) - 1);
            const bool in = mine == s;
            todo &= ~__ballot_sync(kFull, in);
            fp_flush(f, in, s_fp, s, lane);
        }
    }
    if (cur >= 0) gram_flush(gram, s_xtx, cur, lane);
    __syncthreads();
    for (int i = threadIdx.x; i < OTHELLO_PHASES * OTHELLO_ACC; i += kThreads) {
        const int s = i / OTHELLO_ACC, k = i % OTHELLO_ACC;
        unsigned long long v = 0ull;
        if (k < kPairs) v = s_xtx[s][k];
        else if (k >= kFpBase && k < kFpBase + 2 * kFp) v = s_fp[s][k - kFpBase];
        if (v) atomicAdd(&acc[i], v);
    }
}

// integers -> the doubles the solver reads: stats[s] = XtX[10][10], Xty[10], n, sum y^2
__global__ void __launch_bounds__(128) stats_kernel(const long long *__restrict__ acc, double *__restrict__ stats)
{
    const int s = blockIdx.x;
    const long long *a = acc + s * OTHELLO_ACC;
    double *out = stats + s * OTHELLO_STATS;
    for (int k = threadIdx.x; k < OTHELLO_STATS; k += blockDim.x) out[k] = stat_from_acc(a, k);
}

}  // namespace

extern "C" int othello_learn_accumulate(const uint64_t *traj_black, const uint64_t *traj_white, const int32_t *nplies,
                                        const uint64_t *final_black, const uint64_t *final_white, int64_t n_games,
                                        int64_t stride, int32_t t_max, const double *decay, int64_t *acc, void *stream)
{
    OB_CHECK_ARGS(n_games >= 0 && t_max >= 0 && acc && decay);
    if (n_games == 0) return 0;
    OB_CHECK_ARGS(traj_black && traj_white && nplies && final_black && final_white && stride >= n_games);
    const int64_t units = ((n_games + kGames - 1) / kGames) * (int64_t)((t_max + kPlyBlock) / kPlyBlock);
    int dev = 0, sms = 148;
    OB_CUDA(cudaGetDevice(&dev));
    OB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t resident = (int64_t)sms * 3;               // persistent: one wave of CTAs (3 per SM at 80 registers)
    const int64_t want = (units + kWarps - 1) / kWarps;
    const unsigned blocks = (unsigned)(want < resident ? want : resident);
    learn_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
        (const u64 *)traj_black, (const u64 *)traj_white, nplies, (const u64 *)final_black, (const u64 *)final_white,
        n_games, stride, t_max, decay, (unsigned long long *)acc);
    return ob_launch_status();
}

extern "C" int othello_learn_stats(const int64_t *acc, double *stats, void *stream)
{
    OB_CHECK_ARGS(acc && stats);
    stats_kernel<<<OTHELLO_PHASES, 128, 0, (cudaStream_t)stream>>>((const long long *)acc, stats);
    return ob_launch_status();
}
