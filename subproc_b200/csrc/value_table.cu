// value_table.cu -- the reference's order-dependent value table, exactly
// (progress_position_moves_learn.py:37-62), from trajectories in HBM.
//
// The reference keeps, in Redis, one float per distinct counts() 10-tuple and updates it once per
// (position, side) in a fixed order: books by ascending id, positions from the terminal one back to
// the start (replearn.py:37-38), side 'O' (Black) then 'X' (White) (:44-47):
//     new = float(value) * (l ** turn_left)                 value = own - opp final discs, l = 0.90
//     V   = new                     if V == 0
//     V   = V * (1 - a) + new * a   otherwise                a = 0.03
// Different keys are independent; the updates of ONE key form a sequential fp64 recurrence whose
// result depends on the order.  So: (1) records_kernel emits (key, new) for every (position, side)
// at its position in the reference's order; (2) the host side groups equal keys with a STABLE sort
// (order inside a key is preserved); (3) smooth_kernel walks each key's run sequentially, one thread
// per key, with the reference's exact operation order (two roundings for the products, one for the
// sum -- no FMA contraction), starting from the value already in the table.
#include "common.cuh"
#include "fastboard.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 256;

// counts() 10-tuple packed into 43 bits: discs(7) mobility(6) a(3) b(4) c(3) d(4) e(4) f(5) g(3) h(4)
__device__ __forceinline__ u64 pack_features(u64 side, u64 other)
{
    u64 k = (u64)__popcll(side | other);
    k = (k << 6) | (u64)__popcll(obf::legal_moves(side, other));
    k = (k << 3) | (u64)__popcll(side & kClassMask[0]);
    k = (k << 4) | (u64)__popcll(side & kClassMask[1]);
    k = (k << 3) | (u64)__popcll(side & kClassMask[2]);
    k = (k << 4) | (u64)__popcll(side & kClassMask[3]);
    k = (k << 4) | (u64)__popcll(side & kClassMask[4]);
    k = (k << 5) | (u64)__popcll(side & kClassMask[5]);
    k = (k << 3) | (u64)__popcll(side & kClassMask[6]);
    k = (k << 4) | (u64)__popcll(side & kClassMask[7]);
    return k;
}

__global__ void __launch_bounds__(kThreads) records_kernel(const u64 *__restrict__ traj_black,
                                                           const u64 *__restrict__ traj_white,
                                                           const int32_t *__restrict__ nplies,
                                                           const u64 *__restrict__ final_black,
                                                           const u64 *__restrict__ final_white, int64_t n_games,
                                                           int64_t stride, int t_max, const double *__restrict__ decay,
                                                           const int64_t *__restrict__ rec_base, u64 *__restrict__ keys,
                                                           double *__restrict__ targets)
{
    const int64_t tiles_per_row = (n_games + kThreads - 1) / kThreads;
    const int64_t tile = blockIdx.x;
    const int t = (int)(tile / tiles_per_row);
    const int64_t g = (tile % tiles_per_row) * kThreads + threadIdx.x;
    if (g >= n_games) return;
    const int len = nplies[g];
    if (t > len || len > t_max) return;
    const u64 b = traj_black[(int64_t)t * stride + g], w = traj_white[(int64_t)t * stride + g];
    const int value_black = __popcll(final_black[g]) - __popcll(final_white[g]);     // :40-42
    const double d = decay[len - t];                                                // l ** turn_left (:55)
    const int64_t at = rec_base[g] + 2 * (int64_t)(len - t);                        // terminal position first
    keys[at] = pack_features(b, w);
    targets[at] = __dmul_rn((double)value_black, d);
    keys[at + 1] = pack_features(w, b);
    targets[at + 1] = __dmul_rn((double)(-value_black), d);
}

__device__ __forceinline__ double smooth_step(double v, double nv, double keep, double a)
{
    return (v == 0.0) ? nv : __dadd_rn(__dmul_rn(v, keep), __dmul_rn(nv, a));      // :56-61
}

// One thread per key walks its run in update order.  The recurrence is sequential by definition; the
// loads are not, so they are issued eight at a time ahead of the dependent fp64 chain (the opening
// positions are visited by every game: their runs are 2 x n_games long).
__global__ void __launch_bounds__(kThreads) smooth_kernel(const double *__restrict__ targets,
                                                          const int64_t *__restrict__ seg_start,
                                                          const double *__restrict__ init, double a,
                                                          double *__restrict__ out, int64_t n_seg)
{
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= n_seg) return;
    const double keep = __dsub_rn(1.0, a);                                          // (1 - self.a)
    double v = init[s];
    int64_t i = seg_start[s];
    const int64_t e = seg_start[s + 1];
    for (; i + 8 <= e; i += 8) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = __ldg(targets + i + j);
#pragma unroll
        for (int j = 0; j < 8; j++) v = smooth_step(v, x[j], keep, a);
    }
    for (; i < e; i++) v = smooth_step(v, targets[i], keep, a);
    out[s] = v;
}

__global__ void __launch_bounds__(kThreads) unpack_kernel(const u64 *__restrict__ keys, int32_t *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    u64 k = keys[i];
    const int widths[10] = {7, 6, 3, 4, 3, 4, 4, 5, 3, 4};
#pragma unroll
    for (int f = 9; f >= 0; f--) {
        out[10 * i + f] = (int32_t)(k & ((1ull << widths[f]) - 1));
        k >>= widths[f];
    }
}

}  // namespace

extern "C" {

int othello_value_records(const uint64_t *traj_black, const uint64_t *traj_white, const int32_t *nplies,
                          const uint64_t *final_black, const uint64_t *final_white, int64_t n_games, int64_t stride,
                          int32_t t_max, const double *decay, const int64_t *rec_base, uint64_t *keys, double *targets,
                          void *stream)
{
    OB_CHECK_ARGS(n_games >= 0 && t_max >= 0);
    if (n_games == 0) return 0;
    OB_CHECK_ARGS(traj_black && traj_white && nplies && final_black && final_white && decay && rec_base && keys &&
                  targets && stride >= n_games);
    const int64_t tiles = ((n_games + kThreads - 1) / kThreads) * (int64_t)(t_max + 1);
    records_kernel<<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(
        (const u64 *)traj_black, (const u64 *)traj_white, nplies, (const u64 *)final_black, (const u64 *)final_white,
        n_games, stride, t_max, decay, rec_base, (u64 *)keys, targets);
    return ob_launch_status();
}

int othello_value_smooth(const double *targets, const int64_t *seg_start, const double *init, double a, double *out,
                         int64_t n_seg, void *stream)
{
    OB_CHECK_ARGS(n_seg >= 0);
    if (n_seg == 0) return 0;
    OB_CHECK_ARGS(targets && seg_start && init && out);
    smooth_kernel<<<ob_blocks(n_seg, kThreads), kThreads, 0, (cudaStream_t)stream>>>(targets, seg_start, init, a, out, n_seg);
    return ob_launch_status();
}

int othello_unpack_keys(const uint64_t *keys, int32_t *features, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0);
    if (n == 0) return 0;
    OB_CHECK_ARGS(keys && features);
    unpack_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)keys, features, n);
    return ob_launch_status();
}

}  // extern "C"
