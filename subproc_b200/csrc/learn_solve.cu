// learn_solve.cu -- the four per-phase regressions of the learner, solved ON the device so that a
// learning iteration (self-play -> statistics -> all-reduce -> refit -> next self-play) never waits
// for the host.
//
// Mirrors subproc_b200/learner.py::solve_shard, i.e. sklearn's LinearRegression(fit_intercept=True)
// as the reference calls it (progress_position_moves_learn.py:167-181): centre the normal equations,
// minimum-norm solution on the centred 9x9 Gram matrix (constant columns get coefficient 0), intercept,
// RMSE and R^2 on the same statistics, then `coef * 127 / max|coef|` (:180-181) and int() truncation
// toward zero (:200).  One warp per shard; the symmetric eigen-decomposition is parallel-order cyclic
// Jacobi in fp64 (9x9: 9 rounds of 4 disjoint rotations per sweep, a handful of sweeps).
#include "common.cuh"
#include "learn_acc.cuh"

namespace {

constexpr int kN = 9;                       // regressors without the intercept
constexpr int kSweeps = 16;

// (prev_weights and weights may be the same table -- a shard without samples copies its own row -- so
// neither carries __restrict__)
__device__ __forceinline__ void solve_shard(const double *row, int s, int lane, const float *prev_weights, float *weights,
                                            int32_t *__restrict__ params, double *__restrict__ fits, double rcond)
{
    __shared__ double A[kN][kN], V[kN][kN], sxy[kN], coef[kN], xbar[kN], lam_inv[kN];
    __shared__ int live[kN];
    const double *xtx = row, *xty = row + 100;
    const double n = row[110], syy = row[111];
    float *w_out = weights + s * OTHELLO_WEIGHTS;
    double *fit = fits + s * 16;
    if (n <= 0.0) {                                             // nothing seen in this phase: keep the old row
        if (lane < OTHELLO_WEIGHTS) w_out[lane] = prev_weights[s * OTHELLO_WEIGHTS + lane];
        if (lane < kN) params[s * kN + lane] = (int32_t)prev_weights[s * OTHELLO_WEIGHTS + lane];
        if (lane < 16) fit[lane] = 0.0;
        return;
    }
    const double ybar = xty[9] / n;
    if (lane < kN) xbar[lane] = xtx[lane * 10 + 9] / n;          // column of ones: sum x_i
    __syncwarp();
    if (lane < kN) {
        for (int j = 0; j < kN; j++) {
            A[lane][j] = xtx[lane * 10 + j] - n * xbar[lane] * xbar[j];     // centred Gram matrix
            V[lane][j] = lane == j ? 1.0 : 0.0;
        }
        sxy[lane] = xty[lane] - n * xbar[lane] * ybar;
    }
    __syncwarp();
    if (lane < kN) live[lane] = A[lane][lane] > 0.0;             // constant columns (classes nobody owns yet)
    __syncwarp();
    if (lane < kN)
        for (int j = 0; j < kN; j++)
            if (!live[lane] || !live[j]) A[lane][j] = 0.0;
    __syncwarp();
    double tr = 0.0;
    for (int j = 0; j < kN; j++) tr += A[j][j];                  // invariant under the rotations
    // Parallel-order cyclic Jacobi: a sweep is 9 rounds of 4 rotations in disjoint planes (round-robin
    // pairing of 9 indices + one bye), so the four (c, s) of a round are computed together from the current
    // matrix and A <- J^T A J is applied as one column pass (lane = row) and one row pass (lane = column).
    __shared__ double rc[4], rs[4];
    __shared__ int rp[4], rq[4];
    for (int sweep = 0; sweep < kSweeps; sweep++) {
        double off = 0.0;
        for (int p = 0; p < kN; p++) for (int q = p + 1; q < kN; q++) off += A[p][q] * A[p][q];
        if (off <= 1e-32 * tr * tr) break;                      // off-diagonal mass below fp64 resolution of the spectrum
        for (int round = 0; round < kN; round++) {
            if (lane < 4) {
                // circle method on 10 slots (slot 9 = bye): pairs (round + k, round - k) mod 9 skip the bye
                int p = (round + lane + 1) % kN, q = (round + kN - lane - 1) % kN;
                if (p > q) { const int t = p; p = q; q = t; }
                const double apq = A[p][q];
                double c = 1.0, sn = 0.0;
                if (apq != 0.0) {
                    const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    sn = t * c;
                }
                rp[lane] = p; rq[lane] = q; rc[lane] = c; rs[lane] = sn;
            }
            __syncwarp();
            if (lane < kN) {                                    // columns p, q of row `lane`, of A and of V
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = rp[k], q = rq[k];
                    const double c = rc[k], sn = rs[k];
                    const double ap = A[lane][p], aq = A[lane][q];
                    A[lane][p] = c * ap - sn * aq; A[lane][q] = sn * ap + c * aq;
                    const double vp = V[lane][p], vq = V[lane][q];
                    V[lane][p] = c * vp - sn * vq; V[lane][q] = sn * vp + c * vq;
                }
            }
            __syncwarp();
            if (lane < kN) {                                    // rows p, q of column `lane`
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = rp[k], q = rq[k];
                    const double c = rc[k], sn = rs[k];
                    const double ap = A[p][lane], aq = A[q][lane];
                    A[p][lane] = c * ap - sn * aq; A[q][lane] = sn * ap + c * aq;
                }
            }
            __syncwarp();
            if (lane < 4) { A[rp[lane]][rq[lane]] = 0.0; A[rq[lane]][rp[lane]] = 0.0; }   // annihilated exactly
            __syncwarp();
        }
    }
    double lmax = 0.0;
    for (int j = 0; j < kN; j++) lmax = fmax(lmax, A[j][j]);
    if (lane < kN) lam_inv[lane] = (A[lane][lane] > rcond * lmax) ? 1.0 / A[lane][lane] : 0.0;
    __syncwarp();
    if (lane < kN) {
        double acc = 0.0;                                       // coef = V diag(1/lambda) V^T sxy
        for (int j = 0; j < kN; j++) {
            double proj = 0.0;
            for (int k = 0; k < kN; k++) proj += V[k][j] * sxy[k];
            acc += V[lane][j] * lam_inv[j] * proj;
        }
        coef[lane] = live[lane] ? acc : 0.0;
    }
    __syncwarp();
    if (lane == 0) {
        double w[10], dot = 0.0;
        for (int i = 0; i < kN; i++) { w[i] = coef[i]; dot += xbar[i] * coef[i]; }
        w[9] = ybar - dot;                                      // intercept
        double wxty = 0.0, wxw = 0.0;
        for (int i = 0; i < 10; i++) {
            wxty += w[i] * xty[i];
            double r = 0.0;
            for (int j = 0; j < 10; j++) r += xtx[i * 10 + j] * w[j];
            wxw += w[i] * r;
        }
        double sse = syy - 2.0 * wxty + wxw;
        if (sse < 0.0) sse = 0.0;
        const double sst = syy - n * ybar * ybar;
        for (int i = 0; i < kN; i++) fit[i] = coef[i];
        fit[9] = w[9];
        fit[10] = sqrt(sse / n);
        fit[11] = sst > 0.0 ? 1.0 - sse / sst : 0.0;
        fit[12] = n;
        fit[13] = fit[14] = fit[15] = 0.0;
        double m = 0.0;
        for (int i = 0; i < kN; i++) m = fmax(m, fabs(coef[i]));
        const double k127 = m > 0.0 ? 127.0 / m : 0.0;          // coef = 127 / max|coef| (:180)
        for (int i = 0; i < kN; i++) {
            const int32_t q = (int32_t)(coef[i] * k127);        // int(x): truncation toward zero (:200)
            params[s * kN + i] = q;
            w_out[i] = (float)q;
        }
        w_out[9] = 0.0f;                                        // the intercept is dropped (:181)
    }
}

__global__ void __launch_bounds__(32) solve_kernel(const double *__restrict__ stats, const float *prev_weights,
                                                   float *weights, int32_t *__restrict__ params,
                                                   double *__restrict__ fits, double rcond)
{
    solve_shard(stats + (size_t)blockIdx.x * OTHELLO_STATS, blockIdx.x, threadIdx.x, prev_weights, weights, params, fits, rcond);
}

// the same from the exact integer accumulators, in one launch: accumulators -> statistics (shared memory, and
// `stats` if wanted) -> regression; `clear` leaves the accumulators zeroed for the next iteration
__global__ void __launch_bounds__(32) refit_kernel(long long *__restrict__ acc, int clear, double *__restrict__ stats,
                                                   const float *prev_weights, float *weights, int32_t *__restrict__ params,
                                                   double *__restrict__ fits, double rcond)
{
    __shared__ double row[OTHELLO_STATS];
    const int s = blockIdx.x, lane = threadIdx.x;
    long long *a = acc + s * OTHELLO_ACC;
    for (int k = lane; k < OTHELLO_STATS; k += 32) {
        row[k] = obl::stat_from_acc(a, k);
        if (stats) stats[s * OTHELLO_STATS + k] = row[k];
    }
    __syncwarp();
    if (clear)
        for (int k = lane; k < OTHELLO_ACC; k += 32) a[k] = 0;
    solve_shard(row, s, lane, prev_weights, weights, params, fits, rcond);
}

}  // namespace

extern "C" int othello_learn_refit(int64_t *acc, int32_t clear, double *stats, const float *prev_weights, float *weights,
                                   int32_t *params, double *fits, void *stream)
{
    OB_CHECK_ARGS(acc && prev_weights && weights && params && fits);
    refit_kernel<<<OTHELLO_PHASES, 32, 0, (cudaStream_t)stream>>>((long long *)acc, clear, stats, prev_weights, weights, params,
                                                                 fits, 1e-12);
    return ob_launch_status();
}

extern "C" int othello_learn_solve(const double *stats, const float *prev_weights, float *weights, int32_t *params,
                                   double *fits, void *stream)
{
    OB_CHECK_ARGS(stats && prev_weights && weights && params && fits);
    solve_kernel<<<OTHELLO_PHASES, 32, 0, (cudaStream_t)stream>>>(stats, prev_weights, weights, params, fits, 1e-12);
    return ob_launch_status();
}
