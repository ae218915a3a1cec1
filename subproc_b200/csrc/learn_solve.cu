// learn_solve.cu -- the four per-phase regressions of the learner, solved ON the device so that a
// learning iteration (self-play -> statistics -> all-reduce -> refit -> next self-play) never waits
// for the host.
//
// Mirrors subproc_b200/learner.py::solve_shard, i.e. sklearn's LinearRegression(fit_intercept=True)
// as the reference calls it (progress_position_moves_learn.py:167-181): centre the normal equations,
// minimum-norm solution on the centred 9x9 Gram matrix (constant columns get coefficient 0), intercept,
// RMSE and R^2 on the same statistics, then `coef * 127 / max|coef|` (:180-181) and int() truncation
// toward zero (:200).  One warp per shard; the symmetric eigen-decomposition is cyclic Jacobi in fp64
// (9x9: a dozen sweeps), rotations applied by 9 lanes in parallel.
#include "common.cuh"

namespace {

constexpr int kN = 9;                       // regressors without the intercept
constexpr int kSweeps = 16;

// (prev_weights and weights may be the same table -- a shard without samples copies its own row -- so
// neither carries __restrict__)
__global__ void __launch_bounds__(32) solve_kernel(const double *__restrict__ stats, const float *prev_weights,
                                                   float *weights, int32_t *__restrict__ params,
                                                   double *__restrict__ fits, double rcond)
{
    __shared__ double A[kN][kN], V[kN][kN], sxy[kN], coef[kN], xbar[kN], lam_inv[kN];
    __shared__ int live[kN];
    const int s = blockIdx.x, lane = threadIdx.x;
    const double *row = stats + (size_t)s * OTHELLO_STATS;
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
    for (int sweep = 0; sweep < kSweeps; sweep++) {
        double off = 0.0;
        for (int p = 0; p < kN; p++) for (int q = p + 1; q < kN; q++) off += A[p][q] * A[p][q];
        if (off <= 1e-34 * tr * tr) break;                      // off-diagonal mass below fp64 resolution of the spectrum
        for (int p = 0; p < kN - 1; p++) {
            for (int q = p + 1; q < kN; q++) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;                       // warp-uniform: every lane reads the same value
                const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
                const double app = A[p][p], aqq = A[q][q];
                double akp = 0.0, akq = 0.0, vkp = 0.0, vkq = 0.0;
                if (lane < kN) {
                    akp = A[lane][p]; akq = A[lane][q];
                    vkp = V[lane][p]; vkq = V[lane][q];
                }
                __syncwarp();
                if (lane < kN) {
                    if (lane != p && lane != q) {
                        const double np_ = c * akp - sn * akq, nq_ = sn * akp + c * akq;
                        A[lane][p] = np_; A[p][lane] = np_;
                        A[lane][q] = nq_; A[q][lane] = nq_;
                    }
                    V[lane][p] = c * vkp - sn * vkq;
                    V[lane][q] = sn * vkp + c * vkq;
                }
                if (lane == 0) {
                    A[p][p] = app - t * apq; A[q][q] = aqq + t * apq;
                    A[p][q] = 0.0; A[q][p] = 0.0;
                }
                __syncwarp();
            }
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

}  // namespace

extern "C" int othello_learn_solve(const double *stats, const float *prev_weights, float *weights, int32_t *params,
                                   double *fits, void *stream)
{
    OB_CHECK_ARGS(stats && prev_weights && weights && params && fits);
    solve_kernel<<<OTHELLO_PHASES, 32, 0, (cudaStream_t)stream>>>(stats, prev_weights, weights, params, fits, 1e-12);
    return ob_launch_status();
}
