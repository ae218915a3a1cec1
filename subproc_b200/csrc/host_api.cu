// host_api.cu -- the host-buffer front end of the C ABI: what a caller without torch (the
// ctypes / cgo style binding of INTEGRATION.md) uses.  A context owns one stream and one
// grow-only device workspace; every call copies its inputs in, launches the same kernels as
// the device-pointer entry points, copies results out and synchronises.
#include <new>
#include "common.cuh"

constexpr int kPipeStreams = 3;     // playout_host pipelines H2D / kernel / D2H of game chunks over these

struct othello_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t pipe[kPipeStreams];
    cudaEvent_t ready, done[kPipeStreams];
    char *ws;
    size_t ws_bytes;
    // last playout trajectory (views into ws)
    uint64_t *traj_black, *traj_white;
    uint8_t *traj_move;
    int64_t traj_stride;
    int32_t traj_t_max;
};

namespace {

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int reserve(othello_ctx *c, size_t bytes)
{
    if (bytes <= c->ws_bytes) return 0;
    if (c->ws) { OB_CUDA(cudaFree(c->ws)); c->ws = nullptr; c->ws_bytes = 0; }
    c->traj_black = c->traj_white = nullptr; c->traj_move = nullptr;
    OB_CUDA(cudaMalloc((void **)&c->ws, bytes));
    c->ws_bytes = bytes;
    return 0;
}

// bump allocator over the workspace
struct Carver {
    char *base; size_t off;
    template <typename T> T *take(size_t count) { T *p = (T *)(base + off); off += align256(count * sizeof(T)); return p; }
};

}  // namespace

extern "C" {

int othello_abi_version(void) { return OTHELLO_ABI_VERSION; }

const char *othello_error_string(int code)
{
    switch (code) {
    case 0: return "ok";
    case OTHELLO_E_INVALID: return "invalid argument";
    case OTHELLO_E_WORKSPACE: return "workspace too small";
    case OTHELLO_E_NO_DEVICE: return "no CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int othello_ctx_create(int device, othello_ctx **out)
{
    OB_CHECK_ARGS(out != nullptr);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return OTHELLO_E_NO_DEVICE;
    OB_CHECK_ARGS(device >= 0 && device < count);
    OB_CUDA(cudaSetDevice(device));
    othello_ctx *c = new (std::nothrow) othello_ctx();
    if (!c) return OTHELLO_E_INVALID;
    c->device = device; c->ws = nullptr; c->ws_bytes = 0;
    c->traj_black = c->traj_white = nullptr; c->traj_move = nullptr; c->traj_stride = 0; c->traj_t_max = 0;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return (int)e; }
    e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    for (int i = 0; i < kPipeStreams && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { delete c; return (int)e; }
    *out = c;
    return 0;
}

void othello_ctx_destroy(othello_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->ws) cudaFree(c->ws);
    for (int i = 0; i < kPipeStreams; i++) { cudaStreamDestroy(c->pipe[i]); cudaEventDestroy(c->done[i]); }
    cudaEventDestroy(c->ready);
    cudaStreamDestroy(c->stream);
    delete c;
}

int othello_legal_host(othello_ctx *c, const uint64_t *own, const uint64_t *opp, uint64_t *legal, int64_t n)
{
    OB_CHECK_ARGS(c && n >= 0 && (n == 0 || (own && opp && legal)));
    if (n == 0) return 0;
    OB_CUDA(cudaSetDevice(c->device));
    int rc = reserve(c, 3 * align256((size_t)n * 8));
    if (rc) return rc;
    Carver k = {c->ws, 0};
    uint64_t *d_own = k.take<uint64_t>(n), *d_opp = k.take<uint64_t>(n), *d_legal = k.take<uint64_t>(n);
    OB_CUDA(cudaMemcpyAsync(d_own, own, n * 8, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(d_opp, opp, n * 8, cudaMemcpyHostToDevice, c->stream));
    rc = othello_legal(d_own, d_opp, d_legal, n, c->stream);
    if (rc) return rc;
    OB_CUDA(cudaMemcpyAsync(legal, d_legal, n * 8, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int othello_step_host(othello_ctx *c, uint64_t *black, uint64_t *white, uint8_t *turn, int32_t *nturn,
                      const uint8_t *move, uint64_t *flips_out, int32_t *ret, uint8_t *flags, int64_t n)
{
    OB_CHECK_ARGS(c && n >= 0 && (n == 0 || (black && white && turn && nturn && move)));
    if (n == 0) return 0;
    OB_CUDA(cudaSetDevice(c->device));
    int rc = reserve(c, 3 * align256((size_t)n * 8) + 2 * align256((size_t)n * 4) + 3 * align256((size_t)n));
    if (rc) return rc;
    Carver k = {c->ws, 0};
    uint64_t *d_b = k.take<uint64_t>(n), *d_w = k.take<uint64_t>(n), *d_f = k.take<uint64_t>(n);
    int32_t *d_nt = k.take<int32_t>(n), *d_ret = k.take<int32_t>(n);
    uint8_t *d_t = k.take<uint8_t>(n), *d_mv = k.take<uint8_t>(n), *d_fl = k.take<uint8_t>(n);
    cudaStream_t s = c->stream;
    OB_CUDA(cudaMemcpyAsync(d_b, black, n * 8, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(d_w, white, n * 8, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(d_t, turn, n, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(d_nt, nturn, n * 4, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(d_mv, move, n, cudaMemcpyHostToDevice, s));
    rc = othello_step(d_b, d_w, d_t, d_nt, d_mv, d_f, d_ret, d_fl, n, s);
    if (rc) return rc;
    OB_CUDA(cudaMemcpyAsync(black, d_b, n * 8, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaMemcpyAsync(white, d_w, n * 8, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaMemcpyAsync(turn, d_t, n, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaMemcpyAsync(nturn, d_nt, n * 4, cudaMemcpyDeviceToHost, s));
    if (flips_out) OB_CUDA(cudaMemcpyAsync(flips_out, d_f, n * 8, cudaMemcpyDeviceToHost, s));
    if (ret) OB_CUDA(cudaMemcpyAsync(ret, d_ret, n * 4, cudaMemcpyDeviceToHost, s));
    if (flags) OB_CUDA(cudaMemcpyAsync(flags, d_fl, n, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int othello_playout_host(othello_ctx *c, uint64_t seed, uint64_t gid0, int64_t n, const uint64_t *black0,
                         const uint64_t *white0, const uint8_t *turn0, int32_t policy, int32_t random_plies,
                         int32_t n_rand_black, int32_t n_rand_white, const float *weights, int32_t policy_white,
                         const float *weights_white, int32_t t_max, uint64_t *traj_black, uint64_t *traj_white, uint8_t *traj_move, int32_t *nplies,
                         uint64_t *final_black, uint64_t *final_white)
{
    OB_CHECK_ARGS(c && n >= 0 && t_max >= 0);
    if (n == 0) return 0;
    OB_CHECK_ARGS(nplies && final_black && final_white);
    OB_CHECK_ARGS((black0 == nullptr) == (white0 == nullptr));
    OB_CHECK_ARGS((traj_black == nullptr) == (traj_white == nullptr) && (traj_black == nullptr) == (traj_move == nullptr));
    OB_CUDA(cudaSetDevice(c->device));
    const size_t row8 = align256((size_t)n * 8), row1 = align256((size_t)n);
    const size_t tb_bytes = align256((size_t)(t_max + 1) * n * 8), tm_bytes = align256((size_t)t_max * n + 1);
    int rc = reserve(c, 4 * row8 + row1 + align256((size_t)n * 4) + 512 + 2 * tb_bytes + tm_bytes);
    if (rc) return rc;
    Carver k = {c->ws, 0};
    uint64_t *d_b0 = k.take<uint64_t>(n), *d_w0 = k.take<uint64_t>(n), *d_fb = k.take<uint64_t>(n), *d_fw = k.take<uint64_t>(n);
    uint8_t *d_t0 = k.take<uint8_t>(n);
    int32_t *d_np = k.take<int32_t>(n);
    float *d_wt = k.take<float>(OTHELLO_PHASES * OTHELLO_WEIGHTS), *d_wt2 = k.take<float>(OTHELLO_PHASES * OTHELLO_WEIGHTS);
    uint64_t *d_tb = k.take<uint64_t>((size_t)(t_max + 1) * n), *d_tw = k.take<uint64_t>((size_t)(t_max + 1) * n);
    uint8_t *d_tm = k.take<uint8_t>((size_t)t_max * n + 1);
    cudaStream_t s = c->stream;
    if (weights) OB_CUDA(cudaMemcpyAsync(d_wt, weights, sizeof(float) * OTHELLO_PHASES * OTHELLO_WEIGHTS, cudaMemcpyHostToDevice, s));
    if (weights_white) OB_CUDA(cudaMemcpyAsync(d_wt2, weights_white, sizeof(float) * OTHELLO_PHASES * OTHELLO_WEIGHTS, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaEventRecord(c->ready, s));

    // Games are independent, so the batch is cut into chunks whose copy-in, kernel and copy-out run
    // on rotating streams: the PCIe traffic of one chunk hides behind the integer work of the others
    // (with pinned host buffers; pageable buffers simply serialise).
    int64_t chunk = (n + 7) / 8;
    if (chunk < 65536) chunk = 65536;
    chunk = (chunk + 127) & ~(int64_t)127;
    int used = 0;
    for (int64_t c0 = 0, i = 0; c0 < n; c0 += chunk, i++) {
        const int64_t m = (n - c0 < chunk) ? n - c0 : chunk;
        const int si = (int)(i % kPipeStreams);
        cudaStream_t st = c->pipe[si];
        if (i < kPipeStreams) { OB_CUDA(cudaStreamWaitEvent(st, c->ready, 0)); used = (int)i + 1; }
        if (black0) {
            OB_CUDA(cudaMemcpyAsync(d_b0 + c0, black0 + c0, m * 8, cudaMemcpyHostToDevice, st));
            OB_CUDA(cudaMemcpyAsync(d_w0 + c0, white0 + c0, m * 8, cudaMemcpyHostToDevice, st));
        }
        if (turn0) OB_CUDA(cudaMemcpyAsync(d_t0 + c0, turn0 + c0, m, cudaMemcpyHostToDevice, st));
        othello_playout_args a;
        a.seed = seed; a.gid0 = gid0 + (uint64_t)c0; a.n_games = m;
        a.black0 = black0 ? d_b0 + c0 : nullptr; a.white0 = black0 ? d_w0 + c0 : nullptr; a.turn0 = turn0 ? d_t0 + c0 : nullptr;
        a.policy = policy; a.random_plies = random_plies; a.n_rand_black = n_rand_black; a.n_rand_white = n_rand_white;
        a.weights = weights ? d_wt : nullptr;
        a.policy_white = policy_white; a.reserved = 0; a.weights_white = weights_white ? d_wt2 : nullptr;
        a.t_max = t_max; a.stride = n;
        a.traj_black = d_tb + c0; a.traj_white = d_tw + c0; a.traj_move = d_tm + c0;
        a.nplies = d_np + c0; a.final_black = d_fb + c0; a.final_white = d_fw + c0;
        rc = othello_playout(&a, st);
        if (rc) return rc;
        OB_CUDA(cudaMemcpyAsync(nplies + c0, d_np + c0, m * 4, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(final_black + c0, d_fb + c0, m * 8, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(final_white + c0, d_fw + c0, m * 8, cudaMemcpyDeviceToHost, st));
        if (traj_black) {
            // host and device trajectories are both [t][n]: a chunk is a column block
            OB_CUDA(cudaMemcpy2DAsync(traj_black + c0, (size_t)n * 8, d_tb + c0, (size_t)n * 8, (size_t)m * 8, t_max + 1, cudaMemcpyDeviceToHost, st));
            OB_CUDA(cudaMemcpy2DAsync(traj_white + c0, (size_t)n * 8, d_tw + c0, (size_t)n * 8, (size_t)m * 8, t_max + 1, cudaMemcpyDeviceToHost, st));
            if (t_max > 0)
                OB_CUDA(cudaMemcpy2DAsync(traj_move + c0, (size_t)n, d_tm + c0, (size_t)n, (size_t)m, t_max, cudaMemcpyDeviceToHost, st));
        }
    }
    c->traj_black = d_tb; c->traj_white = d_tw; c->traj_move = d_tm; c->traj_stride = n; c->traj_t_max = t_max;
    for (int i = 0; i < used; i++) {
        OB_CUDA(cudaEventRecord(c->done[i], c->pipe[i]));
        OB_CUDA(cudaStreamWaitEvent(s, c->done[i], 0));
    }
    OB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int othello_ctx_trajectory(othello_ctx *c, uint64_t **traj_black, uint64_t **traj_white, uint8_t **traj_move,
                           int64_t *stride, int32_t *t_max)
{
    OB_CHECK_ARGS(c && c->traj_black);
    if (traj_black) *traj_black = c->traj_black;
    if (traj_white) *traj_white = c->traj_white;
    if (traj_move) *traj_move = c->traj_move;
    if (stride) *stride = c->traj_stride;
    if (t_max) *t_max = c->traj_t_max;
    return 0;
}

}  // extern "C"
