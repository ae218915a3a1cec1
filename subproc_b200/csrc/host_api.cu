// host_api.cu -- the host-buffer front end of the C ABI: what a caller without torch (the
// ctypes / cgo style binding of INTEGRATION.md) uses.  A context owns one stream and one
// grow-only device workspace; every call copies its inputs in, launches the same kernels as
// the device-pointer entry points, copies results out and synchronises.
#include <new>
#include "common.cuh"

struct othello_ctx {
    int device;
    cudaStream_t stream;
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
    *out = c;
    return 0;
}

void othello_ctx_destroy(othello_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->ws) cudaFree(c->ws);
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
                         int32_t n_rand_black, int32_t n_rand_white, const float *weights, int32_t t_max,
                         uint64_t *traj_black, uint64_t *traj_white, uint8_t *traj_move, int32_t *nplies,
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
    int rc = reserve(c, 4 * row8 + row1 + align256((size_t)n * 4) + 256 + 2 * tb_bytes + tm_bytes);
    if (rc) return rc;
    Carver k = {c->ws, 0};
    uint64_t *d_b0 = k.take<uint64_t>(n), *d_w0 = k.take<uint64_t>(n), *d_fb = k.take<uint64_t>(n), *d_fw = k.take<uint64_t>(n);
    uint8_t *d_t0 = k.take<uint8_t>(n);
    int32_t *d_np = k.take<int32_t>(n);
    float *d_wt = k.take<float>(OTHELLO_PHASES * OTHELLO_WEIGHTS);
    uint64_t *d_tb = k.take<uint64_t>((size_t)(t_max + 1) * n), *d_tw = k.take<uint64_t>((size_t)(t_max + 1) * n);
    uint8_t *d_tm = k.take<uint8_t>((size_t)t_max * n + 1);
    cudaStream_t s = c->stream;
    if (black0) {
        OB_CUDA(cudaMemcpyAsync(d_b0, black0, n * 8, cudaMemcpyHostToDevice, s));
        OB_CUDA(cudaMemcpyAsync(d_w0, white0, n * 8, cudaMemcpyHostToDevice, s));
    }
    if (turn0) OB_CUDA(cudaMemcpyAsync(d_t0, turn0, n, cudaMemcpyHostToDevice, s));
    if (weights) OB_CUDA(cudaMemcpyAsync(d_wt, weights, sizeof(float) * OTHELLO_PHASES * OTHELLO_WEIGHTS, cudaMemcpyHostToDevice, s));

    othello_playout_args a;
    a.seed = seed; a.gid0 = gid0; a.n_games = n;
    a.black0 = black0 ? d_b0 : nullptr; a.white0 = black0 ? d_w0 : nullptr; a.turn0 = turn0 ? d_t0 : nullptr;
    a.policy = policy; a.random_plies = random_plies; a.n_rand_black = n_rand_black; a.n_rand_white = n_rand_white;
    a.weights = weights ? d_wt : nullptr;
    a.t_max = t_max; a.stride = n;
    a.traj_black = d_tb; a.traj_white = d_tw; a.traj_move = d_tm;
    a.nplies = d_np; a.final_black = d_fb; a.final_white = d_fw;
    rc = othello_playout(&a, s);
    if (rc) return rc;
    c->traj_black = d_tb; c->traj_white = d_tw; c->traj_move = d_tm; c->traj_stride = n; c->traj_t_max = t_max;

    OB_CUDA(cudaMemcpyAsync(nplies, d_np, n * 4, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaMemcpyAsync(final_black, d_fb, n * 8, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaMemcpyAsync(final_white, d_fw, n * 8, cudaMemcpyDeviceToHost, s));
    if (traj_black) {
        OB_CUDA(cudaMemcpyAsync(traj_black, d_tb, (size_t)(t_max + 1) * n * 8, cudaMemcpyDeviceToHost, s));
        OB_CUDA(cudaMemcpyAsync(traj_white, d_tw, (size_t)(t_max + 1) * n * 8, cudaMemcpyDeviceToHost, s));
        OB_CUDA(cudaMemcpyAsync(traj_move, d_tm, (size_t)t_max * n, cudaMemcpyDeviceToHost, s));
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
