// perft.cu -- legal-move enumeration (BASELINE config 2): node counts to a given depth.
//
// Counting convention: a pass is one ply, a position where neither side
// can move (Board.is_game_over, board.py:57-58) is one leaf whatever depth is left.
//
// Plan: breadth-first expansion on the device while the frontier is small (children are
// appended through one warp-aggregated atomic per warp), then one thread per frontier node
// counts its remaining subtree depth-first entirely in registers, with the last ply counted in
// bulk as popc(legal).  Positions are stored mover-relative (own, opp).
#include "common.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 128;
constexpr int kMaxDfs = 6;                  // deepest register-resident DFS instantiated
constexpr int64_t kFrontierTarget = 296 * 1024;   // stop expanding once ~2048 nodes per SM exist

struct Ctl {
    unsigned long long n_out;               // children appended so far
    unsigned long long leaves;              // nodes counted
    unsigned int overflow;                  // frontier did not fit
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads) expand_kernel(const u64 *__restrict__ in_own, const u64 *__restrict__ in_opp,
                                                          int64_t n_in, u64 *__restrict__ out_own,
                                                          u64 *__restrict__ out_opp, int64_t cap, Ctl *ctl)
{
    __shared__ u64 ray_s[obf::kRayDirs * 64];
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    u64 own = 0, opp = 0, legal = 0;
    int cnt = 0;
    unsigned long long leaf = 0;
    bool pass = false;
    if (i < n_in) {
        own = in_own[i]; opp = in_opp[i];
        legal = obf::legal_moves(own, opp);
        if (legal) cnt = __popcll(legal);
        else if (obf::legal_moves(opp, own)) { pass = true; cnt = 1; }
        else leaf = 1;                                       // game over: one leaf
    }
    // warp-exclusive prefix of cnt, one atomic per warp
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(&ctl->n_out, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    leaf = warp_sum(leaf);
    if (lane == 0 && leaf) atomicAdd(&ctl->leaves, leaf);
    if (!cnt) return;
    int64_t o = (int64_t)base + incl - cnt;
    if (o + cnt > cap) { ctl->overflow = 1; return; }
    if (pass) { out_own[o] = opp; out_opp[o] = own; return; }
    const u64 own_r = obf::rev64(own), opp_r = obf::rev64(opp);
    for (u64 rem = legal; rem; rem &= rem - 1, o++) {
        const int sq = __ffsll((long long)rem) - 1;
        const u64 x = 1ull << sq;
        const u64 f = obf::flips_for(sq, own, opp, own_r, opp_r, rays);
        out_own[o] = opp & ~f;                               // the child is seen by its own mover
        out_opp[o] = own | f | x;
    }
}

template <int R> struct Dfs {
    static __device__ unsigned long long run(u64 own, u64 opp, const Rays &rays)
    {
        const u64 legal = obf::legal_moves(own, opp);
        if (!legal) return obf::legal_moves(opp, own) ? Dfs<R - 1>::run(opp, own, rays) : 1ull;
        unsigned long long total = 0;
        const u64 own_r = obf::rev64(own), opp_r = obf::rev64(opp);
        for (u64 rem = legal; rem; rem &= rem - 1) {
            const int sq = __ffsll((long long)rem) - 1;
            const u64 x = 1ull << sq;
            const u64 f = obf::flips_for(sq, own, opp, own_r, opp_r, rays);
            total += Dfs<R - 1>::run(opp & ~f, own | f | x, rays);
        }
        return total;
    }
};
template <> struct Dfs<1> {
    // one ply left: every move, or the pass, or the game-over node itself, is exactly one leaf
    static __device__ unsigned long long run(u64 own, u64 opp, const Rays &)
    {
        const u64 legal = obf::legal_moves(own, opp);
        return legal ? (unsigned long long)__popcll(legal) : 1ull;
    }
};

template <int R>
__global__ void __launch_bounds__(kThreads) dfs_kernel(const u64 *__restrict__ own, const u64 *__restrict__ opp,
                                                       int64_t n, Ctl *ctl)
{
    __shared__ u64 ray_s[obf::kRayDirs * 64];
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    unsigned long long c = 0;
    if (i < n) c = Dfs<R>::run(own[i], opp[i], rays);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->leaves, c);
}

int launch_dfs(int r, const u64 *own, const u64 *opp, int64_t n, Ctl *ctl, cudaStream_t s)
{
    const unsigned blocks = ob_blocks(n, kThreads);
    switch (r) {
    case 1: dfs_kernel<1><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    case 2: dfs_kernel<2><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    case 3: dfs_kernel<3><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    case 4: dfs_kernel<4><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    case 5: dfs_kernel<5><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    case 6: dfs_kernel<6><<<blocks, kThreads, 0, s>>>(own, opp, n, ctl); break;
    default: return OTHELLO_E_INVALID;
    }
    return ob_launch_status();
}

constexpr int64_t kCtlBytes = 256;
constexpr int64_t kDefaultNodes = 4 * 1024 * 1024;

}  // namespace

extern "C" int64_t othello_perft_workspace_bytes(int depth)
{
    (void)depth;
    return kCtlBytes + 2 * kDefaultNodes * 2 * (int64_t)sizeof(u64);
}

extern "C" int othello_perft(uint64_t black, uint64_t white, int turn, int depth, void *workspace,
                             int64_t workspace_bytes, uint64_t *result, void *stream)
{
    OB_CHECK_ARGS(result && depth >= 0 && depth <= 60 && (turn == OTHELLO_BLACK || turn == OTHELLO_WHITE));
    if (depth == 0) { *result = 1; return 0; }
    OB_CHECK_ARGS(workspace != nullptr);
    const int64_t cap = (workspace_bytes - kCtlBytes) / (4 * (int64_t)sizeof(u64));
    if (cap < 64) return OTHELLO_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    Ctl *ctl = (Ctl *)workspace;
    u64 *buf = (u64 *)((char *)workspace + kCtlBytes);
    u64 *own[2] = {buf, buf + 2 * cap}, *opp[2] = {buf + cap, buf + 3 * cap};

    const u64 root[2] = {turn == OTHELLO_BLACK ? (u64)black : (u64)white, turn == OTHELLO_BLACK ? (u64)white : (u64)black};
    Ctl h = {0ull, 0ull, 0u, 0u};
    OB_CUDA(cudaMemcpyAsync(ctl, &h, sizeof h, cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(own[0], &root[0], sizeof(u64), cudaMemcpyHostToDevice, s));
    OB_CUDA(cudaMemcpyAsync(opp[0], &root[1], sizeof(u64), cudaMemcpyHostToDevice, s));

    int64_t n = 1;
    int cur = 0, left = depth;
    // breadth-first while the frontier is small and more than one ply is left
    while (n > 0 && left > 1 && (n < kFrontierTarget || left > kMaxDfs)) {
        const unsigned long long zero = 0;
        OB_CUDA(cudaMemcpyAsync(&ctl->n_out, &zero, sizeof zero, cudaMemcpyHostToDevice, s));
        expand_kernel<<<ob_blocks(n, kThreads), kThreads, 0, s>>>(own[cur], opp[cur], n, own[cur ^ 1], opp[cur ^ 1], cap,
                                                                  ctl);
        OB_CUDA(cudaGetLastError());
        OB_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, s));
        OB_CUDA(cudaStreamSynchronize(s));
        if (h.overflow) return OTHELLO_E_WORKSPACE;          // the next level does not fit: caller must give more scratch
        n = (int64_t)h.n_out;
        cur ^= 1;
        left -= 1;
    }
    if (n > 0) {
        int rc = launch_dfs(left, own[cur], opp[cur], n, ctl, s);
        if (rc) return rc;
    }
    OB_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, s));
    OB_CUDA(cudaStreamSynchronize(s));
    *result = h.leaves;
    return 0;
}
