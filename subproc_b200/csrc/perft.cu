// perft.cu -- legal-move enumeration (BASELINE config 2): node counts to a given depth.
//
// Counting convention: a pass is one ply, a position where neither side
// can move (Board.is_game_over, board.py:57-58) is one leaf whatever depth is left.
//
// Plan: breadth-first expansion on the device while the frontier is small (children are
// appended through one warp-aggregated atomic per warp), then the frontier nodes are counted
// depth-first entirely in registers, with the last ply counted in bulk as popc(legal).
// Positions are stored mover-relative (own, opp).
//
// The host never looks at the frontier: a call enqueues a FIXED sequence of launches
// (depth - 1 expansion steps + one depth-first count) and every kernel reads the frontier size, the
// buffer in use and the plies left from a control block in device memory.  An expansion step that
// is not needed (frontier already large enough, or one ply left) returns at once; the last CTA of a
// step that did expand publishes the new frontier.  No host round trip per level, so the sequence is
// stream-ordered, capturable in a CUDA graph, and its result can stay on the device (multi-GPU: every
// rank counts the frontier nodes whose position hashes to its part and the ranks add their counts).
#include "common.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 128;
constexpr int kMaxDfs = 6;                  // deepest register-resident DFS instantiated
constexpr unsigned long long kFrontierTarget = 296 * 1024;   // stop expanding once ~2048 nodes per SM exist
constexpr int kExpandBlocks = 148 * 8;
constexpr int kDfsBlocks = 148 * 16;

struct Ctl {
    unsigned long long n;                   // nodes in the current frontier
    unsigned long long n_out;               // children appended by the running expansion step
    unsigned long long leaves;              // nodes counted so far
    unsigned long long next;                // work cursor of the depth-first count
    unsigned int cur;                       // frontier buffer in use (0 / 1)
    int left;                               // plies left below the current frontier
    unsigned int overflow;                  // the next frontier did not fit / too deep for the DFS
    unsigned int arrived;                   // CTAs that finished the running expansion step
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void init_kernel(Ctl *ctl, u64 *own0, u64 *opp0, u64 own, u64 opp, int depth)
{
    ctl->n = 1; ctl->n_out = 0; ctl->leaves = 0; ctl->next = 0; ctl->cur = 0; ctl->left = depth;
    ctl->overflow = 0; ctl->arrived = 0;
    own0[0] = own; opp0[0] = opp;
}

// one breadth-first level, if one is still wanted: frontier[cur] -> frontier[cur ^ 1]
__global__ void __launch_bounds__(kThreads) expand_kernel(u64 *buf, int64_t cap, Ctl *ctl, int count_leaves)
{
    __shared__ u64 ray_s[obf::kRayBasic64];
    __shared__ bool s_last;
    // every thread reads the same control words; they are only written by the last CTA of a step
    const unsigned long long n_in = ctl->n;
    const int left = ctl->left;
    const unsigned cur = ctl->cur;
    if (left <= 1 || ctl->overflow || (n_in >= kFrontierTarget && left <= kMaxDfs)) return;   // nothing to do (uniform)
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const u64 *in_own = buf + (size_t)(2 * cur) * cap, *in_opp = in_own + cap;
    u64 *out_own = buf + (size_t)(2 * (cur ^ 1)) * cap, *out_opp = out_own + cap;
    const int lane = threadIdx.x & 31;
    const unsigned long long span = (unsigned long long)gridDim.x * kThreads;
    const unsigned long long rounds = (n_in + span - 1) / span;
    for (unsigned long long r = 0; r < rounds; r++) {
        const unsigned long long i = r * span + (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
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
        if (lane == 0 && leaf && count_leaves) atomicAdd(&ctl->leaves, leaf);
        if (!cnt) continue;
        int64_t o = (int64_t)base + incl - cnt;
        if (o + cnt > cap) { ctl->overflow = 1; continue; }
        if (pass) { out_own[o] = opp; out_opp[o] = own; continue; }
        const u64 own_r = obf::rev64(own), opp_r = obf::rev64(opp);
        for (u64 rem = legal; rem; rem &= rem - 1, o++) {
            const int sq = __ffsll((long long)rem) - 1;
            const u64 x = 1ull << sq;
            const u64 f = obf::flips_for(sq, own, opp, own_r, opp_r, rays);
            out_own[o] = opp & ~f;                               // the child is seen by its own mover
            out_opp[o] = own | f | x;
        }
    }
    // the last CTA to get here publishes the new frontier for the next launch
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->arrived, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long n_out = atomicAdd(&ctl->n_out, 0ull);
        ctl->n = n_out; ctl->n_out = 0; ctl->cur = cur ^ 1; ctl->left = left - 1; ctl->arrived = 0;
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

// which of the nparts shares counts the subtree of a frontier node: a function of the POSITION, because the
// order in which a breadth-first expansion appends nodes differs from run to run and from GPU to GPU
__device__ __forceinline__ unsigned part_of(u64 own, u64 opp, unsigned nparts)
{
    u64 h = own * 0x9E3779B97F4A7C15ull ^ opp * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return (unsigned)(((h & 0xffffffffull) * nparts) >> 32);
}

// count the subtrees of this part's frontier nodes; warps fetch 32 nodes at a time
template <int R>
__device__ __forceinline__ void dfs_all(const u64 *own, const u64 *opp, unsigned long long n, unsigned part,
                                        unsigned nparts, Ctl *ctl, const Rays &rays)
{
    const int lane = threadIdx.x & 31;
    unsigned long long c = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->next, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const unsigned long long i = base + lane;
        if (i < n) {
            const u64 o = own[i], p = opp[i];
            if (nparts == 1 || part_of(o, p, nparts) == part) c += Dfs<R>::run(o, p, rays);
        }
    }
    c = warp_sum(c);
    if (lane == 0 && c) atomicAdd(&ctl->leaves, c);
}

__global__ void __launch_bounds__(kThreads) dfs_kernel(const u64 *buf, int64_t cap, Ctl *ctl, unsigned part, unsigned nparts,
                                                       unsigned long long *result)
{
    __shared__ u64 ray_s[obf::kRayBasic64];
    __shared__ bool s_last;
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const unsigned long long n = ctl->n;
    const int left = ctl->left;
    const u64 *own = buf + (size_t)(2 * ctl->cur) * cap, *opp = own + cap;
    if (!ctl->overflow) {
        switch (left) {
        case 1: dfs_all<1>(own, opp, n, part, nparts, ctl, rays); break;
        case 2: dfs_all<2>(own, opp, n, part, nparts, ctl, rays); break;
        case 3: dfs_all<3>(own, opp, n, part, nparts, ctl, rays); break;
        case 4: dfs_all<4>(own, opp, n, part, nparts, ctl, rays); break;
        case 5: dfs_all<5>(own, opp, n, part, nparts, ctl, rays); break;
        case 6: dfs_all<6>(own, opp, n, part, nparts, ctl, rays); break;
        default: if (n > 0 && threadIdx.x == 0) ctl->overflow = 1; break;     // (cannot happen: see the host loop)
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->arrived, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        result[0] = atomicAdd(&ctl->leaves, 0ull);
        result[1] = ctl->overflow;
    }
}

constexpr int64_t kCtlBytes = 256;
constexpr int64_t kMaxNodes = 4 * 1024 * 1024;

// the frontier after k levels has at most 33^k nodes (a position has at most 33 legal moves)
int64_t frontier_cap(int depth)
{
    int64_t cap = 64;
    for (int k = 1; k < depth && cap < kMaxNodes; k++) cap *= 33;
    return cap < kMaxNodes ? cap : kMaxNodes;
}

}  // namespace

extern "C" int64_t othello_perft_workspace_bytes(int depth)
{
    return kCtlBytes + 16 + 4 * frontier_cap(depth) * (int64_t)sizeof(u64);
}

extern "C" int othello_perft_async(uint64_t black, uint64_t white, int turn, int depth, int part, int nparts,
                                   void *workspace, int64_t workspace_bytes, uint64_t *result, void *stream)
{
    OB_CHECK_ARGS(result && depth >= 1 && depth <= 60 && (turn == OTHELLO_BLACK || turn == OTHELLO_WHITE));
    OB_CHECK_ARGS(nparts >= 1 && part >= 0 && part < nparts && workspace != nullptr);
    const int64_t cap = (workspace_bytes - kCtlBytes) / (4 * (int64_t)sizeof(u64));
    if (cap < 64) return OTHELLO_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    Ctl *ctl = (Ctl *)workspace;
    u64 *buf = (u64 *)((char *)workspace + kCtlBytes);          // [buffer 0 own, opp][buffer 1 own, opp], cap nodes each
    const u64 own = turn == OTHELLO_BLACK ? (u64)black : (u64)white, opp = turn == OTHELLO_BLACK ? (u64)white : (u64)black;
    init_kernel<<<1, 1, 0, s>>>(ctl, buf, buf + cap, own, opp, depth);
    // at most depth - 1 breadth-first levels are ever wanted; each launch decides for itself on the device.
    // Game-over leaves met while expanding are counted once, by part 0.
    for (int level = 0; level < depth - 1; level++)
        expand_kernel<<<kExpandBlocks, kThreads, 0, s>>>(buf, cap, ctl, part == 0 ? 1 : 0);
    dfs_kernel<<<kDfsBlocks, kThreads, 0, s>>>(buf, cap, ctl, (unsigned)part, (unsigned)nparts, (unsigned long long *)result);
    return ob_launch_status();
}

extern "C" int othello_perft(uint64_t black, uint64_t white, int turn, int depth, void *workspace,
                             int64_t workspace_bytes, uint64_t *result, void *stream)
{
    OB_CHECK_ARGS(result && depth >= 0 && depth <= 60 && (turn == OTHELLO_BLACK || turn == OTHELLO_WHITE));
    if (depth == 0) { *result = 1; return 0; }
    OB_CHECK_ARGS(workspace != nullptr);
    if (workspace_bytes < kCtlBytes + 16 + 4 * 64 * (int64_t)sizeof(u64)) return OTHELLO_E_WORKSPACE;
    // the last 16 bytes of the workspace receive {nodes, overflow}
    uint64_t *d_res = (uint64_t *)((char *)workspace + ((workspace_bytes - 16) & ~(int64_t)7));
    int rc = othello_perft_async(black, white, turn, depth, 0, 1, workspace, workspace_bytes - 16, d_res, stream);
    if (rc) return rc;
    uint64_t h[2] = {0, 0};
    OB_CUDA(cudaMemcpyAsync(h, d_res, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (h[1]) return OTHELLO_E_WORKSPACE;                    // a frontier did not fit: the caller must give more scratch
    *result = h[0];
    return 0;
}
