// rules.cu -- batched Board rules, features and linear evaluation on explicit positions.
//
// One thread per position; loads and stores are unit-stride (SoA arrays of u64 / u8 / i32),
// everything between them is register arithmetic from fastboard.cuh / bitboard.cuh.
#include "common.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 256;

// Board.puttables(piece) (board.py:46-52)
__global__ void __launch_bounds__(kThreads) legal_kernel(const u64 *__restrict__ own, const u64 *__restrict__ opp,
                                                         u64 *__restrict__ legal, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n) legal[i] = obf::legal_moves(own[i], opp[i]);
}

// Board.put(piece, x, y) flip set (board.py:161-174)
__global__ void __launch_bounds__(kThreads) flips_kernel(const u64 *__restrict__ own, const u64 *__restrict__ opp,
                                                         const uint8_t *__restrict__ square, u64 *__restrict__ flips,
                                                         int64_t n)
{
    __shared__ u64 ray_s[obf::kRayBasic64];
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const u64 a = own[i], b = opp[i];
    const unsigned s = square[i];
    u64 f = 0;
    if (s < 64) {
        const u64 x = 1ull << s;
        if (!((a | b) & x)) f = put_flips((int)s, a, b, rays);   // occupied => put returns 0 (board.py:162-163)
    }
    flips[i] = f;
}

// Board.put_s for the side to move (board.py:192-209) + is_game_over of the result (board.py:57-58)
__global__ void __launch_bounds__(kThreads) step_kernel(u64 *__restrict__ black, u64 *__restrict__ white,
                                                        uint8_t *__restrict__ turn, int32_t *__restrict__ nturn,
                                                        const uint8_t *__restrict__ move, u64 *__restrict__ flips_out,
                                                        int32_t *__restrict__ ret, uint8_t *__restrict__ flags,
                                                        int64_t n)
{
    __shared__ u64 ray_s[obf::kRayBasic64];
    fill_rays(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    u64 b = black[i], w = white[i];
    int t = turn[i];
    const unsigned mv = move[i];
    const bool black_moves = (t == OTHELLO_BLACK);
    u64 own = black_moves ? b : w, opp = black_moves ? w : b;
    u64 f = 0;
    int out = -1;
    if (mv == OTHELLO_PASS) {
        out = 0;                                              // 'ps'/'PS' is never validated (board.py:194-195)
    } else if (mv < 64) {
        const u64 x = 1ull << mv;
        if (!((own | opp) & x)) f = put_flips((int)mv, own, opp, rays);
        const int c = __popcll(f);
        if (c) { out = c; own |= f | x; opp &= ~f; }          // put() == 0 => -1, state untouched (board.py:199-201)
    }
    if (out >= 0) {
        t = black_moves ? OTHELLO_WHITE : OTHELLO_BLACK;
        b = black_moves ? own : opp;
        w = black_moves ? opp : own;
        black[i] = b; white[i] = w;
        turn[i] = (uint8_t)t;
        nturn[i] += 1;
    }
    if (flips_out) flips_out[i] = f;
    if (ret) ret[i] = out;
    if (flags) {
        const bool bm = (t == OTHELLO_BLACK);
        const u64 mover = bm ? b : w, other = bm ? w : b;
        uint8_t fl = 0;
        if (obf::legal_moves(mover, other) == 0)
            fl = obf::legal_moves(other, mover) == 0 ? OTHELLO_F_GAME_OVER : OTHELLO_F_MUST_PASS;
        flags[i] = fl;
    }
}

// n_black / n_white / n_empty (board.py:37-44)
__global__ void __launch_bounds__(kThreads) counts_kernel(const u64 *__restrict__ black, const u64 *__restrict__ white,
                                                          int32_t *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const u64 b = black[i], w = white[i];
    out[3 * i + 0] = __popcll(b);
    out[3 * i + 1] = __popcll(w & ~b);
    out[3 * i + 2] = 64 - __popcll(b | w);
}

// Board.mask_count(color, mask) (board.py:74-81)
__global__ void __launch_bounds__(kThreads) mask_count_kernel(const u64 *__restrict__ black, const u64 *__restrict__ white,
                                                              const uint8_t *__restrict__ color,
                                                              const u64 *__restrict__ mask, int32_t *__restrict__ out,
                                                              int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const u64 b = black[i], w = white[i];
    const int c = color[i];
    const u64 discs = c == OTHELLO_BLACK ? b : (c == OTHELLO_WHITE ? w : ~(b | w));
    out[i] = __popcll(discs & mask[i]);
}

// counts(a_book, side) (parameter_progress_position_moves_learn.py:5-17).  The [n][10] result is
// staged through shared memory so that the CTA writes its 10 KB with unit-stride stores (a thread
// writing its own 40-byte record would touch 32 sectors per store instruction).
__global__ void __launch_bounds__(kThreads) features_kernel(const u64 *__restrict__ black, const u64 *__restrict__ white,
                                                            const uint8_t *__restrict__ side, int32_t *__restrict__ out,
                                                            int64_t n)
{
    __shared__ int32_t tile[kThreads * OTHELLO_FEATURES];
    const int64_t base = (int64_t)blockIdx.x * kThreads;
    const int64_t i = base + threadIdx.x;
    if (i < n) {
        const u64 b = black[i], w = white[i];
        const bool is_black = side[i] == OTHELLO_BLACK;
        int f[10];
        features10(is_black ? b : w, is_black ? w : b, f);
#pragma unroll
        for (int k = 0; k < 10; k++) tile[threadIdx.x * OTHELLO_FEATURES + k] = f[k];
    }
    __syncthreads();
    const int64_t rows = (n - base < kThreads) ? n - base : kThreads;
    for (int j = threadIdx.x; j < rows * OTHELLO_FEATURES; j += kThreads) out[base * OTHELLO_FEATURES + j] = tile[j];
}

// linear phase-weighted evaluation; the 160-byte weight table is staged in shared memory
__global__ void __launch_bounds__(kThreads) eval_kernel(const u64 *__restrict__ black, const u64 *__restrict__ white,
                                                        const uint8_t *__restrict__ side, const float *__restrict__ weights,
                                                        float *__restrict__ out, int64_t n)
{
    __shared__ float w_s[OTHELLO_PHASES * OTHELLO_WEIGHTS];
    if (threadIdx.x < OTHELLO_PHASES * OTHELLO_WEIGHTS) w_s[threadIdx.x] = weights[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const u64 b = black[i], w = white[i];
    const bool is_black = side[i] == OTHELLO_BLACK;
    out[i] = eval_linear(is_black ? b : w, is_black ? w : b, w_s);
}

// Board.serialize_board (board.py:223-243): 64 characters per position, row-major, 'O' = Black,
// 'X' = White, '-' = empty: an HBM-bound codec (16 B read, 64 B written per position).
// The 8 bits of a rank are spread to 8 bytes with multiplies (SWAR, mostly FMA pipe): per nibble
// (v * 0x01010101) & 0x08040201 leaves bit k alone in byte k; +0x7F.. >> 7 turns it into 0 / 1.
__device__ __forceinline__ u32 spread4(u32 nibble)          // 4 bits -> 4 bytes of 0 / 1
{
    const u32 t = (nibble * 0x01010101u) & 0x08040201u;
    return ((t + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
}

// One thread converts one position (16 u32 of characters); the warp's 2 KB go through shared memory so
// that every store instruction writes 512 contiguous bytes (a thread storing its own 64-byte record
// would touch 32 half-filled sectors per instruction).
__global__ void __launch_bounds__(kThreads) serialize_kernel(const u64 *__restrict__ black, const u64 *__restrict__ white,
                                                             uint4 *__restrict__ out /* [n][4] */, int64_t n)
{
    __shared__ uint4 tile[kThreads / 32][32 * 4 + 4];          // per warp: 32 positions x 4 uint4 (+ padding)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = ((int64_t)blockIdx.x * kThreads + warp * 32);
    const int64_t i = warp_base + lane;
    if (i < n) {
        const u64 b = black[i], w = white[i] & ~b;
#pragma unroll
        for (int q = 0; q < 4; q++) {                         // two ranks (16 characters) per uint4
            u32 c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 nb = (u32)(b >> (16 * q + 4 * k)) & 0xFu, nw = (u32)(w >> (16 * q + 4 * k)) & 0xFu;
                // '-' = 0x2D, 'O' = 0x2D + 0x22, 'X' = 0x2D + 0x2B: no byte overflows, plain multiplies and adds
                c[k] = 0x2D2D2D2Du + spread4(nb) * 0x22u + spread4(nw) * 0x2Bu;
            }
            tile[warp][lane * 4 + q + (lane >> 3)] = make_uint4(c[0], c[1], c[2], c[3]);
        }
    }
    __syncwarp();
    const int64_t rows = n - warp_base;                         // positions of this warp that exist
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int j = r * 32 + lane;                            // uint4 index inside the warp's 2 KB
        if ((j >> 2) < rows) out[warp_base * 4 + j] = tile[warp][j + ((j >> 2) >> 3)];
    }
}

// Board.deserialize (board.py:253-262) of the 64-character board string; any character other than
// 'O' / 'X' is an empty square (turn_from_string, board.py:245-251).  One thread per position reads
// its 64 bytes as eight 8-byte words; per word the bytes equal to a character are found with the
// exact SWAR zero-byte test and their flags gathered into 8 bits with one multiply.
__device__ __forceinline__ u32 match4(u32 chars, u32 pattern)      // 4 bytes -> 4 bits (byte k == pattern byte)
{
    const u32 t = chars ^ pattern;
    const u32 z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;      // bit 7 of byte k set iff byte k == 0
    return (((z >> 7) * 0x01020408u) >> 24) & 0xFu;                           // gather the four flags: byte k -> bit k
}

__global__ void __launch_bounds__(kThreads) deserialize_kernel(const u64 *__restrict__ in /* [n][8] */,
                                                               u64 *__restrict__ black, u64 *__restrict__ white, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(in + i * 8);
    u64 b = 0, w = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {                             // 16 bytes = two ranks per load
        const uint4 v = src[q];
        const u32 word[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b |= (u64)match4(word[k], 0x4F4F4F4Fu) << (16 * q + 4 * k);      // squares 16q + 4k .. + 3, 'O'
            w |= (u64)match4(word[k], 0x58585858u) << (16 * q + 4 * k);      // 'X'
        }
    }
    black[i] = b; white[i] = w;
}

}  // namespace

extern "C" {

int othello_serialize_boards(const uint64_t *black, const uint64_t *white, char *out, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && out)) && ((uintptr_t)out & 15) == 0);
    if (n == 0) return 0;
    serialize_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)black, (const u64 *)white,
                                                                                  (uint4 *)out, n);
    return ob_launch_status();
}

int othello_deserialize_boards(const char *in, uint64_t *black, uint64_t *white, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && in)) && ((uintptr_t)in & 15) == 0);
    if (n == 0) return 0;
    deserialize_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)in, (u64 *)black,
                                                                                    (u64 *)white, n);
    return ob_launch_status();
}

int othello_legal(const uint64_t *own, const uint64_t *opp, uint64_t *legal, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (own && opp && legal)));
    if (n == 0) return 0;
    legal_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)own, (const u64 *)opp,
                                                                               (u64 *)legal, n);
    return ob_launch_status();
}

int othello_flips(const uint64_t *own, const uint64_t *opp, const uint8_t *square, uint64_t *flips, int64_t n,
                  void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (own && opp && square && flips)));
    if (n == 0) return 0;
    flips_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)own, (const u64 *)opp,
                                                                               square, (u64 *)flips, n);
    return ob_launch_status();
}

int othello_step(uint64_t *black, uint64_t *white, uint8_t *turn, int32_t *nturn, const uint8_t *move,
                 uint64_t *flips_out, int32_t *ret, uint8_t *flags, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && turn && nturn && move)));
    if (n == 0) return 0;
    step_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((u64 *)black, (u64 *)white, turn, nturn,
                                                                              move, (u64 *)flips_out, ret, flags, n);
    return ob_launch_status();
}

int othello_counts(const uint64_t *black, const uint64_t *white, int32_t *out, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && out)));
    if (n == 0) return 0;
    counts_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)black, (const u64 *)white,
                                                                                out, n);
    return ob_launch_status();
}

int othello_mask_count(const uint64_t *black, const uint64_t *white, const uint8_t *color, const uint64_t *mask,
                       int32_t *out, int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && color && mask && out)));
    if (n == 0) return 0;
    mask_count_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        (const u64 *)black, (const u64 *)white, color, (const u64 *)mask, out, n);
    return ob_launch_status();
}

int othello_features(const uint64_t *black, const uint64_t *white, const uint8_t *side, int32_t *out, int64_t n,
                     void *stream)
{
    OB_CHECK_ARGS(n >= 0 && (n == 0 || (black && white && side && out)));
    if (n == 0) return 0;
    features_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)black,
                                                                                  (const u64 *)white, side, out, n);
    return ob_launch_status();
}

int othello_eval(const uint64_t *black, const uint64_t *white, const uint8_t *side, const float *weights, float *out,
                 int64_t n, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && weights && (n == 0 || (black && white && side && out)));
    if (n == 0) return 0;
    eval_kernel<<<ob_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const u64 *)black, (const u64 *)white,
                                                                              side, weights, out, n);
    return ob_launch_status();
}

}  // extern "C"
