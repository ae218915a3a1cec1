// bitboard.cuh -- register-resident Othello rules on a pair of 64-bit bitboards (sm_100a).
//
// Replaces the list-of-lists ray walks of the reference (board.py:124-174: hands_for_direc,
// is_puttable_at, put) with 8-direction parallel-prefix (Kogge-Stone) shift-and-mask floods.
// Bit s = x + 8*y (board.py:79): +1 = one file to the right, +8 = one rank down the board.
//
// Everything here is __device__ __forceinline__ integer ALU work: no memory traffic at all.
#pragma once
#include <stdint.h>

namespace ob {

typedef unsigned long long u64;
typedef unsigned int u32;

// files b..g: a disc that a horizontal/diagonal flood may pass THROUGH (a run can never
// continue across the a/h edge, hands_for_direc stops at is_within_board, board.py:131-137)
__device__ constexpr u64 kInner = 0x7E7E7E7E7E7E7E7Eull;

// square-class masks a..h (parameter_progress_position_moves_learn.py:9-16)
__device__ constexpr u64 kClassMask[8] = {
    0x8100000000000081ull, 0x4281000000008142ull, 0x0042000000004200ull, 0x2400810000810024ull,
    0x1800008181000018ull, 0x003C424242423C00ull, 0x0000240000240000ull, 0x0000183C3C180000ull};

template <int D> __device__ __forceinline__ u64 up(u64 v) { return v << D; }
template <int D> __device__ __forceinline__ u64 dn(u64 v) { return v >> D; }

// One axis (two opposite directions).  m = opponent discs a flood may pass through.
// Returns the squares just beyond a run of opponent discs that starts next to an own disc.
template <int D> __device__ __forceinline__ u64 axis_moves(u64 own, u64 m)
{
    u64 fu = m & up<D>(own), fd = m & dn<D>(own);
    fu |= m & up<D>(fu);
    fd |= m & dn<D>(fd);
    const u64 pu = m & up<D>(m), pd = dn<D>(pu);        // pairs of adjacent opponent discs
    fu |= pu & up<2 * D>(fu);
    fd |= pd & dn<2 * D>(fd);
    fu |= pu & up<2 * D>(fu);
    fd |= pd & dn<2 * D>(fd);
    return up<D>(fu) | dn<D>(fd);
}

// Board.puttables(piece) as a mask (board.py:46-52).
__device__ __forceinline__ u64 legal_moves(u64 own, u64 opp)
{
    const u64 m = opp & kInner;
    u64 r = axis_moves<1>(own, m);
    r |= axis_moves<8>(own, opp);
    r |= axis_moves<7>(own, m);
    r |= axis_moves<9>(own, m);
    return r & ~(own | opp);
}

// One axis of put(): the opponent runs starting next to square bit `x` that end on an own disc.
template <int D> __device__ __forceinline__ u64 axis_flips(u64 x, u64 own, u64 m)
{
    u64 fu = m & up<D>(x), fd = m & dn<D>(x);
    fu |= m & up<D>(fu);
    fd |= m & dn<D>(fd);
    const u64 pu = m & up<D>(m), pd = dn<D>(pu);
    fu |= pu & up<2 * D>(fu);
    fd |= pd & dn<2 * D>(fd);
    fu |= pu & up<2 * D>(fu);
    fd |= pd & dn<2 * D>(fd);
    // hands_for_direc keeps a run only when an own disc closes it (board.py:134-138)
    const u64 cu = own & up<D>(fu), cd = own & dn<D>(fd);
    return (cu ? fu : 0ull) | (cd ? fd : 0ull);
}

// Discs flipped by placing an `own` disc on the EMPTY square bit x (board.py:161-174).
// The caller checks emptiness (put returns 0 on an occupied square, board.py:162-163).
__device__ __forceinline__ u64 flips_for(u64 x, u64 own, u64 opp)
{
    const u64 m = opp & kInner;
    return axis_flips<1>(x, own, m) | axis_flips<8>(x, own, opp) | axis_flips<7>(x, own, m) |
           axis_flips<9>(x, own, m);
}

// index of the k-th (0-based, ascending) set bit of a non-empty mask; k < popc(mask).
// This is puttables()[k] (ascending s = x + 8*y, board.py:48-51).
__device__ __forceinline__ int kth_set_bit(u64 mask, int k)
{
    u32 lo = (u32)mask, hi = (u32)(mask >> 32);
    int c = __popc(lo);
    int base = 0;
    u32 v = lo;
    if (k >= c) { k -= c; v = hi; base = 32; }
    c = __popc(v & 0xFFFFu);
    if (k >= c) { k -= c; v >>= 16; base += 16; }
    c = __popc(v & 0xFFu);
    if (k >= c) { k -= c; v >>= 8; base += 8; }
    c = __popc(v & 0xFu);
    if (k >= c) { k -= c; v >>= 4; base += 4; }
    c = __popc(v & 0x3u);
    if (k >= c) { k -= c; v >>= 2; base += 2; }
    c = (int)(v & 1u);
    if (k >= c) { base += 1; }
    return base;
}

// ---- counter-based RNG (DESIGN.md "RNG"; oracle restatement: orc_rng_*) ----------------------
__device__ __forceinline__ u32 fmix32(u32 h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ u32 rng_key(u64 seed, u64 gid)
{
    u32 h = fmix32((u32)seed ^ 0x9E3779B9u);
    h = fmix32(h ^ (u32)(seed >> 32));
    h = fmix32(h ^ (u32)gid);
    h = fmix32(h ^ (u32)(gid >> 32));
    return h;
}
__device__ __forceinline__ u32 rng_draw(u32 key, u32 ply, u32 stream)
{
    return fmix32(key + ply * 0x9E3779B9u + stream * 0x632BE5ABu);
}
__device__ __forceinline__ u32 rng_below(u32 r, u32 n) { return __umulhi(r, n); }

// ---- features / evaluation -------------------------------------------------------------------
// phase row of a disc count: shards (0,16),(17,32),(33,48),(49,64) inclusive
// (progress_position_moves_learn.py:75,112-113)
__device__ __forceinline__ int phase_row(int discs)
{
    int r = (discs + 15) >> 4;
    r -= 1;
    return r < 0 ? 0 : (r > 3 ? 3 : r);
}

// counts(a_book, side) (parameter_progress_position_moves_learn.py:5-17) for side = `own`
__device__ __forceinline__ void features10(u64 own, u64 opp, int f[10])
{
    f[0] = __popcll(own | opp);
    f[1] = __popcll(legal_moves(own, opp));
#pragma unroll
    for (int k = 0; k < 8; k++) f[2 + k] = __popcll(own & kClassMask[k]);
}

// w[phase] . (mobility, a..h) + w[phase][9]; w = [4][10] floats (shared or global)
__device__ __forceinline__ float eval_linear(u64 own, u64 opp, const float *__restrict__ w)
{
    const int discs = __popcll(own | opp);
    const float *row = w + 10 * phase_row(discs);
    float acc = row[9];
    acc = fmaf(row[0], (float)__popcll(legal_moves(own, opp)), acc);
#pragma unroll
    for (int k = 0; k < 8; k++) acc = fmaf(row[1 + k], (float)__popcll(own & kClassMask[k]), acc);
    return acc;
}

}  // namespace ob
