// playout.cu -- GameRunner.play_a_game (game_runner.py:165-201) for millions of games at once.
//
// One thread owns one game from the first ply to the last: the position lives in four 32-bit
// registers (own/opp bitboards), a ply is ~385 integer instructions (fastboard.cuh) and touches memory
// only to append the position and the move to the SoA trajectory [t][game] (a warp writes
// 32 consecutive u64 = 256 B per array per ply).  The warp runs in lock step; games that end
// early idle until the longest game of the warp is over.  Nothing is read from HBM after the
// start position.
//
// The loop is the reference's: positions are recorded before every ply and once at the end
// (recorder.add, game_runner.py:170,159); the side to move asks go_for (game_runner.py:133-152),
// which may substitute a uniformly random move for the engine's, an engine with no move answers
// 'ps' and the pass is a ply of its own (board.py:194-195,203-208); the game stops when neither
// colour can move (board.py:57-58).
#include "playout_common.cuh"

using namespace ob;
using namespace obp;

namespace {

// Table look-ups on the idle LSU pipe instead of ALU-pipe instructions (A/B on B200, profiles/playout_variants_r02.txt):
// put() as four line look-ups (obf::flips_lut, +15 % over eight carry-chain rays), the last three rounds of the
// k-th-set-bit search from the byte table (+2.3 %)
constexpr bool kKthLut = true;

// the random engine (uniform over puttables()); the greedy engine lives in greedy.cu
struct Game {
    u64 own, opp;
    unsigned off;               // element index t * stride + game into the three trajectory arrays
    int t;
    bool passed;                // the previous ply was a pass
};

// One ply.  BLACK: 1 = Black moves, 0 = White moves (known at compile time when every game of the
// launch starts with the same colour: Black and White then alternate strictly, passes included),
// -1 = read `black_moves`.  Returns false when the game is over.
// TRAJ: 0 = no trajectory; 1 = trajectory with capacity checks; 2 = a trajectory that is known to fit, so the two
// checks per ply fall away: from the standard opening the first two plies are moves and a pass can only follow a move,
// so a game has at most 60 moves + 59 passes = 119 plies; with the pass that is taken back at the end of a game (below)
// the kernel writes position rows 0 .. plies + 1 <= 120 and move rows 0 .. plies <= 119, i.e. t_max >= 120 holds them.
// The three trajectory arrays are addressed with ONE 32-bit element index t * stride + game (a 32-bit add per
// ply instead of three 64-bit pointer increments), hence (t_max + 2) * stride < 2^32 (othello_playout checks).
template <int TRAJ, int BLACK>
__device__ __forceinline__ bool play_ply(Game &g, bool black_moves, u32 key, int t_max, int64_t stride, const Rays &rays,
                                         u64 *__restrict__ traj_black = nullptr, u64 *__restrict__ traj_white = nullptr,
                                         uint8_t *__restrict__ traj_move = nullptr)
{
    const bool bm = BLACK < 0 ? black_moves : (BLACK == 1);
    if (TRAJ == 2 || (TRAJ == 1 && g.t <= t_max)) {
        __stcs(traj_black + g.off, bm ? g.own : g.opp);
        __stcs(traj_white + g.off, bm ? g.opp : g.own);
    }
    const u64 legal = obf::legal_moves(g.own, g.opp);
    int move = OTHELLO_PASS;
    u64 f = 0, x = 0;
    if (legal == 0) {
        // is_game_over (board.py:57-58) needs the other colour's moves -- which the NEXT ply computes anyway, with the
        // whole warp instead of the few lanes that are here (that second move generation ran 6 % of the kernel's warp
        // instructions at 7 active lanes, profiles/playout_segments_r02.txt).  So the ply is played as a pass; if the
        // next one finds no move either, the game was over before the pass: it is taken back (rows past a game's
        // end are unspecified, and the rows up to it are what they would have been).
        if (g.passed) { g.t--; return false; }
        g.passed = true;
    } else {
        g.passed = false;
        const int n = __popcll(legal);
        const u32 r1 = rng_draw_fma(key, (u32)g.t, 1u, obf::kOpaqueOne);
        // go_for's substitution (game_runner.py:134-150) replaces the engine's move by a uniformly
        // random one; behind a random engine both draw the same k-th move from stream 1, so the
        // budgets n_rand_* do not change any game played by this kernel.
        move = obf::kth_set_bit<kKthLut>(legal, (int)rng_below(r1, (u32)n), rays);
        x = rays(obf::kRayDirs, move);
        f = obf::flips_lut(move, g.own, g.opp, rays, obf::kOpaqueOne);
    }
    if (TRAJ == 2 || (TRAJ == 1 && g.t < t_max)) __stcs(traj_move + g.off, (uint8_t)move);
    if (TRAJ) g.off += (unsigned)stride;
    // put_s: place, flip, nturn += 1, turn toggles (board.py:203-208)
    const u64 moved = g.own | f | x;
    g.own = g.opp & ~f;
    g.opp = moved;
    g.t++;
    return true;
}

// UNIFORM: no per-game turn0 -- every game starts with Black, so the loop is unrolled over the two colours
template <int TRAJ, bool UNIFORM>
// (128 threads x >= 10 CTAs per SM measured best on B200: 64/128/256 threads and 9..12 CTAs are within 2 %)
#if !defined(OB_PLAYOUT_CTAS)
#define OB_PLAYOUT_CTAS 10
#endif
__global__ void __launch_bounds__(kThreads, OB_PLAYOUT_CTAS) playout_kernel(const othello_playout_args a)
{
    __shared__ __align__(16) u64 ray_s[obf::kRayTable64];
    ob::fill_tables<kThreads>(ray_s);
    __syncthreads();
    const Rays rays = {ray_s};
    const int64_t gi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool live = gi < a.n_games;                  // lanes past the batch only take part in the totals
    int plies = 0, n_black = 0, n_white = 0;
    if (live) {
        const u64 b0 = a.black0 ? a.black0[gi] : OTHELLO_START_BLACK;
        const u64 w0 = a.white0 ? a.white0[gi] : OTHELLO_START_WHITE;
        bool black_moves = UNIFORM ? true : (a.turn0[gi] == OTHELLO_BLACK);
        Game g;
        g.own = black_moves ? b0 : w0; g.opp = black_moves ? w0 : b0;
        g.t = 0;
        g.passed = false;
        g.off = (unsigned)gi;
        const u32 key = rng_key(a.seed, a.gid0 + (u64)gi);
        const int t_max = a.t_max;
        const int64_t stride = a.stride;

        if (UNIFORM) {
            for (;;) {
                if (!play_ply<TRAJ, 1>(g, true, key, t_max, stride, rays, (u64 *)a.traj_black, (u64 *)a.traj_white, a.traj_move)) { black_moves = true; break; }
                if (!play_ply<TRAJ, 0>(g, false, key, t_max, stride, rays, (u64 *)a.traj_black, (u64 *)a.traj_white, a.traj_move)) { black_moves = false; break; }
            }
        } else {
            while (play_ply<TRAJ, -1>(g, black_moves, key, t_max, stride, rays, (u64 *)a.traj_black, (u64 *)a.traj_white, a.traj_move)) black_moves = !black_moves;
        }
        const u64 fb = black_moves ? g.own : g.opp, fw = black_moves ? g.opp : g.own;
        a.nplies[gi] = g.t;
        a.final_black[gi] = fb;
        a.final_white[gi] = fw;
        plies = g.t; n_black = __popcll(fb); n_white = __popcll(fw);
        if (a.summary) a.summary[gi] = game_summary(plies, n_black, n_white);
    }
    __syncwarp();
    add_totals(a.totals, live, plies, n_black, n_white);
}

int launch(const othello_playout_args &a, cudaStream_t s)
{
    const unsigned blocks = ob_blocks(a.n_games, kThreads);
    if (a.traj_black) {
        if ((int64_t)(a.t_max + 2) * a.stride >= (1ll << 32)) return OTHELLO_E_INVALID;   // 32-bit trajectory index
        const bool fits = a.black0 == nullptr && a.t_max >= 120;             // every game from the standard opening fits
        if (a.turn0) {
            if (fits) playout_kernel<2, false><<<blocks, kThreads, 0, s>>>(a);
            else playout_kernel<1, false><<<blocks, kThreads, 0, s>>>(a);
        } else {
            if (fits) playout_kernel<2, true><<<blocks, kThreads, 0, s>>>(a);
            else playout_kernel<1, true><<<blocks, kThreads, 0, s>>>(a);
        }
    } else {
        if (a.turn0) playout_kernel<0, false><<<blocks, kThreads, 0, s>>>(a);
        else playout_kernel<0, true><<<blocks, kThreads, 0, s>>>(a);
    }
    return ob_launch_status();
}

}  // namespace

extern "C" int othello_playout(const othello_playout_args *args, void *stream)
{
    OB_CHECK_ARGS(args != nullptr);
    const othello_playout_args &a = *args;
    OB_CHECK_ARGS(a.n_games >= 0);
    if (a.n_games == 0) return 0;
    OB_CHECK_ARGS(a.nplies && a.final_black && a.final_white);
    OB_CHECK_ARGS((a.black0 == nullptr) == (a.white0 == nullptr));
    OB_CHECK_ARGS(a.policy == OTHELLO_POLICY_RANDOM || a.policy == OTHELLO_POLICY_GREEDY);
    OB_CHECK_ARGS(a.policy_white == -1 || a.policy_white == OTHELLO_POLICY_RANDOM || a.policy_white == OTHELLO_POLICY_GREEDY);
    const int policy_white = a.policy_white == -1 ? a.policy : a.policy_white;
    OB_CHECK_ARGS(a.policy != OTHELLO_POLICY_GREEDY || a.weights != nullptr);
    OB_CHECK_ARGS(policy_white != OTHELLO_POLICY_GREEDY || a.weights != nullptr || a.weights_white != nullptr);
    OB_CHECK_ARGS(a.n_rand_black >= 0 && a.n_rand_white >= 0 && a.random_plies >= 0);
    if (a.traj_black || a.traj_white || a.traj_move)       // (a capacity of 0 plies records only the start position)
        OB_CHECK_ARGS(a.traj_black && a.traj_white && (a.traj_move || a.t_max == 0) && a.t_max >= 0 && a.stride >= a.n_games);
    const bool greedy = a.policy == OTHELLO_POLICY_GREEDY || policy_white == OTHELLO_POLICY_GREEDY;
    cudaStream_t s = (cudaStream_t)stream;
    if (greedy) return ob_launch_greedy(a, s);
    return launch(a, s);
}
