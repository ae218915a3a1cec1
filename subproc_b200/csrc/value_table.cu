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
// at its position in the reference's order; (2) a hand-written STABLE radix sort groups equal keys and
// keeps the order inside a key; (3) every run of equal keys is walked sequentially by one thread with the
// reference's exact operation order (two roundings for the products, one for the sum -- no FMA
// contraction), starting from the value already in the table, which is an open-addressing hash table in
// HBM (key -> dense index) next to dense key / value arrays: no merge, no re-sort of the table per batch.
#include "common.cuh"
#include "fastboard.cuh"

using namespace ob;

namespace {

constexpr int kThreads = 256;

// counts() 10-tuple packed into 43 bits: discs(7) mobility(6) a(3) b(4) c(3) d(4) e(4) f(5) g(3) h(4)
__device__ __forceinline__ u64 pack_features(u64 side, u64 other)
{
    u64 k = (u64)__popcll(side | other);
    k = (k << 6) | (u64)obf::mobility(side, other);
    k = (k << 3) | (u64)class_count(side, 0);
    k = (k << 4) | (u64)class_count(side, 1);
    k = (k << 3) | (u64)class_count(side, 2);
    k = (k << 4) | (u64)class_count(side, 3);
    k = (k << 4) | (u64)class_count(side, 4);
    k = (k << 5) | (u64)class_count(side, 5);
    k = (k << 3) | (u64)class_count(side, 6);
    k = (k << 4) | (u64)class_count(side, 7);
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

// ---- stable LSD radix sort of (key, value) records ------------------------------------------------
// Grouping the records of one key while keeping their update order is a STABLE sort by key.  43-bit
// keys = 6 passes of 8 bits; every pass is histogram -> per-digit row scan -> stable scatter.  A tile is
// 4096 consecutive records; inside a tile warp w owns records [512 w, 512 (w + 1)) in 16 coalesced rows,
// ranks them per digit with match.any (lanes holding the same digit find each other, the lowest lane
// advances the warp's running counter), and the 8 warps are ordered through a per-digit prefix.
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRows = 16;
constexpr int kTile = kSortThreads * kSortRows;
constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ unsigned owner_of(u64 key, unsigned world)
{
    const u64 h = key * 0x9E3779B97F4A7C15ull;
    return (unsigned)(((h >> 32) * (u64)world) >> 32);
}
// MODE 0: 8 bits of the key at `shift`; MODE 1: the rank that owns the key in a table sharded over `world` GPUs
template <int MODE> __device__ __forceinline__ unsigned digit_of(u64 key, int shift, unsigned world)
{
    return MODE == 0 ? (unsigned)(key >> shift) & 255u : owner_of(key, world);
}

template <int MODE>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const u64 *__restrict__ keys, int64_t n, int shift,
                                                                 unsigned world, unsigned *__restrict__ counts, int64_t tiles)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll 4
    for (int r = 0; r < kSortRows; r++) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[digit_of<MODE>(keys[i], shift, world)], 1u);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];          // digit-major
}

// exclusive scan of 256 values held one per thread (CTA of 256 threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ unsigned block_scan_256(unsigned v, unsigned *s_warp, unsigned *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kAll, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
        const unsigned c = s_warp[w];
        if (w < warp) before += c;
        sum += c;
    }
    __syncthreads();
    if (total) *total = sum;
    return before + incl - v;
}

// one CTA per digit: exclusive scan of its row of per-tile counts (in place), row total -> totals[digit]
__global__ void __launch_bounds__(kSortThreads) sort_rowscan_kernel(unsigned *__restrict__ counts, int64_t tiles,
                                                                    unsigned *__restrict__ totals)
{
    __shared__ unsigned s_warp[kSortWarps];
    unsigned *row = counts + (int64_t)blockIdx.x * tiles;
    unsigned carry = 0;
    for (int64_t base = 0; base < tiles; base += kSortThreads * 8) {
        unsigned v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t i = base + (int64_t)threadIdx.x * 8 + j;
            v[j] = i < tiles ? row[i] : 0u;
            sum += v[j];
        }
        unsigned total;
        unsigned at = carry + block_scan_256(sum, s_warp, &total);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t i = base + (int64_t)threadIdx.x * 8 + j;
            if (i < tiles) row[i] = at;
            at += v[j];
        }
        carry += total;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

template <int MODE>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const u64 *__restrict__ keys_in,
                                                                    const double *__restrict__ vals_in,
                                                                    u64 *__restrict__ keys_out, double *__restrict__ vals_out,
                                                                    int64_t n, int shift, unsigned world,
                                                                    const unsigned *__restrict__ counts,
                                                                    const unsigned *__restrict__ totals, int64_t tiles)
{
    __shared__ unsigned s_warp[kSortWarps];
    __shared__ unsigned whist[kSortWarps][256];       // per warp and digit: count, then running output position
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // where this tile's records of digit d start: all smaller digits + this digit's records of earlier tiles
    const unsigned dbase = block_scan_256(totals[threadIdx.x], s_warp, nullptr) +
                           counts[(int64_t)threadIdx.x * tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) whist[w][threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile + warp * (kSortRows * 32) + lane;
    u64 key[kSortRows];
#pragma unroll
    for (int r = 0; r < kSortRows; r++) {
        const int64_t i = base + r * 32;
        key[r] = i < n ? keys_in[i] : 0ull;
        if (i < n) atomicAdd(&whist[warp][digit_of<MODE>(key[r], shift, world)], 1u);
    }
    __syncthreads();
    {
        unsigned at = dbase;                          // thread d orders the warps' runs of digit d
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const unsigned c = whist[w][threadIdx.x];
            whist[w][threadIdx.x] = at;
            at += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRows; r++) {
        const int64_t i = base + r * 32;
        const bool valid = i < n;
        const unsigned vmask = __ballot_sync(kAll, valid);
        if (valid) {
            const unsigned d = digit_of<MODE>(key[r], shift, world);
            const unsigned peers = __match_any_sync(vmask, d);
            const int leader = __ffs(peers) - 1;
            unsigned at = 0;
            if (lane == leader) { at = whist[warp][d]; whist[warp][d] = at + __popc(peers); }
            at = __shfl_sync(peers, at, leader) + __popc(peers & ((1u << lane) - 1u));
            keys_out[at] = key[r];
            vals_out[at] = vals_in[i];
        }
        __syncwarp();
    }
}

__global__ void export_totals_kernel(const unsigned *__restrict__ totals, int64_t *__restrict__ out, int count)
{
    if (threadIdx.x < count) out[threadIdx.x] = (int64_t)totals[threadIdx.x];
}

// ---- the table: open addressing key -> dense index, dense key / value arrays --------------------------
// slot_keys[h] = key (0 = empty: a key's top bits are its disc count >= 4), slot_idx[h] = index into the
// dense arrays.  Linear probing from the high bits of a multiplicative hash; capacity is a power of two.
__device__ __forceinline__ u64 slot_of(u64 key, int log2cap) { return (key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap); }

__device__ __forceinline__ int table_find(u64 key, const u64 *slot_keys, const int32_t *slot_idx, int log2cap)
{
    const u64 mask = (1ull << log2cap) - 1ull;
    for (u64 h = slot_of(key, log2cap);; h = (h + 1) & mask) {
        const u64 cur = slot_keys[h];
        if (cur == key) return slot_idx[h];
        if (cur == 0ull) return -1;
    }
}

// ---- applying a sorted batch --------------------------------------------------------------------------
// probe : which records start a run of equal keys (a "head"), which heads are new keys, which runs are long
// scan  : new heads per tile -> exclusive prefix (+ the total, which sizes the table on the host)
// assign: new keys get dense indices in sorted-key order (the table's layout does not depend on scheduling),
//         are entered into the hash, and start from 0 (`if not exists: set(key, 0)`, :52-53)
// apply : every run is folded into its value in update order (:56-61).  Short runs: one thread per head, a
//         plain loop.  Long runs (the opening positions are visited by every game): one warp per run, 32 targets
//         per coalesced load, the recurrence evaluated by all lanes in step.
constexpr int kLongRun = 128;                        // runs of at least this many records take the warp path
constexpr int kLongFlag = 1 << 30;                   // tag bit: head of a long run
constexpr int kLongBlocks = 64;                      // CTAs of the apply launch that serve the long runs

struct TableCtl {
    unsigned n_long;                                 // long runs found by the probe
    unsigned next_long;                              // work cursor of the apply kernel
};

// tag[i] = -2: not a head; head of a known key: its dense index (| kLongFlag); head of a new key: -1 (short), -3 (long)
__global__ void __launch_bounds__(kSortThreads) table_probe_kernel(const u64 *__restrict__ keys, int64_t n,
                                                                   const u64 *__restrict__ slot_keys,
                                                                   const int32_t *__restrict__ slot_idx, int log2cap,
                                                                   int32_t *__restrict__ tag, unsigned *__restrict__ tile_new,
                                                                   int32_t *__restrict__ long_start, TableCtl *ctl)
{
    __shared__ unsigned s_new;
    if (threadIdx.x == 0) s_new = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
    unsigned mine = 0;
    for (int r = 0; r < kSortRows; r++) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i >= n) break;
        const u64 k = keys[i];
        int t = -2;
        if (i == 0 || keys[i - 1] != k) {
            t = table_find(k, slot_keys, slot_idx, log2cap);
            const bool is_long = i + kLongRun - 1 < n && keys[i + kLongRun - 1] == k;      // sorted: the whole stretch is k
            if (is_long) long_start[atomicAdd(&ctl->n_long, 1u)] = (int32_t)i;
            if (t < 0) { mine++; t = is_long ? -3 : -1; }
            else if (is_long) t |= kLongFlag;
        }
        tag[i] = t;
    }
    if (mine) atomicAdd(&s_new, mine);
    __syncthreads();
    if (threadIdx.x == 0) tile_new[blockIdx.x] = s_new;
}

// exclusive scan of tile_new (one CTA), total -> counters[1]
__global__ void __launch_bounds__(kSortThreads) table_scan_kernel(unsigned *__restrict__ tile_new, int64_t tiles,
                                                                  int64_t *__restrict__ counters)
{
    __shared__ unsigned s_warp[kSortWarps];
    unsigned carry = 0;
    for (int64_t base = 0; base < tiles; base += kSortThreads) {
        const int64_t i = base + threadIdx.x;
        const unsigned v = i < tiles ? tile_new[i] : 0u;
        unsigned total;
        const unsigned at = carry + block_scan_256(v, s_warp, &total);
        if (i < tiles) tile_new[i] = at;
        carry += total;
    }
    if (threadIdx.x == 0) counters[1] = (int64_t)carry;
}

// thread t looks at records [16 t, 16 t + 16) of the tile: consecutive, so a per-thread count + block scan ranks the new heads
__global__ void __launch_bounds__(kSortThreads) table_assign_kernel(const u64 *__restrict__ keys, int64_t n,
                                                                    int32_t *__restrict__ tag,
                                                                    const unsigned *__restrict__ tile_base,
                                                                    u64 *__restrict__ slot_keys, int32_t *__restrict__ slot_idx,
                                                                    int log2cap, u64 *__restrict__ dense_keys,
                                                                    double *__restrict__ dense_values, int64_t n_before)
{
    __shared__ unsigned s_warp[kSortWarps];
    const u64 mask = (1ull << log2cap) - 1ull;
    const int64_t first = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kSortRows;
    int32_t tg[kSortRows];
    unsigned mine = 0;
#pragma unroll
    for (int r = 0; r < kSortRows; r++) {
        tg[r] = first + r < n ? tag[first + r] : -2;
        mine += tg[r] == -1 || tg[r] == -3;
    }
    unsigned rank = tile_base[blockIdx.x] + block_scan_256(mine, s_warp, nullptr);
    if (!mine) return;
#pragma unroll
    for (int r = 0; r < kSortRows; r++) {
        if (tg[r] != -1 && tg[r] != -3) continue;
        const int64_t i = first + r;
        const u64 k = keys[i];
        const int64_t di = n_before + (int64_t)rank++;
        dense_keys[di] = k;
        dense_values[di] = 0.0;                                                     // `set(key, 0)` (:52-53)
        for (u64 h = slot_of(k, log2cap);; h = (h + 1) & mask)                      // keys of one batch are distinct: plain CAS claim
            if (atomicCAS((unsigned long long *)&slot_keys[h], 0ull, (unsigned long long)k) == 0ull) { slot_idx[h] = (int32_t)di; break; }
        tag[i] = (int32_t)di | (tg[r] == -3 ? kLongFlag : 0);
    }
}

__global__ void __launch_bounds__(kSortThreads) table_apply_kernel(const u64 *__restrict__ keys,
                                                                   const double *__restrict__ targets, int64_t n, double a,
                                                                   const int32_t *__restrict__ tag,
                                                                   const int32_t *__restrict__ long_start, TableCtl *ctl,
                                                                   double *__restrict__ dense_values)
{
    const double keep = __dsub_rn(1.0, a);                                          // (1 - self.a)
    if (blockIdx.x < kLongBlocks) {
        // ---- long runs: one warp per run ------------------------------------------------------------
        const int lane = threadIdx.x & 31;
        const unsigned n_long = ctl->n_long;
        for (;;) {
            unsigned which = 0;
            if (lane == 0) which = atomicAdd(&ctl->next_long, 1u);
            which = __shfl_sync(kAll, which, 0);
            if (which >= n_long) break;
            const int64_t i = long_start[which];
            const u64 k = keys[i];
            const int64_t di = tag[i] & (kLongFlag - 1);
            // where the run ends: gallop, then bisect, on the sorted keys (every lane the same walk)
            int64_t lo = i + kLongRun - 1, hi = lo + 1;                             // keys[lo] == k; the end is in (lo, hi]
            for (int64_t step = kLongRun; hi < n && keys[hi] == k; step <<= 1) { lo = hi; hi = lo + step < n ? lo + step : n; }
            while (hi - lo > 1) {
                const int64_t mid = lo + (hi - lo) / 2;
                if (keys[mid] == k) lo = mid; else hi = mid;
            }
            const int64_t end = hi;
            // The recurrence is sequential by definition; the loads are not.  A window of 32 targets is one
            // coalesced load (the next window is requested before this one is folded in), the products new * a
            // are formed by all lanes at once, and every lane then runs the same chain V*(1-a) + p over the window.
            // The `V == 0` test of the rule (:56) is almost never true after a key's first record, so it is kept
            // off the dependent chain (DMUL -> DADD per record): a zero met on the way is noticed beside the chain
            // and the window redone with the test.  Same operations, same roundings, same results.
            double v = dense_values[di];
            double x = i + lane < end ? __ldg(targets + i + lane) : 0.0;
            for (int64_t j = i; j < end; j += 32) {
                const double nx = j + 32 + lane < end ? __ldg(targets + j + 32 + lane) : 0.0;
                const double p = __dmul_rn(x, a);
                if (j + 32 <= end) {
                    double r = v;
                    bool zero = false;
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        zero |= r == 0.0;
                        r = __dadd_rn(__dmul_rn(r, keep), __shfl_sync(kAll, p, q));
                    }
                    if (!zero) v = r;
                    else {
#pragma unroll 4
                        for (int q = 0; q < 32; q++) v = smooth_step(v, __shfl_sync(kAll, x, q), keep, a);
                    }
                } else {
                    const int m = (int)(end - j);
                    for (int q = 0; q < m; q++) v = smooth_step(v, __shfl_sync(kAll, x, q), keep, a);
                }
                x = nx;
            }
            if (lane == 0) dense_values[di] = v;
        }
        return;
    }
    // ---- short runs: one thread per head ------------------------------------------------------------
    const int64_t base = (int64_t)(blockIdx.x - kLongBlocks) * kTile;
#pragma unroll 1
    for (int r = 0; r < kSortRows; r++) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i >= n) break;
        const int32_t tg = tag[i];
        if (tg < 0 || (tg & kLongFlag)) continue;
        const u64 k = keys[i];
        double v = dense_values[tg];
        int64_t j = i;
        do { v = smooth_step(v, __ldg(targets + j), keep, a); j++; } while (j < n && keys[j] == k);
        dense_values[tg] = v;
    }
}

__global__ void table_rehash_kernel(const u64 *__restrict__ dense_keys, int64_t n, u64 *__restrict__ slot_keys,
                                    int32_t *__restrict__ slot_idx, int log2cap)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = dense_keys[i], mask = (1ull << log2cap) - 1ull;
    for (u64 h = slot_of(k, log2cap);; h = (h + 1) & mask)
        if (atomicCAS((unsigned long long *)&slot_keys[h], 0ull, (unsigned long long)k) == 0ull) { slot_idx[h] = (int32_t)i; break; }
}

__global__ void table_lookup_kernel(const u64 *__restrict__ query, int64_t n, const u64 *__restrict__ slot_keys,
                                    const int32_t *__restrict__ slot_idx, int log2cap, int32_t *__restrict__ index)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) index[i] = table_find(query[i], slot_keys, slot_idx, log2cap);
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

// one pass: histogram -> row scan -> stable scatter (in -> out)
template <int MODE>
int sort_pass(const u64 *kin, const double *vin, u64 *kout, double *vout, int64_t n, int shift, unsigned world,
                     unsigned *counts, unsigned *totals, cudaStream_t s)
{
    const int64_t tiles = (n + kTile - 1) / kTile;
    sort_hist_kernel<MODE><<<(unsigned)tiles, kSortThreads, 0, s>>>(kin, n, shift, world, counts, tiles);
    sort_rowscan_kernel<<<256, kSortThreads, 0, s>>>(counts, tiles, totals);
    sort_scatter_kernel<MODE><<<(unsigned)tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, shift, world, counts, totals, tiles);
    return ob_launch_status();
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

int64_t othello_sort_workspace_bytes(int64_t n)
{
    const int64_t tiles = n > 0 ? (n + kTile - 1) / kTile : 1;
    return (256 * tiles + 256) * (int64_t)sizeof(unsigned);
}

int othello_sort_records(uint64_t *keys, double *values, uint64_t *keys_alt, double *values_alt, int64_t n,
                         int32_t key_bits, void *workspace, int64_t workspace_bytes, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && key_bits >= 1 && key_bits <= 64);
    if (n <= 1) return 0;
    OB_CHECK_ARGS(keys && values && keys_alt && values_alt && workspace && n < (1ll << 32));
    if (workspace_bytes < othello_sort_workspace_bytes(n)) return OTHELLO_E_WORKSPACE;
    const int64_t tiles = (n + kTile - 1) / kTile;
    unsigned *counts = (unsigned *)workspace, *totals = counts + 256 * tiles;
    int passes = (key_bits + 7) / 8;
    passes += passes & 1;                                   // an even number of passes ends in the caller's buffers
    u64 *k[2] = {(u64 *)keys, (u64 *)keys_alt};
    double *v[2] = {values, values_alt};
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        int rc = shift < 64 ? sort_pass<0>(k[p & 1], v[p & 1], k[(p + 1) & 1], v[(p + 1) & 1], n, shift, 0u, counts, totals,
                                           (cudaStream_t)stream)
                            : OTHELLO_E_INVALID;
        if (rc) return rc;
    }
    return 0;
}

int othello_partition_records(const uint64_t *keys, const double *values, uint64_t *keys_out, double *values_out,
                              int64_t n, int32_t world, int64_t *owner_counts, void *workspace, int64_t workspace_bytes,
                              void *stream)
{
    OB_CHECK_ARGS(n >= 0 && world >= 1 && world <= 256 && owner_counts);
    OB_CHECK_ARGS(workspace && workspace_bytes >= othello_sort_workspace_bytes(n));
    const int64_t tiles = n > 0 ? (n + kTile - 1) / kTile : 1;
    unsigned *counts = (unsigned *)workspace, *totals = counts + 256 * tiles;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        OB_CUDA(cudaMemsetAsync(owner_counts, 0, sizeof(int64_t) * world, s));
        return 0;
    }
    OB_CHECK_ARGS(keys && values && keys_out && values_out && n < (1ll << 32));
    int rc = sort_pass<1>((const u64 *)keys, values, (u64 *)keys_out, values_out, n, 0, (unsigned)world, counts, totals, s);
    if (rc) return rc;
    export_totals_kernel<<<1, 256, 0, s>>>(totals, owner_counts, world);
    return ob_launch_status();
}

// workspace of a probe / apply pair: tag[n] | tile_new[tiles] | long_start[n / kLongRun + 1] | TableCtl
static size_t ws_align(size_t v) { return (v + 255) & ~(size_t)255; }
struct TableWs {
    int32_t *tag; unsigned *tile_new; int32_t *long_start; TableCtl *ctl; size_t bytes;
};
static TableWs table_ws(void *workspace, int64_t n)
{
    const int64_t tiles = n > 0 ? (n + kTile - 1) / kTile : 1;
    char *p = (char *)workspace;
    TableWs w;
    size_t off = 0;
    w.tag = (int32_t *)(p + off); off += ws_align((size_t)n * sizeof(int32_t));
    w.tile_new = (unsigned *)(p + off); off += ws_align((size_t)tiles * sizeof(unsigned));
    w.long_start = (int32_t *)(p + off); off += ws_align((size_t)(n / kLongRun + 1) * sizeof(int32_t));
    w.ctl = (TableCtl *)(p + off); off += 256;
    w.bytes = off;
    return w;
}

int64_t othello_table_workspace_bytes(int64_t n) { return (int64_t)table_ws(nullptr, n < 0 ? 0 : n).bytes; }

int othello_table_probe(const uint64_t *sorted_keys, int64_t n, const uint64_t *slot_keys, const int32_t *slot_idx,
                        int32_t log2_capacity, void *workspace, int64_t workspace_bytes, int64_t *counters, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && n < (1ll << 31) && counters && log2_capacity >= 4 && log2_capacity <= 30);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { OB_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(int64_t), s)); return 0; }
    OB_CHECK_ARGS(sorted_keys && slot_keys && slot_idx && workspace && workspace_bytes >= othello_table_workspace_bytes(n));
    const int64_t tiles = (n + kTile - 1) / kTile;
    const TableWs w = table_ws(workspace, n);
    OB_CUDA(cudaMemsetAsync(w.ctl, 0, sizeof(TableCtl), s));
    table_probe_kernel<<<(unsigned)tiles, kSortThreads, 0, s>>>((const u64 *)sorted_keys, n, (const u64 *)slot_keys, slot_idx,
                                                               log2_capacity, w.tag, w.tile_new, w.long_start, w.ctl);
    table_scan_kernel<<<1, kSortThreads, 0, s>>>(w.tile_new, tiles, counters);
    return ob_launch_status();
}

int othello_table_apply(const uint64_t *sorted_keys, const double *sorted_targets, int64_t n, double a,
                        uint64_t *slot_keys, int32_t *slot_idx, int32_t log2_capacity, uint64_t *dense_keys,
                        double *dense_values, int64_t n_before, void *workspace, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && n < (1ll << 31) && n_before >= 0 && log2_capacity >= 4 && log2_capacity <= 30);
    if (n == 0) return 0;
    OB_CHECK_ARGS(sorted_keys && sorted_targets && slot_keys && slot_idx && dense_keys && dense_values && workspace);
    const int64_t tiles = (n + kTile - 1) / kTile;
    const TableWs w = table_ws(workspace, n);
    cudaStream_t s = (cudaStream_t)stream;
    table_assign_kernel<<<(unsigned)tiles, kSortThreads, 0, s>>>((const u64 *)sorted_keys, n, w.tag, w.tile_new,
                                                                (u64 *)slot_keys, slot_idx, log2_capacity, (u64 *)dense_keys,
                                                                dense_values, n_before);
    table_apply_kernel<<<(unsigned)(tiles + kLongBlocks), kSortThreads, 0, s>>>((const u64 *)sorted_keys, sorted_targets, n, a,
                                                                                w.tag, w.long_start, w.ctl, dense_values);
    return ob_launch_status();
}

int othello_table_rehash(const uint64_t *dense_keys, int64_t n, uint64_t *slot_keys, int32_t *slot_idx,
                         int32_t log2_capacity, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && log2_capacity >= 4 && log2_capacity <= 30 && n <= (1ll << log2_capacity) / 2);
    if (n == 0) return 0;
    OB_CHECK_ARGS(dense_keys && slot_keys && slot_idx);
    table_rehash_kernel<<<ob_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const u64 *)dense_keys, n, (u64 *)slot_keys,
                                                                              slot_idx, log2_capacity);
    return ob_launch_status();
}

int othello_table_lookup(const uint64_t *query, int64_t n, const uint64_t *slot_keys, const int32_t *slot_idx,
                         int32_t log2_capacity, int32_t *index, void *stream)
{
    OB_CHECK_ARGS(n >= 0 && log2_capacity >= 4 && log2_capacity <= 31);
    if (n == 0) return 0;
    OB_CHECK_ARGS(query && slot_keys && slot_idx && index);
    table_lookup_kernel<<<ob_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const u64 *)query, n, (const u64 *)slot_keys,
                                                                             slot_idx, log2_capacity, index);
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
