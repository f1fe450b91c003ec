// binning.cu -- per-tile offsets and the per-tile depth sort.
//
// Reference semantics (CR/rasterizer_impl.cu:276-319): every tile walks its Gaussians ordered by
// (view depth as uint32 bits, Gaussian index) -- the order a stable LSD radix sort of
// tile<<32|depth keys produces after emitting in index order.  Here every tile owns a bucket
// [tile_offset[t], tile_offset[t+1]) of unordered 64-bit keys depth<<32|index; one CTA per tile
// sorts its bucket in shared memory (bitonic network) and writes the Gaussian ids.  Total
// instance traffic is 8 B written + 8 B read + 4 B written, against six 24-byte radix passes.
#include "kernels.cuh"

namespace g4s {

constexpr int SCAN_THREADS = 1024;
constexpr int ORDER_BUCKETS = 64;
constexpr int LONG_BUCKET_END = 47;   // bucket_of(c) < 47  <=>  c >= 256 (lg = 8, lower half -> 64 - 2 - 16 = 46)

// Single-CTA exclusive scan over the tile counts + bucketed longest-first tile order.
__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(TileScanArgs a) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t bucket_count[ORDER_BUCKETS];
    __shared__ uint32_t bucket_base[ORDER_BUCKETS];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.num_tiles;
    const int per = (T + SCAN_THREADS - 1) / SCAN_THREADS;
    const int begin = min(T, tid * per), end = min(T, begin + per);
    if (tid < ORDER_BUCKETS) bucket_count[tid] = 0;
    if (tid == 0) s_max = 0;
    __syncthreads();

    // bucket: 2 buckets per octave of the list length, longest first
    auto bucket_of = [](uint32_t c) -> int {
        if (c == 0) return ORDER_BUCKETS - 1;
        const int lg = 31 - __clz(c);
        const int half = (lg > 0) ? ((c >> (lg - 1)) & 1) : 0;
        const int b = 2 * lg + half;             // 0 .. 63 ascending with c
        return max(0, ORDER_BUCKETS - 2 - b);    // descending
    };
    uint32_t local = 0, lmax = 0;
    for (int t = begin; t < end; t++) {
        const uint32_t c = a.tile_count[t];
        local += c;
        lmax = max(lmax, c);
        atomicAdd(&bucket_count[bucket_of(c)], 1u);
    }
    // block exclusive scan of `local`
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    atomicMax(&s_max, lmax);
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        warp_sums[lane] = wi - w;  // exclusive
        // exclusive scan of the bucket counts (64 buckets, 2 per lane)
        const uint32_t b0 = bucket_count[2 * lane], b1 = bucket_count[2 * lane + 1];
        uint32_t bi = b0 + b1;
        const uint32_t bsum = bi;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, bi, o);
            if (lane >= o) bi += v;
        }
        bucket_base[2 * lane] = bi - bsum;
        bucket_base[2 * lane + 1] = bi - bsum + b0;
        // tiles with >= 256 entries fill the buckets below LONG_BUCKET_END: their number is that bucket's base, i.e. the
        // prefix of tile_order the long-list sort has to look at
        if (2 * lane + 1 == LONG_BUCKET_END) a.counters[CNT_LONG_TILES] = (int32_t)(bi - bsum + b0);
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + (incl - local);
    for (int t = begin; t < end; t++) {
        const uint32_t c = a.tile_count[t];
        a.tile_offset[t] = run;
        run += c;
        a.tile_count[t] = 0;  // becomes the scatter cursor
        const uint32_t slot = atomicAdd(&bucket_base[bucket_of(c)], 1u);
        a.tile_order[slot] = (uint32_t)t;
    }
    if (tid == SCAN_THREADS - 1) {
        a.tile_offset[T] = run;
        // The instance count is scanned in 32 bits.  A total above INT32_MAX cannot be represented in the counter the
        // guards of every later kernel compare with `capacity`: saturate it (the guards then always refuse, the host
        // sees a count no capacity can satisfy and fails the call) instead of wrapping negative and passing them.
        a.counters[CNT_RENDERED] = run > 0x7fffffffu ? 0x7fffffff : (int32_t)run;
    }
    if (tid == 0) a.counters[CNT_MAXLEN] = (int32_t)s_max;
}

// ---------------------------------------------------------------------------------------------
// The per-tile sort is ONE launch of 128-thread CTAs with two roles (tile_sort_kernel below):
//   * lists of up to 256 entries (nearly all of them: the mean list holds ~160) -- one warp per list, in registers;
//   * longer lists -- one CTA per list: 256-key register chunks merged through shared memory up to 1024 entries, the
//     shared-memory network below up to 2048, the same network in global memory (L2-resident) above.
// tile_scan_kernel orders the tiles longest first and counts the long ones, so the long-list CTAs (first in the grid,
// persistent over that prefix of tile_order) start before the short-list CTAs and nothing is launched per empty slot.
// History (ms per view at c2): one 256-thread CTA per tile for the long lists + a second launch for the warps 0.064
// (8160 CTAs, seven in eight of which looked at their tile and left); this kernel 0.048.

// Bitonic sorting network in its "all ascending" form: the first step of every merge compares
// element i of the lower half with its mirror in the upper half, the remaining steps are the
// usual half-cleaners.  Because every compare-exchange orders (lo < hi) ascending, virtual +inf
// padding above n never moves, so lists of any length sort in place by skipping hi >= n.
template <typename KeyArray>
__device__ __forceinline__ void bitonic_sort_ascending(KeyArray keys, int n) {
    int m = 1, log_m = 0;
    while (m < n) { m <<= 1; log_m++; }
    const int half = m >> 1;
    // all block sizes are powers of two: index arithmetic is shifts and masks, no division
    for (int lk = 1; lk <= log_m; lk++) {          // k = 1 << lk
        const int k = 1 << lk, hk = k >> 1;
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const int within = i & (hk - 1);
            const int blk_base = (i >> (lk - 1)) << lk;
            const int lo = blk_base + within, hi = blk_base + k - 1 - within;
            if (hi < n) {
                const unsigned long long x = keys[lo], y = keys[hi];
                if (x > y) { keys[lo] = y; keys[hi] = x; }
            }
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; lj--) {     // j = 1 << lj
            const int j = 1 << lj;
            for (int i = threadIdx.x; i < half; i += blockDim.x) {
                const int lo = ((i >> lj) << (lj + 1)) | (i & (j - 1)), hi = lo + j;
                if (hi < n) {
                    const unsigned long long x = keys[lo], y = keys[hi];
                    if (x > y) { keys[lo] = y; keys[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// Lists of up to WARP_SORT_MAX = 256 entries are sorted by ONE warp in registers: element e = r * 32 + lane lives in
// register r of its lane, compare-exchange partners at distance < 32 are reached with a shuffle, larger distances are
// other registers of the same lane.  No shared memory, no block barrier.  Padding keys are all-ones (greater than
// any depth<<32|index key), so they stay behind the n real entries.
constexpr int WARP_SORT_MAX = 256;   // (512 with 16 registers per lane was measured: the long lists then form a single-warp tail, 0.063 -> 0.078 ms.
                                     //  A "blocked" layout e = lane * NREG + r halves the shuffle steps and gained 1.7 us of 48: the kernel is bound
                                     //  by its ~20 M warp instructions and the dependent loads, not by the shuffle rate; not kept.)

template <int NREG>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&key)[NREG], int lane) {
    constexpr int N = NREG * 32;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < NREG; r++) {
                    if ((r & jr) == 0) {
                        const bool asc = (((r * 32) & k) == 0);     // bit of k lies in the register index (k >= 64 here)
                        const unsigned long long x = key[r], y = key[r | jr];
                        const bool swap = asc ? (x > y) : (x < y);
                        key[r] = swap ? y : x;
                        key[r | jr] = swap ? x : y;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < NREG; r++) {
                    const bool asc = (k >= 32) ? (((r * 32) & k) == 0) : ((lane & k) == 0);   // bit of k: register index or lane
                    const unsigned long long x = key[r];
                    const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
                    const bool want_min = (lower == asc);
                    key[r] = want_min ? (x < y ? x : y) : (x > y ? x : y);
                }
            }
        }
    }
}

template <int NREG>
__device__ __forceinline__ void warp_sort_tile(const unsigned long long* __restrict__ gk, uint32_t* __restrict__ out,
                                               int n, int lane) {
    unsigned long long key[NREG];
#pragma unroll
    for (int r = 0; r < NREG; r++) {
        const int e = r * 32 + lane;
        key[r] = e < n ? gk[e] : ~0ull;
    }
    warp_bitonic_sort<NREG>(key, lane);
#pragma unroll
    for (int r = 0; r < NREG; r++) {
        const int e = r * 32 + lane;
        if (e < n) out[e] = (uint32_t)key[r];
    }
}

// ---- one launch for both list classes ----------------------------------------------------------------------
// Ascending merge tail of the all-ascending network: the half-cleaners at distances 128 .. 1 inside a 256-key chunk
// held as 8 registers per lane (element e = r * 32 + lane).
__device__ __forceinline__ void warp_merge_tail(unsigned long long (&key)[8], int lane) {
#pragma unroll
    for (int jr = 4; jr > 0; jr >>= 1) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if ((r & jr) == 0) {
                const unsigned long long x = key[r], y = key[r | jr];
                const bool swap = x > y;
                key[r] = swap ? y : x;
                key[r | jr] = swap ? x : y;
            }
        }
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const unsigned long long x = key[r];
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
            key[r] = lower ? (x < y ? x : y) : (x > y ? x : y);
        }
    }
}

// Lists of 257 .. 1024 entries: warp w sorts chunk w (256 keys) in registers, then the chunks are merged by the
// all-ascending network -- only the steps whose partner lies in ANOTHER chunk (the mirror step of a merge and its
// half-cleaners at distances >= 256) go through shared memory, one block barrier each (two buffers, used alternately);
// everything below distance 256 is the register tail above.  A 794-entry list takes 3 barriers instead of the 55 of
// the shared-memory network.  Chunks that lie wholly above n are virtual +inf keys: an exchange with one never changes
// the real chunk (it is the lower side), so it is skipped.
constexpr int HYBRID_MAX = 1024;          // four warps x 256 keys
__device__ __forceinline__ void hybrid_sort_tile(const unsigned long long* __restrict__ gk, uint32_t* __restrict__ out, int n,
                                                 unsigned long long* s_buf /* [2][HYBRID_MAX] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int C = (n + 255) >> 8;
    unsigned long long key[8];
    if (w < C) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int e = w * 256 + r * 32 + lane;
            key[r] = e < n ? gk[e] : ~0ull;
        }
        warp_bitonic_sort<8>(key, lane);
    }
    int phase = 0;
    for (int kc = 2; kc < 2 * C; kc <<= 1) {                 // chunks per merged block
        // step 0: the mirror step; steps 1..: half-cleaners at chunk distances kc / 4, kc / 8, .. 1
        for (int step = 0;; step++) {
            const bool mirror = step == 0;
            const int dist = mirror ? 0 : (kc >> (step + 1));
            if (!mirror && dist == 0) break;
            unsigned long long* buf = s_buf + phase * HYBRID_MAX;
            phase ^= 1;
            if (w < C) {
#pragma unroll
                for (int r = 0; r < 8; r++) buf[w * 256 + r * 32 + lane] = key[r];
            }
            __syncthreads();
            if (w < C) {
                const int cb = w & (kc - 1);
                const int partner = mirror ? (w - cb + (kc - 1 - cb)) : (w ^ dist);
                const bool lower = mirror ? (cb < (kc >> 1)) : ((w & dist) == 0);
                if (partner < C) {
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        const int e = r * 32 + lane;
                        const unsigned long long y = buf[partner * 256 + (mirror ? 255 - e : e)];
                        const unsigned long long x = key[r];
                        key[r] = lower ? (x < y ? x : y) : (x > y ? x : y);
                    }
                }
            }
        }
        if (w < C) warp_merge_tail(key, lane);
    }
    if (w < C) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int e = w * 256 + r * 32 + lane;
            if (e < n) out[e] = (uint32_t)key[r];
        }
    }
}

constexpr int PS_THREADS = 128;            // four warps: one long list (<= 1024 entries by register chunks) or four short ones
constexpr int PS_SMEM_KEYS = 2 * HYBRID_MAX;
constexpr int LONG_CTAS = 2072;            // 14 x 148: more than the SMs hold at once (10 per SM at 48 registers); 1024 .. 4096 measured within 3 us
__global__ void __launch_bounds__(PS_THREADS) tile_sort_kernel(TileSortArgs a) {
    if ((int64_t)a.counters[CNT_RENDERED] > a.capacity) return;
    __shared__ unsigned long long s_keys[PS_SMEM_KEYS];
    // long lists first: tile_order starts with them, longest first, and with ~1500 CTAs resident each normally gets
    // its own CTA; the short-list CTAs fill the SMs as those retire
    const int b = (int)blockIdx.x;
    if (b < LONG_CTAS) {
        const int n_long = a.counters[CNT_LONG_TILES];
        for (int i = b; i < n_long; i += LONG_CTAS) {
            const int tile = (int)a.tile_order[i];
            const uint32_t off = a.tile_offset[tile];
            const int n = (int)(a.tile_offset[tile + 1] - off);
            if (n <= WARP_SORT_MAX) continue;              // exactly 256: a warp below takes it
            unsigned long long* gk = a.keys + off;
            if (n <= HYBRID_MAX) {
                hybrid_sort_tile(gk, a.list + off, n, s_keys);
            } else if (n <= PS_SMEM_KEYS) {
                for (int j = threadIdx.x; j < n; j += PS_THREADS) s_keys[j] = gk[j];
                __syncthreads();
                bitonic_sort_ascending(s_keys, n);
                for (int j = threadIdx.x; j < n; j += PS_THREADS) a.list[off + j] = (uint32_t)s_keys[j];
            } else {
                __syncthreads();
                bitonic_sort_ascending(gk, n);
                for (int j = threadIdx.x; j < n; j += PS_THREADS) a.list[off + j] = (uint32_t)gk[j];
            }
            __syncthreads();                               // the shared buffers are reused by the next tile
        }
        return;
    }
    const int slot = (b - LONG_CTAS) * (PS_THREADS / 32) + (threadIdx.x >> 5);
    if (slot >= a.num_tiles) return;
    const int lane = threadIdx.x & 31;
    const int tile = (int)a.tile_order[slot];
    const uint32_t off = a.tile_offset[tile];
    const int n = (int)(a.tile_offset[tile + 1] - off);
    if (n == 0 || n > WARP_SORT_MAX) return;
    const unsigned long long* gk = a.keys + off;
    uint32_t* out = a.list + off;
    if (n <= 32) warp_sort_tile<1>(gk, out, n, lane);
    else if (n <= 64) warp_sort_tile<2>(gk, out, n, lane);
    else if (n <= 128) warp_sort_tile<4>(gk, out, n, lane);
    else warp_sort_tile<8>(gk, out, n, lane);
}

void launch_tile_scan(const TileScanArgs& a, cudaStream_t s) {
    tile_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(a);
    count_launch();
}
void launch_tile_sort(const TileSortArgs& a, cudaStream_t s) {
    if (a.num_tiles <= 0) return;
    const int short_ctas = (a.num_tiles + PS_THREADS / 32 - 1) / (PS_THREADS / 32);
    tile_sort_kernel<<<LONG_CTAS + short_ctas, PS_THREADS, 0, s>>>(a);
    count_launch();
}

}  // namespace g4s
