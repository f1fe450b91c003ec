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
constexpr int SORT_THREADS = 256;
constexpr int SORT_SMEM_KEYS = 4096;  // 32 KB of shared memory; longer lists sort in global memory

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
        for (int i = threadIdx.x; i < half; i += SORT_THREADS) {
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
            for (int i = threadIdx.x; i < half; i += SORT_THREADS) {
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

// Lists of up to WARP_SORT_MAX = 256 entries (nearly all of them: the mean list holds ~160) are sorted by ONE
// warp in registers: element e = r * 32 + lane lives in register r of its lane, compare-exchange
// partners at distance < 32 are reached with a shuffle, larger distances are other registers of the
// same lane.  No shared memory, no block barrier; eight tiles per CTA.  Padding keys are all-ones
// (greater than any depth<<32|index key), so they stay behind the n real entries.
constexpr int WARP_SORT_MAX = 256;   // (512 with 16 registers per lane was measured: the long lists then form a single-warp tail, 0.063 -> 0.078 ms)

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

__global__ void __launch_bounds__(SORT_THREADS) tile_sort_warp_kernel(TileSortArgs a) {
    if ((int64_t)a.counters[CNT_RENDERED] > a.capacity) return;
    const int slot = blockIdx.x * (SORT_THREADS / 32) + (threadIdx.x >> 5);
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

// Lists longer than WARP_SORT_MAX: one CTA per tile, in shared memory.
__global__ void __launch_bounds__(SORT_THREADS) tile_sort_kernel(TileSortArgs a) {
    if ((int64_t)a.counters[CNT_RENDERED] > a.capacity) return;
    extern __shared__ unsigned long long s_keys[];
    const int tile = (int)a.tile_order[blockIdx.x];
    const uint32_t off = a.tile_offset[tile];
    const int n = (int)(a.tile_offset[tile + 1] - off);
    if (n <= WARP_SORT_MAX) return;
    unsigned long long* gk = a.keys + off;
    if (n <= SORT_SMEM_KEYS) {
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) s_keys[i] = gk[i];
        __syncthreads();
        bitonic_sort_ascending(s_keys, n);
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) a.list[off + i] = (uint32_t)s_keys[i];
    } else {
        // rare: very long list -- same network, in place in global memory (L2-resident)
        __syncthreads();
        bitonic_sort_ascending(gk, n);
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) a.list[off + i] = (uint32_t)gk[i];
    }
}

void launch_tile_scan(const TileScanArgs& a, cudaStream_t s) {
    tile_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(a);
    count_launch();
}
void launch_tile_sort(const TileSortArgs& a, cudaStream_t s) {
    if (a.num_tiles <= 0) return;
    // long lists first (they sit at the front of tile_order), then everything else by warps
    tile_sort_kernel<<<a.num_tiles, SORT_THREADS, SORT_SMEM_KEYS * sizeof(unsigned long long), s>>>(a);
    count_launch();
    const int warps_per_cta = SORT_THREADS / 32;
    tile_sort_warp_kernel<<<(a.num_tiles + warps_per_cta - 1) / warps_per_cta, SORT_THREADS, 0, s>>>(a);
    count_launch();
}

}  // namespace g4s
