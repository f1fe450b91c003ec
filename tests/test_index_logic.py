"""Index logic of two kernels restated in Python and checked exhaustively / on random inputs (no GPU):
the register bitonic network and the chunk merge of `tile_sort_kernel` and the owner search of `scatter_kernel`
(g4splat_b200/csrc/binning.cu, project.cu).  The GPU parity tests check the kernels themselves; these keep
the reasoning behind their index arithmetic executable."""
import random


def warp_bitonic(keys, nreg):
    """binning.cu: warp_bitonic_sort<NREG> -- element e = r * 32 + lane lives in register r of its lane."""
    n_total = nreg * 32
    pad = (1 << 64) - 1
    key = [[keys[r * 32 + l] if r * 32 + l < len(keys) else pad for l in range(32)] for r in range(nreg)]
    k = 2
    while k <= n_total:
        j = k >> 1
        while j > 0:
            if j >= 32:                                   # partner = another register of the same lane
                jr = j >> 5
                for r in range(nreg):
                    if (r & jr) == 0:
                        asc = ((r * 32) & k) == 0
                        for lane in range(32):
                            x, y = key[r][lane], key[r | jr][lane]
                            if (x > y) if asc else (x < y):
                                key[r][lane], key[r | jr][lane] = y, x
            else:                                         # partner = lane ^ j, reached with a shuffle
                for r in range(nreg):
                    new = [None] * 32
                    for lane in range(32):
                        lower = (lane & j) == 0
                        asc = (((r * 32) & k) == 0) if k >= 32 else ((lane & k) == 0)
                        x, y = key[r][lane], key[r][lane ^ j]
                        new[lane] = min(x, y) if lower == asc else max(x, y)
                    key[r] = new
            j >>= 1
        k <<= 1
    return [key[e // 32][e % 32] for e in range(len(keys))]


def test_warp_bitonic_network_sorts_every_length():
    rng = random.Random(1)
    for nreg in (1, 2, 4, 8):
        for n in sorted({1, 2, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256} & set(range(1, nreg * 32 + 1))):
            keys = [rng.getrandbits(60) for _ in range(n)]
            assert warp_bitonic(keys, nreg) == sorted(keys), (nreg, n)
    # equal depths: (depth << 32 | index) keys are distinct, ties in depth resolve by index
    keys = [(7 << 32) | i for i in rng.sample(range(1000), 200)]
    assert warp_bitonic(keys, 8) == sorted(keys)


def hybrid_sort(keys):
    """binning.cu: hybrid_sort_tile -- 256-key chunks sorted in registers (one warp each), merged by the all-ascending
    network: the mirror step and the half-cleaners at chunk distances go through shared memory, the distances below
    256 are warp_merge_tail.  Chunks wholly above n are virtual +inf and are never exchanged with."""
    pad = (1 << 64) - 1
    n = len(keys)
    C = (n + 255) >> 8
    ch = [[keys[w * 256 + e] if w * 256 + e < n else pad for e in range(256)] for w in range(C)]
    ch = [warp_bitonic([k for k in c], 8) for c in ch]

    def merge_tail(c):
        c = list(c)
        j = 128
        while j > 0:
            for lo in range(256):
                if (lo & j) == 0 and c[lo] > c[lo + j]:
                    c[lo], c[lo + j] = c[lo + j], c[lo]
            j >>= 1
        return c

    kc = 2
    while kc < 2 * C:
        step = 0
        while True:
            mirror = step == 0
            dist = 0 if mirror else (kc >> (step + 1))
            if not mirror and dist == 0:
                break
            buf = [list(c) for c in ch]                   # the shared-memory copy every warp reads its partner from
            for w in range(C):
                cb = w & (kc - 1)
                partner = (w - cb + (kc - 1 - cb)) if mirror else (w ^ dist)
                lower = (cb < (kc >> 1)) if mirror else ((w & dist) == 0)
                if partner < C:
                    for e in range(256):
                        y = buf[partner][255 - e if mirror else e]
                        ch[w][e] = min(ch[w][e], y) if lower else max(ch[w][e], y)
            step += 1
        ch = [merge_tail(c) for c in ch]
        kc <<= 1
    return [ch[e >> 8][e & 255] for e in range(n)]


def test_hybrid_chunk_merge_sorts_every_chunk_count():
    rng = random.Random(2)
    for n in (257, 300, 511, 512, 513, 700, 768, 769, 794, 1000, 1023, 1024):
        keys = [rng.getrandbits(60) for _ in range(n)]
        assert hybrid_sort(keys) == sorted(keys), n
        keys = [(rng.randrange(5) << 32) | i for i in rng.sample(range(5000), n)]     # heavy depth ties
        assert hybrid_sort(keys) == sorted(keys), n
    # the network itself is valid up to eight chunks (the kernel uses four)
    for n in (1025, 1537, 2048):
        keys = [rng.getrandbits(60) for _ in range(n)]
        assert hybrid_sort(keys) == sorted(keys), n


def scatter_owner(excl, i):
    """project.cu: scatter_kernel -- the last lane whose exclusive offset is <= i, by binary search."""
    lo = 0
    for step in (16, 8, 4, 2, 1):
        probe = lo + step
        if probe < 32 and excl[probe & 31] <= i:
            lo = probe
    return lo


def test_scatter_owner_search_lands_on_the_live_lane():
    rng = random.Random(3)
    for _ in range(500):
        counts = [rng.choice([0, 0, 0, 1, 2, 5, 9, 40, 700]) for _ in range(32)]
        excl = [sum(counts[:l]) for l in range(32)]
        for i in range(sum(counts)):
            lane = scatter_owner(excl, i)
            assert counts[lane] > 0 and excl[lane] <= i < excl[lane] + counts[lane]


def bucket_of(c):
    """binning.cu: tile_scan_kernel -- two buckets per octave of the list length, longest first; empty tiles last."""
    if c == 0:
        return 63
    lg = c.bit_length() - 1
    half = ((c >> (lg - 1)) & 1) if lg > 0 else 0
    return max(0, 64 - 2 - (2 * lg + half))


def test_long_list_prefix_of_the_tile_order():
    """tile_sort_kernel's long-list CTAs walk tile_order[0 : counters[CNT_LONG_TILES]], the base of bucket LONG_BUCKET_END
    = 47: that prefix must hold every tile with more than 256 entries (and may hold tiles with exactly 256, which the
    long-list role skips and a warp sorts)."""
    assert all(bucket_of(c) < 47 for c in list(range(256, 5000)) + [10 ** 5, 10 ** 6, 2 ** 31 - 1])
    assert all(bucket_of(c) >= 47 for c in range(0, 256))
    # buckets descend with the length, so a counting sort by bucket is longest-first to within one half-octave
    lengths = [0, 1, 2, 3, 5, 31, 32, 100, 255, 256, 257, 383, 384, 794, 1024, 4097, 70000]
    buckets = [bucket_of(c) for c in lengths]
    assert buckets == sorted(buckets, reverse=True)
