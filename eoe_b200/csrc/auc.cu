// ROC-AUC (+ PRC / average precision) on the device, bit-exact with scikit-learn as the reference calls it
// (src/eoe/training/ad_trainer.py:453-454, 517-521): roc_curve(labels, scores) -> auc(fpr, tpr).
//
// Pipeline (all on the caller's stream, no host sync; every size after the sort is a DEVICE value):
//   1 keys      score -> descending-sortable u32 key (-0.0 == +0.0), label bit, 4x256 digit histogram, counts
//   2 sort      4 LSD radix passes (8 bit), one kernel each: per-warp match_any ranking + decoupled look-back, the tile
//               reordered by digit in shared memory so that the scatter writes whole runs
//   3 distinct  tie merge: flag last element of every run of equal keys, inclusive label count -> (tps, fps)
//               at distinct thresholds, compacted with a single-pass (look-back) scan         [_binary_clf_curve]
//               (warp-striped layout: ballot + popc scans, coalesced loads and stores; 4 likewise)
//   4 corners   drop_intermediate: keep points whose 2nd difference of fps or tps is non-zero [roc_curve]
//   5 terms     fpr = fps/fps[-1], tpr = tps/tps[-1]; term_i = (fpr[i+1]-fpr[i]) * (tpr[i+1]+tpr[i]) / 2.0
//   6 leaves    numpy pairwise-sum leaves (<=128 terms: 8 strided accumulators)              [np.trapezoid -> sum]
//   7 tree      numpy pairwise-sum internal nodes (split at n/2 rounded down to a multiple of 8), top levels in shared
//               memory; the last launch also writes the counts / status words (info_out)
// The float64 summation ORDER is what makes the result bit-identical to sklearn; see oracle/auc.py.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace eoe {

constexpr int kSortThreads = 256;
// Keys per thread of a sort tile: 16 (4096-key tiles) or, up to kSortSmallTileMax scores, 8 (2048-key tiles: below ~150
// large tiles part of the SMs would idle; measured 8 - 17 % faster up to 512 k scores, a tie at 640 - 768 k, 7 % slower at
// 1 M, within +-5 % beyond -- profiles/r2_auc_tile_size_sweep.jsonl)
#ifndef EOE_AUC_SORT_ITEMS
#define EOE_AUC_SORT_ITEMS 0                           // 8 / 16: force one tile size (A/B builds)
#endif
constexpr int64_t kSortSmallTileMax = 640 * 1024;
static inline int sort_items_for(int64_t n) {
    if (EOE_AUC_SORT_ITEMS) return EOE_AUC_SORT_ITEMS;
    return n <= kSortSmallTileMax ? 8 : 16;
}
#ifndef EOE_AUC_BALLOT_RANK
#define EOE_AUC_BALLOT_RANK 0                          // 1: rank with eight ballots per key instead of match.any (A/B builds)
#endif
#ifndef EOE_AUC_KEYS_V2
#define EOE_AUC_KEYS_V2 1                              // 0: the round-1 keys kernel (A/B builds)
#endif
constexpr int kKeysThreads = EOE_AUC_KEYS_V2 ? 1024 : 256;
constexpr int kKeysBatch = 4;                          // rows per thread and trip (loads issued together)
constexpr int kScanThreads = 256;
#ifndef EOE_AUC_SCAN_ITEMS
#define EOE_AUC_SCAN_ITEMS 8
#endif
constexpr int kScanItems = EOE_AUC_SCAN_ITEMS;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048
constexpr uint32_t kFlagAgg = 1u, kFlagIncl = 2u;

// Programmatic dependent launch between the kernels of the tiled pipeline: every kernel lets its successor be scheduled at
// once (its blocks become resident as slots free up) and itself waits, before touching anything, until its predecessor
// has completed and its writes are visible -- the ~2 us of launch latency and block scheduling per boundary overlap the
// predecessor's tail.  Completion is transitive (a kernel cannot finish before its predecessor did), so data produced
// several kernels back is covered as well.  Both instructions are no-ops in a launch without the attribute.
#ifndef EOE_AUC_PDL
#define EOE_AUC_PDL 1
#endif
__device__ __forceinline__ void pdl_enter() {
#if EOE_AUC_PDL
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_dependent(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = EOE_AUC_PDL ? 1 : 0;
    const cudaError_t r = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    if (r != cudaSuccess) set_cuda_error(r, "auc pipeline (dependent launch)");
    return r;
}
constexpr int kSpinLimit = 1 << 20;                     // bounded spins: a protocol bug must not hang the GPU

struct AucControl {            // zeroed by one memset per call
    uint32_t hist[4 * 256];
    uint32_t tickets[8];       // 0-3 sort passes, 4 distinct scan, 5 corner scan, 6 prc
    uint32_t status;           // EOE_AUC_STATUS_* | 0x100 internal protocol error
    uint32_t pad0[7];
    unsigned long long n_valid, n_pos, n_distinct, n_kept;   // n_kept excludes the prepended origin
    unsigned long long pad1[4];
};

struct AucLayout {
    size_t control, sort_status, scan1_status, scan2_status, control_bytes;
    size_t keys_a, keys_b, labs_a, labs_b, d_tps, d_fps, k_tps, k_fps, terms, nodes, total;
    int sort_tiles, scan_tiles, max_depth;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Upper bound, monotone in n, of the depth of numpy's pairwise tree over any T <= n terms.  The exact depth
// (right spine: len -> len - ((len/2) & ~7) <= (len+15)/2) is NOT monotone in T (65443 terms are one level
// deeper than 65536), so buffers and grids are sized with the bound; kernels use the exact depth of T.
static int pairwise_depth(int64_t n) {
    int d = 0;
    while (n > 128) { n = (n + 15) / 2; ++d; }
    return d;
}

static AucLayout auc_layout(int64_t n) {
    AucLayout L;
    const int64_t sort_tile = (int64_t)kSortThreads * sort_items_for(n);
    L.sort_tiles = (int)((n + sort_tile - 1) / sort_tile);
    L.scan_tiles = (int)((n + kScanTile - 1) / kScanTile);
    L.max_depth = pairwise_depth(n);
    size_t o = 0;
    L.control = o; o = align_up(o + sizeof(AucControl), 256);
    // (reserved for the larger of the two tile counts a smaller n could need: eoe_auc_workspace_bytes stays monotone in n,
    //  which grow-only callers rely on)
    const int64_t small_n = n < kSortSmallTileMax ? n : kSortSmallTileMax;
    const int64_t small_tiles = EOE_AUC_SORT_ITEMS ? 0 : (small_n + kSortThreads * 8 - 1) / (kSortThreads * 8);
    const int64_t status_tiles = L.sort_tiles > small_tiles ? L.sort_tiles : small_tiles;
    L.sort_status = o; o = align_up(o + (size_t)4 * status_tiles * 256 * 4, 256);
    L.scan1_status = o; o = align_up(o + (size_t)L.scan_tiles * 8, 256);
    L.scan2_status = o; o = align_up(o + (size_t)L.scan_tiles * 8, 256);
    L.control_bytes = o;
    L.keys_a = o; o = align_up(o + (size_t)n * 4, 256);
    L.keys_b = o; o = align_up(o + (size_t)n * 4 + 64, 256);
    L.labs_a = o; o = align_up(o + (size_t)n, 256);
    L.labs_b = o; o = align_up(o + (size_t)n, 256);
    L.d_tps = o; o = align_up(o + (size_t)(n + 2) * 4, 256);
    L.d_fps = o; o = align_up(o + (size_t)(n + 2) * 4, 256);
    L.k_tps = o; o = align_up(o + (size_t)(n + 2) * 4, 256);
    L.k_fps = o; o = align_up(o + (size_t)(n + 2) * 4, 256);
    L.terms = o; o = align_up(o + (size_t)(n + 2) * 8, 256);
    L.nodes = o; o = align_up(o + ((size_t)2 << L.max_depth) * 8 + 64, 256);
    L.total = o;
    return L;
}

__device__ __forceinline__ uint32_t desc_key(float f) {
    uint32_t b = __float_as_uint(f);
    if (f == 0.0f) b = 0u;                                       // -0.0 and +0.0 are one threshold
    const uint32_t asc = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    return ~asc;                                                 // ascending key order == descending score
}
__device__ __forceinline__ float key_to_score(uint32_t key) {
    const uint32_t asc = ~key;
    const uint32_t b = (asc & 0x80000000u) ? (asc ^ 0x80000000u) : ~asc;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ------------------------------------------------------------------------------------------ 1 keys
template <typename T>
__global__ void __launch_bounds__(kKeysThreads)
auc_keys_kernel(const T* __restrict__ scores, const int64_t* __restrict__ labels, int64_t n, int flags,
                uint32_t* __restrict__ keys, uint8_t* __restrict__ labs, AucControl* c) {
    pdl_enter();
    __shared__ uint32_t s_hist[4 * 256];
    __shared__ unsigned int s_cnt[3];
    for (int i = threadIdx.x; i < 1024; i += kKeysThreads) s_hist[i] = 0;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned int nv = 0, np = 0, bad = 0;
    const bool ignore_neg = flags & EOE_AUC_IGNORE_NEGATIVE_LABELS;
#if EOE_AUC_KEYS_V2
    // One block of 1024 threads per SM at most: every block ends with up to 1024 atomics onto the same 4 KB of global
    // counters, so fewer, larger blocks (148 x 1024 adds instead of 489 x 1024 at 1 M scores); kKeysBatch rows per thread
    // are loaded before the first is used (48 KB in flight per SM)
    for (int64_t i0 = (int64_t)blockIdx.x * (kKeysThreads * kKeysBatch); i0 < n; i0 += (int64_t)gridDim.x * (kKeysThreads * kKeysBatch)) {
        float fv[kKeysBatch];
        int64_t lv[kKeysBatch];
#pragma unroll
        for (int q = 0; q < kKeysBatch; ++q) {
            const int64_t i = i0 + q * kKeysThreads + threadIdx.x;
            fv[q] = (i < n) ? to_f32<T>(scores[i]) : 0.f;
            lv[q] = (i < n) ? labels[i] : 0;
        }
#pragma unroll
        for (int q = 0; q < kKeysBatch; ++q) {
            const int64_t i = i0 + q * kKeysThreads + threadIdx.x;
            if (i >= n) continue;
            const float f = fv[q];
            const int64_t l = lv[q];
#else
    {
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
            const float f = to_f32<T>(scores[i]);
            const int64_t l = labels[i];
#endif
            const bool valid = !(ignore_neg && l < 0);
            uint32_t key = 0xffffffffu;      // dropped rows sort behind every finite score
            uint8_t lb = 0;
            if (valid) {
                if (!isfinite(f)) bad = 1;
                key = desc_key(f);
                lb = (l == 1);
                nv++;
                np += lb;
            }
            keys[i] = key;
            labs[i] = lb;
            atomicAdd(&s_hist[key & 255], 1u);
            atomicAdd(&s_hist[256 + ((key >> 8) & 255)], 1u);
            atomicAdd(&s_hist[512 + ((key >> 16) & 255)], 1u);
            atomicAdd(&s_hist[768 + (key >> 24)], 1u);
        }
    }
    nv = __reduce_add_sync(kFullMask, nv);
    np = __reduce_add_sync(kFullMask, np);
    bad = __reduce_or_sync(kFullMask, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], nv);
        atomicAdd(&s_cnt[1], np);
        atomicOr(&s_cnt[2], bad);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += kKeysThreads)
        if (s_hist[i]) atomicAdd(&c->hist[i], s_hist[i]);
    if (threadIdx.x == 0) {
        atomicAdd(&c->n_valid, (unsigned long long)s_cnt[0]);
        atomicAdd(&c->n_pos, (unsigned long long)s_cnt[1]);
        if (s_cnt[2]) atomicOr(&c->status, (uint32_t)EOE_AUC_STATUS_NONFINITE);
    }
}

// block-wide exclusive scan of one u32 per thread (256 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_tmp /*[8]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(kFullMask, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_tmp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t t = s_tmp[w];
        if (w < warp) woff += t;
        tot += t;
    }
    __syncthreads();
    if (total) *total = tot;
    return woff + inc - v;
}

// ------------------------------------------------------------------------------------------ 2 sort pass
// Scatter: with a random 8-bit digit the 32 keys of a warp belong to ~28 different digit runs, so writing them straight to
// their global positions costs one 32-byte sector per 4-byte key and another per label byte (measured: 31 us per pass at
// 1 M scores against 15 us for the top-digit pass, whose runs are long).  The tile is therefore reordered by digit in
// shared memory first (local position = start of the digit's run within the tile + rank), and written out in local
// order: the keys of a digit are consecutive there AND at their destination, so a warp's stores cover whole sectors.
#ifndef EOE_AUC_DIRECT_SCATTER
#define EOE_AUC_DIRECT_SCATTER 0                        // 1: the round-1 form (registers -> global), kept for A/B builds
#endif
template <int kSortItems>
__global__ void __launch_bounds__(kSortThreads)
auc_sort_pass_kernel(const uint32_t* __restrict__ keys_in, const uint8_t* __restrict__ labs_in,
                     uint32_t* __restrict__ keys_out, uint8_t* __restrict__ labs_out, int64_t n, int pass,
                     AucControl* c, uint32_t* status /* [tiles][256] for this pass */) {
    pdl_enter();
    __shared__ uint32_t s_warp_hist[8][256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_tmp[8];
    __shared__ unsigned int s_tile;
    constexpr int kSortTile = kSortThreads * kSortItems;
#if !EOE_AUC_DIRECT_SCATTER
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint8_t s_labs[kSortTile];
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&c->tickets[pass], 1u);
    for (int i = tid; i < 8 * 256; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned int tile = s_tile;
    const int shift = pass * 8;
    const int64_t base = (int64_t)tile * kSortTile + warp * (32 * kSortItems);
    uint32_t key[kSortItems];
    uint8_t lab[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const int64_t idx = base + j * 32 + lane;
        const bool valid = idx < n;
        key[j] = valid ? __ldg(keys_in + idx) : 0xffffffffu;
        lab[j] = valid ? __ldg(labs_in + idx) : (uint8_t)0;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const bool valid = (base + j * 32 + lane) < n;
        const uint32_t d = (key[j] >> shift) & 255u;
#if EOE_AUC_BALLOT_RANK
        uint32_t mask = __ballot_sync(kFullMask, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(kFullMask, bit);
            mask &= bit ? bal : ~bal;
        }
        if (!valid) mask = 1u << lane;
#else
        const uint32_t mask = __match_any_sync(kFullMask, valid ? d : (256u + lane));
#endif
        const int leader = __ffs(mask) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + __popc(mask);
        }
        old = __shfl_sync(kFullMask, old, leader);
        rank[j] = (uint16_t)(old + __popc(mask & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    // thread = digit: exclusive prefix over the 8 warps, tile count, global digit start, look-back
    const uint32_t d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t t = s_warp_hist[w][d];
        s_warp_hist[w][d] = run;
        run += t;
    }
    const uint32_t gstart = block_excl_scan_256(c->hist[pass * 256 + d], s_tmp, nullptr);
    uint32_t* my = status + (size_t)tile * 256 + d;
    // published before the local reordering: successors can use the aggregate (tile 0: the inclusive value) from here on
    st_volatile_u32(my, run | ((tile == 0 ? kFlagIncl : kFlagAgg) << 30));
#if !EOE_AUC_DIRECT_SCATTER
    // start of digit d's run within the tile, folded into the per-warp offsets; then every key goes to its local position
    const uint32_t lstart = block_excl_scan_256(run, s_tmp, nullptr);
#pragma unroll
    for (int w = 0; w < 8; ++w) s_warp_hist[w][d] += lstart;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        if ((base + j * 32 + lane) < n) {
            const uint32_t lp = s_warp_hist[warp][(key[j] >> shift) & 255u] + rank[j];
            s_keys[lp] = key[j];
            s_labs[lp] = lab[j];
        }
    }
#else
    const uint32_t lstart = 0;
#endif
    uint32_t excl = 0;
    if (tile != 0) {
        int64_t t = (int64_t)tile - 1;
        bool done = false;
        int spins = 0;
        while (!done) {
            // batches of up to 8 independent (volatile) loads so that the walk is latency- not chain-bound
            uint32_t v[8];
#pragma unroll
            for (int b = 0; b < 8; ++b)
                v[b] = (t - b >= 0) ? ld_volatile_u32(status + (size_t)(t - b) * 256 + d) : (kFlagIncl << 30);
            int consumed = 0;
            bool stall = false;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const uint32_t f = v[b] >> 30;
                if (!done && !stall) {
                    if (f == 0) {
                        stall = true;                  // predecessor not published yet: re-poll from here
                    } else {
                        excl += v[b] & 0x3fffffffu;
                        ++consumed;
                        if (f == kFlagIncl) done = true;
                    }
                }
            }
            t -= consumed;
            if (stall && ++spins > kSpinLimit) { atomicOr(&c->status, 0x100u); done = true; }
        }
        st_volatile_u32(my, (excl + run) | (kFlagIncl << 30));
    }
    s_base[d] = gstart + excl - lstart;                // global position of local position lp of digit d: s_base[d] + lp
    __syncthreads();
#if !EOE_AUC_DIRECT_SCATTER
    const int64_t left = n - (int64_t)tile * kSortTile;
    const int cnt = left < kSortTile ? (int)left : kSortTile;          // valid keys of the tile = local positions in use
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const int lp = j * kSortThreads + tid;
        if (lp < cnt) {
            const uint32_t k = s_keys[lp];
            const uint32_t pos = s_base[(k >> shift) & 255u] + (uint32_t)lp;
            keys_out[pos] = k;
            labs_out[pos] = s_labs[lp];
        }
    }
#else
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        if ((base + j * 32 + lane) < n) {
            const uint32_t dj = (key[j] >> shift) & 255u;
            const uint32_t pos = s_base[dj] + s_warp_hist[warp][dj] + rank[j];
            keys_out[pos] = key[j];
            labs_out[pos] = lab[j];
        }
    }
#endif
}

// Warp-parallel decoupled look-back over packed 64-bit tile states: [63:62] flag, [61:31] a, [30:0] b.
// Called by warp 0 of the block; returns the exclusive prefix (a, b) and publishes the inclusive state.
__device__ __forceinline__ void lookback_pair(unsigned long long* status, unsigned int tile, uint32_t agg_a,
                                              uint32_t agg_b, uint32_t& excl_a, uint32_t& excl_b, AucControl* c) {
    const int lane = threadIdx.x & 31;
    auto pack = [](uint32_t f, uint32_t a, uint32_t b) {
        return ((unsigned long long)f << 62) | ((unsigned long long)a << 31) | (unsigned long long)b;
    };
    excl_a = 0; excl_b = 0;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(status, pack(kFlagIncl, agg_a, agg_b));
        return;
    }
    if (lane == 0) st_volatile_u64(status + tile, pack(kFlagAgg, agg_a, agg_b));
    int64_t t = (int64_t)tile - 1;      // lane 0 looks at t, lane 1 at t-1, ...
    int spins = 0;
    while (true) {
        const int64_t mine = t - lane;
        unsigned long long v = (mine >= 0) ? ld_volatile_u64(status + mine) : pack(kFlagIncl, 0, 0);
        const uint32_t f = (uint32_t)(v >> 62);
        const unsigned pending = __ballot_sync(kFullMask, f == 0);
        const unsigned incl = __ballot_sync(kFullMask, f == kFlagIncl);
        // usable window: lanes before the first pending one, cut after the first inclusive one
        int limit = pending ? (__ffs(pending) - 1) : 32;
        const int first_incl = incl ? (__ffs(incl) - 1) : 32;
        const bool finish = first_incl < limit;
        if (finish) limit = first_incl + 1;
        uint32_t a = (lane < limit) ? (uint32_t)((v >> 31) & 0x7fffffffu) : 0u;
        uint32_t b = (lane < limit) ? (uint32_t)(v & 0x7fffffffu) : 0u;
        excl_a += __reduce_add_sync(kFullMask, a);
        excl_b += __reduce_add_sync(kFullMask, b);
        if (finish) break;
        t -= limit;
        if (limit == 0 && ++spins > kSpinLimit) { if (lane == 0) atomicOr(&c->status, 0x100u); break; }
    }
    if (lane == 0) st_volatile_u64(status + tile, pack(kFlagIncl, excl_a + agg_a, excl_b + agg_b));
}

// ------------------------------------------------------------------------------------------ 3 distinct
// sklearn _binary_clf_curve: distinct_value_indices = where(diff(sorted_scores)); threshold_idxs = r_[., n-1];
// tps = cumsum(y)[idxs]; fps = 1 + idxs - tps.
// Layout of both scan kernels: a warp owns 32 * kScanItems consecutive elements, lane l the elements l, 32 + l, ... of
// them (every load a full 128-byte line, neighbours by shuffle).  Flags and label bits are single bits, so the scan inside
// a warp is ballot + popc, and the kept elements of a warp leave as dense runs (lane order = element order): coalesced
// stores.  (The thread-contiguous form this replaces touched 32 sectors per load and per store instruction and took
// 24 + 18 us at 1 M scores.)
// exclusive prefix over the 8 warps of (a, b) given per warp by lane 0; totals in tot_a / tot_b
__device__ __forceinline__ void warp_totals_excl(uint32_t a, uint32_t b, uint32_t (*s_w)[8], uint32_t& ex_a, uint32_t& ex_b,
                                                 uint32_t& tot_a, uint32_t& tot_b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_w[0][warp] = a; s_w[1][warp] = b; }
    __syncthreads();
    ex_a = ex_b = tot_a = tot_b = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t ta = s_w[0][w], tb = s_w[1][w];
        if (w < warp) { ex_a += ta; ex_b += tb; }
        tot_a += ta; tot_b += tb;
    }
}

__global__ void __launch_bounds__(kScanThreads)
auc_distinct_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ labs, AucControl* c,
                    unsigned long long* status, uint32_t* __restrict__ d_tps, uint32_t* __restrict__ d_fps,
                    uint32_t* __restrict__ d_key) {
    pdl_enter();
    __shared__ uint32_t s_w[2][8];
    __shared__ uint32_t s_excl[2];
    __shared__ unsigned int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&c->tickets[4], 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t nv = (int64_t)c->n_valid;
    if ((int64_t)tile * kScanTile >= nv) return;          // whole tile beyond the kept rows (uniform per block)
    const int64_t wbase = (int64_t)tile * kScanTile + warp * (32 * kScanItems);
    uint32_t k[kScanItems], lb[kScanItems], fb[kScanItems];            // lb / fb: warp-uniform ballots (lb: the label first)
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        k[j] = (i < nv) ? keys[i] : 0u;
        lb[j] = (i < nv) ? (uint32_t)labs[i] : 0u;        // (all loads go out before the first vote below)
    }
    const int64_t after = wbase + 32 * kScanItems;                      // first element of the next warp's chunk
    const uint32_t k_after = (after < nv) ? keys[after] : 0u;
    uint32_t fc = 0, lc = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        const uint32_t dn = __shfl_down_sync(kFullMask, k[j], 1);
        const uint32_t first_next = (j + 1 < kScanItems) ? __shfl_sync(kFullMask, k[(j + 1 < kScanItems) ? j + 1 : j], 0) : k_after;
        const uint32_t kn = (lane == 31) ? first_next : dn;
        const bool f = (i < nv) && (i == nv - 1 || k[j] != kn);
        fb[j] = __ballot_sync(kFullMask, f);
        lb[j] = __ballot_sync(kFullMask, lb[j] != 0u);
        fc += __popc(fb[j]);
        lc += __popc(lb[j]);
    }
    uint32_t ex_f, ex_l, tot_f, tot_l;
    warp_totals_excl(fc, lc, s_w, ex_f, ex_l, tot_f, tot_l);
    if (tid < 32) {
        uint32_t ea, eb;
        lookback_pair(status, tile, tot_f, tot_l, ea, eb, c);
        if (tid == 0) { s_excl[0] = ea; s_excl[1] = eb; }
    }
    __syncthreads();
    uint32_t slot0 = s_excl[0] + ex_f;                   // slot of the warp's next kept element
    uint32_t tps0 = s_excl[1] + ex_l;                    // positives before the warp's next row of 32
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if ((fb[j] >> lane) & 1u) {
            const int64_t i = wbase + j * 32 + lane;
            const uint32_t slot = slot0 + __popc(fb[j] & lt_mask);
            const uint32_t tps = tps0 + __popc(lb[j] & le_mask);
            d_tps[slot] = tps;
            d_fps[slot] = 1u + (uint32_t)i - tps;
            d_key[slot] = k[j];
            if (i == nv - 1) c->n_distinct = slot + 1;    // the last row always closes a run
        }
        slot0 += __popc(fb[j]);
        tps0 += __popc(lb[j]);
    }
}

// ------------------------------------------------------------------------------------------ 4 corners
// roc_curve(drop_intermediate=True): optimal_idxs = where(r_[True, logical_or(diff(fps,2), diff(tps,2)), True])
// (only if len(fps) > 2); then tps = r_[0, tps], fps = r_[0, fps], thresholds = r_[inf, thresholds].
__global__ void __launch_bounds__(kScanThreads)
auc_corner_kernel(const uint32_t* __restrict__ d_tps, const uint32_t* __restrict__ d_fps,
                  const uint32_t* __restrict__ d_key, AucControl* c, unsigned long long* status,
                  uint32_t* __restrict__ k_tps, uint32_t* __restrict__ k_fps, float* __restrict__ thr_out) {
    pdl_enter();
    __shared__ uint32_t s_w[2][8];
    __shared__ uint32_t s_excl;
    __shared__ unsigned int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&c->tickets[5], 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t m = (int64_t)c->n_distinct;
    if ((int64_t)tile * kScanTile >= m) return;
    if (tile == 0 && tid == 0) {
        k_tps[0] = 0; k_fps[0] = 0;
        if (thr_out) thr_out[0] = INFINITY;
    }
    const int64_t wbase = (int64_t)tile * kScanTile + warp * (32 * kScanItems);
    uint32_t tp[kScanItems], fp[kScanItems], kb[kScanItems];           // kb: warp-uniform ballot of the kept points
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        tp[j] = (i < m) ? d_tps[i] : 0u;
        fp[j] = (i < m) ? d_fps[i] : 0u;
    }
    // the points just before and just after the warp's chunk (values outside [0, m) are never used: the first and the
    // last point are kept unconditionally)
    const int64_t before = wbase - 1, after = wbase + 32 * kScanItems;
    const uint32_t tp_before = (before >= 0 && before < m) ? d_tps[before] : 0u, fp_before = (before >= 0 && before < m) ? d_fps[before] : 0u;
    const uint32_t tp_after = (after < m) ? d_tps[after] : 0u, fp_after = (after < m) ? d_fps[after] : 0u;
    uint32_t kc = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        const uint32_t tu = __shfl_up_sync(kFullMask, tp[j], 1), fu = __shfl_up_sync(kFullMask, fp[j], 1);
        const uint32_t td = __shfl_down_sync(kFullMask, tp[j], 1), fd = __shfl_down_sync(kFullMask, fp[j], 1);
        const uint32_t tl = (j > 0) ? __shfl_sync(kFullMask, tp[(j > 0) ? j - 1 : 0], 31) : tp_before;
        const uint32_t fl = (j > 0) ? __shfl_sync(kFullMask, fp[(j > 0) ? j - 1 : 0], 31) : fp_before;
        const uint32_t tn = (j + 1 < kScanItems) ? __shfl_sync(kFullMask, tp[(j + 1 < kScanItems) ? j + 1 : j], 0) : tp_after;
        const uint32_t fn = (j + 1 < kScanItems) ? __shfl_sync(kFullMask, fp[(j + 1 < kScanItems) ? j + 1 : j], 0) : fp_after;
        const int64_t t0 = (lane == 0) ? tl : tu, f0 = (lane == 0) ? fl : fu;       // point i - 1
        const int64_t t2 = (lane == 31) ? tn : td, f2 = (lane == 31) ? fn : fd;     // point i + 1
        const int64_t t1 = tp[j], f1 = fp[j];
        bool kp = false;
        if (i < m) kp = (m <= 2) || i == 0 || i == m - 1 || (f0 - 2 * f1 + f2 != 0) || (t0 - 2 * t1 + t2 != 0);
        kb[j] = __ballot_sync(kFullMask, kp);
        kc += __popc(kb[j]);
    }
    uint32_t ex, ex_unused, tot, tot_unused;
    warp_totals_excl(kc, 0u, s_w, ex, ex_unused, tot, tot_unused);
    if (tid < 32) {
        uint32_t ea, eb;
        lookback_pair(status, tile, tot, 0u, ea, eb, c);
        if (tid == 0) s_excl = ea;
    }
    __syncthreads();
    uint32_t slot0 = s_excl + ex + 1;           // +1: the prepended origin
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if ((kb[j] >> lane) & 1u) {
            const int64_t i = wbase + j * 32 + lane;
            const uint32_t slot = slot0 + __popc(kb[j] & lt_mask);
            k_tps[slot] = tp[j];
            k_fps[slot] = fp[j];
            if (thr_out) thr_out[slot] = key_to_score(d_key[i]);
            if (i == m - 1) c->n_kept = slot;     // the last point is always kept; the count excludes the origin
        }
        slot0 += __popc(kb[j]);
    }
}

// ------------------------------------------------------------------------------------------ 5 terms
// fpr = fps / fps[-1]; tpr = tps / tps[-1]; np.trapezoid integrand d * (y[1:] + y[:-1]) / 2.0 (fp64, this order)
__global__ void __launch_bounds__(256)
auc_terms_kernel(const uint32_t* __restrict__ k_tps, const uint32_t* __restrict__ k_fps, AucControl* c,
                 double* __restrict__ terms, double* __restrict__ fpr_out, double* __restrict__ tpr_out) {
    pdl_enter();
    const int64_t P = (int64_t)c->n_kept + 1;        // points incl. origin
    const double ftot = (double)(c->n_valid - c->n_pos);
    const double ttot = (double)c->n_pos;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < P; i += (int64_t)gridDim.x * 256) {
        const double f0 = __ddiv_rn((double)k_fps[i], ftot), t0 = __ddiv_rn((double)k_tps[i], ttot);
        if (fpr_out) { fpr_out[i] = f0; tpr_out[i] = t0; }
        if (i + 1 < P) {
            const double f1 = __ddiv_rn((double)k_fps[i + 1], ftot), t1 = __ddiv_rn((double)k_tps[i + 1], ttot);
            terms[i] = __ddiv_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 2.0);
        }
    }
}

// ------------------------------------------------------------------------------------------ 6/7 pairwise sum
__device__ __forceinline__ int dev_pairwise_depth(int64_t n) {
    int d = 0;
    while (n > 128) { n -= (n >> 1) & ~(int64_t)7; ++d; }
    return d;
}

// One group of 8 lanes per leaf slot g of the virtual complete tree of depth D over T terms; lane j owns accumulator r[j].
// All 8 lanes of a group take the same path.
// (no __restrict__ on `a`: the single-launch kernel reads terms it wrote itself -- they must not go through the
// non-coherent load path)
__device__ __forceinline__ void pairwise_leaf_slot(const double* a, int64_t T, int D, int64_t g, int j,
                                                   unsigned gmask, double* nodes) {
    if (g >= ((int64_t)1 << D)) return;
    int64_t off = 0, len = T;
    int lvl = 0;
    while (len > 128) {
        const int64_t h = (len >> 1) & ~(int64_t)7;
        if ((g >> (D - 1 - lvl)) & 1) { off += h; len -= h; } else len = h;
        ++lvl;
    }
    const int rem = D - lvl;
    if (g & (((int64_t)1 << rem) - 1)) return;        // this leaf is represented by its left-most slot only
    const int64_t heap = ((int64_t)1 << lvl) + (g >> rem);
    const double* p = a + off;
    double res;
    if (len < 8) {
        if (j != 0) return;
        res = 0.0;
        for (int64_t i = 0; i < len; ++i) res = __dadd_rn(res, p[i]);
        nodes[heap] = res;
        return;
    }
    double r = p[j];
    const int64_t body = len - (len & 7);
#pragma unroll 8
    for (int64_t i = 8; i < body; i += 8) r = __dadd_rn(r, p[i + j]);      // (unrolled: the loads of a batch go out together)
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
    double o = __shfl_xor_sync(gmask, r, 1);
    r = (j & 1) ? __dadd_rn(o, r) : __dadd_rn(r, o);
    o = __shfl_xor_sync(gmask, r, 2);
    r = (j & 2) ? __dadd_rn(o, r) : __dadd_rn(r, o);
    o = __shfl_xor_sync(gmask, r, 4);
    r = (j & 4) ? __dadd_rn(o, r) : __dadd_rn(r, o);
    if (j == 0) {
        res = r;
        for (int64_t i = body; i < len; ++i) res = __dadd_rn(res, p[i]);
        nodes[heap] = res;
    }
}

__global__ void __launch_bounds__(256)
pairwise_leaves_kernel(const double* __restrict__ a, const unsigned long long* n_ptr, int64_t n_minus,
                       double* __restrict__ nodes) {
    pdl_enter();
    const int64_t T = (int64_t)(*n_ptr) - n_minus;   // number of terms
    if (T <= 0) { if (blockIdx.x == 0 && threadIdx.x == 0) nodes[1] = 0.0; return; }
    const int D = dev_pairwise_depth(T);
    const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;
    pairwise_leaf_slot(a, T, D, g, threadIdx.x & 7, 0xffu << (threadIdx.x & 24), nodes);
}

// internal nodes of the numpy pairwise tree, bottom-up, by one block of NT threads (barrier per level)
// levels D - 1 ... stop (a node's existence and heap slot depend on T and its own level only)
template <int NT>
__device__ __forceinline__ void pairwise_tree_levels(int64_t T, int D, double* nodes, int stop = 0) {
    for (int lvl = D - 1; lvl >= stop; --lvl) {
        for (int64_t p = threadIdx.x; p < ((int64_t)1 << lvl); p += NT) {
            int64_t len = T;
            bool exists = true;
            for (int s = 0; s < lvl; ++s) {
                if (len <= 128) { exists = false; break; }
                const int64_t h = (len >> 1) & ~(int64_t)7;
                len = ((p >> (lvl - 1 - s)) & 1) ? (len - h) : h;
            }
            if (exists && len > 128) {
                const int64_t heap = ((int64_t)1 << lvl) + p;
                nodes[heap] = __dadd_rn(nodes[2 * heap], nodes[2 * heap + 1]);
            }
        }
        __syncthreads();
    }
}

// Internal nodes bottom-up (single block); finally writes the result (NaN if the curve is undefined).
// The top kTreeSmemDepth levels run in shared memory: in global memory every level is a dependent load + store + barrier
// (an L2 round trip per level, 13 of them at 1 M terms) and every node re-derives its length by walking down from the
// root.  Here the heap [1, 2 << Dc) is fetched once (batched loads), the node lengths are produced top-down (a child's
// length follows from its parent's), and the sums go bottom-up, all behind __syncthreads(): same additions, same order.
// Levels >= kTreeSmemDepth (more than ~2 M terms) are reduced in global memory first.
constexpr int kTreeSmemDepth = 13;              // 2 << 13 doubles (128 KB) + 1 << 13 lengths (32 KB)
__global__ void __launch_bounds__(1024)
pairwise_tree_kernel(const unsigned long long* n_ptr, int64_t n_minus, double* __restrict__ nodes, AucControl* c,
                     double* __restrict__ out, int negate_clip, int64_t* info /* last launch of a call: counts + status */) {
    pdl_enter();
    extern __shared__ double s_nodes[];
    const int64_t T = (int64_t)(*n_ptr) - n_minus;
    const int D = (T > 0) ? dev_pairwise_depth(T) : 0;
    const int Dc = D < kTreeSmemDepth ? D : kTreeSmemDepth;
    uint32_t* s_len = reinterpret_cast<uint32_t*>(s_nodes + ((size_t)2 << Dc));   // [1 << Dc): lengths of levels < Dc
    pairwise_tree_levels<1024>(T, D, nodes, Dc);                     // (no-op unless D > kTreeSmemDepth; ends with a barrier)
    const int total = 2 << Dc;
    for (int i0 = 0; i0 < total; i0 += 8 * 1024) {                    // up to 8 independent loads per thread in flight
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * 1024 + threadIdx.x;
            v[q] = (i >= 1 && i < total) ? nodes[i] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * 1024 + threadIdx.x;
            if (i >= 1 && i < total) s_nodes[i] = v[q];
        }
    }
    if (threadIdx.x == 0 && Dc > 0) s_len[1] = (uint32_t)T;           // (Dc > 0 implies 128 < T; T < 2^31)
    __syncthreads();
    for (int lvl = 0; lvl + 1 < Dc; ++lvl) {                         // lengths of level lvl + 1 from level lvl
        for (int p = threadIdx.x; p < (1 << lvl); p += 1024) {
            const int heap = (1 << lvl) + p;
            const uint32_t len = s_len[heap];
            const uint32_t h = (len > 128u) ? ((len >> 1) & ~7u) : 0u;   // 0: no such node (or a leaf's absent children)
            s_len[2 * heap] = h;
            s_len[2 * heap + 1] = (len > 128u) ? (len - h) : 0u;
        }
        __syncthreads();
    }
    for (int lvl = Dc - 1; lvl >= 0; --lvl) {
        for (int p = threadIdx.x; p < (1 << lvl); p += 1024) {
            const int heap = (1 << lvl) + p;
            if (s_len[heap] > 128u) s_nodes[heap] = __dadd_rn(s_nodes[2 * heap], s_nodes[2 * heap + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) nodes[1] = s_nodes[1];
    if (threadIdx.x == 0) {
        double v = nodes[1];
        if (negate_clip) v = fmax(0.0, -v);
        const bool single = (c->n_pos == 0) || (c->n_pos == c->n_valid);
        if (single) atomicOr(&c->status, (uint32_t)EOE_AUC_STATUS_SINGLE_CLASS);
        if (single || (c->status & ~(uint32_t)EOE_AUC_STATUS_SINGLE_CLASS)) v = __longlong_as_double(0x7ff8000000000000LL);
        *out = v;
        if (info) {
            info[0] = (int64_t)c->n_valid; info[1] = (int64_t)c->n_pos; info[2] = (int64_t)c->n_distinct;
            info[3] = (int64_t)c->n_kept + 1; info[4] = (int64_t)c->status; info[5] = info[6] = info[7] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------ PRC / AP
// precision_recall_curve (reversed, (1,0) appended) and average_precision_score = max(0, -sum(diff(recall)*precision[:-1])).
// With m distinct thresholds j = 0..m-1 (descending score): reversed index i = m-1-j.
__global__ void __launch_bounds__(256)
auc_prc_terms_kernel(const uint32_t* __restrict__ d_tps, const uint32_t* __restrict__ d_fps,
                     const uint32_t* __restrict__ d_key, AucControl* c, double* __restrict__ terms,
                     double* __restrict__ prec_out, double* __restrict__ rec_out, float* __restrict__ pthr_out) {
    pdl_enter();
    const int64_t m = (int64_t)c->n_distinct;
    const double ttot = (double)c->n_pos;
    auto prec = [&](int64_t j) {
        const double tp = (double)d_tps[j], ps = __dadd_rn(tp, (double)d_fps[j]);
        return ps != 0.0 ? __ddiv_rn(tp, ps) : 0.0;
    };
    auto rec = [&](int64_t j) { return ttot == 0.0 ? 1.0 : __ddiv_rn((double)d_tps[j], ttot); };
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (int64_t)gridDim.x * 256) {
        const int64_t j = m - 1 - i;
        const double p = prec(j), r = rec(j);
        const double r_next = (i + 1 < m) ? rec(j - 1) : 0.0;
        terms[i] = __dmul_rn(__dsub_rn(r_next, r), p);
        if (prec_out) { prec_out[i] = p; rec_out[i] = r; }
        if (pthr_out) pthr_out[i] = key_to_score(d_key[j]);     // thresholds[::-1]: the distinct scores, increasing
    }
    if (prec_out && blockIdx.x == 0 && threadIdx.x == 0) { prec_out[m] = 1.0; rec_out[m] = 0.0; }
}


// ------------------------------------------------------------------------------------------ single-launch path
// The reference evaluates the AUC on 3 000 - 10 000 scores per class and epoch (ad_trainer.py:452-455, 516-522): there the
// multi-kernel pipeline above is pure launch latency (10+ launches, ~0.1 ms).  For n <= kSmallMax the whole computation
// runs in ONE launch of ONE CTA with everything but the fp64 terms in shared memory: keys and label bits, 4 (or fewer)
// LSD radix passes, the tie / corner scans, then the fp64 terms and numpy's pairwise tree, phase after phase behind
// __syncthreads().  Every arithmetic step is the one of the multi-kernel path (same __d*_rn sequence, same tree), so the
// result is bit-identical to it and to scikit-learn.
// Ranking uses eight warp ballots per key (the digit's peer mask is the AND of the per-bit ballots): match.any has a
// throughput of one warp instruction per ~64 clocks per SM, which made a one-SM sort 10x slower than the ballots.
constexpr int kSmallThreads = 1024;
constexpr int kSmallItems = 16;
constexpr int kSmallMax = kSmallThreads * kSmallItems;        // 16 384 scores

struct SmallShared {
    uint32_t keys[kSmallMax + 8];      // sorted keys; after the distinct phase: kept ROC points packed (tps << 16 | fps)
    uint16_t d_tps[kSmallMax];         // distinct thresholds: true positives / false positives (<= 16 384: 16 bits)
    uint16_t d_fps[kSmallMax];
    uint32_t cnt[8 * kSmallThreads + 8 * kSmallThreads / 32];   // 16 digit counters per thread, two per word, padded
    uint8_t labs[kSmallMax];           // label bits; after the distinct phase: keep flags of the corner filter
    uint32_t scan_tmp[32];
    uint32_t and_all, or_all, n_valid, n_pos, status, pad[3];
    double nodes[1024];                // pairwise tree: depth <= 8 for <= 16 385 terms
};
static_assert(sizeof(SmallShared) <= 227 * 1024, "single-launch AUC: shared memory budget");

// exclusive prefix of one u32 per thread over the 1024 threads of the block; *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* s_tmp /*[32]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_tmp[warp] = inc;
    __syncthreads();
    const uint32_t wt = s_tmp[lane];
    uint32_t winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, winc, o);
        if (lane >= o) winc += t;
    }
    const uint32_t woff = __shfl_sync(kFullMask, winc - wt, warp);
    if (total) *total = __shfl_sync(kFullMask, winc, 31);
    __syncthreads();
    return woff + inc - v;
}

// numpy pairwise sum of a[0..T) by the whole block (nodes: shared memory, >= 2 << depth doubles); result in nodes[1]
__device__ __forceinline__ double block_pairwise_sum(const double* a, int64_t T, double* nodes) {
    if (T <= 0) return 0.0;
    const int D = dev_pairwise_depth(T);
    for (int64_t g0 = 0; g0 < ((int64_t)1 << D); g0 += kSmallThreads / 8)
        pairwise_leaf_slot(a, T, D, g0 + (threadIdx.x >> 3), threadIdx.x & 7, 0xffu << (threadIdx.x & 24), nodes);
    __syncthreads();
    pairwise_tree_levels<kSmallThreads>(T, D, nodes);
    const double v = nodes[1];
    __syncthreads();
    return v;
}

// MAXS: compile-time bound of the keys a thread owns (ceil(n / 1024) | 1): the per-key loops are fully unrolled, so small
// inputs get an instantiation without the dead iterations
template <typename T, int MAXS>
__global__ void __launch_bounds__(kSmallThreads, 1)
auc_small_kernel(const T* __restrict__ scores, const int64_t* __restrict__ labels, int n, int flags,
                 uint32_t* __restrict__ d_key, double* terms, double* __restrict__ auc_out,
                 int64_t* __restrict__ info_out, double* __restrict__ fpr_out, double* __restrict__ tpr_out,
                 float* __restrict__ thr_out, double* __restrict__ prec_out, double* __restrict__ rec_out,
                 float* __restrict__ pthr_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int items = (n + kSmallThreads - 1) / kSmallThreads;      // <= kSmallItems, uniform
    const int wbase = warp * items * 32;                             // this warp's contiguous chunk
    if (tid == 0) { sh.and_all = 0xffffffffu; sh.or_all = 0u; sh.n_valid = 0; sh.n_pos = 0; sh.status = 0; }
    __syncthreads();
    const long long clk0 = clock64();       // phase clocks (thread 0) -> info_out[5..7]: diagnostics of this one-CTA pipeline

    // ---- 1 keys: score -> descending-sortable key, label bit, counts, constant-bit masks (to skip radix passes)
    {
        const bool ignore_neg = flags & EOE_AUC_IGNORE_NEGATIVE_LABELS;
        uint32_t nv = 0, np = 0, bad = 0, a_and = 0xffffffffu, a_or = 0u;
        constexpr int NI = MAXS < kSmallItems ? MAXS : kSmallItems;  // items <= NI
        float f[NI];
        int64_t l[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j) {                               // all loads first: one memory round trip
            const int idx = wbase + j * 32 + lane;
            const bool in = j < items && idx < n;
            f[j] = in ? to_f32<T>(scores[idx]) : 0.f;
            l[j] = in ? labels[idx] : 0;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int idx = wbase + j * 32 + lane;
            if (j < items && idx < n) {
                uint32_t key = 0xffffffffu;          // dropped rows sort behind every finite score
                uint8_t lb = 0;
                if (!(ignore_neg && l[j] < 0)) {
                    if (!isfinite(f[j])) bad = 1;
                    key = desc_key(f[j]);
                    lb = (l[j] == 1);
                    nv++;
                    np += lb;
                }
                sh.keys[idx] = key;
                sh.labs[idx] = lb;
                a_and &= key;
                a_or |= key;
            }
        }
        nv = __reduce_add_sync(kFullMask, nv);
        np = __reduce_add_sync(kFullMask, np);
        bad = __reduce_or_sync(kFullMask, bad);
        a_and = __reduce_and_sync(kFullMask, a_and);
        a_or = __reduce_or_sync(kFullMask, a_or);
        if (lane == 0) {
            atomicAdd(&sh.n_valid, nv);
            atomicAdd(&sh.n_pos, np);
            if (bad) atomicOr(&sh.status, (uint32_t)EOE_AUC_STATUS_NONFINITE);
            atomicAnd(&sh.and_all, a_and);
            atomicOr(&sh.or_all, a_or);
        }
    }
    __syncthreads();
    const uint32_t varying = sh.and_all ^ sh.or_all;               // bits that differ between at least two keys

    // ---- 2 sort: LSD radix, 4 bits per pass, in place in shared memory.  Thread t owns the S consecutive keys
    // [t * S, t * S + S) (S odd: conflict-free) and ranks them with 16 thread-private counters packed two per word
    // (digit w in the low half, digit w + 8 in the high half of word w: no warp votes, no atomics -- one shared-memory
    // read-modify-write per key); a raking scan over the (digit, thread)-ordered counters turns them into positions.
    // Passes whose four bits are the same in every key are skipped (fp16 scores: 13 constant low bits).
    {
        const int S = items | 1;
        const int kbase = tid * S;
        // counter of digit d for thread t: 16-bit element 2 * (w * 1056 + t + t / 32) + (d >> 3), w = d & 7  (word w * 1024
        // + t of the (digit word, thread) order, one padding word per 32)
        uint16_t* c16 = reinterpret_cast<uint16_t*>(sh.cnt);
        const uint32_t tb2 = 2u * (uint32_t)(tid + (tid >> 5));
#pragma unroll 1
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = pass * 4;
            if (((varying >> shift) & 15u) == 0u) continue;             // the pass would be the identity
#pragma unroll
            for (int w = 0; w < 8; ++w) sh.cnt[w * 1056 + (tb2 >> 1)] = 0u;
            uint32_t key[MAXS];
            uint32_t rl[MAXS];                 // rank among the thread's equal digits | label << 8 | counter index << 16
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                const bool valid = j < S && kbase + j < n;
                key[j] = valid ? sh.keys[kbase + j] : 0xffffffffu;
                rl[j] = valid ? ((uint32_t)sh.labs[kbase + j] << 8) : 0u;
            }
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                if (j < S && kbase + j < n) {
                    const uint32_t d = (key[j] >> shift) & 15u;
                    const uint32_t ci = (d & 7u) * 2112u + (d >> 3) + tb2;
                    const uint32_t old = c16[ci];
                    c16[ci] = (uint16_t)(old + 1u);
                    rl[j] |= old | (ci << 16);
                }
            }
            __syncthreads();                                            // every key is in registers, every count is final
            // raking scan: thread r owns the 8 consecutive entries [8 r, 8 r + 8) of the (digit word, thread) order
            uint32_t c8[8], tot = 0;
            {
                const int lin0 = tid * 8, p0 = lin0 + (lin0 >> 5);     // 8 consecutive entries never straddle a padding slot
#pragma unroll
                for (int i = 0; i < 8; ++i) { c8[i] = sh.cnt[p0 + i]; tot += c8[i]; }
                uint32_t block_tot;
                uint32_t run = block_excl_scan_1024(tot, sh.scan_tmp, &block_tot);     // packed halves: no carry (<= 16 384)
                const uint32_t low_total = block_tot & 0xffffu;         // keys whose digit is 0..7 precede every digit 8..15
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sh.cnt[p0 + i] = run + (low_total << 16);
                    run += c8[i];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                if (j < S && kbase + j < n) {
                    const uint32_t pos = (uint32_t)c16[rl[j] >> 16] + (rl[j] & 0xffu);
                    sh.keys[pos] = key[j];
                    sh.labs[pos] = (uint8_t)((rl[j] >> 8) & 1u);
                }
            }
            __syncthreads();
        }
    }

    const long long clk1 = clock64();
    const int nv = (int)sh.n_valid;
    const int npos = (int)sh.n_pos;
    // ---- 3 distinct thresholds (blocked arrangement: thread t owns rows [t * per, (t + 1) * per), per odd and <= MAXS).
    // One pass over shared memory: flags and label bits are kept as bit masks, (flag count, label count) scanned packed.
    int m = 0;
    {
        const int per = ((nv + kSmallThreads - 1) / kSmallThreads) | 1;     // odd: the strided walks are bank-conflict free
        const int lo = min(nv, tid * per);
        uint32_t fmask = 0, lmask = 0;
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            const int i = lo + j;
            if (j < per && i < nv) {
                if ((i == nv - 1) || (sh.keys[i] != sh.keys[i + 1])) fmask |= 1u << j;
                if (sh.labs[i]) lmask |= 1u << j;
            }
        }
        uint32_t tot;
        const uint32_t ex = block_excl_scan_1024(((uint32_t)__popc(fmask) << 16) | (uint32_t)__popc(lmask), sh.scan_tmp, &tot);
        m = (int)(tot >> 16);
        uint32_t slot = ex >> 16, tps = ex & 0xffffu;
        const bool want_keys = thr_out || pthr_out;
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            tps += (lmask >> j) & 1u;
            if ((fmask >> j) & 1u) {
                const int i = lo + j;
                sh.d_tps[slot] = (uint16_t)tps;
                sh.d_fps[slot] = (uint16_t)(1u + (uint32_t)i - tps);
                if (want_keys) d_key[slot] = sh.keys[i];
                ++slot;
            }
        }
    }
    __syncthreads();
    // ---- 4 corners (roc_curve drop_intermediate=True), origin prepended; kept points -> sh.keys as (tps << 16 | fps)
    int kept = 0;
    {
        const int per = ((m + kSmallThreads - 1) / kSmallThreads) | 1;
        const int lo = min(m, tid * per);
        uint32_t kmask = 0;
        uint32_t pk[MAXS];
        if (lo < m) {
            // sliding window over (fps, tps) of points lo-1 .. lo+per: one shared-memory read per point
            int f0 = lo > 0 ? sh.d_fps[lo - 1] : 0, t0 = lo > 0 ? sh.d_tps[lo - 1] : 0;
            int f1 = sh.d_fps[lo], t1 = sh.d_tps[lo];
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                const int i = lo + j;
                if (j < per && i < m) {
                    const bool has_next = i + 1 < m;
                    const int f2 = has_next ? sh.d_fps[i + 1] : 0, t2 = has_next ? sh.d_tps[i + 1] : 0;
                    const bool kp = (m <= 2 || i == 0 || i == m - 1) || (f0 - 2 * f1 + f2 != 0) || (t0 - 2 * t1 + t2 != 0);
                    if (kp) kmask |= 1u << j;
                    pk[j] = ((uint32_t)t1 << 16) | (uint32_t)f1;
                    f0 = f1; t0 = t1; f1 = f2; t1 = t2;
                }
            }
        }
        uint32_t tot;
        uint32_t slot = block_excl_scan_1024((uint32_t)__popc(kmask), sh.scan_tmp, &tot) + 1;   // +1: the prepended origin
        kept = (int)tot;
        if (tid == 0) {
            sh.keys[0] = 0u;
            if (thr_out) thr_out[0] = INFINITY;
        }
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            if ((kmask >> j) & 1u) {
                sh.keys[slot] = pk[j];
                if (thr_out) thr_out[slot] = key_to_score(d_key[lo + j]);
                ++slot;
            }
        }
    }
    __syncthreads();
    const long long clk2 = clock64();
    // ---- 5 terms (same operation sequence as auc_terms_kernel): a warp takes 31 consecutive terms = 32 points, one
    // pair of fp64 divisions per point, the right neighbour's (fpr, tpr) by shuffle
    {
        const int P = kept + 1;
        const double ftot = (double)(nv - npos), ttot = (double)npos;
        for (int base = warp * 31; base < P; base += 32 * 31) {
            const int i = base + lane;
            double f0 = 0.0, t0 = 0.0;
            if (i < P) {
                const uint32_t pk = sh.keys[i];
                f0 = __ddiv_rn((double)(pk & 0xffffu), ftot);
                t0 = __ddiv_rn((double)(pk >> 16), ttot);
                // lane 31's point is lane 0's point of the next chunk and is written there
                if (fpr_out && lane < 31) { fpr_out[i] = f0; tpr_out[i] = t0; }
            }
            const double f1 = __shfl_down_sync(kFullMask, f0, 1), t1 = __shfl_down_sync(kFullMask, t0, 1);
            if (lane < 31 && i + 1 < P)
                terms[i] = __dmul_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 0.5);   // x / 2.0 == x * 0.5, bit for bit
        }
    }
    __syncthreads();
    // ---- 6/7 numpy pairwise sum
    const bool single = (npos == 0) || (npos == nv);
    const uint32_t status = sh.status | (single ? (uint32_t)EOE_AUC_STATUS_SINGLE_CLASS : 0u);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const bool undefined = single || (status & ~(uint32_t)EOE_AUC_STATUS_SINGLE_CLASS);
    {
        const double v = block_pairwise_sum(terms, kept, sh.nodes);
        if (tid == 0) auc_out[0] = undefined ? qnan : v;
    }
    // ---- PRC / average precision (same operation sequence as auc_prc_terms_kernel)
    if (flags & EOE_AUC_WITH_PRC) {
        const double ttot = (double)npos;
        auto prec = [&](int j) {
            const double tp = (double)sh.d_tps[j], ps = __dadd_rn(tp, (double)sh.d_fps[j]);
            return ps != 0.0 ? __ddiv_rn(tp, ps) : 0.0;
        };
        auto rec = [&](int j) { return ttot == 0.0 ? 1.0 : __ddiv_rn((double)sh.d_tps[j], ttot); };
        for (int i = tid; i < m; i += kSmallThreads) {
            const int j = m - 1 - i;
            const double p = prec(j), r = rec(j);
            const double r_next = (i + 1 < m) ? rec(j - 1) : 0.0;
            terms[i] = __dmul_rn(__dsub_rn(r_next, r), p);
            if (prec_out) { prec_out[i] = p; rec_out[i] = r; }
            if (pthr_out) pthr_out[i] = key_to_score(d_key[j]);
        }
        if (prec_out && tid == 0) { prec_out[m] = 1.0; rec_out[m] = 0.0; }
        __syncthreads();
        const double v = block_pairwise_sum(terms, m, sh.nodes);
        if (tid == 0) auc_out[1] = undefined ? qnan : fmax(0.0, -v);
    }
    if (info_out && tid == 0) {
        info_out[0] = nv; info_out[1] = npos; info_out[2] = m; info_out[3] = kept + 1; info_out[4] = (int64_t)status;
        info_out[5] = clk1 - clk0; info_out[6] = clk2 - clk1; info_out[7] = clock64() - clk2;
    }
}

template <typename T, int MAXS>
static int auc_launch_small(const void* scores, const int64_t* labels, int64_t n, int flags, char* ws, const AucLayout& L,
                            double* auc_out, int64_t* info_out, double* fpr_out, double* tpr_out, float* thr_out,
                            double* prec_out, double* rec_out, float* pthr_out, cudaStream_t st) {
    static bool attr_set = false;                      // per process and instantiation; the attribute is per function
    auto kern = auc_small_kernel<T, MAXS>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
        if (e != cudaSuccess) { set_cuda_error(e, "auc small attribute"); return EOE_ERR_CUDA; }
        attr_set = true;
    }
    kern<<<1, kSmallThreads, sizeof(SmallShared), st>>>(
        (const T*)scores, labels, (int)n, flags, (uint32_t*)(ws + L.keys_b), (double*)(ws + L.terms), auc_out,
        info_out, fpr_out, tpr_out, thr_out, prec_out, rec_out, pthr_out);
    return check_launch("auc (single launch)", 1);
}

template <typename T>
static int auc_run_small(const void* scores, const int64_t* labels, int64_t n, int flags, char* ws, const AucLayout& L,
                         double* auc_out, int64_t* info_out, double* fpr_out, double* tpr_out, float* thr_out,
                         double* prec_out, double* rec_out, float* pthr_out, cudaStream_t st) {
    const int S = (int)((n + kSmallThreads - 1) / kSmallThreads) | 1;      // keys per thread, odd
#define EOE_AUC_SMALL(M) return auc_launch_small<T, M>(scores, labels, n, flags, ws, L, auc_out, info_out, fpr_out, tpr_out, \
                                                      thr_out, prec_out, rec_out, pthr_out, st)
    if (S <= 5) EOE_AUC_SMALL(5);
    if (S <= 9) EOE_AUC_SMALL(9);
    if (S <= 13) EOE_AUC_SMALL(13);
    EOE_AUC_SMALL(17);
#undef EOE_AUC_SMALL
}

// ------------------------------------------------------------------------------------------ cluster path
// The one-CTA kernel above is bound by ONE SM's shared-memory port (7 radix passes x 7 accesses per key).  Here a thread
// block CLUSTER of 8 CTAs (8 SMs, one launch) shares the work: CTA c owns slots [c P, (c + 1) P) of the global order in
// its shared memory, ranks its keys exactly as the one-CTA kernel does, and the passes are stitched together through
// distributed shared memory -- per pass every CTA publishes its 16 digit totals, reads the other seven's, and scatters
// its keys straight into the owning CTA's shared memory (st.shared::cluster).  Two cluster barriers per pass (~380
// clocks each).  The scans, terms and pairwise leaves are split the same way through the global workspace; CTA 0 finishes
// the tree.  Same arithmetic, same order: bit-identical to both other paths.  Up to 8 x 16 384 scores.
constexpr int kClusterCtas = 8;
constexpr int kClusterMax = kClusterCtas * kSmallMax;
// measured cross-overs (tools/microbench_latency.py, profiles/r2_latency_reference_sizes.jsonl; device time per call):
// one CTA wins up to ~12 k scores (20 us at 3 000, 42 us at 10 000 -- the cluster's DSMEM scatter and its ~20 cluster
// barriers cost what its 8 SMs save), the cluster between 12 k and 48 k (50 vs 61 us at 16 384; 73 us at 32 768 against
// 78 us device / 115 us per call for the then 11 launches of the tiled pipeline), the tiled pipeline above
constexpr int kSingleCtaBelow = 12288;
constexpr int kClusterUseMax = 49152;

struct ClusterShared {
    uint32_t keys[kSmallMax + 8];
    uint32_t cnt[8 * kSmallThreads + 8 * kSmallThreads / 32];
    uint8_t labs[kSmallMax];
    uint32_t scan_tmp[32];
    uint32_t dtot[16];                 // this CTA's digit totals of the running pass (read by the other CTAs)
    int32_t adj[16];                   // global position = local exclusive prefix + rank + adj[digit]
    uint32_t pub[4];                   // per-phase totals published to the other CTAs
    uint32_t and_all, or_all, n_valid, n_pos, status;      // CTA 0's copies accumulate the whole cluster's
    uint32_t g_varying, g_nv, g_npos, g_status;            // every CTA's copy of the cluster-wide values
};

template <typename T, int MAXS>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kSmallThreads, 1)
auc_cluster_kernel(const T* __restrict__ scores, const int64_t* __restrict__ labels, int n, int flags,
                   uint32_t* d_tps, uint32_t* d_fps, uint32_t* d_key, uint32_t* k_tps, uint32_t* k_fps, double* terms,
                   double* nodes, double* nodes2, double* __restrict__ auc_out, int64_t* __restrict__ info_out,
                   double* __restrict__ fpr_out, double* __restrict__ tpr_out, float* __restrict__ thr_out,
                   double* __restrict__ prec_out, double* __restrict__ rec_out, float* __restrict__ pthr_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ClusterShared& sh = *reinterpret_cast<ClusterShared*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = (n + kClusterCtas - 1) / kClusterCtas;           // slots per CTA
    const unsigned long long invP = (((unsigned long long)1 << 32) + (unsigned)P - 1) / (unsigned)P;    // pos / P = (pos * invP) >> 32
    const int lo_g = c * P;
    const int nloc = max(0, min(n - lo_g, P));
    const int items = (P + kSmallThreads - 1) / kSmallThreads;     // uniform over the cluster
    const int S = items | 1;
    const int kbase = tid * S;
    ClusterShared* sh0 = cluster.map_shared_rank(&sh, 0);
    if (tid == 0) { sh.and_all = 0xffffffffu; sh.or_all = 0u; sh.n_valid = 0; sh.n_pos = 0; sh.status = 0; }
    cluster.sync();
    const long long clk0 = clock64();

    // ---- 1 keys
    {
        const bool ignore_neg = flags & EOE_AUC_IGNORE_NEGATIVE_LABELS;
        const int wbase = warp * items * 32;
        uint32_t nv = 0, np = 0, bad = 0, a_and = 0xffffffffu, a_or = 0u;
        constexpr int NI = MAXS < kSmallItems ? MAXS : kSmallItems;
        float f[NI];
        int64_t l[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int idx = wbase + j * 32 + lane;
            const bool in = j < items && idx < nloc;
            f[j] = in ? to_f32<T>(scores[lo_g + idx]) : 0.f;
            l[j] = in ? labels[lo_g + idx] : 0;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int idx = wbase + j * 32 + lane;
            if (j < items && idx < nloc) {
                uint32_t key = 0xffffffffu;
                uint8_t lb = 0;
                if (!(ignore_neg && l[j] < 0)) {
                    if (!isfinite(f[j])) bad = 1;
                    key = desc_key(f[j]);
                    lb = (l[j] == 1);
                    nv++;
                    np += lb;
                }
                sh.keys[idx] = key;
                sh.labs[idx] = lb;
                a_and &= key;
                a_or |= key;
            }
        }
        nv = __reduce_add_sync(kFullMask, nv);
        np = __reduce_add_sync(kFullMask, np);
        bad = __reduce_or_sync(kFullMask, bad);
        a_and = __reduce_and_sync(kFullMask, a_and);
        a_or = __reduce_or_sync(kFullMask, a_or);
        if (lane == 0) {                                   // straight into CTA 0's accumulators
            atomicAdd(&sh0->n_valid, nv);
            atomicAdd(&sh0->n_pos, np);
            if (bad) atomicOr(&sh0->status, (uint32_t)EOE_AUC_STATUS_NONFINITE);
            atomicAnd(&sh0->and_all, a_and);
            atomicOr(&sh0->or_all, a_or);
        }
    }
    cluster.sync();
    if (tid == 0) {
        sh.g_varying = sh0->and_all ^ sh0->or_all;
        sh.g_nv = sh0->n_valid; sh.g_npos = sh0->n_pos; sh.g_status = sh0->status;
    }
    __syncthreads();
    const uint32_t varying = sh.g_varying;
    const int nv = (int)sh.g_nv, npos = (int)sh.g_npos;

    // ---- 2 sort (see auc_small_kernel; the digit totals of the 8 CTAs are combined through distributed shared memory)
    {
        uint16_t* c16 = reinterpret_cast<uint16_t*>(sh.cnt);
        const uint32_t tb2 = 2u * (uint32_t)(tid + (tid >> 5));
#pragma unroll 1
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = pass * 4;
            if (((varying >> shift) & 15u) == 0u) continue;             // cluster-uniform
#pragma unroll
            for (int w = 0; w < 8; ++w) sh.cnt[w * 1056 + (tb2 >> 1)] = 0u;
            uint32_t key[MAXS];
            uint32_t rl[MAXS];
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                const bool valid = j < S && kbase + j < nloc;
                key[j] = valid ? sh.keys[kbase + j] : 0xffffffffu;
                rl[j] = valid ? ((uint32_t)sh.labs[kbase + j] << 8) : 0u;
            }
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                if (j < S && kbase + j < nloc) {
                    const uint32_t d = (key[j] >> shift) & 15u;
                    const uint32_t ci = (d & 7u) * 2112u + (d >> 3) + tb2;
                    const uint32_t old = c16[ci];
                    c16[ci] = (uint16_t)(old + 1u);
                    rl[j] |= old | (ci << 16);
                }
            }
            __syncthreads();
            uint32_t c8[8], tot = 0, block_tot;
            {
                const int lin0 = tid * 8, p0 = lin0 + (lin0 >> 5);
#pragma unroll
                for (int i = 0; i < 8; ++i) { c8[i] = sh.cnt[p0 + i]; tot += c8[i]; }
                uint32_t run = block_excl_scan_1024(tot, sh.scan_tmp, &block_tot);
                const uint32_t low_total = block_tot & 0xffffu;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sh.cnt[p0 + i] = run + (low_total << 16);
                    run += c8[i];
                }
                block_tot += low_total << 16;              // "row 8" of the high halves, with the same offset
            }
            __syncthreads();
            // digit d's total and first local position in this CTA, from the prefixes of thread 0 in rows w and w + 1
            uint32_t lstart = 0;
            if (tid < 16) {
                const int w = tid & 7, hs = (tid >> 3) * 16;
                const uint32_t e0 = (sh.cnt[w * 1056] >> hs) & 0xffffu;
                const uint32_t e1 = ((w == 7 ? block_tot : sh.cnt[(w + 1) * 1056]) >> hs) & 0xffffu;
                sh.dtot[tid] = e1 - e0;
                lstart = e0;
            }
            cluster.sync();                                // every CTA's keys are in registers, every dtot is published
            if (tid < 32) {
                uint32_t before = 0, rowsum = 0;           // keys with my digit in CTAs < c / in all CTAs
                if (tid < 16) {
#pragma unroll
                    for (int cc = 0; cc < kClusterCtas; ++cc) {
                        const uint32_t v = cluster.map_shared_rank(&sh, cc)->dtot[tid];
                        rowsum += v;
                        if (cc < c) before += v;
                    }
                }
                uint32_t inc = rowsum;                     // exclusive prefix over digits: keys with a smaller digit
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFullMask, inc, o);
                    if (lane >= o) inc += t;
                }
                if (tid < 16) sh.adj[tid] = (int32_t)(inc - rowsum + before) - (int32_t)lstart;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                if (j < S && kbase + j < nloc) {
                    const uint32_t d = (key[j] >> shift) & 15u;
                    const uint32_t pos = (uint32_t)((int32_t)((uint32_t)c16[rl[j] >> 16] + (rl[j] & 0xffu)) + sh.adj[d]);
                    const uint32_t dest = (uint32_t)(((unsigned long long)pos * invP) >> 32);     // exact for pos * P < 2^32
                    const uint32_t slot = pos - dest * (uint32_t)P;
                    ClusterShared* r = cluster.map_shared_rank(&sh, dest);
                    r->keys[slot] = key[j];
                    r->labs[slot] = (uint8_t)((rl[j] >> 8) & 1u);
                }
            }
            cluster.sync();                                // every slice holds the keys of this pass's order
        }
    }
    const long long clk1 = clock64();

    // ---- 3 distinct thresholds -> global d_tps / d_fps / d_key
    const int nvl = max(0, min(nv - lo_g, nloc));                  // kept rows of this slice
    int m = 0;
    {
        uint32_t fmask = 0, lmask = 0;
        const uint32_t next_first = (c + 1 < kClusterCtas) ? cluster.map_shared_rank(&sh, c + 1)->keys[0] : 0u;
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            const int i = kbase + j;
            if (j < S && i < nvl) {
                const int gi = lo_g + i;
                const uint32_t nxt = (i + 1 < P) ? sh.keys[i + 1] : next_first;
                if (gi == nv - 1 || sh.keys[i] != nxt) fmask |= 1u << j;
                if (sh.labs[i]) lmask |= 1u << j;
            }
        }
        uint32_t tf, tl;
        uint32_t slot = block_excl_scan_1024((uint32_t)__popc(fmask), sh.scan_tmp, &tf);
        uint32_t tps = block_excl_scan_1024((uint32_t)__popc(lmask), sh.scan_tmp, &tl);
        if (tid == 0) { sh.pub[0] = tf; sh.pub[1] = tl; }
        cluster.sync();
        for (int cc = 0; cc < kClusterCtas; ++cc) {
            const ClusterShared* r = cluster.map_shared_rank(&sh, cc);
            const uint32_t vf = r->pub[0], vl = r->pub[1];
            m += (int)vf;
            if (cc < c) { slot += vf; tps += vl; }
        }
        const bool want_keys = thr_out || pthr_out;
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            tps += (lmask >> j) & 1u;
            if ((fmask >> j) & 1u) {
                const int gi = lo_g + kbase + j;
                d_tps[slot] = tps;
                d_fps[slot] = 1u + (uint32_t)gi - tps;
                if (want_keys) d_key[slot] = sh.keys[kbase + j];
                ++slot;
            }
        }
    }
    cluster.sync();
    // ---- 4 corners -> global k_tps / k_fps (+ thresholds); CTA c takes points [c Q, (c + 1) Q)
    int kept = 0;
    {
        const int Q = (m + kClusterCtas - 1) / kClusterCtas;
        const int q_lo = min(m, c * Q), q_hi = min(m, q_lo + Q);
        const int per = ((Q + kSmallThreads - 1) / kSmallThreads) | 1;      // <= S
        const int lo = min(q_hi, q_lo + tid * per);
        uint32_t kmask = 0;
        uint32_t pt[MAXS], pf[MAXS];
        if (lo < q_hi) {
            int f0 = lo > 0 ? (int)__ldcg(d_fps + lo - 1) : 0, t0 = lo > 0 ? (int)__ldcg(d_tps + lo - 1) : 0;
            int f1 = (int)__ldcg(d_fps + lo), t1 = (int)__ldcg(d_tps + lo);
#pragma unroll
            for (int j = 0; j < MAXS; ++j) {
                const int i = lo + j;
                if (j < per && i < q_hi) {
                    const bool has_next = i + 1 < m;
                    const int f2 = has_next ? (int)__ldcg(d_fps + i + 1) : 0, t2 = has_next ? (int)__ldcg(d_tps + i + 1) : 0;
                    const bool kp = (m <= 2 || i == 0 || i == m - 1) || (f0 - 2 * f1 + f2 != 0) || (t0 - 2 * t1 + t2 != 0);
                    if (kp) kmask |= 1u << j;
                    pt[j] = (uint32_t)t1; pf[j] = (uint32_t)f1;
                    f0 = f1; t0 = t1; f1 = f2; t1 = t2;
                }
            }
        }
        uint32_t tk;
        uint32_t slot = block_excl_scan_1024((uint32_t)__popc(kmask), sh.scan_tmp, &tk) + 1;     // +1: the prepended origin
        if (tid == 0) sh.pub[2] = tk;
        cluster.sync();
        for (int cc = 0; cc < kClusterCtas; ++cc) {
            const uint32_t v = cluster.map_shared_rank(&sh, cc)->pub[2];
            kept += (int)v;
            if (cc < c) slot += v;
        }
        if (c == 0 && tid == 0) {
            k_tps[0] = 0; k_fps[0] = 0;
            if (thr_out) thr_out[0] = INFINITY;
        }
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
            if ((kmask >> j) & 1u) {
                k_tps[slot] = pt[j];
                k_fps[slot] = pf[j];
                if (thr_out) thr_out[slot] = key_to_score(__ldcg(d_key + lo + j));
                ++slot;
            }
        }
    }
    cluster.sync();
    const long long clk2 = clock64();
    // ---- 5 terms: warp chunks of 31 terms, dealt round-robin over the cluster's 256 warps
    {
        const int Ppts = kept + 1;
        const double ftot = (double)(nv - npos), ttot = (double)npos;
        for (int base = (c * 32 + warp) * 31; base < Ppts; base += kClusterCtas * 32 * 31) {
            const int i = base + lane;
            double f0 = 0.0, t0 = 0.0;
            if (i < Ppts) {
                f0 = __ddiv_rn((double)__ldcg(k_fps + i), ftot);
                t0 = __ddiv_rn((double)__ldcg(k_tps + i), ttot);
                if (fpr_out && lane < 31) { fpr_out[i] = f0; tpr_out[i] = t0; }
            }
            const double f1 = __shfl_down_sync(kFullMask, f0, 1), t1 = __shfl_down_sync(kFullMask, t0, 1);
            if (lane < 31 && i + 1 < Ppts)
                terms[i] = __dmul_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 0.5);
        }
    }
    cluster.sync();
    // ---- 6 leaves of numpy's pairwise tree over the cluster, 7 the tree itself by CTA 0
    const bool single = (npos == 0) || (npos == nv);
    const uint32_t status = sh.g_status | (single ? (uint32_t)EOE_AUC_STATUS_SINGLE_CLASS : 0u);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const bool undefined = single || (status & ~(uint32_t)EOE_AUC_STATUS_SINGLE_CLASS);
    auto leaves = [&](const double* a, int64_t Tn, double* nd) {
        if (Tn <= 0) return;
        const int D = dev_pairwise_depth(Tn);
        for (int64_t g0 = (int64_t)c * (kSmallThreads / 8); g0 < ((int64_t)1 << D); g0 += (int64_t)kClusterCtas * (kSmallThreads / 8))
            pairwise_leaf_slot(a, Tn, D, g0 + (tid >> 3), tid & 7, 0xffu << (tid & 24), nd);
    };
    auto tree = [&](int64_t Tn, double* nd) {              // CTA 0 only (block-uniform call)
        if (Tn <= 0) return 0.0;
        pairwise_tree_levels<kSmallThreads>(Tn, dev_pairwise_depth(Tn), nd);
        return __ldcg(nd + 1);
    };
    leaves(terms, kept, nodes);
    cluster.sync();
    if (c == 0) {
        const double v = tree(kept, nodes);
        if (tid == 0) auc_out[0] = undefined ? qnan : v;
    }
    if (flags & EOE_AUC_WITH_PRC) {
        const double ttot = (double)npos;
        auto prec = [&](int j) {
            const double tp = (double)__ldcg(d_tps + j), ps = __dadd_rn(tp, (double)__ldcg(d_fps + j));
            return ps != 0.0 ? __ddiv_rn(tp, ps) : 0.0;
        };
        auto rec = [&](int j) { return ttot == 0.0 ? 1.0 : __ddiv_rn((double)__ldcg(d_tps + j), ttot); };
        for (int i = c * kSmallThreads + tid; i < m; i += kClusterCtas * kSmallThreads) {
            const int j = m - 1 - i;
            const double p = prec(j), r = rec(j);
            const double r_next = (i + 1 < m) ? rec(j - 1) : 0.0;
            terms[i] = __dmul_rn(__dsub_rn(r_next, r), p);
            if (prec_out) { prec_out[i] = p; rec_out[i] = r; }
            if (pthr_out) pthr_out[i] = key_to_score(__ldcg(d_key + j));
        }
        if (prec_out && c == 0 && tid == 0) { prec_out[m] = 1.0; rec_out[m] = 0.0; }
        cluster.sync();
        leaves(terms, m, nodes2);
        cluster.sync();
        if (c == 0) {
            const double v = tree(m, nodes2);
            if (tid == 0) auc_out[1] = undefined ? qnan : fmax(0.0, -v);
        }
    }
    if (info_out && c == 0 && tid == 0) {
        info_out[0] = nv; info_out[1] = npos; info_out[2] = m; info_out[3] = kept + 1; info_out[4] = (int64_t)status;
        info_out[5] = clk1 - clk0; info_out[6] = clk2 - clk1; info_out[7] = clock64() - clk2;
    }
    cluster.sync();                                        // no CTA exits while a peer may still read its shared memory
}

template <typename T, int MAXS>
static int auc_launch_cluster(const void* scores, const int64_t* labels, int64_t n, int flags, char* ws, const AucLayout& L,
                              double* auc_out, int64_t* info_out, double* fpr_out, double* tpr_out, float* thr_out,
                              double* prec_out, double* rec_out, float* pthr_out, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = auc_cluster_kernel<T, MAXS>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ClusterShared));
        if (e != cudaSuccess) { set_cuda_error(e, "auc cluster attribute"); return EOE_ERR_CUDA; }
        attr_set = true;
    }
    // PRC tree nodes live in the k_tps / k_fps region (dead once the ROC terms exist): (n + 2) * 8 bytes >= (2 << depth) * 8
    kern<<<kClusterCtas, kSmallThreads, sizeof(ClusterShared), st>>>(
        (const T*)scores, labels, (int)n, flags, (uint32_t*)(ws + L.d_tps), (uint32_t*)(ws + L.d_fps), (uint32_t*)(ws + L.keys_b),
        (uint32_t*)(ws + L.k_tps), (uint32_t*)(ws + L.k_fps), (double*)(ws + L.terms), (double*)(ws + L.nodes),
        (double*)(ws + L.k_tps), auc_out, info_out, fpr_out, tpr_out, thr_out, prec_out, rec_out, pthr_out);
    return check_launch("auc (one cluster launch)", 1);
}

template <typename T>
static int auc_run_cluster(const void* scores, const int64_t* labels, int64_t n, int flags, char* ws, const AucLayout& L,
                           double* auc_out, int64_t* info_out, double* fpr_out, double* tpr_out, float* thr_out,
                           double* prec_out, double* rec_out, float* pthr_out, cudaStream_t st) {
    const int P = (int)((n + kClusterCtas - 1) / kClusterCtas);
    const int S = ((P + kSmallThreads - 1) / kSmallThreads) | 1;
#define EOE_AUC_CLUSTER(M) return auc_launch_cluster<T, M>(scores, labels, n, flags, ws, L, auc_out, info_out, fpr_out, tpr_out, \
                                                          thr_out, prec_out, rec_out, pthr_out, st)
    if (S <= 3) EOE_AUC_CLUSTER(3);
    if (S <= 5) EOE_AUC_CLUSTER(5);
    if (S <= 9) EOE_AUC_CLUSTER(9);
    EOE_AUC_CLUSTER(17);
#undef EOE_AUC_CLUSTER
}

template <typename T>
static int auc_run(const void* scores, const int64_t* labels, int64_t n, int flags, char* ws, const AucLayout& L,
                   double* auc_out, int64_t* info_out, double* fpr_out, double* tpr_out, float* thr_out,
                   double* prec_out, double* rec_out, float* pthr_out, cudaStream_t st) {
    AucControl* c = (AucControl*)(ws + L.control);
    uint32_t* keys_a = (uint32_t*)(ws + L.keys_a);
    uint32_t* keys_b = (uint32_t*)(ws + L.keys_b);
    uint8_t* labs_a = (uint8_t*)(ws + L.labs_a);
    uint8_t* labs_b = (uint8_t*)(ws + L.labs_b);
    uint32_t* sort_status = (uint32_t*)(ws + L.sort_status);
    cudaError_t e = cudaMemsetAsync(ws, 0, L.control_bytes, st);
    if (e != cudaSuccess) { set_cuda_error(e, "auc memset"); return EOE_ERR_CUDA; }
#if EOE_AUC_KEYS_V2
    int kgrid = (int)((n + kKeysThreads * kKeysBatch - 1) / (kKeysThreads * kKeysBatch));
    if (kgrid > kNumSMs) kgrid = kNumSMs;
#else
    int kgrid = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (kgrid > kNumSMs * 8) kgrid = kNumSMs * 8;
#endif
    auc_keys_kernel<T><<<kgrid, kKeysThreads, 0, st>>>((const T*)scores, labels, n, flags, keys_a, labs_a, c);
    cudaError_t launch_err = cudaSuccess;
#define AUC_LAUNCH(...) do { if (launch_err == cudaSuccess) launch_err = launch_dependent(__VA_ARGS__); } while (0)
    for (int pass = 0; pass < 4; ++pass) {
        const bool fwd = (pass & 1) == 0;
        auto sort_pass = sort_items_for(n) == 8 ? auc_sort_pass_kernel<8> : auc_sort_pass_kernel<16>;
        AUC_LAUNCH(sort_pass, L.sort_tiles, kSortThreads, 0, st, fwd ? keys_a : keys_b, fwd ? labs_a : labs_b,
                   fwd ? keys_b : keys_a, fwd ? labs_b : labs_a, n, pass, c, sort_status + (size_t)pass * L.sort_tiles * 256);
    }
    // sorted data is back in (keys_a, labs_a); keys_b is free and receives the distinct thresholds
    uint32_t* d_tps = (uint32_t*)(ws + L.d_tps);
    uint32_t* d_fps = (uint32_t*)(ws + L.d_fps);
    uint32_t* k_tps = (uint32_t*)(ws + L.k_tps);
    uint32_t* k_fps = (uint32_t*)(ws + L.k_fps);
    double* terms = (double*)(ws + L.terms);
    double* nodes = (double*)(ws + L.nodes);
    AUC_LAUNCH(auc_distinct_kernel, L.scan_tiles, kScanThreads, 0, st, keys_a, labs_a, c,
               (unsigned long long*)(ws + L.scan1_status), d_tps, d_fps, keys_b);
    AUC_LAUNCH(auc_corner_kernel, L.scan_tiles, kScanThreads, 0, st, d_tps, d_fps, keys_b, c,
               (unsigned long long*)(ws + L.scan2_status), k_tps, k_fps, thr_out);
    int tgrid = (int)((n + 1 + 255) / 256);
    if (tgrid > kNumSMs * 4) tgrid = kNumSMs * 4;
    AUC_LAUNCH(auc_terms_kernel, tgrid, 256, 0, st, k_tps, k_fps, c, terms, fpr_out, tpr_out);
    const int lgrid = (int)((((int64_t)8 << L.max_depth) + 255) / 256);
    const int tdepth = L.max_depth < kTreeSmemDepth ? L.max_depth : kTreeSmemDepth;
    const size_t tree_smem = ((size_t)2 << tdepth) * sizeof(double) + ((size_t)1 << tdepth) * sizeof(uint32_t);
    if (tree_smem > 48 * 1024) {                 // per device and idempotent; only past ~0.2 M scores
        e = cudaFuncSetAttribute(pairwise_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(((size_t)2 << kTreeSmemDepth) * sizeof(double) + ((size_t)1 << kTreeSmemDepth) * sizeof(uint32_t)));
        if (e != cudaSuccess) { set_cuda_error(e, "pairwise_tree smem attr"); return EOE_ERR_CUDA; }
    }
    const unsigned long long* n_kept = &c->n_kept;
    const unsigned long long* n_distinct = &c->n_distinct;
    AUC_LAUNCH(pairwise_leaves_kernel, lgrid, 256, 0, st, terms, n_kept, (int64_t)0, nodes);
    const bool prc = (flags & EOE_AUC_WITH_PRC) != 0;
    AUC_LAUNCH(pairwise_tree_kernel, 1, 1024, tree_smem, st, n_kept, (int64_t)0, nodes, c, auc_out, 0,
               prc ? (int64_t*)nullptr : info_out);
    if (prc) {
        AUC_LAUNCH(auc_prc_terms_kernel, tgrid, 256, 0, st, d_tps, d_fps, keys_b, c, terms, prec_out, rec_out, pthr_out);
        AUC_LAUNCH(pairwise_leaves_kernel, lgrid, 256, 0, st, terms, n_distinct, (int64_t)0, nodes);
        AUC_LAUNCH(pairwise_tree_kernel, 1, 1024, tree_smem, st, n_distinct, (int64_t)0, nodes, c, auc_out + 1, 1, info_out);
    }
#undef AUC_LAUNCH
    if (launch_err != cudaSuccess) return EOE_ERR_CUDA;
    return check_launch("auc pipeline", 10 + (prc ? 3 : 0));
}

}  // namespace eoe

using namespace eoe;

extern "C" size_t eoe_auc_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    return auc_layout(n).total;
}

extern "C" int eoe_auc(const void* scores, int score_dtype, const int64_t* labels, int64_t n, int flags,
                       void* workspace, size_t workspace_bytes, double* auc_out, int64_t* info_out,
                       double* fpr_out, double* tpr_out, float* thr_out, double* prec_out, double* rec_out,
                       float* prc_thr_out, void* stream) {
    if (!scores || !labels || !auc_out || n <= 0) return EOE_ERR_ARG;
    if (n >= ((int64_t)1 << 30)) return EOE_ERR_SHAPE;
    if ((fpr_out == nullptr) != (tpr_out == nullptr)) return EOE_ERR_ARG;
    if ((prec_out == nullptr) != (rec_out == nullptr)) return EOE_ERR_ARG;
    const AucLayout L = auc_layout(n);
    if (!workspace || workspace_bytes < L.total) return EOE_ERR_WORKSPACE;
    if ((uintptr_t)workspace % 256 != 0) return EOE_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    // one launch up to 49 152 scores: one CTA for the reference's sizes (<= 12 288), a cluster of 8 CTAs above that; the
    // tiled pipeline beyond (the cluster kernel itself handles up to 131 072: EOE_AUC_FORCE_CLUSTER, used by the tests)
    const bool force_cluster = (flags & EOE_AUC_FORCE_CLUSTER) && n <= kClusterMax;
    const bool tiled = !force_cluster && (n > kClusterUseMax || (flags & EOE_AUC_FORCE_TILED));
    const bool one_cta = !tiled && !force_cluster && n <= kSmallMax && ((flags & EOE_AUC_FORCE_SINGLE_CTA) || n <= kSingleCtaBelow);
    NvtxRange nvtx(tiled ? "eoe:auc (tiled pipeline)" : (one_cta ? "eoe:auc (one CTA)" : "eoe:auc (one cluster)"));
#define EOE_AUC_DISPATCH(T)                                                                                              \
    return tiled ? auc_run<T>(scores, labels, n, flags, ws, L, auc_out, info_out, fpr_out, tpr_out, thr_out, prec_out,    \
                              rec_out, prc_thr_out, st)                                                                   \
                 : (one_cta ? auc_run_small<T>(scores, labels, n, flags, ws, L, auc_out, info_out, fpr_out, tpr_out,      \
                                               thr_out, prec_out, rec_out, prc_thr_out, st)                               \
                            : auc_run_cluster<T>(scores, labels, n, flags, ws, L, auc_out, info_out, fpr_out, tpr_out,    \
                                                 thr_out, prec_out, rec_out, prc_thr_out, st))
    switch (score_dtype) {
        case EOE_F32: EOE_AUC_DISPATCH(float);
        case EOE_F16: EOE_AUC_DISPATCH(__half);
        case EOE_BF16: EOE_AUC_DISPATCH(__nv_bfloat16);
        default: return EOE_ERR_DTYPE;
    }
#undef EOE_AUC_DISPATCH
}
