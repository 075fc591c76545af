// tsnap_transpose_tma_kernel: kModeTransposeTma members (t() / permute() views whose bases and strides are multiples of
// 16 B, 2/4/8-byte elements) — the copy that `.contiguous()` does on the reference's path (T:io_preparers/tensor.py:266-281)
// when a transposed tensor is staged, here written straight into the wire image.
//
//   global --TMA tile load (tensor map of the strided source, one op per 32 KiB tile)--> shared [B][A]
//          --16 B-block transposition in registers: V x LDS.128 + V x STS.128 per V x V block, V = 16 B / element-->
//          shared [A][B] --TMA tile store (tensor map of the destination)--> global
//
// Persistent, one CTA per SM, 4 load stages + 2 store stages of 32 KiB: 128 KiB of loads stay in flight per SM whatever
// the consumers do, nothing is staged in registers across the memory latency, tile edges are clipped by the TMA unit
// (zero fill on load, no write past the extent on store).  The threads only touch shared memory: blocks are assigned to
// lanes along diagonals of the 8 x 8 block grid, so that the 8 lanes of a quarter-warp hit 8 different 16 B bank groups
// both when they read rows of [B][A] and when they write rows of [A][B] — no padding, no swizzle, no bank conflicts.
// HBM-bound: algorithmic traffic = 2 x payload bytes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace tsnap {

constexpr int kTtConsumerWarps = 8;
constexpr int kTtConsumers = kTtConsumerWarps * 32;  // threads that move data
constexpr int kTtThreads = kTtConsumers + 32;        // + the producer warp (one lane drives the TMA unit)
constexpr int kTtIn = 4;   // load stages
constexpr int kTtOut = 2;  // store stages
constexpr uint32_t kTtSmemBytes = (kTtIn + kTtOut) * kTmaTileBytes + 1024 /* alignment slack */ + 512 /* barriers, per-stage tile records */;

// 5-D tile load: global (tensor map) -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const void* tmap, const int32_t (&c)[5], uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_dst), "l"(tmap), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(mbar)
        : "memory");
}
// 5-D tile store: shared -> global (tensor map), tracked by the issuing thread's bulk async-group (SASS: UTMASTG)
__device__ __forceinline__ void tma_store_5d(const void* tmap, const int32_t (&c)[5], uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                 ::"l"(tmap), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_src)
                 : "memory");
}
// the tensor maps were written by a host copy (generic proxy): order them before their first use by the TMA unit
__device__ __forceinline__ void tensormap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
// bounded wait: a tile that never arrives (a rejected tensor map) traps instead of hanging the device
__device__ __forceinline__ void mbar_wait_bounded(uint32_t mbar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(mbar, parity)) {
        if (clock64() - t0 > (4ll << 30)) __trap();  // ~2 s
    }
}

// coordinates of tile `index` in the member's two tensor maps (same enumeration as the LSU transpose: B fastest, then A,
// then the remaining dims from the highest index down)
__device__ __forceinline__ void tt_coords(const Member& m, uint32_t index, int32_t (&cs)[5], int32_t (&cd)[5]) {
    const uint32_t A = m.shift & 255, B = (m.shift >> 8) & 255;
    const uint32_t var = transpose_tma_variant_of(m.shift);
    const uint32_t kA = transpose_tma_side_a(m.unit, var), kB = transpose_tma_side_b(m.unit, var);
    const uint32_t tilesA = (uint32_t)(((uint64_t)m.osize[A] + kA - 1) / kA), tilesB = (uint32_t)(((uint64_t)m.osize[B] + kB - 1) / kB);
    const uint32_t ib = index % tilesB;
    uint32_t rest = index / tilesB;
    const uint32_t ia = rest % tilesA;
    rest /= tilesA;
    cs[0] = cd[1] = (int32_t)(ia * kA);
    cs[1] = cd[0] = (int32_t)(ib * kB);
    cs[2] = cs[3] = cs[4] = cd[2] = cd[3] = cd[4] = 0;
    int k = 2;
    for (int i = (int)m.nouter - 1; i >= 0; --i) {
        if ((uint32_t)i == A || (uint32_t)i == B) continue;
        const uint32_t sz = (uint32_t)m.osize[i];
        cs[k] = cd[k] = (int32_t)(rest % sz);
        rest /= sz;
        ++k;
    }
}

// The V x V element blocks of a tile, V = 16 / ESZ: block (x, y) = 16 B chunk column x of the rows V*y .. V*y+V-1 of
// in[B][A], written as chunk column y of the rows V*x .. V*x+V-1 of out[A][B].  Work unit u -> (x, y): the 8 lanes of a
// quarter-warp share one 8 x 8 sub-grid and one diagonal d, lane q takes (x, y) = (q, q + d mod 8) of it.
template <int ESZ, int VAR>
struct TtGeom {
    static constexpr int V = 16 / ESZ;
    static constexpr int kA0 = ESZ == 2 ? 128 : 64, kB0 = ESZ == 8 ? 64 : 128;  // plan.h transpose_tma_side_a / _b
    static constexpr int kA = VAR == 1 ? kA0 * 2 : VAR == 2 ? kA0 / 2 : kA0, kB = VAR == 1 ? kB0 / 2 : VAR == 2 ? kB0 * 2 : kB0;
    static constexpr int CA = kA / V;         // 16 B chunks per row of in[B][A]
    static constexpr int NY = kB / V;         // 16 B chunks per row of out[A][B] = block rows of in
    static constexpr int NYH = NY / 8;
    static constexpr int UPT = CA * NY / kTtConsumers;  // blocks per thread
    static_assert(CA % 8 == 0 && NY % 8 == 0 && UPT * kTtConsumers == CA * NY && UPT * V == 8, "tile geometry");
    static_assert(kA * kB * ESZ == (int)kTmaTileBytes && kA <= 256 && kB <= 256, "tile payload / TMA box limit");
};

template <int ESZ, int VAR>
__device__ __forceinline__ void tt_unit(uint32_t u, uint32_t* x, uint32_t* y) {
    using G = TtGeom<ESZ, VAR>;
    const uint32_t q = u & 7, g = u >> 3, d = g & 7, sub = g >> 3;
    *x = 8 * (sub / G::NYH) + q;
    *y = 8 * (sub % G::NYH) + ((q + d) & 7);
}

template <int ESZ, int VAR>
__device__ __forceinline__ void tt_read_blocks(const unsigned char* in, uint32_t (&r)[8][4]) {
    using G = TtGeom<ESZ, VAR>;
#pragma unroll
    for (int n = 0; n < G::UPT; ++n) {
        uint32_t x, y;
        tt_unit<ESZ, VAR>(threadIdx.x + n * kTtConsumers, &x, &y);
#pragma unroll
        for (int i = 0; i < G::V; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(in + ((G::V * y + i) * G::CA + x) * 16);
            r[n * G::V + i][0] = v.x;
            r[n * G::V + i][1] = v.y;
            r[n * G::V + i][2] = v.z;
            r[n * G::V + i][3] = v.w;
        }
    }
}

template <int ESZ, int VAR>
__device__ __forceinline__ void tt_write_blocks(unsigned char* out, const uint32_t (&r)[8][4]) {
    using G = TtGeom<ESZ, VAR>;
#pragma unroll
    for (int n = 0; n < G::UPT; ++n) {
        uint32_t x, y;
        tt_unit<ESZ, VAR>(threadIdx.x + n * kTtConsumers, &x, &y);
#pragma unroll
        for (int j = 0; j < G::V; ++j) {
            uint4 v;
            if (ESZ == 4) {  // out row j = element j of each of the 4 in rows
                v = make_uint4(r[n * 4 + 0][j], r[n * 4 + 1][j], r[n * 4 + 2][j], r[n * 4 + 3][j]);
            } else if (ESZ == 8) {  // 2 x 2 blocks of 64-bit elements
                v = make_uint4(r[n * 2 + 0][2 * j], r[n * 2 + 0][2 * j + 1], r[n * 2 + 1][2 * j], r[n * 2 + 1][2 * j + 1]);
            } else {  // 8 x 8 blocks of 16-bit elements: word w of out row j = halves j of in rows 2w, 2w+1
                const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
                v = make_uint4(__byte_perm(r[0][j >> 1], r[1][j >> 1], sel), __byte_perm(r[2][j >> 1], r[3][j >> 1], sel),
                               __byte_perm(r[4][j >> 1], r[5][j >> 1], sel), __byte_perm(r[6][j >> 1], r[7][j >> 1], sel));
            }
            *reinterpret_cast<uint4*>(out + ((G::V * x + j) * G::NY + y) * 16) = v;
        }
    }
}

// tile kind = element size | variant << 4; one switch per phase
#define TT_DISPATCH(kind, FN, ...)                                  \
    switch (kind) {                                                 \
        case 0x04: FN<4, 0>(__VA_ARGS__); break;                    \
        case 0x14: FN<4, 1>(__VA_ARGS__); break;                    \
        case 0x24: FN<4, 2>(__VA_ARGS__); break;                    \
        case 0x02: FN<2, 0>(__VA_ARGS__); break;                    \
        case 0x12: FN<2, 1>(__VA_ARGS__); break;                    \
        case 0x22: FN<2, 2>(__VA_ARGS__); break;                    \
        case 0x08: FN<8, 0>(__VA_ARGS__); break;                    \
        case 0x18: FN<8, 1>(__VA_ARGS__); break;                    \
        default: FN<8, 2>(__VA_ARGS__); break;                      \
    }

// per load stage: what the consumers and the store of the same tile need to know
struct TtMeta {
    uint32_t kind;       // element size | variant << 4
    int32_t cd[5];       // tile coordinates in the destination tensor map
    const void* dmap;
};

__global__ void __launch_bounds__(kTtThreads, 1)
tsnap_transpose_tma_kernel(const Member* __restrict__ members, const Tile* __restrict__ tiles, const TmaPair* __restrict__ maps, uint32_t ntiles) {
    extern __shared__ unsigned char tt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tt_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* in_buf = smem;                                    // [kTtIn][32 KiB]
    unsigned char* out_buf = smem + kTtIn * kTmaTileBytes;           // [kTtOut][32 KiB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (kTtIn + kTtOut) * kTmaTileBytes);
    uint64_t* full = bars;                      // [kTtIn]  TMA load landed              (producer -> consumers)
    uint64_t* empty = full + kTtIn;             // [kTtIn]  every consumer warp has read  (consumers -> producer)
    uint64_t* out_full = empty + kTtIn;         // [kTtOut] every consumer warp has written (consumers -> producer)
    uint64_t* out_free = out_full + kTtOut;     // [kTtOut] the TMA store has read the stage (producer -> consumers)
    TtMeta* meta = reinterpret_cast<TtMeta*>(out_free + kTtOut);  // [kTtIn]

    const uint32_t first = blockIdx.x, step = gridDim.x;
    const uint32_t n_my = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTtIn; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(empty + s), kTtConsumerWarps);
        }
        for (int o = 0; o < kTtOut; ++o) {
            mbar_init(smem_u32(out_full + o), kTtConsumerWarps);
            mbar_init(smem_u32(out_free + o), 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kTtConsumerWarps) {
        // ---- producer: one lane drives the TMA unit; tile decode and descriptor fetches stay off the consumers' path
        if (lane != 0) return;
        uint32_t fenced_member = 0xffffffffu;
        auto issue_load = [&](uint32_t k) {
            const Tile tl = tiles[first + k * step];
            const Member& m = members[tl.member];
            const TmaPair* mp = maps + (uint32_t)m.q_zero_point;
            if (tl.member != fenced_member) {
                tensormap_acquire(mp->src);
                tensormap_acquire(mp->dst);
                fenced_member = tl.member;
            }
            const uint32_t s = k % kTtIn;
            int32_t cs[5];
            tt_coords(m, tl.index, cs, meta[s].cd);
            meta[s].kind = m.unit | (transpose_tma_variant_of(m.shift) << 4);
            meta[s].dmap = mp->dst;
            const uint32_t bar = smem_u32(full + s);
            mbar_expect_tx(bar, kTmaTileBytes);
            tma_load_5d(smem_u32(in_buf + s * kTmaTileBytes), mp->src, cs, bar);
        };
        for (uint32_t k = 0; k < n_my && k < (uint32_t)kTtIn; ++k) issue_load(k);
        for (uint32_t k = 0; k < n_my; ++k) {
            const uint32_t s = k % kTtIn, o = k % kTtOut;
            mbar_wait_bounded(smem_u32(out_full + o), (k / kTtOut) & 1);
            tma_store_5d(meta[s].dmap, meta[s].cd, smem_u32(out_buf + o * kTmaTileBytes));
            bulk_commit();
            bulk_wait_read<kTtOut - 1>();  // the store issued kTtOut - 1 tiles ago has read its stage
            if (k >= (uint32_t)(kTtOut - 1)) mbar_arrive(smem_u32(out_free + (k - (kTtOut - 1)) % kTtOut));
            if (k + kTtIn < n_my) {
                mbar_wait_bounded(smem_u32(empty + s), (k / kTtIn) & 1);
                issue_load(k + kTtIn);
            }
        }
        bulk_wait_all<0>();
        return;
    }

    // ---- consumers: 8 warps, each moving its own blocks; no CTA-wide barrier
    for (uint32_t k = 0; k < n_my; ++k) {
        const uint32_t s = k % kTtIn, o = k % kTtOut;
        mbar_wait_bounded(smem_u32(full + s), (k / kTtIn) & 1);
        const uint32_t kind = meta[s].kind;
        const unsigned char* in = in_buf + s * kTmaTileBytes;
        unsigned char* out = out_buf + o * kTmaTileBytes;
        uint32_t r[8][4];
        TT_DISPATCH(kind, tt_read_blocks, in, r)
        mbar_wait_bounded(smem_u32(out_free + o), ((k / kTtOut) & 1) ^ 1);  // passes at once on the stage's first use
        TT_DISPATCH(kind, tt_write_blocks, out, r)
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            // the load stage is released only now: its values have provably left shared memory (the stores above consumed
            // them), and the producer refills it after this tile's store anyway, so nothing is lost by not signalling earlier
            mbar_arrive(smem_u32(empty + s));
            mbar_arrive(smem_u32(out_full + o));
        }
    }
}

// ---- tsnap_rows_tma_kernel: kModeRowsTma ------------------------------------------------------------------------
// Column shards / narrow views whose runs are short (<= kRowsTmaMaxRun = 1 KiB): the per-run copy-engine requests of the rows kernel are
// bound by the request rate (~46 cycles each), a tensor map moves a box of up to 256 runs with ONE request on each side.
// One lane per CTA drives a ring of 32 KiB stages, exactly like the dense bulk kernel: TMA tile load -> mbarrier ->
// TMA tile store -> bulk group; nothing passes through registers.
constexpr int kRtStages = 3;
constexpr int kRtCtasPerSm = 2;
constexpr uint32_t kRtSmemBytes = kRtStages * kTmaTileBytes + 1024 + 256;

struct RtMeta {
    int32_t c[5];
    const void* dmap;
};

__device__ __forceinline__ void rt_coords(const Member& m, uint32_t index, int32_t (&c)[5]) {
    const uint32_t rows = rows_tma_box_rows(m.inner);
    const uint32_t last = m.nouter - 1;
    const uint32_t tiles_r = (uint32_t)(((uint64_t)m.osize[last] + rows - 1) / rows);
    uint32_t rest = index / tiles_r;
    c[0] = 0;
    c[1] = (int32_t)((index - rest * tiles_r) * rows);
    c[2] = c[3] = c[4] = 0;
    int k = 2;
    for (int i = (int)last - 1; i >= 0; --i) {
        const uint32_t sz = (uint32_t)m.osize[i];
        c[k++] = (int32_t)(rest % sz);
        rest /= sz;
    }
}

__global__ void __launch_bounds__(32, kRtCtasPerSm)
tsnap_rows_tma_kernel(const Member* __restrict__ members, const Tile* __restrict__ tiles, const TmaPair* __restrict__ maps, uint32_t ntiles) {
    extern __shared__ unsigned char rt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(rt_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kRtStages * kTmaTileBytes);
    RtMeta* meta = reinterpret_cast<RtMeta*>(full + kRtStages);
    if (threadIdx.x != 0) return;  // one elected lane drives the TMA unit

    for (int s = 0; s < kRtStages; ++s) mbar_init(smem_u32(full + s), 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    const uint32_t first = blockIdx.x, step = gridDim.x;
    const uint32_t n_my = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    uint32_t fenced_member = 0xffffffffu;
    auto issue_load = [&](uint32_t k) {
        const Tile tl = tiles[first + k * step];
        const Member& m = members[tl.member];
        const TmaPair* mp = maps + (uint32_t)m.q_zero_point;
        if (tl.member != fenced_member) {
            tensormap_acquire(mp->src);
            tensormap_acquire(mp->dst);
            fenced_member = tl.member;
        }
        const uint32_t s = k % kRtStages;
        rt_coords(m, tl.index, meta[s].c);
        meta[s].dmap = mp->dst;
        const uint32_t bar = smem_u32(full + s);
        mbar_expect_tx(bar, (uint32_t)m.inner * rows_tma_box_rows(m.inner));  // the whole box, clipped parts are zero-filled
        tma_load_5d(smem_u32(smem + s * kTmaTileBytes), mp->src, meta[s].c, bar);
    };
    for (uint32_t k = 0; k < n_my && k < (uint32_t)kRtStages; ++k) issue_load(k);
    for (uint32_t k = 0; k < n_my; ++k) {
        const uint32_t s = k % kRtStages;
        mbar_wait_bounded(smem_u32(full + s), (k / kRtStages) & 1);
        fence_proxy_async_smem();
        tma_store_5d(meta[s].dmap, meta[s].c, smem_u32(smem + s * kTmaTileBytes));
        bulk_commit();
        // the stage stored one iteration ago is free once all but the newest store group have read shared memory
        if (k >= 1 && k - 1 + kRtStages < n_my) {
            bulk_wait_read<1>();
            issue_load(k - 1 + kRtStages);
        }
    }
    bulk_wait_all<0>();
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool transpose_tma_enabled() {
    // read per transposed member (a rare path), so that one process can A/B the two transposes
    const char* e = getenv("TSNAP_B200_TMA_TRANSPOSE");
    return !(e && e[0] == '0') && encode_tiled_fn() != nullptr;
}

bool make_tma_pair(const Member& m, TmaPair* out, uint32_t* variant) {
    static_assert(sizeof(CUtensorMap) == 128 && alignof(TmaPair) >= 64, "tensor map layout");
    if (m.mode != kModeTranspose || !transpose_tma_enabled()) return false;
    const uint32_t esz = m.unit;
    if (esz != 2 && esz != 4 && esz != 8) return false;
    const uint32_t A = m.shift & 255, B = (m.shift >> 8) & 255;
    if (m.nouter < 2 || m.nouter > 2 + kTmaMaxOther) return false;
    if ((m.src | m.dst) & 15) return false;
    // dims in tensor-map order: {A, B, others from the highest index down}; the same order on both sides except A <-> B
    int order[5], n = 0;
    order[n++] = (int)A;
    order[n++] = (int)B;
    for (int i = (int)m.nouter - 1; i >= 0; --i)
        if ((uint32_t)i != A && (uint32_t)i != B) order[n++] = i;
    cuuint64_t dim_s[5], dim_d[5], str_s[4], str_d[4];
    uint64_t extent_s = esz, extent_d = esz;
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        const int64_t sz = m.osize[i], ss = m.sstride[i], ds = m.dstride[i];
        if (sz <= 0 || sz >= (int64_t(1) << 32)) return false;
        if (k >= 1 && (ss <= 0 || (ss & 15) || ss >= (int64_t(1) << 40))) return false;  // A's source stride is the element size
        if (k != 1 && (ds <= 0 || (ds & 15) || ds >= (int64_t(1) << 40))) return false;  // B's destination stride is the element size
        extent_s = std::max<uint64_t>(extent_s, uint64_t(sz) * uint64_t(ss));
        extent_d = std::max<uint64_t>(extent_d, uint64_t(sz) * uint64_t(ds));
    }
    if (m.sstride[A] != int64_t(esz) || m.dstride[B] != int64_t(esz)) return false;
    const uint64_t pad_s = (extent_s + 15) & ~uint64_t(15), pad_d = (extent_d + 15) & ~uint64_t(15);
    if (pad_s >= (uint64_t(1) << 40) || pad_d >= (uint64_t(1) << 40)) return false;
    // source map: {A, B, o...}
    for (int k = 0; k < 5; ++k) {
        dim_s[k] = k < n ? cuuint64_t(m.osize[order[k]]) : 1;
        if (k >= 1) str_s[k - 1] = k < n ? cuuint64_t(m.sstride[order[k]]) : pad_s;
    }
    // destination map: {B, A, o...}
    int order_d[5];
    for (int k = 0; k < n; ++k) order_d[k] = order[k];
    order_d[0] = (int)B;
    order_d[1] = (int)A;
    for (int k = 0; k < 5; ++k) {
        dim_d[k] = k < n ? cuuint64_t(m.osize[order_d[k]]) : 1;
        if (k >= 1) str_d[k - 1] = k < n ? cuuint64_t(m.dstride[order_d[k]]) : pad_d;
    }
    // the tile shape that pads the two extents least (ties: the square-ish one, then the one long along A)
    uint32_t var = 0;
    uint64_t best = ~uint64_t(0);
    for (uint32_t v = 0; v < kTmaVariants; ++v) {
        const uint64_t ka = transpose_tma_side_a(esz, v), kb = transpose_tma_side_b(esz, v);
        const uint64_t padded = (uint64_t(m.osize[A]) + ka - 1) / ka * ((uint64_t(m.osize[B]) + kb - 1) / kb);
        if (padded < best) {
            best = padded;
            var = v;
        }
    }
    *variant = var;
    const cuuint32_t kA = transpose_tma_side_a(esz, var), kB = transpose_tma_side_b(esz, var);
    const cuuint32_t box_s[5] = {kA, kB, 1, 1, 1}, box_d[5] = {kB, kA, 1, 1, 1}, ones[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    EncodeTiledFn enc = encode_tiled_fn();
    CUtensorMap ms, md;
    if (enc(&ms, dt, 5, reinterpret_cast<void*>(uintptr_t(m.src)), dim_s, str_s, box_s, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (enc(&md, dt, 5, reinterpret_cast<void*>(uintptr_t(m.dst)), dim_d, str_d, box_d, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    memcpy(out->src, &ms, 128);
    memcpy(out->dst, &md, 128);
    return true;
}

cudaError_t init_transpose_tma() {
    cudaError_t e = cudaFuncSetAttribute(tsnap_transpose_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTtSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tsnap_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRtSmemBytes);
    return e;
}
bool rows_tma_enabled() {
    const char* e = getenv("TSNAP_B200_TMA_ROWS");
    return !(e && e[0] == '0') && encode_tiled_fn() != nullptr;
}

bool make_rows_tma_pair(const Member& m, TmaPair* out) {
    if (m.mode != kModeRows || !rows_tma_enabled()) return false;
    if (m.inner == 0 || m.inner > kRowsTmaMaxRun || (m.inner & 15)) return false;
    if (m.nouter < 1 || m.nouter > (uint32_t)kRowsTmaMaxOuter) return false;
    if ((m.src | m.dst) & 15) return false;
    cuuint64_t dim[5], str_s[4], str_d[4];
    dim[0] = m.inner / 8;
    uint64_t extent_s = m.inner, extent_d = m.inner;
    int n = 1;
    for (int i = (int)m.nouter - 1; i >= 0; --i, ++n) {
        const int64_t sz = m.osize[i], ss = m.sstride[i], ds = m.dstride[i];
        if (sz <= 0 || sz >= (int64_t(1) << 32)) return false;
        if (ss <= 0 || (ss & 15) || ss >= (int64_t(1) << 40) || ds <= 0 || (ds & 15) || ds >= (int64_t(1) << 40)) return false;
        dim[n] = cuuint64_t(sz);
        str_s[n - 1] = cuuint64_t(ss);
        str_d[n - 1] = cuuint64_t(ds);
        extent_s = std::max<uint64_t>(extent_s, uint64_t(sz) * uint64_t(ss));
        extent_d = std::max<uint64_t>(extent_d, uint64_t(sz) * uint64_t(ds));
    }
    const uint64_t pad_s = (extent_s + 15) & ~uint64_t(15), pad_d = (extent_d + 15) & ~uint64_t(15);
    if (pad_s >= (uint64_t(1) << 40) || pad_d >= (uint64_t(1) << 40)) return false;
    for (; n < 5; ++n) {
        dim[n] = 1;
        str_s[n - 1] = pad_s;
        str_d[n - 1] = pad_d;
    }
    const cuuint32_t box[5] = {cuuint32_t(m.inner / 8), rows_tma_box_rows(m.inner), 1, 1, 1}, ones[5] = {1, 1, 1, 1, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    CUtensorMap ms, md;
    if (enc(&ms, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, reinterpret_cast<void*>(uintptr_t(m.src)), dim, str_s, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (enc(&md, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, reinterpret_cast<void*>(uintptr_t(m.dst)), dim, str_d, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    memcpy(out->src, &ms, 128);
    memcpy(out->dst, &md, 128);
    return true;
}

cudaError_t launch_rows_tma(const Member* d_members, const Tile* d_tiles, const TmaPair* d_maps, uint32_t ntiles, int sm_count,
                            cudaStream_t stream) {
    if (ntiles == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)sm_count * kRtCtasPerSm;
    if (grid > ntiles) grid = ntiles;
    tsnap_rows_tma_kernel<<<grid, 32, kRtSmemBytes, stream>>>(d_members, d_tiles, d_maps, ntiles);
    return cudaGetLastError();
}


cudaError_t launch_transpose_tma(const Member* d_members, const Tile* d_tiles, const TmaPair* d_maps, uint32_t ntiles, int sm_count,
                                 cudaStream_t stream) {
    if (ntiles == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)sm_count;
    if (grid > ntiles) grid = ntiles;
    tsnap_transpose_tma_kernel<<<grid, kTtThreads, kTtSmemBytes, stream>>>(d_members, d_tiles, d_maps, ntiles);
    return cudaGetLastError();
}

}  // namespace tsnap
