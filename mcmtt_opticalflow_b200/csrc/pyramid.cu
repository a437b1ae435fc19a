// K1: fused Gaussian pyrDown + Scharr derivative of one pyramid level, batched over camera-frames.
//
// Replaces the pyramid / derivative construction that cv::calcOpticalFlowPyrLK repeats inside every
// call at psn_where/PSNWhere_Tracker2D.cpp:776 and :871:
// here it runs once per pushed frame.  Arithmetic is integer and bit-exact with OpenCV:
//   pyrDown : [1 4 6 4 1]^2, REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2)
//   Scharr  : dx = t0[x+1]-t0[x-1], t0 = 3(r[y-1]+r[y+1]) + 10 r[y];  dy = 3(t1[x-1]+t1[x+1]) + 10 t1[x],
//             t1 = r[y+1]-r[y-1]; interleaved int16 (dx,dy), REFLECT_101.
//
// One CTA (4 warps) owns a 128x32 tile of level l.  The u8 tile plus halo (160x35 bytes) is fetched by a
// single TMA 3-D box (x, y, plane) that completes on an mbarrier; out-of-image halo bytes arrive as zeros
// and edge tiles rewrite them with the REFLECT_101 mirror in shared memory.  Each lane then owns 4
// adjacent columns and slides down 8 output rows: the horizontal taps are dp4a dot products on the packed
// u8 words (no byte unpacking), the vertical taps run on registers.  Stores are 16-byte (4 x short2) per
// lane = 512 contiguous bytes per warp row for the derivative image and 2 bytes per lane for the next
// u8 level.  HBM traffic per level-l pixel: 1 B read, 4 B + 0.25 B written.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pyr_common.cuh"

namespace psnw {

namespace {

constexpr int kRowWords = kBoxW / 4;   // 40
constexpr int kLaneWord0 = kHaloL / 4 - 1;   // word holding columns x-4..x-1 of lane 0
constexpr int kRowsPerWarp = 8;

// REFLECT_101 apron of the levels >= 1 (common.cuh): what copyMakeBorder builds around every pyramid level inside
// cv::calcOpticalFlowPyrLK, written once per pushed frame so that the LK kernel stages coarse-level patches without border
// arithmetic.  A thread writes one aligned 4-byte word of the apron: LevelGeom::apron rows above and below the image over the
// padded width, and that many columns (rounded up to words) left and right of every image row; sources are the level's own
// pixels (L2-resident).
// The work is cut into regions of one CTA each (apron_regions): the regions of level l ride as extra CTAs in that level's
// own pyr_scharr_kernel launch (level l is complete by then and nothing in the launch writes it), so the apron costs
// neither a launch nor a serial tail.  (The fused multi-level kernel, pyramid_fused.cu, scatters the apron from the
// pixels each CTA holds in shared memory instead.)

// a mirrored row is written as whole words: cols = apron rounded up to 4 pixels left of the image; right of it the words
// start at the one that holds column w (w & ~3), one more than on the left
__host__ __device__ inline int apron_cols(int apron) { return (apron + 3) & ~3; }
// word slots per image row in the side bands: 16 while left + right words fit, else 32, 64, ...
__host__ __device__ inline int apron_slot_shift(int apron)
{
    const int words = 2 * (apron_cols(apron) / 4) + 1;
    int sh = 4;
    while ((1 << sh) < words) ++sh;
    return sh;
}

__device__ __forceinline__ int reflect101(int p, int n)
{
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);          // only positions beyond a single reflection (never sampled) are clamped
}

// the word of four REFLECT_101 sources of pixels (tx .. tx+3, ty).  The source columns are contiguous (ascending inside
// the image, descending where mirrored) unless the word straddles a fold: then two aligned word loads, a funnel shift and
// a byte reversal replace four byte loads.
__device__ __forceinline__ uint32_t apron_word_load(const uint8_t *plane0, int pitch, int W, int H, int tx, int ty)
{
    const uint8_t *srow = plane0 + (ptrdiff_t)reflect101(ty, H) * pitch;
    const int c0 = reflect101(tx, W), c3 = reflect101(tx + 3, W);
    uint32_t v = 0;
    if (c3 - c0 == 3 || c0 - c3 == 3) {
        const int lo = min(c0, c3);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(srow + (lo & ~3));
        const uint32_t a = __ldcg(wp), b = (lo & 3) ? __ldcg(wp + 1) : 0u;
        v = __funnelshift_r(a, b, 8 * (lo & 3));
        if (c0 > c3) v = __byte_perm(v, 0u, 0x0123);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v |= (uint32_t)__ldcg(srow + reflect101(tx + k, W)) << (8 * k);
    }
    return v;
}

// CTAs needed for the apron of a level with h rows: one per apron row above / below the image, then bands of
// nthreads >> slot_shift image rows
__host__ __device__ inline int apron_regions(int h, int apron, int nthreads)
{
    const int band = nthreads >> apron_slot_shift(apron);
    return 2 * apron + (h + band - 1) / band;
}

// region < 2 * apron: one apron row above / below the image over the padded width (threads stride over the words of the
// row); else a band of image rows, 2^slot_shift word slots per row = the words left of the row and those right of it (a
// word that straddles column w rewrites the image pixels it holds with their own values).
__device__ __forceinline__ void apron_region(const LevelGeom &L, int plane, int region, int tid, int nthreads)
{
    const int W = L.w, H = L.h, pitch = L.ipitch, A = L.apron, cols = apron_cols(A), side_l = cols / 4, side_r = side_l + 1;
    uint8_t *plane0 = L.img + (size_t)plane * L.iplane_px;
    if (region < 2 * A) {
        const int row_words = (W + 3) / 4 + 2 * side_l;
        const int ty = region < A ? region - A : H + (region - A);
        uint32_t *dst = reinterpret_cast<uint32_t *>(plane0 + (ptrdiff_t)ty * pitch - cols);
        for (int i = tid; i < row_words; i += nthreads) dst[i] = apron_word_load(plane0, pitch, W, H, 4 * i - cols, ty);
    } else {
        const int sh = apron_slot_shift(A), c = tid & ((1 << sh) - 1);
        if (c >= side_l + side_r) return;
        const int tx = c < side_l ? 4 * c - cols : (W & ~3) + 4 * (c - side_l);
        const int ty = (region - 2 * A) * (nthreads >> sh) + (tid >> sh);
        if (ty < H) *reinterpret_cast<uint32_t *>(plane0 + (ptrdiff_t)ty * pitch + tx) = apron_word_load(plane0, pitch, W, H, tx, ty);
    }
}

// Fix-up + arithmetic of one 128x32 tile of level `cur` held (with its halo) in `tile`: Scharr derivatives of the tile and,
// with kHasNext, the 64x16 tile of the next level.  Shared by the per-level kernel and the level walker below.
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
// the 128 compute threads of a CTA that also holds a helper warp (level walker)
struct ComputeSync { __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, 128;" ::: "memory"); } };

template <bool kHasNext, typename Sync = CtaSync>
__device__ __forceinline__ void tile_body(uint32_t *tile, const LevelGeom &cur, const LevelGeom &nxt, int plane, int x0, int y0, int tid)
{
    const Sync sync;
    // ---- REFLECT_101 fix-up of halo bytes that fell outside the image (TMA wrote zeros there) ----
    const int W = cur.w, H = cur.h;
    const bool edge_l = (x0 == 0), edge_r = (W - x0 + kHaloL + 1 < kBoxW);
    const bool edge_t = (y0 == 0), edge_b = (H - y0 + 2 < kBoxH);
    if (edge_l || edge_r || edge_t || edge_b) {
        uint8_t *tb = reinterpret_cast<uint8_t *>(tile);
        if (edge_l || edge_r) {
            if (tid < kBoxH) {
                uint8_t *row = tb + tid * kBoxW;
                if (edge_l) { row[kHaloL - 1] = row[kHaloL + 1]; row[kHaloL - 2] = row[kHaloL + 2]; }   // x=-1 <- 1, x=-2 <- 2
                if (edge_r) {
                    const int c = W - x0 + kHaloL;                                  // column of x = W
                    row[c] = row[c - 2];                                            // x=W   <- W-2
                    row[c + 1] = row[c - 3];                                        // x=W+1 <- W-3
                }
            }
            sync();
        }
        if (edge_t || edge_b) {
            for (int i = tid; i < kRowWords; i += 128) {
                if (edge_t) { tile[1 * kRowWords + i] = tile[3 * kRowWords + i]; tile[i] = tile[4 * kRowWords + i]; }
                if (edge_b) {
                    const int r = H - y0 + kHaloT;                                  // row of y = H
                    tile[r * kRowWords + i] = tile[(r - 2) * kRowWords + i];        // y=H   <- H-2
                    if (r + 1 < kBoxH) tile[(r + 1) * kRowWords + i] = tile[(r - 3) * kRowWords + i];   // y=H+1 <- H-3
                }
            }
            sync();
        }
    }

    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t *base = tile + (warp * kRowsPerWarp) * kRowWords + kLaneWord0 + lane;   // words (prev, cur, next) = base[0..2]
    const int ybase = y0 + warp * kRowsPerWarp;                               // first output row of this warp

    int hd[3][4], hs[3][4];     // horizontal difference / 3-10-3 smooth of the last three rows
    int pg[5][2];               // horizontal 1-4-6-4-1 of the last five rows (even columns only)

    uint4 *dxy_out = reinterpret_cast<uint4 *>(cur.dxy + (size_t)plane * cur.plane_px + (size_t)ybase * cur.pitch + x0) + lane;
    const int dxy_row_u4 = cur.pitch / 4;
    const bool pad_cols = x0 + kTileW > W;

#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp + 3; ++rr) {
        const uint32_t wp = base[rr * kRowWords], wc = base[rr * kRowWords + 1], wn = base[rr * kRowWords + 2];
        if (rr >= 1) {
            const uint32_t sa = __funnelshift_r(wp, wc, 24);    // bytes x-1 .. x+2
            const uint32_t sb = __funnelshift_r(wc, wn, 8);     // bytes x+1 .. x+4
            int *d = hd[rr % 3], *s = hs[rr % 3];
            d[0] = dp4a_us(sa, kDiffLo, 0);   d[1] = dp4a_us(sa, kDiffHi, 0);
            d[2] = dp4a_us(sb, kDiffLo, 0);   d[3] = dp4a_us(sb, kDiffHi, 0);
            s[0] = dp4a_us(sa, kSmoothLo, 0); s[1] = dp4a_us(sa, kSmoothHi, 0);
            s[2] = dp4a_us(sb, kSmoothLo, 0); s[3] = dp4a_us(sb, kSmoothHi, 0);
        }
        if (kHasNext) {
            const uint32_t sc = __funnelshift_r(wp, wc, 16);    // bytes x-2 .. x+1
            pg[rr % 5][0] = dp4a_us(sc, kGauss, (int)((wc >> 16) & 0xffu));   // + byte x+2
            pg[rr % 5][1] = dp4a_us(wc, kGauss, (int)(wn & 0xffu));           // centre x+2: bytes x..x+3, + byte x+4
        }
        if (rr >= 3) {
            const int k = rr - 3;                               // output row ybase + k: rows rr-2, rr-1, rr
            const int *dm = hd[(rr - 2) % 3], *d0 = hd[(rr - 1) % 3], *dp = hd[rr % 3];
            const int *sm = hs[(rr - 2) % 3], *sp = hs[rr % 3];
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int dx = 3 * (dm[i] + dp[i]) + 10 * d0[i];
                const int dy = sp[i] - sm[i];
                o[i] = __byte_perm((uint32_t)dx, (uint32_t)dy, 0x5410);
                if (pad_cols && x0 + 4 * lane + i >= W) o[i] = 0u;   // columns [w, pitch) stay zero: the LK kernel reads them
            }
            if (ybase + k < H) dxy_out[(size_t)k * dxy_row_u4] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (kHasNext && rr >= 4 && (rr & 1) == 0) {
            const int m = (rr - 4) >> 1;                        // rows rr-4 .. rr, centre rr-2
            const int j = (ybase >> 1) + m;
            int v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                v[i] = (pg[(rr - 4) % 5][i] + pg[rr % 5][i]) + 4 * (pg[(rr - 3) % 5][i] + pg[(rr - 1) % 5][i]) +
                       6 * pg[(rr - 2) % 5][i];
            const int ox = (x0 >> 1) + 2 * lane;
            if (j < nxt.h && ox < nxt.w) {
                const int p0 = (v[0] + 128) >> 8, p1 = (v[1] + 128) >> 8;
                uint8_t *g = nxt.img + (size_t)plane * nxt.iplane_px + (size_t)j * nxt.ipitch + ox;
                const bool two = ox + 1 < nxt.w;
                if (two) *reinterpret_cast<uint16_t *>(g) = (uint16_t)(p0 | (p1 << 8));
                else     g[0] = (uint8_t)p0;
            }
        }
    }
}

template <bool kHasNext>
__global__ void __launch_bounds__(128)
pyr_scharr_kernel(const __grid_constant__ CUtensorMap map, const LevelGeom cur, const LevelGeom nxt,
                  const PlaneList planes, const int tiles_y, const int pdl_mode)
{
    // pdl_mode 0: release the dependents at once, wait before the first read (the next level's launch is set up meanwhile)
    //          1: late_wait -- level 0 of planes that no kernel before it in the stream touches (see launch_pyr_fused): the
    //             launch runs beside its predecessors and only completes after them
    //          2: wait first, THEN release the dependents -- the last launch of a build, so that a late_wait launch of the
    //             next build never runs beside anything older than this launch's wait
    const bool late_wait = pdl_mode == 1;
    if ((int)blockIdx.y >= tiles_y) {
        // extra CTAs of a level >= 1: the REFLECT_101 apron of this (input) level, complete since the previous launch
        // (never launched with late_wait: level 0 has no apron)
        if (pdl_mode != 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (pdl_mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const int region = ((int)blockIdx.y - tiles_y) * (int)gridDim.x + (int)blockIdx.x;
        if (region < apron_regions(cur.h, cur.apron, 128)) apron_region(cur, planes.plane[blockIdx.z], region, threadIdx.x, 128);
        return;
    }
    __shared__ __align__(128) uint32_t tile[kBoxH * kRowWords];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
    const int plane = planes.plane[blockIdx.z];
    const uint32_t bar_a = smem_u32(&bar);

#ifndef PSNW_DEBUG_NO_TMA
    // Programmatic dependent launch: let the next level's grid start its prologue now, and wait for the previous
    // level's grid (which wrote this level's u8 image) to complete before the first read.
    if (pdl_mode != 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) mbar_init(bar_a, 1);
    __syncthreads();
    if (!late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pdl_mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        mbar_expect_tx(bar_a, kBoxW * kBoxH);
        tma_load_3d(smem_u32(tile), &map, x0 - kHaloL, y0 - kHaloT, plane, bar_a);
    }
    mbar_wait(bar_a, 0);
#else   // bring-up aid only (never built into libpsnw_klt.so): same tile contents through plain loads
    (void)bar_a;
    for (int i = tid; i < kBoxW * kBoxH; i += 128) {
        const int r = i / kBoxW, c = i - r * kBoxW;
        const int gx = x0 - kHaloL + c, gy = y0 - kHaloT + r;
        uint8_t v = 0;
        if (gx >= 0 && gy >= 0 && gx < cur.w && gy < cur.h) v = cur.img[(size_t)plane * cur.iplane_px + (size_t)gy * cur.ipitch + gx];
        reinterpret_cast<uint8_t *>(tile)[i] = v;
    }
    __syncthreads();
#endif

    tile_body<kHasNext>(tile, cur, nxt, plane, x0, y0, tid);
    if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- level walker: every level of every frame of a batch in ONE persistent launch ------------------------------------------
// The per-level kernel is bandwidth-efficient, but a batch pays one dependent launch (and one tail) per level, and the coarse
// levels of small frames are mostly launch latency.  Here the tiles of all levels form one ordered work list (level-major,
// then plane, then tile) drawn from a device counter by a grid that just fills the machine.  A tile of level l >= 1 may start
// once every tile of level l-1 of its frame has signalled completion (one counter per (level, frame); items are drawn in
// order and an item only ever waits for items drawn before it, so every wait ends).  Level l of one frame runs while level
// l-1 of the following frames is still being produced: no launch boundary, no per-level tail.
//
// A CTA is a small pipeline: four compute warps and one helper warp around kWalkStages tile buffers.
//   helper (one thread)  draws items, polls their dependency, issues the TMA box of each into a free stage (mbarrier `full`),
//                        and retires finished tiles: it watches the `done` barrier of every stage in flight and publishes
//                        their completion with ONE fence + one reduction per tile.  The fence (which has to wait for the
//                        tile's stores to reach L2) and the draw therefore never hold up the compute warps or the copies.
//   compute (128 threads) wait for a stage, run the tile arithmetic of the per-level kernel, scatter the REFLECT_101 apron of
//                        levels >= 1 from the tile they hold (pyr_common.cuh), arrive on `done`, move to the next stage.
// The inputs of the levels >= 1 were written with generic stores by other CTAs and are read by the TMA unit: the helper
// orders its acquire before the copy with fence.proxy.async.
struct WalkParams {
    CUtensorMap map[PSNW_MAX_LEVELS];
    LevelGeom lv[PSNW_MAX_LEVELS];
    int tiles_x[PSNW_MAX_LEVELS], tiles[PSNW_MAX_LEVELS];    // tiles per row / per plane of a level
    int begin[PSNW_MAX_LEVELS + 1];                             // first work item of a level
    int levels, n_planes;
    int base;                                                   // first level of the walk (levels below it are complete in HBM)
    unsigned *ctl;                                              // [0] next item, [1] CTAs that left, [2 + l * n_planes + i] finished tiles
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(phase)
        : "memory");
    return ok != 0;
}

constexpr int kWalkCompute = 128, kWalkThreads = kWalkCompute + 32, kWalkStages = 3, kWalkCtasPerSm = 7;
constexpr int kWalkSigRing = 16, kWalkSigBatch = 8;            // finished tiles are published in batches: one fence per batch
constexpr int kWalkStageWords = (kBoxH * kRowWords + 31) & ~31;      // 128-byte aligned tile buffers

struct WalkItem {
    int level;         // < 0: no more work
    int plane, x0, y0;
    int edge;          // the tile holds pixels whose mirror images lie in the level's apron
};

__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm)
pyr_walk_kernel(const __grid_constant__ WalkParams P, const __grid_constant__ PlaneList planes)
{
    __shared__ __align__(128) uint32_t tiles[kWalkStages * kWalkStageWords];
    __shared__ __align__(8) uint64_t bar_full[kWalkStages], bar_done[kWalkStages];
    __shared__ ApronWork apron[kWalkStages];
    __shared__ WalkItem item[kWalkStages];
    __shared__ int s_last;
    const int tid = threadIdx.x;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int s = 0; s < kWalkStages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_done[s]), kWalkCompute / 32);
        }
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the control block and the frames belong to the launches before

    if (tid >= kWalkCompute) {
        // ---------------------------------------------------------------- helper warp
        if (tid == kWalkCompute) {
            const int total = P.begin[P.levels];
            int k_load = 0, k_free = 0, k_sig = 0;               // items loaded / finished by the compute warps / published
            int sig[kWalkSigRing];                               // completion counter of item k (index into ctl), or -1
            int it = (int)atomicAdd(&P.ctl[0], 1u);              // two draws ahead: the counter's latency stays off the path
            int it_next = (int)atomicAdd(&P.ctl[0], 1u);
            for (;;) {
                if (k_free < k_load && mbar_test(smem_u32(&bar_done[k_free % kWalkStages]), (uint32_t)(k_free / kWalkStages) & 1u)) ++k_free;
                const bool room = k_load - k_free < kWalkStages && k_load - k_sig < kWalkSigRing;
                bool loaded = false;
                if (it < total && room) {
                    int l = P.base;
#pragma unroll
                    for (int k = 1; k < PSNW_MAX_LEVELS; ++k) l += (k > P.base && k < P.levels && it >= P.begin[k]) ? 1 : 0;
                    const int q = it - P.begin[l], pi = q / P.tiles[l], t = q - pi * P.tiles[l];
                    if (l == P.base || ld_acquire_u32(&P.ctl[2 + (l - 1) * P.n_planes + pi]) >= (unsigned)P.tiles[l - 1]) {
                        const int ty = t / P.tiles_x[l], tx = t - ty * P.tiles_x[l];
                        const int x0 = tx * kTileW, y0 = ty * kTileH, A = P.lv[l].apron, s = k_load % kWalkStages;
                        WalkItem w;
                        w.level = l; w.plane = planes.plane[pi]; w.x0 = x0; w.y0 = y0;
                        w.edge = A > 0 && (x0 <= A || x0 + kTileW >= P.lv[l].w - 1 - A || y0 <= A || y0 + kTileH >= P.lv[l].h - 1 - A);
                        item[s] = w;
                        asm volatile("fence.proxy.async;" ::: "memory");     // generic writes (other CTAs' pixels, fix-ups of the stage) before the copy
                        const uint32_t full = smem_u32(&bar_full[s]);
                        mbar_expect_tx(full, kBoxW * kBoxH);
                        tma_load_3d(smem_u32(tiles + s * kWalkStageWords), &P.map[l], x0 - kHaloL, y0 - kHaloT, w.plane, full);
#pragma unroll
                        for (int j = 0; j < kWalkSigRing; ++j)
                            if (j == k_load % kWalkSigRing) sig[j] = l + 1 < P.levels ? 2 + l * P.n_planes + pi : -1;
                        ++k_load;
                        it = it_next;
                        it_next = (int)atomicAdd(&P.ctl[0], 1u);
                        loaded = true;
                    }
                }
                // finished tiles are published in batches (one fence each), at once when there is nothing to load
                if (k_sig < k_free && (!loaded || k_free - k_sig >= kWalkSigBatch)) {
                    __threadfence();
                    for (; k_sig < k_free; ++k_sig) {
                        int f = -1;
#pragma unroll
                        for (int j = 0; j < kWalkSigRing; ++j)
                            if (j == k_sig % kWalkSigRing) f = sig[j];
                        if (f >= 0) atomicAdd(&P.ctl[f], 1u);
                    }
                    continue;
                }
                if (loaded) continue;
                if (it >= total && k_sig == k_load) break;
                __nanosleep(20);
            }
            // every stage is free: tell the compute warps to leave
            item[k_load % kWalkStages].level = -1;
            mbar_arrive(smem_u32(&bar_full[k_load % kWalkStages]));
            __threadfence();
            s_last = atomicAdd(&P.ctl[1], 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncwarp();
        // the last CTA to leave returns the control block to zero for the next launch
        if (s_last)
            for (int i = tid - kWalkCompute; i < 2 + P.levels * P.n_planes; i += 32) P.ctl[i] = 0u;
        return;
    }

    // -------------------------------------------------------------------- compute warps
    for (int k = 0;; ++k) {
        const int s = k % kWalkStages;
        mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(k / kWalkStages) & 1u);
        const WalkItem w = item[s];
        if (w.level < 0) break;
        const LevelGeom &cur = P.lv[w.level];
        const bool has_next = w.level + 1 < P.levels;
        const LevelGeom &nxt = P.lv[has_next ? w.level + 1 : w.level];
        uint32_t *tile = tiles + s * kWalkStageWords;
        if (w.edge) {
            if (tid == 32) apron_work_fill(apron[s], cur.w, cur.h, cur.apron, w.x0, w.y0, kTileW, kTileH);
            ComputeSync()();
        }
        if (has_next) tile_body<true, ComputeSync>(tile, cur, nxt, w.plane, w.x0, w.y0, tid);
        else          tile_body<false, ComputeSync>(tile, cur, nxt, w.plane, w.x0, w.y0, tid);
        if (w.edge)
            apron_scatter(apron[s], reinterpret_cast<const uint8_t *>(tile), kBoxW, w.x0 - kHaloL, w.y0 - kHaloT, w.x0, kTileW, 3, cur.w, cur.h,
                          cur.apron, cur.img + (size_t)w.plane * cur.iplane_px, cur.ipitch, tid, kWalkCompute);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(&bar_done[s]));      // the warp's reads of the stage and its stores are behind it
    }
}

// BGR (8UC3) -> gray, cv::cvtColor(BGR2GRAY) fixed point (Tracker2D.cpp:257):
//   15 bit: (3735 B + 19235 G + 9798 R + 2^14) >> 15 (OpenCV 4.x);  14 bit: (1868 B + 9617 G + 4899 R + 2^13) >> 14 (2.4.6)
__global__ void __launch_bounds__(256)
bgr2gray_kernel(const uint8_t *__restrict__ bgr, size_t bgr_pitch, uint8_t *__restrict__ gray, int gray_pitch,
                int w, int h, int luma_bits)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= w || y >= h) return;
    const uint8_t *src = bgr + (size_t)y * bgr_pitch + (size_t)x4 * 3;
    const int cb = luma_bits == 14 ? 1868 : 3735, cg = luma_bits == 14 ? 9617 : 19235, cr = luma_bits == 14 ? 4899 : 9798;
    const int rnd = 1 << (luma_bits - 1);
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x4 + i < w) {
            const int v = (cb * src[3 * i] + cg * src[3 * i + 1] + cr * src[3 * i + 2] + rnd) >> luma_bits;
            out |= (uint32_t)v << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t *>(gray + (size_t)y * gray_pitch + x4) = out;
}

}  // namespace

int launch_pyr_level(const PyrDesc &pyr, int level, const CUtensorMap *map_in, const PlaneList &planes, int with_apron, int pdl_mode,
                     cudaStream_t stream)
{
    const LevelGeom &cur = pyr.lv[level];
    const bool has_next = level + 1 < pyr.levels;
    const LevelGeom &nxt = pyr.lv[has_next ? level + 1 : level];
    cudaLaunchConfig_t cfg = {};
    const int tiles_x = (cur.w + kTileW - 1) / kTileW, tiles_y = (cur.h + kTileH - 1) / kTileH;
    const int apron_y = (cur.apron && with_apron) ? (apron_regions(cur.h, cur.apron, 128) + tiles_x - 1) / tiles_x : 0;   // extra grid rows
    cfg.gridDim = dim3(tiles_x, tiles_y + apron_y, planes.n);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // PDL: overlap this launch with the previous level's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = has_next ? cudaLaunchKernelEx(&cfg, pyr_scharr_kernel<true>, *map_in, cur, nxt, planes, tiles_y, pdl_mode)
                                   : cudaLaunchKernelEx(&cfg, pyr_scharr_kernel<false>, *map_in, cur, nxt, planes, tiles_y, pdl_mode);
    return e == cudaSuccess ? 1 : -1;
}

// all levels of the planes in one persistent launch (pyr_walk_kernel).  ctl: device control block of walk_ctl_words(levels,
// planes.n) zeroed words, exclusively this launch's until it completes (the kernel leaves it zeroed).
size_t walk_ctl_words(int levels, int n_planes) { return 2 + (size_t)levels * n_planes; }

int launch_pyr_walk(const PyrDesc &pyr, int base, const CUtensorMap *maps, const PlaneList &planes, unsigned *ctl, int ctas_per_sm, int sm_count,
                    cudaStream_t stream)
{
    static int occ = 0;
    if (!occ && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pyr_walk_kernel, kWalkThreads, 0) != cudaSuccess) return -1;
    WalkParams P;
    memset(&P, 0, sizeof P);
    P.levels = pyr.levels;
    P.n_planes = planes.n;
    P.ctl = ctl;
    P.base = base;
    int begin = 0;
    for (int l = 0; l < pyr.levels; ++l) {
        P.map[l] = maps[l];
        P.lv[l] = pyr.lv[l];
        P.tiles_x[l] = (pyr.lv[l].w + kTileW - 1) / kTileW;
        P.tiles[l] = P.tiles_x[l] * ((pyr.lv[l].h + kTileH - 1) / kTileH);
        P.begin[l] = begin;
        if (l >= base) begin += P.tiles[l] * planes.n;
    }
    for (int l = pyr.levels; l <= PSNW_MAX_LEVELS; ++l) P.begin[l] = begin;
    const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(std::min(begin, per_sm * sm_count));
    cfg.blockDim = dim3(kWalkThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pyr_walk_kernel, P, planes) == cudaSuccess ? 1 : -1;
}

int launch_bgr2gray(const uint8_t *bgr, size_t bgr_pitch, uint8_t *gray, int gray_pitch, int w, int h,
                    int luma_bits, cudaStream_t stream)
{
    dim3 grid((w + 1023) / 1024, h);
    bgr2gray_kernel<<<grid, 256, 0, stream>>>(bgr, bgr_pitch, gray, gray_pitch, w, h, luma_bits);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
