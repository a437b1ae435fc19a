// Device helpers shared by the per-level (pyramid.cu) and the fused multi-level (pyramid_fused.cu) pyramid kernels:
// TMA / mbarrier wrappers and the dp4a coefficient words of the separable integer filters.
#pragma once
#include "common.cuh"

namespace psnw {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(phase)
            : "memory");
    } while (!ok);
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// ---- REFLECT_101 apron of a level >= 1 (common.cuh), scattered from the pixels a CTA holds in shared memory --------------
// What copyMakeBorder builds around every pyramid level inside cv::calcOpticalFlowPyrLK, written once per pushed frame so that
// the LK kernel stages coarse-level patches without border arithmetic.  Every CTA of a pyramid kernel owns a region of the
// level (its tile) and holds it -- plus a few halo columns -- in shared memory; it writes the mirror images of ITS OWN pixels
// that lie within `apron` pixels of a frame edge: no extra launch, no extra CTAs, no dependency between CTAs.
// A CTA handles every apron word (4 aligned pixels) that mirrors at least one pixel of its region.  If all four sources lie in
// the part of its buffer that holds true pixels (region + hv columns) the word is stored whole; otherwise only the bytes whose
// sources it owns (the neighbouring CTA stores the others).  Words stored by two CTAs carry identical values.
//   x classes: 0 identity columns (only with mirrored rows), 1 mirrored about x = 0, 2 mirrored about x = W-1; y alike.
// The destination rectangles of the 8 classes (words x rows) are tabulated once per CTA by a single thread while the tile is
// still in flight: {rows, words per row, first word, first row}.
struct ApronWork {
    int4 cls[8];
    int total;
};

__device__ __forceinline__ int reflect101c(int p, int n)
{
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);
}

// region [ox, ox + own_w) x [oy, oy + own_h) of a W x H level with an apron of A pixels
__device__ __forceinline__ void apron_work_fill(ApronWork &T, int W, int H, int A, int ox, int oy, int own_w, int own_h)
{
    const int Ax = min(A, W - 1), Ay = min(A, H - 1);
    const int ox1 = min(ox + own_w, W), oy1 = min(oy + own_h, H);                   // region clipped to the frame
    int wlo[3], nw[3], rlo[3], nr[3];
    wlo[0] = ox >> 2; nw[0] = ((ox1 - 1) >> 2) - wlo[0] + 1;
    { const int a = max(ox, 1), b = min(ox1 - 1, Ax); wlo[1] = (-b) >> 2; nw[1] = a <= b ? ((-a) >> 2) - wlo[1] + 1 : 0; }
    { const int a = max(ox, W - 1 - Ax), b = min(ox1 - 1, W - 2); wlo[2] = (2 * (W - 1) - b) >> 2; nw[2] = a <= b ? ((2 * (W - 1) - a) >> 2) - wlo[2] + 1 : 0; }
    rlo[0] = oy; nr[0] = oy1 - oy;
    { const int a = max(oy, 1), b = min(oy1 - 1, Ay); rlo[1] = -b; nr[1] = a <= b ? b - a + 1 : 0; }
    { const int a = max(oy, H - 1 - Ay), b = min(oy1 - 1, H - 2); rlo[2] = 2 * (H - 1) - b; nr[2] = a <= b ? b - a + 1 : 0; }
    const bool none = A <= 0 || ox >= ox1 || oy >= oy1;
    int total = 0;
#pragma unroll
    for (int c = 1; c < 9; ++c) {
        const int xc = c % 3, yc = c / 3;
        const int rows = none || nw[xc] <= 0 ? 0 : max(nr[yc], 0);
        T.cls[c - 1] = make_int4(rows, max(nw[xc], 1), wlo[xc], rlo[yc]);
        total += rows * max(nw[xc], 0);
    }
    T.total = total;
}

// four consecutive bytes of shared memory from any address, as one word (byte 0 = lowest address)
__device__ __forceinline__ uint32_t lds_unaligned(const uint8_t *p)
{
    const uint32_t a = smem_u32(p), sh = (a & 3u) * 8u;
    uint32_t w0, w1 = 0u;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a & ~3u));
    if (sh) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"((a & ~3u) + 4u));
    return __funnelshift_r(w0, w1, sh);
}

// buf: the CTA's shared-memory buffer of the level, row pitch `pitch` bytes, holding pixel (bx, by) at its first byte; hv =
// columns beside the region that hold true pixels (where they are inside the frame); plane0 = pixel (0,0) of the plane in HBM.
// The nt threads (a multiple of 32; tid = position in them) share the work: the eight classes go to the warps in turn (with
// 16 warps, two per class), inside a warp groups of 8 / 16 / 32 lanes take a row of words each.  A word whose four sources
// are adjacent buffer bytes that hold true pixels (the rule, away from the words that straddle a fold or the frame edge) is one
// unaligned shared-memory read, reversed where mirrored, and one store; it is stored whole even where it reaches past the
// apron into the row margin (nothing reads those bytes).
__device__ __forceinline__ void apron_scatter(const ApronWork &T, const uint8_t *buf, int pitch, int bx, int by, int ox, int own_w, int hv,
                                              int W, int H, int A, uint8_t *plane0, int ipitch, int tid, int nt)
{
    if (T.total == 0) return;
    const int Ax = min(A, W - 1), Aw = (Ax + 3) & ~3;
    const int ox1 = min(ox + own_w, W);
    const int vx0 = max(ox - hv, 0), vx1 = min(ox + own_w + hv, W);
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    const int per_class = nwarps >= 8 ? nwarps >> 3 : 1;          // warps that share one class
    const int sub = nwarps >= 8 ? warp >> 3 : 0;
#pragma unroll 1
    for (int e = warp & 7; e < 8; e += nwarps >= 8 ? 8 : nwarps) {
        const int4 C = T.cls[e];                                  // {rows, words per row, first word, first row}
        if (C.x == 0) continue;
        const int xc = (e + 1) % 3, yc = (e + 1) / 3;
        const int sh = C.y <= 8 ? 3 : C.y <= 16 ? 4 : 5, cw = 1 << sh, gpw = 32 >> sh;
        const int c0 = lane & (cw - 1);
#pragma unroll 1
        for (int r = sub * gpw + (lane >> sh); r < C.x; r += per_class * gpw) {
            const int ty = C.w + r;
            const uint8_t *srow = buf + (reflect101c(ty, H) - by) * pitch - bx;
            uint8_t *drow = plane0 + (ptrdiff_t)ty * ipitch;
#pragma unroll 1
            for (int c = c0; c < C.y; c += cw) {
                const int tx = 4 * (C.z + c);
                if (xc == 0) {
                    if (tx + 4 <= W) {                            // identity columns, mirrored row: one aligned word
                        *reinterpret_cast<uint32_t *>(drow + tx) = *reinterpret_cast<const uint32_t *>(srow + tx);
                        continue;
                    }
                } else {
                    // destinations tx .. tx+3 all beyond the frame on this side; their sources lo .. lo+3 descend with x
                    const int lo = (xc == 1 ? 0 : 2 * (W - 1)) - tx - 3;
                    const bool beyond = xc == 1 ? (tx >= -Aw && tx + 3 < 0) : (tx >= W && tx < W + Ax);
                    if (beyond && lo >= vx0 && lo + 4 <= vx1) {
                        *reinterpret_cast<uint32_t *>(drow + tx) = __byte_perm(lds_unaligned(srow + lo), 0u, 0x0123);
                        continue;
                    }
                }
                // the words at a fold, at the frame edge or at the end of the apron: byte by byte
                uint32_t v = 0;
                bool all = true, any = false, own[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int px = tx + k;
                    const bool okd = px >= -Ax && px < W + Ax && (yc != 0 || px < 0 || px >= W);
                    const int sx = reflect101c(px, W);
                    own[k] = okd && sx >= ox && sx < ox1;
                    all = all && okd && sx >= vx0 && sx < vx1;
                    any = any || own[k];
                    v |= (uint32_t)srow[min(max(sx, vx0), vx1 - 1)] << (8 * k);
                }
                if (all && any) {
                    *reinterpret_cast<uint32_t *>(drow + tx) = v;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (own[k]) drow[tx + k] = (uint8_t)(v >> (8 * k));
                }
            }
        }
    }
}

// dp4a coefficient words (byte 0 = lowest address)
constexpr int kDiffLo = 0x000100FF;     // (-1, 0, 1, 0)
constexpr int kDiffHi = 0x0100FF00;     // ( 0,-1, 0, 1)
constexpr int kSmoothLo = 0x00030A03;   // ( 3,10, 3, 0)
constexpr int kSmoothHi = 0x030A0300;   // ( 0, 3,10, 3)
constexpr int kGauss = 0x04060401;      // ( 1, 4, 6, 4)

}  // namespace
}  // namespace psnw
