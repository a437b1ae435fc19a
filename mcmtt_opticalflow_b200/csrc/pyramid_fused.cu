// K1f: the whole Gaussian pyramid + Scharr derivatives + REFLECT_101 aprons of a batch of frames in ONE launch (up to 4 levels).
//
// The per-level kernel (pyramid.cu) is bandwidth-efficient for large frames but a rig of small cameras (4 x 768x576 =
// 12 MB, or 64 of them) pays one dependent launch per level, each with its own tail, for a few microseconds of HBM traffic.
// Here one CTA owns a 192x64 region of level 0 (96x32, 48x16, 24x8 of the next levels) and produces every level for it
// from a single TMA box of level 0: the halo each coarser level needs (1 px for Scharr, 2n+2 px for the pyrDown that
// feeds a level needing n) is recomputed locally -- 22 / 10 / 4 / 1 px at levels 0..3 -- so CTAs never wait for each
// other.  HBM traffic is the algorithmic N(6 S_L - 1) bytes plus the aprons; the overlap between neighbouring boxes is
// served by L2.
//
// Two phases per CTA:
//   1. the chain: level k -> level k+1 (pyrDown) in shared memory, one block barrier per level.  Every thread owns one
//      word column (4 pixels) and slides down a band of rows: three shared-memory words per input row, horizontal taps
//      as dp4a, the five-row window in registers, one 2-pixel store per output row.
//   2. the outputs, without any barrier: Scharr derivatives of the own region of every level (uniform work items of 4
//      columns x 4 rows drawn from one flat list over all levels, so that every thread carries the same load), the u8
//      levels >= 1, and the REFLECT_101 apron (common.cuh) of the levels >= 1: every CTA writes the mirror images of ITS
//      OWN pixels that lie within `apron` pixels of a frame edge -- no second launch, no dependency between CTAs.
//
// Arithmetic is the same integer dp4a code as pyramid.cu: bit-exact with cv::pyrDown / calcSharrDeriv, REFLECT_101
// applied at every level's own image border (x = -1, -2, W, W+1 and the same rows) inside shared memory.
#include "common.cuh"
#include "pyr_common.cuh"

namespace psnw {

namespace {

constexpr int kTopW = 24;                            // width of the region owned at the coarsest fused level (192 px of level 0:
                                                     // with its halo the widest TMA box, 256 bytes)
constexpr int kScharrRows = 4;                       // rows per Scharr work item (every level)

// NL fused levels; TH = height of the region owned at the coarsest level: 8 (192x64 of level 0: a rig frame of a few small
// cameras spreads over every SM), 16 or 32 (192x128 / 192x256: large batches, where the halo rows recomputed by every CTA
// -- 22 above and below at level 0 -- cost less the taller the region is)
template <int NL, int TH> struct FG {
    static __host__ __device__ constexpr int own_w(int k) { return kTopW << (NL - 1 - k); }
    static __host__ __device__ constexpr int own_h(int k) { return TH << (NL - 1 - k); }
    static __host__ __device__ constexpr int need(int k) { return k >= NL - 1 ? 1 : 2 * need(k + 1) + 2; }
    // columns kept left (= right) of the own region: >= need, multiple of 4 (words), of 16 at level 0 (TMA start)
    static __host__ __device__ constexpr int left(int k) { return k == 0 ? (need(0) + 15) / 16 * 16 : (need(k) + 3) / 4 * 4; }
    static __host__ __device__ constexpr int pitch(int k) { return own_w(k) + 2 * left(k); }
    static __host__ __device__ constexpr int rows(int k) { return own_h(k) + 2 * need(k); }
    static __host__ __device__ constexpr int offset(int k) { return k == 0 ? 128 : offset(k - 1) + pitch(k - 1) * rows(k - 1); }
    // (the tail pad absorbs the reads of a partial last band of the chain, whose results are never stored)
    static __host__ __device__ constexpr int bytes() { return offset(NL - 1) + pitch(NL - 1) * rows(NL - 1) + 12288; }
    // level-0 box of the TMA loads: the buffer rows in chunks of at most 256
    static __host__ __device__ constexpr int n_box() { return (rows(0) + 255) / 256; }
    static __host__ __device__ constexpr int box_h() { return rows(0) / n_box(); }
    // Scharr work items of level k: 4 columns x kScharrRows rows
    static __host__ __device__ constexpr int sch_cols(int k) { return own_w(k) / 4; }
    static __host__ __device__ constexpr int sch_items(int k) { return sch_cols(k) * (own_h(k) / kScharrRows); }
    static __host__ __device__ constexpr int sch_begin(int k) { return k == 0 ? 0 : sch_begin(k - 1) + sch_items(k - 1); }
    // u8 copy items (words) of level k >= 1
    static __host__ __device__ constexpr int cp_items(int k) { return k == 0 ? 0 : (own_w(k) / 4) * own_h(k); }
    static __host__ __device__ constexpr int cp_begin(int k) { return k <= 1 ? 0 : cp_begin(k - 1) + cp_items(k - 1); }
};

struct FusedLevels {
    LevelGeom lv[4];
    int base;          // pyramid level of lv[0]; > 0: the chain starts from a level that carries an apron, which is written here too
    int late_wait;     // the launch shares no data with the kernels before it in its stream (a build of other planes): it starts
                       // at once and waits for them only before it completes, so that stream order still holds for what follows
};

#ifdef PSNW_PYR_TIMELINE          // bring-up aid (tools/pyr_timeline.cu): per-CTA time stamps at the phase boundaries
__device__ unsigned long long g_timeline[8 * 65536];
#define PSNW_STAMP(slot)                                                                                           \
    do {                                                                                                           \
        __syncthreads();                                                                                           \
        if (threadIdx.x == 0) {                                                                                    \
            unsigned long long t_;                                                                                 \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                 \
            g_timeline[8 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) + (slot)] = t_;        \
        }                                                                                                          \
    } while (0)
#else
#define PSNW_STAMP(slot) do { } while (0)
#endif

// REFLECT_101 at the border of a level held in shared memory: x = -1 (and -2 when `two`), x = W (and W+1), same rows.
// Columns first, then whole rows, so that the corners come out right.  Ends with a barrier when it wrote anything.
template <int NT>
__device__ __forceinline__ void fix_border(uint8_t *buf, int pitch, int rows, int bx, int by, int W, int H, bool two, int tid)
{
    const bool el = bx < 0, er = W - bx < pitch, et = by < 0, eb = H - by < rows;
    if (el || er) {
        for (int r = tid; r < rows; r += NT) {
            uint8_t *row = buf + r * pitch;
            if (el) {
                row[-1 - bx] = row[1 - bx];
                if (two && -2 - bx >= 0) row[-2 - bx] = row[2 - bx];
            }
            if (er) {
                const int c = W - bx;
                if (c >= 2) {
                    row[c] = row[c - 2];
                    if (two && c + 1 < pitch && c >= 3) row[c + 1] = row[c - 3];
                }
            }
        }
    }
    if (el || er || et || eb) __syncthreads();
    if (et || eb) {
        uint32_t *w = reinterpret_cast<uint32_t *>(buf);
        const int pw = pitch >> 2;
        for (int i = tid; i < pw; i += NT) {
            if (et) {
                w[(-1 - by) * pw + i] = w[(1 - by) * pw + i];
                if (two && -2 - by >= 0) w[(-2 - by) * pw + i] = w[(2 - by) * pw + i];
            }
            if (eb) {
                const int r = H - by;
                if (r >= 2) {
                    w[r * pw + i] = w[(r - 2) * pw + i];
                    if (two && r + 1 < rows && r >= 3) w[(r + 1) * pw + i] = w[(r - 3) * pw + i];
                }
            }
        }
        __syncthreads();
    }
}

// ---- phase 1: level K -> level K+1 in shared memory --------------------------------------------------------------
// Buffer geometry: column j of buffer K+1 <-> column 2j + 2*left(K+1) - left(K) ... of buffer K; with word column c of
// buffer K (4 pixels 4u..4u+3) producing outputs 2u, 2u+1 = columns 2c + dj, 2c + dj + 1 of buffer K+1,
// dj = left(K+1) - left(K)/2; row i of buffer K+1 is centred on row 2i + 2 of buffer K (rows 2i .. 2i+4).
template <int NL, int TH, int K, int NT>
__device__ __forceinline__ void pyr_down_level(uint8_t *smem, int tid)
{
    using G = FG<NL, TH>;
    constexpr int pw = G::pitch(K) / 4, R1 = G::rows(K + 1), P1 = G::pitch(K + 1);
    constexpr int dj = G::left(K + 1) - G::left(K) / 2;
    constexpr int nbands = NT / pw, OB = (R1 + nbands - 1) / nbands;          // output rows per band
    static_assert(nbands >= 1 && 2 * OB * G::pitch(K) <= 12288 + (K + 2 < NL ? G::pitch(K + 2) * G::rows(K + 2) : 0) + G::pitch(K + 1) * G::rows(K + 1),
                  "partial last band would read past the shared-memory allocation");
    const int band = tid / pw, c = tid - band * pw;
    const int i0 = band * OB;
    if (i0 >= R1) return;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(smem + G::offset(K)) + (2 * i0) * pw;
    uint8_t *dst = smem + G::offset(K + 1) + i0 * P1;
    const int cm = max(c - 1, 0), cp = min(c + 1, pw - 1);
    const int j = 2 * c + dj;
    const bool jok = j >= 0 && j + 1 < P1;
    int pg[5][2];
#pragma unroll
    for (int rr = 0; rr < 2 * OB + 3; ++rr) {
        const uint32_t wp = src[rr * pw + cm], wc = src[rr * pw + c], wn = src[rr * pw + cp];
        const uint32_t sc = __funnelshift_r(wp, wc, 16);                      // bytes x-2 .. x+1
        pg[rr % 5][0] = dp4a_us(sc, kGauss, (int)((wc >> 16) & 0xffu));       // + byte x+2
        pg[rr % 5][1] = dp4a_us(wc, kGauss, (int)(wn & 0xffu));               // centre x+2: bytes x..x+3, + byte x+4
        if (rr >= 4 && (rr & 1) == 0) {
            const int o = (rr - 4) >> 1;
            int v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                v[i] = ((pg[(rr - 4) % 5][i] + pg[rr % 5][i]) + 4 * (pg[(rr - 3) % 5][i] + pg[(rr - 1) % 5][i]) +
                        6 * pg[(rr - 2) % 5][i] + 128) >> 8;
            if (jok && i0 + o < R1) *reinterpret_cast<uint16_t *>(dst + o * P1 + j) = (uint16_t)(v[0] | (v[1] << 8));
        }
    }
}

template <int NL, int TH, int K, int NT>
__device__ __forceinline__ void chain(uint8_t *smem, const FusedLevels &L, int tid)
{
    using G = FG<NL, TH>;
    const LevelGeom &lv = L.lv[K];
    const int bx = (int)blockIdx.x * G::own_w(K) - G::left(K), by = (int)blockIdx.y * G::own_h(K) - G::need(K);
    fix_border<NT>(smem + G::offset(K), G::pitch(K), G::rows(K), bx, by, lv.w, lv.h, K + 1 < NL, tid);
    if constexpr (K + 1 < NL) {
        pyr_down_level<NL, TH, K, NT>(smem, tid);
        __syncthreads();
        chain<NL, TH, K + 1, NT>(smem, L, tid);
    }
}

// ---- phase 2: outputs ---------------------------------------------------------------------------------------------
// Scharr of 4 columns x kScharrRows rows of level K's own region (work item `it` of the level) -> HBM
template <int NL, int TH, int K>
__device__ __forceinline__ void scharr_item(const uint8_t *smem, const FusedLevels &L, int plane, int it)
{
    using G = FG<NL, TH>;
    constexpr int ncg = G::sch_cols(K), pw = G::pitch(K) / 4, RB = kScharrRows;
    const LevelGeom &lv = L.lv[K];
    const int W = lv.w, H = lv.h;
    const int band = it / ncg, cg = it - band * ncg;
    const int x = (int)blockIdx.x * G::own_w(K) + 4 * cg, y0 = (int)blockIdx.y * G::own_h(K) + band * RB;
    if (x >= W || y0 >= H) return;                    // (columns [w, pitch) were zeroed at creation and are never written)
    const uint32_t *src = reinterpret_cast<const uint32_t *>(smem + G::offset(K)) + (G::need(K) - 1 + band * RB) * pw + G::left(K) / 4 + cg;
    uint4 *dst = reinterpret_cast<uint4 *>(lv.dxy + (size_t)plane * lv.plane_px + (size_t)y0 * lv.pitch + x);
    const int dst_row_u4 = lv.pitch / 4;
    const int nvalid = W - x;                         // >= 4 except in the group that straddles the right edge
    int hd[3][4], hs[3][4];
#pragma unroll
    for (int rr = 0; rr < RB + 2; ++rr) {
        const uint32_t wp = src[rr * pw - 1], wc = src[rr * pw], wn = src[rr * pw + 1];
        const uint32_t sa = __funnelshift_r(wp, wc, 24), sb = __funnelshift_r(wc, wn, 8);
        int *d = hd[rr % 3], *s = hs[rr % 3];
        d[0] = dp4a_us(sa, kDiffLo, 0);   d[1] = dp4a_us(sa, kDiffHi, 0);
        d[2] = dp4a_us(sb, kDiffLo, 0);   d[3] = dp4a_us(sb, kDiffHi, 0);
        s[0] = dp4a_us(sa, kSmoothLo, 0); s[1] = dp4a_us(sa, kSmoothHi, 0);
        s[2] = dp4a_us(sb, kSmoothLo, 0); s[3] = dp4a_us(sb, kSmoothHi, 0);
        if (rr >= 2) {
            const int *dm = hd[(rr - 2) % 3], *d0 = hd[(rr - 1) % 3], *dp = hd[rr % 3];
            const int *sm = hs[(rr - 2) % 3], *sp = hs[rr % 3];
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int dx = 3 * (dm[i] + dp[i]) + 10 * d0[i], dy = sp[i] - sm[i];
                o[i] = __byte_perm((uint32_t)dx, (uint32_t)dy, 0x5410);
            }
            if (y0 + rr - 2 < H) {
                if (nvalid >= 4) {
                    dst[(size_t)(rr - 2) * dst_row_u4] = make_uint4(o[0], o[1], o[2], o[3]);
                } else {                                  // pad columns stay zero
                    dst[(size_t)(rr - 2) * dst_row_u4] = make_uint4(o[0], nvalid > 1 ? o[1] : 0u, nvalid > 2 ? o[2] : 0u, 0u);
                }
            }
        }
    }
}

template <int NL, int TH, int K>
__device__ __forceinline__ void scharr_dispatch(const uint8_t *smem, const FusedLevels &L, int plane, int it)
{
    using G = FG<NL, TH>;
    if (it < G::sch_begin(K) + G::sch_items(K)) {
        scharr_item<NL, TH, K>(smem, L, plane, it - G::sch_begin(K));
    } else if constexpr (K + 1 < NL) {
        scharr_dispatch<NL, TH, K + 1>(smem, L, plane, it);
    }
}

// one aligned word of level K's own region (K >= 1): shared memory -> HBM
template <int NL, int TH, int K>
__device__ __forceinline__ void copy_item(const uint8_t *smem, const FusedLevels &L, int plane, int it)
{
    using G = FG<NL, TH>;
    constexpr int nw = G::own_w(K) / 4;
    const LevelGeom &lv = L.lv[K];
    const int r = it / nw, c = it - r * nw;
    const int x = (int)blockIdx.x * G::own_w(K) + 4 * c, y = (int)blockIdx.y * G::own_h(K) + r;
    if (x >= lv.w || y >= lv.h) return;
    const uint32_t v = *reinterpret_cast<const uint32_t *>(smem + G::offset(K) + (G::need(K) + r) * G::pitch(K) + G::left(K) + 4 * c);
    uint8_t *g = lv.img + (size_t)plane * lv.iplane_px + (size_t)y * lv.ipitch + x;
    if (x + 4 <= lv.w) {
        *reinterpret_cast<uint32_t *>(g) = v;
    } else {                                          // the frame ends inside this word: the bytes past it belong to the apron
        for (int k = 0; k < lv.w - x; ++k) g[k] = (uint8_t)(v >> (8 * k));
    }
}

template <int NL, int TH, int K>
__device__ __forceinline__ void copy_dispatch(const uint8_t *smem, const FusedLevels &L, int plane, int it)
{
    using G = FG<NL, TH>;
    if constexpr (K < NL) {
        if (it < G::cp_begin(K) + G::cp_items(K)) copy_item<NL, TH, K>(smem, L, plane, it - G::cp_begin(K));
        else copy_dispatch<NL, TH, K + 1>(smem, L, plane, it);
    }
}

// REFLECT_101 apron of the levels that carry one, scattered from the CTA's own pixels (pyr_common.cuh)
struct ApronTable {
    ApronWork lv[4];
};

template <int NL, int TH>
__device__ __forceinline__ void apron_table(ApronTable &T, const FusedLevels &L, int K)
{
    const LevelGeom &lv = L.lv[K];
    const int own_w = kTopW << (NL - 1 - K), own_h = TH << (NL - 1 - K);
    apron_work_fill(T.lv[K], lv.w, lv.h, lv.apron, (int)blockIdx.x * own_w, (int)blockIdx.y * own_h, own_w, own_h);
}

template <int NL, int TH, int K, int NT>
__device__ __forceinline__ void apron_level(const uint8_t *smem, const ApronTable &T, const FusedLevels &L, int plane, int tid)
{
    using G = FG<NL, TH>;
    const LevelGeom &lv = L.lv[K];
    const int ox = (int)blockIdx.x * G::own_w(K), oy = (int)blockIdx.y * G::own_h(K);
    constexpr int hv = G::need(K) >= 3 ? 3 : G::need(K);                          // columns beside the own region that hold true pixels
    apron_scatter(T.lv[K], smem + G::offset(K), G::pitch(K), ox - G::left(K), oy - G::need(K), ox, G::own_w(K), hv, lv.w, lv.h, lv.apron,
                  lv.img + (size_t)plane * lv.iplane_px, lv.ipitch, tid, NT);
}

template <int NL, int TH, int K, int NT>
__device__ __forceinline__ void apron_all(const uint8_t *smem, const ApronTable &T, const FusedLevels &L, int plane, int tid)
{
    if constexpr (K < NL) {
        apron_level<NL, TH, K, NT>(smem, T, L, plane, tid);
        apron_all<NL, TH, K + 1, NT>(smem, T, L, plane, tid);
    }
}

#ifndef PSNW_FUSED_MINB
#define PSNW_FUSED_MINB 4            // resident CTAs per SM of the 256-thread variant (half of it with 512 threads)
#endif
template <int NL, int TH, int NT>
__global__ void __launch_bounds__(NT, NT <= 256 ? PSNW_FUSED_MINB : PSNW_FUSED_MINB / 2)
pyr_fused_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ FusedLevels L, const __grid_constant__ PlaneList planes)
{
    using G = FG<NL, TH>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ ApronTable aprons;
    const int tid = threadIdx.x;
    const int plane = planes.plane[blockIdx.z];
    const uint32_t bar_a = smem_u32(&bar);

    PSNW_STAMP(0);
    // Programmatic dependent launch.  A launch that depends on its predecessors waits first and only then lets its own
    // dependents start: a dependent with late_wait therefore never runs beside anything older than this launch's wait.
    if (L.late_wait) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) mbar_init(bar_a, 1);
    __syncthreads();
    if (!L.late_wait) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (tid == 0) {
        mbar_expect_tx(bar_a, G::pitch(0) * G::box_h() * G::n_box());
#pragma unroll
        for (int b = 0; b < G::n_box(); ++b)
            tma_load_3d(smem_u32(smem + G::offset(0) + b * G::pitch(0) * G::box_h()), &map0, (int)blockIdx.x * G::own_w(0) - G::left(0),
                        (int)blockIdx.y * G::own_h(0) - G::need(0) + b * G::box_h(), plane, bar_a);
    }
    // (while the box is in flight) the apron work lists of this CTA; the barriers of the chain publish them
    if (tid >= 32 && tid < 32 + NL) apron_table<NL, TH>(aprons, L, tid - 32);      // (level 0 of a frame has no apron: empty list)
    mbar_wait(bar_a, 0);
    PSNW_STAMP(1);

    // phase 1: every level of the region into shared memory
    chain<NL, TH, 0, NT>(smem, L, tid);
    PSNW_STAMP(2);

    // phase 2: outputs (no barriers; the rotation of the thread index spreads the last, partial round of Scharr items)
    constexpr int n_sch = G::sch_begin(NL - 1) + G::sch_items(NL - 1);
    for (int it = tid; it < n_sch; it += NT) scharr_dispatch<NL, TH, 0>(smem, L, plane, it);
    PSNW_STAMP(3);
    constexpr int n_cp = G::cp_begin(NL - 1) + G::cp_items(NL - 1);
    const int rot = (tid + NT - (n_sch % NT)) % NT;       // threads that did one Scharr item less go first
    for (int it = rot; it < n_cp; it += NT) copy_dispatch<NL, TH, 1>(smem, L, plane, it);
    PSNW_STAMP(4);
    apron_all<NL, TH, 0, NT>(smem, aprons, L, plane, tid);
    PSNW_STAMP(5);
    if (L.late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int NL, int TH, int NT>
int launch_fused_t(const CUtensorMap *map0, const FusedLevels &L, const PlaneList &planes, cudaStream_t stream)
{
    using G = FG<NL, TH>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((L.lv[0].w + G::own_w(0) - 1) / G::own_w(0), (L.lv[0].h + G::own_h(0) - 1) / G::own_h(0), planes.n);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = G::bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pyr_fused_kernel<NL, TH, NT>, *map0, L, planes) == cudaSuccess ? 1 : -1;
}

template <int NL, int TH, int NT>
int configure_t()
{
    return cudaFuncSetAttribute(pyr_fused_kernel<NL, TH, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FG<NL, TH>::bytes()) == cudaSuccess ? 0 : -1;
}

}  // namespace

// region variants: index 0 / 1 / 2 = 64 / 128 / 256 level-0 rows (4 fused levels), half of that with 3
void fused_geometry(int nl, int variant, int *box_w, int *box_h, int *own_w, int *own_h)
{
#define PSNW_FG(NL_, TH_) { using G = FG<NL_, TH_>; *box_w = G::pitch(0); *box_h = G::box_h(); *own_w = G::own_w(0); *own_h = G::own_h(0); }
    if (nl == 4) { if (variant == 0) PSNW_FG(4, 8) else if (variant == 1) PSNW_FG(4, 16) else PSNW_FG(4, 32) }
    else         { if (variant == 0) PSNW_FG(3, 8) else if (variant == 1) PSNW_FG(3, 16) else PSNW_FG(3, 32) }
#undef PSNW_FG
}

int configure_fused_kernels()
{
    return configure_t<4, 8, 256>() | configure_t<4, 16, 256>() | configure_t<4, 32, 256>() | configure_t<3, 8, 256>() | configure_t<3, 16, 256>() |
           configure_t<3, 32, 256>() | configure_t<4, 8, 512>() | configure_t<4, 16, 512>() | configure_t<4, 32, 512>() | configure_t<3, 8, 512>() |
           configure_t<3, 16, 512>() | configure_t<3, 32, 512>();
}

// levels base .. base+nl-1 (nl = 3 or 4) of the planes in one launch, aprons included, from level `base` already in HBM
// (base = 0: the frame itself); `map0` is the tensor map of level `base` with the box of region variant `variant`
// (fused_geometry).  threads per CTA: 256 (four CTAs per SM: large launches of small regions) or 512 (two).  (1024 threads
// and one CTA per SM were measured for the single-wave launch: no gain, the phases of a lone CTA are latency-bound.)
int launch_pyr_fused(const PyrDesc &pyr, int base, int nl, int variant, int threads, const CUtensorMap *map0, const PlaneList &planes, int late_wait,
                     cudaStream_t stream)
{
    FusedLevels L;
    for (int k = 0; k < 4; ++k) L.lv[k] = pyr.lv[base + (k < nl ? k : nl - 1)];
    L.base = base;
    L.late_wait = late_wait;
#define PSNW_LAUNCH(NL_, TH_)                                                                               \
    return threads <= 256 ? launch_fused_t<NL_, TH_, 256>(map0, L, planes, stream) : launch_fused_t<NL_, TH_, 512>(map0, L, planes, stream)
    if (nl == 4) {
        if (variant == 0) { PSNW_LAUNCH(4, 8); }
        if (variant == 1) { PSNW_LAUNCH(4, 16); }
        PSNW_LAUNCH(4, 32);
    }
    if (variant == 0) { PSNW_LAUNCH(3, 8); }
    if (variant == 1) { PSNW_LAUNCH(3, 16); }
    PSNW_LAUNCH(3, 32);
#undef PSNW_LAUNCH
}

}  // namespace psnw
