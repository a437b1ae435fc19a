// K1f: the whole Gaussian pyramid + Scharr derivatives of a frame in ONE launch (up to 4 levels).
//
// The per-level kernel (pyramid.cu) is bandwidth-efficient for large launches but a small rig frame (4 x 768x576 =
// 12 MB) pays four dependent launches of ~3.5 us each for ~2 us of HBM traffic.  Here one CTA owns a 192x64 region of
// level 0 (96x32, 48x16, 24x8 of the next levels) and produces every level for it from a single TMA box of level 0:
// the halo each coarser level needs (1 px for Scharr, 2n+2 px for the pyrDown that feeds a level needing n) is
// recomputed locally -- 22 / 10 / 4 / 1 px at levels 0..3 -- so CTAs never wait for each other.  Intermediate levels
// live in shared memory only (plus their own part written to HBM); HBM traffic is the algorithmic N(6 S_L - 1) bytes, the
// overlap between neighbouring boxes is served by L2.
//
// Arithmetic is the same integer dp4a code as pyramid.cu: bit-exact with cv::pyrDown / calcSharrDeriv, REFLECT_101
// applied at every level's own image border (x = -1, -2, W, W+1 and the same rows) inside shared memory.
#include "common.cuh"
#include "pyr_common.cuh"

namespace psnw {

namespace {

constexpr int kFusedThreads = 512;
constexpr int kTopW = 24, kTopH = 8;                 // region owned at the coarsest fused level

template <int NL> struct FG {
    static __host__ __device__ constexpr int own_w(int k) { return kTopW << (NL - 1 - k); }
    static __host__ __device__ constexpr int own_h(int k) { return kTopH << (NL - 1 - k); }
    static __host__ __device__ constexpr int need(int k) { return k >= NL - 1 ? 1 : 2 * need(k + 1) + 2; }
    // columns kept left (= right) of the own region: >= need, multiple of 4 (words), of 16 at level 0 (TMA start)
    static __host__ __device__ constexpr int left(int k) { return k == 0 ? (need(0) + 15) / 16 * 16 : (need(k) + 3) / 4 * 4; }
    static __host__ __device__ constexpr int pitch(int k) { return own_w(k) + 2 * left(k); }
    static __host__ __device__ constexpr int rows(int k) { return own_h(k) + 2 * need(k); }
    static __host__ __device__ constexpr int offset(int k) { return k == 0 ? 128 : offset(k - 1) + pitch(k - 1) * rows(k - 1); }
    static __host__ __device__ constexpr int bytes() { return offset(NL - 1) + pitch(NL - 1) * rows(NL - 1) + 16; }
    static __host__ __device__ constexpr int rb(int k) { return own_h(k) / 8; }      // Scharr rows per item: 8 bands per level
};

struct FusedLevels {
    LevelGeom lv[4];
};

// REFLECT_101 at the border of a level held in shared memory: x = -1 (and -2 when `two`), x = W (and W+1), same rows.
// Columns first, then whole rows, so that the corners come out right.  Ends with a barrier.
__device__ __forceinline__ void fix_border(uint8_t *buf, int pitch, int rows, int bx, int by, int W, int H, bool two, int tid)
{
    const bool el = bx < 0, er = W - bx < pitch, et = by < 0, eb = H - by < rows;
    if (el || er) {
        for (int r = tid; r < rows; r += kFusedThreads) {
            uint8_t *row = buf + r * pitch;
            if (el) {
                row[-1 - bx] = row[1 - bx];
                if (two && -2 - bx >= 0) row[-2 - bx] = row[2 - bx];
            }
            if (er) {
                const int c = W - bx;
                row[c] = row[c - 2];
                if (two && c + 1 < pitch) row[c + 1] = row[c - 3];
            }
        }
    }
    if (el || er || et || eb) __syncthreads();
    if (et || eb) {
        uint32_t *w = reinterpret_cast<uint32_t *>(buf);
        const int pw = pitch >> 2;
        for (int i = tid; i < pw; i += kFusedThreads) {
            if (et) {
                w[(-1 - by) * pw + i] = w[(1 - by) * pw + i];
                if (two && -2 - by >= 0) w[(-2 - by) * pw + i] = w[(2 - by) * pw + i];
            }
            if (eb) {
                const int r = H - by;
                w[r * pw + i] = w[(r - 2) * pw + i];
                if (two && r + 1 < rows) w[(r + 1) * pw + i] = w[(r - 3) * pw + i];
            }
        }
        __syncthreads();
    }
}

// Scharr of 4 columns x RB rows starting at (x, y0); `src` points at the word holding columns x..x+3 of buffer row y0-1.
template <int RB>
__device__ __forceinline__ void scharr_item(const uint32_t *src, int pw, uint4 *dst, int dst_row_u4, int x, int y0, int W, int H)
{
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
                o[i] = (x + i < W) ? __byte_perm((uint32_t)dx, (uint32_t)dy, 0x5410) : 0u;   // pad columns stay zero
            }
            if (y0 + rr - 2 < H) dst[(size_t)(rr - 2) * dst_row_u4] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// one level of the fused chain: border fix-up, Scharr of the own region -> HBM, pyrDown into the next level's buffer
template <int NL, int K>
__device__ __forceinline__ void fused_level(uint8_t *smem, const FusedLevels &L, int plane, int tid)
{
    using G = FG<NL>;
    const LevelGeom &lv = L.lv[K];
    const int W = lv.w, H = lv.h;
    const int ox = (int)blockIdx.x * G::own_w(K), oy = (int)blockIdx.y * G::own_h(K);
    const int bx = ox - G::left(K), by = oy - G::need(K);
    uint8_t *buf = smem + G::offset(K);
    constexpr int pitch = G::pitch(K), pw = G::pitch(K) / 4;
    fix_border(buf, pitch, G::rows(K), bx, by, W, H, K + 1 < NL, tid);

    // ---- Scharr of the own region -> HBM ----
    {
        constexpr int RB = G::rb(K), ncg = G::own_w(K) / 4;
        if (tid < ncg * 8) {
            const int band = tid / ncg, cg = tid - band * ncg;
            const int x = ox + 4 * cg, y0 = oy + band * RB;
            if (x < lv.pitch && x < ((W + 3) & ~3) && y0 < H) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(buf) + (y0 - 1 - by) * pw + ((x - bx) >> 2);
                uint4 *dst = reinterpret_cast<uint4 *>(lv.dxy + (size_t)plane * lv.plane_px + (size_t)y0 * lv.pitch + x);
                scharr_item<RB>(src, pw, dst, lv.pitch / 4, x, y0, W, H);
            }
        }
    }

    // ---- pyrDown: level K+1 over its own region plus the halo the coarser levels need -> shared memory (+ HBM) ----
    if constexpr (K + 1 < NL) {
        const LevelGeom &nx = L.lv[K + 1];
        const int W1 = nx.w, H1 = nx.h;
        const int ox1 = (int)blockIdx.x * G::own_w(K + 1), oy1 = (int)blockIdx.y * G::own_h(K + 1);
        constexpr int n1 = G::need(K + 1), w1 = G::own_w(K + 1), h1 = G::own_h(K + 1);
        const int bx1 = ox1 - G::left(K + 1), by1 = oy1 - n1;
        uint8_t *buf1 = smem + G::offset(K + 1);
        constexpr int pitch1 = G::pitch(K + 1);
        const int xlo = max(ox1 - n1, 0), xhi = min(ox1 + w1 + n1, W1);          // outputs wanted: [xlo, xhi) x [ylo, yhi)
        const int ylo = max(oy1 - n1, 0), yhi = min(oy1 + h1 + n1, H1);
        const int u0 = (ox1 - n1) >> 1;                                          // lane unit u -> outputs 2u, 2u+1
        constexpr int nu = (w1 + 2 * n1) / 2 + 1, nyi = (h1 + 2 * n1 + 3) / 4;
        const uint32_t *srcw = reinterpret_cast<const uint32_t *>(buf);
        for (int it = tid; it < nu * nyi; it += kFusedThreads) {
            const int yi = it / nu, u = u0 + (it - yi * nu);
            const int X = 2 * u, Y0 = (oy1 - n1) + 4 * yi;
            if (X + 1 < xlo || X >= xhi || Y0 >= yhi || Y0 + 3 < ylo) continue;
            const int wi = (4 * u - bx) >> 2;                                    // word of input columns 4u..4u+3
            const int r0 = 2 * Y0 - 2 - by;                                      // buffer row of input row 2*Y0 - 2
            int pg[5][2];
#pragma unroll
            for (int rr = 0; rr < 11; ++rr) {
                const int r = min(max(r0 + rr, 0), G::rows(K) - 1);
                const uint32_t wp = srcw[r * pw + wi - 1], wc = srcw[r * pw + wi], wn = srcw[r * pw + wi + 1];
                const uint32_t sc = __funnelshift_r(wp, wc, 16);
                pg[rr % 5][0] = dp4a_us(sc, kGauss, (int)((wc >> 16) & 0xffu));
                pg[rr % 5][1] = dp4a_us(wc, kGauss, (int)(wn & 0xffu));
                if (rr >= 4 && (rr & 1) == 0) {
                    const int Y = Y0 + ((rr - 4) >> 1);
                    int v[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        v[i] = ((pg[(rr - 4) % 5][i] + pg[rr % 5][i]) + 4 * (pg[(rr - 3) % 5][i] + pg[(rr - 1) % 5][i]) +
                                6 * pg[(rr - 2) % 5][i] + 128) >> 8;
                    if (Y >= ylo && Y < yhi) {
                        const bool a0 = X >= xlo && X < xhi, a1 = X + 1 >= xlo && X + 1 < xhi;
                        uint8_t *d = buf1 + (Y - by1) * pitch1 + (X - bx1);
                        if (a0) d[0] = (uint8_t)v[0];
                        if (a1) d[1] = (uint8_t)v[1];
                        if (Y >= oy1 && Y < oy1 + h1 && X >= ox1 && X < ox1 + w1) {          // own part of level K+1 -> HBM
                            uint8_t *g = nx.img + (size_t)plane * nx.iplane_px + (size_t)Y * nx.ipitch + X;
                            if (a0 && a1) *reinterpret_cast<uint16_t *>(g) = (uint16_t)(v[0] | (v[1] << 8));
                            else if (a0) g[0] = (uint8_t)v[0];
                        }
                    }
                }
            }
        }
        __syncthreads();
        fused_level<NL, K + 1>(smem, L, plane, tid);
    }
}

template <int NL>
__global__ void __launch_bounds__(kFusedThreads, 1)
pyr_fused_kernel(const __grid_constant__ CUtensorMap map0, const FusedLevels L, const PlaneList planes)
{
    using G = FG<NL>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int plane = planes.plane[blockIdx.z];
    const uint32_t bar_a = smem_u32(&bar);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) mbar_init(bar_a, 1);
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        mbar_expect_tx(bar_a, G::pitch(0) * G::rows(0));
        tma_load_3d(smem_u32(smem + G::offset(0)), &map0, (int)blockIdx.x * G::own_w(0) - G::left(0),
                    (int)blockIdx.y * G::own_h(0) - G::need(0), plane, bar_a);
    }
    mbar_wait(bar_a, 0);
    fused_level<NL, 0>(smem, L, plane, tid);
}

template <int NL>
int launch_fused_t(const CUtensorMap *map0, const FusedLevels &L, const PlaneList &planes, cudaStream_t stream)
{
    using G = FG<NL>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((L.lv[0].w + G::own_w(0) - 1) / G::own_w(0), (L.lv[0].h + G::own_h(0) - 1) / G::own_h(0), planes.n);
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = G::bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pyr_fused_kernel<NL>, *map0, L, planes) == cudaSuccess ? 1 : -1;
}

}  // namespace

void fused_box(int nl, int *box_w, int *box_h)
{
    if (nl == 4) { *box_w = FG<4>::pitch(0); *box_h = FG<4>::rows(0); }
    else         { *box_w = FG<3>::pitch(0); *box_h = FG<3>::rows(0); }
}

int configure_fused_kernels()
{
    if (cudaFuncSetAttribute(pyr_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FG<4>::bytes()) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(pyr_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FG<3>::bytes()) != cudaSuccess) return -1;
    return 0;
}

// levels 0 .. nl-1 (nl = 3 or 4) of the planes in one launch; `map0` is the level-0 tensor map with the fused box
int launch_pyr_fused(const PyrDesc &pyr, int nl, const CUtensorMap *map0, const PlaneList &planes, cudaStream_t stream)
{
    FusedLevels L;
    for (int k = 0; k < 4; ++k) L.lv[k] = pyr.lv[k < nl ? k : nl - 1];
    return nl == 4 ? launch_fused_t<4>(map0, L, planes, stream) : launch_fused_t<3>(map0, L, planes, stream);
}

}  // namespace psnw
