// K5: Shi-Tomasi feature detection inside detection boxes.
//
// Replaces m_detector->detect(gray, keypoints, mask) at psn_where/PSNWhere_Tracker2D.cpp:741 (mask = one
// detection box, :735-742) with north_star's detector: cv::goodFeaturesToTrack semantics (SURVEY.md
// Appendix A.4) -- cornerMinEigenVal(blockSize 3, Sobel 3) -> max under the mask -> THRESH_TOZERO at
// quality * max -> 3x3 non-maximum suppression -> strongest-first greedy min-distance selection
// (ties: higher raster address first, as cv2 4.13's greaterThanPtr).
//
//   min_eig_kernel        dense fp32 map of level 0, HBM-bound: reads N bytes, writes 4N.  Float operations
//                         are kept in the oracle's order (sequential 3x3 sums, unfused) so the map is
//                         bit-identical to oracle/klt_oracle.c::psnwo_min_eig_map.
//   detect_select_kernel  one CTA per target: max-reduce over the ROI, candidate compaction into shared memory,
//                         bitonic sort, then one warp walks the sorted list 32 candidates at a time (distance test
//                         against the accepted list + in-order ballot resolution).  Lists that do not fit take
//                         max_per_target rounds of block-wide arg-max over a global-memory list instead.
#include <float.h>

#include "common.cuh"

namespace psnw {

namespace {

// min_eig_kernel: a warp owns a strip of 120 columns and slides down kEigRows rows.  Every lane holds 4 adjacent columns as
// one packed word per source row (lanes 0 and 31 are halo lanes: they compute products, no outputs); the words left and right
// come from the neighbouring lanes by shuffle.  Sobel taps are dp4a dot products on the packed bytes (horizontal difference
// (-1,0,1), horizontal smooth (1,2,1)) combined vertically in registers over a ring of three source rows; the three product
// maps (fx*fx, fx*fy, fy*fy) of the last three rows stay in registers for the lane's 4 columns + 1 on each side (shuffled in),
// and every output pixel adds its 3x3 products in the oracle's order (rows, then columns, starting from 0).
// Borders: Sobel reads REFLECT_101 source pixels; the box filter reads REFLECT_101 PRODUCTS -- the product at the mirrored
// pixel.  Computed from mirrored sources, a gradient just outside the frame equals the one at the mirrored pixel except for
// the sign of its normal component, so the products are identical except for the sign of fx*fy: it is flipped back for
// product rows / columns outside the frame (bit-identical, zeros included).
constexpr int kEigWarps = 4, kEigRows = 8, kEigStripW = 120;

__device__ __forceinline__ int refl_clamp(int p, int n)
{
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);
}

__device__ __forceinline__ int dp4a_us_d(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// pixels c .. c+3 of a row as one word, REFLECT_101 outside [0, w)
__device__ __forceinline__ uint32_t eig_row_word(const uint8_t *__restrict__ row, int c, int w)
{
    if (c >= 0 && c + 4 <= w) return __ldg(reinterpret_cast<const uint32_t *>(row + c));
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v |= (uint32_t)__ldg(row + refl_clamp(c + k, w)) << (8 * k);
    return v;
}

__global__ void __launch_bounds__(kEigWarps * 32)
min_eig_kernel(const uint8_t *__restrict__ img_base, int w, int h, int pitch, int plane_px, const EigList list)
{
    const uint8_t *__restrict__ img = img_base + (size_t)list.plane[blockIdx.z] * plane_px;
    float *__restrict__ eig = list.eig[blockIdx.z];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * kEigStripW + 4 * (lane - 1);            // first of the lane's 4 columns
    const int y0 = (blockIdx.y * kEigWarps + warp) * kEigRows;            // first output row of the warp
    if (y0 >= h) return;
    constexpr int kDiffLo = 0x000100FF, kDiffHi = 0x0100FF00;             // (-1,0,1,0), (0,-1,0,1)
    constexpr int kSmLo = 0x00010201, kSmHi = 0x01020100;                 // ( 1,2,1,0), (0, 1,2,1)
    const float scale = (float)(1.0 / (255.0 * 4.0 * 3.0));
    const bool out_lane = lane >= 1 && lane <= 30 && x < w;
    // columns x-1 .. x+4 of the product window: those outside the frame carry a flipped fx*fy
    bool col_out[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) col_out[i] = x - 1 + i < 0 || x - 1 + i >= w;

    int hd[3][4], hs[3][4];               // horizontal taps of the last three source rows
    float pxx[3][6], pxy[3][6], pyy[3][6];   // products of the last three product rows, columns x-1 .. x+4
    // source rows y0-2 .. y0+kEigRows+1 -> product rows y0-1 .. y0+kEigRows -> output rows y0 .. y0+kEigRows-1
#pragma unroll
    for (int rr = 0; rr < kEigRows + 4; ++rr) {
        const int sy = y0 - 2 + rr;
        const uint32_t wc = eig_row_word(img + (size_t)refl_clamp(sy, h) * pitch, x, w);
        const uint32_t wp = __shfl_up_sync(0xffffffffu, wc, 1), wn = __shfl_down_sync(0xffffffffu, wc, 1);
        const uint32_t sa = __funnelshift_r(wp, wc, 24), sb = __funnelshift_r(wc, wn, 8);     // bytes x-1 .. x+2, x+1 .. x+4
        int *d = hd[rr % 3], *sm = hs[rr % 3];
        d[0] = dp4a_us_d(sa, kDiffLo, 0);  d[1] = dp4a_us_d(sa, kDiffHi, 0);
        d[2] = dp4a_us_d(sb, kDiffLo, 0);  d[3] = dp4a_us_d(sb, kDiffHi, 0);
        sm[0] = dp4a_us_d(sa, kSmLo, 0);   sm[1] = dp4a_us_d(sa, kSmHi, 0);
        sm[2] = dp4a_us_d(sb, kSmLo, 0);   sm[3] = dp4a_us_d(sb, kSmHi, 0);
        if (rr < 2) continue;
        // product row py = sy - 1 from the source rows rr-2, rr-1, rr
        const int py = sy - 1;
        const bool row_out = py < 0 || py >= h;
        float *qxx = pxx[rr % 3], *qxy = pxy[rr % 3], *qyy = pyy[rr % 3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gx = hd[(rr - 2) % 3][i] + 2 * hd[(rr - 1) % 3][i] + hd[rr % 3][i];
            const int gy = hs[rr % 3][i] - hs[(rr - 2) % 3][i];
            const float fx = __fmul_rn((float)gx, scale), fy = __fmul_rn((float)gy, scale);
            const float xy = __fmul_rn(fx, fy);
            qxx[i + 1] = __fmul_rn(fx, fx);
            qxy[i + 1] = (row_out != col_out[i + 1]) ? -xy : xy;
            qyy[i + 1] = __fmul_rn(fy, fy);
        }
        // columns x-1 and x+4 from the neighbouring lanes (their products carry their own column flags; the row flag is shared)
        qxx[0] = __shfl_up_sync(0xffffffffu, qxx[4], 1);   qxx[5] = __shfl_down_sync(0xffffffffu, qxx[1], 1);
        qxy[0] = __shfl_up_sync(0xffffffffu, qxy[4], 1);   qxy[5] = __shfl_down_sync(0xffffffffu, qxy[1], 1);
        qyy[0] = __shfl_up_sync(0xffffffffu, qyy[4], 1);   qyy[5] = __shfl_down_sync(0xffffffffu, qyy[1], 1);
        if (rr < 4) continue;
        // output row oy = py - 1: product rows rr-2, rr-1, rr
        const int oy = py - 1;
        if (!out_lane || oy >= h) continue;
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    a = __fadd_rn(a, pxx[(rr - 2 + j) % 3][i + k]);
                    b = __fadd_rn(b, pxy[(rr - 2 + j) % 3][i + k]);
                    c = __fadd_rn(c, pyy[(rr - 2 + j) % 3][i + k]);
                }
            a = __fmul_rn(a, 0.5f);
            c = __fmul_rn(c, 0.5f);
            const float dd = __fsub_rn(a, c);
            o[i] = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(b, b))));
        }
        float *dst = eig + (size_t)oy * pitch + x;
        if (x + 4 <= w) {
            *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < w) dst[i] = o[i];
        }
    }
}

constexpr int kSelThreads = 512;
constexpr float kUnusedSlot = -1.0e6f;
constexpr int kMaxAccepted = 1024;        // accepted-corner list kept in shared memory by the sorted path
constexpr int kHistBins = 2048;           // radix-select histogram: 11 bits of the score per pass

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// Radix-select step over a histogram of kHistBins counts in shared memory: the kept set is the largest set of TOP bins whose
// total stays <= budget.  Returns (through sel[0], sel[1]) the lowest kept bin (kHistBins if none fits) and the total of the
// kept bins.  Every thread owns 4 adjacent bins; suffix sums across threads by a reverse scan (shuffles + one array).
__device__ __forceinline__ void top_bins(const int *hist, int budget, int *sel, int *scratch, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int c0 = hist[4 * tid], c1 = hist[4 * tid + 1], c2 = hist[4 * tid + 2], c3 = hist[4 * tid + 3];
    const int mine = c0 + c1 + c2 + c3;
    int incl = mine;                                   // sum over this lane and the lanes above it in the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
    }
    if (lane == 0) scratch[warp] = incl;               // total of the warp
    if (tid == 0) { sel[0] = kHistBins; sel[1] = 0; }
    __syncthreads();
    int above = incl - mine;                           // every bin above this thread's four
    for (int w = warp + 1; w < kSelThreads / 32; ++w) above += scratch[w];
    // suffix(i) = total of the bins >= i grows as i falls: bin i is kept iff suffix(i) <= budget, the kept set is an upper set
    const int s3 = above + c3, s2 = s3 + c2, s1 = s2 + c1, s0 = s1 + c0;
    int lo = kHistBins;
    if (s3 <= budget) lo = 4 * tid + 3;
    if (s2 <= budget) lo = 4 * tid + 2;
    if (s1 <= budget) lo = 4 * tid + 1;
    if (s0 <= budget) lo = 4 * tid;
    if (lo < kHistBins) atomicMin(&sel[0], lo);
    __syncthreads();
    const int L = sel[0];
    if (L >= 4 * tid && L < 4 * tid + 4) sel[1] = L == 4 * tid ? s0 : L == 4 * tid + 1 ? s1 : L == 4 * tid + 2 ? s2 : s3;
    __syncthreads();
}

// key = (float bits of the score << 32) | raster index: scores are positive, so unsigned order == (score, address)
// order and "largest key first" is exactly cv2's greaterThanPtr (score descending, ties: higher address first).
__global__ void __launch_bounds__(kSelThreads, 2)
detect_select_kernel(int w, int h, int pitch, const DetectTarget *__restrict__ targets,
                     int max_per_target, float quality, float min_dist, unsigned long long *__restrict__ cand_all,
                     int smem_cap, int tile_cap)
{
    extern __shared__ __align__(16) unsigned long long s_list[];      // smem_cap keys (power of two), then the ROI tile
    __shared__ unsigned long long s_red[kSelThreads / 32];
    __shared__ float s_maxw[kSelThreads / 32];
    __shared__ __align__(16) unsigned s_acc[kMaxAccepted + 4];   // accepted corners (x | y << 16), padded with far-away sentinels
    __shared__ int s_nc, s_nout;
    __shared__ __align__(16) unsigned s_near[32];                // chunk-local "too close to an earlier lane" masks of the walk
    __shared__ unsigned s_alive[kSelThreads / 32];               // per-warp survivor masks of the chunk
    __shared__ __align__(16) unsigned long long s_top[kSelThreads];   // the strongest <= 512 candidates of a long list (radix select)
    __shared__ int s_hist[kHistBins];
    __shared__ int s_sel[4];                                      // {lowest kept bin, kept count above the bin, compaction cursor, -}

    const DetectTarget T = targets[blockIdx.x];
    const float *__restrict__ eig = T.eig;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kSelThreads / 32;
    float2 *my_out = T.out;
    float *tile = reinterpret_cast<float *>(s_list + smem_cap);

    if (T.x1 <= T.x0 || T.y1 <= T.y0 || max_per_target <= 0) {
        if (tid == 0) *T.count = 0;
        for (int i = tid; i < max_per_target; i += kSelThreads) my_out[i] = make_float2(kUnusedSlot, kUnusedSlot);
        return;
    }
    if (tid == 0) { s_nc = 0; s_nout = 0; }
    for (int i = tid; i < kMaxAccepted + 4; i += kSelThreads) s_acc[i] = 0xffffffffu;

    // 0. ROI plus a one-pixel ring (inside the frame) into shared memory: many loads in flight instead of one
    //    dependent load per loop trip.  ROIs that do not fit read the map through L1 instead.
    const int tx0 = max(T.x0 - 1, 0), ty0 = max(T.y0 - 1, 0), tx1 = min(T.x1 + 1, w), ty1 = min(T.y1 + 1, h);
    const int tw = tx1 - tx0, th = ty1 - ty0;
    const bool tiled = tw * th <= tile_cap;
    float m = -FLT_MAX;                                          // 1. max of the map under the mask ...
    if (tiled) {
        // ... taken while the tile is loaded: the threads walk the tile linearly (every lane busy whatever its width)
        int y = tid / tw, x = tid - y * tw;
        const int sy = kSelThreads / tw, sx = kSelThreads - sy * tw;
#pragma unroll 4
        for (int i = tid; i < tw * th; i += kSelThreads) {
            const float v = eig[(size_t)(ty0 + y) * pitch + tx0 + x];
            tile[i] = v;
            if (tx0 + x >= T.x0 && tx0 + x < T.x1 && ty0 + y >= T.y0 && ty0 + y < T.y1) m = fmaxf(m, v);
            x += sx; y += sy;
            if (x >= tw) { x -= tw; y++; }
        }
    }
    const float *mp = tiled ? tile - (ty0 * tw + tx0) : eig;     // map pointer such that mp[y * ms + x] is pixel (x, y)
    const int ms = tiled ? tw : pitch;
    if (!tiled) {
        // linear walk over the ROI (every lane busy whatever its width), four independent loads in flight per thread
        const int rw = T.x1 - T.x0, npx = rw * (T.y1 - T.y0);
        const int sy = kSelThreads / rw, sx = kSelThreads - sy * rw;
        int y = tid / rw, x = tid - y * rw;
        float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = tid; i < npx; i += 4 * kSelThreads) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + u * kSelThreads < npx) m4[u] = fmaxf(m4[u], eig[(size_t)(T.y0 + y) * pitch + T.x0 + x]);
                x += sx; y += sy;
                if (x >= rw) { x -= rw; y++; }
            }
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_maxw[warp] = m;
    __syncthreads();
    m = s_maxw[0];
#pragma unroll
    for (int i = 1; i < nwarps; ++i) m = fmaxf(m, s_maxw[i]);
    const float thr = __fmul_rn(m, quality);

    // 2. candidates: above threshold, 3x3 local maximum of the thresholded map, not on the frame border
    unsigned long long *glist = cand_all + T.cand_begin;
    const int cx0 = max(T.x0, 1), cx1 = min(T.x1, w - 1), cy0 = max(T.y0, 1), cy1 = min(T.y1, h - 1);
    const int cw = max(cx1 - cx0, 1);
    // A thread tests a run of 4 adjacent pixels: 3 rows x 6 columns loaded once serve the four 3x3 neighbourhoods.
    //   cv2: THRESH_TOZERO at thr, then "equal to its 3x3 dilation" -- with thr >= 0 a neighbour at or below thr can not
    //   exceed a pixel above it, so the neighbours need no thresholding; a negative threshold (a ROI without any positive
    //   response) keeps the literal rule.
    const int cw4 = (cw + 3) >> 2, nruns = (cx1 > cx0 && cy1 > cy0) ? cw4 * (cy1 - cy0) : 0;
    const int rsy = kSelThreads / cw4, rsx = kSelThreads - rsy * cw4;
    const int xlim = (tiled ? tx1 : w) - 1;                       // last column that may be read
    const bool literal = !(thr >= 0.f);
    // Pass 0 keeps the keys in shared memory only; the global list (for the rounds of 3b) is written by a second pass in
    // the rare case that they do not fit -- the usual target sends nothing to HBM here.
    int nc = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int ry = tid / cw4, rx = tid - ry * cw4;
        // uniform trip count: the list slots of a warp's candidates are reserved with one atomic per warp and step
        for (int i0 = 0; i0 < nruns; i0 += kSelThreads, rx += rsx, ry += rsy) {
            if (rx >= cw4) { rx -= cw4; ry++; }
            const int x = cx0 + 4 * rx, y = cy0 + ry;
            unsigned mine = 0u;                                   // bit k: pixel x + k is a candidate
            float v4[4] = {0.f, 0.f, 0.f, 0.f};
            if (i0 + tid < nruns) {
                const float *p = mp + y * ms;
                float t[6], c[6], b[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int col = min(x - 1 + k, xlim);
                    t[k] = p[col - ms]; c[k] = p[col]; b[k] = p[col + ms];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) v4[k] = c[k + 1];
                if (literal) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        t[k] = t[k] > thr ? t[k] : 0.f; c[k] = c[k] > thr ? c[k] : 0.f; b[k] = b[k] > thr ? b[k] : 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float nb = fmaxf(fmaxf(fmaxf(t[k], t[k + 1]), fmaxf(t[k + 2], c[k])),
                                           fmaxf(fmaxf(b[k], b[k + 1]), fmaxf(b[k + 2], c[k + 2])));
                    if (x + k < cx1 && v4[k] > thr && !(nb > v4[k])) mine |= 1u << k;
                }
            }
            unsigned f[4];
            int before = 0, total = 0;                            // candidates of lower lanes / lower k; of the whole warp
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[k] = __ballot_sync(0xffffffffu, (mine >> k) & 1u);
                total += __popc(f[k]);
            }
            if (total) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_nc, total);
                pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((mine >> k) & 1u) {
                        const int slot = pos + before + __popc(f[k] & ((1u << lane) - 1u));
                        const unsigned long long key = ((unsigned long long)__float_as_uint(v4[k]) << 32) | (unsigned)(y * w + x + k);
                        if (pass == 0) { if (slot < smem_cap) s_list[slot] = key; }
                        else if (slot < T.cand_cap) glist[slot] = key;
                    }
                    before += __popc(f[k]);
                }
            }
        }
        __syncthreads();
        nc = min(s_nc, T.cand_cap);
        if (pass == 1 || (nc <= smem_cap && max_per_target <= kMaxAccepted)) break;      // uniform: the sorted path serves it
        __syncthreads();
        if (tid == 0) s_nc = 0;
        __syncthreads();
    }
    const float md2 = __fmul_rn(min_dist, min_dist);
    const bool suppress = min_dist >= 1.f;
    int n_out = 0;

    if (nc <= smem_cap && max_per_target <= kMaxAccepted) {
        // ---- 3a. sorted path: bitonic sort (descending) in shared memory, then one warp walks the sorted list in
        //          chunks of 32: test against the accepted list, resolve the chunk in order with ballots ----
        // A long list (the large ROIs of a 1080p / 4K rig: thousands of local maxima) is first cut to its strongest <= 512
        // keys by a two-pass radix select on the score bits (11 + 11 bits, histograms in shared memory): the greedy rule only
        // ever looks at stronger candidates, so the walk over that prefix takes exactly the decisions of the walk over the
        // whole list; if it ends short of max_per_target with candidates left out, the whole list is sorted after all.
        unsigned long long *LST = s_list;
        int m = nc;
        const bool try_top = nc > kSelThreads && 2 * max_per_target <= 3 * kSelThreads / 2;
        for (int attempt = try_top ? 0 : 1; attempt < 2; ++attempt) {
        if (attempt == 0) {
            int budget = kSelThreads, base_bits = 0;         // keys whose score bits >= base_bits are kept
            for (int pass = 0; pass < 2; ++pass) {
                const int shift = pass ? 10 : 21;
                for (int i = tid; i < kHistBins; i += kSelThreads) s_hist[i] = 0;
                __syncthreads();
                for (int i = tid; i < nc; i += kSelThreads) {
                    const unsigned sb = (unsigned)(s_list[i] >> 32);
                    // pass 1 counts only the keys of the bin pass 0 could not take whole
                    if (pass == 0 || (int)(sb >> 21) == base_bits - 1) atomicAdd(&s_hist[(sb >> shift) & (kHistBins - 1)], 1);
                }
                __syncthreads();
                top_bins(s_hist, budget, s_sel, reinterpret_cast<int *>(s_red), tid);
                const int lo = s_sel[0], kept = s_sel[1];
                __syncthreads();
                if (pass == 0) { base_bits = lo; budget -= kept; }        // bins >= lo whole; bin lo - 1 goes to pass 1
                else           { base_bits = lo < kHistBins ? (((base_bits - 1) << 11) | lo) : (base_bits << 11); }
                if (pass == 0 && lo == 0) { base_bits = 0; break; }       // (cannot happen with nc > 512: kept for safety)
            }
            // base_bits is now a 22-bit prefix threshold (score bits >> 10), or an 11-bit one shifted when pass 1 kept nothing
            if (tid == 0) s_sel[2] = 0;
            __syncthreads();
            for (int i0 = 0; i0 < nc; i0 += kSelThreads) {
                const int i = i0 + tid;
                const unsigned long long key = i < nc ? s_list[i] : 0ull;
                const bool keep = i < nc && (int)((unsigned)(key >> 32) >> 10) >= base_bits;
                const unsigned found = __ballot_sync(0xffffffffu, keep);
                if (found) {
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(&s_sel[2], __popc(found));
                    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(found & ((1u << lane) - 1u));
                    if (keep && pos < kSelThreads) s_top[pos] = key;
                }
            }
            __syncthreads();
            m = min(s_sel[2], kSelThreads);
            LST = s_top;
            if (m < max_per_target && m < nc) continue;       // the prefix cannot fill the quota: sort everything
        } else {
            LST = s_list;
            m = nc;
            n_out = 0;
            __syncthreads();
        }
        int n2 = kSelThreads;
        while (n2 < m) n2 <<= 1;
        for (int i = m + tid; i < n2; i += kSelThreads) LST[i] = 0ull;
        __syncthreads();
        // each warp owns a contiguous segment of n2 / nwarps keys: compare distances inside a segment only need
        // __syncwarp; the few stages that cross segments take a block barrier (6 instead of ~60)
        const int seg = n2 / nwarps, sbase = warp * seg;
        if (n2 == kSelThreads) {
            // the usual case, one key per thread: the key stays in a register, partners inside the warp are exchanged with
            // shuffles (35 of the 45 stages), partners in other warps through shared memory
            unsigned long long key = LST[tid];
            for (int k = 2; k <= n2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    unsigned long long other;
                    if (j >= 32) {
                        LST[tid] = key;
                        __syncthreads();
                        other = LST[tid ^ j];
                        __syncthreads();
                    } else {
                        other = __shfl_xor_sync(0xffffffffu, key, j);
                    }
                    // descending blocks keep the larger key at the lower index of a pair, ascending ones the smaller
                    const bool keep_max = ((tid & j) == 0) == ((tid & k) == 0);
                    key = keep_max ? (key > other ? key : other) : (key < other ? key : other);
                }
            LST[tid] = key;
        } else
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= seg) {
                    __syncthreads();
                    for (int i = tid; i < n2; i += kSelThreads) {
                        const int l = i ^ j;
                        if (l > i) {
                            const unsigned long long a = LST[i], b = LST[l];
                            if (((i & k) == 0) ? (a < b) : (a > b)) { LST[i] = b; LST[l] = a; }
                        }
                    }
                    __syncthreads();
                } else {
                    for (int t = lane; t < seg; t += 32) {
                        const int i = sbase + t, l = i ^ j;
                        if (l > i) {
                            const unsigned long long a = LST[i], b = LST[l];
                            if (((i & k) == 0) ? (a < b) : (a > b)) { LST[i] = b; LST[l] = a; }
                        }
                    }
                    __syncwarp();
                }
            }
        __syncthreads();
        {
            // The CTA walks the sorted list 32 candidates at a time.  (a) drop candidates closer than min_dist to a corner
            // accepted in an earlier chunk: every warp tests the chunk against ITS slice of the accepted list (4 corners per
            // shared-memory load) and the survivor masks are AND-ed -- the list grows to max_per_target corners and one warp
            // walking all of it kept the other 15 (half an SM's registers) at a barrier for 40 % of the kernel at a quota of
            // 256; (b) meanwhile warp 1 builds, per lane, the bit mask of EARLIER lanes of the chunk that are too close;
            // (c) warp 0 resolves the chunk: the in-order greedy rule is a 32-step bit recurrence that all lanes evaluate
            // identically -- lane j is accepted iff it survived (a) and none of its close predecessors was accepted.
            // Coordinates are integers: distances are exact in int32 and d2 < min_dist^2 <=> d2 < ceil(min_dist^2).
            const int md2i = (int)fminf(ceilf(md2), 2.0e9f);
            if (tid == 0) s_nout = 0;
            __syncthreads();
            int n_acc = 0;
            for (int base = 0; base < m && n_acc < max_per_target; base += 32) {
                const unsigned long long key = (base + lane < m) ? LST[base + lane] : 0ull;
                const unsigned idx = (unsigned)key;
                const int xi = (int)(idx % (unsigned)w), yi = (int)(idx / (unsigned)w);
                const unsigned xy = (unsigned)xi | ((unsigned)yi << 16);
                bool alive = key != 0ull;
                if (suppress) {
                    for (int a = 4 * warp; a < n_acc; a += 4 * nwarps) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(&s_acc[a]);
                        const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int dx = xi - (int)(qq[u] & 0xffffu), dy = yi - (int)(qq[u] >> 16);   // |d| < 32768: no overflow
                            if (a + u < n_acc && dx * dx + dy * dy < md2i) alive = false;
                        }
                    }
                }
                if (warp == 1) {
                    unsigned near = 0u;
                    if (suppress) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const unsigned q = __shfl_sync(0xffffffffu, xy, i);
                            const int dx = xi - (int)(q & 0xffffu), dy = yi - (int)(q >> 16);
                            if (i < lane && dx * dx + dy * dy < md2i) near |= 1u << i;
                        }
                    }
                    s_near[lane] = near;
                }
                const unsigned mine = __ballot_sync(0xffffffffu, alive);
                if (lane == 0) s_alive[warp] = mine;
                __syncthreads();
                if (warp == 0) {
                    unsigned pre = 0xffffffffu;
#pragma unroll
                    for (int i = 0; i < nwarps; ++i) pre &= s_alive[i];
                    unsigned acc = 0u;
                    int room = max_per_target - n_acc;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const uint4 nq = *reinterpret_cast<const uint4 *>(&s_near[4 * j4]);     // broadcast reads
                        const unsigned nj4[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = 4 * j4 + u;
                            if (room > 0 && ((pre >> j) & 1u) && (nj4[u] & acc) == 0u) { acc |= 1u << j; --room; }
                        }
                    }
                    if ((acc >> lane) & 1u) {
                        const int slot = n_acc + __popc(acc & ((1u << lane) - 1u));
                        s_acc[slot] = xy;
                        my_out[slot] = make_float2((float)xi, (float)yi);
                    }
                    if (lane == 0) s_nout = n_acc + __popc(acc);
                }
                __syncthreads();
                n_acc = s_nout;
            }
        }
        __syncthreads();
        n_out = s_nout;
        if (attempt == 0 && n_out < max_per_target && m < nc) continue;      // prefix exhausted short of the quota: whole list
        break;
        }
    } else {
        // ---- 3b. any size: max_per_target rounds of block-wide arg-max over the candidate list in global memory,
        //          killing the candidates within min_dist of the last accepted corner on the way ----
        unsigned long long last = 0;
        float lx = 0.f, ly = 0.f;
        while (n_out < max_per_target) {
            unsigned long long best = 0;
            for (int c = tid; c < nc; c += kSelThreads) {
                unsigned long long key = glist[c];
                if (key == 0) continue;
                if (last != 0) {
                    bool kill = (key == last);
                    if (suppress && !kill) {
                        const unsigned idx = (unsigned)key;
                        const float dx = (float)(int)(idx % (unsigned)w) - lx, dy = (float)(int)(idx / (unsigned)w) - ly;
                        kill = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < md2;
                    }
                    if (kill) { glist[c] = 0; continue; }
                }
                best = key > best ? key : best;
            }
            best = warp_max_u64(best);
            if (lane == 0) s_red[warp] = best;
            __syncthreads();
            best = s_red[0];
#pragma unroll
            for (int i = 1; i < nwarps; ++i) best = s_red[i] > best ? s_red[i] : best;
            __syncthreads();
            if (best == 0) break;
            const unsigned idx = (unsigned)best;
            lx = (float)(int)(idx % (unsigned)w);
            ly = (float)(int)(idx / (unsigned)w);
            if (tid == 0) my_out[n_out] = make_float2(lx, ly);
            last = best;
            ++n_out;
        }
    }
    if (tid == 0) *T.count = n_out;
    // unused slots get a far-away sentinel: fed to psnw_klt_track they fail the bounds test at once (status 0)
    for (int i = n_out + tid; i < max_per_target; i += kSelThreads) my_out[i] = make_float2(kUnusedSlot, kUnusedSlot);
}

}  // namespace

int launch_min_eig(const LevelGeom &l0, const EigList &list, cudaStream_t stream)
{
    if (list.n <= 0) return 0;
    dim3 grid((l0.w + kEigStripW - 1) / kEigStripW, (l0.h + kEigWarps * kEigRows - 1) / (kEigWarps * kEigRows), list.n);
    min_eig_kernel<<<grid, kEigWarps * 32, 0, stream>>>(l0.img, l0.w, l0.h, l0.pitch, l0.plane_px, list);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int configure_detect_kernels()
{
    return cudaFuncSetAttribute(detect_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8 + 16384 * 4) == cudaSuccess ? 0 : -1;
}

int launch_detect_select(int w, int h, int pitch, const DetectTarget *targets, int n_targets,
                         int max_per_target, float quality, float min_dist, unsigned long long *cand,
                         int max_roi_area, int max_tile_px, cudaStream_t stream)
{
    if (n_targets <= 0) return 0;
    // shared-memory candidate capacity: local maxima of a 3x3 NMS are at most ~1/4 of the ROI pixels; larger lists
    // (or > kMaxAccepted corners per target) take the global-memory path inside the kernel
    int cap = 1024;
    while (cap < max_roi_area / 3 && cap < 8192) cap <<= 1;
    const int tile_cap = max_tile_px <= 16384 ? max_tile_px : 0;          // ROI tile in shared memory when it fits (64 KB)
    const size_t smem = (size_t)cap * sizeof(unsigned long long) + (size_t)tile_cap * sizeof(float);
    detect_select_kernel<<<n_targets, kSelThreads, smem, stream>>>(w, h, pitch, targets, max_per_target, quality,
                                                                   min_dist, cand, cap, tile_cap);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
