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

constexpr int kEigTW = 64, kEigTH = 16;
constexpr int kSrcW = kEigTW + 4, kSrcH = kEigTH + 4;       // halo 2
constexpr int kPrW = kEigTW + 2, kPrH = kEigTH + 2;         // halo 1

__device__ __forceinline__ int refl_clamp(int p, int n)
{
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);
}

__global__ void __launch_bounds__(256)
min_eig_kernel(const uint8_t *__restrict__ img_base, int w, int h, int pitch, int plane_px, const EigList list)
{
    const uint8_t *__restrict__ img = img_base + (size_t)list.plane[blockIdx.z] * plane_px;
    float *__restrict__ eig = list.eig[blockIdx.z];
    __shared__ uint8_t src[kSrcH][kSrcW];
    __shared__ float pxx[kPrH][kPrW], pxy[kPrH][kPrW], pyy[kPrH][kPrW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kEigTW, y0 = blockIdx.y * kEigTH;

    for (int i = tid; i < kSrcW * kSrcH; i += 256) {
        const int sy = i / kSrcW, sx = i - sy * kSrcW;
        src[sy][sx] = img[(size_t)refl_clamp(y0 - 2 + sy, h) * pitch + refl_clamp(x0 - 2 + sx, w)];
    }
    __syncthreads();

    const float scale = (float)(1.0 / (255.0 * 4.0 * 3.0));
    for (int i = tid; i < kPrW * kPrH; i += 256) {
        const int py = i / kPrW, px = i - py * kPrW;
        // the covariance maps are image-sized and box-filtered with REFLECT_101: a product outside the frame
        // is the product AT the mirrored pixel
        const int X = refl_clamp(x0 - 1 + px, w), Y = refl_clamp(y0 - 1 + py, h);
        const int sx = min(max(X - (x0 - 2), 1), kSrcW - 2), sy = min(max(Y - (y0 - 2), 1), kSrcH - 2);
        const int gx = (src[sy - 1][sx + 1] + 2 * src[sy][sx + 1] + src[sy + 1][sx + 1]) -
                       (src[sy - 1][sx - 1] + 2 * src[sy][sx - 1] + src[sy + 1][sx - 1]);
        const int gy = (src[sy + 1][sx - 1] + 2 * src[sy + 1][sx] + src[sy + 1][sx + 1]) -
                       (src[sy - 1][sx - 1] + 2 * src[sy - 1][sx] + src[sy - 1][sx + 1]);
        const float fx = __fmul_rn((float)gx, scale), fy = __fmul_rn((float)gy, scale);
        pxx[py][px] = __fmul_rn(fx, fx);
        pxy[py][px] = __fmul_rn(fx, fy);
        pyy[py][px] = __fmul_rn(fy, fy);
    }
    __syncthreads();

    for (int i = tid; i < kEigTW * kEigTH; i += 256) {
        const int oy = i / kEigTW, ox = i - oy * kEigTW;
        const int X = x0 + ox, Y = y0 + oy;
        if (X >= w || Y >= h) continue;
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a = __fadd_rn(a, pxx[oy + j][ox + k]);
                b = __fadd_rn(b, pxy[oy + j][ox + k]);
                c = __fadd_rn(c, pyy[oy + j][ox + k]);
            }
        a = __fmul_rn(a, 0.5f);
        c = __fmul_rn(c, 0.5f);
        const float d = __fsub_rn(a, c);
        eig[(size_t)Y * pitch + X] = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    }
}

constexpr int kSelThreads = 512;
constexpr float kUnusedSlot = -1.0e6f;
constexpr int kMaxAccepted = 1024;        // accepted-corner list kept in shared memory by the sorted path

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// key = (float bits of the score << 32) | raster index: scores are positive, so unsigned order == (score, address)
// order and "largest key first" is exactly cv2's greaterThanPtr (score descending, ties: higher address first).
__global__ void __launch_bounds__(kSelThreads)
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
#pragma unroll 2
        for (int y = T.y0 + warp; y < T.y1; y += nwarps)
            for (int x = T.x0 + lane; x < T.x1; x += 32) m = fmaxf(m, mp[y * ms + x]);
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
    const int cw = max(cx1 - cx0, 1), n_cand_px = (cx1 > cx0 && cy1 > cy0) ? cw * (cy1 - cy0) : 0;
    const int csy = kSelThreads / cw, csx = kSelThreads - csy * cw;
    // Pass 0 keeps the keys in shared memory only; the global list (for the rounds of 3b) is written by a second pass in
    // the rare case that they do not fit -- the usual target sends nothing to HBM here.
    int nc = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int cyy = cy0 + tid / cw, cxx = cx0 + (tid - (tid / cw) * cw);
        // uniform trip count: the list slots of a warp's candidates are reserved with one atomic per warp and step
        for (int i0 = 0; i0 < n_cand_px; i0 += kSelThreads, cxx += csx, cyy += csy) {
            if (cxx >= cx1) { cxx -= cw; cyy++; }
            const int x = cxx, y = cyy;
            bool is_max = false;
            float v = 0.f;
            if (i0 + tid < n_cand_px) {
                const float *p = mp + y * ms + x;
                v = p[0];
                if (v > thr) {
                    is_max = true;
#pragma unroll
                    for (int j = -1; j <= 1; ++j)
#pragma unroll
                        for (int i = -1; i <= 1; ++i) {
                            float u = p[j * ms + i];
                            u = (u > thr) ? u : 0.f;
                            is_max = is_max && !(u > v);
                        }
                }
            }
            const unsigned found = __ballot_sync(0xffffffffu, is_max);
            if (found) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_nc, __popc(found));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(found & ((1u << lane) - 1u));
                if (is_max) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(y * w + x);
                    if (pass == 0) { if (pos < smem_cap) s_list[pos] = key; }
                    else if (pos < T.cand_cap) glist[pos] = key;
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
        int n2 = kSelThreads;
        while (n2 < nc) n2 <<= 1;
        for (int i = nc + tid; i < n2; i += kSelThreads) s_list[i] = 0ull;
        __syncthreads();
        // each warp owns a contiguous segment of n2 / nwarps keys: compare distances inside a segment only need
        // __syncwarp; the few stages that cross segments take a block barrier (6 instead of ~60)
        const int seg = n2 / nwarps, sbase = warp * seg;
        if (n2 == kSelThreads) {
            // the usual case, one key per thread: the key stays in a register, partners inside the warp are exchanged with
            // shuffles (35 of the 45 stages), partners in other warps through shared memory
            unsigned long long key = s_list[tid];
            for (int k = 2; k <= n2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    unsigned long long other;
                    if (j >= 32) {
                        s_list[tid] = key;
                        __syncthreads();
                        other = s_list[tid ^ j];
                        __syncthreads();
                    } else {
                        other = __shfl_xor_sync(0xffffffffu, key, j);
                    }
                    // descending blocks keep the larger key at the lower index of a pair, ascending ones the smaller
                    const bool keep_max = ((tid & j) == 0) == ((tid & k) == 0);
                    key = keep_max ? (key > other ? key : other) : (key < other ? key : other);
                }
            s_list[tid] = key;
        } else
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= seg) {
                    __syncthreads();
                    for (int i = tid; i < n2; i += kSelThreads) {
                        const int l = i ^ j;
                        if (l > i) {
                            const unsigned long long a = s_list[i], b = s_list[l];
                            if (((i & k) == 0) ? (a < b) : (a > b)) { s_list[i] = b; s_list[l] = a; }
                        }
                    }
                    __syncthreads();
                } else {
                    for (int t = lane; t < seg; t += 32) {
                        const int i = sbase + t, l = i ^ j;
                        if (l > i) {
                            const unsigned long long a = s_list[i], b = s_list[l];
                            if (((i & k) == 0) ? (a < b) : (a > b)) { s_list[i] = b; s_list[l] = a; }
                        }
                    }
                    __syncwarp();
                }
            }
        __syncthreads();
        if (warp == 0) {
            // One warp walks the sorted list 32 candidates at a time.  (a) drop candidates closer than min_dist to a
            // corner accepted in an earlier chunk (4 accepted corners per shared-memory load); (b) every lane builds
            // the bit mask of EARLIER lanes of the chunk that are too close; (c) the in-order greedy rule becomes a
            // 32-step bit recurrence that all lanes evaluate identically: lane j is accepted iff it survived (a) and
            // none of its close predecessors was accepted.
            // Coordinates are integers: distances are exact in int32 and d2 < min_dist^2 <=> d2 < ceil(min_dist^2).
            const int md2i = (int)fminf(ceilf(md2), 2.0e9f);
            for (int base = 0; base < nc && n_out < max_per_target; base += 32) {
                const unsigned long long key = (base + lane < nc) ? s_list[base + lane] : 0ull;
                const unsigned idx = (unsigned)key;
                const int xi = (int)(idx % (unsigned)w), yi = (int)(idx / (unsigned)w);
                const unsigned xy = (unsigned)xi | ((unsigned)yi << 16);
                bool alive = key != 0ull;
                unsigned near = 0u;
                if (suppress) {
                    for (int a = 0; a < n_out; a += 4) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(&s_acc[a]);
                        const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int dx = xi - (int)(qq[u] & 0xffffu), dy = yi - (int)(qq[u] >> 16);   // |d| < 32768: no overflow
                            if (a + u < n_out && dx * dx + dy * dy < md2i) alive = false;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const unsigned q = __shfl_sync(0xffffffffu, xy, i);
                        const int dx = xi - (int)(q & 0xffffu), dy = yi - (int)(q >> 16);
                        if (i < lane && dx * dx + dy * dy < md2i) near |= 1u << i;
                    }
                }
                const unsigned pre = __ballot_sync(0xffffffffu, alive);
                s_near[lane] = near;
                __syncwarp();
                unsigned acc = 0u;
                int room = max_per_target - n_out;
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
                    const int slot = n_out + __popc(acc & ((1u << lane) - 1u));
                    s_acc[slot] = xy;
                    my_out[slot] = make_float2((float)xi, (float)yi);
                }
                n_out += __popc(acc);
                __syncwarp();
            }
            if (lane == 0) s_nout = n_out;
        }
        __syncthreads();
        n_out = s_nout;
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
    dim3 grid((l0.w + kEigTW - 1) / kEigTW, (l0.h + kEigTH - 1) / kEigTH, list.n);
    min_eig_kernel<<<grid, 256, 0, stream>>>(l0.img, l0.w, l0.h, l0.pitch, l0.plane_px, list);
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
