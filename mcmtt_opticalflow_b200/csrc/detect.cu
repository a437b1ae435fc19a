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
//   detect_select_kernel  one CTA per target: max-reduce over the ROI, candidate compaction, then
//                         max_per_target rounds of block-wide arg-max + distance suppression in shared memory
//                         (equivalent to sort + sequential greedy filter, without the sort).
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
min_eig_kernel(const uint8_t *__restrict__ img, int w, int h, int pitch, float *__restrict__ eig)
{
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

constexpr int kSelThreads = 256;
constexpr float kUnusedSlot = -1.0e6f;
constexpr int kSmemCand = 4096;

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

__global__ void __launch_bounds__(kSelThreads)
detect_select_kernel(const float *__restrict__ eig, int w, int h, int pitch, const DetectTarget *__restrict__ targets,
                     int max_per_target, float quality, float min_dist, unsigned long long *__restrict__ cand_all,
                     float2 *__restrict__ out, int *__restrict__ counts)
{
    __shared__ unsigned long long s_list[kSmemCand];
    __shared__ unsigned long long s_red[kSelThreads / 32];
    __shared__ float s_maxw[kSelThreads / 32];
    __shared__ int s_nc;

    const DetectTarget T = targets[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kSelThreads / 32;
    float2 *my_out = out + (size_t)blockIdx.x * max_per_target;

    if (T.x1 <= T.x0 || T.y1 <= T.y0 || max_per_target <= 0) {
        if (tid == 0) counts[blockIdx.x] = 0;
        for (int i = tid; i < max_per_target; i += kSelThreads) my_out[i] = make_float2(kUnusedSlot, kUnusedSlot);
        return;
    }

    // 1. max of the map under the mask
    float m = -FLT_MAX;
    for (int y = T.y0 + warp; y < T.y1; y += nwarps)
        for (int x = T.x0 + lane; x < T.x1; x += 32) m = fmaxf(m, eig[(size_t)y * pitch + x]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_maxw[warp] = m;
    if (tid == 0) s_nc = 0;
    __syncthreads();
    m = s_maxw[0];
#pragma unroll
    for (int i = 1; i < nwarps; ++i) m = fmaxf(m, s_maxw[i]);
    const float thr = __fmul_rn(m, quality);

    // 2. candidates: above threshold, 3x3 local maximum of the thresholded map, not on the frame border
    unsigned long long *glist = cand_all + T.cand_begin;
    const int cx0 = max(T.x0, 1), cx1 = min(T.x1, w - 1), cy0 = max(T.y0, 1), cy1 = min(T.y1, h - 1);
    for (int y = cy0 + warp; y < cy1; y += nwarps)
        for (int x = cx0 + lane; x < cx1; x += 32) {
            const float *p = eig + (size_t)y * pitch + x;
            const float v = p[0];
            if (!(v > thr)) continue;
            bool is_max = true;
#pragma unroll
            for (int j = -1; j <= 1; ++j)
#pragma unroll
                for (int i = -1; i <= 1; ++i) {
                    float u = p[j * pitch + i];
                    u = (u > thr) ? u : 0.f;
                    is_max = is_max && !(u > v);
                }
            if (is_max) {
                const int pos = atomicAdd(&s_nc, 1);
                if (pos < T.cand_cap)
                    glist[pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(y * w + x);
            }
        }
    __syncthreads();
    const int nc = min(s_nc, T.cand_cap);
    unsigned long long *list = glist;
    if (nc <= kSmemCand) {
        for (int c = tid; c < nc; c += kSelThreads) s_list[c] = glist[c];
        list = s_list;
    }
    __syncthreads();

    // 3. strongest-first greedy selection with min-distance suppression
    const float md2 = __fmul_rn(min_dist, min_dist);
    const bool suppress = min_dist >= 1.f;
    int n_out = 0;
    unsigned long long last = 0;
    float lx = 0.f, ly = 0.f;
    while (n_out < max_per_target) {
        unsigned long long best = 0;
        for (int c = tid; c < nc; c += kSelThreads) {
            unsigned long long key = list[c];
            if (key == 0) continue;
            if (last != 0) {
                bool kill = (key == last);
                if (suppress && !kill) {
                    const unsigned idx = (unsigned)key;
                    const float dx = (float)(int)(idx % (unsigned)w) - lx, dy = (float)(int)(idx / (unsigned)w) - ly;
                    kill = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < md2;
                }
                if (kill) { list[c] = 0; continue; }
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
    if (tid == 0) counts[blockIdx.x] = n_out;
    // unused slots get a far-away sentinel: fed to psnw_klt_track they fail the bounds test at once (status 0)
    for (int i = n_out + tid; i < max_per_target; i += kSelThreads) my_out[i] = make_float2(kUnusedSlot, kUnusedSlot);
}

}  // namespace

int launch_min_eig(const LevelGeom &l0, int plane, float *eig, cudaStream_t stream)
{
    dim3 grid((l0.w + kEigTW - 1) / kEigTW, (l0.h + kEigTH - 1) / kEigTH);
    min_eig_kernel<<<grid, 256, 0, stream>>>(l0.img + (size_t)plane * l0.plane_px, l0.w, l0.h, l0.pitch, eig);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_detect_select(const float *eig, int w, int h, int pitch, const DetectTarget *targets, int n_targets,
                         int max_per_target, float quality, float min_dist, unsigned long long *cand,
                         float2 *out, int *counts, cudaStream_t stream)
{
    if (n_targets <= 0) return 0;
    detect_select_kernel<<<n_targets, kSelThreads, 0, stream>>>(eig, w, h, pitch, targets, max_per_target, quality,
                                                                min_dist, cand, out, counts);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
