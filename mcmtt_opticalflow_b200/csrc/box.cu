// K4: per-target box motion from tracked features == LocalSearchKLT,
// psn_where/PSNWhere_Tracker2D.cpp:452-554 (callers :787 backward, :893 forward); SURVEY.md Appendix B.
//
// One CTA per target.  All comparisons are in double like the reference (PSN_Point2D, PSNWhere_Types.h:18-66);
// the displacement itself is a float subtraction promoted to double (cv::Point2f operator-, :479).
// std::sort + "first strict maximum" (:498-537) is restated without a sort: the estimate on each axis is
// the SMALLEST displacement among those with the largest neighbour count.
#include <algorithm>

#include "common.cuh"

namespace psnw {

namespace {

constexpr int kBoxThreads = 256;
constexpr int kMaxPts = PSNW_MAX_PTS_PER_TARGET;

struct Best {
    int cnt;
    double val;
};

__device__ __forceinline__ Best better(Best a, Best b)
{
    if (b.cnt > a.cnt || (b.cnt == a.cnt && b.cnt > 0 && b.val < a.val)) return b;
    return a;
}

__device__ __forceinline__ Best warp_best(Best v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best w;
        w.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o);
        w.val = __shfl_xor_sync(0xffffffffu, v.val, o);
        v = better(v, w);
    }
    return v;
}

// Shared memory: the MOVING features only, compacted in arbitrary order (neighbour counts, the smallest displacement among
// the largest counts and the inlier test do not depend on the order): displacement x, y as double and the feature index.
// The neighbour count of feature k is split over S adjacent lanes (S = power of two <= 8, chosen so that a typical
// 100-feature target keeps the whole CTA busy) and combined with shuffles; the per-axis winners are reduced with
// shuffles as well -- no serial section.
__global__ void __launch_bounds__(kBoxThreads)
box_update_kernel(const double *__restrict__ boxes_in, const float2 *__restrict__ pre, const float2 *__restrict__ cur,
                  const uint8_t *__restrict__ status, const int *__restrict__ offsets,
                  double *__restrict__ boxes_out, uint8_t *__restrict__ mask, int *__restrict__ counts, int cap)
{
    extern __shared__ __align__(16) unsigned char box_smem[];
    __shared__ int s_n, s_moving, s_inl;
    __shared__ Best s_bx[kBoxThreads / 32], s_by[kBoxThreads / 32];

    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int beg = offsets[t], end = offsets[t + 1];
    const int n_all = min(max(end - beg, 0), cap);
    double *mvx = reinterpret_cast<double *>(box_smem);
    double *mvy = mvx + n_all;
    int *idx = reinterpret_cast<int *>(mvy + n_all);
    const double bx = boxes_in[4 * t], by = boxes_in[4 * t + 1], bw = boxes_in[4 * t + 2], bh = boxes_in[4 * t + 3];

    if (tid == 0) { s_n = 0; s_moving = 0; s_inl = 0; }
    __syncthreads();

    // :477-490  moving features (disparity >= 0.1 px)
    for (int i0 = 0; i0 < n_all; i0 += kBoxThreads) {
        const int i = i0 + tid;
        bool valid = false, m = false;
        double x = 0.0, y = 0.0;
        if (i < n_all) {
            valid = status ? (status[beg + i] != 0) : true;
            const float2 a = pre[beg + i], b = cur[beg + i];
            x = (double)__fsub_rn(b.x, a.x); y = (double)__fsub_rn(b.y, a.y);
            const double d = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
            m = valid && !(d < 0.1);
            mask[beg + i] = 0;
        }
        const unsigned vm = __ballot_sync(0xffffffffu, valid), mm = __ballot_sync(0xffffffffu, m);
        int base = 0;
        if (lane == 0) {
            if (vm) atomicAdd(&s_n, __popc(vm));
            if (mm) base = atomicAdd(&s_moving, __popc(mm));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (m) {
            const int pos = base + __popc(mm & ((1u << lane) - 1u));
            mvx[pos] = x; mvy[pos] = y; idx[pos] = i;
        }
    }
    for (int i = n_all + tid; i < end - beg; i += kBoxThreads) mask[beg + i] = 0;   // points past the capacity never vote
    __syncthreads();
    const int n = s_n, n_moving = s_moving;

    if ((double)n_moving < __dmul_rn((double)n, 0.5)) {          // :493-496 static target: box unchanged, no inliers
        if (tid == 0) {
            boxes_out[4 * t] = bx; boxes_out[4 * t + 1] = by; boxes_out[4 * t + 2] = bw; boxes_out[4 * t + 3] = bh;
            counts[t] = 0;
        }
        return;
    }

    // :502-537  per-axis mode: most neighbours within 0.2 * box.w, ties -> smallest displacement
    const double win = __dmul_rn(bw, 0.2);
    int S = 1;
    while (S < 8 && 2 * S * n_moving <= kBoxThreads) S *= 2;
    const int slice = tid & (S - 1), kstep = kBoxThreads / S;
    Best kx = {0, 0.0}, ky = {0, 0.0};
    for (int k0 = 0; k0 < n_moving; k0 += kstep) {               // uniform trip count: the shuffles below stay convergent
        const int k = k0 + tid / S;
        const bool live = k < n_moving;
        const double vx = live ? mvx[k] : 0.0, vy = live ? mvy[k] : 0.0;
        int cx = 0, cy = 0;
        if (live) {
#pragma unroll 4
            for (int j = slice; j < n_moving; j += S) {
                cx += fabs(__dsub_rn(vx, mvx[j])) < win;
                cy += fabs(__dsub_rn(vy, mvy[j])) < win;
            }
        }
        for (int o = 1; o < S; o <<= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, o);
            cy += __shfl_xor_sync(0xffffffffu, cy, o);
        }
        if (live) {
            kx = better(kx, Best{cx, vx});
            ky = better(ky, Best{cy, vy});
        }
    }
    kx = warp_best(kx); ky = warp_best(ky);
    if (lane == 0) { s_bx[tid >> 5] = kx; s_by[tid >> 5] = ky; }
    __syncthreads();
    Best ax = {0, 0.0}, ay = {0, 0.0};
#pragma unroll
    for (int i = 0; i < kBoxThreads / 32; ++i) { ax = better(ax, s_bx[i]); ay = better(ay, s_by[i]); }
    const double ex = ax.cnt > 0 ? ax.val : 0.0;      // estimatedDisparity starts at (0,0), Types.h:27
    const double ey = ay.cnt > 0 ? ay.val : 0.0;

    // :540-546 inliers among the moving features
    int inl = 0;
    for (int k = tid; k < n_moving; k += kBoxThreads) {
        const double rx = __dsub_rn(mvx[k], ex), ry = __dsub_rn(mvy[k], ey);
        if (sqrt(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry))) < win) { mask[beg + idx[k]] = 1; ++inl; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) inl += __shfl_xor_sync(0xffffffffu, inl, o);
    if (lane == 0 && inl) atomicAdd(&s_inl, inl);
    __syncthreads();
    if (tid == 0) {                                              // :549-553
        boxes_out[4 * t] = __dadd_rn(bx, ex); boxes_out[4 * t + 1] = __dadd_rn(by, ey);
        boxes_out[4 * t + 2] = bw; boxes_out[4 * t + 3] = bh;
        counts[t] = s_inl;
    }
}

// "next" row f-3: tracked points of every tracker inside every detection box -- the counts the reference accumulates in
// featuresInDetectionBox (Tracker2D.cpp:939-944) for its optical-flow majority check (:981-1022), here a by-product on the
// device so that the tracked points need not travel to the host.  PSN_Rect::contain (PSNWhere_Types.h:157) in double.
__global__ void __launch_bounds__(128)
points_in_boxes_kernel(const float2 *__restrict__ pts, const int *__restrict__ offsets, const double *__restrict__ det_boxes,
                       int n_det, int n_trk, int *__restrict__ counts)
{
    __shared__ int s_cnt;
    const int t = blockIdx.x, d = blockIdx.y, tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const double bx = det_boxes[4 * d], by = det_boxes[4 * d + 1], bw = det_boxes[4 * d + 2], bh = det_boxes[4 * d + 3];
    const double x1 = __dadd_rn(bx, bw), y1 = __dadd_rn(by, bh);
    int c = 0;
    for (int i = offsets[t] + tid; i < offsets[t + 1]; i += 128) {
        const float2 p = pts[i];
        const double x = (double)p.x, y = (double)p.y;
        c += (x >= bx && x < x1 && y >= by && y < y1) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (tid == 0) counts[(size_t)d * n_trk + t] = s_cnt;
}

}  // namespace

int launch_points_in_boxes(const float2 *pts, const int *offsets, int n_trk, const double *det_boxes, int n_det, int *counts,
                           cudaStream_t stream)
{
    if (n_trk <= 0 || n_det <= 0) return 0;
    points_in_boxes_kernel<<<dim3(n_trk, n_det), 128, 0, stream>>>(pts, offsets, det_boxes, n_det, n_trk, counts);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// opt-in shared-memory sizes are a per-device attribute: psnw_create calls this for the context's device
int configure_box_kernels()
{
    return cudaFuncSetAttribute(box_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPts * 20 + 16) == cudaSuccess ? 0 : -1;
}

int launch_box_update(const double *boxes_in, int n_targets, const float2 *pre, const float2 *cur,
                      const uint8_t *status, const int *offsets, double *boxes_out, uint8_t *mask,
                      int *counts, int max_pts_per_target, cudaStream_t stream)
{
    if (n_targets <= 0) return 0;
    const int cap = std::min(std::max(max_pts_per_target, 1), kMaxPts);
    const size_t smem = (size_t)cap * 20 + 16;
    box_update_kernel<<<n_targets, kBoxThreads, smem, stream>>>(boxes_in, pre, cur, status, offsets, boxes_out, mask, counts, cap);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
