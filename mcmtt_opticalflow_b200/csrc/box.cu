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
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (launched with programmatic stream serialisation)
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


// ---- device-resident backward chain (Tracker2D.cpp:763-811) --------------------------------------------------------
// Points of target t live in the padded slots [t * cap, (t + 1) * cap); `live` marks the slots in use (the LK status is
// ignored in the backward pass, :787: the status array handed to box_update_kernel is this liveness).  Dead slots hold a
// far-away sentinel, which the LK kernels reject at once.
constexpr float kDeadSlot = -1.0e6f;

// first step: liveness and target index of every slot from the per-target counts; dead slots get the sentinel
__global__ void chain_init_kernel(float2 *__restrict__ pts, const int *__restrict__ counts, int n_targets, int cap,
                                  uint8_t *__restrict__ live, int *__restrict__ tgt, int *__restrict__ offsets,
                                  int *__restrict__ alive, int *__restrict__ steps_done)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_targets) offsets[i] = i * cap;
    if (i < n_targets) { alive[i] = 1; steps_done[i] = 0; }
    if (i >= n_targets * cap) return;
    const int t = i / cap, k = i - t * cap;
    const bool on = k < min(max(counts[t], 0), cap);
    live[i] = on ? 1 : 0;
    tgt[i] = t;
    if (!on) pts[i] = make_float2(kDeadSlot, kDeadSlot);
}

// after LocalSearchKLT of step k: a target with >= min_inliers inliers records its box and continues with the inliers'
// tracked positions, compacted in index order (:788-807); any other target stops for good (:788 break)
__global__ void __launch_bounds__(256)
chain_compact_kernel(const float2 *__restrict__ tracked, const uint8_t *__restrict__ mask, const int *__restrict__ n_inl,
                     const double *__restrict__ box_step, int cap, int min_inliers, int step,
                     float2 *__restrict__ next_pts, uint8_t *__restrict__ next_live, int *__restrict__ next_count,
                     double *__restrict__ boxes_steps, int *__restrict__ alive, int *__restrict__ steps_done)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool go = alive[t] != 0 && n_inl[t] >= min_inliers;
    if (tid == 0) s_base = 0;
    __syncthreads();
    if (tid == 0) {
        if (go) {
            steps_done[t] = step;
            for (int k = 0; k < 4; ++k) boxes_steps[4 * t + k] = box_step[4 * t + k];
        } else {
            alive[t] = 0;
        }
        next_count[t] = go ? n_inl[t] : 0;
    }
    for (int i0 = 0; i0 < cap; i0 += 256) {
        const int i = i0 + tid;
        const bool in = go && i < cap && mask[(size_t)t * cap + i] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (in) next_pts[(size_t)t * cap + before + __popc(b & ((1u << lane) - 1u))] = tracked[(size_t)t * cap + i];
        __syncthreads();
        if (tid == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += s_warp[w]; s_base += s; }
        __syncthreads();
    }
    const int n = go ? s_base : 0;
    for (int i = tid; i < cap; i += 256) {
        next_live[(size_t)t * cap + i] = i < n ? 1 : 0;
        if (i >= n) next_pts[(size_t)t * cap + i] = make_float2(kDeadSlot, kDeadSlot);
    }
}

// ---- optical-flow majority check on the device ("next" row f-3; Tracker2D.cpp:931-944 and :981-1022) ----------------
// counts[d][t] = tracked points of tracker t inside detection box d, for the trackers that reach the comparison (ok[t]:
// >= 4 status-ok pairs, :889) and whose new box overlaps the detection box (:937); 0 otherwise.
__global__ void __launch_bounds__(128)
gated_points_in_boxes_kernel(const float2 *__restrict__ pts, const int *__restrict__ offsets, const double *__restrict__ trk_boxes,
                             const uint8_t *__restrict__ ok, const double *__restrict__ det_boxes, int n_det, int n_trk,
                             int *__restrict__ counts)
{
    __shared__ int s_cnt;
    const int t = blockIdx.x, d = blockIdx.y, tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const double bx = det_boxes[4 * d], by = det_boxes[4 * d + 1], bw = det_boxes[4 * d + 2], bh = det_boxes[4 * d + 3];
    const double tx = trk_boxes[4 * t], ty = trk_boxes[4 * t + 1], tw = trk_boxes[4 * t + 2], th = trk_boxes[4 * t + 3];
    // PSN_Rect::overlap (PSNWhere_Types.h:160-163), newBox.overlap(detection.box)
    const bool gate = ok[t] != 0 &&
                      (__dsub_rn(fmax(__dadd_rn(tx, tw), __dadd_rn(bx, bw)), fmin(tx, bx)) < __dadd_rn(tw, bw)) &&
                      (__dsub_rn(fmax(__dadd_rn(ty, th), __dadd_rn(by, bh)), fmin(ty, by)) < __dadd_rn(th, bh));
    const double x1 = __dadd_rn(bx, bw), y1 = __dadd_rn(by, bh);
    int c = 0;
    if (gate) {
        for (int i = offsets[t] + tid; i < offsets[t + 1]; i += 128) {
            const float2 p = pts[i];
            const double x = (double)p.x, y = (double)p.y;
            c += (x >= bx && x < x1 && y >= by && y < y1) ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (tid == 0) counts[(size_t)d * n_trk + t] = s_cnt;
}

// The reference walks featuresInDetectionBox[d] -- tracker indices, one entry per point, in tracker order -- with a
// run-length rule (:991-1013): the first run counts all its points, a later run one less (the counter restarts at 0 on
// the run's first point), the LAST run is never compared unless it is the only one; the detection's row of the cost
// matrix is wiped unless the largest counted run exceeds ratio x all points.  One thread per detection restates it on
// the per-tracker counts.
__global__ void majority_kernel(const int *__restrict__ counts, int n_det, int n_trk, double ratio, uint8_t *__restrict__ wipe)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_det) return;
    int total = 0, runs = 0, num_major = 0, num_cur = 0;
    for (int t = 0; t < n_trk; ++t) {
        const int c = counts[(size_t)d * n_trk + t];
        if (c <= 0) continue;
        if (runs > 0) {                          // switch to a new tracker: the run before is compared, the counter restarts
            if (num_cur > num_major) num_major = num_cur;
            num_cur = c - 1;
        } else {
            num_cur = c;
        }
        total += c;
        ++runs;
    }
    if (runs == 1) num_major = num_cur;          // sole tracker, :1008-1011
    wipe[d] = (total > 0 && !((double)num_major > __dmul_rn((double)total, ratio))) ? 1 : 0;
}

}  // namespace

int launch_points_in_boxes(const float2 *pts, const int *offsets, int n_trk, const double *det_boxes, int n_det, int *counts,
                           cudaStream_t stream)
{
    if (n_trk <= 0 || n_det <= 0) return 0;
    points_in_boxes_kernel<<<dim3(n_trk, n_det), 128, 0, stream>>>(pts, offsets, det_boxes, n_det, n_trk, counts);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_chain_init(float2 *pts, const int *counts, int n_targets, int cap, uint8_t *live, int *tgt, int *offsets, int *alive,
                      int *steps_done, cudaStream_t stream)
{
    const int n = std::max(n_targets * cap, n_targets + 1);
    chain_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, counts, n_targets, cap, live, tgt, offsets, alive, steps_done);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_chain_compact(const float2 *tracked, const uint8_t *mask, const int *n_inl, const double *box_step, int n_targets, int cap,
                         int min_inliers, int step, float2 *next_pts, uint8_t *next_live, int *next_count, double *boxes_steps,
                         int *alive, int *steps_done, cudaStream_t stream)
{
    if (n_targets <= 0) return 0;
    chain_compact_kernel<<<n_targets, 256, 0, stream>>>(tracked, mask, n_inl, box_step, cap, min_inliers, step, next_pts, next_live,
                                                         next_count, boxes_steps, alive, steps_done);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_flow_majority(const float2 *pts, const int *offsets, int n_trk, const double *trk_boxes, const uint8_t *ok,
                         const double *det_boxes, int n_det, double ratio, int *counts, uint8_t *wipe, cudaStream_t stream)
{
    if (n_trk <= 0 || n_det <= 0) return 0;
    gated_points_in_boxes_kernel<<<dim3(n_trk, n_det), 128, 0, stream>>>(pts, offsets, trk_boxes, ok, det_boxes, n_det, n_trk, counts);
    majority_kernel<<<(n_det + 127) / 128, 128, 0, stream>>>(counts, n_det, n_trk, ratio, wipe);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

// opt-in shared-memory sizes are a per-device attribute: psnw_create calls this for the context's device
// ------------------------------------------------------------------------------------------------------------------
// PSNW_BOX_MEDIANFLOW: median displacement + median pairwise-distance ratio (Kalal, Mikolajczyk, Matas: "Forward-Backward
// Error", the estimator north_star names).  Not what the reference runs (LocalSearchKLT above is),
// so it has no reference parity: its oracle is oracle/klt_oracle.c::psnwo_box_medianflow, bit for bit.  Definition in
// include/psnw_klt.h.  One CTA per target; every median is an exact order statistic found by an 8-pass radix select over
// order-preserving 64-bit keys of the doubles in shared memory (no sort).
constexpr int kMfThreads = 256;
constexpr int kMfPairPts = 128;                                // the scale uses at most this many (evenly strided) points
constexpr int kMfMaxKeys = kMfPairPts * (kMfPairPts - 1) / 2;  // 8128 pairs  (> kMaxPts)

__device__ __forceinline__ unsigned long long mf_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double mf_val(unsigned long long k)
{
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// k-th smallest (0-based) of keys[0..n): block-wide radix select, 8 bits per pass from the top.  All threads return it.
__device__ unsigned long long mf_select(const unsigned long long *keys, int n, int k, int *hist, int *ctl, int tid)
{
    unsigned long long prefix = 0;
#pragma unroll 1
    for (int shift = 56; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kMfThreads) {
            const unsigned long long key = keys[i];
            if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(int)((key >> shift) & 255u)], 1);
        }
        __syncthreads();
        // exclusive prefix of the 256 bins (one per thread): warp scan + the warps' totals
        const int c = hist[tid];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) ctl[2 + (tid >> 5)] = incl;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < (tid >> 5); ++w) before += ctl[2 + w];
        const int excl = before + incl - c;
        if (k >= excl && k < excl + c) { ctl[0] = tid; ctl[1] = k - excl; }
        __syncthreads();
        prefix |= (unsigned long long)ctl[0] << shift;
        k = ctl[1];
        __syncthreads();
    }
    return prefix;
}

// median of n keys: the middle one, or the mean of the two middle ones
__device__ double mf_median(const unsigned long long *keys, int n, int *hist, int *ctl, int tid)
{
    const double hi = mf_val(mf_select(keys, n, n >> 1, hist, ctl, tid));
    if (n & 1) return hi;
    const double lo = mf_val(mf_select(keys, n, (n >> 1) - 1, hist, ctl, tid));
    return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

__global__ void __launch_bounds__(kMfThreads)
box_medianflow_kernel(const double *__restrict__ boxes_in, const float2 *__restrict__ pre, const float2 *__restrict__ cur,
                      const uint8_t *__restrict__ status, const int *__restrict__ offsets,
                      double *__restrict__ boxes_out, uint8_t *__restrict__ mask, int *__restrict__ counts, int cap)
{
    extern __shared__ __align__(16) unsigned char box_smem[];
    __shared__ int hist[256], ctl[2 + kMfThreads / 32], s_n, s_pairs, s_inl;
    __shared__ double s_pts[kMfPairPts][4];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(box_smem);              // kMfMaxKeys
    unsigned short *vidx = reinterpret_cast<unsigned short *>(keys + kMfMaxKeys);              // cap
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int beg = offsets[t], end = offsets[t + 1];
    const int n_all = min(max(end - beg, 0), cap);
    const double bx = boxes_in[4 * t], by = boxes_in[4 * t + 1], bw = boxes_in[4 * t + 2], bh = boxes_in[4 * t + 3];
    if (tid == 0) { s_n = 0; s_pairs = 0; s_inl = 0; }
    for (int i = tid; i < end - beg; i += kMfThreads) mask[beg + i] = 0;
    __syncthreads();

    // valid points in index order: chunks of 256 in turn, ballot-ordered inside a chunk
    for (int i0 = 0; i0 < n_all; i0 += kMfThreads) {
        const int i = i0 + tid;
        const bool valid = i < n_all && (status ? status[beg + i] != 0 : true);
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) ctl[2 + (tid >> 5)] = __popc(vm);
        __syncthreads();
        int before = s_n;
        for (int w = 0; w < (tid >> 5); ++w) before += ctl[2 + w];
        if (valid) vidx[before + __popc(vm & ((1u << lane) - 1u))] = (unsigned short)i;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < kMfThreads / 32; ++w) tot += ctl[2 + w]; s_n += tot; }
        __syncthreads();
    }
    const int n = s_n;
    if (n == 0) {
        if (tid == 0) {
            boxes_out[4 * t] = bx; boxes_out[4 * t + 1] = by; boxes_out[4 * t + 2] = bw; boxes_out[4 * t + 3] = bh;
            counts[t] = 0;
        }
        return;
    }
    // median displacement per axis
    for (int j = tid; j < n; j += kMfThreads) {
        const int i = beg + vidx[j];
        keys[j] = mf_key(__dsub_rn((double)cur[i].x, (double)pre[i].x));
    }
    __syncthreads();
    const double mdx = mf_median(keys, n, hist, ctl, tid);
    for (int j = tid; j < n; j += kMfThreads) {
        const int i = beg + vidx[j];
        keys[j] = mf_key(__dsub_rn((double)cur[i].y, (double)pre[i].y));
    }
    __syncthreads();
    const double mdy = mf_median(keys, n, hist, ctl, tid);
    // residuals to the median displacement, their median, the inlier mask
    for (int j = tid; j < n; j += kMfThreads) {
        const int i = beg + vidx[j];
        const double rx = __dsub_rn(__dsub_rn((double)cur[i].x, (double)pre[i].x), mdx);
        const double ry = __dsub_rn(__dsub_rn((double)cur[i].y, (double)pre[i].y), mdy);
        keys[j] = mf_key(sqrt(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry))));
    }
    __syncthreads();
    const double mres = mf_median(keys, n, hist, ctl, tid);
    int inl = 0;
    for (int j = tid; j < n; j += kMfThreads) {
        if (mf_val(keys[j]) <= mres) { mask[beg + vidx[j]] = 1; ++inl; }
    }
    if (inl) atomicAdd(&s_inl, inl);
    __syncthreads();
    // scale: median ratio of pairwise distances over at most kMfPairPts evenly strided valid points
    const int ns = min(n, kMfPairPts);
    if (tid < ns) {
        const int j = n <= kMfPairPts ? tid : (int)(((long long)tid * n) / kMfPairPts);
        const int i = beg + vidx[j];
        s_pts[tid][0] = (double)pre[i].x; s_pts[tid][1] = (double)pre[i].y;
        s_pts[tid][2] = (double)cur[i].x; s_pts[tid][3] = (double)cur[i].y;
    }
    __syncthreads();
    const int n_pairs = ns * (ns - 1) / 2;
    int good = 0;
    for (int p = tid; p < n_pairs; p += kMfThreads) {
        // pair p -> (a, b), a < b, row-major over the strict upper triangle
        int a = (int)((2.0 * ns - 1.0 - sqrt((2.0 * ns - 1.0) * (2.0 * ns - 1.0) - 8.0 * p)) * 0.5);
        while (a * (2 * ns - a - 1) / 2 > p) --a;
        while ((a + 1) * (2 * ns - a - 2) / 2 <= p) ++a;
        const int b = a + 1 + (p - a * (2 * ns - a - 1) / 2);
        const double px = __dsub_rn(s_pts[a][0], s_pts[b][0]), py = __dsub_rn(s_pts[a][1], s_pts[b][1]);
        const double cx = __dsub_rn(s_pts[a][2], s_pts[b][2]), cy = __dsub_rn(s_pts[a][3], s_pts[b][3]);
        const double dp = sqrt(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)));
        const double dc = sqrt(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)));
        if (dp > 0.0) { keys[p] = mf_key(__ddiv_rn(dc, dp)); ++good; }
        else          keys[p] = ~0ull;                        // coincident source points: sorts last, never counted
    }
    if (good) atomicAdd(&s_pairs, good);
    __syncthreads();
    const int np = s_pairs;
    double scale = 1.0;
    if (np > 0) {
        // the np valid keys are the np smallest of the array: their median = order statistics np/2 (and np/2 - 1) of the whole
        const double hi = mf_val(mf_select(keys, n_pairs, np >> 1, hist, ctl, tid));
        scale = (np & 1) ? hi : __dmul_rn(__dadd_rn(mf_val(mf_select(keys, n_pairs, (np >> 1) - 1, hist, ctl, tid)), hi), 0.5);
    }
    if (tid == 0) {
        const double g = __dmul_rn(0.5, __dsub_rn(scale, 1.0));
        boxes_out[4 * t] = __dsub_rn(__dadd_rn(bx, mdx), __dmul_rn(g, bw));
        boxes_out[4 * t + 1] = __dsub_rn(__dadd_rn(by, mdy), __dmul_rn(g, bh));
        boxes_out[4 * t + 2] = __dmul_rn(bw, scale);
        boxes_out[4 * t + 3] = __dmul_rn(bh, scale);
        counts[t] = s_inl;
    }
}

int configure_box_kernels()
{
    if (cudaFuncSetAttribute(box_medianflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMfMaxKeys * 8 + kMaxPts * 2) != cudaSuccess) return -1;
    return cudaFuncSetAttribute(box_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPts * 20 + 16) == cudaSuccess ? 0 : -1;
}

int launch_box_update(const double *boxes_in, int n_targets, const float2 *pre, const float2 *cur,
                      const uint8_t *status, const int *offsets, double *boxes_out, uint8_t *mask,
                      int *counts, int max_pts_per_target, cudaStream_t stream, int mode)
{
    if (n_targets <= 0) return 0;
    const int cap = std::min(std::max(max_pts_per_target, 1), kMaxPts);
    if (mode == PSNW_BOX_MEDIANFLOW) {
        box_medianflow_kernel<<<n_targets, kMfThreads, (size_t)kMfMaxKeys * 8 + (size_t)cap * 2, stream>>>(boxes_in, pre, cur, status, offsets,
                                                                                                         boxes_out, mask, counts, cap);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    const size_t smem = (size_t)cap * 20 + 16;
    // programmatic dependent launch: set up under the tail of the LK launch whose points it reads (the kernel waits first)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_targets);
    cfg.blockDim = dim3(kBoxThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, box_update_kernel, boxes_in, pre, cur, status, offsets, boxes_out, mask, counts, cap) == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
