// K6 ("next" row f-1): GridFAST, the detector the reference really uses --
// cv::FeatureDetector::create("GridFAST") at psn_where/PSNWhere_Tracker2D.cpp:142, called with the detection-box mask
// at :735-742 -- i.e. OpenCV 2.4 GridAdaptedFeatureDetector(FastFeatureDetector(threshold 10, NMS), 1000, 4 x 4):
// every grid cell is detected as its own sub-image (3 px dead band along cell edges, NMS inside the cell), keypoints
// outside the mask are dropped, each cell keeps its maxTotal/16 = 62 strongest INCLUDING ties of the boundary response.
// The reference then shuffles with rand() and keeps 100 (:752-757); this path keeps the strongest max_per_target
// instead (ties cut in raster order) and returns them in raster order.
//
//   fast_map_kernel         dense u8 map of NMS'd FAST-9/16 scores: reads N bytes, writes N bytes (HBM-bound, once per
//                           pushed frame, shared by all detection boxes of the frame -- the reference re-runs FAST on the
//                           whole frame for every detection).
//   gridfast_select_kernel  one CTA per detection box: per-cell response histograms in shared memory -> per-cell
//                           keepStrongest thresholds -> global top-K threshold -> ordered compaction by block scans.
#include "common.cuh"

namespace psnw {

namespace {

constexpr int kFTW = 64, kFTH = 16;                 // output tile
constexpr int kFSW = kFTW + 8, kFSH = kFTH + 8;     // source tile, halo 4 (radius-3 circle around the halo-1 scores)
constexpr int kFRW = kFTW + 2, kFRH = kFTH + 2;     // raw-score tile, halo 1

__constant__ int c_fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// cornerScore<16> of features2d/fast_score.cpp on the differences d[k] = v - p_k (k = 0..24, wrapping)
__device__ __forceinline__ int fast_corner_score(const int *d, int threshold)
{
    int a0 = threshold;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int a = min(min(d[k + 1], d[k + 2]), d[k + 3]);
        if (a <= a0) continue;
        a = min(a, min(min(d[k + 4], d[k + 5]), min(d[k + 6], min(d[k + 7], d[k + 8]))));
        a0 = max(a0, min(a, d[k]));
        a0 = max(a0, min(a, d[k + 9]));
    }
    int b0 = -a0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int b = max(max(d[k + 1], d[k + 2]), d[k + 3]);
        if (b >= b0) continue;
        b = max(b, max(max(d[k + 4], d[k + 5]), max(d[k + 6], max(d[k + 7], d[k + 8]))));
        b0 = min(b0, max(b, d[k]));
        b0 = min(b0, max(b, d[k + 9]));
    }
    return -b0 - 1;
}

__global__ void __launch_bounds__(256)
fast_map_kernel(const uint8_t *__restrict__ img, int w, int h, int pitch, int threshold, int grid_rows, int grid_cols,
                uint8_t *__restrict__ out)
{
    __shared__ uint8_t src[kFSH][kFSW];
    __shared__ uint8_t raw[kFRH][kFRW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFTW, y0 = blockIdx.y * kFTH;
    for (int i = tid; i < kFSW * kFSH; i += 256) {
        const int sy = i / kFSW, sx = i - sy * kFSW;
        const int gx = min(max(x0 - 4 + sx, 0), w - 1), gy = min(max(y0 - 4 + sy, 0), h - 1);   // clamped: never used by a valid corner
        src[sy][sx] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < kFRW * kFRH; i += 256) {
        const int ry = i / kFRW, rx = i - ry * kFRW;
        const int X = x0 - 1 + rx, Y = y0 - 1 + ry;
        int score = 0;
        if (X >= 0 && Y >= 0 && X < w && Y < h) {
            // grid cell of the pixel and its position inside the cell (the cell is detected as a sub-image)
            int cj = (X * grid_cols) / w, ci = (Y * grid_rows) / h;
            while ((cj * w) / grid_cols > X) --cj;
            while (((cj + 1) * w) / grid_cols <= X) ++cj;
            while ((ci * h) / grid_rows > Y) --ci;
            while (((ci + 1) * h) / grid_rows <= Y) ++ci;
            const int c0 = (cj * w) / grid_cols, c1 = ((cj + 1) * w) / grid_cols, r0 = (ci * h) / grid_rows, r1 = ((ci + 1) * h) / grid_rows;
            if (X >= c0 + 3 && X < c1 - 3 && Y >= r0 + 3 && Y < r1 - 3) {
                const int sx = rx + 3, sy = ry + 3;                      // (X, Y) inside the source tile
                const int v = src[sy][sx];
                unsigned dark = 0, bright = 0;
                int d[25];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    d[k] = v - (int)src[sy + c_fast_dy[k]][sx + c_fast_dx[k]];
                    dark |= (d[k] > threshold ? 1u : 0u) << k;
                    bright |= (d[k] < -threshold ? 1u : 0u) << k;
                }
                // 9 contiguous circle pixels: AND of the mask with its 8 successive circular rotations
                unsigned md = dark | (dark << 16), mb = bright | (bright << 16), rd = md, rb = mb;
#pragma unroll
                for (int s = 1; s <= 8; ++s) { rd &= md >> s; rb &= mb >> s; }
                if (((rd | rb) & 0xffffu) != 0u) {
#pragma unroll
                    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
                    score = fast_corner_score(d, threshold);
                }
            }
        }
        raw[ry][rx] = (uint8_t)min(score, 255);
    }
    __syncthreads();
    for (int i = tid; i < kFTW * kFTH; i += 256) {
        const int oy = i / kFTW, ox = i - oy * kFTW;
        const int X = x0 + ox, Y = y0 + oy;
        if (X >= w || Y >= h) continue;
        const int s = raw[oy + 1][ox + 1];
        bool keep = s > 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
                if (dx != 1 || dy != 1) keep = keep && s > (int)raw[oy + dy][ox + dx];
        out[(size_t)Y * pitch + X] = keep ? (uint8_t)s : (uint8_t)0;
    }
}

constexpr int kGSThreads = 256;
constexpr int kMaxCells = 64;

__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_tot, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kGSThreads / 32; ++i) { if (i < warp) base += warp_tot[i]; tot += warp_tot[i]; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kGSThreads)
gridfast_select_kernel(int w, int h, int pitch, const GridFastTarget *__restrict__ targets, int max_per_target,
                       int grid_rows, int grid_cols, int per_cell)
{
    __shared__ int hist[kMaxCells][64];          // responses are binned by 4 first (coarse pass) ...
    __shared__ int fine[kMaxCells][4];           // ... then the boundary bin of every cell is resolved exactly
    __shared__ int cell_thr[kMaxCells], cell_bin[kMaxCells], cell_above[kMaxCells];
    __shared__ int hist_all[256];
    __shared__ int warp_tot[kGSThreads / 32];
    __shared__ int s_thr, s_room;
    const GridFastTarget T = targets[blockIdx.x];
    const uint8_t *__restrict__ map = T.map;
    const int tid = threadIdx.x, ncell = grid_rows * grid_cols;
    float2 *my_out = T.out;
    const int rw = T.x1 - T.x0, rh = T.y1 - T.y0;
    if (rw <= 0 || rh <= 0 || max_per_target <= 0) {
        if (tid == 0) *T.count = 0;
        for (int i = tid; i < max_per_target; i += kGSThreads) my_out[i] = make_float2(-1.0e6f, -1.0e6f);
        return;
    }
    for (int i = tid; i < kMaxCells * 64; i += kGSThreads) (&hist[0][0])[i] = 0;
    for (int i = tid; i < kMaxCells * 4; i += kGSThreads) (&fine[0][0])[i] = 0;
    for (int i = tid; i < 256; i += kGSThreads) hist_all[i] = 0;
    __syncthreads();
    const int npx = rw * rh;
    auto cell_of = [&](int X, int Y) {
        int cj = (X * grid_cols) / w, ci = (Y * grid_rows) / h;
        while ((cj * w) / grid_cols > X) --cj;
        while (((cj + 1) * w) / grid_cols <= X) ++cj;
        while ((ci * h) / grid_rows > Y) --ci;
        while (((ci + 1) * h) / grid_rows <= Y) ++ci;
        return ci * grid_cols + cj;
    };
    // pass 1: coarse histogram (bins of 4 responses) per cell over the masked keypoints
    for (int p = tid; p < npx; p += kGSThreads) {
        const int y = p / rw, x = p - y * rw, X = T.x0 + x, Y = T.y0 + y;
        const int s = map[(size_t)Y * pitch + X];
        if (s) atomicAdd(&hist[cell_of(X, Y)][s >> 2], 1);
    }
    __syncthreads();
    if (tid < ncell) {                           // coarse bin holding the per_cell-th strongest response of the cell
        int acc = 0, b = 63, n = 0;
        for (int i = 0; i < 64; ++i) n += hist[tid][i];
        if (n > per_cell && per_cell > 0) {
            for (b = 63; b >= 0; --b) { if (acc + hist[tid][b] >= per_cell) break; acc += hist[tid][b]; }
            cell_bin[tid] = b; cell_above[tid] = acc;
        } else {
            cell_bin[tid] = -1; cell_above[tid] = 0;          // keep everything (threshold 1), or nothing when per_cell == 0
        }
        cell_thr[tid] = (n > per_cell && per_cell <= 0) ? 256 : 1;
    }
    __syncthreads();
    // pass 2: exact histogram of the boundary bin
    for (int p = tid; p < npx; p += kGSThreads) {
        const int y = p / rw, x = p - y * rw, X = T.x0 + x, Y = T.y0 + y;
        const int s = map[(size_t)Y * pitch + X];
        if (!s) continue;
        const int c = cell_of(X, Y);
        if (cell_bin[c] == (s >> 2)) atomicAdd(&fine[c][s & 3], 1);
    }
    __syncthreads();
    if (tid < ncell && cell_bin[tid] >= 0) {
        int acc = cell_above[tid], k = 3;
        for (k = 3; k >= 0; --k) { acc += fine[tid][k]; if (acc >= per_cell) break; }
        cell_thr[tid] = cell_bin[tid] * 4 + max(k, 0);          // keepStrongest: every response >= the boundary one
    }
    __syncthreads();
    // pass 3: histogram of everything the cells keep -> global threshold for max_per_target
    for (int p = tid; p < npx; p += kGSThreads) {
        const int y = p / rw, x = p - y * rw, X = T.x0 + x, Y = T.y0 + y;
        const int s = map[(size_t)Y * pitch + X];
        if (s && s >= cell_thr[cell_of(X, Y)]) atomicAdd(&hist_all[s], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int i = 1; i < 256; ++i) total += hist_all[i];
        int thr = 1, room = 1 << 30;
        if (total > max_per_target) {
            int acc = 0;
            for (thr = 255; thr >= 1; --thr) { if (acc + hist_all[thr] >= max_per_target) break; acc += hist_all[thr]; }
            room = max_per_target - acc;
        }
        s_thr = thr; s_room = room;
    }
    __syncthreads();
    const int thr = s_thr, room = s_room;
    // pass 4: ordered compaction in raster order; ties at the global threshold are admitted first come first served
    int n_out = 0, n_tie = 0;
    for (int base = 0; base < npx; base += kGSThreads) {
        const int p = base + tid;
        int s = 0, X = 0, Y = 0;
        if (p < npx) {
            const int y = p / rw, x = p - y * rw;
            X = T.x0 + x; Y = T.y0 + y;
            s = map[(size_t)Y * pitch + X];
            if (s && !(s >= cell_thr[cell_of(X, Y)] && s >= thr)) s = 0;
        }
        const int is_tie = (s != 0 && s == thr) ? 1 : 0;
        int tie_total, out_total;
        const int tie_rank = n_tie + block_exclusive_scan(is_tie, warp_tot, tie_total);
        const int take = (s != 0 && (!is_tie || tie_rank < room)) ? 1 : 0;
        const int pos = n_out + block_exclusive_scan(take, warp_tot, out_total);
        if (take && pos < max_per_target) my_out[pos] = make_float2((float)X, (float)Y);
        n_out += out_total; n_tie += tie_total;
    }
    n_out = min(n_out, max_per_target);
    if (tid == 0) *T.count = n_out;
    for (int i = n_out + tid; i < max_per_target; i += kGSThreads) my_out[i] = make_float2(-1.0e6f, -1.0e6f);
}

}  // namespace

int launch_fast_map(const LevelGeom &l0, int plane, int threshold, int grid_rows, int grid_cols, uint8_t *out, cudaStream_t stream)
{
    dim3 grid((l0.w + kFTW - 1) / kFTW, (l0.h + kFTH - 1) / kFTH);
    fast_map_kernel<<<grid, 256, 0, stream>>>(l0.img + (size_t)plane * l0.plane_px, l0.w, l0.h, l0.pitch, threshold, grid_rows,
                                               grid_cols, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_gridfast_select(int w, int h, int pitch, const GridFastTarget *targets, int n_targets, int max_per_target,
                           int grid_rows, int grid_cols, int per_cell, cudaStream_t stream)
{
    if (n_targets <= 0) return 0;
    if (grid_rows * grid_cols > kMaxCells) return -2;
    gridfast_select_kernel<<<n_targets, kGSThreads, 0, stream>>>(w, h, pitch, targets, max_per_target, grid_rows, grid_cols, per_cell);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
