// K2/K3: sparse pyramidal Lucas-Kanade, one warp per point, all pyramid levels inside the warp, with the
// forward-backward pass fused into the same launch.
//
// Replaces cv::calcOpticalFlowPyrLK at psn_where/PSNWhere_Tracker2D.cpp:776-782 (backward chain) and
// :871-877 (forward), SURVEY.md section 8 rows a4/a5.  The control flow and the fixed-point arithmetic are
// those of OpenCV's LKTrackerInvoker: 14-bit bilinear weights rounded half to
// even, I patch = u8 x 32 as int16, Ix/Iy = Scharr x 32 descaled by 14 bits, structure tensor and mismatch
// vector scaled by 2^-20, float32 solve with unfused multiplies (the file is compiled with -fmad=false and
// the solve uses explicit _rn intrinsics), the same bounds / min-eigenvalue / epsilon / oscillation exits.
// Sums over the window are accumulated exactly in integers and rounded to float once (OpenCV accumulates
// in float; the difference is its own rounding noise, ~1e-4 px).
//
// Two kernels:
//   lk_generic_kernel   any window size (reference-faithful (int)box.w x (int)box.h windows): template patch
//                       in shared memory, image reads through L1 with on-the-fly REFLECT_101 / zero borders.
//   lk_w21_kernel       21x21 window (BASELINE configs): the 441 pixels are 63 horizontal runs of 7, two runs
//                       per lane; the template (I, Ix, Iy) lives in registers; the 22x22 u8 patch of the
//                       target image is staged per iteration into shared memory with aligned 32-bit loads
//                       and bilinearly sampled with dp2a on packed bytes; reductions use REDUX.
#include <float.h>

#include "common.cuh"

namespace psnw {

namespace {

constexpr float kFltScale = 1.f / (1 << 20);

__device__ __forceinline__ int refl(int p, int n)
{
    p = p < 0 ? -p : p;
    return p >= n ? 2 * (n - 1) - p : p;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int &w00, int &w01, int &w10, int &w11)
{
    const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(ia, ib), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ib), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(ia, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exact warp sum of per-lane int32 partials whose total may exceed 32 bits: two REDUX
__device__ __forceinline__ long long warp_sum_i32(int v)
{
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xffff));
    return ((long long)hi << 16) + (long long)lo;
}

struct Solve {
    float x, y;   // nextPts
    float err;
    int st;
};

// 2x2 system shared by both kernels; returns false when the level is rejected (min-eig / determinant gate)
__device__ __forceinline__ bool make_system(long long s11, long long s12, long long s22, int ww, int wh, float min_eig_thr,
                                            float &A11, float &A12, float &A22, float &Dinv)
{
    A11 = __fmul_rn(__ll2float_rn(s11), kFltScale);
    A12 = __fmul_rn(__ll2float_rn(s12), kFltScale);
    A22 = __fmul_rn(__ll2float_rn(s22), kFltScale);
    const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rad), (float)(2 * ww * wh));
    if (min_eig < min_eig_thr || D < FLT_EPSILON) return false;
    Dinv = __fdiv_rn(1.f, D);
    return true;
}

// one Gauss-Newton update; returns true when the iteration loop must stop
__device__ __forceinline__ bool lk_step(long long sb1, long long sb2, float A11, float A12, float A22, float Dinv,
                                        float half_x, float half_y, double eps2, int j,
                                        float &nx, float &ny, float &pdx, float &pdy, float &next_x, float &next_y)
{
    const float b1 = __fmul_rn(__ll2float_rn(sb1), kFltScale), b2 = __fmul_rn(__ll2float_rn(sb2), kFltScale);
    const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dinv);
    const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dinv);
    nx = __fadd_rn(nx, dx);
    ny = __fadd_rn(ny, dy);
    next_x = __fadd_rn(nx, half_x);
    next_y = __fadd_rn(ny, half_y);
    if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2) return true;
    if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
        next_x = __fsub_rn(next_x, __fmul_rn(dx, 0.5f));
        next_y = __fsub_rn(next_y, __fmul_rn(dy, 0.5f));
        return true;
    }
    pdx = dx;
    pdy = dy;
    return false;
}

// ------------------------------------------------------------------------------------------------
// generic window
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sample_u8(const uint8_t *img, int pitch, int w, int h, int X, int Y,
                                         int w00, int w01, int w10, int w11)
{
    const int X0 = refl(X, w), X1 = refl(X + 1, w);
    const uint8_t *r0 = img + (size_t)refl(Y, h) * pitch, *r1 = img + (size_t)refl(Y + 1, h) * pitch;
    return (r0[X0] * w00 + r0[X1] * w01 + r1[X0] * w10 + r1[X1] * w11 + (1 << 8)) >> 9;
}

__device__ void solve_generic(const LkParams &P, int plane_a, int plane_b, int ww, int wh, int max_level,
                              float px, float py, short *Iw, short2 *dIw, int lane, Solve &o)
{
    const float half_x = (ww - 1) * 0.5f, half_y = (wh - 1) * 0.5f;
    const int npx = ww * wh;
    float next_x = 0.f, next_y = 0.f, er = 0.f;
    int st = 1;
    int lx0 = lane, ly0 = 0;                   // window coordinates of this lane's first pixel
    while (lx0 >= ww) { lx0 -= ww; ly0++; }

    for (int level = max_level; level >= 0; --level) {
        const LevelGeom &L = P.lv[level];
        const uint8_t *IA = L.img + (size_t)plane_a * L.plane_px;
        const short2 *DA = L.dxy + (size_t)plane_a * L.plane_px;
        const uint8_t *JB = L.img + (size_t)plane_b * L.plane_px;
        const float sc = __int_as_float((127 - level) << 23);
        float prev_x = __fmul_rn(px, sc), prev_y = __fmul_rn(py, sc);
        float nx, ny;
        if (level == max_level) { nx = prev_x; ny = prev_y; }
        else                    { nx = __fmul_rn(next_x, 2.f); ny = __fmul_rn(next_y, 2.f); }
        next_x = nx; next_y = ny;

        prev_x = __fsub_rn(prev_x, half_x); prev_y = __fsub_rn(prev_y, half_y);
        const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
        if (ipx < -ww || ipx >= L.w || ipy < -wh || ipy >= L.h) {
            if (level == 0) { st = 0; er = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(prev_x, (float)ipx), __fsub_rn(prev_y, (float)ipy), w00, w01, w10, w11);

        long long s11 = 0, s12 = 0, s22 = 0;
        {
            int x = lx0, y = ly0;
            for (int p = lane; p < npx; p += 32) {
                const int X = ipx + x, Y = ipy + y;
                const int ival = sample_u8(IA, L.pitch, L.w, L.h, X, Y, w00, w01, w10, w11);
                const bool xi0 = (unsigned)X < (unsigned)L.w, xi1 = (unsigned)(X + 1) < (unsigned)L.w;
                const bool yi0 = (unsigned)Y < (unsigned)L.h, yi1 = (unsigned)(Y + 1) < (unsigned)L.h;
                const short2 z = make_short2(0, 0);
                const short2 *d0 = DA + (ptrdiff_t)Y * L.pitch + X, *d1 = d0 + L.pitch;
                const short2 d00 = (xi0 && yi0) ? d0[0] : z, d01 = (xi1 && yi0) ? d0[1] : z;
                const short2 d10 = (xi0 && yi1) ? d1[0] : z, d11 = (xi1 && yi1) ? d1[1] : z;
                const int ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
                const int iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
                Iw[p] = (short)ival;
                dIw[p] = make_short2((short)ix, (short)iy);
                s11 += (long long)(ix * ix); s12 += (long long)(ix * iy); s22 += (long long)(iy * iy);
                x += 32;
                while (x >= ww) { x -= ww; y++; }
            }
        }
        s11 = warp_sum_ll(s11); s12 = warp_sum_ll(s12); s22 = warp_sum_ll(s22);
        float A11, A12, A22, Dinv;
        if (!make_system(s11, s12, s22, ww, wh, P.min_eig, A11, A12, A22, Dinv)) {
            if (level == 0) st = 0;
            continue;
        }

        nx = __fsub_rn(nx, half_x); ny = __fsub_rn(ny, half_y);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_iters; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -ww || inx >= L.w || iny < -wh || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            long long sb1 = 0, sb2 = 0;
            int x = lx0, y = ly0;
            for (int p = lane; p < npx; p += 32) {
                const int diff = sample_u8(JB, L.pitch, L.w, L.h, inx + x, iny + y, w00, w01, w10, w11) - Iw[p];
                const short2 d = dIw[p];
                sb1 += (long long)(diff * d.x); sb2 += (long long)(diff * d.y);
                x += 32;
                while (x >= ww) { x -= ww; y++; }
            }
            sb1 = warp_sum_ll(sb1); sb2 = warp_sum_ll(sb2);
            if (lk_step(sb1, sb2, A11, A12, A22, Dinv, half_x, half_y, P.eps2, j, nx, ny, pdx, pdy, next_x, next_y)) break;
        }

        if (st && level == 0) {
            const float qx = __fsub_rn(next_x, half_x), qy = __fsub_rn(next_y, half_y);
            const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
            if (inx < -ww || inx >= L.w || iny < -wh || iny >= L.h) { st = 0; continue; }
            bilinear_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            long long se = 0;
            int x = lx0, y = ly0;
            for (int p = lane; p < npx; p += 32) {
                const int diff = sample_u8(JB, L.pitch, L.w, L.h, inx + x, iny + y, w00, w01, w10, w11) - Iw[p];
                se += diff < 0 ? -diff : diff;
                x += 32;
                while (x >= ww) { x -= ww; y++; }
            }
            se = warp_sum_ll(se);
            er = __fdiv_rn(__ll2float_rn(se), (float)(32 * ww * wh));
        }
    }
    o.x = next_x; o.y = next_y; o.st = st; o.err = st ? er : 0.f;
}

// point index -> job (n_jobs is small; linear scan is uniform across the warp)
__device__ __forceinline__ int find_job(const TrackJob *jobs, int n_jobs, int i)
{
    int j = 0;
    while (j + 1 < n_jobs && i >= jobs[j + 1].pt_begin) ++j;
    return j;
}

__device__ __forceinline__ void finish_point(const LkParams &P, const TrackJob &job, int k, float2 p0,
                                             const Solve &f, const Solve &b, int lane)
{
    if (lane != 0) return;
    int st = f.st;
    float fb = 0.f;
    if (P.do_fb) {
        const float dx = __fsub_rn(p0.x, b.x), dy = __fsub_rn(p0.y, b.y);
        fb = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        st = (f.st && b.st && fb <= P.fb_thresh) ? 1 : 0;
        if (job.fb_err) job.fb_err[k] = fb;
    }
    job.out[k] = make_float2(f.x, f.y);
    job.status[k] = (uint8_t)st;
    job.err[k] = f.err;
}

__global__ void __launch_bounds__(128)
lk_generic_kernel(const __grid_constant__ LkParams P, int max_win_px)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= P.n_pts) return;
    const int stride = (max_win_px + 7) & ~7;
    short2 *dIw = reinterpret_cast<short2 *>(smem_raw) + (size_t)wib * stride;
    short *Iw = reinterpret_cast<short *>(reinterpret_cast<short2 *>(smem_raw) + (size_t)(blockDim.x >> 5) * stride) + (size_t)wib * stride;

    const TrackJob &job = P.jobs[find_job(P.jobs, P.n_jobs, i)];
    const int k = i - job.pt_begin;
    const int t = job.target_of_pt ? job.target_of_pt[k] : 0;
    const TrackGroup g = P.groups[job.group_base + min(max(t, 0), job.n_targets - 1)];
    const float2 p0 = job.pts[k];

    Solve f, b;
    b.x = b.y = b.err = 0.f; b.st = 1;
    solve_generic(P, g.plane_from, g.plane_to, g.win_w, g.win_h, g.max_level, p0.x, p0.y, Iw, dIw, lane, f);
    if (P.do_fb) solve_generic(P, g.plane_to, g.plane_from, g.win_w, g.win_h, g.max_level, f.x, f.y, Iw, dIw, lane, b);
    finish_point(P, job, k, p0, f, b, lane);
}

// ------------------------------------------------------------------------------------------------
// 21 x 21 window
// ------------------------------------------------------------------------------------------------
constexpr int kW = 21;                    // window
constexpr int kPatch = kW + 1;            // 22 x 22 samples
constexpr int kPatchRowWords = 8;         // u8 patch row = 32 bytes (22 + up to 3 alignment bytes)
constexpr int kU8Words = kPatch * kPatchRowWords;            // 176 words
constexpr int kDxyWords = kPatch * kPatch;                   // 484 words (short2 per sample)
constexpr int kWarpSmemWords = kU8Words + kDxyWords;         // 660 words = 2640 B per warp
constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ int dp2a_lo_su(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Stage the 22x22 u8 patch whose top-left sample is (ix, iy) into `patch` (row pitch 32 B).  Returns the byte
// offset of sample column 0 inside a patch row.  Interior patches use aligned 32-bit loads; patches touching
// the frame border gather bytes through REFLECT_101 (the apron OpenCV keeps around every level).
__device__ __forceinline__ int stage_u8_patch(const uint8_t *img, int pitch, int w, int h, int ix, int iy,
                                              uint32_t *patch, int lane)
{
    if (ix >= 0 && iy >= 0 && ix + kW < w && iy + kW < h) {
        const int off = ix & 3;
        const int nwords = (off + kPatch + 3) >> 2;                    // 6 or 7
        const uint32_t *src = reinterpret_cast<const uint32_t *>(img + (size_t)iy * pitch + (ix - off));
        const int pitch_w = pitch >> 2;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int wi = lane + 32 * r;                              // 22 rows x 7 words = 154
            const int row = (wi * 147) >> 10;                          // wi / 7 for wi < 160
            const int c = wi - row * 7;
            if (wi < kPatch * 7 && c < nwords) patch[row * kPatchRowWords + c] = __ldg(src + (size_t)row * pitch_w + c);
        }
        return off;
    }
    uint8_t *pb = reinterpret_cast<uint8_t *>(patch);
    for (int bi = lane; bi < kPatch * kPatch; bi += 32) {
        const int row = bi / kPatch, col = bi - row * kPatch;
        pb[row * (kPatchRowWords * 4) + col] = img[(size_t)refl(iy + row, h) * pitch + refl(ix + col, w)];
    }
    return 0;
}

// Pre-shift bilinear sums (rounding constant included) of a horizontal run of 7 samples starting at byte
// column `bc` of patch row `row`; wt/wb = (w_left | w_right << 16) of the top / bottom row.
__device__ __forceinline__ void bilinear_run7(const uint32_t *patch, int row, int bc, int wt, int wb, int v[7])
{
    const uint32_t *p = patch + row * kPatchRowWords + (bc >> 2);
    const int sh = (bc & 3) * 8;
    const uint32_t t0 = __funnelshift_r(p[0], p[1], sh), t1 = __funnelshift_r(p[1], p[2], sh);
    const uint32_t u0 = __funnelshift_r(p[kPatchRowWords], p[kPatchRowWords + 1], sh);
    const uint32_t u1 = __funnelshift_r(p[kPatchRowWords + 1], p[kPatchRowWords + 2], sh);
    const uint32_t ts = __funnelshift_r(t0, t1, 8), us = __funnelshift_r(u0, u1, 8);   // bytes 1..4
    const uint32_t t5 = t1 >> 8, u5 = u1 >> 8;                                         // bytes 5..7
    v[0] = dp2a_lo_su(wb, u0, dp2a_lo_su(wt, t0, 256));
    v[1] = dp2a_lo_su(wb, us, dp2a_lo_su(wt, ts, 256));
    v[2] = dp2a_hi_su(wb, u0, dp2a_hi_su(wt, t0, 256));
    v[3] = dp2a_hi_su(wb, us, dp2a_hi_su(wt, ts, 256));
    v[4] = dp2a_lo_su(wb, u1, dp2a_lo_su(wt, t1, 256));
    v[5] = dp2a_lo_su(wb, u5, dp2a_lo_su(wt, t5, 256));
    v[6] = dp2a_hi_su(wb, u1, dp2a_hi_su(wt, t1, 256));
}

__device__ void solve_w21(const LkParams &P, int plane_a, int plane_b, int max_level, float px, float py,
                          uint32_t *u8p, uint32_t *dxp, int lane, Solve &o)
{
    constexpr float half = (kW - 1) * 0.5f;
    float next_x = 0.f, next_y = 0.f, er = 0.f;
    int st = 1;
    // this lane's two runs: run r -> (row r / 3, first column 7 * (r % 3)); lane 31 owns a single run
    const bool two = lane < 31;
    const int r0 = lane, r1 = two ? lane + 32 : 62;
    const int row0 = r0 / 3, col0 = 7 * (r0 - 3 * row0);
    const int row1 = r1 / 3, col1 = 7 * (r1 - 3 * row1);

    for (int level = max_level; level >= 0; --level) {
        const LevelGeom &L = P.lv[level];
        const uint8_t *IA = L.img + (size_t)plane_a * L.plane_px;
        const short2 *DA = L.dxy + (size_t)plane_a * L.plane_px;
        const uint8_t *JB = L.img + (size_t)plane_b * L.plane_px;
        const float sc = __int_as_float((127 - level) << 23);
        float prev_x = __fmul_rn(px, sc), prev_y = __fmul_rn(py, sc);
        float nx, ny;
        if (level == max_level) { nx = prev_x; ny = prev_y; }
        else                    { nx = __fmul_rn(next_x, 2.f); ny = __fmul_rn(next_y, 2.f); }
        next_x = nx; next_y = ny;

        prev_x = __fsub_rn(prev_x, half); prev_y = __fsub_rn(prev_y, half);
        const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
        if (ipx < -kW || ipx >= L.w || ipy < -kW || ipy >= L.h) {
            if (level == 0) { st = 0; er = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(prev_x, (float)ipx), __fsub_rn(prev_y, (float)ipy), w00, w01, w10, w11);

        // ---- template: stage I (u8) and dI (short2) patches, then sample them into registers ----
        __syncwarp();
        const int offI = stage_u8_patch(IA, L.pitch, L.w, L.h, ipx, ipy, u8p, lane);
        for (int wi = lane; wi < kDxyWords; wi += 32) {
            const int row = wi / kPatch, col = wi - row * kPatch;
            const int X = ipx + col, Y = ipy + row;
            uint32_t v = 0;
            if ((unsigned)X < (unsigned)L.w && (unsigned)Y < (unsigned)L.h)
                v = __ldg(reinterpret_cast<const uint32_t *>(DA + (size_t)Y * L.pitch + X));
            dxp[wi] = v;
        }
        __syncwarp();

        int tI[14], tIx[14], tIy[14];
        int a11 = 0, a12 = 0, a22 = 0;
        const int wt = (w00 & 0xffff) | (w01 << 16), wb = (w10 & 0xffff) | (w11 << 16);
#pragma unroll
        for (int run = 0; run < 2; ++run) {
            const int row = run ? row1 : row0, col = run ? col1 : col0;
            int v[7];
            bilinear_run7(u8p, row, offI + col, wt, wb, v);
            const uint32_t *d0 = dxp + row * kPatch + col, *d1 = d0 + kPatch;
            uint32_t q0 = d0[0], q1 = d1[0];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const uint32_t n0 = d0[k + 1], n1 = d1[k + 1];
                const int x00 = (int)(q0 << 16) >> 16, y00 = (int)q0 >> 16, x01 = (int)(n0 << 16) >> 16, y01 = (int)n0 >> 16;
                const int x10 = (int)(q1 << 16) >> 16, y10 = (int)q1 >> 16, x11 = (int)(n1 << 16) >> 16, y11 = (int)n1 >> 16;
                int ix = (x00 * w00 + x01 * w01 + x10 * w10 + x11 * w11 + (1 << 13)) >> 14;
                int iy = (y00 * w00 + y01 * w01 + y10 * w10 + y11 * w11 + (1 << 13)) >> 14;
                if (run && !two) { ix = 0; iy = 0; }
                tI[run * 7 + k] = v[k] >> 9;
                tIx[run * 7 + k] = ix;
                tIy[run * 7 + k] = iy;
                a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                q0 = n0; q1 = n1;
            }
        }
        const long long s11 = warp_sum_i32(a11), s12 = warp_sum_i32(a12), s22 = warp_sum_i32(a22);
        float A11, A12, A22, Dinv;
        if (!make_system(s11, s12, s22, kW, kW, P.min_eig, A11, A12, A22, Dinv)) {
            if (level == 0) st = 0;
            continue;
        }

        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_iters; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -kW || inx >= L.w || iny < -kW || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const int jt = (w00 & 0xffff) | (w01 << 16), jb = (w10 & 0xffff) | (w11 << 16);
            __syncwarp();
            const int offJ = stage_u8_patch(JB, L.pitch, L.w, L.h, inx, iny, u8p, lane);
            __syncwarp();
            int b1 = 0, b2 = 0;
#pragma unroll
            for (int run = 0; run < 2; ++run) {
                int v[7];
                bilinear_run7(u8p, run ? row1 : row0, offJ + (run ? col1 : col0), jt, jb, v);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int diff = (v[k] >> 9) - tI[run * 7 + k];
                    b1 += diff * tIx[run * 7 + k];
                    b2 += diff * tIy[run * 7 + k];
                }
            }
            const long long sb1 = warp_sum_i32(b1), sb2 = warp_sum_i32(b2);
            if (lk_step(sb1, sb2, A11, A12, A22, Dinv, half, half, P.eps2, j, nx, ny, pdx, pdy, next_x, next_y)) break;
        }

        if (st && level == 0) {
            const float qx = __fsub_rn(next_x, half), qy = __fsub_rn(next_y, half);
            const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
            if (inx < -kW || inx >= L.w || iny < -kW || iny >= L.h) { st = 0; continue; }
            bilinear_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            const int jt = (w00 & 0xffff) | (w01 << 16), jb = (w10 & 0xffff) | (w11 << 16);
            __syncwarp();
            const int offJ = stage_u8_patch(JB, L.pitch, L.w, L.h, inx, iny, u8p, lane);
            __syncwarp();
            int e = 0;
#pragma unroll
            for (int run = 0; run < 2; ++run) {
                int v[7];
                bilinear_run7(u8p, run ? row1 : row0, offJ + (run ? col1 : col0), jt, jb, v);
                if (run == 0 || two) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const int diff = (v[k] >> 9) - tI[run * 7 + k];
                        e += diff < 0 ? -diff : diff;
                    }
                }
            }
            const long long se = warp_sum_i32(e);
            er = __fdiv_rn(__ll2float_rn(se), (float)(32 * kW * kW));
        }
    }
    o.x = next_x; o.y = next_y; o.st = st; o.err = st ? er : 0.f;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
lk_w21_kernel(const __grid_constant__ LkParams P)
{
    __shared__ __align__(16) uint32_t smem[kWarpsPerCta * kWarpSmemWords];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarpsPerCta + wib;
    if (i >= P.n_pts) return;
    uint32_t *u8p = smem + wib * kWarpSmemWords, *dxp = u8p + kU8Words;

    const TrackJob &job = P.jobs[find_job(P.jobs, P.n_jobs, i)];
    const int k = i - job.pt_begin;
    const int t = job.target_of_pt ? job.target_of_pt[k] : 0;
    const TrackGroup g = P.groups[job.group_base + min(max(t, 0), job.n_targets - 1)];
    const float2 p0 = job.pts[k];

    Solve f, b;
    b.x = b.y = b.err = 0.f; b.st = 1;
    solve_w21(P, g.plane_from, g.plane_to, g.max_level, p0.x, p0.y, u8p, dxp, lane, f);
    if (P.do_fb) solve_w21(P, g.plane_to, g.plane_from, g.max_level, f.x, f.y, u8p, dxp, lane, b);
    finish_point(P, job, k, p0, f, b, lane);
}

}  // namespace

int launch_lk(const LkParams &P, int all_21x21, int max_win_px, cudaStream_t stream)
{
    if (P.n_pts <= 0) return 0;
    if (all_21x21) {
        const int grid = (P.n_pts + kWarpsPerCta - 1) / kWarpsPerCta;
        lk_w21_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(P);
    } else {
        const int stride = (max_win_px + 7) & ~7;
        int warps = 4;
        size_t smem = (size_t)warps * stride * 6;
        while (warps > 1 && smem > 200 * 1024) { warps >>= 1; smem = (size_t)warps * stride * 6; }
        if (smem > 200 * 1024) return -2;
        static size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            if (cudaFuncSetAttribute(lk_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
            configured = 200 * 1024;
        }
        const int grid = (P.n_pts + warps - 1) / warps;
        lk_generic_kernel<<<grid, warps * 32, smem, stream>>>(P, max_win_px);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
