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
//   lk_w21_kernel       21x21 window (BASELINE configs): persistent warps draw points from a device counter; the
//                       441 pixels are 63 horizontal runs of 7, two runs per lane; the template (I, Ix, Iy) lives in
//                       registers; patches are staged by TMA boxes completing on per-warp mbarriers (source
//                       patches one level ahead, the target patch once per level with a 4-px margin) and
//                       bilinearly sampled with dp2a on packed bytes; reductions use REDUX.
#include <float.h>

#include "common.cuh"
#include "pyr_common.cuh"

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

// exact warp sum of per-lane int32 partials whose total may exceed 32 bits: two REDUX (high and low halves);
// the same sum as a float, rounded once: hi * 2^16 + lo is formed exactly inside the fma (|hi| < 2^24, lo < 2^24), so the
// result equals __ll2float_rn of the 64-bit sum without 64-bit integer arithmetic or the slow I2F.S64
struct WarpSum {
    int hi;
    unsigned lo;
};
__device__ __forceinline__ WarpSum warp_sum_split(int v)
{
    WarpSum r;
    r.hi = __reduce_add_sync(0xffffffffu, v >> 16);
    r.lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xffff));
    return r;
}
__device__ __forceinline__ float sum_to_float(WarpSum s) { return __fmaf_rn((float)s.hi, 65536.f, (float)s.lo); }

struct Solve {
    float x, y;   // nextPts
    float err;
    int st;
};

// 2x2 system shared by both kernels; returns false when the level is rejected (min-eig / determinant gate)
__device__ __forceinline__ bool make_system(float f11, float f12, float f22, int ww, int wh, float min_eig_thr,
                                            float &A11, float &A12, float &A22, float &Dinv)
{
    A11 = __fmul_rn(f11, kFltScale);
    A12 = __fmul_rn(f12, kFltScale);
    A22 = __fmul_rn(f22, kFltScale);
    const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rad), (float)(2 * ww * wh));
    if (min_eig < min_eig_thr || D < FLT_EPSILON) return false;
    Dinv = __fdiv_rn(1.f, D);
    return true;
}

// one Gauss-Newton update; returns true when the iteration loop must stop.
//   delta.ddot(delta) <= eps^2 is a double comparison in OpenCV; the float value decides it whenever it is further
//   than 1e-5 (relative) from the threshold, the double product is only formed inside that band.
//   fabs((double)(float)s) < 0.01 is exactly fabsf(s) <= 0.01f (0.01f is the largest float below the double 0.01).
__device__ __forceinline__ bool lk_step(const LkParams &P, float fb1, float fb2, float A11, float A12, float A22,
                                        float Dinv, float half_x, float half_y, int j,
                                        float &nx, float &ny, float &pdx, float &pdy, float &next_x, float &next_y)
{
    const float b1 = __fmul_rn(fb1, kFltScale), b2 = __fmul_rn(fb2, kFltScale);
    const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dinv);
    const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dinv);
    nx = __fadd_rn(nx, dx);
    ny = __fadd_rn(ny, dy);
    next_x = __fadd_rn(nx, half_x);
    next_y = __fadd_rn(ny, half_y);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    bool small = d2 < P.eps2_lo;
    if (!small && !(d2 > P.eps2_hi))
        small = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= P.eps2;
    if (small) return true;
    if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
        next_x = __fsub_rn(next_x, __fmul_rn(dx, 0.5f));
        next_y = __fsub_rn(next_y, __fmul_rn(dy, 0.5f));
        return true;
    }
    pdx = dx;
    pdy = dy;
    return false;
}

// packed-byte helpers of both kernels
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

__device__ __forceinline__ int sext16(uint32_t w) { return (int)(short)(w & 0xffffu); }

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

constexpr int kGenThreads = 128;           // generic-window kernel: one CTA per point

// exact block-wide sums of three int64 values, result broadcast to every thread; `sel` alternates between two
// scratch slots so that one barrier per call is enough
__device__ __forceinline__ void block_sum3(long long &a, long long &b, long long &c, long long *red, int &sel, int tid)
{
    a = warp_sum_ll(a); b = warp_sum_ll(b); c = warp_sum_ll(c);
    long long *r = red + sel * (3 * (kGenThreads / 32));
    if ((tid & 31) == 0) { r[(tid >> 5) * 3] = a; r[(tid >> 5) * 3 + 1] = b; r[(tid >> 5) * 3 + 2] = c; }
    __syncthreads();
    a = 0; b = 0; c = 0;
#pragma unroll
    for (int w = 0; w < kGenThreads / 32; ++w) { a += r[w * 3]; b += r[w * 3 + 1]; c += r[w * 3 + 2]; }
    sel ^= 1;
}

__device__ void solve_generic(const LkParams &P, int plane_a, int plane_b, int ww, int wh, int max_level,
                              float px, float py, short *Iw, short2 *dIw, int tid, long long *red, Solve &o)
{
    const float half_x = (ww - 1) * 0.5f, half_y = (wh - 1) * 0.5f;
    const int npx = ww * wh;
    float next_x = 0.f, next_y = 0.f, er = 0.f;
    int st = 1;
    int lx0 = tid, ly0 = 0;                    // window coordinates of this thread's first pixel
    int rsel = 0;                              // reduction scratch toggles between two slots
    const bool acc32 = npx <= 60 * kGenThreads;   // 60 * 8160 * 4080 < 2^31: per-thread partial sums fit 32 bits
    while (lx0 >= ww) { lx0 -= ww; ly0++; }
    const int step_y = kGenThreads / ww, step_x = kGenThreads - step_y * ww;   // one stride of kGenThreads pixels in (x, y)
    // template rows are padded to a multiple of 4 pixels (index q = y * wp + x): the 4-pixel inner loop reads them as vectors
    const int wp = (ww + 3) & ~3, sq = step_y * wp + step_x, q0 = ly0 * wp + lx0;
    // the 4-pixel inner loop walks groups of 4 adjacent pixels: G groups per row, this thread's first group (gx0, gy0)
    const int G = wp >> 2, gy0 = tid / G, gx0 = tid - gy0 * G, sgy = kGenThreads / G, sgx = kGenThreads - sgy * G;
    const bool acc32_4 = wp * wh <= 60 * kGenThreads;

    for (int level = max_level; level >= 0; --level) {
        const LevelGeom &L = P.lv[level];
        const uint8_t *IA = L.img + (size_t)plane_a * L.iplane_px;
        const short2 *DA = L.dxy + (size_t)plane_a * L.plane_px;
        const uint8_t *JB = L.img + (size_t)plane_b * L.iplane_px;
        const int ipitch = L.ipitch;
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
            int x = lx0, y = ly0, q = q0;
            const int apt = L.apron;
            if (acc32_4 && ipx >= -apt && ipy >= -apt && ipx + ww < L.w + apt && ipy + wh < L.h + apt &&
                ipx + wp + 4 <= ipitch - L.apron_l) {
                // Window inside the (padded) frame: four adjacent pixels per step like the iteration path below -- two aligned
                // image words per row, funnel shifts to the window's byte phase, dp2a on the packed bytes; five derivative words
                // per row serve the four bilinear taps; the template goes out as vectors.  The image taps of a window that leaves
                // the frame read the REFLECT_101 apron kept in HBM; its derivative taps outside the frame are zero (OpenCV's
                // constant border of the derivative image) and take the guarded loads.
                const int sh8 = (ipx & 3) * 8;
                const int wt = (w00 & 0xffff) | (w01 << 16), wb = (w10 & 0xffff) | (w11 << 16);
                const uint8_t *base = IA + (ptrdiff_t)ipy * ipitch + (ipx - (ipx & 3));
                int a11 = 0, a12 = 0, a22 = 0, gx = gx0, gy = gy0;
                for (int it = tid; it < G * wh; it += kGenThreads) {
                    const uint32_t *r0 = reinterpret_cast<const uint32_t *>(base + (ptrdiff_t)gy * ipitch + 4 * gx);
                    const uint32_t *r1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(r0) + ipitch);
                    const uint32_t i0 = __ldg(r0), i1 = __ldg(r0 + 1), z0 = __ldg(r1), z1 = __ldg(r1 + 1);
                    const uint32_t t0 = __funnelshift_r(i0, i1, sh8), t4 = i1 >> sh8;          // bytes 0..3, byte 4
                    const uint32_t u0 = __funnelshift_r(z0, z1, sh8), u4 = z1 >> sh8;
                    const uint32_t ts = __funnelshift_r(t0, t4, 8), us = __funnelshift_r(u0, u4, 8);   // bytes 1..4
                    int iv[4];
                    iv[0] = dp2a_lo_su(wb, u0, dp2a_lo_su(wt, t0, 256)) >> 9; iv[1] = dp2a_lo_su(wb, us, dp2a_lo_su(wt, ts, 256)) >> 9;
                    iv[2] = dp2a_hi_su(wb, u0, dp2a_hi_su(wt, t0, 256)) >> 9; iv[3] = dp2a_hi_su(wb, us, dp2a_hi_su(wt, ts, 256)) >> 9;
                    const int X = ipx + 4 * gx, Y = ipy + gy;
                    const uint32_t *d0 = reinterpret_cast<const uint32_t *>(DA) + (ptrdiff_t)Y * L.pitch + X, *d1 = d0 + L.pitch;
                    uint32_t e0[5], e1[5];
                    if (X >= 0 && X + 4 < L.w && Y >= 0 && Y + 1 < L.h) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) { e0[k] = __ldg(d0 + k); e1[k] = __ldg(d1 + k); }
                    } else {
                        const bool y0in = (unsigned)Y < (unsigned)L.h, y1in = (unsigned)(Y + 1) < (unsigned)L.h;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const bool xin = (unsigned)(X + k) < (unsigned)L.w;
                            e0[k] = (xin && y0in) ? __ldg(d0 + k) : 0u;
                            e1[k] = (xin && y1in) ? __ldg(d1 + k) : 0u;
                        }
                    }
                    const int valid = ww - 4 * gx;           // >= 4 except in the last group of a row
                    int ixv[4], iyv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int ix = (sext16(e0[k]) * w00 + sext16(e0[k + 1]) * w01 + sext16(e1[k]) * w10 + sext16(e1[k + 1]) * w11 + (1 << 13)) >> 14;
                        int iy = (((int)e0[k] >> 16) * w00 + ((int)e0[k + 1] >> 16) * w01 + ((int)e1[k] >> 16) * w10 + ((int)e1[k + 1] >> 16) * w11 + (1 << 13)) >> 14;
                        if (k >= valid) { ix = 0; iy = 0; }  // padding columns of the template row: never part of a sum
                        ixv[k] = ix; iyv[k] = iy;
                        a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    }
                    const int q4 = gy * wp + 4 * gx;
                    *reinterpret_cast<uint2 *>(Iw + q4) = make_uint2((uint32_t)(iv[0] & 0xffff) | ((uint32_t)iv[1] << 16), (uint32_t)(iv[2] & 0xffff) | ((uint32_t)iv[3] << 16));
                    *reinterpret_cast<uint4 *>(dIw + q4) = make_uint4((uint32_t)(ixv[0] & 0xffff) | ((uint32_t)iyv[0] << 16), (uint32_t)(ixv[1] & 0xffff) | ((uint32_t)iyv[1] << 16),
                                                                     (uint32_t)(ixv[2] & 0xffff) | ((uint32_t)iyv[2] << 16), (uint32_t)(ixv[3] & 0xffff) | ((uint32_t)iyv[3] << 16));
                    gx += sgx; gy += sgy;
                    if (gx >= G) { gx -= G; gy++; }
                }
                s11 = a11; s12 = a12; s22 = a22;
            } else if (acc32 && ipx >= 0 && ipy >= 0 && ipx + ww < L.w && ipy + wh < L.h) {
                // window (plus its bilinear ring) inside the frame: no border arithmetic, 32-bit partial sums
                int a11 = 0, a12 = 0, a22 = 0;
                // running offsets instead of a 2-D address per pixel: one stride of kGenThreads pixels, minus a row wrap
                int oi = (ipy + y) * ipitch + (ipx + x), od = (ipy + y) * L.pitch + (ipx + x);
                const int si = step_y * ipitch + step_x, sd = step_y * L.pitch + step_x;
                for (int p = tid; p < npx; p += kGenThreads) {
                    const uint8_t *r0 = IA + oi, *r1 = r0 + ipitch;
                    const int ival = (r0[0] * w00 + r0[1] * w01 + r1[0] * w10 + r1[1] * w11 + (1 << 8)) >> 9;
                    const short2 *d0 = DA + od, *d1 = d0 + L.pitch;
                    const short2 d00 = d0[0], d01 = d0[1], d10 = d1[0], d11 = d1[1];
                    const int ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
                    const int iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
                    Iw[q] = (short)ival;
                    dIw[q] = make_short2((short)ix, (short)iy);
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    x += step_x; q += sq; oi += si; od += sd;
                    if (x >= ww) { x -= ww; q += wp - ww; oi += ipitch - ww; od += L.pitch - ww; }
                }
                s11 = a11; s12 = a12; s22 = a22;
            } else {
            // the image taps of a window that stays inside the padded frame (apron levels) need no mirror arithmetic
            const int ap = L.apron;
            const bool img_direct = ipx >= -ap && ipy >= -ap && ipx + ww < L.w + ap && ipy + wh < L.h + ap;
            for (int p = tid; p < npx; p += kGenThreads) {
                const int X = ipx + x, Y = ipy + y;
                int ival;
                if (img_direct) {
                    const uint8_t *r0 = IA + Y * ipitch + X, *r1 = r0 + ipitch;
                    ival = (r0[0] * w00 + r0[1] * w01 + r1[0] * w10 + r1[1] * w11 + (1 << 8)) >> 9;
                } else {
                    ival = sample_u8(IA, ipitch, L.w, L.h, X, Y, w00, w01, w10, w11);
                }
                const bool xi0 = (unsigned)X < (unsigned)L.w, xi1 = (unsigned)(X + 1) < (unsigned)L.w;
                const bool yi0 = (unsigned)Y < (unsigned)L.h, yi1 = (unsigned)(Y + 1) < (unsigned)L.h;
                const short2 z = make_short2(0, 0);
                const short2 *d0 = DA + (ptrdiff_t)Y * L.pitch + X, *d1 = d0 + L.pitch;
                const short2 d00 = (xi0 && yi0) ? d0[0] : z, d01 = (xi1 && yi0) ? d0[1] : z;
                const short2 d10 = (xi0 && yi1) ? d1[0] : z, d11 = (xi1 && yi1) ? d1[1] : z;
                const int ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
                const int iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
                Iw[q] = (short)ival;
                dIw[q] = make_short2((short)ix, (short)iy);
                s11 += (long long)(ix * ix); s12 += (long long)(ix * iy); s22 += (long long)(iy * iy);
                x += step_x; y += step_y; q += sq;
                if (x >= ww) { x -= ww; y++; q += wp - ww; }
            }
            }
        }
        block_sum3(s11, s12, s22, red, rsel, tid);
        float A11, A12, A22, Dinv;
        if (!make_system(__ll2float_rn(s11), __ll2float_rn(s12), __ll2float_rn(s22), ww, wh, P.min_eig, A11, A12, A22, Dinv)) {
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
            int x = lx0, y = ly0, q = q0;
            // levels >= 1 carry L.apron mirrored pixels around the image in HBM (common.cuh): a window that leaves the frame
            // by no more than that reads them like ordinary pixels
            const int ap = L.apron;
            if (acc32_4 && inx >= -ap && iny >= -ap && inx + ww < L.w + ap && iny + wh < L.h + ap &&
                inx + wp + 4 <= ipitch - L.apron_l) {
                // Four adjacent pixels per step, the inner loop of the 21x21 kernel: two aligned words per image row, funnel
                // shifts to the window's byte phase (inx & 3, the same for every row: pitches are multiples of 4), dp2a on the
                // packed bytes, template read as vectors.  The bytes past the window in the last group of a row (and the
                // few the aligned words over-read, inside the row pitch) never reach the sums.
                const int sh8 = (inx & 3) * 8;
                const int wt = (w00 & 0xffff) | (w01 << 16), wb = (w10 & 0xffff) | (w11 << 16);
                const uint8_t *base = JB + (ptrdiff_t)iny * ipitch + (inx - (inx & 3));
                // Running offsets instead of a 2-D address per group (the template offset of group `it` is 4 * it: rows are padded
                // to whole groups), and TWO groups per trip with all eight image words requested before the first is used: the
                // loads come from L1 / L2 every iteration and their latency was the kernel's largest stall.
                int a1 = 0, a2 = 0, gx = gx0;
                int oj = gy0 * ipitch + 4 * gx0;
                const int sj = sgy * ipitch + 4 * sgx, wrapj = ipitch - 4 * G, n_groups = G * wh;
                auto group = [&](uint32_t i0, uint32_t i1, uint32_t z0, uint32_t z1, int it, int gxx) {
                    const uint32_t t0 = __funnelshift_r(i0, i1, sh8), t4 = i1 >> sh8;          // bytes 0..3, byte 4
                    const uint32_t u0 = __funnelshift_r(z0, z1, sh8), u4 = z1 >> sh8;
                    const uint32_t ts = __funnelshift_r(t0, t4, 8), us = __funnelshift_r(u0, u4, 8);   // bytes 1..4
                    const int v0 = dp2a_lo_su(wb, u0, dp2a_lo_su(wt, t0, 256)), v1 = dp2a_lo_su(wb, us, dp2a_lo_su(wt, ts, 256));
                    const int v2 = dp2a_hi_su(wb, u0, dp2a_hi_su(wt, t0, 256)), v3 = dp2a_hi_su(wb, us, dp2a_hi_su(wt, ts, 256));
                    const uint2 iw = *reinterpret_cast<const uint2 *>(Iw + 4 * it);
                    const uint4 dw = *reinterpret_cast<const uint4 *>(dIw + 4 * it);
                    int d0 = (v0 >> 9) - sext16(iw.x), d1 = (v1 >> 9) - ((int)iw.x >> 16);
                    int d2 = (v2 >> 9) - sext16(iw.y), d3 = (v3 >> 9) - ((int)iw.y >> 16);
                    const int valid = ww - 4 * gxx;          // >= 4 except in the last group of a row
                    if (valid < 4) { d3 = 0; if (valid < 3) d2 = 0; if (valid < 2) d1 = 0; }
                    a1 += d0 * sext16(dw.x) + d1 * sext16(dw.y) + d2 * sext16(dw.z) + d3 * sext16(dw.w);
                    a2 += d0 * ((int)dw.x >> 16) + d1 * ((int)dw.y >> 16) + d2 * ((int)dw.z >> 16) + d3 * ((int)dw.w >> 16);
                };
                int it = tid;
                for (; it + kGenThreads < n_groups; it += 2 * kGenThreads) {
                    const uint32_t *r0 = reinterpret_cast<const uint32_t *>(base + oj);
                    const uint32_t *r1 = reinterpret_cast<const uint32_t *>(base + oj + ipitch);
                    const int gxa = gx;
                    gx += sgx; oj += sj;
                    if (gx >= G) { gx -= G; oj += wrapj; }
                    const uint32_t *s0 = reinterpret_cast<const uint32_t *>(base + oj);
                    const uint32_t *s1 = reinterpret_cast<const uint32_t *>(base + oj + ipitch);
                    const int gxb = gx;
                    gx += sgx; oj += sj;
                    if (gx >= G) { gx -= G; oj += wrapj; }
                    const uint32_t a0 = __ldg(r0), a1w = __ldg(r0 + 1), b0 = __ldg(r1), b1w = __ldg(r1 + 1);
                    const uint32_t c0 = __ldg(s0), c1w = __ldg(s0 + 1), e0 = __ldg(s1), e1w = __ldg(s1 + 1);
                    group(a0, a1w, b0, b1w, it, gxa);
                    group(c0, c1w, e0, e1w, it + kGenThreads, gxb);
                }
                if (it < n_groups) {
                    const uint32_t *r0 = reinterpret_cast<const uint32_t *>(base + oj);
                    const uint32_t *r1 = reinterpret_cast<const uint32_t *>(base + oj + ipitch);
                    group(__ldg(r0), __ldg(r0 + 1), __ldg(r1), __ldg(r1 + 1), it, gx);
                }
                sb1 = a1; sb2 = a2;
            } else if (acc32 && inx >= -ap && iny >= -ap && inx + ww < L.w + ap && iny + wh < L.h + ap) {
                // window inside the (padded) frame and <= 60 pixels per thread: no border arithmetic, 32-bit partial sums
                int a1 = 0, a2 = 0;
                int oj = (iny + y) * ipitch + (inx + x);
                const int sj = step_y * ipitch + step_x;
                for (int p = tid; p < npx; p += kGenThreads) {
                    const uint8_t *r0 = JB + oj, *r1 = r0 + ipitch;
                    const int diff = ((r0[0] * w00 + r0[1] * w01 + r1[0] * w10 + r1[1] * w11 + (1 << 8)) >> 9) - Iw[q];
                    const short2 d = dIw[q];
                    a1 += diff * d.x; a2 += diff * d.y;
                    x += step_x; q += sq; oj += sj;
                    if (x >= ww) { x -= ww; q += wp - ww; oj += ipitch - ww; }
                }
                sb1 = a1; sb2 = a2;
            } else {
                for (int p = tid; p < npx; p += kGenThreads) {
                    const int diff = sample_u8(JB, ipitch, L.w, L.h, inx + x, iny + y, w00, w01, w10, w11) - Iw[q];
                    const short2 d = dIw[q];
                    sb1 += (long long)(diff * d.x); sb2 += (long long)(diff * d.y);
                    x += step_x; y += step_y; q += sq;
                    if (x >= ww) { x -= ww; y++; q += wp - ww; }
                }
            }
            { long long z = 0; block_sum3(sb1, sb2, z, red, rsel, tid); }
            if (lk_step(P, __ll2float_rn(sb1), __ll2float_rn(sb2), A11, A12, A22, Dinv, half_x, half_y, j, nx, ny, pdx, pdy, next_x, next_y)) break;
        }

        if (st && level == 0) {
            const float qx = __fsub_rn(next_x, half_x), qy = __fsub_rn(next_y, half_y);
            const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
            if (inx < -ww || inx >= L.w || iny < -wh || iny >= L.h) { st = 0; continue; }
            bilinear_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            long long se = 0;
            int x = lx0, y = ly0, q = q0;
            const int ap = L.apron;
            const bool img_direct = inx >= -ap && iny >= -ap && inx + ww < L.w + ap && iny + wh < L.h + ap;
            for (int p = tid; p < npx; p += kGenThreads) {
                int diff;
                if (img_direct) {
                    const uint8_t *r0 = JB + (iny + y) * ipitch + (inx + x), *r1 = r0 + ipitch;
                    diff = ((r0[0] * w00 + r0[1] * w01 + r1[0] * w10 + r1[1] * w11 + (1 << 8)) >> 9) - Iw[q];
                } else {
                    diff = sample_u8(JB, ipitch, L.w, L.h, inx + x, iny + y, w00, w01, w10, w11) - Iw[q];
                }
                se += diff < 0 ? -diff : diff;
                x += step_x; y += step_y; q += sq;
                if (x >= ww) { x -= ww; y++; q += wp - ww; }
            }
            { long long z0 = 0, z1 = 0; block_sum3(se, z0, z1, red, rsel, tid); }
            er = __fdiv_rn(__ll2float_rn(se), (float)(32 * ww * wh));
        }
    }
    o.x = next_x; o.y = next_y; o.st = st; o.err = st ? er : 0.f;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

__global__ void __launch_bounds__(kGenThreads)
lk_generic_kernel(const __grid_constant__ LkParams P, int max_win_px)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long red[2 * 3 * (kGenThreads / 32)];
    const int tid = threadIdx.x;
    const int i = blockIdx.x;                                   // one CTA per point: the template of a large window
    const int stride = (max_win_px + 7) & ~7;                   // (6 B per pixel) is shared by 128 threads
    short2 *dIw = reinterpret_cast<short2 *>(smem_raw);
    short *Iw = reinterpret_cast<short *>(dIw + stride);

    const TrackJob *jobs = P.jobs ? P.jobs : P.ejobs;
    const TrackJob &job = jobs[find_job(jobs, P.n_jobs, i)];
    const int k = i - job.pt_begin;
    const int t = job.target_of_pt ? job.target_of_pt[k] : 0;
    const TrackGroup g = (P.jobs ? P.groups : P.egroups)[job.group_base + min(max(t, 0), job.n_targets - 1)];
    const float2 p0 = job.pts[k];

    Solve f, b;
    b.x = b.y = b.err = 0.f; b.st = 1;
    if (g.max_level < 0) {                       // untrackable window (api.cu): the point comes back unmoved with status 0
        f.x = p0.x; f.y = p0.y; f.err = 0.f; f.st = 0;
        b.x = p0.x; b.y = p0.y;
        finish_point(P, job, k, p0, f, b, tid);
        return;
    }
    solve_generic(P, g.plane_from, g.plane_to, g.win_w, g.win_h, g.max_level, p0.x, p0.y, Iw, dIw, tid, red, f);
    // (a point the forward pass lost keeps status 0 whatever the way back says: no backward pass, fb_err = 0)
    if (P.do_fb && f.st) solve_generic(P, g.plane_to, g.plane_from, g.win_w, g.win_h, g.max_level, f.x, f.y, Iw, dIw, tid, red, b);
    else { b.x = p0.x; b.y = p0.y; }
    finish_point(P, job, k, p0, f, b, tid);
}

// ------------------------------------------------------------------------------------------------
// 21 x 21 window
// ------------------------------------------------------------------------------------------------
constexpr int kW = 21;                            // window
constexpr int kPatch = kW + 1;                    // 22 x 22 samples feed the 21 x 21 bilinear window
constexpr int kMargin = 4;                        // extra samples staged around the target-image patch
constexpr int kJPatch = kPatch + 2 * kMargin;     // 30 x 30: the window may drift 8 px inside it before a re-stage
constexpr int kU8Pitch = 12;                      // words per staged u8 row: 48 B = three 16-byte groups >= 15 + 30
constexpr int kDxPitch = 28;                      // words per staged derivative row = 7 groups of 4 pixels
// per-warp shared memory: three TMA destinations, each 128-byte aligned
constexpr int kDBytes = 2560;                     // 22 rows x 112 B = 2464, rounded up
constexpr int kIBytes = 1152;                     // 22 rows x 48 B = 1056, rounded up
constexpr int kJBytes = 1536;                     // 30 rows x 48 B = 1440, rounded up
constexpr int kWarpSmemBytes = kDBytes + kIBytes + kJBytes;    // 5248 B per warp
constexpr int kIBoxBytes = kPatch * kU8Pitch * 4, kJBoxBytes = kJPatch * kU8Pitch * 4, kDBoxBytes = kPatch * kDxPitch * 4;
#ifndef PSNW_LK_WARPS
#define PSNW_LK_WARPS 4
#endif
#ifndef PSNW_LK_MINB
#define PSNW_LK_MINB 4
#endif
constexpr int kWarpsPerCta = PSNW_LK_WARPS;

__device__ __forceinline__ int refl_clamp(int p, int n)
{
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);
}

// window origin (ix, iy) rejected by LKTrackerInvoker's bounds test: ix < -win || ix >= cols || ...
__device__ __forceinline__ bool window_oob(int ix, int iy, int w, int h)
{
    return (unsigned)(ix + kW) >= (unsigned)(w + kW) || (unsigned)(iy + kW) >= (unsigned)(h + kW);
}

// per-lane roles, fixed for the whole kernel
struct WarpRole {
    int lane;
    // the lane's two runs: run r -> (row r / 3, first column 7 * (r % 3)); lane 31 owns one
    int off0, off1;               // byte offset of the runs' first sample in a staged u8 patch: row * 48 + column
    int doff0, doff1;             // word offset of the runs' first pixel in the staged derivative patch: row * 28 + column
    bool two;
};

// Per-warp staging state.  Patches are fetched by TMA (cp.async.bulk.tensor, one instruction issued by lane 0 per patch)
// and complete on two mbarriers: `src` for the source-frame patches (I and its derivatives, prefetched one level ahead)
// and `tgt` for the target-image patch.  ph* = parity of the next completion, pend* = a completion is outstanding.
struct Stage {
    uint32_t bar_src, bar_tgt;
    uint32_t ph_src, ph_tgt;
    bool pend_src, pend_tgt;
};

__device__ __forceinline__ void stage_wait_src(Stage &S)
{
    if (S.pend_src) { mbar_wait(S.bar_src, S.ph_src); S.ph_src ^= 1u; S.pend_src = false; }
}
__device__ __forceinline__ void stage_wait_tgt(Stage &S)
{
    if (S.pend_tgt) { mbar_wait(S.bar_tgt, S.ph_tgt); S.ph_tgt ^= 1u; S.pend_tgt = false; }
}

// tensor maps of level l (LkParams::tmaps): [3l] u8 image with the 48 x 22 box, [3l+1] u8 image with the 48 x 30 box,
// [3l+2] (dx,dy) image with the 28 x 22 box.  Apron levels describe the padded plane (the apron is ordinary memory to the
// copy), level 0 the bare image; the derivative map zero-fills outside the image = the BORDER_CONSTANT apron of OpenCV.
enum { kMapI = 0, kMapJ = 1, kMapD = 2 };

// ---- staging of an N x N u8 patch with top-left sample (x0, y0) into rows of kU8Pitch words; returns the byte offset of
// sample column 0 inside a staged row.  A patch that lies inside the frame, or anywhere on a level with a REFLECT_101
// apron, is one TMA box of three aligned 16-byte groups per row (*tma_bytes receives its size; the caller has armed the
// mbarrier).  On level 0 a patch touching the border is gathered byte by byte through REFLECT_101 (plain stores).
template <int N>
__device__ __forceinline__ int stage_u8(const LkParams &P, int level, int plane, int x0, int y0, uint32_t *patch,
                                        uint32_t bar, const WarpRole &R, int *tma_bytes)
{
    const LevelGeom &L = P.lv[level];
    const bool fast = L.apron || (x0 >= 0 && y0 >= 0 && x0 + N <= L.w && y0 + N <= L.h);
    if (fast) {
        const int off = x0 & 15;
        *tma_bytes = N * kU8Pitch * 4;
        if (R.lane == 0)
            tma_load_3d(smem_u32(patch), P.tmaps + 3 * level + (N == kPatch ? kMapI : kMapJ),
                        x0 - off + L.apron_l, y0 + L.apron, plane, bar);
        return off;
    }
    *tma_bytes = 0;
    if (R.lane < N) {
        const uint8_t *col = L.img + (size_t)plane * L.iplane_px + refl_clamp(x0 + R.lane, L.w);
        uint8_t *pb = reinterpret_cast<uint8_t *>(patch) + R.lane;
#pragma unroll 2
        for (int row = 0; row < N; ++row) pb[row * (kU8Pitch * 4)] = col[(size_t)refl_clamp(y0 + row, L.h) * L.ipitch];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // these generic stores precede later TMA writes to the buffer
    return 0;
}

// Bilinear sums of a horizontal run of 7 samples starting at byte column `bc` of patch row `row`:
//   v[k] = c[k] + w_tl*p(k) + w_tr*p(k+1) + w_bl*q(k) + w_br*q(k+1),  wt/wb = (w_left | w_right << 16) of the top /
// bottom row.  The 8+8 bytes are aligned with funnel shifts and fed to dp2a without unpacking.
__device__ __forceinline__ void bilinear_run7(const uint32_t *patch, int t, int wt, int wb, const int c[7], int v[7])
{
    // t = byte offset of the first sample inside the patch (row * 48 + byte column); the funnel shift only looks at the low
    // five bits of its count, so (t << 3) is the bit offset inside the aligned word
    const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(patch) + (t & ~3));
    const int sh = t << 3;
    const uint32_t t0 = __funnelshift_r(p[0], p[1], sh), t1 = __funnelshift_r(p[1], p[2], sh);
    const uint32_t u0 = __funnelshift_r(p[kU8Pitch], p[kU8Pitch + 1], sh);
    const uint32_t u1 = __funnelshift_r(p[kU8Pitch + 1], p[kU8Pitch + 2], sh);
    const uint32_t ts = __funnelshift_r(t0, t1, 8), us = __funnelshift_r(u0, u1, 8);   // bytes 1..4
    const uint32_t t5 = t1 >> 8, u5 = u1 >> 8;                                         // bytes 5..7
    v[0] = dp2a_lo_su(wb, u0, dp2a_lo_su(wt, t0, c[0]));
    v[1] = dp2a_lo_su(wb, us, dp2a_lo_su(wt, ts, c[1]));
    v[2] = dp2a_hi_su(wb, u0, dp2a_hi_su(wt, t0, c[2]));
    v[3] = dp2a_hi_su(wb, us, dp2a_hi_su(wt, ts, c[3]));
    v[4] = dp2a_lo_su(wb, u1, dp2a_lo_su(wt, t1, c[4]));
    v[5] = dp2a_lo_su(wb, u5, dp2a_lo_su(wt, t5, c[5]));
    v[6] = dp2a_hi_su(wb, u1, dp2a_hi_su(wt, t1, c[6]));
}

// Origin of the staged 30 x 30 target patch for a window at (inx, iny): centred (margin 4 on every side), but slid
// back inside the frame when the window itself is inside, so that the aligned fast path serves it.
__device__ __forceinline__ int patch_origin(int ix, int n)
{
    int s = ix - kMargin;
    if (ix >= 0 && ix + kPatch <= n && n >= kJPatch) s = min(max(s, 0), n - kJPatch);
    return s;
}

// Source point at `level`: (sx, sy) = prevPt * 2^-level, window origin (ix, iy) = floor((sx, sy) - halfWin);
// false when LKTrackerInvoker skips the level (window outside the frame).
__device__ __forceinline__ bool src_origin(float px, float py, int level, int Lw, int Lh, float &sx, float &sy, int &ix, int &iy)
{
    constexpr float half = (kW - 1) * 0.5f;
    const float sc = __int_as_float((127 - level) << 23);
    sx = __fmul_rn(px, sc);
    sy = __fmul_rn(py, sc);
    ix = __float2int_rd(__fsub_rn(sx, half));
    iy = __float2int_rd(__fsub_rn(sy, half));
    return !window_oob(ix, iy, Lw, Lh);
}

// Stages the source-frame patches (I and its derivatives) of `level` on the `src` barrier.  Called one level ahead: by
// then the template of the current level is in registers, so the copies land while the current level iterates and the
// next level starts without waiting for memory.
__device__ __forceinline__ void prefetch_src(const LkParams &P, int plane_a, int level, float px, float py,
                                             uint32_t *ip, uint32_t *dxp, Stage &S, const WarpRole &R)
{
    if (level < 0) return;
    const LevelGeom &L = P.lv[level];
    float sx, sy;
    int ix, iy;
    if (!src_origin(px, py, level, L.w, L.h, sx, sy, ix, iy)) return;
    const bool fastI = L.apron || (ix >= 0 && iy >= 0 && ix + kPatch <= L.w && iy + kPatch <= L.h);
    if (R.lane == 0) mbar_expect_tx(S.bar_src, kDBoxBytes + (fastI ? kIBoxBytes : 0));
    int nb;
    stage_u8<kPatch>(P, level, plane_a, ix, iy, ip, S.bar_src, R, &nb);
    if (R.lane == 0) tma_load_3d(smem_u32(dxp), P.tmaps + 3 * level + kMapD, ix - (ix & 3), iy, plane_a, S.bar_src);
    S.pend_src = true;
}

// target patch at (sx0, sy0) on the `tgt` barrier; returns the byte offset of sample column 0
__device__ __forceinline__ int stage_tgt(const LkParams &P, int level, int plane_b, int sx0, int sy0, uint32_t *jp,
                                         Stage &S, const WarpRole &R)
{
    const LevelGeom &L = P.lv[level];
    const bool fast = L.apron || (sx0 >= 0 && sy0 >= 0 && sx0 + kJPatch <= L.w && sy0 + kJPatch <= L.h);
    if (fast && R.lane == 0) mbar_expect_tx(S.bar_tgt, kJBoxBytes);
    int nb;
    const int off = stage_u8<kJPatch>(P, level, plane_b, sx0, sy0, jp, S.bar_tgt, R, &nb);
    S.pend_tgt = nb != 0;
    return off;
}

// kStats (bench / inspection build of the kernel only, psnw_lk_stats): counts the template passes (levels on which the
// 2x2 system was formed) and the Gauss-Newton iterations of the solve in n_tmpl / n_iter.
template <bool kStats>
__device__ __forceinline__ void solve_w21(const LkParams &P, int plane_a, int plane_b, int max_level, float px, float py,
                                       uint32_t *ip, uint32_t *jp, uint32_t *dxp, Stage &S, const WarpRole &R, bool want_err,
                                       Solve &o, unsigned &n_tmpl, unsigned &n_iter)
{
    constexpr float half = (kW - 1) * 0.5f;
    float next_x = 0.f, next_y = 0.f, er = 0.f;
    int st = 1;

    prefetch_src(P, plane_a, max_level, px, py, ip, dxp, S, R);
#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelGeom &L = P.lv[level];
        const int Lw = L.w, Lh = L.h;
        stage_wait_src(S);                       // this level's I / dI patches have landed
        __syncwarp();

        float prev_x, prev_y, nx, ny;
        int ipx, ipy;
        const bool src_ok = src_origin(px, py, level, Lw, Lh, prev_x, prev_y, ipx, ipy);
        if (level == max_level) { nx = prev_x; ny = prev_y; }
        else                    { nx = __fmul_rn(next_x, 2.f); ny = __fmul_rn(next_y, 2.f); }
        next_x = nx; next_y = ny;
        if (!src_ok) {
            if (level == 0) { st = 0; er = 0.f; }
            prefetch_src(P, plane_a, level - 1, px, py, ip, dxp, S, R);
            continue;
        }
        prev_x = __fsub_rn(prev_x, half); prev_y = __fsub_rn(prev_y, half);
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(prev_x, (float)ipx), __fsub_rn(prev_y, (float)ipy), w00, w01, w10, w11);
        const bool fastI = L.apron || (ipx >= 0 && ipy >= 0 && ipx + kPatch <= Lw && ipy + kPatch <= Lh);
        const int offI = fastI ? (ipx & 15) : 0, offD = ipx & 3;

        // ---- target patch around the first position: in flight while the template is computed ----
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        int sx0 = 0, sy0 = 0, offJ = 0;
        bool staged = !window_oob(__float2int_rd(nx), __float2int_rd(ny), Lw, Lh);
        if (staged) {
            sx0 = patch_origin(__float2int_rd(nx), Lw); sy0 = patch_origin(__float2int_rd(ny), Lh);
            offJ = stage_tgt(P, level, plane_b, sx0, sy0, jp, S, R);
        }

        // ---- template into registers ----
        int tC[14], tIx[14], tIy[14];          // tC = 256 - 512 * I: the dp2a chain then yields (J - I) << 9 directly
        int a11[2] = {0, 0}, a12[2] = {0, 0}, a22[2] = {0, 0};     // one accumulator chain per run: more ILP
        {
            const int wt = (w00 & 0xffff) | (w01 << 16), wb = (w10 & 0xffff) | (w11 << 16);
            const int c256[7] = {256, 256, 256, 256, 256, 256, 256};
#pragma unroll
            for (int run = 0; run < 2; ++run) {
                int v[7];
                bilinear_run7(ip, (run ? R.off1 : R.off0) + offI, wt, wb, c256, v);
                const uint32_t *d0 = dxp + (run ? R.doff1 : R.doff0) + offD, *d1 = d0 + kDxPitch;
                uint32_t q0 = d0[0], q1 = d1[0];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const uint32_t n0 = d0[k + 1], n1 = d1[k + 1];
                    const int ix = (sext16(q0) * w00 + sext16(n0) * w01 + sext16(q1) * w10 + sext16(n1) * w11 + (1 << 13)) >> 14;
                    const int iy = (((int)q0 >> 16) * w00 + ((int)n0 >> 16) * w01 + ((int)q1 >> 16) * w10 + ((int)n1 >> 16) * w11 + (1 << 13)) >> 14;
                    tC[run * 7 + k] = 256 - (v[k] & ~511);      // 256 - 512 * I
                    tIx[run * 7 + k] = ix;
                    tIy[run * 7 + k] = iy;
                    a11[run] += ix * ix; a12[run] += ix * iy; a22[run] += iy * iy;
                    q0 = n0; q1 = n1;
                }
            }
            // lane 31 owns a single run: its second one re-reads run 62 and is dropped from every sum
            if (!R.two) { a11[1] = 0; a12[1] = 0; a22[1] = 0; }
        }
        if (kStats) ++n_tmpl;
        __syncwarp();                            // every lane is done with the I / dI patches ...
        prefetch_src(P, plane_a, level - 1, px, py, ip, dxp, S, R);   // ... so the next level's copies may overwrite them
        stage_wait_tgt(S);                       // the target patch has landed
        __syncwarp();
        const float s11 = sum_to_float(warp_sum_split(a11[0] + a11[1])), s12 = sum_to_float(warp_sum_split(a12[0] + a12[1])),
                    s22 = sum_to_float(warp_sum_split(a22[0] + a22[1]));
        float A11, A12, A22, Dinv;
        if (!make_system(s11, s12, s22, kW, kW, P.min_eig, A11, A12, A22, Dinv)) {
            if (level == 0) st = 0;
            continue;
        }

        float pdx = 0.f, pdy = 0.f;
#pragma unroll 1
        for (int j = 0; j < P.max_iters; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (window_oob(inx, iny, Lw, Lh)) {
                if (level == 0) st = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const int jt = (w00 & 0xffff) | (w01 << 16), jb = (w10 & 0xffff) | (w11 << 16);
            if (kStats) ++n_iter;
            if (!staged || (unsigned)(inx - sx0) > 2u * kMargin || (unsigned)(iny - sy0) > 2u * kMargin) {
                sx0 = patch_origin(inx, Lw); sy0 = patch_origin(iny, Lh);
                __syncwarp();
                offJ = stage_tgt(P, level, plane_b, sx0, sy0, jp, S, R);
                stage_wait_tgt(S);
                __syncwarp();
                staged = true;
            }
            const int tj = (iny - sy0) * (kU8Pitch * 4) + offJ + (inx - sx0);
            int b1[2] = {0, 0}, b2[2] = {0, 0};
#pragma unroll
            for (int run = 0; run < 2; ++run) {
                int v[7];
                bilinear_run7(jp, tj + (run ? R.off1 : R.off0), jt, jb, &tC[run * 7], v);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int diff = v[k] >> 9;
                    b1[run] += diff * tIx[run * 7 + k];
                    b2[run] += diff * tIy[run * 7 + k];
                }
            }
            if (!R.two) { b1[1] = 0; b2[1] = 0; }
            // (a vote-guarded single REDUX per sum -- all partials below 2^26 -- was measured: 0.2555 against 0.2513 ms per c2
            // launch, the vote and the branch cost more than the second REDUX and conversion)
            const float sb1 = sum_to_float(warp_sum_split(b1[0] + b1[1])), sb2 = sum_to_float(warp_sum_split(b2[0] + b2[1]));
            if (lk_step(P, sb1, sb2, A11, A12, A22, Dinv, half, half, j, nx, ny, pdx, pdy, next_x, next_y)) break;
        }

        if (st && level == 0) {
            const float qx = __fsub_rn(next_x, half), qy = __fsub_rn(next_y, half);
            const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
            if (window_oob(inx, iny, Lw, Lh)) { st = 0; continue; }
            if (!want_err) continue;             // the backward pass of the FB check only needs position and status
            bilinear_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            const int jt = (w00 & 0xffff) | (w01 << 16), jb = (w10 & 0xffff) | (w11 << 16);
            if (!staged || (unsigned)(inx - sx0) > 2u * kMargin || (unsigned)(iny - sy0) > 2u * kMargin) {
                sx0 = patch_origin(inx, Lw); sy0 = patch_origin(iny, Lh);
                __syncwarp();
                offJ = stage_tgt(P, level, plane_b, sx0, sy0, jp, S, R);
                stage_wait_tgt(S);
                __syncwarp();
            }
            const int tj = (iny - sy0) * (kU8Pitch * 4) + offJ + (inx - sx0);
            int e = 0;
#pragma unroll
            for (int run = 0; run < 2; ++run) {
                int v[7];
                bilinear_run7(jp, tj + (run ? R.off1 : R.off0), jt, jb, &tC[run * 7], v);
                if (run == 0 || R.two) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const int diff = v[k] >> 9;
                        e += diff < 0 ? -diff : diff;
                    }
                }
            }
            er = __fdiv_rn(sum_to_float(warp_sum_split(e)), (float)(32 * kW * kW));
        }
    }
    stage_wait_src(S);                           // (a skipped level may leave a prefetch outstanding)
    stage_wait_tgt(S);
    __syncwarp();                                // the patches are free for the next solve
    o.x = next_x; o.y = next_y; o.st = st; o.err = st ? er : 0.f;
}

// Persistent warps: every warp draws point indices from a device counter until the batch is exhausted, so a slow point
// (many iterations, re-staging) never holds three idle warps of its CTA and the last wave has no tail of empty slots.
// The last warp to finish resets the two counters, which keeps launches on one context self-contained.
template <bool kStats>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PSNW_LK_MINB)
lk_w21_kernel(const __grid_constant__ LkParams P)
{
    __shared__ __align__(128) uint8_t smem[kWarpsPerCta * kWarpSmemBytes];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *dxp = reinterpret_cast<uint32_t *>(smem + wib * kWarpSmemBytes);
    uint32_t *ip = dxp + kDBytes / 4, *jp = ip + kIBytes / 4;

    WarpRole R;
    R.lane = lane;
    R.two = lane < 31;
    {
        const int r0 = lane, r1 = R.two ? lane + 32 : 62;
        const int row0 = r0 / 3, col0 = 7 * (r0 - 3 * row0), row1 = r1 / 3, col1 = 7 * (r1 - 3 * row1);
        R.off0 = row0 * (kU8Pitch * 4) + col0; R.off1 = row1 * (kU8Pitch * 4) + col1;
        R.doff0 = row0 * kDxPitch + col0; R.doff1 = row1 * kDxPitch + col1;
    }
    Stage S;
    S.bar_src = smem_u32(&bars[wib][0]); S.bar_tgt = smem_u32(&bars[wib][1]);
    S.ph_src = S.ph_tgt = 0u;
    S.pend_src = S.pend_tgt = false;
    if (lane == 0) { mbar_init(S.bar_src, 1); mbar_init(S.bar_tgt, 1); }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (launched with programmatic stream serialisation, launch_lk)

    // Work items: point i < split_from is one item (forward + backward pass); the points of the tail are two items each --
    // forward passes first, [split_from, n_pts), then their backward passes, [n_pts, n_pts + n_tail) -- so that the last
    // items of the launch are half as long and the grid drains sooner.  A forward item hands its result over through the
    // point's output slots and a flag (1 + status); items are drawn in order and a forward item never waits, so the
    // backward item that spins on the flag always makes progress.
    const bool persistent = P.work != nullptr;
    const int n_items = P.n_pts + (P.n_pts - P.split_from);
    for (int it = 0;; ++it) {
        int i = blockIdx.x * kWarpsPerCta + wib;
        if (persistent) {
            if (lane == 0) i = (int)atomicAdd(P.work, 1u);
            i = __shfl_sync(0xffffffffu, i, 0);
        } else if (it) {
            break;                               // one point per warp: the hardware CTA scheduler balances the load
        }
        if (i >= n_items) break;
        const bool bwd_only = i >= P.n_pts;
        if (bwd_only) i = P.split_from + (i - P.n_pts);
        const bool fwd_only = !bwd_only && i >= P.split_from;

        const TrackJob *jobs = P.jobs ? P.jobs : P.ejobs;
        const TrackJob &job = jobs[find_job(jobs, P.n_jobs, i)];
        const int k = i - job.pt_begin;
        const int t = job.target_of_pt ? job.target_of_pt[k] : 0;
        const TrackGroup g = (P.jobs ? P.groups : P.egroups)[job.group_base + min(max(t, 0), job.n_targets - 1)];
        const float2 p0 = job.pts[k];

        // forward pass, then (fused forward-backward check) the same solver from the tracked position back into the
        // source frame; one copy of the solver code keeps the kernel inside the instruction cache
        Solve f, b, cur;
        f.x = f.y = f.err = 0.f; f.st = 1;
        b = f;
        unsigned *flag = P.work + 2 + (i - P.split_from);
        if (bwd_only) {
            unsigned v = 0;
            if (lane == 0) {
                while ((v = ld_acquire(flag)) == 0u) __nanosleep(100);
                *flag = 0u;                      // self-cleaning: the next launch finds zeros
                const float2 q = __ldcg(&job.out[k]);
                f.x = q.x; f.y = q.y; f.err = __ldcg(&job.err[k]);
            }
            f.st = (int)__shfl_sync(0xffffffffu, v, 0) - 1;
            f.x = __shfl_sync(0xffffffffu, f.x, 0); f.y = __shfl_sync(0xffffffffu, f.y, 0);
        }
        if (g.max_level < 0) {                   // untrackable window (api.cu; no split tail then): unmoved, status 0
            f.x = p0.x; f.y = p0.y; f.st = 0;
            b.x = p0.x; b.y = p0.y;
            finish_point(P, job, k, p0, f, b, lane);
            continue;
        }
        const int first = bwd_only ? 1 : 0, last = (P.do_fb && !fwd_only) ? 2 : 1;
        b.x = p0.x; b.y = p0.y;                  // without a backward pass fb_err = 0
#pragma unroll 1
        for (int pass = first; pass < last; ++pass) {
            // a point the forward pass lost keeps status 0 whatever the way back says: the backward pass is not run for it
            // (7 % of the seeds of the c2 rig, 20 % at c3) and fb_err is defined as 0 (oracle: psnwo_fb_check)
            if (pass && !f.st) break;
            const float sx = pass ? f.x : p0.x, sy = pass ? f.y : p0.y;
            unsigned n_tmpl = 0, n_iter = 0;
            solve_w21<kStats>(P, pass ? g.plane_to : g.plane_from, pass ? g.plane_from : g.plane_to, g.max_level, sx, sy, ip, jp, dxp, S, R, pass == 0, cur, n_tmpl, n_iter);
            if (pass) b = cur; else f = cur;
            if (kStats && lane == 0) {
                atomicAdd(P.stats, 1ull);
                atomicAdd(P.stats + 1, (unsigned long long)n_tmpl);
                atomicAdd(P.stats + 2, (unsigned long long)n_iter);
            }
        }
        if (fwd_only) {
            if (lane == 0) {
                job.out[k] = make_float2(f.x, f.y);
                job.err[k] = f.err;
                __threadfence();
                st_release(flag, 1u + (unsigned)f.st);
            }
        } else {
            finish_point(P, job, k, p0, f, b, lane);
        }
    }
    if (persistent && lane == 0) {
        const unsigned total = gridDim.x * kWarpsPerCta;
        if (atomicAdd(P.work + 1, 1u) + 1 == total) { P.work[0] = 0; P.work[1] = 0; }
    }
}

}  // namespace

int configure_lk_kernels()
{
    return cudaFuncSetAttribute(lk_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess ? 0 : -1;
}

// CTAs of the persistent 21x21 kernel that are resident at once on the current device (SM count x occupancy)
int lk_w21_resident_ctas(int max_per_sm)
{
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lk_w21_kernel<false>, kWarpsPerCta * 32, 0) != cudaSuccess)
        return -1;
    if (max_per_sm > 0) per_sm = min(per_sm, max_per_sm);
    return sms * per_sm;
}

int launch_lk(const LkParams &P, int all_21x21, int max_win_px, int resident_ctas, cudaStream_t stream)
{
    if (P.n_pts <= 0) return 0;
    if (all_21x21) {
        int grid = (P.n_pts + kWarpsPerCta - 1) / kWarpsPerCta;
        if (P.work) grid = min(grid, max(resident_ctas, 1));
        // programmatic dependent launch: the grid is set up (and its prologue runs) under the tail of the kernel before it in
        // the stream; the kernel waits for that kernel's completion before its first global access
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kWarpsPerCta * 32);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = P.stats ? cudaLaunchKernelEx(&cfg, lk_w21_kernel<true>, P) : cudaLaunchKernelEx(&cfg, lk_w21_kernel<false>, P);
        return e == cudaSuccess ? 1 : -1;
    } else {
        const int stride = (max_win_px + 7) & ~7;
        const size_t smem = (size_t)stride * 6;
        if (smem > 200 * 1024) return -2;
        lk_generic_kernel<<<P.n_pts, kGenThreads, smem, stream>>>(P, max_win_px);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace psnw
