/*
 * klt_oracle.c -- CPU restatement of the Tracker2D KLT hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library.  Nothing under mcmtt_opticalflow_b200/ links, imports or calls it.
 *
 * What it restates
 * ----------------
 * The reference (blackpigg/MCMTT_OPTICALFLOW, "PSN_Where") runs this path on the CPU through OpenCV
 * 2.4.6, which is NOT vendored in /root/reference (psn_where/PSN_Where.vcxproj:97,104 only name the
 * import libs).  The arithmetic below is therefore a restatement of OpenCV's published algorithm
 * (video/lkpyramid.cpp, imgproc/pyramids.cpp, imgproc/corner.cpp, imgproc/featureselect.cpp) anchored on
 * the reference's call sites:
 *   - cv::calcOpticalFlowPyrLK        psn_where/PSNWhere_Tracker2D.cpp:776-782 (backward), :871-877 (forward)
 *   - LocalSearchKLT (reference code) psn_where/PSNWhere_Tracker2D.cpp:452-554
 *   - feature detection call site     psn_where/PSNWhere_Tracker2D.cpp:735-757 (north_star: Shi-Tomasi)
 *   - PSN_Rect / PSN_Point2D          psn_where/PSNWhere_Types.h:18-66, :112-182
 *
 * Pinning status
 * --------------
 * The reference has no tests, golden vectors or fixtures for this path, and the
 * reference itself cannot be compiled here (MSVC/MFC/MSXML/OpenCV-2.4.6 SDK): "parity unpinned" by the
 * reference's OWN artefacts.  The restatement is instead pinned against the only runnable
 * implementation of the third-party arithmetic, python cv2 4.13.0 (tests/test_oracle_vs_cv2.py and the
 * committed fixtures tests/golden/ made by oracle/gen_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; float expressions must stay unfused).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define PSNWO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* helpers                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else       p = 2 * (len - 1) - p;
    }
    return p;
}

/* cvRound: round half to even (SSE cvtss2si / lrint under the default rounding mode) */
static int cv_round(float v) { return (int)lrintf(v); }
static int cv_floor(float v) { int i = (int)v; return i - (v < (float)i); }

/* ------------------------------------------------------------------------------------------------ */
/* A.1  pyrDown  (imgproc/pyramids.cpp; used by buildOpticalFlowPyramid inside calcOpticalFlowPyrLK, */
/*      call sites PSNWhere_Tracker2D.cpp:776,:871)                                                  */
/*      5x5 [1 4 6 4 1]^2, REFLECT_101, sample even coords, (sum + 128) >> 8, dst = ((w+1)/2,(h+1)/2) */
/* ------------------------------------------------------------------------------------------------ */
PSNWO_API void psnwo_pyr_down(const uint8_t *src, int w, int h, int sstride,
                              uint8_t *dst, int dstride)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int *rows = (int *)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const uint8_t *s = src + (size_t)reflect101(2 * y - 2 + k, h) * sstride;
            int *r = rows + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                int x0 = reflect101(2 * x - 2, w), x1 = reflect101(2 * x - 1, w), x2 = 2 * x,
                    x3 = reflect101(2 * x + 1, w), x4 = reflect101(2 * x + 2, w);
                r[x] = s[x0] + 4 * s[x1] + 6 * s[x2] + 4 * s[x3] + s[x4];
            }
        }
        for (int x = 0; x < dw; x++) {
            int v = rows[x] + 4 * rows[dw + x] + 6 * rows[2 * dw + x] + 4 * rows[3 * dw + x] + rows[4 * dw + x];
            dst[(size_t)y * dstride + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rows);
}

/* ------------------------------------------------------------------------------------------------ */
/* A.2  Scharr derivative (calcSharrDeriv in video/lkpyramid.cpp): interleaved int16 (dx,dy),        */
/*      unnormalised 3/10/3 kernel, REFLECT_101 at the image edge.                                   */
/* ------------------------------------------------------------------------------------------------ */
PSNWO_API void psnwo_scharr(const uint8_t *src, int w, int h, int sstride,
                            int16_t *dxy, int dstride_px)
{
    int *t0 = (int *)malloc(sizeof(int) * (size_t)w * 2), *t1 = t0 + w;
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + (size_t)reflect101(y - 1, h) * sstride;
        const uint8_t *r1 = src + (size_t)y * sstride;
        const uint8_t *r2 = src + (size_t)reflect101(y + 1, h) * sstride;
        for (int x = 0; x < w; x++) {
            t0[x] = 3 * (r0[x] + r2[x]) + 10 * r1[x];
            t1[x] = r2[x] - r0[x];
        }
        int16_t *d = dxy + (size_t)y * dstride_px * 2;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            d[2 * x]     = (int16_t)(t0[xp] - t0[xm]);
            d[2 * x + 1] = (int16_t)(3 * (t1[xm] + t1[xp]) + 10 * t1[x]);
        }
    }
    free(t0);
}

/* ------------------------------------------------------------------------------------------------ */
/* A.3  Sparse pyramidal LK, one direction (LKTrackerInvoker in video/lkpyramid.cpp).                */
/*      Reference call sites: PSNWhere_Tracker2D.cpp:776-782 and :871-877 (defaults: maxLevel 3,     */
/*      30 iterations, eps 0.01, flags 0, minEigThreshold 1e-4; window passed per target).           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *img;   /* u8 level */
    const int16_t *dxy;   /* interleaved (dx,dy) of the same level, may be NULL for the 'to' frame */
    int w, h;
    int img_stride;       /* bytes */
    int dxy_stride_px;    /* pixels (2 int16 each) */
} psnwo_level;

/* image read with the REFLECT_101 apron OpenCV keeps around each level */
static inline int img_at(const psnwo_level *L, int x, int y)
{
    return L->img[(size_t)reflect101(y, L->h) * L->img_stride + reflect101(x, L->w)];
}
/* derivative read: the apron is BORDER_CONSTANT 0 */
static inline void dxy_at(const psnwo_level *L, int x, int y, int *dx, int *dy)
{
    if (x < 0 || y < 0 || x >= L->w || y >= L->h) { *dx = 0; *dy = 0; return; }
    const int16_t *p = L->dxy + ((size_t)y * L->dxy_stride_px + x) * 2;
    *dx = p[0]; *dy = p[1];
}

#define W_BITS 14
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static void bilinear_weights(float a, float b, int *w00, int *w01, int *w10, int *w11)
{
    *w00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
    *w01 = cv_round(a * (1.f - b) * (1 << W_BITS));
    *w10 = cv_round((1.f - a) * b * (1 << W_BITS));
    *w11 = (1 << W_BITS) - *w00 - *w01 - *w10;
}

/* number of the last pyramid level OpenCV keeps: levels are added while both dims stay > win */
PSNWO_API int psnwo_effective_max_level(int w, int h, int win_w, int win_h, int max_level)
{
    int level = 0;
    while (level < max_level) {
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= win_w || h <= win_h) break;
        level++;
    }
    return level;
}

/*
 * exact_sums = 0: float accumulation in raster order (OpenCV's scalar path).
 * exact_sums = 1: integer-exact sums rounded to float once (what the CUDA kernel does; differs from
 *                 mode 0 only by OpenCV's own accumulation rounding, ~1e-4 px).
 */
PSNWO_API void psnwo_lk_track(const psnwo_level *from, const psnwo_level *to, int max_level,
                              const float *pts, int n, int win_w, int win_h,
                              int max_iters, double epsilon, float min_eig_thresh, int exact_sums,
                              float *out, uint8_t *status, float *err)
{
    const float FLT_SCALE = 1.f / (1 << 20);
    const float half_x = (win_w - 1) * 0.5f, half_y = (win_h - 1) * 0.5f;
    if (max_iters < 0) max_iters = 0;
    if (max_iters > 100) max_iters = 100;
    if (epsilon < 0) epsilon = 0;
    if (epsilon > 10) epsilon = 10;
    const double eps2 = epsilon * epsilon;
    const size_t npx = (size_t)win_w * win_h;
    int16_t *Iw = (int16_t *)malloc(sizeof(int16_t) * npx * 3), *dIw = Iw + npx;

    for (int i = 0; i < n; i++) {
        uint8_t st = 1;
        float er = 0.f;
        float next_x = 0.f, next_y = 0.f;      /* nextPts[i], carried across levels */
        for (int level = max_level; level >= 0; level--) {
            const psnwo_level *I = &from[level], *J = &to[level];
            const float sc = (float)(1. / (1 << level));
            float prev_x = pts[2 * i] * sc, prev_y = pts[2 * i + 1] * sc;
            float nx, ny;
            if (level == max_level) { nx = prev_x; ny = prev_y; }
            else                    { nx = next_x * 2.f; ny = next_y * 2.f; }
            next_x = nx; next_y = ny;

            prev_x -= half_x; prev_y -= half_y;
            int ipx = cv_floor(prev_x), ipy = cv_floor(prev_y);
            if (ipx < -win_w || ipx >= I->w || ipy < -win_h || ipy >= I->h) {
                if (level == 0) { st = 0; er = 0.f; }
                continue;
            }
            float a = prev_x - ipx, b = prev_y - ipy;
            int w00, w01, w10, w11;
            bilinear_weights(a, b, &w00, &w01, &w10, &w11);

            float fA11 = 0, fA12 = 0, fA22 = 0;
            int64_t iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win_h; y++)
                for (int x = 0; x < win_w; x++) {
                    int X = ipx + x, Y = ipy + y;
                    int ival = DESCALE(img_at(I, X, Y) * w00 + img_at(I, X + 1, Y) * w01 +
                                       img_at(I, X, Y + 1) * w10 + img_at(I, X + 1, Y + 1) * w11, W_BITS - 5);
                    int dx00, dy00, dx01, dy01, dx10, dy10, dx11, dy11;
                    dxy_at(I, X, Y, &dx00, &dy00);         dxy_at(I, X + 1, Y, &dx01, &dy01);
                    dxy_at(I, X, Y + 1, &dx10, &dy10);     dxy_at(I, X + 1, Y + 1, &dx11, &dy11);
                    int ixval = DESCALE(dx00 * w00 + dx01 * w01 + dx10 * w10 + dx11 * w11, W_BITS);
                    int iyval = DESCALE(dy00 * w00 + dy01 * w01 + dy10 * w10 + dy11 * w11, W_BITS);
                    size_t p = (size_t)y * win_w + x;
                    Iw[p] = (int16_t)ival; dIw[2 * p] = (int16_t)ixval; dIw[2 * p + 1] = (int16_t)iyval;
                    fA11 += (float)(ixval * ixval); fA12 += (float)(ixval * iyval); fA22 += (float)(iyval * iyval);
                    iA11 += (int64_t)ixval * ixval; iA12 += (int64_t)ixval * iyval; iA22 += (int64_t)iyval * iyval;
                }
            float A11, A12, A22;
            if (exact_sums) { A11 = (float)iA11 * FLT_SCALE; A12 = (float)iA12 * FLT_SCALE; A22 = (float)iA22 * FLT_SCALE; }
            else            { A11 = fA11 * FLT_SCALE;        A12 = fA12 * FLT_SCALE;        A22 = fA22 * FLT_SCALE; }

            float D = A11 * A22 - A12 * A12;
            float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win_w * win_h);
            if (min_eig < min_eig_thresh || D < FLT_EPSILON) {
                if (level == 0) st = 0;
                continue;
            }
            D = 1.f / D;

            nx -= half_x; ny -= half_y;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < max_iters; j++) {
                int inx = cv_floor(nx), iny = cv_floor(ny);
                if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) {
                    if (level == 0) st = 0;
                    break;
                }
                a = nx - inx; b = ny - iny;
                bilinear_weights(a, b, &w00, &w01, &w10, &w11);
                float fb1 = 0, fb2 = 0;
                int64_t ib1 = 0, ib2 = 0;
                for (int y = 0; y < win_h; y++)
                    for (int x = 0; x < win_w; x++) {
                        int X = inx + x, Y = iny + y;
                        size_t p = (size_t)y * win_w + x;
                        int diff = DESCALE(img_at(J, X, Y) * w00 + img_at(J, X + 1, Y) * w01 +
                                           img_at(J, X, Y + 1) * w10 + img_at(J, X + 1, Y + 1) * w11, W_BITS - 5) - Iw[p];
                        fb1 += (float)(diff * dIw[2 * p]); fb2 += (float)(diff * dIw[2 * p + 1]);
                        ib1 += (int64_t)diff * dIw[2 * p]; ib2 += (int64_t)diff * dIw[2 * p + 1];
                    }
                float b1, b2;
                if (exact_sums) { b1 = (float)ib1 * FLT_SCALE; b2 = (float)ib2 * FLT_SCALE; }
                else            { b1 = fb1 * FLT_SCALE;        b2 = fb2 * FLT_SCALE; }
                float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
                nx += dx; ny += dy;
                next_x = nx + half_x; next_y = ny + half_y;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_x -= dx * 0.5f; next_y -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }

            if (st && level == 0) {
                float px = next_x - half_x, py = next_y - half_y;
                int inx = cv_floor(px), iny = cv_floor(py);
                if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) { st = 0; continue; }
                a = px - inx; b = py - iny;
                bilinear_weights(a, b, &w00, &w01, &w10, &w11);
                float errval = 0.f;
                int64_t ierr = 0;
                for (int y = 0; y < win_h; y++)
                    for (int x = 0; x < win_w; x++) {
                        int X = inx + x, Y = iny + y;
                        int diff = DESCALE(img_at(J, X, Y) * w00 + img_at(J, X + 1, Y) * w01 +
                                           img_at(J, X, Y + 1) * w10 + img_at(J, X + 1, Y + 1) * w11, W_BITS - 5)
                                   - Iw[(size_t)y * win_w + x];
                        errval += fabsf((float)diff);
                        ierr += diff < 0 ? -diff : diff;
                    }
                if (exact_sums) errval = (float)ierr;
                er = errval * 1.f / (float)(32 * win_w * win_h);
            }
        }
        out[2 * i] = next_x; out[2 * i + 1] = next_y;
        status[i] = st;
        err[i] = st ? er : 0.f;      /* OpenCV leaves err undefined for failed points; pinned to 0 here */
    }
    free(Iw);
}

/* ------------------------------------------------------------------------------------------------ */
/* a5  forward-backward check (north_star definition)                     */
/*     status = st_f & st_b & (||p - p''|| <= thr); thr < 0 disables the distance gate               */
/* ------------------------------------------------------------------------------------------------ */
PSNWO_API void psnwo_fb_check(const float *p, const float *p2, const uint8_t *st_f, const uint8_t *st_b,
                              int n, float thr, uint8_t *status, float *fb_err)
{
    for (int i = 0; i < n; i++) {
        float dx = p[2 * i] - p2[2 * i], dy = p[2 * i + 1] - p2[2 * i + 1];
        float e = sqrtf(dx * dx + dy * dy);
        /* a point the forward pass lost has status 0 whatever the backward pass says: the backward pass is not run for it
           and fb_err is defined as 0 (north_star defines the check; the reference has none) */
        if (!st_f[i]) e = 0.f;
        fb_err[i] = e;
        status[i] = (uint8_t)(st_f[i] && st_b[i] && (thr < 0.f || e <= thr));
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* f-3  optical-flow majority check  (reference code, psn_where/PSNWhere_Tracker2D.cpp:931-944, :981-1022) */
/* Literal restatement on the feature lists: featuresInDetectionBox[d] receives tracker index t once per   */
/* tracked point of t inside detection box d, for the trackers that reach the comparison (ok[t]: >= 4       */
/* status-ok pairs, :889) and whose new box overlaps the detection box (:937); then the run-length walk.    */
/* counts[d][t] (may be NULL) = entries of t in the list of d; wipe[d] = 1 where :1018-1021 resets the row. */
/* ------------------------------------------------------------------------------------------------ */
static int rect_overlap(const double *a, const double *b)                           /* PSN_Rect::overlap, PSNWhere_Types.h:160-163 */
{
    double ax1 = a[0] + a[2], bx1 = b[0] + b[2], ay1 = a[1] + a[3], by1 = b[1] + b[3];
    double mx = ax1 > bx1 ? ax1 : bx1, nx = a[0] < b[0] ? a[0] : b[0];
    double my = ay1 > by1 ? ay1 : by1, ny = a[1] < b[1] ? a[1] : b[1];
    return (mx - nx < a[2] + b[2]) && (my - ny < a[3] + b[3]);
}

PSNWO_API void psnwo_flow_majority(const float *pts, const int *offsets, int n_trk, const double *trk_boxes, const uint8_t *ok,
                                   const double *det_boxes, int n_det, double ratio, int *counts, uint8_t *wipe)
{
    int total_pts = offsets[n_trk];
    int *list = (int *)malloc(sizeof(int) * (size_t)(total_pts > 0 ? total_pts : 1));
    for (int d = 0; d < n_det; d++) {
        const double *db = det_boxes + 4 * d;
        int n = 0;
        for (int t = 0; t < n_trk; t++) {                                           /* tracker loop :860, detection loop :934 */
            int c = 0;
            if (ok[t] && rect_overlap(trk_boxes + 4 * t, db)) {                     /* :889, :937 */
                for (int i = offsets[t]; i < offsets[t + 1]; i++) {                 /* :940-944, PSN_Rect::contain(cv::Point2f) */
                    double x = (double)pts[2 * i], y = (double)pts[2 * i + 1];
                    if (x >= db[0] && x < db[0] + db[2] && y >= db[1] && y < db[1] + db[3]) { list[n++] = t; c++; }
                }
            }
            if (counts) counts[(size_t)d * n_trk + t] = c;
        }
        wipe[d] = 0;
        if (n == 0) continue;                                                       /* :985 */
        int numMajor = 0, numCur = 0, curIdx = list[0];                             /* :987-990 */
        for (int f = 0; f < n; f++) {                                               /* :991-1006 */
            if (curIdx == list[f]) { numCur++; continue; }
            if (numCur > numMajor) numMajor = numCur;
            curIdx = list[f];
            numCur = 0;
        }
        if (list[0] == curIdx) numMajor = numCur;                                   /* :1009-1012 sole tracker */
        if ((double)numMajor > (double)n * ratio) continue;                         /* :1014-1015 */
        wipe[d] = 1;                                                                /* :1017-1020 */
    }
    free(list);
}

/* ------------------------------------------------------------------------------------------------ */
/* a6  LocalSearchKLT  (reference code, psn_where/PSNWhere_Tracker2D.cpp:452-554)                    */
/*     box = {x,y,w,h} doubles (PSN_Rect, PSNWhere_Types.h:112-122).  Returns the inlier count and    */
/*     writes ascending original indices to inlier_idx.  `valid` (may be NULL) restates the forward   */
/*     caller's status filter (:883-888); the backward caller passes every point (:787).              */
/* ------------------------------------------------------------------------------------------------ */
/* ---- PSNW_BOX_MEDIANFLOW (include/psnw_klt.h): median displacement + median pairwise-distance ratio.  Not reference code
 * (the reference runs LocalSearchKLT, below): this restatement IS the definition the CUDA kernel is checked against.  Medians
 * are order statistics of the doubles under the total order of their IEEE bit patterns (-0 < +0), so that the result does
 * not depend on a sort's handling of equal values. ---- */
static unsigned long long mf_key(double v)
{
    unsigned long long b;
    memcpy(&b, &v, 8);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static double mf_val(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}
static int cmp_u64(const void *a, const void *b)
{
    const unsigned long long x = *(const unsigned long long *)a, y = *(const unsigned long long *)b;
    return (x > y) - (x < y);
}
static double mf_median(unsigned long long *keys, int n)        /* sorts keys */
{
    qsort(keys, (size_t)n, sizeof(unsigned long long), cmp_u64);
    if (n & 1) return mf_val(keys[n >> 1]);
    return (mf_val(keys[(n >> 1) - 1]) + mf_val(keys[n >> 1])) * 0.5;
}

PSNWO_API int psnwo_box_medianflow(const double *box_in, const float *pre, const float *cur, const uint8_t *valid, int n_all,
                                   double *box_out, uint8_t *inlier_mask)
{
    enum { PAIR_PTS = 128 };
    int n = 0, inl = 0;
    int *vidx = (int *)malloc(sizeof(int) * (size_t)(n_all + 1));
    unsigned long long *keys = (unsigned long long *)malloc(sizeof(unsigned long long) * (size_t)(n_all + PAIR_PTS * PAIR_PTS));
    double *res = (double *)malloc(sizeof(double) * (size_t)(n_all + 1));
    for (int i = 0; i < n_all; ++i) {
        inlier_mask[i] = 0;
        if (!valid || valid[i]) vidx[n++] = i;
    }
    memcpy(box_out, box_in, sizeof(double) * 4);
    if (n == 0) { free(vidx); free(keys); free(res); return 0; }
    for (int j = 0; j < n; ++j) keys[j] = mf_key((double)cur[2 * vidx[j]] - (double)pre[2 * vidx[j]]);
    const double mdx = mf_median(keys, n);
    for (int j = 0; j < n; ++j) keys[j] = mf_key((double)cur[2 * vidx[j] + 1] - (double)pre[2 * vidx[j] + 1]);
    const double mdy = mf_median(keys, n);
    for (int j = 0; j < n; ++j) {
        const int i = vidx[j];
        const double rx = ((double)cur[2 * i] - (double)pre[2 * i]) - mdx, ry = ((double)cur[2 * i + 1] - (double)pre[2 * i + 1]) - mdy;
        res[j] = sqrt(rx * rx + ry * ry);
        keys[j] = mf_key(res[j]);
    }
    const double mres = mf_median(keys, n);
    for (int j = 0; j < n; ++j)
        if (res[j] <= mres) { inlier_mask[vidx[j]] = 1; ++inl; }
    /* scale */
    const int ns = n < PAIR_PTS ? n : PAIR_PTS;
    int np = 0;
    for (int a = 0; a < ns; ++a)
        for (int b = a + 1; b < ns; ++b) {
            const int ia = vidx[n <= PAIR_PTS ? a : (int)(((long long)a * n) / PAIR_PTS)], ib = vidx[n <= PAIR_PTS ? b : (int)(((long long)b * n) / PAIR_PTS)];
            const double px = (double)pre[2 * ia] - (double)pre[2 * ib], py = (double)pre[2 * ia + 1] - (double)pre[2 * ib + 1];
            const double cx = (double)cur[2 * ia] - (double)cur[2 * ib], cy = (double)cur[2 * ia + 1] - (double)cur[2 * ib + 1];
            const double dp = sqrt(px * px + py * py), dc = sqrt(cx * cx + cy * cy);
            if (dp > 0.0) keys[np++] = mf_key(dc / dp);
        }
    const double s = np > 0 ? mf_median(keys, np) : 1.0;
    const double g = 0.5 * (s - 1.0);
    box_out[0] = (box_in[0] + mdx) - g * box_in[2];
    box_out[1] = (box_in[1] + mdy) - g * box_in[3];
    box_out[2] = box_in[2] * s;
    box_out[3] = box_in[3] * s;
    free(vidx); free(keys); free(res);
    return inl;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

PSNWO_API int psnwo_local_search_klt(const double *box_in, const float *pre, const float *cur,
                                     const uint8_t *valid, int n_all, double *box_out, int *inlier_idx)
{
    const double MIN_MOVEMENT = 0.1, WINDOW_RATIO = 0.2, OF_SCALE = 1.0;   /* :453-454, :18 */
    int n = 0, n_moving = 0;
    double *mvx = (double *)malloc(sizeof(double) * (size_t)(n_all + 1) * 4);
    double *mvy = mvx + n_all + 1, *dxs = mvy + n_all + 1, *dys = dxs + n_all + 1;
    int *midx = (int *)malloc(sizeof(int) * (size_t)(n_all + 1));
    for (int i = 0; i < n_all; i++) {
        if (valid && !valid[i]) continue;
        n++;
        /* cv::Point2f subtraction is float; PSN_Point2D promotes the result (:479, Types.h:29) */
        float fx = cur[2 * i] - pre[2 * i], fy = cur[2 * i + 1] - pre[2 * i + 1];
        double x = (double)fx, y = (double)fy;
        if (sqrt(x * x + y * y) < MIN_MOVEMENT * OF_SCALE) continue;       /* :480-484 */
        mvx[n_moving] = x; mvy[n_moving] = y; dxs[n_moving] = x; dys[n_moving] = y;
        midx[n_moving] = i; n_moving++;
    }
    memcpy(box_out, box_in, sizeof(double) * 4);
    int n_inl = 0;
    if ((double)n_moving < (double)n * 0.5) goto done;                      /* :493-496 */

    qsort(dxs, (size_t)n_moving, sizeof(double), cmp_double);               /* :498-499 */
    qsort(dys, (size_t)n_moving, sizeof(double), cmp_double);
    {
        const double win = box_in[2] * WINDOW_RATIO * OF_SCALE;             /* :502 */
        int max_nx = 0, max_ny = 0;
        double ex = 0.0, ey = 0.0;
        for (int k = 0; k < n_moving; k++) {                                /* :505-537 */
            int nx = 0, ny = 0;
            for (int j = 0; j < n_moving; j++) {
                if (fabs(dxs[k] - dxs[j]) < win) nx++;
                if (fabs(dys[k] - dys[j]) < win) ny++;
            }
            if (max_nx < nx) { ex = dxs[k]; max_nx = nx; }
            if (max_ny < ny) { ey = dys[k]; max_ny = ny; }
        }
        for (int k = 0; k < n_moving; k++) {                                /* :540-546 */
            double rx = mvx[k] - ex, ry = mvy[k] - ey;
            if (sqrt(rx * rx + ry * ry) < win) inlier_idx[n_inl++] = midx[k];
        }
        box_out[0] += ex; box_out[1] += ey;                                 /* :549-553 */
    }
done:
    free(mvx); free(midx);
    return n_inl;
}

/* ------------------------------------------------------------------------------------------------ */
/* A.4  Shi-Tomasi min-eigenvalue map (cornerMinEigenVal, blockSize 3, Sobel 3) and                  */
/*      goodFeaturesToTrack selection under a box mask (north_star's detector; call site :735-757).  */
/* ------------------------------------------------------------------------------------------------ */
PSNWO_API void psnwo_min_eig_map(const uint8_t *src, int w, int h, int sstride, float *eig, int estride_px)
{
    const float scale = (float)(1.0 / (255.0 * 4.0 * 3.0));
    float *dxx = (float *)malloc(sizeof(float) * (size_t)w * h * 3);
    float *dxy = dxx + (size_t)w * h, *dyy = dxy + (size_t)w * h;
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + (size_t)reflect101(y - 1, h) * sstride;
        const uint8_t *r1 = src + (size_t)y * sstride;
        const uint8_t *r2 = src + (size_t)reflect101(y + 1, h) * sstride;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int gx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
            int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
            float fx = (float)gx * scale, fy = (float)gy * scale;
            size_t p = (size_t)y * w + x;
            dxx[p] = fx * fx; dxy[p] = fx * fy; dyy[p] = fy * fy;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0, b = 0, c = 0;
            for (int j = -1; j <= 1; j++) {
                size_t row = (size_t)reflect101(y + j, h) * w;
                for (int i = -1; i <= 1; i++) {
                    size_t p = row + reflect101(x + i, w);
                    a += dxx[p]; b += dxy[p]; c += dyy[p];
                }
            }
            a *= 0.5f; c *= 0.5f;
            eig[(size_t)y * estride_px + x] = (a + c) - sqrtf((a - c) * (a - c) + b * b);
        }
    free(dxx);
}

/*
 * goodFeaturesToTrack(eig-map form) restricted to box (x0,y0,x1,y1) = cropped integer ROI, half-open.
 * Returns the number of corners written to out_xy (float x,y pairs, strongest first).
 * Tie order: equal scores are taken HIGHER raster address first -- cv2 4.13's greaterThanPtr
 * ((*a > *b) ? true : (*a < *b) ? false : (a > b)); pinned by
 * tests/test_oracle_vs_cv2.py::test_gftt_tie_order.  (OpenCV 2.4.6 left ties to std::sort.)
 */
typedef struct { float v; int idx; } cand_t;
static int cmp_cand(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return (x->idx < y->idx) - (x->idx > y->idx);
}

PSNWO_API int psnwo_gftt_select(const float *eig, int w, int h, int estride_px,
                                int x0, int y0, int x1, int y1,
                                int max_corners, float quality, float min_dist, float *out_xy)
{
    if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > w) x1 = w; if (y1 > h) y1 = h;
    if (x1 <= x0 || y1 <= y0 || max_corners <= 0) return 0;
    float maxv = 0.f;
    int any = 0;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            float v = eig[(size_t)y * estride_px + x];
            if (!any || v > maxv) { maxv = v; any = 1; }
        }
    const float thr = maxv * quality;
    cand_t *cands = (cand_t *)malloc(sizeof(cand_t) * (size_t)(x1 - x0) * (y1 - y0));
    int nc = 0;
    for (int y = (y0 > 1 ? y0 : 1); y < (y1 < h - 1 ? y1 : h - 1); y++)
        for (int x = (x0 > 1 ? x0 : 1); x < (x1 < w - 1 ? x1 : w - 1); x++) {
            float v = eig[(size_t)y * estride_px + x];
            if (!(v > thr)) continue;                 /* THRESH_TOZERO keeps v > thr */
            int is_max = 1;
            for (int j = -1; j <= 1 && is_max; j++)
                for (int i = -1; i <= 1; i++) {
                    float u = eig[(size_t)(y + j) * estride_px + (x + i)];
                    if (!(u > thr)) u = 0.f;          /* neighbours are thresholded before the dilate */
                    if (u > v) { is_max = 0; break; }
                }
            if (is_max) { cands[nc].v = v; cands[nc].idx = y * w + x; nc++; }
        }
    qsort(cands, (size_t)nc, sizeof(cand_t), cmp_cand);
    int n_out = 0;
    const float md2 = min_dist * min_dist;
    for (int k = 0; k < nc && n_out < max_corners; k++) {
        int x = cands[k].idx % w, y = cands[k].idx / w;
        int good = 1;
        if (min_dist >= 1.f)
            for (int j = 0; j < n_out; j++) {
                float dx = (float)x - out_xy[2 * j], dy = (float)y - out_xy[2 * j + 1];
                if (dx * dx + dy * dy < md2) { good = 0; break; }
            }
        if (good) { out_xy[2 * n_out] = (float)x; out_xy[2 * n_out + 1] = (float)y; n_out++; }
    }
    free(cands);
    return n_out;
}

/* ------------------------------------------------------------------------------------------------ */
/* A.6  GridFAST -- the reference's actual detector (next row f-1):                                  */
/*      cv::FeatureDetector::create("GridFAST"), psn_where/PSNWhere_Tracker2D.cpp:142, used at :741  */
/*      = GridAdaptedFeatureDetector(FastFeatureDetector(threshold 10, NMS), 1000 keypoints, 4 x 4). */
/*      OpenCV 2.4.6 is not vendored; FAST-9/16 is restated from features2d/fast.cpp + fast_score.cpp */
/*      and pinned against cv2 4.13's FastFeatureDetector run per grid cell (tests/test_gridfast.py). */
/* ------------------------------------------------------------------------------------------------ */
static const int fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* score of pixel (x, y): 0 when it is not a FAST-9/16 corner at `threshold`, else cornerScore<16> (>= threshold) */
static int fast_score_at(const uint8_t *img, int stride, int x, int y, int threshold)
{
    const int v = img[(size_t)y * stride + x];
    int d[25];
    for (int k = 0; k < 25; k++) d[k] = v - img[(size_t)(y + fast_dy[k & 15]) * stride + (x + fast_dx[k & 15])];
    /* corner test: 9 contiguous circle pixels all darker than v - t or all brighter than v + t */
    int is_corner = 0, run_d = 0, run_b = 0;
    for (int k = 0; k < 25 && !is_corner; k++) {
        run_d = d[k] > threshold ? run_d + 1 : 0;
        run_b = d[k] < -threshold ? run_b + 1 : 0;
        if (run_d >= 9 || run_b >= 9) is_corner = 1;
    }
    if (!is_corner) return 0;
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] < a) a = d[k + 3];
        if (a <= a0) continue;
        for (int q = 4; q <= 8; q++) if (d[k + q] < a) a = d[k + q];
        int m = a < d[k] ? a : d[k];
        if (m > a0) a0 = m;
        m = a < d[k + 9] ? a : d[k + 9];
        if (m > a0) a0 = m;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] > b) b = d[k + 3];
        if (b >= b0) continue;
        for (int q = 4; q <= 8; q++) if (d[k + q] > b) b = d[k + q];
        int m = b > d[k] ? b : d[k];
        if (m < b0) b0 = m;
        m = b > d[k + 9] ? b : d[k + 9];
        if (m < b0) b0 = m;
    }
    return -b0 - 1;
}

/*
 * Dense map of non-max-suppressed FAST scores with the grid-cell semantics of GridAdaptedFeatureDetector: every
 * cell is detected as its own sub-image, so pixels within 3 px of a cell edge are never corners and NMS never
 * looks across a cell edge.  out[y][x] = score (1..255) for a kept keypoint, 0 otherwise.
 */
PSNWO_API void psnwo_gridfast_map(const uint8_t *img, int w, int h, int stride, int threshold, int grid_rows, int grid_cols,
                                  uint8_t *out, int ostride)
{
    int *raw = (int *)calloc((size_t)w * h, sizeof(int));
    for (int i = 0; i < grid_rows; i++) {
        const int r0 = (i * h) / grid_rows, r1 = ((i + 1) * h) / grid_rows;
        for (int j = 0; j < grid_cols; j++) {
            const int c0 = (j * w) / grid_cols, c1 = ((j + 1) * w) / grid_cols;
            for (int y = r0 + 3; y < r1 - 3; y++)
                for (int x = c0 + 3; x < c1 - 3; x++) raw[(size_t)y * w + x] = fast_score_at(img, stride, x, y, threshold);
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int s = raw[(size_t)y * w + x];
            int keep = s > 0;
            for (int dy = -1; dy <= 1 && keep; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (!dx && !dy) continue;
                    const int yy = y + dy, xx = x + dx;
                    const int u = (yy < 0 || xx < 0 || yy >= h || xx >= w) ? 0 : raw[(size_t)yy * w + xx];
                    if (!(s > u)) { keep = 0; break; }
                }
            out[(size_t)y * ostride + x] = (uint8_t)(keep ? (s > 255 ? 255 : s) : 0);
        }
    free(raw);
}

/*
 * Keypoints of one detection box from the NMS'd score map: per grid cell keep the keypoints inside the ROI mask,
 * then keepStrongest(max_total / cells) INCLUDING every tie of the boundary response (KeyPointsFilter::retainBest).
 * If more than max_out survive over all cells, the strongest max_out are returned (ties cut in raster order) -- the
 * reference instead draws a random subset (std::random_shuffle, :752-757).  Output in raster order.
 */
PSNWO_API int psnwo_gridfast_select(const uint8_t *map, int w, int h, int mstride, int grid_rows, int grid_cols, int max_total,
                                    int x0, int y0, int x1, int y1, int max_out, float *out_xy)
{
    if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > w) x1 = w; if (y1 > h) y1 = h;
    if (x1 <= x0 || y1 <= y0 || max_out <= 0) return 0;
    const int per_cell = max_total / (grid_rows * grid_cols);
    uint8_t *keep = (uint8_t *)calloc((size_t)w * h, 1);
    int hist_all[256] = {0}, total = 0;
    for (int i = 0; i < grid_rows; i++) {
        const int r0 = (i * h) / grid_rows, r1 = ((i + 1) * h) / grid_rows;
        for (int j = 0; j < grid_cols; j++) {
            const int c0 = (j * w) / grid_cols, c1 = ((j + 1) * w) / grid_cols;
            const int ya = r0 > y0 ? r0 : y0, yb = r1 < y1 ? r1 : y1, xa = c0 > x0 ? c0 : x0, xb = c1 < x1 ? c1 : x1;
            if (ya >= yb || xa >= xb) continue;
            int hist[256] = {0}, n = 0;
            for (int y = ya; y < yb; y++)
                for (int x = xa; x < xb; x++) { const int s = map[(size_t)y * mstride + x]; if (s) { hist[s]++; n++; } }
            int thr = 1;
            if (n > per_cell) {                      /* boundary response = the per_cell-th strongest; keep >= it */
                int acc = 0;
                for (thr = 255; thr >= 1; thr--) { acc += hist[thr]; if (acc >= per_cell) break; }
                if (per_cell <= 0) thr = 256;
            }
            for (int y = ya; y < yb; y++)
                for (int x = xa; x < xb; x++) {
                    const int s = map[(size_t)y * mstride + x];
                    if (s && s >= thr) { keep[(size_t)y * w + x] = (uint8_t)s; hist_all[s]++; total++; }
                }
        }
    }
    int thr = 1, room_at_thr = 1 << 30;
    if (total > max_out) {
        int acc = 0;
        for (thr = 255; thr >= 1; thr--) { if (acc + hist_all[thr] >= max_out) break; acc += hist_all[thr]; }
        room_at_thr = max_out - acc;                 /* how many keypoints of response == thr still fit */
    }
    int n_out = 0;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const int s = keep[(size_t)y * w + x];
            if (!s || s < thr) continue;
            if (s == thr) { if (room_at_thr <= 0) continue; room_at_thr--; }
            out_xy[2 * n_out] = (float)x; out_xy[2 * n_out + 1] = (float)y; n_out++;
        }
    free(keep);
    return n_out;
}
