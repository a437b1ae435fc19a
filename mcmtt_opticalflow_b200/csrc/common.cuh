// Shared device/host declarations of the psnw KLT library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "psnw_klt.h"

namespace psnw {

// One pyramid level of the whole context: planes are [camera][ring slot], row pitch in u8 bytes ==
// row pitch of the derivative image in short2 elements.
struct LevelGeom {
    uint8_t *img;      // pixel (0,0) of plane 0; plane p, pixel (x,y) at img[p * iplane_px + y * ipitch + x]
    short2 *dxy;       // [planes][h][pitch] (dx,dy) int16, Scharr x32 (calcSharrDeriv layout)
    int w, h, pitch;   // pitch: derivative row pitch in short2 elements (also the pitch of per-camera level-0 maps)
    int plane_px;      // h * pitch
    int ipitch;        // image row pitch in bytes
    int iplane_px;     // image plane stride in bytes
    int apron;         // > 0: a REFLECT_101 apron of that many pixels surrounds every image plane in HBM (levels >= 1)
    int apron_l;       // physical left margin of an image row in bytes (0 without apron); pixel (0,0) stays 16-byte aligned
};

// Levels >= 1 are stored with a physical REFLECT_101 apron (written by the pyramid kernels), so the LK kernels stage or
// read coarse-level patches with aligned copies and no border arithmetic.  LevelGeom::apron pixels are mirrored on each
// side (psnw_params::apron; at least kApronMin = 26, what the 21x21 window with its 4-pixel staging margin needs; a context
// that tracks the reference's large per-target windows asks for more); apron_l = apron_left(a) bytes of left margin keep
// pixel (0,0) 16-byte aligned under the round-down of a staged row to 16 bytes, apron_right(a) covers its round-up.
// Level 0 is stored bare: frames are uploaded into it with one linear copy.
constexpr int kApronMin = 26, kApronMax = 160;
__host__ __device__ constexpr int apron_left(int a) { return (a + 6 + 15) & ~15; }
__host__ __device__ constexpr int apron_right(int a) { return a + 22; }

struct PyrDesc {
    LevelGeom lv[PSNW_MAX_LEVELS];
    int levels;
};

constexpr int kMaxPlanesPerLaunch = 255;
struct PlaneList {
    int n;
    int plane[kMaxPlanesPerLaunch];
};

// One (job, target) of a klt_track launch.
struct TrackGroup {
    int plane_from, plane_to;
    int win_w, win_h;
    int max_level;     // after buildOpticalFlowPyramid's truncation rule
    int pad;
};

// One job of a klt_track_batch launch (device-visible pointers).
struct TrackJob {
    const float2 *pts;
    const int *target_of_pt;   // may be null
    float2 *out;
    uint8_t *status;
    float *err;
    float *fb_err;             // may be null
    int pt_begin, n_pts;
    int group_base, n_targets;
};

// small batches carry their descriptors inside the kernel parameters (no copy ahead of the launch)
constexpr int kEmbedJobs = 8, kEmbedGroups = 16;

struct LkParams {
    LevelGeom lv[PSNW_MAX_LEVELS];
    const TrackJob *jobs;      // null: ejobs / egroups below
    const TrackGroup *groups;
    TrackJob ejobs[kEmbedJobs];
    TrackGroup egroups[kEmbedGroups];
    int n_jobs, n_pts;
    int max_iters;
    double eps2;
    float eps2_lo, eps2_hi;   // eps2 * (1 -/+ 1e-5): outside this band the float value decides the epsilon test
    float min_eig;
    float fb_thresh;
    int do_fb;
    const CUtensorMap *tmaps; // [3 * level + {0: I box 48x22, 1: J box 48x30, 2: (dx,dy) box 28x22}] of the 21x21 kernel (lk.cu)
    unsigned *work;           // [2 + kLkTailMax] item counter / finished-warp counter of the persistent 21x21 kernel, then the
                              // hand-over flags of the split tail (all zero between launches)
    int split_from;           // points >= split_from are tracked as two work items (forward, then backward); = n_pts: none
    unsigned long long *stats; // null, or [3] counters {solves, template passes, iterations} (psnw_lk_stats: the counting build of the 21x21 kernel)
};
constexpr int kLkTailMax = 16384;   // capacity of the split tail (points)

// per-device kernel attributes (opt-in dynamic shared memory); called by psnw_create on the context's device
int configure_lk_kernels();
int configure_box_kernels();
int configure_detect_kernels();

// --- launchers (each returns the number of kernels it enqueued, or -1 with cudaGetLastError set) ---
// with_apron: the launch also writes the REFLECT_101 apron of its input level (levels >= 1) with extra CTAs
int launch_pyr_level(const PyrDesc &pyr, int level, const CUtensorMap *map_in, const PlaneList &planes, int with_apron, int pdl_mode,
                     cudaStream_t stream);
// fused multi-level variant (pyramid_fused.cu): levels 0..nl-1 AND their aprons in one launch.  Three region heights
// (variant 0 / 1 / 2); fused_geometry = TMA box of the variant's level-0 map and the level-0 region one CTA owns
constexpr int kFusedVariants = 3;
void fused_geometry(int nl, int variant, int *box_w, int *box_h, int *own_w, int *own_h);
int configure_fused_kernels();
// late_wait: the launch touches no data of the kernels before it in the stream (pyramid builds of other planes): it runs
// beside them and only completes after them
int launch_pyr_fused(const PyrDesc &pyr, int base, int nl, int variant, int threads, const CUtensorMap *map0, const PlaneList &planes, int late_wait,
                     cudaStream_t stream);
// level walker (pyramid.cu): the levels >= base and every apron of the planes in one persistent launch (levels < base complete); ctl = walk_ctl_words() zeroed words
size_t walk_ctl_words(int levels, int n_planes);
int launch_pyr_walk(const PyrDesc &pyr, int base, const CUtensorMap *maps, const PlaneList &planes, unsigned *ctl, int ctas_per_sm, int sm_count,
                    cudaStream_t stream);
int launch_bgr2gray(const uint8_t *bgr, size_t bgr_pitch, uint8_t *gray, int gray_pitch, int w, int h,
                    int luma_bits, cudaStream_t stream);
int lk_w21_resident_ctas(int max_per_sm);   // CTAs of the persistent 21x21 kernel resident at once on the current device (capped per SM when > 0), or -1
int launch_lk(const LkParams &P, int all_21x21, int max_win_px, int resident_ctas, cudaStream_t stream);
int launch_box_update(const double *boxes_in, int n_targets, const float2 *pre, const float2 *cur,
                      const uint8_t *status, const int *offsets, double *boxes_out, uint8_t *mask,
                      int *counts, int max_pts_per_target, cudaStream_t stream, int mode = PSNW_BOX_REFERENCE);
int launch_chain_init(float2 *pts, const int *counts, int n_targets, int cap, uint8_t *live, int *tgt, int *offsets, int *alive,
                      int *steps_done, cudaStream_t stream);
int launch_chain_compact(const float2 *tracked, const uint8_t *mask, const int *n_inl, const double *box_step, int n_targets, int cap,
                         int min_inliers, int step, float2 *next_pts, uint8_t *next_live, int *next_count, double *boxes_steps,
                         int *alive, int *steps_done, cudaStream_t stream);
int launch_flow_majority(const float2 *pts, const int *offsets, int n_trk, const double *trk_boxes, const uint8_t *ok,
                         const double *det_boxes, int n_det, double ratio, int *counts, uint8_t *wipe, cudaStream_t stream);
struct EigList {
    int n;
    int plane[24];
    float *eig[24];
};
int launch_points_in_boxes(const float2 *pts, const int *offsets, int n_trk, const double *det_boxes, int n_det, int *counts,
                           cudaStream_t stream);
int launch_min_eig(const LevelGeom &l0, const EigList &list, cudaStream_t stream);

struct DetectTarget {
    int x0, y0, x1, y1;        // cropped ROI, half open
    int cand_begin, cand_cap;  // slice of the candidate scratch
    const float *eig;          // min-eigenvalue map of the target's camera (level-0 pitch)
    float2 *out;               // [max_per_target]
    int *count;
};
int launch_detect_select(int w, int h, int pitch, const DetectTarget *targets, int n_targets,
                         int max_per_target, float quality, float min_dist, unsigned long long *cand,
                         int max_roi_area, int max_tile_px, cudaStream_t stream);

struct GridFastTarget {
    int x0, y0, x1, y1;        // cropped ROI, half open
    const uint8_t *map;        // NMS'd FAST score map of the target's camera (level-0 pitch)
    float2 *out;               // [max_per_target]
    int *count;
};
int launch_fast_map(const LevelGeom &l0, int plane, int threshold, int grid_rows, int grid_cols, uint8_t *out, cudaStream_t stream);
int launch_gridfast_select(int w, int h, int pitch, const GridFastTarget *targets, int n_targets, int max_per_target,
                           int grid_rows, int grid_cols, int per_cell, cudaStream_t stream);

// tile geometry of the pyramid kernel (the TMA box must match on host and device)
constexpr int kTileW = 128, kTileH = 32;
// The innermost TMA start coordinate must be 16-byte aligned (u8 elements: a multiple of 16 pixels; a start at
// x0 - 4 faults with "illegal instruction" on sm_100a), so the box starts 16 px left of the tile: 16 + 128 + 16.
constexpr int kBoxW = 160, kBoxH = kTileH + 3;
constexpr int kHaloL = 16, kHaloT = 2;

}  // namespace psnw
