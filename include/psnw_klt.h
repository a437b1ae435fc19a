/*
 * psnw_klt.h -- C ABI of the B200-native KLT hot path of PSN_Where's 2D tracker.
 *
 * Drop-in boundary.  The reference has no FFI layer; the numeric callees of
 * CPSNWhere_Tracker2D::Run (psn_where/PSNWhere_Tracker2D.h:140, .cpp:251-373) are replaced one-for-one:
 *
 *   reference callee                                             call site (Tracker2D.cpp)   export
 *   cv::cvtColor + cv::resize into m_vecPtGrayFrameBuffer        :256-263, ring :70,:311-316  psnw_push_frame(s)
 *   cv::FeatureDetector::detect(img, keypoints, mask)            :741 (mask :735-742)         psnw_klt_detect
 *   cv::calcOpticalFlowPyrLK(prev,next,pts,out,status,err,win)   :776-782, :871-877           psnw_klt_track(_batch)
 *   LocalSearchKLT(box, pre, cur, inlierIdx)                     :452-554 (callers :787,:893) psnw_box_update
 *   Initialize / Finalize                                        :99-163 / :176-237           psnw_create / psnw_destroy
 *
 * Conventions: extern "C", POD arguments only, every call returns 0 or a negative PSNW_E* code, never
 * throws, never exits.  A context owns one GPU's shard of the camera rig: n_cameras cameras, each with
 * a ring of `ring_slots` frames kept on the device as Gaussian pyramid + Scharr derivatives.  Calls on
 * one context are serialised by the caller (as Tracker2D instances are); different
 * contexts may be driven from different host threads or processes (one per GPU).
 *
 * Memory flags select how caller pointers are interpreted:
 *   PSNW_MEM_HOST    pageable or pinned host memory; the call returns with outputs written (synchronous,
 *                    like the OpenCV calls it replaces).
 *   PSNW_MEM_PINNED  page-locked host memory; copies are enqueued on the context stream and the call
 *                    returns immediately -- psnw_sync() before reading outputs or reusing inputs.
 *   PSNW_MEM_DEVICE  device pointers; nothing is copied, everything is enqueued (track -> box_update ->
 *                    track chains stay on the GPU).
 * There is no CPU fallback: without a CUDA device psnw_create fails with PSNW_ECUDA.
 */
#ifndef PSNW_KLT_H_
#define PSNW_KLT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PSNW_API __attribute__((visibility("default")))
#else
#define PSNW_API
#endif

#define PSNW_VERSION 100

enum {
    PSNW_OK = 0,
    PSNW_EARG = -1,   /* bad argument */
    PSNW_ECUDA = -2,  /* CUDA runtime / driver failure (psnw_last_error has the text) */
    PSNW_ENOMEM = -3, /* allocation failed or a capacity given at create time was exceeded */
    PSNW_ESTATE = -4  /* call not valid in this state (e.g. slot not filled yet) */
};

enum { PSNW_MEM_HOST = 0, PSNW_MEM_PINNED = 1, PSNW_MEM_DEVICE = 2 };
/* OR-ed into mem_flags (PINNED / DEVICE only; ignored with HOST).
 *  psnw_klt_detect(_batch): the detection runs on an auxiliary stream forked from the context stream at the call, so the
 *    psnw_klt_track / psnw_box_update calls that FOLLOW overlap with it (they must not consume its outputs).  The outputs
 *    are complete, for work enqueued afterwards on the context stream, after psnw_join (device-side join, no host block) or the
 *    next plain push / detect call; the HOST may read them only after psnw_sync.
 *  psnw_push_frame(s)(_bgr): the upload and the pyramid build run on an upload stream that waits only for the launches
 *    already enqueued that read the ring slots being recycled -- not for the rest of the context stream -- so the NEXT
 *    frame is ingested while the current one is still being tracked (one frame in flight).  The ring rotates at the call:
 *    slot numbers in later calls refer to the new ring, and every later call that reads frames (track, detect, read_*)
 *    first makes the context stream wait for the push (device-side, no host block).  The caller keeps the source
 *    frames unchanged until psnw_sync (psnw_join only orders device work, it does not wait on the host). */
#define PSNW_MEM_OVERLAP 0x100
/* OR-ed into the mem_flags of psnw_push_frame(s): also compute the Shi-Tomasi min-eigenvalue map of the pushed frames
 * (a psnw_klt_detect on them will follow and then finds it ready).  Each camera keeps two maps, so the map of a frame
 * ingested ahead (PSNW_MEM_OVERLAP) never disturbs a detection still reading the map of the frame before. */
#define PSNW_PUSH_EIG 0x200

/* psnw_box_update estimators */
enum {
    PSNW_BOX_REFERENCE = 0, /* LocalSearchKLT semantics, Tracker2D.cpp:452-554 (parity-graded) */
    /* Median flow (Kalal et al.; the "median-displacement/scale reduction" of north_star).  NOT what
     * the reference runs, so it has no reference parity; its oracle is oracle/klt_oracle.c::psnwo_box_medianflow (bit-exact).
     *   V = the target's points with status != 0 (all when status is NULL), in index order, n = |V|;  n == 0: box unchanged.
     *   d_i = (double)cur_i - (double)pre_i.   median of m values = middle one (m odd) or mean of the two middle ones.
     *   (mdx, mdy) = per-axis medians of d_i;  res_i = |d_i - (mdx, mdy)|;  inlier <=> res_i <= median(res).
     *   scale s = median of |cur_a - cur_b| / |pre_a - pre_b| over the pairs a < b with |pre_a - pre_b| > 0 among S = V (n <= 128)
     *   or S = { V[floor(k n / 128)], k = 0..127 } (n > 128);  s = 1 without such a pair.
     *   box' = { x + mdx - (s-1) w / 2,  y + mdy - (s-1) h / 2,  w s,  h s }. */
    PSNW_BOX_MEDIANFLOW = 1
};

#define PSNW_MAX_LEVELS 8
#define PSNW_MAX_PTS_PER_TARGET 4096

typedef struct psnw_ctx psnw_ctx;

typedef struct psnw_params {
    int device;           /* CUDA ordinal */
    int n_cameras;        /* cameras owned by this context */
    int width, height;    /* frame size (camera model W x H, Tracker2D.cpp:118-119) */
    int levels;           /* pyramid levels built per frame = maxLevel + 1; OpenCV default 4 */
    int ring_slots;       /* frames kept per camera; 4 == PSN_2D_BACKTRACKING_INTERVAL (Tracker2D.cpp:16) */
    int win_w, win_h;     /* default LK window when no per-target window is given (21 x 21) */
    int max_iters;        /* 30  (OpenCV default TermCriteria) */
    double epsilon;       /* 0.01 (double, like cv::TermCriteria::epsilon) */
    float min_eig_thresh; /* 1e-4 */
    float fb_thresh;      /* forward-backward gate in px; < 0: single direction only (no backward pass) */
    int max_points;       /* capacity of one klt_track(_batch) call, all jobs together */
    int max_targets;      /* capacity of one klt_track(_batch) / box_update / detect call */
    void *stream;         /* cudaStream_t to enqueue on; NULL: the context creates its own */
    int apron;            /* mirrored pixels kept around the pyramid levels >= 1 in device memory (the REFLECT_101 border
                             calcOpticalFlowPyrLK builds around every level); 0 = 26, what the 21 x 21 window needs.  A
                             context that tracks the reference's per-target windows (up to ~55 x 130) sets about half the
                             largest window side (64): such windows then run without border arithmetic on the coarse
                             levels.  Results do not depend on it. */
} psnw_params;

/* Fills *p with the defaults: OpenCV's calcOpticalFlowPyrLK defaults as the reference uses them (4 levels, 21x21, 30
 * iterations, 0.01, 1e-4), 4 ring slots, and north_star's forward-backward gate fb_thresh = 1.0 px.  The reference itself
 * has NO per-point forward-backward gate: a caller that mirrors Tracker2D.cpp:776 / :871
 * literally sets fb_thresh < 0 (mcmtt_opticalflow_b200's Tracker2D mirror refuses a context that does not). */
PSNW_API int psnw_default_params(psnw_params *p);

PSNW_API int psnw_create(psnw_ctx **out, const psnw_params *p);   /* Tracker2D::Initialize, .cpp:99-163 */
/* The parameters the context runs with (after psnw_create's clamping of max_iters / epsilon / apron). */
PSNW_API int psnw_get_params(psnw_ctx *ctx, psnw_params *out);
PSNW_API int psnw_destroy(psnw_ctx *ctx);                         /* Tracker2D::Finalize,   .cpp:176-237 */

/*
 * Frame intake (replaces .cpp:256-263 + ring rotation :311-316).  `gray` is 8-bit single channel,
 * `pitch` bytes per row.  The frame goes into the camera's next ring slot; its pyramid (pyrDown) and
 * Scharr derivatives of every level are built once, on the device.  After the call slot 0 is this
 * frame, slot 1 the previous one, ... (slot k == *(m_vecPtGrayFrameBuffer.rbegin() + k)).
 */
PSNW_API int psnw_push_frame(psnw_ctx *ctx, int cam, const uint8_t *gray, size_t pitch,
                             uint32_t frame_idx, int mem_flags);
/* Same for several cameras at once (one batched launch per level; CPSNWhere::TrackPeople's camera loop,
 * psn_where/PSNWhere.cpp:257-266). cams[i] must be distinct. */
PSNW_API int psnw_push_frames(psnw_ctx *ctx, int n, const int *cams, const uint8_t *const *grays,
                              size_t pitch, uint32_t frame_idx, int mem_flags);
/* BGR (8UC3, cv::imread layout) intake: cvtColor(BGR2GRAY) of .cpp:257 fused into the upload.
 * luma_bits 14 = OpenCV 2.4.6 rounding, 15 = OpenCV 4.x. */
PSNW_API int psnw_push_frames_bgr(psnw_ctx *ctx, int n, const int *cams, const uint8_t *const *bgr,
                                  size_t pitch, uint32_t frame_idx, int luma_bits, int mem_flags);

/* Number of frames currently held for `cam` (0..ring_slots). */
PSNW_API int psnw_frames_held(psnw_ctx *ctx, int cam);

/*
 * Feature detection inside detection boxes (replaces m_detector->detect + mask, .cpp:735-757).
 * Shi-Tomasi min-eigenvalue corners (north_star) == cv::goodFeaturesToTrack(img, max_per_target, quality,
 * min_dist, mask = box cropped like PSN_Rect::cropWithSize, blockSize 3).  boxes: PSN_Rect layout
 * {x,y,w,h} doubles (PSNWhere_Types.h:112-122), ALWAYS host memory (detections come from the host-side
 * detector, psn_where/PSNWhere.cpp:261).  pts_xy: [n_targets][max_per_target][2] floats, strongest first;
 * counts[n_targets].  Slots >= counts[t] are set to (-1e6,-1e6), which psnw_klt_track rejects at once
 * (status 0), so the padded array can be tracked as is without a host round trip.
 */
PSNW_API int psnw_klt_detect(psnw_ctx *ctx, int cam, int slot, const double *boxes, int n_targets,
                             int max_per_target, float quality, float min_dist,
                             float *pts_xy, int *counts, int mem_flags);

typedef struct psnw_detect_job {
    int cam, slot;
    const double *boxes;     /* [n_targets][4], host memory */
    int n_targets;
    float *pts_xy;           /* [n_targets][max_per_target][2] */
    int *counts;             /* [n_targets] */
} psnw_detect_job;

/* Several psnw_klt_detect calls (different cameras) as one launch: CPSNWhere::TrackPeople's camera loop. */
PSNW_API int psnw_klt_detect_batch(psnw_ctx *ctx, const psnw_detect_job *jobs, int n_jobs, int max_per_target,
                                   float quality, float min_dist, int mem_flags);

/*
 * GridFAST detection ("next" row): the detector the reference actually creates at
 * Tracker2D.cpp:142 -- OpenCV 2.4 GridAdaptedFeatureDetector(FastFeatureDetector(threshold, NMS), max_total, grid_rows x
 * grid_cols), reference values (10, 1000, 4, 4) -- applied under the detection-box mask of :735-742.  Each cell keeps
 * its max_total/(rows*cols) strongest keypoints including ties; if more than max_per_target survive, the strongest are
 * kept (ties in raster order; the reference draws a random subset with std::random_shuffle, :752-757).  Output in
 * raster order, unused slots (-1e6,-1e6).  `boxes` is host memory.
 */
PSNW_API int psnw_gridfast_detect(psnw_ctx *ctx, int cam, int slot, const double *boxes, int n_targets,
                                  int max_per_target, int threshold, int grid_rows, int grid_cols, int max_total,
                                  float *pts_xy, int *counts, int mem_flags);
/* Copies the NMS'd FAST score map last computed for `cam` to host u8[h][w] (inspection). */
PSNW_API int psnw_read_fast_map(psnw_ctx *ctx, int cam, uint8_t *map_out);

/*
 * Sparse pyramidal LK from ring slot `slot_from` to `slot_to` of camera `cam` (replaces
 * cv::calcOpticalFlowPyrLK at .cpp:776-782 [slot_from = newer frame] and :871-877 [slot_from = 1, slot_to = 0]).
 *   pts_xy[n_pts][2]      seed points in slot_from
 *   target_of_pt[n_pts]   target index of each point, or NULL (all points use target 0 / the default window)
 *   win_wh[n_targets][2]  per-target window (reference: (int)box.w x (int)box.w backward, x (int)box.h forward),
 *                         or NULL for the context default
 *   out_xy, status, err   as calcOpticalFlowPyrLK's nextPts / status / err (err = 0 where status = 0)
 *   fb_err                if the context was created with fb_thresh >= 0 the kernel also runs the backward
 *                         pass out_xy -> p'' and sets status = st_f & st_b & (|p - p''| <= fb_thresh),
 *                         fb_err = |p - p''| (may be NULL to drop it).  A point the forward pass already lost keeps
 *                         status 0: the backward pass is not run for it and its fb_err is 0.  With fb_thresh < 0
 *                         fb_err is ignored.
 * Pyramid levels are truncated per window exactly like buildOpticalFlowPyramid (level kept while both
 * dimensions stay > window).  A target whose window cannot be tracked here (side < 1, not smaller than the frame, or more
 * than ~34 000 pixels) is skipped, not an error: its points come back unmoved with status 0, the other targets of the
 * batch are tracked as usual.
 */
PSNW_API int psnw_klt_track(psnw_ctx *ctx, int cam, int slot_from, int slot_to,
                            const float *pts_xy, const int *target_of_pt, int n_pts,
                            const int *win_wh, int n_targets,
                            float *out_xy, uint8_t *status, float *err, float *fb_err, int mem_flags);

typedef struct psnw_track_job {
    int cam, slot_from, slot_to;
    int n_pts;
    const float *pts_xy;
    const int *target_of_pt; /* may be NULL */
    int n_targets;           /* >= 1 */
    const int *win_wh;       /* [n_targets][2] or NULL */
    float *out_xy;
    uint8_t *status;
    float *err;
    float *fb_err;           /* may be NULL */
} psnw_track_job;

/* Several psnw_klt_track calls (any cameras / slot pairs) as ONE upload, ONE launch, ONE download. */
PSNW_API int psnw_klt_track_batch(psnw_ctx *ctx, const psnw_track_job *jobs, int n_jobs, int mem_flags);

/*
 * Per-target box motion (replaces LocalSearchKLT, .cpp:452-554; callers :787 and :893).
 *   boxes_in[n_targets][4]       PSN_Rect {x,y,w,h}
 *   pre_xy, cur_xy               all targets' points, CSR by target_offsets[n_targets + 1]
 *   status                       NULL: every point takes part (backward caller, :787); else only points with
 *                                status != 0 (forward caller's filter, :883-888, or the previous step's inliers)
 *   boxes_out[n_targets][4]      translated box (w,h unchanged)
 *   inlier_mask[n_points]        1 for the points LocalSearchKLT lists in inlierFeatureIndex
 *   inlier_counts[n_targets]
 */
PSNW_API int psnw_box_update(psnw_ctx *ctx, const double *boxes_in, int n_targets,
                             const float *pre_xy, const float *cur_xy, const uint8_t *status,
                             const int *target_offsets, int mode,
                             double *boxes_out, uint8_t *inlier_mask, int *inlier_counts, int mem_flags);

/*
 * psnw_klt_track immediately followed by psnw_box_update on its result -- the pair the reference always runs together
 * (.cpp:776-787 backward, :871-893 forward) -- as one call: one upload, two launches, one download, one wait.  The
 * points of target t are pts_xy[target_offsets[t] .. target_offsets[t+1]) (so target_of_pt is implied), the window of
 * target t is win_wh[t] (NULL: context default); use_status = 0 lets every point vote for the box (backward caller,
 * :787), 1 only those the tracker kept (forward caller, :883-888).  Outputs as in the two separate calls.
 */
PSNW_API int psnw_klt_track_update(psnw_ctx *ctx, int cam, int slot_from, int slot_to, const float *pts_xy,
                                   const int *target_offsets, int n_targets, const int *win_wh, const double *boxes_in,
                                   int use_status, float *out_xy, uint8_t *status, float *err, double *boxes_out,
                                   uint8_t *inlier_mask, int *inlier_counts, int mem_flags);

/*
 * "Next" row: counts[d][t] = number of tracker t's tracked points inside detection box d
 * (PSN_Rect::contain) -- the per-pair counts behind the reference's optical-flow majority check
 * (Tracker2D.cpp:939-944, :981-1022) -- computed next to the LK output so the points need not leave the device.
 * pts_xy: all trackers' tracked points, CSR by tracker_offsets[n_trackers + 1]; det_boxes[n_detections][4];
 * counts[n_detections][n_trackers] (the layout of matchingCostArray, :859).
 */
PSNW_API int psnw_points_in_boxes(psnw_ctx *ctx, const float *pts_xy, const int *tracker_offsets, int n_trackers,
                                  const double *det_boxes, int n_detections, int *counts, int mem_flags);

/*
 * The backward chain of a frame's detections (Tracker2D.cpp:763-811) as ONE call: no host round trip between its
 * dependent steps.  Step k = 1 .. n_steps (clamped to the frames held - 1, the reference stops at the first empty FIFO
 * slot, :768) tracks every live target's current points from ring slot k-1 into slot k (window win_wh[t]; the LK status is
 * ignored like at :787), runs LocalSearchKLT with boxes_in[t] -- always the detection box, :770 -- and
 *   - a target with fewer than min_inliers inliers stops for good (:788; reference value 4),
 *   - else its box of step k is recorded and its points become the inliers' tracked positions in index order (:802-807).
 * Points are padded per target: pts_xy[t][0 .. counts[t]) of [n_targets][cap][2] (the layout psnw_klt_detect writes).
 *   boxes_steps[n_steps][n_targets][4]   box of step k at [k-1] (zeros where the target had stopped)
 *   steps_done[n_targets]                last step the target completed (0 .. n_steps)
 *   pts_steps[n_steps][n_targets][cap][2], counts_steps[n_steps][n_targets]
 *                                        the target's points AFTER step k (= vecvecTrackedFeatures[k], :807)
 *   first_inlier_mask[n_targets][cap]    inliers of step 1 among the input points (the reference also keeps those, :792-800)
 * mem_flags: PSNW_MEM_HOST / PSNW_MEM_PINNED / PSNW_MEM_DEVICE (win_wh is always host memory).  n_targets * cap <= max_points.
 */
PSNW_API int psnw_klt_backward_chain(psnw_ctx *ctx, int cam, int n_steps, const float *pts_xy, const int *counts,
                                     int n_targets, int cap, const int *win_wh, const double *boxes_in, int min_inliers,
                                     double *boxes_steps, int *steps_done, float *pts_steps, int *counts_steps,
                                     uint8_t *first_inlier_mask, int mem_flags);

/*
 * "Next" row, the whole check: the optical-flow majority rule of Tracker2D.cpp:931-944 and :981-1022 on the
 * device.  counts[d][t] = tracked points of tracker t (pts_xy, CSR by tracker_offsets) inside detection box d, for the
 * trackers that reach the comparison (tracker_ok[t] != 0: at least 4 status-ok pairs, :889) and whose new box
 * tracker_boxes[t] overlaps the detection box (PSN_Rect::overlap, :937), else 0; wipe[d] = 1 where the reference resets the
 * detection's row of the cost matrix to infinity: its run-length rule over featuresInDetectionBox (first run whole, later
 * runs minus one, last run never compared unless alone) leaves no run above ratio x all points (reference ratio 0.5).
 */
PSNW_API int psnw_flow_majority(psnw_ctx *ctx, const float *pts_xy, const int *tracker_offsets, int n_trackers,
                                const double *tracker_boxes, const uint8_t *tracker_ok, const double *det_boxes,
                                int n_detections, double ratio, int *counts, uint8_t *wipe, int mem_flags);

PSNW_API int psnw_sync(psnw_ctx *ctx);
/* Device-side join of an overlapped detect / push into the context stream (see PSNW_MEM_OVERLAP): work enqueued on the
 * context stream AFTER this call is ordered after them.  It does NOT block the host: host-visible buffers (pinned
 * source frames, pinned or pageable outputs) may be reused or read only after psnw_sync -- or after the caller has
 * synchronised the context stream (psnw_get_stream) itself once psnw_join has returned. */
PSNW_API int psnw_join(psnw_ctx *ctx);
PSNW_API const char *psnw_last_error(psnw_ctx *ctx);
PSNW_API void *psnw_get_stream(psnw_ctx *ctx);

/* --- inspection (tests, bench) ---------------------------------------------------------------- */
/* Geometry of pyramid level `level`: width, height, u8 row pitch in bytes (dxy pitch = 4 * that). */
PSNW_API int psnw_level_info(psnw_ctx *ctx, int level, int *w, int *h, int *pitch);
/* Copies one level of (cam, slot) to host: img_out[h][w] u8 and/or dxy_out[h][w][2] int16 (NULL to skip). */
PSNW_API int psnw_read_level(psnw_ctx *ctx, int cam, int slot, int level, uint8_t *img_out, int16_t *dxy_out);
/* Levels >= 1 are kept in HBM with a REFLECT_101 apron (the border cv::calcOpticalFlowPyrLK builds around every pyramid
 * level with copyMakeBorder): *apron receives its width in pixels (0 for a level stored bare) and, when `out` is not NULL,
 * out[h + 2a][w + 2a] the image including the apron. */
PSNW_API int psnw_read_level_apron(psnw_ctx *ctx, int cam, int slot, int level, uint8_t *out, int *apron);
/* Copies the Shi-Tomasi min-eigenvalue map of (cam, slot) level 0 to host float[h][w] (computes it if needed). */
PSNW_API int psnw_read_min_eig(psnw_ctx *ctx, int cam, int slot, float *eig_out);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
PSNW_API uint64_t psnw_launch_count(psnw_ctx *ctx);
/* Algorithmic bytes one psnw_push_frame moves: N*(6*S_L - 1), SURVEY.md section 8d. */
PSNW_API uint64_t psnw_push_frame_bytes(psnw_ctx *ctx);
/* Re-runs the pyramid + Scharr build of the frames already in slot 0 of `cams` (no copy): lets bench.py time
 * the HBM-bound kernels alone with CUDA events on the context stream. */
PSNW_API int psnw_rebuild_pyramids(psnw_ctx *ctx, int n, const int *cams);
/* The same for the frames in ring slot `slot`: bench.py walks the ring so that a stream of launches touches more bytes than
 * the L2 holds. */
PSNW_API int psnw_rebuild_pyramids_slot(psnw_ctx *ctx, int n, const int *cams, int slot);
/* Work counters of the 21 x 21 LK kernel.  enable != 0: the launches that follow run the counting build of the kernel
 * (same results, three atomics per solve); enable == 0: back to the product kernel.  out (may be NULL) receives the
 * counters gathered since the previous call -- {LK solves, template passes (levels on which the 2x2 system was formed),
 * Gauss-Newton iterations} -- and they are reset.  bench.py derives the kernel's arithmetic roofline from them. */
PSNW_API int psnw_lk_stats(psnw_ctx *ctx, int enable, uint64_t *out);
/* Levels actually built (levels whose smaller side would drop under 8 px are skipped). */
PSNW_API int psnw_levels_built(psnw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PSNW_KLT_H_ */
