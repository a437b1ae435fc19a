// Host shim behind the C ABI of include/psnw_klt.h: context (= one GPU's shard of cameras), device ring of
// pyramids, staging, and the launch sequences.  No tracking arithmetic happens on the host; without a CUDA
// device every entry point fails with PSNW_ECUDA (there is no CPU fallback).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

using namespace psnw;

namespace {

struct DevBuf {
    char *p = nullptr;
    size_t cap = 0;
};

struct StageSlot {
    char *h = nullptr;       // pinned
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
    bool pending = false;
};

constexpr int kStageSlots = 4;
constexpr size_t kGenericSmemMax = 200 * 1024;   // template memory of lk_generic_kernel (6 B per pixel, rows padded to 4)

}  // namespace

struct psnw_ctx {
    psnw_params p;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t aux = nullptr;           // PSNW_MEM_OVERLAP detect calls run here, forked from / joined to `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool aux_pending = false;
    // PSNW_MEM_OVERLAP pushes (upload + pyramid of the NEXT frame while the current one is tracked) run on `up`.  They
    // are ordered after the already enqueued readers of the planes they recycle through reader events: every launch that
    // reads pyramid planes records an event from a small ring and stamps the planes with its sequence number.
    cudaStream_t up = nullptr;
    cudaEvent_t ev_up = nullptr;
    bool up_pending = false;
    static constexpr int kReaderRing = 32;
    cudaEvent_t rd_ev[kReaderRing] = {};
    uint64_t rd_seq = 0;
    std::vector<uint64_t> plane_rd_seq;
    PyrDesc pyr;
    uint8_t *img_alloc[PSNW_MAX_LEVELS] = {};   // allocation behind pyr.lv[l].img (which points at pixel (0,0) of plane 0)
    CUtensorMap maps[PSNW_MAX_LEVELS];
    CUtensorMap map_fused[kFusedVariants];   // level-0 maps with the boxes of the fused multi-level kernel's region variants
    CUtensorMap map_hyb[kFusedVariants];     // the same for level hyb_base: the per-level kernel builds levels < hyb_base, the fused one the rest
    int hyb_base = 0, hyb_levels = 0;        // hyb_levels = 3 or 4 when that split exists (pyramids of >= 4 levels)
    CUtensorMap *lk_maps = nullptr;       // device copy of the LK kernel's patch maps: 3 per level (LkParams::tmaps)
    int fused_levels = 0;                 // 3 or 4 when the fused kernel is usable, else 0
    size_t fused_max_px = 2200000;   // auto mode: fused kernel up to this many level-0 pixels per launch
    size_t fused_chained_max_px = 10000000;   // auto mode: a build that runs beside its predecessor takes the fused kernel up to this size
    int pyr_mode = 0;                     // 0 auto, 1 per-level kernels only, 2 fused from level 0, 3 hybrid, 4 level walker (PSNW_PYR_MODE, bring-up)
    static constexpr int kWalkBlocks = 8; // control blocks of the level walker, used in turn (each launch leaves its block zeroed)
    unsigned *walk_ctl = nullptr;
    size_t walk_ctl_stride = 0;
    unsigned walk_next = 0;
    int walk_base = 0;                    // levels below this one take the per-level kernel ahead of the walker (PSNW_WALK_BASE, bring-up)
    int walk_ctas = 0;                    // CTAs per SM of the walker (0: what fits; PSNW_WALK_CTAS, bring-up)
    size_t walk_min_px = 2200000;         // auto mode: launches above this many level-0 pixels take the walker
    // Consecutive fused builds of disjoint planes on one stream overlap (launch_pyr_fused late_wait): the count of this
    // context's launches after the last such build, its stream, and the planes built since the last build that waited first
    uint64_t pyr_chain_launches = ~(uint64_t)0;
    cudaStream_t pyr_chain_stream = nullptr;
    std::vector<uint8_t> pyr_chain_planes;
    bool pyr_overlap = true;              // PSNW_PYR_OVERLAP=0 (bring-up): every build waits for its predecessors first
    int fused_variant = -1;               // -1 auto, else the region variant of the fused kernel (PSNW_FUSED_VARIANT, bring-up)
    int fused_wide = -1;                  // -1 auto, else threads per CTA of the fused kernel: 256 / 512 / 1024 (PSNW_FUSED_THREADS, bring-up)
    int sm_count = 148;
    int planes = 0;
    std::vector<int> head, held;
    std::vector<uint64_t> plane_seq;     // push counter per plane (min-eig cache tag)
    uint64_t seq = 0;
    DevBuf arena;                         // per-call device scratch (stream ordered)
    size_t arena_off = 0;
    DevBuf desc;                          // job / group descriptors (track)
    unsigned *lk_work = nullptr;          // [2] work counters of the persistent 21x21 LK kernel (self-resetting)
    int lk_ctas = 0;                      // resident CTAs of that kernel on this device
    int lk_tail = 0;                      // points at the end of a large batch whose two passes are separate work items
    unsigned long long *lk_stats = nullptr;   // psnw_lk_stats: device counters of the counting build of the 21x21 kernel, or null
    bool lk_persistent = true;            // persistent warps + work counter; PSNW_LK_PERSISTENT=0 (bring-up): one point per warp
    DevBuf det_desc;                      // target descriptors + staged outputs (detect; may run on the aux stream)
    DevBuf cand;                          // detect candidate scratch
    DevBuf bgr;                           // BGR staging for psnw_push_frames_bgr
    // Shi-Tomasi maps (level-0 pitch), lazily allocated: TWO per camera ([2 cam], [2 cam + 1]) so that the map of the frame
    // just pushed can be computed (PSNW_PUSH_EIG) while a detection still reads the map of the frame before; eig_tag =
    // push counter of the frame a buffer holds; readers are tracked like ring planes (plane_rd_seq[planes + index])
    std::vector<float *> eig;
    std::vector<uint64_t> eig_tag;
    std::vector<uint8_t *> fmap;          // per camera NMS'd FAST score map (GridFAST), lazily allocated
    std::vector<uint64_t> fmap_tag;
    int fmap_cfg[3] = {-1, -1, -1};       // (threshold, grid_rows, grid_cols) the cached maps were built with
    StageSlot stage[kStageSlots];
    int stage_next = 0;
    uint64_t launches = 0;
    std::string err;
};

namespace {

int fail(psnw_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call)                                                            \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) return fail(ctx, PSNW_ECUDA, #call, e_);     \
    } while (0)

int ensure(psnw_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return PSNW_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    const size_t want = std::max(bytes, (size_t)1 << 16);
    if (cudaMalloc(&b.p, want) != cudaSuccess) return fail(ctx, PSNW_ENOMEM, "cudaMalloc scratch", cudaGetLastError());
    b.cap = want;
    return PSNW_OK;
}

template <typename T>
T *take(psnw_ctx *ctx, size_t n)
{
    const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T *r = reinterpret_cast<T *>(ctx->arena.p + ctx->arena_off);
    ctx->arena_off += bytes;
    return r;
}

// pinned staging slot for small descriptor uploads; waits for the previous upload from the same slot
int stage_acquire(psnw_ctx *ctx, size_t bytes, StageSlot **out)
{
    StageSlot &s = ctx->stage[ctx->stage_next];
    ctx->stage_next = (ctx->stage_next + 1) % kStageSlots;
    if (s.pending) { CK(cudaEventSynchronize(s.ev)); s.pending = false; }
    if (bytes > s.cap) {
        if (s.h) CK(cudaFreeHost(s.h));
        s.h = nullptr; s.cap = 0;
        const size_t want = std::max(bytes, (size_t)1 << 14);
        if (cudaMallocHost(&s.h, want) != cudaSuccess) return fail(ctx, PSNW_ENOMEM, "cudaMallocHost", cudaGetLastError());
        s.cap = want;
    }
    if (!s.ev) CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    *out = &s;
    return PSNW_OK;
}

int stage_upload(psnw_ctx *ctx, StageSlot *s, void *dst, size_t bytes, cudaStream_t stream = nullptr)
{
    if (!stream) stream = ctx->stream;
    CK(cudaMemcpyAsync(dst, s->h, bytes, cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(s->ev, stream));
    s->pending = true;
    return PSNW_OK;
}

// PSNW_MEM_HOST calls move their arguments as ONE pinned blob each way instead of one pageable copy per array: the stage
// slot mirrors the scratch arena byte for byte, the caller's inputs are written at the offsets of their device copies,
// [0, arena_off) goes up in one copy and comes back in one copy after the kernels, and the outputs are read from the slot.
constexpr int kBlobMaxPts = 100000;       // larger batches keep one copy per array (no multi-megabyte pinned slots)
template <typename T>
T *mirror(const psnw_ctx *ctx, const StageSlot *s, const T *dev_ptr)
{
    return reinterpret_cast<T *>(s->h + (reinterpret_cast<const char *>(dev_ptr) - reinterpret_cast<const char *>(ctx->arena.p)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_maps(psnw_ctx *ctx)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, PSNW_ECUDA, "cuTensorMapEncodeTiled not available");
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
    for (int l = 0; l < ctx->pyr.levels; ++l) {
        const LevelGeom &g = ctx->pyr.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)ctx->planes};
        const cuuint64_t strides[2] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.iplane_px};
        const cuuint32_t box[3] = {(cuuint32_t)kBoxW, (cuuint32_t)kBoxH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&ctx->maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g.img, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char msg[96];
            snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(level %d) -> CUresult %d", l, (int)r);
            return fail(ctx, PSNW_ECUDA, msg);
        }
    }
    ctx->fused_levels = ctx->pyr.levels >= 4 ? 4 : ctx->pyr.levels == 3 ? 3 : 0;
    ctx->hyb_levels = ctx->pyr.levels >= 4 ? std::min(4, ctx->pyr.levels - 1) : 0;
    ctx->hyb_base = ctx->hyb_levels ? ctx->pyr.levels - ctx->hyb_levels : 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int nl = pass ? ctx->hyb_levels : ctx->fused_levels, base = pass ? ctx->hyb_base : 0;
        for (int v = 0; nl && v < kFusedVariants; ++v) {
            int bw = 0, bh = 0, ow = 0, oh = 0;
            fused_geometry(nl, v, &bw, &bh, &ow, &oh);
            const LevelGeom &g = ctx->pyr.lv[base];
            const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)ctx->planes};
            const cuuint64_t strides[2] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.iplane_px};
            const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = encode(pass ? &ctx->map_hyb[v] : &ctx->map_fused[v], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g.img, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ctx, PSNW_ECUDA, "cuTensorMapEncodeTiled(fused kernel map)");
        }
    }
    // patch maps of the 21x21 LK kernel: u8 image with the I box (48 x 22) and the J box (48 x 30), (dx,dy) image with the
    // 28 x 22 box.  Apron levels describe the whole padded plane -- the apron is ordinary memory to the copy -- and the
    // kernel offsets its coordinates by (apron_l, apron); the derivative map zero-fills outside the image.
    std::vector<CUtensorMap> lk(3 * ctx->pyr.levels);
    for (int l = 0; l < ctx->pyr.levels; ++l) {
        const LevelGeom &g = ctx->pyr.lv[l];
        const cuuint32_t estr[3] = {1, 1, 1};
        const cuuint64_t idims[3] = {(cuuint64_t)(g.apron ? g.ipitch : g.w), (cuuint64_t)(g.h + 2 * g.apron), (cuuint64_t)ctx->planes};
        const cuuint64_t istr[2] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.iplane_px};
        void *ibase = g.apron ? (void *)ctx->img_alloc[l] : (void *)g.img;
        for (int k = 0; k < 2; ++k) {
            const cuuint32_t box[3] = {48, (cuuint32_t)(k ? 30 : 22), 1};
            if (encode(&lk[3 * l + k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ibase, idims, istr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return fail(ctx, PSNW_ECUDA, "cuTensorMapEncodeTiled(LK image patch map)");
        }
        const cuuint64_t ddims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.h, (cuuint64_t)ctx->planes};
        const cuuint64_t dstr[2] = {(cuuint64_t)g.pitch * 4, (cuuint64_t)g.plane_px * 4};
        const cuuint32_t dbox[3] = {28, 22, 1};
        if (encode(&lk[3 * l + 2], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, g.dxy, ddims, dstr, dbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(ctx, PSNW_ECUDA, "cuTensorMapEncodeTiled(LK derivative patch map)");
    }
    if (cudaMalloc(&ctx->lk_maps, sizeof(CUtensorMap) * lk.size()) != cudaSuccess) return fail(ctx, PSNW_ENOMEM, "cudaMalloc LK tensor maps");
    CK(cudaMemcpy(ctx->lk_maps, lk.data(), sizeof(CUtensorMap) * lk.size(), cudaMemcpyHostToDevice));
    return PSNW_OK;
}

inline int plane_of(const psnw_ctx *c, int cam, int slot)
{
    const int S = c->p.ring_slots;
    return cam * S + ((c->head[cam] - slot) % S + S) % S;
}

bool slot_ok(const psnw_ctx *c, int cam, int slot)
{
    return cam >= 0 && cam < c->p.n_cameras && slot >= 0 && slot < c->held[cam];
}

// every entry point runs on the context's device, whatever device the calling thread had current
int enter(psnw_ctx *ctx)
{
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == ctx->p.device) return PSNW_OK;
    CK(cudaSetDevice(ctx->p.device));
    return PSNW_OK;
}

// make the main stream wait for an overlapped push: every consumer of ring slots calls this first (no host block)
int join_up(psnw_ctx *ctx)
{
    if (ctx->up_pending) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_up, 0));
        ctx->up_pending = false;
    }
    return PSNW_OK;
}

// a launch on `st` reads the listed planes: overlapped pushes that recycle one of them wait for it
int note_readers(psnw_ctx *ctx, cudaStream_t st, const int *planes, int n)
{
    if (n <= 0) return PSNW_OK;
    const uint64_t seq = ++ctx->rd_seq;
    CK(cudaEventRecord(ctx->rd_ev[seq % psnw_ctx::kReaderRing], st));
    for (int i = 0; i < n; ++i) ctx->plane_rd_seq[planes[i]] = seq;
    return PSNW_OK;
}

// make the main stream wait for an overlapped detect and an overlapped push (no host block)
int join_aux(psnw_ctx *ctx)
{
    if (int rc = join_up(ctx)) return rc;
    if (ctx->aux_pending) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        ctx->aux_pending = false;
    }
    return PSNW_OK;
}

// buildOpticalFlowPyramid's truncation: a level is kept while both dimensions stay > window
int effective_max_level(const psnw_ctx *c, int win_w, int win_h)
{
    int level = 0;
    while (level + 1 < c->pyr.levels) {
        const LevelGeom &n = c->pyr.lv[level + 1];
        if (n.w <= win_w || n.h <= win_h) break;
        ++level;
    }
    return level;
}

inline bool do_fb_early(const psnw_ctx *c) { return c->p.fb_thresh >= 0.f; }

cudaMemcpyKind h2d(int mem_flags) { return mem_flags == PSNW_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }

// region variant of the fused kernel for a launch over n planes of level `base`: the tallest region that still gives every
// SM its two CTAs for a few rounds -- the halo rows every CTA recomputes weigh less the taller its region -- but never one
// that leaves SMs without work.  *threads: 512 per CTA (two CTAs per SM) from level 0, 256 (four per SM) for the smaller
// regions of the coarse-level launch of the hybrid builder.
int fused_variant_for(const psnw_ctx *ctx, int base, int nl, int n, int *threads, bool chained = false)
{
    const LevelGeom &g = ctx->pyr.lv[base];
    int variant = 0;
    for (int v = kFusedVariants - 1; v > 0 && variant == 0; --v) {
        int bw, bh, ow, oh;
        fused_geometry(nl, v, &bw, &bh, &ow, &oh);
        const size_t ctas = (size_t)((g.w + ow - 1) / ow) * ((g.h + oh - 1) / oh) * n;
        if (ctas >= (size_t)ctx->sm_count * 6) variant = v;
        // a build that runs beside its predecessor (late_wait) is one of a stream of builds: throughput counts, not the latency
        // of this launch -- half as many CTAs of twice the height recompute fewer halo rows (4 x 768x576 in-stream: 4.8 us
        // against 5.5; alone 16.8 us against 14.7)
        if (chained && v == 1 && ctas >= (size_t)ctx->sm_count / 2) variant = v;
    }
    if (ctx->fused_variant >= 0) variant = ctx->fused_variant;
    *threads = (base > 0 || (chained && variant >= 1)) ? 256 : 512;
    if (ctx->fused_wide >= 0) *threads = ctx->fused_wide;
    return variant;
}

int build_pyramids(psnw_ctx *ctx, const std::vector<int> &planes, cudaStream_t st)
{
    for (size_t b = 0; b < planes.size(); b += kMaxPlanesPerLaunch) {
        PlaneList pl;
        pl.n = (int)std::min<size_t>(kMaxPlanesPerLaunch, planes.size() - b);
        for (int i = 0; i < pl.n; ++i) pl.plane[i] = planes[b + i];
        // Three builders (PSNW_PYR_MODE forces one):
        //   fused   levels 0..3 AND their aprons in ONE launch of independent CTAs: a rig frame of a few small cameras, where
        //           one launch per level would be pure latency;
        //   hybrid  the per-level kernel for the finest level(s) (no halo recomputed where the pixels are), then ONE fused
        //           launch for the remaining 3-4 levels and every apron (PSNW_PYR_MODE=hybrid only);
        //   level   one launch per level (each writes the apron of its input level with extra CTAs): everything larger.
        const LevelGeom &l0 = ctx->pyr.lv[0];
        const size_t px = (size_t)l0.w * l0.h * pl.n;
        int mode = ctx->pyr_mode;                                    // 1 level, 2 fused, 3 hybrid
        // (measured, tools/pyr_sweep.py: 4 x 768x576 fused 14.0 us / level 15.9; 32 x 1080p level 93 us / hybrid 99; 16 x 4K
        // level 163 us / hybrid 173)
        // round 2, final kernels (tools/pyr_sweep.py, isolated): 16 / 32 / 64 / 128 x 768x576 level 24.1 / 34.7 / 57.2 / 98.2 us against
        // hybrid 25.5 / 37.5 / 60.3 / 107.4; 8 x 1080p 34.7 against 39.7 -- the hybrid builder is no longer chosen by itself
        if (mode == 0) mode = (ctx->fused_levels && px <= ctx->fused_max_px) ? 2 : 1;
        // one of a stream of builds (it may run beside its predecessor, below): the single fused launch overlaps better than the
        // hybrid's dependent pair (16 x 768x576 in-stream 17.7 us against 19.2, 64 x 57.9 against 58.7; alone it is the slower one)
        // (in-stream 16 x 768x576: fused 17.3 us, level 19.8; 64 x: level 54.2, fused 56.0)
        if (ctx->pyr_mode == 0 && mode == 1 && px <= ctx->fused_chained_max_px && ctx->fused_levels == ctx->pyr.levels && ctx->pyr_overlap &&
            ctx->launches == ctx->pyr_chain_launches && st == ctx->pyr_chain_stream && ctx->pyr_chain_planes.size() == (size_t)ctx->planes) {
            bool disjoint = true;
            for (int i = 0; disjoint && i < pl.n; ++i) disjoint = !ctx->pyr_chain_planes[pl.plane[i]];
            if (disjoint) mode = 2;
        }
        if (mode == 4) {
            // the finest `walk_base` levels by the per-level kernel (no apron CTAs: the walker scatters the apron of every
            // level it reads), the rest by the walker
            const int wb = std::min(ctx->walk_base, ctx->pyr.levels - 1);
            for (int l = 0; l < wb; ++l) {
                const int r = launch_pyr_level(ctx->pyr, l, &ctx->maps[l], pl, l > 0 ? 1 : 0, 0, st);
                if (r < 0) return fail(ctx, PSNW_ECUDA, "pyr_scharr_kernel launch", cudaGetLastError());
                ctx->launches += r;
            }
            unsigned *ctl = ctx->walk_ctl + (ctx->walk_next++ % psnw_ctx::kWalkBlocks) * ctx->walk_ctl_stride;
            const int r = launch_pyr_walk(ctx->pyr, wb, ctx->maps, pl, ctl, ctx->walk_ctas, ctx->sm_count, st);
            if (r < 0) return fail(ctx, PSNW_ECUDA, "pyr_walk_kernel launch", cudaGetLastError());
            ctx->launches += r;
            continue;
        }
        if (mode == 2 && !ctx->fused_levels) mode = 1;
        if (mode == 3 && !ctx->hyb_levels) mode = ctx->fused_levels ? 2 : 1;
        int first = 0, last = ctx->pyr.levels;                       // levels [first, last) take the per-level kernel
        // A build reads only the frames (complete: a copy or an earlier call put them there) and writes only the levels of its
        // own planes.  Straight after a build of OTHER planes on the same stream whose last launch was the fused kernel (which
        // releases its dependents only after its own wait), the first launch of this build need not wait for it (late_wait):
        // consecutive builds overlap.  pyr_chain_planes = planes built since the last build that waited first.
        bool late = (mode >= 1 && mode <= 3) && ctx->pyr_overlap && ctx->launches == ctx->pyr_chain_launches && st == ctx->pyr_chain_stream;
        if (ctx->pyr_chain_planes.size() != (size_t)ctx->planes) ctx->pyr_chain_planes.assign(ctx->planes, 0);
        for (int i = 0; late && i < pl.n; ++i) late = !ctx->pyr_chain_planes[pl.plane[i]];
        if (!late) std::fill(ctx->pyr_chain_planes.begin(), ctx->pyr_chain_planes.end(), 0);
        for (int i = 0; i < pl.n; ++i) ctx->pyr_chain_planes[pl.plane[i]] = 1;
        bool ends_fused = false;
        if (mode == 2) {
            int wide = 0;
            const int variant = fused_variant_for(ctx, 0, ctx->fused_levels, pl.n, &wide, late);
            // deeper pyramids: level 3 again (per-level kernel) only to feed level 4; it rewrites its apron with the same values
            first = ctx->fused_levels < ctx->pyr.levels ? ctx->fused_levels - 1 : ctx->pyr.levels;
            const int r = launch_pyr_fused(ctx->pyr, 0, ctx->fused_levels, variant, wide, &ctx->map_fused[variant], pl, late ? 1 : 0, st);
            if (r < 0) return fail(ctx, PSNW_ECUDA, "pyr_fused_kernel launch", cudaGetLastError());
            ctx->launches += r;
            ends_fused = first == ctx->pyr.levels;
            late = false;
        } else if (mode == 3) {
            last = ctx->hyb_base;
        }
        for (int l = first; l < last; ++l) {
            // (the last launch of a per-level build waits before it releases its dependents, like the fused kernel)
            const bool closes = mode == 1 && l + 1 == last && l > 0;
            const int r = launch_pyr_level(ctx->pyr, l, &ctx->maps[l], pl, 1, (late && l == 0) ? 1 : closes ? 2 : 0, st);
            if (r < 0) return fail(ctx, PSNW_ECUDA, "pyr_scharr_kernel launch", cudaGetLastError());
            ctx->launches += r;
            if (closes) ends_fused = true;
        }
        if (mode == 3) {
            int wide = 0;
            const int variant = fused_variant_for(ctx, ctx->hyb_base, ctx->hyb_levels, pl.n, &wide);
            const int r = launch_pyr_fused(ctx->pyr, ctx->hyb_base, ctx->hyb_levels, variant, wide, &ctx->map_hyb[variant], pl, 0, st);
            if (r < 0) return fail(ctx, PSNW_ECUDA, "pyr_fused_kernel launch", cudaGetLastError());
            ctx->launches += r;
            ends_fused = true;
        }
        ctx->pyr_chain_launches = ends_fused ? ctx->launches : ~(uint64_t)0;
        ctx->pyr_chain_stream = st;
    }
    return PSNW_OK;
}

int advance_ring(psnw_ctx *ctx, int n, const int *cams, std::vector<int> &planes, bool overlap = false)
{
    for (int i = 0; i < n; ++i) {
        const int cam = cams[i];
        if (cam < 0 || cam >= ctx->p.n_cameras) return fail(ctx, PSNW_EARG, "camera index out of range");
        for (int j = 0; j < i; ++j)
            if (cams[j] == cam) return fail(ctx, PSNW_EARG, "duplicate camera in psnw_push_frames");
    }
    // a push may recycle the slot an overlapped detect still reads; an overlapped push orders itself after the readers
    // of exactly the planes it recycles instead (push_stream)
    if (!overlap)
        if (int rc = join_aux(ctx)) return rc;
    planes.resize(n);
    for (int i = 0; i < n; ++i) {
        const int cam = cams[i];
        ctx->head[cam] = (ctx->head[cam] + 1) % ctx->p.ring_slots;
        ctx->held[cam] = std::min(ctx->held[cam] + 1, ctx->p.ring_slots);
        planes[i] = plane_of(ctx, cam, 0);
        ctx->plane_seq[planes[i]] = ++ctx->seq;
    }
    return PSNW_OK;
}

// A push that fails after advance_ring must not leave slot 0 pointing at a recycled plane with stale pixels: the
// transaction restores head / held / plane_seq / seq of the cameras it rotated unless commit() is reached.
struct RingTxn {
    psnw_ctx *ctx;
    std::vector<int> cams, head, held, planes;
    std::vector<uint64_t> pseq;
    uint64_t seq;
    bool done = false;
    RingTxn(psnw_ctx *c, int n, const int *cs) : ctx(c), seq(c->seq)
    {
        for (int i = 0; i < n; ++i) {
            const int cam = cs[i];
            if (cam < 0 || cam >= c->p.n_cameras) continue;       // advance_ring rejects the call before touching anything
            cams.push_back(cam); head.push_back(c->head[cam]); held.push_back(c->held[cam]);
            const int S = c->p.ring_slots, next = cam * S + (c->head[cam] + 1) % S;
            planes.push_back(next); pseq.push_back(c->plane_seq[next]);
        }
    }
    void commit() { done = true; }
    ~RingTxn()
    {
        if (done) return;
        for (size_t i = 0; i < cams.size(); ++i) {
            ctx->head[cams[i]] = head[i]; ctx->held[cams[i]] = held[i]; ctx->plane_seq[planes[i]] = pseq[i];
        }
        ctx->seq = seq;
    }
};

// make `st` wait for the launches already enqueued that read the listed resources (ring planes / min-eig buffers)
int wait_readers(psnw_ctx *ctx, cudaStream_t st, const std::vector<int> &ids)
{
    uint64_t newest = 0;
    for (int id : ids) newest = std::max(newest, ctx->plane_rd_seq[id]);
    if (newest == 0) return PSNW_OK;
    if (newest + psnw_ctx::kReaderRing > ctx->rd_seq) {
        CK(cudaStreamWaitEvent(st, ctx->rd_ev[newest % psnw_ctx::kReaderRing], 0));
    } else {                                            // its event was reused since: wait for every recent reader
        for (int i = 0; i < psnw_ctx::kReaderRing; ++i) CK(cudaStreamWaitEvent(st, ctx->rd_ev[i], 0));
    }
    return PSNW_OK;
}

// Shi-Tomasi map of every (camera, slot) of `cs`, computed on `stream` unless a buffer already holds it; bufs[i] receives
// the index of the buffer (into ctx->eig) with the map of cs[i].  A map is written over the older of the camera's two
// buffers.  `side_stream`: the caller is not ordered after earlier detections (the eager path of an overlapped push), so
// the stream first waits for the readers of the buffers it overwrites.
int ensure_min_eig(psnw_ctx *ctx, const std::vector<std::pair<int, int>> &cs, cudaStream_t stream,
                          std::vector<int> *bufs = nullptr, bool side_stream = false)
{
    const LevelGeom &l0 = ctx->pyr.lv[0];
    EigList el;
    el.n = 0;
    std::vector<int> todo_plane, todo_buf;
    if (bufs) bufs->assign(cs.size(), -1);
    for (size_t i = 0; i < cs.size(); ++i) {
        const int cam = cs[i].first, plane = plane_of(ctx, cam, cs[i].second);
        for (int b = 0; b < 2; ++b)
            if (!ctx->eig[2 * cam + b]) {
                if (cudaMalloc(&ctx->eig[2 * cam + b], sizeof(float) * (size_t)l0.plane_px) != cudaSuccess)
                    return fail(ctx, PSNW_ENOMEM, "cudaMalloc min-eig map", cudaGetLastError());
                ctx->eig_tag[2 * cam + b] = 0;
            }
        const uint64_t want = ctx->plane_seq[plane];
        int b = ctx->eig_tag[2 * cam] == want ? 2 * cam : ctx->eig_tag[2 * cam + 1] == want ? 2 * cam + 1 : -1;
        if (b < 0) {
            b = ctx->eig_tag[2 * cam] <= ctx->eig_tag[2 * cam + 1] ? 2 * cam : 2 * cam + 1;      // the older map goes
            ctx->eig_tag[b] = want;
            todo_plane.push_back(plane); todo_buf.push_back(b);
        }
        if (bufs) (*bufs)[i] = b;
    }
    if (todo_buf.empty()) return PSNW_OK;
    if (side_stream) {
        std::vector<int> ids;
        for (int b : todo_buf) ids.push_back(ctx->planes + b);
        if (int rc = wait_readers(ctx, stream, ids)) return rc;
    }
    for (size_t i = 0; i < todo_buf.size(); ++i) {
        el.plane[el.n] = todo_plane[i];
        el.eig[el.n] = ctx->eig[todo_buf[i]];
        if (++el.n == 24 || i + 1 == todo_buf.size()) {
            const int r = launch_min_eig(l0, el, stream);
            if (r < 0) return fail(ctx, PSNW_ECUDA, "min_eig_kernel launch", cudaGetLastError());
            ctx->launches += r;
            el.n = 0;
        }
    }
    return PSNW_OK;
}

// Stream a push runs on.  PSNW_MEM_OVERLAP (pinned / device memory only): the upload stream, made to wait for the launches
// already enqueued that read the recycled planes -- and for nothing else, so the push overlaps the LK launch in flight.
int push_stream(psnw_ctx *ctx, bool overlap, const std::vector<int> &planes, cudaStream_t *st)
{
    *st = ctx->stream;
    if (!overlap) return PSNW_OK;
    *st = ctx->up;
    return wait_readers(ctx, ctx->up, planes);
}

int push_done(psnw_ctx *ctx, bool overlap)
{
    if (overlap) {
        CK(cudaEventRecord(ctx->ev_up, ctx->up));
        ctx->up_pending = true;
    }
    return PSNW_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int psnw_default_params(psnw_params *p)
{
    if (!p) return PSNW_EARG;
    memset(p, 0, sizeof *p);
    p->device = 0;
    p->n_cameras = 1;
    p->width = 768; p->height = 576;
    p->levels = 4;
    p->ring_slots = 4;
    p->win_w = 21; p->win_h = 21;
    p->max_iters = 30;
    p->epsilon = 0.01;
    p->min_eig_thresh = 1e-4f;
    p->fb_thresh = 1.0f;
    p->max_points = 1 << 16;
    p->max_targets = 1 << 12;
    p->stream = nullptr;
    p->apron = 0;
    return PSNW_OK;
}

int psnw_create(psnw_ctx **out, const psnw_params *p)
{
    if (!out || !p) return PSNW_EARG;
    *out = nullptr;
    if (p->n_cameras < 1 || p->width < 16 || p->height < 16 || p->width > 32768 || p->height > 32768 || p->levels < 1 || p->levels > PSNW_MAX_LEVELS ||
        p->ring_slots < 1 || p->win_w < 1 || p->win_h < 1 || p->max_iters < 0 || p->max_points < 1 || p->max_targets < 1)
        return PSNW_EARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev) return PSNW_ECUDA;
    if (cudaSetDevice(p->device) != cudaSuccess) return PSNW_ECUDA;

    psnw_ctx *ctx = new psnw_ctx();
    ctx->p = *p;
    ctx->p.max_iters = std::min(std::max(p->max_iters, 0), 100);                 // calcOpticalFlowPyrLK clamps
    ctx->p.epsilon = std::min(std::max(p->epsilon, 0.0), 10.0);
    ctx->p.apron = std::min(std::max(p->apron, kApronMin), kApronMax);           // 0 = default
    ctx->planes = p->n_cameras * p->ring_slots;
    ctx->head.assign(p->n_cameras, p->ring_slots - 1);
    ctx->held.assign(p->n_cameras, 0);
    ctx->plane_seq.assign(ctx->planes, 0);
    ctx->plane_rd_seq.assign(ctx->planes + 2 * p->n_cameras, 0);   // ring planes, then the min-eig buffers
    ctx->eig.assign(2 * p->n_cameras, nullptr);
    ctx->eig_tag.assign(2 * p->n_cameras, 0);
    ctx->fmap.assign(p->n_cameras, nullptr);
    ctx->fmap_tag.assign(p->n_cameras, 0);
    memset(&ctx->pyr, 0, sizeof ctx->pyr);

    auto bail = [&](int code) { psnw_destroy(ctx); return code; };

    if (p->stream) {
        ctx->stream = (cudaStream_t)p->stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PSNW_ECUDA);
        ctx->own_stream = true;
    }

    if (configure_lk_kernels() || configure_box_kernels() || configure_detect_kernels() || configure_fused_kernels()) return bail(PSNW_ECUDA);
    ctx->lk_ctas = lk_w21_resident_ctas(0);
    if (ctx->lk_ctas <= 0) return bail(PSNW_ECUDA);
    if (cudaMalloc(&ctx->lk_work, (2 + kLkTailMax) * sizeof(unsigned)) != cudaSuccess) return bail(PSNW_ENOMEM);
    ctx->walk_ctl_stride = (walk_ctl_words(PSNW_MAX_LEVELS, kMaxPlanesPerLaunch) + 63) & ~(size_t)63;
    if (cudaMalloc(&ctx->walk_ctl, psnw_ctx::kWalkBlocks * ctx->walk_ctl_stride * sizeof(unsigned)) != cudaSuccess) return bail(PSNW_ENOMEM);
    cudaMemsetAsync(ctx->walk_ctl, 0, psnw_ctx::kWalkBlocks * ctx->walk_ctl_stride * sizeof(unsigned), ctx->stream);
    cudaMemsetAsync(ctx->lk_work, 0, (2 + kLkTailMax) * sizeof(unsigned), ctx->stream);
    ctx->lk_tail = 2 * ctx->lk_ctas * 4;  // two waves of warps
    if (const char *m = getenv("PSNW_LK_TAIL")) ctx->lk_tail = atoi(m);
    ctx->lk_tail = std::min(std::max(ctx->lk_tail, 0), kLkTailMax);
    if (const char *m = getenv("PSNW_LK_PERSISTENT")) ctx->lk_persistent = m[0] == '1';
    if (const char *m = getenv("PSNW_PYR_MODE")) ctx->pyr_mode = !strcmp(m, "level") ? 1 : !strcmp(m, "fused") ? 2 : !strcmp(m, "hybrid") ? 3 : !strcmp(m, "walk") ? 4 : 0;
    if (const char *m = getenv("PSNW_PYR_OVERLAP")) ctx->pyr_overlap = atoi(m) != 0;
    if (const char *m = getenv("PSNW_WALK_CTAS")) ctx->walk_ctas = atoi(m);
    if (const char *m = getenv("PSNW_WALK_BASE")) ctx->walk_base = std::max(atoi(m), 0);
    if (const char *m = getenv("PSNW_WALK_MIN_PX")) ctx->walk_min_px = (size_t)atoll(m);
    if (const char *m = getenv("PSNW_FUSED_THREADS")) ctx->fused_wide = atoi(m);
    if (const char *m = getenv("PSNW_FUSED_VARIANT")) ctx->fused_variant = std::min(std::max(atoi(m), 0), kFusedVariants - 1);
    if (const char *m = getenv("PSNW_FUSED_MAX_PX")) ctx->fused_max_px = (size_t)atoll(m);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, p->device);
    // the side streams outrank the context stream: their short kernels (pyramid, detection) take the SM slots that LK
    // warps free one by one instead of queueing behind the whole LK grid
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->up, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_up, cudaEventDisableTiming) != cudaSuccess)
        return bail(PSNW_ECUDA);
    for (cudaEvent_t &e : ctx->rd_ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(PSNW_ECUDA);

    // level geometry; levels smaller than 8 px are not built (no window fits them anyway)
    int w = p->width, h = p->height, levels = 0;
    for (int l = 0; l < p->levels; ++l) {
        if (l > 0) { w = (w + 1) / 2; h = (h + 1) / 2; }
        if (w < 8 || h < 8) break;
        LevelGeom &g = ctx->pyr.lv[l];
        g.w = w; g.h = h;
        g.pitch = (w + 127) & ~127;
        g.plane_px = g.pitch * h;
        // level 0 is stored bare (one linear upload per frame); levels >= 1 with a REFLECT_101 apron (common.cuh)
        g.apron = l > 0 ? ctx->p.apron : 0;
        g.apron_l = g.apron ? apron_left(g.apron) : 0;
        g.ipitch = g.apron ? (g.apron_l + w + apron_right(g.apron) + 63) & ~63 : g.pitch;
        g.iplane_px = g.apron ? g.ipitch * (h + 2 * g.apron) : g.plane_px;
        const size_t front = g.apron ? (size_t)g.apron * g.ipitch + g.apron_l : 0;
        const size_t px = (size_t)g.plane_px * ctx->planes, ipx = (size_t)g.iplane_px * ctx->planes;
        if (cudaMalloc(&ctx->img_alloc[l], ipx) != cudaSuccess || cudaMalloc(&g.dxy, px * sizeof(short2)) != cudaSuccess) {
            *out = nullptr;
            return bail(PSNW_ENOMEM);
        }
        g.img = ctx->img_alloc[l] + front;
        cudaMemsetAsync(ctx->img_alloc[l], 0, ipx, ctx->stream);
        cudaMemsetAsync(g.dxy, 0, px * sizeof(short2), ctx->stream);
        ++levels;
    }
    ctx->pyr.levels = levels;
    if (make_tensor_maps(ctx) != PSNW_OK) {
        fprintf(stderr, "psnw_create: %s\n", ctx->err.c_str());
        return bail(PSNW_ECUDA);
    }
    // scratch sized from the capacities given at create time
    const size_t arena = (size_t)p->max_points * 48 + (size_t)p->max_targets * 2048 + (1 << 16);
    if (ensure(ctx, ctx->arena, arena) != PSNW_OK || ensure(ctx, ctx->desc, (size_t)p->max_targets * 64 + (1 << 16)) != PSNW_OK)
        return bail(PSNW_ENOMEM);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(PSNW_ECUDA);
    *out = ctx;
    return PSNW_OK;
}

int psnw_get_params(psnw_ctx *ctx, psnw_params *out)
{
    if (!ctx || !out) return PSNW_EARG;
    *out = ctx->p;
    out->stream = (void *)ctx->stream;
    return PSNW_OK;
}

int psnw_destroy(psnw_ctx *ctx)
{
    if (!ctx) return PSNW_EARG;
    cudaSetDevice(ctx->p.device);
    if (ctx->aux) cudaStreamSynchronize(ctx->aux);
    if (ctx->up) cudaStreamSynchronize(ctx->up);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ev_up) cudaEventDestroy(ctx->ev_up);
    for (cudaEvent_t e : ctx->rd_ev) if (e) cudaEventDestroy(e);
    if (ctx->up) cudaStreamDestroy(ctx->up);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->det_desc.p) cudaFree(ctx->det_desc.p);
    for (int l = 0; l < PSNW_MAX_LEVELS; ++l) {
        if (ctx->img_alloc[l]) cudaFree(ctx->img_alloc[l]);
        if (ctx->pyr.lv[l].dxy) cudaFree(ctx->pyr.lv[l].dxy);
    }
    for (float *e : ctx->eig) if (e) cudaFree(e);
    for (uint8_t *e : ctx->fmap) if (e) cudaFree(e);
    if (ctx->arena.p) cudaFree(ctx->arena.p);
    if (ctx->desc.p) cudaFree(ctx->desc.p);
    if (ctx->lk_work) cudaFree(ctx->lk_work);
    if (ctx->walk_ctl) cudaFree(ctx->walk_ctl);
    if (ctx->lk_stats) cudaFree(ctx->lk_stats);
    if (ctx->lk_maps) cudaFree(ctx->lk_maps);
    if (ctx->cand.p) cudaFree(ctx->cand.p);
    if (ctx->bgr.p) cudaFree(ctx->bgr.p);
    for (StageSlot &s : ctx->stage) {
        if (s.h) cudaFreeHost(s.h);
        if (s.ev) cudaEventDestroy(s.ev);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PSNW_OK;
}

int psnw_push_frames(psnw_ctx *ctx, int n, const int *cams, const uint8_t *const *grays, size_t pitch,
                     uint32_t frame_idx, int mem_flags)
{
    (void)frame_idx;
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n <= 0 || !cams || !grays || pitch < (size_t)ctx->p.width) return fail(ctx, PSNW_EARG, "psnw_push_frames arguments");
    for (int i = 0; i < n; ++i)
        if (!grays[i]) return fail(ctx, PSNW_EARG, "null frame pointer");
    const bool overlap = (mem_flags & PSNW_MEM_OVERLAP) != 0 && (mem_flags & 0xff) != PSNW_MEM_HOST;
    const bool eager_eig = (mem_flags & PSNW_PUSH_EIG) != 0;
    mem_flags &= 0xff;
    std::vector<int> planes;
    RingTxn txn(ctx, n, cams);
    int rc = advance_ring(ctx, n, cams, planes, overlap);
    if (rc) { txn.commit(); return rc; }                 // (advance_ring validates before it rotates)
    cudaStream_t st;
    rc = push_stream(ctx, overlap, planes, &st);
    if (rc) return rc;
    const LevelGeom &l0 = ctx->pyr.lv[0];
    // Frames that sit at a constant stride in the caller's memory and go to planes at a constant stride (the usual
    // case: a [cameras][h][w] block, all rings in lockstep) travel as ONE copy, each frame being one "row" of it.
    bool one_copy = n > 1 && pitch == (size_t)l0.pitch && l0.w == l0.pitch;
    ptrdiff_t sstride = 0, dstride = 0;
    if (one_copy) {
        sstride = grays[1] - grays[0];
        dstride = ((ptrdiff_t)planes[1] - planes[0]) * l0.plane_px;
        // (a 2-D copy takes pitches up to cudaDevAttrMaxPitch = 2^31 - 1: frames that merely happen to lie at equal distances in
        // separate, far-apart allocations -- any two frames do -- keep one copy each)
        constexpr ptrdiff_t kMaxCopyPitch = 0x7fffffff;
        one_copy = sstride >= (ptrdiff_t)l0.plane_px && dstride >= (ptrdiff_t)l0.plane_px && sstride <= kMaxCopyPitch && dstride <= kMaxCopyPitch;
        for (int i = 2; one_copy && i < n; ++i)
            one_copy = (grays[i] - grays[i - 1]) == sstride && ((ptrdiff_t)planes[i] - planes[i - 1]) * l0.plane_px == dstride;
    }
    if (one_copy) {
        // Frames at equal distances are usually slices of one block -- but any two frames are "at equal distances", and a copy
        // whose source rows lie in DIFFERENT host allocations is refused by the runtime (invalid argument, nothing enqueued):
        // those take one copy each like everything else.
        if (cudaMemcpy2DAsync(l0.img + (size_t)planes[0] * l0.plane_px, (size_t)dstride, grays[0], (size_t)sstride, (size_t)l0.plane_px, n,
                              h2d(mem_flags), st) != cudaSuccess) {
            (void)cudaGetLastError();
            one_copy = false;
        }
    }
    if (!one_copy) {
        for (int i = 0; i < n; ++i)
            CK(cudaMemcpy2DAsync(l0.img + (size_t)planes[i] * l0.plane_px, l0.pitch, grays[i], pitch, l0.w, l0.h,
                                 h2d(mem_flags), st));
    }
    rc = build_pyramids(ctx, planes, st);
    if (rc) return rc;
    if (eager_eig) {                          // PSNW_PUSH_EIG: a detection on these frames will follow
        std::vector<std::pair<int, int>> cs;
        for (int i = 0; i < n; ++i) cs.emplace_back(cams[i], 0);
        rc = ensure_min_eig(ctx, cs, st, nullptr, overlap);
        if (rc) return rc;
        if (overlap) {                        // (on the context stream the image planes are ordered by the stream itself)
            rc = note_readers(ctx, st, planes.data(), (int)planes.size());
            if (rc) return rc;
        }
    }
    if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    rc = push_done(ctx, overlap);
    if (rc == PSNW_OK) txn.commit();
    return rc;
}

int psnw_push_frame(psnw_ctx *ctx, int cam, const uint8_t *gray, size_t pitch, uint32_t frame_idx, int mem_flags)
{
    return psnw_push_frames(ctx, 1, &cam, &gray, pitch, frame_idx, mem_flags);
}

int psnw_push_frames_bgr(psnw_ctx *ctx, int n, const int *cams, const uint8_t *const *bgr, size_t pitch,
                         uint32_t frame_idx, int luma_bits, int mem_flags)
{
    (void)frame_idx;
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n <= 0 || !cams || !bgr || pitch < (size_t)ctx->p.width * 3 || (luma_bits != 14 && luma_bits != 15))
        return fail(ctx, PSNW_EARG, "psnw_push_frames_bgr arguments");
    for (int i = 0; i < n; ++i)
        if (!bgr[i]) return fail(ctx, PSNW_EARG, "null frame pointer");
    const bool overlap = (mem_flags & PSNW_MEM_OVERLAP) != 0 && (mem_flags & 0xff) != PSNW_MEM_HOST;
    mem_flags &= 0xff;
    const LevelGeom &l0 = ctx->pyr.lv[0];
    const size_t row = ((size_t)l0.w * 3 + 255) & ~(size_t)255, frame = row * l0.h;
    if (mem_flags != PSNW_MEM_DEVICE) {
        int rc = ensure(ctx, ctx->bgr, frame * n);
        if (rc) return rc;
    }
    std::vector<int> planes;
    RingTxn txn(ctx, n, cams);
    int rc = advance_ring(ctx, n, cams, planes, overlap);
    if (rc) { txn.commit(); return rc; }
    cudaStream_t st;
    rc = push_stream(ctx, overlap, planes, &st);
    if (rc) return rc;
    if (overlap && mem_flags != PSNW_MEM_DEVICE) {      // the BGR staging buffer may still feed a conversion on the main stream
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(st, ctx->ev_fork, 0));
    }
    for (int i = 0; i < n; ++i) {
        const uint8_t *src = bgr[i];
        size_t src_pitch = pitch;
        if (mem_flags != PSNW_MEM_DEVICE) {
            uint8_t *dst = reinterpret_cast<uint8_t *>(ctx->bgr.p) + frame * i;
            CK(cudaMemcpy2DAsync(dst, row, bgr[i], pitch, (size_t)l0.w * 3, l0.h, cudaMemcpyHostToDevice, st));
            src = dst; src_pitch = row;
        }
        const int r = launch_bgr2gray(src, src_pitch, l0.img + (size_t)planes[i] * l0.plane_px, l0.pitch, l0.w, l0.h,
                                      luma_bits, st);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "bgr2gray_kernel launch", cudaGetLastError());
        ctx->launches += r;
    }
    rc = build_pyramids(ctx, planes, st);
    if (rc) return rc;
    if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    rc = push_done(ctx, overlap);
    if (rc == PSNW_OK) txn.commit();
    return rc;
}

int psnw_rebuild_pyramids(psnw_ctx *ctx, int n, const int *cams)
{
    if (!ctx || n <= 0 || !cams) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    std::vector<int> planes(n);
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(ctx, cams[i], 0)) return fail(ctx, PSNW_ESTATE, "psnw_rebuild_pyramids: camera holds no frame");
        planes[i] = plane_of(ctx, cams[i], 0);
    }
    if (int rc = join_aux(ctx)) return rc;
    return build_pyramids(ctx, planes, ctx->stream);
}

int psnw_rebuild_pyramids_slot(psnw_ctx *ctx, int n, const int *cams, int slot)
{
    if (!ctx || n <= 0 || !cams) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    std::vector<int> planes(n);
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(ctx, cams[i], slot)) return fail(ctx, PSNW_ESTATE, "psnw_rebuild_pyramids_slot: camera/slot holds no frame");
        planes[i] = plane_of(ctx, cams[i], slot);
    }
    if (int rc = join_aux(ctx)) return rc;
    return build_pyramids(ctx, planes, ctx->stream);
}

int psnw_lk_stats(psnw_ctx *ctx, int enable, uint64_t *out)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->lk_stats && out) CK(cudaMemcpy(out, ctx->lk_stats, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    else if (out) out[0] = out[1] = out[2] = 0;
    if (enable && !ctx->lk_stats) {
        if (cudaMalloc(&ctx->lk_stats, 3 * sizeof(unsigned long long)) != cudaSuccess) return fail(ctx, PSNW_ENOMEM, "cudaMalloc LK counters", cudaGetLastError());
    }
    if (ctx->lk_stats) CK(cudaMemset(ctx->lk_stats, 0, 3 * sizeof(unsigned long long)));
    if (!enable && ctx->lk_stats) { CK(cudaFree(ctx->lk_stats)); ctx->lk_stats = nullptr; }
    return PSNW_OK;
}

int psnw_frames_held(psnw_ctx *ctx, int cam)
{
    if (!ctx || cam < 0 || cam >= ctx->p.n_cameras) return PSNW_EARG;
    return ctx->held[cam];
}

int psnw_klt_track_batch(psnw_ctx *ctx, const psnw_track_job *jobs, int n_jobs, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (!jobs || n_jobs <= 0) return fail(ctx, PSNW_EARG, "psnw_klt_track_batch: no jobs");
    if (int rc_up = join_up(ctx)) return rc_up;          // an overlapped push may still be filling the slots tracked here
    mem_flags &= 0xff;
    const bool dev = mem_flags == PSNW_MEM_DEVICE;
    int total = 0, total_groups = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const psnw_track_job &J = jobs[j];
        if (J.n_pts < 0 || J.n_targets < 1 || (J.n_pts > 0 && (!J.pts_xy || !J.out_xy || !J.status || !J.err)))
            return fail(ctx, PSNW_EARG, "psnw_klt_track: bad job arrays");
        if (!slot_ok(ctx, J.cam, J.slot_from) || !slot_ok(ctx, J.cam, J.slot_to))
            return fail(ctx, PSNW_ESTATE, "psnw_klt_track: camera/slot holds no frame");
        total += J.n_pts;
        total_groups += J.n_targets;
    }
    if (total > ctx->p.max_points || total_groups > ctx->p.max_targets)
        return fail(ctx, PSNW_ENOMEM, "psnw_klt_track: max_points / max_targets given at create time exceeded");
    if (total == 0) return PSNW_OK;
    const bool host = mem_flags == PSNW_MEM_HOST && total <= kBlobMaxPts;    // one-blob transfers (see mirror())

    const bool do_fb = ctx->p.fb_thresh >= 0.f;
    std::vector<TrackJob> dj(n_jobs);
    std::vector<TrackGroup> dg(total_groups);
    ctx->arena_off = 0;
    // host / pinned arrays are staged in the arena BY ARRAY TYPE, the jobs' points back to back inside each: jobs whose host
    // arrays are slices of one array (a rig's cameras) then travel as one copy per array instead of one per job and array
    float2 *a_pts = nullptr, *a_out = nullptr;
    int *a_tgt = nullptr;
    uint8_t *a_st = nullptr;
    float *a_err = nullptr, *a_fb = nullptr;
    if (!dev) {
        bool any_tgt = false, any_fb = false;
        for (int j = 0; j < n_jobs; ++j) { any_tgt = any_tgt || jobs[j].target_of_pt; any_fb = any_fb || (do_fb_early(ctx) && jobs[j].fb_err); }
        a_pts = take<float2>(ctx, total); a_out = take<float2>(ctx, total);
        a_tgt = any_tgt ? take<int>(ctx, total) : nullptr;
        a_st = take<uint8_t>(ctx, total); a_err = take<float>(ctx, total);
        a_fb = any_fb ? take<float>(ctx, total) : nullptr;
        if (ctx->arena_off > ctx->arena.cap) return fail(ctx, PSNW_ENOMEM, "psnw_klt_track: scratch arena exhausted");
    }
    bool all21 = true, any_skipped = false;
    int max_win_px = 1;
    int pt = 0, gb = 0;
    const LevelGeom &l0 = ctx->pyr.lv[0];
    for (int j = 0; j < n_jobs; ++j) {
        const psnw_track_job &J = jobs[j];
        const int pf = plane_of(ctx, J.cam, J.slot_from), pt_to = plane_of(ctx, J.cam, J.slot_to);
        for (int t = 0; t < J.n_targets; ++t) {
            const int ww = J.win_wh ? J.win_wh[2 * t] : ctx->p.win_w, wh = J.win_wh ? J.win_wh[2 * t + 1] : ctx->p.win_h;
            TrackGroup &g = dg[gb + t];
            g.plane_from = pf; g.plane_to = pt_to; g.win_w = ww; g.win_h = wh;
            g.pad = 0;
            // A window this path cannot track -- degenerate (< 1), not smaller than the frame, or larger than the template
            // memory of the generic kernel -- disables ITS target only: the points come back unmoved with status 0 (the
            // reference tracks every target in its own call, Tracker2D.cpp:776, :871; one bad box must not void the frame).
            const bool untrackable = ww < 1 || wh < 1 || ww >= l0.w || wh >= l0.h || (((size_t)((ww + 3) & ~3) * wh + 7) & ~(size_t)7) * 6 > kGenericSmemMax;
            if (untrackable) { g.max_level = -1; any_skipped = true; continue; }
            g.max_level = effective_max_level(ctx, ww, wh);
            all21 = all21 && ww == 21 && wh == 21;
            max_win_px = std::max(max_win_px, ((ww + 3) & ~3) * wh);      // template rows are padded to 4 pixels (lk.cu)
        }
        TrackJob &D = dj[j];
        D.pt_begin = pt; D.n_pts = J.n_pts; D.group_base = gb; D.n_targets = J.n_targets;
        if (dev) {
            D.pts = reinterpret_cast<const float2 *>(J.pts_xy); D.target_of_pt = J.target_of_pt;
            D.out = reinterpret_cast<float2 *>(J.out_xy); D.status = J.status; D.err = J.err; D.fb_err = do_fb ? J.fb_err : nullptr;
        } else {
            D.pts = a_pts + pt; D.target_of_pt = J.target_of_pt ? a_tgt + pt : nullptr;
            D.out = a_out + pt; D.status = a_st + pt; D.err = a_err + pt;
            D.fb_err = (do_fb && J.fb_err) ? a_fb + pt : nullptr;
        }
        pt += J.n_pts; gb += J.n_targets;
    }

    if (!dev && !host) {                      // pinned (or oversize pageable) arrays: one upload per run of adjacent host slices
        for (int j = 0; j < n_jobs;) {
            int e = j + 1;
            size_t npts = (size_t)jobs[j].n_pts;
            while (e < n_jobs && jobs[e].pts_xy == jobs[j].pts_xy + 2 * npts) npts += (size_t)jobs[e++].n_pts;
            if (npts) CK(cudaMemcpyAsync(const_cast<float2 *>(dj[j].pts), jobs[j].pts_xy, sizeof(float2) * npts, cudaMemcpyHostToDevice, ctx->stream));
            j = e;
        }
        for (int j = 0; j < n_jobs; ++j)
            if (jobs[j].target_of_pt && jobs[j].n_pts > 0)
                CK(cudaMemcpyAsync(const_cast<int *>(dj[j].target_of_pt), jobs[j].target_of_pt, sizeof(int) * jobs[j].n_pts, cudaMemcpyHostToDevice, ctx->stream));
    }
    StageSlot *blob = nullptr;                // PSNW_MEM_HOST: all inputs up as one pinned blob (see mirror())
    if (host) {
        int rcb = stage_acquire(ctx, ctx->arena_off, &blob);
        if (rcb) return rcb;
        for (int j = 0; j < n_jobs; ++j) {
            const psnw_track_job &J = jobs[j];
            if (J.n_pts == 0) continue;
            memcpy(mirror(ctx, blob, dj[j].pts), J.pts_xy, sizeof(float2) * J.n_pts);
            if (dj[j].target_of_pt) memcpy(mirror(ctx, blob, dj[j].target_of_pt), J.target_of_pt, sizeof(int) * J.n_pts);
        }
        CK(cudaMemcpyAsync(ctx->arena.p, blob->h, ctx->arena_off, cudaMemcpyHostToDevice, ctx->stream));
    }

    // descriptors: inside the kernel parameters when they fit, else through a pinned staging slot -> device
    const bool embed = n_jobs <= kEmbedJobs && total_groups <= kEmbedGroups;
    const size_t jb = sizeof(TrackJob) * n_jobs, gbytes = sizeof(TrackGroup) * total_groups;
    const size_t goff = (jb + 255) & ~(size_t)255;
    int rc = PSNW_OK;
    if (!embed) {
        rc = ensure(ctx, ctx->desc, goff + gbytes);
        if (rc) return rc;
        StageSlot *s;
        rc = stage_acquire(ctx, goff + gbytes, &s);
        if (rc) return rc;
        memcpy(s->h, dj.data(), jb);
        memcpy(s->h + goff, dg.data(), gbytes);
        rc = stage_upload(ctx, s, ctx->desc.p, goff + gbytes);
        if (rc) return rc;
    }

    LkParams P;
    memcpy(P.lv, ctx->pyr.lv, sizeof P.lv);
    P.jobs = embed ? nullptr : reinterpret_cast<const TrackJob *>(ctx->desc.p);
    P.groups = embed ? nullptr : reinterpret_cast<const TrackGroup *>(ctx->desc.p + goff);
    if (embed) {
        memcpy(P.ejobs, dj.data(), jb);
        memcpy(P.egroups, dg.data(), gbytes);
    }
    P.n_jobs = n_jobs; P.n_pts = total;
    P.max_iters = ctx->p.max_iters;
    P.eps2 = ctx->p.epsilon * ctx->p.epsilon;
    P.eps2_lo = (float)(P.eps2 * (1.0 - 1e-5));
    P.eps2_hi = (float)(P.eps2 * (1.0 + 1e-5));
    P.min_eig = ctx->p.min_eig_thresh;
    P.fb_thresh = ctx->p.fb_thresh;
    P.do_fb = do_fb ? 1 : 0;
    P.work = ctx->lk_persistent ? ctx->lk_work : nullptr;
    // (a batch that is all tail would only add the hand-over: small batches keep one item per point)
    P.split_from = (P.work && P.do_fb && all21 && !any_skipped && P.n_pts >= 2 * ctx->lk_tail) ? P.n_pts - ctx->lk_tail : P.n_pts;
    P.tmaps = ctx->lk_maps;
    P.stats = ctx->lk_stats;
    const int r = launch_lk(P, all21 ? 1 : 0, max_win_px, ctx->lk_ctas, ctx->stream);
    if (r == -2) return fail(ctx, PSNW_EARG, "psnw_klt_track: window too large for shared memory");      // (unreachable: such targets are skipped)
    if (r < 0) return fail(ctx, PSNW_ECUDA, "lk kernel launch", cudaGetLastError());
    ctx->launches += r;
    {
        std::vector<int> rd;
        for (const TrackGroup &g : dg) {
            if (rd.empty() || rd[rd.size() - 2] != g.plane_from || rd.back() != g.plane_to) { rd.push_back(g.plane_from); rd.push_back(g.plane_to); }
        }
        rc = note_readers(ctx, ctx->stream, rd.data(), (int)rd.size());
        if (rc) return rc;
    }

    if (host) {                               // ... and all outputs back as one
        CK(cudaMemcpyAsync(blob->h, ctx->arena.p, ctx->arena_off, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < n_jobs; ++j) {
            const psnw_track_job &J = jobs[j];
            const TrackJob &D = dj[j];
            if (J.n_pts == 0) continue;
            memcpy(J.out_xy, mirror(ctx, blob, D.out), sizeof(float2) * J.n_pts);
            memcpy(J.status, mirror(ctx, blob, D.status), J.n_pts);
            memcpy(J.err, mirror(ctx, blob, D.err), sizeof(float) * J.n_pts);
            if (D.fb_err) memcpy(J.fb_err, mirror(ctx, blob, D.fb_err), sizeof(float) * J.n_pts);
        }
        return PSNW_OK;
    }
    if (!dev) {
        // downloads: one copy per array and run of jobs whose host arrays are adjacent slices
        for (int j = 0; j < n_jobs;) {
            const psnw_track_job &J = jobs[j];
            int e = j + 1;
            size_t npts = (size_t)J.n_pts;
            while (e < n_jobs && jobs[e].out_xy == J.out_xy + 2 * npts && jobs[e].status == J.status + npts && jobs[e].err == J.err + npts &&
                   ((dj[j].fb_err != nullptr) == (dj[e].fb_err != nullptr)) && (!dj[j].fb_err || jobs[e].fb_err == J.fb_err + npts))
                npts += (size_t)jobs[e++].n_pts;
            if (npts) {
                const TrackJob &D = dj[j];
                CK(cudaMemcpyAsync(J.out_xy, D.out, sizeof(float2) * npts, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(J.status, D.status, npts, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(J.err, D.err, sizeof(float) * npts, cudaMemcpyDeviceToHost, ctx->stream));
                if (D.fb_err) CK(cudaMemcpyAsync(J.fb_err, D.fb_err, sizeof(float) * npts, cudaMemcpyDeviceToHost, ctx->stream));
            }
            j = e;
        }
        if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    }
    return PSNW_OK;
}

int psnw_klt_track(psnw_ctx *ctx, int cam, int slot_from, int slot_to, const float *pts_xy, const int *target_of_pt,
                   int n_pts, const int *win_wh, int n_targets, float *out_xy, uint8_t *status, float *err,
                   float *fb_err, int mem_flags)
{
    psnw_track_job j;
    j.cam = cam; j.slot_from = slot_from; j.slot_to = slot_to;
    j.n_pts = n_pts; j.pts_xy = pts_xy; j.target_of_pt = target_of_pt;
    j.n_targets = n_targets < 1 ? 1 : n_targets; j.win_wh = win_wh;
    j.out_xy = out_xy; j.status = status; j.err = err; j.fb_err = fb_err;
    return psnw_klt_track_batch(ctx, &j, 1, mem_flags);
}

int psnw_box_update(psnw_ctx *ctx, const double *boxes_in, int n_targets, const float *pre_xy, const float *cur_xy,
                    const uint8_t *status, const int *target_offsets, int mode, double *boxes_out,
                    uint8_t *inlier_mask, int *inlier_counts, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_targets == 0) return PSNW_OK;
    if (n_targets < 0 || !boxes_in || !target_offsets || !boxes_out || !inlier_mask || !inlier_counts || (mode != PSNW_BOX_REFERENCE && mode != PSNW_BOX_MEDIANFLOW))
        return fail(ctx, PSNW_EARG, "psnw_box_update arguments");
    if (n_targets > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_box_update: max_targets exceeded");
    mem_flags &= 0xff;
    int r;
    if (mem_flags == PSNW_MEM_DEVICE) {
        r = launch_box_update(boxes_in, n_targets, reinterpret_cast<const float2 *>(pre_xy), reinterpret_cast<const float2 *>(cur_xy),
                              status, target_offsets, boxes_out, inlier_mask, inlier_counts, PSNW_MAX_PTS_PER_TARGET, ctx->stream, mode);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "box_update_kernel launch", cudaGetLastError());
        ctx->launches += r;
        return PSNW_OK;
    }
    const int n_pts = target_offsets[n_targets];
    int max_per = 1;
    for (int t = 0; t < n_targets; ++t) {
        const int k = target_offsets[t + 1] - target_offsets[t];
        if (k < 0 || k > PSNW_MAX_PTS_PER_TARGET) return fail(ctx, PSNW_EARG, "psnw_box_update: offsets not ascending or > PSNW_MAX_PTS_PER_TARGET points in a target");
        max_per = std::max(max_per, k);
    }
    if (n_pts > ctx->p.max_points) return fail(ctx, PSNW_ENOMEM, "psnw_box_update: max_points exceeded");
    if (n_pts > 0 && (!pre_xy || !cur_xy)) return fail(ctx, PSNW_EARG, "psnw_box_update: null points");
    ctx->arena_off = 0;
    double *d_in = take<double>(ctx, (size_t)n_targets * 4), *d_out = take<double>(ctx, (size_t)n_targets * 4);
    int *d_off = take<int>(ctx, n_targets + 1), *d_cnt = take<int>(ctx, n_targets);
    float2 *d_pre = take<float2>(ctx, n_pts), *d_cur = take<float2>(ctx, n_pts);
    uint8_t *d_st = status ? take<uint8_t>(ctx, n_pts) : nullptr, *d_mask = take<uint8_t>(ctx, n_pts);
    if (ctx->arena_off > ctx->arena.cap) return fail(ctx, PSNW_ENOMEM, "psnw_box_update: scratch arena exhausted");
    if (mem_flags == PSNW_MEM_HOST && n_pts <= kBlobMaxPts) {
        StageSlot *s;
        int rc = stage_acquire(ctx, ctx->arena_off, &s);
        if (rc) return rc;
        memcpy(mirror(ctx, s, d_in), boxes_in, sizeof(double) * 4 * n_targets);
        memcpy(mirror(ctx, s, d_off), target_offsets, sizeof(int) * (n_targets + 1));
        if (n_pts > 0) {
            memcpy(mirror(ctx, s, d_pre), pre_xy, sizeof(float2) * n_pts);
            memcpy(mirror(ctx, s, d_cur), cur_xy, sizeof(float2) * n_pts);
            if (d_st) memcpy(mirror(ctx, s, d_st), status, n_pts);
        }
        CK(cudaMemcpyAsync(ctx->arena.p, s->h, ctx->arena_off, cudaMemcpyHostToDevice, ctx->stream));
        r = launch_box_update(d_in, n_targets, d_pre, d_cur, d_st, d_off, d_out, d_mask, d_cnt, max_per, ctx->stream, mode);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "box_update_kernel launch", cudaGetLastError());
        ctx->launches += r;
        CK(cudaMemcpyAsync(s->h, ctx->arena.p, ctx->arena_off, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        memcpy(boxes_out, mirror(ctx, s, d_out), sizeof(double) * 4 * n_targets);
        memcpy(inlier_counts, mirror(ctx, s, d_cnt), sizeof(int) * n_targets);
        if (n_pts > 0) memcpy(inlier_mask, mirror(ctx, s, d_mask), n_pts);
        return PSNW_OK;
    }
    CK(cudaMemcpyAsync(d_in, boxes_in, sizeof(double) * 4 * n_targets, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off, target_offsets, sizeof(int) * (n_targets + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (n_pts > 0) {
        CK(cudaMemcpyAsync(d_pre, pre_xy, sizeof(float2) * n_pts, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_cur, cur_xy, sizeof(float2) * n_pts, cudaMemcpyHostToDevice, ctx->stream));
        if (d_st) CK(cudaMemcpyAsync(d_st, status, n_pts, cudaMemcpyHostToDevice, ctx->stream));
    }
    r = launch_box_update(d_in, n_targets, d_pre, d_cur, d_st, d_off, d_out, d_mask, d_cnt, max_per, ctx->stream, mode);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "box_update_kernel launch", cudaGetLastError());
    ctx->launches += r;
    CK(cudaMemcpyAsync(boxes_out, d_out, sizeof(double) * 4 * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(inlier_counts, d_cnt, sizeof(int) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_pts > 0) CK(cudaMemcpyAsync(inlier_mask, d_mask, n_pts, cudaMemcpyDeviceToHost, ctx->stream));
    if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return PSNW_OK;
}

// min-eigenvalue maps of the (cam, slot) frames in `cs`, computed once per pushed frame (tag = push sequence)

int psnw_klt_detect_batch(psnw_ctx *ctx, const psnw_detect_job *jobs, int n_jobs, int max_per_target,
                          float quality, float min_dist, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_jobs <= 0) return PSNW_OK;
    if (!jobs || max_per_target < 1 || !(quality > 0.f) || min_dist < 0.f) return fail(ctx, PSNW_EARG, "psnw_klt_detect arguments");
    int total = 0;
    std::vector<std::pair<int, int>> cs;
    for (int j = 0; j < n_jobs; ++j) {
        const psnw_detect_job &J = jobs[j];
        if (J.n_targets < 0 || (J.n_targets > 0 && (!J.boxes || !J.pts_xy || !J.counts))) return fail(ctx, PSNW_EARG, "psnw_klt_detect: bad job arrays");
        if (J.n_targets == 0) continue;
        if (!slot_ok(ctx, J.cam, J.slot)) return fail(ctx, PSNW_ESTATE, "psnw_klt_detect: camera/slot holds no frame");
        for (int k = 0; k < j; ++k)
            if (jobs[k].n_targets > 0 && jobs[k].cam == J.cam && jobs[k].slot != J.slot)
                return fail(ctx, PSNW_EARG, "psnw_klt_detect_batch: one slot per camera per call");
        total += J.n_targets;
        cs.emplace_back(J.cam, J.slot);
    }
    if (total == 0) return PSNW_OK;
    if (total > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_klt_detect: max_targets exceeded");
    const LevelGeom &l0 = ctx->pyr.lv[0];
    // PSNW_MEM_OVERLAP: run on the auxiliary stream, forked from the main stream here and joined by psnw_join /
    // psnw_sync / the next push or detect, so that the following track + box_update calls overlap with it
    const bool overlap = (mem_flags & PSNW_MEM_OVERLAP) != 0 && (mem_flags & 0xff) != PSNW_MEM_HOST;
    mem_flags &= 0xff;
    int rc = join_aux(ctx);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    if (overlap) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
        st = ctx->aux;
    }
    std::vector<int> eig_buf;                 // per cs entry: the buffer holding its Shi-Tomasi map
    rc = ensure_min_eig(ctx, cs, st, &eig_buf);
    if (rc) return rc;
    std::vector<int> eig_of_cam(ctx->p.n_cameras, -1);
    for (size_t i = 0; i < cs.size(); ++i) eig_of_cam[cs[i].first] = eig_buf[i];
    {
        std::vector<int> rd;
        for (const auto &c : cs) rd.push_back(plane_of(ctx, c.first, c.second));
        rc = note_readers(ctx, st, rd.data(), (int)rd.size());
        if (rc) return rc;
    }

    // mask ROI = box.cropWithSize(cols, rows).cv()  (PSNWhere_Types.h:146-153,181; Tracker2D.cpp:735).  `boxes` is
    // always host memory: detections come from the host-side detector files (PSNWhere.cpp:261).
    const bool dev = mem_flags == PSNW_MEM_DEVICE;
    const size_t tb = (sizeof(DetectTarget) * total + 255) & ~(size_t)255;
    const size_t out_b = dev ? 0 : (((size_t)total * max_per_target * sizeof(float2) + 255) & ~(size_t)255);
    const size_t cnt_b = dev ? 0 : ((sizeof(int) * total + 255) & ~(size_t)255);
    rc = ensure(ctx, ctx->det_desc, tb + out_b + cnt_b);
    if (rc) return rc;
    float2 *s_out = reinterpret_cast<float2 *>(ctx->det_desc.p + tb);
    int *s_cnt = reinterpret_cast<int *>(ctx->det_desc.p + tb + out_b);

    std::vector<DetectTarget> T(total);
    size_t cand_total = 0;
    int max_area = 1, max_tile = 1, t0 = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const psnw_detect_job &J = jobs[j];
        for (int t = 0; t < J.n_targets; ++t) {
            const double x = J.boxes[4 * t], y = J.boxes[4 * t + 1], w = J.boxes[4 * t + 2], h = J.boxes[4 * t + 3];
            const double nx = std::max(0.0, x), ny = std::max(0.0, y);
            const double nw = std::min((double)l0.w - nx - 1, w), nh = std::min((double)l0.h - ny - 1, h);
            DetectTarget &d = T[t0 + t];
            d.x0 = (int)nx; d.y0 = (int)ny;
            d.x1 = d.x0 + std::max((int)nw, 0); d.y1 = d.y0 + std::max((int)nh, 0);
            d.x0 = std::min(d.x0, l0.w); d.y0 = std::min(d.y0, l0.h);
            d.x1 = std::min(std::max(d.x1, d.x0), l0.w); d.y1 = std::min(std::max(d.y1, d.y0), l0.h);
            d.cand_begin = (int)cand_total;
            d.cand_cap = (d.x1 - d.x0) * (d.y1 - d.y0);
            cand_total += (size_t)d.cand_cap;
            max_area = std::max(max_area, d.cand_cap);
            max_tile = std::max(max_tile, (d.x1 - d.x0 + 2) * (d.y1 - d.y0 + 2));
            d.eig = ctx->eig[eig_of_cam[J.cam]];
            d.out = dev ? reinterpret_cast<float2 *>(J.pts_xy) + (size_t)t * max_per_target : s_out + (size_t)(t0 + t) * max_per_target;
            d.count = dev ? J.counts + t : s_cnt + t0 + t;
        }
        t0 += J.n_targets;
    }
    if (cand_total > ((size_t)1 << 30)) return fail(ctx, PSNW_ENOMEM, "psnw_klt_detect: candidate scratch too large");
    rc = ensure(ctx, ctx->cand, cand_total * sizeof(unsigned long long) + 256);
    if (rc) return rc;
    StageSlot *s;
    rc = stage_acquire(ctx, tb, &s);
    if (rc) return rc;
    memcpy(s->h, T.data(), sizeof(DetectTarget) * total);
    rc = stage_upload(ctx, s, ctx->det_desc.p, sizeof(DetectTarget) * total, st);
    if (rc) return rc;
    const int r = launch_detect_select(l0.w, l0.h, l0.pitch, reinterpret_cast<const DetectTarget *>(ctx->det_desc.p), total,
                                       max_per_target, quality, min_dist, reinterpret_cast<unsigned long long *>(ctx->cand.p),
                                       max_area, max_tile, st);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "detect_select_kernel launch", cudaGetLastError());
    ctx->launches += r;
    {                                         // the selection reads the min-eig buffers: an eager map of a later push waits for it
        std::vector<int> rd;
        for (int b : eig_buf) rd.push_back(ctx->planes + b);
        rc = note_readers(ctx, st, rd.data(), (int)rd.size());
        if (rc) return rc;
    }
    if (!dev) {
        t0 = 0;
        for (int j = 0; j < n_jobs;) {                 // one copy per array and run of jobs whose host arrays are adjacent slices
            const psnw_detect_job &J = jobs[j];
            int e = j + 1;
            size_t nt = (size_t)std::max(J.n_targets, 0);
            while (e < n_jobs && (jobs[e].n_targets == 0 || (nt > 0 && jobs[e].pts_xy == J.pts_xy + 2 * nt * max_per_target && jobs[e].counts == J.counts + nt)))
                nt += (size_t)std::max(jobs[e++].n_targets, 0);
            if (nt) {
                CK(cudaMemcpyAsync(J.pts_xy, s_out + (size_t)t0 * max_per_target, nt * max_per_target * sizeof(float2), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(J.counts, s_cnt + t0, sizeof(int) * nt, cudaMemcpyDeviceToHost, st));
            }
            t0 += (int)nt;
            j = e;
        }
        if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    }
    if (overlap) {
        CK(cudaEventRecord(ctx->ev_join, ctx->aux));
        ctx->aux_pending = true;
    }
    return PSNW_OK;
}

int psnw_klt_detect(psnw_ctx *ctx, int cam, int slot, const double *boxes, int n_targets, int max_per_target,
                    float quality, float min_dist, float *pts_xy, int *counts, int mem_flags)
{
    psnw_detect_job j;
    j.cam = cam; j.slot = slot; j.boxes = boxes; j.n_targets = n_targets; j.pts_xy = pts_xy; j.counts = counts;
    return psnw_klt_detect_batch(ctx, &j, 1, max_per_target, quality, min_dist, mem_flags);
}

// "next" row f-3: per (detection, tracker) count of tracked points inside the detection box (Tracker2D.cpp:939-944)
int psnw_klt_track_update(psnw_ctx *ctx, int cam, int slot_from, int slot_to, const float *pts_xy, const int *target_offsets,
                          int n_targets, const int *win_wh, const double *boxes_in, int use_status, float *out_xy,
                          uint8_t *status, float *err, double *boxes_out, uint8_t *inlier_mask, int *inlier_counts, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_targets == 0) return PSNW_OK;
    if (n_targets < 0 || !target_offsets || !boxes_in || !boxes_out || !inlier_mask || !inlier_counts)
        return fail(ctx, PSNW_EARG, "psnw_klt_track_update arguments");
    mem_flags &= 0xff;
    if (target_offsets[0] != 0) return fail(ctx, PSNW_EARG, "psnw_klt_track_update: target_offsets[0] must be 0");
    for (int t = 0; t < n_targets; ++t) {
        const int k = target_offsets[t + 1] - target_offsets[t];
        if (k < 0 || k > PSNW_MAX_PTS_PER_TARGET) return fail(ctx, PSNW_EARG, "psnw_klt_track_update: offsets not ascending or > PSNW_MAX_PTS_PER_TARGET points in a target");
    }
    if (mem_flags != PSNW_MEM_HOST) {
        // pinned / device memory: nothing to merge, the two calls already queue back to back on the context stream
        if (mem_flags == PSNW_MEM_DEVICE) return fail(ctx, PSNW_EARG, "psnw_klt_track_update: with device memory call psnw_klt_track_batch and psnw_box_update (target_offsets must be readable on the host)");
        const int n = target_offsets[n_targets];
        std::vector<int> tgt(std::max(n, 1));
        for (int t = 0; t < n_targets; ++t)
            for (int i = target_offsets[t]; i < target_offsets[t + 1]; ++i) tgt[i] = t;
        // (pinned memory is host memory: the pair runs as two synchronous calls, target_of_pt lives on this stack frame)
        int rc = psnw_klt_track(ctx, cam, slot_from, slot_to, pts_xy, tgt.data(), n, win_wh, n_targets, out_xy, status, err, nullptr, PSNW_MEM_HOST);
        if (rc) return rc;
        return psnw_box_update(ctx, boxes_in, n_targets, pts_xy, out_xy, use_status ? status : nullptr, target_offsets, PSNW_BOX_REFERENCE,
                               boxes_out, inlier_mask, inlier_counts, PSNW_MEM_HOST);
    }
    const int n_pts = target_offsets[n_targets];
    int max_per = 1;
    for (int t = 0; t < n_targets; ++t) max_per = std::max(max_per, target_offsets[t + 1] - target_offsets[t]);
    if (n_pts <= 0)                               // nothing to track: the boxes come back as LocalSearchKLT returns them for no features
        return psnw_box_update(ctx, boxes_in, n_targets, pts_xy, out_xy, nullptr, target_offsets, PSNW_BOX_REFERENCE, boxes_out,
                               inlier_mask, inlier_counts, PSNW_MEM_HOST);
    if (!pts_xy || !out_xy || !status || !err) return fail(ctx, PSNW_EARG, "psnw_klt_track_update: null point arrays");
    if (n_pts > ctx->p.max_points || n_targets > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_klt_track_update: max_points / max_targets exceeded");
    if (!slot_ok(ctx, cam, slot_from) || !slot_ok(ctx, cam, slot_to)) return fail(ctx, PSNW_ESTATE, "psnw_klt_track_update: camera/slot holds no frame");

    // scratch layout (one blob up, one blob down -- see mirror()); the track call below runs on these device arrays
    ctx->arena_off = 0;
    float2 *d_pts = take<float2>(ctx, n_pts), *d_out = take<float2>(ctx, n_pts);
    int *d_tgt = take<int>(ctx, n_pts), *d_off = take<int>(ctx, n_targets + 1), *d_cnt = take<int>(ctx, n_targets);
    uint8_t *d_st = take<uint8_t>(ctx, n_pts), *d_mask = take<uint8_t>(ctx, n_pts);
    float *d_err = take<float>(ctx, n_pts);
    float *d_fb = ctx->p.fb_thresh >= 0.f ? take<float>(ctx, n_pts) : nullptr;
    double *d_in = take<double>(ctx, (size_t)n_targets * 4), *d_bout = take<double>(ctx, (size_t)n_targets * 4);
    const size_t blob_bytes = ctx->arena_off;
    if (blob_bytes > ctx->arena.cap) return fail(ctx, PSNW_ENOMEM, "psnw_klt_track_update: scratch arena exhausted");
    StageSlot *s;
    int rc = stage_acquire(ctx, blob_bytes, &s);
    if (rc) return rc;
    memcpy(mirror(ctx, s, d_pts), pts_xy, sizeof(float2) * n_pts);
    memcpy(mirror(ctx, s, d_off), target_offsets, sizeof(int) * (n_targets + 1));
    memcpy(mirror(ctx, s, d_in), boxes_in, sizeof(double) * 4 * n_targets);
    int *h_tgt = mirror(ctx, s, d_tgt);
    for (int t = 0; t < n_targets; ++t)
        for (int i = target_offsets[t]; i < target_offsets[t + 1]; ++i) h_tgt[i] = t;
    CK(cudaMemcpyAsync(ctx->arena.p, s->h, blob_bytes, cudaMemcpyHostToDevice, ctx->stream));

    psnw_track_job j;
    j.cam = cam; j.slot_from = slot_from; j.slot_to = slot_to;
    j.n_pts = n_pts; j.pts_xy = reinterpret_cast<const float *>(d_pts); j.target_of_pt = d_tgt;
    j.n_targets = n_targets; j.win_wh = win_wh;
    j.out_xy = reinterpret_cast<float *>(d_out); j.status = d_st; j.err = d_err; j.fb_err = d_fb;
    rc = psnw_klt_track_batch(ctx, &j, 1, PSNW_MEM_DEVICE);         // (takes nothing from the arena with device pointers)
    if (rc) return rc;
    const int r = launch_box_update(d_in, n_targets, d_pts, d_out, use_status ? d_st : nullptr, d_off, d_bout, d_mask, d_cnt, max_per, ctx->stream);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "box_update_kernel launch", cudaGetLastError());
    ctx->launches += r;
    CK(cudaMemcpyAsync(s->h, ctx->arena.p, blob_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(out_xy, mirror(ctx, s, d_out), sizeof(float2) * n_pts);
    memcpy(status, mirror(ctx, s, d_st), n_pts);
    memcpy(err, mirror(ctx, s, d_err), sizeof(float) * n_pts);
    memcpy(boxes_out, mirror(ctx, s, d_bout), sizeof(double) * 4 * n_targets);
    memcpy(inlier_mask, mirror(ctx, s, d_mask), n_pts);
    memcpy(inlier_counts, mirror(ctx, s, d_cnt), sizeof(int) * n_targets);
    return PSNW_OK;
}

// The reference's backward chain (Tracker2D.cpp:763-811) for every detection of a camera without a host round trip between
// the dependent steps: per step one LK launch, one LocalSearchKLT launch and one compaction launch, all enqueued at once.
int psnw_klt_backward_chain(psnw_ctx *ctx, int cam, int n_steps, const float *pts_xy, const int *counts, int n_targets, int cap,
                            const int *win_wh, const double *boxes_in, int min_inliers, double *boxes_steps, int *steps_done,
                            float *pts_steps, int *counts_steps, uint8_t *first_inlier_mask, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_targets == 0) return PSNW_OK;
    if (n_targets < 0 || cap < 1 || cap > PSNW_MAX_PTS_PER_TARGET || n_steps < 1 || !pts_xy || !counts || !boxes_in || !boxes_steps ||
        !steps_done || !pts_steps || !counts_steps || !first_inlier_mask)
        return fail(ctx, PSNW_EARG, "psnw_klt_backward_chain arguments");
    if (cam < 0 || cam >= ctx->p.n_cameras) return fail(ctx, PSNW_EARG, "psnw_klt_backward_chain: camera index out of range");
    mem_flags &= 0xff;
    const bool dev = mem_flags == PSNW_MEM_DEVICE;
    const size_t n = (size_t)n_targets * cap;
    if (n > (size_t)ctx->p.max_points || n_targets > ctx->p.max_targets)
        return fail(ctx, PSNW_ENOMEM, "psnw_klt_backward_chain: n_targets * cap exceeds max_points (or n_targets max_targets)");
    // the reference stops at the first empty FIFO slot (:768-769): steps beyond the frames held are simply not run
    const int steps = std::min(n_steps, ctx->held[cam] - 1);
    if (ctx->held[cam] < 1) return fail(ctx, PSNW_ESTATE, "psnw_klt_backward_chain: camera holds no frame");

    // scratch (the LK launches below run on device pointers and take nothing from the arena)
    const size_t a256 = 256;
    size_t need = 16 * a256 + n * 8 /*P0*/ + n * (8 + 1 + 4 + 4 + 1 + 1 + 4) /*O, st, err, fb, live x2, tgt*/ + (size_t)n_targets * (32 + 16) + (size_t)(n_targets + 1) * 4;
    if (!dev) need += (size_t)n_steps * (n * 8 + (size_t)n_targets * (32 + 4) + 2 * a256) + n + (size_t)n_targets * 36 + 4 * a256;
    int rc = ensure(ctx, ctx->arena, need + (1 << 16));
    if (rc) return rc;
    ctx->arena_off = 0;
    float2 *P0 = take<float2>(ctx, n), *O = take<float2>(ctx, n);
    uint8_t *st = take<uint8_t>(ctx, n), *liveA = take<uint8_t>(ctx, n), *liveB = take<uint8_t>(ctx, n), *mask_k = take<uint8_t>(ctx, n);
    float *err = take<float>(ctx, n), *fb = ctx->p.fb_thresh >= 0.f ? take<float>(ctx, n) : nullptr;
    int *tgt = take<int>(ctx, n), *offs = take<int>(ctx, n_targets + 1), *alive = take<int>(ctx, n_targets), *ninl = take<int>(ctx, n_targets);
    double *box_k = take<double>(ctx, (size_t)n_targets * 4);
    int *d_counts = dev ? const_cast<int *>(counts) : take<int>(ctx, n_targets);
    double *d_boxes_in = dev ? const_cast<double *>(boxes_in) : take<double>(ctx, (size_t)n_targets * 4);
    double *d_boxes_steps = dev ? boxes_steps : take<double>(ctx, (size_t)n_steps * n_targets * 4);
    int *d_steps_done = dev ? steps_done : take<int>(ctx, n_targets);
    float2 *d_pts_steps = dev ? reinterpret_cast<float2 *>(pts_steps) : take<float2>(ctx, (size_t)n_steps * n);
    int *d_counts_steps = dev ? counts_steps : take<int>(ctx, (size_t)n_steps * n_targets);
    uint8_t *d_first = dev ? first_inlier_mask : take<uint8_t>(ctx, n);
    if (ctx->arena_off > ctx->arena.cap) return fail(ctx, PSNW_ENOMEM, "psnw_klt_backward_chain: scratch arena exhausted");

    CK(cudaMemcpyAsync(P0, pts_xy, n * sizeof(float2), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (!dev) {
        CK(cudaMemcpyAsync(d_counts, counts, sizeof(int) * n_targets, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_boxes_in, boxes_in, sizeof(double) * 4 * n_targets, cudaMemcpyHostToDevice, ctx->stream));
    }
    // steps that are not run report no points / no boxes
    CK(cudaMemsetAsync(d_counts_steps, 0, sizeof(int) * (size_t)n_steps * n_targets, ctx->stream));
    CK(cudaMemsetAsync(d_boxes_steps, 0, sizeof(double) * 4 * (size_t)n_steps * n_targets, ctx->stream));
    CK(cudaMemsetAsync(d_first, 0, n, ctx->stream));
    int r = launch_chain_init(P0, d_counts, n_targets, cap, liveA, tgt, offs, alive, d_steps_done, ctx->stream);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "chain_init_kernel launch", cudaGetLastError());
    ctx->launches += r;

    const float2 *cur = P0;
    uint8_t *live = liveA, *live_next = liveB;
    for (int k = 1; k <= steps; ++k) {
        psnw_track_job j;
        j.cam = cam; j.slot_from = k - 1; j.slot_to = k;
        j.n_pts = (int)n; j.pts_xy = reinterpret_cast<const float *>(cur); j.target_of_pt = tgt;
        j.n_targets = n_targets; j.win_wh = win_wh;
        j.out_xy = reinterpret_cast<float *>(O); j.status = st; j.err = err; j.fb_err = fb;
        rc = psnw_klt_track_batch(ctx, &j, 1, PSNW_MEM_DEVICE);
        if (rc) return rc;
        uint8_t *mask = k == 1 ? d_first : mask_k;
        r = launch_box_update(d_boxes_in, n_targets, cur, O, live, offs, box_k, mask, ninl, cap, ctx->stream);       // :787, always the detection box (:770)
        if (r < 0) return fail(ctx, PSNW_ECUDA, "box_update_kernel launch", cudaGetLastError());
        ctx->launches += r;
        float2 *next = d_pts_steps + (size_t)(k - 1) * n;
        r = launch_chain_compact(O, mask, ninl, box_k, n_targets, cap, min_inliers, k, next, live_next, d_counts_steps + (size_t)(k - 1) * n_targets,
                                 d_boxes_steps + (size_t)(k - 1) * n_targets * 4, alive, d_steps_done, ctx->stream);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "chain_compact_kernel launch", cudaGetLastError());
        ctx->launches += r;
        cur = next;
        std::swap(live, live_next);
    }
    if (!dev) {
        CK(cudaMemcpyAsync(boxes_steps, d_boxes_steps, sizeof(double) * 4 * (size_t)n_steps * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(steps_done, d_steps_done, sizeof(int) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        if (steps > 0) CK(cudaMemcpyAsync(pts_steps, d_pts_steps, sizeof(float2) * (size_t)steps * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(counts_steps, d_counts_steps, sizeof(int) * (size_t)n_steps * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(first_inlier_mask, d_first, n, cudaMemcpyDeviceToHost, ctx->stream));
        if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    }
    return PSNW_OK;
}

// "next" row f-3 in full: the optical-flow majority check of Tracker2D.cpp:931-944 / :981-1022 on the device
int psnw_flow_majority(psnw_ctx *ctx, const float *pts_xy, const int *tracker_offsets, int n_trackers, const double *tracker_boxes,
                       const uint8_t *tracker_ok, const double *det_boxes, int n_detections, double ratio, int *counts,
                       uint8_t *wipe, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_trackers == 0 || n_detections == 0) return PSNW_OK;
    if (n_trackers < 0 || n_detections < 0 || !tracker_offsets || !tracker_boxes || !tracker_ok || !det_boxes || !counts || !wipe)
        return fail(ctx, PSNW_EARG, "psnw_flow_majority arguments");
    if (n_trackers > ctx->p.max_targets || n_detections > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_flow_majority: max_targets exceeded");
    mem_flags &= 0xff;
    int r;
    if (mem_flags == PSNW_MEM_DEVICE) {
        r = launch_flow_majority(reinterpret_cast<const float2 *>(pts_xy), tracker_offsets, n_trackers, tracker_boxes, tracker_ok, det_boxes,
                                 n_detections, ratio, counts, wipe, ctx->stream);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "flow majority launch", cudaGetLastError());
        ctx->launches += r;
        return PSNW_OK;
    }
    if (tracker_offsets[0] != 0) return fail(ctx, PSNW_EARG, "psnw_flow_majority: tracker_offsets[0] must be 0");
    for (int t = 0; t < n_trackers; ++t)
        if (tracker_offsets[t + 1] < tracker_offsets[t]) return fail(ctx, PSNW_EARG, "psnw_flow_majority: offsets not ascending");
    const int n_pts = tracker_offsets[n_trackers];
    if (n_pts > ctx->p.max_points) return fail(ctx, PSNW_ENOMEM, "psnw_flow_majority: max_points exceeded");
    if (n_pts > 0 && !pts_xy) return fail(ctx, PSNW_EARG, "psnw_flow_majority: null points");
    const size_t ncnt = (size_t)n_trackers * n_detections;
    int rc = ensure(ctx, ctx->arena, (size_t)n_pts * 8 + (size_t)(n_trackers + 1) * 4 + (size_t)(n_detections + n_trackers) * 33 + ncnt * 4 + n_detections + 8192);
    if (rc) return rc;
    ctx->arena_off = 0;
    float2 *d_pts = take<float2>(ctx, n_pts);
    int *d_off = take<int>(ctx, n_trackers + 1), *d_cnt = take<int>(ctx, ncnt);
    double *d_det = take<double>(ctx, (size_t)n_detections * 4), *d_trk = take<double>(ctx, (size_t)n_trackers * 4);
    uint8_t *d_ok = take<uint8_t>(ctx, n_trackers), *d_wipe = take<uint8_t>(ctx, n_detections);
    if (n_pts > 0) CK(cudaMemcpyAsync(d_pts, pts_xy, sizeof(float2) * n_pts, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off, tracker_offsets, sizeof(int) * (n_trackers + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_det, det_boxes, sizeof(double) * 4 * n_detections, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_trk, tracker_boxes, sizeof(double) * 4 * n_trackers, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_ok, tracker_ok, n_trackers, cudaMemcpyHostToDevice, ctx->stream));
    r = launch_flow_majority(d_pts, d_off, n_trackers, d_trk, d_ok, d_det, n_detections, ratio, d_cnt, d_wipe, ctx->stream);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "flow majority launch", cudaGetLastError());
    ctx->launches += r;
    CK(cudaMemcpyAsync(counts, d_cnt, sizeof(int) * ncnt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(wipe, d_wipe, n_detections, cudaMemcpyDeviceToHost, ctx->stream));
    if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return PSNW_OK;
}

int psnw_points_in_boxes(psnw_ctx *ctx, const float *pts_xy, const int *tracker_offsets, int n_trackers,
                         const double *det_boxes, int n_detections, int *counts, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_trackers == 0 || n_detections == 0) return PSNW_OK;
    if (n_trackers < 0 || n_detections < 0 || !tracker_offsets || !det_boxes || !counts)
        return fail(ctx, PSNW_EARG, "psnw_points_in_boxes arguments");
    if (n_trackers > ctx->p.max_targets || n_detections > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_points_in_boxes: max_targets exceeded");
    mem_flags &= 0xff;
    int r;
    if (mem_flags == PSNW_MEM_DEVICE) {
        r = launch_points_in_boxes(reinterpret_cast<const float2 *>(pts_xy), tracker_offsets, n_trackers, det_boxes, n_detections, counts, ctx->stream);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "points_in_boxes_kernel launch", cudaGetLastError());
        ctx->launches += r;
        return PSNW_OK;
    }
    const int n_pts = tracker_offsets[n_trackers];
    if (n_pts < 0 || n_pts > ctx->p.max_points) return fail(ctx, PSNW_ENOMEM, "psnw_points_in_boxes: max_points exceeded");
    if (n_pts > 0 && !pts_xy) return fail(ctx, PSNW_EARG, "psnw_points_in_boxes: null points");
    const size_t ncnt = (size_t)n_trackers * n_detections;
    int rc = ensure(ctx, ctx->arena, (size_t)n_pts * 8 + (size_t)(n_trackers + 1) * 4 + (size_t)n_detections * 32 + ncnt * 4 + 4096);
    if (rc) return rc;
    ctx->arena_off = 0;
    float2 *d_pts = take<float2>(ctx, n_pts);
    int *d_off = take<int>(ctx, n_trackers + 1), *d_cnt = take<int>(ctx, ncnt);
    double *d_box = take<double>(ctx, (size_t)n_detections * 4);
    if (n_pts > 0) CK(cudaMemcpyAsync(d_pts, pts_xy, sizeof(float2) * n_pts, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off, tracker_offsets, sizeof(int) * (n_trackers + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_box, det_boxes, sizeof(double) * 4 * n_detections, cudaMemcpyHostToDevice, ctx->stream));
    r = launch_points_in_boxes(d_pts, d_off, n_trackers, d_box, n_detections, d_cnt, ctx->stream);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "points_in_boxes_kernel launch", cudaGetLastError());
    ctx->launches += r;
    CK(cudaMemcpyAsync(counts, d_cnt, sizeof(int) * ncnt, cudaMemcpyDeviceToHost, ctx->stream));
    if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return PSNW_OK;
}

// GridFAST detect ("next" row f-1): the reference's own detector, Tracker2D.cpp:142 / :735-757
int psnw_gridfast_detect(psnw_ctx *ctx, int cam, int slot, const double *boxes, int n_targets, int max_per_target,
                         int threshold, int grid_rows, int grid_cols, int max_total, float *pts_xy, int *counts, int mem_flags)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (n_targets == 0) return PSNW_OK;
    if (n_targets < 0 || max_per_target < 1 || !boxes || !pts_xy || !counts || grid_rows < 1 || grid_cols < 1 ||
        grid_rows * grid_cols > 64 || max_total < 0)
        return fail(ctx, PSNW_EARG, "psnw_gridfast_detect arguments");
    if (!slot_ok(ctx, cam, slot)) return fail(ctx, PSNW_ESTATE, "psnw_gridfast_detect: camera/slot holds no frame");
    if (n_targets > ctx->p.max_targets) return fail(ctx, PSNW_ENOMEM, "psnw_gridfast_detect: max_targets exceeded");
    int rc = join_aux(ctx);
    if (rc) return rc;
    threshold = std::min(std::max(threshold, 0), 255);
    const LevelGeom &l0 = ctx->pyr.lv[0];
    const int plane = plane_of(ctx, cam, slot);
    if (!ctx->fmap[cam]) {
        if (cudaMalloc(&ctx->fmap[cam], (size_t)l0.plane_px) != cudaSuccess) return fail(ctx, PSNW_ENOMEM, "cudaMalloc FAST map", cudaGetLastError());
        ctx->fmap_tag[cam] = 0;
    }
    if (ctx->fmap_cfg[0] != threshold || ctx->fmap_cfg[1] != grid_rows || ctx->fmap_cfg[2] != grid_cols) {
        std::fill(ctx->fmap_tag.begin(), ctx->fmap_tag.end(), 0);
        ctx->fmap_cfg[0] = threshold; ctx->fmap_cfg[1] = grid_rows; ctx->fmap_cfg[2] = grid_cols;
    }
    if (ctx->fmap_tag[cam] != ctx->plane_seq[plane]) {
        const int r = launch_fast_map(l0, plane, threshold, grid_rows, grid_cols, ctx->fmap[cam], ctx->stream);
        if (r < 0) return fail(ctx, PSNW_ECUDA, "fast_map_kernel launch", cudaGetLastError());
        ctx->launches += r;
        ctx->fmap_tag[cam] = ctx->plane_seq[plane];
        rc = note_readers(ctx, ctx->stream, &plane, 1);
        if (rc) return rc;
    }
    const bool dev = mem_flags == PSNW_MEM_DEVICE;
    const size_t tb = (sizeof(GridFastTarget) * n_targets + 255) & ~(size_t)255;
    const size_t out_b = dev ? 0 : (((size_t)n_targets * max_per_target * sizeof(float2) + 255) & ~(size_t)255);
    const size_t cnt_b = dev ? 0 : ((sizeof(int) * n_targets + 255) & ~(size_t)255);
    rc = ensure(ctx, ctx->det_desc, tb + out_b + cnt_b);
    if (rc) return rc;
    float2 *s_out = reinterpret_cast<float2 *>(ctx->det_desc.p + tb);
    int *s_cnt = reinterpret_cast<int *>(ctx->det_desc.p + tb + out_b);
    std::vector<GridFastTarget> T(n_targets);
    for (int t = 0; t < n_targets; ++t) {                         // mask ROI = box.cropWithSize(cols, rows).cv(), :735
        const double x = boxes[4 * t], y = boxes[4 * t + 1], w = boxes[4 * t + 2], h = boxes[4 * t + 3];
        const double nx = std::max(0.0, x), ny = std::max(0.0, y);
        const double nw = std::min((double)l0.w - nx - 1, w), nh = std::min((double)l0.h - ny - 1, h);
        GridFastTarget &d = T[t];
        d.x0 = std::min((int)nx, l0.w); d.y0 = std::min((int)ny, l0.h);
        d.x1 = std::min(std::max((int)nx + std::max((int)nw, 0), d.x0), l0.w);
        d.y1 = std::min(std::max((int)ny + std::max((int)nh, 0), d.y0), l0.h);
        d.map = ctx->fmap[cam];
        d.out = dev ? reinterpret_cast<float2 *>(pts_xy) + (size_t)t * max_per_target : s_out + (size_t)t * max_per_target;
        d.count = dev ? counts + t : s_cnt + t;
    }
    StageSlot *s;
    rc = stage_acquire(ctx, tb, &s);
    if (rc) return rc;
    memcpy(s->h, T.data(), sizeof(GridFastTarget) * n_targets);
    rc = stage_upload(ctx, s, ctx->det_desc.p, sizeof(GridFastTarget) * n_targets);
    if (rc) return rc;
    const int r = launch_gridfast_select(l0.w, l0.h, l0.pitch, reinterpret_cast<const GridFastTarget *>(ctx->det_desc.p), n_targets,
                                         max_per_target, grid_rows, grid_cols, max_total / (grid_rows * grid_cols), ctx->stream);
    if (r < 0) return fail(ctx, PSNW_ECUDA, "gridfast_select_kernel launch", cudaGetLastError());
    ctx->launches += r;
    if (!dev) {
        CK(cudaMemcpyAsync(pts_xy, s_out, (size_t)n_targets * max_per_target * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(counts, s_cnt, sizeof(int) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        if (mem_flags == PSNW_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    }
    return PSNW_OK;
}

int psnw_read_fast_map(psnw_ctx *ctx, int cam, uint8_t *map_out)
{
    if (!ctx || !map_out || cam < 0 || cam >= ctx->p.n_cameras) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (!ctx->fmap[cam] || !ctx->fmap_tag[cam]) return fail(ctx, PSNW_ESTATE, "psnw_read_fast_map: no map computed yet");
    const LevelGeom &l0 = ctx->pyr.lv[0];
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy2D(map_out, l0.w, ctx->fmap[cam], l0.pitch, l0.w, l0.h, cudaMemcpyDeviceToHost));
    return PSNW_OK;
}

int psnw_join(psnw_ctx *ctx)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    return join_aux(ctx);
}

int psnw_sync(psnw_ctx *ctx)
{
    if (!ctx) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (int rc = join_aux(ctx)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return PSNW_OK;
}

const char *psnw_last_error(psnw_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void *psnw_get_stream(psnw_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int psnw_level_info(psnw_ctx *ctx, int level, int *w, int *h, int *pitch)
{
    if (!ctx || level < 0 || level >= ctx->pyr.levels) return PSNW_EARG;
    if (w) *w = ctx->pyr.lv[level].w;
    if (h) *h = ctx->pyr.lv[level].h;
    if (pitch) *pitch = ctx->pyr.lv[level].pitch;
    return PSNW_OK;
}

int psnw_levels_built(psnw_ctx *ctx) { return ctx ? ctx->pyr.levels : PSNW_EARG; }

int psnw_read_level(psnw_ctx *ctx, int cam, int slot, int level, uint8_t *img_out, int16_t *dxy_out)
{
    if (!ctx || level < 0 || level >= ctx->pyr.levels) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (!slot_ok(ctx, cam, slot)) return fail(ctx, PSNW_ESTATE, "psnw_read_level: camera/slot holds no frame");
    const LevelGeom &g = ctx->pyr.lv[level];
    const int plane = plane_of(ctx, cam, slot);
    if (int rc_up = join_up(ctx)) return rc_up;
    CK(cudaStreamSynchronize(ctx->stream));
    if (img_out) CK(cudaMemcpy2D(img_out, g.w, g.img + (size_t)plane * g.iplane_px, g.ipitch, g.w, g.h, cudaMemcpyDeviceToHost));
    if (dxy_out) CK(cudaMemcpy2D(dxy_out, (size_t)g.w * 4, g.dxy + (size_t)plane * g.plane_px, (size_t)g.pitch * 4, (size_t)g.w * 4, g.h, cudaMemcpyDeviceToHost));
    return PSNW_OK;
}

int psnw_read_level_apron(psnw_ctx *ctx, int cam, int slot, int level, uint8_t *out, int *apron)
{
    if (!ctx || level < 0 || level >= ctx->pyr.levels) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (!slot_ok(ctx, cam, slot)) return fail(ctx, PSNW_ESTATE, "psnw_read_level_apron: camera/slot holds no frame");
    const LevelGeom &g = ctx->pyr.lv[level];
    const int a = g.apron;
    if (apron) *apron = a;
    if (!out) return PSNW_OK;
    if (int rc_up = join_up(ctx)) return rc_up;
    CK(cudaStreamSynchronize(ctx->stream));
    const uint8_t *src = g.img + (size_t)plane_of(ctx, cam, slot) * g.iplane_px - (ptrdiff_t)a * g.ipitch - a;
    CK(cudaMemcpy2D(out, g.w + 2 * a, src, g.ipitch, g.w + 2 * a, g.h + 2 * a, cudaMemcpyDeviceToHost));
    return PSNW_OK;
}

int psnw_read_min_eig(psnw_ctx *ctx, int cam, int slot, float *eig_out)
{
    if (!ctx || !eig_out) return PSNW_EARG;
    if (int rc_enter = enter(ctx)) return rc_enter;
    if (!slot_ok(ctx, cam, slot)) return fail(ctx, PSNW_ESTATE, "psnw_read_min_eig: camera/slot holds no frame");
    const LevelGeom &l0 = ctx->pyr.lv[0];
    int rc = join_aux(ctx);
    if (rc) return rc;
    std::vector<int> buf;
    rc = ensure_min_eig(ctx, {{cam, slot}}, ctx->stream, &buf);
    if (rc) return rc;
    float *eig = ctx->eig[buf[0]];
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy2D(eig_out, (size_t)l0.w * 4, eig, (size_t)l0.pitch * 4, (size_t)l0.w * 4, l0.h, cudaMemcpyDeviceToHost));
    return PSNW_OK;
}

uint64_t psnw_launch_count(psnw_ctx *ctx) { return ctx ? ctx->launches : 0; }

uint64_t psnw_push_frame_bytes(psnw_ctx *ctx)
{
    if (!ctx) return 0;
    uint64_t b = 0;
    for (int l = 0; l < ctx->pyr.levels; ++l) {
        const uint64_t n = (uint64_t)ctx->pyr.lv[l].w * ctx->pyr.lv[l].h;
        b += n + 4 * n + (l > 0 ? n : 0);      // read u8, write short2, write u8 (levels >= 1)
    }
    return b;
}

}  // extern "C"
