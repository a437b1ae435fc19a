// Bring-up aid: per-CTA timeline of the fused pyramid kernel (pyramid_fused.cu compiled with time stamps at its phase
// boundaries).  usage: pyr_timeline [frames] [variant] [wide] [w h]      -- prints the median / max of every phase in ns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I include -o /tmp/pyr_timeline tools/pyr_timeline.cu -lcuda
#define PSNW_PYR_TIMELINE
#include "../mcmtt_opticalflow_b200/csrc/pyramid_fused.cu"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace psnw;

__global__ void empty_kernel() {}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 4, variant = argc > 2 ? atoi(argv[2]) : 0, wide = argc > 3 ? atoi(argv[3]) : 0;
    const int W = argc > 5 ? atoi(argv[4]) : 768, H = argc > 5 ? atoi(argv[5]) : 576, levels = 4, apron = 26;
    PyrDesc pyr = {};
    pyr.levels = levels;
    int w = W, h = H;
    for (int l = 0; l < levels; ++l) {
        if (l) { w = (w + 1) / 2; h = (h + 1) / 2; }
        LevelGeom &g = pyr.lv[l];
        g.w = w; g.h = h; g.pitch = (w + 127) & ~127; g.plane_px = g.pitch * h;
        g.apron = l ? apron : 0; g.apron_l = g.apron ? apron_left(g.apron) : 0;
        g.ipitch = g.apron ? (g.apron_l + w + apron_right(g.apron) + 63) & ~63 : g.pitch;
        g.iplane_px = g.apron ? g.ipitch * (h + 2 * g.apron) : g.plane_px;
        uint8_t *img;
        CK(cudaMalloc(&img, (size_t)g.iplane_px * frames));
        CK(cudaMalloc(&g.dxy, (size_t)g.plane_px * frames * 4));
        CK(cudaMemset(img, 0x55, (size_t)g.iplane_px * frames));
        g.img = img + (g.apron ? (size_t)g.apron * g.ipitch + g.apron_l : 0);
    }
    int bw, bh, ow, oh;
    fused_geometry(levels, variant, &bw, &bh, &ow, &oh);
    CUtensorMap map;
    {
        const LevelGeom &g = pyr.lv[0];
        const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.iplane_px};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, estr[3] = {1, 1, 1};
        cudaFree(0);
        if (cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g.img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("tensor map failed\n");
            return 1;
        }
    }
    if (configure_fused_kernels()) { printf("configure failed\n"); return 1; }
    PlaneList pl;
    pl.n = frames;
    for (int i = 0; i < frames; ++i) pl.plane[i] = i;
    const int ctas = ((W + ow - 1) / ow) * ((H + oh - 1) / oh) * frames;
    void *flush;
    CK(cudaMalloc(&flush, 256 << 20));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<unsigned long long> tl(8 * (size_t)ctas);
    for (int rep = 0; rep < 6; ++rep) {
        if (rep < 4) CK(cudaMemsetAsync(flush, rep, 256 << 20, 0));          // reps 4, 5: warm caches (no flush)
        CK(cudaEventRecord(e0, 0));
        if (launch_pyr_fused(pyr, levels, variant, wide, &map, pl, 0) < 0) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        CK(cudaEventRecord(e1, 0));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpyFromSymbol(tl.data(), g_timeline, sizeof(unsigned long long) * tl.size()));
        unsigned long long first = ~0ull, last = 0;
        for (int c = 0; c < ctas; ++c) { first = std::min(first, tl[8 * c]); last = std::max(last, tl[8 * c + 5]); }
        printf("rep %d: event %.2f us, first CTA start -> last CTA end %.2f us, %d CTAs\n", rep, ms * 1e3, (last - first) * 1e-3, ctas);
        if (rep == 3 || rep == 5) {
            const char *names[6] = {"start after first", "tma wait", "chain", "scharr", "copies", "apron"};
            for (int p = 0; p < 6; ++p) {
                std::vector<double> v;
                for (int c = 0; c < ctas; ++c) v.push_back(p == 0 ? (double)(tl[8 * c] - first) : (double)(tl[8 * c + p] - tl[8 * c + p - 1]));
                std::sort(v.begin(), v.end());
                printf("  %-18s median %7.0f ns   p10 %7.0f   max %7.0f\n", names[p], v[v.size() / 2], v[v.size() / 10], v.back());
            }
            std::vector<double> v;
            for (int c = 0; c < ctas; ++c) v.push_back((double)(tl[8 * c + 5] - tl[8 * c]));
            std::sort(v.begin(), v.end());
            printf("  %-18s median %7.0f ns   p10 %7.0f   max %7.0f\n", "whole CTA", v[v.size() / 2], v[v.size() / 10], v.back());
        }
    }
    for (int rep = 0; rep < 3; ++rep) {          // floor of the measurement: an empty launch between the same events
        CK(cudaMemsetAsync(flush, rep, 256 << 20, 0));
        CK(cudaEventRecord(e0, 0));
        empty_kernel<<<144, 512, 0, 0>>>();
        CK(cudaEventRecord(e1, 0));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("empty kernel after flush: event %.2f us\n", ms * 1e3);
    }
    return 0;
}
