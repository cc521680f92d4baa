// Falsifiable experiment: can a STAGED sampler - the source box of every output tile brought into
// shared memory by ONE tensor-map TMA instruction (cp.async.bulk.tensor.2d, UTMALDG in SASS), taps gathered with LDS - pass
// the texture kernel (57 us per 12-view README image overlapped, 67 us serialised)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tma_box_probe tools/tma_box_probe.cu && /tmp/tma_box_probe
//
// Workload = BASELINE configs[1]: 8192 x 4096 panorama (RGBA-packed uint32, linear, 8224-texel rows as in the library's
// slot), FOV 120, 12 views 1920 x 1080 (yaw 0 / 90 / 180 / 270 x pitch 30 / 60 / 90).  The output is cut into 32 x 8 pixel
// tiles; the host computes every tile's source bounding box with the projection's formulas (double precision - this probe
// measures data movement, not parity) and the per-pixel tap origin inside it.  A tensor map has ONE box size, so tiles are
// served by the smallest of a few box classes that covers them; tiles no class covers (around the poles) are counted and
// skipped (the real design would send them to the texture path).  Box origins sit on 16-byte boundaries: a
// cp.async.bulk.tensor whose inner coordinate is not a multiple of 4 texels faults with "illegal instruction" on this part.
//
// Modes:  copy    one TMA box per tile, wait on the mbarrier, nothing else (one word read so the copy cannot be elided)
//         gather  copy + 4 LDS taps per pixel from a precomputed 4-byte tap descriptor (no coordinate arithmetic at all),
//                 the exact integer blend and the packed output stores
// Prints one JSON line per mode: microseconds per 12-view image (CUDA events, 8 rotating panoramas = 1.1 GB > L2), bytes
// moved by the boxes, and the share of tiles covered.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int WP = 8192, HP = 4096, PITCH_TEX = 8224, W = 1920, H = 1080, FOV = 120;
constexpr int TW = 32, TH = 8;                       // output tile
constexpr int N_CLASS = 3;
static const int kBoxW[N_CLASS] = {64, 96, 128};     // texels (x 4 bytes: multiples of 16 bytes)
static const int kBoxH[N_CLASS] = {24, 40, 64};      // a 32 x 8 tile needs 45 x 23 texels at the median, 86 x 42 at p90

struct Tile {
    int32_t x0, y0;       // box origin in the panorama (x0 already rolled by the yaw, may wrap: see below)
    int32_t view, tx, ty; // output tile
    int32_t cls;          // box class, -1 = not covered
};

// per-pixel tap descriptor: offset of the top-left tap inside the box (x | y << 8) and the 5-bit fractions
struct __align__(4) Tap {
    uint8_t x, y, fx, fy;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t blend4(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t fx, uint32_t fy) {
    const uint32_t t = (32u - fx) | (fx << 16);
    const uint32_t wA = t * (32u - fy), wB = t * fy;
    const uint32_t t0 = __byte_perm(p00, p01, 0x5140), t1 = __byte_perm(p10, p11, 0x5140);
    const uint32_t t2 = __byte_perm(p00, p01, 0x6262), t3 = __byte_perm(p10, p11, 0x6262);
    const uint32_t sb = __dp2a_lo(wB, t1, __dp2a_lo(wA, t0, 512u));
    const uint32_t sg = __dp2a_hi(wB, t1, __dp2a_hi(wA, t0, 512u));
    const uint32_t sr = __dp2a_lo(wB, t3, __dp2a_lo(wA, t2, 512u));
    return (sb >> 10) | ((sg >> 2) & 0xFF00u) | ((sr << 6) & 0xFF0000u);
}

// one CTA = one tile (32 x 8 threads).  BW x BH box, dynamic shared memory: box (128-byte aligned) + mbarrier
template <bool GATHER>
__global__ void __launch_bounds__(TW *TH) probe_kernel(const __grid_constant__ CUtensorMap map, const Tile *__restrict__ tiles,
                                                      const Tap *__restrict__ taps, uint8_t *__restrict__ out, int BW, int BH,
                                                      unsigned *__restrict__ sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *box = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)BW * BH * 4);
    const Tile t = tiles[blockIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(BW * BH * 4));
        tma_load_2d(box, &map, t.x0, t.y0, bar);
    }
    mbar_wait(bar, 0);
    if (!GATHER) {
        if (threadIdx.x == 0 && box[0] == 0xDEADBEEFu) atomicAdd(sink, 1u);   // keeps the copy alive
        return;
    }
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int u = t.tx * TW + lx, v = t.ty * TH + ly;
    if (u >= W || v >= H) return;
    const size_t px = ((size_t)t.view * H + v) * W + u;
    const Tap tp = taps[px];
    const uint32_t *r0 = box + (size_t)tp.y * BW + tp.x;
    const uint32_t q = blend4(r0[0], r0[1], r0[BW], r0[BW + 1], tp.fx, tp.fy);
    // packed stores like the library's store_quad: 4 pixels -> 3 words
    const uint32_t nxt = __shfl_down_sync(0xffffffffu, q, 1);
    const int j = lx & 3;
    if (j < 3) {
        const uint32_t word = __funnelshift_r(q << 8, nxt, 8 * (j + 1));
        __stcs(reinterpret_cast<uint32_t *>(out + px * 3 - 3 * j + 4 * j), word);
    }
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    // ---- geometry on the host (double precision; the probe measures data movement) ----
    const int yaws[4] = {0, 90, 180, 270}, pitches[3] = {30, 60, 90};
    const double f = 0.5 * W / tan(FOV * M_PI / 360.0);
    const int ntx = (W + TW - 1) / TW, nty = (H + TH - 1) / TH;
    std::vector<Tile> tiles;
    std::vector<Tap> taps((size_t)12 * W * H);
    std::vector<int> need_w, need_h;
    long long uncovered = 0;
    for (int pi = 0; pi < 3; ++pi) {
        const double p = pitches[pi] * M_PI / 180.0, c = cos(p), s = sin(p);
        std::vector<float> U((size_t)W * H), V((size_t)W * H);
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                const double x = u - W / 2.0, y = H / 2.0 - v, n = sqrt(x * x + y * y + f * f);
                const double xn = x / n, yn = y / n, zn = f / n;
                const double yr = c * yn - s * zn, zr = s * yn + c * zn;
                double phi = atan2(yr, xn);
                if (phi < 0) phi += 2 * M_PI;
                U[(size_t)v * W + u] = (float)std::min(std::max(phi * WP / (2 * M_PI), 0.0), (double)(WP - 1));
                V[(size_t)v * W + u] = (float)std::min(std::max(acos(std::min(1.0, std::max(-1.0, zr))) * HP / M_PI, 0.0), (double)(HP - 1));
            }
        for (int yi = 0; yi < 4; ++yi) {
            const int view = yi * 3 + pi, shift = yaws[yi] * WP / 360;
            for (int ty = 0; ty < nty; ++ty)
                for (int tx = 0; tx < ntx; ++tx) {
                    int x0 = 1 << 30, x1 = -1, y0 = 1 << 30, y1 = -1;
                    for (int v = ty * TH; v < std::min(H, ty * TH + TH); ++v)
                        for (int u = tx * TW; u < std::min(W, tx * TW + TW); ++u) {
                            const int sx = (int)lrintf(U[(size_t)v * W + u] * 32.f), sy = (int)lrintf(V[(size_t)v * W + u] * 32.f);
                            x0 = std::min(x0, sx >> 5); x1 = std::max(x1, (sx >> 5) + 1);
                            y0 = std::min(y0, sy >> 5); y1 = std::max(y1, (sy >> 5) + 1);
                        }
                    x0 &= ~3;   // box origins on 16-byte boundaries (the widths below are measured from the aligned origin)
                    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
                    need_w.push_back(bw);
                    need_h.push_back(bh);
                    Tile t;
                    t.view = view; t.tx = tx; t.ty = ty; t.cls = -1;
                    for (int k = 0; k < N_CLASS && t.cls < 0; ++k)
                        if (bw <= kBoxW[k] && bh <= kBoxH[k]) t.cls = k;
                    // the yaw roll moves the box; a box that would straddle the seam is not split here (the real design
                    // appends wrap columns): for the timing it is clamped inside the row
                    t.x0 = std::min((x0 + shift) % WP, PITCH_TEX - (t.cls >= 0 ? kBoxW[t.cls] : 0)) & ~3;
                    t.y0 = y0;
                    if (t.cls < 0) { ++uncovered; continue; }
                    for (int v = ty * TH; v < std::min(H, ty * TH + TH); ++v)
                        for (int u = tx * TW; u < std::min(W, tx * TW + TW); ++u) {
                            const int sx = (int)lrintf(U[(size_t)v * W + u] * 32.f), sy = (int)lrintf(V[(size_t)v * W + u] * 32.f);
                            Tap &tp = taps[((size_t)view * H + v) * W + u];
                            tp.x = (uint8_t)((sx >> 5) - x0); tp.y = (uint8_t)((sy >> 5) - y0);
                            tp.fx = (uint8_t)(sx & 31); tp.fy = (uint8_t)(sy & 31);
                        }
                    tiles.push_back(t);
                }
        }
    }
    std::vector<int> sw = need_w, sh = need_h;
    std::sort(sw.begin(), sw.end());
    std::sort(sh.begin(), sh.end());
    const size_t nt = need_w.size();
    fprintf(stderr, "tiles %zu, needed box width median %d p90 %d p99 %d max %d; height median %d p90 %d p99 %d max %d; uncovered %lld\n",
            nt, sw[nt / 2], sw[nt * 9 / 10], sw[nt * 99 / 100], sw.back(), sh[nt / 2], sh[nt * 9 / 10], sh[nt * 99 / 100], sh.back(), uncovered);

    // ---- device data ----
    const int n_pano = 8;
    const size_t pano_bytes = (size_t)PITCH_TEX * (HP + 32) * 4;
    std::vector<uint32_t *> d_pano(n_pano);
    {
        std::vector<uint32_t> h(pano_bytes / 4);
        for (int i = 0; i < n_pano; ++i) {
            uint32_t x = 1234567u + i;
            for (size_t k = 0; k < h.size(); ++k) { x = x * 1664525u + 1013904223u; h[k] = x & 0x00FFFFFFu; }
            CK(cudaMalloc(&d_pano[i], pano_bytes));
            CK(cudaMemcpy(d_pano[i], h.data(), pano_bytes, cudaMemcpyHostToDevice));
        }
    }
    encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
    Tap *d_taps; uint8_t *d_out; unsigned *d_sink;
    CK(cudaMalloc(&d_taps, taps.size() * sizeof(Tap)));
    CK(cudaMemcpy(d_taps, taps.data(), taps.size() * sizeof(Tap), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)12 * W * H * 3 + 64));
    CK(cudaMalloc(&d_sink, 4));
    CK(cudaMemset(d_sink, 0, 4));
    // tiles by class
    std::vector<std::vector<Tile>> by_cls(N_CLASS);
    for (const Tile &t : tiles) by_cls[t.cls].push_back(t);
    std::vector<Tile *> d_tiles(N_CLASS, nullptr);
    double box_bytes = 0;
    for (int k = 0; k < N_CLASS; ++k) {
        if (by_cls[k].empty()) continue;
        CK(cudaMalloc(&d_tiles[k], by_cls[k].size() * sizeof(Tile)));
        CK(cudaMemcpy(d_tiles[k], by_cls[k].data(), by_cls[k].size() * sizeof(Tile), cudaMemcpyHostToDevice));
        box_bytes += (double)by_cls[k].size() * kBoxW[k] * kBoxH[k] * 4;
    }
    for (int gather = 0; gather < 2; ++gather) {
        std::vector<CUtensorMap> maps((size_t)n_pano * N_CLASS);
        for (int i = 0; i < n_pano; ++i)
            for (int k = 0; k < N_CLASS; ++k) {
                const cuuint64_t dims[2] = {(cuuint64_t)PITCH_TEX, (cuuint64_t)(HP + 32)};
                const cuuint64_t strides[1] = {(cuuint64_t)PITCH_TEX * 4};
                const cuuint32_t box[2] = {(cuuint32_t)kBoxW[k], (cuuint32_t)kBoxH[k]};
                const cuuint32_t estr[2] = {1, 1};
                CUresult r = encode(&maps[(size_t)i * N_CLASS + k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_pano[i], dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
            }
        auto launch_image = [&](int img) {
            for (int k = 0; k < N_CLASS; ++k) {
                if (by_cls[k].empty()) continue;
                const size_t smem = (size_t)kBoxW[k] * kBoxH[k] * 4 + 16;
                if (gather) {
                    CK(cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    probe_kernel<true><<<(unsigned)by_cls[k].size(), TW * TH, smem>>>(maps[(size_t)img * N_CLASS + k], d_tiles[k], d_taps,
                                                                                      d_out, kBoxW[k], kBoxH[k], d_sink);
                } else {
                    CK(cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    probe_kernel<false><<<(unsigned)by_cls[k].size(), TW * TH, smem>>>(maps[(size_t)img * N_CLASS + k], d_tiles[k], d_taps,
                                                                                       d_out, kBoxW[k], kBoxH[k], d_sink);
                }
            }
        };
        launch_image(0);
        CK(cudaDeviceSynchronize());
        for (int i = 0; i < n_pano; ++i) launch_image(i);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        const int reps = 5;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r)
            for (int i = 0; i < n_pano; ++i) launch_image(i);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / (reps * n_pano);
        printf("{\"mode\": \"%s\", \"image_us\": %.2f, \"tiles\": %zu, \"tiles_covered\": %zu, \"covered_frac\": %.4f, "
               "\"box_MB_per_image\": %.1f, \"box_classes\": \"64x24 / 96x40 / 128x64 texels\", \"tiles_per_class\": [%zu, %zu, %zu], "
               "\"box_GBs\": %.0f, \"cuda_error\": \"%s\"}\n",
               gather ? "copy + LDS gather + blend + stores (precomputed taps, no coordinate math)" : "copy only (one TMA box per tile)",
               us, nt, tiles.size(), (double)tiles.size() / nt, box_bytes / 1e6, by_cls[0].size(), by_cls[1].size(), by_cls[2].size(),
               box_bytes / (us * 1e-6) / 1e9, cudaGetErrorString(cudaGetLastError()));
        fflush(stdout);
    }
    return 0;
}
