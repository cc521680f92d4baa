// p2p_api.cu - C ABI (include/p2p.h) over the sm_100a kernels in p2p_kernels.cuh.
// Host logic only: contexts, panorama slots (stream + device buffers), launches, transfers.
#include "p2p.h"
#include "p2p_kernels.cuh"
#include "p2p_jpeg_host.cuh"
#include "p2p_jpegdec.cuh"
#include "p2p_png.cuh"

#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace p2p;

namespace {

// NVTX range over one entry point of the C ABI (per image and stage: upload / decode, project, encode, replicate)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define P2P_NVTX(name) NvtxRange nvtx_range_(name)

struct Slot {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t owned = nullptr;  // created by us, replaced through p2p_set_stream
    uint8_t *d_bgr = nullptr;  // staging copy of the caller's BGR rows
    size_t bgr_cap = 0;
    uint32_t *d_rgba = nullptr;  // packed panorama
    size_t rgba_cap = 0;
    int Wp = 0, Hp = 0, pitch_tex = 0;
    bool valid = false;
    int row0 = 0, row1 = 0;  // packed rows row0 .. row1 (<= Hp, the clamp row) hold data; a full upload has 0 .. Hp
    cudaArray_t arr = nullptr;  // gather-enabled array (sampler 1)
    int arrW = 0, arrH = 0;
    cudaTextureObject_t tex = 0;
    cudaSurfaceObject_t surf = 0;  // the same array, written directly by the pack kernel
    bool tex_current = false;
    uint8_t *d_out = nullptr;  // device outputs when the caller wants them on the host
    size_t out_cap = 0;
    int32_t *d_tab = nullptr;  // yaw table (ix | fx), 2 * Wp
    size_t tab_cap = 0;
    // JPEG encoder scratch (p2p_encode_jpeg): coefficients, per-block bits / offsets, bit string, stuffed files
    int16_t *j_coef = nullptr;
    size_t j_coef_cap = 0;
    uint32_t *j_bits = nullptr;   // [2][n * n_blocks]: code bits, exclusive offsets
    size_t j_bits_cap = 0;
    uint32_t *j_stream = nullptr;
    size_t j_stream_cap = 0;
    uint32_t *j_cnt = nullptr;    // [2][n * chunks]: 0xFF counts, exclusive offsets
    size_t j_cnt_cap = 0;
    uint8_t *j_out = nullptr;
    size_t j_out_cap = 0;
    unsigned long long *j_tot = nullptr;  // [3][n]: total bits, total 0xFF, n_chunks (as u32 pairs)
    size_t j_tot_cap = 0;
    // PNG encoder scratch (p2p_encode_png)
    uint8_t *pg_F = nullptr;
    size_t pg_F_cap = 0;
    uint32_t *pg_S = nullptr;
    size_t pg_S_cap = 0;
    uint16_t *pg_tlen = nullptr;
    size_t pg_tlen_cap = 0;
    uint32_t *pg_blk = nullptr;      // blockpos | blkoff | ntok | lfreq
    size_t pg_blk_cap = 0;
    p2ppng::BlockInfo *pg_info = nullptr;
    size_t pg_info_cap = 0;
    uint32_t *pg_Z = nullptr;
    size_t pg_Z_cap = 0;
    unsigned long long *pg_sums = nullptr;   // [2 n] Adler partial sums | [n] zbits
    size_t pg_sums_cap = 0;
    // JPEG decoder (p2p_upload_pano_jpeg): pinned coefficient staging, device coefficients and component planes
    int16_t *jd_coef_h = nullptr;
    size_t jd_coef_h_cap = 0;
    int16_t *jd_coef_d = nullptr;
    size_t jd_coef_d_cap = 0;
    uint8_t *jd_planes = nullptr;
    size_t jd_planes_cap = 0;
    // device Huffman stage: destuffed scan, subsequence states / block counts, tables, DC differences
    uint32_t *jd_stream = nullptr;
    size_t jd_stream_cap = 0;
    unsigned long long *jd_states = nullptr;  // [2][n_sub]: start, end
    size_t jd_states_cap = 0;
    uint32_t *jd_nblk = nullptr;              // [2][n_sub rounded]: blocks per subsequence, exclusive offsets
    size_t jd_nblk_cap = 0;
    p2pjdec::DevHuff *jd_tables = nullptr;    // [3]
    int32_t *jd_dc = nullptr;                 // [2][3][dc_stride]: DC differences, exclusive sums
    size_t jd_dc_cap = 0;
    uint32_t *jd_tiles = nullptr;             // tile sums / offsets of the three-phase DC scan
    size_t jd_tiles_cap = 0;
    unsigned long long *jd_tot_d = nullptr;   // scan totals (device) [4]
    struct JdFlags { int changed[8]; int bad; int out_of_range; unsigned long long total; } *jd_flags_h = nullptr, *jd_flags_d = nullptr;  // mapped
    uint8_t *jd_sub = nullptr;                // subsequence layout + first subsequence of every restart interval
    size_t jd_sub_cap = 0;
    unsigned long long *j_sizes_h = nullptr;  // mapped host memory: file sizes
    unsigned long long *j_sizes_d = nullptr;
    int j_sizes_n = 0;
};

// memoised tap-row range of one view geometry (no yaw, no image: the key of the reference's pitch map cache)
struct RowRange {
    bool valid = false;
    int W = 0, H = 0, Wp = 0, Hp = 0, trig = 0;
    std::vector<p2p_pitch_consts> pc;
    int lo = 0, hi = 0;  // min / max tap row iy over all pixels (the sampler reads rows iy and iy + 1)
};

}  // namespace

struct p2p_ctx {
    int device = 0;
    int n_slots = 0;
    Slot *slots = nullptr;
    std::mutex mu;
    int opt_sampler = 1;
    int opt_warp_w = 32;
    int opt_ny = 4;
    int opt_nb = 1;
    int opt_mirror = 2;        // 2: row-segment kernel (all-word stores, view groups), 1: round-1 mirror kernel, 0: none
    int opt_seg_chunks = 4;    // chunks of 32 pixel pairs a warp of the row-segment kernel walks
    int opt_trig = 0;          // 0: NumPy-exact (SVML) acos / atan2, 1: own minimax fits
    int opt_interp = 0;
    int opt_seam_wrap = 0;     // exact-bilinear mode only: interpolate across the 0 / 360 degree seam instead of clamping
    int opt_partial = 1;       // p2p_process_image transfers only the panorama rows its views can touch
    int opt_gpu_huffman = 1;   // JPEG inputs without restart markers: Huffman decoding on the device
    long long gpu_huffman_used = 0, gpu_huffman_fallback = 0;
    uint32_t *d_crc_table = nullptr;    // CRC-32 table of the PNG encoder
    p2pjpeg::Tables *d_jtab = nullptr;  // JPEG tables + header of (jW, jH, jQ): the entry of jtabs in use
    int jW = 0, jH = 0, jQ = 0;
    struct JTab { int W, H, Q; p2pjpeg::Tables *d; };
    std::vector<JTab> jtabs;            // one device copy per (size, quality) seen: a folder of mixed sizes never waits
    int *j_err_h = nullptr, *j_err_d = nullptr;  // mapped: set by the encoder kernels when a file does not fit
    std::vector<RowRange> rows;   // memoised per geometry (the reference's pitch_mapping_cache key, ref :55-73)
    int *d_range = nullptr;
    long long launches = 0;
    uint4 *d_flush = nullptr;
    size_t flush_cap = 0;
};

namespace {

// The last error is kept PER CALLING THREAD (like errno): a context is driven by many host threads at once (one per image
// in flight), several entry points fail before or after they hold the context lock, and a message shared through the
// context could be overwritten - or freed - by another thread between the failing call and p2p_last_error.
thread_local std::string tl_err;
thread_local const p2p_ctx *tl_err_ctx = nullptr;

int fail(p2p_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        tl_err = what;
        if (e != cudaSuccess) {
            tl_err += ": ";
            tl_err += cudaGetErrorString(e);
        }
        tl_err_ctx = ctx;
    }
    return code;
}

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) {                                          \
            cudaGetLastError();                                           \
            return fail(ctx, (e_ == cudaErrorMemoryAllocation) ? P2P_ERR_NOMEM : P2P_ERR_CUDA, #call, e_); \
        }                                                                 \
    } while (0)

template <typename T>
int ensure(p2p_ctx *ctx, T **ptr, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return P2P_OK;
    if (*ptr) {
        // a buffer that has to grow (a larger image than this slot has seen) grows with 50 % headroom, and never
        // shrinks: cudaFree / cudaMalloc synchronise the whole device, so a folder of mixed sizes must not pay them per image
        bytes += bytes / 2;
        CK(cudaFree(*ptr));
        *ptr = nullptr;
        *cap = 0;
    }
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    *ptr = static_cast<T *>(p);
    *cap = bytes;
    return P2P_OK;
}

// for buffers whose size follows the CONTENT of a file (compressed scan length ...): grow with 50 % headroom, so that a
// folder of similar files does not free / allocate (= synchronise the device) on every slightly larger one
template <typename T>
int ensure_grow(p2p_ctx *ctx, T **ptr, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return P2P_OK;
    return ensure(ctx, ptr, cap, *ptr ? bytes : bytes + bytes / 2);  // ensure() adds the headroom itself when it regrows
}

int slot_ok(p2p_ctx *ctx, int slot) { return ctx && slot >= 0 && slot < ctx->n_slots; }

int check_dims(p2p_ctx *ctx, int Wp, int Hp) {
    if (Wp <= 0 || Hp <= 0) return fail(ctx, P2P_ERR_INVALID, "panorama size must be positive");
    // cv::remap asserts every dimension < SHRT_MAX (SURVEY 8b "limits inherited")
    if (Wp >= 32767 || Hp >= 32767) return fail(ctx, P2P_ERR_LIMIT, "panorama dimension >= 32767");
    return P2P_OK;
}

int prepare_slot(p2p_ctx *ctx, Slot &s, int Wp, int Hp) {
    const int pitch_tex = ((Wp + 1) + 31) & ~31;  // 128-byte aligned rows
    const size_t bytes = (size_t)pitch_tex * (size_t)(Hp + 1) * 4;
    int rc = ensure(ctx, &s.d_rgba, &s.rgba_cap, bytes);
    if (rc) return rc;
    s.Wp = Wp;
    s.Hp = Hp;
    s.pitch_tex = pitch_tex;
    s.tex_current = false;
    return P2P_OK;
}

int ensure_array(p2p_ctx *ctx, Slot &s);

// pack panorama rows y0 .. y1 (y1 <= Hp: row Hp is the clamp row) of the staging image into the device layout
int launch_pack(p2p_ctx *ctx, Slot &s, const uint8_t *d_src, size_t stride, int y0, int y1) {
    const int groups = s.Wp / 4 + 1;
    dim3 block(256), grid((groups + 255) / 256, y1 - y0 + 1);
    const int aligned4 = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 3) == 0);
    // with the texture sampler the pack kernel also writes the gather array through a surface,
    // so no device-to-device copy is needed before the projection
    cudaSurfaceObject_t surf = 0;
    if (ctx->opt_sampler == 1) {
        int rc = ensure_array(ctx, s);
        if (rc) return rc;
        surf = s.surf;
    }
    pack_kernel<<<grid, block, 0, s.stream>>>(d_src, stride, s.d_rgba, s.pitch_tex, s.Wp, s.Hp, aligned4, surf, y0);
    ctx->launches++;
    CK(cudaGetLastError());
    s.valid = true;
    s.row0 = y0;
    s.row1 = y1;
    s.tex_current = (surf != 0);
    return P2P_OK;
}

bool slot_is_partial(const Slot &s) { return s.row0 > 0 || s.row1 < s.Hp; }

// Tap-row range of a view set on a Wp x Hp panorama: the sampler reads rows lo .. hi + 1.  Evaluated once per
// geometry on `st` (one small kernel + an 8-byte readback) and memoised in the context.
int view_row_range(p2p_ctx *ctx, cudaStream_t st, int n_pitch, const p2p_pitch_consts *pitch, int W, int H,
                   int Wp, int Hp, int *lo, int *hi) {
    RowRange *found = nullptr;
    for (RowRange &c : ctx->rows) {
        bool hit = c.valid && c.W == W && c.H == H && c.Wp == Wp && c.Hp == Hp && c.trig == ctx->opt_trig &&
                   (int)c.pc.size() == n_pitch;
        for (int j = 0; hit && j < n_pitch; ++j) hit = memcmp(&c.pc[j], &pitch[j], sizeof(p2p_pitch_consts)) == 0;
        if (hit) found = &c;
    }
    const bool hit = found != nullptr;
    if (!hit) {
        if (ctx->rows.size() >= 64) ctx->rows.erase(ctx->rows.begin());  // oldest geometry out
        ctx->rows.emplace_back();
        found = &ctx->rows.back();
    }
    RowRange &r = *found;
    if (!hit) {
        if ((H + 7) / 8 > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
        if (!ctx->d_range) CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_range), 2 * sizeof(int)));
        const int init[2] = {INT_MAX, INT_MIN};
        CK(cudaMemcpyAsync(ctx->d_range, init, sizeof(init), cudaMemcpyHostToDevice, st));
        RowRangeParams P;
        memset(&P, 0, sizeof(P));
        P.W = W;
        P.H = H;
        P.halfW = (float)(W / 2.0);
        P.halfH = (float)(H / 2.0);
        P.Hp_f = (float)Hp;
        P.Vmax = (float)(Hp - 1);
        P.numpy_trig = (ctx->opt_trig == 0);
        for (int p0 = 0; p0 < n_pitch; p0 += kMaxPitchPerLaunch) {
            const int np_l = (n_pitch - p0 < kMaxPitchPerLaunch) ? n_pitch - p0 : kMaxPitchPerLaunch;
            for (int j = 0; j < np_l; ++j) P.pc[j] = PitchC{pitch[p0 + j].f, pitch[p0 + j].c, pitch[p0 + j].s};
            tap_rows_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, np_l), 256, 0, st>>>(P, ctx->d_range);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        int got[2] = {0, 0};
        CK(cudaMemcpyAsync(got, ctx->d_range, sizeof(got), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (got[0] > got[1]) got[0] = got[1] = 0;  // every pixel NaN: nothing is read
        r.W = W; r.H = H; r.Wp = Wp; r.Hp = Hp; r.trig = ctx->opt_trig;
        r.pc.assign(pitch, pitch + n_pitch);
        r.lo = got[0];
        r.hi = got[1];
        r.valid = true;
    }
    *lo = r.lo;
    *hi = r.hi;
    return P2P_OK;
}

// gather-enabled array of a slot (texture for the sampler, surface for the pack kernel)
int ensure_array(p2p_ctx *ctx, Slot &s) {
    const int aw = s.Wp, ah = s.Hp;  // wrap in x / clamp in y replace the duplicated column and row
    if (!s.arr || s.arrW != aw || s.arrH != ah) {
        if (s.tex) {
            CK(cudaDestroyTextureObject(s.tex));
            s.tex = 0;
        }
        if (s.surf) {
            CK(cudaDestroySurfaceObject(s.surf));
            s.surf = 0;
        }
        if (s.arr) {
            CK(cudaFreeArray(s.arr));
            s.arr = nullptr;
        }
        cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindUnsigned);
        CK(cudaMallocArray(&s.arr, &fd, aw, ah, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
        s.arrW = aw;
        s.arrH = ah;
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = s.arr;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeWrap;
        td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 1;
        CK(cudaCreateTextureObject(&s.tex, &rd, &td, nullptr));
        CK(cudaCreateSurfaceObject(&s.surf, &rd));
    }
    return P2P_OK;
}

// make the texture of a slot current: normally the pack kernel has written the array already; a
// panorama produced by the rotate kernel (linear buffer only) is copied into it
int ensure_texture(p2p_ctx *ctx, Slot &s) {
    if (s.tex_current) return P2P_OK;
    int rc = ensure_array(ctx, s);
    if (rc) return rc;
    CK(cudaMemcpy2DToArrayAsync(s.arr, 0, 0, s.d_rgba, (size_t)s.pitch_tex * 4, (size_t)s.Wp * 4, s.Hp,
                                cudaMemcpyDeviceToDevice, s.stream));
    s.tex_current = true;
    return P2P_OK;
}

typedef void (*proj_fn)(const ProjParams);

template <int WARP_W, int NB, int SAMPLER, bool QUAD>
proj_fn pick_ny(int ny) {
    switch (ny) {
        case 1: return project_kernel<WARP_W, 1, NB, SAMPLER, QUAD>;
        case 2: return project_kernel<WARP_W, 2, NB, SAMPLER, QUAD>;
        case 3: return project_kernel<WARP_W, 3, NB, SAMPLER, QUAD>;
        default: return project_kernel<WARP_W, 4, NB, SAMPLER, QUAD>;
    }
}

template <int NB, int SAMPLER>
proj_fn pick_w(int warp_w, int ny) {
    return (warp_w == 8) ? pick_ny<8, NB, SAMPLER, true>(ny) : pick_ny<32, NB, SAMPLER, true>(ny);
}

// the packed-store variants; outputs with W % 4 != 0 (or unaligned) use one generic byte-store kernel
template <int SAMPLER>
proj_fn pick_kernel(bool quad, int nb, int warp_w, int ny) {
    if (!quad) return pick_ny<32, 1, SAMPLER, false>(ny);
    switch (nb) {
        case 4: return pick_w<4, SAMPLER>(warp_w, ny);
        case 2: return pick_w<2, SAMPLER>(warp_w, ny);
        default: return pick_w<1, SAMPLER>(warp_w, ny);
    }
}

// ---- row-segment kernel: any flat list of views in one launch ----------------------------------
typedef void (*rows_fn)(const RowsParams);

template <bool TRIG, bool FULL>
rows_fn pick_rows_ny(int ny) {
    switch (ny) {
        case 1: return project_rows_kernel<1, TRIG, FULL>;
        case 2: return project_rows_kernel<2, TRIG, FULL>;
        case 3: return project_rows_kernel<3, TRIG, FULL>;
        default: return project_rows_kernel<4, TRIG, FULL>;
    }
}

rows_fn pick_rows(bool numpy_trig, bool full, int ny) {
    if (numpy_trig) return full ? pick_rows_ny<true, true>(ny) : pick_rows_ny<true, false>(ny);
    return full ? pick_rows_ny<false, true>(ny) : pick_rows_ny<false, false>(ny);
}

bool rows_kernel_usable(const p2p_ctx *ctx, int W, int H, int n_views, const void *d_out) {
    return ctx->opt_mirror == 2 && ctx->opt_sampler == 1 && ctx->opt_interp == 0 && (W & 7) == 0 &&
           (reinterpret_cast<uintptr_t>(d_out) & 3) == 0 &&
           (unsigned long long)W * H * 3 * (unsigned long long)n_views < (1ull << 32);  // 32-bit byte offsets in the kernel
}

// views[i] = (yaw roll, pitch constants) -> d_out + out_index[i] * W * H * 3.  Views with bit-identical pitch constants
// share one coordinate evaluation (up to 4 per group: the key of the reference's pitch_mapping_cache, ref :55-73, has no
// yaw in it); the groups of the whole list go out in one launch (kMaxViewGroups per launch).
int launch_rows(p2p_ctx *ctx, Slot &s, int n_views, const int32_t *yaw_shift, const p2p_pitch_consts *pitch,
                const int *out_index, int W, int H, uint8_t *d_out, int row_begin = 0, int row_end = -1) {
    if (row_end < 0) row_end = H;
    if (row_begin >= row_end) return P2P_OK;
    int rc = ensure_texture(ctx, s);
    if (rc) return rc;
    const int band = row_end - row_begin;
    if ((band + 7) / 8 > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
    RowsParams P;
    memset(&P, 0, sizeof(P));
    P.tex = s.tex;
    P.out = d_out;
    P.W = W;
    P.H = H;
    P.v_begin = row_begin;
    P.v_end = row_end;
    P.n_chunks = (W / 2 + 1 + 31) / 32;
    P.seg_chunks = ctx->opt_seg_chunks;
    P.halfW = (float)(W / 2.0);
    P.halfH = (float)(H / 2.0);
    P.Wp_f = (float)s.Wp;
    P.Hp_f = (float)s.Hp;
    P.Umax = (float)(s.Wp - 1);
    P.Vmax = (float)(s.Hp - 1);
    P.inv_Wp = (float)(1.0 / (double)s.Wp);
    P.inv_Hp = (float)(1.0 / (double)s.Hp);
    const unsigned long long view_bytes = (unsigned long long)W * H * 3;
    std::vector<char> used((size_t)n_views, 0);
    int ng = 0, ny_max = 0;
    bool full = true;
    auto flush = [&]() -> int {
        if (ng == 0) return P2P_OK;
        for (int g = 0; g < ng; ++g) full = full && (P.grp[g].ny == ny_max);
        dim3 grid((P.n_chunks + P.seg_chunks - 1) / P.seg_chunks, (band + 7) / 8, ng);
        pick_rows(ctx->opt_trig == 0, full, ny_max)<<<grid, 256, 0, s.stream>>>(P);
        ctx->launches++;
        CK(cudaGetLastError());
        ng = 0;
        ny_max = 0;
        full = true;
        return P2P_OK;
    };
    for (int i = 0; i < n_views; ++i) {
        if (used[i]) continue;
        ViewGroup &G = P.grp[ng];
        memset(&G, 0, sizeof(G));
        G.pc = PitchC{pitch[i].f, pitch[i].c, pitch[i].s};
        for (int k = i; k < n_views && G.ny < 4; ++k) {
            if (used[k] || memcmp(&pitch[k], &pitch[i], sizeof(p2p_pitch_consts)) != 0) continue;
            used[k] = 1;
            G.shift_n[G.ny] = (float)((double)yaw_shift[k] / (double)s.Wp);
            G.out_off32[G.ny] = (unsigned)((unsigned long long)out_index[k] * view_bytes);
            G.ny++;
        }
        ny_max = (G.ny > ny_max) ? G.ny : ny_max;
        if (++ng == kMaxViewGroups) {
            rc = flush();
            if (rc) return rc;
        }
    }
    return flush();
}

// One or several (nb = 1, 2, 4) same-sized resident panoramas -> their view batches.  All launches
// go to the stream of the first slot.
int launch_project(p2p_ctx *ctx, Slot *const *sl, int nb, int n_yaw, const int32_t *yaw_shift, int n_pitch,
                   const p2p_pitch_consts *pitch, int W, int H, uint8_t *const *d_out) {
    Slot &s = *sl[0];
    for (int b = 0; b < nb; ++b) {
        // a slot filled by p2p_process_image holds only the rows its own views touch
        if (!slot_is_partial(*sl[b])) continue;
        if (ctx->opt_interp != 0)
            return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama: exact interpolation needs a full upload");
        int lo = 0, hi = 0;
        int rc = view_row_range(ctx, sl[b]->stream, n_pitch, pitch, W, H, sl[b]->Wp, sl[b]->Hp, &lo, &hi);
        if (rc) return rc;
        if (lo < sl[b]->row0 || hi + 1 > sl[b]->row1)
            return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama that does not cover these views: upload it again");
    }
    if (ctx->opt_sampler != 0 && ctx->opt_interp == 0) {
        for (int b = 0; b < nb; ++b) {
            int rc = ensure_texture(ctx, *sl[b]);
            if (rc) return rc;
        }
    }
    if (nb == 1 && rows_kernel_usable(ctx, W, H, n_yaw * n_pitch, d_out[0])) {
        // view (k, j) = yaw k, pitch j -> output index k * n_pitch + j, listed pitch-major so that the yaws of one
        // pitch land in the same group
        std::vector<int32_t> ys((size_t)n_yaw * n_pitch);
        std::vector<p2p_pitch_consts> ps((size_t)n_yaw * n_pitch);
        std::vector<int> oi((size_t)n_yaw * n_pitch);
        int n = 0;
        for (int j = 0; j < n_pitch; ++j)
            for (int k = 0; k < n_yaw; ++k, ++n) {
                ys[n] = yaw_shift[k];
                ps[n] = pitch[j];
                oi[n] = k * n_pitch + j;
            }
        return launch_rows(ctx, s, n, ys.data(), ps.data(), oi.data(), W, H, d_out[0]);
    }
    int ny_max = ctx->opt_ny < 1 ? 1 : (ctx->opt_ny > 4 ? 4 : ctx->opt_ny);
    // the kernel adds k * yaw_stride as a 32-bit offset: fall back to fewer yaws per launch for huge outputs
    const unsigned long long ys = (unsigned long long)W * H * 3 * (unsigned long long)n_pitch;
    while (ny_max > 1 && (unsigned long long)(ny_max - 1) * ys >= (1ull << 32)) --ny_max;
    ProjParams P;
    memset(&P, 0, sizeof(P));
    bool aligned = true;
    for (int b = 0; b < nb; ++b) {
        P.pano[b] = sl[b]->d_rgba;
        P.tex[b] = sl[b]->tex;
        P.out[b] = d_out[b];
        aligned = aligned && ((reinterpret_cast<uintptr_t>(d_out[b]) & 3) == 0);
    }
    P.view_stride = (unsigned long long)W * H * 3;
    P.yaw_stride = P.view_stride * (unsigned long long)n_pitch;
    P.yaw_stride32 = (ny_max > 1) ? (unsigned)P.yaw_stride : 0u;
    P.pitch_tex = s.pitch_tex;
    P.Wp = s.Wp;
    P.Hp = s.Hp;
    P.W = W;
    P.H = H;
    P.halfW = (float)(W / 2.0);
    P.halfH = (float)(H / 2.0);
    P.Wp_f = (float)s.Wp;
    P.Hp_f = (float)s.Hp;
    P.Umax = (float)(s.Wp - 1);
    // exact-bilinear mode with the seam-wrap option: U is limited to [0, Wp) instead of [0, Wp - 1], so a pixel whose azimuth
    // falls between the last and the first column interpolates between them (column Wp of the packed layout is column 0)
    if (ctx->opt_interp == 1 && ctx->opt_seam_wrap) P.Umax = nextafterf((float)s.Wp, 0.0f);
    P.Vmax = (float)(s.Hp - 1);
    P.inv_Wp = (float)(1.0 / (double)s.Wp);
    P.inv_Hp = (float)(1.0 / (double)s.Hp);
    P.numpy_trig = (ctx->opt_trig == 0);
    const bool quad = ((W & 3) == 0) && aligned;
    if (!quad && nb > 1) return fail(ctx, P2P_ERR_INVALID, "multi-image launches need W % 4 == 0 and aligned outputs");
    // chunk over yaws (<= 4 share one coordinate evaluation) and pitches (grid.z) so any list length works
    for (int y0 = 0; y0 < n_yaw; y0 += ny_max) {
        const int ny_l = (n_yaw - y0 < ny_max) ? n_yaw - y0 : ny_max;
        for (int k = 0; k < 4; ++k) {
            P.shift[k] = (k < ny_l) ? yaw_shift[y0 + k] : 0;
            P.shift_n[k] = (float)((double)P.shift[k] / (double)s.Wp);
        }
        P.yaw_off = y0;
        for (int p0 = 0; p0 < n_pitch; p0 += kMaxPitchPerLaunch) {
            const int np_l = (n_pitch - p0 < kMaxPitchPerLaunch) ? n_pitch - p0 : kMaxPitchPerLaunch;
            P.n_pitch = np_l;
            P.pitch_off = p0;
            for (int j = 0; j < np_l; ++j) {
                P.pc[j].f = pitch[p0 + j].f;
                P.pc[j].c = pitch[p0 + j].c;
                P.pc[j].s = pitch[p0 + j].s;
            }
            if (ctx->opt_interp == 1) {  // exact-bilinear mode (scipy map_coordinates order=1 arithmetic)
                if (nb != 1) return fail(ctx, P2P_ERR_INVALID, "exact interpolation mode renders one image per launch");
                dim3 egrid((W + 31) / 32, (H + 7) / 8, np_l);
                if (egrid.y > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
                switch (ny_l) {
                    case 1: project_exact_kernel<1><<<egrid, kThreads, 0, s.stream>>>(P); break;
                    case 2: project_exact_kernel<2><<<egrid, kThreads, 0, s.stream>>>(P); break;
                    case 3: project_exact_kernel<3><<<egrid, kThreads, 0, s.stream>>>(P); break;
                    default: project_exact_kernel<4><<<egrid, kThreads, 0, s.stream>>>(P); break;
                }
                ctx->launches++;
                CK(cudaGetLastError());
                continue;
            }
            // mirror-symmetric kernel: texture sampler, one image per launch, vector-store friendly sizes
            if (ctx->opt_mirror == 1 && ctx->opt_sampler == 1 && nb == 1 && quad && (W & 7) == 0) {
                dim3 mgrid((W / 2 + 1 + 31) / 32, (H + kMirRows - 1) / kMirRows, np_l);
                if (mgrid.y > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
                if (P.numpy_trig) {
                    switch (ny_l) {
                        case 1: project_mirror_kernel<1, true><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        case 2: project_mirror_kernel<2, true><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        case 3: project_mirror_kernel<3, true><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        default: project_mirror_kernel<4, true><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                    }
                } else {
                    switch (ny_l) {
                        case 1: project_mirror_kernel<1, false><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        case 2: project_mirror_kernel<2, false><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        case 3: project_mirror_kernel<3, false><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                        default: project_mirror_kernel<4, false><<<mgrid, kMirThreads, 0, s.stream>>>(P); break;
                    }
                }
                ctx->launches++;
                CK(cudaGetLastError());
                continue;
            }
            dim3 grid((W + 31) / 32, (H + 7) / 8, np_l);
            if (grid.y > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
            proj_fn fn = (ctx->opt_sampler == 1) ? pick_kernel<1>(quad, nb, ctx->opt_warp_w, ny_l)
                                                 : pick_kernel<0>(quad, nb, ctx->opt_warp_w, ny_l);
            fn<<<grid, kThreads, 0, s.stream>>>(P);
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    return P2P_OK;
}

int check_project_args(p2p_ctx *ctx, int slot, int n_yaw, const int32_t *yaw_shift, int n_pitch,
                       const p2p_pitch_consts *pitch, int W, int H, const void *out, int Wp) {
    if (!slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
    if (n_yaw <= 0 || n_pitch <= 0 || !yaw_shift || !pitch || !out)
        return fail(ctx, P2P_ERR_INVALID, "null or empty view list / output");
    if (W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "output size must be positive");
    if (W >= 32767 || H >= 32767) return fail(ctx, P2P_ERR_LIMIT, "output dimension >= 32767");
    for (int k = 0; k < n_yaw; ++k)
        if (yaw_shift[k] < 0 || yaw_shift[k] >= Wp) return fail(ctx, P2P_ERR_INVALID, "yaw_shift outside [0, Wp)");
    return P2P_OK;
}

// host BGR rows y0 .. min(y1, Hp - 1) -> staging -> packed rows y0 .. y1 (y1 == Hp adds the clamp row); caller holds the lock
int upload_rows(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride, int y0, int y1) {
    int rc = check_dims(ctx, Wp, Hp);
    if (rc) return rc;
    if (row_stride < (size_t)Wp * 3) return fail(ctx, P2P_ERR_INVALID, "row_stride smaller than Wp * 3");
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[slot];
    // tight device staging copy (row stride rounded to 4 bytes so the packer can use word loads)
    const size_t dstride = ((size_t)Wp * 3 + 3) & ~(size_t)3;
    rc = ensure(ctx, &s.d_bgr, &s.bgr_cap, dstride * Hp);
    if (rc) return rc;
    rc = prepare_slot(ctx, s, Wp, Hp);
    if (rc) return rc;
    const int ys1 = (y1 < Hp) ? y1 : Hp - 1;
    const size_t nrows = (size_t)(ys1 - y0 + 1);
    if (row_stride == dstride) {
        CK(cudaMemcpyAsync(s.d_bgr + (size_t)y0 * dstride, bgr + (size_t)y0 * row_stride, dstride * nrows,
                           cudaMemcpyHostToDevice, s.stream));
    } else {
        CK(cudaMemcpy2DAsync(s.d_bgr + (size_t)y0 * dstride, dstride, bgr + (size_t)y0 * row_stride, row_stride,
                             (size_t)Wp * 3, nrows, cudaMemcpyHostToDevice, s.stream));
    }
    return launch_pack(ctx, s, s.d_bgr, dstride, y0, y1);
}

// p2p_project_views with the context lock held
int project_views_locked(p2p_ctx *ctx, int slot, int n_yaw, const int32_t *yaw_shift, int n_pitch,
                         const p2p_pitch_consts *pitch, int W, int H, uint8_t *out, int out_on_device) {
    if (!slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
    Slot &s = ctx->slots[slot];
    if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
    int rc = check_project_args(ctx, slot, n_yaw, yaw_shift, n_pitch, pitch, W, H, out, s.Wp);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n_yaw * n_pitch * W * H * 3;
    uint8_t *d_out = out;
    if (!out_on_device) {
        rc = ensure(ctx, &s.d_out, &s.out_cap, bytes);
        if (rc) return rc;
        d_out = s.d_out;
    }
    Slot *sl[1] = {&s};
    uint8_t *outs[1] = {d_out};
    rc = launch_project(ctx, sl, 1, n_yaw, yaw_shift, n_pitch, pitch, W, H, outs);
    if (rc) return rc;
    if (!out_on_device) CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
    return P2P_OK;
}

// ---- JPEG encoder (p2p_jpeg.cuh) ---------------------------------------------------------------
// enqueue the encoder for n device images on the slot's stream; the files land in s.j_out, the sizes in s.j_sizes_h
int enqueue_jpeg(p2p_ctx *ctx, Slot &s, const uint8_t *d_bgr, int n, int W, int H, int quality, p2pjpeg::Geometry &G) {
    using namespace p2pjpeg;
    if (W >= 65536 || H >= 65536) return fail(ctx, P2P_ERR_LIMIT, "JPEG dimensions must be < 65536");
    G = make_geometry(W, H);
    if ((size_t)G.n_blocks * 64ull * 27ull >= (1ull << 32)) return fail(ctx, P2P_ERR_LIMIT, "image too large for the JPEG encoder");
    if (!ctx->j_err_h) {
        CK(cudaHostAlloc(reinterpret_cast<void **>(&ctx->j_err_h), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *ctx->j_err_h = 0;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&ctx->j_err_d), ctx->j_err_h, 0));
    }
    if (!ctx->d_jtab || ctx->jW != W || ctx->jH != H || ctx->jQ != quality) {
        // tables + file header per (size, quality), each in its own device buffer: launches of other slots that still read
        // another entry are not disturbed (no stream is synchronised)
        Tables *found = nullptr;
        for (const p2p_ctx::JTab &t : ctx->jtabs)
            if (t.W == W && t.H == H && t.Q == quality) found = t.d;
        if (!found) {
            if (ctx->jtabs.size() >= 64) {  // a pathological stream of sizes: start over once everything has drained
                for (int i = 0; i < ctx->n_slots; ++i) CK(cudaStreamSynchronize(ctx->slots[i].stream));
                for (const p2p_ctx::JTab &t : ctx->jtabs) cudaFree(t.d);
                ctx->jtabs.clear();
                ctx->d_jtab = nullptr;
            }
            Tables T;
            build_tables(W, H, quality, T);
            CK(cudaMalloc(reinterpret_cast<void **>(&found), sizeof(Tables)));
            CK(cudaMemcpy(found, &T, sizeof(T), cudaMemcpyHostToDevice));
            ctx->jtabs.push_back(p2p_ctx::JTab{W, H, quality, found});
        }
        ctx->d_jtab = found;
        ctx->jW = W; ctx->jH = H; ctx->jQ = quality;
    }
    const size_t nb = (size_t)n * G.blk_stride, chunks = G.cap_bits_words / 4;
    int rc = ensure(ctx, &s.j_coef, &s.j_coef_cap, (size_t)n * G.n_blocks * 64 * sizeof(int16_t));
    if (!rc) rc = ensure(ctx, &s.j_bits, &s.j_bits_cap, 2 * nb * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.j_stream, &s.j_stream_cap, (size_t)n * G.cap_bits_words * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.j_cnt, &s.j_cnt_cap, 2 * (size_t)n * chunks * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.j_out, &s.j_out_cap, (size_t)n * G.cap_out);
    if (!rc) rc = ensure(ctx, &s.j_tot, &s.j_tot_cap, 3 * (size_t)n * sizeof(unsigned long long));
    if (rc) return rc;
    if (s.j_sizes_n < n) {
        if (s.j_sizes_h) CK(cudaFreeHost(s.j_sizes_h));
        s.j_sizes_h = nullptr;
        CK(cudaHostAlloc(reinterpret_cast<void **>(&s.j_sizes_h), (size_t)n * sizeof(unsigned long long),
                         cudaHostAllocMapped | cudaHostAllocPortable));
        CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.j_sizes_d), s.j_sizes_h, 0));
        s.j_sizes_n = n;
    }
    uint32_t *bits = s.j_bits, *offs = s.j_bits + nb;
    uint32_t *cnt = s.j_cnt, *ffoff = s.j_cnt + (size_t)n * chunks;
    unsigned long long *tot_bits = s.j_tot, *tot_ff = s.j_tot + n;
    uint32_t *n_chunks = reinterpret_cast<uint32_t *>(s.j_tot + 2 * (size_t)n);
    cudaStream_t st = s.stream;
    CK(cudaMemsetAsync(s.j_stream, 0, (size_t)n * G.cap_bits_words * sizeof(uint32_t), st));
    jpeg_dct_kernel<<<dim3((G.mcux + kMcuPerCta - 1) / kMcuPerCta, G.mcuy, n), 64 * kMcuPerCta, 0, st>>>(d_bgr, s.j_coef, ctx->d_jtab, G);
    jpeg_size_kernel<<<dim3((G.n_blocks + 255) / 256, n), 256, 0, st>>>(s.j_coef, bits, ctx->d_jtab, G);
    jpeg_scan_kernel<<<n, 1024, 0, st>>>(bits, offs, nullptr, (uint32_t)G.n_blocks, (size_t)G.blk_stride, tot_bits);
    jpeg_emit_kernel<<<dim3((G.n_blocks + 255) / 256, n), 256, 0, st>>>(s.j_coef, offs, s.j_stream, ctx->d_jtab, G, tot_bits,
                                                                        ctx->j_err_d);
    const unsigned cgrid = (unsigned)((chunks + 255) / 256);
    jpeg_ffcount_kernel<<<dim3(cgrid, n), 256, 0, st>>>(s.j_stream, cnt, n_chunks, G, tot_bits);
    jpeg_scan_kernel<<<n, 1024, 0, st>>>(cnt, ffoff, n_chunks, 0u, chunks, tot_ff);
    jpeg_stuff_kernel<<<dim3(cgrid, n), 256, 0, st>>>(s.j_stream, ffoff, s.j_out, G, ctx->d_jtab, tot_bits, tot_ff, ctx->j_err_d);
    jpeg_finish_kernel<<<n, 256, 0, st>>>(s.j_out, G, ctx->d_jtab, tot_bits, tot_ff, s.j_sizes_d);
    ctx->launches += 8;
    CK(cudaGetLastError());
    return P2P_OK;
}

// wait for the slot's encoder and copy the files out (called WITHOUT the context lock: only stream calls)
int collect_jpeg(p2p_ctx *ctx, Slot &s, int n, const p2pjpeg::Geometry &G, uint8_t *out_host, size_t out_stride, size_t *sizes) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return P2P_ERR_CUDA;
    int rc = P2P_OK;
    for (int i = 0; i < n; ++i) {
        const unsigned long long sz = s.j_sizes_h[i];
        sizes[i] = (size_t)sz;
        if (sz == 0 || sz > out_stride) {
            rc = P2P_ERR_LIMIT;
            sizes[i] = 0;
            continue;
        }
        e = cudaMemcpyAsync(out_host + (size_t)i * out_stride, s.j_out + (size_t)i * G.cap_out, (size_t)sz,
                            cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) return P2P_ERR_CUDA;
    }
    e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return P2P_ERR_CUDA;
    (void)ctx;
    return rc;
}

// ---- JPEG decoder (p2p_jpegdec.cuh) ------------------------------------------------------------
// Huffman stage on the device (p2pjdec::huff_*): fills s.jd_coef_d.  Returns P2P_OK, an error, or 1 = "not handled"
// (no convergence, inconsistent block counts, unexpected markers): the caller then runs the host decoder.
// The destuffing pass runs on the calling thread; the lock is held only while enqueueing.
int ensure_jd_flags(p2p_ctx *ctx, Slot &s) {
    if (s.jd_flags_h) return P2P_OK;
    CK(cudaHostAlloc(reinterpret_cast<void **>(&s.jd_flags_h), sizeof(*s.jd_flags_h), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(s.jd_flags_h, 0, sizeof(*s.jd_flags_h));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.jd_flags_d), s.jd_flags_h, 0));
    return P2P_OK;
}

int device_huffman(p2p_ctx *ctx, Slot &s, const uint8_t *file, size_t len, const p2pjdec::Parsed &P) {
    using namespace p2pjdec;
    const Info &I = P.info;
    const uint32_t nb = (I.ncomp == 1) ? 1u : (uint32_t)(I.hmax * I.vmax + 2);
    const uint32_t total_mcus = (uint32_t)I.mcux * I.mcuy;
    const uint32_t total_blocks = total_mcus * nb;
    const uint32_t ivl_mcus = P.dri ? (uint32_t)P.dri : total_mcus;
    const uint32_t n_ivl = (total_mcus + ivl_mcus - 1) / ivl_mcus;
    // destuff into the pinned staging buffer (FF 00 -> FF; RSTn starts the next interval, byte-aligned; any other
    // marker ends the scan), then store the words MSB-first so a 32-bit window is one funnel shift
    uint8_t *dst = reinterpret_cast<uint8_t *>(s.jd_coef_h);
    const size_t cap = s.jd_coef_h_cap;
    size_t n = 0;
    std::vector<uint32_t> ivl_byte(1, 0u);   // byte offset of every interval in the destuffed stream
    {
        const uint8_t *p = file + P.ecs, *end = file + len;
        while (p < end) {
            const uint8_t *ff = static_cast<const uint8_t *>(memchr(p, 0xFF, (size_t)(end - p)));
            const size_t run = ff ? (size_t)(ff - p) : (size_t)(end - p);
            if (n + run + 16 > cap) return 1;
            memcpy(dst + n, p, run);
            n += run;
            if (!ff || ff + 1 >= end) return 1;      // no EOI: truncated file
            const uint8_t m = ff[1];
            if (m == 0) {
                dst[n++] = 0xFF;
                p = ff + 2;
            } else if (m >= 0xD0 && m <= 0xD7 && P.dri) {
                if (m != 0xD0 + ((ivl_byte.size() - 1) & 7)) return 1;   // out of sequence: damaged (see decode_scan)
                if (n >= (1ull << 29)) return 1;
                ivl_byte.push_back((uint32_t)n);
                p = ff + 2;
            } else if (m == 0xFF) {
                p = ff + 1;                          // fill byte before a marker
            } else {
                break;                               // EOI for a complete file
            }
        }
    }
    if (n == 0 || n * 8 >= (1ull << 32) || ivl_byte.size() != n_ivl) return 1;
    const size_t n_words = (n + 3) / 4 + 3;
    memset(dst + n, 0, n_words * 4 - n);
    uint32_t *w = reinterpret_cast<uint32_t *>(dst);
    for (size_t i = 0; i < n_words; ++i) w[i] = __builtin_bswap32(w[i]);
    // subsequences: a regular kSubBits grid inside every interval
    ivl_byte.push_back((uint32_t)n);
    std::vector<SubSeq> subs;
    std::vector<uint32_t> ivl_first(n_ivl + 1);
    subs.reserve(n * 8 / kSubBits + n_ivl + 1);
    for (uint32_t k = 0; k < n_ivl; ++k) {
        ivl_first[k] = (uint32_t)subs.size();
        const uint32_t b0 = ivl_byte[k] * 8u, b1 = ivl_byte[k + 1] * 8u;
        if (b1 <= b0) return 1;
        for (uint32_t b = b0; b < b1; b += kSubBits) {
            SubSeq q;
            q.begin = b;
            q.end = (b + kSubBits < b1) ? b + kSubBits : b1;
            q.ivl = k;
            q.first = (b == b0) ? 1u : 0u;
            subs.push_back(q);
        }
    }
    ivl_first[n_ivl] = (uint32_t)subs.size();

    HuffGeom G;
    memset(&G, 0, sizeof(G));
    G.n_bits = (uint32_t)(n * 8);
    G.n_sub = (uint32_t)subs.size();
    G.nb = (int)nb;
    G.n_luma = I.hmax * I.vmax;
    G.hmax = I.hmax; G.vmax = I.vmax; G.mcux = I.mcux;
    G.total_blocks = total_blocks;
    G.n_ivl = n_ivl;
    G.ivl_blocks = ivl_mcus * nb;
    uint32_t max_dc = 0;
    for (int c = 0; c < 3; ++c) {
        G.bw[c] = I.bw[c];
        G.coef_off[c] = I.coef_off[c];
        G.dc_count[c] = (c >= I.ncomp) ? 0u : total_mcus * (c ? 1u : (uint32_t)(I.hmax * I.vmax));
        max_dc = G.dc_count[c] > max_dc ? G.dc_count[c] : max_dc;
    }
    G.dc_stride = (max_dc + 3) & ~3u;
    const size_t nsub4 = ((size_t)G.n_sub + 3) & ~(size_t)3;
    // device tables: the three per-component tables, then the unified look-up of the synchronisation rounds
    struct DevTables {
        DevHuff T[3];
        SyncLut L;
    };
    std::vector<unsigned char> tables_mem(sizeof(DevTables));
    DevTables &DT = *reinterpret_cast<DevTables *>(tables_mem.data());
    DevHuff *T = DT.T;
    for (int c = 0; c < 3; ++c) build_sync_lut(P.dc[P.td[c]], P.ac[P.ta[c]], DT.L.e[c][0], DT.L.e[c][1], DT.L.w[c][0], DT.L.w[c][1]);
    for (int c = 0; c < 3; ++c) {
        const HuffTable &d = P.dc[P.td[c]], &a = P.ac[P.ta[c]];
        memcpy(T[c].dc_look, d.look, sizeof(d.look));
        memcpy(T[c].dc_maxcode, d.maxcode, sizeof(d.maxcode));
        memcpy(T[c].dc_valoff, d.valoff, sizeof(d.valoff));
        memcpy(T[c].dc_vals, d.vals, sizeof(T[c].dc_vals));
        memcpy(T[c].ac_fast, a.fast_ac, sizeof(a.fast_ac));
        memcpy(T[c].ac_look, a.look, sizeof(a.look));
        memcpy(T[c].ac_maxcode, a.maxcode, sizeof(a.maxcode));
        memcpy(T[c].ac_valoff, a.valoff, sizeof(a.valoff));
        memcpy(T[c].ac_vals, a.vals, sizeof(a.vals));
    }
    cudaStream_t st = s.stream;
    const unsigned sgrid = (G.n_sub + 127) / 128;
    SubSeq *d_sub = nullptr;
    uint32_t *d_ivl_first = nullptr;
    const SyncLut *d_lut = nullptr;
    bool fast_rounds = true;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        fast_rounds = ctx->opt_gpu_huffman != 2;   // 2 = the plain rounds (tables in global memory), the tests' yardstick
        CK(cudaSetDevice(ctx->device));
        const size_t sub_bytes = ((size_t)G.n_sub * sizeof(SubSeq) + 15) & ~(size_t)15;
        int rc = ensure_grow(ctx, &s.jd_stream, &s.jd_stream_cap, n_words * 4);
        if (!rc) rc = ensure_grow(ctx, &s.jd_states, &s.jd_states_cap, 2 * (size_t)G.n_sub * sizeof(unsigned long long));
        if (!rc) rc = ensure_grow(ctx, &s.jd_nblk, &s.jd_nblk_cap, 2 * nsub4 * sizeof(uint32_t));
        if (!rc) rc = ensure(ctx, &s.jd_dc, &s.jd_dc_cap, 2 * 3 * (size_t)G.dc_stride * sizeof(int32_t));
        if (!rc) rc = ensure(ctx, &s.jd_tiles, &s.jd_tiles_cap, 2 * 3 * ((((size_t)G.dc_stride + 4095) / 4096 + 3) & ~(size_t)3) * sizeof(uint32_t));
        if (!rc) rc = ensure(ctx, &s.jd_coef_d, &s.jd_coef_d_cap, I.n_coef * sizeof(int16_t));
        if (!rc) rc = ensure_grow(ctx, &s.jd_sub, &s.jd_sub_cap, sub_bytes + ((size_t)n_ivl + 1) * sizeof(uint32_t));
        if (rc) return rc;
        d_sub = reinterpret_cast<SubSeq *>(s.jd_sub);
        d_ivl_first = reinterpret_cast<uint32_t *>(s.jd_sub + sub_bytes);
        if (!s.jd_tables) CK(cudaMalloc(reinterpret_cast<void **>(&s.jd_tables), sizeof(DevTables)));
        if (!s.jd_tot_d) CK(cudaMalloc(reinterpret_cast<void **>(&s.jd_tot_d), 4 * sizeof(unsigned long long)));
        int frc = ensure_jd_flags(ctx, s);
        if (frc) return frc;
        CK(cudaMemcpyAsync(s.jd_stream, w, n_words * 4, cudaMemcpyHostToDevice, st));
        // the tables below live in pageable memory: cudaMemcpyAsync stages them before it returns
        CK(cudaMemcpyAsync(s.jd_tables, &DT, sizeof(DevTables), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_sub, subs.data(), (size_t)G.n_sub * sizeof(SubSeq), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_ivl_first, ivl_first.data(), ((size_t)n_ivl + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(s.jd_coef_d, 0, I.n_coef * sizeof(int16_t), st));
        s.jd_flags_h->bad = 0;
        d_lut = reinterpret_cast<const SyncLut *>(reinterpret_cast<const unsigned char *>(s.jd_tables) + offsetof(DevTables, L));
        if (fast_rounds)
            huff_sync_fast_kernel<<<sgrid, kSyncThreads, 0, st>>>(s.jd_stream, s.jd_tables, d_lut, G, d_sub, s.jd_states,
                                                                 s.jd_states + G.n_sub, s.jd_nblk, 1, &s.jd_flags_d->changed[0]);
        else
            huff_sync_kernel<<<sgrid, 128, 0, st>>>(s.jd_stream, s.jd_tables, G, d_sub, s.jd_states, s.jd_states + G.n_sub,
                                                  s.jd_nblk, 1, &s.jd_flags_d->changed[0]);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    // synchronisation rounds: each needs the "anything changed" flag back on the host
    bool converged = false;
    for (int round = 0; round < kMaxSyncRounds; round += kRoundsPerCheck) {
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            CK(cudaSetDevice(ctx->device));
            // several rounds per host check (a round in which nothing moves costs one early-exit pass); the flag that
            // decides convergence is the one of the LAST round of the batch
            for (int r = 0; r < kRoundsPerCheck; ++r) {
                s.jd_flags_h->changed[r] = 0;   // nothing of this slot is running: the stream was drained above
                if (fast_rounds)
                    huff_sync_fast_kernel<<<sgrid, kSyncThreads, 0, st>>>(s.jd_stream, s.jd_tables, d_lut, G, d_sub, s.jd_states,
                                                                         s.jd_states + G.n_sub, s.jd_nblk, 0,
                                                                         &s.jd_flags_d->changed[r]);
                else
                    huff_sync_kernel<<<sgrid, 128, 0, st>>>(s.jd_stream, s.jd_tables, G, d_sub, s.jd_states,
                                                          s.jd_states + G.n_sub, s.jd_nblk, 0, &s.jd_flags_d->changed[r]);
            }
            ctx->launches += kRoundsPerCheck;
            CK(cudaGetLastError());
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) return P2P_ERR_CUDA;
        if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->changed[kRoundsPerCheck - 1]) == 0) {
            converged = true;
            break;
        }
    }
    if (!converged) return 1;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        p2pjpeg::jpeg_scan_kernel<<<1, 1024, 0, st>>>(s.jd_nblk, s.jd_nblk + nsub4, nullptr, G.n_sub, nsub4, &s.jd_flags_d->total);
        huff_check_kernel<<<(n_ivl + 255) / 256, 256, 0, st>>>(s.jd_nblk + nsub4, s.jd_nblk, d_ivl_first, G, &s.jd_flags_d->bad);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return P2P_ERR_CUDA;
    // every interval must hold its quota of blocks (its padding bits may decode as a few more): else damaged data
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->bad) != 0) return 1;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    int32_t *dcdiff = s.jd_dc;
    uint32_t *dcsum = reinterpret_cast<uint32_t *>(s.jd_dc + 3 * (size_t)G.dc_stride);
    if (fast_rounds)
        huff_write_fast_kernel<<<sgrid, kSyncThreads, 0, st>>>(s.jd_stream, s.jd_tables, d_lut, G, d_sub, d_ivl_first, s.jd_states,
                                                              s.jd_nblk + nsub4, s.jd_coef_d, dcdiff, &s.jd_flags_d->out_of_range);
    else
        huff_write_kernel<<<sgrid, 128, 0, st>>>(s.jd_stream, s.jd_tables, G, d_sub, d_ivl_first, s.jd_states, s.jd_nblk + nsub4,
                                               s.jd_coef_d, dcdiff, &s.jd_flags_d->out_of_range);
    // per component: exclusive sums of the DC differences in scan order (mod 2^32 arithmetic = two's complement sums)
    CK(cudaMemcpyAsync(s.jd_nblk, G.dc_count, 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));  // reuse as n_per_image
    {   // three-phase scan over the whole GPU (the luma plane of an 8K file has 524,288 differences)
        const uint32_t n_tiles = (G.dc_stride + 4095u) / 4096u;
        const size_t tiles_stride = ((size_t)n_tiles + 3) & ~(size_t)3;
        uint32_t *tile_sums = s.jd_tiles, *tile_offs = s.jd_tiles + 3 * tiles_stride;
        const uint32_t *in = reinterpret_cast<const uint32_t *>(dcdiff);
        p2pjpeg::scan_tile_sums_kernel<<<dim3(n_tiles, 3), 1024, 0, st>>>(in, s.jd_nblk, 0u, (size_t)G.dc_stride, tile_sums,
                                                                         tiles_stride);
        p2pjpeg::jpeg_scan_kernel<<<3, 1024, 0, st>>>(tile_sums, tile_offs, nullptr, n_tiles, tiles_stride, s.jd_tot_d);
        p2pjpeg::scan_tiles_apply_kernel<<<dim3(n_tiles, 3), 1024, 0, st>>>(in, dcsum, s.jd_nblk, 0u, (size_t)G.dc_stride,
                                                                           tile_offs, tiles_stride);
        ctx->launches += 2;
    }
    huff_dc_kernel<<<dim3((G.dc_stride + 255) / 256, 3), 256, 0, st>>>(dcdiff, dcsum, G, s.jd_coef_d);
    ctx->launches += 3;
    CK(cudaGetLastError());
    ctx->gpu_huffman_used++;
    return P2P_OK;
}

// Decode `file` into the slot's BGR staging image (s.d_bgr, row stride = Wp * 3 rounded to 4).  The Huffman stage
// runs on the device (files without restart markers) or on the calling thread WITHOUT the context lock; the lock is
// only taken to size buffers and to enqueue.
int decode_jpeg_to_staging(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, p2pjdec::Parsed &P, size_t *dstride) {
    using namespace p2pjdec;
    if (parse_headers(file, len, P)) return P2P_ERR_UNSUPPORTED;
    const Info &I = P.info;
    Slot &s = ctx->slots[slot];
    int use_gpu = 0;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        int rc = check_dims(ctx, I.W, I.H);
        if (rc) return rc;
        CK(cudaSetDevice(ctx->device));
        use_gpu = ctx->opt_gpu_huffman;
        rc = ensure_jd_flags(ctx, s);
        if (rc) return rc;
        const size_t bytes = I.n_coef * sizeof(int16_t);
        if (s.jd_coef_h_cap < bytes) {
            CK(cudaStreamSynchronize(s.stream));  // an earlier upload may still read the old staging buffer
            if (s.jd_coef_h) CK(cudaFreeHost(s.jd_coef_h));
            s.jd_coef_h = nullptr;
            s.jd_coef_h_cap = 0;
            CK(cudaHostAlloc(reinterpret_cast<void **>(&s.jd_coef_h), bytes, cudaHostAllocPortable));
            s.jd_coef_h_cap = bytes;
        } else {
            CK(cudaStreamSynchronize(s.stream));
        }
        s.jd_flags_h->out_of_range = 0;   // "damaged data" flag of the write pass and the IDCT; the stream is drained
    }
    bool coef_on_device = false;
    if (use_gpu) {
        const int rc = device_huffman(ctx, s, file, len, P);
        if (rc == P2P_OK) coef_on_device = true;
        else if (rc != 1) return rc;
        else {
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->gpu_huffman_fallback++;
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(s.stream);   // the staging buffer was used for the stream upload
        }
    }
    if (!coef_on_device && decode_scan(file, len, P, s.jd_coef_h)) return P2P_ERR_UNSUPPORTED;  // damaged: leave it to libjpeg
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    size_t plane_off[3], plane_bytes = 0;
    for (int k = 0; k < 3; ++k) {
        plane_off[k] = plane_bytes;
        plane_bytes += (size_t)I.bw[k] * I.bh[k] * 64;
    }
    *dstride = ((size_t)I.W * 3 + 3) & ~(size_t)3;
    int rc = ensure(ctx, &s.jd_coef_d, &s.jd_coef_d_cap, I.n_coef * sizeof(int16_t));
    if (!rc) rc = ensure(ctx, &s.jd_planes, &s.jd_planes_cap, plane_bytes);
    if (!rc) rc = ensure(ctx, &s.d_bgr, &s.bgr_cap, *dstride * I.H);
    if (rc) return rc;
    if (!coef_on_device)
        CK(cudaMemcpyAsync(s.jd_coef_d, s.jd_coef_h, I.n_coef * sizeof(int16_t), cudaMemcpyHostToDevice, s.stream));
    for (int k = 0; k < 3; ++k) {
        Quant Q;
        memcpy(Q.q, I.quant[k], sizeof(Q.q));
        const int nb = I.bw[k] * I.bh[k];
        jpegdec_idct_kernel<<<(nb + 31) / 32, 256, 0, s.stream>>>(s.jd_coef_d + I.coef_off[k], s.jd_planes + plane_off[k], Q, nb,
                                                                  I.bw[k], I.bw[k] * 8, &s.jd_flags_d->out_of_range);
    }
    ColorParams C;
    C.y = s.jd_planes + plane_off[0];
    C.cb = s.jd_planes + plane_off[1];
    C.cr = s.jd_planes + plane_off[2];
    C.pitch_y = I.bw[0] * 8;
    C.pitch_c = I.bw[1] * 8;
    C.W = I.W; C.H = I.H; C.hmax = I.hmax; C.vmax = I.vmax; C.cw = I.cw; C.ch = I.ch;
    C.bgr = s.d_bgr;
    C.stride = *dstride;
    if (I.H > 65535) return fail(ctx, P2P_ERR_LIMIT, "image too tall for one grid");
    jpegdec_color_kernel<<<dim3(((I.W + 3) / 4 + 255) / 256, I.H), 256, 0, s.stream>>>(C);
    ctx->launches += 4;
    CK(cudaGetLastError());
    return P2P_OK;
}

// ---- PNG encoder (p2p_png.cuh) -----------------------------------------------------------------
// enqueue the encoder for n device images on the slot's stream; files land in s.j_out, sizes in s.j_sizes_h
// (size 0 = this image is not handled on the device: the caller uses cv2.imwrite for it)
int enqueue_png(p2p_ctx *ctx, Slot &s, const uint8_t *d_bgr, int n, int W, int H, p2ppng::Geom &G) {
    using namespace p2ppng;
    if (W >= 32767 || H >= 32767) return fail(ctx, P2P_ERR_LIMIT, "PNG dimensions must be < 32767");
    memset(&G, 0, sizeof(G));
    G.W = W; G.H = H; G.n = n;
    G.row_bytes = 1u + 3u * (uint32_t)W;
    const unsigned long long N64 = (unsigned long long)H * G.row_bytes;
    if (N64 >= (1ull << 31)) return fail(ctx, P2P_ERR_LIMIT, "image too large for the PNG encoder");
    G.N = (uint32_t)N64;
    G.Npad = (G.N + 4095u) & ~4095u;
    G.max_blk = G.N / (uint32_t)kSymPerBlock + 2u;
    G.img_stride = (size_t)W * H * 3;
    G.z_cap = (((size_t)G.N + G.N / 8 + 1024) + 15) & ~(size_t)15;
    G.out_cap = (G.z_cap + (G.z_cap / kIdat + 2) * 12 + 64 + 15) & ~(size_t)15;
    if (!ctx->d_crc_table) {
        uint32_t table[5 * 256];   // CRC-32 byte table + the 4 byte tables of "advance the register by 256 zero bytes"
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        for (int j = 0; j < 4; ++j)
            for (uint32_t b = 0; b < 256; ++b) {
                uint32_t c = b << (8 * j);
                for (int k = 0; k < 256; ++k) c = table[c & 0xFFu] ^ (c >> 8);
                table[256 * (j + 1) + b] = c;
            }
        CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_crc_table), sizeof(table)));
        CK(cudaMemcpy(ctx->d_crc_table, table, sizeof(table), cudaMemcpyHostToDevice));
    }
    const size_t npos = (size_t)n * G.Npad, nblk = (size_t)n * G.max_blk;
    const size_t chunk_words = (size_t)n * ((G.N + kScanChunk - 1) / kScanChunk);
    const size_t blk_words = nblk * 2 + (((size_t)n + 3) & ~(size_t)3) + nblk * (kLCodes + 2) + 2 * chunk_words;
    int rc = ensure(ctx, &s.pg_F, &s.pg_F_cap, npos);
    if (!rc) rc = ensure(ctx, &s.pg_S, &s.pg_S_cap, npos * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.pg_tlen, &s.pg_tlen_cap, npos * sizeof(uint16_t));
    if (!rc) rc = ensure(ctx, &s.pg_blk, &s.pg_blk_cap, blk_words * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.pg_info, &s.pg_info_cap, nblk * sizeof(BlockInfo));
    if (!rc) rc = ensure(ctx, &s.pg_Z, &s.pg_Z_cap, (size_t)n * G.z_cap);
    if (!rc) rc = ensure(ctx, &s.pg_sums, &s.pg_sums_cap, 3 * (size_t)n * sizeof(unsigned long long));
    if (!rc) rc = ensure(ctx, &s.j_out, &s.j_out_cap, (size_t)n * G.out_cap);
    if (rc) return rc;
    if (s.j_sizes_n < n) {
        if (s.j_sizes_h) CK(cudaFreeHost(s.j_sizes_h));
        s.j_sizes_h = nullptr;
        CK(cudaHostAlloc(reinterpret_cast<void **>(&s.j_sizes_h), (size_t)n * sizeof(unsigned long long),
                         cudaHostAllocMapped | cudaHostAllocPortable));
        CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.j_sizes_d), s.j_sizes_h, 0));
        s.j_sizes_n = n;
    }
    uint32_t *blockpos = s.pg_blk, *blkoff = s.pg_blk + nblk, *ntok = s.pg_blk + 2 * nblk;
    uint32_t *lfreq = ntok + (((size_t)n + 3) & ~(size_t)3);
    uint32_t *chunk_agg = lfreq + nblk * (kLCodes + 2), *chunk_carry = chunk_agg + chunk_words;
    unsigned long long *sums = s.pg_sums, *zbits = s.pg_sums + 2 * (size_t)n;
    cudaStream_t st = s.stream;
    if (H > 65535) return fail(ctx, P2P_ERR_LIMIT, "image too tall for one grid");
    CK(cudaMemsetAsync(s.pg_Z, 0, (size_t)n * G.z_cap, st));
    CK(cudaMemsetAsync(s.pg_sums, 0, 3 * (size_t)n * sizeof(unsigned long long), st));
    png_filter_kernel<<<dim3((W + 255) / 256, H, n), 256, 0, st>>>(d_bgr, s.pg_F, G);
    const unsigned n_chunks = (G.N + kScanChunk - 1) / kScanChunk;
    const dim3 sgrid(n_chunks, n);
    png_scan_kernel<0, false><<<sgrid, 1024, 0, st>>>(s.pg_F, s.pg_S, nullptr, nullptr, chunk_agg, nullptr, G);
    png_chunk_carry_kernel<0><<<(n + 31) / 32, 32, 0, st>>>(chunk_agg, chunk_carry, nullptr, G);
    png_scan_kernel<0, true><<<sgrid, 1024, 0, st>>>(s.pg_F, s.pg_S, nullptr, nullptr, nullptr, chunk_carry, G);
    png_scan_kernel<1, false><<<sgrid, 1024, 0, st>>>(s.pg_F, s.pg_S, nullptr, nullptr, chunk_agg, nullptr, G);
    png_chunk_carry_kernel<1><<<(n + 31) / 32, 32, 0, st>>>(chunk_agg, chunk_carry, ntok, G);
    png_scan_kernel<1, true><<<sgrid, 1024, 0, st>>>(s.pg_F, s.pg_S, s.pg_tlen, blockpos, nullptr, chunk_carry, G);
    png_hist_kernel<<<dim3(G.max_blk, n), 256, 0, st>>>(s.pg_F, s.pg_tlen, blockpos, ntok, lfreq, G);
    png_tree_kernel<<<dim3((G.max_blk + 31) / 32, n), 32, 0, st>>>(lfreq, ntok, blockpos, s.pg_info, G);
    png_layout_kernel<<<(n + 31) / 32, 32, 0, st>>>(s.pg_info, ntok, blkoff, zbits, G);
    png_emit_kernel<<<dim3(G.max_blk, n), 256, 0, st>>>(s.pg_F, s.pg_tlen, blockpos, ntok, s.pg_info, blkoff, zbits, s.pg_Z, G);
    png_adler_kernel<<<dim3(64, n), 256, 0, st>>>(s.pg_F, sums, G);
    png_pack_kernel<<<dim3(128, n), 256, 0, st>>>(s.pg_Z, zbits, sums, s.j_out, G);
    const unsigned max_chunks = (unsigned)(G.z_cap / kIdat + 1);
    png_finish_kernel<<<dim3((max_chunks + 7) / 8, n), 256, 0, st>>>(s.j_out, zbits, ctx->d_crc_table, s.j_sizes_d, G);
    ctx->launches += 14;
    CK(cudaGetLastError());
    return P2P_OK;
}

// wait and copy the PNG files out; sizes[i] = 0 marks an image the device encoder does not handle
int collect_png(Slot &s, int n, size_t cap_per_image, uint8_t *out_host, size_t out_stride, size_t *sizes) {
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return P2P_ERR_CUDA;
    int rc = P2P_OK;
    for (int i = 0; i < n; ++i) {
        const unsigned long long sz = s.j_sizes_h[i];
        sizes[i] = (size_t)sz;
        if (sz == 0) continue;
        if (sz > out_stride) {
            rc = P2P_ERR_LIMIT;
            sizes[i] = 0;
            continue;
        }
        if (cudaMemcpyAsync(out_host + (size_t)i * out_stride, s.j_out + (size_t)i * cap_per_image, (size_t)sz,
                            cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
            return P2P_ERR_CUDA;
    }
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return P2P_ERR_CUDA;
    return rc;
}

}  // namespace

// ============================================================================================
extern "C" {

int p2p_abi_version(void) { return P2P_ABI_VERSION; }

int p2p_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *p2p_status_string(int status) {
    switch (status) {
        case P2P_OK: return "ok";
        case P2P_ERR_INVALID: return "invalid argument";
        case P2P_ERR_CUDA: return "CUDA error";
        case P2P_ERR_NOMEM: return "out of memory";
        case P2P_ERR_STATE: return "bad slot state";
        case P2P_ERR_LIMIT: return "size limit exceeded";
        case P2P_ERR_UNSUPPORTED: return "file outside the supported subset (use cv2.imread)";
        default: return "unknown status";
    }
}

int p2p_create(int device, int n_slots, p2p_ctx **out) {
    if (!out || n_slots <= 0 || n_slots > 64) return P2P_ERR_INVALID;
    *out = nullptr;
    int n = p2p_device_count();
    if (n <= 0 || device < 0 || device >= n) return P2P_ERR_CUDA;  // no CPU fallback
    p2p_ctx *ctx = new (std::nothrow) p2p_ctx;
    if (!ctx) return P2P_ERR_NOMEM;
    ctx->device = device;
    ctx->n_slots = n_slots;
    ctx->slots = new (std::nothrow) Slot[n_slots];
    if (!ctx->slots) {
        delete ctx;
        return P2P_ERR_NOMEM;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        delete[] ctx->slots;
        delete ctx;
        return P2P_ERR_CUDA;
    }
    for (int i = 0; i < n_slots; ++i) {
        if (cudaStreamCreateWithFlags(&ctx->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < i; ++j) cudaStreamDestroy(ctx->slots[j].stream);
            delete[] ctx->slots;
            delete ctx;
            return P2P_ERR_CUDA;
        }
        ctx->slots[i].own_stream = true;
    }
    *out = ctx;
    return P2P_OK;
}

void p2p_destroy(p2p_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < ctx->n_slots; ++i) {
        Slot &s = ctx->slots[i];
        if (s.tex) cudaDestroyTextureObject(s.tex);
        if (s.surf) cudaDestroySurfaceObject(s.surf);
        if (s.arr) cudaFreeArray(s.arr);
        cudaFree(s.d_bgr);
        cudaFree(s.d_rgba);
        cudaFree(s.d_out);
        cudaFree(s.d_tab);
        cudaFree(s.j_coef);
        cudaFree(s.j_bits);
        cudaFree(s.j_stream);
        cudaFree(s.j_cnt);
        cudaFree(s.j_out);
        cudaFree(s.j_tot);
        if (s.j_sizes_h) cudaFreeHost(s.j_sizes_h);
        if (s.jd_coef_h) cudaFreeHost(s.jd_coef_h);
        cudaFree(s.jd_coef_d);
        cudaFree(s.jd_planes);
        cudaFree(s.pg_F);
        cudaFree(s.pg_S);
        cudaFree(s.pg_tlen);
        cudaFree(s.pg_blk);
        cudaFree(s.pg_info);
        cudaFree(s.pg_Z);
        cudaFree(s.pg_sums);
        cudaFree(s.jd_stream);
        cudaFree(s.jd_states);
        cudaFree(s.jd_nblk);
        cudaFree(s.jd_tables);
        cudaFree(s.jd_dc);
        cudaFree(s.jd_tiles);
        cudaFree(s.jd_tot_d);
        cudaFree(s.jd_sub);
        if (s.jd_flags_h) cudaFreeHost(s.jd_flags_h);
        if (s.own_stream && s.stream) cudaStreamDestroy(s.stream);
        if (s.owned) cudaStreamDestroy(s.owned);
    }
    cudaFree(ctx->d_flush);
    cudaFree(ctx->d_range);
    for (const p2p_ctx::JTab &t : ctx->jtabs) cudaFree(t.d);
    cudaFree(ctx->d_crc_table);
    if (ctx->j_err_h) cudaFreeHost(ctx->j_err_h);
    cudaGetLastError();
    delete[] ctx->slots;
    delete ctx;
}

const char *p2p_last_error(p2p_ctx *ctx) {
    if (!ctx) return "null context";
    return (tl_err_ctx == ctx) ? tl_err.c_str() : "";  // valid until this thread's next failing call
}

int p2p_set_option(p2p_ctx *ctx, int key, int value) {
    if (!ctx) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    switch (key) {
        case P2P_OPT_SAMPLER:
            if (value != 0 && value != 1) return fail(ctx, P2P_ERR_INVALID, "sampler must be 0 or 1");
            ctx->opt_sampler = value;
            return P2P_OK;
        case P2P_OPT_WARP_W:
            if (value != 8 && value != 32) return fail(ctx, P2P_ERR_INVALID, "warp_w must be 8 or 32");
            ctx->opt_warp_w = value;
            return P2P_OK;
        case P2P_OPT_YAWS_PER_THREAD:
            if (value < 1 || value > 4) return fail(ctx, P2P_ERR_INVALID, "yaws per thread must be 1..4");
            ctx->opt_ny = value;
            return P2P_OK;
        case P2P_OPT_IMAGES_PER_LAUNCH:
            if (value != 1 && value != 2 && value != 4) return fail(ctx, P2P_ERR_INVALID, "images per launch must be 1, 2 or 4");
            ctx->opt_nb = value;
            return P2P_OK;
        case P2P_OPT_MIRROR:
            if (value < 0 || value > 2) return fail(ctx, P2P_ERR_INVALID, "mirror must be 0, 1 or 2");
            ctx->opt_mirror = value;
            return P2P_OK;
        case P2P_OPT_SEG_CHUNKS:
            if (value < 1 || value > 64) return fail(ctx, P2P_ERR_INVALID, "seg_chunks must be 1..64");
            ctx->opt_seg_chunks = value;
            return P2P_OK;
        case P2P_OPT_INTERP:
            if (value != 0 && value != 1) return fail(ctx, P2P_ERR_INVALID, "interp must be 0 (cv2 fixed point) or 1 (exact bilinear)");
            ctx->opt_interp = value;
            return P2P_OK;
        case P2P_OPT_SEAM_WRAP:
            if (value != 0 && value != 1) return fail(ctx, P2P_ERR_INVALID, "seam_wrap must be 0 or 1");
            ctx->opt_seam_wrap = value;
            return P2P_OK;
        case P2P_OPT_TRIG:
            if (value != 0 && value != 1) return fail(ctx, P2P_ERR_INVALID, "trig must be 0 (NumPy-exact) or 1 (minimax)");
            ctx->opt_trig = value;
            return P2P_OK;
        case P2P_OPT_PARTIAL_UPLOAD:
            if (value != 0 && value != 1) return fail(ctx, P2P_ERR_INVALID, "partial_upload must be 0 or 1");
            ctx->opt_partial = value;
            return P2P_OK;
        case P2P_OPT_GPU_HUFFMAN:
            if (value < 0 || value > 2) return fail(ctx, P2P_ERR_INVALID, "gpu_huffman must be 0, 1 or 2");
            ctx->opt_gpu_huffman = value;
            return P2P_OK;

        default:
            return fail(ctx, P2P_ERR_INVALID, "unknown option");
    }
}

int p2p_get_option(p2p_ctx *ctx, int key, int *value) {
    if (!ctx || !value) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    switch (key) {
        case P2P_OPT_SAMPLER: *value = ctx->opt_sampler; return P2P_OK;
        case P2P_OPT_WARP_W: *value = ctx->opt_warp_w; return P2P_OK;
        case P2P_OPT_YAWS_PER_THREAD: *value = ctx->opt_ny; return P2P_OK;
        case P2P_OPT_COUNT_LAUNCHES: *value = (int)ctx->launches; return P2P_OK;
        case P2P_OPT_IMAGES_PER_LAUNCH: *value = ctx->opt_nb; return P2P_OK;
        case P2P_OPT_MIRROR: *value = ctx->opt_mirror; return P2P_OK;
        case P2P_OPT_SEG_CHUNKS: *value = ctx->opt_seg_chunks; return P2P_OK;
        case P2P_OPT_INTERP: *value = ctx->opt_interp; return P2P_OK;
        case P2P_OPT_SEAM_WRAP: *value = ctx->opt_seam_wrap; return P2P_OK;
        case P2P_OPT_TRIG: *value = ctx->opt_trig; return P2P_OK;
        case P2P_OPT_PARTIAL_UPLOAD: *value = ctx->opt_partial; return P2P_OK;
        case P2P_OPT_GPU_HUFFMAN: *value = ctx->opt_gpu_huffman; return P2P_OK;
        case P2P_OPT_GPU_HUFFMAN_COUNT: *value = (int)ctx->gpu_huffman_used; return P2P_OK;

        default: return fail(ctx, P2P_ERR_INVALID, "unknown option");
    }
}

// ---- host helpers --------------------------------------------------------------------------
int p2p_pitch_constants(double fov_deg, double pitch_deg, int W, p2p_pitch_consts *out) {
    if (!out || W <= 0) return P2P_ERR_INVALID;
    const double deg = M_PI / 180.0;  // np.radians(x) == x * (pi / 180)
    const double fov = fov_deg * deg, p = pitch_deg * deg;
    out->f = (float)((0.5 * (double)W) / tan(fov / 2.0));
    out->c = (float)cos(p);
    out->s = (float)sin(p);
    return P2P_OK;
}

int p2p_yaw_table(int Wp, double yaw_deg, int32_t *ix, int32_t *fx, int32_t *shift) {
    if (Wp <= 0 || !ix || !fx) return P2P_ERR_INVALID;
    const double two_pi = 2.0 * M_PI;
    const double yaw = yaw_deg * (M_PI / 180.0);  // f64 scalar: the sum below is f64 (NEP 50)
    const float two_pi_f = (float)two_pi;
    bool roll = true;
    for (int u = 0; u < Wp; ++u) {
        // ref :95   phi = (2*pi*u / Wp) in f32: f32(2pi) * f32(u), then / f32(Wp)
        volatile float m = two_pi_f * (float)u;
        const float phi = m / (float)Wp;
        // ref :98   (phi + yaw) % 2pi in f64, Python-style remainder (result takes divisor's sign)
        double r = fmod((double)phi + yaw, two_pi);
        if (r != 0.0) {
            if (r < 0.0) r += two_pi;
        } else {
            r = 0.0;
        }
        // ref :101-105
        double U = (r * (double)Wp) / two_pi;
        if (U < 0.0) U = 0.0;
        if (U > (double)(Wp - 1)) U = (double)(Wp - 1);
        const float Uf = (float)U;
        volatile float sc = Uf * 32.0f;
        const int sx = (int)lrintf(sc);  // cvRound: round half even
        ix[u] = sx >> 5;
        fx[u] = sx & 31;
        if (fx[u] != 0) roll = false;
    }
    if (roll) {
        const int s0 = ix[0];
        for (int u = 0; u < Wp && roll; ++u) roll = (ix[u] == (u + s0) % Wp);
        if (shift) *shift = roll ? s0 : -1;
    } else if (shift) {
        *shift = -1;
    }
    return P2P_OK;
}

int p2p_host_alloc(void **ptr, size_t bytes) {
    if (!ptr || bytes == 0) return P2P_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return P2P_ERR_NOMEM;
    }
    return P2P_OK;
}

int p2p_host_free(void *ptr) {
    if (!ptr) return P2P_OK;
    if (cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return P2P_ERR_CUDA;
    }
    return P2P_OK;
}

int p2p_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return P2P_ERR_INVALID;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return P2P_ERR_CUDA;
    }
    return P2P_OK;
}

int p2p_host_unregister(void *ptr) {
    if (!ptr) return P2P_ERR_INVALID;
    if (cudaHostUnregister(ptr) != cudaSuccess) {
        cudaGetLastError();
        return P2P_ERR_CUDA;
    }
    return P2P_OK;
}

// ---- panorama upload -----------------------------------------------------------------------
int p2p_upload_pano(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride) {
    P2P_NVTX("p2p_upload_pano");
    if (!slot_ok(ctx, slot) || !bgr) return fail(ctx, P2P_ERR_INVALID, "bad slot or null panorama");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return upload_rows(ctx, slot, bgr, Wp, Hp, row_stride, 0, Hp);
}

int p2p_upload_pano_device(p2p_ctx *ctx, int slot, const void *d_bgr, int Wp, int Hp, size_t row_stride) {
    P2P_NVTX("p2p_upload_pano_device");
    if (!slot_ok(ctx, slot) || !d_bgr) return fail(ctx, P2P_ERR_INVALID, "bad slot or null panorama");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_dims(ctx, Wp, Hp);
    if (rc) return rc;
    if (row_stride < (size_t)Wp * 3) return fail(ctx, P2P_ERR_INVALID, "row_stride smaller than Wp * 3");
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[slot];
    rc = prepare_slot(ctx, s, Wp, Hp);
    if (rc) return rc;
    return launch_pack(ctx, s, static_cast<const uint8_t *>(d_bgr), row_stride, 0, Hp);
}

int p2p_rotate_pano(p2p_ctx *ctx, int src_slot, int dst_slot, const int32_t *ix, const int32_t *fx) {
    P2P_NVTX("p2p_rotate_pano");
    if (!slot_ok(ctx, src_slot) || !slot_ok(ctx, dst_slot) || src_slot == dst_slot || !ix || !fx)
        return fail(ctx, P2P_ERR_INVALID, "bad slots or null table");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Slot &a = ctx->slots[src_slot];
    Slot &d = ctx->slots[dst_slot];
    if (!a.valid) return fail(ctx, P2P_ERR_STATE, "source slot holds no panorama");
    if (slot_is_partial(a)) return fail(ctx, P2P_ERR_STATE, "source slot holds a partial panorama (p2p_process_image)");
    if (ctx->opt_interp != 0)
        return fail(ctx, P2P_ERR_INVALID, "fractional yaws are only defined for the cv2 fixed-point interpolation mode");
    for (int u = 0; u < a.Wp; ++u)
        if (ix[u] < 0 || ix[u] >= a.Wp || fx[u] < 0 || fx[u] > 31) return fail(ctx, P2P_ERR_INVALID, "yaw table entry out of range");
    int rc = prepare_slot(ctx, d, a.Wp, a.Hp);
    if (rc) return rc;
    rc = ensure(ctx, &d.d_tab, &d.tab_cap, (size_t)a.Wp * 2 * sizeof(int32_t));
    if (rc) return rc;
    // the table is tiny; the source must be complete before the destination stream reads it
    CK(cudaStreamSynchronize(a.stream));
    CK(cudaMemcpyAsync(d.d_tab, ix, (size_t)a.Wp * 4, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.d_tab + a.Wp, fx, (size_t)a.Wp * 4, cudaMemcpyHostToDevice, d.stream));
    CK(cudaStreamSynchronize(d.stream));  // ix / fx are caller memory (pageable): copy must be done
    cudaSurfaceObject_t surf = 0;
    if (ctx->opt_sampler == 1) {
        rc = ensure_array(ctx, d);
        if (rc) return rc;
        surf = d.surf;
    }
    dim3 block(256), grid((a.Wp + 1 + 255) / 256, a.Hp + 1);
    rotate_kernel<<<grid, block, 0, d.stream>>>(a.d_rgba, d.d_rgba, a.pitch_tex, a.Wp, a.Hp, d.d_tab, d.d_tab + a.Wp, surf);
    ctx->launches++;
    CK(cudaGetLastError());
    d.valid = true;
    d.row0 = 0;
    d.row1 = a.Hp;
    d.tex_current = (surf != 0);
    return P2P_OK;
}

// ---- hot path ------------------------------------------------------------------------------
int p2p_project_views(p2p_ctx *ctx, int slot, int n_yaw, const int32_t *yaw_shift, int n_pitch,
                      const p2p_pitch_consts *pitch, int W, int H, uint8_t *out, int out_on_device) {
    P2P_NVTX("p2p_project_views");
    if (!ctx) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return project_views_locked(ctx, slot, n_yaw, yaw_shift, n_pitch, pitch, W, H, out, out_on_device);
}

int p2p_project_batch(p2p_ctx *ctx, int n_images, const int32_t *slots, int n_yaw, const int32_t *yaw_shift,
                      int n_pitch, const p2p_pitch_consts *pitch, int W, int H, uint8_t *const *outs,
                      int out_on_device) {
    P2P_NVTX("p2p_project_batch");
    if (!ctx || n_images <= 0 || !slots || !outs) return fail(ctx, P2P_ERR_INVALID, "bad batch arguments");
    int nb = 1;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        nb = ctx->opt_nb;
    }
    int i = 0;
    while (i < n_images) {
        // images that share one launch must be resident, equally sized and write to device memory
        int g = 1;
        if (out_on_device && nb > 1 && i + nb <= n_images && (W & 3) == 0) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            bool ok = true;
            for (int b = 0; b < nb && ok; ++b) {
                ok = slot_ok(ctx, slots[i + b]) && ctx->slots[slots[i + b]].valid && outs[i + b] &&
                     (reinterpret_cast<uintptr_t>(outs[i + b]) & 3) == 0 &&
                     ctx->slots[slots[i + b]].Wp == ctx->slots[slots[i]].Wp &&
                     ctx->slots[slots[i + b]].Hp == ctx->slots[slots[i]].Hp;
                for (int c = 0; c < b && ok; ++c) ok = slots[i + c] != slots[i + b];
            }
            if (ok) g = nb;
        }
        if (g == 1) {
            int rc = p2p_project_views(ctx, slots[i], n_yaw, yaw_shift, n_pitch, pitch, W, H, outs[i], out_on_device);
            if (rc) return rc;
        } else {
            std::lock_guard<std::mutex> lk(ctx->mu);
            int rc = check_project_args(ctx, slots[i], n_yaw, yaw_shift, n_pitch, pitch, W, H, outs[i],
                                        ctx->slots[slots[i]].Wp);
            if (rc) return rc;
            CK(cudaSetDevice(ctx->device));
            Slot *sl[kMaxImagesPerLaunch];
            uint8_t *o[kMaxImagesPerLaunch];
            for (int b = 0; b < g; ++b) {
                sl[b] = &ctx->slots[slots[i + b]];
                o[b] = outs[i + b];
                // the launch runs on the first slot's stream: the others must have finished uploading
                if (b > 0 && sl[b]->stream != sl[0]->stream) CK(cudaStreamSynchronize(sl[b]->stream));
            }
            rc = launch_project(ctx, sl, g, n_yaw, yaw_shift, n_pitch, pitch, W, H, o);
            if (rc) return rc;
        }
        i += g;
    }
    return P2P_OK;
}

// Flat view list, optional row band: view i = (yaw_shift[i], pitch[i]) -> out + i * W * H * 3, rows row_begin .. row_end - 1.
int p2p_project_view_list(p2p_ctx *ctx, int slot, int n_views, const int32_t *yaw_shift, const p2p_pitch_consts *pitch,
                          int W, int H, int row_begin, int row_end, uint8_t *out, int out_on_device) {
    P2P_NVTX("p2p_project_view_list");
    if (!ctx) return P2P_ERR_INVALID;
    if (!slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
    if (n_views <= 0 || !yaw_shift || !pitch || !out) return fail(ctx, P2P_ERR_INVALID, "null or empty view list / output");
    if (W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "output size must be positive");
    if (W >= 32767 || H >= 32767) return fail(ctx, P2P_ERR_LIMIT, "output dimension >= 32767");
    if (row_begin < 0 || row_end > H || row_begin > row_end) return fail(ctx, P2P_ERR_INVALID, "row band outside [0, H]");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->slots[slot];
    if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
    for (int i = 0; i < n_views; ++i)
        if (yaw_shift[i] < 0 || yaw_shift[i] >= s.Wp) return fail(ctx, P2P_ERR_INVALID, "yaw_shift outside [0, Wp)");
    if (row_begin == row_end) return P2P_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t view_bytes = (size_t)W * H * 3;
    uint8_t *d_out = out;
    int rc;
    if (!out_on_device) {
        rc = ensure(ctx, &s.d_out, &s.out_cap, view_bytes * n_views);
        if (rc) return rc;
        d_out = s.d_out;
    }
    if (slot_is_partial(s)) {  // a slot filled by p2p_process_image holds only the rows its own views touch
        if (ctx->opt_interp != 0)
            return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama: exact interpolation needs a full upload");
        std::vector<p2p_pitch_consts> uniq;
        for (int i = 0; i < n_views; ++i) {
            bool seen = false;
            for (const p2p_pitch_consts &u : uniq) seen = seen || memcmp(&u, &pitch[i], sizeof(u)) == 0;
            if (!seen) uniq.push_back(pitch[i]);
        }
        int lo = 0, hi = 0;
        rc = view_row_range(ctx, s.stream, (int)uniq.size(), uniq.data(), W, H, s.Wp, s.Hp, &lo, &hi);
        if (rc) return rc;
        if (lo < s.row0 || hi + 1 > s.row1)
            return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama that does not cover these views: upload it again");
    }
    if (rows_kernel_usable(ctx, W, H, n_views, d_out)) {
        std::vector<int> oi((size_t)n_views);
        for (int i = 0; i < n_views; ++i) oi[i] = i;
        rc = launch_rows(ctx, s, n_views, yaw_shift, pitch, oi.data(), W, H, d_out, row_begin, row_end);
        if (rc) return rc;
    } else {
        // geometries the row-segment kernel does not take (W % 8 != 0, LDG sampler, exact-bilinear mode ...): whole views,
        // one generic launch each; the band is cut out by the copy below
        Slot *sl[1] = {&s};
        for (int i = 0; i < n_views; ++i) {
            uint8_t *o[1] = {d_out + (size_t)i * view_bytes};
            rc = launch_project(ctx, sl, 1, 1, &yaw_shift[i], 1, &pitch[i], W, H, o);
            if (rc) return rc;
        }
    }
    if (!out_on_device) {
        const size_t off = (size_t)row_begin * W * 3, width = (size_t)(row_end - row_begin) * W * 3;
        CK(cudaMemcpy2DAsync(out + off, view_bytes, d_out + off, view_bytes, width, (size_t)n_views,
                             cudaMemcpyDeviceToHost, s.stream));
    }
    return P2P_OK;
}

// Packed rows [y0, y1] (inclusive, y1 <= Hp: row Hp is the clamp row) now hold data in slot `s` of size Wp x Hp: merge them
// with the rows it held before (`was_valid`, old range) when both ranges touch - a panorama can be assembled from pieces
// (p2p_upload_pano_rows, p2p_copy_pano_rows) - else the slot holds just the new piece.
static std::atomic<bool> peer_tried[64][64];   // peer access of (destination device, source device) has been requested

static void merge_rows(Slot &s, bool was_valid, int old_Wp, int old_Hp, int old0, int old1, int y0, int y1) {
    if (was_valid && old_Wp == s.Wp && old_Hp == s.Hp && y0 <= old1 + 1 && y1 >= old0 - 1) {
        s.row0 = (old0 < y0) ? old0 : y0;
        s.row1 = (old1 > y1) ? old1 : y1;
    } else {
        s.row0 = y0;
        s.row1 = y1;
    }
    s.valid = true;
}

int p2p_upload_pano_rows(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride, int row_begin,
                         int row_end) {
    P2P_NVTX("p2p_upload_pano_rows");
    if (!slot_ok(ctx, slot) || !bgr) return fail(ctx, P2P_ERR_INVALID, "bad slot or null panorama");
    if (row_begin < 0 || row_end > Hp || row_begin >= row_end) return fail(ctx, P2P_ERR_INVALID, "row range outside [0, Hp]");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->slots[slot];
    const bool was_valid = s.valid;
    const int oW = s.Wp, oH = s.Hp, o0 = s.row0, o1 = s.row1;
    const size_t old_cap = s.rgba_cap;
    const uint32_t *old_ptr = s.d_rgba;
    // the last piece also writes the clamp row Hp (a copy of row Hp - 1)
    const int y1 = (row_end == Hp) ? Hp : row_end - 1;
    int rc = upload_rows(ctx, slot, bgr, Wp, Hp, row_stride, row_begin, y1);
    if (rc) return rc;
    merge_rows(s, was_valid && old_ptr == s.d_rgba && old_cap == s.rgba_cap, oW, oH, o0, o1, row_begin, y1);
    return P2P_OK;
}

// Copy packed rows of src's slot into dst's slot (another device of the same box: cudaMemcpyPeerAsync, NVLink when peer
// access is available).  row_begin < 0: every row the source holds.  Asynchronous on the destination slot's stream,
// ordered after everything enqueued on the source slot so far.
int p2p_copy_pano_rows(p2p_ctx *dst, int dst_slot, p2p_ctx *src, int src_slot, int row_begin, int row_end) {
    P2P_NVTX("p2p_copy_pano_rows");
    if (!dst || !src) return P2P_ERR_INVALID;
    if (!slot_ok(dst, dst_slot) || !slot_ok(src, src_slot)) return fail(dst, P2P_ERR_INVALID, "bad slot");
    if (dst == src && dst_slot == src_slot) return fail(dst, P2P_ERR_INVALID, "source and destination are the same slot");
    std::unique_lock<std::mutex> l1(dst->mu, std::defer_lock), l2(src->mu, std::defer_lock);
    if (dst == src) l1.lock(); else std::lock(l1, l2);
    p2p_ctx *ctx = dst;  // CK reports on the destination context
    Slot &a = src->slots[src_slot];
    Slot &d = dst->slots[dst_slot];
    if (!a.valid) return fail(dst, P2P_ERR_STATE, "source slot holds no panorama");
    int y0 = a.row0, y1 = a.row1;
    if (row_begin >= 0) {
        y0 = row_begin;
        y1 = row_end - 1;   // half-open [row_begin, row_end) over the packed rows 0 .. Hp (Hp = the clamp row)
        if (y0 > y1 || y0 < a.row0 || y1 > a.row1) return fail(dst, P2P_ERR_STATE, "source slot does not hold these rows");
    }
    const bool was_valid = d.valid;
    const int oW = d.Wp, oH = d.Hp, o0 = d.row0, o1 = d.row1;
    const size_t old_cap = d.rgba_cap;
    const uint32_t *old_ptr = d.d_rgba;
    cudaEvent_t ev = nullptr;
    CK(cudaSetDevice(src->device));
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, a.stream);
    if (e == cudaSuccess) e = cudaSetDevice(dst->device);
    if (e == cudaSuccess && dst->device != src->device && dst->device < 64 && src->device < 64 &&
        !peer_tried[dst->device][src->device].exchange(true)) {   // once per device pair and process
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dst->device, src->device) == cudaSuccess && can) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);  // direct NVLink path; staged through the host otherwise
            if (pe != cudaSuccess) cudaGetLastError();                    // already enabled / not supported: the copy still works
        }
    }
    int rc = P2P_OK;
    if (e == cudaSuccess) rc = prepare_slot(dst, d, a.Wp, a.Hp);
    if (e == cudaSuccess && rc == P2P_OK) e = cudaStreamWaitEvent(d.stream, ev, 0);
    if (e == cudaSuccess && rc == P2P_OK) {
        const size_t row_bytes = (size_t)a.pitch_tex * 4;
        const size_t off = (size_t)y0 * row_bytes, bytes = (size_t)(y1 - y0 + 1) * row_bytes;
        e = cudaMemcpyPeerAsync(reinterpret_cast<uint8_t *>(d.d_rgba) + off, dst->device,
                                reinterpret_cast<const uint8_t *>(a.d_rgba) + off, src->device, bytes, d.stream);
    }
    cudaEventDestroy(ev);  // deferred by the runtime until the wait has consumed it
    if (rc) return rc;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(dst, P2P_ERR_CUDA, "p2p_copy_pano_rows", e);
    }
    merge_rows(d, was_valid && old_ptr == d.d_rgba && old_cap == d.rgba_cap, oW, oH, o0, o1, y0, y1);
    d.tex_current = false;  // the gather array of the destination is refreshed from the linear copy before its next launch
    return P2P_OK;
}

int p2p_copy_pano(p2p_ctx *dst, int dst_slot, p2p_ctx *src, int src_slot) {
    if (dst && slot_ok(dst, dst_slot)) {   // a replica replaces whatever the destination held
        std::lock_guard<std::mutex> lk(dst->mu);
        dst->slots[dst_slot].valid = false;
    }
    return p2p_copy_pano_rows(dst, dst_slot, src, src_slot, -1, -1);
}

int p2p_process_image(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride,
                      int n_yaw, const int32_t *yaw_shift, int n_pitch, const p2p_pitch_consts *pitch,
                      int W, int H, uint8_t *out_host) {
    P2P_NVTX("p2p_process_image");
    if (!slot_ok(ctx, slot) || !bgr) return fail(ctx, P2P_ERR_INVALID, "bad slot or null panorama");
    if (n_pitch <= 0 || !pitch || W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "null or empty view list / output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_dims(ctx, Wp, Hp);
    if (rc) return rc;
    // the views are known before the transfer: move only the panorama rows they can touch (the pitch map does
    // not depend on the yaw or the image, so the range is memoised per geometry like the reference's map cache)
    int y0 = 0, y1 = Hp;
    if (ctx->opt_partial && ctx->opt_interp == 0) {
        CK(cudaSetDevice(ctx->device));
        int lo = 0, hi = 0;
        rc = view_row_range(ctx, ctx->slots[slot].stream, n_pitch, pitch, W, H, Wp, Hp, &lo, &hi);
        if (rc) return rc;
        y0 = lo;
        y1 = hi + 1;  // second tap row; Hp = the clamp row (a copy of row Hp - 1)
    }
    rc = upload_rows(ctx, slot, bgr, Wp, Hp, row_stride, y0, y1);
    if (rc) return rc;
    return project_views_locked(ctx, slot, n_yaw, yaw_shift, n_pitch, pitch, W, H, out_host, 0);
}

int p2p_view_row_range(p2p_ctx *ctx, int n_pitch, const p2p_pitch_consts *pitch, int W, int H, int Wp, int Hp,
                       int *first_row, int *last_row) {
    if (!ctx || n_pitch <= 0 || !pitch || W <= 0 || H <= 0 || !first_row || !last_row)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_dims(ctx, Wp, Hp);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    int lo = 0, hi = 0;
    rc = view_row_range(ctx, ctx->slots[0].stream, n_pitch, pitch, W, H, Wp, Hp, &lo, &hi);
    if (rc) return rc;
    *first_row = lo;
    *last_row = (hi + 1 < Hp) ? hi + 1 : Hp - 1;
    return P2P_OK;
}

// ---- JPEG files of the views (the encode side of cv2.imwrite, ref :277) -------------------------------
int p2p_encode_jpeg(p2p_ctx *ctx, int slot, const uint8_t *bgr, int on_device, int n_images, int W, int H, int quality,
                    uint8_t *out_host, size_t out_stride, size_t *sizes) {
    P2P_NVTX("p2p_encode_jpeg");
    if (!slot_ok(ctx, slot) || !bgr || n_images <= 0 || W <= 0 || H <= 0 || !out_host || !sizes)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    p2pjpeg::Geometry G;
    Slot &s = ctx->slots[slot];
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        const uint8_t *d_src = bgr;
        if (!on_device) {
            const size_t bytes = (size_t)n_images * W * H * 3;
            int rc = ensure(ctx, &s.d_out, &s.out_cap, bytes);
            if (rc) return rc;
            CK(cudaMemcpyAsync(s.d_out, bgr, bytes, cudaMemcpyHostToDevice, s.stream));
            d_src = s.d_out;
        }
        int rc = enqueue_jpeg(ctx, s, d_src, n_images, W, H, quality, G);
        if (rc) return rc;
    }
    cudaSetDevice(ctx->device);
    int rc = collect_jpeg(ctx, s, n_images, G, out_host, out_stride, sizes);
    if (rc == P2P_ERR_LIMIT) return fail(ctx, rc, "a JPEG file does not fit its output buffer (out_stride) or the encoder's capacity");
    if (rc) return fail(ctx, rc, "JPEG encoder: CUDA error");
    return P2P_OK;
}

int p2p_project_views_jpeg(p2p_ctx *ctx, int slot, int n_yaw, const int32_t *yaw_shift, int n_pitch,
                           const p2p_pitch_consts *pitch, int W, int H, int quality, uint8_t *out_host,
                           size_t out_stride, size_t *sizes) {
    P2P_NVTX("p2p_project_views_jpeg");
    return p2p_process_image_jpeg(ctx, slot, nullptr, 0, 0, 0, n_yaw, yaw_shift, n_pitch, pitch, W, H, quality, out_host,
                                  out_stride, sizes);
}

int p2p_process_image_jpeg(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride, int n_yaw,
                           const int32_t *yaw_shift, int n_pitch, const p2p_pitch_consts *pitch, int W, int H,
                           int quality, uint8_t *out_host, size_t out_stride, size_t *sizes) {
    P2P_NVTX("p2p_process_image_jpeg");
    if (!slot_ok(ctx, slot) || !out_host || !sizes) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    if (n_yaw <= 0 || n_pitch <= 0 || !pitch || W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "null or empty view list / output");
    p2pjpeg::Geometry G;
    Slot &s = ctx->slots[slot];
    const int n = n_yaw * n_pitch;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        int rc = P2P_OK;
        if (bgr) {  // upload first: only the rows these views can touch (see p2p_process_image)
            rc = check_dims(ctx, Wp, Hp);
            if (rc) return rc;
            int y0 = 0, y1 = Hp;
            if (ctx->opt_partial && ctx->opt_interp == 0) {
                int lo = 0, hi = 0;
                rc = view_row_range(ctx, s.stream, n_pitch, pitch, W, H, Wp, Hp, &lo, &hi);
                if (rc) return rc;
                y0 = lo;
                y1 = hi + 1;
            }
            rc = upload_rows(ctx, slot, bgr, Wp, Hp, row_stride, y0, y1);
            if (rc) return rc;
        }
        if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
        rc = check_project_args(ctx, slot, n_yaw, yaw_shift, n_pitch, pitch, W, H, out_host, s.Wp);
        if (rc) return rc;
        rc = ensure(ctx, &s.d_out, &s.out_cap, (size_t)n * W * H * 3);
        if (rc) return rc;
        Slot *sl[1] = {&s};
        uint8_t *outs[1] = {s.d_out};
        rc = launch_project(ctx, sl, 1, n_yaw, yaw_shift, n_pitch, pitch, W, H, outs);
        if (rc) return rc;
        rc = enqueue_jpeg(ctx, s, s.d_out, n, W, H, quality, G);
        if (rc) return rc;
    }
    cudaSetDevice(ctx->device);
    int rc = collect_jpeg(ctx, s, n, G, out_host, out_stride, sizes);
    if (rc == P2P_ERR_LIMIT) return fail(ctx, rc, "a JPEG file does not fit its output buffer (out_stride) or the encoder's capacity");
    if (rc) return fail(ctx, rc, "JPEG encoder: CUDA error");
    return P2P_OK;
}

// ---- JPEG panoramas decoded on the device (the decode side of cv2.imread, ref :244) -----------------------
int p2p_jpeg_probe(const uint8_t *file, size_t len, int *W, int *H) {
    if (!file || !W || !H) return P2P_ERR_INVALID;
    p2pjdec::Parsed P;
    if (p2pjdec::parse_headers(file, len, P)) return P2P_ERR_UNSUPPORTED;
    *W = P.info.W;
    *H = P.info.H;
    return P2P_OK;
}

int p2p_jpeg_coefficients(const uint8_t *file, size_t len, int16_t *coef, size_t capacity, int32_t *layout) {
    if (!file || !layout) return P2P_ERR_INVALID;
    p2pjdec::Parsed P;
    if (p2pjdec::parse_headers(file, len, P)) return P2P_ERR_UNSUPPORTED;
    const p2pjdec::Info &I = P.info;
    layout[0] = I.W; layout[1] = I.H; layout[2] = I.hmax; layout[3] = I.vmax;
    for (int k = 0; k < 3; ++k) {
        layout[4 + 2 * k] = I.bw[k];
        layout[5 + 2 * k] = I.bh[k];
    }
    if (!coef) return P2P_OK;
    if (capacity < I.n_coef) return P2P_ERR_INVALID;
    return p2pjdec::decode_scan(file, len, P, coef) ? P2P_ERR_UNSUPPORTED : P2P_OK;
}

int p2p_upload_pano_jpeg(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, int *Wp, int *Hp) {
    P2P_NVTX("p2p_upload_pano_jpeg");
    if (!slot_ok(ctx, slot) || !file || !Wp || !Hp) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    p2pjdec::Parsed P;
    size_t dstride = 0;
    int rc = decode_jpeg_to_staging(ctx, slot, file, len, P, &dstride);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        rc = prepare_slot(ctx, s, P.info.W, P.info.H);
        if (rc) return rc;
        *Wp = P.info.W;
        *Hp = P.info.H;
        rc = launch_pack(ctx, s, s.d_bgr, dstride, 0, P.info.H);
        if (rc) return rc;
    }
    // the IDCT reports coefficient blocks no 8-bit encoder produces (damaged data): wait for it outside the lock
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "JPEG decoder: CUDA error");
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->out_of_range)) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        s.valid = false;
        return fail(ctx, P2P_ERR_UNSUPPORTED, "JPEG data out of range (damaged file: fall back to cv2.imread)");
    }
    return P2P_OK;
}

int p2p_decode_jpeg(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, uint8_t *bgr_host, size_t row_stride,
                    size_t capacity_rows) {
    P2P_NVTX("p2p_decode_jpeg");
    if (!slot_ok(ctx, slot) || !file || !bgr_host) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    p2pjdec::Parsed P;
    size_t dstride = 0;
    int rc = decode_jpeg_to_staging(ctx, slot, file, len, P, &dstride);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
    if (rc) return rc;
    if (row_stride < (size_t)P.info.W * 3 || capacity_rows < (size_t)P.info.H)
        return fail(ctx, P2P_ERR_INVALID, "output buffer smaller than the image (see p2p_jpeg_probe)");
    Slot &s = ctx->slots[slot];
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        s.valid = false;  // the staging image changed under whatever panorama the slot held
        CK(cudaMemcpy2DAsync(bgr_host, row_stride, s.d_bgr, dstride, (size_t)P.info.W * 3, P.info.H, cudaMemcpyDeviceToHost,
                             s.stream));
    }
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "JPEG decoder: CUDA error");
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->out_of_range))
        return fail(ctx, P2P_ERR_UNSUPPORTED, "JPEG data out of range (damaged file: fall back to cv2.imread)");
    return P2P_OK;
}

// ---- PNG files of the views (cv2.imwrite(<name>.png, view), ref :277, the default output format) ----------------
int p2p_encode_png(p2p_ctx *ctx, int slot, const uint8_t *bgr, int on_device, int n_images, int W, int H,
                   uint8_t *out_host, size_t out_stride, size_t *sizes) {
    P2P_NVTX("p2p_encode_png");
    if (!slot_ok(ctx, slot) || !bgr || n_images <= 0 || W <= 0 || H <= 0 || !out_host || !sizes)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    p2ppng::Geom G;
    Slot &s = ctx->slots[slot];
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        const uint8_t *d_src = bgr;
        if (!on_device) {
            const size_t bytes = (size_t)n_images * W * H * 3;
            int rc = ensure(ctx, &s.d_out, &s.out_cap, bytes);
            if (rc) return rc;
            CK(cudaMemcpyAsync(s.d_out, bgr, bytes, cudaMemcpyHostToDevice, s.stream));
            d_src = s.d_out;
        }
        int rc = enqueue_png(ctx, s, d_src, n_images, W, H, G);
        if (rc) return rc;
    }
    cudaSetDevice(ctx->device);
    int rc = collect_png(s, n_images, G.out_cap, out_host, out_stride, sizes);
    if (rc == P2P_ERR_LIMIT) return fail(ctx, rc, "a PNG file does not fit its output buffer (out_stride)");
    if (rc) return fail(ctx, rc, "PNG encoder: CUDA error");
    return P2P_OK;
}

int p2p_process_image_png(p2p_ctx *ctx, int slot, const uint8_t *bgr, int Wp, int Hp, size_t row_stride, int n_yaw,
                          const int32_t *yaw_shift, int n_pitch, const p2p_pitch_consts *pitch, int W, int H,
                          uint8_t *out_host, size_t out_stride, size_t *sizes, uint8_t *pixels_host) {
    P2P_NVTX("p2p_process_image_png");
    if (!slot_ok(ctx, slot) || !out_host || !sizes) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    if (n_yaw <= 0 || n_pitch <= 0 || !pitch || W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "null or empty view list / output");
    p2ppng::Geom G;
    Slot &s = ctx->slots[slot];
    const int n = n_yaw * n_pitch;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        int rc = P2P_OK;
        if (bgr) {
            rc = check_dims(ctx, Wp, Hp);
            if (rc) return rc;
            int y0 = 0, y1 = Hp;
            if (ctx->opt_partial && ctx->opt_interp == 0) {
                int lo = 0, hi = 0;
                rc = view_row_range(ctx, s.stream, n_pitch, pitch, W, H, Wp, Hp, &lo, &hi);
                if (rc) return rc;
                y0 = lo;
                y1 = hi + 1;
            }
            rc = upload_rows(ctx, slot, bgr, Wp, Hp, row_stride, y0, y1);
            if (rc) return rc;
        }
        if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
        rc = check_project_args(ctx, slot, n_yaw, yaw_shift, n_pitch, pitch, W, H, out_host, s.Wp);
        if (rc) return rc;
        rc = ensure(ctx, &s.d_out, &s.out_cap, (size_t)n * W * H * 3);
        if (rc) return rc;
        Slot *sl[1] = {&s};
        uint8_t *outs[1] = {s.d_out};
        rc = launch_project(ctx, sl, 1, n_yaw, yaw_shift, n_pitch, pitch, W, H, outs);
        if (rc) return rc;
        rc = enqueue_png(ctx, s, s.d_out, n, W, H, G);
        if (rc) return rc;
        // the pixels too, if the caller wants them (needed for the views the device encoder does not handle)
        if (pixels_host) CK(cudaMemcpyAsync(pixels_host, s.d_out, (size_t)n * W * H * 3, cudaMemcpyDeviceToHost, s.stream));
    }
    cudaSetDevice(ctx->device);
    int rc = collect_png(s, n, G.out_cap, out_host, out_stride, sizes);
    if (rc == P2P_ERR_LIMIT) return fail(ctx, rc, "a PNG file does not fit its output buffer (out_stride)");
    if (rc) return fail(ctx, rc, "PNG encoder: CUDA error");
    return P2P_OK;
}

int p2p_sync(p2p_ctx *ctx, int slot) {
    if (!ctx) return P2P_ERR_INVALID;
    // the streams are read under the lock, the wait itself happens outside it: a thread waiting for its slot must
    // not keep the other threads (other slots) from enqueueing
    std::vector<cudaStream_t> streams;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (slot >= 0 && !slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
        for (int i = 0; i < ctx->n_slots; ++i)
            if (slot < 0 || i == slot) streams.push_back(ctx->slots[i].stream);
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    for (size_t i = 0; i < streams.size() && e == cudaSuccess; ++i) e = cudaStreamSynchronize(streams[i]);
    if (e != cudaSuccess) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        cudaGetLastError();
        return fail(ctx, P2P_ERR_CUDA, "cudaStreamSynchronize", e);
    }
    return P2P_OK;
}

int p2p_set_stream(p2p_ctx *ctx, int slot, void *cuda_stream) {
    if (!slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[slot];
    CK(cudaStreamSynchronize(s.stream));
    // the slot's own stream is kept (another slot may have borrowed it through p2p_get_stream)
    // and released by p2p_destroy
    if (s.own_stream) s.owned = s.stream;
    s.stream = static_cast<cudaStream_t>(cuda_stream);
    s.own_stream = false;
    return P2P_OK;
}

int p2p_get_stream(p2p_ctx *ctx, int slot, void **cuda_stream) {
    if (!slot_ok(ctx, slot) || !cuda_stream) return fail(ctx, P2P_ERR_INVALID, "bad slot or null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    *cuda_stream = ctx->slots[slot].stream;
    return P2P_OK;
}

// ---- timing --------------------------------------------------------------------------------
int p2p_event_create(p2p_ctx *ctx, void **event) {
    if (!ctx || !event) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    *event = e;
    return P2P_OK;
}

int p2p_event_destroy(p2p_ctx *ctx, void *event) {
    if (!ctx || !event) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
    return P2P_OK;
}

int p2p_event_record(p2p_ctx *ctx, void *event, int slot) {
    if (!slot_ok(ctx, slot) || !event) return fail(ctx, P2P_ERR_INVALID, "bad slot or event");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(static_cast<cudaEvent_t>(event), ctx->slots[slot].stream));
    return P2P_OK;
}

int p2p_event_wait(p2p_ctx *ctx, void *event, int slot) {
    if (!slot_ok(ctx, slot) || !event) return fail(ctx, P2P_ERR_INVALID, "bad slot or null event");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamWaitEvent(ctx->slots[slot].stream, static_cast<cudaEvent_t>(event), 0));
    return P2P_OK;
}

int p2p_event_elapsed_ms(p2p_ctx *ctx, void *start, void *stop, float *ms) {
    if (!ctx || !start || !stop || !ms) return P2P_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
    CK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
    return P2P_OK;
}

int p2p_flush_l2(p2p_ctx *ctx, int slot, size_t bytes) {
    if (!slot_ok(ctx, slot) || bytes == 0) return fail(ctx, P2P_ERR_INVALID, "bad slot or size");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    bytes = (bytes + 15) & ~(size_t)15;
    int rc = ensure(ctx, &ctx->d_flush, &ctx->flush_cap, bytes);
    if (rc) return rc;
    fill_kernel<<<148 * 8, 256, 0, ctx->slots[slot].stream>>>(ctx->d_flush, bytes / 16, (uint32_t)ctx->launches);
    ctx->launches++;
    CK(cudaGetLastError());
    return P2P_OK;
}

// ---- debug exports -------------------------------------------------------------------------
int p2p_selftest(p2p_ctx *ctx, const p2p_pitch_consts *pitch, int W, int H, int exhaustive_div,
                 unsigned long long *ray_mismatches, unsigned long long *div_mismatches) {
    if (!ctx || !pitch || W <= 0 || H <= 0 || !ray_mismatches || !div_mismatches)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaStream_t st = ctx->slots[0].stream;
    cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), st);
    if (e == cudaSuccess) {
        PitchC k{pitch->f, pitch->c, pitch->s};
        dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
        selftest_ray_kernel<<<grid, block, 0, st>>>(k, W, H, (float)(W / 2.0), (float)(H / 2.0), d);
        ctx->launches++;
        if (exhaustive_div) {
            selftest_constdiv_kernel<<<148 * 16, 256, 0, st>>>(d);
            ctx->launches++;
        }
        e = cudaGetLastError();
    }
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "p2p_selftest", e);
    *ray_mismatches = h[0];
    *div_mismatches = h[1];
    return P2P_OK;
}

int p2p_coords(p2p_ctx *ctx, const p2p_pitch_consts *pitch, int W, int H, int Wp, int Hp,
               float *U_host, float *V_host) {
    if (!ctx || !pitch || !U_host || !V_host || W <= 0 || H <= 0) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_dims(ctx, Wp, Hp);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * H;
    float *d = nullptr;
    CK(cudaMalloc(&d, n * 2 * sizeof(float)));
    PitchC k{pitch->f, pitch->c, pitch->s};
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    cudaStream_t st = ctx->slots[0].stream;
    coords_kernel<<<grid, block, 0, st>>>(k, W, H, (float)(W / 2.0), (float)(H / 2.0), (float)Wp, (float)Hp,
                                          (float)(Wp - 1), (float)(Hp - 1), d, d + n, ctx->opt_trig == 0);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(U_host, d, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(V_host, d + n, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "p2p_coords", e);
    return P2P_OK;
}

int p2p_sample_with_maps(p2p_ctx *ctx, int slot, int yaw_shift, const float *U_host, const float *V_host,
                         int W, int H, uint8_t *out_host) {
    if (!slot_ok(ctx, slot) || !U_host || !V_host || !out_host || W <= 0 || H <= 0)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->slots[slot];
    if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
    if (yaw_shift < 0 || yaw_shift >= s.Wp) return fail(ctx, P2P_ERR_INVALID, "yaw_shift outside [0, Wp)");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * H;
    float *d = nullptr;
    uint8_t *d_o = nullptr;
    CK(cudaMalloc(&d, n * 2 * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_o, n * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, U_host, n * sizeof(float), cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, V_host, n * sizeof(float), cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) {
        dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
        sample_maps_kernel<<<grid, block, 0, s.stream>>>(s.d_rgba, s.pitch_tex, s.Wp, s.Hp, yaw_shift, d, d + n, W, H, d_o,
                                                        ctx->opt_interp, ctx->opt_interp == 1 && ctx->opt_seam_wrap);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_o, n * 3, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d);
    cudaFree(d_o);
    if (e != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "p2p_sample_with_maps", e);
    return P2P_OK;
}

int p2p_download_pano(p2p_ctx *ctx, int slot, uint8_t *bgr_host, size_t row_stride) {
    if (!slot_ok(ctx, slot) || !bgr_host) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->slots[slot];
    if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
    if (slot_is_partial(s)) return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama (p2p_process_image)");
    if (row_stride < (size_t)s.Wp * 3) return fail(ctx, P2P_ERR_INVALID, "row_stride smaller than Wp * 3");
    CK(cudaSetDevice(ctx->device));
    uint8_t *d = nullptr;
    const size_t dstride = (size_t)s.Wp * 3;
    CK(cudaMalloc(&d, dstride * s.Hp));
    dim3 block(256), grid((s.Wp + 255) / 256, s.Hp);
    unpack_kernel<<<grid, block, 0, s.stream>>>(s.d_rgba, s.pitch_tex, s.Wp, s.Hp, d, dstride);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(bgr_host, row_stride, d, dstride, dstride, s.Hp, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "p2p_download_pano", e);
    return P2P_OK;
}

}  // extern "C"
