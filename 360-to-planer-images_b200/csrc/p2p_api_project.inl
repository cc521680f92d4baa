// p2p_api_project.inl - host side of the projection path: packing, textures, kernel selection and launches, panorama upload / rotation / replication, the hot-path entry points of include/p2p.h
// Part of the single translation unit p2p_api.cu (textual include, after p2p_ctx.cuh).

namespace {

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
    if ((band + kRowsWarps - 1) / kRowsWarps > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
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
        dim3 grid((P.n_chunks + P.seg_chunks - 1) / P.seg_chunks, (band + kRowsWarps - 1) / kRowsWarps, ng);
        pick_rows(ctx->opt_trig == 0, full, ny_max)<<<grid, 32 * kRowsWarps, 0, s.stream>>>(P);
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

}  // namespace

extern "C" {

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

// Fractional yaws without the materialised yaw pass (project_frac_kernel): yaw k is its Wp-entry column table.
int p2p_project_views_table(p2p_ctx *ctx, int slot, int n_yaw, const int32_t *const *ix, const int32_t *const *fx, int n_pitch,
                            const p2p_pitch_consts *pitch, int W, int H, uint8_t *out, int out_on_device) {
    P2P_NVTX("p2p_project_views_table");
    if (!ctx) return P2P_ERR_INVALID;
    if (!slot_ok(ctx, slot)) return fail(ctx, P2P_ERR_INVALID, "bad slot");
    if (n_yaw <= 0 || n_pitch <= 0 || !ix || !fx || !pitch || !out || W <= 0 || H <= 0)
        return fail(ctx, P2P_ERR_INVALID, "bad argument");
    if (W >= 32767 || H >= 32767) return fail(ctx, P2P_ERR_LIMIT, "output dimension >= 32767");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->slots[slot];
    if (!s.valid) return fail(ctx, P2P_ERR_STATE, "slot holds no panorama");
    if (slot_is_partial(s)) return fail(ctx, P2P_ERR_STATE, "slot holds a partial panorama (p2p_process_image)");
    if (ctx->opt_interp != 0)
        return fail(ctx, P2P_ERR_INVALID, "fractional yaws are only defined for the cv2 fixed-point interpolation mode");
    const int Wp = s.Wp;
    if (Wp > 65535) return fail(ctx, P2P_ERR_LIMIT, "panorama too wide for the packed yaw table");
    // tab[k][c] = ix | fx << 16, entry Wp repeats entry 0
    std::vector<uint32_t> tabs((size_t)n_yaw * (Wp + 1));
    for (int k = 0; k < n_yaw; ++k) {
        if (!ix[k] || !fx[k]) return fail(ctx, P2P_ERR_INVALID, "null yaw table");
        uint32_t *t = tabs.data() + (size_t)k * (Wp + 1);
        for (int c = 0; c < Wp; ++c) {
            if (ix[k][c] < 0 || ix[k][c] >= Wp || fx[k][c] < 0 || fx[k][c] > 31)
                return fail(ctx, P2P_ERR_INVALID, "yaw table entry out of range");
            t[c] = (uint32_t)ix[k][c] | ((uint32_t)fx[k][c] << 16);
        }
        t[Wp] = t[0];
    }
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_texture(ctx, s);
    if (rc) return rc;
    rc = ensure(ctx, &s.d_tab, &s.tab_cap, tabs.size() * sizeof(uint32_t));
    if (rc) return rc;
    const size_t bytes = (size_t)n_yaw * n_pitch * W * H * 3;
    uint8_t *d_out = out;
    if (!out_on_device) {
        rc = ensure(ctx, &s.d_out, &s.out_cap, bytes);
        if (rc) return rc;
        d_out = s.d_out;
    }
    // (pageable source: the copy is staged before the call returns, and ordered behind earlier launches that read d_tab)
    CK(cudaMemcpyAsync(s.d_tab, tabs.data(), tabs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    ProjParams P;
    memset(&P, 0, sizeof(P));
    P.tex[0] = s.tex;
    P.out[0] = d_out;
    P.view_stride = (unsigned long long)W * H * 3;
    P.yaw_stride = P.view_stride * (unsigned long long)n_pitch;
    P.pitch_tex = s.pitch_tex;
    P.Wp = s.Wp; P.Hp = s.Hp; P.W = W; P.H = H;
    P.halfW = (float)(W / 2.0);
    P.halfH = (float)(H / 2.0);
    P.Wp_f = (float)s.Wp;
    P.Hp_f = (float)s.Hp;
    P.Umax = (float)(s.Wp - 1);
    P.Vmax = (float)(s.Hp - 1);
    P.inv_Wp = (float)(1.0 / (double)s.Wp);
    P.inv_Hp = (float)(1.0 / (double)s.Hp);
    P.numpy_trig = (ctx->opt_trig == 0);
    const bool quad = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0);
    for (int y0 = 0; y0 < n_yaw; y0 += 4) {
        const int ny_l = (n_yaw - y0 < 4) ? n_yaw - y0 : 4;
        FracTabs T;
        for (int k = 0; k < 4; ++k)
            T.tab[k] = reinterpret_cast<const uint32_t *>(s.d_tab) + (size_t)(y0 + (k < ny_l ? k : 0)) * (Wp + 1);
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
            dim3 grid((W + 31) / 32, (H + 7) / 8, np_l);
            if (grid.y > 65535) return fail(ctx, P2P_ERR_LIMIT, "output too large for one grid");
            switch (ny_l * 2 + (quad ? 1 : 0)) {
                case 2: project_frac_kernel<1, false><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 3: project_frac_kernel<1, true><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 4: project_frac_kernel<2, false><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 5: project_frac_kernel<2, true><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 6: project_frac_kernel<3, false><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 7: project_frac_kernel<3, true><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                case 8: project_frac_kernel<4, false><<<grid, kThreads, 0, s.stream>>>(P, T); break;
                default: project_frac_kernel<4, true><<<grid, kThreads, 0, s.stream>>>(P, T); break;
            }
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    if (!out_on_device) CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
    return P2P_OK;
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

}  // extern "C"
