// p2p_api_jpeg.inl - host side of the JPEG encoder (csrc/p2p_jpeg.cuh) and its entry points
// Part of the single translation unit p2p_api.cu (textual include, after p2p_ctx.cuh).

namespace {

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
    cudaError_t e = wait_slot(ctx, s);
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
    e = wait_slot(ctx, s);
    if (e != cudaSuccess) return P2P_ERR_CUDA;
    (void)ctx;
    return rc;
}

}  // namespace

extern "C" {

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

}  // extern "C"
