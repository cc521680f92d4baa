// p2p_api_png.inl - host side of the PNG encoder (csrc/p2p_png.cuh) and its entry points
// Part of the single translation unit p2p_api.cu (textual include, after p2p_ctx.cuh).

namespace {

// CRC-32 byte table + the 4 byte tables of "advance the register by 256 zero bytes" (encoder), on the device once per context
int ensure_crc_table(p2p_ctx *ctx) {
    if (ctx->d_crc_table) return P2P_OK;
    uint32_t table[5 * 256];
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
    {
        const int rc_t = ensure_crc_table(ctx);
        if (rc_t) return rc_t;
    }
    const size_t npos = (size_t)n * G.Npad, nblk = (size_t)n * G.max_blk;
    const size_t tile_words = npos / kTile;   // per-tile arrays: last change, inherited run start, tokens, first token index
    const size_t blk_words = nblk * 2 + (((size_t)n + 3) & ~(size_t)3) + nblk * (kLCodes + 2) + 4 * tile_words;
    int rc = ensure(ctx, &s.pg_F, &s.pg_F_cap, npos);
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
    uint32_t *tile_last = lfreq + nblk * (kLCodes + 2), *tile_carry = tile_last + tile_words;
    uint32_t *tile_cnt = tile_carry + tile_words, *tile_first = tile_cnt + tile_words;
    unsigned long long *sums = s.pg_sums, *zbits = s.pg_sums + 2 * (size_t)n;
    cudaStream_t st = s.stream;
    if (H > 65535) return fail(ctx, P2P_ERR_LIMIT, "image too tall for one grid");
    CK(cudaMemsetAsync(s.pg_Z, 0, (size_t)n * G.z_cap, st));
    CK(cudaMemsetAsync(s.pg_sums, 0, 3 * (size_t)n * sizeof(unsigned long long), st));
    png_filter_kernel<<<dim3((W + 255) / 256, H, n), 256, 0, st>>>(d_bgr, s.pg_F, G);
    const dim3 tgrid(G.Npad / kTile, n);
    png_tile_kernel<<<tgrid, 256, 0, st>>>(s.pg_F, tile_last, sums, G);
    png_tile_scan_kernel<0><<<n, 1024, 0, st>>>(tile_last, tile_carry, nullptr, G);
    png_token_kernel<<<tgrid, 256, 0, st>>>(s.pg_F, tile_carry, s.pg_tlen, tile_cnt, G);
    png_tile_scan_kernel<1><<<n, 1024, 0, st>>>(tile_cnt, tile_first, ntok, G);
    png_blockpos_kernel<<<dim3((G.max_blk + 3) / 4, n), 128, 0, st>>>(s.pg_tlen, tile_first, ntok, blockpos, G);
    png_hist_kernel<<<dim3(G.max_blk, n), 256, 0, st>>>(s.pg_F, s.pg_tlen, blockpos, ntok, lfreq, G);
    png_tree_kernel<<<dim3((G.max_blk + kTreeWarps - 1) / kTreeWarps, n), kTreeWarps * 32, 0, st>>>(lfreq, ntok, blockpos, s.pg_info, G);
    png_layout_kernel<<<n, 32, 0, st>>>(s.pg_info, ntok, blkoff, zbits, G);
    png_emit_kernel<<<dim3(G.max_blk, n), 256, 0, st>>>(s.pg_F, s.pg_tlen, blockpos, ntok, s.pg_info, blkoff, zbits, s.pg_Z, G);
    const unsigned max_chunks = (unsigned)(G.z_cap / kIdat + 1);
    png_finish_kernel<<<dim3((max_chunks + kFinWarps - 1) / kFinWarps, n), kFinWarps * 32, 0, st>>>(s.pg_Z, zbits, sums, s.j_out,
                                                                                                ctx->d_crc_table, s.j_sizes_d, G);
    ctx->launches += 11;
    CK(cudaGetLastError());
    return P2P_OK;
}

// wait and copy the PNG files out; sizes[i] = 0 marks an image the device encoder does not handle
int collect_png(p2p_ctx *ctx, Slot &s, int n, size_t cap_per_image, uint8_t *out_host, size_t out_stride, size_t *sizes) {
    if (wait_slot(ctx, s) != cudaSuccess) return P2P_ERR_CUDA;
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
    if (wait_slot(ctx, s) != cudaSuccess) return P2P_ERR_CUDA;
    return rc;
}

}  // namespace

extern "C" {

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
    int rc = collect_png(ctx, s, n_images, G.out_cap, out_host, out_stride, sizes);
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
    int rc = collect_png(ctx, s, n, G.out_cap, out_host, out_stride, sizes);
    if (rc == P2P_ERR_LIMIT) return fail(ctx, rc, "a PNG file does not fit its output buffer (out_stride)");
    if (rc) return fail(ctx, rc, "PNG encoder: CUDA error");
    return P2P_OK;
}

}  // extern "C"
