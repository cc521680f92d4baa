// p2p_api_jpegdec.inl - host side of the JPEG decoder (csrc/p2p_jpegdec.cuh: destuffing, device Huffman stage, fallback) and its entry points
// Part of the single translation unit p2p_api.cu (textual include, after p2p_ctx.cuh).

namespace {

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
    // (the words are byte-swapped to MSB-first on the device, right after the copy: one pass less on the host)
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
        p2pjdec::bswap_words_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(s.jd_stream, (uint32_t)n_words);
        ctx->launches++;
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

}  // namespace

extern "C" {

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

}  // extern "C"
