// p2p_api_jpegdec.inl - host side of the JPEG decoder (csrc/p2p_jpegdec.cuh: destuffing, device Huffman stage, fallback) and its entry points
// Part of the single translation unit p2p_api.cu (textual include, after p2p_ctx.cuh).

namespace {

// ---- JPEG decoder (p2p_jpegdec.cuh) ------------------------------------------------------------
// flags the decoder's kernels and the host share: page-locked, mapped into the device's address space
int ensure_jd_flags(p2p_ctx *ctx, Slot &s) {
    if (s.jd_flags_h) return P2P_OK;
    CK(cudaHostAlloc(reinterpret_cast<void **>(&s.jd_flags_h), sizeof(*s.jd_flags_h), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(s.jd_flags_h, 0, sizeof(*s.jd_flags_h));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.jd_flags_d), s.jd_flags_h, 0));
    return P2P_OK;
}

// The Huffman stage in four parts, so that it can be queued either with a host check after every few rounds (modes 2 / 3)
// or optimistically, without waiting for anything but the destuffed length (mode 1, the default):
//   huff_begin          destuffing (device or calling thread), subsequence layout, tables, round 0
//   huff_rounds_checked rounds in batches of kRoundsPerCheck until one moves nothing (the host reads the flag)
//   huff_count_check    prefix sum of the blocks per subsequence + quota check per interval (+ the device-side gate)
//   huff_finish         write pass, DC sums, DC values
constexpr int kOptimisticRounds = 16;   // rounds queued after round 0 before the verdict is looked at (photographs: ~10)

void launch_round(Slot &s, int first, int *changed) {
    using namespace p2pjdec;
    Slot::JdRun &R = s.jd_run;
    if (R.fast)
        huff_sync_fast_kernel<<<R.sgrid, kSyncThreads, 0, s.stream>>>(s.jd_stream, s.jd_tables, R.d_lut, R.G, R.d_sub, s.jd_states,
                                                                     s.jd_states + R.G.n_sub, s.jd_nblk, first, changed);
    else
        huff_sync_kernel<<<R.sgrid, 128, 0, s.stream>>>(s.jd_stream, s.jd_tables, R.G, R.d_sub, s.jd_states,
                                                      s.jd_states + R.G.n_sub, s.jd_nblk, first, changed);
}

// Destuffing on the device: files without restart markers whose scan runs up to the file's final EOI.  The raw segment
// goes through the slot's pinned staging buffer; the kernels drop the stuffed zeros and raise `irregular` for anything
// else behind a 0xFF (markers, fill bytes), which sends the file down the calling-thread pass below.  The one thing the
// host needs back is the destuffed length.  Returns 0 = done (*n_out set, s.jd_stream filled MSB-first), 1 = not taken.
int destuff_on_device(p2p_ctx *ctx, Slot &s, const uint8_t *file, size_t len, const p2pjdec::Parsed &P, size_t *n_out) {
    using namespace p2pjdec;
    if (P.dri || len < P.ecs + 3 || file[len - 2] != 0xFF || file[len - 1] != 0xD9) return 1;
    const size_t n_raw = len - 2 - P.ecs;
    if (n_raw + 16 > s.jd_coef_h_cap || n_raw >= (1ull << 29)) return 1;
    uint8_t *stage = reinterpret_cast<uint8_t *>(s.jd_coef_h);
    memcpy(stage, file + P.ecs, n_raw);
    const unsigned chunks = (unsigned)((n_raw + kDestuffChunk - 1) / kDestuffChunk);
    const size_t chunks4 = ((size_t)chunks + 3) & ~(size_t)3;
    const size_t n_words = (n_raw + 3) / 4 + 3;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        int rc = ensure_grow(ctx, &s.jd_raw, &s.jd_raw_cap, n_raw + 16);
        if (!rc) rc = ensure_grow(ctx, &s.jd_stream, &s.jd_stream_cap, n_words * 4);
        if (!rc) rc = ensure_grow(ctx, &s.jd_dcnt, &s.jd_dcnt_cap, 2 * chunks4 * sizeof(uint32_t));
        if (rc) return rc;
        cudaStream_t st = s.stream;
        s.jd_flags_h->irregular = 0;
        s.jd_flags_h->total = 0;
        CK(cudaMemcpyAsync(s.jd_raw, stage, n_raw, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(s.jd_stream, 0, n_words * 4, st));
        destuff_count_kernel<<<chunks, 256, 0, st>>>(s.jd_raw, (uint32_t)n_raw, s.jd_dcnt, &s.jd_flags_d->irregular);
        p2pjpeg::jpeg_scan_kernel<<<1, 1024, 0, st>>>(s.jd_dcnt, s.jd_dcnt + chunks4, nullptr, chunks, chunks4, &s.jd_flags_d->total);
        destuff_scatter_kernel<<<chunks, 256, 0, st>>>(s.jd_raw, (uint32_t)n_raw, s.jd_dcnt + chunks4,
                                                       reinterpret_cast<uint8_t *>(s.jd_stream));
        ctx->launches += 3;
        CK(cudaGetLastError());
    }
    if (wait_slot(ctx, s) != cudaSuccess) return P2P_ERR_CUDA;
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->irregular)) return 1;
    const size_t removed = (size_t)*reinterpret_cast<volatile unsigned long long *>(&s.jd_flags_h->total);
    if (removed >= n_raw) return 1;
    *n_out = n_raw - removed;
    return 0;
}

// Returns P2P_OK (round 0 queued, s.jd_run describes the stage), an error, or 1 = "not handled".
int huff_begin(p2p_ctx *ctx, Slot &s, const uint8_t *file, size_t len, const p2pjdec::Parsed &P, int mode) {
    using namespace p2pjdec;
    const Info &I = P.info;
    const uint32_t nb = (I.ncomp == 1) ? 1u : (uint32_t)(I.hmax * I.vmax + 2);
    const uint32_t total_mcus = (uint32_t)I.mcux * I.mcuy;
    const uint32_t total_blocks = total_mcus * nb;
    const uint32_t ivl_mcus = P.dri ? (uint32_t)P.dri : total_mcus;
    const uint32_t n_ivl = (total_mcus + ivl_mcus - 1) / ivl_mcus;
    uint8_t *dst = reinterpret_cast<uint8_t *>(s.jd_coef_h);
    const size_t cap = s.jd_coef_h_cap;
    size_t n = 0;
    std::vector<uint32_t> ivl_byte(1, 0u);   // byte offset of every interval in the destuffed stream
    bool stream_on_device = false;
    if (mode == 1) {
        const int rc = destuff_on_device(ctx, s, file, len, P, &n);
        if (rc == 0) stream_on_device = true;
        else if (rc != 1) return rc;
    }
    if (!stream_on_device) {
        // destuff into the pinned staging buffer (FF 00 -> FF; RSTn starts the next interval, byte-aligned; any other
        // marker ends the scan); the words are swapped to MSB-first on the device so a 32-bit window is one funnel shift
        n = 0;
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
    if (!stream_on_device) memset(dst + n, 0, n_words * 4 - n);
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

    Slot::JdRun &R = s.jd_run;
    HuffGeom &G = R.G;
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
    R.nsub4 = ((size_t)G.n_sub + 3) & ~(size_t)3;
    R.sgrid = (G.n_sub + 127) / 128;
    R.pending = false;
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    R.fast = mode != 2;   // 2 = the plain rounds (tables in global memory), the tests' yardstick
    CK(cudaSetDevice(ctx->device));
    const size_t sub_bytes = ((size_t)G.n_sub * sizeof(SubSeq) + 15) & ~(size_t)15;
    int rc = ensure_grow(ctx, &s.jd_stream, &s.jd_stream_cap, n_words * 4);
    if (!rc) rc = ensure_grow(ctx, &s.jd_states, &s.jd_states_cap, 2 * (size_t)G.n_sub * sizeof(unsigned long long));
    if (!rc) rc = ensure_grow(ctx, &s.jd_nblk, &s.jd_nblk_cap, 2 * R.nsub4 * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.jd_dc, &s.jd_dc_cap, 2 * 3 * (size_t)G.dc_stride * sizeof(int32_t));
    if (!rc) rc = ensure(ctx, &s.jd_tiles, &s.jd_tiles_cap, 2 * 3 * ((((size_t)G.dc_stride + 4095) / 4096 + 3) & ~(size_t)3) * sizeof(uint32_t));
    if (!rc) rc = ensure(ctx, &s.jd_coef_d, &s.jd_coef_d_cap, I.n_coef * sizeof(int16_t));
    if (!rc) rc = ensure_grow(ctx, &s.jd_sub, &s.jd_sub_cap, sub_bytes + ((size_t)n_ivl + 1) * sizeof(uint32_t));
    if (rc) return rc;
    R.d_sub = reinterpret_cast<SubSeq *>(s.jd_sub);
    R.d_ivl_first = reinterpret_cast<uint32_t *>(s.jd_sub + sub_bytes);
    if (!s.jd_tables) CK(cudaMalloc(reinterpret_cast<void **>(&s.jd_tables), sizeof(DevTables)));
    if (!s.jd_tot_d) CK(cudaMalloc(reinterpret_cast<void **>(&s.jd_tot_d), 4 * sizeof(unsigned long long)));
    if (!s.jd_gate_d) CK(cudaMalloc(reinterpret_cast<void **>(&s.jd_gate_d), 8 * sizeof(int)));   // [0] gate, [4..6] blocks per component
    if (!stream_on_device) {
        CK(cudaMemcpyAsync(s.jd_stream, dst, n_words * 4, cudaMemcpyHostToDevice, st));
        bswap_words_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(s.jd_stream, (uint32_t)n_words);
        ctx->launches++;
    }
    // the tables below live in pageable memory: cudaMemcpyAsync stages them before it returns
    CK(cudaMemcpyAsync(s.jd_tables, &DT, sizeof(DevTables), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sub, subs.data(), (size_t)G.n_sub * sizeof(SubSeq), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_ivl_first, ivl_first.data(), ((size_t)n_ivl + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s.jd_coef_d, 0, I.n_coef * sizeof(int16_t), st));
    s.jd_flags_h->bad = 0;
    R.d_lut = reinterpret_cast<const SyncLut *>(reinterpret_cast<const unsigned char *>(s.jd_tables) + offsetof(DevTables, L));
    launch_round(s, 1, &s.jd_flags_d->changed[0]);
    ctx->launches++;
    CK(cudaGetLastError());
    return P2P_OK;
}

// synchronisation rounds, each batch followed by the "anything changed" flag coming back to the host.
// P2P_OK = fixed point reached, 1 = not within kMaxSyncRounds.  (Nothing of this slot is running when a batch is queued.)
int huff_rounds_checked(p2p_ctx *ctx, Slot &s) {
    using namespace p2pjdec;
    cudaStream_t st = s.stream;
    for (int round = 0; round < kMaxSyncRounds; round += kRoundsPerCheck) {
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            CK(cudaSetDevice(ctx->device));
            // several rounds per host check (a round in which nothing moves costs one early-exit pass); the flag that
            // decides convergence is the one of the LAST round of the batch
            for (int r = 0; r < kRoundsPerCheck; ++r) {
                s.jd_flags_h->changed[r] = 0;
                launch_round(s, 0, &s.jd_flags_d->changed[r]);
            }
            ctx->launches += kRoundsPerCheck;
            CK(cudaGetLastError());
        }
        if (wait_slot(ctx, s) != cudaSuccess) return P2P_ERR_CUDA;
        if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->changed[kRoundsPerCheck - 1]) == 0) return P2P_OK;
    }
    return 1;
}

// blocks per subsequence -> exclusive offsets; every interval must hold its quota of blocks (its padding bits may decode
// as a few more): else damaged data.  With `gate` the verdict also stays on the device for the write pass queued behind.
// Caller holds the lock.
int huff_count_check(p2p_ctx *ctx, Slot &s, bool gate) {
    using namespace p2pjdec;
    Slot::JdRun &R = s.jd_run;
    cudaStream_t st = s.stream;
    s.jd_flags_h->bad = 0;   // only the check kernel queued below writes it (an optimistic run may have left its own verdict)
    p2pjpeg::jpeg_scan_kernel<<<1, 1024, 0, st>>>(s.jd_nblk, s.jd_nblk + R.nsub4, nullptr, R.G.n_sub, R.nsub4, &s.jd_flags_d->total);
    huff_check_kernel<<<(R.G.n_ivl + 255) / 256, 256, 0, st>>>(s.jd_nblk + R.nsub4, s.jd_nblk, R.d_ivl_first, R.G, &s.jd_flags_d->bad);
    ctx->launches += 2;
    if (gate) {
        huff_gate_kernel<<<1, 1, 0, st>>>(&s.jd_flags_d->changed[1], &s.jd_flags_d->bad, s.jd_gate_d, &s.jd_flags_d->gate);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    return P2P_OK;
}

// write pass + DC values.  Caller holds the lock.
int huff_finish(p2p_ctx *ctx, Slot &s, bool gated) {
    using namespace p2pjdec;
    Slot::JdRun &R = s.jd_run;
    const HuffGeom &G = R.G;
    cudaStream_t st = s.stream;
    const int *gate = gated ? s.jd_gate_d : nullptr;
    int32_t *dcdiff = s.jd_dc;
    uint32_t *dcsum = reinterpret_cast<uint32_t *>(s.jd_dc + 3 * (size_t)G.dc_stride);
    if (R.fast)
        huff_write_fast_kernel<<<R.sgrid, kSyncThreads, 0, st>>>(s.jd_stream, s.jd_tables, R.d_lut, G, R.d_sub, R.d_ivl_first, s.jd_states,
                                                                s.jd_nblk + R.nsub4, s.jd_coef_d, dcdiff, &s.jd_flags_d->out_of_range, gate);
    else
        huff_write_kernel<<<R.sgrid, 128, 0, st>>>(s.jd_stream, s.jd_tables, G, R.d_sub, R.d_ivl_first, s.jd_states, s.jd_nblk + R.nsub4,
                                                 s.jd_coef_d, dcdiff, &s.jd_flags_d->out_of_range, gate);
    // per component: exclusive sums of the DC differences in scan order (mod 2^32 arithmetic = two's complement sums)
    // (the counts per subsequence in s.jd_nblk stay as they are: a gated run that did not pass continues from them)
    uint32_t *n_per_comp = reinterpret_cast<uint32_t *>(s.jd_gate_d + 4);
    CK(cudaMemcpyAsync(n_per_comp, G.dc_count, 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    {   // three-phase scan over the whole GPU (the luma plane of an 8K file has 524,288 differences)
        const uint32_t n_tiles = (G.dc_stride + 4095u) / 4096u;
        const size_t tiles_stride = ((size_t)n_tiles + 3) & ~(size_t)3;
        uint32_t *tile_sums = s.jd_tiles, *tile_offs = s.jd_tiles + 3 * tiles_stride;
        const uint32_t *in = reinterpret_cast<const uint32_t *>(dcdiff);
        p2pjpeg::scan_tile_sums_kernel<<<dim3(n_tiles, 3), 1024, 0, st>>>(in, n_per_comp, 0u, (size_t)G.dc_stride, tile_sums,
                                                                         tiles_stride);
        p2pjpeg::jpeg_scan_kernel<<<3, 1024, 0, st>>>(tile_sums, tile_offs, nullptr, n_tiles, tiles_stride, s.jd_tot_d);
        p2pjpeg::scan_tiles_apply_kernel<<<dim3(n_tiles, 3), 1024, 0, st>>>(in, dcsum, n_per_comp, 0u, (size_t)G.dc_stride,
                                                                           tile_offs, tiles_stride);
    }
    huff_dc_kernel<<<dim3((G.dc_stride + 255) / 256, 3), 256, 0, st>>>(dcdiff, dcsum, G, s.jd_coef_d);
    ctx->launches += 5;
    CK(cudaGetLastError());
    return P2P_OK;
}

// Host-checked form (modes 2 and 3, and the continuation of an optimistic run that needs more rounds): P2P_OK = the
// coefficients are queued, 1 = not handled (no convergence, block counts off).
int huff_checked_tail(p2p_ctx *ctx, Slot &s) {
    int rc = huff_rounds_checked(ctx, s);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        rc = huff_count_check(ctx, s, false);
        if (rc) return rc;
    }
    if (wait_slot(ctx, s) != cudaSuccess) return P2P_ERR_CUDA;
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->bad) != 0) return 1;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    rc = huff_finish(ctx, s, false);
    if (rc) return rc;
    ctx->gpu_huffman_used++;
    return P2P_OK;
}

// Huffman stage on the device (p2pjdec::huff_*): fills s.jd_coef_d.  Returns P2P_OK, an error, or 1 = "not handled"
// (no convergence, inconsistent block counts, unexpected markers): the caller then runs the host decoder.
// Mode 1 queues everything optimistically and leaves s.jd_run.pending set: the verdict is read by jd_resolve after the
// caller's wait.  The lock is held only while enqueueing.
int device_huffman(p2p_ctx *ctx, Slot &s, const uint8_t *file, size_t len, const p2pjdec::Parsed &P, int mode) {
    using namespace p2pjdec;
    int rc = huff_begin(ctx, s, file, len, P, mode);
    if (rc) return rc;
    if (mode != 1) return huff_checked_tail(ctx, s);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    // nothing of this slot ran when huff_begin queued round 0 (the stream was drained for the destuffed length or by the
    // caller), and the flags are only written by rounds queued below
    s.jd_flags_h->changed[0] = 0;
    s.jd_flags_h->changed[1] = 0;
    s.jd_flags_h->gate = 0;
    for (int r = 0; r < kOptimisticRounds; ++r)
        launch_round(s, 0, &s.jd_flags_d->changed[r + 1 == kOptimisticRounds ? 1 : 0]);
    ctx->launches += kOptimisticRounds;
    CK(cudaGetLastError());
    rc = huff_count_check(ctx, s, true);
    if (!rc) rc = huff_finish(ctx, s, true);
    if (rc) return rc;
    s.jd_run.pending = true;
    ctx->gpu_huffman_used++;
    return P2P_OK;
}

// IDCT, upsampling and colour conversion of the coefficients in s.jd_coef_d (or s.jd_coef_h when the host decoded them)
// into the slot's BGR staging image (s.d_bgr, row stride = Wp * 3 rounded to 4) or, `packed`, straight into the slot's
// panorama (what launch_pack would make of the staging image: p2p_upload_pano_jpeg never needs the BGR form).
int jd_pixels(p2p_ctx *ctx, Slot &s, const p2pjdec::Parsed &P, bool coef_on_device, size_t *dstride, bool packed) {
    using namespace p2pjdec;
    const Info &I = P.info;
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
    if (!rc && !packed) rc = ensure(ctx, &s.d_bgr, &s.bgr_cap, *dstride * I.H);
    if (packed) s.valid = false;   // the colour kernel writes the panorama itself: whatever the slot held is gone from here on
    if (!rc && packed) rc = prepare_slot(ctx, s, I.W, I.H);
    cudaSurfaceObject_t surf = 0;
    if (!rc && packed && ctx->opt_sampler == 1) {
        rc = ensure_array(ctx, s);
        surf = s.surf;
    }
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
    C.rgba = s.d_rgba;
    C.pitch_tex = s.pitch_tex;
    C.surf = surf;
    if (I.H > 65535) return fail(ctx, P2P_ERR_LIMIT, "image too tall for one grid");
    const dim3 cgrid(((I.W + 3) / 4 + 255) / 256, I.H);
    if (packed) {
        jpegdec_color_kernel<true><<<cgrid, 256, 0, s.stream>>>(C);
        s.valid = true;
        s.row0 = 0;
        s.row1 = I.H;
        s.tex_current = (surf != 0);
    } else {
        jpegdec_color_kernel<false><<<cgrid, 256, 0, s.stream>>>(C);
    }
    ctx->launches += 4;
    CK(cudaGetLastError());
    return P2P_OK;
}

int jd_host_scan(p2p_ctx *ctx, Slot &s, const uint8_t *file, size_t len, const p2pjdec::Parsed &P) {
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->gpu_huffman_fallback++;
        cudaSetDevice(ctx->device);
        wait_slot(ctx, s);   // the staging buffer was used for the stream upload
    }
    // (the block-norm check of a progressive file is left to the IDCT kernel, which makes it anyway)
    return p2pjdec::decode_scan(file, len, P, s.jd_coef_h, false) ? P2P_ERR_UNSUPPORTED : P2P_OK;  // damaged: leave it to libjpeg
}

// Decode `file` into the slot's BGR staging image.  The Huffman stage runs on the device (queued optimistically in the
// default mode: see jd_resolve) or on the calling thread WITHOUT the context lock; the lock is only taken to size
// buffers and to enqueue.
int decode_jpeg_to_staging(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, p2pjdec::Parsed &P, size_t *dstride,
                           bool packed) {
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
            CK(wait_slot(ctx, s));  // an earlier upload may still read the old staging buffer
            if (s.jd_coef_h) CK(cudaFreeHost(s.jd_coef_h));
            s.jd_coef_h = nullptr;
            s.jd_coef_h_cap = 0;
            CK(cudaHostAlloc(reinterpret_cast<void **>(&s.jd_coef_h), bytes, cudaHostAllocPortable));
            s.jd_coef_h_cap = bytes;
        } else {
            CK(wait_slot(ctx, s));
        }
        s.jd_flags_h->out_of_range = 0;   // "damaged data" flag of the write pass and the IDCT; the stream is drained
        s.jd_run.pending = false;
    }
    bool coef_on_device = false;
    if (use_gpu && !P.progressive) {   // (a progressive file's scans are decoded on the calling thread)
        const int rc = device_huffman(ctx, s, file, len, P, use_gpu);
        if (rc == P2P_OK) coef_on_device = true;
        else if (rc != 1) return rc;
    }
    if (!coef_on_device) {
        const int rc = jd_host_scan(ctx, s, file, len, P);
        if (rc) return rc;
    }
    return jd_pixels(ctx, s, P, coef_on_device, dstride, packed);
}

// After the caller's wait on the slot's stream: the verdict of an optimistically queued Huffman stage.  P2P_OK = the
// staging image stands; 2 = the stage had to be completed (more rounds on the device, or the host decoder) and the
// staging image was queued again: the caller repeats what it queued behind it and waits once more; else an error.
int jd_resolve(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, const p2pjdec::Parsed &P, size_t *dstride, bool packed) {
    Slot &s = ctx->slots[slot];
    if (!s.jd_run.pending) return P2P_OK;
    s.jd_run.pending = false;
    const int gate = *reinterpret_cast<volatile int *>(&s.jd_flags_h->gate);
    if (gate == 1) return P2P_OK;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->gpu_huffman_used--;              // counted when it was queued
        s.jd_flags_h->out_of_range = 0;       // whatever ran behind the closed gate saw no coefficients
    }
    bool coef_on_device = false;
    if (gate == 0) {                          // not at the fixed point yet: continue from the states reached
        const int rc = huff_checked_tail(ctx, s);
        if (rc == P2P_OK) coef_on_device = true;
        else if (rc != 1) return rc;
    }
    if (!coef_on_device) {
        const int rc = jd_host_scan(ctx, s, file, len, P);
        if (rc) return rc;
    }
    const int rc = jd_pixels(ctx, s, P, coef_on_device, dstride, packed);
    return rc ? rc : 2;
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
    Slot &s = ctx->slots[slot];
    int rc = decode_jpeg_to_staging(ctx, slot, file, len, P, &dstride, true);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
    if (rc) return rc;
    *Wp = P.info.W;
    *Hp = P.info.H;
    for (int attempt = 0; attempt < 2; ++attempt) {
        // the IDCT reports coefficient blocks no 8-bit encoder produces (damaged data), an optimistically queued Huffman
        // stage its verdict: wait for both outside the lock
        cudaSetDevice(ctx->device);
        if (wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "JPEG decoder: CUDA error");
        rc = jd_resolve(ctx, slot, file, len, P, &dstride, true);
        if (rc == P2P_OK) break;
        if (rc != 2) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            s.valid = false;
            if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
            return rc;
        }
    }
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
    int rc = decode_jpeg_to_staging(ctx, slot, file, len, P, &dstride, false);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
    if (rc) return rc;
    if (row_stride < (size_t)P.info.W * 3 || capacity_rows < (size_t)P.info.H)
        return fail(ctx, P2P_ERR_INVALID, "output buffer smaller than the image (see p2p_jpeg_probe)");
    Slot &s = ctx->slots[slot];
    for (int attempt = 0; attempt < 2; ++attempt) {
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            CK(cudaSetDevice(ctx->device));
            s.valid = false;  // the staging image changed under whatever panorama the slot held
            CK(cudaMemcpy2DAsync(bgr_host, row_stride, s.d_bgr, dstride, (size_t)P.info.W * 3, P.info.H, cudaMemcpyDeviceToHost,
                                 s.stream));
        }
        cudaSetDevice(ctx->device);
        if (wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "JPEG decoder: CUDA error");
        rc = jd_resolve(ctx, slot, file, len, P, &dstride, false);
        if (rc == P2P_OK) break;
        if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, "JPEG file outside the supported subset (fall back to cv2.imread)");
        if (rc != 2) return rc;
    }
    if (*reinterpret_cast<volatile int *>(&s.jd_flags_h->out_of_range))
        return fail(ctx, P2P_ERR_UNSUPPORTED, "JPEG data out of range (damaged file: fall back to cv2.imread)");
    return P2P_OK;
}

}  // extern "C"
