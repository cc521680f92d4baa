// p2p_api_pngdec.inl - host side of the PNG decoder (csrc/p2p_pngdec.cuh: candidate search, chain walk, symbolic inflate, unfilter) and its entry points
// Part of the single translation unit p2p_api.cu (textual include, after p2p_api_png.inl).

namespace {

// where the decoder's small arrays sit in the slot's table buffers (device: s.pd_tab, page-locked: s.pd_tab_h)
struct PdLayout {
    uint32_t n_segs = 0, cap_list = 0, cap_cand = 0, cap_blocks = 0;
    size_t d_segs = 0, d_list = 0, d_cands = 0, d_blocks = 0, d_bands = 0, d_prog = 0, d_ctr = 0, d_total = 0;
    size_t h_segs = 0, h_cands = 0, h_blocks = 0, h_ctr = 0, h_total = 0;
};

constexpr uint32_t kPdSeg = 4096;                 // bytes per CRC segment
constexpr uint64_t kPdHostBudget = 8ull << 20;    // compressed bits of fixed / unannounced blocks the chain walk measures on the host

PdLayout pd_layout(const p2ppdec::Parsed &P) {
    using namespace p2ppdec;
    PdLayout L;
    for (const Idat &c : P.idat) L.n_segs += (c.len + kPdSeg - 1) / kPdSeg;
    const size_t n = P.info.stream_len;
    L.cap_list = (uint32_t)std::min<size_t>(std::max<size_t>(n / 8, 1u << 16), 1u << 26);
    L.cap_cand = (uint32_t)std::min<size_t>(n / 512 + 1024, 1u << 24);
    L.cap_blocks = L.cap_cand + (uint32_t)std::min<size_t>(P.info.raw_bytes / 65535 + 1024, 1u << 24);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t at = 0;
    L.d_segs = at; at = up(at + (size_t)L.n_segs * sizeof(CrcSeg));
    L.d_list = at; at = up(at + (size_t)L.cap_list * sizeof(uint64_t));
    L.d_cands = at; at = up(at + (size_t)L.cap_cand * sizeof(Cand));
    L.d_blocks = at; at = up(at + (size_t)L.cap_blocks * sizeof(Block));
    L.d_bands = at; at = up(at + (size_t)(L.cap_blocks + 1) * sizeof(uint32_t));   // marks in front of each group of the tail pass
    L.d_prog = at; at = up(at + ((size_t)P.info.H + 1) * sizeof(uint32_t));   // progress per band | ticket
    L.d_ctr = at; at = up(at + sizeof(Slot::PdCtr));
    L.d_total = at;
    at = 0;
    L.h_segs = at; at = up(at + (size_t)L.n_segs * sizeof(CrcSeg));
    L.h_cands = at; at = up(at + (size_t)L.cap_cand * sizeof(Cand));
    L.h_blocks = at; at = up(at + (size_t)L.cap_blocks * sizeof(Block));
    L.h_ctr = at; at = up(at + sizeof(Slot::PdCtr));
    L.h_total = at;
    return L;
}

template <typename T>
int ensure_pinned(p2p_ctx *ctx, T **ptr, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return P2P_OK;
    if (*ptr) CK(cudaFreeHost(*ptr));
    *ptr = nullptr;
    *cap = 0;
    bytes += bytes / 2;
    void *p = nullptr;
    CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
    *ptr = static_cast<T *>(p);
    *cap = bytes;
    return P2P_OK;
}

// Decode `file` into the slot's BGR staging image (row stride *dstride).  Two waits on the slot's stream: after the
// candidate search (the host walks the chain) and - by the caller - after everything else, followed by pd_verdict.
// P2P_ERR_UNSUPPORTED = declined (outside the subset, damaged, or not worth it): the caller reads the file with cv2.
int png_to_staging(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, p2ppdec::Parsed &P, size_t *dstride_out) {
    using namespace p2ppdec;
    if (parse_png(file, len, P)) return P2P_ERR_UNSUPPORTED;
    const Info &I = P.info;
    Slot &s = ctx->slots[slot];
    const PdLayout L = pd_layout(P);
    if (L.n_segs > (1u << 22)) return P2P_ERR_UNSUPPORTED;   // millions of tiny IDAT chunks: not worth a table of them
    const uint64_t n_words = stream_words(P), stream_bits = (uint64_t)(I.stream_len - 4) * 8;
    const size_t dstride = ((size_t)I.W * 3 + 3) & ~(size_t)3;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        int rc = check_dims(ctx, I.W, I.H);
        if (rc) return rc;
        CK(cudaSetDevice(ctx->device));
        CK(wait_slot(ctx, s));  // earlier work on this slot may still read the page-locked buffers
        rc = ensure_crc_table(ctx);
        if (!rc) rc = ensure_pinned(ctx, &s.pd_zs_h, &s.pd_zs_h_cap, n_words * 4);
        if (!rc) rc = ensure_pinned(ctx, &s.pd_tab_h, &s.pd_tab_h_cap, L.h_total);
        if (!rc) rc = ensure_grow(ctx, &s.pd_zs, &s.pd_zs_cap, n_words * 4);
        if (!rc) rc = ensure_grow(ctx, &s.pd_tab, &s.pd_tab_cap, L.d_total);
        if (!rc) rc = ensure(ctx, &s.pd_raw, &s.pd_raw_cap, I.raw_bytes + 64);
        if (!rc) rc = ensure(ctx, &s.pd_ref, &s.pd_ref_cap, (I.raw_bytes + 64) * sizeof(uint16_t));
        if (!rc) rc = ensure(ctx, &s.d_bgr, &s.bgr_cap, dstride * I.H);
        if (rc) return rc;
    }
    // the slot belongs to the calling thread: fill the page-locked buffers outside the lock
    gather_stream(file, P, reinterpret_cast<uint8_t *>(s.pd_zs_h));
    CrcSeg *segs_h = reinterpret_cast<CrcSeg *>(s.pd_tab_h + L.h_segs);
    {
        uint32_t k = 0;
        for (const Idat &c : P.idat)
            for (uint32_t o = 0; o < c.len; o += kPdSeg) segs_h[k++] = CrcSeg{c.stream_off + o, std::min(kPdSeg, c.len - o), 0u};
    }
    Cand *cands_h = reinterpret_cast<Cand *>(s.pd_tab_h + L.h_cands);
    Block *blocks_h = reinterpret_cast<Block *>(s.pd_tab_h + L.h_blocks);
    Slot::PdCtr *ctr_h = reinterpret_cast<Slot::PdCtr *>(s.pd_tab_h + L.h_ctr);
    CrcSeg *segs_d = reinterpret_cast<CrcSeg *>(s.pd_tab + L.d_segs);
    uint64_t *list_d = reinterpret_cast<uint64_t *>(s.pd_tab + L.d_list);
    Cand *cands_d = reinterpret_cast<Cand *>(s.pd_tab + L.d_cands);
    Block *blocks_d = reinterpret_cast<Block *>(s.pd_tab + L.d_blocks);
    uint32_t *prog_d = reinterpret_cast<uint32_t *>(s.pd_tab + L.d_prog);
    uint32_t *gmarks_d = reinterpret_cast<uint32_t *>(s.pd_tab + L.d_bands);
    Slot::PdCtr *ctr_d = reinterpret_cast<Slot::PdCtr *>(s.pd_tab + L.d_ctr);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = s.stream;
        CK(cudaMemcpyAsync(s.pd_zs, s.pd_zs_h, n_words * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(segs_d, segs_h, (size_t)L.n_segs * sizeof(CrcSeg), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctr_d, 0, sizeof(Slot::PdCtr), st));
        pd_crc_kernel<<<(L.n_segs + 127) / 128, 128, 0, st>>>(reinterpret_cast<const uint8_t *>(s.pd_zs), segs_d, L.n_segs, ctx->d_crc_table);
        const uint64_t words_scanned = (stream_bits + 31) / 32;
        pd_quick_kernel<<<(unsigned)((words_scanned + 255) / 256), 256, 0, st>>>(s.pd_zs, n_words, 16, stream_bits, list_d, L.cap_list, &ctr_d->n_list);
        pd_full_kernel<<<148 * 8, 128, 0, st>>>(s.pd_zs, n_words, list_d, &ctr_d->n_list, L.cap_list, cands_d, L.cap_cand, &ctr_d->n_cand);
        pd_measure_kernel<<<148 * 16, 32, 0, st>>>(s.pd_zs, n_words, stream_bits, cands_d, &ctr_d->n_cand, L.cap_cand, I.raw_bytes, I.wsize);
        ctx->launches += 4;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctr_h, ctr_d, sizeof(Slot::PdCtr), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(segs_h, segs_d, (size_t)L.n_segs * sizeof(CrcSeg), cudaMemcpyDeviceToHost, st));
    }
    if (wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "PNG decoder: CUDA error");
    // the candidates (a second, short copy now that their number is known), then the chain
    const uint32_t n_cand = std::min(ctr_h->n_cand, L.cap_cand);
    if (n_cand) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpyAsync(cands_h, cands_d, (size_t)n_cand * sizeof(Cand), cudaMemcpyDeviceToHost, s.stream));
    }
    if (n_cand && wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "PNG decoder: CUDA error");
    std::vector<Cand> cands(cands_h, cands_h + n_cand);
    std::vector<Block> blocks;
    uint64_t end_bit = 0;
    if (walk_chain(s.pd_zs_h, P, cands, blocks, kPdHostBudget, &end_bit, L.cap_blocks)) return P2P_ERR_UNSUPPORTED;
    P.info.adler = be32(reinterpret_cast<const uint8_t *>(s.pd_zs_h) + I.stream_len - 4);
    memcpy(blocks_h, blocks.data(), blocks.size() * sizeof(Block));
    const uint32_t nb = (uint32_t)blocks.size();
    const uint64_t n_match_total = blocks.back().match_off + blocks.back().n_matches;
    const size_t stride = 1 + I.row_bytes;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = s.stream;
        const int rc_m = ensure_grow(ctx, &s.pd_match, &s.pd_match_cap, (size_t)(n_match_total + 1) * sizeof(Match));
        if (rc_m) return rc_m;
        CK(cudaMemcpyAsync(blocks_d, blocks_h, (size_t)nb * sizeof(Block), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(s.pd_ref, 0, I.raw_bytes * sizeof(uint16_t), st));   // no history marks yet
        pd_decode_kernel<<<(nb + kDecoders - 1) / kDecoders, 32, 0, st>>>(s.pd_zs, n_words, stream_bits, blocks_d, nb, s.pd_raw, s.pd_match, I.raw_bytes,
                                                                         I.wsize, &ctr_d->bad);
        pd_copy_kernel<<<(nb + 7) / 8, 256, 0, st>>>(blocks_d, nb, s.pd_match, s.pd_raw, s.pd_ref);
        const uint32_t per = group_size(nb);
        pd_tails_group_kernel<<<(nb + per - 1) / per, 1024, 0, st>>>(blocks_d, nb, per, s.pd_raw, s.pd_ref, gmarks_d);
        pd_tails_chain_kernel<<<1, 1024, 0, st>>>(blocks_d, nb, per, s.pd_raw, s.pd_ref, gmarks_d);
        pd_tails_finish_kernel<<<nb, 256, 0, st>>>(blocks_d, per, s.pd_raw, s.pd_ref);
        pd_resolve_kernel<<<dim3(nb, 4), 256, 0, st>>>(blocks_d, s.pd_raw, s.pd_ref);
        pd_adler_kernel<<<(unsigned)((I.raw_bytes + 4095) / 4096), 256, 0, st>>>(s.pd_raw, I.raw_bytes, ctr_d->sums);
        // the history marks are done with: their buffer takes the reconstructed rows (rows of rstride bytes, word aligned)
        uint8_t *recon = reinterpret_cast<uint8_t *>(s.pd_ref);
        const size_t rstride = (I.row_bytes + 3) & ~(size_t)3;
        CK(cudaMemsetAsync(prog_d, 0, ((size_t)I.H + 1) * sizeof(uint32_t), st));
        const unsigned ugrid = (unsigned)(((I.H + 31) / 32 + 7) / 8);   // a warp per band of 32 rows
        uint32_t *ticket = prog_d + I.H;
        // 4 pixels per lane and step (8: 1.98 instead of 1.75 ms on files without row dependencies, 10.0 instead of 9.1 ms with
        // adaptive filters, profiles/r2_png_decoder_chunk8.jsonl)
#define P2P_UNFILTER(BPP) pd_unfilter_kernel<BPP, 4><<<ugrid, 256, 0, st>>>(s.pd_raw, recon, I.W, I.H, stride, rstride, ticket, prog_d, &ctr_d->bad)
        switch (I.bpp) {
            case 1: P2P_UNFILTER(1); break;
            case 2: P2P_UNFILTER(2); break;
            case 3: P2P_UNFILTER(3); break;
            default: P2P_UNFILTER(4); break;
        }
#undef P2P_UNFILTER
        pd_bgr_kernel<<<dim3((I.W + 255) / 256, I.H), 256, 0, st>>>(recon, I.W, I.H, rstride, I.bpp, s.d_bgr, dstride);
        ctx->launches += 9;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctr_h, ctr_d, sizeof(Slot::PdCtr), cudaMemcpyDeviceToHost, st));
        s.valid = false;  // the staging image changed under whatever panorama the slot held
    }
    // chunk checksums: the segments' CRCs (read back before the chain walk) folded per IDAT chunk while the device works on the
    // passes queued above (libpng refuses a file with a damaged IDAT chunk: so does this decoder, whatever the passes produce)
    {
        const uint32_t x_full = crc_xpow8n(kPdSeg);
        const uint32_t idat0 = crc32_host(reinterpret_cast<const uint8_t *>("IDAT"), 4);
        uint32_t k = 0;
        for (const Idat &c : P.idat) {
            uint32_t crc = idat0;
            for (uint32_t o = 0; o < c.len; o += kPdSeg, ++k) {
                const uint32_t n = segs_h[k].len;
                crc = crc_mulmod(n == kPdSeg ? x_full : crc_xpow8n(n), crc) ^ segs_h[k].crc;
            }
            if (crc != c.crc) return P2P_ERR_UNSUPPORTED;
        }
    }
    *dstride_out = dstride;
    return P2P_OK;
}

// after the caller's wait: did every block decode, does the Adler-32 of the inflated image match, were all filter types valid
int pd_verdict(Slot &s, const p2ppdec::Parsed &P) {
    const PdLayout L = pd_layout(P);
    const Slot::PdCtr *c = reinterpret_cast<const Slot::PdCtr *>(s.pd_tab_h + L.h_ctr);
    if (c->bad) return P2P_ERR_UNSUPPORTED;
    const uint32_t s1 = (uint32_t)((1ull + c->sums[0]) % 65521ull);
    const uint32_t s2 = (uint32_t)((P.info.raw_bytes % 65521ull + c->sums[1]) % 65521ull);
    return (((s2 << 16) | s1) == P.info.adler) ? P2P_OK : P2P_ERR_UNSUPPORTED;
}

const char *kPngDeclined = "PNG file outside the supported subset (fall back to cv2.imread)";

}  // namespace

extern "C" {

// ---- PNG panoramas decoded on the device (the decode side of cv2.imread for .png inputs, ref :244) ----------------
int p2p_png_probe(const uint8_t *file, size_t len, int *W, int *H) {
    if (!file || !W || !H) return P2P_ERR_INVALID;
    p2ppdec::Parsed P;
    if (p2ppdec::parse_png(file, len, P)) return P2P_ERR_UNSUPPORTED;
    *W = P.info.W;
    *H = P.info.H;
    return P2P_OK;
}

int p2p_png_decode_host(const uint8_t *file, size_t len, uint8_t *bgr, size_t row_stride, size_t capacity_rows, uint64_t *stats) {
    if (!file || !bgr) return P2P_ERR_INVALID;
    p2ppdec::Parsed P;
    if (p2ppdec::parse_png(file, len, P)) return P2P_ERR_UNSUPPORTED;
    if (row_stride < (size_t)P.info.W * 3 || capacity_rows < (size_t)P.info.H) return P2P_ERR_INVALID;
    return p2ppdec::decode_host_model(file, len, bgr, row_stride, stats) ? P2P_ERR_UNSUPPORTED : P2P_OK;
}

int p2p_upload_pano_png(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, int *Wp, int *Hp) {
    P2P_NVTX("p2p_upload_pano_png");
    if (!slot_ok(ctx, slot) || !file || !Wp || !Hp) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    p2ppdec::Parsed P;
    size_t dstride = 0;
    Slot &s = ctx->slots[slot];
    int rc = png_to_staging(ctx, slot, file, len, P, &dstride);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, kPngDeclined);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        rc = prepare_slot(ctx, s, P.info.W, P.info.H);
        if (!rc) rc = launch_pack(ctx, s, s.d_bgr, dstride, 0, P.info.H);
        if (rc) return rc;
    }
    cudaSetDevice(ctx->device);
    if (wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "PNG decoder: CUDA error");
    if (pd_verdict(s, P)) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        s.valid = false;
        return fail(ctx, P2P_ERR_UNSUPPORTED, "PNG data damaged (fall back to cv2.imread)");
    }
    *Wp = P.info.W;
    *Hp = P.info.H;
    return P2P_OK;
}

int p2p_decode_png(p2p_ctx *ctx, int slot, const uint8_t *file, size_t len, uint8_t *bgr_host, size_t row_stride,
                   size_t capacity_rows) {
    P2P_NVTX("p2p_decode_png");
    if (!slot_ok(ctx, slot) || !file || !bgr_host) return fail(ctx, P2P_ERR_INVALID, "bad argument");
    {
        p2ppdec::Parsed P0;
        if (p2ppdec::parse_png(file, len, P0)) return fail(ctx, P2P_ERR_UNSUPPORTED, kPngDeclined);
        if (row_stride < (size_t)P0.info.W * 3 || capacity_rows < (size_t)P0.info.H)
            return fail(ctx, P2P_ERR_INVALID, "output buffer smaller than the image (see p2p_png_probe)");
    }
    p2ppdec::Parsed P;
    size_t dstride = 0;
    Slot &s = ctx->slots[slot];
    int rc = png_to_staging(ctx, slot, file, len, P, &dstride);
    if (rc == P2P_ERR_UNSUPPORTED) return fail(ctx, rc, kPngDeclined);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpy2DAsync(bgr_host, row_stride, s.d_bgr, dstride, (size_t)P.info.W * 3, P.info.H, cudaMemcpyDeviceToHost, s.stream));
    }
    cudaSetDevice(ctx->device);
    if (wait_slot(ctx, s) != cudaSuccess) return fail(ctx, P2P_ERR_CUDA, "PNG decoder: CUDA error");
    if (pd_verdict(s, P)) return fail(ctx, P2P_ERR_UNSUPPORTED, "PNG data damaged (fall back to cv2.imread)");
    return P2P_OK;
}

}  // extern "C"
