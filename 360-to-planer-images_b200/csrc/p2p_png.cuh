// p2p_png.cuh - sm_100a PNG encoder for the projected views (the encode side, default --output_format).
//
// The reference writes every view with cv2.imwrite (ref app/panorama_to_plane-pitch.py:277); for ".png" (the default,
// ref :400-405) that is libpng + zlib at OpenCV's settings: filter Sub on every row, zlib level 1, strategy Z_RLE,
// memLevel 8, 32 KiB window, IDAT chunks of 8192 bytes.  Everything is deterministic integer work, restated here so the
// file is byte-identical to cv2.imwrite's (oracle/png_model.py is the Python restatement, pinned against zlib and
// cv2.imencode):
//   deflate.c  deflate_rle: greedy matches at distance 1 - a pure function of the maximal byte runs, so the tokens of
//              every position follow from "where did my run start" (one max-scan) and two look-ahead bytes
//   trees.c    one thread per deflate block (16383 symbols) runs build_tree / gen_bitlen / gen_codes / scan_tree /
//              send_tree / build_bl_tree exactly (heap order and depth tie-breaks included) and picks static / dynamic
//   emission   one CTA per block: local prefix sum of code lengths, LSB-first bits OR-ed into the stream
//   pngwrite.c Adler-32, 8192-byte IDAT chunks, CRC-32 per chunk, IHDR / IEND
// Also restated: blocks zlib stores uncompressed (_tr_stored_block: white noise), the window bits libpng writes into the
// zlib header of small images (optimize_cmf, <= 16384 bytes of data), filter type 0 for images one pixel wide, streams
// ending exactly on an IDAT boundary.  sizes[i] = 0 ("fall back to cv2.imwrite") is only left for a stream that would
// not fit its buffer, which the buffer sizing excludes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p2ppng {

constexpr int kSymPerBlock = 16383;     // lit_bufsize - 1 with memLevel 8
constexpr int kLCodes = 286, kDCodes = 30, kBlCodes = 19, kHeapSize = 2 * kLCodes + 1, kMaxBits = 15;
constexpr int kHdrWords = 80;           // block header (tree description) capacity in 32-bit words
constexpr int kIdat = 8192;

struct Geom {
    int W, H, n;                 // image size, number of images
    uint32_t row_bytes;          // 1 + 3 W
    uint32_t N;                  // filtered bytes per image = H * row_bytes
    uint32_t Npad;               // per-image stride of the per-position arrays (multiple of 4096)
    uint32_t max_blk;            // per-image stride of the per-block arrays
    size_t img_stride;           // bytes between input images
    size_t z_cap;                // bytes of the zlib stream buffer per image (multiple of 4)
    size_t out_cap;              // bytes of the output file buffer per image
};

struct BlockInfo {
    uint32_t lcode[kLCodes];     // code | len << 16 for literal / length symbols
    uint32_t dcode0;             // code | len << 16 of distance code 0
    uint32_t hdr[kHdrWords];     // header bits, LSB first
    uint32_t hdr_bits;           // 3 block-type bits + tree description
    uint32_t bits;               // whole block: header + symbols + end-of-block
    uint32_t kind;               // 1 static, 2 dynamic, 0 stored
    uint32_t pad;
};

__constant__ uint8_t kExtraL[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t kExtraD[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t kExtraBl[19] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
__constant__ uint8_t kBlOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
__constant__ uint16_t kBaseLen[29] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128,
                                      160, 192, 224, 255};

// trees.c _length_code for lc = match length - 3
__device__ __forceinline__ int length_code(int lc) {
    if (lc == 255) return 28;
    if (lc < 8) return lc;
    const int b = 31 - __clz(lc);          // 3..7
    return 4 * (b - 1) + ((lc >> (b - 2)) & 3);
}

// ---- filter ---------------------------------------------------------------------------------------------------------
// F[y][0] = 1 (Sub), F[y][1 + i] = rgb[i] - rgb[i - 3] (mod 256), RGB order (png_set_bgr).  Images one pixel wide get
// filter type 0: libpng drops Sub / Avg / Paeth for width 1 (same bytes, other type)
__global__ void __launch_bounds__(256)
png_filter_kernel(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ F, const Geom G) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;   // pixel
    const uint32_t y = blockIdx.y;
    const int img = blockIdx.z;
    if (x >= (uint32_t)G.W) return;
    const uint8_t *p = bgr + (size_t)img * G.img_stride + ((size_t)y * G.W + x) * 3;
    uint8_t *o = F + (size_t)img * G.Npad + (size_t)y * G.row_bytes + 1 + (size_t)x * 3;
    int b = p[0], g = p[1], r = p[2];
    if (x > 0) {
        r -= p[-1];
        g -= p[-2];
        b -= p[-3];
    } else {
        o[-1] = (G.W == 1) ? 0 : 1;
    }
    o[0] = (uint8_t)r;
    o[1] = (uint8_t)g;
    o[2] = (uint8_t)b;
}

// ---- chunked scans over the positions of an image ------------------------------------------------------------------
// A CTA owns one chunk of kScanChunk positions (tiles of 4096).  WRITE = false computes only the chunk aggregate
// (chunk_agg), png_chunk_carry_kernel turns the aggregates into the carry every chunk starts from, WRITE = true
// repeats the scan with that carry and writes the results.
// MODE 0: s[i] = start of the maximal run of equal bytes containing i (inclusive max-scan of run-start indices)
// MODE 1: token flags from (F, s) and their exclusive sum (the token index, not stored): writes tlen[i] (0 none,
//         1 literal, 3..258 match), the position of every kSymPerBlock-th token (= deflate block starts), the token count
constexpr uint32_t kScanChunk = 65536;

template <int MODE, bool WRITE>
__global__ void __launch_bounds__(1024)
png_scan_kernel(const uint8_t *__restrict__ F, uint32_t *__restrict__ S, uint16_t *__restrict__ tlen,
                uint32_t *__restrict__ blockpos, uint32_t *__restrict__ chunk_agg, const uint32_t *__restrict__ chunk_carry,
                const Geom G) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int img = blockIdx.y;
    const uint32_t n_chunks = (G.N + kScanChunk - 1) / kScanChunk;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n_chunks) return;
    const uint8_t *f = F + (size_t)img * G.Npad;
    uint32_t *s = S + (size_t)img * G.Npad;
    const uint32_t N = G.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = WRITE ? chunk_carry[(size_t)img * n_chunks + chunk] : 0u;
    __syncthreads();
    const uint32_t c_end = ((chunk + 1) * kScanChunk < N) ? (chunk + 1) * kScanChunk : N;
    for (uint32_t base = chunk * kScanChunk; base < c_end; base += 4096u) {
        const uint32_t i0 = base + threadIdx.x * 4u;
        uint32_t v[4];
        uint32_t tl[4];
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k;
                v[k] = (i < N && i > 0 && f[i] != f[i - 1]) ? i : 0u;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k;
                uint32_t t = 0;
                if (i < N) {
                    const uint32_t j = i - s[i];            // offset inside the run
                    if (j == 0) {
                        t = 1;                              // first byte of a run: literal
                    } else {
                        const uint32_t within = (j - 1) % 258u, p = i - within;   // chunk of <= 258 bytes after the first
                        const bool match = (p + 2 < N) && f[p + 1] == f[p] && f[p + 2] == f[p];
                        if (!match) {
                            t = 1;                          // fewer than 3 equal bytes left: literals
                        } else if (within == 0) {
                            uint32_t e = p + 3;
                            const uint32_t lim = (p + 258u < N) ? p + 258u : N;
                            while (e < lim && f[e] == f[p]) ++e;
                            t = e - p;                      // match length 3..258
                        }
                    }
                }
                tl[k] = t;
                v[k] = t ? 1u : 0u;
            }
        }
        uint32_t agg, inc;
        if (MODE == 0) {
            agg = max(max(v[0], v[1]), max(v[2], v[3]));
            inc = agg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = max(inc, u);
            }
        } else {
            agg = v[0] + v[1] + v[2] + v[3];
            inc = agg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w = (MODE == 0) ? max(w, u) : w + u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (!WRITE) {
            // aggregate only
        } else if (MODE == 0) {
            // exclusive prefix (max) of everything before this thread's 4 items
            uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
            ex = (lane == 0) ? 0u : ex;
            ex = max(ex, warp ? s_warp[warp - 1] : 0u);
            ex = max(ex, carry);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ex = max(ex, v[k]);
                if (i0 + k < N) s[i0 + k] = ex;
            }
        } else {
            uint32_t ex = carry + (inc - agg) + (warp ? s_warp[warp - 1] : 0u);
            uint16_t *tlo = tlen + (size_t)img * G.Npad;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k;
                if (i < N) {
                    tlo[i] = (uint16_t)tl[k];
                    if (v[k] && ex % (uint32_t)kSymPerBlock == 0u)
                        blockpos[(size_t)img * G.max_blk + ex / (uint32_t)kSymPerBlock] = i;
                    ex += v[k];
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = (MODE == 0) ? max(carry, s_warp[31]) : carry + s_warp[31];
        __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) chunk_agg[(size_t)img * n_chunks + chunk] = s_carry;
}

// carry of every chunk = aggregate of the chunks before it (max for MODE 0, sum for MODE 1); total[img] = whole image
template <int MODE>
__global__ void png_chunk_carry_kernel(const uint32_t *__restrict__ chunk_agg, uint32_t *__restrict__ chunk_carry,
                                       uint32_t *__restrict__ total, const Geom G) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= G.n) return;
    const uint32_t n_chunks = (G.N + kScanChunk - 1) / kScanChunk;
    uint32_t acc = 0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        chunk_carry[(size_t)img * n_chunks + c] = acc;
        const uint32_t a = chunk_agg[(size_t)img * n_chunks + c];
        acc = (MODE == 0) ? max(acc, a) : acc + a;
    }
    if (total) total[img] = acc;
}

// position range of deflate block b of an image (the last block may be empty)
__device__ __forceinline__ void block_range(const uint32_t *__restrict__ blockpos, uint32_t T, uint32_t N, uint32_t b,
                                            uint32_t &p0, uint32_t &p1) {
    const uint32_t nblk = T / (uint32_t)kSymPerBlock + 1u;
    p0 = (b * (uint32_t)kSymPerBlock < T) ? blockpos[b] : N;
    p1 = (b + 1 < nblk && (b + 1) * (uint32_t)kSymPerBlock < T) ? blockpos[b + 1] : N;
}

// ---- per-block symbol histogram --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
png_hist_kernel(const uint8_t *__restrict__ F, const uint16_t *__restrict__ tlen, const uint32_t *__restrict__ blockpos,
                const uint32_t *__restrict__ ntok, uint32_t *__restrict__ lfreq, const Geom G) {
    __shared__ uint32_t h[kLCodes + 2];   // [kLCodes] = number of matches (distance code 0)
    const int img = blockIdx.y;
    const uint32_t b = blockIdx.x, T = ntok[img];
    if (b >= T / (uint32_t)kSymPerBlock + 1u) return;
    for (int i = threadIdx.x; i < kLCodes + 2; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t p0, p1;
    block_range(blockpos + (size_t)img * G.max_blk, T, G.N, b, p0, p1);
    const uint8_t *f = F + (size_t)img * G.Npad;
    const uint16_t *tl = tlen + (size_t)img * G.Npad;
    for (uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
        const uint32_t t = tl[i];
        if (t == 1) atomicAdd(&h[f[i]], 1u);
        else if (t >= 3) {
            atomicAdd(&h[257 + length_code((int)t - 3)], 1u);
            atomicAdd(&h[kLCodes], 1u);
        }
    }
    __syncthreads();
    uint32_t *o = lfreq + ((size_t)img * G.max_blk + b) * (kLCodes + 2);
    for (int i = threadIdx.x; i < kLCodes + 2; i += blockDim.x) o[i] = (i == 256) ? 1u : h[i];   // END_BLOCK once
}

// ---- trees.c, one thread per deflate block ---------------------------------------------------------------------------
struct TreeWork {
    uint32_t freq[kHeapSize];
    uint16_t len[kHeapSize], dad[kHeapSize];
    uint8_t depth[kHeapSize];
    // heap entries carry their sort key: (freq << 8 | depth) << 16 | node, so that trees.c's smaller(n, m) =
    // "freq[n] < freq[m] || (freq[n] == freq[m] && depth[n] <= depth[m])" is one compare of the keys (no indirection)
    unsigned long long heap[kHeapSize + 1];
    uint16_t bl_count[kMaxBits + 1];
};

struct BitAcc {      // LSB-first bit writer into a word array
    uint32_t *w;
    uint32_t nbits;
    __device__ __forceinline__ void send(uint32_t value, int length) {
        const uint32_t i = nbits >> 5, sh = nbits & 31;
        w[i] |= value << sh;
        if (sh + length > 32) w[i + 1] |= value >> (32 - sh);
        nbits += (uint32_t)length;
    }
};

__device__ __forceinline__ uint32_t bi_reverse(uint32_t code, int len) {
    return __brev(code) >> (32 - len);
}

// build_tree + gen_bitlen + gen_codes for the `elems` symbols whose frequencies sit in t.freq[0 .. elems).  Code lengths
// end up in t.len, codes in codes[]; returns max_code.  opt_len / static_len are updated like in trees.c.
__device__ int build_tree(TreeWork &t, int elems, const uint8_t *stree_len, const uint8_t *extra, int base, int max_length,
                          long long &opt_len, long long &static_len, uint16_t *codes) {
    int heap_len = 0, heap_max = kHeapSize, max_code = -1;
    auto entry = [&](int n) { return ((((unsigned long long)t.freq[n] << 8) | t.depth[n]) << 16) | (unsigned long long)n; };
    for (int n = 0; n < elems; ++n) {
        if (t.freq[n] != 0) {
            t.depth[n] = 0;
            t.heap[++heap_len] = entry(max_code = n);
        } else {
            t.len[n] = 0;
        }
    }
    while (heap_len < 2) {
        const int node = (max_code < 2) ? ++max_code : 0;
        t.freq[node] = 1;
        t.depth[node] = 0;
        t.heap[++heap_len] = entry(node);
        opt_len--;
        if (stree_len) static_len -= stree_len[node];
    }
    auto pqdownheap = [&](int k) {
        const unsigned long long v = t.heap[k];
        int j = k << 1;
        while (j <= heap_len) {
            unsigned long long hj = t.heap[j];
            if (j < heap_len) {
                const unsigned long long hj1 = t.heap[j + 1];
                if ((hj1 >> 16) <= (hj >> 16)) {   // smaller(heap[j + 1], heap[j])
                    hj = hj1;
                    j++;
                }
            }
            if ((v >> 16) <= (hj >> 16)) break;   // smaller(v, heap[j])
            t.heap[k] = hj;
            k = j;
            j <<= 1;
        }
        t.heap[k] = v;
    };
    for (int n = heap_len / 2; n >= 1; --n) pqdownheap(n);
    int node = elems;
    do {
        const int n = (int)(t.heap[1] & 0xFFFFu);
        t.heap[1] = t.heap[heap_len--];
        pqdownheap(1);
        const int m = (int)(t.heap[1] & 0xFFFFu);
        t.heap[--heap_max] = (unsigned long long)n;   // below heap_max only the node numbers are used
        t.heap[--heap_max] = (unsigned long long)m;
        t.freq[node] = t.freq[n] + t.freq[m];
        t.depth[node] = (uint8_t)((t.depth[n] >= t.depth[m] ? t.depth[n] : t.depth[m]) + 1);
        t.dad[n] = t.dad[m] = (uint16_t)node;
        t.heap[1] = entry(node);
        node++;
        pqdownheap(1);
    } while (heap_len >= 2);
    t.heap[--heap_max] = t.heap[1] & 0xFFFFu;
    // gen_bitlen
    for (int b = 0; b <= kMaxBits; ++b) t.bl_count[b] = 0;
    int overflow = 0;
    t.len[(int)t.heap[heap_max]] = 0;
    int h;
    for (h = heap_max + 1; h < kHeapSize; ++h) {
        const int n = (int)t.heap[h];
        int bits = t.len[t.dad[n]] + 1;
        if (bits > max_length) {
            bits = max_length;
            overflow++;
        }
        t.len[n] = (uint16_t)bits;
        if (n > max_code) continue;
        t.bl_count[bits]++;
        const int xbits = (n >= base) ? extra[n - base] : 0;
        const long long f = t.freq[n];
        opt_len += f * (bits + xbits);
        if (stree_len) static_len += f * (stree_len[n] + xbits);
    }
    if (overflow > 0) {
        do {
            int bits = max_length - 1;
            while (t.bl_count[bits] == 0) bits--;
            t.bl_count[bits]--;
            t.bl_count[bits + 1] += 2;
            t.bl_count[max_length]--;
            overflow -= 2;
        } while (overflow > 0);
        for (int bits = max_length; bits != 0; --bits) {
            int n = t.bl_count[bits];
            while (n != 0) {
                const int m = (int)t.heap[--h];
                if (m > max_code) continue;
                if (t.len[m] != bits) {
                    opt_len += ((long long)bits - t.len[m]) * t.freq[m];
                    t.len[m] = (uint16_t)bits;
                }
                n--;
            }
        }
    }
    // gen_codes
    uint32_t next_code[kMaxBits + 1];
    uint32_t code = 0;
    next_code[0] = 0;
    for (int bits = 1; bits <= kMaxBits; ++bits) {
        code = (code + t.bl_count[bits - 1]) << 1;
        next_code[bits] = code;
    }
    for (int n = 0; n <= max_code; ++n) {
        const int l = t.len[n];
        codes[n] = l ? (uint16_t)bi_reverse(next_code[l]++, l) : 0;
    }
    for (int n = max_code + 1; n < elems; ++n) codes[n] = 0;
    return max_code;
}

// scan_tree (SEND = false: count the code-length symbols) / send_tree (SEND = true: emit them)
template <bool SEND>
__device__ void walk_tree(const uint16_t *lens, int max_code, uint32_t *blfreq, const uint16_t *bllen, const uint16_t *blcode,
                          BitAcc *out) {
    int prevlen = -1, nextlen = lens[0], count = 0;
    int max_count = 7, min_count = 4;
    if (nextlen == 0) max_count = 138, min_count = 3;
    for (int n = 0; n <= max_code; ++n) {
        const int curlen = nextlen;
        nextlen = (n + 1 <= max_code) ? lens[n + 1] : 0xffff;
        if (++count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
            if (SEND) for (int k = 0; k < count; ++k) out->send(blcode[curlen], bllen[curlen]);
            else blfreq[curlen] += (uint32_t)count;
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (SEND) {
                    out->send(blcode[curlen], bllen[curlen]);
                    count--;
                } else {
                    blfreq[curlen]++;
                }
            }
            if (SEND) {
                out->send(blcode[16], bllen[16]);
                out->send((uint32_t)(count - 3), 2);
            } else {
                blfreq[16]++;
            }
        } else if (count <= 10) {
            if (SEND) {
                out->send(blcode[17], bllen[17]);
                out->send((uint32_t)(count - 3), 3);
            } else {
                blfreq[17]++;
            }
        } else {
            if (SEND) {
                out->send(blcode[18], bllen[18]);
                out->send((uint32_t)(count - 11), 7);
            } else {
                blfreq[18]++;
            }
        }
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) max_count = 138, min_count = 3;
        else if (curlen == nextlen) max_count = 6, min_count = 3;
        else max_count = 7, min_count = 4;
    }
}

__device__ __forceinline__ int static_l_len(int n) { return n < 144 ? 8 : (n < 256 ? 9 : (n < 280 ? 7 : 8)); }

// _tr_flush_block for one block: trees, static / dynamic decision, header bits, code tables (one thread per block:
// the work is one long data-dependent sequential loop nest)
__global__ void __launch_bounds__(32)
png_tree_kernel(const uint32_t *__restrict__ lfreq, const uint32_t *__restrict__ ntok, const uint32_t *__restrict__ blockpos,
                BlockInfo *__restrict__ info, const Geom G) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    const uint32_t T = ntok[img], nblk = T / (uint32_t)kSymPerBlock + 1u;
    if (b >= nblk) return;
    const uint32_t *fr = lfreq + ((size_t)img * G.max_blk + b) * (kLCodes + 2);
    BlockInfo &bi = info[(size_t)img * G.max_blk + b];
    const uint32_t last = (b + 1 == nblk) ? 1u : 0u;
    TreeWork t;
    uint8_t sl[kLCodes];
    for (int n = 0; n < kLCodes; ++n) sl[n] = (uint8_t)static_l_len(n);
    uint8_t sd[kDCodes];
    for (int n = 0; n < kDCodes; ++n) sd[n] = 5;
    long long opt_len = 0, static_len = 0;
    // literal / length tree
    uint16_t llen[kLCodes], lcodes[kLCodes];
    for (int n = 0; n < kLCodes; ++n) t.freq[n] = fr[n];
    const int lmax = build_tree(t, kLCodes, sl, kExtraL, 257, kMaxBits, opt_len, static_len, lcodes);
    for (int n = 0; n < kLCodes; ++n) llen[n] = (n <= lmax) ? t.len[n] : 0;
    // distance tree (only code 0 can occur: every match has distance 1)
    uint16_t dlen[kDCodes], dcodes[kDCodes];
    for (int n = 0; n < kDCodes; ++n) t.freq[n] = 0;
    t.freq[0] = fr[kLCodes];
    const int dmax = build_tree(t, kDCodes, sd, kExtraD, 0, kMaxBits, opt_len, static_len, dcodes);
    for (int n = 0; n < kDCodes; ++n) dlen[n] = (n <= dmax) ? t.len[n] : 0;
    // bit-length tree
    uint32_t blfreq[kBlCodes];
    for (int n = 0; n < kBlCodes; ++n) blfreq[n] = 0;
    walk_tree<false>(llen, lmax, blfreq, nullptr, nullptr, nullptr);
    walk_tree<false>(dlen, dmax, blfreq, nullptr, nullptr, nullptr);
    uint16_t bllen[kBlCodes], blcodes[kBlCodes];
    for (int n = 0; n < kBlCodes; ++n) t.freq[n] = blfreq[n];
    const int blmax = build_tree(t, kBlCodes, nullptr, kExtraBl, 0, 7, opt_len, static_len, blcodes);
    for (int n = 0; n < kBlCodes; ++n) bllen[n] = (n <= blmax) ? t.len[n] : 0;
    int max_blindex = kBlCodes - 1;
    while (max_blindex >= 3 && bllen[kBlOrder[max_blindex]] == 0) max_blindex--;
    opt_len += 3 * (max_blindex + 1) + 5 + 5 + 4;
    long long opt_lenb = (opt_len + 3 + 7) >> 3;
    const long long static_lenb = (static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    uint32_t p0, p1;
    block_range(blockpos + (size_t)img * G.max_blk, T, G.N, b, p0, p1);
    const long long stored_len = (long long)p1 - (long long)p0;
    for (int i = 0; i < kHdrWords; ++i) bi.hdr[i] = 0;
    BitAcc out;
    out.w = bi.hdr;
    out.nbits = 0;
    if (stored_len + 4 <= opt_lenb) {
        // trees.c _tr_stored_block: 3 header bits, pad to a byte, LEN, ~LEN, the bytes.  (zlib also needs the block's
        // bytes to be still in its window, block_start >= 0: a block chosen here has < 16383 + 1300 bytes - more match
        // bytes and the Huffman form is shorter -, far less than the 32506 that could slide out.)
        bi.kind = 0;
        bi.bits = (uint32_t)stored_len;   // bytes; png_layout_kernel knows the alignment and turns this into bits
        out.send((0u << 1) + last, 3);
        bi.hdr_bits = out.nbits;
        return;
    }
    if (static_lenb == opt_lenb) {
        bi.kind = 1;
        out.send((1u << 1) + last, 3);
        // static codes: canonical codes of the fixed lengths
        uint32_t next_code[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        next_code[7] = 0;                 // 24 codes of 7 bits: 0000000 ..
        next_code[8] = 0x30;              // 152 codes of 8 bits start at 00110000
        next_code[9] = 0x190;             // 112 codes of 9 bits start at 110010000
        for (int n = 0; n < kLCodes; ++n) {
            const int l = static_l_len(n);
            bi.lcode[n] = bi_reverse(next_code[l]++, l) | ((uint32_t)l << 16);
        }
        bi.dcode0 = 0u | (5u << 16);
        bi.bits = (uint32_t)(static_len + 3);
    } else {
        bi.kind = 2;
        out.send((2u << 1) + last, 3);
        out.send((uint32_t)(lmax + 1 - 257), 5);
        out.send((uint32_t)(dmax + 1 - 1), 5);
        out.send((uint32_t)(max_blindex + 1 - 4), 4);
        for (int rank = 0; rank <= max_blindex; ++rank) out.send(bllen[kBlOrder[rank]], 3);
        walk_tree<true>(llen, lmax, nullptr, bllen, blcodes, &out);
        walk_tree<true>(dlen, dmax, nullptr, bllen, blcodes, &out);
        for (int n = 0; n < kLCodes; ++n) bi.lcode[n] = (uint32_t)lcodes[n] | ((uint32_t)llen[n] << 16);
        bi.dcode0 = (uint32_t)dcodes[0] | ((uint32_t)dlen[0] << 16);
        bi.bits = (uint32_t)(opt_len + 3);
    }
    bi.hdr_bits = out.nbits;
}

// bit offset of every block in the zlib stream (after the 2 header bytes), total length, "handled" flag
__global__ void png_layout_kernel(BlockInfo *__restrict__ info, const uint32_t *__restrict__ ntok, uint32_t *__restrict__ blkoff,
                                  unsigned long long *__restrict__ zbits, const Geom G) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= G.n) return;
    const uint32_t nblk = ntok[img] / (uint32_t)kSymPerBlock + 1u;
    unsigned long long off = 16;   // CMF + FLG
    bool ok = true;
    for (uint32_t b = 0; b < nblk; ++b) {
        BlockInfo &bi = info[(size_t)img * G.max_blk + b];
        blkoff[(size_t)img * G.max_blk + b] = (uint32_t)off;
        if (bi.kind == 0) {   // stored: header, padding to the next byte, LEN + ~LEN, data
            const uint32_t pad = (8u - (uint32_t)((off + 3ull) & 7ull)) & 7u;
            bi.bits = 3u + pad + 32u + 8u * bi.bits;
        }
        off += bi.bits;
        if (off + 64 > (unsigned long long)G.z_cap * 8ull) {
            ok = false;
            break;
        }
    }
    zbits[img] = ok ? off : 0ull;   // 0 = not handled (the stream would not fit its buffer)
}

// ---- emission: one CTA per deflate block ------------------------------------------------------------------------------
__device__ __forceinline__ void or_bits(uint32_t *__restrict__ z, unsigned long long pos, unsigned long long value, int length) {
    const uint32_t i = (uint32_t)(pos >> 5), sh = (uint32_t)(pos & 31);
    atomicOr(z + i, (uint32_t)(value << sh));
    if (sh + length > 32) atomicOr(z + i + 1, (uint32_t)(value >> (32 - sh)));
    if (sh + length > 64) atomicOr(z + i + 2, (uint32_t)(value >> (64 - sh)));
}

__global__ void __launch_bounds__(256)
png_emit_kernel(const uint8_t *__restrict__ F, const uint16_t *__restrict__ tlen, const uint32_t *__restrict__ blockpos,
                const uint32_t *__restrict__ ntok, const BlockInfo *__restrict__ info, const uint32_t *__restrict__ blkoff,
                const unsigned long long *__restrict__ zbits, uint32_t *__restrict__ Z, const Geom G) {
    __shared__ uint32_t s_code[kLCodes];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    const int img = blockIdx.y;
    const uint32_t b = blockIdx.x, T = ntok[img];
    if (b >= T / (uint32_t)kSymPerBlock + 1u || zbits[img] == 0ull) return;
    const BlockInfo &bi = info[(size_t)img * G.max_blk + b];
    for (int i = threadIdx.x; i < kLCodes; i += blockDim.x) s_code[i] = bi.lcode[i];
    uint32_t *z = Z + (size_t)img * (G.z_cap / 4);
    const unsigned long long base = blkoff[(size_t)img * G.max_blk + b];
    // header bits
    for (uint32_t w = threadIdx.x; w * 32 < bi.hdr_bits; w += blockDim.x) {
        const int nb = (bi.hdr_bits - w * 32 < 32) ? (int)(bi.hdr_bits - w * 32) : 32;
        or_bits(z, base + (unsigned long long)w * 32, bi.hdr[w], nb);
    }
    uint32_t p0, p1;
    block_range(blockpos + (size_t)img * G.max_blk, T, G.N, b, p0, p1);
    const uint8_t *f = F + (size_t)img * G.Npad;
    if (bi.kind == 0) {   // stored block: LEN, ~LEN and the filtered bytes themselves, byte aligned
        const unsigned long long data = (base + 3ull + 7ull) & ~7ull;
        const uint32_t len = p1 - p0;
        if (threadIdx.x == 0) or_bits(z, data, (unsigned long long)((len & 0xFFFFu) | ((~len & 0xFFFFu) << 16)), 32);
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) or_bits(z, data + 32ull + 8ull * i, f[p0 + i], 8);
        return;
    }
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint16_t *tl = tlen + (size_t)img * G.Npad;
    const uint32_t d0 = bi.dcode0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t0 = p0; t0 < p1; t0 += blockDim.x) {
        const uint32_t i = t0 + threadIdx.x;
        unsigned long long val = 0;
        int nb = 0;
        if (i < p1) {
            const uint32_t t = tl[i];
            if (t == 1) {
                const uint32_t c = s_code[f[i]];
                val = c & 0xFFFFu;
                nb = (int)(c >> 16);
            } else if (t >= 3) {
                const int lc = (int)t - 3, code = length_code(lc);
                const uint32_t c = s_code[257 + code];
                val = c & 0xFFFFu;
                nb = (int)(c >> 16);
                const int ex = kExtraL[code];
                if (ex) {
                    val |= (unsigned long long)(lc - kBaseLen[code]) << nb;
                    nb += ex;
                }
                val |= (unsigned long long)(d0 & 0xFFFFu) << nb;
                nb += (int)(d0 >> 16);
            }
        }
        uint32_t inc = (uint32_t)nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wsum = 0, tile = 0;
        for (int k = 0; k < 8; ++k) {
            if (k < warp) wsum += s_warp[k];
            tile += s_warp[k];
        }
        const uint32_t carry = s_carry;
        if (nb) or_bits(z, base + bi.hdr_bits + carry + wsum + (inc - (uint32_t)nb), val, nb);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tile;
        __syncthreads();
    }
    if (threadIdx.x == 0) {   // END_BLOCK
        const uint32_t c = s_code[256];
        or_bits(z, base + bi.hdr_bits + s_carry, c & 0xFFFFu, (int)(c >> 16));
    }
}

// ---- Adler-32 of the filtered data: s1 = 1 + sum F[i], s2 = N + sum (N - i) F[i]  (mod 65521) ------------------------
__global__ void __launch_bounds__(256)
png_adler_kernel(const uint8_t *__restrict__ F, unsigned long long *__restrict__ sums, const Geom G) {
    const int img = blockIdx.y;
    const uint8_t *f = F + (size_t)img * G.Npad;
    unsigned long long a = 0, b = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < G.N; i += gridDim.x * blockDim.x) {
        const unsigned long long v = f[i];
        a += v;
        b += v * (unsigned long long)(G.N - i);
    }
    a %= 65521ull;
    b %= 65521ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[2 * img], a % 65521ull);
        atomicAdd(&sums[2 * img + 1], b % 65521ull);
    }
}

// ---- file assembly ------------------------------------------------------------------------------------------------
// zlib stream = CMF FLG (78 01) | deflate bits | Adler-32 (big endian); PNG = signature | IHDR | IDAT x k (8192 bytes each) | IEND
__device__ __forceinline__ size_t zlib_len(unsigned long long zb) { return (size_t)((zb + 7) >> 3) + 4; }

// CMF (low byte) and FLG (high byte): 78 01, or the smaller window libpng writes into the header of a stream of at most
// 16384 bytes of data (pngwutil.c optimize_cmf); the deflate bits do not depend on it (every match has distance 1)
__device__ __forceinline__ uint32_t zlib_header(uint32_t data_size) {
    uint32_t cinfo = 7, half = 1u << 14;
    if (data_size <= 16384u) {
        do {
            half >>= 1;
            --cinfo;
        } while (cinfo > 0 && data_size <= half);
    }
    const uint32_t cmf = 0x08u | (cinfo << 4);
    const uint32_t flg = 0x1Fu - ((cmf << 8) % 0x1Fu);
    return cmf | (flg << 8);
}

__global__ void __launch_bounds__(256)
png_pack_kernel(const uint32_t *__restrict__ Z, const unsigned long long *__restrict__ zbits,
                const unsigned long long *__restrict__ sums, uint8_t *__restrict__ out, const Geom G) {
    const int img = blockIdx.y;
    const unsigned long long zb = zbits[img];
    if (zb == 0ull) return;
    const size_t L = zlib_len(zb);
    const uint8_t *z = reinterpret_cast<const uint8_t *>(Z + (size_t)img * (G.z_cap / 4));
    uint8_t *o = out + (size_t)img * G.out_cap + 33;   // after signature (8) + IHDR chunk (25)
    const uint32_t s1 = (uint32_t)((1ull + sums[2 * img]) % 65521ull);
    const uint32_t s2 = (uint32_t)(((unsigned long long)G.N + sums[2 * img + 1]) % 65521ull);
    const uint32_t adler = (s2 << 16) | s1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (size_t)gridDim.x * blockDim.x) {
        uint8_t v;
        if (i == 0) v = (uint8_t)zlib_header(G.N);
        else if (i == 1) v = (uint8_t)(zlib_header(G.N) >> 8);
        else if (i >= L - 4) v = (uint8_t)(adler >> (8 * (L - 1 - i)));
        else v = z[i];
        const size_t c = i / kIdat;
        o[c * (kIdat + 12) + 8 + (i - c * kIdat)] = v;
    }
}

__device__ __forceinline__ uint32_t crc_update(uint32_t crc, uint8_t byte, const uint32_t *__restrict__ table) {
    return table[(crc ^ byte) & 0xFFu] ^ (crc >> 8);
}

// chunk lengths, types and CRCs, signature, IHDR, IEND, file size.  One warp per IDAT chunk: a full chunk's 8192 data
// bytes are 32 segments of 256 bytes, every lane runs the CRC register over its segment from state 0, and lane 0 folds
// the partial states with the linear "advance by 256 zero bytes" operator (crc_table[256 ..] = its 4 byte tables):
//   state(s, A || B) = advance_|B|(state(s, A)) ^ state(0, B).
// The last (short) chunk is done by lane 0 alone.
__device__ __forceinline__ uint32_t crc_advance256(uint32_t c, const uint32_t *__restrict__ t) {
    return t[256 + (c & 0xFFu)] ^ t[512 + ((c >> 8) & 0xFFu)] ^ t[768 + ((c >> 16) & 0xFFu)] ^ t[1024 + (c >> 24)];
}

__global__ void __launch_bounds__(256)
png_finish_kernel(uint8_t *__restrict__ out, const unsigned long long *__restrict__ zbits,
                  const uint32_t *__restrict__ crc_table, unsigned long long *__restrict__ sizes, const Geom G) {
    const int img = blockIdx.y;
    const unsigned long long zb = zbits[img];
    const int lane = threadIdx.x & 31;
    const size_t c = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // chunk = warp
    if (zb == 0ull) {
        if (c == 0 && lane == 0) sizes[img] = 0ull;
        return;
    }
    const size_t L = zlib_len(zb);
    const size_t nchunks = (L + kIdat - 1) / kIdat;
    uint8_t *o = out + (size_t)img * G.out_cap;
    if (c < nchunks) {
        uint8_t *ch = o + 33 + c * (kIdat + 12);
        const uint32_t len = (uint32_t)((c + 1 < nchunks) ? kIdat : L - c * kIdat);
        uint32_t crc = 0xFFFFFFFFu;
        const uint8_t typ[4] = {'I', 'D', 'A', 'T'};
        for (int k = 0; k < 4; ++k) crc = crc_update(crc, typ[k], crc_table);
        if (len == kIdat) {
            const uint8_t *seg = ch + 8 + lane * 256;
            uint32_t r = 0;
            for (int k = 0; k < 256; ++k) r = crc_update(r, seg[k], crc_table);
            for (int k = 0; k < 32; ++k) {
                const uint32_t rk = __shfl_sync(0xffffffffu, r, k);
                crc = crc_advance256(crc, crc_table) ^ rk;
            }
        } else if (lane == 0) {
            for (uint32_t k = 0; k < len; ++k) crc = crc_update(crc, ch[8 + k], crc_table);
        }
        if (lane == 0) {
            ch[0] = (uint8_t)(len >> 24); ch[1] = (uint8_t)(len >> 16); ch[2] = (uint8_t)(len >> 8); ch[3] = (uint8_t)len;
            ch[4] = 'I'; ch[5] = 'D'; ch[6] = 'A'; ch[7] = 'T';
            crc ^= 0xFFFFFFFFu;
            uint8_t *e = ch + 8 + len;
            e[0] = (uint8_t)(crc >> 24); e[1] = (uint8_t)(crc >> 16); e[2] = (uint8_t)(crc >> 8); e[3] = (uint8_t)crc;
        }
    }
    if (c == 0 && lane == 0) {
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        for (int k = 0; k < 8; ++k) o[k] = sig[k];
        uint8_t *h = o + 8;
        h[0] = 0; h[1] = 0; h[2] = 0; h[3] = 13;
        h[4] = 'I'; h[5] = 'H'; h[6] = 'D'; h[7] = 'R';
        h[8] = (uint8_t)(G.W >> 24); h[9] = (uint8_t)(G.W >> 16); h[10] = (uint8_t)(G.W >> 8); h[11] = (uint8_t)G.W;
        h[12] = (uint8_t)(G.H >> 24); h[13] = (uint8_t)(G.H >> 16); h[14] = (uint8_t)(G.H >> 8); h[15] = (uint8_t)G.H;
        h[16] = 8; h[17] = 2; h[18] = 0; h[19] = 0; h[20] = 0;
        uint32_t crc = 0xFFFFFFFFu;
        for (int k = 4; k < 21; ++k) crc = crc_update(crc, h[k], crc_table);
        crc ^= 0xFFFFFFFFu;
        h[21] = (uint8_t)(crc >> 24); h[22] = (uint8_t)(crc >> 16); h[23] = (uint8_t)(crc >> 8); h[24] = (uint8_t)crc;
        const size_t body = 33 + (nchunks - 1) * (kIdat + 12) + 12 + (L - (nchunks - 1) * kIdat);
        uint8_t *e = o + body;
        const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
        for (int k = 0; k < 12; ++k) e[k] = iend[k];
        sizes[img] = (unsigned long long)(body + 12);
    }
}

}  // namespace p2ppng
