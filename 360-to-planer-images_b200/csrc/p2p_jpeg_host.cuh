// p2p_jpeg_host.cuh - host side of the JPEG encoder: tables (quantisation, Huffman, file header) exactly as
// libjpeg-turbo forms them at OpenCV's defaults, per-slot scratch, launches.  Included by p2p_api.cu.
#pragma once
#include "p2p_jpeg.cuh"

#include <string.h>

#include <vector>

namespace p2pjpeg {

// Annex K tables (jcparam.c std_huff_tables) and quantisation tables (jcparam.c std_luminance_quant_tbl, ...)
static const uint8_t kZigzagNat[64] = {  // zigzag position -> natural index (jutils.c jpeg_natural_order)
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63};
static const uint8_t kLumaQ[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
    80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
    95, 98, 112, 100, 103, 99};
static const uint8_t kChromaQ[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99};
static const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
    0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
    0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
    0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
    0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};
static const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
    0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
    0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};

// jchuff.c jpeg_make_c_derived_tbl: (code << 8) | length per symbol
inline void derive(const uint8_t *bits, const uint8_t *vals, uint32_t *table) {
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) table[vals[k++]] = (code++ << 8) | (uint32_t)len;
        code <<= 1;
    }
}

inline void put_seg(std::vector<uint8_t> &o, uint8_t marker, const std::vector<uint8_t> &payload) {
    o.push_back(0xFF);
    o.push_back(marker);
    const size_t len = payload.size() + 2;
    o.push_back((uint8_t)(len >> 8));
    o.push_back((uint8_t)len);
    o.insert(o.end(), payload.begin(), payload.end());
}

// everything that depends on (W, H, quality): jcparam.c jpeg_set_quality(force_baseline), jcmarker.c headers
inline void build_tables(int W, int H, int quality, Tables &T) {
    memset(&T, 0, sizeof(T));
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = (quality < 50) ? 5000 / quality : 200 - quality * 2;
    uint8_t qy[64], qc[64];
    for (int i = 0; i < 64; ++i) {
        long a = ((long)kLumaQ[i] * scale + 50) / 100, b = ((long)kChromaQ[i] * scale + 50) / 100;
        a = a < 1 ? 1 : (a > 255 ? 255 : a);
        b = b < 1 ? 1 : (b > 255 ? 255 : b);
        qy[i] = (uint8_t)a;
        qc[i] = (uint8_t)b;
        T.div_y[i] = (uint16_t)(a * 8);
        T.div_c[i] = (uint16_t)(b * 8);
        T.rcp_y[i] = (uint32_t)((1ull << 32) / (unsigned long long)(a * 8)) + 1u;
        T.rcp_c[i] = (uint32_t)((1ull << 32) / (unsigned long long)(b * 8)) + 1u;
    }
    derive(kDcLumaBits, kDcVals, T.dc[0]);
    derive(kDcChromaBits, kDcVals, T.dc[1]);
    derive(kAcLumaBits, kAcLumaVals, T.ac[0]);
    derive(kAcChromaBits, kAcChromaVals, T.ac[1]);
    std::vector<uint8_t> o = {0xFF, 0xD8};
    put_seg(o, 0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; ++t) {
        std::vector<uint8_t> p = {(uint8_t)t};
        for (int i = 0; i < 64; ++i) p.push_back((t ? qc : qy)[kZigzagNat[i]]);
        put_seg(o, 0xDB, p);
    }
    put_seg(o, 0xC0, {8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    const struct { uint8_t id; const uint8_t *bits, *vals; int n; } huff[4] = {
        {0x00, kDcLumaBits, kDcVals, 12}, {0x10, kAcLumaBits, kAcLumaVals, 162},
        {0x01, kDcChromaBits, kDcVals, 12}, {0x11, kAcChromaBits, kAcChromaVals, 162}};
    for (const auto &h : huff) {
        std::vector<uint8_t> p = {h.id};
        p.insert(p.end(), h.bits, h.bits + 16);
        p.insert(p.end(), h.vals, h.vals + h.n);
        put_seg(o, 0xC4, p);
    }
    put_seg(o, 0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    T.header_len = (int)o.size();
    memcpy(T.header, o.data(), o.size());
}

inline Geometry make_geometry(int W, int H) {
    Geometry G;
    memset(&G, 0, sizeof(G));
    G.W = W;
    G.H = H;
    G.mcux = (W + 15) / 16;
    G.mcuy = (H + 15) / 16;
    G.n_mcu = G.mcux * G.mcuy;
    G.n_blocks = G.n_mcu * 6;
    G.blk_stride = (G.n_blocks + 3) & ~3;
    G.ybw = (W + 7) / 8;
    G.ybh = (H + 7) / 8;
    G.cw = (W + 1) / 2;
    G.ch = (H + 1) / 2;
    G.img_stride = (size_t)W * H * 3;
    // capacity: the raw image size (a q95 file of white noise needs 1.2 bytes / pixel; the kernels report overflow)
    const size_t raw = (size_t)G.mcux * 16 * G.mcuy * 16 * 3 + 1024;
    G.cap_bits_words = ((raw / 4) + 15) & ~(size_t)15;   // 16-byte chunks; chunk arrays stay 16-byte aligned per image
    G.cap_out = (raw + kHeaderMax + 16 + 15) & ~(size_t)15;
    return G;
}

}  // namespace p2pjpeg
