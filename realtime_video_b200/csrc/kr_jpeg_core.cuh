// kr_jpeg_core.cuh — per-thread bodies of the device-side JPEG encoder (kr_jpeg.cu).
//
// Frame egress, second half: the reference turns every decoded frame into a JPEG on the HOST
//     TF.to_pil_image(frames[0, idx], "RGB").save(io, format='JPEG', quality=90)        release_server.py:973
// (Pillow -> libjpeg-turbo, 24-thread pool, after a 57.5 MB fp32 device->host copy per 12-frame block).  Here the
// same byte stream is produced on the device, so only the compressed files (~1-3 MB per block) cross PCIe.
// "Same" means byte-identical to Pillow's output: integer colour conversion, 2x2 chroma box filter with the
// alternating bias, the "islow" forward DCT, round-half-away quantisation, baseline Huffman coding with the
// Annex K tables, 0xFF byte stuffing and the marker layout libjpeg writes.
//
// The encoder is four passes; every pass is "one thread = one independent unit", written here as plain functions
// of the thread's index so that the SAME code also compiles for the host (tests/jpeg_emulate.cpp runs the passes
// with loops in place of the grid and is compared with Pillow on the CPU):
//   1. dct_thread    one thread per 8x8 block: pixels -> YCbCr -> (chroma: 2x2 downsample) -> DCT -> quantise ->
//                    zigzag coefficients (int16 x 64, scan order) + the block's AC code length in bits
//   2. scan_*        per frame: DC differences (need the predecessor block) -> bits per block -> exclusive prefix
//                    sum = the block's bit offset in the frame's entropy-coded segment
//   3. emit_thread   one thread per block: Huffman codes OR-ed into the zeroed bit stream at the block's offset
//   4. stuff_*       per frame: header + byte stuffing (0xFF -> 0xFF 0x00, a prefix sum over 0xFF counts) + EOI
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KRJ_HD __host__ __device__ __forceinline__
#else
#define KRJ_HD inline
#endif

namespace krj {

struct HuffTab {
  uint16_t code[256];
  uint8_t size[256];
};
struct Tables {
  HuffTab dc[2];   // 0 luma, 1 chroma; symbols 0..11
  HuffTab ac[2];   // symbol = (run << 4) | size
};

// ITU-T T.81 Annex K.3 tables (the ones jpeg_set_defaults installs, jcparam.c std_huff_tables)
struct HuffSpec {
  uint8_t bits[16];
  uint8_t vals[162];
  int nvals;
};
constexpr HuffSpec kDcLuma = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}, 12};
constexpr HuffSpec kDcChroma = {{0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}, 12};
constexpr HuffSpec kAcLuma = {
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},
    {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
     0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
     0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
     0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
     0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
     0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
     0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
     0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
     0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa},
    162};
constexpr HuffSpec kAcChroma = {
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77},
    {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
     0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
     0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
     0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
     0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
     0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
     0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
     0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
     0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa},
    162};

// canonical code assignment in order of increasing length (jchuff.c jpeg_make_c_derived_tbl)
constexpr HuffTab derive(const HuffSpec& s) {
  HuffTab t{};
  unsigned code = 0;
  int p = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < s.bits[len - 1] && p < s.nvals; ++i, ++p) {
      t.code[s.vals[p]] = static_cast<uint16_t>(code);
      t.size[s.vals[p]] = static_cast<uint8_t>(len);
      ++code;
    }
    code <<= 1;
  }
  return t;
}
constexpr Tables make_tables() { return Tables{{derive(kDcLuma), derive(kDcChroma)}, {derive(kAcLuma), derive(kAcChroma)}}; }

constexpr uint8_t kStdQuant[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
     69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
     81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72, 92,  95,  98,  112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
     99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

// zigzag position -> natural (row-major) index (jutils.c jpeg_natural_order); k is a constant after unrolling
KRJ_HD constexpr int natural(int k) {
  constexpr uint8_t t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return t[k];
}

// ------------------------------------------------------------------------------------------------------------------
// geometry / workspace
// ------------------------------------------------------------------------------------------------------------------
constexpr int kHeaderBytes = 623;        // SOI APP0 DQT DQT SOF0 DHT DHT DHT DHT SOS (fixed for 3-component 4:2:0)
constexpr int kHeaderWords = 160;        // header storage passed by value to the last pass (640 bytes)
constexpr int kMaxBlockWords = 54;       // worst-case code length of one block: 22 (DC) + 63 * 26 (AC) = 1660 bits

struct Geometry {
  int frames, H, W;
  int mcus_x, mcus_y;
  int nblk;                // 8x8 blocks per frame in scan order: (Y00 Y01 Y10 Y11 Cb Cr) per 16x16 MCU
  int raw_words;           // capacity (32-bit words) of one frame's unstuffed bit stream
};
KRJ_HD Geometry make_geometry(int frames, int H, int W) {
  Geometry g;
  g.frames = frames; g.H = H; g.W = W;
  g.mcus_x = W / 16; g.mcus_y = H / 16;
  g.nblk = g.mcus_x * g.mcus_y * 6;
  g.raw_words = ((g.nblk * kMaxBlockWords + 64) + 63) / 64 * 64;
  return g;
}

struct Workspace {          // device pointers carved from the caller's workspace
  int16_t* coefs;           // [frames, nblk, 64] zigzag order
  uint32_t* bits;           // [frames, nblk]  pass 1: AC bits; pass 2: total bits -> exclusive bit offset
  uint32_t* frame_bits;     // [frames] bits of the entropy-coded segment before padding
  uint32_t* raw;            // [frames, raw_words] unstuffed bit stream, big-endian bytes
};

struct QuantTables {        // natural order, jcparam.c jpeg_add_quant_table(scale, force_baseline)
  uint16_t q[2][64];
};
inline QuantTables make_quant(int quality) {
  QuantTables t;
  if (quality <= 0) quality = 1;
  if (quality > 100) quality = 100;
  const long scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < 64; ++i) {
      long v = (static_cast<long>(kStdQuant[c][i]) * scale + 50) / 100;
      if (v <= 0) v = 1;
      if (v > 255) v = 255;
      t.q[c][i] = static_cast<uint16_t>(v);
    }
  return t;
}

struct Header {
  uint32_t w[kHeaderWords];   // kHeaderBytes bytes in memory order, zero padded
};
// jcmarker.c: write_file_header (SOI, JFIF APP0), write_frame_header (DQT per table, SOF0), write_scan_header
// (DHT per table in component order, SOS)
inline Header make_header(int H, int W, const QuantTables& qt) {
  Header h;
  uint8_t* b = reinterpret_cast<uint8_t*>(h.w);
  for (int i = 0; i < kHeaderWords * 4; ++i) b[i] = 0;
  int n = 0;
  auto put = [&](int v) { b[n++] = static_cast<uint8_t>(v); };
  put(0xFF); put(0xD8);
  const uint8_t app0[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  for (uint8_t v : app0) put(v);
  for (int t = 0; t < 2; ++t) {
    put(0xFF); put(0xDB); put(0); put(67); put(t);
    for (int k = 0; k < 64; ++k) put(qt.q[t][natural(k)]);
  }
  put(0xFF); put(0xC0); put(0); put(17); put(8);
  put(H >> 8); put(H & 0xFF); put(W >> 8); put(W & 0xFF);
  put(3);
  put(1); put(0x22); put(0);
  put(2); put(0x11); put(1);
  put(3); put(0x11); put(1);
  const HuffSpec* specs[4] = {&kDcLuma, &kAcLuma, &kDcChroma, &kAcChroma};
  const int ids[4] = {0x00, 0x10, 0x01, 0x11};
  for (int s = 0; s < 4; ++s) {
    put(0xFF); put(0xC4);
    const int len = 2 + 1 + 16 + specs[s]->nvals;
    put(len >> 8); put(len & 0xFF); put(ids[s]);
    for (int i = 0; i < 16; ++i) put(specs[s]->bits[i]);
    for (int i = 0; i < specs[s]->nvals; ++i) put(specs[s]->vals[i]);
  }
  const uint8_t sos[] = {0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
  for (uint8_t v : sos) put(v);
  // n == kHeaderBytes (checked by the launcher)
  h.w[kHeaderWords - 1] = static_cast<uint32_t>(n);
  return h;
}

// ------------------------------------------------------------------------------------------------------------------
// pass 1: pixels -> quantised zigzag coefficients
// ------------------------------------------------------------------------------------------------------------------
KRJ_HD int bit_length(unsigned v) {
#if defined(__CUDA_ARCH__)
  return 32 - __clz(static_cast<int>(v));
#else
  int n = 0;
  while (v) { ++n; v >>= 1; }
  return n;
#endif
}

// the reference's host-side normalisation + torchvision to_pil_image (release_server.py:979-983):
// byte = trunc(clamp((x + 1) * 0.5, 0, 1) * 255), every operation rounded to fp32 separately
KRJ_HD int px_to_u8(float x) {
#if defined(__CUDA_ARCH__)
  float v = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  return static_cast<int>(__float2uint_rz(__fmul_rn(v, 255.0f)));
#else
  volatile float s = x + 1.0f;
  volatile float v = s * 0.5f;
  float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  if (!(v == v)) c = v;                        // NaN propagates like fminf/fmaxf would not; irrelevant for pixels
  volatile float m = c * 255.0f;
  return static_cast<int>(m);
#endif
}

// 8 consecutive pixels of one row, x0 % 8 == 0
struct LoaderF32 {          // fp32 planar [frames, 3, H, W] in [-1, 1] (the VAE decoder's output)
  const float* px;
  int H, W;
  KRJ_HD void load8(int frame, int y, int x0, int* r, int* g, int* b) const {
    const long plane = static_cast<long>(H) * W;
    const float* p = px + static_cast<long>(frame) * 3 * plane + static_cast<long>(y) * W + x0;
#if defined(__CUDA_ARCH__)
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(p)), r1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p + plane)),
                 g1 = __ldg(reinterpret_cast<const float4*>(p + plane + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p + 2 * plane)),
                 b1 = __ldg(reinterpret_cast<const float4*>(p + 2 * plane + 4));
    const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = px_to_u8(rr[i]); g[i] = px_to_u8(gg[i]); b[i] = px_to_u8(bb[i]); }
#else
    for (int i = 0; i < 8; ++i) {
      r[i] = px_to_u8(p[i]); g[i] = px_to_u8(p[plane + i]); b[i] = px_to_u8(p[2 * plane + i]);
    }
#endif
  }
};
struct LoaderRgb8 {         // packed bytes [frames, H, W, 3] (kr_frames_to_rgb8's output / PIL 'RGB' layout)
  const uint8_t* rgb;
  int H, W;
  KRJ_HD void load8(int frame, int y, int x0, int* r, int* g, int* b) const {
    const uint8_t* p = rgb + ((static_cast<long>(frame) * H + y) * W + x0) * 3;     // 24 bytes, 8-byte aligned
#if defined(__CUDA_ARCH__)
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p)), c = __ldg(reinterpret_cast<const uint2*>(p + 8)),
                d = __ldg(reinterpret_cast<const uint2*>(p + 16));
    const uint32_t w[6] = {a.x, a.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[i] = (w[(3 * i) >> 2] >> (8 * ((3 * i) & 3))) & 0xFF;
      g[i] = (w[(3 * i + 1) >> 2] >> (8 * ((3 * i + 1) & 3))) & 0xFF;
      b[i] = (w[(3 * i + 2) >> 2] >> (8 * ((3 * i + 2) & 3))) & 0xFF;
    }
#else
    for (int i = 0; i < 8; ++i) { r[i] = p[3 * i]; g[i] = p[3 * i + 1]; b[i] = p[3 * i + 2]; }
#endif
  }
};

// jccolor.c rgb_ycc_convert, SCALEBITS = 16 (FIX(x) = x * 65536 + 0.5)
KRJ_HD int rgb_to_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
KRJ_HD int rgb_to_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
KRJ_HD int rgb_to_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

KRJ_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c jpeg_fdct_islow, one 8-point pass (CONST_BITS = 13, PASS1_BITS = 2); 32-bit arithmetic is exact for
// 8-bit samples (the library's own INT32 guarantee)
template <int kPass>
KRJ_HD void fdct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
  const int tmp0 = d0 + d7, tmp7 = d0 - d7, tmp1 = d1 + d6, tmp6 = d1 - d6;
  const int tmp2 = d2 + d5, tmp5 = d2 - d5, tmp3 = d3 + d4, tmp4 = d3 - d4;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  constexpr int sh = kPass == 0 ? 13 - 2 : 13 + 2;
  if (kPass == 0) {
    d0 = (tmp10 + tmp11) * 4;
    d4 = (tmp10 - tmp11) * 4;
  } else {
    d0 = descale(tmp10 + tmp11, 2);
    d4 = descale(tmp10 - tmp11, 2);
  }
  int z1 = (tmp12 + tmp13) * 4433;
  d2 = descale(z1 + tmp13 * 6270, sh);
  d6 = descale(z1 + tmp12 * (-15137), sh);
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * 9633;
  const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
  z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
  z3 += z5; z4 += z5;
  d7 = descale(t4 + z1 + z3, sh);
  d5 = descale(t5 + z2 + z4, sh);
  d3 = descale(t6 + z2 + z3, sh);
  d1 = descale(t7 + z1 + z4, sh);
}

// jcdctmgr.c forward_DCT: divide by (q << 3), round half away from zero
KRJ_HD int quantise(int v, int q) {
  const unsigned qv = static_cast<unsigned>(q) << 3;
  if (v < 0) return -static_cast<int>((static_cast<unsigned>(-v) + (qv >> 1)) / qv);
  return static_cast<int>((static_cast<unsigned>(v) + (qv >> 1)) / qv);
}

// position of a block in scan order
KRJ_HD int luma_pos(int bx, int by, int mcus_x) { return ((by >> 1) * mcus_x + (bx >> 1)) * 6 + (by & 1) * 2 + (bx & 1); }
KRJ_HD int chroma_pos(int cx, int cy, int c, int mcus_x) { return (cy * mcus_x + cx) * 6 + 4 + c; }

// idx in [0, 4M): luma block, raster order over the (W/8) x (H/8) block grid (adjacent threads read adjacent
// 32-byte spans of the same pixel rows); idx in [4M, 6M): chroma block, (component, cy, cx) with cx fastest.
template <class Loader>
KRJ_HD void dct_thread(const Loader& ld, const Geometry& g, const QuantTables& qt, const Tables& T, int frame, int idx,
                       const Workspace& ws) {
  const int nluma = g.mcus_x * g.mcus_y * 4;
  int d[64];
  int pos, comp;
  if (idx < nluma) {
    const int bw = g.mcus_x * 2;
    const int by = idx / bw, bx = idx - by * bw;
    pos = luma_pos(bx, by, g.mcus_x);
    comp = 0;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      int r[8], gg[8], b[8];
      ld.load8(frame, by * 8 + y, bx * 8, r, gg, b);
#pragma unroll
      for (int x = 0; x < 8; ++x) d[8 * y + x] = rgb_to_y(r[x], gg[x], b[x]) - 128;
    }
  } else {
    const int j = idx - nluma;
    const int per = g.mcus_x * g.mcus_y;
    const int c = j / per, m = j - c * per;
    const int cy = m / g.mcus_x, cx = m - cy * g.mcus_x;
    pos = chroma_pos(cx, cy, c, g.mcus_x);
    comp = 1;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int r0[8], g0[8], b0[8], r1[8], g1[8], b1[8];
        ld.load8(frame, cy * 16 + 2 * y, cx * 16 + half * 8, r0, g0, b0);
        ld.load8(frame, cy * 16 + 2 * y + 1, cx * 16 + half * 8, r1, g1, b1);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          int s;
          if (c == 0)
            s = rgb_to_cb(r0[2 * x], g0[2 * x], b0[2 * x]) + rgb_to_cb(r0[2 * x + 1], g0[2 * x + 1], b0[2 * x + 1]) +
                rgb_to_cb(r1[2 * x], g1[2 * x], b1[2 * x]) + rgb_to_cb(r1[2 * x + 1], g1[2 * x + 1], b1[2 * x + 1]);
          else
            s = rgb_to_cr(r0[2 * x], g0[2 * x], b0[2 * x]) + rgb_to_cr(r0[2 * x + 1], g0[2 * x + 1], b0[2 * x + 1]) +
                rgb_to_cr(r1[2 * x], g1[2 * x], b1[2 * x]) + rgb_to_cr(r1[2 * x + 1], g1[2 * x + 1], b1[2 * x + 1]);
          // jcsample.c h2v2_downsample: bias 1, 2, 1, 2, ... along the output row
          d[8 * y + half * 4 + x] = ((s + ((x & 1) ? 2 : 1)) >> 2) - 128;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    fdct8<0>(d[8 * i], d[8 * i + 1], d[8 * i + 2], d[8 * i + 3], d[8 * i + 4], d[8 * i + 5], d[8 * i + 6], d[8 * i + 7]);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    fdct8<1>(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i]);

  // quantise in zigzag order, count the AC code length (jchuff.c encode_one_block without the DC term)
  const HuffTab& ac = T.ac[comp];
  uint32_t packed[32];
  unsigned bits = 0;
  int run = 0;
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const int n = natural(k);
    const int v = quantise(d[n], qt.q[comp][n]);
    if (k & 1) packed[k >> 1] |= static_cast<uint32_t>(v & 0xFFFF) << 16;
    else packed[k >> 1] = static_cast<uint32_t>(v & 0xFFFF);
    if (k > 0) {
      if (v == 0) {
        ++run;
      } else {
        bits += static_cast<unsigned>(run >> 4) * ac.size[0xF0];
        const int nb = bit_length(static_cast<unsigned>(v < 0 ? -v : v));
        bits += ac.size[((run & 15) << 4) + nb] + nb;
        run = 0;
      }
    }
  }
  if (run > 0) bits += ac.size[0];
  const long blk = static_cast<long>(frame) * g.nblk + pos;
  uint32_t* out = reinterpret_cast<uint32_t*>(ws.coefs + blk * 64);
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<uint4*>(out)[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
#else
  for (int i = 0; i < 32; ++i) out[i] = packed[i];
#endif
  ws.bits[blk] = bits;
}

// ------------------------------------------------------------------------------------------------------------------
// pass 2: bits per block (adds the DC term) and their exclusive prefix sum, per frame
// ------------------------------------------------------------------------------------------------------------------
// predecessor of block b (same component) in scan order, -1 for the first block of a component
KRJ_HD int dc_pred_block(int b) {
  const int m = b / 6, k = b - m * 6;
  if (k == 0) return m > 0 ? b - 3 : -1;
  if (k < 4) return b - 1;
  return m > 0 ? b - 6 : -1;
}
KRJ_HD int dc_diff(const int16_t* frame_coefs, int b) {
  const int p = dc_pred_block(b);
  return frame_coefs[static_cast<long>(b) * 64] - (p >= 0 ? frame_coefs[static_cast<long>(p) * 64] : 0);
}
KRJ_HD int chunk_len(int nblk, int nthreads) { return (nblk + nthreads - 1) / nthreads; }

// thread tid owns the blocks [tid*chunk, min(nblk, (tid+1)*chunk)); returns the bits of its blocks (and stores
// every block's total in place of its AC count)
KRJ_HD unsigned scan_sum_thread(const Geometry& g, const Tables& T, int frame, int tid, int nthreads, const Workspace& ws) {
  const int chunk = chunk_len(g.nblk, nthreads);
  const int16_t* fc = ws.coefs + static_cast<long>(frame) * g.nblk * 64;
  uint32_t* fb = ws.bits + static_cast<long>(frame) * g.nblk;
  unsigned sum = 0;
  for (int b = tid * chunk; b < g.nblk && b < (tid + 1) * chunk; ++b) {
    const int diff = dc_diff(fc, b);
    const int nb = bit_length(static_cast<unsigned>(diff < 0 ? -diff : diff));
    const unsigned tot = fb[b] + T.dc[(b % 6) < 4 ? 0 : 1].size[nb] + nb;
    fb[b] = tot;
    sum += tot;
  }
  return sum;
}
// second walk: replace every block's total by its exclusive bit offset (excl = bits of all earlier threads' blocks)
KRJ_HD void scan_write_thread(const Geometry& g, int frame, int tid, int nthreads, unsigned excl, const Workspace& ws) {
  const int chunk = chunk_len(g.nblk, nthreads);
  uint32_t* fb = ws.bits + static_cast<long>(frame) * g.nblk;
  unsigned off = excl;
  for (int b = tid * chunk; b < g.nblk && b < (tid + 1) * chunk; ++b) {
    const unsigned t = fb[b];
    fb[b] = off;
    off += t;
  }
}
// words of the raw stream that pass 3 will OR into (everything up to and including the padded last byte)
KRJ_HD unsigned raw_words_used(unsigned total_bits) { return (total_bits + 7 + 31) / 32 + 1; }
// jchuff.c flush_bits: the last byte is padded with one-bits.  Big-endian word holding bits [32w, 32w+32).
KRJ_HD uint32_t pad_word(unsigned total_bits, unsigned w) {
  const unsigned end = (total_bits + 7) & ~7u;              // first bit after the padded byte
  uint32_t v = 0;
  for (unsigned bit = total_bits; bit < end; ++bit)
    if ((bit >> 5) == w) v |= 0x80000000u >> (bit & 31);
  return v;
}
KRJ_HD uint32_t to_memory_order(uint32_t big_endian_word) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(big_endian_word, 0, 0x0123);
#else
  return ((big_endian_word & 0xFF) << 24) | ((big_endian_word & 0xFF00) << 8) | ((big_endian_word >> 8) & 0xFF00) |
         (big_endian_word >> 24);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// pass 3: Huffman codes of one block, OR-ed into the frame's bit stream at the block's offset
// ------------------------------------------------------------------------------------------------------------------
template <class OrWord>     // or_word(uint32_t* word, uint32_t value_in_memory_order)
struct BitSink {
  uint32_t* raw;
  unsigned word;            // next word to flush
  unsigned n;               // bits held in acc (low n bits), n < 32 between calls
  uint64_t acc;
  OrWord or_word;
  KRJ_HD void put(unsigned code, int size) {
    acc = (acc << size) | (code & ((1u << size) - 1u));
    n += size;
    if (n >= 32) {
      n -= 32;
      or_word(raw + word, to_memory_order(static_cast<uint32_t>(acc >> n)));
      ++word;
      acc &= (1ull << n) - 1ull;
    }
  }
  KRJ_HD void flush() {
    if (n > 0) or_word(raw + word, to_memory_order(static_cast<uint32_t>(acc << (32 - n))));
  }
};

template <class OrWord>
KRJ_HD void emit_thread(const Geometry& g, const Tables& T, int frame, int b, const Workspace& ws, OrWord or_word) {
  const int16_t* fc = ws.coefs + static_cast<long>(frame) * g.nblk * 64;
  const int16_t* zz = fc + static_cast<long>(b) * 64;
  const unsigned off = ws.bits[static_cast<long>(frame) * g.nblk + b];
  const int comp = (b % 6) < 4 ? 0 : 1;
  const HuffTab& dc = T.dc[comp];
  const HuffTab& ac = T.ac[comp];
  // the bits in front of this block inside its first word belong to other blocks: they enter as zeros
  BitSink<OrWord> s{ws.raw + static_cast<long>(frame) * g.raw_words, off >> 5, off & 31, 0ull, or_word};
  int temp = dc_diff(fc, b), temp2 = temp;
  if (temp < 0) { temp = -temp; temp2--; }
  int nb = bit_length(static_cast<unsigned>(temp));
  s.put(dc.code[nb], dc.size[nb]);
  if (nb) s.put(static_cast<unsigned>(temp2), nb);
  int run = 0;
  for (int k = 1; k < 64; ++k) {
    temp = zz[k];
    if (temp == 0) { ++run; continue; }
    while (run > 15) { s.put(ac.code[0xF0], ac.size[0xF0]); run -= 16; }
    temp2 = temp;
    if (temp < 0) { temp = -temp; temp2--; }
    nb = bit_length(static_cast<unsigned>(temp));
    const int sym = (run << 4) + nb;
    s.put(ac.code[sym], ac.size[sym]);
    s.put(static_cast<unsigned>(temp2), nb);
    run = 0;
  }
  if (run > 0) s.put(ac.code[0], ac.size[0]);
  s.flush();
}

// ------------------------------------------------------------------------------------------------------------------
// pass 4: header + byte stuffing + EOI.  The frame's stream is walked in tiles of nthreads * 16 bytes; inside a
// tile thread t owns bytes [16t, 16t+16).
// ------------------------------------------------------------------------------------------------------------------
KRJ_HD unsigned stream_bytes(unsigned total_bits) { return (total_bits + 7) >> 3; }

struct Chunk16 {
  uint32_t w[4];
  KRJ_HD unsigned byte(int i) const { return (w[i >> 2] >> (8 * (i & 3))) & 0xFF; }
};
KRJ_HD Chunk16 load_chunk(const uint32_t* frame_raw, long byte0) {
  Chunk16 c;
#if defined(__CUDA_ARCH__)
  const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(frame_raw) + byte0);
  c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
#else
  for (int i = 0; i < 4; ++i) c.w[i] = frame_raw[byte0 / 4 + i];
#endif
  return c;
}
// number of 0xFF bytes among the thread's valid bytes
KRJ_HD unsigned stuff_count(const Chunk16& c, long byte0, unsigned nbytes) {
  unsigned cnt = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (byte0 + i < static_cast<long>(nbytes) && c.byte(i) == 0xFF) ++cnt;
  return cnt;
}
// write the thread's bytes (with their stuffed zeros) at out[dst...]; never beyond cap
KRJ_HD void stuff_write(const Chunk16& c, long byte0, unsigned nbytes, uint8_t* out, long dst, long cap) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (byte0 + i >= static_cast<long>(nbytes)) break;
    const unsigned v = c.byte(i);
    if (dst < cap) out[dst] = static_cast<uint8_t>(v);
    ++dst;
    if (v == 0xFF) {
      if (dst < cap) out[dst] = 0;
      ++dst;
    }
  }
}

}  // namespace krj
