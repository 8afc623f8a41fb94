/* CPU restatement of baseline JPEG encoding as the reference performs it on every outgoing frame —
 * TEST INFRASTRUCTURE ONLY (called by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by
 * the product path).
 *
 * Reference call site: release_server.py:973
 *     TF.to_pil_image(frames[0, idx], "RGB").save(io, format='JPEG', quality=90)
 * run in a 24-thread pool (release_server.py:945-946) for each of the 12 frames of a block.  The algorithm lives in
 * a third-party dependency that is not vendored in /root/reference: Pillow (12.2.0 in this image) -> its bundled
 * libjpeg-turbo (libjpeg API 6.2, features.version('jpg') == '6.2').  This file restates the published algorithm of
 * that library for exactly the parameters Pillow passes for `quality=90` on an 'RGB' image:
 *
 *   jpeg_set_defaults + jpeg_set_quality(q, force_baseline=TRUE)     jcparam.c  (Annex K tables, scaling)
 *   RGB -> YCbCr, 16-bit fixed point                                  jccolor.c  rgb_ycc_convert
 *   chroma 2x2 box downsample, alternating bias 1,2                   jcsample.c h2v2_downsample
 *   forward DCT "islow" (CONST_BITS 13, PASS1_BITS 2)                 jfdctint.c jpeg_fdct_islow
 *   quantisation: round-half-away division by (q << 3)               jcdctmgr.c forward_DCT
 *   baseline Huffman coding with the Annex K tables, 0xFF stuffing    jchuff.c   encode_one_block / emit_bits
 *   marker layout SOI APP0(JFIF 1.01) DQT DQT SOF0 DHTx4 SOS .. EOI   jcmarker.c
 *
 * Image dimensions must be multiples of 16 (every resolution of the hot path is: pixels = 8 x latent, latent
 * dimensions even), so libjpeg's edge replication / dummy blocks never occur.
 *
 * PINNED: tests/test_oracle_jpeg.py compares this file's output byte for byte with Pillow's for noise, gradients,
 * flat and extreme images at several sizes and qualities (Pillow is importable here and on the GPU box).
 */
#include <stdint.h>
#include <string.h>

static const uint8_t kStdLumaQ[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
    69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
    81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72, 92,  95,  98,  112, 100, 103, 99};
static const uint8_t kStdChromaQ[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
    99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
/* zigzag position k -> natural (row-major) index, jutils.c jpeg_natural_order */
static const uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

typedef struct {
  uint16_t code[256];
  uint8_t size[256];
} HuffTable;

/* jchuff.c jpeg_make_c_derived_tbl: canonical codes in order of increasing length */
static void derive(const uint8_t* bits, const uint8_t* vals, int nvals, HuffTable* t) {
  memset(t, 0, sizeof(*t));
  unsigned code = 0;
  int p = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < bits[len - 1] && p < nvals; ++i, ++p) {
      t->code[vals[p]] = (uint16_t)code;
      t->size[vals[p]] = (uint8_t)len;
      ++code;
    }
    code <<= 1;
  }
}

typedef struct {
  uint8_t* out;
  long cap, n;
  uint64_t acc; /* bit accumulator, newest bits at the low end */
  int nbits;
  int overflow;
} BitWriter;

static void put_byte(BitWriter* w, int b) {
  if (w->n < w->cap) w->out[w->n] = (uint8_t)b;
  else w->overflow = 1;
  w->n++;
}

/* jchuff.c emit_bits: MSB first; a 0xFF data byte is followed by a stuffed 0x00 */
static void emit_bits(BitWriter* w, unsigned code, int size) {
  if (size == 0) return;
  w->acc = (w->acc << size) | (code & ((1u << size) - 1u));
  w->nbits += size;
  while (w->nbits >= 8) {
    int b = (int)((w->acc >> (w->nbits - 8)) & 0xFF);
    put_byte(w, b);
    if (b == 0xFF) put_byte(w, 0);
    w->nbits -= 8;
  }
}

static int bit_length(int v) {
  int n = 0;
  while (v) { ++n; v >>= 1; }
  return n;
}

/* jchuff.c encode_one_block */
static int encode_block(BitWriter* w, const int16_t* zz, int last_dc, const HuffTable* dc, const HuffTable* ac) {
  int temp = zz[0] - last_dc, temp2 = temp;
  if (temp < 0) { temp = -temp; temp2--; }
  int nbits = bit_length(temp);
  emit_bits(w, dc->code[nbits], dc->size[nbits]);
  if (nbits) emit_bits(w, (unsigned)temp2, nbits);
  int r = 0;
  for (int k = 1; k < 64; ++k) {
    temp = zz[k];
    if (temp == 0) { ++r; continue; }
    while (r > 15) { emit_bits(w, ac->code[0xF0], ac->size[0xF0]); r -= 16; }
    temp2 = temp;
    if (temp < 0) { temp = -temp; temp2--; }
    nbits = bit_length(temp);
    int sym = (r << 4) + nbits;
    emit_bits(w, ac->code[sym], ac->size[sym]);
    emit_bits(w, (unsigned)temp2, nbits);
    r = 0;
  }
  if (r > 0) emit_bits(w, ac->code[0], ac->size[0]);
  return zz[0];
}

#define DESCALE(x, n) (((x) + (1L << ((n)-1))) >> (n))

/* jfdctint.c jpeg_fdct_islow, data = samples - 128; output scaled up by 8 */
static void fdct_islow(long* data) {
  const long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633,
             F1_501 = 12299, F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 8; ++i) {
      long* d = pass == 0 ? data + 8 * i : data + i;
      const int s = pass == 0 ? 1 : 8;
      long tmp0 = d[0] + d[7 * s], tmp7 = d[0] - d[7 * s];
      long tmp1 = d[1 * s] + d[6 * s], tmp6 = d[1 * s] - d[6 * s];
      long tmp2 = d[2 * s] + d[5 * s], tmp5 = d[2 * s] - d[5 * s];
      long tmp3 = d[3 * s] + d[4 * s], tmp4 = d[3 * s] - d[4 * s];
      long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
      const int sh = pass == 0 ? 13 - 2 : 13 + 2;
      if (pass == 0) {
        d[0] = (tmp10 + tmp11) << 2;
        d[4 * s] = (tmp10 - tmp11) << 2;
      } else {
        d[0] = DESCALE(tmp10 + tmp11, 2);
        d[4 * s] = DESCALE(tmp10 - tmp11, 2);
      }
      long z1 = (tmp12 + tmp13) * F0_541;
      d[2 * s] = DESCALE(z1 + tmp13 * F0_765, sh);
      d[6 * s] = DESCALE(z1 + tmp12 * (-F1_847), sh);
      z1 = tmp4 + tmp7;
      long z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7, z5 = (z3 + z4) * F1_175;
      tmp4 *= F0_298; tmp5 *= F2_053; tmp6 *= F3_072; tmp7 *= F1_501;
      z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
      z3 += z5; z4 += z5;
      d[7 * s] = DESCALE(tmp4 + z1 + z3, sh);
      d[5 * s] = DESCALE(tmp5 + z2 + z4, sh);
      d[3 * s] = DESCALE(tmp6 + z2 + z3, sh);
      d[1 * s] = DESCALE(tmp7 + z1 + z4, sh);
    }
  }
}

/* 8x8 block of plane (pitch) at (bx, by) in blocks -> quantised coefficients in zigzag order */
static void block_coefs(const uint8_t* plane, int pitch, int bx, int by, const uint16_t* q, int16_t* zz) {
  long d[64];
  for (int y = 0; y < 8; ++y)
    for (int x = 0; x < 8; ++x) d[8 * y + x] = (long)plane[(by * 8 + y) * pitch + bx * 8 + x] - 128;
  fdct_islow(d);
  for (int k = 0; k < 64; ++k) {
    const int i = kNatural[k];
    long t = d[i], qv = (long)q[i] << 3;
    if (t < 0) { t = -t; t += qv >> 1; t /= qv; t = -t; }
    else { t += qv >> 1; t /= qv; }
    zz[k] = (int16_t)t;
  }
}

static void quant_table(const uint8_t* basic, int quality, uint16_t* q) {
  if (quality <= 0) quality = 1;
  if (quality > 100) quality = 100;
  const long scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
  for (int i = 0; i < 64; ++i) {
    long t = ((long)basic[i] * scale + 50) / 100;
    if (t <= 0) t = 1;
    if (t > 255) t = 255; /* force_baseline */
    q[i] = (uint16_t)t;
  }
}

static void put_marker_dht(BitWriter* w, int tc_th, const uint8_t* bits, const uint8_t* vals, int nvals) {
  put_byte(w, 0xFF); put_byte(w, 0xC4);
  const int len = 2 + 1 + 16 + nvals;
  put_byte(w, len >> 8); put_byte(w, len & 0xFF);
  put_byte(w, tc_th);
  for (int i = 0; i < 16; ++i) put_byte(w, bits[i]);
  for (int i = 0; i < nvals; ++i) put_byte(w, vals[i]);
}

/* Number of header bytes written (SOI .. SOS), for callers that want the scan data offset. */
long jpeg_oracle_header(int H, int W, int quality, uint8_t* out, long cap) {
  BitWriter w = {out, cap, 0, 0, 0, 0};
  uint16_t q[2][64];
  quant_table(kStdLumaQ, quality, q[0]);
  quant_table(kStdChromaQ, quality, q[1]);
  put_byte(&w, 0xFF); put_byte(&w, 0xD8);
  /* APP0: JFIF 1.01, density 1:1, no thumbnail (jcmarker.c emit_jfif_app0) */
  static const uint8_t app0[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  for (unsigned i = 0; i < sizeof(app0); ++i) put_byte(&w, app0[i]);
  for (int t = 0; t < 2; ++t) {
    put_byte(&w, 0xFF); put_byte(&w, 0xDB); put_byte(&w, 0); put_byte(&w, 67); put_byte(&w, t);
    for (int k = 0; k < 64; ++k) put_byte(&w, q[t][kNatural[k]]);
  }
  put_byte(&w, 0xFF); put_byte(&w, 0xC0); put_byte(&w, 0); put_byte(&w, 17); put_byte(&w, 8);
  put_byte(&w, H >> 8); put_byte(&w, H & 0xFF); put_byte(&w, W >> 8); put_byte(&w, W & 0xFF);
  put_byte(&w, 3);
  put_byte(&w, 1); put_byte(&w, 0x22); put_byte(&w, 0);
  put_byte(&w, 2); put_byte(&w, 0x11); put_byte(&w, 1);
  put_byte(&w, 3); put_byte(&w, 0x11); put_byte(&w, 1);
  put_marker_dht(&w, 0x00, kDcLumaBits, kDcVals, 12);
  put_marker_dht(&w, 0x10, kAcLumaBits, kAcLumaVals, 162);
  put_marker_dht(&w, 0x01, kDcChromaBits, kDcVals, 12);
  put_marker_dht(&w, 0x11, kAcChromaBits, kAcChromaVals, 162);
  static const uint8_t sos[] = {0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
  for (unsigned i = 0; i < sizeof(sos); ++i) put_byte(&w, sos[i]);
  return w.overflow ? -w.n : w.n;
}

/* rgb: [H, W, 3] uint8.  scratch: caller-provided H*W*3/2 bytes (Y plane, then downsampled Cb, Cr).
 * Returns the number of bytes of the complete JPEG file, or -(bytes needed) if it does not fit into cap.
 * coefs (optional): int16 [blocks, 64] zigzag coefficients in scan order (Y00 Y01 Y10 Y11 Cb Cr per MCU). */
long jpeg_oracle_encode_rgb8(const uint8_t* rgb, int H, int W, int quality, uint8_t* scratch, int16_t* coefs,
                             uint8_t* out, long cap) {
  if (H <= 0 || W <= 0 || H % 16 || W % 16 || H > 65535 || W > 65535) return 0;
  uint8_t* Y = scratch;
  uint8_t* Cb = Y + (long)H * W;
  uint8_t* Cr = Cb + (long)(H / 2) * (W / 2);
  const int Wc = W / 2;
  /* jccolor.c rgb_ycc_convert (SCALEBITS 16) + jcsample.c h2v2_downsample */
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const uint8_t* p = rgb + ((long)y * W + x) * 3;
      Y[(long)y * W + x] = (uint8_t)((19595L * p[0] + 38470L * p[1] + 7471L * p[2] + 32768L) >> 16);
    }
  for (int y = 0; y < H / 2; ++y)
    for (int x = 0; x < Wc; ++x) {
      long sb = 0, sr = 0;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          const uint8_t* p = rgb + ((long)(2 * y + dy) * W + 2 * x + dx) * 3;
          sb += (-11059L * p[0] - 21709L * p[1] + 32768L * p[2] + (128L << 16) + 32767L) >> 16;
          sr += (32768L * p[0] - 27439L * p[1] - 5329L * p[2] + (128L << 16) + 32767L) >> 16;
        }
      const long bias = (x & 1) ? 2 : 1;
      Cb[(long)y * Wc + x] = (uint8_t)((sb + bias) >> 2);
      Cr[(long)y * Wc + x] = (uint8_t)((sr + bias) >> 2);
    }
  uint16_t q[2][64];
  quant_table(kStdLumaQ, quality, q[0]);
  quant_table(kStdChromaQ, quality, q[1]);
  HuffTable dcL, acL, dcC, acC;
  derive(kDcLumaBits, kDcVals, 12, &dcL);
  derive(kAcLumaBits, kAcLumaVals, 162, &acL);
  derive(kDcChromaBits, kDcVals, 12, &dcC);
  derive(kAcChromaBits, kAcChromaVals, 162, &acC);

  BitWriter w = {out, cap, 0, 0, 0, 0};
  w.n = jpeg_oracle_header(H, W, quality, out, cap);
  if (w.n < 0) { w.n = -w.n; w.overflow = 1; }
  int last[3] = {0, 0, 0};
  long blk = 0;
  int16_t zz[64];
  for (int my = 0; my < H / 16; ++my)
    for (int mx = 0; mx < W / 16; ++mx) {
      for (int b = 0; b < 4; ++b) {
        block_coefs(Y, W, 2 * mx + (b & 1), 2 * my + (b >> 1), q[0], zz);
        if (coefs) memcpy(coefs + 64 * blk, zz, sizeof(zz));
        ++blk;
        last[0] = encode_block(&w, zz, last[0], &dcL, &acL);
      }
      block_coefs(Cb, Wc, mx, my, q[1], zz);
      if (coefs) memcpy(coefs + 64 * blk, zz, sizeof(zz));
      ++blk;
      last[1] = encode_block(&w, zz, last[1], &dcC, &acC);
      block_coefs(Cr, Wc, mx, my, q[1], zz);
      if (coefs) memcpy(coefs + 64 * blk, zz, sizeof(zz));
      ++blk;
      last[2] = encode_block(&w, zz, last[2], &dcC, &acC);
    }
  emit_bits(&w, 0x7F, 7); /* jchuff.c flush_bits: pad the last byte with ones */
  w.acc = 0; w.nbits = 0;
  put_byte(&w, 0xFF); put_byte(&w, 0xD9);
  return w.overflow ? -w.n : w.n;
}
