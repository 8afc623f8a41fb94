// Host emulation of realtime_video_b200/csrc/kr_jpeg.cu — TEST INFRASTRUCTURE (tests/test_jpeg_emulation_cpu.py).
//
// The device encoder's per-thread bodies live in kr_jpeg_core.cuh and compile for the host as well.  This file
// replaces the four kernels' grids by loops (one iteration per thread, the CTA-wide prefix sums done serially) so the
// exact code the GPU threads run — index arithmetic, colour conversion, DCT, quantisation, bit offsets, the OR-ed bit
// stream, byte stuffing — is checked against Pillow / the oracle on a box without a GPU.  What it cannot cover: the
// launch configuration, the warp-shuffle scan and the atomics (tests/test_zz_jpeg_gpu.py does, on the B200).
//
// Build: g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -x c++ tests/jpeg_emulate.cpp
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../realtime_video_b200/csrc/kr_jpeg_core.cuh"

namespace {
const krj::Tables kTables = krj::make_tables();
}

extern "C" int jpeg_emulate_header_bytes() { return krj::kHeaderBytes; }

// kind 0: fp32 planar [frames, 3, H, W]; 1: RGB bytes [frames, H, W, 3].  frame_threads = CTA size of passes 2 / 4.
extern "C" int jpeg_emulate(const void* src, int kind, int frames, int H, int W, int quality, uint8_t* out, long cap,
                            int* sizes, int frame_threads, int16_t* coefs_out) {
  if (H % 16 || W % 16 || frames <= 0) return 1;
  const krj::Geometry g = krj::make_geometry(frames, H, W);
  std::vector<int16_t> coefs(static_cast<size_t>(frames) * g.nblk * 64, 0x7FFF);
  std::vector<uint32_t> bits(static_cast<size_t>(frames) * g.nblk, 0xDEADBEEF);
  std::vector<uint32_t> frame_bits(frames, 0xDEADBEEF);
  std::vector<uint32_t> raw(static_cast<size_t>(frames) * g.raw_words, 0xA5A5A5A5u);   // garbage: pass 2 must zero
  krj::Workspace ws{coefs.data(), bits.data(), frame_bits.data(), raw.data()};
  const krj::QuantTables qt = krj::make_quant(quality);
  const krj::Header hdr = krj::make_header(H, W, qt);
  if (hdr.w[krj::kHeaderWords - 1] != static_cast<uint32_t>(krj::kHeaderBytes)) return 2;

  // pass 1: grid (ceil(nblk / 128), frames) x 128
  for (int f = 0; f < frames; ++f)
    for (int idx = 0; idx < g.nblk; ++idx) {
      if (kind == 0) krj::dct_thread(krj::LoaderF32{static_cast<const float*>(src), H, W}, g, qt, kTables, f, idx, ws);
      else krj::dct_thread(krj::LoaderRgb8{static_cast<const uint8_t*>(src), H, W}, g, qt, kTables, f, idx, ws);
    }
  // pass 2: one CTA of frame_threads per frame
  for (int f = 0; f < frames; ++f) {
    std::vector<unsigned> mine(frame_threads);
    for (int t = 0; t < frame_threads; ++t) mine[t] = krj::scan_sum_thread(g, kTables, f, t, frame_threads, ws);
    unsigned run = 0;
    for (int t = 0; t < frame_threads; ++t) {
      krj::scan_write_thread(g, f, t, frame_threads, run, ws);
      run += mine[t];
    }
    ws.frame_bits[f] = run;
    uint32_t* fr = ws.raw + static_cast<long>(f) * g.raw_words;
    const unsigned used = krj::raw_words_used(run);
    if (used > static_cast<unsigned>(g.raw_words)) return 3;
    for (unsigned w = 0; w < used; ++w) fr[w] = krj::to_memory_order(krj::pad_word(run, w));
  }
  // pass 3: one thread per block, any order (run backwards to prove order independence)
  for (int f = frames - 1; f >= 0; --f)
    for (int b = g.nblk - 1; b >= 0; --b)
      krj::emit_thread(g, kTables, f, b, ws, [](uint32_t* word, uint32_t v) { *word |= v; });
  // pass 4: one CTA per frame, tiles of frame_threads * 16 bytes
  for (int f = 0; f < frames; ++f) {
    uint8_t* dst = out + static_cast<long>(f) * cap;
    const uint32_t* fr = ws.raw + static_cast<long>(f) * g.raw_words;
    const unsigned nbytes = krj::stream_bytes(ws.frame_bits[f]);
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(hdr.w);
    for (int i = 0; i < krj::kHeaderBytes; ++i)
      if (i < cap) dst[i] = hb[i];
    long running = krj::kHeaderBytes;
    const long tile = static_cast<long>(frame_threads) * 16;
    for (long tile0 = 0; tile0 < static_cast<long>(nbytes); tile0 += tile) {
      std::vector<unsigned> cnt(frame_threads, 0);
      std::vector<krj::Chunk16> ch(frame_threads);
      for (int t = 0; t < frame_threads; ++t) {
        const long byte0 = tile0 + 16L * t;
        if (byte0 < static_cast<long>(nbytes)) {
          ch[t] = krj::load_chunk(fr, byte0);
          cnt[t] = krj::stuff_count(ch[t], byte0, nbytes);
        }
      }
      unsigned excl = 0;
      for (int t = 0; t < frame_threads; ++t) {
        const long byte0 = tile0 + 16L * t;
        if (byte0 < static_cast<long>(nbytes)) krj::stuff_write(ch[t], byte0, nbytes, dst, running + 16L * t + excl, cap);
        excl += cnt[t];
      }
      const long tile_bytes = (static_cast<long>(nbytes) - tile0) < tile ? (static_cast<long>(nbytes) - tile0) : tile;
      running += tile_bytes + excl;
    }
    if (running < cap) dst[running] = 0xFF;
    if (running + 1 < cap) dst[running + 1] = 0xD9;
    const long size = running + 2;
    sizes[f] = size <= cap ? static_cast<int>(size) : -static_cast<int>(size);
  }
  if (coefs_out) std::memcpy(coefs_out, coefs.data(), coefs.size() * sizeof(int16_t));
  return 0;
}
