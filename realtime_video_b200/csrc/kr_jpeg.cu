// kr_jpeg.cu — device-side baseline JPEG encoder for the frame egress (second half).
//
// Replaces, byte for byte, the reference's host-side
//     TF.to_pil_image(frames[0, idx], "RGB").save(io, format='JPEG', quality=90)        release_server.py:973
// (and the normalisation in front of it, release_server.py:979-983) for a whole block of frames: four launches,
// bodies in kr_jpeg_core.cuh.  HBM-bound integer/byte work: per 12-frame 832x480 block 57.5 MB of fp32 pixels are
// read once by the luma threads and twice more (from L2) by the chroma threads, 14.4 MB of coefficients are written
// and read twice, and ~1-3 MB of JPEG bytes leave the device instead of 57.5 MB (fp32) / 14.4 MB (RGB8).
//   pass 1  jpeg_dct_kernel    one thread per 8x8 block, no shared memory (each thread streams its own pixel rows)
//   pass 2  jpeg_scan_kernel   one 1024-thread CTA per frame: bits per block -> exclusive prefix sum; zeroes the
//                              part of the bit stream pass 3 will OR into and sets the final padding bits
//   pass 3  jpeg_emit_kernel   one thread per block: Huffman codes OR-ed (atomicOr, words shared with neighbours)
//   pass 4  jpeg_stuff_kernel  one 1024-thread CTA per frame: header, 0xFF stuffing via a running prefix sum, EOI, size
#include "kr_common.cuh"
#include "kr_jpeg_core.cuh"
#include "kr_ops.h"

namespace kr {

namespace {

__constant__ krj::Tables c_tables = krj::make_tables();

constexpr int kBlockThreads = 128;     // passes 1 and 3
constexpr int kFrameThreads = 1024;    // passes 2 and 4

template <class Loader>
__global__ void __launch_bounds__(kBlockThreads)
jpeg_dct_kernel(const Loader ld, const krj::Geometry g, const krj::QuantTables qt, const krj::Workspace ws) {
  const int idx = blockIdx.x * kBlockThreads + threadIdx.x;
  if (idx >= g.nblk) return;
  krj::dct_thread(ld, g, qt, c_tables, static_cast<int>(blockIdx.y), idx, ws);
}

// exclusive prefix sum of one value per thread over a 1024-thread CTA; *total = sum over the CTA
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* smem33, unsigned* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();                       // smem33 may still be read from a previous call
  if (lane == 31) smem33[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned w = smem33[lane];     // kFrameThreads / 32 == 32 warps
    unsigned winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += n;
    }
    smem33[lane] = winc - w;             // exclusive offset of warp `lane`
    if (lane == 31) smem33[32] = winc;
  }
  __syncthreads();
  *total = smem33[32];
  return smem33[warp] + inc - v;
}

__global__ void __launch_bounds__(kFrameThreads)
jpeg_scan_kernel(const krj::Geometry g, const krj::Workspace ws) {
  __shared__ unsigned smem[33];
  const int frame = blockIdx.x;
  const unsigned mine = krj::scan_sum_thread(g, c_tables, frame, threadIdx.x, kFrameThreads, ws);
  unsigned total;
  const unsigned excl = block_exclusive_scan(mine, smem, &total);
  krj::scan_write_thread(g, frame, threadIdx.x, kFrameThreads, excl, ws);
  if (threadIdx.x == 0) ws.frame_bits[frame] = total;
  // zero the words pass 3 ORs into; the words holding the padding bits start from the padding pattern
  uint32_t* raw = ws.raw + static_cast<long>(frame) * g.raw_words;
  const unsigned used = krj::raw_words_used(total);
  for (unsigned w = threadIdx.x; w < used; w += kFrameThreads)
    raw[w] = krj::to_memory_order(krj::pad_word(total, w));
}

__global__ void __launch_bounds__(kBlockThreads)
jpeg_emit_kernel(const krj::Geometry g, const krj::Workspace ws) {
  const int b = blockIdx.x * kBlockThreads + threadIdx.x;
  if (b >= g.nblk) return;
  krj::emit_thread(g, c_tables, static_cast<int>(blockIdx.y), b, ws,
                   [](uint32_t* word, uint32_t v) { if (v) atomicOr(word, v); });
}

__global__ void __launch_bounds__(kFrameThreads)
jpeg_stuff_kernel(const krj::Geometry g, const krj::Workspace ws, const krj::Header hdr, uint8_t* out, long cap,
                  int* sizes) {
  __shared__ unsigned smem[33];
  const int frame = blockIdx.x;
  uint8_t* dst = out + static_cast<long>(frame) * cap;
  const uint32_t* raw = ws.raw + static_cast<long>(frame) * g.raw_words;
  const unsigned nbytes = krj::stream_bytes(ws.frame_bits[frame]);
  // header: 4-byte words (cap % 4 == 0 and out 4-byte aligned: checked by the launcher), tail bytes singly
  for (int i = threadIdx.x; i < krj::kHeaderBytes / 4; i += kFrameThreads)
    if (4L * i + 4 <= cap) reinterpret_cast<uint32_t*>(dst)[i] = hdr.w[i];
  if (threadIdx.x < krj::kHeaderBytes % 4) {
    const int i = krj::kHeaderBytes / 4 * 4 + threadIdx.x;
    if (i < cap) dst[i] = static_cast<uint8_t>(hdr.w[i >> 2] >> (8 * (i & 3)));
  }
  long running = krj::kHeaderBytes;       // output position of the current tile's first byte
  constexpr long kTile = static_cast<long>(kFrameThreads) * 16;
  for (long tile0 = 0; tile0 < static_cast<long>(nbytes); tile0 += kTile) {
    const long byte0 = tile0 + 16L * threadIdx.x;
    krj::Chunk16 c;
    unsigned cnt = 0;
    const bool live = byte0 < static_cast<long>(nbytes);
    if (live) {
      c = krj::load_chunk(raw, byte0);
      cnt = krj::stuff_count(c, byte0, nbytes);
    }
    unsigned tile_ff;
    const unsigned excl = block_exclusive_scan(cnt, smem, &tile_ff);
    if (live) krj::stuff_write(c, byte0, nbytes, dst, running + 16L * threadIdx.x + excl, cap);
    const long tile_bytes = (static_cast<long>(nbytes) - tile0) < kTile ? (static_cast<long>(nbytes) - tile0) : kTile;
    running += tile_bytes + tile_ff;
  }
  if (threadIdx.x == 0) {
    if (running < cap) dst[running] = 0xFF;
    if (running + 1 < cap) dst[running + 1] = 0xD9;
    const long size = running + 2;
    sizes[frame] = size <= cap ? static_cast<int>(size) : -static_cast<int>(size);
  }
}

size_t align256(size_t n) { return (n + 255) / 256 * 256; }

struct Carve {
  size_t coefs, bits, frame_bits, raw, total;
};
Carve carve(const krj::Geometry& g) {
  Carve c;
  size_t off = 0;
  c.coefs = off; off += align256(static_cast<size_t>(g.frames) * g.nblk * 64 * sizeof(int16_t));
  c.bits = off; off += align256(static_cast<size_t>(g.frames) * g.nblk * sizeof(uint32_t));
  c.frame_bits = off; off += align256(static_cast<size_t>(g.frames) * sizeof(uint32_t));
  c.raw = off; off += align256(static_cast<size_t>(g.frames) * g.raw_words * sizeof(uint32_t));
  c.total = off;
  return c;
}

bool shape_ok(int frames, int height, int width) {
  return frames > 0 && height > 0 && width > 0 && height % 16 == 0 && width % 16 == 0 && height < 65536 &&
         width < 65536 && frames < 65536;
}

}  // namespace

size_t jpeg_workspace_bytes(int frames, int height, int width) {
  if (!shape_ok(frames, height, width)) return 0;
  return carve(krj::make_geometry(frames, height, width)).total;
}

// src_kind 0: fp32 planar [frames, 3, H, W] in [-1, 1]; 1: packed RGB bytes [frames, H, W, 3]
int frames_to_jpeg(const void* src, int src_kind, int frames, int height, int width, int quality, uint8_t* out,
                   long cap, int* sizes, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!shape_ok(frames, height, width)) {
    set_last_error("frames_to_jpeg: frames=%d H=%d W=%d unsupported (H and W must be multiples of 16: pixels = 8 x "
                   "latent with even latent dimensions on every hot-path resolution)", frames, height, width);
    return KR_ERR_INVALID_ARG;
  }
  if (quality < 1 || quality > 100) {
    set_last_error("frames_to_jpeg: quality %d outside 1..100", quality);
    return KR_ERR_INVALID_ARG;
  }
  const krj::Geometry g = krj::make_geometry(frames, height, width);
  const Carve cv = carve(g);
  if (workspace == nullptr || workspace_bytes < cv.total || reinterpret_cast<uintptr_t>(workspace) % 256 != 0) {
    set_last_error("frames_to_jpeg: workspace of %zu bytes (256-byte aligned) needed, got %zu", cv.total,
                   workspace_bytes);
    return KR_ERR_INVALID_ARG;
  }
  if (cap < krj::kHeaderBytes + 2 || cap % 4 != 0 || cap > 0x7FFFFFF0L || reinterpret_cast<uintptr_t>(out) % 4 != 0) {
    set_last_error("frames_to_jpeg: per-frame capacity %ld must be a multiple of 4, >= %d and < 2 GiB; out 4-byte "
                   "aligned", cap, krj::kHeaderBytes + 2);
    return KR_ERR_INVALID_ARG;
  }
  const uintptr_t align = src_kind == 0 ? 16 : 8;
  if (reinterpret_cast<uintptr_t>(src) % align != 0) {
    set_last_error("frames_to_jpeg: source must be %d-byte aligned", static_cast<int>(align));
    return KR_ERR_INVALID_ARG;
  }
  uint8_t* base = static_cast<uint8_t*>(workspace);
  krj::Workspace ws;
  ws.coefs = reinterpret_cast<int16_t*>(base + cv.coefs);
  ws.bits = reinterpret_cast<uint32_t*>(base + cv.bits);
  ws.frame_bits = reinterpret_cast<uint32_t*>(base + cv.frame_bits);
  ws.raw = reinterpret_cast<uint32_t*>(base + cv.raw);
  const krj::QuantTables qt = krj::make_quant(quality);
  const krj::Header hdr = krj::make_header(height, width, qt);
  if (hdr.w[krj::kHeaderWords - 1] != static_cast<uint32_t>(krj::kHeaderBytes)) {
    set_last_error("frames_to_jpeg: internal header length mismatch");
    return KR_ERR_INVALID_ARG;
  }
  const dim3 per_block((g.nblk + kBlockThreads - 1) / kBlockThreads, frames);
  if (src_kind == 0) {
    const krj::LoaderF32 ld{static_cast<const float*>(src), height, width};
    jpeg_dct_kernel<<<per_block, kBlockThreads, 0, stream>>>(ld, g, qt, ws);
  } else {
    const krj::LoaderRgb8 ld{static_cast<const uint8_t*>(src), height, width};
    jpeg_dct_kernel<<<per_block, kBlockThreads, 0, stream>>>(ld, g, qt, ws);
  }
  jpeg_scan_kernel<<<frames, kFrameThreads, 0, stream>>>(g, ws);
  jpeg_emit_kernel<<<per_block, kBlockThreads, 0, stream>>>(g, ws);
  jpeg_stuff_kernel<<<frames, kFrameThreads, 0, stream>>>(g, ws, hdr, out, cap, sizes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("frames_to_jpeg: launch failed: %s", cudaGetErrorString(e));
    return KR_ERR_CUDA;
  }
  return KR_OK;
}

}  // namespace kr
