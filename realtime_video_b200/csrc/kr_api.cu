// kr_api.cu — the extern "C" surface declared in include/krea_b200.h.
#include "../../include/krea_b200.h"
#include "kr_common.cuh"
#include "kr_ops.h"

#include <cmath>


#define KR_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      kr::set_last_error("%s: %s", __func__, msg); \
      return KR_ERR_INVALID_ARG;              \
    }                                         \
  } while (0)

extern "C" {

int kr_version(void) { return 100; }
const char* kr_last_error(void) { return kr::last_error(); }

int kr_gemm_kernel_id(int epilogue, int M, int N, int K) { return kr::gemm_plan(epilogue, M, N, K) + 1; }
int kr_gemm_kernel_id_ws(int epilogue, int M, int N, int K, int have_workspace) {
  return kr::gemm_plan(epilogue, M, N, K, have_workspace != 0) + 1;
}
size_t kr_gemm_workspace_bytes(void) { return kr::gemm_sk_workspace_bytes(); }

int kr_gemm(int dtype, int epilogue, const void* a, int lda, const void* w, int ldw,
            const void* bias, void* out, int ldc, int M, int N, int K, const void* residual,
            int ldr, const void* gate, int gate_stride, int rows_per_gate, float alpha,
            void* out2, int ldc2, int n_split, int row_offset, void* stream) {
  return kr_gemm_ws(dtype, epilogue, a, lda, w, ldw, bias, out, ldc, M, N, K, residual, ldr, gate, gate_stride,
                    rows_per_gate, alpha, out2, ldc2, n_split, row_offset, nullptr, 0, stream);
}

int kr_gemm_ws(int dtype, int epilogue, const void* a, int lda, const void* w, int ldw,
               const void* bias, void* out, int ldc, int M, int N, int K, const void* residual,
               int ldr, const void* gate, int gate_stride, int rows_per_gate, float alpha,
               void* out2, int ldc2, int n_split, int row_offset, void* workspace, size_t workspace_bytes,
               void* stream) {
  KR_REQUIRE(a && w && out, "null a/w/out");
  KR_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  kr::GemmParams p;
  p.out = out; p.bias = bias; p.residual = residual; p.gate = gate;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldr = ldr; p.gate_stride = gate_stride;
  p.rows_per_gate = rows_per_gate; p.alpha = alpha;
  p.out2 = out2; p.ldc2 = ldc2; p.n_split = n_split; p.row_offset = row_offset;
  return kr::gemm_tn(dtype, epilogue, a, lda, w, ldw, p, static_cast<cudaStream_t>(stream), workspace, workspace_bytes);
}

int kr_fp8_quantize(const void* x, int ldx, int rows, int cols, void* q, int ldq, float* state, void* stream) {
  KR_REQUIRE(x && q && state, "null x/q/state");
  return kr::fp8_quantize(x, ldx, rows, cols, q, ldq, state, static_cast<cudaStream_t>(stream));
}

int kr_gemm_fp8(int epilogue, const void* a, int lda, const void* w, int ldw, const float* scale_a, float scale_w,
                const void* bias, void* out, int ldc, int M, int N, int K, const void* residual, int ldr,
                const void* gate, int gate_stride, int rows_per_gate, void* out2, int ldc2, int n_split, int row_offset,
                void* stream) {
  KR_REQUIRE(a && w && out, "null a/w/out");
  kr::GemmParams p;
  p.out = out; p.bias = bias; p.residual = residual; p.gate = gate;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldr = ldr; p.gate_stride = gate_stride;
  p.rows_per_gate = rows_per_gate; p.alpha = 1.0f;
  p.out2 = out2; p.ldc2 = ldc2; p.n_split = n_split; p.row_offset = row_offset;
  return kr::gemm_fp8_tn(epilogue, a, lda, w, ldw, p, scale_a, scale_w, static_cast<cudaStream_t>(stream));
}

int kr_attn_fwd(int dtype, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                void* out, int ldo, int Lq, int Lkv, int heads, float softmax_scale, int mask_mode,
                int block_len, int window, int pad_keys, void* stream) {
  KR_REQUIRE(q && k && v && out, "null q/k/v/out");
  KR_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  KR_REQUIRE(mask_mode == 0 || mask_mode == 1, "mask_mode must be 0 or 1");
  kr::AttnParams p;
  p.out = out; p.ldo = ldo; p.Lq = Lq; p.Lkv = Lkv; p.heads = heads;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.mask_mode = mask_mode; p.block_len = block_len; p.window = window; p.pad_keys = pad_keys;
  return kr::attn_fwd(dtype, q, ldq, k, ldk, v, ldv, p, static_cast<cudaStream_t>(stream));
}

int kr_t5_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int L,
               int heads, const void* bias_delta, const void* key_mask, void* stream) {
  KR_REQUIRE(q && k && v && out && bias_delta, "null q/k/v/out/bias");
  kr::T5AttnParams p;
  p.out = out; p.ldo = ldo; p.L = L; p.heads = heads; p.bias_delta = bias_delta; p.key_mask = key_mask;
  return kr::t5_attn(q, ldq, k, ldk, v, ldv, p, static_cast<cudaStream_t>(stream));
}

int kr_ln_modulate(const void* x, int ldx, void* out, int ldo, int rows, int D, float eps,
                   const void* w, const void* b, const void* mod, int mod_rows, int shift_idx,
                   int scale_idx, int rows_per_frame, int row_offset, void* stream) {
  KR_REQUIRE(x && out, "null x/out");
  return kr::ln_modulate(x, ldx, out, ldo, rows, D, eps, w, b, mod, mod_rows, shift_idx, scale_idx,
                         rows_per_frame, row_offset, static_cast<cudaStream_t>(stream));
}

int kr_qkv_norm_rope(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                     const void* wq, const void* wk, void* q_out, int ldqo, void* k_out, int ldko,
                     void* v_out, int ldvo, const void* rope, int rows, int D, int head_dim,
                     int grid_h, int grid_w, int start_frame, int row_offset, float eps, void* stream) {
  KR_REQUIRE(q && k && wq && wk && q_out && k_out, "null q/k/weights/outputs");
  KR_REQUIRE((v == nullptr) == (v_out == nullptr), "v and v_out must both be given or both null");
  kr::QkvPostParams p;
  p.q = static_cast<const uint16_t*>(q); p.k = static_cast<const uint16_t*>(k);
  p.v = static_cast<const uint16_t*>(v);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.wq = static_cast<const uint16_t*>(wq); p.wk = static_cast<const uint16_t*>(wk);
  p.q_out = static_cast<uint16_t*>(q_out); p.ldqo = ldqo;
  p.k_out = static_cast<uint16_t*>(k_out); p.ldko = ldko;
  p.v_out = static_cast<uint16_t*>(v_out); p.ldvo = ldvo;
  p.rope = static_cast<const float2*>(rope);
  p.D = D; p.head_dim = head_dim; p.grid_h = grid_h; p.grid_w = grid_w; p.start_frame = start_frame;
  p.row_offset = row_offset;
  p.eps = eps;
  p.peer_cols = 0;
  return kr::qkv_post(p, rows, static_cast<cudaStream_t>(stream));
}

int kr_qkv_norm_rope_p2p(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                         const void* wq, const void* wk, void* const* q_peer, int ldqo, void* const* k_peer,
                         int ldko, void* const* v_peer, int ldvo, int world, int peer_cols, const void* rope,
                         int rows, int D, int head_dim, int grid_h, int grid_w, int start_frame, int row_offset,
                         float eps, void* stream) {
  KR_REQUIRE(q && k && v && wq && wk && q_peer && k_peer && v_peer, "null q/k/v/weights/peer tables");
  KR_REQUIRE(world >= 1 && world <= 8, "world must be 1..8");
  KR_REQUIRE(peer_cols > 0 && peer_cols % head_dim == 0 && peer_cols * world == D, "peer_cols * world must equal D");
  kr::QkvPostParams p;
  p.q = static_cast<const uint16_t*>(q); p.k = static_cast<const uint16_t*>(k);
  p.v = static_cast<const uint16_t*>(v);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.wq = static_cast<const uint16_t*>(wq); p.wk = static_cast<const uint16_t*>(wk);
  p.q_out = nullptr; p.k_out = nullptr; p.v_out = nullptr;
  p.ldqo = ldqo; p.ldko = ldko; p.ldvo = ldvo;
  p.rope = static_cast<const float2*>(rope);
  p.D = D; p.head_dim = head_dim; p.grid_h = grid_h; p.grid_w = grid_w; p.start_frame = start_frame;
  p.row_offset = row_offset;
  p.eps = eps;
  p.peer_cols = peer_cols;
  for (int i = 0; i < 8; ++i) {
    p.q_peer[i] = i < world ? static_cast<uint16_t*>(q_peer[i]) : nullptr;
    p.k_peer[i] = i < world ? static_cast<uint16_t*>(k_peer[i]) : nullptr;
    p.v_peer[i] = i < world ? static_cast<uint16_t*>(v_peer[i]) : nullptr;
    if (i < world) KR_REQUIRE(p.q_peer[i] && p.k_peer[i] && p.v_peer[i], "null peer pointer");
  }
  return kr::qkv_post(p, rows, static_cast<cudaStream_t>(stream));
}

int kr_comm_scatter_rows(const void* src, int ld_src, void* const* dst_peer, int ld_dst, int rows, int cols,
                         int rows_per_peer, int world, void* stream) {
  KR_REQUIRE(src && dst_peer, "null src / peer table");
  return kr::p2p_scatter_rows(src, ld_src, dst_peer, ld_dst, rows, cols, rows_per_peer, world,
                              static_cast<cudaStream_t>(stream));
}

int kr_kv_roll(void* cache, int ld, int width, int dst_row, int src_row, int rows, void* stream) {
  KR_REQUIRE(cache, "null cache");
  return kr::kv_roll(cache, ld, width, dst_row, src_row, rows, static_cast<cudaStream_t>(stream));
}

int kr_rmsnorm(const void* x, int ldx, void* out, int ldo, const void* w, int rows, int D,
               float eps, void* stream) {
  KR_REQUIRE(x && out && w, "null x/out/w");
  return kr::rmsnorm_rows(x, ldx, out, ldo, w, rows, D, eps, static_cast<cudaStream_t>(stream));
}

int kr_add_modulation(const void* modulation, const void* e0, int lde0_frame, void* out,
                      int frames, int mod_rows, int D, void* stream) {
  KR_REQUIRE(modulation && e0 && out, "null pointer");
  return kr::add_modulation(modulation, e0, lde0_frame, out, frames, mod_rows, D,
                            static_cast<cudaStream_t>(stream));
}

int kr_activation(const void* x, void* y, size_t n, int kind, void* stream) {
  KR_REQUIRE(x && y, "null pointer");
  KR_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (silu) or 1 (gelu-tanh)");
  return kr::activation(x, y, n, kind, static_cast<cudaStream_t>(stream));
}

int kr_patchify(const void* x, long sc, long sf, long sh, long sw, void* out, int C, int F, int H,
                int W, void* stream) {
  KR_REQUIRE(x && out, "null pointer");
  return kr::patchify(x, sc, sf, sh, sw, out, C, F, H, W, static_cast<cudaStream_t>(stream));
}

int kr_unpatchify_x0(const void* head_out, int ldh, const void* xt, const double* sigma,
                     void* flow, void* x0, int C, int F, int H, int W, void* stream) {
  KR_REQUIRE(head_out && flow, "null pointer");
  KR_REQUIRE(x0 == nullptr || (xt && sigma), "x0 requested without xt/sigma");
  return kr::unpatchify_x0(head_out, ldh, xt, sigma, flow, x0, C, F, H, W,
                           static_cast<cudaStream_t>(stream));
}

int kr_vae_conv3d(int dtype, int cin, int n, const void* in, int t_in, const void* weight,
                  int w_rows, const void* bias, int cout, int T, int H, int W, int tile_w, int tile_h,
                  int kt, int kh, int kw, void* out_raw, long raw_pix, long raw_frame, void* out_norm,
                  long norm_pix, long norm_frame, const void* gamma, const void* residual,
                  long res_pix, long res_frame, float* out_pix, int sub2, void* stream) {
  KR_REQUIRE(in && weight, "null input/weight");
  KR_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  KR_REQUIRE(out_raw || out_norm || out_pix, "no output requested");
  kr::ConvParams p;
  p.T = T; p.H = H; p.W = W; p.TW = tile_w; p.TH = tile_h; p.KT = kt; p.KH = kh; p.KW = kw;
  p.cout = cout;
  p.out_raw = out_raw; p.raw_pix = raw_pix; p.raw_frame = raw_frame;
  p.out_norm = out_norm; p.norm_pix = norm_pix; p.norm_frame = norm_frame;
  p.out_pix = out_pix; p.bias = bias;
  p.residual = residual; p.res_pix = res_pix; p.res_frame = res_frame;
  p.gamma = gamma; p.norm_scale = sqrtf(static_cast<float>(cout)); p.sub2 = sub2;
  return kr::vae_conv(dtype, cin, n, in, t_in, weight, w_rows, p, static_cast<cudaStream_t>(stream));
}

int kr_vae_rmsnorm_silu(int dtype, const void* x, void* y, const void* gamma, long pixels, int C,
                        int do_silu, void* stream) {
  KR_REQUIRE(x && y && gamma, "null pointer");
  return kr::vae_rmsnorm_silu(dtype, x, y, gamma, pixels, C, do_silu, static_cast<cudaStream_t>(stream));
}

int kr_vae_upsample2x(const void* in, void* out, int T, int H, int W, int C, void* stream) {
  KR_REQUIRE(in && out, "null pointer");
  return kr::vae_upsample2x(in, out, T, H, W, C, static_cast<cudaStream_t>(stream));
}

int kr_vae_scale_input(int dtype, const void* z, long zt, long zc, long zh, long zw,
                       const void* mean, const void* inv_std, const void* w2, const void* b2,
                       void* out, int T, int H, int W, void* stream) {
  KR_REQUIRE(z && mean && inv_std && w2 && b2 && out, "null pointer");
  return kr::vae_scale_input(dtype, z, zt, zc, zh, zw, mean, inv_std, w2, b2, out, T, H, W,
                             static_cast<cudaStream_t>(stream));
}

int kr_softmax_rows(int dtype, const float* s, long ld, void* p, long ldo, int rows, int cols,
                    void* stream) {
  KR_REQUIRE(s && p, "null pointer");
  return kr::softmax_rows(dtype, s, ld, p, ldo, rows, cols, static_cast<cudaStream_t>(stream));
}

int kr_frames_to_rgb8(const float* pixels, uint8_t* rgb, int frames, int height, int width, void* stream) {
  KR_REQUIRE(pixels && rgb, "null pointer");
  return kr::frames_to_rgb8(pixels, rgb, frames, height, width, static_cast<cudaStream_t>(stream));
}

size_t kr_jpeg_workspace_bytes(int frames, int height, int width) {
  return kr::jpeg_workspace_bytes(frames, height, width);
}

int kr_frames_to_jpeg(const float* pixels, int frames, int height, int width, int quality, unsigned char* out,
                      long cap, int* sizes, void* workspace, size_t workspace_bytes, void* stream) {
  KR_REQUIRE(pixels && out && sizes, "null pointer");
  return kr::frames_to_jpeg(pixels, 0, frames, height, width, quality, out, cap, sizes, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

int kr_rgb8_to_jpeg(const unsigned char* rgb, int frames, int height, int width, int quality, unsigned char* out,
                    long cap, int* sizes, void* workspace, size_t workspace_bytes, void* stream) {
  KR_REQUIRE(rgb && out && sizes, "null pointer");
  return kr::frames_to_jpeg(rgb, 1, frames, height, width, quality, out, cap, sizes, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

}  // extern "C"
