// kr_dit_block.cu — kr_dit_block_fwd: the whole CausalWanAttentionBlock forward (causal_model.py:440-492 with the
// self-attention of :218-397 and the cross-attention of model.py:171-228) as ONE C-ABI call (SURVEY.md 8b "minimum
// export set").  Host code only: it issues the same 14 launches, with the same arguments, as the per-op schedule of
// realtime_video_b200/dit.py (_block / _self_attention / _cross_attention) — tests/test_block_fwd_cpu.py proves the
// two launch sequences identical call by call.  What a caller saves is the host side: one FFI crossing and one
// argument struct per block instead of 14 crossings, 14 argument marshals and 8 temporary tensor allocations.
//
//   add_modulation -> ln_modulate -> GEMM(to_qkv; q,k -> scratch, v -> V-cache slot) -> RMSNorm+RoPE (q -> scratch,
//   k -> K-cache slot) -> attention -> GEMM(o, gate+residual -> x) -> LayerNorm(affine) -> GEMM(q) -> RMSNorm
//   -> attention(text K/V) -> GEMM(o, residual -> x) -> ln_modulate -> GEMM(ffn.0, GELU) -> GEMM(ffn.2, gate+residual -> x)
//
// Scope: bf16, fused to_qkv, one GPU, prompt K/V already cached (crossattn_cache is_init) — the steady state of the
// server loop.  Everything else (first pass of a prompt, FP8 tier, multi-GPU exchange) stays on the per-op entry points.
#include "../../include/krea_b200.h"
#include "kr_common.cuh"
#include "kr_ops.h"

#include <cmath>

namespace {

size_t align256(size_t n) { return (n + 255) / 256 * 256; }

struct Scratch {          // byte offsets inside the caller's workspace
  size_t emod, h, qk, rq, y, hid, total;
};
Scratch carve(long L, long D, long ffn, long frames) {
  Scratch s;
  size_t off = 0;
  s.emod = off; off += align256(static_cast<size_t>(frames) * 6 * D * 2);
  s.h = off; off += align256(static_cast<size_t>(L) * D * 2);          // LayerNorm output (reused three times)
  s.qk = off; off += align256(static_cast<size_t>(L) * 2 * D * 2);     // q | k of the fused projection
  s.rq = off; off += align256(static_cast<size_t>(L) * D * 2);         // normalised + rotated q, later the cross-attention q
  s.y = off; off += align256(static_cast<size_t>(L) * D * 2);          // attention output (self, then cross)
  s.hid = off; off += align256(static_cast<size_t>(L) * ffn * 2);      // FFN hidden
  s.total = off;
  return s;
}

}  // namespace

extern "C" {

size_t kr_dit_block_workspace_bytes(int L, int D, int ffn, int frames) {
  if (L <= 0 || D <= 0 || ffn <= 0 || frames <= 0) return 0;
  return carve(L, D, ffn, frames).total;
}

#define KR_BLK_REQUIRE(cond, msg)                              \
  do {                                                         \
    if (!(cond)) {                                             \
      kr::set_last_error("kr_dit_block_fwd: %s", msg);         \
      return KR_ERR_INVALID_ARG;                               \
    }                                                          \
  } while (0)
#define KR_BLK_CALL(expr)          \
  do {                             \
    const int rc_ = (expr);        \
    if (rc_ != KR_OK) return rc_;  \
  } while (0)

int kr_dit_block_fwd(const KrDitBlockParams* p, void* stream_) {
  KR_BLK_REQUIRE(p != nullptr, "null params");
  KR_BLK_REQUIRE(p->x && p->e0 && p->modulation && p->w_qkv && p->w_o && p->w_cq && p->w_co && p->w_ffn0 && p->w_ffn2,
                 "null tensor pointer");
  KR_BLK_REQUIRE(p->norm_q && p->norm_k && p->norm_cq && p->k_cache && p->v_cache && p->ck && p->cv, "null tensor pointer");
  KR_BLK_REQUIRE(p->L > 0 && p->D > 0 && p->ffn > 0 && p->heads > 0 && p->head_dim == 128 && p->heads * p->head_dim == p->D,
                 "bad geometry (head_dim must be 128, heads * head_dim == D)");
  KR_BLK_REQUIRE(p->rows_per_frame > 0 && p->L % p->rows_per_frame == 0 && p->frames == p->L / p->rows_per_frame,
                 "L must be frames * rows_per_frame");
  KR_BLK_REQUIRE((2 * p->D) % 256 == 0, "the split QKV output needs 2*D to be a multiple of 256");
  KR_BLK_REQUIRE(p->local_start >= 0 && p->local_end == p->local_start + p->L, "cache slot must hold exactly L rows");
  KR_BLK_REQUIRE(p->mask_mode == 0 || p->mask_mode == 1, "mask_mode must be 0 or 1");
  KR_BLK_REQUIRE(p->mask_mode == 1 || (p->attn_lo >= 0 && p->attn_lo < p->local_end), "attn_lo outside the cache prefix");
  KR_BLK_REQUIRE(!p->cross_attn_norm || (p->norm3_w && p->norm3_b), "cross_attn_norm needs norm3 weight and bias");
  const Scratch sc = carve(p->L, p->D, p->ffn, p->frames);
  KR_BLK_REQUIRE(p->workspace && p->workspace_bytes >= sc.total && reinterpret_cast<uintptr_t>(p->workspace) % 256 == 0,
                 "workspace too small or not 256-byte aligned (kr_dit_block_workspace_bytes)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L = p->L, D = p->D, fs = p->rows_per_frame;
  uint8_t* ws = static_cast<uint8_t*>(p->workspace);
  uint16_t* emod = reinterpret_cast<uint16_t*>(ws + sc.emod);
  uint16_t* h = reinterpret_cast<uint16_t*>(ws + sc.h);
  uint16_t* qk = reinterpret_cast<uint16_t*>(ws + sc.qk);
  uint16_t* rq = reinterpret_cast<uint16_t*>(ws + sc.rq);
  uint16_t* y = reinterpret_cast<uint16_t*>(ws + sc.y);
  uint16_t* hid = reinterpret_cast<uint16_t*>(ws + sc.hid);
  uint16_t* x = static_cast<uint16_t*>(p->x);
  uint16_t* k_slot = static_cast<uint16_t*>(p->k_cache) + static_cast<size_t>(p->local_start) * p->ld_cache;
  uint16_t* v_slot = static_cast<uint16_t*>(p->v_cache) + static_cast<size_t>(p->local_start) * p->ld_cache;
  // 1/sqrt(head_dim) rounded to fp32 like the per-op entry point receives it, then log2(e) (kr_attn_fwd)
  const float scale_log2 = static_cast<float>(1.0 / std::sqrt(128.0)) * 1.4426950408889634f;

  auto gemm = [&](int epi, const void* a, int lda, const void* w, int K, int N, const void* bias, void* out, int ldc,
                  const void* residual, int ldr, const void* gate, void* out2, int ldc2, int n_split) {
    kr::GemmParams g;
    g.out = out; g.bias = bias; g.residual = residual; g.gate = gate;
    g.M = L; g.N = N; g.K = K; g.ldc = ldc; g.ldr = ldr;
    g.gate_stride = gate ? 6 * D : 0;
    g.rows_per_gate = gate ? fs : 0;
    g.row_offset = 0;
    g.alpha = 1.0f;
    g.out2 = out2; g.ldc2 = ldc2; g.n_split = n_split;
    return kr::gemm_tn(0, epi, a, lda, w, K, g, stream, p->gemm_workspace, p->gemm_workspace_bytes);
  };

  // time modulation of this block: emod[f] = modulation + e0[f]  (causal_model.py:466)
  KR_BLK_CALL(kr::add_modulation(p->modulation, p->e0, p->lde0_frame, emod, p->frames, 6, D, stream));
  // ---- self-attention (causal_model.py:467-476) -------------------------------------------------------------
  KR_BLK_CALL(kr::ln_modulate(x, p->ldx, h, D, L, D, p->eps_block, nullptr, nullptr, emod, 6, 0, 1, fs, 0, stream));
  KR_BLK_CALL(gemm(kr::EPI_BIAS, h, D, p->w_qkv, D, 3 * D, p->b_qkv, qk, 2 * D, nullptr, 0, nullptr, v_slot, p->ld_cache,
                   2 * D));
  {
    kr::QkvPostParams q;
    q.q = qk; q.k = qk + D; q.v = nullptr;
    q.ldq = 2 * D; q.ldk = 2 * D; q.ldv = 0;
    q.wq = static_cast<const uint16_t*>(p->norm_q); q.wk = static_cast<const uint16_t*>(p->norm_k);
    q.q_out = rq; q.ldqo = D;
    q.k_out = k_slot; q.ldko = p->ld_cache;
    q.v_out = nullptr; q.ldvo = 0;
    q.rope = static_cast<const float2*>(p->rope);
    q.D = D; q.head_dim = p->head_dim; q.grid_h = p->grid_h; q.grid_w = p->grid_w; q.start_frame = p->start_frame;
    q.row_offset = 0;
    q.eps = p->eps_qk;
    q.peer_cols = 0;
    for (int i = 0; i < 8; ++i) q.q_peer[i] = q.k_peer[i] = q.v_peer[i] = nullptr;
    KR_BLK_CALL(kr::qkv_post(q, L, stream));
  }
  {
    kr::AttnParams a;
    a.out = y; a.ldo = D; a.Lq = L; a.heads = p->heads; a.scale_log2 = scale_log2;
    const uint16_t* kc = static_cast<const uint16_t*>(p->k_cache);
    const uint16_t* vc = static_cast<const uint16_t*>(p->v_cache);
    if (p->mask_mode == 1) {      // recompute branch (:305-348): keys = rows [0, L) under the block-causal rule
      a.Lkv = L; a.mask_mode = 1; a.block_len = p->block_len; a.window = p->window; a.pad_keys = p->pad_keys;
      KR_BLK_CALL(kr::attn_fwd(0, rq, D, kc, p->ld_cache, vc, p->ld_cache, a, stream));
    } else {                      // cache branch (:386-390): keys = rows [attn_lo, local_end)
      a.Lkv = p->local_end - p->attn_lo; a.mask_mode = 0; a.block_len = 0; a.window = 0; a.pad_keys = 0;
      KR_BLK_CALL(kr::attn_fwd(0, rq, D, kc + static_cast<size_t>(p->attn_lo) * p->ld_cache, p->ld_cache,
                               vc + static_cast<size_t>(p->attn_lo) * p->ld_cache, p->ld_cache, a, stream));
    }
  }
  KR_BLK_CALL(gemm(kr::EPI_BIAS_GATE_RES, y, D, p->w_o, D, D, p->b_o, x, p->ldx, x, p->ldx, emod + 2 * D, nullptr, 0, 0));
  // ---- cross-attention (causal_model.py:478-480, model.py:171-228) ------------------------------------------
  const uint16_t* hx = x;
  int ldh = p->ldx;
  if (p->cross_attn_norm) {
    KR_BLK_CALL(kr::ln_modulate(x, p->ldx, h, D, L, D, p->eps_norm3, p->norm3_w, p->norm3_b, nullptr, 0, 0, 1, 0, 0, stream));
    hx = h;
    ldh = D;
  }
  KR_BLK_CALL(gemm(kr::EPI_BIAS, hx, ldh, p->w_cq, D, D, p->b_cq, rq, D, nullptr, 0, nullptr, nullptr, 0, 0));
  KR_BLK_CALL(kr::rmsnorm_rows(rq, D, rq, D, p->norm_cq, L, D, p->eps_cross, stream));
  {
    kr::AttnParams a;
    a.out = y; a.ldo = D; a.Lq = L; a.Lkv = p->text_len; a.heads = p->heads; a.scale_log2 = scale_log2;
    a.mask_mode = 0; a.block_len = 0; a.window = 0; a.pad_keys = 0;
    KR_BLK_CALL(kr::attn_fwd(0, rq, D, p->ck, p->ld_ck, p->cv, p->ld_cv, a, stream));
  }
  KR_BLK_CALL(gemm(kr::EPI_BIAS_RES, y, D, p->w_co, D, D, p->b_co, x, p->ldx, x, p->ldx, nullptr, nullptr, 0, 0));
  // ---- FFN (causal_model.py:482-490) ------------------------------------------------------------------------
  KR_BLK_CALL(kr::ln_modulate(x, p->ldx, h, D, L, D, p->eps_block, nullptr, nullptr, emod, 6, 3, 4, fs, 0, stream));
  KR_BLK_CALL(gemm(kr::EPI_BIAS_GELU, h, D, p->w_ffn0, D, p->ffn, p->b_ffn0, hid, p->ffn, nullptr, 0, nullptr, nullptr, 0, 0));
  KR_BLK_CALL(gemm(kr::EPI_BIAS_GATE_RES, hid, p->ffn, p->w_ffn2, p->ffn, D, p->b_ffn2, x, p->ldx, x, p->ldx, emod + 5 * D,
                   nullptr, 0, 0));
  return KR_OK;
}

}  // extern "C"
