// vlk_attn_fwd / vlk_attn_bwd: argument checks and the shape -> kernel map of the attention family.
//
//   Tq, Tk <= 64                     attention_pair.cu     two (batch, head) problems per tcgen05 tile; fwd + bwd;
//                                                          attention-probability dropout (Q-Former)
//   64 < Tk <= 272 and Tq >= 64      attention_tcgen05.cu  persistent warp-specialised forward (CLIP ViT-L/14, 257 tokens)
//   Tk > 272 (forward), T > 64 (bwd) attention_flash.cu    streaming tcgen05 forward / backward (GPT-2 pretraining, T = 1024)
//   Tq < 64 against Tk > 64          attention_simt.cu     CUDA-core forward (KV-cached decode rows)
//
// The map is a pure function of the shapes: no environment switches, no alternative implementations in the library.
#include "common.cuh"

namespace vlk {

int attn_fwd_simt(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, cudaStream_t stream, int q_row0);
int attn_mid_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                 long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs,
                 int causal, float scale, cudaStream_t stream);
int attn_pair_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                  unsigned int stream_id, cudaStream_t stream);
int attn_pair_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                  void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                  int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                  long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float dropout_p,
                  const unsigned long long* seed_state, unsigned int stream_id, cudaStream_t stream);
int attn_flash_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                   long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                   int o_rs, int causal, float scale, cudaStream_t stream, int q_rows);
int attn_flash_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                   int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                   long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float* delta,
                   cudaStream_t stream);

namespace {
constexpr int kPairMax = 64;    // both sequence lengths of the head-pair kernel
constexpr int kMidMaxKeys = 272;
inline bool mult8(long long a, long long b, long long c, long long d) { return !((a | b | c | d) & 7); }
}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq,
                            int Tk, long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs,
                            long long o_bs, int o_rs, int causal, float scale, float dropout_p,
                            const unsigned long long* seed_state, unsigned int stream_id, void* stream) {
    VLK_REQUIRE(q && k && v && o, VLK_ERR_INVALID_ARG, "vlk_attn_fwd: null pointer");
    VLK_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, VLK_ERR_INVALID_ARG, "vlk_attn_fwd: B=%d H=%d Tq=%d Tk=%d", B, H,
                Tq, Tk);
    VLK_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || seed_state), VLK_ERR_INVALID_ARG,
                "vlk_attn_fwd: dropout_p=%f needs a seed state", dropout_p);
    const bool small = Tq <= kPairMax && Tk <= kPairMax;
    VLK_REQUIRE(dropout_p == 0.f || small, VLK_ERR_UNSUPPORTED,
                "vlk_attn_fwd: attention dropout is only implemented for Tq, Tk <= 64 (the Q-Former shapes)");
    VLK_REQUIRE(mult8(q_rs, k_rs, v_rs, o_rs) && mult8(q_bs, k_bs, v_bs, o_bs), VLK_ERR_ALIGNMENT,
                "vlk_attn_fwd: strides must be multiples of 8 elements");
    VLK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), VLK_ERR_ALIGNMENT,
                "vlk_attn_fwd: 16B alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (small)
        return attn_pair_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal, scale,
                             dropout_p, seed_state, stream_id, s);
    if (Tk > kMidMaxKeys) {
        // whole 128-row query blocks on the tensor cores; a remainder of <= 8 rows (CLIP's 257th token) would waste a
        // 128-row tile per (batch, head): those rows go to the CUDA-core few-rows kernel
        const int tail = Tq % 128;
        const bool split = tail > 0 && tail <= 8 && Tq > 128 && Tk <= 288;
        int rc = attn_flash_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal, scale,
                                s, split ? Tq - tail : Tq);
        if (rc || !split) return rc;
        return attn_fwd_simt(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal, scale, s,
                             Tq - tail);
    }
    if (Tq >= 64 && Tk > 64)
        return attn_mid_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal, scale, s);
    return attn_fwd_simt(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal, scale, s,
                         0);
}

extern "C" int vlk_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                            const float* lse, void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk,
                            long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs,
                            long long o_bs, int o_rs, long long dq_bs, int dq_rs, long long dk_bs, int dk_rs,
                            long long dv_bs, int dv_rs, int causal, float scale, float* delta, float dropout_p,
                            const unsigned long long* seed_state, unsigned int stream_id, void* stream) {
    VLK_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && delta, VLK_ERR_INVALID_ARG,
                "vlk_attn_bwd: null pointer");
    VLK_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, VLK_ERR_INVALID_ARG, "vlk_attn_bwd: B=%d H=%d Tq=%d Tk=%d", B, H,
                Tq, Tk);
    VLK_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || seed_state), VLK_ERR_INVALID_ARG,
                "vlk_attn_bwd: dropout_p=%f needs a seed state", dropout_p);
    const bool small = Tq <= kPairMax && Tk <= kPairMax;
    VLK_REQUIRE(dropout_p == 0.f || small, VLK_ERR_UNSUPPORTED,
                "vlk_attn_bwd: attention dropout is only implemented for Tq, Tk <= 64 (the Q-Former shapes)");
    VLK_REQUIRE(mult8(q_rs, k_rs, v_rs, o_rs) && mult8(q_bs, k_bs, v_bs, o_bs) && mult8(dq_rs, dk_rs, dv_rs, 0) &&
                    mult8(dq_bs, dk_bs, dv_bs, 0),
                VLK_ERR_ALIGNMENT, "vlk_attn_bwd: strides must be multiples of 8 elements");
    VLK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o) && aligned16(dq) &&
                    aligned16(dk) && aligned16(dv),
                VLK_ERR_ALIGNMENT, "vlk_attn_bwd: 16B alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (small)
        return attn_pair_bwd(q, k, v, o, d_o, lse, dq, dk, dv, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs,
                             dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, causal, scale, dropout_p, seed_state, stream_id, s);
    return attn_flash_bwd(q, k, v, o, d_o, lse, dq, dk, dv, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs,
                          dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, causal, scale, delta, s);
}
