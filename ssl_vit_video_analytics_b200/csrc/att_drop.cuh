// Shared by the attention kernels (include inside namespace vmae).
#pragma once
// Attention-probability dropout (nn.MultiheadAttention(dropout=p), mae_vit_adapter.py:40-47): a stateless hash of
// (seed, batch*head, query, key pair) gives two 16-bit uniforms, so the forward and both backward kernels regenerate the same
// keep mask without storing it. Element (q, k) is dropped when its half of the word is below thr = round(p * 65536).
struct AttnDrop {
  uint32_t seed_lo, seed_hi, thr;
  float rp;   // 1 / (1 - thr / 65536)
};
__device__ __forceinline__ uint32_t drop_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
constexpr uint32_t DROP_CQ = 0x9E3779B1u, DROP_CK = 0x85EBCA77u;
// word for query row id `rid` (= (b*H + h)*L + q) and key pair kp (= k >> 1): low half -> even k, high half -> odd k
__device__ __forceinline__ uint32_t drop_word(const AttnDrop& d, uint32_t rid, uint32_t kp) {
  return drop_mix32(rid * DROP_CQ + kp * DROP_CK + d.seed_lo) ^ d.seed_hi;
}
static inline AttnDrop make_attn_drop(float p_drop, unsigned long long seed) {
  AttnDrop d;
  d.seed_lo = static_cast<uint32_t>(seed);
  d.seed_hi = static_cast<uint32_t>(seed >> 32);
  d.thr = static_cast<uint32_t>(p_drop * 65536.0f + 0.5f);
  d.rp = 65536.0f / (65536.0f - static_cast<float>(d.thr));
  return d;
}
