// stdrng.hpp — the reference's seeded RNG stream, host side.
// rand 0.7.3 `StdRng` (Cargo.toml:12; third-party, not vendored) is rand_chacha's ChaCha20Rng:
//   seed_from_u64 -> PCG32 expansion of the u64 into a 256-bit key; 20-round ChaCha block function,
//   64-bit block counter starting at 0, stream id 0; u64 = two consecutive little-endian u32 words;
//   gen::<f64>() = (u64 >> 11) * 2^-53.   Call sites: src/lib.rs:38, src/swarm.rs:118.
#pragma once
#include <array>
#include <cstdint>

namespace lightdock {

class StdRng {
 public:
  static StdRng seed_from_u64(uint64_t state) {
    StdRng r;
    for (auto &word : r.key_) {
      state = state * 6364136223846793005ULL + 11634580027462260723ULL;
      const uint32_t xorshifted = static_cast<uint32_t>(((state >> 18) ^ state) >> 27);
      const uint32_t rot = static_cast<uint32_t>(state >> 59);
      word = (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    return r;
  }
  uint32_t next_u32() {
    if (pos_ == 16) refill();
    return block_[pos_++];
  }
  uint64_t next_u64() {
    const uint64_t lo = next_u32();
    const uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
  double gen_f64() { return static_cast<double>(next_u64() >> 11) * (1.0 / 9007199254740992.0); }

 private:
  static uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  static void quarter(std::array<uint32_t, 16> &s, int a, int b, int c, int d) {
    s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
    s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
    s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
    s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
  }
  void refill() {
    std::array<uint32_t, 16> in{};
    in[0] = 0x61707865u; in[1] = 0x3320646eu; in[2] = 0x79622d32u; in[3] = 0x6b206574u;
    for (int i = 0; i < 8; ++i) in[4 + i] = key_[i];
    in[12] = static_cast<uint32_t>(counter_);
    in[13] = static_cast<uint32_t>(counter_ >> 32);
    std::array<uint32_t, 16> w = in;
    for (int round = 0; round < 20; round += 2) {
      quarter(w, 0, 4, 8, 12); quarter(w, 1, 5, 9, 13); quarter(w, 2, 6, 10, 14); quarter(w, 3, 7, 11, 15);
      quarter(w, 0, 5, 10, 15); quarter(w, 1, 6, 11, 12); quarter(w, 2, 7, 8, 13); quarter(w, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) block_[i] = w[i] + in[i];
    ++counter_;
    pos_ = 0;
  }
  std::array<uint32_t, 8> key_{};
  std::array<uint32_t, 16> block_{};
  uint64_t counter_ = 0;
  int pos_ = 16;
};

}  // namespace lightdock
