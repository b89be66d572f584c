// refrng.h -- the reference's random sources, restated (host side, plain C++).
//
// Key generation and encryption of the reference draw from two generators:
//   * a BLAKE2Xb-based PRNG (fhe-cmplr/rtlib/ant/include/util/prng.h:42-90, src/util/prng.c:13-70):
//     buffers of 1024 32-bit words, buffer k = BLAKE2Xb(out = 4096 bytes, in = the 64-bit counter
//     k, key = the 64-byte seed); Uniform_uint_prng() rejects words above the largest multiple
//     of the range and divides.  Uniform residues (Sample_uniform, random_sample.c:38-76) and the
//     ternary secret (Sample_ternary, :99-152) come from it;
//   * glibc rand() (random_sample.c:20-36): Sample_triangle (:78-97) re-seeds it with
//     srand(time) and takes rand() % 4 per coefficient.
// To reproduce the reference's keys bit for bit from the same seeds the runtime
// needs the same streams consumed in the same order.  BLAKE2b / BLAKE2Xb are written from
// RFC 7693 and the BLAKE2X specification (the reference links the BLAKE2 team's ref code,
// fhe-cmplr/third-party/BLAKE2/ref); glibc's random() is the TYPE_3 additive feedback generator
// r[i] = r[i-3] + r[i-31] with the documented seeding.  The secure default sampler of the
// runtime does not use this file (client.cu).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace ace {
namespace refrng {

// ---------------------------------------------------------------- BLAKE2b (RFC 7693)
struct Blake2b {
  uint64_t h[8], t[2], f[2];
  uint8_t  buf[128];
  size_t   buflen, outlen;

  static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  static uint64_t load64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
  }
  static const uint64_t* iv() {
    static const uint64_t v[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                  0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                  0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    return v;
  }
  // the 64-byte parameter block (RFC 7693 2.5 / BLAKE2X: the 8-byte node offset of BLAKE2b holds
  // a 32-bit node offset and the 32-bit XOF length)
  void init_param(uint8_t digest_len, uint8_t key_len, uint8_t fanout, uint8_t depth, uint32_t leaf_len,
                  uint32_t node_offset, uint32_t xof_len, uint8_t node_depth, uint8_t inner_len) {
    uint8_t p[64] = {0};
    p[0] = digest_len; p[1] = key_len; p[2] = fanout; p[3] = depth;
    for (int i = 0; i < 4; i++) {
      p[4 + i]  = (uint8_t)(leaf_len >> (8 * i));
      p[8 + i]  = (uint8_t)(node_offset >> (8 * i));
      p[12 + i] = (uint8_t)(xof_len >> (8 * i));
    }
    p[16] = node_depth; p[17] = inner_len;
    for (int i = 0; i < 8; i++) h[i] = iv()[i] ^ load64(p + 8 * i);
    t[0] = t[1] = f[0] = f[1] = 0;
    buflen = 0;
    outlen = digest_len;
    memset(buf, 0, sizeof(buf));
  }
  void compress(const uint8_t* block) {
    static const uint8_t sigma[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = iv()[i]; }
    v[12] ^= t[0]; v[13] ^= t[1]; v[14] ^= f[0]; v[15] ^= f[1];
#define ACE_B2_G(r, i, a, b, c, d)                \
    a = a + b + m[sigma[r][2 * i]];     d = rotr(d ^ a, 32); \
    c = c + d;                          b = rotr(b ^ c, 24); \
    a = a + b + m[sigma[r][2 * i + 1]]; d = rotr(d ^ a, 16); \
    c = c + d;                          b = rotr(b ^ c, 63);
    for (int r = 0; r < 12; r++) {
      ACE_B2_G(r, 0, v[0], v[4], v[8], v[12])  ACE_B2_G(r, 1, v[1], v[5], v[9], v[13])
      ACE_B2_G(r, 2, v[2], v[6], v[10], v[14]) ACE_B2_G(r, 3, v[3], v[7], v[11], v[15])
      ACE_B2_G(r, 4, v[0], v[5], v[10], v[15]) ACE_B2_G(r, 5, v[1], v[6], v[11], v[12])
      ACE_B2_G(r, 6, v[2], v[7], v[8], v[13])  ACE_B2_G(r, 7, v[3], v[4], v[9], v[14])
    }
#undef ACE_B2_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
  }
  void update(const uint8_t* in, size_t len) {
    while (len > 0) {
      if (buflen == 128) {  // a full block that is known not to be the last one
        t[0] += 128;
        if (t[0] < 128) t[1]++;
        compress(buf);
        buflen = 0;
      }
      size_t take = 128 - buflen < len ? 128 - buflen : len;
      memcpy(buf + buflen, in, take);
      buflen += take; in += take; len -= take;
    }
  }
  void final(uint8_t* out) {
    t[0] += buflen;
    if (t[0] < buflen) t[1]++;
    f[0] = ~0ull;
    memset(buf + buflen, 0, 128 - buflen);
    compress(buf);
    for (size_t i = 0; i < outlen; i++) out[i] = (uint8_t)(h[i / 8] >> (8 * (i % 8)));
  }
};

// BLAKE2Xb with a key: out_len bytes (< 2^32 - 1) from `in` under `key` (key_len <= 64).  Root hash
// H0 = BLAKE2b(64 bytes; fanout 1, depth 1, XOF length = out_len); output block i =
// BLAKE2b(H0; fanout 0, depth 0, leaf length 64, node offset i, XOF length out_len, inner length 64).
inline void blake2xb(uint8_t* out, uint32_t out_len, const uint8_t* in, size_t in_len, const uint8_t* key,
                     uint8_t key_len) {
  Blake2b S;
  S.init_param(64, key_len, 1, 1, 0, 0, out_len, 0, 0);
  if (key_len) {
    uint8_t block[128] = {0};
    memcpy(block, key, key_len);
    S.update(block, 128);
  }
  S.update(in, in_len);
  uint8_t root[64];
  S.final(root);
  uint32_t left = out_len;
  for (uint32_t i = 0; left > 0; i++) {
    const uint8_t n = left < 64 ? (uint8_t)left : 64;
    Blake2b C;
    C.init_param(n, 0, 0, 0, 64, i, out_len, 0, 64);
    C.update(root, 64);
    C.final(out);
    out += n; left -= n;
  }
}

// ---------------------------------------------------------------- the reference's PRNG (prng.h)
struct Blake2Prng {
  static constexpr uint32_t kSeedWords = 16, kBufWords = 1024;  // SEED_CNT, PRNG_BUFFER_SIZE
  uint32_t seed[kSeedWords];
  uint64_t counter = 0;
  uint32_t buf[kBufWords];
  uint32_t idx = 0;  // _buffer_idx

  void pin(const uint32_t* seed16, uint64_t ctr) {
    memcpy(seed, seed16, sizeof(seed));
    counter = ctr;
    idx = 0;
  }
  uint32_t next() {  // Get_prng_value
    if (idx == kBufWords) idx = 0;
    if (idx == 0) {
      uint8_t in[8];
      for (int i = 0; i < 8; i++) in[i] = (uint8_t)(counter >> (8 * i));
      blake2xb(reinterpret_cast<uint8_t*>(buf), kBufWords * 4, in, 8, reinterpret_cast<const uint8_t*>(seed),
               kSeedWords * 4);
      counter++;
    }
    return buf[idx++];
  }
  uint32_t uniform_uint(uint32_t mn, uint32_t mx) {  // Uniform_uint_prng
    uint32_t range = mx - mn, ret;
    if (range < 0xFFFFFFFFu) {
      range += 1;
      const uint32_t scaling = 0xFFFFFFFFu / range, past = range * scaling;
      do ret = next(); while (ret >= past);
      ret /= scaling;
    } else {
      ret = next();
    }
    return ret + mn;
  }
  // Sample_uniform (random_sample.c:52-75): n residues below `bound`
  void sample_uniform(int64_t* out, size_t n, uint64_t bound) {
    uint32_t bits = 0;
    while ((bound >> (bits + 1)) != 0) bits++;         // (uint32_t)log2(bound)
    const uint32_t per = bits / 32, shift_chunk = per * 32;
    const uint32_t last_bound = (uint32_t)(bound >> shift_chunk);
    for (size_t i = 0; i < n; i++) {
      for (;;) {
        uint64_t r = 0;
        for (uint32_t k = 0, sh = 0; k < per; k++, sh += 32) r += (uint64_t)uniform_uint(0, 0xFFFFFFFFu) << sh;
        r += (uint64_t)uniform_uint(0, last_bound) << shift_chunk;
        if (r < bound) { out[i] = (int64_t)r; break; }
      }
    }
  }
  // Sample_ternary (random_sample.c:99-152)
  void sample_ternary(int64_t* out, size_t n, int64_t hw) {
    if (hw == 0) {
      for (size_t i = 0; i < n; i++) out[i] = (int64_t)(int32_t)uniform_uint((uint32_t)-1, 1);
      return;
    }
    if ((size_t)hw > n) hw = (int64_t)n;
    int32_t ones = 0;
    while (ones < hw / 2 - 1 || ones > hw / 2 + 1) {
      ones = 0;
      memset(out, 0, n * sizeof(int64_t));
      int64_t total = 0;
      while (total < hw) {
        const uint32_t at = uniform_uint(0, (uint32_t)n - 1);
        if (out[at] == 0) {
          if ((int32_t)uniform_uint(0, 1) == 0) out[at] = -1;
          else { out[at] = 1; ones++; }
          total++;
        }
      }
    }
  }
};

// The same stream for bulk consumers (key generation: ~9 M residues per switch key): buffers are
// independent of each other (the counter is the hash input), so the next kBatch of them are
// computed by several host threads at once; consumption stays sequential, as it must (a rejected
// word shifts everything behind it).
struct BulkPrng {
  static constexpr size_t kBatch = 2048;  // buffers per refill (8 MiB of words)
  uint32_t              seed[16];
  uint64_t              counter = 0;      // of the next buffer to generate
  std::vector<uint32_t> words;
  size_t                pos = 0;
  unsigned              threads = 8;

  void pin(const uint32_t* seed16, uint64_t ctr) {
    memcpy(seed, seed16, sizeof(seed));
    counter = ctr;
    words.clear();
    pos = 0;
  }
  void refill();  // defined in client.cu (std::thread)
  uint32_t next() {
    if (pos == words.size()) refill();
    return words[pos++];
  }
  uint32_t uniform_uint(uint32_t mn, uint32_t mx) {
    uint32_t range = mx - mn, ret;
    if (range < 0xFFFFFFFFu) {
      range += 1;
      const uint32_t scaling = 0xFFFFFFFFu / range, past = range * scaling;
      do ret = next(); while (ret >= past);
      ret /= scaling;
    } else {
      ret = next();
    }
    return ret + mn;
  }
  void sample_uniform(int64_t* out, size_t n, uint64_t bound) {
    uint32_t bits = 0;
    while ((bound >> (bits + 1)) != 0) bits++;
    const uint32_t per = bits / 32, shift_chunk = per * 32;
    const uint32_t last_bound = (uint32_t)(bound >> shift_chunk);
    const uint32_t range = last_bound + 1, scaling = last_bound == 0xFFFFFFFFu ? 1 : 0xFFFFFFFFu / range,
                   past = range * scaling;
    for (size_t i = 0; i < n; i++) {
      for (;;) {
        uint64_t r = 0;
        for (uint32_t k = 0, sh = 0; k < per; k++, sh += 32) r += (uint64_t)next() << sh;
        uint32_t hi;
        if (last_bound == 0xFFFFFFFFu) hi = next();
        else {
          do hi = next(); while (hi >= past);
          hi /= scaling;
        }
        r += (uint64_t)hi << shift_chunk;
        if (r < bound) { out[i] = (int64_t)r; break; }
      }
    }
  }
  void sample_ternary(int64_t* out, size_t n, int64_t hw) {
    if (hw == 0) {
      for (size_t i = 0; i < n; i++) out[i] = (int64_t)(int32_t)uniform_uint((uint32_t)-1, 1);
      return;
    }
    if ((size_t)hw > n) hw = (int64_t)n;
    int32_t ones = 0;
    while (ones < hw / 2 - 1 || ones > hw / 2 + 1) {
      ones = 0;
      memset(out, 0, n * sizeof(int64_t));
      int64_t total = 0;
      while (total < hw) {
        const uint32_t at = uniform_uint(0, (uint32_t)n - 1);
        if (out[at] == 0) {
          if ((int32_t)uniform_uint(0, 1) == 0) out[at] = -1;
          else { out[at] = 1; ones++; }
          total++;
        }
      }
    }
  }
};

// ---------------------------------------------------------------- glibc srandom() / random()
// TYPE_3: 31 words, r[i] = r[i-3] + r[i-31]; seeding: r[0] = seed, r[i] = 16807 r[i-1] mod (2^31-1)
// by Schrage's method, 310 values discarded; random() returns the new word >> 1.  rand() is
// random() in glibc.
struct GlibcRandom {
  int32_t r[31];
  int     f = 3, b = 0;
  void srandom(uint32_t seed) {
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
      const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      long word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      r[i] = (int32_t)word;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; i++) random();
  }
  int32_t random() {
    const uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
    r[f] = (int32_t)v;
    if (++f == 31) f = 0;
    if (++b == 31) b = 0;
    return (int32_t)(v >> 1);
  }
  // Sample_triangle (random_sample.c:78-97) after srand(seed): rand() % 4 -> -1, +1, 0, 0
  void sample_triangle(int64_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
      const int64_t x = random() % 4;
      out[i] = x == 0 ? -1 : (x == 1 ? 1 : 0);
    }
  }
};

}  // namespace refrng
}  // namespace ace
