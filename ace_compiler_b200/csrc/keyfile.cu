// keyfile.cu -- key and ciphertext files of the B200 runtime.
//
// The reference has no serialisation: every process regenerates its keys in Prepare_context
// (fhe-cmplr/rtlib/ant/src/rtlib/context.c:29-86; minutes for the 227 switch keys of ResNet-20).
// SURVEY 8(f1) asks for a format; this is it.  Everything is little-endian, limbs are the same
// canonical residues that sit in HBM (POLYNOMIAL._data layout: limb-major, Q limbs then P limbs).
//
//   header  : magic[8] "ACEB200K" (keys) / "ACEB200C" (ciphertext), u32 version = 1, u32 N,
//             u32 L, u32 K, u32 dnum, u32 flags, u64 modulus[L + K]
//   keys    : flags bit 0 = secret key present, bit 1 = public key present, bit 2 = relin key
//             [secret: (L+K) limbs] [public: L + L limbs] [relin: dnum x 2 x (L+K) limbs]
//             u32 n_rot, then per rotation key: u32 automorphism index, dnum x 2 x (L+K) limbs
//             (per digit: Pk0_at limbs, then Pk1_at limbs)
//   ct      : u32 level, u32 slots, u32 sf_degree, u32 pad, f64 scale, level limbs of c0, level of c1
// A file only loads into a context whose moduli are the ones recorded (same parameter set).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "context.h"

namespace ace {
namespace {

struct File {
  FILE* f;
  explicit File(const char* path, const char* mode) : f(fopen(path, mode)) {
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  }
  ~File() { if (f) fclose(f); }
  void put(const void* p, size_t n) {
    if (fwrite(p, 1, n, f) != n) throw std::runtime_error("short write");
  }
  void get(void* p, size_t n) {
    if (fread(p, 1, n, f) != n) throw std::runtime_error("truncated file");
  }
  template <class T> void put(const T& v) { put(&v, sizeof(T)); }
  template <class T> T    get() { T v; get(&v, sizeof(T)); return v; }
};

void put_header(File& F, const Context& c, const char* magic, u32 flags) {
  F.put(magic, 8);
  F.put<u32>(1); F.put<u32>(c.N); F.put<u32>((u32)c.L); F.put<u32>((u32)c.K); F.put<u32>((u32)c.dnum);
  F.put<u32>(flags);
  for (size_t g = 0; g < c.G; g++) F.put<u64>(c.mod[g]);
}
u32 get_header(File& F, const Context& c, const char* magic) {
  char m[8];
  F.get(m, 8);
  if (memcmp(m, magic, 8) != 0) throw std::runtime_error("not an ace_b200 file of this kind");
  if (F.get<u32>() != 1) throw std::runtime_error("unknown file version");
  const u32 N = F.get<u32>(), L = F.get<u32>(), K = F.get<u32>(), dnum = F.get<u32>(), flags = F.get<u32>();
  if (N != c.N || L != c.L || K != c.K || dnum != c.dnum) throw std::runtime_error("file is for another parameter set");
  for (size_t g = 0; g < c.G; g++)
    if (F.get<u64>() != c.mod[g]) throw std::runtime_error("file is for other moduli");
  return flags;
}

// device limbs <-> file, through one staging buffer
struct Mover {
  Context&         c;
  File&            F;
  std::vector<u64> host;
  void out(const u64* dev, size_t limbs) {
    host.resize(limbs * c.N);
    c.download(host.data(), dev, limbs);
    F.put(host.data(), host.size() * sizeof(u64));
  }
  void in(u64* dev, size_t limbs, const u64* moduli_of_limb /* may be null */, size_t g0) {
    host.resize(limbs * c.N);
    F.get(host.data(), host.size() * sizeof(u64));
    // residues must be canonical: a corrupt file must not poison the lazy arithmetic
    for (size_t l = 0; l < limbs; l++) {
      const u64 q = moduli_of_limb ? moduli_of_limb[l] : c.mod[g0 + l];
      const u64* p = host.data() + l * c.N;
      for (u32 i = 0; i < c.N; i++)
        if (p[i] >= q) throw std::runtime_error("file holds a residue outside its modulus");
    }
    c.upload(dev, host.data(), limbs);
  }
};

}  // namespace

void Context::save_keys(const char* path, bool with_secret) {
  ACE_CUDA(cudaSetDevice(device));
  File F(path, "wb");
  const u32 flags = ((with_secret && sk_ntt) ? 1u : 0u) | (pk0 ? 2u : 0u) | (relin_key.k0 ? 4u : 0u);
  put_header(F, *this, "ACEB200K", flags);
  Mover M{*this, F, {}};
  const size_t per = G;
  if (flags & 1) M.out(sk_ntt, G);
  if (flags & 2) { M.out(pk0, L); M.out(pk1, L); }
  auto put_swk = [&](const SwitchKey& k) {
    for (size_t j = 0; j < dnum; j++) {
      M.out(k.k0 + j * per * N, G);
      M.out(k.k1 + j * per * N, G);
    }
  };
  if (flags & 4) put_swk(relin_key);
  std::vector<u32> idx = rot_key_indices();
  std::sort(idx.begin(), idx.end());
  F.put<u32>((u32)idx.size());
  for (u32 k : idx) {
    F.put<u32>(k);
    put_swk(rot_key(k));
  }
}

void Context::load_keys(const char* path) {
  ACE_CUDA(cudaSetDevice(device));
  File F(path, "rb");
  const u32 flags = get_header(F, *this, "ACEB200K");
  Mover M{*this, F, {}};
  const size_t per = G * (size_t)N;
  if (flags & 1) {
    if (!sk_ntt) ACE_CUDA(cudaMalloc(&sk_ntt, per * sizeof(u64)));
    M.in(sk_ntt, G, nullptr, 0);
  }
  if (flags & 2) {
    if (!pk0) ACE_CUDA(cudaMalloc(&pk0, L * (size_t)N * sizeof(u64)));
    if (!pk1) ACE_CUDA(cudaMalloc(&pk1, L * (size_t)N * sizeof(u64)));
    M.in(pk0, L, nullptr, 0);
    M.in(pk1, L, nullptr, 0);
  }
  auto get_swk = [&](SwitchKey& k) {
    if (!k.k0) ACE_CUDA(cudaMalloc(&k.k0, dnum * per * sizeof(u64)));
    if (!k.k1) ACE_CUDA(cudaMalloc(&k.k1, dnum * per * sizeof(u64)));
    for (size_t j = 0; j < dnum; j++) {
      M.in(k.k0 + j * per, G, nullptr, 0);
      M.in(k.k1 + j * per, G, nullptr, 0);
    }
  };
  if (flags & 4) get_swk(relin_key);
  const u32 n_rot = F.get<u32>();
  if (n_rot > 4 * N) throw std::runtime_error("implausible number of rotation keys");
  for (u32 i = 0; i < n_rot; i++) {
    const u32 k = F.get<u32>();
    if (k >= 2 * N || !(k & 1)) throw std::runtime_error("bad automorphism index in key file");
    get_swk(rot_key(k));
  }
  sync();
}

void Context::save_ct(const char* path, const u64* c0, const u64* c1, u32 level, u32 slots, u32 sf_degree,
                      double scale) {
  ACE_CUDA(cudaSetDevice(device));
  File F(path, "wb");
  put_header(F, *this, "ACEB200C", 0);
  F.put<u32>(level); F.put<u32>(slots); F.put<u32>(sf_degree); F.put<u32>(0);
  F.put<double>(scale);
  Mover M{*this, F, {}};
  M.out(c0, level);
  M.out(c1, level);
}

void Context::load_ct(const char* path, u64* c0, u64* c1, u32 max_level, u32* level, u32* slots,
                      u32* sf_degree, double* scale) {
  ACE_CUDA(cudaSetDevice(device));
  File F(path, "rb");
  get_header(F, *this, "ACEB200C");
  const u32 lv = F.get<u32>(), sl = F.get<u32>(), sfd = F.get<u32>();
  F.get<u32>();
  const double sc = F.get<double>();
  if (lv == 0 || lv > L || lv > max_level) throw std::runtime_error("ciphertext level does not fit");
  Mover M{*this, F, {}};
  M.in(c0, lv, nullptr, 0);
  M.in(c1, lv, nullptr, 0);
  if (level) *level = lv;
  if (slots) *slots = sl;
  if (sf_degree) *sf_degree = sfd;
  if (scale) *scale = sc;
  sync();
}

}  // namespace ace
