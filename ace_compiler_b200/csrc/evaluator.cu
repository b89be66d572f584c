// evaluator.cu -- ciphertext-level evaluator (ckks_evaluator.c) on HBM-resident ciphertexts.
// See evaluator.h for the reference map.  Bootstrap lives in bootstrap.cu, the Chebyshev
// evaluation in chebyshev.cu.
#include "evaluator.h"

#include <cmath>
#include <cstring>

#include "host_math.h"

namespace ace {

Evaluator::~Evaluator() {
  for (auto& kv : precom_) {
    for (auto* tab : {&kv.second->c2s, &kv.second->s2c})
      for (auto& step : *tab)
        for (u64* p : step)
          if (p) cudaFree(p);
  }
}

// ------------------------------------------------------------------------------ memory
void Evaluator::reserve(Ct& x, u32 nq, u32 np) {
  if (x.cap < nq + np || x.c0 == nullptr) {
    release(x);
    x.c0  = c->alloc_limbs(nq + np, false);
    x.c1  = c->alloc_limbs(nq + np, false);
    x.cap = nq + np;
  }
  x.nq = nq;
  x.np = np;
}

void Evaluator::release(Ct& x) {
  if (x.c0) c->free_limbs(x.c0);
  if (x.c1) c->free_limbs(x.c1);
  x.c0 = x.c1 = nullptr;
  x.cap = 0;
}

void Evaluator::copy(Ct& dst, const Ct& src) {
  if (&dst == &src) return;
  reserve(dst, src.nq, src.np);
  size_t bytes = (size_t)(src.nq + src.np) * c->N * sizeof(u64);
  ACE_CUDA(cudaMemcpyAsync(dst.c0, src.c0, bytes, cudaMemcpyDeviceToDevice, c->stream));
  ACE_CUDA(cudaMemcpyAsync(dst.c1, src.c1, bytes, cudaMemcpyDeviceToDevice, c->stream));
  dst.sf = src.sf; dst.sfd = src.sfd; dst.slots = src.slots;
}

void Evaluator::move(Ct& dst, Ct& src) {
  if (&dst == &src) return;
  release(dst);
  dst = src;
  src.c0 = src.c1 = nullptr;
  src.cap = 0;
}

// ------------------------------------------------------------------------------ add / sub
// Result level = the smaller operand level (Adjust_level, ciphertext.h:283-319); when the
// result aliases an operand its level drops with it, exactly like the reference where the
// adjusted operand is only restored if it is not the result.
static void binary_meta(Evaluator* ev, Ct& res, Ct& a, Ct& b, u32& level) {
  if (a.np != b.np) throw std::runtime_error("add/sub: p primes do not match");
  level = a.nq < b.nq ? a.nq : b.nq;
  if (a.np && a.nq != b.nq) throw std::runtime_error("add/sub: extended operands differ in level");
  if (&res != &a && &res != &b) {
    const Ct& small = a.nq < b.nq ? a : b;  // smaller_ciph (ckks_evaluator.c:53-58)
    double sf = small.sf; u32 sfd = small.sfd, slots = small.slots;
    ev->reserve(res, level, a.np);
    res.sf = sf; res.sfd = sfd; res.slots = slots;
  }
}

void Evaluator::add(Ct& res, Ct& a, Ct& b) {
  u32 level;
  binary_meta(this, res, a, b, level);
  Basis bs{level, a.np, (u32)c->L};
  launch_ew_basis2(c->T, EW_ADD, res.c0, res.c1, a.c0, a.c1, b.c0, b.c1, bs, c->stream);
  c->tr(Context::TR_LIMB_ADD, 0, 2 * bs.width());
  c->launches += 1;
  res.nq = level;
}

void Evaluator::sub(Ct& res, Ct& a, Ct& b) {
  u32 level;
  binary_meta(this, res, a, b, level);
  if (a.np) throw std::runtime_error("sub: extended ciphertexts are not supported");
  Basis bs{level, 0, (u32)c->L};
  launch_ew_basis2(c->T, EW_SUB, res.c0, res.c1, a.c0, a.c1, b.c0, b.c1, bs, c->stream);
  c->tr(Context::TR_LIMB_ADD, 0, 2 * bs.width());
  c->launches += 1;
  res.nq = level;
}

// ------------------------------------------------------------------------------ constants
static void fill_pack(ScalarPack& sp, const std::vector<u64>& v, const std::vector<u64>& mod) {
  for (size_t i = 0; i < v.size(); i++) {
    sp.v[i]  = v[i];
    sp.sh[i] = hm::shoup(v[i], mod[i]);
  }
}

// Add_plaintext (ckks_evaluator.c:103-118) with the constant plaintext of Encode_val_at_level
// (ckks_encoder.c:464-528): in NTT form every coefficient of limb l is the same residue
void Evaluator::add_const_sfd(Ct& res, Ct& a, double v, u32 sfd) {
  if (a.np) throw std::runtime_error("add_const: extended ciphertext");
  ScalarPack sp;
  fill_pack(sp, c->value_residues(v, a.nq, sfd), c->mod);
  if (&res != &a) {
    reserve(res, a.nq, 0);
    res.sf = a.sf; res.sfd = a.sfd; res.slots = a.slots;
    ACE_CUDA(cudaMemcpyAsync(res.c1, a.c1, (size_t)a.nq * c->N * sizeof(u64),
                             cudaMemcpyDeviceToDevice, c->stream));
  }
  launch_add_scalar(c->T, res.c0, a.c0, sp, 0, a.nq, c->stream);
  c->tr(Context::TR_LIMB_ADD, 0, a.nq);
  c->launches++;
}

void Evaluator::add_const(Ct& res, Ct& a, double v) { add_const_sfd(res, a, v, a.sfd); }

void Evaluator::mul_const(Ct& res, Ct& a, double v) {
  if (a.np) throw std::runtime_error("mul_const: extended ciphertext");
  ScalarPack sp;
  fill_pack(sp, c->value_residues(v, a.nq, 1), c->mod);
  const double delta = (double)((u64)1 << c->params.scaling_mod_size);
  double sf = a.sf * pow(delta, 1);  // plain->_scaling_factor = pow(scaling_factor, sf_degree)
  u32 sfd = a.sfd + 1, slots = a.slots, nq = a.nq;
  if (&res != &a) reserve(res, nq, 0);
  Basis bs{nq, 0, (u32)c->L};
  launch_mul_scalar_pack2(c->T, res.c0, res.c1, a.c0, a.c1, sp, bs, c->stream);
  c->tr(Context::TR_LIMB_MUL, 0, 2 * bs.width());
  c->launches += 1;
  res.sf = sf; res.sfd = sfd; res.slots = slots;
}

void Evaluator::mul_integer(Ct& res, Ct& a, u32 power) {
  ScalarPack sp;
  Basis bs = basis(a);
  for (u32 y = 0; y < bs.width(); y++) {
    u64 q = c->mod[bs.g(y)];
    sp.v[y]  = (u64)power % q;
    sp.sh[y] = hm::shoup(sp.v[y], q);
  }
  if (&res != &a) {
    reserve(res, a.nq, a.np);
    res.sf = a.sf; res.sfd = a.sfd; res.slots = a.slots;
  }
  launch_mul_scalar_pack2(c->T, res.c0, res.c1, a.c0, a.c1, sp, bs, c->stream);
  c->tr(Context::TR_LIMB_MUL, 0, 2 * bs.width());
  c->launches += 1;
}

// Mul_by_monomial: multiply by X^power; the monomial is built in coefficient form
// (+1 at power mod N, or q-1 when power mod 2N >= N), transformed and multiplied in
void Evaluator::mul_monomial(Ct& res, Ct& a, u32 power) {
  if (a.np) throw std::runtime_error("mul_monomial: extended ciphertext");
  const u32 N = c->N, nq = a.nq;
  u64* mono = c->alloc_limbs(nq, false);
  launch_monomial(c->T, mono, power % N, (power % (2 * N)) >= N, nq, c->stream);
  c->ntt(mono, 0, nq);
  if (&res != &a) {
    reserve(res, nq, 0);
    res.sf = a.sf; res.sfd = a.sfd; res.slots = a.slots;
  }
  launch_ew(c->T, EW_MUL, res.c0, a.c0, mono, 0, nq, c->stream);
  launch_ew(c->T, EW_MUL, res.c1, a.c1, mono, 0, nq, c->stream);
  c->tr(Context::TR_LIMB_MUL, 0, 2 * nq);
  c->launches += 3;
  c->free_limbs(mono);
}

// ------------------------------------------------------------------------------ mul / rescale
void Evaluator::mul(Ct& res, Ct& a, Ct& b) {
  if (a.np || b.np) throw std::runtime_error("mul: extended ciphertext");
  const u32 level = a.nq < b.nq ? a.nq : b.nq;
  const Ct& small = a.nq < b.nq ? a : b;
  double sf = a.sf * b.sf;
  u32 sfd = a.sfd + b.sfd, slots = small.slots;
  Ct t;
  reserve(t, level, 0);
  c->ct_mul_relin(t.c0, t.c1, a.c0, a.c1, b.c0, b.c1, level);
  move(res, t);
  res.sf = sf; res.sfd = sfd; res.slots = slots;
}

void Evaluator::rescale(Ct& res, Ct& a) {
  if (a.np) throw std::runtime_error("rescale: extended ciphertext");
  if (a.nq < 2) throw std::runtime_error("rescale: multiply level is not big enough");
  const double delta = (double)((u64)1 << c->params.scaling_mod_size);
  double sf = a.sf / delta;
  u32 sfd = a.sfd - 1, slots = a.slots, nq = a.nq;
  Ct t;
  reserve(t, nq, 0);
  RescaleJob jobs[2] = {{t.c0, a.c0, nq}, {t.c1, a.c1, nq}};
  c->rescale_batch(jobs, 2);
  move(res, t);
  res.nq = nq - 1;
  res.sf = sf; res.sfd = sfd; res.slots = slots;
}

// ------------------------------------------------------------------------------ rotations
const SwitchKey& Evaluator::rot_key(int32_t rot) {
  u32 k = c->auto_index(rot);
  if (!c->has_rot_key(k)) throw std::runtime_error("cannot find auto key for rotation " + std::to_string(rot));
  return c->rot_key(k);
}

void Evaluator::rotate(Ct& res, Ct& a, int32_t rot) {
  if (a.np) throw std::runtime_error("rotate: extended ciphertext");
  double sf = a.sf; u32 sfd = a.sfd, slots = a.slots, nq = a.nq;
  Ct t;
  reserve(t, nq, 0);
  c->ct_rotate(t.c0, t.c1, a.c0, a.c1, nq, rot);
  move(res, t);
  res.sf = sf; res.sfd = sfd; res.slots = slots;
}

void Evaluator::conjugate(Ct& res, Ct& a) { rotate(res, a, (int32_t)(2 * c->N - 1)); }

}  // namespace ace
