// evaluator.h -- ciphertext-level CKKS evaluator and bootstrap on the B200 runtime.
//
// Restates, on device-resident ciphertexts, what the reference's CKKS_EVALUATOR and
// CKKS_BTS_CTX do on the host (paths under fhe-cmplr/rtlib/ant/):
//   src/util/ckks_evaluator.c            Add/Sub/Mul/Rescale/Rotate/Conjugate/... (a11 helpers)
//   src/util/ckks_bootstrap_context.c    Bootstrap_setup, Find_rot_indices, Rotate_iteration,
//                                        Coeffs_to_slots / Slots_to_coeffs, Eval_bootstrap
//   src/util/ckks_chebyshev.c            Eval_chebyshev_ps (Paterson-Stockmeyer on Chebyshev basis)
// Control flow and all FP64 set-up arithmetic follow the reference operation for operation,
// so every limb the GPU produces equals the reference's.  Exact ring operations between two
// approximate base conversions are batched/fused freely: canonical residues do not depend on
// the order of exact modular additions and multiplications.
#pragma once
#include <complex>
#include <map>
#include <memory>
#include <vector>

#include "context.h"

namespace ace {

// A ciphertext in HBM: c0 and c1 are laid out [nq Q limbs | np P limbs] (np = 0 or K).
struct Ct {
  u64*   c0    = nullptr;
  u64*   c1    = nullptr;
  u32    nq    = 0;
  u32    np    = 0;
  u32    cap   = 0;  // limbs allocated per polynomial
  double sf    = 0;  // CIPHERTEXT._scaling_factor
  u32    sfd   = 0;  // CIPHERTEXT._sf_degree
  u32    slots = 0;
  bool   empty() const { return c0 == nullptr; }
};

// CKKS_BOOT_PARAMS (include/util/ckks_bootstrap_context.h:289-304)
struct BtsFftParams {
  int32_t level_budget, layers_coll, layers_rem, num_rot, b, g, num_rot_rem, b_rem, g_rem;
};

// CKKS_BTS_PRECOM (ckks_bootstrap_context.h:307-319); plaintexts live in HBM
struct BtsPrecom {
  u32          slots = 0;
  BtsFftParams enc{}, dec{};
  // [step][index] -> device plaintext of (level[step] + K) limbs, nullptr where unused
  std::vector<std::vector<u64*>> c2s, s2c;
  std::vector<u32>               c2s_level, s2c_level;
};

class Evaluator {
 public:
  explicit Evaluator(Context* ctx) : c(ctx) {}
  // evaluator of a worker context: uses the bootstrap tables of `primary` (read-only)
  Evaluator(Context* ctx, Evaluator* primary) : c(ctx), shared_(primary) {}
  ~Evaluator();
  Context* c;

  // ---- memory
  void reserve(Ct& x, u32 nq, u32 np);  // buffers for nq+np limbs; contents undefined
  void release(Ct& x);
  void copy(Ct& dst, const Ct& src);    // Copy_ciphertext
  void move(Ct& dst, Ct& src);          // dst takes src's buffers; src becomes empty

  // ---- ckks_evaluator.c
  void add(Ct& res, Ct& a, Ct& b);                       // Add_ciphertext :46-75
  void sub(Ct& res, Ct& a, Ct& b);                       // Sub_ciphertext :77-101
  void add_const(Ct& res, Ct& a, double v);              // Add_const :120-131
  void add_const_sfd(Ct& res, Ct& a, double v, u32 sfd); // Add_plaintext of a constant plaintext
  void mul_const(Ct& res, Ct& a, double v);              // Mul_const :208-217
  void mul_integer(Ct& res, Ct& a, u32 power);           // Mul_integer :219-232
  void mul_monomial(Ct& res, Ct& a, u32 power);          // Mul_by_monomial :234-264
  void mul(Ct& res, Ct& a, Ct& b);                       // Mul_ciphertext :167-179 (relin key)
  void rescale(Ct& res, Ct& a);                          // Rescale_ciphertext :324-343
  void rotate(Ct& res, Ct& a, int32_t rot);              // Eval_fast_rotate :527-535
  void conjugate(Ct& res, Ct& a);                        // Conjugate :579-604

  // ---- ckks_bootstrap_context.c / cipher_eval.c:366-404
  static u32 bootstrap_depth(size_t hamming_weight);     // Get_bootstrap_depth, budget {3,3}
  bool       bootstrap_supported() const;                // Get_mult_depth > bts_depth
  void       bootstrap_setup(u32 slots);                 // Bootstrap_setup :1050-1192
  std::vector<int32_t> bootstrap_rot_indices(u32 slots); // Find_rot_indices :280-299
  void       bootstrap(Ct& res, Ct& in, u32 level_after_bts);  // Bootstrap + Eval_bootstrap

  // Coeffs_to_slots / Slots_to_coeffs alone (:1494-1504) and read access to the diagonal tables
  void linear_transform(Ct& res, Ct& in, bool encoding);
  const u64* diagonal_plain(u32 slots, bool encoding, u32 step, u32 idx, u32* level);

  // collapsed FFT diagonals of C2S/S2C as the set-up computes them (host FP64 only)
  static size_t fft_diagonals(u32 slots, u32 budget, bool flag, bool encoding, double* out);

  // ---- ckks_chebyshev.c
  void eval_chebyshev(Ct& out, Ct& in, const std::vector<double>& coeffs, double a, double b);

 private:
  typedef std::complex<double>    cd;
  typedef std::vector<cd>         vcd;
  typedef std::vector<double>     vd;
  std::map<u32, std::unique_ptr<BtsPrecom>> precom_;
  Evaluator* shared_ = nullptr;
  std::map<u32, std::unique_ptr<BtsPrecom>>& precom() { return shared_ ? shared_->precom_ : precom_; }

  Basis basis(const Ct& x) const { return Basis{x.nq, x.np, (u32)c->L}; }
  const SwitchKey& rot_key(int32_t rot);

  // bootstrap set-up (host FP64)
  static BtsFftParams fft_params(u32 slots, u32 level_budget, u32 dim1);
  static std::vector<u32> select_layers(u32 log_slots, u32 budget);
  static std::vector<vcd> coeff_one_level(const vcd& ksi, const std::vector<u32>& rot_group,
                                          bool flag, bool encoding);
  static std::vector<std::vector<vcd>> coeff_collapse(const vcd& ksi,
                                                      const std::vector<u32>& rot_group,
                                                      u32 level_budget, bool flag,
                                                      bool encoding);
  void rotate_precomp(BtsPrecom& pc, std::vector<std::vector<vcd>>& coeffs, u32 slots,
                      double scale, u32 level, bool encoding);
  void find_coeffslots_rot_index(std::vector<int32_t>& out, const BtsFftParams& p, u32 slots,
                                 u32 m, bool encoding);

  // bootstrap evaluation
  void coeff_slots_transform(Ct& result, Ct& in, BtsPrecom& pc, bool encoding);
  void rotate_iteration(Ct& result, BtsPrecom& pc, const std::vector<std::vector<int32_t>>& rin,
                        const std::vector<std::vector<int32_t>>& rout, int32_t step,
                        bool encoding, bool is_rem);
  void apply_double_angle(Ct& x, u32 num_iter);
  void eval_approx_mod(Ct& out, Ct& in, const vd& coeffs);
  void eval_bootstrap(Ct& res, Ct& in, u32 raise_level, BtsPrecom& pc);

  // chebyshev
  void eval_linear_wsum(Ct& out, std::vector<Ct>& list, size_t size, const double* weights);
  void eval_quot_or_rem(Ct& out, std::vector<Ct>& t, const vd& qr, u32 k, bool is_quot,
                        bool in_rec);
  void inner_eval_ps(Ct& out, const vd& coeffs, u32 k, u32 m, std::vector<Ct>& t,
                     std::vector<Ct>& t2, bool in_rec);
};

}  // namespace ace
