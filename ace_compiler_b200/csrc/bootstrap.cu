// bootstrap.cu -- CKKS bootstrap (ModRaise -> CoeffsToSlots -> EvalMod -> SlotsToCoeffs) on the
// B200 runtime.  Restates fhe-cmplr/rtlib/ant/src/util/ckks_bootstrap_context.c:
//   set-up:   Select_layers :513-549, Get_colls_fft_params :551-610, Coeff_enc/dec_one_level
//             :419-511, Coeff_collapse :612-776, Coeffs2slots/Slots2coeffs_precomp :778-922,
//             Rotate_precomp :301-417, Bootstrap_setup :1050-1192, Find_rot_indices :214-299
//   evaluate: Rotate_iteration :1237-1381, Coeff_slots_transform :1383-1492,
//             Apply_double_angle_iterations :1512-1524, Eval_approx_mod :1553-1582,
//             Eval_bootstrap :1584-1860, Bootstrap (src/ckks/cipher_eval.c:366-404)
// The FP64 set-up runs on the host with the same operations in the same order as the reference
// (its plaintext tables must round to the same integers); the 6 x 63 diagonal plaintexts are
// encoded on the GPU and stay in HBM in the extended basis Q u P.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "evaluator.h"
#include "host_math.h"
#include "sin_coeffs.h"

namespace ace {

typedef std::complex<double>           cd;
typedef std::vector<cd>                vcd;
typedef std::vector<std::vector<vcd>>  vvvcd;

// complex product as the reference's C compiler evaluates it: (ac - bd, ad + bc), no FMA
static inline cd cmul(cd a, cd b) {
  return cd(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
static inline cd cscale(cd a, double s) { return cd(a.real() * s, a.imag() * s); }

// ------------------------------------------------------------------------------ sine polynomial
struct SinPoly { u32 k; u32 r; u32 n; const double* coeff; bool even_kind; };

static bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e != nullptr && atoi(e) != 0;
}

// Get_eval_sin_poly_info (:39-58); UNIFORM_TERNARY secrets
static SinPoly sin_poly(size_t hw) {
  const bool under = hw > 0 && hw <= 192;
  if (env_flag("RTLIB_BTS_EVEN_POLY"))
    return under ? SinPoly{32, 3, 55, kSinEvenHw192, true} : SinPoly{512, 7, 55, kSinEven, true};
  return under ? SinPoly{32, 3, 55, kSinUniformHw192, false}
               : SinPoly{512, 6, 89, kSinUniform, false};
}

// Get_depth_by_degree (:60-75) on a normalised series, plus the double-angle iterations
static u32 approx_mod_depth(size_t hw) {
  SinPoly sp  = sin_poly(hw);
  u32 deg     = sp.n - 1, d;
  if (deg <= 4) d = 3;
  else if (deg == 5) d = 4;
  else if (deg <= 13) d = 5;
  else if (deg <= 27) d = 6;
  else if (deg <= 59) d = 7;
  else if (deg <= 119) d = 8;
  else if (deg <= 247) d = 9;
  else if (deg <= 495) d = 10;
  else if (deg <= 1007) d = 11;
  else d = 12;
  return d - 1 + sp.r;
}

u32 Evaluator::bootstrap_depth(size_t hw) { return approx_mod_depth(hw) + 3 + 3; }
bool Evaluator::bootstrap_supported() const {
  return c->params.mul_depth > bootstrap_depth(c->params.hamming_weight);
}

// Reduce_rotation (:195-207)
static u32 reduce_rotation(int32_t index, u32 slots) {
  int32_t is = (int32_t)slots;
  if ((slots & (slots - 1)) == 0) {
    int32_t n = (int32_t)log2(slots);
    if (index >= 0) return index - ((index >> n) << n);
    return index + is + (((-index) >> n) << n);
  }
  return (is + index % is) % is;
}

// ------------------------------------------------------------------------------ FFT parameters
std::vector<u32> Evaluator::select_layers(u32 log_slots, u32 budget) {
  u32 layers = (u32)ceil((double)log_slots / budget);
  u32 rows = log_slots / layers, rem = log_slots % layers;
  u32 dim = rem != 0 ? rows + 1 : rows;
  if (dim < budget) {
    layers -= 1;
    rows = log_slots / layers;
    rem  = log_slots - rows * layers;
    dim  = rem != 0 ? rows + 1 : rows;
    while (dim != budget) {
      rows -= 1;
      rem = log_slots - rows * layers;
      dim = rem != 0 ? rows + 1 : rows;
    }
  }
  return {layers, rows, rem};
}

BtsFftParams Evaluator::fft_params(u32 slots, u32 level_budget, u32 dim1) {
  u32 log_slots = (u32)log2(slots);
  std::vector<u32> dims = select_layers(log_slots, level_budget);
  int32_t layers = (int32_t)dims[0], rem = (int32_t)dims[2];
  u32 num_rot = (1u << (layers + 1)) - 1, num_rot_rem = (1u << (rem + 1)) - 1;
  int32_t g;
  if (dim1 == 0 || dim1 > num_rot) g = num_rot > 7 ? (1 << (layers / 2 + 2)) : (1 << (layers / 2 + 1));
  else g = (int32_t)dim1;
  int32_t b = (int32_t)(num_rot + 1) / g, b_rem = 0, g_rem = 0;
  if (rem != 0) {
    g_rem = num_rot_rem > 7 ? (1 << (rem / 2 + 2)) : (1 << (rem / 2 + 1));
    b_rem = (int32_t)(num_rot_rem + 1) / g_rem;
  }
  return BtsFftParams{(int32_t)level_budget, layers, rem, (int32_t)num_rot, b, g,
                      (int32_t)num_rot_rem, b_rem, g_rem};
}

// Coeff_enc_one_level / Coeff_dec_one_level: rows [0,log) shifted right, [log,2log) unshifted,
// [2log,3log) shifted left
std::vector<vcd> Evaluator::coeff_one_level(const vcd& ksi, const std::vector<u32>& rot_group,
                                            bool flag, bool encoding) {
  const u32 dim = (u32)ksi.size() - 1, slots = (u32)rot_group.size();
  const u32 log_slots = (u32)log2(slots);
  std::vector<vcd> coeff(3 * log_slots, vcd(slots, cd(0, 0)));
  for (u32 m = slots; m > 1; m >>= 1) {
    u32  s  = (u32)log2(m) - 1;
    vcd& c0 = coeff[s];
    vcd& c1 = coeff[s + log_slots];
    vcd& c2 = coeff[s + 2 * log_slots];
    for (u32 k = 0; k < slots; k += m) {
      u32 lenh = m >> 1, lenq = m << 2;
      for (u32 j = 0; j < lenh; j++) {
        u32 tw = encoding ? (lenq - rot_group[j] % lenq) * (dim / lenq)
                          : (rot_group[j] % lenq) * (dim / lenq);
        if (flag && m == 2) {
          cd val = std::exp(encoding ? cd(-0.0, -M_PI / 2) : cd(0.0, M_PI / 2));
          cd w   = cmul(val, ksi[tw]);
          c1[j + k] = val;
          c1[j + k + lenh] = -w;
          if (encoding) { c2[j + k] = val; c0[j + k + lenh] = w; }
          else          { c2[j + k] = w;   c0[j + k + lenh] = val; }
        } else {
          cd w = ksi[tw];
          c1[j + k] = cd(1, 0);
          c1[j + k + lenh] = -w;
          if (encoding) { c2[j + k] = cd(1, 0); c0[j + k + lenh] = w; }
          else          { c2[j + k] = w;        c0[j + k + lenh] = cd(1, 0); }
        }
      }
    }
  }
  return coeff;
}

vvvcd Evaluator::coeff_collapse(const vcd& ksi, const std::vector<u32>& rot_group,
                                u32 level_budget, bool flag, bool encoding) {
  const u32 slots = (u32)rot_group.size(), log_slots = (u32)log2(slots);
  std::vector<u32> dims = select_layers(log_slots, level_budget);
  const int32_t layers = (int32_t)dims[0], rem = (int32_t)dims[2];
  const int32_t dim_coll = (int32_t)level_budget;
  const bool flag_rem = rem != 0;
  const u32 num_rot = (1u << (layers + 1)) - 1, num_rot_rem = (1u << (rem + 1)) - 1;

  std::vector<vcd> coeff1 = coeff_one_level(ksi, rot_group, flag, encoding);
  vvvcd coeff(dim_coll);
  for (int32_t i = 0; i < dim_coll; i++) {
    bool after = (encoding && i >= 1) || (!encoding && i < (int32_t)level_budget - 1);
    u32 rows = (flag_rem && !after) ? num_rot_rem : num_rot;
    coeff[i].assign(rows, vcd(slots, cd(0, 0)));
  }
  for (int32_t s = 0; s < dim_coll; s++) {
    int32_t top = encoding ? (int32_t)log_slots - (dim_coll - 1 - s) * layers - 1 : s * layers;
    bool is_rem = flag_rem && ((encoding && s == 0) || (!encoding && s == dim_coll - 1));
    int32_t end_l = is_rem ? rem : layers;
    for (int32_t l = 0; l < end_l; l++) {
      if (l == 0) {
        coeff[s][0] = coeff1[top];
        coeff[s][1] = coeff1[top + log_slots];
        coeff[s][2] = coeff1[top + 2 * log_slots];
        continue;
      }
      std::vector<vcd> temp(coeff[s].size(), vcd(slots, cd(0, 0)));
      if (encoding) {
        u32 t = 0;
        const vcd& a0 = coeff1[top - l];
        const vcd& a1 = coeff1[top - l + log_slots];
        const vcd& a2 = coeff1[top - l + 2 * log_slots];
        for (int32_t u = 0; u < (1 << (l + 1)) - 1; u++) {
          const vcd& tu = coeff[s][u];
          for (u32 k = 0; k < slots; k++) {
            u32 r1 = reduce_rotation((int32_t)k - (1 << (top - l)), slots);
            u32 r2 = reduce_rotation((int32_t)k + (1 << (top - l)), slots);
            temp[u + t][k]     += cmul(a0[k], tu[r1]);
            temp[u + t + 1][k] += cmul(a1[k], tu[k]);
            temp[u + t + 2][k] += cmul(a2[k], tu[r2]);
          }
          t += 1;
        }
      } else {
        for (u32 t = 0; t < 3; t++) {
          const vcd& a = coeff1[top + l + t * log_slots];
          const u32 shift = t == 0 ? 0 : (t == 1 ? (1u << l) : (1u << (l + 1)));
          for (int32_t u = 0; u < (1 << (l + 1)) - 1; u++) {
            const vcd& tu = coeff[s][u];
            vcd& dst = temp[u + shift];
            for (u32 k = 0; k < slots; k++) dst[k] += cmul(a[k], tu[k]);
          }
        }
      }
      coeff[s] = temp;
    }
  }
  return coeff;
}

// Rotate_precomp: rotate every diagonal by the giant-step offset, apply `scale` on one level,
// encode over Q u P.  The reference reads the *encode* parameter set in both directions.
void Evaluator::rotate_precomp(BtsPrecom& pc, vvvcd& coeffs, u32 slots, double scale, u32 level,
                               bool encoding) {
  const BtsFftParams& P = pc.enc;
  const u32 m = 2 * c->N, q_cnt = (u32)c->L, K = (u32)c->K;
  const int32_t budget = P.level_budget;
  const int32_t flag_rem = P.layers_rem != 0 ? 1 : 0, stop = flag_rem ? 0 : -1;
  auto& tab = encoding ? pc.c2s : pc.s2c;
  auto& lev = encoding ? pc.c2s_level : pc.s2c_level;
  tab.assign(budget, {});
  lev.assign(budget, 0);
  const u32 rem_index = encoding ? 0 : budget - 1;
  for (int32_t i = 0; i < budget; i++)
    tab[i].assign((flag_rem && (u32)i == rem_index) ? P.num_rot_rem : P.num_rot, nullptr);

  const int32_t start = encoding ? stop + 1 : 0;
  const int32_t end   = encoding ? budget : budget - flag_rem;
  const int32_t cond  = encoding ? start : end - 1;
  const u32 enc_level = level ? level + 1 : q_cnt - budget + 1;
  const u32 dec_level = level ? level + budget : q_cnt;
  vcd rot_vl(slots);
  auto encode_row = [&](vcd& vl, u32 rot, u32 plain_level, u64*& slot) {
    for (u32 idx = 0; idx < slots; idx++) rot_vl[idx] = vl[(idx + rot) % slots];
    u64* d = nullptr;
    ACE_CUDA(cudaMalloc(&d, (size_t)(plain_level + K) * c->N * sizeof(u64)));
    c->encode_cplx(d, rot_vl.data(), slots, plain_level, slots, 1, K);
    slot = d;
  };
  for (int32_t s = start; s < end; s++) {
    u32 plain_level = encoding ? enc_level + s : dec_level - s;
    lev[s] = plain_level;
    for (int32_t i = 0; i < P.b; i++) {
      for (int32_t j = 0; j < P.g; j++) {
        int32_t dim2 = P.g * i + j;
        if (dim2 == P.num_rot) continue;
        int32_t shift = encoding ? ((s - flag_rem) * P.layers_coll + P.layers_rem) : (s * P.layers_coll);
        u32 rot = reduce_rotation(-P.g * i * (1 << shift), m / 4);
        vcd& vl = coeffs[s][dim2];
        if (flag_rem == 0 && s == cond)
          for (cd& x : vl) x = cscale(x, scale);
        encode_row(vl, rot, plain_level, tab[s][dim2]);
      }
    }
  }
  if (flag_rem) {
    int32_t dim1  = encoding ? stop : budget - flag_rem;
    int32_t shift = encoding ? 1 : (1 << (dim1 * P.layers_coll));
    u32 plain_level = encoding ? enc_level : dec_level - budget + flag_rem;
    lev[dim1] = plain_level;
    for (int32_t i = 0; i < P.b_rem; i++) {
      for (int32_t j = 0; j < P.g_rem; j++) {
        int32_t dim2 = P.g_rem * i + j;
        if (dim2 == P.num_rot_rem) continue;
        u32 rot = reduce_rotation(-P.g_rem * i * shift, m / 4);
        vcd& vl = coeffs[dim1][dim2];
        for (cd& x : vl) x = cscale(x, scale);
        encode_row(vl, rot, plain_level, tab[dim1][dim2]);
      }
    }
  }
}

void Evaluator::bootstrap_setup(u32 num_slots) {
  const u32 N = c->N, m = 2 * N;
  const u32 slots = num_slots == 0 ? m / 4 : num_slots;
  if (precom().count(slots)) return;
  if (shared_) throw std::runtime_error("bootstrap tables for a new slot count must be set up on the primary context");
  if (!bootstrap_supported()) throw std::runtime_error("bootstrap: need a larger multiply depth");
  std::unique_ptr<BtsPrecom> pcp(new BtsPrecom);
  BtsPrecom& pc = *pcp;
  pc.slots = slots;
  u32 budget[2] = {3, 3};  // Bootstrap_precom (src/rtlib/context.c:162-185)
  double log_slots = log2(slots);
  for (u32& v : budget) {
    if (v > log_slots) v = (u32)log_slots;
    if (v < 1) v = 1;
  }
  pc.enc = fft_params(slots, budget[0], 0);
  pc.dec = fft_params(slots, budget[1], 0);
  if (pc.enc.level_budget == 1 && pc.dec.level_budget == 1)
    throw std::runtime_error("bootstrap: linear-transform mode is unimplemented (as in the reference)");

  const u32 slots4 = 4 * slots;
  const bool sparse = m != slots4;
  std::vector<u32> rot_group(slots);
  u32 five = 1;
  for (u32 i = 0; i < slots; i++) {
    rot_group[i] = five;
    five *= 5;
    five %= slots4;
  }
  vcd ksi(slots4 + 1);
  for (size_t i = 0; i < slots4; i++) {
    double angle = 2.0 * M_PI * i / slots4;
    ksi[i] = cd(cos(angle), sin(angle));
  }
  ksi[slots4] = ksi[0];

  const u64 q0 = c->mod[0];
  const double dbl_q0 = (double)q0;
  const double pow2 = ldexp(1.0, (int)(u32)round(log2(dbl_q0)));  // (UINT128)1 << round(log2 q0)
  const double pre = dbl_q0 / pow2;
  const double scale_enc = pre / 1.0, scale_dec = 1 / pre;

  const u32 bts_depth = approx_mod_depth(c->params.hamming_weight) + pc.enc.level_budget + pc.dec.level_budget;
  const u32 level_0 = (u32)c->params.mul_depth + 1;
  if (level_0 <= (u32)pc.enc.level_budget) throw std::runtime_error("bootstrap: not enough levels");
  if (level_0 <= bts_depth) throw std::runtime_error("bootstrap: need set a larger multiply depth");
  const u32 level_enc = level_0 - pc.enc.level_budget, level_dec = level_0 - bts_depth;

  // Coeffs2slots_precomp :778-865
  auto collapse = [&](bool encoding) {
    if (!sparse) return coeff_collapse(ksi, rot_group, pc.enc.level_budget, false, encoding);
    vvvcd a = coeff_collapse(ksi, rot_group, pc.enc.level_budget, false, encoding);
    vvvcd b = coeff_collapse(ksi, rot_group, pc.enc.level_budget, true, encoding);
    for (size_t i = 0; i < a.size(); i++)
      for (size_t j = 0; j < a[i].size(); j++)
        a[i][j].insert(a[i][j].end(), b[i][j].begin(), b[i][j].end());
    return a;
  };
  {
    vvvcd coeffs = collapse(true);
    SinPoly sp = sin_poly(c->params.hamming_weight);
    double factor = 1.0 / N;
    factor /= sp.k;
    const double sf = (double)((u64)1 << c->params.scaling_mod_size);
    double ratio = round(log2(dbl_q0 / sf));
    factor /= pow(2, ratio);
    factor = pow(factor, 1. / pc.enc.level_budget);
    for (auto& lvl : coeffs)
      for (auto& row : lvl)
        for (cd& x : row) x = cscale(x, factor);
    rotate_precomp(pc, coeffs, sparse ? 2 * slots : slots, scale_enc, level_enc, true);
  }
  {
    vvvcd coeffs = collapse(false);
    rotate_precomp(pc, coeffs, sparse ? 2 * slots : slots, scale_dec, level_dec, false);
  }
  c->sync();
  precom()[slots] = std::move(pcp);
}

// Find_coeffslots_rot_index :214-278
void Evaluator::find_coeffslots_rot_index(std::vector<int32_t>& out, const BtsFftParams& p,
                                          u32 slots, u32 m, bool encoding) {
  const int32_t flag_rem = p.layers_rem != 0 ? 1 : 0, stop = flag_rem ? 0 : -1;
  const u32 mdiv4 = m / 4;
  const int32_t start = encoding ? stop + 1 : 0, end = p.level_budget;
  const int32_t slots_value = encoding ? (int32_t)slots : (int32_t)mdiv4;
  for (int32_t s = start; s < end; s++) {
    int32_t shift = encoding ? 1 << ((s - flag_rem) * p.layers_coll + p.layers_rem)
                             : 1 << (s * p.layers_coll);
    for (int32_t j = 0; j < p.g; j++)
      out.push_back((int32_t)reduce_rotation((j - ((p.num_rot + 1) / 2) + 1) * shift, slots_value));
    for (int32_t i = 0; i < p.b; i++)
      out.push_back((int32_t)reduce_rotation((p.g * i) * shift, mdiv4));
  }
  if (flag_rem) {
    int32_t s = p.level_budget - flag_rem;
    int32_t shift = encoding ? 1 : 1 << (s * p.layers_coll);
    for (int32_t j = 0; j < p.g_rem; j++)
      out.push_back((int32_t)reduce_rotation((j - ((p.num_rot_rem + 1) / 2) + 1) * shift, slots_value));
    for (int32_t i = 0; i < p.b_rem; i++)
      out.push_back((int32_t)reduce_rotation(p.g_rem * i * shift, mdiv4));
  }
  const u32 slots4 = slots * 4;
  if (slots4 != m)
    for (u32 j = 1; j < m / slots4; j <<= 1) out.push_back((int32_t)(j * slots));
}

std::vector<int32_t> Evaluator::bootstrap_rot_indices(u32 num_slots) {
  const u32 m = 2 * c->N, slots = num_slots == 0 ? m / 4 : num_slots;
  u32 budget[2] = {3, 3};
  double log_slots = log2(slots);
  for (u32& v : budget) {
    if (v > log_slots) v = (u32)log_slots;
    if (v < 1) v = 1;
  }
  BtsFftParams enc = fft_params(slots, budget[0], 0), dec = fft_params(slots, budget[1], 0);
  std::vector<int32_t> v;
  find_coeffslots_rot_index(v, enc, slots, m, true);
  find_coeffslots_rot_index(v, dec, slots, m, false);
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  v.erase(std::remove_if(v.begin(), v.end(), [&](int32_t x) { return x == 0 || x == (int32_t)(m / 4); }),
          v.end());
  return v;
}

// ------------------------------------------------------------------------------ linear transforms
// One baby-step/giant-step level of CoeffsToSlots / SlotsToCoeffs with hoisted rotations.
// Reference dataflow (Rotate_iteration): ModUp(c1) once; g giant... "fast" rotations in the
// extended basis (key inner product + P*c0, automorphism, no ModDown); b inner sums against the
// diagonal plaintexts; every inner sum but the first is key-switched by its baby-step rotation;
// two ModDowns at the very end.
void Evaluator::rotate_iteration(Ct& result, BtsPrecom& pc,
                                 const std::vector<std::vector<int32_t>>& rin,
                                 const std::vector<std::vector<int32_t>>& rout, int32_t step,
                                 bool encoding, bool is_rem) {
  const BtsFftParams& P = encoding ? pc.enc : pc.dec;
  const int32_t giant = is_rem ? P.g_rem : P.g, baby = is_rem ? P.b_rem : P.b;
  const int32_t num_rot = is_rem ? P.num_rot_rem : P.num_rot;
  const int32_t level_idx = encoding ? P.level_budget - 1 : 0;
  if (is_rem || step != level_idx) rescale(result, result);

  const u32 N = c->N, K = (u32)c->K, nq = result.nq, W = nq + K;
  const size_t WN = (size_t)W * N;
  const Basis ext_b{nq, K, (u32)c->L}, q_b{nq, 0, (u32)c->L};
  const auto& tab = encoding ? pc.c2s : pc.s2c;
  const u32 pt_level = (encoding ? pc.c2s_level : pc.s2c_level)[step];
  if (nq > pt_level) throw std::runtime_error("bootstrap: ciphertext level exceeds the plaintext table");
  const u32 beta = (u32)c->num_decomp(nq);

  u64* ext = c->alloc_limbs((size_t)beta * W, false);
  c->modup_all(ext, result.c1, nq);

  // giant-step rotations in the extended basis: rot[j] = (r0, r1), W limbs each
  u64* rot = c->alloc_limbs((size_t)giant * 2 * W, false);
  u64* acc = c->alloc_limbs(2 * (size_t)W, false);
  for (int32_t j = 0; j < giant; j++) {
    u64 *r0 = rot + (size_t)j * 2 * WN, *r1 = r0 + WN;
    int32_t val = rin[step][j];
    if (val != 0) {
      // Fast_rotate_ext in one pass: key inner product, + P*c0, automorphism (written through
      // the inverse table)
      const SwitchKey& key = rot_key(val);
      if (!key.k0 || !key.k1) throw std::runtime_error("switch key not loaded");
      const int64_t* inv = c->auto_order_inv(c->auto_index(val));
      launch_ksw_inner_rot(c->T, r0, r1, ext, result.c1, (u32)c->part_size, key.k0, key.k1, beta, nq,
                           (u32)c->L, K, result.c0, c->pmodq_, c->pmodq_sh_, inv, false, false,
                           c->stream);
      {
        const uint64_t n = 2ull * beta * W;  // what ksw_acc counts
        c->tr(Context::TR_LIMB_MUL, 0, n);
        c->tr(Context::TR_LIMB_ADD, 0, n);
      }
      c->tr(Context::TR_LIMB_MUL, 0, nq);
      c->tr(Context::TR_LIMB_ADD, 0, nq);
      c->tr(Context::TR_LIMB_ROT, 0, 2 * W);
      c->launches += 1;
    } else {  // Switch_key_ext: lift (c0, c1) by P, P limbs are zero
      launch_mul_scalar(c->T, r0, result.c0, c->pmodq_, c->pmodq_sh_, 0, nq, c->stream);
      launch_mul_scalar(c->T, r1, result.c1, c->pmodq_, c->pmodq_sh_, 0, nq, c->stream);
      ACE_CUDA(cudaMemsetAsync(r0 + (size_t)nq * N, 0, (size_t)K * N * sizeof(u64), c->stream));
      ACE_CUDA(cudaMemsetAsync(r1 + (size_t)nq * N, 0, (size_t)K * N * sizeof(u64), c->stream));
      c->tr(Context::TR_LIMB_MUL, 0, 2 * nq);
      c->launches += 2;
    }
  }

  u64* first = c->alloc_limbs(W, false);
  u64* outer = c->alloc_limbs(2 * (size_t)W, false);   // (outer0, outer1)
  u64* tmp   = c->alloc_limbs(2 * (size_t)W, false);
  bool outer0_zero = true;
  if (baby <= kMaxDotBaby && giant <= kMaxDotGiant) {
    // ---- batched form.  The baby steps only meet in the final accumulations, so: all inner sums
    // in one pass over the giant-step rotations, ONE ModDown over every inner.c1, ONE ModUp over
    // all their digits, then per step the key inner product (its own key) fused with the
    // automorphism and the accumulation.  Exact ring additions commute, the approximate base
    // conversions see the same inputs as in the step-by-step form: same limbs.
    u64* inner_all = c->alloc_limbs((size_t)(baby > 1 ? baby - 1 : 1) * 2 * W, false);
    u64* zero_pt   = c->alloc_limbs((size_t)pt_level + K, true);  // stands in for the absent diagonal
    DotAllArgs da;
    da.rot = rot; da.rot_stride = 2 * WN; da.c1_offset = WN;
    da.b = (u32)baby; da.g = (u32)giant; da.pt_pstart = pt_level;
    for (int32_t i = 0; i < baby; i++) {
      u32 used = 0;
      for (int32_t j = 0; j < giant; j++) {
        const u64* pt = (giant * i + j == num_rot) ? nullptr : tab[step][giant * i + j];
        if (giant * i + j != num_rot && !pt) throw std::runtime_error("bootstrap: missing diagonal plaintext");
        da.pt[i * giant + j] = pt ? pt : zero_pt;  // absent term: multiply by zero
        used += pt ? 1 : 0;
      }
      // first = inner_0.c0, outer = (0, inner_0.c1)
      da.out0[i] = i == 0 ? first : inner_all + (size_t)(i - 1) * 2 * WN;
      da.out1[i] = i == 0 ? outer + WN : inner_all + (size_t)(i - 1) * 2 * WN + WN;
      c->tr(Context::TR_LIMB_MUL, 0, 2ull * used * W);
      c->tr(Context::TR_LIMB_ADD, 0, 2ull * (used - 1) * W);
    }
    launch_pt_dot_all(c->T, da, ext_b, c->stream);
    c->launches++;
    std::vector<int32_t> sw;  // baby steps that need a key switch
    for (int32_t i = 1; i < baby; i++) {
      u64* inner = inner_all + (size_t)(i - 1) * 2 * WN;
      if (rout[step][i] != 0) {
        sw.push_back(i);
      } else {
        launch_ew_basis(c->T, EW_ADD, first, first, inner, ext_b, c->stream);
        launch_ew_basis(c->T, EW_ADD, outer + WN, outer + WN, inner + WN, ext_b, c->stream);
        c->tr(Context::TR_LIMB_ADD, 0, 2 * W);
        c->launches += 2;
      }
    }
    if (!sw.empty()) {
      const size_t ns = sw.size();
      u64* red_all = c->alloc_limbs(ns * nq, false);
      u64* ext_all = c->alloc_limbs(ns * (size_t)beta * W, false);
      std::vector<ModdownJob> md(ns);
      std::vector<ModupJob>   mu;
      for (size_t k = 0; k < ns; k++) {
        u64* inner = inner_all + (size_t)(sw[k] - 1) * 2 * WN;
        u64* red   = red_all + k * (size_t)nq * N;
        md[k] = ModdownJob{red, inner + WN, nq};
        for (u32 j = 0; j < beta; j++)
          mu.push_back(ModupJob{ext_all + (k * beta + j) * WN, red + (size_t)c->digit_start(j) * N, nq, j,
                                false});
      }
      c->moddown_batch(md.data(), ns);
      c->modup_batch(mu.data(), mu.size());
      for (size_t k = 0; k < ns; k++) {
        const int32_t val = rout[step][sw[k]];
        u64* inner = inner_all + (size_t)(sw[k] - 1) * 2 * WN;
        const int64_t* order = c->auto_order(c->auto_index(val));
        const int64_t* inv   = c->auto_order_inv(c->auto_index(val));
        launch_gather_add_basis(c->T, first, inner, order, ext_b, c->stream);  // first += rot(inner.c0)
        const SwitchKey& key = rot_key(val);
        if (!key.k0 || !key.k1) throw std::runtime_error("switch key not loaded");
        launch_ksw_inner_rot(c->T, outer, outer + WN, ext_all + k * beta * WN, red_all + k * (size_t)nq * N,
                             (u32)c->part_size, key.k0, key.k1, beta, nq, (u32)c->L, K, nullptr, nullptr,
                             nullptr, inv, !outer0_zero, true, c->stream);
        outer0_zero = false;
        const uint64_t n = 2ull * beta * W;  // what ksw_acc counts
        c->tr(Context::TR_LIMB_MUL, 0, n);
        c->tr(Context::TR_LIMB_ADD, 0, n);
        c->tr(Context::TR_LIMB_ROT, 0, 3 * W);
        c->tr(Context::TR_LIMB_ADD, 0, 3 * W);
        c->launches += 2;
      }
      c->free_limbs(red_all);
      c->free_limbs(ext_all);
    }
    c->free_limbs(inner_all);
    c->free_limbs(zero_pt);
  } else {
  // ---- step-by-step form (more baby or giant steps than the batched kernel holds)
  u64* inner = c->alloc_limbs(2 * (size_t)W, false);
  u64* red   = c->alloc_limbs(nq, false);
  for (int32_t i = 0; i < baby; i++) {
    const int32_t gbase = giant * i;
    DotArgs da;
    da.n = 0;
    da.pt_pstart = pt_level;
    for (int32_t j = 0; j < giant; j++) {
      if (gbase + j == num_rot) continue;
      const u64* pt = tab[step][gbase + j];
      if (!pt) throw std::runtime_error("bootstrap: missing diagonal plaintext");
      da.a0[da.n] = rot + (size_t)j * 2 * WN;
      da.a1[da.n] = rot + (size_t)j * 2 * WN + WN;
      da.pt[da.n] = pt;
      da.n++;
    }
    u64 *in0 = i == 0 ? first : inner, *in1 = i == 0 ? outer + WN : inner + WN;
    launch_pt_dot(c->T, in0, in1, da, ext_b, c->stream);
    c->tr(Context::TR_LIMB_MUL, 0, 2ull * da.n * W);   // Mul_plaintext per rotation ...
    c->tr(Context::TR_LIMB_ADD, 0, 2ull * (da.n - 1) * W);  // ... and Add_ciphertext
    c->launches++;
    if (i == 0) continue;  // first = inner.c0, outer = (0, inner.c1)
    int32_t val = rout[step][i];
    if (val != 0) {
      const int64_t* order = c->auto_order(c->auto_index(val));
      const int64_t* inv   = c->auto_order_inv(c->auto_index(val));
      // first += rot(inner.c0)
      launch_gather_add_basis(c->T, first, inner, order, ext_b, c->stream);
      // inner.c1 -> Q basis -> digits -> key switch in the extended basis -> rotate -> outer:
      // the inner product writes through the inverse automorphism table and accumulates
      c->mod_down_pair(red, nullptr, inner + WN, nullptr, nq, nullptr);
      c->modup_all(ext, red, nq);
      const SwitchKey& key = rot_key(val);
      if (!key.k0 || !key.k1) throw std::runtime_error("switch key not loaded");
      launch_ksw_inner_rot(c->T, outer, outer + WN, ext, red, (u32)c->part_size, key.k0, key.k1, beta,
                           nq, (u32)c->L, K, nullptr, nullptr, nullptr, inv, !outer0_zero, true,
                           c->stream);
      outer0_zero = false;
      {
        const uint64_t n = 2ull * beta * W;  // what ksw_acc counts
        c->tr(Context::TR_LIMB_MUL, 0, n);
        c->tr(Context::TR_LIMB_ADD, 0, n);
      }
      c->tr(Context::TR_LIMB_ROT, 0, 3 * W);
      c->tr(Context::TR_LIMB_ADD, 0, 3 * W);
      c->launches += 2;
    } else {
      launch_ew_basis(c->T, EW_ADD, first, first, inner, ext_b, c->stream);
      launch_ew_basis(c->T, EW_ADD, outer + WN, outer + WN, inner + WN, ext_b, c->stream);
      c->tr(Context::TR_LIMB_ADD, 0, 2 * W);
      c->launches += 2;
    }
  }
  c->free_limbs(inner);
  c->free_limbs(red);
  }
  // outer.c0 += first; ModDown both
  if (outer0_zero) {
    ACE_CUDA(cudaMemcpyAsync(outer, first, WN * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
  } else {
    launch_ew_basis(c->T, EW_ADD, outer, outer, first, ext_b, c->stream);
    c->launches++;
  }
  c->tr(Context::TR_LIMB_ADD, 0, W);
  (void)q_b;
  c->mod_down_pair(result.c0, result.c1, outer, outer + WN, nq, nullptr);
  const double delta = (double)((u64)1 << c->params.scaling_mod_size);
  result.sf  = result.sf * pow(delta, 1);  // Mul_plaintext: sf * plain sf, degree + 1
  result.sfd = result.sfd + 1;
  for (u64* p : {ext, rot, acc, first, outer, tmp}) c->free_limbs(p);
}

void Evaluator::coeff_slots_transform(Ct& result, Ct& in, BtsPrecom& pc, bool encoding) {
  const BtsFftParams& P = encoding ? pc.enc : pc.dec;
  const u32 order = 2 * c->N, slots = in.slots;
  const int32_t budget = P.level_budget;
  const int32_t flag_rem = P.layers_rem != 0 ? 1 : 0, stop = flag_rem ? 0 : -1;
  const int32_t start = encoding ? stop + 1 : 0;
  const int32_t end   = encoding ? budget : budget - flag_rem;
  const int32_t slots_value = encoding ? (int32_t)slots : (int32_t)(order / 4);
  const u32 rem_index = encoding ? 0 : budget - 1;
  std::vector<std::vector<int32_t>> rin(budget), rout(budget);
  for (int32_t i = 0; i < budget; i++) {
    rin[i].assign(((flag_rem && (u32)i == rem_index) ? P.num_rot_rem : P.num_rot) + 1, 0);
    rout[i].assign(P.b + P.b_rem, 0);
  }
  for (int32_t s = start; s < end; s++) {
    int32_t shift = encoding ? ((s - flag_rem) * P.layers_coll + P.layers_rem) : (s * P.layers_coll);
    for (int32_t j = 0; j < P.g; j++)
      rin[s][j] = (int32_t)reduce_rotation((j - ((P.num_rot + 1) / 2) + 1) * (1 << shift), slots_value);
    for (int32_t i = 0; i < P.b; i++)
      rout[s][i] = (int32_t)reduce_rotation((P.g * i) * (1 << shift), order / 4);
  }
  if (flag_rem) {
    int32_t s = encoding ? stop : budget - flag_rem;
    int32_t shift = encoding ? 1 : (1 << (s * P.layers_coll));
    for (int32_t j = 0; j < P.g_rem; j++)
      rin[s][j] = (int32_t)reduce_rotation((j - ((P.num_rot_rem + 1) / 2) + 1) * shift, slots_value);
    for (int32_t i = 0; i < P.b_rem; i++)
      rout[s][i] = (int32_t)reduce_rotation((P.g_rem * i) * shift, order / 4);
  }
  copy(result, in);
  if (encoding) {
    for (int32_t s = end - 1; s > start - 1; s--) rotate_iteration(result, pc, rin, rout, s, true, false);
  } else {
    for (int32_t s = start; s < end; s++) rotate_iteration(result, pc, rin, rout, s, false, false);
  }
  if (flag_rem) rotate_iteration(result, pc, rin, rout, encoding ? stop : budget - flag_rem, encoding, true);
}

// ------------------------------------------------------------------------------ EvalMod
void Evaluator::apply_double_angle(Ct& x, u32 num_iter) {
  const int32_t r = (int32_t)num_iter;
  for (int32_t j = 1; j < r + 1; j++) {
    mul(x, x, x);
    add(x, x, x);
    double scalar = -1.0 / pow((2.0 * M_PI), pow(2.0, j - r));
    add_const(x, x, scalar);
    rescale(x, x);
  }
}

void Evaluator::eval_approx_mod(Ct& out, Ct& in, const std::vector<double>& coeffs) {
  SinPoly sp = sin_poly(c->params.hamming_weight);
  if (sp.even_kind) add_const(in, in, -1. / (4. * sp.k));  // y = x - 1/(4K)
  eval_chebyshev(out, in, coeffs, -1, 1);
  apply_double_angle(out, sp.r);
}

// ------------------------------------------------------------------------------ Eval_bootstrap
void Evaluator::eval_bootstrap(Ct& res, Ct& in, u32 raise_level, BtsPrecom& pc) {
  const u32 N = c->N, m = 2 * N, slots = in.slots, q_cnt = (u32)c->L;
  if (!c->has_rot_key(m - 1)) throw std::runtime_error("cannot find conj key");
  const double sf = (double)((u64)1 << c->params.scaling_mod_size);
  const int32_t deg = (int32_t)round(log2((double)(int64_t)c->mod[0] / sf));
  const u32 init_q = in.nq;

  Ct raised;
  copy(raised, in);
  while (raised.sfd > 1) rescale(raised, raised);
  if (!raise_level) raise_level = q_cnt;
  if (raise_level > q_cnt) throw std::runtime_error("the raise level must be <= q_cnt");

  // ModRaise: limb 0 in coefficient form, centred lift to every q_i, back to NTT form
  Ct nw;
  reserve(nw, raise_level, 0);
  nw.sf = raised.sf; nw.sfd = raised.sfd; nw.slots = slots;
  u64* coef = c->alloc_limbs(2, false);
  c->intt_from(coef, raised.c0, 0, 1);
  c->intt_from(coef + N, raised.c1, 0, 1);
  launch_mod_raise(c->T, nw.c0, coef, raise_level, c->stream);
  launch_mod_raise(c->T, nw.c1, coef + N, raise_level, c->stream);
  c->tr(Context::TR_LIMB_NTT, 0, 2);         // the two INTTs above (intt_from does not count)
  c->tr(Context::TR_LIMB_ADD, 0, 2 * raise_level);  // Switch_modulus pass ~ one limb add each
  c->launches += 2;
  c->ntt(nw.c0, 0, raise_level);
  c->ntt(nw.c1, 0, raise_level);
  c->free_limbs(coef);
  release(raised);

  SinPoly sp = sin_poly(c->params.hamming_weight);
  std::vector<double> coeffs(sp.coeff, sp.coeff + sp.n);

  if (slots == m / 4) {  // fully packed
    Ct enc, conj, encs;
    coeff_slots_transform(enc, nw, pc, true);
    conjugate(conj, enc);
    sub(encs, enc, conj);
    add(enc, enc, conj);
    mul_monomial(encs, encs, 3 * m / 4);
    while (enc.sfd > 1) {
      rescale(enc, enc);
      rescale(encs, encs);
    }
    eval_approx_mod(enc, enc, coeffs);
    eval_approx_mod(encs, encs, coeffs);
    mul_monomial(encs, encs, m / 4);
    add(enc, enc, encs);
    coeff_slots_transform(res, enc, pc, false);
    release(enc); release(conj); release(encs);
  } else {  // sparsely packed
    Ct temp;
    for (u32 j = 1; j < N / (2 * slots); j <<= 1) {
      rotate(temp, nw, (int32_t)(j * slots));
      add(nw, nw, temp);
    }
    release(temp);
    Ct enc, conj;
    coeff_slots_transform(enc, nw, pc, true);
    conjugate(conj, enc);
    add(enc, enc, conj);
    release(conj);
    while (enc.sfd > 1) rescale(enc, enc);
    eval_approx_mod(enc, enc, coeffs);
    coeff_slots_transform(res, enc, pc, false);
    Ct r;
    rotate(r, res, (int32_t)slots);
    add(res, res, r);
    release(enc); release(r);
  }

  if (env_flag("RT_BTS_CLEAR_IMAG") && deg >= 1) {
    Ct conj;
    conjugate(conj, res);
    add(res, res, conj);
    release(conj);
    u64 ratio = (u64)pow(2., deg - 1);
    if (ratio > 1) mul_integer(res, res, (u32)ratio);
  } else {
    u64 ratio = (u64)pow(2., deg);
    mul_integer(res, res, (u32)ratio);
  }
  release(nw);
  while (res.sfd > 1) rescale(res, res);
  if (res.nq <= init_q) copy(res, in);  // bootstrapping earned nothing: return the input
}

void Evaluator::linear_transform(Ct& res, Ct& in, bool encoding) {
  if (!precom().count(in.slots)) bootstrap_setup(in.slots);
  coeff_slots_transform(res, in, *precom()[in.slots], encoding);
}

const u64* Evaluator::diagonal_plain(u32 slots, bool encoding, u32 step, u32 idx, u32* level) {
  if (slots == 0) slots = c->N / 2;
  if (!precom().count(slots)) bootstrap_setup(slots);
  BtsPrecom& pc = *precom()[slots];
  auto& tab = encoding ? pc.c2s : pc.s2c;
  if (step >= tab.size() || idx >= tab[step].size()) return nullptr;
  *level = (encoding ? pc.c2s_level : pc.s2c_level)[step];
  return tab[step][idx];
}

// host-only view of the collapsed FFT diagonals (no device work): [level][row][slots] (re, im)
size_t Evaluator::fft_diagonals(u32 slots, u32 budget, bool flag, bool encoding, double* out) {
  const u32 slots4 = 4 * slots;
  std::vector<u32> rot_group(slots);
  u32 five = 1;
  for (u32 i = 0; i < slots; i++) {
    rot_group[i] = five;
    five *= 5;
    five %= slots4;
  }
  vcd ksi(slots4 + 1);
  for (size_t i = 0; i < slots4; i++) {
    double angle = 2.0 * M_PI * i / slots4;
    ksi[i] = cd(cos(angle), sin(angle));
  }
  ksi[slots4] = ksi[0];
  vvvcd co = coeff_collapse(ksi, rot_group, budget, flag, encoding);
  size_t n = 0;
  for (auto& lvl : co)
    for (auto& row : lvl)
      for (cd& v : row) {
        out[2 * n] = v.real();
        out[2 * n + 1] = v.imag();
        n++;
      }
  return n;
}

// Bootstrap (src/ckks/cipher_eval.c:366-404)
void Evaluator::bootstrap(Ct& res, Ct& in, u32 level_after_bts) {
  const u32 slots = in.slots, q_cnt = (u32)c->L;
  if (!precom().count(slots)) bootstrap_setup(slots);
  BtsPrecom& pc = *precom()[slots];
  const u32 bts_depth = bootstrap_depth(c->params.hamming_weight);
  if (in.sfd == 1 && in.nq >= level_after_bts) {
    copy(res, in);
    return;
  }
  if (level_after_bts && level_after_bts > q_cnt - bts_depth)
    throw std::runtime_error("The level set after bootstrapping is excessively high");
  u32 raise_level = level_after_bts ? level_after_bts + bts_depth : q_cnt;
  if (&res == &in) {
    Ct out;
    eval_bootstrap(out, in, raise_level, pc);
    move(res, out);
  } else {
    eval_bootstrap(res, in, raise_level, pc);
  }
}

}  // namespace ace
