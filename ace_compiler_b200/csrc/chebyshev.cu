// chebyshev.cu -- homomorphic evaluation of a Chebyshev series with the Paterson-Stockmeyer
// recursion, as the reference does in fhe-cmplr/rtlib/ant/src/util/ckks_chebyshev.c.
// The double-precision coefficient manipulation (long division in the Chebyshev basis) decides
// which plaintext constants get encoded, so it follows the reference step by step:
//   Get_degree_from_coeffs :42-53, Is_even_poly :56-64, Compute_degree_ps :96-134,
//   Long_div_chebyshev :155-270, Eval_linear_wsum(_mutable) :282-323, Eval_quot_or_rem :325-389,
//   Inner_eval_chebyshev_ps :393-501, Eval_chebyshev_ps :503-672.
#include <cmath>
#include <cstdlib>

#include "evaluator.h"
#include "host_math.h"

namespace ace {

typedef std::vector<double> vd;

// index of the last non-zero coefficient (0 for an all-zero list)
static u32 series_degree(const vd& v) {
  if (v.empty()) return 0;
  u32 drop = 1;
  for (int64_t i = (int64_t)v.size() - 1; i > 0; i--) {
    if (v[i] == 0) drop++;
    else break;
  }
  return (u32)v.size() - drop;
}

static bool series_is_even(const vd& v) {
  u32 d = series_degree(v);
  for (u32 i = 1; i <= d; i += 2)
    if (v[i] != 0.) return false;
  return true;
}

// Paterson-Stockmeyer split (k, m) for degree n; table for n <= 2204, heuristic above
static void ps_split(u32 n, u32& k, u32& m) {
  static const u32 upto[16] = {2, 11, 13, 17, 55, 59, 76, 239, 247, 284, 991, 1007, 1083, 2015, 2031, 2204};
  static const u32 mval[16] = {1, 2, 3, 2, 3, 4, 3, 4, 5, 4, 5, 6, 5, 6, 7, 6};
  if (n <= 2204) {
    u32 idx = n - 1, t = 0;
    while (idx >= upto[t]) t++;
    m = mval[t];
    k = (u32)floor(n / ((1 << m) - 1)) + 1;
    return;
  }
  double sqn2 = floor(log2(sqrt(n / 2)));
  u32 best = 0xffffffffu;
  k = m = 0;
  for (u32 kk = 1; kk <= n; kk++) {
    for (u32 mm = 1; mm <= ceil(log2(n / kk) + 1) + 1; mm++) {
      if (((int32_t)n - (int32_t)kk * ((1 << mm) - 1)) < 0 && abs(floor(log2(kk)) - sqn2) <= 1) {
        u32 mul = kk + 2 * mm + (1 << (mm - 1)) - 4;
        if (best > mul) { best = mul; k = kk; m = mm; }
      }
    }
  }
}

static const double kPrec = 9.5367431640625e-07;  // 2^-20
static bool differs_from_one(double v) { return (1 - kPrec >= v) || (1 + kPrec <= v); }

// f / g in the Chebyshev basis: quotient q and remainder r (c0 convention, not c0/2)
static void cheb_long_div(vd& q, vd& r, const vd& f, const vd& g) {
  u32 n = series_degree(f), k = series_degree(g);
  if (n != f.size() - 1 || k != g.size() - 1)
    throw std::runtime_error("chebyshev division: dominant coefficient is zero");
  r = f;
  if (n < k) { q.assign(1, 0.0); return; }
  q.assign(n - k + 1, 0.0);
  const double g_back = g.back();
  while (n > k) {
    q[n - k] = 2 * r.back();
    if (differs_from_one(g[k])) q[n - k] = 2 * r.back() / g_back;
    vd d(n + 1, 0.0);
    if (k == n - k) {
      d[0] = 2 * g[n - k];
      for (u32 i = 1; i < 2 * k + 1; i++) d[i] = g[abs((int32_t)(n - k - i))];
    } else if ((int32_t)k > (int32_t)(n - k)) {
      d[0] = 2 * g[n - k];
      for (u32 i = 1; i < k - (n - k) + 1; i++)
        d[i] = g[abs((int32_t)(n - k - i))] + g[(int32_t)(n - k + i)];
      for (u32 i = k - (n - k) + 1; i < n + 1; i++) d[i] = g[abs((int32_t)(i - n + k))];
    } else {
      d[n - k] = g[0];
      for (u32 i = n - 2 * k; i < n + 1; i++) d[i] = g[abs((int32_t)(i - n + k))];
    }
    const double r_back = r.back();
    if (differs_from_one(r_back))
      for (double& x : d) x = x * r_back;
    if (differs_from_one(g_back))
      for (double& x : d) x = x / g_back;
    for (size_t i = 0; i < r.size(); i++) r[i] = r[i] - d[i];
    if (r.size() > 1) {
      n = series_degree(r);
      r.resize(n + 1, 0.0);
    }
  }
  if (n == k) {
    const double r_back = r.back();
    q[0] = r_back;
    if (differs_from_one(g_back)) q[0] = r_back / g_back;
    vd d(g);
    if (differs_from_one(r_back))
      for (double& x : d) x = x * r_back;
    if (differs_from_one(g_back))
      for (double& x : d) x = x / g_back;
    for (size_t i = 0; i < r.size(); i++) r[i] = r[i] - d[i];
    if (r.size() > 1) {
      n = series_degree(r);
      r.resize(n + 1, 0.0);
    }
  }
  q[0] = q[0] * 2;  // c0 convention
}

// out = rescale(sum_i weights[i] * list[i]) over the first `size` ciphertexts
void Evaluator::eval_linear_wsum(Ct& out, std::vector<Ct>& list, size_t size,
                                 const double* weights) {
  // One kernel per 12 terms instead of Mul_const + Add_ciph (+ a copy) per term: the same canonical
  // residues (modular sums do not depend on their order), a fifth of the traffic, and a bootstrap
  // loses ~200 of its ~960 launches.  ACE_B200_NO_WSUM=1: term by term, as the reference.
  static const bool fused = getenv("ACE_B200_NO_WSUM") == nullptr;
  bool usable = fused && c->L + 0 <= (size_t)kMaxWsumLimbs;
  for (size_t i = 0; i < size && usable; i++)
    if (weights[i] != 0. && (list[i].np || &list[i] == &out)) usable = false;
  if (usable) {
    const double delta = (double)((u64)1 << c->params.scaling_mod_size);
    u32  level = 0;
    bool any = false;
    for (size_t i = 0; i < size; i++)
      if (weights[i] != 0.) { level = any ? (list[i].nq < level ? list[i].nq : level) : list[i].nq; any = true; }
    if (!any) throw std::runtime_error("polynomial has no non-zero coefficient");
    static thread_local WsumArgs args;
    args.n = 0; args.acc = 0;
    bool first_term = true;
    u32  run_level = 0;
    for (size_t i = 0; i < size; i++) {
      if (weights[i] == 0.) continue;
      Ct& a = list[i];
      if (first_term) {  // what copy(out, Mul_const(list[i])) leaves in out's header
        reserve(out, a.nq, 0);
        out.sf = a.sf * pow(delta, 1); out.sfd = a.sfd + 1; out.slots = a.slots;
        run_level = a.nq;
        first_term = false;
      } else {
        run_level = a.nq < run_level ? a.nq : run_level;
        c->tr(Context::TR_LIMB_ADD, 0, 2 * run_level);
      }
      c->tr(Context::TR_LIMB_MUL, 0, 2 * a.nq);
      const std::vector<u64> res = c->value_residues(weights[i], a.nq, 1);
      const u32 t = args.n++;
      args.c0[t] = a.c0; args.c1[t] = a.c1;
      for (u32 y = 0; y < level; y++) { args.w[t][y] = res[y]; args.wsh[t][y] = hm::shoup(res[y], c->mod[y]); }
      if (args.n == (u32)kMaxWsum) {
        launch_ct_wsum(c->T, out.c0, out.c1, args, level, c->stream);
        c->launches++;
        args.n = 0; args.acc = 1;
      }
    }
    if (args.n) { launch_ct_wsum(c->T, out.c0, out.c1, args, level, c->stream); c->launches++; }
    out.nq = level;
    rescale(out, out);
    return;
  }
  bool first = true;
  Ct tmp;
  for (size_t i = 0; i < size; i++) {
    if (weights[i] == 0.) continue;
    mul_const(tmp, list[i], weights[i]);
    if (first) { copy(out, tmp); first = false; }
    else add(out, out, tmp);
  }
  if (first) throw std::runtime_error("polynomial has no non-zero coefficient");
  rescale(out, out);
  release(tmp);
}

void Evaluator::eval_quot_or_rem(Ct& out, std::vector<Ct>& t, const vd& qr, u32 k, bool is_quot,
                                 bool in_rec) {
  vd cp(qr);
  cp.resize(k, 0.0);
  Ct& tk1 = t[k - 1];
  size_t dg = series_degree(cp);
  if (dg > 0) {
    eval_linear_wsum(out, t, dg, cp.data() + 1);
    if (is_quot) {
      if (in_rec) {
        // leading coefficient is a power of two: add T_k that many times by doubling
        double last = qr.back();
        Ct sum;
        copy(sum, tk1);
        for (u32 i = 0; i < log2(last); i++) add(sum, sum, sum);
        add(out, out, sum);
        release(sum);
      } else {
        add(out, out, tk1);
        add(out, out, tk1);
      }
    } else {
      add(out, out, tk1);
    }
  } else {
    copy(out, tk1);
    if (is_quot) {
      double last = qr.back();
      u32 end = in_rec ? (u32)log2(last) : (u32)last;
      for (u32 i = 0; i < end; i++) add(out, out, tk1);
    }
  }
  add_const(out, out, qr[0] / 2);
}

void Evaluator::inner_eval_ps(Ct& out, const vd& coeffs, u32 k, u32 m, std::vector<Ct>& t,
                              std::vector<Ct>& t2, bool in_rec) {
  const u32 k2m2k = k * (1 << (m - 1)) - k;
  vd tkm(k2m2k + k + 1, 0.0);
  tkm.back() = 1;
  vd div_q, div_r;
  cheb_long_div(div_q, div_r, coeffs, tkm);

  vd r2(div_r);
  if (k2m2k <= series_degree(div_r)) {
    r2[k2m2k] = r2[k2m2k] - 1;
    r2.resize(series_degree(r2) + 1, 0.0);
  } else {
    r2.resize(k2m2k + 1, 0.0);
    r2.back() = -1;
  }
  vd cq, cr;
  cheb_long_div(cq, cr, r2, div_q);

  size_t s2_len = cr.size() > (size_t)(k2m2k + 1) ? cr.size() : (size_t)(k2m2k + 1);
  vd s2(cr);
  s2.resize(s2_len, 0.0);
  s2[s2_len - 1] = 1;

  Ct cu;
  u32 dc = series_degree(cq);
  bool have_c = false;
  if (dc >= 1) {
    if (dc == 1) {
      if (cq[1] != 1) {
        mul_const(cu, t[0], cq[1]);
        rescale(cu, cu);
      } else {
        copy(cu, t[0]);
      }
    } else {
      eval_linear_wsum(cu, t, dc, cq.data() + 1);
    }
    add_const(cu, cu, cq[0] / 2);
    have_c = true;
  }

  Ct qu, su;
  if (series_degree(div_q) > k) inner_eval_ps(qu, div_q, k, m - 1, t, t2, true);
  else eval_quot_or_rem(qu, t, div_q, k, true, in_rec);
  if (series_degree(s2) > k) inner_eval_ps(su, s2, k, m - 1, t, t2, true);
  else eval_quot_or_rem(su, t, s2, k, false, in_rec);

  Ct& t2m = t2[m - 1];
  if (have_c) {
    cu.nq = t2m.nq;  // Set_ciph_level(cu, level of T2[m-1])
    add(out, t2m, cu);
  } else {
    add_const(out, t2m, cq[0] / 2);
  }
  mul(out, out, qu);
  rescale(out, out);
  add(out, out, su);
  release(qu); release(su); release(cu);
}

void Evaluator::eval_chebyshev(Ct& out, Ct& in, const vd& coeffs, double a, double b) {
  const u32 n = series_degree(coeffs);
  if (n < 5) throw std::runtime_error("Eval_chebyshev_linear: not implemented (as in the reference)");
  const bool even = series_is_even(coeffs);
  vd f2;
  if (coeffs.back() == 0 && !std::signbit(coeffs.back())) f2.assign(coeffs.begin(), coeffs.begin() + n + 1);
  else f2 = coeffs;

  u32 k, m;
  ps_split(n, k, m);
  if (even && (k % 2 == 1)) k += 1;

  std::vector<Ct> t(k);
  double rnd_a = round(a), rnd_b = round(b);
  if ((rnd_a == -1) && (rnd_b == 1) && (a - rnd_a < 1E-10) && (b - rnd_b < 1E-10)) {
    copy(t[0], in);
  } else {
    double alpha = 2 / (b - a), beta = alpha * a;
    mul_const(t[0], in, alpha);
    rescale(t[0], t[0]);
    add_const(t[0], t[0], -1.0 - beta);
  }
  Ct y, prod;
  copy(y, t[0]);
  const u32 neg1_sfd = t[0].sfd;  // the -1 plaintext is encoded once with T_1's sf degree

  for (u32 i = 2; i <= k; i++) {
    Ct& tj = t[i - 1];
    if (!(i & (i - 1))) {  // power of two: T_i = 2 T_{i/2}^2 - 1
      Ct& h = t[i / 2 - 1];
      mul(prod, h, h);
      add(tj, prod, prod);
      rescale(tj, tj);
      add_const_sfd(tj, tj, -1.0, neg1_sfd);
    } else if (i % 2 == 1) {  // odd: T_i = 2 T_{(i-1)/2} T_{(i+1)/2} - T_1
      if (even) continue;
      mul(prod, t[i / 2 - 1], t[i / 2]);
      add(tj, prod, prod);
      rescale(tj, tj);
      sub(tj, tj, y);
    } else {  // even, not a power of two
      u32 h1 = i / 2;
      if (even && (h1 % 2 == 1)) h1 += 1;
      u32 h2 = i - h1;
      mul(prod, t[h1 - 1], t[h2 - 1]);
      add(tj, prod, prod);
      rescale(tj, tj);
      if (h1 == h2) add_const_sfd(tj, tj, -1.0, neg1_sfd);
      else sub(tj, tj, t[1]);
    }
  }
  // bring every T_i to the level of T_k (FIXED_MANUAL branch)
  for (size_t i = 1; i < k; i++) {
    if (even && i % 2 == 1) continue;
    if (t[i - 1].nq > t[k - 1].nq) t[i - 1].nq = t[k - 1].nq;
  }

  std::vector<Ct> t2(m);
  copy(t2[0], t[k - 1]);
  for (u32 i = 1; i < m; i++) {
    mul(prod, t2[i - 1], t2[i - 1]);
    add(t2[i], prod, prod);
    rescale(t2[i], t2[i]);
    add_const_sfd(t2[i], t2[i], -1.0, neg1_sfd);
  }
  Ct t2km1;
  copy(t2km1, t2[0]);
  for (u32 i = 1; i < m; i++) {
    mul(prod, t2km1, t2[i]);
    add(t2km1, prod, prod);
    rescale(t2km1, t2km1);
    sub(t2km1, t2km1, t2[0]);
  }

  const u32 k2m2k = k * (1 << (m - 1)) - k;
  f2.resize(2 * k2m2k + k + 1, 0.0);
  f2.back() = 1;
  inner_eval_ps(out, f2, k, m, t, t2, false);
  sub(out, out, t2km1);

  release(y); release(prod); release(t2km1);
  for (Ct& x : t) release(x);
  for (Ct& x : t2) release(x);
}

}  // namespace ace
