// Exact inference on a pedigree blanket + the reference's children heuristic, host and device.
//
// Replaces the pgmpy VariableElimination.query call sites of correct_genotypes3.py:119-121,242-245,
// 388-390 on the network bayesnet/model.py:173-230 builds (founders of the sub-pedigree get the
// allele prior [1/4,1/2,1/4], every other node the 3x3x3 Mendel table of model.py:28-30), and
// bayesnet/model.py:65-89 generate_probs_kids.
//
// Method (not pgmpy's): "frontier elimination".  Blanket nodes are stored parents-before-children
// (DFS post-order from the focal animals).  Walking that order we keep ONE joint table over the
// currently open unobserved variables; a node multiplies its Mendel factor in (adding a base-3 digit
// if it is unobserved) and a variable is summed out right after its last child.  Hard evidence never
// adds a digit, so with mostly-genotyped ancestors the table has 3..27 entries.  Factors that are
// not connected to the query through unobserved variables are skipped, which is exactly pgmpy's
// behaviour of dropping factors whose scope became empty (contradictory evidence that is d-separated
// from the query cannot produce NaN; contradictory evidence inside the query's component gives 0/0).
// Every unnormalised quantity is a dyadic rational that binary64 holds exactly, so the result of the
// single final division does not depend on the elimination order.
#pragma once
#include <stdint.h>
#include <math.h>
#include <cuda_fp16.h>

#define GFX_WMAX 6            // open unobserved variables the per-thread table holds (fast path)
#define GFX_TAB 729           // 3^GFX_WMAX
#define GFX_WMAX_BIG 9        // rare wide blankets rerun on a pooled global-memory table
#define GFX_TAB_BIG 19683     // 3^GFX_WMAX_BIG
#define GFX_WMAX_HUGE 13      // last resort: a few 12.75 MB tables (suspect filters can open most of a pair blanket)
#define GFX_TAB_HUGE 1594323  // 3^GFX_WMAX_HUGE
#define GFX_MAXN 64           // max nodes in one blanket (bit masks are uint64)
#define GFX_HD __host__ __device__ __forceinline__

// P(kid = k | sire = s, dam = d): bayesnet/model.py:28-30, generated from allele transmission
// probabilities s/2 and d/2 (all values are exact dyadics, identical to the table).
GFX_HD double gfx_mendel(int k, int s, int d) {
  const double ps = 0.5 * (double)s, pd = 0.5 * (double)d;
  if (k == 0) return (1.0 - ps) * (1.0 - pd);
  if (k == 2) return ps * pd;
  return ps * (1.0 - pd) + (1.0 - ps) * pd;
}
GFX_HD double gfx_prior(int k) { return k == 1 ? 0.5 : 0.25; }  // model.py:24

GFX_HD int gfx_pow3(int k) {
  int r = 1;
  for (int i = 0; i < k; ++i) r *= 3;
  return r;
}

struct GfxBlanket {
  int n;              // nodes, topological (parents first)
  const int8_t* p1;   // local index of sire or -1 (both parents or none are inside the blanket)
  const int8_t* p2;   // local index of dam  or -1
  const int8_t* last; // local index of the node's last child inside the blanket (own index if none)
};

GFX_HD int gfx_lowbit(uint64_t m) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)m) - 1;
#else
  return __builtin_ctzll(m);
#endif
}
GFX_HD int gfx_highbit(uint64_t m) {
#ifdef __CUDA_ARCH__
  return 63 - __clzll((long long)m);
#else
  return 63 - __builtin_clzll(m);
#endif
}

// Core of the inference: sum-product over the component `comp` of the query q (bit mask of the
// unobserved variables connected to q), visiting only the candidate factors in `cand` (any superset
// of comp and the children of comp; nodes outside are never referenced, so code[] / obs only need
// to be valid for comp, its children and their parents).
//
// Three reductions keep the table small without changing the (exact, dyadic) unnormalised values:
//   * only factors connected to q through unobserved variables are used (pgmpy drops the others);
//   * barren nodes -- unobserved, not q, no evidence below them -- are skipped: their factor sums to 1;
//   * an unobserved founder with a single remaining child is folded into that child's factor
//     (sum_p prior(p) T(x | p, .)) instead of opening a digit for it.
// Returns 0, or 1 if more than `wmax` variables would be open at once (out = NaN).
// tab[3] receives the unnormalised total P(evidence in the component).
// TS = distance between consecutive table entries in doubles: 1 for a private table, the number of threads that
// interleave their tables in one shared-memory array otherwise (entry i of a thread at tab[i * TS]: bank = thread).
template <int TS = 1>
__host__ __device__ inline int gfx_eliminate(const GfxBlanket& b, int q, const uint8_t* code, uint64_t obs,
                                             uint64_t comp, uint64_t cand, double* tab, double out[3], int wmax) {
  // needed factors: connected to q and not barren (reverse topological sweep)
  uint64_t need = 1ull << q, inc = 0;
  int8_t nkids[GFX_MAXN], dlast[GFX_MAXN];
  for (uint64_t m = cand; m;) {
    const int v = gfx_highbit(m);
    m &= ~(1ull << v);
    nkids[v] = 0;
    dlast[v] = (int8_t)v;
    const int a1 = b.p1[v];
    if (a1 >= 0) { nkids[a1] = 0; nkids[b.p2[v]] = 0; dlast[a1] = (int8_t)a1; dlast[b.p2[v]] = (int8_t)b.p2[v]; }
  }
  for (uint64_t m = cand; m;) {
    const int v = gfx_highbit(m);
    m &= ~(1ull << v);
    uint64_t sc = 1ull << v;
    const int a1 = b.p1[v], a2 = b.p2[v];
    if (a1 >= 0) sc |= (1ull << a1) | (1ull << a2);
    if (!((sc & ~obs) & comp)) continue;
    inc |= 1ull << v;
    if ((obs >> v) & 1) need |= 1ull << v;  // evidence that touches the component
    if (!((need >> v) & 1)) continue;
    if (a1 >= 0) {
      need |= (1ull << a1) | (1ull << a2);
      ++nkids[a1]; ++nkids[a2];
      if (dlast[a1] < v) dlast[a1] = (int8_t)v;
      if (dlast[a2] < v) dlast[a2] = (int8_t)v;
    }
  }
  need &= inc;  // observed parents were marked too; their own factors count only if connected
  int8_t open[GFX_WMAX_HUGE];
  int w = 0, size = 1;
  uint64_t folded = 0;
  double fv[GFX_MAXN * 3];  // belief vectors of folded nodes (only touched entries are used)
  tab[0] = 1.0;
  for (uint64_t m = need; m;) {
    const int v = gfx_lowbit(m);
    m &= m - 1;
    const int a1 = b.p1[v], a2 = b.p2[v];
    const bool v_obs = (obs >> v) & 1;
    if (a1 < 0 && v_obs) continue;               // prior of an observed founder: scalar, dropped
    // where do the parents' states come from: evidence, a folded belief vector, or a table digit
    int s_pos = -1, d_pos = -1, s_val = 0, d_val = 0, s_str = 1, d_str = 1;
    bool s_fold = false, d_fold = false;
    if (a1 >= 0) {
      if ((obs >> a1) & 1) s_val = code[a1];
      else if ((folded >> a1) & 1) s_fold = true;
      else { for (int k = 0; k < w; ++k) if (open[k] == a1) s_pos = k; s_str = gfx_pow3(s_pos); }
      if ((obs >> a2) & 1) d_val = code[a2];
      else if ((folded >> a2) & 1) d_fold = true;
      else { for (int k = 0; k < w; ++k) if (open[k] == a2) d_pos = k; d_str = gfx_pow3(d_pos); }
    }
    // T(x | s, d) factorises over the transmitted alleles: with al(.) = P(parent passes allele 0) and
    // be(.) = P(passes allele 1), T(0|.) = al_s al_d, T(2|.) = be_s be_d, T(1|.) = al_s be_d + be_s al_d.
    // Per parent these masses come from the evidence, from a folded belief vector, or from a table digit.
    double sa[3], sb[3], da[3], db[3];
    const int na = s_pos >= 0 ? 3 : 1, nb = d_pos >= 0 ? 3 : 1;
    if (a1 >= 0) {
      if (s_pos >= 0) { sa[0] = 1.0; sa[1] = 0.5; sa[2] = 0.0; sb[0] = 0.0; sb[1] = 0.5; sb[2] = 1.0; }
      else if (s_fold) { sa[0] = fv[a1 * 3] + 0.5 * fv[a1 * 3 + 1]; sb[0] = 0.5 * fv[a1 * 3 + 1] + fv[a1 * 3 + 2]; }
      else { sa[0] = 1.0 - 0.5 * s_val; sb[0] = 0.5 * s_val; }
      if (d_pos >= 0) { da[0] = 1.0; da[1] = 0.5; da[2] = 0.0; db[0] = 0.0; db[1] = 0.5; db[2] = 1.0; }
      else if (d_fold) { da[0] = fv[a2 * 3] + 0.5 * fv[a2 * 3 + 1]; db[0] = 0.5 * fv[a2 * 3 + 1] + fv[a2 * 3 + 2]; }
      else { da[0] = 1.0 - 0.5 * d_val; db[0] = 0.5 * d_val; }
    }
    const int x_lo = v_obs ? code[v] : 0, x_hi = v_obs ? code[v] : 2;
    // an unobserved node with a single remaining child whose own factor involves no table digit is
    // carried as a belief vector (plain peeling) instead of opening a digit
    if (!v_obs && v != q && nkids[v] == 1 && s_pos < 0 && d_pos < 0) {
      if (a1 < 0) { fv[v * 3] = 0.25; fv[v * 3 + 1] = 0.5; fv[v * 3 + 2] = 0.25; }
      else {
        fv[v * 3] = sa[0] * da[0];
        fv[v * 3 + 1] = sa[0] * db[0] + sb[0] * da[0];
        fv[v * 3 + 2] = sb[0] * db[0];
      }
      folded |= 1ull << v;
      continue;
    }
    if (!v_obs && w == wmax) { out[0] = out[1] = out[2] = nan(""); return 1; }
    // The factor value depends on the digits of the two parents only.  The table is walked three entries at a time --
    // the three values of digit 0 --: three independent loads, products and stores per step, and the parents' digits
    // advance once per step.  [The version before walked runs of the smaller parent stride with the factor hoisted: a
    // parent at digit 0 meant runs of ONE entry with the digit bookkeeping on every entry; these loops were 46 % of the
    // instructions and 50 % of the stall samples of k_correct_jobs_general (profiles/r2_stalls_gen4_general_before.txt),
    // and the correction stage of a 10 000 x 10 000 chunk took 10.45 ms against 9.41 ms now.]  Every entry still
    // receives the single product tab[i] * f of the per-entry formulation: results are unchanged.
    const bool s_d0 = na == 3 && s_str == 1, d_d0 = nb == 3 && d_str == 1;   // parent's digit is digit 0
    const bool s_hi = na == 3 && !s_d0, d_hi = nb == 3 && !d_d0;             // parent's digit is a higher one
    // All states x of an unobserved v are written in ONE pass over the table (one load and three stores per entry; three
    // passes with a load each before), and the next three entries are requested before the current ones are stored: the
    // compiler cannot move a load above the stores itself (dst may alias tab; it never does for the entries ahead).
    double f27[27];
    for (int x = x_lo; x <= x_hi; ++x) {
      double* f9 = f27 + x * 9;
      if (a1 < 0) f9[0] = gfx_prior(x);
      else
        for (int a = 0; a < na; ++a)
          for (int c = 0; c < nb; ++c)
            f9[a * 3 + c] = x == 0 ? sa[a] * da[c] : (x == 2 ? sb[a] * db[c] : sa[a] * db[c] + sb[a] * da[c]);
    }
    const size_t xstr = v_obs ? 0 : (size_t)size * TS;   // table of state x starts at tab + x * xstr
    if (size == 1) {
      const double t = tab[0];
      for (int x = x_hi; x >= x_lo; --x) tab[x * xstr] = t * f27[x * 9];
    } else {
      int ia = 0, ca = 0, ib = 0, cb = 0;
      const int fstep = (s_d0 ? 3 : 0) + (d_d0 ? 1 : 0);
      double t0 = tab[0], t1 = tab[TS], t2 = tab[2 * TS];
      for (int i = 0; i < size; i += 3) {
        const int fb = (s_hi ? ia : 0) * 3 + (d_hi ? ib : 0);
        const double c0 = t0, c1 = t1, c2 = t2;
        if (i + 3 < size) { t0 = tab[(i + 3) * TS]; t1 = tab[(i + 4) * TS]; t2 = tab[(i + 5) * TS]; }
        for (int x = x_hi; x >= x_lo; --x) {
          double* dst = tab + x * xstr;
          const double* f = f27 + x * 9 + fb;
          dst[i * TS] = c0 * f[0];
          dst[(i + 1) * TS] = c1 * f[fstep];
          dst[(i + 2) * TS] = c2 * f[2 * fstep];
        }
        if (s_hi) { ca += 3; if (ca == s_str) { ca = 0; ia = ia == 2 ? 0 : ia + 1; } }
        if (d_hi) { cb += 3; if (cb == d_str) { cb = 0; ib = ib == 2 ? 0 : ib + 1; } }
      }
    }
    if (!v_obs) { open[w++] = (int8_t)v; size *= 3; }
    // sum out every open variable whose last needed child is this node
    for (int k = w - 1; k >= 0; --k) {
      const int u = open[k];
      if (u == q || dlast[u] > v) continue;
      const int str = gfx_pow3(k);
      const int nsize = size / 3;
      if (str == 1) {
        int i = 0;
        for (; i + 3 <= nsize; i += 3) {
          double t[9];
          for (int e = 0; e < 9; ++e) t[e] = tab[(3 * i + e) * TS];
          tab[i * TS] = (t[0] + t[1]) + t[2];
          tab[(i + 1) * TS] = (t[3] + t[4]) + t[5];
          tab[(i + 2) * TS] = (t[6] + t[7]) + t[8];
        }
        for (; i < nsize; ++i) tab[i * TS] = (tab[3 * i * TS] + tab[(3 * i + 1) * TS]) + tab[(3 * i + 2) * TS];
      } else {
        for (int i = 0, src = 0; i < nsize; src += 2 * str)
          for (int lo = 0; lo < str; lo += 3, i += 3, src += 3) {
            double t[9];
            for (int e = 0; e < 3; ++e) {
              t[e] = tab[(src + e) * TS];
              t[3 + e] = tab[(src + str + e) * TS];
              t[6 + e] = tab[(src + 2 * str + e) * TS];
            }
            for (int e = 0; e < 3; ++e) tab[(i + e) * TS] = (t[e] + t[3 + e]) + t[6 + e];
          }
      }
      for (int mm = k; mm + 1 < w; ++mm) open[mm] = open[mm + 1];
      --w;
      size = nsize;
    }
  }
  const double s = (tab[0] + tab[1 * TS]) + tab[2 * TS];
  out[0] = tab[0] / s;
  out[1] = tab[1 * TS] / s;
  out[2] = tab[2 * TS] / s;
  tab[3 * TS] = s;
  return 0;
}

// Marginal of node q given hard evidence on the nodes whose bit is set in `obs` (their states in
// code[]), every node's status known up front.  tab: scratch of 3^wmax doubles (>= 4).
__host__ __device__ inline int gfx_infer(const GfxBlanket& b, int q, const uint8_t* code, uint64_t obs,
                                         double* tab, double out[3], int wmax = GFX_WMAX) {
  const int n = b.n;
  obs &= ~(1ull << q);
  // component of q in the interaction graph of unobserved variables
  uint64_t comp = 1ull << q;
  for (bool grew = true; grew;) {
    grew = false;
    for (int v = 0; v < n; ++v) {
      uint64_t sc = 1ull << v;
      if (b.p1[v] >= 0) sc |= (1ull << b.p1[v]) | (1ull << b.p2[v]);
      sc &= ~obs;
      if ((sc & comp) && (sc & ~comp)) { comp |= sc; grew = true; }
    }
    for (int v = n - 1; v >= 0 && grew; --v) {
      uint64_t sc = 1ull << v;
      if (b.p1[v] >= 0) sc |= (1ull << b.p1[v]) | (1ull << b.p2[v]);
      sc &= ~obs;
      if ((sc & comp) && (sc & ~comp)) comp |= sc;
    }
  }
  const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1);
  return gfx_eliminate(b, q, code, obs, comp, all, tab, out, wmax);
}

// correctPair queries [sire, dam] jointly with joint=False (correct_genotypes3.py:242-245): pgmpy
// multiplies every remaining factor, marginalises to each variable and normalises.  The marginals
// equal the single-variable queries unless the OTHER animal's component has P(evidence) = 0, in
// which case the joint is all zero and both marginals are 0/0 = NaN.
__host__ __device__ inline int gfx_infer_pair(const GfxBlanket& b, int q0, int q1, const uint8_t* code, uint64_t obs,
                                              double* tab, double out0[3], double out1[3], int wmax = GFX_WMAX) {
  obs &= ~((1ull << q0) | (1ull << q1));
  int bad = gfx_infer(b, q0, code, obs, tab, out0, wmax);
  const double s0 = tab[3];
  bad |= gfx_infer(b, q1, code, obs, tab, out1, wmax);
  const double s1 = tab[3];
  if (!bad && (s0 == 0.0 || s1 == 0.0)) {
    out0[0] = out0[1] = out0[2] = nan("");
    out1[0] = out1[1] = out1[2] = nan("");
  }
  return bad;
}

// ---- float16 score pieces of correct_genotypes3.py:121,126-127,149-155 -------------------------
// numpy half arithmetic = compute in float32, round once to half.
GFX_HD float gfx_h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
GFX_HD uint16_t gfx_f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// f16 bits of 2 * (nanmax(anc16) - anc16[obs]) where anc16 = float16(anc)
GFX_HD uint16_t gfx_anc_score(const double anc[3], int obs) {
  float f[3];
  for (int i = 0; i < 3; ++i) f[i] = __half2float(__double2half(anc[i]));
  float mx = nanf("");
  for (int i = 0; i < 3; ++i)
    if (f[i] == f[i] && !(mx >= f[i])) mx = f[i];
  const uint16_t diff = gfx_f2h(mx - f[obs]);
  return gfx_f2h(2.0f * gfx_h2f(diff));
}
// f16 bits of the summed children differences: m units of f16(1/3) = 1365/4096, summed in float32
GFX_HD uint16_t gfx_child_score(uint32_t m) { return gfx_f2h((float)(m * 1365u) * (1.0f / 4096.0f)); }
GFX_HD uint16_t gfx_half_add(uint16_t a, uint16_t b) { return gfx_f2h(gfx_h2f(a) + gfx_h2f(b)); }

// ---- children heuristic, bayesnet/model.py:65-89 ------------------------------------------------
// kv[((k*4 + p)*3 + a)*3 + x]: Mendel row of a child in state k, masked by the other parent's state
// p (3 = unknown) on the sire axis a, divided by the row sum (model.py:78-84); x is the dam axis.
__host__ __device__ inline void gfx_fill_kv(double* kv) {
  for (int k = 0; k < 3; ++k)
    for (int p = 0; p < 4; ++p) {
      double row[9];
      for (int a = 0; a < 3; ++a)
        for (int x = 0; x < 3; ++x) row[3 * a + x] = (p < 3 && a != p) ? 0.0 : gfx_mendel(k, a, x);
      // np.sum over 9 contiguous doubles: 8 unrolled accumulators, then the tail
      double s = ((row[0] + row[1]) + (row[2] + row[3])) + ((row[4] + row[5]) + (row[6] + row[7]));
      s += row[8];
      for (int i = 0; i < 9; ++i) kv[(k * 4 + p) * 9 + i] = s > 0 ? row[i] / s : row[i];
    }
}

// numpy's pairwise summation (loops_utils.h pairwise_sum) streamed over `n` elements that each carry
// three values (the three dam-axis columns are summed in lock step).  next(v) yields the next element.
template <class Next>
__host__ __device__ inline void gfx_pw_leaf(Next& next, int n, double res[3]) {
  double v[3];
  if (n < 8) {
    res[0] = res[1] = res[2] = 0.0;
    for (int i = 0; i < n; ++i) { next(v); res[0] += v[0]; res[1] += v[1]; res[2] += v[2]; }
    return;
  }
  double r[8][3];
  for (int j = 0; j < 8; ++j) { next(v); r[j][0] = v[0]; r[j][1] = v[1]; r[j][2] = v[2]; }
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) { next(v); r[j][0] += v[0]; r[j][1] += v[1]; r[j][2] += v[2]; }
  for (int c = 0; c < 3; ++c)
    res[c] = ((r[0][c] + r[1][c]) + (r[2][c] + r[3][c])) + ((r[4][c] + r[5][c]) + (r[6][c] + r[7][c]));
  for (; i < n; ++i) { next(v); res[0] += v[0]; res[1] += v[1]; res[2] += v[2]; }
}

template <class Next>
__host__ __device__ inline void gfx_pairwise3(Next& next, int n, double res[3]) {
  if (n <= 128) { gfx_pw_leaf(next, n, res); return; }
  // explicit post-order walk of numpy's recursive halving (n2 = n/2 rounded down to a multiple of 8)
  int fn[24], fstage[24];
  double fleft[24][3];
  int sp = 0;
  fn[0] = n; fstage[0] = 0;
  double val[3] = {0, 0, 0};
  while (sp >= 0) {
    const int m = fn[sp];
    if (m <= 128) {
      gfx_pw_leaf(next, m, val);
      --sp;
    } else if (fstage[sp] == 0) {
      int n2 = m / 2; n2 -= n2 % 8;
      fstage[sp] = 1;
      fn[sp + 1] = n2; fstage[sp + 1] = 0;
      ++sp;
      continue;
    } else if (fstage[sp] == 1) {
      int n2 = m / 2; n2 -= n2 % 8;
      fleft[sp][0] = val[0]; fleft[sp][1] = val[1]; fleft[sp][2] = val[2];
      fstage[sp] = 2;
      fn[sp + 1] = m - n2; fstage[sp + 1] = 0;
      ++sp;
      continue;
    } else {
      val[0] = fleft[sp][0] + val[0]; val[1] = fleft[sp][1] + val[1]; val[2] = fleft[sp][2] + val[2];
      --sp;
    }
  }
  res[0] = val[0]; res[1] = val[1]; res[2] = val[2];
}

// kid_code(i) -> state of the i-th genotyped child (0..2, anything else = not usable);
// par_code(i) -> state of that child's other parent (0..2, anything else = unknown).
// np.nanmean(table[:, dam_axis == x]): the fancy-indexed block is column-major, so memory order is
// sire-axis a = 0,1,2 outermost and the usable children innermost.
template <class KidCode, class ParCode>
struct GfxKidStream {
  int a, cur, nkids;
  KidCode& kid_code;
  ParCode& par_code;
  const double* kv;
  __host__ __device__ void operator()(double v[3]) {
    for (;;) {
      if (cur == nkids) { cur = 0; ++a; }
      const int k = kid_code(cur);
      if (k <= 2) {
        int p = par_code(cur);
        if (p > 2) p = 3;
        const double* e = kv + (k * 4 + p) * 9 + 3 * a;
        v[0] = e[0]; v[1] = e[1]; v[2] = e[2];
        ++cur;
        return;
      }
      ++cur;
    }
  }
};

// mean over the 3*ninc entries of each dam-axis column, then divide by the Python sum (model.py:85-88)
GFX_HD void gfx_kids_finish(const double tot[3], int ninc, double out[3]) {
  const double cnt = (double)(3 * ninc);
  double r0 = tot[0] / cnt, r1 = tot[1] / cnt, r2 = tot[2] / cnt;
  const double s = (r0 + r1) + r2;
  if (s > 0) { r0 /= s; r1 /= s; r2 /= s; }
  out[0] = r0; out[1] = r1; out[2] = r2;
}

// Returns false when there is no genotyped child at all (the reference's `probs = None`).
template <class KidCode, class ParCode>
__host__ __device__ inline bool gfx_kids_term(int nkids, KidCode kid_code, ParCode par_code, const double* kv,
                                              double out[3]) {
  if (nkids <= 0) return false;
  int ninc = 0;
  for (int i = 0; i < nkids; ++i) ninc += kid_code(i) <= 2;
  if (ninc == 0) { out[0] = out[1] = out[2] = 1.0 / 3.0; return true; }
  GfxKidStream<KidCode, ParCode> next{0, 0, nkids, kid_code, par_code, kv};
  double tot[3];
  gfx_pairwise3(next, 3 * ninc, tot);
  gfx_kids_finish(tot, ninc, out);
  return true;
}

// Same sum over a compacted list of usable children: cls[i] = 4 * state + other-parent class.
struct GfxClsStream {
  int a, cur, ninc;
  const uint8_t* cls;
  const double* kv;
  __host__ __device__ void operator()(double v[3]) {
    if (cur == ninc) { cur = 0; ++a; }
    const double* e = kv + (int)cls[cur] * 9 + 3 * a;
    v[0] = e[0]; v[1] = e[1]; v[2] = e[2];
    ++cur;
  }
};
__host__ __device__ inline void gfx_kids_term_cls(int ninc, const uint8_t* cls, const double* kv, double out[3]) {
  if (ninc == 0) { out[0] = out[1] = out[2] = 1.0 / 3.0; return; }
  GfxClsStream next{0, 0, ninc, cls, kv};
  double tot[3];
  gfx_pairwise3(next, 3 * ninc, tot);
  gfx_kids_finish(tot, ninc, out);
}

// correct_genotypes3.py:280-304 / :410-422 : product (or whichever exists), divided by its sum if > 0
GFX_HD void gfx_combine(bool has_anc, const double anc[3], bool has_kids, const double kids[3], double p[3]) {
  for (int i = 0; i < 3; ++i) p[i] = (has_anc && has_kids) ? anc[i] * kids[i] : (has_anc ? anc[i] : kids[i]);
  const double s = (p[0] + p[1]) + p[2];
  if (s > 0) { p[0] /= s; p[1] /= s; p[2] /= s; }
}

// Decision inputs of correct_genotypes3.py:741-744,809 (pairs) / :932 (singles), computed from the
// generation-start snapshot: bits 0-2 tie set, bit 3 (max - P[obs]) >= threshold, bit 7 valid.
GFX_HD uint8_t gfx_decision_bits(const double p[3], int obs_snapshot, double tie, double threshold) {
  double mx = nan("");
  for (int i = 0; i < 3; ++i)
    if (p[i] == p[i] && !(mx >= p[i])) mx = p[i];
  uint8_t r = 0x80;
  const double cut = mx - tie;
  for (int i = 0; i < 3; ++i)
    if (p[i] >= cut) r |= (uint8_t)(1u << i);
  const double op = obs_snapshot <= 2 ? p[obs_snapshot] : nan("");
  if ((mx - op) >= threshold) r |= 8u;
  return r;
}
