// See gfx_schedule.h.  Plain C++ (no CUDA): one pass over the pedigree with stamp arrays, so building
// the schedule is O(animals * blanket size) instead of the reference's O(animals^2) get_subset
// (pedigree/pedigree_dag.py:149-152 scans every kid2sire/kid2dam entry per call).
#include "gfx_schedule.h"

#include <algorithm>
#include <unordered_set>

namespace {

struct Builder {
  const gfx_pedigree_desc* ped;
  std::vector<int32_t> stamp;   // stamp[a] == cur  <=> a is in the current node set
  std::vector<int32_t> local;   // local index of a in the current blanket
  int32_t cur = 0;
  std::string err;
  int err_code = GFX_OK;

  explicit Builder(const gfx_pedigree_desc* p) : ped(p), stamp(p->n_ped, 0), local(p->n_ped, -1) {}

  // Appends a blanket for `focal` (1 or 2 animals) at `depth` to `bs`.  `plain` (optional) is set
  // when the blanket is exactly the duplicate-free ancestor tree (no closure additions, no extra
  // edges), i.e. when per-branch recursion equals exact inference.  Returns blanket id or <0.
  int add_blanket(const int32_t* focal, int nf, int depth, GfxBlanketSet& bs, bool* plain) {
    ++cur;
    std::vector<int32_t> nodes;
    bool dup = false;
    for (int i = 0; i < nf; ++i) {
      if (stamp[focal[i]] != cur) { stamp[focal[i]] = cur; nodes.push_back(focal[i]); }
    }
    // level-by-level ancestors (pedigree_dag.py:221-231); duplicates flag loops
    std::vector<int32_t> level(focal, focal + nf), next;
    std::vector<int32_t> last_level;
    for (int d = 0; d < depth && !level.empty(); ++d) {
      next.clear();
      for (int32_t k : level) {
        const int32_t par[2] = {ped->sire[k], ped->dam[k]};
        for (int32_t p : par) {
          if (p < 0) continue;
          next.push_back(p);
          if (stamp[p] == cur) dup = true;
          else { stamp[p] = cur; nodes.push_back(p); }
        }
      }
      level.swap(next);
      if (d == depth - 1) last_level = level;
    }
    // closure under "partner of an included parent" (pedigree_dag.py:124-139)
    bool added = false;
    for (bool changed = true; changed;) {
      changed = false;
      const size_t n0 = nodes.size();
      for (size_t i = 0; i < n0; ++i) {
        const int32_t k = nodes[i];
        const int32_t s = ped->sire[k], d = ped->dam[k];
        if (s < 0 && d < 0) continue;
        if (s < 0 || d < 0) {
          err = "animal " + std::to_string((long long)ped->ped_id[k]) +
                " has exactly one known parent (outside the reference's domain: KeyError in "
                "pedigree_dag.py:126-128)";
          err_code = GFX_E_PEDIGREE;
          return -1;
        }
        const bool si = stamp[s] == cur, di = stamp[d] == cur;
        if (si && !di) { stamp[d] = cur; nodes.push_back(d); changed = added = true; }
        if (di && !si) { stamp[s] = cur; nodes.push_back(s); changed = added = true; }
      }
    }
    if ((int)nodes.size() > GFX_MAXN_HOST) {
      err = "blanket with " + std::to_string(nodes.size()) + " nodes exceeds the supported maximum";
      err_code = GFX_E_TOO_WIDE;
      return -1;
    }
    // DFS post-order from the focal animals over in-blanket parents -> topological order
    std::vector<int32_t> order;
    order.reserve(nodes.size());
    for (int32_t k : nodes) local[k] = -1;
    std::vector<std::pair<int32_t, int>> st;
    auto has_par = [&](int32_t k) {
      const int32_t s = ped->sire[k], d = ped->dam[k];
      return s >= 0 && d >= 0 && stamp[s] == cur && stamp[d] == cur;
    };
    auto dfs = [&](int32_t root) {
      if (local[root] != -1) return;
      st.push_back({root, 0});
      local[root] = -2;  // on stack
      while (!st.empty()) {
        auto& top = st.back();
        const int32_t k = top.first;
        if (top.second < 2 && has_par(k)) {
          const int32_t p = top.second == 0 ? ped->sire[k] : ped->dam[k];
          ++top.second;
          if (local[p] == -1) { local[p] = -2; st.push_back({p, 0}); }
        } else {
          local[k] = (int32_t)order.size();
          order.push_back(k);
          st.pop_back();
        }
      }
    };
    for (int i = 0; i < nf; ++i) dfs(focal[i]);
    for (int32_t k : nodes) dfs(k);  // closure additions are always reachable, but stay safe
    const int n = (int)order.size();
    const size_t base = bs.row.size();
    bs.row.resize(base + n);
    bs.p1.resize(base + n);
    bs.p2.resize(base + n);
    bs.last.resize(base + n);
    bs.kidmask.resize(base + n, 0);
    for (int i = 0; i < n; ++i) {
      const int32_t k = order[i];
      bs.row[base + i] = ped->row[k];
      bs.last[base + i] = (int8_t)i;
      if (has_par(k)) {
        bs.p1[base + i] = (int8_t)local[ped->sire[k]];
        bs.p2[base + i] = (int8_t)local[ped->dam[k]];
      } else {
        bs.p1[base + i] = bs.p2[base + i] = -1;
      }
    }
    for (int i = 0; i < n; ++i) {
      if (bs.p1[base + i] >= 0) {
        bs.last[base + bs.p1[base + i]] = std::max<int8_t>(bs.last[base + bs.p1[base + i]], (int8_t)i);
        bs.last[base + bs.p2[base + i]] = std::max<int8_t>(bs.last[base + bs.p2[base + i]], (int8_t)i);
        bs.kidmask[base + bs.p1[base + i]] |= 1ull << i;
        bs.kidmask[base + bs.p2[base + i]] |= 1ull << i;
      }
    }
    // worst-case frontier width (nothing observed)
    {
      int w = 0, mw = 0;
      std::vector<int> open;
      for (int i = 0; i < n; ++i) {
        open.push_back(i);
        mw = std::max(mw, (int)open.size());
        for (int k = (int)open.size() - 1; k >= 0; --k) {
          const int u = open[k];
          const bool is_focal = (u == local[focal[0]]) || (nf > 1 && u == local[focal[1]]);
          if (!is_focal && bs.last[base + u] <= i) open.erase(open.begin() + k);
        }
      }
      (void)w;
      bs.max_width = std::max(bs.max_width, mw);
    }
    bs.max_nodes = std::max(bs.max_nodes, n);
    bs.q0.push_back((int8_t)local[focal[0]]);
    bs.q1.push_back(nf > 1 ? (int8_t)local[focal[1]] : (int8_t)-1);
    bs.off.push_back((int32_t)(base + n));
    {
      uint8_t tr = 0;
      for (int fi = 0; fi < nf; ++fi) {
        const int qf = local[focal[fi]];
        // the OTHER focal animal is never evidence (correct_genotypes3.py:230-238 drops both) and has no child in the
        // blanket: as somebody's child it is barren (its factor sums to 1) and does not couple anything
        const int qo_t = nf > 1 ? local[focal[1 - fi]] : -1;
        const uint64_t drop_t = (qo_t >= 0 && bs.kidmask[base + qo_t] == 0) ? (1ull << qo_t) : 0ull;
        bool ok = bs.kidmask[base + qf] == 0 && bs.p1[base + qf] >= 0;
        std::vector<int> stk;
        if (ok) { stk.push_back(bs.p1[base + qf]); stk.push_back(bs.p2[base + qf]); }
        while (ok && !stk.empty()) {
          const int u = stk.back();
          stk.pop_back();
          if (__builtin_popcountll(bs.kidmask[base + u] & ~drop_t) != 1) { ok = false; break; }
          if (bs.p1[base + u] >= 0) { stk.push_back(bs.p1[base + u]); stk.push_back(bs.p2[base + u]); }
        }
        if (ok) tr |= (uint8_t)(1u << fi);
      }
      bs.tree.push_back(tr);
    }
    for (int fi = 0; fi < 2; ++fi) {
      int32_t srow[14];
      int snode[14];
      for (int i = 0; i < 14; ++i) { srow[i] = -3; snode[i] = -1; }
      uint16_t bad = 0;
      if (fi >= nf) bad = 0x8000;
      else {
        const int qf = local[focal[fi]];
        const int qo = nf > 1 ? local[focal[1 - fi]] : -1;
        if (bs.kidmask[base + qf] != 0 || bs.p1[base + qf] < 0) bad = 0x8000;
        else {
          snode[0] = bs.p1[base + qf];
          snode[1] = bs.p2[base + qf];
          for (int sl = 0; sl < 14; ++sl) {
            const int u = snode[sl];
            if (u < 0) continue;
            srow[sl] = u == qo ? -2 : (bs.row[base + u] >= 0 ? bs.row[base + u] : -1);
            // (the other focal animal as a child is barren: see the tree flags above)
            const uint64_t drop = (qo >= 0 && bs.kidmask[base + qo] == 0) ? (1ull << qo) : 0ull;
            if (__builtin_popcountll(bs.kidmask[base + u] & ~drop) != 1) bad |= (uint16_t)(1u << sl);
            if (bs.p1[base + u] >= 0) {
              if (sl < 6) { snode[2 * sl + 2] = bs.p1[base + u]; snode[2 * sl + 3] = bs.p2[base + u]; }
              else bad |= (uint16_t)(1u << sl);   // parents beyond the slot tree
            }
          }
        }
      }
      for (int i = 0; i < 14; ++i) bs.slotrow.push_back(srow[i]);
      bs.slotbad.push_back(bad);
    }
    if (plain) {
      bool extra = false;
      for (int32_t k : last_level)
        if (has_par(k)) extra = true;
      *plain = !dup && !added && !extra;
    }
    return bs.count() - 1;
  }
  static constexpr int GFX_MAXN_HOST = 64;
};

inline uint64_t pair_key(int32_t s, int32_t d) { return ((uint64_t)(uint32_t)s << 32) | (uint32_t)d; }

}  // namespace

int gfx_build_schedule(const gfx_pedigree_desc* ped, GfxHostSchedule& out, std::string& err) {
  const int np = ped->n_ped, nr = ped->n_rows, back = ped->back;
  if (np <= 0 || nr <= 0 || back < 1) { err = "n_ped, n_rows and back must be positive"; return GFX_E_INVALID; }
  std::vector<int32_t> ped_of_row(nr, -1);
  for (int a = 0; a < np; ++a) {
    const int r = ped->row[a];
    if (r >= nr) { err = "row index out of range"; return GFX_E_INVALID; }
    if (r >= 0) {
      if (ped_of_row[r] != -1) { err = "two pedigree animals map to the same genotype row"; return GFX_E_INVALID; }
      ped_of_row[r] = a;
    }
    if (ped->sire[a] >= np || ped->dam[a] >= np) { err = "parent index out of range"; return GFX_E_INVALID; }
  }
  for (int r = 0; r < nr; ++r)
    if (ped_of_row[r] < 0) { err = "genotype row " + std::to_string(r) + " has no pedigree animal"; return GFX_E_INVALID; }

  out = GfxHostSchedule();
  out.n_ped = np; out.n_rows = nr; out.back = back;
  out.bl.off.push_back(0);
  Builder B(ped);
  auto row_of = [&](int32_t a) { return a >= 0 ? ped->row[a] : -1; };

  // ---- rank: per genotyped animal -------------------------------------------------------------
  out.anc6.assign((size_t)nr * 6, -1);
  out.rank_flags.assign(nr, 0);
  out.rank_blanket.assign(nr, -1);
  out.prow.assign((size_t)nr * 2, -1);
  int n_loopy = 0, n_trios = 0, n_lone = 0;
  for (int r = 0; r < nr; ++r) {
    const int32_t a = ped_of_row[r];
    const int32_t s = ped->sire[a], d = ped->dam[a];
    out.prow[2 * r] = row_of(s);
    out.prow[2 * r + 1] = row_of(d);
    if (s >= 0 && d >= 0 && row_of(s) >= 0 && row_of(d) >= 0) { out.trio_row.push_back(r); ++n_trios; }
    if (s < 0 && d < 0) continue;  // no ancestors: anc term absent (correct_genotypes3.py:95,129-131)
    bool plain = false;
    const int id = B.add_blanket(&a, 1, back, out.bl, &plain);
    if (id < 0) { err = B.err; return B.err_code; }
    out.rank_blanket[r] = id;
    uint8_t fl = 1;
    if (!plain || back > 2) { fl |= 2; ++n_loopy; }
    out.rank_flags[r] = fl;
    int32_t* a6 = &out.anc6[(size_t)r * 6];
    a6[0] = row_of(s);
    a6[1] = row_of(d);
    if (back >= 2) {
      if (s >= 0) { a6[2] = row_of(ped->sire[s]); a6[3] = row_of(ped->dam[s]); }
      if (d >= 0) { a6[4] = row_of(ped->sire[d]); a6[5] = row_of(ped->dam[d]); }
    }
  }

  // trios grouped by family (same sire and dam), families of one sire adjacent: the scan kernel stages the
  // parents' rows once per family and the sire's row stays cache-hot across his families
  {
    std::vector<int32_t>& tr = out.trio_row;
    std::sort(tr.begin(), tr.end(), [&](int32_t x, int32_t y) {
      const int32_t sx = out.prow[2 * x], sy = out.prow[2 * y];
      if (sx != sy) return sx < sy;
      const int32_t dx = out.prow[2 * x + 1], dy = out.prow[2 * y + 1];
      if (dx != dy) return dx < dy;
      return x < y;
    });
    out.fam_off.push_back(0);
    for (size_t i = 1; i <= tr.size(); ++i)
      if (i == tr.size() || out.prow[2 * tr[i]] != out.prow[2 * tr[i - 1]] || out.prow[2 * tr[i] + 1] != out.prow[2 * tr[i - 1] + 1])
        out.fam_off.push_back((int32_t)i);
  }

  // ---- genotyped children CSR, reference iteration order ----------------------------------------
  out.child_off.assign(nr + 1, 0);
  for (int r = 0; r < nr; ++r) {
    const int32_t a = ped_of_row[r];
    for (int32_t i = ped->kid_off[a]; i < ped->kid_off[a + 1]; ++i) {
      const int32_t c = ped->kid[i];
      if (ped->row[c] < 0) continue;
      // other parent: the known parent of c that is not a (correct_genotypes3.py:261,273,403)
      int32_t other = -1, cnt = 0;
      if (ped->sire[c] >= 0 && ped->sire[c] != a) { other = ped->sire[c]; ++cnt; }
      if (ped->dam[c] >= 0 && ped->dam[c] != a) { other = ped->dam[c]; ++cnt; }
      out.child_row.push_back(ped->row[c]);
      out.child_other.push_back(cnt == 1 ? row_of(other) : -1);
    }
    out.child_off[r + 1] = (int32_t)out.child_row.size();
  }

  // ---- partner pairs, generations, singles ------------------------------------------------------
  std::unordered_set<uint64_t> seen;
  std::vector<std::pair<int32_t, int32_t>> pairs;  // pedigree indices, both genotyped
  std::vector<uint8_t> paired(np, 0);
  for (int r = 0; r < nr; ++r) {
    const int32_t a = ped_of_row[r];
    const int32_t s = ped->sire[a], d = ped->dam[a];
    if (s < 0 || d < 0) continue;
    paired[s] = paired[d] = 1;
    if (seen.insert(pair_key(s, d)).second && ped->row[s] >= 0 && ped->row[d] >= 0) pairs.push_back({s, d});
  }
  std::sort(pairs.begin(), pairs.end(), [&](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) {
    if (ped->ped_id[x.first] != ped->ped_id[y.first]) return ped->ped_id[x.first] < ped->ped_id[y.first];
    return ped->ped_id[x.second] < ped->ped_id[y.second];
  });
  int max_gen = 0;
  for (int a = 0; a < np; ++a) max_gen = std::max(max_gen, ped->generation[a]);
  const int ng = max_gen + 1;
  for (auto& pr : pairs) {
    out.pair_sire_row.push_back(ped->row[pr.first]);
    out.pair_dam_row.push_back(ped->row[pr.second]);
    const int32_t s = pr.first, d = pr.second;
    const bool s_has = ped->sire[s] >= 0 || ped->dam[s] >= 0;
    const bool d_has = ped->sire[d] >= 0 || ped->dam[d] >= 0;
    int id = -1;
    if (s_has && d_has) {  // model only if both have a known parent (correct_genotypes3.py:191)
      const int32_t f[2] = {s, d};
      id = B.add_blanket(f, 2, back + 1, out.bl, nullptr);
      if (id < 0) { err = B.err; return B.err_code; }
    }
    out.pair_blanket.push_back(id);
  }
  out.gen_off.assign(ng + 1, 0);
  out.app_gen_off.assign(ng + 1, 0);
  out.app_focal_off.push_back(0);
  int max_jobs = 0;
  for (int g = 0; g < ng; ++g) {
    const int j0 = (int)out.job_pair.size();
    for (size_t i = 0; i < pairs.size(); ++i)
      if (ped->generation[pairs[i].first] == g || ped->generation[pairs[i].second] == g)
        out.job_pair.push_back((int32_t)i);
    const int nj = (int)out.job_pair.size() - j0;
    max_jobs = std::max(max_jobs, nj);
    out.gen_off[g + 1] = (int32_t)out.job_pair.size();
    // apply lists: entries of a focal row in job order
    std::vector<std::pair<int32_t, int32_t>> ent;  // (row, t*2+side)
    ent.reserve((size_t)nj * 2);
    for (int t = 0; t < nj; ++t) {
      const int32_t p = out.job_pair[j0 + t];
      ent.push_back({out.pair_sire_row[p], t * 2});
      ent.push_back({out.pair_dam_row[p], t * 2 + 1});
    }
    std::stable_sort(ent.begin(), ent.end(), [](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) {
      return x.first < y.first;
    });
    for (size_t i = 0; i < ent.size(); ++i) {
      if (i == 0 || ent[i].first != ent[i - 1].first) {
        if (i != 0) out.app_focal_off.push_back((int32_t)out.app_entry.size());
        out.app_focal_row.push_back(ent[i].first);
      }
      out.app_entry.push_back(ent[i].second);
    }
    if (!ent.empty()) out.app_focal_off.push_back((int32_t)out.app_entry.size());
    out.app_gen_off[g + 1] = (int32_t)out.app_focal_row.size();
    // heaviest animals first: cost ~ jobs * (blanket walk) + genotyped children
    {
      const int f0 = out.app_gen_off[g], nf = out.app_gen_off[g + 1] - f0;
      std::vector<std::pair<int64_t, int32_t>> wgt(nf);
      for (int i = 0; i < nf; ++i) {
        const int32_t r = out.app_focal_row[f0 + i];
        const int64_t jobs = out.app_focal_off[f0 + i + 1] - out.app_focal_off[f0 + i];
        wgt[i] = {-(jobs * 12 + (out.child_off[r + 1] - out.child_off[r])), i};
      }
      std::sort(wgt.begin(), wgt.end());
      for (int i = 0; i < nf; ++i) out.app_order.push_back(wgt[i].second);
    }
  }
  for (int r = 0; r < nr; ++r) {
    const int32_t a = ped_of_row[r];
    if (paired[a]) continue;
    // an animal with neither a blanket nor genotyped children can never yield a result
    // (correct_genotypes3.py:417-419 returns None for it): leave it out of the job list
    if (out.rank_blanket[r] < 0 && out.child_off[r + 1] == out.child_off[r]) { ++n_lone; continue; }
    out.single_row.push_back(r);
    out.single_blanket.push_back(out.rank_blanket[r]);
  }

  {
    const int ns = (int)out.single_row.size();
    std::vector<std::pair<int64_t, int32_t>> wgt(ns);
    for (int i = 0; i < ns; ++i) {
      const int32_t r = out.single_row[i];
      wgt[i] = {-(int64_t)(out.child_off[r + 1] - out.child_off[r]), i};
    }
    std::sort(wgt.begin(), wgt.end());
    for (int i = 0; i < ns; ++i) out.single_order.push_back(wgt[i].second);
  }
  gfx_schedule_info_t& I = out.info;
  I.n_rows = nr;
  I.n_generations = ng;
  I.n_pairs = (int32_t)pairs.size();
  I.n_pair_jobs = (int32_t)out.job_pair.size();
  I.n_singles = (int32_t)out.single_row.size();
  I.n_loopy = n_loopy;
  I.n_blankets = out.bl.count();
  I.max_blanket_nodes = out.bl.max_nodes;
  I.max_width = out.bl.max_width;
  I.max_pairs_in_generation = max_jobs;
  I.n_trios = n_trios;
  I.n_lone = n_lone;
  return GFX_OK;
}
