// CUDA kernels (sm_100a) + C ABI of genofix_b200.  See include/genofix_b200.h for the contract and
// DESIGN.md for the data layout and the roofline of each kernel.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/genofix_b200.h"
#include "gfx_infer.cuh"
#include "gfx_schedule.h"

#define GFX_BIG_SLOTS 64
#define GFX_HUGE_SLOTS 4

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define GFX_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(GFX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)
#define GFX_LAUNCHED()                                                                   \
  do {                                                                                   \
    ++g_launches;                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) return fail(GFX_E_CUDA, std::string("launch: ") + cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* gfx_last_error(void) { return g_err.c_str(); }
extern "C" int gfx_version(void) { return 100; }
extern "C" int64_t gfx_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// device-side schedule
// ------------------------------------------------------------------------------------------------
// per-row record of the rank kernel, heaviest rows (most children) first: one 32-byte load per task
struct RankRow {
  int32_t row, c0, nk, fl;   // fl: bit0 has ancestors, bit1 loopy blanket
  int32_t rS, rD, pad0, pad1;
};

struct gfx_schedule {
  GfxHostSchedule h;
  int device = 0;
  int num_sms = 148;
  // device copies
  int32_t *anc6 = nullptr, *rank_blanket = nullptr, *prow = nullptr;
  uint8_t* rank_flags = nullptr;
  int32_t *child_off = nullptr, *child_row = nullptr, *child_other = nullptr;
  int32_t* rank_kids = nullptr;      // children rows per rank row, every list starting at a multiple of 4 entries (16-byte bulk copies)
  int32_t *bl_off = nullptr, *bl_row = nullptr;
  int8_t *bl_p1 = nullptr, *bl_p2 = nullptr, *bl_last = nullptr, *bl_q0 = nullptr, *bl_q1 = nullptr;
  uint64_t* bl_kid = nullptr;
  uint8_t* bl_tree = nullptr;
  int32_t* bl_slotrow = nullptr;
  uint16_t* bl_slotbad = nullptr;
  int32_t *job_row0 = nullptr, *job_row1 = nullptr, *job_bl = nullptr;        // pair jobs, all generations
  int32_t *app_focal_row = nullptr, *app_focal_off = nullptr, *app_entry = nullptr, *app_order = nullptr;
  int32_t* single_order = nullptr;
  int32_t *single_row = nullptr, *single_bl = nullptr, *single_off = nullptr, *single_entry = nullptr;
  int32_t *loopy_rows = nullptr, *trio_row = nullptr, *fam_off = nullptr;
  int n_loopy_rows = 0;
  int debug = 0;
  uint2* rank_lut = nullptr;   // 4096 entries: f16 scores for obs = 0,1,2 (+ pad)
  uint32_t* rank_lut6 = nullptr;     // 64 entries (obs | sire << 2 | dam << 4): score | class << 16
  uint16_t* rank_class_sc = nullptr; // score of each of the <= 8 classes
  // work lists of k_mendel_rank3 / k_mendel_rank_generic_wl: a ring of 8 (lanes share the schedule, a few launches are in
  // flight at a time); a slot is handed out again only behind the event of its last consumer.  Headers are all zero at rest.
  struct WlSlot { uint2* buf = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool used = false; };
  WlSlot rank_wl[8];
  uint32_t* rank_wl_hdr = nullptr;
  uint32_t rank_wl_next = 0;
  cudaEvent_t rank_ev0 = nullptr, rank_ev1 = nullptr;   // gfx_rank_timing: recorded around the dominant rank kernel
  bool rank_timing = false;
  std::atomic<int> too_wide{0};         // sticky: a GFX_E_TOO_WIDE seen by any lane of this schedule fails every later check
  std::atomic<int> rank_form_pref{0};   // gfx_rank_form: 0 = third form when possible, 2 = second form
  bool rank_form3 = false;           // the score table is the one k_mendel_rank3's boolean network computes
  int rank_ncls = 0;
  std::vector<RankRow> rank_rows_h;     // per-row records of the rank kernel, heaviest first
  struct RankTasks { int64_t nwords; int form; RankRow* dev; uint32_t n; };
  std::vector<RankTasks> rank_tasks;    // task tables, one per matrix width seen so far
  std::mutex rank_mutex;
  double* kv = nullptr;        // 108 doubles, children-heuristic element table
  int* status = nullptr;
  unsigned long long* counters = nullptr;  // 16 diagnostic counters (gfx_debug_counters)
  double* big_pool = nullptr;  // GFX_BIG_SLOTS tables of GFX_TAB_BIG doubles for rare wide blankets
  int* big_flags = nullptr;    // GFX_BIG_SLOTS flags, then GFX_HUGE_SLOTS flags
  double* huge_pool = nullptr; // GFX_HUGE_SLOTS tables of GFX_TAB_HUGE doubles
};

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  const size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)dst, bytes);
  if (e != cudaSuccess) return e;
  if (v.size()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// Table for the plain 2-generation tree: index = 2-bit codes of S,D,SS,SD,DS,DD (3 = unobserved).
static void build_rank_lut(std::vector<uint2>& lut) {
  lut.resize(4096);
  // topological order: SS SD S DS DD D K
  const int8_t p1[7] = {-1, -1, 0, -1, -1, 3, 2};
  const int8_t p2[7] = {-1, -1, 1, -1, -1, 4, 5};
  const int8_t last[7] = {2, 2, 6, 5, 5, 6, 6};
  GfxBlanket b{7, p1, p2, last};
  double tab[GFX_TAB];
  for (int idx = 0; idx < 4096; ++idx) {
    const int c[6] = {idx & 3, (idx >> 2) & 3, (idx >> 4) & 3, (idx >> 6) & 3, (idx >> 8) & 3, (idx >> 10) & 3};
    uint8_t code[7] = {(uint8_t)c[2], (uint8_t)c[3], (uint8_t)c[0], (uint8_t)c[4], (uint8_t)c[5], (uint8_t)c[1], 3};
    uint64_t obs = 0;
    for (int v = 0; v < 6; ++v)
      if (code[v] < 3) obs |= 1ull << v;
    double anc[3];
    gfx_infer(b, 6, code, obs, tab, anc);
    uint16_t s[3];
    for (int o = 0; o < 3; ++o) s[o] = gfx_anc_score(anc, o);
    lut[idx] = make_uint2((uint32_t)s[0] | ((uint32_t)s[1] << 16), (uint32_t)s[2]);
  }
}

// Fast table of the rank kernel.  With both parents observed the grand-parents are d-separated from the focal
// animal, so the ancestor score depends on (obs, sire, dam) only: idx6 = obs | sire << 2 | dam << 4.  This is
// CHECKED here against all 256 grand-parent variants of the full table.  Distinct scores get a class id
// (class 0 = score 0) for the (class, m) histogram; obs = 3 maps to the marker 0xFFFF.
static bool build_rank_fast(const std::vector<uint2>& lut, std::vector<uint32_t>& lut6, std::vector<uint16_t>& class_sc,
                            int& ncls_out, std::string& err) {
  lut6.assign(64, 0);
  class_sc.assign(8, 0);
  int ncls = 1;
  auto pick = [](const uint2& e, int obs) -> uint16_t {
    return obs == 0 ? (uint16_t)(e.x & 0xFFFFu) : (obs == 1 ? (uint16_t)(e.x >> 16) : (uint16_t)(e.y & 0xFFFFu));
  };
  for (int idx = 0; idx < 64; ++idx) {
    const int obs = idx & 3, sv = (idx >> 2) & 3, dv = (idx >> 4) & 3;
    uint16_t sc;
    if (obs == 3) sc = 0xFFFFu;
    else if (sv == 3 || dv == 3) continue;   // not reachable on the fast path
    else {
      sc = pick(lut[sv | (dv << 2) | 0xFF0], obs);
      for (int gp = 0; gp < 256; ++gp)
        if (pick(lut[sv | (dv << 2) | (gp << 4)], obs) != sc) { err = "rank table depends on grand-parents below observed parents"; return false; }
      // exactness precondition of the packed-half path (half add == float32 add + one rounding):
      // finite scores are non-negative multiples of 2^-13 below 16
      const float f = gfx_h2f(sc);
      if (f == f && (f < 0.f || f >= 16.f || f * 8192.f != (float)(int)(f * 8192.f))) { err = "rank table score outside the exact range"; return false; }
    }
    int c = -1;
    for (int i = 0; i < ncls; ++i) if (class_sc[i] == sc) c = i;
    if (c < 0) {
      if (ncls == 8) { err = "more than 8 distinct scores in the fast rank table"; return false; }
      class_sc[ncls] = sc;
      c = ncls++;
    }
    lut6[idx] = (uint32_t)sc | ((uint32_t)c << 16);
  }
  ncls_out = ncls;
  return true;
}

// k_mendel_rank3 computes the score class of a cell with both parents observed from a fixed boolean network over the
// six code bits (classes 0, 0.5, 1, 2 -- the values 2 * (max - own) takes on Mendel's table, bayesnet/model.py:27-38).
// This is the same network on the host; the schedule checks it against the table built by inference and only then
// lets the kernel use it.
static int rank3_class(int o, int sv, int dv) {
  const int o0 = o & 1, o1 = o >> 1, s0 = sv & 1, s1 = sv >> 1, d0 = dv & 1, d1 = dv >> 1;
  const int u = o0 ^ s1 ^ d1, v = o1 ^ (s1 & d1), t3 = (u | v) & (s0 ^ 1), aa = s1 | d1;
  const int q = (aa & ((o0 | o1) ^ 1)) | ((aa ^ 1) & o1);
  const int c2 = q & (s0 ^ d0), c1 = s0 & d0 & (o0 ^ 1);
  return ((t3 & (d0 ^ 1)) | c1) | (((t3 & (d0 ^ 1)) | c2) << 1);
}
static bool rank3_network_matches(const std::vector<uint32_t>& lut6) {
  static const uint16_t sc[4] = {0x0000u, 0x3800u, 0x3C00u, 0x4000u};
  for (int o = 0; o < 3; ++o)
    for (int sv = 0; sv < 3; ++sv)
      for (int dv = 0; dv < 3; ++dv)
        if ((lut6[o | (sv << 2) | (dv << 4)] & 0xFFFFu) != sc[rank3_class(o, sv, dv)]) return false;
  return true;
}

extern "C" void gfx_schedule_destroy(gfx_schedule* s) {
  if (!s) return;
  void* ptrs[] = {s->anc6, s->rank_blanket, s->prow, s->rank_flags, s->child_off, s->child_row, s->child_other,
                  s->bl_off, s->bl_row, s->bl_p1, s->bl_p2, s->bl_last, s->bl_q0, s->bl_q1, s->job_row0, s->job_row1,
                  s->job_bl, s->app_focal_row, s->app_focal_off, s->app_entry, s->single_row, s->single_bl,
                  s->single_off, s->single_entry, s->loopy_rows, s->trio_row, s->rank_lut, s->kv, s->status,
                  s->big_pool, s->big_flags, s->huge_pool, s->bl_kid, s->counters, s->bl_tree, s->app_order, s->single_order, s->bl_slotrow, s->bl_slotbad, s->fam_off, s->rank_lut6, s->rank_class_sc, s->rank_kids, s->rank_wl_hdr};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& t : s->rank_tasks)
    if (t.dev) cudaFree(t.dev);
  for (auto& w : s->rank_wl) {
    if (w.buf) cudaFree(w.buf);
    if (w.done) cudaEventDestroy(w.done);
  }
  if (s->rank_ev0) cudaEventDestroy(s->rank_ev0);
  if (s->rank_ev1) cudaEventDestroy(s->rank_ev1);
  delete s;
}

extern "C" int gfx_schedule_create(const gfx_pedigree_desc* ped, gfx_schedule** out) {
  if (!ped || !out) return fail(GFX_E_INVALID, "null argument");
  *out = nullptr;
  gfx_schedule* s = new (std::nothrow) gfx_schedule();
  if (!s) return fail(GFX_E_NOMEM, "out of host memory");
  std::string err;
  int rc = gfx_build_schedule(ped, s->h, err);
  if (rc != GFX_OK) { delete s; return fail(rc, err); }
  GfxHostSchedule& h = s->h;
  cudaError_t e = cudaGetDevice(&s->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, s->device);
  std::vector<int32_t> job_row0, job_row1, job_bl;
  for (int32_t p : h.job_pair) {
    job_row0.push_back(h.pair_sire_row[p]);
    job_row1.push_back(h.pair_dam_row[p]);
    job_bl.push_back(h.pair_blanket[p]);
  }
  std::vector<int32_t> single_off(h.single_row.size() + 1), single_entry(h.single_row.size());
  for (size_t i = 0; i < h.single_row.size(); ++i) { single_off[i] = (int32_t)i; single_entry[i] = (int32_t)i; }
  single_off[h.single_row.size()] = (int32_t)h.single_row.size();
  std::vector<int32_t> loopy;
  for (int r = 0; r < h.n_rows; ++r)
    if (h.rank_flags[r] & 2) loopy.push_back(r);
  s->n_loopy_rows = (int)loopy.size();
  std::vector<uint2> lut;
  build_rank_lut(lut);
  std::vector<uint32_t> lut6;
  std::vector<uint16_t> class_sc;
  if (!build_rank_fast(lut, lut6, class_sc, s->rank_ncls, err)) { delete s; return fail(GFX_E_INVALID, err); }
  s->rank_form3 = rank3_network_matches(lut6);
  std::vector<RankRow>& rank_rows = s->rank_rows_h;
  rank_rows.resize(h.n_rows);
  std::vector<int32_t> rank_kids;
  {
    std::vector<int32_t> order(h.n_rows);
    for (int r = 0; r < h.n_rows; ++r) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
      return h.child_off[x + 1] - h.child_off[x] > h.child_off[y + 1] - h.child_off[y];
    });
    for (int i = 0; i < h.n_rows; ++i) {
      const int r = order[i];
      const int nk = h.child_off[r + 1] - h.child_off[r];
      rank_rows[i] = RankRow{r, (int32_t)rank_kids.size(), nk, h.rank_flags[r],
                             h.anc6[(size_t)r * 6], h.anc6[(size_t)r * 6 + 1], 0, 0};
      for (int c = 0; c < nk; ++c) rank_kids.push_back(h.child_row[h.child_off[r] + c]);
      while (rank_kids.size() % 4) rank_kids.push_back(-1);
    }
    for (int c = 0; c < 32; ++c) rank_kids.push_back(-1);
  }
  std::vector<double> kv(108);
  gfx_fill_kv(kv.data());
  std::vector<int> status(1, 0);
#define UP(field, vec) if (e == cudaSuccess) e = upload(&s->field, vec)
  UP(anc6, h.anc6); UP(rank_blanket, h.rank_blanket); UP(prow, h.prow); UP(rank_flags, h.rank_flags);
  UP(child_off, h.child_off); UP(child_row, h.child_row); UP(child_other, h.child_other);
  UP(bl_off, h.bl.off); UP(bl_row, h.bl.row); UP(bl_p1, h.bl.p1); UP(bl_p2, h.bl.p2); UP(bl_last, h.bl.last);
  UP(bl_q0, h.bl.q0); UP(bl_q1, h.bl.q1); UP(bl_kid, h.bl.kidmask); UP(bl_tree, h.bl.tree); UP(bl_slotrow, h.bl.slotrow); UP(bl_slotbad, h.bl.slotbad);
  UP(job_row0, job_row0); UP(job_row1, job_row1); UP(job_bl, job_bl);
  UP(app_focal_row, h.app_focal_row); UP(app_focal_off, h.app_focal_off); UP(app_entry, h.app_entry);
  UP(single_row, h.single_row); UP(single_bl, h.single_blanket); UP(single_off, single_off);
  UP(single_entry, single_entry);
  UP(app_order, h.app_order); UP(single_order, h.single_order);
  UP(loopy_rows, loopy); UP(trio_row, h.trio_row); UP(fam_off, h.fam_off); UP(rank_lut, lut); UP(rank_lut6, lut6); UP(rank_class_sc, class_sc); UP(kv, kv); UP(status, status);
  UP(rank_kids, rank_kids);
#undef UP
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->big_pool, sizeof(double) * (size_t)GFX_TAB_BIG * GFX_BIG_SLOTS);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->big_flags, sizeof(int) * (GFX_BIG_SLOTS + GFX_HUGE_SLOTS));
  if (e == cudaSuccess) e = cudaMemset(s->big_flags, 0, sizeof(int) * (GFX_BIG_SLOTS + GFX_HUGE_SLOTS));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->huge_pool, sizeof(double) * (size_t)GFX_TAB_HUGE * GFX_HUGE_SLOTS);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->rank_wl_hdr, sizeof(uint32_t) * 64 * 32);
  if (e == cudaSuccess) e = cudaMemset(s->rank_wl_hdr, 0, sizeof(uint32_t) * 64 * 32);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->counters, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess) e = cudaMemset(s->counters, 0, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) {
    gfx_schedule_destroy(s);
    return fail(GFX_E_CUDA, std::string("schedule upload: ") + cudaGetErrorString(e));
  }
  *out = s;
  return GFX_OK;
}

// Host-only variant (no CUDA calls): builds the index arrays so that gfx_schedule_info / _copy can be
// inspected on a machine without a GPU.  The handle must not be passed to any kernel entry point.
extern "C" int gfx_schedule_create_host(const gfx_pedigree_desc* ped, gfx_schedule** out) {
  if (!ped || !out) return fail(GFX_E_INVALID, "null argument");
  *out = nullptr;
  gfx_schedule* s = new (std::nothrow) gfx_schedule();
  if (!s) return fail(GFX_E_NOMEM, "out of host memory");
  std::string err;
  int rc = gfx_build_schedule(ped, s->h, err);
  if (rc != GFX_OK) { delete s; return fail(rc, err); }
  s->device = -1;
  *out = s;
  return GFX_OK;
}

extern "C" int gfx_schedule_info(const gfx_schedule* s, gfx_schedule_info_t* out) {
  if (!s || !out) return fail(GFX_E_INVALID, "null argument");
  *out = s->h.info;
  return GFX_OK;
}

extern "C" int gfx_schedule_copy(const gfx_schedule* s, int which, int32_t* out, int64_t cap) {
  if (!s || !out) return fail(GFX_E_INVALID, "null argument");
  const GfxHostSchedule& h = s->h;
  std::vector<int32_t> tmp;
  const std::vector<int32_t>* src = nullptr;
  switch (which) {
    case 0:
      for (size_t i = 0; i < h.pair_sire_row.size(); ++i) { tmp.push_back(h.pair_sire_row[i]); tmp.push_back(h.pair_dam_row[i]); }
      src = &tmp; break;
    case 1: src = &h.gen_off; break;
    case 2: src = &h.single_row; break;
    case 3: src = &h.job_pair; break;
    case 4: for (uint8_t f : h.rank_flags) tmp.push_back(f); src = &tmp; break;
    default: return fail(GFX_E_INVALID, "unknown array selector");
  }
  if ((int64_t)src->size() > cap) return fail(GFX_E_INVALID, "output buffer too small");
  if (!src->empty()) memcpy(out, src->data(), src->size() * sizeof(int32_t));
  return (int)src->size();
}

extern "C" int gfx_debug_counters(gfx_schedule* s, int enable, uint64_t* out16_host) {
  if (!s || s->device < 0) return fail(GFX_E_INVALID, "bad schedule");
  if (out16_host) GFX_CUDA(cudaMemcpy(out16_host, s->counters, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost));
  GFX_CUDA(cudaMemset(s->counters, 0, sizeof(uint64_t) * 16));
  s->debug = enable;
  return GFX_OK;
}

extern "C" int gfx_check_device_status(const gfx_schedule* s, void* stream) {
  if (!s) return fail(GFX_E_INVALID, "null argument");
  int st = 0;
  GFX_CUDA(cudaMemcpyAsync(&st, s->status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GFX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  gfx_schedule* ms = const_cast<gfx_schedule*>(s);
  if (st != 0) {
    int zero = 0;
    GFX_CUDA(cudaMemcpy(s->status, &zero, sizeof(int), cudaMemcpyHostToDevice));
    ms->too_wide.store(1);
  }
  // lanes share the schedule and its one status word: the lane that reads the word first clears it, so the flag stays up on
  // the host side and every lane's check fails from then on (no chunk with unevaluated cells is handed out as valid)
  if (ms->too_wide.load())
    return fail(GFX_E_TOO_WIDE, "a blanket needed more than GFX_WMAX_HUGE (13) simultaneously unobserved variables");
  return GFX_OK;
}

extern "C" int gfx_reset_device_status(gfx_schedule* s) {
  if (!s || s->device < 0) return fail(GFX_E_INVALID, "bad schedule");
  int zero = 0;
  GFX_CUDA(cudaMemcpy(s->status, &zero, sizeof(int), cudaMemcpyHostToDevice));
  s->too_wide.store(0);
  return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
// pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void k_pack2(const uint8_t* __restrict__ codes, int64_t n, int64_t s, int64_t ld, uint32_t* __restrict__ packed,
                        int64_t wpr, bool vec) {
  const int64_t total = n * wpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / wpr, w = i - r * wpr;
    const int64_t j0 = w * 16;
    uint32_t out = 0;
    const uint8_t* src = codes + r * ld + j0;
    if (vec && j0 + 16 <= s) {
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t c = (q[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
        out |= (c > 2u ? 3u : c) << (2 * k);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        uint32_t c = 3u;
        if (j0 + k < s) { c = src[k]; if (c > 2u) c = 3u; }
        out |= c << (2 * k);
      }
    }
    packed[i] = out;
  }
}

__global__ void k_unpack2(const uint32_t* __restrict__ packed, int64_t n, int64_t s, int64_t wpr,
                          uint8_t* __restrict__ codes, int64_t ld, bool vec) {
  const int64_t total = n * wpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / wpr, w = i - r * wpr;
    const int64_t j0 = w * 16;
    if (j0 >= s) continue;
    const uint32_t p = packed[i];
    uint8_t* dst = codes + r * ld + j0;
    if (vec && j0 + 16 <= s) {
      uint32_t q[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c = (p >> (2 * (4 * g + k))) & 3u;
          o |= (c == 3u ? 9u : c) << (8 * k);
        }
        q[g] = o;
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(q[0], q[1], q[2], q[3]);
    } else {
      for (int k = 0; k < 16 && j0 + k < s; ++k) {
        const uint32_t c = (p >> (2 * k)) & 3u;
        dst[k] = (uint8_t)(c == 3u ? 9u : c);
      }
    }
  }
}

static int grid_for(int64_t work_items, int threads, int num_sms, int per_sm) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

extern "C" int gfx_pack2(const uint8_t* codes, int64_t n, int64_t s, int64_t ld, uint32_t* packed, int64_t wpr, void* stream) {
  if (!codes || !packed || n <= 0 || s <= 0 || wpr * 16 < s || ld < s) return fail(GFX_E_INVALID, "gfx_pack2: bad shape");
  const bool vec = (ld % 16 == 0) && (((uintptr_t)codes) % 16 == 0);
  k_pack2<<<grid_for(n * wpr, 256, device_sms(), 16), 256, 0, (cudaStream_t)stream>>>(codes, n, s, ld, packed, wpr, vec);
  GFX_LAUNCHED();
  return GFX_OK;
}

extern "C" int gfx_unpack2(const uint32_t* packed, int64_t n, int64_t s, int64_t wpr, uint8_t* codes, int64_t ld, void* stream) {
  if (!codes || !packed || n <= 0 || s <= 0 || wpr * 16 < s || ld < s) return fail(GFX_E_INVALID, "gfx_unpack2: bad shape");
  const bool vec = (ld % 16 == 0) && (((uintptr_t)codes) % 16 == 0);
  k_unpack2<<<grid_for(n * wpr, 256, device_sms(), 16), 256, 0, (cudaStream_t)stream>>>(packed, n, s, wpr, codes, ld, vec);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
// Mendel rank kernel, fused with the histogram of the scores
// ------------------------------------------------------------------------------------------------
// One thread = one packed word = 16 SNPs of one animal; a warp takes (row, 128-word chunk) tasks from a global
// counter and covers 32 consecutive words per iteration (128 B of codes in, 1 KB of scores out).
//
// Fast path (both parents genotyped and observed in all 16 cells of the word): the grand-parents are
// d-separated, the score is a function of (obs, sire, dam).  The three words are transposed with 15 bitwise
// operations into bytes obs | sire << 2 | dam << 4, one per cell, and looked up in a LANE-PRIVATE table
// s_fast[idx6][lane] -- address formed by one PRMT, no bank conflicts.  The children term (m units of
// float16(1/3), counted bit-parallel) is added with packed-half arithmetic: child = fma(1024 + m, c, -1024 c)
// (one rounding, = float16(m * 1365 / 4096)), score = child + anc (half add of multiples of 2^-13 below 2^10
// is float32 add + one rounding: what numpy does).
// The histogram is fused: cells without children are counted per idx6 in lane-private counters next to the
// table entry, cells with children per (score class, m) -- both conflict-free -- and mapped to float16
// patterns when the block retires.  Anything else (missing parent cells, ragged last word, m >= 32, founders
// with missing data) goes through the per-cell slow path with the full 4096-entry table.
struct RankArgs {
  const uint32_t* packed;
  int64_t wpr;        // words per row
  int nwords;         // words that hold SNPs: ceil(n_snps/16)
  int64_t n_snps;
  uint16_t* raw;
  int64_t ld_raw;
  int n_rows;
  const int32_t* anc6;
  const RankRow* tasks;       // one record per task (pad0 = first word, pad1 = end word)
  uint32_t n_tasks;
  const int32_t* child_row;
  const uint2* lut;
  const uint32_t* lut6;
  const uint16_t* class_sc;
  unsigned long long* hist;   // 65537 bins or nullptr (HIST = false)
  uint32_t* wl_hdr;           // k_mendel_rank3: work list of k_mendel_rank_generic_wl: [0] number of entries (zero at launch), or nullptr
  uint2* wl;                  //                 its (row, word) entries
  uint32_t wl_cap;
};

// 896 threads per SM: ptxas gets 72 registers instead of 64, which ends the spilling of the prefetched words and the
// per-step re-derivation of the row pointers (1024 threads: 100.3 us per 10k x 10k chunk, 896: 93.1 us; 768: 94.0 us;
// 640 / 512: 99.8 / 103.5 us -- profiles/r1_rank_variants_{a,b}.jsonl)
#ifndef R2_THREADS
#define R2_THREADS 896
#endif
// words per task by number of children of the row (measured at 10k x 10k and 100k x 4096)
#ifndef R2_WORDS_HEAVY
#define R2_WORDS_HEAVY 32    // >= 8 children
#endif
#ifndef R2_WORDS_MID
#define R2_WORDS_MID 128     // 1..7 children
#endif
#ifndef R2_WORDS_LIGHT
#define R2_WORDS_LIGHT 128   // childless
#endif
#define R2_CM_OFF 0u          // u32 [8 classes][32 m][32 lanes]
#define R2_FAST_OFF 32768u    // u32 [64 idx6][64]: lanes 0..31 table entries, 32..63 counters
#define R2_BINS_OFF 49152u    // u32 [8192] patterns 0x3000..0x4FFF
#define R2_SMEM 81920u
#define R2_BIN_LO 0x3000u
#define R2_BIN_N 0x2000u

__device__ __forceinline__ uint32_t ldg_word(const uint32_t* p, int64_t row, int64_t wpr, int64_t w) {
  return row >= 0 ? __ldg(p + row * wpr + w) : 0xFFFFFFFFu;
}
__device__ __forceinline__ __half2 r2_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t r2_u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void r2_store(uint16_t* dst, const uint32_t* o) {
  __stcs(reinterpret_cast<uint4*>(dst), make_uint4(o[0], o[1], o[2], o[3]));
  __stcs(reinterpret_cast<uint4*>(dst) + 1, make_uint4(o[4], o[5], o[6], o[7]));
}

// children of a row at word w: per cell m = sum over genotyped children of the float16 difference in units of
// 1/3 (bayesnet/model.py:93-116 with the other parent unknown, D7): child 0 -> [0,1,2][obs], child 1 -> 0,
// child 2 -> [2,1,0][obs].  m16[k] = m(cell 2k) | m(cell 2k+1) << 16; anyv bit 2i: cell i has a genotyped child.
// Counted bit-parallel: 2-bit contributions -> nibble lanes (7 children) -> byte lanes (17 flushes) -> 16-bit lanes.
__device__ __forceinline__ void r2_children(const RankArgs& a, int c0, int c1, int64_t w, uint32_t w0, uint32_t m16[8],
                                            uint32_t& anyv) {
  const uint32_t lo = w0, hi = w0 >> 1;
  const uint32_t o0 = ~(lo | hi) & 0x55555555u;
  const uint32_t o1 = (lo & ~hi) & 0x55555555u;
  const uint32_t o2 = (hi & ~lo) & 0x55555555u;
#pragma unroll
  for (int k = 0; k < 8; ++k) m16[k] = 0;
  uint32_t bl0 = 0, bl1 = 0, bh0 = 0, bh1 = 0;   // byte t: cells 4t, 4t + 2, 4t + 1, 4t + 3
  int pend8 = 0;
  anyv = 0;
  // four children per round; the next round's rows are loaded while the current round is counted (absent children
  // read as all-missing).  The loop is bound by the latency of these loads, not by the arithmetic.
  uint32_t nx[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = c0 + u < c1 ? __ldg(a.child_row + c0 + u) : -1;
    nx[u] = r >= 0 ? __ldg(a.packed + (int64_t)r * a.wpr + w) : 0xFFFFFFFFu;
  }
  for (int c = c0; c < c1; c += 4) {
    uint32_t wc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wc[u] = nx[u];
    if (c + 4 < c1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = c + 4 + u < c1 ? __ldg(a.child_row + c + 4 + u) : -1;
        nx[u] = r >= 0 ? __ldg(a.packed + (int64_t)r * a.wpr + w) : 0xFFFFFFFFu;
      }
    }
    uint32_t acc_lo = 0, acc_hi = 0;          // nibble n: cell 2n / cell 2n + 1 (4 children * 2 units = 8 < 16)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t cl = wc[u], ch = wc[u] >> 1;
      const uint32_t is0 = ~(cl | ch) & 0x55555555u;
      const uint32_t is2 = (ch & ~cl) & 0x55555555u;
      anyv |= ~(cl & ch);
      const uint32_t one = (is0 | is2) & o1;
      const uint32_t two = (is2 & o0) | (is0 & o2);
      const uint32_t contrib = one | (two << 1);  // 2-bit lanes holding 0,1,2
      acc_lo += contrib & 0x33333333u;
      acc_hi += (contrib >> 2) & 0x33333333u;
    }
    bl0 += acc_lo & 0x0F0F0F0Fu; bl1 += (acc_lo >> 4) & 0x0F0F0F0Fu;
    bh0 += acc_hi & 0x0F0F0F0Fu; bh1 += (acc_hi >> 4) & 0x0F0F0F0Fu;
    if (++pend8 == 31) {                        // 31 * 8 = 248 < 256
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        m16[2 * t] += __byte_perm(bl0, bh0, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
        m16[2 * t + 1] += __byte_perm(bl1, bh1, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
      }
      bl0 = bl1 = bh0 = bh1 = 0;
      pend8 = 0;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    // halves: low = byte t of bl, high = byte t of bh
    m16[2 * t] += __byte_perm(bl0, bh0, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
    m16[2 * t + 1] += __byte_perm(bl1, bh1, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
  }
  anyv &= 0x55555555u;
}

// histogram of one word's final patterns (words whose cells do not fit the lane-private counters)
// zeros and missing cells are tallied in two scratch bins (patterns that cannot be scores) and mapped when the block retires
#define R2_SLOT_ZERO (R2_BIN_N - 1u)   // 0x4FFF
#define R2_SLOT_MISS (R2_BIN_N - 2u)   // 0x4FFE
__device__ __noinline__ void r2_hist_patterns(unsigned long long* hist, uint32_t* s_bins, const uint32_t outw[8], uint32_t miss,
                                              int lim) {
  uint32_t zeros = 0, nmiss = 0;
#pragma unroll 1
  for (int i = 0; i < lim; ++i) {
    const uint32_t r = (outw[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
    if ((miss >> (2 * i)) & 1u) ++nmiss;
    else if (r == 0u) ++zeros;
    else if (r - R2_BIN_LO < R2_BIN_N - 2u) atomicAdd(&s_bins[r - R2_BIN_LO], 1u);
    else atomicAdd(&hist[r], 1ull);
  }
  if (zeros) atomicAdd(&s_bins[R2_SLOT_ZERO], zeros);
  if (nmiss) atomicAdd(&s_bins[R2_SLOT_MISS], nmiss);
}

// Per-cell path with the full table; handles every case (and is the definition the fast path is tested against).
template <bool HIST>
__device__ __noinline__ void r2_slow_word(const RankArgs& a, uint32_t* s_bins, int64_t row, int64_t w, uint32_t w0, bool has_anc,
                                          int c0, int c1) {
  uint32_t wa[6];
  if (has_anc) {
    const int32_t* a6 = a.anc6 + row * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) wa[k] = ldg_word(a.packed, __ldg(a6 + k), a.wpr, w);
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) wa[k] = 0xFFFFFFFFu;
  }
  uint32_t m16[8], anyv;
  r2_children(a, c0, c1, w, w0, m16, anyv);
  const int lim = (int)min((int64_t)16, a.n_snps - w * 16);
  uint32_t outw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t pair = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * k + h;
      const uint32_t obs = (w0 >> (2 * i)) & 3u;
      uint32_t r = 0x3C00u;  // 1.0: unobserved, or neither ancestors nor usable children (:102)
      if (obs != 3u) {
        const bool kids = (anyv >> (2 * i)) & 1u;
        const uint32_t m = (m16[k] >> (16 * h)) & 0xFFFFu;
        if (has_anc) {
          const uint32_t idx = ((wa[0] >> (2 * i)) & 3u) | (((wa[1] >> (2 * i)) & 3u) << 2) |
                               (((wa[2] >> (2 * i)) & 3u) << 4) | (((wa[3] >> (2 * i)) & 3u) << 6) |
                               (((wa[4] >> (2 * i)) & 3u) << 8) | (((wa[5] >> (2 * i)) & 3u) << 10);
          const uint2 e = __ldg(a.lut + idx);   // 32 KB table, L1 resident; only this path needs it
          const uint16_t sc = obs == 0 ? (uint16_t)(e.x & 0xFFFFu) : (obs == 1 ? (uint16_t)(e.x >> 16) : (uint16_t)(e.y & 0xFFFFu));
          r = kids ? gfx_half_add(gfx_child_score(m), sc) : sc;
        } else if (kids) {
          r = gfx_child_score(m);
        }
      } else if (HIST) {
        r = 0xFFFFu;   // E = 1 by definition (:603)
      }
      pair |= r << (16 * h);
    }
    outw[k] = pair;
  }
  if (HIST)
    r2_hist_patterns(a.hist, s_bins, outw, w0 & (w0 >> 1) & 0x55555555u, lim);
  r2_store(a.raw + row * a.ld_raw + w * 16, outw);
}

#ifndef R2_CTAS
#define R2_CTAS (R2_THREADS <= 512 ? 2 : 1)
#endif
#define R2_DYN_SMEM (R2_SMEM + 16u)
template <bool HIST>
__global__ void __launch_bounds__(R2_THREADS, R2_CTAS) k_mendel_rank(RankArgs a) {
  extern __shared__ __align__(16) char smem[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    uint32_t* s_fast = reinterpret_cast<uint32_t*>(smem + R2_FAST_OFF);
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
      const int idx = i >> 6, l = i & 63;
      uint32_t v = 0;
      if (l < 32) {
        const uint32_t e = a.lut6[idx];
        uint32_t sc = e & 0xFFFFu;
        if (!HIST && sc == 0xFFFFu) sc = 0x3C00u;   // plain mode: unobserved cells score 1.0 (:102)
        v = sc | ((R2_CM_OFF + (e >> 16) * 4096u + (uint32_t)l * 4u) << 16);
      }
      s_fast[i] = v;
    }
    uint32_t* s_z = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (int)(R2_FAST_OFF / 4); i += blockDim.x) s_z[i] = 0;                 // (class, m) counters
    uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem + R2_BINS_OFF);
    for (int i = threadIdx.x; i < (int)R2_BIN_N; i += blockDim.x) s_bins[i] = 0;
    if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(smem + R2_SMEM) = 0u;   // block-local ticket counter
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lane4 = (uint32_t)lane * 4u;
  const __half2 C3 = r2_h2(0x35553555u);    // float16(1/3) = 1365 / 4096
  const __half2 K3 = r2_h2(0xDD55DD55u);    // -1024 * float16(1/3) = -341.25
  uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem + R2_BINS_OFF);
  // Task = 32-byte record (row, children range, flags, parent rows, first word, end word), built on the host per
  // matrix width: rows sorted heaviest first and cut into chunks -- short ones for rows with many children (their
  // children loop is latency bound and would otherwise be the tail), long ones for childless rows (per-task set-up).
  // Tasks are dealt round-robin to the blocks (static, balanced because the order is by cost); inside a block the
  // warps draw tickets from a shared-memory counter (dynamic, no global same-address atomics).  Two tickets are kept
  // in flight so that the next task's record is loaded while the current task is processed.
  uint32_t* s_ticket = reinterpret_cast<uint32_t*>(smem + R2_SMEM);
#define R2_TASK_OF(k) ((k) * gridDim.x + blockIdx.x)
  const uint32_t tasks = a.n_tasks;
  uint32_t tk = 0;
  if (lane == 0) tk = atomicAdd(s_ticket, 2u);
  tk = __shfl_sync(0xffffffffu, tk, 0);
  uint32_t task = R2_TASK_OF(tk), ntask = R2_TASK_OF(tk + 1u);
  int4 q0 = make_int4(0, 0, 0, 0), q1 = q0;
  if (task < tasks) {
    q0 = __ldg(reinterpret_cast<const int4*>(a.tasks + task));
    q1 = __ldg(reinterpret_cast<const int4*>(a.tasks + task) + 1);
  }
  while (task < tasks) {
    int4 nq0 = make_int4(0, 0, 0, 0), nq1 = nq0;
    if (ntask < tasks) {
      nq0 = __ldg(reinterpret_cast<const int4*>(a.tasks + ntask));
      nq1 = __ldg(reinterpret_cast<const int4*>(a.tasks + ntask) + 1);
    }
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(s_ticket, 1u);
    const int wbeg = q1.z, wend = q1.w;
    const int row = q0.x, c0 = q0.y, c1 = q0.y + q0.z;
    const bool has_anc = q0.w & 1, loopy = q0.w & 2;
    const int rS = q1.x, rD = q1.y;
    const bool parents_ok = has_anc && rS >= 0 && rD >= 0;
    if (!(loopy && !parents_ok)) {   // else: every word of this row belongs to k_mendel_rank_generic
      const uint32_t* prow0 = a.packed + (int64_t)row * a.wpr;
      const uint32_t* prowS = a.packed + (int64_t)(parents_ok ? rS : row) * a.wpr;
      const uint32_t* prowD = a.packed + (int64_t)(parents_ok ? rD : row) * a.wpr;
      // keep the three row pointers live: without this ptxas re-derives them (six 64-bit products) every warp step
      asm volatile("" : "+l"(prow0), "+l"(prowS), "+l"(prowD));
      int wn = wbeg + lane;
      uint32_t n0 = 0, nS = 0, nD = 0;
      if (wn < wend) { n0 = __ldg(prow0 + wn); nS = __ldg(prowS + wn); nD = __ldg(prowD + wn); }
      for (; wn - lane < wend; ) {
        const int w = wn;
        const uint32_t w0 = n0, wS = nS, wD = nD;
        wn += 32;
        if (wn < wend) { n0 = __ldg(prow0 + wn); nS = __ldg(prowS + wn); nD = __ldg(prowD + wn); }
        if (w >= wend) continue;
        uint16_t* dst = a.raw + (int64_t)row * a.ld_raw + (int64_t)w * 16;
        const uint32_t miss = w0 & (w0 >> 1) & 0x55555555u;
        const bool tail = HIST && w == a.nwords - 1 && (a.n_snps & 15);
        uint32_t outw[8];
        if (!has_anc && c0 == c1) {
          // neither ancestors nor children: every score is 1.0 (:102)
          // (words with a missing cell take the per-cell path: patching the marker in here costs 32 predicated
          // instructions on every word, missing or not)
          if (tail || (HIST && miss)) { r2_slow_word<HIST>(a, s_bins, row, w, w0, false, c0, c1); continue; }
#pragma unroll
          for (int k = 0; k < 8; ++k) outw[k] = 0x3C003C00u;
          if (HIST) {
            // one shared-memory update per warp step
            const uint32_t am = __activemask();
            if (lane == __ffs(am) - 1) atomicAdd(&s_bins[0x3C00u - R2_BIN_LO], 16u * __popc(am));
          }
          r2_store(dst, outw);
          continue;
        }
        bool slow = tail || (HIST && miss && c1 > c0);   // NaN marker + children term: per-cell path
        if (has_anc) {
          if (!parents_ok) slow = true;
          else if (((wS & (wS >> 1)) | (wD & (wD >> 1))) & 0x55555555u) {
            if (loopy) continue;   // k_mendel_rank_generic owns this word
            slow = true;
          }
        }
        uint32_t m16[8], anyv = 0;
        bool big = false;
        if (!slow && c1 > c0) {
          r2_children(a, c0, c1, w, w0, m16, anyv);
          uint32_t any = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) any |= m16[k];
          big = any & 0xFFE0FFE0u;                                                          // (class, m) counters hold m < 32
          if (any & 0xFC00FC00u) slow = true;                                               // packed-half path: m < 1024
          if (!has_anc && (miss || anyv != 0x55555555u)) slow = true;                        // 1.0 cells among child scores
        }
        if (slow) { r2_slow_word<HIST>(a, s_bins, row, w, w0, has_anc, c0, c1); continue; }
        if (has_anc) {
          // 4-row transposition of 2-bit fields: one byte obs | sire << 2 | dam << 4 per cell
          const uint32_t M2 = 0x33333333u, M4 = 0x0F0F0F0Fu;
          const uint32_t A = (w0 & M2) | ((wS << 2) & ~M2);          // nibble n: cell 2n
          const uint32_t B = ((w0 >> 2) & M2) | (wS & ~M2);          // nibble n: cell 2n + 1
          const uint32_t Cc = wD & M2, Dd = (wD >> 2) & M2;
          const uint32_t X = (A & M4) | ((Cc << 4) & ~M4);           // byte t: cell 4t
          const uint32_t Y = ((A >> 4) & M4) | (Cc & ~M4);           // byte t: cell 4t + 2
          const uint32_t Z = (B & M4) | ((Dd << 4) & ~M4);           // byte t: cell 4t + 1
          const uint32_t W = ((B >> 4) & M4) | (Dd & ~M4);           // byte t: cell 4t + 3
          if (c1 == c0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t offa = __byte_perm((k & 1) ? Y : X, lane4, 0x5504u | ((uint32_t)(k >> 1) << 4));
              const uint32_t offb = __byte_perm((k & 1) ? W : Z, lane4, 0x5504u | ((uint32_t)(k >> 1) << 4));
              const uint32_t ea = *reinterpret_cast<const uint32_t*>(smem + R2_FAST_OFF + offa);
              const uint32_t eb = *reinterpret_cast<const uint32_t*>(smem + R2_FAST_OFF + offb);
              if (HIST) {
                atomicAdd(reinterpret_cast<uint32_t*>(smem + R2_FAST_OFF + 128u + offa), 1u);
                atomicAdd(reinterpret_cast<uint32_t*>(smem + R2_FAST_OFF + 128u + offb), 1u);
              }
              outw[k] = __byte_perm(ea, eb, 0x5410u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t offa = __byte_perm((k & 1) ? Y : X, lane4, 0x5504u | ((uint32_t)(k >> 1) << 4));
              const uint32_t offb = __byte_perm((k & 1) ? W : Z, lane4, 0x5504u | ((uint32_t)(k >> 1) << 4));
              const uint32_t ea = *reinterpret_cast<const uint32_t*>(smem + R2_FAST_OFF + offa);
              const uint32_t eb = *reinterpret_cast<const uint32_t*>(smem + R2_FAST_OFF + offb);
              if (HIST && !big) {
                atomicAdd(reinterpret_cast<uint32_t*>(smem + (ea >> 16) + ((m16[k] & 0xFFFFu) << 7)), 1u);
                atomicAdd(reinterpret_cast<uint32_t*>(smem + (eb >> 16) + ((m16[k] >> 16) << 7)), 1u);
              }
              const __half2 cs = __hfma2(r2_h2(m16[k] + 0x64006400u), C3, K3);
              outw[k] = r2_u(__hadd2(cs, r2_h2(__byte_perm(ea, eb, 0x5410u))));
            }
            if (HIST && big) {
              uint32_t tmp[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) tmp[k] = outw[k];
              r2_hist_patterns(a.hist, s_bins, tmp, 0u, 16);
            }
          }
        } else {
          // founder with children, nothing missing: the score is the children term alone (class 0 = score 0)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (HIST && !big) {
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R2_CM_OFF + lane4 + ((m16[k] & 0xFFFFu) << 7)), 1u);
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R2_CM_OFF + lane4 + ((m16[k] >> 16) << 7)), 1u);
            }
            outw[k] = r2_u(__hfma2(r2_h2(m16[k] + 0x64006400u), C3, K3));
          }
          if (HIST && big) {
            uint32_t tmp[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) tmp[k] = outw[k];
            r2_hist_patterns(a.hist, s_bins, tmp, 0u, 16);
          }
        }
        r2_store(dst, outw);
      }
    }
    task = ntask; q0 = nq0; q1 = nq1;
    next = __shfl_sync(0xffffffffu, next, 0);
    ntask = R2_TASK_OF(next);
  }
  if (!HIST) return;
  // ---- retire: counters -> float16 patterns -> global histogram ------------------------------------
  __syncthreads();
  {
    const uint32_t* s_fast = reinterpret_cast<const uint32_t*>(smem + R2_FAST_OFF);
    const uint32_t* s_cm = reinterpret_cast<const uint32_t*>(smem + R2_CM_OFF);
    const int t = threadIdx.x;
    if (t < 64) {
      unsigned long long sum = 0;
      for (int l = 0; l < 32; ++l) sum += s_fast[t * 64 + 32 + ((l + t) & 31)];
      const uint32_t sc = a.lut6[t] & 0xFFFFu;
      if (sum) atomicAdd(&a.hist[sc == 0xFFFFu ? 65536u : sc], sum);
    } else if (t < 64 + 256) {
      const int cm = t - 64, cls = cm >> 5, m = cm & 31;
      unsigned long long sum = 0;
      for (int l = 0; l < 32; ++l) sum += s_cm[cm * 32 + ((l + t) & 31)];
      const uint16_t sc = a.class_sc[cls];
      // the marker class only ever has m = 0 (an unobserved cell receives no children units)
      const uint32_t pat = sc == 0xFFFFu ? 65536u : (uint32_t)gfx_half_add(gfx_child_score((uint32_t)m), sc);
      if (sum) atomicAdd(&a.hist[pat], sum);
    }
    for (int i = t; i < (int)R2_BIN_N; i += blockDim.x) {
      const uint32_t v = s_bins[i];
      if (!v) continue;
      const uint32_t bin = i == (int)R2_SLOT_ZERO ? 0u : (i == (int)R2_SLOT_MISS ? 65536u : R2_BIN_LO + i);
      atomicAdd(&a.hist[bin], (unsigned long long)v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Mendel rank kernel, third form: score classes from a boolean network, scores from PRMT byte tables, TMA-staged inputs
// ------------------------------------------------------------------------------------------------
// Same results as k_mendel_rank (bit for bit, scores and histogram; scripts/bench_rank_forms.py compares the two on a
// dozen data sets).  What changes is where the time goes (profiles/r2_ncu_lines_rank.txt: 349 warp instructions per
// 32-word step in the form above, 25 % task / loop overhead, 25 % the children loop, 13 % the one-cell table lookups and
// 8 % shared-memory atomics; 32 shared-memory operations per step):
//  * with both parents observed the ancestor score of a cell is 2 * (max - own) of one row of Mendel's table: four values
//    (0, 0.5, 1, 2).  Its 2-bit class is computed for the 16 cells of a word at once by a network of nine LOP3 over the six
//    code bit planes (rank3_class is the same network on the host; the schedule checks it against the table built by
//    inference).  PRMT then serves as the table: the class nibbles of four cells select the high bytes of their float16
//    scores from a 4-byte constant, a second PRMT interleaves them with zero low bytes -- an output word (two scores) per
//    PRMT, no shared memory.  [Two earlier versions of this kernel looked the scores up two cells at a time in a 4096-entry
//    shared-memory table: 6.6-8.7 wavefronts per LDS.64 (bank conflicts on an index that is mostly the own genotype), the
//    LSU pipe 46 % busy, 2.3 short-scoreboard stalls per issue -- profiles/r2_ncu_rank3_table.txt.]
//  * childless rows (2/3 of the cells of config 2) tally their histogram as three popcounts of class masks per word, kept
//    in registers and added to the lane-private (class, m = 0) counters once per TASK;
//  * one specialised loop per kind of row (the dispatch tests of the form above ran on every step);
//  * TMA-staged inputs, pipelined two tasks deep (below): nothing of the prefetch pipeline lives in registers;
//  * a child word address is one IMAD.WIDE (row * bytes-per-row + the lane's column pointer), child rows are broadcast
//    by shuffles from one register per lane.
#ifndef R3_THREADS
#define R3_THREADS 640   // 96 registers: no spills (a spilled pipeline variable costs a memory latency per task)
#endif
// words per task by number of children of the row (at most R3_CHUNK = 128: one staging stage)
#ifndef R3_WORDS_LIGHT
#define R3_WORDS_LIGHT 128   // childless
#endif
#ifndef R3_WORDS_MID
#define R3_WORDS_MID 128     // 1..7 children
#endif
#ifndef R3_WORDS_HEAVY
#define R3_WORDS_HEAVY 32    // 8..63 children
#endif
#ifndef R3_WORDS_HUGE
#define R3_WORDS_HUGE 32     // >= 64 children
#endif
#define R3_CM_OFF 0u          // u32 [4 classes][64 m][32 lanes]
#define R3_BINS_OFF 32768u    // u32 [8192] patterns 0x3000..0x4FFF
#define R3_SMEM 65536u

// prmt.b32 with the full selector: bit 3 of a selector nibble replicates the sign of the chosen byte
__device__ __forceinline__ uint32_t r3_prmt(uint32_t x, uint32_t y, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(sel));
  return r;
}

// x >> k on the FMA pipe (IMAD.HI): the kernel is bound by the ALU pipe (LOP3 / SHF / PRMT: one warp instruction per two
// cycles per scheduler, profiles/r2_ncu_rank3.txt: 59 % busy against 17 % for the FMA pipe)
#ifndef R3_FMA_SHIFTS
#define R3_FMA_SHIFTS 0   // measured: 80.3 us with, 78.2 us without (IMAD.HI is no cheaper than SHF here)
#endif
__device__ __forceinline__ uint32_t r3_shr(uint32_t x, int k) {
#if R3_FMA_SHIFTS
  return __umulhi(x, 1u << (32 - k));
#else
  return x >> k;
#endif
}
struct R3Kids {
  uint32_t o0, o1, o2;             // parent's cell is 0 / 1 / 2 (bit 2i)
  uint32_t bl0, bl1, bh0, bh1;     // byte t: cells 4t, 4t + 2, 4t + 1, 4t + 3
  uint32_t allmiss;
};
// four child words: units of float16(1/3) per cell (bayesnet/model.py:93-116, other parent unknown, D7)
__device__ __forceinline__ void r3_round(R3Kids& K, const uint32_t wc[4]) {
  uint32_t acc_lo = 0, acc_hi = 0;          // nibble n: cell 2n / cell 2n + 1 (4 children * 2 units = 8 < 16)
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t cl = wc[u], ch = r3_shr(wc[u], 1);
    K.allmiss &= cl & ch;
    const uint32_t one = ~cl & K.o1;                            // child 0 or 2 under a heterozygous parent: 1 unit
    const uint32_t two = ~cl & ((ch & K.o0) | (~ch & K.o2));    // child 2 under parent 0, child 0 under parent 2: 2 units
    const uint32_t contrib = two * 2u + one;                    // disjoint bits: the sum is the union (one IMAD)
    acc_lo += contrib & 0x33333333u;
    acc_hi += r3_shr(contrib, 2) & 0x33333333u;
  }
  K.bl0 += acc_lo & 0x0F0F0F0Fu; K.bl1 += r3_shr(acc_lo, 4) & 0x0F0F0F0Fu;
  K.bh0 += acc_hi & 0x0F0F0F0Fu; K.bh1 += r3_shr(acc_hi, 4) & 0x0F0F0F0Fu;
}
__device__ __forceinline__ void r3_flush(R3Kids& K, uint32_t m16[8]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    m16[2 * t] += __byte_perm(K.bl0, K.bh0, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
    m16[2 * t + 1] += __byte_perm(K.bl1, K.bh1, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
  }
  K.bl0 = K.bl1 = K.bh0 = K.bh1 = 0;
}
#define R3_LOAD_FULL(c)                                                                     \
  _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                           \
    const int r = __shfl_sync(0xffffffffu, kid, (c) + u);                                   \
    nx[u] = __ldg(reinterpret_cast<const uint32_t*>(pw + (int64_t)r * wprb));               \
  }
#define R3_LOAD_TAIL(c)                                                                     \
  _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                           \
    const int r = __shfl_sync(0xffffffffu, kid, ((c) + u) & 31);                            \
    nx[u] = 0xFFFFFFFFu;                                                                    \
    if (u < 3 && r >= 0) nx[u] = __ldg(reinterpret_cast<const uint32_t*>(pw + (int64_t)r * wprb)); \
  }
// the first round (children 0..3) of the word column pw points to: requested one step ahead of its use
__device__ __forceinline__ void r3_children_first(const char* pw, int wprb, int nk, int kid, uint32_t nx[4]) {
  if (nk >= 4) { R3_LOAD_FULL(0) } else { R3_LOAD_TAIL(0) }
}
// m16[k] = m(cell 2k) | m(cell 2k + 1) << 16 over the nk genotyped children of the row, at the word column pw points
// to; allmiss bit 2i: every child is missing at cell i.  kid: lane i holds the row of child i (i < 32, -1 beyond nk);
// nx: the words of children 0..3 (r3_children_first).  All 32 lanes must call this (shuffles).
// over: bit 0 some m >= 64 (beyond the (class, m) counters), bit 1 some m >= 1024 (beyond the packed-half path).
__device__ __forceinline__ void r3_children(const char* pw, int wprb, const int32_t* __restrict__ child_row, int c0, int nk,
                                            int kid, int lane, uint32_t w0, uint32_t nx[4], uint32_t m16[8], uint32_t& allmiss,
                                            uint32_t& over) {
  const uint32_t M = 0x55555555u;
  R3Kids K;
  K.o0 = ~(w0 | (w0 >> 1)) & M;
  K.o1 = (w0 & ~(w0 >> 1)) & M;
  K.o2 = ((w0 >> 1) & ~w0) & M;
  K.bl0 = K.bl1 = K.bh0 = K.bh1 = 0;
  K.allmiss = 0xFFFFFFFFu;
  uint32_t wc[4];
  if (nk <= 32) {
    // one batch (every row of a livestock dam, most sires): no outer loop, the byte counters (at most 64) are the result
    const int full = nk >> 2, rem = nk & 3;
    for (int i = 1; i < full; ++i) {        // a full round followed by a full round
#pragma unroll
      for (int u = 0; u < 4; ++u) wc[u] = nx[u];
      R3_LOAD_FULL(4 * i)
      r3_round(K, wc);
    }
    if (full > 0) {                           // the last full round
#pragma unroll
      for (int u = 0; u < 4; ++u) wc[u] = nx[u];
      if (rem) { R3_LOAD_TAIL(4 * full) }
      r3_round(K, wc);
    }
    if (rem) r3_round(K, nx);                 // 1..3 children (the others read as missing)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      m16[2 * t] = __byte_perm(K.bl0, K.bh0, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
      m16[2 * t + 1] = __byte_perm(K.bl1, K.bh1, 0x0400u + 0x0101u * t) & 0x00FF00FFu;
    }
    over = ((K.bl0 | K.bl1 | K.bh0 | K.bh1) & 0xC0C0C0C0u) ? 1u : 0u;
    allmiss = K.allmiss & M;
    return;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) m16[k] = 0;
  for (int base = 0; base < nk; base += 32) {
    const int cnt = min(32, nk - base);
    int kidn = -1;
    if (base + 32 < nk && lane < nk - base - 32) kidn = __ldg(child_row + c0 + base + 32 + lane);   // next batch of 32
    const int full = cnt >> 2, rem = cnt & 3;
    if (base > 0) {
      if (full > 0) { R3_LOAD_FULL(0) } else { R3_LOAD_TAIL(0) }
    }
    for (int i = 1; i < full; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) wc[u] = nx[u];
      R3_LOAD_FULL(4 * i)
      r3_round(K, wc);
    }
    if (full > 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) wc[u] = nx[u];
      if (rem) { R3_LOAD_TAIL(4 * full) }
      r3_round(K, wc);
    }
    if (rem) r3_round(K, nx);
    r3_flush(K, m16);                         // 8 rounds * 8 units = 64 per byte
    kid = kidn;
  }
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) any |= m16[k];
  over = ((any & 0xFFC0FFC0u) ? 1u : 0u) | ((any & 0xFC00FC00u) ? 2u : 0u);
  allmiss = K.allmiss & M;
}

// Score classes of the 16 cells of a word whose own, sire and dam cells are all observed: bit 2i of b0 / b1 = low / high
// bit of the class of cell i (0 -> 0, 1 -> 0.5, 2 -> 1, 3 -> 2); the odd bits are garbage.
__device__ __forceinline__ void r3_classes(uint32_t w0, uint32_t wS, uint32_t wD, uint32_t& b0, uint32_t& b1) {
  const uint32_t o0 = w0, o1 = r3_shr(w0, 1), s0 = wS, s1 = r3_shr(wS, 1), d0 = wD, d1 = r3_shr(wD, 1);
  const uint32_t u = o0 ^ s1 ^ d1;                  // both parents homozygous: own low bit differs from the expected child's
  const uint32_t v = o1 ^ (s1 & d1);                //                          own high bit differs
  const uint32_t t3 = (u | v) & ~s0;
  const uint32_t aa = s1 | d1;
  const uint32_t q = (aa & ~(o0 | o1)) | (~aa & o1);   // one heterozygous parent: own is the homozygote the other parent excludes
  const uint32_t c2 = q & (s0 ^ d0);
  const uint32_t c1 = s0 & d0 & ~o0;                // both heterozygous, own homozygous: 0.5
  b1 = (t3 & ~d0) | c2;
  b0 = (t3 & ~d0) | c1;
}
// class nibbles: V nibble n = class of cell 2n, W nibble n = class of cell 2n + 1
#define R3_NIBBLES(b0, b1)                                                              \
  const uint32_t C_ = ((b0) & 0x55555555u) | (((b1) * 2u) & 0xAAAAAAAAu);               \
  const uint32_t V = C_ & 0x33333333u, W = r3_shr(C_, 2) & 0x33333333u
// four bytes TAB[class] for the cells 0, 2, 4, 6 (V) or 1, 3, 5, 7 (W) of the low / high half of the nibble word
#define R3_BYTES(TAB, X, hi) r3_prmt(TAB, 0u, (hi) ? r3_shr(X, 16) : (X))
// word k (cells 2k, 2k + 1) = [0, byte k of E, 0, byte k of O]; every table byte has a clear sign bit, so its sign is the zero
#define R3_WORD(E, O, k) r3_prmt(E, O, 0x4C08u + 0x1111u * (uint32_t)(k))
#define R3_SCORE_TAB 0x403C3800u    // high bytes of float16 0, 0.5, 1, 2
#define R3_CLASS_TAB 0x60402000u    // class << 5: as byte 1 / 3 of a word it is the offset class * 8192 of the (class, m) counters

// One childless word with own, sire and dam cells observed: 16 scores out, class counts up.
template <bool HIST>
__device__ __forceinline__ void r3_leaf_word(uint32_t w0, uint32_t wS, uint32_t wD, uint16_t* dst, uint32_t& n1, uint32_t& n2,
                                             uint32_t& n3) {
  const uint32_t M = 0x55555555u;
  uint32_t b0, b1;
  r3_classes(w0, wS, wD, b0, b1);
  R3_NIBBLES(b0, b1);
  const uint32_t e0 = R3_BYTES(R3_SCORE_TAB, V, 0), e1 = R3_BYTES(R3_SCORE_TAB, V, 1);
  const uint32_t o0 = R3_BYTES(R3_SCORE_TAB, W, 0), o1 = R3_BYTES(R3_SCORE_TAB, W, 1);
  uint32_t outw[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { outw[k] = R3_WORD(e0, o0, k); outw[4 + k] = R3_WORD(e1, o1, k); }
  r2_store(dst, outw);
  if (HIST) {
    n1 += __popc(b0 & ~b1 & M);
    n2 += __popc(~b0 & b1 & M);
    n3 += __popc(b0 & b1 & M);
  }
}

struct R3Task {
  int row, c0, nk, fl, rS, rD, wbeg, wend;
};
__device__ __forceinline__ R3Task r3_decode(const int4& q0, const int4& q1) {
  return R3Task{q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
}

// ---- TMA staging (cp.async.bulk + mbarrier), one private area per warp -------------------------------------
// A task covers at most R3_CHUNK words of a row.  Its inputs -- the own, sire and dam row segments (<= 512 bytes each)
// and the rows of its first 32 children (128 bytes) -- are fetched by four bulk copies that one lane issues while the
// PREVIOUS task runs; the 32-byte task records travel the same way one task further ahead.  Nothing of this pipeline
// lives in registers: the first version kept the records and first words in flight in registers, ptxas spilled them
// and every spill store waited for its load (profiles/r2_ncu_rank3_regpipe.txt: 42 % of the stall samples).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

#ifndef R3_PROXY_FENCE
#define R3_PROXY_FENCE 1   // measured: 77.4 us per call with and without
#endif
#ifndef R3_XSTEP
#define R3_XSTEP 1       // the first child words of the next step are requested before the current step is scored
#endif
#ifndef R3_CHUNK
#define R3_CHUNK 128   // words per task at most
#endif
#define R3_STAGE_BYTES (3u * R3_CHUNK * 4u + 128u)     // own, sire, dam segments, child rows
#define R3_REC_OFF (2u * R3_STAGE_BYTES)               // two 32-byte record slots
#define R3_BAR_OFF (R3_REC_OFF + 64u)                  // mbarriers: data[2], record[2]
#define R3_WARP_BYTES ((R3_BAR_OFF + 32u + 127u) / 128u * 128u)
#define R3_WARP_OFF (R3_SMEM + 128u)
#define R3_DYN_SMEM (R3_WARP_OFF + (R3_THREADS / 32) * R3_WARP_BYTES)

__device__ __forceinline__ void r3_issue_record(const RankArgs& a, char* wsm, uint32_t task, int slot) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + R3_BAR_OFF) + 2 + slot;
  mbar_expect_tx(bar, 32u);
  tma_load_1d(wsm + R3_REC_OFF + 32u * slot, a.tasks + task, 32u, bar);
}
__device__ __forceinline__ void r3_issue_data(const RankArgs& a, char* wsm, const R3Task& T, int stage) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + R3_BAR_OFF) + stage;
  char* st = wsm + (uint32_t)stage * R3_STAGE_BYTES;
  const uint32_t bytes = (uint32_t)((T.wend - T.wbeg + 3) & ~3) * 4u;
  const bool pk = (T.fl & 1) && T.rS >= 0 && T.rD >= 0;
  const uint32_t kb = T.nk > 0 ? (uint32_t)((min(T.nk, 32) + 3) & ~3) * 4u : 0u;
  mbar_expect_tx(bar, bytes * (pk ? 3u : 1u) + kb);
  tma_load_1d(st, a.packed + (int64_t)T.row * a.wpr + T.wbeg, bytes, bar);
  if (pk) {
    tma_load_1d(st + R3_CHUNK * 4u, a.packed + (int64_t)T.rS * a.wpr + T.wbeg, bytes, bar);
    tma_load_1d(st + 2u * R3_CHUNK * 4u, a.packed + (int64_t)T.rD * a.wpr + T.wbeg, bytes, bar);
  }
  if (kb) tma_load_1d(st + 3u * R3_CHUNK * 4u, a.child_row + T.c0, kb, bar);
}

// a word of a loopy row in which a parent cell is missing belongs to the general inference: the active lanes of the warp
// append theirs to the work list with one atomic
__device__ __forceinline__ void r3_push(const RankArgs& a, int row, uint32_t w) {
  if (!a.wl_hdr) return;   // scanning form of the consumer
  const uint32_t act = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(act) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(a.wl_hdr, (uint32_t)__popc(act));
  base = __shfl_sync(act, base, leader);
  const uint32_t at = base + (uint32_t)__popc(act & ((1u << lane) - 1u));
  if (at < a.wl_cap) a.wl[at] = make_uint2((uint32_t)row, w);
}

template <bool HIST>
__global__ void __launch_bounds__(R3_THREADS, 1) k_mendel_rank3(RankArgs a) {
  extern __shared__ __align__(128) char smem[];
  // k_mendel_rank_generic (loopy rows, disjoint cells) may start as soon as SMs free up: it waits for this grid at its end
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  char* const wsm = smem + R3_WARP_OFF + (threadIdx.x >> 5) * R3_WARP_BYTES;
  {
    if (HIST) {
      uint4* z = reinterpret_cast<uint4*>(smem + R3_CM_OFF);
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
      z = reinterpret_cast<uint4*>(smem + R3_BINS_OFF);
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(smem + R3_SMEM) = 0u;   // block-local ticket counter
    if (lane == 0) {
      uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + R3_BAR_OFF);
      for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const uint32_t lane4 = (uint32_t)lane * 4u;
  const uint32_t lane44 = lane4 * 0x00010001u;
  const uint32_t M = 0x55555555u;
  const __half2 C3 = r2_h2(0x35553555u);    // float16(1/3) = 1365 / 4096
  const __half2 K3 = r2_h2(0xDD55DD55u);    // -1024 * float16(1/3) = -341.25
  uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem + R3_BINS_OFF);
  uint32_t utailw = (HIST && (a.n_snps & 15)) ? (uint32_t)(a.nwords - 1) : 0xFFFFFFFFu;   // ragged last word: per-cell path
  asm volatile("" : "+r"(utailw));
  const int wprb = (int)(a.wpr * 4);
  // Tasks: records dealt round-robin to the blocks (static, balanced because the table is ordered by cost), tickets from a
  // shared-memory counter inside the block (as in k_mendel_rank).  [One global counter instead: same-address atomics
  // retire at ~2 ns each, 54 000 draws took longer than the kernel; draws of 4 or 8 consecutive tasks left a tail of one
  // draw, 25-50 % of a warp's life -- 112 / 144 / 215 us per chunk against 87, profiles/r2_rank3_variants.md.]
  // Task number i of a warp uses data stage i & 1 and record slot i & 1; the phase parity of both is (i >> 1) & 1.
  uint32_t* s_ticket = reinterpret_cast<uint32_t*>(smem + R3_SMEM);
  const uint32_t tasks = a.n_tasks;
  uint32_t tk = 0;
  if (lane == 0) tk = atomicAdd(s_ticket, 3u);
  tk = __shfl_sync(0xffffffffu, tk, 0);
  uint32_t t0 = R2_TASK_OF(tk), t1 = R2_TASK_OF(tk + 1u), t2 = R2_TASK_OF(tk + 2u);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(wsm + R3_BAR_OFF);
  R3Task T = R3Task{0, 0, 0, 0, 0, 0, 0, 0};
  if (t0 < tasks) {
    if (lane == 0) {
      r3_issue_record(a, wsm, t0, 0);
      if (t1 < tasks) r3_issue_record(a, wsm, t1, 1);
    }
    mbar_wait(bars + 2, 0u);
    T = r3_decode(*reinterpret_cast<const int4*>(wsm + R3_REC_OFF), *reinterpret_cast<const int4*>(wsm + R3_REC_OFF + 16u));
    __syncwarp();
    if (lane == 0) r3_issue_data(a, wsm, T, 0);
  }
  for (uint32_t it = 0; t0 < tasks; ++it) {
    // ---- pipeline: the next task's inputs, the record of the one after, the next ticket -------------------
    if (t1 < tasks) {
      const uint32_t sl = (it + 1u) & 1u;
      mbar_wait(bars + 2 + sl, ((it + 1u) >> 1) & 1u);
      if (lane == 0) {
        // the record stays in its slot until this task is over (it is read again when it becomes the current one)
        const R3Task T1 = r3_decode(*reinterpret_cast<const int4*>(wsm + R3_REC_OFF + 32u * sl), *reinterpret_cast<const int4*>(wsm + R3_REC_OFF + 32u * sl + 16u));
        // the stage refilled here was read two iterations ago and the shuffle that closes every iteration has all lanes past
        // those reads; the proxy fence (generic reads before async writes of the same shared memory) costs nothing measurable
#if R3_PROXY_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        r3_issue_data(a, wsm, T1, (int)sl);
      }
    }
    uint32_t next = 0;
    if (lane == 0) {
      if (t2 < tasks) r3_issue_record(a, wsm, t2, (int)(it & 1u));
      next = atomicAdd(s_ticket, 1u);
    }
    // ---- the task ---------------------------------------------------------------------------------------
    const int wbeg = T.wbeg, row = T.row, c0 = T.c0, nk = T.nk;
    const uint32_t uwend = (uint32_t)T.wend;
    const bool has_anc = T.fl & 1, loopy = T.fl & 2;
    const bool parents_ok = has_anc && T.rS >= 0 && T.rD >= 0;
    uint16_t* rawrow = a.raw + (int64_t)row * a.ld_raw;
    asm volatile("" : "+l"(rawrow));   // pinned: ptxas otherwise re-derives the 64-bit product on every step
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(wsm + (it & 1u) * R3_STAGE_BYTES) + lane;   // own; + R3_CHUNK sire; + 2 R3_CHUNK dam
    mbar_wait(bars + (it & 1u), (it >> 1) & 1u);
    uint32_t w = (uint32_t)(wbeg + lane);
    if (loopy && !parents_ok) {
      // every word of this row belongs to k_mendel_rank_generic
      for (; w < uwend; w += 32u) r3_push(a, row, w);
    } else if (!has_anc && nk == 0) {
      // ---- neither ancestors nor children: every score is 1.0 (:102) ---------------------------------
      uint32_t ones = 0;
      for (; w < uwend; w += 32u, sw += 32) {
        if (HIST) {
          const uint32_t w0 = sw[0];
          if (w == utailw || (w0 & (w0 >> 1) & M)) { r2_slow_word<HIST>(a, s_bins, row, w, w0, false, c0, c0); continue; }
          ones += 16u;
        }
        const uint4 v = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
        __stcs(reinterpret_cast<uint4*>(rawrow + (uint64_t)w * 16u), v);
        __stcs(reinterpret_cast<uint4*>(rawrow + (uint64_t)w * 16u) + 1, v);
      }
      if (HIST) {
        ones = __reduce_add_sync(0xffffffffu, ones);
        if (lane == 0 && ones) atomicAdd(&s_bins[0x3C00u - R2_BIN_LO], ones);
      }
    } else if (has_anc && !parents_ok) {
      // ---- a parent without a genotype: the grand-parents matter, per-cell path with the full table --
      for (; w < uwend; w += 32u, sw += 32) r2_slow_word<HIST>(a, s_bins, row, w, sw[0], true, c0, c0 + nk);
    } else if (has_anc && nk == 0) {
      // ---- childless, both parents genotyped ----------------------------------------------------------
      uint32_t n1 = 0, n2 = 0, n3 = 0, nw = 0;   // cells of class 1, 2, 3; words
      for (; w < uwend; w += 32u, sw += 32) {
        const uint32_t w0 = sw[0], wS = sw[R3_CHUNK], wD = sw[2 * R3_CHUNK];
        const uint32_t pm = ((wS & (wS >> 1)) | (wD & (wD >> 1))) & M;
        if (pm != 0u || w == utailw || (w0 & (w0 >> 1) & M)) {   // a missing own cell: marker (or 1.0) on the per-cell path
          if (pm != 0u && loopy) r3_push(a, row, w);
          else r2_slow_word<HIST>(a, s_bins, row, w, w0, true, c0, c0);
          continue;
        }
        r3_leaf_word<HIST>(w0, wS, wD, rawrow + (uint64_t)w * 16u, n1, n2, n3);
        nw += 16u;
      }
      if (HIST) {
        const uint32_t n0 = nw - n1 - n2 - n3;
        if (n0) atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + lane4), n0);
        if (n1) atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + 8192u + lane4), n1);
        if (n2) atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + 16384u + lane4), n2);
        if (n3) atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + 24576u + lane4), n3);
      }
    } else {
      // ---- rows with children: both parents genotyped, or a founder (score = the children term alone, class 0) ----
      const int kid = lane < nk ? (int)sw[3 * R3_CHUNK] : -1;
      uint32_t nx[4];
#if R3_XSTEP
      r3_children_first(reinterpret_cast<const char*>(a.packed + min(w, uwend - 1u)), wprb, nk, kid, nx);
#endif
      for (uint32_t wb = (uint32_t)wbeg; wb < uwend; wb += 32u, w += 32u, sw += 32) {
        const uint32_t w0 = sw[0];
        uint32_t m16[8], allmiss, over;
#if !R3_XSTEP
        r3_children_first(reinterpret_cast<const char*>(a.packed + min(w, uwend - 1u)), wprb, nk, kid, nx);
#endif
        r3_children(reinterpret_cast<const char*>(a.packed + min(w, uwend - 1u)), wprb, a.child_row, c0, nk, kid, lane, w0, nx, m16, allmiss, over);
        // the next step's first child words travel while this step is scored
#if R3_XSTEP
        if (wb + 32u < uwend) r3_children_first(reinterpret_cast<const char*>(a.packed + min(w + 32u, uwend - 1u)), wprb, nk, kid, nx);
#endif
        if (w >= uwend) continue;
        const bool big = over & 1u;                            // (class, m) counters hold m < 64
        const uint32_t miss = w0 & (w0 >> 1) & M;
        uint32_t outw[8];
        if (has_anc) {
          const uint32_t wS = sw[R3_CHUNK], wD = sw[2 * R3_CHUNK];
          const uint32_t pm = ((wS & (wS >> 1)) | (wD & (wD >> 1))) & M;
          if (pm != 0u || w == utailw || miss || (over & 2u)) {   // packed-half path: m < 1024; own observed
            if (pm != 0u && loopy) r3_push(a, row, w);
            else r2_slow_word<HIST>(a, s_bins, row, w, w0, true, c0, c0 + nk);
            continue;
          }
          uint32_t b0, b1;
          r3_classes(w0, wS, wD, b0, b1);
          R3_NIBBLES(b0, b1);
          const uint32_t e0 = R3_BYTES(R3_SCORE_TAB, V, 0), e1 = R3_BYTES(R3_SCORE_TAB, V, 1);
          const uint32_t o0 = R3_BYTES(R3_SCORE_TAB, W, 0), o1 = R3_BYTES(R3_SCORE_TAB, W, 1);
          const uint32_t ce0 = R3_BYTES(R3_CLASS_TAB, V, 0), ce1 = R3_BYTES(R3_CLASS_TAB, V, 1);
          const uint32_t co0 = R3_BYTES(R3_CLASS_TAB, W, 0), co1 = R3_BYTES(R3_CLASS_TAB, W, 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t anc = (k < 4) ? R3_WORD(e0, o0, k & 3) : R3_WORD(e1, o1, k & 3);
            if (HIST && !big) {
              const uint32_t cw = (k < 4) ? R3_WORD(ce0, co0, k & 3) : R3_WORD(ce1, co1, k & 3);
              const uint32_t t = cw + (m16[k] << 7) + lane44;
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + (t & 0xFFFFu)), 1u);
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + (t >> 16)), 1u);
            }
            const __half2 cs = __hfma2(r2_h2(m16[k] + 0x64006400u), C3, K3);
            outw[k] = r2_u(__hadd2(cs, r2_h2(anc)));
          }
        } else {
          // a cell that is missing or has no genotyped child scores 1.0, not the children term: per-cell path
          if (w == utailw || miss || allmiss != 0u || (over & 2u)) {
            r2_slow_word<HIST>(a, s_bins, row, w, w0, false, c0, c0 + nk);
            continue;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (HIST && !big) {
              const uint32_t t = (m16[k] << 7) + lane44;
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + (t & 0xFFFFu)), 1u);
              atomicAdd(reinterpret_cast<uint32_t*>(smem + R3_CM_OFF + (t >> 16)), 1u);
            }
            outw[k] = r2_u(__hfma2(r2_h2(m16[k] + 0x64006400u), C3, K3));
          }
        }
        if (HIST && big) {
          uint32_t tmp[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) tmp[k] = outw[k];
          r2_hist_patterns(a.hist, s_bins, tmp, 0u, 16);
        }
        r2_store(rawrow + (uint64_t)w * 16u, outw);
      }
    }
    // ---- rotate the pipeline ------------------------------------------------------------------------------
    {
      const uint32_t sl = (it + 1u) & 1u;
      T = r3_decode(*reinterpret_cast<const int4*>(wsm + R3_REC_OFF + 32u * sl), *reinterpret_cast<const int4*>(wsm + R3_REC_OFF + 32u * sl + 16u));
    }
    t0 = t1; t1 = t2;
    next = __shfl_sync(0xffffffffu, next, 0);
    t2 = R2_TASK_OF(next);
  }
  if (!HIST) return;
  // ---- retire: counters -> float16 patterns -> global histogram ------------------------------------
  __syncthreads();
  {
    const uint32_t* s_cm = reinterpret_cast<const uint32_t*>(smem + R3_CM_OFF);
    const int t = threadIdx.x;
    if (t < 256) {
      const int cls = t >> 6, m = t & 63;
      unsigned long long sum = 0;
      for (int l = 0; l < 32; ++l) sum += s_cm[t * 32 + ((l + t) & 31)];
      const uint16_t sc = (uint16_t)((R3_SCORE_TAB >> (8 * cls)) & 0xFFu) << 8;
      if (sum) atomicAdd(&a.hist[gfx_half_add(gfx_child_score((uint32_t)m), sc)], sum);
    }
    for (int i = t; i < (int)R2_BIN_N; i += blockDim.x) {
      const uint32_t v = s_bins[i];
      if (!v) continue;
      const uint32_t bin = i == (int)R2_SLOT_ZERO ? 0u : (i == (int)R2_SLOT_MISS ? 65536u : R2_BIN_LO + i);
      atomicAdd(&a.hist[bin], (unsigned long long)v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic (loopy blanket) rank, pair/single inference, apply
// ------------------------------------------------------------------------------------------------
struct CellCtx {
  const uint32_t* live;   // packed matrix being read (the generation-start snapshot in the correct kernels)
  int64_t wpr;
  const uint16_t* raw;    // float16 scores; 0xFFFF marks a cell that was missing in the input (E = 1, :603)
  int64_t ld_raw;
  const double* elut;     // E per raw pattern (only the rank gate needs the values themselves)
  const uint16_t* cutraw; // per pmax pattern (65536 = "E is 1"): first raw pattern whose E > pmax - suspect_t
  uint32_t tp_raw;        // first raw pattern whose E >= test_p
};

__device__ __forceinline__ int cell_code(const uint32_t* p, int64_t wpr, int row, int64_t j) {
  return (int)((__ldg(p + (int64_t)row * wpr + (j >> 4)) >> (2 * (j & 15))) & 3u);
}
// raw pattern of a cell, 0xFFFF if it was missing in the input; patterns of non-negative halves order like values
__device__ __forceinline__ uint32_t cell_rx(const CellCtx& c, int row, int64_t j) {
  return __ldg(c.raw + (int64_t)row * c.ld_raw + j);
}
__device__ __forceinline__ double rx_E(const CellCtx& c, uint32_t rx) { return rx == 0xFFFFu ? 1.0 : __ldg(c.elut + rx); }

struct BlanketDev {
  const int32_t* off;
  const int32_t* row;
  const int8_t *p1, *p2, *last, *q0, *q1;
  double* big_pool;
  int* big_flags;
  double* huge_pool;
  unsigned long long* counters;
};
// diagnostic counters: 0 queries, 1 answered from the two parents, 2 eliminations, 3 big-table reruns,
// 4 nodes touched, 5 suspect cells, 6 job evaluations, 7 children visited
#ifdef GFX_DEBUG_COUNTERS
#define GFX_COUNT_PTR(p, k) do { if (p) atomicAdd((p) + (k), 1ull); } while (0)
#define GFX_COUNT(bl, k, v) do { if ((bl).counters) atomicAdd((bl).counters + (k), (unsigned long long)(v)); } while (0)
#else
#define GFX_COUNT_PTR(p, k) do { } while (0)
#define GFX_COUNT(bl, k, v) do { } while (0)   // each site costs ~40 SASS instructions of warp-aggregated atomic: debug builds only
#endif

// Evidence is gathered lazily, walking outwards from the query only through unobserved nodes: the
// usual case (both parents observed and not suspect) touches two rows instead of the whole blanket.
struct LazyBlanket {
  const CellCtx& c;
  const int32_t* rows;
  const int8_t* p1;
  const int8_t* p2;
  const uint64_t* kid;
  int n;
  int64_t j;
  uint32_t cut;           // raw pattern from which a node is "suspect" and dropped from the evidence
  bool filter;
  unsigned long long* dbg = nullptr;            // false: plain evidence (rank), true: drop suspect nodes (correct)
#ifdef GFX_DEBUG_COUNTERS
  uint32_t* xflags = nullptr;   // events of one xpeel query: 1 observed extra child, 2 its other parent unobserved, 4 DOWN frame, 8 depth > 3
#endif
  // statuses of all nodes fetched up front (prefetch()): 2 bits per node.  The walks below are chains of dependent
  // steps; with the lazy form every step waited for its own two loads (the deferred-job kernels ran at 20 % issue
  // utilisation on long-scoreboard stalls), with the prefetch all loads of a job are in flight at once.
  uint64_t pre_lo = 0, pre_hi = 0;
  bool pre = false;
  __device__ __forceinline__ int status(int t) const {
    if (pre) return (int)(((t < 32 ? pre_lo >> (2 * t) : pre_hi >> (2 * (t - 32)))) & 3ull);
    const int r = __ldg(rows + t);
    if (r < 0) return 3;
    const int cd = cell_code(c.live, c.wpr, r, j);
    if (cd == 3) return 3;
    if (filter && cell_rx(c, r, j) >= cut) return 3;   // suspect node: not evidence (:230-238)
    return cd;
  }
  __device__ __forceinline__ void prefetch() {
    uint64_t lo = 0, hi = 0;
    const int64_t wo = j >> 4;
    const int sh = 2 * (int)(j & 15);
#pragma unroll 4
    for (int t = 0; t < n; ++t) {
      const int r = __ldg(rows + t);
      uint32_t st = 3u;
      if (r >= 0) {
        const uint32_t cd = (__ldg(c.live + (int64_t)r * c.wpr + wo) >> sh) & 3u;
        const uint32_t rx = filter ? (uint32_t)__ldg(c.raw + (int64_t)r * c.ld_raw + j) : 0u;
        st = (cd == 3u || (filter && rx >= cut)) ? 3u : cd;
      }
      if (t < 32) lo |= (uint64_t)st << (2 * t);
      else hi |= (uint64_t)st << (2 * (t - 32));
    }
    pre_lo = lo; pre_hi = hi; pre = true;
  }
};

// Plain peeling for tree-shaped ancestry: belief of node t from its ancestors, evidence fetched on the
// way.  Returns false as soon as an unobserved node has more than one child inside the blanket (a loop
// or evidence below it): the caller then runs the general elimination.  T(x|s,d) factorises over the
// transmitted alleles, so the contraction needs the two "transmits 0" / "transmits 1" masses only.
// (explicit stack instead of recursion: depth <= 5 levels of unobserved ancestors, else the general path)
__device__ __noinline__ bool peel(const LazyBlanket& L, int root, int qother, double* out) {
  int8_t node[6], stage[6];
  double sv[6][3];
  double r0 = 0, r1 = 0, r2 = 0;
  int sp = 0;
  node[0] = (int8_t)root;
  stage[0] = 0;
  while (sp >= 0) {
    const int t = node[sp];
    if (stage[sp] == 0) {
      const int st = t == qother ? 3 : L.status(t);
      if (st != 3) { r0 = st == 0 ? 1.0 : 0.0; r1 = st == 1 ? 1.0 : 0.0; r2 = st == 2 ? 1.0 : 0.0; --sp; continue; }
      if (__popcll(L.kid[t]) != 1) return false;   // an unobserved node shared by two branches couples them
      const int a1 = L.p1[t];
      if (a1 < 0) { r0 = 0.25; r1 = 0.5; r2 = 0.25; --sp; continue; }
      if (sp == 5) return false;
      stage[sp] = 1;
      node[sp + 1] = (int8_t)a1;
      stage[sp + 1] = 0;
      ++sp;
    } else if (stage[sp] == 1) {
      sv[sp][0] = r0; sv[sp][1] = r1; sv[sp][2] = r2;
      stage[sp] = 2;
      node[sp + 1] = L.p2[t];
      stage[sp + 1] = 0;
      ++sp;
    } else {
      const double sa = sv[sp][0] + 0.5 * sv[sp][1], sb = 0.5 * sv[sp][1] + sv[sp][2];
      const double da = r0 + 0.5 * r1, db = 0.5 * r1 + r2;
      r0 = sa * da; r2 = sb * db; r1 = sa * db + sb * da;
      --sp;
    }
  }
  out[0] = r0; out[1] = r1; out[2] = r2;
  return true;
}
// Plain peeling on the precomputed ancestor slots of a focal animal (3 levels, level order).  The evidence
// of the first two levels (6 rows) is fetched in one batch of independent loads, the third level only
// under unobserved grand-parents.  A node passes allele 0 with probability al = 1 - c/2 when observed with
// state c, 1/2 when it is an unobserved founder of the blanket, and the mean of its parents' al otherwise
// (all exact dyadics), so P(q) = [al_S al_D, al_S be_D + be_S al_D, be_S be_D].  Returns false when an
// unobserved node on the way has several children in the blanket (bit in `bad`): not a tree there.
__device__ __forceinline__ int slot_status(const CellCtx& c, int r, int64_t wo, int sh, int64_t j, uint32_t cut) {
  if (r < 0) return r == -3 ? 4 : 3;
  const int cd = (int)((__ldg(c.live + (int64_t)r * c.wpr + wo) >> sh) & 3u);
  const uint32_t rx = __ldg(c.raw + (int64_t)r * c.ld_raw + j);
  return (cd == 3 || rx >= cut) ? 3 : cd;
}
__device__ __noinline__ bool slot_peel(const CellCtx& c, const int32_t* __restrict__ srow, uint32_t bad, int64_t j,
                                          uint32_t cut, double out[3]) {
  if (bad & 0x8000u) return false;
  const int64_t wo = j >> 4;
  const int sh = 2 * (int)(j & 15);
  int st[6];
  st[0] = slot_status(c, __ldg(srow + 0), wo, sh, j, cut);
  st[1] = slot_status(c, __ldg(srow + 1), wo, sh, j, cut);
  // second level only under an unobserved parent (both parents' grand-parents in one batch)
  const bool need2 = st[0] > 2 || st[1] > 2;
#pragma unroll
  for (int s = 2; s < 6; ++s) st[s] = need2 ? slot_status(c, __ldg(srow + s), wo, sh, j, cut) : 4;
  double al[2];
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (st[f] <= 2) { al[f] = 1.0 - 0.5 * st[f]; continue; }
    if ((bad >> f) & 1u) return false;
    double a2[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int s2 = 2 * f + 2 + g;
      const int t2 = st[s2];
      if (t2 <= 2) { a2[g] = 1.0 - 0.5 * t2; continue; }
      if (t2 == 4) { a2[g] = -1.0; continue; }          // no parents: f is a founder of the blanket
      if ((bad >> s2) & 1u) return false;
      const int u0 = slot_status(c, __ldg(srow + 2 * s2 + 2), wo, sh, j, cut);
      const int u1 = slot_status(c, __ldg(srow + 2 * s2 + 3), wo, sh, j, cut);
      if (u0 == 4) { a2[g] = 0.5; continue; }           // unobserved founder
      if ((u0 == 3 && ((bad >> (2 * s2 + 2)) & 1u)) || (u1 == 3 && ((bad >> (2 * s2 + 3)) & 1u))) return false;
      const double x0 = u0 <= 2 ? 1.0 - 0.5 * u0 : 0.5, x1 = u1 <= 2 ? 1.0 - 0.5 * u1 : 0.5;
      a2[g] = 0.5 * (x0 + x1);
    }
    al[f] = a2[0] < 0.0 ? 0.5 : 0.5 * (a2[0] + a2[1]);
  }
  const double be0 = 1.0 - al[0], be1 = 1.0 - al[1];
  out[0] = al[0] * al[1];
  out[2] = be0 * be1;
  out[1] = al[0] * be1 + be0 * al[1];
  return true;
}

// Extended peeling: the same upward recursion, but an unobserved node may have further children inside
// the blanket as long as they are observed (their Mendel factor becomes a likelihood on the node, with
// the other parent's belief obtained the same way) or are childless and unobserved (barren).  Exact as
// long as no unobserved node is reached twice; otherwise `ok` is cleared and the caller falls back to
// the general elimination.  Beliefs are unnormalised: their sum is P(evidence in the component).
#define GFX_XPEEL_DEPTH 16
// Iterative form (explicit frames, no device recursion) of sum-product message passing on the part of the
// blanket that is a tree around the query.  Two kinds of frames:
//   UP(t, from):  belief of the unobserved node t from its ancestors, times the messages of its other
//                 children (everything except the child `from` we came through);
//   DOWN(k):      product of the messages of all children of the unobserved node k, plus a flag telling
//                 whether any evidence sits below k (if not, k is barren: its factor sums to 1 and the
//                 whole branch -- including the child's other parent -- is dropped, as pgmpy does).
// A child c of node v with other parent o sends  m(x_v) = sum_cx P(cx | x_v, o) D_c(cx)  where D_c is the
// indicator of its state when observed and DOWN(c) otherwise; P(cx | x, o) only needs o's allele masses.
// Returns false when the general path is needed (a node reached twice, or too deep).
__device__ __noinline__ bool belief_up(const LazyBlanket& L, int q, int qother, uint64_t& visited, int root, int root_from,
                                       double* out) {
  int8_t f_node[GFX_XPEEL_DEPTH], f_from[GFX_XPEEL_DEPTH], f_stage[GFX_XPEEL_DEPTH], f_sk[GFX_XPEEL_DEPTH],
      f_kid[GFX_XPEEL_DEPTH], f_flag[GFX_XPEEL_DEPTH];
  uint64_t f_m[GFX_XPEEL_DEPTH];
  double f_b[GFX_XPEEL_DEPTH][3], f_s[GFX_XPEEL_DEPTH][3];
  double r0 = 0, r1 = 0, r2 = 0;
  bool rflag = false;
  int sp = 0;
  f_node[0] = (int8_t)root; f_from[0] = (int8_t)root_from; f_stage[0] = 0;
  // stages: 0 enter UP, 1 sire belief back, 2 dam belief back, 3 child loop, 4 other parent of an observed child
  // back, 5 DOWN(child) back, 6 other parent of an unobserved child back, 7 enter DOWN
  while (sp >= 0) {
    const int t = f_node[sp];
    const int stage = f_stage[sp];
    if (stage == 0) {
      const int st = (t == q || t == qother) ? 3 : L.status(t);
      if (st != 3) { r0 = st == 0 ? 1.0 : 0.0; r1 = st == 1 ? 1.0 : 0.0; r2 = st == 2 ? 1.0 : 0.0; --sp; continue; }
#ifdef GFX_DEBUG_COUNTERS
      if (L.xflags && sp > 0 && (f_stage[sp - 1] == 4 || f_stage[sp - 1] == 6)) *L.xflags |= 2u;
      if (L.xflags && sp > 3) *L.xflags |= 8u;
#endif
      if ((visited >> t) & 1) { GFX_COUNT_PTR(L.dbg, 9); return false; }
      visited |= 1ull << t;
      const int a1 = L.p1[t];
      if (a1 < 0) {
        f_b[sp][0] = 0.25; f_b[sp][1] = 0.5; f_b[sp][2] = 0.25;
        f_m[sp] = L.kid[t] & ~(1ull << f_from[sp]);
        f_flag[sp] = 0;
        f_stage[sp] = 3;
      } else {
        if (sp + 1 >= GFX_XPEEL_DEPTH) return false;
        f_stage[sp] = 1;
        f_node[sp + 1] = (int8_t)a1; f_from[sp + 1] = (int8_t)t; f_stage[sp + 1] = 0;
        ++sp;
      }
    } else if (stage == 1) {
      f_s[sp][0] = r0; f_s[sp][1] = r1; f_s[sp][2] = r2;
      f_stage[sp] = 2;
      f_node[sp + 1] = L.p2[t]; f_from[sp + 1] = (int8_t)t; f_stage[sp + 1] = 0;
      ++sp;
    } else if (stage == 2) {
      const double sa = f_s[sp][0] + 0.5 * f_s[sp][1], sb = 0.5 * f_s[sp][1] + f_s[sp][2];
      const double da = r0 + 0.5 * r1, db = 0.5 * r1 + r2;
      f_b[sp][0] = sa * da; f_b[sp][2] = sb * db; f_b[sp][1] = sa * db + sb * da;
      f_m[sp] = L.kid[t] & ~(1ull << f_from[sp]);
      f_flag[sp] = 0;
      f_stage[sp] = 3;
    } else if (stage == 7) {
#ifdef GFX_DEBUG_COUNTERS
      if (L.xflags) *L.xflags |= 4u;
#endif
      if ((visited >> t) & 1) { GFX_COUNT_PTR(L.dbg, 9); return false; }
      visited |= 1ull << t;
      f_b[sp][0] = 1.0; f_b[sp][1] = 1.0; f_b[sp][2] = 1.0;
      f_m[sp] = L.kid[t];
      f_flag[sp] = 0;
      f_stage[sp] = 3;
    } else {
      if (stage == 4 || stage == 6) {
        // message of the child whose other parent's belief just came back: o passes allele 0 / 1 with oa / ob
        const double oa = r0 + 0.5 * r1, ob = 0.5 * r1 + r2;
        double d0, d1, d2;   // D_c: indicator of the observed state, or the saved DOWN(c)
        if (stage == 4) { const int sk = f_sk[sp]; d0 = sk == 0 ? 1.0 : 0.0; d1 = sk == 1 ? 1.0 : 0.0; d2 = sk == 2 ? 1.0 : 0.0; }
        else { d0 = f_s[sp][0]; d1 = f_s[sp][1]; d2 = f_s[sp][2]; }
        // x passes allele 0 with xa = 1, 1/2, 0:  P(c=0|x,o) = xa oa, P(c=2|x,o) = xb ob, P(c=1|x,o) = xa ob + xb oa
        const double m0 = oa * d0 + ob * d1;                                   // x = 0
        const double m1 = 0.5 * oa * d0 + (0.5 * ob + 0.5 * oa) * d1 + 0.5 * ob * d2;  // x = 1
        const double m2 = oa * d1 + ob * d2;                                   // x = 2
        f_b[sp][0] *= m0; f_b[sp][1] *= m1; f_b[sp][2] *= m2;
        f_stage[sp] = 3;
      } else if (stage == 5) {
        if (rflag) {
          // evidence below the child: keep its DOWN vector, fetch the other parent's belief
          if (sp + 1 >= GFX_XPEEL_DEPTH) return false;
          f_s[sp][0] = r0; f_s[sp][1] = r1; f_s[sp][2] = r2;
          f_flag[sp] = 1;
          const int k = f_kid[sp];
          f_stage[sp] = 6;
          f_node[sp + 1] = L.p1[k] == t ? L.p2[k] : L.p1[k];
          f_from[sp + 1] = (int8_t)k;
          f_stage[sp + 1] = 0;
          ++sp;
          continue;
        }
        f_stage[sp] = 3;   // barren branch: dropped
      }
      bool pushed = false;
      uint64_t m = f_m[sp];
      while (m) {
        const int k = gfx_lowbit(m);
        m &= m - 1;
        const int sk = (k == q || k == qother) ? 3 : L.status(k);
        if (sk == 3 && L.kid[k] == 0) continue;   // childless and unobserved: barren, sums to 1
        if (sp + 1 >= GFX_XPEEL_DEPTH) return false;
        f_m[sp] = m;
        f_kid[sp] = (int8_t)k;
        if (sk != 3) {
#ifdef GFX_DEBUG_COUNTERS
          if (L.xflags) *L.xflags |= 1u;
#endif
          f_sk[sp] = (int8_t)sk;
          f_flag[sp] = 1;
          f_stage[sp] = 4;
          f_node[sp + 1] = L.p1[k] == t ? L.p2[k] : L.p1[k];
          f_from[sp + 1] = (int8_t)k;
          f_stage[sp + 1] = 0;
        } else {
          f_stage[sp] = 5;
          f_node[sp + 1] = (int8_t)k;
          f_from[sp + 1] = (int8_t)t;
          f_stage[sp + 1] = 7;
        }
        ++sp;
        pushed = true;
        break;
      }
      if (!pushed) { r0 = f_b[sp][0]; r1 = f_b[sp][1]; r2 = f_b[sp][2]; rflag = f_flag[sp] != 0; --sp; }
    }
  }
  out[0] = r0; out[1] = r1; out[2] = r2;
  return true;
}
// out = normalised marginal, *total = unnormalised sum (0 => the reference's joint is 0/0)
__device__ __noinline__ bool xpeel_query(const LazyBlanket& L, int q, int qother, double out[3], double* total) {
  if (L.kid[q] != 0 || L.p1[q] < 0) { GFX_COUNT_PTR(L.dbg, 11); return false; }
  uint64_t visited = 1ull << q;
  double s[3], d[3];
#ifdef GFX_DEBUG_COUNTERS
  uint32_t xf = 0;
  LazyBlanket& LM = const_cast<LazyBlanket&>(L);
  LM.xflags = &xf;
  const bool ok_s = belief_up(L, q, qother, visited, L.p1[q], q, s);
  const bool ok_d = ok_s && belief_up(L, q, qother, visited, L.p2[q], q, d);
  LM.xflags = nullptr;
  if (L.dbg) {
    // 12: pure tree walk (slot peeling gave up for depth / barren children)  13: observed extra children with observed other parents
    // 14: needs recursion (other parent unobserved or DOWN frames)           15: the same classes among the FAILED walks (<< 2 bits)
    const int cls = (xf & 6u) ? 14 : ((xf & 1u) ? 13 : 12);
    if (ok_d) atomicAdd(L.dbg + cls, 1ull);
    else atomicAdd(L.dbg + 15, 1ull << (20 * (cls - 12)));
  }
  if (!ok_d) return false;
#else
  if (!belief_up(L, q, qother, visited, L.p1[q], q, s)) return false;
  if (!belief_up(L, q, qother, visited, L.p2[q], q, d)) return false;
#endif
  const double sa = s[0] + 0.5 * s[1], sb = 0.5 * s[1] + s[2], da = d[0] + 0.5 * d[1], db = 0.5 * d[1] + d[2];
  const double u0 = sa * da, u2 = sb * db, u1 = sa * db + sb * da;
  const double t = (u0 + u1) + u2;
  out[0] = u0 / t; out[1] = u1 / t; out[2] = u2 / t;
  *total = t;
  return true;
}

__device__ __forceinline__ bool peel_query(const LazyBlanket& L, int q, int qother, double out[3]) {
  if (L.kid[q] != 0 || L.p1[q] < 0) return false;
  double s[3], d[3];
  if (!peel(L, L.p1[q], qother, s)) return false;
  if (!peel(L, L.p2[q], qother, d)) return false;
  const double sa = s[0] + 0.5 * s[1], sb = 0.5 * s[1] + s[2], da = d[0] + 0.5 * d[1], db = 0.5 * d[1] + d[2];
  out[0] = sa * da; out[2] = sb * db; out[1] = sa * db + sb * da;   // sums to exactly 1
  return true;
}

// Marginal of q (qother: the second focal animal of a pair, forced unobserved, or -1).
// *total = unnormalised P(evidence in the component).
// Optional first tier of the general elimination (-DGFX_GENERAL_SMEM_TIER): tables of up to 3^GFX_WMAX_S entries in
// SHARED memory, one per job of the block, interleaved (entry i of slot t at stab[i * GFX_STAB_SLOTS + t]: conflict
// free) -- the 729-entry private tables live in local memory, 93 KB per warp of 16 jobs, and miss L1 four times out of
// five.  MEASURED AND NOT ENABLED: the kernel's duration is set by its widest jobs, which fail the first tier and run
// twice; k_correct_jobs_general 2.66 -> 5.30 ms per 4096-SNP chunk, exact either way (profiles/r2_summary.md).
#define GFX_WMAX_S 4
#define GFX_TAB_S 81
#define GFX_STAB_SLOTS 64     // 4 warps x KG_ITEMS jobs per block of k_correct_jobs_general
__device__ __noinline__ int query_lazy(const LazyBlanket& L, const BlanketDev& bl, int q, int qother, uint8_t* code,
                                       double* tab, double out[3], double* total, double* stab);
// general query with its own scratch (kept out of line so the common path stays small)
__device__ __noinline__ int query_general(const LazyBlanket& L, const BlanketDev& bl, int q, int qother, double out[3],
                                          double* total, double* stab) {
  double tab[GFX_TAB];
  uint8_t code[GFX_MAXN];
  return query_lazy(L, bl, q, qother, code, tab, out, total, stab);
}
__device__ __noinline__ int query_lazy(const LazyBlanket& L, const BlanketDev& bl, int q, int qother, uint8_t* code,
                                       double* tab, double out[3], double* total, double* stab) {
  uint64_t known = 1ull << q, obs = 0, comp = 1ull << q, queue = 1ull << q, cand = 0;
  if (qother >= 0) known |= 1ull << qother;
  while (queue) {
    const int u = gfx_lowbit(queue);
    queue &= queue - 1;
    uint64_t touch = 0;
    if (L.p1[u] >= 0) touch |= (1ull << L.p1[u]) | (1ull << L.p2[u]);
    const uint64_t km = L.kid[u];
    touch |= km;
    cand |= km;
    for (uint64_t m = km; m; m &= m - 1) {
      const int k = gfx_lowbit(m);
      touch |= (1ull << L.p1[k]) | (1ull << L.p2[k]);
    }
    for (uint64_t m = touch & ~known; m; m &= m - 1) {
      const int t = gfx_lowbit(m);
      const int st = L.status(t);
      code[t] = (uint8_t)st;
      if (st != 3) obs |= 1ull << t;
    }
    known |= touch;
    const uint64_t add = touch & ~obs & ~comp;
    comp |= add;
    queue |= add;
  }
  cand |= comp;
  GFX_COUNT(bl, 0, 1);
  GFX_COUNT(bl, 4, __popcll(known));
  if (comp == (1ull << q) && L.kid[q] == 0 && L.p1[q] >= 0) {
    GFX_COUNT(bl, 1, 1);
    // both parents observed, nothing below: the Mendel column, already normalised
    const int sv = code[L.p1[q]], dv = code[L.p2[q]];
    out[0] = gfx_mendel(0, sv, dv); out[1] = gfx_mendel(1, sv, dv); out[2] = gfx_mendel(2, sv, dv);
    *total = 1.0;
    return 0;
  }
  GfxBlanket b{L.n, L.p1, L.p2, nullptr};
  if (stab) {
    if (!gfx_eliminate<GFX_STAB_SLOTS>(b, q, code, obs, comp, cand, stab, out, GFX_WMAX_S)) {
      GFX_COUNT(bl, 2, 1);
      *total = stab[3 * GFX_STAB_SLOTS];
      return 0;
    }
  }
  int bad = gfx_eliminate(b, q, code, obs, comp, cand, tab, out, GFX_WMAX);
  GFX_COUNT(bl, 2, 1);
  if (bad) {
    GFX_COUNT(bl, 3, 1);
    int slot = (int)((blockIdx.x * blockDim.x + threadIdx.x) % GFX_BIG_SLOTS);
    for (;;) {
      if (atomicCAS(bl.big_flags + slot, 0, 1) == 0) break;
      slot = (slot + 1) % GFX_BIG_SLOTS;
      __nanosleep(200);
    }
    double* big = bl.big_pool + (size_t)slot * GFX_TAB_BIG;
    bad = gfx_eliminate(b, q, code, obs, comp, cand, big, out, GFX_WMAX_BIG);
    *total = big[3];
    __threadfence();
    atomicExch(bl.big_flags + slot, 0);
    if (bad) {
      // still too wide: one of the few huge tables (milliseconds per query, but a result instead of an error)
      int hs = (int)((blockIdx.x * blockDim.x + threadIdx.x) % GFX_HUGE_SLOTS);
      for (;;) {
        if (atomicCAS(bl.big_flags + GFX_BIG_SLOTS + hs, 0, 1) == 0) break;
        hs = (hs + 1) % GFX_HUGE_SLOTS;
        __nanosleep(1000);
      }
      double* huge = bl.huge_pool + (size_t)hs * GFX_TAB_HUGE;
      bad = gfx_eliminate(b, q, code, obs, comp, cand, huge, out, GFX_WMAX_HUGE);
      *total = huge[3];
      __threadfence();
      atomicExch(bl.big_flags + GFX_BIG_SLOTS + hs, 0);
    }
    return bad;
  }
  *total = tab[3];
  return 0;
}

struct GenericRankArgs {
  const uint32_t* packed;
  int64_t wpr, n_snps;
  uint16_t* raw;
  int64_t ld_raw;
  const int32_t* rows;
  int n_rows_sel;
  const int32_t* rank_blanket;
  const int32_t* child_off;
  const int32_t* child_row;
  const int32_t* anc6;
  BlanketDev bl;
  const uint64_t* bl_kid;
  int* status;
  unsigned long long* hist;   // nullptr: plain scores (missing cells 1.0), else histogram + 0xFFFF marker
};

// Rows whose blanket is not the plain tree.  A warp scans 32 words of such a row: words in which both parents
// are observed in all 16 cells are d-separated from the loop and belong to k_mendel_rank (same test there);
// the others are processed two at a time, one cell per lane, with the full inference.
// one cell per lane (two words of 16 cells per warp step): full inference on the row's blanket
__device__ __forceinline__ void generic_rank_cell(const GenericRankArgs& a, int row, int64_t j, bool mine, int lane, double* tab,
                                                  uint8_t* code) {
  uint32_t r = 0x3C00u;
  if (mine) {
    const int obs = j < a.n_snps ? cell_code(a.packed, a.wpr, row, j) : 3;
    if (obs != 3) {
      const int bid = a.rank_blanket[row];
      const int off = a.bl.off[bid];
      CellCtx cc{a.packed, a.wpr, nullptr, 0, nullptr, nullptr, 0};
      LazyBlanket LB{cc, a.bl.row + off, a.bl.p1 + off, a.bl.p2 + off, a.bl_kid + off, a.bl.off[bid + 1] - off, j, 0u, false};
      double anc[3], tot;
      if (query_lazy(LB, a.bl, a.bl.q0[bid], -1, code, tab, anc, &tot, nullptr)) atomicExch(a.status, 1);
      const uint16_t sc = gfx_anc_score(anc, obs);
      uint32_t m = 0;
      bool kids = false;
      for (int c = a.child_off[row]; c < a.child_off[row + 1]; ++c) {
        const int k = cell_code(a.packed, a.wpr, a.child_row[c], j);
        if (k != 3) kids = true;
        if (k == 0) m += obs;            // [0,1,2][obs]
        else if (k == 2) m += 2 - obs;   // [2,1,0][obs]
      }
      r = kids ? gfx_half_add(gfx_child_score(m), sc) : sc;
    } else if (a.hist) {
      r = 0xFFFFu;
    }
    a.raw[(int64_t)row * a.ld_raw + j] = (uint16_t)r;
  }
  if (a.hist) {
    // one atomic per distinct pattern of the warp
    const bool count = mine && j < a.n_snps;
    const uint32_t act = __ballot_sync(0xffffffffu, count);
    if (count) {
      const uint32_t peers = __match_any_sync(act, r);
      if (lane == __ffs(peers) - 1) atomicAdd(&a.hist[r == 0xFFFFu ? 65536u : r], (unsigned long long)__popc(peers));
    }
  }
}

// Scanning form (behind k_mendel_rank, and behind k_mendel_rank3 when the work list would be too large): a warp scans 32
// words of a loopy row and processes the flagged ones two at a time.
__global__ void __launch_bounds__(128) k_mendel_rank_generic(GenericRankArgs a) {
  const int nwords = (int)((a.n_snps + 15) / 16);
  const int chunks = (nwords + 31) / 32;
  double tab[GFX_TAB];
  uint8_t code[GFX_MAXN];
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < (int64_t)a.n_rows_sel * chunks; t += nwarps) {
    const int sel = (int)(t / chunks);
    const int wbase = (int)(t - (int64_t)sel * chunks) * 32;
    const int row = a.rows[sel];
    const int rS = a.anc6[(int64_t)row * 6], rD = a.anc6[(int64_t)row * 6 + 1];
    bool flag = false;
    if (wbase + lane < nwords) {
      flag = true;
      if (rS >= 0 && rD >= 0) {
        const uint32_t wS = __ldg(a.packed + (int64_t)rS * a.wpr + wbase + lane), wD = __ldg(a.packed + (int64_t)rD * a.wpr + wbase + lane);
        flag = (((wS & (wS >> 1)) | (wD & (wD >> 1))) & 0x55555555u) != 0u;
      }
    }
    uint32_t fm = __ballot_sync(0xffffffffu, flag);
    while (fm) {
      const int wa = __ffs(fm) - 1;
      fm &= fm - 1;
      int wb = -1;
      if (fm) { wb = __ffs(fm) - 1; fm &= fm - 1; }
      const int wsel = lane < 16 ? wa : wb;
      const bool mine = wsel >= 0;
      const int64_t j = (int64_t)(wbase + (mine ? wsel : 0)) * 16 + (lane & 15);
      generic_rank_cell(a, row, j, mine, lane, tab, code);
    }
  }
  // launched with programmatic stream serialisation behind k_mendel_rank / k_mendel_rank3: the two grids write disjoint
  // cells, so this one runs in the shadow of the other's tail and only has to outlive it (stream order for what follows)
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Work-list form (behind k_mendel_rank3, which meets every flagged word of a loopy row anyway and records it): hdr[0] is the
// number of entries (row, word).  With nothing on the list -- genotypes without missing calls -- the grid ends as soon as it
// has seen the count; its launch is hidden behind the producer (programmatic dependent launch).  The last block to leave
// puts the header back to zero: headers live in a ring owned by the schedule and are not cleared by the host.
__global__ void __launch_bounds__(128) k_mendel_rank_generic_wl(GenericRankArgs a, uint32_t* hdr, const uint2* __restrict__ wl) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the list is complete when the producer grid has finished
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = *reinterpret_cast<volatile uint32_t*>(hdr);
  __syncthreads();
  const uint32_t n = s_n;
  if (n != 0u) {
    double tab[GFX_TAB];
    uint8_t code[GFX_MAXN];
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t e = 2u * ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < n; e += 2u * nwarps) {
      const uint32_t mye = e + (lane >> 4);
      const bool mine = mye < n;
      int row = 0;
      int64_t j = 0;
      if (mine) {
        const uint2 ent = wl[mye];
        row = (int)ent.x;
        j = (int64_t)ent.y * 16 + (lane & 15);
      }
      generic_rank_cell(a, row, j, mine, lane, tab, code);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(hdr + 1, 1u) == gridDim.x - 1u) { hdr[1] = 0u; hdr[0] = 0u; }
}

static int launch_mendel_rank(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr, uint16_t* raw,
                              int64_t ld_raw, uint64_t* hist, cudaStream_t stream) {
  if (!s || !packed || !raw) return fail(GFX_E_INVALID, "null argument");
  if (s->device < 0) return fail(GFX_E_INVALID, "host-only schedule passed to a kernel entry point");
  const int64_t nwords = (n_snps + 15) / 16;
  if (n_snps <= 0 || wpr < nwords || ld_raw < nwords * 16 || (ld_raw % 8) != 0 || (((uintptr_t)raw) % 16) != 0)
    return fail(GFX_E_INVALID, "gfx_mendel_rank: need wpr >= ceil(s/16), ld_raw >= 16*ceil(s/16), ld_raw % 8 == 0, raw 16-byte aligned");
  // k_mendel_rank3 needs Mendel's table (rank3_network_matches); GFX_RANK_FORM=2 selects the previous form (A/B runs)
  const char* form_env = getenv("GFX_RANK_FORM");
  // ... and rows that bulk copies can address: 16-byte aligned segments, at most 2^31 bytes per row
  const bool form3 = s->rank_form3 && !(form_env ? form_env[0] == '2' : s->rank_form_pref.load() == 2) && (wpr % 4) == 0 && (((uintptr_t)packed) % 16) == 0 &&
                     wpr <= 0x1FFFFFFFll;
  // task table for this matrix width (built once, cached in the schedule)
  const RankRow* tasks_dev = nullptr;
  uint32_t n_tasks = 0;
  {
    std::lock_guard<std::mutex> lock(const_cast<gfx_schedule*>(s)->rank_mutex);
    gfx_schedule* ms = const_cast<gfx_schedule*>(s);
    for (auto& t : ms->rank_tasks)
      if (t.nwords == nwords && t.form == (form3 ? 3 : 2)) { tasks_dev = t.dev; n_tasks = t.n; }
    if (!tasks_dev) {
      std::vector<RankRow> tasks;
      tasks.reserve((size_t)s->h.n_rows * ((nwords + 127) / 128 + 1));
      for (const RankRow& r : s->rank_rows_h) {
        int64_t cw;
        if (form3) {
          // rows are cut into equal parts of at most R3_WORDS_* words (multiples of 32): long tasks for cheap rows (per-task
          // set-up), short ones for rows with many children (the heaviest tasks start first, but a 1 000-kid sire must not be one)
          const int64_t maxw = r.nk == 0 ? R3_WORDS_LIGHT : (r.nk < 8 ? R3_WORDS_MID : (r.nk < 64 ? R3_WORDS_HEAVY : R3_WORDS_HUGE));
          const int64_t parts = (nwords + maxw - 1) / maxw;
          cw = ((nwords + parts - 1) / parts + 31) / 32 * 32;
        } else {
          cw = r.nk >= 8 ? R2_WORDS_HEAVY : (r.nk >= 1 ? R2_WORDS_MID : R2_WORDS_LIGHT);
        }
        for (int64_t w = 0; w < nwords; w += cw) {
          RankRow t = r;
          t.pad0 = (int32_t)w;
          t.pad1 = (int32_t)std::min<int64_t>(w + cw, nwords);
          tasks.push_back(t);
        }
      }
      if (tasks.size() >= 0xFFFF0000ull || nwords > 0x7FFFFF00ll)
        return fail(GFX_E_INVALID, "gfx_mendel_rank: matrix too large for one call (more than 2^32 tasks)");
      RankRow* dev = nullptr;
      GFX_CUDA(cudaMalloc((void**)&dev, tasks.size() * sizeof(RankRow)));
      cudaError_t e = cudaMemcpy(dev, tasks.data(), tasks.size() * sizeof(RankRow), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { cudaFree(dev); return fail(GFX_E_CUDA, std::string("task table upload: ") + cudaGetErrorString(e)); }
      // tables are never evicted: another lane (host thread on the same schedule) may be about to launch on one, and a
      // table is 32 bytes per task -- a few MB per distinct matrix width
      ms->rank_tasks.push_back({nwords, form3 ? 3 : 2, dev, (uint32_t)tasks.size()});
      tasks_dev = dev;
      n_tasks = (uint32_t)tasks.size();
    }
  }
  RankArgs a{packed, wpr, (int)nwords, n_snps, raw, ld_raw, s->h.n_rows, s->anc6, tasks_dev, n_tasks, s->rank_kids,
             s->rank_lut, s->rank_lut6, s->rank_class_sc, (unsigned long long*)hist, nullptr, nullptr, 0u};
  {
    // once per process (one process per GPU); guarded: several lanes call this concurrently
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, []() {
      attr_err = cudaFuncSetAttribute(k_mendel_rank<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_DYN_SMEM);
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(k_mendel_rank<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_DYN_SMEM);
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(k_mendel_rank3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R3_DYN_SMEM);
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(k_mendel_rank3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R3_DYN_SMEM);
    });
    GFX_CUDA(attr_err);
  }
  if (hist) GFX_CUDA(cudaMemsetAsync(hist, 0, 65537 * sizeof(uint64_t), stream));
  uint2* wl = nullptr;
  uint32_t* wl_hdr = nullptr;
  int wl_slot = -1;
  if (form3) {
    int blocks = (int)std::min<int64_t>(((int64_t)n_tasks + (R3_THREADS / 32) - 1) / (R3_THREADS / 32), s->num_sms);
    if (blocks < 1) blocks = 1;
    // work list of the loopy rows: worst case every word of every loopy row; beyond 256 MB the consumer scans instead
    const int64_t wl_cap = (int64_t)s->n_loopy_rows * nwords;
    if (s->n_loopy_rows > 0 && wl_cap <= (32ll << 20)) {
      gfx_schedule* ms = const_cast<gfx_schedule*>(s);
      std::lock_guard<std::mutex> lock(ms->rank_mutex);
      wl_slot = (int)(ms->rank_wl_next++ % 8u);
      gfx_schedule::WlSlot& W = ms->rank_wl[wl_slot];
      if (!W.done) GFX_CUDA(cudaEventCreateWithFlags(&W.done, cudaEventDisableTiming));
      if (W.cap < (size_t)wl_cap) {
        if (W.used) GFX_CUDA(cudaEventSynchronize(W.done));
        if (W.buf) cudaFree(W.buf);
        W.buf = nullptr; W.cap = 0; W.used = false;
        GFX_CUDA(cudaMalloc((void**)&W.buf, sizeof(uint2) * (size_t)wl_cap));
        W.cap = (size_t)wl_cap;
      }
      if (W.used) GFX_CUDA(cudaStreamWaitEvent(stream, W.done, 0));
      wl = W.buf;
      wl_hdr = s->rank_wl_hdr + 32 * wl_slot;
      a.wl_hdr = wl_hdr; a.wl = wl; a.wl_cap = (uint32_t)wl_cap;
    }
    if (s->rank_timing) GFX_CUDA(cudaEventRecord(s->rank_ev0, stream));
    if (hist) k_mendel_rank3<true><<<blocks, R3_THREADS, R3_DYN_SMEM, stream>>>(a);
    else k_mendel_rank3<false><<<blocks, R3_THREADS, R3_DYN_SMEM, stream>>>(a);
    if (s->rank_timing) GFX_CUDA(cudaEventRecord(s->rank_ev1, stream));
  } else {
    int blocks = (int)std::min<int64_t>(((int64_t)n_tasks + (R2_THREADS / 32) - 1) / (R2_THREADS / 32), s->num_sms * R2_CTAS);
    if (blocks < 1) blocks = 1;
    if (s->rank_timing) GFX_CUDA(cudaEventRecord(s->rank_ev0, stream));
    if (hist) k_mendel_rank<true><<<blocks, R2_THREADS, R2_DYN_SMEM, stream>>>(a);
    else k_mendel_rank<false><<<blocks, R2_THREADS, R2_DYN_SMEM, stream>>>(a);
    if (s->rank_timing) GFX_CUDA(cudaEventRecord(s->rank_ev1, stream));
  }
  GFX_LAUNCHED();
  if (s->n_loopy_rows > 0) {
    GenericRankArgs g{packed, wpr, n_snps, raw, ld_raw, s->loopy_rows, s->n_loopy_rows, s->rank_blanket, s->child_off, s->child_row,
                      s->anc6,
                      BlanketDev{s->bl_off, s->bl_row, s->bl_p1, s->bl_p2, s->bl_last, s->bl_q0, s->bl_q1, s->big_pool, s->big_flags, s->huge_pool, s->debug ? s->counters : nullptr},
                      s->bl_kid, s->status, (unsigned long long*)hist};
    const int64_t total = (int64_t)s->n_loopy_rows * ((nwords + 31) / 32) * 32;   // one warp per 32 words
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(wl ? s->num_sms * 4 : grid_for(total, 128, s->num_sms, 8)));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool pdl = []() { const char* e = getenv("GFX_RANK_PDL"); return !(e && e[0] == '0'); }();
    cfg.numAttrs = pdl ? 1 : 0;
    if (wl) {
      GFX_CUDA(cudaLaunchKernelEx(&cfg, k_mendel_rank_generic_wl, g, wl_hdr, (const uint2*)wl));
      gfx_schedule* ms = const_cast<gfx_schedule*>(s);
      std::lock_guard<std::mutex> lock(ms->rank_mutex);
      GFX_CUDA(cudaEventRecord(ms->rank_wl[wl_slot].done, stream));
      ms->rank_wl[wl_slot].used = true;
    } else {
      GFX_CUDA(cudaLaunchKernelEx(&cfg, k_mendel_rank_generic, g));
    }
    GFX_LAUNCHED();
  }
  return GFX_OK;
}

// k_mendel_rank3 is the kernel for matrices whose words are mostly complete (its fast paths need own, sire and dam cells
// observed); on matrices full of missing calls nearly every word takes the per-cell path, for which the second form is the
// leaner host (no staging, 80 KB instead of 134 KB of shared memory: config 4 runs 15.4 ms per step with it, 20.5 ms with the
// third form).  The caller knows the missing fraction from the previous chunk's histogram (bin 65536) and says which to use.
extern "C" int gfx_rank_form(gfx_schedule* s, int form) {
  if (!s) return fail(GFX_E_INVALID, "null argument");
  if (form != 0 && form != 2 && form != 3) return fail(GFX_E_INVALID, "gfx_rank_form: 0 (automatic), 2 or 3");
  s->rank_form_pref.store(form == 3 ? 0 : form);
  return GFX_OK;
}

extern "C" int gfx_rank_timing(gfx_schedule* s, int enable) {
  if (!s || s->device < 0) return fail(GFX_E_INVALID, "bad schedule");
  if (enable && !s->rank_ev0) {
    GFX_CUDA(cudaEventCreate(&s->rank_ev0));
    GFX_CUDA(cudaEventCreate(&s->rank_ev1));
  }
  s->rank_timing = enable != 0;
  return GFX_OK;
}

extern "C" int gfx_rank_last_kernel_ms(gfx_schedule* s, float* ms) {
  if (!s || !ms || !s->rank_ev0) return fail(GFX_E_INVALID, "gfx_rank_last_kernel_ms: timing was never switched on");
  GFX_CUDA(cudaEventSynchronize(s->rank_ev1));
  GFX_CUDA(cudaEventElapsedTime(ms, s->rank_ev0, s->rank_ev1));
  return GFX_OK;
}

extern "C" int gfx_mendel_rank(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr,
                               uint16_t* raw, int64_t ld_raw, void* stream) {
  return launch_mendel_rank(s, packed, n_snps, wpr, raw, ld_raw, nullptr, (cudaStream_t)stream);
}

// Fused variant: scores + their 65 537-bin histogram in one pass; missing cells carry the marker 0xFFFF
// (what gfx_mendel_rank followed by gfx_rank_hist produces).
extern "C" int gfx_mendel_rank_hist(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr,
                                    uint16_t* raw, int64_t ld_raw, uint64_t* hist, void* stream) {
  if (!hist) return fail(GFX_E_INVALID, "gfx_mendel_rank_hist: null histogram");
  return launch_mendel_rank(s, packed, n_snps, wpr, raw, ld_raw, hist, (cudaStream_t)stream);
}

// ---- histogram of raw scores ----------------------------------------------------------------------
// Scores are multiples of float16(1/3) and float16(1/4)-ish units, so a handful of patterns carry almost
// all the mass: zeros and the four hottest patterns are counted in registers, patterns in [0x3000, 0x5000)
// (values 0.125 .. 32) in a 32 KB shared-memory histogram, anything else (NaN, tiny, huge) in global memory.
#define GFX_HIST_LO 0x3000u
#define GFX_HIST_N 0x2000u
__global__ void __launch_bounds__(256) k_rank_hist(uint16_t* __restrict__ raw, int64_t ld_raw,
                                                   const uint32_t* __restrict__ packed, int64_t wpr, int64_t n,
                                                   int64_t n_snps, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t s_hist[GFX_HIST_N];
  for (int i = threadIdx.x; i < (int)GFX_HIST_N; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int64_t nwords = (n_snps + 15) / 16;
  const int64_t total = n * nwords;
  const uint32_t H0 = 0x3555u, H1 = 0x3955u, H2 = 0x3800u, H3 = 0x3C00u;   // 1/3, 2/3, 1/2, 1
  uint32_t zeros = 0, missing = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nwords, w = i - r * nwords;
    const uint32_t pw = __ldg(packed + r * wpr + w);
    const uint4* src = reinterpret_cast<const uint4*>(raw + r * ld_raw + w * 16);
    const uint4 v0 = src[0], v1 = src[1];
    const uint32_t q[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const int lim = (int)min((int64_t)16, n_snps - w * 16);
    uint32_t miss_mask = pw & (pw >> 1) & 0x55555555u;   // code 3
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k >= lim) break;
      const uint32_t h = (q[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
      if ((miss_mask >> (2 * k)) & 1u) { ++missing; raw[r * ld_raw + w * 16 + k] = 0xFFFFu; }  // E = 1 by definition (:603)
      else if (h == 0u) ++zeros;
      else if (h == H0) ++c0;
      else if (h == H1) ++c1;
      else if (h == H2) ++c2;
      else if (h == H3) ++c3;
      else if (h - GFX_HIST_LO < GFX_HIST_N) atomicAdd(&s_hist[h - GFX_HIST_LO], 1u);
      else atomicAdd(&hist[h], 1ull);
    }
  }
  // warp-reduce the register counters
  for (int o = 16; o > 0; o >>= 1) {
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    missing += __shfl_xor_sync(0xffffffffu, missing, o);
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (zeros) atomicAdd(&hist[0], (unsigned long long)zeros);
    if (missing) atomicAdd(&hist[65536], (unsigned long long)missing);
    if (c0) atomicAdd(&s_hist[H0 - GFX_HIST_LO], c0);
    if (c1) atomicAdd(&s_hist[H1 - GFX_HIST_LO], c1);
    if (c2) atomicAdd(&s_hist[H2 - GFX_HIST_LO], c2);
    if (c3) atomicAdd(&s_hist[H3 - GFX_HIST_LO], c3);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (int)GFX_HIST_N; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[GFX_HIST_LO + i], (unsigned long long)s_hist[i]);
}

extern "C" int gfx_rank_hist(uint16_t* raw, int64_t ld_raw, const uint32_t* packed, int64_t wpr, int64_t n,
                             int64_t n_snps, uint64_t* hist, void* stream) {
  if (!raw || !packed || !hist || n <= 0 || n_snps <= 0) return fail(GFX_E_INVALID, "gfx_rank_hist: bad argument");
  GFX_CUDA(cudaMemsetAsync(hist, 0, 65537 * sizeof(uint64_t), (cudaStream_t)stream));
  const int64_t total = n * ((n_snps + 15) / 16);
  k_rank_hist<<<grid_for(total, 256, device_sms(), 6), 256, 0, (cudaStream_t)stream>>>(
      raw, ld_raw, packed, wpr, n, n_snps, (unsigned long long*)hist);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ---- per-animal sum of E ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rank_rowsum(const uint16_t* __restrict__ raw, int64_t ld_raw,
                                                     const uint32_t* __restrict__ packed, int64_t wpr, int64_t n,
                                                     int64_t n_snps, const double* __restrict__ elut, double missing_value,
                                                     double* __restrict__ out) {
  // one warp per row, fixed lane-strided order + fixed shuffle tree => deterministic sums
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    double acc = 0.0;
    for (int64_t j = lane; j < n_snps; j += 32) {
      const uint32_t c = (__ldg(packed + r * wpr + (j >> 4)) >> (2 * (j & 15))) & 3u;
      acc += c == 3u ? missing_value : __ldg(elut + raw[r * ld_raw + j]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc;
  }
}

extern "C" int gfx_rank_rowsum(const uint16_t* raw, int64_t ld_raw, const uint32_t* packed, int64_t wpr, int64_t n,
                               int64_t n_snps, const double* elut, double missing_value, double* out, void* stream) {
  if (!raw || !packed || !elut || !out || n <= 0 || n_snps <= 0) return fail(GFX_E_INVALID, "gfx_rank_rowsum: bad argument");
  k_rank_rowsum<<<grid_for(n * 32, 256, device_sms(), 8), 256, 0, (cudaStream_t)stream>>>(raw, ld_raw, packed, wpr, n, n_snps,
                                                                                          elut, missing_value, out);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ---- per-animal tally exactly as pandas computes it in the reference ------------------------------------
// error_checker.py:193-204 / preindex.py:259-266 keep the transformed rank matrix in FLOAT16 and sum it with
// DataFrame.sum(axis=1): a float16 accumulator per animal that takes the SNP columns one after the other, every
// addition carried out in float32 and rounded to float16 (numpy's half add).  That chain is order dependent, so it
// is reproduced literally: one thread per selected animal walks its row in column order; the next 64 bytes of
// scores are in flight while the current 32 cells are added.  e16[raw pattern] is the float16 E of every score
// pattern (built on the host with numpy); e16[0xFFFF] is the value of cells that were missing in the input.
__global__ void __launch_bounds__(128) k_rank_rowsum_f16(const uint16_t* __restrict__ raw, int64_t ld_raw, int64_t n_snps,
                                                         const uint16_t* __restrict__ e16, const int32_t* __restrict__ rows,
                                                         int n_sel, uint16_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sel) return;
  const int64_t r = rows ? __ldg(rows + t) : t;
  const uint16_t* src = raw + r * ld_raw;          // ld_raw is a multiple of 64: rows are 128-byte aligned
  const int64_t nblk = n_snps >> 5;                 // blocks of 32 cells = 4 x uint4
  float acc = 0.0f;                                 // always holds a float16 value
  uint4 nx[4];
  if (nblk > 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) nx[u] = __ldg(reinterpret_cast<const uint4*>(src) + u);
  }
  for (int64_t b = 0; b < nblk; ++b) {
    uint4 cur[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nx[u];
    if (b + 1 < nblk) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = __ldg(reinterpret_cast<const uint4*>(src + (b + 1) * 32) + u);
    }
    float x[32];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t wv[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x[u * 8 + 2 * k] = __half2float(__ushort_as_half(__ldg(e16 + (wv[k] & 0xFFFFu))));
        x[u * 8 + 2 * k + 1] = __half2float(__ushort_as_half(__ldg(e16 + (wv[k] >> 16))));
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc = __half2float(__float2half_rn(acc + x[i]));
  }
  for (int64_t j = nblk << 5; j < n_snps; ++j)
    acc = __half2float(__float2half_rn(acc + __half2float(__ushort_as_half(__ldg(e16 + __ldg(src + j))))));
  out[t] = __half_as_ushort(__float2half_rn(acc));
}

extern "C" int gfx_rank_rowsum_f16(const uint16_t* raw, int64_t ld_raw, int64_t n, int64_t n_snps, const uint16_t* e16,
                                   const int32_t* rows, int32_t n_sel, uint16_t* out, void* stream) {
  if (!raw || !e16 || !out || n <= 0 || n_snps <= 0 || n_sel < 0 || ld_raw < n_snps || (ld_raw & 63))
    return fail(GFX_E_INVALID, "gfx_rank_rowsum_f16: bad argument");
  if (!rows && n_sel != n) return fail(GFX_E_INVALID, "gfx_rank_rowsum_f16: n_sel must equal n without a row list");
  if (n_sel == 0) return GFX_OK;
  k_rank_rowsum_f16<<<(n_sel + 127) / 128, 128, 0, (cudaStream_t)stream>>>(raw, ld_raw, n_snps, e16, rows, n_sel, out);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ---- correctPair / correctSingle + apply, fused per focal animal -----------------------------------
//
// One warp owns 32 consecutive SNPs of one focal animal of the generation and walks that animal's
// jobs (mating pairs it takes part in, canonical order).  Everything except the animal's own cell is
// read from the generation-start snapshot, exactly like the reference's workers
// (correct_genotypes3.py:687-689); the own cell is carried in a register and updated job after job,
// like the serial apply loop (:714-845).  The children term of the focal animal does not depend on
// the mate, so it is evaluated at most once per (animal, SNP).
struct FocalArgs {
  CellCtx c;              // c.live = snapshot
  uint32_t* live;         // matrix being corrected
  int64_t n_snps;
  int nf;                 // 2 = pairs, 1 = singles
  int n_focal;
  const int32_t* focal_row;
  const int32_t* focal_off;
  const int32_t* entry;   // job * nf + side
  const int32_t* job_row0;
  const int32_t* job_row1;
  const int32_t* job_bl;
  BlanketDev bl;
  const uint64_t* bl_kid;
  const uint8_t* bl_tree;
  const int32_t* bl_slotrow;
  const uint16_t* bl_slotbad;
  const int32_t* child_off;
  const int32_t* child_row;
  const int32_t* child_other;
  const int32_t* prow;
  const double* kv;
  gfx_correct_params p;
  double* post;           // optional [(job*nf + side) * n_snps + j] * 3
  int32_t* tallies;
  int n_rows;
  int* status;
  uint32_t* deferred;     // overflow bitmap, 1 bit per (focal, SNP): cell needs the general kernel
  int64_t dwords;         // bitmap words per focal animal
  unsigned long long* wl_count;  // deferred cells so far
  unsigned long long* task_count; // next (animal, chunk) task of the lean kernel
  unsigned long long* item_count; // deferred jobs so far
  const int32_t* focal_order;    // focal indices, heaviest first
  uint2* wl;              // per deferred cell: (focal, snp)
  uint2* st;              // per deferred cell: x = live state | first unresolved job << 8, y = applied | excluded << 16
  int4* skc;              // per deferred cell: class counts of the children term (K0, K1, K2, usable children)
  uint8_t* sbits;         // per deferred cell: decision bits of every job (0xFF = to be computed), at sboff[cell]
  uint32_t* sboff;        // per deferred cell: first byte of its run in sbits (one byte per job of its animal)
  unsigned long long sb_cap;
  uint2* items;           // deferred jobs: (cell slot, job index within the animal)
  unsigned long long wl_cap, item_cap;
  uint2* items_sorted;    // the same items grouped by job entry (same blanket, same side): convergent warps
  uint32_t *key_count, *key_off, *key_cur;   // per job entry of the generation: items, first position, cursor
  int n_keys, e_base;     // keys = entry index - e_base
  int maxj;
};

// Children term for litters of up to 42 genotyped children: classes packed 4 bits each in registers,
// numpy's 8-accumulator block sum run straight from them (one leaf of the pairwise scheme).
#define GFX_KIDS_FAST 96
#define GFX_KIDS_WORDS (GFX_KIDS_FAST / 8)
// classes of the usable children, 4 bits each, in shared memory: word w of thread t at cls[w * 128 + t]
struct NibStream {
  const uint32_t* cls;   // already offset by the thread index
  uint32_t word;
  int a, cur, ninc;
  const double* kv;
  __device__ __forceinline__ void operator()(double v[3]) {
    if (cur == ninc) { cur = 0; ++a; }
    if ((cur & 7) == 0) word = cls[(cur >> 3) * 128];
    const int c = (int)((word >> ((cur & 7) * 4)) & 15u);
    const double* e = kv + c * 9 + 3 * a;
    v[0] = e[0]; v[1] = e[1]; v[2] = e[2];
    ++cur;
  }
};
// one copy of numpy's block sum for the whole kernel (instruction-cache footprint matters here)
__device__ __noinline__ void nib_leaf(NibStream& next, int n, double* res) { gfx_pw_leaf(next, n, res); }
__device__ __noinline__ void nib_pairwise(NibStream& next, int n, double* res) {
  if (n <= 128) { nib_leaf(next, n, res); return; }
  int fn[16], fstage[16];
  double fleft[16][3];
  int sp = 0;
  fn[0] = n; fstage[0] = 0;
  double val[3] = {0, 0, 0};
  while (sp >= 0) {
    const int m = fn[sp];
    if (m <= 128) { nib_leaf(next, m, val); --sp; continue; }
    int n2 = m / 2; n2 -= n2 % 8;
    if (fstage[sp] == 0) { fstage[sp] = 1; fn[sp + 1] = n2; fstage[sp + 1] = 0; ++sp; }
    else if (fstage[sp] == 1) {
      fleft[sp][0] = val[0]; fleft[sp][1] = val[1]; fleft[sp][2] = val[2];
      fstage[sp] = 2; fn[sp + 1] = m - n2; fstage[sp + 1] = 0; ++sp;
    } else {
      val[0] = fleft[sp][0] + val[0]; val[1] = fleft[sp][1] + val[1]; val[2] = fleft[sp][2] + val[2];
      --sp;
    }
  }
  res[0] = val[0]; res[1] = val[1]; res[2] = val[2];
}
__device__ __noinline__ void kids_term_fast(const CellCtx& c, const int32_t* child_row, const int32_t* child_other, int nk,
                                            int64_t j, const double* kv, uint32_t* cls, double out[3]) {
  int ninc = 0;
  uint32_t word = 0;
  const int64_t wo = j >> 4;
  const int sh = 2 * (int)(j & 15);
  for (int i = 0; i < nk; ++i) {
    const int k = (int)((__ldg(c.live + (int64_t)__ldg(child_row + i) * c.wpr + wo) >> sh) & 3u);
    if (k > 2) continue;
    const int o = __ldg(child_other + i);
    const int pc = o >= 0 ? (int)((__ldg(c.live + (int64_t)o * c.wpr + wo) >> sh) & 3u) : 3;
    word |= (uint32_t)(k * 4 + pc) << ((ninc & 7) * 4);
    ++ninc;
    if ((ninc & 7) == 0) { cls[((ninc >> 3) - 1) * 128] = word; word = 0; }
  }
  if (ninc & 7) cls[(ninc >> 3) * 128] = word;
  if (ninc == 0) { out[0] = out[1] = out[2] = 1.0 / 3.0; return; }
  NibStream next{cls, 0u, 0, 0, ninc, kv};
  double tot[3];
  nib_pairwise(next, 3 * ninc, tot);
  gfx_kids_finish(tot, ninc, out);
}

#define GFX_KIDS_LOCAL 160
// children term of `row` at SNP j from the snapshot (bayesnet/model.py:65-89); false = no genotyped child
__device__ __noinline__ bool kids_term_dev(const CellCtx& c, const int32_t* child_row, const int32_t* child_other, int nk,
                                           int64_t j, const double* kv, double out[3]) {
  if (nk <= 0) return false;
  if (nk <= GFX_KIDS_LOCAL) {
    uint8_t cls[GFX_KIDS_LOCAL];
    int ninc = 0;
    for (int i = 0; i < nk; ++i) {
      const int k = cell_code(c.live, c.wpr, child_row[i], j);
      if (k > 2) continue;
      const int o = child_other[i];
      const int pc = o >= 0 ? cell_code(c.live, c.wpr, o, j) : 3;
      cls[ninc++] = (uint8_t)(k * 4 + pc);
    }
    gfx_kids_term_cls(ninc, cls, kv, out);
    return true;
  }
  struct KC { const uint32_t* live; int64_t wpr; const int32_t* rows; int64_t j;
    __device__ int operator()(int i) const { return cell_code(live, wpr, rows[i], j); } };
  struct PC { const uint32_t* live; int64_t wpr; const int32_t* rows; int64_t j;
    __device__ int operator()(int i) const { const int r = rows[i]; return r >= 0 ? cell_code(live, wpr, r, j) : 3; } };
  KC kc{c.live, c.wpr, child_row, j};
  PC pc{c.live, c.wpr, child_other, j};
  return gfx_kids_term(nk, kc, pc, kv, out);
}

#define GFX_CORRECT_CHUNK 1024

// ---- children term as class counts, decisions in exact arithmetic ------------------------------------------------
// bayesnet/model.py:65-89 turns every usable child into one of four vectors: A = [2/3, 1/3, 0],
// B = [1/3, 1/3, 1/3], C = [0, 1/3, 2/3] or Z = 0, selected by (child state, other parent's state); the children
// term is their mean divided by its sum.  In exact arithmetic that is K / sum(K) with the INTEGERS
// K = sum over the usable children of 3 * vector, so the posterior of a job is N / sum(N), N[x] = anc[x] * K[x]
// (anc: exact dyadic rationals), and both decision tests of correct_genotypes3.py:741-744,809 / :932 compare
// (max N - N[i]) / sum(N) with a constant.  The reference evaluates the same quantities in float64 with a relative
// error below 1e-14, so the exact comparison gives the reference's answer whenever the margin exceeds
// GFX_GUARD_EPS; inside that band (exact ties like 0.75 - 0.25 vs 0.5 are common) the job is replayed literally,
// in numpy's summation order, by the deferred-job kernels.  Calls stay bit-exact, the fast path needs no float64
// children sum at all.
// Flags of a deferred job item (high bits of its job number): what the kernels before have already found out about it,
// so that no walk that is known to fail is repeated (the deferred kernels re-ran the plain peeling of every item and the
// general round re-ran the message passing that had just failed).
#define GFX_ITEM_LITERAL 0x80000000u      // plain peeling answered the job, only the literal children sum is missing
#define GFX_ITEM_OWN_NOPEEL 0x40000000u   // plain peeling fails on the focal animal's side
#define GFX_ITEM_MATE_NOPEEL 0x20000000u  // ... succeeds there and fails on the mate's side
#define GFX_ITEM_OWN_NOX 0x10000000u      // message passing fails on the focal animal's side
#define GFX_ITEM_MATE_NOX 0x08000000u     // ... succeeds there and fails on the mate's side
#define GFX_ITEM_FLAGS 0xF8000000u
#define GFX_GUARD_EPS 1e-9
#define GFX_BITS_GUARD 0xFFFFFFFFu
struct KidCounts { int k0, k1, k2, ninc; };   // ninc: usable (genotyped, non-missing) children
// s_ku[k * 4 + p]: x = K0 | K1 << 16, y = K2 | 1 << 16 of a child in state k whose other parent shows p (3 = unknown)
__device__ __forceinline__ void kid_units_init(uint2* s_ku) {
  if (threadIdx.x < 16) {
    const int k = threadIdx.x >> 2, pp = threadIdx.x & 3;
    int v = 3;                                   // 0 A, 1 B, 2 C, 3 Z / not usable
    if (k == 0) v = pp == 2 ? 3 : 0;
    else if (k == 1) v = pp == 0 ? 2 : (pp == 2 ? 0 : 1);
    else if (k == 2) v = pp == 0 ? 3 : 2;
    const uint32_t k0 = v == 0 ? 2u : (v == 1 ? 1u : 0u), k1 = v == 3 ? 0u : 1u, k2 = v == 2 ? 2u : (v == 1 ? 1u : 0u);
    s_ku[threadIdx.x] = make_uint2(k0 | (k1 << 16), k2 | (k < 3 ? 1u << 16 : 0u));
  }
}
#define GFX_KIDS_COUNT_MAX 30000   // 16-bit fields
__device__ __forceinline__ KidCounts kids_counts(const CellCtx& c, const int32_t* __restrict__ child_row,
                                                 const int32_t* __restrict__ child_other, int nk, int64_t wo, int sh,
                                                 const uint2* s_ku) {
  uint32_t a01 = 0, a2n = 0;
  for (int i = 0; i < nk; ++i) {   // (#pragma unroll 4 measured: no change)
    const int r = __ldg(child_row + i), o = __ldg(child_other + i);
    const uint32_t k = (__ldg(c.live + (int64_t)r * c.wpr + wo) >> sh) & 3u;
    const uint32_t pp = o >= 0 ? ((__ldg(c.live + (int64_t)o * c.wpr + wo) >> sh) & 3u) : 3u;
    const uint2 u = s_ku[k * 4u + pp];
    a01 += u.x;
    a2n += u.y;
  }
  return KidCounts{(int)(a01 & 0xFFFFu), (int)(a01 >> 16), (int)(a2n & 0xFFFFu), (int)(a2n >> 16)};
}
// Decision bits (gfx_decision_bits layout) from unnormalised posterior numerators, or GFX_BITS_GUARD when a test
// falls inside the guard band.  n >= 0, at least one > 0.
__device__ __forceinline__ uint32_t guard_bits(double n0, double n1, double n2, int obs, double tie, double thr) {
  const double z = (n0 + n1) + n2;
  const double mx = fmax(n0, fmax(n1, n2));
  const double eps = GFX_GUARD_EPS * z, bound = tie * z;
  uint32_t r = 0x80u;
  const double d0 = mx - n0, d1 = mx - n1, d2 = mx - n2;
  bool guard = fabs(d0 - bound) <= eps || fabs(d1 - bound) <= eps || fabs(d2 - bound) <= eps;
  if (d0 <= bound) r |= 1u;
  if (d1 <= bound) r |= 2u;
  if (d2 <= bound) r |= 4u;
  if (obs <= 2) {
    const double dd = obs == 0 ? d0 : (obs == 1 ? d1 : d2), bt = thr * z;
    guard = guard || fabs(dd - bt) <= eps;
    if (dd >= bt) r |= 8u;
  }
  return guard ? GFX_BITS_GUARD : r;
}
// The cases in which the children term is known without summing: no genotyped child (term absent), no usable child
// ([1/3, 1/3, 1/3], model.py:89) or only vanished rows (zeros).  Returns true and the literal term then.
__device__ __forceinline__ bool kids_literal_known(int nk, const KidCounts& K, bool* has_kids, double kids[3]) {
  *has_kids = nk > 0;
  if (nk <= 0) { kids[0] = kids[1] = kids[2] = 0.0; return true; }
  if (K.ninc == 0) { kids[0] = kids[1] = kids[2] = 1.0 / 3.0; return true; }
  if (K.k0 + K.k1 + K.k2 == 0) { kids[0] = kids[1] = kids[2] = 0.0; return true; }
  return false;
}

// Posterior of ONE deferred job of a cell, reduced to the decision bits of gfx_decision_bits (0 = the job is not
// suspect at this SNP or yields no result).  Independent of the live state, so jobs of a cell can be evaluated in
// any order / by different threads; only the apply walk is sequential.
// MODE 1: slot peeling + tree message passing; 2: + general elimination.  false = needs a higher mode.
// The decision uses the class counts of the parked cell; the children term is summed literally (numpy's order) only
// when a test falls inside the guard band, when the fast kernel already found that (`literal`), or for posteriors.
template <int MODE>
__device__ __forceinline__ bool job_bits(const FocalArgs& a, int row, int ent, int64_t j, int obs0, uint32_t rxf,
                                         const KidCounts& K, uint32_t flags, uint32_t* bits_out, double* stab,
                                         uint32_t* fail_flag = nullptr) {
  const bool literal = (flags & GFX_ITEM_LITERAL) != 0u;
  const uint32_t tp = a.c.tp_raw;
  const uint32_t idxf = rxf == 0xFFFFu ? 65536u : rxf;
  const int job = a.nf == 2 ? (ent >> 1) : ent;
  const int side = a.nf == 2 ? (ent & 1) : 0;
  *bits_out = 0;
  bool sus = rxf >= tp;                                     // :181-185 / :342
  uint32_t pidx;                                            // pattern of pmax (:220-228 / :374-376)
  if (a.nf == 2) {
    const int mrow = side == 0 ? a.job_row1[job] : a.job_row0[job];
    const uint32_t rxm = cell_rx(a.c, mrow, j);
    sus = sus || (rxm >= tp);
    if (!sus) return true;
    const uint32_t idxm = rxm == 0xFFFFu ? 65536u : rxm;
    const int obs_m = cell_code(a.c.live, a.c.wpr, mrow, j);
    if (obs0 != 3 && obs_m != 3) pidx = idxf > idxm ? idxf : idxm;
    else if (obs0 == 3 && obs_m == 3) pidx = 65536u;
    else pidx = obs0 != 3 ? idxf : idxm;
  } else {
    if (!sus) return true;
    pidx = obs0 != 3 ? idxf : 65536u;
  }
  GFX_COUNT(a.bl, 6, 1);
  const int bid = a.job_bl[job];
  double anc[3] = {0, 0, 0};
  if (bid >= 0) {
    const int off = a.bl.off[bid];
    LazyBlanket LB{a.c, a.bl.row + off, a.bl.p1 + off, a.bl.p2 + off, a.bl_kid + off, a.bl.off[bid + 1] - off, j,
                   (uint32_t)__ldg(a.c.cutraw + pidx), true, a.bl.counters};
    const int q0 = a.bl.q0[bid], q1 = a.bl.q1[bid];
    const int qme = side == 0 ? q0 : q1, qot = a.nf == 2 ? (side == 0 ? q1 : q0) : -1;
    const uint32_t tree = a.bl_tree[bid];                   // bit s: ancestry of focal s is a plain tree
    double tot_me = 1.0, tot_ot = 1.0;
    const int sme = a.nf == 2 ? side : 0;
    if ((flags & GFX_ITEM_OWN_NOPEEL) ||
        !slot_peel(a.c, a.bl_slotrow + ((int64_t)bid * 2 + sme) * 14, a.bl_slotbad[bid * 2 + sme], j, LB.cut, anc)) {
      LB.prefetch();
      if (!(flags & GFX_ITEM_OWN_NOX) && xpeel_query(LB, qme, qot, anc, &tot_me)) { GFX_COUNT(a.bl, 8, 1); }
      else if constexpr (MODE == 1) { if (fail_flag) *fail_flag = GFX_ITEM_OWN_NOX; return false; }
      else if (query_general(LB, a.bl, qme, qot, anc, &tot_me, stab)) atomicExch(a.status, 1);
    } else {
      GFX_COUNT(a.bl, 1, 1);
    }
    if (qot >= 0 && !((tree >> (side ^ 1)) & 1u)) {
      // the joint of correctPair is 0/0 when EITHER animal's component has P(evidence) = 0; a
      // tree-shaped ancestry has no evidence below unobserved nodes, so its total is exactly 1
      double tmp[3];
      if ((flags & GFX_ITEM_MATE_NOPEEL) ||
          !slot_peel(a.c, a.bl_slotrow + ((int64_t)bid * 2 + (side ^ 1)) * 14, a.bl_slotbad[bid * 2 + (side ^ 1)], j, LB.cut, tmp)) {
        if (!LB.pre) LB.prefetch();
        if ((flags & GFX_ITEM_MATE_NOX) || !xpeel_query(LB, qot, qme, tmp, &tot_ot)) {
          if constexpr (MODE == 1) { if (fail_flag) *fail_flag = GFX_ITEM_MATE_NOX; return false; }
          else if (query_general(LB, a.bl, qot, qme, tmp, &tot_ot, stab)) atomicExch(a.status, 1);
        }
      }
    }
    if (tot_me == 0.0 || tot_ot == 0.0) anc[0] = anc[1] = anc[2] = nan("");
  }
  const int k0 = a.child_off[row], nk = a.child_off[row + 1] - k0;
  if (bid < 0 && nk <= 0) return true;                      // lone node: no result (:287-288 / :417-419)
  bool has_kids;
  double kids[3];
  const bool known = kids_literal_known(nk, K, &has_kids, kids);
  if (!known && !literal && !a.post) {
    if (bid >= 0 && !(anc[0] == anc[0])) { *bits_out = 0x80u; return true; }   // NaN posterior: nothing passes a test
    const double n0 = bid >= 0 ? anc[0] * (double)K.k0 : (double)K.k0, n1 = bid >= 0 ? anc[1] * (double)K.k1 : (double)K.k1,
                 n2 = bid >= 0 ? anc[2] * (double)K.k2 : (double)K.k2;
    if ((n0 + n1) + n2 > 0.0) {
      const uint32_t b = guard_bits(n0, n1, n2, obs0, a.p.tiethreshold, a.p.threshold);
      if (b != GFX_BITS_GUARD) { *bits_out = b; return true; }
    } else {
      // every product has an exact zero factor: the literal posterior is the unnormalised zero vector (:280-304)
      const double z3[3] = {0.0, 0.0, 0.0};
      *bits_out = gfx_decision_bits(z3, obs0, a.p.tiethreshold, a.p.threshold);
      return true;
    }
  }
  if (!known) kids_term_dev(a.c, a.child_row + k0, a.child_other + k0, nk, j, a.kv, kids);
  double p[3];
  gfx_combine(bid >= 0, anc, has_kids, kids, p);
  *bits_out = gfx_decision_bits(p, obs0, a.p.tiethreshold, a.p.threshold);
  if (a.post) {
    double* o = a.post + ((int64_t)ent * a.n_snps + j) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
  }
  return true;
}

// The serial apply step of one job on the live state (:809-824 / :932-938) with the lazily evaluated
// rank gate (:733-737,817 / :928-933).
struct ApplyState {
  int cur, applied, excluded;
  bool gate_done, gate;
};
__device__ __forceinline__ void apply_bits(const FocalArgs& a, int row, int64_t j, uint32_t rxf, uint32_t bits, ApplyState& S) {
  if (!bits) return;
  const bool cond = S.cur == 3 || (!((bits >> S.cur) & 1u) && (bits & 8u));
  if (!cond) return;
  if (!S.gate_done) {
    double mr = 0.0;
    for (int q = 0; q < 2; ++q) {
      const int pr = a.prow[2 * row + q];
      if (pr >= 0) { const double ev = rx_E(a.c, cell_rx(a.c, pr, j)); if (ev > mr) mr = ev; }
    }
    S.gate = rx_E(a.c, rxf) * a.p.err_thresh >= mr;
    S.gate_done = true;
  }
  if (S.gate) {
    const uint32_t set = bits & 7u;
    S.cur = __popc(set) == 1 ? (__ffs(set) - 1) : 3;
    ++S.applied;
  } else {
    ++S.excluded;
  }
}
__device__ __forceinline__ void write_cell(const FocalArgs& a, int row, int64_t j, int obs0, const ApplyState& S) {
  if (S.cur != obs0) {
    uint32_t* wp = a.live + (int64_t)row * a.c.wpr + (j >> 4);
    const int sh = 2 * (int)(j & 15);
    atomicAnd(wp, ~(3u << sh));
    atomicOr(wp, (uint32_t)S.cur << sh);
  }
  if (S.applied) atomicAdd(a.tallies + row, S.applied);
  if (S.excluded && a.nf == 2) atomicAdd(a.tallies + a.n_rows + row, S.excluded);
}

// Plain peeling on the ancestor slots in integer arithmetic (the exact counterpart of slot_peel): al8 = 8 * P(node
// passes allele 0) is 8 - 4 c for an observed state c, 4 for an unobserved founder of the blanket and the mean of
// its parents' otherwise; anc64 = 64 * P(q) = [al_S al_D, al_S be_D + be_S al_D, be_S be_D].  The statuses of the
// focal animal's own parents (slots 0 / 1 of every blanket of this animal) come from the per-cell cache (rows pS / pD,
// codes, raw patterns): they differ between the jobs of a cell only through `cut`.
#ifndef GFX_FAST_ANC_EAGER
#define GFX_FAST_ANC_EAGER 0
#endif
struct ParentCache { int rS, rD; uint32_t cS, cD, xS, xD; };
__device__ __forceinline__ int slot_status_cached(const CellCtx& c, int r, const ParentCache& P, int64_t wo, int sh, int64_t j,
                                                  uint32_t cut) {
  if (r >= 0 && r == P.rS) return (P.cS == 3u || P.xS >= cut) ? 3 : (int)P.cS;
  if (r >= 0 && r == P.rD) return (P.cD == 3u || P.xD >= cut) ? 3 : (int)P.cD;
  return slot_status(c, r, wo, sh, j, cut);
}
__device__ __forceinline__ bool fast_anc(const CellCtx& c, const int32_t* __restrict__ srow, uint32_t bad, const ParentCache& P,
                                         int64_t wo, int sh, int64_t j, uint32_t cut, int anc64[3]) {
  if (bad & 0x8000u) return false;
  int st[2];
  st[0] = slot_status_cached(c, __ldg(srow + 0), P, wo, sh, j, cut);
  st[1] = slot_status_cached(c, __ldg(srow + 1), P, wo, sh, j, cut);
#if GFX_FAST_ANC_EAGER
  // the grand-parents' statuses are requested together with the parents' (one memory latency instead of two when a
  // parent is unobserved or suspect; wasted loads when both parents count)
  int st2[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) st2[g] = slot_status(c, __ldg(srow + 2 + g), wo, sh, j, cut);
#endif
  int al[2];
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (st[f] <= 2) { al[f] = 8 - 4 * st[f]; continue; }
    if ((bad >> f) & 1u) return false;
    int a2[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int s2 = 2 * f + 2 + g;
#if GFX_FAST_ANC_EAGER
      const int t2 = st2[2 * f + g];
#else
      const int t2 = slot_status(c, __ldg(srow + s2), wo, sh, j, cut);
#endif
      if (t2 <= 2) { a2[g] = 8 - 4 * t2; continue; }
      if (t2 == 4) { a2[g] = -1; continue; }            // no parents: f is a founder of the blanket
      if ((bad >> s2) & 1u) return false;
      const int u0 = slot_status(c, __ldg(srow + 2 * s2 + 2), wo, sh, j, cut);
      const int u1 = slot_status(c, __ldg(srow + 2 * s2 + 3), wo, sh, j, cut);
      if (u0 == 4) { a2[g] = 4; continue; }             // unobserved founder
      if ((u0 == 3 && ((bad >> (2 * s2 + 2)) & 1u)) || (u1 == 3 && ((bad >> (2 * s2 + 3)) & 1u))) return false;
      const int x0 = u0 <= 2 ? 8 - 4 * u0 : 4, x1 = u1 <= 2 ? 8 - 4 * u1 : 4;
      a2[g] = (x0 + x1) >> 1;                           // multiples of 2: exact
    }
    al[f] = a2[0] < 0 ? 4 : ((a2[0] + a2[1]) >> 1);     // both multiples of 2: exact
  }
  const int be0 = 8 - al[0], be1 = 8 - al[1];
  anc64[0] = al[0] * al[1];
  anc64[2] = be0 * be1;
  anc64[1] = al[0] * be1 + be0 * al[1];
  return true;
}

// Fast per-cell walk.  Jobs are evaluated with integer slot peeling + class counts and applied in order; from the first
// job that needs more (message passing, general elimination, or a literal replay inside the guard band), the cell
// gets a slot in the deferred list: its state is parked, the remaining easy jobs still deposit their bits there, the
// hard jobs become items for k_correct_jobs, and k_correct_apply finishes the walk.  Returns false if the whole cell
// must go to the fallback kernel (no slot left, posteriors requested), in which case nothing has been written.
__device__ __forceinline__ bool lean_cell(const FocalArgs& a, const uint2* s_ku, int f, int row, int e0, int e1, int64_t j) {
  const int64_t wo = j >> 4;
  const int sh = 2 * (int)(j & 15);
  const int obs0 = (int)((__ldg(a.c.live + (int64_t)row * a.c.wpr + wo) >> sh) & 3u);   // snapshot state of the focal animal
  const uint32_t rxf = cell_rx(a.c, row, j);
  const uint32_t idxf = rxf == 0xFFFFu ? 65536u : rxf;
  const uint32_t tp = a.c.tp_raw;
  GFX_COUNT(a.bl, 5, 1);
  // children term: the same for every job of this animal, counted once with the warp converged
  const int k0 = a.child_off[row], nk = a.child_off[row + 1] - k0;
  if (nk > GFX_KIDS_COUNT_MAX) return false;
  const KidCounts K = kids_counts(a.c, a.child_row + k0, a.child_other + k0, nk, wo, sh, s_ku);
  bool has_kids;
  double kids_known[3];
  const bool known = kids_literal_known(nk, K, &has_kids, kids_known);
  ParentCache P;
  P.rS = a.prow[2 * row]; P.rD = a.prow[2 * row + 1];
  P.cS = P.cD = 3u; P.xS = P.xD = 0u;
  if (P.rS >= 0) { P.cS = (__ldg(a.c.live + (int64_t)P.rS * a.c.wpr + wo) >> sh) & 3u; P.xS = cell_rx(a.c, P.rS, j); }
  if (P.rD >= 0) { P.cD = (__ldg(a.c.live + (int64_t)P.rD * a.c.wpr + wo) >> sh) & 3u; P.xD = cell_rx(a.c, P.rD, j); }
  const ParentCache none{-9, -9, 3u, 3u, 0u, 0u};
  ApplyState S{obs0, 0, 0, false, false};
  bool parked = false;
  unsigned long long slot = 0, sb = 0;
  int first = 0;
  for (int e = e0; e < e1; ++e) {
    const int ent = a.entry[e];
    const int job = a.nf == 2 ? (ent >> 1) : ent;
    const int side = a.nf == 2 ? (ent & 1) : 0;
    bool sus = rxf >= tp;                                     // :181-185 / :342
    uint32_t pidx;                                            // pattern of pmax (:220-228 / :374-376)
    if (a.nf == 2) {
      const int mrow = side == 0 ? a.job_row1[job] : a.job_row0[job];
      const uint32_t rxm = cell_rx(a.c, mrow, j);
      sus = sus || (rxm >= tp);
      const uint32_t idxm = rxm == 0xFFFFu ? 65536u : rxm;
      const int obs_m = (int)((__ldg(a.c.live + (int64_t)mrow * a.c.wpr + wo) >> sh) & 3u);
      if (obs0 != 3 && obs_m != 3) pidx = idxf > idxm ? idxf : idxm;
      else if (obs0 == 3 && obs_m == 3) pidx = 65536u;
      else pidx = obs0 != 3 ? idxf : idxm;
    } else {
      pidx = obs0 != 3 ? idxf : 65536u;
    }
    uint32_t bits = 0, flag = 0;
    bool ok = true;
    const int bid = a.job_bl[job];
    if (sus && !(bid < 0 && nk <= 0)) {                       // else: not suspect here, or a lone node without result
      GFX_COUNT(a.bl, 6, 1);
      int anc64[3] = {1, 1, 1};
      if (bid >= 0) {
        const uint32_t cut = __ldg(a.c.cutraw + pidx);
        const int sme = a.nf == 2 ? side : 0;
        ok = fast_anc(a.c, a.bl_slotrow + ((int64_t)bid * 2 + sme) * 14, a.bl_slotbad[bid * 2 + sme], P, wo, sh, j, cut, anc64);
        if (!ok) flag = GFX_ITEM_OWN_NOPEEL;
        if (ok && a.nf == 2 && !((a.bl_tree[bid] >> (side ^ 1)) & 1u)) {
          // the mate's component must have P(evidence) > 0 (:242-245 joint = False): certain when it peels plainly
          int tmp[3];
          ok = fast_anc(a.c, a.bl_slotrow + ((int64_t)bid * 2 + (side ^ 1)) * 14, a.bl_slotbad[bid * 2 + (side ^ 1)], none, wo, sh, j,
                        cut, tmp);
          if (!ok) flag = GFX_ITEM_MATE_NOPEEL;
        }
      }
      if (!ok) GFX_COUNT(a.bl, 8, 1);
      if (ok) {
        if (known) {
          GFX_COUNT(a.bl, 10, 1);
          // no children sum needed: the literal arithmetic of :280-304 on exact inputs
          const double anc[3] = {anc64[0] * (1.0 / 64.0), anc64[1] * (1.0 / 64.0), anc64[2] * (1.0 / 64.0)};
          double p[3];
          gfx_combine(bid >= 0, anc, has_kids, kids_known, p);
          bits = gfx_decision_bits(p, obs0, a.p.tiethreshold, a.p.threshold);
        } else {
          const int n0 = anc64[0] * K.k0, n1 = anc64[1] * K.k1, n2 = anc64[2] * K.k2;
          if (n0 + n1 + n2 > 0) {
            bits = guard_bits((double)n0, (double)n1, (double)n2, obs0, a.p.tiethreshold, a.p.threshold);
            if (bits == GFX_BITS_GUARD) { ok = false; flag = GFX_ITEM_LITERAL; GFX_COUNT(a.bl, 9, 1); }
          } else {
            const double z3[3] = {0.0, 0.0, 0.0};   // exact zeros: the unnormalised zero vector (:280-304)
            bits = gfx_decision_bits(z3, obs0, a.p.tiethreshold, a.p.threshold);
          }
        }
      }
    }
    if (ok && !parked) { apply_bits(a, row, j, rxf, bits, S); continue; }
    if (!parked) {
      // one byte of decision bits per job of this animal (a many-mate sire has dozens, a dam one), then the cell's slot
      // (one atomic for both: slot number in the low word of the counter, bytes handed out in the high word)
      // The lanes of the warp that park at this job draw together (same animal, so the same e1 - e0): one same-address
      // atomic with a returned value per lane was 8 % of the kernel's stall samples (profiles/r2_stalls_k_correct_gen0.txt).
      const uint32_t act = __activemask();
      const int ln = threadIdx.x & 31, leader = __ffs(act) - 1, nact = __popc(act);
      unsigned long long got = 0;
      if (ln == leader)
        got = atomicAdd(a.wl_count, ((unsigned long long)(e1 - e0) * (unsigned long long)nact << 32) | (unsigned long long)nact);
      got = __shfl_sync(act, got, leader);
      const unsigned long long rk = (unsigned long long)__popc(act & ((1u << ln) - 1u));
      slot = (got & 0xFFFFFFFFull) + rk;
      sb = (got >> 32) + rk * (unsigned long long)(e1 - e0);
      if (slot >= a.wl_cap || sb + (unsigned long long)(e1 - e0) > a.sb_cap) return false;
      parked = true;
      first = e - e0;
    }
    if (ok) { a.sbits[sb + (e - e0)] = (uint8_t)bits; continue; }
    a.sbits[sb + (e - e0)] = 0xFF;
    const unsigned long long it = atomicAdd(a.item_count, 1ull);
    if (it < a.item_cap) {
      a.items[it] = make_uint2((uint32_t)slot, (uint32_t)(e - e0) | flag);
      atomicAdd(a.key_count + (e - a.e_base), 1u);
    }
    else first |= 0x8000;   // no room for the job item: the cell is redone by the fallback kernel
  }
  if (!parked) { write_cell(a, row, j, obs0, S); return true; }
  a.wl[slot] = make_uint2((uint32_t)f, (uint32_t)j);
  a.st[slot] = make_uint2((uint32_t)S.cur | ((uint32_t)(first & 0x7FFF) << 8) | ((first & 0x8000) ? 0x80000000u : 0u) |
                              (S.gate_done ? 0x40000000u : 0u) | (S.gate ? 0x20000000u : 0u),
                          (uint32_t)S.applied | ((uint32_t)S.excluded << 16));
  a.skc[slot] = make_int4(K.k0, K.k1, K.k2, K.ninc);
  a.sboff[slot] = (uint32_t)sb;
  return (first & 0x8000) == 0;
}

template <bool GENERAL>
__device__ __forceinline__ bool correct_cell(const FocalArgs& a, const double* s_kv, uint32_t* s_cls, int row, int e0, int e1,
                                             int64_t j) {
  double* const stab = nullptr;   // the whole-cell fallback keeps its tables private

  const uint32_t tp = a.c.tp_raw;
  const int obs0 = cell_code(a.c.live, a.c.wpr, row, j);   // snapshot state of the focal animal
  int cur = obs0;                                          // live state, updated job after job
  const uint32_t rxf = cell_rx(a.c, row, j);
  const uint32_t idxf = rxf == 0xFFFFu ? 65536u : rxf;
  bool gate_done = false, gate = false, has_kids = false;
  double kids[3] = {0, 0, 0};
  int applied = 0, excluded = 0;
  GFX_COUNT(a.bl, 5, 1);
  {
    // children term: the same for every job of this animal, evaluated once with the warp converged
    const int k0 = a.child_off[row], nk = a.child_off[row + 1] - k0;
    if (nk > 0 && nk <= GFX_KIDS_FAST) {
      kids_term_fast(a.c, a.child_row + k0, a.child_other + k0, nk, j, s_kv, s_cls, kids);
      has_kids = true;
    } else if (nk > 0) {
      if constexpr (!GENERAL) return false;
      else has_kids = kids_term_dev(a.c, a.child_row + k0, a.child_other + k0, nk, j, a.kv, kids);
    }
    GFX_COUNT(a.bl, 7, nk);
  }
  for (int e = e0; e < e1; ++e) {
    const int ent = a.entry[e];
    const int job = a.nf == 2 ? (ent >> 1) : ent;
    const int side = a.nf == 2 ? (ent & 1) : 0;
    bool sus = rxf >= tp;
    uint32_t pidx;                                          // pattern of pmax (:220-228 / :374-376)
    if (a.nf == 2) {
      const int mrow = side == 0 ? a.job_row1[job] : a.job_row0[job];
      const uint32_t rxm = cell_rx(a.c, mrow, j);
      sus = sus || (rxm >= tp);
      if (!sus) continue;
      const uint32_t idxm = rxm == 0xFFFFu ? 65536u : rxm;
      const int obs_m = cell_code(a.c.live, a.c.wpr, mrow, j);
      if (obs0 != 3 && obs_m != 3) pidx = idxf > idxm ? idxf : idxm;
      else if (obs0 == 3 && obs_m == 3) pidx = 65536u;
      else pidx = obs0 != 3 ? idxf : idxm;
    } else {
      if (!sus) continue;
      pidx = obs0 != 3 ? idxf : 65536u;
    }
    GFX_COUNT(a.bl, 6, 1);
    const int bid = a.job_bl[job];
    double anc[3] = {0, 0, 0};
    if (bid >= 0) {
      const int off = a.bl.off[bid];
      LazyBlanket LB{a.c, a.bl.row + off, a.bl.p1 + off, a.bl.p2 + off, a.bl_kid + off, a.bl.off[bid + 1] - off, j,
                     (uint32_t)__ldg(a.c.cutraw + pidx), true, a.bl.counters};
      const int q0 = a.bl.q0[bid], q1 = a.bl.q1[bid];
      const int qme = side == 0 ? q0 : q1, qot = a.nf == 2 ? (side == 0 ? q1 : q0) : -1;
      const uint32_t tree = a.bl_tree[bid];                 // bit s: ancestry of focal s is a plain tree
      double tot_me = 1.0, tot_ot = 1.0;
      const int sme = a.nf == 2 ? side : 0;
      if (!slot_peel(a.c, a.bl_slotrow + ((int64_t)bid * 2 + sme) * 14, a.bl_slotbad[bid * 2 + sme], j, LB.cut, anc)) {
        if (xpeel_query(LB, qme, qot, anc, &tot_me)) { GFX_COUNT(a.bl, 8, 1); }
        else if constexpr (!GENERAL) return false;
        else if (query_general(LB, a.bl, qme, qot, anc, &tot_me, stab)) atomicExch(a.status, 1);
      } else {
        GFX_COUNT(a.bl, 1, 1);
      }
      if (qot >= 0 && !((tree >> (side ^ 1)) & 1u)) {
        // the joint of correctPair is 0/0 when EITHER animal's component has P(evidence) = 0; a
        // tree-shaped ancestry has no evidence below unobserved nodes, so its total is exactly 1
        double tmp[3];
        if (!slot_peel(a.c, a.bl_slotrow + ((int64_t)bid * 2 + (side ^ 1)) * 14, a.bl_slotbad[bid * 2 + (side ^ 1)], j, LB.cut, tmp) &&
            !xpeel_query(LB, qot, qme, tmp, &tot_ot)) {
          if constexpr (!GENERAL) return false;
          else if (query_general(LB, a.bl, qot, qme, tmp, &tot_ot, stab)) atomicExch(a.status, 1);
        }
      }
      if (tot_me == 0.0 || tot_ot == 0.0) anc[0] = anc[1] = anc[2] = nan("");
    }
    if (bid < 0 && !has_kids) continue;                     // lone node: no result (:287-288 / :417-419)
    double p[3];
    gfx_combine(bid >= 0, anc, has_kids, kids, p);
    const uint32_t bits = gfx_decision_bits(p, obs0, a.p.tiethreshold, a.p.threshold);
    if (a.post) {
      double* o = a.post + ((int64_t)ent * a.n_snps + j) * 3;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    }
    // decision on the LIVE state (:809-824 / :932-938)
    const bool cond = cur == 3 || (!((bits >> cur) & 1u) && (bits & 8u));
    if (!cond) continue;
    if (!gate_done) {                                       // rank gate (:733-737,817 / :928-933)
      double mr = 0.0;
      for (int q = 0; q < 2; ++q) {
        const int pr = a.prow[2 * row + q];
        if (pr >= 0) { const double ev = rx_E(a.c, cell_rx(a.c, pr, j)); if (ev > mr) mr = ev; }
      }
      gate = rx_E(a.c, rxf) * a.p.err_thresh >= mr;
      gate_done = true;
    }
    if (gate) {
      const uint32_t set = bits & 7u;
      cur = __popc(set) == 1 ? (__ffs(set) - 1) : 3;
      ++applied;
    } else {
      ++excluded;
    }
  }
  if (cur != obs0) {
    uint32_t* wp = a.live + (int64_t)row * a.c.wpr + (j >> 4);
    const int sh = 2 * (int)(j & 15);
    atomicAnd(wp, ~(3u << sh));
    atomicOr(wp, (uint32_t)cur << sh);
  }
  if (applied) atomicAdd(a.tallies + row, applied);
  if (excluded && a.nf == 2) atomicAdd(a.tallies + a.n_rows + row, excluded);
  return true;
}

// Lean kernel.  A warp takes (focal animal, GFX_LEAN_CHUNK-SNP chunk) tasks from a global counter -- animals are
// ordered heaviest first, so the many-mate sires do not end up as the tail --, compacts the suspect
// SNPs of the chunk with ballots and then runs one cell per lane.  Cells that need the general path
// go to the work list (or, if that is full, the overflow bitmap) untouched.
// SNPs per warp task of the pair generations (measured on config 2: 256 -> 16.8 ms, 128 -> 16.3, 64 -> 15.1, 32 -> 14.8 ms per
// chunk: short tasks balance better and the suspect cells are dense enough that one ballot round fills the warp).
// The singles are the opposite case: thousands of childless animals with ~1 % suspect cells, so a 32-SNP task finds
// 0.4 cells and runs the cell walk with one lane (266 M warp instructions for 336 k job evaluations at 10 k x 4096);
// their tasks cover 1024 SNPs and compact ~12 cells per walk.
#ifndef GFX_LEAN_CHUNK
#define GFX_LEAN_CHUNK 32
#endif
#ifndef GFX_LEAN_CHUNK_SINGLES
#define GFX_LEAN_CHUNK_SINGLES 1024
#endif
#ifndef KC_BLOCKS
#define KC_BLOCKS 8   // 64 registers (44 bytes of spills) and 32 warps per SM: 8.49 ms per chunk alone against 8.81 with 6 blocks
#endif                // (78 registers), 39.2 against 40.1 ms per 3-lane bench step; 7 blocks: 8.61 / 38.8 ms

template <int GFX_LEAN_CHUNK_T>
__global__ void __launch_bounds__(128, KC_BLOCKS) k_correct(FocalArgs a) {
  __shared__ uint16_t s_list[4][GFX_LEAN_CHUNK_T];
  __shared__ uint2 s_ku[16];
  kid_units_init(s_ku);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t chunks = (a.n_snps + GFX_LEAN_CHUNK_T - 1) / GFX_LEAN_CHUNK_T;
  const int64_t tasks = (int64_t)a.n_focal * chunks;
  const uint32_t tp = a.c.tp_raw;
  for (;;) {
    // one (animal, SNP range) task per draw, drawn when needed.  [Measured and rejected: requesting the next ticket before
    // the current task is worked on (10.43 vs 10.45 ms per chunk), draws of 2 / 4 consecutive tasks (+0.1 / +0.3 ms).]
    unsigned long long task = 0;
    if (lane == 0) task = atomicAdd(a.task_count, 1ull);
    task = __shfl_sync(0xffffffffu, task, 0);
    if ((int64_t)task >= tasks) break;
    // animal-major task order, heaviest animals first: the long tasks of the many-mate sires start at once
    // (measured: chunk-major order is 40 % slower because the last chunk's heavy tasks start late)
    const int f = a.focal_order[task / chunks];
    const int64_t j0 = (int64_t)(task % chunks) * GFX_LEAN_CHUNK_T;
    const int row = a.focal_row[f];
    const int e0 = a.focal_off[f], e1 = a.focal_off[f + 1];
    // pass 1: SNPs where at least one job of this animal is suspect (:181-185 / :342); E >= test_p is a
    // comparison of raw patterns because E is monotone in the score
    int cnt = 0;
    for (int t = 0; t < GFX_LEAN_CHUNK_T / 32; ++t) {
      if (j0 + t * 32 >= a.n_snps) break;
      const int64_t j = j0 + t * 32 + lane;
      bool sus = false;
      if (j < a.n_snps) {
        sus = cell_rx(a.c, row, j) >= tp;
        if (a.nf == 2)
          // [measured and rejected: the mates' rows kept in shared memory per task and four mates' scores requested
          //  together: 10.60 vs 10.45 ms per chunk]
          for (int e = e0; e < e1 && !sus; ++e) {
            const int ent = a.entry[e];
            const int mrow = (ent & 1) == 0 ? a.job_row1[ent >> 1] : a.job_row0[ent >> 1];
            sus = cell_rx(a.c, mrow, j) >= tp;
          }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, sus);
      if (sus) s_list[wib][cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(t * 32 + lane);
      cnt += __popc(m);
    }
    __syncwarp();
    for (int k0 = 0; k0 < cnt; k0 += 32) {
      const int k = k0 + lane;
      int64_t j = 0;
      bool defer = false;
      if (k < cnt) {
        j = j0 + s_list[wib][k];
        // posteriors requested: every cell takes the literal whole-cell kernel (the reference's own arithmetic)
        defer = a.post != nullptr || !lean_cell(a, s_ku, f, row, e0, e1, j);
      }
      // (rare) whole cell to the fallback kernel through the bitmap
      if (defer) {
        atomicOr(a.deferred + (int64_t)f * a.dwords + (j >> 5), 1u << (j & 31));
        atomicAdd(a.wl_count + 3, 1ull);
      }
    }
    __syncwarp();
  }
}

// Deferred jobs are grouped by job entry before they run: the threads of a warp then walk the same blanket from
// the same query and differ only in the evidence of their SNP (measured before: 3 of 32 lanes active).
__global__ void __launch_bounds__(1024) k_jobs_scan(FocalArgs a) {
  __shared__ uint32_t s_part[1024];
  const int t = threadIdx.x;
  const int per = (a.n_keys + 1023) / 1024;
  uint32_t sum = 0;
  for (int i = t * per; i < min((t + 1) * per, a.n_keys); ++i) sum += a.key_count[i];
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = t >= o ? s_part[t - o] : 0u;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  uint32_t run = s_part[t] - sum;
  for (int i = t * per; i < min((t + 1) * per, a.n_keys); ++i) {
    a.key_off[i] = run;
    a.key_cur[i] = 0u;
    run += a.key_count[i];
  }
}
__global__ void __launch_bounds__(256) k_jobs_scatter(FocalArgs a) {
  unsigned long long n = *a.item_count;
  if (n > a.item_cap) n = a.item_cap;
  unsigned long long ncell = *a.wl_count & 0xFFFFFFFFull;
  if (ncell > a.wl_cap) ncell = a.wl_cap;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 it = a.items[i];
    // items of cells that found no slot were never counted; they keep their cell on the fallback path
    if (it.x >= ncell) continue;
    const int key = a.focal_off[a.wl[it.x].x] + (int)(it.y & ~GFX_ITEM_FLAGS) - a.e_base;
    const uint32_t pos = a.key_off[key] + atomicAdd(a.key_cur + key, 1u);
    a.items_sorted[pos] = it;
  }
}

// Deferred jobs, one thread each, in two rounds so that a warp never mixes the cheap and the expensive kind:
// round A = tree message passing; what it cannot answer is handed to round B = general elimination.
// Warps draw 32 items at a time from a global counter (the general eliminations have a long tail).
template <int MODE, int PER>
__device__ __forceinline__ void jobs_round(const FocalArgs& a, const uint2* list, unsigned long long n, unsigned long long* ticket,
                                           double* stab = nullptr) {
  unsigned long long ncell = *a.wl_count & 0xFFFFFFFFull;
  if (ncell > a.wl_cap) ncell = a.wl_cap;
  const int lane = threadIdx.x & 31;
  for (;;) {
    // [measured and rejected: the next ticket requested before these items are worked on -- a busy warp then sits on a
    //  batch that an idle one would have taken: 11.12 vs 10.45 ms per chunk]
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t * (unsigned long long)PER >= n) break;
    const unsigned long long i = t * (unsigned long long)PER + lane;
    bool again = false;
    uint2 it = make_uint2(0u, 0u);
    if (lane < PER && i < n) {
      it = list[i];
      if (it.x < ncell) {
        const uint2 c = a.wl[it.x];
        const int f = (int)c.x;
        const int64_t j = (int64_t)c.y;
        const int row = a.focal_row[f];
        const int obs0 = cell_code(a.c.live, a.c.wpr, row, j);
        const uint32_t rxf = cell_rx(a.c, row, j);
        const int4 kc = a.skc[it.x];
        const KidCounts K{kc.x, kc.y, kc.z, kc.w};
        const int ji = (int)(it.y & ~GFX_ITEM_FLAGS);
        uint32_t bits, why = 0u;
        if (job_bits<MODE>(a, row, a.entry[a.focal_off[f] + ji], j, obs0, rxf, K, it.y & GFX_ITEM_FLAGS, &bits, stab, &why)) {
          a.sbits[(unsigned long long)a.sboff[it.x] + ji] = (uint8_t)bits;
        } else {
          again = true;
          // what the general round need not try again; when the focal side was answered by message passing here, its
          // plain peeling is known to fail as well
          it.y |= why | (why == GFX_ITEM_MATE_NOX ? GFX_ITEM_MATE_NOPEEL : GFX_ITEM_OWN_NOPEEL);
        }
      }
    }
    if (MODE == 1) {
      // hand over to round B (a.items is free again: the scatter pass has copied it)
      const uint32_t m = __ballot_sync(0xffffffffu, again);
      if (m) {
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(a.wl_count + 5, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (again) a.items[base + __popc(m & ((1u << lane) - 1u))] = it;
      }
    }
  }
}
#ifndef KJ_BLOCKS
#define KJ_BLOCKS 5   // 96 registers, no spills, 20 warps per SM: 36.9 against 38.5 ms per 3-lane bench step with 4 blocks (128 registers)
#endif
#ifndef KJ_ITEMS
#define KJ_ITEMS 32   // jobs per ticket of the tree-message-passing round
#endif
// Threads per block of the two deferred-job kernels.  One-warp blocks (so that the long tail of a few slow items
// does not pin the registers of whole 4-warp blocks next to the other lanes' kernels) were measured: no gain in the
// 3-lane bench step (47.7 vs 48.3 ms), slower alone -- profiles/r2_summary.md.
#ifndef KJ_THREADS
#define KJ_THREADS 128
#endif
__global__ void __launch_bounds__(KJ_THREADS, KJ_BLOCKS * 128 / KJ_THREADS) k_correct_jobs(FocalArgs a) {
  unsigned long long n = *a.item_count;
  if (n > a.item_cap) n = a.item_cap;
  jobs_round<1, KJ_ITEMS>(a, a.items_sorted, n, a.wl_count + 4);
}
#ifndef KG_BLOCKS
#define KG_BLOCKS 4
#endif
#ifndef KG_ITEMS
#define KG_ITEMS 32   // jobs per ticket of the general round (with the one-pass table products: 32 -> 9.03 ms per chunk, 16 -> 9.2, 8 -> 9.75)
#endif
__global__ void __launch_bounds__(KJ_THREADS, KG_BLOCKS * 128 / KJ_THREADS) k_correct_jobs_general(FocalArgs a) {
#ifdef GFX_GENERAL_SMEM_TIER
  // one shared-memory table slot per job in flight: warp w, lane l < KG_ITEMS -> slot w * KG_ITEMS + l
  __shared__ double s_tab[GFX_TAB_S * GFX_STAB_SLOTS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* stab = (lane < KG_ITEMS && wib * KG_ITEMS + lane < GFX_STAB_SLOTS) ? s_tab + wib * KG_ITEMS + lane : nullptr;
  jobs_round<2, KG_ITEMS>(a, a.items, a.wl_count[5], a.wl_count + 6, stab);
#else
  jobs_round<2, KG_ITEMS>(a, a.items, a.wl_count[5], a.wl_count + 6, nullptr);
#endif
}

// Finishes the serial walk of every parked cell once all its jobs have bits.
__global__ void __launch_bounds__(128) k_correct_apply(FocalArgs a) {
  unsigned long long n = *a.wl_count & 0xFFFFFFFFull;
  if (n > a.wl_cap) n = a.wl_cap;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 st = a.st[i];
    if (st.x & 0x80000000u) continue;   // redone by the fallback kernel
    const uint2 c = a.wl[i];
    const int f = (int)c.x;
    const int64_t j = (int64_t)c.y;
    const int row = a.focal_row[f];
    const int e0 = a.focal_off[f], e1 = a.focal_off[f + 1];
    const int obs0 = cell_code(a.c.live, a.c.wpr, row, j);
    const uint32_t rxf = cell_rx(a.c, row, j);
    ApplyState S{(int)(st.x & 0xFFu), (int)(st.y & 0xFFFFu), (int)(st.y >> 16), (st.x & 0x40000000u) != 0,
                 (st.x & 0x20000000u) != 0};
    const int first = (int)((st.x >> 8) & 0x7FFFu);
    const unsigned long long sb = a.sboff[i];
    for (int e = e0 + first; e < e1; ++e) apply_bits(a, row, j, rxf, a.sbits[sb + (e - e0)], S);
    write_cell(a, row, j, obs0, S);
  }
}

// Only when the work list overflowed: the remaining deferred cells from the bitmap.
__global__ void __launch_bounds__(128) k_correct_general_overflow(FocalArgs a) {
  __shared__ int s_list[GFX_CORRECT_CHUNK];
  __shared__ int s_cnt;
  __shared__ double s_kv[108];
  __shared__ uint32_t s_cls[GFX_KIDS_WORDS * 128];
  if (a.wl_count[3] == 0) return;   // no cell was sent to the fallback
  for (int i = threadIdx.x; i < 108; i += blockDim.x) s_kv[i] = a.kv[i];
  const int64_t chunks = (a.n_snps + GFX_CORRECT_CHUNK - 1) / GFX_CORRECT_CHUNK;
  const int64_t tasks = (int64_t)a.n_focal * chunks;
  for (int64_t task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int f = (int)(task / chunks);
    const int64_t j0 = (task - (int64_t)f * chunks) * GFX_CORRECT_CHUNK;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (threadIdx.x < GFX_CORRECT_CHUNK / 32) {
      uint32_t m = a.deferred[(int64_t)f * a.dwords + (j0 >> 5) + threadIdx.x];
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        s_list[atomicAdd(&s_cnt, 1)] = threadIdx.x * 32 + b;
      }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt) {
      const int row = a.focal_row[f];
      const int e0 = a.focal_off[f], e1 = a.focal_off[f + 1];
      for (int k = threadIdx.x; k < cnt; k += blockDim.x)
        correct_cell<true>(a, s_kv, s_cls + threadIdx.x, row, e0, e1, j0 + s_list[k]);
    }
    __syncthreads();
  }
}

static int run_correct(const gfx_schedule* s, int nf, int n_jobs, const int32_t* job_row0, const int32_t* job_row1,
                       const int32_t* job_bl, int n_focal, const int32_t* focal_row, const int32_t* focal_off,
                       const int32_t* focal_order, const int32_t* entry, uint32_t* live, int64_t n_snps, int64_t wpr,
                       const uint16_t* raw, int64_t ld_raw, const double* elut, const gfx_correct_params* p,
                       uint32_t* snapshot, int64_t scratch_words, double* post, int32_t* tallies, int maxj, int e_base, int n_keys,
                       cudaStream_t stream) {
  if (n_jobs <= 0 || n_focal <= 0) return GFX_OK;
  const int64_t dwords = ((n_snps + GFX_CORRECT_CHUNK - 1) / GFX_CORRECT_CHUNK) * (GFX_CORRECT_CHUNK / 32);
  const int64_t fixed = (int64_t)s->h.n_rows * wpr + (int64_t)s->h.n_rows * dwords + 4;
  if (scratch_words < fixed + 64 * 1024)
    return fail(GFX_E_INVALID, "scratch too small: need n_rows * (wpr + 32 * ceil(n_snps / 1024)) words plus >= 256 KB of lists");
  GFX_CUDA(cudaMemcpyAsync(snapshot, live, sizeof(uint32_t) * (size_t)s->h.n_rows * wpr, cudaMemcpyDeviceToDevice, stream));
  FocalArgs a;
  a.c = CellCtx{snapshot, wpr, raw, ld_raw, elut, p->cutraw_dev, p->tp_raw};
  a.live = live;
  a.n_snps = n_snps; a.nf = nf; a.n_focal = n_focal;
  a.focal_row = focal_row; a.focal_off = focal_off; a.entry = entry;
  a.job_row0 = job_row0; a.job_row1 = job_row1; a.job_bl = job_bl;
  a.bl = BlanketDev{s->bl_off, s->bl_row, s->bl_p1, s->bl_p2, s->bl_last, s->bl_q0, s->bl_q1, s->big_pool, s->big_flags, s->huge_pool, s->debug ? s->counters : nullptr};
  a.bl_kid = s->bl_kid;
  a.bl_tree = s->bl_tree;
  a.bl_slotrow = s->bl_slotrow;
  a.bl_slotbad = s->bl_slotbad;
  a.child_off = s->child_off; a.child_row = s->child_row; a.child_other = s->child_other;
  a.prow = s->prow; a.kv = s->kv; a.p = *p; a.post = post; a.tallies = tallies; a.n_rows = s->h.n_rows; a.status = s->status;
  // scratch layout: [snapshot: n_rows * wpr][fallback bitmap: n_rows * dwords][4 counters][parked cells: children
  // terms, (focal, snp), state, job items, decision bits]
  a.dwords = dwords;
  a.deferred = snapshot + (size_t)s->h.n_rows * wpr;
  uint32_t* tail = a.deferred + (size_t)s->h.n_rows * dwords;
  tail += (4 - (((uintptr_t)tail / 4) & 3)) & 3;   // 16-byte alignment
  a.wl_count = (unsigned long long*)tail;
  a.item_count = a.wl_count + 2;
  (void)maxj;
  a.n_keys = n_keys; a.e_base = e_base;
  const int64_t nk4 = ((int64_t)n_keys + 3) / 4 * 4;
  a.key_count = tail + 16;            // zeroed with the 8 counters
  a.key_off = a.key_count + nk4;
  a.key_cur = a.key_off + nk4;
  char* base = (char*)(a.key_cur + nk4);
  const int64_t avail = (char*)(snapshot + scratch_words) - base;
  const int64_t per_slot = 16 + 8 + 8 + 4 + 16 + 16 + 12;   // class counts, cell, state, sbits offset, items x 2, 12 decision bytes
  const int64_t scap = avail > 0 ? avail / per_slot : 0;
  if (scap < 1024) return fail(GFX_E_INVALID, "scratch too small for the deferred-cell lists");
  a.skc = (int4*)base;                           base += scap * 16;
  a.wl = (uint2*)base;                           base += scap * 8;
  a.st = (uint2*)base;                           base += scap * 8;
  a.items = (uint2*)base;                        base += scap * 16;
  a.items_sorted = (uint2*)base;                 base += scap * 16;
  a.sboff = (uint32_t*)base;                     base += scap * 4;
  a.sbits = (uint8_t*)base;
  a.sb_cap = (unsigned long long)scap * 12ull < 0x7FFFFFF0ull ? (unsigned long long)scap * 12ull : 0x7FFFFFF0ull;   // 31 bits: the
  // high word of the slot counter keeps counting after the lists are full
  a.wl_cap = (unsigned long long)scap;
  a.item_cap = (unsigned long long)scap * 2;
  a.maxj = 0;
  GFX_CUDA(cudaMemsetAsync(a.deferred, 0, ((char*)(a.key_count + nk4)) - (char*)a.deferred, stream));  // bitmap + counters + key counts
  a.task_count = a.wl_count + 1;
  a.focal_order = focal_order;
  const int64_t tasks = (int64_t)n_focal * ((n_snps + GFX_CORRECT_CHUNK - 1) / GFX_CORRECT_CHUNK);
  // the fallback kernel returns at once unless a cell was sent to it (lists full, posteriors requested): a small grid,
  // its blocks stride over the tasks [24 blocks per SM cost 34 us per generation for nothing]
  int64_t blocks = tasks;
  const int64_t cap = (int64_t)s->num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int lean_chunk = nf == 2 ? GFX_LEAN_CHUNK : GFX_LEAN_CHUNK_SINGLES;
  const int64_t wtasks = (int64_t)n_focal * ((n_snps + lean_chunk - 1) / lean_chunk);
  int64_t lblocks = (wtasks + 3) / 4;
  if (lblocks > (int64_t)s->num_sms * KC_BLOCKS) lblocks = (int64_t)s->num_sms * KC_BLOCKS;
  if (nf == 2) k_correct<GFX_LEAN_CHUNK><<<(int)lblocks, 128, 0, stream>>>(a);
  else k_correct<GFX_LEAN_CHUNK_SINGLES><<<(int)lblocks, 128, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_jobs_scan<<<1, 1024, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_jobs_scatter<<<s->num_sms * 4, 256, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_correct_jobs<<<s->num_sms * 8 * (128 / KJ_THREADS), KJ_THREADS, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_correct_jobs_general<<<s->num_sms * 8 * (128 / KJ_THREADS), KJ_THREADS, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_correct_apply<<<s->num_sms * 4, 128, 0, stream>>>(a);
  GFX_LAUNCHED();
  k_correct_general_overflow<<<(int)blocks, 128, 0, stream>>>(a);
  GFX_LAUNCHED();
  return GFX_OK;
}

extern "C" int gfx_correct_pairs(const gfx_schedule* s, int32_t generation, uint32_t* live,
                                 int64_t n_snps, int64_t wpr, const uint16_t* raw, int64_t ld_raw, const double* elut,
                                 const gfx_correct_params* p, uint32_t* snapshot, int64_t scratch_words, double* post,
                                 int32_t* tallies, void* stream) {
  if (!s || !live || !raw || !elut || !p || !p->cutraw_dev || !snapshot || !tallies) return fail(GFX_E_INVALID, "null argument");
  if (s->device < 0) return fail(GFX_E_INVALID, "host-only schedule passed to a kernel entry point");
  if (generation < 0 || generation >= s->h.info.n_generations) return fail(GFX_E_INVALID, "generation out of range");
  const int j0 = s->h.gen_off[generation], nj = s->h.gen_off[generation + 1] - j0;
  const int f0 = s->h.app_gen_off[generation], nfoc = s->h.app_gen_off[generation + 1] - f0;
  int maxj = 1;
  for (int i = 0; i < nfoc; ++i)
    maxj = std::max(maxj, s->h.app_focal_off[f0 + i + 1] - s->h.app_focal_off[f0 + i]);
  // focal_off holds global offsets into app_entry: pass the un-shifted entry array
  return run_correct(s, 2, nj, s->job_row0 + j0, s->job_row1 + j0, s->job_bl + j0, nfoc, s->app_focal_row + f0,
                     s->app_focal_off + f0, s->app_order + f0, s->app_entry, live, n_snps, wpr, raw, ld_raw, elut, p, snapshot, scratch_words,
                     post, tallies, maxj, s->h.app_focal_off[f0], s->h.app_focal_off[f0 + nfoc] - s->h.app_focal_off[f0], (cudaStream_t)stream);
}

extern "C" int gfx_correct_singles(const gfx_schedule* s, uint32_t* live, int64_t n_snps, int64_t wpr,
                                   const uint16_t* raw, int64_t ld_raw, const double* elut, const gfx_correct_params* p,
                                   uint32_t* snapshot, int64_t scratch_words, double* post, int32_t* tallies, void* stream) {
  if (!s || !live || !raw || !elut || !p || !p->cutraw_dev || !snapshot || !tallies) return fail(GFX_E_INVALID, "null argument");
  if (s->device < 0) return fail(GFX_E_INVALID, "host-only schedule passed to a kernel entry point");
  const int n = (int)s->h.single_row.size();
  return run_correct(s, 1, n, s->single_row, nullptr, s->single_bl, n, s->single_row, s->single_off, s->single_order,
                     s->single_entry,
                     live, n_snps, wpr, raw, ld_raw, elut, p, snapshot, scratch_words, post, tallies, 1, 0, n, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// trio consistency scan: bit-parallel over 16 SNPs per word
// ------------------------------------------------------------------------------------------------
// One warp per FAMILY (trios with the same sire and dam are adjacent in trio_row, fam_off delimits them): per
// 16-byte quad the lane loads the sire's and the dam's codes once and then the codes of up to 4 kids, whose tallies
// live in registers, so the parents cross the L2 -> SM link once per family instead of once per kid (the per-trio
// form moved 3 rows per kid and was bound by that link: 7 TB/s of L2 reads for 2.6 TB/s of algorithmic bytes).
// Families are walked in groups of 4 kids (more per group costs occupancy).  No atomics: a kid belongs to one family.
// Round 2 tried a TILE-MAJOR task order (all families of a 2048-SNP column tile, then the next tile) so that a dam's
// second read -- as a parent, about one generation of rows after her read as a kid -- would hit L2: DRAM read stayed at
// 1.29x the algorithmic bytes (4.82 GB vs 4.85 GB for 3.75 GB at 300k animals) while the per-task index chain and the
// per-(tile, kid) reductions doubled the instructions: 0.58 -> 0.35 of the roofline; not kept (profiles/r2_summary.md).
#ifndef GFX_FAM_BLOCKS
#define GFX_FAM_BLOCKS 4
#endif
// One 16-byte quad (64 SNPs) of a family: the parents' four words against the four words of each of NK kids.  Both scan
// kernels are bound by the ALU pipe (LOP3 / SHF issue every second cycle per scheduler), not by memory, so the quad is
// written for few logic operations: the masks of the kid states the parents exclude (no0 / no1 / no2) are pre-masked with
// "both parents observed and the SNP exists", which makes a kid's inconsistency bit three LOP3 (a missing kid matches no
// state) and lets `bad & present` go; the bits of two words are merged (even | odd << 1, an IMAD) before each POPC; the
// ragged tail of a row is a separate instantiation, the full quads carry no validity arithmetic.
template <int NK, bool RAGGED>
__device__ __forceinline__ void trio_quad(const uint4& sq, const uint4& dq, const uint4 (&kq)[NK], int64_t first_snp, int64_t n_snps,
                                          int (&n_inc)[NK], int (&n_tst)[NK]) {
  const uint32_t M = 0x55555555u;
  const uint32_t sw[4] = {sq.x, sq.y, sq.z, sq.w}, dw[4] = {dq.x, dq.y, dq.z, dq.w};
  uint32_t b_even[NK], p_even[NK];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = sw[i], d = dw[i], s1 = s >> 1, d1 = d >> 1;
    uint32_t valid = M;
    if (RAGGED) {
      const int64_t rem = n_snps - (first_snp + 16 * i);
      valid = rem >= 16 ? M : (rem <= 0 ? 0u : (((1u << (2 * rem)) - 1u) & M));
    }
    const uint32_t par_ok = ~((s & s1) | (d & d1)) & valid;
    const uint32_t s0 = ~(s | s1), s2 = s1 & ~s, d0 = ~(d | d1), d2 = d1 & ~d;
    const uint32_t no0 = (s2 | d2) & par_ok, no2 = (s0 | d0) & par_ok, no1 = ((s0 & d0) | (s2 & d2)) & par_ok;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const uint32_t kk = i == 0 ? kq[k].x : (i == 1 ? kq[k].y : (i == 2 ? kq[k].z : kq[k].w));
      const uint32_t kh = kk >> 1;
      const uint32_t lo = (kk & no1) | (~kk & no0);               // kid 1 -> no1, kid 0 -> no0 (low bit of the code selects)
      const uint32_t bad = (kh & ~kk & no2) | (~kh & lo);         // P(kid | sire, dam) = 0  (bayesnet/model.py:28-30)
      const uint32_t present = ~(kk & kh) & par_ok;
      if ((i & 1) == 0) { b_even[k] = bad; p_even[k] = present; }
      else {
        n_inc[k] += __popc(bad * 2u + b_even[k]);                 // disjoint bit positions: one popcount for two words
        n_tst[k] += __popc(present * 2u + p_even[k]);
      }
    }
  }
}

// NK kids of one family (compile-time count: no predicated work for absent kids)
template <int NK>
__device__ __forceinline__ void trio_family_group(const uint32_t* __restrict__ packed, int64_t wpr, int64_t nquads, int64_t n_snps,
                                                  const uint4* ps, const uint4* pd, const int32_t* __restrict__ rows, int lane,
                                                  long long* __restrict__ inc, long long* __restrict__ tested) {
  const uint4* pk[NK];
  int n_inc[NK], n_tst[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    pk[k] = reinterpret_cast<const uint4*>(packed + (int64_t)rows[k] * wpr);
    n_inc[k] = 0; n_tst[k] = 0;
  }
  for (int64_t q = lane; q < nquads; q += 32) {
    const uint4 sq = __ldg(ps + q), dq = __ldg(pd + q);
    uint4 kq[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) kq[k] = __ldg(pk[k] + q);
    if ((q + 1) * 64 <= n_snps) trio_quad<NK, false>(sq, dq, kq, q * 64, n_snps, n_inc, n_tst);
    else trio_quad<NK, true>(sq, dq, kq, q * 64, n_snps, n_inc, n_tst);
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    int a = n_inc[k], b = n_tst[k];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { inc[rows[k]] += a; tested[rows[k]] += b; }   // a kid belongs to one family: no other writer in this launch
  }
}

__global__ void __launch_bounds__(256, GFX_FAM_BLOCKS) k_trio_scan(const uint32_t* __restrict__ packed, int64_t wpr, int64_t nwords,
                                                   int64_t n_snps, const int32_t* __restrict__ trio_row,
                                                   const int32_t* __restrict__ fam_off, int n_fam,
                                                   const int32_t* __restrict__ prow, long long* __restrict__ inc,
                                                   long long* __restrict__ tested) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int64_t nquads = (nwords + 3) >> 2;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n_fam; f += warps) {
    const int t0 = fam_off[f], t1 = fam_off[f + 1];
    const int r0 = trio_row[t0];
    const uint4* ps = reinterpret_cast<const uint4*>(packed + (int64_t)prow[2 * r0] * wpr);
    const uint4* pd = reinterpret_cast<const uint4*>(packed + (int64_t)prow[2 * r0 + 1] * wpr);
    for (int g0 = t0; g0 < t1; g0 += 4) {
      const int nk = min(4, t1 - g0);
      const int32_t* rows = trio_row + g0;
      if (nk == 4) trio_family_group<4>(packed, wpr, nquads, n_snps, ps, pd, rows, lane, inc, tested);
      else if (nk == 3) trio_family_group<3>(packed, wpr, nquads, n_snps, ps, pd, rows, lane, inc, tested);
      else if (nk == 2) trio_family_group<2>(packed, wpr, nquads, n_snps, ps, pd, rows, lane, inc, tested);
      else trio_family_group<1>(packed, wpr, nquads, n_snps, ps, pd, rows, lane, inc, tested);
    }
  }
}

// ---- TMA-staged variant of the scan ----------------------------------------------------------------------------------
// The same family walk, but the row segments come through shared memory: lane 0 of every warp issues 1-D bulk copies
// (cp.async.bulk global -> shared, SASS UBLKCP, completion counted in bytes on an mbarrier) of the next 512-byte segment
// of the sire's, the dam's and the kids' rows into a warp-private ring of GFX_TMA_STAGES stages while the warp scans the
// current segment from shared memory.  No per-lane address arithmetic, no registers held by loads in flight, and
// GFX_TMA_STAGES - 1 segments of every row in flight per warp.
#ifndef GFX_TMA_STAGES
#define GFX_TMA_STAGES 3
#endif
#ifndef GFX_TRIO_TMA_DEFAULT
#define GFX_TRIO_TMA_DEFAULT 0
#endif
#define GFX_TMA_ROWS 6        // sire, dam, up to 4 kids
#define GFX_TMA_WARPS 8
#define GFX_TMA_SMEM (GFX_TMA_WARPS * GFX_TMA_STAGES * GFX_TMA_ROWS * 512 + GFX_TMA_WARPS * GFX_TMA_STAGES * 8)
template <int NK>
__device__ __forceinline__ void trio_family_group_tma(const uint32_t* __restrict__ packed, int64_t wpr, int64_t nquads, int64_t n_snps,
                                                      int row_s, int row_d, const int32_t* __restrict__ rows, int lane, uint4* buf,
                                                      uint64_t* bars, uint32_t& parity, long long* __restrict__ inc,
                                                      long long* __restrict__ tested) {
  const char* rp[2 + NK];
  rp[0] = reinterpret_cast<const char*>(packed + (int64_t)row_s * wpr);
  rp[1] = reinterpret_cast<const char*>(packed + (int64_t)row_d * wpr);
  int n_inc[NK], n_tst[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    rp[2 + k] = reinterpret_cast<const char*>(packed + (int64_t)rows[k] * wpr);
    n_inc[k] = 0; n_tst[k] = 0;
  }
  const int64_t row_bytes = nquads * 16;
  const int64_t nsteps = (nquads + 31) >> 5;
  auto issue = [&](int64_t step) {
    if (lane == 0) {
      const int st = (int)(step % GFX_TMA_STAGES);
      const int64_t off = step * 512;
      const uint32_t bytes = (uint32_t)(row_bytes - off < 512 ? row_bytes - off : 512);
      mbar_expect_tx(bars + st, bytes * (2 + NK));
#pragma unroll
      for (int r = 0; r < 2 + NK; ++r) tma_load_1d(buf + (st * GFX_TMA_ROWS + r) * 32, rp[r] + off, bytes, bars + st);
    }
  };
  for (int64_t p = 0; p < GFX_TMA_STAGES - 1 && p < nsteps; ++p) issue(p);
  for (int64_t step = 0; step < nsteps; ++step) {
    const int st = (int)(step % GFX_TMA_STAGES);
    // the stage refilled here was read in the previous iteration; the __syncwarp below closed those reads
    if (step + GFX_TMA_STAGES - 1 < nsteps) issue(step + GFX_TMA_STAGES - 1);
    mbar_wait(bars + st, (parity >> st) & 1u);
    parity ^= 1u << st;
    const int64_t q = step * 32 + lane;
    if (q < nquads) {
      const uint4 sq = buf[(st * GFX_TMA_ROWS + 0) * 32 + lane], dq = buf[(st * GFX_TMA_ROWS + 1) * 32 + lane];
      uint4 kq[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) kq[k] = buf[(st * GFX_TMA_ROWS + 2 + k) * 32 + lane];
      if ((q + 1) * 64 <= n_snps) trio_quad<NK, false>(sq, dq, kq, q * 64, n_snps, n_inc, n_tst);
      else trio_quad<NK, true>(sq, dq, kq, q * 64, n_snps, n_inc, n_tst);
    }
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    int a = n_inc[k], b = n_tst[k];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { inc[rows[k]] += a; tested[rows[k]] += b; }
  }
}

__global__ void __launch_bounds__(GFX_TMA_WARPS * 32) k_trio_scan_tma(const uint32_t* __restrict__ packed, int64_t wpr, int64_t nwords,
                                                                      int64_t n_snps, const int32_t* __restrict__ trio_row,
                                                                      const int32_t* __restrict__ fam_off, int n_fam,
                                                                      const int32_t* __restrict__ prow, long long* __restrict__ inc,
                                                                      long long* __restrict__ tested) {
  extern __shared__ __align__(128) unsigned char s_tma[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint4* buf = reinterpret_cast<uint4*>(s_tma) + (size_t)wib * GFX_TMA_STAGES * GFX_TMA_ROWS * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tma + GFX_TMA_WARPS * GFX_TMA_STAGES * GFX_TMA_ROWS * 512) + wib * GFX_TMA_STAGES;
  if (lane == 0) {
    for (int i = 0; i < GFX_TMA_STAGES; ++i) mbar_init(bars + i, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;   // bit s: phase the next wait on stage s looks for
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int64_t nquads = (nwords + 3) >> 2;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n_fam; f += warps) {
    const int t0 = fam_off[f], t1 = fam_off[f + 1];
    const int r0 = trio_row[t0];
    const int rs = prow[2 * r0], rd = prow[2 * r0 + 1];
    for (int g0 = t0; g0 < t1; g0 += 4) {
      const int nk = min(4, t1 - g0);
      const int32_t* rows = trio_row + g0;
      if (nk == 4) trio_family_group_tma<4>(packed, wpr, nquads, n_snps, rs, rd, rows, lane, buf, bars, parity, inc, tested);
      else if (nk == 3) trio_family_group_tma<3>(packed, wpr, nquads, n_snps, rs, rd, rows, lane, buf, bars, parity, inc, tested);
      else if (nk == 2) trio_family_group_tma<2>(packed, wpr, nquads, n_snps, rs, rd, rows, lane, buf, bars, parity, inc, tested);
      else trio_family_group_tma<1>(packed, wpr, nquads, n_snps, rs, rd, rows, lane, buf, bars, parity, inc, tested);
    }
  }
}

static int launch_trio_scan(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr, int64_t* inc,
                            int64_t* tested, bool zero, cudaStream_t stream) {
  if (!s || !packed || !inc || !tested || n_snps <= 0) return fail(GFX_E_INVALID, "gfx_trio_scan: bad argument");
  if (s->device < 0) return fail(GFX_E_INVALID, "host-only schedule passed to a kernel entry point");
  if (wpr % 4 || (((uintptr_t)packed) % 16)) return fail(GFX_E_INVALID, "gfx_trio_scan: rows must be 16-byte aligned (wpr % 4 == 0)");
  if (zero) {
    GFX_CUDA(cudaMemsetAsync(inc, 0, sizeof(int64_t) * s->h.n_rows, stream));
    GFX_CUDA(cudaMemsetAsync(tested, 0, sizeof(int64_t) * s->h.n_rows, stream));
  }
  const int nt = (int)s->h.trio_row.size();
  if (nt == 0) return GFX_OK;
  const int nfam = (int)s->h.fam_off.size() - 1;
  static const int use_tma = []() { const char* e = getenv("GFX_TRIO_TMA"); return e ? atoi(e) : GFX_TRIO_TMA_DEFAULT; }();
  if (use_tma) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, []() {
      attr_err = cudaFuncSetAttribute(k_trio_scan_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GFX_TMA_SMEM);
    });
    GFX_CUDA(attr_err);
    const int per_sm = (int)(220 * 1024 / GFX_TMA_SMEM) < 1 ? 1 : (int)(220 * 1024 / GFX_TMA_SMEM);
    k_trio_scan_tma<<<grid_for((int64_t)nfam * 32, GFX_TMA_WARPS * 32, s->num_sms, per_sm), GFX_TMA_WARPS * 32, GFX_TMA_SMEM, stream>>>(
        packed, wpr, (n_snps + 15) / 16, n_snps, s->trio_row, s->fam_off, nfam, s->prow, (long long*)inc, (long long*)tested);
  } else {
    k_trio_scan<<<grid_for((int64_t)nfam * 32, 256, s->num_sms, 8), 256, 0, stream>>>(
        packed, wpr, (n_snps + 15) / 16, n_snps, s->trio_row, s->fam_off, nfam, s->prow, (long long*)inc, (long long*)tested);
  }
  GFX_LAUNCHED();
  return GFX_OK;
}

extern "C" int gfx_trio_scan(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr, int64_t* inc,
                             int64_t* tested, void* stream) {
  return launch_trio_scan(s, packed, n_snps, wpr, inc, tested, true, (cudaStream_t)stream);
}

// The same scan ADDED to the counters: for a matrix held as column chunks (the pipeline's own resident layout), one
// call per chunk.
extern "C" int gfx_trio_scan_add(const gfx_schedule* s, const uint32_t* packed, int64_t n_snps, int64_t wpr, int64_t* inc,
                                 int64_t* tested, void* stream) {
  return launch_trio_scan(s, packed, n_snps, wpr, inc, tested, false, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// LD-window joint counts
// ------------------------------------------------------------------------------------------------
// Layout of one window's counters: [3^len full window][3^(len-2) inner window 1][3^(len-4) inner window 2]...,
// inner windows 1 .. int((len - 2) / 2) - 1 -- the nested terms of jalleledist3.py:133-138 (the i = 0 pass of
// that loop is the empty slice, D12, and needs no counter).  A row is in the reference's `subset` when ALL cells of
// the window pass the mask; the inner window i then counts the rows whose INNER cells match, whatever the outer
// cells hold (missing calls included), so it is not a marginal of the full histogram.
__host__ __device__ __forceinline__ int gfx_window_levels(int len) { const int k = (len - 2) / 2; return k < 1 ? 1 : k; }
__global__ void __launch_bounds__(256) k_window_counts(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ mask,
                                                       int64_t n, int64_t wpr, int64_t mwpr, const int32_t* __restrict__ win_start,
                                                       const int32_t* __restrict__ win_len, int n_windows, int stride,
                                                       uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t s_cnt[];
  for (int wi = blockIdx.x; wi < n_windows; wi += gridDim.x) {
    const int start = win_start[wi], len = win_len[wi];
    const int nlev = gfx_window_levels(len);
    int total = 0;
    for (int l = 0; l < nlev; ++l) total += gfx_pow3(len - 2 * l);
    for (int i = threadIdx.x; i < total; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
      uint32_t code[16];
      bool ok = len > 0;
      for (int c = 0; c < len; ++c) {
        const int64_t j = start + c;
        code[c] = (__ldg(packed + r * wpr + (j >> 4)) >> (2 * (j & 15))) & 3u;
        if (mask && !((__ldg(mask + r * mwpr + (j >> 5)) >> (j & 31)) & 1u)) { ok = false; break; }
      }
      if (!ok) continue;
      int off = 0;
      for (int l = 0; l < nlev; ++l) {
        int idx = 0;
        bool present = true;
        for (int c = l; c < len - l; ++c) {
          if (code[c] == 3u) { present = false; break; }
          idx = idx * 3 + (int)code[c];
        }
        if (present) atomicAdd(&s_cnt[off + idx], 1u);
        off += gfx_pow3(len - 2 * l);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) counts[(int64_t)wi * stride + i] = s_cnt[i];
    __syncthreads();
  }
}

extern "C" int gfx_window_counts(const uint32_t* packed, const uint32_t* mask, int64_t n, int64_t n_snps, int64_t wpr,
                                 int64_t mwpr, const int32_t* win_start, const int32_t* win_len, int32_t n_windows,
                                 int32_t stride, uint32_t* counts, void* stream) {
  if (!packed || !win_start || !win_len || !counts || n <= 0 || n_windows <= 0 || stride <= 0)
    return fail(GFX_E_INVALID, "gfx_window_counts: bad argument");
  (void)n_snps;
  const size_t smem = (size_t)stride * sizeof(uint32_t);
  if (smem > 200 * 1024) return fail(GFX_E_INVALID, "window too wide for shared memory (surround <= 4)");
  {
    // stride must hold the full histogram and the inner-window histograms of the widest window (<= 16 SNPs)
    std::vector<int32_t> lens((size_t)n_windows);
    GFX_CUDA(cudaMemcpyAsync(lens.data(), win_len, sizeof(int32_t) * (size_t)n_windows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GFX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int32_t len : lens) {
      if (len < 0 || len > 16) return fail(GFX_E_INVALID, "gfx_window_counts: window length out of range");
      int64_t total = 0;
      for (int l = 0; l < gfx_window_levels(len); ++l) total += gfx_pow3(len - 2 * l);
      if (total > stride) return fail(GFX_E_INVALID, "gfx_window_counts: stride smaller than the window's counters");
    }
  }
  GFX_CUDA(cudaFuncSetAttribute(k_window_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = n_windows < device_sms() * 4 ? n_windows : device_sms() * 4;
  k_window_counts<<<blocks, 256, smem, (cudaStream_t)stream>>>(packed, mask, n, wpr, mwpr, win_start, win_len, n_windows,
                                                              stride, counts);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
// the empirical (LD-window) blend of correctMatrix(empC != None): correct_genotypes3.py:619-667 (ranks), :746-779 and
// :883-908 (decisions); jalleledist3.py:101-117 (getCountTable)
// ------------------------------------------------------------------------------------------------
#define GFX_EMP_WMAX 9
struct EmpDev {
  const int32_t* wstart;
  const int32_t* wlen;
  const int64_t* foff;
  const double* freq;
  double weight;
};
// getCountTable: stored frequency of each state of SNP j given the calls of one animal (row pointer rp, 2-bit codes) in
// j's window; a missing call in the window is summed over its completions, first window SNP = slowest digit, plain
// left-to-right additions (the reference raises there, SURVEY D4: the oracle defines this case).  When the window does
// not hold SNP j itself (the last SNP of a chromosome, D11) the three values are the same entry.
__device__ __forceinline__ void emp_count_table(const uint32_t* __restrict__ rp, int64_t j, const EmpDev& X, double c[3]) {
  const int a = X.wstart[j], w = X.wlen[j];
  const double* __restrict__ f = X.freq + X.foff[j];
  int base = 0, tstride = 0, nmiss = 0, mstride[GFX_EMP_WMAX];
  int stride = 1;
  for (int k = w - 1; k >= 0; --k) {      // last window SNP = fastest digit
    const int64_t col = a + k;
    if (col == j) tstride = stride;
    else {
      const int v = (int)((rp[col >> 4] >> (2 * (col & 15))) & 3u);
      if (v == 3) mstride[nmiss++] = stride;   // collected fastest digit first
      else base += v * stride;
    }
    stride *= 3;
  }
  if (nmiss == 0) {
    c[0] = f[base]; c[1] = f[base + tstride]; c[2] = f[base + 2 * tstride];
    return;
  }
  int total = 1;
  for (int k = 0; k < nmiss; ++k) total *= 3;
  for (int s = 0; s < 3; ++s) {
    double acc = 0.0;
    for (int m = 0; m < total; ++m) {       // m counts in C order: the slowest missing digit is the most significant
      int off = base + s * tstride, r = m;
      for (int k = 0; k < nmiss; ++k) { off += (r % 3) * mstride[k]; r /= 3; }
      acc = m == 0 ? f[off] : acc + f[off];
    }
    c[s] = acc;
  }
}
// :752-755 / :888-889: the table normalised to sum 1; false when the sum is not positive (:754 keeps the previous value
// then; with a positive pseudocount it cannot happen)
__device__ __forceinline__ bool emp_state_probs(const uint32_t* rp, int64_t j, const EmpDev& X, double nrm[3]) {
  double c[3];
  emp_count_table(rp, j, X, c);
  const double tot = (c[0] + c[1]) + c[2];
  if (!(tot > 0.0)) return false;
  nrm[0] = c[0] / tot; nrm[1] = c[1] / tot; nrm[2] = c[2] / tot;
  return true;
}

// ranks (:619-656): E(cell) = mean(E, weight * (max - own) of the normalised window frequencies) for every observed cell of a
// SNP whose window holds it; the division by the new maximum, the missing cells and the quantile follow on the tensors.
__global__ void __launch_bounds__(256) k_emp_rank_blend(const uint32_t* __restrict__ packed, int64_t n, int64_t n_snps, int64_t wpr,
                                                        const uint16_t* __restrict__ raw, int64_t ld_raw,
                                                        const double* __restrict__ elut, EmpDev X, double* __restrict__ e_out,
                                                        int64_t ld_e) {
  const int64_t total = n * n_snps;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / n_snps, j = t - r * n_snps;
    const uint32_t* rp = packed + r * wpr;
    const int obs = (int)((rp[j >> 4] >> (2 * (j & 15))) & 3u);
    const uint32_t rx = raw[r * ld_raw + j];
    double E = (obs == 3 || rx == 0xFFFFu) ? 1.0 : elut[rx];
    const int a = X.wstart[j], w = X.wlen[j];
    if (obs != 3 && j >= a && j < a + w) {
      double nrm[3];
      if (emp_state_probs(rp, j, X, nrm)) {
        double mx = nrm[0];
        if (nrm[1] > mx) mx = nrm[1];
        if (nrm[2] > mx) mx = nrm[2];
        const double empdiff = mx - nrm[obs];
        E = (E + empdiff * X.weight) / 2.0;      // np.mean of the two
      }
    }
    e_out[r * ld_e + j] = E;
  }
}

// decisions (:714-845 pairs, :866-944 singles) with an index: one thread per focal animal walks its jobs in canonical order
// and, inside a job, the SNPs in column order -- the window evidence is the animal's LIVE row, so its cells are no longer
// independent of each other (a correction at SNP j changes the evidence of j + 1 and of the next job).  The posteriors of
// the jobs come from the inference kernels (post, computed on the generation-start snapshot, sentinel = no result).
struct EmpApplyArgs {
  int nf, n_focal;
  const int32_t* focal_row;
  const int32_t* focal_off;
  const int32_t* entry;
  uint32_t* live;
  const uint32_t* snapshot;
  int64_t n_snps, wpr;
  const uint16_t* raw;
  int64_t ld_raw;
  const double* elut;
  const int32_t* prow;
  EmpDev X;
  double tie, threshold, err_thresh, sentinel;
  const double* post;
  int32_t* tallies;
  int n_rows;
};
__global__ void __launch_bounds__(64) k_emp_apply(EmpApplyArgs a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n_focal) return;
  const int row = a.focal_row[f];
  uint32_t* rp = a.live + (int64_t)row * a.wpr;
  const uint32_t* sp = a.snapshot + (int64_t)row * a.wpr;
  int applied = 0, excluded = 0;
  for (int e = a.focal_off[f]; e < a.focal_off[f + 1]; ++e) {
    const int ent = a.entry[e];
    const double* P = a.post + (int64_t)ent * a.n_snps * 3;
    for (int64_t j = 0; j < a.n_snps; ++j) {
      const double p0 = P[3 * j];
      if (p0 == a.sentinel) continue;
      const double p[3] = {p0, P[3 * j + 1], P[3 * j + 2]};
      const int cur = (int)((rp[j >> 4] >> (2 * (j & 15))) & 3u);
      const int obs0 = (int)((sp[j >> 4] >> (2 * (j & 15))) & 3u);
      double nrm[3];
      if (!emp_state_probs(rp, j, a.X, nrm)) continue;
      double comb[3];
      for (int i = 0; i < 3; ++i) {
        const double nw = nrm[i] * a.X.weight;
        comb[i] = p[i] == p[i] ? (p[i] + nw) / 2.0 : nw;       // np.nanmean over the two rows
      }
      double nsum = 0.0;
      for (int i = 0; i < 3; ++i) if (comb[i] == comb[i]) nsum += comb[i];
      if (a.nf == 1 || nsum > 0.0) {
        const double tot = (comb[0] + comb[1]) + comb[2];
        comb[0] /= tot; comb[1] /= tot; comb[2] /= tot;
      }
      double mx = nan("");
      for (int i = 0; i < 3; ++i) if (comb[i] == comb[i] && !(mx >= comb[i])) mx = comb[i];
      uint32_t states = 0;
      const double cut = mx - a.tie;
      for (int i = 0; i < 3; ++i) if (comb[i] >= cut) states |= 1u << i;
      double op;
      if (a.nf == 2) op = obs0 <= 2 ? p[obs0] : nan("");      // pairs keep the Mendelian probability of the snapshot state (:811)
      else op = cur == 3 ? 0.0 : comb[cur];                   // singles: the blended one (:904)
      const bool cond = cur == 3 || (!((states >> cur) & 1u) && (mx - op) >= a.threshold);
      if (!cond) continue;
      double mr = 0.0;
      for (int q = 0; q < 2; ++q) {
        const int pr = a.prow[2 * row + q];
        if (pr >= 0) {
          const uint32_t rx = a.raw[(int64_t)pr * a.ld_raw + j];
          const double ev = rx == 0xFFFFu ? 1.0 : a.elut[rx];
          if (ev > mr) mr = ev;
        }
      }
      const uint32_t rxf = a.raw[(int64_t)row * a.ld_raw + j];
      const double ef = rxf == 0xFFFFu ? 1.0 : a.elut[rxf];
      if (ef * a.err_thresh >= mr) {
        const int nw = __popc(states) == 1 ? (__ffs(states) - 1) : 3;
        const int sh = 2 * (int)(j & 15);
        rp[j >> 4] = (rp[j >> 4] & ~(3u << sh)) | ((uint32_t)nw << sh);   // the row belongs to this thread
        ++applied;
      } else {
        ++excluded;
      }
    }
  }
  if (applied) atomicAdd(a.tallies + row, applied);
  if (excluded && a.nf == 2) atomicAdd(a.tallies + a.n_rows + row, excluded);
}

static int emp_index_check(const gfx_emp_index* idx) {
  if (!idx || !idx->wstart_dev || !idx->wlen_dev || !idx->foff_dev || !idx->freq_dev) return fail(GFX_E_INVALID, "null LD index");
  if (idx->max_window > GFX_EMP_WMAX) return fail(GFX_E_INVALID, "LD windows of more than 9 SNPs are not supported");
  return GFX_OK;
}

extern "C" int gfx_emp_rank_blend(const uint32_t* packed, int64_t n, int64_t n_snps, int64_t wpr, const uint16_t* raw,
                                  int64_t ld_raw, const double* elut, const gfx_emp_index* idx, double* e_out, int64_t ld_e,
                                  void* stream) {
  if (!packed || !raw || !elut || !e_out || n <= 0 || n_snps <= 0) return fail(GFX_E_INVALID, "null argument");
  int rc = emp_index_check(idx);
  if (rc != GFX_OK) return rc;
  EmpDev X{idx->wstart_dev, idx->wlen_dev, idx->foff_dev, idx->freq_dev, idx->weight};
  k_emp_rank_blend<<<grid_for(n * n_snps, 256, device_sms(), 8), 256, 0, (cudaStream_t)stream>>>(packed, n, n_snps, wpr, raw, ld_raw,
                                                                                                 elut, X, e_out, ld_e);
  GFX_LAUNCHED();
  return GFX_OK;
}

extern "C" int gfx_emp_apply(const gfx_schedule* s, int32_t generation, uint32_t* live, const uint32_t* snapshot, int64_t n_snps,
                             int64_t wpr, const uint16_t* raw, int64_t ld_raw, const double* elut, const gfx_emp_index* idx,
                             const gfx_correct_params* p, const double* post, double post_sentinel, int32_t* tallies, void* stream) {
  if (!s || !live || !snapshot || !raw || !elut || !p || !post || !tallies) return fail(GFX_E_INVALID, "null argument");
  if (s->device < 0) return fail(GFX_E_INVALID, "host-only schedule passed to a kernel entry point");
  int rc = emp_index_check(idx);
  if (rc != GFX_OK) return rc;
  EmpApplyArgs a;
  if (generation >= 0) {
    if (generation >= s->h.info.n_generations) return fail(GFX_E_INVALID, "generation out of range");
    const int f0 = s->h.app_gen_off[generation];
    a.nf = 2; a.n_focal = s->h.app_gen_off[generation + 1] - f0;
    a.focal_row = s->app_focal_row + f0; a.focal_off = s->app_focal_off + f0; a.entry = s->app_entry;
  } else {
    a.nf = 1; a.n_focal = (int)s->h.single_row.size();
    a.focal_row = s->single_row; a.focal_off = s->single_off; a.entry = s->single_entry;
  }
  if (a.n_focal <= 0) return GFX_OK;
  a.live = live; a.snapshot = snapshot; a.n_snps = n_snps; a.wpr = wpr; a.raw = raw; a.ld_raw = ld_raw; a.elut = elut;
  a.prow = s->prow;
  a.X = EmpDev{idx->wstart_dev, idx->wlen_dev, idx->foff_dev, idx->freq_dev, idx->weight};
  a.tie = p->tiethreshold; a.threshold = p->threshold; a.err_thresh = p->err_thresh; a.sentinel = post_sentinel;
  a.post = post; a.tallies = tallies; a.n_rows = s->h.n_rows;
  k_emp_apply<<<(a.n_focal + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a);
  GFX_LAUNCHED();
  return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int gfx_build_cut_table(const double* elut, double test_p, double suspect_t, uint16_t* cutraw, uint32_t* tp_raw) {
  if (!elut || !cutraw || !tp_raw) return fail(GFX_E_INVALID, "null argument");
  // finite non-negative half patterns 0 .. 0x7C00 order like their values and E is monotone in the score
  const uint32_t top = 0x7C00u;  // +inf: never reached by a score, used as "no pattern qualifies"
  uint32_t t = top + 1;
  for (uint32_t p = 0; p <= top; ++p)
    if (elut[p] >= test_p) { t = p; break; }
  *tp_raw = t;
  // cutraw[r] = first pattern whose E > E(r) - suspect_t; two-pointer sweep (both sides monotone)
  uint32_t q = 0;
  for (uint32_t r = 0; r <= 65536u; ++r) {
    if (r > top && r < 65536u) { cutraw[r] = (uint16_t)(top + 1); continue; }  // NaN / negative patterns: unused
    const double cut = (r == 65536u ? 1.0 : elut[r]) - suspect_t;
    if (r == 65536u) q = 0;
    while (q <= top && !(elut[q] > cut)) ++q;
    cutraw[r] = (uint16_t)q;
  }
  return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
// Simulator pieces on the device: gene drop, error injection, evaluation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t gfx_mix64(uint64_t x) {   // splitmix64 finaliser: counter-based, no state
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t gfx_rand(uint64_t seed, uint64_t a, uint64_t b) {
  return gfx_mix64(gfx_mix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b);
}

// One level of the pedigree (parents are in earlier levels).  hap word: bit 2i = paternal, bit 2i+1 = maternal allele
// of SNP i.  Every SNP segregates independently (free recombination): the transmitted allele of a parent is one
// of its two alleles with probability 1/2 each (simulate.py:187-230 with quickgsim's meiosis reduced to independent
// loci); founders draw both alleles with frequency 1/2 (simulate.py:206-207).
__global__ void __launch_bounds__(256) k_gene_drop(const int32_t* __restrict__ order, int n_level, const int32_t* __restrict__ sire_row,
                                                   const int32_t* __restrict__ dam_row, int64_t nwords, int64_t wpr, uint64_t seed,
                                                   uint32_t* __restrict__ hap) {
  const int64_t total = (int64_t)n_level * nwords;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / nwords, w = i - k * nwords;
    const int r = order[k];
    const int rs = sire_row[r], rd = dam_row[r];
    const uint64_t rnd = gfx_rand(seed, (uint64_t)r, (uint64_t)w);
    const uint32_t pick_s = (uint32_t)rnd & 0x55555555u, pick_d = (uint32_t)(rnd >> 32) & 0x55555555u;
    uint32_t pat, mat;   // alleles at the even bit positions
    if (rs >= 0) {
      const uint32_t h = hap[(int64_t)rs * wpr + w];
      pat = ((h & ~pick_s) | ((h >> 1) & pick_s)) & 0x55555555u;
    } else {
      pat = (uint32_t)(gfx_mix64(rnd) & 0x55555555u);
    }
    if (rd >= 0) {
      const uint32_t h = hap[(int64_t)rd * wpr + w];
      mat = ((h & ~pick_d) | ((h >> 1) & pick_d)) & 0x55555555u;
    } else {
      mat = (uint32_t)((gfx_mix64(rnd) >> 32) & 0x55555555u);
    }
    hap[(int64_t)r * wpr + w] = pat | (mat << 1);
  }
}

extern "C" int gfx_gene_drop(int32_t n, const int32_t* sire_row_dev, const int32_t* dam_row_dev, const int32_t* order_dev,
                             const int32_t* level_off_host, int32_t n_levels, int64_t n_snps, int64_t wpr, uint64_t seed,
                             uint32_t* hap_dev, void* stream) {
  if (n <= 0 || !sire_row_dev || !dam_row_dev || !order_dev || !level_off_host || n_levels <= 0 || !hap_dev || n_snps <= 0)
    return fail(GFX_E_INVALID, "gfx_gene_drop: bad argument");
  const int64_t nwords = (n_snps + 15) / 16;
  if (wpr < nwords) return fail(GFX_E_INVALID, "gfx_gene_drop: wpr < ceil(n_snps / 16)");
  for (int l = 0; l < n_levels; ++l) {
    const int cnt = level_off_host[l + 1] - level_off_host[l];
    if (cnt <= 0) continue;
    k_gene_drop<<<grid_for((int64_t)cnt * nwords, 256, device_sms(), 8), 256, 0, (cudaStream_t)stream>>>(
        order_dev + level_off_host[l], cnt, sire_row_dev, dam_row_dev, nwords, wpr, seed, hap_dev);
    GFX_LAUNCHED();
  }
  return GFX_OK;
}

// ---- gene drop with linkage: quickgsim's meiosis (quickgsim/genome.py:74-115) ------------------------------------------
// Per gamete and chromosome: n = max(1, Poisson(morgans)) crossovers (one obligatory, :75-79), placed on distinct interior
// markers 1 .. n_markers - 2 with probability proportional to the genetic distance to the previous marker (:81-86, Chrom.cm_to_p
// :35-41), sorted; the gamete starts on a random strand and switches at every crossover marker (:98-113).  Everything a
// gamete needs is a function of (seed, animal, side, chromosome) through the counter-based generator, so every thread
// (animal, word) recomputes the few crossovers of the chromosomes its word touches: no shared state, no ordering between
// threads, results independent of the launch geometry.  Successive weighted draws with rejection of repeats are the same
// distribution as numpy's choice(p, replace=False).
#define GFX_MAX_XOVER 12
struct LinkMap {
  int n_chrom;
  const int32_t* chrom_start;   // [n_chrom + 1] first SNP of every chromosome (SNPs in map order)
  const float* morgans;         // [n_chrom]
  const float* cum_p;           // [n_snps] per chromosome: cumulative crossover probability up to and including the marker
};                              //          (0 at the first marker, 1 at the last but one and the last)
__device__ __forceinline__ int gfx_poisson(float lambda, uint64_t& st) {
  // inversion by sequential search (lambda is a chromosome length in Morgans: ~1)
  const float u = (float)((gfx_mix64(st += 0x9E3779B97F4A7C15ull) >> 40) + 0.5f) * (1.0f / 16777216.0f);
  float p = __expf(-lambda), c = p;
  int k = 0;
  while (u > c && k < 64) { ++k; p *= lambda / (float)k; c += p; }
  return k;
}
// crossover markers (local indices, ascending) and the start strand of one gamete on one chromosome
__device__ __forceinline__ int gfx_gamete_plan(const LinkMap& L, int c, uint64_t seed, uint64_t animal, int side, int* brk, int* start) {
  const int s0 = L.chrom_start[c], nm = L.chrom_start[c + 1] - s0;
  uint64_t st = gfx_rand(seed ^ 0xC2B2AE3D27D4EB4Full, animal * 2ull + (uint64_t)side, (uint64_t)c);
  *start = (int)(gfx_mix64(st += 0x9E3779B97F4A7C15ull) >> 63);
  if (nm < 3) return 0;                                   // no interior marker: the whole chromosome is one segment (:83-86)
  int n = gfx_poisson(L.morgans[c], st);
  if (n < 1) n = 1;                                        // obligatory crossover
  if (n > nm - 2) n = nm - 2;
  if (n > GFX_MAX_XOVER) n = GFX_MAX_XOVER;
  const float* cp = L.cum_p + s0;
  int got = 0;
  for (int tries = 0; got < n && tries < 8 * GFX_MAX_XOVER; ++tries) {
    const float u = (float)((gfx_mix64(st += 0x9E3779B97F4A7C15ull) >> 40) + 0.5f) * (1.0f / 16777216.0f);
    int lo = 1, hi = nm - 2;                               // first marker whose cumulative probability reaches u
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cp[mid] >= u) hi = mid; else lo = mid + 1; }
    bool dup = false;
    for (int k = 0; k < got; ++k) dup |= brk[k] == lo;
    if (!dup) brk[got++] = lo;
  }
  for (int i = 1; i < got; ++i) {                          // insertion sort
    const int v = brk[i];
    int k = i - 1;
    while (k >= 0 && brk[k] > v) { brk[k + 1] = brk[k]; --k; }
    brk[k + 1] = v;
  }
  return got;
}
__global__ void __launch_bounds__(128) k_gene_drop_linked(const int32_t* __restrict__ order, int n_level, const int32_t* __restrict__ sire_row,
                                                          const int32_t* __restrict__ dam_row, int64_t nwords, int64_t wpr, int64_t n_snps,
                                                          uint64_t seed, LinkMap L, uint32_t* __restrict__ hap) {
  const int64_t total = (int64_t)n_level * nwords;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / nwords, w = i - k * nwords;
    const int r = order[k];
    const int prow[2] = {sire_row[r], dam_row[r]};
    const uint64_t rnd = gfx_rand(seed, (uint64_t)r, (uint64_t)w);
    uint32_t out = 0;
    // chromosome of the word's first SNP
    const int64_t j0 = w * 16;
    int c = 0;
    {
      int lo = 0, hi = L.n_chrom - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (L.chrom_start[mid] <= j0) lo = mid; else hi = mid - 1; }
      c = lo;
    }
    for (int side = 0; side < 2; ++side) {
      uint32_t alle;   // transmitted allele of every SNP of the word at the even bit positions
      if (prow[side] < 0) {
        alle = (uint32_t)((side == 0 ? gfx_mix64(rnd) : (gfx_mix64(rnd) >> 32)) & 0x55555555u);     // founder: frequency 1/2
      } else {
        const uint32_t h = hap[(int64_t)prow[side] * wpr + w];
        uint32_t pick = 0;                                  // bit 2i set: SNP i comes from the parent's maternal strand
        int cc = c, brk[GFX_MAX_XOVER], start = 0, nb = -1;
        for (int b = 0; b < 16; ++b) {
          const int64_t j = j0 + b;
          if (j >= n_snps) break;
          while (cc + 1 < L.n_chrom && j >= L.chrom_start[cc + 1]) { ++cc; nb = -1; }
          if (nb < 0) nb = gfx_gamete_plan(L, cc, seed, (uint64_t)r, side, brk, &start);
          const int local = (int)(j - L.chrom_start[cc]);
          int strand = start;
          for (int q = 0; q < nb; ++q) strand ^= (int)(brk[q] <= local);
          pick |= (uint32_t)strand << (2 * b);
        }
        alle = ((h & ~pick) | ((h >> 1) & pick)) & 0x55555555u;
      }
      out |= alle << side;
    }
    hap[(int64_t)r * wpr + w] = out;
  }
}

extern "C" int gfx_gene_drop_linked(int32_t n, const int32_t* sire_row_dev, const int32_t* dam_row_dev, const int32_t* order_dev,
                                    const int32_t* level_off_host, int32_t n_levels, int64_t n_snps, int64_t wpr, uint64_t seed,
                                    int32_t n_chrom, const int32_t* chrom_start_dev, const float* morgans_dev, const float* cum_p_dev,
                                    uint32_t* hap_dev, void* stream) {
  if (n <= 0 || !sire_row_dev || !dam_row_dev || !order_dev || !level_off_host || n_levels <= 0 || !hap_dev || n_snps <= 0 ||
      n_chrom <= 0 || !chrom_start_dev || !morgans_dev || !cum_p_dev)
    return fail(GFX_E_INVALID, "gfx_gene_drop_linked: bad argument");
  const int64_t nwords = (n_snps + 15) / 16;
  if (wpr < nwords) return fail(GFX_E_INVALID, "gfx_gene_drop_linked: wpr too small");
  LinkMap L{n_chrom, chrom_start_dev, morgans_dev, cum_p_dev};
  for (int l = 0; l < n_levels; ++l) {
    const int cnt = level_off_host[l + 1] - level_off_host[l];
    if (cnt <= 0) continue;
    k_gene_drop_linked<<<grid_for((int64_t)cnt * nwords, 128, device_sms(), 16), 128, 0, (cudaStream_t)stream>>>(
        order_dev + level_off_host[l], cnt, sire_row_dev, dam_row_dev, nwords, wpr, n_snps, seed, L, hap_dev);
    GFX_LAUNCHED();
  }
  return GFX_OK;
}

// genotype code = number of 1-alleles; cells beyond n_snps (and the padding words) become 3 = missing
__global__ void __launch_bounds__(256) k_haps_to_codes(const uint32_t* __restrict__ hap, int64_t n, int64_t n_snps, int64_t wpr,
                                                       uint32_t* __restrict__ packed) {
  const int64_t total = n * wpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = i % wpr;
    const uint32_t h = hap[i];
    const uint32_t p = h & 0x55555555u, m = (h >> 1) & 0x55555555u;
    uint32_t code = (p ^ m) | ((p & m) << 1);
    const int64_t rem = n_snps - w * 16;
    if (rem <= 0) code = 0xFFFFFFFFu;
    else if (rem < 16) code |= 0xFFFFFFFFu << (2 * rem);
    packed[i] = code;
  }
}
extern "C" int gfx_haps_to_codes(const uint32_t* hap_dev, int64_t n, int64_t n_snps, int64_t wpr, uint32_t* packed_dev, void* stream) {
  if (!hap_dev || !packed_dev || n <= 0 || n_snps <= 0 || wpr < (n_snps + 15) / 16) return fail(GFX_E_INVALID, "gfx_haps_to_codes: bad argument");
  k_haps_to_codes<<<grid_for(n * wpr, 256, device_sms(), 8), 256, 0, (cudaStream_t)stream>>>(hap_dev, n, n_snps, wpr, packed_dev);
  GFX_LAUNCHED();
  return GFX_OK;
}

// Every observed cell is hit with probability `rate` and then shows one of the two OTHER genotypes with equal
// probability (simulate.py:318-331 draws ceil(N S rate) cells with replacement; here the count is binomial).
__global__ void __launch_bounds__(256) k_inject_errors(uint32_t* __restrict__ packed, int64_t n, int64_t n_snps, int64_t wpr,
                                                       uint32_t thresh, uint64_t seed) {
  const int64_t nwords = (n_snps + 15) / 16;
  const int64_t total = n * nwords;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nwords, w = i - r * nwords;
    uint32_t v = packed[r * wpr + w];
    bool dirty = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint64_t rnd = gfx_rand(seed, (uint64_t)i, (uint64_t)k);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t u = (uint32_t)(rnd >> (32 * h));
        const int c = 2 * k + h;
        const uint32_t code = (v >> (2 * c)) & 3u;
        if ((u >> 1) < thresh && code != 3u) {
          const uint32_t other = (code + 1u + (u & 1u)) % 3u;
          v = (v & ~(3u << (2 * c))) | (other << (2 * c));
          dirty = true;
        }
      }
    }
    if (dirty) packed[r * wpr + w] = v;
  }
}
extern "C" int gfx_inject_errors(uint32_t* packed_dev, int64_t n, int64_t n_snps, int64_t wpr, double rate, uint64_t seed, void* stream) {
  if (!packed_dev || n <= 0 || n_snps <= 0 || wpr < (n_snps + 15) / 16 || !(rate >= 0.0 && rate <= 1.0))
    return fail(GFX_E_INVALID, "gfx_inject_errors: bad argument");
  const uint32_t thresh = (uint32_t)(rate * 2147483648.0);   // 31-bit uniform
  k_inject_errors<<<grid_for(n * ((n_snps + 15) / 16), 256, device_sms(), 8), 256, 0, (cudaStream_t)stream>>>(packed_dev, n, n_snps, wpr,
                                                                                                            thresh, seed);
  GFX_LAUNCHED();
  return GFX_OK;
}

// Evaluation of a correction run (simulate.py:426-515): bit-parallel comparison of truth / observed / corrected.
// out[0] errors (observed != truth)  [1] fixed (error, corrected == truth)  [2] blanked (error, corrected missing)
// out[3] wrong fix (error, changed, neither)  [4] new errors (no error, changed)  [5] missed (error, unchanged)  [6] cells
__global__ void __launch_bounds__(256) k_confusion(const uint32_t* __restrict__ truth, const uint32_t* __restrict__ obs,
                                                   const uint32_t* __restrict__ fixed, int64_t n, int64_t n_snps, int64_t wpr,
                                                   unsigned long long* __restrict__ out) {
  const int64_t nwords = (n_snps + 15) / 16;
  const int64_t total = n * nwords;
  unsigned int c[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nwords, w = i - r * nwords;
    const uint32_t t = truth[r * wpr + w], o = obs[r * wpr + w], k = fixed[r * wpr + w];
    uint32_t valid = 0x55555555u;
    const int64_t rem = n_snps - w * 16;
    if (rem < 16) valid = ((1u << (2 * rem)) - 1u) & 0x55555555u;
    // per-cell "differs" masks at the even bit positions
    const uint32_t d_to = ((t ^ o) | ((t ^ o) >> 1)) & valid, d_ok = ((o ^ k) | ((o ^ k) >> 1)) & valid;
    const uint32_t d_tk = ((t ^ k) | ((t ^ k) >> 1)) & valid, kmiss = k & (k >> 1) & valid;
    c[0] += __popc(d_to);
    c[1] += __popc(d_to & ~d_tk);
    c[2] += __popc(d_to & d_ok & kmiss);
    c[3] += __popc(d_to & d_ok & d_tk & ~kmiss);
    c[4] += __popc(~d_to & valid & d_ok);
    c[5] += __popc(d_to & ~d_ok);
    c[6] += __popc(valid);
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    unsigned int v = c[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + q, (unsigned long long)v);
  }
}
extern "C" int gfx_confusion(const uint32_t* truth_dev, const uint32_t* obs_dev, const uint32_t* fixed_dev, int64_t n, int64_t n_snps,
                             int64_t wpr, uint64_t* out8_dev, void* stream) {
  if (!truth_dev || !obs_dev || !fixed_dev || !out8_dev || n <= 0 || n_snps <= 0 || wpr < (n_snps + 15) / 16)
    return fail(GFX_E_INVALID, "gfx_confusion: bad argument");
  GFX_CUDA(cudaMemsetAsync(out8_dev, 0, 8 * sizeof(uint64_t), (cudaStream_t)stream));
  k_confusion<<<grid_for(n * ((n_snps + 15) / 16), 256, device_sms(), 4), 256, 0, (cudaStream_t)stream>>>(
      truth_dev, obs_dev, fixed_dev, n, n_snps, wpr, (unsigned long long*)out8_dev);
  GFX_LAUNCHED();
  return GFX_OK;
}

extern "C" int gfx_founders(const gfx_pedigree_desc* ped, const uint8_t* restrict_host, int32_t* out_idx, int32_t* out_n) {
  if (!ped || !out_idx || !out_n) return fail(GFX_E_INVALID, "null argument");
  int n = 0;
  for (int a = 0; a < ped->n_ped; ++a) {
    if (restrict_host && !restrict_host[a]) continue;
    const int32_t s = ped->sire[a], d = ped->dam[a];
    const bool has_s = s >= 0 && (!restrict_host || restrict_host[s]);
    const bool has_d = d >= 0 && (!restrict_host || restrict_host[d]);
    if (!has_s || !has_d) out_idx[n++] = a;
  }
  *out_n = n;
  return GFX_OK;
}

extern "C" int gfx_infer_host(int32_t n, const int8_t* p1, const int8_t* p2, const uint8_t* code, int32_t query, double* out3) {
  if (n <= 0 || n > GFX_MAXN || !p1 || !p2 || !code || !out3 || query < 0 || query >= n)
    return fail(GFX_E_INVALID, "gfx_infer_host: bad argument");
  std::vector<int8_t> last(n);
  for (int i = 0; i < n; ++i) last[i] = (int8_t)i;
  for (int i = 0; i < n; ++i)
    if (p1[i] >= 0) {
      if (p1[i] >= i || p2[i] >= i || p2[i] < 0) return fail(GFX_E_INVALID, "nodes must be topologically ordered");
      last[p1[i]] = (int8_t)i;
      last[p2[i]] = (int8_t)i;
    }
  uint64_t obs = 0;
  for (int i = 0; i < n; ++i)
    if (code[i] < 3 && i != query) obs |= 1ull << i;
  GfxBlanket b{n, p1, p2, last.data()};
  std::vector<double> tab(GFX_TAB_BIG);
  if (gfx_infer(b, query, code, obs, tab.data(), out3, GFX_WMAX_BIG)) return fail(GFX_E_TOO_WIDE, "too many open variables");
  return GFX_OK;
}

extern "C" int gfx_infer_pair_host(int32_t n, const int8_t* p1, const int8_t* p2, const uint8_t* code, int32_t q0, int32_t q1,
                                   double* out6) {
  if (n <= 0 || n > GFX_MAXN || !p1 || !p2 || !code || !out6 || q0 < 0 || q0 >= n || q1 < 0 || q1 >= n || q0 == q1)
    return fail(GFX_E_INVALID, "gfx_infer_pair_host: bad argument");
  std::vector<int8_t> last(n);
  for (int i = 0; i < n; ++i) last[i] = (int8_t)i;
  for (int i = 0; i < n; ++i)
    if (p1[i] >= 0) {
      if (p1[i] >= i || p2[i] >= i || p2[i] < 0) return fail(GFX_E_INVALID, "nodes must be topologically ordered");
      last[p1[i]] = (int8_t)i;
      last[p2[i]] = (int8_t)i;
    }
  uint64_t obs = 0;
  for (int i = 0; i < n; ++i)
    if (code[i] < 3) obs |= 1ull << i;
  GfxBlanket b{n, p1, p2, last.data()};
  std::vector<double> tab(GFX_TAB_BIG);
  if (gfx_infer_pair(b, q0, q1, code, obs, tab.data(), out6, out6 + 3, GFX_WMAX_BIG))
    return fail(GFX_E_TOO_WIDE, "too many open variables");
  return GFX_OK;
}

extern "C" int gfx_kids_term_host(int32_t n, const uint8_t* kid, const uint8_t* par, double* out3) {
  if (n < 0 || !out3) return fail(GFX_E_INVALID, "gfx_kids_term_host: bad argument");
  double kv[108];
  gfx_fill_kv(kv);
  struct K { const uint8_t* p; __host__ __device__ int operator()(int i) const { return p[i]; } };
  K kc{kid}, pc{par};
  if (!gfx_kids_term(n, kc, pc, kv, out3)) { out3[0] = out3[1] = out3[2] = nan(""); return 1; }
  return GFX_OK;
}
