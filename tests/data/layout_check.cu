// Host-only check of the shared-memory layout rules of rs_device.cuh (tests/test_layout_cpu.py compiles and runs it; no GPU):
// alignment, no overlap between regions that are live at the same time, the two documented aliases, the size limits
// behind "nine / ten cells per SM", and FixedShape::layout(algo) == what rs_create computes for the headline cell.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>
#define RS_NS rs
#define RS_THREADS 128
#define RS_MIN_BLOCKS 8
#define RS_CORE_ONLY
#include "rs_device.cuh"

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { ++fails; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

struct Region { const char* name; int off, len; };

static void check_layout(int S, int U, int G, int m_cap, int cq, int ng, int min_n, int nb, bool share, const char* what) {
  const rs::Layout L = rs::make_layout(S, U, G, m_cap, cq, ng, min_n, nb, share);
  const int n = std::max(S * G, min_n), nw = (n + 31) / 32;
  const bool mt_alias = L.mtab == L.posl, den_alias = L.den == L.cnt;
  std::vector<Region> r = {
      {"avg", L.avg, 8 * U * nb}, {"off", L.off, 8 * S}, {"tval", L.tval, 128}, {"posl", L.posl, 2 * n}, {"posr", L.posr, 2 * n},
      {"tx", L.tx, 4 * U * nb}, {"utr", L.utr, 4 * U}, {"mask", L.mask, 8 * U}, {"seg0", L.seg0, 4 * (n / 16 + 2)},
      {"seg1", L.seg1, 4 * (n / 16 + 2)}, {"target", L.target, 4 * S}, {"quota", L.quota, 4 * S}, {"frb", L.frb, 4 * S},
      {"wd", L.wd, 4 * S}, {"misc", L.misc, 4 * 48}, {"a", L.a, 2 * n}, {"win", L.win, 2 * n}, {"cnt", L.cnt, 2 * 16 * nw},
      {"outsl", L.outsl, G}, {"sptr", L.sptr, 4 * (S + 1)}, {"sues", L.sues, 2 * U}, {"cq", L.cq, cq}, {"mbar", L.mbar, 8},
      {"done", L.done, U}};
  if (!mt_alias) r.push_back({"mtab", L.mtab, 8 * rs::kMStride * m_cap});
  if (!den_alias) r.push_back({"den", L.den, 8 * U});
  if (ng) {
    r.push_back({"ng_red", L.ng_red, 16 * (RS_THREADS / 32)});
    r.push_back({"ng_list", L.ng_list, 2 * ng});
    r.push_back({"ng_hc", L.ng_hc, ng});
    r.push_back({"ng_mcs", L.ng_mcs, ng * RS_THREADS});
  }
  for (size_t i = 0; i < r.size(); ++i) {
    CHECK(r[i].off >= 0 && r[i].off + r[i].len <= L.total, "%s: %s [%d, +%d) outside total %d", what, r[i].name, r[i].off, r[i].len, L.total);
    for (size_t j = i + 1; j < r.size(); ++j)
      CHECK(r[i].len == 0 || r[j].len == 0 || r[i].off + r[i].len <= r[j].off || r[j].off + r[j].len <= r[i].off,
            "%s: %s [%d, +%d) overlaps %s [%d, +%d)", what, r[i].name, r[i].off, r[i].len, r[j].name, r[j].off, r[j].len);
  }
  for (int o : {L.avg, L.den, L.off, L.tval, L.mtab, L.mbar}) CHECK(o % 8 == 0, "%s: offset %d of a 64-bit array is not 8-aligned", what, o);
  for (int o : {L.tx, L.utr, L.mask, L.seg0, L.seg1, L.target, L.quota, L.frb, L.wd, L.misc, L.sptr, L.ng_mcs}) CHECK(o % 4 == 0, "%s: offset %d not 4-aligned", what, o);
  CHECK(L.cq % 16 == 0, "%s: the staged CQI must be 16-byte aligned for the bulk copy", what);
  if (mt_alias) CHECK(8 * rs::kMStride * m_cap <= 4 * n && L.posr == L.posl + 2 * n, "%s: the metric table does not fit the slot arrays", what);
  CHECK(den_alias == (share && 8 * U <= 2 * 16 * nw), "%s: den/cnt sharing rule", what);
  CHECK(L.total % 16 == 0, "%s: total %d", what, L.total);
}

int main() {
  // the headline cell, every scheduler id with a compile-time-shape kernel, both CQI layouts
  using U8 = rs::FixedShape<20, 5, 64, 8, 0>;
  using Nib = rs::FixedShape<20, 5, 64, 8, 2>;
  for (int algo : {9, 8, 10, 101, 103, 7, 11}) {
    const bool nvs = algo == 7 || algo == 11;
    // what rs_create computes (rs_sched.cu): chunks of metric_table_cap() UEs for the transport ids, the served slice for NVS
    const int cap = rs::metric_table_cap(5, 100, 20 * 64), per = cap / 5;
    const int m_cap = nvs ? 5 : std::min(20, per) * 5, chunks = nvs ? 1 : (20 + per - 1) / per;
    const int min_n = algo == 10 ? std::max(16 * 64, 1024) : ((algo == 101 || algo == 103) ? (rs::inter_scratch_bytes(64, 20) + 3) / 4 : 0);
    const rs::Layout h0 = rs::make_layout(20, 100, 64, m_cap, 100 * 64, nvs ? 5 : 0, min_n, 1, rs::den_shares_cnt(algo));
    const rs::Layout h2 = rs::make_layout(20, 100, 64, m_cap, 100 * 32, nvs ? 5 : 0, min_n, 1, rs::den_shares_cnt(algo));
    const rs::Layout f0 = U8::layout(algo), f2 = Nib::layout(algo);
    CHECK(memcmp(&h0, &f0, sizeof h0) == 0, "id %d: FixedShape<...,0>::layout differs from the host's (the handle would fall back to the general kernel)", algo);
    CHECK(memcmp(&h2, &f2, sizeof h2) == 0, "id %d: FixedShape<...,2>::layout differs from the host's", algo);
    CHECK(U8::chunks(algo) == chunks && U8::mcap(algo) == m_cap, "id %d: chunks %d/%d m_cap %d/%d", algo, U8::chunks(algo), chunks, U8::mcap(algo), m_cap);
    check_layout(20, 100, 64, m_cap, 100 * 64, nvs ? 5 : 0, min_n, 1, rs::den_shares_cnt(algo), "headline u8");
    check_layout(20, 100, 64, m_cap, 100 * 32, nvs ? 5 : 0, min_n, 1, rs::den_shares_cnt(algo), "headline packed");
    // cells per SM: 233472 B of shared memory per SM, 1024 B reserved per CTA
    const int fit0 = 233472 / (f0.total + 1024), fit2 = 233472 / (f2.total + 1024);
    CHECK(fit0 >= (algo == 10 || nvs ? 9 : 10), "id %d: only %d u8-layout cells fit an SM (%d B each)", algo, fit0, f0.total);
    CHECK(fit2 >= 10, "id %d: only %d packed-layout cells fit an SM (%d B each)", algo, fit2, f2.total);
  }
  // other shapes: small, big slices, many slices, two bearers, no staging, id 11 scratch
  check_layout(5, 10, 64, 10, 0, 0, 0, 1, true, "5x2");
  check_layout(5, 200, 64, 40, 200 * 64, 0, 0, 1, true, "5x40");
  check_layout(50, 100, 64, rs::metric_table_cap(2, 100, 50 * 64), 100 * 32, 0, 0, 1, true, "50x2");
  check_layout(50, 2000, 64, 50, 0, 0, 0, 1, true, "50x40 direct");
  check_layout(64, 192, 64, rs::metric_table_cap(3, 192, 64 * 64), 192 * 64, 0, 0, 2, true, "64x3 two bearers");
  check_layout(20, 100, 64, 5, 100 * 64, 5, 0, 1, true, "id 11");
  check_layout(6, 19, 64, 19, 19 * 64, 0, 1024, 2, false, "id 10 two bearers");
  check_layout(1, 0, 1280, 0, 0, 0, 0, 1, false, "sort test");
  printf(fails ? "%d failure(s)\n" : "layout ok\n", fails);
  return fails ? 1 : 0;
}
