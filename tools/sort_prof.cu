// Development aid: per-level clock64() split of the std::sort emulation alone (rs_sort_test_kernel built with
// -DRS_SORT_TIMING), n = 1280 entries, keys = max of 5 draws from the CQI trace histogram (the headline cell's winners).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRS_SORT_TIMING -I radiosaber_b200/csrc -o build/sort_prof tools/sort_prof.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define RS_NS rs
#define RS_THREADS 128
#define RS_MIN_BLOCKS 8
#include "rs_device.cuh"

static void build_eq(int nmax, std::vector<unsigned short>* tab) {   // rs_sched.cu build_eq_table
  tab->assign(nmax >= 17 ? rs::eq_offset(nmax + 1) : 1, 0);
  std::vector<unsigned short> a(nmax);
  std::vector<std::pair<int, int>> st;
  for (int len = 17; len <= nmax; ++len) {
    for (int i = 0; i < len; ++i) a[i] = (unsigned short)i;
    st.clear();
    st.emplace_back(0, len);
    while (!st.empty()) {
      int f = st.back().first, l = st.back().second;
      st.pop_back();
      while (l - f > 16) {
        std::swap(a[f], a[f + (l - f) / 2]);
        for (int k = 0; f + 1 + k < l - 1 - k; ++k) std::swap(a[f + 1 + k], a[l - 1 - k]);
        const int cut = f + 1 + (l - f - 1) / 2;
        st.emplace_back(cut, l);
        l = cut;
      }
    }
    unsigned short* out = tab->data() + rs::eq_offset(len);
    const int mid = len / 2;
    for (int i = 0; i < len; ++i) { const int x = a[i]; out[i] = (unsigned short)(x == 0 ? mid : (x == mid ? 0 : x)); }
  }
}

int main(int argc, char** argv) {
  const int n = 1280;
  static const unsigned long long hist[15] = {19075, 7082, 33860, 261099, 438688, 199446, 518174, 661977, 237928, 861279, 596358, 355319, 447453, 12000, 153462};
  unsigned long long tot = 0, cdf[15];
  for (int i = 0; i < 15; ++i) { tot += hist[i]; cdf[i] = tot; }
  std::vector<unsigned short> eq;
  build_eq(n, &eq);
  unsigned short* de;
  cudaMalloc(&de, eq.size() * 2);
  cudaMemcpy(de, eq.data(), eq.size() * 2, cudaMemcpyHostToDevice);
  const rs::Layout L = rs::make_layout(1, 0, n, 0);
  cudaFuncSetAttribute(rs::rs_sort_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  for (int cells : {148, 1184, 4736}) {
    std::vector<unsigned char> keys((size_t)cells * n);
    srand(1);
    for (auto& k : keys) {
      int best = 0;
      for (int d = 0; d < 5; ++d) {
        const unsigned long long r = (((unsigned long long)rand() << 31) ^ rand()) % tot;
        int c = 0;
        while (cdf[c] <= r) ++c;
        best = std::max(best, c + 1);
      }
      k = (unsigned char)best;
    }
    unsigned char* dk; int* dp;
    cudaMalloc(&dk, keys.size());
    cudaMalloc(&dp, keys.size() * 4);
    cudaMemcpy(dk, keys.data(), keys.size(), cudaMemcpyHostToDevice);
    int depth = 0; for (int m = n; m > 1; m >>= 1) depth++; depth *= 2;
    const int reps = 10;
    rs::rs_sort_test_kernel<<<cells, rs::kThreads, L.total>>>(dk, n, depth, dp, de, n, L);
    long long zero[64] = {0};
#ifdef RS_SORT_TIMING
    cudaMemcpyToSymbol(rs::g_sort_prof, zero, sizeof zero);
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) rs::rs_sort_test_kernel<<<cells, rs::kThreads, L.total>>>(dk, n, depth, dp, de, n, L);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long prof[64] = {0};
#ifdef RS_SORT_TIMING
    cudaMemcpyFromSymbol(prof, rs::g_sort_prof, sizeof prof);
#endif
    printf("cells %d: %s, %.1f us per launch; CTA 0, cycles per sort:\n  init %lld\n", cells, cudaGetErrorString(e), ms * 1e3 / reps, prof[0] / reps);
    long long sum = prof[0];
    for (int lv = 0; lv <= 12; ++lv) {
      const long long* p = prof + 1 + 4 * lv;
      if (p[0] + p[1] + p[2] + p[3] == 0) continue;
      printf("  level %2d%s: ranges %6lld | barrier %5lld\n", lv, lv == 12 ? "+" : " ", p[0] / reps, p[3] / reps);
      sum += p[0] + p[1] + p[2] + p[3];
    }
    printf("  exit %lld, counting sort %lld, total %lld\n", prof[60] / reps, prof[61] / reps, (sum + prof[60] + prof[61]) / reps);
    {   // the real std::sort (same libstdc++ as the reference build) on a few of the arrays
      std::vector<int> got((size_t)cells * n);
      cudaMemcpy(got.data(), dp, got.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, checked = 0;
      for (int c = 0; c < cells; c += 7, ++checked) {
        const unsigned char* k = keys.data() + (size_t)c * n;
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return k[x] > k[y]; });
        for (int i = 0; i < n; ++i) if (idx[i] != got[(size_t)c * n + i]) { ++bad; break; }
      }
      printf("  std::sort check: %d of %d arrays differ\n", bad, checked);
    }
    cudaFree(dk); cudaFree(dp);
  }
  return 0;
}
