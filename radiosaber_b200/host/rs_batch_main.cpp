/* rs_batch -- C++ host program for batch runs: thousands of independent cells of one slice configuration on
 * one GPU through the C ABI of include/rs_sched.h.
 *
 * It takes what the reference's SingleCellWithI scenario takes -- the scheduler id and the JSON slice config
 * (src/scenarios/single-cell-with-interference.h:94-123, 214-248; the scheduler constructors parse the same file,
 * downlink-transport-scheduler.cpp:55-97) -- plus a batch size, and either replays the cqi-traces-noise0 files
 * through per-cell UE->trace mappings (enb-mac-entity.cc:42-56, 160-193) or draws synthetic CQI on the device.
 * Output: one JSON line with the per-slice totals the reference's plotters compute from its logs
 * (plot_throughput.py:26-98) and, with --log-cell, the reference's own log text for one cell of the batch.
 *
 *   rs_batch --algo 9 --config cfg.json --cells 4096 --ttis 1000 [--seed 1]
 *            [--traces DIR --mapping FILE]      replay DIR/ue<id>.log; cell b, UE u replays map[(u + 7 b) % n]
 *            [--trace-rows N]                   lines per trace file (475 in cqi-traces-noise0)
 *            [--log-cell B --log-prefix P]      P.stdout / P.stderr as the reference prints them (synthetic CQI only)
 *
 * Host logic only; every scheduling decision is made by the CUDA kernels behind the ABI (no CPU fallback).
 */
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "rs_sched.h"

namespace {

/* ---- the little JSON the slice configs use: objects, arrays, numbers (strings are skipped over) ---------- */
struct Json {
  enum Kind { Null, Num, Arr, Obj, Str } kind = Null;
  double num = 0;
  std::vector<Json> arr;
  std::map<std::string, Json> obj;
  const Json& operator[](const char* k) const {
    static const Json none;
    auto it = obj.find(k);
    return it == obj.end() ? none : it->second;
  }
  int as_int() const { return (int)num; }
};

struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& text) : s(text) {}
  void ws() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
  std::string str() {
    std::string out;
    ++i;
    while (i < s.size() && s[i] != '"') { if (s[i] == '\\') ++i; out += s[i++]; }
    ++i;
    return out;
  }
  Json value() {
    ws();
    Json v;
    if (i >= s.size()) throw std::runtime_error("config: unexpected end");
    if (s[i] == '{') {
      v.kind = Json::Obj;
      ++i;
      for (ws(); s[i] != '}'; ws()) {
        if (s[i] == ',') { ++i; continue; }
        std::string k = str();
        ws();
        if (s[i] != ':') throw std::runtime_error("config: ':' expected");
        ++i;
        v.obj[k] = value();
      }
      ++i;
    } else if (s[i] == '[') {
      v.kind = Json::Arr;
      ++i;
      for (ws(); s[i] != ']'; ws()) {
        if (s[i] == ',') { ++i; continue; }
        v.arr.push_back(value());
      }
      ++i;
    } else if (s[i] == '"') {
      v.kind = Json::Str;
      str();
    } else {
      char* end = nullptr;
      v.kind = Json::Num;
      v.num = strtod(s.c_str() + i, &end);
      if (end == s.c_str() + i) throw std::runtime_error("config: value expected");
      i = (size_t)(end - s.c_str());
    }
    return v;
  }
};

void check(int rc, const char* what) {
  if (rc != RS_OK) throw std::runtime_error(std::string(what) + ": " + rs_last_error());
}
void cu(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

struct Args {
  int algo = 9, cells = 4096, ttis = 1000, log_cell = -1, trace_rows = 475;   /* 475 lines: enb-mac-entity.cc:181 */
  uint64_t seed = 1;
  std::string config, traces, mapping, log_prefix;
};

}  // namespace

int main(int argc, char** argv) {
  try {
    Args a;
    for (int i = 1; i < argc; ++i) {
      const std::string k = argv[i];
      auto next = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value for " + k); return argv[++i]; };
      if (k == "--algo") a.algo = atoi(next().c_str());
      else if (k == "--config") a.config = next();
      else if (k == "--cells") a.cells = atoi(next().c_str());
      else if (k == "--ttis") a.ttis = atoi(next().c_str());
      else if (k == "--seed") a.seed = strtoull(next().c_str(), nullptr, 10);
      else if (k == "--traces") a.traces = next();
      else if (k == "--mapping") a.mapping = next();
      else if (k == "--trace-rows") a.trace_rows = atoi(next().c_str());
      else if (k == "--log-cell") a.log_cell = atoi(next().c_str());
      else if (k == "--log-prefix") a.log_prefix = next();
      else throw std::runtime_error("unknown argument " + k);
    }
    if (a.config.empty() || a.cells < 1 || a.ttis < 1 || a.trace_rows < 1) {
      fprintf(stderr, "usage: rs_batch --algo ID --config cfg.json --cells B --ttis T [--seed s] [--traces DIR --mapping FILE [--trace-rows N]] "
                      "[--log-cell b --log-prefix P]\n");
      return 2;
    }
    /* ---- the slice config, expanded like the scheduler constructors do --------------------------------- */
    std::ifstream ifs(a.config);
    if (!ifs.is_open()) throw std::runtime_error("Fail to open configuration file.");   /* transport.cpp:58-60 */
    std::stringstream ss;
    ss << ifs.rdbuf();
    const std::string text = ss.str();
    const Json cfgj = Parser(text).value();
    std::vector<double> weight;
    std::vector<int32_t> params, u2s;
    for (const Json& grp : cfgj["slices"].arr)
      for (int j = 0; j < grp["n_slices"].as_int(); ++j) {
        weight.push_back(grp["weight"].num);
        params.push_back(grp["algo_alpha"].as_int());
        params.push_back(grp["algo_beta"].as_int());
        params.push_back(grp["algo_epsilon"].as_int());
        params.push_back(grp["algo_psi"].as_int());
      }
    const std::vector<Json>& ups = cfgj["ues_per_slice"].arr;
    for (size_t s = 0; s < ups.size(); ++s)
      for (int j = 0; j < ups[s].as_int(); ++j) u2s.push_back((int32_t)s);
    const int S = (int)ups.size(), U = (int)u2s.size(), B = a.cells, T = a.ttis;
    if ((int)weight.size() != S || U < 1) throw std::runtime_error("config: slices and ues_per_slice do not match");
    const int R = 512, RBG = 8, G = R / RBG;   /* 100 MHz: bandwidth-manager.cpp:98-102, eesm-effective-sinr.h:82-103 */

    rs_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.algo = a.algo;
    cfg.n_slices = S;
    cfg.n_ues = U;
    cfg.n_rbs = R;
    cfg.rbg_size = RBG;
    cfg.cqi_per_rb = 2;   /* 4 bits per RBG: what a CQI is on the air */
    cfg.data_to_transmit = 100000000;
    cfg.weight = weight.data();
    cfg.params = params.data();
    cfg.ue_to_slice = u2s.data();
    rs_handle* h = nullptr;
    check(rs_create(&cfg, B, 0, &h), "rs_create");
    const int n_draws = rs_rand_draws_per_cell_tti(h);

    /* ---- CQI source ------------------------------------------------------------------------------------ */
    const bool replay = !a.traces.empty();
    const int n_rows = a.trace_rows;
    if (replay) {
      if (a.mapping.empty()) throw std::runtime_error("--traces needs --mapping");
      int32_t n_map = 0;
      check(rs_parse_mapping_file(a.mapping.c_str(), nullptr, 0, &n_map), "mapping");
      if (n_map < 1) throw std::runtime_error("empty mapping file");
      std::vector<int32_t> map(n_map);
      check(rs_parse_mapping_file(a.mapping.c_str(), map.data(), n_map, &n_map), "mapping");
      int n_traces = 0;
      for (int t : map) n_traces = std::max(n_traces, t + 1);
      std::vector<uint8_t> traces((size_t)n_traces * n_rows * R, 10);
      std::vector<char> have(n_traces, 0);
      for (int t : map) {
        if (have[t]) continue;
        have[t] = 1;
        const std::string path = a.traces + "/ue" + std::to_string(t) + ".log";
        check(rs_parse_trace_file(path.c_str(), n_rows, R, traces.data() + (size_t)t * n_rows * R), path.c_str());
      }
      std::vector<int32_t> ue_trace((size_t)B * U);
      for (int b = 0; b < B; ++b)   /* every cell its own mapping (the reference: map[u % n] for its one cell) */
        for (int u = 0; u < U; ++u) ue_trace[(size_t)b * U + u] = map[(size_t)(u + 7 * (int64_t)b) % n_map];
      check(rs_set_traces(h, traces.data(), n_traces, n_rows, ue_trace.data()), "rs_set_traces");
    }

    /* ---- the TTI clock of the reference (simulator.cc:116-126): t += 0.001 in double from the first TTI >= 0.1 s */
    std::vector<double> now(T), dt(T);
    {
      double t = 0.0;
      while (t < 0.1) t = t + 0.001;
      double last = 0.1;
      for (int k = 0; k < T; ++k) { now[k] = t; dt[k] = t - last; last = t; t = t + 0.001; }
    }

    /* ---- run in blocks of TB TTIs with everything resident on the device ------------------------------- */
    const int TB = 16;
    uint8_t* d_cqi = nullptr;
    int32_t* d_draws = nullptr;
    int16_t *d_rbg = nullptr, *d_gue = nullptr, *d_grbg = nullptr;   /* d_g*: id 10's grant list */
    int32_t* d_gn = nullptr;
    int32_t *d_bits = nullptr, *d_tgt = nullptr, *d_quo = nullptr;
    uint8_t* d_fc = nullptr;
    const size_t row = G / 2;
    if (!replay) cu(cudaMalloc(&d_cqi, (size_t)TB * B * U * row), "cudaMalloc");
    if (n_draws > 0) cu(cudaMalloc(&d_draws, sizeof(int32_t) * (size_t)TB * B * n_draws), "cudaMalloc");
    const bool want_log = a.log_cell >= 0 && a.log_cell < B && !a.log_prefix.empty();
    rs_log* lg = nullptr;
    std::vector<int16_t> h_rbg, h_gue, h_grbg;
    std::vector<int32_t> h_bits, h_tgt, h_quo;
    std::vector<uint8_t> h_fc, h_cqi;
    rs_outputs out;
    memset(&out, 0, sizeof out);
    if (want_log) {
      check(rs_log_create(&cfg, &lg), "rs_log_create");
      cu(cudaMalloc(&d_rbg, sizeof(int16_t) * (size_t)TB * B * G), "cudaMalloc");
      cu(cudaMalloc(&d_bits, sizeof(int32_t) * (size_t)TB * B * U), "cudaMalloc");
      cu(cudaMalloc(&d_fc, (size_t)TB * B * U), "cudaMalloc");
      cu(cudaMalloc(&d_tgt, sizeof(int32_t) * (size_t)TB * B * S), "cudaMalloc");
      cu(cudaMalloc(&d_quo, sizeof(int32_t) * (size_t)TB * B * S), "cudaMalloc");
      out.rbg_to_ue = d_rbg; out.tbs_bits = d_bits; out.final_cqi = d_fc; out.slice_target = d_tgt; out.slice_quota = d_quo;
      if (a.algo == 10) {
        cu(cudaMalloc(&d_gn, sizeof(int32_t) * (size_t)TB * B), "cudaMalloc");
        cu(cudaMalloc(&d_gue, sizeof(int16_t) * (size_t)TB * B * 2 * G), "cudaMalloc");
        cu(cudaMalloc(&d_grbg, sizeof(int16_t) * (size_t)TB * B * 2 * G), "cudaMalloc");
        out.alloc_n = d_gn; out.alloc_ue = d_gue; out.alloc_rbg = d_grbg;
        h_gue.resize(2 * (size_t)G); h_grbg.resize(2 * (size_t)G);
      }
      h_rbg.resize(G); h_bits.resize(U); h_fc.resize(U); h_tgt.resize(S); h_quo.resize(S); h_cqi.resize((size_t)U * row);
    }
    std::vector<int32_t> trow(TB);
    double ms_total = 0;   /* host clock around the synchronised scheduling calls (generators and log copies excluded) */
    for (int t0 = 0; t0 < T; t0 += TB) {
      const int n = std::min(TB, T - t0);
      if (!replay) check(rs_synth_cqi(h, a.seed, 0, t0, n, d_cqi), "rs_synth_cqi");
      if (n_draws > 0) check(rs_synth_rand2(h, a.seed, 0, t0, n, d_draws), "rs_synth_rand2");
      check(rs_sync(h), "rs_sync");
      const auto w0 = std::chrono::steady_clock::now();
      if (replay) {
        /* all UEs report in the same TTI, every 40 TTIs from the first (phy/ue-lte-phy.cpp:215-232) */
        for (int k = 0; k < n; ++k) trow[k] = rs_trace_row(now[(t0 + k) - (t0 + k) % 40], n_rows);
        check(rs_run_traces_device(h, n, trow.data(), d_draws, nullptr, 0, dt.data() + t0, want_log ? &out : nullptr, TB), "rs_run_traces_device");
      } else {
        check(rs_run_device(h, n, d_cqi, (int64_t)B * U * row, 1, d_draws, nullptr, 0, dt.data() + t0, want_log ? &out : nullptr, TB), "rs_run_device");
      }
      check(rs_sync(h), "rs_sync");
      ms_total += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
      if (want_log && !replay) {
        const int c = a.log_cell;
        for (int k = 0; k < n; ++k) {
          const size_t tb = (size_t)k * B + c;
          cu(cudaMemcpy(h_cqi.data(), d_cqi + tb * U * row, (size_t)U * row, cudaMemcpyDeviceToHost), "copy");
          cu(cudaMemcpy(h_rbg.data(), d_rbg + tb * G, sizeof(int16_t) * G, cudaMemcpyDeviceToHost), "copy");
          cu(cudaMemcpy(h_bits.data(), d_bits + tb * U, sizeof(int32_t) * U, cudaMemcpyDeviceToHost), "copy");
          cu(cudaMemcpy(h_fc.data(), d_fc + tb * U, U, cudaMemcpyDeviceToHost), "copy");
          cu(cudaMemcpy(h_tgt.data(), d_tgt + tb * S, sizeof(int32_t) * S, cudaMemcpyDeviceToHost), "copy");
          cu(cudaMemcpy(h_quo.data(), d_quo + tb * S, sizeof(int32_t) * S, cudaMemcpyDeviceToHost), "copy");
          /* PacketScheduler::m_ts counts TTIs since the eNB was created: 100 at the first TTI with bearers */
          const uint64_t ts = 100 + (uint64_t)(t0 + k);
          if (a.algo == 10) {
            int32_t n_grants = 0;
            cu(cudaMemcpy(&n_grants, d_gn + tb, sizeof n_grants, cudaMemcpyDeviceToHost), "copy");
            cu(cudaMemcpy(h_gue.data(), d_gue + tb * 2 * G, sizeof(int16_t) * 2 * G, cudaMemcpyDeviceToHost), "copy");
            cu(cudaMemcpy(h_grbg.data(), d_grbg + tb * 2 * G, sizeof(int16_t) * 2 * G, cudaMemcpyDeviceToHost), "copy");
            check(rs_log_tti_grants(lg, ts, h_cqi.data(), n_grants, h_gue.data(), h_grbg.data(), h_bits.data(), h_fc.data(),
                                    h_tgt.data(), h_quo.data()), "rs_log_tti_grants");
          } else {
            check(rs_log_tti(lg, ts, h_cqi.data(), h_rbg.data(), h_bits.data(), h_fc.data(), h_tgt.data(), h_quo.data()), "rs_log_tti");
          }
        }
      }
    }
    std::vector<uint64_t> stats((size_t)4 * S);
    check(rs_get_stats(h, stats.data()), "rs_get_stats");
    printf("{\"algo\": %d, \"cells\": %d, \"ttis\": %d, \"slices\": %d, \"ues\": %d, \"cqi\": \"%s\", \"cell_ttis_per_s\": %.1f, \"slice_bytes\": [",
           a.algo, B, T, S, U, replay ? "trace replay" : "synthetic", (double)B * T / (ms_total * 1e-3));
    for (int s = 0; s < S; ++s) printf("%s%llu", s ? ", " : "", (unsigned long long)stats[s]);
    printf("], \"slice_rbs\": [");
    for (int s = 0; s < S; ++s) printf("%s%llu", s ? ", " : "", (unsigned long long)stats[S + s]);
    printf("], \"slice_mbps_per_cell\": [");
    for (int s = 0; s < S; ++s) printf("%s%.3f", s ? ", " : "", (double)stats[s] * 8 / 1e6 / (T * 1e-3) / B);
    printf("]}\n");
    if (want_log && !replay) {
      FILE* fo = fopen((a.log_prefix + ".stdout").c_str(), "w");
      FILE* fe = fopen((a.log_prefix + ".stderr").c_str(), "w");
      if (!fo || !fe) throw std::runtime_error("cannot write the log files");
      fputs(rs_log_stdout(lg, nullptr), fo);
      fputs(rs_log_stderr(lg, nullptr), fe);
      fclose(fo);
      fclose(fe);
    }
    if (lg) rs_log_destroy(lg);
    rs_destroy(h);
    cudaFree(d_cqi); cudaFree(d_draws); cudaFree(d_rbg); cudaFree(d_bits); cudaFree(d_fc); cudaFree(d_tgt); cudaFree(d_quo); cudaFree(d_gn); cudaFree(d_gue); cudaFree(d_grbg);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "rs_batch: %s\n", e.what());
    return 1;
  }
}
