/* oracle/unittest_probe.cpp -- runs the reference's own two unit-test programs' functions on their own inputs
 * (unittest/test_tp_algos.cpp:104-121: a 5 RBG x 3 slice CQI matrix with quotas {1,3,1};
 * unittest/test_effective_sinr.cpp:229: seven RBs at 20 dB and one at 8 dB) and prints the results as JSON.
 * The reference records no expected output for them, and test_tp_algos.cpp has its MaximizeCell call commented
 * out (:123); including the files where they lie and calling the functions is how they become golden vectors
 * (tests/golden/unittest_vectors.json, tools/make_unittest_vectors.py).  TEST INFRASTRUCTURE ONLY. */
#define main tp_algos_main
#include "unittest/test_tp_algos.cpp"
#undef main
#define main effective_sinr_main
#include "unittest/test_effective_sinr.cpp"
#undef main

#include <cstdio>

int main() {
  const int nb_rbgs = 5, nb_slices = 3;
  int** flow_cqi = new int*[nb_rbgs];
  for (int i = 0; i < nb_rbgs; ++i) flow_cqi[i] = new int[nb_slices];
  /* the matrix of test_tp_algos.cpp:110-117 */
  flow_cqi[0][0] = 8;
  flow_cqi[1][0] = flow_cqi[2][0] = flow_cqi[3][0] = flow_cqi[4][0] = 2;
  flow_cqi[0][1] = flow_cqi[1][1] = flow_cqi[2][1] = flow_cqi[3][1] = 10;
  flow_cqi[4][1] = 10;
  flow_cqi[0][2] = 9;
  flow_cqi[1][2] = 7;
  flow_cqi[2][2] = flow_cqi[3][2] = flow_cqi[4][2] = 2;
  std::vector<int> quota = {1, 3, 1};
  std::vector<int> mc = MaximizeCell(flow_cqi, quota, nb_rbgs, nb_slices);
  std::vector<int> vg = VogelApproximate(flow_cqi, quota, nb_rbgs, nb_slices);
  std::vector<double> sinr = {20, 20, 20, 20, 20, 20, 20, 8};
  const double eff = GetEesmEffectiveSinr(sinr);
  const int cqi = GetCQIFromSinr(eff);
  const int mcs = GetMCSFromCQI(cqi);
  printf("{\"cqi_matrix\": [");
  for (int i = 0; i < nb_rbgs; ++i)
    printf("%s[%d, %d, %d]", i ? ", " : "", flow_cqi[i][0], flow_cqi[i][1], flow_cqi[i][2]);
  printf("], \"quota\": [1, 3, 1], \"maximize_cell\": [");
  for (size_t i = 0; i < mc.size(); ++i) printf("%s%d", i ? ", " : "", mc[i]);
  printf("], \"vogel_int_cqi\": [");
  for (size_t i = 0; i < vg.size(); ++i) printf("%s%d", i ? ", " : "", vg[i]);
  printf("], \"sinr_db\": [20, 20, 20, 20, 20, 20, 20, 8], \"eesm_effective_sinr\": %.17g, \"cqi\": %d, \"mcs\": %d, \"tbs_8rb\": %d}\n",
         eff, cqi, mcs, GetTBSizeFromMCS(mcs, (int)sinr.size()));
  return 0;
}
