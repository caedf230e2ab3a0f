/* A plain C program that drives libblangpt.so through include/blangpt.h alone -- no Python, no torch:
 * create -> set_model -> initialize -> run_scans -> get_round_stats -> adapt -> perform_inference -> destroy,
 * i.e. the calls the Java shim (java/blang/engines/gpu/GpuPT.java) makes, on BASELINE config 0's shape (Normal
 * model, 8 chains).  tests/test_c_abi_driver.py compiles it with gcc and runs it: on a box without a GPU it must
 * fail in blangpt_create with the "no CPU fallback" message (exit code 3), on a B200 it prints the round summary.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "blangpt.h"

#define CHECK(h, call) do { int rc_ = (call); if (rc_ != BLANGPT_OK) { \
    fprintf(stderr, "%s -> status %d: %s\n", #call, rc_, blangpt_last_error(h)); return rc_ == BLANGPT_ECUDA ? 3 : 2; } } while (0)

int main(void) {
  enum { N_OBS = 1000, N_CHAINS = 8, N_SCANS = 40 };
  static double y[N_OBS];
  unsigned long long s = 20240001ULL;                      /* a small LCG: the data only need to be fixed */
  for (int i = 0; i < N_OBS; i++) {
    double u = 0.0;
    for (int k = 0; k < 12; k++) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; u += (double)(s >> 11) / 9007199254740992.0; }
    y[i] = 1.5 + 2.0 * (u - 6.0);                          /* ~ N(1.5, 4) */
  }
  blangpt_config cfg;
  blangpt_default_config(&cfg);
  cfg.nChains = N_CHAINS; cfg.random = 7;
  blangpt_handle* h = NULL;
  CHECK(NULL, blangpt_create(&cfg, &h));
  blangpt_array data = {y, N_OBS, BLANGPT_F64, N_OBS, 1};
  const double hyper[3] = {0.0, 100.0, 0.1};               /* mu ~ Normal(0, 100), sigma2 ~ Exponential(0.1) */
  CHECK(h, blangpt_set_model(h, BLANGPT_FAM_NORMAL, &data, 1, hyper, 3));
  CHECK(h, blangpt_initialize(h, BLANGPT_INIT_SCM));

  static double energy[N_SCANS * N_CHAINS], samples[N_SCANS * 2];
  static int32_t indicators[N_SCANS * N_CHAINS];
  blangpt_scan_outputs out;
  memset(&out, 0, sizeof out);
  out.energy = energy; out.swapIndicators = indicators; out.samples = samples;
  CHECK(h, blangpt_run_scans(h, N_SCANS, &out));

  double accept[N_CHAINS], Lambda = 0.0, logz = 0.0;
  int64_t trips = 0;
  blangpt_round_stats st;
  memset(&st, 0, sizeof st);
  st.swapAcceptMean = accept; st.Lambda = &Lambda; st.logZSteppingStone = &logz; st.roundTrips = &trips;
  CHECK(h, blangpt_get_round_stats(h, &st));
  double mean_mu = 0.0;
  for (int i = N_SCANS / 2; i < N_SCANS; i++) mean_mu += samples[2 * i] / (N_SCANS - N_SCANS / 2);
  printf("after %d scans: Lambda %.4f  logZ(stepping stone) %.3f  round trips %lld  mean mu (2nd half) %.3f\n",
         N_SCANS, Lambda, logz, (long long)trips, mean_mu);
  if (!(Lambda >= 0.0) || !isfinite(logz) || fabs(mean_mu - 1.5) > 0.5) { fprintf(stderr, "implausible results\n"); return 4; }

  double ladder[N_CHAINS];
  CHECK(h, blangpt_adapt(h, NULL));
  CHECK(h, blangpt_get_ladder(h, ladder, N_CHAINS));
  if (ladder[0] != 1.0 || ladder[N_CHAINS - 1] != 0.0) { fprintf(stderr, "adapted ladder does not span [0, 1]\n"); return 4; }

  int32_t rn[64], ra[64];
  const int n_rounds = blangpt_rounds(200, 0.5, rn, ra);
  int total = 0;
  for (int r = 0; r < n_rounds; r++) total += rn[r];
  double lam[64], lz[64];
  int64_t rt[64];
  blangpt_inference_summary sum;
  memset(&sum, 0, sizeof sum);
  sum.Lambda = lam; sum.logZ = lz; sum.roundTrips = rt;
  CHECK(h, blangpt_perform_inference(h, 200, 0.5, 1, NULL, &sum));
  printf("performInference(200, 0.5): %d rounds, %d scans in total, final round: Lambda %.4f  logZ %.3f  round trips %lld\n",
         sum.nRounds, total, lam[sum.nRounds - 1], lz[sum.nRounds - 1], (long long)rt[sum.nRounds - 1]);
  uint64_t counters[4];
  CHECK(h, blangpt_counters(h, counters));
  printf("log-density evaluations %llu, scans %llu, kernel launches %llu\n", (unsigned long long)counters[0],
         (unsigned long long)counters[1], (unsigned long long)counters[2]);
  blangpt_destroy(h);
  puts("C ABI driver: OK");
  return 0;
}
