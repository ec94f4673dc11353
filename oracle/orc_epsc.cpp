// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the EPSC ring-sector loop-closure descriptors and their scoring:
//   calculateEPSC   src/core/epscGeneration.cpp:478-520     calculateSEPSC :522-562
//   calculateFEPSC  :591-607                                 calculateDistance :633-660
//   constants       src/include/epscGeneration.h:9-43 (rings 20, sectors 80, min 3 m, max 60 m,
//                   DISTANCE_THRESHOLD 0.75), label map config/label.yaml:187-206 (using_label).
// Quirk Q4 reproduced: the esc/psc counters are unsigned char (wrap at 256) and the quotient
// 100*psc/(1+esc) is narrowed to unsigned char modulo 256; FEPSC truncates 0.4*sepsc + 0.6*epsc.
// atan2f is taken as the correctly rounded float of the double routine (see DESIGN.md numerics).
#include "orc_api.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
const int RINGS = 20, SECTORS = 80;
const double MAX_DIS = 60, MIN_DIS = 3;
const double RING_STEP = (MAX_DIS - MIN_DIS) / RINGS;
const double SECTOR_STEP = 2 * M_PI / SECTORS;

inline bool bin_of(float x, float y, int& ring_id, int& sector_id) {
  double distance = std::sqrt(x * x + y * y);            // float sqrt, widened
  if (distance >= MAX_DIS || distance < MIN_DIS) return false;
  ring_id = (int)std::floor((distance - MIN_DIS) / RING_STEP);
  double angle = M_PI + (double)(float)std::atan2((double)y, (double)x);
  sector_id = (int)std::floor(angle / SECTOR_STEP);
  if (ring_id >= RINGS || ring_id < 0) return false;
  if (sector_id >= SECTORS || sector_id < 0) return false;
  return true;
}
}  // namespace

extern "C" {

// using_map: 256-entry LUT label -> class (config/label.yaml using_label; 0 where absent)
void orc_epsc_describe(const float* corner4, int32_t nc, const float* surf4, int32_t ns,
                       const float* sem4, const uint16_t* sem_label, int32_t nsem, const uint8_t* using_map,
                       uint8_t* epsc, uint8_t* sepsc, uint8_t* fepsc) {
  uint8_t esc[1600], psc[1600];
  memset(esc, 0, sizeof(esc)); memset(psc, 0, sizeof(psc));
  int r, s;
  for (int i = 0; i < nc; i++) if (bin_of(corner4[4 * (size_t)i], corner4[4 * (size_t)i + 1], r, s)) esc[r * SECTORS + s]++;
  for (int i = 0; i < ns; i++) if (bin_of(surf4[4 * (size_t)i], surf4[4 * (size_t)i + 1], r, s)) psc[r * SECTORS + s]++;
  for (int i = 0; i < 1600; i++) epsc[i] = (uint8_t)(100 * psc[i] / (1 + esc[i]));
  memset(esc, 0, sizeof(esc)); memset(psc, 0, sizeof(psc));
  for (int i = 0; i < nsem; i++) {
    if (!bin_of(sem4[4 * (size_t)i], sem4[4 * (size_t)i + 1], r, s)) continue;
    unsigned l = sem_label[i];
    int cls = l < 256 ? using_map[l] : 0;
    if (cls == 40 || cls == 50) psc[r * SECTORS + s]++;
    else if (cls == 81) esc[r * SECTORS + s]++;
  }
  for (int i = 0; i < 1600; i++) sepsc[i] = (uint8_t)(100 * psc[i] / (1 + esc[i]));
  for (int i = 0; i < 1600; i++) fepsc[i] = (uint8_t)(sepsc[i] * 0.4 + epsc[i] * 0.6);
}

// calculateDistance: returns the score; *best_shift = the winning i in [-10, 10) (first minimum wins)
double orc_epsc_distance(const uint8_t* d1, const uint8_t* d2, int32_t* best_shift, int32_t* min_sad) {
  double difference = 1.0;
  int bs = 0, bsad = -1;
  for (int i = -10; i < 10; i++) {
    int match_count = 0;
    for (int p = 0; p < SECTORS; p++) {
      int new_col = p + i;
      if (new_col >= SECTORS) new_col -= SECTORS;
      if (new_col < 0) new_col += SECTORS;
      for (int q = 0; q < RINGS; q++) match_count += std::abs((int)d1[q * SECTORS + p] - (int)d2[q * SECTORS + new_col]);
    }
    double diff_temp = ((double)match_count) / (SECTORS * RINGS * 255);
    if (diff_temp < difference) { difference = diff_temp; bs = i; bsad = match_count; }
  }
  if (best_shift) *best_shift = bs;
  if (min_sad) *min_sad = bsad;
  return 1 - difference;
}

// Loop-detection scoring of every frame q against its history j < q (loopDetection :736-860 with the
// travel gate left to the caller): keeps, per query, the topk candidates with score > 0.75 (score
// descending, then lower j).  idx/score/shift are N x topk (idx = -1 when fewer qualify).
void orc_epsc_score_all(const uint8_t* desc, int32_t N, int32_t topk, int32_t* idx, float* score, int8_t* shift, int32_t n_threads) {
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(dynamic, 8)
  for (int q = 0; q < N; q++) {
    std::vector<std::pair<int, int>> cand;   // (sad, j)
    std::vector<int> sh(q > 0 ? q : 1);
    for (int j = 0; j < q; j++) {
      int bs, sad;
      double sc = orc_epsc_distance(desc + 1600 * (size_t)j, desc + 1600 * (size_t)q, &bs, &sad);
      sh[j] = bs;
      if (sad >= 0 && sc > 0.75) cand.push_back({sad, j});
    }
    std::sort(cand.begin(), cand.end());
    for (int k = 0; k < topk; k++) {
      if (k < (int)cand.size()) {
        idx[(size_t)q * topk + k] = cand[k].second;
        score[(size_t)q * topk + k] = (float)(1.0 - (double)cand[k].first / (SECTORS * RINGS * 255));
        shift[(size_t)q * topk + k] = (int8_t)sh[cand[k].second];
      } else { idx[(size_t)q * topk + k] = -1; score[(size_t)q * topk + k] = 0.f; shift[(size_t)q * topk + k] = 0; }
    }
  }
}

}  // extern "C"
