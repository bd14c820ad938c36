// BatchNorm finalize arithmetic shared by the elementwise kernels (deferred finalize in their prologue) and the fused
// grid-synchronous epilogue of the flat convolution: every user derives the same per-channel constants bit for bit.
#pragma once
#include "common.cuh"

namespace cilrs {

struct BnStat {
  float scale, shift, mean, rstd, unbiased_var;
};
CILRS_DEVINL BnStat bn_stat_from_sums(double S0, double S1, double inv_count, double unbias, float eps, float gamma, float beta) {
  const double mean_d = S0 * inv_count;
  double var_d = S1 * inv_count - mean_d * mean_d;
  if (var_d < 0.0) var_d = 0.0;
  BnStat r;
  r.mean = (float)mean_d;
  r.unbiased_var = (float)(var_d * unbias);
  r.rstd = 1.0f / sqrtf((float)var_d + eps);
  r.scale = gamma * r.rstd;
  r.shift = beta - r.mean * r.scale;
  return r;
}
CILRS_DEVINL float bn_bdot_from_sums(double S0, double S1, float mean, float rstd) {
  return (float)((double)rstd * (S1 - (double)mean * S0));
}

}  // namespace cilrs
