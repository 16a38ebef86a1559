// Kernel parameter block of the fused VP-SDE SuperDiff step.
#pragma once
#include "../../include/superdiff_b200.h"

namespace sdb {

struct StepParams {
  const float* x;
  const float* noise;
  const float* s[SD_MAX_MODELS];
  float* logq;
  float* x_out;
  float* weights;
  const float* logp_bias;
  const float* sched;
  const int* step_counter;
  int M, B, D;
  float a, b, sigma, dt;
  int mode, dlogq_mode;
  float temperature, ito_scale;
  float mix_scale;            // dx = -dt*a*x + mix_scale*dt*b*mix + c*noise: 2 for the reverse SDE, 1 for the probability-flow ODE
  const float* dlogq_add;     // optional [B][M], added to the per-model increment before the max-subtraction (ODE: dt * div_i)
};

}  // namespace sdb
