// Training-side kernels: weighted cross-entropy, backward passes, Keras Adam.
#pragma once
#include "lu_defs.h"

struct TrainState {
  size_t off_loss_acc = 0;     // double[2]: sum(weighted ce), sum(valid)
};

// Keras (TF2 OptimizerV2) Adam, epsilon outside the bias-corrected sqrt (train2D.py:61,93)
struct LuAdam {
  float* p; const float* g; float* m; float* v; float lr_t, b1, b2, eps;
  LU_HD void operator()(int64_t i) const {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
};
