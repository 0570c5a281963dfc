// Layout of the activations the training forward of the neural renderer keeps for its backward (neural_render.cu, nr_train.cu).
#pragma once
#include <cstddef>

namespace gnrf {
struct NrTrainPlan {
  size_t t1[8], sh[8], net[8];   // byte offsets per level: LReLU(layer_1) [N][2ci][s^2]; PSU output [N][ci][4s^2]; level output [N][co][4s^2]
  size_t bl, rgb_a, rgb_b;       // forward scratch
  size_t total;
};
NrTrainPlan nr_train_plan(int N, int C, int S, int n_blocks, int min_feat);
}  // namespace gnrf
