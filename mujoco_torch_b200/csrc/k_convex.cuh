// k_convex: convex (box / mesh) narrowphase -- plane_convex, sphere_convex, capsule_convex, convex_convex.
#pragma once
#include "k_collision.cuh"

template <typename T>
__device__ void convex_pair(const ModelView<T>& m, const EnvBuf<T>& d, int fn, int g1, int g2, const int* slot) {
  // placeholder until the convex narrowphase lands: mark every slot as separated
  T frame[9] = {0, 0, 1, 0, 1, 0, -1, 0, 0};
  for (int k = 0; k < 4; ++k) write_contact(m, d, slot[k], T(1), V3<T>(0, 0, 0), frame);
}
