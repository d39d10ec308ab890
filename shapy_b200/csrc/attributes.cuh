// B2A attribute head arithmetic, shared by the CUDA kernel (attributes.cu) and the host-compiled
// copy the tests build from the same source (oracle/attributes_host.cpp).
//
// Reference: attributes/attributes/attributes_betas/polynomial.py:21-140 -- Polynomial(input_dim, output_dim, degree=2):
//   features A = [x_i for i in 0..n) ++ [x_i x_j for i <= j, lexicographic]   (itertools.combinations_with_replacement,
//   polynomial.py:55-69), output = Linear(len(A), output_dim)(A) = A W^T + b  (polynomial.py:137-140).
#pragma once
#include <stddef.h>

#ifndef SHAPY_HD
#define SHAPY_HD __host__ __device__ __forceinline__
#endif

namespace shapy {

// number of degree-2 polynomial features of n inputs (no bias term)
SHAPY_HD int b2a_num_features(int n) { return n + n * (n + 1) / 2; }

// output `o` for one body: x (n), W (n_out, n_feat) row-major, bias (n_out)
SHAPY_HD float b2a_output(const float *x, int n, const float *W, const float *bias, int o) {
  const float *w = W + (size_t)o * b2a_num_features(n);
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += x[i] * w[i];
  int k = n;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) acc += (x[i] * x[j]) * w[k++];
  return acc + bias[o];
}

}  // namespace shapy
