// Shared helpers for the shapy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/shapy_b200.h"

namespace shapy {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define SHAPY_CUDA_TRY(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      shapy::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define SHAPY_LAUNCH_CHECK()                                                              \
  do {                                                                                    \
    shapy::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      shapy::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define SHAPY_REQUIRE(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      shapy::set_error(__VA_ARGS__);                                                      \
      return SHAPY_ERR_ARG;                                                               \
    }                                                                                     \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// RAII-less device buffer owned by long-lived objects (model constants).
template <typename T>
struct DevBuf {
  T *ptr = nullptr;
  size_t n = 0;
  cudaError_t upload(const std::vector<T> &h) {
    n = h.size();
    cudaError_t e = cudaMalloc((void **)&ptr, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpy(ptr, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
    return e;
  }
  cudaError_t alloc(size_t count) {
    n = count;
    return cudaMalloc((void **)&ptr, std::max<size_t>(n, 1) * sizeof(T));
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    n = 0;
  }
};

}  // namespace shapy
