// On-GPU input stage: uint8 full images + per-person crop windows -> the normalised fp32 NCHW
// tensor HighResolutionNet.forward consumes.  Replaces the reference's CPU DataLoader work per person
// (read_img's /255, transf_utils.crop with cv2.resize, ToTensor, Normalize) and lets the host ship 1 byte per
// channel (4x less PCIe / NVLink traffic than fp32 crops).  HBM-bound: reads <= the window's bytes, writes
// 3 * size^2 * 4 bytes per person.  One thread per output pixel (3 channels), consecutive threads = consecutive x.
#include "common.cuh"
#include "preprocess.cuh"

namespace shapy {

struct PreDesc {   // == shapy_image_desc_t (include/shapy_b200.h)
  long long offset;
  int height, width;
  int ul_x, ul_y, br_x, br_y;
};

__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t *__restrict__ images, const PreDesc *__restrict__ descs, int B,
                                                         int size, float m0, float m1, float m2, float s0, float s1, float s2,
                                                         float *__restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || i >= size * size) return;
  const PreDesc d = descs[b];
  const int y = i / size, x = i - y * size;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  float v[3];
  preprocess_pixel(images + d.offset, d.height, d.width, d.ul_x, d.ul_y, d.br_x, d.br_y, size, y, x, mean, stdv, v);
  float *o = out + (size_t)b * 3 * size * size + i;
  o[0] = v[0];
  o[(size_t)size * size] = v[1];
  o[(size_t)2 * size * size] = v[2];
}

}  // namespace shapy

using namespace shapy;

extern "C" int shapy_preprocess_forward(const unsigned char *images, const shapy_image_desc_t *descs, int B, int size,
                                        const float *mean, const float *stdv, float *out, void *stream) {
  SHAPY_REQUIRE(images && descs && mean && stdv && out, "shapy_preprocess_forward: null argument");
  SHAPY_REQUIRE(B > 0 && size > 0 && B <= 65535, "shapy_preprocess_forward: bad sizes");
  static_assert(sizeof(PreDesc) == sizeof(shapy_image_desc_t), "descriptor layouts must match");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(size * size, 256), B);
  preprocess_kernel<<<grid, 256, 0, st>>>(images, reinterpret_cast<const PreDesc *>(descs), B, size, mean[0], mean[1], mean[2],
                                          stdv[0], stdv[1], stdv[2], out);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}
