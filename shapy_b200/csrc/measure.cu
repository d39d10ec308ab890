// Virtual measurements (mass, height, chest / waist / hip circumference) in ONE launch, one CTA per body.
//
// Replaces, for the fixed query the reference issues (a 2-triangle horizontal quad per measurement):
//   BodyMeasurements.forward           mesh-mesh-intersection/body_measurements/body_measurements.py:217-246
//   3 x (LBVH build + traversal)       mesh-mesh-intersection/src/mesh_mesh_intersect_cuda_op.cu:969-1079
//   3 x scipy.spatial.ConvexHull (CPU) body_measurements.py:160-173
// A plane touches O(sqrt(F)) triangles, so one streaming pass over the faces (read v_shaped once,
// 125 KB / body) beats building a 3 MB BVH three times: every face is tested against the query quad of
// each plane with exactly the reference's predicates (inclusive AABB test, 11-axis SAT with its
// fallback axes, "first ray hit" point selection -- see mmi_device.cuh), hits are appended to a
// shared-memory point list, and a warp per plane gift-wraps the 2-D hull (exact fp64 orientation
// predicate on the fp32 points, as qhull sees them) and sums the 3-D edge lengths in fp32.
// Compiled with -fmad=false (bit parity of the predicates with oracle/mmi_oracle.c).
#include "common.cuh"
#include "mmi_device.cuh"

namespace shapy {

constexpr int kMaxPts = 512;
constexpr int kMaxCand = 4096;  // (face, plane, query triangle) candidates that pass the AABB filter, per body

struct MeasureArgs {
  const float *verts;  // (B,V,3) or null
  const int *faces;    // (F,3)
  const float *tris;   // (B,F,3,3) or null
  int B, V, F;
  shapy_measure_landmarks_t lm;
  float *out;          // (B,5)
  float *pts_out;      // (B,3,maxp,3) or null
  int *cnt_out;        // (B,3) or null
  int maxp;
  int *status;
};

__device__ __forceinline__ mmi::Tri load_tri(const MeasureArgs &a, int b, int f) {
  mmi::Tri t;
  if (a.tris) {
    const float *p = a.tris + ((size_t)b * a.F + f) * 9;
    t.v0 = make_float3(p[0], p[1], p[2]);
    t.v1 = make_float3(p[3], p[4], p[5]);
    t.v2 = make_float3(p[6], p[7], p[8]);
  } else {
    const int *fi = a.faces + 3 * (size_t)f;
    const float *vb = a.verts + (size_t)b * a.V * 3;
    int i0 = fi[0], i1 = fi[1], i2 = fi[2];
    t.v0 = make_float3(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]);
    t.v1 = make_float3(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]);
    t.v2 = make_float3(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]);
  }
  return t;
}

__device__ __forceinline__ bool better_next(double cx, double cz, float ax, float az, float bx, float bz) {
  // true when candidate b should replace a as the next hull vertex after c (counter-clockwise walk):
  // b is strictly to the right of c->a, or collinear and farther.
  double ux = (double)ax - cx, uz = (double)az - cz, wx = (double)bx - cx, wz = (double)bz - cz;
  double o = ux * wz - uz * wx;
  if (o < 0) return true;
  if (o > 0) return false;
  return (wx * wx + wz * wz) > (ux * ux + uz * uz);
}

constexpr int kMeasThreads = 512;
__global__ void __launch_bounds__(kMeasThreads) measure_kernel(MeasureArgs a) {
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ float pts[3][kMaxPts][3];
  __shared__ int cnt[3];
  __shared__ float hs[3];
  __shared__ float lmy[2];
  __shared__ float red[kMeasThreads / 32];
  __shared__ int cand[kMaxCand];
  __shared__ int ncand;
  if (t < 3) cnt[t] = 0;
  if (t == 0) ncand = 0;
  if (t < 5) {
    mmi::Tri tr = load_tri(a, b, a.lm.face_idx[t]);
    // (tri * bc.reshape(3,1)).sum(0): v0*bc0 + v1*bc1 + v2*bc2, y component
    float y = tr.v0.y * a.lm.bc[t][0] + tr.v1.y * a.lm.bc[t][1] + tr.v2.y * a.lm.bc[t][2];
    if (t < 2) lmy[t] = y; else hs[t - 2] = y;
  }
  __syncthreads();
  const float h0 = hs[0], h1 = hs[1], h2 = hs[2];
  float vol = 0.f;
  // ---- phase 1: stream all faces once: signed volume + coarse (AABB) filter against the three planes.
  // Survivors go to a shared-memory queue so that the expensive exact predicates (phase 2) run with one
  // candidate per thread instead of one diverged lane per warp.
  for (int f = t; f < a.F; f += blockDim.x) {
    mmi::Tri T = load_tri(a, b, f);
    // compute_mass, body_measurements.py:207-214
    vol += -T.v2.x * T.v1.y * T.v0.z + T.v1.x * T.v2.y * T.v0.z + T.v2.x * T.v0.y * T.v1.z -
           T.v0.x * T.v2.y * T.v1.z - T.v1.x * T.v0.y * T.v2.z + T.v0.x * T.v1.y * T.v2.z;
    mmi::Box tb = mmi::tri_box(T);
    bool xz = (-1.f <= tb.hi.x) && (1.f >= tb.lo.x) && (-1.f <= tb.hi.z) && (1.f >= tb.lo.z);
    if (!xz || f == 0) continue;  // python keeps collision_faces > 0 only (body_measurements.py:161)
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float h = p == 0 ? h0 : (p == 1 ? h1 : h2);
      if (!((h <= tb.hi.y) && (h >= tb.lo.y))) continue;
      int slot = atomicAdd(&ncand, 2);
      if (slot + 1 < kMaxCand) { cand[slot] = (f << 3) | (p << 1); cand[slot + 1] = (f << 3) | (p << 1) | 1; }
    }
  }
  __syncthreads();
  // ---- phase 2: exact SAT + "first ray hit" point selection, one (face, plane, query triangle) per thread
  {
    const int n = min(ncand, kMaxCand);
    if (ncand > kMaxCand && t == 0 && a.status) atomicExch(a.status, 1);
    for (int i = t; i < n; i += blockDim.x) {
      const int c = cand[i], f = c >> 3, p = (c >> 1) & 3, q = c & 1;
      const float h = p == 0 ? h0 : (p == 1 ? h1 : h2);
      mmi::Tri T = load_tri(a, b, f);
      mmi::Tri Q;
      Q.v0 = make_float3(-1.f, h, -1.f);
      Q.v1 = q == 0 ? make_float3(1.f, h, -1.f) : make_float3(1.f, h, 1.f);
      Q.v2 = q == 0 ? make_float3(1.f, h, 1.f) : make_float3(-1.f, h, 1.f);
      if (!mmi::sat11(Q, T)) continue;
      float3 b1 = make_float3(0, 0, 0), b2 = b1;
      mmi::isect_points(Q, T, b1, b2);  // no hit -> barycentrics stay 0 (first-body semantics)
      int slot = atomicAdd(&cnt[p], 1);
      if (slot < kMaxPts) {
        pts[p][slot][0] = T.v0.x * b1.x + T.v1.x * b1.y + T.v2.x * b1.z;
        pts[p][slot][1] = T.v0.y * b1.x + T.v1.y * b1.y + T.v2.y * b1.z;
        pts[p][slot][2] = T.v0.z * b1.x + T.v1.z * b1.y + T.v2.z * b1.z;
      }
    }
  }
  for (int o = 16; o; o >>= 1) vol += __shfl_xor_sync(0xffffffffu, vol, o);
  if (lane == 0) red[warp] = vol;
  __syncthreads();
  if (t == 0) {
    float v = 0.f;
    for (int i = 0; i < kMeasThreads / 32; ++i) v += red[i];
    a.out[(size_t)b * 5 + 0] = fabsf(v) / 6.0f * 985.0f;
    a.out[(size_t)b * 5 + 1] = fabsf(lmy[0] - lmy[1]);
  }
  if (warp < 3) {
    const int p = warp;
    int n = cnt[p];
    if (n > kMaxPts) {
      if (lane == 0 && a.status) atomicExch(a.status, 1);
      n = kMaxPts;
    }
    if (a.cnt_out && lane == 0) a.cnt_out[(size_t)b * 3 + p] = n;
    if (a.pts_out) {
      int m = min(n, a.maxp);
      for (int i = lane; i < m * 3; i += 32)
        a.pts_out[(((size_t)b * 3 + p) * a.maxp) * 3 + i] = pts[p][i / 3][i % 3];
      if (n > a.maxp && lane == 0 && a.status) atomicExch(a.status, 1);
    }
    float perim = 0.f;
    if (n < 3) {
      perim = __int_as_float(0x7fc00000);  // qhull would raise on a degenerate input
    } else {
      // start: lexicographic min (x, z)
      int best = -1;
      float bx = 0.f, bz = 0.f;
      for (int i = lane; i < n; i += 32) {
        float x = pts[p][i][0], z = pts[p][i][2];
        if (best < 0 || x < bx || (x == bx && z < bz)) { best = i; bx = x; bz = z; }
      }
      for (int o = 16; o; o >>= 1) {
        int ob = __shfl_xor_sync(0xffffffffu, best, o);
        float ox = __shfl_xor_sync(0xffffffffu, bx, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
        if (ob >= 0 && (best < 0 || ox < bx || (ox == bx && oz < bz) || (ox == bx && oz == bz && ob < best))) {
          best = ob; bx = ox; bz = oz;
        }
      }
      const int start = best;
      const float sx = bx, sz = bz;
      int cur = start;
      for (int it = 0; it < n + 1; ++it) {
        const float cx = pts[p][cur][0], cz = pts[p][cur][2];
        int nb = -1;
        float nx = 0.f, nz = 0.f;
        for (int i = lane; i < n; i += 32) {
          float x = pts[p][i][0], z = pts[p][i][2];
          if (x == cx && z == cz) continue;  // the point itself and its duplicates
          if (nb < 0 || better_next(cx, cz, nx, nz, x, z)) { nb = i; nx = x; nz = z; }
        }
        for (int o = 16; o; o >>= 1) {
          int ob = __shfl_xor_sync(0xffffffffu, nb, o);
          float ox = __shfl_xor_sync(0xffffffffu, nx, o), oz = __shfl_xor_sync(0xffffffffu, nz, o);
          if (ob >= 0) {
            bool take = nb < 0 || better_next(cx, cz, nx, nz, ox, oz) ||
                        (ox == nx && oz == nz && ob < nb);
            if (take) { nb = ob; nx = ox; nz = oz; }
          }
        }
        if (nb < 0) break;  // all points coincide
        float dx = pts[p][nb][0] - pts[p][cur][0], dy = pts[p][nb][1] - pts[p][cur][1],
              dz = pts[p][nb][2] - pts[p][cur][2];
        perim += sqrtf(dx * dx + dy * dy + dz * dz);
        cur = nb;
        if (nx == sx && nz == sz) break;
      }
    }
    if (lane == 0) a.out[(size_t)b * 5 + 2 + p] = perim;
  }
}

}  // namespace shapy

using namespace shapy;

static int launch_measure(MeasureArgs a, void *stream) {
  SHAPY_REQUIRE(a.out && a.B > 0 && a.F > 0, "shapy_measure: bad argument");
  for (int i = 0; i < 5; ++i)
    SHAPY_REQUIRE(a.lm.face_idx[i] >= 0 && a.lm.face_idx[i] < a.F, "landmark face %d out of range", a.lm.face_idx[i]);
  if (a.status) SHAPY_CUDA_TRY(cudaMemsetAsync(a.status, 0, sizeof(int), (cudaStream_t)stream));
  measure_kernel<<<a.B, kMeasThreads, 0, (cudaStream_t)stream>>>(a);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

extern "C" int shapy_measure_forward(const float *v_shaped, const int32_t *faces, int batch, int num_verts,
                                     int num_faces, const shapy_measure_landmarks_t *lm, float *out,
                                     float *plane_points, int32_t *plane_counts, int max_points, int32_t *status,
                                     void *stream) {
  SHAPY_REQUIRE(v_shaped && faces && lm, "shapy_measure_forward: null argument");
  MeasureArgs a{v_shaped, faces, nullptr, batch, num_verts, num_faces, *lm, out, plane_points, plane_counts,
                max_points, status};
  return launch_measure(a, stream);
}

extern "C" int shapy_measure_forward_tris(const float *triangles, int batch, int num_faces,
                                          const shapy_measure_landmarks_t *lm, float *out, float *plane_points,
                                          int32_t *plane_counts, int max_points, int32_t *status, void *stream) {
  SHAPY_REQUIRE(triangles && lm, "shapy_measure_forward_tris: null argument");
  MeasureArgs a{nullptr, nullptr, triangles, batch, 0, num_faces, *lm, out, plane_points, plane_counts, max_points,
                status};
  return launch_measure(a, stream);
}
