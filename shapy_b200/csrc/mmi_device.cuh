// Device-side triangle/triangle predicates with the exact semantics of the reference operator
// mesh-mesh-intersection/src/mesh_mesh_intersect_cuda_op.cu (quirks included):
//   CMP 91-92, SatCrossEdge 151-169, point_to_barycentric 186-200, ray_triangle_intersect 202-232,
//   TriangleTriangleIsectSepAxis 270-341, checkOverlap 362-373,
//   find_triangle_triangle_intersection_points 375-518.
// Translation units including this header are compiled with -fmad=false so that the fp32 results
// are identical to the plain-C oracle (oracle/mmi_oracle.c).
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace shapy {
namespace mmi {

struct Tri {
  float3 v0, v1, v2;
};

__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 mul3(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ bool cmp0(float x) {  // CMP(x, 0)
  return fabsf(x) <= FLT_EPSILON * fmaxf(1.0f, fabsf(x));
}

__device__ __forceinline__ float3 sat_cross_edge(float3 a, float3 b, float3 c, float3 d) {
  float3 ab = sub3(b, a), cd = sub3(d, c);
  float3 result = cross3(ab, cd);
  if (!cmp0(dot3(ab, cd))) return result;
  float3 axis = cross3(ab, sub3(c, a));
  result = cross3(ab, axis);
  if (!cmp0(dot3(result, result))) return result;
  return make_float3(0.f, 0.f, 0.f);
}

__device__ __forceinline__ bool axis_separates(float3 ax, const Tri &t1, const Tri &t2) {
  float p = dot3(ax, t1.v0), l1 = p, h1 = p;
  p = dot3(ax, t1.v1); l1 = fminf(l1, p); h1 = fmaxf(h1, p);
  p = dot3(ax, t1.v2); l1 = fminf(l1, p); h1 = fmaxf(h1, p);
  p = dot3(ax, t2.v0);
  float l2 = p, h2 = p;
  p = dot3(ax, t2.v1); l2 = fminf(l2, p); h2 = fmaxf(h2, p);
  p = dot3(ax, t2.v2); l2 = fminf(l2, p); h2 = fmaxf(h2, p);
  bool overlap = (l1 <= h2) && (l2 <= h1);
  return !overlap && !cmp0(dot3(ax, ax));
}

__device__ inline bool sat11(const Tri &t1, const Tri &t2) {
  if (axis_separates(sat_cross_edge(t1.v0, t1.v1, t1.v1, t1.v2), t1, t2)) return false;
  if (axis_separates(sat_cross_edge(t2.v0, t2.v1, t2.v1, t2.v2), t1, t2)) return false;
  const float3 a0[3] = {t1.v0, t1.v1, t1.v2}, a1[3] = {t1.v1, t1.v2, t1.v0};
  const float3 b0[3] = {t2.v0, t2.v1, t2.v2}, b1[3] = {t2.v1, t2.v2, t2.v0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (axis_separates(sat_cross_edge(a0[i], a1[i], b0[j], b1[j]), t1, t2)) return false;
  return true;
}

__device__ __forceinline__ bool ray_tri(float3 orig, float3 dir, float3 v0, float3 v1, float3 v2, float &t, float3 &p) {
  float3 v0v1 = sub3(v1, v0), v0v2 = sub3(v2, v0);
  float3 pvec = cross3(dir, v0v2);
  float det = dot3(v0v1, pvec);
  if (fabs((double)det) < 1e-4) return false;
  float inv = 1 / det;
  float3 tvec = sub3(orig, v0);
  float u = dot3(tvec, pvec) * inv;
  if (u < 0 || u > 1) return false;
  float3 qvec = cross3(tvec, v0v1);
  float v = dot3(dir, qvec) * inv;
  if (v < 0 || u + v > 1) return false;
  t = dot3(v0v2, qvec) * inv;
  p = add3(mul3(t, dir), orig);
  return true;
}

__device__ __forceinline__ float3 to_bary(float3 p, float3 a, float3 b, float3 c) {
  float3 v0 = sub3(b, a), v1 = sub3(c, a), v2 = sub3(p, a);
  float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
  float denom = d00 * d11 - d01 * d01;
  float y = (d11 * d20 - d01 * d21) / denom;
  float z = (d00 * d21 - d01 * d20) / denom;
  return make_float3((float)(1.0 - y - z), y, z);
}

// Returns true when a first intersection point exists; bc1/bc2 written only then.
__device__ inline bool isect_points(const Tri &Q, const Tri &T, float3 &bc1, float3 &bc2) {
  const float3 qe[3] = {sub3(Q.v1, Q.v0), sub3(Q.v2, Q.v1), sub3(Q.v0, Q.v2)};
  const float3 qo[3] = {Q.v0, Q.v1, Q.v2};
  const float3 te[3] = {sub3(T.v1, T.v0), sub3(T.v2, T.v1), sub3(T.v0, T.v2)};
  const float3 to[3] = {T.v0, T.v1, T.v2};
  float tmin = FLT_MAX;
  bool found_first = false, found_second = false;
  float3 ip = make_float3(0, 0, 0), ip1 = ip, ip2 = ip;
  float t = 0.0f;  // uninitialised in the reference; 0 / NaN give identical results
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    bool hit = ray_tri(qo[i], qe[i], T.v0, T.v1, T.v2, t, ip);
    if (t > 1 || t < 0) continue;
    if (hit && !found_first) { ip1 = ip; found_first = true; tmin = t; }
    float s = (float)((double)t + 1e-4);
    hit = ray_tri(add3(qo[i], mul3(s, qe[i])), qe[i], T.v1, T.v1, T.v2, t, ip2);  // degenerate: never hits
    if (t > 1 || t < 0) continue;
    if (hit && found_first && t > tmin && !found_second) { ip2 = ip; found_second = true; }
  }
  if (found_first) bc1 = to_bary(ip1, T.v0, T.v1, T.v2);
  if (found_second) { bc2 = to_bary(ip2, T.v0, T.v1, T.v2); return true; }
  tmin = FLT_MAX;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    bool hit = ray_tri(to[i], te[i], Q.v0, Q.v1, Q.v2, t, ip);
    if (t > 1 || t < 0) continue;
    if (hit && !found_first) { ip1 = ip; tmin = t; found_first = true; }
    float s = (float)((double)t + 1e-4);
    hit = ray_tri(add3(to[i], mul3(s, te[i])), te[i], Q.v0, Q.v1, Q.v2, t, ip);
    if (t > 1 || t < 0) continue;
    if (hit && found_first && t > tmin && !found_second) { ip2 = ip; found_second = true; }
  }
  if (found_first) bc1 = to_bary(ip1, T.v0, T.v1, T.v2);
  if (found_second) { bc2 = to_bary(ip2, T.v0, T.v1, T.v2); return true; }
  if (found_first) bc2 = bc1;
  return found_first;
}

struct Box {
  float3 lo, hi;
};
__device__ __forceinline__ Box tri_box(const Tri &t) {
  Box b;
  b.lo = make_float3(fminf(t.v0.x, fminf(t.v1.x, t.v2.x)), fminf(t.v0.y, fminf(t.v1.y, t.v2.y)),
                     fminf(t.v0.z, fminf(t.v1.z, t.v2.z)));
  b.hi = make_float3(fmaxf(t.v0.x, fmaxf(t.v1.x, t.v2.x)), fmaxf(t.v0.y, fmaxf(t.v1.y, t.v2.y)),
                     fmaxf(t.v0.z, fmaxf(t.v1.z, t.v2.z)));
  return b;
}
__device__ __forceinline__ bool box_overlap(const Box &a, const Box &b) {
  return (a.lo.x <= b.hi.x) && (a.hi.x >= b.lo.x) && (a.lo.y <= b.hi.y) && (a.hi.y >= b.lo.y) &&
         (a.lo.z <= b.hi.z) && (a.hi.z >= b.lo.z);
}

}  // namespace mmi
}  // namespace shapy
