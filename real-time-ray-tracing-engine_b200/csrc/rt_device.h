// rt_device.h — device-side data layout and per-thread functions of the B200 path tracer.
//
// Everything here is a plain inline function of one thread's data (RT_HD = __host__ __device__), so
// the kernels in rt_kernels.cu stay thin and the same bodies can be compiled by a host compiler for
// the unit tests in tests/emu (a TEST build; librt_b200.so contains device code only and has no CPU
// path).  FP32 throughout, except the *_exact functions, which evaluate the reference's FP64
// primitive tests in the reference's operation order without fused multiply-adds.
//
// Reference semantics mirrored (paths relative to the reference's src/):
//   camera ray          core/camera/Camera.cpp:186-230, CameraKernels.cu:61-95
//   sphere / quad hit   scene/objects/Sphere.cpp:101-143, Plane.cpp:78-112
//   constant medium     scene/mediums/ConstantMedium.cpp:25-94
//   slab test           optimization/AABB.cpp:141-164
//   scatter / emit      scene/materials/*.cpp, utils/math/PDF.hpp, ONB.hpp, Vec3Utility.cuh:57-70
//   textures            scene/textures/*.cpp, utils/math/PerlinNoise.hpp:43-79,186-201
//   integrator step     core/camera/Camera.cpp:232-309, CameraKernels.cu:106-202
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
struct float4 {
  float x, y, z, w;
};
struct float2 {
  float x, y;
};
struct int4 {
  int x, y, z, w;
};
struct uint2 {
  uint32_t x, y;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

#define RT_PI_F 3.14159265358979323846f
#define RT_INF_F (__builtin_huge_valf())
#define RT_T_MIN 0.001f /* Camera.cpp:242 */

// ---------------------------------------------------------------------------------------------------
// bit casts
// ---------------------------------------------------------------------------------------------------
RT_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i;
  __builtin_memcpy(&i, &f, 4);
  return i;
#endif
}
RT_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f;
  __builtin_memcpy(&f, &i, 4);
  return f;
#endif
}

// ---------------------------------------------------------------------------------------------------
// float3 math
// ---------------------------------------------------------------------------------------------------
struct f3 {
  float x, y, z;
};
RT_HD f3 F3(float x, float y, float z) {
  f3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
RT_HD f3 F3(float4 v) { return F3(v.x, v.y, v.z); }
RT_HD f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD f3 operator-(f3 a) { return F3(-a.x, -a.y, -a.z); }
RT_HD f3 operator*(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD f3 operator*(float t, f3 a) { return F3(t * a.x, t * a.y, t * a.z); }
RT_HD f3 operator*(f3 a, float t) { return F3(t * a.x, t * a.y, t * a.z); }
RT_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD f3 cross(f3 a, f3 b) { return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_HD float length(f3 a) { return sqrtf(dot(a, a)); }
// Single-instruction reciprocal / square root / reciprocal square root (<= 2 ulp) for the shading code, where
// IEEE rounding buys nothing (the FP32 result is a Monte-Carlo sample); ray-box and ray-primitive tests
// keep the correctly rounded operations.
RT_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
RT_HD float fast_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
RT_HD float fast_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
RT_HD f3 normalize(f3 a) { // Vec3::normalize (Vec3.hpp:141-149): |a| <= 1e-8 -> (1, 0, 0)
  float len2 = dot(a, a);
  if (len2 > 1e-16f)
    return fast_rsqrt(len2) * a;
  return F3(1.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------
// Division by a launch-invariant integer (Granlund & Montgomery round-up method): x / d for x < 2^31 as
// (mulhi(x, mul) + x) >> shift.  Path and pixel indexing divides by the film width, the tile height and
// the owned-pixel count for every path at every bounce; a hardware-less 32-bit division is ~20
// instructions, this is 3.
// ---------------------------------------------------------------------------------------------------
struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv fastdiv_make(uint32_t d) { // d >= 1
  FastDiv f;
  uint32_t l = 0;
  while (l < 32 && ((uint64_t)1 << l) < d)
    l++;
  f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - d)) / d + 1);
  f.shift = l;
  f.d = d;
  return f;
}
RT_HD uint32_t fastdiv(const FastDiv &f, uint32_t x) { // x < 2^31
#if defined(__CUDA_ARCH__)
  return (__umulhi(x, f.mul) + x) >> f.shift;
#else
  return (uint32_t)((((uint64_t)x * f.mul) >> 32) + x) >> f.shift;
#endif
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, keyed by the render seed, counter = (pixel, sample, bounce, stream << 16 | block).
// Replaces the reference's per-pixel curand XORWOW state (CameraKernels.cu:15-25): no state in
// memory, and a sample's random numbers do not depend on launch shape, queue order or GPU count.
// ---------------------------------------------------------------------------------------------------
enum { RT_STREAM_CAMERA = 0, RT_STREAM_SHADE = 1, RT_STREAM_MEDIUM0 = 2 };

RT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct Uniform4 {
  float x, y, z, w;
};

RT_HD Uniform4 philox_uniform4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t stream,
                               uint32_t block) {
  uint32_t c0 = pixel, c1 = sample, c2 = bounce, c3 = (stream << 16) | block;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float s = 1.0f / 16777216.0f; // 24-bit uniforms in [0,1), exact in FP32
  Uniform4 u;
  u.x = (float)(c0 >> 8) * s;
  u.y = (float)(c1 >> 8) * s;
  u.z = (float)(c2 >> 8) * s;
  u.w = (float)(c3 >> 8) * s;
  return u;
}

// ---------------------------------------------------------------------------------------------------
// Device scene layout
// ---------------------------------------------------------------------------------------------------
// BVH4 node: 8 x float4 = 128 B (one L1/L2 line), SoA over the four children.
//   [0] lo.x  [1] hi.x  [2] lo.y  [3] hi.y  [4] lo.z  [5] hi.z  [6] child refs (int bits)  [7] spare
// child ref >= 0: node index;  < 0: leaf, primitive index = ~ref;  RT_EMPTY: unused slot (box inverted).
// RT_EMPTY is the bit pattern of +inf: read as a float, a child ref is a tiny positive denormal (node), a
// NaN or negative number (leaf) or +inf (empty), so max(entry distance, ref-as-float) leaves valid
// children alone and makes an empty slot unreachable for any ray, NaN rays included, at no extra cost.
#define RT_NODE_F4 8
#define RT_EMPTY ((int)0x7f800000)

// Primitive record: 4 x float4 = 64 B, in BVH leaf (Morton) order, world space (instances baked).
//   sphere: [0] c0.xyz, radius       [1] center_dir.xyz, colour.g   [2] material header  [3] colour.b, typemat, id, object
//           [2] = copy of the material record's [0]; for metal / solid textures [2].w = colour.r as well
//   quad:   [0] normal.xyz, D        [1] A.xyz, Q.x                 [2] B.xyz, Q.y       [3] Q.z, typemat, id, object
//           alpha = A . (p - Q), beta = B . (p - Q)   with A = v x w, B = w x u  (Plane.cpp:93-94 rearranged)
//   medium: [0] -1/density, first boundary record, n boundary records, medium index      [3] -, typemat, id, object
// typemat = type << 28 | material index.  id = unified primitive id of include/rt_b200.h.
#define RT_PRIM_F4 4
enum { RT_PT_SPHERE = 0, RT_PT_QUAD = 1, RT_PT_MEDIUM = 2 };

// Material record: 3 x float4.
//   [0] type, texture type (int bits), p0, p1   [1] color A   [2] color B
//   lambertian / isotropic / diffuse_light: solid -> A; checker -> p0 = scale, A = even, B = odd;
//                                           noise -> p0 = scale, p1 = perlin table index (int bits)
//   metal: A = albedo, p0 = fuzz.   dielectric: p0 = refraction index.
#define RT_MAT_F4 3

// Light record: 4 x float4.  quad: [0] corner, shape  [1] u, area  [2] v, D  [3] normal, -  ([4],[5] below)
// sphere: [0] center, shape  [1] radius
#define RT_LIGHT_F4 6

struct DScene {
  const float4 *nodes;
  const float4 *prims;
  const float4 *bprims; // boundary records of constant media (same layout as prims)
  const float4 *mats;
  const float4 *lights;
  const float4 *perlin_grad;        // 256 float4 per table
  const unsigned char *perlin_perm; // 3 x 256 bytes per table (x, y, z)
  const uint32_t *texels;           // image textures, one 0x00BBGGRR word per texel, all images back to back
  int n_prims;
  int n_lights;
  int n_media;
  float bg[3];
};

struct DCamera {
  float center[3];
  float p00c[3]; // pixel00_loc - center (so the direction needs no large cancelling subtraction)
  float du[3], dv[3];
  float disk_u[3], disk_v[3];
  int defocus; // defocus_angle > 0
  int width, height;
};

// ---------------------------------------------------------------------------------------------------
// Ray / hit
// ---------------------------------------------------------------------------------------------------
struct Ray {
  f3 o, d;
  float time;
};

struct Hit {
  float t;
  int prim; // leaf-order primitive index, -1 = miss
};

// Camera::get_ray (Camera.cpp:186-205) with the reference CUDA path's polar disk sampler
// (Vec3Utility.cuh:57-61).  u0 = (jitter x, jitter y, disk r^2, disk angle), u1.x = time.
// Every multiply-add is written as an explicit fmaf: the same camera ray is computed in more than one kernel
// (the first extend launch generates it, the first shade launch re-derives it instead of reading it back from
// memory), and explicit fused operations leave the compiler no contraction choice that could differ between them.
RT_HD float fma_add(float a, float b, float c) { // a * b + c in one rounding
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
RT_HD Ray camera_ray(const DCamera &cam, int i, int j, int s_i, int s_j, float recip_sqrt_spp, Uniform4 u0,
                     Uniform4 u1) {
  float px = fma_add((float)s_i + u0.x, recip_sqrt_spp, -0.5f);
  float py = fma_add((float)s_j + u0.y, recip_sqrt_spp, -0.5f);
  float fx = (float)i + px, fy = (float)j + py;
  // pixel sample - center
  f3 rel = F3(fma_add(fy, cam.dv[0], fma_add(fx, cam.du[0], cam.p00c[0])),
              fma_add(fy, cam.dv[1], fma_add(fx, cam.du[1], cam.p00c[1])),
              fma_add(fy, cam.dv[2], fma_add(fx, cam.du[2], cam.p00c[2])));
  f3 lens = F3(0.f, 0.f, 0.f);
  if (cam.defocus) {
    float r = sqrtf(u0.z);
    float th = 2.0f * RT_PI_F * u0.w;
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(th, &sn, &cs);
#else
    sn = sinf(th);
    cs = cosf(th);
#endif
    float lu = r * cs, lv = r * sn;
    lens = F3(fma_add(lv, cam.disk_v[0], lu * cam.disk_u[0]), fma_add(lv, cam.disk_v[1], lu * cam.disk_u[1]),
              fma_add(lv, cam.disk_v[2], lu * cam.disk_u[2]));
  }
  Ray ray;
  ray.o = F3(cam.center[0], cam.center[1], cam.center[2]) + lens;
  ray.d = rel - lens;
  ray.time = u1.x;
  return ray;
}

// ---------------------------------------------------------------------------------------------------
// Primitive tests (FP32 render path)
// ---------------------------------------------------------------------------------------------------
RT_HD float4 ldg4(const float4 *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
// Two adjacent 16-byte rows (32-byte aligned) in one 256-bit load (sm_100: LDG.E.256).
struct F8 {
  float4 a, b;
};
RT_HD F8 ldg8(const float4 *p) {
  F8 v;
#if defined(__CUDA_ARCH__)
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
      : "l"(p));
#else
  v.a = p[0];
  v.b = p[1];
#endif
  return v;
}
// Read-only loads the compiler may not move: issued where they are written, so that independent fetches of
// a latency-bound kernel go out together instead of being sunk below the first branch that uses one of them.
RT_HD float4 ldg4_now(const float4 *p) {
#if defined(__CUDA_ARCH__)
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return *p;
#endif
}
RT_HD float2 ldg2_now(const float2 *p) {
#if defined(__CUDA_ARCH__)
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#else
  return *p;
#endif
}

// Sphere::hit (Sphere.cpp:101-143): roots (h -+ sqrt(disc)) / a on the open interval (tmin, tmax).
// The discriminant is evaluated as a * (r^2 - |oc - (h/a) d|^2), which is the same quantity as
// h^2 - a c without the catastrophic cancellation of c = |oc|^2 - r^2 in FP32 (radius-1000 ground).
#ifndef RT_FAST_PRIM
#define RT_FAST_PRIM 1
#endif
// 1 / x and sqrt of the primitive tests: the reciprocal instruction + one Newton step (<= 1 ulp) and the approximate
// square root (<= 2 ulp) instead of the IEEE sequences (~10 instructions each, 1.7 times per segment).  Measured on
// B200 (profiles/r02_experiments.md): frames 1.3-1.9 % faster, and the FP32-vs-FP64 primitive mismatch audit of every
// BASELINE config counts the same mismatches as with the IEEE operations (RT_FAST_PRIM=0 keeps those).
// FAST = false (the light-sampling pdf in the shading code, which calls the same tests once per light): IEEE operations.
template <bool FAST> RT_HD float prim_rcp(float a) {
  if (FAST) {
    float r = fast_rcp(a);
    return fmaf(fmaf(-a, r, 1.0f), r, r);
  }
  return 1.0f / a;
}
template <bool FAST> RT_HD float prim_sqrt(float x) { return FAST ? fast_sqrt(x) : sqrtf(x); }
template <bool FAST = (RT_FAST_PRIM != 0)>
RT_HD bool sphere_hit(float4 r0, float4 r1, const Ray &ray, float tmin, float tmax, float &t_out) {
  f3 center = F3(r0.x, r0.y, r0.z) + ray.time * F3(r1.x, r1.y, r1.z);
  float radius = r0.w;
  f3 oc = center - ray.o;
  float a = dot(ray.d, ray.d);
  float h = dot(ray.d, oc);
  float inv_a = prim_rcp<FAST>(a);
  f3 perp = oc - (h * inv_a) * ray.d;
  float disc = a * (radius * radius - dot(perp, perp));
  if (disc < 0.f)
    return false;
  float sq = prim_sqrt<FAST>(disc);
  float root = (h - sq) * inv_a;
  if (!(tmin < root && root <= tmax)) { // closed at tmax: ties are decided by leaf_test, not by the order of the tests
    root = (h + sq) * inv_a;
    if (!(tmin < root && root <= tmax))
      return false;
  }
  t_out = root;
  return true;
}

// The same for a ray that STARTS on this sphere (the segment before ended on it): of the two roots, the one at the
// origin is not a hit.  In the reference's FP64 that root is ~1e-13 and the interval's lower bound 0.001 removes
// it (Camera.cpp:242); in FP32 the origin lies up to a few 1e-5 off a radius-1000 sphere and, at grazing
// directions, the spurious root lands on either side of 0.001 by rounding noise alone.  So the decision is taken
// geometrically: a ray leaving towards the outside (h <= 0) cannot meet a sphere again, a ray entering it meets
// the far root only.
RT_HD bool sphere_hit_from_surface(float4 r0, float4 r1, const Ray &ray, float tmin, float tmax, float &t_out) {
  f3 center = F3(r0.x, r0.y, r0.z) + ray.time * F3(r1.x, r1.y, r1.z);
  float radius = r0.w;
  f3 oc = center - ray.o;
  float h = dot(ray.d, oc);
  if (!(h > 0.f))
    return false;
  float a = dot(ray.d, ray.d);
  float inv_a = prim_rcp<RT_FAST_PRIM != 0>(a);
  f3 perp = oc - (h * inv_a) * ray.d;
  float disc = a * (radius * radius - dot(perp, perp));
  float root = (h + prim_sqrt<RT_FAST_PRIM != 0>(fmaxf(disc, 0.f))) * inv_a; // on the surface |perp| <= radius up to rounding
  if (!(tmin < root && root <= tmax))
    return false;
  t_out = root;
  return true;
}

// Plane::hit (Plane.cpp:78-112): closed interval on t and on the planar coordinates.
template <bool FAST = (RT_FAST_PRIM != 0)>
RT_HD bool quad_hit(float4 r0, float4 r1, float4 r2, float qz, const Ray &ray, float tmin, float tmax,
                    float &t_out) {
  f3 n = F3(r0.x, r0.y, r0.z);
  f3 q = F3(r1.w, r2.w, qz);
  float denom = dot(n, ray.d);
  if (fabsf(denom) < 1e-8f)
    return false;
  // (D - n.o) / denom with the subtraction done on the point
  float t = FAST ? dot(n, q - ray.o) * prim_rcp<true>(denom) : dot(n, q - ray.o) / denom;
  if (!(tmin <= t && t <= tmax))
    return false;
  f3 hp = (ray.o - q) + t * ray.d;
  float alpha = dot(F3(r1.x, r1.y, r1.z), hp);
  float beta = dot(F3(r2.x, r2.y, r2.z), hp);
  if (!(0.f <= alpha && alpha <= 1.f) || !(0.f <= beta && beta <= 1.f))
    return false;
  t_out = t;
  return true;
}

// Closest boundary crossing of a constant medium's convex boundary on the open/closed interval the
// member type uses.
RT_HD bool boundary_hit(const DScene &sc, int first, int count, const Ray &ray, float tmin, float tmax,
                        float &t_out) {
  bool any = false;
  float closest = tmax;
  for (int k = 0; k < count; k++) {
    const float4 *rec = sc.bprims + (size_t)(first + k) * RT_PRIM_F4;
    float4 r0 = ldg4(rec), r1 = ldg4(rec + 1), r3 = ldg4(rec + 3);
    int type = (uint32_t)f2i(r3.y) >> 28;
    float t;
    bool ok;
    if (type == RT_PT_SPHERE)
      ok = sphere_hit(r0, r1, ray, tmin, closest, t);
    else
      ok = quad_hit(r0, r1, ldg4(rec + 2), r3.x, ray, tmin, closest, t);
    if (ok) {
      any = true;
      closest = t;
    }
  }
  t_out = closest;
  return any;
}

// ConstantMedium::hit (ConstantMedium.cpp:25-94).  xi is the segment's uniform for this medium.
RT_HD bool medium_hit(const DScene &sc, float4 r0, const Ray &ray, float tmin, float tmax, float xi,
                      float &t_out) {
  int first = f2i(r0.y), count = f2i(r0.z);
  float t1, t2;
  if (!boundary_hit(sc, first, count, ray, -RT_INF_F, RT_INF_F, t1))
    return false;
  if (!boundary_hit(sc, first, count, ray, t1 + 0.0001f, RT_INF_F, t2))
    return false;
  if (t1 < tmin)
    t1 = tmin;
  if (t2 > tmax)
    t2 = tmax;
  if (t1 >= t2)
    return false;
  if (t1 < 0.f)
    t1 = 0.f;
  float ray_length = length(ray.d);
  float distance_inside = (t2 - t1) * ray_length;
  float hit_distance = r0.x * logf(xi); // -1/density * log(xi)
  if (hit_distance > distance_inside)
    return false;
  t_out = t1 + hit_distance / ray_length;
  return true;
}

// One leaf primitive against the current interval; updates `hit` when closer.
struct RayKey {
  uint64_t seed;
  uint32_t pixel, sample, bounce;
};

// Traversal statistics hooks: empty in the product, counters in the host test build (tests/emu).
#ifndef RT_STAT_NODE
#define RT_STAT_NODE()
#define RT_STAT_LEAF()
#endif

// start_prim: the surface primitive the ray starts on (the previous segment's hit; -1 for camera rays, rays
// scattered inside a medium and the parity hook).  A flat primitive cannot be met again, a sphere only at its
// far root (sphere_hit_from_surface).
// Is a hit of primitive `prim` at distance t better than the best so far?  Closer wins; at exactly the same distance
// (coincident surfaces: the two faces neighbouring boxes share, a quad listed twice) the primitive that comes LATER in
// the scene description wins - what the reference does for planes, whose interval is closed (Plane.cpp:88, later objects
// replace earlier ones at equal t) - decided by the unified primitive id stored in the records.  The rule does not
// depend on the order in which primitives are tested, so every traversal order, every tree builder and a traversal
// shared between lanes (k_tail) name the same primitive.  Ties are rare: the ids are only fetched when one occurs.
RT_HD bool tie_wins(const DScene &sc, int prim, int old_prim) {
  if (old_prim < 0)
    return true;
  const int id_new = f2i(ldg4(sc.prims + (size_t)prim * RT_PRIM_F4 + 3).z);
  const int id_old = f2i(ldg4(sc.prims + (size_t)old_prim * RT_PRIM_F4 + 3).z);
  return id_new > id_old;
}
RT_HD bool closer_hit(const DScene &sc, float t, int prim, const Hit &hit) {
  return t < hit.t || (t == hit.t && tie_wins(sc, prim, hit.prim));
}
// TIES: decide hits at exactly the same distance by closer_hit's order-independent rule (needed where a traversal is
// shared between lanes: k_tail<.., SHARE>, and then in every kernel that traces the same scene).  Without it the last
// primitive TESTED at that distance wins, which depends on the traversal order - as it does in the reference.  The
// rule costs 2 % of the traversal kernels (measured), so scenes that do not need it run without.
template <bool TIES = false>
RT_HD void leaf_test(const DScene &sc, int prim, const Ray &ray, float tmin, Hit &hit, int start_prim,
                     const RayKey &key) {
  RT_STAT_LEAF();
  const float4 *rec = sc.prims + (size_t)prim * RT_PRIM_F4;
#ifndef RT_LEAF_EARLY_LOAD
#define RT_LEAF_EARLY_LOAD 1
#endif
  float4 r0 = ldg4(rec), r3 = ldg4(rec + 3);
#if RT_LEAF_EARLY_LOAD
  float4 r1 = ldg4(rec + 1); // spheres and quads both read it: issued with the other two instead of behind the type test
#endif
  int type = (uint32_t)f2i(r3.y) >> 28;
  float t;
  bool ok;
  if (type == RT_PT_SPHERE) {
#if !RT_LEAF_EARLY_LOAD
    float4 r1 = ldg4(rec + 1);
#endif
    ok = prim == start_prim ? sphere_hit_from_surface(r0, r1, ray, tmin, hit.t, t) : sphere_hit(r0, r1, ray, tmin, hit.t, t);
  } else if (type == RT_PT_QUAD) {
#if !RT_LEAF_EARLY_LOAD
    float4 r1 = ldg4(rec + 1);
#endif
    ok = prim != start_prim && quad_hit(r0, r1, ldg4(rec + 2), r3.x, ray, tmin, hit.t, t);
  } else {
    Uniform4 u = philox_uniform4(key.seed, key.pixel, key.sample, key.bounce,
                                 (uint32_t)(RT_STREAM_MEDIUM0 + f2i(r0.w)), 0);
    ok = medium_hit(sc, r0, ray, tmin, hit.t, u.x, t);
  }
  // the closest hit (the tests accept t <= hit.t); a tie is decided by closer_hit's order-independent rule
  if (ok) {
    if (TIES && __builtin_expect(t == hit.t, 0))
      ok = tie_wins(sc, prim, hit.prim);
    if (ok) {
      hit.t = t;
      hit.prim = prim;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// BVH4 traversal (FP32).  Slab test as AABB::hit (AABB.cpp:141-164) but with the reciprocal
// direction hoisted, NaN-safe min/max, and the far plane widened so that no box the FP64 reference
// would enter is culled (Ize, "Robust BVH ray traversal"; see RayTrav below).
// The stack holds (child ref, entry distance); `stack` is caller-provided storage of RT_STACK entries
// (shared-memory short stack in the kernels, spilling to local memory beyond RT_STACK_SMEM).
// ---------------------------------------------------------------------------------------------------
#define RT_STACK 64 // > 3 pushes per level of a 63-bit Morton tree collapsed to 4-wide nodes; only the first
                    // RT_STACK_SMEM entries are ever touched on typical scenes

struct StackEntry {
  int ref;
  float t;
};

// Per-ray traversal constants.  The slab distances are evaluated as plane * inv - o * inv (one FMA per
// plane); the near / far plane of every axis is picked by the direction's sign bit, which selects the
// node row to load, so no per-axis min/max is needed.
// Why this form never rejects a box the exact test enters (u = 2^-24; checked against rational arithmetic in
// tests/test_device_functions_host.py): with inv = (1 + d1) / d, oi = o inv (1 + d2) and the FMA's own rounding d3,
//     computed t(plane) = (plane - o (1 + d2)) / d * (1 + d1)(1 + d3),
// i.e. the exact distance for an origin moved by at most u |o| on that axis, times 1 +- 2u.  Every box in a node
// is at least two ulps (4u |plane|) larger than the geometry it bounds (to_build_box, rt_flatten.h; unions keep
// that).  If |plane| >= |o| / 4 the padding outweighs the moved origin; otherwise |plane - o| > 3/4 |o| and the
// moved origin is a relative error below 4/3 u of the distance.  Either way both sides are off by at most
// (1 +- 3.34u), so widening the exit distance by 1 + 10u (1.0000006f) keeps entry <= exit whenever it holds
// exactly.  No absolute allowance is needed - an earlier version added 2.5e-7 max|o / d| to every exit distance,
// which let a ray with one tiny direction component (|o / d| huge on that axis) enter every box of the scene.
struct RayTrav {
  f3 inv, oi;          // 1 / d and o / d
  unsigned nx, ny, nz; // row of the near plane: x 0/1, y 2/3, z 4/5
  unsigned fxr, fyr, fzr; // row of the far plane = near row ^ 1, kept by node_visit<.., true> (k_extend: three registers
                          // for three XORs per visit, extend -2.2 ... -3.6 %; k_tail has no registers to spare and
                          // recomputes them - the members are dead there)
};
RT_HD bool sign_bit(float f) { return f2i(f) < 0; }
RT_HD RayTrav trav_from(f3 inv, f3 oi) {
  RayTrav t;
  t.inv = inv;
  t.oi = oi;
  t.nx = sign_bit(inv.x) ? 1 : 0; // 1 / d has the sign of d
  t.ny = sign_bit(inv.y) ? 3 : 2;
  t.nz = sign_bit(inv.z) ? 5 : 4;
  t.fxr = t.nx ^ 1u;
  t.fyr = t.ny ^ 1u;
  t.fzr = t.nz ^ 1u;
  return t;
}
// A zero (or denormal, or NaN) direction component is replaced by +-1e-20 for the box tests only: the planes of that
// axis are then 1e20 x (plane - o) away, i.e. behind or beyond everything unless the origin lies between them - what
// a parallel ray should see - instead of inf - inf = NaN, which constrains nothing.  unit_vector_polar returns exact
// zeros with probability ~2^-22 per draw; such a ray used to walk the whole tree (1645 nodes + 3409 primitives on the
// final scene: 2 ms of one lane, a third of the frames).  Primitive tests use the ray's own direction.
RT_HD float trav_component(float d) { return copysignf(fminf(fmaxf(fabsf(d), 1e-20f), 1e20f), d); }
// 1 / d by the single-instruction reciprocal (relative error <= 2^-23 = 2u): the per-side bound of the argument
// above becomes 1 +- (2u + u + 4/3 u), the two sides together 1 + 8.7u - still inside the 1 + 10u widening - and a
// ray's set-up loses three IEEE divisions.  (The clamp keeps the operand and the result normal numbers.)
RT_HD RayTrav make_trav(f3 o, f3 d) {
  f3 inv = F3(fast_rcp(trav_component(d.x)), fast_rcp(trav_component(d.y)), fast_rcp(trav_component(d.z)));
  return trav_from(inv, F3(o.x * inv.x, o.y * inv.y, o.z * inv.z));
}
RT_HD float fma_sub(float a, float b, float c) { // a * b - c in one rounding
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, -c);
#else
  return fmaf(a, b, -c);
#endif
}

// Visits inner node `node`: tests its four child boxes, returns the nearest hit child in `next` (false
// when no child is hit) and pushes the other hit children.
#ifndef RT_LOAD256
#define RT_LOAD256 0 // measured slower (extend +7.5 % on C2): see below and profiles/r02_experiments.md
#endif
template <class Stack, bool FAR_ROW_REGS = false>
RT_HD bool node_visit(const DScene &sc, int node, const RayTrav &rt, float tmin, float tmax, Stack &stack, int &sp,
                      int &next) {
  RT_STAT_NODE();
  // rows are addressed with one 32-bit index each (node * 8 + row): a single wide multiply-add per load
  const unsigned n = (unsigned)node * RT_NODE_F4;
#if RT_LOAD256
  // experiment: the lo / hi rows of an axis share a 32-byte sector - one 256-bit load per axis (sm_100's LDG.E.256)
  // and one address computation per node instead of two 128-bit loads picked by the sign, the near / far planes then
  // selected in registers (24 FSEL).  4 L1 requests per visit instead of 7 - and 7.5 % slower: the L1 data pipe is
  // not what holds the kernel back, the instruction count (+17 per visit) is
  const float4 *np = sc.nodes + n;
  F8 ax = ldg8(np), ay = ldg8(np + 2), az = ldg8(np + 4);
  float4 cr = ldg4(np + 6);
  const bool sx = rt.nx & 1u, sy = rt.ny & 1u, sz = rt.nz & 1u;
  float4 nrx = sx ? ax.b : ax.a, frx = sx ? ax.a : ax.b, nry = sy ? ay.b : ay.a, fry = sy ? ay.a : ay.b,
         nrz = sz ? az.b : az.a, frz = sz ? az.a : az.b;
#else
  const unsigned fxr = FAR_ROW_REGS ? rt.fxr : rt.nx ^ 1u, fyr = FAR_ROW_REGS ? rt.fyr : rt.ny ^ 1u,
                 fzr = FAR_ROW_REGS ? rt.fzr : rt.nz ^ 1u;
  float4 nrx = ldg4(sc.nodes + (n + rt.nx)), frx = ldg4(sc.nodes + (n + fxr)),
         nry = ldg4(sc.nodes + (n + rt.ny)), fry = ldg4(sc.nodes + (n + fyr)),
         nrz = ldg4(sc.nodes + (n + rt.nz)), frz = ldg4(sc.nodes + (n + fzr)), cr = ldg4(sc.nodes + (n + 6u));
#endif
  float tn[4];
  int cref[4] = {f2i(cr.x), f2i(cr.y), f2i(cr.z), f2i(cr.w)};
  const float nx[4] = {nrx.x, nrx.y, nrx.z, nrx.w}, fx[4] = {frx.x, frx.y, frx.z, frx.w};
  const float ny[4] = {nry.x, nry.y, nry.z, nry.w}, fy[4] = {fry.x, fry.y, fry.z, fry.w};
  const float nz[4] = {nrz.x, nrz.y, nrz.z, nrz.w}, fz[4] = {frz.x, frz.y, frz.z, frz.w};
  const float tmax_wide = tmax * 1.0000006f; // 1 + 10u: see RayTrav
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float tnear = fmaxf(fmaxf(fma_sub(nx[c], rt.inv.x, rt.oi.x), fma_sub(ny[c], rt.inv.y, rt.oi.y)),
                        fmaxf(fma_sub(nz[c], rt.inv.z, rt.oi.z), fmaxf(tmin, i2f(cref[c]))));
    float tfar = fminf(fminf(fma_sub(fx[c], rt.inv.x, rt.oi.x), fma_sub(fy[c], rt.inv.y, rt.oi.y)),
                       fma_sub(fz[c], rt.inv.z, rt.oi.z));
    tfar = fminf(tfar * 1.0000006f, tmax_wide);
    tn[c] = tnear <= tfar ? tnear : RT_INF_F;
  }
  // three comparators bring the nearest child to slot 0; the others are pushed as they lie (sorting them
  // too costs more instructions per visit than the better pop order saves: measured)
#define RT_CSWAP(a, b)                                                                                       \
  if (tn[b] < tn[a]) {                                                                                       \
    float tt = tn[a];                                                                                        \
    tn[a] = tn[b];                                                                                           \
    tn[b] = tt;                                                                                              \
    int ti = cref[a];                                                                                        \
    cref[a] = cref[b];                                                                                       \
    cref[b] = ti;                                                                                            \
  }
  RT_CSWAP(0, 1)
  RT_CSWAP(2, 3)
  RT_CSWAP(0, 2)
#undef RT_CSWAP
  if (Stack::has_fast_push && sp + 3 <= Stack::fast_depth) {
    // the usual case: all three slots are in the fast part of the stack, so the pushes are three
    // predicated stores instead of three divergent branches
#pragma unroll
    for (int c = 3; c >= 1; c--) {
      bool p = tn[c] < RT_INF_F;
      stack.set_fast_if(sp, cref[c], tn[c], p);
      sp += p ? 1 : 0;
    }
  } else {
#pragma unroll
    for (int c = 3; c >= 1; c--)
      if (tn[c] < RT_INF_F) {
        if (sp < RT_STACK) {
          stack.set(sp, cref[c], tn[c]);
          sp++;
        }
      }
  }
  next = cref[0];
  return tn[0] < RT_INF_F;
}

// Pops the next stack entry whose entry distance is still inside the interval.
// `floor`: entries below it are not this traversal's any more (k_tail's end-game sharing hands the oldest - largest -
// pending subtrees of a long traversal to idle lanes).
template <class Stack> RT_HD bool stack_pop(Stack &stack, int &sp, float tmax, int &ref, int floor = 0) {
  while (sp > floor) {
    float t;
    sp--;
    stack.get(sp, ref, t);
    if (t <= tmax)
      return true;
  }
  return false;
}

template <class Stack, bool TIES = false>
RT_HD void traverse(const DScene &sc, const Ray &ray, float tmin, Hit &hit, int skip_prim, const RayKey &key,
                    Stack &stack) {
  RayTrav rt = make_trav(ray.o, ray.d);
  int sp = 0;
  int ref = 0; // root node
  for (;;) {
    if (ref >= 0) {
      if (node_visit(sc, ref, rt, tmin, hit.t, stack, sp, ref))
        continue;
    } else {
      leaf_test<TIES>(sc, ~ref, ray, tmin, hit, skip_prim, key);
    }
    if (!stack_pop(stack, sp, hit.t, ref))
      return;
  }
}

// The inexact FP32 slab test above may only ever ENTER more boxes than the exact one, so a simple
// array stack is enough for host-side use.
struct LocalStack {
  static constexpr bool has_fast_push = false;
  static constexpr int fast_depth = 0;
  StackEntry e[RT_STACK];
  RT_HD void set_fast_if(int i, int ref, float t, bool p) {
    if (p)
      set(i, ref, t);
  }
  RT_HD void set(int i, int ref, float t) {
    e[i].ref = ref;
    e[i].t = t;
  }
  RT_HD void get(int i, int &ref, float &t) const {
    ref = e[i].ref;
    t = e[i].t;
  }
};

// ---------------------------------------------------------------------------------------------------
// Textures and materials
// ---------------------------------------------------------------------------------------------------
// PerlinNoise::noise / turb (PerlinNoise.hpp:43-79,186-201).
RT_HD float perlin_noise(const DScene &sc, int table, f3 p) {
  const float4 *grad = sc.perlin_grad + (size_t)table * 256;
  const unsigned char *perm = sc.perlin_perm + (size_t)table * 768;
  float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  float u = p.x - fx, v = p.y - fy, w = p.z - fz;
  int xi = (int)fx, yi = (int)fy, zi = (int)fz;
  float uu = u * u * (3.f - 2.f * u), vv = v * v * (3.f - 2.f * v), ww = w * w * (3.f - 2.f * w);
  float accum = 0.f;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int k = 0; k < 2; k++) {
        int idx = perm[(xi + i) & 255] ^ perm[256 + ((yi + j) & 255)] ^ perm[512 + ((zi + k) & 255)];
        float4 g = ldg4(grad + idx);
        float wx = u - (float)i, wy = v - (float)j, wz = w - (float)k;
        accum += (i ? uu : 1.f - uu) * (j ? vv : 1.f - vv) * (k ? ww : 1.f - ww) * (g.x * wx + g.y * wy + g.z * wz);
      }
  return accum;
}

RT_HD float perlin_turb(const DScene &sc, int table, f3 p, int depth) {
  float accum = 0.f, weight = 1.f;
  for (int i = 0; i < depth; i++) {
    accum += weight * perlin_noise(sc, table, p);
    weight *= 0.5f;
    p = 2.f * p;
  }
  return fabsf(accum);
}

enum { RT_DTEX_SOLID = 0, RT_DTEX_CHECKER = 1, RT_DTEX_NOISE = 2, RT_DTEX_IMAGE = 3 };

// Surface coordinates of a hit, as the reference's primitives compute them (Sphere.cpp:136-140 from the
// outward unit normal; Plane.cpp:93-102: the planar coordinates).  Only image textures read them.
RT_HD void sphere_uv(f3 outward, float &u, float &v) {
  float theta = acosf(fminf(fmaxf(-outward.y, -1.f), 1.f));
  float phi = atan2f(-outward.z, outward.x) + RT_PI_F;
  u = phi * (0.5f / RT_PI_F);
  v = theta * (1.0f / RT_PI_F);
}

// Texture::value for the material's texture (SolidColorTexture.cpp:8-10, CheckerTexture.cpp:43-54,
// NoiseTexture.cpp:29-30).  None of the reference's textures reads (u, v); the image texture (not in the
// reference, "The Next Week" image_texture::value) does.
RT_HD f3 material_texture(const DScene &sc, const float4 *m, float4 m0, f3 p, float u, float v) {
  int tex = f2i(m0.y);
  if (tex == RT_DTEX_SOLID)
    return F3(ldg4(m + 1));
  if (tex == RT_DTEX_CHECKER) {
    float inv_scale = fast_rcp(m0.z);
    int xi = (int)floorf(inv_scale * p.x), yi = (int)floorf(inv_scale * p.y), zi = (int)floorf(inv_scale * p.z);
    bool even = ((xi + yi + zi) % 2) == 0;
    return F3(ldg4(m + (even ? 1 : 2)));
  }
  if (tex == RT_DTEX_IMAGE) {
    int width = f2i(m0.z), height = f2i(m0.w);
    if (width <= 0 || height <= 0)
      return F3(0.f, 1.f, 1.f);
    float uc = fminf(fmaxf(u, 0.f), 1.f), vc = 1.0f - fminf(fmaxf(v, 0.f), 1.f);
    int i = (int)(uc * (float)width), j = (int)(vc * (float)height);
    i = i > width - 1 ? width - 1 : i;
    j = j > height - 1 ? height - 1 : j;
    uint32_t texel = sc.texels[(size_t)f2i(ldg4(m + 1).x) + (size_t)j * (size_t)width + (size_t)i]; // 0x00BBGGRR
    const float s = 1.0f / 255.0f;
    return F3(s * (float)(texel & 255u), s * (float)((texel >> 8) & 255u), s * (float)((texel >> 16) & 255u));
  }
  float f = 1.f + sinf(m0.z * p.z + 10.f * perlin_turb(sc, f2i(m0.w), p, 7));
  return F3(0.5f * f, 0.5f * f, 0.5f * f);
}

// ONB (ONB.hpp:33-36,64)
struct Onb {
  f3 u, v, w;
};
RT_HD Onb onb_make(f3 n) {
  Onb b;
  b.w = normalize(n);
  f3 a = fabsf(b.w.x) > 0.9f ? F3(0.f, 1.f, 0.f) : F3(1.f, 0.f, 0.f);
  b.v = normalize(cross(b.w, a));
  b.u = cross(b.w, b.v);
  return b;
}
RT_HD f3 onb_transform(const Onb &b, f3 a) { return a.x * b.u + a.y * b.v + a.z * b.w; }

RT_HD void sincos_2pi(float u, float &sn, float &cs) {
#if defined(__CUDA_ARCH__)
  sincospif(2.0f * u, &sn, &cs);
#else
  sn = sinf(2.0f * RT_PI_F * u);
  cs = cosf(2.0f * RT_PI_F * u);
#endif
}

// cuda_vec3_random_unit_vector (Vec3Utility.cuh:65-70)
RT_HD f3 unit_vector_polar(float u1, float u2) {
  float z = -1.0f + 2.0f * u1;
  float r = fast_sqrt(fmaxf(0.f, 1.0f - z * z));
  float sn, cs;
  sincos_2pi(u2, sn, cs);
  return F3(r * cs, r * sn, z);
}
// random_cosine_direction (Vec3Utility.hpp:94-104)
RT_HD f3 cosine_direction(float r1, float r2) {
  float sn, cs;
  sincos_2pi(r1, sn, cs);
  float s = fast_sqrt(r2);
  return F3(cs * s, sn * s, fast_sqrt(1.f - r2));
}

// Lights: HittablePDF over the light list (PDF.hpp:82-113, HittableList.cpp:44-63).
//   quad   [0] corner, shape   [1] u, area   [2] v, -   [3] normal, D   [4] A, -   [5] B, -
//   sphere [0] center, shape   [1] radius
RT_HD f3 light_random(const float4 *l, f3 origin, float r1, float r2) {
  float4 l0 = ldg4(l), l1 = ldg4(l + 1);
  if (f2i(l0.w) == 1) { // quad: Plane::random (Plane.cpp:128-133)
    float4 l2 = ldg4(l + 2);
    f3 p = F3(l0) + r1 * F3(l1) + r2 * F3(l2);
    return p - origin;
  }
  // Sphere::random (Sphere.cpp:161-179)
  f3 direction = F3(l0) - origin;
  float distance_squared = dot(direction, direction);
  Onb uvw = onb_make(direction);
  float radius = l1.x;
  float z = 1.f + r2 * (fast_sqrt(fmaxf(0.f, 1.f - radius * radius * fast_rcp(distance_squared))) - 1.f);
  float sn, cs;
  sincos_2pi(r1, sn, cs);
  float s = fast_sqrt(fmaxf(0.f, 1.f - z * z));
  return onb_transform(uvw, F3(cs * s, sn * s, z));
}

RT_HD float light_pdf_value(const float4 *l, f3 origin, f3 dir) {
  float4 l0 = ldg4(l), l1 = ldg4(l + 1);
  Ray r;
  r.o = origin;
  r.d = dir;
  r.time = 0.f;
  float t;
  if (f2i(l0.w) == 1) { // Plane::pdf_value (Plane.cpp:115-126)
    float4 l2 = ldg4(l + 2), l3 = ldg4(l + 3), l4 = ldg4(l + 4), l5 = ldg4(l + 5);
    float4 q0 = make_float4(l3.x, l3.y, l3.z, l3.w);
    float4 q1 = make_float4(l4.x, l4.y, l4.z, l0.x);
    float4 q2 = make_float4(l5.x, l5.y, l5.z, l0.y);
    (void)l2;
    if (!quad_hit<false>(q0, q1, q2, l0.z, r, RT_T_MIN, RT_INF_F, t))
      return 0.f;
    float dd = dot(dir, dir);
    float distance_squared = t * t * dd;
    float cosine = fabsf(dot(dir, F3(l3)) * fast_rsqrt(dd));
    return distance_squared * fast_rcp(cosine * l1.w);
  }
  // Sphere::pdf_value (Sphere.cpp:145-159)
  float4 s0 = make_float4(l0.x, l0.y, l0.z, l1.x);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!sphere_hit<false>(s0, s1, r, RT_T_MIN, RT_INF_F, t))
    return 0.f;
  f3 oc = F3(l0) - origin;
  float cos_theta_max = fast_sqrt(fmaxf(0.f, 1.f - l1.x * l1.x * fast_rcp(dot(oc, oc))));
  float solid_angle = 2.f * RT_PI_F * (1.f - cos_theta_max);
  return fast_rcp(solid_angle);
}

// ---------------------------------------------------------------------------------------------------
// One integrator step (Camera::ray_color unrolled into a loop, Camera.cpp:232-309; pdf guard of
// CameraKernels.cu:192).  Returns true when the path continues with `next` (throughput already
// multiplied in); otherwise `radiance_out` is the path's contribution (throughput * emitted or
// throughput * background) and the path ends.
// ---------------------------------------------------------------------------------------------------
struct ShadeResult {
  Ray next;
  f3 throughput;
  f3 radiance;
  int next_skip_prim;
};

RT_HD bool shade_segment(const DScene &sc, const Ray &ray, Hit hit, f3 throughput, const RayKey &key,
                         bool last_bounce, ShadeResult &out) {
  out.radiance = F3(0.f, 0.f, 0.f);
  out.next_skip_prim = -1;
  if (hit.prim < 0) {
    out.radiance = throughput * F3(sc.bg[0], sc.bg[1], sc.bg[2]);
    return false;
  }
  const float4 *rec = sc.prims + (size_t)hit.prim * RT_PRIM_F4;
  float4 r0 = ldg4_now(rec), r1 = ldg4_now(rec + 1), r2 = ldg4_now(rec + 2), r3 = ldg4_now(rec + 3);
  uint32_t typemat = (uint32_t)f2i(r3.y);
  int type = typemat >> 28;
  const float4 *m = sc.mats + (size_t)(typemat & 0x0fffffffu) * RT_MAT_F4;
  // a sphere record carries its material header and, for metal / solid textures, the colour
  // (rt_flatten.h): the hit is shaded after one dependent fetch instead of three
  float4 m0 = type == RT_PT_SPHERE ? r2 : ldg4(m);
  int mtype = f2i(m0.x);
  const bool inline_color = type == RT_PT_SPHERE && (mtype == 1 || f2i(m0.y) == RT_DTEX_SOLID);
  const f3 color_a = inline_color ? F3(m0.w, r1.w, r3.x) : F3(0.f, 0.f, 0.f);

  // hit point, normal, front face (HitRecord::set_face_normal, HitRecord.hpp:36-39)
  f3 p = ray.o + hit.t * ray.d;
  f3 normal;
  bool front;
  float tex_u = 0.f, tex_v = 0.f; // surface coordinates, only computed for image textures
  const bool wants_uv = f2i(m0.y) == RT_DTEX_IMAGE;
  if (type == RT_PT_SPHERE) {
    f3 center = F3(r0) + ray.time * F3(r1);
    f3 outward = normalize(p - center); // (p - center) / radius (Sphere.cpp:123), renormalised
    p = center + r0.w * outward; // keep the point on the surface (FP32 drift on large spheres)
    front = dot(ray.d, outward) < 0.f;
    normal = front ? outward : -outward;
    out.next_skip_prim = hit.prim; // the next ray starts on this sphere: only its far root counts (leaf_test)
    if (wants_uv)
      sphere_uv(outward, tex_u, tex_v);
  } else if (type == RT_PT_QUAD) {
    f3 outward = F3(r0);
    front = dot(ray.d, outward) < 0.f;
    normal = front ? outward : -outward;
    out.next_skip_prim = hit.prim; // a ray leaving a flat primitive cannot hit it again
    if (wants_uv) {
      f3 hp = p - F3(r1.w, r2.w, r3.x);
      tex_u = dot(F3(r1), hp);
      tex_v = dot(F3(r2), hp);
    }
  } else {
    normal = F3(1.f, 0.f, 0.f); // ConstantMedium.cpp:86-88
    front = true;
  }

  if (mtype == 3) { // diffuse light: emits on the front face, never scatters (DiffuseLightMaterial.cpp:12-19)
    if (front)
      out.radiance = throughput * (inline_color ? color_a : material_texture(sc, m, m0, p, tex_u, tex_v));
    return false;
  }
  if (last_bounce) // the scattered ray would be traced with depth 0 and contribute nothing
    return false;

  Uniform4 u = philox_uniform4(key.seed, key.pixel, key.sample, key.bounce, RT_STREAM_SHADE, 0);
  out.next.o = p;
  out.next.time = ray.time;

  if (mtype == 1) { // metal (MetalMaterial.cpp:46-61)
    f3 reflected = ray.d - (2.f * dot(ray.d, normal)) * normal;
    out.next.d = normalize(reflected) + m0.z * unit_vector_polar(u.y, u.z);
    out.throughput = throughput * (inline_color ? color_a : F3(ldg4(m + 1)));
    return true;
  }
  if (mtype == 2) { // dielectric (DielectricMaterial.cpp:62-84, Vec3Utility.hpp:76-89)
    float ri = front ? fast_rcp(m0.z) : m0.z;
    f3 unit_direction = normalize(ray.d);
    float cos_theta = fminf(dot(-unit_direction, normal), 1.0f);
    float sin_theta = fast_sqrt(fmaxf(0.f, 1.0f - cos_theta * cos_theta));
    bool cannot_refract = ri * sin_theta > 1.0f;
    float r0s = (1.f - ri) * fast_rcp(1.f + ri);
    r0s = r0s * r0s;
    float c1 = 1.f - cos_theta;
    float reflectance = r0s + (1.f - r0s) * (c1 * c1 * c1 * c1 * c1);
    if (cannot_refract || reflectance > u.x) {
      out.next.d = unit_direction - (2.f * dot(unit_direction, normal)) * normal;
    } else {
      f3 perp = ri * (unit_direction + cos_theta * normal);
      f3 parallel = (-fast_sqrt(fabsf(1.0f - dot(perp, perp)))) * normal;
      out.next.d = perp + parallel;
    }
    out.throughput = throughput;
    return true;
  }

  // lambertian (LambertianMaterial.cpp:15-59) / isotropic (IsotropicMaterial.cpp:12-31)
  bool lambert = mtype == 0;
  f3 attenuation = inline_color ? color_a : material_texture(sc, m, m0, p, tex_u, tex_v);
  Onb uvw;
  if (lambert)
    uvw = onb_make(normal);
  f3 dir;
  if (u.x < 0.5f && sc.n_lights > 0) { // MixturePDF::generate (PDF.hpp:135-139)
    int pick = (int)(u.w * (float)sc.n_lights);
    pick = pick > sc.n_lights - 1 ? sc.n_lights - 1 : pick;
    dir = light_random(sc.lights + (size_t)pick * RT_LIGHT_F4, p, u.y, u.z);
  } else if (lambert) {
    dir = onb_transform(uvw, cosine_direction(u.y, u.z));
  } else {
    dir = unit_vector_polar(u.y, u.z);
  }
  f3 unit_dir = normalize(dir);
  float mat_pdf = lambert ? fmaxf(0.f, dot(unit_dir, uvw.w) * (1.0f / RT_PI_F)) : 1.0f / (4.0f * RT_PI_F);
  float first_pdf = mat_pdf;
  if (sc.n_lights > 0) { // HittableList::pdf_value (HittableList.cpp:44-55)
    float weight = 1.0f / (float)sc.n_lights;
    first_pdf = 0.f;
    for (int i = 0; i < sc.n_lights; i++)
      first_pdf += weight * light_pdf_value(sc.lights + (size_t)i * RT_LIGHT_F4, p, dir);
  }
  float pdf_value = 0.5f * first_pdf + 0.5f * mat_pdf;
  float scattering_pdf;
  if (lambert) {
    float cos_theta = dot(normal, unit_dir);
    scattering_pdf = cos_theta < 0.f ? 0.f : cos_theta * (1.0f / RT_PI_F);
  } else {
    scattering_pdf = 1.0f / (4.0f * RT_PI_F);
  }
  if (!(pdf_value > 1e-8f) || !(scattering_pdf > 0.f)) // zero weight: nothing further can contribute
    return false;
  out.next.d = dir;
  out.throughput = throughput * ((scattering_pdf * fast_rcp(pdf_value)) * attenuation);
  return true;
}

// ---------------------------------------------------------------------------------------------------
// Film
// ---------------------------------------------------------------------------------------------------
// Scanline-tile ownership: tile k (tile_rows consecutive scanlines) belongs to rank k % n_ranks; a
// rank stores its tiles compactly in tile order.
struct DFilmMap {
  int width, height, rank, n_ranks, tile_rows;
};
RT_HD int owned_row_to_global(const DFilmMap &m, int local_row) {
  int tile_local = local_row / m.tile_rows;
  return (tile_local * m.n_ranks + m.rank) * m.tile_rows + local_row % m.tile_rows;
}
RT_HD int owned_rows(int height, int rank, int n_ranks, int tile_rows) {
  int rows = 0;
  int n_tiles = (height + tile_rows - 1) / tile_rows;
  for (int t = rank; t < n_tiles; t += n_ranks) {
    int r0 = t * tile_rows;
    int r1 = r0 + tile_rows < height ? r0 + tile_rows : height;
    rows += r1 - r0;
  }
  return rows;
}

// Path numbering of a pass: path = sample_in_pass * n_owned + k, and k enumerates the rank's owned pixels either
// scanline by scanline or - when the film's shape allows it - by 8 x 4 pixel blocks, so that 32 consecutive paths
// (a warp of camera rays, and the hit points its scattered rays start from) form a compact bundle.  The film
// itself is always row-major over the owned scanlines.
struct PathMap {
  DFilmMap map;
  int n_owned;
  int tiled, blocks_x; // blocks_x = blocks per band of 4 owned scanlines
  FastDiv div_owned, div_width, div_tile_rows, div_blocks_x;
};
inline PathMap pathmap_make(const DFilmMap &map, long long n_owned, bool allow_blocks) {
  PathMap m;
  m.map = map;
  m.n_owned = (int)n_owned;
  const long long owned_scanlines = map.width > 0 ? n_owned / map.width : 0;
  m.tiled = allow_blocks && map.width > 0 && map.width % 8 == 0 && owned_scanlines % 4 == 0 && map.tile_rows % 4 == 0;
  m.blocks_x = map.width / 8 > 1 ? map.width / 8 : 1;
  m.div_owned = fastdiv_make((uint32_t)(n_owned > 1 ? n_owned : 1));
  m.div_width = fastdiv_make((uint32_t)(map.width > 1 ? map.width : 1));
  m.div_tile_rows = fastdiv_make((uint32_t)(map.tile_rows > 1 ? map.tile_rows : 1));
  m.div_blocks_x = fastdiv_make((uint32_t)m.blocks_x);
  return m;
}
// path -> sample index within the pass, film index of the pixel (row-major over the owned scanlines), global
// scanline and column
RT_HD void path_to_pixel(const PathMap &m, uint32_t path, uint32_t &sample_local, uint32_t &owned_pixel, int &row, int &col) {
  sample_local = fastdiv(m.div_owned, path);
  uint32_t k = path - sample_local * (uint32_t)m.n_owned;
  uint32_t local_row;
  if (m.tiled) {
    uint32_t block = k >> 5, lane = k & 31u;
    uint32_t band = fastdiv(m.div_blocks_x, block);
    uint32_t bx = block - band * (uint32_t)m.blocks_x;
    local_row = band * 4u + (lane >> 3);
    col = (int)(bx * 8u + (lane & 7u));
    k = local_row * (uint32_t)m.map.width + (uint32_t)col;
  } else {
    local_row = fastdiv(m.div_width, k);
    col = (int)(k - local_row * (uint32_t)m.map.width);
  }
  // owned_row_to_global: tile t of this rank is global tile t * n_ranks + rank
  uint32_t tile_local = fastdiv(m.div_tile_rows, local_row);
  uint32_t in_tile = local_row - tile_local * (uint32_t)m.map.tile_rows;
  row = (int)((tile_local * (uint32_t)m.map.n_ranks + (uint32_t)m.map.rank) * (uint32_t)m.map.tile_rows + in_tile);
  owned_pixel = k;
}

// to_byte (utils/ColorUtility.hpp:11-26) in the reference's FP64.
RT_HD unsigned char to_byte_f64(double v) {
  double x = v > 0 ? sqrt(v) : 0;
  if (x < 0.000)
    x = 0.000;
  if (x > 0.999)
    x = 0.999;
  return (unsigned char)(256 * x);
}
