// rt_exact.cu — kernel of the FP64 parity path (rt_trace_rays, RT_TRACE_EXACT_F64).
// Compiled with --fmad=false: every multiply and add rounds separately, as in the reference's
// scalar FP64 code (see rt_exact.h).
#include "rt_internal.h"

__global__ void __launch_bounds__(128)
    k_trace_exact(const __grid_constant__ ExactScene sc, const rt_ray *__restrict__ rays, long long n, uint64_t seed,
                  rt_hit *__restrict__ hits) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const rt_ray &in = rays[q];
    RayD r;
    for (int k = 0; k < 3; k++) {
      r.o[k] = in.origin[k];
      r.d[k] = in.direction[k];
    }
    r.time = in.time;
    RayKey key;
    key.seed = seed;
    key.pixel = in.rng_pixel;
    key.sample = in.rng_sample;
    key.bounce = in.rng_bounce;
    HitD h;
    traverse_exact(sc, r, in.t_min, in.t_max, h, key);
    rt_hit out;
    out.t = h.prim >= 0 ? h.t : (double)RT_INF_F;
    out.prim = h.id;
    out.object = h.object;
    out.front_face = h.prim >= 0 ? h.front : 0;
    out.pad_ = 0;
    hits[q] = out;
  }
}

void launch_trace_exact(cudaStream_t s, const ExactScene &sc, const rt_ray *d_rays, int64_t n, uint64_t seed,
                        rt_hit *d_hits) {
  if (n <= 0)
    return;
  long long blocks = (n + 127) / 128;
  if (blocks > 148 * 16)
    blocks = 148 * 16;
  k_trace_exact<<<(int)blocks, 128, 0, s>>>(sc, d_rays, n, seed, d_hits);
}

// ---------------------------------------------------------------------------------------------------
// Parity audit (rt_context_set_audit): the FP64 traversal of every queued ray, and its comparison with
// what the FP32 extend kernel answered for the same queue slot.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ RayD queue_ray(const float4 a, const float4 b) {
  RayD r;
  r.o[0] = a.x, r.o[1] = a.y, r.o[2] = a.z; // float -> double is exact: the very ray the FP32 kernel traces
  r.d[0] = b.x, r.d[1] = b.y, r.d[2] = b.z;
  r.time = a.w;
  return r;
}

// Before k_extend of `bounce`: hit[q].y still holds the primitive the ray must skip.
__global__ void __launch_bounds__(128)
    k_audit_trace(const __grid_constant__ ExactScene sc, const __grid_constant__ PassParams pp,
                  const float4 *__restrict__ ray_a, const float4 *__restrict__ ray_b, const float2 *__restrict__ hit,
                  const unsigned int *__restrict__ counts, int bounce, int has_media, int2 *__restrict__ audit_prim,
                  double *__restrict__ audit_t) {
  const unsigned int n = counts[bounce];
  const unsigned int stride = gridDim.x * blockDim.x;
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    float4 a = ray_a[q], b = ray_b[q];
    RayD r = queue_ray(a, b);
    RayKey key;
    key.seed = pp.seed;
    key.pixel = key.sample = 0;
    key.bounce = (uint32_t)bounce;
    if (has_media) {
      int k;
      path_to_key(pp, __float_as_int(b.w), bounce, key, k);
    }
    const int skip = bounce == 0 ? -1 : __float_as_int(hit[q].y);
    HitD h;
    traverse_exact(sc, r, 0.001, (double)RT_INF_F, h, key, skip); // Interval(0.001, inf), Camera.cpp:242
    audit_prim[q] = make_int2(h.prim, skip);
    audit_t[q] = h.t;
  }
}

// After k_extend of `bounce`: hit[q] = (t, primitive) of the FP32 traversal.
__global__ void __launch_bounds__(128)
    k_audit_compare(const float4 *__restrict__ ray_a, const float4 *__restrict__ ray_b, const float2 *__restrict__ hit,
                    const unsigned int *__restrict__ counts, int bounce, const int2 *__restrict__ audit_prim,
                    const double *__restrict__ audit_t, const int *__restrict__ leaf_id,
                    unsigned long long *__restrict__ tallies, rt_audit_sample *__restrict__ samples) {
  const unsigned int n = counts[bounce];
  const unsigned int stride = gridDim.x * blockDim.x;
  unsigned int mismatch = 0, flips = 0, t_off = 0;
  float max_rel = 0.f;
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    float2 h = hit[q];
    const int2 ap = audit_prim[q];
    int fast = __float_as_int(h.y), exact = ap.x;
    if (fast != exact) {
      mismatch++;
      flips += (fast < 0) != (exact < 0);
      unsigned long long slot = atomicAdd(&tallies[RT_AUDIT_SAMPLES], 1ull);
      if (slot < RT_AUDIT_MAX_SAMPLES) {
        float4 a = ray_a[q], b = ray_b[q];
        rt_audit_sample s;
        s.origin[0] = a.x, s.origin[1] = a.y, s.origin[2] = a.z, s.time = a.w;
        s.direction[0] = b.x, s.direction[1] = b.y, s.direction[2] = b.z;
        s.bounce = bounce;
        s.fast_prim = fast >= 0 ? leaf_id[fast] : -1;
        s.exact_prim = exact >= 0 ? leaf_id[exact] : -1;
        s.fast_t = h.x;
        s.skip_prim = ap.y >= 0 ? leaf_id[ap.y] : -1;
        s.exact_t = audit_t[q];
        samples[slot] = s;
      }
    } else if (fast >= 0) {
      double te = audit_t[q];
      double denom = fabs(te) > 1e-3 ? fabs(te) : 1e-3;
      float rel = (float)(fabs((double)h.x - te) / denom);
      max_rel = fmaxf(max_rel, rel);
      t_off += rel > 1e-4f;
    }
  }
  // block totals -> one atomic per tally
  __shared__ unsigned int s_sum[3];
  __shared__ unsigned int s_max;
  if (threadIdx.x == 0) {
    s_sum[0] = s_sum[1] = s_sum[2] = 0u;
    s_max = 0u;
  }
  __syncthreads();
  if (mismatch)
    atomicAdd(&s_sum[0], mismatch);
  if (flips)
    atomicAdd(&s_sum[1], flips);
  if (t_off)
    atomicAdd(&s_sum[2], t_off);
  atomicMax(&s_max, __float_as_uint(max_rel)); // non-negative floats order like their bit patterns
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_sum[0]) {
      atomicAdd(&tallies[RT_AUDIT_MISMATCH], (unsigned long long)s_sum[0]);
      if (bounce == 0)
        atomicAdd(&tallies[RT_AUDIT_PRIMARY_MISMATCH], (unsigned long long)s_sum[0]);
    }
    if (s_sum[1])
      atomicAdd(&tallies[RT_AUDIT_HIT_MISS], (unsigned long long)s_sum[1]);
    if (s_sum[2])
      atomicAdd(&tallies[RT_AUDIT_T_ABOVE_1E4], (unsigned long long)s_sum[2]);
    atomicMax(&tallies[RT_AUDIT_MAX_REL_T], (unsigned long long)s_max);
    if (blockIdx.x == 0) {
      atomicAdd(&tallies[RT_AUDIT_SEGMENTS], (unsigned long long)n);
      if (bounce == 0)
        atomicAdd(&tallies[RT_AUDIT_PRIMARY], (unsigned long long)n);
    }
  }
}

void launch_audit_trace(const rt_context *ctx, const ExactScene &sc, const PassParams &pp, WaveBuffers &w, int bounce,
                        int has_media) {
  int b = bounce & 1;
  long long blocks = ((long long)pp.n_paths + 127) / 128;
  if (blocks > ctx->sm_count * 16)
    blocks = ctx->sm_count * 16;
  k_audit_trace<<<(int)blocks, 128, 0, ctx->stream>>>(sc, pp, w.ray_a[b], w.ray_b[b], w.hit[b], w.counts, bounce, has_media,
                                                      w.audit_prim, w.audit_t);
}

void launch_audit_compare(const rt_context *ctx, const PassParams &pp, WaveBuffers &w, int bounce, const int *leaf_id) {
  int b = bounce & 1;
  long long blocks = ((long long)pp.n_paths + 127) / 128;
  if (blocks > ctx->sm_count * 16)
    blocks = ctx->sm_count * 16;
  k_audit_compare<<<(int)blocks, 128, 0, ctx->stream>>>(w.ray_a[b], w.ray_b[b], w.hit[b], w.counts, bounce, w.audit_prim,
                                                        w.audit_t, leaf_id, w.audit_stats, w.audit_samples);
}
