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
