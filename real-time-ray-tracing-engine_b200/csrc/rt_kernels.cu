// rt_kernels.cu — the CUDA kernels of the wavefront path tracer and of the LBVH build, for sm_100a.
//
// Pipeline per pass (replaces the reference's one-thread-per-pixel recursive megakernels,
// core/camera/CameraKernels.cu:206-278):
//   k_generate    camera rays for all paths of the pass             -> ray queue 0
//   k_extend      BVH4 traversal + sphere / quad / medium tests     -> hit per queue slot
//   k_shade       emit / scatter / light sampling, Philox draws     -> compacted ray queue of the next bounce
//   k_tail        after the first bounces: the thin rest of the path population, traced and shaded to
//                 completion in one persistent launch
//   k_accumulate  multi-sample passes: per-pixel sum of the pass's samples, in order -> film (in one-sample
//                 passes the kernel that ends a path adds its radiance to the film itself)
// Queue lengths stay on the device (counts[bounce]); every kernel is a persistent grid sized in
// multiples of the SM count that pulls work from the queue, so the host never synchronises between
// bounces.  Compaction uses one ballot per warp and one atomic per block.  Paths are numbered by 8 x 4 pixel
// blocks, so a warp of camera rays is a compact bundle.
#include "rt_internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <string>
#include <utility>

#define RT_BLOCK 128
#define RT_STACK_SMEM 8

// Programmatic dependent launch (sm_90+), RT_PDL=1: the extend / shade / tail kernels of a pass are launched with the
// programmatic-serialization attribute, so the next kernel's launch is processed - and its blocks take the SM slots the
// previous kernel's blocks free - while the previous kernel still runs its last rays; every kernel waits for its
// predecessor's completion (and memory) before it touches anything, and lets its successor start launching at once.
// Measured on B200 inside the pass graph (profiles/r02_experiments.md): images identical, frames 0.4-1.2 % SLOWER (the
// graph already leaves no launch gap to hide, and waiting blocks hold SM slots), so it is off unless asked for.
__device__ __forceinline__ void rt_pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
template <class... KArgs, class... Args>
static void rt_launch(void (*kernel)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t stream, bool dependent,
                      Args &&...args) {
  static const bool pdl = getenv("RT_PDL") && atoi(getenv("RT_PDL")) != 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && dependent) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#ifndef RT_TAIL_SHARE_MIN_PRIMS
#define RT_TAIL_SHARE_MIN_PRIMS 32768 // end-game sharing in k_tail from this many primitives on (0: always, -1: never)
#endif
// Scenes whose longest traversals can hold a launch up get k_tail's end-game sharing - and with it leaf_test's
// order-independent tie rule in every kernel that traces them.  RT_TAIL_SHARE=0 / 1 forces both off / on; 2 = the tie
// rule without the sharing (what a shared traversal must reproduce bit for bit: tools/env_check.py).
static int rt_share_mode(const DScene &sc) {
  static const int share_env = getenv("RT_TAIL_SHARE") ? atoi(getenv("RT_TAIL_SHARE")) : -1;
  if (share_env >= 0)
    return share_env;
  return RT_TAIL_SHARE_MIN_PRIMS >= 0 && sc.n_prims >= RT_TAIL_SHARE_MIN_PRIMS ? 1 : 0;
}
static bool rt_scene_tie_rule(const DScene &sc) { return rt_share_mode(sc) != 0; }
static bool rt_scene_shares_traversals(const DScene &sc) { return rt_share_mode(sc) == 1; }

LaunchShape rt_persistent_shape(const rt_context *ctx, int threads, int blocks_per_sm) {
  LaunchShape s;
  s.threads = threads;
  s.blocks = ctx->sm_count * blocks_per_sm;
  return s;
}

// ---------------------------------------------------------------------------------------------------
// Short traversal stack: the first RT_STACK_SMEM entries of every thread live in shared memory
// (column-interleaved: no bank conflicts), deeper entries spill to local memory.
// ---------------------------------------------------------------------------------------------------
// One (reference, entry distance) pair per 8-byte word.  Every thread keeps the shared-window address of
// its own column in a register (laundered through an empty asm so that it is not recomputed from the CTA's
// window base at every use): a push or pop is one multiply-add and one 64-bit shared-memory access.
__shared__ int2 rt_stack_smem[RT_STACK_SMEM * RT_BLOCK];
struct SmemStack {
  static constexpr bool has_fast_push = true;
  static constexpr int fast_depth = RT_STACK_SMEM;
  unsigned base;
  StackEntry spill[RT_STACK - RT_STACK_SMEM];
  __device__ __forceinline__ void init() {
    base = (unsigned)__cvta_generic_to_shared(&rt_stack_smem[threadIdx.x]);
    asm volatile("" : "+r"(base));
  }
  __device__ __forceinline__ void set_fast_if(int i, int ref, float t, bool p) { // i < RT_STACK_SMEM
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t}"
                 :
                 : "r"(base + (unsigned)i * (RT_BLOCK * 8u)), "r"(ref), "r"(__float_as_int(t)), "r"((int)p)
                 : "memory");
  }
  __device__ __forceinline__ void set(int i, int ref, float t) {
    if (i < RT_STACK_SMEM) {
      set_fast_if(i, ref, t, true);
    } else {
      spill[i - RT_STACK_SMEM].ref = ref;
      spill[i - RT_STACK_SMEM].t = t;
    }
  }
  __device__ __forceinline__ void get(int i, int &ref, float &t) const {
    if (i < RT_STACK_SMEM) {
      int tb;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                   : "=r"(ref), "=r"(tb)
                   : "r"(base + (unsigned)i * (RT_BLOCK * 8u))
                   : "memory");
      t = __int_as_float(tb);
    } else {
      ref = spill[i - RT_STACK_SMEM].ref;
      t = spill[i - RT_STACK_SMEM].t;
    }
  }
  // entry i (< RT_STACK_SMEM) of the column `lane_offset` threads away (another lane of the same warp)
  __device__ __forceinline__ void get_from(int lane_offset, int i, int &ref, float &t) const {
    int tb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                 : "=r"(ref), "=r"(tb)
                 : "r"(base + (unsigned)(lane_offset * 8) + (unsigned)i * (RT_BLOCK * 8u))
                 : "memory");
    t = __int_as_float(tb);
  }
};
#define RT_DECLARE_STACK(stack)                                                                              \
  SmemStack stack;                                                                                           \
  stack.init()

// A path ends exactly once.  In a one-sample pass it is the only path of its pixel, so its contribution goes
// straight into the film (bit-identical to k_accumulate's film + radiance); multi-sample passes park it per
// path and k_accumulate sums each pixel's samples in order.
__device__ __forceinline__ void path_ends(const PassParams &pp, float4 *__restrict__ radiance, int path, int owned_pixel,
                                          f3 rad) {
  if (pp.film_direct) {
#if RT_FILM_RED
    // no other path of this pass touches the pixel: three reductions give the same sums as load-add-store
    // without waiting for the load (subnormal sums would be flushed; radiance never gets there)
    float *f = reinterpret_cast<float *>(pp.film_direct + owned_pixel);
    atomicAdd(f + 0, rad.x);
    atomicAdd(f + 1, rad.y);
    atomicAdd(f + 2, rad.z);
#else
    float4 f = pp.film_direct[owned_pixel];
    f.x += rad.x;
    f.y += rad.y;
    f.z += rad.z;
    pp.film_direct[owned_pixel] = f;
#endif
  } else {
    radiance[path] = make_float4(rad.x, rad.y, rad.z, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------
// generate
// ---------------------------------------------------------------------------------------------------
// The camera ray of path p (Camera::get_ray, Camera.cpp:186-205): a pure function of the pass parameters and the
// path's Philox key, so any kernel can derive it instead of reading it from memory.
__device__ __forceinline__ Ray path_camera_ray(const PassParams &pp, int p) {
  RayKey key;
  int k, row, col;
  path_to_key(pp, p, 0, key, k, row, col);
  int s_j = (int)fastdiv(pp.div_sqrt_spp, key.sample), s_i = (int)key.sample - s_j * pp.sqrt_spp;
  Uniform4 u0 = philox_uniform4(pp.seed, key.pixel, key.sample, 0, RT_STREAM_CAMERA, 0);
  Uniform4 u1 = philox_uniform4(pp.seed, key.pixel, key.sample, 0, RT_STREAM_CAMERA, 1);
  return camera_ray(pp.cam, col, row, s_i, s_j, pp.recip_sqrt_spp, u0, u1);
}

// Only launched when the first bounce does not generate its own rays (tail-only schedules, the parity audit).
__global__ void __launch_bounds__(RT_BLOCK)
    k_generate(const __grid_constant__ PassParams pp, float4 *__restrict__ ray_a, float4 *__restrict__ ray_b,
               unsigned int *__restrict__ counts) {
  rt_pdl_enter();
  int stride = gridDim.x * blockDim.x;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < pp.n_paths; p += stride) {
    Ray r = path_camera_ray(pp, pp.path_base + p);
    ray_a[p] = make_float4(r.o.x, r.o.y, r.o.z, r.time);
    ray_b[p] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(pp.path_base + p));
    // no per-path initialisation is written: at bounce 0 the throughput is 1 and no primitive is skipped
    // (the kernels know), and every path writes its radiance exactly once when it ends
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    counts[0] = (unsigned int)pp.n_paths;
}

// ---------------------------------------------------------------------------------------------------
// extend
// ---------------------------------------------------------------------------------------------------
// Persistent warps with dynamic ray fetch and a while-while traversal loop (Aila & Laine, "Understanding
// the efficiency of ray traversal on GPUs"): every warp keeps pulling rays from the bounce's queue through a
// device-side cursor; lanes first descend inner nodes only, then test leaves only, so a warp never
// executes box code and primitive code in the same iteration; when fewer than RT_REFILL lanes still hold a
// ray, the idle lanes fetch new ones (one atomicAdd per warp).
#ifndef RT_REFILL
#define RT_REFILL 16 // measured: 16 beats 22 and 28 by 1-2 %
#endif
#ifndef RT_TAIL_BLOCKS
#define RT_TAIL_BLOCKS 5 // resident blocks per SM of k_tail (register budget = 65536 / (128 * RT_TAIL_BLOCKS))
#endif
#ifndef RT_FILM_RED
#define RT_FILM_RED 1
#endif
#ifndef RT_TAIL_REFILL
#define RT_TAIL_REFILL 8 // k_tail leaves the traversal loop to shade / refill below this many busy lanes
#endif
#define RT_DONE 0x7fffffff

// STATS: the instrumented instantiation (rt_context_set_stats) counts node visits and primitive tests; the
// product instantiation carries no counting code.
// STATS only: the longest single traversal ([4]) and how many rays needed more than 512 node visits ([5])
template <bool STATS> __device__ __forceinline__ void ray_stats(unsigned long long *stats, unsigned int ray_nodes) {
  if (STATS) {
    atomicMax(&stats[4], (unsigned long long)ray_nodes);
    if (ray_nodes > 512u)
      atomicAdd(&stats[5], 1ull);
  }
}
template <bool STATS>
__device__ __forceinline__ void traversal_stats(unsigned long long *stats, unsigned int n_nodes, unsigned int n_tests) {
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, o);
      n_tests += __shfl_xor_sync(0xffffffffu, n_tests, o);
    }
    if ((threadIdx.x & 31u) == 0 && (n_nodes | n_tests)) {
      atomicAdd(&stats[1], (unsigned long long)n_nodes);
      atomicAdd(&stats[2], (unsigned long long)n_tests);
    }
  }
}

// GEN (bounce 0 only): the kernel derives the camera ray of path q itself instead of reading queue 0, which is
// then never written (k_generate is not launched, k_shade<GEN> re-derives the ray as well): one launch and
// 2 x 32 B per path of queue traffic less.  The ray a lane is tracing is parked in shared memory for the
// primitive tests (origin, direction and time are not needed by node visits and would not fit in registers).
#ifndef RT_RAY_SMEM
#define RT_RAY_SMEM 0 // 1: bounces > 0 also park the fetched ray in shared memory instead of re-reading the queue
#endif
__shared__ float4 rt_ray_smem_a[RT_BLOCK], rt_ray_smem_b[RT_BLOCK];

template <bool STATS, bool GEN, bool TIES>
__global__ void __launch_bounds__(RT_BLOCK, 8)
    k_extend(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp,
             const float4 *__restrict__ ray_a, const float4 *__restrict__ ray_b, float2 *__restrict__ hit,
             unsigned int *__restrict__ counts, unsigned int *__restrict__ cursor, int bounce, int has_media,
             unsigned long long *stats) {
  rt_pdl_enter();
  RT_DECLARE_STACK(stack);
  constexpr bool PARK = GEN || RT_RAY_SMEM;
  const unsigned int n = GEN ? (unsigned int)pp.n_paths : counts[bounce];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&stats[0], (unsigned long long)n);
    if (GEN)
      counts[0] = n; // the length of the (virtual) queue 0, read by the shade launch that follows
  }
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;

  // Per-lane state kept in registers across node visits: the slab constants, the closest hit, the stack
  // pointer and the current reference.  Origin, direction, time and the skip primitive are only needed by
  // primitive tests (1.65 per segment vs 6.5 node visits), so they are re-read from the queue there.
  RayTrav rt = make_trav(F3(0.f, 0.f, 0.f), F3(1.f, 1.f, 1.f));
  Hit best;
  int sp = 0, ref = RT_DONE;
  unsigned int q = 0, n_nodes = 0, n_tests = 0, ray_nodes = 0;
  bool exhausted = false; // the queue has no more rays to hand out
  best.t = -1.0f;         // no ray held
  best.prim = -1;

  for (;;) {
    // ---- fetch: idle lanes take the next rays of the queue ----
    unsigned int idle = __ballot_sync(0xffffffffu, ref == RT_DONE);
    if (idle && !exhausted) {
      unsigned int base = 0;
      int leader = __ffs(idle) - 1;
      if ((int)lane == leader)
        base = atomicAdd(cursor, (unsigned int)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (ref == RT_DONE) {
        unsigned int mine = base + (unsigned int)__popc(idle & lt_mask);
        if (mine < n) {
          q = mine;
          float4 a, b;
          if (GEN) {
            Ray r = path_camera_ray(pp, pp.path_base + (int)q);
            a = make_float4(r.o.x, r.o.y, r.o.z, r.time);
            b = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(pp.path_base + (int)q));
          } else {
            a = ray_a[q];
            b = ray_b[q];
          }
          if (PARK) {
            rt_ray_smem_a[threadIdx.x] = a;
            rt_ray_smem_b[threadIdx.x] = b;
          }
          rt = make_trav(F3(a.x, a.y, a.z), F3(b.x, b.y, b.z));
          best.t = RT_INF_F;
          best.prim = -1;
          sp = 0;
          ref = 0; // root
          ray_nodes = 0;
        }
      }
      exhausted = base + (unsigned int)__popc(idle) >= n;
    }
    if (__all_sync(0xffffffffu, ref == RT_DONE))
      break;

    // ---- traverse until too few lanes are busy ----
    for (;;) {
      // inner nodes only
      while (ref >= 0 && ref != RT_DONE) {
        if (STATS) {
          n_nodes++;
          ray_nodes++;
        }
        if (!node_visit<SmemStack, true>(sc, ref, rt, RT_T_MIN, best.t, stack, sp, ref))
          if (!stack_pop(stack, sp, best.t, ref))
            ref = RT_DONE;
      }
      // leaves only
      if (ref < 0) {
        float4 a, b;
        if (PARK) {
          a = rt_ray_smem_a[threadIdx.x];
          b = rt_ray_smem_b[threadIdx.x];
        } else {
          a = ray_a[q];
          b = ray_b[q];
        }
        Ray r;
        r.o = F3(a.x, a.y, a.z);
        r.d = F3(b.x, b.y, b.z);
        r.time = a.w;
        int skip = bounce == 0 ? -1 : __float_as_int(hit[q].y);
        RayKey key;
        key.seed = pp.seed;
        key.pixel = key.sample = 0;
        key.bounce = (uint32_t)bounce;
        if (has_media) {
          int k;
          path_to_key(pp, __float_as_int(b.w), bounce, key, k);
        }
        do {
          if (STATS)
            n_tests++;
          leaf_test<TIES>(sc, ~ref, r, RT_T_MIN, best, skip, key);
          if (!stack_pop(stack, sp, best.t, ref))
            ref = RT_DONE;
        } while (ref < 0);
      }
      bool finished = ref == RT_DONE && best.t != -1.0f;
      if (finished) { // write the result once
        ray_stats<STATS>(stats, ray_nodes);
        hit[q] = make_float2(best.t, __int_as_float(best.prim));
        best.t = -1.0f; // marks "already written / no ray held"
      }
      unsigned int busy = __ballot_sync(0xffffffffu, ref != RT_DONE);
      if (busy == 0 || (!exhausted && __popc(busy) < RT_REFILL))
        break;
    }
  }
  traversal_stats<STATS>(stats, n_nodes, n_tests);
}

// ---------------------------------------------------------------------------------------------------
// extend, two rays per lane (RT_EXTEND_MUX=1) - an EXPERIMENT, measured slower, not the default
// ---------------------------------------------------------------------------------------------------
// The while-while loop above keeps a lane idle from the moment its ray reaches a leaf until every other lane of
// the warp has reached one too (ncu: 17 of 32 lanes in the node loop on scattered rays).  Here every lane owns TWO
// rays: the one in registers and a parked one whose traversal state (11 words) and short stack live in shared
// memory.  A lane whose ray leaves the node phase swaps to its parked ray if that one still has nodes to visit,
// so the node loop only runs out of work for a lane when BOTH its rays wait for the leaf phase; the leaf phase
// then serves both.  Same node visits, same primitive tests, same answers per ray (images bit-identical,
// tools/mux_check.py) - only their interleaving changes.
// Measured on B200 (profiles/r02_experiments.md, ncu capture r02_mux_ncu.md): the node loop does run at 24.6 of 32
// lanes instead of 17.2 and takes 30 % fewer iterations, exactly as intended - and the kernel is still 1.5-1.6 x
// SLOWER (C2 extend 0.49-0.54 ms against 0.33): the swap puts shared-memory round trips and a warp-wide vote on
// the critical path of every node iteration, issue rate falls from 2.6 to 1.3 instructions per clock (long-
// scoreboard stalls 2.6 -> 17), i.e. the issue-bound kernel becomes latency bound, and the swap / vote
// instructions give back a third of the saved iterations.  Kept selectable for the record; exit / refill
// thresholds and 6 blocks per SM without spills move it by < 10 %.
#ifndef RT_MUX_NODE_EXIT
#define RT_MUX_NODE_EXIT 12 // leave the node phase when fewer lanes than this still have a node to visit
#endif
#ifndef RT_MUX_REFILL
#define RT_MUX_REFILL 20 // refill when at least this many of the warp's 64 ray slots are empty
#endif
__shared__ __align__(16384) int2 rt_mux_stack_smem[2 * RT_STACK_SMEM * RT_BLOCK];
__shared__ float4 rt_mux_park[3 * RT_BLOCK];
struct MuxStack { // SmemStack with two regions; `which` selects the region of the ray in registers
  static constexpr bool has_fast_push = true;
  static constexpr int fast_depth = RT_STACK_SMEM;
  unsigned base0, base;
  int spill_off;
  StackEntry spill[2 * (RT_STACK - RT_STACK_SMEM)];
  __device__ __forceinline__ void init() {
    base0 = (unsigned)__cvta_generic_to_shared(&rt_mux_stack_smem[threadIdx.x]);
    asm volatile("" : "+r"(base0));
    select(0);
  }
  __device__ __forceinline__ void select(int which) {
    base = base0 + (unsigned)which * (RT_STACK_SMEM * RT_BLOCK * 8u);
    spill_off = which * (RT_STACK - RT_STACK_SMEM);
  }
  __device__ __forceinline__ void set_fast_if(int i, int ref, float t, bool p) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t}"
                 :
                 : "r"(base + (unsigned)i * (RT_BLOCK * 8u)), "r"(ref), "r"(__float_as_int(t)), "r"((int)p)
                 : "memory");
  }
  __device__ __forceinline__ void set(int i, int ref, float t) {
    if (i < RT_STACK_SMEM) {
      set_fast_if(i, ref, t, true);
    } else {
      spill[spill_off + i - RT_STACK_SMEM].ref = ref;
      spill[spill_off + i - RT_STACK_SMEM].t = t;
    }
  }
  __device__ __forceinline__ void get(int i, int &ref, float &t) const {
    if (i < RT_STACK_SMEM) {
      int tb;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                   : "=r"(ref), "=r"(tb)
                   : "r"(base + (unsigned)i * (RT_BLOCK * 8u))
                   : "memory");
      t = __int_as_float(tb);
    } else {
      ref = spill[spill_off + i - RT_STACK_SMEM].ref;
      t = spill[spill_off + i - RT_STACK_SMEM].t;
    }
  }
};

enum { RT_SLOT_EMPTY = 0, RT_SLOT_NODE = 1, RT_SLOT_LEAF = 2 };

// traversal state of one ray: what moves between the registers and the parked slot
struct MuxRay {
  f3 inv, oi;
  Hit best;
  int sp, ref;
  unsigned int q;
};
__device__ __forceinline__ RayTrav mux_trav(const MuxRay &r) { // the derived parts of RayTrav are recomputed, not parked
  return trav_from(r.inv, r.oi);
}
__device__ __forceinline__ int mux_state(int ref) { return ref == RT_DONE ? RT_SLOT_EMPTY : (ref >= 0 ? RT_SLOT_NODE : RT_SLOT_LEAF); }

#ifndef RT_MUX_BLOCKS
#define RT_MUX_BLOCKS 8
#endif
// registers <-> parked slot, one 16-byte row at a time so that only four temporaries are live
#define RT_MUX_SWAP()                                                                                        \
  do {                                                                                                       \
    if (cur.ref == RT_DONE && cur.best.t != -1.0f) { /* a ray that ended writes its answer before it leaves */ \
      hit[cur.q] = make_float2(cur.best.t, __int_as_float(cur.best.prim));                                   \
      cur.best.t = -1.0f;                                                                                    \
    }                                                                                                        \
    float4 row = rt_mux_park[threadIdx.x];                                                                   \
    rt_mux_park[threadIdx.x] = make_float4(cur.inv.x, cur.inv.y, cur.inv.z, cur.best.t);                     \
    cur.inv = F3(row.x, row.y, row.z);                                                                       \
    cur.best.t = row.w;                                                                                      \
    row = rt_mux_park[RT_BLOCK + threadIdx.x];                                                               \
    rt_mux_park[RT_BLOCK + threadIdx.x] = make_float4(cur.oi.x, cur.oi.y, cur.oi.z, __int_as_float(cur.best.prim)); \
    cur.oi = F3(row.x, row.y, row.z);                                                                        \
    cur.best.prim = __float_as_int(row.w);                                                                   \
    row = rt_mux_park[2 * RT_BLOCK + threadIdx.x];                                                           \
    rt_mux_park[2 * RT_BLOCK + threadIdx.x] =                                                                \
        make_float4(__int_as_float(cur.sp), __int_as_float(cur.ref), __uint_as_float(cur.q), 0.f);           \
    const int new_parked = mux_state(cur.ref);                                                               \
    cur.sp = __float_as_int(row.x);                                                                          \
    cur.ref = parked == RT_SLOT_EMPTY ? RT_DONE : __float_as_int(row.y);                                     \
    cur.q = __float_as_uint(row.z);                                                                          \
    if (parked == RT_SLOT_EMPTY)                                                                             \
      cur.best.t = -1.0f;                                                                                    \
    parked = new_parked;                                                                                     \
    stack.base ^= stack_toggle;                                                                              \
    stack.spill_off ^= (RT_STACK - RT_STACK_SMEM);                                                           \
    rt = mux_trav(cur);                                                                                      \
  } while (0)

template <bool STATS>
__global__ void __launch_bounds__(RT_BLOCK, RT_MUX_BLOCKS)
    k_extend_mux(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp, const float4 *__restrict__ ray_a,
                 const float4 *__restrict__ ray_b, float2 *__restrict__ hit, unsigned int *__restrict__ counts,
                 unsigned int *__restrict__ cursor, int bounce, int has_media, unsigned long long *stats) {
  MuxStack stack;
  stack.init();
  // the two stack regions are RT_STACK_SMEM * RT_BLOCK * 8 bytes apart: with the first one aligned to that size
  // (a power of two) the region is switched by flipping one address bit
  static_assert((RT_STACK_SMEM * RT_BLOCK * 8) == 8192, "stack region toggle assumes 8 KiB regions");
  const unsigned stack_toggle = 8192u; // region 0 starts 16 KiB-aligned (rt_mux_stack_smem), so bit 13 selects the region
  const unsigned int n = counts[bounce];
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&stats[0], (unsigned long long)n);
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;

  MuxRay cur; // the ray in registers
  cur.inv = cur.oi = F3(1.f, 1.f, 1.f);
  cur.best.t = -1.f;
  cur.best.prim = -1;
  cur.sp = 0;
  cur.ref = RT_DONE;
  cur.q = 0;
  RayTrav rt = mux_trav(cur);
  int parked = RT_SLOT_EMPTY; // state of the parked ray
  unsigned int n_nodes = 0, n_tests = 0;
  bool exhausted = false;

  for (;;) {
    // ---- refill: empty slots take the next rays of the queue (registers first, then the parked slots) ----
    if (!exhausted) {
      unsigned int empty_cur = __ballot_sync(0xffffffffu, cur.ref == RT_DONE);
      unsigned int empty_parked = __ballot_sync(0xffffffffu, parked == RT_SLOT_EMPTY);
      if (__popc(empty_cur) + __popc(empty_parked) >= RT_MUX_REFILL) {
#pragma unroll 1
        for (int round = 0; round < 2 && !exhausted; round++) {
          const unsigned int want = round == 0 ? empty_cur : empty_parked;
          if (!want)
            continue;
          const bool mine_wanted = round == 0 ? cur.ref == RT_DONE : parked == RT_SLOT_EMPTY;
          unsigned int base = 0;
          int leader = __ffs(want) - 1;
          if ((int)lane == leader)
            base = atomicAdd(cursor, (unsigned int)__popc(want));
          base = __shfl_sync(0xffffffffu, base, leader);
          exhausted = base + (unsigned int)__popc(want) >= n;
          unsigned int mine = base + (unsigned int)__popc(want & lt_mask);
          if (mine_wanted && mine < n) {
            float4 a = ray_a[mine], b = ray_b[mine];
            RayTrav t = make_trav(F3(a.x, a.y, a.z), F3(b.x, b.y, b.z));
            if (round == 0) {
              cur.inv = t.inv;
              cur.oi = t.oi;
              cur.best.t = RT_INF_F;
              cur.best.prim = -1;
              cur.sp = 0;
              cur.ref = 0; // root
              cur.q = mine;
              rt = t;
            } else {
              rt_mux_park[threadIdx.x] = make_float4(t.inv.x, t.inv.y, t.inv.z, RT_INF_F);
              rt_mux_park[RT_BLOCK + threadIdx.x] = make_float4(t.oi.x, t.oi.y, t.oi.z, __int_as_float(-1));
              rt_mux_park[2 * RT_BLOCK + threadIdx.x] = make_float4(__int_as_float(0), __int_as_float(0), __uint_as_float(mine), 0.f);
              parked = RT_SLOT_NODE;
            }
          }
        }
      }
    }
    if (__all_sync(0xffffffffu, cur.ref == RT_DONE && parked == RT_SLOT_EMPTY))
      break;

    for (;;) {
      // ---- node phase: a lane whose ray stops at a leaf (or ends) continues with its parked ray ----
      for (;;) {
        if (!(cur.ref >= 0 && cur.ref != RT_DONE) && parked == RT_SLOT_NODE)
          RT_MUX_SWAP();
        const bool work = cur.ref >= 0 && cur.ref != RT_DONE;
        unsigned int busy = __ballot_sync(0xffffffffu, work);
        if (__popc(busy) < RT_MUX_NODE_EXIT) {
          // the stragglers finish their current node run (as in the plain loop) unless leaf work is waiting
          if (busy == 0 || __any_sync(0xffffffffu, (cur.ref < 0) || parked == RT_SLOT_LEAF))
            break;
        }
        if (work) {
          if (STATS)
            n_nodes++;
          if (!node_visit(sc, cur.ref, rt, RT_T_MIN, cur.best.t, stack, cur.sp, cur.ref))
            if (!stack_pop(stack, cur.sp, cur.best.t, cur.ref))
              cur.ref = RT_DONE;
        }
      }
      // ---- leaf phase: the ray in registers, then the parked one ----
#pragma unroll 1
      for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
          if (!__any_sync(0xffffffffu, parked == RT_SLOT_LEAF))
            break;
          if (parked == RT_SLOT_LEAF)
            RT_MUX_SWAP();
        }
        if (cur.ref < 0) {
          float4 a = ray_a[cur.q], b = ray_b[cur.q];
          Ray r;
          r.o = F3(a.x, a.y, a.z);
          r.d = F3(b.x, b.y, b.z);
          r.time = a.w;
          int skip = bounce == 0 ? -1 : __float_as_int(hit[cur.q].y);
          RayKey key;
          key.seed = pp.seed;
          key.pixel = key.sample = 0;
          key.bounce = (uint32_t)bounce;
          if (has_media) {
            int k;
            path_to_key(pp, __float_as_int(b.w), bounce, key, k);
          }
          do {
            if (STATS)
              n_tests++;
            leaf_test(sc, ~cur.ref, r, RT_T_MIN, cur.best, skip, key);
            if (!stack_pop(stack, cur.sp, cur.best.t, cur.ref))
              cur.ref = RT_DONE;
          } while (cur.ref < 0);
        }
        if (cur.ref == RT_DONE && cur.best.t != -1.0f) { // a ray that ended writes its answer once and frees its slot
          hit[cur.q] = make_float2(cur.best.t, __int_as_float(cur.best.prim));
          cur.best.t = -1.0f;
        }
      }
      unsigned int node_work = __ballot_sync(0xffffffffu, (cur.ref >= 0 && cur.ref != RT_DONE) || parked == RT_SLOT_NODE);
      unsigned int empty = __popc(__ballot_sync(0xffffffffu, cur.ref == RT_DONE)) + __popc(__ballot_sync(0xffffffffu, parked == RT_SLOT_EMPTY));
      if (node_work == 0 || (!exhausted && empty >= RT_MUX_REFILL))
        break;
    }
  }
  traversal_stats<STATS>(stats, n_nodes, n_tests);
}
#undef RT_MUX_SWAP

// The straightforward variant (one ray per thread, grid stride, single if/else loop); kept for A/B
// measurements (RT_EXTEND=simple).
__global__ void __launch_bounds__(RT_BLOCK)
    k_extend_simple(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp,
                    const float4 *__restrict__ ray_a, const float4 *__restrict__ ray_b, float2 *__restrict__ hit,
                    const unsigned int *__restrict__ counts, int bounce, int has_media, unsigned long long *stats) {
  RT_DECLARE_STACK(stack);
  const int n = (int)counts[bounce];
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&stats[0], (unsigned long long)n);
  int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    float4 a = ray_a[q], b = ray_b[q];
    float2 h = hit[q];
    Ray r;
    r.o = F3(a.x, a.y, a.z);
    r.d = F3(b.x, b.y, b.z);
    r.time = a.w;
    RayKey key;
    key.seed = pp.seed;
    key.pixel = key.sample = 0;
    key.bounce = (uint32_t)bounce;
    if (has_media) {
      int k;
      path_to_key(pp, __float_as_int(b.w), bounce, key, k);
    }
    Hit best;
    best.t = RT_INF_F;
    best.prim = -1;
    traverse(sc, r, RT_T_MIN, best, bounce == 0 ? -1 : __float_as_int(h.y), key, stack);
    hit[q] = make_float2(best.t, __int_as_float(best.prim));
  }
}

// ---------------------------------------------------------------------------------------------------
// shade + queue compaction
// ---------------------------------------------------------------------------------------------------
#ifndef RT_SHADE_BLOCKS
#define RT_SHADE_BLOCKS 8 // resident blocks per SM (64 registers): k_shade is latency bound, occupancy wins over a few spills
#endif
#ifndef RT_SHADE_THREADS
#define RT_SHADE_THREADS 256
#endif
#define RT_SHADE_WARPS (RT_SHADE_THREADS / 32)
#ifndef RT_SHADE_SORT
#define RT_SHADE_SORT 0
#endif
#ifndef RT_SHADE_DYNAMIC
#define RT_SHADE_DYNAMIC 1
#endif
#if RT_SHADE_SORT && RT_SHADE_DYNAMIC
#error "RT_SHADE_SORT keeps the static partition: build with -DRT_SHADE_DYNAMIC=0"
#endif
// Queue compaction: every warp counts its continuing paths with a ballot, the block adds the counts up in
// shared memory and reserves the slots of all its warps with ONE atomicAdd on the next queue's length.  All
// atomics of a launch hit the same address and the L2 serialises them (~0.85 clocks each): one per warp
// (65 k per launch at 1080p) kept every warp waiting on that queue; one per block is 8 times fewer.
// GEN (bounce 0 of a pass whose first extend launch generated the camera rays): queue slot q is path q and its
// ray is re-derived from the path's Philox key instead of being read.
template <bool GEN>
__global__ void __launch_bounds__(RT_SHADE_THREADS, RT_SHADE_BLOCKS * RT_BLOCK / RT_SHADE_THREADS)
    k_shade(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp, const float4 *__restrict__ ray_a,
            const float4 *__restrict__ ray_b, const float2 *__restrict__ hit, float4 *__restrict__ next_a,
            float4 *__restrict__ next_b, float2 *__restrict__ next_hit, const float4 *__restrict__ thr,
            float4 *__restrict__ next_thr, float4 *__restrict__ radiance, unsigned int *__restrict__ counts,
            unsigned int *__restrict__ cursor, int bounce) {
  __shared__ unsigned int s_count[RT_SHADE_WARPS], s_first[RT_SHADE_WARPS];
  __shared__ unsigned int s_base; // the block's next 256 queue entries (dynamic fetch: RT_SHADE_DYNAMIC)
  rt_pdl_enter();
#if RT_SHADE_SORT
  __shared__ unsigned int s_sort[8 * RT_SHADE_WARPS];
#endif
  const int n = (int)counts[bounce];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool last_bounce = bounce + 1 >= pp.max_depth;
#if RT_SHADE_DYNAMIC
  // Blocks take their 256 entries from a device-side cursor instead of a fixed stride: iterations differ in cost
  // (misses are cheap, glass and textures are not), and with a static partition the SMs idled 15-20 % of the launch
  // waiting for the slowest blocks.  The next fetch rides on the barrier pair the compaction needs anyway.
  if (threadIdx.x == 0)
    s_base = atomicAdd(cursor, (unsigned int)RT_SHADE_THREADS);
  __syncthreads();
  for (int block_base = (int)s_base; block_base < n;) {
#else
  const int stride = gridDim.x * blockDim.x;
  // the same trip count for every warp of the block: the loop body has block-wide barriers
  for (int block_base = blockIdx.x * blockDim.x; block_base < n; block_base += stride) {
#endif
    int q = block_base + (int)threadIdx.x;
    bool active = q < n;
    bool cont = false;
    ShadeResult res;
    int path = 0;
    if (active) {
      // all three queue reads go out before anything waits on one of them
      float2 h = ldg2_now(hit + q);
      Ray r;
      Hit ht;
      ht.t = h.x;
      if (GEN) {
        path = pp.path_base + q;
        r = path_camera_ray(pp, path);
        // opaque to the optimiser from here on, like the loaded ray of the other instantiation
        asm volatile("" : "+f"(r.o.x), "+f"(r.o.y), "+f"(r.o.z), "+f"(r.d.x), "+f"(r.d.y), "+f"(r.d.z), "+f"(r.time));
        ht.prim = __float_as_int(h.y);
      } else {
        float4 a = ldg4_now(ray_a + q), b = ldg4_now(ray_b + q);
        path = __float_as_int(b.w);
        r.o = F3(a.x, a.y, a.z);
        r.d = F3(b.x, b.y, b.z);
        r.time = a.w;
        // a ray with a NaN time counts as a miss; the test also ties the miss branch to ray_a, which keeps
        // that read next to the other two instead of behind the branch (one exposed memory latency less)
        ht.prim = a.w == a.w ? __float_as_int(h.y) : -1;
      }
      RayKey key;
      int k;
      path_to_key(pp, path, bounce, key, k);
      // the path throughput travels with the ray, in queue order (sequential 16-byte accesses instead of a
      // per-path array read and written at random: half-used 32-byte sectors both ways)
      float4 tp = bounce == 0 ? make_float4(1.f, 1.f, 1.f, 0.f) : ldg4_now(thr + q);
      cont = shade_segment(sc, r, ht, F3(tp.x, tp.y, tp.z), key, last_bounce, res);
      if (!cont)
        path_ends(pp, radiance, path, k, res.radiance);
    }
#if RT_SHADE_SORT
    // Block-local counting sort of the continuing rays by direction octant: the block's slots are handed out
    // octant by octant (and warp by warp within an octant), so 32 consecutive queue entries - a warp of the next
    // extend launch - mostly share the signs of their direction, i.e. the order in which they walk the tree.
    static_assert(RT_SHADE_WARPS == 8, "the 64-entry scan below assumes 8 warps per block");
    const unsigned int oct = cont ? ((res.next.d.x < 0.f ? 1u : 0u) | (res.next.d.y < 0.f ? 2u : 0u) |
                                     (res.next.d.z < 0.f ? 4u : 0u))
                                  : 8u;
    unsigned int rank = 0, my_count = 0;
#pragma unroll
    for (unsigned int k = 0; k < 8; k++) {
      unsigned int m = __ballot_sync(0xffffffffu, oct == k);
      if (oct == k)
        rank = (unsigned int)__popc(m & ((1u << lane) - 1u));
      if ((unsigned int)lane == k)
        my_count = (unsigned int)__popc(m);
    }
    if (lane < 8)
      s_sort[lane * RT_SHADE_WARPS + warp] = my_count; // octant-major
    __syncthreads();
    if (warp == 0) {
      unsigned int a = s_sort[lane], b = s_sort[lane + 32];
      unsigned int ia = a, ib = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) {
          ia += ta;
          ib += tb;
        }
      }
      unsigned int total_a = __shfl_sync(0xffffffffu, ia, 31), total = total_a + __shfl_sync(0xffffffffu, ib, 31);
      unsigned int first = 0;
      if (lane == 0 && total)
        first = atomicAdd(&counts[bounce + 1], total);
      first = __shfl_sync(0xffffffffu, first, 0);
      s_sort[lane] = first + ia - a;
      s_sort[lane + 32] = first + total_a + ib - b;
    }
    __syncthreads();
    if (cont) {
      unsigned int slot = s_sort[oct * RT_SHADE_WARPS + warp] + rank;
      next_a[slot] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
      next_b[slot] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
      next_hit[slot] = make_float2(0.f, __int_as_float(res.next_skip_prim));
      next_thr[slot] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
    }
    __syncthreads(); // s_sort is rewritten by the next iteration
    continue;
#endif
    unsigned int mask = __ballot_sync(0xffffffffu, cont);
    if (lane == 0)
      s_count[warp] = (unsigned int)__popc(mask);
    __syncthreads();
    if (warp == 0) { // exclusive prefix over the warps' counts + the block's reservation
      unsigned int c = lane < RT_SHADE_WARPS ? s_count[lane] : 0u;
      unsigned int incl = c;
#pragma unroll
      for (int o = 1; o < RT_SHADE_WARPS; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
          incl += t;
      }
      unsigned int total = __shfl_sync(0xffffffffu, incl, RT_SHADE_WARPS - 1);
      unsigned int first = 0;
      if (lane == 0 && total)
        first = atomicAdd(&counts[bounce + 1], total);
      first = __shfl_sync(0xffffffffu, first, 0);
      if (lane < RT_SHADE_WARPS)
        s_first[lane] = first + incl - c;
#if RT_SHADE_DYNAMIC
      if (lane == 0)
        s_base = atomicAdd(cursor, (unsigned int)RT_SHADE_THREADS);
#endif
    }
    __syncthreads();
#if RT_SHADE_DYNAMIC
    block_base = (int)s_base; // read by every thread before warp 0 can rewrite it behind the next iteration's first barrier
#endif
    if (cont) {
      unsigned int slot = s_first[warp] + (unsigned int)__popc(mask & ((1u << lane) - 1u));
      next_a[slot] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
      next_b[slot] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
      next_hit[slot] = make_float2(0.f, __int_as_float(res.next_skip_prim));
      next_thr[slot] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// tail: the thin end of the path population in ONE launch
// ---------------------------------------------------------------------------------------------------
// After the first bounces the path population thins out (C2: the bounces from the fourth on are 15 % of the
// frame's segments, spread over five bounces), and every wavefront launch ends with its slowest rays - a cost
// that no longer amortises over a small queue and that grows with the scene (rt_api.cu, wave_depth).  k_tail takes the queue of bounce `first_bounce` and runs every remaining path to
// its end inside the kernel: persistent warps fetch paths dynamically, traverse (same while-while loop as
// k_extend) and, when too few lanes are still traversing, shade the finished segments in place; a lane
// whose path continues re-enters traversal with the scattered ray (written back to its own queue slot), a
// lane whose path ended fetches the next path.  Philox keys carry the lane's own bounce index, so the
// image is identical to the all-wavefront schedule.
#ifndef RT_TAIL_NOINLINE
#define RT_TAIL_NOINLINE 0 // 1: the tail kernel calls the shading step as a function (its registers are then live only
                           // during the call, so the traversal loop can run at a higher occupancy)
#endif
#if RT_TAIL_NOINLINE
__device__ __noinline__ bool shade_segment_call(const DScene &sc, const Ray &ray, Hit hit, f3 throughput, const RayKey &key,
                                                bool last_bounce, ShadeResult &out) {
  return shade_segment(sc, ray, hit, throughput, key, last_bounce, out);
}
#else
#define shade_segment_call shade_segment
#endif

// SHARE: end-game sharing of long traversals between the lanes of a warp (below); the launch wrapper turns it on for
// scenes whose stragglers matter (rt_kernels.cu launch_tail), the others run the instantiation without its bookkeeping.
template <bool STATS, bool SHARE, bool TIES>
__global__ void __launch_bounds__(RT_BLOCK, RT_TAIL_BLOCKS)
    k_tail(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp, float4 *__restrict__ ray_a,
           float4 *__restrict__ ray_b, float2 *__restrict__ hit, float4 *__restrict__ next_a,
           float4 *__restrict__ next_b, float2 *__restrict__ next_hit, float4 *__restrict__ thr,
           float4 *__restrict__ next_thr, float4 *__restrict__ radiance, unsigned int *__restrict__ counts,
           unsigned int *__restrict__ cursor,
           int first_bounce, int end_bounce, int has_media, unsigned long long *stats) {
  rt_pdl_enter();
  RT_DECLARE_STACK(stack);
  const unsigned int n = counts[first_bounce];
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;

  RayTrav rt = make_trav(F3(0.f, 0.f, 0.f), F3(1.f, 1.f, 1.f));
  Hit best;
  int sp = 0, ref = RT_DONE, bounce = first_bounce;
  unsigned int q = 0, segments = 0, n_nodes = 0, n_tests = 0, ray_nodes = 0;
  bool exhausted = false;
  best.t = -1.0f; // no segment held
  best.prim = -1;
  // End-game sharing (below): `owner` = the lane whose segment this lane is tracing (itself, or the lane it helps),
  // `helpers` = how many lanes are still tracing subtrees of this lane's own segment.  (Dead without SHARE.)
  // `lo` = floor of this lane's stack: entries below it were handed to helpers.
  int owner = (int)lane, helpers = 0, lo = 0;

  for (;;) {
    // ---- end-game sharing: once the queue is empty, a warp's last long traversals would run on one lane each while
    // the other lanes - and soon most of the GPU - idle (10^6 spheres: the SMs were active 72 % of the launch).  A
    // lane without work takes one pending subtree (a stack entry) of the busiest lane and traces it with that lane's
    // ray; results come back through shuffles and are merged with leaf_test's order-independent rule, so the answer
    // does not depend on who traced what.
    if (SHARE) {
      // helpers that are done hand their result to the owner and become free lanes
      // (a helper that has given subtrees away itself waits for those first)
      unsigned int done_helpers = __ballot_sync(0xffffffffu, owner != (int)lane && ref == RT_DONE && helpers == 0);
      while (done_helpers) {
        const int h = __ffs(done_helpers) - 1;
        done_helpers &= done_helpers - 1;
        const int o = __shfl_sync(0xffffffffu, owner, h);
        const float ht = __shfl_sync(0xffffffffu, best.t, h);
        const int hp = __shfl_sync(0xffffffffu, best.prim, h);
        if ((int)lane == o) {
          helpers--;
          if (hp >= 0 && hp != best.prim && closer_hit(sc, ht, hp, best)) {
            best.t = ht;
            best.prim = hp;
          }
        }
        if ((int)lane == h) {
          owner = (int)lane;
          best.t = -1.0f;
          best.prim = -1;
        }
      }
      if (exhausted) {
        const unsigned int free_lanes = __ballot_sync(0xffffffffu, ref == RT_DONE && best.t == -1.0f);
        // the donor: the lane with the most pending subtrees in the shared-memory part of its stack; it gives away the
        // OLDEST ones (pushed nearest to the root: the largest subtrees, and entries that a deep traversal would only
        // come back to at its very end)
        const int avail = ref != RT_DONE ? (sp < RT_STACK_SMEM ? sp : RT_STACK_SMEM) - lo : 0;
        const int key = avail >= 1 ? ((avail << 5) | (int)lane) : -1;
        const int top = __reduce_max_sync(0xffffffffu, key);
        if (free_lanes && top >= 0) {
          const int donor = top & 31, davail = top >> 5;
          const int n_free = __popc(free_lanes);
          const int k = n_free < davail ? n_free : davail;
          const int rank = __popc(free_lanes & lt_mask);
          const bool take = ((free_lanes >> lane) & 1u) && rank < k;
          __syncwarp(); // the donor's pushes are visible to the lanes that read its column
          // the donor's ray, interval and identity
          const float ix = __shfl_sync(0xffffffffu, rt.inv.x, donor), iy = __shfl_sync(0xffffffffu, rt.inv.y, donor),
                      iz = __shfl_sync(0xffffffffu, rt.inv.z, donor), ox = __shfl_sync(0xffffffffu, rt.oi.x, donor),
                      oy = __shfl_sync(0xffffffffu, rt.oi.y, donor), oz = __shfl_sync(0xffffffffu, rt.oi.z, donor);
          const float dt = __shfl_sync(0xffffffffu, best.t, donor);
          const int dp = __shfl_sync(0xffffffffu, best.prim, donor);
          const unsigned int dq = __shfl_sync(0xffffffffu, q, donor);
          const int db = __shfl_sync(0xffffffffu, bounce, donor);
          const int dlo = __shfl_sync(0xffffffffu, lo, donor);
          if (take) {
            rt = trav_from(F3(ix, iy, iz), F3(ox, oy, oz));
            best.t = dt;
            best.prim = dp;
            q = dq;
            bounce = db;
            owner = donor;
            sp = 0;
            lo = 0;
            // entry dlo + rank of the donor's column (columns are 8 bytes apart, rows 1 KB)
            int eref;
            float et;
            stack.get_from(donor - (int)lane, dlo + rank, eref, et);
            ref = et <= best.t ? eref : RT_DONE;
          }
          __syncwarp(); // ... and read before the donor pushes over them
          if ((int)lane == donor) {
            lo += k;
            helpers += k;
          }
        }
      }
    }
    // ---- fetch: lanes without a path take the next ones of the queue ----
    unsigned int idle = __ballot_sync(0xffffffffu, ref == RT_DONE && best.t == -1.0f);
    if (idle && !exhausted) {
      unsigned int base = 0;
      int leader = __ffs(idle) - 1;
      if ((int)lane == leader)
        base = atomicAdd(cursor, (unsigned int)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (ref == RT_DONE && best.t == -1.0f) {
        unsigned int mine = base + (unsigned int)__popc(idle & lt_mask);
        if (mine < n) {
          q = mine;
          float4 a = ray_a[q], b = ray_b[q];
          rt = make_trav(F3(a.x, a.y, a.z), F3(b.x, b.y, b.z));
          best.t = RT_INF_F;
          best.prim = -1;
          sp = 0;
          lo = 0;
          ref = 0;
          bounce = first_bounce;
          segments++;
        }
      }
      exhausted = base + (unsigned int)__popc(idle) >= n;
    }
    if (__all_sync(0xffffffffu, ref == RT_DONE && best.t == -1.0f))
      break;

    // ---- traverse until too few lanes are still inside the tree ----
    for (;;) {
      while (ref >= 0 && ref != RT_DONE) {
        if (STATS) {
          n_nodes++;
          ray_nodes++;
        }
        if (!node_visit(sc, ref, rt, RT_T_MIN, best.t, stack, sp, ref))
          if (!stack_pop(stack, sp, best.t, ref, SHARE ? lo : 0))
            ref = RT_DONE;
      }
      if (ref < 0) {
        float4 a = ray_a[q], b = ray_b[q];
        Ray r;
        r.o = F3(a.x, a.y, a.z);
        r.d = F3(b.x, b.y, b.z);
        r.time = a.w;
        int skip = bounce == 0 ? -1 : __float_as_int(hit[q].y);
        RayKey key;
        key.seed = pp.seed;
        key.pixel = key.sample = 0;
        key.bounce = (uint32_t)bounce;
        if (has_media) {
          int k;
          path_to_key(pp, __float_as_int(b.w), bounce, key, k);
        }
        do {
          if (STATS)
            n_tests++;
          leaf_test<TIES>(sc, ~ref, r, RT_T_MIN, best, skip, key); // sharing needs the order-independent tie rule
          if (!stack_pop(stack, sp, best.t, ref, SHARE ? lo : 0))
            ref = RT_DONE;
        } while (ref < 0);
      }
      unsigned int busy = __ballot_sync(0xffffffffu, ref != RT_DONE);
      if (__popc(busy) < RT_TAIL_REFILL)
        break;
    }

    // ---- shade the segments whose traversal is complete ----
    if (ref == RT_DONE && best.t != -1.0f && (!SHARE || (owner == (int)lane && helpers == 0))) {
      ray_stats<STATS>(stats, ray_nodes);
      ray_nodes = 0;
      float4 a = ray_a[q], b = ray_b[q];
      int path = __float_as_int(b.w);
      Ray r;
      r.o = F3(a.x, a.y, a.z);
      r.d = F3(b.x, b.y, b.z);
      r.time = a.w;
      RayKey key;
      int k;
      path_to_key(pp, path, bounce, key, k);
      float4 tp = bounce == 0 ? make_float4(1.f, 1.f, 1.f, 0.f) : thr[q]; // the lane's own queue slot
      ShadeResult res;
      bool cont = shade_segment_call(sc, r, best, F3(tp.x, tp.y, tp.z), key, bounce + 1 >= pp.max_depth, res);
      if (cont && bounce + 1 >= end_bounce) {
        // the launch covers bounces [first_bounce, end_bounce): survivors are queued for the next launch,
        // so one very long path (glass, mirrors) cannot keep a whole launch waiting on a single lane
        unsigned int slot = atomicAdd(&counts[end_bounce], 1u);
        next_a[slot] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
        next_b[slot] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
        next_hit[slot] = make_float2(0.f, __int_as_float(res.next_skip_prim));
        next_thr[slot] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
        best.t = -1.0f;
      } else if (cont) {
        ray_a[q] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
        ray_b[q] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
        hit[q] = make_float2(0.f, __int_as_float(res.next_skip_prim));
        thr[q] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
        rt = make_trav(res.next.o, res.next.d);
        best.t = RT_INF_F;
        best.prim = -1;
        sp = 0;
        lo = 0;
        ref = 0;
        bounce++;
        segments++;
      } else {
        path_ends(pp, radiance, path, k, res.radiance);
        best.t = -1.0f;
      }
    }
  }
  // segments traced by this warp -> stats[3]
  for (int o = 16; o > 0; o >>= 1)
    segments += __shfl_xor_sync(0xffffffffu, segments, o);
  if (lane == 0 && segments)
    atomicAdd(&stats[3], (unsigned long long)segments);
  traversal_stats<STATS>(stats, n_nodes, n_tests);
}

// ---------------------------------------------------------------------------------------------------
// tail with warp-local regrouping (k_tail_regroup)
// ---------------------------------------------------------------------------------------------------
// k_tail couples a lane to its path: a lane whose segment is traced waits until fewer than RT_TAIL_REFILL lanes
// are still traversing, then shades with whoever else is waiting, then re-enters traversal (ncu: 10-11 of 32 lanes
// per instruction, the lowest of all kernels).  Here the warp owns a small POOL of paths instead:
//   * a lane that finishes a segment drops (slot, t, primitive, bounce) into the warp's pending list in shared
//     memory and immediately takes another ray - first from the warp's own list of scattered rays, then from the
//     launch's queue - so the traversal loop always runs with at least RT_REFILL busy lanes while work exists;
//   * shading runs when 32 segments are pending: a full warp of hits, whatever their bounces, at 32 of 32 lanes;
//     the scattered rays go back to their own queue slots and onto the warp's to-trace list;
//   * the traversal registers of the rays in flight are parked in shared memory around the shading code, so the
//     two register-hungry phases do not add up (64 registers, the occupancy of k_extend, instead of 96).
// A warp holds at most 32 rays in flight + 63 pending hits + the to-trace list; it only fetches from the launch's
// queue when its to-trace list cannot fill its idle lanes, so the three together never exceed 95 paths.
// Same segments, same Philox keys (per-path bounce index), same film sums: images are bit-identical to k_tail's.
#ifndef RT_TAIL2_BLOCKS
#define RT_TAIL2_BLOCKS 8
#endif
#ifndef RT_TAIL2_PARTIAL
#define RT_TAIL2_PARTIAL 16 // with nothing left to fetch: shade a partial batch once this many hits are pending ...
#endif
#ifndef RT_TAIL2_LOW
#define RT_TAIL2_LOW 8 // ... or when fewer lanes than this are still traversing
#endif
#define RT_PEND_CAP 64
#define RT_TRAV_CAP 96
#define RT_WARPS_PER_BLOCK (RT_BLOCK / 32)
__shared__ int4 rt_pend_smem[RT_WARPS_PER_BLOCK][RT_PEND_CAP];
__shared__ int2 rt_trav_smem[RT_WARPS_PER_BLOCK][RT_TRAV_CAP];
__shared__ float4 rt_park_smem[3][RT_BLOCK];

template <bool STATS>
__global__ void __launch_bounds__(RT_BLOCK, RT_TAIL2_BLOCKS)
    k_tail_regroup(const __grid_constant__ DScene sc, const __grid_constant__ PassParams pp, float4 *__restrict__ ray_a,
                   float4 *__restrict__ ray_b, float2 *__restrict__ hit, float4 *__restrict__ next_a,
                   float4 *__restrict__ next_b, float2 *__restrict__ next_hit, float4 *__restrict__ thr,
                   float4 *__restrict__ next_thr, float4 *__restrict__ radiance, unsigned int *__restrict__ counts,
                   unsigned int *__restrict__ cursor, int first_bounce, int end_bounce, int has_media,
                   unsigned long long *stats) {
  RT_DECLARE_STACK(stack);
  const unsigned int n = counts[first_bounce];
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;
  int4 *pend = rt_pend_smem[threadIdx.x >> 5];
  int2 *trav = rt_trav_smem[threadIdx.x >> 5];

  RayTrav rt = make_trav(F3(0.f, 0.f, 0.f), F3(1.f, 1.f, 1.f));
  Hit best;
  int sp = 0, ref = RT_DONE, bounce = first_bounce;
  unsigned int q = 0, segments = 0, n_nodes = 0, n_tests = 0;
  int n_pend = 0, n_trav = 0; // warp-uniform
  bool exhausted = false;     // the launch's queue has no more paths to hand out
  best.t = -1.0f;
  best.prim = -1;

  for (;;) {
    // ---- refill: lanes without a ray take one, the warp's own scattered rays first ----
    unsigned int idle = __ballot_sync(0xffffffffu, ref == RT_DONE);
    if (idle && (n_trav > 0 || !exhausted)) {
      const int n_idle = __popc(idle), rank = __popc(idle & lt_mask);
      const int local = n_idle < n_trav ? n_idle : n_trav;
      const int wanted = n_idle - local;
      unsigned int base = 0;
      if (wanted && !exhausted) {
        int leader = __ffs(idle) - 1;
        if ((int)lane == leader)
          base = atomicAdd(cursor, (unsigned int)wanted);
        base = __shfl_sync(0xffffffffu, base, leader);
        exhausted = base + (unsigned int)wanted >= n;
      } else {
        base = n; // nothing to fetch
      }
      if (ref == RT_DONE) {
        bool got = false;
        if (rank < local) {
          int2 e = trav[n_trav - 1 - rank];
          q = (unsigned int)e.x;
          bounce = e.y;
          got = true;
        } else {
          unsigned int mine = base + (unsigned int)(rank - local);
          if (mine < n) {
            q = mine;
            bounce = first_bounce;
            got = true;
          }
        }
        if (got) {
          float4 a = ray_a[q], b = ray_b[q];
          rt = make_trav(F3(a.x, a.y, a.z), F3(b.x, b.y, b.z));
          best.t = RT_INF_F;
          best.prim = -1;
          sp = 0;
          ref = 0;
          segments++;
        }
      }
      n_trav -= local;
      __syncwarp();
    }
    unsigned int busy = __ballot_sync(0xffffffffu, ref != RT_DONE);

    // ---- shade: a full warp of pending hits, or what is left when nothing else can feed the idle lanes ----
    const bool no_source = n_trav == 0 && exhausted;
    if (n_pend >= 32 || (n_pend > 0 && no_source && (n_pend >= RT_TAIL2_PARTIAL || __popc(busy) < RT_TAIL2_LOW))) {
      // park the rays in flight: nothing of the traversal state stays in registers across the shading code
      rt_park_smem[0][threadIdx.x] = make_float4(rt.inv.x, rt.inv.y, rt.inv.z, best.t);
      rt_park_smem[1][threadIdx.x] = make_float4(rt.oi.x, rt.oi.y, rt.oi.z, __int_as_float(best.prim));
      rt_park_smem[2][threadIdx.x] = make_float4(__int_as_float(sp), __int_as_float(ref), __uint_as_float(q), __int_as_float(bounce));
      const int batch = n_pend < 32 ? n_pend : 32;
      n_pend -= batch;
      bool cont = false, to_next = false;
      ShadeResult res;
      unsigned int eq = 0;
      int ebounce = 0, path = 0;
      if ((int)lane < batch) {
        int4 e = pend[n_pend + (int)lane];
        eq = (unsigned int)e.x;
        ebounce = e.w;
        Hit h;
        h.t = __int_as_float(e.y);
        h.prim = e.z;
        float4 a = ray_a[eq], b = ray_b[eq];
        path = __float_as_int(b.w);
        Ray r;
        r.o = F3(a.x, a.y, a.z);
        r.d = F3(b.x, b.y, b.z);
        r.time = a.w;
        RayKey key;
        int k;
        path_to_key(pp, path, ebounce, key, k);
        float4 tp = ebounce == 0 ? make_float4(1.f, 1.f, 1.f, 0.f) : thr[eq];
        cont = shade_segment(sc, r, h, F3(tp.x, tp.y, tp.z), key, ebounce + 1 >= pp.max_depth, res);
        if (!cont)
          path_ends(pp, radiance, path, k, res.radiance);
        to_next = cont && ebounce + 1 >= end_bounce;
      }
      // survivors of the launch's last bounce are queued for the next launch (one atomic per warp)
      unsigned int m_next = __ballot_sync(0xffffffffu, to_next);
      if (m_next) {
        unsigned int first = 0;
        int leader = __ffs(m_next) - 1;
        if ((int)lane == leader)
          first = atomicAdd(&counts[end_bounce], (unsigned int)__popc(m_next));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (to_next) {
          unsigned int slot = first + (unsigned int)__popc(m_next & lt_mask);
          next_a[slot] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
          next_b[slot] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
          next_hit[slot] = make_float2(0.f, __int_as_float(res.next_skip_prim));
          next_thr[slot] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
        }
      }
      // the others go back to their own queue slots and onto the warp's to-trace list
      const bool stay = cont && !to_next;
      unsigned int m_stay = __ballot_sync(0xffffffffu, stay);
      if (stay) {
        ray_a[eq] = make_float4(res.next.o.x, res.next.o.y, res.next.o.z, res.next.time);
        ray_b[eq] = make_float4(res.next.d.x, res.next.d.y, res.next.d.z, __int_as_float(path));
        hit[eq] = make_float2(0.f, __int_as_float(res.next_skip_prim));
        thr[eq] = make_float4(res.throughput.x, res.throughput.y, res.throughput.z, 0.f);
        trav[n_trav + __popc(m_stay & lt_mask)] = make_int2((int)eq, ebounce + 1);
      }
      n_trav += __popc(m_stay);
      __syncwarp(); // the list entries and the rays in the queue slots are visible to the lanes that take them
      // unpark
      float4 p0 = rt_park_smem[0][threadIdx.x], p1 = rt_park_smem[1][threadIdx.x], p2 = rt_park_smem[2][threadIdx.x];
      MuxRay m;
      m.inv = F3(p0.x, p0.y, p0.z);
      m.oi = F3(p1.x, p1.y, p1.z);
      rt = mux_trav(m);
      best.t = p0.w;
      best.prim = __float_as_int(p1.w);
      sp = __float_as_int(p2.x);
      ref = __float_as_int(p2.y);
      q = __float_as_uint(p2.z);
      bounce = __float_as_int(p2.w);
      continue; // refill from the new to-trace entries
    }
    if (busy == 0)
      break; // nothing in flight, nothing pending, nothing to fetch

    // ---- traverse until a lane count or a list length asks for one of the steps above ----
    for (;;) {
      while (ref >= 0 && ref != RT_DONE) {
        if (STATS)
          n_nodes++;
        if (!node_visit(sc, ref, rt, RT_T_MIN, best.t, stack, sp, ref))
          if (!stack_pop(stack, sp, best.t, ref))
            ref = RT_DONE;
      }
      if (ref < 0) {
        float4 a = ray_a[q], b = ray_b[q];
        Ray r;
        r.o = F3(a.x, a.y, a.z);
        r.d = F3(b.x, b.y, b.z);
        r.time = a.w;
        int skip = bounce == 0 ? -1 : __float_as_int(hit[q].y);
        RayKey key;
        key.seed = pp.seed;
        key.pixel = key.sample = 0;
        key.bounce = (uint32_t)bounce;
        if (has_media) {
          int k;
          path_to_key(pp, __float_as_int(b.w), bounce, key, k);
        }
        do {
          if (STATS)
            n_tests++;
          leaf_test(sc, ~ref, r, RT_T_MIN, best, skip, key);
          if (!stack_pop(stack, sp, best.t, ref))
            ref = RT_DONE;
        } while (ref < 0);
      }
      // a traced segment leaves the lane at once
      const bool finished = ref == RT_DONE && best.t != -1.0f;
      unsigned int m_fin = __ballot_sync(0xffffffffu, finished);
      if (finished) {
        pend[n_pend + __popc(m_fin & lt_mask)] = make_int4((int)q, __float_as_int(best.t), best.prim, bounce);
        best.t = -1.0f;
      }
      n_pend += __popc(m_fin);
      busy = __ballot_sync(0xffffffffu, ref != RT_DONE);
      if (busy == 0 || n_pend >= 32)
        break;
      if (__popc(busy) < RT_REFILL && (n_trav > 0 || !exhausted || n_pend >= RT_TAIL2_PARTIAL || __popc(busy) < RT_TAIL2_LOW))
        break;
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1)
    segments += __shfl_xor_sync(0xffffffffu, segments, o);
  if (lane == 0 && segments)
    atomicAdd(&stats[3], (unsigned long long)segments);
  traversal_stats<STATS>(stats, n_nodes, n_tests);
}

// ---------------------------------------------------------------------------------------------------
// accumulate / resolve
// ---------------------------------------------------------------------------------------------------
__global__ void k_accumulate(const __grid_constant__ PassParams pp, const float4 *__restrict__ radiance,
                             float4 *__restrict__ film) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pp.n_owned; j += stride) {
    RayKey key;
    int k; // film index of the pixel whose paths are j, j + n_owned, ...
    path_to_key(pp, j, 0, key, k);
    float4 acc = film[k];
    for (int s = 0; s < pp.n_samples; s++) {
      float4 r = radiance[(size_t)s * pp.n_owned + j];
      acc.x += r.x;
      acc.y += r.y;
      acc.z += r.z;
    }
    film[k] = acc;
  }
}

// to_byte(scale * sum) per channel in the reference's FP64 (ColorUtility.hpp:11-26,
// DynamicCamera.cpp:280-306).
__global__ void k_resolve_rgb8(const float4 *__restrict__ film, long long n, double scale, uint8_t *__restrict__ out) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    float4 v = film[k];
    out[k * 3 + 0] = to_byte_f64(scale * (double)v.x);
    out[k * 3 + 1] = to_byte_f64(scale * (double)v.y);
    out[k * 3 + 2] = to_byte_f64(scale * (double)v.z);
  }
}

__global__ void k_resolve_rgb(const float4 *__restrict__ film, long long n, double scale, float *__restrict__ out) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    float4 v = film[k];
    out[k * 3 + 0] = (float)(scale * (double)v.x);
    out[k * 3 + 1] = (float)(scale * (double)v.y);
    out[k * 3 + 2] = (float)(scale * (double)v.z);
  }
}

// Rank-major gathered compact films -> one row-major image (float4 radiance sums or RGB8 frames).
#define RT_MAX_RANKS 64
struct RankBases {
  long long first_pixel[RT_MAX_RANKS]; // offset of rank r's block in the gathered buffer, in pixels
};
struct Rgb8 {
  unsigned char r, g, b;
};

template <class Pixel>
__global__ void k_scatter_gathered(int width, int height, int n_ranks, int tile_rows, const __grid_constant__ RankBases bases,
                                   const Pixel *__restrict__ gathered, Pixel *__restrict__ full) {
  long long total = (long long)width * height;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
    int row = (int)(g / width), col = (int)(g - (long long)row * width);
    int tile = row / tile_rows;
    int rank = tile % n_ranks;
    int tile_local = tile / n_ranks;
    // rows this rank owns before `row`: full tiles before tile_local (only the image's last tile can be short)
    int local_row = tile_local * tile_rows + (row - tile * tile_rows);
    full[g] = gathered[bases.first_pixel[rank] + (long long)local_row * width + col];
  }
}

// ---------------------------------------------------------------------------------------------------
// parity hook: FP32 closest hit for explicit rays, through the same traverse() as k_extend
// ---------------------------------------------------------------------------------------------------
template <bool TIES>
__global__ void __launch_bounds__(RT_BLOCK)
    k_trace_fast(const __grid_constant__ DScene sc, const rt_ray *__restrict__ rays, long long n, uint64_t seed,
                 const int *__restrict__ leaf_object, const int *__restrict__ leaf_id, rt_hit *__restrict__ hits) {
  RT_DECLARE_STACK(stack);
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const rt_ray &in = rays[q];
    Ray r;
    r.o = F3((float)in.origin[0], (float)in.origin[1], (float)in.origin[2]);
    r.d = F3((float)in.direction[0], (float)in.direction[1], (float)in.direction[2]);
    r.time = (float)in.time;
    RayKey key;
    key.seed = seed;
    key.pixel = in.rng_pixel;
    key.sample = in.rng_sample;
    key.bounce = in.rng_bounce;
    Hit best;
    best.t = (float)in.t_max;
    best.prim = -1;
    traverse<SmemStack, TIES>(sc, r, (float)in.t_min, best, -1, key, stack);
    rt_hit out;
    out.t = best.prim >= 0 ? (double)best.t : (double)RT_INF_F;
    out.prim = best.prim >= 0 ? leaf_id[best.prim] : -1;
    out.object = best.prim >= 0 ? leaf_object[best.prim] : -1;
    out.front_face = 0;
    out.pad_ = 0;
    if (best.prim >= 0) { // front face as the shader derives it
      const float4 *rec = sc.prims + (size_t)best.prim * RT_PRIM_F4;
      float4 r0 = rec[0], r3 = rec[3];
      int type = (uint32_t)__float_as_int(r3.y) >> 28;
      if (type == RT_PT_SPHERE) {
        float4 r1 = rec[1];
        f3 center = F3(r0) + r.time * F3(r1);
        f3 p = r.o + best.t * r.d;
        out.front_face = dot(r.d, p - center) < 0.f;
      } else if (type == RT_PT_QUAD) {
        out.front_face = dot(r.d, F3(r0)) < 0.f;
      } else {
        out.front_face = 1;
      }
    }
    hits[q] = out;
  }
}

// ---------------------------------------------------------------------------------------------------
// LBVH build kernels
// ---------------------------------------------------------------------------------------------------
__global__ void k_morton(const BuildBox *__restrict__ boxes, int n, const float *__restrict__ scene_lo,
                         const float *__restrict__ scene_inv, uint64_t *__restrict__ codes, uint32_t *__restrict__ index) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  codes[i] = morton_body(boxes[i], scene_lo, scene_inv);
  index[i] = (uint32_t)i;
}

__global__ void k_gather_boxes(const BuildBox *__restrict__ in, const uint32_t *__restrict__ index, BuildBox *__restrict__ out,
                               int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[index[i]];
}

__global__ void k_gather_records(const uint4 *__restrict__ in, const uint32_t *__restrict__ index, uint4 *__restrict__ out, int n,
                                 int vec_per) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)n * vec_per;
  if (t >= total)
    return;
  int i = (int)(t / vec_per), v = (int)(t - (long long)i * vec_per);
  out[(size_t)i * vec_per + v] = in[(size_t)index[i] * vec_per + v];
}

__global__ void k_hierarchy(const uint64_t *__restrict__ codes, BinTree t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < t.n - 1)
    hierarchy_body(codes, t, i);
}

__global__ void k_refit(BinTree t, const BuildBox *__restrict__ leaf_boxes) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t.n)
    return;
  int n_internal = t.n - 1;
  int node = t.parent[n_internal + j];
  while (node >= 0) {
    __threadfence();
    unsigned int old = atomicAdd(&t.visits[node], 1u);
    if (old == 0)
      return; // the sibling subtree is not finished yet; its thread will continue upwards
    __threadfence();
    const volatile BuildBox *vb = t.box;
    int l = t.left[node], r = t.right[node];
    BuildBox a, b;
    if (l >= 0) {
      for (int k = 0; k < 3; k++) {
        a.lo[k] = vb[l].lo[k];
        a.hi[k] = vb[l].hi[k];
      }
    } else {
      a = leaf_boxes[~l];
    }
    if (r >= 0) {
      for (int k = 0; k < 3; k++) {
        b.lo[k] = vb[r].lo[k];
        b.hi[k] = vb[r].hi[k];
      }
    } else {
      b = leaf_boxes[~r];
    }
    t.box[node] = box_union(a, b);
    node = t.parent[node];
  }
}

__global__ void k_collapse(BinTree t, const BuildBox *__restrict__ leaf_boxes, float4 *__restrict__ nodes,
                           const CollapseItem *__restrict__ items, int n_items, CollapseItem *__restrict__ next,
                           int *next_count, int *wide_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items)
    return;
  CollapseItem it = items[i];
  int child[4];
  int n_child = collapse_gather(t, leaf_boxes, it.bin, child);
  int n_inner = 0;
  for (int k = 0; k < n_child; k++)
    n_inner += child[k] >= 0;
  int wide_ref[4] = {0, 0, 0, 0};
  if (n_inner) {
    int first_wide = atomicAdd(wide_count, n_inner);
    int first_slot = atomicAdd(next_count, n_inner);
    int m = 0;
    for (int k = 0; k < n_child; k++)
      if (child[k] >= 0) {
        wide_ref[k] = first_wide + m;
        next[first_slot + m].bin = child[k];
        next[first_slot + m].wide = first_wide + m;
        next[first_slot + m].up = it.wide * 4 + k;
        m++;
      }
  }
  collapse_write(t, leaf_boxes, nodes, it.wide, child, n_child, wide_ref, it.up);
}

// ---- PLOC rounds (rt_bvh.h) ----
__global__ void k_ploc_init(const BuildBox *__restrict__ leaf_boxes, int n, PlocCluster *__restrict__ clusters) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n)
    clusters[j] = PlocCluster{leaf_boxes[j], ~j, 0};
}

__global__ void k_ploc_nearest(const PlocCluster *__restrict__ clusters, int count, int *__restrict__ nearest) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count)
    nearest[i] = ploc_nearest_body(clusters, count, i);
}

// high word: this cluster creates a node; low word: this cluster has a successor.  One inclusive scan of the
// packed words numbers both the new nodes and the slots of the next round.
__global__ void k_ploc_roles(const int *__restrict__ nearest, int count, unsigned long long *__restrict__ packed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    int role = ploc_role(nearest, i);
    packed[i] = ((unsigned long long)(role > 0) << 32) | (unsigned long long)(role >= 0);
  }
}

__global__ void k_ploc_merge(const PlocCluster *__restrict__ clusters, const int *__restrict__ nearest,
                             const unsigned long long *__restrict__ scanned, int count, int first_node,
                             PlocCluster *__restrict__ next, BinTree t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
    return;
  int role = ploc_role(nearest, i);
  unsigned long long incl = scanned[i];
  int merges_before = (int)(incl >> 32) - (role > 0), slot = (int)(incl & 0xffffffffu) - (role >= 0);
  ploc_merge_body(clusters, nearest, i, role, slot, first_node - merges_before, next, t);
}

// Sum of the surface areas of the binary tree's internal nodes (the tree-dependent part of its SAH cost).
__global__ void k_tree_area(const BuildBox *__restrict__ box, int n_internal, double *__restrict__ sum) {
  double local = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_internal; i += gridDim.x * blockDim.x)
    local += (double)box_area(box[i]);
  for (int o = 16; o > 0; o >>= 1)
    local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local != 0.0)
    atomicAdd(sum, local);
}

// ---- refit of the BVH4 after primitive updates (rt_scene_update_spheres) ----
__global__ void k_leaf_links(const float4 *__restrict__ nodes, int n_nodes, int *__restrict__ leaf_up) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes)
    leaf_links_body(nodes, i, leaf_up);
}

// Scatters `count` updated records to their leaves and writes each new leaf box into its parent's slot.
__global__ void k_update_leaves(const float4 *__restrict__ records, const PrimExact *__restrict__ exact,
                                const BuildBox *__restrict__ boxes, const int *__restrict__ leaf, int count,
                                const int *__restrict__ leaf_up, float4 *__restrict__ prims,
                                PrimExact *__restrict__ ex_prims, float4 *__restrict__ nodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
    return;
  int j = leaf[i];
  for (int k = 0; k < RT_PRIM_F4; k++)
    prims[(size_t)j * RT_PRIM_F4 + k] = records[(size_t)i * RT_PRIM_F4 + k];
  ex_prims[j] = exact[i];
  int up = leaf_up[j];
  node_set_slot_box(nodes, up >> 2, up & 3, boxes[i]);
}

// One thread per leaf climbs towards the root; at every node the last child to arrive recomputes the node's
// box from its slots and stores it in the parent's slot (every node is completed exactly once).
__global__ void k_refit_wide(float4 *nodes, const int *__restrict__ leaf_up, unsigned int *arrivals, int n_leaf) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_leaf)
    return;
  int node = leaf_up[j] >> 2;
  for (;;) {
    __threadfence(); // the slot this thread (or k_update_leaves) wrote is visible before the arrival is counted
    const float4 *n = nodes + (size_t)node * RT_NODE_F4;
    const float4 cr = __ldcg(n + 6);
    const int n_children = (f2i(cr.x) != RT_EMPTY) + (f2i(cr.y) != RT_EMPTY) + (f2i(cr.z) != RT_EMPTY) + (f2i(cr.w) != RT_EMPTY);
    if (atomicAdd(&arrivals[node], 1u) + 1u < (unsigned int)n_children)
      return;
    __threadfence();
    int up = f2i(__ldcg(n + 7).x);
    if (up < 0)
      return;
    float4 rows[6];
    for (int r = 0; r < 6; r++)
      rows[r] = __ldcg(n + r);
    node_set_slot_box(nodes, up >> 2, up & 3, node_bounds(rows, cr));
    node = up >> 2;
  }
}

// ---------------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------------
static inline int ceil_div(long long a, int b) { return (int)((a + b - 1) / b); }

void launch_morton(cudaStream_t s, const BuildBox *boxes, int n, const float *scene_lo, const float *scene_inv,
                   uint64_t *codes, uint32_t *index) {
  k_morton<<<ceil_div(n, 256), 256, 0, s>>>(boxes, n, scene_lo, scene_inv, codes, index);
}

void launch_leaf_links(cudaStream_t s, const float4 *nodes, int n_nodes, int *leaf_up) {
  if (n_nodes > 0)
    k_leaf_links<<<ceil_div(n_nodes, 256), 256, 0, s>>>(nodes, n_nodes, leaf_up);
}
void launch_update_leaves(cudaStream_t s, const float4 *records, const PrimExact *exact, const BuildBox *boxes,
                          const int *leaf, int count, const int *leaf_up, float4 *prims, PrimExact *ex_prims, float4 *nodes) {
  if (count > 0)
    k_update_leaves<<<ceil_div(count, 128), 128, 0, s>>>(records, exact, boxes, leaf, count, leaf_up, prims, ex_prims, nodes);
}
void launch_refit_wide(cudaStream_t s, float4 *nodes, const int *leaf_up, unsigned int *arrivals, int n_leaf) {
  if (n_leaf > 0)
    k_refit_wide<<<ceil_div(n_leaf, 256), 256, 0, s>>>(nodes, leaf_up, arrivals, n_leaf);
}

int sort_pairs(cudaStream_t s, uint64_t *keys_in, uint64_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, int n) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, n, 0, 63, s);
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "cub::DeviceRadixSort::SortPairs (size query)");
  void *tmp = nullptr;
  e = cudaMalloc(&tmp, bytes ? bytes : 16);
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "cudaMalloc (sort scratch)");
  e = cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, n, 0, 63, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  cudaFree(tmp);
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
  if (e2 != cudaSuccess)
    return rt_cuda_fail(e2, "cudaStreamSynchronize (sort)");
  return RT_OK;
}

// All PLOC rounds: sorted leaf boxes -> BinTree (left / right / parent / box, root = node 0).  clusters[2] and
// nearest / packed are caller-provided scratch of n entries.  The host reads one 8-byte total per round.
int ploc_build(cudaStream_t s, const BuildBox *leaf_boxes, int n, BinTree t, PlocCluster *clusters[2], int *nearest,
               unsigned long long *packed, int *rounds_out) {
  size_t scan_bytes = 0;
  cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, packed, packed, n, s);
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "cub::DeviceScan::InclusiveSum (size query)");
  void *scan_tmp = nullptr;
  if ((e = cudaMalloc(&scan_tmp, scan_bytes ? scan_bytes : 16)) != cudaSuccess)
    return rt_cuda_fail(e, "cudaMalloc (scan scratch)");
  k_ploc_init<<<ceil_div(n, 256), 256, 0, s>>>(leaf_boxes, n, clusters[0]);
  int count = n, next_node = n - 2, cur = 0, rounds = 0, st = RT_OK;
  while (count > 1) {
    int blocks = ceil_div(count, 256);
    k_ploc_nearest<<<blocks, 256, 0, s>>>(clusters[cur], count, nearest);
    k_ploc_roles<<<blocks, 256, 0, s>>>(nearest, count, packed);
    e = cub::DeviceScan::InclusiveSum(scan_tmp, scan_bytes, packed, packed, count, s);
    if (e != cudaSuccess) {
      st = rt_cuda_fail(e, "cub::DeviceScan::InclusiveSum");
      break;
    }
    k_ploc_merge<<<blocks, 256, 0, s>>>(clusters[cur], nearest, packed, count, next_node, clusters[cur ^ 1], t);
    unsigned long long total = 0;
    if ((e = cudaMemcpyAsync(&total, packed + (count - 1), sizeof total, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) {
      st = rt_cuda_fail(e, "PLOC round");
      break;
    }
    int merges = (int)(total >> 32), survivors = (int)(total & 0xffffffffu);
    if (merges <= 0 || survivors != count - merges) { // cannot happen: the globally closest pair is always mutual
      rt_set_error("PLOC round made no progress");
      st = RT_ERR_CUDA;
      break;
    }
    next_node -= merges;
    count = survivors;
    cur ^= 1;
    rounds++;
  }
  cudaFree(scan_tmp);
  if (rounds_out)
    *rounds_out = rounds;
  return st;
}

int tree_area(cudaStream_t s, const BuildBox *box, int n_internal, double *d_sum, double *out) {
  cudaError_t e = cudaMemsetAsync(d_sum, 0, sizeof(double), s);
  if (e == cudaSuccess) {
    int blocks = ceil_div(n_internal, 256);
    k_tree_area<<<blocks < 1024 ? blocks : 1024, 256, 0, s>>>(box, n_internal, d_sum);
    e = cudaMemcpyAsync(out, d_sum, sizeof(double), cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(s);
  return e == cudaSuccess ? RT_OK : rt_cuda_fail(e, "tree_area");
}

void launch_gather_boxes(cudaStream_t s, const BuildBox *in, const uint32_t *index, BuildBox *out, int n) {
  k_gather_boxes<<<ceil_div(n, 256), 256, 0, s>>>(in, index, out, n);
}

void launch_gather_records(cudaStream_t s, const void *in, const uint32_t *index, void *out, int n, int bytes_per) {
  int vec_per = bytes_per / 16;
  long long total = (long long)n * vec_per;
  k_gather_records<<<ceil_div(total, 256), 256, 0, s>>>((const uint4 *)in, index, (uint4 *)out, n, vec_per);
}

void launch_hierarchy(cudaStream_t s, const uint64_t *codes, BinTree t) {
  if (t.n >= 2)
    k_hierarchy<<<ceil_div(t.n - 1, 256), 256, 0, s>>>(codes, t);
}

void launch_refit(cudaStream_t s, BinTree t, const BuildBox *leaf_boxes) {
  if (t.n >= 2)
    k_refit<<<ceil_div(t.n, 256), 256, 0, s>>>(t, leaf_boxes);
}

void launch_collapse(cudaStream_t s, BinTree t, const BuildBox *leaf_boxes, float4 *nodes, const CollapseItem *items,
                     int n_items, CollapseItem *next, int *next_count, int *wide_count) {
  k_collapse<<<ceil_div(n_items, 128), 128, 0, s>>>(t, leaf_boxes, nodes, items, n_items, next, next_count, wide_count);
}

void launch_generate(const rt_context *ctx, const PassParams &pp, WaveBuffers &w) {
  LaunchShape sh = rt_persistent_shape(ctx, RT_BLOCK, 16);
  int need = ceil_div(pp.n_paths, RT_BLOCK);
  k_generate<<<need < sh.blocks ? need : sh.blocks, RT_BLOCK, 0, ctx->pass_stream>>>(pp, w.ray_a[0], w.ray_b[0], w.counts);
}

void launch_extend(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int bounce, bool gen) {
  int b = bounce & 1;
  static const bool simple = getenv("RT_EXTEND") && std::string(getenv("RT_EXTEND")) == "simple";
  if (simple) {
    LaunchShape sh = rt_persistent_shape(ctx, RT_BLOCK, 8);
    int need = ceil_div(pp.n_paths, RT_BLOCK);
    k_extend_simple<<<need < sh.blocks ? need : sh.blocks, RT_BLOCK, 0, ctx->pass_stream>>>(
        sc, pp, w.ray_a[b], w.ray_b[b], w.hit[b], w.counts, bounce, sc.n_media > 0, w.stats);
    return;
  }
  // persistent warps: exactly the resident set (8 blocks of 4 warps per SM), each pulling rays from the queue
  LaunchShape sh = rt_persistent_shape(ctx, RT_BLOCK, 8);
  int need = ceil_div(pp.n_paths, RT_BLOCK);
  unsigned int *cursor = w.counts + (pp.max_depth + 2) + bounce;
  const int blocks = need < sh.blocks ? need : sh.blocks;
  static const bool mux = getenv("RT_EXTEND_MUX") && atoi(getenv("RT_EXTEND_MUX")) != 0;
  if (mux && !gen) {
    auto k = ctx->stats ? k_extend_mux<true> : k_extend_mux<false>;
    k<<<blocks, RT_BLOCK, 0, ctx->pass_stream>>>(sc, pp, w.ray_a[b], w.ray_b[b], w.hit[b], w.counts, cursor, bounce, sc.n_media > 0,
                                            w.stats);
    return;
  }
  const bool ties = rt_scene_tie_rule(sc); // the tie rule of leaf_test goes with k_tail's end-game sharing
  auto kernel = ties ? (gen ? (ctx->stats ? k_extend<true, true, true> : k_extend<false, true, true>)
                            : (ctx->stats ? k_extend<true, false, true> : k_extend<false, false, true>))
                     : (gen ? (ctx->stats ? k_extend<true, true, false> : k_extend<false, true, false>)
                            : (ctx->stats ? k_extend<true, false, false> : k_extend<false, false, false>));
  // RT_EXTEND_DYN_SMEM=<bytes>: unused dynamic shared memory per block (experiment aid: how much the kernel
  // depends on the L1 capacity that shared memory is carved out of)
  static const size_t dyn_smem = getenv("RT_EXTEND_DYN_SMEM") ? (size_t)atol(getenv("RT_EXTEND_DYN_SMEM")) : 0;
  rt_launch(kernel, blocks, RT_BLOCK, dyn_smem, ctx->pass_stream, true, sc, pp, (const float4 *)w.ray_a[b], (const float4 *)w.ray_b[b],
            w.hit[b], w.counts, cursor, bounce, (int)(sc.n_media > 0), w.stats);
}

void launch_shade(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int bounce, bool gen) {
  LaunchShape sh = rt_persistent_shape(ctx, RT_SHADE_THREADS, RT_SHADE_BLOCKS * RT_BLOCK / RT_SHADE_THREADS);
  int need = ceil_div(pp.n_paths, RT_SHADE_THREADS);
  int b = bounce & 1, nb = b ^ 1;
  auto kernel = gen ? k_shade<true> : k_shade<false>;
  unsigned int *cursor = w.counts + 2 * (pp.max_depth + 2) + bounce; // third block of the count words: shade fetch cursors
  rt_launch(kernel, need < sh.blocks ? need : sh.blocks, RT_SHADE_THREADS, 0, ctx->pass_stream, true, sc, pp,
            (const float4 *)w.ray_a[b], (const float4 *)w.ray_b[b], (const float2 *)w.hit[b], w.ray_a[nb], w.ray_b[nb], w.hit[nb],
            (const float4 *)w.thr[b], w.thr[nb], w.radiance, w.counts, cursor, bounce);
}

void launch_tail(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int first_bounce,
                 int end_bounce, int buffer) {
  LaunchShape sh = rt_persistent_shape(ctx, RT_BLOCK, RT_TAIL_BLOCKS);
  int need = ceil_div(pp.n_paths, RT_BLOCK);
  int b = buffer, nb = buffer ^ 1;
  unsigned int *cursor = w.counts + (pp.max_depth + 2) + first_bounce;
  const int blocks = need < sh.blocks ? need : sh.blocks;
  // RT_TAIL=regroup: the warp-pool variant (k_tail_regroup: bit-identical, measured slower) for A/B measurements
  static const bool regroup = getenv("RT_TAIL") && std::string(getenv("RT_TAIL")) == "regroup";
  if (regroup) {
    LaunchShape sh2 = rt_persistent_shape(ctx, RT_BLOCK, RT_TAIL2_BLOCKS);
    const int blocks2 = need < sh2.blocks ? need : sh2.blocks;
    auto k = ctx->stats ? k_tail_regroup<true> : k_tail_regroup<false>;
    k<<<blocks2, RT_BLOCK, 0, ctx->pass_stream>>>(sc, pp, w.ray_a[b], w.ray_b[b], w.hit[b], w.ray_a[nb], w.ray_b[nb], w.hit[nb],
                                             w.thr[b], w.thr[nb], w.radiance, w.counts, cursor, first_bounce, end_bounce,
                                             sc.n_media > 0, w.stats);
    return;
  }
  // End-game sharing pays where single traversals get long enough to hold a launch up - large scenes (10^6 spheres:
  // tail 1.54 -> 1.27 ms) - and costs 2-4 % of the kernel elsewhere (its bookkeeping in a 96-register kernel).
  // RT_TAIL_SHARE=0 / 1 forces it off / on.
  const bool share = rt_scene_shares_traversals(sc), ties = rt_scene_tie_rule(sc);
  auto kernel = ctx->stats ? (share ? k_tail<true, true, true> : (ties ? k_tail<true, false, true> : k_tail<true, false, false>))
                           : (share ? k_tail<false, true, true> : (ties ? k_tail<false, false, true> : k_tail<false, false, false>));
  rt_launch(kernel, blocks, RT_BLOCK, 0, ctx->pass_stream, true, sc, pp, w.ray_a[b], w.ray_b[b], w.hit[b], w.ray_a[nb], w.ray_b[nb],
            w.hit[nb], w.thr[b], w.thr[nb], w.radiance, w.counts, cursor, first_bounce, end_bounce, (int)(sc.n_media > 0), w.stats);
}

void launch_accumulate(const rt_context *ctx, const PassParams &pp, WaveBuffers &w, float4 *film) {
  LaunchShape sh = rt_persistent_shape(ctx, 256, 8);
  int need = ceil_div(pp.n_owned, 256);
  k_accumulate<<<need < sh.blocks ? need : sh.blocks, 256, 0, ctx->pass_stream>>>(pp, w.radiance, film);
}

void launch_resolve_rgb8(cudaStream_t s, const float4 *film, int64_t n, double scale, uint8_t *out) {
  if (n > 0)
    k_resolve_rgb8<<<ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096, 256, 0, s>>>(film, n, scale, out);
}

void launch_resolve_rgb(cudaStream_t s, const float4 *film, int64_t n, double scale, float *out) {
  if (n > 0)
    k_resolve_rgb<<<ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096, 256, 0, s>>>(film, n, scale, out);
}

void launch_scatter_gathered(cudaStream_t s, int width, int height, int n_ranks, int tile_rows, const void *gathered,
                             void *full, int bytes_per_pixel) {
  long long total = (long long)width * height;
  if (total <= 0)
    return;
  RankBases bases;
  long long first = 0;
  for (int r = 0; r < n_ranks; r++) {
    bases.first_pixel[r] = first;
    first += (long long)owned_rows(height, r, n_ranks, tile_rows) * width;
  }
  int blocks = ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096;
  if (bytes_per_pixel == 16)
    k_scatter_gathered<float4><<<blocks, 256, 0, s>>>(width, height, n_ranks, tile_rows, bases, (const float4 *)gathered,
                                                       (float4 *)full);
  else
    k_scatter_gathered<Rgb8><<<blocks, 256, 0, s>>>(width, height, n_ranks, tile_rows, bases, (const Rgb8 *)gathered,
                                                     (Rgb8 *)full);
}

void launch_trace_fast(const rt_context *ctx, const DScene &sc, const rt_ray *d_rays, int64_t n, uint64_t seed,
                       const int *leaf_object, const int *leaf_id, rt_hit *d_hits) {
  if (n <= 0)
    return;
  LaunchShape sh = rt_persistent_shape(ctx, RT_BLOCK, 8);
  int need = ceil_div(n, RT_BLOCK);
  auto kernel = rt_scene_tie_rule(sc) ? k_trace_fast<true> : k_trace_fast<false>;
  kernel<<<need < sh.blocks ? need : sh.blocks, RT_BLOCK, 0, ctx->stream>>>(sc, d_rays, n, seed, leaf_object, leaf_id, d_hits);
}
