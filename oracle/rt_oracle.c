/*
 * oracle/rt_oracle.c — TEST INFRASTRUCTURE ONLY (see rt_oracle.h).
 *
 * Plain-C, FP64, scalar restatement of the reference's CPU path-tracing path.  Every function names
 * the reference source it follows (paths relative to the reference's src/).  The operation order of
 * the reference's *scalar fallback* branches is kept so that, compiled with -ffp-contract=off, the
 * mt19937 mode reproduces the reference bit for bit (checked by tests/test_oracle_vs_ref.py against
 * oracle/_ref/libref_harness.so and by the committed fixtures in tests/golden/).
 *
 * Two random sources:
 *   mt19937 + rejection samplers  : what the reference CPU does (Utility.hpp:16-37,
 *                                   Vec3Utility.hpp:41-62)
 *   Philox4x32-10 + polar samplers: the stream the CUDA kernels draw from, keyed by
 *                                   (seed; pixel, sample, bounce, stream) with a fixed slot layout,
 *                                   using the reference CUDA path's polar samplers
 *                                   (Vec3Utility.cuh:57-70) and its pdf guard (CameraKernels.cu:192)
 */
#include "rt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORA_PI 3.1415926535897932385 /* Utility.hpp:8 */
#define ORA_INF INFINITY

/* ------------------------------------------------------------------------------------------------
 * Vec3 (Vec3.hpp scalar branches, Vec3Utility.hpp:6-38)
 * ---------------------------------------------------------------------------------------------- */
typedef struct v3 {
  double x, y, z;
} v3;

static inline v3 V(double x, double y, double z) {
  v3 r = {x, y, z};
  return r;
}
static inline v3 vfrom(const double *p) { return V(p[0], p[1], p[2]); }
static inline void vto(double *p, v3 a) {
  p[0] = a.x;
  p[1] = a.y;
  p[2] = a.z;
}
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(double t, v3 a) { return V(t * a.x, t * a.y, t * a.z); }
static inline v3 vdiv(v3 a, double t) { return vscale(1 / t, a); } /* Vec3Utility.hpp:24 */
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double vlen2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double vlen(v3 a) { return sqrt(vlen2(a)); }
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* Vec3::normalize (Vec3.hpp:141-149) */
static inline v3 vunit(v3 a) {
  double len = vlen(a);
  if (len > 1e-8) {
    double s = 1.0 / len;
    return V(a.x * s, a.y * s, a.z * s);
  }
  return V(1.0, 0.0, 0.0);
}
static inline double vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* ------------------------------------------------------------------------------------------------
 * Random sources
 * ---------------------------------------------------------------------------------------------- */
void ora_mt_seed(ora_mt19937 *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

uint32_t ora_mt_next(ora_mt19937 *g) {
  if (g->idx >= 624) {
    uint32_t *mt = g->mt;
    for (int k = 0; k < 624; k++) {
      uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* std::generate_canonical<double,53>(mt19937) as libstdc++ implements it: two engine calls,
 * (g1 + g2 * 2^32) / 2^64, clamped below 1.  This is what uniform_real_distribution<double>(0,1)
 * returns (Utility.hpp:22-25). */
double ora_mt_canonical(ora_mt19937 *g) {
  double sum = 0.0, tmp = 1.0;
  sum += (double)ora_mt_next(g) * tmp;
  tmp *= 4294967296.0;
  sum += (double)ora_mt_next(g) * tmp;
  tmp *= 4294967296.0;
  double ret = sum / tmp;
  if (ret >= 1.0)
    ret = nextafter(1.0, 0.0);
  return ret;
}

/* std::uniform_int_distribution<int>(lo,hi)(mt19937) as libstdc++ (GCC >= 11) implements it for a
 * 32-bit engine: Lemire's nearly-divisionless method (Utility.hpp:34-37). */
int ora_mt_uniform_int(ora_mt19937 *g, int lo, int hi) {
  uint32_t urange = (uint32_t)hi - (uint32_t)lo;
  if (urange == 0xffffffffu)
    return (int)((uint32_t)lo + ora_mt_next(g));
  uint32_t range = urange + 1u;
  uint64_t product = (uint64_t)ora_mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)ora_mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (int)((uint32_t)lo + (uint32_t)(product >> 32));
}

/* Philox4x32-10 (Salmon et al., SC'11; the constants every implementation shares). */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

typedef struct rng_t {
  int kind;    /* ORA_RNG_* */
  int sampler; /* ORA_SAMPLER_* */
  ora_mt19937 *mt;
  uint64_t seed;
  uint32_t pixel, sample;
  uint64_t draws;
} rng_t;

/* random_double() (Utility.hpp:22-25): canonical * (1 - 0) + 0. */
static double rnd(rng_t *r) {
  r->draws++;
  return ora_mt_canonical(r->mt) * (1.0 - 0.0) + 0.0;
}
/* random_double(min,max) (Utility.hpp:28-31): canonical * (max - min) + min. */
static double rnd_range(rng_t *r, double lo, double hi) {
  r->draws++;
  return ora_mt_canonical(r->mt) * (hi - lo) + lo;
}
static int rnd_int(rng_t *r, int lo, int hi) {
  r->draws++;
  return ora_mt_uniform_int(r->mt, lo, hi);
}

/* Four 24-bit uniforms in [0,1) of block `block` of stream `stream` at bounce `bounce`. */
static void philox_uniforms(rng_t *r, uint32_t bounce, uint32_t stream, uint32_t block, double out[4]) {
  uint32_t ctr[4] = {r->pixel, r->sample, bounce, (stream << 16) | block};
  uint32_t key[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
  uint32_t o[4];
  ora_philox4x32_10(ctr, key, o);
  for (int i = 0; i < 4; i++)
    out[i] = (double)(o[i] >> 8) * (1.0 / 16777216.0);
  r->draws += 4;
}

/* ------------------------------------------------------------------------------------------------
 * Scene
 * ---------------------------------------------------------------------------------------------- */
typedef struct box_t {
  double lo[3], hi[3];
} box_t;

typedef struct osphere {
  v3 c0, cdir;
  double radius;
  int material, id;
} osphere;

typedef struct oquad {
  v3 corner, u, v, normal, w;
  double D, area;
  int material, id;
} oquad;

enum { OBJ_SURFACE = 0, OBJ_MEDIUM = 1 };

typedef struct oobject {
  int kind;
  int shape;     /* RT_SHAPE_* of the member primitives */
  int first, n;  /* member primitives in the spheres / quads array */
  int xform;     /* -1 = none */
  int medium;    /* index into media when kind == OBJ_MEDIUM */
  int closed;    /* 1: members use Interval::contains on t (quads), 0: surrounds (spheres, media) */
  box_t box;
} oobject;

typedef struct onode {
  box_t box;
  int left, right; /* children: >= 0 node index, < 0: ~object index */
} onode;

struct ora_scene {
  rt_scene_desc d; /* deep copies of the arrays */
  osphere *spheres;
  oquad *quads;
  oobject *objects;
  int n_objects;
  onode *nodes;
  int n_nodes, root;
  ora_counters cnt;
};

typedef struct oray {
  v3 o, d;
  double time;
} oray;

typedef struct ohit {
  v3 p, normal;
  double t, u, v;
  int front, material, prim, object;
} ohit;

static void *dup_array(const void *src, size_t n, size_t sz) {
  if (!n)
    return NULL;
  void *p = malloc(n * sz);
  memcpy(p, src, n * sz);
  return p;
}

/* AABB(Point3,Point3) + pad_to_minimums (AABB.cpp:12-21,167-176). */
static void box_pad(box_t *b) {
  const double delta = 0.0001;
  for (int a = 0; a < 3; a++)
    if (b->hi[a] - b->lo[a] < delta) {
      double padding = delta * 0.5;
      b->lo[a] = b->lo[a] - padding;
      b->hi[a] = b->hi[a] + padding;
    }
}
static box_t box_points(v3 p1, v3 p2) {
  box_t b;
  for (int a = 0; a < 3; a++) {
    double x = vget(p1, a), y = vget(p2, a);
    b.lo[a] = x <= y ? x : y;
    b.hi[a] = x <= y ? y : x;
  }
  box_pad(&b);
  return b;
}
/* AABB(AABB,AABB) (AABB.cpp:22-26, Interval.hpp:36-37). */
static box_t box_union(box_t a, box_t b) {
  box_t r;
  for (int k = 0; k < 3; k++) {
    r.lo[k] = a.lo[k] <= b.lo[k] ? a.lo[k] : b.lo[k];
    r.hi[k] = a.hi[k] >= b.hi[k] ? a.hi[k] : b.hi[k];
  }
  return r;
}
static box_t box_empty(void) {
  box_t b;
  for (int k = 0; k < 3; k++) {
    b.lo[k] = ORA_INF;
    b.hi[k] = -ORA_INF;
  }
  return b;
}

/* Sphere bounding box (Sphere.cpp:8-23). */
static box_t sphere_box(const osphere *s) {
  v3 rv = V(s->radius, s->radius, s->radius);
  v3 c1 = vadd(s->c0, vscale(1.0, s->cdir));
  box_t b0 = box_points(vsub(s->c0, rv), vadd(s->c0, rv));
  box_t b1 = box_points(vsub(c1, rv), vadd(c1, rv));
  return box_union(b0, b1);
}
/* Plane bounding box (Plane.cpp:17-20). */
static box_t quad_box(const oquad *q) {
  box_t d1 = box_points(q->corner, vadd(vadd(q->corner, q->u), q->v));
  box_t d2 = box_points(vadd(q->corner, q->u), vadd(q->corner, q->v));
  return box_union(d1, d2);
}
/* RotateY / Translate bounding boxes (RotateY.cpp:5-37, Translate.cpp:7-10), innermost wrapper first. */
static box_t xform_box(const ora_scene *s, int xf, box_t b) {
  if (xf < 0)
    return b;
  const rt_xform *x = &s->d.xforms[xf];
  for (int k = x->n_ops - 1; k >= 0; k--) {
    const rt_xform_op *op = &s->d.xform_ops[x->first_op + k];
    if (op->type == RT_XF_TRANSLATE) {
      for (int a = 0; a < 3; a++) {
        b.lo[a] += op->offset[a];
        b.hi[a] += op->offset[a];
      }
      box_pad(&b);
    } else {
      double mn[3] = {ORA_INF, ORA_INF, ORA_INF}, mx[3] = {-ORA_INF, -ORA_INF, -ORA_INF};
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          for (int kk = 0; kk < 2; kk++) {
            double x0 = i * b.hi[0] + (1 - i) * b.lo[0];
            double y0 = j * b.hi[1] + (1 - j) * b.lo[1];
            double z0 = kk * b.hi[2] + (1 - kk) * b.lo[2];
            double nx = op->cos_theta * x0 + op->sin_theta * z0;
            double nz = -op->sin_theta * x0 + op->cos_theta * z0;
            double t[3] = {nx, y0, nz};
            for (int c = 0; c < 3; c++) {
              mn[c] = fmin(mn[c], t[c]);
              mx[c] = fmax(mx[c], t[c]);
            }
          }
      b = box_points(V(mn[0], mn[1], mn[2]), V(mx[0], mx[1], mx[2]));
    }
  }
  return b;
}

static int cmp_axis;
static const ora_scene *cmp_scene;
static int cmp_objects(const void *a, const void *b) {
  const box_t *ba = &cmp_scene->objects[*(const int *)a].box, *bb = &cmp_scene->objects[*(const int *)b].box;
  double ca = (ba->lo[cmp_axis] + ba->hi[cmp_axis]) * 0.5, cb = (bb->lo[cmp_axis] + bb->hi[cmp_axis]) * 0.5;
  if (ca < cb)
    return -1;
  if (ca > cb)
    return 1;
  return *(const int *)a - *(const int *)b;
}

/* Median split on the longest axis (the tree shape does not matter: the closest hit is
 * tree-independent; ties are resolved with the list's rule, see bvh_hit). */
static int build_nodes(ora_scene *s, int *idx, int lo, int hi) {
  if (hi - lo == 1)
    return ~idx[lo];
  box_t b = box_empty();
  for (int i = lo; i < hi; i++)
    b = box_union(b, s->objects[idx[i]].box);
  int axis = 0;
  double ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
  if (ex > ey)
    axis = ex > ez ? 0 : 2;
  else
    axis = ey > ez ? 1 : 2;
  cmp_axis = axis;
  cmp_scene = s;
  qsort(idx + lo, (size_t)(hi - lo), sizeof(int), cmp_objects);
  int mid = lo + (hi - lo) / 2;
  int me = s->n_nodes++;
  s->nodes[me].box = b;
  int l = build_nodes(s, idx, lo, mid);
  int r = build_nodes(s, idx, mid, hi);
  s->nodes[me].left = l;
  s->nodes[me].right = r;
  return me;
}

ora_scene *ora_scene_create(const rt_scene_desc *desc) {
  ora_scene *s = (ora_scene *)calloc(1, sizeof *s);
  s->d = *desc;
  s->d.spheres = (const rt_sphere *)dup_array(desc->spheres, (size_t)desc->n_spheres, sizeof(rt_sphere));
  s->d.quads = (const rt_quad *)dup_array(desc->quads, (size_t)desc->n_quads, sizeof(rt_quad));
  s->d.xform_ops = (const rt_xform_op *)dup_array(desc->xform_ops, (size_t)desc->n_xform_ops, sizeof(rt_xform_op));
  s->d.xforms = (const rt_xform *)dup_array(desc->xforms, (size_t)desc->n_xforms, sizeof(rt_xform));
  s->d.media = (const rt_medium *)dup_array(desc->media, (size_t)desc->n_media, sizeof(rt_medium));
  s->d.materials = (const rt_material *)dup_array(desc->materials, (size_t)desc->n_materials, sizeof(rt_material));
  s->d.textures = (const rt_texture *)dup_array(desc->textures, (size_t)desc->n_textures, sizeof(rt_texture));
  s->d.perlins = (const rt_perlin *)dup_array(desc->perlins, (size_t)desc->n_perlins, sizeof(rt_perlin));
  s->d.lights = (const rt_light *)dup_array(desc->lights, (size_t)desc->n_lights, sizeof(rt_light));
  rt_image *images = (rt_image *)dup_array(desc->images, (size_t)desc->n_images, sizeof(rt_image));
  for (int i = 0; i < desc->n_images; i++)
    images[i].rgb = (const uint8_t *)dup_array(desc->images[i].rgb,
                                               (size_t)desc->images[i].width * (size_t)desc->images[i].height, 3);
  s->d.images = images;

  int ns = desc->n_spheres, nq = desc->n_quads;
  s->spheres = (osphere *)calloc((size_t)(ns ? ns : 1), sizeof(osphere));
  s->quads = (oquad *)calloc((size_t)(nq ? nq : 1), sizeof(oquad));
  for (int i = 0; i < ns; i++) {
    const rt_sphere *p = &desc->spheres[i];
    osphere *o = &s->spheres[i];
    o->c0 = vfrom(p->center0);
    o->cdir = vfrom(p->center_dir);
    o->radius = fmax(0, p->radius); /* Sphere.cpp:9 */
    o->material = p->material;
    o->id = i;
  }
  for (int i = 0; i < nq; i++) { /* Plane ctor, Plane.cpp:6-21 */
    const rt_quad *p = &desc->quads[i];
    oquad *o = &s->quads[i];
    o->corner = vfrom(p->corner);
    o->u = vfrom(p->u);
    o->v = vfrom(p->v);
    v3 n = vcross(o->u, o->v);
    o->normal = vunit(n);
    o->D = vdot(o->normal, o->corner);
    o->w = vdiv(n, vdot(n, n));
    o->area = vlen(n);
    o->material = p->material;
    o->id = ns + i;
  }

  /* Top-level objects: primitives sharing an `object` id are one HittableList (a make_box) under
   * one wrapper chain; a medium's boundary primitives belong to the medium's object. */
  s->n_objects = desc->n_objects;
  s->objects = (oobject *)calloc((size_t)(s->n_objects ? s->n_objects : 1), sizeof(oobject));
  for (int k = 0; k < s->n_objects; k++) {
    s->objects[k].first = -1;
    s->objects[k].medium = -1;
    s->objects[k].xform = -1;
    s->objects[k].box = box_empty();
  }
  for (int i = 0; i < ns; i++) {
    oobject *o = &s->objects[desc->spheres[i].object];
    if (o->first < 0) {
      o->first = i;
      o->shape = RT_SHAPE_SPHERE;
      o->xform = desc->spheres[i].xform;
      o->closed = 0;
    }
    o->n++;
    o->box = box_union(o->box, sphere_box(&s->spheres[i]));
  }
  for (int i = 0; i < nq; i++) {
    oobject *o = &s->objects[desc->quads[i].object];
    if (o->first < 0) {
      o->first = i;
      o->shape = RT_SHAPE_QUAD;
      o->xform = desc->quads[i].xform;
      o->closed = 1;
    }
    o->n++;
    o->box = box_union(o->box, quad_box(&s->quads[i]));
  }
  for (int m = 0; m < desc->n_media; m++) {
    oobject *o = &s->objects[desc->media[m].object];
    o->kind = OBJ_MEDIUM;
    o->medium = m;
    o->closed = 0;
  }
  for (int k = 0; k < s->n_objects; k++)
    s->objects[k].box = xform_box(s, s->objects[k].xform, s->objects[k].box);

  if (s->n_objects > 0) {
    int *idx = (int *)malloc(sizeof(int) * (size_t)s->n_objects);
    for (int k = 0; k < s->n_objects; k++)
      idx[k] = k;
    s->nodes = (onode *)calloc((size_t)s->n_objects, sizeof(onode));
    s->n_nodes = 0;
    s->root = build_nodes(s, idx, 0, s->n_objects);
    free(idx);
  }
  return s;
}

void ora_scene_destroy(ora_scene *s) {
  if (!s)
    return;
  free((void *)s->d.spheres);
  free((void *)s->d.quads);
  free((void *)s->d.xform_ops);
  free((void *)s->d.xforms);
  free((void *)s->d.media);
  free((void *)s->d.materials);
  free((void *)s->d.textures);
  free((void *)s->d.perlins);
  free((void *)s->d.lights);
  for (int i = 0; i < s->d.n_images; i++)
    free((void *)s->d.images[i].rgb);
  free((void *)s->d.images);
  free(s->spheres);
  free(s->quads);
  free(s->objects);
  free(s->nodes);
  free(s);
}

/* ------------------------------------------------------------------------------------------------
 * Primitive intersection
 * ---------------------------------------------------------------------------------------------- */
static inline v3 ray_at(const oray *r, double t) { /* Ray.hpp:48-51 */
  return V(r->o.x + t * r->d.x, r->o.y + t * r->d.y, r->o.z + t * r->d.z);
}
/* HitRecord::set_face_normal (HitRecord.hpp:36-39). */
static inline void set_face_normal(ohit *h, const oray *r, v3 outward) {
  h->front = vdot(r->d, outward) < 0;
  h->normal = h->front ? outward : vneg(outward);
}

/* Sphere::hit (Sphere.cpp:101-143).  closed_max = 0 is the reference's Interval::surrounds; 1 also
 * accepts root == t_max (used only by the BVH walker, which then applies the list's tie rule). */
static int sphere_hit(ora_scene *s, const osphere *sp, const oray *r, double tmin, double tmax, int closed_max,
                      ohit *h) {
  s->cnt.sphere_tests++;
  v3 center = V(sp->c0.x + r->time * sp->cdir.x, sp->c0.y + r->time * sp->cdir.y, sp->c0.z + r->time * sp->cdir.z);
  v3 oc = vsub(center, r->o);
  double a = vlen2(r->d);
  double hh = vdot(r->d, oc);
  double c = vlen2(oc) - sp->radius * sp->radius;
  double disc = hh * hh - a * c;
  if (disc < 0)
    return 0;
  double sqrtd = sqrt(disc);
  double root = (hh - sqrtd) / a;
  if (!(tmin < root && (closed_max ? root <= tmax : root < tmax))) {
    root = (hh + sqrtd) / a;
    if (!(tmin < root && (closed_max ? root <= tmax : root < tmax)))
      return 0;
  }
  h->t = root;
  h->p = ray_at(r, root);
  v3 outward = vdiv(vsub(h->p, center), sp->radius);
  set_face_normal(h, r, outward);
  h->material = sp->material;
  double theta = acos(-outward.y);
  double phi = atan2(-outward.z, outward.x) + ORA_PI;
  h->u = phi / (2 * ORA_PI);
  h->v = theta / ORA_PI;
  h->prim = sp->id;
  return 1;
}

void ora_sphere_uv(const double n[3], double *u, double *v) {
  double theta = acos(-n[1]);
  double phi = atan2(-n[2], n[0]) + ORA_PI;
  *u = phi / (2 * ORA_PI);
  *v = theta / ORA_PI;
}

/* Plane::hit (Plane.cpp:78-112). */
static int quad_hit(ora_scene *s, const oquad *q, const oray *r, double tmin, double tmax, ohit *h) {
  s->cnt.quad_tests++;
  double denom = vdot(q->normal, r->d);
  if (fabs(denom) < 1e-8)
    return 0;
  double t = (q->D - vdot(q->normal, r->o)) / denom;
  if (!(tmin <= t && t <= tmax))
    return 0;
  v3 p = ray_at(r, t);
  v3 hp = vsub(p, q->corner);
  double alpha = vdot(q->w, vcross(hp, q->v));
  double beta = vdot(q->w, vcross(q->u, hp));
  if (!(0 <= alpha && alpha <= 1) || !(0 <= beta && beta <= 1))
    return 0;
  h->u = alpha;
  h->v = beta;
  h->t = t;
  h->p = p;
  h->material = q->material;
  set_face_normal(h, r, q->normal);
  h->prim = q->id;
  return 1;
}

/* The member list of one object under its wrapper chain: Translate::hit (Translate.cpp:17-28),
 * RotateY::hit (RotateY.cpp:41-79), HittableList::hit (HittableList.cpp:26-42). */
static int members_hit(ora_scene *s, const oobject *o, const oray *ray, double tmin, double tmax, int closed_max,
                       ohit *h) {
  oray r = *ray;
  const rt_xform *x = o->xform >= 0 ? &s->d.xforms[o->xform] : NULL;
  int nops = x ? x->n_ops : 0;
  for (int k = 0; k < nops; k++) {
    const rt_xform_op *op = &s->d.xform_ops[x->first_op + k];
    if (op->type == RT_XF_TRANSLATE) {
      r.o = vsub(r.o, vfrom(op->offset));
    } else {
      double c = op->cos_theta, sn = op->sin_theta;
      v3 o2 = V((c * r.o.x) - (sn * r.o.z), r.o.y, (sn * r.o.x) + (c * r.o.z));
      v3 d2 = V((c * r.d.x) - (sn * r.d.z), r.d.y, (sn * r.d.x) + (c * r.d.z));
      r.o = o2;
      r.d = d2;
    }
  }
  int any = 0;
  double closest = tmax;
  ohit tmp;
  for (int i = 0; i < o->n; i++) {
    int ok;
    if (o->shape == RT_SHAPE_SPHERE)
      ok = sphere_hit(s, &s->spheres[o->first + i], &r, tmin, closest, closed_max, &tmp);
    else
      ok = quad_hit(s, &s->quads[o->first + i], &r, tmin, closest, &tmp);
    if (ok) {
      any = 1;
      closest = tmp.t;
      *h = tmp;
    }
  }
  if (!any)
    return 0;
  for (int k = nops - 1; k >= 0; k--) {
    const rt_xform_op *op = &s->d.xform_ops[x->first_op + k];
    if (op->type == RT_XF_TRANSLATE) {
      h->p = vadd(h->p, vfrom(op->offset));
    } else {
      double c = op->cos_theta, sn = op->sin_theta;
      h->p = V((c * h->p.x) + (sn * h->p.z), h->p.y, (-sn * h->p.x) + (c * h->p.z));
      h->normal = V((c * h->normal.x) + (sn * h->normal.z), h->normal.y, (-sn * h->normal.x) + (c * h->normal.z));
    }
  }
  return 1;
}

/* ConstantMedium::hit (ConstantMedium.cpp:25-94, scalar branch :71-84). */
static int medium_hit(ora_scene *s, const oobject *o, const oray *r, double tmin, double tmax, rng_t *rng,
                      uint32_t bounce, ohit *h) {
  const rt_medium *m = &s->d.media[o->medium];
  ohit rec1, rec2;
  if (!members_hit(s, o, r, -ORA_INF, ORA_INF, 0, &rec1))
    return 0;
  if (!members_hit(s, o, r, rec1.t + 0.0001, ORA_INF, 0, &rec2))
    return 0;
  if (rec1.t < tmin)
    rec1.t = tmin;
  if (rec2.t > tmax)
    rec2.t = tmax;
  if (rec1.t >= rec2.t)
    return 0;
  if (rec1.t < 0)
    rec1.t = 0;
  double ray_length = vlen(r->d);
  double distance_inside = (rec2.t - rec1.t) * ray_length;
  double neg_inv_density = -1.0 / m->density;
  double xi;
  if (rng->kind == ORA_RNG_MT19937) {
    xi = rnd(rng);
  } else {
    double u[4];
    philox_uniforms(rng, bounce, (uint32_t)(ORA_STREAM_MEDIUM0 + o->medium), 0, u);
    xi = u[0];
  }
  double hit_distance = neg_inv_density * log(xi);
  if (hit_distance > distance_inside)
    return 0;
  h->t = rec1.t + hit_distance / ray_length;
  h->p = ray_at(r, h->t);
  h->normal = V(1, 0, 0);
  h->front = 1;
  h->material = m->material;
  h->u = rec2.u; /* the reference leaves u,v as whatever the record held; unused by the phase function */
  h->v = rec2.v;
  h->prim = s->d.n_spheres + s->d.n_quads + o->medium;
  return 1;
}

static int object_hit(ora_scene *s, int k, const oray *r, double tmin, double tmax, int closed_max, rng_t *rng,
                      uint32_t bounce, ohit *h) {
  const oobject *o = &s->objects[k];
  int ok = o->kind == OBJ_MEDIUM ? medium_hit(s, o, r, tmin, tmax, rng, bounce, h)
                                 : members_hit(s, o, r, tmin, tmax, closed_max, h);
  if (ok)
    h->object = k;
  return ok;
}

/* HittableList::hit over the top-level objects (HittableList.cpp:26-42). */
static int list_hit(ora_scene *s, const oray *r, double tmin, double tmax, rng_t *rng, uint32_t bounce, ohit *h) {
  int any = 0;
  double closest = tmax;
  ohit tmp;
  for (int k = 0; k < s->n_objects; k++)
    if (object_hit(s, k, r, tmin, closest, 0, rng, bounce, &tmp)) {
      any = 1;
      closest = tmp.t;
      *h = tmp;
    }
  return any;
}

/* AABB::hit (AABB.cpp:141-164). */
static int box_hit(ora_scene *s, const box_t *b, const oray *r, double tmin, double tmax) {
  s->cnt.node_tests++;
  for (int axis = 0; axis < 3; axis++) {
    double dir_inv = 1.0 / vget(r->d, axis);
    double org = vget(r->o, axis);
    double t0 = (b->lo[axis] - org) * dir_inv;
    double t1 = (b->hi[axis] - org) * dir_inv;
    if (t0 < t1) {
      if (t0 > tmin)
        tmin = t0;
      if (t1 < tmax)
        tmax = t1;
    } else {
      if (t1 > tmin)
        tmin = t1;
      if (t0 < tmax)
        tmax = t0;
    }
    if (tmax <= tmin)
      return 0;
  }
  return 1;
}

/* Pointer-tree traversal in the shape of BVHNode::hit (BVHNode.cpp:133-146).  Candidates are tested
 * on the closed interval and exact-t ties are then resolved the way HittableList's in-order scan
 * resolves them: of two objects hitting at the same t the later one wins iff its members use the
 * closed Interval::contains test (quads, Plane.cpp:86); spheres and media use the open test
 * (Sphere.cpp:116), so a later one loses. */
static void bvh_walk(ora_scene *s, int node, const oray *r, double tmin, double tmax, rng_t *rng, uint32_t bounce,
                     ohit *best, int *have) {
  if (node < 0) {
    int k = ~node;
    ohit tmp;
    double limit = *have ? best->t : tmax;
    if (!object_hit(s, k, r, tmin, limit, *have ? 1 : 0, rng, bounce, &tmp))
      return;
    if (*have && tmp.t == best->t) {
      int later = tmp.object > best->object;
      int wins = later ? s->objects[tmp.object].closed : !s->objects[best->object].closed;
      if (!wins)
        return;
    }
    *best = tmp;
    *have = 1;
    return;
  }
  if (!box_hit(s, &s->nodes[node].box, r, tmin, *have ? best->t : tmax))
    return;
  bvh_walk(s, s->nodes[node].left, r, tmin, tmax, rng, bounce, best, have);
  bvh_walk(s, s->nodes[node].right, r, tmin, tmax, rng, bounce, best, have);
}

/* Optional log of every segment ray_color traces (test aid: realistic secondary-ray sets). */
static rt_ray *g_ray_log = NULL;
static int64_t g_ray_log_cap = 0, g_ray_log_n = 0;
void ora_ray_log(rt_ray *buf, int64_t cap) {
  g_ray_log = buf;
  g_ray_log_cap = cap;
  g_ray_log_n = 0;
}
int64_t ora_ray_log_count(void) { return g_ray_log_n; }

static int world_hit(ora_scene *s, int use_bvh, const oray *r, double tmin, double tmax, rng_t *rng, uint32_t bounce,
                     ohit *h) {
  if (g_ray_log && g_ray_log_n < g_ray_log_cap) {
    rt_ray *q = &g_ray_log[g_ray_log_n++];
    memset(q, 0, sizeof *q);
    vto(q->origin, r->o);
    vto(q->direction, r->d);
    q->time = r->time;
    q->t_min = tmin;
    q->t_max = tmax;
    q->rng_pixel = rng->pixel;
    q->rng_sample = rng->sample;
    q->rng_bounce = bounce;
  }
  if (!use_bvh || s->n_objects == 0)
    return list_hit(s, r, tmin, tmax, rng, bounce, h);
  int have = 0;
  bvh_walk(s, s->root, r, tmin, tmax, rng, bounce, h, &have);
  return have;
}

/* ------------------------------------------------------------------------------------------------
 * Textures (SolidColorTexture.cpp:8-10, CheckerTexture.cpp:43-54, NoiseTexture.cpp:29-30,
 * PerlinNoise.hpp:43-79,186-201)
 * ---------------------------------------------------------------------------------------------- */
static double perlin_noise(const rt_perlin *pn, v3 p) {
  double uf = p.x - floor(p.x), vf = p.y - floor(p.y), wf = p.z - floor(p.z);
  int xi = (int)floor(p.x), yi = (int)floor(p.y), zi = (int)floor(p.z);
  double uu = uf * uf * (3 - 2 * uf), vv = vf * vf * (3 - 2 * vf), ww = wf * wf * (3 - 2 * wf);
  double accum = 0.0;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        const double *g =
            pn->rand_vec[pn->perm_x[(xi + i) & 255] ^ pn->perm_y[(yi + j) & 255] ^ pn->perm_z[(zi + k) & 255]];
        v3 wv = V(uf - i, vf - j, wf - k);
        accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                 vdot(vfrom(g), wv);
      }
  return accum;
}
static double perlin_turb(const rt_perlin *pn, v3 p, int depth) {
  double accum = 0.0, weight = 1.0;
  v3 tp = p;
  for (int i = 0; i < depth; i++) {
    accum += weight * perlin_noise(pn, tp);
    weight *= 0.5;
    tp = V(tp.x * 2, tp.y * 2, tp.z * 2);
  }
  return fabs(accum);
}
static v3 texture_value(const ora_scene *s, int tex, double u, double v, v3 p) {
  const rt_texture *t = &s->d.textures[tex];
  if (t->type == RT_TEX_SOLID)
    return vfrom(t->color);
  if (t->type == RT_TEX_CHECKER) {
    double inv_scale = 1.0 / t->scale;
    int xi = (int)floor(inv_scale * p.x), yi = (int)floor(inv_scale * p.y), zi = (int)floor(inv_scale * p.z);
    int even = (xi + yi + zi) % 2 == 0;
    return texture_value(s, even ? t->even : t->odd, u, v, p);
  }
  if (t->type == RT_TEX_IMAGE) {
    /* not in the reference: "Ray Tracing: The Next Week" image_texture::value over the (u, v) the
     * reference's primitives compute (Sphere.cpp:136-140, Plane.cpp:101-102) */
    const rt_image *im = &s->d.images[t->perlin];
    if (im->height <= 0 || im->width <= 0)
      return V(0, 1, 1);
    double uc = u < 0 ? 0 : (u > 1 ? 1 : u);
    double vc = 1.0 - (v < 0 ? 0 : (v > 1 ? 1 : v));
    int i = (int)(uc * im->width), j = (int)(vc * im->height);
    if (i > im->width - 1)
      i = im->width - 1;
    if (j > im->height - 1)
      j = im->height - 1;
    const uint8_t *px = im->rgb + 3 * ((size_t)j * (size_t)im->width + (size_t)i);
    const double color_scale = 1.0 / 255.0;
    return V(color_scale * px[0], color_scale * px[1], color_scale * px[2]);
  }
  double f = 1 + sin(t->scale * p.z + 10 * perlin_turb(&s->d.perlins[t->perlin], p, 7));
  return vscale(f, V(0.5, 0.5, 0.5));
}

/* ------------------------------------------------------------------------------------------------
 * Samplers
 * ---------------------------------------------------------------------------------------------- */
/* The order in which GCC evaluates the random_double() calls inside one constructor / operator
 * expression is unspecified by the language; these are the orders the compiled reference shows
 * (pinned against oracle/_ref by tests/test_oracle_vs_ref.py). */
#ifndef ORA_ARGS_RIGHT_TO_LEFT
#define ORA_ARGS_RIGHT_TO_LEFT 1
#endif

/* Vec3::random(min,max) (Vec3.hpp:112-115). */
static v3 vec3_random_range(rng_t *r, double lo, double hi) {
#if ORA_ARGS_RIGHT_TO_LEFT
  double z = rnd_range(r, lo, hi), y = rnd_range(r, lo, hi), x = rnd_range(r, lo, hi);
#else
  double x = rnd_range(r, lo, hi), y = rnd_range(r, lo, hi), z = rnd_range(r, lo, hi);
#endif
  return V(x, y, z);
}
/* random_unit_vector, rejection (Vec3Utility.hpp:53-62). */
static v3 random_unit_vector_rejection(rng_t *r) {
  for (;;) {
    v3 v = vec3_random_range(r, -1, 1);
    double lensq = vlen2(v);
    if (1e-160 < lensq && lensq <= 1.0)
      return vdiv(v, sqrt(lensq));
  }
}
/* cuda_vec3_random_unit_vector, polar (Vec3Utility.cuh:65-70) from two uniforms. */
static v3 unit_vector_polar(double u1, double u2) {
  double z = -1.0 + 2.0 * u1;
  double a = 2.0 * ORA_PI * u2;
  double rr = sqrt(1.0 - z * z);
  return V(rr * cos(a), rr * sin(a), z);
}
/* random_cosine_direction (Vec3Utility.hpp:94-104). */
static v3 cosine_direction(double r1, double r2) {
  double phi = 2 * ORA_PI * r1;
  double x = cos(phi) * sqrt(r2);
  double y = sin(phi) * sqrt(r2);
  double z = sqrt(1 - r2);
  return V(x, y, z);
}
/* ONB (ONB.hpp:33-36,64). */
typedef struct onb {
  v3 u, v, w;
} onb;
static onb onb_make(v3 n) {
  onb b;
  b.w = vunit(n);
  v3 a = fabs(b.w.x) > 0.9 ? V(0, 1, 0) : V(1, 0, 0);
  b.v = vunit(vcross(b.w, a));
  b.u = vcross(b.w, b.v);
  return b;
}
static v3 onb_transform(const onb *b, v3 a) {
  return vadd(vadd(vscale(a.x, b->u), vscale(a.y, b->v)), vscale(a.z, b->w));
}

/* ------------------------------------------------------------------------------------------------
 * Lights: HittablePDF over the light list (PDF.hpp:82-113, HittableList.cpp:44-63, Plane.cpp:115-133,
 * Sphere.cpp:145-179)
 * ---------------------------------------------------------------------------------------------- */
typedef struct olight {
  int shape;
  osphere sp;
  oquad q;
} olight;

static olight light_make(const rt_light *l) {
  olight o;
  memset(&o, 0, sizeof o);
  o.shape = l->shape;
  if (l->shape == RT_SHAPE_SPHERE) {
    o.sp.c0 = vfrom(l->a);
    o.sp.cdir = V(0, 0, 0);
    o.sp.radius = fmax(0, l->radius);
    o.sp.material = -1;
  } else {
    o.q.corner = vfrom(l->a);
    o.q.u = vfrom(l->b);
    o.q.v = vfrom(l->c);
    v3 n = vcross(o.q.u, o.q.v);
    o.q.normal = vunit(n);
    o.q.D = vdot(o.q.normal, o.q.corner);
    o.q.w = vdiv(n, vdot(n, n));
    o.q.area = vlen(n);
    o.q.material = -1;
  }
  return o;
}

static double light_pdf_value(ora_scene *s, const rt_light *l, v3 origin, v3 dir) {
  olight o = light_make(l);
  oray r = {origin, dir, 0.0}; /* the reference's 2-arg Ray leaves the time unset; lights are static */
  ohit h;
  if (o.shape == RT_SHAPE_QUAD) {
    if (!quad_hit(s, &o.q, &r, 0.001, ORA_INF, &h))
      return 0;
    double distance_squared = h.t * h.t * vlen2(dir);
    double cosine = fabs(vdot(dir, h.normal) / vlen(dir));
    return distance_squared / (cosine * o.q.area);
  }
  if (!sphere_hit(s, &o.sp, &r, 0.001, ORA_INF, 0, &h))
    return 0;
  double dist_squared = vlen2(vsub(o.sp.c0, origin));
  double cos_theta_max = sqrt(1 - o.sp.radius * o.sp.radius / dist_squared);
  double solid_angle = 2 * ORA_PI * (1 - cos_theta_max);
  return 1 / solid_angle;
}

/* r1, r2: the two uniforms in the order the reference draws them. */
static v3 light_random(const rt_light *l, v3 origin, double r1, double r2) {
  olight o = light_make(l);
  if (o.shape == RT_SHAPE_QUAD) {
    v3 p = vadd(vadd(o.q.corner, vscale(r1, o.q.u)), vscale(r2, o.q.v));
    return vsub(p, origin);
  }
  v3 direction = vsub(o.sp.c0, origin);
  double distance_squared = vlen2(direction);
  onb uvw = onb_make(direction);
  double z = 1 + r2 * (sqrt(1 - o.sp.radius * o.sp.radius / distance_squared) - 1);
  double phi = 2 * ORA_PI * r1;
  double x = cos(phi) * sqrt(1 - z * z);
  double y = sin(phi) * sqrt(1 - z * z);
  return onb_transform(&uvw, V(x, y, z));
}

static double lights_pdf_value(ora_scene *s, v3 origin, v3 dir) {
  int n = s->d.n_lights;
  double weight = 1.0 / n;
  double sum = 0.0;
  for (int i = 0; i < n; i++)
    sum += weight * light_pdf_value(s, &s->d.lights[i], origin, dir);
  return sum;
}

/* ------------------------------------------------------------------------------------------------
 * Integrator: Camera::ray_color (Camera.cpp:232-309)
 * ---------------------------------------------------------------------------------------------- */
typedef struct render_ctx {
  ora_scene *s;
  rng_t *rng;
  int use_bvh;
  int max_depth;
  v3 background;
} render_ctx;

static v3 ray_color(render_ctx *c, const oray *ray, int depth) {
  ora_scene *s = c->s;
  rng_t *rng = c->rng;
  if (depth <= 0)
    return V(0, 0, 0);
  uint32_t bounce = (uint32_t)(c->max_depth - depth);
  s->cnt.segments++;
  ohit rec;
  if (!world_hit(s, c->use_bvh, ray, 0.001, ORA_INF, rng, bounce, &rec))
    return c->background;

  const rt_material *m = &s->d.materials[rec.material];
  int philox = rng->kind == ORA_RNG_PHILOX;
  double u4[4] = {0, 0, 0, 0};
  if (philox)
    philox_uniforms(rng, bounce, ORA_STREAM_SHADE, 0, u4);

  /* emitted: DiffuseLightMaterial.cpp:12-19, zero for every other material (Material.hpp) */
  v3 emitted = V(0, 0, 0);
  if (m->type == RT_MAT_DIFFUSE_LIGHT) {
    if (rec.front)
      emitted = texture_value(s, m->texture, rec.u, rec.v, rec.p);
    return emitted; /* no scatter */
  }

  if (m->type == RT_MAT_METAL) { /* MetalMaterial.cpp:46-61 */
    v3 reflected = vsub(ray->d, vscale(2 * vdot(ray->d, rec.normal), rec.normal));
    v3 fuzz_dir = philox ? unit_vector_polar(u4[1], u4[2]) : random_unit_vector_rejection(rng);
    reflected = vadd(vunit(reflected), vscale(m->fuzz, fuzz_dir));
    oray next = {rec.p, reflected, ray->time};
    return vmul(vfrom(m->albedo), ray_color(c, &next, depth - 1));
  }

  if (m->type == RT_MAT_DIELECTRIC) { /* DielectricMaterial.cpp:62-84 (+ incoming time, see header) */
    double ri = rec.front ? (1.0 / m->ior) : m->ior;
    v3 unit_direction = vunit(ray->d);
    double cos_theta = fmin(vdot(vneg(unit_direction), rec.normal), 1.0);
    double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    int cannot_refract = ri * sin_theta > 1.0;
    double r0 = (1 - ri) / (1 + ri);
    r0 = r0 * r0;
    double reflectance = r0 + (1 - r0) * pow((1 - cos_theta), 5);
    v3 direction;
    if (cannot_refract || reflectance > (philox ? u4[0] : rnd(rng))) {
      direction = vsub(unit_direction, vscale(2 * vdot(unit_direction, rec.normal), rec.normal));
    } else { /* refract, Vec3Utility.hpp:82-89 */
      double ct = fmin(vdot(vneg(unit_direction), rec.normal), 1.0);
      v3 perp = vscale(ri, vadd(unit_direction, vscale(ct, rec.normal)));
      v3 parallel = vscale(-sqrt(fabs(1.0 - vlen2(perp))), rec.normal);
      direction = vadd(perp, parallel);
    }
    oray next = {rec.p, direction, ray->time};
    return vmul(V(1.0, 1.0, 1.0), ray_color(c, &next, depth - 1));
  }

  /* Lambertian (LambertianMaterial.cpp:15-59) / Isotropic (IsotropicMaterial.cpp:12-31) */
  int lambert = m->type == RT_MAT_LAMBERTIAN;
  v3 attenuation = texture_value(s, m->texture, rec.u, rec.v, rec.p);
  onb uvw;
  if (lambert)
    uvw = onb_make(rec.normal); /* CosinePDF(record.normal) */
  int n_lights = s->d.n_lights;

  /* MixturePDF(light_or_material, material).generate() (PDF.hpp:135-143) */
  v3 dir;
  if (!philox) {
    int choose_first = rnd(rng) < 0.5;
    if (choose_first && n_lights > 0) { /* HittableList::random (HittableList.cpp:57-63) */
      int pick = rnd_int(rng, 0, n_lights - 1);
      const rt_light *l = &s->d.lights[pick];
      double r1, r2;
      if (l->shape == RT_SHAPE_QUAD) {
        /* Plane::random (Plane.cpp:128-133): the two draws sit in one expression */
#if ORA_ARGS_RIGHT_TO_LEFT
        r2 = rnd(rng);
        r1 = rnd(rng);
#else
        r1 = rnd(rng);
        r2 = rnd(rng);
#endif
      } else {
        r1 = rnd(rng);
        r2 = rnd(rng);
      }
      dir = light_random(l, rec.p, r1, r2);
    } else if (lambert) {
      double r1 = rnd(rng);
      double r2 = rnd(rng);
      v3 cd = cosine_direction(r1, r2);
      dir = onb_transform(&uvw, cd);
    } else {
      dir = random_unit_vector_rejection(rng);
    }
  } else {
    int choose_first = u4[0] < 0.5;
    if (choose_first && n_lights > 0) {
      int pick = (int)(u4[3] * n_lights);
      if (pick > n_lights - 1)
        pick = n_lights - 1;
      dir = light_random(&s->d.lights[pick], rec.p, u4[1], u4[2]);
    } else if (lambert) {
      v3 cd = cosine_direction(u4[1], u4[2]);
      dir = onb_transform(&uvw, cd);
    } else {
      dir = unit_vector_polar(u4[1], u4[2]);
    }
  }

  /* MixturePDF.value (PDF.hpp:131-133) */
  double mat_pdf = lambert ? fmax(0, vdot(vunit(dir), uvw.w) / ORA_PI) : 1.0 / (4.0 * ORA_PI);
  double first_pdf = n_lights > 0 ? lights_pdf_value(s, rec.p, dir) : mat_pdf;
  double pdf_value = 0.5 * first_pdf + 0.5 * mat_pdf;

  double scattering_pdf;
  if (lambert) {
    double cos_theta = vdot(rec.normal, vunit(dir));
    scattering_pdf = cos_theta < 0 ? 0 : cos_theta / ORA_PI;
  } else {
    scattering_pdf = 1 / (4 * ORA_PI);
  }

  if (philox && !(pdf_value > 1e-8)) /* reference CUDA path's guard (CameraKernels.cu:192) */
    return emitted;

  oray next = {rec.p, dir, ray->time};
  v3 sample_color = ray_color(c, &next, depth - 1);
  v3 scattered = vdiv(vmul(vscale(scattering_pdf, attenuation), sample_color), pdf_value);
  return vadd(emitted, scattered);
}

/* ------------------------------------------------------------------------------------------------
 * Camera (Camera.cpp:31-73,186-230)
 * ---------------------------------------------------------------------------------------------- */
void ora_camera_init(const rt_camera_config *cfg, rt_camera *out) {
  int W = cfg->image_width;
  int H = (int)(W / cfg->aspect_ratio);
  H = (H < 1) ? 1 : H;
  v3 center = vfrom(cfg->lookfrom);
  double theta = cfg->vfov * ORA_PI / 180.0;
  double h = tan(theta / 2);
  double viewport_height = 2 * h * cfg->focus_dist;
  double viewport_width = viewport_height * ((double)W / H);
  v3 w = vunit(vsub(vfrom(cfg->lookfrom), vfrom(cfg->lookat)));
  v3 u = vunit(vcross(vfrom(cfg->vup), w));
  v3 v = vcross(w, u);
  v3 viewport_u = vscale(viewport_width, u);
  v3 viewport_v = vscale(viewport_height, vneg(v));
  v3 du = vdiv(viewport_u, W);
  v3 dv = vdiv(viewport_v, H);
  v3 upper_left =
      vsub(vsub(vsub(center, vscale(cfg->focus_dist, w)), vdiv(viewport_u, 2)), vdiv(viewport_v, 2));
  v3 p00 = vadd(upper_left, vscale(0.5, vadd(du, dv)));
  double defocus_radius = cfg->focus_dist * tan((cfg->defocus_angle / 2) * ORA_PI / 180.0);
  memset(out, 0, sizeof *out);
  out->image_width = W;
  out->image_height = H;
  vto(out->center, center);
  vto(out->pixel00_loc, p00);
  vto(out->pixel_delta_u, du);
  vto(out->pixel_delta_v, dv);
  vto(out->defocus_disk_u, vscale(defocus_radius, u));
  vto(out->defocus_disk_v, vscale(defocus_radius, v));
  out->defocus_angle = cfg->defocus_angle;
  memcpy(out->background, cfg->background, sizeof out->background);
}

/* Camera::get_ray (Camera.cpp:186-205). */
static oray get_ray(const rt_camera *cam, int sqrt_spp, rng_t *rng, int i, int j, int s_i, int s_j) {
  double recip = 1.0 / sqrt_spp;
  double u[8] = {0};
  int philox = rng->kind == ORA_RNG_PHILOX;
  if (philox) {
    philox_uniforms(rng, 0, ORA_STREAM_CAMERA, 0, u);
    philox_uniforms(rng, 0, ORA_STREAM_CAMERA, 1, u + 4);
  }
  double px = ((s_i + (philox ? u[0] : rnd(rng))) * recip) - 0.5;
  double py = ((s_j + (philox ? u[1] : rnd(rng))) * recip) - 0.5;
  v3 p00 = vfrom(cam->pixel00_loc), du = vfrom(cam->pixel_delta_u), dv = vfrom(cam->pixel_delta_v);
  v3 pixel_sample = vadd(vadd(p00, vscale(i + px, du)), vscale(j + py, dv));
  v3 origin = vfrom(cam->center);
  if (!(cam->defocus_angle <= 0)) {
    double dx, dy;
    if (philox) { /* cuda_vec3_random_in_unit_disk (Vec3Utility.cuh:57-61) */
      double rr = sqrt(u[2]);
      double th = 2.0 * ORA_PI * u[3];
      dx = rr * cos(th);
      dy = rr * sin(th);
    } else { /* random_in_unit_disk (Vec3Utility.hpp:41-49) */
      for (;;) {
#if ORA_ARGS_RIGHT_TO_LEFT
        dy = rnd_range(rng, -1, 1);
        dx = rnd_range(rng, -1, 1);
#else
        dx = rnd_range(rng, -1, 1);
        dy = rnd_range(rng, -1, 1);
#endif
        if (dx * dx + dy * dy + 0.0 * 0.0 < 1)
          break;
      }
    }
    origin = vadd(vadd(vfrom(cam->center), vscale(dx, vfrom(cam->defocus_disk_u))),
                  vscale(dy, vfrom(cam->defocus_disk_v)));
  }
  oray r;
  r.o = origin;
  r.d = vsub(pixel_sample, origin);
  r.time = philox ? u[4] : rnd(rng);
  return r;
}

void ora_primary_rays(const rt_camera_config *cfg, int rng_kind, int sampler, uint64_t seed, int s_i, int s_j,
                      rt_ray *out) {
  rt_camera cam;
  ora_camera_init(cfg, &cam);
  int sqrt_spp = (int)sqrt((double)cfg->samples_per_pixel);
  ora_mt19937 mt;
  ora_mt_seed(&mt, (uint32_t)seed);
  rng_t rng = {rng_kind, sampler, &mt, seed, 0, 0, 0};
  int W = cam.image_width, H = cam.image_height;
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      rng.pixel = (uint32_t)j * (uint32_t)W + (uint32_t)i;
      rng.sample = (uint32_t)(s_j * sqrt_spp + s_i);
      oray r = get_ray(&cam, sqrt_spp, &rng, i, j, s_i, s_j);
      rt_ray *o = &out[(size_t)j * W + i];
      memset(o, 0, sizeof *o);
      vto(o->origin, r.o);
      vto(o->direction, r.d);
      o->time = r.time;
      o->t_min = 0.001;
      o->t_max = ORA_INF;
      o->rng_pixel = rng.pixel;
      o->rng_sample = rng.sample;
      o->rng_bounce = 0;
    }
}

void ora_trace(ora_scene *s, const rt_ray *rays, int64_t n, int use_bvh, int rng_kind, uint64_t seed,
               rt_hit *hits) {
  ora_mt19937 mt;
  ora_mt_seed(&mt, (uint32_t)seed);
  rng_t rng = {rng_kind, ORA_SAMPLER_REJECTION, &mt, seed, 0, 0, 0};
  for (int64_t k = 0; k < n; k++) {
    const rt_ray *q = &rays[k];
    oray r = {vfrom(q->origin), vfrom(q->direction), q->time};
    rng.pixel = q->rng_pixel;
    rng.sample = q->rng_sample;
    ohit h;
    int ok = world_hit(s, use_bvh, &r, q->t_min, q->t_max, &rng, q->rng_bounce, &h);
    hits[k].t = ok ? h.t : ORA_INF;
    hits[k].prim = ok ? h.prim : -1;
    hits[k].object = ok ? h.object : -1;
    hits[k].front_face = ok ? h.front : 0;
    hits[k].pad_ = 0;
  }
}

double ora_render(ora_scene *s, const rt_camera_config *cfg, int rng_kind, int sampler, uint64_t seed, int use_bvh,
                  int row0, int row1, int single_stratum, double *out, ora_counters *counters) {
  rt_camera cam;
  ora_camera_init(cfg, &cam);
  int W = cam.image_width, H = cam.image_height;
  if (row1 > H || row1 < 0)
    row1 = H;
  int sqrt_spp = (int)sqrt((double)cfg->samples_per_pixel);
  double scale = single_stratum >= 0 ? 1.0 : 1.0 / cfg->samples_per_pixel;
  ora_mt19937 mt;
  ora_mt_seed(&mt, (uint32_t)seed);
  rng_t rng = {rng_kind, sampler, &mt, seed, 0, 0, 0};
  render_ctx c = {s, &rng, use_bvh, cfg->max_depth, vfrom(cfg->background)};
  memset(&s->cnt, 0, sizeof s->cnt);

  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int j = row0; j < row1; j++)
    for (int i = 0; i < W; i++) {
      v3 pixel = V(0, 0, 0);
      rng.pixel = (uint32_t)j * (uint32_t)W + (uint32_t)i;
      if (single_stratum >= 0) {
        int s_i = single_stratum % sqrt_spp, s_j = single_stratum / sqrt_spp;
        rng.sample = (uint32_t)single_stratum;
        oray r = get_ray(&cam, sqrt_spp, &rng, i, j, s_i, s_j);
        pixel = vadd(pixel, ray_color(&c, &r, cfg->max_depth));
        s->cnt.paths++;
      } else {
        for (int s_j = 0; s_j < sqrt_spp; ++s_j)
          for (int s_i = 0; s_i < sqrt_spp; ++s_i) {
            rng.sample = (uint32_t)(s_j * sqrt_spp + s_i);
            oray r = get_ray(&cam, sqrt_spp, &rng, i, j, s_i, s_j);
            pixel = vadd(pixel, ray_color(&c, &r, cfg->max_depth));
            s->cnt.paths++;
          }
      }
      v3 col = vscale(scale, pixel);
      double *o = out + ((size_t)(j - row0) * W + i) * 3;
      o[0] = col.x;
      o[1] = col.y;
      o[2] = col.z;
    }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  s->cnt.rng_draws = rng.draws;
  if (counters)
    *counters = s->cnt;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* to_byte (ColorUtility.hpp:11-26). */
int ora_to_byte(double v) {
  double x = v > 0 ? sqrt(v) : 0;
  if (x < 0.000)
    x = 0.000;
  if (x > 0.999)
    x = 0.999;
  return (int)(unsigned char)(256 * x);
}
