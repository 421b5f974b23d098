// JSON scene files.  The reference advertises a JSON scene format but only ever *dumps* JSON for
// debugging (core/camera/Camera.cpp:75-150 and every class's json() method); this loader defines the
// input format by mirroring the dump's class and field names:
//
//   { "camera":   {"type":"CameraConfig","aspect_ratio":..,"vfov":..,"defocus_angle":..,"focus_dist":..,
//                  "lookfrom":{"x":..,"y":..,"z":..},"lookat":{..},"vup":{..},"background":{..}},
//     "perlins":  [{"type":"PerlinNoise","rand_vec":[[x,y,z] x256],"perm_x":[..],"perm_y":[..],"perm_z":[..]}],
//     "textures": [{"type":"SolidColorTexture","albedo":{..}},
//                  {"type":"CheckerTexture","scale":..,"even_texture":<index>,"odd_texture":<index>},
//                  {"type":"NoiseTexture","scale":..,"perlin":<index>},
//                  {"type":"ImageTexture","image":<index>}],           (not in the reference: rt_image)
//     "images":   [{"type":"Image","file":"<name>.ppm"}],              (binary or ASCII PPM beside the JSON)
//     "materials":[{"type":"LambertianMaterial","texture":<index>}, {"type":"MetalMaterial","albedo":{..},"fuzz":..},
//                  {"type":"DielectricMaterial","refraction_index":..}, {"type":"DiffuseLightMaterial","texture":<index>},
//                  {"type":"IsotropicMaterial","texture":<index>}],
//     "world":    [ <hittable>, ... ],     "lights": [ <Sphere or Plane without material>, ... ] }
//   <hittable> = {"type":"Sphere","center":{"origin":{..},"direction":{..}},"radius":..,"material":<index>}
//              | {"type":"Plane","corner":{..},"u_side":{..},"v_side":{..},"material":<index>}
//              | {"type":"HittableList","objects":[<Sphere or Plane>, ...]}          (e.g. the six sides of a box)
//              | {"type":"Translate","offset":{..},"object":<hittable>} | {"type":"RotateY","angle":..,"object":<hittable>}
//              | {"type":"ConstantMedium","density":..,"boundary":<hittable>,"phase_function":<material index>}
// Numbers are written with 17 significant digits, so a saved scene loads back bit for bit.
#include "../../include/rt_host.h"
#include "scene_builder.hpp"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace rth {
void set_error(const std::string &msg);

namespace {

// ---- a small JSON value + recursive-descent parser ----
struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  double number = 0;
  bool boolean = false;
  std::string string;
  std::vector<Json> array;
  std::vector<std::pair<std::string, Json>> object;

  const Json *find(const std::string &key) const {
    for (const auto &kv : object)
      if (kv.first == key)
        return &kv.second;
    return nullptr;
  }
  const Json &at(const std::string &key) const {
    const Json *v = find(key);
    if (!v)
      throw std::runtime_error("missing field \"" + key + "\"");
    return *v;
  }
  double num() const {
    if (kind != Number)
      throw std::runtime_error("expected a number");
    return number;
  }
  int integer() const { return int(num()); }
  const std::string &str() const {
    if (kind != String)
      throw std::runtime_error("expected a string");
    return string;
  }
};

class Parser {
public:
  explicit Parser(const std::string &text) : m_text(text) {}
  Json parse() {
    Json v = value();
    skip();
    if (m_pos != m_text.size())
      fail("trailing characters");
    return v;
  }

private:
  const std::string &m_text;
  size_t m_pos = 0;
  int m_depth = 0;
  [[noreturn]] void fail(const std::string &what) {
    throw std::runtime_error("JSON: " + what + " at offset " + std::to_string(m_pos));
  }
  void skip() {
    while (m_pos < m_text.size() && std::isspace((unsigned char)m_text[m_pos]))
      m_pos++;
  }
  char peek() {
    skip();
    if (m_pos >= m_text.size())
      fail("unexpected end");
    return m_text[m_pos];
  }
  void expect(char c) {
    if (peek() != c)
      fail(std::string("expected '") + c + "'");
    m_pos++;
  }
  Json value() {
    // scene files nest a handful of levels (world -> wrapper chain -> object -> vector); a bound keeps a
    // damaged or hostile file from exhausting the stack of this recursive parser
    struct Depth {
      int &d;
      explicit Depth(int &depth) : d(depth) { ++d; }
      ~Depth() { --d; }
    } guard(m_depth);
    if (m_depth > 256)
      fail("nesting deeper than 256 levels");
    char c = peek();
    Json v;
    if (c == '{') {
      v.kind = Json::Object;
      m_pos++;
      if (peek() == '}') {
        m_pos++;
        return v;
      }
      for (;;) {
        Json key = value();
        if (key.kind != Json::String)
          fail("object key must be a string");
        expect(':');
        v.object.emplace_back(key.string, value());
        if (peek() == ',') {
          m_pos++;
          continue;
        }
        expect('}');
        return v;
      }
    }
    if (c == '[') {
      v.kind = Json::Array;
      m_pos++;
      if (peek() == ']') {
        m_pos++;
        return v;
      }
      for (;;) {
        v.array.push_back(value());
        if (peek() == ',') {
          m_pos++;
          continue;
        }
        expect(']');
        return v;
      }
    }
    if (c == '"') {
      v.kind = Json::String;
      m_pos++;
      while (m_pos < m_text.size() && m_text[m_pos] != '"') {
        if (m_text[m_pos] == '\\' && m_pos + 1 < m_text.size())
          m_pos++;
        v.string.push_back(m_text[m_pos++]);
      }
      if (m_pos >= m_text.size())
        fail("unterminated string");
      m_pos++;
      return v;
    }
    if (m_text.compare(m_pos, 4, "true") == 0) {
      v.kind = Json::Bool;
      v.boolean = true;
      m_pos += 4;
      return v;
    }
    if (m_text.compare(m_pos, 5, "false") == 0) {
      v.kind = Json::Bool;
      m_pos += 5;
      return v;
    }
    if (m_text.compare(m_pos, 4, "null") == 0) {
      m_pos += 4;
      return v;
    }
    const char *start = m_text.c_str() + m_pos;
    char *end = nullptr;
    double d = std::strtod(start, &end);
    if (end == start)
      fail("unexpected character");
    m_pos += size_t(end - start);
    v.kind = Json::Number;
    v.number = d;
    return v;
  }
};

Vec vec_of(const Json &j) { return Vec(j.at("x").num(), j.at("y").num(), j.at("z").num()); }

// ---- loading ----
struct Loader {
  SceneBuilder &s;

  void leaf(const Json &j, int xf, int object, int flags, int forced_material, bool has_forced) {
    const std::string &type = j.at("type").str();
    int material = has_forced ? forced_material : j.at("material").integer();
    if (type == "Sphere") {
      const Json &c = j.at("center");
      rt_sphere sp{};
      store(sp.center0, vec_of(c.at("origin")));
      store(sp.center_dir, vec_of(c.at("direction")));
      sp.radius = std::fmax(0, j.at("radius").num());
      sp.material = material;
      sp.xform = xf;
      sp.object = object;
      sp.flags = flags;
      s.spheres.push_back(sp);
    } else if (type == "Plane") {
      rt_quad q{};
      store(q.corner, vec_of(j.at("corner")));
      store(q.u, vec_of(j.at("u_side")));
      store(q.v, vec_of(j.at("v_side")));
      q.material = material;
      q.xform = xf;
      q.object = object;
      q.flags = flags;
      s.quads.push_back(q);
    } else {
      throw std::runtime_error("unsupported primitive type \"" + type + "\"");
    }
  }

  // Peels instance wrappers off `j`, then emits the primitives of the innermost object.
  void members(const Json &j, std::vector<XformOp> chain, int object, int flags, int forced_material, bool has_forced) {
    const Json *cur = &j;
    for (;;) {
      const std::string &type = cur->at("type").str();
      if (type == "Translate") {
        chain.push_back(translate(vec_of(cur->at("offset"))));
        cur = &cur->at("object");
      } else if (type == "RotateY") {
        chain.push_back(rotate_y(cur->at("angle").num()));
        cur = &cur->at("object");
      } else {
        break;
      }
    }
    int xf = s.xform(chain);
    if (cur->at("type").str() == "HittableList") {
      for (const Json &m : cur->at("objects").array)
        leaf(m, xf, object, flags, forced_material, has_forced);
    } else {
      leaf(*cur, xf, object, flags, forced_material, has_forced);
    }
  }

  void world_object(const Json &j) {
    // a ConstantMedium may itself sit under wrappers in the reference's object graph; the flat format
    // keeps the wrappers on the boundary, which is equivalent
    const Json *cur = &j;
    std::vector<XformOp> chain;
    for (;;) {
      const std::string &type = cur->at("type").str();
      if (type == "Translate") {
        chain.push_back(translate(vec_of(cur->at("offset"))));
        cur = &cur->at("object");
      } else if (type == "RotateY") {
        chain.push_back(rotate_y(cur->at("angle").num()));
        cur = &cur->at("object");
      } else {
        break;
      }
    }
    int object = s.n_objects++;
    if (cur->at("type").str() == "ConstantMedium") {
      size_t s0 = s.spheres.size(), q0 = s.quads.size();
      members(cur->at("boundary"), chain, object, RT_PRIM_BOUNDARY, -1, true);
      size_t ns = s.spheres.size() - s0, nq = s.quads.size() - q0;
      if ((ns == 0) == (nq == 0))
        throw std::runtime_error("a medium boundary must consist of spheres or of planes");
      rt_medium m{};
      m.density = cur->at("density").num();
      m.shape = ns ? RT_SHAPE_SPHERE : RT_SHAPE_QUAD;
      m.first_prim = int(ns ? s0 : q0);
      m.n_prims = int(ns ? ns : nq);
      m.material = cur->at("phase_function").integer();
      m.object = object;
      s.media.push_back(m);
    } else {
      members(*cur, chain, object, 0, 0, false);
    }
  }
};

// PPM reader for image textures: binary P6 or ASCII P3, maxval 255.
bool read_ppm(const std::string &path, int &w, int &h, std::vector<uint8_t> &rgb) {
  std::ifstream in(path, std::ios::binary);
  if (!in)
    return false;
  std::string magic;
  in >> magic;
  auto next_int = [&]() {
    for (;;) {
      in >> std::ws;
      if (in.peek() == '#') {
        std::string comment;
        std::getline(in, comment);
      } else {
        break;
      }
    }
    int v = -1;
    in >> v;
    return v;
  };
  w = next_int();
  h = next_int();
  int maxval = next_int();
  if ((magic != "P6" && magic != "P3") || w <= 0 || h <= 0 || maxval != 255)
    return false;
  if (int64_t(w) * int64_t(h) > (int64_t(1) << 28)) // the scene description's own limit (rt_flatten.h)
    return false;
  rgb.resize(size_t(w) * size_t(h) * 3);
  if (magic == "P6") {
    in.get(); // the single whitespace after maxval
    in.read(reinterpret_cast<char *>(rgb.data()), std::streamsize(rgb.size()));
    return size_t(in.gcount()) == rgb.size();
  }
  for (size_t k = 0; k < rgb.size(); k++) {
    int v = -1;
    in >> v;
    if (v < 0 || v > 255)
      return false;
    rgb[k] = uint8_t(v);
  }
  return true;
}

void load_scene(const Json &root, SceneBuilder &s, const std::string &base_dir) {
  if (const Json *arr = root.find("images"))
    for (const Json &im : arr->array) {
      int w = 0, h = 0;
      std::vector<uint8_t> rgb;
      std::string file = im.at("file").str();
      std::string path = (!file.empty() && file[0] == '/') ? file : base_dir + file;
      if (!read_ppm(path, w, h, rgb))
        throw std::runtime_error("cannot read image texture " + path + " (binary or ASCII PPM, maxval 255)");
      s.image(w, h, rgb);
    }
  const Json &cam = root.at("camera");
  s.camera = rt_camera_config{};
  s.camera.aspect_ratio = cam.at("aspect_ratio").num();
  s.camera.vfov = cam.at("vfov").num();
  s.camera.defocus_angle = cam.at("defocus_angle").num();
  s.camera.focus_dist = cam.at("focus_dist").num();
  store(s.camera.lookfrom, vec_of(cam.at("lookfrom")));
  store(s.camera.lookat, vec_of(cam.at("lookat")));
  store(s.camera.vup, vec_of(cam.at("vup")));
  store(s.camera.background, vec_of(cam.at("background")));
  if (const Json *v = cam.find("image_width"))
    s.camera.image_width = v->integer();
  if (const Json *v = cam.find("samples_per_pixel"))
    s.camera.samples_per_pixel = v->integer();
  if (const Json *v = cam.find("max_depth"))
    s.camera.max_depth = v->integer();

  if (const Json *arr = root.find("perlins"))
    for (const Json &p : arr->array) {
      rt_perlin t{};
      const Json &rv = p.at("rand_vec");
      if (rv.array.size() != RT_PERLIN_POINTS)
        throw std::runtime_error("perlin rand_vec must have 256 entries");
      for (int i = 0; i < RT_PERLIN_POINTS; i++)
        for (int a = 0; a < 3; a++)
          t.rand_vec[i][a] = rv.array[i].array.at(a).num();
      const char *names[3] = {"perm_x", "perm_y", "perm_z"};
      int32_t *perms[3] = {t.perm_x, t.perm_y, t.perm_z};
      for (int k = 0; k < 3; k++) {
        const Json &pm = p.at(names[k]);
        if (pm.array.size() != RT_PERLIN_POINTS)
          throw std::runtime_error("perlin permutation must have 256 entries");
        for (int i = 0; i < RT_PERLIN_POINTS; i++)
          perms[k][i] = pm.array[i].integer();
      }
      s.perlins.push_back(t);
    }
  if (const Json *arr = root.find("textures"))
    for (const Json &t : arr->array) {
      const std::string &type = t.at("type").str();
      rt_texture r{};
      r.even = r.odd = r.perlin = -1;
      if (type == "SolidColorTexture") {
        r.type = RT_TEX_SOLID;
        store(r.color, vec_of(t.at("albedo")));
      } else if (type == "CheckerTexture") {
        r.type = RT_TEX_CHECKER;
        r.scale = t.at("scale").num();
        r.even = t.at("even_texture").integer();
        r.odd = t.at("odd_texture").integer();
      } else if (type == "NoiseTexture") {
        r.type = RT_TEX_NOISE;
        r.scale = t.at("scale").num();
        r.perlin = t.at("perlin").integer();
      } else if (type == "ImageTexture") {
        r.type = RT_TEX_IMAGE;
        r.perlin = t.at("image").integer();
      } else {
        throw std::runtime_error("unknown texture type \"" + type + "\"");
      }
      s.textures.push_back(r);
    }
  if (const Json *arr = root.find("materials"))
    for (const Json &m : arr->array) {
      const std::string &type = m.at("type").str();
      rt_material r{};
      r.texture = -1;
      if (type == "LambertianMaterial") {
        r.type = RT_MAT_LAMBERTIAN;
        r.texture = m.at("texture").integer();
      } else if (type == "MetalMaterial") {
        r.type = RT_MAT_METAL;
        store(r.albedo, vec_of(m.at("albedo")));
        r.fuzz = m.at("fuzz").num();
      } else if (type == "DielectricMaterial") {
        r.type = RT_MAT_DIELECTRIC;
        r.ior = m.at("refraction_index").num();
      } else if (type == "DiffuseLightMaterial") {
        r.type = RT_MAT_DIFFUSE_LIGHT;
        r.texture = m.at("texture").integer();
      } else if (type == "IsotropicMaterial") {
        r.type = RT_MAT_ISOTROPIC;
        r.texture = m.at("texture").integer();
      } else {
        throw std::runtime_error("unknown material type \"" + type + "\"");
      }
      s.materials.push_back(r);
    }
  Loader loader{s};
  for (const Json &o : root.at("world").array)
    loader.world_object(o);
  if (const Json *arr = root.find("lights"))
    for (const Json &l : arr->array) {
      const std::string &type = l.at("type").str();
      if (type == "Sphere")
        s.light_sphere(vec_of(l.at("center").at("origin")), l.at("radius").num());
      else if (type == "Plane")
        s.light_quad(vec_of(l.at("corner")), vec_of(l.at("u_side")), vec_of(l.at("v_side")));
      else
        throw std::runtime_error("unsupported light type \"" + type + "\"");
    }
  s.finalize();
}

// ---- saving ----
std::string num(double v) {
  char buf[40];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  return buf;
}
std::string vec_json(const double *p) {
  return "{\"type\":\"Vec3\",\"x\":" + num(p[0]) + ",\"y\":" + num(p[1]) + ",\"z\":" + num(p[2]) + "}";
}

void save_scene(const SceneBuilder &s, std::ostream &out, const std::string &json_path) {
  // image textures are written beside the JSON as binary PPM files and referenced by file name
  out << "{\n\"images\":[";
  for (size_t i = 0; i < s.images.size(); i++) {
    std::string file = json_path + ".image" + std::to_string(i) + ".ppm";
    std::ofstream img(file, std::ios::binary);
    img << "P6\n" << s.images[i].width << " " << s.images[i].height << "\n255\n";
    img.write(reinterpret_cast<const char *>(s.image_data[i].data()), std::streamsize(s.image_data[i].size()));
    size_t slash = file.find_last_of('/');
    out << (i ? "," : "") << "{\"type\":\"Image\",\"width\":" << s.images[i].width << ",\"height\":" << s.images[i].height
        << ",\"file\":\"" << (slash == std::string::npos ? file : file.substr(slash + 1)) << "\"}";
  }
  out << "],\n";
  const rt_camera_config &c = s.camera;
  out << "\"camera\":{\"type\":\"CameraConfig\",\"aspect_ratio\":" << num(c.aspect_ratio) << ",\"vfov\":" << num(c.vfov)
      << ",\"defocus_angle\":" << num(c.defocus_angle) << ",\"focus_dist\":" << num(c.focus_dist)
      << ",\"lookfrom\":" << vec_json(c.lookfrom) << ",\"lookat\":" << vec_json(c.lookat) << ",\"vup\":" << vec_json(c.vup)
      << ",\"background\":" << vec_json(c.background) << "},\n";
  out << "\"perlins\":[";
  for (size_t i = 0; i < s.perlins.size(); i++) {
    const rt_perlin &p = s.perlins[i];
    out << (i ? ",\n" : "\n") << "{\"type\":\"PerlinNoise\",\"rand_vec\":[";
    for (int k = 0; k < RT_PERLIN_POINTS; k++)
      out << (k ? "," : "") << "[" << num(p.rand_vec[k][0]) << "," << num(p.rand_vec[k][1]) << "," << num(p.rand_vec[k][2]) << "]";
    out << "]";
    const char *names[3] = {"perm_x", "perm_y", "perm_z"};
    const int32_t *perms[3] = {p.perm_x, p.perm_y, p.perm_z};
    for (int a = 0; a < 3; a++) {
      out << ",\"" << names[a] << "\":[";
      for (int k = 0; k < RT_PERLIN_POINTS; k++)
        out << (k ? "," : "") << perms[a][k];
      out << "]";
    }
    out << "}";
  }
  out << "],\n\"textures\":[";
  for (size_t i = 0; i < s.textures.size(); i++) {
    const rt_texture &t = s.textures[i];
    out << (i ? ",\n" : "\n");
    if (t.type == RT_TEX_SOLID)
      out << "{\"type\":\"SolidColorTexture\",\"albedo\":" << vec_json(t.color) << "}";
    else if (t.type == RT_TEX_CHECKER)
      out << "{\"type\":\"CheckerTexture\",\"scale\":" << num(t.scale) << ",\"even_texture\":" << t.even
          << ",\"odd_texture\":" << t.odd << "}";
    else if (t.type == RT_TEX_IMAGE)
      out << "{\"type\":\"ImageTexture\",\"image\":" << t.perlin << "}";
    else
      out << "{\"type\":\"NoiseTexture\",\"scale\":" << num(t.scale) << ",\"perlin\":" << t.perlin << "}";
  }
  out << "],\n\"materials\":[";
  for (size_t i = 0; i < s.materials.size(); i++) {
    const rt_material &m = s.materials[i];
    out << (i ? ",\n" : "\n");
    switch (m.type) {
    case RT_MAT_LAMBERTIAN:
      out << "{\"type\":\"LambertianMaterial\",\"texture\":" << m.texture << "}";
      break;
    case RT_MAT_METAL:
      out << "{\"type\":\"MetalMaterial\",\"albedo\":" << vec_json(m.albedo) << ",\"fuzz\":" << num(m.fuzz) << "}";
      break;
    case RT_MAT_DIELECTRIC:
      out << "{\"type\":\"DielectricMaterial\",\"refraction_index\":" << num(m.ior) << "}";
      break;
    case RT_MAT_DIFFUSE_LIGHT:
      out << "{\"type\":\"DiffuseLightMaterial\",\"texture\":" << m.texture << "}";
      break;
    default:
      out << "{\"type\":\"IsotropicMaterial\",\"texture\":" << m.texture << "}";
    }
  }
  out << "],\n\"world\":[";

  // group primitives by top-level object, keeping array order inside an object
  std::vector<std::vector<int>> obj_spheres(s.n_objects), obj_quads(s.n_objects);
  std::vector<int> obj_medium(s.n_objects, -1);
  for (size_t i = 0; i < s.spheres.size(); i++)
    obj_spheres[s.spheres[i].object].push_back(int(i));
  for (size_t i = 0; i < s.quads.size(); i++)
    obj_quads[s.quads[i].object].push_back(int(i));
  for (size_t m = 0; m < s.media.size(); m++)
    obj_medium[s.media[m].object] = int(m);
  auto sphere_json = [&](const rt_sphere &sp, bool with_material) {
    std::string r = "{\"type\":\"Sphere\",\"center\":{\"type\":\"Ray\",\"origin\":" + vec_json(sp.center0) +
                    ",\"direction\":" + vec_json(sp.center_dir) + "},\"radius\":" + num(sp.radius);
    if (with_material)
      r += ",\"material\":" + std::to_string(sp.material);
    return r + "}";
  };
  auto quad_json = [&](const rt_quad &q, bool with_material) {
    std::string r = "{\"type\":\"Plane\",\"corner\":" + vec_json(q.corner) + ",\"u_side\":" + vec_json(q.u) +
                    ",\"v_side\":" + vec_json(q.v);
    if (with_material)
      r += ",\"material\":" + std::to_string(q.material);
    return r + "}";
  };
  for (int k = 0; k < s.n_objects; k++) {
    bool medium = obj_medium[k] >= 0;
    std::vector<std::string> members;
    int xf = -1;
    for (int i : obj_spheres[k]) {
      members.push_back(sphere_json(s.spheres[i], !medium));
      xf = s.spheres[i].xform;
    }
    for (int i : obj_quads[k]) {
      members.push_back(quad_json(s.quads[i], !medium));
      xf = s.quads[i].xform;
    }
    std::string inner;
    if (members.size() == 1) {
      inner = members[0];
    } else {
      inner = "{\"type\":\"HittableList\",\"objects\":[";
      for (size_t i = 0; i < members.size(); i++)
        inner += (i ? "," : "") + members[i];
      inner += "]}";
    }
    if (xf >= 0) {
      const rt_xform &x = s.xforms[xf];
      for (int o = x.n_ops - 1; o >= 0; o--) { // innermost wrapper first
        const rt_xform_op &op = s.xform_ops[x.first_op + o];
        if (op.type == RT_XF_TRANSLATE)
          inner = "{\"type\":\"Translate\",\"offset\":" + vec_json(op.offset) + ",\"object\":" + inner + "}";
        else
          inner = "{\"type\":\"RotateY\",\"angle\":" + num(op.angle_deg) + ",\"object\":" + inner + "}";
      }
    }
    if (medium) {
      const rt_medium &m = s.media[obj_medium[k]];
      inner = "{\"type\":\"ConstantMedium\",\"density\":" + num(m.density) + ",\"boundary\":" + inner +
              ",\"phase_function\":" + std::to_string(m.material) + "}";
    }
    out << (k ? ",\n" : "\n") << inner;
  }
  out << "],\n\"lights\":[";
  for (size_t i = 0; i < s.lights.size(); i++) {
    const rt_light &l = s.lights[i];
    out << (i ? ",\n" : "\n");
    if (l.shape == RT_SHAPE_SPHERE) {
      double zero[3] = {0, 0, 0};
      out << "{\"type\":\"Sphere\",\"center\":{\"type\":\"Ray\",\"origin\":" << vec_json(l.a) << ",\"direction\":" << vec_json(zero)
          << "},\"radius\":" << num(l.radius) << "}";
    } else {
      out << "{\"type\":\"Plane\",\"corner\":" << vec_json(l.a) << ",\"u_side\":" << vec_json(l.b) << ",\"v_side\":" << vec_json(l.c)
          << "}";
    }
  }
  out << "]\n}\n";
}

} // namespace
} // namespace rth

extern "C" {

rth_scene *rth_scene_load_json(const char *path) {
  std::ifstream in(path ? path : "");
  if (!in) {
    rth::set_error(std::string("cannot open ") + (path ? path : "(null)"));
    return nullptr;
  }
  std::stringstream ss;
  ss << in.rdbuf();
  std::string text = ss.str();
  rth_scene *s = new rth_scene();
  try {
    rth::Json root = rth::Parser(text).parse();
    std::string p(path);
    size_t slash = p.find_last_of('/');
    rth::load_scene(root, s->builder, slash == std::string::npos ? std::string() : p.substr(0, slash + 1));
  } catch (const std::exception &e) {
    rth::set_error(std::string(path) + ": " + e.what());
    delete s;
    return nullptr;
  }
  return s;
}

int rth_scene_save_json(const rth_scene *scene, const char *path) {
  std::ofstream out(path ? path : "");
  if (!out) {
    rth::set_error(std::string("cannot open ") + (path ? path : "(null)"));
    return 1;
  }
  const_cast<rth_scene *>(scene)->builder.finalize();
  rth::save_scene(scene->builder, out, path);
  return out.good() ? 0 : 1;
}

} // extern "C"
