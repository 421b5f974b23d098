// Progressive window of the dynamic camera (DynamicCamera.cpp:62-91,196-306): an SDL3 window with a streaming
// RGB24 texture that is refreshed from the frame the GPU resolved, the reference's key handling (W/A/S/D held =
// move, = / - pressed = samples per pixel, ESC / window close = quit) and the FPS line (in the window title: the
// reference draws it with SDL_ttf, which this build does not require).
//
// SDL3 is not linked: libSDL3 is opened at run time (RT_SDL3_LIB overrides the file name), so the host library
// has no build- or load-time dependency on it and a machine without SDL3 gets a clear error and the headless
// mode.  The handful of SDL 3.2 declarations used here are restated below (values from SDL_init.h,
// SDL_pixels.h, SDL_render.h, SDL_events.h, SDL_keycode.h, SDL_scancode.h of SDL 3.2.x, the release the
// reference builds against, build.sh:56-59).
#include "../../include/rt_host.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>

namespace rth {
void set_error(const std::string &msg);
}

namespace {

// --- SDL 3.2 ABI subset -----------------------------------------------------------------------------
constexpr uint32_t kSDL_INIT_VIDEO = 0x00000020u;
constexpr uint32_t kSDL_PIXELFORMAT_RGB24 = 0x17101803u;
constexpr int kSDL_TEXTUREACCESS_STREAMING = 1;
constexpr uint32_t kSDL_EVENT_QUIT = 0x100u;
constexpr uint32_t kSDL_EVENT_KEY_DOWN = 0x300u;
constexpr uint32_t kSDLK_ESCAPE = 0x1Bu, kSDLK_EQUALS = 0x3Du, kSDLK_MINUS = 0x2Du;
constexpr int kSDL_SCANCODE_A = 4, kSDL_SCANCODE_D = 7, kSDL_SCANCODE_S = 22, kSDL_SCANCODE_W = 26;

struct SdlKeyboardEvent { // SDL_KeyboardEvent
  uint32_t type, reserved;
  uint64_t timestamp;
  uint32_t window_id, which, scancode, key;
  uint16_t mod, raw;
  uint8_t down, repeat;
};
union SdlEvent { // SDL_Event: 128 bytes
  uint32_t type;
  SdlKeyboardEvent key;
  uint8_t padding[128];
};
static_assert(sizeof(SdlEvent) == 128, "SDL_Event is 128 bytes");

struct Sdl {
  void *lib = nullptr;
  bool (*Init)(uint32_t) = nullptr;
  void (*Quit)() = nullptr;
  const char *(*GetError)() = nullptr;
  void *(*CreateWindow)(const char *, int, int, uint64_t) = nullptr;
  void *(*CreateRenderer)(void *, const char *) = nullptr;
  void *(*CreateTexture)(void *, uint32_t, int, int, int) = nullptr;
  bool (*UpdateTexture)(void *, const void *, const void *, int) = nullptr;
  bool (*RenderClear)(void *) = nullptr;
  bool (*RenderTexture)(void *, void *, const void *, const void *) = nullptr;
  bool (*RenderPresent)(void *) = nullptr;
  bool (*PollEvent)(SdlEvent *) = nullptr;
  const bool *(*GetKeyboardState)(int *) = nullptr;
  bool (*SetWindowTitle)(void *, const char *) = nullptr;
  void (*DestroyTexture)(void *) = nullptr;
  void (*DestroyRenderer)(void *) = nullptr;
  void (*DestroyWindow)(void *) = nullptr;
};

template <class F> bool resolve(void *lib, const char *name, F &fn, std::string &missing) {
  fn = reinterpret_cast<F>(dlsym(lib, name));
  if (!fn)
    missing += std::string(missing.empty() ? "" : ", ") + name;
  return fn != nullptr;
}

} // namespace

struct rth_presenter {
  Sdl sdl;
  void *window = nullptr, *renderer = nullptr, *texture = nullptr;
  int width = 0, height = 0;
};

extern "C" {

rth_presenter *rth_presenter_open(int width, int height, const char *title) {
  if (width < 1 || height < 1) {
    rth::set_error("presenter: image size must be positive");
    return nullptr;
  }
  const char *override_name = std::getenv("RT_SDL3_LIB");
  const char *names[] = {override_name, "libSDL3.so.0", "libSDL3.so"};
  void *lib = nullptr;
  std::string tried;
  for (const char *name : names) {
    if (!name || !*name)
      continue;
    lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (lib)
      break;
    tried += std::string(tried.empty() ? "" : "; ") + dlerror();
  }
  if (!lib) {
    rth::set_error("SDL3 is not available (" + tried + "); run the dynamic camera headless with --frames <n>");
    return nullptr;
  }
  rth_presenter *p = new rth_presenter;
  Sdl &s = p->sdl;
  s.lib = lib;
  std::string missing;
  resolve(lib, "SDL_Init", s.Init, missing);
  resolve(lib, "SDL_Quit", s.Quit, missing);
  resolve(lib, "SDL_GetError", s.GetError, missing);
  resolve(lib, "SDL_CreateWindow", s.CreateWindow, missing);
  resolve(lib, "SDL_CreateRenderer", s.CreateRenderer, missing);
  resolve(lib, "SDL_CreateTexture", s.CreateTexture, missing);
  resolve(lib, "SDL_UpdateTexture", s.UpdateTexture, missing);
  resolve(lib, "SDL_RenderClear", s.RenderClear, missing);
  resolve(lib, "SDL_RenderTexture", s.RenderTexture, missing);
  resolve(lib, "SDL_RenderPresent", s.RenderPresent, missing);
  resolve(lib, "SDL_PollEvent", s.PollEvent, missing);
  resolve(lib, "SDL_GetKeyboardState", s.GetKeyboardState, missing);
  resolve(lib, "SDL_SetWindowTitle", s.SetWindowTitle, missing);
  resolve(lib, "SDL_DestroyTexture", s.DestroyTexture, missing);
  resolve(lib, "SDL_DestroyRenderer", s.DestroyRenderer, missing);
  resolve(lib, "SDL_DestroyWindow", s.DestroyWindow, missing);
  if (!missing.empty()) {
    rth::set_error("the SDL3 library lacks: " + missing);
    dlclose(lib);
    delete p;
    return nullptr;
  }
  auto fail = [&](const char *what) {
    rth::set_error(std::string(what) + " failed: " + (s.GetError ? s.GetError() : "?"));
    rth_presenter_close(p);
    return static_cast<rth_presenter *>(nullptr);
  };
  if (!s.Init(kSDL_INIT_VIDEO)) // DynamicCamera.cpp:62
    return fail("SDL_Init");
  p->width = width;
  p->height = height;
  p->window = s.CreateWindow(title ? title : "Dynamic Camera", width, height, 0); // :66-67
  if (!p->window)
    return fail("SDL_CreateWindow");
  p->renderer = s.CreateRenderer(p->window, nullptr); // :68
  if (!p->renderer)
    return fail("SDL_CreateRenderer");
  p->texture = s.CreateTexture(p->renderer, kSDL_PIXELFORMAT_RGB24, kSDL_TEXTUREACCESS_STREAMING, width, height); // :69-71
  if (!p->texture)
    return fail("SDL_CreateTexture");
  return p;
}

// DynamicCamera::handle_events (:204-278): drains the event queue, then reads the held movement keys.
int rth_presenter_poll(rth_presenter *p, rth_input *out) {
  if (!p || !out)
    return 1;
  std::memset(out, 0, sizeof *out);
  const bool *state = p->sdl.GetKeyboardState(nullptr);
  SdlEvent e;
  std::memset(&e, 0, sizeof e);
  while (p->sdl.PollEvent(&e)) {
    if (e.type == kSDL_EVENT_QUIT)
      out->quit = 1;
    if (e.type == kSDL_EVENT_KEY_DOWN) {
      if (e.key.key == kSDLK_ESCAPE)
        out->quit = 1;
      if (e.key.key == kSDLK_EQUALS)
        out->spp_delta += 1;
      if (e.key.key == kSDLK_MINUS)
        out->spp_delta -= 1;
    }
  }
  if (state) {
    if (state[kSDL_SCANCODE_W])
      out->move_z += 1, out->moved = 1;
    if (state[kSDL_SCANCODE_S])
      out->move_z -= 1, out->moved = 1;
    if (state[kSDL_SCANCODE_A])
      out->move_x -= 1, out->moved = 1;
    if (state[kSDL_SCANCODE_D])
      out->move_x += 1, out->moved = 1;
  }
  return 0;
}

// update_texture's upload (:303-305) + the end of the render loop (:196-200).
int rth_presenter_present(rth_presenter *p, const uint8_t *rgb8, const char *status_line) {
  if (!p || !rgb8)
    return 1;
  const Sdl &s = p->sdl;
  bool ok = s.UpdateTexture(p->texture, nullptr, rgb8, p->width * 3);
  if (status_line)
    s.SetWindowTitle(p->window, status_line);
  ok = s.RenderClear(p->renderer) && ok;
  ok = s.RenderTexture(p->renderer, p->texture, nullptr, nullptr) && ok;
  ok = s.RenderPresent(p->renderer) && ok;
  if (!ok) {
    rth::set_error(std::string("SDL present failed: ") + s.GetError());
    return 1;
  }
  return 0;
}

void rth_presenter_close(rth_presenter *p) { // ~DynamicCamera (:25-37)
  if (!p)
    return;
  const Sdl &s = p->sdl;
  if (p->texture && s.DestroyTexture)
    s.DestroyTexture(p->texture);
  if (p->renderer && s.DestroyRenderer)
    s.DestroyRenderer(p->renderer);
  if (p->window && s.DestroyWindow)
    s.DestroyWindow(p->window);
  if (s.Quit)
    s.Quit();
  if (s.lib)
    dlclose(s.lib);
  delete p;
}

} // extern "C"
