/* TEST DOUBLE for libSDL3 (tests/test_presenter.py): implements the SDL 3.2 entry points host/presenter.cpp
 * resolves with dlopen, logs every call to $FAKE_SDL_LOG, plays a key script from $FAKE_SDL_KEYS (one character
 * per poll round: w/a/s/d = key held, '=' / '-' = key-down event, e = ESC key-down, q = window-close event,
 * anything else = nothing; a window-close event follows the end of the script) and writes the pixels of the
 * last SDL_UpdateTexture to $FAKE_SDL_FRAME.  Not part of the product. */
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint32_t type, reserved;
  uint64_t timestamp;
  uint32_t window_id, which, scancode, key;
  uint16_t mod, raw;
  uint8_t down, repeat;
} KeyboardEvent;
typedef union {
  uint32_t type;
  KeyboardEvent key;
  uint8_t padding[128];
} Event;

static bool g_keys[512];
static int g_round = -1, g_event_sent = 0, g_tex_w = 0, g_tex_h = 0;
static int g_window, g_renderer, g_texture;

static void logf_(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static void logf_(const char *fmt, ...) {
  const char *path = getenv("FAKE_SDL_LOG");
  if (!path)
    return;
  FILE *f = fopen(path, "a");
  if (!f)
    return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(f, fmt, ap);
  va_end(ap);
  fputc('\n', f);
  fclose(f);
}
static char script_char(void) {
  const char *s = getenv("FAKE_SDL_KEYS");
  if (!s || g_round < 0 || (size_t)g_round >= strlen(s))
    return 'q';
  return s[g_round];
}

bool SDL_Init(uint32_t flags) {
  logf_("SDL_Init 0x%x", flags);
  return getenv("FAKE_SDL_FAIL_INIT") == NULL;
}
void SDL_Quit(void) { logf_("SDL_Quit"); }
const char *SDL_GetError(void) { return "fake SDL error"; }
void *SDL_CreateWindow(const char *title, int w, int h, uint64_t flags) {
  logf_("SDL_CreateWindow \"%s\" %d %d %llu", title, w, h, (unsigned long long)flags);
  return &g_window;
}
void *SDL_CreateRenderer(void *window, const char *name) {
  logf_("SDL_CreateRenderer %d %s", window == &g_window, name ? name : "(null)");
  return &g_renderer;
}
void *SDL_CreateTexture(void *renderer, uint32_t format, int access, int w, int h) {
  logf_("SDL_CreateTexture %d 0x%x %d %d %d", renderer == &g_renderer, format, access, w, h);
  g_tex_w = w;
  g_tex_h = h;
  return &g_texture;
}
bool SDL_UpdateTexture(void *texture, const void *rect, const void *pixels, int pitch) {
  logf_("SDL_UpdateTexture %d rect=%d pitch=%d", texture == &g_texture, rect != NULL, pitch);
  const char *path = getenv("FAKE_SDL_FRAME");
  if (path) {
    FILE *f = fopen(path, "wb");
    if (f) {
      fwrite(pixels, 1, (size_t)pitch * (size_t)g_tex_h, f);
      fclose(f);
    }
  }
  return true;
}
bool SDL_RenderClear(void *renderer) {
  logf_("SDL_RenderClear %d", renderer == &g_renderer);
  return true;
}
bool SDL_RenderTexture(void *renderer, void *texture, const void *src, const void *dst) {
  logf_("SDL_RenderTexture %d %d %d %d", renderer == &g_renderer, texture == &g_texture, src != NULL, dst != NULL);
  return true;
}
bool SDL_RenderPresent(void *renderer) {
  logf_("SDL_RenderPresent %d", renderer == &g_renderer);
  return true;
}
const bool *SDL_GetKeyboardState(int *numkeys) { /* one call per poll round: advance the script */
  g_round++;
  g_event_sent = 0;
  memset(g_keys, 0, sizeof g_keys);
  switch (script_char()) {
  case 'a': g_keys[4] = true; break;
  case 'd': g_keys[7] = true; break;
  case 's': g_keys[22] = true; break;
  case 'w': g_keys[26] = true; break;
  default: break;
  }
  if (numkeys)
    *numkeys = 512;
  return g_keys;
}
bool SDL_PollEvent(Event *e) {
  if (g_event_sent)
    return false;
  g_event_sent = 1;
  char c = script_char();
  memset(e, 0, sizeof *e);
  if (c == 'q') {
    e->type = 0x100;
    return true;
  }
  if (c == '=' || c == '-' || c == 'e') {
    e->type = 0x300;
    e->key.key = c == 'e' ? 0x1B : (uint32_t)c;
    e->key.down = 1;
    return true;
  }
  return false;
}
bool SDL_SetWindowTitle(void *window, const char *title) {
  logf_("SDL_SetWindowTitle %d \"%s\"", window == &g_window, title);
  return true;
}
void SDL_DestroyTexture(void *t) { logf_("SDL_DestroyTexture %d", t == &g_texture); }
void SDL_DestroyRenderer(void *r) { logf_("SDL_DestroyRenderer %d", r == &g_renderer); }
void SDL_DestroyWindow(void *w) { logf_("SDL_DestroyWindow %d", w == &g_window); }
