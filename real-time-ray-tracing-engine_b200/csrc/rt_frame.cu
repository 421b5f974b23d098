// rt_frame.cu — the displayed frame of a multi-GPU render: every rank tone-maps its own tiles and stores them
// straight into the frame owner's row-major RGB8 image through a peer-mapped pointer (NVLink / NVSwitch stores;
// CUDA IPC between processes, peer access inside one process), then raises an arrival flag in the owner's memory.
// No collective, no staging copy, no scatter pass: the tone map, the transfer and the tile placement are one
// kernel per rank (replaces the reference's full-buffer D2H + host tone map per frame, DynamicCamera.cpp:280-306,
// 519-554, and this library's earlier resolve -> NCCL gather -> scatter sequence).
//
// Protocol per use of a frame (ticket t = 1, 2, ... counted by every handle on its own):
//   remote rank: rt_film_present   ONE kernel: waits until the owner consumed ticket t-1 (first thread of every
//                                  block polls the owner's word), stores the tiles, the last block sets flags[rank] = t
//   owner      : rt_film_present   one kernel, ordered after the frame's previous download by a CUDA event
//                rt_frame_wait     one tiny kernel: the stream waits until all ranks' flags are >= t
//                rt_frame_download (copy to host memory on the frame's copy stream, then consumed = t)
//             or rt_frame_release  (consumed = t without a copy)
//   A frame with one rank needs none of the flags: present is a single launch, wait and release launch nothing.
// Waits are bounded (RT_FRAME_TIMEOUT_CYCLES): a rank that never shows up sets the frame's error word instead of
// hanging the GPU.
#include "rt_internal.h"

#include <algorithm>
#include <cstring>
#include <new>

#define RT_FRAME_FLAG_WORDS 128 // [0, 64) arrival tickets, [64] consumed ticket, [65] error
#define RT_FRAME_CONSUMED 64
#define RT_FRAME_ERROR 65
#define RT_FRAME_TIMEOUT_CYCLES (6000000000ll) // ~3 s at 2 GHz

struct rt_frame {
  rt_context *ctx = nullptr; // the context whose stream launches kernels on this handle
  int width = 0, height = 0, n_ranks = 1;
  uint8_t *rgb8 = nullptr;      // row-major frame, in the owner's memory (mapped into this process / device)
  uint32_t *flags = nullptr;    // RT_FRAME_FLAG_WORDS words behind the image
  void *base = nullptr;         // what cudaMalloc / cudaIpcOpenMemHandle returned
  bool owner = false, ipc = false;
  unsigned int *blocks_done = nullptr; // on ctx's device: block counter of the present kernel
  uint32_t ticket = 0;
  cudaStream_t copy_stream = nullptr;  // owner: downloads run here, next to the next frame's render
  cudaEvent_t ready = nullptr, copied = nullptr;
  bool download_pending = false; // owner: the copy stream may still be reading the frame
};

static size_t frame_image_bytes(int width, int height) {
  size_t b = (size_t)width * height * 3;
  return (b + 255) / 256 * 256;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Thread r waits until word[r] has reached `ticket` (wrap-around safe), or gives up after the timeout.
// One block.  release != nullptr: once every word has arrived, *release = ticket (the owner consumed the frame
// without copying it anywhere: rt_frame_wait_release).
__global__ void k_frame_wait(const uint32_t *words, int n_words, uint32_t ticket, uint32_t *error, uint32_t *release) {
  int r = threadIdx.x;
  if (r < n_words) {
    long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(words + r) - ticket) < 0) {
      if (clock64() - t0 > RT_FRAME_TIMEOUT_CYCLES) {
        atomicExch(error, 1u);
        break;
      }
      __nanosleep(128);
    }
  }
  if (release) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys(release, ticket);
    }
  }
}

// First thread of a block polls until *word has reached `value`; the block continues together.
__device__ __forceinline__ void block_wait_for(const uint32_t *word, uint32_t value, uint32_t *error) {
  if (word) {
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(word) - value) < 0) {
        if (clock64() - t0 > RT_FRAME_TIMEOUT_CYCLES) {
          atomicExch(error, 1u);
          break;
        }
        __nanosleep(128);
      }
    }
    __syncthreads();
  }
}

__global__ void k_frame_signal(uint32_t *word, uint32_t value) {
  __threadfence_system();
  st_release_sys(word, value);
}

// Byte offset of compact-film byte `b` (rank's owned tiles, tile order) in the row-major frame.
__device__ __forceinline__ size_t frame_offset(size_t b, size_t tile_bytes, int rank, int n_ranks) {
  size_t tile_local = b / tile_bytes;
  return (tile_local * (size_t)n_ranks + (size_t)rank) * tile_bytes + (b - tile_local * tile_bytes);
}

// The film value of owned pixel k.  When a multi-sample pass left its per-path radiance unsummed (PendingSum), this
// is where k_accumulate's in-order sum happens - same operands, same order, so the same bits - and the film is
// brought up to date on the way.
__device__ __forceinline__ float4 film_value(float4 *__restrict__ film, long long k, const PendingSum &pending) {
  float4 v = film[k];
  if (pending.n_samples) {
    const float4 *r = pending.radiance + film_index_to_path(pending, (uint32_t)k);
    for (int s = 0; s < pending.n_samples; s++) {
      float4 a = __ldg(r + (size_t)s * pending.n_owned);
      v.x += a.x;
      v.y += a.y;
      v.z += a.z;
    }
    film[k] = v;
  }
  return v;
}

// to_byte(scale * sum) (ColorUtility.hpp:11-26, FP64 like k_resolve_rgb8) of the film's owned pixels, staged per
// block in shared memory and stored as 16-byte pieces into the frame rows the tiles belong to.  1024 pixels per
// block iteration.  Requires (tile_rows * width * 3) % 16 == 0 (every piece stays inside one tile).
#define RT_PRESENT_THREADS 256
#define RT_PRESENT_PIXELS 1024
__global__ void __launch_bounds__(RT_PRESENT_THREADS)
    k_present_rgb8(float4 *__restrict__ film, long long n_owned, double scale, uint8_t *__restrict__ frame,
                   size_t tile_bytes, int rank, int n_ranks, unsigned int *blocks_done, uint32_t *flag, uint32_t ticket,
                   const uint32_t *consumed, uint32_t *error, const __grid_constant__ PendingSum pending) {
  __shared__ __align__(16) uint8_t stage[RT_PRESENT_PIXELS * 3];
  block_wait_for(consumed, ticket - 1, error);
  const long long n_chunks = (n_owned + RT_PRESENT_PIXELS - 1) / RT_PRESENT_PIXELS;
  const size_t total_bytes = (size_t)n_owned * 3;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
#pragma unroll
    for (int i = 0; i < RT_PRESENT_PIXELS / RT_PRESENT_THREADS; i++) {
      int local = i * RT_PRESENT_THREADS + threadIdx.x;
      long long k = chunk * RT_PRESENT_PIXELS + local;
      if (k < n_owned) {
        float4 v = film_value(film, k, pending);
        stage[local * 3 + 0] = to_byte_f64(scale * (double)v.x);
        stage[local * 3 + 1] = to_byte_f64(scale * (double)v.y);
        stage[local * 3 + 2] = to_byte_f64(scale * (double)v.z);
      }
    }
    __syncthreads();
    if (threadIdx.x < RT_PRESENT_PIXELS * 3 / 16) {
      size_t b = (size_t)chunk * (RT_PRESENT_PIXELS * 3) + (size_t)threadIdx.x * 16;
      if (b < total_bytes) // total_bytes is a multiple of 16 here
        *reinterpret_cast<uint4 *>(frame + frame_offset(b, tile_bytes, rank, n_ranks)) =
            *reinterpret_cast<const uint4 *>(stage + threadIdx.x * 16);
    }
    __syncthreads();
  }
  if (flag) { // the last block to finish raises the rank's arrival flag
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int done = atomicAdd(blocks_done, 1u);
      if (done == gridDim.x - 1) {
        *blocks_done = 0;
        __threadfence_system();
        st_release_sys(flag, ticket);
      }
    }
  }
}

// Any geometry: one pixel per thread, three byte stores.
__global__ void k_present_rgb8_any(float4 *__restrict__ film, long long n_owned, double scale,
                                   uint8_t *__restrict__ frame, int width, int tile_rows, int rank, int n_ranks,
                                   unsigned int *blocks_done, uint32_t *flag, uint32_t ticket, const uint32_t *consumed,
                                   uint32_t *error, const __grid_constant__ PendingSum pending) {
  block_wait_for(consumed, ticket - 1, error);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_owned; k += stride) {
    long long local_row = k / width;
    int col = (int)(k - local_row * width);
    long long tile_local = local_row / tile_rows;
    long long row = (tile_local * n_ranks + rank) * tile_rows + (local_row - tile_local * tile_rows);
    float4 v = film_value(film, k, pending);
    uint8_t *out = frame + ((size_t)row * width + col) * 3;
    out[0] = to_byte_f64(scale * (double)v.x);
    out[1] = to_byte_f64(scale * (double)v.y);
    out[2] = to_byte_f64(scale * (double)v.z);
  }
  if (flag) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int done = atomicAdd(blocks_done, 1u);
      if (done == gridDim.x - 1) {
        *blocks_done = 0;
        __threadfence_system();
        st_release_sys(flag, ticket);
      }
    }
  }
}

// Owned pixels of a film (compact tile order) -> RGB8.  (rank, n_ranks) = the film's own: every tile lands in its
// rows of a row-major frame; (0, 1): the output keeps the compact order (rt_film_resolve_rgb8_device).
// flag == nullptr: no arrival flag.  consumed != nullptr: every block first waits until *consumed >= ticket - 1.
void launch_present_rgb8(const rt_context *ctx, cudaStream_t stream, const float4 *accum, int64_t n_owned, int width,
                         int tile_rows, int rank, int n_ranks, double scale, uint8_t *frame, unsigned int *blocks_done,
                         uint32_t *flag, uint32_t ticket, const uint32_t *consumed, uint32_t *error, float4 *accum_rw,
                         const PendingSum &pending) {
  // a pending sum is folded into the film by the kernel, which then needs the film writable
  float4 *film = accum_rw ? accum_rw : const_cast<float4 *>(accum);
  const PendingSum sum = accum_rw ? pending : PendingSum();
  const size_t tile_bytes = (size_t)tile_rows * width * 3;
  if (n_owned == 0) {
    if (flag)
      k_frame_signal<<<1, 1, 0, stream>>>(flag, ticket);
    return;
  }
  const bool whole_pieces = n_ranks == 1 ? ((size_t)n_owned * 3) % 16 == 0 : tile_bytes % 16 == 0;
  if (whole_pieces && ((uintptr_t)frame & 15) == 0) {
    long long chunks = (n_owned + RT_PRESENT_PIXELS - 1) / RT_PRESENT_PIXELS;
    int blocks = (int)std::min<long long>(chunks, (long long)ctx->sm_count * 8);
    k_present_rgb8<<<blocks, RT_PRESENT_THREADS, 0, stream>>>(film, n_owned, scale, frame,
                                                              n_ranks == 1 ? (size_t)n_owned * 3 : tile_bytes, rank, n_ranks,
                                                              blocks_done, flag, ticket, consumed, error, sum);
  } else {
    int blocks = (int)std::min<long long>((n_owned + 255) / 256, (long long)ctx->sm_count * 8);
    k_present_rgb8_any<<<blocks, 256, 0, stream>>>(film, n_owned, scale, frame, width, tile_rows, rank, n_ranks, blocks_done,
                                                   flag, ticket, consumed, error, sum);
  }
}

namespace {
int frame_invalid(const char *msg) {
  rt_set_error(msg);
  return RT_ERR_INVALID;
}

int frame_finish_handle(rt_frame *f) {
  RT_CUDA(cudaMalloc((void **)&f->blocks_done, sizeof(unsigned int)));
  RT_CUDA(cudaMemsetAsync(f->blocks_done, 0, sizeof(unsigned int), f->ctx->stream));
  if (f->owner) {
    RT_CUDA(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
    RT_CUDA(cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming));
    RT_CUDA(cudaEventCreateWithFlags(&f->copied, cudaEventDisableTiming));
  }
  RT_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return RT_OK;
}
} // namespace

extern "C" {

int rt_frame_create(rt_context *ctx, int width, int height, int n_ranks, rt_frame **out) {
  if (!ctx || !out || width < 1 || height < 1 || n_ranks < 1 || n_ranks > 64 || (int64_t)width * height > (int64_t)1 << 30)
    return frame_invalid("rt_frame_create: bad argument (1..64 ranks)");
  *out = nullptr;
  RT_CUDA(cudaSetDevice(ctx->device));
  rt_frame *f = new (std::nothrow) rt_frame();
  if (!f)
    return frame_invalid("out of host memory");
  f->ctx = ctx;
  f->width = width;
  f->height = height;
  f->n_ranks = n_ranks;
  f->owner = true;
  const size_t image = frame_image_bytes(width, height), total = image + RT_FRAME_FLAG_WORDS * sizeof(uint32_t);
  cudaError_t e = cudaMalloc(&f->base, total);
  if (e != cudaSuccess) {
    delete f;
    return rt_cuda_fail(e, "cudaMalloc (frame)");
  }
  f->rgb8 = static_cast<uint8_t *>(f->base);
  f->flags = reinterpret_cast<uint32_t *>(f->rgb8 + image);
  e = cudaMemsetAsync(f->base, 0, total, ctx->stream);
  int st = e == cudaSuccess ? frame_finish_handle(f) : rt_cuda_fail(e, "cudaMemsetAsync (frame)");
  if (st != RT_OK) {
    rt_frame_destroy(f);
    return st;
  }
  *out = f;
  return RT_OK;
}

int rt_frame_export(rt_frame *frame, unsigned char handle[64]) {
  if (!frame || !handle || !frame->owner)
    return frame_invalid("rt_frame_export: needs the owner's frame");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  RT_CUDA(cudaSetDevice(frame->ctx->device));
  cudaIpcMemHandle_t h;
  RT_CUDA(cudaIpcGetMemHandle(&h, frame->base));
  std::memcpy(handle, &h, 64);
  return RT_OK;
}

int rt_frame_open(rt_context *ctx, const unsigned char handle[64], int width, int height, int n_ranks, rt_frame **out) {
  if (!ctx || !handle || !out || width < 1 || height < 1 || n_ranks < 1 || n_ranks > 64)
    return frame_invalid("rt_frame_open: bad argument");
  *out = nullptr;
  RT_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  void *base = nullptr;
  RT_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  rt_frame *f = new (std::nothrow) rt_frame();
  if (!f) {
    cudaIpcCloseMemHandle(base);
    return frame_invalid("out of host memory");
  }
  f->ctx = ctx;
  f->width = width;
  f->height = height;
  f->n_ranks = n_ranks;
  f->ipc = true;
  f->base = base;
  f->rgb8 = static_cast<uint8_t *>(base);
  f->flags = reinterpret_cast<uint32_t *>(f->rgb8 + frame_image_bytes(width, height));
  int st = frame_finish_handle(f);
  if (st != RT_OK) {
    rt_frame_destroy(f);
    return st;
  }
  *out = f;
  return RT_OK;
}

int rt_frame_attach(rt_context *ctx, rt_frame *owner_frame, rt_frame **out) {
  if (!ctx || !owner_frame || !out || !owner_frame->owner)
    return frame_invalid("rt_frame_attach: needs a context and the owner's frame");
  *out = nullptr;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (ctx->device != owner_frame->ctx->device) {
    int can = 0;
    RT_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, owner_frame->ctx->device));
    if (!can) {
      rt_set_error("rt_frame_attach: the device has no peer access to the frame owner's device");
      return RT_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(owner_frame->ctx->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled)
      cudaGetLastError();
    else if (e != cudaSuccess)
      return rt_cuda_fail(e, "cudaDeviceEnablePeerAccess");
  }
  rt_frame *f = new (std::nothrow) rt_frame();
  if (!f)
    return frame_invalid("out of host memory");
  f->ctx = ctx;
  f->width = owner_frame->width;
  f->height = owner_frame->height;
  f->n_ranks = owner_frame->n_ranks;
  f->rgb8 = owner_frame->rgb8;
  f->flags = owner_frame->flags;
  int st = frame_finish_handle(f);
  if (st != RT_OK) {
    rt_frame_destroy(f);
    return st;
  }
  *out = f;
  return RT_OK;
}

void rt_frame_destroy(rt_frame *frame) {
  if (!frame)
    return;
  cudaSetDevice(frame->ctx->device);
  cudaStreamSynchronize(frame->ctx->stream);
  if (frame->copy_stream) {
    cudaStreamSynchronize(frame->copy_stream);
    cudaStreamDestroy(frame->copy_stream);
  }
  if (frame->ready)
    cudaEventDestroy(frame->ready);
  if (frame->copied)
    cudaEventDestroy(frame->copied);
  cudaFree(frame->blocks_done);
  if (frame->ipc)
    cudaIpcCloseMemHandle(frame->base);
  else if (frame->owner)
    cudaFree(frame->base);
  delete frame;
}

uint64_t rt_frame_device_ptr(rt_frame *frame) { return frame ? (uint64_t)(uintptr_t)frame->rgb8 : 0; }

int rt_film_present(rt_film *film, double scale, rt_frame *frame) {
  if (!film || !frame)
    return frame_invalid("rt_film_present: null argument");
  if (film->ctx != frame->ctx)
    return frame_invalid("rt_film_present: the film and the frame handle belong to different contexts");
  if (film->map.width != frame->width || film->map.height != frame->height || film->map.n_ranks != frame->n_ranks)
    return frame_invalid("rt_film_present: film and frame geometry differ");
  rt_context *ctx = film->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  const uint32_t ticket = ++frame->ticket;
  const bool alone = frame->n_ranks == 1;
  const uint32_t *consumed = nullptr;
  if (frame->owner) {
    // the frame's previous content may still be on its way to host memory: order behind that copy (no kernel)
    if (frame->download_pending) {
      RT_CUDA(cudaStreamWaitEvent(ctx->stream, frame->copied, 0));
      frame->download_pending = false;
    }
  } else if (ticket > 1) {
    consumed = frame->flags + RT_FRAME_CONSUMED; // polled by the present kernel itself
  }
  uint32_t *flag = frame->flags + film->map.rank; // raised by the kernel's last block
  // a multi-sample pass that deferred its in-order sum: the present kernel does it while it tone-maps
  const PendingSum pending = film->pending;
  film->pending = PendingSum();
  if (ctx->pending_film == film)
    ctx->pending_film = nullptr;
  launch_present_rgb8(ctx, ctx->stream, film->accum, film->n_owned, film->map.width, film->map.tile_rows, film->map.rank,
                      film->map.n_ranks, scale, frame->rgb8, frame->blocks_done, alone ? nullptr : flag, ticket, consumed,
                      frame->flags + RT_FRAME_ERROR, film->accum, pending);
  ctx->counters.kernel_launches += 1;
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}

int rt_frame_wait(rt_frame *frame) {
  if (!frame || !frame->owner)
    return frame_invalid("rt_frame_wait: needs the owner's frame");
  rt_context *ctx = frame->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (frame->n_ranks > 1) {
    k_frame_wait<<<1, 64, 0, ctx->stream>>>(frame->flags, frame->n_ranks, frame->ticket, frame->flags + RT_FRAME_ERROR, nullptr);
    ctx->counters.kernel_launches += 1;
    RT_CUDA(cudaGetLastError());
  }
  return RT_OK;
}

int rt_frame_wait_release(rt_frame *frame) {
  if (!frame || !frame->owner)
    return frame_invalid("rt_frame_wait_release: needs the owner's frame");
  rt_context *ctx = frame->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (frame->n_ranks > 1) {
    k_frame_wait<<<1, 64, 0, ctx->stream>>>(frame->flags, frame->n_ranks, frame->ticket, frame->flags + RT_FRAME_ERROR,
                                            frame->flags + RT_FRAME_CONSUMED);
    ctx->counters.kernel_launches += 1;
    RT_CUDA(cudaGetLastError());
  }
  return RT_OK;
}

int rt_frame_release(rt_frame *frame) {
  if (!frame || !frame->owner)
    return frame_invalid("rt_frame_release: needs the owner's frame");
  rt_context *ctx = frame->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (frame->n_ranks > 1) {
    k_frame_signal<<<1, 1, 0, ctx->stream>>>(frame->flags + RT_FRAME_CONSUMED, frame->ticket);
    ctx->counters.kernel_launches += 1;
    RT_CUDA(cudaGetLastError());
  }
  return RT_OK;
}

int rt_frame_download(rt_frame *frame, uint8_t *host_rgb8) {
  if (!frame || !frame->owner || !host_rgb8)
    return frame_invalid("rt_frame_download: needs the owner's frame and a host buffer");
  rt_context *ctx = frame->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  RT_CUDA(cudaEventRecord(frame->ready, ctx->stream)); // everything queued so far: render, present, wait
  RT_CUDA(cudaStreamWaitEvent(frame->copy_stream, frame->ready, 0));
  RT_CUDA(cudaMemcpyAsync(host_rgb8, frame->rgb8, (size_t)frame->width * frame->height * 3, cudaMemcpyDeviceToHost,
                          frame->copy_stream));
  if (frame->n_ranks > 1) {
    k_frame_signal<<<1, 1, 0, frame->copy_stream>>>(frame->flags + RT_FRAME_CONSUMED, frame->ticket);
    ctx->counters.kernel_launches += 1;
    RT_CUDA(cudaGetLastError());
  }
  RT_CUDA(cudaEventRecord(frame->copied, frame->copy_stream));
  frame->download_pending = true;
  return RT_OK;
}

int rt_frame_download_wait(rt_frame *frame) {
  if (!frame || !frame->owner)
    return frame_invalid("rt_frame_download_wait: needs the owner's frame");
  RT_CUDA(cudaSetDevice(frame->ctx->device));
  RT_CUDA(cudaStreamSynchronize(frame->copy_stream));
  uint32_t err = 0;
  RT_CUDA(cudaMemcpy(&err, frame->flags + RT_FRAME_ERROR, sizeof err, cudaMemcpyDeviceToHost));
  if (err) {
    rt_set_error("frame: a rank did not present its tiles in time");
    return RT_ERR_CUDA;
  }
  return RT_OK;
}

int rt_host_alloc(size_t bytes, void **out) {
  if (!out)
    return frame_invalid("rt_host_alloc: null output");
  *out = nullptr;
  RT_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return RT_OK;
}

void rt_host_free(void *p) {
  if (p)
    cudaFreeHost(p);
}

int rt_frame_error(rt_frame *frame) {
  if (!frame)
    return -1;
  if (cudaSetDevice(frame->ctx->device) != cudaSuccess || cudaStreamSynchronize(frame->ctx->stream) != cudaSuccess)
    return -1;
  uint32_t err = 0;
  if (cudaMemcpy(&err, frame->flags + RT_FRAME_ERROR, sizeof err, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return (int)err;
}

} // extern "C"
