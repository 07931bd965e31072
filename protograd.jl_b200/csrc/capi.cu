// capi.cu — lifecycle, memory, stream/graph plumbing of the C ABI (include/protograd_b200.h).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void pg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int pg_ws_ensure(pg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return PG_OK;
  if (ctx->capturing) {
    pg_set_error("workspace of %zu bytes needed during graph capture (have %zu): call pg_reserve_workspace first", bytes, ctx->ws_bytes);
    return PG_ERR_NOMEM;
  }
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->ws) PG_CHECK_CUDA(cudaFree(ctx->ws));
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  size_t want = bytes + (bytes >> 2);
  PG_CHECK_CUDA(cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return PG_OK;
}

extern "C" {

const char* pg_last_error(void) { return g_err; }
int pg_version(void) { return 100; }

int pg_device_count(int* n) {
  PG_REQUIRE(n != nullptr, "null out pointer");
  *n = 0;
  PG_CHECK_CUDA(cudaGetDeviceCount(n));
  return PG_OK;
}

int pg_init(int device, pg_ctx** out) {
  PG_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  int n = 0;
  PG_CHECK_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) {
    pg_set_error("no CUDA device: this library has no CPU fallback");
    return PG_ERR_CUDA;
  }
  PG_REQUIRE(device >= 0 && device < n, "device index out of range");
  PG_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    pg_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return PG_ERR_UNSUPPORTED;
  }
  pg_ctx* c = new pg_ctx();
  c->device = device;
  c->num_sms = c->real_sms = prop.multiProcessorCount;
  // a failure half-way leaves no handle behind: everything created so far is released again
  auto setup = [&]() -> int {
    PG_CHECK_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    PG_CHECK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    PG_CHECK_CUDA(cudaMalloc((void**)&c->counters, 256 * sizeof(unsigned int)));
    PG_CHECK_CUDA(cudaMemset(c->counters, 0, 256 * sizeof(unsigned int)));
    PG_CHECK_CUDA(cudaMalloc((void**)&c->result_slot, 16 * sizeof(double)));
    PG_CHECK_CUDA(cudaMallocHost((void**)&c->result_host, 16 * sizeof(double)));
    return pg_ws_ensure(c, (size_t)32 << 20);
  };
  const int rc = setup();
  if (rc != PG_OK) {
    pg_destroy(c);          // (keeps the error message: pg_destroy sets none)
    return rc;
  }
  *out = c;
  return PG_OK;
}

int pg_destroy(pg_ctx* ctx) {
  if (!ctx) return PG_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->switch_event) cudaEventDestroy(ctx->switch_event);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->aux) cudaFree(ctx->aux);
  if (ctx->counters) cudaFree(ctx->counters);
  if (ctx->result_slot) cudaFree(ctx->result_slot);
  if (ctx->result_host) cudaFreeHost(ctx->result_host);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return PG_OK;
}

int pg_set_stream(pg_ctx* ctx, void* s) {
  PG_CHECK_CTX(ctx);
  cudaStream_t next = s ? (cudaStream_t)s : ctx->own_stream;
  if (next != ctx->stream && !ctx->capturing) {
    // The context's scratch (workspace, padded operand copies, reduction and tile-scheduler counters) assumes
    // ONE queue of work: whatever was enqueued on the previous stream must finish before anything enqueued on
    // the new one starts.
    if (!ctx->switch_event) PG_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->switch_event, cudaEventDisableTiming));
    // (a borrowed stream may have been destroyed by its owner since: then there is nothing to wait for)
    if (cudaEventRecord(ctx->switch_event, ctx->stream) == cudaSuccess) {
      PG_CHECK_CUDA(cudaStreamWaitEvent(next, ctx->switch_event, 0));
    } else {
      cudaGetLastError();
    }
  }
  ctx->stream = next;
  return PG_OK;
}
int pg_get_stream(pg_ctx* ctx, void** s) {
  PG_CHECK_CTX_NF(ctx);
  PG_REQUIRE(s != nullptr, "null out pointer");
  *s = (void*)ctx->stream;
  return PG_OK;
}
int pg_sync(pg_ctx* ctx) {
  PG_CHECK_CTX(ctx);
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return PG_OK;
}
int pg_launch_count(pg_ctx* ctx, int64_t* n) {
  PG_CHECK_CTX_NF(ctx);
  PG_REQUIRE(n != nullptr, "null out pointer");
  *n = ctx->launches;
  return PG_OK;
}
int pg_set_sm_limit(pg_ctx* ctx, int n) {
  PG_CHECK_CTX(ctx);
  if (n < 0) n = ctx->real_sms + n;            // negative: leave |n| SMs free
  ctx->num_sms = (n <= 0 || n > ctx->real_sms) ? ctx->real_sms : n;
  return PG_OK;
}
int pg_reserve_workspace(pg_ctx* ctx, size_t bytes) {
  PG_CHECK_CTX(ctx);
  return pg_ws_ensure(ctx, bytes);
}

int pg_malloc(pg_ctx* ctx, size_t bytes, void** dptr) {
  PG_CHECK_CTX_NF(ctx);
  PG_REQUIRE(dptr != nullptr, "null out pointer");
  PG_CHECK_CUDA(cudaSetDevice(ctx->device));
  PG_CHECK_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
  return PG_OK;
}
int pg_free(pg_ctx* ctx, void* dptr) {
  PG_CHECK_CTX_NF(ctx);
  if (dptr) PG_CHECK_CUDA(cudaFree(dptr));
  return PG_OK;
}
int pg_memset0(pg_ctx* ctx, void* dptr, size_t bytes) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, dptr, bytes)) pg_flush_deferred(ctx);
  PG_CHECK_CUDA(cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
  return PG_OK;
}
int pg_copy_d2d(pg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, dst, bytes) || pg_deferred_overlaps(ctx, src, bytes)) pg_flush_deferred(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return PG_OK;
}
int pg_upload(pg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, dst, bytes)) pg_flush_deferred(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PG_OK;
}
int pg_download_async(pg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, src, bytes)) pg_flush_deferred(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return PG_OK;
}
int pg_download(pg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  int rc = pg_download_async(ctx, dst, src, bytes);
  if (rc != PG_OK) return rc;
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return PG_OK;
}
int pg_host_alloc(size_t bytes, void** p) {
  PG_REQUIRE(p != nullptr, "null out pointer");
  PG_CHECK_CUDA(cudaMallocHost(p, bytes ? bytes : 16));
  return PG_OK;
}
int pg_host_free(void* p) {
  if (p) PG_CHECK_CUDA(cudaFreeHost(p));
  return PG_OK;
}

static cudaStream_t pick_stream(pg_ctx* ctx, int which) { return which == 1 ? ctx->copy_stream : ctx->stream; }

int pg_event_create(pg_ctx* ctx, void** event) {
  PG_CHECK_CTX_NF(ctx);
  PG_REQUIRE(event != nullptr, "null out pointer");
  cudaEvent_t e;
  PG_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = (void*)e;
  return PG_OK;
}
int pg_event_destroy(pg_ctx* ctx, void* event) {
  PG_CHECK_CTX_NF(ctx);
  if (event) PG_CHECK_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return PG_OK;
}
int pg_event_record(pg_ctx* ctx, void* event, int which_stream) {
  PG_CHECK_CTX_NF(ctx);
  PG_CHECK_CUDA(cudaEventRecord((cudaEvent_t)event, pick_stream(ctx, which_stream)));
  return PG_OK;
}
int pg_stream_wait_event(pg_ctx* ctx, int which_stream, void* event) {
  PG_CHECK_CTX_NF(ctx);
  PG_CHECK_CUDA(cudaStreamWaitEvent(pick_stream(ctx, which_stream), (cudaEvent_t)event, 0));
  return PG_OK;
}
int pg_event_sync(pg_ctx* ctx, void* event) {
  PG_CHECK_CTX_NF(ctx);
  PG_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return PG_OK;
}
int pg_upload_on(pg_ctx* ctx, int which_stream, void* dst, const void* src, size_t bytes) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, dst, bytes)) pg_flush_deferred(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pick_stream(ctx, which_stream)));
  return PG_OK;
}

int pg_graph_begin(pg_ctx* ctx) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(!ctx->capturing, "already capturing");
  PG_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  ctx->capturing = true;
  return PG_OK;
}
int pg_graph_end(pg_ctx* ctx, void** graph_exec) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(ctx->capturing, "not capturing");
  PG_REQUIRE(graph_exec != nullptr, "null out pointer");
  ctx->capturing = false;
  cudaGraph_t g = nullptr;
  PG_CHECK_CUDA(cudaStreamEndCapture(ctx->stream, &g));
  cudaGraphExec_t ge = nullptr;
  PG_CHECK_CUDA(cudaGraphInstantiate(&ge, g, 0));
  PG_CHECK_CUDA(cudaGraphDestroy(g));
  *graph_exec = (void*)ge;
  return PG_OK;
}
int pg_graph_launch(pg_ctx* ctx, void* graph_exec) {
  PG_CHECK_CTX(ctx);
  PG_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
  return PG_OK;
}
int pg_graph_destroy(pg_ctx* ctx, void* graph_exec) {
  PG_CHECK_CTX(ctx);
  if (graph_exec) PG_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return PG_OK;
}

int pg_loss_read(pg_ctx* ctx, int dtype, const void* loss_dev, double* result) {
  PG_CHECK_CTX_NF(ctx);
  if (pg_deferred_overlaps(ctx, loss_dev, 8)) pg_flush_deferred(ctx);
  PG_REQUIRE(result != nullptr && loss_dev != nullptr, "null pointer");
  PG_REQUIRE(dtype == PG_F32 || dtype == PG_F64 || dtype == PG_F16 || dtype == PG_BF16, "pg_loss_read: dtype must be a floating-point type");
  const size_t sz = dtype == PG_F64 ? 8 : (dtype == PG_F32 ? 4 : 2);
  PG_CHECK_CUDA(cudaMemcpyAsync(ctx->result_host, loss_dev, sz, cudaMemcpyDeviceToHost, ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (dtype == PG_F64) {
    *result = *(double*)ctx->result_host;
  } else if (dtype == PG_F32) {
    *result = (double)*(float*)ctx->result_host;
  } else {
    const uint16_t h = *(uint16_t*)ctx->result_host;
    uint32_t bits;
    if (dtype == PG_BF16) {
      bits = (uint32_t)h << 16;
    } else {                                  // IEEE binary16 -> binary32 (exact)
      const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
      if (e == 0) {
        if (m == 0) bits = sign;
        else {                                // subnormal: renormalise
          int sh = 0; uint32_t mm = m;
          while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
          bits = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3ffu) << 13);
        }
      } else if (e == 31) bits = sign | 0x7f800000u | (m << 13);
      else bits = sign | ((e + 112) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    *result = (double)f;
  }
  return PG_OK;
}

}  // extern "C"
