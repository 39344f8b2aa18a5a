// api.cu — context, error reporting and engine dispatch of the C-ABI (include/m1b200.h)
#include "common.cuh"

static thread_local char g_err[1024] = "";

void m1_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* m1_last_error(void) { return g_err; }
extern "C" int m1_version(void) { return 100; }

extern "C" int m1_ctx_create(int device, m1_ctx** out) {
  M1_CHECK(out != nullptr, "m1_ctx_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  M1_CHECK(e == cudaSuccess && count > 0,
           "m1_ctx_create: no CUDA device (%s) - libm1b200 has no CPU fallback",
           cudaGetErrorString(e));
  M1_CHECK(device >= 0 && device < count, "m1_ctx_create: device %d out of range (%d)", device, count);
  cudaDeviceProp prop;
  M1_CUDA(cudaGetDeviceProperties(&prop, device));
  M1_CHECK(prop.major == 10, "m1_ctx_create: device %d is sm_%d%d, libm1b200 is built for sm_100a only",
           device, prop.major, prop.minor);
  M1_CUDA(cudaSetDevice(device));
  m1_ctx* c = new m1_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->launches = 0;
  c->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) c->encode_tiled = fn;
  c->scratch_bytes = 1 << 20;
  M1_CUDA(cudaMalloc(&c->scratch, c->scratch_bytes));
  c->partial_bytes = 160u << 20;     // deterministic-reduction partials / split-K partial tiles of the weight gradient
  M1_CUDA(cudaMalloc(&c->partial, c->partial_bytes));
  c->counter_bytes = 64u << 10;
  M1_CUDA(cudaMalloc(&c->counters, c->counter_bytes));
  M1_CUDA(cudaMemset(c->counters, 0, c->counter_bytes));
  *out = c;
  return 0;
}

extern "C" int m1_ctx_destroy(m1_ctx* ctx) {
  if (!ctx) return 0;
  cudaFree(ctx->scratch);
  cudaFree(ctx->partial);
  cudaFree(ctx->counters);
  delete ctx;
  return 0;
}

extern "C" int64_t m1_ctx_launch_count(m1_ctx* ctx, int reset) {
  if (!ctx) return -1;
  int64_t n = ctx->launches;
  if (reset) ctx->launches = 0;
  return n;
}

extern "C" int m1_conv3d(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                         const float* const* w, const void* w_packed, const float* const* bias,
                         void* const* outs, void* stream) {
  M1_CHECK(ctx && d && srcs && outs, "m1_conv3d: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int engine = d->engine;
  if (engine == M1_ENGINE_AUTO)
    engine = (w_packed != nullptr && m1_conv3d_tc_supported(d)) ? M1_ENGINE_TCGEN05 : M1_ENGINE_SIMT;
  if (engine == M1_ENGINE_TCGEN05) return m1_conv3d_tc(ctx, d, srcs, w_packed, bias, outs, st);
  M1_CHECK(w != nullptr, "m1_conv3d: SIMT engine needs the fp32 master weights");
  return m1_conv3d_simt(ctx, d, srcs, w, bias, outs, st);
}

extern "C" int m1_conv3d_plan_info(const m1_conv_desc* d, int which, int32_t* out) {
  if (!d || !out) return 0;
  if (which == 0) return m1_conv3d_tc_plan_info(d, out);
  if (which == 1) return m1_conv3d_halo_plan_info(d, out);
  if (which == 2) return m1_conv3d_wgrad_plan_info(d, out);
  return 0;
}
