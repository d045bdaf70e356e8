// api.cu -- context lifetime and error reporting of liba3d.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void a3d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int a3d_version(void) { return A3D_VERSION; }
extern "C" const char* a3d_last_error(void) { return g_err; }

extern "C" int a3d_create(int device, a3d_ctx** out) {
  if (!out) return A3D_EINVAL;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    a3d_set_error("a3d_create: no CUDA device (%s); liba3d has no CPU fallback", cudaGetErrorString(e));
    return A3D_ENODEV;
  }
  A3D_REQUIRE(device >= 0 && device < ndev, "a3d_create: device %d out of range (%d devices)", device, ndev);
  A3D_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  A3D_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    a3d_set_error("a3d_create: device %d is sm_%d%d; liba3d is built for sm_100a only", device, prop.major, prop.minor);
    return A3D_ENODEV;
  }
  a3d_ctx* c = new a3d_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  cudaDriverGetVersion(&c->driver_version);
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->fn_encode_tiled, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) c->fn_encode_tiled = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &c->fn_encode_im2col, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) c->fn_encode_im2col = nullptr;
  cudaGetLastError();
  if (!c->fn_encode_tiled || !c->fn_encode_im2col) {
    delete c;
    a3d_set_error("a3d_create: cuTensorMapEncode* driver entry points not found");
    return A3D_ENODEV;
  }
  *out = c;
  return 0;
}

extern "C" int a3d_destroy(a3d_ctx* ctx) {
  if (!ctx) return 0;
  a3d_comm_destroy(ctx);
  delete ctx;
  return 0;
}

extern "C" int a3d_sm_count(a3d_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" uint64_t a3d_launch_count(a3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
