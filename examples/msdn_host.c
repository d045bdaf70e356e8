/* A host WITHOUT Python: one MSDN train step and one inference through liba3d's C-ABI (include/a3d.h).
 * What a maintainer of a C/C++/Go/Java/... driver binds instead of `models.msdn(inputs, targets)` + `session.run`
 * (src/ann3depth.py:126-127,143-145).  Build (see examples/Makefile):
 *   gcc msdn_host.c -I../include -I/usr/local/cuda/include -L../ann3depth_b200 -la3d -L/usr/local/cuda/lib64 -lcudart -lm
 * Random glorot-like weights in the packed arena, synthetic U[0,1) images: prints the two losses of the step (the random-init
 * sanity band of docs/documentation.md:391-394 is 1e2 .. 1e6) and the range of the predicted depth map. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include "a3d.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ < 0) { fprintf(stderr, "%s failed: %d (%s)\n", #x, rc_, a3d_last_error()); return 1; } } while (0)

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 2, H = 480, W = 640, DH = 55, DW = 73;
  a3d_ctx* ctx = NULL;
  CHECK(a3d_create(0, &ctx));
  size_t bytes = a3d_msdn_workspace_bytes(ctx, B, H, W, DH, DW, 1);
  void* ws = NULL;
  if (cudaMalloc(&ws, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", bytes); return 1; }
  a3d_msdn* net = NULL;
  CHECK(a3d_msdn_create(ctx, B, H, W, DH, DW, 1, ws, bytes, NULL, &net));
  CHECK(a3d_msdn_configure(net, 0.999f, 7));              /* a sane beta2 so that the step moves the weights */

  /* weights: uniform +-sqrt(6 / (fan_in + fan_out)) per segment, written straight into the packed arena */
  float *w_dev = NULL; size_t total = 0;
  CHECK(a3d_msdn_arena(net, &w_dev, NULL, NULL, NULL, NULL, &total));
  float* w = (float*)calloc(total, sizeof(float));
  unsigned seed = 1;
  int nseg = a3d_msdn_segment(net, -1, NULL, NULL, NULL, NULL);
  for (int i = 0; i < nseg; ++i) {
    const char* name; size_t off, numel; int shape[4];
    a3d_msdn_segment(net, i, &name, &off, &numel, shape);
    if (!strstr(name, "/kernel")) { if (strstr(name, "dense_1/bias") || strstr(name, "third/bias")) for (size_t k = 0; k < numel; ++k) w[off + k] = 1.0f; continue; }
    double fan = shape[1] ? (double)numel / shape[0] + (double)shape[0] * (shape[2] ? shape[1] * shape[2] : 1) : numel;
    float lim = (float)sqrt(6.0 / fan);
    for (size_t k = 0; k < numel; ++k) w[off + k] = (2.f * frand(&seed) - 1.f) * lim;
  }
  cudaMemcpy(w_dev, w, total * sizeof(float), cudaMemcpyHostToDevice);
  CHECK(a3d_msdn_sync_weights(net, NULL));

  /* one synthetic batch */
  size_t ni = (size_t)B * H * W * 3, nd = (size_t)B * DH * DW;
  float *im = (float*)malloc(ni * 4), *dp = (float*)malloc(nd * 4), *im_dev, *dp_dev, *loss_dev, *fine_dev;
  for (size_t k = 0; k < ni; ++k) im[k] = frand(&seed);
  for (size_t k = 0; k < nd; ++k) dp[k] = 0.05f + 0.95f * frand(&seed);
  cudaMalloc((void**)&im_dev, ni * 4); cudaMalloc((void**)&dp_dev, nd * 4); cudaMalloc((void**)&loss_dev, 8);
  cudaMalloc((void**)&fine_dev, (size_t)B * 55 * 74 * 4);
  cudaMemcpy(im_dev, im, ni * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dp_dev, dp, nd * 4, cudaMemcpyHostToDevice);

  int phase = a3d_msdn_step(net, im_dev, dp_dev, NULL /* device dropout RNG */, loss_dev, NULL);
  CHECK(phase);
  float loss[2];
  cudaMemcpy(loss, loss_dev, 8, cudaMemcpyDeviceToHost);
  CHECK(a3d_msdn_infer(net, im_dev, fine_dev, NULL, NULL));
  float* fine = (float*)malloc((size_t)B * 55 * 74 * 4);
  if (cudaMemcpy(fine, fine_dev, (size_t)B * 55 * 74 * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "CUDA error\n"); return 1; }
  float lo = fine[0], hi = fine[0];
  for (int k = 0; k < B * 55 * 74; ++k) { if (fine[k] < lo) lo = fine[k]; if (fine[k] > hi) hi = fine[k]; }
  printf("a3d_msdn_step: phase %d global_step %lld loss_coarse %.3f loss_fine %.3f | a3d_msdn_infer: depth map in [%.4f, %.4f] | launches %llu\n",
         phase, a3d_msdn_global_step(net), loss[0], loss[1], lo, hi, (unsigned long long)a3d_launch_count(ctx));
  int ok = phase == 1 && a3d_msdn_global_step(net) == 1 && isfinite(loss[0]) && isfinite(loss[1]) && loss[0] > 1e2f && loss[0] < 1e6f && isfinite(lo) && isfinite(hi);
  a3d_msdn_destroy(net);
  a3d_destroy(ctx);
  return ok ? 0 : 2;
}
