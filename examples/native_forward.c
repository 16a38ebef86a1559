/* The score network from plain C: no Python, no PyTorch - libsuperdiff_b200.so, the CUDA runtime, and two files written by
 * super_diffusion_b200.native.NativeScoreNet.save() (a weight blob) / any tool (a raw fp32 NHWC input).
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/native_forward.c -o /tmp/native_forward \
 *       -L super_diffusion_b200 -lsuperdiff_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/super_diffusion_b200
 *   /tmp/native_forward model.bin x.bin out.bin <B> <t> [conditioned: labels.bin]
 *
 * Configuration = the reference's cifar/configs/sm/cifar/vpsde.py (nf 128, ch_mult (1,2,2,2), 2 blocks per level, attention
 * at 16x16 and 8x8); this is the model_fn(t, x, y) seam of cifar/models/utils.py:86-96 as three C calls. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "superdiff_b200.h"

static void* read_file(const char* path, size_t* bytes) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  *bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc(*bytes);
  if (fread(p, 1, *bytes, f) != *bytes) { fprintf(stderr, "short read on %s\n", path); exit(2); }
  fclose(f);
  return p;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(3); } } while (0)
#define SD(call) do { int r_ = (call); if (r_ != SD_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, r_, sd_last_error()); exit(4); } } while (0)

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s model.bin x.bin out.bin B t [labels.bin]\n", argv[0]); return 1; }
  const int B = atoi(argv[4]);
  const float t = (float)atof(argv[5]);
  sd_scorenet_desc d = {0};
  d.image_size = 32; d.channels = 3; d.nf = 128; d.num_res_blocks = 2;
  d.n_levels = 4; d.ch_mult[0] = 1; d.ch_mult[1] = 2; d.ch_mult[2] = 2; d.ch_mult[3] = 2;
  d.n_attn_res = 2; d.attn_resolutions[0] = 16; d.attn_resolutions[1] = 8;
  d.conditioned = argc > 6; d.num_classes = 10;

  size_t wbytes = 0, need = 0, xbytes = 0, ws_bytes = 0;
  void* wh = read_file(argv[1], &wbytes);
  SD(sd_scorenet_weights_bytes(&d, &need));
  if (need != wbytes) { fprintf(stderr, "%s holds %zu bytes, this configuration needs %zu\n", argv[1], wbytes, need); return 2; }
  float* xh = (float*)read_file(argv[2], &xbytes);
  const size_t n = (size_t)B * 32 * 32 * 3;
  if (xbytes != n * 4) { fprintf(stderr, "%s: expected %zu bytes\n", argv[2], n * 4); return 2; }
  SD(sd_scorenet_workspace_bytes(&d, B, 0, &ws_bytes));

  void *wd, *ws, *xd, *od, *td, *yd = NULL;
  CU(cudaMalloc(&wd, wbytes)); CU(cudaMalloc(&ws, ws_bytes)); CU(cudaMalloc(&xd, n * 4)); CU(cudaMalloc(&od, n * 4));
  CU(cudaMalloc(&td, 4));
  CU(cudaMemcpy(wd, wh, wbytes, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(xd, xh, n * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(td, &t, 4, cudaMemcpyHostToDevice));
  if (d.conditioned) {
    size_t yb = 0;
    void* yh = read_file(argv[6], &yb);
    if (yb != (size_t)B * 4) { fprintf(stderr, "%s: expected %d int32 labels\n", argv[6], B); return 2; }
    CU(cudaMalloc(&yd, yb));
    CU(cudaMemcpy(yd, yh, yb, cudaMemcpyHostToDevice));
  }
  d.weights = wd; d.weights_bytes = wbytes;
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  SD(sd_scorenet_forward(&d, (const float*)td, 0, (const float*)xd, (const int*)yd, B, (float*)od, ws, ws_bytes, SD_PRECISION_BF16, st));
  CU(cudaStreamSynchronize(st));
  float* oh = (float*)malloc(n * 4);
  CU(cudaMemcpy(oh, od, n * 4, cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[3], "wb");
  fwrite(oh, 4, n, f);
  fclose(f);
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += (double)oh[i] * oh[i];
  printf("score-net forward: B = %d, t = %g, workspace %.1f MB, |out|^2 = %.6e\n", B, (double)t, ws_bytes / 1e6, s);
  return 0;
}
