/* SuperDiff-OR sampling (the loop of cifar/eval_utils.py:72-86 over get_joint_stoch_vf, cifar/dynamics.py:115-136) from plain C:
 * per timestep M x sd_scorenet_forward + one sd_step_vpsde.  No Python, no PyTorch in the process.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/native_sampler.c -o /tmp/native_sampler \
 *       -L super_diffusion_b200 -lsuperdiff_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/super_diffusion_b200
 *   /tmp/native_sampler x0.bin noise.bin out_x.bin out_logq.bin B n_steps dt modelA.bin modelB.bin [...]
 *
 * x0: fp32 [B,32,32,3]; noise: fp32 [n_steps][B,32,32,3] (the caller supplies the noise, as everywhere in this library);
 * models: weight blobs written by super_diffusion_b200.native.NativeScoreNet.save() for the reference's vpsde.py configuration. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "superdiff_b200.h"

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(3); } } while (0)
#define SD(call) do { int r_ = (call); if (r_ != SD_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, r_, sd_last_error()); exit(4); } } while (0)

static void* read_file(const char* path, size_t expect) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  void* p = malloc(expect);
  if (fread(p, 1, expect, f) != expect) { fprintf(stderr, "%s: expected %zu bytes\n", path, expect); exit(2); }
  fclose(f);
  return p;
}
static void write_file(const char* path, const void* p, size_t bytes) {
  FILE* f = fopen(path, "wb");
  fwrite(p, 1, bytes, f);
  fclose(f);
}

/* VP-SDE schedule of cifar/dynamics.py:101-110 (beta_0 = 0.1, beta_1 = 20; sigma(t) = t) */
static double dlog_alphadt(double t) { return -0.5 * 0.1 - 0.5 * t * (20.0 - 0.1); }
static double beta(double t) { return 1.0 + 0.5 * t * 0.1 + 0.5 * t * t * (20.0 - 0.1); }

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: %s x0.bin noise.bin out_x.bin out_logq.bin B n_steps dt model.bin [model.bin ...]\n", argv[0]); return 1; }
  const int B = atoi(argv[5]), n_steps = atoi(argv[6]), M = argc - 8, D = 32 * 32 * 3;
  const double dt = atof(argv[7]);
  if (M > SD_MAX_MODELS) { fprintf(stderr, "at most %d models\n", SD_MAX_MODELS); return 1; }
  sd_scorenet_desc d[SD_MAX_MODELS];
  size_t wbytes = 0, ws_bytes = 0;
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  for (int m = 0; m < M; ++m) {
    sd_scorenet_desc c = {0};
    c.image_size = 32; c.channels = 3; c.nf = 128; c.num_res_blocks = 2;
    c.n_levels = 4; c.ch_mult[0] = 1; c.ch_mult[1] = 2; c.ch_mult[2] = 2; c.ch_mult[3] = 2;
    c.n_attn_res = 2; c.attn_resolutions[0] = 16; c.attn_resolutions[1] = 8;
    SD(sd_scorenet_weights_bytes(&c, &wbytes));
    void* wh = read_file(argv[8 + m], wbytes);
    void* wd;
    CU(cudaMalloc(&wd, wbytes));
    CU(cudaMemcpy(wd, wh, wbytes, cudaMemcpyHostToDevice));
    free(wh);
    c.weights = wd; c.weights_bytes = wbytes;
    d[m] = c;
  }
  SD(sd_scorenet_workspace_bytes(&d[0], B, 0, &ws_bytes));
  const size_t n = (size_t)B * D;
  float* x0 = (float*)read_file(argv[1], n * 4);
  float* noise = (float*)read_file(argv[2], (size_t)n_steps * n * 4);
  float *x, *nz, *logq, *w, *tdev, *scores[SD_MAX_MODELS];
  void* ws;
  CU(cudaMalloc((void**)&x, n * 4)); CU(cudaMalloc((void**)&nz, n * 4)); CU(cudaMalloc(&ws, ws_bytes));
  CU(cudaMalloc((void**)&logq, (size_t)B * M * 4)); CU(cudaMalloc((void**)&w, (size_t)B * M * 4)); CU(cudaMalloc((void**)&tdev, 4));
  for (int m = 0; m < M; ++m) CU(cudaMalloc((void**)&scores[m], n * 4));
  CU(cudaMemcpy(x, x0, n * 4, cudaMemcpyHostToDevice));
  CU(cudaMemset(logq, 0, (size_t)B * M * 4));                     /* logq_0 = 0 (eval_utils.py:75) */

  double t = 1.0;                                                  /* Python-float t, t += -dt (eval_utils.py:76,85) */
  for (int i = 0; i < n_steps; ++i) {
    const float tf = (float)t;
    CU(cudaMemcpyAsync(tdev, &tf, 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(nz, noise + (size_t)i * n, n * 4, cudaMemcpyHostToDevice, st));
    for (int m = 0; m < M; ++m)
      SD(sd_scorenet_forward(&d[m], tdev, 0, x, NULL, B, scores[m], ws, ws_bytes, SD_PRECISION_BF16, st));
    SD(sd_step_vpsde(x, nz, (const float* const*)scores, M, B, D, (float)dlog_alphadt(t), (float)beta(t), (float)t, (float)dt, NULL, NULL,
                     SD_MODE_OR, SD_DLOGQ_CIFAR_MAXSUB, 1e6f, NULL, 0.f, logq, x /* in place */, w, st));
    CU(cudaStreamSynchronize(st));                                 /* tf / the pageable noise slice are reused next step */
    t += -dt;
  }
  float* xh = (float*)malloc(n * 4);
  float* lh = (float*)malloc((size_t)B * M * 4);
  CU(cudaMemcpy(xh, x, n * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(lh, logq, (size_t)B * M * 4, cudaMemcpyDeviceToHost));
  write_file(argv[3], xh, n * 4);
  write_file(argv[4], lh, (size_t)B * M * 4);
  printf("SuperDiff-OR: %d models, B = %d, %d steps, logq[0] = (%g, %g)\n", M, B, n_steps, (double)lh[0], (double)lh[M > 1 ? 1 : 0]);
  return 0;
}
