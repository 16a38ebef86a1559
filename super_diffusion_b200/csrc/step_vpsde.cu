// Host dispatch + C ABI of the fused VP-SDE SuperDiff step (kernels: step_vpsde_kernel.cuh,
// instantiated per model count in step_vpsde_m*.cu so the build parallelises).
#include "common.cuh"
#include "../../include/superdiff_b200.h"

#include "step_vpsde_params.cuh"
#include <cstdlib>

namespace sdb {

template <int M> cudaError_t launch_m(const StepParams& p, int threads, int nv, int cluster, int vec, cudaStream_t st);
template <int M> cudaError_t launch_small(const StepParams& p, cudaStream_t st);
template <int M> cudaError_t launch_and_stream(const StepParams& p, int threads, int nv, int vec, cudaStream_t st);
template <int M> cudaError_t launch_and_smem(const StepParams& p, int threads, cudaStream_t st);

// mirrors and_smem_bytes() in step_vpsde_kernel.cuh
static size_t and_smem_need(int M, int D, int threads) {
  const int K = (M - 1) * M / 2 + 2 * (M - 1) + 2;
  return (size_t)(M + 1) * D * 4 + sizeof(double) * ((size_t)(threads / 32 + 2) * K + M) + 16;
}

// AND keeps the sample in shared memory (bulk-copy kernel) from this model count on (SDB_AND_SMEM_MIN_M overrides)
static int and_smem_min_m() {
  static const int v = [] {
    const char* e = getenv("SDB_AND_SMEM_MIN_M");
    return e && *e ? atoi(e) : 3;
  }();
  return v;
}

// AND switches from the register-resident kernel to the two-pass streaming kernel at this model count
// (SDB_AND_STREAM_MIN_M overrides; measured crossover in profiles/r01d_step_sweep.txt).
static int and_stream_min_m() {
  static const int v = [] {
    const char* e = getenv("SDB_AND_STREAM_MIN_M");
    return e && *e ? atoi(e) : 5;
  }();
  return v;
}

}  // namespace sdb

namespace sdb {

static int step_vpsde_impl(const float* x, const float* noise, const float* const* scores, int M, int B, int D,
                           float a_t, float b_t, float sigma_t, float dt, const float* sched, const int* step_counter,
                           int mode, int dlogq_mode, float temperature, const float* logp_bias, float ito_scale,
                           float* logq, float* x_out, float* weights, void* stream, int threads, int nv, int cluster,
                           float mix_scale = 2.f, const float* dlogq_add = nullptr, bool ode = false) {
  if (M < 1 || M > SD_MAX_MODELS) return fail(kErrInvalidArg, "sd_step_vpsde: M must be in [1, 8]");
  if (B < 0 || D < 1) return fail(kErrInvalidArg, "sd_step_vpsde: B >= 0 and D >= 1 required");
  if (mode < SD_MODE_OR || mode > SD_MODE_FIXED) return fail(kErrInvalidArg, "sd_step_vpsde: unknown mode");
  if (dlogq_mode < SD_DLOGQ_CIFAR_MAXSUB || dlogq_mode > SD_DLOGQ_NONE)
    return fail(kErrInvalidArg, "sd_step_vpsde: unknown dlogq_mode");
  if (B == 0) return SD_OK;  // empty batch: nothing to launch (pointers may be null)
  if (!x || (!noise && !ode) || !scores || !x_out || !weights || (!logq && (mode == SD_MODE_OR || dlogq_mode != SD_DLOGQ_NONE)))
    return fail(kErrInvalidArg, "sd_step_vpsde: null pointer argument");
  if (ode && mode == SD_MODE_AND)
    return fail(kErrUnsupported, "sd_step_vpsde_ode: the AND weights are defined by the noise of the stochastic step (superposition_edu.ipynb:899-905)");
  StepParams p{};
  p.x = x; p.noise = noise;
  for (int i = 0; i < M; ++i) {
    if (!scores[i]) return fail(kErrInvalidArg, "sd_step_vpsde: null score pointer");
    p.s[i] = scores[i];
  }
  p.logq = logq; p.x_out = x_out; p.weights = weights; p.logp_bias = logp_bias;
  p.sched = sched; p.step_counter = step_counter;
  p.M = M; p.B = B; p.D = D;
  p.a = a_t; p.b = b_t; p.sigma = sigma_t; p.dt = dt;
  p.mode = mode; p.dlogq_mode = dlogq_mode; p.temperature = temperature; p.ito_scale = ito_scale;
  p.mix_scale = mix_scale; p.dlogq_add = dlogq_add;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t err;

  if (D <= 64 && threads == 0) {
#define SDB_SMALL(Mv) case Mv: err = launch_small<Mv>(p, st); break;
    switch (M) {
      SDB_SMALL(1) SDB_SMALL(2) SDB_SMALL(3) SDB_SMALL(4) SDB_SMALL(5) SDB_SMALL(6) SDB_SMALL(7) SDB_SMALL(8)
      default: err = cudaErrorInvalidValue;
    }
#undef SDB_SMALL
    return check_cuda(err, "sd_step_vpsde (small-D launch)");
  }

  // vector width: float4 when every base pointer is 16B aligned and D % 4 == 0
  bool aligned = (D % 4 == 0) && (((uintptr_t)x | (uintptr_t)(noise ? noise : x) | (uintptr_t)x_out) % 16 == 0);
  for (int i = 0; i < M; ++i) aligned = aligned && ((uintptr_t)scores[i] % 16 == 0);
  const int vec = aligned ? 4 : 1;
  const int nunits = D / vec;
  const bool is_and = mode == SD_MODE_AND;
  bool and_stream = is_and && cluster == -1;                       // explicit requests through sd_step_vpsde_ex
  bool and_smem = is_and && cluster == -2;
  if (is_and && !and_stream && !and_smem && (threads == 0 || nv == 0 || cluster == 0)) {
    const int nv_cap = max(1, min(4, 24 / (M + 2)));
    const bool smem_ok = vec == 4 && nunits <= 4 * 256 && and_smem_need(M, D, 256) <= 227 * 1024;
    if (M >= and_smem_min_m() && smem_ok) {                        // sample resident in shared memory
      and_smem = true;
      threads = 256;
    } else if (M >= and_stream_min_m() || 256L * nv_cap < nunits) { // does not stay resident in one CTA
      and_stream = true;
      threads = 256;
      nv = M <= 4 ? 2 : 1;
    }
  }
  if (and_smem) {
    if (threads < 32 || threads > 256 || threads % 32)
      return fail(kErrInvalidArg, "sd_step_vpsde_ex: bad launch shape (shared-memory AND: threads 32..256)");
    if (vec != 4 || nunits > 4 * threads || and_smem_need(M, D, threads) > 227 * 1024)
      return fail(kErrUnsupported, "sd_step_vpsde_ex: shared-memory AND needs 16-byte aligned rows, D <= 16*threads and (M+1)*D*4 bytes of shared memory");
#define SDB_AM(Mv) case Mv: err = launch_and_smem<Mv>(p, threads, st); break;
    switch (M) {
      SDB_AM(1) SDB_AM(2) SDB_AM(3) SDB_AM(4) SDB_AM(5) SDB_AM(6) SDB_AM(7) SDB_AM(8)
      default: err = cudaErrorInvalidValue;
    }
#undef SDB_AM
    return check_cuda(err, "sd_step_vpsde launch (shared-memory AND)");
  }
  if (and_stream) {
    if (threads < 32 || threads > 256 || threads % 32 || nv < 1 || nv > 2)
      return fail(kErrInvalidArg, "sd_step_vpsde_ex: bad launch shape (streaming AND: threads 32..256, 1 or 2 vectors per thread)");
#define SDB_AS(Mv) case Mv: err = launch_and_stream<Mv>(p, threads, nv, vec, st); break;
    switch (M) {
      SDB_AS(1) SDB_AS(2) SDB_AS(3) SDB_AS(4) SDB_AS(5) SDB_AS(6) SDB_AS(7) SDB_AS(8)
      default: err = cudaErrorInvalidValue;
    }
#undef SDB_AS
    return check_cuda(err, "sd_step_vpsde launch (streaming AND)");
  }
  if ((threads == 0 || nv == 0 || cluster == 0) && !is_and && nunits >= 384 && B >= (M == 2 ? 768 : 384)) {
    // Weights known up front (OR / AVG / FIXED, stochastic or ODE) at batches that fill the chip: SMALL CTAs that walk the
    // sample in rounds - 128 threads x 3 float4 (x 1 for M >= 5).  ~1000 CTA slots instead of 444 make the last wave cheap and
    // keep more independent loads in flight per SM; measured (tools/step_sweep.py --shape, profiles/r01d_step_shapes*.txt)
    // 4-17 % ahead of one 256 x 3 round at batch 1024-4096 and, at batch 8192, OR at 1.01-1.02 of the copy peak for
    // M = 2, 3, 4, 8 (was 0.98 / 0.89 / 0.92 / 0.92) - what had looked like the cost of the softmax was the launch shape.
    threads = 128;
    nv = M <= 4 ? 3 : 1;
    cluster = 1;
  }
  if (threads == 0 || nv == 0 || cluster == 0) {
    // Heuristic: keep (M+2)*NV float4 registers per thread <= 24, prefer 128..256 threads and
    // enough CTAs (>= ~8 per SM) that the 148 SMs stay balanced.
    const int nv_cap = max(1, min(4, 24 / (M + 2)));
    int best_t = 256, best_nv = nv_cap, best_c = 8;
    long best_cost = -1;
    for (int c = 1; c <= 8; c *= 2)
      for (int t = 64; t <= 256; t += 32)
        for (int n = 1; n <= nv_cap; ++n) {
          const long cap = (long)c * t * n;
          if (is_and && cap < nunits) continue;            // AND needs the sample resident
          const long rounds = (nunits + cap - 1) / cap;
          const long waste = rounds * cap - nunits;          // idle lanes
          const long ctas = (long)B * c;
          // measured on B200 (tools/time_forward.py): one CTA per sample beats cluster splits at every
          // batch size that fills the chip, so clusters are only for residency (AND, large D) or tiny batches
          long cost = waste * 4 + (c > 1 ? 2000L * c : 0) + rounds * 16 + n;
          if (ctas < 148) cost += (148 - ctas) * 50;
          if (t < 128) cost += 32;
          // wave quantisation: estimated registers/thread = 4 per float4 held + 28; a second, nearly empty wave
          // (e.g. 512 CTAs on 444 resident slots) costs more than slightly fatter threads
          const long regs = 4L * (M + 2) * n + 28 + (is_and ? 16 : 0);
          long occ = 65536 / (regs * t);
          if (occ > 2048 / t) occ = 2048 / t;
          if (occ < 1) occ = 1;
          const long slots = 148 * occ;
          const long waves = (ctas + slots - 1) / slots;
          if (waves <= 3) cost += (waves * slots - ctas) * 2;     // idle CTA slots in the last wave
          if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_t = t; best_nv = n; best_c = c; }
        }
    if (best_cost < 0)
      return fail(kErrUnsupported, "sd_step_vpsde: D too large for the register-resident AND step (max 8*256*4 float4 per sample)");
    threads = best_t; nv = best_nv; cluster = best_c;
  }
  if (threads < 32 || threads > 256 || threads % 32 || nv < 1 || nv > 4 ||
      !(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8))
    return fail(kErrInvalidArg, "sd_step_vpsde_ex: bad launch shape");
  if (is_and && (long)cluster * threads * nv < nunits)
    return fail(kErrUnsupported, "sd_step_vpsde_ex: launch shape does not hold one sample (AND mode)");
#define SDB_M(Mv) case Mv: err = launch_m<Mv>(p, threads, nv, cluster, vec, st); break;
  switch (M) {
    SDB_M(1) SDB_M(2) SDB_M(3) SDB_M(4) SDB_M(5) SDB_M(6) SDB_M(7) SDB_M(8)
    default: err = cudaErrorInvalidValue;
  }
#undef SDB_M
  return check_cuda(err, "sd_step_vpsde launch");
}

__global__ void counter_add_kernel(int* c, int d) { *c += d; }
__global__ void counter_add_sat_kernel(int* c, int d, int last) { const int v = *c + d; *c = v > last ? last : (v < 0 ? 0 : v); }

}  // namespace sdb

extern "C" {

int sd_step_vpsde(const float* x, const float* noise, const float* const* scores_host, int M, int B, int D,
                  float a_t, float b_t, float sigma_t, float dt, const float* sched, const int* step_counter,
                  int mode, int dlogq_mode, float temperature, const float* logp_bias, float ito_scale,
                  float* logq, float* x_out, float* weights, void* stream) {
  return sdb::step_vpsde_impl(x, noise, scores_host, M, B, D, a_t, b_t, sigma_t, dt, sched, step_counter, mode,
                              dlogq_mode, temperature, logp_bias, ito_scale, logq, x_out, weights, stream, 0, 0, 0);
}

int sd_step_vpsde_ex(const float* x, const float* noise, const float* const* scores_host, int M, int B, int D,
                     float a_t, float b_t, float sigma_t, float dt, const float* sched, const int* step_counter,
                     int mode, int dlogq_mode, float temperature, const float* logp_bias, float ito_scale,
                     float* logq, float* x_out, float* weights, void* stream, int threads, int vec_per_thread,
                     int cluster) {
  return sdb::step_vpsde_impl(x, noise, scores_host, M, B, D, a_t, b_t, sigma_t, dt, sched, step_counter, mode,
                              dlogq_mode, temperature, logp_bias, ito_scale, logq, x_out, weights, stream, threads,
                              vec_per_thread, cluster);
}

int sd_step_vpsde_ode(const float* x, const float* const* scores_host, int M, int B, int D,
                      float a_t, float b_t, float sigma_eps, float dt, const float* sched, const int* step_counter,
                      int mode, int dlogq_mode, float temperature, const float* logp_bias, const float* dlogq_add,
                      float* logq, float* x_out, float* weights, void* stream) {
  return sdb::step_vpsde_impl(x, nullptr, scores_host, M, B, D, a_t, b_t, sigma_eps, dt, sched, step_counter, mode,
                              dlogq_mode, temperature, logp_bias, 0.f, logq, x_out, weights, stream, 0, 0, 0, 1.f,
                              dlogq_add, true);
}

int sd_counter_add(int* counter, int delta, void* stream) {
  if (!counter) return sdb::fail(sdb::kErrInvalidArg, "sd_counter_add: null counter");
  sdb::counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  return sdb::check_cuda(cudaGetLastError(), "sd_counter_add launch");
}

int sd_counter_add_sat(int* counter, int delta, int rows, void* stream) {
  if (!counter || rows < 1) return sdb::fail(sdb::kErrInvalidArg, "sd_counter_add_sat: null counter or rows < 1");
  sdb::counter_add_sat_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta, rows - 1);
  return sdb::check_cuda(cudaGetLastError(), "sd_counter_add_sat launch");
}

}  // extern "C"
