// Fused VP-SDE SuperDiff step for sm_100a: ONE launch per timestep.
//
// Replaces the body of get_joint_stoch_vf.joint_vf (reference cifar/dynamics.py:123-136),
// get_avg_vf.joint_vf (:155-171) and the toy notebook cells
// (notebooks/superposition_edu.ipynb:813-819, :899-905, :938-946).
//
// Layout: one thread-block CLUSTER per sample.  Each CTA owns a contiguous
// slice of the sample's D elements, keeps x / noise / all M scores for that
// slice in registers (one HBM read per element), and the cluster exchanges
// the K per-sample partial reductions through distributed shared memory.
//   weights known up front (OR / AVG / FIXED): K = M          (R_i * sigma)
//   AND (weights depend on the reductions):    K = M(M+1)/2+M (Gram G_ij, N_i)
// HBM bytes per sample: 4*D*(M+3)  (read x, noise, M scores; write x_out).
#pragma once
#include "common.cuh"
#include "../../include/superdiff_b200.h"
#include "step_vpsde_params.cuh"

namespace sdb {


__device__ __forceinline__ void ldv4(float (&d)[4], const float* p) {
  const float4 t = ld_stream4(p);
  d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
}

struct StepScalars {
  float a, b, sigma, dt;
};

__device__ __forceinline__ StepScalars load_scalars(const StepParams& p) {
  StepScalars s{p.a, p.b, p.sigma, p.dt};
  if (p.sched != nullptr) {
    const int row = p.step_counter ? *p.step_counter : 0;
    const float4 v = *reinterpret_cast<const float4*>(p.sched + 4 * (size_t)row);
    s.a = v.x; s.b = v.y; s.sigma = v.z; s.dt = v.w;
  }
  return s;
}

// Mixing weights that do not need this step's reductions.
template <int M>
__device__ __forceinline__ void known_weights(const StepParams& p, int sample, float (&w)[M]) {
  if (p.mode == SD_MODE_OR) {
    float z[M], zmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float l = p.logq[(size_t)sample * M + i];
      if (p.logp_bias) l += p.logp_bias[i];
      z[i] = p.temperature * l;
      zmax = fmaxf(zmax, z[i]);
    }
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) { z[i] = expf(z[i] - zmax); den += z[i]; }
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = z[i] / den;
  } else if (p.mode == SD_MODE_AVG) {
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = 1.0f / (float)M;
  } else {  // FIXED
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = p.weights[(size_t)sample * M + i];
  }
}

// AND weights in the cancellation-free difference form.  With d_i = s_i - s_M (i < M-1... Md = M-1 of them),
// D_ij = <d_i, d_j>, E_i = <d_i, noise>, equalising the Ito increments R_i reduces to the
// symmetric (M-1)x(M-1) system   2 dt b sum_j D_ij kappa_j = dt b D_ii - c E_i,   kappa_M = 1 - sum_j kappa_j.
// For M = 2 this is the notebook's select_kappa (superposition_edu.ipynb:899-905):
//   kappa_1 = 1/2 - c <s1-s2, noise> / (2 dt b |s1-s2|^2).
// The differences are formed element-wise in fp32 (exact-ish, like the reference's (s1-s2)**2), so nearly equal
// models (t ~ 1) do not lose the denominator to cancellation the way a Gram-matrix formulation does.
// D packed upper triangular over Md: idx(i,j), i <= j.
template <int M>
__device__ void and_solve(const double* D, const double* E, double dtb, double c, double* kappa) {
  constexpr int Md = M - 1;
  auto d = [&](int i, int j) {
    if (i > j) { int t = i; i = j; j = t; }
    return D[i * Md - (i * (i - 1)) / 2 + (j - i)];
  };
  if (M == 1) { kappa[0] = 1.0; return; }
  if (M == 2) {
    kappa[0] = (dtb * D[0] - c * E[0]) / (2.0 * dtb * D[0]);
    kappa[1] = 1.0 - kappa[0];
    return;
  }
  double A[Md > 0 ? Md : 1][Md + 1];
  for (int i = 0; i < Md; ++i) {
    for (int j = 0; j < Md; ++j) A[i][j] = 2.0 * dtb * d(i, j);
    A[i][Md] = dtb * d(i, i) - c * E[i];
  }
  for (int col = 0; col < Md; ++col) {  // Gaussian elimination, partial pivoting
    int piv = col;
    double best = fabs(A[col][col]);
    for (int r = col + 1; r < Md; ++r)
      if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
    if (piv != col)
      for (int j = 0; j <= Md; ++j) { double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
    const double inv = 1.0 / A[col][col];
    for (int r = col + 1; r < Md; ++r) {
      const double f = A[r][col] * inv;
      for (int j = col; j <= Md; ++j) A[r][j] -= f * A[col][j];
    }
  }
  double sum = 0.0;
  for (int i = Md - 1; i >= 0; --i) {
    double acc = A[i][Md];
    for (int j = i + 1; j < Md; ++j) acc -= A[i][j] * kappa[j];
    kappa[i] = acc / A[i][i];
  }
  for (int i = 0; i < Md; ++i) sum += kappa[i];
  kappa[Md] = 1.0 - sum;
}

// Ito increments under AND from the difference reductions:  R_M = [2dtb(G_MM + sum_j kappa_j F_j) + c N_M - dtb G_MM]/sigma
// with F_j = <s_M, d_j>, and R_i = R_M + (row-i residual of the system)/sigma, which is 0 up to rounding.
// layout of `tot`: D (Md(Md+1)/2) | E (Md) | F (Md) | G_MM | N_M
template <int M>
__device__ void and_increments(const double* tot, const double* kappa, double dtb, double c, double sigma, double* R) {
  constexpr int Md = M - 1;
  constexpr int ND = Md * (Md + 1) / 2;
  const double* D = tot;
  const double* E = tot + ND;
  const double* F = tot + ND + Md;
  const double Gmm = tot[ND + 2 * Md], Nm = tot[ND + 2 * Md + 1];
  auto d = [&](int i, int j) {
    if (i > j) { int t = i; i = j; j = t; }
    return D[i * Md - (i * (i - 1)) / 2 + (j - i)];
  };
  double mixF = 0.0;
  for (int j = 0; j < Md; ++j) mixF += kappa[j] * F[j];
  const double Rm = (2.0 * dtb * (Gmm + mixF) + c * Nm - dtb * Gmm) / sigma;
  R[Md] = Rm;
  for (int i = 0; i < Md; ++i) {
    double acc = 0.0;
    for (int j = 0; j < Md; ++j) acc += kappa[j] * d(i, j);
    R[i] = Rm + (2.0 * dtb * acc - dtb * d(i, i) + c * E[i]) / sigma;
  }
}

__device__ __forceinline__ void write_logq(const StepParams& p, const StepScalars& sc, int sample, int M,
                                           const double* R /* already divided by sigma */) {
  if (p.dlogq_mode == SD_DLOGQ_NONE) return;
  double Ra[SD_MAX_MODELS];
  for (int i = 0; i < M; ++i) Ra[i] = R[i] + (p.dlogq_add ? (double)p.dlogq_add[(size_t)sample * M + i] : 0.0);
  R = Ra;
  double sub = 0.0;
  if (p.dlogq_mode == SD_DLOGQ_CIFAR_MAXSUB) {
    sub = -R[0];
    for (int i = 1; i < M; ++i) sub = fmin(sub, -R[i]);   // -max_i R_i
  } else {
    sub = (double)p.ito_scale * (double)sc.dt * (double)sc.a;
  }
  for (int i = 0; i < M; ++i) {
    float* q = p.logq + (size_t)sample * M + i;
    // the reference accumulates logq in fp32 (cifar/eval_utils.py:84)
    *q = *q + (float)(R[i] + sub);
  }
}

// VEC = 4: float4 accesses (D % 4 == 0, 16B-aligned bases); VEC = 1: scalar.
template <int M, int NV, int VEC, bool AND, bool CLUSTER>
__global__ void __launch_bounds__(256) step_vpsde_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ double scratch[];
  unsigned csize = 1, crank = 0;
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    csize = cluster.num_blocks();
    crank = cluster.block_rank();
  }
  const int sample = blockIdx.x / csize;
  const StepScalars sc = load_scalars(p);
  const float dta = sc.dt * sc.a, dtb = sc.dt * sc.b;
  const float c = p.noise ? sqrtf(2.f * sc.sigma * sc.b * sc.dt) : 0.f;     // no noise tensor: deterministic (ODE) step
  const float mixc = p.mix_scale * dtb;
  const int nunits = p.D / VEC;                                   // float4 (or scalar) units per sample
  const int per_cta = (nunits + csize - 1) / csize;
  const int u0 = crank * per_cta;
  const int u1 = min(nunits, u0 + per_cta);
  const size_t base = (size_t)sample * p.D;

  constexpr int KAND = (M - 1) * M / 2 + 2 * (M - 1) + 2;   // D | E | F | G_MM | N_M
  constexpr int K = AND ? KAND : M;
  float part[K];
#pragma unroll
  for (int k = 0; k < K; ++k) part[k] = 0.f;

  if (!AND) {
    float w[M];
    known_weights<M>(p, sample, w);
    // weights known: stream the slice in rounds of NV units per thread
    for (int r0 = u0; r0 < u1; r0 += NV * blockDim.x) {
      float xv[NV][VEC], ev[NV][VEC], sv[M][NV][VEC];
      bool ok[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int u = r0 + j * blockDim.x + threadIdx.x;
        ok[j] = u < u1;
        if (ok[j]) {
          const size_t off = base + (size_t)u * VEC;
          if constexpr (VEC == 4) {
            ldv4(xv[j], p.x + off);
            if (p.noise) ldv4(ev[j], p.noise + off);
            else { ev[j][0] = 0.f; ev[j][1] = 0.f; ev[j][2] = 0.f; ev[j][3] = 0.f; }
#pragma unroll
            for (int i = 0; i < M; ++i) ldv4(sv[i][j], p.s[i] + off);
          } else {
            xv[j][0] = ld_stream1(p.x + off);
            ev[j][0] = p.noise ? ld_stream1(p.noise + off) : 0.f;
#pragma unroll
            for (int i = 0; i < M; ++i) sv[i][j][0] = ld_stream1(p.s[i] + off);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (!ok[j]) continue;
        float o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float mix = 0.f;
#pragma unroll
          for (int i = 0; i < M; ++i) mix = fmaf(w[i], sv[i][j][e], mix);
          const float xe = xv[j][e];
          const float dx = -dta * xe + mixc * mix + c * ev[j][e];
          o[e] = xe + dx;
          const float q = dx + dta * xe;                // dx + dt*a*x
#pragma unroll
          for (int i = 0; i < M; ++i) {
            const float s = sv[i][j][e];
            part[i] = fmaf(s, q - dtb * s, part[i]);    // s_i*(dx + dt*a*x - dt*b*s_i)
          }
        }
        const size_t off = base + (size_t)(r0 + j * blockDim.x + threadIdx.x) * VEC;
        if constexpr (VEC == 4) st4(p.x_out + off, make_float4(o[0], o[1], o[2], o[3]));
        else p.x_out[off] = o[0];
      }
    }
    const double* tot = block_cluster_sum<K, CLUSTER>(part, scratch);
    if (crank == 0 && threadIdx.x == 0) {
      double R[M];
#pragma unroll
      for (int i = 0; i < M; ++i) R[i] = tot[i] / (double)sc.sigma;
      write_logq(p, sc, sample, M, R);
      if (p.mode != SD_MODE_FIXED)
        for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
    }
  } else {
    // AND: the slice stays resident in registers between the reduction pass and the write pass
    float xv[NV][VEC], ev[NV][VEC], sv[M][NV][VEC];
    bool ok[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = u0 + j * blockDim.x + threadIdx.x;
      ok[j] = u < u1;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        xv[j][e] = 0.f; ev[j][e] = 0.f;
#pragma unroll
        for (int i = 0; i < M; ++i) sv[i][j][e] = 0.f;
      }
      if (ok[j]) {
        const size_t off = base + (size_t)u * VEC;
        if constexpr (VEC == 4) {
          ldv4(xv[j], p.x + off);
          ldv4(ev[j], p.noise + off);
#pragma unroll
          for (int i = 0; i < M; ++i) ldv4(sv[i][j], p.s[i] + off);
        } else {
          xv[j][0] = ld_stream1(p.x + off);
          ev[j][0] = ld_stream1(p.noise + off);
#pragma unroll
          for (int i = 0; i < M; ++i) sv[i][j][0] = ld_stream1(p.s[i] + off);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        constexpr int Md = M - 1;
        constexpr int ND = Md * (Md + 1) / 2;
        const float sm = sv[M - 1][j][e];
        float d[Md > 0 ? Md : 1];
#pragma unroll
        for (int i = 0; i < Md; ++i) d[i] = sv[i][j][e] - sm;
        int k = 0;
#pragma unroll
        for (int i = 0; i < Md; ++i)
#pragma unroll
          for (int l = i; l < Md; ++l) { part[k] = fmaf(d[i], d[l], part[k]); ++k; }
#pragma unroll
        for (int i = 0; i < Md; ++i) {
          part[ND + i] = fmaf(d[i], ev[j][e], part[ND + i]);
          part[ND + Md + i] = fmaf(sm, d[i], part[ND + Md + i]);
        }
        part[ND + 2 * Md] = fmaf(sm, sm, part[ND + 2 * Md]);
        part[ND + 2 * Md + 1] = fmaf(sm, ev[j][e], part[ND + 2 * Md + 1]);
      }
    const double* tot = block_cluster_sum<K, CLUSTER>(part, scratch);
    // kappa: closed form for M <= 2 (every thread), one solver thread + smem broadcast otherwise
    double* kappa_sh = scratch + ((blockDim.x + 31) / 32 + 2) * K;
    double kappa[M];
    if (M <= 2) {
      and_solve<M>(tot, tot + (M - 1) * M / 2, (double)sc.dt * (double)sc.b, (double)c, kappa);
    } else {
      if (threadIdx.x == 0) {
        and_solve<M>(tot, tot + (M - 1) * M / 2, (double)sc.dt * (double)sc.b, (double)c, kappa);
        for (int i = 0; i < M; ++i) kappa_sh[i] = kappa[i];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < M; ++i) kappa[i] = kappa_sh[i];
    }
    float w[M];
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = (float)kappa[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (!ok[j]) continue;
      float o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        // s_M + sum_j kappa_j (s_j - s_M): the reference's form (superposition_edu.ipynb:942), well conditioned
        // for the unclipped |kappa| >> 1 that occurs while the models still agree (t ~ 1)
        const float sm = sv[M - 1][j][e];
        float mix = sm;
#pragma unroll
        for (int i = 0; i < M - 1; ++i) mix = fmaf(w[i], sv[i][j][e] - sm, mix);
        o[e] = xv[j][e] + (-dta * xv[j][e] + 2.f * dtb * mix + c * ev[j][e]);
      }
      const size_t off = base + (size_t)(u0 + j * blockDim.x + threadIdx.x) * VEC;
      if constexpr (VEC == 4) st4(p.x_out + off, make_float4(o[0], o[1], o[2], o[3]));
      else p.x_out[off] = o[0];
    }
    if (crank == 0 && threadIdx.x == 0) {
      double R[M];
      and_increments<M>(tot, kappa, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, R);
      write_logq(p, sc, sample, M, R);
      for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
    }
  }
}

// Tiny-D path (toy 2-D notebook, D <= 64): one thread per sample, fp64 reductions.
template <int M>
__global__ void __launch_bounds__(128) step_vpsde_small_kernel(const __grid_constant__ StepParams p) {
  const int sample = blockIdx.x * blockDim.x + threadIdx.x;
  if (sample >= p.B) return;
  const StepScalars sc = load_scalars(p);
  const float dta = sc.dt * sc.a, dtb = sc.dt * sc.b;
  const float c = p.noise ? sqrtf(2.f * sc.sigma * sc.b * sc.dt) : 0.f;
  const float mixc = p.mix_scale * dtb;
  const size_t base = (size_t)sample * p.D;
  float w[M];
  double R[M];
  if (p.mode == SD_MODE_AND) {
    constexpr int Md = M - 1;
    constexpr int ND = Md * (Md + 1) / 2;
    double Dm[ND > 0 ? ND : 1], E[Md > 0 ? Md : 1];
    for (int k = 0; k < ND; ++k) Dm[k] = 0.0;
    for (int i = 0; i < Md; ++i) E[i] = 0.0;
    for (int d_ = 0; d_ < p.D; ++d_) {
      const float sm = p.s[M - 1][base + d_];
      const float e = p.noise[base + d_];
      float df[Md > 0 ? Md : 1];
#pragma unroll
      for (int i = 0; i < Md; ++i) df[i] = p.s[i][base + d_] - sm;
      int k = 0;
#pragma unroll
      for (int i = 0; i < Md; ++i) {
#pragma unroll
        for (int l = i; l < Md; ++l) Dm[k++] += (double)df[i] * (double)df[l];
        E[i] += (double)df[i] * (double)e;
      }
    }
    double kappa[M];
    and_solve<M>(Dm, E, (double)sc.dt * (double)sc.b, (double)c, kappa);
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = (float)kappa[i];
  } else {
    known_weights<M>(p, sample, w);
  }
#pragma unroll
  for (int i = 0; i < M; ++i) R[i] = 0.0;
  for (int d = 0; d < p.D; ++d) {
    float s[M];
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = p.s[i][base + d];
    float mix = s[M - 1];             // s_M + sum_j kappa_j (s_j - s_M), see the AND path of the vector kernel
#pragma unroll
    for (int i = 0; i < M - 1; ++i) mix = fmaf(w[i], s[i] - s[M - 1], mix);
    const float xe = p.x[base + d];
    const float dx = -dta * xe + mixc * mix + (p.noise ? c * p.noise[base + d] : 0.f);
    const float q = dx + dta * xe;
#pragma unroll
    for (int i = 0; i < M; ++i) R[i] += (double)s[i] * (double)(q - dtb * s[i]);
    p.x_out[base + d] = xe + dx;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) R[i] /= (double)sc.sigma;
  write_logq(p, sc, sample, M, R);
  if (p.mode != SD_MODE_FIXED)
    for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
}

template <int M, int NV, int VEC, bool AND>
static cudaError_t launch_cfg(const StepParams& p, int threads, int cluster, cudaStream_t st) {
  constexpr int K = AND ? ((M - 1) * M / 2 + 2 * (M - 1) + 2) : M;
  const size_t smem = sizeof(double) * ((size_t)(threads / 32 + 2) * K + M);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.B * cluster);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (cluster > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, step_vpsde_kernel<M, NV, VEC, AND, true>, p);
  }
  return cudaLaunchKernelEx(&cfg, step_vpsde_kernel<M, NV, VEC, AND, false>, p);
}

template <int M>
cudaError_t launch_m(const StepParams& p, int threads, int nv, int cluster, int vec, cudaStream_t st) {
  const bool is_and = p.mode == SD_MODE_AND;
#define SDB_CASE(NVv)                                                                        \
  case NVv:                                                                                  \
    if (vec == 4) return is_and ? launch_cfg<M, NVv, 4, true>(p, threads, cluster, st)        \
                                : launch_cfg<M, NVv, 4, false>(p, threads, cluster, st);      \
    return is_and ? launch_cfg<M, NVv, 1, true>(p, threads, cluster, st)                      \
                  : launch_cfg<M, NVv, 1, false>(p, threads, cluster, st);
  switch (nv) {
    SDB_CASE(1)
    SDB_CASE(2)
    SDB_CASE(3)
    SDB_CASE(4)
  }
#undef SDB_CASE
  return cudaErrorInvalidValue;
}

template <int M>
cudaError_t launch_small(const StepParams& p, cudaStream_t st) {
  const int threads = 128;
  step_vpsde_small_kernel<M><<<(p.B + threads - 1) / threads, threads, 0, st>>>(p);
  return cudaGetLastError();
}


}  // namespace sdb
