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
#include "tcgen05_util.cuh"   // mbarrier wrappers
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

// Per-step scalars come either from the arguments or from the device-side schedule table (`sched[row]`, row = *step_counter,
// so a captured CUDA graph replays for every timestep).  The table form is a chain of two dependent loads; issued and
// consumed at the top of the kernel, every CTA spent two memory round trips with no data load in flight.  So it is
// split: begin_scalars ISSUES the counter load (volatile asm keeps it above the streaming loads), finish_scalars - placed
// after the sample's loads have been issued - reads the row and forms the scalars.
__device__ __forceinline__ int begin_scalars(const StepParams& p) {
  int row = 0;
  if (p.sched != nullptr && p.step_counter != nullptr)
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(row) : "l"(p.step_counter));
  return row;
}
__device__ __forceinline__ StepScalars finish_scalars(const StepParams& p, int row) {
  StepScalars s{p.a, p.b, p.sigma, p.dt};
  if (p.sched != nullptr) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p.sched + 4 * (size_t)row));
    s.a = v.x; s.b = v.y; s.sigma = v.z; s.dt = v.w;
  }
  return s;
}
__device__ __forceinline__ StepScalars load_scalars(const StepParams& p) { return finish_scalars(p, begin_scalars(p)); }

// Mixing weights that do not need this step's reductions.  finish_weights consumes values whose loads were ISSUED before the
// sample's streaming loads (LogqState::old, and the caller's predicated load of the FIXED weights) and must itself come
// after them - see the call site: with the softmax ahead of the streaming loads every CTA starts with a dependent DRAM
// round trip during which it has nothing in flight (worth 3-5 % at 2 CTAs / SM; the larger part of the old OR-vs-AVG gap
// turned out to be the launch shape, see step_vpsde.cu).
template <int M>
__device__ __forceinline__ void finish_weights(const StepParams& p, const float (&logq_old)[M], const float (&wfix)[M],
                                               float (&w)[M]) {
  if (p.mode == SD_MODE_OR) {
    float z[M], zmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      z[i] = p.temperature * (logq_old[i] + (p.logp_bias ? p.logp_bias[i] : 0.f));
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
    for (int i = 0; i < M; ++i) w[i] = wfix[i];
  }
}

template <int M>
__device__ __forceinline__ void known_weights(const StepParams& p, int sample, float (&w)[M]) {
  float lq[M], wfix[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    lq[i] = (p.mode == SD_MODE_OR) ? p.logq[(size_t)sample * M + i] : 0.f;
    wfix[i] = (p.mode == SD_MODE_FIXED) ? p.weights[(size_t)sample * M + i] : 0.f;
  }
  finish_weights<M>(p, lq, wfix, w);
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
__device__ __forceinline__ void and_solve(const double* D, const double* E, double dtb, double c, double* kappa) {
  constexpr int Md = M - 1;
  if (M == 1) { kappa[0] = 1.0; return; }
  if (M == 2) {
    kappa[0] = (dtb * D[0] - c * E[0]) * fast_drcp(2.0 * dtb * D[0]);
    kappa[1] = 1.0 - kappa[0];
    return;
  }
  // The matrix 2 dt b D is a Gram matrix (symmetric positive definite unless two differences are linearly dependent), so
  // elimination needs no pivoting; only the upper triangle is kept and every index below is a compile-time constant, so
  // the whole solve lives in registers.  (The first version pivoted over a local-memory array with run-time indices:
  // ~10 us of dependent local-memory accesses by one thread per sample at M = 8 - longer than moving the sample.)
  double U[Md > 0 ? Md : 1][Md > 0 ? Md : 1], rhs[Md > 0 ? Md : 1];
#pragma unroll
  for (int i = 0; i < Md; ++i) {
#pragma unroll
    for (int j = i; j < Md; ++j) U[i][j] = 2.0 * dtb * D[i * Md - (i * (i - 1)) / 2 + (j - i)];
    rhs[i] = dtb * D[i * Md - (i * (i - 1)) / 2] - c * E[i];
  }
  double pinv[Md > 0 ? Md : 1];                   // reciprocal pivots, reused by the back substitution
#pragma unroll
  for (int col = 0; col < Md; ++col) {
    const double inv = fast_drcp(U[col][col]);
    pinv[col] = inv;
#pragma unroll
    for (int r = col + 1; r < Md; ++r) {
      const double f = U[col][r] * inv;          // A[r][col] = A[col][r]: the trailing block stays symmetric
#pragma unroll
      for (int j = r; j < Md; ++j) U[r][j] -= f * U[col][j];
      rhs[r] -= f * rhs[col];
    }
  }
  double sum = 0.0;
#pragma unroll
  for (int i = Md - 1; i >= 0; --i) {
    double acc = rhs[i];
#pragma unroll
    for (int j = i + 1; j < Md; ++j) acc -= U[i][j] * kappa[j];
    kappa[i] = acc * pinv[i];
  }
#pragma unroll
  for (int i = 0; i < Md; ++i) sum += kappa[i];
  kappa[Md] = 1.0 - sum;
}

// The same solve by ONE WARP, for the vector kernels with M > 2.  Serial, by thread 0, it was a chain of ~450 dependent fp64
// instructions (plus spills of the 35-entry triangle): ncu showed 57 % of all stall samples of the M = 8 kernel at the
// barrier behind it - the solve took longer than moving the sample.  Here lane r owns row r of the augmented system
// [2 dt b D | rhs] (Md + 1 doubles, compile-time indices only); Gauss-Jordan without pivoting (the matrix is a Gram matrix):
// per column one pivot-row broadcast by shuffle, one reciprocal, and every other lane updates its row in parallel, so there
// is no back-substitution chain.  Called by all 32 lanes of warp 0; writes kappa[0..M) to shared memory.
template <int M>
__device__ __forceinline__ void and_solve_warp(const double* tot, double dtb, double c, double* kappa_sh) {
  constexpr int Md = M - 1;
  constexpr int ND = Md * (Md + 1) / 2;
  const int lane = threadIdx.x & 31;
  const int r = lane < Md ? lane : Md - 1;          // spare lanes mirror the last row; their results are discarded
  double a[Md + 1];
#pragma unroll
  for (int j = 0; j < Md; ++j) {
    const int lo = r < j ? r : j, hi = r < j ? j : r;
    a[j] = 2.0 * dtb * tot[lo * Md - (lo * (lo - 1)) / 2 + (hi - lo)];
  }
  a[Md] = dtb * tot[r * Md - (r * (r - 1)) / 2] - c * tot[ND + r];
  double dinv = 1.0;
#pragma unroll
  for (int col = 0; col < Md; ++col) {
    const double inv = fast_drcp(__shfl_sync(0xffffffffu, a[col], col));
    const double f = a[col] * inv;
    if (lane == col) dinv = inv;                    // own (final) pivot, reciprocal
#pragma unroll
    for (int j = col + 1; j <= Md; ++j) {
      const double pj = __shfl_sync(0xffffffffu, a[j], col);
      if (lane != col) a[j] -= f * pj;
    }
  }
  const double x = a[Md] * dinv;
  double sum = lane < Md ? x : 0.0;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);     // Md <= 7: lanes 0..7
  if (lane < Md) kappa_sh[lane] = x;
  if (lane == 0) kappa_sh[Md] = 1.0 - sum;
}

// Ito increments by warp 0 (lane i forms R_i), written to R_sh[0..M) in shared memory; see and_increments for the algebra.
template <int M>
__device__ __forceinline__ void and_increments_warp(const double* tot, const double* kappa_sh, double dtb, double c,
                                                    double sigma, double* R_sh) {
  constexpr int Md = M - 1;
  constexpr int ND = Md * (Md + 1) / 2;
  const int lane = threadIdx.x & 31;
  const double* E = tot + ND;
  const double* F = tot + ND + Md;
  const double Gmm = tot[ND + 2 * Md], Nm = tot[ND + 2 * Md + 1];
  double mixF = 0.0;
#pragma unroll
  for (int j = 0; j < Md; ++j) mixF += kappa_sh[j] * F[j];
  const double inv_sigma = fast_drcp(sigma);
  const double Rm = (2.0 * dtb * (Gmm + mixF) + c * Nm - dtb * Gmm) * inv_sigma;
  if (lane < Md) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < Md; ++j) {
      const int lo = lane < j ? lane : j, hi = lane < j ? j : lane;
      acc += kappa_sh[j] * tot[lo * Md - (lo * (lo - 1)) / 2 + (hi - lo)];
    }
    const double dii = tot[lane * Md - (lane * (lane - 1)) / 2];
    R_sh[lane] = Rm + (2.0 * dtb * acc - dtb * dii + c * E[lane]) * inv_sigma;
  } else if (lane == Md) {
    R_sh[Md] = Rm;
  }
  __syncwarp();
}

// Ito increments under AND from the difference reductions:  R_M = [2dtb(G_MM + sum_j kappa_j F_j) + c N_M - dtb G_MM]/sigma
// with F_j = <s_M, d_j>, and R_i = R_M + (row-i residual of the system)/sigma, which is 0 up to rounding.
// layout of `tot`: D (Md(Md+1)/2) | E (Md) | F (Md) | G_MM | N_M
template <int M>
__device__ __forceinline__ void and_increments(const double* tot, const double* kappa, double dtb, double c, double sigma, double* R) {
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
#pragma unroll
  for (int j = 0; j < Md; ++j) mixF += kappa[j] * F[j];
  const double inv_sigma = fast_drcp(sigma);
  const double Rm = (2.0 * dtb * (Gmm + mixF) + c * Nm - dtb * Gmm) * inv_sigma;
  R[Md] = Rm;
#pragma unroll
  for (int i = 0; i < Md; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < Md; ++j) acc += kappa[j] * d(i, j);
    R[i] = Rm + (2.0 * dtb * acc - dtb * d(i, i) + c * E[i]) * inv_sigma;
  }
}

// The old log-densities (and the optional additive term) are loaded at the START of the kernel (load_logq_state), next to
// the streaming loads: read here, at the tail, they were a dependent ~0.7 us DRAM round trip by one thread while the CTA
// still held its slot on the SM.
template <int M>
struct LogqState {
  float old[M];
  float add[M];
};

template <int M>
__device__ __forceinline__ void load_logq_state(const StepParams& p, int sample, LogqState<M>& st) {
  const bool need = p.logq != nullptr && (p.dlogq_mode != SD_DLOGQ_NONE || p.mode == SD_MODE_OR);
#pragma unroll
  for (int i = 0; i < M; ++i) {
    st.old[i] = need ? p.logq[(size_t)sample * M + i] : 0.f;
    st.add[i] = p.dlogq_add ? p.dlogq_add[(size_t)sample * M + i] : 0.f;
  }
}

template <int M>
__device__ __forceinline__ void write_logq(const StepParams& p, const StepScalars& sc, int sample, const LogqState<M>& st,
                                           const double* R /* already divided by sigma */) {
  if (p.dlogq_mode == SD_DLOGQ_NONE) return;
  double Ra[M];
#pragma unroll
  for (int i = 0; i < M; ++i) Ra[i] = R[i] + (double)st.add[i];
  double sub = 0.0;
  if (p.dlogq_mode == SD_DLOGQ_CIFAR_MAXSUB) {
    sub = -Ra[0];
#pragma unroll
    for (int i = 1; i < M; ++i) sub = fmin(sub, -Ra[i]);   // -max_i R_i
  } else {
    sub = (double)p.ito_scale * (double)sc.dt * (double)sc.a;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) {
    // the reference accumulates logq in fp32 (cifar/eval_utils.py:84)
    p.logq[(size_t)sample * M + i] = st.old[i] + (float)(Ra[i] + sub);
  }
}

// VEC = 4: float4 accesses (D % 4 == 0, 16B-aligned bases); VEC = 1: scalar.
// (M+2)*NV <= 12 float4 per thread (M = 2 at D = 3072): hold the kernel to 3 CTAs of 256 threads per SM (<= 85 registers)
template <int M, int NV, int VEC, bool AND, bool CLUSTER>
__global__ void __launch_bounds__(256, ((M + 2) * NV * (VEC / 4) <= 12 && VEC == 4) ? 3 : 1) step_vpsde_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ double scratch[];
  unsigned csize = 1, crank = 0;
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    csize = cluster.num_blocks();
    crank = cluster.block_rank();
  }
  const int sample = blockIdx.x / csize;
  const int sched_row = begin_scalars(p);
  StepScalars sc{};
  float dta = 0.f, dtb = 0.f, c = 0.f, mixc = 0.f;
  auto form_scalars = [&]() {
    sc = finish_scalars(p, sched_row);
    dta = sc.dt * sc.a; dtb = sc.dt * sc.b;
    c = p.noise ? sqrtf(2.f * sc.sigma * sc.b * sc.dt) : 0.f;     // no noise tensor: deterministic (ODE) step
    mixc = p.mix_scale * dtb;
  };
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
    float w[M], wfix[M];
    LogqState<M> lqs;
    load_logq_state<M>(p, sample, lqs);          // predicated loads, issued here and consumed after the first round's data loads
#pragma unroll
    for (int i = 0; i < M; ++i) wfix[i] = (p.mode == SD_MODE_FIXED) ? p.weights[(size_t)sample * M + i] : 0.f;
    bool have_w = false;
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
      if (!have_w) {
        // Nothing may touch the loaded log-densities before this point: any use (even a register move merging the
        // per-mode sources) waits for the ~1 us uncached load while the CTA has no data load in flight.
        // The empty volatile asm pins the softmax below the streaming loads (volatile
        // asm statements keep their order; otherwise the loop-invariant softmax is hoisted to the loop pre-header).
#pragma unroll
        for (int i = 0; i < M; ++i) asm volatile("" : "+f"(lqs.old[i]), "+f"(wfix[i]));
        form_scalars();
        finish_weights<M>(p, lqs.old, wfix, w);
        have_w = true;
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
    if (!have_w) { form_scalars(); finish_weights<M>(p, lqs.old, wfix, w); }     // empty slice (cluster rank past the end of a short sample)
    const double* tot = block_cluster_sum<K, CLUSTER>(part, scratch);
    if (crank == 0 && threadIdx.x == 0) {
      double R[M];
      const double inv_sigma = fast_drcp((double)sc.sigma);       // one reciprocal on the CTA's tail, not M divisions
#pragma unroll
      for (int i = 0; i < M; ++i) R[i] = tot[i] * inv_sigma;
      write_logq<M>(p, sc, sample, lqs, R);
      if (p.mode != SD_MODE_FIXED) {
#pragma unroll
        for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
      }
    }
  } else {
    // AND: the slice stays resident in registers between the reduction pass and the write pass
    LogqState<M> lqs;
    load_logq_state<M>(p, sample, lqs);
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
    form_scalars();                      // after the sample's loads have been issued
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
    // kappa: closed form for M <= 2 (every thread), warp 0 solves + smem broadcast otherwise
    double* kappa_sh = scratch + ((blockDim.x + 31) / 32 + 2) * K;
    double kappa[M];
    if constexpr (M <= 2) {
      and_solve<M>(tot, tot + (M - 1) * M / 2, (double)sc.dt * (double)sc.b, (double)c, kappa);
    } else {
      if (threadIdx.x < 32) and_solve_warp<M>(tot, (double)sc.dt * (double)sc.b, (double)c, kappa_sh);
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
    if constexpr (M <= 2) {
      if (crank == 0 && threadIdx.x == 0) {
        double R[M];
        and_increments<M>(tot, kappa, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, R);
        write_logq<M>(p, sc, sample, lqs, R);
#pragma unroll
        for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
      }
    } else if (crank == 0 && threadIdx.x < 32) {
      // scratch[0..M): the warp partials of the block sum are dead (no cluster peer reads them)
      and_increments_warp<M>(tot, kappa_sh, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, scratch);
      if (threadIdx.x == 0) {
        write_logq<M>(p, sc, sample, lqs, scratch);
#pragma unroll
        for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
      }
    }
  }
}

// AND for model counts whose sample does not stay register-resident in one CTA (M >= 5 at D = 3072: the resident form
// needed a 2-CTA cluster at 254 registers and ran at 0.23 of the copy peak for M = 8).  Two streaming passes by ONE CTA
// per sample: pass 1 reads noise + the M scores in rounds of NV units per thread and accumulates the K difference-form
// reductions; the CTA solves for kappa; pass 2 re-reads x, noise and the scores - (M+2)*4*D bytes, <= 120 KB, which the
// same CTA touched microseconds earlier, so they come from L2, not HBM - and writes x'.  Same reductions, same solve, same
// mix expression as the resident kernel: results are bit-identical for a given CTA shape.  No limit on D.
template <int M, int NV, int VEC>
__global__ void __launch_bounds__(256, 2) step_vpsde_and_stream_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ double scratch[];
  const int sample = blockIdx.x;
  const int sched_row = begin_scalars(p);
  const int nunits = p.D / VEC;
  const size_t base = (size_t)sample * p.D;
  constexpr int Md = M - 1;
  constexpr int ND = Md * (Md + 1) / 2;
  constexpr int K = ND + 2 * Md + 2;                       // D | E | F | G_MM | N_M
  float part[K];
#pragma unroll
  for (int k = 0; k < K; ++k) part[k] = 0.f;
  LogqState<M> lqs;
  load_logq_state<M>(p, sample, lqs);

  for (int r0 = 0; r0 < nunits; r0 += NV * blockDim.x) {
    float ev[NV][VEC], sv[M][NV][VEC];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = r0 + j * blockDim.x + threadIdx.x;
      if (u < nunits) {
        const size_t off = base + (size_t)u * VEC;
        if constexpr (VEC == 4) {
          ldv4(ev[j], p.noise + off);
#pragma unroll
          for (int i = 0; i < M; ++i) ldv4(sv[i][j], p.s[i] + off);
        } else {
          ev[j][0] = ld_stream1(p.noise + off);
#pragma unroll
          for (int i = 0; i < M; ++i) sv[i][j][0] = ld_stream1(p.s[i] + off);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          ev[j][e] = 0.f;
#pragma unroll
          for (int i = 0; i < M; ++i) sv[i][j][e] = 0.f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
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
  }
  const StepScalars sc = finish_scalars(p, sched_row);
  const float dta = sc.dt * sc.a, dtb = sc.dt * sc.b;
  const float c = sqrtf(2.f * sc.sigma * sc.b * sc.dt);
  const double* tot = block_cluster_sum<K, false>(part, scratch);
  double* kappa_sh = scratch + ((blockDim.x + 31) / 32 + 2) * K;
  if (threadIdx.x < 32) {
    if constexpr (M <= 2) {
      if (threadIdx.x == 0) {
        double kappa[M];
        and_solve<M>(tot, tot + ND, (double)sc.dt * (double)sc.b, (double)c, kappa);
#pragma unroll
        for (int i = 0; i < M; ++i) kappa_sh[i] = kappa[i];
      }
    } else {
      and_solve_warp<M>(tot, (double)sc.dt * (double)sc.b, (double)c, kappa_sh);
    }
  }
  __syncthreads();
  float w[M];
#pragma unroll
  for (int i = 0; i < M; ++i) w[i] = (float)kappa_sh[i];

  for (int r0 = 0; r0 < nunits; r0 += NV * blockDim.x) {
    float xv[NV][VEC], ev[NV][VEC], sv[M][NV][VEC];
    bool ok[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = r0 + j * blockDim.x + threadIdx.x;
      ok[j] = u < nunits;
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
    for (int j = 0; j < NV; ++j) {
      if (!ok[j]) continue;
      float o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float sm = sv[M - 1][j][e];
        float mix = sm;
#pragma unroll
        for (int i = 0; i < M - 1; ++i) mix = fmaf(w[i], sv[i][j][e] - sm, mix);
        o[e] = xv[j][e] + (-dta * xv[j][e] + 2.f * dtb * mix + c * ev[j][e]);
      }
      const size_t off = base + (size_t)(r0 + j * blockDim.x + threadIdx.x) * VEC;
      if constexpr (VEC == 4) st4(p.x_out + off, make_float4(o[0], o[1], o[2], o[3]));
      else p.x_out[off] = o[0];
    }
  }
  if constexpr (M <= 2) {
    if (threadIdx.x == 0) {
      double kappa[M], R[M];
#pragma unroll
      for (int i = 0; i < M; ++i) kappa[i] = kappa_sh[i];
      and_increments<M>(tot, kappa, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, R);
      write_logq<M>(p, sc, sample, lqs, R);
#pragma unroll
      for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
    }
  } else if (threadIdx.x < 32) {
    // scratch[0..M): the warp partials of the block sum are dead by now
    and_increments_warp<M>(tot, kappa_sh, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, scratch);
    if (threadIdx.x == 0) {
      write_logq<M>(p, sc, sample, lqs, scratch);
#pragma unroll
      for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
    }
  }
}

// 1-D bulk copy global -> shared memory (the TMA engine without a tensor map); completion is counted in bytes on `bar`.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// AND with the sample resident in SHARED memory, for model counts the register file cannot hold (M >= 3 at D = 3072):
// one elected thread issues M+1 bulk copies (noise and the M scores of this sample, D*4 bytes each) that land in shared
// memory asynchronously - every byte of the sample is in flight from the first instruction, with no registers and no
// per-thread load instructions spent on it - while all threads fetch their x slice into registers.  Pass 1 (reductions)
// and pass 2 (mix + Euler-Maruyama update) both read shared memory, so HBM is read exactly once, like the register-
// resident kernel.  (M+1)*D*4 bytes + scratch per CTA: 110.6 KB at M = 8, D = 3072 -> two CTAs per SM.
// Thread t owns units t, t+T, t+2T, ... in that order: the same accumulation order as the register-resident kernel.
constexpr int kAndSmemNVX = 4;          // x units prefetched per thread: D <= 4 * 4 * threads
template <int M>
__global__ void __launch_bounds__(256, (M <= 4 ? 3 : 2)) step_vpsde_and_smem_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int Md = M - 1;
  constexpr int ND = Md * (Md + 1) / 2;
  constexpr int K = ND + 2 * Md + 2;                       // D | E | F | G_MM | N_M
  const int nwarps = (blockDim.x + 31) >> 5;
  const int nunits = p.D / 4;
  float* tile = reinterpret_cast<float*>(smem_raw);        // [noise | s_0 | ... | s_{M-1}], D floats each
  double* scratch = reinterpret_cast<double*>(tile + (size_t)(M + 1) * p.D);
  double* kappa_sh = scratch + (nwarps + 2) * K;
  uint64_t* bar = reinterpret_cast<uint64_t*>(kappa_sh + M);
  const int sample = blockIdx.x;
  const size_t base = (size_t)sample * p.D;
  const uint32_t row_bytes = (uint32_t)p.D * 4u;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, row_bytes * (M + 1));
    bulk_load_1d(tile, p.noise + base, row_bytes, bar);
#pragma unroll
    for (int i = 0; i < M; ++i) bulk_load_1d(tile + (size_t)(i + 1) * p.D, p.s[i] + base, row_bytes, bar);
  }
  const int sched_row = begin_scalars(p);
  LogqState<M> lqs;
  load_logq_state<M>(p, sample, lqs);
  float xv[kAndSmemNVX][4];
#pragma unroll
  for (int j = 0; j < kAndSmemNVX; ++j) {
    const int u = j * blockDim.x + threadIdx.x;
    if (u < nunits) ldv4(xv[j], p.x + base + (size_t)u * 4);
    else { xv[j][0] = 0.f; xv[j][1] = 0.f; xv[j][2] = 0.f; xv[j][3] = 0.f; }
  }
  const StepScalars sc = finish_scalars(p, sched_row);      // x loads and bulk copies are in flight
  const float dta = sc.dt * sc.a, dtb = sc.dt * sc.b;
  const float c = sqrtf(2.f * sc.sigma * sc.b * sc.dt);
  __syncthreads();                       // barrier initialisation visible to every waiting thread
  mbar_wait(bar, 0);

  const float4* t4 = reinterpret_cast<const float4*>(tile);
  float part[K];
#pragma unroll
  for (int k = 0; k < K; ++k) part[k] = 0.f;
  for (int u = threadIdx.x; u < nunits; u += blockDim.x) {
    const float4 e4 = t4[u];
    const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
    float sv[M][4];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const float4 v = t4[(size_t)(i + 1) * nunits + u];
      sv[i][0] = v.x; sv[i][1] = v.y; sv[i][2] = v.z; sv[i][3] = v.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float sm = sv[M - 1][e];
      float d[Md > 0 ? Md : 1];
#pragma unroll
      for (int i = 0; i < Md; ++i) d[i] = sv[i][e] - sm;
      int k = 0;
#pragma unroll
      for (int i = 0; i < Md; ++i)
#pragma unroll
        for (int l = i; l < Md; ++l) { part[k] = fmaf(d[i], d[l], part[k]); ++k; }
#pragma unroll
      for (int i = 0; i < Md; ++i) {
        part[ND + i] = fmaf(d[i], ev[e], part[ND + i]);
        part[ND + Md + i] = fmaf(sm, d[i], part[ND + Md + i]);
      }
      part[ND + 2 * Md] = fmaf(sm, sm, part[ND + 2 * Md]);
      part[ND + 2 * Md + 1] = fmaf(sm, ev[e], part[ND + 2 * Md + 1]);
    }
  }
  const double* tot = block_cluster_sum<K, false>(part, scratch);
  if (threadIdx.x < 32) {
    if constexpr (M <= 2) {
      if (threadIdx.x == 0) {
        double kappa[M];
        and_solve<M>(tot, tot + ND, (double)sc.dt * (double)sc.b, (double)c, kappa);
#pragma unroll
        for (int i = 0; i < M; ++i) kappa_sh[i] = kappa[i];
      }
    } else {
      and_solve_warp<M>(tot, (double)sc.dt * (double)sc.b, (double)c, kappa_sh);
    }
  }
  __syncthreads();
  float w[M];
#pragma unroll
  for (int i = 0; i < M; ++i) w[i] = (float)kappa_sh[i];

#pragma unroll
  for (int j = 0; j < kAndSmemNVX; ++j) {
    const int u = j * blockDim.x + threadIdx.x;
    if (u >= nunits) continue;
    const float4 e4 = t4[u];
    const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
    float mixv[4];
    {
      const float4 v = t4[(size_t)M * nunits + u];          // s_M
      mixv[0] = v.x; mixv[1] = v.y; mixv[2] = v.z; mixv[3] = v.w;
    }
    const float smv[4] = {mixv[0], mixv[1], mixv[2], mixv[3]};
#pragma unroll
    for (int i = 0; i < M - 1; ++i) {
      const float4 v = t4[(size_t)(i + 1) * nunits + u];
      // s_M + sum_j kappa_j (s_j - s_M): the reference's form (superposition_edu.ipynb:942)
      mixv[0] = fmaf(w[i], v.x - smv[0], mixv[0]);
      mixv[1] = fmaf(w[i], v.y - smv[1], mixv[1]);
      mixv[2] = fmaf(w[i], v.z - smv[2], mixv[2]);
      mixv[3] = fmaf(w[i], v.w - smv[3], mixv[3]);
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = xv[j][e] + (-dta * xv[j][e] + 2.f * dtb * mixv[e] + c * ev[e]);
    st4(p.x_out + base + (size_t)u * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
  if constexpr (M <= 2) {
    if (threadIdx.x == 0) {
      double kappa[M], R[M];
#pragma unroll
      for (int i = 0; i < M; ++i) kappa[i] = kappa_sh[i];
      and_increments<M>(tot, kappa, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, R);
      write_logq<M>(p, sc, sample, lqs, R);
#pragma unroll
      for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
    }
  } else if (threadIdx.x < 32) {
    // scratch[0..M): the warp partials of the block sum are dead by now
    and_increments_warp<M>(tot, kappa_sh, (double)sc.dt * (double)sc.b, (double)c, (double)sc.sigma, scratch);
    if (threadIdx.x == 0) {
      write_logq<M>(p, sc, sample, lqs, scratch);
#pragma unroll
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
  LogqState<M> lqs;
  load_logq_state<M>(p, sample, lqs);
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
  write_logq<M>(p, sc, sample, lqs, R);
  if (p.mode != SD_MODE_FIXED) {
#pragma unroll
    for (int i = 0; i < M; ++i) p.weights[(size_t)sample * M + i] = w[i];
  }
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
cudaError_t launch_and_stream(const StepParams& p, int threads, int nv, int vec, cudaStream_t st) {
  constexpr int K = (M - 1) * M / 2 + 2 * (M - 1) + 2;
  const size_t smem = sizeof(double) * ((size_t)(threads / 32 + 2) * K + M);
  if (vec == 4) {
    if (nv == 2) step_vpsde_and_stream_kernel<M, 2, 4><<<p.B, threads, smem, st>>>(p);
    else step_vpsde_and_stream_kernel<M, 1, 4><<<p.B, threads, smem, st>>>(p);
  } else {
    step_vpsde_and_stream_kernel<M, 1, 1><<<p.B, threads, smem, st>>>(p);
  }
  return cudaGetLastError();
}

// bytes of dynamic shared memory of step_vpsde_and_smem_kernel<M>
inline size_t and_smem_bytes(int M, int D, int threads) {
  const int K = (M - 1) * M / 2 + 2 * (M - 1) + 2;
  return (size_t)(M + 1) * D * 4 + sizeof(double) * ((size_t)(threads / 32 + 2) * K + M) + 16;
}

template <int M>
cudaError_t launch_and_smem(const StepParams& p, int threads, cudaStream_t st) {
  const size_t smem = and_smem_bytes(M, p.D, threads);
  if (smem > 48 * 1024) {     // per device and per instantiation, so set it on every launch (host-side only, legal during capture)
    cudaError_t e = cudaFuncSetAttribute(step_vpsde_and_smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  step_vpsde_and_smem_kernel<M><<<p.B, threads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int M>
cudaError_t launch_small(const StepParams& p, cudaStream_t st) {
  const int threads = 128;
  step_vpsde_small_kernel<M><<<(p.B + threads - 1) / threads, threads, 0, st>>>(p);
  return cudaGetLastError();
}


}  // namespace sdb
