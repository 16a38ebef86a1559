// Fused EDM-sigma SuperDiff step on Stable-Diffusion latents (sm_100a).
//
// Replaces reference applications/images/clip_eval.py:395-413 ("and", "or")
// and :417-424 ("avg"): ~25 eager PyTorch elementwise/reduce launches become
// one kernel that reads latents, z, v_obj, v_bg, v_unc once and writes the new
// latents, kappa and the two log-likelihoods.
// One thread-block cluster per sample (D = 16384 for 64x64x4 latents); the
// slice stays in registers between the reduction and the write pass.
// HBM bytes per sample: 4*D*6.
#include "common.cuh"
#include <cstdlib>
#include <cstdio>
#include "../../include/superdiff_b200.h"

namespace sdb {

struct EdmParams {
  const float* x; const float* z; const float* vo; const float* vb; const float* vu;
  float* ll; float* x_out; float* kappa_out;
  int B, D;
  float sigma, dsigma, g, lift_term, temperature, logp, kappa_fixed;
  int mode;
  const float* dlog;     // SD_EDM_MODE_AND_ODE: [B][2] = (dlog_obj, dlog_bg), the Hutchinson divergence estimates (clip_eval.py:103)
};

constexpr int SD_EDM_MODE_AND_ODE = 100;   // internal: deterministic AND (clip_eval.py:377-391), reached through sd_step_edm_ode

// reductions (all per sample):
//  0 dd=<d,d>  1 bd=<base,d>  2 zd=<z,d>  3 oo=<vo,vo>  4 bb=<vb,vb>  5 ob=<vo,base>  6 od=<vo,d>  7 oz=<vo,z>
//  with d = vo - vb, base = vu + g (vb - vu)
template <int NV, bool CLUSTER, int MAXT = 256>
__global__ void __launch_bounds__(MAXT) step_edm_kernel(const __grid_constant__ EdmParams p) {
  extern __shared__ double scratch[];
  unsigned csize = 1, crank = 0;
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    csize = cluster.num_blocks();
    crank = cluster.block_rank();
  }
  const int sample = blockIdx.x / csize;
  const int nunits = p.D / 4;
  const int per_cta = (nunits + csize - 1) / csize;
  const int u0 = crank * per_cta, u1 = min(nunits, u0 + per_cta);
  const size_t base_off = (size_t)sample * p.D;
  const bool ode = p.mode == SD_EDM_MODE_AND_ODE;
  const float cn = ode ? 0.f : sqrtf(2.f * fabsf(p.dsigma) * p.sigma);

  float4 x[NV], z[NV], d[NV], bs[NV], vo[NV];
  bool ok[NV];
  float part[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) part[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int u = u0 + j * blockDim.x + threadIdx.x;
    ok[j] = u < u1;
    x[j] = z[j] = d[j] = bs[j] = vo[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok[j]) {
      const size_t off = base_off + (size_t)u * 4;
      x[j] = ld_stream4(p.x + off);
      if (p.z) z[j] = ld_stream4(p.z + off);
      vo[j] = ld_stream4(p.vo + off);
      const float4 vb = ld_stream4(p.vb + off);
      const float4 vu = ld_stream4(p.vu + off);
      d[j] = make_float4(vo[j].x - vb.x, vo[j].y - vb.y, vo[j].z - vb.z, vo[j].w - vb.w);
      bs[j] = make_float4(vu.x + p.g * (vb.x - vu.x), vu.y + p.g * (vb.y - vu.y),
                          vu.z + p.g * (vb.z - vu.z), vu.w + p.g * (vb.w - vu.w));
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float dd[4] = {d[j].x, d[j].y, d[j].z, d[j].w};
    const float bb[4] = {bs[j].x, bs[j].y, bs[j].z, bs[j].w};
    const float zz[4] = {z[j].x, z[j].y, z[j].z, z[j].w};
    const float oo[4] = {vo[j].x, vo[j].y, vo[j].z, vo[j].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float vbe = oo[e] - dd[e];
      part[0] = fmaf(dd[e], dd[e], part[0]);
      part[1] = fmaf(bb[e], dd[e], part[1]);
      part[2] = fmaf(zz[e], dd[e], part[2]);
      part[3] = fmaf(oo[e], oo[e], part[3]);
      part[4] = fmaf(vbe, vbe, part[4]);
      part[5] = fmaf(oo[e], bb[e], part[5]);
      part[6] = fmaf(oo[e], dd[e], part[6]);
      part[7] = fmaf(oo[e], zz[e], part[7]);
    }
  }
  const double* t = block_cluster_sum<8, CLUSTER>(part, scratch);
  const double DD = t[0], BD = t[1], ZD = t[2], OO = t[3], BB = t[4], OB = t[5], OD = t[6], OZ = t[7];
  const double ds = p.dsigma, sg = p.sigma, g = p.g;
  double kappa;
  if (ode) {
    // clip_eval.py:383-385: kappa = [sigma (dlog_o - dlog_b) + <d, vo + vb> + lift/dsigma*sigma/N - <d, base>] / (g <d,d>)
    const double dl = (double)p.dlog[2 * sample] - (double)p.dlog[2 * sample + 1];
    kappa = (sg * dl + (OO - BB) + (double)p.lift_term - BD) * fast_drcp(g * DD);
  } else if (p.mode == SD_MODE_AND) {
    // clip_eval.py:398-400 with dx_ind = 2 dsigma base + cn z
    const double num = fabs(ds) * (BB - OO) - (2.0 * ds * BD + (double)cn * ZD) + (double)p.lift_term;
    kappa = num * fast_drcp(2.0 * ds * g * DD);
  } else if (p.mode == SD_MODE_OR) {
    // clip_eval.py:402: softmax([T (ll_obj + logp), T ll_bg])[0], fp32 like the reference
    const float z0 = p.temperature * (p.ll[2 * sample] + p.logp), z1 = p.temperature * p.ll[2 * sample + 1];
    const float m = fmaxf(z0, z1);
    const float e0 = expf(z0 - m), e1 = expf(z1 - m);
    kappa = (double)(e0 / (e0 + e1));
  } else {
    kappa = (double)p.kappa_fixed;
  }
  const float kf = (float)kappa;
  const float two_ds = ode ? p.dsigma : 2.f * p.dsigma;      // ODE: latents += dsigma * vf (clip_eval.py:388)
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (!ok[j]) continue;
    float4 o;
    // vf = base + g kappa d ; dx = 2 dsigma vf + cn z   (clip_eval.py:404-405)
    o.x = x[j].x + (two_ds * (bs[j].x + p.g * kf * d[j].x) + cn * z[j].x);
    o.y = x[j].y + (two_ds * (bs[j].y + p.g * kf * d[j].y) + cn * z[j].y);
    o.z = x[j].z + (two_ds * (bs[j].z + p.g * kf * d[j].z) + cn * z[j].z);
    o.w = x[j].w + (two_ds * (bs[j].w + p.g * kf * d[j].w) + cn * z[j].w);
    st4(p.x_out + base_off + (size_t)(u0 + j * blockDim.x + threadIdx.x) * 4, o);
  }
  if (crank == 0 && threadIdx.x == 0 && ode) {
    // clip_eval.py:389-390: ll_k += dsigma (dlog_k - sum (-v_k/sigma)(v_k - vf)),  <vo,vf> = OB + g k OD, <d,vf> = BD + g k DD
    const double o_vf = OB + g * kappa * OD, b_vf = o_vf - (BD + g * kappa * DD);
    p.ll[2 * sample] = p.ll[2 * sample] + (float)(ds * ((double)p.dlog[2 * sample] + (OO - o_vf) * fast_drcp(sg)));
    p.ll[2 * sample + 1] = p.ll[2 * sample + 1] + (float)(ds * ((double)p.dlog[2 * sample + 1] + (BB - b_vf) * fast_drcp(sg)));
    p.kappa_out[sample] = kf;
  } else if (crank == 0 && threadIdx.x == 0) {
    // <vo,dx> = 2ds(OB + g k OD) + cn OZ ; <vb,dx> = <vo,dx> - <d,dx>, <d,dx> = 2ds(BD + g k DD) + cn ZD
    const double o_dx = 2.0 * ds * (OB + g * kappa * OD) + (double)cn * OZ;
    const double d_dx = 2.0 * ds * (BD + g * kappa * DD) + (double)cn * ZD;
    const double b_dx = o_dx - d_dx;
    const double q = (p.mode == SD_MODE_OR) ? (-ds * fast_drcp(sg)) : (-fabs(ds) * fast_drcp(sg));   // :412-413 vs :409-410
    p.ll[2 * sample] = p.ll[2 * sample] + (float)(-o_dx * fast_drcp(sg) + q * OO);
    p.ll[2 * sample + 1] = p.ll[2 * sample + 1] + (float)(-b_dx * fast_drcp(sg) + q * BB);
    p.kappa_out[sample] = kf;
  }
}

// Large batches (B >= 256: enough samples to fill the chip with one CTA each): ONE CTA per sample, streaming.  None of the
// eight reductions depends on kappa (the log-likelihood updates are written in Gram terms), so the modes whose kappa is
// known up front (OR: softmax of the old ll; AVG: fixed) are a single pass - read the five tensors once, accumulate and
// write in the same round.  AND / AND-ODE need <d,d>, <base,d>, ... before the first write: pass 1 reads z, v_obj, v_bg,
// v_unc, the CTA forms kappa, pass 2 re-reads the sample (<= 320 KB that this CTA touched microseconds earlier: L2, not
// HBM) and writes.  No clusters: the cluster-per-sample kernel above ran at 0.55-0.59 of the copy peak at batch 512-2048.
// Old log-likelihoods / divergences are loaded at the start and consumed after the streaming loads (see step_vpsde_kernel).
template <int NV>
__global__ void __launch_bounds__(256) step_edm_stream_kernel(const __grid_constant__ EdmParams p) {
  extern __shared__ double scratch[];
  const int sample = blockIdx.x;
  const int nunits = p.D / 4;
  const size_t base_off = (size_t)sample * p.D;
  const bool ode = p.mode == SD_EDM_MODE_AND_ODE;
  const bool two_pass = ode || p.mode == SD_MODE_AND;
  const float cn = ode ? 0.f : sqrtf(2.f * fabsf(p.dsigma) * p.sigma);
  const float two_ds = ode ? p.dsigma : 2.f * p.dsigma;
  float ll0 = p.ll[2 * sample], ll1 = p.ll[2 * sample + 1];
  float dl0 = 0.f, dl1 = 0.f;
  if (ode) { dl0 = p.dlog[2 * sample]; dl1 = p.dlog[2 * sample + 1]; }
  float part[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) part[k] = 0.f;

  auto accumulate = [&](const float4& z4, const float4& vo4, const float4& vb4, const float4& vu4, float4& d4, float4& bs4) {
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, oo[4] = {vo4.x, vo4.y, vo4.z, vo4.w};
    const float vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w}, vu[4] = {vu4.x, vu4.y, vu4.z, vu4.w};
    float dd[4], bb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dd[e] = oo[e] - vb[e];
      bb[e] = vu[e] + p.g * (vb[e] - vu[e]);
      const float vbe = oo[e] - dd[e];          // as in the resident kernel
      part[0] = fmaf(dd[e], dd[e], part[0]);
      part[1] = fmaf(bb[e], dd[e], part[1]);
      part[2] = fmaf(zz[e], dd[e], part[2]);
      part[3] = fmaf(oo[e], oo[e], part[3]);
      part[4] = fmaf(vbe, vbe, part[4]);
      part[5] = fmaf(oo[e], bb[e], part[5]);
      part[6] = fmaf(oo[e], dd[e], part[6]);
      part[7] = fmaf(oo[e], zz[e], part[7]);
    }
    d4 = make_float4(dd[0], dd[1], dd[2], dd[3]);
    bs4 = make_float4(bb[0], bb[1], bb[2], bb[3]);
  };
  auto update = [&](const float4& x4, const float4& z4, const float4& d4, const float4& bs4, float kf, size_t off) {
    float4 o;
    // vf = base + g kappa d ; dx = 2 dsigma vf + cn z   (clip_eval.py:404-405; ODE: dsigma vf, :388)
    o.x = x4.x + (two_ds * (bs4.x + p.g * kf * d4.x) + cn * z4.x);
    o.y = x4.y + (two_ds * (bs4.y + p.g * kf * d4.y) + cn * z4.y);
    o.z = x4.z + (two_ds * (bs4.z + p.g * kf * d4.z) + cn * z4.z);
    o.w = x4.w + (two_ds * (bs4.w + p.g * kf * d4.w) + cn * z4.w);
    st4(p.x_out + off, o);
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  double kappa = (double)p.kappa_fixed;
  float kf = p.kappa_fixed;

  if (!two_pass) {
    bool have_k = false;
    for (int r0 = 0; r0 < nunits; r0 += NV * blockDim.x) {
      float4 x[NV], z[NV], vo[NV], vb[NV], vu[NV];
      bool ok[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int u = r0 + j * blockDim.x + threadIdx.x;
        ok[j] = u < nunits;
        if (ok[j]) {
          const size_t off = base_off + (size_t)u * 4;
          x[j] = ld_stream4(p.x + off); z[j] = ld_stream4(p.z + off);
          vo[j] = ld_stream4(p.vo + off); vb[j] = ld_stream4(p.vb + off); vu[j] = ld_stream4(p.vu + off);
        }
      }
      if (!have_k) {
        asm volatile("" : "+f"(ll0), "+f"(ll1));        // the softmax stays below the loads just issued
        if (p.mode == SD_MODE_OR) {
          // clip_eval.py:402: softmax([T (ll_obj + logp), T ll_bg])[0], fp32 like the reference
          const float z0 = p.temperature * (ll0 + p.logp), z1 = p.temperature * ll1;
          const float m = fmaxf(z0, z1);
          const float e0 = expf(z0 - m), e1 = expf(z1 - m);
          kappa = (double)(e0 / (e0 + e1));
          kf = (float)kappa;
        }
        have_k = true;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (!ok[j]) continue;
        float4 d4, bs4;
        accumulate(z[j], vo[j], vb[j], vu[j], d4, bs4);
        update(x[j], z[j], d4, bs4, kf, base_off + (size_t)(r0 + j * blockDim.x + threadIdx.x) * 4);
      }
    }
  } else {
    for (int r0 = 0; r0 < nunits; r0 += NV * blockDim.x) {
      float4 z[NV], vo[NV], vb[NV], vu[NV];
      bool ok[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int u = r0 + j * blockDim.x + threadIdx.x;
        ok[j] = u < nunits;
        if (ok[j]) {
          const size_t off = base_off + (size_t)u * 4;
          z[j] = p.z ? ld_stream4(p.z + off) : zero4;
          vo[j] = ld_stream4(p.vo + off); vb[j] = ld_stream4(p.vb + off); vu[j] = ld_stream4(p.vu + off);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (!ok[j]) continue;
        float4 d4, bs4;
        accumulate(z[j], vo[j], vb[j], vu[j], d4, bs4);
      }
    }
  }
  const double* t = block_cluster_sum<8, false>(part, scratch);
  const double DD = t[0], BD = t[1], ZD = t[2], OO = t[3], BB = t[4], OB = t[5], OD = t[6], OZ = t[7];
  const double ds = p.dsigma, sg = p.sigma, g = p.g;
  if (two_pass) {
    if (ode) {
      // clip_eval.py:383-385
      kappa = (sg * ((double)dl0 - (double)dl1) + (OO - BB) + (double)p.lift_term - BD) * fast_drcp(g * DD);
    } else {
      // clip_eval.py:398-400 with dx_ind = 2 dsigma base + cn z
      const double num = fabs(ds) * (BB - OO) - (2.0 * ds * BD + (double)cn * ZD) + (double)p.lift_term;
      kappa = num * fast_drcp(2.0 * ds * g * DD);
    }
    kf = (float)kappa;
    for (int r0 = 0; r0 < nunits; r0 += NV * blockDim.x) {
      float4 x[NV], z[NV], vo[NV], vb[NV], vu[NV];
      bool ok[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int u = r0 + j * blockDim.x + threadIdx.x;
        ok[j] = u < nunits;
        if (ok[j]) {
          const size_t off = base_off + (size_t)u * 4;
          x[j] = ld_stream4(p.x + off);
          z[j] = p.z ? ld_stream4(p.z + off) : zero4;
          vo[j] = ld_stream4(p.vo + off); vb[j] = ld_stream4(p.vb + off); vu[j] = ld_stream4(p.vu + off);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (!ok[j]) continue;
        const float4 d4 = make_float4(vo[j].x - vb[j].x, vo[j].y - vb[j].y, vo[j].z - vb[j].z, vo[j].w - vb[j].w);
        const float4 bs4 = make_float4(vu[j].x + p.g * (vb[j].x - vu[j].x), vu[j].y + p.g * (vb[j].y - vu[j].y),
                                       vu[j].z + p.g * (vb[j].z - vu[j].z), vu[j].w + p.g * (vb[j].w - vu[j].w));
        update(x[j], z[j], d4, bs4, kf, base_off + (size_t)(r0 + j * blockDim.x + threadIdx.x) * 4);
      }
    }
  }
  if (threadIdx.x == 0) {
    if (ode) {
      // clip_eval.py:389-390
      const double o_vf = OB + g * kappa * OD, b_vf = o_vf - (BD + g * kappa * DD);
      p.ll[2 * sample] = ll0 + (float)(ds * ((double)dl0 + (OO - o_vf) * fast_drcp(sg)));
      p.ll[2 * sample + 1] = ll1 + (float)(ds * ((double)dl1 + (BB - b_vf) * fast_drcp(sg)));
    } else {
      const double o_dx = 2.0 * ds * (OB + g * kappa * OD) + (double)cn * OZ;
      const double d_dx = 2.0 * ds * (BD + g * kappa * DD) + (double)cn * ZD;
      const double b_dx = o_dx - d_dx;
      const double q = (p.mode == SD_MODE_OR) ? (-ds * fast_drcp(sg)) : (-fabs(ds) * fast_drcp(sg));   // :412-413 vs :409-410
      p.ll[2 * sample] = ll0 + (float)(-o_dx * fast_drcp(sg) + q * OO);
      p.ll[2 * sample + 1] = ll1 + (float)(-b_dx * fast_drcp(sg) + q * BB);
    }
    p.kappa_out[sample] = kf;
  }
}

// AND / AND-ODE at large batches: kappa needs <d,d>, <base,d>, ... before the first write, and the two-pass streaming form
// above re-reads the sample through L2 (0.54-0.60 of the copy peak).  Here ONE 1024-thread CTA per sample keeps what the
// write pass needs - d = v_obj - v_bg, base = v_unc + g (v_bg - v_unc) and z, 3 x D x 4 bytes = 192 KB at D = 16384 - in
// SHARED memory and x in registers, so HBM is read exactly once (4*D*6 bytes per sample, like the resident kernel) with no
// cluster.  One CTA per SM (the sample fills the shared memory): the reduction and the solve are not overlapped by a
// second CTA, which is what separates it from the single-pass `or` / `avg` form.
constexpr int kEdmSmemThreads = 1024, kEdmSmemNX = 4;      // x units held per thread: D <= 4 * 4 * 1024
__global__ void __launch_bounds__(kEdmSmemThreads, 1) step_edm_smem_kernel(const __grid_constant__ EdmParams p) {
  extern __shared__ __align__(16) unsigned char edm_smem[];
  const int sample = blockIdx.x;
  const int nunits = p.D / 4;
  const size_t base_off = (size_t)sample * p.D;
  const bool ode = p.mode == SD_EDM_MODE_AND_ODE;
  const float cn = ode ? 0.f : sqrtf(2.f * fabsf(p.dsigma) * p.sigma);
  const float two_ds = ode ? p.dsigma : 2.f * p.dsigma;
  float4* d_sh = reinterpret_cast<float4*>(edm_smem);
  float4* bs_sh = d_sh + nunits;
  float4* z_sh = bs_sh + nunits;
  double* scratch = reinterpret_cast<double*>(z_sh + nunits);
  float ll0 = p.ll[2 * sample], ll1 = p.ll[2 * sample + 1];
  float dl0 = 0.f, dl1 = 0.f;
  if (ode) { dl0 = p.dlog[2 * sample]; dl1 = p.dlog[2 * sample + 1]; }
  float part[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) part[k] = 0.f;
  float4 x[kEdmSmemNX];
#pragma unroll
  for (int j = 0; j < kEdmSmemNX; ++j) {
    const int u = j * kEdmSmemThreads + threadIdx.x;
    x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < nunits) {
      const size_t off = base_off + (size_t)u * 4;
      x[j] = ld_stream4(p.x + off);
      const float4 z4 = p.z ? ld_stream4(p.z + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 vo4 = ld_stream4(p.vo + off), vb4 = ld_stream4(p.vb + off), vu4 = ld_stream4(p.vu + off);
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, oo[4] = {vo4.x, vo4.y, vo4.z, vo4.w};
      const float vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w}, vu[4] = {vu4.x, vu4.y, vu4.z, vu4.w};
      float dd[4], bb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dd[e] = oo[e] - vb[e];
        bb[e] = vu[e] + p.g * (vb[e] - vu[e]);
        const float vbe = oo[e] - dd[e];
        part[0] = fmaf(dd[e], dd[e], part[0]);
        part[1] = fmaf(bb[e], dd[e], part[1]);
        part[2] = fmaf(zz[e], dd[e], part[2]);
        part[3] = fmaf(oo[e], oo[e], part[3]);
        part[4] = fmaf(vbe, vbe, part[4]);
        part[5] = fmaf(oo[e], bb[e], part[5]);
        part[6] = fmaf(oo[e], dd[e], part[6]);
        part[7] = fmaf(oo[e], zz[e], part[7]);
      }
      d_sh[u] = make_float4(dd[0], dd[1], dd[2], dd[3]);
      bs_sh[u] = make_float4(bb[0], bb[1], bb[2], bb[3]);
      z_sh[u] = z4;
    }
  }
  const double* t = block_cluster_sum<8, false>(part, scratch);
  const double DD = t[0], BD = t[1], ZD = t[2], OO = t[3], BB = t[4], OB = t[5], OD = t[6], OZ = t[7];
  const double ds = p.dsigma, sg = p.sigma, g = p.g;
  double kappa;
  if (ode) kappa = (sg * ((double)dl0 - (double)dl1) + (OO - BB) + (double)p.lift_term - BD) * fast_drcp(g * DD);        // clip_eval.py:383-385
  else kappa = (fabs(ds) * (BB - OO) - (2.0 * ds * BD + (double)cn * ZD) + (double)p.lift_term) / (2.0 * ds * g * DD);   // :398-400
  const float kf = (float)kappa;
#pragma unroll
  for (int j = 0; j < kEdmSmemNX; ++j) {
    const int u = j * kEdmSmemThreads + threadIdx.x;
    if (u >= nunits) continue;
    const float4 d4 = d_sh[u], b4 = bs_sh[u], z4 = z_sh[u];      // this thread's own writes: no barrier needed
    float4 o;
    o.x = x[j].x + (two_ds * (b4.x + p.g * kf * d4.x) + cn * z4.x);
    o.y = x[j].y + (two_ds * (b4.y + p.g * kf * d4.y) + cn * z4.y);
    o.z = x[j].z + (two_ds * (b4.z + p.g * kf * d4.z) + cn * z4.z);
    o.w = x[j].w + (two_ds * (b4.w + p.g * kf * d4.w) + cn * z4.w);
    st4(p.x_out + base_off + (size_t)u * 4, o);
  }
  if (threadIdx.x == 0) {
    if (ode) {
      const double o_vf = OB + g * kappa * OD, b_vf = o_vf - (BD + g * kappa * DD);     // clip_eval.py:389-390
      p.ll[2 * sample] = ll0 + (float)(ds * ((double)dl0 + (OO - o_vf) * fast_drcp(sg)));
      p.ll[2 * sample + 1] = ll1 + (float)(ds * ((double)dl1 + (BB - b_vf) * fast_drcp(sg)));
    } else {
      const double o_dx = 2.0 * ds * (OB + g * kappa * OD) + (double)cn * OZ;
      const double d_dx = 2.0 * ds * (BD + g * kappa * DD) + (double)cn * ZD;
      const double b_dx = o_dx - d_dx;
      const double q = -fabs(ds) / sg;                                                     // :409-410
      p.ll[2 * sample] = ll0 + (float)(-o_dx * fast_drcp(sg) + q * OO);
      p.ll[2 * sample + 1] = ll1 + (float)(-b_dx * fast_drcp(sg) + q * BB);
    }
    p.kappa_out[sample] = kf;
  }
}

template <int NV>
static cudaError_t launch_edm(const EdmParams& p, int threads, int cluster, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.B * cluster);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = sizeof(double) * (size_t)(threads / 32 + 2) * 8;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (cluster > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (threads > 512) {
      if constexpr (NV <= 2) return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, true, 1024>, p);
      else return cudaErrorInvalidConfiguration;          // 4 float4 per thread x 5 arrays do not fit 64 registers
    }
    if (threads > 256) return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, true, 512>, p);
    return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, true>, p);
  }
  if (threads > 512) {
    if constexpr (NV <= 2) return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, false, 1024>, p);
    else return cudaErrorInvalidConfiguration;
  }
  if (threads > 256) return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, false, 512>, p);
  return cudaLaunchKernelEx(&cfg, step_edm_kernel<NV, false>, p);
}

}  // namespace sdb

static int step_edm_impl(const float* latents, const float* z, const float* v_obj, const float* v_bg,
                         const float* v_unc, int B, int D, float sigma, float dsigma, float guidance,
                         float lift_term, int mode, float temperature, float logp, float kappa_fixed, float* ll,
                         float* latents_out, float* kappa_out, const float* dlog, void* stream) {
  using namespace sdb;
  const bool ode = mode == SD_EDM_MODE_AND_ODE;
  if (B == 0) return SD_OK;
  if (!latents || (!z && !ode) || !v_obj || !v_bg || !v_unc || !ll || !latents_out || !kappa_out || (ode && !dlog))
    return fail(kErrInvalidArg, "sd_step_edm_cfg: null pointer argument");
  if (B < 0 || D < 4 || D % 4) return fail(kErrInvalidArg, "sd_step_edm_cfg: D must be a positive multiple of 4");
  if (!(mode == SD_MODE_AND || mode == SD_MODE_OR || mode == SD_MODE_AVG || ode))
    return fail(kErrInvalidArg, "sd_step_edm_cfg: mode must be AND, OR or AVG");
  if (!(sigma > 0.f)) return fail(kErrInvalidArg, "sd_step_edm_cfg: sigma must be > 0");
  if ((((uintptr_t)latents | (uintptr_t)(z ? z : latents) | (uintptr_t)v_obj | (uintptr_t)v_bg | (uintptr_t)v_unc |
        (uintptr_t)latents_out) % 16) != 0)
    return fail(kErrInvalidArg, "sd_step_edm_cfg: tensors must be 16-byte aligned");
  EdmParams p{latents, z, v_obj, v_bg, v_unc, ll, latents_out, kappa_out, B, D,
              sigma, dsigma, guidance, lift_term, temperature, logp, kappa_fixed, mode, dlog};
  const int nunits = D / 4;
  // enough samples to fill the chip with one CTA each: the streaming kernel (SDB_EDM_STREAM_MIN_B overrides the threshold)
  static const int stream_min_b = [] { const char* e = getenv("SDB_EDM_STREAM_MIN_B"); return e && *e ? atoi(e) : 256; }();
  // AND / AND-ODE with one sample per SM resident in shared memory (batch >= 148: a CTA for every SM)
  static const int smem_and = [] { const char* e = getenv("SDB_EDM_SMEM_AND"); return e && *e ? atoi(e) : 1; }();   // tuning knob
  const size_t smem_need = (size_t)D * 12 + sizeof(double) * (size_t)(kEdmSmemThreads / 32 + 2) * 8;
  if (smem_and && (ode || mode == SD_MODE_AND) && B >= 148 && nunits <= kEdmSmemNX * kEdmSmemThreads && smem_need <= 227 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(step_edm_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need);
    if (e != cudaSuccess) return check_cuda(e, "sd_step_edm_cfg (shared-memory AND)");
    step_edm_smem_kernel<<<B, kEdmSmemThreads, smem_need, (cudaStream_t)stream>>>(p);
    return check_cuda(cudaGetLastError(), "sd_step_edm_cfg launch (shared-memory AND)");
  }
  if (B >= stream_min_b) {
    const size_t smem = sizeof(double) * (size_t)(256 / 32 + 2) * 8;
    static const int stream_nv = [] { const char* e = getenv("SDB_EDM_STREAM_NV"); return e && *e ? atoi(e) : 2; }();   // tuning knob: 2 measured 3-5 % ahead of 1 at batch 512
    if (stream_nv == 2 && nunits >= 512) step_edm_stream_kernel<2><<<B, 256, smem, (cudaStream_t)stream>>>(p);
    else step_edm_stream_kernel<1><<<B, 256, smem, (cudaStream_t)stream>>>(p);
    return check_cuda(cudaGetLastError(), "sd_step_edm_cfg launch (streaming)");
  }
  // smallest (cluster, threads, NV <= 2) that keeps the sample resident; prefer more CTAs when B is small
  int best_c = 0, best_t = 0, best_nv = 0;
  long best_cost = -1;
  for (int c = 1; c <= 8; c *= 2)
    for (int t = 64; t <= 256; t *= 2)
      for (int n = 1; n <= 2; ++n) {
        const long cap = (long)c * t * n;
        if (cap < nunits) continue;
        long cost = (cap - nunits) * 4 + (c > 1 ? 64 * c : 0);
        const long ctas = (long)B * c;
        if (ctas < 148 * 4) cost += (148 * 4 - ctas);
        if (t < 128) cost += 32;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = c; best_t = t; best_nv = n; }
      }
  if (best_cost < 0) return fail(kErrUnsupported, "sd_step_edm_cfg: D too large (max 8*256*2 float4 per sample)");
  // few samples (BASELINE config 4: batch 64): clusters of 4 CTAs with four float4 per thread instead of clusters of 8 with two --
  // half as many CTAs to schedule, the same bytes in flight per SM: 10.9 -> 8.6 us at batch 64; at batch 96 the order flips
  // (12.7 vs 14.5 us), so only while the grid stays within two CTAs per SM (profiles/r04j_edm_shapes.txt)
  if (best_c == 8 && best_t == 256 && best_nv == 2 && (long)B * 4 <= 2L * 148 && (long)4 * 256 * 4 >= nunits) { best_c = 4; best_nv = 4; }
  // SDB_EDM_SHAPE="cluster,threads,nv" forces a launch shape (tools/edm_sweep.py --shape); it must hold the sample
  static const char* shape_env = getenv("SDB_EDM_SHAPE");
  if (shape_env && *shape_env) {
    int c = 0, t = 0, n = 0;
    if (sscanf(shape_env, "%d,%d,%d", &c, &t, &n) == 3 && (c == 1 || c == 2 || c == 4 || c == 8) && t >= 64 && t <= 1024 && (t % 32) == 0 &&
        (n == 1 || n == 2 || n == 4) && (long)c * t * n >= nunits && !(n == 4 && t > 512)) {
      best_c = c; best_t = t; best_nv = n;
    }
  }
  cudaError_t err = best_nv == 1 ? launch_edm<1>(p, best_t, best_c, (cudaStream_t)stream)
                  : best_nv == 2 ? launch_edm<2>(p, best_t, best_c, (cudaStream_t)stream)
                                 : launch_edm<4>(p, best_t, best_c, (cudaStream_t)stream);
  return check_cuda(err, "sd_step_edm_cfg launch");
}

extern "C" int sd_step_edm_cfg(const float* latents, const float* z, const float* v_obj, const float* v_bg,
                               const float* v_unc, int B, int D, float sigma, float dsigma, float guidance,
                               float lift_term, int mode, float temperature, float logp, float kappa_fixed, float* ll,
                               float* latents_out, float* kappa_out, void* stream) {
  if (!(mode == SD_MODE_AND || mode == SD_MODE_OR || mode == SD_MODE_AVG))
    return sdb::fail(sdb::kErrInvalidArg, "sd_step_edm_cfg: mode must be AND, OR or AVG");
  return step_edm_impl(latents, z, v_obj, v_bg, v_unc, B, D, sigma, dsigma, guidance, lift_term, mode, temperature, logp,
                       kappa_fixed, ll, latents_out, kappa_out, nullptr, stream);
}

extern "C" int sd_step_edm_ode(const float* latents, const float* v_obj, const float* v_bg, const float* v_unc,
                               const float* dlog, int B, int D, float sigma, float dsigma, float guidance,
                               float lift_term, float* ll, float* latents_out, float* kappa_out, void* stream) {
  return step_edm_impl(latents, nullptr, v_obj, v_bg, v_unc, B, D, sigma, dsigma, guidance, lift_term,
                       sdb::SD_EDM_MODE_AND_ODE, 1.f, 0.f, 0.5f, ll, latents_out, kappa_out, dlog, stream);
}
