// HBM-bound glue between the tap-GEMMs: BatchNorm statistics / finalize, the fused
// "BN-apply + PReLU + circular phase shift + reflect halo" producer of conv inputs and its
// backward, layout converters, the discriminator FC tail, and the loss tails.
// All 16-bit tensors are NLC ([positions][C], C innermost); threads own 8 consecutive channels
// (one 16-byte vector) so every access is a coalesced 128-bit transaction.
#include "common.cuh"

namespace sg {

__device__ __forceinline__ V8 ldv8(const void* p, int64_t elem_off) {
  return *reinterpret_cast<const V8*>(reinterpret_cast<const uint16_t*>(p) + elem_off);
}
__device__ __forceinline__ void stv8(void* p, int64_t elem_off, const V8& v) {
  *reinterpret_cast<V8*>(reinterpret_cast<uint16_t*>(p) + elem_off) = v;
}

// block-level reduction of per-thread 8-channel partials: threads with the same channel group
// (tid % cgs) are summed, then one double atomic per channel.
template <int NS>
__device__ __forceinline__ void block_reduce_channels(float (&part)[NS][8], int cgs, int C, double* out,
                                                      float* smem /* [256][8] */) {
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  for (int s = 0; s < NS; ++s) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) smem[tid * 8 + j] = part[s][j];
    __syncthreads();
    if (tid < cgs) {
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int t = tid; t < blockDim.x; t += cgs)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += (double)smem[t * 8 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(out + (int64_t)s * C + cg * 8 + j, acc[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Streaming kernels.  Thread = 4 adjacent channels (one 8-byte load per stream) of a row, C/4
// threads per row, 256/(C/4) rows per CTA iteration, two rows in flight per thread.  Measured
// trade-off on B200: 2 channels/thread is instruction-bound (index math per element), 8 channels/
// thread with all per-channel constants in registers drops to 1 CTA/SM; 4 channels keeps ~64
// registers (4 CTAs/SM) with ~15 instructions per element.
// ------------------------------------------------------------------------------------------
constexpr int VEC = 4;
constexpr int EW_UNROLL = 2;
// per-channel statistics are accumulated into SG_STAT_SLICES interleaved copies (slice = CTA % 8):
// ~450 CTAs hitting one fp64 address serialise at ~60 ns each (measured: ~30 us tail per launch);
// consumers add the slices up.
constexpr int SL = SG_STAT_SLICES;

struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const void* p, int64_t elem, int dtype) {
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + elem);
  F4 r;
  if (dtype == SG_F16) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  } else {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  }
  return r;
}
__device__ __forceinline__ void st4(void* p, int64_t elem, const float (&x)[4], int dtype) {
  uint2 u;
  if (dtype == SG_F16) {
    __half2 a = __floats2half2_rn(x[0], x[1]), b = __floats2half2_rn(x[2], x[3]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  }
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + elem) = u;
}

// per-thread partial sums of NS statistics for 4 channels -> smem combine over the CTA's threads
// that own the same channels -> one double atomic per channel and CTA
template <int NS>
__device__ __forceinline__ void block_stats_flush(float (&part)[NS][VEC], int cgs, int C, double* out,
                                                  float* smem /* [256][VEC] */) {
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  for (int s = 0; s < NS; ++s) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VEC; ++j) smem[tid * VEC + j] = part[s][j];
    __syncthreads();
    if (tid < cgs) {
      double acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = 0;
      for (int t = tid; t < (int)blockDim.x; t += cgs)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += (double)smem[t * VEC + j];
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        atomicAdd(out + ((int64_t)(blockIdx.x % SL) * NS + s) * C + cg * VEC + j, acc[j]);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_stats_kernel(const void* __restrict__ a, int dtype, int64_t rows64, int C, double* __restrict__ stats) {
  __shared__ float red[256 * VEC];
  const int cgs = C / VEC;
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  const int rpb = 256 / cgs;
  const int rows = (int)rows64;
  const int stride = gridDim.x * rpb;
  float part[2][VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { part[0][j] = 0.f; part[1][j] = 0.f; }
  for (int r0 = blockIdx.x * rpb + tid / cgs; r0 < rows; r0 += 4 * stride) {
    F4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * stride;
      if (r < rows) v[u] = ld4(a, (int64_t)r * C + cg * VEC, dtype);
      else { v[u].v[0] = v[u].v[1] = v[u].v[2] = v[u].v[3] = 0.f; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        part[0][j] += v[u].v[j];
        part[1][j] = fmaf(v[u].v[j], v[u].v[j], part[1][j]);
      }
  }
  block_stats_flush<2>(part, cgs, C, stats, red);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0, s1 = 0;
  for (int i = 0; i < SL; ++i) { s0 += stats[(int64_t)i * 2 * C + c]; s1 += stats[(int64_t)i * 2 * C + C + c]; }
  const double mean = s0 / count;
  double var = s1 / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - (float)mean * sc;
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = invstd;
  if (rmean) {
    const double unbiased = count > 1 ? var * count / (count - 1.0) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
  }
}

// ------------------------------------------------------------------------------------------
// h[b][q + H][c] = act(a[b][src(q)][c] * scale + shift),  q in [-H, L + H),
// src(q) = unroll(reflect(q))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
act_fwd_kernel(const void* __restrict__ a, int dtype, int batch, int L, int C,
               const float* __restrict__ scale_shift, const float* __restrict__ slope, int act, int roll, int H,
               void* __restrict__ h, void* __restrict__ h_bf16, void* __restrict__ a_bf16) {
  const int cgs = C / VEC;
  const int Lh = L + 2 * H;
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  const int rpb = 256 / cgs;
  const int rows = batch * Lh;           // output rows (incl. halo)
  const int stride = gridDim.x * rpb;
  float sc[VEC], sh[VEC], sl[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c = cg * VEC + j;
    sc[j] = scale_shift ? scale_shift[c] : 1.f;
    sh[j] = scale_shift ? scale_shift[C + c] : 0.f;
    sl[j] = (act == SG_ACT_PRELU) ? slope[c] : 1.f;
  }
  for (int r0 = blockIdx.x * rpb + tid / cgs; r0 < rows; r0 += EW_UNROLL * stride) {
    F4 v[EW_UNROLL];
    int srcs[EW_UNROLL], qhs[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * stride;
      srcs[u] = 0; qhs[u] = 0;
      if (r < rows) {
        const int b = r / Lh;
        qhs[u] = r - b * Lh;
        srcs[u] = b * L + unroll_idx(reflect_idx(qhs[u] - H, L), roll, L);
        v[u] = ld4(a, (int64_t)srcs[u] * C + cg * VEC, dtype);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * stride;
      if (r < rows) {
        float y[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          y[j] = fmaf(v[u].v[j], sc[j], sh[j]);
          if (act == SG_ACT_PRELU) y[j] = y[j] > 0.f ? y[j] : sl[j] * y[j];
        }
        st4(h, (int64_t)r * C + cg * VEC, y, dtype);
        // bf16 twins: operands of the weight-gradient tap-GEMM (tcgen05 kind::f16 cannot mix f16 x bf16)
        if (h_bf16) st4(h_bf16, (int64_t)r * C + cg * VEC, y, SG_BF16);
        if (a_bf16 && qhs[u] >= H && qhs[u] < H + L) st4(a_bf16, (int64_t)srcs[u] * C + cg * VEC, v[u].v, SG_BF16);
      }
    }
  }
}

// gradient w.r.t. the activation output at exact position l: the consumer-view gradient at the
// rolled position plus its reflect-halo mirrors
__device__ __forceinline__ F4 gather_gy(const void* g_h, int ldh, int H, int roll, int b, int l, int L, int c) {
  F4 g;
  g.v[0] = g.v[1] = g.v[2] = g.v[3] = 0.f;
  if (g_h) {
    const int Lh = L + 2 * H;
    int q0 = l + roll;
    if (q0 >= L) q0 -= L;
    if (q0 < 0) q0 += L;
    const int64_t base = (int64_t)b * Lh + H;
    g = ld4(g_h, (base + q0) * ldh + c, SG_BF16);
    if (H > 0) {
      if (q0 >= 1 && q0 <= H) {
        const F4 m = ld4(g_h, (base - q0) * ldh + c, SG_BF16);
#pragma unroll
        for (int j = 0; j < 4; ++j) g.v[j] += m.v[j];
      }
      if (q0 >= L - 1 - H && q0 <= L - 2) {
        const F4 m = ld4(g_h, (base + 2 * (L - 1) - q0) * ldh + c, SG_BF16);
#pragma unroll
        for (int j = 0; j < 4; ++j) g.v[j] += m.v[j];
      }
    }
  }
  return g;
}

// MODE 0: reductions (and, when g_a_out != null, g_pre written in the same pass: final without BN)
// MODE 1: apply (BN backward) using the reductions
// The skip-connection gradient g_add is w.r.t. the PRE-activation (generator.py:185,191) and joins
// after the activation derivative.
template <int MODE>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const void* __restrict__ g_h, int ldh, int H, int roll, const void* __restrict__ g_add, int lda,
               const void* __restrict__ a, int dtype, int batch, int L, int C,
               const float* __restrict__ scale_shift, const float* __restrict__ mean_invstd,
               const float* __restrict__ slope, int act, double* __restrict__ red, int use_bn,
               void* __restrict__ g_a_out) {
  __shared__ float sred[256 * VEC];
  const int cgs = C / VEC;
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  const int c0 = cg * VEC;
  const int rpb = 256 / cgs;
  const int rows = batch * L;
  const int stride = gridDim.x * rpb;
  // per-channel constants folded as far as possible:
  //   y = x*sc + sh (sign only) ; ahat = x*is - mi ; MODE 1: ga = sc*gpre + ka*x + kb
  float sc[VEC], sh[VEC], sl[VEC], p0[VEC], p1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c = c0 + j;
    sc[j] = scale_shift ? scale_shift[c] : 1.f;
    sh[j] = scale_shift ? scale_shift[C + c] : 0.f;
    sl[j] = (act == SG_ACT_PRELU) ? slope[c] : 1.f;
    const float mu = mean_invstd ? mean_invstd[c] : 0.f;
    const float is = mean_invstd ? mean_invstd[C + c] : 1.f;
    if (MODE == 0) {
      p0[j] = is;
      p1[j] = mu * is;
    } else {
      // MODE 1 receives the slice-SUMMED statistics [3][C] (the caller adds the slices of pass 1)
      const float r1 = (float)(red[C + c] / (double)rows);
      const float r2 = (float)(red[2 * C + c] / (double)rows);
      p0[j] = use_bn ? -sc[j] * r2 * is : 0.f;                       // ka
      p1[j] = use_bn ? sc[j] * (r2 * is * mu - r1) : 0.f;           // kb
    }
  }
  float part[3][VEC];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[s][j] = 0.f;
  for (int r0 = blockIdx.x * rpb + tid / cgs; r0 < rows; r0 += EW_UNROLL * stride) {
    F4 gy[EW_UNROLL], gs[EW_UNROLL], av[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * stride;
#pragma unroll
      for (int j = 0; j < 4; ++j) gs[u].v[j] = 0.f;
      if (r < rows) {
        const int b = r / L, l = r - b * L;
        gy[u] = gather_gy(g_h, ldh, H, roll, b, l, L, c0);
        if (g_add) gs[u] = ld4(g_add, (int64_t)r * lda + c0, SG_BF16);
        av[u] = ld4(a, (int64_t)r * C + c0, dtype);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * stride;
      if (r < rows) {
        float out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float x = av[u].v[j];
          const float y = fmaf(x, sc[j], sh[j]);
          const float g = gy[u].v[j];
          float gpre = g;
          if (act == SG_ACT_PRELU && y <= 0.f) {
            if (MODE == 0) part[0][j] = fmaf(g, y, part[0][j]);
            gpre = g * sl[j];
          }
          gpre += gs[u].v[j];
          if (MODE == 0) {
            part[1][j] += gpre;
            part[2][j] = fmaf(gpre, fmaf(x, p0[j], -p1[j]), part[2][j]);
            out[j] = gpre;
          } else {
            out[j] = use_bn ? fmaf(sc[j], gpre, fmaf(p0[j], x, p1[j])) : gpre;
          }
        }
        if (g_a_out) st4(g_a_out, (int64_t)r * C + c0, out, SG_BF16);
      }
    }
  }
  if (MODE == 0) block_stats_flush<3>(part, cgs, C, red, sred);
}

// ------------------------------------------------------------------------------------------
// layout converters (32 x 32 smem transpose tiles)
// ------------------------------------------------------------------------------------------
__global__ void ncl_to_nlc_kernel(const float* __restrict__ src, int C, int L, void* __restrict__ dst, int dtype) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    tile[i][tx] = (c < C && l < L) ? src[((int64_t)b * C + c) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    if (c < C && l < L) st16(dst, ((int64_t)b * L + l) * C + c, tile[tx][i], dtype);
  }
}
__global__ void nlc_to_ncl_kernel(const void* __restrict__ src, int dtype, int C, int L, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    tile[i][tx] = (c < C && l < L) ? ld16(src, ((int64_t)b * L + l) * C + c, dtype) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    if (c < C && l < L) dst[((int64_t)b * C + c) * L + l] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ a, int dtype, int64_t rows64, int C, double* __restrict__ tmp) {
  __shared__ float red[256 * VEC];
  const int cgs = C / VEC;
  const int tid = threadIdx.x;
  const int cg = tid % cgs;
  const int rpb = 256 / cgs;
  const int rows = (int)rows64;
  const int stride = gridDim.x * rpb;
  float part[1][VEC] = {{0.f, 0.f, 0.f, 0.f}};
  for (int r = blockIdx.x * rpb + tid / cgs; r < rows; r += stride) {
    const F4 v = ld4(a, (int64_t)r * C + cg * VEC, dtype);
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[0][j] += v.v[j];
  }
  block_stats_flush<1>(part, cgs, C, tmp, red);
}
__global__ void colsum_fold_kernel(const double* __restrict__ tmp, int C, int mod, float* __restrict__ out,
                                   int accumulate) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= mod) return;
  double s = 0;
  for (int i = 0; i < SL; ++i)
    for (int c = m; c < C; c += mod) s += tmp[(int64_t)i * C + c];
  out[m] = (accumulate ? out[m] : 0.f) + (float)s;
}

// ------------------------------------------------------------------------------------------
// Discriminator head after fc.0 (discriminator.py:111-117)
// ------------------------------------------------------------------------------------------
constexpr int FC1 = 256, FC2 = 128;

__global__ void __launch_bounds__(256)
fc_tail_fwd_kernel(const float* __restrict__ fc0_acc, const float* __restrict__ b0, const float* __restrict__ s1,
                   const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ s3,
                   const float* __restrict__ w4, const float* __restrict__ b4, float* __restrict__ z1,
                   float* __restrict__ z2, float* __restrict__ logit) {
  __shared__ float h1[FC1];
  __shared__ float h2[FC2];
  __shared__ float wred[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  {
    const float z = fc0_acc[(int64_t)b * FC1 + tid] + b0[tid];
    z1[(int64_t)b * FC1 + tid] = z;
    h1[tid] = z > 0.f ? z : s1[tid] * z;
  }
  __syncthreads();
  // z2[j] = b2[j] + sum_i w2[j][i] h1[i] : one warp per 16 outputs, lanes stride the 256 inputs
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp * 16; j < warp * 16 + 16; ++j) {
    float s = 0.f;
    for (int i = lane; i < FC1; i += 32) s = fmaf(w2[j * FC1 + i], h1[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float z = s + b2[j];
      z2[(int64_t)b * FC2 + j] = z;
      h2[j] = z > 0.f ? z : s3[j] * z;
    }
  }
  __syncthreads();
  float s = tid < FC2 ? w4[tid] * h2[tid] : 0.f;
  s = warp_sum(s);
  if (lane == 0) wred[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float t = b4[0];
    for (int i = 0; i < 8; ++i) t += wred[i];
    logit[b] = t;
  }
}

// per-row backward: g_z2 [B][128], g_z1 [B][256], g_h1 [B][256] (fp32 workspaces) + bf16 copy of g_z1
__global__ void __launch_bounds__(256)
fc_tail_bwd_rows_kernel(const float* __restrict__ z1, const float* __restrict__ z2, const float* __restrict__ logit,
                        const float* __restrict__ g_logit_in, float target, float weight,
                        const float* __restrict__ s1, const float* __restrict__ w2,
                        const float* __restrict__ s3, const float* __restrict__ w4, int batch,
                        float* __restrict__ loss_out, float* __restrict__ g_logit_ws, float* __restrict__ g_z2_ws,
                        float* __restrict__ g_z1_ws, float* __restrict__ g_h1_ws, void* __restrict__ g_z1_bf16) {
  __shared__ float gz2[FC2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float diff = logit[b] - target;
  const float gl = g_logit_in ? g_logit_in[b] : 2.f * diff / (float)batch * weight;
  if (tid == 0) {
    g_logit_ws[b] = gl;
    if (loss_out) atomicAdd(loss_out, diff * diff / (float)batch * weight);
  }
  if (tid < FC2) {
    const float z = z2[(int64_t)b * FC2 + tid];
    const float gh2 = gl * w4[tid];
    const float g = z > 0.f ? gh2 : gh2 * s3[tid];
    gz2[tid] = g;
    g_z2_ws[(int64_t)b * FC2 + tid] = g;
  }
  __syncthreads();
  float gh1 = 0.f;
  for (int j = 0; j < FC2; ++j) gh1 = fmaf(gz2[j], w2[j * FC1 + tid], gh1);
  const float z = z1[(int64_t)b * FC1 + tid];
  const float g = z > 0.f ? gh1 : gh1 * s1[tid];
  g_z1_ws[(int64_t)b * FC1 + tid] = g;
  g_h1_ws[(int64_t)b * FC1 + tid] = gh1;
  st16(g_z1_bf16, (int64_t)b * FC1 + tid, g, SG_BF16);
}

// parameter gradients of the head: blockIdx.y = chunk of 16 batch rows, one thread per output
// element, partial sums merged with one atomic per (element, chunk)
constexpr int FC_CHUNK = 16;
__global__ void __launch_bounds__(256)
fc_tail_bwd_params_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                          const float* __restrict__ g_logit, const float* __restrict__ g_z2,
                          const float* __restrict__ g_z1, const float* __restrict__ g_h1,
                          const float* __restrict__ s1, const float* __restrict__ s3,
                          const float* __restrict__ w4, int batch, float* __restrict__ g_b0, float* __restrict__ g_s1, float* __restrict__ g_w2,
                          float* __restrict__ g_b2, float* __restrict__ g_s3, float* __restrict__ g_w4,
                          float* __restrict__ g_b4) {
  __shared__ float h1s[FC_CHUNK][FC1];
  __shared__ float gz2s[FC_CHUNK][FC2];
  const int b0 = blockIdx.y * FC_CHUNK;
  const int nb = min(FC_CHUNK, batch - b0);
  const int tid = threadIdx.x;
  for (int i = tid; i < nb * FC1; i += 256) {
    const int r = i / FC1, c = i % FC1;
    const float z = z1[(int64_t)(b0 + r) * FC1 + c];
    h1s[r][c] = z > 0.f ? z : s1[c] * z;
  }
  for (int i = tid; i < nb * FC2; i += 256) gz2s[i / FC2][i % FC2] = g_z2[(int64_t)(b0 + i / FC2) * FC2 + i % FC2];
  __syncthreads();
  if (blockIdx.x < FC2) {                  // g_w2 row j = blockIdx.x, column i = tid
    const int j = blockIdx.x, i = tid;
    float s = 0.f;
    for (int r = 0; r < nb; ++r) s = fmaf(gz2s[r][j], h1s[r][i], s);
    atomicAdd(g_w2 + j * FC1 + i, s);
    return;
  }
  // last x-block: the vector gradients
  {
    const int i = tid;                     // FC1 outputs: g_b0, g_s1
    float sb = 0.f, ss = 0.f;
    for (int r = 0; r < nb; ++r) {
      const int64_t o = (int64_t)(b0 + r) * FC1 + i;
      sb += g_z1[o];
      const float z = z1[o];
      if (z <= 0.f) ss = fmaf(g_h1[o], z, ss);
    }
    atomicAdd(g_b0 + i, sb);
    atomicAdd(g_s1 + i, ss);
  }
  if (tid < FC2) {
    const int j = tid;
    float sb = 0.f, ss = 0.f, sw = 0.f;
    const float sl = s3[j];
    for (int r = 0; r < nb; ++r) {
      const float z = z2[(int64_t)(b0 + r) * FC2 + j];
      const float gl = g_logit[b0 + r];
      const float g = gz2s[r][j];
      sb += g;
      const float h = z > 0.f ? z : sl * z;
      sw = fmaf(gl, h, sw);
      if (z <= 0.f) ss = fmaf(gl * w4[j], z, ss);        // d s3 = sum g_h2 * z [z<=0], g_h2 = g_logit*w4
    }
    atomicAdd(g_b2 + j, sb);
    atomicAdd(g_s3 + j, ss);
    atomicAdd(g_w4 + j, sw);
  }
  if (tid == 0) {
    float s = 0.f;
    for (int r = 0; r < nb; ++r) s += g_logit[b0 + r];
    atomicAdd(g_b4, s);
  }
}

__global__ void l1_loss_bwd_kernel(const float* __restrict__ y, const float* __restrict__ clean, int64_t n,
                                   float weight, float* __restrict__ loss_out, float* __restrict__ gy,
                                   int accumulate) {
  float s = 0.f;
  const float gscale = weight / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = y[i] - clean[i];
    s += fabsf(d);
    const float g = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
    if (gy) gy[i] = accumulate ? gy[i] + g : g;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && loss_out) atomicAdd(loss_out, s * gscale);
}

static inline int ew_grid(int64_t work_items, int per_block, int cap_per_sm = 16) {
  int64_t g = cdiv(work_items, per_block);
  const int64_t cap = (int64_t)cap_per_sm * NUM_SMS;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
// kernels that end in a per-block reduction (smem + one double atomic per channel and block):
// keep the grid at 2 CTAs/SM so the same-address atomics stay in the hundreds, not thousands
constexpr int RED_CAP = 3;
// grid for the warp-per-64-channel-chunk streaming kernels: multiple of 4 so that warps/chunks is integral
static inline int stream_grid(int64_t rows, int C, int cap_per_sm) {
  const int rpb = 256 / (C / VEC);
  int64_t g = cdiv(rows, (int64_t)rpb * EW_UNROLL);
  const int64_t cap = (int64_t)cap_per_sm * NUM_SMS;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace sg

using namespace sg;

#define ST ((cudaStream_t)stream)

extern "C" int sg_bn_stats(const void* a, int dtype, int64_t rows_total, int C, double* stats, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 1024 && a && stats && rows_total < (1ll << 31));
  bn_stats_kernel<<<stream_grid(rows_total, C, RED_CAP), 256, 0, ST>>>(a, dtype, rows_total, C, stats);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_bn_finalize(const double* stats, int64_t count, int C, const float* gamma, const float* beta,
                              float eps, float momentum, float* running_mean, float* running_var,
                              float* scale_shift, float* mean_invstd, void* stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(stats, (double)count, C, gamma, beta, eps, momentum,
                                                      running_mean, running_var, scale_shift, mean_invstd);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_act_fwd(const void* a, int dtype, int batch, int L, int C, const float* scale_shift,
                          const float* slope, int act, int roll, int out_halo_pos, void* h, void* h_bf16,
                          void* a_bf16, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 1024 && (out_halo_pos == 0 || L >= 32));
  SG_CHECK_ARG(act == SG_ACT_NONE || (act == SG_ACT_PRELU && slope));
  act_fwd_kernel<<<stream_grid((int64_t)batch * (L + 2 * out_halo_pos), C, 16), 256, 0, ST>>>(a, dtype, batch, L, C, scale_shift, slope, act, roll,
                                                          out_halo_pos, h, h_bf16, a_bf16);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_act_bwd_reduce(const void* g_h, int g_h_ld, int in_halo_pos, int roll, const void* g_add,
                                 int g_add_ld, const void* a,
                                 int dtype, int batch, int L, int C, const float* scale_shift,
                                 const float* mean_invstd, const float* slope, int act, double* red,
                                 void* g_a_out, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 1024 && red);
  act_bwd_kernel<0><<<stream_grid((int64_t)batch * L, C, RED_CAP), 256, 0, ST>>>(
      g_h, g_h_ld > 0 ? g_h_ld : C, in_halo_pos, roll, g_add, g_add_ld > 0 ? g_add_ld : C, a, dtype, batch, L, C,
      scale_shift, mean_invstd, slope, act, red, 0, g_a_out);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_act_bwd_apply(const void* g_h, int g_h_ld, int in_halo_pos, int roll, const void* g_add,
                                int g_add_ld, const void* a,
                                int dtype, int batch, int L, int C, const float* scale_shift,
                                const float* mean_invstd, const float* slope, int act, const double* red,
                                int use_bn, void* g_a, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 1024 && red && g_a);
  act_bwd_kernel<1><<<stream_grid((int64_t)batch * L, C, 16), 256, 0, ST>>>(
      g_h, g_h_ld > 0 ? g_h_ld : C, in_halo_pos, roll, g_add, g_add_ld > 0 ? g_add_ld : C, a, dtype, batch, L, C,
      scale_shift, mean_invstd, slope, act, const_cast<double*>(red), use_bn, g_a);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_ncl_to_nlc(const float* src, int batch, int C, int L, void* dst, int dtype, void* stream) {
  dim3 grid((L + 31) / 32, (C + 31) / 32, batch), block(32, 8);
  ncl_to_nlc_kernel<<<grid, block, 0, ST>>>(src, C, L, dst, dtype);
  SG_CHECK_LAUNCH();
  return SG_OK;
}
extern "C" int sg_nlc_to_ncl(const void* src, int dtype, int batch, int C, int L, float* dst, void* stream) {
  dim3 grid((L + 31) / 32, (C + 31) / 32, batch), block(32, 8);
  nlc_to_ncl_kernel<<<grid, block, 0, ST>>>(src, dtype, C, L, dst);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_colsum(const void* a, int dtype, int64_t rows, int C, int mod, float* out, int accumulate,
                         double* tmp, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 1024 && tmp && C % mod == 0);
  SG_CHECK_CUDA(cudaMemsetAsync(tmp, 0, sizeof(double) * C * SL, ST));
  colsum_kernel<<<stream_grid(rows, C, RED_CAP), 256, 0, ST>>>(a, dtype, rows, C, tmp);
  SG_CHECK_LAUNCH();
  colsum_fold_kernel<<<(mod + 127) / 128, 128, 0, ST>>>(tmp, C, mod, out, accumulate);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_fc_tail_fwd(const float* fc0_acc, const float* b0, const float* s1, const float* w2,
                              const float* b2, const float* s3, const float* w4, const float* b4, int batch,
                              float* z1, float* z2, float* logit, void* stream) {
  fc_tail_fwd_kernel<<<batch, 256, 0, ST>>>(fc0_acc, b0, s1, w2, b2, s3, w4, b4, z1, z2, logit);
  SG_CHECK_LAUNCH();
  return SG_OK;
}

extern "C" int sg_fc_tail_bwd(const float* z1, const float* z2, const float* logit, const float* g_logit_in,
                              float target, float weight,
                              const float* s1, const float* w2, const float* s3, const float* w4, int batch,
                              float* loss_out, void* g_z1_bf16, float* ws /* [B*(1+128+256+256)] */, float* g_b0,
                              float* g_s1, float* g_w2, float* g_b2, float* g_s3, float* g_w4, float* g_b4,
                              void* stream) {
  SG_CHECK_ARG(ws && g_z1_bf16);
  float* g_logit = ws;
  float* g_z2 = ws + batch;
  float* g_z1 = g_z2 + (int64_t)batch * FC2;
  float* g_h1 = g_z1 + (int64_t)batch * FC1;
  fc_tail_bwd_rows_kernel<<<batch, 256, 0, ST>>>(z1, z2, logit, g_logit_in, target, weight, s1, w2, s3, w4, batch,
                                                 loss_out, g_logit, g_z2, g_z1, g_h1, g_z1_bf16);
  SG_CHECK_LAUNCH();
  if (g_w2) {
    dim3 grid(FC2 + 1, (batch + FC_CHUNK - 1) / FC_CHUNK);
    fc_tail_bwd_params_kernel<<<grid, 256, 0, ST>>>(z1, z2, g_logit, g_z2, g_z1, g_h1, s1, s3, w4, batch, g_b0, g_s1,
                                                    g_w2, g_b2, g_s3, g_w4, g_b4);
    SG_CHECK_LAUNCH();
  }
  return SG_OK;
}

extern "C" int sg_l1_loss_bwd(const float* y, const float* clean, int64_t n, float weight, float* loss_out,
                              float* gy, int accumulate, void* stream) {
  l1_loss_bwd_kernel<<<2 * NUM_SMS, 256, 0, ST>>>(y, clean, n, weight, loss_out, gy, accumulate);
  SG_CHECK_LAUNCH();
  return SG_OK;
}
