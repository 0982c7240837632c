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
// Streaming kernels: a warp owns one 64-channel chunk (lane = 2 adjacent channels, one 128-byte
// line per row) and strides over rows, 4 rows in flight per thread.  Per-channel constants and
// partial sums are 2 registers each, so 6+ CTAs are resident per SM and HBM latency is hidden by
// occupancy x ILP rather than by wide per-thread vectors.
// ------------------------------------------------------------------------------------------
constexpr int EW_UNROLL = 8;   // rows in flight per thread (each a 128-byte warp request per stream)

__device__ __forceinline__ float2 ld2(const void* p, int64_t elem, int dtype) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p) + elem);
  if (dtype == SG_F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void st2(void* p, int64_t elem, float x, float y, int dtype) {
  uint32_t u;
  if (dtype == SG_F16) { __half2 h = __floats2half2_rn(x, y); u = *reinterpret_cast<uint32_t*>(&h); }
  else { __nv_bfloat162 h = __floats2bfloat162_rn(x, y); u = *reinterpret_cast<uint32_t*>(&h); }
  *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p) + elem) = u;
}

struct WarpWork {
  int chunk, c, row0, row_stride, lane;
};
__device__ __forceinline__ WarpWork warp_work(int C) {
  WarpWork w;
  const int chunks = C / 64;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  w.lane = threadIdx.x & 31;
  w.chunk = gw % chunks;
  w.c = w.chunk * 64 + w.lane * 2;
  w.row0 = gw / chunks;
  w.row_stride = nw / chunks;
  return w;
}
// merges per-lane partial sums of NS statistics for channels (c, c+1): warps of a block that own the
// same chunk are combined in smem first (chunks <= 8), then one double atomic per channel
template <int NS>
__device__ __forceinline__ void warp_stats_flush(float (&part)[NS][2], const WarpWork& w, int C, double* out,
                                                 float* smem /* [8][NS][64] */) {
  const int chunks = C / 64;
  const int wib = threadIdx.x >> 5;
  if (chunks >= 8) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      atomicAdd(out + (int64_t)s * C + w.c, (double)part[s][0]);
      atomicAdd(out + (int64_t)s * C + w.c + 1, (double)part[s][1]);
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    smem[(wib * NS + s) * 64 + w.lane * 2] = part[s][0];
    smem[(wib * NS + s) * 64 + w.lane * 2 + 1] = part[s][1];
  }
  __syncthreads();
  if (wib < chunks) {       // warp `wib` owns chunk (first_chunk + wib) % chunks == w.chunk
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double a0 = 0, a1 = 0;
      for (int o = wib; o < 8; o += chunks) {
        a0 += (double)smem[(o * NS + s) * 64 + w.lane * 2];
        a1 += (double)smem[(o * NS + s) * 64 + w.lane * 2 + 1];
      }
      atomicAdd(out + (int64_t)s * C + w.c, a0);
      atomicAdd(out + (int64_t)s * C + w.c + 1, a1);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_stats_kernel(const void* __restrict__ a, int dtype, int64_t rows64, int C, double* __restrict__ stats) {
  __shared__ float red[8 * 2 * 64];
  const WarpWork w = warp_work(C);
  const int rows = (int)rows64;
  float part[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int r0 = w.row0; r0 < rows; r0 += EW_UNROLL * w.row_stride) {
    float2 v[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      v[u] = r < rows ? ld2(a, (int64_t)r * C + w.c, dtype) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      part[0][0] += v[u].x; part[0][1] += v[u].y;
      part[1][0] = fmaf(v[u].x, v[u].x, part[1][0]);
      part[1][1] = fmaf(v[u].y, v[u].y, part[1][1]);
    }
  }
  warp_stats_flush<2>(part, w, C, stats, red);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;
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
  const WarpWork w = warp_work(C);
  const int Lh = L + 2 * H;
  const int rows = batch * Lh;           // output rows (incl. halo)
  const float sc0 = scale_shift ? scale_shift[w.c] : 1.f, sc1 = scale_shift ? scale_shift[w.c + 1] : 1.f;
  const float sh0 = scale_shift ? scale_shift[C + w.c] : 0.f, sh1 = scale_shift ? scale_shift[C + w.c + 1] : 0.f;
  const float sl0 = act == SG_ACT_PRELU ? slope[w.c] : 1.f, sl1 = act == SG_ACT_PRELU ? slope[w.c + 1] : 1.f;
  for (int r0 = w.row0; r0 < rows; r0 += EW_UNROLL * w.row_stride) {
    float2 v[EW_UNROLL];
    int srcs[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      srcs[u] = 0;
      v[u] = make_float2(0.f, 0.f);
      if (r < rows) {
        const int b = r / Lh, qh = r - b * Lh;
        srcs[u] = b * L + unroll_idx(reflect_idx(qh - H, L), roll, L);
        v[u] = ld2(a, (int64_t)srcs[u] * C + w.c, dtype);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      if (r < rows) {
        const int b = r / Lh, qh = r - b * Lh;
        float y0 = fmaf(v[u].x, sc0, sh0), y1 = fmaf(v[u].y, sc1, sh1);
        if (act == SG_ACT_PRELU) { y0 = y0 > 0.f ? y0 : sl0 * y0; y1 = y1 > 0.f ? y1 : sl1 * y1; }
        st2(h, (int64_t)r * C + w.c, y0, y1, dtype);
        // bf16 twins: operands of the weight-gradient tap-GEMM (tcgen05 kind::f16 cannot mix f16 x bf16)
        if (h_bf16) st2(h_bf16, (int64_t)r * C + w.c, y0, y1, SG_BF16);
        if (a_bf16 && qh >= H && qh < H + L) st2(a_bf16, (int64_t)srcs[u] * C + w.c, v[u].x, v[u].y, SG_BF16);
      }
    }
  }
}

// gradient w.r.t. the activation output at exact position l: the consumer-view gradient at the
// rolled position plus its reflect-halo mirrors (row indices are warp-uniform)
__device__ __forceinline__ float2 gather_gy(const void* g_h, int ldh, int H, int roll, int b, int l, int L, int c) {
  float2 g = make_float2(0.f, 0.f);
  if (g_h) {
    const int Lh = L + 2 * H;
    int q0 = l + roll;
    if (q0 >= L) q0 -= L;
    if (q0 < 0) q0 += L;
    const int64_t base = (int64_t)b * Lh + H;
    g = ld2(g_h, (base + q0) * ldh + c, SG_BF16);
    if (H > 0) {
      if (q0 >= 1 && q0 <= H) {
        const float2 m = ld2(g_h, (base - q0) * ldh + c, SG_BF16);
        g.x += m.x; g.y += m.y;
      }
      if (q0 >= L - 1 - H && q0 <= L - 2) {
        const float2 m = ld2(g_h, (base + 2 * (L - 1) - q0) * ldh + c, SG_BF16);
        g.x += m.x; g.y += m.y;
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
  __shared__ float sred[8 * 3 * 64];
  const WarpWork w = warp_work(C);
  const int rows = batch * L;
  float sc[2], sh[2], mu[2], is[2], sl[2], r1[2] = {0.f, 0.f}, r2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = w.c + j;
    sc[j] = scale_shift ? scale_shift[c] : 1.f;
    sh[j] = scale_shift ? scale_shift[C + c] : 0.f;
    mu[j] = mean_invstd ? mean_invstd[c] : 0.f;
    is[j] = mean_invstd ? mean_invstd[C + c] : 1.f;
    sl[j] = (act == SG_ACT_PRELU) ? slope[c] : 1.f;
    if (MODE == 1) {
      r1[j] = (float)(red[C + c] / (double)rows);
      r2[j] = (float)(red[2 * C + c] / (double)rows);
    }
  }
  float part[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  for (int r0 = w.row0; r0 < rows; r0 += EW_UNROLL * w.row_stride) {
    float2 gy[EW_UNROLL], gs[EW_UNROLL], av[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      gy[u] = gs[u] = av[u] = make_float2(0.f, 0.f);
      if (r < rows) {
        const int b = r / L, l = r - b * L;
        gy[u] = gather_gy(g_h, ldh, H, roll, b, l, L, w.c);
        if (g_add) gs[u] = ld2(g_add, (int64_t)r * lda + w.c, SG_BF16);
        av[u] = ld2(a, (int64_t)r * C + w.c, dtype);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      if (r < rows) {
        float out[2];
        const float xs[2] = {av[u].x, av[u].y}, gys[2] = {gy[u].x, gy[u].y}, gss[2] = {gs[u].x, gs[u].y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float y = fmaf(xs[j], sc[j], sh[j]);
          const float ahat = (xs[j] - mu[j]) * is[j];
          float gpre = gys[j];
          if (act == SG_ACT_PRELU && y <= 0.f) {
            if (MODE == 0) part[0][j] = fmaf(gys[j], y, part[0][j]);
            gpre = gys[j] * sl[j];
          }
          gpre += gss[j];
          if (MODE == 0) {
            part[1][j] += gpre;
            part[2][j] = fmaf(gpre, ahat, part[2][j]);
            out[j] = gpre;
          } else {
            out[j] = use_bn ? sc[j] * (gpre - r1[j] - ahat * r2[j]) : gpre;
          }
        }
        if (g_a_out) st2(g_a_out, (int64_t)r * C + w.c, out[0], out[1], SG_BF16);
      }
    }
  }
  if (MODE == 0) warp_stats_flush<3>(part, w, C, red, sred);
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
  __shared__ float red[8 * 64];
  const WarpWork w = warp_work(C);
  const int rows = (int)rows64;
  float part[1][2] = {{0.f, 0.f}};
  for (int r0 = w.row0; r0 < rows; r0 += EW_UNROLL * w.row_stride) {
    float2 v[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int r = r0 + u * w.row_stride;
      v[u] = r < rows ? ld2(a, (int64_t)r * C + w.c, dtype) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) { part[0][0] += v[u].x; part[0][1] += v[u].y; }
  }
  warp_stats_flush<1>(part, w, C, tmp, red);
}
__global__ void colsum_fold_kernel(const double* __restrict__ tmp, int C, int mod, float* __restrict__ out,
                                   int accumulate) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= mod) return;
  double s = 0;
  for (int c = m; c < C; c += mod) s += tmp[c];
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
constexpr int RED_CAP = 4;
// grid for the warp-per-64-channel-chunk streaming kernels: multiple of 4 so that warps/chunks is integral
static inline int stream_grid(int64_t rows, int C, int cap_per_sm) {
  const int64_t units = rows * (C / 64);
  int64_t g = cdiv(units, 8 * EW_UNROLL);
  const int64_t cap = (int64_t)cap_per_sm * NUM_SMS;
  if (g > cap) g = cap;
  g = cdiv(g, 4) * 4;
  return (int)(g < 4 ? 4 : g);
}

}  // namespace sg

using namespace sg;

#define ST ((cudaStream_t)stream)

extern "C" int sg_bn_stats(const void* a, int dtype, int64_t rows_total, int C, double* stats, void* stream) {
  SG_CHECK_ARG(C % 64 == 0 && C <= 2048 && a && stats && rows_total < (1ll << 31));
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
  SG_CHECK_ARG(C % 64 == 0 && C <= 2048 && (out_halo_pos == 0 || L >= 32));
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
  SG_CHECK_ARG(C % 64 == 0 && C <= 2048 && red);
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
  SG_CHECK_ARG(C % 64 == 0 && C <= 2048 && red && g_a);
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
  SG_CHECK_ARG(C % 64 == 0 && C <= 2048 && tmp && C % mod == 0);
  SG_CHECK_CUDA(cudaMemsetAsync(tmp, 0, sizeof(double) * C, ST));
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
