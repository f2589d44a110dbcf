// GroupNorm (channels-last, two launches) and LayerNorm.  HBM-bound: 16-byte vector loads, fp32 math,
// fp64 only for the cross-block accumulation of the GroupNorm statistics.
#include "common.cuh"
#include "kernels.h"

namespace ca {

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[n][g] = {sum x, sum x^2} over rows x (C/groups) channels.
// grid (slabs, n); each thread owns one 8-channel vector and strides over rows of the slab.
// ---------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, int c0, const __nv_bfloat16* __restrict__ x1,
                                int c1, long long rows, int groups, int rows_per_block, double* __restrict__ sums) {
  __shared__ float s_acc[64][2];
  const int C = c0 + c1;
  const int nvec = C >> 3;
  const int cpg = C / groups;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  const int rows_per_iter = blockDim.x / nvec;
  const int vec = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  if (rsub < rows_per_iter) {
    const long long r_begin = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r_end = min(rows, r_begin + rows_per_block);
    const int ch = vec * 8;
    const bool second = ch >= c0;
    const __nv_bfloat16* base = second ? x1 + (ch - c0) : x0 + ch;
    const int cs = second ? c1 : c0;
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
    for (long long r = r_begin + rsub; r < r_end; r += rows_per_iter) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(n) * rows + r) * cs));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        s[2 * e] += f.x; ss[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; ss[2 * e + 1] += f.y * f.y;
      }
    }
    // fold the 8 channels into (at most a few) groups
    int g_cur = ch / cpg;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (ch + e) / cpg;
      if (g != g_cur) {
        atomicAdd(&s_acc[g_cur][0], a);
        atomicAdd(&s_acc[g_cur][1], b);
        a = 0.f; b = 0.f; g_cur = g;
      }
      a += s[e]; b += ss[e];
    }
    atomicAdd(&s_acc[g_cur][0], a);
    atomicAdd(&s_acc[g_cur][1], b);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&sums[static_cast<long long>(n) * groups * 2 + i], static_cast<double>((&s_acc[0][0])[i]));
}

cudaError_t launch_gn_stats(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1, int n, long long rows,
                            int groups, double* sums, cudaStream_t stream) {
  const int C = c0 + c1;
  if (groups > 64 || C % groups != 0 || (C & 7) != 0 || (c0 & 7) != 0 || (C >> 3) > 512) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * n * groups * 2, stream);
  if (e != cudaSuccess) return e;
  const int threads = 512;
  // aim for >= ~4 waves of blocks while keeping >= 64 rows per block
  long long want_blocks = (148LL * 8 + n - 1) / n;
  long long rpb = (rows + want_blocks - 1) / want_blocks;
  if (rpb < 64) rpb = 64;
  const int slabs = static_cast<int>((rows + rpb - 1) / rpb);
  gn_stats_kernel<<<dim3(slabs, n), threads, 0, stream>>>(x0, c0, x1, c1, rows, groups, static_cast<int>(rpb), sums);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU, +2x nearest upsample, +concat of two sources), bf16 out.
// grid (pixel slabs, images); image i uses the statistics of sample i / imgs_per_sample.
// ---------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int c0, const __nv_bfloat16* __restrict__ x1,
                                int c1, int h, int w, int imgs_per_sample, int groups, float eps,
                                const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int silu, int up2x, int pix_per_block,
                                __nv_bfloat16* __restrict__ y) {
  __shared__ float s_mean[64], s_rstd[64];
  const int C = c0 + c1;
  const int nvec = C >> 3;
  const int cpg = C / groups;
  const int img = blockIdx.y;
  const int sample = img / imgs_per_sample;
  const long long hw = static_cast<long long>(h) * w;
  if (threadIdx.x < groups) {
    const double cnt = static_cast<double>(hw) * imgs_per_sample * cpg;
    const double s = sums[(static_cast<long long>(sample) * groups + threadIdx.x) * 2];
    const double ss = sums[(static_cast<long long>(sample) * groups + threadIdx.x) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const long long p_begin = static_cast<long long>(blockIdx.x) * pix_per_block;
  const long long p_end = min(hw, p_begin + pix_per_block);
  const long long total = (p_end - p_begin) * nvec;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const long long pix = p_begin + idx / nvec;
    const int vec = static_cast<int>(idx % nvec);
    const int ch = vec * 8;
    const bool second = ch >= c0;
    const __nv_bfloat16* src = second ? x1 + (static_cast<long long>(img) * hw + pix) * c1 + (ch - c0)
                                      : x0 + (static_cast<long long>(img) * hw + pix) * c0 + ch;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&u);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ch + e;
      const int g = c / cpg;
      float v = (__bfloat162float(hv[e]) - s_mean[g]) * s_rstd[g];
      v = v * __ldg(gamma + c) + __ldg(beta + c);
      if (silu) v = v / (1.0f + __expf(-v));
      f[e] = v;
    }
    const uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                               pack_bf16x2(f[6], f[7]));
    if (!up2x) {
      *reinterpret_cast<uint4*>(y + (static_cast<long long>(img) * hw + pix) * C + ch) = o;
    } else {
      const int py = static_cast<int>(pix / w), px = static_cast<int>(pix % w);
      const long long ow = 2LL * w;
      __nv_bfloat16* dst = y + ((static_cast<long long>(img) * 2 * h + 2 * py) * ow + 2 * px) * C + ch;
      *reinterpret_cast<uint4*>(dst) = o;
      *reinterpret_cast<uint4*>(dst + C) = o;
      *reinterpret_cast<uint4*>(dst + ow * C) = o;
      *reinterpret_cast<uint4*>(dst + ow * C + C) = o;
    }
  }
}

cudaError_t launch_gn_apply(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1, int n, int h, int w,
                            int imgs_per_sample, int groups, float eps, const double* sums, const float* gamma,
                            const float* beta, int silu, int up2x, __nv_bfloat16* y, cudaStream_t stream) {
  const int C = c0 + c1;
  if (groups > 64 || C % groups != 0 || (C & 7) != 0 || (c0 & 7) != 0) return cudaErrorInvalidValue;
  const long long hw = static_cast<long long>(h) * w;
  long long want_blocks = (148LL * 8 + n - 1) / n;
  long long ppb = (hw + want_blocks - 1) / want_blocks;
  if (ppb < 16) ppb = 16;
  const int slabs = static_cast<int>((hw + ppb - 1) / ppb);
  gn_apply_kernel<<<dim3(slabs, n), 256, 0, stream>>>(x0, c0, x1, c1, h, w, imgs_per_sample, groups, eps, sums, gamma,
                                                      beta, silu, up2x, static_cast<int>(ppb), y);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two passes over the row (second pass hits L1).
// Optional fused pre-add of a broadcast row vector (frame position embedding).
// ---------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int c, float eps,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const __nv_bfloat16* __restrict__ add_rowvec, long long rows_per_vec,
                                 __nv_bfloat16* __restrict__ y_sum, __nv_bfloat16* __restrict__ y) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = c >> 3;
  const __nv_bfloat16* xr = x + row * c;
  const __nv_bfloat16* av = add_rowvec ? add_rowvec + (row / rows_per_vec) * c : nullptr;
  float s = 0.f, ss = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + v * 8);
    const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&u);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(hv[e]);
    if (av) {
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(av + v * 8));
      const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&ua);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = round_bf16(f[e] + __bfloat162float(ha[e]));
      if (y_sum) {
        *reinterpret_cast<uint4*>(y_sum + row * c + v * 8) = make_uint4(
            pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += f[e]; ss += f[e] * f[e]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const float mean = s / c;
  const float var = fmaxf(ss / c - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int v = lane; v < nvec; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + v * 8);
    const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&u);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(hv[e]);
    if (av) {
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(av + v * 8));
      const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&ua);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = round_bf16(f[e] + __bfloat162float(ha[e]));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
      f[e] = (f[e] - mean) * rstd * __ldg(gamma + v * 8 + e) + __ldg(beta + v * 8 + e);
    *reinterpret_cast<uint4*>(y + row * c + v * 8) = make_uint4(
        pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

cudaError_t launch_layernorm(const __nv_bfloat16* x, long long rows, int c, float eps, const float* gamma,
                             const float* beta, const __nv_bfloat16* add_rowvec, long long rows_per_vec,
                             __nv_bfloat16* y_sum, __nv_bfloat16* y, cudaStream_t stream) {
  if ((c & 7) != 0) return cudaErrorInvalidValue;
  const int warps = 8;
  const long long blocks = (rows + warps - 1) / warps;
  layernorm_kernel<<<static_cast<unsigned>(blocks), warps * 32, 0, stream>>>(x, rows, c, eps, gamma, beta, add_rowvec,
                                                                           rows_per_vec > 0 ? rows_per_vec : 1,
                                                                           y_sum, y);
  return cudaGetLastError();
}

}  // namespace ca
