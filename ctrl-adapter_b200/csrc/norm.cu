// GroupNorm (channels-last, two launches) and LayerNorm.  HBM-bound: 16-byte vector loads with several loads in
// flight per thread, fp32 math, fp64 only for the cross-block accumulation of the GroupNorm statistics.
#include "common.cuh"
#include "kernels.h"

namespace ca {

__device__ __forceinline__ uint4 ld_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[n][g] = {sum x, sum x^2} over rows x (C/groups) channels.
// grid (slabs, n); each thread owns one 8-channel vector column and strides over the rows of the slab,
// 4 rows per iteration so that 4 independent 16-byte loads are in flight.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, int c0, const __nv_bfloat16* __restrict__ x1, int c1,
                long long rows, int groups, int rows_per_block, double* __restrict__ sums) {
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
    const int cs = second ? c1 : c0;
    const __nv_bfloat16* base = (second ? x1 + (ch - c0) : x0 + ch) + static_cast<long long>(n) * rows * cs;
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
    long long r = r_begin + rsub;
    for (; r + 3LL * rows_per_iter < r_end; r += 4LL * rows_per_iter) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_nc(base + (r + static_cast<long long>(k) * rows_per_iter) * cs);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          s[2 * e] += f.x; ss[2 * e] = fmaf(f.x, f.x, ss[2 * e]);
          s[2 * e + 1] += f.y; ss[2 * e + 1] = fmaf(f.y, f.y, ss[2 * e + 1]);
        }
      }
    }
    for (; r < r_end; r += rows_per_iter) {
      const uint4 u = ld_nc(base + r * cs);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        s[2 * e] += f.x; ss[2 * e] = fmaf(f.x, f.x, ss[2 * e]);
        s[2 * e + 1] += f.y; ss[2 * e + 1] = fmaf(f.y, f.y, ss[2 * e + 1]);
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
  // >= ~4 resident blocks per SM overall, while keeping the per-thread fp32 partial sums short (<= ~64 rows)
  const int rows_per_iter = threads / (C >> 3);
  long long want_blocks = (148LL * 4 + n - 1) / n;
  long long rpb = (rows + want_blocks - 1) / want_blocks;
  const long long min_rpb = 16LL * rows_per_iter, max_rpb = 64LL * rows_per_iter;
  if (rpb < min_rpb) rpb = min_rpb;
  if (rpb > max_rpb) rpb = max_rpb;
  const int slabs = static_cast<int>((rows + rpb - 1) / rpb);
  gn_stats_kernel<<<dim3(slabs, n), threads, 0, stream>>>(x0, c0, x1, c1, rows, groups, static_cast<int>(rpb), sums);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU, +2x nearest upsample, +concat of two sources), bf16 out.
// grid (pixel slabs, images); image i uses the statistics of sample i / imgs_per_sample.
// Per-channel scale/shift (gamma*rstd, beta - mean*gamma*rstd) are staged in shared memory once per block.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int c0, const __nv_bfloat16* __restrict__ x1, int c1, int h,
                int w, int imgs_per_sample, int groups, float eps, const double* __restrict__ sums,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu, int up2x, int pix_per_block,
                __nv_bfloat16* __restrict__ y) {
  extern __shared__ float s_ab[];  // [C] scale, [C] shift
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
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = s_rstd[g] * __ldg(gamma + c);
    s_ab[c] = a;
    s_ab[C + c] = __ldg(beta + c) - s_mean[g] * a;
  }
  __syncthreads();
  const long long p_begin = static_cast<long long>(blockIdx.x) * pix_per_block;
  const long long p_end = min(hw, p_begin + pix_per_block);
  const long long total = (p_end - p_begin) * nvec;
  const __nv_bfloat16* img0 = x0 + static_cast<long long>(img) * hw * c0;
  const __nv_bfloat16* img1 = x1 ? x1 + static_cast<long long>(img) * hw * c1 : nullptr;
  for (long long idx0 = threadIdx.x; idx0 < total; idx0 += 2LL * blockDim.x) {
    uint4 u[2];
    long long pix[2];
    int ch[2];
    bool ok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long idx = idx0 + static_cast<long long>(k) * blockDim.x;
      ok[k] = idx < total;
      pix[k] = p_begin + (ok[k] ? idx / nvec : 0);
      ch[k] = static_cast<int>((ok[k] ? idx % nvec : 0)) * 8;
      const __nv_bfloat16* src = (ch[k] >= c0) ? img1 + pix[k] * c1 + (ch[k] - c0) : img0 + pix[k] * c0 + ch[k];
      if (ok[k]) u[k] = ld_nc(src);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!ok[k]) continue;
      const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&u[k]);
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = fmaf(__bfloat162float(hv[e]), s_ab[ch[k] + e], s_ab[C + ch[k] + e]);
        if (silu) v = v / (1.0f + __expf(-v));
        f[e] = v;
      }
      const uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                 pack_bf16x2(f[6], f[7]));
      if (!up2x) {
        *reinterpret_cast<uint4*>(y + (static_cast<long long>(img) * hw + pix[k]) * C + ch[k]) = o;
      } else {
        const int py = static_cast<int>(pix[k] / w), px = static_cast<int>(pix[k] % w);
        const long long ow = 2LL * w;
        __nv_bfloat16* dst = y + ((static_cast<long long>(img) * 2 * h + 2 * py) * ow + 2 * px) * C + ch[k];
        *reinterpret_cast<uint4*>(dst) = o;
        *reinterpret_cast<uint4*>(dst + C) = o;
        *reinterpret_cast<uint4*>(dst + ow * C) = o;
        *reinterpret_cast<uint4*>(dst + ow * C + C) = o;
      }
    }
  }
}

cudaError_t launch_gn_apply(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1, int n, int h, int w,
                            int imgs_per_sample, int groups, float eps, const double* sums, const float* gamma,
                            const float* beta, int silu, int up2x, __nv_bfloat16* y, cudaStream_t stream) {
  const int C = c0 + c1;
  if (groups > 64 || C % groups != 0 || (C & 7) != 0 || (c0 & 7) != 0 || C > 4096) return cudaErrorInvalidValue;
  const long long hw = static_cast<long long>(h) * w;
  long long want_blocks = (148LL * 8 + n - 1) / n;
  long long ppb = (hw + want_blocks - 1) / want_blocks;
  if (ppb < 16) ppb = 16;
  const int slabs = static_cast<int>((hw + ppb - 1) / ppb);
  gn_apply_kernel<<<dim3(slabs, n), 256, 2 * C * sizeof(float), stream>>>(x0, c0, x1, c1, h, w, imgs_per_sample, groups,
                                                                          eps, sums, gamma, beta, silu, up2x,
                                                                          static_cast<int>(ppb), y);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row is held in registers (NV 16-byte vectors per lane): one global read,
// two-pass mean / variance from registers, one write.  Optional fused pre-add of a broadcast row vector
// (frame position embedding / single-token cross-attention output).
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int c, float eps,
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
  uint4 u[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) u[k] = ld_nc(xr + v * 8);
  }
  float f[NV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&u[k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[k][e] = __bfloat162float(hv[e]);
      if (av) {
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(av + v * 8));
        const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&ua);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[k][e] = round_bf16(f[k][e] + __bfloat162float(ha[e]));
        if (y_sum) {
          *reinterpret_cast<uint4*>(y_sum + row * c + v * 8) =
              make_uint4(pack_bf16x2(f[k][0], f[k][1]), pack_bf16x2(f[k][2], f[k][3]), pack_bf16x2(f[k][4], f[k][5]),
                         pack_bf16x2(f[k][6], f[k][7]));
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += f[k][e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / c;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (lane + 32 * k < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = f[k][e] - mean; ss = fmaf(d, d, ss); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / c + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (f[k][e] - mean) * rstd * gg[e] + bb[e];
      *reinterpret_cast<uint4*>(y + row * c + v * 8) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
}

cudaError_t launch_layernorm(const __nv_bfloat16* x, long long rows, int c, float eps, const float* gamma,
                             const float* beta, const __nv_bfloat16* add_rowvec, long long rows_per_vec,
                             __nv_bfloat16* y_sum, __nv_bfloat16* y, cudaStream_t stream) {
  if ((c & 7) != 0 || c > 8 * 32 * 8) return cudaErrorInvalidValue;
  const int warps = 8;
  const unsigned blocks = static_cast<unsigned>((rows + warps - 1) / warps);
  const long long rpv = rows_per_vec > 0 ? rows_per_vec : 1;
  const int nv = ((c >> 3) + 31) / 32;
#define CA_LN(NV) layernorm_kernel<NV><<<blocks, warps * 32, 0, stream>>>(x, rows, c, eps, gamma, beta, add_rowvec, rpv, y_sum, y)
  switch (nv) {
    case 1: CA_LN(1); break;
    case 2: CA_LN(2); break;
    case 3: CA_LN(3); break;
    case 4: CA_LN(4); break;
    case 5: CA_LN(5); break;
    case 6: CA_LN(6); break;
    case 7: CA_LN(7); break;
    default: CA_LN(8); break;
  }
#undef CA_LN
  return cudaGetLastError();
}

}  // namespace ca
