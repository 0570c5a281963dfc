// Ray generation, coarse depth edges and hierarchical fine sampling (tiny HBM-bound kernels).
// Reference: utils/model_utils.py:283-375 (GenSamplePoints), :378-490 (FineSample).
#include "common.cuh"

namespace gnrf {

// d = normalize(R * (Kinv * (x,y,1))), l = -1/d_z.  One thread per ray.
// Products are accumulated k = 0,1,2 with fused multiply-adds (what a BLAS 3x3 bmm does on the host).
__global__ void ray_setup_kernel(const float* __restrict__ xy, const float* __restrict__ rmats,
                                 const float* __restrict__ kinv, int B, int N_r, float4* __restrict__ ray_dl) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N_r) return;
  int b = idx / N_r, r = idx - b * N_r;
  const float* K = kinv + b * 9;
  const float* R = rmats + b * 9;
  float x = xy[(b * 2 + 0) * N_r + r];
  float y = xy[(b * 2 + 1) * N_r + r];
  float p[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = fmaf(K[i * 3 + 2], 1.0f, fmaf(K[i * 3 + 1], y, __fmul_rn(K[i * 3 + 0], x)));
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = fmaf(R[i * 3 + 2], p[2], fmaf(R[i * 3 + 1], p[1], __fmul_rn(R[i * 3 + 0], p[0])));
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  float dx = __fdiv_rn(d[0], n), dy = __fdiv_rn(d[1], n), dz = __fdiv_rn(d[2], n);
  float l = __fdiv_rn(-1.0f, dz);
  ray_dl[idx] = make_float4(dx, dy, dz, l);
}

__device__ __forceinline__ float coarse_z(float rel1, float rel2, float t) {
  // rela_z1 * (1 - t) + rela_z2 * t, each op individually rounded (utils/model_utils.py:354-356)
  return __fadd_rn(__fmul_rn(rel1, __fsub_rn(1.0f, t)), __fmul_rn(rel2, t));
}

__global__ void coarse_depths_kernel(const float* __restrict__ tvecs, const float* __restrict__ t_vals,
                                     const float* __restrict__ jitter_u, int B, int N_r, int N_s, float z1, float z2,
                                     float* __restrict__ z_edges) {
  int n_e = N_s + 1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * N_r * n_e;
  if (idx >= total) return;
  int k = (int)(idx % n_e);
  int b = (int)(idx / ((long long)N_r * n_e));
  float oz = tvecs[b * 3 + 2];
  float rel1 = __fsub_rn(oz, z1), rel2 = __fsub_rn(oz, z2);
  float z = coarse_z(rel1, rel2, t_vals[k]);
  if (jitter_u != nullptr) {
    // stratified jitter (utils/model_utils.py:302-307)
    float zlo = z, zhi = z;
    if (k > 0) zlo = __fmul_rn(0.5f, __fadd_rn(z, coarse_z(rel1, rel2, t_vals[k - 1])));
    if (k < N_s) zhi = __fmul_rn(0.5f, __fadd_rn(coarse_z(rel1, rel2, t_vals[k + 1]), z));
    z = __fadd_rn(zlo, __fmul_rn(__fsub_rn(zhi, zlo), jitter_u[idx]));
  }
  z_edges[idx] = z;
}

// One CTA per ray. n_c coarse samples -> m = n_c - 2 interior weights -> cdf[m+1]; n_f1 fine depths; bitonic sort.
constexpr int kFineThreads = 128;
constexpr int kFineMax = 512;  // n_c + n_f1 <= 512

__global__ void __launch_bounds__(kFineThreads) fine_depths_kernel(const float* __restrict__ weights,
                                                                    const float* __restrict__ z_edges_c,
                                                                    const float* __restrict__ u_in, int u_per_ray, int n_c,
                                                                    int n_f1, long long* __restrict__ inds_out,
                                                                    float* __restrict__ z_out) {
  __shared__ float s_cdf[kFineMax];
  __shared__ float s_zc[kFineMax];
  __shared__ float s_sort[kFineMax];
  const int ray = blockIdx.x;  // flattened (b, r)
  const int tid = threadIdx.x;
  const int m = n_c - 2;
  const float* w = weights + (size_t)ray * n_c;
  const float* zc = z_edges_c + (size_t)ray * (n_c + 1);
  for (int i = tid; i < n_c; i += kFineThreads) s_zc[i] = zc[i];
  if (tid == 0) {
    // pdf = w / sum(w + 1e-5); cdf = [0, cumsum(pdf)]  (utils/model_utils.py:417-421).
    // torch's CPU cumsum accumulates float inputs in double and rounds every prefix to float; the total is the
    // correctly rounded float sum.
    double tot = 0.0;
    for (int j = 0; j < m; ++j) tot += (double)__fadd_rn(w[j + 1], 1e-5f);
    float s = (float)tot;
    double acc = 0.0;
    s_cdf[0] = 0.0f;
    for (int j = 0; j < m; ++j) {
      acc += (double)__fdiv_rn(w[j + 1], s);
      s_cdf[j + 1] = (float)acc;
    }
  }
  __syncthreads();
  int n_tot = n_c + n_f1;
  int n_pad = 1;
  while (n_pad < n_tot) n_pad <<= 1;
  for (int i = tid; i < n_pad; i += kFineThreads) {
    float v = INFINITY;
    if (i < n_c) {
      v = s_zc[i];
    } else if (i < n_tot) {
      int j = i - n_c;
      float u = u_per_ray ? u_in[(size_t)ray * n_f1 + j] : u_in[j];
      // searchsorted(cdf, u, right=True): number of cdf entries <= u  (cdf is non-decreasing, m+1 entries)
      int lo = 0, hi = m + 1;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      int ind = lo;
      if (inds_out != nullptr) inds_out[(size_t)ray * n_f1 + j] = ind;
      int below = max(ind - 1, 0), above = min(ind, m);
      float cb = s_cdf[below], ca = s_cdf[above];
      float bb = __fmul_rn(0.5f, __fadd_rn(s_zc[below + 1], s_zc[below]));
      float ba = __fmul_rn(0.5f, __fadd_rn(s_zc[above + 1], s_zc[above]));
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      float t = __fdiv_rn(__fsub_rn(u, cb), denom);
      v = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
    }
    s_sort[i] = v;
  }
  __syncthreads();
  // bitonic sort ascending
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_pad; i += kFineThreads) {
        int ixj = i ^ j;
        if (ixj > i) {
          float a = s_sort[i], c = s_sort[ixj];
          bool up = ((i & k) == 0);
          if ((a > c) == up) { s_sort[i] = c; s_sort[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n_tot; i += kFineThreads) z_out[(size_t)ray * n_tot + i] = s_sort[i];
}

}  // namespace gnrf

using namespace gnrf;

extern "C" int gnrf_ray_setup(const float* xy, const float* rmats, const float* inv_inmats, int B, int N_r, float* ray_dl,
                              gnrf_stream_t stream) {
  GNRF_CHECK_ARG(xy && rmats && inv_inmats && ray_dl);
  GNRF_CHECK_ARG(B > 0 && N_r > 0);
  int n = B * N_r;
  ray_setup_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(xy, rmats, inv_inmats, B, N_r,
                                                                    reinterpret_cast<float4*>(ray_dl));
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_coarse_depths(const float* tvecs, const float* t_vals, const float* jitter_u, int B, int N_r, int N_s,
                                  float world_z1, float world_z2, float* z_edges, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(tvecs && t_vals && z_edges);
  GNRF_CHECK_ARG(B > 0 && N_r > 0 && N_s > 0);
  long long total = (long long)B * N_r * (N_s + 1);
  coarse_depths_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(tvecs, t_vals, jitter_u, B, N_r, N_s,
                                                                                     world_z1, world_z2, z_edges);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}

extern "C" int gnrf_fine_depths(const float* weights, const float* z_edges_coarse, const float* u, int u_per_ray, int B,
                                int N_r, int N_c, int N_f1, int64_t* inds, float* z_edges_fine, gnrf_stream_t stream) {
  GNRF_CHECK_ARG(weights && z_edges_coarse && u && z_edges_fine);
  GNRF_CHECK_ARG(B > 0 && N_r > 0 && N_c >= 3 && N_f1 >= 1);
  if (N_c + N_f1 > kFineMax)
    return fail(GNRF_ERR_UNSUPPORTED, "gnrf_fine_depths: N_c + N_f1 = %d exceeds %d", N_c + N_f1, kFineMax);
  fine_depths_kernel<<<B * N_r, kFineThreads, 0, as_stream(stream)>>>(weights, z_edges_coarse, u, u_per_ray, N_c, N_f1,
                                                                      reinterpret_cast<long long*>(inds), z_edges_fine);
  GNRF_LAUNCH_CHECK();
  count_launches(1);
  return GNRF_OK;
}
