// Codebook path: pre_vq projection (+l2 normalise), exact-fp32 nearest-neighbour search with
// the reference's  (sum z^2 - 2 z.E) + sum E^2  association and first-min tie rule, and the
// decode-side gather + post_vq projection.
#include "omt_common.cuh"

namespace omt {

// ---------------------------------------------------------------------------------------
// pre_vq: z[M, CD] = x[M, C] . W[CD, C]^T + b   (C = 512 -> 4 float4 per lane), optional l2 norm.
// One warp per row; W staged in shared memory.
// ---------------------------------------------------------------------------------------
template <int CD>
__global__ void __launch_bounds__(256) pre_vq_kernel(const float* __restrict__ x, int ldx,
                                                     const float* __restrict__ Wt,
                                                     const float* __restrict__ b,
                                                     float* __restrict__ z, int M, int C, int l2) {
  pdl_sync();
  extern __shared__ float4 wsm4[];   // [CD][C/4]
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < CD * C4; i += blockDim.x) wsm4[i] = reinterpret_cast<const float4*>(Wt)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int row = blockIdx.x * warps + (threadIdx.x >> 5); row < M; row += gridDim.x * warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float acc[CD];
#pragma unroll
    for (int j = 0; j < CD; ++j) acc[j] = 0.f;
    for (int c = lane; c < C4; c += 32) {
      const float4 xv = xr[c];
#pragma unroll
      for (int j = 0; j < CD; ++j) {
        const float4 wv = wsm4[j * C4 + c];
        acc[j] = fmaf(xv.x, wv.x, acc[j]);
        acc[j] = fmaf(xv.y, wv.y, acc[j]);
        acc[j] = fmaf(xv.z, wv.z, acc[j]);
        acc[j] = fmaf(xv.w, wv.w, acc[j]);
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CD; ++j) {
      acc[j] = warp_sum(acc[j]) + b[j];
      ss = fmaf(acc[j], acc[j], ss);
    }
    float den = 1.f;
    if (l2) den = fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < CD; ++j) z[(size_t)row * CD + j] = l2 ? acc[j] / den : acc[j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// VQ search.  grid = (row blocks, 4 code quarters); each CTA stages its quarter of the table
// (n_codes/4 x 8 fp32, 64 KiB for 8192 codes) plus sum E^2 in shared memory, one thread per row.
// Partial (best_d, best_idx) per quarter go to the workspace; the combine kernel takes the
// first minimum across quarters (ascending index order == torch.argmin tie rule).
// ---------------------------------------------------------------------------------------
constexpr int VQ_SPLIT = 4;

__global__ void __launch_bounds__(256) vq_search_kernel(const float* __restrict__ z,
                                                        const float* __restrict__ E,
                                                        const float* __restrict__ e2, int M,
                                                        int n_codes, float* __restrict__ pd,
                                                        int* __restrict__ pi) {
  pdl_sync();
  extern __shared__ float4 esm[];            // [per][2] float4 + e2[per]
  const int per = n_codes / VQ_SPLIT;
  const int k0 = blockIdx.y * per;
  float* e2s = reinterpret_cast<float*>(esm + 2 * per);
  for (int i = threadIdx.x; i < 2 * per; i += blockDim.x)
    esm[i] = reinterpret_cast<const float4*>(E + (size_t)k0 * 8)[i];
  for (int i = threadIdx.x; i < per; i += blockDim.x) e2s[i] = e2[k0 + i];
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const float4 za = reinterpret_cast<const float4*>(z + (size_t)row * 8)[0];
  const float4 zb = reinterpret_cast<const float4*>(z + (size_t)row * 8)[1];
  // sum z^2 exactly as torch's (z**2).sum(dim=1): sequential over the 8 channels
  float zz = za.x * za.x;
  zz += za.y * za.y; zz += za.z * za.z; zz += za.w * za.w;
  zz += zb.x * zb.x; zz += zb.y * zb.y; zz += zb.z * zb.z; zz += zb.w * zb.w;
  // (2*z) @ E^T : the factor 2 is exact, fold it into z
  const float z0 = 2.f * za.x, z1 = 2.f * za.y, z2 = 2.f * za.z, z3 = 2.f * za.w;
  const float z4 = 2.f * zb.x, z5 = 2.f * zb.y, z6 = 2.f * zb.z, z7 = 2.f * zb.w;
  float best = INFINITY;
  int bi = 0;
#pragma unroll 4
  for (int k = 0; k < per; ++k) {
    const float4 ea = esm[2 * k], eb = esm[2 * k + 1];
    float dot = z0 * ea.x;
    dot = fmaf(z1, ea.y, dot); dot = fmaf(z2, ea.z, dot); dot = fmaf(z3, ea.w, dot);
    dot = fmaf(z4, eb.x, dot); dot = fmaf(z5, eb.y, dot); dot = fmaf(z6, eb.z, dot);
    dot = fmaf(z7, eb.w, dot);
    const float d = (zz - dot) + e2s[k];
    if (d < best) { best = d; bi = k; }      // strict <  => first minimum wins
  }
  pd[(size_t)blockIdx.y * M + row] = best;
  pi[(size_t)blockIdx.y * M + row] = k0 + bi;
}

__global__ void __launch_bounds__(256) vq_combine_kernel(const float* __restrict__ pd,
                                                         const int* __restrict__ pi, int M,
                                                         int64_t* __restrict__ idx,
                                                         int32_t* __restrict__ counts) {
  pdl_sync();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float best = pd[row];
  int bi = pi[row];
#pragma unroll
  for (int s = 1; s < VQ_SPLIT; ++s) {
    const float d = pd[(size_t)s * M + row];
    if (d < best) { best = d; bi = pi[(size_t)s * M + row]; }
  }
  idx[row] = bi;
  if (counts != nullptr) atomicAdd(counts + bi, 1);
}

// ---------------------------------------------------------------------------------------
// post_vq: X[M, C] = zrow[M, CD] . W[C, CD]^T + b, zrow = E[idx] | zc | (E[idx]-z)+z.
// Block = C/4 threads, thread owns 4 output channels (its 4 x CD weights live in registers).
// ---------------------------------------------------------------------------------------
constexpr int POSTVQ_ROWS = 32;

template <int CD>
__global__ void __launch_bounds__(128) post_vq_kernel(const int64_t* __restrict__ idx,
                                                      const float* __restrict__ E,
                                                      const float* __restrict__ zc,
                                                      const float* __restrict__ zst,
                                                      float* __restrict__ zq_out,
                                                      const float* __restrict__ Wt,
                                                      const float* __restrict__ b,
                                                      float* __restrict__ X, int M, int C) {
  pdl_sync();
  __shared__ float rows[POSTVQ_ROWS][CD];
  const int r0 = blockIdx.x * POSTVQ_ROWS;
  for (int i = threadIdx.x; i < POSTVQ_ROWS * CD; i += blockDim.x) {
    const int r = r0 + i / CD, j = i % CD;
    float v = 0.f;
    if (r < M) {
      if (idx != nullptr) {
        v = E[(size_t)idx[r] * CD + j];
        if (zst != nullptr) {                       // straight-through rounding (codebook.py:120)
          const float zz = zst[(size_t)r * CD + j];
          v = (v - zz) + zz;
        }
      } else {
        v = zc[(size_t)r * CD + j];
      }
      if (zq_out != nullptr) zq_out[(size_t)r * CD + j] = v;
    }
    rows[i / CD][j] = v;
  }
  __syncthreads();
  const int c = threadIdx.x * 4;
  if (c >= C) return;
  float w[4][CD];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CD; ++j) w[i][j] = Wt[(size_t)(c + i) * CD + j];
  const float4 bb = *reinterpret_cast<const float4*>(b + c);
  for (int rr = 0; rr < POSTVQ_ROWS; ++rr) {
    const int r = r0 + rr;
    if (r >= M) break;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < CD; ++j) {
      const float zv = rows[rr][j];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(zv, w[i][j], o[i]);
    }
    *reinterpret_cast<float4*>(X + (size_t)r * C + c) =
        make_float4(o[0] + bb.x, o[1] + bb.y, o[2] + bb.z, o[3] + bb.w);
  }
}

}  // namespace omt

using namespace omt;

extern "C" int omt_pre_vq(const float* x, int ldx, const float* Wt, const float* b, float* z, int M, int C,
                          int cd, int l2, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(x && Wt && b && z, "omt_pre_vq: null pointer");
  OMT_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && C <= 1024, "omt_pre_vq: bad C/ldx");
  OMT_REQUIRE(cd == 8 || cd == 16, "omt_pre_vq: codebook_dim %d unsupported (8 or 16)", cd);
  if (M == 0) return OMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)cd * C * sizeof(float);
  int blocks = (M + 7) / 8;
  const int cap = omt::sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (cd == 8) {
    OMT_CUDA(launch_k(pre_vq_kernel<8>, dim3(blocks), dim3(256), smem, st, x, ldx, Wt, b, z, M, C, l2));
  } else {
    static bool set16 = false;
    if (!set16) {
      OMT_CUDA(cudaFuncSetAttribute(pre_vq_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
      set16 = true;
    }
    OMT_CUDA(launch_k(pre_vq_kernel<16>, dim3(blocks), dim3(256), smem, st, x, ldx, Wt, b, z, M, C, l2));
  }
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_vq_search(const float* z, const float* E, const float* e2, int M, int n_codes,
                             int64_t* idx, int32_t* counts, void* workspace, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(z && E && e2 && idx && workspace, "omt_vq_search: null pointer");
  OMT_REQUIRE(n_codes % VQ_SPLIT == 0 && n_codes >= VQ_SPLIT, "omt_vq_search: n_codes %% 4 != 0");
  const int per = n_codes / VQ_SPLIT;
  const size_t smem = (size_t)per * 36;
  OMT_REQUIRE(smem <= 200 * 1024, "omt_vq_search: n_codes=%d too large for the shared-memory table", n_codes);
  if (M == 0) return OMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    OMT_CUDA(cudaFuncSetAttribute(vq_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  float* pd = reinterpret_cast<float*>(workspace);
  int* pi = reinterpret_cast<int*>(pd + (size_t)VQ_SPLIT * M);
  dim3 grid((M + 255) / 256, VQ_SPLIT);
  OMT_CUDA(launch_k(vq_search_kernel, grid, dim3(256), smem, st, z, E, e2, M, n_codes, pd, pi));
  OMT_LAUNCH_CHECK();
  OMT_CUDA(launch_k(vq_combine_kernel, dim3((M + 255) / 256), dim3(256), 0, st, (const float*)pd, (const int*)pi, M, idx, counts));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_post_vq(const int64_t* idx, const float* E, const float* zc, const float* z_st_from,
                           float* zq_out, const float* Wt, const float* b, float* X, int M, int C, int cd,
                           omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(Wt && b && X, "omt_post_vq: null pointer");
  OMT_REQUIRE((idx != nullptr && E != nullptr) || zc != nullptr, "omt_post_vq: need idx+E or zc");
  OMT_REQUIRE(C % 4 == 0 && C <= 512, "omt_post_vq: C=%d unsupported", C);
  OMT_REQUIRE(cd == 8, "omt_post_vq: codebook_dim %d unsupported (8)", cd);
  if (M == 0) return OMT_OK;
  OMT_CUDA(launch_k(post_vq_kernel<8>, dim3((M + POSTVQ_ROWS - 1) / POSTVQ_ROWS), dim3(128), 0, (cudaStream_t)stream,
                    idx, E, zc, z_st_from, zq_out, Wt, b, X, M, C));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}
