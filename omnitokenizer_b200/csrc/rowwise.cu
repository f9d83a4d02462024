// Row-wise / HBM-bound kernels of the OmniTokenizer encode/decode path:
// LayerNorm, patch gather+LN, un-patchify, PEG gather-stencil, rope+l2norm+scale.
// All of them stream the canonical X[B][T'][N][C] buffer once with 16-byte accesses.
#include "omt_common.cuh"
#include <string.h>

namespace omt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int g_peg_kernel = 4;   // omt_set_option("peg_kernel", 3|4): 4 = peg_tile4_kernel (cp.async gather + FFMA2, default), 3 = peg_tile_kernel
int g_pdl = 0;   // measured on B200: PDL made the step 2-4 % slower (dependent CTAs hold SM resources during the tail), so it is opt-in
static int g_dev_ok[64];   // 0 unknown, 1 ok, -1 bad
static int g_sms[64];

int check_device() {
  int dev = 0;
  OMT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return OMT_E_ARG; }
  if (g_dev_ok[dev] == 0) {
    cudaDeviceProp p;
    OMT_CUDA(cudaGetDeviceProperties(&p, dev));
    g_sms[dev] = p.multiProcessorCount;
    g_dev_ok[dev] = (p.major == 10) ? 1 : -1;
  }
  if (g_dev_ok[dev] < 0) {
    set_error("omnitok_b200 kernels are built for sm_100a only (no fallback path)");
    return OMT_E_ARCH;
  }
  return OMT_OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64 && g_sms[dev] > 0) ? g_sms[dev] : 148;
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, whole row in registers (C <= 1024), two-pass statistics.
// ------------------------------------------------------------------------------------------
struct LnPlanes {          // optional fp16 hi / lo operand planes (omt_layernorm_h), written at the LOGICAL row
  uint16_t* y_hi; uint16_t* y_lo;    // normalised row
  uint16_t* x_hi; uint16_t* x_lo;    // raw input row (Attention.forward projects k, v from it)
  int lds;
  float* y_rs; float* x_rs;          // non-NULL: row-scaled planes (omt_common.cuh), the inverse row scale goes here
};

// PAIR (C a multiple of 256, NV even): a lane owns 8 consecutive columns per 256-column block (two adjacent float4 chunks), so
// the row-scaled planes leave as 16-byte stores (512 B per warp instruction) instead of 8-byte ones.
template <int NV, bool PAIR>   // float4 chunks per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int ldx,
                                                        float* __restrict__ y, int ldy,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ b, int M, int C,
                                                        float eps, int seg, int seg_stride, int seg_off,
                                                        const LnPlanes pl) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int lrow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (lrow >= M) return;
  const long long row = map_row(lrow, seg, seg_stride, seg_off);
  const float* xr = x + (size_t)row * ldx;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = PAIR ? ((i >> 1) * 32 + lane) * 8 + (i & 1) * 4 : (i * 32 + lane) * 4;
    if (c < C) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      if (pl.x_hi != nullptr && pl.x_rs == nullptr) store_split4(pl.x_hi, pl.x_lo, (size_t)lrow * pl.lds + c, v[i]);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (pl.x_hi != nullptr && pl.x_rs != nullptr) {      // row-scaled planes of the raw row
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) mx = fmaxf(mx, max4abs(v[i]));
    float sc, inv;
    row_scale(warp_max(mx), sc, inv);
    if constexpr (PAIR) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const int c = ((i >> 1) * 32 + lane) * 8;
        if (c < C) store_split8u(pl.x_hi, pl.x_lo, (size_t)lrow * pl.lds + c, v[i], v[i + 1], sc);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) store_split4u(pl.x_hi, pl.x_lo, (size_t)lrow * pl.lds + c, v[i], sc);
      }
    }
    if (lane == 0) pl.x_rs[lrow] = inv;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = PAIR ? ((i >> 1) * 32 + lane) * 8 + (i & 1) * 4 : (i * 32 + lane) * 4;
    if (c < C) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
  float* yr = y + (size_t)row * ldy;
  float omx = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = PAIR ? ((i >> 1) * 32 + lane) * 8 + (i & 1) * 4 : (i * 32 + lane) * 4;
    if (c < C) {
      const float4 g = *reinterpret_cast<const float4*>(w + c);
      float4 o;
      o.x = v[i].x * rstd * g.x; o.y = v[i].y * rstd * g.y;
      o.z = v[i].z * rstd * g.z; o.w = v[i].w * rstd * g.w;
      if (b != nullptr) {
        const float4 bb = *reinterpret_cast<const float4*>(b + c);
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (y != nullptr) *reinterpret_cast<float4*>(yr + c) = o;
      if (pl.y_hi != nullptr && pl.y_rs == nullptr) store_split4(pl.y_hi, pl.y_lo, (size_t)lrow * pl.lds + c, o);
      v[i] = o;                                        // kept for the row-scaled form below
      omx = fmaxf(omx, max4abs(o));
    }
  }
  if (pl.y_hi != nullptr && pl.y_rs != nullptr) {      // row-scaled planes of the normalised row
    float sc, inv;
    row_scale(warp_max(omx), sc, inv);
    if constexpr (PAIR) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const int c = ((i >> 1) * 32 + lane) * 8;
        if (c < C) store_split8u(pl.y_hi, pl.y_lo, (size_t)lrow * pl.lds + c, v[i], v[i + 1], sc);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) store_split4u(pl.y_hi, pl.y_lo, (size_t)lrow * pl.lds + c, v[i], sc);
      }
    }
    if (lane == 0) pl.y_rs[lrow] = inv;
  }
}

// ------------------------------------------------------------------------------------------
// Patch gather + LayerNorm.  One warp per patch row; K = Cin*p*p (first frame) or Cin*pt*p*p.
// Feature f = ((c*PT + dt)*p + p1)*p + p2 ; p2 is contiguous in the video (p % 4 == 0).
// ------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) patchify_ln_kernel(const float* __restrict__ video,
                                                          float* __restrict__ A, uint16_t* __restrict__ A_hi,
                                                          uint16_t* __restrict__ A_lo, float* __restrict__ A_rs,
                                                          const float* __restrict__ lw,
                                                          const float* __restrict__ lb, int rows,
                                                          int Cin, int T, int H, int W, int p, int pt,
                                                          int first, float eps) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int hh = H / p, ww = W / p;
  const int PT = first ? 1 : pt;
  const int K = Cin * PT * p * p;
  int r = row;
  const int wi = r % ww; r /= ww;
  const int hi = r % hh; r /= hh;
  int ti = 0;
  if (!first) { const int tn = (T - 1) / pt; ti = r % tn; r /= tn; }
  const int bi = r;
  const int t0 = first ? 0 : 1 + ti * pt;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int f = (i * 32 + lane) * 4;
    if (f < K) {
      const int p2 = f % p;
      const int p1 = (f / p) % p;
      const int dt = (f / (p * p)) % PT;
      const int c = f / (p * p * PT);
      const size_t off = ((((size_t)bi * Cin + c) * T + (t0 + dt)) * H + (hi * p + p1)) * W + wi * p + p2;
      v[i] = *reinterpret_cast<const float4*>(video + off);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const size_t rbase = (size_t)row * K;
  // write a finished row: fp32, 2^11-scaled planes, or row-scaled planes (+ the inverse row scale)
  auto emit = [&](float4 (&o)[NV]) {
    float sc = 1.f, inv = 1.f;
    if (A_rs != nullptr) {
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) mx = fmaxf(mx, max4abs(o[i]));
      row_scale(warp_max(mx), sc, inv);
      if (lane == 0) A_rs[row] = inv;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = (i * 32 + lane) * 4;
      if (f < K) {
        if (A_rs != nullptr) store_split4u(A_hi, A_lo, rbase + f, o[i], sc);
        else if (A_hi != nullptr) store_split4(A_hi, A_lo, rbase + f, o[i]);
        else *reinterpret_cast<float4*>(A + rbase + f) = o[i];
      }
    }
  };
  if (lw == nullptr) {        // plain im2col (patch_embed='cnn': the strided Conv3d is a GEMM on raw patch vectors)
    emit(v);
    return;
  }
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int f = (i * 32 + lane) * 4;
    if (f < K) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int f = (i * 32 + lane) * 4;
    if (f < K) {
      const float4 g = *reinterpret_cast<const float4*>(lw + f);
      const float4 bb = *reinterpret_cast<const float4*>(lb + f);
      float4 o;
      o.x = v[i].x * rstd * g.x + bb.x; o.y = v[i].y * rstd * g.y + bb.y;
      o.z = v[i].z * rstd * g.z + bb.z; o.w = v[i].w * rstd * g.w + bb.w;
      v[i] = o;
    }
  }
  emit(v);
}

__global__ void __launch_bounds__(256) unpatchify_kernel(const float* __restrict__ P,
                                                         float* __restrict__ video, long long total4,
                                                         int Cin, int T, int H, int W, int p, int pt,
                                                         int first) {
  pdl_sync();
  const int hh = H / p, ww = W / p;
  const int PT = first ? 1 : pt;
  const int K4 = Cin * PT * p * p / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % K4) * 4;
    int r = (int)(i / K4);
    const int wi = r % ww; r /= ww;
    const int hi = r % hh; r /= hh;
    int ti = 0;
    if (!first) { const int tn = (T - 1) / pt; ti = r % tn; r /= tn; }
    const int bi = r;
    const int t0 = first ? 0 : 1 + ti * pt;
    const int p2 = f % p;
    const int p1 = (f / p) % p;
    const int dt = (f / (p * p)) % PT;
    const int c = f / (p * p * PT);
    const size_t off = ((((size_t)bi * Cin + c) * T + (t0 + dt)) * H + (hi * p + p1)) * W + wi * p + p2;
    *reinterpret_cast<float4*>(video + off) = *reinterpret_cast<const float4*>(P + i * 4);
  }
}

// ------------------------------------------------------------------------------------------
// Un-patchify fused with the consumer's uint8 conversion (vqgan_eval.py:139,147-148; Latte sample_ddp.py:206;
// DiT sample_ddp.py:163):  u8 = trunc( clamp(x * mul + add, lo, hi) * post )  written channels-LAST
// (b, t, H, W, c) -- the layout every consumer permutes to before .byte().  mul / add / post are applied as separate
// fp32 roundings (no fma contraction) so the bytes equal torch's elementwise expression on the fp32 reconstruction.
// One thread = 4 consecutive pixels of one patch line, all Cin channels (Cin * 4 bytes contiguous in the output).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpatchify_u8_kernel(const float* __restrict__ P, uint8_t* __restrict__ out,
                                                            long long total, int Cin, int T, int H, int W, int p,
                                                            int pt, int first, float mul, float add, float lo,
                                                            float hi, float post) {
  pdl_sync();
  const int hh = H / p, ww = W / p;
  const int PT = first ? 1 : pt;
  const int per_row = PT * p * (p / 4);           // (dt, p1, p2-quad) items per patch row
  const int K = Cin * PT * p * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int it = (int)(i % per_row);
    int r = (int)(i / per_row);
    const int row = r;
    const int q4 = it % (p / 4); it /= (p / 4);
    const int p1 = it % p;
    const int dt = it / p;
    const int wi = r % ww; r /= ww;
    const int hi_ = r % hh; r /= hh;
    int ti = 0;
    if (!first) { const int tn = (T - 1) / pt; ti = r % tn; r /= tn; }
    const int bi = r;
    const int t = first ? 0 : 1 + ti * pt + dt;
    const size_t pix = (((size_t)bi * T + t) * H + (hi_ * p + p1)) * W + wi * p + q4 * 4;
    uint8_t* o = out + pix * Cin;
    for (int c = 0; c < Cin; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(P + (size_t)row * K + ((c * PT + dt) * p + p1) * p + q4 * 4);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float tv = __fadd_rn(__fmul_rn(e[j], mul), add);
        tv = fminf(fmaxf(tv, lo), hi);
        o[j * Cin + c] = (uint8_t)__float2uint_rz(__fmul_rn(tv, post));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// PEG: y = x + bias + sum_k w[k] * x[nbr[k]].  Block = C/4 threads (one float4 channel group
// each), loops over ROWS consecutive rows; the 27 x C weight table sits in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int PEG_ROWS = 16;

__global__ void __launch_bounds__(128) peg_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                  const float* __restrict__ w27,
                                                  const float* __restrict__ bias,
                                                  const int32_t* __restrict__ nbr, int rows_per_b,
                                                  int C, long long M) {
  pdl_sync();
  extern __shared__ float4 wsm[];   // [27][C/4]
  __shared__ int32_t nsm[PEG_ROWS][27];
  const int c4 = threadIdx.x;       // channel group
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < 27 * C4; i += blockDim.x)
    wsm[i] = reinterpret_cast<const float4*>(w27)[i];
  const long long r0 = (long long)blockIdx.x * PEG_ROWS;
  for (int i = threadIdx.x; i < PEG_ROWS * 27; i += blockDim.x) {
    const long long r = r0 + i / 27;
    nsm[i / 27][i % 27] = (r < M) ? nbr[(r % rows_per_b) * 27 + (i % 27)] : -1;
  }
  __syncthreads();
  if (c4 >= C4) return;
  const float4 bb = reinterpret_cast<const float4*>(bias)[c4];
  for (int rr = 0; rr < PEG_ROWS; ++rr) {
    const long long r = r0 + rr;
    if (r >= M) break;
    const long long base = (r / rows_per_b) * rows_per_b;
    const float4 xv = reinterpret_cast<const float4*>(x + r * C)[c4];
    float4 acc = bb;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const int n = nsm[rr][k];
      if (n >= 0) {
        const float4 nv = __ldg(reinterpret_cast<const float4*>(x + (base + n) * C) + c4);
        const float4 wv = wsm[k * C4 + c4];
        acc.x = fmaf(nv.x, wv.x, acc.x); acc.y = fmaf(nv.y, wv.y, acc.y);
        acc.z = fmaf(nv.z, wv.z, acc.z); acc.w = fmaf(nv.w, wv.w, acc.w);
      }
    }
    acc.x += xv.x; acc.y += xv.y; acc.z += xv.z; acc.w += xv.w;
    reinterpret_cast<float4*>(y + r * C)[c4] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// PEG, tiled form.  The stencil lives in "volume space" (t2,h2,w2): for spatial transformers that
// is the true token grid; for temporal ones it is the reference's literal reshape of the
// '(b h w) t d' tensor (flat f = n*T + tau unravelled over (T,h,w)).  Either way volume position f
// maps to canonical row  temporal ? (f % T) * N + f / T : f.
// One CTA stages a (TT+2) x (HB+2) x (w+2) x 16-channel halo tile in shared memory (zeros where the
// reference pads), then every thread owns one (plane, row, channel-pair) strip and slides a
// 3x3x3 register window along w: 9 shared loads per output instead of 27 global ones.
// ------------------------------------------------------------------------------------------
constexpr int PEG_CC = 16;       // channels per CTA

__global__ void __launch_bounds__(256) peg_tile_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       const float* __restrict__ w27,
                                                       const float* __restrict__ bias, int T, int h, int w,
                                                       int C, int temporal, int causal, int TT, int HB, int RS) {
  pdl_sync();
  extern __shared__ __align__(16) float tile[];      // [(TT+2)][(HB+2)] rows of RS floats ((w+2)*16 + pad)
  const int N = h * w;
  const int n_hblk = (h + HB - 1) / HB;
  const int t0 = (blockIdx.x / n_hblk) * TT, h0 = (blockIdx.x % n_hblk) * HB;
  const int c0 = blockIdx.y * PEG_CC;
  const long long bbase = (long long)blockIdx.z * T * N;
  const int pad_lo = causal ? 2 : 1;
  const int rows = (TT + 2) * (HB + 2);
  const int nth = blockDim.x;
  // ---- halo tile: the (few) integer divisions happen once per (plane,row) pair and once per position,
  //      not once per 16-byte load: int2 {global row or -1, smem float offset} per position
  int2* pmap = reinterpret_cast<int2*>(tile + rows * RS);
  const int P = rows * (w + 2);
  for (int pos = threadIdx.x; pos < P; pos += nth) {
    const int pw = pos % (w + 2);
    const int pr = pos / (w + 2);
    const int ph = pr % (HB + 2), pt = pr / (HB + 2);
    const int t2 = t0 - pad_lo + pt, h2 = h0 - 1 + ph, w2 = pw - 1;
    int row = -1;
    if (t2 >= 0 && t2 < T && h2 >= 0 && h2 < h && w2 >= 0 && w2 < w) {
      const int f = (t2 * h + h2) * w + w2;
      row = temporal ? (f % T) * N + f / T : f;
    }
    pmap[pos] = make_int2(row, pr * RS + pw * PEG_CC);
  }
  __syncthreads();
  // gathers are issued in batches of 8 per thread before any shared store so ~8 x 16 B are in flight per thread
  for (int i0 = threadIdx.x; i0 < P * 4; i0 += nth * 8) {
    float4 v[8];
    int so[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * nth;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      so[u] = -1;
      if (i < P * 4) {
        const int2 pm = pmap[i >> 2];
        so[u] = pm.y + (i & 3) * 4;
        if (pm.x >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(x + (bbase + pm.x) * C + c0) + (i & 3));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (so[u] >= 0) *reinterpret_cast<float4*>(tile + so[u]) = v[u];
  }
  __syncthreads();
  // ---- strips
  const int cp = threadIdx.x & 7;                  // channel pair inside the 16-channel slab
  const int strip = threadIdx.x >> 3;
  const int sh = strip % HB, st = strip / HB;
  if (st >= TT || t0 + st >= T || h0 + sh >= h) return;
  float2 wt[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) wt[k] = *reinterpret_cast<const float2*>(w27 + (size_t)k * C + c0 + 2 * cp);
  const float2 bb = *reinterpret_cast<const float2*>(bias + c0 + 2 * cp);
  float2 win[9][3];
  const float* tp = tile + 2 * cp;
#pragma unroll
  for (int r9 = 0; r9 < 9; ++r9) {
    const float* rp = tp + ((st + r9 / 3) * (HB + 2) + sh + r9 % 3) * RS;
    win[r9][1] = *reinterpret_cast<const float2*>(rp);
    win[r9][2] = *reinterpret_cast<const float2*>(rp + PEG_CC);
  }
  const int fbase = ((t0 + st) * h + (h0 + sh)) * w;
  int tau = fbase % T, nn = fbase / T;               // temporal: volume position f <-> canonical (tau, n), advanced incrementally
  for (int w2 = 0; w2 < w; ++w2) {
    float2 acc = bb;
#pragma unroll
    for (int r9 = 0; r9 < 9; ++r9) {
      const float* rp = tp + ((st + r9 / 3) * (HB + 2) + sh + r9 % 3) * RS + (w2 + 2) * PEG_CC;
      win[r9][0] = win[r9][1]; win[r9][1] = win[r9][2];
      win[r9][2] = *reinterpret_cast<const float2*>(rp);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        acc.x = fmaf(win[r9][kw].x, wt[r9 * 3 + kw].x, acc.x);
        acc.y = fmaf(win[r9][kw].y, wt[r9 * 3 + kw].y, acc.y);
      }
    }
    const float2 ctr = causal ? win[7][1] : win[4][1];   // the un-shifted token itself (residual)
    acc.x += ctr.x; acc.y += ctr.y;
    const int row = temporal ? tau * N + nn : fbase + w2;
    *reinterpret_cast<float2*>(y + (bbase + row) * C + c0 + 2 * cp) = acc;
    if (++tau == T) { tau = 0; ++nn; }
  }
}

// ------------------------------------------------------------------------------------------
// PEG, tiled form v4: same tile geometry and the SAME fma order as peg_tile_kernel (bit-identical output), with
// the instruction count cut ~3x (the v3 kernel is issue-bound: ncu 67 M warp instructions, 102 per output):
//   * halo gather by cp.async (16 B, zero-fill where the reference pads): no register staging, no position-map
//     pass; one warp walks one halo row at a time so the only divisions are per row (warp-uniform) and the
//     temporal  f -> (f % T, f / T)  split is a multiply-shift on the in-row offset;
//   * the 3x3x3 register window rotates by renaming (w loop unrolled by 3) instead of 36 MOVs per output;
//   * packed fma.rn.f32x2 (FFMA2): one instruction per channel PAIR and tap.
// Requires T <= 64, w <= 254 (multiply-shift range) and 16-byte aligned x; the host falls back to v3 otherwise.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
// One output position of a strip.  R = window rotation: tap kw of window row r9 lives in slot (kw + R) % 3;
// the new halo column is read at a[r9] + R * 64 bytes (a[] points at the trip's first new column).
template <int R>
__device__ __forceinline__ float2 peg_step(float2 (&win)[9][3], const float2 (&wt)[27], const float2 bb,
                                           const uint32_t (&a)[9], const bool causal) {
  float2 acc = bb;
#pragma unroll
  for (int r9 = 0; r9 < 9; ++r9) {
    win[r9][(2 + R) % 3] = lds_f2(a[r9] + R * (PEG_CC * 4));
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) acc = ffma2(win[r9][(kw + R) % 3], wt[r9 * 3 + kw], acc);
  }
  const float2 ctr = causal ? win[7][(1 + R) % 3] : win[4][(1 + R) % 3];   // the un-shifted token itself (residual)
  acc.x += ctr.x; acc.y += ctr.y;
  return acc;
}

__global__ void __launch_bounds__(160, 3) peg_tile4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ w27,
                                                           const float* __restrict__ bias, int T, int h, int w,
                                                           int C, int temporal, int causal, int TT, int HB, int RS, int zrow) {
  pdl_sync();
  // [valid planes of the tile][(HB+2)] rows of RS floats ((w+2)*16 + pad), then ONE all-zero row at index zrow: planes
  // outside the volume (the causal pad in front, the halo behind the last plane) are not stored -- every window row that
  // falls into one reads the zero row instead.  5 of 7 planes at T' = 5: 69 KB instead of 94 KB, three CTAs per SM.
  extern __shared__ __align__(16) float tile[];
  const int N = h * w;
  const int n_hblk = (h + HB - 1) / HB;
  const int t0 = (blockIdx.x / n_hblk) * TT, h0 = (blockIdx.x % n_hblk) * HB;
  const int c0 = blockIdx.y * PEG_CC;
  const long long bbase = (long long)blockIdx.z * T * N;
  const int pad_lo = causal ? 2 : 1;
  const int rows = (TT + 2) * (HB + 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t inv_T = (65536u + (uint32_t)T - 1u) / (uint32_t)T;       // floor(v / T) == (v * inv_T) >> 16 for v < 65536 / T
  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  const int chunks = (w + 2) * 4;                                          // 16-byte chunks per halo row
  const int pz_lo = t0 < pad_lo ? pad_lo - t0 : 0;                         // leading planes of the tile that lie before t = 0
  for (int i = threadIdx.x; i < RS / 4; i += blockDim.x)
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(tile_s + (uint32_t)(zrow * RS + 4 * i) * 4u), "f"(0.f) : "memory");
  // ---- halo tile: warp <-> halo row; lane <-> 16-byte chunk (consecutive lanes write consecutive shared addresses)
  for (int pr = warp; pr < rows; pr += nwarps) {
    const int ph = pr % (HB + 2), pt = pr / (HB + 2);
    const int t2 = t0 - pad_lo + pt, h2 = h0 - 1 + ph;
    if (t2 < 0 || t2 >= T) continue;                                       // plane outside the volume: not stored
    const bool row_ok = h2 >= 0 && h2 < h;
    const int fb = row_ok ? (t2 * h + h2) * w : 0;                         // volume position of (t2, h2, w2 = 0)
    const int tau0 = fb % T, nn0 = fb / T;
    const uint32_t dst_row = tile_s + (uint32_t)(((pt - pz_lo) * (HB + 2) + ph) * RS) * 4u;
    for (int j = lane; j < chunks; j += 32) {
      const int w2 = (j >> 2) - 1;
      const bool ok = row_ok && w2 >= 0 && w2 < w;
      long long row = 0;
      if (ok) {
        if (temporal) {
          const uint32_t v = (uint32_t)(tau0 + w2);
          const uint32_t q = (v * inv_T) >> 16;
          row = (long long)(v - q * (uint32_t)T) * N + nn0 + (int)q;
        } else {
          row = fb + w2;
        }
      }
      const float* src = x + (bbase + row) * C + c0 + (j & 3) * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_row + (uint32_t)j * 16u), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- strips: weights are fetched while the gather is in flight
  const int cp = threadIdx.x & 7;                  // channel pair inside the 16-channel slab
  const int strip = threadIdx.x >> 3;
  const int sh = strip % HB, st = strip / HB;
  const bool active = st < TT && t0 + st < T && h0 + sh < h;
  float2 wt[27];
  float2 bb = make_float2(0.f, 0.f);
  if (active) {
#pragma unroll
    for (int k = 0; k < 27; ++k) wt[k] = __ldg(reinterpret_cast<const float2*>(w27 + (size_t)k * C + c0 + 2 * cp));
    bb = __ldg(reinterpret_cast<const float2*>(bias + c0 + 2 * cp));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!active) return;
  // shared byte addresses of halo column 0 of the 9 window rows of this strip
  uint32_t a[9];
#pragma unroll
  for (int r9 = 0; r9 < 9; ++r9) {
    const int pt = st + r9 / 3;                     // tile plane of this window row
    const int t2 = t0 - pad_lo + pt;
    const int prow = (t2 < 0 || t2 >= T) ? zrow : (pt - pz_lo) * (HB + 2) + sh + r9 % 3;
    a[r9] = tile_s + (uint32_t)((prow * RS + 2 * cp) * 4);
  }
  float2 win[9][3];
#pragma unroll
  for (int r9 = 0; r9 < 9; ++r9) {
    win[r9][0] = lds_f2(a[r9]);                     // halo column 0 (w2 = -1)
    win[r9][1] = lds_f2(a[r9] + PEG_CC * 4);        // halo column 1 (w2 = 0)
    a[r9] += 2 * PEG_CC * 4;                        // -> the first new column of trip 0 (halo column 2)
  }
  // output pointer, advanced incrementally: spatial rows are consecutive; temporal rows follow the literal
  // reshape  f -> (tau, n) = (f % T, f / T)  ->  canonical row tau * N + n
  const int fbase = ((t0 + st) * h + (h0 + sh)) * w;
  int tau = fbase % T;
  const long long row0 = temporal ? (long long)tau * N + fbase / T : (long long)fbase;
  float* yp = y + (bbase + row0) * C + c0 + 2 * cp;
  const long long inc = temporal ? (long long)N * C : (long long)C;
  const long long wrap = (long long)T * N * C - C;   // temporal: tau T-1 -> 0 moves back T planes and on one token
  const bool cz = causal != 0;
#define OMT_PEG_STEP(R)                                                                             \
  {                                                                                                 \
    const float2 acc = peg_step<R>(win, wt, bb, a, cz);                                             \
    *reinterpret_cast<float2*>(yp) = acc;                                                           \
    yp += inc;                                                                                      \
    if (temporal && ++tau == T) { tau = 0; yp -= wrap; }                                            \
  }
  for (int wb = 0; wb < w; wb += 3) {
    OMT_PEG_STEP(0)
    if (wb + 1 < w) OMT_PEG_STEP(1)
    if (wb + 2 < w) OMT_PEG_STEP(2)
#pragma unroll
    for (int r9 = 0; r9 < 9; ++r9) a[r9] += 3 * PEG_CC * 4;
  }
#undef OMT_PEG_STEP
}

// ------------------------------------------------------------------------------------------
// rope + l2norm + scale, in place on q and k.  One warp per row; lane l owns the complex pair
// (2l, 2l+1) of every head.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void qk_prep_one(float* p, const float2 cs, bool rope, const float2 sc) {
  float2 v = *reinterpret_cast<float2*>(p);
  if (rope) {
    const float a = v.x * cs.x - v.y * cs.y;
    const float b = v.x * cs.y + v.y * cs.x;
    v.x = a; v.y = b;
  }
  const float ss = warp_sum(v.x * v.x + v.y * v.y);
  const float den = fmaxf(sqrtf(ss), 1e-12f);
  v.x = v.x / den * sc.x;
  v.y = v.y / den * sc.y;
  *reinterpret_cast<float2*>(p) = v;
}

__global__ void __launch_bounds__(256) qk_prep_kernel(float* __restrict__ q, int ldq,
                                                      float* __restrict__ k, int ldk,
                                                      const float* __restrict__ qs,
                                                      const float* __restrict__ ks,
                                                      const float* __restrict__ rc,
                                                      const float* __restrict__ rs, int M, int N,
                                                      int heads) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const bool rope = rc != nullptr;
  float2 cs = make_float2(1.f, 0.f);
  if (rope) {
    const int pos = row % N;
    cs.x = rc[pos * 32 + lane];
    cs.y = rs[pos * 32 + lane];
  }
  const float2 sq = *reinterpret_cast<const float2*>(qs + 2 * lane);
  const float2 sk = *reinterpret_cast<const float2*>(ks + 2 * lane);
  for (int h = 0; h < heads; ++h) {
    qk_prep_one(q + (size_t)row * ldq + h * 64 + 2 * lane, cs, rope, sq);
    qk_prep_one(k + (size_t)row * ldk + h * 64 + 2 * lane, cs, rope, sk);
  }
}

}  // namespace omt

using namespace omt;

extern "C" int omt_abi_version(void) { return OMT_ABI_VERSION; }
extern "C" const char* omt_last_error(void) { return omt::g_err; }

extern "C" int omt_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  OMT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  OMT_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return OMT_OK;
}

static int layernorm_impl(const char* who, const float* x, int ldx, float* y, int ldy, const omt::LnPlanes& pl, const float* w,
                          const float* b, int M, int C, float eps, int seg, int seg_stride, int seg_off, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(x && w && (y || pl.y_hi), "%s: null pointer", who);
  OMT_REQUIRE(M >= 0 && C > 0 && C % 4 == 0 && C <= 1024, "%s: C=%d must be a multiple of 4, <= 1024", who, C);
  OMT_REQUIRE(ldx % 4 == 0 && ldx >= C && (y == nullptr || (ldy % 4 == 0 && ldy >= C)), "%s: bad leading dims", who);
  OMT_REQUIRE((pl.y_hi == nullptr) == (pl.y_lo == nullptr) && (pl.x_hi == nullptr) == (pl.x_lo == nullptr), "%s: planes come in hi / lo pairs", who);
  if (pl.y_hi != nullptr || pl.x_hi != nullptr) {
    OMT_REQUIRE(pl.lds % 4 == 0 && pl.lds >= C, "%s: plane leading dimension %d", who, pl.lds);
    OMT_REQUIRE(((uintptr_t)pl.y_hi | (uintptr_t)pl.y_lo | (uintptr_t)pl.x_hi | (uintptr_t)pl.x_lo) % 8 == 0, "%s: planes must be 8-byte aligned", who);
  }
  if (M == 0) return OMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (C / 4 + 31) / 32;
  // 8 consecutive columns per lane (16-byte plane stores) when the row splits into whole 256-column blocks and the planes allow it
  const bool pair = C % 256 == 0 && pl.lds % 8 == 0 &&
                    ((uintptr_t)pl.y_hi | (uintptr_t)pl.y_lo | (uintptr_t)pl.x_hi | (uintptr_t)pl.x_lo) % 16 == 0;
  dim3 grid((M + 7) / 8), block(256);
  switch (nv) {
    case 1: OMT_CUDA(launch_k(layernorm_kernel<1, false>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl)); break;
    case 2:
      if (pair) OMT_CUDA(launch_k(layernorm_kernel<2, true>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl));
      else OMT_CUDA(launch_k(layernorm_kernel<2, false>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl));
      break;
    case 3: OMT_CUDA(launch_k(layernorm_kernel<3, false>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl)); break;
    case 4:
      if (pair) OMT_CUDA(launch_k(layernorm_kernel<4, true>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl));
      else OMT_CUDA(launch_k(layernorm_kernel<4, false>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl));
      break;
    default: OMT_CUDA(launch_k(layernorm_kernel<8, false>, grid, block, 0, st, x, ldx, y, ldy, w, b, M, C, eps, seg, seg_stride, seg_off, pl)); break;
  }
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_layernorm(const float* x, int ldx, float* y, int ldy, const float* w, const float* b,
                             int M, int C, float eps, int seg, int seg_stride, int seg_off,
                             omt_stream_t stream) {
  OMT_REQUIRE(y != nullptr, "omt_layernorm: null pointer");
  omt::LnPlanes pl{nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  return layernorm_impl("omt_layernorm", x, ldx, y, ldy, pl, w, b, M, C, eps, seg, seg_stride, seg_off, stream);
}

extern "C" int omt_layernorm_h(const float* x, int ldx, float* y, int ldy, uint16_t* y_hi, uint16_t* y_lo, float* y_rs,
                               uint16_t* x_hi, uint16_t* x_lo, float* x_rs, int lds, const float* w, const float* b,
                               int M, int C, float eps, int seg, int seg_stride, int seg_off, omt_stream_t stream) {
  OMT_REQUIRE((y_rs == nullptr || y_hi != nullptr) && (x_rs == nullptr || x_hi != nullptr), "omt_layernorm_h: row scales without planes");
  omt::LnPlanes pl{y_hi, y_lo, x_hi, x_lo, lds, y_rs, x_rs};
  return layernorm_impl("omt_layernorm_h", x, ldx, y, ldy, pl, w, b, M, C, eps, seg, seg_stride, seg_off, stream);
}

extern "C" int omt_patchify_ln(const float* video, float* A, uint16_t* A_hi, uint16_t* A_lo, float* A_rs, const float* ln_w,
                               const float* ln_b, int B, int Cin, int T, int H, int W, int p, int pt, int first,
                               float eps, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(video && (A || A_hi) && ((ln_w == nullptr) == (ln_b == nullptr)) && ((A_hi == nullptr) == (A_lo == nullptr)),
              "omt_patchify_ln: null pointer");
  OMT_REQUIRE(((uintptr_t)A_hi | (uintptr_t)A_lo) % 8 == 0, "omt_patchify_ln: planes must be 8-byte aligned");
  OMT_REQUIRE(A_rs == nullptr || A_hi != nullptr, "omt_patchify_ln: row scales without planes");
  OMT_REQUIRE(p % 4 == 0 && H % p == 0 && W % p == 0, "omt_patchify_ln: patch %d must be a multiple of 4 dividing %dx%d", p, H, W);
  OMT_REQUIRE(first || (T > 1 && (T - 1) % pt == 0), "omt_patchify_ln: (T-1) %% pt != 0");
  const int PT = first ? 1 : pt;
  const int K = Cin * PT * p * p;
  OMT_REQUIRE(K <= 1024, "omt_patchify_ln: patch vector %d > 1024", K);
  const long long rows = (long long)B * (first ? 1 : (T - 1) / pt) * (H / p) * (W / p);
  if (rows == 0) return OMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((rows + 7) / 8)), block(256);
  const int nv = (K / 4 + 31) / 32;
  if (nv <= 2)
    OMT_CUDA(launch_k(patchify_ln_kernel<2>, grid, block, 0, st, video, A, A_hi, A_lo, A_rs, ln_w, ln_b, (int)rows, Cin, T, H, W, p, pt, first, eps));
  else if (nv <= 6)
    OMT_CUDA(launch_k(patchify_ln_kernel<6>, grid, block, 0, st, video, A, A_hi, A_lo, A_rs, ln_w, ln_b, (int)rows, Cin, T, H, W, p, pt, first, eps));
  else
    OMT_CUDA(launch_k(patchify_ln_kernel<8>, grid, block, 0, st, video, A, A_hi, A_lo, A_rs, ln_w, ln_b, (int)rows, Cin, T, H, W, p, pt, first, eps));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_unpatchify(const float* P, float* video, int B, int Cin, int T, int H, int W, int p,
                              int pt, int first, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(P && video, "omt_unpatchify: null pointer");
  OMT_REQUIRE(p % 4 == 0 && H % p == 0 && W % p == 0, "omt_unpatchify: bad patch size");
  OMT_REQUIRE(first || (T > 1 && (T - 1) % pt == 0), "omt_unpatchify: (T-1) %% pt != 0");
  const int PT = first ? 1 : pt;
  const long long rows = (long long)B * (first ? 1 : (T - 1) / pt) * (H / p) * (W / p);
  const long long total4 = rows * (Cin * PT * p * p / 4);
  if (total4 == 0) return OMT_OK;
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  OMT_CUDA(launch_k(unpatchify_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, P, video, total4, Cin, T, H, W, p, pt, first));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_unpatchify_u8(const float* P, uint8_t* out, int B, int Cin, int T, int H, int W, int p, int pt,
                                 int first, float mul, float add, float lo, float hi, float post, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(P && out, "omt_unpatchify_u8: null pointer");
  OMT_REQUIRE(p % 4 == 0 && H % p == 0 && W % p == 0, "omt_unpatchify_u8: bad patch size");
  OMT_REQUIRE(first || (T > 1 && (T - 1) % pt == 0), "omt_unpatchify_u8: (T-1) %% pt != 0");
  OMT_REQUIRE(lo >= 0.f && hi * post < 256.f, "omt_unpatchify_u8: clamp range [%g, %g] x %g does not fit a byte", lo, hi, post);
  const int PT = first ? 1 : pt;
  const long long rows = (long long)B * (first ? 1 : (T - 1) / pt) * (H / p) * (W / p);
  const long long total = rows * PT * p * (p / 4);
  if (total == 0) return OMT_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  OMT_CUDA(launch_k(unpatchify_u8_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, P, out, total, Cin, T, H, W, p,
                    pt, first, mul, add, lo, hi, post));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_peg(const float* x, float* y, const float* w27, const float* bias, const int32_t* nbr,
                       int B, int rows_per_b, int C, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(x && y && w27 && bias && nbr, "omt_peg: null pointer");
  OMT_REQUIRE(x != y, "omt_peg: in-place is not supported (stencil)");
  OMT_REQUIRE(C % 4 == 0 && C / 4 <= 128, "omt_peg: C=%d unsupported (need C %% 4 == 0, C <= 512)", C);
  const long long M = (long long)B * rows_per_b;
  if (M == 0) return OMT_OK;
  static bool attr_set[64];          // the attribute is per device
  const size_t smem = (size_t)27 * C * sizeof(float);
  int dev0 = 0;
  cudaGetDevice(&dev0);
  if (dev0 >= 0 && dev0 < 64 && !attr_set[dev0]) {
    OMT_CUDA(cudaFuncSetAttribute(peg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 27 * 512 * 4));
    attr_set[dev0] = true;
  }
  const unsigned blocks = (unsigned)((M + PEG_ROWS - 1) / PEG_ROWS);
  peg_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(x, y, w27, bias, nbr, rows_per_b, C, M);
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_peg_volume(const float* x, float* y, const float* w27, const float* bias, int B, int T, int h,
                              int w, int C, int temporal, int causal, omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(x && y && w27 && bias, "omt_peg_volume: null pointer");
  OMT_REQUIRE(x != y, "omt_peg_volume: in-place is not supported (stencil)");
  OMT_REQUIRE(C % PEG_CC == 0 && C / PEG_CC <= 65535 && B <= 65535, "omt_peg_volume: C=%d must be a multiple of 16", C);
  OMT_REQUIRE(T >= 1 && h >= 1 && w >= 1, "omt_peg_volume: bad volume");
  if (B == 0) return OMT_OK;
  // tile geometry: planes per CTA (TT) and rows per CTA (HB) so that threads <= 256 and smem <= ~100 KB
  int RS = (w + 2) * PEG_CC;
  RS += ((16 - RS % 32) + 32) % 32;                  // row stride == 16 (mod 32) floats: 2-way minimum bank pattern
  int TT = T < 5 ? T : 5, HB = 4;
  auto smem_of = [&](int tt, int hb) { return (size_t)(tt + 2) * (hb + 2) * (RS * sizeof(float) + (size_t)(w + 2) * 8); };
  while (HB > 1 && (smem_of(TT, HB) > 112 * 1024 || TT * HB * 8 > 256)) --HB;
  while (TT > 1 && (smem_of(TT, HB) > 112 * 1024 || TT * HB * 8 > 256)) --TT;
  const size_t smem = smem_of(TT, HB);
  OMT_REQUIRE(smem <= 200 * 1024, "omt_peg_volume: row of %d tokens does not fit the shared-memory tile", w);
  static size_t smem_set[64];        // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && smem > smem_set[dev]) {
    OMT_CUDA(cudaFuncSetAttribute(peg_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[dev] = smem;
  }
  const int threads = ((TT * HB * 8 + 31) / 32) * 32;
  dim3 grid(((T + TT - 1) / TT) * ((h + HB - 1) / HB), C / PEG_CC, B);
  const bool fast_ok = T <= 64 && w <= 254 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const bool v4 = g_peg_kernel == 4 && fast_ok;
  if (v4) {
    // planes of a tile that lie inside the volume (the others are one shared zero row): the maximum over the t-blocks
    const int pad_lo = causal ? 2 : 1;
    int vp = 1;
    for (int t0 = 0; t0 < T; t0 += TT) {
      const int lo = t0 - pad_lo < 0 ? 0 : t0 - pad_lo, hi = t0 - pad_lo + TT + 2 > T ? T : t0 - pad_lo + TT + 2;
      if (hi - lo > vp) vp = hi - lo;
    }
    const int zrow = vp * (HB + 2);
    const size_t smem4 = (size_t)(zrow + 1) * RS * sizeof(float);
    static size_t smem4_set[64];
    if (dev >= 0 && dev < 64 && smem4 > smem4_set[dev]) {
      OMT_CUDA(cudaFuncSetAttribute(peg_tile4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
      smem4_set[dev] = smem4;
    }
    OMT_CUDA(launch_k(peg_tile4_kernel, grid, dim3(threads), smem4, (cudaStream_t)stream, x, y, w27, bias, T, h, w, C, temporal, causal, TT, HB, RS, zrow));
  } else {
    OMT_CUDA(launch_k(peg_tile_kernel, grid, dim3(threads), smem, (cudaStream_t)stream, x, y, w27, bias, T, h, w, C, temporal, causal, TT, HB, RS));
  }
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

extern "C" int omt_qk_prep(float* q, int ldq, float* k, int ldk, const float* q_scale, const float* k_scale,
                           const float* rope_cos, const float* rope_sin, int M, int N, int heads,
                           omt_stream_t stream) {
  OMT_ENTER();
  OMT_REQUIRE(q && k && q_scale && k_scale, "omt_qk_prep: null pointer");
  OMT_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "omt_qk_prep: cos/sin must both be given");
  OMT_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && N > 0, "omt_qk_prep: bad leading dims");
  if (M == 0) return OMT_OK;
  OMT_CUDA(launch_k(qk_prep_kernel, dim3((M + 7) / 8), dim3(256), 0, (cudaStream_t)stream, q, ldq, k, ldk, q_scale, k_scale,
                    rope_cos, rope_sin, M, N, heads));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

namespace omt { extern int g_attn_kernel; extern int g_f16_bn; extern int g_attn_f16_ctas; }

extern "C" int omt_set_option(const char* name, int value) {
  if (name == nullptr) return OMT_E_ARG;
  if (strcmp(name, "pdl") == 0) { omt::g_pdl = value ? 1 : 0; return OMT_OK; }
  if (strcmp(name, "peg_kernel") == 0) {
    if (value != 3 && value != 4) { omt::set_error("peg_kernel must be 3 or 4"); return OMT_E_ARG; }
    omt::g_peg_kernel = value;
    return OMT_OK;
  }
  if (strcmp(name, "attn_kernel") == 0) {
    if (value != 1 && value != 3) { omt::set_error("attn_kernel must be 1 or 3"); return OMT_E_ARG; }
    omt::g_attn_kernel = value;
    return OMT_OK;
  }
  if (strcmp(name, "f16_bn") == 0) {
    if (value != 0 && value != 128 && value != 256) { omt::set_error("f16_bn must be 0, 128 or 256"); return OMT_E_ARG; }
    omt::g_f16_bn = value;
    return OMT_OK;
  }
  if (strcmp(name, "attn_f16_ctas") == 0) {
    if (value != 1 && value != 2) { omt::set_error("attn_f16_ctas must be 1 or 2"); return OMT_E_ARG; }
    omt::g_attn_f16_ctas = value;
    return OMT_OK;
  }
  omt::set_error("unknown option %s", name);
  return OMT_E_ARG;
}
