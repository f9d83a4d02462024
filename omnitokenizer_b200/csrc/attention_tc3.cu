// Attention v3: spatial (full, non-causal) attention core on tcgen05 with 3xTF32 compensation and
// TMEM-resident A operands (tcgen05.mma "TS" form).  Same contract as attention_tc.cu:
//   O = softmax(scale * Q K^T) V  per (sequence, head), head dim 64, N % 128 == 0
//   (F.scaled_dot_product_attention at modules/attention.py:451).
//
// Why: with A = Q / P in shared memory every N=64 MMA re-reads a 4 KiB A tile, so the MMAs are
// shared-memory-bandwidth bound (6 KiB per ~32-cycle instruction; measured 814 us per layer).  Here Q
// (once per CTA) and P (every key tile) live in TENSOR MEMORY as tf32 hi / lo column blocks and are
// consumed as the TMEM A operand; only the K / V^T B tiles (2 KiB per MMA) come from shared memory, and
// the freed shared memory double-buffers them.
//   TMEM columns: S[2] 0-127 | O[2] 128-255 | P[2] x (hi | lo) 256-511  -- P is double-buffered so that the
//                 softmax of tile j+1 never waits for P.V of tile j (the measured ~430 us of pure hand-off latency)
//   smem bytes  : K_hi[2] | K_lo[2] | V^T_hi[2] | V^T_lo[2] | V_raw[2] | Q_hi | Q_lo
//                 K tiles land directly in their K_hi stage (split in place); every TMA load is issued a full
//                 tile ahead of its consumer, so the load latency is off the per-tile critical path
// Roles: warp 0 TMA, warp 1 MMA issue + TMEM alloc, warps 2-5 transform (Q -> TMEM, K split, V transpose+split),
//        warps 6-13 softmax: TWO threads per query row (32 keys / 32 output dims each; they only exchange the
//        row max through smem), S from TMEM, P hi/lo back to TMEM, O accumulated in registers.
#include "omt_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>

namespace omt {
namespace atc3 {
using namespace omt::ptx;

constexpr int QT = 128, KT = 64, D = 64;
constexpr int Q_BYTES = QT * D * 4;     // 32 KiB
constexpr int K_BYTES = KT * D * 4;     // 16 KiB
constexpr int OFF_KH = 0, OFF_KL = 2 * K_BYTES, OFF_VH = 4 * K_BYTES, OFF_VL = 6 * K_BYTES;
constexpr int OFF_VR = 8 * K_BYTES;                                  // V_raw[2]: TMA landing, two tiles deep
constexpr int OFF_QH = 10 * K_BYTES, OFF_QL = OFF_QH + Q_BYTES;      // Q lands in OFF_QH, split in place
constexpr int OFF_CTRL = OFF_QL + Q_BYTES;                           // barriers, TMEM pointer, row-max exchange
constexpr int SMEM = OFF_CTRL + 2048;                                // 226 KiB: no static shared memory, no slack
constexpr int TM_S = 0, TM_O = 128, TM_P = 256;                      // S[2] | O[2] | P[2] x (hi 64 | lo 64)
constexpr int THREADS = 448;   // TMA, MMA, 4 transform warps, 8 softmax warps (two per TMEM lane quarter)
// tf32 x tf32 -> f32, M=128, N=64; bit 16 = B is MN-major (used for V)
constexpr uint32_t IDESC_KK = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Args {
  float* o; int ldo;
  uint16_t* o_hi; uint16_t* o_lo;   // optional operand planes instead of o
  int N;
  float scale_log2;
};

__global__ void __launch_bounds__(THREADS, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Args a) {
  // All shared memory is dynamic (the kernel needs 226 of the 227 KiB): with no static allocation the dynamic
  // window starts 1024-byte aligned, which the SW128 operand tiles require -- checked, not assumed.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_CTRL);
  uint64_t& q_full = bars[0]; uint64_t& q_ready = bars[1];
  uint64_t* k_full = bars + 2; uint64_t* v_full = bars + 4; uint64_t* vr_free = bars + 26;
  uint64_t* k_ready = bars + 6;  uint64_t* k_empty = bars + 8;  uint64_t* v_ready = bars + 10; uint64_t* v_empty = bars + 12;
  uint64_t* s_full = bars + 14;  uint64_t* s_empty = bars + 16; uint64_t* o_full = bars + 18;  uint64_t* o_empty = bars + 20;
  uint64_t* p_full = bars + 22;
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(bars + 28);
  float (*xch)[QT] = reinterpret_cast<float (*)[QT]>(smem + OFF_CTRL + 256);   // [2 key halves][row] row-max exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int ntiles = a.N / KT;
  const int row_q0 = seq * a.N + qt * QT;
  const int row_k0 = seq * a.N;
  const int col0 = head * D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    mbar_init(&q_full, 1); mbar_init(&q_ready, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&vr_free[i], 4);
      mbar_init(&k_ready[i], 4); mbar_init(&k_empty[i], 1);
      mbar_init(&v_ready[i], 4); mbar_init(&v_empty[i], 1);
      mbar_init(&p_full[i], 8);      // per P stage: softmax warps may run one tile ahead of the tensor pipe
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();      // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(&q_full, Q_BYTES);
      tma_load_2d(&tmQ, &q_full, smem + OFF_QH, col0, row_q0);
      tma_load_2d(&tmQ, &q_full, smem + OFF_QH + Q_BYTES / 2, col0 + 32, row_q0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        const uint32_t ph2 = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph2 ^ 1);        // S of tile j-2 retired: its K stage is free -> K_j lands one tile early
        mbar_expect_tx(&k_full[st], K_BYTES);
        tma_load_2d(&tmK, &k_full[st], smem + OFF_KH + st * K_BYTES, col0, row_k0 + j * KT);
        tma_load_2d(&tmK, &k_full[st], smem + OFF_KH + st * K_BYTES + K_BYTES / 2, col0 + 32, row_k0 + j * KT);
        mbar_wait(&vr_free[st], ph2 ^ 1);
        mbar_expect_tx(&v_full[st], K_BYTES);
        tma_load_2d(&tmV, &v_full[st], smem + OFF_VR + st * K_BYTES, col0, row_k0 + j * KT);
        tma_load_2d(&tmV, &v_full[st], smem + OFF_VR + st * K_BYTES + K_BYTES / 2, col0 + 32, row_k0 + j * KT);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged; one elected lane issues) =================
    {
      const uint32_t sb = smem_u32(smem);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t ph2 = (j >> 1) & 1;
        mbar_wait(&k_ready[st], ph2);
        mbar_wait(&s_empty[st], ph2 ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + TM_S + st * 64;
          const uint64_t kh0 = desc_kmajor(sb + OFF_KH + st * K_BYTES), kl0 = desc_kmajor(sb + OFF_KL + st * K_BYTES);
          const uint64_t qh0 = desc_kmajor(sb + OFF_QH), ql0 = desc_kmajor(sb + OFF_QL);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {          // 8 tf32 of the head dim per MMA
            const uint64_t adv = (uint64_t)(((kk >> 2) * (K_BYTES / 2) + (kk & 3) * 32) >> 4);
            const uint64_t qadv = (uint64_t)(((kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32) >> 4);
            mma_tf32(d, ql0 + qadv, kh0 + adv, IDESC_KK, kk != 0);
            mma_tf32(d, qh0 + qadv, kl0 + adv, IDESC_KK, 1);
            mma_tf32(d, qh0 + qadv, kh0 + adv, IDESC_KK, 1);
          }
          tc_commit(&s_full[st]);
          tc_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(&q_ready, 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        const int st = j & 1;
        const uint32_t ph2 = (j >> 1) & 1;
        mbar_wait(&p_full[st], ph2);
        mbar_wait(&v_ready[st], ph2);
        mbar_wait(&o_empty[st], ph2 ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + TM_O + st * 64;
          const uint64_t vh0 = desc_kmajor(sb + OFF_VH + st * K_BYTES), vl0 = desc_kmajor(sb + OFF_VL + st * K_BYTES);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {          // 8 keys per MMA
            const uint64_t adv = (uint64_t)(((kk >> 2) * (K_BYTES / 2) + (kk & 3) * 32) >> 4);
            const uint32_t pbase = tmem_base + TM_P + st * 128;
            mma_tf32_ts(d, pbase + 64 + kk * 8, vh0 + adv, IDESC_KK, kk != 0);
            mma_tf32_ts(d, pbase + kk * 8, vl0 + adv, IDESC_KK, 1);
            mma_tf32_ts(d, pbase + kk * 8, vh0 + adv, IDESC_KK, 1);
          }
          tc_commit(&o_full[st]);
          tc_commit(&v_empty[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ================= transform =================
    const int t = threadIdx.x - 64;
    // ---- Q: tf32 hi (in place) / lo in shared memory
    mbar_wait(&q_full, 0);
    {
      float4* h = reinterpret_cast<float4*>(smem + OFF_QH);
      float4* l = reinterpret_cast<float4*>(smem + OFF_QL);
#pragma unroll
      for (int i = 0; i < Q_BYTES / 16 / 128; ++i) {
        const int idx = t + i * 128;
        const float4 v = h[idx];
        float4 hi, lo;
        hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        h[idx] = hi;
        l[idx] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_ready);
    }
    auto split_k = [&](int jj) {
      const int st = jj & 1;
      mbar_wait(&k_full[st], (jj >> 1) & 1);
      float4* h = reinterpret_cast<float4*>(smem + OFF_KH + st * K_BYTES);
      const float4* src = h;                     // split in place
      float4* l = reinterpret_cast<float4*>(smem + OFF_KL + st * K_BYTES);
#pragma unroll
      for (int i = 0; i < K_BYTES / 16 / 128; ++i) {
        const int idx = t + i * 128;
        const float4 v = src[idx];
        float4 hi, lo;
        hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        h[idx] = hi;
        l[idx] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_ready[st]);
    };
    split_k(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) split_k(j + 1);
      const int st = j & 1;
      mbar_wait(&v_full[st], (j >> 1) & 1);
      mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + t;
        const int key = idx & 63, d4 = idx >> 6;
        const float4 v = *reinterpret_cast<const float4*>(smem + OFF_VR + st * K_BYTES + (d4 >> 3) * (K_BYTES / 2) + key * 128 +
                                                          (((d4 & 7) ^ (key & 7)) << 4));
        const float e[4] = {v.x, v.y, v.z, v.w};
        const int c = key >> 5, kk = key & 31;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = d4 * 4 + i;
          const int off = st * K_BYTES + c * (K_BYTES / 2) + d * 128 + ((((kk >> 2) ^ (d & 7)) << 4) | ((kk & 3) << 2));
          const float hi = tf32_rn(e[i]);
          *reinterpret_cast<float*>(smem + OFF_VH + off) = hi;
          *reinterpret_cast<float*>(smem + OFF_VL + off) = e[i] - hi;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&v_ready[st]); mbar_arrive(&vr_free[st]); }
    }
  } else {
    // ================= softmax + output accumulation =================
    // thread (q, lane, half): query row r = 32q + lane, keys [32*half, +32) of every tile and output dims
    // [32*half, +32).  The pair of a row sits in warps w and w+4 (same TMEM lane quarter).
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int bar_id = 2 + q;                        // named barrier of this warp pair (64 threads)
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    for (int j = 0; j < ntiles; ++j) {
      float s[32];
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld32(tmem_base + TM_S + lane_addr + (j & 1) * 64 + half * 32, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, s[i]);
      xch[half][r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, xch[half ^ 1][r]);
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");   // partner has read before the slot is reused
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * a.scale_log2);
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { s[i] = exp2f((s[i] - m_new) * a.scale_log2); psum += s[i]; }
      l_run = l_run * alpha + psum;                  // partial row sum over this thread's keys
      m_run = m_new;
      // publish P_j first (its TMEM stage was released when tile j-2 was folded), then fold O_{j-1}:
      // the tensor pipe starts P_j.V_j while this thread is still accumulating the previous tile
      {
        float hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { hi[i] = tf32_rn(s[i]); lo[i] = s[i] - hi[i]; }
        const uint32_t pbase = tmem_base + lane_addr + TM_P + (j & 1) * 128 + half * 32;
        tmem_st32(pbase, hi);
        tmem_st32(pbase + 64, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      if (j > 0) {
        const int jp = j - 1;
        mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
        tc_fence_after();
        float oj[32];
        tmem_ld32(tmem_base + TM_O + lane_addr + (jp & 1) * 64 + half * 32, oj);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha_prev, oj[i]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[jp & 1]);
      }
      alpha_prev = alpha;
    }
    {
      const int jp = ntiles - 1;
      mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
      tc_fence_after();
      float oj[32];
      tmem_ld32(tmem_base + TM_O + lane_addr + (jp & 1) * 64 + half * 32, oj);
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha_prev, oj[i]);
      tc_fence_before();
    }
    // total row sum = the two partial sums (same running max on both sides)
    xch[half][r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float inv = 1.0f / (l_run + xch[half ^ 1][r]);
    const size_t ooff = (size_t)(row_q0 + r) * a.ldo + col0 + half * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 ov = make_float4(o_acc[i] * inv, o_acc[i + 1] * inv, o_acc[i + 2] * inv, o_acc[i + 3] * inv);
      if (a.o_hi != nullptr) store_split4(a.o_hi, a.o_lo, ooff + i, ov);
      else *reinterpret_cast<float4*>(a.o + ooff + i) = ov;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode2d(CUtensorMap* m, const float* base, int cols, long long rows, int ld, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return OMT_E_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return OMT_E_CUDA; }
  return OMT_OK;
}

}  // namespace atc3

int launch_attn_tc3(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, uint16_t* o_hi,
                    uint16_t* o_lo, int ldo, int n_seq, int N, int heads, float scale, cudaStream_t st) {
  using namespace atc3;
  CUtensorMap tmQ, tmK, tmV;
  const long long rows = (long long)n_seq * N;
  int rc = encode2d(&tmQ, q, heads * D, rows, ldq, QT);
  if (rc) return rc;
  rc = encode2d(&tmK, k, heads * D, rows, ldk, KT);
  if (rc) return rc;
  rc = encode2d(&tmV, v, heads * D, rows, ldv, KT);
  if (rc) return rc;
  static bool attr[64];      // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    OMT_CUDA(cudaFuncSetAttribute(attn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr[dev] = true;
  }
  Args a{o, ldo, o_hi, o_lo, N, scale * 1.4426950408889634f};
  dim3 grid(N / QT, heads, n_seq);
  OMT_CUDA(launch_k(attn_tc3_kernel, grid, dim3(THREADS), SMEM, st, tmQ, tmK, tmV, a));
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

}  // namespace omt
