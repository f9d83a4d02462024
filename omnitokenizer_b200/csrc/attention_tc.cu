// Spatial (full, non-causal) attention core on the tcgen05 tensor cores with 3xTF32 compensation:
//   O = softmax(scale * Q K^T) V   per (sequence, head), head dim 64, N % 128 == 0.
// (reference: F.scaled_dot_product_attention at modules/attention.py:451; q, k already carry
//  rope + l2norm + per-dim scale from omt_qk_prep.)
//
// One CTA = 128 queries of one (sequence, head); keys/values are streamed in tiles of 64.
//   warp 0      TMA producer: Q once, then K / V tiles straight out of the [M, 1536] QKV buffer
//               (2-D tensor maps, one 32-column SW128 box per half of the head dim)
//   warp 1      TMEM alloc + single-thread MMA issue:
//                 S_j  = Q K_j^T        A = Q (K-major), B = K_j (K-major)          -> TMEM S[j&1]
//                 O_j  = P_j V_j        A = P_j (K-major, written by the softmax warps),
//                                       B = V_j^T (K-major; V is transposed + split by the transform warps --
//                                       an MN-major descriptor on V as loaded returned zeros on hardware) -> TMEM O[j&1]
//               every product is 3 MMAs: lo.hi + hi.lo + hi.hi (tf32 operands, fp32 accumulate)
//   warps 2-5   transform: split Q, K_j, V_j into tf32 hi (in place) / lo in shared memory
//   warps 6-9   softmax: one thread per query row; S_j from TMEM, online max / sum in fp32 (exp2),
//               P_j -> shared memory as tf32 hi / lo, O accumulated in REGISTERS:
//               O = O * alpha_j + O_j  (O_j read back from TMEM), so no TMEM rescale pass exists.
#include "omt_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>

namespace omt {
namespace atc {
using namespace omt::ptx;

constexpr int QT = 128;                 // queries per CTA
constexpr int KT = 64;                  // keys per tile
constexpr int D = 64;
constexpr int Q_BYTES = QT * D * 4;     // 32 KiB (two 16 KiB column halves)
constexpr int K_BYTES = KT * D * 4;     // 16 KiB
constexpr int P_BYTES = QT * KT * 4;    // 32 KiB (two 16 KiB key halves)
// smem map (bytes): Q_hi | Q_lo | K_hi | K_lo | V^T_hi | V^T_lo | P_hi | P_lo | V_raw | K_raw
constexpr int OFF_QH = 0, OFF_QL = Q_BYTES, OFF_KH = 2 * Q_BYTES, OFF_KL = OFF_KH + K_BYTES;
constexpr int OFF_VH = OFF_KL + K_BYTES, OFF_VL = OFF_VH + K_BYTES, OFF_PH = OFF_VL + K_BYTES, OFF_PL = OFF_PH + P_BYTES;
constexpr int OFF_VR = OFF_PL + P_BYTES, OFF_KR = OFF_VR + K_BYTES;   // TMA landing buffers (raw fp32)
constexpr int SMEM = OFF_KR + K_BYTES + 1024;        // 224 KiB + alignment slack
constexpr int THREADS = 320;
// tf32 x tf32 -> f32, M=128, N=64, both operands K-major
constexpr uint32_t IDESC_KK = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Args {
  float* o; int ldo;
  int N;            // tokens per sequence
  float scale_log2; // scale * log2(e)
  int dbg;          // debug knob (omt_set_option("attn_debug", n)); 0 in production
};

__global__ void __launch_bounds__(THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t q_full, q_ready, k_full, k_ready, k_empty, v_full, v_ready, v_empty, vr_free, kr_free, p_full;
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], o_full[2], o_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int ntiles = a.N / KT;
  const int row_q0 = seq * a.N + qt * QT;      // first query row in the [M, ld] buffer
  const int row_k0 = seq * a.N;
  const int col0 = head * D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    mbar_init(&q_full, 1); mbar_init(&q_ready, 4);
    mbar_init(&k_full, 1); mbar_init(&k_ready, 4); mbar_init(&k_empty, 1);
    mbar_init(&v_full, 1); mbar_init(&v_ready, 4); mbar_init(&v_empty, 1); mbar_init(&vr_free, 4); mbar_init(&kr_free, 4);
    mbar_init(&p_full, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tm_s = tmem_base;             // S[2] : 2 x 64 columns
  const uint32_t tm_o = tmem_base + 128;       // O_j[2] : 2 x 64 columns

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(&q_full, Q_BYTES);
      tma_load_2d(&tmQ, &q_full, smem + OFF_QH, col0, row_q0);
      tma_load_2d(&tmQ, &q_full, smem + OFF_QH + Q_BYTES / 2, col0 + 32, row_q0);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(&kr_free, ph ^ 1);             // K_j lands in the raw buffer while S_{j-1} is still running
        mbar_expect_tx(&k_full, K_BYTES);
        tma_load_2d(&tmK, &k_full, smem + OFF_KR, col0, row_k0 + j * KT);
        tma_load_2d(&tmK, &k_full, smem + OFF_KR + K_BYTES / 2, col0 + 32, row_k0 + j * KT);
        mbar_wait(&vr_free, ph ^ 1);
        mbar_expect_tx(&v_full, K_BYTES);
        tma_load_2d(&tmV, &v_full, smem + OFF_VR, col0, row_k0 + j * KT);
        tma_load_2d(&tmV, &v_full, smem + OFF_VR + K_BYTES / 2, col0 + 32, row_k0 + j * KT);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t sb = smem_u32(smem);
      auto issue_s = [&](int j) {
        mbar_wait(&k_ready, j & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tm_s + (j & 1) * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {             // halves of the head dim (32 columns = one swizzle row)
          const uint64_t qh = desc_kmajor(sb + OFF_QH + c * (Q_BYTES / 2)), ql = desc_kmajor(sb + OFF_QL + c * (Q_BYTES / 2));
          const uint64_t kh = desc_kmajor(sb + OFF_KH + c * (K_BYTES / 2)), kl = desc_kmajor(sb + OFF_KL + c * (K_BYTES / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            mma_tf32(d, ql + adv, kh + adv, IDESC_KK, (c | k) != 0);
            mma_tf32(d, qh + adv, kl + adv, IDESC_KK, 1);
            mma_tf32(d, qh + adv, kh + adv, IDESC_KK, 1);
          }
        }
        tc_commit(&s_full[j & 1]);
        tc_commit(&k_empty);
      };
      mbar_wait(&q_ready, 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);       // S of the next tile overlaps the softmax of this one
        mbar_wait(&p_full, j & 1);
        mbar_wait(&v_ready, j & 1);
        mbar_wait(&o_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tm_o + (j & 1) * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {             // halves of the key tile (32 keys = one swizzle row of P)
          const uint64_t ph_ = desc_kmajor(sb + OFF_PH + c * (P_BYTES / 2)), pl_ = desc_kmajor(sb + OFF_PL + c * (P_BYTES / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            // V^T chunk c: [64 d rows][32 keys] K-major (transposed + split by the transform warps)
            const uint64_t vh = desc_kmajor(sb + OFF_VH + c * (K_BYTES / 2)) + adv, vl = desc_kmajor(sb + OFF_VL + c * (K_BYTES / 2)) + adv;
            mma_tf32(d, pl_ + adv, vh, IDESC_KK, (c | k) != 0);
            mma_tf32(d, ph_ + adv, vl, IDESC_KK, 1);
            mma_tf32(d, ph_ + adv, vh, IDESC_KK, 1);
          }
        }
        tc_commit(&o_full[j & 1]);
        tc_commit(&v_empty);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================= transform: tf32 hi (in place) / lo =================
    const int t = threadIdx.x - 64;
    auto split = [&](int off_hi, int off_lo, int bytes) {
      float4* h = reinterpret_cast<float4*>(smem + off_hi);
      float4* l = reinterpret_cast<float4*>(smem + off_lo);
      for (int idx = t; idx < bytes / 16; idx += 128) {
        const float4 v = h[idx];
        float4 hi, lo;
        hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        h[idx] = hi;
        l[idx] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    split(OFF_QH, OFF_QL, Q_BYTES);
    if (lane == 0) mbar_arrive(&q_ready);
    // K_{j+1} is split BEFORE V_j is transposed: its operand buffers free up as soon as S_j retires, whereas
    // V^T must wait for P.V of tile j-1 -- doing them in tile order serialised S_{j+1} behind that wait.
    auto split_k = [&](int jj) {
      mbar_wait(&k_full, jj & 1);
      mbar_wait(&k_empty, (jj & 1) ^ 1);          // S_{j-1} no longer reads the K hi / lo operand buffers
      {
        const float4* src = reinterpret_cast<const float4*>(smem + OFF_KR);
        float4* h = reinterpret_cast<float4*>(smem + OFF_KH);
        float4* l = reinterpret_cast<float4*>(smem + OFF_KL);
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
      }
      if (lane == 0) { mbar_arrive(&k_ready); mbar_arrive(&kr_free); }
    };
    split_k(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) split_k(j + 1);
      // V_j: [keys][d] as landed (SW128 per 32-wide d half) -> V^T [d][keys] K-major hi / lo
      mbar_wait(&v_full, j & 1);
      mbar_wait(&v_empty, (j & 1) ^ 1);          // P.V of tile j-1 no longer reads the V^T buffers
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + t;
        const int key = idx & 63, d4 = idx >> 6;            // a warp = 32 consecutive keys, one d quad
        const float4 v = *reinterpret_cast<const float4*>(smem + OFF_VR + (d4 >> 3) * (K_BYTES / 2) + key * 128 +
                                                          (((d4 & 7) ^ (key & 7)) << 4));
        const float e[4] = {v.x, v.y, v.z, v.w};
        const int c = key >> 5, kk = key & 31;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = d4 * 4 + i;
          const int off = c * (K_BYTES / 2) + d * 128 + ((((kk >> 2) ^ (d & 7)) << 4) | ((kk & 3) << 2));
          const float hi = tf32_rn(e[i]);
          *reinterpret_cast<float*>(smem + OFF_VH + off) = hi;
          *reinterpret_cast<float*>(smem + OFF_VL + off) = e[i] - hi;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&v_ready); mbar_arrive(&vr_free); }
    }
  } else {
    // ================= softmax + output accumulation =================
    const int q = warp & 3;                        // TMEM lane quarter
    const int r = q * 32 + lane;                   // query row inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float o_acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    for (int j = 0; j < ntiles; ++j) {
      float s[KT];
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld32(tm_s + lane_addr + (j & 1) * 64, s);
      tmem_ld32(tm_s + lane_addr + (j & 1) * 64 + 32, s + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < KT; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * a.scale_log2);
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < KT; ++i) { s[i] = exp2f((s[i] - m_new) * a.scale_log2); psum += s[i]; }
      l_run = l_run * alpha + psum;
      m_run = m_new;
      // fold the previous tile's P.V (also proves the P buffer is free again)
      if (j > 0) {
        const int jp = j - 1;
        mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
        tc_fence_after();
        float oj[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld32(tm_o + lane_addr + (jp & 1) * 64 + hh * 32, oj);
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[hh * 32 + i] = fmaf(o_acc[hh * 32 + i], alpha_prev, oj[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[jp & 1]);
      }
      alpha_prev = alpha;
      // P_j -> smem (K-major SW128: row r, 16-byte unit u of key-half c at (u ^ (r & 7)))
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint8_t* ph_ = smem + OFF_PH + c * (P_BYTES / 2) + r * 128;
        uint8_t* pl_ = smem + OFF_PL + c * (P_BYTES / 2) + r * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 hi, lo;
          const float* sv = s + c * 32 + u * 4;
          hi.x = tf32_rn(sv[0]); hi.y = tf32_rn(sv[1]); hi.z = tf32_rn(sv[2]); hi.w = tf32_rn(sv[3]);
          lo.x = sv[0] - hi.x; lo.y = sv[1] - hi.y; lo.z = sv[2] - hi.z; lo.w = sv[3] - hi.w;
          const int su = (u ^ (r & 7)) * 16;
          *reinterpret_cast<float4*>(ph_ + su) = hi;
          *reinterpret_cast<float4*>(pl_ + su) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
    }
    // last tile's P.V
    {
      const int jp = ntiles - 1;
      mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
      tc_fence_after();
      float oj[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tmem_ld32(tm_o + lane_addr + (jp & 1) * 64 + hh * 32, oj);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[hh * 32 + i] = fmaf(o_acc[hh * 32 + i], alpha_prev, oj[i]);
      }
      tc_fence_before();
    }
    float inv = 1.0f / l_run;
    if (a.dbg == 3) { inv = 1.0f; o_acc[0] = l_run; o_acc[1] = m_run; o_acc[2] = alpha_prev; }
    float* op = a.o + (size_t)(row_q0 + r) * a.ldo + col0;
#pragma unroll
    for (int i = 0; i < D; i += 4)
      *reinterpret_cast<float4*>(op + i) = make_float4(o_acc[i] * inv, o_acc[i + 1] * inv, o_acc[i + 2] * inv, o_acc[i + 3] * inv);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
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

}  // namespace atc

int g_attn_debug = 0;

int launch_attn_tc(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo,
                   int n_seq, int N, int heads, float scale, cudaStream_t st) {
  using namespace atc;
  CUtensorMap tmQ, tmK, tmV;
  const long long rows = (long long)n_seq * N;
  int rc = encode2d(&tmQ, q, heads * D, rows, ldq, QT);
  if (rc) return rc;
  rc = encode2d(&tmK, k, heads * D, rows, ldk, KT);
  if (rc) return rc;
  rc = encode2d(&tmV, v, heads * D, rows, ldv, KT);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    OMT_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  Args a{o, ldo, N, scale * 1.4426950408889634f, g_attn_debug};
  dim3 grid(N / QT, heads, n_seq);
  attn_tc_kernel<<<grid, THREADS, SMEM, st>>>(tmQ, tmK, tmV, a);
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

}  // namespace omt
