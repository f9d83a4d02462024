// tcgen05 / TMA / TMEM GEMM for every nn.Linear on the path (sm_100a):
//   C[M,N] = A[M,K] . W[N,K]^T (+bias)(+residual) | GEGLU epilogue, fp32 in / fp32 out.
//
// Math modes
//   3xTF32 : A = A_hi + A_lo, W = W_hi + W_lo (tf32 round-to-nearest split; W split offline,
//            A split in shared memory by the transform warps); D += A_lo.W_hi + A_hi.W_lo + A_hi.W_hi
//            with fp32 accumulation in TMEM.  fp32-grade accuracy (code indices bit-exact).
//   TF32   : single pass on the raw operands (throughput mode).
//
// Structure (one 128 x BN output tile per CTA, BLOCK_K = 32 fp32 = one 128-byte swizzle row):
//   warp 0      : TMA producer   (A: two 64-row boxes through a 3-D map that also encodes the
//                                 first-frame / rest-frames row map; W_hi, W_lo: one box each)
//   warp 1      : TMEM alloc + single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 2..5  : hi/lo transform of the A stage in smem (generic proxy -> fence.proxy.async),
//                 then the epilogue: tcgen05.ld 32x32b -> smem transpose -> coalesced 16 B stores.
#include "omt_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace omt {

namespace tc {
using namespace omt::ptx;

constexpr int BM = 128;
constexpr int BK = 32;                      // fp32 elements per k-block (128 bytes)
constexpr int A_BYTES = BM * BK * 4;        // 16 KiB
constexpr int SMEM_BUDGET = 216 * 1024;

template <int BN, bool SPLIT>
struct Cfg {
  static constexpr int W_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = SPLIT ? 2 * A_BYTES + 2 * W_BYTES : A_BYTES + W_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TX_BYTES = A_BYTES + (SPLIT ? 2 : 1) * W_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024;
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmWlo, const GemmArgs g, const int epilogue) {
  using C = Cfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t ready_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int num_kb = g.K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    if (SPLIT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWlo)) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * C::STAGE_BYTES; };

  if (warp == 0) {
    if (lane == 0) {
      // logical rows m0 .. m0+127 as two 64-row boxes; (row in segment, segment) coordinates
      int c1[2], c2[2];
      for (int hf = 0; hf < 2; ++hf) {
        const int r = m0 + hf * 64;
        if (g.a_seg > 0) { c1[hf] = r % g.a_seg; c2[hf] = r / g.a_seg; }
        else { c1[hf] = r; c2[hf] = 0; }
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], C::TX_BYTES);
        uint8_t* sp = stage_ptr(s);
        tma_load_3d(&tmA, &full_bar[s], sp, kb * BK, c1[0], c2[0]);
        tma_load_3d(&tmA, &full_bar[s], sp + A_BYTES / 2, kb * BK, c1[1], c2[1]);
        if (SPLIT) {
          tma_load_2d(&tmW, &full_bar[s], sp + 2 * A_BYTES, kb * BK, n0);
          tma_load_2d(&tmWlo, &full_bar[s], sp + 2 * A_BYTES + C::W_BYTES, kb * BK, n0);
        } else {
          tma_load_2d(&tmW, &full_bar[s], sp + A_BYTES, kb * BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (SPLIT) mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_ptr(s));
        if (SPLIT) {
          const uint64_t d_ahi = desc_kmajor(sa), d_alo = desc_kmajor(sa + A_BYTES);
          const uint64_t d_whi = desc_kmajor(sa + 2 * A_BYTES), d_wlo = desc_kmajor(sa + 2 * A_BYTES + C::W_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);   // 32 bytes per k-step inside the swizzle row
            mma_tf32(tmem_base, d_alo + adv, d_whi + adv, C::IDESC, (kb | k) != 0);
            mma_tf32(tmem_base, d_ahi + adv, d_wlo + adv, C::IDESC, 1);
            mma_tf32(tmem_base, d_ahi + adv, d_whi + adv, C::IDESC, 1);
          }
        } else {
          const uint64_t d_a = desc_kmajor(sa), d_w = desc_kmajor(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            mma_tf32(tmem_base, d_a + adv, d_w + adv, C::IDESC, (kb | k) != 0);
          }
        }
        tc_commit(&empty_bar[s]);     // frees the smem stage once these MMAs retire
      }
      tc_commit(&tmem_full_bar);      // accumulator complete
    }
    __syncwarp();
  } else {
    const int t = threadIdx.x - 64;   // 0..127
    if (SPLIT) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        float4* a = reinterpret_cast<float4*>(stage_ptr(s));
        float4* alo = reinterpret_cast<float4*>(stage_ptr(s) + A_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + i * 128;       // elementwise: the 128B swizzle is position-agnostic
          const float4 v = a[idx];
          float4 hi, lo;
          hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
          lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
          a[idx] = hi;
          alo[idx] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to tcgen05
        mbar_arrive(&ready_bar[s]);
      }
    }
    // ---------------- epilogue ----------------
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    float* stg = reinterpret_cast<float*>(smem) + q * (32 * 33);
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= g.N) break;
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + (lane >> 3), col = (lane & 7) * 4;
        const int m = m0 + q * 32 + rl, n = n0 + c * 32 + col;
        if (m < g.M && n < g.N) {
          const float* sp = stg + rl * 33 + col;
          float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
          if (g.bias != nullptr) {
            const float4 bb = *reinterpret_cast<const float4*>(g.bias + n);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          const long long prow = map_row(m, g.c_seg, g.c_seg_stride, g.c_seg_off);
          if (epilogue == OMT_EPI_GEGLU) {
            float2 o;
            o.x = gelu_erf(v.y) * v.x;
            o.y = gelu_erf(v.w) * v.z;
            *reinterpret_cast<float2*>(g.C + prow * g.ldc + (n >> 1)) = o;
          } else {
            if (g.residual != nullptr) {
              const float4 rr = *reinterpret_cast<const float4*>(g.residual + prow * g.ldr + n);
              v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            *reinterpret_cast<float4*>(g.C + prow * g.ldc + n) = v;
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return OMT_E_CUDA; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return OMT_E_CUDA;
  }
  return OMT_OK;
}

template <int BN, bool SPLIT>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmWlo, const GemmArgs& g,
                  int epilogue, cudaStream_t st) {
  using C = Cfg<BN, SPLIT>;
  static bool attr = false;
  if (!attr) {
    OMT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  gemm_tc_kernel<BN, SPLIT><<<grid, 192, C::SMEM, st>>>(tmA, tmW, tmWlo, g, epilogue);
  OMT_LAUNCH_CHECK();
  return OMT_OK;
}

}  // namespace tc

static int g_tc_bn = 128;   // tile-N selector of the v1 kernel (set through omt_set_option for tuning)
static int g_tc_kernel = 2; // 2 = persistent 2-CTA kernel (gemm_tc2.cu) for 3xTF32, 1 = one-tile-per-CTA kernel

int launch_gemm_tc2(const GemmArgs& g, const float* W_lo, int epilogue, cudaStream_t st, const float* A2, int n_split);

int tc_fuses_qkprep(int math) { return math == OMT_MATH_3XTF32 && g_tc_kernel == 2; }

int launch_gemm_tc(const GemmArgs& g, const float* W_lo, int epilogue, int math, cudaStream_t st, const float* A2, int n_split) {
  using namespace tc;
  OMT_REQUIRE(g.K % BK == 0, "omt_linear(tcgen05): K=%d must be a multiple of 32", g.K);
  OMT_REQUIRE(g.lda % 4 == 0, "omt_linear(tcgen05): lda %% 4");
  if (g.a_seg > 0) {
    OMT_REQUIRE(g.a_seg % 64 == 0 && g.M % g.a_seg == 0, "omt_linear(tcgen05): A row-map segment %d must be a multiple of 64 dividing M=%d", g.a_seg, g.M);
  }
  const bool split = (math == OMT_MATH_3XTF32);
  if (split && g_tc_kernel == 2) return launch_gemm_tc2(g, W_lo, epilogue, st, A2, n_split);
  OMT_REQUIRE(A2 == nullptr, "omt_linear2: the dual-A form needs the v2 tcgen05 kernel or the fp32 path");
  const int n_pad = (g.N + 127) / 128 * 128;
  CUtensorMap tmA, tmW, tmWlo;
  {
    const int seg = g.a_seg > 0 ? g.a_seg : g.M;
    const int nseg = g.a_seg > 0 ? g.M / g.a_seg : 1;
    const long long sstride = g.a_seg > 0 ? g.a_seg_stride : g.M;
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)seg, (cuuint64_t)nseg};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * 4, (cuuint64_t)sstride * g.lda * 4};
    cuuint32_t box[3] = {BK, 64, 1};
    const float* base = g.A + (size_t)(g.a_seg > 0 ? g.a_seg_off : 0) * g.lda;
    int rc = encode_map(&tmA, base, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int bn = (g_tc_bn == 256 && n_pad % 256 == 0) ? 256 : 128;
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)g.K * 4};
    cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    int rc = encode_map(&tmW, g.W, 2, dims, strides, box);
    if (rc) return rc;
    rc = encode_map(&tmWlo, split ? W_lo : g.W, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (bn == 256) {
    return split ? launch<256, true>(tmA, tmW, tmWlo, g, epilogue, st) : launch<256, false>(tmA, tmW, tmWlo, g, epilogue, st);
  }
  return split ? launch<128, true>(tmA, tmW, tmWlo, g, epilogue, st) : launch<128, false>(tmA, tmW, tmWlo, g, epilogue, st);
}

}  // namespace omt

namespace omt { extern int g_attn_kernel; extern int g_attn_debug; extern int g_peg_kernel; extern int g_tc2_arrive_cta; }

extern "C" int omt_set_option(const char* name, int value) {
  if (name == nullptr) return OMT_E_ARG;
  if (strcmp(name, "pdl") == 0) { omt::g_pdl = value ? 1 : 0; return OMT_OK; }
  if (strcmp(name, "tc_arrive_cta") == 0) { omt::g_tc2_arrive_cta = value ? 1 : 0; return OMT_OK; }
  if (strcmp(name, "peg_kernel") == 0) {
    if (value != 3 && value != 4) { omt::set_error("peg_kernel must be 3 or 4"); return OMT_E_ARG; }
    omt::g_peg_kernel = value;
    return OMT_OK;
  }
  if (strcmp(name, "attn_debug") == 0) { omt::g_attn_debug = value; return OMT_OK; }
  if (strcmp(name, "attn_kernel") == 0) {
    if (value < 1 || value > 3) { omt::set_error("attn_kernel must be 1, 2 or 3"); return OMT_E_ARG; }
    omt::g_attn_kernel = value;
    return OMT_OK;
  }
  if (strcmp(name, "tc_block_n") == 0) {
    if (value != 128 && value != 256) { omt::set_error("tc_block_n must be 128 or 256"); return OMT_E_ARG; }
    omt::g_tc_bn = value;
    return OMT_OK;
  }
  if (strcmp(name, "tc_kernel") == 0) {
    if (value != 1 && value != 2) { omt::set_error("tc_kernel must be 1 or 2"); return OMT_E_ARG; }
    omt::g_tc_kernel = value;
    return OMT_OK;
  }
  omt::set_error("unknown option %s", name);
  return OMT_E_ARG;
}
