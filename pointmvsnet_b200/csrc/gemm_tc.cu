// tcgen05 (5th-gen tensor core) GEMM for the per-point contractions of EdgeConv / flow_mlp:
//   Y[r, 0:N] = f(X[r, 0:K]) * W[0:N, 0:K]^T        r = points (rows), fp32 in / fp32 out
// (reference: nn.Conv1d 1x1 in networks.py:13-14,22-23,51-52 and nn/conv.py:21-30 followed by
// train-mode BatchNorm + ReLU, which is fused here as the INPUT transform of the next layer).
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      : TMEM allocation, single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 1..4  : producers - coalesced 128-bit global loads of a 128 x 32 fp32 slab of X, fused
//                 BatchNorm+ReLU, split into TF32 "hi" and residual "lo" planes, stored to shared
//                 memory in the UMMA K-major SWIZZLE_128B layout, fence.proxy.async, mbarrier arrive
//   warps 5..8  : epilogue - tcgen05.ld the 128 x N fp32 accumulator from TMEM, stage through
//                 shared memory, coalesced stores, per-column sum / sum-of-squares (BN statistics)
// The weight matrix (hi and lo planes) stays resident in shared memory for the CTA's lifetime.
// Accumulators are double buffered in TMEM so the MMA of tile t+1 overlaps the epilogue of tile t.
//
// Precision: kind::tf32 has a 10-bit mantissa.  NSPLIT = 3 runs the error-compensated product
// A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo ("3xTF32", residuals exact in fp32), which keeps fp32-level
// accuracy (measured in tests/test_gpu_parity.py); NSPLIT = 1 is plain TF32.
#include <algorithm>

#include "common.cuh"

namespace pmvs {

namespace tc {

constexpr int BM = 128;            // rows per tile == UMMA M
constexpr int KC = 32;             // fp32 columns per K chunk == one 128-byte swizzle row
constexpr int NSTAGE = 2;          // A-operand pipeline depth
constexpr int NUM_THREADS = 288;   // 1 MMA warp + 4 producer warps + 4 epilogue warps
constexpr int A_PLANE_BYTES = BM * KC * 4;  // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start
// address >> 4 in bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B
// (one 8-row x 128 B swizzle atom) >> 4 in [32,46), version 1 in [46,48), layout 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits
// 7 and 10), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TF32 split: hi keeps the 10 explicit mantissa bits the tensor core reads, lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = __fsub_rn(x, hi);
}

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int N_OUT, int NSPLIT>
struct Smem {
  static constexpr int NPL = NSPLIT == 3 ? 2 : 1;                       // planes (hi [, lo])
  static constexpr int B_CHUNK_BYTES = N_OUT * KC * 4;                  // one K chunk of one plane
  static constexpr int D_PITCH = N_OUT + 4;                             // floats
  static __host__ __device__ constexpr size_t b_bytes(int K) { return (size_t)NPL * round_up(K, KC) / KC * B_CHUNK_BYTES; }
  static __host__ __device__ constexpr size_t a_bytes() { return (size_t)NSTAGE * NPL * A_PLANE_BYTES; }
  static __host__ __device__ constexpr size_t d_bytes() { return (size_t)BM * D_PITCH * 4; }
  static __host__ __device__ constexpr size_t total(int K) {
    return 1024 /*align slack*/ + b_bytes(K) + a_bytes() + d_bytes() + 4 * 224 * 4 /*BN coeffs*/ + 256 /*barriers*/;
  }
};

template <int N_OUT, int NSPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const GemmArgs a, int tiles_per_group) {
  using S = Smem<N_OUT, NSPLIT>;
  constexpr int NPL = S::NPL;
  constexpr int TMEM_COLS = 2 * N_OUT < 32 ? 32 : 2 * N_OUT;  // two accumulators, power of two >= 32
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  const int K = a.cin;
  const int nch = (K + KC - 1) / KC;
  unsigned char* sB = smem;                                   // [plane][chunk][N_OUT x 128 B swizzled]
  unsigned char* sA = sB + S::b_bytes(K);                     // [stage][plane][128 x 128 B swizzled]
  float* sD = (float*)(sA + S::a_bytes());                    // [128][D_PITCH]
  float* sBN = sD + BM * S::D_PITCH;                          // mean, istd, gamma, beta  x 224
  uint64_t* bars = (uint64_t*)(sBN + 4 * 224);
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);
  // barriers: full[NSTAGE] 0.., empty[NSTAGE], tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NSTAGE);
  const uint32_t bar_tfull = smem_u32(bars + 2 * NSTAGE), bar_tempty = smem_u32(bars + 2 * NSTAGE + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // contiguous tile range of this CTA
  const long long total_tiles = (long long)a.groups * tiles_per_group;
  const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(bar_full + 8 * s, 128);
        mbar_init(bar_empty + 8 * s, 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_tfull + 8 * s, 1);
        mbar_init(bar_tempty + 8 * s, 128);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights: W [N_OUT, K] -> hi / lo planes, K-major SWIZZLE_128B, zero padded to 32 columns
  {
    const int pieces = nch * N_OUT * 8;  // 16-byte pieces per plane
    for (int e = tid; e < pieces; e += NUM_THREADS) {
      const int c = e / (N_OUT * 8);
      const int rem = e - c * (N_OUT * 8);
      const int n = rem >> 3, pc = rem & 7;
      const int k0 = c * KC + pc * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 < K) w = ldg4(a.w + (size_t)n * K + k0);
      float4 hi, lo;
      split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y);
      split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
      const int off = c * S::B_CHUNK_BYTES + (n >> 3) * 1024 + (n & 7) * 128 + ((pc ^ (n & 7)) << 4);
      *reinterpret_cast<float4*>(sB + off) = NSPLIT == 3 ? hi : w;
      if (NSPLIT == 3) *reinterpret_cast<float4*>(sB + nch * S::B_CHUNK_BYTES + off) = lo;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, N_OUT);
      uint32_t stage = 0, phase = 0;
      uint32_t acc_phase[2] = {0, 0};
      int it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int as = it & 1;
        mbar_wait(bar_tempty + 8 * as, acc_phase[as] ^ 1);  // epilogue drained this accumulator
        acc_phase[as] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * N_OUT;
        for (int c = 0; c < nch; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const int ksteps = min(KC, K - c * KC) / 8;
          const uint32_t a_hi = smem_u32(sA + (size_t)stage * NPL * A_PLANE_BYTES);
          const uint32_t a_lo = a_hi + A_PLANE_BYTES;
          const uint32_t b_hi = smem_u32(sB + (size_t)c * S::B_CHUNK_BYTES);
          const uint32_t b_lo = b_hi + nch * S::B_CHUNK_BYTES;
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t first = (c == 0 && j == 0) ? 0u : 1u;
            if (NSPLIT == 3) {
              umma_tf32(d_tmem, make_desc(a_lo + j * 32), make_desc(b_hi + j * 32), idesc, first);
              umma_tf32(d_tmem, make_desc(a_hi + j * 32), make_desc(b_lo + j * 32), idesc, 1u);
              umma_tf32(d_tmem, make_desc(a_hi + j * 32), make_desc(b_hi + j * 32), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc(a_hi + j * 32), make_desc(b_hi + j * 32), idesc, first);
            }
          }
          umma_commit(bar_empty + 8 * stage);  // frees the A stage once these MMAs have read it
          if (c == nch - 1) umma_commit(bar_tfull + 8 * as);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp <= 4) {
    // ============================ producers ============================
    const int pw = warp - 1;           // 0..3: rows pw*32 .. pw*32+31
    const int rsub = lane >> 3, pc = lane & 7;
    const bool in_bn = a.in_stats != nullptr;
    const int ptid = tid - 32;         // 0..127
    uint32_t stage = 0, phase = 0;
    int cur_group = -1;
    for (long long t = t_begin; t < t_end; ++t) {
      const int g = (int)(t / tiles_per_group);
      const int row0 = (int)(t - (long long)g * tiles_per_group) * BM;
      const int rows_valid = min(BM, a.rows_per_group - row0);
      const size_t grow0 = (size_t)g * a.rows_per_group + row0;
      if (in_bn && g != cur_group) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // nobody still reads the old coefficients
        const double* s = a.in_stats + (size_t)g * 2 * K;
        for (int c = ptid; c < K; c += 128) {
          BnCoef k = bn_coef(s[c], s[K + c], a.in_count, a.eps);
          sBN[c] = k.mean;
          sBN[224 + c] = k.invstd;
          sBN[448 + c] = a.in_gamma[c];
          sBN[672 + c] = a.in_beta[c];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      cur_group = g;
      for (int c = 0; c < nch; ++c) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        unsigned char* aHi = sA + (size_t)stage * NPL * A_PLANE_BYTES;
        const int k0 = c * KC + pc * 4;
        const bool kvalid = k0 < K;
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = pw * 32 + i * 4 + rsub;
          x[i] = (kvalid && r < rows_valid) ? ldg4(a.x + (grow0 + r) * a.ldx + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = pw * 32 + i * 4 + rsub;
          float4 v = x[i];
          if (in_bn && kvalid && r < rows_valid) {
            v.x = fmaxf(bn_apply(v.x, sBN[k0 + 0], sBN[224 + k0 + 0], sBN[448 + k0 + 0], sBN[672 + k0 + 0]), 0.f);
            v.y = fmaxf(bn_apply(v.y, sBN[k0 + 1], sBN[224 + k0 + 1], sBN[448 + k0 + 1], sBN[672 + k0 + 1]), 0.f);
            v.z = fmaxf(bn_apply(v.z, sBN[k0 + 2], sBN[224 + k0 + 2], sBN[448 + k0 + 2], sBN[672 + k0 + 2]), 0.f);
            v.w = fmaxf(bn_apply(v.w, sBN[k0 + 3], sBN[224 + k0 + 3], sBN[448 + k0 + 3], sBN[672 + k0 + 3]), 0.f);
          }
          const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((pc ^ (r & 7)) << 4);
          if (NSPLIT == 3) {
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(aHi + off) = hi;
            *reinterpret_cast<float4*>(aHi + A_PLANE_BYTES + off) = lo;
          } else {
            *reinterpret_cast<float4*>(aHi + off) = v;
          }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(bar_full + 8 * stage);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue ============================
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int etid = tid - 160;        // 0..127
    const int row = q * 32 + lane;
    uint32_t acc_phase[2] = {0, 0};
    double cs1 = 0.0, cs2 = 0.0;       // column statistics of (column etid % N_OUT, row part etid / N_OUT)
    int cur_group = -1;
    int it = 0;
    auto flush_stats = [&](int g) {
      if (a.out_stats != nullptr && g >= 0 && etid / N_OUT < (128 / N_OUT > 0 ? 128 / N_OUT : 1)) {
        double* o = a.out_stats + (size_t)g * 2 * a.cout;
        atomicAdd(o + etid % N_OUT, cs1);
        atomicAdd(o + a.cout + etid % N_OUT, cs2);
      }
      cs1 = 0.0;
      cs2 = 0.0;
    };
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int as = it & 1;
      const int g = (int)(t / tiles_per_group);
      const int row0 = (int)(t - (long long)g * tiles_per_group) * BM;
      const int rows_valid = min(BM, a.rows_per_group - row0);
      const size_t grow0 = (size_t)g * a.rows_per_group + row0;
      if (g != cur_group) { flush_stats(cur_group); cur_group = g; }
      mbar_wait(bar_tfull + 8 * as, acc_phase[as]);
      acc_phase[as] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * N_OUT;
      if (N_OUT >= 32) {
#pragma unroll
        for (int cb = 0; cb < N_OUT / 32; ++cb) {
          float v[32];
          tmem_ld32(taddr + cb * 32, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(sD + row * S::D_PITCH + cb * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else {
        float v[16];
        tmem_ld16(taddr, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(sD + row * S::D_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * as);               // accumulator may be overwritten
      asm volatile("bar.sync 2, 128;" ::: "memory");   // staged tile complete
      constexpr int C4 = N_OUT / 4;
      for (int e = etid; e < BM * C4; e += 128) {
        const int r = e / C4, c4 = e - r * C4;
        if (r < rows_valid)
          st4(a.y + (grow0 + r) * a.ldy + c4 * 4, *reinterpret_cast<const float4*>(sD + r * S::D_PITCH + c4 * 4));
      }
      if (a.out_stats != nullptr) {
        // column statistics: thread -> (column, row part); 4 independent accumulators per sum
        constexpr int PARTS = 128 / N_OUT > 0 ? 128 / N_OUT : 1;
        constexpr int RPP = BM / PARTS;
        const int col = etid % N_OUT, part = etid / N_OUT;
        if (part < PARTS) {
          float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
          const int r0 = part * RPP;
          const float* colp = sD + col;
#pragma unroll 4
          for (int r = 0; r < RPP; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int rr = r0 + r + u;
              const float v = rr < rows_valid ? colp[rr * S::D_PITCH] : 0.f;
              s1[u] += v;
              s2[u] = fmaf(v, v, s2[u]);
            }
          }
          cs1 += (double)((s1[0] + s1[1]) + (s1[2] + s1[3]));
          cs2 += (double)((s2[0] + s2[1]) + (s2[2] + s2[3]));
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");   // staging buffer free for the next tile
    }
    flush_stats(cur_group);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

static int g_mode = 3;  // 0: SIMT fp32 GEMM, 1: TF32 tensor cores, 3: 3xTF32 tensor cores (default)

template <int N_OUT, int NSPLIT>
static int launch_one(const GemmArgs& a, cudaStream_t st, const char* name) {
  using S = Smem<N_OUT, NSPLIT>;
  const size_t smem = S::total(a.cin);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<N_OUT, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      cudaGetLastError();
      set_error("gemm_tc: cannot reserve %zu bytes of shared memory", smem);
      return PMVS_ERR_CUDA;
    }
    configured = smem;
  }
  const int tiles_per_group = cdiv(a.rows_per_group, BM);
  const long long total = (long long)a.groups * tiles_per_group;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static int cached_sms = 0;
  if (cached_sms == 0) cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
  sms = cached_sms > 0 ? cached_sms : 148;
  const int grid = (int)std::min<long long>(total, sms);
  prof_begin(name, st);
  gemm_tc_kernel<N_OUT, NSPLIT><<<grid, NUM_THREADS, smem, st>>>(a, tiles_per_group);
  return check_launch("gemm_tc_kernel", st);
}

}  // namespace tc

// returns -1 if the tensor-core path does not apply (caller falls back to the SIMT kernel)
int launch_gemm_tc(const GemmArgs& a, cudaStream_t st, const char* name) {
  if (tc::g_mode == 0) return -1;
  if (a.cin % 8 != 0 || a.cin > 224 || a.ldx % 4 != 0 || a.ldy % 4 != 0) return -1;
  if (((uintptr_t)a.x & 15) || ((uintptr_t)a.y & 15) || ((uintptr_t)a.w & 15)) return -1;
  const bool x3 = tc::g_mode == 3;
  switch (a.cout) {
    case 16: return x3 ? tc::launch_one<16, 3>(a, st, name) : tc::launch_one<16, 1>(a, st, name);
    case 64: return x3 ? tc::launch_one<64, 3>(a, st, name) : tc::launch_one<64, 1>(a, st, name);
    case 128: return x3 ? tc::launch_one<128, 3>(a, st, name) : tc::launch_one<128, 1>(a, st, name);
  }
  return -1;
}

}  // namespace pmvs

extern "C" int pmvs_set_gemm_mode(int mode) {
  if (mode != 0 && mode != 1 && mode != 3) {
    pmvs::set_error("set_gemm_mode: mode must be 0 (SIMT fp32), 1 (TF32) or 3 (3xTF32)");
    return PMVS_ERR_ARG;
  }
  pmvs::tc::g_mode = mode;
  return PMVS_OK;
}
extern "C" int pmvs_get_gemm_mode(void) { return pmvs::tc::g_mode; }
