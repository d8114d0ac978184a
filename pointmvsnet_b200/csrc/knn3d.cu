// Structured-grid kNN (reference: utils/torch_utils.py:16-61, get_knn_3d).
//
// The reference materialises dist[B,3*k^3,D,H,W] with a one-hot conv3d and runs topk.
// Here a CTA stages an xyz tile plus a zero-filled halo in shared memory (the conv's zero
// padding, torch_utils.py:44: an out-of-grid neighbour IS the zero vector), every thread
// owns one point and keeps a sorted top-K of 64-bit keys (fp32 distance bits << 32 | candidate
// id, id = d*k*k + h*k + w as in torch_utils.py:32-38) in registers: non-negative floats order
// like their bit patterns, so an unsigned key comparison IS the canonical order (distance, then
// candidate id) and the candidates can be scanned centre-out - the near ones first, so that most
// later candidates fail the single "worse than the current K-th" test and are never inserted.
// Distances use the reference's rounding sequence: single-rounded differences,
// (dx^2 + dy^2) + dz^2 with no FMA contraction.
#include "common.cuh"

namespace pmvs {

constexpr int KNN_TX = 32, KNN_TY = 2, KNN_TD = 5;  // 320 threads: 3 CTAs/SM at 62 registers

template <int K>
__device__ __forceinline__ void knn_insert(unsigned (&kd)[K], unsigned (&ki)[K], unsigned d, unsigned j) {
  // keys are (kd, ki) pairs compared as one 64-bit unsigned number
  const unsigned long long key = ((unsigned long long)d << 32) | j;
#pragma unroll
  for (int p = K - 1; p > 0; --p) {
    const unsigned long long prev = ((unsigned long long)kd[p - 1] << 32) | ki[p - 1];
    const unsigned long long cur = ((unsigned long long)kd[p] << 32) | ki[p];
    const bool lt_prev = key < prev;
    const bool lt_cur = key < cur;
    kd[p] = lt_prev ? kd[p - 1] : (lt_cur ? d : kd[p]);
    ki[p] = lt_prev ? ki[p - 1] : (lt_cur ? j : ki[p]);
  }
  const unsigned long long first = ((unsigned long long)kd[0] << 32) | ki[0];
  if (key < first) {
    kd[0] = d;
    ki[0] = j;
  }
}

// centre-out scan order of the (dh, dw) window offsets and of the depth offsets
__constant__ signed char c_ring5[25][2] = {{0, 0},  {-1, 0},  {0, -1}, {0, 1},  {1, 0},   {-1, -1}, {-1, 1}, {1, -1}, {1, 1},
                                           {-2, 0}, {0, -2},  {0, 2},  {2, 0},  {-2, -1}, {-2, 1},  {-1, -2}, {-1, 2}, {1, -2},
                                           {1, 2},  {2, -1},  {2, 1},  {-2, -2}, {-2, 2}, {2, -2},  {2, 2}};
__constant__ signed char c_ring3[9][2] = {{0, 0}, {-1, 0}, {0, -1}, {0, 1}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};

template <int KS, int K, typename IdxT>
__global__ void __launch_bounds__(KNN_TX* KNN_TY* KNN_TD)
    knn3d_kernel(const float* __restrict__ xyz, IdxT* __restrict__ idx_out, int D, int H, int W, int dtiles
#if PMVS_EDGE_TILE
                 , unsigned char* __restrict__ cand_out  // optional [clouds*D*H*W, K] candidate ids (K == 16)
#endif
    ) {
  constexpr int HK = KS / 2;
  constexpr int SX = KNN_TX + 2 * HK, SY = KNN_TY + 2 * HK, SZ = KNN_TD + 2 * HK;
  __shared__ float tile[3][SZ][SY][SX];

  const int cloud = blockIdx.z / dtiles;
  const int z0 = (blockIdx.z % dtiles) * KNN_TD;
  const int y0 = blockIdx.y * KNN_TY;
  const int x0 = blockIdx.x * KNN_TX;
  const long long HW = (long long)H * W;
  const long long DHW = HW * D;
  const float* base = xyz + (long long)cloud * 3 * DHW;

  const int tid = (threadIdx.z * KNN_TY + threadIdx.y) * KNN_TX + threadIdx.x;
  constexpr int NT = KNN_TX * KNN_TY * KNN_TD;
  for (int e = tid; e < 3 * SZ * SY * SX; e += NT) {
    int sx = e % SX;
    int r = e / SX;
    int sy = r % SY;
    r /= SY;
    int sz = r % SZ;
    int c = r / SZ;
    int gx = x0 + sx - HK, gy = y0 + sy - HK, gz = z0 + sz - HK;
    float v = 0.f;  // zero padding
    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D)
      v = __ldg(base + c * DHW + gz * HW + (long long)gy * W + gx);
    (&tile[0][0][0][0])[e] = v;
  }
  __syncthreads();

  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y, z = z0 + threadIdx.z;
  if (x >= W || y >= H || z >= D) return;

  const int cxs = threadIdx.x + HK, cys = threadIdx.y + HK, czs = threadIdx.z + HK;
  const float cx = tile[0][czs][cys][cxs];
  const float cy = tile[1][czs][cys][cxs];
  const float cz = tile[2][czs][cys][cxs];

  unsigned kd[K], ki[K];   // distance bits / candidate id, sorted ascending as 64-bit keys
#pragma unroll
  for (int p = 0; p < K; ++p) {
    kd[p] = 0x7f800000u;  // +inf
    ki[p] = 0xffffffffu;
  }

#pragma unroll 1
  for (int ring = 0; ring < KS * KS; ++ring) {
    const int dh = KS == 5 ? c_ring5[ring][0] : c_ring3[ring][0];
    const int dw = KS == 5 ? c_ring5[ring][1] : c_ring3[ring][1];
    const int sy = threadIdx.y + HK + dh, sx = threadIdx.x + HK + dw;
#pragma unroll
    for (int q = 0; q < KS; ++q) {
      // depth offsets centre-out: 0, -1, +1, -2, +2
      const int dd = (q == 0) ? 0 : ((q & 1) ? -((q + 1) >> 1) : (q >> 1));
      const int sz = threadIdx.z + HK + dd;
      const float ex = __fsub_rn(cx, tile[0][sz][sy][sx]);
      const float ey = __fsub_rn(cy, tile[1][sz][sy][sx]);
      const float ez = __fsub_rn(cz, tile[2][sz][sy][sx]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      const unsigned db = __float_as_uint(d2);
      const unsigned j = (unsigned)(((dd + HK) * KS + (dh + HK)) * KS + (dw + HK));
      // worse than the current K-th (distance, then id)?  NaN distances never enter (db > +inf bits)
      if (db < kd[K - 1] || (db == kd[K - 1] && j < ki[K - 1])) knn_insert<K>(kd, ki, db, j);
    }
  }
  int bi[K];
#pragma unroll
  for (int p = 0; p < K; ++p) bi[p] = (int)ki[p];

  // candidate id -> linear index with the reference's global clamp (torch_utils.py:51-59)
  const long long n = (long long)z * HW + (long long)y * W + x;
  IdxT* dst = idx_out + ((long long)cloud * DHW + n) * K;
  __align__(16) IdxT vals[K];
#pragma unroll
  for (int p = 0; p < K; ++p) {
    const int j = bi[p];
    const int od = j / (KS * KS) - HK;
    const int oh = (j % (KS * KS)) / KS - HK;
    const int ow = j % KS - HK;
    long long t = n + od * HW + (long long)oh * W + ow;
    t = t < 0 ? 0 : (t > DHW - 1 ? DHW - 1 : t);
    vals[p] = (IdxT)t;
  }
  constexpr int VEC = 16 / sizeof(IdxT);
#pragma unroll
  for (int p = 0; p < K; p += VEC) {
    *reinterpret_cast<int4*>(dst + p) = *reinterpret_cast<const int4*>(&vals[p]);
  }
#if PMVS_EDGE_TILE
  if (K == 16 && cand_out != nullptr) {
    // candidate id of every pick, 255 when the candidate lies outside the grid (its linear index is
    // clamped / aliases another row, torch_utils.py:51-59, and must be taken from idx_out)
    unsigned w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int p = 0; p < K; ++p) {
      const int j = bi[p];
      const int od = j / (KS * KS) - HK, oh = (j % (KS * KS)) / KS - HK, ow = j % KS - HK;
      const bool in = z + od >= 0 && z + od < D && y + oh >= 0 && y + oh < H && x + ow >= 0 && x + ow < W;
      w4[p >> 2] |= (in ? (unsigned)j : 255u) << (8 * (p & 3));
    }
    *reinterpret_cast<uint4*>(cand_out + ((long long)cloud * DHW + n) * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
#endif
}

template <int KS, int K>
static int launch_ks_k(const float* xyz, int64_t* idx64, int32_t* idx32, int clouds, int D, int H, int W,
                       cudaStream_t st) {
  const int dtiles = cdiv(D, KNN_TD);
  dim3 block(KNN_TX, KNN_TY, KNN_TD);
  dim3 grid(cdiv(W, KNN_TX), cdiv(H, KNN_TY), clouds * dtiles);
  prof_begin("knn3d", st);
#if PMVS_EDGE_TILE
  if (idx32)
    knn3d_kernel<KS, K, int32_t><<<grid, block, 0, st>>>(xyz, idx32, D, H, W, dtiles, nullptr);
  else
    knn3d_kernel<KS, K, int64_t><<<grid, block, 0, st>>>(xyz, idx64, D, H, W, dtiles, nullptr);
#else
  if (idx32)
    knn3d_kernel<KS, K, int32_t><<<grid, block, 0, st>>>(xyz, idx32, D, H, W, dtiles);
  else
    knn3d_kernel<KS, K, int64_t><<<grid, block, 0, st>>>(xyz, idx64, D, H, W, dtiles);
#endif
  return check_launch("knn3d_kernel", st);
}

template <int KS>
static int launch_ks(const float* xyz, int64_t* idx64, int32_t* idx32, int clouds, int D, int H, int W, int knn,
                     cudaStream_t st) {
  switch (knn) {
    case 4: return launch_ks_k<KS, 4>(xyz, idx64, idx32, clouds, D, H, W, st);
    case 8: return launch_ks_k<KS, 8>(xyz, idx64, idx32, clouds, D, H, W, st);
    case 16: return launch_ks_k<KS, 16>(xyz, idx64, idx32, clouds, D, H, W, st);
    case 20: return launch_ks_k<KS, 20>(xyz, idx64, idx32, clouds, D, H, W, st);
    case 32: return launch_ks_k<KS, 32>(xyz, idx64, idx32, clouds, D, H, W, st);
  }
  set_error("knn3d: unsupported knn=%d (supported: 4, 8, 16, 20, 32)", knn);
  return PMVS_ERR_ARG;
}

#if PMVS_EDGE_TILE
int launch_knn3d_cand(const float* xyz, int32_t* idx32, unsigned char* cand, int clouds, int D, int H, int W,
                      cudaStream_t st) {
  PMVS_REQUIRE(xyz && idx32 && cand, "knn3d_cand: NULL pointer");
  PMVS_REQUIRE(clouds > 0 && D > 0 && H > 0 && W > 0, "knn3d: empty input");
  PMVS_REQUIRE((long long)clouds * cdiv(D, KNN_TD) <= 65535, "knn3d: too many clouds");
  PMVS_REQUIRE((long long)D * H * W < (1ll << 31), "knn3d: cloud too large for int32 indices");
  const int dtiles = cdiv(D, KNN_TD);
  dim3 block(KNN_TX, KNN_TY, KNN_TD);
  dim3 grid(cdiv(W, KNN_TX), cdiv(H, KNN_TY), clouds * dtiles);
  prof_begin("knn3d", st);
  knn3d_kernel<5, 16, int32_t><<<grid, block, 0, st>>>(xyz, idx32, D, H, W, dtiles, cand);
  return check_launch("knn3d_kernel", st);
}
#endif

int launch_knn3d(const float* xyz, int64_t* idx64, int32_t* idx32, int clouds, int D, int H, int W, int ksize,
                 int knn, cudaStream_t st) {
  PMVS_REQUIRE(xyz && (idx64 || idx32) && !(idx64 && idx32), "knn3d: need xyz and exactly one output");
  PMVS_REQUIRE(clouds > 0 && D > 0 && H > 0 && W > 0, "knn3d: empty input");
  PMVS_REQUIRE(knn <= ksize * ksize * ksize, "knn3d: knn=%d exceeds window size %d^3", knn, ksize);
  PMVS_REQUIRE((long long)clouds * cdiv(D, KNN_TD) <= 65535, "knn3d: too many clouds");
  if (idx32) PMVS_REQUIRE((long long)D * H * W < (1ll << 31), "knn3d: cloud too large for int32 indices");
  if (ksize == 5) return launch_ks<5>(xyz, idx64, idx32, clouds, D, H, W, knn, st);
  if (ksize == 3) return launch_ks<3>(xyz, idx64, idx32, clouds, D, H, W, knn, st);
  set_error("knn3d: unsupported kernel_size=%d (supported: 3, 5)", ksize);
  return PMVS_ERR_ARG;
}

}  // namespace pmvs

extern "C" int pmvs_knn3d(const float* xyz, int64_t* idx64, int32_t* idx32, int B, int D, int H, int W, int ksize,
                          int knn, pmvs_stream_t stream) {
  return pmvs::launch_knn3d(xyz, idx64, idx32, B, D, H, W, ksize, knn, (cudaStream_t)stream);
}
