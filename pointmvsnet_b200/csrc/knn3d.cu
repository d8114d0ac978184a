// Structured-grid kNN (reference: utils/torch_utils.py:16-61, get_knn_3d).
//
// The reference materialises dist[B,3*k^3,D,H,W] with a one-hot conv3d and runs topk.
// Here a CTA stages an xyz tile plus a zero-filled halo in shared memory (the conv's zero
// padding, torch_utils.py:44: an out-of-grid neighbour IS the zero vector) and every thread
// owns one point.  Keys are 64-bit: (fp32 distance bits << 32) | candidate id with
// id = d*k*k + h*k + w (torch_utils.py:32-38); non-negative floats order like their bit
// patterns, so the unsigned key order IS the canonical order (distance, then candidate id).
// Distances use the reference's rounding sequence: single-rounded differences,
// (dx^2 + dy^2) + dz^2 with no FMA contraction.
//
// Two selection kernels:
//  * knn3d_merge_kernel (kernel_size 5, knn 16 - the hot path, OPT_KNN = 1): the 125 candidates are
//    taken centre-out in 9 batches of 16.  A batch whose smallest distance is larger than the
//    current 16th best IN EVERY LANE of the warp (one warp vote) is dropped after 16 distance
//    evaluations; otherwise it is sorted by a 63-comparator odd-even merge network, merged with
//    the running sorted top-16 by one min step (L[i] = min(L[i], B[15-i])) and re-sorted by a
//    32-comparator bitonic merger.  No data-dependent branches inside a batch, so the lanes of a
//    warp never diverge (the round-1 kernel's sorted insertion ran for the whole warp whenever one
//    lane had to insert: ~10 000 instructions per point).
//  * knn3d_kernel: sorted insertion, every (kernel_size, knn) combination of the public operator.
#include "common.cuh"

namespace pmvs {

constexpr int KNN_TX = 32, KNN_TY = 2, KNN_TD = 5;  // 320 threads per CTA

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

// 16-bit neighbour code consumed by the tile EdgeConv kernels (edge_tile.cu): for a candidate inside the grid the
// row offset (dd+2)*96 + (dh+2)*12 + (dw+2) inside an (8+4) x (4+4) x 5 halo tile, so that the gather address is one
// shift-add; for a candidate OUTSIDE the grid (zero-vector candidate, torch_utils.py:44) bit 15 and the candidate id.
__device__ __forceinline__ unsigned knn_code16(unsigned fd, unsigned fh, unsigned fw, unsigned id, bool in_grid) {
  return in_grid ? fd * 96u + fh * 12u + fw : (0x8000u | id);
}

// picks -> linear indices with the reference's global clamp (torch_utils.py:51-59), plus (optional)
// the 16-bit neighbour codes the tile EdgeConv kernels consume
template <int KS, int K, typename IdxT>
__device__ __forceinline__ void knn_emit(const int (&bi)[K], IdxT* __restrict__ idx_out,
                                         unsigned short* __restrict__ cand_out, long long point, long long n,
                                         int x, int y, int z, int D, int H, int W) {
  constexpr int HK = KS / 2;
  const long long HW = (long long)H * W, DHW = HW * D;
  if (idx_out != nullptr) {
    IdxT* dst = idx_out + point * K;
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
    for (int p = 0; p < K; p += VEC) *reinterpret_cast<int4*>(dst + p) = *reinterpret_cast<const int4*>(&vals[p]);
  }
  if (K == 16 && KS == 5 && cand_out != nullptr) {
    unsigned w8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int p = 0; p < K; ++p) {
      const int j = bi[p];
      const int od = j / (KS * KS) - HK, oh = (j % (KS * KS)) / KS - HK, ow = j % KS - HK;
      const bool in = z + od >= 0 && z + od < D && y + oh >= 0 && y + oh < H && x + ow >= 0 && x + ow < W;
      w8[p >> 1] |= knn_code16((unsigned)(od + HK), (unsigned)(oh + HK), (unsigned)(ow + HK), (unsigned)j, in) << (16 * (p & 1));
    }
    uint4* dst = reinterpret_cast<uint4*>(cand_out + point * 16);
    dst[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
    dst[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
  }
}

template <int KS, int K, typename IdxT>
__global__ void __launch_bounds__(KNN_TX* KNN_TY* KNN_TD)
    knn3d_kernel(const float* __restrict__ xyz, IdxT* __restrict__ idx_out, int D, int H, int W, int dtiles,
                 unsigned short* __restrict__ cand_out) {
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
  const long long n = (long long)z * HW + (long long)y * W + x;
  knn_emit<KS, K, IdxT>(bi, idx_out, cand_out, (long long)cloud * DHW + n, n, x, y, z, D, H, W);
}

// =======================================================================================
// kernel_size 5, knn 16: batched sort + bitonic merge
// =======================================================================================
// CTA = 16 x 4 pixels x 5 layers (320 threads); a warp = 8 x 4 pixels of one layer, so that few warps contain
// lanes next to the cloud border (those need candidates of the outer ring, see the batch order below).
constexpr int KM_TX = 16, KM_TY = 4, KM_TD = 5;
constexpr int KM_SX = KM_TX + 4, KM_SY = KM_TY + 4, KM_SZ = KM_TD + 4;
constexpr int KM_BATCHES = 9;
static_assert(KM_SX == 20 && KM_SY == 8, "c_knn_cand was generated for a 20 x 8 x 9 tile");
// 9 batches of 16 {float4 offset inside the shared-memory tile relative to the point itself, candidate id}:
// batches 0-2 = the 45 candidates of the 3x3 columns (centre-out), then the outer ring BY DIRECTION: 3 = left
// (dw = -2), 4 = right (dw = +2), 5-6 = top (dh = -2), 7-8 = bottom (dh = +2).  A lane at the left border of the
// cloud lacks its left neighbours and needs the `right` batch only, and so on: most warps sort 3-4 batches.
__constant__ int2 c_knn_cand[KM_BATCHES * 16] = {
    {0, 62}, {-160, 37}, {160, 87}, {-320, 12},
    {320, 112}, {-20, 57}, {-180, 32}, {140, 82},
    {-340, 7}, {300, 107}, {-1, 61}, {-161, 36},
    {159, 86}, {-321, 11}, {319, 111}, {1, 63},
    {-159, 38}, {161, 88}, {-319, 13}, {321, 113},
    {20, 67}, {-140, 42}, {180, 92}, {-300, 17},
    {340, 117}, {-21, 56}, {-181, 31}, {139, 81},
    {-341, 6}, {299, 106}, {-19, 58}, {-179, 33},
    {141, 83}, {-339, 8}, {301, 108}, {19, 66},
    {-141, 41}, {179, 91}, {-301, 16}, {339, 116},
    {21, 68}, {-139, 43}, {181, 93}, {-299, 18},
    {341, 118}, {0, 127}, {0, 127}, {0, 127},
    {-2, 60}, {-162, 35}, {158, 85}, {-322, 10},
    {318, 110}, {-22, 55}, {-182, 30}, {138, 80},
    {-342, 5}, {298, 105}, {18, 65}, {-142, 40},
    {178, 90}, {-302, 15}, {338, 115}, {0, 127},
    {2, 64}, {-158, 39}, {162, 89}, {-318, 14},
    {322, 114}, {-18, 59}, {-178, 34}, {142, 84},
    {-338, 9}, {302, 109}, {22, 69}, {-138, 44},
    {182, 94}, {-298, 19}, {342, 119}, {0, 127},
    {-40, 52}, {-200, 27}, {120, 77}, {-360, 2},
    {280, 102}, {-41, 51}, {-201, 26}, {119, 76},
    {-361, 1}, {279, 101}, {-39, 53}, {-199, 28},
    {121, 78}, {-359, 3}, {281, 103}, {-42, 50},
    {-202, 25}, {118, 75}, {-362, 0}, {278, 100},
    {-38, 54}, {-198, 29}, {122, 79}, {-358, 4},
    {282, 104}, {0, 127}, {0, 127}, {0, 127},
    {0, 127}, {0, 127}, {0, 127}, {0, 127},
    {40, 72}, {-120, 47}, {200, 97}, {-280, 22},
    {360, 122}, {39, 71}, {-121, 46}, {199, 96},
    {-281, 21}, {359, 121}, {41, 73}, {-119, 48},
    {201, 98}, {-279, 23}, {361, 123}, {38, 70},
    {-122, 45}, {198, 95}, {-282, 20}, {358, 120},
    {42, 74}, {-118, 49}, {202, 99}, {-278, 24},
    {362, 124}, {0, 127}, {0, 127}, {0, 127},
    {0, 127}, {0, 127}, {0, 127}, {0, 127},
};
__constant__ unsigned c_knn_pad[KM_BATCHES] = {0, 0, 57344, 32768, 32768, 0, 65024, 0, 65024};  // padding slots

#define KNN_SORT16(X) \
  X(0, 1) X(2, 3) X(0, 2) X(1, 3) X(1, 2) X(4, 5) X(6, 7) X(4, 6) \
  X(5, 7) X(5, 6) X(0, 4) X(2, 6) X(2, 4) X(1, 5) X(3, 7) X(3, 5) \
  X(1, 2) X(3, 4) X(5, 6) X(8, 9) X(10, 11) X(8, 10) X(9, 11) X(9, 10) \
  X(12, 13) X(14, 15) X(12, 14) X(13, 15) X(13, 14) X(8, 12) X(10, 14) X(10, 12) \
  X(9, 13) X(11, 15) X(11, 13) X(9, 10) X(11, 12) X(13, 14) X(0, 8) X(4, 12) \
  X(4, 8) X(2, 10) X(6, 14) X(6, 10) X(2, 4) X(6, 8) X(10, 12) X(1, 9) \
  X(5, 13) X(5, 9) X(3, 11) X(7, 15) X(7, 11) X(3, 5) X(7, 9) X(11, 13) \
  X(1, 2) X(3, 4) X(5, 6) X(7, 8) X(9, 10) X(11, 12) X(13, 14)
#define KNN_BITONIC_MERGE16(X) \
  X(0, 8) X(1, 9) X(2, 10) X(3, 11) X(4, 12) X(5, 13) X(6, 14) X(7, 15) \
  X(0, 4) X(1, 5) X(2, 6) X(3, 7) X(8, 12) X(9, 13) X(10, 14) X(11, 15) \
  X(0, 2) X(1, 3) X(4, 6) X(5, 7) X(8, 10) X(9, 11) X(12, 14) X(13, 15) \
  X(0, 1) X(2, 3) X(4, 5) X(6, 7) X(8, 9) X(10, 11) X(12, 13) X(14, 15)

// Keys are compared as fp64: (distance bits << 32 | id) read as a positive double orders exactly like the unsigned
// 64-bit integer (finite fp32 distances give finite doubles), and DSETP issues on the fp64 pipe - the selection
// network is bound by the integer ALU pipe (SEL), so this takes a third of its instructions off that pipe.
constexpr unsigned KNN_PAD_D = 0x7fe00000u;

template <typename IdxT>
__global__ void __launch_bounds__(KM_TX* KM_TY* KM_TD, 2)
    knn3d_merge_kernel(const float* __restrict__ xyz, IdxT* __restrict__ idx_out, int D, int H, int W, int dtiles,
                       unsigned short* __restrict__ cand_out) {
  __shared__ float4 tile[KM_SZ * KM_SY * KM_SX];  // (x, y, z, 0) incl. the zero-filled 2-wide halo
  __shared__ unsigned short s_dec[128];           // candidate id -> (od + 2) | (oh + 2) << 4 | (ow + 2) << 8

  const int cloud = blockIdx.z / dtiles;
  const int z0 = (blockIdx.z % dtiles) * KM_TD;
  const int y0 = blockIdx.y * KM_TY;
  const int x0 = blockIdx.x * KM_TX;
  const long long HW = (long long)H * W;
  const long long DHW = HW * D;
  const float* base = xyz + (long long)cloud * 3 * DHW;

  const int tid = threadIdx.x;
  constexpr int NT = KM_TX * KM_TY * KM_TD;
  if (tid < 125) s_dec[tid] = (unsigned short)((tid / 25) | (((tid % 25) / 5) << 4) | ((tid % 5) << 8));
  for (int e = tid; e < KM_SZ * KM_SY * KM_SX; e += NT) {
    const int sx = e % KM_SX;
    int r = e / KM_SX;
    const int sy = r % KM_SY;
    const int sz = r / KM_SY;
    const int gx = x0 + sx - 2, gy = y0 + sy - 2, gz = z0 + sz - 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // zero padding
    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D) {
      const float* p = base + gz * HW + (long long)gy * W + gx;
      v.x = __ldg(p);
      v.y = __ldg(p + DHW);
      v.z = __ldg(p + 2 * DHW);
    }
    tile[e] = v;
  }
  __syncthreads();

  // warp w: layer w / 2, x half w % 2; lane: 8 x 4 pixels
  const int warp = tid >> 5, lane = tid & 31;
  const int tz = warp >> 1, tx = (warp & 1) * 8 + (lane & 7), ty = lane >> 3;
  const int x = x0 + tx, y = y0 + ty, z = z0 + tz;
  const bool valid = x < W && y < H && z < D;  // invalid threads run along (warp votes below) and skip the store
  const float4* self = tile + ((tz + 2) * KM_SY + ty + 2) * KM_SX + tx + 2;
  const float4 c = *self;

  double L[16];  // running top-16, ascending
#pragma unroll 1
  for (int b = 0; b < KM_BATCHES; ++b) {
    double k[16];
    unsigned dmin = 0xffffffffu;
    const unsigned pad = c_knn_pad[b];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int2 ce = c_knn_cand[b * 16 + q];  // warp-uniform
      const float4 nb = self[ce.x];
      const float ex = __fsub_rn(c.x, nb.x), ey = __fsub_rn(c.y, nb.y), ez = __fsub_rn(c.z, nb.z);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      unsigned db = __float_as_uint(d2);
      if (q >= 9 && ((pad >> q) & 1u)) db = KNN_PAD_D;  // padding slots only occur at q >= 9
      k[q] = __hiloint2double((int)db, ce.y);
      dmin = min(dmin, db);
    }
    if (b > 0) {
      // nothing of this batch can enter any lane's top-16 (equal distances are kept: the id decides)
      if (__all_sync(0xffffffffu, !valid || dmin > (unsigned)__double2hiint(L[15]))) continue;
    }
#define KNN_CE(A, B)                         \
  {                                          \
    const bool sw = k[B] < k[A];             \
    const double t = sw ? k[B] : k[A];       \
    k[B] = sw ? k[A] : k[B];                 \
    k[A] = t;                                \
  }
    KNN_SORT16(KNN_CE)
    if (b == 0) {
#pragma unroll
      for (int q = 0; q < 16; ++q) L[q] = k[q];
    } else {
      // the 16 smallest of the two sorted lists: L[q] = min(L[q], B[15 - q]) is bitonic; re-sort it
#pragma unroll
      for (int q = 0; q < 16; ++q) L[q] = k[15 - q] < L[q] ? k[15 - q] : L[q];
#define KNN_CEL(A, B)                        \
  {                                          \
    const bool sw = L[B] < L[A];             \
    const double t = sw ? L[B] : L[A];       \
    L[B] = sw ? L[A] : L[B];                 \
    L[A] = t;                                \
  }
      KNN_BITONIC_MERGE16(KNN_CEL)
    }
  }
#undef KNN_CE
#undef KNN_CEL
  if (!valid) return;

  const long long n = (long long)z * HW + (long long)y * W + x;
  const long long point = (long long)cloud * DHW + n;
  // in-grid masks: bit (o + 2) set <=> coordinate + o is inside the grid
  unsigned mz = 0u, my = 0u, mx = 0u;
#pragma unroll
  for (int o = -2; o <= 2; ++o) {
    mz |= (z + o >= 0 && z + o < D) ? 1u << (o + 2) : 0u;
    my |= (y + o >= 0 && y + o < H) ? 1u << (o + 2) : 0u;
    mx |= (x + o >= 0 && x + o < W) ? 1u << (o + 2) : 0u;
  }
  unsigned w8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  __align__(16) IdxT vals[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const unsigned j = (unsigned)__double2loint(L[p]);
    const unsigned f = s_dec[j & 127u];
    const unsigned fd = f & 15u, fh = (f >> 4) & 15u, fw = f >> 8;
    const unsigned in = (mz >> fd) & (my >> fh) & (mx >> fw) & 1u;
    w8[p >> 1] |= knn_code16(fd, fh, fw, j & 127u, in != 0u) << (16 * (p & 1));
    if (idx_out != nullptr) {
      // linear index with the reference's global clamp (torch_utils.py:51-59)
      long long t = n + ((long long)fd - 2) * HW + ((long long)fh - 2) * W + ((long long)fw - 2);
      t = t < 0 ? 0 : (t > DHW - 1 ? DHW - 1 : t);
      vals[p] = (IdxT)t;
    }
  }
  if (idx_out != nullptr) {
    constexpr int VEC = 16 / sizeof(IdxT);
    IdxT* dst = idx_out + point * 16;
#pragma unroll
    for (int p = 0; p < 16; p += VEC) *reinterpret_cast<int4*>(dst + p) = *reinterpret_cast<const int4*>(&vals[p]);
  }
  if (cand_out != nullptr) {
    uint4* dst = reinterpret_cast<uint4*>(cand_out + point * 16);
    dst[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
    dst[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
  }
}

template <int KS, int K>
static int launch_ks_k(const float* xyz, int64_t* idx64, int32_t* idx32, unsigned short* cand, int clouds, int D, int H,
                       int W, cudaStream_t st) {
  const int dtiles = cdiv(D, KNN_TD);
  dim3 block(KNN_TX, KNN_TY, KNN_TD);
  dim3 grid(cdiv(W, KNN_TX), cdiv(H, KNN_TY), clouds * dtiles);
  prof_begin("knn3d", st);
  if (KS == 5 && K == 16 && opt(OPT_KNN) != 0) {
    dim3 mgrid(cdiv(W, KM_TX), cdiv(H, KM_TY), clouds * dtiles);
    if (idx64)
      knn3d_merge_kernel<int64_t><<<mgrid, KM_TX * KM_TY * KM_TD, 0, st>>>(xyz, idx64, D, H, W, dtiles, cand);
    else
      knn3d_merge_kernel<int32_t><<<mgrid, KM_TX * KM_TY * KM_TD, 0, st>>>(xyz, idx32, D, H, W, dtiles, cand);
  } else if (idx64) {
    knn3d_kernel<KS, K, int64_t><<<grid, block, 0, st>>>(xyz, idx64, D, H, W, dtiles, cand);
  } else {
    knn3d_kernel<KS, K, int32_t><<<grid, block, 0, st>>>(xyz, idx32, D, H, W, dtiles, cand);
  }
  return check_launch("knn3d_kernel", st);
}

template <int KS>
static int launch_ks(const float* xyz, int64_t* idx64, int32_t* idx32, int clouds, int D, int H, int W, int knn,
                     cudaStream_t st) {
  switch (knn) {
    case 4: return launch_ks_k<KS, 4>(xyz, idx64, idx32, nullptr, clouds, D, H, W, st);
    case 8: return launch_ks_k<KS, 8>(xyz, idx64, idx32, nullptr, clouds, D, H, W, st);
    case 16: return launch_ks_k<KS, 16>(xyz, idx64, idx32, nullptr, clouds, D, H, W, st);
    case 20: return launch_ks_k<KS, 20>(xyz, idx64, idx32, nullptr, clouds, D, H, W, st);
    case 32: return launch_ks_k<KS, 32>(xyz, idx64, idx32, nullptr, clouds, D, H, W, st);
  }
  set_error("knn3d: unsupported knn=%d (supported: 4, 8, 16, 20, 32)", knn);
  return PMVS_ERR_ARG;
}

int launch_knn3d_cand(const float* xyz, int32_t* idx32, unsigned short* cand, int clouds, int D, int H, int W,
                      cudaStream_t st) {
  PMVS_REQUIRE(xyz && cand, "knn3d_cand: NULL pointer");
  PMVS_REQUIRE(clouds > 0 && D > 0 && H > 0 && W > 0, "knn3d: empty input");
  PMVS_REQUIRE((long long)clouds * cdiv(D, KNN_TD) <= 65535, "knn3d: too many clouds");
  PMVS_REQUIRE((long long)D * H * W < (1ll << 31), "knn3d: cloud too large for int32 indices");
  return launch_ks_k<5, 16>(xyz, nullptr, idx32, cand, clouds, D, H, W, st);
}

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
