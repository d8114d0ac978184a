// Homography warp + multi-view feature fetch.
//
//  (1) pmvs_feature_fetch*: the stand-alone FeatureFetcher operator
//      (reference utils/feature_fetcher.py:13-60), generic C / NCHW.
//  (2) fused_fetch_kernel: rows a2-a9 of the hot path (reference model.py:153-204) in one
//      launch: nearest depth upsample, pixel grid, hypothesis un-projection, projection into
//      every view, pyramid resize composed with the bilinear fetch, variance over views,
//      xyz normalisation, and the 136-channel point feature written points-major in the
//      sub-cloud order the EdgeConv kernels consume.  One warp owns one pixel: lanes span
//      the 112 pyramid channels (16 lanes x float4 on conv3, 8 on conv2, 4 on conv1), so a
//      tap is one fully used 256/128/64-byte segment of the channels-last pyramid.
//      Camera matrices for the CTA's batch element are staged into shared memory with one
//      cp.async.bulk (TMA bulk copy) completing on an mbarrier.
#include "common.cuh"

namespace pmvs {

// ---------------------------------------------------------------------------------------
// camera block, one per batch element (floats)
// ---------------------------------------------------------------------------------------
constexpr int CB_KINV = 0;    // inverse of the scaled reference intrinsics, 3x3 row-major
constexpr int CB_R0INV = 9;   // inverse reference rotation
constexpr int CB_T0 = 18;     // reference translation
constexpr int CB_MEAN = 21;
constexpr int CB_STD = 24;
constexpr int CB_INTERVAL = 27;
constexpr int CB_VIEW = 28;   // per view: R[9], t[3], K[9] (scaled), pad[3]
constexpr int CB_VSTRIDE = 24;
__host__ __device__ constexpr int cam_block_floats(int V) { return CB_VIEW + CB_VSTRIDE * V; }

__device__ void inv3x3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double r = 1.0 / det;
  o[0] = A * r;
  o[1] = -(b * i - c * h) * r;
  o[2] = (b * f - c * e) * r;
  o[3] = B * r;
  o[4] = (a * i - c * g) * r;
  o[5] = -(a * f - c * d) * r;
  o[6] = C * r;
  o[7] = -(a * h - b * g) * r;
  o[8] = (a * e - b * d) * r;
}

// cam_params [B,V,2,4,4] (io.py:31-45) -> camera blocks.  Mirrors model.py:54-57 (R, t,
// R_inv), :159-163 (K rows 0,1 scaled), :169 (inverse of the reference K).
__global__ void cam_setup_kernel(const float* __restrict__ cam_params, const float* __restrict__ interval,
                                 const float* __restrict__ mean, const float* __restrict__ stdv,
                                 float* __restrict__ blocks, int B, int V, float kscale, float iscale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* out = blocks + (size_t)b * cam_block_floats(V);
  for (int v = 0; v < V; ++v) {
    const float* ext = cam_params + ((size_t)(b * V + v) * 2 + 0) * 16;
    const float* intr = cam_params + ((size_t)(b * V + v) * 2 + 1) * 16;
    float* o = out + CB_VIEW + v * CB_VSTRIDE;
    float K[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        o[r * 3 + c] = ext[r * 4 + c];
        float k = intr[r * 4 + c];
        if (r < 2) k = __fmul_rn(k, kscale);
        K[r * 3 + c] = k;
        o[12 + r * 3 + c] = k;
      }
    for (int r = 0; r < 3; ++r) o[9 + r] = ext[r * 4 + 3];
    o[21] = o[22] = o[23] = 0.f;
    if (v == 0) {
      double m[9], inv[9];
      for (int q = 0; q < 9; ++q) m[q] = (double)K[q];
      inv3x3(m, inv);
      for (int q = 0; q < 9; ++q) out[CB_KINV + q] = (float)inv[q];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) m[r * 3 + c] = (double)ext[r * 4 + c];
      inv3x3(m, inv);
      for (int q = 0; q < 9; ++q) out[CB_R0INV + q] = (float)inv[q];
      for (int r = 0; r < 3; ++r) out[CB_T0 + r] = ext[r * 4 + 3];
    }
  }
  for (int r = 0; r < 3; ++r) {
    out[CB_MEAN + r] = mean ? mean[b * 3 + r] : 0.f;
    out[CB_STD + r] = stdv ? stdv[b * 3 + r] : 1.f;
  }
  out[CB_INTERVAL] = interval ? __fmul_rn(iscale, interval[b]) : 0.f;
}

// ---------------------------------------------------------------------------------------
// shared projection math (feature_fetcher.py:36-53 + grid_sample's un-normalisation)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3(const float* r, float x, float y, float z) {
  return fmaf(r[2], z, fmaf(r[1], y, __fmul_rn(r[0], x)));
}

// pixel coordinate in the sampled map (align_corners=True round trip, feature_fetcher.py:51-53
// then ATen grid_sampler_unnormalize): ((g + 1) / 2) * (size - 1), g = (u - .5)/(size-1)*2 - 1
__device__ __forceinline__ float grid_coord(float u, int size) {
  const float sm1 = (float)(size - 1);
  const float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(u, 0.5f), sm1), 2.f), 1.f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), sm1);
}

__device__ __forceinline__ void project(const float* R, const float* t, const float* K, float wx, float wy, float wz,
                                        float& u, float& v) {
  float xc = wx, yc = wy, zc = wz;
  if (R != nullptr) {
    xc = __fadd_rn(dot3(R + 0, wx, wy, wz), t[0]);
    yc = __fadd_rn(dot3(R + 3, wx, wy, wz), t[1]);
    zc = __fadd_rn(dot3(R + 6, wx, wy, wz), t[2]);
  }
  const float nx = __fdiv_rn(xc, zc), ny = __fdiv_rn(yc, zc);
  u = dot3(K + 0, nx, ny, 1.f);
  v = dot3(K + 3, nx, ny, 1.f);
}

__device__ __forceinline__ bool usable(float c) { return fabsf(c) < 1.0e8f; }  // false for NaN/inf

// ---------------------------------------------------------------------------------------
// (1) stand-alone FeatureFetcher
// ---------------------------------------------------------------------------------------
struct Taps {
  int x0, y0;
  float nw, ne, sw, se;
  bool ok_w, ok_e, ok_n, ok_s;
};
__device__ __forceinline__ Taps make_taps(float ix, float iy, int W, int H) {
  Taps t;
  const float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  const float ex = fx + 1.f, ey = fy + 1.f;
  t.nw = __fmul_rn(__fsub_rn(ex, ix), __fsub_rn(ey, iy));
  t.ne = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(ey, iy));
  t.sw = __fmul_rn(__fsub_rn(ex, ix), __fsub_rn(iy, fy));
  t.se = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
  t.ok_w = t.x0 >= 0 && t.x0 < W;
  t.ok_e = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.ok_n = t.y0 >= 0 && t.y0 < H;
  t.ok_s = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  return t;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256)
    feature_fetch_kernel(const float* __restrict__ maps, const float* __restrict__ pts, const float* __restrict__ Kmat,
                         const float* __restrict__ Emat, float* __restrict__ out, float* __restrict__ grad_maps,
                         int V, int C, int H, int W, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int bv = blockIdx.y;
  if (n >= N) return;
  const int b = bv / V;
  const float wx = pts[((size_t)b * 3 + 0) * N + n];
  const float wy = pts[((size_t)b * 3 + 1) * N + n];
  const float wz = pts[((size_t)b * 3 + 2) * N + n];
  float R[9], t[3], K[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) K[q] = Kmat[(size_t)bv * 9 + q];
  if (Emat != nullptr) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) R[r * 3 + c] = Emat[(size_t)bv * 12 + r * 4 + c];
      t[r] = Emat[(size_t)bv * 12 + r * 4 + 3];
    }
  }
  float u, v;
  project(Emat ? R : nullptr, t, K, wx, wy, wz, u, v);
  const float ix = grid_coord(u, W), iy = grid_coord(v, H);
  const bool ok = usable(ix) && usable(iy);
  Taps tp = make_taps(ok ? ix : -10.f, ok ? iy : -10.f, W, H);
  const size_t plane = (size_t)H * W;
  const size_t o_nw = (size_t)tp.y0 * W + tp.x0;
  for (int c = 0; c < C; ++c) {
    const size_t cb = ((size_t)bv * C + c) * plane;
    if (!BACKWARD) {
      const float* m = maps + cb;
      float acc = 0.f;
      if (tp.ok_n && tp.ok_w) acc = __fmul_rn(__ldg(m + o_nw), tp.nw);
      if (tp.ok_n && tp.ok_e) acc = fmaf(__ldg(m + o_nw + 1), tp.ne, acc);
      if (tp.ok_s && tp.ok_w) acc = fmaf(__ldg(m + o_nw + W), tp.sw, acc);
      if (tp.ok_s && tp.ok_e) acc = fmaf(__ldg(m + o_nw + W + 1), tp.se, acc);
      out[((size_t)bv * C + c) * N + n] = acc;
    } else {
      const float g = out[((size_t)bv * C + c) * N + n];  // `out` carries grad_out here
      float* m = grad_maps + cb;
      if (tp.ok_n && tp.ok_w) atomicAdd(m + o_nw, g * tp.nw);
      if (tp.ok_n && tp.ok_e) atomicAdd(m + o_nw + 1, g * tp.ne);
      if (tp.ok_s && tp.ok_w) atomicAdd(m + o_nw + W, g * tp.sw);
      if (tp.ok_s && tp.ok_e) atomicAdd(m + o_nw + W + 1, g * tp.se);
    }
  }
}

// ---------------------------------------------------------------------------------------
// (2) fused warp + fetch + variance
// ---------------------------------------------------------------------------------------

// One axis of "bilinear sample of the bilinearly resized map" as <= 4 (index, weight)
// pairs into the NATIVE map.  The resized map (F.interpolate, align_corners=False,
// model.py:184) is never materialised: tap r of the sampled map is
//   l0 * native[p0] + l1 * native[p1]   (ATen upsample_bilinear2d source index rule)
// and the two sample taps (floor, floor+1) carry the grid_sample weights
// (ATen grid_sampler_2d, zeros padding: out-of-range taps contribute nothing).
// Duplicate native indices are folded and the non-zero entries are compacted to the front,
// so the consumer loops stop at the first zero weight.
struct Axis {
  int i[4];
  float w[4];
};
__device__ __forceinline__ void axis_entries(float coord, int out_size, int in_size, float scale, bool valid,
                                             Axis& a) {
  const float f = floorf(coord);
  const int r0 = (int)f;
  const float wt[2] = {__fsub_rn(f + 1.f, coord), __fsub_rn(coord, f)};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int r = r0 + s;
    const bool inb = valid && r >= 0 && r < out_size;
    float src = __fsub_rn(__fmul_rn(scale, (float)r + 0.5f), 0.5f);
    if (src < 0.f) src = 0.f;
    int p0 = (int)src;
    if (p0 > in_size - 1) p0 = in_size - 1;
    const int p1 = p0 + (p0 < in_size - 1 ? 1 : 0);
    const float l1 = __fsub_rn(src, (float)p0);
    const float l0 = __fsub_rn(1.f, l1);
    a.i[2 * s] = p0;
    a.i[2 * s + 1] = p1;
    a.w[2 * s] = inb ? __fmul_rn(wt[s], l0) : 0.f;
    a.w[2 * s + 1] = inb ? __fmul_rn(wt[s], l1) : 0.f;
  }
  // fold duplicate native indices so every native texel is loaded once
  if (a.i[1] == a.i[0]) { a.w[0] += a.w[1]; a.w[1] = 0.f; }
  if (a.i[3] == a.i[2]) { a.w[2] += a.w[3]; a.w[3] = 0.f; }
  if (a.i[2] == a.i[0]) { a.w[0] += a.w[2]; a.w[2] = 0.f; }
  else if (a.i[2] == a.i[1]) { a.w[1] += a.w[2]; a.w[2] = 0.f; }
  if (a.i[3] == a.i[1]) { a.w[1] += a.w[3]; a.w[3] = 0.f; }
  else if (a.i[3] == a.i[0]) { a.w[0] += a.w[3]; a.w[3] = 0.f; }
  // compaction: bubble zero weights to the back (stable for the non-zero entries)
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
    for (int j = 0; j < 3 - pass; ++j) {
      const bool sw = a.w[j] == 0.f;
      const float tw = a.w[j];
      const int ti = a.i[j];
      a.w[j] = sw ? a.w[j + 1] : tw;
      a.i[j] = sw ? a.i[j + 1] : ti;
      a.w[j + 1] = sw ? tw : a.w[j + 1];
      a.i[j + 1] = sw ? ti : a.i[j + 1];
    }
  }
}

constexpr int FETCH_WARPS = 8;
constexpr int FETCH_TRIPLES_PER_ROUND = 10;  // 30 lanes build 10 (hypothesis, view) triples of level descriptors

// Sampling descriptor of one (hypothesis, view, level): element offsets (already multiplied
// by the channel count / row pitch) and weights per axis, plus the loop bounds shared by the
// axis tap counts (the tap loops run to the warp-wide maximum of the step).
struct __align__(16) Desc {
  unsigned xo[4];   // byte offsets of the native texels along x (index * C * 4)
  float xw[4];
  unsigned yo[4];   // byte offsets of (view, row)
  float yw[4];
  int nx, ny, pad0, pad1;
};
static_assert(sizeof(Desc) == 80, "Desc layout");

__host__ __device__ constexpr size_t fetch_smem_bytes(int V) {  // descriptors
  return (size_t)FETCH_WARPS * PMVS_NUM_HYP * V * 3 * sizeof(Desc);
}
__host__ __device__ constexpr size_t fetch_smem_total(int V) {  // + 16 floats of xyz per warp
  return fetch_smem_bytes(V) + FETCH_WARPS * 16 * sizeof(float);
}

// One pyramid level L (0: conv1 16 ch, 1: conv2 32 ch, 2: conv3 64 ch) for the 5 hypotheses of a
// pixel; see "phase 2" in fused_fetch_kernel.  A level with C channels needs C/4 lanes per sample,
// so 32/(C/4) HYPOTHESES are sampled side by side (2 on conv3, 4 on conv2, 8 >= 5 on conv1) and the
// views are walked sequentially: every lane owns its hypothesis' running sum / sum of squares in
// view order (model.py:188-189) and no cross-lane reduction is needed.  The hypotheses of a pixel
// project to nearly the same place in a view, so their tap counts agree and the warp-uniform
// loop bounds are tight.
template <int L>
__device__ __forceinline__ void level_pass(const FusedFetchParams& p, const Desc* desc, int b, int V, int lane,
                                           float rV, float* frow0, size_t fstep) {
  constexpr int C = 16 << L;
  constexpr int G = C / 4;        // lanes per sample
  constexpr int S = 32 / G;       // hypotheses sampled per step
  constexpr int CH_OFF = L == 2 ? 48 : (L == 1 ? 16 : 0);
  const int grp = lane / G, cq = lane % G;
  const char* lbase = reinterpret_cast<const char*>(p.pyr[L] + cq * 4 + (size_t)b * V * p.hl[L] * p.wl[L] * C);
#pragma unroll 1
  for (int m0 = 0; m0 < PMVS_NUM_HYP; m0 += S) {
    const int m = m0 + grp;
    const bool active = m < PMVS_NUM_HYP;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      Desc dd = desc[((active ? m : 0) * V + v) * 3 + L];
      if (!active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { dd.xw[j] = 0.f; dd.yw[j] = 0.f; }
        dd.nx = 0; dd.ny = 0;
      }
      const int nxm = __reduce_max_sync(0xffffffffu, dd.nx);
      const int nym = __reduce_max_sync(0xffffffffu, dd.ny);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ey = 0; ey < 4; ++ey) {
        if (ey < nym) {  // warp-uniform bound
          const char* row = lbase + dd.yo[ey];
#pragma unroll
          for (int ex = 0; ex < 4; ++ex) {
            if (ex < nxm) {  // warp-uniform bound
              const float wgt = __fmul_rn(dd.yw[ey], dd.xw[ex]);
              const float4 t = __ldg(reinterpret_cast<const float4*>(row + dd.xo[ex]));
              acc.x = fmaf(wgt, t.x, acc.x);
              acc.y = fmaf(wgt, t.y, acc.y);
              acc.z = fmaf(wgt, t.z, acc.z);
              acc.w = fmaf(wgt, t.w, acc.w);
            }
          }
        }
      }
      // model.py:188-189: sums over views of x and x**2, in view order
      s1.x = __fadd_rn(s1.x, acc.x); s1.y = __fadd_rn(s1.y, acc.y);
      s1.z = __fadd_rn(s1.z, acc.z); s1.w = __fadd_rn(s1.w, acc.w);
      s2.x = __fadd_rn(s2.x, __fmul_rn(acc.x, acc.x)); s2.y = __fadd_rn(s2.y, __fmul_rn(acc.y, acc.y));
      s2.z = __fadd_rn(s2.z, __fmul_rn(acc.z, acc.z)); s2.w = __fadd_rn(s2.w, __fmul_rn(acc.w, acc.w));
    }
    if (active) {
      // model.py:190: mean(x^2) - mean(x)^2 (difference unfused); mean = sum * (1/V) as ATen's
      // CUDA mean kernel computes it (identical to sum / V for V a power of two)
      float4 o;
      float a;
      a = __fmul_rn(s1.x, rV); o.x = __fsub_rn(__fmul_rn(s2.x, rV), __fmul_rn(a, a));
      a = __fmul_rn(s1.y, rV); o.y = __fsub_rn(__fmul_rn(s2.y, rV), __fmul_rn(a, a));
      a = __fmul_rn(s1.z, rV); o.z = __fsub_rn(__fmul_rn(s2.z, rV), __fmul_rn(a, a));
      a = __fmul_rn(s1.w, rV); o.w = __fsub_rn(__fmul_rn(s2.w, rV), __fmul_rn(a, a));
      st4(frow0 + m * fstep + CH_OFF + cq * 4, o);
    }
  }
}

__global__ void __launch_bounds__(FETCH_WARPS * 32) fused_fetch_kernel(const FusedFetchParams p) {
  __shared__ __align__(16) float cam[cam_block_floats(PMVS_MAX_VIEWS)];
  __shared__ __align__(8) unsigned long long bar;
  extern __shared__ __align__(16) unsigned char dyn_smem[];

  const int b = blockIdx.y;
  const int V = p.V;
  // --- stage this batch element's camera block: one TMA bulk copy + mbarrier ------------
  const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
  const unsigned cam_addr = (unsigned)__cvta_generic_to_shared(cam);
  const unsigned bytes = (unsigned)(cam_block_floats(V) * sizeof(float));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float* src = p.cam_blocks + (size_t)b * cam_block_floats(V);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cam_addr),
        "l"(src), "r"(bytes), "r"(bar_addr)
        : "memory");
  }
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar_addr)
          : "memory");
    }
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * FETCH_WARPS + warp;
  const int h = p.h, w = p.w;
  if (pix >= h * w) return;  // warp-uniform
  const int Y = pix / w, X = pix - Y * w;
  const int ntriples = PMVS_NUM_HYP * V;
  Desc* desc = reinterpret_cast<Desc*>(dyn_smem) + (size_t)warp * ntriples * 3;

  // nearest upsample of the previous depth (model.py:153-158; ATen nearest index rule)
  const float nsy = (float)p.hp / (float)h, nsx = (float)p.wp / (float)w;
  int ys = (int)floorf((float)Y * nsy), xs = (int)floorf((float)X * nsx);
  ys = ys < p.hp - 1 ? ys : p.hp - 1;
  xs = xs < p.wp - 1 ? xs : p.wp - 1;
  const float dprev = __ldg(p.depth_prev + ((size_t)b * p.hp + ys) * p.wp + xs);

  // uv = K_ref^-1 * (x + .5, y + .5, 1)   (functions.py:128-138, model.py:165-170)
  const float px = (float)X + 0.5f, py = (float)Y + 0.5f;
  const float uvx = dot3(cam + CB_KINV + 0, px, py, 1.f);
  const float uvy = dot3(cam + CB_KINV + 3, px, py, 1.f);
  const float uvz = dot3(cam + CB_KINV + 6, px, py, 1.f);
  const float interval = cam[CB_INTERVAL];

  auto world_point = [&](int m, float& wx, float& wy, float& wz) {
    const float dm = __fadd_rn(dprev, __fmul_rn(interval, (float)(m - 2)));  // model.py:174
    const float cx = __fsub_rn(__fmul_rn(uvx, dm), cam[CB_T0 + 0]);
    const float cy = __fsub_rn(__fmul_rn(uvy, dm), cam[CB_T0 + 1]);
    const float cz = __fsub_rn(__fmul_rn(uvz, dm), cam[CB_T0 + 2]);
    wx = dot3(cam + CB_R0INV + 0, cx, cy, cz);  // model.py:177
    wy = dot3(cam + CB_R0INV + 3, cx, cy, cz);
    wz = dot3(cam + CB_R0INV + 6, cx, cy, cz);
  };

  // ---- phase 1: 30 lanes build the (hypothesis, view, level) sampling descriptors ----------
  {
    const int tl = lane / 3, l = lane - tl * 3;  // triple inside the round, level
    const int hl = lane < 30 ? p.hl[l] : 1, wl = lane < 30 ? p.wl[l] : 1;
    const int Cl = 16 << l;
    const float sxl = (float)wl / (float)w, syl = (float)hl / (float)h;  // ATen area_pixel_compute_scale
    for (int t0 = 0; t0 < ntriples; t0 += FETCH_TRIPLES_PER_ROUND) {
      const int t = t0 + tl;
      const bool act = lane < 30 && t < ntriples;
      const int m = act ? t / V : 0, v = act ? t - m * V : 0;
      float wx, wy, wz;
      world_point(m, wx, wy, wz);
      const float* cv = cam + CB_VIEW + v * CB_VSTRIDE;
      float u, vv;
      project(cv, cv + 9, cv + 12, wx, wy, wz, u, vv);
      const float ix = grid_coord(u, w), iy = grid_coord(vv, h);
      const bool ok = act && usable(ix) && usable(iy);
      Axis ax, ay;
      axis_entries(ok ? ix : 0.f, w, wl, sxl, ok, ax);
      axis_entries(ok ? iy : 0.f, h, hl, syl, ok, ay);
      int nx = (ax.w[0] != 0.f) + (ax.w[1] != 0.f) + (ax.w[2] != 0.f) + (ax.w[3] != 0.f);
      int ny = (ay.w[0] != 0.f) + (ay.w[1] != 0.f) + (ay.w[2] != 0.f) + (ay.w[3] != 0.f);
      if (act) {
        Desc dd;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dd.xo[j] = (unsigned)(ax.i[j] * Cl) * 4u;
          dd.xw[j] = ax.w[j];
          dd.yo[j] = (unsigned)((ay.i[j] + v * hl) * wl * Cl) * 4u;
          dd.yw[j] = ay.w[j];
        }
        dd.nx = nx; dd.ny = ny; dd.pad0 = 0; dd.pad1 = 0;
        desc[t * 3 + l] = dd;
      }
    }
  }
  // normalised xyz of the 5 hypothesis points (model.py:46-48,193): lane m computes point m
  float* xyzs = reinterpret_cast<float*>(dyn_smem + fetch_smem_bytes(V)) + warp * 16;
  if (lane < PMVS_NUM_HYP) {
    float wx, wy, wz;
    world_point(lane, wx, wy, wz);
    xyzs[lane * 3 + 0] = __fdiv_rn(__fsub_rn(wx, cam[CB_MEAN + 0]), cam[CB_STD + 0]);
    xyzs[lane * 3 + 1] = __fdiv_rn(__fsub_rn(wy, cam[CB_MEAN + 1]), cam[CB_STD + 1]);
    xyzs[lane * 3 + 2] = __fdiv_rn(__fsub_rn(wz, cam[CB_MEAN + 2]), cam[CB_STD + 2]);
  }
  __syncwarp();

  // ---- phase 2: one pyramid level at a time, all 32 lanes on that level ----------------------
  // (see level_pass: hypotheses side by side in lane groups, views sequentially)
  const int r = p.ratio;
  const int hs = h / r, wsub = w / r;
  const int yy = Y / r, ii = Y - yy * r, xx = X / r, jj = X - xx * r;
  const int cloud = (ii * r + jj) * p.B + b;
  const int Npts = PMVS_NUM_HYP * hs * wsub;
  const float rV = __frcp_rn((float)V);
  float* frow0 = p.feature + ((size_t)cloud * Npts + (size_t)yy * wsub + xx) * PMVS_FEAT_CH;  // hypothesis 0
  const size_t fstep = (size_t)hs * wsub * PMVS_FEAT_CH;                                       // next hypothesis

  level_pass<2>(p, desc, b, V, lane, rV, frow0, fstep);
  level_pass<1>(p, desc, b, V, lane, rV, frow0, fstep);
  level_pass<0>(p, desc, b, V, lane, rV, frow0, fstep);

  // normalised xyz: tiled 8x into channels 112..135 (model.py:193-197) and kept planar for the kNN
#pragma unroll
  for (int m = 0; m < PMVS_NUM_HYP; ++m) {
    const float nx = xyzs[m * 3 + 0], ny = xyzs[m * 3 + 1], nz = xyzs[m * 3 + 2];
    if (lane < 6) {
      const int ph = lane % 3;  // float4 #q starts at component (4q) % 3 = q % 3
      float4 o;
      o.x = ph == 0 ? nx : (ph == 1 ? ny : nz);
      o.y = ph == 0 ? ny : (ph == 1 ? nz : nx);
      o.z = ph == 0 ? nz : (ph == 1 ? nx : ny);
      o.w = o.x;
      st4(frow0 + m * fstep + 112 + lane * 4, o);
    } else if (lane < 9) {
      const int comp = lane - 6;
      const int n = (m * hs + yy) * wsub + xx;
      p.xyz[((size_t)cloud * 3 + comp) * Npts + n] = comp == 0 ? nx : (comp == 1 ? ny : nz);
    }
  }
}

// ---------------------------------------------------------------------------------------
// (3) coarse-stage plane sweep: fetch + variance -> cost volume  (reference model.py:81-113)
// ---------------------------------------------------------------------------------------
// One thread per hypothesis point (d, y, x); channels in chunks of 16 so that sum / sum of squares
// stay in registers; the source-view projection is recomputed per chunk (cheap next to 64 taps).
// The reference view contributes its un-warped feature (model.py:103-106).  NCHW reads and the
// [B,C,D,h,w] writes are coalesced across x.
constexpr int CV_CH = 16;
__global__ void __launch_bounds__(256)
    cost_volume_kernel(const float* __restrict__ feat, const float* __restrict__ cam_params,
                       const float* __restrict__ cam_blocks, float* __restrict__ cost, int V, int C, int h, int w,
                       int D) {
  __shared__ float cam[cam_block_floats(PMVS_MAX_VIEWS)];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < cam_block_floats(V); i += blockDim.x)
    cam[i] = cam_blocks[(size_t)b * cam_block_floats(V) + i];
  __syncthreads();
  const int hw = h * w;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= D * hw) return;
  const int d = p / hw, pix = p - d * hw;
  const int y = pix / w, x = pix - y * w;
  // depth hypotheses: torch.linspace(depth_start, depth_end, D) (model.py:81-85; ATen's symmetric rule)
  const float* cp = cam_params + ((size_t)(b * V) * 2 + 1) * 16 + 12;
  const float dstart = cp[0], dint = cp[1];
  const float dend = __fadd_rn(dstart, __fmul_rn((float)(D - 1), dint));  // model.py:67
  const float step = D > 1 ? __fdiv_rn(__fsub_rn(dend, dstart), (float)(D - 1)) : 0.f;
  const float depth = d < D / 2 ? __fadd_rn(dstart, __fmul_rn(step, (float)d))
                                : __fsub_rn(dend, __fmul_rn(step, (float)(D - 1 - d)));
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const float cx = __fsub_rn(__fmul_rn(dot3(cam + CB_KINV + 0, px, py, 1.f), depth), cam[CB_T0 + 0]);
  const float cy = __fsub_rn(__fmul_rn(dot3(cam + CB_KINV + 3, px, py, 1.f), depth), cam[CB_T0 + 1]);
  const float cz = __fsub_rn(__fmul_rn(dot3(cam + CB_KINV + 6, px, py, 1.f), depth), cam[CB_T0 + 2]);
  const float wx = dot3(cam + CB_R0INV + 0, cx, cy, cz);
  const float wy = dot3(cam + CB_R0INV + 3, cx, cy, cz);
  const float wz = dot3(cam + CB_R0INV + 6, cx, cy, cz);
  const float fV = (float)V;
  const size_t plane = (size_t)hw;
  for (int c0 = 0; c0 < C; c0 += CV_CH) {
    float s1[CV_CH], s2[CV_CH];
    const float* ref = feat + ((size_t)(b * V) * C + c0) * plane + pix;
#pragma unroll
    for (int c = 0; c < CV_CH; ++c) {
      const float f0 = __ldg(ref + c * plane);
      s1[c] = f0;
      s2[c] = __fmul_rn(f0, f0);
    }
    for (int v = 1; v < V; ++v) {
      const float* cv = cam + CB_VIEW + v * CB_VSTRIDE;
      float u, vv;
      project(cv, cv + 9, cv + 12, wx, wy, wz, u, vv);
      const float ix = grid_coord(u, w), iy = grid_coord(vv, h);
      const bool ok = usable(ix) && usable(iy);
      const Taps tp = make_taps(ok ? ix : -10.f, ok ? iy : -10.f, w, h);
      const float* m = feat + ((size_t)(b * V + v) * C + c0) * plane + (size_t)tp.y0 * w + tp.x0;
#pragma unroll
      for (int c = 0; c < CV_CH; ++c) {
        const float* mc = m + c * plane;
        float acc = 0.f;
        if (tp.ok_n && tp.ok_w) acc = __fmul_rn(__ldg(mc), tp.nw);
        if (tp.ok_n && tp.ok_e) acc = fmaf(__ldg(mc + 1), tp.ne, acc);
        if (tp.ok_s && tp.ok_w) acc = fmaf(__ldg(mc + w), tp.sw, acc);
        if (tp.ok_s && tp.ok_e) acc = fmaf(__ldg(mc + w + 1), tp.se, acc);
        s1[c] = __fadd_rn(s1[c], acc);
        s2[c] = __fadd_rn(s2[c], __fmul_rn(acc, acc));
      }
    }
#pragma unroll
    for (int c = 0; c < CV_CH; ++c) {
      const float a = __fdiv_rn(s1[c], fV);  // model.py:108-111, unfused
      cost[(((size_t)b * C + c0 + c) * D + d) * plane + pix] = __fsub_rn(__fdiv_rn(s2[c], fV), __fmul_rn(a, a));
    }
  }
}

int launch_cam_setup(const float* cam_params, const float* interval, const float* mean, const float* stdv,
                     float* blocks, int B, int V, float kscale, float iscale, cudaStream_t st) {
  prof_begin("cam_setup", st);
  cam_setup_kernel<<<cdiv(B, 32), 32, 0, st>>>(cam_params, interval, mean, stdv, blocks, B, V, kscale, iscale);
  return check_launch("cam_setup_kernel", st);
}

int launch_fused_fetch(const FusedFetchParams& p, cudaStream_t st) {
  dim3 grid(cdiv((long long)p.h * p.w, FETCH_WARPS), p.B);
  const size_t smem = fetch_smem_total(p.V);
  static size_t smem_set = 0;  // per-process high-water mark of the opt-in dynamic smem size
  if (smem > 40 * 1024 && smem > smem_set) {  // static smem (camera block) counts against the 48 KB default
    if (cudaFuncSetAttribute(fused_fetch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      cudaGetLastError();
      set_error("fused_fetch: cannot reserve %zu bytes of shared memory (V=%d)", smem, p.V);
      return PMVS_ERR_CUDA;
    }
    smem_set = smem;
  }
  prof_begin("fused_fetch", st);
  fused_fetch_kernel<<<grid, FETCH_WARPS * 32, smem, st>>>(p);
  return check_launch("fused_fetch_kernel", st);
}

size_t cam_block_bytes(int B, int V) { return (size_t)B * cam_block_floats(V) * sizeof(float); }

}  // namespace pmvs

extern "C" int pmvs_feature_fetch(const float* feature_maps, const float* pts, const float* intrinsics,
                                  const float* extrinsics, float* out, int B, int V, int C, int H, int W, int N,
                                  pmvs_stream_t stream) {
  using namespace pmvs;
  PMVS_REQUIRE(feature_maps && pts && intrinsics && out, "feature_fetch: NULL pointer");
  PMVS_REQUIRE(B > 0 && V > 0 && C > 0 && H > 1 && W > 1 && N >= 0, "feature_fetch: bad shape");
  PMVS_REQUIRE((long long)B * V <= 65535, "feature_fetch: B*V too large");
  if (N == 0) return PMVS_OK;
  dim3 grid(cdiv(N, 256), B * V);
  feature_fetch_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(feature_maps, pts, intrinsics, extrinsics, out,
                                                                    nullptr, V, C, H, W, N);
  return check_launch("feature_fetch_kernel");
}

extern "C" int pmvs_feature_fetch_backward(const float* grad_out, const float* pts, const float* intrinsics,
                                           const float* extrinsics, float* grad_maps, int B, int V, int C, int H,
                                           int W, int N, pmvs_stream_t stream) {
  using namespace pmvs;
  PMVS_REQUIRE(grad_out && pts && intrinsics && grad_maps, "feature_fetch_backward: NULL pointer");
  PMVS_REQUIRE(B > 0 && V > 0 && C > 0 && H > 1 && W > 1 && N >= 0, "feature_fetch_backward: bad shape");
  PMVS_REQUIRE((long long)B * V <= 65535, "feature_fetch_backward: B*V too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(grad_maps, 0, (size_t)B * V * C * H * W * sizeof(float), st) != cudaSuccess) {
    set_error("feature_fetch_backward: memset failed");
    return PMVS_ERR_CUDA;
  }
  if (N == 0) return PMVS_OK;
  dim3 grid(cdiv(N, 256), B * V);
  feature_fetch_kernel<true><<<grid, 256, 0, st>>>(nullptr, pts, intrinsics, extrinsics, const_cast<float*>(grad_out),
                                                   grad_maps, V, C, H, W, N);
  return check_launch("feature_fetch_backward_kernel");
}

extern "C" int pmvs_cost_volume(const float* features, const float* cam_params, float* cost, void* workspace,
                                size_t workspace_bytes, int B, int V, int C, int h, int w, int D, int is_test,
                                pmvs_stream_t stream) {
  using namespace pmvs;
  PMVS_REQUIRE(features && cam_params && cost && workspace, "cost_volume: NULL pointer");
  PMVS_REQUIRE(B > 0 && B <= 65535 && V > 0 && V <= PMVS_MAX_VIEWS && h > 1 && w > 1 && D > 0, "cost_volume: bad shape");
  PMVS_REQUIRE(C > 0 && C % CV_CH == 0, "cost_volume: channels must be a multiple of %d", CV_CH);
  PMVS_REQUIRE((long long)D * h * w < (1ll << 31), "cost_volume: volume too large");
  if (workspace_bytes < cam_block_bytes(B, V)) {
    set_error("cost_volume: workspace %zu bytes < required %zu", workspace_bytes, cam_block_bytes(B, V));
    return PMVS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // model.py:58-61: K rows 0,1 divided by 2, and by 4 more at test time
  PMVS_TRY(launch_cam_setup(cam_params, nullptr, nullptr, nullptr, (float*)workspace, B, V, is_test ? 0.125f : 0.5f, 1.f,
                            st));
  dim3 grid(cdiv((long long)D * h * w, 256), B);
  prof_begin("cost_volume", st);
  cost_volume_kernel<<<grid, 256, 0, st>>>(features, cam_params, (const float*)workspace, cost, V, C, h, w, D);
  return check_launch("cost_volume_kernel", st);
}
