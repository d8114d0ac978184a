// EXPERIMENTAL (compiled only with -DPMVS_EDGE_TILE=1; not part of the default build, never run on a
// GPU yet - DESIGN.md section 8, item 1).
//
// EdgeConv statistics / apply with the neighbour rows gathered from SHARED memory.  The default
// edge_kernel issues 16 independent 128/256-byte L2->L1 gathers per point and sits at ~50 % of every
// unit.  Here a CTA owns a tile of ET_TX x ET_TY pixels x 5 hypothesis layers of one sub-cloud,
// loads the `edge` half of the rows of the tile plus its 2-pixel halo once (3 rows per point instead
// of 16 gathers), 32 channels at a time, and resolves neighbour k of a point through the 1-byte
// CANDIDATE id the kNN kernel emits (id = (dd+2)*25 + (dh+2)*5 + (dw+2), torch_utils.py:32-38): the
// halo row is base(point) + lut[id].  Picks outside the grid (zero-vector candidates whose linear
// index aliases another row, torch_utils.py:51-59; 0.04 % of the picks) carry id 255 and fall back to
// the global row index.  Arithmetic is the same as edge_kernel's.
#pragma once

constexpr int ET_TX = 8, ET_TY = 4, ET_D = PMVS_NUM_HYP;  // 160 points per tile
constexpr int ET_HX = ET_TX + 4, ET_HY = ET_TY + 4;       // with the 2-pixel halo of the 5x5x5 window
constexpr int ET_ROWS = ET_D * ET_HY * ET_HX;             // 480 halo rows
constexpr int ET_CP = 32;                                 // channels per slab (61 440 B of shared memory)
constexpr int ET_THREADS = 256;
constexpr int ET_LPP = ET_CP / 4;                         // lanes per point
constexpr int ET_PPW = 32 / ET_LPP;                       // points per warp step
constexpr int ET_WARPS = ET_THREADS / 32;
constexpr size_t ET_SMEM = (size_t)ET_ROWS * ET_CP * sizeof(float);

template <int COUT, bool APPLY>
__global__ void __launch_bounds__(ET_THREADS) edge_tile_kernel(const EdgeArgs a) {
  static_assert(COUT % ET_CP == 0, "edge_tile: channels must be a multiple of 32");
  constexpr int LD = 2 * COUT;
  extern __shared__ __align__(16) float halo[];  // [ET_ROWS][ET_CP]
  __shared__ int lut[125];
  __shared__ float part[APPLY ? 1 : ET_WARPS][APPLY ? 1 : 4 * ET_CP];
  __shared__ float c_mean[2][COUT], c_istd[2][COUT], c_g[2][COUT], c_b[2][COUT];

  const int g = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane / ET_LPP, cl = (lane % ET_LPP) * 4;
  const int gh = a.gh, gw = a.gw;
  const int tiles_x = (gw + ET_TX - 1) / ET_TX;
  const int y0 = (blockIdx.x / tiles_x) * ET_TY, x0 = (blockIdx.x % tiles_x) * ET_TX;
  const size_t cloud_base = (size_t)g * a.rows_per_group + (size_t)blockIdx.y * a.N;

  for (int c = tid; c < 125; c += ET_THREADS)
    lut[c] = ((c / 25 - 2) * ET_HY + ((c % 25) / 5 - 2)) * ET_HX + (c % 5 - 2);
  if (APPLY) {
    const double* s = a.stats + (size_t)g * 4 * COUT;
    const double cnt_c = (double)a.rows_per_group, cnt_n = (double)a.rows_per_group * 16;
    for (int c = tid; c < COUT; c += ET_THREADS) {
      BnCoef kn = bn_coef(s[2 * COUT + c], s[3 * COUT + c], cnt_n, a.eps);
      const int gn = a.concat_central ? COUT + c : c;
      c_mean[1][c] = kn.mean; c_istd[1][c] = kn.invstd; c_g[1][c] = a.gamma[gn]; c_b[1][c] = a.beta[gn];
      if (a.concat_central) {
        BnCoef kc = bn_coef(s[c], s[COUT + c], cnt_c, a.eps);
        c_mean[0][c] = kc.mean; c_istd[0][c] = kc.invstd; c_g[0][c] = a.gamma[c]; c_b[0][c] = a.beta[c];
      }
    }
  }

#pragma unroll 1
  for (int slab = 0; slab < COUT / ET_CP; ++slab) {
    const int ch0 = slab * ET_CP;
    __syncthreads();  // previous slab's readers are done (and lut / coefficients are visible)
    // ---- halo of `edge` rows, channels [ch0, ch0+32): one float4 per thread and step ----------------
    for (int e = tid; e < ET_ROWS * ET_LPP; e += ET_THREADS) {
      const int r = e / ET_LPP, q = e - r * ET_LPP;
      const int hx = r % ET_HX, t = r / ET_HX;
      const int hy = t % ET_HY, d = t / ET_HY;
      const int gy = y0 + hy - 2, gx = x0 + hx - 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < gh && gx >= 0 && gx < gw)
        v = ldg4(a.le + (cloud_base + (size_t)(d * gh + gy) * gw + gx) * LD + COUT + ch0 + q * 4);
      *reinterpret_cast<float4*>(halo + (size_t)r * ET_CP + q * 4) = v;
    }
    __syncthreads();

    float4 sc1 = make_float4(0.f, 0.f, 0.f, 0.f), sc2 = sc1, sn1 = sc1, sn2 = sc1;
    for (int it = warp * ET_PPW + sub; it < ET_D * ET_TY * ET_TX; it += ET_WARPS * ET_PPW) {
      const int tx = it % ET_TX, t = it / ET_TX;
      const int ty = t % ET_TY, d = t / ET_TY;
      const int y = y0 + ty, x = x0 + tx;
      if (y >= gh || x >= gw) continue;
      const size_t row = cloud_base + (size_t)(d * gh + y) * gw + x;
      const float4 loc = ldg4(a.le + row * LD + ch0 + cl);
      const int4 cw = __ldg(reinterpret_cast<const int4*>(a.cand + row * 16));
      const bool esc = ((cw.x | cw.y | cw.z | cw.w) & 0x80808080) != 0;  // ids are <= 124 unless escaped
      const int base_p = (d * ET_HY + ty + 2) * ET_HX + tx + 2;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 A4, c0;
      if (APPLY) {
        const float4 m = *reinterpret_cast<const float4*>(&c_mean[1][ch0 + cl]);
        const float4 is = *reinterpret_cast<const float4*>(&c_istd[1][ch0 + cl]);
        const float4 gm = *reinterpret_cast<const float4*>(&c_g[1][ch0 + cl]);
        const float4 bt = *reinterpret_cast<const float4*>(&c_b[1][ch0 + cl]);
        A4 = make_float4(is.x * gm.x, is.y * gm.y, is.z * gm.z, is.w * gm.w);
        c0 = make_float4(fmaf(-(m.x + loc.x), A4.x, bt.x), fmaf(-(m.y + loc.y), A4.y, bt.y),
                         fmaf(-(m.z + loc.z), A4.z, bt.z), fmaf(-(m.w + loc.w), A4.w, bt.w));
      }
      const int words[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int c = (words[k >> 2] >> (8 * (k & 3))) & 255;
        float4 e;
        if (esc && c == 255) {
          const int nb = __ldg(a.idx + row * 16 + k);
          e = ldg4(a.le + (cloud_base + (size_t)nb) * LD + COUT + ch0 + cl);
        } else {
          e = *reinterpret_cast<const float4*>(halo + (size_t)(base_p + lut[c]) * ET_CP + cl);
        }
        if (APPLY) {
          o.x += fmaxf(fmaf(e.x, A4.x, c0.x), 0.f);
          o.y += fmaxf(fmaf(e.y, A4.y, c0.y), 0.f);
          o.z += fmaxf(fmaf(e.z, A4.z, c0.z), 0.f);
          o.w += fmaxf(fmaf(e.w, A4.w, c0.w), 0.f);
        } else {
          const float dx = __fsub_rn(e.x, loc.x), dy = __fsub_rn(e.y, loc.y);
          const float dz = __fsub_rn(e.z, loc.z), dw = __fsub_rn(e.w, loc.w);
          sn1.x += dx; sn1.y += dy; sn1.z += dz; sn1.w += dw;
          sn2.x = fmaf(dx, dx, sn2.x); sn2.y = fmaf(dy, dy, sn2.y);
          sn2.z = fmaf(dz, dz, sn2.z); sn2.w = fmaf(dw, dw, sn2.w);
        }
      }
      if (APPLY) {
        float* orow = a.out + row * a.ldo;
        if (a.concat_central) {
          float4 c;
          c.x = fmaxf(bn_apply(loc.x, c_mean[0][ch0 + cl + 0], c_istd[0][ch0 + cl + 0], c_g[0][ch0 + cl + 0], c_b[0][ch0 + cl + 0]), 0.f);
          c.y = fmaxf(bn_apply(loc.y, c_mean[0][ch0 + cl + 1], c_istd[0][ch0 + cl + 1], c_g[0][ch0 + cl + 1], c_b[0][ch0 + cl + 1]), 0.f);
          c.z = fmaxf(bn_apply(loc.z, c_mean[0][ch0 + cl + 2], c_istd[0][ch0 + cl + 2], c_g[0][ch0 + cl + 2], c_b[0][ch0 + cl + 2]), 0.f);
          c.w = fmaxf(bn_apply(loc.w, c_mean[0][ch0 + cl + 3], c_istd[0][ch0 + cl + 3], c_g[0][ch0 + cl + 3], c_b[0][ch0 + cl + 3]), 0.f);
          st4(orow + ch0 + cl, c);
          orow += COUT;
        }
        st4(orow + ch0 + cl, make_float4(__fdiv_rn(o.x, 16.f), __fdiv_rn(o.y, 16.f), __fdiv_rn(o.z, 16.f), __fdiv_rn(o.w, 16.f)));
      } else {
        sc1.x += loc.x; sc1.y += loc.y; sc1.z += loc.z; sc1.w += loc.w;
        sc2.x = fmaf(loc.x, loc.x, sc2.x); sc2.y = fmaf(loc.y, loc.y, sc2.y);
        sc2.z = fmaf(loc.z, loc.z, sc2.z); sc2.w = fmaf(loc.w, loc.w, sc2.w);
      }
    }

    if (!APPLY) {
      float v[16] = {sc1.x, sc1.y, sc1.z, sc1.w, sc2.x, sc2.y, sc2.z, sc2.w,
                     sn1.x, sn1.y, sn1.z, sn1.w, sn2.x, sn2.y, sn2.z, sn2.w};
#pragma unroll
      for (int off = ET_LPP; off < 32; off <<= 1) {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
      }
      if (sub == 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) part[warp][(q >> 2) * ET_CP + cl + (q & 3)] = v[q];
      }
      __syncthreads();
      double* o = a.stats + (size_t)g * 4 * COUT;
      for (int c = tid; c < 4 * ET_CP; c += ET_THREADS) {
        double t = 0.0;
#pragma unroll
        for (int wq = 0; wq < ET_WARPS; ++wq) t += (double)part[wq][c];
        atomicAdd(o + (c / ET_CP) * COUT + ch0 + (c % ET_CP), t);  // [sum_c | sumsq_c | sum_n | sumsq_n][COUT]
      }
    }
  }
}

template <bool APPLY>
static int launch_edge_tile(const EdgeArgs& a, cudaStream_t st) {
  PMVS_REQUIRE(a.K == 16 && (a.cout == 32 || a.cout == 64) && a.cand != nullptr, "edge_tile: unsupported shape");
  PMVS_REQUIRE(a.N == PMVS_NUM_HYP * a.gh * a.gw && a.rows_per_group % a.N == 0, "edge_tile: bad cloud shape");
  const int clouds = a.rows_per_group / a.N;
  PMVS_REQUIRE(a.groups <= 65535 && clouds <= 65535, "edge_tile: too many clouds");
  dim3 grid(cdiv(a.gw, ET_TX) * cdiv(a.gh, ET_TY), clouds, a.groups);
  static const char* const names[2][2] = {{"edge_stats_32", "edge_stats_64"}, {"edge_apply_32", "edge_apply_64"}};
  static bool attr_set[2][2] = {{false, false}, {false, false}};
  const int ci = a.cout == 32 ? 0 : 1;
  auto kern = a.cout == 32 ? edge_tile_kernel<32, APPLY> : edge_tile_kernel<64, APPLY>;
  if (!attr_set[APPLY ? 1 : 0][ci]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ET_SMEM) != cudaSuccess) {
      cudaGetLastError();
      set_error("edge_tile: cannot reserve %zu bytes of shared memory", ET_SMEM);
      return PMVS_ERR_CUDA;
    }
    attr_set[APPLY ? 1 : 0][ci] = true;
  }
  prof_begin(names[APPLY ? 1 : 0][ci], st);
  kern<<<grid, ET_THREADS, ET_SMEM, st>>>(a);
  return check_launch(APPLY ? "edge_tile_apply_kernel" : "edge_tile_stats_kernel", st);
}
