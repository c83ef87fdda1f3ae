// Mask refinement on the device (reference: manga_translator/mask_refinement/__init__.py:9-31 and
// text_mask_utils.py:64-190).  The reference runs this stage on the CPU between OCR and inpainting: cv2.resize of page and raw mask,
// connected components of the raw mask, a per-text-line DenseCRF (pydensecrf: 5 mean-field iterations, a spatial and a bilateral
// permutohedral-lattice filter per iteration), elliptical dilations, and a resize back.  Kernels here:
//   resize_linear_u8      cv2.resize(INTER_LINEAR) for uint8, bit-exact (fixed-point scheme of OpenCV's resize.cpp: 11-bit coefficients,
//                         horizontal pass in int, vertical ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2; x clamps its
//                         coefficient at the border, y clamps the row index), optional "> 0 -> 255"
//   cut_rects             cv2.rectangle(mask, (x, y), (x+w, y+h), 0, 1) for every text line (separates components at line boxes)
//   cc_*                  8-connected component labelling by union-find (init / merge with the W, NW, N, NE neighbours / flatten),
//                         component ids + bounding boxes + areas (the `stats` of cv2.connectedComponentsWithStats)
//   crf_*                 the batched DenseCRF: every text line's region is one segment of a shared pixel list and owns one segment
//                         of a shared hash table; lattice construction (elevate, round to the remainder-0 point, rank, barycentric
//                         weights, key insertion with atomicCAS), blur-neighbour lookup, then per iteration splat (atomicAdd) ->
//                         blur along the d+1 axes -> slice for both kernels and the softmax update.  Arithmetic follows densecrf's
//                         permutohedral.cpp statement by statement in fp32; the only freedom taken is the summation ORDER of the
//                         splat (atomics), which perturbs Q in the last bits.
//   dilate_lines / dilate_se / owner_map   the per-line ellipse dilation + OR into the page mask, the final dilation
#include <cuda_runtime.h>
#include <math.h>
#include "mitb_internal.h"

namespace mitb {

namespace {

// ------------------------------------------------------------------------------------------------------------------- resize
__device__ __forceinline__ void lin_coeff(int d, int dn, int sn, bool clamp_coeff, int& i0, int& i1, int& a0, int& a1) {
  const double scale = 1.0 / ((double)dn / (double)sn);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int i = (int)floorf(f);
  f -= (float)i;
  if (clamp_coeff) {
    if (i < 0) { i = 0; f = 0.f; }
    if (i >= sn - 1) { i = sn - 1; f = 0.f; }
  }
  i0 = i < 0 ? 0 : i > sn - 1 ? sn - 1 : i;
  i1 = i + 1 < 0 ? 0 : i + 1 > sn - 1 ? sn - 1 : i + 1;
  a0 = __float2int_rn((1.f - f) * 2048.f);
  a1 = __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, int cn, uint8_t* __restrict__ dst,
                                                               int dh, int dw, int binarize) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= dw || dy >= dh) return;
  int x0, x1, xa0, xa1, y0, y1, ya0, ya1;
  lin_coeff(dx, dw, sw, true, x0, x1, xa0, xa1);
  lin_coeff(dy, dh, sh, false, y0, y1, ya0, ya1);
  const uint8_t* r0 = src + (size_t)y0 * sw * cn;
  const uint8_t* r1 = src + (size_t)y1 * sw * cn;
  uint8_t* o = dst + ((size_t)dy * dw + dx) * cn;
  for (int k = 0; k < cn; ++k) {
    const int s0 = r0[x0 * cn + k] * xa0 + r0[x1 * cn + k] * xa1;
    const int s1 = r1[x0 * cn + k] * xa0 + r1[x1 * cn + k] * xa1;
    int v = (((ya0 * (s0 >> 4)) >> 16) + ((ya1 * (s1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : v > 255 ? 255 : v;
    o[k] = (uint8_t)(binarize ? (v > 0 ? 255 : 0) : v);
  }
}

// ------------------------------------------------------------------------------------------------------------------- rectangles
__global__ void __launch_bounds__(128) cut_rects_kernel(uint8_t* __restrict__ mask, int h, int w, const int* __restrict__ rects, int n) {
  const int r = blockIdx.x;
  if (r >= n) return;
  const int x0 = rects[4 * r], y0 = rects[4 * r + 1], x1 = x0 + rects[4 * r + 2], y1 = y0 + rects[4 * r + 3];
  for (int x = x0 + (int)threadIdx.x; x <= x1; x += blockDim.x) {
    if (x < 0 || x >= w) continue;
    if (y0 >= 0 && y0 < h) mask[(size_t)y0 * w + x] = 0;
    if (y1 >= 0 && y1 < h) mask[(size_t)y1 * w + x] = 0;
  }
  for (int y = y0 + (int)threadIdx.x; y <= y1; y += blockDim.x) {
    if (y < 0 || y >= h) continue;
    if (x0 >= 0 && x0 < w) mask[(size_t)y * w + x0] = 0;
    if (x1 >= 0 && x1 < w) mask[(size_t)y * w + x1] = 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------- components
__device__ __forceinline__ int uf_find(const int* parent, int i) {
  int p = parent[i];
  while (p != i) { i = p; p = parent[i]; }
  return i;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }            // a > b: hook the larger root under the smaller
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;                                                 // somebody else hooked a meanwhile: merge their target with b
  }
}
__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = mask[i] ? i : -1;
}
__global__ void __launch_bounds__(256) cc_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int h, int w) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int i = y * w + x;
  if (!mask[i]) return;
  if (x > 0 && mask[i - 1]) uf_union(parent, i, i - 1);
  if (y > 0) {
    if (mask[i - w]) uf_union(parent, i, i - w);
    if (x > 0 && mask[i - w - 1]) uf_union(parent, i, i - w - 1);
    if (x + 1 < w && mask[i - w + 1]) uf_union(parent, i, i - w + 1);
  }
}
// flatten + number the roots (component id in `comp_of_root`, indexed by the root pixel)
__global__ void __launch_bounds__(256) cc_flatten_kernel(int* __restrict__ parent, int n, int* __restrict__ comp_of_root, int* __restrict__ ncomp, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || parent[i] < 0) return;
  const int r = uf_find(parent, i);
  parent[i] = r;
  if (r == i) { const int id = atomicAdd(ncomp, 1); comp_of_root[i] = id < cap ? id : -1; }
}
// stats[id] = {x0, y0, x1, y1, area}; labels[i] = component id (or -1)
__global__ void __launch_bounds__(256) cc_stats_kernel(const int* __restrict__ parent, const int* __restrict__ comp_of_root, int h, int w,
                                                       int* __restrict__ labels, int* __restrict__ stats) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int i = y * w + x;
  int id = -1;
  if (parent[i] >= 0) {
    id = comp_of_root[parent[i]];
    if (id >= 0) {
      int* s = stats + 5 * id;
      atomicMin(s + 0, x); atomicMin(s + 1, y); atomicMax(s + 2, x); atomicMax(s + 3, y); atomicAdd(s + 4, 1);
    }
  }
  labels[i] = id;
}
__global__ void __launch_bounds__(256) cc_stats_init_kernel(int* __restrict__ stats, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { stats[5 * i] = 0x7fffffff; stats[5 * i + 1] = 0x7fffffff; stats[5 * i + 2] = -1; stats[5 * i + 3] = -1; stats[5 * i + 4] = 0; }
}
// owner_map[i] = owner[labels[i]] (text line that owns the pixel's component) or -1
__global__ void __launch_bounds__(256) owner_map_kernel(const int* __restrict__ labels, const int* __restrict__ owner, int n, int* __restrict__ omap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int l = labels[i]; omap[i] = l >= 0 ? owner[l] : -1; }
}

// ------------------------------------------------------------------------------------------------------------------- dense CRF
// Per line: region rect (x, y, w, h) in the working image, first pixel / first table slot of its segments, table capacity (power of two)
struct CrfLine { int x, y, w, h; int pix0; int slot0; int cap; int pad; };

constexpr unsigned long long kEmpty = ~0ull;

template <int D>
__device__ __forceinline__ unsigned long long pack_key(const int (&k)[D], int* err) {
  constexpr int BITS = D <= 3 ? 16 : 12;
  constexpr int BIAS = 1 << (BITS - 1);
  unsigned long long key = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int v = k[i] + BIAS;
    if (v < 0 || v >= (1 << BITS)) *err = 1;
    key |= (unsigned long long)(v & ((1 << BITS) - 1)) << (BITS * i);
  }
  return key;
}
__device__ __forceinline__ unsigned hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int table_insert(unsigned long long* keys, int slot0, int cap, unsigned long long key) {
  unsigned s = hash64(key) & (unsigned)(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    const unsigned long long cur = atomicCAS(&keys[slot0 + s], kEmpty, key);
    if (cur == kEmpty || cur == key) return slot0 + (int)s;
    s = (s + 1) & (unsigned)(cap - 1);
  }
  return -1;
}
__device__ __forceinline__ int table_find(const unsigned long long* keys, int slot0, int cap, unsigned long long key) {
  unsigned s = hash64(key) & (unsigned)(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    const unsigned long long cur = keys[slot0 + s];
    if (cur == key) return slot0 + (int)s;
    if (cur == kEmpty) return -1;
    s = (s + 1) & (unsigned)(cap - 1);
  }
  return -1;
}

// Lattice construction for one kernel (D = 2: positions / sxy; D = 5: positions / sxy and colours / srgb), densecrf Permutohedral::init.
// One thread per region pixel; grid.y = line.
template <int D>
__global__ void __launch_bounds__(128) crf_lattice_kernel(const CrfLine* __restrict__ lines, const uint8_t* __restrict__ img, int img_w,
                                                          float sxy, float srgb, unsigned long long* __restrict__ keys,
                                                          int* __restrict__ offs, float* __restrict__ bary, int* __restrict__ err) {
  const CrfLine L = lines[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.w * L.h) return;
  const int lx = p % L.w, ly = p / L.w;
  float f[D];
  f[0] = (float)lx / sxy; f[1] = (float)ly / sxy;
  if (D == 5) {
    const uint8_t* c = img + ((size_t)(L.y + ly) * img_w + (L.x + lx)) * 3;
    f[2] = (float)c[0] / srgb; f[3] = (float)c[1] / srgb; f[4] = (float)c[2] / srgb;
  }
  const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)(D + 1));
  float elevated[D + 1], rem0[D + 1];
  int rank[D + 1];
  {
    float sm = 0.f;
#pragma unroll
    for (int j = D; j > 0; --j) {
      const float scale = (float)(1.0 / sqrt((double)((j + 1) * j)) * (double)inv_std_dev);
      const float cf = __fmul_rn(f[j - 1], scale);
      elevated[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
      sm = __fadd_rn(sm, cf);
    }
    elevated[0] = sm;
  }
  const float down = 1.0f / (float)(D + 1), up = (float)(D + 1);
  int sum = 0;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(down, elevated[i]);
    const float u = __fmul_rn(ceilf(v), up), dn = __fmul_rn(floorf(v), up);
    const float rd = (__fsub_rn(u, elevated[i]) < __fsub_rn(elevated[i], dn)) ? (float)(short)u : (float)(short)dn;
    rem0[i] = rd;
    sum = (int)__fadd_rn((float)sum, __fmul_rn(rd, down));
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) rank[i] = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float di = __fsub_rn(elevated[i], rem0[i]);
#pragma unroll
    for (int j = i + 1; j <= D; ++j) {
      if (di < __fsub_rn(elevated[j], rem0[j])) rank[i]++; else rank[j]++;
    }
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    rank[i] += sum;
    if (rank[i] < 0) { rank[i] += D + 1; rem0[i] += (float)(D + 1); }
    else if (rank[i] > D) { rank[i] -= D + 1; rem0[i] -= (float)(D + 1); }
  }
  float b[D + 2];
#pragma unroll
  for (int i = 0; i <= D + 1; ++i) b[i] = 0.f;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(__fsub_rn(elevated[i], rem0[i]), down);
    // b[D - rank[i]] += v; b[D - rank[i] + 1] -= v  (rank is data dependent: select instead of indexing, keeps b in registers)
#pragma unroll
    for (int t = 0; t <= D + 1; ++t) {
      if (t == D - rank[i]) b[t] = __fadd_rn(b[t], v);
      if (t == D - rank[i] + 1) b[t] = __fsub_rn(b[t], v);
    }
  }
  b[0] = (float)(1.0 + (double)b[D + 1] + (double)b[0]);
  const size_t o = ((size_t)L.pix0 + p) * (D + 1);
#pragma unroll
  for (int r = 0; r <= D; ++r) {
    int key[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      // canonical[r][rank[i]] = rank[i] <= D - r ? r : r - (D + 1)
      key[i] = (int)rem0[i] + (rank[i] <= D - r ? r : r - (D + 1));
    }
    const int slot = table_insert(keys, L.slot0, L.cap, pack_key<D>(key, err));
    if (slot < 0) *err = 2;
    offs[o + r] = slot;
    bary[o + r] = b[r];
  }
}

// blur neighbours of every occupied slot: nbr[(j * nslots + slot) * 2 + {0,1}]
template <int D>
__global__ void __launch_bounds__(256) crf_neighbors_kernel(const CrfLine* __restrict__ lines, const unsigned long long* __restrict__ keys,
                                                            int* __restrict__ nbr, int nslots) {
  const CrfLine L = lines[blockIdx.y];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.cap) return;
  const int slot = L.slot0 + s;
  const unsigned long long key = keys[slot];
  if (key == kEmpty) return;
  constexpr int BITS = D <= 3 ? 16 : 12;
  constexpr int BIAS = 1 << (BITS - 1);
  int k[D];
#pragma unroll
  for (int i = 0; i < D; ++i) k[i] = (int)((key >> (BITS * i)) & ((1 << BITS) - 1)) - BIAS;
  int dummy = 0;
#pragma unroll
  for (int j = 0; j <= D; ++j) {
    int n1[D], n2[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { n1[i] = k[i] - 1; n2[i] = k[i] + 1; }
    if (j < D) { n1[j] = k[j] + D; n2[j] = k[j] - D; }
    nbr[((size_t)j * nslots + slot) * 2 + 0] = table_find(keys, L.slot0, L.cap, pack_key<D>(n1, &dummy));
    nbr[((size_t)j * nslots + slot) * 2 + 1] = table_find(keys, L.slot0, L.cap, pack_key<D>(n2, &dummy));
  }
}

// Q0 = softmax(-unary): unary from the line's component mask (owner map == line): {0, u_on} / {u_on, 0}
__global__ void __launch_bounds__(128) crf_init_q_kernel(const CrfLine* __restrict__ lines, const int* __restrict__ omap, int img_w, float u_on,
                                                         float2* __restrict__ unary, float2* __restrict__ Q) {
  const CrfLine L = lines[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.w * L.h) return;
  const int lx = p % L.w, ly = p / L.w;
  const bool on = omap[(size_t)(L.y + ly) * img_w + (L.x + lx)] == (int)blockIdx.y;
  // mask_softmax = [1 - m, m] -> unary = -log(clip(., 1e-5, 1)): on: (u_on, 0), off: (0, u_on)
  const float2 u = on ? make_float2(u_on, 0.f) : make_float2(0.f, u_on);
  unary[L.pix0 + p] = u;
  const float a = -u.x, b = -u.y, m = fmaxf(a, b);
  const float ea = expf(a - m), eb = expf(b - m), s = ea + eb;
  Q[L.pix0 + p] = make_float2(ea / s, eb / s);
}

template <int D>
__global__ void __launch_bounds__(128) crf_splat_kernel(const CrfLine* __restrict__ lines, const int* __restrict__ offs, const float* __restrict__ bary,
                                                        const float2* __restrict__ Q, float2* __restrict__ vals) {
  const CrfLine L = lines[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.w * L.h) return;
  const float2 q = Q[L.pix0 + p];
  const size_t o = ((size_t)L.pix0 + p) * (D + 1);
#pragma unroll
  for (int r = 0; r <= D; ++r) {
    const int s = offs[o + r];
    const float w = bary[o + r];
    atomicAdd(&vals[s].x, __fmul_rn(w, q.x));
    atomicAdd(&vals[s].y, __fmul_rn(w, q.y));
  }
}

__global__ void __launch_bounds__(256) crf_blur_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ nbr_j, int nslots,
                                                       const float2* __restrict__ src, float2* __restrict__ dst) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  if (keys[s] == kEmpty) return;
  const int n1 = nbr_j[2 * (size_t)s], n2 = nbr_j[2 * (size_t)s + 1];
  const float2 a = n1 >= 0 ? src[n1] : make_float2(0.f, 0.f), b = n2 >= 0 ? src[n2] : make_float2(0.f, 0.f), c = src[s];
  dst[s] = make_float2(__fadd_rn(c.x, __fmul_rn(0.5f, __fadd_rn(a.x, b.x))), __fadd_rn(c.y, __fmul_rn(0.5f, __fadd_rn(a.y, b.y))));
}

// slice one kernel's lattice into out[pix] (float2 per pixel)
template <int D>
__global__ void __launch_bounds__(128) crf_slice_kernel(const CrfLine* __restrict__ lines, const int* __restrict__ offs, const float* __restrict__ bary,
                                                        const float2* __restrict__ vals, float2* __restrict__ out) {
  const CrfLine L = lines[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.w * L.h) return;
  const float alpha = 1.0f / (1.0f + exp2f(-(float)D));
  const size_t o = ((size_t)L.pix0 + p) * (D + 1);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r <= D; ++r) {
    const float2 v = vals[offs[o + r]];
    const float w = bary[o + r];
    acc.x = __fadd_rn(acc.x, __fmul_rn(__fmul_rn(w, v.x), alpha));
    acc.y = __fadd_rn(acc.y, __fmul_rn(__fmul_rn(w, v.y), alpha));
  }
  out[L.pix0 + p] = acc;
}

// Q = softmax(-unary + wg * Kg + wb * Kb); on the last iteration also writes the refined mask (argmax) into `refined`
__global__ void __launch_bounds__(256) crf_update_kernel(const float2* __restrict__ unary, const float2* __restrict__ kg, const float2* __restrict__ kb,
                                                         float wg, float wb, int npix, float2* __restrict__ Q, uint8_t* __restrict__ refined) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float2 u = unary[i], g = kg[i], b = kb[i];
  // tmp1 = -unary; tmp1 -= (-wg * Kg); tmp1 -= (-wb * Kb)
  const float a0 = __fsub_rn(__fsub_rn(-u.x, __fmul_rn(-wg, g.x)), __fmul_rn(-wb, b.x));
  const float a1 = __fsub_rn(__fsub_rn(-u.y, __fmul_rn(-wg, g.y)), __fmul_rn(-wb, b.y));
  const float m = fmaxf(a0, a1);
  const float e0 = expf(a0 - m), e1 = expf(a1 - m), s = e0 + e1;
  const float2 q = make_float2(e0 / s, e1 / s);
  Q[i] = q;
  if (refined) refined[i] = q.y > q.x ? 255 : 0;               // np.argmax: ties -> label 0
}

// ------------------------------------------------------------------------------------------------------------------- dilation
// Per text line i: cc_i = refined mask inside rect1, (owner map == i) elsewhere; dilate by the line's ellipse inside rect2 (pixels
// outside rect2 do not exist for cv2.dilate of the sub-image) and OR into `final`.
struct DilLine { int x1, y1, w1, h1; int x2, y2, w2, h2; int pix0; int se0, ksize; int pad; };

__global__ void __launch_bounds__(128) dilate_lines_kernel(const DilLine* __restrict__ lines, const int* __restrict__ omap, const uint8_t* __restrict__ refined,
                                                           const uint8_t* __restrict__ se, int img_w, uint8_t* __restrict__ final_mask) {
  const DilLine L = lines[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.w2 * L.h2) return;
  const int x = L.x2 + p % L.w2, y = L.y2 + p / L.w2;
  const int a = L.ksize / 2;
  bool hit = false;
  for (int ky = 0; ky < L.ksize && !hit; ++ky) {
    const int yy = y + ky - a;
    if (yy < L.y2 || yy >= L.y2 + L.h2) continue;
    for (int kx = 0; kx < L.ksize; ++kx) {
      if (!se[L.se0 + ky * L.ksize + kx]) continue;
      const int xx = x + kx - a;
      if (xx < L.x2 || xx >= L.x2 + L.w2) continue;
      bool on;
      if (xx >= L.x1 && xx < L.x1 + L.w1 && yy >= L.y1 && yy < L.y1 + L.h1) on = refined[L.pix0 + (yy - L.y1) * L.w1 + (xx - L.x1)] != 0;
      else on = omap[(size_t)yy * img_w + xx] == (int)blockIdx.y;
      if (on) { hit = true; break; }
    }
  }
  if (hit) final_mask[(size_t)y * img_w + x] = 255;
}

// cv2.dilate(src, se) of a whole uint8 image with an arbitrary structuring element (anchor at the centre, border ignored)
__global__ void __launch_bounds__(256) dilate_se_kernel(const uint8_t* __restrict__ src, int h, int w, const uint8_t* __restrict__ se, int ksize,
                                                        uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int a = ksize / 2;
  int best = 0;
  for (int ky = 0; ky < ksize; ++ky) {
    const int yy = y + ky - a;
    if (yy < 0 || yy >= h) continue;
    for (int kx = 0; kx < ksize; ++kx) {
      const int xx = x + kx - a;
      if (xx < 0 || xx >= w || !se[ky * ksize + kx]) continue;
      const int v = src[(size_t)yy * w + xx];
      best = v > best ? v : best;
    }
  }
  dst[(size_t)y * w + x] = (uint8_t)best;
}

}  // namespace

void launch_resize_linear_u8(const uint8_t* src, int sh, int sw, int cn, uint8_t* dst, int dh, int dw, int binarize, cudaStream_t st) {
  MITB_CHECK(sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && dh <= 65535 && (cn == 1 || cn == 3), "resize_linear_u8: bad shape");
  dim3 grid((unsigned)((dw + 255) / 256), (unsigned)dh);
  resize_linear_u8_kernel<<<grid, 256, 0, st>>>(src, sh, sw, cn, dst, dh, dw, binarize);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

void launch_cut_rects(uint8_t* mask, int h, int w, const int* rects, int n, cudaStream_t st) {
  if (n <= 0) return;
  cut_rects_kernel<<<n, 128, 0, st>>>(mask, h, w, rects, n);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

// labels int32 [h*w] (component id or -1), stats int32 [cap][5] = {x0, y0, x1, y1, area}, ncomp int32 [1]; scratch: 2 * h * w ints
void launch_cc_label(const uint8_t* mask, int h, int w, int* labels, int* stats, int* ncomp, int cap, int* scratch, cudaStream_t st) {
  MITB_CHECK(h >= 1 && w >= 1 && h <= 65535 && (long)h * w < (1l << 30) && cap >= 1, "cc_label: bad shape");
  const int n = h * w;
  int* parent = scratch;
  int* comp_of_root = scratch + n;
  CUDA_OK(cudaMemsetAsync(ncomp, 0, sizeof(int), st));
  cc_stats_init_kernel<<<(cap + 255) / 256, 256, 0, st>>>(stats, cap);
  cc_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(mask, parent, n);
  dim3 grid((unsigned)((w + 255) / 256), (unsigned)h);
  cc_merge_kernel<<<grid, 256, 0, st>>>(mask, parent, h, w);
  cc_flatten_kernel<<<(n + 255) / 256, 256, 0, st>>>(parent, n, comp_of_root, ncomp, cap);
  cc_stats_kernel<<<grid, 256, 0, st>>>(parent, comp_of_root, h, w, labels, stats);
  for (int i = 0; i < 5; ++i) count_launch();
  CUDA_OK(cudaGetLastError());
}

void launch_owner_map(const int* labels, const int* owner, int n, int* omap, cudaStream_t st) {
  owner_map_kernel<<<(n + 255) / 256, 256, 0, st>>>(labels, owner, n, omap);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

// The batched DenseCRF of refine_mask (text_mask_utils.py:71-94) for `nlines` regions.
//   lines: device CrfLine[nlines] (8 ints each); img: bilateral-filtered working image uint8 [h, w, 3]; omap: owner map int32 [h, w]
//   max_pix: the largest region's pixel count; npix / nslots: totals of the pixel / slot segments
//   work (device scratch, sized by mitb_op_crf_workspace): keys | offs2 bary2 nbr2 | offs5 bary5 nbr5 | vals a/b | unary Q kg kb
//   refined: uint8 [npix] (255 where the CRF labels the pixel as text); err: int32 device flag (key range / table overflow)
size_t crf_workspace_bytes(long npix, long nslots2, long nslots5) {
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 255) & ~(size_t)255; };
  add(8 * (size_t)nslots2); add(8 * (size_t)nslots5);
  add(4 * (size_t)npix * 3); add(4 * (size_t)npix * 3); add(4 * (size_t)nslots2 * 3 * 2);
  add(4 * (size_t)npix * 6); add(4 * (size_t)npix * 6); add(4 * (size_t)nslots5 * 6 * 2);
  const size_t ns = (size_t)(nslots2 > nslots5 ? nslots2 : nslots5);
  add(8 * ns); add(8 * ns);
  add(8 * (size_t)npix); add(8 * (size_t)npix); add(8 * (size_t)npix); add(8 * (size_t)npix);
  return b + 1024;
}

void launch_crf(const int* lines2 /*CrfLine with the d=2 table segments*/, const int* lines5, int nlines, const uint8_t* img, const int* omap, int img_w,
                int max_pix, int max_cap2, int max_cap5, long npix, long nslots2, long nslots5, int iters, float sxy_g, float w_g, float sxy_b,
                float srgb, float w_b, float u_on, void* work, uint8_t* refined, int* err, cudaStream_t st) {
  MITB_CHECK(nlines >= 1 && nlines <= 65535 && npix >= 1 && npix < (1l << 28) && nslots2 < (1l << 30) && nslots5 < (1l << 30) && iters >= 1, "crf: bad sizes");
  char* base = static_cast<char*>(work);
  auto take = [&](size_t bytes) { char* p = base; base += (bytes + 255) & ~(size_t)255; return p; };
  auto* keys2 = reinterpret_cast<unsigned long long*>(take(8 * (size_t)nslots2));
  auto* keys5 = reinterpret_cast<unsigned long long*>(take(8 * (size_t)nslots5));
  int* offs2 = reinterpret_cast<int*>(take(4 * (size_t)npix * 3)); float* bary2 = reinterpret_cast<float*>(take(4 * (size_t)npix * 3));
  int* nbr2 = reinterpret_cast<int*>(take(4 * (size_t)nslots2 * 3 * 2));
  int* offs5 = reinterpret_cast<int*>(take(4 * (size_t)npix * 6)); float* bary5 = reinterpret_cast<float*>(take(4 * (size_t)npix * 6));
  int* nbr5 = reinterpret_cast<int*>(take(4 * (size_t)nslots5 * 6 * 2));
  const size_t ns = (size_t)(nslots2 > nslots5 ? nslots2 : nslots5);
  float2* va = reinterpret_cast<float2*>(take(8 * ns)); float2* vb = reinterpret_cast<float2*>(take(8 * ns));
  float2* unary = reinterpret_cast<float2*>(take(8 * (size_t)npix)); float2* Q = reinterpret_cast<float2*>(take(8 * (size_t)npix));
  float2* kg = reinterpret_cast<float2*>(take(8 * (size_t)npix)); float2* kb = reinterpret_cast<float2*>(take(8 * (size_t)npix));
  const CrfLine* L2 = reinterpret_cast<const CrfLine*>(lines2);
  const CrfLine* L5 = reinterpret_cast<const CrfLine*>(lines5);
  ProfScope ps("mask_crf", 0.0, 0.0, st);
  CUDA_OK(cudaMemsetAsync(keys2, 0xff, 8 * (size_t)nslots2, st));
  CUDA_OK(cudaMemsetAsync(keys5, 0xff, 8 * (size_t)nslots5, st));
  CUDA_OK(cudaMemsetAsync(err, 0, sizeof(int), st));
  const dim3 gp((unsigned)((max_pix + 127) / 128), (unsigned)nlines);
  crf_lattice_kernel<2><<<gp, 128, 0, st>>>(L2, img, img_w, sxy_g, 1.f, keys2, offs2, bary2, err);
  crf_lattice_kernel<5><<<gp, 128, 0, st>>>(L5, img, img_w, sxy_b, srgb, keys5, offs5, bary5, err);
  crf_neighbors_kernel<2><<<dim3((unsigned)((max_cap2 + 255) / 256), (unsigned)nlines), 256, 0, st>>>(L2, keys2, nbr2, (int)nslots2);
  crf_neighbors_kernel<5><<<dim3((unsigned)((max_cap5 + 255) / 256), (unsigned)nlines), 256, 0, st>>>(L5, keys5, nbr5, (int)nslots5);
  crf_init_q_kernel<<<gp, 128, 0, st>>>(L2, omap, img_w, u_on, unary, Q);
  for (int i = 0; i < 5; ++i) count_launch();
  for (int it = 0; it < iters; ++it) {
    // spatial kernel
    CUDA_OK(cudaMemsetAsync(va, 0, 8 * (size_t)nslots2, st));
    crf_splat_kernel<2><<<gp, 128, 0, st>>>(L2, offs2, bary2, Q, va);
    float2* s = va; float2* d = vb;
    for (int j = 0; j <= 2; ++j) {
      crf_blur_kernel<<<(unsigned)((nslots2 + 255) / 256), 256, 0, st>>>(keys2, nbr2 + (size_t)j * nslots2 * 2, (int)nslots2, s, d);
      float2* t = s; s = d; d = t;
    }
    crf_slice_kernel<2><<<gp, 128, 0, st>>>(L2, offs2, bary2, s, kg);
    // bilateral kernel
    CUDA_OK(cudaMemsetAsync(va, 0, 8 * (size_t)nslots5, st));
    crf_splat_kernel<5><<<gp, 128, 0, st>>>(L5, offs5, bary5, Q, va);
    s = va; d = vb;
    for (int j = 0; j <= 5; ++j) {
      crf_blur_kernel<<<(unsigned)((nslots5 + 255) / 256), 256, 0, st>>>(keys5, nbr5 + (size_t)j * nslots5 * 2, (int)nslots5, s, d);
      float2* t = s; s = d; d = t;
    }
    crf_slice_kernel<5><<<gp, 128, 0, st>>>(L5, offs5, bary5, s, kb);
    crf_update_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(unary, kg, kb, w_g, w_b, (int)npix, Q, it + 1 == iters ? refined : nullptr);
    for (int i = 0; i < 14; ++i) count_launch();
  }
  CUDA_OK(cudaGetLastError());
}

void launch_dilate_lines(const int* lines /*DilLine, 12 ints each*/, int nlines, int max_pix2, const int* omap, const uint8_t* refined, const uint8_t* se,
                         int img_w, uint8_t* final_mask, cudaStream_t st) {
  if (nlines <= 0) return;
  MITB_CHECK(nlines <= 65535, "dilate_lines: too many lines");
  dilate_lines_kernel<<<dim3((unsigned)((max_pix2 + 127) / 128), (unsigned)nlines), 128, 0, st>>>(reinterpret_cast<const DilLine*>(lines), omap, refined, se,
                                                                                                 img_w, final_mask);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

void launch_dilate_se(const uint8_t* src, int h, int w, const uint8_t* se, int ksize, uint8_t* dst, cudaStream_t st) {
  MITB_CHECK(h >= 1 && h <= 65535 && w >= 1 && ksize >= 1 && (ksize & 1), "dilate_se: bad shape");
  dilate_se_kernel<<<dim3((unsigned)((w + 255) / 256), (unsigned)h), 256, 0, st>>>(src, h, w, se, ksize, dst);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace mitb
