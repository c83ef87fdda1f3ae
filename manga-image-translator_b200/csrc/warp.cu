// Text-line crops and CTC collapse on the device (rows O3 and O8).
//
// warp_lines_kernel replaces, for every text line of an OCR chunk, `cv2.warpPerspective(img[y1:y2, x1:x2], M, (w, h))` (+
// `cv2.rotate(.., ROTATE_90_COUNTERCLOCKWISE)` for vertical lines) of Quadrilateral.get_transformed_region
// (utils/generic.py:445-481) and the zero-padded canvas packing of Model48pxCTCOCR._infer (ocr/model_48px_ctc.py:86-92): the
// page stays on the device, the host only solves the 4-point homography.  Bit-exact with OpenCV's own arithmetic
// (third-party, restated from OpenCV 4.x modules/imgproc/src/imgwarp.cpp; oracle/warp_ref.py is the numpy restatement pinned
// against the installed cv2):
//   * WarpPerspectiveInvoker walks the destination in blocks of bw0 columns; per pixel, in doubles and in THIS order,
//       X0 = M0*xb + M1*y + M2,  W = (M6*xb + M7*y + M8) + M6*x1,  W = W ? 32/W : 0,  fX = (X0 + M0*x1) * W   (xb = block start,
//       x1 = x - xb), clamped to the int range and rounded half-to-even; integer quads put many pixels exactly on rounding
//       boundaries, so the association order is part of the result (measured: the flat formula differs in ~1e-5 of the pixels);
//   * remapBilinear with the fixed-point table of initInterTab2D: weights (32-ay)(32-ax)*32 etc. (sum 32768), except the
//     integer-aligned entry, which saturates to 32767 and gets its missing 1 added to the *bottom-right* weight: {32767,0,0,1};
//     result = (sum + 16384) >> 15, samples outside the crop are 0 (BORDER_CONSTANT).
// ctc_collapse_kernel: greedy CTC collapse of decode_ctc_top1 (model_48px_ctc.py:466-478): keep step t iff argmax[t] != 0 and
// argmax[t] != argmax[t-1]; kept steps are compacted per line together with their log-probabilities and colours.
#include <cuda_runtime.h>
#include "mitb_internal.h"

namespace mitb {

namespace {

// one record per line, 16 doubles: Minv[9] (inverse homography, row major), x1, y1 (crop origin in the page), cw, ch (crop size),
// w, h (size of the warp output BEFORE the rotation), rot (1: vertical line, rotate 90 degrees counter-clockwise)
constexpr int kWarpRec = 16;

__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }

__global__ void __launch_bounds__(256) warp_lines_kernel(const uint8_t* __restrict__ page, int H, int W, const double* __restrict__ lines,
                                                         int n, uint8_t* __restrict__ canvas, int canvas_h, int canvas_w) {
  const int line = blockIdx.z;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= canvas_w || r >= canvas_h) return;
  const double* L = lines + (size_t)line * kWarpRec;
  const int ox = (int)L[9], oy = (int)L[10], cw = (int)L[11], ch = (int)L[12], w = (int)L[13], h = (int)L[14], rot = (int)L[15];
  uint8_t* dst = canvas + (((size_t)line * canvas_h + r) * canvas_w + c) * 3;
  // canvas (r, c) -> warp output (x, y): horizontal lines are copied, vertical ones were rotated: out[i][j] = region[j][w - 1 - i]
  const int x = rot ? w - 1 - r : c, y = rot ? c : r;
  if (x < 0 || x >= w || y < 0 || y >= h || cw <= 0 || ch <= 0) { dst[0] = 0; dst[1] = 0; dst[2] = 0; return; }
  int bh0 = h < 16 ? h : 16;
  int bw0 = 1024 / bh0; if (bw0 > w) bw0 = w;
  const int xb = (x / bw0) * bw0;
  const double dxb = (double)xb, dy = (double)y, dx1 = (double)(x - xb);
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(L[0], dxb), __dmul_rn(L[1], dy)), L[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(L[3], dxb), __dmul_rn(L[4], dy)), L[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(L[6], dxb), __dmul_rn(L[7], dy)), L[8]);
  double Wd = __dadd_rn(W0, __dmul_rn(L[6], dx1));
  Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(L[0], dx1)), Wd)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(L[3], dx1)), Wd)));
  const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  const int ax = X & 31, ay = Y & 31;
  int w00, w01, w10, w11;
  if ((ax | ay) == 0) { w00 = 32767; w01 = 0; w10 = 0; w11 = 1; }
  else { w00 = (32 - ay) * (32 - ax) * 32; w01 = (32 - ay) * ax * 32; w10 = ay * (32 - ax) * 32; w11 = ay * ax * 32; }
  const bool x0ok = sx >= 0 && sx < cw, x1ok = sx + 1 >= 0 && sx + 1 < cw, y0ok = sy >= 0 && sy < ch, y1ok = sy + 1 >= 0 && sy + 1 < ch;
  const uint8_t* p00 = page + ((long)(oy + sy) * W + (ox + sx)) * 3;            // only dereferenced where the flags allow
  const uint8_t* p10 = p00 + (size_t)W * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int v00 = (x0ok && y0ok) ? p00[k] : 0, v01 = (x1ok && y0ok) ? p00[3 + k] : 0;
    const int v10 = (x0ok && y1ok) ? p10[k] : 0, v11 = (x1ok && y1ok) ? p10[3 + k] : 0;
    const int s = (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
    dst[k] = (uint8_t)(s < 0 ? 0 : s > 255 ? 255 : s);
  }
}

// one warp per line: ballot-compaction of the kept time steps
__global__ void __launch_bounds__(32) ctc_collapse_kernel(const int* __restrict__ argmax, const float* __restrict__ logprob,
                                                          const float* __restrict__ colors, int n, int T, int* __restrict__ counts,
                                                          int* __restrict__ steps, int* __restrict__ chars, float* __restrict__ lp_out,
                                                          float* __restrict__ col_out) {
  const int line = blockIdx.x, lane = threadIdx.x;
  if (line >= n) return;
  const int* a = argmax + (size_t)line * T;
  int total = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false; int ch = 0;
    if (t < T) { ch = a[t]; keep = ch != 0 && (t == 0 || a[t - 1] != ch); }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int o = total + __popc(m & ((1u << lane) - 1u));
      const size_t d = (size_t)line * T + o;
      steps[d] = t; chars[d] = ch;
      if (lp_out) lp_out[d] = logprob[(size_t)line * T + t];
      if (col_out) {
#pragma unroll
        for (int k = 0; k < 6; ++k) col_out[d * 6 + k] = colors[((size_t)line * T + t) * 6 + k];
      }
    }
    total += __popc(m);
  }
  if (lane == 0) counts[line] = total;
}

}  // namespace

void launch_warp_lines(const uint8_t* page, int H, int W, const double* lines, int n, uint8_t* canvas, int canvas_h, int canvas_w,
                       cudaStream_t st) {
  MITB_CHECK(n >= 0 && n <= 65535 && canvas_h >= 1 && canvas_h <= 65535 && canvas_w >= 1 && H >= 1 && W >= 1, "warp_lines: bad shape");
  if (n == 0) return;
  ProfScope ps("warp_lines", 0.0, 15.0 * n * canvas_h * (double)canvas_w, st);
  dim3 grid((unsigned)((canvas_w + 255) / 256), (unsigned)canvas_h, (unsigned)n);
  warp_lines_kernel<<<grid, 256, 0, st>>>(page, H, W, lines, n, canvas, canvas_h, canvas_w);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

void launch_ctc_collapse(const int* argmax, const float* logprob, const float* colors, int n, int T, int* counts, int* steps, int* chars,
                         float* lp_out, float* col_out, cudaStream_t st) {
  MITB_CHECK(n >= 0 && T >= 1, "ctc_collapse: bad shape");
  if (n == 0) return;
  ctc_collapse_kernel<<<n, 32, 0, st>>>(argmax, logprob, colors, n, T, counts, steps, chars, lp_out, col_out);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace mitb

// ---------------------------------------------------------------------------------------------------------------------
// the O(n^2) part of the text-line graph - `quadrilateral_can_merge_region` (utils/generic.py:653-698) for every pair of
// lines, as used by the OCR direction graph (ocr/common.py:12-39) and by textline_merge (textline_merge/__init__.py:110-126).
// One thread per pair (i < j).  Per line the host passes 16 doubles: corners (8), AABB x, y, w, h, font_size, aspect_ratio, angle,
// flags (bit 0: approximately axis aligned, bit 1: convex).  Arithmetic mirrors mit_b200/host/geometry.py statement by statement in
// IEEE doubles without contraction (the host port is pinned by the reference's known-answer tests); a pair with a non-convex quad
// is reported as 2 ("ask the host": its hull has fewer vertices than the quad).
namespace mitb {
namespace {

struct PairParams { double ratio, gap, tol, tol2, fs_tol, ar_tol; };

__device__ __forceinline__ double orient_d(const double* p, const double* q, const double* r) {
  return __dsub_rn(__dmul_rn(__dsub_rn(q[0], p[0]), __dsub_rn(r[1], p[1])), __dmul_rn(__dsub_rn(q[1], p[1]), __dsub_rn(r[0], p[0])));
}
// inside or on the boundary (cv2.pointPolygonTest(..) >= 0): boundary points have distance 0 to an edge anyway, so only the strict
// interior matters for the caller
__device__ bool inside_quad(const double* poly, const double* pt) {
  bool in = false;
  for (int i = 0, j = 3; i < 4; j = i++) {
    const double xi = poly[2 * i], yi = poly[2 * i + 1], xj = poly[2 * j], yj = poly[2 * j + 1];
    if ((yi > pt[1]) != (yj > pt[1]) && pt[0] < (xj - xi) * (pt[1] - yi) / (yj - yi) + xi) in = !in;
  }
  return in;
}
__device__ double pts_to_segs(const double* P, const double* S) {      // min over the 4 points of P and the 4 edges S[j] -> S[j+1]
  double best = 1e300;
  for (int j = 0; j < 4; ++j) {
    const double* s0 = S + 2 * j; const double* s1 = S + 2 * ((j + 1) & 3);
    const double abx = __dsub_rn(s1[0], s0[0]), aby = __dsub_rn(s1[1], s0[1]);
    const double den = __dadd_rn(__dmul_rn(abx, abx), __dmul_rn(aby, aby));
    for (int i = 0; i < 4; ++i) {
      const double apx = __dsub_rn(P[2 * i], s0[0]), apy = __dsub_rn(P[2 * i + 1], s0[1]);
      double t = 0.0;
      if (den != 0.0) { t = __ddiv_rn(__dadd_rn(__dmul_rn(apx, abx), __dmul_rn(apy, aby)), den); t = t < 0.0 ? 0.0 : t > 1.0 ? 1.0 : t; }
      const double ox = __dsub_rn(P[2 * i], __dadd_rn(s0[0], __dmul_rn(t, abx))), oy = __dsub_rn(P[2 * i + 1], __dadd_rn(s0[1], __dmul_rn(t, aby)));
      const double d = sqrt(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)));
      best = d < best ? d : best;
    }
  }
  return best;
}
__device__ double polygon_distance_d(const double* p1, const double* p2) {
  for (int i = 0; i < 4; ++i) {
    const double* a = p1 + 2 * i; const double* b = p1 + 2 * ((i + 1) & 3);
    for (int j = 0; j < 4; ++j) {
      const double* c = p2 + 2 * j; const double* d = p2 + 2 * ((j + 1) & 3);
      if (__dmul_rn(orient_d(a, b, c), orient_d(a, b, d)) < 0.0 && __dmul_rn(orient_d(c, d, a), orient_d(c, d, b)) < 0.0) return 0.0;
    }
  }
  if (inside_quad(p1, p2) || inside_quad(p2, p1)) return 0.0;
  const double d1 = pts_to_segs(p1, p2), d2 = pts_to_segs(p2, p1);
  return d1 < d2 ? d1 : d2;
}

__global__ void __launch_bounds__(128) textline_pairs_kernel(const double* __restrict__ q, int n, PairParams pp, uint8_t* __restrict__ adj) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  if (i >= j) { if (i == j) adj[idx] = 0; return; }
  const double* A = q + 16 * (size_t)i; const double* B = q + 16 * (size_t)j;
  const double x1 = A[8], y1 = A[9], w1 = A[10], h1 = A[11], x2 = B[8], y2 = B[9], w2 = B[10], h2 = B[11];
  const double fa = A[12], fb = B[12], ara = A[13], arb = B[13];
  const int fla = (int)A[15], flb = (int)B[15];
  uint8_t res = 0;
  // Types follow the host port under numpy 2 (NEP 50): font sizes and aspect ratios are float32 scalars, so products / quotients with
  // Python numbers stay float32 and a Python-float distance is rounded to float32 when compared with them; AABB numbers are int64 /
  // float64 and compare in float64.  (The pinned case: dist 51.0 against 85 * 0.6 = 51.000004f.)
  const float fa32 = (float)fa, fb32 = (float)fb, ara32 = (float)ara, arb32 = (float)arb;
  do {
    const float cs32 = fminf(fa32, fb32);
    const double cs = (double)cs32;
    const double gx = fmax(0.0, __dsub_rn(fmax(x1, x2), fmin(__dadd_rn(x1, w1), __dadd_rn(x2, w2))));
    const double gy = fmax(0.0, __dsub_rn(fmax(y1, y2), fmin(__dadd_rn(y1, h1), __dadd_rn(y2, h2))));
    if (sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy))) > (double)__fmul_rn((float)pp.gap, cs32)) break;
    if (!(fla & 2) || !(flb & 2)) { res = 2; break; }                     // a non-convex quad: let the host decide this pair
    const double dist = polygon_distance_d(A, B);
    const float dist32 = (float)dist;
    if (dist32 > __fmul_rn((float)pp.gap, cs32)) break;
    if (__fdiv_rn(fmaxf(fa32, fb32), cs32) > (float)pp.fs_tol) break;
    const float inv_ar = (float)(1.0 / pp.ar_tol);
    if (ara32 > (float)pp.ar_tol && arb32 < inv_ar) break;
    if (arb32 > (float)pp.ar_tol && ara32 < inv_ar) break;
    if ((fla & 1) && (flb & 1)) {
      if (dist32 >= __fmul_rn(cs32, (float)pp.tol)) break;
      if (fabs(__dsub_rn(__dadd_rn(x1, floor(w1 / 2.0)), __dadd_rn(x2, floor(w2 / 2.0)))) < pp.tol2) { res = 1; break; }
      if ((w1 > __dmul_rn(h1, pp.ratio) && h2 > __dmul_rn(w2, pp.ratio)) || (w2 > __dmul_rn(h2, pp.ratio) && h1 > __dmul_rn(w1, pp.ratio))) break;
      const double lim = (double)__fmul_rn(cs32, (float)pp.tol2);
      if (w1 > __dmul_rn(h1, pp.ratio) || w2 > __dmul_rn(h2, pp.ratio)) {
        res = (fabs(__dsub_rn(x1, x2)) < lim || fabs(__dsub_rn(__dadd_rn(x1, w1), __dadd_rn(x2, w2))) < lim) ? 1 : 0; break;
      }
      if (h1 > __dmul_rn(w1, pp.ratio) || h2 > __dmul_rn(w2, pp.ratio)) {
        res = (fabs(__dsub_rn(y1, y2)) < lim || fabs(__dsub_rn(__dadd_rn(y1, h1), __dadd_rn(y2, h2))) < lim) ? 1 : 0; break;
      }
      break;
    }
    if (fabs(__dsub_rn(A[14], B[14])) < 15.0 * 3.14159265358979323846 / 180.0) {
      if (dist32 > __fmul_rn(cs32, (float)pp.tol2)) break;               // poly_distance == polygon_distance for convex quads
      res = __fdiv_rn(fabsf(__fsub_rn(fa32, fb32)), cs32) <= 0.25f ? 1 : 0;
    }
    (void)cs;
  } while (false);
  adj[idx] = res;
  adj[(size_t)j * n + i] = res;
}

}  // namespace

void launch_textline_pairs(const double* quads, int n, const double* params6, uint8_t* adj, cudaStream_t st) {
  MITB_CHECK(n >= 0 && n <= 16384, "textline_pairs: too many lines");
  if (n == 0) return;
  PairParams pp{params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
  const long total = (long)n * n;
  textline_pairs_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(quads, n, pp, adj);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace mitb
