// Full-resolution 7x7 convolution with a thin output (Cout <= 4): LaMa's final ReflectionPad2d(3) + Conv2d(64->3, k7) +
// Sigmoid (inpainting_lama_mpe.py:596-601).  59 GFLOP of fp32 per 2048x1536 page, but as an implicit GEMM every input
// value would be re-fetched 49 times from L2 (39 GB of gather traffic).  Here a CTA stages the input tile with its 3-pixel
// halo in shared memory ONCE per 8-channel slab (planar [c][y][x] layout, reflect/zero padding resolved while staging)
// and each thread produces 4 horizontally adjacent pixels x Cout channels from registers:
//   per (channel, tap row): 3 LDS.128 of inputs + 7 broadcast LDS.128 of weights feed 84 FFMA  ->  FMA-pipe bound.
#include "mitb_internal.h"

namespace mitb {

namespace {
constexpr int TW = 64, TH = 16, HALO = 3, KS = 7;
constexpr int SW = TW + 2 * HALO + 2;            // 72: row pitch (x from -4 .. 67 so that float4 loads stay aligned)
constexpr int SH = TH + 2 * HALO;                // 22
constexpr int CCH = 8;                           // channels per slab

struct ThinParams {
  const float* in; int N, H, W, in_cs, in_coff, Cin;
  const float* w;                                // [tap][Cin][4] (K-major fp32, ldw = 4)
  float* out; int out_cs, out_coff, Cout, out_planar;
  const float* shift; int act, pad;
};

__device__ __forceinline__ float act_thin(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SILU: return v / (1.f + expf(-v));
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(256, 2) conv7_thin_kernel(const ThinParams p) {
  extern __shared__ __align__(16) float thin_smem[];
  float (*tile)[SH][SW] = reinterpret_cast<float (*)[SH][SW]>(thin_smem);                       // [CCH][22][72] = 50688 B
  float (*wsm)[KS * KS][4] = reinterpret_cast<float (*)[KS * KS][4]>(thin_smem + CCH * SH * SW);  // [CCH][49][4] = 6272 B
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, n = blockIdx.z;
  const int tid = threadIdx.x;
  const int lx = (tid & 15) * 4, ly = tid >> 4;          // this thread's 4 output pixels: (ty0+ly, tx0+lx .. +3)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < p.Cin; c0 += CCH) {
    __syncthreads();
    // ---- stage the slab: pixels (y in [ty0-3, ty0+TH+3), x in [tx0-4, tx0+TW+4)), 8 channels, padding resolved here
    for (int i = tid; i < SH * SW; i += 256) {
      const int sy = i / SW, sx = i - sy * SW;
      int gy = ty0 + sy - HALO, gx = tx0 + sx - 4;
      bool ok = true;
      if (p.pad == PAD_REFLECT) {
        if (gy < 0) gy = -gy; if (gy >= p.H) gy = 2 * p.H - 2 - gy;
        if (gx < 0) gx = -gx; if (gx >= p.W) gx = 2 * p.W - 2 - gx;
        ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;      // far outside the image (tile overhang): unused values
      } else ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (ok) {
        const float* src = p.in + ((size_t)(n * p.H + gy) * p.W + gx) * p.in_cs + p.in_coff + c0;
        a = __ldg(reinterpret_cast<const float4*>(src)); b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      }
      tile[0][sy][sx] = a.x; tile[1][sy][sx] = a.y; tile[2][sy][sx] = a.z; tile[3][sy][sx] = a.w;
      tile[4][sy][sx] = b.x; tile[5][sy][sx] = b.y; tile[6][sy][sx] = b.z; tile[7][sy][sx] = b.w;
    }
    for (int i = tid; i < CCH * KS * KS; i += 256) {
      const int c = i / (KS * KS), t = i - c * (KS * KS);
      *reinterpret_cast<float4*>(&wsm[c][t][0]) = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)t * p.Cin + c0 + c) * 4));
    }
    __syncthreads();
    // ---- accumulate
#pragma unroll 1
    for (int c = 0; c < CCH; ++c) {
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        // inputs x = lx-3 .. lx+6 live at smem columns (lx+1) .. (lx+10); load the aligned span [lx, lx+12)
        const float4 v0 = *reinterpret_cast<const float4*>(&tile[c][ly + dy][lx]);
        const float4 v1 = *reinterpret_cast<const float4*>(&tile[c][ly + dy][lx + 4]);
        const float4 v2 = *reinterpret_cast<const float4*>(&tile[c][ly + dy][lx + 8]);
        const float in[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4 wv = *reinterpret_cast<const float4*>(&wsm[c][dy * KS + dx][0]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = in[1 + i + dx];
            acc[i][0] = fmaf(x, wv.x, acc[i][0]); acc[i][1] = fmaf(x, wv.y, acc[i][1]);
            acc[i][2] = fmaf(x, wv.z, acc[i][2]); acc[i][3] = fmaf(x, wv.w, acc[i][3]);
          }
        }
      }
    }
  }
  const int oy = ty0 + ly;
  if (oy >= p.H) return;
  const size_t plane = (size_t)p.H * p.W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = tx0 + lx + i;
    if (ox >= p.W) continue;
    const size_t pix = (size_t)oy * p.W + ox;
    for (int j = 0; j < p.Cout; ++j) {
      float v = acc[i][j] + (p.shift ? p.shift[j] : 0.f);
      v = act_thin(v, p.act);
      if (p.out_planar) p.out[((size_t)n * p.out_cs + p.out_coff + j) * plane + pix] = v;
      else p.out[((size_t)n * plane + pix) * p.out_cs + p.out_coff + j] = v;
    }
  }
}
}  // namespace

bool conv_thin_supported(const ConvOp& op) {
  if (op.in.planar || op.stat_max || op.ntaps != KS * KS || op.out.C > 4 || op.ldw != 4) return false;
  if (op.sy != 1 || op.sx != 1 || op.in.C % CCH != 0 || op.in.cs % 4 != 0 || op.in.coff % 4 != 0) return false;
  if (op.Ho != op.in.H || op.Wo != op.in.W || op.oy_mul != 1 || op.ox_mul != 1 || op.oy_add || op.ox_add) return false;
  if (op.in_scale || op.add0.p || op.add1.p || op.scale || op.mul1) return false;
  if (op.act != ACT_NONE && op.act != ACT_RELU && op.act != ACT_SILU && op.act != ACT_SIGMOID) return false;
  for (int t = 0; t < op.ntaps; ++t)
    if (op.tdy[t] != t / KS - HALO || op.tdx[t] != t % KS - HALO) return false;
  return op.in.H >= 4 && op.in.W >= 4;
}

void launch_conv_thin(const ConvOp& op, cudaStream_t st) {
  ThinParams p;
  p.in = op.in.p; p.N = op.in.N; p.H = op.in.H; p.W = op.in.W; p.in_cs = op.in.cs; p.in_coff = op.in.coff; p.Cin = op.in.C;
  p.w = op.w; p.out = op.out.p; p.out_cs = op.out.cs; p.out_coff = op.out.coff; p.Cout = op.out.C; p.out_planar = op.out.planar;
  p.shift = op.shift; p.act = op.act; p.pad = op.pad;
  const size_t smem = (size_t)(CCH * SH * SW + CCH * KS * KS * 4) * sizeof(float);
  static bool attr = false;
  if (!attr) { CUDA_OK(cudaFuncSetAttribute(conv7_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  dim3 grid((p.W + TW - 1) / TW, (p.H + TH - 1) / TH, p.N);
  conv7_thin_kernel<<<grid, 256, smem, st>>>(p);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace mitb
