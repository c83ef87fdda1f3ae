// Device helpers shared by the tcgen05 convolution kernels (conv_tc.cu: register-gather producers; conv_tma.cu: TMA-fed
// operands): mbarrier / TMA / tcgen05 / TMEM wrappers, the UMMA shared-memory descriptor, the bf16 hi/mid split and the
// epilogue activations.  Included inside namespace mitb { namespace { ... } } of each translation unit.
#pragma once

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();          // ~2 s at 2 GHz: far beyond any legitimate wait
    }
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1) | [32,46) stride byte
//   offset >> 4 (1024 >> 4) | [46,48) descriptor version 1 (sm_100) | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float apply_act_tc(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_SILU: return v / (1.f + expf(-v));
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACT_SIGMOID2: { float s = 1.f / (1.f + expf(-v)); return 1.f / (1.f + expf(-s)); }
    case ACT_CLAMP01: return fminf(fmaxf(v, 0.f), 1.f);
    default: return v;
  }
}
__device__ __forceinline__ int reflect_tc(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// split 8 fp32 values into bf16 hi / mid packs (16 bytes each): 6 instructions per pair
// (F2FP pack-convert for hi, shift/mask to get hi back as fp32, two FADD for the remainder, F2FP for mid)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);     // .x (low half) = v[2i]
    const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&hb);
    const float h0 = __uint_as_float(hbits << 16), h1 = __uint_as_float(hbits & 0xffff0000u);
    const __nv_bfloat162 mb = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
    h[i] = hbits;
    m[i] = *reinterpret_cast<const uint32_t*>(&mb);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}


// split 4 fp32 values into bf16 hi / mid packs (8 bytes each)
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& mid) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&h0), b1 = *reinterpret_cast<const uint32_t*>(&h1);
  const __nv_bfloat162 m0 = __floats2bfloat162_rn(v.x - __uint_as_float(b0 << 16), v.y - __uint_as_float(b0 & 0xffff0000u));
  const __nv_bfloat162 m1 = __floats2bfloat162_rn(v.z - __uint_as_float(b1 << 16), v.w - __uint_as_float(b1 & 0xffff0000u));
  hi = make_uint2(b0, b1);
  mid = make_uint2(*reinterpret_cast<const uint32_t*>(&m0), *reinterpret_cast<const uint32_t*>(&m1));
}

// exact-erf GELU with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): two MUFU + ~14 FP32 instructions instead of the
// ~35 of erff(); the GELU epilogue of the ConvNeXt fc1 layers is otherwise longer than their 2..8 K-block main loops
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.44269504088896340736f));
  const float erf_abs = fmaf(-pl, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// The same GELU on the packed fp32x2 pipe of sm_100 (fma.rn.f32x2 / mul.f32x2), two elements per call, rewritten so that no
// sign transfer is needed:  gelu(x) = max(x, 0) - 0.5 |x| p(t) exp(-z^2),  z = |x| / sqrt 2,  t = 1 / (1 + 0.3275911 z)
// (x >= 0: x (1 - pe/2); x < 0: x pe/2 = -|x| pe/2).  The polynomial coefficients carry the factor -1/2; exp(-z^2) =
// ex2(-(|x| sqrt(log2(e)/2))^2).  19 instructions per PAIR (4 of them MUFU) against 17 per element for gelu_fast.
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = __ffma2_rn(ax, make_float2(0.23164189028714973f, 0.23164189028714973f), make_float2(1.f, 1.f));   // 0.3275911 / sqrt 2
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  float2 pl = __ffma2_rn(t, make_float2(-0.5307027145f, -0.5307027145f), make_float2(0.7265760135f, 0.7265760135f));
  pl = __ffma2_rn(pl, t, make_float2(-0.7107068705f, -0.7107068705f));
  pl = __ffma2_rn(pl, t, make_float2(0.142248368f, 0.142248368f));
  pl = __ffma2_rn(pl, t, make_float2(-0.127414796f, -0.127414796f));
  pl = __fmul2_rn(pl, t);                                   // -p(t) / 2
  const float2 u = __fmul2_rn(ax, make_float2(0.84932180028801904f, 0.84932180028801904f));      // sqrt(log2(e) / 2)
  const float2 w = __fmul2_rn(u, u);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(-w.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(-w.y));
  const float2 h = __fmul2_rn(ax, pl);
  return __ffma2_rn(h, e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// split 4 fp32 values (two packed pairs) into bf16 hi / mid packs with the remainder on the packed pipe: 10 instructions
__device__ __forceinline__ void split4p(float2 a, float2 b, uint2& hi, uint2& mid) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(b.x, b.y);
  const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&h0), b1 = *reinterpret_cast<const uint32_t*>(&h1);
  const float2 m1 = make_float2(-1.f, -1.f);
  const float2 r0 = __ffma2_rn(make_float2(__uint_as_float(b0 << 16), __uint_as_float(b0 & 0xffff0000u)), m1, a);   // a - hi, exact
  const float2 r1 = __ffma2_rn(make_float2(__uint_as_float(b1 << 16), __uint_as_float(b1 & 0xffff0000u)), m1, b);
  const __nv_bfloat162 q0 = __floats2bfloat162_rn(r0.x, r0.y), q1 = __floats2bfloat162_rn(r1.x, r1.y);
  hi = make_uint2(b0, b1);
  mid = make_uint2(*reinterpret_cast<const uint32_t*>(&q0), *reinterpret_cast<const uint32_t*>(&q1));
}

template <int ACT>
__device__ __forceinline__ float act_t(float v, int act_rt);
template <int ACT>
__device__ __forceinline__ float2 act_t2(float2 v, int act_rt) {
  if (ACT == ACT_NONE) return v;
  if (ACT == ACT_RELU) return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  if (ACT == ACT_GELU) return gelu_fast2(v);
  return make_float2(act_t<ACT>(v.x, act_rt), act_t<ACT>(v.y, act_rt));
}

template <int ACT>
__device__ __forceinline__ float act_t(float v, int act_rt) {
  if (ACT == ACT_NONE) return v;
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_GELU) return gelu_fast(v);
  if (ACT == ACT_SILU) return v / (1.f + expf(-v));
  return apply_act_tc(v, act_rt);               // ACT == -1: rare activations, runtime switch
}

