// sph_math.cuh -- SPH kernel functions and pair geometry, in two arithmetic modes.
//
// This header is included by sph_sweeps.cu, which is compiled TWICE:
//   -DSPH_STRICT=1 -fmad=false : namespace sph_strict.  IEEE division / sqrt, no FMA contraction and
//                                 the reference's operation order (SB:74-103), so every per-particle
//                                 sum is bit-identical to the strict-fp32 oracle.
//   -DSPH_STRICT=0 -fmad=true  : namespace sph_fast.  FMA contraction, rsqrt / reciprocal
//                                 approximations; same neighbour sets (the cull is always exact),
//                                 fields within 1e-5 of the oracle.
#pragma once
#include "sph_common.cuh"

#ifndef SPH_STRICT
#error "compile with -DSPH_STRICT=0 or 1"
#endif

#if SPH_STRICT
#define SPH_NS sph_strict
#else
#define SPH_NS sph_fast
#endif

#define PI_F_DEV ((float)3.141592653589793)

namespace SPH_NS {

struct f3 {
	float x, y, z;
};
__device__ __forceinline__ f3 F3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 xyz(const float4 &a) { return F3(a.x, a.y, a.z); }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return F3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return F3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 neg(f3 a) { return F3(-a.x, -a.y, -a.z); }
// Taichi: dot = (a0*b0 + a1*b1) + a2*b2
__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
	return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float4 F4(f3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

// Pair geometry.  r2 is ALWAYS computed with separate multiplies and adds (no FMA) so that the
// cull  sqrt(r2) > h  <=>  r2 > cull_t  (PS:466) selects exactly the reference's
// neighbour set in both modes.
struct Pair {
	f3 r;     // x_i - x_j
	float r2; // (dx*dx + dy*dy) + dz*dz
};
__device__ __forceinline__ Pair make_pair(const float4 &pi, const float4 &pj) {
	Pair p;
	p.r.x = __fsub_rn(pi.x, pj.x);
	p.r.y = __fsub_rn(pi.y, pj.y);
	p.r.z = __fsub_rn(pi.z, pj.z);
	p.r2 = __fadd_rn(__fadd_rn(__fmul_rn(p.r.x, p.r.x), __fmul_rn(p.r.y, p.r.y)), __fmul_rn(p.r.z, p.r.z));
	return p;
}
__device__ __forceinline__ bool culled(const Pair &p, const SphConsts &c) { return p.r2 > c.cull_t; }

// Pair geometry inside the list-walking sweeps.  The neighbour set is already fixed by the lists, so
// the fast kernels may contract r2 into FMAs; the strict kernels keep the reference arithmetic.
__device__ __forceinline__ Pair sweep_pair(const float4 &pi, const float4 &pj) {
#if SPH_STRICT
	return make_pair(pi, pj);
#else
	Pair p;
	p.r.x = pi.x - pj.x;
	p.r.y = pi.y - pj.y;
	p.r.z = pi.z - pj.z;
	p.r2 = fmaf(p.r.z, p.r.z, fmaf(p.r.y, p.r.y, p.r.x * p.r.x));
	return p;
#endif
}

// SB:74-88 cubic_kernel(|r|, h)
__device__ __forceinline__ float cubic_w(const Pair &p, const SphConsts &c) {
#if SPH_STRICT
	float r = sqrtf(p.r2);
	float q = r / c.h;
	float ret = 0.0f;
	if (0.0f <= q && q <= 0.5f) {
		float q2 = q * q;
		float q3 = q2 * q;
		ret = c.kW * (6.0f * (q3 - q2) + 1.0f);
	} else if (0.5f < q && q <= 1.0f) {
		float t = 1.0f - q;
		ret = (2.0f * c.kW) * (t * (t * t));
	}
	return ret;
#else
	float q = sqrtf(p.r2) * c.inv_h;
	float t = 1.0f - q;
	float a = c.kW * (6.0f * (q * q) * (q - 1.0f) + 1.0f);
	float b = (2.0f * c.kW) * (t * t * t);
	return q <= 0.5f ? a : (q <= 1.0f ? b : 0.0f);
#endif
}

// SB:74-88 with an explicit distance (PCISPH evaluates W on predicted positions, PC:141-142)
__device__ __forceinline__ float cubic_w_r(float r, const SphConsts &c) {
	float q = r / c.h;
	float ret = 0.0f;
	if (0.0f <= q && q <= 0.5f) {
		float q2 = q * q;
		float q3 = q2 * q;
		ret = c.kW * (6.0f * (q3 - q2) + 1.0f);
	} else if (0.5f < q && q <= 1.0f) {
		float t = 1.0f - q;
		ret = (2.0f * c.kW) * (t * (t * t));
	}
	return ret;
}

// SB:90-103 cubic_kernel_derivative(r, h), including the reference's extra factor 6
__device__ __forceinline__ f3 cubic_dw(const Pair &p, const SphConsts &c) {
#if SPH_STRICT
	float r_norm = sqrtf(p.r2);
	float q = r_norm / c.h;
	f3 ret = F3(0.0f, 0.0f, 0.0f);
	if (1e-5f < q && q <= 0.5f) {
		float q2 = q * q;
		float co = c.kDW6 * (3.0f * q2 - 2.0f * q);
		float den = c.h * r_norm;
		ret = (co * p.r) / den;
	} else if (0.5f < q && q <= 1.0f) {
		float t = 1.0f - q;
		float co = c.nkDW6 * (t * t);
		float den = c.h * r_norm;
		ret = (co * p.r) / den;
	}
	return ret;
#else
	// grad W = g * r with g = coef(q) / (h |r|).  For q <= 1/2: kDW6 q (3q - 2) / (h |r|) = (kDW6 / h^2)(3q - 2)
	// (finite at r = 0, where r itself vanishes); for 1/2 < q <= 1: nkDW6 (1 - q)^2 / (h |r|).
	float rinv = rsqrtf(fmaxf(p.r2, 1e-30f));
	float q = (p.r2 * rinv) * c.inv_h;
	float t = 1.0f - q;
	float ga = fmaf(3.0f * c.dwA, q, -2.0f * c.dwA);
	float gb = (c.dwB * (t * t)) * rinv;
	float g = q <= 0.5f ? ga : (q <= 1.0f ? gb : 0.0f);
	return g * p.r;
#endif
}

// fast-mode division helper: approximate reciprocal is well inside the 1e-5 budget
__device__ __forceinline__ float sdiv(float a, float b) {
#if SPH_STRICT
	return a / b;
#else
	return __fdividef(a, b);
#endif
}

// ---- block reductions (hierarchical: warp shuffles, then one shared-memory hop) ---------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
	return v;
}

// every thread of the block must call; thread 0 writes the block's partial
__device__ __forceinline__ void block_partial(double sum, int cnt, float mx, SphPartial *out) {
	__shared__ double ssum[SPH_BLOCK / 32];
	__shared__ int scnt[SPH_BLOCK / 32];
	__shared__ float smax[SPH_BLOCK / 32];
	sum = warp_sum_d(sum);
	cnt = warp_sum_i(cnt);
	mx = warp_max_f(mx);
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (lane == 0) { ssum[w] = sum; scnt[w] = cnt; smax[w] = mx; }
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		int n = 0;
		float m = -INFINITY;
#pragma unroll
		for (int k = 0; k < SPH_BLOCK / 32; ++k) { s += ssum[k]; n += scnt[k]; m = fmaxf(m, smax[k]); }
		out[blockIdx.x].sum = s;
		out[blockIdx.x].cnt = n;
		out[blockIdx.x].maxv = m;
	}
}

} // namespace SPH_NS
