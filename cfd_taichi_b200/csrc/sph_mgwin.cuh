// sph_mgwin.cuh -- layout and access primitives of the CUDA-IPC peer windows (multi-GPU slabs), shared by the
// exchange kernels (sph_multigpu.cu) and the sweeps that push their edge values from their own epilogue (sph_sweeps.cu).
//
// "LL16" protocol: every 16-byte slot carries the exchange's epoch in its last word and is written with ONE 128-bit
// store, which NVLink delivers atomically (the assumption NCCL's LL / LL128 protocols make); the receiver polls the
// slot until the tag matches.  No fence, no flag, no proxy thread, no host.
#pragma once
#include <stdint.h>

#define SPH_MG_MAX_RANKS 16

// ---- window layout (identical on every rank) ----------------------------------------------------------
//   [0, 4096)            loop partials: rslot[2 parity][MAX_RANKS] = {sum | tag}, {count, max | tag}; sync slots
//   [4096, ...)          float4 xr[2 parity][2 side][cap_halo], .w = epoch tag
struct MgCtlWin {
	double rslot[2][SPH_MG_MAX_RANKS][4]; // two tagged 16-byte slots per (parity, source rank)
	uint4 sync[2][2];                     // [parity][side]: "my neighbour on that side has entered this exchange"
};
static_assert(sizeof(MgCtlWin) <= 4096, "window control block");
__host__ __device__ static inline MgCtlWin *win_ctl(char *w) { return (MgCtlWin *)w; }
__host__ __device__ static inline float4 *win_xr(char *w, int cap, int parity, int side) {
	return (float4 *)(w + 4096) + ((size_t)parity * 2 + (size_t)side) * (size_t)cap;
}

// particle messages (migration = kind 0, ghost particles = kind 1), one area per (kind, receiving side), behind the
// value slots: [count slot | cap x {pos.xyz | tag} | cap x {vel.xyz | tag} | cap x {vel.w, gid, - | tag}]
__host__ __device__ static inline size_t win_pm_bytes(int cap) { return 64 + (size_t)48 * (size_t)cap; }
__host__ __device__ static inline char *win_pm(char *w, int cap, int kind, int side) {
	return w + 4096 + (size_t)64 * (size_t)cap + ((size_t)kind * 2 + (size_t)side) * win_pm_bytes(cap);
}
__host__ __device__ static inline size_t win_total_bytes(int cap) { return 4096 + (size_t)64 * (size_t)cap + 4 * win_pm_bytes(cap); }

__device__ __forceinline__ void st_slot(void *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
	asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_slot(const void *p) {
	uint4 r;
	asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
	return r;
}

// What a producing sweep needs to push its edge values itself (instead of the exchange kernel reading them back
// after the sweep has ended): the values are on the wire while the sweep is still running, so the exchange kernel
// that follows only polls and unpacks.  tag == nullptr on one GPU (the sweeps test it once per thread).
struct SphMgPush {
	const int2 *tag;      // per sorted slot: (index in the left neighbour's receive block, index in the right one's), -1 = not sent
	char *peer_l, *peer_r; // the neighbours' windows (null at the domain ends)
	int *pushed_epoch;    // device word: the producer records the epoch it pushed (a gated-off producer does not)
	int cap, epoch;
};
// one value (x, y, z) of sorted particle s to whichever neighbours hold a ghost copy of it
__device__ __forceinline__ void mg_push(const SphMgPush &pu, int s, float x, float y, float z) {
	if (!pu.tag) return;
	int2 t = pu.tag[s];
	const int parity = pu.epoch & 1;
	// my left neighbour receives on its right side (1), my right neighbour on its left side (0)
	if (t.x >= 0) st_slot(&win_xr(pu.peer_l, pu.cap, parity, 1)[t.x], __float_as_uint(x), __float_as_uint(y), __float_as_uint(z), (uint32_t)pu.epoch);
	if (t.y >= 0) st_slot(&win_xr(pu.peer_r, pu.cap, parity, 0)[t.y], __float_as_uint(x), __float_as_uint(y), __float_as_uint(z), (uint32_t)pu.epoch);
}
// thread 0 of block 0 of a producer that really ran
__device__ __forceinline__ void mg_push_mark(const SphMgPush &pu) {
	if (pu.tag && blockIdx.x == 0 && threadIdx.x == 0) *pu.pushed_epoch = pu.epoch;
}
