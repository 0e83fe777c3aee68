// sph_grid.cu -- uniform-grid neighbour search for sm_100a:
//   warp-aggregated cell hash + count  ->  block-scan exclusive prefix sum  ->  counting-sort
//   scatter  ->  per-cell index fix-up (stable order)  ->  reorder into cell-contiguous float4 SoA.
//
// Replaces ParticleSystem.reset_grid / update_grid_fluid_particles / update_grid_rigid_particles /
// update_boundary_grids (PS:322-335, 368-407): the reference's per-cell dynamic lists become a CSR
// (cell_start[G+1] + sorted_id[N]) whose order inside a cell is ascending original index, i.e. the
// arrival order of ti.cpu with one thread.  Cell ids are bit-exact: floor(pos / h) with an IEEE
// f32 division (PS:490-494, box_min ignored), 1-D id = x + gx*gz*y + gx*z (PS:102, 486-488).
#include "sph_internal.h"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

// ---------------------------------------------------------------------------------------------
// K0: cell hash + warp-aggregated count
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hash_count(const float4 *__restrict__ pos, int n, SphConsts c,
                                                     int *__restrict__ cell_of, int *__restrict__ cell_cnt,
                                                     SphCtl *ctl) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int idx = -1;
	if (i < n) {
		float4 p = pos[i];
		// PS:494 ti.floor(pos / support_radius, ti.i32): true division, then floor
		int cx = (int)floorf(__fdiv_rn(p.x, c.h));
		int cy = (int)floorf(__fdiv_rn(p.y, c.h));
		int cz = (int)floorf(__fdiv_rn(p.z, c.h));
		bool bad = cx < 0 || cy < 0 || cz < 0 || cx >= c.gx || cy >= c.gy || cz >= c.gz ||
		           !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
		if (bad) {
			// the reference only prints (PS:393-395, 492-493); latch a flag and keep the particle
			// addressable by clamping it into the grid
			atomicOr(&ctl->error_flags, SPH_ERR_OUT_OF_GRID);
			cx = min(max(cx, 0), c.gx - 1);
			cy = min(max(cy, 0), c.gy - 1);
			cz = min(max(cz, 0), c.gz - 1);
		}
		idx = cx + cy * c.gxz + cz * c.gx;
		cell_of[i] = idx;
	}
	// warp-aggregated increment: lanes that hit the same cell elect one leader
	unsigned peers = __match_any_sync(0xffffffffu, idx);
	if (idx >= 0) {
		int leader = __ffs(peers) - 1;
		if ((threadIdx.x & 31) == leader) atomicAdd(&cell_cnt[idx], __popc(peers));
	}
}

// ---------------------------------------------------------------------------------------------
// K1: exclusive prefix sum over the cell counts (three launches, block scan with warp shuffles)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) >= o) v += t;
	}
	return v;
}

// inclusive scan of one value per thread across the block; returns inclusive value, total in *total
__device__ __forceinline__ int block_incl_scan(int v, int *total) {
	__shared__ int wsum[32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int inc = warp_incl_scan(v);
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		int nw = (blockDim.x + 31) >> 5;
		int x = lane < nw ? wsum[lane] : 0;
		x = warp_incl_scan(x);
		wsum[lane] = x;
	}
	__syncthreads();
	int base = w > 0 ? wsum[w - 1] : 0;
	*total = wsum[((blockDim.x + 31) >> 5) - 1];
	__syncthreads();
	return inc + base;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int *__restrict__ cnt, int G,
                                                               int *__restrict__ sums) {
	int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	int s = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		int g = base + k;
		if (g < G) s += cnt[g];
	}
	int total;
	block_incl_scan(s, &total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int *sums, int nb) {
	__shared__ int carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (int base = 0; base < nb; base += 1024) {
		int i = base + threadIdx.x;
		int v = i < nb ? sums[i] : 0;
		int total;
		int inc = block_incl_scan(v, &total);
		int carry = carry_s;
		if (i < nb) sums[i] = carry + inc - v;
		__syncthreads();
		if (threadIdx.x == 0) carry_s = carry + total;
		__syncthreads();
	}
	if (threadIdx.x == 0) sums[nb] = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int *__restrict__ cnt, int G,
                                                              const int *__restrict__ sums, int nb,
                                                              int *__restrict__ start) {
	int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	int v[SCAN_ITEMS];
	int s = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		int g = base + k;
		v[k] = g < G ? cnt[g] : 0;
		s += v[k];
	}
	int total;
	int inc = block_incl_scan(s, &total);
	int run = sums[blockIdx.x] + inc - s;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		int g = base + k;
		if (g < G) {
			start[g] = run;
			cnt[g] = 0; // becomes the fill cursor of the scatter
		}
		run += v[k];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) start[G] = sums[nb];
}

// ---------------------------------------------------------------------------------------------
// K2: counting-sort scatter (arrival order inside a cell is arbitrary here ...)
// K3: ... and is made canonical (ascending original index) by a per-cell insertion sort
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(const int *__restrict__ cell_of, int n,
                                                  const int *__restrict__ start, int *__restrict__ fill,
                                                  int *__restrict__ sorted_id) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int c = cell_of[i];
	int slot = start[c] + atomicAdd(&fill[c], 1);
	sorted_id[slot] = i;
}

// gid (optional): global particle id used as the tie-break when the handle holds one slab of a
// multi-GPU domain, so that the order inside a cell equals the single-domain order.
// The scatter's atomics leave the particles of a cell in arbitrary order; the reference's order inside a cell is
// ascending particle index (append order of ti.cpu with one thread, PS:382-407).  Cells of up to 16 particles are
// rank-sorted in registers (all loads independent, n^2 compares, no dependent memory traffic: the thread-per-cell
// insertion sort it replaces was a chain of dependent global loads, 35 us at 650 k cells and 13 of 32 lanes busy);
// larger cells fall back to the insertion sort.
template <int M>
__device__ __forceinline__ void cell_rank_sort(int *__restrict__ ids, int n, const int *__restrict__ gid) {
	int id[M], key[M];
#pragma unroll
	for (int i = 0; i < M; ++i) {
		id[i] = i < n ? ids[i] : 0x7fffffff;
		key[i] = (i < n && gid) ? gid[id[i]] : id[i];
	}
#pragma unroll
	for (int i = 0; i < M; ++i) {
		int r = 0;
#pragma unroll
		for (int j = 0; j < M; ++j) r += key[j] < key[i] ? 1 : 0; // keys are unique; the padding ranks last
		if (i < n) ids[r] = id[i];
	}
}
__global__ void __launch_bounds__(256) k_cell_fix(const int *__restrict__ start, int G,
                                                   int *__restrict__ sorted_id, const int *__restrict__ gid) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= G) return;
	int a = start[c], b = start[c + 1];
	int n = b - a;
	if (n <= 1) return;
	if (n <= 8) { cell_rank_sort<8>(sorted_id + a, n, gid); return; }
	if (n <= 16) { cell_rank_sort<16>(sorted_id + a, n, gid); return; }
	for (int i = a + 1; i < b; ++i) {
		int key = sorted_id[i];
		int kk = gid ? gid[key] : key;
		int j = i - 1;
		while (j >= a) {
			int t = sorted_id[j];
			int tk = gid ? gid[t] : t;
			if (tk <= kk) break;
			sorted_id[j + 1] = t;
			--j;
		}
		sorted_id[j + 1] = key;
	}
}

// ---------------------------------------------------------------------------------------------
// K4: reorder into cell-contiguous float4 SoA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_fluid(const int *__restrict__ sorted_id,
                                                       const int *__restrict__ cell_of, int n,
                                                       const float4 *__restrict__ pos,
                                                       const float4 *__restrict__ vel,
                                                       float4 *__restrict__ spos, float4 *__restrict__ svel,
                                                       int *__restrict__ scell, int *__restrict__ slot_of) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int i = sorted_id[s];
	float4 p = pos[i];
	p.w = 0.0f;
	spos[s] = p;
	svel[s] = vel[i];
	scell[s] = cell_of[i];
	slot_of[i] = s;
}

__global__ void __launch_bounds__(256) k_gather_pos(const int *__restrict__ sorted_id,
                                                     const int *__restrict__ cell_of, int n,
                                                     const float4 *__restrict__ pos,
                                                     float4 *__restrict__ spos, int *__restrict__ scell) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int i = sorted_id[s];
	spos[s] = pos[i];
	scell[s] = cell_of[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_unsort(const int *__restrict__ sorted_id, int n,
                                                 const T *__restrict__ in, T *__restrict__ out) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	out[sorted_id[s]] = in[s];
}

__global__ void __launch_bounds__(256) k_writeback(const int *__restrict__ sorted_id, int n, int n_owned,
                                                    const float4 *__restrict__ spos,
                                                    const float4 *__restrict__ svel,
                                                    float4 *__restrict__ pos, float4 *__restrict__ vel) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int i = sorted_id[s];
	if (i >= n_owned) return; // ghost particles (multi-GPU) are read-only copies
	float4 p = spos[s];
	p.w = 0.0f;
	pos[i] = p;
	vel[i] = svel[s];
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void sphg_build(SphHandle *h, SphGrid &g, const float4 *pos, int n, cudaStream_t st, const int *gid) {
	const SphConsts &c = h->c;
	g.n = n;
	cudaMemsetAsync(g.cell_cnt, 0, sizeof(int) * (size_t)c.G, st);
	if (n > 0) {
		k_hash_count<<<cdiv(n, 256), 256, 0, st>>>(pos, n, c, g.cell_of, g.cell_cnt, h->ctl);
		h->launches++;
	}
	int nb = cdiv(c.G, SCAN_TILE);
	k_scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(g.cell_cnt, c.G, h->scan_sums);
	k_scan_sums<<<1, 1024, 0, st>>>(h->scan_sums, nb);
	k_scan_apply<<<nb, SCAN_THREADS, 0, st>>>(g.cell_cnt, c.G, h->scan_sums, nb, g.cell_start);
	h->launches += 3;
	if (n > 0) {
		k_scatter<<<cdiv(n, 256), 256, 0, st>>>(g.cell_of, n, g.cell_start, g.cell_cnt, g.sorted_id);
		k_cell_fix<<<cdiv(c.G, 256), 256, 0, st>>>(g.cell_start, c.G, g.sorted_id, gid);
		h->launches += 2;
	}
}

void sphg_gather_fluid(SphHandle *h, cudaStream_t st) {
	int n = h->fg.n;
	if (n <= 0) return;
	k_gather_fluid<<<cdiv(n, 256), 256, 0, st>>>(h->fg.sorted_id, h->fg.cell_of, n, h->pos, h->vel,
	                                              h->a4[A4_POS], h->a4[A4_VEL], h->fg.scell, h->fg.slot_of);
	h->launches++;
}

// The same in two parts, for a step whose velocities are still on their way from the host (sph_upload_state_xyz
// defers them behind the grid and list build): positions + the solver-persistent scalar vel.w now (it never leaves
// the device; the concurrent unpack kernel rewrites it with the same value), velocities when they have arrived.
__global__ void __launch_bounds__(256) k_gather_fluid_pos(const int *__restrict__ sorted_id, const int *__restrict__ cell_of, int n,
                                                           const float4 *__restrict__ pos, const float4 *vel,
                                                           float4 *__restrict__ spos, float4 *__restrict__ svel,
                                                           int *__restrict__ scell, int *__restrict__ slot_of) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int i = sorted_id[s];
	float4 p = pos[i];
	p.w = 0.0f;
	spos[s] = p;
	svel[s] = make_float4(0.0f, 0.0f, 0.0f, reinterpret_cast<const volatile float *>(vel + i)[3]);
	scell[s] = cell_of[i];
	slot_of[i] = s;
}
__global__ void __launch_bounds__(256) k_gather_vel(const int *__restrict__ sorted_id, int n, const float4 *__restrict__ vel,
                                                     float4 *__restrict__ svel) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s < n) svel[s] = vel[sorted_id[s]];
}
void sphg_gather_fluid_pos(SphHandle *h, cudaStream_t st) {
	int n = h->fg.n;
	if (n <= 0) return;
	k_gather_fluid_pos<<<cdiv(n, 256), 256, 0, st>>>(h->fg.sorted_id, h->fg.cell_of, n, h->pos, h->vel, h->a4[A4_POS],
	                                                  h->a4[A4_VEL], h->fg.scell, h->fg.slot_of);
	h->launches++;
}
void sphg_gather_vel(SphHandle *h, cudaStream_t st) {
	int n = h->fg.n;
	if (n <= 0) return;
	k_gather_vel<<<cdiv(n, 256), 256, 0, st>>>(h->fg.sorted_id, n, h->vel, h->a4[A4_VEL]);
	h->launches++;
}

void sphg_gather_boundary(SphHandle *h, cudaStream_t st) {
	int n = h->bg.n;
	if (n <= 0) return;
	k_gather_pos<<<cdiv(n, 256), 256, 0, st>>>(h->bg.sorted_id, h->bg.cell_of, n, h->bpos, h->bspos,
	                                            h->bg.scell);
	h->launches++;
}

void sphg_gather_rigid(SphHandle *h, cudaStream_t st) {
	int n = h->rg.n;
	if (n <= 0) return;
	k_gather_pos<<<cdiv(n, 256), 256, 0, st>>>(h->rg.sorted_id, h->rg.cell_of, n, h->rpos, h->rspos, h->rg.scell);
	h->launches++;
}

void sphg_unsort_f1(SphHandle *h, const SphGrid &g, const float *in, float *out, int n, cudaStream_t st) {
	if (n <= 0) return;
	k_unsort<float><<<cdiv(n, 256), 256, 0, st>>>(g.sorted_id, n, in, out);
	h->launches++;
}
void sphg_unsort_i1(SphHandle *h, const SphGrid &g, const int *in, int *out, int n, cudaStream_t st) {
	if (n <= 0) return;
	k_unsort<int><<<cdiv(n, 256), 256, 0, st>>>(g.sorted_id, n, in, out);
	h->launches++;
}
void sphg_unsort_f4(SphHandle *h, const SphGrid &g, const float4 *in, float4 *out, int n, cudaStream_t st) {
	if (n <= 0) return;
	k_unsort<float4><<<cdiv(n, 256), 256, 0, st>>>(g.sorted_id, n, in, out);
	h->launches++;
}
void sphg_writeback(SphHandle *h, const float4 *spos, const float4 *svel, cudaStream_t st) {
	int n = h->c.N;
	if (n <= 0) return;
	k_writeback<<<cdiv(n, 256), 256, 0, st>>>(h->fg.sorted_id, n, h->c.N_owned, spos, svel, h->pos, h->vel);
	h->launches++;
}

// ---------------------------------------------------------------------------------------------
// SB:219-245 visualize_rho / visualize_neighbour: min / max of a per-particle quantity (hierarchical:
// warp shuffles, one smem hop, one atomic per block on order-independent integer keys), then the colour
// map  rgb[i] = (0, 0.28, (q_i - min) / (max - min))  written in ORIGINAL particle order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2key(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; } // monotone
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

template <bool IS_INT>
__global__ void __launch_bounds__(256) k_vis_minmax(const void *__restrict__ q, const int *__restrict__ sorted_id, int n,
                                                     int n_owned, int *__restrict__ mm) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	float v = 0.0f;
	bool ok = s < n && sorted_id[s] < n_owned;
	if (ok) v = IS_INT ? (float)((const int *)q)[s] : ((const float *)q)[s];
	float lo = ok ? v : INFINITY, hi = ok ? v : -INFINITY;
	for (int o = 16; o > 0; o >>= 1) {
		lo = fminf(lo, __shfl_down_sync(0xffffffffu, lo, o));
		hi = fmaxf(hi, __shfl_down_sync(0xffffffffu, hi, o));
	}
	__shared__ float slo[8], shi[8];
	if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int k = 1; k < 8; ++k) { lo = fminf(lo, slo[k]); hi = fmaxf(hi, shi[k]); }
		atomicMin(&mm[0], f2key(lo));
		atomicMax(&mm[1], f2key(hi));
	}
}
__global__ void k_vis_init(int *mm) { mm[0] = f2key(INFINITY); mm[1] = f2key(-INFINITY); }

template <bool IS_INT>
__global__ void __launch_bounds__(256) k_vis_colour(const void *__restrict__ q, const int *__restrict__ sorted_id, int n,
                                                     int n_owned, const int *__restrict__ mm, float *__restrict__ rgb,
                                                     int stride) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || sorted_id[s] >= n_owned) return; // ghost copies (multi-GPU slabs) have no colour of their own
	float lo = key2f(mm[0]), hi = key2f(mm[1]);
	if (!(hi - lo > 0.0f)) return; // SB:230, 244: colours stay as they are
	float v = IS_INT ? (float)((const int *)q)[s] : ((const float *)q)[s];
	float *o = rgb + (size_t)sorted_id[s] * stride;
	o[0] = 0.0f; o[1] = 0.28f; o[2] = (v - lo) / (hi - lo);
}

// what: 0 = rho (SB:219-232), 1 = neighbour count (SB:234-245); rgb: n x stride floats, original order
void sphg_visualize(SphHandle *h, int what, float *rgb, int stride, cudaStream_t st) {
	int n = h->c.N, no = h->c.N_owned; // sorted slots hold owned particles and (multi-GPU) ghost copies
	if (n <= 0) return;
	int *mm = (int *)h->red; // 32 bytes of device scratch that no kernel uses between steps
	k_vis_init<<<1, 1, 0, st>>>(mm);
	if (what == 0) {
		k_vis_minmax<false><<<cdiv(n, 256), 256, 0, st>>>(h->a1[A1_RHO], h->fg.sorted_id, n, no, mm);
		k_vis_colour<false><<<cdiv(n, 256), 256, 0, st>>>(h->a1[A1_RHO], h->fg.sorted_id, n, no, mm, rgb, stride);
	} else {
		k_vis_minmax<true><<<cdiv(n, 256), 256, 0, st>>>(h->nbr_count, h->fg.sorted_id, n, no, mm);
		k_vis_colour<true><<<cdiv(n, 256), 256, 0, st>>>(h->nbr_count, h->fg.sorted_id, n, no, mm, rgb, stride);
	}
	h->launches += 3;
}

// ---------------------------------------------------------------------------------------------
// Device-side scene initialisation: the fluid lattice (PS:142-151) and the one-layer
// boundary shell (PS:155-195) of init_particle_pos, written straight into the caller's float4 arrays.
// Stateless (no handle: the reference fills the positions before anything else exists, PS:119).
// Arithmetic is the reference's, operation by operation, with un-fused IEEE intrinsics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_init_fluid_lattice(long long total, const int *__restrict__ ids, long long n, float x_num, float z_num, float xz_num,
                     int xi, int zi, float rad, float sx, float sy, float sz, float4 *__restrict__ pos) {
	long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	long long id = ids ? (long long)ids[k] : k;
	float x, y, z;
	if (total < (1LL << 24)) {
		float fi = (float)id;                                            // PS:146-149: f32 index arithmetic
		float qx = floorf(__fdiv_rn(fi, x_num));
		x = __fsub_rn(fi, __fmul_rn(x_num, qx));                         // i % x_num
		z = __fsub_rn(qx, __fmul_rn(z_num, floorf(__fdiv_rn(qx, z_num)))); // (i // x_num) % z_num
		y = (float)(int)__fdiv_rn(fi, xz_num);                           // int(i / xz_num)
	} else { // the f32 index arithmetic is inexact beyond 2^24: integer lattice
		x = (float)(id % xi);
		z = (float)((id / xi) % zi);
		y = (float)(id / ((long long)xi * zi));
	}
	pos[k] = make_float4(__fadd_rn(__fmul_rn(__fmul_rn(x, rad), 2.0f), sx), __fadd_rn(__fmul_rn(__fmul_rn(y, rad), 2.0f), sy),
	                     __fadd_rn(__fmul_rn(__fmul_rn(z, rad), 2.0f), sz), 0.0f); // PS:150
}

__global__ void __launch_bounds__(256)
k_init_boundary_shell(long long nb, long long x_cnt, long long z_cnt, long long bottom, long long one_round, float d,
                      float box_max_y, float4 *__restrict__ bpos) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nb) return;
	const long long xr = x_cnt - 1, zr = z_cnt - 1;
	float x = 0.0f, y = 0.0f, z = 0.0f;
	if (i < bottom) {                                                    // PS:164-168 floor
		x = __fmul_rn((float)(i % x_cnt), d);
		z = __fmul_rn(floorf(__fdiv_rn((float)i, (float)x_cnt)), d);
	} else if (i < nb - bottom) {                                        // PS:169-189 walls, layer by layer
		long long idx = i - bottom;
		long long layer = (long long)floorf(__fdiv_rn((float)idx, (float)one_round));
		y = __fmul_rn(d, (float)(layer + 1));
		idx = idx - layer * one_round + 1;
		if (idx <= xr) x = __fmul_rn((float)(idx % xr), d);
		else if (idx <= xr + zr) { x = __fmul_rn((float)xr, d); z = __fmul_rn((float)((idx - x_cnt) % zr), d); }
		else if (idx <= 2 * xr + zr) { x = __fmul_rn((float)((2 * xr + zr - idx) % xr + 1), d); z = __fmul_rn((float)zr, d); }
		else if (idx <= 2 * (xr + zr)) z = __fmul_rn((float)((2 * (xr + zr) - idx) % zr + 1), d);
	} else {                                                             // PS:190-195 lid
		long long i2 = i - (nb - bottom);
		x = __fmul_rn((float)(i2 % x_cnt), d);
		y = box_max_y;
		z = __fmul_rn((float)(long long)__fdiv_rn((float)i2, (float)x_cnt), d);
	}
	bpos[i] = make_float4(x, y, z, 0.0f);
}

extern "C" int sph_init_fluid_lattice(const SphLattice *lat, long long particle_num_total, const int32_t *dev_ids, size_t n,
                                      void *dev_pos4, int device, void *stream) {
	if (!lat) return SPH_EINVAL;
	if (n == 0) return SPH_OK;
	if (!dev_pos4) return SPH_EINVAL;
	if (cudaSetDevice(device) != cudaSuccess) return SPH_ECUDA;
	double d = lat->particle_radius * 2;
	double x_num_d = lat->water_size[0] / d, z_num_d = lat->water_size[2] / d;
	k_init_fluid_lattice<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
	    particle_num_total, dev_ids, (long long)n, (float)x_num_d, (float)z_num_d, (float)x_num_d * (float)z_num_d, // PS:145: f32 locals
	    (int)llround(x_num_d), (int)llround(z_num_d), (float)lat->particle_radius, (float)lat->start_pos[0],
	    (float)lat->start_pos[1], (float)lat->start_pos[2], (float4 *)dev_pos4);
	return cudaGetLastError() == cudaSuccess ? SPH_OK : SPH_ECUDA;
}

extern "C" int sph_init_boundary_shell(const SphLattice *lat, size_t nb, void *dev_bpos4, int device, void *stream) {
	if (!lat) return SPH_EINVAL;
	if (nb == 0) return SPH_OK;
	if (!dev_bpos4) return SPH_EINVAL;
	if (cudaSetDevice(device) != cudaSuccess) return SPH_ECUDA;
	double dd = lat->particle_radius * 2;
	// PS:155-158: `box` is bound to a kernel local, i.e. an f32 vector, so the two counts are f32 arithmetic in the kernel
	// although compute_boundary_particles_count (PS:129-137, host, fp64) sized the array; they disagree for some boxes
	// (5.2 / 0.05: 105 on the host, 104 here) and the reference lays its shell out with THESE counts (quirk B-18)
	const float df = (float)dd;
	long long x_cnt = (long long)((float)(lat->box_max[0] - lat->box_min[0]) / df + 1.0f);
	long long z_cnt = (long long)((float)(lat->box_max[2] - lat->box_min[2]) / df + 1.0f);
	long long bottom = x_cnt * z_cnt, one_round = x_cnt * z_cnt - (x_cnt - 2) * (z_cnt - 2);
	k_init_boundary_shell<<<(unsigned)((nb + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
	    (long long)nb, x_cnt, z_cnt, bottom, one_round, (float)dd, (float)lat->box_max[1], (float4 *)dev_bpos4);
	return cudaGetLastError() == cudaSuccess ? SPH_OK : SPH_ECUDA;
}

// ---------------------------------------------------------------------------------------------
// Host state as the reference's callers see it (pos / vel as N x 3, main.py:190): packed xyz staging so
// that the PCIe transfers carry 12 instead of 16 bytes per vector.  vel.w (the solver-persistent scalar:
// DFSPH warm_start_k, IISPH p_past) never leaves the device.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_xyz(const float *__restrict__ p3, const float *__restrict__ v3, int n,
                                                     float4 *__restrict__ pos, float4 *__restrict__ vel) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (p3) pos[i] = make_float4(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], 0.0f);
	if (v3) vel[i] = make_float4(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2], vel[i].w);
}
__global__ void __launch_bounds__(256) k_pack_xyz(const float4 *__restrict__ pos, const float4 *__restrict__ vel, int n,
                                                   float *__restrict__ p3, float *__restrict__ v3) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (p3) { float4 p = pos[i]; p3[3 * i] = p.x; p3[3 * i + 1] = p.y; p3[3 * i + 2] = p.z; }
	if (v3) { float4 v = vel[i]; v3[3 * i] = v.x; v3[3 * i + 1] = v.y; v3[3 * i + 2] = v.z; }
}
void sphg_unpack_xyz(SphHandle *h, const float *p3, const float *v3, int n, cudaStream_t st) {
	if (n <= 0) return;
	k_unpack_xyz<<<cdiv(n, 256), 256, 0, st>>>(p3, v3, n, h->pos, h->vel);
	h->launches++;
}
void sphg_pack_xyz(SphHandle *h, float *p3, float *v3, int n, cudaStream_t st) {
	if (n <= 0) return;
	k_pack_xyz<<<cdiv(n, 256), 256, 0, st>>>(h->pos, h->vel, n, p3, v3);
	h->launches++;
}
