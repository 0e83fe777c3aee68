// sph_sweeps.cu -- neighbour-list construction and the for_all_neighbor sweeps of every solver.
// Compiled twice (see sph_math.cuh): namespace sph_strict (-fmad=false) and sph_fast.
//
// Design: positions are frozen inside a solver step (they change only in the
// final integrate kernel: DF:238, PC:206, II:191, WC:52), so the 27-cell traversal with its exact
// distance cull (PS:447-469, 337-366) is done ONCE per step by k_build_lists, which emits compact
// per-particle neighbour lists (strict kernels: the reference's canonical visiting order; fast kernels:
// address order) and, for DFSPH, the per-pair gradient cache.  Every later sweep walks those lists: one
// coalesced 128-bit load of four indices and one float4 gather per neighbour.  Quantities a
// neighbour contributes as a single scalar (k_j/rho_j, p_j/rho_j^2 ...) are pre-divided by their
// producer kernel and ride in the .w lane of a position copy, so one 16-byte gather per pair
// brings everything.  No per-pair atomics anywhere; reductions are warp-shuffle + one smem hop.
#include "sph_math.cuh"
#include "sph_internal.h"
#include "sph_ctl.cuh"
#include "sph_mgwin.cuh"

#ifndef SPH_MINB
#define SPH_MINB 1 // minimum resident blocks per SM requested from ptxas for the list-walking sweeps
#endif

namespace SPH_NS {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- per-pair gradient cache (DFSPH) -------------------------------------------------------------------------
// Positions are frozen inside a step, so grad W_ij is too.  The two sweeps that need a neighbour's position AND
// velocity (k_df_drho, k_df_rho_adv: a 32-byte gather per pair, the costliest access pattern of the step) read
// [j | grad W_ij] records that k_build_lists wrote once per step instead -- a coalesced 16-byte stream -- and
// gather only the velocity.  The stored gradient is bit-identical to a recomputation (same function, same
// inputs), so the strict kernels stay bit-exact.  The sweeps that need one scalar of the neighbour keep the
// 16-byte (pos, payload) gather: for them the stream would cost more DRAM time than the arithmetic it saves.
__host__ __device__ inline size_t sph_gw_index(int s, int cap, int k) {
	return ((size_t)(s >> 5) * (size_t)cap + (size_t)k) * 32u + (size_t)(s & 31);
}
// The two streaming sweeps are HBM-bound on this stream (ncu: 19x their algorithmic bytes, 77 % of the HBM peak),
// so the FAST kernels shrink it: the index is already in the neighbour list (4 bytes), and the gradient is stored
// as three 21-bit fixed-point numbers over [-S, S], S = max |grad W| = kDW6 / (3 h) (8 bytes): 12 bytes per pair
// instead of 16, and k_build_lists writes 8 instead of 16.  Quantisation error <= S 2^-21 per component, i.e.
// 5e-7 of the largest gradient; measured per-sweep error against the strict kernels ~1e-6 (tolerance 1e-5).
#if !SPH_STRICT
__device__ __forceinline__ uint2 gq_pack(f3 g, const SphConsts &c) {
	// |g| <= S / 1.001 analytically, so the biased values stay inside 21 bits without a clamp (the mask only keeps a
	// non-finite gradient -- flagged elsewhere -- from spilling into its neighbours' bits)
	const float B = 1048576.0f;
	uint32_t mx = (uint32_t)__float2int_rn(fmaf(g.x, c.gq_inv, B)) & 0x1FFFFFu;
	uint32_t my = (uint32_t)__float2int_rn(fmaf(g.y, c.gq_inv, B)) & 0x1FFFFFu;
	uint32_t mz = (uint32_t)__float2int_rn(fmaf(g.z, c.gq_inv, B)) & 0x1FFFFFu;
	return make_uint2(mx | (my << 21), (my >> 11) | (mz << 10)); // bits 0-20 x, 21-41 y, 42-62 z
}
__device__ __forceinline__ f3 gq_unpack(uint32_t lo, uint32_t hi, float scale, float bias) {
	uint32_t mx = lo & 0x1FFFFFu, my = __funnelshift_r(lo, hi, 21) & 0x1FFFFFu, mz = (hi >> 10) & 0x1FFFFFu;
	return F3(fmaf((float)mx, scale, bias), fmaf((float)my, scale, bias), fmaf((float)mz, scale, bias));
}
__device__ __forceinline__ void gq_store(uint4 *gq, int s, int cap, int k, uint2 w) {
	uint2 *p = reinterpret_cast<uint2 *>(gq) + ((((size_t)(s >> 5) * (size_t)(cap >> 1) + (size_t)(k >> 1)) * 32u + (size_t)(s & 31)) << 1) + (size_t)(k & 1);
	*p = w;
}
#endif


// One thread per sorted particle.  (Slabs with overlap: the edge launch maps its threads through the edge list, the interior launch skips those
// particles; a skipped thread behaves like one past the end)
__device__ __forceinline__ int sph_split_slot(const SphLists &L, int n) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (L.split_mode == 1) s = s < L.n_edge ? L.edge_list[s] : n;
	else if (L.split_mode == 2 && s < n && ((L.edge_mask[s >> 5] >> (s & 31)) & 1u)) s = n;
	return s;
}

// decode the 1-D cell id (PS:102) back into (x, y, z)
__device__ __forceinline__ void cell_xyz(int cid, const SphConsts &c, int &cx, int &cy, int &cz) {
	cy = cid / c.gxz;
	int rem = cid - cy * c.gxz;
	cz = rem / c.gx;
	cx = rem - cz * c.gx;
}

__device__ __forceinline__ f3 ld3(const float *p) { return F3(p[0], p[1], p[2]); }

// velocity of a rigid particle as seen by the coupling terms (DF:168-169, 292-293; II:328-330):
// v_j = (vel + acc*dt) + cross(omega [+ alpha*dt], pos_j - centroid)
__device__ __forceinline__ f3 rigid_velocity(const SphRigidState *st, f3 pos_j, float dt, bool with_alpha) {
	f3 om = ld3(st->omega);
	if (with_alpha) om = om + ld3(st->alpha) * dt;
	f3 v_omega = cross(om, pos_j - ld3(st->centroid));
	return (ld3(st->vel) + ld3(st->acc) * dt) + v_omega;
}

static inline SphRigidArgs rigid_args(const SphHandle *h) {
	SphRigidArgs a;
	a.rspos = h->rspos;
	a.rstart = h->rg.cell_start;
	a.rsorted_id = h->rg.sorted_id;
	a.st = h->rstate;
	a.pos_orig = h->pos;
	a.slot_of = h->fg.slot_of;
	a.quirk = mg_rigid_quirk(h);
	a.gid = a.quirk ? h->gid : nullptr;
	a.active = (h->c.Nr > 0 && h->c.active_rigid && h->rigid_ready) ? 1 : 0;
	return a;
}

// Iterate the 27 cells around (cx,cy,cz) in the reference's order: ndrange((-1,2),(-1,2),(-1,2)),
// dx outermost, dz innermost (PS:452), skipping out-of-range cells (PS:453-456).
// The fast DFSPH kernels walk address-ordered lists (cell id = x + gx z + gx gz y ascending: dy outermost,
// dx innermost); any other consumer keeps the canonical order below.
#define SPH_FOR_27_ADDR(c, cx, cy, cz, C1)                                              \
	for (int dy_ = -1; dy_ <= 1; ++dy_)                                                 \
		for (int dz_ = -1; dz_ <= 1; ++dz_)                                             \
			for (int dx_ = -1; dx_ <= 1; ++dx_)                                         \
				if ((unsigned)((cx) + dx_) < (unsigned)(c).gx && (unsigned)((cy) + dy_) < (unsigned)(c).gy && \
				    (unsigned)((cz) + dz_) < (unsigned)(c).gz)                          \
					for (int C1 = ((cx) + dx_) + ((cy) + dy_) * (c).gxz + ((cz) + dz_) * (c).gx, once_ = 1; once_; once_ = 0)
#define SPH_FOR_27(c, cx, cy, cz, C1)                                                   \
	for (int dx_ = -1; dx_ <= 1; ++dx_)                                                 \
		for (int dy_ = -1; dy_ <= 1; ++dy_)                                             \
			for (int dz_ = -1; dz_ <= 1; ++dz_)                                         \
				if ((unsigned)((cx) + dx_) < (unsigned)(c).gx && (unsigned)((cy) + dy_) < (unsigned)(c).gy && \
				    (unsigned)((cz) + dz_) < (unsigned)(c).gz)                          \
					for (int C1 = ((cx) + dx_) + ((cy) + dy_) * (c).gxz + ((cz) + dz_) * (c).gx, once_ = 1; once_; once_ = 0)

// ---------------------------------------------------------------------------------------------
// Akinci boundary volumes, once at start-up (PS:309-320): V_b = 1 / sum_{b' != b} W
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_BLOCK) k_boundary_volume(SphConsts c, float4 *__restrict__ bspos,
                                                                const int *__restrict__ bscell,
                                                                const int *__restrict__ bstart,
                                                                const int *__restrict__ bsorted_id,
                                                                float4 *__restrict__ bpos_user) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.Nb) return;
	float4 pi = bspos[s];
	int cx, cy, cz;
	cell_xyz(bscell[s], c, cx, cy, cz);
	float volume = 0.0f;
	SPH_FOR_27(c, cx, cy, cz, c1) {
		int a = bstart[c1], b = bstart[c1 + 1];
		for (int e = a; e < b; ++e) {
			if (e == s) continue; // PS:362 (same material: skip self)
			float4 pj = bspos[e];
			Pair p = make_pair(pi, pj);
			if (culled(p, c)) continue;
			volume += cubic_w(p, c);
		}
	}
	float v = 1.0f / volume;
	// .w of neighbours is not read in this kernel, so the in-place update is race free
	bspos[s].w = v;
	bpos_user[bsorted_id[s]].w = v;
}

void boundary_volume(SphHandle *h, cudaStream_t st) {
	if (h->c.Nb <= 0) return;
	k_boundary_volume<<<cdiv(h->c.Nb, SPH_BLOCK), SPH_BLOCK, 0, st>>>(h->c, h->bspos, h->bg.scell, h->bg.cell_start,
	                                                                   h->bg.sorted_id, h->bpos);
	h->launches++;
}

// Cull the candidates [a, b) of one contiguous segment of a sorted array against pi, 32 at a time: the loop
// over candidates only builds a hit mask (branch-free body: with a ~15 % hit rate a conditional append
// inside it runs for almost every candidate of a warp at a few active lanes), then the mask is expanded in
// ascending order.  `self` is the slot to skip (PS:461) or NO_SELF.
#define NO_SELF (-0x40000000)
template <class Emit>
__device__ __forceinline__ void scan_segment(const SphConsts &c, const float4 &pi, const float4 *__restrict__ arr, int a,
                                             int b, int self, Emit &&emit, int limit = 0x7fffffff, int *err = nullptr) {
	SPH_BOUNDS_OK(a >= 0 && a <= b && b <= limit, err);
#if SPH_DEBUG_BOUNDS
	if (!(a >= 0 && a <= b && b <= limit)) return;
#endif
	for (int base = a; base < b; base += 32) {
		int len = min(32, b - base);
		uint32_t mask = 0;
		for (int k = 0; k < len; ++k) {
			Pair p = make_pair(pi, arr[base + k]);
			mask |= (uint32_t)(!culled(p, c)) << k; // PS:466 / PS:364
		}
		uint32_t d = (uint32_t)(self - base);
		if (d < 32u) mask &= ~(1u << d);
		while (mask) {
			int k = __ffs(mask) - 1;
			mask &= mask - 1;
			emit(base + k);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// k_build_lists: the only 27-cell traversal of a step.  Emits the fluid and boundary neighbour
// lists (canonical order), the neighbour count of get_neighbour_count (PS:424-445), rho
// (SB:41-72) and, for DFSPH, alpha (DF:32-89) -- all of which depend on positions only.
// ---------------------------------------------------------------------------------------------
template <bool ALPHA, bool RIGID>
#ifndef SPH_MINB_LISTS
#define SPH_MINB_LISTS 8
#endif
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB_LISTS) // 64 registers: this kernel is issue-bound and wants the warps
k_build_lists(SphConsts c, const float4 *__restrict__ spos, const float4 *__restrict__ svel,
              const int *__restrict__ scell, const int *__restrict__ cstart, const int *__restrict__ sorted_id,
              const float4 *__restrict__ bspos, const int *__restrict__ bstart, SphLists L, SphRigidArgs rg,
              int *__restrict__ nbr_count, float *__restrict__ rho, float *__restrict__ alpha,
              float4 *__restrict__ posR, float4 *__restrict__ posT1, SphCtl *ctl, SphMgPush pu) {
	int s = sph_split_slot(L, c.N);
	int nf = 0, nb = 0;
	mg_push_mark(pu);
	if (s < c.N && c.N != c.N_owned && sorted_id[s] >= c.N_owned) {
		// ghost copy of a neighbour rank's particle (multi-GPU): never a centre particle; its rho / alpha /
		// payloads arrive through the halo exchange.  fcount < 0 is the ownership flag every sweep tests.
		float4 pi = spos[s];
		L.fcount[s] = -1;
		L.bcount[s] = 0;
		nbr_count[s] = 0;
		if (!L.skip_ghost_fill) { // (overlap: the exchange beside this launch writes the complete records)
			posR[s] = make_float4(pi.x, pi.y, pi.z, 0.0f);
			if (ALPHA) posT1[s] = make_float4(pi.x, pi.y, pi.z, 0.0f);
		}
	} else if (s < c.N) {
		float4 pi = spos[s];
		int cx, cy, cz;
		cell_xyz(scell[s], c, cx, cy, cz);
		uint32_t *fl = L.flist, *bl = L.blist;
#define SPH_FLW(n) fl[sph_list_word(s, c.kstride, n)]
#define SPH_BLW(n) bl[sph_list_word(s, c.kbstride, n)]
		// ---- phase 1: the 27-cell traversal only culls and appends (a 15 % hit rate would otherwise run
		// ---- the kernel-function arithmetic at 15 % lane utilisation on every candidate) -----------------
		int ncount = 0; // get_neighbour_count (PS:424-445)
		int i_orig = RIGID ? (rg.gid ? rg.gid[sorted_id[s]] : sorted_id[s]) : 0; // the reference compares GLOBAL indices
		if (!RIGID) {
			// append cursors: the word of entry n+1 is 1 further, 125 further after a quad (sph_list_word)
			uint32_t *fcur = &SPH_FLW(0), *bcur = &SPH_BLW(0);
			auto emit_f = [&](int e) {
				if (nf < c.kmax) *fcur = (uint32_t)e;
				fcur += (nf & 3) == 3 ? 125 : 1;
				nf++;
			};
			auto emit_b = [&](int e) {
				if (nb < c.kbmax) *bcur = (uint32_t)e;
				bcur += (nb & 3) == 3 ? 125 : 1;
				nb++;
			};
#if SPH_STRICT
			// canonical order of the reference: cell by cell, (dx, dy, dz) with dz fastest
			SPH_FOR_27(c, cx, cy, cz, c1) { scan_segment(c, pi, spos, cstart[c1], cstart[c1 + 1], s, emit_f, c.N, L.err); }
			if (c.boundary_handle == 1)
				SPH_FOR_27(c, cx, cy, cz, c1) { scan_segment(c, pi, bspos, bstart[c1], bstart[c1 + 1], NO_SELF, emit_b, c.Nb, L.err); }
#else
			// The 27 cells are 9 runs of up to three x-adjacent cells, contiguous in the sorted arrays (cell id =
			// x + gx z + gx gz y).  The 18 run bounds of a grid are fetched up front (independent loads instead of
			// 27 dependent load pairs), then each run is one segment; the list comes out in address order.
			const int x0 = max(cx - 1, 0), xw = min(cx + 1, c.gx - 1) - x0 + 1;
			{
				int ra[9], rb[9];
#pragma unroll
				for (int r = 0; r < 9; ++r) {
					int y = cy + r / 3 - 1, z = cz + r % 3 - 1;
					bool ok = (unsigned)y < (unsigned)c.gy && (unsigned)z < (unsigned)c.gz;
					int c0 = x0 + y * c.gxz + z * c.gx;
					ra[r] = ok ? cstart[c0] : 0;
					rb[r] = ok ? cstart[c0 + xw] : 0;
				}
#pragma unroll
				for (int r = 0; r < 9; ++r) scan_segment(c, pi, spos, ra[r], rb[r], s, emit_f, c.N, L.err);
			}
			if (c.boundary_handle == 1) {
				int qa[9], qb[9];
#pragma unroll
				for (int r = 0; r < 9; ++r) {
					int y = cy + r / 3 - 1, z = cz + r % 3 - 1;
					bool ok = (unsigned)y < (unsigned)c.gy && (unsigned)z < (unsigned)c.gz;
					int c0 = x0 + y * c.gxz + z * c.gx;
					qa[r] = ok ? bstart[c0] : 0;
					qb[r] = ok ? bstart[c0 + xw] : 0;
				}
#pragma unroll
				for (int r = 0; r < 9; ++r) scan_segment(c, pi, bspos, qa[r], qb[r], NO_SELF, emit_b, c.Nb, L.err);
			}
#endif
			ncount = nf; // no rigid entries: get_neighbour_count (PS:424-445) equals the fluid hits
		} else {
			// scenes with an active rigid body: per-cell traversal, rigid particles after the fluid ones of a cell
#if SPH_STRICT
			SPH_FOR_27(c, cx, cy, cz, c1) {
#else
			SPH_FOR_27_ADDR(c, cx, cy, cz, c1) {
#endif
				int a = cstart[c1], b = cstart[c1 + 1];
				for (int e = a; e < b; ++e) {
					if (e == s) continue; // PS:461
					Pair p = make_pair(pi, spos[e]);
					if (culled(p, c)) continue; // PS:466
					if (nf < c.kmax) SPH_FLW(nf) = (uint32_t)e;
					nf++;
					ncount++;
				}
				// rigid particles follow the fluid ones inside a cell (second append kernel, PS:385-386)
				int ra = rg.rstart[c1], rb = rg.rstart[c1 + 1];
				for (int e = ra; e < rb; ++e) {
					// PS:440-442 quirk: the count compares the rigid-LOCAL index with i and measures the
					// distance to the FLUID particle with that index (SURVEY B-7)
					int k = rg.rsorted_id[e];
					if (k != i_orig) {
						Pair q = make_pair(pi, rg.quirk ? rg.quirk[k] : rg.pos_orig[min(k, c.N_owned - 1)]);
						if (!culled(q, c)) ncount++;
					}
					Pair p = make_pair(pi, rg.rspos[e]);
					if (culled(p, c)) continue;
					if (c.fs_couple != 1) continue; // SB:64: the tasks return 0
					if (nf < c.kmax) SPH_FLW(nf) = (uint32_t)e | SPH_RIGID_BIT;
					nf++;
				}
			}
			if (c.boundary_handle == 1) {
#if SPH_STRICT
				SPH_FOR_27(c, cx, cy, cz, c1) {
#else
				SPH_FOR_27_ADDR(c, cx, cy, cz, c1) {
#endif
					int a = bstart[c1], b = bstart[c1 + 1];
					for (int e = a; e < b; ++e) {
						Pair p = make_pair(pi, bspos[e]);
						if (culled(p, c)) continue; // PS:364
						if (nb < c.kbmax) SPH_BLW(nb) = (uint32_t)e;
						nb++;
					}
				}
			}
		}
		int nfl = min(nf, c.kmax), nbl = min(nb, c.kbmax);
#if SPH_STRICT
#define SPH_GRAD_STORE(k, j, dw) if (L.gw) L.gw[sph_gw_index(s, c.kstride, k)] = make_float4(__uint_as_float(j), (dw).x, (dw).y, (dw).z)
#else
#define SPH_GRAD_STORE(k, j, dw) if (L.gq) gq_store(L.gq, s, c.kstride, k, gq_pack(dw, c))
#endif
		// ---- phase 2: walk the fresh lists (canonical order, every lane busy): rho (SB:41-72), alpha (DF:32-89)
		float rho_f = 0.001f; // SB:44
		f3 ss = F3(0.0f, 0.0f, 0.0f);
		float sq = 0.0f;
		for (int k = 0; k < nfl; ++k) {
			uint32_t j = SPH_FLW(k);
			if (RIGID && (j & SPH_RIGID_BIT)) {
				float4 pj = rg.rspos[j & ~SPH_RIGID_BIT];
				Pair p = make_pair(pi, pj);
				rho_f += (pj.w * cubic_w(p, c)) * SPH_RHO0; // SB:65
				if (ALPHA) {
					f3 dw = cubic_dw(p, c);
					SPH_GRAD_STORE(k, j, dw);
					f3 g = (pj.w * SPH_RHO0) * dw; // DF:62, 75
					ss = ss + g;
					sq += dot(g, g);
				}
				continue;
			}
			Pair p = make_pair(pi, spos[j]);
			rho_f += c.m * cubic_w(p, c); // SB:62
			if (ALPHA) {
				f3 dw = cubic_dw(p, c);
				SPH_GRAD_STORE(k, j, dw);
				f3 g = c.m * dw; // DF:58, 70
				ss = ss + g;
				sq += dot(g, g);
			}
		}
		float rho_i = rho_f;
		float den = dot(ss, ss) + sq;
		if (c.boundary_handle == 1) {
			float rho_b = 0.0f;
			f3 ssb = F3(0.0f, 0.0f, 0.0f);
			float sqb = 0.0f;
			for (int k = 0; k < nbl; ++k) {
				float4 pj = bspos[SPH_BLW(k)];
				Pair p = make_pair(pi, pj);
				rho_b += pj.w * cubic_w(p, c); // SB:71
				if (ALPHA) {
					f3 g = (pj.w * SPH_RHO0) * cubic_dw(p, c); // DF:82, 88
					ssb = ssb + g;
					sqb += dot(g, g);
				}
			}
			rho_i = rho_f + rho_b * SPH_RHO0; // SB:49
			den = ((dot(ss, ss) + sq) + sqb) + dot(ssb, ssb); // DF:45
		}
		L.fcount[s] = nfl;
		L.bcount[s] = nbl;
		nbr_count[s] = ncount;
		rho[s] = rho_i;
		posR[s] = make_float4(pi.x, pi.y, pi.z, rho_i);
		if (ALPHA) {
			float al = fabsf(den) < 1e-6f ? 0.0f : rho_i / den; // DF:48-51
			alpha[s] = al;
			// payload of the warm start (DF:333-337): (k / dt) / rho
			float k = svel[s].w;
			float t1 = (k / ctl->dt) / rho_i;
			posT1[s] = make_float4(pi.x, pi.y, pi.z, t1);
			mg_push(pu, s, t1, rho_i, 0.0f); // slabs: (warm-start payload, rho) of an edge particle, MG_F4_T1R
		}
#undef SPH_FLW
#undef SPH_BLW
#undef SPH_GRAD_STORE
	}
	int mf = warp_max_i(nf), mb = warp_max_i(nb);
	if ((threadIdx.x & 31) == 0) {
		if (mf > c.kmax) atomicOr(&ctl->error_flags, SPH_ERR_LIST_OVERFLOW);
		if (mb > c.kbmax) atomicOr(&ctl->error_flags, SPH_ERR_BLIST_OVERFLOW);
		if (mf > ctl->max_nbr) atomicMax(&ctl->max_nbr, mf);
		if (mb > ctl->max_bnbr) atomicMax(&ctl->max_bnbr, mb);
	}
}

// One sweep + the exchange of the values it produces for the neighbour ranks (`what`, MG_NONE: none).
// `launch(grid, stream, lists, push)` issues the kernel.  Classic: one launch over all particles, then the exchange.
// Slabs with overlap (mg_split): the edge particles first, on the exchange stream, with the exchange right behind
// them, while the interior runs on the main stream; the main stream joins before anything reads the ghost values.
// Returns the number of block partials the sweep wrote (interior blocks, then edge blocks); *exchanged tells the
// caller whether the data exchange is done (a sweep with a loop decision still owes the reduction).
template <class Launch>
static int sweep(SphHandle *h, int what, int kc, cudaStream_t st, bool exchange_here, bool *exchanged, Launch &&launch) {
	const int nb = cdiv(h->c.N, SPH_BLOCK);
	MgSplit sp = mg_split(h);
	SphMgPush nopush;
	memset(&nopush, 0, sizeof(nopush));
	if (!sp.on) {
		sph_prof_begin(h, kc, st);
		launch(nb, st, h->L, what == MG_NONE ? nopush : mg_push_args(h));
		sph_prof_end(h, st);
		h->launches += 1;
		if (exchange_here && what != MG_NONE) mg_exchange(h, what, st);
		if (exchanged) *exchanged = exchange_here;
		return nb;
	}
	SphLists Le = h->L, Li = h->L;
	Le.split_mode = 1; Le.edge_list = sp.edge_list; Le.n_edge = sp.n_edge; Le.partial_offset = nb;
	Li.split_mode = 2; Li.edge_mask = sp.edge_mask; Li.skip_ghost_fill = what != MG_NONE ? 1 : 0;
	const int nbe = cdiv(sp.n_edge, SPH_BLOCK);
	mg_fork(h, st);
	if (nbe > 0) launch(nbe, sp.xs, Le, nopush);
	if (what != MG_NONE) mg_exchange(h, what, sp.xs);
	sph_prof_begin(h, kc, st);
	launch(nb, st, Li, nopush);
	sph_prof_end(h, st);
	mg_join(h, st);
	h->launches += 1 + (nbe > 0 ? 1 : 0);
	if (exchanged) *exchanged = true;
	return nb + nbe;
}
void build_lists(SphHandle *h, cudaStream_t st, bool exchange) {
	const SphConsts &c = h->c;
	if (c.N <= 0) { // an empty slab still takes part in the exchange (its neighbours wait for its handshake)
		if (exchange) mg_exchange(h, MG_F4_T1R, st);
		return;
	}
	SphRigidArgs rg = rigid_args(h);
	bool al = c.solver == SPH_SOLVER_DFSPH;
	// `exchange`: the first phase of a step follows the build with the exchange of (payload, rho) of the edge
	// particles (MG_F4_T1R); on slabs with overlap that exchange runs behind the interior part of the build
	sweep(h, exchange ? MG_F4_T1R : MG_NONE, KC_LISTS, st, true, nullptr,
	      [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu_) {
		SphMgPush pu = pu_;
		if (!al) memset(&pu, 0, sizeof(pu)); // the other solvers' T1R payload is not written by this kernel
#define SPH_BL(A, R)                                                                                              \
	k_build_lists<A, R><<<grid, SPH_BLOCK, 0, s>>>(c, h->a4[A4_POS], h->a4[A4_VEL], h->fg.scell, h->fg.cell_start, \
	                                               h->fg.sorted_id, h->bspos, h->bg.cell_start, L, rg,             \
	                                               h->nbr_count, h->a1[A1_RHO], h->a1[A1_ALPHA], h->a4[A4_PR],     \
	                                               h->a4[A4_T1], h->ctl, pu)
		if (al && rg.active) SPH_BL(true, true);
		else if (al) SPH_BL(true, false);
		else if (rg.active) SPH_BL(false, true);
		else SPH_BL(false, false);
#undef SPH_BL
	});
	h->lists_valid = true;
	if (rg.active) mg_rigid_quirk_update(h, 1, st); // slabs: rho of the quirk particles is known now
}

// Every solver's first phase starts with the step's lists (and the ghosts' density on slabs); a caller that
// drives single sweeps has already built them with SPH_PH_BUILD_LISTS.
void first_phase_lists(SphHandle *h, cudaStream_t st) {
	if (!h->lists_fresh) build_lists(h, st, true); // slabs: + rho (posR.w) and the payload of the ghost particles
	h->lists_fresh = false;
	sph_finish_deferred_vel(h, st); // e2e path: the velocities travelled behind the grid and list build (sph_api.cu)
}

// list walkers ---------------------------------------------------------------------------------
// Lists are quad-interleaved (sph_list_word): one 128-bit load brings four entries of a particle and a
// warp reads 512 contiguous bytes.  The list stream comes from DRAM (it is larger than L2), so the
// walker keeps two quads in flight ahead of the one being consumed; entries bypass L1 allocation
// (LDG.E.NA) and are marked evict-first in L2 so that the stream does not displace the particle
// records the gathers hit.  The four bodies of a full quad are straight-line code, so ptxas issues
// their gathers back to back; the tail (count % 4 entries) is guarded.
#ifndef SPH_LIST_HINTS
#define SPH_LIST_HINTS 2 // 0: plain ld.global.nc, 1: + L1::no_allocate, 2: + L2 evict-first policy
#endif
__device__ __forceinline__ uint64_t list_policy() {
	uint64_t pol = 0;
#if SPH_LIST_HINTS == 2
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
	return pol;
}
__device__ __forceinline__ uint4 ld_list(const uint4 *p, uint64_t pol) {
	uint4 r;
#if SPH_LIST_HINTS == 2
	asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
	    : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	    : "l"(p), "l"(pol));
#elif SPH_LIST_HINTS == 1
	asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
	asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
	return r;
}
struct ListRange {
	const uint4 *p; // quad 0 of this lane
	int n;          // entries of this lane
#if SPH_DEBUG_BOUNDS
	int lim, lim_rigid, *err;
#endif
};
// bounds-checked build: entry j of a list must address an existing particle of the array it indexes
#if SPH_DEBUG_BOUNDS
__device__ __forceinline__ bool sph_entry_ok(uint32_t j, int lim, int lim_rigid, int *err) {
	bool ok = (j & SPH_RIGID_BIT) ? (int)(j & ~SPH_RIGID_BIT) < lim_rigid : (int)j < lim;
	if (!ok) atomicOr(err, SPH_ERR_BOUNDS);
	return ok; // an invalid entry is reported and skipped, never dereferenced
}
#endif
__device__ __forceinline__ ListRange list_range(const SphLists &L, bool boundary, int cap, int s, int n) {
	ListRange r;
	r.p = reinterpret_cast<const uint4 *>(boundary ? L.blist : L.flist) + ((size_t)(s >> 5) * (size_t)(cap >> 2)) * 32u + (size_t)(s & 31);
	r.n = n;
#if SPH_DEBUG_BOUNDS
	r.lim = boundary ? L.n_boundary : L.n_fluid;
	r.lim_rigid = boundary ? 0 : L.n_rigid;
	r.err = L.err;
	SPH_BOUNDS_OK(n <= (boundary ? L.cap_b : L.cap_f) && cap == (boundary ? L.cap_b : L.cap_f), L.err);
#endif
	return r;
}
template <class F>
__device__ __forceinline__ void operator<<(ListRange r, F &&f) {
	if (r.n <= 0) return;
	const uint64_t pol = list_policy();
	const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
	const uint4 *p = r.p;
	uint4 cur = ld_list(p, pol);
	uint4 nx1 = r.n > 4 ? ld_list(p + 32, pol) : zero;
	int k = 0;
#if SPH_DEBUG_BOUNDS
#define SPH_CALL(j) do { if (sph_entry_ok(j, r.lim, r.lim_rigid, r.err)) f(j); } while (0)
#else
#define SPH_CALL(j) f(j)
#endif
	for (; k + 4 <= r.n; k += 4) {
		uint4 nx2 = k + 8 < r.n ? ld_list(p + 64, pol) : zero;
		p += 32;
		SPH_CALL(cur.x); SPH_CALL(cur.y); SPH_CALL(cur.z); SPH_CALL(cur.w);
		cur = nx1;
		nx1 = nx2;
	}
	int m = r.n - k;
	if (m > 0) {
		SPH_CALL(cur.x);
		if (m > 1) {
			SPH_CALL(cur.y);
			if (m > 2) SPH_CALL(cur.z);
		}
	}
#undef SPH_CALL
}
// usage:  SPH_FOR_FLUID(L, c, s, j) { ...body, `return` skips to the next neighbour... };
#define SPH_FOR_FLUID(L, c, s, J) list_range(L, false, (c).kstride, s, (L).fcount[s]) << [&](uint32_t J)
#define SPH_FOR_BOUNDARY(L, c, s, J) list_range(L, true, (c).kbstride, s, (L).bcount[s]) << [&](uint32_t J)
// the same with the count already in a register
#define SPH_FOR_FLUID_N(L, c, s, n, J) list_range(L, false, (c).kstride, s, n) << [&](uint32_t J)
#define SPH_FOR_BOUNDARY_N(L, c, s, n, J) list_range(L, true, (c).kbstride, s, n) << [&](uint32_t J)

__device__ __forceinline__ float4 ld_gw(const float4 *p, uint64_t pol) {
	float4 r;
#if SPH_LIST_HINTS == 2
	asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
	    : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
	    : "l"(p), "l"(pol));
#else
	asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
#endif
	return r;
}
// f(j, grad W_ij) for the n entries of sorted particle s; four records in flight ahead of the four in use
template <class F>
__device__ __forceinline__ void walk_gw(const SphLists &L, const float4 *__restrict__ gw, int cap, int s, int n, F &&f_) {
	if (n <= 0) return;
#if SPH_DEBUG_BOUNDS
	if (n > L.cap_f || cap != L.cap_f) { atomicOr(L.err, SPH_ERR_BOUNDS); return; }
	auto f = [&](uint32_t j, f3 dw) { if (sph_entry_ok(j, L.n_fluid, L.n_rigid, L.err)) f_(j, dw); };
#else
	F &f = f_;
#endif
	const uint64_t pol = list_policy();
	const float4 *p = gw + sph_gw_index(s, cap, 0);
	const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float4 c0 = ld_gw(p, pol), c1 = n > 1 ? ld_gw(p + 32, pol) : z, c2 = n > 2 ? ld_gw(p + 64, pol) : z,
	       c3 = n > 3 ? ld_gw(p + 96, pol) : z;
	int k = 0;
	for (; k + 4 <= n; k += 4) {
		p += 128;
		float4 n0 = k + 4 < n ? ld_gw(p, pol) : z, n1 = k + 5 < n ? ld_gw(p + 32, pol) : z,
		       n2 = k + 6 < n ? ld_gw(p + 64, pol) : z, n3 = k + 7 < n ? ld_gw(p + 96, pol) : z;
		f(__float_as_uint(c0.x), F3(c0.y, c0.z, c0.w));
		f(__float_as_uint(c1.x), F3(c1.y, c1.z, c1.w));
		f(__float_as_uint(c2.x), F3(c2.y, c2.z, c2.w));
		f(__float_as_uint(c3.x), F3(c3.y, c3.z, c3.w));
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
	}
	int m = n - k;
	if (m > 0) {
		f(__float_as_uint(c0.x), F3(c0.y, c0.z, c0.w));
		if (m > 1) {
			f(__float_as_uint(c1.x), F3(c1.y, c1.z, c1.w));
			if (m > 2) f(__float_as_uint(c2.x), F3(c2.y, c2.z, c2.w));
		}
	}
}

// (A variant that staged the float4 records through shared memory with TMA bulk copies -- cp.async.bulk + mbarrier,
// a per-warp ring of 2 KB stages -- was measured slower on B200, 111-124 us against 107 us: the ring takes L1 away
// from the velocity gathers.  profiles/r1d_experiments.md section 9; removed in round 2.)

#ifndef SPH_GQ_AHEAD
#define SPH_GQ_AHEAD 2 // quads of (indices, gradients) in flight ahead of the one in use
#endif
#ifndef SPH_MINB_STREAM
// minimum resident blocks per SM of the two streaming sweeps (k_df_drho, k_df_rho_adv).  They are latency-bound on the
// stream: measured on B200 at 10^6 particles (profiles/r2_experiments.md) k_df_drho 108 us at 110 registers (4 blocks),
// 99 us at 96 (5), 90 us at 80 (6), 89 us at 72 (7), 107 us at 64 (8: spills).  Strict kernels keep ptxas' choice.
#if SPH_STRICT
#define SPH_MINB_STREAM SPH_MINB
#else
#define SPH_MINB_STREAM 7
#endif
#endif
#if !SPH_STRICT
// fast kernels: f(j, grad W_ij) from the neighbour list (index) + the quantised gradient stream; per quad of
// entries three coalesced 128-bit loads (one of indices, two of gradients), two quads in flight ahead of the one
// in use, L1 no-allocate, L2 evict-first
struct GqQuad {
	uint4 j, a, b;
};
template <class F>
__device__ __forceinline__ void walk_gq(const SphLists &L, const SphConsts &c, int s, int n, F &&f_) {
	if (n <= 0) return;
#if SPH_DEBUG_BOUNDS
	if (n > L.cap_f || c.kstride != L.cap_f) { atomicOr(L.err, SPH_ERR_BOUNDS); return; }
	auto f = [&](uint32_t j, f3 dw) { if (sph_entry_ok(j, L.n_fluid, L.n_rigid, L.err)) f_(j, dw); };
#else
	F &f = f_;
#endif
	const uint64_t pol = list_policy();
	const uint4 *pl = reinterpret_cast<const uint4 *>(L.flist) + ((size_t)(s >> 5) * (size_t)(c.kstride >> 2)) * 32u + (size_t)(s & 31);
	const uint4 *pg = L.gq + ((size_t)(s >> 5) * (size_t)(c.kstride >> 1)) * 32u + (size_t)(s & 31);
	const float scale = c.gq_scale, bias = -1048576.0f * c.gq_scale;
	const uint4 z = make_uint4(0u, 0u, 0u, 0u);
	auto load = [&](const uint4 *l, const uint4 *g) {
		GqQuad r;
		r.j = ld_list(l, pol);
		r.a = ld_list(g, pol);
		r.b = ld_list(g + 32, pol);
		return r;
	};
	GqQuad cur = load(pl, pg), nx1;
	if (n > 4) nx1 = load(pl + 32, pg + 64);
	else { nx1.j = z; nx1.a = z; nx1.b = z; }
	int k = 0;
	for (; k + 4 <= n; k += 4) {
#if SPH_GQ_AHEAD == 2
		GqQuad nx2;
		if (k + 8 < n) nx2 = load(pl + 64, pg + 128);
		else { nx2.j = z; nx2.a = z; nx2.b = z; }
#endif
		pl += 32;
		pg += 64;
		f(cur.j.x, gq_unpack(cur.a.x, cur.a.y, scale, bias));
		f(cur.j.y, gq_unpack(cur.a.z, cur.a.w, scale, bias));
		f(cur.j.z, gq_unpack(cur.b.x, cur.b.y, scale, bias));
		f(cur.j.w, gq_unpack(cur.b.z, cur.b.w, scale, bias));
		cur = nx1;
#if SPH_GQ_AHEAD == 2
		nx1 = nx2;
#else
		if (k + 8 < n) nx1 = load(pl + 32, pg + 64);
#endif
	}
	int m = n - k;
	if (m > 0) {
		f(cur.j.x, gq_unpack(cur.a.x, cur.a.y, scale, bias));
		if (m > 1) {
			f(cur.j.y, gq_unpack(cur.a.z, cur.a.w, scale, bias));
			if (m > 2) f(cur.j.z, gq_unpack(cur.b.x, cur.b.y, scale, bias));
		}
	}
}
#define SPH_WALK_GW(L, c, s, n, ...) walk_gq(L, c, s, n, __VA_ARGS__)
#else
#define SPH_WALK_GW(L, c, s, n, ...) walk_gw(L, (L).gw, (c).kstride, s, n, __VA_ARGS__)
#endif

// =============================================================================================
// DFSPH (dfsph_solver.py)
// =============================================================================================

// A rigid neighbour entry of the fluid list: sorted rigid slot with SPH_RIGID_BIT set.
#define SPH_IS_RIGID(j) (RIGID && ((j) & SPH_RIGID_BIT))
#define SPH_RIGID_SLOT(j) ((j) & ~SPH_RIGID_BIT)

// ---- DFSPH sweeps: one thread per sorted particle; ghost copies (multi-GPU) carry fcount < 0 -------------
#define SPH_DF_THREAD()                                        \
	const int s = sph_split_slot(L, c.N);                      \
	const int nf_ = s < c.N ? L.fcount[s] : -1;                \
	const bool live = nf_ >= 0;                                \
	const int nb_ = live ? L.bcount[s] : 0;                    \
	(void)nb_

// DF:314-355 divergence_warm_start.  Reads neighbour payload t1 = (k/dt)/rho from posT1.w.
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB)
k_df_warm_start(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT1,
                const float4 *__restrict__ bspos, const float *__restrict__ rho, float4 *__restrict__ svel,
                const SphCtl *__restrict__ ctl, SphMgPush pu) {
	SPH_DF_THREAD();
	mg_push_mark(pu);
	float dt = ctl->dt;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vi = pi;
	float rho_i = 1.0f;
	if (live) { pi = posT1[s]; vi = svel[s]; rho_i = rho[s]; }
	float k_i = vi.w / dt; // DF:333, 342, 353
	f3 va = F3(0.0f, 0.0f, 0.0f);
#if SPH_STRICT
	// strict kernels: IEEE sqrt / division make the gradient the costly part, so every DFSPH sweep reads it
	// from the per-pair cache (bit-identical to a recomputation) and gathers only the neighbour's payload
	walk_gw(L, L.gw, c.kstride, s, nf_, [&](uint32_t j, f3 dw) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * dw; // DF:345
			return;
		}
		va = va + (c.m * (pi.w + __ldg(&posT1[j]).w)) * dw; // DF:337
	});
#else
	SPH_FOR_FLUID_N(L, c, s, nf_, j) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			Pair p = make_pair(pi, pj);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * cubic_dw(p, c); // DF:345
			return;
		}
		float4 pj = __ldg(&posT1[j]);
		Pair p = make_pair(pi, pj);
		va = va + (c.m * (pi.w + pj.w)) * cubic_dw(p, c); // DF:337
	};
#endif
	f3 vb = F3(0.0f, 0.0f, 0.0f);
	if (c.boundary_handle == 1) {
		SPH_FOR_BOUNDARY_N(L, c, s, nb_, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			vb = vb + ((pj.w * k_i) / rho_i) * cubic_dw(p, c); // DF:354
		};
	}
	if (!live) return;
	f3 v = xyz(vi);
	if (c.boundary_handle == 1) v = v - (va + vb * SPH_RHO0) * dt; // DF:322
	else v = v - va * dt;                                          // DF:324
	svel[s] = F4(v, 0.0f); // DF:325 warm_start_k.fill(0)
	mg_push(pu, s, v.x, v.y, v.z);
}

// DF:252-300 derivative_iter_all_rho.  Writes drho and the payload t2 = ((drho*alpha)/dt)/rho of
// the following divergence iteration (DF:363-367); block partials feed the device-side average.
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB_STREAM)
k_df_drho(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ spos,
          const float4 *__restrict__ svel, const float4 *__restrict__ bspos,
          const int *__restrict__ nbr_count,
          const float *__restrict__ rho, const float *__restrict__ alpha, float *__restrict__ drho,
          float4 *__restrict__ posT2, const SphCtl *__restrict__ ctl, SphPartial *__restrict__ partials, int gated,
          SphMgPush pu) {
	if (gated && !ctl->div_active) return;
	SPH_DF_THREAD();
	mg_push_mark(pu);
	double psum = 0.0;
	int pcnt = 0;
	float dt = ctl->dt;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	f3 vi = F3(0.0f, 0.0f, 0.0f);
	bool enough = false;
	if (live) {
		pi = spos[s];
		vi = xyz(svel[s]);
		enough = nbr_count[s] >= 20; // DF:258-261
	}
	float rd = 0.0f, rdb = 0.0f;
	if (enough) {
		SPH_WALK_GW(L, c, s, nf_, [&](uint32_t j, f3 dw) {
			if (SPH_IS_RIGID(j)) {
				float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
				f3 v_j = rigid_velocity(rg.st, xyz(pj), dt, false);             // DF:292-293
				rd += (pj.w * SPH_RHO0) * dot(vi - v_j, dw);                    // DF:294
				return;
			}
			rd += c.m * dot(vi - xyz(__ldg(&svel[j])), dw); // DF:287
		});
		if (c.boundary_handle == 1) {
			SPH_FOR_BOUNDARY_N(L, c, s, nb_, j) {
				float4 pj = __ldg(&bspos[j]);
				Pair p = make_pair(pi, pj);
				rdb += pj.w * dot(vi, cubic_dw(p, c)); // DF:300
			};
		}
	}
	if (live) {
		float out = 0.0f;
		if (enough) out = c.boundary_handle == 1 ? fmaxf(rd + rdb * SPH_RHO0, 0.0f) : fmaxf(rd, 0.0f); // DF:267
		drho[s] = out;
		float t2 = ((out * alpha[s]) / dt) / rho[s];
		posT2[s] = make_float4(pi.x, pi.y, pi.z, t2);
		mg_push(pu, s, t2, 0.0f, 0.0f);
		if (out > 0.0f) { psum = (double)out; pcnt = 1; } // DF:275-277
	}
	block_partial(psum, pcnt, 0.0f, partials + L.partial_offset);
}

// DF:302-312, 357-391 divergence_iter_all_vel_adv fused with DF:381-384 sum_up_stiff
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB)
k_df_div_iter(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT2,
              const float4 *__restrict__ bspos, const float *__restrict__ rho, const float *__restrict__ alpha,
              const float *__restrict__ drho, float4 *__restrict__ svel, const SphCtl *__restrict__ ctl, SphMgPush pu) {
	if (!ctl->div_active) return;
	SPH_DF_THREAD();
	mg_push_mark(pu);
	float dt = ctl->dt;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float da = 0.0f, rho_i = 1.0f;
	if (live) { pi = posT2[s]; da = drho[s] * alpha[s]; rho_i = rho[s]; }
	float k_i = da / dt; // DF:363, 374, 388
	f3 va = F3(0.0f, 0.0f, 0.0f);
#if SPH_STRICT
	walk_gw(L, L.gw, c.kstride, s, nf_, [&](uint32_t j, f3 dw) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * dw; // DF:377
			return;
		}
		float f = pi.w + __ldg(&posT2[j]).w;
		if (f > 1e-5f) va = va + (c.m * f) * dw; // DF:367-369
	});
#else
	SPH_FOR_FLUID_N(L, c, s, nf_, j) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			Pair p = make_pair(pi, pj);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * cubic_dw(p, c); // DF:377
			return;
		}
		float4 pj = __ldg(&posT2[j]);
		Pair p = make_pair(pi, pj);
		float f = pi.w + pj.w;
		f3 dw = cubic_dw(p, c);
		if (f > 1e-5f) va = va + (c.m * f) * dw; // DF:367-369 (a select instead of the branch measured 4 us slower)
	};
#endif
	f3 vb = F3(0.0f, 0.0f, 0.0f);
	if (c.boundary_handle == 1) {
		SPH_FOR_BOUNDARY_N(L, c, s, nb_, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			vb = vb + ((pj.w * k_i) / rho_i) * cubic_dw(p, c); // DF:390
		};
	}
	if (!live) return;
	float4 vi = svel[s];
	f3 v = xyz(vi);
	if (c.boundary_handle == 1) v = v - (va + vb * SPH_RHO0) * dt; // DF:310
	else v = v - va * dt;
	svel[s] = F4(v, vi.w + da); // DF:384
	mg_push(pu, s, v.x, v.y, v.z);
}

// SB:190-201: viscosity contribution of a rigid neighbour (uses rho[particle_j.index], SURVEY B-6)
template <bool RIGID>
__device__ __forceinline__ void rigid_viscosity(const SphConsts &c, const SphRigidArgs &rg, uint32_t j,
                                                const float4 &pi, const f3 &vi, float rho_i,
                                                const float *__restrict__ rho, f3 &visc) {
	uint32_t r = SPH_RIGID_SLOT(j);
	float4 pj = __ldg(&rg.rspos[r]);
	Pair p = make_pair(pi, pj);
	f3 v_ij = vi - ld3(rg.st->vel);
	float shear = dot(v_ij, p.r);
	if (shear < 0.0f) {
		float q = sqrtf(p.r2);
		float q2 = q * q;
		int jr = min(rg.rsorted_id[r], c.N_owned - 1);
		float rho_q = rg.quirk ? rg.quirk[rg.rsorted_id[r]].w : rho[rg.slot_of[jr]];
		float nu = c.visc_num / (rho_i + rho_q);
		float pi_ij = ((-nu) * shear) / (q2 + c.visc_eps_h2);
		visc = visc + ((-SPH_RHO0 * pj.w) * pi_ij) * cubic_dw(p, c); // SB:201
	}
}

// DF:91-122: tension (SB:204-217) + viscosity (SB:170-202) + f_ext + v* = v + dt f / m, and the
// block maxima of |v*| for the adaptive time step.
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB)
k_df_ext_force(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posR,
               const float4 *__restrict__ svel, const float *__restrict__ rho, float4 *__restrict__ svadv,
               float4 *__restrict__ fext, const SphCtl *__restrict__ ctl, SphPartial *__restrict__ partials, SphMgPush pu) {
	SPH_DF_THREAD();
	mg_push_mark(pu);
	float vmax = -INFINITY;
	float dt = ctl->dt;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
	f3 vi = F3(0.0f, 0.0f, 0.0f);
	if (live) { pi = posR[s]; vi = xyz(svel[s]); }
	f3 ten = F3(0.0f, 0.0f, 0.0f), visc = F3(0.0f, 0.0f, 0.0f);
	SPH_FOR_FLUID_N(L, c, s, nf_, j) {
		if (SPH_IS_RIGID(j)) {
			rigid_viscosity<RIGID>(c, rg, j, pi, vi, pi.w, rho, visc);
			return;
		}
		float4 pj = __ldg(&posR[j]);
		f3 vj = xyz(__ldg(&svel[j]));
		Pair p = make_pair(pi, pj);
		ten = ten + (c.tension_coef * cubic_w(p, c)) * p.r; // SB:216
		f3 v_ij = vi - vj;
		float shear = dot(v_ij, p.r); // SB:183
		if (shear < 0.0f) {
#if SPH_STRICT
			float q = sqrtf(p.r2);
			float q2 = q * q;
#else
			float q2 = p.r2;
#endif
			float nu = c.visc_num / (pi.w + pj.w);                 // SB:187
			float pi_ij = ((-nu) * shear) / (q2 + c.visc_eps_h2);  // SB:188
			visc = visc + (c.neg_m * pi_ij) * cubic_dw(p, c);      // SB:189
		}
	};
	if (live) {
		f3 tension = ten * c.m;  // SB:209
		f3 viscosity = visc * c.m; // SB:175
		f3 g = F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
		f3 f = (g + tension) + viscosity; // DF:96
		f3 va = vi + (dt * f) / c.m;       // DF:102
		fext[s] = F4(f, 0.0f);
		svadv[s] = F4(va, 0.0f);
		mg_push(pu, s, va.x, va.y, va.z);
		vmax = sqrtf(dot(va, va)); // DF:103
	}
	block_partial(0.0, 0, vmax, partials + L.partial_offset);
}

// DF:124-176 compute_all_rho_adv.  Writes rho_adv and the payload t3 = (((rho_adv-rho0)*alpha)/dt2)/rho
// of iter_all_vel_adv (DF:199-203).
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB_STREAM)
k_df_rho_adv(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ spos,
             const float4 *__restrict__ svadv, const float4 *__restrict__ bspos, const float *__restrict__ rho,
             const float *__restrict__ alpha, float *__restrict__ rho_adv, float4 *__restrict__ posT3,
             const SphCtl *__restrict__ ctl, SphPartial *__restrict__ partials, int gated, SphMgPush pu) {
	if (gated && !ctl->den_active) return;
	SPH_DF_THREAD();
	mg_push_mark(pu);
	double psum = 0.0;
	int pcnt = 0;
	float dt = ctl->dt, dt2 = ctl->dt2;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	f3 vi = F3(0.0f, 0.0f, 0.0f);
	if (live) { pi = spos[s]; vi = xyz(svadv[s]); }
	float delta = 0.0f, db = 0.0f;
	SPH_WALK_GW(L, c, s, nf_, [&](uint32_t j, f3 dw) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			f3 v_j = rigid_velocity(rg.st, xyz(pj), dt, true);                  // DF:168-169
			delta += (pj.w * SPH_RHO0) * dot(vi - v_j, dw);                     // DF:170
			return;
		}
		delta += c.m * dot(vi - xyz(__ldg(&svadv[j])), dw); // DF:162
	});
	if (c.boundary_handle == 1) {
		SPH_FOR_BOUNDARY_N(L, c, s, nb_, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			db += pj.w * dot(vi, cubic_dw(p, c)); // DF:176
		};
	}
	if (live) {
		float rho_i = rho[s];
		float ra;
		if (c.boundary_handle == 1) ra = fmaxf(rho_i + dt * (delta + db * SPH_RHO0), SPH_RHO0); // DF:135
		else ra = fmaxf(rho_i + dt * delta, SPH_RHO0);                                          // DF:137
		rho_adv[s] = ra;
		float t3 = (((ra - SPH_RHO0) * alpha[s]) / dt2) / rho_i;
		posT3[s] = make_float4(pi.x, pi.y, pi.z, t3);
		mg_push(pu, s, t3, 0.0f, 0.0f);
		if (!(ra == SPH_RHO0)) { psum = (double)ra; pcnt = 1; } // DF:139-141
	}
	block_partial(psum, pcnt, 0.0f, partials + L.partial_offset);
}

// DF:178-219 iter_all_vel_adv (fluid + boundary part; the rigid force scatter DF:212 is the gather
// kernel k_rigid_force_df in sph_rigid.cuh)
template <bool RIGID>
__global__ void __launch_bounds__(SPH_BLOCK, SPH_MINB)
k_df_vel_adv_iter(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT3,
                  const float4 *__restrict__ bspos, const float *__restrict__ rho, const float *__restrict__ alpha,
                  const float *__restrict__ rho_adv, float4 *__restrict__ svadv, const SphCtl *__restrict__ ctl,
                  int gated, SphMgPush pu) {
	if (gated && !ctl->den_active) return;
	SPH_DF_THREAD();
	mg_push_mark(pu);
	float dt = ctl->dt, dt2 = ctl->dt2;
	float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float rho_i = 1.0f, k_i = 0.0f;
	if (live) {
		pi = posT3[s];
		rho_i = rho[s];
		k_i = ((rho_adv[s] - SPH_RHO0) * alpha[s]) / dt2; // DF:199, 208, 217
	}
	f3 va = F3(0.0f, 0.0f, 0.0f);
#if SPH_STRICT
	walk_gw(L, L.gw, c.kstride, s, nf_, [&](uint32_t j, f3 dw) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * dw; // DF:211
			return;
		}
		va = va + (c.m * (pi.w + __ldg(&posT3[j]).w)) * dw; // DF:203
	});
#else
	SPH_FOR_FLUID_N(L, c, s, nf_, j) {
		if (SPH_IS_RIGID(j)) {
			float4 pj = __ldg(&rg.rspos[SPH_RIGID_SLOT(j)]);
			Pair p = make_pair(pi, pj);
			va = va + (((pj.w * SPH_RHO0) * k_i) / rho_i) * cubic_dw(p, c); // DF:211
			return;
		}
		float4 pj = __ldg(&posT3[j]);
		Pair p = make_pair(pi, pj);
		va = va + (c.m * (pi.w + pj.w)) * cubic_dw(p, c); // DF:203
	};
#endif
	f3 vb = F3(0.0f, 0.0f, 0.0f);
	if (c.boundary_handle == 1) {
		SPH_FOR_BOUNDARY_N(L, c, s, nb_, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			vb = vb + ((pj.w * k_i) / rho_i) * cubic_dw(p, c); // DF:219
		};
	}
	if (!live) return;
	f3 delta = c.boundary_handle == 1 ? va + vb * SPH_RHO0 : va; // DF:187
	f3 v = xyz(svadv[s]) - delta * dt; // DF:191
	svadv[s] = F4(v, 0.0f);
	mg_push(pu, s, v.x, v.y, v.z);
}

// DF:235-250 compute_all_position, fused with the write-back into the caller's original-order state
__global__ void __launch_bounds__(SPH_BLOCK)
k_df_position(SphConsts c, const int *__restrict__ sorted_id, const float4 *__restrict__ spos,
              const float4 *__restrict__ svel, const float4 *__restrict__ svadv, float4 *__restrict__ pos,
              float4 *__restrict__ vel, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	int i = sorted_id[s];
	if (i >= c.N_owned) return;
	float dt = ctl->dt;
	f3 va = xyz(svadv[s]);
	f3 x = xyz(spos[s]) + (dt * va) * 0.9999f; // DF:238
	f3 v = va * 0.9999f;                       // DF:239
	if (c.boundary_handle == 0) {              // DF:241-250
		float *xp = &x.x, *vp = &v.x;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			if (xp[k] <= c.clamp_lo[k]) { xp[k] = c.clamp_lo[k]; vp[k] *= -0.5f; }
			if (xp[k] >= c.clamp_hi[k]) { xp[k] = c.clamp_hi[k]; vp[k] *= -0.5f; }
		}
	}
	pos[i] = F4(x, 0.0f);
	vel[i] = F4(v, svel[s].w);
}

// ---- controllers: one block reduces the block partials deterministically and takes the reference's
// ---- host-side loop decision on the device (sph_ctl.cuh).  On several GPUs the partials travel with the
// ---- halo exchange and the receive kernel decides (sph_multigpu.cu). -------------------------------------
__device__ __forceinline__ void reduce_partials(const SphPartial *p, int n, double &sum, int &cnt, float &mx) {
	sph_reduce_partials<256>(p, n, sum, cnt, mx);
}
__global__ void __launch_bounds__(1024) k_df_ctl(int kind, SphCtl *ctl, const SphPartial *partials, int n, SphCtlArgs args) {
	if ((kind == SPH_CTL_DIV_ITER && !ctl->div_active) || (kind == SPH_CTL_DEN && !ctl->den_active)) return;
	double sum; int cnt; float mx;
	sph_reduce_partials<1024>(partials, n, sum, cnt, mx);
	if (threadIdx.x == 0) sph_ctl_apply(kind, ctl, sum, cnt, mx, args);
}
// DF:225 after iter_all_vel_adv of iteration den_iters: does the next iteration run?
__global__ void k_df_ctl_den_next(SphCtl *ctl) {
	if (!ctl->den_active) return;
	int it = ctl->den_iters + 1;
	ctl->den_iters = it;
	ctl->den_active = (it < 2 || (double)ctl->den_avg - 1000.0 > 0.1 * 1000 * 0.01) ? 1 : 0; // DF:225
	if (it >= 1000) { ctl->den_active = 0; atomicOr(&ctl->error_flags, SPH_ERR_DENSITY_CAP); }
}
// the loop decision that follows a sweep: one controller launch, or (several GPUs) part of the exchange
static void df_decide(SphHandle *h, int what, int kind, int nb, cudaStream_t st) {
	if (h->comm) { mg_exchange_reduce(h, what, kind, nb, st); return; }
	SphCtlArgs a;
	a.dt_cfl_c1 = h->c.dt_cfl_c1;
	a.rs = h->rstate;
	a.rigid_exists = (h->c.Nr > 0 && h->rigid_ready) ? 1 : 0;
	k_df_ctl<<<1, 1024, 0, st>>>(kind, h->ctl, h->partials, nb, a);
	h->launches++;
}

// ---- DFSPH drivers -----------------------------------------------------------------------------
// launch a kernel templated on RIGID with the instantiation the scene needs
#define SPH_LAUNCH_R(K, GRID, STREAM, ...)                                                \
	do {                                                                                  \
		if (rg.active) K<true><<<GRID, SPH_BLOCK, 0, STREAM>>>(__VA_ARGS__);              \
		else K<false><<<GRID, SPH_BLOCK, 0, STREAM>>>(__VA_ARGS__);                       \
	} while (0)

void rigid_lists(SphHandle *h, cudaStream_t st);
void rigid_force_df(SphHandle *h, int gated, cudaStream_t st);

// the loop decision after a sweep that wrote `n_partials` block partials; `what` still has to travel unless `exchanged`
static void df_decide_after(SphHandle *h, int what, bool exchanged, int kind, int n_partials, cudaStream_t st) {
	df_decide(h, exchanged ? MG_NONE : what, kind, n_partials, st);
}

// DF:314-355 divergence_warm_start
static void df_warm_start(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	(void)nb;
	sweep(h, MG_F4_VEL, KC_DF_WARM, st, true, nullptr, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_warm_start, grid, s, c, L, rg, h->a4[A4_T1], h->bspos, h->a1[A1_RHO], h->a4[A4_VEL], h->ctl, pu);
	});
}

// DF:252-280 derivative_iter_all_rho + the loop decision that follows it (DF:398-399 before the loop, DF:406-414
// inside it, where the sweep is gated on ctl->div_active)
static void df_drho(SphHandle *h, int in_loop, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	(void)nb;
	bool done = false;
	int np = sweep(h, MG_F4_T2, KC_DF_DRHO, st, false, &done, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_drho, grid, s, c, L, rg, h->a4[A4_POS], h->a4[A4_VEL], h->bspos, h->nbr_count, h->a1[A1_RHO],
		             h->a1[A1_ALPHA], h->a1[A1_DRHO], h->a4[A4_T2], h->ctl, h->partials, in_loop, pu);
	});
	df_decide_after(h, MG_F4_T2, done, in_loop ? SPH_CTL_DIV_ITER : SPH_CTL_DIV_FIRST, np, st);
}

// DF:302-312 divergence_iter_all_vel_adv + DF:381-384 sum_up_stiff, gated on ctl->div_active
static void df_div_vel(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	(void)nb;
	sweep(h, MG_F4_VEL, KC_DF_DIV, st, true, nullptr, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_div_iter, grid, s, c, L, rg, h->a4[A4_T2], h->bspos, h->a1[A1_RHO], h->a1[A1_ALPHA], h->a1[A1_DRHO],
		             h->a4[A4_VEL], h->ctl, pu);
	});
}

// DF:393-416 correct_divergence_error: the 15 passes (max_iteration_density_divergence, DF:24) are enqueued
// unconditionally and gate themselves on the device flag
static void df_divergence(SphHandle *h, cudaStream_t st) {
	df_warm_start(h, st);
	df_drho(h, 0, st);
	for (int it = 0; it < 15; ++it) {
		df_div_vel(h, st);
		df_drho(h, 1, st);
	}
}

static void df_ext_force_vel_adv(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	(void)nb;
	bool done = false;
	int np = sweep(h, MG_F4_VADV, KC_DF_EXT, st, false, &done, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_ext_force, grid, s, c, L, rg, h->a4[A4_PR], h->a4[A4_VEL], h->a1[A1_RHO], h->a4[A4_VADV], h->a4[A4_FA],
		             h->ctl, h->partials, pu);
	});
	// DF:105-110 loops over all rigid particles whenever a rigid body exists, active or not
	df_decide_after(h, MG_F4_VADV, done, SPH_CTL_DT, np, st);
}

// DF:124-152 compute_all_rho_adv of pass `it` + the average the loop condition reads (DF:225)
static void df_den_rho(SphHandle *h, int it, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	int gated = it >= 2 ? 1 : 0; // min_iteration_density (DF:21)
	(void)nb;
	bool done = false;
	int np = sweep(h, MG_F4_T3, KC_DF_RHOADV, st, false, &done, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_rho_adv, grid, s, c, L, rg, h->a4[A4_POS], h->a4[A4_VADV], h->bspos, h->a1[A1_RHO], h->a1[A1_ALPHA],
		             h->a1[A1_RHOADV], h->a4[A4_T3], h->ctl, h->partials, gated, pu);
	});
	df_decide_after(h, MG_F4_T3, done, SPH_CTL_DEN, np, st);
}

// DF:178-219 iter_all_vel_adv of pass `it` (+ the fluid -> rigid forces, DF:212) and the decision whether pass it+1 runs
static void df_den_vel(SphHandle *h, int it, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	int gated = it >= 2 ? 1 : 0;
	(void)nb;
	sweep(h, MG_F4_VADV, KC_DF_VELADV, st, true, nullptr, [&](int grid, cudaStream_t s, const SphLists &L, const SphMgPush &pu) {
		SPH_LAUNCH_R(k_df_vel_adv_iter, grid, s, c, L, rg, h->a4[A4_T3], h->bspos, h->a1[A1_RHO], h->a1[A1_ALPHA], h->a1[A1_RHOADV],
		             h->a4[A4_VADV], h->ctl, gated, pu);
	});
	if (rg.active) rigid_force_df(h, gated, st); // DF:212, gather form
	k_df_ctl_den_next<<<1, 1, 0, st>>>(h->ctl);
	h->launches += 1;
}

static void df_density_iters(SphHandle *h, int first, int count, cudaStream_t st) {
	for (int it = first; it < first + count; ++it) {
		df_den_rho(h, it, st);
		df_den_vel(h, it, st);
	}
}

static int df_density(SphHandle *h, cudaStream_t st) {
	// The reference loop has no iteration cap (DF:225).  Iterations are enqueued in chunks gated by the
	// device flag; the host looks at the flag once per chunk (not per iteration).
	int chunk = h->last_den_chunk > 0 ? h->last_den_chunk : 3;
	int done = 0;
	if (rigid_args(h).active) rigid_lists(h, st);
	for (;;) {
		df_density_iters(h, done, chunk, st);
		done += chunk;
		cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(SphCtl), cudaMemcpyDeviceToHost, st);
		cudaStreamSynchronize(st);
		if (!h->ctl_host->den_active) break;
		chunk = 2;
	}
	h->last_den_chunk = h->ctl_host->den_iters + 1;
	return 0;
}

static void df_position(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_DF_POS, st);
	k_df_position<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->fg.sorted_id, h->a4[A4_POS], h->a4[A4_VEL],
	                                                          h->a4[A4_VADV], h->pos, h->vel, h->ctl);
	sph_prof_end(h, st);
	h->launches++;
}

void df_phase(SphHandle *h, int phase, cudaStream_t st) {
	switch (phase) {
	case SPH_PH_DF_INITIALIZE: first_phase_lists(h, st); break;
	case SPH_PH_DF_DIVERGENCE: df_divergence(h, st); break;
	case SPH_PH_DF_EXT_FORCE_VEL_ADV: df_ext_force_vel_adv(h, st); h->den_piece = 0; break;
	case SPH_PH_DF_DENSITY: df_density(h, st); break;
	case SPH_PH_DF_POSITION: df_position(h, st); break;
	// the same step one sweep at a time (single-sweep parity tests drive these; the loop decisions stay on the device)
	case SPH_PH_DF_WARM_START: df_warm_start(h, st); break;
	case SPH_PH_DF_DRHO_FIRST: df_drho(h, 0, st); break;
	case SPH_PH_DF_DIV_VEL: df_div_vel(h, st); break;
	case SPH_PH_DF_DIV_DRHO: df_drho(h, 1, st); break;
	case SPH_PH_DF_DEN_RHO:
		if (h->den_piece == 0 && rigid_args(h).active) rigid_lists(h, st);
		df_den_rho(h, h->den_piece, st);
		break;
	case SPH_PH_DF_DEN_VEL: df_den_vel(h, h->den_piece, st); h->den_piece++; break;
	default: break;
	}
}

void df_step(SphHandle *h, cudaStream_t st) {
	first_phase_lists(h, st);
	df_divergence(h, st);
	df_ext_force_vel_adv(h, st);
	df_density(h, st);
	df_position(h, st);
}

// placeholders filled in by the other solver sections below
void wc_phase(SphHandle *h, int phase, cudaStream_t st);
void pc_phase(SphHandle *h, int phase, cudaStream_t st);
void pc_precompute(SphHandle *h, cudaStream_t st);
void pc_set_delta(SphHandle *h, int target, cudaStream_t st);
void ii_phase(SphHandle *h, int phase, cudaStream_t st);
void pbf_phase(SphHandle *h, int phase, cudaStream_t st);

} // namespace SPH_NS

#include "sph_rigid.cuh"
#include "sph_sweeps_other.cuh"
#include "sph_sweeps_pbf.cuh"

