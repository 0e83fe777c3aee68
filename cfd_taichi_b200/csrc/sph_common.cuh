// sph_common.cuh -- shared declarations of the B200 SPH hot path (internal; the public C-ABI is
// include/sph_b200.h).  Compiled for sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sph_b200.h"

#ifndef SPH_BLOCK
#define SPH_BLOCK 128            // threads per sweep block (one sorted particle per thread)
#endif
#define SPH_RHO0 1000.0f         // solver_base.rho_0 (SB:19)

// error flag bits latched on the device (SphStats.error_flags)
#define SPH_ERR_OUT_OF_GRID 1
#define SPH_ERR_LIST_OVERFLOW 2
#define SPH_ERR_BLIST_OVERFLOW 4
#define SPH_ERR_DENSITY_CAP 8
#define SPH_ERR_NONFINITE 16
#define SPH_ERR_COMM_TIMEOUT 32 // a peer did not raise its exchange flag within 10 s (multi-GPU windows)
#define SPH_ERR_BOUNDS 64       // an index check of the bounds-checked build (-DSPH_DEBUG_BOUNDS=1) failed

// compute-sanitizer is closed on the GPU pool this is developed on, so the library carries its own index checks:
// a build with -DSPH_DEBUG_BOUNDS=1 (cfd_taichi_b200/build.py: _build_debug/) validates every neighbour-list entry,
// list length and candidate segment before it is used and latches SPH_ERR_BOUNDS; tests/test_gpu_bounds.py runs every
// solver, the coupled rigid scene and restart on small scenes under it.
#ifndef SPH_DEBUG_BOUNDS
#define SPH_DEBUG_BOUNDS 0
#endif
#if SPH_DEBUG_BOUNDS
#define SPH_BOUNDS_OK(cond, errptr) do { if (!(cond)) atomicOr((errptr), SPH_ERR_BOUNDS); } while (0)
#else
#define SPH_BOUNDS_OK(cond, errptr) ((void)0)
#endif

// Constants the reference evaluates in Python scope (fp64) and then casts to f32 where they meet
// an f32 expression.  Filled once on the host in sph_api.cu.
struct SphConsts {
	float h;            // support_radius = 4 r (PS:82)
	float inv_h;        // fast mode only
	float cull_t;       // largest f32 t with sqrtf(t) <= h  (sqrt-free form of PS:466)
	float r, d;         // particle_radius, particle_diameter
	float m;            // particle_m (PS:83)
	float kW;           // 8 / (pi h^3)   (SB:79)
	float kDW;          // 48 / (pi h^3)  (SB:95)
	float kDW6, nkDW6;  // (k*6) and ((-k)*6) of SB:98,100
	float dwA, dwB;     // fast kernels: kDW6 / h^2 and nkDW6 / h
	float gq_inv, gq_scale; // fast kernels: fixed-point step of the quantised gradient stream (sph_sweeps.cu) and its inverse
	float gravity;
	float visc_num;     // 2 * alpha * h * c_s  (SB:187)
	float visc_eps_h2;  // eps * h * h          (SB:188)
	float neg_m;        // -particle_m          (SB:189)
	float tension_coef; // -k_t / m * m         (SB:216)
	float dt_cfl_c1;    // 0.4 * r * 2          (DF:112)
	float clamp_lo[3], clamp_hi[3]; // box_min + margin, box_max - margin (clamp boundary mode)
	float pc_beta;      // PC:23
	int gx, gy, gz;     // grid_num (PS:101)
	int gxz;            // gx * gz = y stride (PS:102)
	int G;              // number of cells
	int N;              // fluid particles handled by this handle (owned + ghost)
	int N_owned;        // owned fluid particles (== N on one GPU)
	int Nb, Nr;
	int kmax, kbmax;        // neighbour-list capacities (entries a particle may hold before the overflow flag)
	int kstride, kbstride;  // storage capacities of the quad-interleaved lists: kmax / kbmax rounded up to 4
	int boundary_handle, fs_couple, solver;
	int active_rigid;
};

// Device-resident solver control block: time step, loop state, reduction results.
// Written by single-thread controller kernels, read by every sweep (no host round trip).
struct SphCtl {
	float dt, dt2, ps_dt;
	int error_flags;
	int simulate_cnt;
	// DFSPH divergence loop (DF:393-416)
	int div_active, div_iters;
	float div_err, div_past, div_first;
	// DFSPH density loop (DF:221-233)
	int den_active, den_iters;
	float den_avg;
	// PCISPH (PC:47-70)
	int pc_active, pc_iters;
	float pc_err, pc_delta;
	int pc_max_index;
	// IISPH (II:78-100)
	int ii_active, ii_iters, ii_have_last;
	float ii_residual, ii_last;
	int max_nbr, max_bnbr;
	float max_vel;
	int graph_cond; // unused
	int blocks_done; // blocks-finished counter of the sweep tails (sph_ctl.cuh), zero between launches
};

// Device-resident rigid-body state (rigid_solver.py:12-31 + the uniform per-particle fields the
// reference fills with .fill(): vel RS:97, omega RS:96, alpha RS:128, acc RS:41).
struct SphRigidState {
	float centroid[3];      // ps.rigid_centriod (PS:271)
	float inertia[9];       // PS:290
	float inertia_inv[9];   // PS:291, rotated every step (RS:141)
	float vel[3], omega[3], alpha[3], acc[3];
	float attitude[3];      // RS:127
	float mass;             // RS:161
	float rs_dt;            // rigid_solver.delta_time (RS:13, 223-224)
	float max_surface_vel;  // DF:104-110
	float force_sum[3], torque[3];
	int collision_cnt;
	int simulate_cnt;
	int active;
};

#define SPH_RIGID_BIT 0x80000000u

// what the sweeps need to evaluate a rigid neighbour
struct SphRigidArgs {
	const float4 *rspos;        // sorted rigid particles: xyz, w = volume
	const int *rstart;          // rigid grid CSR
	const int *rsorted_id;      // sorted rigid slot -> rigid-local index
	const SphRigidState *st;
	const float4 *pos_orig;     // fluid positions in ORIGINAL order (neighbour-count quirk, PS:440-442)
	const int *slot_of;         // fluid original index -> sorted slot (viscosity quirk, SB:199)
	const int *gid;             // multi-GPU slabs only: global id of the fluid particle with local original index i; else null
	const float4 *quirk;        // multi-GPU slabs only: (pos, rho) of the fluid particles with GLOBAL id < Nr, replicated
	                            // on every rank (the two quirks index FLUID arrays with a rigid-local index); else null
	int active;
};

struct __align__(16) SphPartial {
	double sum;
	int cnt;
	float maxv;
};

// Per-step neighbour lists, quad-interleaved per warp: entries 4q .. 4q+3 of sorted particle s are the
// four words of the uint4 at  list4[((s >> 5) * (cap / 4) + q) * 32 + (s & 31)]  (cap is a multiple of 4),
// so one 128-bit load brings four entries and a warp reads 512 contiguous bytes per request.
struct SphLists {
	uint32_t *flist; int *fcount;   // fluid neighbours (indices into the sorted fluid arrays)
	uint32_t *blist; int *bcount;   // boundary neighbours (indices into the sorted boundary arrays)
	// DFSPH only (null otherwise): per-pair cache [index | grad W_ij] of the fluid list, one float4 per entry,
	// entry k of sorted particle s at gw[((s >> 5) * cap + k) * 32 + (s & 31)] (a warp reads 512 contiguous bytes)
	float4 *gw;
	// fast DFSPH kernels instead: grad W_ij as 3 x 21-bit fixed point in 8 bytes (the index comes from flist);
	// entries 2p, 2p+1 of sorted particle s in the uint4 at gq[((s >> 5) * (cap / 2) + p) * 32 + (s & 31)]
	uint4 *gq;
	// multi-GPU overlap (sph_multigpu.cu: mg_split): a sweep is launched twice -- split_mode 1 over the edge particles
	// (edge_list: the sorted slots of the particles a neighbour rank holds a ghost copy of), split_mode 2 over all
	// particles except those (edge_mask: one bit per sorted slot) -- so that the halo exchange of the edge values runs
	// behind the interior launch.  split_mode 0: one launch over everything (one GPU, and the default on slabs).
	const int *edge_list;
	const uint32_t *edge_mask;
	int n_edge, split_mode, partial_offset;
	int skip_ghost_fill; // the exchange that runs beside this launch fills the ghosts' records: the list build must not touch them
	int *err;       // &ctl->error_flags (bounds-checked build)
	int n_fluid, n_rigid, n_boundary, cap_f, cap_b; // limits the bounds-checked build validates list entries against
};

__host__ __device__ inline size_t sph_list_base(int s, int cap) {
	return ((size_t)(s >> 5) * (size_t)cap) * 32u + (size_t)(s & 31);
}

// word offset of entry n of sorted particle s in a quad-interleaved list
__host__ __device__ inline size_t sph_list_word(int s, int cap, int n) {
	return ((((size_t)(s >> 5) * (size_t)(cap >> 2) + (size_t)(n >> 2)) * 32u + (size_t)(s & 31)) << 2) + (size_t)(n & 3);
}

#define SPH_CUDA_CHECK(h, expr)                                                         \
	do {                                                                                \
		cudaError_t e_ = (expr);                                                        \
		if (e_ != cudaSuccess) return sph_fail_cuda((h), e_, #expr, __FILE__, __LINE__); \
	} while (0)
