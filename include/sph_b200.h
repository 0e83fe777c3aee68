/*
 * sph_b200.h -- C-ABI of the B200-native SPH hot path (drop-in for CFD_Taichi's
 * ParticleSystem / *_solver kernels).  Plain pointers and sizes only; no torch types.
 *
 * The reference has no FFI: its "operator API" is the set of @ti.kernel methods that main.py
 * reaches through ParticleSystem / solver_base / <name>_solver.step().  Each entry point below
 * cites the reference method(s) it replaces (file:line in the upstream repository).
 *
 * Conventions
 *   - every function returns 0 on success, a negative SPH_E* code on failure; the message is
 *     available from sph_last_error().  No exceptions cross the boundary.
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*).  Host synchronisation:
 *     sph_create / sph_destroy, sph_read_stats, sph_rigid_state / sph_rigid_set_state, sph_profile_end and the
 *     sph_download_* calls synchronise by contract.  sph_step / sph_phase of the three iterative solvers
 *     synchronise ONCE per step (never per solver iteration): the uncapped DFSPH density loop (DF:225) and the
 *     PCISPH / IISPH pressure loops are enqueued in chunks gated by a device flag, and the host reads that flag
 *     once per chunk (one chunk per step in steady state).  WCSPH, PBF and every single-sweep phase are fully
 *     asynchronous.  On several GPUs each step additionally reads the migration / ghost counts back once.
 *   - one host thread per handle (as in the reference: one Python thread drives main.py:95-206).
 *   - the caller (Python/torch) owns the particle state buffers bound with sph_bind; the
 *     library owns scratch only (cell arrays, neighbour lists, sorted work buffers).
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_OK 0
#define SPH_EINVAL (-1)
#define SPH_ECUDA (-2)
#define SPH_ENOMEM (-3)
#define SPH_ESTATE (-4)
#define SPH_ENOTBOUND (-5)

/* solver ids (main.py:65-68 resolves "<name>_solver") */
#define SPH_SOLVER_WCSPH 0
#define SPH_SOLVER_PCISPH 1
#define SPH_SOLVER_IISPH 2
#define SPH_SOLVER_DFSPH 3
#define SPH_SOLVER_PBF 4   /* pbf_solver.py with index-based task semantics (stale in the reference, SURVEY B-14) */

/* Mirrors the JSON blocks read by ParticleSystem.__init__ (PS:31-127), solver_base.__init__
 * (SB:7-39) and rigid_solver.__init__ (RS:6-31).  Host-side derived numbers (particle counts,
 * grid_num) are computed by the Python mirror with the reference's own fp64 formulas and
 * passed in, exactly as ParticleSystem does before it allocates its fields. */
typedef struct SphConfig {
	double box_min[3];
	double box_max[3];
	double particle_radius;   /* scene.particle_radius */
	double gravity;           /* scene.gravity */
	double delta_time;        /* solver.delta_time */
	int32_t boundary_handle;  /* solver.boundary_handle (1 = Akinci particles, 0 = box clamp) */
	int32_t fs_couple;        /* solver.fs_couple */
	int32_t solver;           /* SPH_SOLVER_* */
	int32_t n_fluid;          /* PS:85-86 particle_num */
	int32_t n_boundary;       /* PS:129-137 */
	int32_t n_rigid;          /* voxel points of solid.mesh (0 = no rigid body) */
	int32_t active_rigid;     /* solid.active */
	int32_t grid_num[3];      /* PS:100-101 */
	int32_t max_neighbors;          /* capacity of the per-step fluid neighbour list (0 = default) */
	int32_t max_boundary_neighbors; /* capacity of the boundary neighbour list (0 = default) */
	int32_t strict;           /* 1 = strict-fp32 kernels (no FMA, IEEE div/sqrt, reference operation
	                             order: bit-exact against the oracle); 0 = fast kernels (same
	                             neighbour sets, results within 1e-5) */
	int32_t n_ghost_capacity; /* multi-GPU: room for ghost fluid particles after the owned ones */
	double rigid_rho;         /* solid.rho_0 */
	int32_t use_graph;        /* reserved, must be 0.  The solver loops are stream-ordered launches gated by device
	                             flags; the loop-free solvers (WCSPH, PBF) replay a whole step as one CUDA graph on
	                             their own (one GPU, captured at the first sph_step; SPH_NO_GRAPH=1 disables it) */
	int32_t reserved;
} SphConfig;

/* Fields that can be bound (caller-owned device memory) or fetched (library scratch copied
 * into caller memory in ORIGINAL particle order). */
enum SphField {
	/* bindable state, float4 per particle */
	SPH_F_FLUID_POS = 0,   /* xyz = fluid_particles.pos (PS:6-20), w unused */
	SPH_F_FLUID_VEL = 1,   /* xyz = fluid_particles.vel, w = solver-persistent scalar
	                          (DFSPH warm_start_k DF:17, IISPH p_past II:21) */
	SPH_F_BOUNDARY_POS = 2,/* xyz = boundary_particles.pos, w = boundary_particles.volume */
	SPH_F_RIGID_POS = 3,   /* xyz = rigid_particles.pos, w = rigid_particles.volume */
	SPH_F_RIGID_VEL = 4,   /* xyz = rigid_particles.vel, w = rigid_particles.mass */
	SPH_F_RIGID_FORCE = 5, /* xyz = rigid_particles.force */
	SPH_F_FLUID_ACC = 6,   /* xyz = fluid_particles.acc (WCSPH only) */
	SPH_F_FLUID_GID = 8,   /* int32 global particle id per fluid particle (multi-GPU slabs) */
	SPH_F_RIGID_VERTICES = 7, /* xyz = ps.rigid_vertices (mesh vertices moved with the body, RS:101-102, 138-139) */
	/* fetchable per-fluid-particle results (float unless noted), original order */
	SPH_F_RHO = 16, SPH_F_ALPHA, SPH_F_RHO_DERIVATIVE, SPH_F_RHO_ADV, SPH_F_VEL_ADV /*float4*/,
	SPH_F_CELL1D /*int32*/, SPH_F_NEIGHBOR_COUNT /*int32*/, SPH_F_BOUNDARY_NEIGHBOR_COUNT /*int32*/,
	SPH_F_PRESSURE, SPH_F_FORCE_A /*float4 generic force/acc buffer*/, SPH_F_FORCE_B /*float4*/,
	SPH_F_SCALAR_A, SPH_F_SCALAR_B, SPH_F_SCALAR_C, SPH_F_VEC_A /*float4*/, SPH_F_VEC_B /*float4*/,
	SPH_F_VEC_C /*float4*/,
	/* the payload copies the sweeps gather (xyz = sorted position, w = the scalar the next sweep reads of a neighbour) */
	SPH_F_PAYLOAD_1 /*float4*/, SPH_F_PAYLOAD_3 /*float4*/, SPH_F_POS_RHO /*float4: xyz, rho*/,
	/* grid arrays (int32), fetched verbatim */
	SPH_F_CELL_START = 64,   /* G+1 exclusive prefix sums == per-cell list offsets */
	SPH_F_SORTED_INDEX,      /* N: original index of the particle in sorted slot s */
	SPH_F_BOUNDARY_CELL_START,
	SPH_F_BOUNDARY_SORTED_INDEX
};

/* Phases, for callers that drive a solver piecewise like the reference's public methods. */
enum SphPhase {
	SPH_PH_BUILD_GRID = 0,        /* PS:368-407 reset_grid + update_grid (+ reorder + lists) */
	/* DFSPH */
	SPH_PH_DF_INITIALIZE = 10,    /* DF:423-426 */
	SPH_PH_DF_DIVERGENCE,         /* DF:393-416 correct_divergence_error */
	SPH_PH_DF_EXT_FORCE_VEL_ADV,  /* DF:91-122 compute_all_ext_force + compute_all_vel_adv */
	SPH_PH_DF_DENSITY,            /* DF:221-233 correct_density_error */
	SPH_PH_DF_POSITION,           /* DF:235-250 compute_all_position */
	/* WCSPH */
	SPH_PH_WC_PRESSURE = 20,      /* WC:32-38 */
	SPH_PH_WC_KINEMATIC,          /* WC:40-63 */
	/* PCISPH */
	SPH_PH_PC_EXT_FORCE = 30,     /* PC:220-226 */
	SPH_PH_PC_ITERATION,          /* PC:47-70 */
	SPH_PH_PC_INTEGRATION,        /* PC:200-218 */
	/* IISPH */
	SPH_PH_II_PREDICT_ADVECTION = 40, /* II:35-75 */
	SPH_PH_II_PRESSURE_SOLVE,         /* II:78-100 */
	SPH_PH_II_INTEGRATION,            /* II:184-206 */
	/* PBF (fetch: rho = SPH_F_RHO, constrain = SCALAR_A, pbf_lambda = SCALAR_B, constrain_derivative =
	 * FORCE_A, delta_pos = FORCE_B, pos_predict = VEC_C) */
	SPH_PH_PBF_PREDICT = 50,          /* PBF:26-30 externel_force_predict_pos */
	SPH_PH_PBF_LAMBDA,                /* PBF:32-52 compute_all_lambda */
	SPH_PH_PBF_DELTA_POS,             /* PBF:55-65 compute_all_delta_pos */
	SPH_PH_PBF_UPDATE_POS,            /* PBF:67-96 update_all_pos (move all, then XSPH) */
	/* commit the sorted work buffers back into the bound original-order state */
	SPH_PH_WRITEBACK = 90,

	/* The same steps ONE SWEEP AT A TIME (cfd_taichi_b200/selfcheck.py, tests/test_gpu_fast_parity.py: every sweep
	 * of the fast kernels is fed the strict kernels' inputs through sph_copy_work_state).  The loop decisions stay
	 * on the device; a sweep of a loop that has ended is a no-op; SphStats reports the loop flags.
	 *   DF_DIVERGENCE    == WARM_START, DRHO_FIRST, 15 x (DIV_VEL, DIV_DRHO)
	 *   DF_DENSITY       == (DEN_RHO, DEN_VEL) while den_active
	 *   WC_PRESSURE      == BUILD_LISTS, WC_EOS, WC_FORCE
	 *   PC_EXT_FORCE     == BUILD_LISTS, PC_EXT_FORCE;   PC_ITERATION == PC_PREDICT, PC_RHO_FIRST, (PC_PRESS_FORCE, PC_RHO) while loop_active
	 *   II_PREDICT_ADVECTION == BUILD_LISTS, II_ADVECT, II_AII;   II_PRESSURE_SOLVE == II_SOLVE_BEGIN, (II_DIJ, II_UPDATE) while loop_active */
	SPH_PH_BUILD_LISTS = 100,     /* PS:447-469 the step's neighbour lists (+ rho SB:41-72, + alpha DF:32-89); the solver's first phase then skips its own build */
	SPH_PH_DF_WARM_START = 110,   /* DF:314-355 */
	SPH_PH_DF_DRHO_FIRST,         /* DF:252-280 + the decision of DF:398-399 */
	SPH_PH_DF_DIV_VEL,            /* DF:302-312 divergence_iter_all_vel_adv + DF:381-384 sum_up_stiff */
	SPH_PH_DF_DIV_DRHO,           /* DF:252-280 + the decision of DF:406-414 */
	SPH_PH_DF_DEN_RHO,            /* DF:124-152 compute_all_rho_adv + the average DF:225 reads */
	SPH_PH_DF_DEN_VEL,            /* DF:178-219 iter_all_vel_adv (+ DF:212 rigid forces) + the decision of DF:225 */
	SPH_PH_WC_EOS = 120,          /* WC:32-38 Tait pressure */
	SPH_PH_WC_FORCE,              /* WC:65-144 pressure gradient, boundary term, viscosity, tension (+ WC:126) */
	SPH_PH_PC_PREDICT = 130,      /* PC:72-87 predict_vel_pos */
	SPH_PH_PC_RHO_FIRST,          /* PC:89-100 predict_rho + PC:123-133 compute_residual + PC:54 */
	SPH_PH_PC_PRESS_FORCE,        /* PC:102-121 iter_press + update_press_force (+ predict_vel_pos, + PC:186) */
	SPH_PH_PC_RHO,                /* PC:89-100 + PC:123-133 + the decision of PC:56 */
	SPH_PH_II_ADVECT = 140,       /* II:35-54 f_adv, v_adv, d_ii */
	SPH_PH_II_AII,                /* II:56-75 rho_adv, a_ii, start iterate */
	SPH_PH_II_SOLVE_BEGIN,        /* II:78-82 */
	SPH_PH_II_DIJ,                /* II:208-250 sum_j d_ij p_j */
	SPH_PH_II_UPDATE              /* II:252-340 update_p + II:102-113 compute_residual + the decision of II:83-93 */
};

/* Read by sph_read_stats (the only synchronising query; replaces the kernel return values
 * DF:125,253 PC:122 II:103 and the prints DF:233,416 PC:70 II:96). */
typedef struct SphStats {
	float delta_time;          /* solver.delta_time[None] after the step (DFSPH: adaptive, DF:112-119) */
	float ps_delta_time;       /* ps.delta_time[None] (DF:119) */
	int32_t simulate_cnt;      /* SB:137 */
	int32_t error_flags;       /* bit0 particle outside grid (PS:393-395), bit1 neighbour list overflow,
	                              bit2 boundary list overflow, bit3 density loop hit the safety cap,
	                              bit4 non-finite value, bit5 a peer rank did not answer a slab exchange
	                              within 10 s (multi-GPU), bit6 an index check failed (bounds-checked build only) */
	int32_t div_iters;         /* DF:416 */
	float div_first_err, div_err;
	int32_t den_iters;         /* DF:233 */
	float den_err;
	int32_t pc_iters;          /* PC:70 */
	float pc_err;
	int32_t ii_iters;          /* II:96 */
	float ii_residual;
	int32_t max_neighbors_seen, max_boundary_neighbors_seen;
	float pc_delta;            /* PC:45 */
	int32_t pc_max_index;      /* PS:409-422 */
	int32_t kernel_launches;   /* launches issued by the library since creation */
	int32_t div_active, den_active; /* DFSPH loop flags on the device (1 while the loop of DF:400 / DF:225 would continue) */
	int32_t loop_active;            /* PCISPH / IISPH loop flag on the device (PC:56 / II:83) */
} SphStats;

typedef struct SphHandle SphHandle;

int sph_create(const SphConfig *cfg, int device, SphHandle **out);
int sph_destroy(SphHandle *h);
const char *sph_last_error(const SphHandle *h); /* h may be NULL: message of the failed sph_create */
int sph_abi_version(void);

/* Borrow caller-owned device memory for a state field (float4 * n). */
int sph_bind(SphHandle *h, int field, void *dev_ptr, size_t n);

/* One-time static boundary set-up: boundary grid (PS:322-335) and Akinci volumes (PS:309-320).
 * Writes the volume into SPH_F_BOUNDARY_POS.w. */
int sph_init_boundary(SphHandle *h, void *stream);
/* One-time rigid set-up (PS:249-295): Akinci volumes and masses of the rigid particles (written to
 * SPH_F_RIGID_POS.w / SPH_F_RIGID_VEL.w), centroid, inertia tensor and its inverse (device state). */
int sph_init_rigid(SphHandle *h, void *stream);

/* PCISPH pre_compute (PC:28-37): grid + neighbour lists/counts of the initial state.  The arg-max of
 * get_max_neighbor_particle_index (PS:409-422) is O(N) host logic on the fetched counts; its result is
 * handed to sph_pcisph_delta, which evaluates pre_compute_delta (PC:39-45) for that particle. */
int sph_pcisph_precompute(SphHandle *h, void *stream);
int sph_pcisph_delta(SphHandle *h, int particle_index, void *stream);
/* multi-GPU slabs: delta is computed by the rank that owns the arg-max particle (sph_pcisph_delta there) and
 * handed to the other ranks as a value; particle_index is the GLOBAL id reported in SphStats.pc_max_index */
int sph_pcisph_set_delta(SphHandle *h, float delta, int particle_index, void *stream);

/* One full solver.step() (SB:136-143 + <name>_solver.step), n_substeps times. */
int sph_step(SphHandle *h, int n_substeps, void *stream);
/* One phase (see enum SphPhase). */
int sph_phase(SphHandle *h, int phase, void *stream);

/* One rigid_solver.step() (RS:216-234), entirely on the device: torque and force reductions over the
 * rigid particles (RS:121-123, 35-38; hierarchical warp/block reductions in the fast kernels, serial
 * reference order in the strict kernels), attitude / rotation (RS:118-141), wall contact scan and
 * impulse (RS:53-94), particle, vertex and centroid update (RS:96-104).  Fluid->rigid forces are
 * GATHERED per rigid particle inside the fluid solver step (no per-pair atomics; replaces DF:212). */
int sph_rigid_step(SphHandle *h, void *stream);
typedef struct SphRigidInfo {
	float centroid[3], inertia[9], inertia_inv[9];
	float vel[3], omega[3], alpha[3], acc[3], attitude[3], force_sum[3], torque[3];
	float mass, delta_time;
	int32_t collision_cnt, simulate_cnt;
	float max_surface_vel;   /* DF:104-110, feeds the adaptive time step of the next DFSPH step */
	int32_t reserved;
} SphRigidInfo;
int sph_rigid_state(SphHandle *h, SphRigidInfo *out); /* synchronises */
/* Restart: put a state read with sph_rigid_state back (the rigid particle arrays are caller-owned and are
 * restored by the caller).  Synchronises. */
int sph_rigid_set_state(SphHandle *h, const SphRigidInfo *in);

int sph_set_delta_time(SphHandle *h, float dt, void *stream);

/* Copy a result field into caller device memory, original particle order. */
int sph_fetch(SphHandle *h, int field, void *dev_out, size_t n, void *stream);

/* Device-side scene initialisation (init_particle_pos, PS:139-195): the fluid lattice and the one-layer
 * boundary shell written straight into caller-owned float4 arrays, with the reference's f32 arithmetic.
 * Stateless: no handle is needed (the reference fills the positions before anything else, PS:119).
 * dev_ids (optional, int32, device): the global lattice indices to generate (multi-GPU slabs); NULL = 0..n-1.
 * particle_num_total selects the reference's f32 index arithmetic (< 2^24 particles, where it is exact)
 * or the integer lattice. */
typedef struct SphLattice {
	double particle_radius;
	double start_pos[3];
	double water_size[3];
	double box_min[3];
	double box_max[3];
} SphLattice;
int sph_init_fluid_lattice(const SphLattice *lat, long long particle_num_total, const int32_t *dev_ids, size_t n,
                           void *dev_pos4, int device, void *stream);
int sph_init_boundary_shell(const SphLattice *lat, size_t nb, void *dev_bpos4, int device, void *stream);

/* SB:219-245 visualize_rho / visualize_neighbour: rgb[i] = (0, 0.28, (q_i - min q) / (max q - min q)) in
 * ORIGINAL particle order, q = rho or get_neighbour_count; dev_rgb holds n rows of stride_floats floats
 * (3 = ps.rgb, 4 = an rgba buffer).  Unchanged when max == min (SB:230, 244).  Uses the density / counts of
 * the last solver step. */
#define SPH_VIS_RHO 0
#define SPH_VIS_NEIGHBOUR 1
int sph_visualize(SphHandle *h, int what, void *dev_rgb, int stride_floats, size_t n, void *stream);

/* Host-buffer entry points (the e2e path): pinned or pageable host memory, float4 * n_fluid. */
/* Headless replacement of the GGUI draw calls of main.py:153-161 (scene.ambient_light, scene.point_light,
 * scene.particles for the fluid and the rigid body with per-vertex colours) with the scene file's camera
 * (cam_pos / cam_look_at / cam_up, main.py:60-62): shaded sphere splats, depth-tested, into a caller-owned device
 * image of width * height RGBA8 pixels (row 0 = top).  dev_*_rgb: per-particle colours, `stride` floats per particle
 * (ps.fluid_particles.rgb / ps.rigid_particles.rgb); dev_depth (optional): float depth along the view direction,
 * +inf where nothing was drawn.  Asynchronous on `stream`. */
typedef struct SphCamera {
	double pos[3], look_at[3], up[3];
	double fov_y_deg;       /* <= 0: GGUI's default, 45 */
	double light_pos[3];    /* main.py:154: (0.5, 1.5, 1.5) */
	double ambient;         /* main.py:153: 0.8 */
	uint8_t background[4];  /* r, g, b, unused */
} SphCamera;
#define SPH_RENDER_FLUID 1  /* key 'f' / 'g' of main.py:134-137 */
#define SPH_RENDER_RIGID 2  /* key 'r' / 't' of main.py:138-141 */
int sph_render(SphHandle *h, const SphCamera *camera, int width, int height, int what, const void *dev_fluid_rgb,
               int fluid_rgb_stride, const void *dev_rigid_rgb, int rigid_rgb_stride, void *dev_rgba8, void *dev_depth,
               void *stream);

int sph_upload_state(SphHandle *h, const float *host_pos4, const float *host_vel4, void *stream);
int sph_download_state(SphHandle *h, float *host_pos4, float *host_vel4, void *stream);
/* The same with the host arrays the reference's callers hold (pos / vel as N x 3 floats, main.py:190):
 * 12 instead of 16 bytes per vector over PCIe; vel.w (DFSPH warm_start_k, IISPH p_past) stays on the device.
 * Either pointer may be NULL. */
int sph_upload_state_xyz(SphHandle *h, const float *host_pos3, const float *host_vel3, void *stream);
int sph_download_state_xyz(SphHandle *h, float *host_pos3, float *host_vel3, void *stream);

int sph_read_stats(SphHandle *h, SphStats *out);

/* Test support (single-sweep parity of the fast kernels against the strict ones, tests/test_gpu_fast_parity.py):
 * copy the in-step work state -- every sorted per-particle array, the control block, the rigid-body state -- of
 * `src` into `dst`.  Both handles must describe the same scene on the same device with their grids built from the
 * same positions.  Neighbour lists are not copied (each arithmetic mode keeps its own order). */
int sph_copy_work_state(SphHandle *dst, SphHandle *src, void *stream);

/* Live per-kernel-class timing: CUDA events recorded on the launching stream around every launch
 * between sph_profile_begin and sph_profile_end (which synchronises).  Class ids are listed in
 * DESIGN.md (0 grid build, 1 neighbour lists, 3 DFSPH drho sweep ...). */
int sph_profile_begin(SphHandle *h);
int sph_profile_end(SphHandle *h, float *ms_by_class, int32_t *launches_by_class, int n_classes);

/* Multi-GPU: one handle per GPU holds one x-slab [col_lo, col_hi) of the global grid.
 * The caller creates the handle with n_fluid = owned-particle capacity and n_ghost_capacity > 0, binds
 * SPH_F_FLUID_GID (int32 global particle ids, the sort tie-break that keeps the single-domain order),
 * sets the initial owned count with sph_set_counts and joins the NCCL communicator:
 *   rank 0: sph_comm_unique_id(id) -> broadcast the 128 bytes (torch.distributed plumbing) -> every
 *   rank: sph_comm_init(h, id, rank, nranks, col_lo, col_hi).
 * sph_step then performs, per step, on the caller's stream: particle migration and the one-column ghost-particle
 * exchange (one kernel each + one count read-back), and -- per sweep -- the ghost-value exchange and the
 * loop-decision reduction as ONE kernel; all of them store into the neighbours' CUDA-IPC peer windows over NVLink
 * and poll their own (no NCCL, no host; SPH_MG_TRANSPORT=nccl selects NCCL instead, for A/B runs). */
int sph_comm_unique_id(char *out128);
int sph_comm_init(SphHandle *h, const char *id128, int rank, int nranks, int col_lo, int col_hi);
int sph_comm_info(SphHandle *h, int32_t *out8); /* owned, ghosts, sent L/R, received L/R, rank, nranks */
int sph_set_counts(SphHandle *h, int n_owned, int n_ghost);

#ifdef __cplusplus
}
#endif
#endif
