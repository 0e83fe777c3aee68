// sph_internal.h -- handle layout and internal launcher prototypes (not part of the C-ABI).
#pragma once
#include "sph_common.cuh"

// float4 work arrays in SORTED (cell-contiguous) order; roles per solver are listed in DESIGN.md.
enum {
	A4_POS = 0, // xyz = position, w = 0
	A4_VEL,     // xyz = velocity, w = solver-persistent scalar (warm_start_k / p_past)
	A4_T1,      // xyz = position, w = neighbour payload #1 (e.g. (k/dt)/rho for the warm start)
	A4_T2,      // xyz = position, w = neighbour payload #2
	A4_T3,      // xyz = position, w = neighbour payload #3
	A4_PR,      // xyz = position, w = rho
	A4_VADV,    // xyz = advected / predicted velocity
	A4_FA,      // generic vec (force_ext, press_force, d_ii ...)
	A4_FB,      // generic vec
	A4_FC,      // generic vec
	A4_FD,      // generic vec
	A4_COUNT
};
enum {
	A1_RHO = 0, A1_ALPHA, A1_DRHO, A1_RHOADV, A1_P, A1_SA, A1_SB, A1_SC, A1_SD, A1_COUNT
};

struct SphGrid {
	int *cell_of;      // per particle (original order): 1-D cell id
	int *cell_cnt;     // per cell: counter / fill cursor
	int *cell_start;   // G + 1 exclusive prefix sums
	int *sorted_id;    // per sorted slot: original index
	int *scell;        // per sorted slot: 1-D cell id
	int *slot_of;      // per particle (original order): sorted slot
	int n;
};

// live per-kernel-class timing with CUDA events on the launching stream (bench.py roofline)
#define SPH_PROF_CAP 8192
#define SPH_PROF_CLASSES 32
enum {
	KC_GRID = 0, KC_LISTS, KC_DF_WARM, KC_DF_DRHO, KC_DF_DIV, KC_DF_EXT, KC_DF_RHOADV, KC_DF_VELADV, KC_DF_POS,
	KC_CTL, KC_WC_FORCE, KC_WC_KIN, KC_PC_EXT, KC_PC_PREDICT, KC_PC_RHO, KC_PC_FORCE, KC_PC_INT,
	KC_II_ADV, KC_II_AII, KC_II_DIJ, KC_II_UPDATE, KC_II_INT, KC_RIGID, KC_OTHER, KC_MG_EXCHANGE, KC_MG_STEP, KC_MG_WAIT
};
struct SphProf {
	bool on;
	int n;
	int kid[SPH_PROF_CAP];
	cudaEvent_t e0[SPH_PROF_CAP], e1[SPH_PROF_CAP];
	bool created;
};

struct SphHandle {
	SphConfig cfg;
	SphConsts c;
	int device;
	char err[512];
	int launches;
	int simulate_cnt; // SB:137
	// caller-owned state
	float4 *pos, *vel, *bpos, *rpos, *rvel, *rforce, *acc;
	size_t n_pos, n_vel, n_bpos, n_rpos, n_rvel, n_rforce, n_acc;
	// grids
	SphGrid fg, bg, rg;
	int *scan_sums;
	int scan_sums_cap;
	// sorted static boundary / rigid
	float4 *bspos;  // xyz, w = volume
	float4 *rspos;  // xyz, w = volume * rho0-free volume
	SphRigidState *rstate;      // device
	float4 *rverts; size_t n_rverts; // caller-owned mesh vertices (float4)
	uint32_t *rl_list; int *rl_count; int rl_cap; // rigid-centric fluid neighbour lists (force gather)
	bool rigid_ready;
	// sorted work arrays
	float4 *a4[A4_COUNT];
	float *a1[A1_COUNT];
	SphLists L;
	int *nbr_count; // get_neighbour_count (PS:424-445), sorted order
	SphCtl *ctl;        // device
	SphCtl *ctl_host;   // pinned mirror
	SphPartial *partials;
	int n_partials;
	double *red; // device: {sum, cnt, max} of the last reduction (all ranks)
	bool grid_valid, boundary_ready, lists_valid;
	cudaGraphExec_t step_graph;   // whole-step CUDA graph of the loop-free solvers (sph_api.cu: step_graph_launch)
	cudaStream_t graph_stream;
	int step_graph_launches;
	cudaStream_t copy_stream;  // deferred velocity upload (sph_upload_state_xyz)
	cudaEvent_t ev_vel_ready, ev_mark;
	bool vel_in_flight;        // a velocity upload is on the copy stream: h->vel is not complete yet
	bool vel_gather_pending;   // the grid of this step was built without the sorted velocities
	unsigned long long *render_zbuf; // sph_render: (depth | colour) per pixel
	size_t render_cap;
	int async_error;      // a library call inside a void helper failed (NCCL transport): the enclosing sph_* call returns it
	bool lists_fresh;     // SPH_PH_BUILD_LISTS has built this step's lists: the solver's first phase skips its own build
	int sweep_blocks;
	int last_den_chunk;
	int den_piece;        // density iterations issued one at a time through SPH_PH_DF_DEN_ONE since the last v* pass
	float *xyz_stage;     // 2 x 3 floats per fluid particle: staging of sph_upload_state_xyz / sph_download_state_xyz
	SphProf *prof;
	struct SphComm *comm; // multi-GPU slab state (sph_multigpu.cu); null on one GPU
	int *gid;             // caller-owned global particle ids (multi-GPU)
	size_t n_gid;
};

int sph_fail(SphHandle *h, int code, const char *fmt, ...);
static inline void sph_prof_begin(SphHandle *h, int kid, cudaStream_t st) {
	SphProf *p = h->prof;
	if (!p || !p->on || p->n >= SPH_PROF_CAP) return;
	p->kid[p->n] = kid;
	cudaEventRecord(p->e0[p->n], st);
}
static inline void sph_prof_end(SphHandle *h, cudaStream_t st) {
	SphProf *p = h->prof;
	if (!p || !p->on || p->n >= SPH_PROF_CAP) return;
	cudaEventRecord(p->e1[p->n], st);
	p->n++;
}
int sph_fail_cuda(SphHandle *h, cudaError_t e, const char *expr, const char *file, int line);

// ---- sph_grid.cu (mode independent) -------------------------------------------------------
void sphg_build(SphHandle *h, SphGrid &g, const float4 *pos, int n, cudaStream_t st, const int *gid = nullptr);
void sphg_gather_fluid(SphHandle *h, cudaStream_t st);
void sphg_gather_fluid_pos(SphHandle *h, cudaStream_t st);   // deferred velocity upload: positions + vel.w now ...
void sphg_gather_vel(SphHandle *h, cudaStream_t st);         // ... the velocities when they have arrived
// sph_api.cu: velocities uploaded by sph_upload_state_xyz travel on a copy stream behind the grid and list build;
// whoever reads the velocity arrays first calls this (waits for the copy, gathers the sorted velocities)
void sph_finish_deferred_vel(SphHandle *h, cudaStream_t st);
void sphg_gather_boundary(SphHandle *h, cudaStream_t st);
void sphg_gather_rigid(SphHandle *h, cudaStream_t st);
void sphg_unsort_f1(SphHandle *h, const SphGrid &g, const float *in, float *out, int n, cudaStream_t st);
void sphg_unsort_i1(SphHandle *h, const SphGrid &g, const int *in, int *out, int n, cudaStream_t st);
void sphg_unsort_f4(SphHandle *h, const SphGrid &g, const float4 *in, float4 *out, int n, cudaStream_t st);
void sphg_writeback(SphHandle *h, const float4 *spos, const float4 *svel, cudaStream_t st);
void sphg_unpack_xyz(SphHandle *h, const float *p3, const float *v3, int n, cudaStream_t st);
void sphg_pack_xyz(SphHandle *h, float *p3, float *v3, int n, cudaStream_t st);
void sphg_visualize(SphHandle *h, int what, float *rgb, int stride, cudaStream_t st);

// ---- sph_multigpu.cu: slab decomposition along x, NCCL halo exchange / migration / allreduce -------
// Every call is a no-op (returns immediately) when h->comm is null.
enum { MG_F4_T1R = 0 /* posT1.w + posR.w */, MG_F4_VEL, MG_F4_T2, MG_F4_VADV, MG_F4_T3,
       MG_F4_T1W /* posT1.w */, MG_NONE = -1 };
#define MG_XYZ(a4_index) (100 + (a4_index)) // xyz of the float4 work array a4[a4_index] (w is not carried)
struct SphMgPush;
SphMgPush mg_push_args(SphHandle *h);
// overlap of the halo exchange with interior compute (sph_sweeps.cu: sweep()): edge particles first on the exchange
// stream `xs`, the interior concurrently on the main stream; on == false: one launch over everything
struct MgSplit {
	bool on;
	cudaStream_t xs;
	const int *edge_list;
	const uint32_t *edge_mask;
	int n_edge;
};
MgSplit mg_split(SphHandle *h);
void mg_fork(SphHandle *h, cudaStream_t main_stream);   // the exchange stream waits for everything enqueued on the main stream
void mg_join(SphHandle *h, cudaStream_t main_stream);   // the main stream waits for the exchange stream                            // the next sweep pushes its edge values itself (sph_mgwin.cuh)
void mg_exchange(SphHandle *h, int what, cudaStream_t st);       // ghost values of one field, both neighbours
// ghost values + the all-reduce of the sweep's n_blocks block partials + the loop decision `ctl_kind`
// (sph_ctl.cuh) applied on every rank
void mg_exchange_reduce(SphHandle *h, int what, int ctl_kind, int n_blocks, cudaStream_t st);
int mg_begin_step(SphHandle *h, cudaStream_t st);                // migration + ghost exchange + counts
void mg_after_grid(SphHandle *h, cudaStream_t st);               // sorted slots of the send / recv lists
void mg_destroy(SphHandle *h);
// rigid bodies over slabs (the body is replicated on every rank)
void mg_allreduce_sum_f32(SphHandle *h, float *dev, size_t n, cudaStream_t st);   // no-op on one GPU
const float4 *mg_rigid_quirk(const SphHandle *h);                                       // null on one GPU
void mg_rigid_quirk_update(SphHandle *h, int with_rho, cudaStream_t st);          // no-op on one GPU

// ---- sph_sweeps.cu, compiled twice (namespace sph_strict with -fmad=false, sph_fast) -------
#define SPH_SWEEP_API(NS)                                                                   \
	namespace NS {                                                                          \
	void boundary_volume(SphHandle *h, cudaStream_t st);                                    \
	void build_lists(SphHandle *h, cudaStream_t st, bool push_lists = false);               \
	void df_step(SphHandle *h, cudaStream_t st);                                            \
	void df_phase(SphHandle *h, int phase, cudaStream_t st);                                \
	void first_phase_lists(SphHandle *h, cudaStream_t st);                                  \
	void wc_phase(SphHandle *h, int phase, cudaStream_t st);                                \
	void pc_phase(SphHandle *h, int phase, cudaStream_t st);                                \
	void pc_precompute(SphHandle *h, cudaStream_t st);                                      \
	void pc_set_delta(SphHandle *h, int target, cudaStream_t st);                           \
	void ii_phase(SphHandle *h, int phase, cudaStream_t st);                                \
	void pbf_phase(SphHandle *h, int phase, cudaStream_t st);                               \
	void rigid_init(SphHandle *h, cudaStream_t st);                                         \
	void rigid_step(SphHandle *h, cudaStream_t st);                                         \
	}
SPH_SWEEP_API(sph_strict)
SPH_SWEEP_API(sph_fast)
