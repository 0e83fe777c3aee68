/*
 * sph_oracle.h -- CPU restatement of the CFD_Taichi SPH hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a strict IEEE-754 binary32 / int32 restatement of the
 * reference's Taichi kernels (ParticleSystem.py, solver_base.py, {wcsph,pcisph,iisph,
 * dfsph}_solver.py, rigid_solver.py), operations in source order, no FMA contraction,
 * canonical neighbour order = ti.cpu with cpu_max_num_threads=1 (27 cells in (dx,dy,dz)
 * lexicographic order with dz fastest, ascending global particle index inside a cell).
 *
 * PARITY UNPINNED for the compiled reference: it has no tests or golden vectors and taichi==1.6.0
 * cannot be imported in this image.  What IS pinned: the reference's own solver sources, imported
 * unmodified from /root/reference and EXECUTED under a stand-in for the Taichi front end
 * (tests/golden/ti_shim, Taichi's scalar rules as read in SURVEY.md appendix A), give the
 * fixtures tests/golden/refshim_*.npz, which this oracle reproduces bit for bit -- every solver
 * array after every step, iteration counts and printed residuals (tests/test_reference_shim.py).
 * That pins the transcription (statements, operand order, loop structure, constants, quirks);
 * Taichi's own code generation stays this repository's reading, hence the first two words.
 * Hand-derived known-answer checks: tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (cfd_taichi_b200/) never does.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcConfig {
	double box_min[3];
	double box_max[3];
	double particle_radius;
	double gravity;
	double start_pos[3];
	double water_size[3];
	double delta_time;
	int boundary_handle;   /* JSON solver.boundary_handle (default true) */
	int fs_couple;         /* JSON solver.fs_couple (default true) */
	int solver;            /* 0 wcsph, 1 pcisph, 2 iisph, 3 dfsph, 4 pbf(index semantics) */
	int exist_rigid;       /* config has a 'solid' block */
	int active_rigid;      /* solid.active */
	double rigid_rho;      /* solid.rho_0 */
	double rigid_pos_offset[3];
	double rigid_att_offset_deg[3];
	int n_rigid_vertices;
} OrcConfig;

typedef struct OrcSim OrcSim;

/* rigid_points: n_rigid*3 floats (voxel centres before rotation/offset), may be NULL.
 * rigid_vertices: n_rigid_vertices*3 floats (mesh vertices), may be NULL. */
OrcSim *orc_create(const OrcConfig *cfg, const float *rigid_points, int n_rigid,
                   const float *rigid_vertices);
void orc_destroy(OrcSim *s);
void orc_set_threads(int n);

/* Derived sizes with the reference's host formulas (ParticleSystem.py:85-86,100-101,129-137). */
void orc_derived_sizes(const OrcConfig *cfg, long long *particle_num, long long *boundary_num,
                       int grid_num[3]);

/* One solver.step() (main.py:166-167). */
void orc_step(OrcSim *s);
/* One rigid_solver.step() (main.py:169-171); no-op without an active rigid body. */
void orc_rigid_step(OrcSim *s);

/* Individual phases, for single-substep parity.  Names follow the reference methods. */
void orc_base_step(OrcSim *s);                 /* solver_base.step: grid rebuild + reset */
void orc_phase(OrcSim *s, const char *name);   /* e.g. "initialize", "divergence_warm_start" */

/* Field access by name: returns 0 on success.  ptr is internal storage (valid until destroy). */
int orc_field(OrcSim *s, const char *name, void **ptr, long long *n, int *ncomp, int *is_float);
double orc_scalar(OrcSim *s, const char *name);
void orc_set_scalar(OrcSim *s, const char *name, double v);

/* Standalone kernel functions (known-answer tests). */
float orc_cubic_kernel(float r, float h);
void orc_cubic_kernel_derivative(const float r[3], float h, float out[3]);
float orc_poly_kernel(float r, float h);                                        /* SB:122-129 */
void orc_spiky_kernel_derivative(const float r[3], float h, float out[3]);      /* SB:113-120 */
/* Largest float t such that sqrtf(t) <= h (the sqrt-free cull threshold). */
float orc_cull_threshold(float h);

#ifdef __cplusplus
}
#endif
#endif
