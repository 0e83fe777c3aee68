/*
 * sph_oracle.c -- CPU restatement of the CFD_Taichi SPH hot path.  TEST INFRASTRUCTURE ONLY.
 * See sph_oracle.h for the contract (strict fp32, canonical order, "parity unpinned").
 *
 * Every function cites the reference file:line it follows (paths relative to the upstream
 * repository: ParticleSystem.py = PS, solver_base.py = SB, dfsph_solver.py = DF,
 * wcsph_solver.py = WC, pcisph_solver.py = PC, iisph_solver.py = II, rigid_solver.py = RS).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (see Makefile).
 * All arithmetic that the reference performs in ti.f32 is written with float operands only;
 * host-side (Python-scope) arithmetic of the reference is done in double and then cast.
 */
#include "sph_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAT_FLUID 0
#define MAT_BOUNDARY 1
#define MAT_SOLID 2

#define SOLVER_WCSPH 0
#define SOLVER_PCISPH 1
#define SOLVER_IISPH 2
#define SOLVER_DFSPH 3
#define SOLVER_PBF 4

static int g_threads = 1;

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

typedef struct { float x, y, z; } v3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float *p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline void st3(float *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 v3_add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 v3_scale(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }   /* s * v */
static inline v3 v3_mul(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }     /* v * s */
static inline v3 v3_div(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
/* Taichi vector ops: dot = (a0*b0 + a1*b1) + a2*b2 ; norm = sqrt(dot(a,a)). */
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float v3_norm(v3 a) { return sqrtf(v3_dot(a, a)); }
static inline v3 v3_cross(v3 a, v3 b) {
	return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

#define PI_F ((float)3.141592653589793)

/* Particle view = what get_particle (PS:496-507) returns, but as pointers (no 116-byte copy). */
typedef struct {
	const float *pos, *vel, *acc, *omega, *alpha;
	const int *cell3;
	float volume;
	int material;
	int index;
} PV;

struct OrcSim {
	OrcConfig cfg;
	/* Python-scope constants of the reference, cast to f32 where they meet f32 expressions */
	float h, d, r, m;
	double m_d, r_d, d_d, h_d;
	float gravity;
	int N, Nb, Nr, Nv;
	int gnum[3];
	long long G;
	int exist_rigid, active_rigid;
	int boundary_handle; /* 1 = akinci, 0 = clamp */
	int fs_couple;
	int solver;
	int error_flags;
	/* fluid */
	float *pos, *vel, *acc;
	int *cell3, *cell1;
	/* boundary */
	float *bpos, *bvol;
	int *bcell3;
	/* rigid */
	float *rpos, *rvel, *racc, *rforce, *romega, *ralpha, *rvol, *rmass, *rverts;
	int *rcell3;
	float centroid[3], inertia[9], inertia_inv[9];
	/* rigid solver state (RS:12-31) */
	float rs_dt, rs_omega[3], rs_attitude[3], rs_mass;
	int rs_run_once, rs_simulate_cnt;
	/* grids: CSR with canonical order (ascending global index per cell) */
	int *cell_start, *cell_items;
	int *bcell_start, *bcell_items;
	int *nbr_count;
	/* solver_base */
	float *rho, *viscosity, *tension;
	float dt, ps_dt;
	int simulate_cnt;
	float visc_cs, tension_k;
	/* dfsph */
	float *alpha, *rho_adv, *rho_derivative, *vel_adv, *vel_adv_delta, *force_ext, *warm_start_k;
	float dt2;
	int df_div_iters, df_den_iters;
	float df_div_first_err, df_div_err, df_den_err;
	/* wcsph */
	float *pressure, *pressure_gradient, *boundary_acc;
	/* pcisph */
	float *pos_predict, *vel_predict, *ext_force, *press_force, *rho_predict, *rho_err, *press_iter;
	float pc_delta;
	double pc_beta;
	int pc_iters, pc_max_index;
	float pc_err;
	/* iisph */
	float *v_adv, *f_adv, *d_ii, *a_ii, *d_ij, *p_iter, *p_past, *p_new_buff, *r_sum, *f_press;
	int ii_iters;
	float ii_residual;
	/* pbf (index-based semantics, sph_oracle_pbf.inc) */
	float *pbf_constrain, *pbf_lambda, *pbf_cd, *pbf_delta_pos;
	int pbf_update_mode;
	/* scratch */
	float *scratch;
};

/* ------------------------------------------------------------------------------------------
 * SPH kernel functions (SB:74-129)
 * ---------------------------------------------------------------------------------------- */

/* SB:74-88 cubic_kernel(r, h) */
static inline float cubic_kernel(float r, float h) {
	float ret = 0.0f;
	float q = r / h;
	float k = 8.0f / (PI_F * (h * (h * h)));
	if (0.0f <= q && q <= 0.5f) {
		float q2 = q * q;
		float q3 = q2 * q;
		ret = k * (6.0f * (q3 - q2) + 1.0f);
	} else if (0.5f < q && q <= 1.0f) {
		float t = 1.0f - q;
		ret = (2.0f * k) * (t * (t * t));
	} else {
		ret = 0.0f;
	}
	return ret;
}

/* SB:90-103 cubic_kernel_derivative(r, h) -- keeps the reference's extra factor 6 */
static inline v3 cubic_dw(v3 r, float h) {
	float r_norm = v3_norm(r);
	float q = r_norm / h;
	v3 ret = V3(0.0f, 0.0f, 0.0f);
	float k = 48.0f / (PI_F * (h * (h * h)));
	if (1e-5f < q && q <= 0.5f) {
		float q2 = q * q;
		float c = (k * 6.0f) * (3.0f * q2 - 2.0f * q);
		float den = h * r_norm;
		ret = v3_div(v3_scale(c, r), den);
	} else if (0.5f < q && q <= 1.0f) {
		float t = 1.0f - q;
		float c = ((-k) * 6.0f) * (t * t);
		float den = h * r_norm;
		ret = v3_div(v3_scale(c, r), den);
	}
	return ret;
}

/* SB:105-111 spiky_kernel ; SB:113-120 spiky_kernel_derivative ; SB:122-129 poly_kernel */
static inline float spiky_kernel(float r, float h) {
	float ret = 0.0f;
	float q = r / h;
	if (q <= 1.0f) {
		float t = 1.0f - q;
		ret = (15.0f * (t * (t * t))) / (((PI_F * h) * h) * h);
	}
	return ret;
}
static inline v3 spiky_dw(v3 r, float h) {
	float r_norm = v3_norm(r);
	float q = r_norm / h;
	v3 ret = V3(0.0f, 0.0f, 0.0f);
	if (q <= 1.0f && q > 0.0f) {
		float t = 1.0f - q;
		float h2 = h * h;
		float c = -(45.0f * (t * t));
		float den = (PI_F * (h2 * h2)) * r_norm;
		ret = v3_div(v3_scale(c, r), den);
	}
	return ret;
}
static inline float poly_kernel(float r, float h) {
	float q = r / h;
	float q2 = q * q;
	float ret = 0.0f;
	if (q <= 1.0f) {
		float t = 1.0f - q2;
		ret = (315.0f / ((64.0f * PI_F) * (h * (h * h)))) * (t * (t * t));
	}
	return ret;
}

float orc_cubic_kernel(float r, float h) { return cubic_kernel(r, h); }
float orc_poly_kernel(float r, float h) { return poly_kernel(r, h); }
void orc_spiky_kernel_derivative(const float r[3], float h, float out[3]) {
	v3 g = spiky_dw(V3(r[0], r[1], r[2]), h);
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}
void orc_cubic_kernel_derivative(const float r[3], float h, float out[3]) {
	st3(out, cubic_dw(ld3(r), h));
}
float orc_cull_threshold(float h) {
	/* largest t with sqrtf(t) <= h: start at h*h and walk ulps */
	float t = h * h;
	while (sqrtf(t) <= h) t = nextafterf(t, INFINITY);
	while (sqrtf(t) > h) t = nextafterf(t, -INFINITY);
	return t;
}

/* ------------------------------------------------------------------------------------------
 * Host-side derived sizes (PS:78-86, 100-101, 129-137) -- Python fp64 arithmetic
 * ---------------------------------------------------------------------------------------- */
void orc_derived_sizes(const OrcConfig *c, long long *particle_num, long long *boundary_num,
                       int grid_num[3]) {
	double r = c->particle_radius;
	double d = r * 2;
	double h = 4 * r;
	/* PS:85-86: int(wx / d * wy / d * wz / d), left to right */
	double pn = c->water_size[0] / d * c->water_size[1] / d * c->water_size[2] / d;
	*particle_num = (long long)pn;
	/* PS:129-137 */
	double bx = c->box_max[0] - c->box_min[0];
	double by = c->box_max[1] - c->box_min[1];
	double bz = c->box_max[2] - c->box_min[2];
	long long x_cnt = (long long)(bx / d + 1);
	long long z_cnt = (long long)(bz / d + 1);
	long long bottom = x_cnt * z_cnt;
	long long round_cnt = x_cnt * z_cnt - (x_cnt - 2) * (z_cnt - 2);
	long long layer = (long long)ceil((by - d) / d);
	*boundary_num = layer * round_cnt + bottom * 2;
	/* PS:100-101 */
	grid_num[0] = (int)ceil(bx / h) + 1;
	grid_num[1] = (int)ceil(by / h) + 1;
	grid_num[2] = (int)ceil(bz / h) + 1;
}

/* ------------------------------------------------------------------------------------------
 * Index helpers (PS:486-507)
 * ---------------------------------------------------------------------------------------- */
static inline long long cell1d(const OrcSim *s, int cx, int cy, int cz) {
	/* PS:102,486-488: dot with (1, gx*gz, gx) */
	return (long long)cx + (long long)cy * ((long long)s->gnum[0] * s->gnum[2]) + (long long)cz * s->gnum[0];
}
static inline void cell3_of(const OrcSim *s, const float *p, int out[3]) {
	/* PS:490-494: floor(pos / support_radius) with f32 true division, box_min ignored */
	out[0] = (int)floorf(p[0] / s->h);
	out[1] = (int)floorf(p[1] / s->h);
	out[2] = (int)floorf(p[2] / s->h);
}
static inline PV get_particle(const OrcSim *s, int gi) {
	PV v;
	if (gi < s->N) {
		v.pos = s->pos + 3 * gi; v.vel = s->vel + 3 * gi; v.acc = s->acc + 3 * gi;
		v.omega = NULL; v.alpha = NULL; v.cell3 = s->cell3 + 3 * gi;
		v.volume = 0.0f; v.material = MAT_FLUID; v.index = gi;
	} else if (gi < s->N + s->Nb) {
		int b = gi - s->N;
		v.pos = s->bpos + 3 * b; v.vel = NULL; v.acc = NULL; v.omega = NULL; v.alpha = NULL;
		v.cell3 = s->bcell3 + 3 * b; v.volume = s->bvol[b]; v.material = MAT_BOUNDARY; v.index = b;
	} else {
		int k = gi - s->N - s->Nb;
		v.pos = s->rpos + 3 * k; v.vel = s->rvel + 3 * k; v.acc = s->racc + 3 * k;
		v.omega = s->romega + 3 * k; v.alpha = s->ralpha + 3 * k; v.cell3 = s->rcell3 + 3 * k;
		v.volume = s->rvol[k]; v.material = MAT_SOLID; v.index = k;
	}
	return v;
}

/* PS:447-469 for_all_neighbor: canonical order, self skipped by GLOBAL index, cull norm > h */
#define FOR_NEIGHBORS(S, GI, PI, PJ, ...)                                                         \
	do {                                                                                          \
		PV PI = get_particle((S), (GI));                                                          \
		const int *cc_ = PI.cell3;                                                                \
		for (int dx_ = -1; dx_ <= 1; ++dx_)                                                       \
			for (int dy_ = -1; dy_ <= 1; ++dy_)                                                   \
				for (int dz_ = -1; dz_ <= 1; ++dz_) {                                             \
					int cx_ = cc_[0] + dx_, cy_ = cc_[1] + dy_, cz_ = cc_[2] + dz_;               \
					if (cx_ >= (S)->gnum[0] || cy_ >= (S)->gnum[1] || cz_ >= (S)->gnum[2]) continue; \
					if (!(cx_ >= 0 && cy_ >= 0 && cz_ >= 0)) continue;                            \
					long long c1_ = cell1d((S), cx_, cy_, cz_);                                   \
					for (int e_ = (S)->cell_start[c1_]; e_ < (S)->cell_start[c1_ + 1]; ++e_) {    \
						int nj_ = (S)->cell_items[e_];                                            \
						if (nj_ == (GI)) continue;                                                \
						PV PJ = get_particle((S), nj_);                                           \
						if (v3_norm(v3_sub(ld3(PI.pos), ld3(PJ.pos))) > (S)->h) continue;         \
						__VA_ARGS__                                                                \
					}                                                                             \
				}                                                                                 \
	} while (0)

/* PS:337-366 for_all_boundary_neighbor: GI is a global index (fluid i, or N + b); tasks get
 * (IL = local index of the centre, J = local boundary index). */
#define FOR_BOUNDARY_NEIGHBORS(S, GI, IL, J, ...)                                                 \
	do {                                                                                          \
		int IL = (GI);                                                                            \
		const int *cc_;                                                                           \
		const float *cp_;                                                                         \
		int same_ = 1;                                                                            \
		if (IL >= (S)->N) { IL -= (S)->N; cc_ = (S)->bcell3 + 3 * IL; cp_ = (S)->bpos + 3 * IL; } \
		else { same_ = 0; cc_ = (S)->cell3 + 3 * IL; cp_ = (S)->pos + 3 * IL; }                   \
		for (int dx_ = -1; dx_ <= 1; ++dx_)                                                       \
			for (int dy_ = -1; dy_ <= 1; ++dy_)                                                   \
				for (int dz_ = -1; dz_ <= 1; ++dz_) {                                             \
					int cx_ = cc_[0] + dx_, cy_ = cc_[1] + dy_, cz_ = cc_[2] + dz_;               \
					if (cx_ >= (S)->gnum[0] || cy_ >= (S)->gnum[1] || cz_ >= (S)->gnum[2]) continue; \
					if (!(cx_ >= 0 && cy_ >= 0 && cz_ >= 0)) continue;                            \
					long long c1_ = cell1d((S), cx_, cy_, cz_);                                   \
					for (int e_ = (S)->bcell_start[c1_]; e_ < (S)->bcell_start[c1_ + 1]; ++e_) {  \
						int J = (S)->bcell_items[e_];                                             \
						if (J == IL && same_ == 1) continue;                                      \
						if (v3_norm(v3_sub(ld3(cp_), ld3((S)->bpos + 3 * J))) > (S)->h) continue; \
						__VA_ARGS__                                                                \
					}                                                                             \
				}                                                                                 \
	} while (0)

#define NTHREADS(S) (((S)->Nr > 0 && (S)->active_rigid) ? 1 : g_threads)
#define PAR_FOR(S) _Pragma("omp parallel for schedule(dynamic, 256) num_threads(nth_)")
#define DECL_NTH(S) int nth_ = NTHREADS(S); (void)nth_

/* ------------------------------------------------------------------------------------------
 * Grid (PS:322-335, 368-407): CSR restatement of the per-cell dynamic lists.
 * Order inside a cell = arrival order with one thread = ascending index; fluid first, then
 * rigid (global index N+Nb+k) because update_grid_rigid_particles is a second kernel.
 * ---------------------------------------------------------------------------------------- */
static void build_boundary_grid(OrcSim *s) {
	long long G = s->G;
	memset(s->bcell_start, 0, sizeof(int) * (G + 1));
	int *c1 = (int *)malloc(sizeof(int) * (s->Nb > 0 ? s->Nb : 1));
	for (int i = 0; i < s->Nb; ++i) {
		int c[3];
		cell3_of(s, s->bpos + 3 * i, c);
		long long idx = cell1d(s, c[0], c[1], c[2]);
		s->bcell3[3 * i] = c[0]; s->bcell3[3 * i + 1] = c[1]; s->bcell3[3 * i + 2] = c[2];
		if (idx < 0 || idx >= G) { s->error_flags |= 2; c1[i] = -1; continue; }
		c1[i] = (int)idx;
		s->bcell_start[idx + 1]++;
	}
	for (long long c = 0; c < G; ++c) s->bcell_start[c + 1] += s->bcell_start[c];
	int *fill = (int *)calloc((size_t)(G > 0 ? G : 1), sizeof(int));
	for (int i = 0; i < s->Nb; ++i) {
		if (c1[i] < 0) continue;
		s->bcell_items[s->bcell_start[c1[i]] + fill[c1[i]]++] = i;
	}
	free(fill);
	free(c1);
}

static void reset_and_update_grid(OrcSim *s) {
	long long G = s->G;
	memset(s->cell_start, 0, sizeof(int) * (G + 1));
	/* PS:388-397 */
	for (int i = 0; i < s->N; ++i) {
		int c[3];
		cell3_of(s, s->pos + 3 * i, c);
		long long idx = cell1d(s, c[0], c[1], c[2]);
		/* PS:393: guard is `< 0 or > G`; idx == G would append out of bounds -> treat as error */
		if (idx < 0 || idx >= G) { s->error_flags |= 1; s->cell1[i] = -1; continue; }
		s->cell1[i] = (int)idx;
		s->cell3[3 * i] = c[0]; s->cell3[3 * i + 1] = c[1]; s->cell3[3 * i + 2] = c[2];
		s->cell_start[idx + 1]++;
	}
	int *rc1 = NULL;
	if (s->exist_rigid == 1 && s->Nr > 0) {
		/* PS:399-407 */
		rc1 = (int *)malloc(sizeof(int) * s->Nr);
		for (int k = 0; k < s->Nr; ++k) {
			rc1[k] = -1;
			if (s->active_rigid == 0) continue;
			int c[3];
			cell3_of(s, s->rpos + 3 * k, c);
			long long idx = cell1d(s, c[0], c[1], c[2]);
			if (idx < 0 || idx >= G) { s->error_flags |= 4; continue; }
			rc1[k] = (int)idx;
			s->rcell3[3 * k] = c[0]; s->rcell3[3 * k + 1] = c[1]; s->rcell3[3 * k + 2] = c[2];
			s->cell_start[idx + 1]++;
		}
	}
	for (long long c = 0; c < G; ++c) s->cell_start[c + 1] += s->cell_start[c];
	int *fill = (int *)calloc((size_t)(G > 0 ? G : 1), sizeof(int));
	for (int i = 0; i < s->N; ++i) {
		int c = s->cell1[i];
		if (c < 0) continue;
		s->cell_items[s->cell_start[c] + fill[c]++] = i;
	}
	if (rc1) {
		for (int k = 0; k < s->Nr; ++k) {
			int c = rc1[k];
			if (c < 0) continue;
			s->cell_items[s->cell_start[c] + fill[c]++] = s->N + s->Nb + k;
		}
		free(rc1);
	}
	free(fill);
}

/* PS:424-445 get_neighbour_count, including the index quirk for rigid entries (SURVEY B-7) */
static inline int get_neighbour_count(const OrcSim *s, int i) {
	int cnt = 0;
	int c[3];
	cell3_of(s, s->pos + 3 * i, c);
	for (int dx = -1; dx <= 1; ++dx)
		for (int dy = -1; dy <= 1; ++dy)
			for (int dz = -1; dz <= 1; ++dz) {
				int cx = c[0] + dx, cy = c[1] + dy, cz = c[2] + dz;
				if (cx >= s->gnum[0] || cy >= s->gnum[1] || cz >= s->gnum[2]) continue;
				if (!(cx >= 0 && cy >= 0 && cz >= 0)) continue;
				long long c1 = cell1d(s, cx, cy, cz);
				for (int e = s->cell_start[c1]; e < s->cell_start[c1 + 1]; ++e) {
					PV pj = get_particle(s, s->cell_items[e]);
					if (pj.index == i) continue;
					int jj = pj.index;
					if (jj >= s->N) jj = s->N - 1; /* reference would read out of bounds */
					if (v3_norm(v3_sub(ld3(s->pos + 3 * i), ld3(s->pos + 3 * jj))) > s->h) continue;
					cnt += 1;
				}
			}
	return cnt;
}

static void compute_all_neighbour_counts(OrcSim *s) {
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) s->nbr_count[i] = get_neighbour_count(s, i);
}

/* PS:409-422 get_max_neighbor_particle_index with one thread: atomic_max returns the OLD max */
static int get_max_neighbor_particle_index(OrcSim *s) {
	int max_count = -1, max_index = -1;
	compute_all_neighbour_counts(s);
	for (int i = 0; i < s->N; ++i) {
		int cnt = s->nbr_count[i];
		int old = max_count;
		if (cnt > max_count) max_count = cnt;
		if (old == cnt) max_index = i;
	}
	return max_index;
}

/* ------------------------------------------------------------------------------------------
 * Initialisation (PS:139-295)
 * ---------------------------------------------------------------------------------------- */
static inline float fmod_py(float a, float b) { return a - b * floorf(a / b); } /* App. A-4 */

static void init_particle_pos(OrcSim *s) {
	const OrcConfig *c = &s->cfg;
	/* PS:143-145: Python-scope fp64, cast to f32 when meeting the i32->f32 loop index */
	double x_num_d = c->water_size[0] / s->d_d;
	double z_num_d = c->water_size[2] / s->d_d;
	float x_num = (float)x_num_d, z_num = (float)z_num_d, xz_num = x_num * z_num; /* PS:145: product of two f32 locals */
	float sx = (float)c->start_pos[0], sy = (float)c->start_pos[1], sz = (float)c->start_pos[2];
	float rad = (float)s->r_d;
	long long xi_num = llround(x_num_d), zi_num = llround(z_num_d);
	for (int i = 0; i < s->N; ++i) {
		float x, z;
		int y;
		if (s->N < (1 << 24)) {
			float fi = (float)i;
			x = fmod_py(fi, x_num);                 /* PS:147 */
			z = fmod_py(floorf(fi / x_num), z_num); /* PS:148 */
			y = (int)(fi / xz_num);                 /* PS:149 */
		} else {
			/* beyond 2^24 the reference's f32 index arithmetic is inexact: integer lattice
			 *  */
			x = (float)(i % xi_num);
			z = (float)((i / xi_num) % zi_num);
			y = (int)(i / (xi_num * zi_num));
		}
		/* PS:150: Vector([x,y,z]) * r * 2 + start_pos */
		s->pos[3 * i + 0] = ((x * rad) * 2.0f) + sx;
		s->pos[3 * i + 1] = (((float)y * rad) * 2.0f) + sy;
		s->pos[3 * i + 2] = ((z * rad) * 2.0f) + sz;
	}
	/* PS:155-195 boundary shell */
	/* PS:155-158: `box = (box_max - box_min)` is a Python-scope (fp64) subtraction, but binding it to a kernel local makes
	 * it an f32 vector, so x_cnt / z_cnt are f32 arithmetic HERE while compute_boundary_particles_count (PS:129-137, host,
	 * fp64) sized the array: the two disagree for some boxes (5.2 / 0.05: 105 on the host, 104 in the kernel), and the
	 * reference then lays its shell out with the kernel's counts over the host's particle count (quirk B-18). */
	float bxf = (float)(c->box_max[0] - c->box_min[0]);
	float bzf = (float)(c->box_max[2] - c->box_min[2]);
	int x_cnt = (int)(bxf / s->d + 1.0f);
	int z_cnt = (int)(bzf / s->d + 1.0f);
	int x_cnt_round = x_cnt - 1;
	int z_cnt_round = z_cnt - 1;
	int bottom = x_cnt * z_cnt;
	int one_round = x_cnt * z_cnt - (x_cnt - 2) * (z_cnt - 2);
	float d = s->d;
	float box_max_y = (float)c->box_max[1];
	for (int i = 0; i < s->Nb; ++i) {
		float x = 0.0f, y = 0.0f, z = 0.0f;
		if (i < bottom) {
			x = (float)(i % x_cnt) * d;
			y = 0.0f;
			z = floorf((float)i / (float)x_cnt) * d;
		} else if (bottom <= i && i < s->Nb - bottom) {
			int index = i - bottom;
			int layer = (int)floorf((float)index / (float)one_round);
			y = d * (float)(layer + 1);
			index -= layer * one_round;
			index += 1;
			if (index <= x_cnt_round) {
				x = (float)(index % x_cnt_round) * d;
				z = 0.0f;
			} else if (x_cnt_round < index && index <= x_cnt_round + z_cnt_round) {
				x = (float)x_cnt_round * d;
				z = (float)((index - x_cnt) % z_cnt_round) * d;
			} else if (x_cnt_round + z_cnt_round < index && index <= 2 * x_cnt_round + z_cnt_round) {
				x = (float)((2 * x_cnt_round + z_cnt_round - index) % x_cnt_round + 1) * d;
				z = (float)z_cnt_round * d;
			} else if (2 * x_cnt_round + z_cnt_round < index && index <= 2 * (x_cnt_round + z_cnt_round)) {
				x = 0.0f;
				z = (float)((2 * (x_cnt_round + z_cnt_round) - index) % z_cnt_round + 1) * d;
			}
		} else {
			int index = i - (s->Nb - bottom);
			x = (float)(index % x_cnt) * d;
			y = box_max_y;
			z = (float)((int)((float)index / (float)x_cnt)) * d;
		}
		s->bpos[3 * i] = x; s->bpos[3 * i + 1] = y; s->bpos[3 * i + 2] = z;
	}
}

/* Taichi 1.6 ti.math.rotation3d(ang_x, ang_y, ang_z) as recalled in SURVEY App. A-11:
 * rot_yaw_pitch_roll(yaw = ang_z, pitch = ang_x, roll = ang_y), GLM yawPitchRoll entries
 * written row-major.  UNVERIFIABLE offline. */
static void rotation3d(float ang_x, float ang_y, float ang_z, float R[9]) {
	float yaw = ang_z, pitch = ang_x, roll = ang_y;
	float ch = cosf(yaw), sh = sinf(yaw);
	float cp = cosf(pitch), sp = sinf(pitch);
	float cb = cosf(roll), sb = sinf(roll);
	R[0] = ch * cb + sh * sp * sb; R[1] = sb * cp;  R[2] = -sh * cb + ch * sp * sb;
	R[3] = -ch * sb + sh * sp * cb; R[4] = cb * cp; R[5] = sb * sh + ch * sp * cb;
	R[6] = sh * cp;                 R[7] = -sp;     R[8] = ch * cp;
}
static inline v3 mat3_mul_v(const float M[9], v3 a) {
	return V3((M[0] * a.x + M[1] * a.y) + M[2] * a.z, (M[3] * a.x + M[4] * a.y) + M[5] * a.z,
	          (M[6] * a.x + M[7] * a.y) + M[8] * a.z);
}
static void mat3_mul(const float A[9], const float B[9], float C[9]) {
	float T[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			T[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
	memcpy(C, T, sizeof(T));
}
static void mat3_transpose(const float A[9], float T[9]) {
	float t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
	memcpy(T, t, sizeof(t));
}
static void mat3_inverse(const float m[9], float inv[9]) {
	/* Taichi's Matrix.inverse() for n = 3 as recalled: inv_determinant = 1.0 / determinant() first,
	 * then entries[j][i] = inv_determinant * (E(i+1,j+1) E(i+2,j+2) - E(i+2,j+1) E(i+1,j+2)) with E(x,y) = a(x % 3, y % 3);
	 * determinant() = a00 (a11 a22 - a21 a12) - a10 (a01 a22 - a21 a02) + a20 (a01 a12 - a11 a02) */
#define MI_E(x, y) m[3 * ((x) % 3) + ((y) % 3)]
	float det = (MI_E(0, 0) * (MI_E(1, 1) * MI_E(2, 2) - MI_E(2, 1) * MI_E(1, 2)) -
	             MI_E(1, 0) * (MI_E(0, 1) * MI_E(2, 2) - MI_E(2, 1) * MI_E(0, 2))) +
	            MI_E(2, 0) * (MI_E(0, 1) * MI_E(1, 2) - MI_E(1, 1) * MI_E(0, 2));
	float inv_det = 1.0f / det;
	float t[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			t[3 * j + i] = inv_det * (MI_E(i + 1, j + 1) * MI_E(i + 2, j + 2) - MI_E(i + 2, j + 1) * MI_E(i + 1, j + 2));
#undef MI_E
	memcpy(inv, t, sizeof(t));
}

/* PS:198-223 */
static void init_rigid_particles_pos(OrcSim *s) {
	const OrcConfig *c = &s->cfg;
	const double pi_d = 3.141592653589793;
	float ax = (float)(c->rigid_att_offset_deg[0] / 180.0 * pi_d);
	float ay = (float)(c->rigid_att_offset_deg[1] / 180.0 * pi_d);
	float az = (float)(c->rigid_att_offset_deg[2] / 180.0 * pi_d);
	float R[9];
	rotation3d(ax, az, ay, R); /* PS:200: rotation3d(att.x, att.z, att.y) */
	v3 off = V3((float)c->rigid_pos_offset[0], (float)c->rigid_pos_offset[1], (float)c->rigid_pos_offset[2]);
	for (int i = 0; i < s->Nr; ++i) {
		v3 p = mat3_mul_v(R, ld3(s->rpos + 3 * i));
		st3(s->rpos + 3 * i, v3_add(p, off));
	}
	for (int i = 0; i < s->Nv; ++i) {
		v3 p = mat3_mul_v(R, ld3(s->rverts + 3 * i));
		st3(s->rverts + 3 * i, v3_add(p, off));
	}
}

/* PS:309-320 compute_all_boundary_volume */
static void compute_all_boundary_volume(OrcSim *s) {
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->Nb; ++i) {
		float volume = 0.0f;
		FOR_BOUNDARY_NEIGHBORS(s, i + s->N, il, j, {
			float q = v3_norm(v3_sub(ld3(s->bpos + 3 * il), ld3(s->bpos + 3 * j)));
			volume += cubic_kernel(q, s->h);
		});
		s->bvol[i] = 1.0f / volume;
	}
}

/* PS:249-295 init_rigid_particles_data */
static void init_rigid_particles_data(OrcSim *s) {
	int off = s->N + s->Nb;
	for (int i = 0; i < s->Nr; ++i) {
		float volume = 0.0f;
		if (s->active_rigid) { /* inactive rigid particles are not in the grid (SURVEY B-R2) */
			FOR_NEIGHBORS(s, i + off, pi, pj, {
				if (pj.material == MAT_SOLID) { /* PS:301-307 */
					float q = v3_norm(v3_sub(ld3(pi.pos), ld3(pj.pos)));
					volume += cubic_kernel(q, s->h);
				}
			});
		}
		s->rvol[i] = (volume < 1e-6f) ? 0.0f : 1.0f / volume;
	}
	float rrho = (float)s->cfg.rigid_rho;
	for (int i = 0; i < s->Nr; ++i) s->rmass[i] = rrho * s->rvol[i];
	v3 cen = V3(0, 0, 0);
	float sum_mass = 0.0f;
	for (int i = 0; i < s->Nr; ++i) {
		cen = v3_add(cen, v3_mul(ld3(s->rpos + 3 * i), s->rmass[i]));
		sum_mass += s->rmass[i];
	}
	st3(s->centroid, v3_div(cen, sum_mass));
	float Ixx = 0, Iyy = 0, Izz = 0, Ixy = 0, Ixz = 0, Iyz = 0;
	for (int i = 0; i < s->Nr; ++i) {
		v3 p = v3_sub(ld3(s->rpos + 3 * i), ld3(s->centroid));
		float mi = s->rmass[i];
		Ixx += mi * (p.y * p.y + p.z * p.z);
		Iyy += mi * (p.x * p.x + p.z * p.z);
		Izz += mi * (p.x * p.x + p.y * p.y);
		Ixy += (-mi) * (p.x * p.y);
		Ixz += (-mi) * (p.x * p.z);
		Iyz += (-mi) * (p.z * p.y);
	}
	float I[9] = {Ixx, Ixy, Ixz, Ixy, Iyy, Iyz, Ixz, Iyz, Izz};
	memcpy(s->inertia, I, sizeof(I));
	mat3_inverse(I, s->inertia_inv);
}

/* ------------------------------------------------------------------------------------------
 * solver_base sweeps (SB:41-72, 170-217)
 * ---------------------------------------------------------------------------------------- */

/* SB:41-51 compute_all_rho ; SB:58-66 compute_rho ; SB:68-72 compute_rho_from_boundary */
static void compute_all_rho(OrcSim *s) {
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		float rho = 0.001f;
		FOR_NEIGHBORS(s, i, pi, pj, {
			float ret = 0.0f;
			if (pj.material == MAT_FLUID) {
				ret = s->m * cubic_kernel(v3_norm(v3_sub(ld3(pi.pos), ld3(pj.pos))), s->h);
			} else if (pj.material == MAT_SOLID) {
				if (s->fs_couple == 1)
					ret = (pj.volume * cubic_kernel(v3_norm(v3_sub(ld3(pi.pos), ld3(pj.pos))), s->h)) * 1000.0f;
			}
			rho += ret;
		});
		if (s->boundary_handle == 1) {
			float rho_boundary = 0.0f;
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				float q = v3_norm(v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j)));
				rho_boundary += s->bvol[j] * cubic_kernel(q, s->h);
			});
			s->rho[i] = rho + rho_boundary * 1000.0f;
		} else {
			s->rho[i] = rho;
		}
	}
}

/* SB:170-202 solve_all_viscosity / compute_viscosity */
static void solve_all_viscosity(OrcSim *s) {
	/* Python-scope constants folded in fp64 (App. A-2) */
	float nu_num = (float)(2 * 0.08 * s->h_d * (double)s->visc_cs);
	float eps_h2 = (float)(0.01 * s->h_d * s->h_d);
	float neg_m = (float)(-s->m_d);
	float neg_rho0 = -1000.0f;
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 visc = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			if (pj.material == MAT_FLUID || (pj.material == MAT_SOLID && s->fs_couple == 1)) {
				v3 v_ij = v3_sub(ld3(pi.vel), ld3(pj.vel));
				v3 x_ij = v3_sub(ld3(pi.pos), ld3(pj.pos));
				float shear = v3_dot(v_ij, x_ij);
				if (shear < 0.0f) {
					float q = v3_norm(x_ij);
					float q2 = q * q;
					int jr = pj.index; /* SB:187,199: rho[particle_j.index] also for rigid j (B-6) */
					if (jr >= s->N) jr = s->N - 1;
					float nu = nu_num / (s->rho[pi.index] + s->rho[jr]);
					float pi_ij = ((-nu) * shear) / (q2 + eps_h2);
					v3 dw = cubic_dw(x_ij, s->h);
					if (pj.material == MAT_FLUID)
						ret = v3_add(ret, v3_scale(neg_m * pi_ij, dw));
					else
						ret = v3_add(ret, v3_scale((neg_rho0 * pj.volume) * pi_ij, dw));
				}
			}
			visc = v3_add(visc, ret);
		});
		st3(s->viscosity + 3 * i, v3_mul(visc, s->m));
	}
}

/* SB:204-217 solve_all_tension / compute_tension */
static void solve_all_tension(OrcSim *s) {
	/* - tension_k / particle_m * particle_m : all Python-scope -> fp64, then f32 */
	float coef = (float)(-(double)s->tension_k / s->m_d * s->m_d);
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 ten = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			if (pj.material == MAT_FLUID) {
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				float w = cubic_kernel(v3_norm(q), s->h);
				ten = v3_add(ten, v3_scale(coef * w, q));
			}
		});
		st3(s->tension + 3 * i, v3_mul(ten, s->m));
	}
}

/* clamp boundary (boundary_handle == false): DF:241-250, PC:78-87/208-217, II:194-203 use
 * particle_radius; WC:54-63 uses particle_diameter. */
static void clamp_box(OrcSim *s, float *pos, float *vel, double margin) {
	for (int i = 0; i < s->N; ++i)
		for (int j = 0; j < 3; ++j) {
			float lo = (float)(s->cfg.box_min[j] + margin);
			float hi = (float)(s->cfg.box_max[j] - margin);
			if (pos[3 * i + j] <= lo) { pos[3 * i + j] = lo; vel[3 * i + j] *= -0.5f; }
			if (pos[3 * i + j] >= hi) { pos[3 * i + j] = hi; vel[3 * i + j] *= -0.5f; }
		}
}

/* SB:136-143 solver_base.step: counter, grid rebuild, reset() */
static void solver_reset(OrcSim *s);
void orc_base_step(OrcSim *s) {
	s->simulate_cnt += 1;
	reset_and_update_grid(s);
	solver_reset(s);
}

/* ------------------------------------------------------------------------------------------
 * DFSPH (DF:32-445)
 * ---------------------------------------------------------------------------------------- */

/* DF:32-89 compute_all_alpha */
static void df_compute_all_alpha(OrcSim *s) {
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 sum_square = V3(0, 0, 0);
		float square_sum = 0.0f;
		float denominator = 0.0f;
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
			if (pj.material == MAT_FLUID) ret = v3_scale(s->m, cubic_dw(q, s->h));
			else if (pj.material == MAT_SOLID && s->fs_couple == 1) ret = v3_scale(pj.volume * 1000.0f, cubic_dw(q, s->h));
			sum_square = v3_add(sum_square, ret);
		});
		FOR_NEIGHBORS(s, i, pi, pj, {
			float ans = 0.0f;
			v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
			if (pj.material == MAT_FLUID) { v3 ret = v3_scale(s->m, cubic_dw(q, s->h)); ans = v3_dot(ret, ret); }
			else if (pj.material == MAT_SOLID && s->fs_couple == 1) { v3 ret = v3_scale(pj.volume * 1000.0f, cubic_dw(q, s->h)); ans = v3_dot(ret, ret); }
			square_sum += ans;
		});
		if (s->boundary_handle == 1) {
			v3 ssb = V3(0, 0, 0);
			float sqb = 0.0f;
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				ssb = v3_add(ssb, v3_scale(s->bvol[j] * 1000.0f, cubic_dw(q, s->h)));
			});
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				v3 ret = v3_scale(s->bvol[j] * 1000.0f, cubic_dw(q, s->h));
				sqb += v3_dot(ret, ret);
			});
			denominator = ((v3_dot(sum_square, sum_square) + square_sum) + sqb) + v3_dot(ssb, ssb);
		} else {
			denominator = v3_dot(sum_square, sum_square) + square_sum;
		}
		if (fabsf(denominator) < 1e-6f) s->alpha[i] = 0.0f;
		else s->alpha[i] = s->rho[i] / denominator;
	}
}

/* DF:423-426 initialize */
static void df_initialize(OrcSim *s) {
	compute_all_rho(s);
	df_compute_all_alpha(s);
}

/* DF:314-355 divergence_warm_start */
static void df_divergence_warm_start(OrcSim *s) {
	DECL_NTH(s);
	float dt = s->dt;
	float *newvel = s->scratch;
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 va = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			if (pj.material == MAT_FLUID) {
				int ii = pi.index, jj = pj.index;
				float k_i = s->warm_start_k[ii] / dt;
				float k_j = s->warm_start_k[jj] / dt;
				v3 q = v3_sub(ld3(s->pos + 3 * ii), ld3(s->pos + 3 * jj));
				ret = v3_scale(s->m * (k_i / s->rho[ii] + k_j / s->rho[jj]), cubic_dw(q, s->h));
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				int ii = pi.index;
				float k_i = s->warm_start_k[ii] / dt;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = v3_scale(((pj.volume * 1000.0f) * k_i) / s->rho[ii], cubic_dw(q, s->h));
			}
			va = v3_add(va, ret);
		});
		v3 v = ld3(s->vel + 3 * i);
		if (s->boundary_handle == 1) {
			v3 vb = V3(0, 0, 0);
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				float k_i = s->warm_start_k[il] / dt;
				vb = v3_add(vb, v3_scale((s->bvol[j] * k_i) / s->rho[il], cubic_dw(q, s->h)));
			});
			v = v3_sub(v, v3_mul(v3_add(va, v3_mul(vb, 1000.0f)), dt));
		} else {
			v = v3_sub(v, v3_mul(va, dt));
		}
		st3(newvel + 3 * i, v);
	}
	memcpy(s->vel, newvel, sizeof(float) * 3 * s->N);
	memset(s->warm_start_k, 0, sizeof(float) * s->N); /* DF:325 */
}

/* DF:252-300 derivative_iter_all_rho */
static float df_derivative_iter_all_rho(OrcSim *s) {
	DECL_NTH(s);
	float dt = s->dt;
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		int nc = get_neighbour_count(s, i);
		s->nbr_count[i] = nc;
		if (nc < 20) { s->rho_derivative[i] = 0.0f; continue; }
		float rd = 0.0f;
		FOR_NEIGHBORS(s, i, pi, pj, {
			float ret = 0.0f;
			if (pj.material == MAT_FLUID) {
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = s->m * v3_dot(v3_sub(ld3(pi.vel), ld3(pj.vel)), cubic_dw(q, s->h));
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				v3 kernel = cubic_dw(q, s->h);
				v3 v_omega = v3_cross(ld3(pj.omega), v3_sub(ld3(pj.pos), ld3(s->centroid)));
				v3 v_j = v3_add(v3_add(ld3(pj.vel), v3_mul(ld3(pj.acc), dt)), v_omega);
				ret = (pj.volume * 1000.0f) * v3_dot(v3_sub(ld3(pi.vel), v_j), kernel);
			}
			rd += ret;
		});
		if (s->boundary_handle == 1) {
			float rdb = 0.0f;
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				rdb += s->bvol[j] * v3_dot(ld3(s->vel + 3 * il), cubic_dw(q, s->h));
			});
			s->rho_derivative[i] = fmaxf(rd + rdb * 1000.0f, 0.0f);
		} else {
			s->rho_derivative[i] = fmaxf(rd, 0.0f);
		}
	}
	float avg = 0.0f, ret = 0.0f;
	int cnt = 0;
	for (int i = 0; i < s->N; ++i)
		if (s->rho_derivative[i] > 0.0f) { cnt += 1; avg += s->rho_derivative[i]; }
	if (cnt > 0) ret = avg / (float)cnt;
	return ret;
}

/* DF:302-312, 357-391 divergence_iter_all_vel_adv */
static void df_divergence_iter_all_vel_adv(OrcSim *s) {
	DECL_NTH(s);
	float dt = s->dt;
	float *newvel = s->scratch;
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 va = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			if (pj.material == MAT_FLUID) {
				int ii = pi.index, jj = pj.index;
				float k_i = (s->rho_derivative[ii] * s->alpha[ii]) / dt;
				float k_j = (s->rho_derivative[jj] * s->alpha[jj]) / dt;
				v3 q = v3_sub(ld3(s->pos + 3 * ii), ld3(s->pos + 3 * jj));
				v3 kernel = cubic_dw(q, s->h);
				float f = k_i / s->rho[ii] + k_j / s->rho[jj];
				if (f > 1e-5f) ret = v3_scale(s->m * f, kernel);
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				int ii = pi.index;
				float k_i = (s->rho_derivative[ii] * s->alpha[ii]) / dt;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = v3_scale(((pj.volume * 1000.0f) * k_i) / s->rho[ii], cubic_dw(q, s->h));
			}
			va = v3_add(va, ret);
		});
		v3 v = ld3(s->vel + 3 * i);
		if (s->boundary_handle == 1) {
			v3 vb = V3(0, 0, 0);
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				float k_i = (s->rho_derivative[il] * s->alpha[il]) / dt;
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				vb = v3_add(vb, v3_scale((s->bvol[j] * k_i) / s->rho[il], cubic_dw(q, s->h)));
			});
			v = v3_sub(v, v3_mul(v3_add(va, v3_mul(vb, 1000.0f)), dt));
		} else {
			v = v3_sub(v, v3_mul(va, dt));
		}
		st3(newvel + 3 * i, v);
	}
	memcpy(s->vel, newvel, sizeof(float) * 3 * s->N);
}

/* DF:381-384 sum_up_stiff */
static void df_sum_up_stiff(OrcSim *s) {
	for (int i = 0; i < s->N; ++i) s->warm_start_k[i] += s->rho_derivative[i] * s->alpha[i];
}

/* DF:393-416 correct_divergence_error (host loop) */
static void df_correct_divergence_error(OrcSim *s) {
	float past = 0.0f;
	int iter_cnt = 0;
	df_divergence_warm_start(s);
	float avg = df_derivative_iter_all_rho(s);
	float first_err = avg;
	while ((iter_cnt < 1 || avg > 10.0f) && iter_cnt < 15) {
		df_divergence_iter_all_vel_adv(s);
		df_sum_up_stiff(s);
		past = avg;
		avg = df_derivative_iter_all_rho(s);
		/* host-side: Python floats (fp64) holding f32 values */
		if (fabs((double)avg - (double)past) < 1e-5) break;
		iter_cnt += 1;
	}
	s->df_div_iters = iter_cnt;
	s->df_div_first_err = first_err;
	s->df_div_err = avg;
}

/* DF:91-96 compute_all_ext_force */
static void df_compute_all_ext_force(OrcSim *s) {
	solve_all_tension(s);
	solve_all_viscosity(s);
	float g = s->gravity;
	v3 gv = V3(g * 0.0f, g * -1.0f, g * 0.0f);
	for (int i = 0; i < s->N; ++i)
		st3(s->force_ext + 3 * i, v3_add(v3_add(gv, ld3(s->tension + 3 * i)), ld3(s->viscosity + 3 * i)));
}

/* DF:98-122 compute_all_vel_adv (adaptive dt) */
static void df_compute_all_vel_adv(OrcSim *s) {
	float max_vel = -INFINITY;
	float dt = s->dt;
	for (int i = 0; i < s->N; ++i) {
		v3 va = v3_add(ld3(s->vel + 3 * i), v3_div(v3_scale(dt, ld3(s->force_ext + 3 * i)), s->m));
		st3(s->vel_adv + 3 * i, va);
		float n = v3_norm(va);
		if (n > max_vel) max_vel = n;
	}
	float max_rigid_vel = 0.0f;
	for (int k = 0; k < s->Nr; ++k) {
		v3 pos = ld3(s->rpos + 3 * k), vel = ld3(s->rvel + 3 * k), omega = ld3(s->romega + 3 * k);
		float n = v3_norm(vel) + v3_norm(v3_cross(omega, v3_sub(pos, ld3(s->centroid))));
		if (n > max_rigid_vel) max_rigid_vel = n;
	}
	max_vel += max_rigid_vel;
	float c1 = (float)(0.4 * s->r_d * 2);
	float max_delta_time = (c1 / max_vel) * 0.2f;
	if (max_delta_time > 1e-3f) s->dt = 1e-3f;
	else s->dt = fmaxf(max_delta_time, 1e-5f);
	s->dt2 = s->dt * s->dt;
	s->ps_dt = s->dt;
}

/* DF:124-176 compute_all_rho_adv */
static float df_compute_all_rho_adv(OrcSim *s) {
	DECL_NTH(s);
	float dt = s->dt;
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		float delta = 0.0f;
		FOR_NEIGHBORS(s, i, pi, pj, {
			float ret = 0.0f;
			if (pj.material == MAT_FLUID) {
				int ii = pi.index, jj = pj.index;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				v3 kernel = cubic_dw(q, s->h);
				ret = s->m * v3_dot(v3_sub(ld3(s->vel_adv + 3 * ii), ld3(s->vel_adv + 3 * jj)), kernel);
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				int ii = pi.index;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				v3 kernel = cubic_dw(q, s->h);
				v3 v_omega = v3_cross(v3_add(ld3(pj.omega), v3_mul(ld3(pj.alpha), dt)), v3_sub(ld3(pj.pos), ld3(s->centroid)));
				v3 v_j = v3_add(v3_add(ld3(pj.vel), v3_mul(ld3(pj.acc), dt)), v_omega);
				ret = (pj.volume * 1000.0f) * v3_dot(v3_sub(ld3(s->vel_adv + 3 * ii), v_j), kernel);
			}
			delta += ret;
		});
		if (s->boundary_handle == 1) {
			float db = 0.0f;
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				db += s->bvol[j] * v3_dot(ld3(s->vel_adv + 3 * il), cubic_dw(q, s->h));
			});
			s->rho_adv[i] = fmaxf(s->rho[i] + dt * (delta + db * 1000.0f), 1000.0f);
		} else {
			s->rho_adv[i] = fmaxf(s->rho[i] + dt * delta, 1000.0f);
		}
	}
	float rho_avg = 0.0f, ret = 1000.0f;
	int cnt = 0;
	for (int i = 0; i < s->N; ++i)
		if (!(s->rho_adv[i] == 1000.0f)) { rho_avg += s->rho_adv[i]; cnt += 1; }
	if (cnt > 0) ret = rho_avg / (float)cnt;
	return ret;
}

/* DF:178-219 iter_all_vel_adv (rigid force scatter DF:212) */
static void df_iter_all_vel_adv(OrcSim *s) {
	DECL_NTH(s);
	float dt = s->dt, dt2 = s->dt2;
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 va = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			if (pj.material == MAT_FLUID) {
				int ii = pi.index, jj = pj.index;
				float k_i = ((s->rho_adv[ii] - 1000.0f) * s->alpha[ii]) / dt2;
				float k_j = ((s->rho_adv[jj] - 1000.0f) * s->alpha[jj]) / dt2;
				v3 q = v3_sub(ld3(s->pos + 3 * ii), ld3(s->pos + 3 * jj));
				ret = v3_scale(s->m * (k_i / s->rho[ii] + k_j / s->rho[jj]), cubic_dw(q, s->h));
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				int ii = pi.index, jj = pj.index;
				float k_i = ((s->rho_adv[ii] - 1000.0f) * s->alpha[ii]) / dt2;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = v3_scale(((pj.volume * 1000.0f) * k_i) / s->rho[ii], cubic_dw(q, s->h));
				v3 f = v3_add(ld3(s->rforce + 3 * jj), v3_mul(ret, s->m));
				st3(s->rforce + 3 * jj, f);
			}
			va = v3_add(va, ret);
		});
		if (s->boundary_handle == 1) {
			v3 vb = V3(0, 0, 0);
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				float k_i = ((s->rho_adv[il] - 1000.0f) * s->alpha[il]) / dt2;
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				vb = v3_add(vb, v3_scale((s->bvol[j] * k_i) / s->rho[il], cubic_dw(q, s->h)));
			});
			st3(s->vel_adv_delta + 3 * i, v3_add(va, v3_mul(vb, 1000.0f)));
		} else {
			st3(s->vel_adv_delta + 3 * i, va);
		}
	}
	for (int i = 0; i < s->N; ++i)
		st3(s->vel_adv + 3 * i, v3_sub(ld3(s->vel_adv + 3 * i), v3_mul(ld3(s->vel_adv_delta + 3 * i), dt)));
}

/* DF:221-233 correct_density_error (no iteration cap in the reference; 1000 is a safety net) */
static void df_correct_density_error(OrcSim *s) {
	double rho_avg = INFINITY;
	int iter_cnt = 0;
	while (iter_cnt < 2 || rho_avg - 1000 > 0.1 * 1000 * 0.01) {
		rho_avg = (double)df_compute_all_rho_adv(s);
		df_iter_all_vel_adv(s);
		iter_cnt += 1;
		if (iter_cnt >= 1000) { s->error_flags |= 8; break; }
	}
	s->df_den_iters = iter_cnt;
	s->df_den_err = (float)(rho_avg - 1000);
}

/* DF:235-250 compute_all_position */
static void df_compute_all_position(OrcSim *s) {
	float dt = s->dt;
	for (int i = 0; i < s->N; ++i) {
		v3 va = ld3(s->vel_adv + 3 * i);
		st3(s->pos + 3 * i, v3_add(ld3(s->pos + 3 * i), v3_mul(v3_scale(dt, va), 0.9999f)));
		st3(s->vel + 3 * i, v3_mul(va, 0.9999f));
	}
	if (s->boundary_handle == 0) clamp_box(s, s->pos, s->vel, s->r_d);
}

static void df_iterate(OrcSim *s) {
	df_correct_divergence_error(s);
	df_compute_all_ext_force(s);
	df_compute_all_vel_adv(s);
	df_correct_density_error(s);
	df_compute_all_position(s);
}

/* ------------------------------------------------------------------------------------------
 * WCSPH (WC:25-129)
 * ---------------------------------------------------------------------------------------- */
static void wc_pressure_phase(OrcSim *s) {
	compute_all_rho(s);
	/* WC:65-68, 86-90 solve_p: B * ((max(rho, rho0)/rho0) ** 7 - 1), pow by squaring (App. A-5) */
	for (int i = 0; i < s->N; ++i) {
		float rho_i = fmaxf(s->rho[i], 1000.0f);
		float a = rho_i / 1000.0f;
		float a2 = a * a, a3 = a * a2, a4 = a2 * a2;
		s->pressure[i] = 70000.0f * (a3 * a4 - 1.0f);
	}
	/* WC:70-84, 92-129 */
	DECL_NTH(s);
	PAR_FOR(s)
	for (int i = 0; i < s->N; ++i) {
		v3 acc = V3(0, 0, 0);
		FOR_NEIGHBORS(s, i, pi, pj, {
			v3 ret = V3(0, 0, 0);
			if (pj.material == MAT_FLUID) {
				int ii = pi.index, jj = pj.index;
				float rho_i = s->rho[ii];
				float rho_i_2 = rho_i * rho_i;
				float p_i = s->pressure[ii], p_j = s->pressure[jj];
				float rho_j = s->rho[jj];
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = v3_sub(ret, v3_scale(s->m * (p_i / rho_i_2 + p_j / (rho_j * rho_j)), cubic_dw(q, s->h)));
			} else if (pj.material == MAT_SOLID && s->fs_couple == 1) {
				int ii = pi.index;
				float p_i = s->pressure[ii];
				float rho_i = s->rho[ii];
				float rho_i_2 = rho_i * rho_i;
				v3 q = v3_sub(ld3(pi.pos), ld3(pj.pos));
				ret = v3_mul(v3_scale(((-pj.volume) * p_i) / rho_i_2, cubic_dw(q, s->h)), 1000.0f);
				v3 f = v3_add(ld3(s->rforce + 3 * pj.index), v3_mul(v3_neg(ret), s->m));
				st3(s->rforce + 3 * pj.index, f);
			}
			acc = v3_add(acc, ret);
		});
		if (s->boundary_handle == 1) {
			v3 bacc = V3(0, 0, 0);
			FOR_BOUNDARY_NEIGHBORS(s, i, il, j, {
				float p_i = s->pressure[il];
				float rho_i = s->rho[il];
				float rho_i_2 = rho_i * rho_i;
				v3 q = v3_sub(ld3(s->pos + 3 * il), ld3(s->bpos + 3 * j));
				v3 ret = v3_sub(V3(0, 0, 0), v3_scale((s->bvol[j] * p_i) / rho_i_2, cubic_dw(q, s->h)));
				bacc = v3_add(bacc, ret);
			});
			st3(s->boundary_acc + 3 * i, v3_mul(bacc, 1000.0f));
		}
		st3(s->pressure_gradient + 3 * i, acc);
	}
	solve_all_viscosity(s);
	solve_all_tension(s);
}

/* WC:40-63 kinematic_phase */
static void wc_kinematic_phase(OrcSim *s) {
	float dt = s->dt;
	for (int i = 0; i < s->N; ++i) {
		v3 a = ld3(s->acc + 3 * i);
		v3 add = v3_add(v3_add(ld3(s->pressure_gradient + 3 * i), ld3(s->viscosity + 3 * i)), ld3(s->tension + 3 * i));
		if (s->boundary_handle == 1) add = v3_add(add, ld3(s->boundary_acc + 3 * i));
		st3(s->acc + 3 * i, v3_add(a, add));
	}
	for (int i = 0; i < s->N; ++i) {
		v3 v = v3_add(ld3(s->vel + 3 * i), v3_mul(ld3(s->acc + 3 * i), dt));
		v = v3_mul(v, 0.9998f);
		st3(s->vel + 3 * i, v);
		st3(s->pos + 3 * i, v3_add(ld3(s->pos + 3 * i), v3_mul(v, dt)));
	}
	if (s->boundary_handle == 0) clamp_box(s, s->pos, s->vel, s->d_d);
}

#include "sph_oracle_solvers2.inc"
#include "sph_oracle_pbf.inc"

/* per-solver reset() (SB:131-134 ; DF:418-421 ; PC:228-231 ; II:31-33) */
static void solver_reset(OrcSim *s) {
	if (s->solver == SOLVER_WCSPH || s->solver == SOLVER_PBF) {
		float g = s->gravity;
		for (int i = 0; i < s->N; ++i) st3(s->acc + 3 * i, V3(g * 0.0f, g * -1.0f, g * 0.0f));
	} else if (s->solver == SOLVER_PCISPH) {
		memset(s->press_iter, 0, sizeof(float) * s->N);
		memset(s->press_force, 0, sizeof(float) * 3 * s->N);
	}
}

void orc_step(OrcSim *s) {
	orc_base_step(s);
	switch (s->solver) {
	case SOLVER_DFSPH: df_initialize(s); df_iterate(s); break;
	case SOLVER_WCSPH: wc_pressure_phase(s); wc_kinematic_phase(s); break;
	case SOLVER_PCISPH: pc_compute_ext_force(s); pc_iteration(s); pc_integration(s); break;
	case SOLVER_IISPH: ii_predict_advection(s); ii_pressure_solve(s); ii_integration(s); break;
	case SOLVER_PBF: pbf_step(s); break;
	default: break;
	}
}

void orc_phase(OrcSim *s, const char *name) {
#define PH(n, call) if (!strcmp(name, n)) { call; return; }
	PH("reset_grid_update_grid", reset_and_update_grid(s))
	PH("compute_all_rho", compute_all_rho(s))
	PH("solve_all_viscosity", solve_all_viscosity(s))
	PH("solve_all_tension", solve_all_tension(s))
	PH("neighbour_counts", compute_all_neighbour_counts(s))
	PH("initialize", df_initialize(s))
	PH("divergence_warm_start", df_divergence_warm_start(s))
	PH("derivative_iter_all_rho", s->df_div_err = df_derivative_iter_all_rho(s))
	PH("divergence_iter_all_vel_adv", df_divergence_iter_all_vel_adv(s))
	PH("sum_up_stiff", df_sum_up_stiff(s))
	PH("correct_divergence_error", df_correct_divergence_error(s))
	PH("compute_all_ext_force", df_compute_all_ext_force(s))
	PH("compute_all_vel_adv", df_compute_all_vel_adv(s))
	PH("compute_all_rho_adv", s->df_den_err = df_compute_all_rho_adv(s) - 1000.0f)
	PH("iter_all_vel_adv", df_iter_all_vel_adv(s))
	PH("correct_density_error", df_correct_density_error(s))
	PH("compute_all_position", df_compute_all_position(s))
	PH("iterate", df_iterate(s))
	PH("pressure_phase", wc_pressure_phase(s))
	PH("kinematic_phase", wc_kinematic_phase(s))
	PH("pc_compute_ext_force", pc_compute_ext_force(s))
	PH("pc_iteration", pc_iteration(s))
	PH("pc_integration", pc_integration(s))
	PH("ii_predict_advection", ii_predict_advection(s))
	PH("ii_pressure_solve", ii_pressure_solve(s))
	PH("ii_integration", ii_integration(s))
	PH("ii_compute_all_d_ij", ii_compute_all_d_ij(s))
	PH("ii_update_p", ii_update_p(s))
	PH("pbf_externel_force_predict_pos", pbf_externel_force_predict_pos(s))
	PH("pbf_compute_all_lambda", pbf_compute_all_lambda(s))
	PH("pbf_compute_all_delta_pos", pbf_compute_all_delta_pos(s))
	PH("pbf_update_all_pos", pbf_update_all_pos(s))
#undef PH
	fprintf(stderr, "orc_phase: unknown phase '%s'\n", name);
	s->error_flags |= 1024;
}

/* ------------------------------------------------------------------------------------------
 * Construction (PS:31-127 + solver __init__)
 * ---------------------------------------------------------------------------------------- */
static float *falloc(long long n) { return (float *)calloc(n > 0 ? n : 1, sizeof(float)); }
static int *ialloc(long long n) { return (int *)calloc(n > 0 ? n : 1, sizeof(int)); }

OrcSim *orc_create(const OrcConfig *cfg, const float *rigid_points, int n_rigid, const float *rigid_vertices) {
	OrcSim *s = (OrcSim *)calloc(1, sizeof(OrcSim));
	s->cfg = *cfg;
	s->r_d = cfg->particle_radius;
	s->d_d = s->r_d * 2;                  /* PS:81 */
	s->h_d = 4 * s->r_d;                  /* PS:82 */
	s->m_d = 1000 * pow(s->r_d, 3) * 8;  /* PS:83: Python float ** int = libm pow */
	s->r = (float)s->r_d; s->d = (float)s->d_d; s->h = (float)s->h_d; s->m = (float)s->m_d;
	s->gravity = (float)cfg->gravity;
	long long pn, bn;
	orc_derived_sizes(cfg, &pn, &bn, s->gnum);
	s->N = (int)pn; s->Nb = (int)bn;
	s->exist_rigid = cfg->exist_rigid ? 1 : 0;
	s->active_rigid = (cfg->exist_rigid && cfg->active_rigid) ? 1 : 0;
	s->Nr = s->exist_rigid ? n_rigid : 0;
	s->Nv = s->exist_rigid ? cfg->n_rigid_vertices : 0;
	s->G = (long long)s->gnum[0] * s->gnum[1] * s->gnum[2];
	s->boundary_handle = cfg->boundary_handle ? 1 : 0;
	s->fs_couple = cfg->fs_couple ? 1 : 0;
	s->solver = cfg->solver;
	s->dt = (float)cfg->delta_time;
	s->dt2 = (float)(cfg->delta_time * cfg->delta_time); /* DF:20: Python-scope f32 field read ** 2 */
	{
		/* DF:20 reads the f32 field back into Python (fp64), squares, stores as f32 */
		double dtf = (double)s->dt;
		s->dt2 = (float)(dtf * dtf);
	}
	s->ps_dt = 0.0f;
	s->visc_cs = (s->solver == SOLVER_WCSPH) ? 10.0f : 13.0f;    /* SB:24 / WC:18 */
	s->tension_k = (s->solver == SOLVER_WCSPH) ? 0.2f : 0.5f;    /* SB:26 / WC:20 */

	int N = s->N, Nb = s->Nb, Nr = s->Nr;
	s->pos = falloc(3LL * N); s->vel = falloc(3LL * N); s->acc = falloc(3LL * N);
	s->cell3 = ialloc(3LL * N); s->cell1 = ialloc(N);
	s->bpos = falloc(3LL * Nb); s->bvol = falloc(Nb); s->bcell3 = ialloc(3LL * Nb);
	s->rpos = falloc(3LL * Nr); s->rvel = falloc(3LL * Nr); s->racc = falloc(3LL * Nr);
	s->rforce = falloc(3LL * Nr); s->romega = falloc(3LL * Nr); s->ralpha = falloc(3LL * Nr);
	s->rvol = falloc(Nr); s->rmass = falloc(Nr); s->rcell3 = ialloc(3LL * Nr);
	s->rverts = falloc(3LL * s->Nv);
	if (Nr > 0 && rigid_points) memcpy(s->rpos, rigid_points, sizeof(float) * 3 * Nr);
	if (s->Nv > 0 && rigid_vertices) memcpy(s->rverts, rigid_vertices, sizeof(float) * 3 * s->Nv);
	s->cell_start = ialloc(s->G + 1); s->cell_items = ialloc((long long)N + Nr);
	s->bcell_start = ialloc(s->G + 1); s->bcell_items = ialloc(Nb);
	s->nbr_count = ialloc(N);
	s->rho = falloc(N); s->viscosity = falloc(3LL * N); s->tension = falloc(3LL * N);
	s->scratch = falloc(3LL * N);
	s->alpha = falloc(N); s->rho_adv = falloc(N); s->rho_derivative = falloc(N);
	s->vel_adv = falloc(3LL * N); s->vel_adv_delta = falloc(3LL * N); s->force_ext = falloc(3LL * N);
	s->warm_start_k = falloc(N);
	s->pressure = falloc(N); s->pressure_gradient = falloc(3LL * N); s->boundary_acc = falloc(3LL * N);
	s->pos_predict = falloc(3LL * N); s->vel_predict = falloc(3LL * N); s->ext_force = falloc(3LL * N);
	s->press_force = falloc(3LL * N); s->rho_predict = falloc(N); s->rho_err = falloc(N); s->press_iter = falloc(N);
	s->v_adv = falloc(3LL * N); s->f_adv = falloc(3LL * N); s->d_ii = falloc(3LL * N); s->a_ii = falloc(N);
	s->d_ij = falloc(3LL * N); s->p_iter = falloc(N); s->p_past = falloc(N); s->p_new_buff = falloc(N);
	s->r_sum = falloc(N); s->f_press = falloc(3LL * N);
	s->pbf_constrain = falloc(N); s->pbf_lambda = falloc(N); s->pbf_cd = falloc(3LL * N); s->pbf_delta_pos = falloc(3LL * N);

	init_particle_pos(s);                                   /* PS:119 */
	if (s->exist_rigid == 1) init_rigid_particles_pos(s);   /* PS:120-121 */
	/* PS:225-247 init_particles_data */
	build_boundary_grid(s);
	reset_and_update_grid(s);
	compute_all_boundary_volume(s);
	if (s->exist_rigid) init_rigid_particles_data(s);

	/* solver __init__ */
	s->rs_dt = (float)cfg->delta_time;
	if (s->solver == SOLVER_PCISPH) pc_init(s);
	return s;
}

void orc_destroy(OrcSim *s) {
	if (!s) return;
	float **fp[] = {&s->pos, &s->vel, &s->acc, &s->bpos, &s->bvol, &s->rpos, &s->rvel, &s->racc, &s->rforce,
	                &s->romega, &s->ralpha, &s->rvol, &s->rmass, &s->rverts, &s->rho, &s->viscosity, &s->tension,
	                &s->scratch, &s->alpha, &s->rho_adv, &s->rho_derivative, &s->vel_adv, &s->vel_adv_delta,
	                &s->force_ext, &s->warm_start_k, &s->pressure, &s->pressure_gradient, &s->boundary_acc,
	                &s->pos_predict, &s->vel_predict, &s->ext_force, &s->press_force, &s->rho_predict, &s->rho_err,
	                &s->press_iter, &s->v_adv, &s->f_adv, &s->d_ii, &s->a_ii, &s->d_ij, &s->p_iter, &s->p_past,
	                &s->p_new_buff, &s->r_sum, &s->f_press, &s->pbf_constrain, &s->pbf_lambda, &s->pbf_cd, &s->pbf_delta_pos};
	for (size_t i = 0; i < sizeof(fp) / sizeof(fp[0]); ++i) free(*fp[i]);
	int **ip[] = {&s->cell3, &s->cell1, &s->bcell3, &s->rcell3, &s->cell_start, &s->cell_items, &s->bcell_start,
	              &s->bcell_items, &s->nbr_count};
	for (size_t i = 0; i < sizeof(ip) / sizeof(ip[0]); ++i) free(*ip[i]);
	free(s);
}

/* ------------------------------------------------------------------------------------------
 * Field / scalar access by name
 * ---------------------------------------------------------------------------------------- */
int orc_field(OrcSim *s, const char *name, void **ptr, long long *n, int *ncomp, int *is_float) {
#define FLD(nm, p, cnt, nc, isf) if (!strcmp(name, nm)) { *ptr = (void *)(p); *n = (cnt); *ncomp = (nc); *is_float = (isf); return 0; }
	FLD("pos", s->pos, s->N, 3, 1) FLD("vel", s->vel, s->N, 3, 1) FLD("acc", s->acc, s->N, 3, 1)
	FLD("cell3", s->cell3, s->N, 3, 0) FLD("cell1", s->cell1, s->N, 1, 0)
	FLD("bpos", s->bpos, s->Nb, 3, 1) FLD("bvol", s->bvol, s->Nb, 1, 1) FLD("bcell3", s->bcell3, s->Nb, 3, 0)
	FLD("rpos", s->rpos, s->Nr, 3, 1) FLD("rvel", s->rvel, s->Nr, 3, 1) FLD("racc", s->racc, s->Nr, 3, 1)
	FLD("rforce", s->rforce, s->Nr, 3, 1) FLD("romega", s->romega, s->Nr, 3, 1) FLD("ralpha", s->ralpha, s->Nr, 3, 1)
	FLD("rvol", s->rvol, s->Nr, 1, 1) FLD("rmass", s->rmass, s->Nr, 1, 1) FLD("rverts", s->rverts, s->Nv, 3, 1)
	FLD("centroid", s->centroid, 1, 3, 1) FLD("inertia", s->inertia, 1, 9, 1) FLD("inertia_inv", s->inertia_inv, 1, 9, 1)
	FLD("rs_omega", s->rs_omega, 1, 3, 1) FLD("rs_attitude", s->rs_attitude, 1, 3, 1)
	FLD("cell_start", s->cell_start, s->G + 1, 1, 0) FLD("cell_items", s->cell_items, s->cell_start[s->G], 1, 0)
	FLD("bcell_start", s->bcell_start, s->G + 1, 1, 0) FLD("bcell_items", s->bcell_items, s->bcell_start[s->G], 1, 0)
	FLD("nbr_count", s->nbr_count, s->N, 1, 0)
	FLD("rho", s->rho, s->N, 1, 1) FLD("viscosity", s->viscosity, s->N, 3, 1) FLD("tension", s->tension, s->N, 3, 1)
	FLD("alpha", s->alpha, s->N, 1, 1) FLD("rho_adv", s->rho_adv, s->N, 1, 1)
	FLD("rho_derivative", s->rho_derivative, s->N, 1, 1) FLD("vel_adv", s->vel_adv, s->N, 3, 1)
	FLD("vel_adv_delta", s->vel_adv_delta, s->N, 3, 1) FLD("force_ext", s->force_ext, s->N, 3, 1)
	FLD("warm_start_k", s->warm_start_k, s->N, 1, 1)
	FLD("pressure", s->pressure, s->N, 1, 1) FLD("pressure_gradient", s->pressure_gradient, s->N, 3, 1)
	FLD("boundary_acc", s->boundary_acc, s->N, 3, 1)
	FLD("pos_predict", s->pos_predict, s->N, 3, 1) FLD("vel_predict", s->vel_predict, s->N, 3, 1)
	FLD("ext_force", s->ext_force, s->N, 3, 1) FLD("press_force", s->press_force, s->N, 3, 1)
	FLD("rho_predict", s->rho_predict, s->N, 1, 1) FLD("rho_err", s->rho_err, s->N, 1, 1)
	FLD("press_iter", s->press_iter, s->N, 1, 1)
	FLD("v_adv", s->v_adv, s->N, 3, 1) FLD("f_adv", s->f_adv, s->N, 3, 1) FLD("d_ii", s->d_ii, s->N, 3, 1)
	FLD("a_ii", s->a_ii, s->N, 1, 1) FLD("d_ij", s->d_ij, s->N, 3, 1) FLD("p_iter", s->p_iter, s->N, 1, 1)
	FLD("p_past", s->p_past, s->N, 1, 1) FLD("r_sum", s->r_sum, s->N, 1, 1) FLD("f_press", s->f_press, s->N, 3, 1)
	FLD("pbf_constrain", s->pbf_constrain, s->N, 1, 1) FLD("pbf_lambda", s->pbf_lambda, s->N, 1, 1)
	FLD("pbf_constrain_derivative", s->pbf_cd, s->N, 3, 1) FLD("pbf_delta_pos", s->pbf_delta_pos, s->N, 3, 1)
#undef FLD
	return -1;
}

double orc_scalar(OrcSim *s, const char *name) {
#define SC(nm, v) if (!strcmp(name, nm)) return (double)(v);
	SC("particle_num", s->N) SC("boundary_particles_num", s->Nb) SC("rigid_particles_num", s->Nr)
	SC("grid_x", s->gnum[0]) SC("grid_y", s->gnum[1]) SC("grid_z", s->gnum[2]) SC("grid_count", s->G)
	SC("particle_m", s->m_d) SC("support_radius", s->h_d) SC("delta_time", s->dt) SC("delta_time_2", s->dt2)
	SC("ps_delta_time", s->ps_dt) SC("simulate_cnt", s->simulate_cnt) SC("error_flags", s->error_flags)
	SC("df_div_iters", s->df_div_iters) SC("df_den_iters", s->df_den_iters)
	SC("df_div_first_err", s->df_div_first_err) SC("df_div_err", s->df_div_err) SC("df_den_err", s->df_den_err)
	SC("pc_delta", s->pc_delta) SC("pc_beta", s->pc_beta) SC("pc_iters", s->pc_iters) SC("pc_err", s->pc_err)
	SC("pc_max_index", s->pc_max_index)
	SC("ii_iters", s->ii_iters) SC("ii_residual", s->ii_residual)
	SC("rs_mass", s->rs_mass) SC("rs_dt", s->rs_dt) SC("active_rigid", s->active_rigid)
#undef SC
	return NAN;
}

void orc_set_scalar(OrcSim *s, const char *name, double v) {
	if (!strcmp(name, "delta_time")) { s->dt = (float)v; return; }
	if (!strcmp(name, "delta_time_2")) { s->dt2 = (float)v; return; }
	if (!strcmp(name, "active_rigid")) { s->active_rigid = (int)v; return; }
	if (!strcmp(name, "pbf_update_mode")) { s->pbf_update_mode = (int)v; return; }
}
