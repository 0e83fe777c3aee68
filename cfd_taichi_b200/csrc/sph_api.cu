// sph_api.cu -- the C-ABI of include/sph_b200.h: handle life cycle, constant folding that mirrors the
// reference's Python-scope arithmetic, scratch allocation, step / phase drivers and host-buffer
// (e2e) entry points.  No torch types; one host thread per handle.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "sph_internal.h"

static char g_create_error[512] = "";

int sph_fail(SphHandle *h, int code, const char *fmt, ...) {
	char *dst = h ? h->err : g_create_error;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(dst, 512, fmt, ap);
	va_end(ap);
	return code;
}

int sph_fail_cuda(SphHandle *h, cudaError_t e, const char *expr, const char *file, int line) {
	return sph_fail(h, SPH_ECUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, expr);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#define PI_F ((float)3.141592653589793)

// Python-scope constant folding of the reference (fp64 first, then one cast to f32).
static void fill_consts(const SphConfig &cfg, SphConsts &c) {
	memset(&c, 0, sizeof(c));
	double r_d = cfg.particle_radius;
	double d_d = r_d * 2;                 // PS:81
	double h_d = 4 * r_d;                 // PS:82 (== SB:17 kernel_h)
	double m_d = 1000 * pow(r_d, 3) * 8;  // PS:83
	bool wc = cfg.solver == SPH_SOLVER_WCSPH;
	double c_s = wc ? 10 : 13;            // SB:24 / WC:18
	double k_t = wc ? 0.2 : 0.5;          // SB:26 / WC:20
	c.h = (float)h_d;
	c.inv_h = 1.0f / c.h;
	c.r = (float)r_d;
	c.d = (float)d_d;
	c.m = (float)m_d;
	{
		// largest f32 t with sqrtf(t) <= h: sqrt(r2) > h  <=>  r2 > t
		float t = c.h * c.h;
		while (sqrtf(t) <= c.h) t = nextafterf(t, INFINITY);
		while (sqrtf(t) > c.h) t = nextafterf(t, -INFINITY);
		c.cull_t = t;
	}
	float h3 = c.h * (c.h * c.h);
	c.kW = 8.0f / (PI_F * h3);    // SB:79
	c.kDW = 48.0f / (PI_F * h3);  // SB:95
	c.kDW6 = c.kDW * 6.0f;        // SB:98
	c.nkDW6 = (-c.kDW) * 6.0f;    // SB:100
	c.dwA = (float)((double)c.kDW6 / (h_d * h_d));
	c.dwB = (float)((double)c.nkDW6 / h_d);
	{
		// |grad W| <= kDW6 / (3 h) (at q = 1/3; the outer branch stays below kDW6 / (4 h)): 21-bit fixed point over
		// [-S, S] with S 0.1 % above the bound, step S / (2^20 - 1)
		double S = (double)c.kDW6 / (3.0 * h_d) * 1.001;
		c.gq_scale = (float)(S / 1048575.0);
		c.gq_inv = (float)(1048575.0 / S);
	}
	c.gravity = (float)cfg.gravity;
	c.visc_num = (float)(2 * 0.08 * h_d * c_s);   // SB:187
	c.visc_eps_h2 = (float)(0.01 * h_d * h_d);    // SB:188
	c.neg_m = (float)(-m_d);                      // SB:189
	c.tension_coef = (float)(-k_t / m_d * m_d);   // SB:216
	c.dt_cfl_c1 = (float)(0.4 * r_d * 2);         // DF:112
	double margin = wc ? d_d : r_d;               // WC:57 vs DF:244 / PC:81 / II:197
	for (int k = 0; k < 3; ++k) {
		c.clamp_lo[k] = (float)(cfg.box_min[k] + margin);
		c.clamp_hi[k] = (float)(cfg.box_max[k] - margin);
	}
	{
		double dtf = (double)(float)cfg.delta_time; // f32 field read back into Python (PC:23)
		c.pc_beta = (float)(dtf * dtf * m_d * m_d * 2 / (1000.0 * 1000.0));
	}
	c.gx = cfg.grid_num[0];
	c.gy = cfg.grid_num[1];
	c.gz = cfg.grid_num[2];
	c.gxz = c.gx * c.gz;
	c.G = c.gx * c.gy * c.gz;
	c.N = cfg.n_fluid;
	c.N_owned = cfg.n_fluid;
	c.Nb = cfg.n_boundary;
	c.Nr = cfg.n_rigid;
	c.kmax = cfg.max_neighbors > 0 ? cfg.max_neighbors : 96;
	c.kbmax = cfg.max_boundary_neighbors > 0 ? cfg.max_boundary_neighbors : 48;
	c.kstride = (c.kmax + 3) & ~3;   // quad-interleaved lists (sph_list_word)
	c.kbstride = (c.kbmax + 3) & ~3;
	c.boundary_handle = cfg.boundary_handle ? 1 : 0;
	c.fs_couple = cfg.fs_couple ? 1 : 0;
	c.solver = cfg.solver;
	c.active_rigid = cfg.active_rigid ? 1 : 0;
}

template <typename T>
static cudaError_t dalloc(T **p, size_t n) {
	*p = nullptr;
	if (n == 0) n = 1;
	return cudaMalloc((void **)p, n * sizeof(T));
}

static int alloc_grid(SphHandle *h, SphGrid &g, size_t n, size_t G) {
	SPH_CUDA_CHECK(h, dalloc(&g.cell_of, n));
	SPH_CUDA_CHECK(h, dalloc(&g.cell_cnt, G + 1));
	SPH_CUDA_CHECK(h, dalloc(&g.cell_start, G + 1));
	SPH_CUDA_CHECK(h, dalloc(&g.sorted_id, n));
	SPH_CUDA_CHECK(h, dalloc(&g.scell, n));
	SPH_CUDA_CHECK(h, dalloc(&g.slot_of, n));
	SPH_CUDA_CHECK(h, cudaMemset(g.cell_start, 0, sizeof(int) * (G + 1)));
	g.n = 0;
	return SPH_OK;
}

static void free_grid(SphGrid &g) {
	cudaFree(g.cell_of); cudaFree(g.cell_cnt); cudaFree(g.cell_start); cudaFree(g.sorted_id); cudaFree(g.scell); cudaFree(g.slot_of);
}

extern "C" int sph_abi_version(void) { return 1; }

extern "C" const char *sph_last_error(const SphHandle *h) { return h ? h->err : g_create_error; }

extern "C" int sph_create(const SphConfig *cfg, int device, SphHandle **out) {
	if (!cfg || !out) return sph_fail(nullptr, SPH_EINVAL, "sph_create: null argument");
	*out = nullptr;
	if (cfg->n_fluid < 0 || cfg->n_boundary < 0 || cfg->n_rigid < 0)
		return sph_fail(nullptr, SPH_EINVAL, "sph_create: negative particle count");
	if (cfg->grid_num[0] <= 0 || cfg->grid_num[1] <= 0 || cfg->grid_num[2] <= 0)
		return sph_fail(nullptr, SPH_EINVAL, "sph_create: grid_num must be positive");
	if ((long long)cfg->grid_num[0] * cfg->grid_num[1] * cfg->grid_num[2] > 0x7fffffffLL - 2)
		return sph_fail(nullptr, SPH_EINVAL, "sph_create: grid too large for int32 cell ids");
	if (cfg->solver < SPH_SOLVER_WCSPH || cfg->solver > SPH_SOLVER_PBF)
		return sph_fail(nullptr, SPH_EINVAL, "sph_create: unknown solver id %d", cfg->solver);
	if (!(cfg->particle_radius > 0)) return sph_fail(nullptr, SPH_EINVAL, "sph_create: particle_radius must be > 0");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0)
		return sph_fail(nullptr, SPH_ECUDA, "sph_create: no CUDA device (%s); this library has no CPU path",
		                cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return sph_fail(nullptr, SPH_EINVAL, "sph_create: bad device %d", device);
	e = cudaSetDevice(device);
	if (e != cudaSuccess) return sph_fail(nullptr, SPH_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

	SphHandle *h = new (std::nothrow) SphHandle();
	if (!h) return sph_fail(nullptr, SPH_ENOMEM, "sph_create: out of host memory");
	memset(h, 0, sizeof(*h));
	h->cfg = *cfg;
	h->device = device;
	fill_consts(*cfg, h->c);
	SphConsts &c = h->c;
	size_t ncap = (size_t)cfg->n_fluid + (size_t)(cfg->n_ghost_capacity > 0 ? cfg->n_ghost_capacity : 0);
	size_t G = (size_t)c.G;
	int rc;
	if ((rc = alloc_grid(h, h->fg, ncap, G)) != SPH_OK) { *out = h; return rc; }
	if ((rc = alloc_grid(h, h->bg, (size_t)c.Nb, G)) != SPH_OK) { *out = h; return rc; }
	if ((rc = alloc_grid(h, h->rg, (size_t)c.Nr, c.Nr > 0 ? G : 1)) != SPH_OK) { *out = h; return rc; }
	*out = h;
	h->scan_sums_cap = cdiv((int)G, 2048) + 2;
	SPH_CUDA_CHECK(h, dalloc(&h->scan_sums, (size_t)h->scan_sums_cap));
	SPH_CUDA_CHECK(h, dalloc(&h->bspos, (size_t)c.Nb));
	SPH_CUDA_CHECK(h, dalloc(&h->rspos, (size_t)c.Nr));
	SPH_CUDA_CHECK(h, dalloc(&h->rstate, 1));
	SPH_CUDA_CHECK(h, cudaMemset(h->rstate, 0, sizeof(SphRigidState)));
	h->rl_cap = 96;
	SPH_CUDA_CHECK(h, dalloc(&h->rl_list, (size_t)((c.Nr + 31) / 32) * 32 * (size_t)h->rl_cap));
	SPH_CUDA_CHECK(h, dalloc(&h->rl_count, (size_t)c.Nr));
	for (int k = 0; k < A4_COUNT; ++k) {
		SPH_CUDA_CHECK(h, dalloc(&h->a4[k], ncap));
		SPH_CUDA_CHECK(h, cudaMemset(h->a4[k], 0, sizeof(float4) * (ncap ? ncap : 1)));
	}
	for (int k = 0; k < A1_COUNT; ++k) {
		SPH_CUDA_CHECK(h, dalloc(&h->a1[k], ncap));
		SPH_CUDA_CHECK(h, cudaMemset(h->a1[k], 0, sizeof(float) * (ncap ? ncap : 1)));
	}
	size_t nwarps = (ncap + 31) / 32;
	SPH_CUDA_CHECK(h, dalloc(&h->L.flist, nwarps * 32 * (size_t)c.kstride));
	SPH_CUDA_CHECK(h, dalloc(&h->L.blist, nwarps * 32 * (size_t)c.kbstride));
	h->L.gw = nullptr;
	h->L.gq = nullptr;
	if (c.solver == SPH_SOLVER_DFSPH) {
		// per-pair gradient stream of the step: strict kernels keep the exact float4 records, fast kernels 8 bytes per pair
		if (cfg->strict) SPH_CUDA_CHECK(h, dalloc(&h->L.gw, nwarps * 32 * (size_t)c.kstride));
		else SPH_CUDA_CHECK(h, dalloc(&h->L.gq, nwarps * 32 * (size_t)(c.kstride >> 1)));
	}
	SPH_CUDA_CHECK(h, dalloc(&h->L.fcount, ncap));
	SPH_CUDA_CHECK(h, dalloc(&h->L.bcount, ncap));
	SPH_CUDA_CHECK(h, cudaMemset(h->L.fcount, 0, sizeof(int) * (ncap ? ncap : 1)));
	SPH_CUDA_CHECK(h, cudaMemset(h->L.bcount, 0, sizeof(int) * (ncap ? ncap : 1)));
	SPH_CUDA_CHECK(h, dalloc(&h->nbr_count, ncap));
	SPH_CUDA_CHECK(h, dalloc(&h->ctl, 1));
	h->L.err = &h->ctl->error_flags;
	h->L.n_fluid = c.N; h->L.n_rigid = c.Nr; h->L.n_boundary = c.Nb; h->L.cap_f = c.kstride; h->L.cap_b = c.kbstride;
	SPH_CUDA_CHECK(h, cudaMallocHost((void **)&h->ctl_host, sizeof(SphCtl)));
	h->n_partials = cdiv((int)(ncap ? ncap : 1), SPH_BLOCK) + 1;
	SPH_CUDA_CHECK(h, dalloc(&h->partials, (size_t)h->n_partials));
	SPH_CUDA_CHECK(h, dalloc(&h->red, 4));
	memset(h->ctl_host, 0, sizeof(SphCtl));
	h->ctl_host->dt = (float)cfg->delta_time;        // SB:16
	{
		double dtf = (double)h->ctl_host->dt;        // DF:20 delta_time[None] ** 2 in Python scope
		h->ctl_host->dt2 = (float)(dtf * dtf);
	}
	h->ctl_host->ps_dt = 0.0f;                       // PS:37
	SPH_CUDA_CHECK(h, cudaMemcpy(h->ctl, h->ctl_host, sizeof(SphCtl), cudaMemcpyHostToDevice));
	{
		SphRigidState rs;
		memset(&rs, 0, sizeof(rs));
		rs.rs_dt = (float)cfg->delta_time; // RS:13
		rs.active = c.active_rigid;
		SPH_CUDA_CHECK(h, cudaMemcpy(h->rstate, &rs, sizeof(rs), cudaMemcpyHostToDevice));
	}
	snprintf(h->err, sizeof(h->err), "ok");
	return SPH_OK;
}

extern "C" int sph_destroy(SphHandle *h) {
	if (!h) return SPH_OK;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	free_grid(h->fg); free_grid(h->bg); free_grid(h->rg);
	cudaFree(h->scan_sums); cudaFree(h->bspos); cudaFree(h->rspos); cudaFree(h->rstate); cudaFree(h->rl_list); cudaFree(h->rl_count);
	for (int k = 0; k < A4_COUNT; ++k) cudaFree(h->a4[k]);
	for (int k = 0; k < A1_COUNT; ++k) cudaFree(h->a1[k]);
	cudaFree(h->render_zbuf); cudaFree(h->L.flist); cudaFree(h->L.blist); cudaFree(h->L.gw); cudaFree(h->L.gq);
	cudaFree(h->L.fcount); cudaFree(h->L.bcount);
	mg_destroy(h);
	cudaFree(h->nbr_count); cudaFree(h->ctl); cudaFree(h->partials); cudaFree(h->red);
	if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
	if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
	cudaFree(h->xyz_stage);
	if (h->copy_stream) { cudaStreamDestroy(h->copy_stream); cudaEventDestroy(h->ev_vel_ready); cudaEventDestroy(h->ev_mark); }
	if (h->ctl_host) cudaFreeHost(h->ctl_host);
	if (h->prof) {
		if (h->prof->created)
			for (int i = 0; i < SPH_PROF_CAP; ++i) { cudaEventDestroy(h->prof->e0[i]); cudaEventDestroy(h->prof->e1[i]); }
		delete h->prof;
	}
	delete h;
	return SPH_OK;
}

extern "C" int sph_bind(SphHandle *h, int field, void *dev_ptr, size_t n) {
	if (!h) return SPH_EINVAL;
	if (!dev_ptr && n > 0) return sph_fail(h, SPH_EINVAL, "sph_bind: null pointer for field %d", field);
	if (((uintptr_t)dev_ptr & 15u) != 0 && field != SPH_F_FLUID_GID) return sph_fail(h, SPH_EINVAL, "sph_bind: field %d is not 16-byte aligned", field);
	size_t ncap = (size_t)h->cfg.n_fluid + (size_t)(h->cfg.n_ghost_capacity > 0 ? h->cfg.n_ghost_capacity : 0);
	if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; } // the graph holds the old pointers
	switch (field) {
	case SPH_F_FLUID_POS:
		if (n < ncap) return sph_fail(h, SPH_EINVAL, "sph_bind: fluid pos needs %zu float4, got %zu", ncap, n);
		h->pos = (float4 *)dev_ptr; h->n_pos = n; break;
	case SPH_F_FLUID_VEL:
		if (n < ncap) return sph_fail(h, SPH_EINVAL, "sph_bind: fluid vel needs %zu float4, got %zu", ncap, n);
		h->vel = (float4 *)dev_ptr; h->n_vel = n; break;
	case SPH_F_FLUID_ACC:
		if (n < ncap) return sph_fail(h, SPH_EINVAL, "sph_bind: fluid acc needs %zu float4, got %zu", ncap, n);
		h->acc = (float4 *)dev_ptr; h->n_acc = n; break;
	case SPH_F_BOUNDARY_POS:
		if (n < (size_t)h->c.Nb) return sph_fail(h, SPH_EINVAL, "sph_bind: boundary pos needs %d float4", h->c.Nb);
		h->bpos = (float4 *)dev_ptr; h->n_bpos = n; h->boundary_ready = false; break;
	case SPH_F_RIGID_POS:
		if (n < (size_t)h->c.Nr) return sph_fail(h, SPH_EINVAL, "sph_bind: rigid pos needs %d float4", h->c.Nr);
		h->rpos = (float4 *)dev_ptr; h->n_rpos = n; break;
	case SPH_F_RIGID_VEL:
		if (n < (size_t)h->c.Nr) return sph_fail(h, SPH_EINVAL, "sph_bind: rigid vel needs %d float4", h->c.Nr);
		h->rvel = (float4 *)dev_ptr; h->n_rvel = n; break;
	case SPH_F_RIGID_FORCE:
		if (n < (size_t)h->c.Nr) return sph_fail(h, SPH_EINVAL, "sph_bind: rigid force needs %d float4", h->c.Nr);
		h->rforce = (float4 *)dev_ptr; h->n_rforce = n; break;
	case SPH_F_RIGID_VERTICES:
		h->rverts = (float4 *)dev_ptr; h->n_rverts = n; break;
	case SPH_F_FLUID_GID:
		if (n < ncap) return sph_fail(h, SPH_EINVAL, "sph_bind: fluid gid needs %zu int32, got %zu", ncap, n);
		h->gid = (int *)dev_ptr; h->n_gid = n; break;
	default:
		return sph_fail(h, SPH_EINVAL, "sph_bind: field %d is not bindable", field);
	}
	return SPH_OK;
}

static int require_state(SphHandle *h) {
	if (!h) return SPH_EINVAL;
	size_t ncap = (size_t)h->cfg.n_fluid + (size_t)(h->cfg.n_ghost_capacity > 0 ? h->cfg.n_ghost_capacity : 0);
	if (ncap > 0 && (!h->pos || !h->vel)) return sph_fail(h, SPH_ENOTBOUND, "fluid pos/vel are not bound (sph_bind)");
	if (h->c.Nb > 0 && h->c.boundary_handle == 1 && !h->boundary_ready)
		return sph_fail(h, SPH_ESTATE, "boundary particles are not initialised (sph_init_boundary)");
	return SPH_OK;
}

static int check_launch(SphHandle *h, const char *what) {
	if (h->async_error) { int rc = h->async_error; h->async_error = 0; return rc; } // message already recorded by sph_fail
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return sph_fail(h, SPH_ECUDA, "%s: %s", what, cudaGetErrorString(e));
	return SPH_OK;
}

extern "C" int sph_init_boundary(SphHandle *h, void *stream) {
	if (!h) return SPH_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	if (h->c.Nb > 0) {
		if (!h->bpos) return sph_fail(h, SPH_ENOTBOUND, "boundary pos is not bound");
		sphg_build(h, h->bg, h->bpos, h->c.Nb, st);   // PS:322-335
		sphg_gather_boundary(h, st);
		if (h->cfg.strict) sph_strict::boundary_volume(h, st); // PS:309-320
		else sph_fast::boundary_volume(h, st);
	} else {
		SPH_CUDA_CHECK(h, cudaMemsetAsync(h->bg.cell_start, 0, sizeof(int) * ((size_t)h->c.G + 1), st));
	}
	h->boundary_ready = true;
	return check_launch(h, "sph_init_boundary");
}

static int require_rigid(SphHandle *h) {
	if (!h) return SPH_EINVAL;
	if (h->c.Nr <= 0) return sph_fail(h, SPH_ESTATE, "no rigid body in this handle");
	if (!h->rpos || !h->rvel || !h->rforce) return sph_fail(h, SPH_ENOTBOUND, "rigid pos/vel/force are not bound");
	return SPH_OK;
}

extern "C" int sph_init_rigid(SphHandle *h, void *stream) {
	int rc = require_rigid(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	// PS:249-295: needs the rigid particles in the grid (only when active, SURVEY B-R2)
	sphg_build(h, h->rg, h->rpos, h->c.Nr, st);
	sphg_gather_rigid(h, st);
	if (h->cfg.strict) sph_strict::rigid_init(h, st); else sph_fast::rigid_init(h, st);
	h->rigid_ready = true;
	return check_launch(h, "sph_init_rigid");
}

extern "C" int sph_rigid_step(SphHandle *h, void *stream) {
	int rc = require_rigid(h);
	if (rc != SPH_OK) return rc;
	if (!h->rigid_ready) return sph_fail(h, SPH_ESTATE, "sph_rigid_step: call sph_init_rigid first");
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	if (h->cfg.strict) sph_strict::rigid_step(h, st); else sph_fast::rigid_step(h, st);
	return check_launch(h, "sph_rigid_step");
}

extern "C" int sph_rigid_state(SphHandle *h, SphRigidInfo *out) {
	if (!h || !out) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	SphRigidState s;
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize());
	SPH_CUDA_CHECK(h, cudaMemcpy(&s, h->rstate, sizeof(s), cudaMemcpyDeviceToHost));
	memset(out, 0, sizeof(*out));
	memcpy(out->centroid, s.centroid, sizeof(s.centroid));
	memcpy(out->inertia, s.inertia, sizeof(s.inertia));
	memcpy(out->inertia_inv, s.inertia_inv, sizeof(s.inertia_inv));
	memcpy(out->vel, s.vel, sizeof(s.vel));
	memcpy(out->omega, s.omega, sizeof(s.omega));
	memcpy(out->alpha, s.alpha, sizeof(s.alpha));
	memcpy(out->acc, s.acc, sizeof(s.acc));
	memcpy(out->attitude, s.attitude, sizeof(s.attitude));
	memcpy(out->force_sum, s.force_sum, sizeof(s.force_sum));
	memcpy(out->torque, s.torque, sizeof(s.torque));
	out->mass = s.mass;
	out->delta_time = s.rs_dt;
	out->collision_cnt = s.collision_cnt;
	out->simulate_cnt = s.simulate_cnt;
	out->max_surface_vel = s.max_surface_vel;
	return SPH_OK;
}

extern "C" int sph_rigid_set_state(SphHandle *h, const SphRigidInfo *in) {
	if (!h || !in) return SPH_EINVAL;
	if (!h->rstate || h->c.Nr <= 0) return sph_fail(h, SPH_ESTATE, "sph_rigid_set_state: the handle has no rigid body");
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	SphRigidState s;
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize());
	SPH_CUDA_CHECK(h, cudaMemcpy(&s, h->rstate, sizeof(s), cudaMemcpyDeviceToHost));
	memcpy(s.centroid, in->centroid, sizeof(s.centroid));
	memcpy(s.inertia, in->inertia, sizeof(s.inertia));
	memcpy(s.inertia_inv, in->inertia_inv, sizeof(s.inertia_inv));
	memcpy(s.vel, in->vel, sizeof(s.vel));
	memcpy(s.omega, in->omega, sizeof(s.omega));
	memcpy(s.alpha, in->alpha, sizeof(s.alpha));
	memcpy(s.acc, in->acc, sizeof(s.acc));
	memcpy(s.attitude, in->attitude, sizeof(s.attitude));
	memcpy(s.force_sum, in->force_sum, sizeof(s.force_sum));
	memcpy(s.torque, in->torque, sizeof(s.torque));
	s.mass = in->mass;
	s.rs_dt = in->delta_time;
	s.collision_cnt = in->collision_cnt;
	s.simulate_cnt = in->simulate_cnt;
	s.max_surface_vel = in->max_surface_vel;
	SPH_CUDA_CHECK(h, cudaMemcpy(h->rstate, &s, sizeof(s), cudaMemcpyHostToDevice));
	return SPH_OK;
}

// Deferred velocity upload (e2e path).  sph_upload_state_xyz puts the velocities on a copy stream; the step that follows
// builds its grid and neighbour lists from the positions while they travel (0.45 ms of PCIe at 10^6 particles behind
// 0.51 ms of kernels) and joins here, right before the first sweep that reads a velocity.
void sph_finish_deferred_vel(SphHandle *h, cudaStream_t st) {
	if (h->vel_in_flight) {
		cudaStreamWaitEvent(st, h->ev_vel_ready, 0);
		h->vel_in_flight = false;
	}
	if (h->vel_gather_pending) {
		sphg_gather_vel(h, st);
		h->vel_gather_pending = false;
	}
}

static int base_step(SphHandle *h, cudaStream_t st) {
	// SB:136-143: simulate_cnt += 1 ; reset_grid ; update_grid ; reset()
	h->simulate_cnt += 1;
	h->L.n_fluid = h->c.N; // slabs: owned + ghost particles of this step
	sph_prof_begin(h, KC_GRID, st);
	sphg_build(h, h->fg, h->pos, h->c.N, st, h->comm ? h->gid : nullptr);
	if (h->vel_in_flight) { // the velocities are still on the copy stream: positions and vel.w now, the rest later
		sphg_gather_fluid_pos(h, st);
		h->vel_gather_pending = true;
	} else {
		sphg_gather_fluid(h, st);
	}
	mg_after_grid(h, st); // multi-GPU: sorted slots of the sent / received particles of this step
	if (h->c.Nr > 0 && h->c.active_rigid) { // PS:385-386, 399-407
		sphg_build(h, h->rg, h->rpos, h->c.Nr, st);
		sphg_gather_rigid(h, st);
		mg_rigid_quirk_update(h, 0, st); // slabs: positions of the fluid particles the count quirk looks at
	}
	sph_prof_end(h, st);
	h->grid_valid = true;
	h->lists_valid = false;
	h->lists_fresh = false;
	return SPH_OK;
}

extern "C" int sph_phase(SphHandle *h, int phase, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	if (h->c.N <= 0 && !h->comm) { // empty fluid block: nothing to sort or sweep
		if (phase == SPH_PH_BUILD_GRID) h->simulate_cnt += 1;
		return SPH_OK;
	}
	if (phase == SPH_PH_BUILD_GRID) {
		if (h->comm) {
			rc = mg_begin_step(h, st);
			if (rc != SPH_OK) return rc;
		}
		base_step(h, st);
		return check_launch(h, "sph_phase(build_grid)");
	}
	if (!h->grid_valid) return sph_fail(h, SPH_ESTATE, "sph_phase: grid not built (SPH_PH_BUILD_GRID first)");
	// a deferred velocity upload is joined by the first phase that reads a velocity: the list-building first phases
	// do it themselves after the build (first_phase_lists), every other phase here
	if (phase != SPH_PH_BUILD_LISTS && phase != SPH_PH_DF_INITIALIZE && phase != SPH_PH_WC_PRESSURE && phase != SPH_PH_PC_EXT_FORCE &&
	    phase != SPH_PH_II_PREDICT_ADVECTION)
		sph_finish_deferred_vel(h, st);
	bool strict = h->cfg.strict != 0;
	if (phase == SPH_PH_WRITEBACK) {
		sphg_writeback(h, h->a4[A4_POS], h->a4[A4_VEL], st);
		return check_launch(h, "sph_phase(writeback)");
	}
	if (phase != SPH_PH_DF_INITIALIZE && phase != SPH_PH_WC_PRESSURE && phase != SPH_PH_PC_EXT_FORCE && phase != SPH_PH_BUILD_LISTS &&
	    phase != SPH_PH_II_PREDICT_ADVECTION && phase != SPH_PH_PBF_PREDICT && phase != SPH_PH_PBF_LAMBDA && !h->lists_valid)
		return sph_fail(h, SPH_ESTATE, "sph_phase: neighbour lists not built (run the solver's first phase)");
	if (phase == SPH_PH_BUILD_LISTS) {
		if (h->c.solver == SPH_SOLVER_PBF) return sph_fail(h, SPH_EINVAL, "sph_phase: PBF builds its lists on the predicted positions (SPH_PH_PBF_LAMBDA)");
		h->lists_fresh = false;
		if (strict) sph_strict::first_phase_lists(h, st); else sph_fast::first_phase_lists(h, st);
		h->lists_fresh = true;
	} else if ((phase >= SPH_PH_DF_INITIALIZE && phase <= SPH_PH_DF_POSITION) || (phase >= SPH_PH_DF_WARM_START && phase <= SPH_PH_DF_DEN_VEL)) {
		if (h->c.solver != SPH_SOLVER_DFSPH) return sph_fail(h, SPH_EINVAL, "sph_phase: handle is not a DFSPH solver");
		if (strict) sph_strict::df_phase(h, phase, st); else sph_fast::df_phase(h, phase, st);
	} else if ((phase >= SPH_PH_WC_PRESSURE && phase <= SPH_PH_WC_KINEMATIC) || phase == SPH_PH_WC_EOS || phase == SPH_PH_WC_FORCE) {
		if (h->c.solver != SPH_SOLVER_WCSPH) return sph_fail(h, SPH_EINVAL, "sph_phase: handle is not a WCSPH solver");
		if (strict) sph_strict::wc_phase(h, phase, st); else sph_fast::wc_phase(h, phase, st);
	} else if ((phase >= SPH_PH_PC_EXT_FORCE && phase <= SPH_PH_PC_INTEGRATION) || (phase >= SPH_PH_PC_PREDICT && phase <= SPH_PH_PC_RHO)) {
		if (h->c.solver != SPH_SOLVER_PCISPH) return sph_fail(h, SPH_EINVAL, "sph_phase: handle is not a PCISPH solver");
		if (strict) sph_strict::pc_phase(h, phase, st); else sph_fast::pc_phase(h, phase, st);
	} else if ((phase >= SPH_PH_II_PREDICT_ADVECTION && phase <= SPH_PH_II_INTEGRATION) || (phase >= SPH_PH_II_ADVECT && phase <= SPH_PH_II_UPDATE)) {
		if (h->c.solver != SPH_SOLVER_IISPH) return sph_fail(h, SPH_EINVAL, "sph_phase: handle is not an IISPH solver");
		if (strict) sph_strict::ii_phase(h, phase, st); else sph_fast::ii_phase(h, phase, st);
	} else if (phase >= SPH_PH_PBF_PREDICT && phase <= SPH_PH_PBF_UPDATE_POS) {
		if (h->c.solver != SPH_SOLVER_PBF) return sph_fail(h, SPH_EINVAL, "sph_phase: handle is not a PBF solver");
		if (strict) sph_strict::pbf_phase(h, phase, st); else sph_fast::pbf_phase(h, phase, st);
	} else {
		return sph_fail(h, SPH_EINVAL, "sph_phase: unknown phase %d", phase);
	}
	return check_launch(h, "sph_phase");
}

// one solver step on `st`: prologue (SB:136-143) + the solver's phases
static void enqueue_step(SphHandle *h, cudaStream_t st, bool strict) {
	base_step(h, st);
	if (h->c.solver == SPH_SOLVER_PBF) sph_finish_deferred_vel(h, st); // its first phase integrates the velocities
	switch (h->c.solver) {
	case SPH_SOLVER_DFSPH:
		if (strict) sph_strict::df_step(h, st); else sph_fast::df_step(h, st);
		break;
	case SPH_SOLVER_WCSPH:
		for (int p = SPH_PH_WC_PRESSURE; p <= SPH_PH_WC_KINEMATIC; ++p)
			if (strict) sph_strict::wc_phase(h, p, st); else sph_fast::wc_phase(h, p, st);
		break;
	case SPH_SOLVER_PCISPH:
		for (int p = SPH_PH_PC_EXT_FORCE; p <= SPH_PH_PC_INTEGRATION; ++p)
			if (strict) sph_strict::pc_phase(h, p, st); else sph_fast::pc_phase(h, p, st);
		break;
	case SPH_SOLVER_IISPH:
		for (int p = SPH_PH_II_PREDICT_ADVECTION; p <= SPH_PH_II_INTEGRATION; ++p)
			if (strict) sph_strict::ii_phase(h, p, st); else sph_fast::ii_phase(h, p, st);
		break;
	case SPH_SOLVER_PBF:
		for (int p = SPH_PH_PBF_PREDICT; p <= SPH_PH_PBF_UPDATE_POS; ++p)
			if (strict) sph_strict::pbf_phase(h, p, st); else sph_fast::pbf_phase(h, p, st);
		break;
	default: break;
	}
}

// CUDA graph of a whole step for the solvers whose step never looks at the host (WCSPH, PBF: no convergence loop).
// Their step is ~13 launches of a few microseconds each on the small scenes the reference ships (BASELINE configs[0]:
// 29 k particles), i.e. launch-bound; the graph is captured once on an internal stream (the caller's may be the
// legacy default stream, which cannot be captured) and replayed into the caller's stream.  Not used on slabs (the
// exchange epochs are launch arguments), while per-launch profiling is on, or while a deferred upload is pending.
static bool step_graphable(const SphHandle *h) {
	static const bool off = getenv("SPH_NO_GRAPH") != nullptr;
	if (off || h->comm || (h->prof && h->prof->on) || h->vel_in_flight || h->vel_gather_pending) return false;
	return h->c.solver == SPH_SOLVER_WCSPH || h->c.solver == SPH_SOLVER_PBF;
}
static int step_graph_launch(SphHandle *h, cudaStream_t st, bool strict) {
	if (!h->step_graph) {
		if (!h->graph_stream) SPH_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
		SPH_CUDA_CHECK(h, cudaStreamSynchronize(st)); // the capture stream is not ordered behind the caller's stream
		int launches0 = h->launches, sim0 = h->simulate_cnt;
		cudaGraph_t g = nullptr;
		SPH_CUDA_CHECK(h, cudaStreamBeginCapture(h->graph_stream, cudaStreamCaptureModeThreadLocal));
		enqueue_step(h, h->graph_stream, strict);
		cudaError_t e = cudaStreamEndCapture(h->graph_stream, &g);
		h->step_graph_launches = h->launches - launches0;
		h->launches = launches0;
		h->simulate_cnt = sim0;
		if (e != cudaSuccess || !g) return sph_fail(h, SPH_ECUDA, "sph_step: graph capture failed: %s", cudaGetErrorString(e));
		e = cudaGraphInstantiate(&h->step_graph, g, 0);
		cudaGraphDestroy(g);
		if (e != cudaSuccess) { h->step_graph = nullptr; return sph_fail(h, SPH_ECUDA, "sph_step: cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
	}
	SPH_CUDA_CHECK(h, cudaGraphLaunch(h->step_graph, st));
	h->launches += h->step_graph_launches;
	h->simulate_cnt += 1;
	h->grid_valid = true; // what enqueue_step leaves behind on the host side
	h->lists_fresh = false;
	return SPH_OK;
}

extern "C" int sph_step(SphHandle *h, int n_substeps, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	bool strict = h->cfg.strict != 0;
	if (h->c.N <= 0 && !h->comm) { // empty fluid block: the counter advances, nothing to sort or sweep
		h->simulate_cnt += n_substeps;
		return SPH_OK;
	}
	for (int k = 0; k < n_substeps; ++k) {
		if (h->comm) { // multi-GPU slab: migration + ghost particles first (positions moved last step)
			sph_finish_deferred_vel(h, st);
			rc = mg_begin_step(h, st);
			if (rc != SPH_OK) return rc;
		}
		if (step_graphable(h)) {
			rc = step_graph_launch(h, st, strict);
			if (rc != SPH_OK) return rc;
		} else {
			enqueue_step(h, st, strict);
		}
		h->grid_valid = false; // positions moved: sorted buffers describe the previous state
		h->lists_valid = false;
	}
	return check_launch(h, "sph_step");
}

extern "C" int sph_pcisph_precompute(SphHandle *h, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	if (h->comm) { // slabs: the ghost layer must be in place for the neighbour counts of the edge columns
		rc = mg_begin_step(h, st);
		if (rc != SPH_OK) return rc;
	}
	sphg_build(h, h->fg, h->pos, h->c.N, st, h->comm ? h->gid : nullptr); // PC:29-31
	sphg_gather_fluid(h, st);
	mg_after_grid(h, st);
	h->grid_valid = true;
	if (h->cfg.strict) sph_strict::pc_precompute(h, st); else sph_fast::pc_precompute(h, st);
	return check_launch(h, "sph_pcisph_precompute");
}

__global__ void k_set_pc_delta(SphCtl *ctl, float delta, int index) {
	ctl->pc_delta = delta;
	ctl->pc_max_index = index;
}

extern "C" int sph_pcisph_set_delta(SphHandle *h, float delta, int particle_index, void *stream) {
	if (!h) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	k_set_pc_delta<<<1, 1, 0, (cudaStream_t)stream>>>(h->ctl, delta, particle_index);
	h->launches++;
	return check_launch(h, "sph_pcisph_set_delta");
}

extern "C" int sph_pcisph_delta(SphHandle *h, int particle_index, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	if (!h->lists_valid) return sph_fail(h, SPH_ESTATE, "sph_pcisph_delta: call sph_pcisph_precompute first");
	if (particle_index < 0 || particle_index >= h->c.N_owned)
		return sph_fail(h, SPH_EINVAL, "sph_pcisph_delta: particle index %d out of range", particle_index);
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	if (h->cfg.strict) sph_strict::pc_set_delta(h, particle_index, st); else sph_fast::pc_set_delta(h, particle_index, st);
	return check_launch(h, "sph_pcisph_delta");
}

__global__ void k_set_dt(SphCtl *ctl, float dt) {
	ctl->dt = dt;
	ctl->dt2 = dt * dt;
}

extern "C" int sph_set_delta_time(SphHandle *h, float dt, void *stream) {
	if (!h) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	k_set_dt<<<1, 1, 0, (cudaStream_t)stream>>>(h->ctl, dt);
	h->launches++;
	return check_launch(h, "sph_set_delta_time");
}

extern "C" int sph_fetch(SphHandle *h, int field, void *dev_out, size_t n, void *stream) {
	if (!h || !dev_out) return SPH_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	const SphConsts &c = h->c;
	size_t N = (size_t)c.N;
	auto f1 = [&](const float *src) -> int {
		if (n < N) return sph_fail(h, SPH_EINVAL, "sph_fetch: output too small (%zu < %zu)", n, N);
		sphg_unsort_f1(h, h->fg, src, (float *)dev_out, c.N, st);
		return SPH_OK;
	};
	auto f4 = [&](const float4 *src) -> int {
		if (n < N) return sph_fail(h, SPH_EINVAL, "sph_fetch: output too small (%zu < %zu)", n, N);
		sphg_unsort_f4(h, h->fg, src, (float4 *)dev_out, c.N, st);
		return SPH_OK;
	};
	auto i1 = [&](const int *src) -> int {
		if (n < N) return sph_fail(h, SPH_EINVAL, "sph_fetch: output too small (%zu < %zu)", n, N);
		sphg_unsort_i1(h, h->fg, src, (int *)dev_out, c.N, st);
		return SPH_OK;
	};
	auto raw = [&](const void *src, size_t cnt, size_t esz) -> int {
		if (n < cnt) return sph_fail(h, SPH_EINVAL, "sph_fetch: output too small (%zu < %zu)", n, cnt);
		SPH_CUDA_CHECK(h, cudaMemcpyAsync(dev_out, src, cnt * esz, cudaMemcpyDeviceToDevice, st));
		return SPH_OK;
	};
	int rc;
	switch (field) {
	case SPH_F_RHO: rc = f1(h->a1[A1_RHO]); break;
	case SPH_F_ALPHA: rc = f1(h->a1[A1_ALPHA]); break;
	case SPH_F_RHO_DERIVATIVE: rc = f1(h->a1[A1_DRHO]); break;
	case SPH_F_RHO_ADV: rc = f1(h->a1[A1_RHOADV]); break;
	case SPH_F_PRESSURE: rc = f1(h->a1[A1_P]); break;
	case SPH_F_SCALAR_A: rc = f1(h->a1[A1_SA]); break;
	case SPH_F_SCALAR_B: rc = f1(h->a1[A1_SB]); break;
	case SPH_F_SCALAR_C: rc = f1(h->a1[A1_SC]); break;
	case SPH_F_VEL_ADV: rc = f4(h->a4[A4_VADV]); break;
	case SPH_F_FORCE_A: rc = f4(h->a4[A4_FA]); break;
	case SPH_F_FORCE_B: rc = f4(h->a4[A4_FB]); break;
	case SPH_F_VEC_A: rc = f4(h->a4[A4_FC]); break;
	case SPH_F_VEC_B: rc = f4(h->a4[A4_FD]); break;
	case SPH_F_VEC_C: rc = f4(h->a4[A4_T2]); break;
	case SPH_F_PAYLOAD_1: rc = f4(h->a4[A4_T1]); break;
	case SPH_F_PAYLOAD_3: rc = f4(h->a4[A4_T3]); break;
	case SPH_F_POS_RHO: rc = f4(h->a4[A4_PR]); break;
	case SPH_F_FLUID_VEL: rc = f4(h->a4[A4_VEL]); break;   // in-step (sorted) velocity, original order
	case SPH_F_FLUID_POS: rc = f4(h->a4[A4_POS]); break;
	case SPH_F_CELL1D: rc = raw(h->fg.cell_of, N, sizeof(int)); break;
	case SPH_F_NEIGHBOR_COUNT: rc = i1(h->nbr_count); break;
	case SPH_F_BOUNDARY_NEIGHBOR_COUNT: rc = i1(h->L.bcount); break;
	case SPH_F_CELL_START: rc = raw(h->fg.cell_start, (size_t)c.G + 1, sizeof(int)); break;
	case SPH_F_SORTED_INDEX: rc = raw(h->fg.sorted_id, N, sizeof(int)); break;
	case SPH_F_BOUNDARY_CELL_START: rc = raw(h->bg.cell_start, (size_t)c.G + 1, sizeof(int)); break;
	case SPH_F_BOUNDARY_SORTED_INDEX: rc = raw(h->bg.sorted_id, (size_t)c.Nb, sizeof(int)); break;
	default: return sph_fail(h, SPH_EINVAL, "sph_fetch: field %d cannot be fetched", field);
	}
	if (rc != SPH_OK) return rc;
	return check_launch(h, "sph_fetch");
}

extern "C" int sph_visualize(SphHandle *h, int what, void *dev_rgb, int stride_floats, size_t n, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	if (!dev_rgb || stride_floats < 3) return sph_fail(h, SPH_EINVAL, "sph_visualize: rgb buffer with a stride of >= 3 floats expected");
	if (what != SPH_VIS_RHO && what != SPH_VIS_NEIGHBOUR) return sph_fail(h, SPH_EINVAL, "sph_visualize: unknown quantity %d", what);
	if (n < (size_t)h->c.N_owned) return sph_fail(h, SPH_EINVAL, "sph_visualize: output too small (%zu < %d)", n, h->c.N_owned);
	if (h->simulate_cnt <= 0 && !h->lists_valid)
		return sph_fail(h, SPH_ESTATE, "sph_visualize: no density / neighbour counts yet (run a solver step first)");
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sphg_visualize(h, what, (float *)dev_rgb, stride_floats, (cudaStream_t)stream);
	return check_launch(h, "sph_visualize");
}

extern "C" int sph_upload_state(SphHandle *h, const float *host_pos4, const float *host_vel4, void *stream) {
	if (!h || !h->pos || !h->vel) return h ? sph_fail(h, SPH_ENOTBOUND, "sph_upload_state: state not bound") : SPH_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	size_t bytes = sizeof(float4) * (size_t)h->c.N_owned;
	if (host_pos4) SPH_CUDA_CHECK(h, cudaMemcpyAsync(h->pos, host_pos4, bytes, cudaMemcpyHostToDevice, st));
	if (host_vel4) SPH_CUDA_CHECK(h, cudaMemcpyAsync(h->vel, host_vel4, bytes, cudaMemcpyHostToDevice, st));
	return SPH_OK;
}

extern "C" int sph_download_state(SphHandle *h, float *host_pos4, float *host_vel4, void *stream) {
	if (!h || !h->pos || !h->vel) return h ? sph_fail(h, SPH_ENOTBOUND, "sph_download_state: state not bound") : SPH_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	size_t bytes = sizeof(float4) * (size_t)h->c.N_owned;
	if (host_pos4) SPH_CUDA_CHECK(h, cudaMemcpyAsync(host_pos4, h->pos, bytes, cudaMemcpyDeviceToHost, st));
	if (host_vel4) SPH_CUDA_CHECK(h, cudaMemcpyAsync(host_vel4, h->vel, bytes, cudaMemcpyDeviceToHost, st));
	SPH_CUDA_CHECK(h, cudaStreamSynchronize(st));
	return SPH_OK;
}

static int xyz_stage(SphHandle *h) {
	if (h->xyz_stage) return SPH_OK;
	size_t n = (size_t)h->cfg.n_fluid;
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&h->xyz_stage, sizeof(float) * 6 * (n ? n : 1)));
	return SPH_OK;
}

extern "C" int sph_upload_state_xyz(SphHandle *h, const float *host_pos3, const float *host_vel3, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	if ((rc = xyz_stage(h)) != SPH_OK) return rc;
	size_t n = (size_t)h->c.N_owned, bytes = sizeof(float) * 3 * n;
	float *p3 = h->xyz_stage, *v3 = h->xyz_stage + 3 * (size_t)h->cfg.n_fluid;
	sph_finish_deferred_vel(h, st); // an earlier deferred upload must have landed before its staging buffer is reused
	// One GPU, positions and velocities together: the positions go first on the caller's stream -- the grid and the
	// neighbour lists of the next step need nothing else -- and the velocities follow on a copy stream; the step joins
	// them right before its first sweep that reads a velocity (sph_finish_deferred_vel).  SPH_E2E_NO_DEFER: A/B knob.
	static const bool no_defer = getenv("SPH_E2E_NO_DEFER") != nullptr;
	bool defer = host_pos3 && host_vel3 && !h->comm && !no_defer;
	if (defer && !h->copy_stream) {
		SPH_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
		SPH_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_vel_ready, cudaEventDisableTiming));
		SPH_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_mark, cudaEventDisableTiming));
	}
	if (host_pos3) SPH_CUDA_CHECK(h, cudaMemcpyAsync(p3, host_pos3, bytes, cudaMemcpyHostToDevice, st));
	if (defer) {
		sphg_unpack_xyz(h, p3, nullptr, (int)n, st);
		SPH_CUDA_CHECK(h, cudaEventRecord(h->ev_mark, st));              // everything the caller enqueued before this call ...
		SPH_CUDA_CHECK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_mark, 0)); // ... has finished with h->vel and the staging buffer
		SPH_CUDA_CHECK(h, cudaMemcpyAsync(v3, host_vel3, bytes, cudaMemcpyHostToDevice, h->copy_stream));
		sphg_unpack_xyz(h, nullptr, v3, (int)n, h->copy_stream);
		SPH_CUDA_CHECK(h, cudaEventRecord(h->ev_vel_ready, h->copy_stream));
		h->vel_in_flight = true;
	} else {
		if (host_vel3) SPH_CUDA_CHECK(h, cudaMemcpyAsync(v3, host_vel3, bytes, cudaMemcpyHostToDevice, st));
		sphg_unpack_xyz(h, host_pos3 ? p3 : nullptr, host_vel3 ? v3 : nullptr, (int)n, st);
	}
	return check_launch(h, "sph_upload_state_xyz");
}

extern "C" int sph_download_state_xyz(SphHandle *h, float *host_pos3, float *host_vel3, void *stream) {
	int rc = require_state(h);
	if (rc != SPH_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	sph_finish_deferred_vel(h, (cudaStream_t)stream);
	if ((rc = xyz_stage(h)) != SPH_OK) return rc;
	size_t n = (size_t)h->c.N_owned, bytes = sizeof(float) * 3 * n;
	float *p3 = h->xyz_stage, *v3 = h->xyz_stage + 3 * (size_t)h->cfg.n_fluid;
	sphg_pack_xyz(h, host_pos3 ? p3 : nullptr, host_vel3 ? v3 : nullptr, (int)n, st);
	if (host_pos3) SPH_CUDA_CHECK(h, cudaMemcpyAsync(host_pos3, p3, bytes, cudaMemcpyDeviceToHost, st));
	if (host_vel3) SPH_CUDA_CHECK(h, cudaMemcpyAsync(host_vel3, v3, bytes, cudaMemcpyDeviceToHost, st));
	SPH_CUDA_CHECK(h, cudaStreamSynchronize(st));
	return SPH_OK;
}

extern "C" int sph_profile_begin(SphHandle *h) {
	if (!h) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	if (!h->prof) {
		h->prof = new (std::nothrow) SphProf();
		if (!h->prof) return sph_fail(h, SPH_ENOMEM, "sph_profile_begin: out of memory");
		memset(h->prof, 0, sizeof(SphProf));
	}
	if (!h->prof->created) {
		for (int i = 0; i < SPH_PROF_CAP; ++i) {
			SPH_CUDA_CHECK(h, cudaEventCreate(&h->prof->e0[i]));
			SPH_CUDA_CHECK(h, cudaEventCreate(&h->prof->e1[i]));
		}
		h->prof->created = true;
	}
	h->prof->n = 0;
	h->prof->on = true;
	return SPH_OK;
}

extern "C" int sph_profile_end(SphHandle *h, float *ms_by_class, int32_t *launches_by_class, int n_classes) {
	if (!h || !ms_by_class || !launches_by_class) return SPH_EINVAL;
	if (!h->prof || !h->prof->on) return sph_fail(h, SPH_ESTATE, "sph_profile_end: profiling is not active");
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize());
	for (int k = 0; k < n_classes; ++k) { ms_by_class[k] = 0.0f; launches_by_class[k] = 0; }
	for (int i = 0; i < h->prof->n; ++i) {
		float ms = 0.0f;
		SPH_CUDA_CHECK(h, cudaEventElapsedTime(&ms, h->prof->e0[i], h->prof->e1[i]));
		int k = h->prof->kid[i];
		if (k >= 0 && k < n_classes) { ms_by_class[k] += ms; launches_by_class[k] += 1; }
	}
	h->prof->on = false;
	h->prof->n = 0;
	return SPH_OK;
}

extern "C" int sph_read_stats(SphHandle *h, SphStats *out) {
	if (!h || !out) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	int sim = h->simulate_cnt;
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize());
	SPH_CUDA_CHECK(h, cudaMemcpy(h->ctl_host, h->ctl, sizeof(SphCtl), cudaMemcpyDeviceToHost));
	const SphCtl &k = *h->ctl_host;
	memset(out, 0, sizeof(*out));
	out->delta_time = k.dt;
	out->ps_delta_time = k.ps_dt;
	out->simulate_cnt = sim;
	out->error_flags = k.error_flags;
	out->div_iters = k.div_iters;
	out->div_first_err = k.div_first;
	out->div_err = k.div_err;
	out->den_iters = k.den_iters;
	out->den_err = (float)((double)k.den_avg - 1000.0);
	out->pc_iters = k.pc_iters;
	out->pc_err = k.pc_err;
	out->pc_delta = k.pc_delta;
	out->pc_max_index = k.pc_max_index;
	out->ii_iters = k.ii_iters;
	out->ii_residual = k.ii_residual;
	out->max_neighbors_seen = k.max_nbr;
	out->max_boundary_neighbors_seen = k.max_bnbr;
	out->kernel_launches = h->launches;
	out->div_active = k.div_active;
	out->den_active = k.den_active;
	out->loop_active = h->c.solver == SPH_SOLVER_PCISPH ? k.pc_active : h->c.solver == SPH_SOLVER_IISPH ? k.ii_active : 0;
	return SPH_OK;
}

#if SPH_DEBUG_BOUNDS
// bounds-checked build only (not part of the ABI): overwrite entry `entry` of sorted particle `slot`'s fluid list, so
// that tests/test_gpu_bounds.py can show the index checks are live
extern "C" int sph_debug_poke_list(SphHandle *h, int slot, int entry, unsigned int value) {
	if (!h || slot < 0 || slot >= h->c.N || entry < 0 || entry >= h->c.kstride) return SPH_EINVAL;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize());
	SPH_CUDA_CHECK(h, cudaMemcpy(h->L.flist + sph_list_word(slot, h->c.kstride, entry), &value, sizeof(value), cudaMemcpyHostToDevice));
	return SPH_OK;
}
#endif

// Single-sweep parity tests: put handle `dst` into the in-step state of handle `src` (same scene, grids built
// from the same positions, so the sorted order is identical): every sorted per-particle work array, the
// neighbour counts, the control block and the rigid-body state.  The neighbour lists are NOT copied -- the
// strict and the fast kernels keep them in different orders; each handle builds its own.
extern "C" int sph_copy_work_state(SphHandle *dst, SphHandle *src, void *stream) {
	if (!dst || !src) return SPH_EINVAL;
	if (dst->c.N != src->c.N || dst->c.solver != src->c.solver || dst->c.Nr != src->c.Nr || dst->device != src->device)
		return sph_fail(dst, SPH_EINVAL, "sph_copy_work_state: the two handles do not describe the same scene on one device");
	if (!dst->grid_valid || !src->grid_valid) return sph_fail(dst, SPH_ESTATE, "sph_copy_work_state: build both grids first");
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(dst, cudaSetDevice(dst->device));
	sph_finish_deferred_vel(dst, st);
	sph_finish_deferred_vel(src, st);
	size_t n = (size_t)src->c.N;
	for (int k = 0; k < A4_COUNT; ++k)
		SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->a4[k], src->a4[k], sizeof(float4) * n, cudaMemcpyDeviceToDevice, st));
	for (int k = 0; k < A1_COUNT; ++k)
		SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->a1[k], src->a1[k], sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
	SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->nbr_count, src->nbr_count, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
	SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->ctl, src->ctl, sizeof(SphCtl), cudaMemcpyDeviceToDevice, st));
	SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->rstate, src->rstate, sizeof(SphRigidState), cudaMemcpyDeviceToDevice, st));
	if (src->c.Nr > 0 && src->rforce && dst->rforce)
		SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->rforce, src->rforce, sizeof(float4) * (size_t)src->c.Nr, cudaMemcpyDeviceToDevice, st));
	dst->den_piece = src->den_piece;
	return SPH_OK;
}

// ---- multi-GPU slab support ------------------------------------------------------------------------
extern "C" int sph_set_counts(SphHandle *h, int n_owned, int n_ghost) {
	if (!h) return SPH_EINVAL;
	size_t ncap = (size_t)h->cfg.n_fluid + (size_t)(h->cfg.n_ghost_capacity > 0 ? h->cfg.n_ghost_capacity : 0);
	if (n_owned < 0 || n_ghost < 0 || (size_t)n_owned + (size_t)n_ghost > ncap)
		return sph_fail(h, SPH_EINVAL, "sph_set_counts: %d owned + %d ghost exceeds capacity %zu", n_owned, n_ghost, ncap);
	h->c.N_owned = n_owned;
	h->c.N = n_owned + n_ghost;
	return SPH_OK;
}
