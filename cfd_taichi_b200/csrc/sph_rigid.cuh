// sph_rigid.cuh -- rigid-body coupling (included by sph_sweeps.cu, both modes):
//   rigid_init      ParticleSystem.init_rigid_particles_data (PS:249-295)
//   rigid_lists     per-step rigid-centric fluid neighbour lists, ordered by original fluid index
//   rigid_force_df  the fluid->rigid force of DFSPH's density solve (DF:212) as a GATHER per rigid
//                   particle: no per-pair atomics, same accumulation order as the one-thread reference
//   rigid_step      rigid_solver.step (RS:216-234) in one single-block kernel: torque / force
//                   reductions, attitude, rotation, wall contact scan, impulse, move.
#pragma once

namespace SPH_NS {

__device__ __forceinline__ void st3(float *p, f3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// 3x3 helpers, row-major, Taichi's left-to-right accumulation
__device__ __forceinline__ f3 mat_mul_v(const float *M, f3 a) {
	return F3((M[0] * a.x + M[1] * a.y) + M[2] * a.z, (M[3] * a.x + M[4] * a.y) + M[5] * a.z,
	          (M[6] * a.x + M[7] * a.y) + M[8] * a.z);
}
__device__ inline void mat_mul(const float *A, const float *B, float *C) {
	float T[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) T[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
	for (int k = 0; k < 9; ++k) C[k] = T[k];
}
// Taichi's Matrix.inverse() for n = 3 (as restated in the oracle): 1 / determinant first, then the cofactor products
__device__ inline void mat_inverse(const float *m, float *inv) {
#define MI_E(x, y) m[3 * ((x) % 3) + ((y) % 3)]
	float det = (MI_E(0, 0) * (MI_E(1, 1) * MI_E(2, 2) - MI_E(2, 1) * MI_E(1, 2)) -
	             MI_E(1, 0) * (MI_E(0, 1) * MI_E(2, 2) - MI_E(2, 1) * MI_E(0, 2))) +
	            MI_E(2, 0) * (MI_E(0, 1) * MI_E(1, 2) - MI_E(1, 1) * MI_E(0, 2));
	float inv_det = 1.0f / det;
	float t[9];
#pragma unroll
	for (int i = 0; i < 3; ++i)
#pragma unroll
		for (int j = 0; j < 3; ++j)
			t[3 * j + i] = inv_det * (MI_E(i + 1, j + 1) * MI_E(i + 2, j + 2) - MI_E(i + 2, j + 1) * MI_E(i + 1, j + 2));
#undef MI_E
	for (int k = 0; k < 9; ++k) inv[k] = t[k];
}
// ti.math.rotation3d as restated in the oracle.  sin / cos are evaluated in fp64
// and rounded once so that they agree with a correctly rounded libm sinf / cosf.
__device__ inline void rotation3d(float ang_x, float ang_y, float ang_z, float *R) {
	float yaw = ang_z, pitch = ang_x, roll = ang_y;
	float ch = (float)cos((double)yaw), sh = (float)sin((double)yaw);
	float cp = (float)cos((double)pitch), sp = (float)sin((double)pitch);
	float cb = (float)cos((double)roll), sb = (float)sin((double)roll);
	R[0] = ch * cb + sh * sp * sb; R[1] = sb * cp;  R[2] = -sh * cb + ch * sp * sb;
	R[3] = -ch * sb + sh * sp * cb; R[4] = cb * cp; R[5] = sb * sh + ch * sp * cb;
	R[6] = sh * cp;                 R[7] = -sp;     R[8] = ch * cp;
}

// ---- init: Akinci volume of the rigid particles over their rigid neighbours (PS:252-259, 301-307) ----
__global__ void __launch_bounds__(SPH_BLOCK)
k_rigid_volume(SphConsts c, float4 *__restrict__ rspos, const int *__restrict__ rscell, const int *__restrict__ rstart,
               const int *__restrict__ rsorted_id, float4 *__restrict__ rpos_user) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.Nr) return;
	float4 pi = rspos[s];
	int cx, cy, cz;
	cell_xyz(rscell[s], c, cx, cy, cz);
	float volume = 0.0f;
	SPH_FOR_27(c, cx, cy, cz, c1) {
		int a = rstart[c1], b = rstart[c1 + 1];
		for (int e = a; e < b; ++e) {
			if (e == s) continue;
			Pair p = make_pair(pi, rspos[e]);
			if (culled(p, c)) continue;
			volume += cubic_w(p, c);
		}
	}
	float v = volume < 1e-6f ? 0.0f : 1.0f / volume; // PS:255-259
	rspos[s].w = v;
	rpos_user[rsorted_id[s]].w = v;
}

// PS:261-291 mass, centroid, inertia tensor and inverse; RS:156-162 total mass.  One thread, the
// reference's index order (one-time set-up).
__global__ void k_rigid_mass_props(SphConsts c, const float4 *__restrict__ rpos, float4 *__restrict__ rvel,
                                   float rigid_rho, SphRigidState *st) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	f3 cen = F3(0.0f, 0.0f, 0.0f);
	float sum_mass = 0.0f;
	for (int i = 0; i < c.Nr; ++i) {
		float4 p = rpos[i];
		float m = rigid_rho * p.w; // PS:263
		rvel[i].w = m;
		cen = cen + xyz(p) * m;    // PS:269
		sum_mass += m;
	}
	cen = cen / sum_mass;          // PS:271
	float Ixx = 0, Iyy = 0, Izz = 0, Ixy = 0, Ixz = 0, Iyz = 0;
	for (int i = 0; i < c.Nr; ++i) {
		float4 p4 = rpos[i];
		f3 p = xyz(p4) - cen;
		float m = rvel[i].w;
		Ixx += m * (p.y * p.y + p.z * p.z);
		Iyy += m * (p.x * p.x + p.z * p.z);
		Izz += m * (p.x * p.x + p.y * p.y);
		Ixy += (-m) * (p.x * p.y);
		Ixz += (-m) * (p.x * p.z);
		Iyz += (-m) * (p.z * p.y);
	}
	float I[9] = {Ixx, Ixy, Ixz, Ixy, Iyy, Iyz, Ixz, Iyz, Izz};
	for (int k = 0; k < 9; ++k) st->inertia[k] = I[k];
	mat_inverse(I, st->inertia_inv);
	st3(st->centroid, cen);
	st->mass = sum_mass;
}

void rigid_init(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	if (c.active_rigid) { // inactive rigid particles are not in the grid: zero volumes (SURVEY B-R2)
		k_rigid_volume<<<cdiv(c.Nr, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->rspos, h->rg.scell, h->rg.cell_start,
		                                                            h->rg.sorted_id, h->rpos);
		h->launches++;
	}
	k_rigid_mass_props<<<1, 32, 0, st>>>(c, h->rpos, h->rvel, (float)h->cfg.rigid_rho, h->rstate);
	h->launches++;
}

// ---- per-step rigid-centric lists: fluid neighbours of every rigid particle, ascending ORIGINAL fluid
// ---- index (the order in which the one-thread reference adds to rigid_particles[j].force) ------------
__global__ void __launch_bounds__(SPH_BLOCK)
k_rigid_lists(SphConsts c, const float4 *__restrict__ rspos, const int *__restrict__ rscell,
              const float4 *__restrict__ spos, const int *__restrict__ cstart, const int *__restrict__ sorted_id,
              uint32_t *__restrict__ rl_list, int *__restrict__ rl_count, int cap, SphCtl *ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.Nr) return;
	float4 pr = rspos[s];
	int cx, cy, cz;
	cell_xyz(rscell[s], c, cx, cy, cz);
	uint32_t *lp = rl_list + sph_list_base(s, cap);
	int n = 0;
	SPH_FOR_27(c, cx, cy, cz, c1) {
		int a = cstart[c1], b = cstart[c1 + 1];
		for (int e = a; e < b; ++e) {
			if (sorted_id[e] >= c.N_owned) continue; // slabs: ghost copies push on their owner's rank
			Pair p = make_pair(spos[e], pr);
			if (culled(p, c)) continue;
			if (n < cap) {
				// insertion by original index
				int key = sorted_id[e];
				int k = n - 1;
				while (k >= 0 && sorted_id[lp[(size_t)k * 32]] > key) { lp[(size_t)(k + 1) * 32] = lp[(size_t)k * 32]; --k; }
				lp[(size_t)(k + 1) * 32] = (uint32_t)e;
			}
			n++;
		}
	}
	if (n > cap) atomicOr(&ctl->error_flags, SPH_ERR_LIST_OVERFLOW);
	rl_count[s] = min(n, cap);
}

void rigid_lists(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	k_rigid_lists<<<cdiv(c.Nr, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->rspos, h->rg.scell, h->a4[A4_POS], h->fg.cell_start,
	                                                           h->fg.sorted_id, h->rl_list, h->rl_count, h->rl_cap, h->ctl);
	h->launches++;
}

// DF:204-212: rigid_particles[j].force += (V_j rho0 k_i / rho_i * grad W_ij) * m, gathered per rigid j
__global__ void __launch_bounds__(SPH_BLOCK)
k_rigid_force_df(SphConsts c, const float4 *__restrict__ rspos, const int *__restrict__ rsorted_id,
                 const uint32_t *__restrict__ rl_list, const int *__restrict__ rl_count, int cap,
                 const float4 *__restrict__ spos, const float *__restrict__ rho, const float *__restrict__ alpha,
                 const float *__restrict__ rho_adv, float4 *__restrict__ rforce, const SphCtl *__restrict__ ctl,
                 int gated) {
	if (gated && !ctl->den_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.Nr) return;
	int n = rl_count[s];
	if (n == 0) return;
	float dt2 = ctl->dt2;
	float4 pr = rspos[s];
	int jr = rsorted_id[s];
	f3 f = xyz(rforce[jr]);
	const uint32_t *lp = rl_list + sph_list_base(s, cap);
	for (int k = 0; k < n; ++k) {
		uint32_t i = lp[(size_t)k * 32];
		float4 pi = spos[i];
		Pair p = make_pair(pi, pr);
		float k_i = ((rho_adv[i] - SPH_RHO0) * alpha[i]) / dt2;       // DF:208
		f3 ret = (((pr.w * SPH_RHO0) * k_i) / rho[i]) * cubic_dw(p, c); // DF:211
		f = f + ret * c.m;                                              // DF:212
	}
	rforce[jr] = F4(f, 0.0f);
}

// The same gather for the other solvers' scatter sites: WC:124-126, PC:185-186, II:158-159.
enum { RF_WC = 1, RF_PC = 2, RF_II = 3 };
template <int MODE>
__global__ void __launch_bounds__(SPH_BLOCK)
k_rigid_force(SphConsts c, const float4 *__restrict__ rspos, const int *__restrict__ rsorted_id,
              const uint32_t *__restrict__ rl_list, const int *__restrict__ rl_count, int cap,
              const float4 *__restrict__ spos, const float *__restrict__ rho, const float *__restrict__ press,
              float4 *__restrict__ rforce, const SphCtl *__restrict__ ctl, int gated) {
	if (gated && MODE == RF_PC && !ctl->pc_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.Nr) return;
	int n = rl_count[s];
	if (n == 0) return;
	float4 pr = rspos[s];
	int jr = rsorted_id[s];
	f3 f = xyz(rforce[jr]);
	const uint32_t *lp = rl_list + sph_list_base(s, cap);
	for (int k = 0; k < n; ++k) {
		uint32_t i = lp[(size_t)k * 32];
		Pair p = make_pair(spos[i], pr);
		float rho_i = rho[i], p_i = press[i];
		f3 dw = cubic_dw(p, c);
		if (MODE == RF_WC) {
			f3 ret = ((((-pr.w) * p_i) / (rho_i * rho_i)) * dw) * SPH_RHO0; // WC:124
			f = f + neg(ret) * c.m;                                          // WC:126
		} else if (MODE == RF_PC) {
			f3 ret = (((pr.w * SPH_RHO0) * p_i) * dw) / (rho_i * rho_i);    // PC:185
			f = f + ret * c.m;                                               // PC:186
		} else {
			f3 force = (((pr.w * SPH_RHO0) / (rho_i * rho_i)) * dw) * p_i;  // II:158
			f = f + force * c.m;                                             // II:159
		}
	}
	rforce[jr] = F4(f, 0.0f);
}

void rigid_force(SphHandle *h, int mode, int gated, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.Nr, SPH_BLOCK);
	sph_prof_begin(h, KC_RIGID, st);
#define RF_ARGS c, h->rspos, h->rg.sorted_id, h->rl_list, h->rl_count, h->rl_cap, h->a4[A4_POS], h->a1[A1_RHO], h->a1[A1_P], h->rforce, h->ctl, gated
	if (mode == RF_WC) k_rigid_force<RF_WC><<<nb, SPH_BLOCK, 0, st>>>(RF_ARGS);
	else if (mode == RF_PC) k_rigid_force<RF_PC><<<nb, SPH_BLOCK, 0, st>>>(RF_ARGS);
	else k_rigid_force<RF_II><<<nb, SPH_BLOCK, 0, st>>>(RF_ARGS);
#undef RF_ARGS
	sph_prof_end(h, st);
	h->launches++;
}

void rigid_force_df(SphHandle *h, int gated, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_RIGID, st);
	k_rigid_force_df<<<cdiv(c.Nr, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->rspos, h->rg.sorted_id, h->rl_list, h->rl_count,
	                                                              h->rl_cap, h->a4[A4_POS], h->a1[A1_RHO], h->a1[A1_ALPHA],
	                                                              h->a1[A1_RHOADV], h->rforce, h->ctl, gated);
	sph_prof_end(h, st);
	h->launches++;
}

// ---- rigid_solver.step (RS:216-234) ---------------------------------------------------------------
#define RB_THREADS 1024

// block-wide sum of a vec3 (+ an int): hierarchical warp shuffle + one shared-memory hop
__device__ inline void block_sum3(f3 &v, int &cnt, float *sh /* >= 4 * 32 floats */) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v.x += __shfl_down_sync(0xffffffffu, v.x, o);
		v.y += __shfl_down_sync(0xffffffffu, v.y, o);
		v.z += __shfl_down_sync(0xffffffffu, v.z, o);
		cnt += __shfl_down_sync(0xffffffffu, cnt, o);
	}
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) { sh[w] = v.x; sh[32 + w] = v.y; sh[64 + w] = v.z; sh[96 + w] = __int_as_float(cnt); }
	__syncthreads();
	if (threadIdx.x == 0) {
		f3 t = F3(0.0f, 0.0f, 0.0f);
		int n = 0;
		for (int k = 0; k < RB_THREADS / 32; ++k) { t.x += sh[k]; t.y += sh[32 + k]; t.z += sh[64 + k]; n += __float_as_int(sh[96 + k]); }
		v = t;
		cnt = n;
	}
	__syncthreads();
}
__device__ inline float block_max1(float v, float *sh) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) sh[w] = v;
	__syncthreads();
	float m = sh[0];
	for (int k = 1; k < RB_THREADS / 32; ++k) m = fmaxf(m, sh[k]);
	__syncthreads();
	return m;
}

__global__ void __launch_bounds__(RB_THREADS)
k_rigid_step(SphConsts c, float4 *__restrict__ rpos, float4 *__restrict__ rvel, float4 *__restrict__ rforce,
             float4 *__restrict__ rverts, int n_verts, SphRigidState *st, const SphCtl *__restrict__ ctl) {
	__shared__ float sh[128];
	__shared__ float sR[9];
	__shared__ float sv[32]; // cen(0..2) disp(3..5) ori(6..8) vel(9..11) omega(12..14) norm(15..17) cp(18..20) cnt(21)
	int tid = threadIdx.x;
	int Nr = c.Nr;
	f3 cen = ld3(st->centroid);
	float dt = st->rs_dt;
	if (ctl->ps_dt > 0.0f) dt = ctl->ps_dt; // RS:223-224

	// RS:118-128 compute_attitude: torque = sum cross(x - c, F)
	f3 torque = F3(0.0f, 0.0f, 0.0f), force = F3(0.0f, 0.0f, 0.0f);
#if SPH_STRICT
	if (tid == 0) {
		for (int i = 0; i < Nr; ++i) torque = torque + cross(xyz(rpos[i]) - cen, xyz(rforce[i]));
		for (int i = 0; i < Nr; ++i) force = force + xyz(rforce[i]); // RS:35-38 (forces do not change in between)
	}
#else
	{
		int dummy = 0;
		for (int i = tid; i < Nr; i += RB_THREADS) {
			f3 F = xyz(rforce[i]);
			torque = torque + cross(xyz(rpos[i]) - cen, F);
			force = force + F;
		}
		block_sum3(torque, dummy, sh);
		block_sum3(force, dummy, sh);
	}
#endif
	if (tid == 0) {
		f3 alpha = mat_mul_v(st->inertia_inv, torque); // RS:125
		f3 omega = ld3(st->omega) + alpha * dt;         // RS:126
		f3 att = omega * dt;                            // RS:127
		st3(st->alpha, alpha);
		st3(st->omega, omega);
		st3(st->attitude, att);
		st3(st->torque, torque);
		st3(st->force_sum, force);
		float R[9], Rt[9], T[9];
		rotation3d(-att.x, -att.z, -att.y, R); // RS:132
		for (int k = 0; k < 9; ++k) sR[k] = R[k];
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R[3 * j + i];
		mat_mul(R, st->inertia_inv, T);
		mat_mul(T, Rt, st->inertia_inv); // RS:141
		// RS:40-46
		f3 acc = force / st->mass + F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
		f3 vel = acc * dt + xyz(rvel[0]); // RS:43: the body velocity IS rigid_particles.vel[0] (a caller's vel.fill() sets it)
		f3 disp = vel * dt;
		st3(st->acc, acc);
		st3(sv + 0, cen); st3(sv + 3, disp); st3(sv + 6, disp); st3(sv + 9, vel); st3(sv + 12, omega);
	}
	__syncthreads();
	// RS:135-139 rotation of particles and mesh vertices about the centroid
	for (int i = tid; i < Nr; i += RB_THREADS) {
		float4 p = rpos[i];
		f3 q = mat_mul_v(sR, xyz(p) - cen) + cen;
		rpos[i] = F4(q, p.w);
		rforce[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // RS:38
	}
	for (int i = tid; i < n_verts; i += RB_THREADS) {
		float4 p = rverts[i];
		rverts[i] = F4(mat_mul_v(sR, xyz(p) - cen) + cen, p.w);
	}
	__syncthreads();
	// RS:53-76 wall contact scan
	f3 ori = ld3(sv + 6), vel = ld3(sv + 9), omega = ld3(sv + 12);
	float lo[3], hi[3];
	for (int j = 0; j < 3; ++j) { lo[j] = c.clamp_lo[j]; hi[j] = c.clamp_hi[j]; } // box +- particle_diameter (rigid margin)
	float dmax[3] = {-INFINITY, -INFINITY, -INFINITY}, dmin[3] = {INFINITY, INFINITY, INFINITY};
	float nrm[3] = {0.0f, 0.0f, 0.0f};
	float dcur[3] = {ori.x, ori.y, ori.z};
	f3 cp = F3(0.0f, 0.0f, 0.0f);
	int cp_cnt = 0;
#if SPH_STRICT
	int i0 = 0, istep = 1, iend = (tid == 0) ? Nr : 0;
#else
	int i0 = tid, istep = RB_THREADS, iend = Nr;
#endif
	for (int i = i0; i < iend; i += istep) {
		f3 x = xyz(rpos[i]);
		float xs[3] = {x.x, x.y, x.z}, os[3] = {ori.x, ori.y, ori.z};
		for (int j = 0; j < 3; ++j) {
			int collision = 0;
			if (xs[j] + os[j] <= lo[j]) { // RS:56
				dmax[j] = fmaxf(dmax[j], lo[j] - xs[j]);
				dcur[j] = fmaxf(dcur[j], lo[j] - xs[j]); // RS:58 in index order
				f3 v = vel + cross(omega, (x + ori) - cen);
				float vj = j == 0 ? v.x : (j == 1 ? v.y : v.z);
				if (vj < 0.0f) { collision = 1; nrm[j] = -1.0f; }
			}
			if (xs[j] + os[j] >= hi[j]) { // RS:65
				dmin[j] = fminf(dmin[j], hi[j] - xs[j]);
				dcur[j] = fminf(dcur[j], hi[j] - xs[j]); // RS:67 in index order
				f3 v = vel + cross(omega, (x + ori) - cen);
				float vj = j == 0 ? v.x : (j == 1 ? v.y : v.z);
				if (vj > 0.0f) { collision = 1; nrm[j] = 1.0f; }
			}
			if (collision == 1) { cp = cp + x; cp_cnt += 1; } // RS:74-76
		}
	}
#if !SPH_STRICT
	block_sum3(cp, cp_cnt, sh);
#endif
	for (int j = 0; j < 3; ++j) {
		float a = block_max1(dmax[j], sh);
		float b = -block_max1(-dmin[j], sh);
		float n1 = block_max1(nrm[j] > 0.0f ? 1.0f : 0.0f, sh);  // any +1 (written after a -1 in index order?)
		float n2 = block_max1(nrm[j] < 0.0f ? 1.0f : 0.0f, sh);
		if (tid == 0) {
			// atomic_max / atomic_min on the displacement (RS:58, 67), starting from ori
#if SPH_STRICT
			sv[3 + j] = dcur[j];
			sv[15 + j] = nrm[j];
			(void)n1; (void)n2; (void)a; (void)b;
#else
			float d = sv[3 + j];
			d = fmaxf(d, a);
			d = fminf(d, b);
			sv[3 + j] = d;
			sv[15 + j] = n1 > 0.0f ? 1.0f : (n2 > 0.0f ? -1.0f : 0.0f);
#endif
		}
	}
	__syncthreads();
	if (tid == 0) {
		f3 disp = ld3(sv + 3);
		st->collision_cnt = cp_cnt;
		if (cp_cnt > 0) {
			cp = (cp + ori) / (float)cp_cnt - cen;                       // RS:81
			f3 cv = vel + cross(omega, cp);                               // RS:83
			f3 n = ld3(sv + 15);
			// RS:106-116 compute_new_vel
			float cmu = (float)(0.8 * (1 + 0.1));
			f3 v_n = dot(cv, n) * n;
			f3 v_t = cv - v_n;
			float a = fmaxf(1.0f - (cmu * sqrtf(dot(v_n, v_n))) / sqrtf(dot(v_t, v_t)), 0.0f);
			f3 cv_new = a * v_t + (-0.1f) * v_n;
			float C[9] = {0.0f, -cp.z, cp.y, cp.z, 0.0f, -cp.x, -cp.y, cp.x, 0.0f};
			float CI[9], CIC[9], K[9], Kinv[9];
			mat_mul(C, st->inertia_inv, CI);
			mat_mul(CI, C, CIC);
			for (int k = 0; k < 9; ++k) {
				float id = (k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f;
				K[k] = id / st->mass - CIC[k]; // RS:89
			}
			mat_inverse(K, Kinv);
			f3 jimp = mat_mul_v(Kinv, cv_new - cv);                       // RS:92
			vel = vel + jimp / st->mass;                                  // RS:93
			omega = omega + mat_mul_v(st->inertia_inv, cross(cp, jimp));  // RS:94
		}
		st3(st->omega, omega);
		st3(st->vel, vel);
		st3(st->centroid, cen + disp); // RS:104
		st3(sv + 3, disp); st3(sv + 9, vel); st3(sv + 12, omega);
		st->rs_dt = dt;
		st->simulate_cnt += 1;
	}
	__syncthreads();
	// RS:96-102 move particles and vertices; refresh the uniform per-particle velocity
	f3 disp = ld3(sv + 3);
	vel = ld3(sv + 9);
	omega = ld3(sv + 12);
	f3 cen2 = cen + disp;
	float vmax = 0.0f;
	for (int i = tid; i < Nr; i += RB_THREADS) {
		float4 p = rpos[i];
		f3 x = xyz(p) + disp;
		rpos[i] = F4(x, p.w);
		rvel[i] = F4(vel, rvel[i].w);
		f3 w = cross(omega, x - cen2);
		vmax = fmaxf(vmax, sqrtf(dot(vel, vel)) + sqrtf(dot(w, w))); // DF:110
	}
	for (int i = tid; i < n_verts; i += RB_THREADS) {
		float4 p = rverts[i];
		rverts[i] = F4(xyz(p) + disp, p.w);
	}
	vmax = block_max1(vmax, sh);
	if (tid == 0) st->max_surface_vel = vmax;
}

void rigid_step(SphHandle *h, cudaStream_t st) {
	SphConsts c = h->c;
	// the rigid solver clamps with particle_diameter (RS:56, 65), whatever the fluid solver uses
	for (int k = 0; k < 3; ++k) {
		c.clamp_lo[k] = (float)(h->cfg.box_min[k] + h->cfg.particle_radius * 2);
		c.clamp_hi[k] = (float)(h->cfg.box_max[k] - h->cfg.particle_radius * 2);
	}
	// slabs: every rank gathered the forces of its OWNED fluid particles; the body is replicated, so the sum over
	// ranks goes to every rank and each integrates the identical state
	mg_allreduce_sum_f32(h, (float *)h->rforce, 4 * (size_t)c.Nr, st);
	sph_prof_begin(h, KC_RIGID, st);
	k_rigid_step<<<1, RB_THREADS, 0, st>>>(c, h->rpos, h->rvel, h->rforce, h->rverts, (int)h->n_rverts, h->rstate, h->ctl);
	sph_prof_end(h, st);
	h->launches++;
}

} // namespace SPH_NS
