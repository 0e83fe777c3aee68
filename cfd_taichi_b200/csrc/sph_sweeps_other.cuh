// sph_sweeps_other.cuh -- WCSPH / PCISPH / IISPH sweeps (included by sph_sweeps.cu, both modes).
// Same structure as the DFSPH section: neighbour lists built once per step by k_build_lists,
// list-walking sweeps with one float4 gather per neighbour and per-neighbour scalars riding in .w.
#pragma once

namespace SPH_NS {

// shared non-pressure force pass: tension (SB:204-217) + viscosity (SB:170-202) for fluid neighbours.
// posR.w = rho (written by k_build_lists).
#define SPH_RIGID_ENTRY(j) (rg.active && ((j) & SPH_RIGID_BIT))
// ghost copy of a neighbour rank's particle (multi-GPU slabs): never a centre particle (k_build_lists marks it)
#define SPH_IS_GHOST(L, s) ((L).fcount[s] < 0)

__device__ __forceinline__ void tension_viscosity(const SphConsts &c, const SphLists &L, const SphRigidArgs &rg, int s,
                                                  const float4 *__restrict__ posR, const float4 *__restrict__ svel,
                                                  const float *__restrict__ rho, const float4 &pi, const f3 &vi,
                                                  f3 &tension, f3 &viscosity) {
	f3 ten = F3(0.0f, 0.0f, 0.0f), visc = F3(0.0f, 0.0f, 0.0f);
	SPH_FOR_FLUID(L, c, s, j) {
		if (SPH_RIGID_ENTRY(j)) { rigid_viscosity<true>(c, rg, j, pi, vi, pi.w, rho, visc); return; } // SB:190-201
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
			float nu = c.visc_num / (pi.w + pj.w);                // SB:187
			float pi_ij = ((-nu) * shear) / (q2 + c.visc_eps_h2); // SB:188
			visc = visc + (c.neg_m * pi_ij) * cubic_dw(p, c);     // SB:189
		}
	};
	tension = ten * c.m;    // SB:209
	viscosity = visc * c.m; // SB:175
}

__device__ __forceinline__ void clamp_box(const SphConsts &c, f3 &x, f3 &v) {
	float *xp = &x.x, *vp = &v.x;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		if (xp[k] <= c.clamp_lo[k]) { xp[k] = c.clamp_lo[k]; vp[k] *= -0.5f; }
		if (xp[k] >= c.clamp_hi[k]) { xp[k] = c.clamp_hi[k]; vp[k] *= -0.5f; }
	}
}

// =============================================================================================
// WCSPH (wcsph_solver.py)
// =============================================================================================

// WC:65-68, 86-90 solve_p, and the payloads of the force pass: posT1.w = p / rho^2, velR.w = rho
__global__ void __launch_bounds__(SPH_BLOCK)
k_wc_pressure(SphConsts c, const float4 *__restrict__ posR, const float4 *__restrict__ svel,
              float *__restrict__ pressure, float4 *__restrict__ posT1, float4 *__restrict__ velR) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pr = posR[s];
	float rho = pr.w;
	float rho_i = fmaxf(rho, SPH_RHO0);
	float a = rho_i / SPH_RHO0;
	float a2 = a * a, a3 = a * a2, a4 = a2 * a2; // ** 7 by squaring
	float p = 70000.0f * (a3 * a4 - 1.0f);       // WC:89
	pressure[s] = p;
	posT1[s] = make_float4(pr.x, pr.y, pr.z, p / (rho * rho));
	float4 v = svel[s];
	velR[s] = make_float4(v.x, v.y, v.z, rho);
}

// WC:70-84 + 92-129 pressure gradient / boundary pressure, fused with viscosity and tension
// (separate accumulators, so each sum keeps the reference's order).
__global__ void __launch_bounds__(SPH_BLOCK)
k_wc_force(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT1,
           const float4 *__restrict__ velR, const float4 *__restrict__ bspos, const float *__restrict__ pressure,
           const float *__restrict__ rho,
           float4 *__restrict__ pgrad, float4 *__restrict__ visc_out, float4 *__restrict__ ten_out,
           float4 *__restrict__ bacc_out) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || L.fcount[s] < 0) return; // ghost copies (multi-GPU slabs) are never centre particles
	float4 pi = posT1[s];
	float4 vr = velR[s];
	f3 vi = xyz(vr);
	float rho_i = vr.w;
	f3 acc = F3(0.0f, 0.0f, 0.0f), ten = F3(0.0f, 0.0f, 0.0f), visc = F3(0.0f, 0.0f, 0.0f);
	float p_i = pressure[s];
	float rho_i_2 = rho_i * rho_i;
	SPH_FOR_FLUID(L, c, s, j) {
		if (SPH_RIGID_ENTRY(j)) {
			float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
			Pair p = make_pair(pi, pj);
			acc = acc + ((((-pj.w) * p_i) / rho_i_2) * cubic_dw(p, c)) * SPH_RHO0; // WC:124
			float4 pr = make_float4(pi.x, pi.y, pi.z, rho_i);
			rigid_viscosity<true>(c, rg, j, pr, vi, rho_i, rho, visc);
			return;
		}
		float4 pj = __ldg(&posT1[j]);
		float4 vj4 = __ldg(&velR[j]);
		Pair p = make_pair(pi, pj);
		f3 dw = cubic_dw(p, c);
		acc = acc - (c.m * (pi.w + pj.w)) * dw;              // WC:116
		ten = ten + (c.tension_coef * cubic_w(p, c)) * p.r;  // SB:216
		f3 v_ij = vi - xyz(vj4);
		float shear = dot(v_ij, p.r);
		if (shear < 0.0f) {
#if SPH_STRICT
			float q = sqrtf(p.r2);
			float q2 = q * q;
#else
			float q2 = p.r2;
#endif
			float nu = c.visc_num / (rho_i + vj4.w);
			float pi_ij = ((-nu) * shear) / (q2 + c.visc_eps_h2);
			visc = visc + (c.neg_m * pi_ij) * dw;
		}
	};
	f3 bacc = F3(0.0f, 0.0f, 0.0f);
	if (c.boundary_handle == 1) {
		SPH_FOR_BOUNDARY(L, c, s, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			bacc = bacc - ((pj.w * p_i) / rho_i_2) * cubic_dw(p, c); // WC:99
		};
		bacc = bacc * SPH_RHO0; // WC:83
	}
	pgrad[s] = F4(acc, 0.0f);
	visc_out[s] = F4(visc * c.m, 0.0f);
	ten_out[s] = F4(ten * c.m, 0.0f);
	bacc_out[s] = F4(bacc, 0.0f);
}

// WC:40-63 kinematic_phase (+ SB:131-134 reset: acc = gravity) and write-back to original order
__global__ void __launch_bounds__(SPH_BLOCK)
k_wc_kinematic(SphConsts c, const int *__restrict__ sorted_id, const float4 *__restrict__ spos,
               const float4 *__restrict__ svel, const float4 *__restrict__ pgrad, const float4 *__restrict__ visc,
               const float4 *__restrict__ ten, const float4 *__restrict__ bacc, float4 *__restrict__ pos,
               float4 *__restrict__ vel, float4 *__restrict__ acc_out, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	int i = sorted_id[s];
	if (i >= c.N_owned) return;
	float dt = ctl->dt;
	f3 a = F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
	f3 add = (xyz(pgrad[s]) + xyz(visc[s])) + xyz(ten[s]);
	if (c.boundary_handle == 1) add = add + xyz(bacc[s]);
	a = a + add;                          // WC:44-47
	float4 v4 = svel[s];
	f3 v = xyz(v4) + a * dt;              // WC:50
	v = v * 0.9998f;                      // WC:51
	f3 x = xyz(spos[s]) + v * dt;         // WC:52
	if (c.boundary_handle == 0) clamp_box(c, x, v); // WC:54-63 (margin = particle_diameter)
	pos[i] = F4(x, 0.0f);
	vel[i] = F4(v, v4.w);
	if (acc_out) acc_out[i] = F4(a, 0.0f);
}

// WC:32-38 Tait pressure per particle (+ the payload copies the force sweep gathers)
static void wc_eos(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_WC_FORCE, st);
	k_wc_pressure<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->a4[A4_PR], h->a4[A4_VEL], h->a1[A1_P], h->a4[A4_T1], h->a4[A4_VADV]);
	sph_prof_end(h, st);
	h->launches += 1;
}
// WC:65-144 pressure gradient + boundary term + viscosity + tension (+ fluid -> rigid forces, WC:126)
static void wc_force(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	SphRigidArgs rg = rigid_args(h);
	sph_prof_begin(h, KC_WC_FORCE, st);
	k_wc_force<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, rg, h->a4[A4_T1], h->a4[A4_VADV], h->bspos, h->a1[A1_P],
	                                                      h->a1[A1_RHO], h->a4[A4_FA], h->a4[A4_FB], h->a4[A4_FC], h->a4[A4_FD]);
	sph_prof_end(h, st);
	h->launches += 1;
	if (rg.active) { rigid_lists(h, st); rigid_force(h, RF_WC, 0, st); } // WC:126, gather form
}

void wc_phase(SphHandle *h, int phase, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	if (phase == SPH_PH_WC_PRESSURE) {
		first_phase_lists(h, st); // slabs: rho of the ghost particles travels here; p follows from it pointwise
		wc_eos(h, st);
		wc_force(h, st);
	} else if (phase == SPH_PH_WC_EOS) {
		wc_eos(h, st);
	} else if (phase == SPH_PH_WC_FORCE) {
		wc_force(h, st);
	} else if (phase == SPH_PH_WC_KINEMATIC) {
		sph_prof_begin(h, KC_WC_KIN, st);
		k_wc_kinematic<<<nba, SPH_BLOCK, 0, st>>>(c, h->fg.sorted_id, h->a4[A4_POS], h->a4[A4_VEL], h->a4[A4_FA],
		                                          h->a4[A4_FB], h->a4[A4_FC], h->a4[A4_FD], h->pos, h->vel, h->acc, h->ctl);
		sph_prof_end(h, st);
		h->launches += 1;
	}
}

// =============================================================================================
// PCISPH (pcisph_solver.py)
// =============================================================================================

// PC:220-226 compute_ext_force (rho comes from k_build_lists) + PC:228-231 reset + first
// PC:72-87 predict_vel_pos (press_force = 0)
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_ext_force(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posR,
               const float4 *__restrict__ svel, const float *__restrict__ rho, float4 *__restrict__ ext_force, float4 *__restrict__ press_force, float *__restrict__ press,
               float4 *__restrict__ posT1) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pi = posR[s];
	press[s] = 0.0f;                                 // PC:230
	press_force[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // PC:231
	posT1[s] = make_float4(pi.x, pi.y, pi.z, 0.0f);  // payload of the force pass: press_iter
	if (SPH_IS_GHOST(L, s)) return;
	f3 vi = xyz(svel[s]);
	f3 tension, viscosity;
	tension_viscosity(c, L, rg, s, posR, svel, rho, pi, vi, tension, viscosity);
	f3 g = F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
	ext_force[s] = F4((g + tension) + viscosity, 0.0f); // PC:226
}

// PC:72-87 predict_vel_pos
__device__ __forceinline__ void pc_predict(const SphConsts &c, float dt, f3 x, f3 v, f3 ext, f3 press,
                                           float4 *pos_predict, float4 *vel_predict, int s) {
	f3 vp = v + (dt * (ext + press)) / c.m; // PC:75
	f3 xp = x + dt * vp;                    // PC:76
	if (c.boundary_handle == 0) clamp_box(c, xp, vp); // PC:78-87
	vel_predict[s] = F4(vp, 0.0f);
	pos_predict[s] = F4(xp, 0.0f);
}

__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_predict(SphConsts c, SphLists L, const float4 *__restrict__ spos, const float4 *__restrict__ svel,
             const float4 *__restrict__ ext_force, const float4 *__restrict__ press_force,
             float4 *__restrict__ pos_predict, float4 *__restrict__ vel_predict, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || SPH_IS_GHOST(L, s)) return;
	pc_predict(c, ctl->dt, xyz(spos[s]), xyz(svel[s]), xyz(ext_force[s]), xyz(press_force[s]), pos_predict,
	           vel_predict, s);
}

// PC:89-101 predict_rho (+ PC:121-133 residual partials, + speculative PC:103-107 iter_press into
// p_next: it is committed by the next force pass only if the loop continues)
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_predict_rho(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ pos_predict,
                 const float4 *__restrict__ bspos,
                 float *__restrict__ rho_predict, float *__restrict__ rho_err, const float *__restrict__ press,
                 float *__restrict__ p_next, const SphCtl *__restrict__ ctl, SphPartial *__restrict__ partials,
                 int gated) {
	if (gated && !ctl->pc_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	double psum = 0.0;
	int pcnt = 0;
	if (s < c.N && !SPH_IS_GHOST(L, s)) {
		float4 pi = pos_predict[s];
		float rp = 0.0f;
		SPH_FOR_FLUID(L, c, s, j) {
			if (SPH_RIGID_ENTRY(j)) {
				float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
				f3 d = xyz(pi) - xyz(pj);
				float q = sqrtf(dot(d, d));                      // PC:146
				rp += (cubic_w_r(q, c) * pj.w) * SPH_RHO0;       // PC:147
				return;
			}
			float4 pj = __ldg(&pos_predict[j]);
			f3 d = xyz(pi) - xyz(pj);
			float q = sqrtf(dot(d, d));          // PC:141
			rp += cubic_w_r(q, c) * c.m;         // PC:142
		};
		float out = rp;
		if (c.boundary_handle == 1) {
			float rb = 0.0f;
			SPH_FOR_BOUNDARY(L, c, s, j) {
				float4 pj = __ldg(&bspos[j]);
				f3 d = xyz(pi) - xyz(pj);
				float q = sqrtf(dot(d, d));      // PC:152
				rb += cubic_w_r(q, c) * pj.w;    // PC:153
			};
			out = rp + rb * SPH_RHO0;            // PC:98
		}
		float err = out - SPH_RHO0;              // PC:101
		rho_predict[s] = out;
		rho_err[s] = err;
		p_next[s] = fmaxf(0.0f, press[s] + err * ctl->pc_delta); // PC:106-107
		float e = fmaxf(err, 0.0f);
		if (e > 0.0f) { psum = (double)e; pcnt = 1; }
	}
	block_partial(psum, pcnt, 0.0f, partials);
}

// commits iter_press (PC:103-107) for every particle and publishes it as the .w payload
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_commit_press(SphConsts c, SphLists L, const float *__restrict__ p_next, float *__restrict__ press,
                  float4 *__restrict__ posT1, const SphCtl *__restrict__ ctl) {
	if (!ctl->pc_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || SPH_IS_GHOST(L, s)) return; // ghosts receive posT1.w through the slab exchange
	float p = p_next[s];
	press[s] = p;
	posT1[s].w = p;
}

// PC:109-119 update_press_force fused with the following PC:72-87 predict_vel_pos
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_press_force(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT1,
                 const float4 *__restrict__ bspos,
                 const float *__restrict__ rho, const float4 *__restrict__ svel, const float4 *__restrict__ ext_force,
                 float4 *__restrict__ press_force, float4 *__restrict__ pos_predict, float4 *__restrict__ vel_predict,
                 const SphCtl *__restrict__ ctl) {
	if (!ctl->pc_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || SPH_IS_GHOST(L, s)) return;
	float4 pi = posT1[s];
	f3 pf = F3(0.0f, 0.0f, 0.0f), pf_fluid = pf;
	(void)pf_fluid;
	float rho_i = rho[s];
	float rho_i_2 = rho_i * rho_i;
	SPH_FOR_FLUID(L, c, s, j) {
		if (SPH_RIGID_ENTRY(j)) {
			float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
			Pair p = make_pair(pi, pj);
			f3 ret = (((pj.w * SPH_RHO0) * pi.w) * cubic_dw(p, c)) / rho_i_2; // PC:185
			pf = pf + ret * c.m;                                                // PC:187
			return;
		}
		float4 pj = __ldg(&posT1[j]);
		Pair p = make_pair(pi, pj);
#if SPH_STRICT
		pf = pf + ((((pi.w + pj.w) * cubic_dw(p, c)) / 1000000.0f) * c.m) * c.m; // PC:177
#else
		pf_fluid = pf_fluid + (pi.w + pj.w) * cubic_dw(p, c); // the constant factor m^2 / rho_0^2 is applied once below
#endif
	};
#if !SPH_STRICT
	pf = pf + pf_fluid * ((c.m * c.m) / 1000000.0f);
#endif
	f3 out = neg(pf);
	if (c.boundary_handle == 1) {
		f3 bacc = F3(0.0f, 0.0f, 0.0f);
		const float p_over_rho2 = pi.w / rho_i_2;
		(void)p_over_rho2;
		SPH_FOR_BOUNDARY(L, c, s, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
#if SPH_STRICT
			bacc = bacc - ((pj.w * pi.w) / rho_i_2) * cubic_dw(p, c); // PC:197
#else
			bacc = bacc - (pj.w * p_over_rho2) * cubic_dw(p, c);
#endif
		};
		out = neg(pf) + (bacc * SPH_RHO0) * c.m; // PC:117
	}
	press_force[s] = F4(out, 0.0f);
	pc_predict(c, ctl->dt, xyz(pi), xyz(svel[s]), xyz(ext_force[s]), out, pos_predict, vel_predict, s);
}

// mode 0: first evaluation (PC:54); mode 1: end of a loop body (PC:68-69)
__global__ void __launch_bounds__(1024) k_pc_ctl(SphCtl *ctl, const SphPartial *partials, int n, int mode) {
	if (mode == 1 && !ctl->pc_active) return;
	double sum; int cnt; float mx;
	sph_reduce_partials<1024>(partials, n, sum, cnt, mx); // 1024 threads, 8 loads in flight: ~4 us instead of ~10 (pure L2 latency)
	if (threadIdx.x != 0) return;
	SphCtlArgs none = {};
	sph_ctl_apply(mode == 0 ? SPH_CTL_PC_FIRST : SPH_CTL_PC_ITER, ctl, sum, cnt, mx, none); // PC:54-56, 68-69
}
// the loop decision after a residual sweep: one controller launch, or (slabs) part of the exchange
static void pc_decide(SphHandle *h, int mode, int nb, cudaStream_t st) {
	if (h->comm) { mg_exchange_reduce(h, MG_NONE, mode == 0 ? SPH_CTL_PC_FIRST : SPH_CTL_PC_ITER, nb, st); return; }
	k_pc_ctl<<<1, 1024, 0, st>>>(h->ctl, h->partials, nb, mode);
	h->launches++;
}

// PC:200-218 integration + write-back
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_integration(SphConsts c, const int *__restrict__ sorted_id, const float4 *__restrict__ spos,
                 const float4 *__restrict__ svel, const float4 *__restrict__ ext_force,
                 const float4 *__restrict__ press_force, float4 *__restrict__ pos, float4 *__restrict__ vel,
                 const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	int i = sorted_id[s];
	if (i >= c.N_owned) return;
	float dt = ctl->dt;
	float4 v4 = svel[s];
	f3 v = xyz(v4) + (dt * (xyz(ext_force[s]) + xyz(press_force[s]))) / c.m; // PC:203-204
	v = v * 0.9999f;                                                          // PC:205
	f3 x = xyz(spos[s]) + dt * v;                                             // PC:206
	if (c.boundary_handle == 0) clamp_box(c, x, v);
	pos[i] = F4(x, 0.0f);
	vel[i] = F4(v, v4.w);
}

// PC:39-45 pre_compute_delta for the particle with original index `target` (un-weighted sums, PC:156-167)
__global__ void __launch_bounds__(SPH_BLOCK)
k_pc_delta(SphConsts c, SphLists L, const float4 *__restrict__ spos, const float4 *__restrict__ rspos,
           const int *__restrict__ sorted_id, int target, SphCtl *ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || sorted_id[s] != target) return;
	float4 pi = spos[s];
	f3 sum = F3(0.0f, 0.0f, 0.0f);
	float sq = 0.0f;
	SPH_FOR_FLUID(L, c, s, j) {
		Pair p = make_pair(pi, (j & SPH_RIGID_BIT) ? rspos[j & ~SPH_RIGID_BIT] : spos[j]); // no material test (PC:156-167)
		f3 dw = cubic_dw(p, c);
		sum = sum + dw;
		sq += dot(dw, dw);
	};
	ctl->pc_delta = 1.0f / ((dot(sum, sum) + sq) * c.pc_beta); // PC:45
	ctl->pc_max_index = target;
}

void pc_precompute(SphHandle *h, cudaStream_t st) {
	// PC:28-37: grid (done by the caller), neighbour counts; the arg-max (PS:409-422) is evaluated by the
	// Python mirror from the fetched counts and passed back through h->cfg.reserved.
	build_lists(h, st);
}

void pc_set_delta(SphHandle *h, int target, cudaStream_t st) {
	const SphConsts &c = h->c;
	k_pc_delta<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, h->a4[A4_POS], h->rspos, h->fg.sorted_id, target, h->ctl);
	h->launches++;
}

// PC:72-87 predict_vel_pos at the current pressure force (the first evaluation of a step; later ones are fused
// into the pressure-force sweep)
static void pc_predict(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	if (rigid_args(h).active) rigid_lists(h, st);
	sph_prof_begin(h, KC_PC_PREDICT, st);
	k_pc_predict<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, h->a4[A4_POS], h->a4[A4_VEL], h->a4[A4_FA], h->a4[A4_FB],
	                                                        h->a4[A4_T2], h->a4[A4_VADV], h->ctl);
	sph_prof_end(h, st);
	mg_exchange(h, MG_XYZ(A4_T2), st); // slabs: predicted positions of the ghost particles
	h->launches += 1;
}
// PC:89-100 predict_rho + PC:123-133 compute_residual + the loop decision (PC:54 before the loop, PC:56 inside it)
static void pc_rho(SphHandle *h, int in_loop, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	sph_prof_begin(h, KC_PC_RHO, st);
	k_pc_predict_rho<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, rigid_args(h), h->a4[A4_T2], h->bspos, h->a1[A1_SA], h->a1[A1_SB],
	                                            h->a1[A1_P], h->a1[A1_SC], h->ctl, h->partials, in_loop);
	sph_prof_end(h, st);
	pc_decide(h, in_loop, nba, st);
	h->launches += 1;
}
// PC:102-121 iter_press + update_press_force (+ predict_vel_pos of the next evaluation, + fluid -> rigid forces PC:186)
static void pc_press_force(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	SphRigidArgs rg = rigid_args(h);
	k_pc_commit_press<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, h->a1[A1_SC], h->a1[A1_P], h->a4[A4_T1], h->ctl);
	mg_exchange(h, MG_F4_T1W, st); // slabs: press_iter of the ghost particles
	sph_prof_begin(h, KC_PC_FORCE, st);
	k_pc_press_force<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, rg, h->a4[A4_T1], h->bspos, h->a1[A1_RHO], h->a4[A4_VEL],
	                                            h->a4[A4_FA], h->a4[A4_FB], h->a4[A4_T2], h->a4[A4_VADV], h->ctl);
	sph_prof_end(h, st);
	mg_exchange(h, MG_XYZ(A4_T2), st);
	if (rg.active) rigid_force(h, RF_PC, 1, st); // PC:186, gather form
	h->launches += 2;
}

// PC:47-55: predicted state at zero pressure correction, its density error and the decision whether the loop starts
static void pc_iteration_begin(SphHandle *h, cudaStream_t st) {
	pc_predict(h, st);
	pc_rho(h, 0, st);
}

// PC:56-70: `count` passes of the loop body, each gated on ctl->pc_active (a no-op once the loop has ended)
static void pc_iteration_passes(SphHandle *h, int count, cudaStream_t st) {
	for (int it = 0; it < count; ++it) {
		pc_press_force(h, st);
		pc_rho(h, 1, st);
	}
}

static void pc_iteration(SphHandle *h, cudaStream_t st) {
	pc_iteration_begin(h, st);
	int done = 0;
	int chunk = h->last_den_chunk > 0 ? h->last_den_chunk : 4;
	for (;;) {
		int n = chunk < 80 - done ? chunk : 80 - done; // max_iteration (PC:21)
		pc_iteration_passes(h, n, st);
		done += n;
		if (done >= 80) break;
		// one look at the device flag per chunk (not per iteration)
		cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(SphCtl), cudaMemcpyDeviceToHost, st);
		cudaStreamSynchronize(st);
		if (!h->ctl_host->pc_active) break;
		chunk = 4;
	}
	if (done < 80) h->last_den_chunk = h->ctl_host->pc_iters + 1;
	else h->last_den_chunk = 80;
}

void pc_phase(SphHandle *h, int phase, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	if (phase == SPH_PH_PC_EXT_FORCE) {
		first_phase_lists(h, st);
		sph_prof_begin(h, KC_PC_EXT, st);
		k_pc_ext_force<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, rigid_args(h), h->a4[A4_PR], h->a4[A4_VEL], h->a1[A1_RHO],
		                                          h->a4[A4_FA], h->a4[A4_FB], h->a1[A1_P], h->a4[A4_T1]);
		sph_prof_end(h, st);
		h->launches++;
	} else if (phase == SPH_PH_PC_ITERATION) {
		pc_iteration(h, st);
	} else if (phase == SPH_PH_PC_PREDICT) { // the loop one sweep at a time (single-sweep parity tests)
		pc_predict(h, st);
	} else if (phase == SPH_PH_PC_RHO_FIRST) {
		pc_rho(h, 0, st);
	} else if (phase == SPH_PH_PC_PRESS_FORCE) {
		pc_press_force(h, st);
	} else if (phase == SPH_PH_PC_RHO) {
		pc_rho(h, 1, st);
	} else if (phase == SPH_PH_PC_INTEGRATION) {
		sph_prof_begin(h, KC_PC_INT, st);
		k_pc_integration<<<nba, SPH_BLOCK, 0, st>>>(c, h->fg.sorted_id, h->a4[A4_POS], h->a4[A4_VEL], h->a4[A4_FA],
		                                            h->a4[A4_FB], h->pos, h->vel, h->ctl);
		sph_prof_end(h, st);
		h->launches++;
	}
}

// =============================================================================================
// IISPH (iisph_solver.py)
// =============================================================================================

// the recurring factor  - dt * dt * m / (rho_i * rho_i)  (II:244-245, 283-284, 291-292, 301-302)
__device__ __forceinline__ float ii_dji_coef(const SphConsts &c, float dt, float rho_i) {
	return (((-dt) * dt) * c.m) / (rho_i * rho_i);
}

// II:42-55: tension, viscosity, f_adv, v_adv and d_ii in one pass over the lists
__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_advect(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posR,
            const float4 *__restrict__ svel, const float *__restrict__ rho, const float4 *__restrict__ bspos, float4 *__restrict__ f_adv, float4 *__restrict__ v_adv,
            float4 *__restrict__ d_ii, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || SPH_IS_GHOST(L, s)) return;
	float dt = ctl->dt;
	float4 pi = posR[s];
	float rho_i = pi.w;
	f3 vi = xyz(svel[s]);
	f3 ten = F3(0.0f, 0.0f, 0.0f), visc = F3(0.0f, 0.0f, 0.0f), dii = F3(0.0f, 0.0f, 0.0f);
	float cf = (-c.m) / (rho_i * rho_i); // II:261
	SPH_FOR_FLUID(L, c, s, j) {
		if (SPH_RIGID_ENTRY(j)) {
			float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
			Pair p = make_pair(pi, pj);
			rigid_viscosity<true>(c, rg, j, pi, vi, rho_i, rho, visc);
			dii = dii + (((-pj.w) * SPH_RHO0) / (rho_i * rho_i)) * cubic_dw(p, c); // II:267
			return;
		}
		float4 pj = __ldg(&posR[j]);
		f3 vj = xyz(__ldg(&svel[j]));
		Pair p = make_pair(pi, pj);
		f3 dw = cubic_dw(p, c);
		ten = ten + (c.tension_coef * cubic_w(p, c)) * p.r;
		f3 v_ij = vi - vj;
		float shear = dot(v_ij, p.r);
		if (shear < 0.0f) {
#if SPH_STRICT
			float q = sqrtf(p.r2);
			float q2 = q * q;
#else
			float q2 = p.r2;
#endif
			float nu = c.visc_num / (rho_i + pj.w);
			float pi_ij = ((-nu) * shear) / (q2 + c.visc_eps_h2);
			visc = visc + (c.neg_m * pi_ij) * dw;
		}
		dii = dii + cf * dw;
	};
	f3 g = F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
	f3 f = (g + ten * c.m) + visc * c.m;      // II:45
	f_adv[s] = F4(f, 0.0f);
	v_adv[s] = F4(vi + (dt * f) / c.m, 0.0f); // II:47
	if (c.boundary_handle == 1) {
		f3 db = F3(0.0f, 0.0f, 0.0f);
		SPH_FOR_BOUNDARY(L, c, s, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			db = db + ((-pj.w) / (rho_i * rho_i)) * cubic_dw(p, c); // II:273
		};
		d_ii[s] = F4(((dii + db * SPH_RHO0) * dt) * dt, 0.0f); // II:53
	} else {
		d_ii[s] = F4((dii * dt) * dt, 0.0f);
	}
}

// II:57-75: rho_adv, p_iter = 0.5 p_past, a_ii
__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_rho_adv_aii(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posR,
                 const float4 *__restrict__ v_adv,
                 const float4 *__restrict__ bspos, const float4 *__restrict__ d_ii, const float4 *__restrict__ svel,
                 float *__restrict__ rho, float *__restrict__ rho_adv, float *__restrict__ a_ii, float *__restrict__ press,
                 float4 *__restrict__ posT1, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pi = posR[s];
	float p0 = 0.5f * svel[s].w; // II:67 (ghost copies carry their p_past in vel.w as well)
	press[s] = p0;
	posT1[s] = make_float4(pi.x, pi.y, pi.z, p0);
	if (SPH_IS_GHOST(L, s)) { rho[s] = pi.w; return; } // slabs: k_ii_dij gathers rho[j] of ghosts too
	float dt = ctl->dt;
	float rho_i = pi.w;
	f3 va = xyz(v_adv[s]);
	f3 dii = xyz(d_ii[s]);
	float coef = ii_dji_coef(c, dt, rho_i);
	float ra = 0.0f, aii = 0.0f;
	SPH_FOR_FLUID(L, c, s, j) {
		if (SPH_RIGID_ENTRY(j)) {
			float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
			Pair p = make_pair(pi, pj);
			f3 dw = cubic_dw(p, c);
			f3 v_j = rigid_velocity(rg.st, xyz(pj), dt, true);      // II:328-330
			ra += (pj.w * dot(va - v_j, dw)) * SPH_RHO0;            // II:333
			f3 d_ji = coef * neg(dw);                               // II:291-292
			aii += (pj.w * dot(dii - d_ji, dw)) * SPH_RHO0;         // II:293
			return;
		}
		float4 pj = __ldg(&posR[j]);
		f3 vj = xyz(__ldg(&v_adv[j]));
		Pair p = make_pair(pi, pj);
		f3 dw = cubic_dw(p, c);
		ra += c.m * dot(va - vj, dw);           // II:324
		f3 d_ji = coef * neg(dw);               // II:283-284: kernel derivative of -q is the negated vector
		aii += c.m * dot(dii - d_ji, dw);       // II:285
	};
	if (c.boundary_handle == 1) {
		float rab = 0.0f, ab = 0.0f;
		SPH_FOR_BOUNDARY(L, c, s, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			f3 dw = cubic_dw(p, c);
			rab += pj.w * dot(va, dw);          // II:340
			f3 d_ji = coef * neg(dw);
			ab += pj.w * dot(dii - d_ji, dw);   // II:303
		};
		rho_adv[s] = (ra + rab * SPH_RHO0) * dt + rho_i; // II:63
		a_ii[s] = aii + ab * SPH_RHO0;                    // II:73
	} else {
		rho_adv[s] = ra * dt + rho_i;
		a_ii[s] = aii;
	}
}

// II:121-126, 305-314 compute_all_d_ij.  posT1.w = p_iter of the neighbour.
__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_dij(SphConsts c, SphLists L, const float4 *__restrict__ posT1, const float *__restrict__ rho,
         float4 *__restrict__ d_ij, const float4 *__restrict__ d_ii, float4 *__restrict__ q_out,
         const SphCtl *__restrict__ ctl) {
	if (!ctl->ii_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || L.fcount[s] < 0) return;
	float dt = ctl->dt;
	float4 pi = posT1[s];
	f3 dij = F3(0.0f, 0.0f, 0.0f);
	SPH_FOR_FLUID(L, c, s, j) {
		if (j & SPH_RIGID_BIT) return; // II:308: fluid neighbours only
		float4 pj = __ldg(&posT1[j]);
		float rho_j = __ldg(&rho[j]);
		Pair p = make_pair(pi, pj);
#if SPH_STRICT
		dij = dij + (((-c.m) * pj.w) * cubic_dw(p, c)) / (rho_j * rho_j); // II:313
#else
		dij = dij + sdiv((-c.m) * pj.w, rho_j * rho_j) * cubic_dw(p, c); // one approximate reciprocal instead of three divisions
#endif
	};
	f3 d = (dij * dt) * dt; // II:126
	d_ij[s] = F4(d, 0.0f);
	// what update_p needs of a NEIGHBOUR j is d_ii_j p_j + sum_k d_jk p_k (II:246): the fast kernels gather this one
	// vector instead of the two (the strict kernels write it too, so that both modes keep the same set of work arrays)
	q_out[s] = F4(xyz(d_ii[s]) * pi.w + d, 0.0f);
}

// II:128-147 update_p (sum_factor II:228-253) + residual partials (II:102-113); p_next is committed
// by k_ii_commit so that neighbours keep reading the current iterate.
__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_update_p(SphConsts c, SphLists L, SphRigidArgs rg, const float4 *__restrict__ posT1,
              const float4 *__restrict__ bspos,
              const float *__restrict__ rho, const float4 *__restrict__ d_ij, const float4 *__restrict__ d_ii,
              const float4 *__restrict__ q_in,
              const float *__restrict__ a_ii, const float *__restrict__ rho_adv, float *__restrict__ r_sum,
              float *__restrict__ p_next, const SphCtl *__restrict__ ctl, SphPartial *__restrict__ partials) {
	if (!ctl->ii_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	double psum = 0.0;
	int pcnt = 0;
	if (s < c.N && !SPH_IS_GHOST(L, s)) {
		float dt = ctl->dt;
		float4 pi = posT1[s];
		float rho_i = rho[s];
		f3 dij_i = xyz(d_ij[s]);
		float coef = ii_dji_coef(c, dt, rho_i);
		float sum = 0.0f;
		SPH_FOR_FLUID(L, c, s, j) {
			if (SPH_RIGID_ENTRY(j)) {
				float4 pj = __ldg(&rg.rspos[j & ~SPH_RIGID_BIT]);
				Pair p = make_pair(pi, pj);
				sum += (dot(dij_i, cubic_dw(p, c)) * pj.w) * SPH_RHO0; // II:252
				return;
			}
			float4 pj = __ldg(&posT1[j]);
			Pair p = make_pair(pi, pj);
			f3 w_ij = cubic_dw(p, c);
			f3 d_ji = (coef * neg(w_ij)) * pi.w;                       // II:244-245
#if SPH_STRICT
			f3 dij_j = xyz(__ldg(&d_ij[j]));
			f3 dii_j = xyz(__ldg(&d_ii[j]));
			f3 t = (dij_i - dii_j * pj.w) - (dij_j - d_ji);            // II:246
#else
			f3 t = (dij_i + d_ji) - xyz(__ldg(&q_in[j]));              // the same terms, the neighbour's two in one gather
#endif
			sum += c.m * dot(t, w_ij);
		};
		float rs = sum;
		if (c.boundary_handle == 1) {
			float bsum = 0.0f;
			SPH_FOR_BOUNDARY(L, c, s, j) {
				float4 pj = __ldg(&bspos[j]);
				Pair p = make_pair(pi, pj);
				bsum += (dot(dij_i, cubic_dw(p, c)) * pj.w) * SPH_RHO0; // II:232
			};
			rs = sum + bsum; // II:136
		}
		r_sum[s] = rs;
		float a = a_ii[s], ra = rho_adv[s];
		float pn;
		if (fabsf(a) > 1e-7f) pn = 0.5f * pi.w + (0.5f * ((SPH_RHO0 - ra) - rs)) / a; // II:140-142
		else pn = 0.0f;
		pn = fmaxf(pn, 0.0f); // II:147
		p_next[s] = pn;
		if (pn > 0.0f) { psum = (double)(((a * pn + rs) + ra) - 1000.0f); pcnt = 1; } // II:108-110
	}
	block_partial(psum, pcnt, 0.0f, partials);
}

__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_commit(SphConsts c, SphLists L, const float *__restrict__ p_next, float *__restrict__ press, float4 *__restrict__ posT1,
            const SphCtl *__restrict__ ctl) {
	if (!ctl->ii_active) return;
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N || SPH_IS_GHOST(L, s)) return; // ghosts receive posT1.w through the slab exchange
	float p = p_next[s];
	press[s] = p;
	posT1[s].w = p;
}

// II:78-100 loop control.  mode 0 = before the loop, mode 1 = after an iteration.
__global__ void __launch_bounds__(1024) k_ii_ctl(SphCtl *ctl, const SphPartial *partials, int n, int mode) {
	if (mode == 0) {
		if (threadIdx.x == 0) { ctl->ii_active = 1; ctl->ii_iters = 0; ctl->ii_have_last = 0; ctl->ii_residual = INFINITY; }
		return;
	}
	if (!ctl->ii_active) return;
	double sum; int cnt; float mx;
	sph_reduce_partials<1024>(partials, n, sum, cnt, mx); // 1024 threads, 8 loads in flight: ~4 us instead of ~10 (pure L2 latency)
	if (threadIdx.x != 0) return;
	SphCtlArgs none = {};
	sph_ctl_apply(SPH_CTL_II_ITER, ctl, sum, cnt, mx, none); // II:83-93, 102-113
}

// II:184-206 intergation (sic) + write-back
__global__ void __launch_bounds__(SPH_BLOCK)
k_ii_integration(SphConsts c, const int *__restrict__ sorted_id, const float4 *__restrict__ spos,
                 const float4 *__restrict__ v_adv, const float4 *__restrict__ d_ij, const float4 *__restrict__ d_ii,
                 const float *__restrict__ press, float4 *__restrict__ f_press, float4 *__restrict__ pos,
                 float4 *__restrict__ vel, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	int i = sorted_id[s];
	if (i >= c.N_owned) return;
	float dt = ctl->dt;
	float p = press[s];
	f3 f = ((xyz(d_ij[s]) + xyz(d_ii[s]) * p) * c.m) / (dt * dt); // II:167
	f_press[s] = F4(f, 0.0f);
	f3 v = xyz(v_adv[s]) + (dt * f) / c.m; // II:189
	v = v * 0.9999f;                       // II:190
	f3 x = xyz(spos[s]) + dt * v;          // II:191
	if (c.boundary_handle == 0) clamp_box(c, x, v);
	pos[i] = F4(x, 0.0f);
	vel[i] = F4(v, p);                     // II:206 p_past = p_iter
}

// II:78-82: residual of the start iterate and the decision whether the loop starts
static void ii_pressure_solve_begin(SphHandle *h, cudaStream_t st) {
	k_ii_ctl<<<1, 1024, 0, st>>>(h->ctl, h->partials, cdiv(h->c.N, SPH_BLOCK), 0);
	h->launches++;
}

// II:208-250 compute_all_d_ij: sum_j d_ij p_j, gated on ctl->ii_active
static void ii_dij(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_II_DIJ, st);
	k_ii_dij<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, h->a4[A4_T1], h->a1[A1_RHO], h->a4[A4_FB], h->a4[A4_FC],
	                                                     h->a4[A4_T2], h->ctl);
	sph_prof_end(h, st);
#if SPH_STRICT
	mg_exchange(h, MG_XYZ(A4_FB), st); // slabs: sum_j d_ij p_j of the ghost particles
#else
	mg_exchange(h, MG_XYZ(A4_T2), st); // slabs: d_ii p + sum_j d_ij p_j of the ghost particles
#endif
	h->launches += 1;
}
// II:252-340 update_p (relaxed Jacobi, omega = 0.5) + II:102-113 compute_residual + the loop decision (II:83-93)
static void ii_update(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	sph_prof_begin(h, KC_II_UPDATE, st);
	k_ii_update_p<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, rigid_args(h), h->a4[A4_T1], h->bspos, h->a1[A1_RHO], h->a4[A4_FB],
	                                         h->a4[A4_FC], h->a4[A4_T2], h->a1[A1_SA], h->a1[A1_RHOADV], h->a1[A1_SB],
	                                         h->a1[A1_SC], h->ctl, h->partials);
	sph_prof_end(h, st);
	k_ii_commit<<<nba, SPH_BLOCK, 0, st>>>(c, h->L, h->a1[A1_SC], h->a1[A1_P], h->a4[A4_T1], h->ctl);
	if (h->comm) {
		mg_exchange(h, MG_F4_T1W, st); // slabs: the new pressure iterate of the ghost particles
		mg_exchange_reduce(h, MG_NONE, SPH_CTL_II_ITER, nba, st);
	} else {
		k_ii_ctl<<<1, 1024, 0, st>>>(h->ctl, h->partials, nba, 1);
	}
	h->launches += 3;
}

// II:83-100: `count` relaxed Jacobi passes, each gated on ctl->ii_active
static void ii_pressure_solve_passes(SphHandle *h, int count, cudaStream_t st) {
	for (int it = 0; it < count; ++it) {
		ii_dij(h, st);
		ii_update(h, st);
	}
}

static void ii_pressure_solve(SphHandle *h, cudaStream_t st) {
	ii_pressure_solve_begin(h, st);
	int done = 0;
	int chunk = h->last_den_chunk > 0 ? h->last_den_chunk : 4;
	for (;;) {
		int n = chunk < 180 - done ? chunk : 180 - done; // max_iter_cnt (II:27)
		ii_pressure_solve_passes(h, n, st);
		done += n;
		if (done >= 180) break;
		cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(SphCtl), cudaMemcpyDeviceToHost, st);
		cudaStreamSynchronize(st);
		if (!h->ctl_host->ii_active) break;
		chunk = 4;
	}
	h->last_den_chunk = done < 180 ? h->ctl_host->ii_iters + 1 : 180;
}

// II:35-54 advection force, v_adv and d_ii in one pass
static void ii_advect(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_II_ADV, st);
	k_ii_advect<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, rigid_args(h), h->a4[A4_PR], h->a4[A4_VEL], h->a1[A1_RHO],
	                                                       h->bspos, h->a4[A4_FA], h->a4[A4_VADV], h->a4[A4_FC], h->ctl);
	sph_prof_end(h, st);
	mg_exchange(h, MG_F4_VADV, st);    // slabs: v_adv and d_ii of the ghost particles
	mg_exchange(h, MG_XYZ(A4_FC), st);
	h->launches += 1;
}
// II:56-75 rho_adv, a_ii and the start iterate p = 0.5 p_past
static void ii_aii(SphHandle *h, cudaStream_t st) {
	const SphConsts &c = h->c;
	sph_prof_begin(h, KC_II_AII, st);
	k_ii_rho_adv_aii<<<cdiv(c.N, SPH_BLOCK), SPH_BLOCK, 0, st>>>(c, h->L, rigid_args(h), h->a4[A4_PR], h->a4[A4_VADV], h->bspos,
	                                                            h->a4[A4_FC], h->a4[A4_VEL], h->a1[A1_RHO], h->a1[A1_RHOADV],
	                                                            h->a1[A1_SA], h->a1[A1_P], h->a4[A4_T1], h->ctl);
	sph_prof_end(h, st);
	h->launches += 1;
}

void ii_phase(SphHandle *h, int phase, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nba = cdiv(c.N, SPH_BLOCK);
	if (phase == SPH_PH_II_PREDICT_ADVECTION) {
		first_phase_lists(h, st);
		ii_advect(h, st);
		ii_aii(h, st);
	} else if (phase == SPH_PH_II_PRESSURE_SOLVE) {
		ii_pressure_solve(h, st);
	} else if (phase == SPH_PH_II_ADVECT) { // the step one sweep at a time (single-sweep parity tests)
		ii_advect(h, st);
	} else if (phase == SPH_PH_II_AII) {
		ii_aii(h, st);
	} else if (phase == SPH_PH_II_SOLVE_BEGIN) {
		ii_pressure_solve_begin(h, st);
	} else if (phase == SPH_PH_II_DIJ) {
		ii_dij(h, st);
	} else if (phase == SPH_PH_II_UPDATE) {
		ii_update(h, st);
	} else if (phase == SPH_PH_II_INTEGRATION) {
		if (rigid_args(h).active) { rigid_lists(h, st); rigid_force(h, RF_II, 0, st); } // II:159, gather form
		sph_prof_begin(h, KC_II_INT, st);
		k_ii_integration<<<nba, SPH_BLOCK, 0, st>>>(c, h->fg.sorted_id, h->a4[A4_POS], h->a4[A4_VADV], h->a4[A4_FB],
		                                            h->a4[A4_FC], h->a1[A1_P], h->a4[A4_FD], h->pos, h->vel, h->ctl);
		sph_prof_end(h, st);
		h->launches++;
	}
}

} // namespace SPH_NS
