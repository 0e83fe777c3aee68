// sph_sweeps_pbf.cuh -- position-based fluids (pbf_solver.py:26-186), included by sph_sweeps.cu (both modes).
//
// The reference's PBF tasks take integer (i, j) indices while for_all_neighbor passes structs (SURVEY
// B-14), so pbf_solver.py cannot compile at the reference's HEAD.  Semantics here are the index-based
// reading fixed by the oracle (oracle/sph_oracle_pbf.inc): fluid and boundary neighbours only, every
// kernel evaluated on the positions of the step's start (PBF:126, 142, 164: pos, never pos_predict), and
// update_all_pos (PBF:67-96) -- one racy loop in the reference -- split into "move all" + "XSPH from the
// moved positions and un-corrected velocities" (the only order a parallel machine can reproduce).
#pragma once

namespace SPH_NS {

#define PBF_EPSILON 1.0e-6f // PBF:16
#define PBF_K 1e-7f         // PBF:18 tension
#define PBF_C 9e-6f         // PBF:19 viscosity

// SB:122-129 poly_kernel(r, h)
__device__ __forceinline__ float poly_w(float r, float h) {
	float q = r / h;
	float q2 = q * q;
	float ret = 0.0f;
	if (q <= 1.0f) {
		float t = 1.0f - q2;
		ret = (315.0f / ((64.0f * PI_F_DEV) * (h * (h * h)))) * (t * (t * t));
	}
	return ret;
}
// PBF:188-196 / SB:113-120 spiky_kernel_derivative(r, h)
__device__ __forceinline__ f3 spiky_dw(const Pair &p, float h) {
	float r_norm = sqrtf(p.r2);
	float q = r_norm / h;
	f3 ret = F3(0.0f, 0.0f, 0.0f);
	if (q <= 1.0f && q > 0.0f) {
		float t = 1.0f - q;
		float h2 = h * h;
		float co = -(45.0f * (t * t));
		float den = (PI_F_DEV * (h2 * h2)) * r_norm;
		ret = (co * p.r) / den;
	}
	return ret;
}
// PBF:164-174: -k (W(r) / W(0.3 h))^4
__device__ __forceinline__ float pbf_s_corr(float r, float h, float r_corr) {
	float s = poly_w(r, h) / poly_w(r_corr, h);
	s *= s;
	s *= s;
	s *= -PBF_K;
	return s;
}

// PBF:26-30 externel_force_predict_pos (acc = gravity, SB:131-134) on the sorted arrays
__global__ void __launch_bounds__(SPH_BLOCK)
k_pbf_predict(SphConsts c, const float4 *__restrict__ spos, float4 *__restrict__ svel, float4 *__restrict__ pos_predict,
              const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float dt = ctl->dt;
	f3 a = F3(c.gravity * 0.0f, c.gravity * -1.0f, c.gravity * 0.0f);
	float4 v4 = svel[s];
	f3 v = xyz(v4) + dt * a;
	svel[s] = F4(v, v4.w);
	pos_predict[s] = F4(xyz(spos[s]) + dt * v, 0.0f);
}

// PBF:32-52 compute_all_lambda: rho (SB:41-51 with the poly6 tasks PBF:154-162), constrain (PBF:137),
// constrain_derivative (PBF:114-131), sum of squared derivatives (PBF:139-150), lambda -- one pass over the
// lists with one accumulator per reference sum (each keeps the reference's order).  posT1.w = lambda.
__global__ void __launch_bounds__(SPH_BLOCK)
k_pbf_lambda(SphConsts c, SphLists L, const float4 *__restrict__ spos, const float4 *__restrict__ bspos,
             float *__restrict__ rho, float *__restrict__ constrain, float *__restrict__ lambda,
             float4 *__restrict__ cd_out, float4 *__restrict__ posT1) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pi = spos[s];
	float rho_f = 0.001f; // SB:44
	f3 cd = F3(0.0f, 0.0f, 0.0f);
	float sum = 0.0f;
	SPH_FOR_FLUID(L, c, s, j) {
		if (j & SPH_RIGID_BIT) return; // index-based tasks cannot address rigid particles
		Pair p = make_pair(pi, __ldg(&spos[j]));
		rho_f += c.m * poly_w(sqrtf(p.r2), c.h); // PBF:155
		f3 g = spiky_dw(p, c.h) / 1000.0f;       // PBF:126, 142
		cd = cd + g;
		sum += dot(g, g);
	};
	float rho_i = rho_f;
	float sum_b = 0.0f;
	if (c.boundary_handle == 1) {
		float rho_b = 0.0f;
		f3 cdb = F3(0.0f, 0.0f, 0.0f);
		SPH_FOR_BOUNDARY(L, c, s, j) {
			float4 pj = __ldg(&bspos[j]);
			Pair p = make_pair(pi, pj);
			rho_b += pj.w * poly_w(sqrtf(p.r2), c.h); // PBF:159-161
			f3 g = spiky_dw(p, c.h) / 1000.0f;        // PBF:130-131, 148
			cdb = cdb + g;
			sum_b += dot(g, g);
		};
		rho_i = rho_f + rho_b * SPH_RHO0; // SB:49
		cd = cd + cdb;                    // PBF:121
	}
	float con = fmaxf(rho_i / 1000.0f - 1.0f, 0.0f); // PBF:137
	float lam = 0.0f;
	if (!(con == 0.0f)) { // PBF:39-52
		float den = c.boundary_handle == 1 ? (dot(cd, cd) + sum) + sum_b : dot(cd, cd) + sum;
		lam = (-con) / (den + PBF_EPSILON);
	}
	rho[s] = rho_i;
	constrain[s] = con;
	lambda[s] = lam;
	cd_out[s] = F4(cd, 0.0f);
	posT1[s] = make_float4(pi.x, pi.y, pi.z, lam);
}

// PBF:55-65 compute_all_delta_pos
__global__ void __launch_bounds__(SPH_BLOCK)
k_pbf_delta_pos(SphConsts c, SphLists L, const float4 *__restrict__ posT1, const float4 *__restrict__ bspos,
                float r_corr, float4 *__restrict__ delta_pos) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pi = posT1[s];
	f3 dp = F3(0.0f, 0.0f, 0.0f);
	SPH_FOR_FLUID(L, c, s, j) {
		if (j & SPH_RIGID_BIT) return;
		float4 pj = __ldg(&posT1[j]);
		Pair p = make_pair(pi, pj);
		float sc = pbf_s_corr(sqrtf(p.r2), c.h, r_corr);
		dp = dp + ((pi.w + pj.w) + sc) * spiky_dw(p, c.h); // PBF:164
	};
	if (c.boundary_handle == 1) {
		f3 dpb = F3(0.0f, 0.0f, 0.0f);
		SPH_FOR_BOUNDARY(L, c, s, j) {
			Pair p = make_pair(pi, __ldg(&bspos[j]));
			float sc = pbf_s_corr(sqrtf(p.r2), c.h, r_corr);
			dpb = dpb + (pi.w + sc) * spiky_dw(p, c.h); // PBF:174
		};
		delta_pos[s] = F4((dp + dpb) / 1000.0f, 0.0f); // PBF:62
	} else {
		delta_pos[s] = F4(dp / 1000.0f, 0.0f); // PBF:64
	}
}

// PBF:69-84: move every particle (constraint correction, velocity from the displacement, clamp boundary)
__global__ void __launch_bounds__(SPH_BLOCK)
k_pbf_move(SphConsts c, const float4 *__restrict__ spos, const float4 *__restrict__ pos_predict,
           const float4 *__restrict__ delta_pos, float4 *__restrict__ new_pos, float4 *__restrict__ svel,
           float lo0, float lo1, float lo2, float hi0, float hi1, float hi2, const SphCtl *__restrict__ ctl) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float dt = ctl->dt;
	f3 pp = xyz(pos_predict[s]) + xyz(delta_pos[s]);
	f3 v = (pp - xyz(spos[s])) / dt;
	if (c.boundary_handle == 0) { // PBF:73-80: margin = particle_radius, velocity decays (no reflection)
		float lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
		float *x = &pp.x, *vv = &v.x;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			if (x[k] <= lo[k]) { x[k] = lo[k]; vv[k] *= 0.5f; }
			if (x[k] >= hi[k]) { x[k] = hi[k]; vv[k] *= 0.5f; }
		}
	}
	new_pos[s] = F4(pp, 0.0f);
	svel[s] = F4(v, svel[s].w);
}

// PBF:87-96: XSPH viscosity from the moved positions.  The neighbour candidates are the step's grid
// (cells of the OLD positions, PS:452), the cull (PS:466) is on the NEW positions, so this is a 27-cell
// traversal of its own; fused with the write-back into the caller's original-order state.
__global__ void __launch_bounds__(SPH_BLOCK)
k_pbf_xsph(SphConsts c, const int *__restrict__ scell, const int *__restrict__ cstart, const int *__restrict__ sorted_id,
           const float4 *__restrict__ new_pos, const float4 *__restrict__ svel, float4 *__restrict__ pos,
           float4 *__restrict__ vel) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= c.N) return;
	float4 pi = new_pos[s];
	float4 vi4 = svel[s];
	f3 vi = xyz(vi4);
	int i = sorted_id[s];
	if (i >= c.N_owned) return; // ghost copy (multi-GPU slabs): moved and corrected on its owner's rank
	int cx, cy, cz;
	cell_xyz(scell[s], c, cx, cy, cz);
	f3 acc = F3(0.0f, 0.0f, 0.0f);
	SPH_FOR_27(c, cx, cy, cz, c1) {
		int a = cstart[c1], b = cstart[c1 + 1];
		for (int e = a; e < b; ++e) {
			if (e == s) continue; // PS:461
			Pair p = make_pair(pi, new_pos[e]);
			if (culled(p, c)) continue; // PS:466
			acc = acc + (xyz(svel[e]) - vi) * poly_w(sqrtf(p.r2), c.h); // PBF:99
		}
	}
	pos[i] = make_float4(pi.x, pi.y, pi.z, 0.0f);   // PBF:84
	vel[i] = F4(vi + PBF_C * acc, vi4.w);           // PBF:94-96 (the boundary sum PBF:91-92 is computed but unused)
}

void pbf_phase(SphHandle *h, int phase, cudaStream_t st) {
	const SphConsts &c = h->c;
	int nb = cdiv(c.N, SPH_BLOCK);
	float4 *pos_predict = h->a4[A4_T2], *new_pos = h->a4[A4_T3];
	if (phase == SPH_PH_PBF_PREDICT) {
		sph_prof_begin(h, KC_OTHER, st);
		k_pbf_predict<<<nb, SPH_BLOCK, 0, st>>>(c, h->a4[A4_POS], h->a4[A4_VEL], pos_predict, h->ctl);
		sph_prof_end(h, st);
		h->launches++;
	} else if (phase == SPH_PH_PBF_LAMBDA) {
		build_lists(h, st);
		sph_prof_begin(h, KC_OTHER, st);
		k_pbf_lambda<<<nb, SPH_BLOCK, 0, st>>>(c, h->L, h->a4[A4_POS], h->bspos, h->a1[A1_RHO], h->a1[A1_SA], h->a1[A1_SB],
		                                       h->a4[A4_FA], h->a4[A4_T1]);
		sph_prof_end(h, st);
		h->launches++;
		mg_exchange(h, MG_F4_T1W, st); // slabs: lambda of the ghost particles
	} else if (phase == SPH_PH_PBF_DELTA_POS) {
		float r_corr = (float)(0.3 * (4.0 * h->cfg.particle_radius)); // PBF:20, 166: s_corr_factor * kernel_h in Python scope
		sph_prof_begin(h, KC_OTHER, st);
		k_pbf_delta_pos<<<nb, SPH_BLOCK, 0, st>>>(c, h->L, h->a4[A4_T1], h->bspos, r_corr, h->a4[A4_FB]);
		sph_prof_end(h, st);
		h->launches++;
	} else if (phase == SPH_PH_PBF_UPDATE_POS) {
		float lo[3], hi[3];
		for (int k = 0; k < 3; ++k) {
			lo[k] = (float)(h->cfg.box_min[k] + h->cfg.particle_radius);
			hi[k] = (float)(h->cfg.box_max[k] - h->cfg.particle_radius);
		}
		sph_prof_begin(h, KC_OTHER, st);
		k_pbf_move<<<nb, SPH_BLOCK, 0, st>>>(c, h->a4[A4_POS], pos_predict, h->a4[A4_FB], new_pos, h->a4[A4_VEL], lo[0], lo[1],
		                                     lo[2], hi[0], hi[1], hi[2], h->ctl);
		mg_exchange(h, MG_XYZ(A4_T3), st); // slabs: moved positions and un-corrected velocities of the ghost particles
		mg_exchange(h, MG_F4_VEL, st);
		k_pbf_xsph<<<nb, SPH_BLOCK, 0, st>>>(c, h->fg.scell, h->fg.cell_start, h->fg.sorted_id, new_pos, h->a4[A4_VEL], h->pos,
		                                     h->vel);
		sph_prof_end(h, st);
		h->launches += 2;
	}
}

} // namespace SPH_NS
