// sph_ctl.cuh -- the DFSPH solver-loop decisions of the reference's host code, evaluated on the device.
// Mode independent (control arithmetic only); included by sph_sweeps.cu (both arithmetic modes) and by
// sph_multigpu.cu.  On one GPU a one-block controller kernel reduces the sweep's block partials and
// decides; on several GPUs the partials ride with the halo exchange and its receive kernel decides after
// summing the ranks' partials in rank order.  (Deciding in the tail of the sweep itself -- "last block
// done" -- was measured slower: every block then waits ~1 us for its atomic before it can retire, +10 % on
// a 125 us sweep, more than the 6 us controller launch it saves.)
#pragma once
#include "sph_common.cuh"

enum { SPH_CTL_NONE = 0, SPH_CTL_DIV_FIRST, SPH_CTL_DIV_ITER, SPH_CTL_DT, SPH_CTL_DEN, SPH_CTL_PC_FIRST, SPH_CTL_PC_ITER,
       SPH_CTL_II_ITER };

// what a controller needs besides the reduced (sum, count, max)
struct SphCtlArgs {
	float dt_cfl_c1;              // 0.4 * r * 2 (DF:112)
	const SphRigidState *rs;      // DF:104-110: max rigid surface speed
	int rigid_exists;
};

__device__ __forceinline__ void sph_ctl_apply(int kind, SphCtl *ctl, double sum, int cnt, float mx, const SphCtlArgs &a) {
	switch (kind) {
	case SPH_CTL_DIV_FIRST: { // DF:398-399
		float avg = cnt > 0 ? (float)(sum / (double)cnt) : 0.0f; // DF:278-279
		ctl->div_first = avg;
		ctl->div_err = avg;
		ctl->div_past = 0.0f;
		ctl->div_iters = 0;
		ctl->div_active = 1; // iter_cnt < min_iteration_density_divergence
		break;
	}
	case SPH_CTL_DIV_ITER: { // DF:406-414
		if (!ctl->div_active) break;
		float avg = cnt > 0 ? (float)(sum / (double)cnt) : 0.0f;
		ctl->div_past = ctl->div_err;
		ctl->div_err = avg;
		if (fabs((double)avg - (double)ctl->div_past) < 1e-5) { // DF:410-412: break before iter_cnt += 1
			ctl->div_active = 0;
		} else {
			int it = ctl->div_iters + 1;
			ctl->div_iters = it;
			ctl->div_active = ((it < 1 || avg > 10.0f) && it < 15) ? 1 : 0; // DF:400
		}
		break;
	}
	case SPH_CTL_DT: { // DF:100-119: max |v*| (+ rigid surface speed) -> adaptive dt on the device
		float max_rigid_vel = a.rigid_exists ? a.rs->max_surface_vel : 0.0f; // DF:104-110 (loops over ALL rigid particles)
		float max_vel = mx + max_rigid_vel;              // DF:111
		// explicit IEEE division: this header is compiled into translation units with and without -prec-div=false,
		// and the time step must not depend on which of them decides (single GPU: the sweep unit, slabs: the exchange unit)
		float max_dt = __fmul_rn(__fdiv_rn(a.dt_cfl_c1, max_vel), 0.2f);   // DF:112
		float dt;
		if (max_dt > 1e-3f) dt = 1e-3f;                  // DF:114-115
		else dt = fmaxf(max_dt, 1e-5f);                  // DF:117
		ctl->max_vel = max_vel;
		ctl->dt = dt;
		ctl->dt2 = dt * dt;                              // DF:118
		ctl->ps_dt = dt;                                 // DF:119
		ctl->den_active = 1;
		ctl->den_iters = 0;
		ctl->den_avg = INFINITY;
		break;
	}
	case SPH_CTL_DEN: { // DF:221-233, evaluated after compute_all_rho_adv of iteration den_iters
		if (!ctl->den_active) break;
		ctl->den_avg = cnt > 0 ? (float)(sum / (double)cnt) : 1000.0f; // DF:128, 148-149
		break;
	}
	case SPH_CTL_PC_FIRST:  // PC:54: residual before the loop
	case SPH_CTL_PC_ITER: { // PC:68-69: end of a loop body
		if (kind == SPH_CTL_PC_ITER && !ctl->pc_active) break;
		float avg = cnt > 0 ? (float)(sum / (double)cnt) : 0.0f; // PC:131-132
		int it = kind == SPH_CTL_PC_FIRST ? 0 : ctl->pc_iters + 1;
		ctl->pc_iters = it;
		ctl->pc_err = avg;
		ctl->pc_active = (((double)avg > 1000 * .1 * 0.01 || it < 1) && it < 80) ? 1 : 0; // PC:56
		break;
	}
	case SPH_CTL_II_ITER: { // II:78-100 after an iteration
		if (!ctl->ii_active) break;
		float res = cnt > 0 ? (float)(sum / (double)cnt) : 0.0f; // II:111-112
		int l = ctl->ii_iters + 1;                                // II:88
		ctl->ii_iters = l;
		ctl->ii_residual = res;
		if (ctl->ii_have_last && (double)res - (double)ctl->ii_last > 0) { ctl->ii_active = 0; break; } // II:91-93
		ctl->ii_last = res;
		ctl->ii_have_last = 1;
		ctl->ii_active = (((double)res > .1 * 1000 * 0.01 || l < 1) && l < 180) ? 1 : 0; // II:83
		break;
	}
	default: break;
	}
}

// Deterministic reduction of the block partials by ONE block of NT threads (fixed strided order, then a tree).
template <int NT>
__device__ __forceinline__ void sph_reduce_partials(const SphPartial *p, int n, double &sum, int &cnt, float &mx) {
	__shared__ double ss[NT];
	__shared__ int sc[NT];
	__shared__ float sm[NT];
	double a = 0.0;
	int b = 0;
	float m = -INFINITY;
	// eight independent loads in flight per thread (the kernel is pure L2 latency), fixed summation order
	int i = threadIdx.x;
	for (; i + 7 * NT < n; i += 8 * NT) {
		SphPartial q[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) q[u] = p[i + u * NT];
#pragma unroll
		for (int u = 0; u < 8; ++u) { a += q[u].sum; b += q[u].cnt; m = fmaxf(m, q[u].maxv); }
	}
	for (; i < n; i += NT) { a += p[i].sum; b += p[i].cnt; m = fmaxf(m, p[i].maxv); }
	ss[threadIdx.x] = a; sc[threadIdx.x] = b; sm[threadIdx.x] = m;
	__syncthreads();
#pragma unroll
	for (int o = NT / 2; o > 0; o >>= 1) {
		if (threadIdx.x < o) {
			ss[threadIdx.x] += ss[threadIdx.x + o];
			sc[threadIdx.x] += sc[threadIdx.x + o];
			sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
		}
		__syncthreads();
	}
	sum = ss[0]; cnt = sc[0]; mx = sm[0];
}

