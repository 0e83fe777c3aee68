// sph_multigpu.cu -- x-slab domain decomposition over the GPUs of one box.
//
// Rank p owns the x-cell columns [col_lo, col_hi) of the GLOBAL grid (cell ids stay the reference's,
// PS:102, so integer outputs equal the single-domain run).  NCCL over NVLink is used for exactly three
// things: (1) particle migration after advection, (2) the one-column ghost (halo) layer -- particle
// copies once per step, then one float4 per ghost after every sweep that produces a quantity a
// neighbour reads, (3) the (sum, count) / max all-reduce behind each solver-loop decision, so that every
// rank executes the single-domain iteration counts.  Nothing here synchronises the host except the one
// count read-back per step in mg_begin_step.  NCCL is bound at run time with dlopen (the torch-bundled
// libnccl.so.2 that the process already maps); a missing library is an error, never a fallback.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "sph_internal.h"

struct NcclApi {
	void *lib;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {};

static int nccl_load(SphHandle *h) {
	if (g_nccl.lib) return SPH_OK;
	void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) return sph_fail(h, SPH_ESTATE, "multi-GPU: cannot dlopen libnccl.so.2 (%s)", dlerror());
#define L(sym) *(void **)(&g_nccl.sym) = dlsym(lib, "nccl" #sym); if (!g_nccl.sym) return sph_fail(h, SPH_ESTATE, "multi-GPU: nccl" #sym " not found")
	L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(GroupStart); L(GroupEnd); L(Send); L(Recv); L(AllReduce); L(GetErrorString);
#undef L
	g_nccl.lib = lib;
	return SPH_OK;
}

// counters kept on the device (read back once per step)
enum { MC_KEEP = 0, MC_MIG_L, MC_MIG_R, MC_HALO_L, MC_HALO_R, MC_OWNED, MC_GHOST_L, MC_GHOST_R, MC_OVERFLOW, MC_COUNT };

struct SphComm {
	ncclComm_t comm;
	int rank, nranks;
	int col_lo, col_hi;
	int cap_halo, cap_mig; // particles per message
	// particle messages: [int4 header (count)] [float4 pos * cap] [float4 vel * cap] [int gid * cap]
	char *msg_send[2], *msg_recv[2]; // [0] = left neighbour, [1] = right neighbour
	size_t msg_bytes_halo, msg_bytes_mig;
	int *counters;          // device, MC_COUNT ints
	int *counters_host;     // pinned
	float4 *tmp_pos, *tmp_vel;
	int *tmp_gid;
	int *send_orig[2];      // original (local) index of the k-th particle packed for each side
	int n_send[2], n_recv[2];
	float4 *xsend[2], *xrecv[2]; // per-sweep ghost values
};

#define NCCL_OK(h, expr)                                                                                 \
	do {                                                                                                 \
		ncclResult_t r_ = (expr);                                                                        \
		if (r_ != ncclSuccess) return sph_fail((h), SPH_ECUDA, "NCCL: %s (%s)", g_nccl.GetErrorString(r_), #expr); \
	} while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t msg_size(int cap) { return 16 + (size_t)cap * (16 + 16 + 4); }
__host__ __device__ static inline float4 *msg_pos(char *m) { return (float4 *)(m + 16); }
__host__ __device__ static inline float4 *msg_vel(char *m, int cap) { return (float4 *)(m + 16 + (size_t)cap * 16); }
__host__ __device__ static inline int *msg_gid(char *m, int cap) { return (int *)(m + 16 + (size_t)cap * 32); }

// ---- kernels -------------------------------------------------------------------------------------
__device__ __forceinline__ int column_of(float x, float h) { return (int)floorf(__fdiv_rn(x, h)); } // PS:494

// migration: classify the owned particles by x-column (after the previous step's advection)
__global__ void __launch_bounds__(256)
k_mg_classify(const float4 *__restrict__ pos, const float4 *__restrict__ vel, const int *__restrict__ gid, int n,
              float h, int col_lo, int col_hi, int has_left, int has_right, float4 *__restrict__ keep_pos,
              float4 *__restrict__ keep_vel, int *__restrict__ keep_gid, char *msg_l, char *msg_r, int cap,
              int *counters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int dest = -1; // 0 keep, 1 left, 2 right
	float4 p, v;
	int g = 0;
	if (i < n) {
		p = pos[i]; v = vel[i]; g = gid[i];
		int col = column_of(p.x, h);
		dest = (col < col_lo && has_left) ? 1 : ((col >= col_hi && has_right) ? 2 : 0);
	}
	// warp-aggregated append: one atomic per destination per warp
	for (int d = 0; d < 3; ++d) {
		unsigned m = __ballot_sync(0xffffffffu, dest == d);
		if (!m) continue;
		int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
		if (lane == leader) base = atomicAdd(&counters[d == 0 ? MC_KEEP : (d == 1 ? MC_MIG_L : MC_MIG_R)], __popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (dest == d) {
			int k = base + __popc(m & ((1u << lane) - 1));
			if (d == 0) { keep_pos[k] = p; keep_vel[k] = v; keep_gid[k] = g; }
			else {
				char *msg = d == 1 ? msg_l : msg_r;
				if (k < cap) { msg_pos(msg)[k] = p; msg_vel(msg, cap)[k] = v; msg_gid(msg, cap)[k] = g; }
				else atomicOr(&counters[MC_OVERFLOW], 1);
			}
		}
	}
}

__global__ void k_mg_header(char *msg_l, char *msg_r, const int *counters, int idx_l, int idx_r, int cap) {
	if (threadIdx.x == 0) {
		((int *)msg_l)[0] = min(counters[idx_l], cap);
		((int *)msg_r)[0] = min(counters[idx_r], cap);
	}
}

// kept particles back into the bound arrays, then the received migrants behind them
__global__ void __launch_bounds__(256)
k_mg_rebuild(const float4 *__restrict__ keep_pos, const float4 *__restrict__ keep_vel, const int *__restrict__ keep_gid,
             const char *rl, const char *rr, int has_left, int has_right, int cap, float4 *__restrict__ pos,
             float4 *__restrict__ vel, int *__restrict__ gid, int capacity, int *counters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int keep = counters[MC_KEEP];
	int nl = has_left ? ((const int *)rl)[0] : 0;
	int nr = has_right ? ((const int *)rr)[0] : 0;
	int total = keep + nl + nr;
	if (i == 0) {
		counters[MC_OWNED] = min(total, capacity);
		if (total > capacity) atomicOr(&counters[MC_OVERFLOW], 2);
	}
	if (i >= total || i >= capacity) return;
	if (i < keep) { pos[i] = keep_pos[i]; vel[i] = keep_vel[i]; gid[i] = keep_gid[i]; }
	else if (i < keep + nl) {
		int k = i - keep;
		pos[i] = msg_pos((char *)rl)[k]; vel[i] = msg_vel((char *)rl, cap)[k]; gid[i] = msg_gid((char *)rl, cap)[k];
	} else {
		int k = i - keep - nl;
		pos[i] = msg_pos((char *)rr)[k]; vel[i] = msg_vel((char *)rr, cap)[k]; gid[i] = msg_gid((char *)rr, cap)[k];
	}
}

// halo: owned particles of the first / last owned column are copied to the left / right neighbour
__global__ void __launch_bounds__(256)
k_mg_pack_halo(const float4 *__restrict__ pos, const float4 *__restrict__ vel, const int *__restrict__ gid,
               float h, int col_lo, int col_hi, int has_left, int has_right, char *msg_l, char *msg_r, int cap,
               int *__restrict__ send_orig_l, int *__restrict__ send_orig_r, int *counters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int n = counters[MC_OWNED];
	bool tl = false, tr = false;
	float4 p, v;
	int g = 0;
	if (i < n) {
		p = pos[i]; v = vel[i]; g = gid[i];
		int col = column_of(p.x, h);
		tl = has_left && col == col_lo;
		tr = has_right && col == col_hi - 1;
	}
	for (int d = 0; d < 2; ++d) {
		bool t = d == 0 ? tl : tr;
		unsigned m = __ballot_sync(0xffffffffu, t);
		if (!m) continue;
		int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
		if (lane == leader) base = atomicAdd(&counters[d == 0 ? MC_HALO_L : MC_HALO_R], __popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (t) {
			int k = base + __popc(m & ((1u << lane) - 1));
			char *msg = d == 0 ? msg_l : msg_r;
			if (k < cap) {
				msg_pos(msg)[k] = p; msg_vel(msg, cap)[k] = v; msg_gid(msg, cap)[k] = g;
				(d == 0 ? send_orig_l : send_orig_r)[k] = i;
			} else atomicOr(&counters[MC_OVERFLOW], 4);
		}
	}
}

__global__ void __launch_bounds__(256)
k_mg_unpack_halo(const char *rl, const char *rr, int has_left, int has_right, int cap, float4 *__restrict__ pos,
                 float4 *__restrict__ vel, int *__restrict__ gid, int capacity, int *counters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int n = counters[MC_OWNED];
	int nl = has_left ? ((const int *)rl)[0] : 0;
	int nr = has_right ? ((const int *)rr)[0] : 0;
	if (n + nl + nr > capacity) { // keep what fits, flag the rest
		if (i == 0) atomicOr(&counters[MC_OVERFLOW], 8);
		nr = max(0, min(nr, capacity - n - nl));
		nl = max(0, min(nl, capacity - n));
	}
	if (i == 0) { counters[MC_GHOST_L] = nl; counters[MC_GHOST_R] = nr; }
	if (i < nl) {
		pos[n + i] = msg_pos((char *)rl)[i]; vel[n + i] = msg_vel((char *)rl, cap)[i]; gid[n + i] = msg_gid((char *)rl, cap)[i];
	} else if (i < nl + nr) {
		int k = i - nl;
		pos[n + i] = msg_pos((char *)rr)[k]; vel[n + i] = msg_vel((char *)rr, cap)[k]; gid[n + i] = msg_gid((char *)rr, cap)[k];
	}
}

// per-sweep ghost values: one float4 per sent particle, both neighbours in one launch
__global__ void __launch_bounds__(256)
k_mg_pack_values(int what, const int *__restrict__ send_orig_l, const int *__restrict__ send_orig_r,
                 const int *__restrict__ slot_of, int nl, int nr, const float4 *__restrict__ a,
                 const float4 *__restrict__ b, float4 *__restrict__ out_l, float4 *__restrict__ out_r) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nl + nr) return;
	bool left = k < nl;
	int kk = left ? k : k - nl;
	int s = slot_of[(left ? send_orig_l : send_orig_r)[kk]];
	float4 v;
	if (what == MG_F4_T1R) v = make_float4(a[s].w, b[s].w, 0.0f, 0.0f); // posT1.w, posR.w
	else if (what == MG_F4_T2 || what == MG_F4_T3) v = make_float4(a[s].w, 0.0f, 0.0f, 0.0f);
	else v = a[s];
	(left ? out_l : out_r)[kk] = v;
}
__global__ void __launch_bounds__(256)
k_mg_unpack_values(int what, const int *__restrict__ slot_of, int first_orig, int nl, int nr,
                   const float4 *__restrict__ in_l, const float4 *__restrict__ in_r,
                   const float4 *__restrict__ spos, float4 *__restrict__ a, float4 *__restrict__ b,
                   float4 *__restrict__ pv) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nl + nr) return;
	int s = slot_of[first_orig + k]; // ghosts are stored left block first, then right block
	float4 v = k < nl ? in_l[k] : in_r[k - nl];
	float4 p = spos[s]; // the payload buffers carry a position copy; ghosts get theirs here
	if (what == MG_F4_T1R) { a[s] = make_float4(p.x, p.y, p.z, v.x); b[s] = make_float4(p.x, p.y, p.z, v.y); }
	else if (what == MG_F4_T2 || what == MG_F4_T3) a[s] = make_float4(p.x, p.y, p.z, v.x);
	else {
		a[s] = v;
		if (what == MG_F4_VEL) pv[2 * (size_t)s + 1] = v; // the 256-bit (pos, vel) records of k_df_drho
	}
}

// ---- host side ------------------------------------------------------------------------------------
extern "C" int sph_comm_unique_id(char *out128) {
	if (!out128) return SPH_EINVAL;
	int rc = nccl_load(nullptr);
	if (rc != SPH_OK) return rc;
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) return sph_fail(nullptr, SPH_ECUDA, "ncclGetUniqueId failed");
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
	memcpy(out128, &id, 128);
	return SPH_OK;
}

extern "C" int sph_comm_init(SphHandle *h, const char *id128, int rank, int nranks, int col_lo, int col_hi) {
	if (!h || !id128) return SPH_EINVAL;
	if (nranks < 2 || rank < 0 || rank >= nranks) return sph_fail(h, SPH_EINVAL, "sph_comm_init: bad rank %d / %d", rank, nranks);
	if (h->c.solver != SPH_SOLVER_DFSPH) return sph_fail(h, SPH_EINVAL, "multi-GPU slabs are built for the DFSPH solver");
	if (h->cfg.n_ghost_capacity <= 0) return sph_fail(h, SPH_EINVAL, "sph_comm_init: create the handle with n_ghost_capacity > 0");
	if (!h->gid) return sph_fail(h, SPH_ENOTBOUND, "sph_comm_init: bind SPH_F_FLUID_GID first");
	int rc = nccl_load(h);
	if (rc != SPH_OK) return rc;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	SphComm *m = new (std::nothrow) SphComm();
	if (!m) return sph_fail(h, SPH_ENOMEM, "sph_comm_init: out of memory");
	memset(m, 0, sizeof(*m));
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	ncclResult_t r = g_nccl.CommInitRank(&m->comm, nranks, id, rank);
	if (r != ncclSuccess) { delete m; return sph_fail(h, SPH_ECUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
	m->rank = rank; m->nranks = nranks; m->col_lo = col_lo; m->col_hi = col_hi;
	m->cap_halo = h->cfg.n_ghost_capacity / 2;
	m->cap_mig = m->cap_halo / 2 > 0 ? m->cap_halo / 2 : 1;
	m->msg_bytes_halo = msg_size(m->cap_halo);
	m->msg_bytes_mig = msg_size(m->cap_mig);
	size_t ncap = (size_t)h->cfg.n_fluid + (size_t)h->cfg.n_ghost_capacity;
	for (int d = 0; d < 2; ++d) {
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->msg_send[d], m->msg_bytes_halo));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->msg_recv[d], m->msg_bytes_halo));
		SPH_CUDA_CHECK(h, cudaMemset(m->msg_send[d], 0, m->msg_bytes_halo));
		SPH_CUDA_CHECK(h, cudaMemset(m->msg_recv[d], 0, m->msg_bytes_halo));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->send_orig[d], sizeof(int) * (size_t)m->cap_halo));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->xsend[d], sizeof(float4) * (size_t)m->cap_halo));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->xrecv[d], sizeof(float4) * (size_t)m->cap_halo));
	}
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->counters, sizeof(int) * MC_COUNT));
	SPH_CUDA_CHECK(h, cudaMallocHost((void **)&m->counters_host, sizeof(int) * MC_COUNT));
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->tmp_pos, sizeof(float4) * ncap));
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->tmp_vel, sizeof(float4) * ncap));
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->tmp_gid, sizeof(int) * ncap));
	h->comm = m;
	return SPH_OK;
}

void mg_destroy(SphHandle *h) {
	SphComm *m = h->comm;
	if (!m) return;
	for (int d = 0; d < 2; ++d) {
		cudaFree(m->msg_send[d]); cudaFree(m->msg_recv[d]); cudaFree(m->send_orig[d]); cudaFree(m->xsend[d]); cudaFree(m->xrecv[d]);
	}
	cudaFree(m->counters); cudaFreeHost(m->counters_host); cudaFree(m->tmp_pos); cudaFree(m->tmp_vel); cudaFree(m->tmp_gid);
	if (m->comm) g_nccl.CommDestroy(m->comm);
	delete m;
	h->comm = nullptr;
}

static int exchange_bytes(SphHandle *h, SphComm *m, size_t bytes, cudaStream_t st) {
	int left = m->rank - 1, right = m->rank + 1;
	NCCL_OK(h, g_nccl.GroupStart());
	if (left >= 0) {
		NCCL_OK(h, g_nccl.Send(m->msg_send[0], bytes, ncclChar, left, m->comm, st));
		NCCL_OK(h, g_nccl.Recv(m->msg_recv[0], bytes, ncclChar, left, m->comm, st));
	}
	if (right < m->nranks) {
		NCCL_OK(h, g_nccl.Send(m->msg_send[1], bytes, ncclChar, right, m->comm, st));
		NCCL_OK(h, g_nccl.Recv(m->msg_recv[1], bytes, ncclChar, right, m->comm, st));
	}
	NCCL_OK(h, g_nccl.GroupEnd());
	return SPH_OK;
}

// migration + ghost particle exchange + one read-back of the new counts (sets c.N_owned / c.N)
int mg_begin_step(SphHandle *h, cudaStream_t st) {
	SphComm *m = h->comm;
	if (!m) return SPH_OK;
	const SphConsts &c = h->c;
	int has_left = m->rank > 0, has_right = m->rank + 1 < m->nranks;
	int capacity_owned = h->cfg.n_fluid;
	int capacity_all = h->cfg.n_fluid + h->cfg.n_ghost_capacity;
	int n = c.N_owned;
	cudaMemsetAsync(m->counters, 0, sizeof(int) * MC_COUNT, st);
	// (1) migration
	k_mg_classify<<<cdiv(n > 0 ? n : 1, 256), 256, 0, st>>>(h->pos, h->vel, h->gid, n, c.h, m->col_lo, m->col_hi, has_left,
	                                                        has_right, m->tmp_pos, m->tmp_vel, m->tmp_gid, m->msg_send[0],
	                                                        m->msg_send[1], m->cap_mig, m->counters);
	k_mg_header<<<1, 32, 0, st>>>(m->msg_send[0], m->msg_send[1], m->counters, MC_MIG_L, MC_MIG_R, m->cap_mig);
	int rc = exchange_bytes(h, m, m->msg_bytes_mig, st);
	if (rc != SPH_OK) return rc;
	k_mg_rebuild<<<cdiv(capacity_owned, 256), 256, 0, st>>>(m->tmp_pos, m->tmp_vel, m->tmp_gid, m->msg_recv[0], m->msg_recv[1],
	                                                        has_left, has_right, m->cap_mig, h->pos, h->vel, h->gid,
	                                                        capacity_owned, m->counters);
	// (2) halo particles
	k_mg_pack_halo<<<cdiv(capacity_owned, 256), 256, 0, st>>>(h->pos, h->vel, h->gid, c.h, m->col_lo, m->col_hi, has_left,
	                                                          has_right, m->msg_send[0], m->msg_send[1], m->cap_halo,
	                                                          m->send_orig[0], m->send_orig[1], m->counters);
	k_mg_header<<<1, 32, 0, st>>>(m->msg_send[0], m->msg_send[1], m->counters, MC_HALO_L, MC_HALO_R, m->cap_halo);
	rc = exchange_bytes(h, m, m->msg_bytes_halo, st);
	if (rc != SPH_OK) return rc;
	k_mg_unpack_halo<<<cdiv(2 * m->cap_halo, 256), 256, 0, st>>>(m->msg_recv[0], m->msg_recv[1], has_left, has_right,
	                                                             m->cap_halo, h->pos, h->vel, h->gid, capacity_all, m->counters);
	h->launches += 6;
	// (3) the one host read-back of the step
	SPH_CUDA_CHECK(h, cudaMemcpyAsync(m->counters_host, m->counters, sizeof(int) * MC_COUNT, cudaMemcpyDeviceToHost, st));
	SPH_CUDA_CHECK(h, cudaStreamSynchronize(st));
	const int *k = m->counters_host;
	if (k[MC_OVERFLOW]) return sph_fail(h, SPH_ESTATE, "multi-GPU: message or slab capacity exceeded (flags %d)", k[MC_OVERFLOW]);
	h->c.N_owned = k[MC_OWNED];
	h->c.N = k[MC_OWNED] + k[MC_GHOST_L] + k[MC_GHOST_R];
	m->n_send[0] = k[MC_HALO_L] < m->cap_halo ? k[MC_HALO_L] : m->cap_halo;
	m->n_send[1] = k[MC_HALO_R] < m->cap_halo ? k[MC_HALO_R] : m->cap_halo;
	m->n_recv[0] = k[MC_GHOST_L];
	m->n_recv[1] = k[MC_GHOST_R];
	return SPH_OK;
}

void mg_after_grid(SphHandle *h, cudaStream_t st) { (void)h; (void)st; }

void sph_reduce_partials_launch(SphHandle *h, int n_blocks, cudaStream_t st); // sph_sweeps.cu

// Ghost values of one field to / from both neighbours; when reduce_blocks > 0 the loop-decision
// all-reduce of the same sweep's block partials rides in the same NCCL group (one launch of the
// communication kernel instead of two).
static void mg_exchange_impl(SphHandle *h, int what, int reduce_blocks, cudaStream_t st) {
	SphComm *m = h->comm;
	if (!m) return;
	const float4 *a = nullptr, *b = nullptr;
	float4 *wa = nullptr, *wb = nullptr;
	switch (what) {
	case MG_F4_T1R: a = wa = h->a4[A4_T1]; b = wb = h->a4[A4_PR]; break;
	case MG_F4_VEL: a = wa = h->a4[A4_VEL]; break;
	case MG_F4_T2: a = wa = h->a4[A4_T2]; break;
	case MG_F4_VADV: a = wa = h->a4[A4_VADV]; break;
	case MG_F4_T3: a = wa = h->a4[A4_T3]; break;
	default: break;
	}
	int ns = m->n_send[0] + m->n_send[1], nr = m->n_recv[0] + m->n_recv[1];
	if (a && ns > 0) {
		k_mg_pack_values<<<cdiv(ns, 256), 256, 0, st>>>(what, m->send_orig[0], m->send_orig[1], h->fg.slot_of, m->n_send[0],
		                                                m->n_send[1], a, b, m->xsend[0], m->xsend[1]);
		h->launches++;
	}
	if (reduce_blocks > 0) sph_reduce_partials_launch(h, reduce_blocks, st);
	int left = m->rank - 1, right = m->rank + 1;
	g_nccl.GroupStart();
	if (a && left >= 0) {
		if (m->n_send[0] > 0) g_nccl.Send(m->xsend[0], (size_t)m->n_send[0] * 4, ncclFloat, left, m->comm, st);
		if (m->n_recv[0] > 0) g_nccl.Recv(m->xrecv[0], (size_t)m->n_recv[0] * 4, ncclFloat, left, m->comm, st);
	}
	if (a && right < m->nranks) {
		if (m->n_send[1] > 0) g_nccl.Send(m->xsend[1], (size_t)m->n_send[1] * 4, ncclFloat, right, m->comm, st);
		if (m->n_recv[1] > 0) g_nccl.Recv(m->xrecv[1], (size_t)m->n_recv[1] * 4, ncclFloat, right, m->comm, st);
	}
	if (reduce_blocks > 0) {
		g_nccl.AllReduce(h->red, h->red, 2, ncclDouble, ncclSum, m->comm, st);
		g_nccl.AllReduce(h->red + 2, h->red + 2, 1, ncclDouble, ncclMax, m->comm, st);
	}
	g_nccl.GroupEnd();
	if (a && nr > 0) {
		k_mg_unpack_values<<<cdiv(nr, 256), 256, 0, st>>>(what, h->fg.slot_of, h->c.N_owned, m->n_recv[0], m->n_recv[1],
		                                                  m->xrecv[0], m->xrecv[1], h->a4[A4_POS], wa, wb, h->pv);
		h->launches++;
	}
}

void mg_exchange(SphHandle *h, int what, cudaStream_t st) { mg_exchange_impl(h, what, 0, st); }
void mg_allreduce(SphHandle *h, int n_blocks, cudaStream_t st) { mg_exchange_impl(h, -1, n_blocks, st); }
void mg_exchange_reduce(SphHandle *h, int what, int n_blocks, cudaStream_t st) { mg_exchange_impl(h, what, n_blocks, st); }

extern "C" int sph_comm_info(SphHandle *h, int32_t *out8) {
	if (!h || !out8) return SPH_EINVAL;
	memset(out8, 0, sizeof(int32_t) * 8);
	out8[0] = h->c.N_owned;
	out8[1] = h->c.N - h->c.N_owned;
	if (h->comm) {
		out8[2] = h->comm->n_send[0]; out8[3] = h->comm->n_send[1];
		out8[4] = h->comm->n_recv[0]; out8[5] = h->comm->n_recv[1];
		out8[6] = h->comm->rank; out8[7] = h->comm->nranks;
	}
	return SPH_OK;
}
