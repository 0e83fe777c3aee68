// sph_multigpu.cu -- x-slab domain decomposition over the GPUs of one box.
//
// Rank p owns the x-cell columns [col_lo, col_hi) of the GLOBAL grid (cell ids stay the reference's,
// PS:102, so integer outputs equal the single-domain run).  Exactly three things cross GPUs over NVLink:
// (1) particle migration after advection, (2) the one-column ghost (halo) layer -- particle
// copies once per step, then one float4 per ghost after every sweep that produces a quantity a
// neighbour reads, (3) the (sum, count) / max all-reduce behind each solver-loop decision, so that every
// rank executes the single-domain iteration counts.  Nothing here synchronises the host except the one
// count read-back per step in mg_begin_step.  NCCL (start-up all-gather of the IPC handles, the rigid-body
// all-reduce, and the optional SPH_MG_TRANSPORT=nccl transport) is bound at run time with dlopen (the torch-bundled
// libnccl.so.2 that the process already maps); a missing library is an error, never a fallback.
//
// Transport of (1), (2) and (3) inside a step: the ~40 exchanges per DFSPH step are 0.3 MB each, i.e. pure
// latency, so they do not go through NCCL.  Every rank owns a "window" of device memory that its peers
// map with CUDA IPC.  One kernel per exchange (k_mg_exchange): each block stores this rank's ghost values
// straight into the neighbours' windows over NVLink, then polls the slots its neighbours fill and scatters
// them into the ghost slots; one extra block reduces the sweep's block partials, stores them into every
// rank's window, polls all ranks' slots, sums them in rank order (deterministic, identical on every rank)
// and takes the loop decision (sph_ctl.cuh).  Every 16-byte slot carries the exchange's epoch in its last
// word and is written with one 128-bit store (the atomicity NCCL's LL protocols rely on), so there is no
// fence, no flag, no host involvement and no proxy thread.  Windows are double-buffered by epoch parity; every
// exchange carries a tagged sync slot to and from both neighbours (whatever the halo counts), so a rank can be
// at most one exchange ahead of a peer.  The two variable-size particle messages per step (migration, ghost particles)
// use the same windows (k_mg_xfer).  SPH_MG_TRANSPORT=nccl selects NCCL for everything (A/B measurements).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "sph_internal.h"
#include "sph_ctl.cuh"
#include "sph_mgwin.cuh"

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
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {};

static int nccl_load(SphHandle *h) {
	if (g_nccl.lib) return SPH_OK;
	void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) return sph_fail(h, SPH_ESTATE, "multi-GPU: cannot dlopen libnccl.so.2 (%s)", dlerror());
#define L(sym) *(void **)(&g_nccl.sym) = dlsym(lib, "nccl" #sym); if (!g_nccl.sym) return sph_fail(h, SPH_ESTATE, "multi-GPU: nccl" #sym " not found")
	L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(GroupStart); L(GroupEnd); L(Send); L(Recv); L(AllReduce); L(AllGather); L(GetErrorString);
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
	float4 *xsend[2], *xrecv[2]; // per-sweep ghost values (NCCL transport)
	// peer-memory transport (CUDA IPC windows over NVLink)
	int p2p;                     // 1: per-sweep values and loop partials go through the windows
	char *win;                   // this rank's window (cudaMalloc, exported with cudaIpcGetMemHandle)
	char *peer_win[SPH_MG_MAX_RANKS]; // every rank's window mapped into this process ([rank] == win)
	size_t win_bytes;
	int epoch;                   // exchanges issued so far (same sequence on every rank)
	int *send_slot[2], *recv_slot; // sorted slots of the particles sent to each side / of the ghosts, per step
	// overlap: edge particles (= every particle sent to a neighbour) as a list and as a bit mask over the sorted slots
	int overlap;                 // SPH_MG_OVERLAP and a slab at least two columns wide (no particle is sent to both sides)
	cudaStream_t xs;             // exchange stream (highest priority)
	cudaEvent_t ev_fork, ev_join;
	int *edge_list;
	uint32_t *edge_mask;
	int2 *push_tag;              // per sorted slot: where the particle's value goes in each neighbour's receive block (-1: nowhere)
	int *pushed_epoch;           // device word written by a producing sweep that pushed its own edge values
	int push_pending;            // host: epoch handed to a producer with mg_push_args, consumed by the next exchange
	float4 *quirk;               // rigid scenes: (pos, rho) of the fluid particles with global id < Nr, replicated
};

#define NCCL_OK(h, expr)                                                                                 \
	do {                                                                                                 \
		ncclResult_t r_ = (expr);                                                                        \
		if (r_ != ncclSuccess) return sph_fail((h), SPH_ECUDA, "NCCL: %s (%s)", g_nccl.GetErrorString(r_), #expr); \
	} while (0)

// inside the void exchange helpers: record the failure; the enclosing sph_step / sph_phase returns it (check_launch)
#define NCCL_LATCH(h, expr)                                                                              \
	do {                                                                                                 \
		ncclResult_t r_ = (expr);                                                                        \
		if (r_ != ncclSuccess && !(h)->async_error)                                                      \
			(h)->async_error = sph_fail((h), SPH_ECUDA, "NCCL: %s (%s)", g_nccl.GetErrorString(r_), #expr); \
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
	else if (what == MG_F4_T2 || what == MG_F4_T3 || what == MG_F4_T1W) v = make_float4(a[s].w, 0.0f, 0.0f, 0.0f);
	else v = a[s];
	(left ? out_l : out_r)[kk] = v;
}
__global__ void __launch_bounds__(256)
k_mg_unpack_values(int what, const int *__restrict__ slot_of, int first_orig, int nl, int nr,
                   const float4 *__restrict__ in_l, const float4 *__restrict__ in_r,
                   const float4 *__restrict__ spos, float4 *__restrict__ a, float4 *__restrict__ b) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nl + nr) return;
	int s = slot_of[first_orig + k]; // ghosts are stored left block first, then right block
	float4 v = k < nl ? in_l[k] : in_r[k - nl];
	float4 p = spos[s]; // the payload buffers carry a position copy; ghosts get theirs here
	if (what == MG_F4_T1R) { a[s] = make_float4(p.x, p.y, p.z, v.x); b[s] = make_float4(p.x, p.y, p.z, v.y); }
	else if (what == MG_F4_T2 || what == MG_F4_T3 || what == MG_F4_T1W) a[s] = make_float4(p.x, p.y, p.z, v.x);
	else a[s] = v;
}

// ---- peer-memory transport ------------------------------------------------------------------------
// "LL16" protocol: every 16-byte slot carries its own epoch tag in the last word and is written with ONE
// 128-bit store, which NVLink delivers atomically (the assumption NCCL's LL / LL128 protocols make for 8-
// and 128-byte stores).  The receiver polls each slot until the tag equals the epoch: no fence, no
// separate flag, no "last block" -- the fence.sys + flag version of this kernel took 23 us per exchange,
// this one is bounded by the NVLink round trip.  Ghost values never need their .w (velocities: the
// warm-start scalar of a ghost is not read; payloads: one or two scalars), so the tag costs no bandwidth.
struct MgPeers {
	char *w[SPH_MG_MAX_RANKS];
};
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
// poll a slot until its tag word (.w) equals `epoch`; gives up after 10 s (a dead peer must not hang the
// GPU) and latches SPH_ERR_COMM_TIMEOUT
__device__ __forceinline__ uint4 wait_slot(const void *p, int epoch, SphCtl *ctl) {
	uint4 v = ld_slot(p);
	if ((int)v.w == epoch) return v;
	// a peer that already timed out once is gone: do not wait 10 s per slot and exchange for the rest of the run
	if (*(volatile int *)&ctl->error_flags & SPH_ERR_COMM_TIMEOUT) return v;
	unsigned long long t0 = global_ns();
	for (;;) {
		v = ld_slot(p);
		if ((int)v.w == epoch) return v;
		if (global_ns() - t0 > 10000000000ull) { atomicOr(&ctl->error_flags, SPH_ERR_COMM_TIMEOUT); return v; }
	}
}

// One kernel per exchange.  Blocks 0 .. gridDim-2 (or all, without a reduction): store this rank's values of
// the sent particles straight into the neighbours' windows, then poll the slots the neighbours fill and
// scatter them into the ghost slots.  Last block (when do_reduce): reduce the sweep's block partials, store
// {sum, count, max} into every rank's window, poll all ranks' slots, sum in rank order and take the loop
// decision.  A block pushes before it polls and the grid is far smaller than one resident wave, so two
// ranks can never wait for each other's unscheduled blocks.
__global__ void __launch_bounds__(256)
k_mg_exchange(int what, const int *__restrict__ send_slot_l, const int *__restrict__ send_slot_r,
              const int *__restrict__ recv_slot, int nsl, int nsr, int nrl, int nrr,
              const float4 *src_a, const float4 *src_b, float4 *dst_a, float4 *dst_b, const float4 *__restrict__ spos,
              MgPeers peers, char *win, int cap, int rank, int nranks, int epoch, int do_reduce,
              const SphPartial *__restrict__ partials, int n_partials, int ctl_kind, SphCtlArgs cargs,
              double *__restrict__ red, SphCtl *ctl, const int *__restrict__ pushed_epoch) {
	const int parity = epoch & 1;
	if (do_reduce && blockIdx.x == gridDim.x - 1) {
		double sum; int cnt; float mx;
		sph_reduce_partials<256>(partials, n_partials, sum, cnt, mx);
		__shared__ double s_sum[SPH_MG_MAX_RANKS];
		__shared__ int s_cnt[SPH_MG_MAX_RANKS];
		__shared__ float s_max[SPH_MG_MAX_RANKS];
		if (threadIdx.x < nranks) { // one thread per destination rank: two tagged 16-byte slots
			char *slot = (char *)win_ctl(peers.w[threadIdx.x])->rslot[parity][rank];
			unsigned long long sb = (unsigned long long)__double_as_longlong(sum);
			st_slot(slot, (uint32_t)sb, (uint32_t)(sb >> 32), 0u, (uint32_t)epoch);
			st_slot(slot + 16, (uint32_t)cnt, __float_as_uint(mx), 0u, (uint32_t)epoch);
		}
		if (threadIdx.x < nranks) { // ... and one thread per source rank
			const char *slot = (const char *)win_ctl(win)->rslot[parity][threadIdx.x];
			uint4 a = wait_slot(slot, epoch, ctl), b = wait_slot(slot + 16, epoch, ctl);
			s_sum[threadIdx.x] = __longlong_as_double((long long)(((unsigned long long)a.y << 32) | a.x));
			s_cnt[threadIdx.x] = (int)b.x;
			s_max[threadIdx.x] = __uint_as_float(b.y);
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			double tsum = 0.0;
			int tcnt = 0;
			float tmax = -INFINITY;
			for (int r = 0; r < nranks; ++r) { tsum += s_sum[r]; tcnt += s_cnt[r]; tmax = fmaxf(tmax, s_max[r]); }
			red[0] = tsum; red[1] = (double)tcnt; red[2] = (double)tmax;
			sph_ctl_apply(ctl_kind, ctl, tsum, tcnt, tmax, cargs); // every rank takes the same decision
		}
		return;
	}
	if (!src_a) return;
	// Handshake with both neighbours on EVERY exchange, whatever the halo counts: a rank that sends to an empty
	// neighbour column would otherwise never wait for that neighbour, run two epochs ahead and overwrite slots of
	// the same parity before they were read.  With it a rank cannot leave exchange e before both neighbours have
	// entered e, i.e. finished reading e-1: one epoch of slack, which is what the two parities cover.
	if (blockIdx.x == 0 && threadIdx.x < 2) {
		int side = threadIdx.x, peer = side == 0 ? rank - 1 : rank + 1;
		if (peer >= 0 && peer < nranks) {
			st_slot(&win_ctl(peers.w[peer])->sync[parity][1 - side], 0u, 0u, 0u, (uint32_t)epoch);
			wait_slot(&win_ctl(win)->sync[parity][side], epoch, ctl);
		}
	}
	const int nblk = (int)gridDim.x - (do_reduce ? 1 : 0);
	const int stride = nblk * (int)blockDim.x;
	const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
	// phase 1: push.  My left neighbour receives on its right side (1), my right neighbour on its left side (0).
	// Skipped when the producing sweep pushed these values from its own epilogue (it recorded the epoch; a producer
	// that was gated off by its loop flag did not, and the -- unchanged -- values are pushed here as before).
	const bool pushed = pushed_epoch && *pushed_epoch == epoch;
	for (int k = t0; !pushed && k < nsl + nsr; k += stride) {
		bool left = k < nsl;
		int kk = left ? k : k - nsl;
		int s = (left ? send_slot_l : send_slot_r)[kk];
		float4 v;
		if (what == MG_F4_T1R) v = make_float4(src_a[s].w, src_b[s].w, 0.0f, 0.0f); // posT1.w, posR.w
		else if (what == MG_F4_T2 || what == MG_F4_T3 || what == MG_F4_T1W) v = make_float4(src_a[s].w, 0.0f, 0.0f, 0.0f);
		else v = src_a[s];
		float4 *dst = left ? win_xr(peers.w[rank - 1], cap, parity, 1) : win_xr(peers.w[rank + 1], cap, parity, 0);
		st_slot(&dst[kk], __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), (uint32_t)epoch);
	}
	// phase 2: poll + unpack.  Ghosts are stored left block first, then right block.
	for (int k = t0; k < nrl + nrr; k += stride) {
		int s = recv_slot[k];
		uint4 u = k < nrl ? wait_slot(&win_xr(win, cap, parity, 0)[k], epoch, ctl)
		                  : wait_slot(&win_xr(win, cap, parity, 1)[k - nrl], epoch, ctl);
		float4 p = spos[s]; // the payload buffers carry a position copy; ghosts get theirs here
		float vx = __uint_as_float(u.x), vy = __uint_as_float(u.y), vz = __uint_as_float(u.z);
		if (what == MG_F4_T1R) { dst_a[s] = make_float4(p.x, p.y, p.z, vx); dst_b[s] = make_float4(p.x, p.y, p.z, vy); }
		else if (what == MG_F4_T2 || what == MG_F4_T3 || what == MG_F4_T1W) dst_a[s] = make_float4(p.x, p.y, p.z, vx);
		else dst_a[s] = make_float4(vx, vy, vz, 0.0f);
	}
}

// Particle messages over the peer windows (migration, ghost particles): the packed message of each side
// (msg_send: count known on the device) goes straight into the neighbour's window, count in the first slot; the
// slots the neighbours fill are polled and copied into msg_recv in the layout k_mg_rebuild / k_mg_unpack_halo read.
// The count slot doubles as the handshake with both neighbours (every rank pushes one, empty or not).
__global__ void __launch_bounds__(256)
k_mg_xfer(int kind, const char *__restrict__ send_l, const char *__restrict__ send_r, char *recv_l, char *recv_r,
          const int *__restrict__ counters, int idx_l, int idx_r, int cap_msg, MgPeers peers, char *win, int cap, int rank,
          int nranks, int epoch, SphCtl *ctl) {
	const bool has_l = rank > 0, has_r = rank + 1 < nranks;
	const int nsl = has_l ? min(counters[idx_l], cap_msg) : 0, nsr = has_r ? min(counters[idx_r], cap_msg) : 0;
	const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
	// push: my left neighbour receives on its right side (1), my right neighbour on its left side (0)
	for (int k = t0; k < nsl + nsr; k += stride) {
		bool left = k < nsl;
		int kk = left ? k : k - nsl;
		const char *msg = left ? send_l : send_r;
		char *area = left ? win_pm(peers.w[rank - 1], cap, kind, 1) : win_pm(peers.w[rank + 1], cap, kind, 0);
		float4 p = msg_pos((char *)msg)[kk], v = msg_vel((char *)msg, cap_msg)[kk];
		int g = msg_gid((char *)msg, cap_msg)[kk];
		uint4 *slots = (uint4 *)(area + 64);
		st_slot(&slots[kk], __float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), (uint32_t)epoch);
		st_slot(&slots[(size_t)cap + kk], __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), (uint32_t)epoch);
		st_slot(&slots[(size_t)2 * cap + kk], __float_as_uint(v.w), (uint32_t)g, 0u, (uint32_t)epoch);
	}
	if (t0 == 0) {
		if (has_l) st_slot(win_pm(peers.w[rank - 1], cap, kind, 1), (uint32_t)nsl, 0u, 0u, (uint32_t)epoch);
		if (has_r) st_slot(win_pm(peers.w[rank + 1], cap, kind, 0), (uint32_t)nsr, 0u, 0u, (uint32_t)epoch);
	}
	// poll: the counts first (every block for itself), then the particles
	__shared__ int s_n[2];
	if (threadIdx.x < 2) {
		bool ex = threadIdx.x == 0 ? has_l : has_r;
		s_n[threadIdx.x] = ex ? (int)wait_slot(win_pm(win, cap, kind, threadIdx.x), epoch, ctl).x : 0;
	}
	__syncthreads();
	const int nrl = min(s_n[0], cap_msg), nrr = min(s_n[1], cap_msg);
	if (t0 == 0) {
		if (has_l) ((int *)recv_l)[0] = nrl;
		if (has_r) ((int *)recv_r)[0] = nrr;
	}
	for (int k = t0; k < nrl + nrr; k += stride) {
		bool left = k < nrl;
		int kk = left ? k : k - nrl;
		const uint4 *slots = (const uint4 *)(win_pm(win, cap, kind, left ? 0 : 1) + 64);
		uint4 a = wait_slot(&slots[kk], epoch, ctl), b = wait_slot(&slots[(size_t)cap + kk], epoch, ctl),
		      c3 = wait_slot(&slots[(size_t)2 * cap + kk], epoch, ctl);
		char *msg = left ? recv_l : recv_r;
		msg_pos(msg)[kk] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), 0.0f);
		msg_vel(msg, cap_msg)[kk] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(c3.x));
		msg_gid(msg, cap_msg)[kk] = (int)c3.y;
	}
}

// NCCL transport: this rank's partial before, and the decision after, the all-reduce
__global__ void __launch_bounds__(256) k_mg_reduce_partials(const SphPartial *partials, int n, double *red) {
	double sum; int cnt; float mx;
	sph_reduce_partials<256>(partials, n, sum, cnt, mx);
	if (threadIdx.x == 0) { red[0] = sum; red[1] = (double)cnt; red[2] = (double)mx; }
}
void sph_reduce_partials_launch(SphHandle *h, int n_blocks, cudaStream_t st) {
	k_mg_reduce_partials<<<1, 256, 0, st>>>(h->partials, n_blocks, h->red);
	h->launches++;
}
__global__ void k_mg_ctl_apply(int ctl_kind, SphCtlArgs cargs, const double *__restrict__ red, SphCtl *ctl) {
	sph_ctl_apply(ctl_kind, ctl, red[0], (int)red[1], (float)red[2], cargs);
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

// Allocate this rank's window, exchange the IPC handles (one NCCL all-gather at start-up) and map the peers'.
static int mg_open_windows(SphHandle *h, SphComm *m) {
	if (m->nranks > SPH_MG_MAX_RANKS) return sph_fail(h, SPH_EINVAL, "multi-GPU: at most %d ranks", SPH_MG_MAX_RANKS);
	m->win_bytes = win_total_bytes(m->cap_halo);
	for (int d = 0; d < 2; ++d) SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->send_slot[d], sizeof(int) * (size_t)m->cap_halo));
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->recv_slot, sizeof(int) * 2 * (size_t)m->cap_halo));
	{
		size_t ncap = (size_t)h->cfg.n_fluid + (size_t)h->cfg.n_ghost_capacity;
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->push_tag, sizeof(int2) * ncap));
		SPH_CUDA_CHECK(h, cudaMemset(m->push_tag, 0xff, sizeof(int2) * ncap));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->pushed_epoch, sizeof(int)));
		SPH_CUDA_CHECK(h, cudaMemset(m->pushed_epoch, 0, sizeof(int)));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->edge_list, sizeof(int) * 2 * (size_t)m->cap_halo));
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->edge_mask, sizeof(uint32_t) * ((ncap + 31) / 32 + 1)));
		SPH_CUDA_CHECK(h, cudaMemset(m->edge_mask, 0, sizeof(uint32_t) * ((ncap + 31) / 32 + 1)));
		int lo = 0, hi = 0;
		SPH_CUDA_CHECK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi)); // hi = numerically lowest = highest priority
		SPH_CUDA_CHECK(h, cudaStreamCreateWithPriority(&m->xs, cudaStreamNonBlocking, hi));
		SPH_CUDA_CHECK(h, cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
		SPH_CUDA_CHECK(h, cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
		// a particle of a one-column slab is sent to both sides and would be swept twice by the edge launch
		m->overlap = (getenv("SPH_MG_OVERLAP") != nullptr && atoi(getenv("SPH_MG_OVERLAP")) != 0 && m->col_hi - m->col_lo >= 2) ? 1 : 0;
	}
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->win, m->win_bytes));
	SPH_CUDA_CHECK(h, cudaMemset(m->win, 0, m->win_bytes));
	cudaIpcMemHandle_t mine;
	SPH_CUDA_CHECK(h, cudaIpcGetMemHandle(&mine, m->win));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
	char *dbuf = nullptr;
	SPH_CUDA_CHECK(h, cudaMalloc((void **)&dbuf, 64 * (size_t)m->nranks));
	SPH_CUDA_CHECK(h, cudaMemcpy(dbuf + 64 * (size_t)m->rank, &mine, 64, cudaMemcpyHostToDevice));
	SPH_CUDA_CHECK(h, cudaDeviceSynchronize()); // the memsets above must have landed before any peer can write
	NCCL_OK(h, g_nccl.AllGather(dbuf + 64 * (size_t)m->rank, dbuf, 64, ncclChar, m->comm, 0));
	SPH_CUDA_CHECK(h, cudaStreamSynchronize(0));
	cudaIpcMemHandle_t all[SPH_MG_MAX_RANKS];
	SPH_CUDA_CHECK(h, cudaMemcpy(all, dbuf, 64 * (size_t)m->nranks, cudaMemcpyDeviceToHost));
	cudaFree(dbuf);
	for (int r = 0; r < m->nranks; ++r) {
		if (r == m->rank) { m->peer_win[r] = m->win; continue; }
		cudaError_t e = cudaIpcOpenMemHandle((void **)&m->peer_win[r], all[r], cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess)
			return sph_fail(h, SPH_ECUDA, "multi-GPU: cudaIpcOpenMemHandle(rank %d) failed: %s (peer access over NVLink is "
			                              "required; SPH_MG_TRANSPORT=nccl selects the NCCL transport)", r, cudaGetErrorString(e));
	}
	m->p2p = 1;
	m->epoch = 0;
	return SPH_OK;
}

extern "C" int sph_comm_init(SphHandle *h, const char *id128, int rank, int nranks, int col_lo, int col_hi) {
	if (!h || !id128) return SPH_EINVAL;
	if (nranks < 2 || rank < 0 || rank >= nranks) return sph_fail(h, SPH_EINVAL, "sph_comm_init: bad rank %d / %d", rank, nranks);
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
	if (h->c.Nr > 0) {
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&m->quirk, sizeof(float4) * (size_t)h->c.Nr));
		SPH_CUDA_CHECK(h, cudaMemset(m->quirk, 0, sizeof(float4) * (size_t)h->c.Nr));
	}
	h->comm = m;
	const char *tr = getenv("SPH_MG_TRANSPORT");
	if (!(tr && strcmp(tr, "nccl") == 0)) {
		int rc2 = mg_open_windows(h, m);
		if (rc2 != SPH_OK) { mg_destroy(h); return rc2; }
	}
	return SPH_OK;
}

void mg_destroy(SphHandle *h) {
	SphComm *m = h->comm;
	if (!m) return;
	for (int d = 0; d < 2; ++d) {
		cudaFree(m->msg_send[d]); cudaFree(m->msg_recv[d]); cudaFree(m->send_orig[d]); cudaFree(m->xsend[d]); cudaFree(m->xrecv[d]);
	}
	cudaFree(m->counters); cudaFreeHost(m->counters_host); cudaFree(m->tmp_pos); cudaFree(m->tmp_vel); cudaFree(m->tmp_gid);
	if (m->p2p) {
		cudaDeviceSynchronize();
		for (int r = 0; r < m->nranks; ++r)
			if (r != m->rank && m->peer_win[r]) cudaIpcCloseMemHandle(m->peer_win[r]);
	}
	cudaFree(m->quirk); cudaFree(m->win); cudaFree(m->send_slot[0]); cudaFree(m->send_slot[1]); cudaFree(m->recv_slot); cudaFree(m->push_tag); cudaFree(m->pushed_epoch); cudaFree(m->edge_list); cudaFree(m->edge_mask);
	if (m->xs) { cudaStreamDestroy(m->xs); cudaEventDestroy(m->ev_fork); cudaEventDestroy(m->ev_join); }
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

// one particle message per side over the peer windows (kind 0 = migration, 1 = ghost particles)
static void xfer_particles(SphHandle *h, SphComm *m, int kind, int idx_l, int idx_r, int cap_msg, int expect, cudaStream_t st) {
	MgPeers peers;
	for (int r = 0; r < SPH_MG_MAX_RANKS; ++r) peers.w[r] = r < m->nranks ? m->peer_win[r] : nullptr;
	int blocks = cdiv(expect > 0 ? expect : 1, 256);
	if (blocks > 296) blocks = 296; // one resident wave: pushing blocks are never queued behind polling ones
	int epoch = ++m->epoch;
	k_mg_xfer<<<blocks, 256, 0, st>>>(kind, m->msg_send[0], m->msg_send[1], m->msg_recv[0], m->msg_recv[1], m->counters, idx_l,
	                                  idx_r, cap_msg, peers, m->win, m->cap_halo, m->rank, m->nranks, epoch, h->ctl);
	h->launches++;
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
	sph_prof_begin(h, KC_MG_STEP, st);
	cudaMemsetAsync(m->counters, 0, sizeof(int) * MC_COUNT, st);
	// (1) migration
	k_mg_classify<<<cdiv(n > 0 ? n : 1, 256), 256, 0, st>>>(h->pos, h->vel, h->gid, n, c.h, m->col_lo, m->col_hi, has_left,
	                                                        has_right, m->tmp_pos, m->tmp_vel, m->tmp_gid, m->msg_send[0],
	                                                        m->msg_send[1], m->cap_mig, m->counters);
	int rc = SPH_OK;
	if (m->p2p) {
		xfer_particles(h, m, 0, MC_MIG_L, MC_MIG_R, m->cap_mig, m->n_send[0] / 8 + 256, st); // a few particles per step
	} else {
		k_mg_header<<<1, 32, 0, st>>>(m->msg_send[0], m->msg_send[1], m->counters, MC_MIG_L, MC_MIG_R, m->cap_mig);
		rc = exchange_bytes(h, m, m->msg_bytes_mig, st);
		if (rc != SPH_OK) { sph_prof_end(h, st); return rc; }
	}
	k_mg_rebuild<<<cdiv(capacity_owned, 256), 256, 0, st>>>(m->tmp_pos, m->tmp_vel, m->tmp_gid, m->msg_recv[0], m->msg_recv[1],
	                                                        has_left, has_right, m->cap_mig, h->pos, h->vel, h->gid,
	                                                        capacity_owned, m->counters);
	// (2) halo particles
	k_mg_pack_halo<<<cdiv(capacity_owned, 256), 256, 0, st>>>(h->pos, h->vel, h->gid, c.h, m->col_lo, m->col_hi, has_left,
	                                                          has_right, m->msg_send[0], m->msg_send[1], m->cap_halo,
	                                                          m->send_orig[0], m->send_orig[1], m->counters);
	if (m->p2p) {
		// grid sized from the previous step's ghost counts (the kernel strides; the true counts are on the device)
		int expect = m->n_send[0] + m->n_send[1] > m->n_recv[0] + m->n_recv[1] ? m->n_send[0] + m->n_send[1] : m->n_recv[0] + m->n_recv[1];
		xfer_particles(h, m, 1, MC_HALO_L, MC_HALO_R, m->cap_halo, expect > 0 ? expect : 4 * m->cap_halo / 8, st);
	} else {
		k_mg_header<<<1, 32, 0, st>>>(m->msg_send[0], m->msg_send[1], m->counters, MC_HALO_L, MC_HALO_R, m->cap_halo);
		rc = exchange_bytes(h, m, m->msg_bytes_halo, st);
		if (rc != SPH_OK) { sph_prof_end(h, st); return rc; }
	}
	k_mg_unpack_halo<<<cdiv(2 * m->cap_halo, 256), 256, 0, st>>>(m->msg_recv[0], m->msg_recv[1], has_left, has_right,
	                                                             m->cap_halo, h->pos, h->vel, h->gid, capacity_all, m->counters);
	h->launches += 6;
	// (3) the one host read-back of the step
	SPH_CUDA_CHECK(h, cudaMemcpyAsync(m->counters_host, m->counters, sizeof(int) * MC_COUNT, cudaMemcpyDeviceToHost, st));
	// the flags latched by the previous step's exchanges travel with the counters (WCSPH / PBF have no other read-back)
	SPH_CUDA_CHECK(h, cudaMemcpyAsync(&h->ctl_host->error_flags, &h->ctl->error_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
	sph_prof_end(h, st);
	SPH_CUDA_CHECK(h, cudaStreamSynchronize(st));
	const int *k = m->counters_host;
	if (h->ctl_host && (h->ctl_host->error_flags & SPH_ERR_COMM_TIMEOUT))
		return sph_fail(h, SPH_ESTATE, "multi-GPU: a peer rank stopped answering (exchange timed out after 10 s)");
	if (k[MC_OVERFLOW]) return sph_fail(h, SPH_ESTATE, "multi-GPU: message or slab capacity exceeded (flags %d)", k[MC_OVERFLOW]);
	h->c.N_owned = k[MC_OWNED];
	h->c.N = k[MC_OWNED] + k[MC_GHOST_L] + k[MC_GHOST_R];
	m->n_send[0] = k[MC_HALO_L] < m->cap_halo ? k[MC_HALO_L] : m->cap_halo;
	m->n_send[1] = k[MC_HALO_R] < m->cap_halo ? k[MC_HALO_R] : m->cap_halo;
	m->n_recv[0] = k[MC_GHOST_L];
	m->n_recv[1] = k[MC_GHOST_R];
	return SPH_OK;
}

// once per step, after the sort: where the sent particles and the ghosts live in the sorted arrays
__global__ void __launch_bounds__(256)
k_mg_slots(const int *__restrict__ send_orig_l, const int *__restrict__ send_orig_r, const int *__restrict__ slot_of,
           int nsl, int nsr, int first_orig, int nr, int *__restrict__ send_slot_l, int *__restrict__ send_slot_r,
           int *__restrict__ recv_slot, int2 *__restrict__ push_tag, int *__restrict__ edge_list, uint32_t *__restrict__ edge_mask) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	// push_tag was filled with (-1, -1); a particle of a one-column slab is sent to both sides (two different threads)
	if (k < nsl) {
		int s = slot_of[send_orig_l[k]];
		send_slot_l[k] = s; push_tag[s].x = k;
		edge_list[k] = s; atomicOr(&edge_mask[s >> 5], 1u << (s & 31));
	}
	if (k < nsr) {
		int s = slot_of[send_orig_r[k]];
		send_slot_r[k] = s; push_tag[s].y = k;
		edge_list[nsl + k] = s; atomicOr(&edge_mask[s >> 5], 1u << (s & 31));
	}
	if (k < nr) recv_slot[k] = slot_of[first_orig + k];
}
void mg_after_grid(SphHandle *h, cudaStream_t st) {
	SphComm *m = h->comm;
	if (!m || !m->p2p) return;
	int nr = m->n_recv[0] + m->n_recv[1];
	int work = nr > m->n_send[0] ? nr : m->n_send[0];
	if (m->n_send[1] > work) work = m->n_send[1];
	if (h->c.N > 0) {
		cudaMemsetAsync(m->push_tag, 0xff, sizeof(int2) * (size_t)h->c.N, st); // (-1, -1): not sent
		cudaMemsetAsync(m->edge_mask, 0, sizeof(uint32_t) * (size_t)((h->c.N + 31) / 32), st);
	}
	if (work <= 0) return;
	k_mg_slots<<<cdiv(work, 256), 256, 0, st>>>(m->send_orig[0], m->send_orig[1], h->fg.slot_of, m->n_send[0], m->n_send[1],
	                                            h->c.N_owned, nr, m->send_slot[0], m->send_slot[1], m->recv_slot, m->push_tag,
	                                            m->edge_list, m->edge_mask);
	h->launches++;
}

// Ghost values of one field to / from both neighbours; when ctl_kind != SPH_CTL_NONE the all-reduce of
// h->red = {sum, count, max} (left there by the tail of the sweep that just ran) rides in the same
// exchange and the loop decision is applied on every rank.
static void mg_exchange_impl(SphHandle *h, int what, int ctl_kind, int reduce_blocks, cudaStream_t st) {
	if (ctl_kind == SPH_CTL_NONE) reduce_blocks = 0;
	SphCtlArgs cargs;
	cargs.dt_cfl_c1 = h->c.dt_cfl_c1;
	cargs.rs = h->rstate;
	cargs.rigid_exists = (h->c.Nr > 0 && h->rigid_ready) ? 1 : 0;
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
	case MG_F4_T1W: a = wa = h->a4[A4_T1]; break;
	default:
		if (what >= MG_XYZ(0) && what < MG_XYZ(A4_COUNT)) a = wa = h->a4[what - MG_XYZ(0)];
		break;
	}
	int ns = m->n_send[0] + m->n_send[1], nr = m->n_recv[0] + m->n_recv[1];
	if (m->p2p) {
		sph_prof_begin(h, KC_MG_EXCHANGE, st);
		int epoch = m->push_pending ? m->push_pending : ++m->epoch; // a producer already holds this exchange's epoch
		m->push_pending = 0;
		MgPeers peers;
		for (int r = 0; r < SPH_MG_MAX_RANKS; ++r) peers.w[r] = r < m->nranks ? m->peer_win[r] : nullptr;
		int work = ns > nr ? ns : nr;
		int blocks = a ? cdiv(work > 0 ? work : 1, 256) : 0;
		if (blocks > 296) blocks = 296; // far below one resident wave: pushing blocks are never queued behind polling ones
		if (reduce_blocks > 0) blocks += 1;
		if (blocks > 0) {
			k_mg_exchange<<<blocks, 256, 0, st>>>(what, m->send_slot[0], m->send_slot[1], m->recv_slot, m->n_send[0], m->n_send[1],
			                                      m->n_recv[0], m->n_recv[1], a, b, wa, wb, h->a4[A4_POS], peers,
			                                      m->win, m->cap_halo, m->rank, m->nranks, epoch, reduce_blocks > 0 ? 1 : 0,
			                                      h->partials, reduce_blocks, ctl_kind, cargs, h->red, h->ctl, m->pushed_epoch);
			h->launches += 1;
		}
		sph_prof_end(h, st);
		return;
	}
	sph_prof_begin(h, KC_MG_EXCHANGE, st);
	if (a && ns > 0) {
		k_mg_pack_values<<<cdiv(ns, 256), 256, 0, st>>>(what, m->send_orig[0], m->send_orig[1], h->fg.slot_of, m->n_send[0],
		                                                m->n_send[1], a, b, m->xsend[0], m->xsend[1]);
		h->launches++;
	}
	if (reduce_blocks > 0) sph_reduce_partials_launch(h, reduce_blocks, st);
	int left = m->rank - 1, right = m->rank + 1;
	NCCL_LATCH(h, g_nccl.GroupStart());
	if (a && left >= 0) {
		if (m->n_send[0] > 0) NCCL_LATCH(h, g_nccl.Send(m->xsend[0], (size_t)m->n_send[0] * 4, ncclFloat, left, m->comm, st));
		if (m->n_recv[0] > 0) NCCL_LATCH(h, g_nccl.Recv(m->xrecv[0], (size_t)m->n_recv[0] * 4, ncclFloat, left, m->comm, st));
	}
	if (a && right < m->nranks) {
		if (m->n_send[1] > 0) NCCL_LATCH(h, g_nccl.Send(m->xsend[1], (size_t)m->n_send[1] * 4, ncclFloat, right, m->comm, st));
		if (m->n_recv[1] > 0) NCCL_LATCH(h, g_nccl.Recv(m->xrecv[1], (size_t)m->n_recv[1] * 4, ncclFloat, right, m->comm, st));
	}
	if (reduce_blocks > 0) {
		NCCL_LATCH(h, g_nccl.AllReduce(h->red, h->red, 2, ncclDouble, ncclSum, m->comm, st));
		NCCL_LATCH(h, g_nccl.AllReduce(h->red + 2, h->red + 2, 1, ncclDouble, ncclMax, m->comm, st));
	}
	NCCL_LATCH(h, g_nccl.GroupEnd());
	if (reduce_blocks > 0) { k_mg_ctl_apply<<<1, 1, 0, st>>>(ctl_kind, cargs, h->red, h->ctl); h->launches++; }
	if (a && nr > 0) {
		k_mg_unpack_values<<<cdiv(nr, 256), 256, 0, st>>>(what, h->fg.slot_of, h->c.N_owned, m->n_recv[0], m->n_recv[1],
		                                                  m->xrecv[0], m->xrecv[1], h->a4[A4_POS], wa, wb);
		h->launches++;
	}
	sph_prof_end(h, st);
}

// The sweep about to be launched produces the values of the NEXT exchange: hand it that exchange's epoch and the
// neighbours' windows so that it stores its edge values itself (sph_mgwin.cuh); the exchange kernel then only polls
// and unpacks.  Must be followed by exactly one mg_exchange / mg_exchange_reduce of those values.
// MEASURED on 2 x B200, 10^6 particles per GPU (profiles/r2_experiments.md): the exchange kernels get shorter
// (0.63 -> 0.52 ms per step) but every sweep pays for the tag load and the peer stores: 4.30 ms per step against
// 4.14 ms.  Hence OPT-IN (SPH_MG_EPILOGUE_PUSH=1); by default tag == nullptr and the exchange kernel pushes.
SphMgPush mg_push_args(SphHandle *h) {
	SphMgPush pu;
	memset(&pu, 0, sizeof(pu));
	SphComm *m = h->comm;
	static const bool enabled = getenv("SPH_MG_EPILOGUE_PUSH") != nullptr;
	if (!m || !m->p2p || !enabled) return pu;
	pu.tag = m->push_tag;
	pu.peer_l = m->rank > 0 ? m->peer_win[m->rank - 1] : nullptr;
	pu.peer_r = m->rank + 1 < m->nranks ? m->peer_win[m->rank + 1] : nullptr;
	pu.pushed_epoch = m->pushed_epoch;
	pu.cap = m->cap_halo;
	pu.epoch = m->push_pending = ++m->epoch;
	return pu;
}

MgSplit mg_split(SphHandle *h) {
	MgSplit sp;
	memset(&sp, 0, sizeof(sp));
	SphComm *m = h->comm;
	if (!m || !m->p2p || !m->overlap) return sp;
	sp.on = true;
	sp.xs = m->xs;
	sp.edge_list = m->edge_list;
	sp.edge_mask = m->edge_mask;
	sp.n_edge = m->n_send[0] + m->n_send[1];
	return sp;
}
void mg_fork(SphHandle *h, cudaStream_t main_stream) {
	SphComm *m = h->comm;
	cudaEventRecord(m->ev_fork, main_stream);
	cudaStreamWaitEvent(m->xs, m->ev_fork, 0);
}
void mg_join(SphHandle *h, cudaStream_t main_stream) {
	SphComm *m = h->comm;
	cudaEventRecord(m->ev_join, m->xs);
	cudaStreamWaitEvent(main_stream, m->ev_join, 0);
}

void mg_exchange(SphHandle *h, int what, cudaStream_t st) { mg_exchange_impl(h, what, SPH_CTL_NONE, 0, st); }
void mg_exchange_reduce(SphHandle *h, int what, int ctl_kind, int n_blocks, cudaStream_t st) {
	mg_exchange_impl(h, what, ctl_kind, n_blocks, st);
}

// ---- rigid bodies over slabs -------------------------------------------------------------------------------
void mg_allreduce_sum_f32(SphHandle *h, float *dev, size_t n, cudaStream_t st) {
	SphComm *m = h->comm;
	if (!m || n == 0) return;
	NCCL_LATCH(h, g_nccl.AllReduce(dev, dev, n, ncclFloat, ncclSum, m->comm, st));
}

// Two reference quirks index FLUID arrays with a rigid-local index k < Nr: the neighbour count measures the
// distance to fluid particle k (PS:440-442, SURVEY B-7) and the viscosity uses its rho (SB:199, B-6).  With
// slabs fluid particle k (a GLOBAL id) lives on some rank: its owner writes (pos, rho) into slot k of a zeroed
// array and one all-reduce replicates the Nr records (x + 0 is exact, so every rank sees the owner's bits).
__global__ void __launch_bounds__(256)
k_mg_quirk_fill(const float4 *__restrict__ pos, const int *__restrict__ gid, const int *__restrict__ slot_of,
                const float *__restrict__ rho, int n_owned, int nr, int with_rho, float4 *__restrict__ quirk) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_owned) return;
	int g = gid[i];
	if (g >= nr) return;
	float4 p = pos[i];
	quirk[g] = make_float4(p.x, p.y, p.z, with_rho ? rho[slot_of[i]] : 0.0f);
}
const float4 *mg_rigid_quirk(const SphHandle *h) { return (h->comm && h->c.Nr > 0) ? h->comm->quirk : nullptr; }
void mg_rigid_quirk_update(SphHandle *h, int with_rho, cudaStream_t st) {
	SphComm *m = h->comm;
	if (!m || h->c.Nr <= 0) return;
	cudaMemsetAsync(m->quirk, 0, sizeof(float4) * (size_t)h->c.Nr, st);
	if (h->c.N_owned > 0)
		k_mg_quirk_fill<<<cdiv(h->c.N_owned, 256), 256, 0, st>>>(h->pos, h->gid, h->fg.slot_of, h->a1[A1_RHO], h->c.N_owned,
		                                                         h->c.Nr, with_rho, m->quirk);
	NCCL_LATCH(h, g_nccl.AllReduce(m->quirk, m->quirk, 4 * (size_t)h->c.Nr, ncclFloat, ncclSum, m->comm, st));
	h->launches++;
}

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
