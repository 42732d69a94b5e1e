// bp_comm.cu -- halo exchange and global reductions.
//
// The reference has no collectives of its own: the halo scatter of MatMult, the
// VecDot all-reduces and the off-process adds of parallel assembly all happen inside
// PETSc/dolfin under `mpirun -np N`.  Here each rank (one process
// per GPU) owns a slab of ice columns; ghost values of u (before assembly) and of the
// CG search direction (before each SpMV) come from the neighbours with grouped
// ncclSend/ncclRecv, and the CG / Newton scalars are all-reduced with ncclAllReduce.
// NCCL is bound at run time with dlopen so that the library loads on a box without it.
//
// That is the fallback.  On one NVSwitch box (one process per GPU) the ranks map each other's
// exchange arena with CUDA IPC at bp_comm_init and talk through PEER MEMORY instead: a halo
// exchange is ONE small kernel per rank that stores the boundary values straight into the
// neighbours' landing buffers (NVLink stores), raises a flag there, waits for the neighbours'
// flags and copies what arrived into the ghost rows; the CG / Newton scalars are all-reduced the
// same way (every rank stores its partial sums into every other rank's slots and adds them up
// in rank order: the same bits everywhere).  No pack + ncclSend/ncclRecv (three launches and a
// proxy round trip, 12-15 us), which was half of a CG iteration on the partitioned multigrid;
// both kernels are plain launches, so they are captured into the V-cycle / CG graphs.
// BP_P2P=0, or an IPC failure on any rank, keeps NCCL.
//
// Several contexts in ONE process (bp_*_group) exchange by device copies instead;
// that is how the partitioned path is exercised on a single GPU.
#include "bp_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <algorithm>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclInt64_ = 4, ncclFloat64_ = 8 };   // ncclDataType_t
enum { ncclMin_ = 3 };
enum { ncclSum_ = 0, ncclMax_ = 2 };   // ncclRedOp_t

struct nccl_api {
	void* handle = nullptr;
	int (*GetUniqueId)(ncclUniqueId*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
static nccl_api g_nccl;

// ---- peer memory ---------------------------------------------------------------------------------
#define P2P_MAX_RANKS   16
#define P2P_AR_DOUBLES  8                       // scalars per small all-reduce
#define P2P_ARENA_BYTES ((size_t)32 << 20)
// arena of every rank: [all-reduce flags: P2P_MAX_RANKS u64][all-reduce data: 2 x P2P_MAX_RANKS x P2P_AR_DOUBLES f64]
// then, bump-allocated per partitioned context (fine level and every coupled multigrid level):
// flags [n_nbr] u64 and the landing buffer [2][n_ghost] double2 (two exchanges alternate)
#define P2P_AR_FLAG_OFF 0
#define P2P_AR_DATA_OFF (P2P_MAX_RANKS * 8)
#define P2P_HEAP_OFF    4096

struct bp_comm_state {
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1;
	bool p2p = false;                           // arenas of all ranks mapped
	char* arena = nullptr;                      // this rank's (cudaMalloc, exported by cudaIpcGetMemHandle)
	size_t arena_used = P2P_HEAP_OFF;
	char* peer[P2P_MAX_RANKS] = { nullptr };    // every rank's arena in this process (peer[rank] = arena)
	char** d_peer = nullptr;                    // the same table on the device
	unsigned long long* d_ar_seq = nullptr;     // all-reduces done so far (device)
	int64_t* d_xchg = nullptr;                  // staging for the offset exchange at attach time
};

// what one partitioned context needs for its exchanges (bp_ctx::p2p)
struct bp_p2p_halo {
	int n_nbr = 0;
	int64_t n_ghost = 0;
	unsigned long long* d_seq = nullptr;        // exchanges done so far
	unsigned long long* my_flag = nullptr;      // [n_nbr] in this rank's arena: raised by neighbour k
	double2* my_land = nullptr;                 // [2][n_ghost] in this rank's arena
	// device tables, one entry per neighbour
	unsigned long long** d_rem_flag = nullptr;  // the neighbour's flag for this rank
	double2** d_rem_land = nullptr;             // where this rank's values go in the neighbour's landing buffer (first of the two)
	int64_t* d_rem_stride = nullptr;            // distance to the second one (the neighbour's n_ghost)
	int64_t* d_send_ptr = nullptr;              // [n_nbr + 1]
};

static int load_nccl(bp_ctx* c, const char* path)
{
	if (g_nccl.handle) return BP_OK;
	const char* cands[] = { path, "libnccl.so.2", "libnccl.so" };
	void* h = nullptr;
	for (const char* p : cands) { if (p && *p) { h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
	if (!h) return bp_fail(c, BP_ERR_COMM, "could not dlopen NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
	#define SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) \
		return bp_fail(c, BP_ERR_COMM, "NCCL symbol %s missing", name); } while (0)
	SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
	SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
	SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
	g_nccl.handle = h;
	return BP_OK;
}
#define BP_NCCL(ctx, call) do { int r_ = (call); if (r_ != ncclSuccess_) \
	return bp_fail(ctx, BP_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

// ---- peer-memory kernels -------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
// wait until a peer has raised *flag to `want`.  Bounded (two minutes): a rank that died must not hang this GPU --
// the caller then poisons its result with NaN and the Krylov / Newton loop reports the failure.
__device__ __forceinline__ bool p2p_wait(const unsigned long long* flag, unsigned long long want)
{
	unsigned long long t0, t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	while (ld_acquire_sys(flag) < want) {
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		if (t - t0 > 120000000000ull) return false;
	}
	return true;
}

struct HaloArgs {
	int n_nbr; int64_t n_ghost;
	unsigned long long* seq; unsigned long long* my_flag; const double2* my_land;
	unsigned long long** rem_flag; double2** rem_land; const int64_t* rem_stride; const int64_t* send_ptr; const int32_t* send_nodes;
};

// one halo exchange of one rank: boundary values -> the neighbours' landing buffers (peer stores), flag, wait for the
// neighbours' flags, landing buffer -> ghost rows.  One CTA: an interface is a few thousand nodes.
__global__ void __launch_bounds__(1024)
k_halo_p2p(const HaloArgs A, const double2* vec, double2* ghost)
{
	__shared__ int s_ok;
	const unsigned long long seq = *A.seq;
	const size_t par = (size_t)(seq & 1ull);
	if (threadIdx.x == 0) s_ok = 1;
	for (int k = 0; k < A.n_nbr; ++k) {
		const int64_t s0 = A.send_ptr[k], s1 = A.send_ptr[k + 1];
		double2* dst = A.rem_land[k] + par * (size_t)A.rem_stride[k];
		for (int64_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) dst[i - s0] = vec[A.send_nodes[i]];
	}
	__threadfence_system();
	__syncthreads();
	if ((int)threadIdx.x < A.n_nbr) {
		st_release_sys(A.rem_flag[threadIdx.x], seq + 1);
		if (!p2p_wait(A.my_flag + threadIdx.x, seq + 1)) s_ok = 0;
	}
	__syncthreads();
	const double2* src = A.my_land + par * (size_t)A.n_ghost;
	const bool ok = s_ok != 0;
	for (int64_t i = threadIdx.x; i < A.n_ghost; i += blockDim.x)
		ghost[i] = ok ? __ldcg(src + i) : make_double2(__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll));
	__syncthreads();
	if (threadIdx.x == 0) *A.seq = seq + 1;
}

// all-reduce (sum or max) of `count` <= P2P_AR_DOUBLES scalars: every rank stores its values into its slot on every
// rank, raises its flag there, waits for everybody's flag here and reduces the slots in rank order
__global__ void __launch_bounds__(32)
k_allreduce_p2p(char* const* peer, const int me, const int n, unsigned long long* seq_p, double* sc, const int count, const int op_max)
{
	__shared__ int s_ok;
	const unsigned long long seq = *seq_p;
	const size_t par = (size_t)(seq & 1ull);
	const int t = threadIdx.x;
	if (t == 0) s_ok = 1;
	__syncwarp();
	if (t < n) {
		double* dst = (double*)(peer[t] + P2P_AR_DATA_OFF) + (par * P2P_MAX_RANKS + me) * P2P_AR_DOUBLES;
		for (int i = 0; i < count; ++i) dst[i] = sc[i];
		__threadfence_system();
		st_release_sys((unsigned long long*)(peer[t] + P2P_AR_FLAG_OFF) + me, seq + 1);
		if (!p2p_wait((const unsigned long long*)(peer[me] + P2P_AR_FLAG_OFF) + t, seq + 1)) s_ok = 0;
	}
	__syncwarp();
	if (t < count) {
		const double* src = (const double*)(peer[me] + P2P_AR_DATA_OFF) + par * P2P_MAX_RANKS * P2P_AR_DOUBLES;
		double s = __ldcg(src + t);
		for (int r = 1; r < n; ++r) { const double v = __ldcg(src + (size_t)r * P2P_AR_DOUBLES + t); s = op_max ? fmax(s, v) : s + v; }
		sc[t] = s_ok ? s : __longlong_as_double(0x7ff8000000000000ll);
	}
	__syncwarp();
	if (t == 0) *seq_p = seq + 1;
}

// map every rank's arena (collective over the communicator).  Any failure on any rank leaves NCCL in charge everywhere.
static int p2p_setup(bp_ctx* c)
{
	bp_comm_state* S = c->comm;
	if (S->n_ranks < 2) return BP_OK;
	cudaStream_t st = c->stream;
	int ok = (S->n_ranks <= P2P_MAX_RANKS && c->env.p2p) ? 1 : 0;
	struct Rec { cudaIpcMemHandle_t h; int64_t ok; };
	Rec mine;
	memset(&mine, 0, sizeof(mine));
	if (ok && cudaMalloc((void**)&S->arena, P2P_ARENA_BYTES) != cudaSuccess) { cudaGetLastError(); S->arena = nullptr; ok = 0; }
	if (ok && cudaMemset(S->arena, 0, P2P_ARENA_BYTES) != cudaSuccess) { cudaGetLastError(); ok = 0; }
	if (ok && cudaIpcGetMemHandle(&mine.h, S->arena) != cudaSuccess) { cudaGetLastError(); ok = 0; }
	mine.ok = ok;
	Rec *d_mine = nullptr, *d_all = nullptr;
	std::vector<Rec> all(S->n_ranks);
	BP_CUDA(c, cudaMalloc((void**)&d_mine, sizeof(Rec)));
	BP_CUDA(c, cudaMalloc((void**)&d_all, sizeof(Rec) * S->n_ranks));
	BP_CUDA(c, cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
	BP_CUDA(c, cudaDeviceSynchronize());                 // the zeroed arena is in place before anybody can learn its handle
	BP_NCCL(c, g_nccl.AllGather(d_mine, d_all, sizeof(Rec), ncclInt8_, S->comm, st));
	BP_CUDA(c, cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * S->n_ranks, cudaMemcpyDeviceToHost, st));
	BP_CUDA(c, cudaStreamSynchronize(st));
	for (int r = 0; r < S->n_ranks; ++r) ok = ok && all[r].ok;
	int64_t bad = 0;
	if (ok) {
		for (int r = 0; r < S->n_ranks; ++r) {
			if (r == S->rank) { S->peer[r] = S->arena; continue; }
			if (cudaIpcOpenMemHandle((void**)&S->peer[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); S->peer[r] = nullptr; bad = 1; }
		}
	}
	// second round: did everybody manage to map everybody?  (re-use d_all as int64 scratch)
	mine.ok = (ok && !bad) ? 1 : 0;
	BP_CUDA(c, cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
	BP_NCCL(c, g_nccl.AllGather(d_mine, d_all, sizeof(Rec), ncclInt8_, S->comm, st));
	BP_CUDA(c, cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * S->n_ranks, cudaMemcpyDeviceToHost, st));
	BP_CUDA(c, cudaStreamSynchronize(st));
	cudaFree(d_mine); cudaFree(d_all);
	bool all_ok = true;
	for (int r = 0; r < S->n_ranks; ++r) all_ok = all_ok && all[r].ok;
	if (!all_ok) {
		for (int r = 0; r < S->n_ranks; ++r) if (r != S->rank && S->peer[r]) { cudaIpcCloseMemHandle(S->peer[r]); S->peer[r] = nullptr; }
		if (S->arena) { cudaFree(S->arena); S->arena = nullptr; }
		if (c->prm.verbose) fprintf(stderr, "[cslvr_b200] rank %d: peer memory not available, exchanges go through NCCL\n", S->rank);
		return BP_OK;
	}
	BP_CUDA(c, cudaMalloc((void**)&S->d_peer, sizeof(char*) * P2P_MAX_RANKS));
	BP_CUDA(c, cudaMemcpy(S->d_peer, S->peer, sizeof(char*) * P2P_MAX_RANKS, cudaMemcpyHostToDevice));
	BP_CUDA(c, cudaMalloc((void**)&S->d_ar_seq, sizeof(unsigned long long)));
	BP_CUDA(c, cudaMemset(S->d_ar_seq, 0, sizeof(unsigned long long)));
	BP_CUDA(c, cudaMalloc((void**)&S->d_xchg, sizeof(int64_t) * 8 * P2P_MAX_RANKS));
	S->p2p = true;
	if (c->prm.verbose) fprintf(stderr, "[cslvr_b200] rank %d: %d peer arenas mapped (CUDA IPC): halo exchanges and scalar all-reduces use NVLink stores\n", S->rank, S->n_ranks - 1);
	return BP_OK;
}

static size_t p2p_align(size_t x) { return (x + 255) & ~(size_t)255; }

// give a partitioned context (the fine level, a coupled multigrid level) its landing buffer and learn where its values
// go on the neighbours.  Collective: every rank of the communicator calls it for the same context in the same order.
int bp_comm_p2p_attach(bp_ctx* c)
{
	if (!c || !c->comm || !c->comm->p2p || c->p2p_tried) return BP_OK;
	c->p2p_tried = true;
	bp_comm_state* S = c->comm;
	const int nb = c->n_nbr;
	const int64_t ng = nb ? c->recv_ptr[nb] : 0;
	BP_CUDA(c, cudaSetDevice(c->device));
	size_t off = p2p_align(S->arena_used);
	const size_t flag_off = off; off += p2p_align((size_t)std::max(nb, 1) * 8);
	const size_t land_off = off; off += p2p_align((size_t)2 * std::max<int64_t>(ng, 1) * 16);
	double fits = (off <= P2P_ARENA_BYTES && nb <= P2P_MAX_RANKS) ? 0.0 : 1.0;
	{	// the same decision on every rank (a context that does not fit anywhere keeps NCCL everywhere)
		double* d = (double*)S->d_xchg;
		BP_CUDA(c, cudaMemcpy(d, &fits, sizeof(double), cudaMemcpyHostToDevice));
		BP_NCCL(c, g_nccl.AllReduce(d, d, 1, ncclFloat64_, ncclSum_, S->comm, 0));
		BP_CUDA(c, cudaMemcpy(&fits, d, sizeof(double), cudaMemcpyDeviceToHost));
	}
	if (fits != 0.0) return BP_OK;
	S->arena_used = off;
	if (nb == 0) return BP_OK;
	std::vector<int64_t> mine((size_t)3 * nb), theirs((size_t)3 * nb);
	for (int k = 0; k < nb; ++k) {
		mine[3 * k] = (int64_t)(land_off + (size_t)c->recv_ptr[k] * 16);   // where neighbour k's values land here (first buffer)
		mine[3 * k + 1] = ng;                                               // distance to the second buffer
		mine[3 * k + 2] = (int64_t)(flag_off + (size_t)k * 8);              // the flag neighbour k raises here
	}
	int64_t* d_m = S->d_xchg + 8;
	int64_t* d_t = S->d_xchg + 8 + 3 * P2P_MAX_RANKS;
	BP_CUDA(c, cudaMemcpy(d_m, mine.data(), mine.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
	BP_NCCL(c, g_nccl.GroupStart());
	for (int k = 0; k < nb; ++k) {
		BP_NCCL(c, g_nccl.Send(d_m + 3 * k, 3, ncclInt64_, c->nbr_rank[k], S->comm, 0));
		BP_NCCL(c, g_nccl.Recv(d_t + 3 * k, 3, ncclInt64_, c->nbr_rank[k], S->comm, 0));
	}
	BP_NCCL(c, g_nccl.GroupEnd());
	BP_CUDA(c, cudaMemcpy(theirs.data(), d_t, theirs.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
	bp_p2p_halo* H = new bp_p2p_halo();
	H->n_nbr = nb; H->n_ghost = ng;
	H->my_flag = (unsigned long long*)(S->arena + flag_off);
	H->my_land = (double2*)(S->arena + land_off);
	std::vector<unsigned long long*> rf(nb);
	std::vector<double2*> rl(nb);
	std::vector<int64_t> rs(nb);
	for (int k = 0; k < nb; ++k) {
		char* base = S->peer[c->nbr_rank[k]];
		rl[k] = (double2*)(base + theirs[3 * k]); rs[k] = theirs[3 * k + 1]; rf[k] = (unsigned long long*)(base + theirs[3 * k + 2]);
	}
	c->p2p = H;
	BP_CUDA(c, cudaMalloc((void**)&H->d_seq, sizeof(unsigned long long)));
	BP_CUDA(c, cudaMemset(H->d_seq, 0, sizeof(unsigned long long)));
	BP_CUDA(c, cudaMalloc((void**)&H->d_rem_flag, sizeof(void*) * nb));
	BP_CUDA(c, cudaMalloc((void**)&H->d_rem_land, sizeof(void*) * nb));
	BP_CUDA(c, cudaMalloc((void**)&H->d_rem_stride, sizeof(int64_t) * nb));
	BP_CUDA(c, cudaMalloc((void**)&H->d_send_ptr, sizeof(int64_t) * (nb + 1)));
	BP_CUDA(c, cudaMemcpy(H->d_rem_flag, rf.data(), sizeof(void*) * nb, cudaMemcpyHostToDevice));
	BP_CUDA(c, cudaMemcpy(H->d_rem_land, rl.data(), sizeof(void*) * nb, cudaMemcpyHostToDevice));
	BP_CUDA(c, cudaMemcpy(H->d_rem_stride, rs.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice));
	BP_CUDA(c, cudaMemcpy(H->d_send_ptr, c->send_ptr.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice));
	return BP_OK;
}

void bp_comm_p2p_detach(bp_ctx* c)
{
	if (!c || !c->p2p) return;
	bp_p2p_halo* H = c->p2p;
	void* ptrs[] = { H->d_seq, H->d_rem_flag, H->d_rem_land, H->d_rem_stride, H->d_send_ptr };
	for (void* p : ptrs) if (p) cudaFree(p);
	delete H;
	c->p2p = nullptr;
	c->p2p_tried = false;
}

extern "C" int bp_comm_unique_id(const char* nccl_library_path, char id_out[128])
{
	int rc = load_nccl(nullptr, nccl_library_path);
	if (rc) return rc;
	ncclUniqueId id;
	BP_NCCL(nullptr, g_nccl.GetUniqueId(&id));
	memcpy(id_out, id.internal, 128);
	return BP_OK;
}

extern "C" int bp_comm_init(bp_ctx* c, const char* nccl_library_path, const char id[128], int rank, int n_ranks)
{
	if (!c) return BP_ERR_ARG;
	int rc = load_nccl(c, nccl_library_path);
	if (rc) return rc;
	BP_CUDA(c, cudaSetDevice(c->device));
	bp_comm_free(c);
	c->comm = new bp_comm_state();
	c->comm->rank = rank; c->comm->n_ranks = n_ranks;
	ncclUniqueId uid;
	memcpy(uid.internal, id, 128);
	BP_NCCL(c, g_nccl.CommInitRank(&c->comm->comm, n_ranks, uid, rank));
	if ((rc = p2p_setup(c))) return rc;
	return bp_comm_p2p_attach(c);
}

void bp_comm_free(bp_ctx* c)
{
	bp_comm_p2p_detach(c);
	if (c->comm) {
		bp_comm_state* S = c->comm;
		if (S->p2p) for (int r = 0; r < S->n_ranks; ++r) if (r != S->rank && S->peer[r]) cudaIpcCloseMemHandle(S->peer[r]);
		void* ptrs[] = { S->arena, S->d_peer, S->d_ar_seq, S->d_xchg };
		for (void* p : ptrs) if (p) cudaFree(p);
		if (c->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->comm);
		delete c->comm;
		c->comm = nullptr;
	}
}

double* bp_vec(bp_ctx* c, int which)
{
	switch (which) { case 0: return c->d_u; case 1: return c->d_p; case 2: return c->d_z; case 3: return c->d_z2; case 4: return c->d_io2; case 5: return c->d_dx; default: return c->d_p; }
}

// ghost values of vector `which` (0: u, 1: p) from the owning ranks
int bp_comm_halo(bp_ctx** ctxs, int n, int which)
{
	if (n == 1) {
		bp_ctx* c = ctxs[0];
		if (c->n_nbr == 0) return BP_OK;
		if (!c->comm) return bp_fail(c, BP_ERR_STATE, "partitioned context without a communicator (bp_comm_init)");
		double* v = bp_vec(c, which);
		if (c->p2p) {
			const bp_p2p_halo* H = c->p2p;
			const HaloArgs A = { H->n_nbr, H->n_ghost, H->d_seq, H->my_flag, H->my_land, H->d_rem_flag, H->d_rem_land, H->d_rem_stride, H->d_send_ptr, c->d_send_nodes };
			k_halo_p2p<<<1, 1024, 0, c->stream>>>(A, (const double2*)v, (double2*)v + c->nn_owned);
			c->launches++;
			return BP_OK;
		}
		bp_launch_pack_halo(c, v);
		BP_NCCL(c, g_nccl.GroupStart());
		for (int k = 0; k < c->n_nbr; ++k) {
			const int64_t ns = c->send_ptr[k + 1] - c->send_ptr[k], nr = c->recv_ptr[k + 1] - c->recv_ptr[k];
			if (ns) BP_NCCL(c, g_nccl.Send(c->d_sendbuf + 2 * c->send_ptr[k], (size_t)(2 * ns), ncclFloat64_, c->nbr_rank[k], c->comm->comm, c->stream));
			if (nr) BP_NCCL(c, g_nccl.Recv(v + 2 * (c->nn_owned + c->recv_ptr[k]), (size_t)(2 * nr), ncclFloat64_, c->nbr_rank[k], c->comm->comm, c->stream));
		}
		BP_NCCL(c, g_nccl.GroupEnd());
		return BP_OK;
	}
	// in-process group on one device: all contexts run on ctxs[0]'s stream
	cudaStream_t s = ctxs[0]->stream;
	for (int r = 0; r < n; ++r) bp_launch_pack_halo(ctxs[r], bp_vec(ctxs[r], which));
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = ctxs[r];
		for (int k = 0; k < c->n_nbr; ++k) {
			const int src = c->nbr_rank[k];
			if (src < 0 || src >= n) return bp_fail(c, BP_ERR_ARG, "neighbour rank %d outside the group", src);
			bp_ctx* o = ctxs[src];
			int kk = -1;
			for (int j = 0; j < o->n_nbr; ++j) if (o->nbr_rank[j] == r) kk = j;
			const int64_t nr = c->recv_ptr[k + 1] - c->recv_ptr[k];
			if (kk < 0 || o->send_ptr[kk + 1] - o->send_ptr[kk] != nr)
				return bp_fail(c, BP_ERR_ARG, "halo lists of ranks %d and %d do not match", r, src);
			if (nr) BP_CUDA(c, cudaMemcpyAsync(bp_vec(c, which) + 2 * (c->nn_owned + c->recv_ptr[k]), o->d_sendbuf + 2 * o->send_ptr[kk],
			                                   (size_t)(2 * nr) * sizeof(double), cudaMemcpyDeviceToDevice, s));
		}
	}
	return BP_OK;
}

int bp_comm_allreduce(bp_ctx** ctxs, int n, int slot, int count)
{
	if (n == 1) {
		bp_ctx* c = ctxs[0];
		if (!c->comm || c->comm->n_ranks == 1) return BP_OK;
		if (c->comm->p2p && count <= P2P_AR_DOUBLES) {
			k_allreduce_p2p<<<1, 32, 0, c->stream>>>(c->comm->d_peer, c->comm->rank, c->comm->n_ranks, c->comm->d_ar_seq, c->d_sc + slot, count, 0);
			c->launches++;
			return BP_OK;
		}
		BP_NCCL(c, g_nccl.AllReduce(c->d_sc + slot, c->d_sc + slot, (size_t)count, ncclFloat64_, ncclSum_, c->comm->comm, c->stream));
		return BP_OK;
	}
	bp_launch_sum_scalars(ctxs, n, slot, count);
	return BP_OK;
}

int bp_comm_allreduce_max(bp_ctx* c, int slot, int count)
{
	if (!c->comm || c->comm->n_ranks == 1) return BP_OK;
	if (c->comm->p2p && count <= P2P_AR_DOUBLES) {
		k_allreduce_p2p<<<1, 32, 0, c->stream>>>(c->comm->d_peer, c->comm->rank, c->comm->n_ranks, c->comm->d_ar_seq, c->d_sc + slot, count, 1);
		c->launches++;
		return BP_OK;
	}
	BP_NCCL(c, g_nccl.AllReduce(c->d_sc + slot, c->d_sc + slot, (size_t)count, ncclFloat64_, ncclMax_, c->comm->comm, c->stream));
	return BP_OK;
}

int bp_comm_rank(const bp_ctx* c) { return c->comm ? c->comm->rank : 0; }
int bp_comm_size(const bp_ctx* c) { return c->comm ? c->comm->n_ranks : 1; }

struct BufPtrs { double* p[16]; };
__global__ void k_allreduce_bufs(BufPtrs P, int n, size_t count, int op_max)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
		double s = P.p[0][i];
		for (int r = 1; r < n; ++r) { const double v = P.p[r][i]; s = op_max ? fmax(s, v) : s + v; }
		for (int r = 0; r < n; ++r) P.p[r][i] = s;
	}
}

int bp_comm_allreduce_buf(bp_ctx** ctxs, int n, double** bufs, size_t count, int op_max)
{
	if (count == 0) return BP_OK;
	if (n == 1) {
		bp_ctx* c = ctxs[0];
		if (!c->comm || c->comm->n_ranks == 1) return BP_OK;
		BP_NCCL(c, g_nccl.AllReduce(bufs[0], bufs[0], count, ncclFloat64_, op_max ? ncclMax_ : ncclSum_, c->comm->comm, c->stream));
		return BP_OK;
	}
	if (n > 16) return bp_fail(ctxs[0], BP_ERR_ARG, "in-process groups are limited to 16 contexts");
	BufPtrs P;
	for (int r = 0; r < n; ++r) P.p[r] = bufs[r];
	const int g = (int)std::min<size_t>((count + 255) / 256, 1184);
	k_allreduce_bufs<<<g, 256, 0, ctxs[0]->stream>>>(P, n, count, op_max);
	ctxs[0]->launches++;
	return BP_OK;
}
