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
// Several contexts in ONE process (bp_*_group) exchange by device copies instead;
// that is how the partitioned path is exercised on a single GPU.
#include "bp_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <algorithm>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclFloat64_ = 8 };   // ncclDataType_t
enum { ncclSum_ = 0, ncclMax_ = 2 };   // ncclRedOp_t

struct nccl_api {
	void* handle = nullptr;
	int (*GetUniqueId)(ncclUniqueId*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
static nccl_api g_nccl;

struct bp_comm_state {
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1;
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
	SYM(AllReduce, "ncclAllReduce"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
	SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
	g_nccl.handle = h;
	return BP_OK;
}
#define BP_NCCL(ctx, call) do { int r_ = (call); if (r_ != ncclSuccess_) \
	return bp_fail(ctx, BP_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

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
	return BP_OK;
}

void bp_comm_free(bp_ctx* c)
{
	if (c->comm) {
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
		BP_NCCL(c, g_nccl.AllReduce(c->d_sc + slot, c->d_sc + slot, (size_t)count, ncclFloat64_, ncclSum_, c->comm->comm, c->stream));
		return BP_OK;
	}
	bp_launch_sum_scalars(ctxs, n, slot, count);
	return BP_OK;
}

int bp_comm_allreduce_max(bp_ctx* c, int slot, int count)
{
	if (!c->comm || c->comm->n_ranks == 1) return BP_OK;
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
