// gdsp_comm.cu -- the multi-GPU plumbing of the operator library: NCCL under the C boundary.
//
// BASELINE north_star: "windowed operators exchange halos of the window radius with neighbouring slabs over
// NVLink with NCCL send/recv, while percentile histograms and named variables use NCCL allreduce/broadcast."
// A communicator is bound to one gdsp_ctx (one GPU, one stream); every call is enqueued on that stream, so a
// halo exchange is ordered with the kernels before and after it without a host synchronisation.  Two ways to
// build the communicators: one process per GPU (gdsp_comm_unique_id on rank 0, the 128 bytes travel by whatever
// the launcher has -- torchrun's store, MPI, a file -- then gdsp_comm_create on every rank), or one process
// driving all GPUs of a box (gdsp_comm_create_all = ncclCommInitAll; what a multi-GPU C host would use).
#include <nccl.h>
#include <vector>
#include "gdsp_common.cuh"

struct gdsp_comm
	{
	gdsp_ctx*  ctx;
	ncclComm_t nccl;
	int        rank, size;
	};

#define GDSP_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
	gdsp_set_error ("%s: %s", #call, ncclGetErrorString (r_));  return GDSP_ERR_CUDA; } } while (0)

extern "C" int gdsp_comm_unique_id (unsigned char* id128)
	{
	GDSP_REQUIRE (id128 != NULL, "gdsp_comm_unique_id: NULL argument");
	static_assert (sizeof (ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	ncclUniqueId id;
	GDSP_NCCL (ncclGetUniqueId (&id));
	memcpy (id128, &id, 128);
	return GDSP_OK;
	}

extern "C" int gdsp_comm_create (gdsp_ctx* c, const unsigned char* id128, int nranks, int rank, gdsp_comm** out)
	{
	GDSP_REQUIRE (c && id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "gdsp_comm_create: bad argument");
	ncclUniqueId id;
	memcpy (&id, id128, 128);
	GDSP_CUDA (cudaSetDevice (c->device));
	gdsp_comm* m = new gdsp_comm ();
	m->ctx = c;  m->rank = rank;  m->size = nranks;  m->nccl = NULL;
	ncclResult_t r = ncclCommInitRank (&m->nccl, nranks, id, rank);
	if (r != ncclSuccess) { gdsp_set_error ("ncclCommInitRank: %s", ncclGetErrorString (r));  delete m;  return GDSP_ERR_CUDA; }
	*out = m;
	return GDSP_OK;
	}

extern "C" int gdsp_comm_create_all (gdsp_ctx** ctxs, int n, gdsp_comm** out)
	{
	GDSP_REQUIRE (ctxs && out && n >= 1, "gdsp_comm_create_all: bad argument");
	std::vector<int> devs (n);
	for (int i = 0; i < n; i++) { GDSP_REQUIRE (ctxs[i] != NULL, "gdsp_comm_create_all: NULL context");  devs[i] = ctxs[i]->device; }
	std::vector<ncclComm_t> comms (n);
	GDSP_NCCL (ncclCommInitAll (comms.data (), n, devs.data ()));
	for (int i = 0; i < n; i++)
		{
		gdsp_comm* m = new gdsp_comm ();
		m->ctx = ctxs[i];  m->rank = i;  m->size = n;  m->nccl = comms[i];
		out[i] = m;
		}
	return GDSP_OK;
	}

extern "C" void gdsp_comm_destroy (gdsp_comm* m)
	{
	if (m == NULL) return;
	if (m->nccl != NULL) { cudaSetDevice (m->ctx->device);  cudaStreamSynchronize (m->ctx->stream);  ncclCommDestroy (m->nccl); }
	delete m;
	}

extern "C" int gdsp_comm_rank (const gdsp_comm* m) { return m ? m->rank : -1; }
extern "C" int gdsp_comm_size (const gdsp_comm* m) { return m ? m->size : 0; }

// one grouped send/recv round with the slab neighbours; all of it on the context's stream
extern "C" int gdsp_comm_exchange_halos (gdsp_comm* m, double* sig, const gdsp_halo* plan, int nplan)
	{
	GDSP_REQUIRE (m && sig && (plan || nplan == 0) && nplan >= 0, "gdsp_comm_exchange_halos: bad argument");
	if (nplan == 0) return GDSP_OK;
	GDSP_CUDA (cudaSetDevice (m->ctx->device));
	GDSP_NCCL (ncclGroupStart ());
	for (int k = 0; k < nplan; k++)
		{
		const gdsp_halo& h = plan[k];
		GDSP_REQUIRE (h.peer >= 0 && h.peer < m->size && h.peer != m->rank && h.send_lo <= h.send_hi && h.recv_lo <= h.recv_hi,
		              "gdsp_comm_exchange_halos: malformed plan entry %d", k);
		if (h.send_hi > h.send_lo) GDSP_NCCL (ncclSend (sig + h.send_lo, h.send_hi - h.send_lo, ncclDouble, h.peer, m->nccl, m->ctx->stream));
		if (h.recv_hi > h.recv_lo) GDSP_NCCL (ncclRecv (sig + h.recv_lo, h.recv_hi - h.recv_lo, ncclDouble, h.peer, m->nccl, m->ctx->stream));
		}
	GDSP_NCCL (ncclGroupEnd ());
	return GDSP_OK;
	}

// the multi-device form for a single-process host: one grouped round over all communicators
extern "C" int gdsp_comm_exchange_halos_all (gdsp_comm** ms, double** sigs, const gdsp_halo* const* plans, const int* nplans, int n)
	{
	GDSP_REQUIRE (ms && sigs && plans && nplans && n >= 1, "gdsp_comm_exchange_halos_all: bad argument");
	GDSP_NCCL (ncclGroupStart ());
	for (int d = 0; d < n; d++)
		for (int k = 0; k < nplans[d]; k++)
			{
			const gdsp_halo& h = plans[d][k];
			gdsp_comm* m = ms[d];
			if (h.send_hi > h.send_lo) GDSP_NCCL (ncclSend (sigs[d] + h.send_lo, h.send_hi - h.send_lo, ncclDouble, h.peer, m->nccl, m->ctx->stream));
			if (h.recv_hi > h.recv_lo) GDSP_NCCL (ncclRecv (sigs[d] + h.recv_lo, h.recv_hi - h.recv_lo, ncclDouble, h.peer, m->nccl, m->ctx->stream));
			}
	GDSP_NCCL (ncclGroupEnd ());
	return GDSP_OK;
	}

// element-wise sum over ranks of n 64-bit counts given and returned on the HOST (percentile region counts, run
// counts, flags): staged through a small device buffer, one ncclAllReduce, synchronises the stream
extern "C" int gdsp_comm_allreduce_sum_u64 (gdsp_comm* m, uint64_t* h_values, int n)
	{
	GDSP_REQUIRE (m && h_values && n >= 1, "gdsp_comm_allreduce_sum_u64: bad argument");
	gdsp_ctx* c = m->ctx;
	GDSP_CUDA (cudaSetDevice (c->device));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 7, sizeof (uint64_t) * (size_t) n, &ws));
	GDSP_CUDA (cudaMemcpyAsync (ws, h_values, sizeof (uint64_t) * (size_t) n, cudaMemcpyHostToDevice, c->stream));
	GDSP_NCCL (ncclAllReduce (ws, ws, (size_t) n, ncclUint64, ncclSum, m->nccl, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (h_values, ws, sizeof (uint64_t) * (size_t) n, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

// n doubles from every rank, host to host (piece carries, per-piece records, min/max): h_out holds size*n values in rank order
extern "C" int gdsp_comm_allgather_f64 (gdsp_comm* m, const double* h_in, int n, double* h_out)
	{
	GDSP_REQUIRE (m && h_in && h_out && n >= 1, "gdsp_comm_allgather_f64: bad argument");
	gdsp_ctx* c = m->ctx;
	GDSP_CUDA (cudaSetDevice (c->device));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 7, sizeof (double) * (size_t) n * (size_t) (m->size + 1), &ws));
	double* d_in = (double*) ws;  double* d_out = d_in + n;
	GDSP_CUDA (cudaMemcpyAsync (d_in, h_in, sizeof (double) * (size_t) n, cudaMemcpyHostToDevice, c->stream));
	GDSP_NCCL (ncclAllGather (d_in, d_out, (size_t) n, ncclDouble, m->nccl, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (h_out, d_out, sizeof (double) * (size_t) n * (size_t) m->size, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

// n doubles from every rank, device to device (percentile samples and candidates); asynchronous on the stream
extern "C" int gdsp_comm_allgather_dev (gdsp_comm* m, const double* d_in, uint64_t n, double* d_out)
	{
	GDSP_REQUIRE (m && d_in && d_out, "gdsp_comm_allgather_dev: bad argument");
	if (n == 0) return GDSP_OK;
	GDSP_CUDA (cudaSetDevice (m->ctx->device));
	GDSP_NCCL (ncclAllGather (d_in, d_out, (size_t) n, ncclDouble, m->nccl, m->ctx->stream));
	return GDSP_OK;
	}

// root's n doubles to everyone, host to host (named variables, thresholds)
extern "C" int gdsp_comm_broadcast_f64 (gdsp_comm* m, double* h_values, int n, int root)
	{
	GDSP_REQUIRE (m && h_values && n >= 1 && root >= 0 && root < m->size, "gdsp_comm_broadcast_f64: bad argument");
	gdsp_ctx* c = m->ctx;
	GDSP_CUDA (cudaSetDevice (c->device));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 7, sizeof (double) * (size_t) n, &ws));
	if (m->rank == root) GDSP_CUDA (cudaMemcpyAsync (ws, h_values, sizeof (double) * (size_t) n, cudaMemcpyHostToDevice, c->stream));
	GDSP_NCCL (ncclBroadcast (ws, ws, (size_t) n, ncclDouble, root, m->nccl, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (h_values, ws, sizeof (double) * (size_t) n, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}
