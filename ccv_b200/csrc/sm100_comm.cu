// sm100_comm.cu -- CCV_NNC_COMM_ALLREDUCE_FORWARD / _BACKWARD of CCV_NNC_BACKEND_GPU_SM100 and the communicator
// plumbing behind it.  Replaces lib/nnc/cmd/comm/gpu/ccv_nnc_comm_gpu_nccl.cu:12-58 (the exec) and
// lib/nnc/gpu/ccv_nnc_compat.cu:1385-1445 (ccv_nnc_nccl_get_comm).
//
// Two deployment shapes, chosen by where the tensors live:
//   * one process per GPU (how bench.py / torchrun drive this backend): every tensor of the call lives on this
//     process's device; the communicator was bound with ccv_nnc_sm100_comm_init_rank(id, world, rank) (the 128-byte id
//     comes from rank 0's ccv_nnc_sm100_comm_unique_id and travels over whatever side channel the host has); each
//     tensor is sum-reduced across the ranks, all tensors of one call inside one NCCL group = one fused launch.
//   * one process, P devices (the reference's shape, ccv_nnc_comm_gpu_nccl.cu:17-47): tensor i lives on device i, the
//     communicators come from ncclCommInitAll(P) created lazily, tensor i is enqueued on the stream of
//     ccv_nnc_stream_context_find_neighbor(stream, device i).
// NCCL is resolved with dlopen at first use so that the library loads (and every other command runs) without it.
#include "../../include/ccv_nnc_sm100.h"
#include "sm100_contract.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace {

// the slice of nccl.h this file uses (stable since NCCL 2.0): opaque communicator, 128-byte unique id, enums
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT16 = 6, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_BFLOAT16 = 9 };
enum { NCCL_SUM = 0 };

struct Nccl {
	void* handle = 0;
	int (*GetUniqueId)(ncclUniqueId*) = 0;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = 0;
	int (*CommInitAll)(ncclComm_t*, int, const int*) = 0;
	int (*CommDestroy)(ncclComm_t) = 0;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = 0;
	int (*GroupStart)(void) = 0;
	int (*GroupEnd)(void) = 0;
	const char* (*GetErrorString)(int) = 0;
};

Nccl g_nccl;
std::mutex g_mutex;
ncclComm_t g_rank_comm = 0; // one process per GPU
int g_rank = -1, g_world = 0, g_rank_device = -1;
std::vector<ncclComm_t> g_all_comms; // one process, P devices

int load_nccl()
{
	if (g_nccl.handle)
		return 0;
	// an already loaded libnccl.so.2 (e.g. the one torch.distributed brought in) is reused by soname
	const char* const names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char* n : names)
		if ((g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
			break;
	if (!g_nccl.handle)
	{
		sm100::set_last_error("dlopen(libnccl.so.2) failed", cudaErrorUnknown);
		return -1;
	}
#define SM100_NCCL_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name); if (!g_nccl.field) { sm100::set_last_error("missing NCCL symbol " name, cudaErrorUnknown); g_nccl.handle = 0; return -1; }
	SM100_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
	SM100_NCCL_SYM(CommInitRank, "ncclCommInitRank")
	SM100_NCCL_SYM(CommInitAll, "ncclCommInitAll")
	SM100_NCCL_SYM(CommDestroy, "ncclCommDestroy")
	SM100_NCCL_SYM(AllReduce, "ncclAllReduce")
	SM100_NCCL_SYM(GroupStart, "ncclGroupStart")
	SM100_NCCL_SYM(GroupEnd, "ncclGroupEnd")
	SM100_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SM100_NCCL_SYM
	return 0;
}

int nccl_fail(const char* const what, const int rc)
{
	static char msg[256];
	snprintf(msg, sizeof(msg), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
	sm100::set_last_error(msg, cudaErrorUnknown);
	return CCV_NNC_EXEC_INVALID;
}

int nccl_datatype(const int datatype)
{
	switch (CCV_GET_DATA_TYPE(datatype))
	{
		case CCV_8U: return NCCL_UINT8;
		case CCV_32S: return NCCL_INT32;
		case CCV_64S: return NCCL_INT64;
		case CCV_16F: return NCCL_FLOAT16;
		case CCV_16BF: return NCCL_BFLOAT16;
		case CCV_32F: return NCCL_FLOAT32;
		case CCV_64F: return NCCL_FLOAT64;
	}
	return -1;
}

size_t datatype_bytes(const int datatype)
{
	switch (CCV_GET_DATA_TYPE(datatype))
	{
		case CCV_8U: return 1;
		case CCV_16F: case CCV_16BF: return 2;
		case CCV_64S: case CCV_64F: return 8;
	}
	return 4;
}

size_t tensor_count(const ccv_nnc_tensor_t* const t)
{
	size_t n = 1;
	for (int i = 0; i < CCV_NNC_MAX_DIM_ALLOC && t->info.dim[i] > 0; i++)
		n *= (size_t)t->info.dim[i];
	return n;
}

}

extern "C" {

int ccv_nnc_sm100_comm_unique_id(void* const id, const size_t size)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (size < sizeof(ncclUniqueId) || load_nccl())
		return -1;
	ncclUniqueId u;
	const int rc = g_nccl.GetUniqueId(&u);
	if (rc != NCCL_SUCCESS)
		return nccl_fail("ncclGetUniqueId", rc);
	memcpy(id, &u, sizeof(u));
	return 0;
}

int ccv_nnc_sm100_comm_init_rank(const void* const id, const size_t size, const int world, const int rank)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (size < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world || load_nccl())
		return -1;
	if (g_rank_comm)
	{
		g_nccl.CommDestroy(g_rank_comm);
		g_rank_comm = 0;
	}
	ncclUniqueId u;
	memcpy(&u, id, sizeof(u));
	cudaGetDevice(&g_rank_device);
	const int rc = g_nccl.CommInitRank(&g_rank_comm, world, u, rank);
	if (rc != NCCL_SUCCESS)
	{
		g_rank_comm = 0;
		return nccl_fail("ncclCommInitRank", rc);
	}
	g_rank = rank, g_world = world;
	return 0;
}

int ccv_nnc_sm100_comm_rank(void) { return g_rank; }
int ccv_nnc_sm100_comm_world(void) { return g_world; }

void ccv_nnc_sm100_comm_destroy(void)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (!g_nccl.handle)
		return;
	if (g_rank_comm)
		g_nccl.CommDestroy(g_rank_comm);
	g_rank_comm = 0, g_rank = -1, g_world = 0;
	for (ncclComm_t c : g_all_comms)
		g_nccl.CommDestroy(c);
	g_all_comms.clear();
}

// lib/nnc/cmd/comm/gpu/ccv_nnc_comm_gpu_nccl.cu:12-58.  Forward and backward of an allreduce are the same sum.
int ccv_nnc_sm100_exec_allreduce(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	const int count = input_size < output_size ? input_size : output_size;
	if (count < 1)
		return CCV_NNC_EXEC_INVALID;
	int device_count = 0, same_device = 1;
	const int first_device = CCV_TENSOR_GET_DEVICE_ID(inputs[0]->info.type);
	for (int i = 0; i < count; i++)
	{
		if (!inputs[i] || !outputs[i] || CCV_IS_TENSOR_VIEW(inputs[i]) || CCV_IS_TENSOR_VIEW(outputs[i]))
			return CCV_NNC_EXEC_INVALID; // contiguous tensors only, as the reference asserts (:21-23)
		if (inputs[i]->info.datatype != outputs[i]->info.datatype || tensor_count(inputs[i]) != tensor_count(outputs[i]) || nccl_datatype(inputs[i]->info.datatype) < 0)
			return CCV_NNC_EXEC_INVALID;
		if (CCV_TENSOR_GET_MEMORY(inputs[i]->info.type) != CCV_TENSOR_GPU_MEMORY || CCV_TENSOR_GET_DEVICE_ID(inputs[i]->info.type) != CCV_TENSOR_GET_DEVICE_ID(outputs[i]->info.type))
			return CCV_NNC_EXEC_INVALID;
		const int device = CCV_TENSOR_GET_DEVICE_ID(inputs[i]->info.type);
		if (device != first_device)
			same_device = 0;
		if (device + 1 > device_count)
			device_count = device + 1;
	}
	std::lock_guard<std::mutex> lock(g_mutex);
	if (same_device && !g_rank_comm && device_count == 1)
	{
		// a single participant (device 0 only, no rank communicator bound): the sum over one replica is a copy
		cudaStream_t stream = (cudaStream_t)ccv_nnc_stream_context_get_stream(stream_context);
		for (int i = 0; i < count; i++)
			if (inputs[i]->data.u8 != outputs[i]->data.u8)
				cudaMemcpyAsync(outputs[i]->data.u8, inputs[i]->data.u8, tensor_count(inputs[i]) * datatype_bytes(inputs[i]->info.datatype), cudaMemcpyDeviceToDevice, stream);
		return CCV_NNC_EXEC_SUCCESS;
	}
	if (load_nccl())
		return CCV_NNC_EXEC_NO_KERNEL;
	int rc;
	if (same_device && g_rank_comm)
	{
		// one process per GPU: reduce every tensor across the ranks, one group = one fused NCCL launch
		cudaStream_t stream = (cudaStream_t)ccv_nnc_stream_context_get_stream(stream_context);
		if ((rc = g_nccl.GroupStart()) != NCCL_SUCCESS)
			return nccl_fail("ncclGroupStart", rc);
		for (int i = 0; i < count; i++)
			if ((rc = g_nccl.AllReduce(inputs[i]->data.u8, outputs[i]->data.u8, tensor_count(inputs[i]), nccl_datatype(inputs[i]->info.datatype), NCCL_SUM, g_rank_comm, stream)) != NCCL_SUCCESS)
			{
				g_nccl.GroupEnd(); // never leave the group open
				return nccl_fail("ncclAllReduce", rc);
			}
		if ((rc = g_nccl.GroupEnd()) != NCCL_SUCCESS)
			return nccl_fail("ncclGroupEnd", rc);
		sm100::count_launch(1);
		return CCV_NNC_EXEC_SUCCESS;
	}
	// one process, P devices: tensor i on device i
	if ((int)g_all_comms.size() != device_count)
	{
		for (ncclComm_t c : g_all_comms)
			g_nccl.CommDestroy(c);
		g_all_comms.assign(device_count, (ncclComm_t)0);
		if ((rc = g_nccl.CommInitAll(g_all_comms.data(), device_count, 0)) != NCCL_SUCCESS)
		{
			g_all_comms.clear();
			return nccl_fail("ncclCommInitAll", rc);
		}
	}
	// Per-device streams: with a stream context, the neighbour the graph runner registered for that device
	// (comm/gpu/ccv_nnc_comm_gpu_nccl.cu:43-44) -- a missing neighbour is an error, not a silent fall to another stream, because
	// nothing would order the reduction with the kernels that produced the gradients there.  Without a stream context
	// (synchronous form), each device's legacy default stream, which is what this host's per-thread default contexts wrap.
	std::vector<cudaStream_t> streams(count, (cudaStream_t)0);
	if (stream_context)
		for (int i = 0; i < count; i++)
		{
			ccv_nnc_stream_context_t* const neighbor = ccv_nnc_stream_context_find_neighbor(stream_context, CCV_TENSOR_GET_DEVICE_ID(inputs[i]->info.type));
			if (!neighbor)
				return CCV_NNC_EXEC_INVALID;
			streams[i] = (cudaStream_t)ccv_nnc_stream_context_get_stream(neighbor);
		}
	int restore = 0;
	cudaGetDevice(&restore);
	if ((rc = g_nccl.GroupStart()) != NCCL_SUCCESS)
		return nccl_fail("ncclGroupStart", rc);
	for (int i = 0; i < count; i++)
	{
		const int device = CCV_TENSOR_GET_DEVICE_ID(inputs[i]->info.type);
		cudaSetDevice(device);
		if ((rc = g_nccl.AllReduce(inputs[i]->data.u8, outputs[i]->data.u8, tensor_count(inputs[i]), nccl_datatype(inputs[i]->info.datatype), NCCL_SUM, g_all_comms[device], streams[i])) != NCCL_SUCCESS)
		{
			g_nccl.GroupEnd();
			cudaSetDevice(restore);
			return nccl_fail("ncclAllReduce", rc);
		}
	}
	rc = g_nccl.GroupEnd();
	cudaSetDevice(restore);
	if (rc != NCCL_SUCCESS)
		return nccl_fail("ncclGroupEnd", rc);
	sm100::count_launch(1);
	return CCV_NNC_EXEC_SUCCESS;
}

}
