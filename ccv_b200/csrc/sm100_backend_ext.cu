// sm100_backend_ext.cu -- command layer for attention, layer / rms norm, upsample and the gradient allreduce.
#include "../../include/ccv_nnc_sm100.h"
#include "sm100_contract.h"
#include "sm100_ew.h"
#include <cuda_runtime.h>
#include <string.h>

#define SM100_EXEC_ARGS const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context

extern "C" {

int ccv_nnc_sm100_exec_sdpa_forw(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_sdpa_back(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_lnorm_forw(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_lnorm_back(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_rmsnorm_forw(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_rmsnorm_back(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_upsample_forw(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_upsample_back(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_allreduce(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }

}
