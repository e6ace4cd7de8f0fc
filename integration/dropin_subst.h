/* integration/dropin_subst.h -- force-included in front of the reference's UNMODIFIED test programs (integration/Makefile).
 * The tests name the backend they exercise (`cmd.backend = CCV_NNC_BACKEND_GPU_CUDNN`, `ccv_nnc_cmd_ok(.., CCV_NNC_BACKEND_GPU_REF)`
 * guards: test/int/nnc/cudnn.tests.c:24-36, cublas.tests.c, sgd.tests.c:16 ...).  After the reference's own headers have defined the
 * enum, every GPU backend name reads CCV_NNC_BACKEND_GPU_SM100, i.e. "the reference's tests with GPU_SM100 substituted". */
#include <ccv.h>
#include <nnc/ccv_nnc.h>
#define CCV_NNC_BACKEND_GPU_CUDNN CCV_NNC_BACKEND_GPU_SM100
#define CCV_NNC_BACKEND_GPU_CUBLAS CCV_NNC_BACKEND_GPU_SM100
#define CCV_NNC_BACKEND_GPU_REF CCV_NNC_BACKEND_GPU_SM100
#define CCV_NNC_BACKEND_GPU_NCCL CCV_NNC_BACKEND_GPU_SM100
