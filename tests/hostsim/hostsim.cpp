// TEST INFRASTRUCTURE ONLY -- storage for the SIMT emulator's per-thread context (see cuda_shim.h).
#include "cuda_shim.h"
namespace hostsim {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local Ctx t_ctx;
}
