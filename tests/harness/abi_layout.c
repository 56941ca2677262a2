/*
 * abi_layout.c - compile-time proof that the driver/NVML types restated in
 * vgpu_manager_b200/csrc/vgpu_internal.h are layout-identical to the real CUDA 12.9 headers
 * (the reference guards its own vendored subset the same way: library/hack/check_struct_layout.py).
 * Compiling this file IS the test (tests/test_abi_surface.py).
 */
#include <cuda.h>
#include <nvml.h>
#include <stddef.h>

/* keep the toolkit's values before our names shadow them */
enum {
  real_SUCCESS = CUDA_SUCCESS, real_INVALID_VALUE = CUDA_ERROR_INVALID_VALUE, real_OOM = CUDA_ERROR_OUT_OF_MEMORY,
  real_NOT_INIT = CUDA_ERROR_NOT_INITIALIZED, real_INVALID_CTX = CUDA_ERROR_INVALID_CONTEXT,
  real_NOT_FOUND = CUDA_ERROR_NOT_FOUND, real_NOT_READY = CUDA_ERROR_NOT_READY, real_NOT_SUPPORTED = CUDA_ERROR_NOT_SUPPORTED,
  real_NV_SUCCESS = NVML_SUCCESS, real_NV_NOT_SUPPORTED = NVML_ERROR_NOT_SUPPORTED, real_NV_NOT_FOUND = NVML_ERROR_NOT_FOUND,
  real_NV_FN_NOT_FOUND = NVML_ERROR_FUNCTION_NOT_FOUND,
  real_ATTACH_GLOBAL = CU_MEM_ATTACH_GLOBAL, real_NONBLOCK = CU_STREAM_NON_BLOCKING, real_PORTABLE = CU_MEMHOSTALLOC_PORTABLE,
  real_DEVICEMAP = CU_MEMHOSTALLOC_DEVICEMAP, real_SMS = CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT,
  real_THR = CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_MULTIPROCESSOR, real_PTDS = CU_GET_PROC_ADDRESS_PER_THREAD_DEFAULT_STREAM,
  real_GEQ = CU_STREAM_WAIT_VALUE_GEQ, real_MEMOPS64 = CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS,
  real_SMEM_ATTR = CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, real_LOC_DEVICE = CU_MEM_LOCATION_TYPE_DEVICE,
  real_CAP_RELAXED = CU_STREAM_CAPTURE_MODE_RELAXED, real_NODE_KERNEL = CU_GRAPH_NODE_TYPE_KERNEL, real_NODE_GRAPH = CU_GRAPH_NODE_TYPE_GRAPH,
  /* graph cost reads gridDimX/Y/Z as 32-bit words 2,3,4 of either kernel-node parameter struct */
  real_KNP1_GX = offsetof(CUDA_KERNEL_NODE_PARAMS_v1, gridDimX), real_KNP1_GZ = offsetof(CUDA_KERNEL_NODE_PARAMS_v1, gridDimZ),
  real_KNP2_GX = offsetof(CUDA_KERNEL_NODE_PARAMS_v2, gridDimX), real_KNP2_GZ = offsetof(CUDA_KERNEL_NODE_PARAMS_v2, gridDimZ),
  real_KNP2_SIZE = sizeof(CUDA_KERNEL_NODE_PARAMS_v2),
};
typedef CUDA_ARRAY_DESCRIPTOR real_arr2;
typedef CUDA_ARRAY3D_DESCRIPTOR real_arr3;
typedef CUmemAllocationProp real_prop;
typedef CUlaunchConfig real_cfg;
typedef CUuuid real_uuid;
typedef CUdeviceptr real_dptr;
typedef nvmlMemory_t real_nvmem;
typedef nvmlMemory_v2_t real_nvmem2;
typedef nvmlUtilization_t real_nvutil;
typedef nvmlProcessInfo_v1_t real_proc1;
typedef nvmlProcessUtilizationSample_t real_sample;

#define CUresult vg_CUresult
#define CUdevice vg_CUdevice
#define CUdeviceptr vg_CUdeviceptr
#define cuuint64_t vg_cuuint64_t
#define CUcontext vg_CUcontext
#define CUstream vg_CUstream
#define CUfunction vg_CUfunction
#define CUmodule vg_CUmodule
#define CUarray vg_CUarray
#define CUmipmappedArray vg_CUmipmappedArray
#define CUmemoryPool vg_CUmemoryPool
#define CUevent vg_CUevent
#define CUgraph vg_CUgraph
#define CUgraphNode vg_CUgraphNode
#define CUgraphExec vg_CUgraphExec
#define CUmemGenericAllocationHandle vg_CUmemGenericAllocationHandle
#define CUuuid vg_CUuuid
#define nvmlReturn_t vg_nvmlReturn_t
#define nvmlDevice_t vg_nvmlDevice_t
#define CUDA_SUCCESS vg_CUDA_SUCCESS
#define CUDA_ERROR_INVALID_VALUE vg_CUDA_ERROR_INVALID_VALUE
#define CUDA_ERROR_OUT_OF_MEMORY vg_CUDA_ERROR_OUT_OF_MEMORY
#define CUDA_ERROR_NOT_INITIALIZED vg_CUDA_ERROR_NOT_INITIALIZED
#define CUDA_ERROR_INVALID_CONTEXT vg_CUDA_ERROR_INVALID_CONTEXT
#define CUDA_ERROR_NOT_FOUND vg_CUDA_ERROR_NOT_FOUND
#define CUDA_ERROR_NOT_READY vg_CUDA_ERROR_NOT_READY
#define CUDA_ERROR_NOT_SUPPORTED vg_CUDA_ERROR_NOT_SUPPORTED
#define NVML_SUCCESS vg_NVML_SUCCESS
#define NVML_ERROR_NOT_SUPPORTED vg_NVML_ERROR_NOT_SUPPORTED
#define NVML_ERROR_NOT_FOUND vg_NVML_ERROR_NOT_FOUND
#define NVML_ERROR_FUNCTION_NOT_FOUND vg_NVML_ERROR_FUNCTION_NOT_FOUND
#include "../../vgpu_manager_b200/csrc/cu_abi_subset.h"

#define SAME(a, b) _Static_assert((a) == (b), #a " != " #b)
SAME(vg_CUDA_SUCCESS, real_SUCCESS); SAME(vg_CUDA_ERROR_INVALID_VALUE, real_INVALID_VALUE);
SAME(vg_CUDA_ERROR_OUT_OF_MEMORY, real_OOM); SAME(vg_CUDA_ERROR_NOT_INITIALIZED, real_NOT_INIT);
SAME(vg_CUDA_ERROR_INVALID_CONTEXT, real_INVALID_CTX); SAME(vg_CUDA_ERROR_NOT_FOUND, real_NOT_FOUND);
SAME(vg_CUDA_ERROR_NOT_READY, real_NOT_READY); SAME(vg_CUDA_ERROR_NOT_SUPPORTED, real_NOT_SUPPORTED);
SAME(vg_NVML_SUCCESS, real_NV_SUCCESS); SAME(vg_NVML_ERROR_NOT_SUPPORTED, real_NV_NOT_SUPPORTED);
SAME(vg_NVML_ERROR_NOT_FOUND, real_NV_NOT_FOUND); SAME(vg_NVML_ERROR_FUNCTION_NOT_FOUND, real_NV_FN_NOT_FOUND);
SAME(VCU_MEM_ATTACH_GLOBAL, real_ATTACH_GLOBAL); SAME(VCU_STREAM_NON_BLOCKING, real_NONBLOCK);
SAME(VCU_MEMHOSTALLOC_PORTABLE, real_PORTABLE); SAME(VCU_MEMHOSTALLOC_DEVICEMAP, real_DEVICEMAP);
SAME(VCU_ATTR_SM_COUNT, real_SMS); SAME(VCU_ATTR_MAX_THREADS_PER_SM, real_THR); SAME(VCU_GET_PROC_PTDS, real_PTDS);
SAME(VCU_WAIT_GEQ, real_GEQ); SAME(122, real_MEMOPS64); SAME(8, real_SMEM_ATTR); SAME(1, real_LOC_DEVICE);

SAME(sizeof(vcu_array_desc_t), sizeof(real_arr2));
SAME(offsetof(vcu_array_desc_t, Format), offsetof(real_arr2, Format));
SAME(offsetof(vcu_array_desc_t, NumChannels), offsetof(real_arr2, NumChannels));
SAME(sizeof(vcu_array3d_desc_t), sizeof(real_arr3));
SAME(offsetof(vcu_array3d_desc_t, Depth), offsetof(real_arr3, Depth));
SAME(offsetof(vcu_array3d_desc_t, Format), offsetof(real_arr3, Format));
SAME(offsetof(vcu_array3d_desc_t, NumChannels), offsetof(real_arr3, NumChannels));
SAME(sizeof(vcu_mem_alloc_prop_t), sizeof(real_prop));
SAME(offsetof(vcu_mem_alloc_prop_t, location), offsetof(real_prop, location));
SAME(sizeof(vcu_launch_config_t), sizeof(real_cfg));
SAME(offsetof(vcu_launch_config_t, gridDimZ), offsetof(real_cfg, gridDimZ));
SAME(offsetof(vcu_launch_config_t, hStream), offsetof(real_cfg, hStream));
SAME(sizeof(vg_CUuuid), sizeof(real_uuid));
SAME(sizeof(vg_CUdeviceptr), sizeof(real_dptr));
SAME(sizeof(vnv_memory_t), sizeof(real_nvmem));
SAME(offsetof(vnv_memory_t, free), offsetof(real_nvmem, free));
SAME(offsetof(vnv_memory_t, used), offsetof(real_nvmem, used));
SAME(sizeof(vnv_memory_v2_t), sizeof(real_nvmem2));
SAME(offsetof(vnv_memory_v2_t, reserved), offsetof(real_nvmem2, reserved));
SAME(offsetof(vnv_memory_v2_t, used), offsetof(real_nvmem2, used));
SAME(sizeof(vnv_utilization_t), sizeof(real_nvutil));
SAME(sizeof(vgpu_proc_t), sizeof(real_proc1));
SAME(offsetof(vgpu_proc_t, used_bytes), offsetof(real_proc1, usedGpuMemory));
SAME(sizeof(vgpu_util_sample_t), sizeof(real_sample));
SAME(offsetof(vgpu_util_sample_t, ts_us), offsetof(real_sample, timeStamp));
SAME(offsetof(vgpu_util_sample_t, sm), offsetof(real_sample, smUtil));
SAME(offsetof(vgpu_util_sample_t, dec), offsetof(real_sample, decUtil));
SAME(VCU_STREAM_CAPTURE_MODE_RELAXED, real_CAP_RELAXED); SAME(VCU_GRAPH_NODE_KERNEL, real_NODE_KERNEL); SAME(VCU_GRAPH_NODE_GRAPH, real_NODE_GRAPH);
SAME(real_KNP1_GX, 8); SAME(real_KNP1_GZ, 16); SAME(real_KNP2_GX, 8); SAME(real_KNP2_GZ, 16);
_Static_assert(real_KNP2_SIZE <= 128, "graph_cost() reads kernel-node parameters into a 128-byte buffer");
int abi_layout_ok = 1;
