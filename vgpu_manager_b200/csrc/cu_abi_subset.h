/*
 * cu_abi_subset.h - the handful of CUDA driver / NVML types and constants the intercepted
 * entry points touch, restated so that the host shim never includes <cuda.h>/<nvml.h> (their
 * versioning macros would rename the very symbols this library must export).
 * tests/harness/abi_layout.c static-asserts every definition here against CUDA 12.9.
 */
#ifndef VGPU_CU_ABI_SUBSET_H
#define VGPU_CU_ABI_SUBSET_H
#include <stddef.h>
#include <stdint.h>
#include "../../include/vgpu_contract.h"

/* ------------------------------------------------------------------ driver ABI subset */
typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef unsigned long long cuuint64_t;
typedef void *CUcontext, *CUstream, *CUfunction, *CUmodule, *CUarray, *CUmipmappedArray,
    *CUmemoryPool, *CUevent, *CUgraph, *CUgraphNode, *CUgraphExec;
enum { VCU_GRAPH_NODE_KERNEL = 0, VCU_GRAPH_NODE_GRAPH = 4 }; /* CUgraphNodeType */
enum { VCU_STREAM_CAPTURE_MODE_RELAXED = 2 };                  /* CUstreamCaptureMode */
typedef unsigned long long CUmemGenericAllocationHandle;
typedef struct { char bytes[16]; } CUuuid;

enum {
  CUDA_SUCCESS = 0,
  CUDA_ERROR_INVALID_VALUE = 1,
  CUDA_ERROR_OUT_OF_MEMORY = 2,
  CUDA_ERROR_NOT_INITIALIZED = 3,
  CUDA_ERROR_INVALID_CONTEXT = 201,
  CUDA_ERROR_NOT_FOUND = 500,
  CUDA_ERROR_NOT_READY = 600,
  CUDA_ERROR_NOT_SUPPORTED = 801,
};

typedef struct {
  size_t Width, Height;
  int Format;
  unsigned int NumChannels;
} vcu_array_desc_t; /* CUDA_ARRAY_DESCRIPTOR_v2 */

typedef struct {
  size_t Width, Height, Depth;
  int Format;
  unsigned int NumChannels, Flags;
} vcu_array3d_desc_t; /* CUDA_ARRAY3D_DESCRIPTOR_v2 */

typedef struct {
  int type; /* CUmemAllocationType */
  int requestedHandleTypes;
  struct { int type; int id; } location; /* CUmemLocation; type 1 == DEVICE */
  void *win32HandleMetaData;
  struct { unsigned char compressionType, gpuDirectRDMACapable; unsigned short usage; unsigned char reserved[4]; } allocFlags;
} vcu_mem_alloc_prop_t; /* CUmemAllocationProp_v1 */

typedef struct {
  struct { int type; int id; } location; /* CUmemLocation */
  int flags;                              /* CUmemAccess_flags */
} vcu_mem_access_desc_t; /* CUmemAccessDesc_v1 */
#define VCU_MEM_LOCATION_DEVICE 1
#define VCU_MEM_LOCATION_HOST_NUMA 3
#define VCU_MEM_ALLOCATION_PINNED 1
#define VCU_MEM_ACCESS_READWRITE 3
#define VCU_MEM_GRANULARITY_MINIMUM 0
#define VCU_ATTR_HOST_NUMA_ID 134

typedef struct {
  unsigned int gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ, sharedMemBytes;
  CUstream hStream;
  void *attrs;
  unsigned int numAttrs;
} vcu_launch_config_t; /* CUlaunchConfig */

#define VCU_MEM_ATTACH_GLOBAL 0x1u
#define VCU_STREAM_NON_BLOCKING 0x1u
#define VCU_MEMHOSTALLOC_PORTABLE 0x1u
#define VCU_MEMHOSTALLOC_DEVICEMAP 0x2u
#define VCU_ATTR_SM_COUNT 16
#define VCU_ATTR_MAX_THREADS_PER_SM 39
#define VCU_GET_PROC_PTDS (1ull << 1)
#define VCU_WAIT_GEQ 0x0u

/* ------------------------------------------------------------------ NVML ABI subset */
typedef int nvmlReturn_t;
typedef void *nvmlDevice_t;
enum {
  NVML_SUCCESS = 0,
  NVML_ERROR_NOT_SUPPORTED = 3,
  NVML_ERROR_NOT_FOUND = 6,
  NVML_ERROR_FUNCTION_NOT_FOUND = 13,
};
typedef struct { unsigned long long total, free, used; } vnv_memory_t;
typedef struct { unsigned int version; unsigned long long total, reserved, free, used; } vnv_memory_v2_t;
typedef struct { unsigned int gpu, memory; } vnv_utilization_t;

#endif
