// Activation / tensor-core operand storage type of the whole library, chosen at build time.
// Default: IEEE fp16 (11-bit significand: 8x finer than bf16 at the same tensor-core rate; backward
// runs under a static loss scale, see ops.py).  -DEOSVOS_ACT_BF16 selects bfloat16.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifdef EOSVOS_ACT_BF16
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
#define EOSVOS_ACT_CODE 0
#define EOSVOS_MMA_FMT 1u
#define EOSVOS_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
__device__ __forceinline__ float2 act22float2(act2_t v) { return __bfloat1622float2(v); }
__device__ __forceinline__ act2_t floats2act2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ act_t float2act(float a) { return __float2bfloat16(a); }
#else
typedef __half act_t;
typedef __half2 act2_t;
#define EOSVOS_ACT_CODE 1
#define EOSVOS_MMA_FMT 0u
#define EOSVOS_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
__device__ __forceinline__ float2 act22float2(act2_t v) { return __half22float2(v); }
__device__ __forceinline__ act2_t floats2act2(float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ act_t float2act(float a) { return __float2half_rn(a); }
#endif
