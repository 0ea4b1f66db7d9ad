/* cqg_rtc.h — what <cstdint>/<stdint.h>/<stddef.h>/<cuda_runtime.h> give the device headers, for NVRTC
 * (kernels specialised per query are compiled at run time from these same headers: cqg_jit in cqg_api.cu). */
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif
