/* Fixed-width integer types for both build modes: the ahead-of-time nvcc/gcc build takes them from
 * <stdint.h>; NVRTC (run-time specialised plane kernels) has no standard headers. */
#ifndef RDDL_STDINT_H
#define RDDL_STDINT_H
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#define INT64_MAX 0x7fffffffffffffffll
#define INT64_MIN (-INT64_MAX - 1)
#else
#include <stdint.h>
#endif
#endif
