// rtc_compat.h — what the device headers need from <stdint.h> / <math.h> / <cuda_runtime.h> when they are compiled by
// NVRTC (the scene-specialised "baked" render kernel, bake.cpp): NVRTC knows the device built-ins but has no C library.
#pragma once
#if defined(__CUDACC_RTC__)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#ifndef NAN
#define NAN __int_as_float(0x7fffffff)
#endif
#else
#include <stdint.h>
#include <stddef.h>
#endif
