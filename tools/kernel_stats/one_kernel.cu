#include "../../include/mcphylo_b200.h"
#include "schedule.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <type_traits>
#include "device_layout.cuh"
#include "device_math.cuh"
#include KW
namespace {
#ifndef I_K
#define I_K 4
#define I_C 1
#define I_NE 3
#endif
#ifndef I_ACCG
#define I_ACCG false
#define I_DYN false
#endif
template __global__ void felsenstein_walk<I_K, I_C, I_DYN, false, I_NE, I_ACCG>(const __grid_constant__ WalkParams);
}
