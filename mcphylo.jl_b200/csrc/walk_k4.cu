// walk_k4.cu — the walk kernels for K = 4 states (one translation unit per state count, see kernel_api.hpp).
#include "walk_inst.cuh"
MCP_DEFINE_KERNEL_TABLE(4)
