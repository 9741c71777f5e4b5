// walk_k6.cu — the walk kernels for K = 6 states (one translation unit per state count, see kernel_api.hpp).
#include "walk_inst.cuh"
MCP_DEFINE_KERNEL_TABLE(6)
