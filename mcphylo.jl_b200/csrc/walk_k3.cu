// walk_k3.cu — the walk kernels for K = 3 states (one translation unit per state count, see kernel_api.hpp).
#include "walk_inst.cuh"
MCP_DEFINE_KERNEL_TABLE(3)
