/*
 * Minimal C program against the C ABI (include/mcphylo_b200.h): what any FFI binding does.
 * No Python, no torch: the library is loaded with dlopen, a 4-taxon JC tree is evaluated.
 *
 *   gcc -O2 -o c_abi_example examples/c_abi_example.c -ldl -lm
 *   ./c_abi_example mcphylo.jl_b200/lib/libmcphylo_b200.so
 *
 * Tree ((a:0.1,b:0.2)e:0.05,(c:0.3,d:0.1)f:0.2)g;  nums a..d = 1..4, e = 5, f = 6, root g = 7.
 * Prints logL and the 6 branch gradients; tests/test_gpu_c_example.py compares them with the oracle.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/mcphylo_b200.h"

#define LOAD(name) \
    __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char **argv)
{
    const char *path = argc > 1 ? argv[1] : "mcphylo.jl_b200/lib/libmcphylo_b200.so";
    void *lib = dlopen(path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(mcp_create) LOAD(mcp_destroy) LOAD(mcp_last_error) LOAD(mcp_alignment_from_codes)
    LOAD(mcp_alignment_destroy) LOAD(mcp_eval)

    mcp_ctx *ctx = NULL;
    if (p_mcp_create(&ctx, 0)) { fprintf(stderr, "mcp_create: %s\n", p_mcp_last_error(NULL)); return 1; }

    /* 4 leaves x 8 sites, states 0..3, 4 = gap */
    const uint8_t codes[4 * 8] = {0, 1, 2, 3, 0, 0, 4, 2,
                                  0, 1, 2, 3, 1, 0, 2, 2,
                                  0, 1, 3, 3, 0, 4, 2, 1,
                                  0, 2, 2, 3, 0, 0, 2, 2};
    const int32_t leaf_nums[4] = {1, 2, 3, 4};
    mcp_alignment *aln = NULL;
    if (p_mcp_alignment_from_codes(ctx, codes, 4, 8, leaf_nums, 4, &aln)) {
        fprintf(stderr, "alignment: %s\n", p_mcp_last_error(ctx));
        return 1;
    }
    const int32_t postorder[7] = {1, 2, 5, 3, 4, 6, 7};
    const int32_t parent[7] = {5, 5, 6, 6, 7, 7, 0};
    const double blv[6] = {0.1, 0.2, 0.3, 0.1, 0.05, 0.2};
    /* Jukes-Cantor, K = 4: Q = 1/4 off-diagonal, -3/4 diagonal; eigenvalues -1 (x3), 0; mu = 4/3.
     * An orthonormal eigenbasis (columns), U^-1 = U^T. */
    const double h = 0.5, a = 0.7071067811865476, z = 0.0;
    const double U[16] = {a, -a, z, z,   z, z, a, -a,   h, h, -h, -h,   h, h, h, h}; /* column-major */
    double Uinv[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Uinv[i + 4 * j] = U[j + 4 * i];
    const double D[4] = {-1.0, -1.0, -1.0, 0.0};
    const double rates[1] = {1.0}, pi[4] = {0.25, 0.25, 0.25, 0.25};
    double ll = 0.0, grad[6];
    if (p_mcp_eval(ctx, aln, 7, postorder, parent, blv, U, D, Uinv, 4.0 / 3.0, rates, 1, pi, 1, &ll, grad)) {
        fprintf(stderr, "mcp_eval: %s\n", p_mcp_last_error(ctx));
        return 1;
    }
    printf("%.17g", ll);
    for (int i = 0; i < 6; ++i) printf(" %.17g", grad[i]);
    printf("\n");
    p_mcp_alignment_destroy(ctx, aln);
    p_mcp_destroy(ctx);
    dlclose(lib);
    return 0;
}
