/* Plain-C consumer of include/gsum_b200.h: proves the header is valid C (no C++ in the signatures), that the library links
 * from C, and — on a machine without a CUDA device — that the context cannot be created (there is no CPU path).
 * Built with gcc by tests/test_abi.py::test_header_is_plain_c_and_links_from_c. */
#include <stddef.h>
#include <stdio.h>
#include "gsum_b200.h"

int main(void) {
    gsum_ctx *ctx = NULL;
    /* take the address of a few entry points of every family so that the linker must resolve them */
    void *syms[] = {(void *)gsum_lml_grid, (void *)gsum_cholesky, (void *)gsum_cho_solve, (void *)gsum_fit_create, (void *)gsum_predict,
                    (void *)gsum_pivoted_cholesky, (void *)gsum_draws, (void *)gsum_grid_normalize, (void *)gsum_comm_init,
                    (void *)gsum_grid_allgather, (void *)gsum_eigh};
    int rc = gsum_ctx_create(0, NULL, &ctx);
    printf("version %d, %d symbols, gsum_ctx_create -> %d\n", gsum_version(), (int)(sizeof(syms) / sizeof(syms[0])), rc);
    if (rc == 0) {
        double X[3] = {0.0, 0.5, 1.0}, ls[1] = {0.3}, K[9];
        rc = gsum_kernel_matrix(ctx, X, 3, NULL, 0, 1, ls, 1, 1.0, 1e-6, K, GSUM_MEM_HOST);
        printf("gsum_kernel_matrix -> %d, K[0][0] = %.9f, K[0][1] = %.9f\n", rc, K[0], K[1]);
        gsum_ctx_destroy(ctx);
        return rc != 0;
    }
    return rc == -2 || rc == -3 ? 0 : 1;      /* -2: no CUDA device — the expected outcome on the CPU builder */
}
