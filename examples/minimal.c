/* Minimal C99 client of the drop-in boundary (include/isdf.h): one discrete collision cost+gradient evaluation.
 *   gcc -std=c99 -Iinclude examples/minimal.c -o minimal implicit-sdf-planner_b200/libisdf_b200.so -Wl,-rpath,$PWD/implicit-sdf-planner_b200 -lm
 * Exit code: 0 = evaluated, 3 = no usable CUDA device (the library has no CPU path and says so), 1 = any other error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "isdf.h"

int main(void) {
    isdf_config cfg;
    isdf_ctx *ctx = NULL;
    enum { N = 2, X = 32 };
    static unsigned char occ[X * X * X];
    double T[N] = {2.5, 2.5};
    double C[18 * N];            /* 6N x 3 column-major: row 6i+k = t^k coefficient of piece i */
    double cost = 0.0, gradC[18 * N], gradT[N];
    const double bmin[3] = {0.0, 0.0, 0.0};
    int i, rc;

    isdf_default_config(&cfg);
    cfg.integral_intervs = 16;
    rc = isdf_create(&cfg, 0, &ctx);
    if (rc != ISDF_OK) {
        fprintf(stderr, "isdf_create: %d (%s)\n", rc, isdf_last_error());
        return rc == ISDF_ERR_CUDA ? 3 : 1;
    }
    memset(occ, 0, sizeof occ);
    for (i = 0; i < X * X; i++) occ[(size_t)16 * X * X + i] = 1;       /* a wall at x = 16 */
    memset(C, 0, sizeof C); memset(gradC, 0, sizeof gradC); memset(gradT, 0, sizeof gradT);
    /* a straight line through the wall: x(t) = 10 + 1.2 t on piece 0, continued on piece 1; y = z = 16 */
    C[0] = 10.0; C[1] = 1.2; C[6] = 13.0; C[7] = 1.2;                  /* x block: pieces 0 and 1 */
    C[6 * N + 0] = 16.0; C[6 * N + 6] = 16.0;                          /* y block */
    C[12 * N + 0] = 16.0; C[12 * N + 6] = 16.0;                        /* z block */
    if ((rc = isdf_set_shape_named(ctx, "Torus", NULL, NULL)) != ISDF_OK ||
        (rc = isdf_set_map_u8(ctx, occ, X, X, X, bmin, 1.0)) != ISDF_OK ||
        (rc = isdf_eval_discrete(ctx, N, T, C, &cost, gradC, gradT)) != ISDF_OK) {
        fprintf(stderr, "isdf: %d (%s)\n", rc, isdf_last_error());
        isdf_destroy(ctx);
        return 1;
    }
    printf("cost %.6f  dcost/dT = (%.6f, %.6f)  dcost/dc_x1[piece 0] = %.6f\n", cost, gradT[0], gradT[1], gradC[1]);
    isdf_destroy(ctx);
    return 0;
}
