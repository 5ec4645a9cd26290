/* Link proof for INTEGRATION.md §1: this program stands in for the object code GHC emits for PCG's seven
 * `foreign import ccall` declarations on the DO path and is linked against libpcgdo_b200.so ALONE — none of the
 * reference's C / C++ objects:
 *   align2d, align2dAffine           DO/Pairwise/FFI.hsc:122-143
 *   setUp2dCostMtx, setUp3dCostMtx   lib/core/tcm/src/Data/TCM/Dense/FFI.hsc:319-334
 *   matrixInit, getCostAndMedian2D,
 *   getCostAndMedian3D               lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:49-69 (and lib/core/tcm/.../Memoized/FFI.hsc:185-205)
 * (the `align3d` import at FFI.hsc:152 sits inside a block comment upstream).
 * Without arguments it exercises the host-only setup calls exactly as `performMatrixAllocation`
 * (Dense/FFI.hsc:461-476) makes them and prints table checksums; with "gpu" it also runs one call of each
 * device-backed import. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcg_do_b200.h"

static alignIO_t *alloc_aio(const unsigned *vals, size_t n, size_t cap)      /* allocInitAlign_io, FFI.hsc:474-481 */
{
    alignIO_t *io = malloc(sizeof *io);
    io->character = calloc(cap, sizeof(elem_t));
    io->length = n;
    io->capacity = cap;
    for (size_t i = 0; i < n; ++i) io->character[cap - n + i] = vals[i];
    return io;
}

int main(int argc, char **argv)
{
    void *imports[] = {(void *)align2d, (void *)align2dAffine, (void *)setUp2dCostMtx, (void *)setUp3dCostMtx,
                       (void *)matrixInit, (void *)getCostAndMedian2D, (void *)getCostAndMedian3D};
    for (size_t i = 0; i < sizeof imports / sizeof *imports; ++i)
        if (!imports[i]) return 2;
    const size_t a = 5;
    unsigned tcm[25];
    for (size_t i = 0; i < a; ++i)
        for (size_t j = 0; j < a; ++j) tcm[i * a + j] = i == j ? 0u : 1u;
    cost_matrices_2d_t *m2 = malloc(sizeof *m2);
    cost_matrices_3d_t *m3 = malloc(sizeof *m3);
    setUp2dCostMtx(m2, tcm, a, 0);
    setUp3dCostMtx(m3, tcm, a, 0);
    unsigned long long s2 = 0, s3 = 0;
    const size_t dim = (size_t)1 << a;
    for (size_t i = 0; i < dim * dim; ++i) s2 = s2 * 31 + m2->cost[i] + 7 * m2->median[i];
    for (size_t i = 0; i < (dim + 1) * (dim + 1) * (dim + 1); ++i) s3 = s3 * 31 + m3->cost[i] + 7 * m3->median[i];
    printf("tables %llu %llu\n", s2, s3);
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        const unsigned x[] = {1, 2, 4, 8, 1, 1, 2}, y[] = {1, 2, 8, 8, 1, 2};
        alignIO_t *c1 = alloc_aio(x, 7, 15), *c2 = alloc_aio(y, 6, 15), *g = alloc_aio(NULL, 0, 15), *u = alloc_aio(NULL, 0, 15);
        const int cost = align2d(c1, c2, g, u, m2, 0, 1, 0);
        printf("align2d %d len %zu\n", cost, g->length);
        unsigned big[9] = {0, 1, 2, 1, 0, 1, 2, 1, 0};
        costMatrix_p memo = matrixInit(3, big);
        uint64_t e1 = 1, e2 = 4, e3 = 2;
        dcElement_t d1 = {3, &e1}, d2 = {3, &e2}, d3 = {3, &e3}, r = {3, NULL};
        const unsigned c2d = getCostAndMedian2D(&d1, &d2, &r, memo);
        printf("overlap2 %u %llu\n", c2d, (unsigned long long)r.element[0]);
        const unsigned c3d = getCostAndMedian3D(&d1, &d2, &d3, &r, memo);
        printf("overlap3 %u %llu\n", c3d, (unsigned long long)r.element[0]);
        matrixDestroy(memo);
    }
    freeCostMtx(m2, 1);
    freeCostMtx(m3, 0);
    return 0;
}
