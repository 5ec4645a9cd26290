// do_tcm.cpp — host-side dense TCM construction (runs once per character metadata).
//
// Same results as the reference's setUp2dCostMtx / distance
// (lib/core/ffi/external-direct-optimization/c_code_alloc_setup.c:25-43, 110-205; tables indexed
// (e1 << a) + e2 as in costMatrix.c:50-55): for every pair of non-empty ambiguity sets, the
// Sankoff cost min_s d(s,e1) + d(s,e2) and the set of symbols s attaining it, with
// d(s,e) = min_{pos in e} sigma[pos*a + s]  (the reference's transposed indexing, kept on purpose).
#include "do_host.h"

#include <climits>
#include <cstdlib>
#include <vector>

namespace pcgdo {

void build_dense_tables(const unsigned int *sigma, size_t a, unsigned int *cost, unsigned int *median)
{
    const size_t dim = (size_t)1 << a;
    std::vector<int> d(dim * a, INT_MAX);
    for (size_t e = 1; e < dim; ++e)
        for (size_t s = 0; s < a; ++s) {
            int best = INT_MAX;
            for (size_t pos = 0; pos < a; ++pos)
                if (e & ((size_t)1 << pos)) {
                    const int c = (int)sigma[pos * a + s];
                    if (c < best) best = c;
                }
            d[e * a + s] = best;
        }
    for (size_t e1 = 1; e1 < dim; ++e1)
        for (size_t e2 = 1; e2 < dim; ++e2) {
            int best = INT_MAX;
            unsigned int med = 0;
            for (size_t s = 0; s < a; ++s) {
                const int c = d[e1 * a + s] + d[e2 * a + s];
                if (c < best)       { best = c; med = 1u << s; }
                else if (c == best) { med |= 1u << s; }
            }
            cost[(e1 << a) + e2] = (unsigned int)best;
            median[(e1 << a) + e2] = med;
        }
}

}  // namespace pcgdo

extern "C" void setUp2dCostMtx(cost_matrices_2d_t *m, unsigned int *tcm, size_t alphSize, unsigned int gap_open)
{
    const size_t dim = (size_t)1 << alphSize;
    m->alphSize = alphSize;
    m->costMatrixDimension = dim;
    m->gap_char = 1u << (alphSize - 1);
    m->cost_model_type = gap_open == 0 ? 0 : 3;      // c_code_alloc_setup.c:119
    m->include_ambiguities = 1;
    m->gap_open_cost = gap_open;
    m->is_metric = 1;
    m->num_elements = dim - 1;
    // same (over-)allocation as costMatrix.c:178-186 so callers that index these stay in bounds
    const size_t bytes = 2 * dim * dim * sizeof(int);
    m->cost = (unsigned int *)calloc(bytes, 1);
    m->worst = (unsigned int *)calloc(bytes, 1);
    m->prepend_cost = (unsigned int *)calloc(bytes, 1);
    m->tail_cost = (unsigned int *)calloc(bytes, 1);
    m->median = (elem_t *)calloc(bytes, 1);
    if (!m->cost || !m->worst || !m->prepend_cost || !m->tail_cost || !m->median) abort();
    pcgdo::build_dense_tables(tcm, alphSize, m->cost, m->median);
    const size_t gap = m->gap_char;
    for (size_t i = 1; i < dim; ++i) {
        m->prepend_cost[i] = m->cost[(gap << alphSize) + i];
        m->tail_cost[i] = m->cost[(i << alphSize) + gap];
    }
}

// c_code_alloc_setup.c:208-270 with cm_alloc_3d (costMatrix.c:408-449): (dim + 1)^3 zero-initialised entries, entry
// (((e1 << a) + e2) << a) + e3 = min over symbols s of d(s,e1) + d(s,e2) + d(s,e3) and the set of symbols attaining it;
// d is the same transposed-index `distance` as in the 2D tables.  The per-element distances are tabulated first
// (the reference recomputes them inside the triple loop).
extern "C" void setUp3dCostMtx(cost_matrices_3d_t *m, unsigned int *tcm, size_t alphSize, unsigned int gap_open)
{
    const size_t a = alphSize, dim = (size_t)1 << a;
    m->alphSize = a;
    m->costMatrixDimension = dim;
    m->gap_char = 1u << (a - 1);
    m->cost_model_type = gap_open == 0 ? 0 : 3;
    m->include_ambiguities = 1;
    m->gap_open_cost = gap_open;
    m->num_elements = dim - 1;
    const size_t size = (dim + 1) * (dim + 1) * (dim + 1);
    m->cost = (unsigned int *)calloc(size * sizeof(int), 1);
    m->median = (elem_t *)calloc(size * sizeof(elem_t), 1);
    if (!m->cost || !m->median) abort();
    std::vector<int> d(dim * a, INT_MAX);
    for (size_t e = 1; e < dim; ++e)
        for (size_t s = 0; s < a; ++s) {
            int best = INT_MAX;
            for (size_t pos = 0; pos < a; ++pos)
                if (e & ((size_t)1 << pos)) {
                    const int c = (int)tcm[pos * a + s];
                    if (c < best) best = c;
                }
            d[e * a + s] = best;
        }
    std::vector<int> d12(a);
    for (size_t e1 = 1; e1 < dim; ++e1)
        for (size_t e2 = 1; e2 < dim; ++e2) {
            for (size_t s = 0; s < a; ++s) d12[s] = d[e1 * a + s] + d[e2 * a + s];
            unsigned int *crow = m->cost + (((e1 << a) + e2) << a);
            elem_t *mrow = m->median + (((e1 << a) + e2) << a);
            for (size_t e3 = 1; e3 < dim; ++e3) {
                int best = INT_MAX;
                unsigned int med = 0;
                const int *d3 = &d[e3 * a];
                for (size_t s = 0; s < a; ++s) {
                    const int c = d12[s] + d3[s];
                    if (c < best)       { best = c; med = 1u << s; }
                    else if (c == best) { med |= 1u << s; }
                }
                crow[e3] = (unsigned int)best;
                mrow[e3] = med;
            }
        }
}
