// do_host.h — internal host-side declarations shared by do_tcm.cpp / do_api.cu.
#pragma once
#include <cstddef>
#include "../../include/pcg_do_b200.h"

namespace pcgdo {
// cost / median: (1 << a)^2 entries each, indexed (e1 << a) + e2; row and column 0 untouched.
void build_dense_tables(const unsigned int *sigma, size_t a, unsigned int *cost, unsigned int *median);
}
