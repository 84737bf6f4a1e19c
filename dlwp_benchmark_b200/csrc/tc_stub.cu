// Tensor-core (tcgen05) variants hook in here; until a shape is claimed the CUDA-core
// kernels handle it.  (*handled = 0 => caller runs the generic kernel.)
#include "common.cuh"

int sb200_tc_tables_create(sb200_plan_s* p) { p->tc = nullptr; return 0; }
void sb200_tc_tables_destroy(sb200_plan_s* p) { p->tc = nullptr; }

int sb200_tc_rowdft_fwd(sb200_plan_t, int, const float*, float*, int64_t, cudaStream_t, int* handled) {
    *handled = 0;
    return 0;
}
