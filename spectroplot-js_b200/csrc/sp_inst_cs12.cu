// generated: cs12 kernels
#define SP_INST_TAG cs12
#define SP_INST_FMT sp::CS12
#include "sp_inst.cuh"
