// libzkfl.so, Pippenger pipeline: the G2 (Fq2 coordinates) instantiation of the accumulate / reduce kernels.
#include "msm_host.cuh"

ZK_INSTANTIATE_MSM(Fq2)
