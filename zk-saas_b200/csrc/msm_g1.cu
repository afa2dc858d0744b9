// msm_g1.cu -- instantiates the MSM pipeline for BN254 G1 (base field Fq).
#include "msm_impl.cuh"
namespace zkg {
ZKG_MSM_DEFINE(g1, Fq)
}
