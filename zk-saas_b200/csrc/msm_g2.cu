// msm_g2.cu -- instantiates the MSM pipeline for BN254 G2 (base field Fq2).
#include "msm_impl.cuh"
namespace zkg {
ZKG_MSM_DEFINE(g2, Fq2)
}
