// msm_g1.cu -- instantiates the MSM pipeline for BN254 G1 (base field Fq).
#include "msm_impl.cuh"
namespace zkg {
ZKG_MSM_DEFINE(g1, Fq)
int msm_pick_c_merged_host(size_t n) { return msm_pick_c_merged(n); }
}
