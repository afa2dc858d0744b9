// group_g2.cu -- instantiates the small group-element operations (group_impl.cuh) for BN254 G2 (base field Fq2).
#include "group_impl.cuh"
namespace zkg {
ZKG_GROUP_DEFINE(g2, Fq2)
}
