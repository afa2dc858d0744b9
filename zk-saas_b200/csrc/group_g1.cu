// group_g1.cu -- instantiates the small group-element operations (group_impl.cuh) for BN254 G1 (base field Fq).
#include "group_impl.cuh"
namespace zkg {
ZKG_GROUP_DEFINE(g1, Fq)
}
