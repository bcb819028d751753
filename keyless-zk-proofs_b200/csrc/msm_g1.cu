// G1 instantiation of the MSM pipeline (see msm.cuh).
#include "msm.cuh"

namespace kzp
{

template struct MsmBases<G1Xyzz>;
template void msm_bases_create<G1Xyzz>(MsmBases<G1Xyzz>&, const uint8_t*, uint64_t, bool, cudaStream_t, uint32_t);
template void msm_bases_destroy<G1Xyzz>(MsmBases<G1Xyzz>&);
template void msm_scratch_create<G1Xyzz>(MsmScratch<G1Xyzz>&, const MsmSort&, uint32_t);
template void msm_scratch_destroy<G1Xyzz>(MsmScratch<G1Xyzz>&);
template void msm_reduce_batch<G1Xyzz>(const MsmSort&, const MsmBases<G1Xyzz>* const*, MsmScratch<G1Xyzz>* const*, int,
                                       cudaStream_t);
template void msm_last_accumulate<G1Xyzz>(const MsmSort&, const MsmScratch<G1Xyzz>&, float*, uint64_t*);

void point_op_g1(int op, const void* p, const void* q, void* out, uint64_t count, cudaStream_t st)
{
    point_op_t<G1Xyzz>(op, p, q, out, count, st);
}

} // namespace kzp
