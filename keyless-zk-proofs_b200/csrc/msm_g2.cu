// G2 instantiation of the MSM pipeline (see msm.cuh).
#include "msm.cuh"

namespace kzp
{

template struct MsmBases<G2Xyzz>;
template void msm_bases_create<G2Xyzz>(MsmBases<G2Xyzz>&, const uint8_t*, uint64_t, bool, cudaStream_t, uint32_t);
template void msm_bases_destroy<G2Xyzz>(MsmBases<G2Xyzz>&);
template void msm_scratch_create<G2Xyzz>(MsmScratch<G2Xyzz>&, const MsmSort&, uint32_t);
template void msm_scratch_destroy<G2Xyzz>(MsmScratch<G2Xyzz>&);
template void msm_reduce_batch<G2Xyzz>(const MsmSort&, const MsmBases<G2Xyzz>* const*, MsmScratch<G2Xyzz>* const*, int,
                                       cudaStream_t);
template void msm_last_accumulate<G2Xyzz>(const MsmSort&, const MsmScratch<G2Xyzz>&, float*, uint64_t*);

void point_op_g2(int op, const void* p, const void* q, void* out, uint64_t count, cudaStream_t st)
{
    point_op_t<G2Xyzz>(op, p, q, out, count, st);
}

} // namespace kzp
