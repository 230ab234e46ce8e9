// k_tail_loop<1> (uniform field): see tail_loop.cuh
#include "tail_loop.cuh"

namespace b200
{
cudaError_t tail_blocks_per_sm_field(int* per_sm)
{
    return tail_blocks_per_sm<1>(per_sm);
}
cudaError_t tail_launch_field(ParamsView const& p, StateView const& s, TailArgs const& a,
                           u32 num_blocks, cudaStream_t stream)
{
    return tail_launch<1>(p, s, a, num_blocks, stream);
}
}  // namespace b200
