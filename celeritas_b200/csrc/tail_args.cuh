//---------------------------------------------------------------------------//
// Arguments and entry points of the device-resident step loop (tail_loop.cuh) shared by the
// translation units that instantiate it and the C-ABI launcher (tail.cu).
//---------------------------------------------------------------------------//
#pragma once

#include "views.cuh"

namespace b200
{
struct TailArgs
{
    u32 max_iterations;
    u32 exit_active;   // leave when the next iteration could hold more tracks than this
    u32* ring;         // mapped host memory: [max_iterations][TAIL_RING_WORDS]
    u32* done;         // mapped host memory: {iterations done, exit reason}
    u32 coop;          // warp-cooperative step when there are at most coop_max tracks
    u32 coop_max;
};

cudaError_t tail_blocks_per_sm_field(int* per_sm);
cudaError_t tail_blocks_per_sm_rzfield(int* per_sm);
cudaError_t tail_blocks_per_sm_nofield(int* per_sm);
cudaError_t tail_launch_field(ParamsView const&, StateView const&, TailArgs const&, u32 num_blocks,
                              cudaStream_t);
cudaError_t tail_launch_rzfield(ParamsView const&, StateView const&, TailArgs const&,
                                u32 num_blocks, cudaStream_t);
cudaError_t tail_launch_nofield(ParamsView const&, StateView const&, TailArgs const&,
                                u32 num_blocks, cudaStream_t);
}  // namespace b200
