//---------------------------------------------------------------------------//
// Helpers shared by the C-ABI launchers (kernels*.cu, tail.cu).
//---------------------------------------------------------------------------//
#pragma once

#include <atomic>

#include "step_device.cuh"

namespace b200
{
//! Kernels launched by this library since it was loaded (b200_launch_count)
extern std::atomic<uint64_t> g_launches;
#define B2_COUNT(n) ::b200::g_launches.fetch_add(n, std::memory_order_relaxed)

inline unsigned grid_for(u32 n)
{
    return n == 0 ? 1u : (n + BLOCK - 1) / BLOCK;
}
inline int check_launch()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}
inline ParamsView const& PV(B200ParamsView const* p)
{
    return *reinterpret_cast<ParamsView const*>(p);
}
inline StateView const& SV(B200StateView const* s)
{
    return *reinterpret_cast<StateView const*>(s);
}
//! Threads needed to cover the active list (host upper bound, capped by slots)
inline u32 active_hint(StateView const& s)
{
    return s.hint_active < s.num_slots ? s.hint_active : s.num_slots;
}
}  // namespace b200
