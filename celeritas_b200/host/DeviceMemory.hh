//---------------------------------------------------------------------------//
// Small RAII helpers for device memory owned by params/state objects.
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace celeritas_b200
{
struct CudaError : std::runtime_error
{
    int code;
    CudaError(cudaError_t e, char const* what)
        : std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e)), code(int(e))
    {
    }
};

#define B2_CUDA_CALL(expr)                                      \
    do                                                          \
    {                                                           \
        cudaError_t b2_err_ = (expr);                           \
        if (b2_err_ != cudaSuccess)                             \
            throw ::celeritas_b200::CudaError(b2_err_, #expr);  \
    } while (0)

//! Owns a set of device allocations released together
class DeviceArena
{
  public:
    DeviceArena() = default;
    DeviceArena(DeviceArena const&) = delete;
    DeviceArena& operator=(DeviceArena const&) = delete;
    ~DeviceArena()
    {
        for (void* p : ptrs_)
            cudaFree(p);
    }

    //! Allocate `count` elements of T (zero-initialised); never returns null
    template<class T>
    T* alloc(size_t count)
    {
        void* p = nullptr;
        size_t bytes = (count ? count : 1) * sizeof(T);
        B2_CUDA_CALL(cudaMalloc(&p, bytes));
        ptrs_.push_back(p);
        B2_CUDA_CALL(cudaMemset(p, 0, bytes));
        bytes_ += bytes;
        return static_cast<T*>(p);
    }

    //! Allocate and fill with one byte value (0xff gives invalid ids)
    template<class T>
    T* alloc_fill(size_t count, int byte)
    {
        T* p = this->alloc<T>(count);
        B2_CUDA_CALL(cudaMemset(p, byte, (count ? count : 1) * sizeof(T)));
        return p;
    }

    //! Upload a host vector
    template<class T>
    T const* upload(std::vector<T> const& v)
    {
        T* p = this->alloc<T>(v.size());
        if (!v.empty())
            B2_CUDA_CALL(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }

    size_t bytes() const { return bytes_; }

  private:
    std::vector<void*> ptrs_;
    size_t bytes_{0};
};
}  // namespace celeritas_b200
