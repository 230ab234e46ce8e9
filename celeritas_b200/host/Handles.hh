//---------------------------------------------------------------------------//
// Definitions of the opaque C-ABI handles (include/celeritas_b200.h), shared by the
// translation units that implement the C entry points.
//---------------------------------------------------------------------------//
#pragma once

#include <memory>
#include <string>

#include "CoreParams.hh"
#include "CoreState.hh"
#include "Stepper.hh"

struct B200Params
{
    std::shared_ptr<celeritas_b200::CoreParams> params;
};

struct B200State
{
    std::unique_ptr<celeritas_b200::CoreState> owned;
    celeritas_b200::CoreState* state;
};

struct B200Stepper
{
    std::shared_ptr<celeritas_b200::Stepper> stepper;
    B200State state_handle;
    uint64_t launches_at_create;
};

namespace celeritas_b200
{
//! Message returned by b200_last_error() on the calling thread
void set_last_error(std::string const& message);
}  // namespace celeritas_b200
