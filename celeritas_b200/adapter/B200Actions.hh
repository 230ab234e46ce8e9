//---------------------------------------------------------------------------//
// Reference-side binding: the B200 step actions behind Celeritas's own plugin surface.
//
// THIS FILE IS COMPILED AGAINST THE REFERENCE'S HEADERS (it is what a Celeritas maintainer adds
// to their tree). It uses nothing of this repository except the C-ABI
// in include/celeritas_b200.h. It is not part of libceleritas_b200.so; oracle/Makefile target
// `dropin` compiles it together with the reference's CUDA build for the `-m gpu` test
// tests/test_gpu_dropin.py.
//
// Pieces, each against the reference interface it implements:
//  * B200StepAction        : CoreStepActionInterface + ConcreteAction
//                            (src/corecel/sys/ActionInterface.hh:175-186, 221-249;
//                             src/celeritas/global/ActionInterface.hh:24-28)
//  * B200AuxParams/AuxState : AuxParamsInterface / AuxStateInterface: the SoA track state of a
//                            stream lives in the reference CoreState's AuxStateVec
//                            (src/corecel/data/AuxInterface.hh:35-92, CoreState.hh:134-138)
//  * make_b200_registry    : an ActionRegistry (src/corecel/sys/ActionRegistry.hh) with the same
//                            ids, labels and StepActionOrder as the problem's own registry
//  * B200StepperAdapter    : StepperInterface (src/celeritas/global/Stepper.hh:72-110) over the
//                            reference's OWN ActionSequence and CoreState<device>
//
// Why a second registry and not the reference's Stepper<device> class: ActionRegistry is
// append-only ("can never be removed", ActionRegistry.hh:40-41) and CoreParams' constructor
// hard-wires its own kernels into the registry it is given (CoreParams.cc:139-207, 253-284), and
// Stepper<M>'s constructor builds its sequence from that registry (Stepper.cc:62-70). The body of
// B200StepperAdapter is Stepper<M>'s, line for line in meaning, with the registry swapped.
//---------------------------------------------------------------------------//
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "corecel/data/AuxInterface.hh"
#include "corecel/sys/ActionInterface.hh"
#include "corecel/sys/ActionRegistry.hh"
#include "celeritas/global/ActionInterface.hh"
#include "celeritas/global/ActionSequence.hh"
#include "celeritas/global/CoreParams.hh"
#include "celeritas/global/CoreState.hh"
#include "celeritas/global/Stepper.hh"
#include "celeritas/phys/Primary.hh"

#include "celeritas_b200.h"

namespace celeritas_b200_adapter
{
using celeritas::ActionId;
using celeritas::AuxId;
using celeritas::CoreParams;
using celeritas::MemSpace;
using CoreStateHost = celeritas::CoreState<MemSpace::host>;
using CoreStateDevice = celeritas::CoreState<MemSpace::device>;
using celeritas::size_type;
using celeritas::StepActionOrder;
using celeritas::StreamId;

//---------------------------------------------------------------------------//
//! Owner of the uploaded problem (B200Params) shared by every action and state
class B200Problem
{
  public:
    //! Upload an image serialized in memory (b200_params_create_from_memory)
    B200Problem(void const* image, std::size_t size);
    ~B200Problem();
    B200Problem(B200Problem const&) = delete;
    B200Problem& operator=(B200Problem const&) = delete;

    ::B200Params* get() const { return params_; }
    B200ParamsView const* view() const { return b200_params_view(params_); }

  private:
    ::B200Params* params_{nullptr};
};

//---------------------------------------------------------------------------//
//! Per-stream B200 track state, held in the reference CoreState's aux vector
class B200AuxState final : public celeritas::AuxStateInterface
{
  public:
    B200AuxState(B200Problem const& problem, StreamId stream, size_type num_track_slots);
    ~B200AuxState() final;

    ::B200Stepper* handle() const { return stepper_; }
    B200StateView const* view() const
    {
        return b200_state_view(b200_stepper_state(stepper_));
    }
    cudaStream_t stream() const { return b200_stepper_stream(stepper_); }

  private:
    ::B200Stepper* stepper_{nullptr};
};

class B200AuxParams final : public celeritas::AuxParamsInterface
{
  public:
    B200AuxParams(AuxId id, std::shared_ptr<B200Problem const> problem)
        : id_(id), problem_(std::move(problem))
    {
    }
    AuxId aux_id() const final { return id_; }
    std::string_view label() const final { return "b200-track-state"; }
    UPState create_state(MemSpace m, StreamId stream, size_type size) const final;

  private:
    AuxId id_;
    std::shared_ptr<B200Problem const> problem_;
};

//---------------------------------------------------------------------------//
//! What an adapter does when the reference's ActionSequence calls step()
enum class B200Role
{
    launch,          //!< one b200_step_* launcher
    begin_iteration, //!< extend-from-primaries: b200_stepper_begin_iteration
    end_iteration,   //!< extend-from-secondaries: launcher + b200_stepper_end_iteration
    merged           //!< kernel already launched by a sibling action (no-op, keeps the id)
};

class B200StepAction final : public celeritas::CoreStepActionInterface,
                             public celeritas::ConcreteAction
{
  public:
    using Launcher = int (*)(B200ParamsView const*, B200StateView const*, cudaStream_t);

    B200StepAction(ActionId id,
                   std::string label,
                   std::string description,
                   StepActionOrder order,
                   B200Role role,
                   Launcher launch,
                   std::shared_ptr<B200Problem const> problem,
                   AuxId state_id);

    StepActionOrder order() const final { return order_; }
    //! No CPU implementation exists: host execution is a configuration error
    void step(CoreParams const&, CoreStateHost&) const final;
    void step(CoreParams const&, CoreStateDevice&) const final;

    //! Stage primaries for the next iteration (ExtendFromPrimariesAction::insert)
    void insert(CoreStateDevice& state, celeritas::Span<celeritas::Primary const> primaries) const;

    B200Role role() const { return role_; }

  private:
    StepActionOrder order_;
    B200Role role_;
    Launcher launch_;
    std::shared_ptr<B200Problem const> problem_;
    AuxId state_id_;
};

//---------------------------------------------------------------------------//
struct B200Registry
{
    std::shared_ptr<celeritas::ActionRegistry> actions;
    std::shared_ptr<B200StepAction const> primaries;  //!< the begin_iteration action
};

//! Mirror of `core.action_reg()` with every step action replaced by its B200 adapter
B200Registry make_b200_registry(CoreParams const& core,
                                std::shared_ptr<B200Problem const> problem,
                                AuxId state_id);

//---------------------------------------------------------------------------//
//! Stepper<MemSpace::device> (src/celeritas/global/Stepper.cc:62-201) over the B200 registry
class B200StepperAdapter final : public celeritas::StepperInterface
{
  public:
    //! \param problem uploaded image of `input.params`
    B200StepperAdapter(Input input, std::shared_ptr<B200Problem const> problem);
    ~B200StepperAdapter() final;

    void warm_up() final;
    celeritas::StepperResult operator()() final;
    celeritas::StepperResult operator()(SpanConstPrimary primaries) final;
    void kill_active() final;
    void reseed(celeritas::UniqueEventId event_id) final;
    ActionSequenceT const& actions() const final { return *actions_; }
    celeritas::CoreStateInterface const& state() const final { return *state_; }
    SPState sp_state() final { return state_; }

    //! Kernel launches issued by the B200 library for this stepper
    std::uint64_t launch_count() const;
    //! The stream's SoA track state (aux state of the reference CoreState)
    B200AuxState& b200_state() const;

  private:
    std::shared_ptr<CoreParams const> params_;
    std::shared_ptr<B200Problem const> problem_;
    AuxId state_id_;
    B200Registry registry_;
    std::shared_ptr<ActionSequenceT> actions_;
    std::shared_ptr<celeritas::CoreState<MemSpace::device>> state_;
};
}  // namespace celeritas_b200_adapter
