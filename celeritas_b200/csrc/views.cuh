//---------------------------------------------------------------------------//
// Device views: plain structs of raw device pointers and scalars.
//
// ParamsView is the immutable problem (what the reference keeps in
// CoreParamsData, /root/reference/src/celeritas/global/CoreTrackData.hh:65-106)
// re-laid out as flat columns in HBM. StateView is the mutable per-track-slot
// state (CoreStateData, CoreTrackData.hh:114-160) as a structure of arrays:
// every field is its own contiguous array indexed by track slot, so that a
// warp touching one field of 32 consecutive slots issues one coalesced access.
// Multi-level fields (ORANGE per-level position/direction/volume) are stored
// level-major: index = level * num_slots + slot.
//---------------------------------------------------------------------------//
#pragma once

#include "types.cuh"

namespace b200
{
//---------------------------------------------------------------------------//
// GEOMETRY (ORANGE)
//---------------------------------------------------------------------------//
enum SurfaceType : u8
{
    SURF_PX = 0, SURF_PY, SURF_PZ, SURF_CXC, SURF_CYC, SURF_CZC, SURF_SC,
    SURF_CX, SURF_CY, SURF_CZ, SURF_P, SURF_S, SURF_KX, SURF_KY, SURF_KZ,
    SURF_SQ, SURF_GQ, SURF_INV, SURF_SIZE
};

enum UniverseType : u8
{
    UNIV_SIMPLE = 0,
    UNIV_RECT_ARRAY = 1
};

enum TransformType : u8
{
    TRANSFORM_NONE = 0,
    TRANSFORM_TRANSLATION = 1,
    TRANSFORM_TRANSFORMATION = 2
};

enum VolumeFlags : u32
{
    VOL_INTERNAL_SURFACES = 0x1,
    VOL_IMPLICIT = 0x2,
    VOL_SIMPLE_SAFETY = 0x4,
    VOL_EMBEDDED_UNIVERSE = 0x8
};

//! Logic tokens (reference logic::OperatorToken, OrangeTypes.hh:244-254)
constexpr u32 LOGIC_BEGIN = 0xfff9u;
constexpr u32 LOGIC_TRUE = 0xfffbu;
constexpr u32 LOGIC_OR = 0xfffcu;
constexpr u32 LOGIC_AND = 0xfffdu;
constexpr u32 LOGIC_NOT = 0xfffeu;

//! One row per simple unit (16 x u32), see Params.cc for the column order
struct SimpleUnit
{
    u32 surf_begin;      // first entry in surface_types / real_ids
    u32 surf_end;
    u32 real_id_begin;   // first entry in real_ids for this unit's surfaces
    u32 conn_begin;      // first connectivity record (indexed by local surface)
    u32 vol_begin;       // first volume record
    u32 vol_count;
    u32 background;      // local volume id or INVALID
    u32 simple_safety;
    u32 bbox_begin;      // first bbox (indexed by local volume)
    u32 inner_begin;
    u32 inner_count;
    u32 leaf_begin;
    u32 leaf_count;
    u32 inf_begin;       // into bih_local_volume_ids
    u32 inf_count;
    u32 pad;
};

struct RectArray
{
    u32 daughter_begin;
    u32 daughter_count;
    u32 dims[3];
    u32 grid_begin[3];  // interleaved below as begin/end pairs in the image
    u32 grid_end[3];
    u32 surf_offsets[4];
    u32 pad;
};

struct GeoParams
{
    u32 max_depth;
    u32 max_faces;
    u32 max_intersections;
    u32 num_universes;
    real tol_rel;
    real tol_abs;

    RO<u8> universe_type;
    RO<u32> universe_index;
    RO<u32> universe_surface_offset;
    RO<u32> universe_volume_offset;

    SimpleUnit const* simple_units;
    RO<u32> rect_arrays;  // 16 u32 per array (see orange.cuh)

    RO<u32> local_surface_ids;
    RO<u32> local_volume_ids;
    RO<u32> real_ids;
    RO<u16> logic_ints;
    RO<real> reals;
    RO<u8> surface_types;

    RO<u32> vol_face_begin;
    RO<u32> vol_face_end;
    RO<u32> vol_logic_begin;
    RO<u32> vol_logic_end;
    RO<u32> vol_max_isect;
    RO<u32> vol_flags;
    RO<u32> vol_daughter;

    RO<u32> conn_begin;
    RO<u32> conn_end;

    RO<u32> daughter_universe;
    RO<u32> daughter_transform;
    RO<u8> transform_type;
    RO<u32> transform_offset;

    float const* bih_bboxes;  // 6 per bbox: lower xyz, upper xyz
    RO<u32> bih_local_volume_ids;
    RO<u32> bih_inner_parent;
    RO<u32> bih_inner_axis;
    float const* bih_inner_left_pos;
    RO<u32> bih_inner_left_child;
    float const* bih_inner_right_pos;
    RO<u32> bih_inner_right_child;
    RO<u32> bih_leaf_parent;
    RO<u32> bih_leaf_vol_begin;
    RO<u32> bih_leaf_vol_end;

    RO<u32> volume_material;  // global volume id -> material id
};

//---------------------------------------------------------------------------//
// MATERIALS / PARTICLES / CUTOFFS
//---------------------------------------------------------------------------//
struct MatParams
{
    u32 num_materials;
    u32 num_elements;
    u32 max_element_components;
    RO<u32> element_z;
    RO<real> element_reals;   // 6 per element: mass, cbrt_z, cbrt_zzp, log_z, coulomb, mass_rad_coeff
    RO<u32> elcomp_element;
    RO<real> elcomp_fraction;
    RO<u32> material_elcomp_begin;
    RO<u32> material_elcomp_end;
    RO<real> material_reals;  // 8 per material: number_density, temperature, zeff, density, electron_density, rad_length, mean_exc_energy, log_mean_exc_energy
};

enum ElementReal { EL_MASS = 0, EL_CBRT_Z, EL_CBRT_ZZP, EL_LOG_Z, EL_COULOMB, EL_MASS_RAD_COEFF, EL_NUM_REALS };
enum MaterialReal { MAT_NUMBER_DENSITY = 0, MAT_TEMPERATURE, MAT_ZEFF, MAT_DENSITY, MAT_ELECTRON_DENSITY, MAT_RAD_LENGTH, MAT_MEAN_EXC, MAT_LOG_MEAN_EXC, MAT_NUM_REALS };

struct ParticleParams
{
    u32 num_particles;
    RO<real> mass;
    RO<real> charge;
    RO<real> decay_constant;
    RO<u8> matter;  // 1 = antiparticle
};

struct CutoffParams
{
    u32 num_particles;
    u32 num_materials;
    u32 apply_post_interaction;
    u32 id_gamma, id_electron, id_positron;
    RO<real> energy;  // [index][material]
    RO<real> range;
    RO<u32> id_to_index;
};

//---------------------------------------------------------------------------//
// PHYSICS TABLES
//---------------------------------------------------------------------------//
struct PhysParams
{
    u32 num_particles;
    u32 max_processes;  // P
    u32 num_materials;
    u32 num_models;

    real min_range;
    real max_step_over_range;
    real min_eprime_over_e;
    real lowest_electron_energy;
    real linear_loss_limit;
    real fixed_step_limiter;
    real lambda_limit;
    real range_factor;
    real safety_factor;

    u32 model_to_action;
    u32 step_limit_algorithm;
    u32 fixed_step_action;

    // Grids
    RO<u32> grid_size;
    RO<real> grid_log_front;
    RO<real> grid_log_back;
    RO<real> grid_log_delta;
    RO<u32> grid_prime;
    RO<u32> grid_value_offset;
    RO<real> reals;
    // Node energies exp(log_front + i * log_delta) of every grid, tabulated by the host
    // loader with the host libm: the same doubles the reference computes at run time
    // (UniformGrid::operator[] + std::exp in XsCalculator), without two exp() calls per
    // lookup on the device
    RO<u32> grid_energy_offset;  // [grid] into grid_energy
    RO<real> grid_energy;

    // Per (particle, ppid)
    RO<u32> pp_num;        // [particle]
    RO<u32> pp_eloss_ppid; // [particle]
    RO<u32> pp_has_at_rest;
    RO<u32> pp_process;    // [particle][P]
    RO<u32> pp_grid;       // [3][particle][P][material]
    RO<u8> pp_integral;    // [particle][P]
    RO<real> pp_energy_max_xs;  // [particle][P][material]
    RO<u32> pp_model_begin;     // [particle][P] -> pm_pmid
    RO<u32> pp_model_count;
    RO<u32> pm_energy_begin;    // [particle][P] -> pm_energy
    RO<real> pm_energy;
    RO<u32> pm_pmid;
    RO<u32> pmid_model;         // [pmid] -> model id
    RO<u32> elsel_begin;        // [pmid][material] -> elsel_grid or INVALID
    RO<u32> elsel_count;
    RO<u32> elsel_grid;

    // Hardwired (on-the-fly xs) models
    u32 hw_photoelectric;       // process id
    u32 hw_livermore_pe;        // model id
    u32 hw_positron_annihilation;
    u32 hw_eplusgg;
    real hw_photoelectric_table_thresh;
};

enum ValueGridType { VGT_MACRO_XS = 0, VGT_ENERGY_LOSS = 1, VGT_RANGE = 2 };

//---------------------------------------------------------------------------//
// MODEL DATA
//---------------------------------------------------------------------------//
struct KleinNishinaParams
{
    u32 action;
    u32 electron;
    u32 gamma;
    real inv_electron_mass;
};

struct MollerBhabhaParams
{
    u32 action;
    u32 electron;
    u32 positron;
    real electron_mass;
};

struct EPlusGGParams
{
    u32 action;
    u32 positron;
    u32 gamma;
    real electron_mass;
};

struct BetheHeitlerParams
{
    u32 action;
    u32 electron;
    u32 positron;
    u32 gamma;
    u32 enable_lpm;
    real electron_mass;
};

//! Seltzer-Berger scaled differential cross sections: one 2D grid per element
//! (log E x exiting fraction), see em/data/SeltzerBergerData.hh
struct SeltzerBergerParams
{
    u32 action;
    u32 electron;
    u32 positron;
    u32 gamma;
    real electron_mass;
    // per element: 8 u32 {x_begin, x_size, y_begin, y_size, values_begin, argmax_begin, 0, 0}
    RO<u32> elements;
    RO<u32> sizes;   // argmax values (y index of the row maximum)
    RO<real> reals;
};

struct RelativisticBremParams
{
    u32 action;
    u32 electron;
    u32 positron;
    u32 gamma;
    u32 enable_lpm;
    real electron_mass;
    // per element: fz, factor1, factor2, gamma_factor, epsilon_factor
    RO<real> elem_data;
};

//! Livermore photoelectric data (em/data/LivermorePEData.hh)
struct LivermorePEParams
{
    u32 action;
    u32 electron;
    u32 gamma;
    real inv_electron_mass;
    // per element: 8 u32 {xs_lo grid begin, xs_lo size, xs_lo value begin, xs_hi grid begin,
    //                     xs_hi size, xs_hi value begin, shell begin, shell count}
    RO<u32> elements;
    RO<real> element_thresh;  // 2 per element: thresh_lo, thresh_hi
    // per shell: 4 u32 {grid begin, size, value begin, 0}
    RO<u32> shells;
    // per shell: 13 reals {binding_energy, param_lo[6], param_hi[6]}
    RO<real> shell_reals;
    RO<real> reals;
};

//! Urban multiple scattering (em/data/UrbanMscData.hh)
struct UrbanMscParams
{
    u32 enabled;
    u32 electron;
    u32 positron;
    real electron_mass;
    real tau_small, tau_big, tau_limit, safety_tol, geom_limit;
    real low_energy_limit, high_energy_limit;
    // per material: 8 reals {stepmin_coeff[2], theta_coeff[2], tail_coeff[3], tail_corr}
    RO<real> material_data;
    // per (material, particle in {e-, e+}): 2 reals {scaled_zeff, d_over_r}
    RO<real> par_mat_data;
    // per (material, particle): xs grid {size, prime, value offset} + log grid
    RO<u32> xs_grid_u32;   // 3 per entry: size, prime_index, value_offset
    RO<real> xs_grid_f64;  // 3 per entry: log_front, log_back, log_delta
    RO<real> reals;
    RO<u32> xs_grid_energy_offset;  // [entry] into grid_energy (host-tabulated node energies)
    RO<real> grid_energy;
};

//! Energy-loss fluctuation (em/data/FluctuationData.hh)
struct FluctuationParams
{
    u32 enabled;
    u32 electron;
    real electron_mass;
    // per material: 6 reals {binding_energy[2], log_binding_energy[2], oscillator_strength[2]}
    RO<real> urban;
};

//! Uniform magnetic field + driver options (field/FieldDriverOptions.hh:26-93)
struct FieldParams
{
    u32 enabled;
    u32 max_nsteps;
    u32 max_substeps;
    real field[3];            // native units (gauss)
    real coeffi_per_charge;   // e / (MeV/c) in native units: dp/ds = coeffi * q * (p x B) / |p|
    real minimum_step;
    real delta_chord;
    real delta_intersection;
    real epsilon_step;
    real epsilon_rel_max;
    real errcon;
    real pgrow;
    real pshrink;
    real safety;
    real max_stepping_increase;
    real max_stepping_decrease;
    // RZ field map (reference: field/RZMapField.hh, RZMapFieldData.hh): rz_values != null
    // selects it; uniform grids {front, back, delta, size} in z and r, element (iz, ir) =
    // {value_z, value_r} at 2 * (iz * rz_size_r + ir)
    real rz_z[3];
    real rz_r[3];
    u32 rz_size_z;
    u32 rz_size_r;
    real const* rz_values;
};

struct PhysConstants
{
    real migdal_constant;
    real lpm_constant;     // MeV / len
    real r_electron;
    real alpha_fine_structure;
};

//! Combined bremsstrahlung (em/data/CombinedBremData.hh): Seltzer-Berger tables and
//! relativistic-model data (held in ModelParams::sb / rb) behind one action
struct CombinedBremParams
{
    u32 action;
    real sb_upper_limit;  // detail::seltzer_berger_upper_limit() = 1 GeV
};

//! Rayleigh scattering (em/data/RayleighData.hh): form-factor fit parameters per element
struct RayleighParams
{
    u32 action;
    u32 gamma;
    real hc_factor;   // units::centimeter / (c_light * h_planck)
    real mev;         // native value of 1 MeV
    RO<real> params;  // [element][a0 a1 a2 b0 b1 b2 n0 n1 n2]
};

//! Single Coulomb scattering (em/data/CoulombScatteringData.hh, WentzelOKVIData.hh) and the
//! isotope columns its executor draws the target from (mat/MaterialData.hh:30-60)
struct CoulombParams
{
    u32 action;
    u32 electron;
    u32 positron;
    u32 is_combined;
    u32 form_factor_type;  // NuclearFormFactorType: none, flat, exponential, gaussian
    real costheta_limit;
    real screening_factor;
    real a_sq_factor;
    real screen_r_sq_elec;  // (hbar / (2 C_TF a_0))^2 [(MeV/c)^2]
    real twopi_mrsq;        // 2 pi (m_e r_e)^2
    real alpha_fine_structure;
    real fm_par_hbar;       // 1 fm / hbar [1 / (MeV/c)]
    RO<real> nuclear_form_prefactor;  // per isotope
    RO<real> mott;                    // [element][electron, positron][theta 5][beta 6]
    RO<real> inv_mass_cbrt_sq;        // per material (combined with Wentzel VI only)
    RO<u32> element_isocomp_range;    // 2 per element
    RO<u32> isocomp_isotope;
    RO<real> isocomp_fraction;
    RO<u32> isotope_za;               // 2 per isotope: Z, A
    RO<real> isotope_nuclear_mass;
};

//! Muon / hadron ionisation (em/data/MuHadIonizationData.hh): four models share one
//! interactor and differ by the delta-ray energy distribution
struct MuHadIonizationParams
{
    // action ids by energy sampler: Bragg (positive) / ICRU73QO (negative), Bethe-Bloch,
    // muon Bethe-Bloch; INVALID when absent
    u32 bragg_action;
    u32 icru73qo_action;
    u32 bethe_bloch_action;
    u32 mu_bethe_bloch_action;
    u32 electron;
    real electron_mass;
    real proton_mass;           // [MeV]
    real alpha_over_twopi;
};

//! Muon bremsstrahlung (em/data/MuBremsstrahlungData.hh)
struct MuBremsstrahlungParams
{
    u32 action;
    u32 gamma;
    u32 mu_minus;
    u32 mu_plus;
    real electron_mass;
    real sqrt_euler;
    real dcs_factor;  // 16 alpha N_A (m_e r_e)^2
};

struct ModelParams
{
    KleinNishinaParams kn;
    MollerBhabhaParams mb;
    EPlusGGParams epgg;
    BetheHeitlerParams bh;
    SeltzerBergerParams sb;
    RelativisticBremParams rb;
    LivermorePEParams pe;
    CombinedBremParams cb;
    RayleighParams rayleigh;
    CoulombParams coulomb;
    MuHadIonizationParams muioni;
    MuBremsstrahlungParams mubrems;
    // any of cb / rayleigh / coulomb / muioni / mubrems present: their interactors are compiled
    // only into the per-action interaction kernels (run_interaction<true>), and such problems
    // do not use the fused step or the device-resident loop
    u32 has_extra_models;
    UrbanMscParams msc;
    FluctuationParams fluct;
    FieldParams field;
    PhysConstants constants;
};

//---------------------------------------------------------------------------//
// RNG / SIM / INIT / CORE SCALARS
//---------------------------------------------------------------------------//
struct RngParams
{
    u32 seed;
    RO<u32> jump;              // [32][5]
    RO<u32> jump_subsequence;  // [32][5]
};

struct SimParams
{
    u32 has_looping;
    RO<u32> looping_steps;  // 2 per particle: max_subthreshold_steps, max_steps
    RO<real> looping_energy;
};

struct CoreScalars
{
    u32 boundary_action;
    u32 propagation_limit_action;
    u32 tracking_cut_action;
    u32 along_step_user_action;
    u32 along_step_neutral_action;
    u32 track_order;
};

struct ParamsView
{
    // element counts of phys.reals / phys.grid_energy (table staging, kernels.cu)
    u32 phys_reals_count;
    u32 phys_energy_count;
    CoreScalars scalars;
    GeoParams geo;
    MatParams mat;
    ParticleParams particle;
    CutoffParams cutoff;
    PhysParams phys;
    ModelParams model;
    RngParams rng;
    SimParams sim;
};

//---------------------------------------------------------------------------//
// STATE
//---------------------------------------------------------------------------//
constexpr int MAX_SECONDARIES = 2;

struct StateView
{
    u32 num_slots;
    u32 max_depth;
    u32 max_processes;

    // sim
    u8* status;
    u32* track_id;
    u32* parent_id;
    u32* event_id;
    u32* num_steps;
    u32* num_looping_steps;
    real* time;
    real* step_length;
    u32* post_step_action;
    u32* along_step_action;

    // particle
    u32* particle_id;
    real* energy;

    // material
    u32* material_id;

    // geometry
    u32* geo_level;
    u32* geo_surface_level;  // INVALID = not on a boundary
    u32* geo_surf;
    u8* geo_sense;
    u8* geo_boundary;        // 1 = exiting, 0 = reentrant
    u32* geo_next_level;
    real* geo_next_step;
    u32* geo_next_surf;
    u8* geo_next_sense;
    real* geo_pos;   // [3][level][slot]
    real* geo_dir;   // [3][level][slot]
    u32* geo_vol;    // [level][slot]
    u32* geo_univ;   // [level][slot]

    // physics
    real* interaction_mfp;
    real* macro_xs;
    real* energy_deposition;
    real* dedx_range;
    real* msc_range;         // [3][slot]: range_init, range_factor, limit_min
    u8* msc_is_displaced;
    real* msc_true_path;
    real* msc_geom_path;
    real* msc_alpha;
    real* per_process_xs;    // [P][slot]
    u32* element;            // sampled element component
    // secondaries produced this step, fixed storage per slot
    u32* sec_particle;       // [MAX_SECONDARIES][slot]; INVALID = none
    real* sec_energy;        // [MAX_SECONDARIES][slot]
    real* sec_dir;           // [MAX_SECONDARIES][3][slot]

    // rng: xorstate[5], weyl
    u32* rng;                // [6][slot]

    // track initialization
    u32* vacancies;          // [slot]
    u32* secondary_counts;   // [slot + 1]
    u32* parents;            // [slot]
    u32* indices;            // [slot]
    u32* track_counters;     // [max_events]
    u32 init_capacity;
    // initializers, SoA
    u32* ti_track_id;
    u32* ti_parent_id;
    u32* ti_event_id;
    u32* ti_particle_id;
    real* ti_time;
    real* ti_energy;
    real* ti_pos;            // [3][capacity]
    real* ti_dir;            // [3][capacity]
    // volume hierarchy of the parent at the secondary's birth point, so that new
    // tracks from secondaries skip the point-location search (ti_level INVALID = locate)
    u32* ti_level;           // [capacity]
    u32* ti_vol;             // [max_depth][capacity]
    u32* ti_univ;            // [max_depth][capacity]

    // Dense lists of active slots, rebuilt at the end of every step: charged tracks
    // are track_slots[0 .. num_charged), neutral tracks track_slots[N-1 .. N-num_neutral]
    u32* track_slots;        // [slot]

    // device-resident counters (mirrors CoreStateCounters) + scratch
    u32* counters;           // see Counter enum
    u32* block_scratch;      // [6 * num_blocks] per-block totals for the end-of-step scans
    u8* slot_class;          // [slot] end-of-step classification, pass 1 -> pass 3
    // TrackOrder::init_charge only: ti_neutral_prefix[i] = neutral initializers in [0, i),
    // kept up to date as the queue grows; gives every starting track its rank among the
    // neutral / charged tracks started in the same step without a partition pass
    u32* ti_neutral_prefix;  // [init_capacity + 1]
    u32 single_event;        // event id if exactly one event is in flight, else INVALID
    // Host-mapped copy of the counters, written by the end-of-step scan as soon as they
    // are final (before the last pass runs), followed by `iteration_seq`: the host picks
    // up the step's result and prepares the next step while the last pass is still running
    u32* host_counters;      // [CTR_SIZE + 1], pinned host memory mapped into the device
    u32 iteration_seq;       // sequence number of this step iteration (host-set)
    // host-side upper bounds used only to size grids (kernels re-check device counters)
    u32 hint_active;
    u32 hint_charged;
    u32 hint_neutral;
    u32 hint_new;
    // End-of-step passes cover slots [slot_begin, num_slots) only (multiple of the block
    // size). Every slot below is inactive, settled, and listed in vacancies[i] = i: new
    // tracks take the highest vacancies first, so the busy slots cluster at the top.
    u32 slot_begin;

    // Device-resident step loop (tail.cu): runs of 32 consecutive slots
    u32* run_vac_prefix;     // [num_slots / 32 + 1] vacancies below each run
    u32* run_vac_mask;       // [num_slots / 32] bit i: slot 32 * run + i is vacant
    u64* run_scan;           // [2 * num_slots / 32] per-run exclusive prefixes inside a super-block
    u32* tail_reset_list;    // [2][num_slots] slots that became inactive (ping-pong)
    u32* tail_ctrl;          // [2] entries in each reset list

    // TrackOrder::reindex_* (SortTracksAction): permutation of ALL track slots sorted by the
    // order's key (null for the other orders), rebuilt by every sort action, and the first
    // index of every key: keys are action ids (num_sort_keys = number of actions), particle
    // ids, or 0 = active / 1 = inactive; slots without a key (inactive) sort last
    u32* sort_slots;         // [slot]
    u32* sort_offsets;       // [num_sort_keys + 2]
    u32* sort_block_counts;  // [num_sort_keys + 1][sort blocks]
    u32 num_sort_keys;

    // Interacting tracks of the current step sorted by model (null: interactions run
    // over the whole active list). Filled by the discrete-select launch.
    u32* interact_list;   // [model][slot]
    u32* interact_count;  // [16] entries per model

    // scoring (null when no detectors are registered)
    u32* pre_volume;                     // [slot] global volume id at the pre-step point
    u32 const* calo_detector_of_volume;  // [volume] detector id or INVALID
    real* calo_edep;                     // [detector] accumulated energy deposition [MeV]
    u32 num_detectors;

    // Step/hit output (reference: StepCollector + DetectorSteps, user/DetectorSteps.cu): null
    // when the problem has no sensitive volumes. Pre-step point per slot, and the compact hit
    // records of the last step in slot order (csrc/kernels_sort.cu: k_hits_gather)
    u32 const* hit_detector_of_volume;  // [volume] detector id or INVALID
    u32 hit_nonzero_edep;               // drop steps without energy deposition
    real* hit_pre;   // [8][slot]: time, pos xyz, dir xyz, energy at the pre-step point
    u32* hit_u32;    // [6][slot]: detector, track, event, parent, track step count, particle
    real* hit_f64;   // [18][slot]: step length, edep, pre {time, pos, dir, energy}, post {..}
    u32* hit_count;  // [1] records written by the last step

    // step counters accumulated on device: {track-steps, step iterations}
    u64* step_counters;

    // diagnostics (user/ActionDiagnostic.hh, user/StepDiagnostic.hh); null when disabled
    u32* diag_action_counts;  // [particle][diag_action_bins]: post-step action of every track-step
    u32 diag_action_bins;     // number of actions
    u32* diag_step_counts;    // [particle][diag_step_bins]: steps per track at its death
    u32 diag_step_bins;       // max_step_bin + 2
};

enum Counter : u32
{
    CTR_NUM_GENERATED = 0,
    CTR_NUM_INITIALIZERS,
    CTR_NUM_VACANCIES,
    CTR_NUM_ACTIVE,
    CTR_NUM_SECONDARIES,
    CTR_NUM_ALIVE,
    CTR_NUM_NEW_TRACKS,
    CTR_ERROR,
    CTR_TRACK_ID_BASE,
    CTR_NUM_CHARGED,   // entries at the front of track_slots
    CTR_NUM_NEUTRAL,   // entries at the back of track_slots
    CTR_SCAN_DONE,     // blocks of the end-of-step scan that have finished
    CTR_SCAN_TOTALS,   // 6 totals
    CTR_FIRST_BUSY_BLOCK = CTR_SCAN_TOTALS + 6,  // first 128-slot block that held a track this step
    CTR_SIZE = 24
};

}  // namespace b200
