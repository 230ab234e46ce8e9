// Hand-written build configuration for compiling the reference's host sources
// directly with g++ (oracle/Makefile). It states the same configuration the
// reference's CMake would generate for: CUDA (sm_100), ORANGE geometry, XORWOW RNG,
// double precision, CGS units, no OpenMP: the configuration of the reference's own CUDA build. (template:
// /root/reference/src/corecel/Config.hh.in)
#pragma once

#define CELERITAS_USE_CUDA 1
#define CELERITAS_USE_GEANT4 0
#define CELERITAS_USE_HEPMC3 0
#define CELERITAS_USE_HIP 0
#define CELERITAS_USE_MPI 0
#define CELERITAS_USE_OPENMP 0
#define CELERITAS_USE_PERFETTO 0
#define CELERITAS_USE_PNG 0
#define CELERITAS_USE_ROOT 0
#define CELERITAS_USE_VECGEOM 0

#define CELERITAS_DEBUG 0
#define CELERITAS_DEVICE_DEBUG 0

#define CELERITAS_REAL_TYPE_DOUBLE 1
#define CELERITAS_REAL_TYPE_FLOAT 2
#define CELERITAS_REAL_TYPE CELERITAS_REAL_TYPE_DOUBLE

#define CELERITAS_UNITS_CGS 1
#define CELERITAS_UNITS_SI 2
#define CELERITAS_UNITS_CLHEP 3
#define CELERITAS_UNITS CELERITAS_UNITS_CGS

#define CELERITAS_OPENMP_DISABLED 0
#define CELERITAS_OPENMP_EVENT 1
#define CELERITAS_OPENMP_TRACK 2
#define CELERITAS_OPENMP CELERITAS_OPENMP_DISABLED

#define CELERITAS_CORE_GEO_VECGEOM 0
#define CELERITAS_CORE_GEO_GEANT4 0
#define CELERITAS_CORE_GEO_ORANGE 1
#define CELERITAS_CORE_GEO CELERITAS_CORE_GEO_ORANGE

#define CELERITAS_CORE_RNG_CURAND 0
#define CELERITAS_CORE_RNG_HIPRAND 0
#define CELERITAS_CORE_RNG_XORWOW 1
#define CELERITAS_CORE_RNG CELERITAS_CORE_RNG_XORWOW

// The reference's CMake sets 256 whenever CUDA is on (/root/reference/CMakeLists.txt:94-97);
// every action kernel is __launch_bounds__(CELERITAS_MAX_BLOCK_SIZE)
#define CELERITAS_MAX_BLOCK_SIZE 256

inline constexpr char celeritas_build_type[] = "Release";
inline constexpr char celeritas_hostname[] = "oracle";
inline constexpr char celeritas_real_type[] = "double";
inline constexpr char celeritas_units[] = "CGS";
inline constexpr char celeritas_openmp[] = "disabled";
inline constexpr char celeritas_core_geo[] = "ORANGE";
inline constexpr char celeritas_core_rng[] = "xorwow";
inline constexpr char celeritas_clhep_version[] = "";
inline constexpr char celeritas_geant4_version[] = "";
inline constexpr char celeritas_vecgeom_version[] = "";

#define CELERITAS_HAVE_ROCTX 0

#define CELERITAS_GEANT4_VERSION 0x000000
#define CELERITAS_VECGEOM_VERSION 0x000000
#define CELERITAS_HEPMC3_VERSION 0x000000
