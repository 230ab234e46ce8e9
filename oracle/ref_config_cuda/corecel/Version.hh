// Hand-written stand-in for the reference's generated corecel/Version.hh
// (template: /root/reference/src/corecel/Version.hh.in); the reference tree
// has no git metadata so its version is unknown.
#pragma once
#define CELERITAS_VERSION 0x000000
inline constexpr char celeritas_version[] = "0.0.0-unknown";
inline constexpr int celeritas_version_major = 0;
inline constexpr int celeritas_version_minor = 0;
inline constexpr int celeritas_version_patch = 0;
