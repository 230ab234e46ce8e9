//---------------------------------------------------------------------------//
// Native ORANGE construction (see OrangeBuilder.cpp): the reference's .org.json geometry
// input -> the geo.* columns of a problem image.
//---------------------------------------------------------------------------//
#pragma once

#include <string>

#include "Image.hh"

namespace celeritas_b200
{
//! Build the geometry image of an ORANGE JSON input file (OrangeParams(OrangeInput&&),
//! /root/reference/src/orange/OrangeParams.cc:137-209). Throws std::runtime_error.
b200::Image build_orange_image(std::string const& org_json_path);
}  // namespace celeritas_b200
