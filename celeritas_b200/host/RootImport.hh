//---------------------------------------------------------------------------//
// Physics-data reader (see RootImport.cpp): the reference's ROOT export of
// celeritas::ImportData -> JSON with the reference's member names.
//---------------------------------------------------------------------------//
#pragma once

#include <string>

namespace b200
{
//! Decode the single `ImportData` entry of a reference physics export
//! (/root/reference/src/celeritas/ext/RootImporter.cc, io/ImportData.hh:55-112).
//! Throws std::runtime_error on files it cannot read.
std::string import_root_to_json(std::string const& path);
}  // namespace b200
