// xyst_b200/host/exodus.hpp -- ExodusII mesh ingest and the diagnostics text writer of the host
// mirror (see exodus.cpp).
#pragma once
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "riecg.hpp"

namespace xyst {

//! A tetrahedron mesh as read from a file: coordinates by (0-based) file node id, tetrahedron
//! connectivity, and the boundary triangles of every side set (node triples, oriented as the file's
//! element sides are)
struct ExoMesh {
  Coords coord;
  std::vector< std::size_t > tets;
  std::map< int, std::vector< std::size_t > > sidetri;
};

ExoMesh readExodus( const std::string& path );

//! Column names of the diagnostics file after it, t, dt (Transporter::diagHeader)
std::vector< std::string > diagNames( const Config& cfg );

//! cf. tk::DiagWriter
class DiagWriter {
  public:
    DiagWriter( const std::string& filename, int precision, const std::vector< std::string >& names );
    void write( const std::vector< real >& row );     //!< row = { it, t, dt, diagnostics... }
  private:
    std::ofstream m_out;
    int m_width;
};

} // xyst::
