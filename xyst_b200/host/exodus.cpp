// xyst_b200/host/exodus.cpp -- mesh ingest and diagnostics output of the host mirror.
//
// readExodus(): the tetrahedron mesh, its side sets and coordinates from an ExodusII file as the
// reference's regression suite uses them (tests/regression/inciter/*/*.exo), cf.
// src/IO/ExodusIIMeshReader.cpp:90-835. ExodusII sits on the NetCDF classic file format; the
// reference links libexodus/libnetcdf, which this image does not have, so the two classic variants
// the files come in -- CDF-1 ("CDF\x01", 32-bit offsets) and CDF-2 ("CDF\x02", 64-bit offsets) --
// are parsed here directly from the published format: big-endian header of dimension, attribute
// and variable lists followed by the (non-record) variable data.
// Side sets are (element, side) pairs: element ids count through the element blocks in file
// order; a side of a tetrahedron is the face tk::expofa gives (ExodusIIMeshReader.cpp:697-735), a
// "side" of a triangle-block element is the triangle itself.
//
// DiagWriter: the text format of src/IO/DiagWriter.cpp:26-112 (header "#  1:it  2:t  3:dt  4:..",
// column width max(20, precision+8), scientific) with the column names of
// Transporter::diagHeader (Transporter.cpp:909-1010).
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstring>
#include <limits>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "exodus.hpp"
#include "problems.hpp"

namespace xyst {

namespace {

struct NcVar {
  std::vector< std::size_t > dims;      // dimension lengths
  int type = 0;                         // 1 byte, 2 char, 3 short, 4 int, 5 float, 6 double
  std::uint64_t begin = 0;
  bool record = false;
  std::map< std::string, std::string > catt;   // character attributes
};

class NcFile {
  public:
    explicit NcFile( const std::string& path ) {
      std::ifstream f( path, std::ios::binary );
      if (!f) throw std::runtime_error( "Cannot open mesh file: " + path );
      m_buf.assign( std::istreambuf_iterator< char >( f ), std::istreambuf_iterator< char >() );
      if (m_buf.size() < 8 || m_buf[0] != 'C' || m_buf[1] != 'D' || m_buf[2] != 'F')
        throw std::runtime_error( "Not a NetCDF classic (ExodusII) file: " + path );
      m_version = static_cast< unsigned char >( m_buf[3] );
      if (m_version != 1 && m_version != 2)
        throw std::runtime_error( "Unsupported NetCDF format variant (only CDF-1 and CDF-2): " + path );
      m_pos = 4;
      m_numrecs = u32();
      // dimensions
      auto tag = u32(); auto n = u32();
      if (tag != 0 && tag != 0x0A) throw std::runtime_error( "NetCDF header: dimension list expected" );
      std::vector< std::size_t > dimlen( n );
      for (std::uint32_t i=0; i<n; ++i) { auto nm = name(); dimlen[i] = u32(); m_dim[nm] = dimlen[i]; if (dimlen[i] == 0) m_recdim = static_cast< int >( i ); }
      attributes( nullptr );              // global attributes
      tag = u32(); n = u32();
      if (tag != 0 && tag != 0x0B) throw std::runtime_error( "NetCDF header: variable list expected" );
      for (std::uint32_t i=0; i<n; ++i) {
        auto nm = name();
        NcVar v;
        auto nd = u32();
        for (std::uint32_t d=0; d<nd; ++d) {
          auto id = u32();
          if (id >= dimlen.size()) throw std::runtime_error( "NetCDF header: bad dimension id" );
          if (static_cast< int >( id ) == m_recdim) v.record = true;
          v.dims.push_back( dimlen[id] );
        }
        attributes( &v );
        v.type = static_cast< int >( u32() );
        u32();                             // vsize
        v.begin = m_version == 1 ? u32() : u64();
        m_var[nm] = std::move( v );
      }
    }
    bool has( const std::string& n ) const { return m_var.count( n ) > 0; }
    std::size_t dim( const std::string& n ) const { auto i = m_dim.find( n ); return i == m_dim.end() ? 0 : i->second; }
    const NcVar& var( const std::string& n ) const {
      auto i = m_var.find( n );
      if (i == m_var.end()) throw std::runtime_error( "ExodusII variable not found: " + n );
      return i->second;
    }
    //! all values of a fixed-size numeric variable as doubles / as 64-bit integers
    template< class T > std::vector< T > read( const std::string& n ) const {
      const auto& v = var( n );
      if (v.record) throw std::runtime_error( "record variables are not read: " + n );
      std::size_t cnt = 1; for (auto d : v.dims) cnt *= d;
      static const std::size_t sz[] = { 0, 1, 1, 2, 4, 4, 8 };
      if (v.type < 1 || v.type > 6) throw std::runtime_error( "NetCDF: bad type of " + n );
      if (v.begin + cnt*sz[v.type] > m_buf.size()) throw std::runtime_error( "NetCDF: data of " + n + " beyond end of file" );
      std::vector< T > out( cnt );
      const unsigned char* p = reinterpret_cast< const unsigned char* >( m_buf.data() ) + v.begin;
      for (std::size_t i=0; i<cnt; ++i) {
        switch (v.type) {
          case 1: case 2: out[i] = static_cast< T >( static_cast< signed char >( p[i] ) ); break;
          case 3: { std::int16_t x = static_cast< std::int16_t >( (p[2*i] << 8) | p[2*i+1] ); out[i] = static_cast< T >( x ); break; }
          case 4: { std::uint32_t x = be32( p + 4*i ); out[i] = static_cast< T >( static_cast< std::int32_t >( x ) ); break; }
          case 5: { std::uint32_t x = be32( p + 4*i ); float fl; std::memcpy( &fl, &x, 4 ); out[i] = static_cast< T >( fl ); break; }
          default: { std::uint64_t x = (static_cast< std::uint64_t >( be32( p + 8*i ) ) << 32) | be32( p + 8*i + 4 ); double d; std::memcpy( &d, &x, 8 ); out[i] = static_cast< T >( d ); }
        }
      }
      return out;
    }
  private:
    static std::uint32_t be32( const unsigned char* p ) {
      return (static_cast< std::uint32_t >( p[0] ) << 24) | (static_cast< std::uint32_t >( p[1] ) << 16) |
             (static_cast< std::uint32_t >( p[2] ) << 8) | p[3];
    }
    std::uint32_t u32() {
      if (m_pos + 4 > m_buf.size()) throw std::runtime_error( "NetCDF header truncated" );
      auto v = be32( reinterpret_cast< const unsigned char* >( m_buf.data() ) + m_pos ); m_pos += 4; return v;
    }
    std::uint64_t u64() { std::uint64_t h = u32(); return (h << 32) | u32(); }
    std::string name() {
      auto n = u32();
      if (m_pos + n > m_buf.size()) throw std::runtime_error( "NetCDF header truncated" );
      std::string s( m_buf.data() + m_pos, n );
      m_pos += (n + 3) & ~std::size_t(3);
      return s;
    }
    void attributes( NcVar* v ) {
      auto tag = u32(); auto n = u32();
      if (tag != 0 && tag != 0x0C) throw std::runtime_error( "NetCDF header: attribute list expected" );
      static const std::size_t sz[] = { 0, 1, 1, 2, 4, 4, 8 };
      for (std::uint32_t i=0; i<n; ++i) {
        auto nm = name();
        auto ty = u32(); auto cnt = u32();
        if (ty < 1 || ty > 6) throw std::runtime_error( "NetCDF header: bad attribute type" );
        std::size_t bytes = cnt*sz[ty];
        if (m_pos + bytes > m_buf.size()) throw std::runtime_error( "NetCDF header truncated" );
        if (v && ty == 2) v->catt[nm] = std::string( m_buf.data() + m_pos, cnt );
        m_pos += (bytes + 3) & ~std::size_t(3);
      }
    }
    std::vector< char > m_buf;
    std::size_t m_pos = 0;
    unsigned m_version = 0;
    std::uint32_t m_numrecs = 0;
    int m_recdim = -1;
    std::map< std::string, std::size_t > m_dim;
    std::map< std::string, NcVar > m_var;
};

} // anonymous

ExoMesh readExodus( const std::string& path )
{
  NcFile f( path );
  ExoMesh m;
  auto npoin = f.dim( "num_nodes" );
  if (!npoin) throw std::runtime_error( "ExodusII file without nodes: " + path );
  if (f.has( "coordx" )) {
    m.coord[0] = f.read< real >( "coordx" );
    m.coord[1] = f.has( "coordy" ) ? f.read< real >( "coordy" ) : std::vector< real >( npoin, 0.0 );
    m.coord[2] = f.has( "coordz" ) ? f.read< real >( "coordz" ) : std::vector< real >( npoin, 0.0 );
  } else {                                   // older files: coord[num_dim][num_nodes]
    auto c = f.read< real >( "coord" );
    auto nd = f.dim( "num_dim" );
    for (std::size_t d=0; d<3; ++d)
      m.coord[d] = d < nd ? std::vector< real >( c.begin()+static_cast< std::ptrdiff_t >( d*npoin ), c.begin()+static_cast< std::ptrdiff_t >( (d+1)*npoin ) )
                          : std::vector< real >( npoin, 0.0 );
  }
  // element blocks in file order: file-internal element id -> (block type, block-relative id)
  auto nblk = f.dim( "num_el_blk" );
  struct Blk { bool tet; std::size_t n, first; };
  std::vector< Blk > blk;
  std::vector< std::size_t > tris;
  for (std::size_t b=1; b<=nblk; ++b) {
    auto nm = "connect" + std::to_string( b );
    const auto& v = f.var( nm );
    auto et = v.catt.count( "elem_type" ) ? v.catt.at( "elem_type" ) : std::string();
    for (auto& ch : et) ch = static_cast< char >( std::toupper( static_cast< unsigned char >( ch ) ) );
    auto conn = f.read< long long >( nm );
    if (et.rfind( "TET", 0 ) == 0) {
      if (v.dims.size() != 2 || v.dims[1] != 4) throw std::runtime_error( "only 4-node tetrahedra are supported" );
      blk.push_back( { true, v.dims[0], m.tets.size()/4 } );
      for (auto c : conn) { if (c < 1 || static_cast< std::size_t >( c ) > npoin) throw std::runtime_error( "connectivity out of range" ); m.tets.push_back( static_cast< std::size_t >( c-1 ) ); }
    } else if (et.rfind( "TRI", 0 ) == 0) {
      if (v.dims.size() != 2 || v.dims[1] != 3) throw std::runtime_error( "only 3-node triangles are supported" );
      blk.push_back( { false, v.dims[0], tris.size()/3 } );
      for (auto c : conn) { if (c < 1 || static_cast< std::size_t >( c ) > npoin) throw std::runtime_error( "connectivity out of range" ); tris.push_back( static_cast< std::size_t >( c-1 ) ); }
    } else throw std::runtime_error( "unsupported element type in ExodusII file: " + et );
  }
  if (m.tets.empty()) throw std::runtime_error( "ExodusII file without tetrahedra: " + path );
  // side sets -> boundary triangles
  static const int expofa[4][3] = { {0,1,3}, {1,2,3}, {0,3,2}, {0,2,1} };   // tk::expofa, DerivedData.hpp
  auto nss = f.dim( "num_side_sets" );
  if (nss) {
    auto ids = f.read< long long >( "ss_prop1" );
    for (std::size_t s=1; s<=nss; ++s) {
      auto el = f.read< long long >( "elem_ss" + std::to_string( s ) );
      auto sd = f.read< long long >( "side_ss" + std::to_string( s ) );
      auto& out = m.sidetri[ static_cast< int >( ids[s-1] ) ];
      for (std::size_t i=0; i<el.size(); ++i) {
        auto e = static_cast< std::size_t >( el[i]-1 );
        std::size_t acc = 0; const Blk* B = nullptr;
        for (const auto& b : blk) { if (e < acc + b.n) { B = &b; break; } acc += b.n; }
        if (!B) throw std::runtime_error( "side set element id out of range" );
        auto r = B->first + (e - acc);
        if (B->tet) {
          auto k = static_cast< std::size_t >( sd[i]-1 );
          if (k > 3) throw std::runtime_error( "side set side id out of range" );
          for (int j=0; j<3; ++j) out.push_back( m.tets[r*4 + static_cast< std::size_t >( expofa[k][j] )] );
        } else
          for (std::size_t j=0; j<3; ++j) out.push_back( tris[r*3+j] );
      }
    }
  }
  return m;
}

std::vector< std::string > diagNames( const Config& cfg )
{
  std::vector< std::string > d;
  if (cfg.solver == "chocg") {               // Transporter.cpp:959-1005
    bool psol = static_cast< bool >( problems::PRESSURE_SOL( cfg ) );
    std::vector< std::string > var{ "p" };
    if (!psol) { var.push_back( "u" ); var.push_back( "v" ); var.push_back( "w" ); }
    for (const auto& v : var) d.push_back( "L2(" + v + ')' );
    for (const auto& v : var) d.push_back( "L2(d" + v + ')' );
    if (psol) { d.push_back( "L2(err:p)" ); d.push_back( "L1(err:p)" ); }
    else if (problems::SOL( cfg )) {
      for (std::size_t i=1; i<var.size(); ++i) d.push_back( "L2(err:" + var[i] + ')' );
      for (std::size_t i=1; i<var.size(); ++i) d.push_back( "L1(err:" + var[i] + ')' );
    }
    return d;
  }
  std::vector< std::string > var{ "r", "ru", "rv", "rw", "rE" };             // :919-957
  for (const auto& v : var) d.push_back( "L2(" + v + ')' );
  for (const auto& v : var) d.push_back( "L2(d" + v + ')' );
  d.push_back( "mE" );
  if (problems::SOL( cfg )) {
    for (auto v : { "r", "u", "v", "w", "e" }) d.push_back( std::string( "L2(err:" ) + v + ')' );
    for (auto v : { "r", "u", "v", "w", "e" }) d.push_back( std::string( "L1(err:" ) + v + ')' );
  }
  return d;
}

DiagWriter::DiagWriter( const std::string& filename, int precision, const std::vector< std::string >& names )
  : m_out( filename ), m_width( std::max( 20, precision+8 ) )
{
  if (!m_out) throw std::runtime_error( "Failed to open file: " + filename );
  m_out << std::scientific;
  if (precision > 0 && precision < std::numeric_limits< real >::digits10+2) m_out << std::setprecision( precision );
  m_out << "#" << std::setw(9) << "1:it" << std::setw(m_width) << "2:t" << std::setw(m_width) << "3:dt";
  std::size_t column = 4;
  for (const auto& n : names) { std::stringstream s; s << column++ << ':' << n; m_out << std::setw(m_width) << s.str(); }
  m_out << std::endl;
}

void DiagWriter::write( const std::vector< real >& row )
{
  if (row.size() < 3) return;
  m_out << std::setw(10) << static_cast< std::uint64_t >( row[0] ) << std::setw(m_width) << row[1] << std::setw(m_width) << row[2];
  for (std::size_t i=3; i<row.size(); ++i) m_out << std::setw(m_width) << row[i];
  m_out << std::endl;
}

} // xyst::
