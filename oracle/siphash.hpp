// oracle/siphash.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// SipHash-2-4 (Aumasson & Bernstein, "SipHash: a fast short-input PRF", 2012),
// restated from the published algorithm. The reference hashes sorted node-id
// tuples with highwayhash's SipHash and a fixed key (src/Mesh/UnsMesh.hpp:33-34,
// :75-92; vendored src/highwayhash/sip_hash.h); the iteration order of its
// unordered containers -- and so the order of triangle/edge superedges
// (src/Inciter/RieCG.cpp:680,706) -- follows from this hash.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <array>
#include <algorithm>

namespace orc {

inline std::uint64_t rotl64( std::uint64_t v, int b ) { return (v << b) | (v >> (64-b)); }

inline std::uint64_t siphash24( std::uint64_t k0, std::uint64_t k1,
                                const unsigned char* in, std::size_t len )
{
  std::uint64_t v0 = 0x736f6d6570736575ULL ^ k0, v1 = 0x646f72616e646f6dULL ^ k1,
                v2 = 0x6c7967656e657261ULL ^ k0, v3 = 0x7465646279746573ULL ^ k1;
  auto round = [&](){
    v0 += v1; v2 += v3; v1 = rotl64(v1,13); v3 = rotl64(v3,16); v1 ^= v0; v3 ^= v2;
    v0 = rotl64(v0,32);
    v2 += v1; v0 += v3; v1 = rotl64(v1,17); v3 = rotl64(v3,21); v1 ^= v2; v3 ^= v0;
    v2 = rotl64(v2,32); };
  auto absorb = [&]( std::uint64_t m ){ v3 ^= m; round(); round(); v0 ^= m; };
  std::size_t nfull = len / 8;
  for (std::size_t i=0; i<nfull; ++i) {
    std::uint64_t m; std::memcpy( &m, in + 8*i, 8 ); absorb( m ); }   // little-endian host
  unsigned char last[8] = {0,0,0,0,0,0,0,0};
  std::memcpy( last, in + 8*nfull, len - 8*nfull );
  last[7] = static_cast< unsigned char >( len & 0xff );
  std::uint64_t m; std::memcpy( &m, last, 8 ); absorb( m );
  v2 ^= 0xff;
  round(); round(); round(); round();
  return (v0 ^ v1) ^ (v2 ^ v3);
}

//! Hash / equality of an (unordered) tuple of node ids, cf. UnsMesh.hpp:75-112
template< std::size_t N > struct IdHash {
  std::size_t operator()( const std::array< std::size_t, N >& p ) const {
    std::array< std::size_t, N > s = p;
    std::sort( s.begin(), s.end() );
    return siphash24( 0x0706050403020100ULL, 0x0F0E0D0C0B0A0908ULL,
                      reinterpret_cast< const unsigned char* >( s.data() ),
                      N*sizeof(std::size_t) );
  }
};
template< std::size_t N > struct IdEq {
  bool operator()( const std::array< std::size_t, N >& l,
                   const std::array< std::size_t, N >& r ) const {
    auto s = l, p = r;
    std::sort( s.begin(), s.end() ); std::sort( p.begin(), p.end() );
    return s == p;
  }
};

} // orc::
