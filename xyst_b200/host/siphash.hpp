// xyst_b200/host/siphash.hpp -- SipHash-2-4 (Aumasson & Bernstein 2012) of sorted node-id
// tuples with the reference's fixed key (src/Mesh/UnsMesh.hpp:33-34,:75-112).
//
// Why the PRODUCT needs it: the reference picks its triangle superedges by walking a
// std::unordered_set of faces in iteration order (src/Inciter/RieCG.cpp:646-651,:680-701),
// and the van Leer limiter's +1e-9 regularisation (src/Physics/Riemann.cpp:92-95) makes an
// edge's flux depend on its ORIENTATION at the 1e-9 level. Which leftover edges end up in a
// triangle (face-cyclic orientation) or alone (low->high global id) therefore changes the
// answer beyond rounding; reproducing the reference means reproducing that walk.
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace xyst {

inline std::uint64_t siphash24_ids( const std::size_t* ids, std::size_t n )
{
  const std::uint64_t k0 = 0x0706050403020100ULL, k1 = 0x0F0E0D0C0B0A0908ULL;
  std::uint64_t v0 = 0x736f6d6570736575ULL ^ k0, v1 = 0x646f72616e646f6dULL ^ k1,
                v2 = 0x6c7967656e657261ULL ^ k0, v3 = 0x7465646279746573ULL ^ k1;
  auto rotl = []( std::uint64_t v, int b ){ return (v << b) | (v >> (64-b)); };
  auto round = [&](){
    v0 += v1; v2 += v3; v1 = rotl(v1,13); v3 = rotl(v3,16); v1 ^= v0; v3 ^= v2; v0 = rotl(v0,32);
    v2 += v1; v0 += v3; v1 = rotl(v1,17); v3 = rotl(v3,21); v1 ^= v2; v3 ^= v0; v2 = rotl(v2,32); };
  for (std::size_t i=0; i<n; ++i) {       // whole 8-byte words (little-endian host)
    std::uint64_t m = ids[i];
    v3 ^= m; round(); round(); v0 ^= m;
  }
  std::uint64_t last = static_cast< std::uint64_t >( (n*8) & 0xff ) << 56;   // length byte, no tail
  v3 ^= last; round(); round(); v0 ^= last;
  v2 ^= 0xff;
  round(); round(); round(); round();
  return (v0 ^ v1) ^ (v2 ^ v3);
}

template< std::size_t N > struct IdHash {
  std::size_t operator()( const std::array< std::size_t, N >& p ) const {
    auto s = p; std::sort( s.begin(), s.end() );
    return siphash24_ids( s.data(), N );
  }
};
template< std::size_t N > struct IdEq {
  bool operator()( const std::array< std::size_t, N >& l, const std::array< std::size_t, N >& r ) const {
    auto s = l, p = r; std::sort( s.begin(), s.end() ); std::sort( p.begin(), p.end() );
    return s == p;
  }
};

} // xyst::
