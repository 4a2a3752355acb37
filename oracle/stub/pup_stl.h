// Test infrastructure only: minimal stand-in for Charm++'s pup_stl.h so that
// the reference's Charm++-free translation units compile under oracle/_ref.
// (recipe from SURVEY.md section 8c)
#pragma once
#include <vector>
#include <string>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <array>
#include <tuple>
#include <cstddef>
#include <functional>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <memory>
#include <list>
#include <deque>
#include <iostream>
#include <sstream>
#include <limits>
namespace PUP {
  class er { public:
    bool isUnpacking() const { return false; }
    bool isSizing() const { return false; }
    bool isPacking() const { return false; } };
  template<class T> inline void operator|( er&, T& ) {}
  template<class T> inline void pup( er&, T& ) {}
  template<class T> inline void PUParray( er&, T*, std::size_t ) {}
}
