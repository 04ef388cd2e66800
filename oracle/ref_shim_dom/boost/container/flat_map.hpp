// Stand-in for boost::container::flat_map (absent from this image; oracle/_ref Domain build only):
// the reference uses it as an ordered associative cache (geometry/LookupTree.h, geometry/Domain.h).
#pragma once
#include <map>
namespace boost::container {
  template <class K, class V, class C = std::less<K>> class flat_map : public std::map<K, V, C> {
   public:
    using std::map<K, V, C>::map;
    bool contains(const K& k) const { return this->find(k) != this->end(); }
  };
}
