// Stand-in for <boost/serialization/serialization.hpp>, TEST INFRASTRUCTURE ONLY (see ../../opencv2/core/core.hpp).
// DBoW2's BowVector / FeatureVector only declare a serialize() member template; it is never instantiated here.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}  // namespace boost::serialization
