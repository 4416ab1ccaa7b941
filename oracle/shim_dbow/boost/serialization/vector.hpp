// TEST INFRASTRUCTURE ONLY - Boost is not in this image; DBoW2's BowVector.h / FeatureVector.h only name these in
// serialize() templates that the oracle never instantiates.
#pragma once
namespace boost { namespace serialization {
class access;
template <class B, class D> B& base_object(D& d) { return d; }
} }
