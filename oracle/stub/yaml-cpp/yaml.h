// Minimal stand-in for <yaml-cpp/yaml.h>, used ONLY to compile the reference's
// OpenMP execution path into oracle/_ref (yaml-cpp is not installed in this image).
// The reference needs it for one thing on this path: include/onika/memory/allocator.h
// specialises YAML::convert< CudaMMVector<T> >.  Nothing here is ever executed.
// Test infrastructure: not part of the shipped product.
#pragma once
#include <cstddef>
#include <string>

namespace YAML
{
  template<class T> struct convert;

  class Node
  {
  public:
    Node() = default;
    template<class T> Node(const T&) {}
    template<class T> void push_back(const T&) {}
    bool IsSequence() const { return false; }
    bool IsMap() const { return false; }
    bool IsScalar() const { return false; }
    bool IsDefined() const { return false; }
    std::size_t size() const { return 0; }
    Node operator[](std::size_t) const { return Node{}; }
    Node operator[](const char*) const { return Node{}; }
    Node operator[](const std::string&) const { return Node{}; }
    template<class T> T as() const { return T{}; }
    Node Clone() const { return *this; }
    explicit operator bool() const { return false; }
  };
}
