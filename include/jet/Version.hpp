#pragma once
#include <string>
namespace Jet {
constexpr size_t MAJOR_VERSION = 0;
constexpr size_t MINOR_VERSION = 2;
constexpr size_t PATCH_VERSION = 3;
/// Version of the Jet API this engine drops in for (/root/reference/include/jet/Version.hpp).
inline std::string Version() { return "0.2.3-dev+b200"; }
} // namespace Jet
