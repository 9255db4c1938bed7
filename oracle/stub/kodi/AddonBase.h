// Minimal stand-in for the Kodi dev-kit header <kodi/AddonBase.h>.
// TEST INFRASTRUCTURE ONLY: lets the reference sources compile in place (variant 1).
// Not part of the product.  Written from the names the reference uses, not from the dev-kit.
#pragma once
#include <string>
#define ATTRIBUTE_HIDDEN __attribute__((visibility("hidden")))
#define ADDONCREATOR(cls) /* the add-on entry point: the harness constructs the class itself */
#ifndef STR
#define STR_(x) #x
#define STR(x) STR_(x)
#endif
#ifndef RTL_RADIOFM_VERSION
#define RTL_RADIOFM_VERSION 0.0.0
#endif
namespace kodi
{
// the add-on's data directory: files live in the stand-in tinyxml's in-memory store, so any prefix will do
inline std::string GetAddonPath(const std::string& append = "") { return "refaddon://" + append; }
namespace addon
{
class CAddonBase
{
public:
  virtual ~CAddonBase() = default;
};
} // namespace addon
} // namespace kodi
