// Minimal stand-in for the Kodi dev-kit header <kodi/AddonBase.h>.
// TEST INFRASTRUCTURE ONLY: lets the reference DSP sources compile in place
// (variant 1).  Not part of the product.
#pragma once
#define ATTRIBUTE_HIDDEN __attribute__((visibility("hidden")))
