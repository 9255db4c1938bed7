// host/Definitions.h -- types of the reference's DSP API (Definitions.h:44-45, IirFilter.h:15) for the drop-in host classes.
#pragma once
#include <complex>
typedef float RealType;
typedef std::complex<RealType> ComplexType;
class cRadioReceiver; // the PVR client (RadioReceiver.h); only carried as an opaque pointer
typedef enum eFilterType // Definitions.h:19-26
{
  ftLP,
  ftHP,
  ftBP,
  ftBR,
  ftConst
} eFilterType;
