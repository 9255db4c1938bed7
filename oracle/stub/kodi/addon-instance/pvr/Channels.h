// Stand-in for <kodi/addon-instance/pvr/Channels.h>: the classes live in ../PVR.h.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../PVR.h"
