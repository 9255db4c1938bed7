// Minimal stand-in for <kodi/General.h> (logging only).  TEST INFRASTRUCTURE ONLY.
#pragma once
enum AddonLog { ADDON_LOG_DEBUG, ADDON_LOG_INFO, ADDON_LOG_WARNING, ADDON_LOG_ERROR, ADDON_LOG_FATAL };
namespace kodi { inline void Log(AddonLog, const char*, ...) {} }
