// Stand-in for <kodi/gui/Window.h>: the base class of the reference's settings dialog (ChannelSettings.h), whose
// members the harness defines as recorders -- no window is ever shown.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
enum ADDON_ACTION { ADDON_ACTION_NONE = 0, ADDON_ACTION_PREVIOUS_MENU = 10, ADDON_ACTION_NAV_BACK = 92 };
namespace kodi
{
namespace gui
{
class CWindow
{
public:
  CWindow() = default;
  CWindow(const std::string&, const std::string&, bool, bool = true) {}
  virtual ~CWindow() = default;
  virtual bool OnInit() { return false; }
  virtual bool OnFocus(int) { return false; }
  virtual bool OnClick(int) { return false; }
  virtual bool OnAction(ADDON_ACTION) { return false; }
};
} // namespace gui
} // namespace kodi
