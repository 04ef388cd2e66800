// Shadow header (oracle/_ref build only): a no-op stand-in for hemelb::log::Logger.
#pragma once
namespace hemelb::log {
  enum LogLevel { Critical, Error, Warning, Info, Debug, Trace };
  enum LogType { Singleton, OnePerCore };
  struct Logger {
    template <LogLevel L, LogType T, class... A> static void Log(A&&...) {}
    template <LogLevel L> static constexpr bool ShouldDisplay() { return false; }
  };
}
