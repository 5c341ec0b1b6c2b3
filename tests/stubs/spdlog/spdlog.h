// Minimal spdlog stub for compiling the reference's examples against include/xrt (logging only).
#pragma once
namespace spdlog {
namespace level { enum level_enum { trace, debug, info, warn, err, critical, off }; }
inline void set_level(level::level_enum) {}
template <typename... A> inline void info(A&&...) {}
template <typename... A> inline void warn(A&&...) {}
template <typename... A> inline void error(A&&...) {}
template <typename... A> inline void debug(A&&...) {}
} // namespace spdlog
