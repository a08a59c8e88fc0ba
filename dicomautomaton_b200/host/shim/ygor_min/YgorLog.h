// Minimal stand-in for Ygor's YgorLog.h: stream-style log macros to stderr (YLOG_VERBOSITY env: 0 silences info).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace ygor_min {
inline bool log_info_enabled() {
    static const bool on = [] { const char *e = std::getenv("YLOG_VERBOSITY"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace ygor_min
#define YLOGDEBUG(x) do { } while (0)
#define YLOGINFO(x) do { if (ygor_min::log_info_enabled()) { std::ostringstream _os; _os << x; std::cerr << "--(I) " << _os.str() << std::endl; } } while (0)
#define YLOGWARN(x) do { std::ostringstream _os; _os << x; std::cerr << "--(W) " << _os.str() << std::endl; } while (0)
#define YLOGERR(x) do { std::ostringstream _os; _os << x; std::cerr << "--(E) " << _os.str() << std::endl; std::abort(); } while (0)
