// ORACLE shim (test infrastructure): logging calls of the reference compile to nothing
#pragma once
#include <memory>
#include <sstream>
#include <string>
namespace spdlog {
namespace level { enum level_enum { trace, debug, info, warn, err, critical, off }; }
class logger {
public:
    template <class... A> void log(level::level_enum, const A&...) {}
    template <class... A> void debug(const A&...) {}
    template <class... A> void info(const A&...) {}
    template <class... A> void warn(const A&...) {}
    template <class... A> void error(const A&...) {}
    template <class... A> void critical(const A&...) {}
    void flush() {}
    void set_level(level::level_enum) {}
    void flush_on(level::level_enum) {}
};
}
#include "../dvshim_fmt.hpp"
