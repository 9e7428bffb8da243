// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/utils/log_utils.h: the reference's logging helpers
// (Debugt / Infot / Warnt / Errort / ... one family per logger) compile to nothing.
#pragma once
#include <memory>
#include <string>

#include <spdlog/spdlog.h>

namespace dynamic_vins {

class MyLogger {
public:
    inline static std::string kLogOutputDir;
    inline static std::shared_ptr<spdlog::logger> vio_logger, tk_logger, sg_logger;
};

#define DVSHIM_LOG_FAMILY(sfx)                                                   \
    template <class... A> inline void Debug##sfx(const A&...) {}                \
    template <class... A> inline void Info##sfx(const A&...) {}                 \
    template <class... A> inline void Warn##sfx(const A&...) {}                 \
    template <class... A> inline void Error##sfx(const A&...) {}                \
    template <class... A> inline void Critical##sfx(const A&...) {}
DVSHIM_LOG_FAMILY(v)
DVSHIM_LOG_FAMILY(s)
DVSHIM_LOG_FAMILY(t)
#undef DVSHIM_LOG_FAMILY

}  // namespace dynamic_vins
