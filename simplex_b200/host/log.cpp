#include "log.h"

#include <cstdlib>
#include <cstring>
#include <ctime>

namespace simplex::log
{
    namespace
    {
        Level from_env()
        {
            const char * e = std::getenv("SIMPLEX_LOG");
            if (!e) return info;
            if (!std::strcmp(e, "off") || !std::strcmp(e, "0")) return off;
            if (!std::strcmp(e, "debug") || !std::strcmp(e, "2")) return debug;
            return info;
        }
        Level g_level = from_env();
    }  // namespace

    Level level() { return g_level; }
    void set_level(Level l) { g_level = l; }
    const char * name(Level l) { return l == debug ? "debug" : (l == info ? "info" : "off"); }

    void emit(Level l, const std::string & text)
    {
        if (l > g_level || l == off) return;
        char stamp[40];
        const std::time_t now = std::time(nullptr);
        std::tm tm_utc{};
        gmtime_r(&now, &tm_utc);
        std::strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S+0000", &tm_utc);
        std::cout << '[' << stamp << "][" << name(l) << "] " << text << '\n';
    }
}  // namespace simplex::log
