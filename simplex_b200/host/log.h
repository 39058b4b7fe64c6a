// Minimal levelled logger of the host side (product code). The reference logs every iteration
// at debug level to stdout through a singleton (reference src/Logger.h:70-112); here the loop
// runs on the GPU, so per-pivot lines are replayed from the device trace after the loop when
// the debug level is on. Level: environment variable SIMPLEX_LOG = off | info | debug
// (default info), or simplex::log::set_level().
#pragma once

#include <iostream>
#include <sstream>

namespace simplex::log
{
    enum Level { off = 0, info = 1, debug = 2 };
    Level level();
    void set_level(Level l);
    const char * name(Level l);
    void emit(Level l, const std::string & text);

    class Line
    {
    public:
        explicit Line(Level l) : m_level(l) {}
        ~Line() { emit(m_level, m_text.str()); }
        template <class T>
        Line & operator<<(const T & v)
        {
            m_text << v;
            return *this;
        }

    private:
        Level m_level;
        std::ostringstream m_text;
    };
}  // namespace simplex::log

#define SB_LOG(lvl) \
    if (::simplex::log::lvl > ::simplex::log::level()) {} else ::simplex::log::Line(::simplex::log::lvl)
