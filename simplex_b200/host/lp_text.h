// Reader of the CPLEX-LP text subset accepted by the reference (reference src/Parser.h:13-16).
#pragma once

#include "model.h"

#include <string_view>

namespace simplex
{
    // Fills `lp` from the text of an .lp file. Returns true when the closing "End" was seen.
    // Throws std::string for errors the reference reports as "Exception: ..." and
    // `const char *` / std::logic_error subclasses for malformed text (see lp_text.cpp).
    bool run_parser_lp(std::string_view input, LinearProgram & lp);
}  // namespace simplex
