// `simplex [LP filename]` — the reference's command line (reference src/main.cpp:15-71) on top of
// the GPU loop: read the .lp file, presolve, solve (Phase I / Phase II on the device), write
// sol_test.xml into the working directory.
//
// Same messages and exit codes as the reference for the argument checks (1), for errors it
// reports as "Exception: ..." (0) and for the errors it does not catch (`const char *`,
// std::logic_error: the reference process dies by SIGABRT; so does this one, after saying why).
// Additions, all optional and after the file name:
//   --pricing first|dantzig   entering rule (default first = what the reference ships with)
//   --device N                CUDA device ordinal
//   --stats                   one JSON line with per-phase iteration counts and device times
#include "log.h"
#include "lp_text.h"
#include "presolve.h"
#include "solver.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace
{
    [[noreturn]] void die_like_an_uncaught_exception(const char * what)
    {
        std::cout.flush();
        std::cerr << "terminate: " << what << std::endl;
        std::abort();
    }
}  // namespace

int main(int argc, char ** argv)
{
    std::cout.precision(4);
    std::cout.setf(std::ios_base::showpoint);

    if (argc < 2)
    {
        std::cout << "You must provide a file name." << std::endl;
        return 1;
    }
    const std::string path(argv[1]);
    if (path.substr(path.find_last_of('.') + 1) != "lp")
    {
        std::cout << "Unrecognized file extension (must be .lp)." << std::endl;
        return 1;
    }
    std::ifstream file(path, std::ios::binary);
    if (!file)
    {
        std::cout << "Invalid input file name." << std::endl;
        return 1;
    }
    std::ostringstream whole;
    whole << file.rdbuf();
    std::string text = whole.str();
    file.close();
    std::cout << "Reading file done.\n";

    simplex::DeviceOptions chosen;
    bool stats = false;
    if (const char * e = std::getenv("SIMPLEX_PRICING"))
        if (!std::strcmp(e, "dantzig") || !std::strcmp(e, "most_negative")) chosen.pricing = simplex::Pricing::most_negative;
    for (int k = 2; k < argc; ++k)
    {
        const std::string arg(argv[k]);
        if (arg == "--pricing" && k + 1 < argc)
        {
            const std::string rule(argv[++k]);
            chosen.pricing = (rule == "dantzig" || rule == "most_negative") ? simplex::Pricing::most_negative
                                                                            : simplex::Pricing::first_eligible;
        }
        else if (arg == "--device" && k + 1 < argc)
            chosen.device = std::atoi(argv[++k]);
        else if (arg == "--stats")
            stats = true;
    }
    if (simplex::log::level() >= simplex::log::debug) chosen.trace_capacity = 1 << 16;

    try
    {
        simplex::LinearProgram lp;
        simplex::run_parser_lp(text, lp);
        text.clear();
        simplex::Presolve pre_solver(lp);
        pre_solver.run();
        simplex::Simplex main_solver(lp);
        main_solver.options() = chosen;
        main_solver.solve();
        main_solver.write("sol_test.xml");
        if (stats)
        {
            std::ostringstream js;
            js.precision(17);
            js << "{\"status\": " << static_cast<int>(main_solver.status()) << ", \"objective\": " << main_solver.objective_value()
               << ", \"calls\": [";
            bool first = true;
            for (const auto & c : main_solver.calls())
            {
                js << (first ? "" : ", ") << "{\"M\": " << c.M << ", \"N\": " << c.N << ", \"term\": " << c.term
                   << ", \"iterations\": " << c.iterations << ", \"pivots\": " << c.pivots << ", \"bound_flips\": " << c.bound_flips
                   << ", \"loop_seconds\": " << c.loop_seconds << ", \"resident\": " << (c.reused_resident_tableau ? "true" : "false")
                   << "}";
                first = false;
            }
            js << "]}";
            std::cout << "STATS " << js.str() << std::endl;
        }
    }
    catch (const std::string & msg)
    {
        std::cout << "Exception: " << msg << std::endl;
    }
    catch (const char * msg)
    {
        die_like_an_uncaught_exception(msg);
    }
    catch (const std::logic_error & e)
    {
        die_like_an_uncaught_exception(e.what());
    }

    std::cout << "Exiting program." << std::endl;
    return 0;
}
